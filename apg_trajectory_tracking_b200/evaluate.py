"""Batched closed-loop evaluation on table references.

Reference: ``QuadEvaluator.follow_trajectory("rand")`` / ``run_eval`` (scripts/evaluate_drone.py:81-194, 236-298), the
per-epoch evaluation of the quadrotor trainer (scripts/train_drone.py:205-216): a batch-1 CPU loop of up to 251 policy
calls per run.  Here N drones walk their reference tables in ONE kernel launch (csrc/eval_kernels.cu): window ->
QuadDataset.prepare_data -> policy -> dynamics step -> divergence / stability -> stop or reset.  CUDA only.
"""
import ctypes

import numpy as np
import torch

from . import _capi, rollout as R
from .ops import _p, _require_cuda, _stream


class TableEvaluator:
    """Owns the workspace for (spec, N); ``follow`` runs one closed-loop evaluation of N drones."""

    def __init__(self, spec: R.RolloutSpec, n_drones: int, device=None):
        if not torch.cuda.is_available():
            raise _capi.ApgError("no CUDA device: the evaluation rollout only runs on the GPU")
        if spec.system != "quad" or spec.net not in ("hutter_conv", "lstm"):
            raise _capi.ApgError("table evaluation is implemented for the quadrotor hutter conv and LSTM policies")
        self.spec, self.n = spec, int(n_drones)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _capi.lib()
        with torch.cuda.device(self.device):
            self.cfg = spec.config(self.n)
            ws = self.lib.apg_workspace_bytes(ctypes.byref(self.cfg))
            if ws == 0:
                raise _capi.ApgError("bad rollout spec for the evaluation rollout")
            self.workspace = torch.empty(ws + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_ptr = ctypes.c_void_p(self.workspace.data_ptr() + off)

    def follow(self, params_flat, tables, init_states=None, table_index=None, steps=251, thresh_div=1.0,
               thresh_stable=1.0, test_time=0, want=("states", "div", "actions"), h0c0=None):
        """tables (T,RL,9) rows [pos, euler, vel]; table_index (N,) int32 or None (drone i walks table i);
        init_states (N,12) or None (at rest on the first table point, like ``zero_reset(*initial_pos)``).
        Returns dict(states (N,steps+1,12), div (N,steps), actions (N,steps,4), n_steps (N,) int32); entries after a
        drone stopped are zero.  LSTM policy: ``h0c0`` (2,N,8) = every drone's hidden / cell state before its first
        policy call (required), and the dict also holds ``hc`` (2,N,8), the state after its last one."""
        _require_cuda(params_flat, tables, init_states, table_index)
        tables = tables if (tables.dtype == torch.float32 and tables.is_contiguous()) else tables.contiguous().float()
        n, dev = self.n, self.device
        if table_index is not None:
            table_index = table_index.to(torch.int32).contiguous()
            if table_index.numel() != n:
                raise ValueError("table_index must have one entry per drone")
        elif tables.shape[0] < n:
            raise ValueError("without table_index there must be one table per drone")
        if init_states is None:
            first = tables[:, 0, :3] if table_index is None else tables[table_index.long(), 0, :3]
            init_states = torch.zeros(n, 12, device=dev)
            init_states[:, :3] = first[:n]
        init_states = init_states.contiguous().float()
        out = {"n_steps": torch.zeros(n, dtype=torch.int32, device=dev)}
        if "states" in want:
            out["states"] = torch.zeros(n, steps + 1, 12, device=dev)
        if "div" in want:
            out["div"] = torch.zeros(n, steps, device=dev)
        if "actions" in want:
            out["actions"] = torch.zeros(n, steps, 4, device=dev)
        opt = lambda k: None if k not in out else _p(out[k])          # noqa: E731
        if self.spec.net == "lstm":
            if h0c0 is None or tuple(h0c0.shape) != (2, n, 8):
                raise ValueError("LSTM policy: h0c0 of shape (2, N, 8) is required")
            _require_cuda(h0c0)
            h0c0 = h0c0.contiguous().float()
            out["hc"] = torch.zeros(2, n, 8, device=dev)
            with torch.cuda.device(dev):
                _capi.check(self.lib.apg_eval_rollout_lstm(
                    ctypes.byref(self.cfg), _p(params_flat), _p(h0c0), _p(tables),
                    None if table_index is None else _p(table_index), int(tables.shape[0]), int(tables.shape[1]),
                    _p(init_states), int(steps), ctypes.c_float(thresh_div), ctypes.c_float(thresh_stable),
                    int(test_time), self._ws_ptr, opt("states"), opt("div"), opt("actions"), _p(out["n_steps"]),
                    _p(out["hc"]), _stream(tables)))
            return out
        with torch.cuda.device(dev):
            _capi.check(self.lib.apg_eval_rollout(
                ctypes.byref(self.cfg), _p(params_flat), _p(tables), None if table_index is None else _p(table_index),
                int(tables.shape[0]), int(tables.shape[1]), _p(init_states), int(steps), ctypes.c_float(thresh_div),
                ctypes.c_float(thresh_stable), int(test_time), self._ws_ptr, opt("states"), opt("div"), opt("actions"),
                _p(out["n_steps"]), _stream(tables)))
        return out


def selfplay_samples(out, tables, table_index=None, horizon=10, take_every_x=1000, thresh_div=1.0, thresh_stable=1.0,
                     test_time=0, action_counter=0):
    """The self-play feed of the evaluation: the raw (state, reference rows) pairs ``NetworkWrapper.predict_actions``
    hands to ``DroneDataset.get_and_add_eval_data(..., add_to_dataset=True)`` (controllers/network_wrapper.py:42-52,
    dataset.py:98-119) when the N runs of one ``TableEvaluator.follow`` call are counted one after the other like the
    reference's sequential runs: policy call number c (1-based, continuing from ``action_counter``) is kept when
    ``c % take_every_x == 0``.

    ``out``: the dict ``follow`` returned (needs "states" and "div"); tables / table_index / thresholds / test_time
    as passed to ``follow``.  The state a policy call saw is the recorded state, or - after a step that diverged or
    was unstable with ``test_time == 0`` - the reference state the drone was reset to (``[table[ci], 0, 0, 0]``,
    random_traj.py:89-92); its reference rows are ``Random.get_ref_traj`` at that step.  Pure index arithmetic and
    gathers on the tensors' device.  Returns (states (M,12), ref_states (M,horizon,9), action counter afterwards)."""
    states, div, n_steps = out["states"], out["div"], out["n_steps"].long()
    dev, h = states.device, int(horizon)
    rl = tables.shape[1]
    sel, counter = _kept_calls(n_steps, take_every_x, action_counter)
    if sel is None:
        return states.new_zeros(0, 12), states.new_zeros(0, h, 9), counter
    run, step = sel
    tab = tables if table_index is None else tables[table_index.long()]
    tab = tab[run].float()                                            # (M,RL,9)
    m = run.numel()
    ar = torch.arange(m, device=dev)
    last = max(rl - h, 0)
    ci = step.clamp(max=last)                                         # walking index before the call's window
    seen = states[run, step]
    if not test_time:
        prev = (step - 1).clamp(min=0)
        bad = (div[run, prev] > thresh_div) | ~((seen[:, 3].abs() < thresh_stable) & (seen[:, 4].abs() < thresh_stable))
        reset = torch.cat((tab[ar, ci], seen.new_zeros(m, 3)), dim=1)
        seen = torch.where(((step > 0) & bad)[:, None], reset, seen)
    end = ci >= rl - h
    start = torch.where(end, ci, ci + 1)
    nreal = torch.where(end, rl - ci, torch.full_like(ci, h))
    r = torch.arange(h, device=dev)[None, :]
    rows = torch.gather(tab, 1, (start[:, None] + r).clamp(max=rl - 1)[:, :, None].expand(m, h, 9))
    pad = torch.zeros_like(rows)
    pad[:, :, :3] = tab[:, -1:, :3]
    rows = torch.where((r < nreal[:, None])[:, :, None], rows, pad)
    return seen, rows, counter


def _kept_calls(n_steps, take_every_x, action_counter):
    """(run, step) of the policy calls whose running number (1-based, continuing from ``action_counter`` over the
    runs taken one after the other) is a multiple of ``take_every_x``; None when there is none"""
    total, x, ac = int(n_steps.sum()), int(take_every_x), int(action_counter)
    first = (ac // x + 1) * x
    if first > ac + total:
        return None, ac + total
    g = torch.arange(first, ac + total + 1, x, device=n_steps.device) - ac - 1     # 0-based index over all runs
    cum = torch.cumsum(n_steps, 0)
    run = torch.searchsorted(cum, g, right=True)
    return (run, g - (cum - n_steps)[run]), ac + total


def wing_selfplay_samples(out, targets, take_every_x=1000, action_counter=0):
    """The fixed-wing counterpart of ``selfplay_samples``: the raw (state, target) pairs
    ``FixedWingNetWrapper.predict_actions`` hands to ``get_and_add_eval_data(..., add_to_dataset=True)``
    (controllers/network_wrapper.py:81-90) for the N flights of one ``WingTargetEvaluator.fly`` call counted one after
    the other.  The state a call saw is the state ``env.step`` returned before it (the evaluator's local ``state`` is
    not refreshed by a reset); its target is the one current at that step: target k+1 from the step after the
    first one whose x position passed target k (evaluate_fixed_wing.py:93-110).
    ``out`` needs "states".  Returns (states (M,12), targets (M,3), action counter afterwards)."""
    states, n_steps = out["states"], out["n_steps"].long()
    dev, n, K = states.device, states.shape[0], targets.shape[1]
    sel, counter = _kept_calls(n_steps, take_every_x, action_counter)
    if sel is None:
        return states.new_zeros(0, 12), states.new_zeros(0, 3), counter
    run, step = sel
    targets = targets.to(dev, torch.float32)
    # step index at which target k was passed (first step >= the previous switch whose new x lies beyond it)
    steps = states.shape[1] - 1
    idx = torch.arange(steps, device=dev)[None, :]
    x_after = states[:, 1:, 0]
    start = torch.zeros(n, dtype=torch.long, device=dev)
    ti = torch.zeros(run.numel(), dtype=torch.long, device=dev)
    for k in range(K - 1):
        cond = (x_after > targets[:, k, 0:1]) & (idx >= start[:, None]) & (idx < n_steps[:, None])
        first = torch.where(cond.any(dim=1), cond.float().argmax(dim=1), torch.full_like(start, steps + 1))
        ti = ti + (step > first[run]).long()
        start = first + 1
    return states[run, step], targets[run, ti], counter


class WingTargetEvaluator:
    """Fixed wing: ``FixedWingEvaluator.fly_to_point`` / ``run_eval`` (scripts/evaluate_fixed_wing.py:46-178) for N
    flights at once.  ``spec``: ``RolloutSpec.wing_concurrent(h, dt_env, modified_params)`` of the EVALUATION
    environment; ``mean`` / ``std`` / ``dt_data``: the dataset's normalisation and delta_t (WingDataset)."""

    def __init__(self, spec: R.RolloutSpec, n_drones: int, mean, std, dt_data=None, device=None):
        if not torch.cuda.is_available():
            raise _capi.ApgError("no CUDA device: the evaluation rollout only runs on the GPU")
        if spec.system != "wing" or spec.net != "hutter_lin":
            raise _capi.ApgError("target evaluation is implemented for the fixed-wing hutter net")
        self.spec, self.n = spec, int(n_drones)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _capi.lib()
        self.mean = np.ascontiguousarray(torch.as_tensor(mean).detach().cpu().numpy(), dtype=np.float32)
        self.std = np.ascontiguousarray(torch.as_tensor(std).detach().cpu().numpy(), dtype=np.float32)
        if self.mean.shape != (12,) or self.std.shape != (12,):
            raise ValueError("mean / std must have 12 entries")
        self.dt_data = float(spec.dt if dt_data is None else dt_data)
        with torch.cuda.device(self.device):
            self.cfg = spec.config(self.n)
            ws = self.lib.apg_workspace_bytes(ctypes.byref(self.cfg))
            if ws == 0:
                raise _capi.ApgError("bad rollout spec for the evaluation rollout")
            self.workspace = torch.empty(ws + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_ptr = ctypes.c_void_p(self.workspace.data_ptr() + off)

    def fly(self, params_flat, targets, init_states=None, steps=1000, thresh_div=10.0, thresh_stable=0.8,
            test_time=0, want=("states", "div_linear", "actions")):
        """targets (N,K,3); init_states (N,12) or None (``zero_reset``: zeros with u = 11.5 m/s).  Returns
        dict(states (N,steps+1,12), div_linear (N,steps), actions (N,steps,4), n_steps (N,) int32,
        div_target_sum (N,), div_target_cnt (N,)): mean target error of flight i = sum / cnt (run_eval, :157)."""
        _require_cuda(params_flat, targets, init_states)
        n, dev = self.n, self.device
        targets = targets.contiguous().float()
        if targets.dim() != 3 or targets.shape[0] != n or targets.shape[2] != 3:
            raise ValueError(f"targets must be (N,K,3) with N = {n}")
        if init_states is None:
            init_states = torch.zeros(n, 12, device=dev)
            init_states[:, 3] = 11.5
        init_states = init_states.contiguous().float()
        out = {"n_steps": torch.zeros(n, dtype=torch.int32, device=dev),
               "div_target_sum": torch.zeros(n, device=dev), "div_target_cnt": torch.zeros(n, device=dev)}
        if "states" in want:
            out["states"] = torch.zeros(n, steps + 1, 12, device=dev)
        if "div_linear" in want:
            out["div_linear"] = torch.zeros(n, steps, device=dev)
        if "actions" in want:
            out["actions"] = torch.zeros(n, steps, 4, device=dev)
        opt = lambda k: None if k not in out else _p(out[k])          # noqa: E731
        with torch.cuda.device(dev):
            _capi.check(self.lib.apg_eval_fly_to_points(
                ctypes.byref(self.cfg), _p(params_flat), _p(targets), int(targets.shape[1]), _p(init_states),
                ctypes.c_void_p(self.mean.ctypes.data), ctypes.c_void_p(self.std.ctypes.data),
                ctypes.c_float(self.dt_data), int(steps), ctypes.c_float(thresh_div), ctypes.c_float(thresh_stable),
                int(test_time), self._ws_ptr, opt("states"), opt("div_linear"), opt("actions"), _p(out["n_steps"]),
                _p(out["div_target_sum"]), _p(out["div_target_cnt"]), _stream(targets)))
        return out


class CartpoleBalanceEvaluator:
    """Cartpole: ``Evaluator.evaluate_in_environment`` (scripts/evaluate_cartpole.py:78-262) for N carts at once.
    ``spec``: ``RolloutSpec.cartpole_concurrent(h, dt_env, modified_params)`` of the EVALUATION environment."""

    def __init__(self, spec: R.RolloutSpec, n_drones: int, device=None):
        if not torch.cuda.is_available():
            raise _capi.ApgError("no CUDA device: the evaluation rollout only runs on the GPU")
        if spec.system != "cartpole" or spec.net != "simple":
            raise _capi.ApgError("balance evaluation is implemented for the cartpole simple net")
        self.spec, self.n = spec, int(n_drones)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _capi.lib()
        with torch.cuda.device(self.device):
            self.cfg = spec.config(self.n)
            ws = self.lib.apg_workspace_bytes(ctypes.byref(self.cfg))
            if ws == 0:
                raise _capi.ApgError("bad rollout spec for the evaluation rollout")
            self.workspace = torch.empty(ws + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_ptr = ctypes.c_void_p(self.workspace.data_ptr() + off)

    def balance(self, params_flat, init_states=None, steps=250, thresh_div=0.21, burn_in_steps=50,
                want=("states", "actions")):
        """init_states (N,4) or None (the reference's start: everything zero, evaluate_cartpole.py:100-113 with
        ``center_at_x``).  Returns dict(states (N,steps,4) as returned by ``_step``, actions (N,steps), n_steps (N,)
        int32, success (N,) = n_steps - 1 (the reference's step index at the end of the run), mean_angle (N,) = mean
        |theta| after the burn-in (100 when there is none, :230), vel_sum (N,))."""
        _require_cuda(params_flat, init_states)
        n, dev = self.n, self.device
        if init_states is None:
            init_states = torch.zeros(n, 4, device=dev)
        init_states = init_states.contiguous().float()
        if tuple(init_states.shape) != (n, 4):
            raise ValueError(f"init_states must be (N,4) with N = {n}")
        out = {"n_steps": torch.zeros(n, dtype=torch.int32, device=dev)}
        ang_sum, ang_cnt = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        out["vel_sum"] = torch.zeros(n, device=dev)
        if "states" in want:
            out["states"] = torch.zeros(n, steps, 4, device=dev)
        if "actions" in want:
            out["actions"] = torch.zeros(n, steps, device=dev)
        opt = lambda k: None if k not in out else _p(out[k])          # noqa: E731
        with torch.cuda.device(dev):
            _capi.check(self.lib.apg_eval_cartpole(
                ctypes.byref(self.cfg), _p(params_flat), _p(init_states), int(steps), ctypes.c_float(thresh_div),
                int(burn_in_steps), self._ws_ptr, opt("states"), opt("actions"), _p(out["n_steps"]), _p(ang_sum),
                _p(ang_cnt), _p(out["vel_sum"]), _stream(init_states)))
        out["success"] = out["n_steps"] - 1
        out["mean_angle"] = torch.where(ang_cnt > 0, ang_sum / ang_cnt.clamp(min=1), torch.full_like(ang_sum, 100.0))
        return out


def cartpole_eval_statistics(n_steps, vel_sum):
    """``evaluate_in_environment``'s result dict (evaluate_cartpole.py:233-238) over the N runs of one ``balance``
    call: mean_vel (mean |x_dot| over ALL steps of all runs), mean_stable, std_stable (of ``success``)"""
    ns = n_steps.detach().cpu().double().numpy()
    succ = ns - 1
    return {"mean_vel": float(vel_sum.detach().cpu().double().sum() / max(ns.sum(), 1.0)),
            "mean_stable": float(succ.mean()), "std_stable": float(succ.std())}


def wing_eval_statistics(div_target_sum, div_target_cnt):
    """``FixedWingEvaluator.run_eval`` (evaluate_fixed_wing.py:157-178): mean and std over the flights of each
    flight's mean target error"""
    m = (div_target_sum / div_target_cnt.clamp(min=1)).detach().cpu().double().numpy()
    return float(m.mean()), float(m.std())


def eval_statistics(div, n_steps, thresh_div):
    """``QuadEvaluator.run_eval`` statistics (evaluate_drone.py:266-298) over the N runs of one ``follow`` call:
    (mean, std of the steps below the divergence threshold, mean, std of the tracking error of the runs that stayed
    below it for their whole length, mean, std of the tracking error of all runs)."""
    d = div.detach().cpu().double().numpy()
    ns = n_steps.detach().cpu().numpy()
    per_run = np.array([d[i, :ns[i]].mean() if ns[i] else 0.0 for i in range(len(ns))])
    stable = np.array([(d[i, :ns[i]] < thresh_div).sum() for i in range(len(ns))])
    full = per_run[stable == ns[-1]]             # max_steps_stable = len(reference_traj) of the LAST run (:276)
    nan = float("nan")
    return (stable.mean(), stable.std(), full.mean() if len(full) else nan, full.std() if len(full) else nan,
            per_run.mean(), per_run.std())
