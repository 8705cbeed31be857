"""Argument marshalling of the Python wrappers of the newer C entry points (prepare.py, evaluate.py,
quad_dynamics_trained.py) exercised WITHOUT a GPU: `_capi.lib()` is replaced by a stand-in whose entry points take
the same ctypes arguments and run the SAME kernel bodies compiled for the host (tests/hostcheck/*.cpp) on the CPU
tensors' memory.  A test double for the host-side plumbing only -- the kernels themselves are checked on the GPU by
tests/test_zz_new_paths_gpu.py."""
import contextlib
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from apg_trajectory_tracking_b200 import _capi, evaluate as EV, ops, prepare as PR, rollout as R, synthetic as SY
from apg_trajectory_tracking_b200.neural_control import dataset as DS
from apg_trajectory_tracking_b200.neural_control.dynamics import quad_dynamics_trained as QT
from oracle import apg_oracle as O
from tests.helpers import golden_params, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp, name):
    out = tmp / f"lib{name}.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I",
                           os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", f"{name}.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def _addr(x):
    """ctypes argument (c_void_p / int / None) -> integer address or None"""
    if x is None:
        return None
    return x.value if isinstance(x, ctypes.c_void_p) else int(x)


def _f(x):
    return x.value if isinstance(x, ctypes.c_float) else float(x)


class _HostLib:
    """same entry-point names and argument lists as libapg_b200.so for the input-side / evaluation / learnt calls"""

    def __init__(self, tmp):
        self.real = _capi.lib()
        self.prep, self.ev, self.ln = _build(tmp, "hostcheck_prep"), _build(tmp, "hostcheck_eval"), \
            _build(tmp, "hostcheck_learnt")
        self.dyn = _build(tmp, "hostcheck")
        self.keep = []

    def __getattr__(self, name):          # host-only queries go to the real library
        return getattr(self.real, name)

    @staticmethod
    def _vp(a):
        return ctypes.c_void_p(a)

    def apg_prepare_quad(self, states, ref, n, rows, in_state, cur, in_ref, ref_out, stream):
        self.prep.hc_prepare_quad(*[self._vp(_addr(x)) for x in (states, ref)], n, rows,
                                  *[self._vp(_addr(x)) for x in (in_state, cur, in_ref, ref_out)])
        return 0

    def apg_prepare_wing(self, states, targets, mean, std, dt, h, n, in_state, cur, in_ref, ref_out, stream):
        self.prep.hc_prepare_wing(*[self._vp(_addr(x)) for x in (states, targets, mean, std)], ctypes.c_float(_f(dt)),
                                  h, n, *[self._vp(_addr(x)) for x in (in_state, cur, in_ref, ref_out)])
        return 0

    def apg_sample_windows(self, traj, rows, cols, L, stride, n, states, refs, stream):
        if n > 0 and (n - 1) * stride + L > rows - 1:
            return -1
        self.prep.hc_sample_windows(self._vp(_addr(traj)), cols, L, stride, n, self._vp(_addr(states)),
                                    self._vp(_addr(refs)))
        return 0

    def apg_poly_reference(self, coef, n, rows, t_first, dt, out, stream):
        self.prep.hc_poly_reference(self._vp(_addr(coef)), n, rows, ctypes.c_float(_f(t_first)),
                                    ctypes.c_float(_f(dt)), self._vp(_addr(out)))
        return 0

    def apg_reference_table(self, traj, rows, cols, nth, speed, zoff, table_rows, out, stream):
        if table_rows > 0 and (table_rows - 1) * nth > rows - 1:
            return -1
        self.prep.hc_reference_table(self._vp(_addr(traj)), cols, nth, ctypes.c_float(_f(speed)),
                                     ctypes.c_float(_f(zoff)), table_rows, self._vp(_addr(out)))
        return 0

    def apg_polynomial_points(self, coef, degree, rot, start, n, x_start, x_range, dist, hover, max_rows, out, ref_len,
                              stream):
        self.prep.hc_polynomial_points(self._vp(_addr(coef)), degree, self._vp(_addr(rot)), self._vp(_addr(start)), n,
                                       ctypes.c_double(x_start.value), ctypes.c_double(x_range.value),
                                       ctypes.c_double(dist.value), hover, max_rows, self._vp(_addr(out)),
                                       self._vp(_addr(ref_len)))
        return 0

    def apg_dynamics_step(self, system, phys, state, action, dt, n, out, stream):
        fn = getattr(self.dyn, "hc_step_%s_f32" % ("quad", "wing", "cartpole")[system])
        fn(self._vp(_addr(state)), self._vp(_addr(action)), ctypes.c_float(_f(dt)), self._vp(_addr(phys)),
           self._vp(_addr(out)), n)
        return 0

    def apg_eval_rollout(self, cfg, params, tables, index, n_tables, rows, init, steps, tdiv, tstab, test_time, ws,
                         states, div, actions, n_steps, stream):
        c = cfg._obj
        n = c.n_drones
        tmp = {k: np.zeros(s, np.float32) for k, s in (("states", (n, steps + 1, 12)), ("div", (n, steps)),
                                                        ("actions", (n, steps, 4)))}
        self.keep.append(tmp)
        ptr = lambda given, k: self._vp(_addr(given) if given is not None else tmp[k].ctypes.data)   # noqa: E731
        phys = (ctypes.c_float * 48)(*c.phys)
        self.ev.hc_eval_rollout(self._vp(_addr(params)), c.horizon, c.out_dim, self._vp(_addr(tables)),
                                self._vp(_addr(index)), rows, self._vp(_addr(init)), n, steps, ctypes.c_float(c.dt),
                                phys, ctypes.c_float(_f(tdiv)), ctypes.c_float(_f(tstab)), int(test_time),
                                ptr(states, "states"), ptr(div, "div"), ptr(actions, "actions"),
                                self._vp(_addr(n_steps)))
        return 0

    def apg_eval_fly_to_points(self, cfg, params, targets, K, init, mean, std, dt_data, steps, tdiv, tstab, test_time,
                               ws, states, div, actions, n_steps, dts, dtc, stream):
        c = cfg._obj
        n = c.n_drones
        tmp = {k: np.zeros(s, np.float32) for k, s in (("states", (n, steps + 1, 12)), ("div", (n, steps)),
                                                        ("actions", (n, steps, 4)))}
        self.keep.append(tmp)
        ptr = lambda given, k: self._vp(_addr(given) if given is not None else tmp[k].ctypes.data)   # noqa: E731
        phys = (ctypes.c_float * 48)(*c.phys)
        self.ev.hc_eval_wing(self._vp(_addr(params)), c.horizon, self._vp(_addr(targets)), K, self._vp(_addr(init)), n,
                             self._vp(_addr(mean)), self._vp(_addr(std)), ctypes.c_float(_f(dt_data)),
                             ctypes.c_float(c.dt), phys, steps, ctypes.c_float(_f(tdiv)), ctypes.c_float(_f(tstab)),
                             int(test_time), ptr(states, "states"), ptr(div, "div"), ptr(actions, "actions"),
                             self._vp(_addr(n_steps)), self._vp(_addr(dts)), self._vp(_addr(dtc)))
        return 0

    def apg_eval_cartpole(self, cfg, params, init, steps, tdiv, burn, ws, states, actions, n_steps, asum, acnt, vsum,
                          stream):
        c = cfg._obj
        n = c.n_drones
        tmp = {"states": np.zeros((n, steps, 4), np.float32), "actions": np.zeros((n, steps), np.float32)}
        self.keep.append(tmp)
        ptr = lambda given, k: self._vp(_addr(given) if given is not None else tmp[k].ctypes.data)   # noqa: E731
        phys = (ctypes.c_float * 48)(*c.phys)
        self.ev.hc_eval_cartpole(self._vp(_addr(params)), c.horizon, self._vp(_addr(init)), n, ctypes.c_float(c.dt),
                                 phys, steps, ctypes.c_float(_f(tdiv)), int(burn), ptr(states, "states"),
                                 ptr(actions, "actions"), *[self._vp(_addr(x)) for x in (n_steps, asum, acnt, vsum)])
        return 0

    def apg_learnt_step(self, system, params, phys, state, action, dt, n, out, stream):
        if system == 0:
            self.ln.hc_learnt_fwd_f32(*[self._vp(_addr(x)) for x in (params, phys, state, action)],
                                      ctypes.c_float(_f(dt)), n, self._vp(_addr(out)))
        else:
            self.ln.hc_learnt_wing_fwd_f32(*[self._vp(_addr(x)) for x in (params, state, action)],
                                           ctypes.c_float(_f(dt)), n, self._vp(_addr(out)))
        return 0

    def apg_learnt_step_adjoint(self, system, params, phys, state, action, dt, n, g, gs, ga, gp, ws, stream):
        if system == 0:
            self.ln.hc_learnt_adj_f32(*[self._vp(_addr(x)) for x in (params, phys, state, action)],
                                      ctypes.c_float(_f(dt)), n, *[self._vp(_addr(x)) for x in (g, gs, ga, gp)])
        else:
            self.ln.hc_learnt_wing_adj_f32(*[self._vp(_addr(x)) for x in (params, state, action)],
                                           ctypes.c_float(_f(dt)), n, *[self._vp(_addr(x)) for x in (g, gs, ga, gp)])
        return 0


@pytest.fixture
def hostlib(monkeypatch, tmp_path_factory):
    lib = _HostLib(tmp_path_factory.mktemp("hostlibs"))
    monkeypatch.setattr(_capi, "lib", lambda: lib)
    for mod in (ops, PR, EV, QT):
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None, raising=False)
        monkeypatch.setattr(mod, "_stream", lambda t: None, raising=False)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    return lib


def test_prepare_wrappers(hostlib):
    g = load_golden("prep_data.npz")
    s, r = torch.tensor(g["quad_raw_states"]).float(), torch.tensor(g["quad_raw_refs"]).float()
    out = PR.prepare_quad(s.clone(), r.clone())
    assert torch.equal(out["cur"], torch.tensor(g["quad_states"])) and torch.equal(out["ref"], torch.tensor(g["quad_ref"]))
    assert torch.equal(out["in_ref"], torch.tensor(g["quad_in_ref"]))
    assert torch.allclose(out["in_state"], torch.tensor(g["quad_in_state"]), atol=1e-6)
    s2, r2 = s.clone(), r.clone()
    o2 = PR.prepare_quad(s2, r2, in_place=True)
    assert o2["cur"] is s2 and o2["ref"] is r2 and torch.equal(s2, out["cur"]) and torch.equal(r2, out["ref"])
    only = PR.prepare_quad(s.clone(), r.clone(), want=("in_ref",))
    assert set(only) == {"in_ref"} and torch.equal(only["in_ref"], out["in_ref"])
    # slices of larger buffers as outputs (what step_host passes)
    big = {"cur": torch.zeros(20, 12), "ref": torch.zeros(20, 10, 9), "in_ref": torch.zeros(20, 10, 9),
           "in_state": torch.zeros(20, 15)}
    big["cur"][8:14], big["ref"][8:14] = s, r
    sl = {k: v[8:14] for k, v in big.items()}
    PR.prepare_quad(sl["cur"], sl["ref"], out=sl)
    assert torch.equal(big["in_ref"][8:14], out["in_ref"]) and torch.equal(big["cur"][8:14], out["cur"])
    assert float(big["in_ref"][:8].abs().sum() + big["in_ref"][14:].abs().sum()) == 0.0
    w = PR.prepare_wing(torch.tensor(g["wing_raw_states"]).float(), torch.tensor(g["wing_targets"]).float(),
                        g["wing_mean"], g["wing_std"], float(g["wing_dt"]), int(g["wing_h"]))
    for k, name in (("in_state", "wing_in_state"), ("in_ref", "wing_in_ref"), ("ref", "wing_ref"), ("cur", "wing_states")):
        assert torch.allclose(w[k], torch.tensor(g[name]), atol=2e-6 * float(np.abs(g[name]).max())), k
    traj = torch.randn(101, 10)
    st, rf = PR.sample_windows(traj, 4, 10, 20)
    assert torch.equal(st[:, :9], traj[[0, 20, 40, 60], :9]) and torch.equal(rf[2, 3], traj[44, :9])
    with pytest.raises(_capi.ApgError):
        PR.sample_windows(traj, 10, 10, 20)
    c = SY.quad_case(5, 10, 0.1, seed=1)
    coef = torch.zeros(5, 3, 6)
    coef[:, :, 1] = 1.0
    rows = PR.poly_reference(coef, 10, 0.1)
    assert torch.allclose(rows[:, :, 0], (torch.arange(10) + 1)[None] * 0.1) and float(rows[:, :, 6].min()) == 1.0
    assert c["ref"].shape == rows.shape


def test_table_evaluator_wrapper(hostlib):
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    name = "fast_stop"
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    ev = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(int(h), dt), 1, "cpu")
    out = ev.follow(R.flatten_params(params), torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None],
                    steps=int(steps), thresh_div=tdiv, thresh_stable=tstab, test_time=int(test_time))
    taken = len(g[f"{name}_div"])
    assert int(out["n_steps"][0]) == taken                 # default init state: at rest on the first table point
    assert np.abs(out["states"][0, :taken + 1].numpy() - g[f"{name}_states"]).max() <= 2e-5
    assert np.abs(out["div"][0, :taken].numpy() - g[f"{name}_div"]).max() <= 2e-5
    # table index + explicit initial states + a subset of outputs
    tabs = torch.tensor(np.stack([g["gentle_table"][:100], g["tight_table"][:100]]), dtype=torch.float32)
    index = torch.tensor([1, 0, 1], dtype=torch.int32)
    init = torch.zeros(3, 12)
    init[:, :3] = tabs[index.long(), 0, :3] + 0.02
    ev3 = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(10, 0.1), 3, "cpu")
    out = ev3.follow(R.flatten_params(params), tabs, init_states=init, table_index=index, steps=30, thresh_div=0.6,
                     thresh_stable=0.5, test_time=1, want=("div",))
    want = O.eval_follow_tables(params, tabs[index.long()], init, 30, 10, 0.1, 0.6, 0.5, 1)
    assert set(out) == {"div", "n_steps"} and torch.equal(out["n_steps"].long(), want["n_steps"])
    assert float((out["div"] - want["div"]).abs().max()) <= 2e-5
    stats = EV.eval_statistics(out["div"], out["n_steps"], 0.6)
    ostats = O.eval_statistics(want["div"], want["n_steps"], 0.6)
    assert np.allclose(np.array(stats), np.array(ostats), rtol=1e-4, equal_nan=True)


def test_learnt_dynamics_wrapper_and_trainer_step(hostlib):
    from apg_trajectory_tracking_b200.scripts.train_base import TrainBase
    g = load_golden("learnt_dyn.npz")
    d = QT.LearntDynamics({"rotational_drag": [float(x) for x in g["b_rot_drag"]]})
    with torch.no_grad():
        for i, (_, p) in enumerate(d.named_parameters()):
            if i not in (1, 2, 3):
                p.copy_(torch.tensor(g[f"b_param_{i}"]))
    s = torch.tensor(g["b_state"]).requires_grad_(True)
    a = torch.tensor(g["b_action"]).requires_grad_(True)
    out = d(s, a, float(g["b_dt"]))
    assert float((out.detach() - torch.tensor(g["b_out"])).abs().max()) <= 5e-6 * float(np.abs(g["b_out"]).max())
    (out * torch.tensor(g["b_cot"])).sum().backward()
    assert torch.allclose(s.grad, torch.tensor(g["b_gstate"]), atol=2e-5 * float(np.abs(g["b_gstate"]).max()))
    for i, (name, p) in enumerate(d.named_parameters()):
        want = torch.tensor(g[f"b_gparam_{i}"])
        if i != 1:
            assert float((p.grad - want).abs().max()) <= 5e-5 * max(float(want.abs().max()), 0.1), name

    class _Target:                         # stands in for the CUDA FlightmareDynamics op on this CPU double
        def __call__(self, state, action, dt):
            return torch.tensor(g["b_dyn_target"])
    tr = TrainBase(d, _Target(), delta_t=float(g["b_dt"]), learning_rate_dynamics=1e-3)
    tr.init_dynamics_optimizer(l2_lambda=0.01)
    loss = tr.train_dynamics_model(torch.tensor(g["b_state"]), torch.tensor(g["b_action"])[:, None, :])
    assert abs(float(loss) - float(g["b_dyn_loss"])) <= 1e-5 * abs(float(g["b_dyn_loss"]))


def test_wing_target_evaluator_wrapper(hostlib):
    from tests.test_oracle_golden import wing_eval_case
    g = load_golden("eval_wing.npz")
    for name in ("two_targets", "tight_reset"):
        params, targets, init, h, dt_data, dt_env, steps, test_time, tdiv, tstab = wing_eval_case(g, name)
        ev = EV.WingTargetEvaluator(R.RolloutSpec.wing_concurrent(h, dt_env), 1, g["mean"], g["std"], dt_data, "cpu")
        out = ev.fly(R.flatten_params(params), targets, steps=steps, thresh_div=tdiv, thresh_stable=tstab,
                     test_time=test_time)
        traj, dl, dtg = g[f"{name}_traj"], g[f"{name}_div_linear"], g[f"{name}_div_target"]
        taken = len(dl)
        assert int(out["n_steps"][0]) == taken and int(out["div_target_cnt"][0]) == len(dtg)
        assert np.abs(out["states"][0, 1:taken + 1].numpy() - traj[:, :12]).max() <= 5e-5 * np.abs(traj[:, :12]).max()
        assert abs(float(out["div_target_sum"][0]) - dtg.sum()) <= 2e-4 * max(dtg.sum(), 1.0)
        m, sd = EV.wing_eval_statistics(out["div_target_sum"], out["div_target_cnt"])
        assert abs(m - dtg.mean()) <= 2e-4 * max(dtg.mean(), 1.0) and sd == 0.0
    slim = ev.fly(R.flatten_params(params), targets, steps=20, want=())
    assert set(slim) == {"n_steps", "div_target_sum", "div_target_cnt"} and int(slim["n_steps"][0]) == 20


def test_learnt_wing_dynamics_wrapper(hostlib):
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import LearntFixedWingDynamics
    g = load_golden("learnt_dyn.npz")
    d = LearntFixedWingDynamics()
    with torch.no_grad():
        for i, (_, p) in enumerate(d.named_parameters()):
            p.copy_(torch.tensor(g[f"wb_param_{i}"]))
    s = torch.tensor(g["wb_state"]).requires_grad_(True)
    a = torch.tensor(g["wb_action"]).requires_grad_(True)
    out = d(s, a, float(g["wb_dt"]))
    assert float((out.detach() - torch.tensor(g["wb_out"])).abs().max()) <= 5e-6 * float(np.abs(g["wb_out"]).max())
    (out * torch.tensor(g["wb_cot"])).sum().backward()
    assert torch.allclose(s.grad, torch.tensor(g["wb_gstate"]), atol=5e-5 * float(np.abs(g["wb_gstate"]).max()))
    assert torch.allclose(a.grad, torch.tensor(g["wb_gaction"]), atol=5e-5 * float(np.abs(g["wb_gaction"]).max()))
    for i, (name, p) in enumerate(d.named_parameters()):
        want = torch.tensor(g[f"wb_gparam_{i}"])
        assert float((p.grad - want).abs().max()) <= 2e-4 * max(float(want.abs().max()), 1e-2), name


def test_cartpole_balance_evaluator_wrapper(hostlib):
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    flat = R.flatten_params(params)
    for name in ("tilted", "falls"):
        steps, tdiv, burn = g[f"{name}_cfg"]
        ev = EV.CartpoleBalanceEvaluator(R.RolloutSpec.cartpole_concurrent(10, 0.05), 1, "cpu")
        out = ev.balance(flat, torch.tensor(g[f"{name}_init"], dtype=torch.float32)[None], steps=int(steps),
                         thresh_div=float(tdiv), burn_in_steps=int(burn))
        want = g[f"{name}_states"]
        taken = len(want)
        assert int(out["n_steps"][0]) == taken and int(out["success"][0]) == int(g[f"{name}_success"][0])
        assert np.abs(out["states"][0, :taken].numpy() - want).max() <= 2e-5 * max(np.abs(want).max(), 1.0)
        late = np.abs(want[int(burn) + 1:, 2])
        assert abs(float(out["mean_angle"][0]) - (late.mean() if len(late) else 100.0)) <= 1e-5
        st = EV.cartpole_eval_statistics(out["n_steps"], out["vel_sum"])
        assert abs(st["mean_vel"] - g[f"{name}_vel"].mean()) <= 1e-4 * max(g[f"{name}_vel"].mean(), 1.0)
        assert st["mean_stable"] == taken - 1 and st["std_stable"] == 0.0
    # the reference's own start (all zeros) for a batch, slim outputs
    ev4 = EV.CartpoleBalanceEvaluator(R.RolloutSpec.cartpole_concurrent(10, 0.05), 4, "cpu")
    slim = ev4.balance(flat, steps=30, want=())
    assert set(slim) == {"n_steps", "vel_sum", "success", "mean_angle"} and slim["n_steps"].tolist() == [30] * 4
    with pytest.raises(ValueError):
        ev4.balance(flat, torch.zeros(3, 4))


def test_selfplay_feed_matches_reference_dataset(hostlib, monkeypatch):
    """evaluation -> self-play slots of the dataset (network_wrapper.py:42-52, dataset.py:98-119): the three runs of
    tests/golden/eval_selfplay.npz taken one after the other with ONE action counter, through the TableEvaluator
    wrapper (host-compiled kernel logic) and evaluate.selfplay_samples, into a DeviceQuadDataset ring"""
    from apg_trajectory_tracking_b200 import device_data as DD
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    g = load_golden("eval_selfplay.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    h, dt, take, n_sampled, n_slots = [float(v) for v in g["cfg"]]
    h, take, n_sampled, n_slots = int(h), int(take), int(n_sampled), int(n_slots)
    tot = n_sampled + n_slots
    ds = DD.DeviceQuadDataset(torch.zeros(tot, 12), torch.zeros(tot, h, 9), "cpu", num_self_play=n_slots)
    assert ds.num_sampled_states == n_sampled and ds.get_eval_index() == n_sampled
    counter, kept_s, kept_r = 0, [], []
    for name in [str(v) for v in g["run_names"]]:
        steps, tdiv, tstab = g[f"{name}_cfg"]
        tab = torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None]
        ev = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(h, dt), 1, "cpu")
        out = ev.follow(R.flatten_params(params), tab, init_states=torch.tensor(g[f"{name}_states"][:1],
                                                                               dtype=torch.float32),
                        steps=int(steps), thresh_div=tdiv, thresh_stable=tstab, test_time=0)
        assert int(out["n_steps"][0]) == len(g[f"{name}_div"])
        s, r, counter = EV.selfplay_samples(out, tab, None, h, take, tdiv, tstab, 0, counter)
        ds.add_self_play(s, r)
        kept_s.append(s)
        kept_r.append(r)
    kept_s, kept_r = torch.cat(kept_s).numpy(), torch.cat(kept_r).numpy()
    assert counter == int(g["action_counter"][0]) and ds.eval_counter == int(g["eval_counter"][0])
    assert kept_s.shape == g["kept_states"].shape
    assert np.abs(kept_s - g["kept_states"]).max() <= 2e-5 and np.abs(kept_r - g["kept_refs"]).max() <= 1e-6
    # one kept call saw the state a reset left behind: exactly a table row with zero body rates
    assert any(np.all(k[9:] == 0) and np.abs(g["b_table"] - k[:9]).max(axis=1).min() < 1e-6 for k in kept_s)
    # the ring: what the reference's dataset holds afterwards (its `states` have the position zeroed and its
    # `ref_states` are relative to it, dataset.py:170-175 - prepare the raw ring rows the same way to compare)
    prep = PR.prepare_quad(ds.states, ds.ref_states, want=("cur", "ref"))
    assert np.abs(prep["cur"].numpy() - g["ds_states"]).max() <= 2e-5
    assert np.abs(prep["ref"].numpy() - g["ds_ref_states"]).max() <= 2e-5
    assert float(ds.states[:n_sampled].abs().sum()) == 0.0              # the sampled rows are untouched


def test_selfplay_samples_batched_selection_matches_sequential_runs():
    """N runs of one follow call = N sequential runs of the reference: selection, reset states and windows against
    the oracle's recorded policy inputs (no kernels involved: selfplay_samples is index arithmetic on the outputs)"""
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    tabs = torch.tensor(np.stack([g["fast_reset_table"][:60], g["tight_table"][:60], g["gentle_table"][:60]]),
                        dtype=torch.float32)
    index = torch.tensor([0, 1, 2, 1, 0], dtype=torch.int32)
    gen = torch.Generator().manual_seed(0)
    init = torch.zeros(5, 12)
    init[:, :3] = tabs[index.long(), 0, :3] + 0.05 * torch.randn(5, 3, generator=gen)
    for test_time, steps in ((0, 70), (1, 40)):
        out = O.eval_follow_tables(params, tabs[index.long()], init, steps, 10, 0.1, 0.4, 0.3, test_time,
                                   record_policy_inputs=True)
        assert test_time or int((out["div"] > 0.4).sum()) > 0           # there are resets in the batch
        for take, ac in ((3, 0), (11, 7), (1000, 0)):
            kept, after = O.selfplay_kept_calls(out["n_steps"], take, ac)
            s, r, counter = EV.selfplay_samples(out, tabs, index, 10, take, 0.4, 0.3, test_time, ac)
            assert counter == after and s.shape[0] == len(kept)
            if kept:
                want_s = torch.stack([out["policy_states"][j, i] for j, i in kept])
                want_r = torch.stack([out["windows"][j, i] for j, i in kept])
                assert torch.equal(s, want_s) and torch.equal(r, want_r)


def test_reference_table_wrapper(hostlib):
    g = load_golden("ref_table.npz")
    for name in ("a", "b", "c"):
        dt, speed = [float(v) for v in g[f"{name}_cfg"]]
        tab = PR.reference_table(torch.tensor(g[f"{name}_raw"]), dt, speed)
        want = g[f"{name}_table"].copy()
        want[:, 2] += 3                                                  # Random.__init__ (random_traj.py:35)
        assert tuple(tab.shape) == want.shape and np.abs(tab.numpy() - want).max() <= 3e-6
    with pytest.raises(ValueError):
        PR.reference_table(torch.zeros(50, 12), 0.1, 0.37)               # dt / 0.01 * speed not an integer
    with pytest.raises(ValueError):
        PR.reference_table(torch.zeros(50, 9), 0.1, 0.4)


def test_polynomial_points_wrapper(hostlib):
    g = load_golden("poly_traj.npz")
    names = [str(v) for v in g["case_names"]]
    for name in names:
        x_range, degree, mdd, h, hover = g[f"{name}_cfg"]
        pts, ref_len = PR.polynomial_points(torch.tensor(g[f"{name}_coef"])[None], torch.tensor(g[f"{name}_rot"])[None],
                                            torch.tensor(g[f"{name}_start"])[None], x_range=x_range,
                                            max_drone_dist=mdd, horizon=int(h), hover_steps=int(hover))
        want = g[f"{name}_points"]
        assert int(ref_len[0]) == len(want) and pts.shape[1] >= len(want)
        assert np.array_equal(pts[0, :len(want)].numpy(), want.astype(np.float32))
    # two trajectories of the same degree in one call, no shift; too small a buffer is reported
    coef = torch.tensor(np.stack([g["a_coef"], g["b_coef"]]))
    rot = torch.tensor(np.stack([g["a_rot"], g["b_rot"]]))
    pts, ref_len = PR.polynomial_points(coef, rot, None, x_range=6, max_drone_dist=0.5, horizon=10, hover_steps=5)
    assert int(ref_len[1]) == len(g["b_points"])
    assert np.allclose(pts[1, :int(ref_len[1])].numpy() - pts[1, 0].numpy(), g["b_points"] - g["b_points"][0], atol=1e-5)
    with pytest.raises(ValueError):
        PR.polynomial_points(coef, rot, None, x_range=6, max_drone_dist=0.5, horizon=10, hover_steps=5, max_rows=20)


def _single_drone_mirrors_on_cpu(monkeypatch):
    """the batch-1 mirrors (controllers / environments) with the host-compiled dynamics and the policy on the CPU"""
    from apg_trajectory_tracking_b200.neural_control import environments as ENV
    from apg_trajectory_tracking_b200.neural_control.controllers import network_wrapper as NW
    from apg_trajectory_tracking_b200.neural_control.models import hutter_model as HM, simple_model as SM
    monkeypatch.setattr(ENV, "compute_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(NW, "_device_of", lambda net: torch.device("cpu"))
    for mod in (HM, SM):
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None)
    return NW


def test_single_drone_mirrors_reproduce_reference_quad_evaluation(hostlib, monkeypatch):
    """NetworkWrapper + QuadDataset + QuadRotorEnvBase + FlightmareDynamics mirrors driven by the reference's
    follow_trajectory loop (scripts/evaluate_drone.py:136-188), against its golden run with resets"""
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.drone_env import QuadRotorEnvBase
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    g = load_golden("eval_rand.npz")
    name = "fast_reset"
    steps, test_time, tdiv, tstab, h, dt = [float(v) for v in g[f"{name}_cfg"]]
    steps, h = int(steps), int(h)
    net = Net(15, h, 9, 4 * h)
    with torch.no_grad():
        for p, q in zip(net.parameters(), golden_params(load_golden("conc_quad_kat4.npz"))):
            p.copy_(q)
    table = g[f"{name}_table"]
    ds = DS.QuadDataset(np.zeros((4, 12)), np.zeros((4, h, 9)), self_play=1.0)
    ctrl = NW.NetworkWrapper(net, ds, horizon=h, dt=dt, take_every_x=5)
    env = QuadRotorEnvBase(FlightmareDynamics(), dt)
    state = env.zero_reset(*table[0, :3])
    ci, traj, divs = torch.zeros(1, dtype=torch.long), [state], []
    for i in range(steps):
        rows, ci = O.eval_window(torch.tensor(table)[None], ci, h)
        action = ctrl.predict_actions(state, rows[0].numpy().copy())
        state, stable = env.step(action[0], thresh=tstab)
        traj.append(state)
        div = float(np.linalg.norm(table[int(ci), :3] - state[:3]))
        divs.append(div)
        if div > tdiv or not stable:
            state = np.hstack((table[int(ci)], np.zeros(3)))
            env._state.from_np(state)
        if i >= len(table):
            break
    want = g[f"{name}_states"]
    assert len(traj) == len(want) and np.abs(np.array(traj) - want).max() <= 5e-5
    assert np.abs(np.array(divs) - g[f"{name}_div"]).max() <= 5e-5 and (np.array(divs) > tdiv).sum() > 0
    assert ctrl.action_counter == steps and ds.eval_counter == steps // 5       # the self-play feed ran


def test_single_drone_mirrors_reproduce_reference_cartpole_evaluation(hostlib, monkeypatch):
    """CartpoleWrapper + CartPoleEnv + CartpoleDynamics mirrors in the loop of evaluate_in_environment
    (scripts/evaluate_cartpole.py:121-228), including the aliasing side effect on the environment state"""
    from apg_trajectory_tracking_b200.neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.cartpole_env import CartPoleEnv
    from apg_trajectory_tracking_b200.neural_control.models.simple_model import Net
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    g = load_golden("eval_cartpole.npz")
    net = Net(4, 10)
    with torch.no_grad():
        for i, p in enumerate(net.parameters()):
            p.copy_(torch.tensor(g[f"param_{i}"]))
    for name in ("tilted", "falls"):
        steps, tdiv, burn = g[f"{name}_cfg"]
        env = CartPoleEnv(CartpoleDynamics(), 0.05, thresh_div=float(tdiv))
        ctrl = NW.CartpoleWrapper(net, horizon=10, action_dim=1)
        env.state = np.array(g[f"{name}_init"], dtype=np.float64)
        new_state, log = env.state, []
        for i in range(int(steps)):
            with torch.no_grad():
                action_seq = ctrl.predict_actions(new_state, None)
            new_state = env._step(action_seq[:, 0], is_torch=True)
            log.append(new_state.copy())
            if not env.is_upright():
                break
        want = g[f"{name}_states"]
        assert len(log) == len(want) and np.abs(np.array(log) - want).max() <= 2e-5


def test_single_drone_mirrors_reproduce_reference_wing_flight(hostlib, monkeypatch):
    """FixedWingNetWrapper + WingDataset + SimpleWingEnv + FixedWingDynamics mirrors in the loop of fly_to_point
    (scripts/evaluate_fixed_wing.py:66-92) on the part of a golden flight before the first target switch"""
    from tests.test_oracle_golden import wing_eval_case
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.wing_env import SimpleWingEnv
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    g = load_golden("eval_wing.npz")
    params, targets, init, h, dt_data, dt_env, steps, test_time, tdiv, tstab = wing_eval_case(g, "one_target")
    net = Net(9, 1, 3, 4 * h, conv=False)
    with torch.no_grad():
        for p, q in zip(net.parameters(), params):
            p.copy_(q)
    ds = DS.WingDataset(np.zeros((2, 12)), np.ones((2, 3)), mean=g["mean"], std=g["std"], delta_t=dt_data, horizon=h)
    ctrl = NW.FixedWingNetWrapper(net, ds, horizon=h)
    env = SimpleWingEnv(FixedWingDynamics(), dt_env)
    env.zero_reset()
    state, target = env._state, targets[0, 0].numpy()
    traj = g["one_target_traj"]
    n_check = 40
    assert traj[:n_check, 0].max() < target[0]                           # still before the target: no switching
    for i in range(n_check):
        action = ctrl.predict_actions(state, target)
        assert np.abs(action[0] - traj[i, 12:]).max() <= 1e-4
        state, stable = env.step(action[0], thresh_stable=tstab)
        assert stable and np.abs(state - traj[i, :12]).max() <= 2e-4 * np.abs(traj[:, :12]).max()
    assert ctrl.action_counter == n_check


def test_evaluator_mirrors_run_eval_batched(hostlib, monkeypatch):
    """scripts.evaluate_drone.QuadEvaluator / evaluate_fixed_wing.FixedWingEvaluator / evaluate_cartpole.Evaluator:
    the reference's constructors and run_eval / evaluate_in_environment results with all runs in one launch"""
    from apg_trajectory_tracking_b200.scripts import evaluate_cartpole as EC, evaluate_drone as ED, \
        evaluate_fixed_wing as EF
    from apg_trajectory_tracking_b200.neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.cartpole_env import CartPoleEnv
    from apg_trajectory_tracking_b200.neural_control.environments.drone_env import QuadRotorEnvBase
    from apg_trajectory_tracking_b200.neural_control.environments.wing_env import SimpleWingEnv
    from apg_trajectory_tracking_b200.neural_control.models import hutter_model as HM, simple_model as SM
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    # ---- quadrotor: two golden runs as one batch
    g = load_golden("eval_rand.npz")
    h, dt = 10, 0.1
    net = HM.Net(15, h, 9, 4 * h)
    with torch.no_grad():
        for p, q in zip(net.parameters(), golden_params(load_golden("conc_quad_kat4.npz"))):
            p.copy_(q)
    ds = DS.QuadDataset(np.zeros((6, 12)), np.zeros((6, h, 9)), self_play=1.0)
    ctrl = NW.NetworkWrapper(net, ds, horizon=h, dt=dt, take_every_x=9)
    ev = ED.QuadEvaluator(ctrl, QuadRotorEnvBase(FlightmareDynamics(), dt), ref_length=h, dt=dt, speed_factor=0.4)
    tables = torch.tensor(np.stack([g["gentle_table"], g["fast_reset_table"]]), dtype=torch.float32)
    got = ev.run_eval("rand", nr_test=2, max_steps=80, thresh_div=1.0, thresh_stable=1.0, tables=tables)
    divs = [g["gentle_div"], g["fast_reset_div"]]
    per_run = np.array([d.mean() for d in divs])
    stable = np.array([(d < 1.0).sum() for d in divs])
    full = per_run[stable == len(divs[-1])]
    want = (stable.mean(), stable.std(), full.mean() if len(full) else np.nan, full.std() if len(full) else np.nan,
            per_run.mean(), per_run.std())
    assert np.allclose(np.array(got, dtype=np.float64), np.array(want, dtype=np.float64), atol=2e-4, equal_nan=True)
    assert ctrl.action_counter == 160 and ds.eval_counter == 160 // 9
    # ---- fixed wing: random targets drawn like run_eval, against the single-flight wrapper
    gw = load_golden("eval_wing.npz")
    hw, dt_data, dt_env = int(gw["cfg"][0]), float(gw["cfg"][1]), float(gw["cfg"][2])
    wnet = HM.Net(9, 1, 3, 4 * hw, conv=False)
    with torch.no_grad():
        for i, p in enumerate(wnet.parameters()):
            p.copy_(torch.tensor(gw[f"param_{i}"]))
    wds = DS.WingDataset(np.zeros((2, 12)), np.ones((2, 3)), mean=gw["mean"], std=gw["std"], delta_t=dt_data,
                         horizon=hw)
    wev = EF.FixedWingEvaluator(NW.FixedWingNetWrapper(wnet, wds, horizon=hw), SimpleWingEnv(FixedWingDynamics(), dt_env),
                                dt=dt_env, horizon=hw, thresh_div=4.0, thresh_stable=0.4)
    np.random.seed(3)
    mean_err, std_err = wev.run_eval(nr_test=3, printout=False)
    np.random.seed(3)
    targets = np.array([[[50.0, *((np.random.rand(2) - .5) * 10)]] for _ in range(3)])
    one = EV.WingTargetEvaluator(R.RolloutSpec.wing_concurrent(hw, dt_env), 3, gw["mean"], gw["std"], dt_data, "cpu")
    params = [torch.tensor(gw[f"param_{i}"]) for i in range(14)]
    ref = one.fly(R.flatten_params(params), torch.tensor(targets, dtype=torch.float32), steps=1000, thresh_div=4.0,
                  thresh_stable=0.4)
    m, sd = EV.wing_eval_statistics(ref["div_target_sum"], ref["div_target_cnt"])
    assert abs(mean_err - m) <= 1e-6 and abs(std_err - sd) <= 1e-6 and np.isfinite(m)
    assert wev.controller.action_counter == int(ref["n_steps"].sum())
    # ---- cartpole: golden starts as one batch
    gc = load_golden("eval_cartpole.npz")
    cnet = SM.Net(4, 10)
    with torch.no_grad():
        for i, p in enumerate(cnet.parameters()):
            p.copy_(torch.tensor(gc[f"param_{i}"]))
    cev = EC.Evaluator(NW.CartpoleWrapper(cnet, horizon=10), CartPoleEnv(CartpoleDynamics(), 0.05, thresh_div=0.21))
    names = ["zero_start", "tilted", "falls", "falls_at_once"]                   # the runs with thresh_div 0.21
    init = np.stack([gc[f"{n}_init"] for n in names])
    succ, vel = cev.evaluate_in_environment(nr_iters=4, max_steps=60, burn_in_steps=5, return_success=1,
                                            init_states=init)
    want_succ = [min(int(gc[f"{n}_success"][0]), 59) for n in names]
    assert succ.tolist() == want_succ
    want_vel = np.concatenate([gc[f"{n}_vel"][:60] for n in names])
    assert len(vel) == len(want_vel) and np.abs(np.array(vel) - want_vel).max() <= 2e-5
    res = cev.evaluate_in_environment(nr_iters=2, max_steps=20)
    assert res["mean_stable"] == 19.0 and res["std_stable"] == 0.0 and res["mean_vel"] >= 0.0


def test_quad_evaluator_loads_tables_like_random_reference(hostlib, monkeypatch, tmp_path):
    """QuadEvaluator.load_tables = Random.__init__ over trajectory files (random_traj.py:29-36)"""
    from apg_trajectory_tracking_b200.scripts import evaluate_drone as ED
    from apg_trajectory_tracking_b200.neural_control.models import hutter_model as HM
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    g = load_golden("ref_table.npz")
    (tmp_path / "train").mkdir()
    np.save(tmp_path / "train" / "traj_0.npy", g["a_raw"].astype(np.float64))
    dt, speed = [float(v) for v in g["a_cfg"]]
    ctrl = NW.NetworkWrapper(HM.Net(15, 10, 9, 40), None, horizon=10, dt=dt)
    ev = ED.QuadEvaluator(ctrl, None, ref_length=10, dt=dt, speed_factor=speed, data_dir=str(tmp_path))
    tabs = ev.load_tables(3, "cpu")
    want = g["a_table"].copy()
    want[:, 2] += 3
    assert tuple(tabs.shape) == (3,) + want.shape and np.abs(tabs[1].numpy() - want).max() <= 3e-6


def test_wing_selfplay_feed_matches_reference_dataset(hostlib, monkeypatch):
    """fixed-wing evaluation -> self-play slots (network_wrapper.py:81-90, dataset.py:98-119): the three flights of
    tests/golden/eval_wing_selfplay.npz through FixedWingEvaluator.fly_to_points (host-compiled kernel logic) with
    ONE controller / dataset, against what the reference's wrapper + dataset hold afterwards"""
    from apg_trajectory_tracking_b200.scripts import evaluate_fixed_wing as EF
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.wing_env import SimpleWingEnv
    from apg_trajectory_tracking_b200.neural_control.models import hutter_model as HM
    NW = _single_drone_mirrors_on_cpu(monkeypatch)
    g, gw = load_golden("eval_wing_selfplay.npz"), load_golden("eval_wing.npz")
    h, dt_data, dt_env, take, n_sampled, n_slots = [float(v) for v in g["cfg"]]
    h, take, n_sampled, n_slots = int(h), int(take), int(n_sampled), int(n_slots)
    net = HM.Net(9, 1, 3, 4 * h, conv=False)
    with torch.no_grad():
        for i, p in enumerate(net.parameters()):
            p.copy_(torch.tensor(gw[f"param_{i}"]))
    tot = n_sampled + n_slots
    ds = DS.WingDataset(np.zeros((tot, 12)), np.ones((tot, 3)), mean=gw["mean"], std=gw["std"], delta_t=dt_data,
                        horizon=h, self_play=n_slots / n_sampled)
    assert ds.num_sampled_states == n_sampled and ds.num_self_play == n_slots
    ctrl = NW.FixedWingNetWrapper(net, ds, horizon=h, take_every_x=take)
    recorded = []
    real = ds.get_and_add_eval_data

    def recording(st, rf, add_to_dataset=False):
        if add_to_dataset:
            recorded.append((np.array(st), np.array(rf)))
        return real(st, rf, add_to_dataset=add_to_dataset)
    ds.get_and_add_eval_data = recording
    for name in [str(v) for v in g["run_names"]]:
        steps, tdiv, tstab = g[f"{name}_cfg"]
        ev = EF.FixedWingEvaluator(ctrl, SimpleWingEnv(FixedWingDynamics(), dt_env), dt=dt_env, horizon=h,
                                   thresh_div=float(tdiv), thresh_stable=float(tstab), test_time=0)
        out = ev.fly_to_points(g[f"{name}_targets"][None], max_steps=int(steps))
        assert int(out["n_steps"][0]) == int(g[f"{name}_n_steps"][0])
    assert ctrl.action_counter == int(g["action_counter"][0]) and ds.eval_counter == int(g["eval_counter"][0])
    ks, kt = np.array([r[0] for r in recorded]), np.array([r[1] for r in recorded])
    assert np.abs(kt - g["kept_targets"]).max() <= 1e-6
    assert np.abs(ks - g["kept_states"]).max() <= 1e-4 * np.abs(g["kept_states"]).max()
    # the ring afterwards (the reference stores the prepared tensors of each kept sample)
    sl = slice(n_sampled, tot)
    assert np.abs(ds.states[sl].numpy() - g["ds_states"][sl]).max() <= 1e-4 * np.abs(g["ds_states"]).max()
    assert np.abs(ds.normed_states[sl].numpy() - g["ds_normed_states"][sl]).max() <= 5e-4
    assert np.abs(ds.ref_states[sl].numpy() - g["ds_ref_states"][sl]).max() <= 1e-4 * np.abs(g["ds_ref_states"]).max()


def test_wing_selfplay_samples_batched_selection_matches_sequential_runs():
    gw = load_golden("eval_wing.npz")
    params = [torch.tensor(gw[f"param_{i}"]) for i in range(14)]
    h, dt_data, dt_env = int(gw["cfg"][0]), float(gw["cfg"][1]), float(gw["cfg"][2])
    n, K, steps = 6, 3, 160
    gen = torch.Generator().manual_seed(2)
    targets = torch.zeros(n, K, 3)
    for k, x in enumerate((20.0, 40.0, 62.0)):
        targets[:, k] = torch.tensor([x, 0, 0]) + (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([4.0, 6, 6])
    init = torch.zeros(n, 12)
    init[:, 3] = 11.5
    out = O.eval_fly_to_points(params, targets, init, gw["mean"], gw["std"], steps, h, dt_data, dt_env, 2.0, 0.4, 0,
                               record_policy_inputs=True)
    assert int(out["target_index"].max()) == K - 1                       # flights that switch targets twice
    for take, ac in ((3, 0), (7, 5), (10000, 0)):
        kept, after = O.selfplay_kept_calls(out["n_steps"], take, ac)
        s, tg, counter = EV.wing_selfplay_samples(out, targets, take, ac)
        assert counter == after and s.shape[0] == len(kept)
        if kept:
            assert torch.equal(s, torch.stack([out["policy_states"][j, i] for j, i in kept]))
            assert torch.equal(tg, torch.stack([targets[j, int(out["target_index"][j, i])] for j, i in kept]))


def test_datasets_with_device_prepare_match_host_containers(hostlib):
    """QuadDataset / WingDataset(device=...) run prepare_data through the prepare kernels (here: their host-compiled
    bodies) and must hand out the same 4-tuples as the host containers, self-play slots included"""
    g = load_golden("prep_data.npz")
    qs, qr = g["quad_raw_states"], g["quad_raw_refs"]
    host, dev = DS.QuadDataset(qs, qr, self_play=0.5), DS.QuadDataset(qs, qr, self_play=0.5, device="cpu")
    assert dev.num_self_play == host.num_self_play > 0
    for a, b in zip(host[3], dev[3]):
        assert torch.allclose(a, b, atol=2e-6)
    for d in (host, dev):
        d.get_and_add_eval_data(qs[1].copy(), qr[1].copy(), add_to_dataset=True)
    at = host.num_sampled_states
    assert torch.allclose(host.states[at], dev.states[at], atol=2e-6) and dev.eval_counter == 1
    assert torch.allclose(host.in_ref_states[at], dev.in_ref_states[at], atol=2e-6)
    ws, wt = g["wing_raw_states"], g["wing_targets"]
    kw = dict(mean=g["wing_mean"], std=g["wing_std"], delta_t=float(g["wing_dt"]), horizon=int(g["wing_h"]))
    wh, wd = DS.WingDataset(ws, wt, **kw), DS.WingDataset(ws, wt, device="cpu", **kw)
    for a, b in zip(wh[2], wd[2]):
        assert torch.allclose(a, b, atol=5e-6)
    loader = torch.utils.data.DataLoader(dev, batch_size=4, shuffle=False)
    assert next(iter(loader))[0].shape == (4, 15)


def test_environment_data_sampling_mirrors(hostlib, monkeypatch, tmp_path):
    """full_state_training_data (drone_env.py:232-269) over trajectory files and construct_states
    (cartpole_env.py:178-236) on the mirrors"""
    from apg_trajectory_tracking_b200.neural_control.environments import cartpole_env as CE, drone_env as DE
    _single_drone_mirrors_on_cpu(monkeypatch)
    g = load_golden("ref_table.npz")
    (tmp_path / "train").mkdir()
    np.save(tmp_path / "train" / "traj_0.npy", g["a_raw"].astype(np.float64))
    dt, speed = [float(v) for v in g["a_cfg"]]
    L = 5
    np.random.seed(0)
    states, refs = DE.full_state_training_data(25, ref_length=L, dt=dt, speed_factor=speed, data_dir=str(tmp_path))
    assert states.shape == (25, 12) and refs.shape == (25, L, 9)
    traj = g["a_table"]                                   # what the reference's load_prepare_trajectory returns
    cut = traj[:-(L + 1)]
    starts = cut[::2 * L]
    n = len(starts)
    want_states = np.hstack((starts, np.zeros((n, 3))))
    want_refs = np.zeros((n, L, 9))
    for i in range(1, L + 1):
        want_refs[:, i - 1] = traj[i::2 * L][:n]
    reps = 25 // n + 1
    assert np.abs(states - np.tile(want_states, (reps, 1))[:25]).max() <= 3e-6
    assert np.abs(refs - np.tile(want_refs, (reps, 1, 1))[:25]).max() <= 3e-6
    np.random.seed(1)
    data = CE.construct_states(60, 0.05)
    assert data.shape == (60, 4) and np.isfinite(data).all() and np.abs(data[:, 2]).max() <= np.pi
    # the first run: 20 pushes from one slowed-down random state, each state the dynamics step of the previous one
    nxt = O.cartpole_step(torch.tensor(data[:19], dtype=torch.float32), torch.zeros(19, 1), 0.05).numpy()
    assert np.abs(nxt[:, 0] - data[1:20, 0]).max() <= 1e-5           # x' = x + x_dot * dt does not depend on the push


def test_wing_flight_sampler_mirror(hostlib, monkeypatch):
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from apg_trajectory_tracking_b200.neural_control.environments import wing_env as WE
    _single_drone_mirrors_on_cpu(monkeypatch)
    np.random.seed(0)
    env = WE.SimpleWingEnv(FixedWingDynamics(), 0.05)
    traj = WE.run_wing_flight(env, traj_len=25)
    assert traj.ndim == 2 and traj.shape[1] == 12 and 1 <= len(traj) <= 25 and np.isfinite(traj).all()
    # first state = one step of the oracle dynamics from zero_reset with the first drawn action
    np.random.seed(0)
    a0 = np.clip(np.random.normal(scale=.15, size=4) + np.array([.25, .5, .5, .5]), 0, 1)
    s0 = torch.zeros(1, 12)
    s0[0, 3] = 11.5
    want = O.wing_step(s0, torch.tensor(a0, dtype=torch.float32)[None], 0.05)[0].numpy()
    assert np.abs(traj[0] - want).max() <= 1e-5 * np.abs(want).max()
    v = WE.generate_unit_vecs(50)
    assert v.shape == (50, 3) and (v[:, 0] >= 0.01).all()
