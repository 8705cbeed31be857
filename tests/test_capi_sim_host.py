"""The WHOLE library on the CPU: tests/hostcheck/capisim builds libapg_b200_sim.so from the unchanged product sources
(csrc/capi.cu, capi_prep.cu and every *_kernels.cu) with -DAPG_SIM: the C-ABI layer runs on a CUDA-runtime shim,
every kernel launch (through the real launchers: grid / block / dynamic shared memory sizes) on the software models
of tests/hostcheck (gpu_sim.h, te_sim.h, tc_sim.h); a canary behind the requested dynamic shared memory catches
kernels that write past the size their launcher computed.  The product's own Python wrappers are pointed at that
library ("device" pointers = host pointers), so the tests below are the GPU parity tests in miniature - what they
add over the per-kernel model tests is the host side: argument checks, workspace plan offsets, weight packing by the
real pack kernels, kernel selection by configuration and by the APG_LEGACY_MMA / peer-exchange switches."""
import contextlib
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import bench as B
from apg_trajectory_tracking_b200 import _capi, evaluate as EV, ops, prepare as PR, rollout as R, synthetic as SY
from apg_trajectory_tracking_b200.neural_control.dynamics import quad_dynamics_trained as QT
# imported HERE, before any fixture patches ops._require_cuda: these modules bind that name at import time, and a first
# import under the patch would leave them without their CUDA-only guard for the rest of the session
from apg_trajectory_tracking_b200 import device_data, train  # noqa: F401
from apg_trajectory_tracking_b200.neural_control import dataset as _ds, drone_loss as _dl  # noqa: F401
from apg_trajectory_tracking_b200.neural_control.dynamics import (cartpole_dynamics, fixed_wing_dynamics,  # noqa: F401
                                                                   quad_dynamics_flightmare)
from apg_trajectory_tracking_b200.neural_control.models import hutter_model, rnn, simple_model  # noqa: F401
from oracle import apg_oracle as O
from tests.helpers import golden_params, load_golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def simlib_path(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("capisim")
    objs = []
    for name in ("host", "te", "tq"):
        obj = tmp / f"{name}.o"
        subprocess.check_call(["g++", "-O1", "-c", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                               "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                               os.path.join(ROOT, "tests", "hostcheck", "capisim", f"{name}.cpp"), "-o", str(obj)])
        objs.append(str(obj))
    out = tmp / "libapg_b200_sim.so"
    subprocess.check_call(["g++", "-shared", "-pthread"] + objs + ["-o", str(out)])
    return str(out)


class _FakeStream:
    cuda_stream = 0


@pytest.fixture
def simlib(simlib_path, monkeypatch):
    lib = ctypes.CDLL(simlib_path)
    for name, (res, args) in _capi.EXPORTS.items():
        fn = getattr(lib, name)                       # every declared entry point exists in the model library too
        fn.restype, fn.argtypes = res, args
    lib.apg_sim_take_errors.restype = ctypes.c_int
    monkeypatch.setattr(_capi, "lib", lambda: lib)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: _FakeStream())
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(R, "_dev_f32", lambda t, name: None if t is None else t.contiguous().float())
    for mod in (ops, PR, EV, QT):
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None, raising=False)
        monkeypatch.setattr(mod, "_stream", lambda t: ctypes.c_void_p(0), raising=False)
    monkeypatch.delenv("APG_LEGACY_MMA", raising=False)
    if os.environ.get("APG_SIM_POISON_SMEM"):
        # stress run: the workspace (stashes, partials, packed weights) starts as NaN patterns, like recycled HBM may
        real_init = R.Rollout.__init__

        def poisoned_init(self, *a, **k):
            real_init(self, *a, **k)
            self.workspace.fill_(0xFF)
        monkeypatch.setattr(R.Rollout, "__init__", poisoned_init)
    yield lib
    buf = ctypes.create_string_buffer(4096)
    n = lib.apg_sim_take_errors(buf, 4096)
    assert n == 0, buf.value.decode()


def _quad_case(n, seed):
    params = B.default_init("quad", 10, seed=seed)
    case = SY.quad_case(n, 10, 0.1, seed=seed)
    want = O.concurrent_value_and_grad("quad", params, case["in_state"], case["cur"], case["in_ref"], case["ref"], 10,
                                       0.1)
    return params, case, want


def _check(loss, grad, params, want, tol=5e-5):
    want_loss, want_grad = want[0], want[1]
    assert abs(float(loss) - float(want_loss)) <= 2e-5 * abs(float(want_loss))
    for got, g in zip(R.split_flat(grad, params), want_grad):
        if g is None:
            assert float(got.abs().max()) == 0
        else:
            assert float((got - g).abs().max()) <= tol * max(float(g.abs().max()), 1e-6)


_SLOW = pytest.mark.slow
@pytest.mark.parametrize("flags", [(), ("APG_LEGACY_MMA",)])
def test_rollout_forward_backward_through_the_c_abi_all_kernel_selections(simlib, monkeypatch, flags):
    """apg_rollout_forward + apg_rollout_backward, quadrotor concurrent, 100 drones on a 3-SM model: the default
    mma.sync kernels and every combination of the optional tcgen05 paths give the oracle's loss and gradient"""
    for k in flags:
        monkeypatch.setenv(k, "1")
    n = 100                                            # two 64-drone stash tiles, one (partial) 128-drone tcgen05 tile
    params, case, want = _quad_case(n, 21)
    runner = R.Rollout(R.RolloutSpec.quad_concurrent(10, 0.1), n, "cpu")
    flat = R.flatten_params(params)
    loss, grad = runner.value_and_grad(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    _check(loss, grad, params, want)


def _learnt_case(n, seed):
    """random policy + the reference-pinned learnt dynamics of tests/golden/learnt_dyn.npz (variant b: non-trivial
    action transform and residual MLP), quadrotor concurrent case"""
    g = load_golden("learnt_dyn.npz")
    params = B.default_init("quad", 10, seed=seed)
    case = SY.quad_case(n, 10, 0.1, seed=seed)
    d = QT.LearntDynamics({"rotational_drag": [float(x) for x in g["b_rot_drag"]]})
    with torch.no_grad():
        for i, (_, p) in enumerate(d.named_parameters()):
            if i not in (1, 2, 3):
                p.copy_(torch.tensor(g[f"b_param_{i}"]))
    lparams = [p.detach().clone() for _, p in d.named_parameters()]
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g["b_rot_drag"]))
    return params, case, d, lparams, cfg


def test_rollout_through_learnt_dynamics_fused(simlib):
    """apg_rollout_forward_learnt + apg_rollout_backward: the quadrotor concurrent rollout with every step taken by
    LearntDynamics (tq_dyn_kernel<true>) gives the oracle's loss, states and policy gradient; and it differs from the
    analytic rollout (the learnt parameters are live)"""
    n = 100
    params, case, d, lparams, cfg = _learnt_case(n, 23)
    want = O.value_and_grad(lambda ps: O.rollout_concurrent_learnt(ps, lparams, case["in_state"], case["cur"],
                                                                   case["in_ref"], case["ref"], 10, 0.1, cfg), params)
    spec = R.RolloutSpec.quad_concurrent(10, 0.1, modified_params=dict(d.cfg))
    runner = R.Rollout(spec, n, "cpu")
    flat = R.flatten_params(params)
    lflat = d._flat().detach()
    loss, states, _ = runner.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"],
                                     want_states=True, learnt_params=lflat)
    loss = loss.clone()
    grad = runner.backward(1.0)
    _check(loss, grad, params, want)
    assert float((states - want[2]).abs().max()) <= 2e-5 * float(want[2].abs().max())
    plain, _, _ = runner.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    assert abs(float(plain) - float(loss)) > 1e-3 * abs(float(loss))
    # configurations without the variant refuse instead of silently integrating the analytic model
    wing = R.Rollout(R.RolloutSpec.wing_concurrent(10, 0.05), 8, "cpu")
    wcase = SY.wing_case(8, 10, 0.05, seed=1) if hasattr(SY, "wing_case") else None
    if wcase is not None:
        wp = R.flatten_params(B.default_init("wing", 10, seed=1))
        with pytest.raises(_capi.ApgError):
            wing.forward(wp, wcase["in_state"], wcase["cur"], wcase["in_ref"], wcase["ref"],
                         learnt_params=torch.zeros(1914))


def test_peer_exchange_entry_points_single_rank(simlib):
    """apg_rollout_backward_p2p + apg_grad_gather_sgd_p2p with world = 1 on a host buffer: the gradient that comes
    out of the slot equals the plain backward's, and the fused SGD update equals the torch ops"""
    n = 130
    params, case, want = _quad_case(n, 5)
    runner = R.Rollout(R.RolloutSpec.quad_concurrent(10, 0.1), n, "cpu")
    flat = R.flatten_params(params).clone()
    loss, grad = runner.value_and_grad(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    npar = flat.numel()
    nbytes = simlib.apg_grad_comm_bytes(1, npar)
    buf = torch.zeros(nbytes // 4)
    so, fo = ctypes.c_size_t(), ctypes.c_size_t()
    assert simlib.apg_grad_comm_offsets(1, npar, 1, ctypes.byref(so), ctypes.byref(fo)) == 0
    slot_tab = torch.tensor([buf.data_ptr() + so.value], dtype=torch.int64)
    flag_tab = torch.tensor([buf.data_ptr() + fo.value], dtype=torch.int64)
    ticket = torch.zeros(1, dtype=torch.int32)
    comm = _capi.ApgGradComm(0, 1, ctypes.c_void_p(slot_tab.data_ptr()), ctypes.c_void_p(flag_tab.data_ptr()), 1,
                             ctypes.c_void_p(ticket.data_ptr()))
    runner.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    runner.backward_p2p(comm)
    g2, mom = torch.zeros(npar), torch.zeros(npar)
    p2 = flat.clone()
    _capi.check(simlib.apg_grad_gather_sgd_p2p(ctypes.byref(comm), ctypes.c_void_p(buf.data_ptr() + so.value), npar,
                                               ctypes.c_void_p(g2.data_ptr()), ctypes.c_void_p(p2.data_ptr()),
                                               ctypes.c_void_p(mom.data_ptr()), ctypes.c_float(1e-6),
                                               ctypes.c_float(0.9), None))
    assert torch.equal(g2, grad) and torch.equal(mom, grad)
    assert torch.allclose(p2, flat - 1e-6 * grad, rtol=0, atol=1e-9) and int(ticket[0]) == 0


@pytest.mark.slow
def test_other_train_modes_through_the_c_abi(simlib):
    """wing and cartpole concurrent, quadrotor autoregressive and LSTM (GPU-verified kernels: their agreement with the
    oracle on the model also validates the model on four more kernel pairs)"""
    for workload in ("wing_concurrent", "cartpole_concurrent", "quad_autoregressive", "quad_lstm"):
        w = dict(B.WORKLOADS[workload])
        n, h = 66, min(w["h"], 4)
        w["h"] = h
        case = B.make_case(w, n, 3, "cpu")
        params = B.default_init(w["system"], h, mode=w.get("mode", "concurrent"))
        spec = B.make_spec(w)
        runner = R.Rollout(spec, n, "cpu")
        flat = R.flatten_params(params)
        loss, grad = runner.value_and_grad(flat, case.get("in_state"), case["cur"], case.get("in_ref"), case.get("ref"),
                                           case.get("h0c0"))
        if w.get("mode", "concurrent") == "concurrent":
            want = O.concurrent_value_and_grad(w["system"], params, case["in_state"], case["cur"], case.get("in_ref"),
                                               case.get("ref"), h, spec.dt)
        else:
            hc = (case["h0c0"][0], case["h0c0"][1]) if case.get("h0c0") is not None else None
            want = O.recurrent_value_and_grad(w["mode"].lower(), params, case["cur"], case["in_ref"], case["ref"], h,
                                              spec.dt, hc0=hc)
        _check(loss, grad, params, want, tol=2e-4)


def test_evaluation_entry_points_through_the_c_abi(simlib):
    """apg_eval_rollout / apg_eval_fly_to_points / apg_eval_cartpole through the product's evaluator classes on the
    model library: host-side argument checks, workspace plan, real pack kernel, real launchers"""
    from tests.test_oracle_golden import wing_eval_case
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    steps = 12
    ev = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(10, 0.1), 2, "cpu")
    tabs = torch.tensor(np.stack([g["gentle_table"], g["fast_reset_table"]]), dtype=torch.float32)
    out = ev.follow(R.flatten_params(params), tabs, steps=steps, thresh_div=1.0, thresh_stable=1.0, test_time=0)
    assert out["n_steps"].tolist() == [steps, steps]
    for k, name in enumerate(("gentle", "fast_reset")):
        assert np.abs(out["states"][k].numpy() - g[f"{name}_states"][:steps + 1]).max() <= 5e-5
    gw = load_golden("eval_wing.npz")
    wparams, targets, init, h, dt_data, dt_env, _, test_time, tdiv, tstab = wing_eval_case(gw, "one_target")
    wev = EV.WingTargetEvaluator(R.RolloutSpec.wing_concurrent(h, dt_env), 1, gw["mean"], gw["std"], dt_data, "cpu")
    wout = wev.fly(R.flatten_params(wparams), targets, steps=15, thresh_div=tdiv, thresh_stable=tstab)
    traj = gw["one_target_traj"]
    assert np.abs(wout["states"][0, 1:16].numpy() - traj[:15, :12]).max() <= 1e-4 * np.abs(traj[:, :12]).max()
    gc = load_golden("eval_cartpole.npz")
    cparams = [torch.tensor(gc[f"param_{i}"]) for i in range(10)]
    cev = EV.CartpoleBalanceEvaluator(R.RolloutSpec.cartpole_concurrent(10, 0.05), 2, "cpu")
    init = torch.tensor(np.stack([gc["falls_init"], gc["tilted_init"]]), dtype=torch.float32)
    cout = cev.balance(R.flatten_params(cparams), init, steps=10, thresh_div=0.21, burn_in_steps=5)
    assert cout["n_steps"].tolist() == [7, 10]
    assert np.abs(cout["states"][0, :7].numpy() - gc["falls_states"]).max() <= 2e-5


def test_evaluation_with_the_lstm_policy_through_the_c_abi(simlib):
    """apg_eval_rollout_lstm (eval_rollout_lstm_kernel): the reference evaluator's runs with an LSTM_NEW policy
    (tests/golden/eval_rand_lstm.npz: states, divergences, actions, final hidden / cell state), one drone per run in
    ONE launch where the tables have one length; then a ragged batch against the oracle"""
    g = load_golden("eval_rand_lstm.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    flat = R.flatten_params(params)
    spec = R.RolloutSpec.quad_recurrent("lstm", 10, 0.1)
    for name in ("gentle", "fast_stop", "loose"):
        steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
        steps = min(int(steps), 14)                      # (the CPU model runs one OS thread per GPU thread)
        taken = min(len(g[f"{name}_div"]), steps)
        ev = EV.TableEvaluator(spec, 1, "cpu")
        h0c0 = torch.tensor(np.stack([g[f"{name}_h0"], g[f"{name}_c0"]]))
        out = ev.follow(flat, torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None],
                        init_states=torch.tensor(g[f"{name}_states"][:1], dtype=torch.float32), steps=steps,
                        thresh_div=tdiv, thresh_stable=tstab, test_time=int(test_time), h0c0=h0c0)
        assert int(out["n_steps"][0]) == taken
        assert np.abs(out["states"][0, :taken + 1].numpy() - g[f"{name}_states"][:taken + 1]).max() <= 1e-4
        assert np.abs(out["div"][0, :taken].numpy() - g[f"{name}_div"][:taken]).max() <= 1e-4
        assert np.abs(out["actions"][0, :taken].numpy() - g[f"{name}_actions"][:taken]).max() <= 1e-4
        if taken == len(g[f"{name}_div"]):
            assert np.abs(out["hc"][0, 0].numpy() - g[f"{name}_h1"][0]).max() <= 1e-4
            assert np.abs(out["hc"][1, 0].numpy() - g[f"{name}_c1"][0]).max() <= 1e-4
    n, steps = 70, 5                                     # two tiles, the second ragged; shared tables through the index
    tabs = torch.tensor(np.stack([g["gentle_table"][:100], g["loose_table"][:100]]), dtype=torch.float32)
    gen = torch.Generator().manual_seed(5)
    index = torch.randint(0, 2, (n,), generator=gen, dtype=torch.int32)
    init = torch.zeros(n, 12)
    init[:, :3] = tabs[index.long(), 0, :3] + 0.05 * torch.randn(n, 3, generator=gen)
    hc0 = torch.randn(2, n, 8, generator=gen)
    want = O.eval_follow_tables(params, tabs[index.long()], init, steps, 10, 0.1, 0.5, 0.4, 1, hc0=(hc0[0], hc0[1]))
    out = EV.TableEvaluator(spec, n, "cpu").follow(flat, tabs, init_states=init, table_index=index, steps=steps,
                                                   thresh_div=0.5, thresh_stable=0.4, test_time=1, h0c0=hc0)
    assert torch.equal(out["n_steps"], want["n_steps"].to(torch.int32))
    assert float((out["states"] - want["states"]).abs().max()) <= 1e-4
    assert float((out["hc"][0] - want["hc"][0]).abs().max()) <= 1e-4
    assert float((out["hc"][1] - want["hc"][1]).abs().max()) <= 1e-4
    with pytest.raises(ValueError):
        EV.TableEvaluator(spec, n, "cpu").follow(flat, tabs, init_states=init, table_index=index, steps=steps)
    # the evaluator mirror with an LSTM net: starts every run from the net's state, leaves the last run's state in it
    import types
    from apg_trajectory_tracking_b200.scripts import evaluate_drone as ED
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    net = rnn.LSTM_NEW(15, 10, 9, 4)
    with torch.no_grad():
        for p_, q in zip(net.parameters(), params):
            p_.copy_(q)
    ctrl = types.SimpleNamespace(net=net, action_counter=0)
    env = types.SimpleNamespace(dynamics=FlightmareDynamics(), dt=0.1)
    qe = ED.QuadEvaluator(ctrl, env, ref_length=10, dt=0.1, speed_factor=0.4, train_mode="LSTM")
    net.hidden_state, net.cell_state = torch.tensor(g["gentle_h0"]), torch.tensor(g["gentle_c0"])
    two = torch.tensor(np.stack([g["gentle_table"][:100], g["gentle_table"][:100]]), dtype=torch.float32)
    res = qe.follow_tables(two, max_nr_steps=6, thresh_stable=1.0, thresh_div=1.0,
                           init_states=torch.tensor(np.repeat(g["gentle_states"][:1], 2, 0), dtype=torch.float32))
    assert np.abs(res["states"][1, :7].numpy() - g["gentle_states"][:7]).max() <= 1e-4
    assert torch.equal(res["states"][0], res["states"][1]) and ctrl.action_counter == 12
    assert torch.equal(net.hidden_state, res["hc"][0, -1:]) and tuple(net.cell_state.shape) == (1, 8)


def test_input_side_and_learnt_entry_points_through_the_c_abi(simlib):
    g = load_golden("prep_data.npz")
    out = PR.prepare_quad(torch.tensor(g["quad_raw_states"]), torch.tensor(g["quad_raw_refs"]))
    assert np.abs(out["in_state"].numpy() - g["quad_in_state"]).max() <= 2e-6
    assert np.abs(out["in_ref"].numpy() - g["quad_in_ref"]).max() <= 2e-6
    wo = PR.prepare_wing(torch.tensor(g["wing_raw_states"]), torch.tensor(g["wing_targets"]), g["wing_mean"],
                         g["wing_std"], float(g["wing_dt"]), int(g["wing_h"]))
    assert np.abs(wo["ref"].numpy() - g["wing_ref"]).max() <= 1e-5 * np.abs(g["wing_ref"]).max()
    gp = load_golden("poly_traj.npz")
    pts, ref_len = PR.polynomial_points(torch.tensor(gp["b_coef"])[None], torch.tensor(gp["b_rot"])[None],
                                        torch.tensor(gp["b_start"])[None], x_range=6, max_drone_dist=0.5, horizon=10,
                                        hover_steps=5)
    assert int(ref_len[0]) == len(gp["b_points"])
    assert np.array_equal(pts[0, :len(gp["b_points"])].numpy(), gp["b_points"].astype(np.float32))
    gt = load_golden("ref_table.npz")
    tab = PR.reference_table(torch.tensor(gt["c_raw"]), float(gt["c_cfg"][0]), float(gt["c_cfg"][1]), z_offset=0.0)
    assert np.abs(tab.numpy() - gt["c_table"]).max() <= 3e-6
    # learnt residual dynamics: forward + adjoint (wing: all 46 physical parameters live)
    gl = load_golden("learnt_dyn.npz")
    flat = torch.tensor(np.concatenate([np.asarray(gl[f"wb_param_{i}"]).reshape(-1) for i in range(42)]),
                        dtype=torch.float32, requires_grad=True)
    s = torch.tensor(gl["wb_state"], dtype=torch.float32, requires_grad=True)
    a = torch.tensor(gl["wb_action"], dtype=torch.float32, requires_grad=True)
    phys = np.zeros(48, np.float32)
    o = QT._LearntStep.apply(flat, s, a, float(gl["wb_dt"]), phys, 1)
    assert rel_err(o.detach(), torch.tensor(gl["wb_out"])) <= 5e-6
    (o * torch.tensor(gl["wb_cot"])).sum().backward()
    want = np.concatenate([np.asarray(gl[f"wb_gparam_{i}"]).reshape(-1) for i in range(42)])
    assert np.abs(flat.grad.numpy() - want).max() <= 1e-4 * np.abs(want).max()
    assert rel_err(s.grad, torch.tensor(gl["wb_gstate"])) <= 5e-5
    # quadrotor: the construction-time simulator constants travel as `phys`
    from apg_trajectory_tracking_b200 import params as P
    qflat = torch.tensor(np.concatenate([np.asarray(gl[f"b_param_{i}"]).reshape(-1) for i in range(8)]),
                         dtype=torch.float32, requires_grad=True)
    qs = torch.tensor(gl["b_state"], dtype=torch.float32, requires_grad=True)
    qa = torch.tensor(gl["b_action"], dtype=torch.float32, requires_grad=True)
    qphys = P.PHYS["quad"]({"rotational_drag": [float(x) for x in gl["b_rot_drag"]]})
    qo = QT._LearntStep.apply(qflat, qs, qa, float(gl["b_dt"]), qphys, 0)
    assert rel_err(qo.detach(), torch.tensor(gl["b_out"])) <= 5e-6
    (qo * torch.tensor(gl["b_cot"])).sum().backward()
    qwant = np.concatenate([np.asarray(gl[f"b_gparam_{i}"]).reshape(-1) for i in range(8)])
    keep = np.ones(qwant.size, bool)
    keep[16] = False                                   # mass: cancels analytically, autograd leaves rounding noise
    assert np.abs(qflat.grad.numpy() - qwant)[keep].max() <= 1e-4 * np.abs(qwant).max()
    assert rel_err(qa.grad, torch.tensor(gl["b_gaction"])) <= 5e-5


@pytest.mark.slow
@pytest.mark.parametrize("system", ["quad", "wing", "cartpole", "autoregressive", "lstm"])
def test_raw_sample_train_step_on_the_model_library(simlib, monkeypatch, system):
    """FusedTrainStep.step_host (raw samples -> chunked staging -> prepare kernels -> forward -> adjoint per chunk)
    with the REAL prepare / rollout kernels of the model library behind it (CUDA streams / events stubbed), against
    the whole-batch step from prepared inputs: 130 drones in chunks of 64"""
    from apg_trajectory_tracking_b200 import train as T
    from tests.test_step_host_logic_cpu import _Event, _Stream as _BaseStream

    class _Stream(_BaseStream):
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items()
                                                                        if kk != "pin_memory"}))
    n = 130
    h, dt = {"quad": (10, 0.1), "wing": (6, 0.05), "cartpole": (5, 0.05), "autoregressive": (4, 0.1),
             "lstm": (4, 0.1)}[system]
    mode = system if system in ("autoregressive", "lstm") else "concurrent"
    system = "quad" if mode != "concurrent" else system
    w = dict(system=system, mode=mode, h=h, dt=dt)
    params = B.default_init(system, h, seed=4, mode=mode)
    spec = B.make_spec(w)
    case = B.make_case(w, n, 21, "cpu")
    a = T.FusedTrainStep(params, spec, n, lr=1e-4, device="cpu", distributed=False)
    b = T.FusedTrainStep(params, spec, n, lr=1e-4, device="cpu", distributed=False)
    a._dev = lambda x: x
    la = a.step(case.get("in_state"), case["cur"], case.get("in_ref"), case.get("ref"), case.get("h0c0"))
    raw = B.raw_host_inputs(w, case)
    lb = b.step_host(chunk=64, **raw)
    assert abs(float(la) - float(lb)) <= 2e-5 * abs(float(la))
    assert rel_err(b.grad, a.grad) <= 2e-4 and rel_err(b.flat, a.flat) <= 1e-6
    # quad concurrent = the tcgen05 path on RAW samples: memset-stamped pack, chain, dynamics, dX chain, dW, reduce and
    # no prepare kernels; the others: prepare (2) + their 5 launches
    assert b.host_launches_per_step == 3 * (5 if system == "cartpole" else (6 if system == "quad" and mode == "concurrent" else 7))


def test_plain_c_demo_against_the_model_library(simlib_path, tmp_path):
    """examples/c_abi_demo.c linked against libapg_b200_sim.so: apg_rollout_value_and_grad_host (device buffer cache,
    H2D / D2H copies, forward, backward) from plain C; its loss against the oracle, its own finite-difference check"""
    import re
    from tests.test_zy_c_abi_example import _lcg_inputs
    exe = str(tmp_path / "c_abi_demo_sim")
    libdir = os.path.dirname(simlib_path)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", libdir, "-lapg_b200_sim",
                           "-lm", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    loss = float(re.search(r"^loss (\S+)", r.stdout, re.M).group(1))
    params, st = _lcg_inputs(256, 5)
    want, _, _, _ = O.concurrent_value_and_grad("cartpole", params, st, st, None, None, 5, 0.05)
    assert abs(loss - float(want)) <= 2e-5 * abs(float(want))


@pytest.mark.slow
def test_two_ranks_exchange_gradients_through_peer_slots(simlib):
    """data-parallel step of two "ranks" (two runners, two symmetric buffers in this process): each rank's
    apg_rollout_backward_p2p scatters into both buffers, each rank's gather sums the slots in rank order - the result
    equals the gradient of the concatenated batch and is bitwise identical on both ranks"""
    world, n_each = 2, 70
    params, case, want = _quad_case(world * n_each, 8)
    flat = R.flatten_params(params)
    npar = flat.numel()
    nbytes = simlib.apg_grad_comm_bytes(world, npar)
    bufs = [torch.zeros(nbytes // 4) for _ in range(world)]
    so, fo = ctypes.c_size_t(), ctypes.c_size_t()
    assert simlib.apg_grad_comm_offsets(world, npar, 1, ctypes.byref(so), ctypes.byref(fo)) == 0
    slot_tab = torch.tensor([b.data_ptr() + so.value for b in bufs], dtype=torch.int64)
    flag_tab = torch.tensor([b.data_ptr() + fo.value for b in bufs], dtype=torch.int64)
    tickets = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
    comms, losses = [], []
    for r in (1, 0):                                                        # arrival order must not matter
        sl = slice(r * n_each, (r + 1) * n_each)
        runner = R.Rollout(R.RolloutSpec.quad_concurrent(10, 0.1), n_each, "cpu")
        shard = [case[k][sl].clone() for k in ("in_state", "cur", "in_ref", "ref")]    # own (aligned) allocations
        loss, _, _ = runner.forward(flat, *shard)
        losses.append(float(loss))
        comm = _capi.ApgGradComm(r, world, ctypes.c_void_p(slot_tab.data_ptr()), ctypes.c_void_p(flag_tab.data_ptr()),
                                 1, ctypes.c_void_p(tickets[r].data_ptr()))
        runner.backward_p2p(comm)
        comms.append((r, comm))
    grads = {}
    for r, comm in comms:
        g = torch.zeros(npar)
        _capi.check(simlib.apg_grad_gather_sgd_p2p(ctypes.byref(comm), ctypes.c_void_p(bufs[r].data_ptr() + so.value),
                                                   npar, ctypes.c_void_p(g.data_ptr()), None, None,
                                                   ctypes.c_float(0), ctypes.c_float(0), None))
        grads[r] = g
    assert torch.equal(grads[0], grads[1])
    _check(sum(losses), grads[0], params, want)


@pytest.mark.slow
def test_device_dataset_epoch_on_the_model_library(simlib):
    """DeviceQuadDataset + run_epoch_device (raw samples -> prepare kernels -> fused rollout -> torch SGD on the module's
    flat views) against the same epoch through the host QuadDataset batches: tests/test_zz_new_paths_gpu.py::
    test_device_dataset_epoch_equals_host_dataset_epoch in miniature on the model library"""
    from apg_trajectory_tracking_b200 import device_data as DD, train as T
    from apg_trajectory_tracking_b200.neural_control import dataset as DS
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    n, h, dt, bs = 150, 10, 0.1, 64
    raw = SY.quad_case(n, h, dt, seed=2)
    off = torch.randn(n, 3, generator=torch.Generator().manual_seed(0))
    states, refs = raw["cur"].clone(), raw["ref"].clone()
    states[:, :3] += off
    refs[:, :, :3] += off[:, None]
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        nets.append(Net(15, h, 9, 4 * h, conv=1))
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    ma, mb = T.ModuleRollout(nets[0], spec, "cpu"), T.ModuleRollout(nets[1], spec, "cpu")
    oa = torch.optim.SGD(nets[0].parameters(), lr=1e-5, momentum=0.9)
    ob = torch.optim.SGD(nets[1].parameters(), lr=1e-5, momentum=0.9)
    host = DS.QuadDataset(states.numpy(), refs.numpy())
    dev = DD.DeviceQuadDataset(states, refs, "cpu")
    tot, i = 0.0, 0
    for i, lo in enumerate(range(0, n, bs)):
        b = [t[lo:lo + bs].clone() for t in (host.normed_states, host.states, host.in_ref_states, host.ref_states)]
        oa.zero_grad()
        tot += float(ma.loss_and_grad(*b))
        oa.step()
    la = tot / max(i, 1)
    lb = DD.run_epoch_device(mb, ob, dev, bs, shuffle=False)
    assert abs(la - lb) <= 2e-5 * abs(la), (la, lb)
    assert rel_err(mb.flat, ma.flat) <= 1e-6
    d3 = DD.DeviceQuadDataset.from_trajectory(torch.randn(301, 9), h, device="cpu")
    assert len(d3) == len(range(0, 301 - (h + 1), 2 * h))


def test_trainer_rollout_uses_the_physics_of_the_train_dynamics_it_was_given(simlib, monkeypatch):
    """the reference differentiates through ``self.train_dynamics`` (train_drone.py:186-190).  A trainer
    built on ``FlightmareDynamics(modified_params=...)`` - with NO ``modified_params`` entry in its config, or with an
    entry that describes the evaluation dynamics instead - must roll out the MODIFIED model: its fused loss equals the
    oracle's with the same constants and differs from the default-physics loss; learnt dynamics with a recurrent
    train mode are refused instead of silently integrating the analytic model."""
    from apg_trajectory_tracking_b200.scripts import train_drone as TD
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_trained import LearntDynamics
    from apg_trajectory_tracking_b200.neural_control import environments as ENV
    import importlib
    monkeypatch.setattr(ENV, "compute_device", lambda: torch.device("cpu"), raising=False)
    for name in ("models.hutter_model", "dynamics.quad_dynamics_flightmare", "dynamics.quad_dynamics_trained", "dataset"):
        mod = importlib.import_module("apg_trajectory_tracking_b200.neural_control." + name)
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None, raising=False)
    monkeypatch.setenv("APG_LEGACY_MMA", "1")             # (the smaller CPU-model run; the constants reach both paths)
    mp = {"translational_drag": [0.3, 0.3, 0.3], "kinv_ang_vel_tau": [20.0, 20.0, 6.0]}   # (the mass cancels in this model)
    # the reference's configs/quad_config.json, minus its modified_params entry
    config = dict(nr_epochs=1, delta_t=0.1, delta_t_train=0.1, epoch_size=64, self_play=0, self_play_every_x=2,
                  batch_size=64, reset_strength=1.2, max_drone_dist=0.25, max_steps=1000, thresh_div_start=0.1,
                  thresh_div_end=2, thresh_stable_start=1, thresh_stable_end=2, state_size=12, horizon=10,
                  train_mode="concurrent", ref_dim=9, action_dim=4, l2_lambda=0.01, learning_rate_controller=1e-5,
                  learning_rate_dynamics=0.001, resample_every=3, suc_up_down=1, speed_factor=0.5, system="quad",
                  sample_in="train_env", device="cpu")  # the "device" memory of the CPU-model library
    n, h, dt = 64, 10, 0.1
    case = SY.quad_case(n, h, dt, seed=11)
    losses = {}
    for key, dyn, cfg_extra in (("default", FlightmareDynamics(), {}),
                                ("modified", FlightmareDynamics(modified_params=mp), {}),
                                ("modified_cfg_is_eval", FlightmareDynamics(modified_params=mp),
                                 {"modified_params": {"translational_drag": [0.9, 0.9, 0.9]}})):
        tr = TD.TrainDrone(dyn, FlightmareDynamics(), dict(config, **cfg_extra))
        torch.manual_seed(5)
        tr.initialize_model()
        assert (tr.modified_params()["translational_drag"] == [0.3, 0.3, 0.3]) == (key != "default")
        params = [p.detach().clone() for p in tr.net.parameters()]
        loss = tr.fused.loss_and_grad(case["in_state"], case["cur"], case["in_ref"], case["ref"])
        losses[key] = (float(loss), params)
    assert abs(losses["modified"][0] - losses["default"][0]) > 1e-3 * abs(losses["default"][0])
    assert losses["modified_cfg_is_eval"][0] == losses["modified"][0]
    cfg_mod = dict(O.QUAD_CFG, **mp)
    monkeypatch.setitem(O.STEP_FN, "quad", lambda s, a, d: O.quad_step(s, a, d, cfg=cfg_mod))
    want, _, _, _ = O.concurrent_value_and_grad("quad", losses["modified"][1], case["in_state"], case["cur"],
                                                case["in_ref"], case["ref"], h, dt)
    assert abs(losses["modified"][0] - float(want)) <= 2e-5 * abs(float(want))
    with pytest.raises(ValueError):
        TD.TrainDrone(LearntDynamics(), FlightmareDynamics(), dict(config, train_mode="autoregressive")).modified_params()


def test_trainer_epoch_through_learnt_dynamics_is_the_fused_rollout(simlib, monkeypatch):
    """``TrainDrone.run_epoch`` with ``train_dynamics = LearntDynamics`` (the controller phase of the
    reference's run_dynamics, train_drone.py:260-278) takes the FUSED rollout through the learnt steps
    (apg_rollout_forward_learnt), follows the oracle's training on the same batches, and the un-fused per-step loop
    (config ``unfused_learnt_rollout``) gives the same epoch loss"""
    from apg_trajectory_tracking_b200.scripts import train_drone as TD
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control import environments as ENV, dataset as DS
    import importlib
    monkeypatch.setattr(ENV, "compute_device", lambda: torch.device("cpu"), raising=False)
    for name in ("models.hutter_model", "dynamics.quad_dynamics_flightmare", "dynamics.quad_dynamics_trained", "dataset",
                 "drone_loss"):
        mod = importlib.import_module("apg_trajectory_tracking_b200.neural_control." + name)
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None, raising=False)
    g = load_golden("learnt_dyn.npz")
    n, h, dt, bs, lr = 128, 10, 0.1, 64, 1e-5
    raw = SY.quad_case(n, h, dt, seed=4)
    config = dict(delta_t=dt, horizon=h, ref_dim=9, action_dim=4, state_size=12, batch_size=bs, system="quad",
                  learning_rate_controller=lr, train_mode="concurrent", device="cpu")
    curves, calls = [], []
    real = simlib.apg_rollout_forward_learnt

    def counted(*a):
        calls.append(1)
        return real(*a)
    monkeypatch.setattr(simlib, "apg_rollout_forward_learnt", counted, raising=False)
    for unfused in (False, True):
        _, _, d, lparams, ocfg = _learnt_case(1, 0)
        tr = TD.TrainDrone(d, FlightmareDynamics(), dict(config, unfused_learnt_rollout=unfused))
        torch.manual_seed(0)
        tr.initialize_model(state_data=DS.QuadDataset(raw["cur"].numpy(), raw["ref"].numpy()))
        tr.trainloader = torch.utils.data.DataLoader(tr.state_data, batch_size=bs, shuffle=False)
        if not unfused:
            p0 = [p.detach().clone() for p in tr.net.parameters()]
        before = len(calls)
        curves.append(tr.run_epoch(epoch=0))
        assert (len(calls) - before) == (0 if unfused else n // bs)
        if not unfused:
            trained = [p.detach().clone() for p in tr.net.parameters()]
    ds = DS.QuadDataset(raw["cur"].numpy(), raw["ref"].numpy())
    ps, bufs, run = [p.clone() for p in p0], [None] * len(p0), 0.0
    for b in range(n // bs):
        ins, cur, inr, ref = (torch.stack(x) for x in zip(*[ds[i] for i in range(b * bs, (b + 1) * bs)]))
        loss, grads, _, _ = O.value_and_grad(
            lambda q: O.rollout_concurrent_learnt(q, lparams, ins, cur, inr, ref, h, dt, ocfg), ps)
        ps, bufs = O.sgd_momentum_step(ps, grads, bufs, lr)
        run += float(loss)
    want = run / (n // bs - 1)
    assert abs(curves[0] - want) <= 2e-5 * abs(want) and abs(curves[1] - want) <= 2e-5 * abs(want), (curves, want)
    for a, w in zip(trained, ps):
        assert rel_err(a, w) <= 1e-4
