"""The evaluation kernels THEMSELVES on the CPU: csrc/eval_kernels.cu (unchanged source, -DAPG_SIM) compiled with g++ on
top of the software model of tests/hostcheck/te_sim.h (one OS thread per GPU thread; mma.sync fragments, TMA bulk
copies, mbarriers, __syncthreads_or in software), against the golden runs of the reference's own evaluators and the
oracle.  Covers what the per-drone host checks (tests/test_eval_math_host.py) cannot: shared-memory carve-up, tile
engine call sites, barriers / early exit, packed-weight layouts, output addressing of whole launches."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from apg_trajectory_tracking_b200 import params as P
from oracle import apg_oracle as O
from tests.helpers import golden_params, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def te(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_tesim") / "libhostcheck_tesim.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck_tesim.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _flat(params):
    return np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)


def test_cartpole_kernel_on_the_model_matches_reference_runs(te):
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    names = ["tilted", "falls", "falls_at_once", "zero_start"]
    steps, n = 12, 70                                                   # two tiles, the second partial
    rng = np.random.default_rng(0)
    init = np.concatenate([np.stack([g[f"{k}_init"] for k in names]),
                           rng.uniform(-1, 1, (n - 4, 4)) * np.array([0.5, 1.5, 0.12, 1.8])]).astype(np.float32)
    states, act = np.zeros((n, steps, 4), np.float32), np.zeros((n, steps), np.float32)
    nst = np.zeros(n, np.int32)
    asum, acnt, vsum = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    err = ctypes.create_string_buffer(2048)
    nerr = te.hc_tesim_eval_cartpole(_p(_flat(params)), 10, _p(init), n, ctypes.c_float(0.05), _p(P.PHYS["cartpole"]()),
                                     steps, ctypes.c_float(0.21), 5, 2, _p(states), _p(act), _p(nst), _p(asum),
                                     _p(acnt), _p(vsum), err, 2048)
    assert nerr == 0, err.value.decode()
    for k, name in enumerate(names):
        want = g[f"{name}_states"][:steps]
        taken = min(len(g[f"{name}_states"]), steps)
        assert int(nst[k]) == taken
        assert np.abs(states[k, :taken] - want[:taken]).max() <= 2e-5
    out = O.eval_cartpole_balance(params, torch.tensor(init), steps, 0.05, 0.21, 5)
    assert np.array_equal(nst, out["n_steps"].numpy()) and len(set(nst.tolist())) > 1
    assert np.abs(states - out["states"].numpy()).max() <= 5e-5
    assert np.abs(vsum - out["vel_sum"].numpy()).max() <= 1e-3


@pytest.mark.slow
def test_quad_eval_kernel_on_the_model_matches_reference_run_with_reset(te):
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    name = "fast_reset"
    steps_all, test_time, tdiv, tstab, h, dt = [float(v) for v in g[f"{name}_cfg"]]
    h = int(h)
    steps = 34                                                           # the first reset happens at step 29
    tabs = np.ascontiguousarray(np.stack([g[f"{name}_table"], g["gentle_table"]]), np.float32)
    n = 66                                                               # two tiles (64 + 2 drones)
    index = np.array([0, 1] * 33, np.int32)
    rng = np.random.default_rng(1)
    init = np.zeros((n, 12), np.float32)
    init[:, :3] = tabs[index, 0, :3] + rng.normal(0, 0.03, (n, 3))
    init[0], init[1] = g[f"{name}_states"][0], g["gentle_states"][0]
    states = np.zeros((n, steps + 1, 12), np.float32)
    div, act = np.zeros((n, steps), np.float32), np.zeros((n, steps, 4), np.float32)
    nst = np.zeros(n, np.int32)
    err = ctypes.create_string_buffer(2048)
    nerr = te.hc_tesim_eval_rollout(_p(_flat(params)), h, 4 * h, _p(tabs), _p(index), tabs.shape[1], _p(init), n, steps,
                                    ctypes.c_float(dt), _p(P.PHYS["quad"]()), ctypes.c_float(tdiv),
                                    ctypes.c_float(tstab), 0, 2, _p(states), _p(div), _p(act), _p(nst), err, 2048)
    assert nerr == 0, err.value.decode()
    assert (nst == steps).all()
    for k, nm in ((0, name), (1, "gentle")):
        assert np.abs(states[k] - g[f"{nm}_states"][:steps + 1]).max() <= 5e-5, nm
        assert np.abs(div[k] - g[f"{nm}_div"][:steps]).max() <= 5e-5
        assert np.abs(act[k] - g[f"{nm}_actions"][:steps]).max() <= 5e-5
    assert (g[f"{name}_div"][:steps] > tdiv).sum() >= 1                  # the run contains a reset
    want = O.eval_follow_tables(params, torch.tensor(tabs)[index.astype(np.int64)], torch.tensor(init), steps, h, dt,
                                tdiv, tstab, 0)
    assert np.abs(states - want["states"].numpy()).max() <= 2e-4


@pytest.mark.slow
def test_wing_eval_kernel_on_the_model_matches_reference_flight(te):
    from tests.test_oracle_golden import wing_eval_case
    g = load_golden("eval_wing.npz")
    params, targets, init1, h, dt_data, dt_env, steps_all, test_time, tdiv, tstab = wing_eval_case(g, "two_targets")
    steps, n = 70, 65                                                    # the target switch happens around step 50
    tg = np.ascontiguousarray(np.repeat(targets.numpy(), n, 0), np.float32)
    rng = np.random.default_rng(2)
    tg[1:, :, 1:] += rng.uniform(-1, 1, (n - 1, tg.shape[1], 2)).astype(np.float32)
    init = np.zeros((n, 12), np.float32)
    init[:, 3] = 11.5
    states = np.zeros((n, steps + 1, 12), np.float32)
    div, act = np.zeros((n, steps), np.float32), np.zeros((n, steps, 4), np.float32)
    nst = np.zeros(n, np.int32)
    dts, dtc = np.zeros(n, np.float32), np.zeros(n, np.float32)
    mean, std = np.ascontiguousarray(g["mean"], np.float32), np.ascontiguousarray(g["std"], np.float32)
    err = ctypes.create_string_buffer(2048)
    nerr = te.hc_tesim_eval_wing(_p(_flat(params)), h, _p(tg), tg.shape[1], _p(init), n, _p(mean), _p(std),
                                 ctypes.c_float(dt_data), ctypes.c_float(dt_env), _p(P.PHYS["wing"]()), steps,
                                 ctypes.c_float(tdiv), ctypes.c_float(tstab), 0, 2, _p(states), _p(div), _p(act),
                                 _p(nst), _p(dts), _p(dtc), err, 2048)
    assert nerr == 0, err.value.decode()
    traj = g["two_targets_traj"]
    scale = np.abs(traj[:, :12]).max()
    assert int(nst[0]) == steps
    assert np.abs(states[0, 1:steps + 1] - traj[:steps, :12]).max() <= 1e-4 * scale
    assert np.abs(act[0] - traj[:steps, 12:]).max() <= 1e-4
    assert dtc[0] >= 1                                                   # the first target was passed
    want = O.eval_fly_to_points(params, torch.tensor(tg), torch.tensor(init), g["mean"], g["std"], steps, h, dt_data,
                                dt_env, tdiv, tstab, 0)
    assert np.array_equal(nst, want["n_steps"].numpy())
    assert np.abs(states - want["states"].numpy()).max() <= 2e-4 * scale
    assert np.abs(dts - want["div_target_sum"].numpy()).max() <= 1e-3


# ---- learnt residual dynamics: csrc/learnt_kernels.cu (unchanged source) on the CPU thread model ------------------
@pytest.fixture(scope="module")
def ln(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_lnsim") / "libhostcheck_lnsim.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck_lnsim.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("system,tag,nparam_tensors", [(0, "b", 8), (1, "wb", 42)])
def test_learnt_kernels_on_the_model_match_reference_autograd(ln, system, tag, nparam_tensors):
    """forward + adjoint launches over several blocks / tiles (the golden batch tiled up to 300 rows): outputs and
    state / action gradients per row, parameter gradient = sum over the rows"""
    g = load_golden("learnt_dyn.npz")
    flat = np.concatenate([np.asarray(g[f"{tag}_param_{i}"]).reshape(-1) for i in range(nparam_tensors)])
    flat = np.ascontiguousarray(flat, np.float32)
    assert ln.hc_lnsim_num_params(system) == flat.size
    pc = P.PHYS["quad"]({"rotational_drag": [float(x) for x in g[f"{tag}_rot_drag"]]}) if system == 0 \
        else P.PHYS["wing"]()
    reps = 300 // len(g[f"{tag}_state"]) + 1
    tile = lambda k: np.ascontiguousarray(np.tile(g[f"{tag}_{k}"], (reps, 1))[:300], np.float32)   # noqa: E731
    s, a, cot = tile("state"), tile("action"), tile("cot")
    n, m = 300, len(g[f"{tag}_state"])
    out, gs, ga, gp = np.zeros_like(s), np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    err = ctypes.create_string_buffer(2048)
    nerr = ln.hc_lnsim_step_and_adjoint(system, _p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(float(g[f"{tag}_dt"])),
                                        n, 2, _p(cot), _p(out), _p(gs), _p(ga), _p(gp), err, 2048)
    assert nerr == 0, err.value.decode()
    rel = lambda x, w: np.abs(x - w).max() / max(np.abs(w).max(), 1e-6)          # noqa: E731
    assert rel(out[:m], g[f"{tag}_out"]) <= 5e-6 and np.array_equal(out[m:2 * m], out[:m])
    assert rel(gs[:m], g[f"{tag}_gstate"]) <= 5e-5 and rel(ga[:m], g[f"{tag}_gaction"]) <= 5e-5
    # parameter gradient of the tiled batch = sum over complete copies + the partial last copy: compare through the
    # oracle-free identity on the first `m` rows only when the batch is one exact multiple
    want = np.concatenate([np.asarray(g[f"{tag}_gparam_{i}"]).reshape(-1) for i in range(nparam_tensors)])
    full = (n // m) * m
    out2, gs2, ga2, gp2 = np.zeros_like(s), np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    nerr = ln.hc_lnsim_step_and_adjoint(system, _p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(float(g[f"{tag}_dt"])),
                                        full, 3, _p(cot), _p(out2), _p(gs2), _p(ga2), _p(gp2), err, 2048)
    assert nerr == 0, err.value.decode()
    scale = np.abs(want).max() * (full // m)
    assert np.abs(gp2 - want * (full // m)).max() <= 1e-4 * scale


# ---- the concurrent hutter rollout kernels on the model -------------------------------------------------------------
TM, TMP = 64, 68


def _build(tmp, name):
    out = tmp / f"lib{name}.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", f"{name}.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="module")
def hk(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("hostcheck_hksim")
    return _build(tmp, "hostcheck_hksim")


def _check_grad(grad, params, want_grad, tol):
    o = 0
    for i, (p, g) in enumerate(zip(params, want_grad)):
        got = grad[o:o + p.numel()].reshape(p.shape)
        o += p.numel()
        if g is None:
            assert np.abs(got).max() == 0
            continue
        scale = max(float(g.abs().max()), 1e-6)
        assert np.abs(got - g.detach().double().numpy()).max() <= tol * scale, i


@pytest.mark.slow
def test_hutter_kernels_on_the_model_reproduce_oracle_loss_and_gradient(hk):
    """hutter_fwd_kernel / hutter_adj_kernel are GPU-verified: that their unchanged source ALSO reproduces the oracle
    on the software model validates the model itself (warp specialisation, named barriers, mbarrier hand-offs, TMA
    bulk loads / stores with deferred reads, mma.sync fragment layout)."""
    import bench as B
    from apg_trajectory_tracking_b200 import synthetic as SY
    lib = hk
    n, h, grid = 150, 10, 2                                              # 3 tiles: CTA 0 gets two, the last is partial
    params = B.default_init("quad", h, seed=4)
    case = SY.quad_case(n, h, 0.1, seed=4)
    flat = _flat(params)
    f32 = lambda t: np.ascontiguousarray(t.numpy(), np.float32)                 # noqa: E731
    ins, cur, inr, ref = f32(case["in_state"]), f32(case["cur"]), f32(case["in_ref"]), f32(case["ref"])
    pc = P.PHYS["quad"]()
    nt = (n + TM - 1) // TM
    nan = lambda rows: np.full(nt * rows * TMP, np.nan, np.float32)             # noqa: E731
    x1, h1, h2, h3, act, sts = nan(224), nan(64), nan(64), nan(64), nan(40), nan(h * 12)
    lossp = np.zeros(grid, np.float32)
    err = ctypes.create_string_buffer(2048)
    common = (_p(flat), _p(ins), _p(cur), _p(inr), _p(ref), n, h, ctypes.c_float(0.1), _p(pc), grid, _p(x1), _p(h1),
              _p(h2), _p(h3), _p(act), _p(sts))
    assert lib.hc_hksim_forward(*common, _p(lossp), err, 2048) == 0, err.value.decode()
    want_loss, want_grad, _, _ = O.concurrent_value_and_grad("quad", params, case["in_state"], case["cur"],
                                                             case["in_ref"], case["ref"], h, 0.1)
    assert abs(float(lossp.sum()) - float(want_loss)) <= 2e-5 * abs(float(want_loss))
    npar = lib.hc_hksim_num_params(h)
    assert npar == flat.size
    parts = np.zeros((grid, npar), np.float32)
    assert lib.hc_hksim_adjoint(*common, _p(parts), err, 2048) == 0, err.value.decode()
    grad = np.zeros(npar, np.float32)
    lib.hc_hksim_reduce(_p(parts), grid, h, _p(grad))
    _check_grad(grad.astype(np.float64), params, want_grad, 5e-5)


def test_quad_eval_kernel_on_the_model_autoregressive_policy(te):
    """Net(15, h, 9, 4) (out_dim 4: the policy is called once per step and its whole output is the action)"""
    import bench as B
    g = load_golden("eval_rand.npz")
    h, dt, steps, n = 10, 0.1, 8, 40
    params = B.default_init("quad", h, seed=3, mode="autoregressive")
    tabs = np.ascontiguousarray(g["tight_table"][None, :60], np.float32)
    index = np.zeros(n, np.int32)
    rng = np.random.default_rng(4)
    init = np.zeros((n, 12), np.float32)
    init[:, :3] = tabs[0, 0, :3] + rng.normal(0, 0.05, (n, 3))
    init[:, 6:9] = rng.normal(0, 0.1, (n, 3))
    states = np.zeros((n, steps + 1, 12), np.float32)
    div, act = np.zeros((n, steps), np.float32), np.zeros((n, steps, 4), np.float32)
    nst = np.zeros(n, np.int32)
    err = ctypes.create_string_buffer(2048)
    nerr = te.hc_tesim_eval_rollout(_p(_flat(params)), h, 4, _p(tabs), _p(index), tabs.shape[1], _p(init), n, steps,
                                    ctypes.c_float(dt), _p(P.PHYS["quad"]()), ctypes.c_float(0.5), ctypes.c_float(0.4),
                                    1, 1, _p(states), _p(div), _p(act), _p(nst), err, 2048)
    assert nerr == 0, err.value.decode()
    want = O.eval_follow_tables(params, torch.tensor(tabs).repeat(n, 1, 1), torch.tensor(init), steps, h, dt, 0.5, 0.4, 1)
    assert np.array_equal(nst, want["n_steps"].numpy())
    assert np.abs(states - want["states"].numpy()).max() <= 1e-4
    assert np.abs(act - want["actions"].numpy()).max() <= 1e-4
