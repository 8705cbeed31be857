"""Per-drone logic of the closed-loop evaluation kernel (csrc/eval_math.cuh: reference window, divergence, stability,
stop / reset) compiled with g++ and driven like the kernel's per-thread code, against the golden runs of the
reference's own QuadEvaluator.follow_trajectory("rand") (tests/golden/eval_rand.npz) and against the CPU oracle."""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import apg_oracle as O
from tests.helpers import golden_params, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("apg_params", os.path.join(ROOT, "apg_trajectory_tracking_b200",
                                                                             "params.py"))
P = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(P)


@pytest.fixture(scope="module")
def he(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_eval") / "libhostcheck_eval.so"
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck_eval.cpp")
    inc = os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I", inc,
                           src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _run(he, params, tables, index, init, steps, h, dt, tdiv, tstab, test_time):
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    tables = np.ascontiguousarray(tables, dtype=np.float32)
    init = np.ascontiguousarray(init, dtype=np.float32)
    n, rl = init.shape[0], tables.shape[1]
    idx = None if index is None else np.ascontiguousarray(index, dtype=np.int32)
    states = np.zeros((n, steps + 1, 12), np.float32)
    div, act = np.zeros((n, steps), np.float32), np.zeros((n, steps, 4), np.float32)
    nst = np.zeros(n, np.int32)
    pc = P.PHYS["quad"]()
    he.hc_eval_rollout(_p(flat), h, params[-1].shape[0], _p(tables), _p(idx), rl, _p(init), n, steps,
                       ctypes.c_float(dt), _p(pc), ctypes.c_float(tdiv), ctypes.c_float(tstab), int(test_time),
                       _p(states), _p(div), _p(act), _p(nst))
    return states, div, act, nst


@pytest.mark.parametrize("name", ["gentle", "fast_reset", "fast_stop", "short_table", "tight"])
def test_kernel_logic_matches_reference_evaluator(he, name):
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    steps, test_time, h = int(steps), int(test_time), int(h)
    ref_states = g[f"{name}_states"]
    states, div, act, nst = _run(he, params, g[f"{name}_table"][None], None, ref_states[:1], steps, h, dt, tdiv, tstab,
                                 test_time)
    taken = len(g[f"{name}_div"])
    assert int(nst[0]) == taken
    assert np.abs(states[0, :taken + 1] - ref_states).max() <= 2e-5
    assert np.abs(div[0, :taken] - g[f"{name}_div"]).max() <= 2e-5
    assert np.abs(act[0, :taken] - g[f"{name}_actions"]).max() <= 2e-5
    assert np.abs(states[0, taken + 1:]).sum() == 0


def test_kernel_logic_matches_oracle_batched_with_table_index(he):
    """several drones sharing tables through the index array, different start offsets, stop mode"""
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    tabs = np.stack([g["gentle_table"][:100], g["tight_table"][:100]]).astype(np.float32)
    index = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    rng = np.random.default_rng(0)
    init = np.zeros((5, 12), np.float32)
    init[:, :3] = tabs[index, 0, :3] + rng.normal(0, 0.05, (5, 3))
    init[:, 6:9] = rng.normal(0, 0.1, (5, 3))
    states, div, act, nst = _run(he, params, tabs, index, init, 50, 10, 0.1, 0.6, 0.5, 1)
    out = O.eval_follow_tables(params, torch.tensor(tabs)[torch.tensor(index, dtype=torch.long)], torch.tensor(init),
                               50, 10, 0.1, 0.6, 0.5, 1)
    assert np.array_equal(nst, out["n_steps"].numpy())
    assert len(set(nst.tolist())) > 1                     # the drones stop at different steps
    assert np.abs(states - out["states"].numpy()).max() <= 2e-5
    assert np.abs(div - out["div"].numpy()).max() <= 2e-5


# ---------------------------------------------------------------------------------------------------------------
# fixed wing: FixedWingEvaluator.fly_to_point
# ---------------------------------------------------------------------------------------------------------------
def _run_wing(he, g, params, targets, init, steps, h, dt_data, dt_env, tdiv, tstab, test_time):
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    targets = np.ascontiguousarray(targets, dtype=np.float32)
    init = np.ascontiguousarray(init, dtype=np.float32)
    n, K = targets.shape[0], targets.shape[1]
    mean, std = np.ascontiguousarray(g["mean"], np.float32), np.ascontiguousarray(g["std"], np.float32)
    states = np.zeros((n, steps + 1, 12), np.float32)
    div, act = np.zeros((n, steps), np.float32), np.zeros((n, steps, 4), np.float32)
    nst, dts, dtc = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    pc = P.PHYS["wing"]()
    he.hc_eval_wing(_p(flat), h, _p(targets), K, _p(init), n, _p(mean), _p(std), ctypes.c_float(dt_data),
                    ctypes.c_float(dt_env), _p(pc), steps, ctypes.c_float(tdiv), ctypes.c_float(tstab), int(test_time),
                    _p(states), _p(div), _p(act), _p(nst), _p(dts), _p(dtc))
    return states, div, act, nst, dts, dtc


@pytest.mark.parametrize("name", ["one_target", "two_targets", "tight_reset", "tight_stop", "unstable", "step_limit"])
def test_wing_kernel_logic_matches_reference_evaluator(he, name):
    from tests.test_oracle_golden import wing_eval_case
    g = load_golden("eval_wing.npz")
    params, targets, init, h, dt_data, dt_env, steps, test_time, tdiv, tstab = wing_eval_case(g, name)
    states, div, act, nst, dts, dtc = _run_wing(he, g, params, targets.numpy(), init.numpy(), steps, h, dt_data, dt_env,
                                                tdiv, tstab, test_time)
    traj, dl, dtg = g[f"{name}_traj"], g[f"{name}_div_linear"], g[f"{name}_div_target"]
    taken = len(dl)
    assert int(nst[0]) == taken
    scale = np.abs(traj[:, :12]).max()
    assert np.abs(states[0, 1:taken + 1] - traj[:, :12]).max() <= 5e-5 * scale
    assert np.abs(act[0, :taken] - traj[:, 12:]).max() <= 5e-5
    assert np.abs(div[0, :taken] - dl).max() <= 5e-5 * max(dl.max(), 1.0)
    assert int(dtc[0]) == len(dtg) and abs(float(dts[0]) - dtg.sum()) <= 2e-4 * max(dtg.sum(), 1.0)


# ---- cartpole: Evaluator.evaluate_in_environment (tests/golden/eval_cartpole.npz) -------------------------------------
CARTPOLE_EVAL_RUNS = ["zero_start", "tilted", "falls", "tight", "falls_at_once"]


def _run_cartpole(he, params, init, steps, dt, tdiv, burn):
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    init = np.ascontiguousarray(init, dtype=np.float32)
    n = init.shape[0]
    states, act = np.zeros((n, steps, 4), np.float32), np.zeros((n, steps), np.float32)
    nst = np.zeros(n, np.int32)
    asum, acnt, vsum = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    pc = P.PHYS["cartpole"]()
    he.hc_eval_cartpole(_p(flat), params[-1].shape[0], _p(init), n, ctypes.c_float(dt), _p(pc), steps,
                        ctypes.c_float(tdiv), int(burn), _p(states), _p(act), _p(nst), _p(asum), _p(acnt), _p(vsum))
    return states, act, nst, asum, acnt, vsum


@pytest.mark.parametrize("name", CARTPOLE_EVAL_RUNS)
def test_cartpole_kernel_logic_matches_reference_evaluator(he, name):
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    steps, tdiv, burn = g[f"{name}_cfg"]
    states, act, nst, asum, acnt, vsum = _run_cartpole(he, params, g[f"{name}_init"][None], int(steps), 0.05, tdiv,
                                                       int(burn))
    want = g[f"{name}_states"]
    taken = len(want)
    assert int(nst[0]) == taken and int(nst[0]) - 1 == int(g[f"{name}_success"][0])
    assert np.abs(states[0, :taken] - want).max() <= 2e-5 * max(np.abs(want).max(), 1.0)
    assert np.abs(states[0, taken:]).sum() == 0
    assert abs(float(vsum[0]) - g[f"{name}_vel"].sum()) <= 1e-4 * max(g[f"{name}_vel"].sum(), 1.0)
    late = np.abs(want[int(burn) + 1:, 2])
    assert int(acnt[0]) == len(late)
    assert abs(float(asum[0]) - late.sum()) <= 1e-4 * max(late.sum(), 1.0)
    # from the second step on the cart position the environment integrates starts from 0 (x' = 0 + x_dot * dt)
    if taken > 2:
        assert np.abs(want[2:, 0] - want[1:-1, 1] * 0.05).max() <= 1e-6


def test_cartpole_kernel_logic_matches_oracle_batched(he):
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    rng = np.random.default_rng(1)
    init = np.concatenate([np.stack([g[f"{n}_init"] for n in CARTPOLE_EVAL_RUNS]),
                           rng.uniform(-1, 1, (27, 4)) * np.array([0.5, 1.5, 0.12, 1.8])]).astype(np.float32)
    states, act, nst, asum, acnt, vsum = _run_cartpole(he, params, init, 60, 0.05, 0.21, 5)
    out = O.eval_cartpole_balance(params, torch.tensor(init), 60, 0.05, 0.21, 5)
    assert np.array_equal(nst, out["n_steps"].numpy())
    assert 1 < len(set(nst.tolist()))                                   # runs of different length in one batch
    assert np.abs(states - out["states"].numpy()).max() <= 5e-5
    assert np.abs(act - out["actions"].numpy()).max() <= 5e-5
    assert np.abs(vsum - out["vel_sum"].numpy()).max() <= 1e-3
    mean_angle = np.where(acnt > 0, asum / np.maximum(acnt, 1), 100.0)
    assert np.abs(mean_angle - out["mean_angle"].numpy()).max() <= 1e-4
