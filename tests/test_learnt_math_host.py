"""Per-drone math of the learnt residual quadrotor dynamics (csrc/learnt_math.cuh) compiled with g++: forward against
the golden outputs of the reference's LearntDynamics, hand-written adjoint (state, action, every parameter) against
the reference's autograd (tests/golden/learnt_dyn.npz) and against fp64 autograd of the oracle."""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import apg_oracle as O
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("apg_params", os.path.join(ROOT, "apg_trajectory_tracking_b200",
                                                                             "params.py"))
P = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(P)


@pytest.fixture(scope="module")
def hl(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_learnt") / "libhostcheck_learnt.so"
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck_learnt.cpp")
    inc = os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I", inc,
                           src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _case(tag, dtype):
    g = load_golden("learnt_dyn.npz")
    flat = np.concatenate([np.asarray(g[f"{tag}_param_{i}"]).reshape(-1) for i in range(8)]).astype(dtype)
    pc = P.PHYS["quad"]({"rotational_drag": [float(x) for x in g[f"{tag}_rot_drag"]]})
    arr = lambda k: np.ascontiguousarray(g[f"{tag}_{k}"], dtype=dtype)          # noqa: E731
    return g, flat, pc, arr("state"), arr("action"), arr("cot"), float(g[f"{tag}_dt"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_forward_and_adjoint_match_reference_fp32(hl, tag):
    g, flat, pc, s, a, cot, dt = _case(tag, np.float32)
    assert hl.hc_learnt_num_params() == flat.size == 1891
    n = s.shape[0]
    out = np.zeros_like(s)
    hl.hc_learnt_fwd_f32(_p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(dt), n, _p(out))
    assert np.abs(out - g[f"{tag}_out"]).max() <= 5e-6 * np.abs(g[f"{tag}_out"]).max()
    gs, ga, gp = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_adj_f32(_p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs), _p(ga), _p(gp))
    assert np.abs(gs - g[f"{tag}_gstate"]).max() <= 2e-5 * np.abs(g[f"{tag}_gstate"]).max()
    assert np.abs(ga - g[f"{tag}_gaction"]).max() <= 2e-5 * np.abs(g[f"{tag}_gaction"]).max()
    want = np.concatenate([np.asarray(g[f"{tag}_gparam_{i}"]).reshape(-1) for i in range(8)])
    off = np.cumsum([0] + [np.asarray(g[f"{tag}_param_{i}"]).size for i in range(8)])
    scale = np.abs(want).max()
    for i in range(8):
        got_i, want_i = gp[off[i]:off[i + 1]], want[off[i]:off[i + 1]]
        if i == 1 or (i == 2 and np.abs(g[f"{tag}_rot_drag"]).max() == 0):
            assert np.abs(got_i).max() == 0.0 and np.abs(want_i).max() <= 1e-4       # analytic zero vs rounding noise
        else:
            assert np.abs(got_i - want_i).max() <= 5e-5 * max(np.abs(want_i).max(), 1e-3 * scale), i


@pytest.mark.parametrize("tag", ["a", "b"])
def test_adjoint_matches_oracle_autograd_fp64(hl, tag):
    g, flat, pc, s, a, cot, dt = _case(tag, np.float64)
    n = s.shape[0]
    gs, ga, gp = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_adj_f64(_p(flat), _p(pc), _p(s), _p(a), ctypes.c_double(dt), n, _p(cot), _p(gs), _p(ga), _p(gp))
    lparams = [torch.tensor(np.asarray(g[f"{tag}_param_{i}"]), dtype=torch.float64, requires_grad=True)
               for i in range(8)]
    # the harness reads the simulator constants as fp32 (like the kernels): give the oracle the same rounded values
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(np.float32(x)) for x in g[f"{tag}_rot_drag"]))
    with torch.no_grad():
        lparams[2].copy_(torch.tensor([float(pc[P_J]) for P_J in (1, 2, 3)], dtype=torch.float64))
        lparams[3].copy_(torch.tensor([float(pc[k]) for k in (4, 5, 6)], dtype=torch.float64))
    ts, ta = torch.tensor(s, requires_grad=True), torch.tensor(a, requires_grad=True)
    out = O.learnt_quad_step(lparams, ts, ta, dt, cfg)
    grads = torch.autograd.grad(out, [ts, ta] + lparams, torch.tensor(cot), allow_unused=True)
    assert np.abs(gs - grads[0].numpy()).max() <= 1e-9 * np.abs(gs).max()
    assert np.abs(ga - grads[1].numpy()).max() <= 1e-9 * np.abs(ga).max()
    off = np.cumsum([0] + [p.numel() for p in lparams])
    scale = max(float(x.abs().max()) for x in grads[2:] if x is not None)
    for i in range(8):
        want = grads[2 + i].numpy().reshape(-1)
        tol = 1e-9 * max(np.abs(want).max(), 1e-3 * scale)
        if i in (1, 2):
            tol = 1e-6 * scale      # mass / inertia: autograd's cancelling terms leave fp64 rounding noise
        assert np.abs(gp[off[i]:off[i + 1]] - want).max() <= tol, i


def test_tiled_factor_rows_reproduce_the_parameter_gradient(hl):
    """the adjoint kernel's tile / factor-row / entry mapping (learnt_entry_rows) against the direct outer products"""
    rng = np.random.default_rng(1)
    g, flat, pc, _, _, _, dt = _case("b", np.float32)
    n = 300                                               # 3 tiles of 128 with a ragged tail
    s = (0.4 * rng.standard_normal((n, 12))).astype(np.float32)
    a = rng.random((n, 4)).astype(np.float32)
    cot = rng.standard_normal((n, 12)).astype(np.float32)
    gs1, ga1, gp1 = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    gs2, ga2, gp2 = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_adj_f32(_p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs1), _p(ga1), _p(gp1))
    hl.hc_learnt_adj_tiled_f32(_p(flat), _p(pc), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs2), _p(ga2),
                               _p(gp2), 128)
    assert np.array_equal(gs1, gs2) and np.array_equal(ga1, ga2)
    assert np.abs(gp1 - gp2).max() <= 2e-5 * np.abs(gp1).max()
    assert np.abs(gp2[16:17]).max() == 0.0 and np.abs(gp2).min() >= 0 and np.count_nonzero(gp2) > 1800


# ---------------------------------------------------------------------------------------------------------------
# fixed wing: LearntFixedWingDynamics (every physical constant a live parameter, general 3x3 inertia matrix)
# ---------------------------------------------------------------------------------------------------------------
def _wing_case(tag, dtype):
    g = load_golden("learnt_dyn.npz")
    flat = np.concatenate([np.asarray(g[f"{tag}_param_{i}"]).reshape(-1) for i in range(42)]).astype(dtype)
    arr = lambda k: np.ascontiguousarray(g[f"{tag}_{k}"], dtype=dtype)          # noqa: E731
    return g, flat, arr("state"), arr("action"), arr("cot"), float(g[f"{tag}_dt"])


@pytest.mark.parametrize("tag", ["wa", "wb"])
def test_wing_forward_and_adjoint_match_reference_fp32(hl, tag):
    g, flat, s, a, cot, dt = _wing_case(tag, np.float32)
    assert hl.hc_learnt_wing_num_params() == flat.size == 1914
    assert [str(x) for x in g["wing_param_names"]][1:38] == ["cfg." + k for k in O.WING_LEARNT_KEYS]
    n = s.shape[0]
    out = np.zeros_like(s)
    hl.hc_learnt_wing_fwd_f32(_p(flat), _p(s), _p(a), ctypes.c_float(dt), n, _p(out))
    assert np.abs(out - g[f"{tag}_out"]).max() <= 5e-6 * np.abs(g[f"{tag}_out"]).max()
    gs, ga, gp = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_wing_adj_f32(_p(flat), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs), _p(ga), _p(gp))
    assert np.abs(gs - g[f"{tag}_gstate"]).max() <= 5e-5 * np.abs(g[f"{tag}_gstate"]).max()
    assert np.abs(ga - g[f"{tag}_gaction"]).max() <= 5e-5 * np.abs(g[f"{tag}_gaction"]).max()
    off = np.cumsum([0] + [np.asarray(g[f"{tag}_param_{i}"]).size for i in range(42)])
    names = [str(x) for x in g["wing_param_names"]]
    for i in range(42):
        got, want = gp[off[i]:off[i + 1]], np.asarray(g[f"{tag}_gparam_{i}"]).reshape(-1)
        if names[i] == "cfg.g":
            assert np.abs(got).max() == 0.0 and np.abs(want).max() == 0.0            # detached in the reference
        else:
            # fp32 sums over 29 drones of terms of mixed sign: relative to the tensor's own scale
            assert np.abs(got - want).max() <= 2e-4 * max(np.abs(want).max(), 1e-2), names[i]


@pytest.mark.parametrize("tag", ["wa", "wb"])
def test_wing_adjoint_matches_oracle_autograd_fp64(hl, tag):
    g, flat, s, a, cot, dt = _wing_case(tag, np.float64)
    n = s.shape[0]
    out = np.zeros_like(s)
    hl.hc_learnt_wing_fwd_f64(_p(flat), _p(s), _p(a), ctypes.c_double(dt), n, _p(out))
    gs, ga, gp = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_wing_adj_f64(_p(flat), _p(s), _p(a), ctypes.c_double(dt), n, _p(cot), _p(gs), _p(ga), _p(gp))
    lparams = [torch.tensor(np.asarray(g[f"{tag}_param_{i}"]), dtype=torch.float64, requires_grad=True)
               for i in range(42)]
    ts, ta = torch.tensor(s, requires_grad=True), torch.tensor(a, requires_grad=True)
    want_out = O.learnt_wing_step(lparams, ts, ta, dt)
    assert np.abs(out - want_out.detach().numpy()).max() <= 1e-12 * np.abs(out).max()
    grads = torch.autograd.grad(want_out, [ts, ta] + lparams, torch.tensor(cot), allow_unused=True)
    assert np.abs(gs - grads[0].numpy()).max() <= 1e-10 * np.abs(gs).max()
    assert np.abs(ga - grads[1].numpy()).max() <= 1e-10 * np.abs(ga).max()
    off = np.cumsum([0] + [p.numel() for p in lparams])
    for i in range(42):
        want = (grads[2 + i] if grads[2 + i] is not None else torch.zeros_like(lparams[i])).numpy().reshape(-1)
        assert np.abs(gp[off[i]:off[i + 1]] - want).max() <= 1e-10 * max(np.abs(want).max(), 1e-3), i


def test_wing_tiled_factor_rows_reproduce_the_parameter_gradient(hl):
    rng = np.random.default_rng(2)
    g, flat, s0, _, _, dt = _wing_case("wb", np.float32)
    n = 300
    s = np.repeat(s0, 11, axis=0)[:n].copy()
    s += (0.02 * rng.standard_normal(s.shape)).astype(np.float32)
    a = rng.random((n, 4)).astype(np.float32)
    cot = rng.standard_normal((n, 12)).astype(np.float32)
    gs1, ga1, gp1 = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    gs2, ga2, gp2 = np.zeros_like(s), np.zeros_like(a), np.zeros_like(flat)
    hl.hc_learnt_wing_adj_f32(_p(flat), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs1), _p(ga1), _p(gp1))
    hl.hc_learnt_wing_adj_tiled_f32(_p(flat), _p(s), _p(a), ctypes.c_float(dt), n, _p(cot), _p(gs2), _p(ga2), _p(gp2),
                                    128)
    assert np.array_equal(gs1, gs2) and np.array_equal(ga1, ga2)
    off = [0, 9, 46, 46 + 1024, 46 + 1088, 46 + 1088 + 768, 1914]
    for lo, hi in zip(off, off[1:]):
        assert np.abs(gp1[lo:hi] - gp2[lo:hi]).max() <= 5e-5 * max(np.abs(gp1[lo:hi]).max(), 1e-3)
