"""The per-drone math that the CUDA kernels inline (csrc/apg_math.cuh) compiled with g++ and checked on the CPU:
forward steps against the golden vectors of the reference, hand-written adjoints against fp64 autograd of the
oracle.  This validates the formulas before any GPU time is spent; the GPU parity tests (-m gpu) validate the
kernels themselves."""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import apg_oracle as O
from tests.helpers import load_golden, max_rel_to_scale

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("apg_params", os.path.join(ROOT, "apg_trajectory_tracking_b200",
                                                                             "params.py"))
P = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(P)


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so"
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    inc = os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I", inc,
                           src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


SYS = {"quad": (O.quad_step, 12, 4), "wing": (O.wing_step, 12, 4), "cartpole": (O.cartpole_step, 4, 1)}


@pytest.mark.parametrize("name", ["quad", "wing", "cartpole"])
def test_step_and_adjoint_fp64(hc, name):
    g = load_golden("steps.npz")
    fn, S, A = SYS[name]
    dt = float(g[f"rand_{name}_dt"])
    s = np.ascontiguousarray(g[f"rand_{name}_state"], dtype=np.float64)
    a = np.ascontiguousarray(g[f"rand_{name}_action"], dtype=np.float64)
    cot = np.ascontiguousarray(g[f"rand_{name}_cot"], dtype=np.float64)
    n = s.shape[0]
    pc = P.PHYS[name]()
    out = np.zeros_like(s)
    getattr(hc, f"hc_step_{name}_f64")(_p(s), _p(a), ctypes.c_double(dt), _p(pc), _p(out), n)
    ts, ta = torch.tensor(s, requires_grad=True), torch.tensor(a, requires_grad=True)
    ref_out = fn(ts, ta, dt)
    # fp64 vs fp64; physical constants are fp32-rounded in the harness (as in the kernels) -> 1e-7 level
    assert max_rel_to_scale(out, ref_out) <= 5e-7
    # and against the fp32 reference output itself
    assert max_rel_to_scale(out, g[f"rand_{name}_out"]) <= 5e-6
    gs, ga = np.zeros_like(s), np.zeros_like(a)
    getattr(hc, f"hc_adj_{name}_f64")(_p(s), _p(a), ctypes.c_double(dt), _p(pc), _p(cot), _p(gs), _p(ga), n)
    rgs, rga = torch.autograd.grad((ref_out * torch.tensor(cot)).sum(), (ts, ta))
    assert max_rel_to_scale(gs, rgs) <= 2e-6, name
    assert max_rel_to_scale(ga, rga) <= 2e-6, name
    # reference autograd (fp32) agrees too
    assert max_rel_to_scale(gs, g[f"rand_{name}_gstate"]) <= 3e-5
    assert max_rel_to_scale(ga, g[f"rand_{name}_gaction"]) <= 3e-5


@pytest.mark.parametrize("name", ["quad", "wing", "cartpole"])
def test_step_fp32_kats(hc, name):
    g = load_golden("steps.npz")
    kats = {"quad": [("kat1_state", "kat1_action", "kat1_out", 0.05), ("kat1_state", "kat1b_action", "kat1b_out", 0.1)],
            "wing": [("kat2_state", "kat2_action", "kat2_out", 0.05), ("kat2b_state", "kat2b_action", "kat2b_out", 0.05)],
            "cartpole": [("kat3_state", "kat3_action", "kat3_out", 0.02)]}[name]
    pc = P.PHYS[name]()
    for ks, ka, ko, dt in kats:
        s = np.ascontiguousarray(g[ks], dtype=np.float32)
        a = np.ascontiguousarray(g[ka], dtype=np.float32)
        out = np.zeros_like(s)
        getattr(hc, f"hc_step_{name}_f32")(_p(s), _p(a), ctypes.c_float(dt), _p(pc), _p(out), 1)
        assert max_rel_to_scale(out, g[ko]) <= 2e-6, (name, ko)


def test_features_and_adjoint(hc):
    g = load_golden("steps.npz")
    s = np.ascontiguousarray(g["feat_state"], dtype=np.float64)
    n = s.shape[0]
    f = np.zeros((n, 15))
    hc.hc_features_f64(_p(s), _p(f), n)
    assert max_rel_to_scale(f, g["feat_out"]) <= 2e-6
    cot = np.ascontiguousarray(g["feat_cot"], dtype=np.float64)
    gs = np.zeros_like(s)
    hc.hc_features_adj_f64(_p(s), _p(cot), _p(gs), n)
    ts = torch.tensor(s, requires_grad=True)
    rg = torch.autograd.grad((O.state_preprocessing(ts) * torch.tensor(cot)).sum(), ts)[0]
    assert max_rel_to_scale(gs, rg) <= 1e-9
    assert max_rel_to_scale(gs, g["feat_gstate"]) <= 1e-5


@pytest.mark.parametrize("name,fname", [("quad", "conc_quad_rand.npz"), ("wing", "conc_wing_rand_h20.npz"),
                                        ("cartpole", "conc_cartpole_rand_b128_h5.npz")])
def test_rollout_loss_and_action_grads(hc, name, fname):
    """horizon rollout + reverse sweep with the kernel math (fp64) vs oracle autograd w.r.t. the action sequence,
    and the loss vs the reference's loss."""
    g = load_golden(fname)
    h, dt = int(g["h"]), float(g["dt"])
    fn, S, A = SYS[name]
    cur = np.ascontiguousarray(g["cur"], dtype=np.float64)
    act = np.ascontiguousarray(g["actions"], dtype=np.float64)
    n = cur.shape[0]
    ref = np.ascontiguousarray(g["ref"], dtype=np.float64) if name != "cartpole" else np.zeros(1)
    pc = P.PHYS[name]()
    gact = np.zeros_like(act)
    states = np.zeros((n, h, S))
    f = getattr(hc, f"hc_rollout_{name}_f64")
    f.restype = ctypes.c_double
    loss = f(_p(cur), _p(act), _p(ref), ctypes.c_double(dt), _p(pc), n, h, _p(gact), _p(states))
    assert abs(loss - float(g["loss"])) <= 5e-6 * abs(float(g["loss"]))
    assert max_rel_to_scale(states, g["states"]) <= 1e-5
    ta = torch.tensor(act, requires_grad=True)
    s = torch.tensor(cur)
    sts = []
    for k in range(h):
        s = fn(s, ta[:, k], dt)
        sts.append(s)
    sts = torch.stack(sts, 1)
    tref = O.cartpole_make_reference(torch.tensor(cur), h) if name == "cartpole" else torch.tensor(ref)
    l = O.LOSS_FN[name](sts, tref, ta)
    rg = torch.autograd.grad(l, ta)[0]
    assert abs(loss - float(l)) <= 1e-6 * abs(float(l))
    assert max_rel_to_scale(gact, rg) <= 2e-6
