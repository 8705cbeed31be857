"""The tcgen05 kernels THEMSELVES on the CPU: csrc/hutter_tc_kernels.cu and csrc/adj_dw_tc_kernels.cu (unchanged source,
-DAPG_TC_SIM) compiled with g++ on top of the software model of tests/hostcheck/tc_sim.h - one OS thread per GPU
thread, TMEM / mbarriers / tcgen05.mma (descriptor-decoding model, executed at commit time) in software.  Forward ->
dX chain -> streaming dW GEMM of the split adjoint, end to end against the oracle's loss, actions, states and policy
gradient.  What it proves: roles, hand-off protocol (no deadlock, no lost phase), TMEM lane / column addressing, issue
order and accumulate flags, epilogue math, stash and gradient addressing.  What it cannot: that the hardware agrees
with the model of the instructions (tools/micro/tcgen05_gemm.cu and the GPU parity tests)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import bench as B
from apg_trajectory_tracking_b200 import params as P, synthetic as SY
from oracle import apg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TM, TMP, H = 64, 68, 10


def _build(tmp, name):
    out = tmp / f"lib{name}.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", f"{name}.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("hostcheck_tcsim")
    return _build(tmp, "hostcheck_tcsim"), _build(tmp, "hostcheck_tcsim_dw")


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _unstash(st, rows, n):
    nt = (n + TM - 1) // TM
    return st.reshape(nt, rows, TMP)[:, :, :TM].transpose(0, 2, 1).reshape(nt * TM, rows)[:n]


@pytest.mark.parametrize("n,grid", [(300, 2), pytest.param(128, 1, marks=pytest.mark.slow),
                                    pytest.param(70, 3, marks=pytest.mark.slow)])
def test_tcgen05_forward_and_split_adjoint_run_end_to_end_on_the_model(sim, n, grid):
    fw, dwl = sim
    params = B.default_init("quad", H, seed=n)
    case = SY.quad_case(n, H, 0.1, seed=n)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    f32 = lambda t: np.ascontiguousarray(t.numpy(), np.float32)                 # noqa: E731
    ins, cur, inr, ref = f32(case["in_state"]), f32(case["cur"]), f32(case["in_ref"]), f32(case["ref"])
    pc = P.PHYS["quad"]()
    blob = np.zeros(fw.hc_sim_blob_bytes(), np.uint8)
    fw.hc_sim_pack(_p(flat), _p(blob))
    nt64 = (n + TM - 1) // TM
    nan = lambda rows: np.full(nt64 * rows * TMP, np.nan, np.float32)           # noqa: E731
    x1, h1, h2, h3, act, sts = nan(224), nan(64), nan(64), nan(64), nan(40), nan(H * 12)
    lossp = np.zeros(grid, np.float32)
    states, actions = np.zeros((n, H, 12), np.float32), np.zeros((n, H, 4), np.float32)
    err = ctypes.create_string_buffer(2048)
    nerr = fw.hc_sim_forward(_p(blob), _p(ins), _p(cur), _p(inr), _p(ref), n, ctypes.c_float(0.1), _p(pc), grid, _p(x1),
                             _p(h1), _p(h2), _p(h3), _p(act), _p(sts), _p(lossp), _p(states), _p(actions), err, 2048)
    assert nerr == 0, err.value.decode()
    want_loss, want_grad, want_states, want_actions = O.concurrent_value_and_grad(
        "quad", params, case["in_state"], case["cur"], case["in_ref"], case["ref"], H, 0.1)
    assert np.isfinite(lossp).all()
    assert abs(float(lossp.sum()) - float(want_loss)) <= 2e-5 * abs(float(want_loss))
    assert np.abs(actions - want_actions.detach().numpy()).max() <= 2e-5
    assert np.abs(states - want_states.detach().numpy()).max() <= 1e-4
    # the stash the adjoint kernels read: actions [k*4 + c], states [k*12 + q]
    assert np.abs(_unstash(act, 40, n) - actions.reshape(n, 40)).max() == 0
    assert np.abs(_unstash(sts, H * 12, n) - states.reshape(n, H * 12)).max() == 0

    # ---- split adjoint: dX chain on the same model, then the streaming dW GEMM
    dzo, dz3, dz2, dz1, dzx = nan(40), nan(64), nan(64), nan(64), nan(224)
    nerr = fw.hc_sim_adj_dx(_p(blob), _p(ins), _p(cur), _p(inr), _p(ref), n, ctypes.c_float(0.1), _p(pc), grid, _p(x1),
                            _p(h1), _p(h2), _p(h3), _p(act), _p(sts), _p(dzo), _p(dz3), _p(dz2), _p(dz1), _p(dzx), err,
                            2048)
    assert nerr == 0, err.value.decode()
    npar = dwl.hc_simdw_num_params()
    assert npar == flat.size
    parts = np.full((grid, npar), np.nan, np.float32)
    nerr = dwl.hc_simdw_adj_dw(_p(ins), _p(inr), n, grid, _p(x1), _p(h1), _p(h2), _p(h3), _p(dzo), _p(dz3), _p(dz2),
                               _p(dz1), _p(dzx), _p(parts), err, 2048)
    assert nerr == 0, err.value.decode()
    assert np.isfinite(parts).all(), "a gradient entry was not written (or a NaN operand leaked into a product)"
    grad = parts.astype(np.float64).sum(0)
    o = 0
    for i, (p, g) in enumerate(zip(params, want_grad)):
        got = grad[o:o + p.numel()].reshape(p.shape)
        o += p.numel()
        if g is None:                                                           # ref_in.*: unused by the conv net
            assert np.abs(got).max() == 0
            continue
        scale = max(float(g.abs().max()), 1e-6)
        assert np.abs(got - g.detach().double().numpy()).max() <= 5e-5 * scale, i
    assert fw.hc_sim_mma_count() > 0
