"""The second-generation tcgen05 kernels THEMSELVES on the CPU (csrc/tq_kernels.cu, csrc/tq_dw_kernels.cu, unchanged
source, -DAPG_TC_SIM) on the software model of tests/hostcheck/tc_sim.h: pack -> forward -> dX chain -> streaming
weight-gradient GEMM through the real launchers, against the oracle's loss, actions, states, stash contents and policy
gradient.  What it proves: the four-group / two-slot hand-over protocol (no deadlock, no lost phase), TMEM column use,
the operand-image stash addressing (128B swizzle) on both the writing and the bulk-copy side, the transposed weight
images, accumulate flags, gradient map.  What it cannot: that the hardware agrees with the model of the instructions -
K-major unswizzled / 128B-swizzled operands and A-from-TMEM were measured on B200 (profiles/r2_tcgen05_*.jsonl), and
the GPU parity tests run the same kernels."""
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
H = 10


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("hostcheck_tqsim")
    out = tmp / "libhostcheck_tqsim.so"
    # APG_SIM_DEFINES="-DX -DY": build variants of the kernels (e.g. -DTQ_WIN_PREFETCH) on the model
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++"] +
                          os.environ.get("APG_SIM_DEFINES", "").split() + [
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck_tqsim.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _aligned(nbytes, fill=0xFF):
    raw = np.full(nbytes + 1024, fill, np.uint8)
    off = (-raw.ctypes.data) % 1024
    return raw, raw[off:off + nbytes]


def unstash(buf, tile_bytes, o_rows, R, n):
    """set (first row o_rows, R rows) of every tile -> [n][R] (undoing panels and the 128B swizzle)"""
    nt = (n + 127) // 128
    out = np.zeros((nt * 128, R), np.float32)
    d = np.arange(128)
    for t in range(nt):
        blk = buf[t * tile_bytes + o_rows * 512: t * tile_bytes + (o_rows + R) * 512].view(np.float32)
        for r in range(R):
            off = ((d >> 5) * (R * 128) + r * 128 + ((((d & 31) >> 2) ^ (r & 7)) << 4) + (d & 3) * 4) // 4
            out[t * 128:(t + 1) * 128, r] = blk[off]
    return out[:n]


@pytest.mark.parametrize("n,grid,raw_mode", [(600, 1, False), (200, 1, True),
                                        pytest.param(300, 2, False, marks=pytest.mark.slow),
                                        pytest.param(70, 3, False, marks=pytest.mark.slow)])
def test_tq_forward_dx_chain_and_streaming_dw_gemm_on_the_model(sim, n, grid, raw_mode):
    params = B.default_init("quad", H, seed=n)
    case = SY.quad_case(n, H, 0.1, seed=n)
    if raw_mode:
        # RAW samples (absolute positions): the kernels run QuadDataset.prepare_data (dataset.py:155-204) in their
        # prologue; the expectation is the oracle on the host-prepared tensors of the same samples
        from apg_trajectory_tracking_b200.neural_control import dataset as DS
        pos = torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 6 - 3
        cur_raw, ref_raw = case["cur"].clone(), case["ref"].clone()
        cur_raw[:, :3] = pos
        ref_raw[:, :, :3] += pos[:, None, :]
        ds = DS.QuadDataset.__new__(DS.QuadDataset)
        ds.device = None
        a, b, c, d = ds.prepare_data(cur_raw.clone(), ref_raw.clone())
        case = {"in_state": a, "cur": b, "in_ref": c, "ref": d}
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    f32 = lambda t: np.ascontiguousarray(t.numpy(), np.float32)                 # noqa: E731
    ins, cur, inr, ref = f32(case["in_state"]), f32(case["cur"]), f32(case["in_ref"]), f32(case["ref"])
    k_ins, k_cur, k_inr, k_ref = (None, f32(cur_raw), None, f32(ref_raw)) if raw_mode else (ins, cur, inr, ref)
    pc = P.PHYS["quad"]()
    sz = (ctypes.c_longlong * 7)()
    sim.hc_tq_sizes(n, sz)
    blob_b, tblob_b, fst_b, zst_b, npar, ftile, ztile = [int(x) for x in sz]
    assert npar == flat.size
    keep = []
    bufs = []
    for nb in (blob_b, tblob_b, fst_b, zst_b):
        raw, view = _aligned(nb)                      # 0xFF fill = NaN patterns: anything read before written shows
        keep.append(raw)
        bufs.append(view)
    blob, tblob, fst, zst = bufs
    dyn_grid = 2
    lossp = np.zeros(dyn_grid, np.float32)
    loss_total = np.zeros(1, np.float32)
    parts = np.full((grid, npar), np.nan, np.float32)
    states, actions = np.zeros((n, H, 12), np.float32), np.zeros((n, H, 4), np.float32)
    err = ctypes.create_string_buffer(4096)
    nerr = sim.hc_tq_step(_p(flat), _p(k_ins), _p(k_cur), _p(k_inr), _p(k_ref), n, ctypes.c_float(0.1), _p(pc), grid, _p(blob),
                          _p(tblob), _p(fst), _p(zst), _p(lossp), _p(parts), _p(states), _p(actions), 3, dyn_grid, _p(loss_total), err, 4096)
    assert nerr == 0, err.value.decode()
    want_loss, want_grad, want_states, want_actions = O.concurrent_value_and_grad(
        "quad", params, case["in_state"], case["cur"], case["in_ref"], case["ref"], H, 0.1)
    assert np.isfinite(lossp).all()
    assert abs(float(lossp.sum()) - float(want_loss)) <= 2e-5 * abs(float(want_loss))
    assert abs(float(loss_total[0]) - float(lossp.astype(np.float64).sum())) <= 1e-6 * abs(float(want_loss))   # last-block sum
    assert np.abs(actions - want_actions.detach().numpy()).max() <= 2e-5
    assert np.abs(states - want_states.detach().numpy()).max() <= 1e-4
    # stash set the dynamics kernel reads: actions [k*4 + c] (first row 592)
    assert np.abs(unstash(fst, ftile, 592, 40, n) - actions.reshape(n, 40)).max() == 0
    assert np.abs(unstash(fst, ftile, 0, 16, n)[:, :15] - ins).max() <= (1e-6 if raw_mode else 0)
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
    assert sim.hc_tq_mma_count() > 0
