"""Host-checkable parts of the optional tcgen05 forward (csrc/tc_layout.cuh): the pack-kernel body (torch-flat
parameters -> (hi, lo) K-major weight images: padding, Toeplitz block of the conv, fc1 column permutation), the op
list and the stash addressing, emulated on the CPU from the PACKED images and compared with the oracle's policy
forward in the layout the (GPU-verified) adjoint kernel reads.  The tcgen05 instructions themselves can only be
checked on the GPU (tests/test_zz_new_paths_gpu.py::test_tc_forward_*)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import bench as B
from apg_trajectory_tracking_b200 import synthetic as SY
from oracle import apg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TM, TMP = 64, 68


@pytest.fixture(scope="module")
def ht(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_tc") / "libhostcheck_tc.so"
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck_tc.cpp")
    inc = os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I", inc,
                           src, "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _unstash(st, rows, n):
    """[ntiles64][rows][TMP] -> (n, rows)"""
    nt = (n + TM - 1) // TM
    a = st.reshape(nt, rows, TMP)[:, :, :TM].transpose(0, 2, 1).reshape(nt * TM, rows)
    return a[:n]


@pytest.mark.parametrize("n", [1, 64, 65, 130, 300])
def test_packed_images_and_op_list_reproduce_the_policy_in_stash_layout(ht, n):
    h = 10
    params = B.default_init("quad", h, seed=n)
    case = SY.quad_case(n, h, 0.1, seed=n)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    blob = np.zeros(ht.hc_tc_blob_bytes(), np.uint8)
    ht.hc_tc_pack(_p(flat), _p(blob))
    nt64 = (n + TM - 1) // TM
    nan = lambda rows: np.full(nt64 * rows * TMP, np.nan, np.float32)          # noqa: E731
    x1, h1, h2, h3, act = nan(224), nan(64), nan(64), nan(64), nan(40)
    ins = np.ascontiguousarray(case["in_state"].numpy(), np.float32)
    inr = np.ascontiguousarray(case["in_ref"].numpy(), np.float32)
    ht.hc_tc_emulate(_p(blob), _p(ins), _p(inr), n, _p(x1), _p(h1), _p(h2), _p(h3), _p(act))
    # oracle activations (fp64), conv outputs reordered channel-major (c*8 + t) -> position-major (t*20 + c)
    ps = [p.double() for p in params]
    ws, bs, wc, bc, wr, br, w1, b1, w2, b2, w3, b3, wo, bo = ps
    s = torch.tanh(case["in_state"].double() @ ws.t() + bs)
    conv = O._conv_encoder(case["in_ref"].double(), wc, bc)                     # (n, 160) channel-major
    conv_pm = conv.reshape(n, 20, 8).transpose(1, 2).reshape(n, 160)
    a1 = torch.tanh(torch.cat((s, conv), 1) @ w1.t() + b1)
    a2 = torch.tanh(a1 @ w2.t() + b2)
    a3 = torch.tanh(a2 @ w3.t() + b3)
    out = torch.sigmoid(a3 @ wo.t() + bo)
    for got, rows, want in ((x1, 224, torch.cat((s, conv_pm), 1)), (h1, 64, a1), (h2, 64, a2), (h3, 64, a3),
                            (act, 40, out)):
        g = _unstash(got, rows, n)
        assert np.isfinite(g).all()
        assert np.abs(g - want.numpy()).max() <= 3e-6, rows
    # every column of every existing 64-drone tile is written (the adjoint multiplies dead columns by zero: they
    # must be finite), the TMP - 64 padding floats are never touched
    full = x1.reshape(nt64, 224, TMP)
    assert np.isfinite(full[:, :, :TM]).all() and np.isnan(full[:, :, TM:]).all()


def test_packed_images_are_hi_lo_splits_with_zero_padding(ht):
    params = B.default_init("quad", 10, seed=5)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    blob = np.full(ht.hc_tc_blob_bytes(), 0xAB, np.uint8)
    ht.hc_tc_pack(_p(flat), _p(blob))
    words = blob.view(np.uint32)
    assert not np.any(words == 0xABABABAB)                   # every word of the blob is written
    img = blob[:228352].view(np.float32)
    assert np.isfinite(img).all()
    # hi images keep 19 significant bits: low 13 mantissa bits clear in the first (hi) half of the first image
    ws_hi = blob[:64 * 16 * 4].view(np.uint32)
    assert np.all((ws_hi & 0x1FFF) == 0)
    # hi + lo reproduces the weight exactly: states_in.weight[3][7] sits at kmajor_off(3, 7, 16)
    off = (3 >> 3) * ((16 >> 2) * 128) + (7 >> 2) * 128 + (3 & 7) * 16 + (7 & 3) * 4
    hi = blob[off:off + 4].view(np.float32)[0]
    lo = blob[64 * 16 * 4 + off:64 * 16 * 4 + off + 4].view(np.float32)[0]
    assert np.float32(hi) + np.float32(lo) == params[0][3, 7].numpy()


# ---------------------------------------------------------------------------------------------------------------
# streaming weight-gradient GEMM of the optional split adjoint (csrc/adj_dw_layout.cuh)
# ---------------------------------------------------------------------------------------------------------------
def _stash(a, rows):
    """(n, rows) -> [ntiles64][rows][TMP] float32 with finite garbage in dead columns and NaN in the padding"""
    n = a.shape[0]
    nt = (n + TM - 1) // TM
    st = np.full((nt, rows, TMP), np.nan, np.float32)
    full = np.zeros((nt * TM, rows), np.float32)
    full[:n] = a
    st[:, :, :TM] = full.reshape(nt, TM, rows).transpose(0, 2, 1)
    return st


@pytest.mark.parametrize("n", [64, 65, 200])
def test_streaming_dw_gemm_reproduces_the_weight_gradient(ht, n):
    h = 10
    params = [p.double() for p in B.default_init("quad", h, seed=100 + n)]
    ws, bs, wc, bc, wr, br, w1, b1, w2, b2, w3, b3, wo, bo = params
    case = SY.quad_case(n, h, 0.1, seed=n)
    ins, inr = case["in_state"].double(), case["in_ref"].double()
    s = torch.tanh(ins @ ws.t() + bs)
    conv = O._conv_encoder(inr, wc, bc)                                     # channel-major (c*8 + t)
    x = torch.cat((s, conv), 1)
    a1 = torch.tanh(x @ w1.t() + b1)
    a2 = torch.tanh(a1 @ w2.t() + b2)
    a3 = torch.tanh(a2 @ w3.t() + b3)
    g = torch.Generator().manual_seed(n)
    dzo = torch.randn(n, 40, generator=g, dtype=torch.float64)
    dz3 = (dzo @ wo) * (1 - a3 ** 2)
    dz2 = (dz3 @ w3) * (1 - a2 ** 2)
    dz1 = (dz2 @ w2) * (1 - a1 ** 2)
    dx = dz1 @ w1
    ds = dx[:, :64] * (1 - s ** 2)
    dconv = dx[:, 64:] * (conv > 0)
    pm = lambda t: t.reshape(n, 20, 8).transpose(1, 2).reshape(n, 160)      # noqa: E731  channel- -> position-major
    f32 = lambda t: t.float().numpy()                                        # noqa: E731
    st_x1, st_h1, st_h2, st_h3 = _stash(f32(torch.cat((s, pm(conv)), 1)), 224), _stash(f32(a1), 64), \
        _stash(f32(a2), 64), _stash(f32(a3), 64)
    z_o, z_3, z_2, z_1 = _stash(f32(dzo), 40), _stash(f32(dz3), 64), _stash(f32(dz2), 64), _stash(f32(dz1), 64)
    z_x = _stash(f32(torch.cat((ds, pm(dconv)), 1)), 224)
    ins32 = np.ascontiguousarray(case["in_state"].numpy(), np.float32)
    inr32 = np.ascontiguousarray(case["in_ref"].numpy(), np.float32)
    P = np.zeros(ht.hc_dw_num_params(), np.float32)
    ht.hc_dw_emulate(_p(st_x1), _p(st_h1), _p(st_h2), _p(st_h3), _p(ins32), _p(inr32), _p(z_o), _p(z_3), _p(z_2),
                     _p(z_1), _p(z_x), n, _p(P))
    assert np.isfinite(P).all(), "an entry was written twice / not at all, or padding leaked into a product"
    gwc = torch.einsum("nct,ntjd->cdj", dconv.reshape(n, 20, 8),
                       torch.stack([inr[:, t:t + 3, :] for t in range(8)], 1))          # [c][ci][j]
    want = [ds.t() @ ins, ds.sum(0), gwc, dconv.reshape(n, 20, 8).sum((0, 2)), torch.zeros(64, 90), torch.zeros(64),
            dz1.t() @ x, dz1.sum(0), dz2.t() @ a1, dz2.sum(0), dz3.t() @ a2, dz3.sum(0), dzo.t() @ a3, dzo.sum(0)]
    o = 0
    for i, w in enumerate(want):
        got = P[o:o + w.numel()].reshape(w.shape)
        o += w.numel()
        scale = max(float(w.abs().max()), 1e-6)
        assert np.abs(got - w.numpy()).max() <= 2e-5 * scale + (0 if i not in (4, 5) else 0), i
    assert o == P.size


# ---------------------------------------------------------------------------------------------------------------
# descriptor-level emulation: the MMAs are executed by a software model that decodes the shared-memory / instruction
# descriptors the kernels build (canonical unswizzled layouts of cute/atom/mma_traits_sm100.hpp, TF32 truncation)
# ---------------------------------------------------------------------------------------------------------------
def _net_and_case(n, seed):
    params = B.default_init("quad", 10, seed=seed)
    case = SY.quad_case(n, 10, 0.1, seed=seed)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), dtype=np.float32)
    return params, case, flat


@pytest.mark.parametrize("n", [128, 77])
def test_forward_issuer_descriptors_reproduce_the_policy(ht, n):
    params, case, flat = _net_and_case(n, 11)
    blob = np.zeros(ht.hc_tc_blob_bytes(), np.uint8)
    ht.hc_tc_pack(_p(flat), _p(blob))
    act = np.zeros((128, 40), np.float32)
    ins = np.ascontiguousarray(case["in_state"].numpy(), np.float32)
    inr = np.ascontiguousarray(case["in_ref"].numpy(), np.float32)
    assert ht.hc_tc_forward_desc(_p(blob), _p(ins), _p(inr), n, _p(act)) == 1, "a descriptor failed to decode"
    want = torch.sigmoid(O.hutter_forward([p.double() for p in params], case["in_state"].double(),
                                          case["in_ref"].double())).numpy()
    assert np.abs(act[:n] - want).max() <= 3e-6            # 3xTF32 with truncating inputs: fp32-level accuracy


def test_dx_issuer_mn_major_descriptors_reproduce_the_dx_chain(ht):
    n = 128
    params, case, flat = _net_and_case(n, 12)
    blob = np.zeros(ht.hc_tc_blob_bytes(), np.uint8)
    ht.hc_tc_pack(_p(flat), _p(blob))
    ps = [p.double() for p in params]
    ws, bs, wc, bc, wr, br, w1, b1, w2, b2, w3, b3, wo, bo = ps
    s = torch.tanh(case["in_state"].double() @ ws.t() + bs)
    conv = O._conv_encoder(case["in_ref"].double(), wc, bc)
    x = torch.cat((s, conv), 1)
    a1 = torch.tanh(x @ w1.t() + b1)
    a2 = torch.tanh(a1 @ w2.t() + b2)
    a3 = torch.tanh(a2 @ w3.t() + b3)
    g = torch.Generator().manual_seed(3)
    dlog = torch.randn(n, 40, generator=g, dtype=torch.float64)
    dz3 = (dlog @ wo) * (1 - a3 ** 2)
    dz2 = (dz3 @ w3) * (1 - a2 ** 2)
    dz1 = (dz2 @ w2) * (1 - a1 ** 2)
    dx = dz1 @ w1
    pm = lambda t: t.reshape(n, 20, 8).transpose(1, 2).reshape(n, 160)      # noqa: E731
    want_x = torch.cat((dx[:, :64] * (1 - s ** 2), pm(dx[:, 64:] * (conv > 0))), 1)
    f32 = lambda t: np.ascontiguousarray(t.float().numpy())                  # noqa: E731
    o3, o2, o1, ox = (np.zeros((n, 64), np.float32) for _ in range(3)), None, None, np.zeros((n, 224), np.float32)
    o3, o2, o1 = list(o3)
    ok = ht.hc_tc_dx_desc(_p(blob), _p(f32(dlog)), _p(f32(a3)), _p(f32(a2)), _p(f32(a1)),
                          _p(f32(torch.cat((s, pm(conv)), 1))), _p(o3), _p(o2), _p(o1), _p(ox))
    assert ok == 1, "a descriptor failed to decode"
    for got, want in ((o3, dz3), (o2, dz2), (o1, dz1), (ox, want_x)):
        assert np.abs(got - want.numpy()).max() <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize("N", [48, 64])
def test_dw_issuer_descriptors_reproduce_one_gemm(ht, N):
    rng = np.random.default_rng(N)
    a = rng.standard_normal((128, 64)).astype(np.float32)
    b = np.zeros((64, 64), np.float32)
    b[:N] = rng.standard_normal((N, 64)).astype(np.float32)
    d = np.zeros((128, N), np.float32)
    assert ht.hc_dw_op_desc(_p(a), _p(b), N, _p(d)) == 1, "a descriptor failed to decode"
    want = a.astype(np.float64) @ b[:N].astype(np.float64).T
    assert np.abs(d - want).max() <= 2e-5 * np.abs(want).max()


@pytest.mark.parametrize("ncta", [1, 3, 4, 7, 148])
def test_sliced_partial_reduction_covers_every_cta_once(ht, ncta):
    rng = np.random.default_rng(ncta)
    n = 1891
    partials = rng.standard_normal((ncta, n)).astype(np.float32)
    grad = np.zeros(n, np.float32)
    ht.hc_reduce4(_p(partials), ncta, n, ctypes.c_float(0.5), _p(grad))
    want = 0.5 * partials.astype(np.float64).sum(0)
    assert np.abs(grad - want).max() <= 1e-5 * np.abs(want).max()
