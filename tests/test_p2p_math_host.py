"""Gradient exchange over peer memory (csrc/p2p_math.cuh / p2p_kernels.cu): layout, reduction order, rank-order sum,
fused SGD update and the two-set / epoch protocol, with all ranks emulated in one process on the CPU.  The
memory-ordering half of the protocol (fences, release / acquire flags over NVLink) needs hardware:
tests/multi_gpu_check.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hp(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostcheck_p2p") / "libhostcheck_p2p.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-std=c++17", "-ffp-contract=off", "-I",
                           os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck_p2p.cpp"), "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    lib.hc_p2p_total_floats.restype = ctypes.c_size_t
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _reduce_slice(partials):
    """four interleaved running sums over the CTAs of one slice, then (s0+s1)+(s2+s3)"""
    ncta, n = partials.shape
    s = np.zeros((4, n), np.float32)
    c = 0
    while c + 3 < ncta:
        for k in range(4):
            s[k] += partials[c + k]
        c += 4
    while c < ncta:
        s[0] += partials[c]
        c += 1
    return (s[0] + s[1]) + (s[2] + s[3])


def _reduce_like_kernel(partials, scale, perm=None):
    """apg_reduce_scatter_p2p_kernel's association (= apg_reduce4_kernel's): four slices of CTAs, each summed with four
    interleaved running sums, then (t0+t1)+(t2+t3)"""
    ncta, n = partials.shape
    per = (ncta + 3) // 4
    t = [_reduce_slice(partials[min(k * per, ncta):min((k + 1) * per, ncta)]) if min(k * per, ncta) < ncta
         else np.zeros(n, np.float32) for k in range(4)]
    out = np.float32(scale) * ((t[0] + t[1]) + (t[2] + t[3]))
    return out if perm is None else out[perm]


@pytest.mark.parametrize("world,n,ncta", [(1, 37, 5), (2, 1000, 148), (8, 333, 37)])
def test_exchange_equals_rank_ordered_sum_and_sgd(hp, world, n, ncta):
    rng = np.random.default_rng(world * 100 + n)
    total = hp.hc_p2p_total_floats(world, n)
    assert total == 2 * (world * n + world)
    mem = np.zeros(world * total, np.float32)
    params = [rng.standard_normal(n).astype(np.float32) for _ in range(world)]
    params = [params[0].copy() for _ in range(world)]                      # replicas start identical
    bufs = [np.zeros(n, np.float32) for _ in range(world)]
    ref_p, ref_b = params[0].copy(), np.zeros(n, np.float32)
    lr, mom = np.float32(1e-3), np.float32(0.9)
    for epoch in range(1, 6):                                              # both sets are reused several times
        parts = [rng.standard_normal((ncta, n)).astype(np.float32) for _ in range(world)]
        order = rng.permutation(world)                                     # arrival order must not matter
        for r in order:
            hp.hc_p2p_reduce_scatter(_p(mem), world, n, int(r), epoch, _p(parts[r]), ncta, ctypes.c_float(0.5), 0, 0, 0)
            if r != order[-1]:                                             # someone is still missing: keep waiting
                g = np.zeros(n, np.float32)
                assert hp.hc_p2p_gather_sgd(_p(mem), world, n, int(order[0]), epoch, _p(g), None, None,
                                            ctypes.c_float(0), ctypes.c_float(0)) == 0
        want = np.zeros(n, np.float32)
        for q in range(world):
            want = want + _reduce_like_kernel(parts[q], 0.5)               # rank order, float32
        ref_b = mom * ref_b + want
        ref_p = ref_p - lr * ref_b
        for r in range(world):
            g = np.zeros(n, np.float32)
            assert hp.hc_p2p_gather_sgd(_p(mem), world, n, r, epoch, _p(g), _p(params[r]), _p(bufs[r]),
                                        ctypes.c_float(lr), ctypes.c_float(mom)) == 1
            assert np.array_equal(g, want)                                 # bitwise the same on every rank
            assert np.array_equal(params[r], ref_p) and np.array_equal(bufs[r], ref_b)
        # optimizer semantics = torch.optim.SGD(momentum=0.9)
    tp = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.SGD([tp], lr=0.1, momentum=0.9)
    for gval in (1.0, 2.0):
        tp.grad = torch.full((3,), gval)
        opt.step()
    b = 0.9 * 1.0 + 2.0
    assert torch.allclose(tp.detach(), torch.full((3,), -(0.1 * 1.0 + 0.1 * b)))


def test_position_major_fc1_block_is_read_in_torch_order(hp):
    """the hutter conv nets' fc1 gradient block is kept position-major by the kernels (apg_reduce_kernel)"""
    npos, k1, off = 8, 64 + 20 * 8, 11
    n = off + 64 * k1 + 5
    rng = np.random.default_rng(0)
    parts = rng.standard_normal((3, n)).astype(np.float32)
    perm = np.arange(n)
    for j in range(64):
        for k in range(64, k1):
            c, t = divmod(k - 64, npos)
            perm[off + j * k1 + k] = off + j * k1 + 64 + t * 20 + c
    total = hp.hc_p2p_total_floats(1, n)
    mem = np.zeros(total, np.float32)
    hp.hc_p2p_reduce_scatter(_p(mem), 1, n, 0, 1, _p(parts), 3, ctypes.c_float(1.0), off, k1, npos)
    g = np.zeros(n, np.float32)
    assert hp.hc_p2p_gather_sgd(_p(mem), 1, n, 0, 1, _p(g), None, None, ctypes.c_float(0), ctypes.c_float(0)) == 1
    assert np.array_equal(g, _reduce_like_kernel(parts, 1.0, perm))


def test_comm_layout_queries_of_the_library():
    from apg_trajectory_tracking_b200 import _capi
    lib = _capi.lib()
    assert lib.apg_grad_comm_bytes(8, 32728) == 4 * 2 * (8 * 32728 + 8)
    so, fo = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.apg_grad_comm_offsets(8, 32728, 1, ctypes.byref(so), ctypes.byref(fo)) == 0
    assert so.value == 4 * (8 * 32728 + 8) and fo.value == so.value + 4 * 8 * 32728
    assert lib.apg_grad_comm_offsets(8, 32728, 2, ctypes.byref(so), ctypes.byref(fo)) != 0
    assert lib.apg_grad_comm_bytes(0, 10) == 0


def test_peer_grad_exchange_tables_and_descriptors(hp, monkeypatch):
    """dist.PeerGradExchange on the CPU with the symmetric-memory allocator and the CUDA context calls stubbed:
    pointer tables, set alternation, descriptor fields; the harness then plays both kernels of a 1-rank step on the
    very memory the tables point to."""
    import contextlib
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm
    from apg_trajectory_tracking_b200 import _capi, dist as D
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29617", rank=0, world_size=1)
    try:
        class _Handle:
            def __init__(self, t):
                self.buffer_ptrs = [t.data_ptr()]
        monkeypatch.setattr(symm, "empty", lambda numel, dtype=None, device=None: torch.zeros(numel, dtype=dtype))
        monkeypatch.setattr(symm, "rendezvous", lambda t, group: _Handle(t))
        monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
        monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
        n = 50
        ex = D.PeerGradExchange(n, "cpu")
        assert ex.world == 1 and ex.rank == 0 and ex.buf.numel() == 2 * (n + 1)
        base = ex.buf.data_ptr()
        assert ex.slot_tab.tolist() == [[base], [base + 4 * (n + 1)]]
        assert ex.flag_tab.tolist() == [[base + 4 * n], [base + 4 * (n + 1) + 4 * n]]
        rng = np.random.default_rng(5)
        mem = ex.buf.numpy()
        for step in (1, 2, 3):
            comm, local = ex.next_step()
            assert (comm.rank, comm.world, comm.epoch) == (0, 1, step)
            assert comm.slot_ptrs == ex.slot_tab[step & 1].data_ptr() and local.value == ex.slot_tab[step & 1, 0].item()
            parts = rng.standard_normal((7, n)).astype(np.float32)
            hp.hc_p2p_reduce_scatter(_p(mem), 1, n, 0, step, _p(parts), 7, ctypes.c_float(1.0), 0, 0, 0)
            want = _reduce_like_kernel(parts, 1.0)
            off = (local.value - base) // 4
            assert np.array_equal(mem[off:off + n], want)                     # landed in the set the table names
            assert mem[off + n:off + n + 1].view(np.uint32)[0] == step
    finally:
        dist.destroy_process_group()


def test_p2p_kernels_themselves_on_the_thread_model(tmp_path_factory):
    """csrc/p2p_kernels.cu (unchanged source, -DAPG_SIM): every rank a launch of 128-thread CTAs on the CPU thread
    model - ticket / last-CTA flag raise, flag wait, rank-ordered sum and fused SGD as written in the kernels"""
    out = tmp_path_factory.mktemp("hostcheck_p2psim") / "libhostcheck_p2psim.so"
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "apg_trajectory_tracking_b200", "csrc"),
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck_p2psim.cpp"), "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    world, n, ncta = 3, 300, 5
    rng = np.random.default_rng(9)
    total = 2 * (world * n + world)
    mem = np.zeros(world * total, np.float32)
    tickets = np.zeros(world, np.uint32)
    params = np.tile(rng.standard_normal(n).astype(np.float32), (world, 1))
    bufs = np.zeros((world, n), np.float32)
    ref_p, ref_b = params[0].copy(), np.zeros(n, np.float32)
    lr, mom = np.float32(1e-2), np.float32(0.9)
    err = ctypes.create_string_buffer(1024)
    for epoch in (1, 2, 3):
        parts = rng.standard_normal((world, ncta, n)).astype(np.float32)
        order = rng.permutation(world).astype(np.int32)
        grads = np.zeros((world, n), np.float32)
        ne = lib.hc_p2psim_step(_p(mem), world, n, epoch, _p(parts), ncta, ctypes.c_float(1.0), _p(order), _p(grads),
                                _p(params), _p(bufs), ctypes.c_float(lr), ctypes.c_float(mom), _p(tickets), err, 1024)
        assert ne == 0, err.value.decode()
        want = np.zeros(n, np.float32)
        for q in range(world):
            want = want + _reduce_like_kernel(parts[q], 1.0)
        ref_b = mom * ref_b + want
        ref_p = ref_p - lr * ref_b
        for r in range(world):
            assert np.array_equal(grads[r], want)
            assert np.array_equal(params[r], ref_p) and np.array_equal(bufs[r], ref_b)
        assert (tickets == 0).all()                                       # reset by the last CTA of each launch
