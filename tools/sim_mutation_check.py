"""How sharp are the CPU models?  Applies known-bad one-line mutations to a scratch copy of the kernel sources, rebuilds
the model harness and reports whether the model run notices (wrong numbers, a recorded violation, or a timeout of the
kernels' own bounded waits).  Usage: python tools/sim_mutation_check.py  (a few minutes; prints one line per mutant);
python tools/sim_mutation_check.py --smem  (a launcher requesting too little shared memory, on the model library)."""
import ctypes
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from apg_trajectory_tracking_b200 import params as P, synthetic as SY  # noqa: E402
from oracle import apg_oracle as O  # noqa: E402

MUTANTS = [
    ("baseline (no mutation)", None, None, None),
    ("forward chain: epilogue does not wait for fc2 before reading its accumulator", "tq_kernels.cu",
     "      wait_d();                                               // fc2\n", "\n"),
    ("forward chain: epilogue warps address TMEM lanes of the wrong quarter", "tq_kernels.cu",
     "    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)((warp & 3) * 32) << 16);\n    const uint32_t d_main",
     "    const uint32_t slot = tmem + s * SLOT_COLS + ((uint32_t)(((warp + 1) & 3) * 32) << 16);\n    const uint32_t d_main"),
    ("dW GEMM: first MMA of a pass never clears the accumulator", "tq_dw_kernels.cu",
     "tcp::mma_ts(d, a_lo + ks * 8, dbh, idesc, (ks > 0 || !clear) ? 1u : 0u);", "tcp::mma_ts(d, a_lo + ks * 8, dbh, idesc, 1u);"),
    ("dW GEMM: the producer of pass 1 does not wait for the MMAs of pass 0 (stages of different size overlap)",
     "tq_dw_kernels.cu",
     "      if (pass > 0 && my_tiles > 0) dwq_wait(smem_u32(&s_bars.done[pass - 1]), 0, abort_flag);\n      for (int j = 0; j < my_tiles; ++j) {\n        const int tile",
     "      for (int j = 0; j < my_tiles; ++j) {\n        const int tile"),
    ("dW GEMM: the B feeder does not wait for its slot to be free", "tq_dw_kernels.cu",
     "          if (u >= NS) dwq_wait(smem_u32(&s_bars.sfree[sl]), (uint32_t)(u / NS - 1) & 1u, abort_flag);\n          TQP(1);\n#ifndef DWQ_PROBE_NO_FEED\n#pragma unroll\n          for (int kk = 0; kk < 16; ++kk) if (lane + kk * 32 < nb) b_lo",
     "          TQP(1);\n#ifndef DWQ_PROBE_NO_FEED\n#pragma unroll\n          for (int kk = 0; kk < 16; ++kk) if (lane + kk * 32 < nb) b_lo"),
    ("dW GEMM: the lo MMA reads the wrong k-step of the B image", "tq_dw_kernels.cu",
     "tcp::mma_ts(d, a_hi + ks * 8, dbl, idesc, 1u);", "tcp::mma_ts(d, a_hi + ks * 8, DESC_HI | (bl - 2 * ks), idesc, 1u);"),
]


def run_step(src_root, n=600, grid=1):
    """pack -> forward chain -> dynamics -> dX chain -> dW GEMM of the (mutated) sources on the CPU model, against the
    oracle: loss, actions, gradient (hostcheck_tqsim.cpp, the harness of tests/test_tq_kernel_sim_host.py)"""
    tmp = tempfile.mkdtemp()
    lib = os.path.join(tmp, "libm.so")
    r = subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                        "-I", os.path.join(src_root, "apg_trajectory_tracking_b200", "csrc"),
                        os.path.join(src_root, "tests", "hostcheck", "hostcheck_tqsim.cpp"), "-o", lib],
                       capture_output=True, text=True)
    if r.returncode != 0:
        return "does not compile"
    sim = ctypes.CDLL(lib)
    h = 10
    params = B.default_init("quad", h, seed=n)
    case = SY.quad_case(n, h, 0.1, seed=n)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), np.float32)
    f32 = lambda t: np.ascontiguousarray(t.numpy(), np.float32)                  # noqa: E731
    ins, cur, inr, ref = f32(case["in_state"]), f32(case["cur"]), f32(case["in_ref"]), f32(case["ref"])
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                              # noqa: E731
    sz = (ctypes.c_longlong * 7)()
    sim.hc_tq_sizes(n, sz)
    blob_b, tblob_b, fst_b, zst_b, npar, _, _ = [int(x) for x in sz]
    keep, bufs = [], []
    for nb in (blob_b, tblob_b, fst_b, zst_b):
        raw = np.full(nb + 1024, 0xFF, np.uint8)
        off = (-raw.ctypes.data) % 1024
        keep.append(raw)
        bufs.append(raw[off:off + nb])
    blob, tblob, fst, zst = bufs
    lossp, lt = np.zeros(2, np.float32), np.zeros(1, np.float32)
    parts = np.full((grid, npar), np.nan, np.float32)
    st, ac = np.zeros((n, h, 12), np.float32), np.zeros((n, h, 4), np.float32)
    err = ctypes.create_string_buffer(2048)
    os.environ["APG_SIM_FAST_TIMEOUT"] = "1"
    ne = sim.hc_tq_step(p(flat), p(ins), p(cur), p(inr), p(ref), n, ctypes.c_float(0.1), p(P.PHYS["quad"]()), grid,
                        p(blob), p(tblob), p(fst), p(zst), p(lossp), p(parts), p(st), p(ac), 3, 2, p(lt), err, 2048)
    want, wg, _, want_act = O.concurrent_value_and_grad("quad", params, case["in_state"], case["cur"], case["in_ref"],
                                                        case["ref"], h, 0.1)
    if ne:
        return "model violation: " + err.value.decode()[:120]
    if not np.isfinite(lossp).all() or not np.isfinite(parts).all():
        return "NaN in loss / gradient (poisoned, uninitialised or in-flight data was read)"
    rel = abs(float(lossp.sum()) - float(want)) / abs(float(want))
    da = float(np.abs(ac - want_act.detach().numpy()).max())
    grad, o, worst = parts.astype(np.float64).sum(0), 0, 0.0
    for prm, g in zip(params, wg):
        got = grad[o:o + prm.numel()].reshape(prm.shape)
        o += prm.numel()
        if g is not None:
            worst = max(worst, float(np.abs(got - g.detach().double().numpy()).max()) / max(float(g.abs().max()), 1e-6))
    ok = rel <= 2e-5 and da <= 2e-5 and worst <= 5e-5
    return "results match the oracle" if ok else f"wrong results (loss rel {rel:.2e}, actions {da:.2e}, gradient {worst:.2e})"


def main():
    for name, fname, old, new in MUTANTS:
        scratch = tempfile.mkdtemp()
        for d in ("apg_trajectory_tracking_b200/csrc", "tests/hostcheck"):
            shutil.copytree(os.path.join(ROOT, d), os.path.join(scratch, d), ignore=shutil.ignore_patterns("*.so", "*.o"))
        if fname and old is None:
            continue
        if fname:
            path = os.path.join(scratch, "apg_trajectory_tracking_b200", "csrc", fname)
            s = open(path).read()
            assert s.count(old) >= 1, (name, "pattern not found")
            open(path, "w").write(s.replace(old, new, 1))
        print(f"{name}: {run_step(scratch)}", flush=True)
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__" and "--smem" not in sys.argv:
    main()


# ---- a launcher that requests too little dynamic shared memory: caught by the canary of the model library ------------
def smem_mutant():
    from apg_trajectory_tracking_b200 import _capi, rollout as R
    scratch = tempfile.mkdtemp()
    for d in ("apg_trajectory_tracking_b200/csrc", "tests/hostcheck", "include"):
        shutil.copytree(os.path.join(ROOT, d), os.path.join(scratch, d), ignore=shutil.ignore_patterns("*.so", "*.o"))
    path = os.path.join(scratch, "apg_trajectory_tracking_b200", "csrc", "eval_kernels.cu")
    s = open(path).read()
    old = "return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + y.rows_total * TMP) + 16;"
    assert old in s
    open(path, "w").write(s.replace(old, "return sizeof(float) * (size_t)(y.f_total + pad4(TM * y.F0) + (y.rows_total - 8) * TMP) + 16;"))
    objs = []
    for name in ("host", "te", "tq"):
        obj = os.path.join(scratch, name + ".o")
        subprocess.check_call(["g++", "-O1", "-c", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                               "-I", os.path.join(scratch, "apg_trajectory_tracking_b200", "csrc"),
                               os.path.join(scratch, "tests", "hostcheck", "capisim", name + ".cpp"), "-o", obj])
        objs.append(obj)
    libp = os.path.join(scratch, "libsim.so")
    subprocess.check_call(["g++", "-shared", "-pthread"] + objs + ["-o", libp])
    lib = ctypes.CDLL(libp)
    lib.apg_workspace_bytes.restype = ctypes.c_size_t
    n, steps = 8, 3
    cfg = R.RolloutSpec.cartpole_concurrent(10, 0.05).config(n)
    ws = torch.zeros(lib.apg_workspace_bytes(ctypes.byref(cfg)) + 512, dtype=torch.uint8)
    wsp = ctypes.c_void_p(ws.data_ptr() + (-ws.data_ptr()) % 256)
    params = torch.cat([p.reshape(-1) for p in B.default_init("cartpole", 10, seed=0)]).contiguous()
    init = torch.zeros(n, 4)
    states, nst = torch.zeros(n, steps, 4), torch.zeros(n, dtype=torch.int32)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())                                 # noqa: E731
    rc = lib.apg_eval_cartpole(ctypes.byref(cfg), vp(params), vp(init), steps, ctypes.c_float(0.21), 1, wsp,
                               vp(states), None, vp(nst), None, None, None, None)
    buf = ctypes.create_string_buffer(1024)
    ne = lib.apg_sim_take_errors(buf, 1024)
    print("launcher of eval_cartpole_kernel requests 8 rows of shared memory too few: rc", rc, "->",
          (buf.value.decode()[:110] if ne else "NOT noticed"), flush=True)
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__" and "--smem" in sys.argv:
    smem_mutant()
