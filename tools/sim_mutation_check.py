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
    ("tcgen05 forward: first MMA of an op never clears the accumulator", "hutter_tc_kernels.cu",
     "mma_ts(d, alo + ks * 8, bh, idesc, (ks > 0 || !op.clear) ? 1u : 0u);", "mma_ts(d, alo + ks * 8, bh, idesc, 1u);"),
    ("tcgen05 forward: epilogue does not wait for fc2 before reading its accumulator", "hutter_tc_kernels.cu",
     "      wait_d();                                               // fc2\n", "\n"),
    ("tcgen05 forward: A operand published before the TMEM stores are fenced (arrive first)", "hutter_tc_kernels.cu",
     "  tcp::wait_st();\n  tcp::fence_before_thread_sync();\n  tc_mbar_arrive(bar);", "  tc_mbar_arrive(bar);\n  tcp::wait_st();"),
    ("tcgen05 forward: hi*lo term reads the wrong k-step of the weight image", "hutter_tc_kernels.cu",
     "mma_ts(d, ahi + ks * 8, bl, idesc, 1u);", "mma_ts(d, ahi + ks * 8, kmajor_desc(wlo, 0, op.K), idesc, 1u);"),
    ("tcgen05 forward: epilogue warps address TMEM lanes of the wrong quarter", "hutter_tc_kernels.cu",
     "((uint32_t)((warp & 3) * 32) << 16);", "((uint32_t)(((warp + 1) & 3) * 32) << 16);"),
]


def run_forward(src_root, n=130, grid=1):
    tmp = tempfile.mkdtemp()
    lib = os.path.join(tmp, "libm.so")
    r = subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-std=c++20", "-pthread", "-ffp-contract=off", "-x", "c++",
                        "-I", os.path.join(src_root, "apg_trajectory_tracking_b200", "csrc"),
                        os.path.join(src_root, "tests", "hostcheck", "hostcheck_tcsim.cpp"), "-o", lib],
                       capture_output=True, text=True)
    if r.returncode != 0:
        return "does not compile"
    fw = ctypes.CDLL(lib)
    h = 10
    params = B.default_init("quad", h, seed=1)
    case = SY.quad_case(n, h, 0.1, seed=1)
    flat = np.ascontiguousarray(torch.cat([p.reshape(-1) for p in params]).numpy(), np.float32)
    f32 = lambda t: np.ascontiguousarray(t.numpy(), np.float32)                  # noqa: E731
    ins, cur, inr, ref = f32(case["in_state"]), f32(case["cur"]), f32(case["in_ref"]), f32(case["ref"])
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)                              # noqa: E731
    blob = np.zeros(fw.hc_sim_blob_bytes(), np.uint8)
    fw.hc_sim_pack(p(flat), p(blob))
    nt = (n + 63) // 64
    nan = lambda r_: np.full(nt * r_ * 68, np.nan, np.float32)                   # noqa: E731
    x1, h1, h2, h3, act, sts = nan(224), nan(64), nan(64), nan(64), nan(40), nan(120)
    lossp = np.zeros(grid, np.float32)
    st, ac = np.zeros((n, h, 12), np.float32), np.zeros((n, h, 4), np.float32)
    err = ctypes.create_string_buffer(2048)
    ne = fw.hc_sim_forward(p(blob), p(ins), p(cur), p(inr), p(ref), n, ctypes.c_float(0.1), p(P.PHYS["quad"]()), grid,
                           p(x1), p(h1), p(h2), p(h3), p(act), p(sts), p(lossp), p(st), p(ac), err, 2048)
    want, _, _, want_act = O.concurrent_value_and_grad("quad", params, case["in_state"], case["cur"], case["in_ref"],
                                                       case["ref"], h, 0.1)
    if ne:
        return "model violation: " + err.value.decode()[:120]
    if not np.isfinite(lossp).all():
        return "NaN loss (poisoned / uninitialised accumulator)"
    rel = abs(float(lossp.sum()) - float(want)) / abs(float(want))
    da = float(np.abs(ac - want_act.detach().numpy()).max())
    return "results match the oracle" if rel <= 2e-5 and da <= 2e-5 else f"wrong results (loss rel {rel:.2e}, actions {da:.2e})"


def main():
    for name, fname, old, new in MUTANTS:
        scratch = tempfile.mkdtemp()
        for d in ("apg_trajectory_tracking_b200/csrc", "tests/hostcheck"):
            shutil.copytree(os.path.join(ROOT, d), os.path.join(scratch, d), ignore=shutil.ignore_patterns("*.so", "*.o"))
        if fname:
            path = os.path.join(scratch, "apg_trajectory_tracking_b200", "csrc", fname)
            s = open(path).read()
            assert s.count(old) >= 1, (name, "pattern not found")
            open(path, "w").write(s.replace(old, new, 1))
        print(f"{name}: {run_forward(scratch)}", flush=True)
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
    for name in ("host", "te", "tc", "dw"):
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
