"""per-kernel device times of the tcgen05 path (apg_debug_timing: CUDA events between the launches) for several N"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY, _capi
import bench

NAMES = ["pack", "fwd chain", "dyn+rev", "loss sum", "dx chain", "dw gemm", "reduce"]

def main():
    h, dt = 10, 0.1
    dev = "cuda:0"
    lib = _capi.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    raw = "--raw" in sys.argv            # RAW samples: prepare_data runs in the forward kernel's prologue
    for n in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [65536]:
        case = SY.quad_case(n, h, dt, seed=1234, device=dev)
        flat = R.flatten_params(bench.default_init("quad", h)).to(dev)
        r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, dev)
        args = (None, case["cur"], None, case["ref"]) if raw else (case["in_state"], case["cur"], case["in_ref"], case["ref"])
        for _ in range(3):
            r.value_and_grad(flat, *args)
        lib.apg_debug_timing(1)
        acc = np.zeros(7)
        reps = 10
        for _ in range(reps):
            flush.zero_()
            r.value_and_grad(flat, *args)
            out = np.zeros(7, np.float32)
            _capi.check(lib.apg_debug_kernel_times(ctypes.c_void_p(out.ctypes.data)))
            acc += out
        lib.apg_debug_timing(0)
        acc /= reps
        print(("raw " if raw else "") + f"N={n:7d} " + "  ".join(f"{nm} {1e3 * t:7.1f}us" for nm, t in zip(NAMES, acc)) + f"  total {1e3 * acc.sum():7.1f}us")

main()
