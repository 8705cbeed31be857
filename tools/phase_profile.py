"""per-phase cycle breakdown of the hutter forward / adjoint kernels (needs the -DAPG_PROFILE build:
python -m apg_trajectory_tracking_b200.build --profile; run with APG_B200_LIB=.../libapg_b200_prof.so)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY, _capi
import bench

FWD = ["setup", "wait inputs", "first layer", "trunk fc1..out (+wait act_empty)", "handoff + store drain"]
ADJ = ["setup", "wait dlog (dyn warps)", "wait h3", "dW out", "dX out", "wait h2", "dW fc3", "dX fc3", "wait h1", "dW fc2", "dX fc2",
       "wait X1", "dW fc1", "dX fc1", "wait inputs", "first-layer dW"]

def main():
    n, h, dt = 65536, 10, 0.1
    dev = "cuda:0"
    case = SY.quad_case(n, h, dt, seed=1234, device=dev)
    flat = R.flatten_params(bench.default_init("quad", h)).to(dev)
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, dev)
    for _ in range(3):
        r.value_and_grad(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    torch.cuda.synchronize()
    out = np.zeros((2, 148, 24), dtype=np.int64)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    rc = lib.apg_debug_profile(ctypes.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    for k, names in ((0, FWD), (1, ADJ)):
        m = out[k].mean(0)
        tot = m.sum()
        print(("forward" if k == 0 else "adjoint"), f"total {tot:.0f} cycles/CTA ({tot / 6.92:.0f} per tile)")
        for i, nm in enumerate(names):
            print(f"  {nm:18s} {m[i]:10.0f} {100 * m[i] / tot:5.1f}%")

main()
