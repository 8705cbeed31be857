"""per-role cycle breakdown of the tq kernels (needs the -DAPG_PROFILE build:
python -m apg_trajectory_tracking_b200.build --profile; run with APG_B200_LIB=.../libapg_b200_prof.so)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY, _capi
import bench

def main():
    n, h, dt = 65536, 10, 0.1
    dev = "cuda:0"
    case = SY.quad_case(n, h, dt, seed=1234, device=dev)
    flat = R.flatten_params(bench.default_init("quad", h)).to(dev)
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, dev)
    for _ in range(3):
        r.value_and_grad(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    torch.cuda.synchronize()
    out = np.zeros((3, 148, 16), dtype=np.int64)
    out_dw = np.zeros((3, 148, 16), dtype=np.int64)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    rc = lib.apg_debug_profile_tq_chain(ctypes.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    rc = lib.apg_debug_profile_tq_dw(ctypes.c_void_p(out_dw.ctypes.data))
    assert rc == 0, rc
    out[2] = out_dw[2]
    names = {0: ["epi wait_d", "epi work"], 1: ["epi wait_d", "epi work"],
             2: ["prod wait rfree", "prod issue", "mma wait lo_ready", "mma issue", "conv wait full", "conv wait lo_free", "conv work"]}
    for k, kn in ((0, "tq_fwd"), (1, "tq_dx"), (2, "tq_dw")):
        m = out[k].mean(0); mx = out[k].max(0)
        print(kn)
        for i, nm in enumerate(names[k]):
            print(f"  {nm:20s} mean {m[i]:10.0f}  max {mx[i]:10.0f} cycles")

main()
