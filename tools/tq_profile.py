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
    out = np.zeros((3, 148, 20), dtype=np.int64)
    out_dw = np.zeros((3, 148, 20), dtype=np.int64)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    rc = lib.apg_debug_profile_tq_chain(ctypes.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    rc = lib.apg_debug_profile_tq_dw(ctypes.c_void_p(out_dw.ctypes.data))
    assert rc == 0, rc
    out[2] = out_dw[2]
    names = {0: ["epi wait_d", "epi other work", "dense: tmem ld16+wait", "dense: tanh/stash/split", "dense: tmem st16 x2", "dense: wait::st+fence+arrive"], 1: ["epi wait_d", "epi work"],
             2: ["prod wait rfree", "prod issue", "mma wait slot ready", "mma issue", "A feeder wait full", "A feeder wait slot free",
                 "A feeder work", "A feeder flush (wait done + TMEM -> partial)", "setup", "main loop + flushes", "tail",
                 "B feeder wait full", "B feeder wait slot free", "B feeder work", "B feeder flush", "-", "-", "-", "-",
                 "mma wait a_ready"]}
    t = out[0][:, 6:12].astype(np.float64)
    seg = [("H1 epilogue of warp 0 (ld, tanh, stash, st, arrive)", t[:, 1] - t[:, 0]),
           ("warp 0 arrived -> issuer saw all 8 warps", t[:, 2] - t[:, 1]),
           ("issue 24 MMAs", t[:, 3] - t[:, 2]), ("commit instruction", t[:, 4] - t[:, 3]),
           ("commit issued -> warp 0 sees d_ready (tensor exec + wake-up)", t[:, 5] - t[:, 4]),
           ("whole hand-off", t[:, 5] - t[:, 0])]
    print("one hand-off of tq_fwd (tile 0, slot 0, H1 -> fc2), cycles, mean / max over CTAs")
    for nm, v in seg:
        print(f"  {nm:62s} {v.mean():8.0f} {v.max():8.0f}")
    for k, kn in ((0, "tq_fwd"), (1, "tq_dx"), (2, "tq_dw")):
        g0, g1 = out[k][:, 16].astype(np.float64), out[k][:, 17].astype(np.float64)
        cyc = (out[k][:, 15] if k < 2 else out[k][:, 18]).astype(np.float64)
        print(f"{kn}: %globaltimer span first CTA entry -> last CTA end {(g1.max() - g0.min()) / 1e3:.1f} us; CTA entry spread "
              f"{(g0.max() - g0.min()) / 1e3:.1f} us; CTA body mean {(g1 - g0).mean() / 1e3:.1f} us max {(g1 - g0).max() / 1e3:.1f} us; "
              f"SM clock from cycles / ns: {(cyc / np.maximum(g1 - g0, 1)).mean() * 1e3:.0f} MHz")
    for k, kn in ((0, "tq_fwd"), (1, "tq_dx")):
        tl = out[k][:, 12:16].astype(np.float64)
        print(f"{kn} timeline of thread 0, cycles from kernel entry, mean / min / max over CTAs")
        for i, nm in enumerate(["setup done (TMEM, barriers, bulk copy issued)", "first accumulator ready",
                                "warp 0 finished its tiles", "all warps finished"]):
            print(f"  {nm:50s} {tl[:, i].mean():9.0f} {tl[:, i].min():9.0f} {tl[:, i].max():9.0f}")
    for k, kn in ((0, "tq_fwd"), (1, "tq_dx"), (2, "tq_dw")):
        m = out[k].mean(0); mx = out[k].max(0)
        print(kn)
        for i, nm in enumerate(names[k]):
            print(f"  {nm:44s} mean {m[i]:10.0f}  max {mx[i]:10.0f} cycles")

main()
