"""device-timed forward / adjoint of the fixed-wing concurrent workload (BASELINE config 4 per-GPU size) and of the
quadrotor concurrent workload on the mma.sync kernels (APG_LEGACY_MMA=1 in the environment) - A/B timing helper"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY


def timeit(name, runner, flat, args, n, h, reps=10):
    grad = torch.empty(runner.n_params, device=flat.device)
    for _ in range(3):
        runner.value_and_grad(flat, *args, out=grad)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for _ in range(reps):
        ev[0].record(); runner.forward(flat, *args); ev[1].record(); runner.backward(1.0, out=grad); ev[2].record()
        torch.cuda.synchronize()
        tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[1].elapsed_time(ev[2]))
    tf.sort(); tb.sort()
    f, b = tf[len(tf) // 2], tb[len(tb) // 2]
    print(f"{name}: N={n} h={h} fwd {f:.3f} ms adj {b:.3f} ms -> {n * h / ((f + b) * 1e-3):.3e} drone-steps/s", flush=True)


dev = "cuda:0"
n, h = 131072, 20
wc = SY.wing_case(n, h, 0.05, seed=1, device=dev)
flat = R.flatten_params(B.default_init("wing", h, seed=1)).to(dev)
r = R.Rollout(R.RolloutSpec.wing_concurrent(h, 0.05), n, dev)
timeit("wing concurrent", r, flat, (wc["in_state"], wc["cur"], wc["in_ref"], wc["ref"]), n, h)
del r
n, h = 65536, 10
qc = SY.quad_case(n, h, 0.1, seed=1, device=dev)
flat = R.flatten_params(B.default_init("quad", h, seed=1)).to(dev)
r = R.Rollout(R.RolloutSpec.quad_concurrent(h, 0.1), n, dev)
timeit("quad concurrent (%s)" % ("tcgen05" if r.tcgen05 else "mma.sync"), r, flat,
       (qc["in_state"], qc["cur"], qc["in_ref"], qc["ref"]), n, h)
