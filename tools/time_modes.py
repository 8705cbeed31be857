"""device-timed sanity numbers for the other train modes (not the contract bench)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY


def rnd(shapes, fans, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(*s, generator=g) * 2 - 1) / f ** 0.5 for s, f in zip(shapes, fans)]


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
    print(f"{name}: N={n} h={h} fwd {f:.3f} ms adj {b:.3f} ms -> {n * h / ((f + b) * 1e-3):.3e} drone-steps/s, loss {runner.loss.item():.3f}", flush=True)


def main():
    dev = "cuda:0"
    n, h, dt = 65536, 10, 0.1
    case = SY.quad_case(n, 2 * h, dt, seed=1, device=dev)
    # autoregressive
    shapes = [(64, 15), (64,), (20, 9, 3), (20,), (64, 90), (64,), (64, 224), (64,), (64, 64), (64,), (64, 64), (64,), (4, 64), (4,)]
    fans = [15, 15, 27, 27, 90, 90, 224, 224, 64, 64, 64, 64, 64, 64]
    flat = R.flatten_params(rnd(shapes, fans)).to(dev)
    r = R.Rollout(R.RolloutSpec.quad_recurrent("autoregressive", h, dt), n, dev)
    timeit("quad autoregressive", r, flat, (None, case["cur"], case["in_ref"], case["ref"]), n, h, reps=5)
    del r
    torch.cuda.empty_cache()
    shapes = [(20, 9, 3), (20,), (64, 90), (64,), (4, 8), (4,), (32, 175), (32, 8), (32,), (32,)]
    fans = [27, 27, 90, 90, 8, 8, 8, 8, 8, 8]
    flat = R.flatten_params(rnd(shapes, fans)).to(dev)
    h0c0 = torch.randn(2, n, 8, device=dev)
    r = R.Rollout(R.RolloutSpec.quad_recurrent("lstm", h, dt), n, dev)
    timeit("quad lstm", r, flat, (None, case["cur"], case["in_ref"], case["ref"], h0c0), n, h, reps=5)
    del r
    torch.cuda.empty_cache()
    # wing h=20 N=131072
    n2, h2 = 131072, 20
    wc = SY.wing_case(n2, h2, 0.05, seed=1, device=dev)
    shapes = [(64, 9), (64,), (20, 3, 3), (20,), (64, 3), (64,), (64, 128), (64,), (64, 64), (64,), (64, 64), (64,), (80, 64), (80,)]
    fans = [9, 9, 9, 9, 3, 3, 128, 128, 64, 64, 64, 64, 64, 64]
    flat = R.flatten_params(rnd(shapes, fans)).to(dev)
    r = R.Rollout(R.RolloutSpec.wing_concurrent(h2, 0.05), n2, dev)
    timeit("wing concurrent", r, flat, (wc["in_state"], wc["cur"], wc["in_ref"], wc["ref"]), n2, h2)


main()
