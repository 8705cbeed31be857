"""quick device-timed sanity run of the quad concurrent train step (not the contract bench)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    h, dt = 10, 0.1
    dev = "cuda:0"
    case = SY.quad_case(n, h, dt, seed=1234, device=dev)
    torch.manual_seed(0)
    shapes = [(64, 15), (64,), (20, 9, 3), (20,), (64, 90), (64,), (64, 224), (64,), (64, 64), (64,), (64, 64), (64,), (40, 64), (40,)]
    params = [(torch.rand(*s) * 2 - 1) / (s[-1] if len(s) > 1 else 64) ** 0.5 for s in shapes]
    flat = R.flatten_params(params).to(dev)
    runner = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, dev)
    grad = torch.empty(runner.n_params, device=dev)
    for _ in range(3):
        runner.value_and_grad(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"], out=grad)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    for _ in range(10):
        ev[0].record()
        runner.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
        ev[1].record()
        runner.backward(1.0, out=grad)
        ev[2].record()
        torch.cuda.synchronize()
        tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[1].elapsed_time(ev[2]))
    tf.sort(); tb.sort()
    f, b = tf[len(tf) // 2], tb[len(tb) // 2]
    print(f"N={n} fwd {f:.3f} ms  adj {b:.3f} ms  total {f + b:.3f} ms  -> {n * h / ((f + b) * 1e-3):.3e} drone-steps/s  loss {runner.loss.item():.4f}")

main()
