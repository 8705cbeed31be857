"""A few untimed iterations of one bench workload (forward + adjoint through the C ABI) - the target of the short ncu
capture bench.py runs for `roofline.traffic` (and of ad-hoc ncu sessions).
    python tools/one_step.py <workload> [drones] [iterations]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench as B  # noqa: E402
from apg_trajectory_tracking_b200 import rollout as R  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "quad_concurrent"
    w = dict(B.WORKLOADS[name])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else w["n"]
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dev = "cuda:0"
    case = B.make_case(w, n, 1234, dev)
    flat = R.flatten_params(B.default_init(w["system"], w["h"], mode=w.get("mode", "concurrent"))).to(dev)
    r = R.Rollout(B.make_spec(w), n, dev)
    for _ in range(iters):
        r.value_and_grad(flat, case.get("in_state"), case["cur"], case.get("in_ref"), case.get("ref"), case.get("h0c0"))
    torch.cuda.synchronize()


main()
