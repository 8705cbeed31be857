"""torchrun helper of tests/test_gpu_api.py::test_two_gpu_sharded_gradient_equals_single_gpu"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apg_trajectory_tracking_b200 import dist as D, rollout as R, synthetic as SY  # noqa: E402
import bench  # noqa: E402


def main():
    rank, world = D.init_from_env("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    n, h, dt = 4096, 10, 0.1
    case = SY.quad_case(n, h, dt, seed=2)
    params = bench.default_init("quad", h, seed=0)
    flat = R.flatten_params(params).to(dev)
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    sh = {k: D.shard(v, rank, world).contiguous().to(dev) for k, v in case.items()}
    runner = R.Rollout(spec, sh["cur"].shape[0], dev)
    loss, grad = runner.value_and_grad(flat, sh["in_state"], sh["cur"], sh["in_ref"], sh["ref"])
    D.allreduce_sum_(grad)
    lt = loss.clone()
    dist.all_reduce(lt)
    if rank == 0:
        full = {k: v.to(dev) for k, v in case.items()}
        r1 = R.Rollout(spec, n, dev)
        l1, g1 = r1.value_and_grad(flat, full["in_state"], full["cur"], full["in_ref"], full["ref"])
        torch.cuda.synchronize()
        print(json.dumps({"grad_rel_err": float((grad - g1).norm() / g1.norm()),
                          "loss_rel_err": abs(float(lt) - float(l1)) / abs(float(l1))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
