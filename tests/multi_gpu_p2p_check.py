"""torchrun helper: the gradient exchange over peer memory (dist.PeerGradExchange, csrc/p2p_kernels.cu) against the
NCCL all-reduce path on the same shards - a few full train steps each, from the same initial parameters.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 tests/multi_gpu_p2p_check.py
Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apg_trajectory_tracking_b200 import dist as D, rollout as R, synthetic as SY, train as T  # noqa: E402
import bench  # noqa: E402


def main():
    rank, world = D.init_from_env("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    n, h, dt, steps = 8192, 10, 0.1, 4
    case = SY.quad_case(n, h, dt, seed=2)
    params = bench.default_init("quad", h, seed=0)
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    sh = {k: D.shard(v, rank, world).contiguous().to(dev) for k, v in case.items()}
    args = (sh["in_state"], sh["cur"], sh["in_ref"], sh["ref"])
    out = {}
    results = {}
    for name, peer in (("nccl", False), ("p2p", True)):
        st = T.FusedTrainStep(params, spec, sh["cur"].shape[0], lr=1e-7, device=dev, peer_exchange=peer)
        losses = []
        for _ in range(steps):
            losses.append(float(st.step(*args).item()))
        _, g = st.value_and_grad(*args)                    # the gather-only form (no update)
        torch.cuda.synchronize()
        results[name] = (st.flat.clone(), g.clone(), losses)
    pn, gn, ln = results["nccl"]
    pp, gp, lp = results["p2p"]
    gathered = [torch.empty_like(pp) for _ in range(world)]
    dist.all_gather(gathered, pp)
    out["params_rel_diff_vs_nccl"] = float((pp - pn).norm() / pn.norm())
    out["grad_rel_diff_vs_nccl"] = float((gp - gn).norm() / gn.norm())
    out["loss_rel_diff_vs_nccl"] = max(abs(a - b) / abs(b) for a, b in zip(lp, ln))
    out["p2p_params_bitwise_equal_across_ranks"] = all(bool(torch.equal(gathered[0], t)) for t in gathered)
    out["finite"] = bool(torch.isfinite(pp).all() and torch.isfinite(gp).all())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
