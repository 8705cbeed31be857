"""world_size-2 gloo test of the multi-process host logic (runs on CPU): drone-axis sharding + ONE sum-allreduce of
the flat gradient + identical SGD-momentum step reproduces the single-process large-batch step.  Per-shard
gradients come from the CPU oracle here (the CUDA path is covered by the -m gpu tests)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import apg_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_problem():
    from apg_trajectory_tracking_b200 import synthetic as SY
    n, h, dt = 24, 6, 0.1
    case = SY.quad_case(n, h, dt, seed=3)
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 15), (64,), (20, 9, 3), (20,), (64, 9 * h), (64,), (64, 64 + 20 * (h - 2)), (64,), (64, 64), (64,),
              (64, 64), (64,), (4 * h, 64), (4 * h,)]
    params = [(torch.rand(*s, generator=g) * 2 - 1) * 0.2 for s in shapes]
    return case, params, h, dt


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from apg_trajectory_tracking_b200 import dist as D
    r, w = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    case, params, h, dt = _make_problem()
    sh = {k: D.shard(v, rank, world) for k, v in case.items()}
    flat = torch.cat([p.reshape(-1) for p in params])
    buf = torch.zeros_like(flat)
    losses = []
    for it in range(2):
        ps, o = [], 0
        for p in params:
            ps.append(flat[o:o + p.numel()].view_as(p)); o += p.numel()
        loss, grads, _, _ = O.concurrent_value_and_grad("quad", ps, sh["in_state"], sh["cur"], sh["in_ref"], sh["ref"],
                                                        h, dt)
        g = torch.cat([(gr if gr is not None else torch.zeros_like(p)).reshape(-1) for gr, p in zip(grads, params)])
        D.allreduce_sum_(g)
        lt = loss.clone()
        dist.all_reduce(lt)
        losses.append(float(lt))
        D.sgd_momentum_step_(flat, g, buf, 1e-5)
    if rank == 0:
        torch.save({"flat": flat, "losses": losses}, out)
    # every rank holds identical parameters after the update
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert all(torch.equal(gathered[0], x) for x in gathered)
    dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_process(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    case, params, h, dt = _make_problem()
    ps, bufs = [p.clone() for p in params], [None] * len(params)
    losses = []
    for it in range(2):
        loss, grads, _, _ = O.concurrent_value_and_grad("quad", ps, case["in_state"], case["cur"], case["in_ref"],
                                                        case["ref"], h, dt)
        losses.append(float(loss))
        ps, bufs = O.sgd_momentum_step(ps, grads, bufs, 1e-5)
    flat = torch.cat([p.reshape(-1) for p in ps])
    for a, b in zip(res["losses"], losses):
        assert abs(a - b) <= 2e-6 * abs(b)
    assert float((res["flat"] - flat).abs().max()) <= 1e-6 * float(flat.abs().max())


def test_shard_bounds_cover_the_batch():
    from apg_trajectory_tracking_b200 import dist as D
    for n, w in ((65536, 8), (10, 3), (7, 8)):
        b = [D.shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _agree_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo")
    import bench
    # the ranks see different elapsed times / step times: they must still agree on ONE count
    n = bench.agree_extra_steps(1.0 if rank == 0 else 0.2, 1e-3 if rank == 0 else 3e-3, "cpu", True)
    m = bench.agree_extra_steps(0.0, 1e-3, "cpu", True)
    torch.save((n, m), f"{out}.{rank}")
    dist.destroy_process_group()


def test_bench_extra_steps_are_rank_consistent(tmp_path):
    """regression test: the clock-sampling extension of bench.py must run the same number of (all-reducing) steps
    on every rank"""
    out = str(tmp_path / "agree")
    mp.spawn(_agree_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0 == r1 and r0[0] == 1001 and r0[1] == 0
