#!/usr/bin/env python
"""Benchmark of the batched differentiable rollout train step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

metric   drone-steps/sec (fwd+bwd) = N_total * h / time of one train iteration
         iteration = policy forward -> sigmoid -> h dynamics steps -> tracking loss -> full backward to the weight
         gradient [-> NCCL sum-allreduce of the flat gradient when G > 1] -> SGD(momentum) step
workload BASELINE configs[1]: quadrotor, concurrent MLP Net(15,10,9,40), horizon 10, N = 65536 drones PER GPU
         (weak scaling), synthetic polynomial reference trajectories, default-initialised policy.
One "step" = one train iteration over the whole (per-GPU) batch.  Device-timed with CUDA events around every
step on the launching stream; the L2 is flushed (256 MiB write) between timed steps, outside the timed region.
"""
import argparse
import datetime
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (system, mode, h, dt, N per GPU, algorithmic bytes per drone-step (fwd+bwd), flops per drone-step)
    "quad_concurrent": dict(system="quad", mode="concurrent", h=10, dt=0.1, n=65536, bytes_per_step=165.6,
                            flops_per_step=17.7e3, label="quad concurrent MLP Net(15,10,9,40) h=10 N=65536/GPU"),
    "wing_concurrent": dict(system="wing", mode="concurrent", h=20, dt=0.05, n=131072, bytes_per_step=33.6,
                            flops_per_step=8.1e3, label="fixed-wing MLP Net(9,1,3,80) h=20 N=131072/GPU"),
    "cartpole_concurrent": dict(system="cartpole", mode="concurrent", h=5, dt=0.05, n=128, bytes_per_step=12.8,
                                flops_per_step=10e3, label="cartpole MLP Net(4,5) h=5 B=128"),
    "quad_autoregressive": dict(system="quad", mode="autoregressive", h=10, dt=0.1, n=65536, bytes_per_step=225.6,
                                flops_per_step=169e3, label="quad autoregressive Net(15,10,9,4) h=10 N=65536/GPU"),
    "quad_lstm": dict(system="quad", mode="lstm", h=10, dt=0.1, n=32768, bytes_per_step=238.4, flops_per_step=62e3,
                      label="quad LSTM_NEW(15,10,9,4) h=10 N=32768/GPU"),
}
LR = {"quad": 1e-5, "wing": 1e-4, "cartpole": 1e-5}


def hutter_shapes(system, h, mode="concurrent"):
    if mode == "autoregressive":
        return [(64, 15), (64,), (20, 9, 3), (20,), (64, 9 * h), (64,), (64, 64 + 20 * (h - 2)), (64,), (64, 64), (64,),
                (64, 64), (64,), (4, 64), (4,)]
    if mode == "lstm":
        return [(20, 9, 3), (20,), (64, 9 * h), (64,), (4, 8), (4,), (32, 15 + 20 * (h - 2)), (32, 8), (32,), (32,)]
    if system == "quad":
        return [(64, 15), (64,), (20, 9, 3), (20,), (64, 9 * h), (64,), (64, 64 + 20 * (h - 2)), (64,), (64, 64), (64,),
                (64, 64), (64,), (4 * h, 64), (4 * h,)]
    if system == "wing":
        return [(64, 9), (64,), (20, 3, 3), (20,), (64, 3), (64,), (64, 128), (64,), (64, 64), (64,), (64, 64), (64,),
                (4 * h, 64), (4 * h,)]
    return [(32, 4), (32,), (64, 32), (64,), (64, 64), (64,), (32, 64), (32,), (h, 32), (h,)]


def default_init(system, h, seed=0, mode="concurrent"):
    """PyTorch default (kaiming-uniform, a=sqrt(5)) initialisation == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    weights and biases of Linear / Conv1d; drawn tensor by tensor from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    out, fan_in = [], 1
    for s in hutter_shapes(system, h, mode):
        if len(s) > 1:
            fan_in = 1
            for d in s[1:]:
                fan_in *= d
        if mode == "lstm" and s[0] == 32:
            fan_in = 8                       # nn.LSTMCell initialises everything with U(-1/sqrt(hidden), ..)
        out.append((torch.rand(*s, generator=g) * 2 - 1) / fan_in ** 0.5)
    return out


def make_case(w, n, seed, device):
    from apg_trajectory_tracking_b200 import synthetic as SY
    if w.get("mode", "concurrent") != "concurrent":
        c = SY.quad_case(n, 2 * w["h"], w["dt"], seed=seed, device=device)
        c["in_state"] = None
        if w["mode"] == "lstm":
            g = torch.Generator().manual_seed(seed + 77)
            c["h0c0"] = torch.stack((torch.randn(n, 8, generator=g), torch.randn(n, 8, generator=g)), 0).to(device)
        return c
    if w["system"] == "quad":
        return SY.quad_case(n, w["h"], w["dt"], seed=seed, device=device)
    if w["system"] == "wing":
        return SY.wing_case(n, w["h"], w["dt"], seed=seed, device=device)
    c = SY.cartpole_case(n, seed=seed, device=device)
    c["in_ref"], c["ref"] = None, None
    return c


def make_spec(w):
    from apg_trajectory_tracking_b200 import rollout as R
    if w.get("mode", "concurrent") != "concurrent":
        return R.RolloutSpec.quad_recurrent(w["mode"], w["h"], w["dt"])
    if w["system"] == "quad":
        return R.RolloutSpec.quad_concurrent(w["h"], w["dt"])
    if w["system"] == "wing":
        return R.RolloutSpec.wing_concurrent(w["h"], w["dt"])
    return R.RolloutSpec.cartpole_concurrent(w["h"], w["dt"])


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower() == "active":
                    reasons.add(nm)
        sm.sort()
        # median of the upper half: samples taken between steps (L2 flush, host gaps) would bias a plain median
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def cpu_reference_step(w, n, params, case, threads):
    """One train iteration of the reference algorithm on the CPU (oracle port: same op chain as the reference's
    PyTorch code, autograd tape and all) -- returns a closure running one step."""
    from oracle import apg_oracle as O
    torch.set_num_threads(threads)
    ps = [p.clone() for p in params]
    bufs = [None] * len(ps)
    lr = LR[w["system"]]

    def step():
        nonlocal ps, bufs
        if w.get("mode", "concurrent") == "concurrent":
            loss, grads, _, _ = O.concurrent_value_and_grad(w["system"], ps, case["in_state"], case["cur"],
                                                            case["in_ref"], case["ref"], w["h"], w["dt"])
        else:
            hc = (case["h0c0"][0], case["h0c0"][1]) if w["mode"] == "lstm" else None
            loss, grads, _, _ = O.recurrent_value_and_grad(w["mode"], ps, case["cur"], case["in_ref"], case["ref"],
                                                           w["h"], w["dt"], hc0=hc)
        ps, bufs = O.sgd_momentum_step(ps, grads, bufs, lr)
        return float(loss)
    return step


def agree_extra_steps(need_s, t_step_s, device, distributed):
    """Number of extra untimed steps to run after the timed region, IDENTICAL on every rank: each rank proposes
    ceil(need / step time) and the MAX over ranks is taken with one all-reduce.  (Every step contains a gradient
    all-reduce; a per-rank time-based loop would issue mismatched collectives and deadlock.)"""
    n = min(int(need_s / max(t_step_s, 1e-5)) + 1, 20000) if need_s > 0 else 0
    t = torch.tensor([n], device=device, dtype=torch.int64)
    if distributed:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def raw_host_inputs(w, host):
    """the RAW samples FusedTrainStep.step_host takes for workload w, picked from the pinned host case"""
    kw = {"cur": host["cur"]}
    if w["system"] == "quad":
        kw["ref"] = host["ref"]
        if w.get("mode") == "lstm":
            kw["h0c0"] = host["h0c0"]
    elif w["system"] == "wing":
        kw["target"] = host["target"]
    return kw


def live_dram_traffic(workload, n, kernel):
    """dram bytes (read + write) of one launch of `kernel` in workload `workload` at n drones, from a short ncu run of
    tools/one_step.py in a child process; None if ncu is missing / fails / finds no such launch"""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           "-k", "regex:" + kernel, "-s", "2", "-c", "1", sys.executable, os.path.join(ROOT, "tools", "one_step.py"),
           workload, str(n), "4"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
        rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
        hdr = rows[0]
        total = 0.0
        for r in rows[1:]:
            if len(r) != len(hdr):
                continue
            unit, val = r[hdr.index("Metric Unit")], float(r[hdr.index("Metric Value")].replace(",", ""))
            total += val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        return total if total > 0 else None
    except Exception:                                               # noqa: BLE001
        return None


def measure_e2e_raw(stepper, w, host, case, e2e_steps, world, dev, barrier):
    """e2e through FusedTrainStep.step_host: per step H2D of the raw samples (chunked, copy stream), device-side
    prepare, forward, adjoint, [allreduce], SGD, D2H of the loss.  Before timing, loss and gradient of this path are
    checked against the prepared-input path on the same parameters; ranks agree on the outcome and on the chunk
    size (each step contains an all-reduce, so every rank must take the same branch)."""
    import torch.distributed as dist
    n, h = w["n"], w["h"]
    kw = raw_host_inputs(w, host)
    h2d = sum(v.numel() * 4 for v in kw.values())

    def agree(x, op):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    ok, check = 1.0, {}
    try:
        la, ga = stepper.runner.value_and_grad(stepper.flat, case["in_state"], case["cur"], case.get("in_ref"),
                                               case.get("ref"), case.get("h0c0"))
        la, ga = float(la.item()), ga.clone()
        # whole batch in one chunk: same kernels, same reduction order as the prepared-input call.  (Cutting the batch
        # into chunks changes the order of an fp32 sum whose terms cancel ~300-fold at this size; that alone moves the
        # gradient by 1e-5 .. 1e-3 of its norm - reported as grad_rel_l2_default_chunks, not judged here; the chunked
        # path is checked against the oracle at sizes where the oracle is exact enough, tests/test_step_host_*.)
        lb, gb = stepper.value_and_grad_host(allreduce=False, chunk=0, **kw)
        lb, gb = float(lb.item()), gb.clone()
        gerr = float((gb - ga).norm() / (ga.norm() + 1e-30))
        check = {"loss_prepared": la, "loss_raw": lb, "grad_rel_l2": gerr}
        _, gchunks = stepper.value_and_grad_host(allreduce=False, **kw)
        check["grad_rel_l2_default_chunks"] = float((gchunks - ga).norm() / (ga.norm() + 1e-30))
        if w["system"] == "quad" and w.get("mode", "concurrent") == "concurrent" and torch.device(dev).type == "cuda":
            # the prepared tensors of `case` were made on the host; the raw path computes the same features in the
            # kernel prologue and the two differ by ulps - which the batch gradient (a sum with ~300x cancellation)
            # amplifies to ~1e-4 of its norm.  Like for like: the prepared-input path on the tensors the DEVICE
            # prepare kernels make from the same raw samples.
            from apg_trajectory_tracking_b200 import prepare as PR
            prep = PR.prepare_quad(kw["cur"].to(dev), kw["ref"].to(dev))
            lc, gc = stepper.runner.value_and_grad(stepper.flat, prep["in_state"], prep["cur"], prep["in_ref"],
                                                   prep["ref"])
            check["grad_rel_l2_vs_host_prepared_inputs"] = gerr
            la, gerr = float(lc.item()), float((gb - gc).norm() / (gc.norm() + 1e-30))
            check.update({"loss_prepared": la, "grad_rel_l2": gerr})
        if not (abs(la - lb) <= 1e-5 * abs(la) and gerr <= 1e-4):
            ok = 0.0
    except Exception as ex:                                       # noqa: BLE001
        ok, check = 0.0, {"error": f"{type(ex).__name__}: {ex}"[:300]}
    if agree(ok, "MIN") < 1.0:
        return {"ok": False, "check": check}

    # chunk size: best of a few candidates (whole batch, 4, 2, 1 waves of 64-drone tiles), agreed over the ranks
    wave = 64 * max(1, stepper.runner.lib.apg_sm_count())
    cands = [c for c in (0, 4 * wave, 2 * wave, wave) if c == 0 or c < n]
    times = {}
    for c in cands:
        for _ in range(2):
            float(stepper.step_host(chunk=c, **kw).item())
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            float(stepper.step_host(chunk=c, **kw).item())
        barrier()
        times[c] = agree((time.perf_counter() - t0) / 3, "MAX")
    chunk = min(times, key=times.get)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        float(stepper.step_host(chunk=chunk, **kw).item())
    barrier()
    t_sync = agree((time.perf_counter() - t0) / e2e_steps, "MAX")
    # the same loop with the loss of step i read after step i+1 has been issued (step_host_async: alternating
    # staging sets, the H2D copies of the next step overlap the last kernels of this one); every step still copies
    # its inputs and has its loss read on the host inside the timed region
    barrier()
    t0 = time.perf_counter()
    prev, sink = None, 0.0
    for _ in range(e2e_steps):
        hnd = stepper.step_host_async(chunk=chunk, **kw)
        if prev is not None:
            sink += prev.item()
        prev = hnd
    sink += prev.item()
    barrier()
    t_async = agree((time.perf_counter() - t0) / e2e_steps, "MAX")
    t = t_sync                          # headline: every step's loss is read before the next step is issued
    api = ("apg_trajectory_tracking_b200.train.FusedTrainStep.step_host / step_host_async(raw host samples): "
           "chunked H2D on a copy stream overlapped with prepare + forward + adjoint of the previous chunk; "
           "value = loss read on the host after EVERY step (value_loss_read_one_step_late: the pipelined "
           "variant, reported next to it)")
    extra = {}
    if world == 1 and torch.device(dev).type == "cuda":
        # The same step as ONE CUDA-graph launch (FusedTrainStep.capture_host / replay_host): eagerly every chunk costs
        # ~15 CUDA calls issued from Python and beyond two chunks that CPU time sets the step time; captured, the batch
        # can be cut finely and little is left to compute after the last copy has landed.  Every replay still copies the
        # step's inputs from the pinned host tensors and the loss is read on the host before the next replay.
        try:
            gtimes = {}
            for c in [c for c in (2 * wave, wave, wave // 2, wave // 4) if 0 < c < n]:
                stepper.capture_host(chunk=c, warmup=1, **kw)
                for _ in range(2):
                    stepper.replay_host().item()
                t0 = time.perf_counter()
                for _ in range(3):
                    stepper.replay_host().item()
                gtimes[c] = (time.perf_counter() - t0) / 3
            if gtimes:
                gchunk = min(gtimes, key=gtimes.get)
                stepper.capture_host(chunk=gchunk, warmup=1, **kw)
                stepper.replay_host().item()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    stepper.replay_host().item()
                t_graph = (time.perf_counter() - t0) / e2e_steps
                extra = {"value_eager_step_host": world * n * h / t_sync, "value_graph_replay_host": world * n * h / t_graph,
                         "graph_chunk": gchunk,
                         "graph_chunk_candidates_ms": {str(c): round(v * 1e3, 4) for c, v in gtimes.items()}}
                if t_graph < t:
                    t = t_graph
                    api = ("apg_trajectory_tracking_b200.train.FusedTrainStep.capture_host / replay_host(raw pinned host "
                           "samples): one CUDA-graph launch per step = chunked H2D (parallel branch) + forward + adjoint "
                           "per chunk + SGD + D2H of the loss; value = loss read on the host after EVERY step "
                           "(value_eager_step_host: the same step issued call by call)")
        except Exception as ex:                                   # noqa: BLE001
            extra = {"graph_error": f"{type(ex).__name__}: {ex}"[:300]}
    out = {"ok": True, "value": world * n * h / t, "unit": "drone-steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "steps": e2e_steps, "chunk": chunk if chunk else n,
           "chunk_candidates_ms": {str(c if c else n): round(v * 1e3, 4) for c, v in times.items()}, "check": check,
           "value_loss_read_every_step": world * n * h / t,
           "value_loss_read_one_step_late": world * n * h / t_async,
           "api": api}
    out.update(extra)
    return out


def physical_cores():
    try:
        import psutil
        c = psutil.cpu_count(logical=False)
        if c:
            return int(c)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args, w, rank):
    """--impl reference: the reference's CPU algorithm for the path (oracle port; the Python reference itself cannot
    travel to the GPU box), all host cores, rank 0 only."""
    if rank != 0:
        return
    threads = physical_cores()
    n = w["n"] if w.get("mode", "concurrent") == "concurrent" else min(w["n"], 8192)   # bounded sample per step
    case = make_case(w, n, 1234, "cpu")
    params = default_init(w["system"], w["h"], mode=w.get("mode", "concurrent"))
    step = cpu_reference_step(w, n, params, case, threads)
    step()
    # bounded sample: shrink the per-step batch until warm-up + K steps fit the budget (throughput is per drone-step,
    # so a smaller sample measures the same metric)
    budget_s = float(os.environ.get("APG_BENCH_REFERENCE_BUDGET_S", "240"))
    while n > 1024:
        t0 = time.perf_counter()
        step()
        if (time.perf_counter() - t0) * (args.steps + args.warmup) <= budget_s:
            break
        n //= 2
        case = make_case(w, n, 1234, "cpu")
        step = cpu_reference_step(w, n, params, case, threads)
        step()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt_s = (time.perf_counter() - t0) / max(args.steps, 1)
    value = n * w["h"] / dt_s
    line = {
        "impl": "reference", "metric": "drone-steps/sec (fwd+bwd)", "value": value, "unit": "drone-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["label"], "horizon": w["h"], "n_drones": n,
                   "note": "CPU arm: oracle port of the reference's PyTorch op chain (autograd tape) + SGD step"},
        "cpu_baseline": {"value": value, "unit": "drone-steps/s", "cores": threads, "kind": "port",
                         "sample": f"N={n} drones x h={w['h']} per step, {args.steps} iterations"},
        "e2e": {"value": value, "unit": "drone-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quad_concurrent", choices=sorted(WORKLOADS))
    ap.add_argument("--drones-per-gpu", "--n", dest="n", type=int, default=0, help="override drones per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-raw-e2e", action="store_true", help="skip the raw-sample (step_host) end-to-end arm")
    ap.add_argument("--no-live-traffic", action="store_true",
                    help="do not measure roofline.traffic with an ncu child process (use profiles/ncu_traffic.json)")
    ap.add_argument("--p2p-grad", action="store_true",
                    help="N>1: exchange the gradient with the package's own kernels over NVLink peer memory "
                         "(APG_P2P_GRAD=1) instead of the NCCL all-reduce")
    ap.add_argument("--nccl-grad", action="store_true",
                    help="N>1: force the NCCL all-reduce + torch optimizer ops (APG_P2P_GRAD=0); default on the tcgen05 "
                         "path: the peer-memory kernels")
    ap.add_argument("--legacy-mma", action="store_true",
                    help="quad_concurrent: run the mma.sync kernels (APG_LEGACY_MMA=1) instead of the tcgen05 path")
    args = ap.parse_args()
    # hard wall-clock bound for the whole process: dump the Python stacks and exit instead of hanging a GPU box
    faulthandler.dump_traceback_later(int(os.environ.get("APG_BENCH_WATCHDOG_S", "1500")), exit=True)
    w = dict(WORKLOADS[args.workload])
    if args.n:
        w["n"] = args.n
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    if args.p2p_grad:
        os.environ["APG_P2P_GRAD"] = "1"
    if args.nccl_grad:
        os.environ["APG_P2P_GRAD"] = "0"
    if args.legacy_mma:
        os.environ["APG_LEGACY_MMA"] = "1"

    if args.warmup < 3:
        args.warmup = 3
    import torch.distributed as dist
    from apg_trajectory_tracking_b200 import train as T
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a mismatched or stuck collective must end the run (NCCL watchdog abort), not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    n, h = w["n"], w["h"]
    case = make_case(w, n, 1234 + rank, dev)
    params = default_init(w["system"], h, mode=w.get("mode", "concurrent"))
    stepper = T.FusedTrainStep(params, make_spec(w), n, lr=LR[w["system"]], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    args_step = (case["in_state"], case["cur"], case.get("in_ref"), case.get("ref"), case.get("h0c0"))

    # clocks are sampled from the warm-up to the end of the timed region (and, if that is shorter than ~1.5 s,
    # over extra identical untimed steps appended after it) so that a 200 ms sampler sees the loaded clocks
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_sampler0 = time.perf_counter()
    for _ in range(args.warmup):
        stepper.step(*args_step)
    barrier()

    # ---- device-timed region: K steps, an event pair (+ one in the middle) around each
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        if flush_buf is not None:
            flush_buf.zero_()
        e = ev[i]
        e[0].record()
        loss, _, _ = stepper.runner.forward(stepper.flat, *args_step)
        e[1].record()
        if stepper.peer is not None:
            # adjoint whose gradient reduction scatters into every rank's slot over peer memory, then one kernel:
            # wait for all ranks, rank-ordered sum, SGD(momentum) update
            comm, local_set = stepper.peer.next_step()
            stepper.runner.backward_p2p(comm)
            e[2].record()
            stepper.peer.gather(comm, local_set, grad_out=stepper.grad, params=stepper.flat,
                                momentum_buf=stepper.buf, lr=stepper.lr, momentum=stepper.momentum)
        elif stepper.runner.tcgen05 and world == 1:
            # tcgen05 path, one device: the SGD(momentum) update rides on the gradient reduction (one launch)
            stepper.runner.backward_sgd(stepper.flat, stepper.buf, stepper.lr, stepper.momentum, out=stepper.grad)
            e[2].record()
        else:
            stepper.runner.backward(1.0, out=stepper.grad)
            e[2].record()
            if world > 1:
                dist.all_reduce(stepper.grad)
            stepper.buf.mul_(stepper.momentum).add_(stepper.grad)
            stepper.flat.add_(stepper.buf, alpha=-stepper.lr)
        e[3].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    # Extra identical untimed steps so that the 200 ms clock sampler sees >= ~1.5 s of load.  The count MUST be the
    # same on every rank (each step contains an all-reduce): it is derived from the measured step time and agreed
    # with one MAX all-reduce -- a time-based loop per rank would issue mismatched collectives and deadlock.
    extra_steps = agree_extra_steps(max(0.0, 1.5 - (time.perf_counter() - t_sampler0)),
                                    t_wall / max(args.steps, 1), dev, world > 1)
    for i in range(extra_steps):
        stepper.step(*args_step)
        if i % 50 == 49:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    clocks["note"] = f"sampled over warm-up + timed region + {extra_steps} identical untimed steps"
    # ---- per-kernel device times of the tcgen05 path (CUDA events between the launches inside the C-ABI calls,
    #      apg_debug_timing): 10 extra steps, outside every timed region, L2 flushed before each like the timed steps
    kernel_ms = None
    if stepper.runner.tcgen05 and stepper.peer is None:
        import ctypes
        import numpy as np
        from apg_trajectory_tracking_b200 import _capi
        lib = _capi.lib()
        lib.apg_debug_timing(1)
        acc, reps = np.zeros(7), 10
        for _ in range(reps):
            if flush_buf is not None:
                flush_buf.zero_()
            stepper.runner.forward(stepper.flat, *args_step)
            if world == 1:      # the launches of the timed loop: the optimizer step rides on the reduction
                stepper.runner.backward_sgd(stepper.flat, stepper.buf, stepper.lr, stepper.momentum, out=stepper.grad)
            else:
                stepper.runner.backward(1.0, out=stepper.grad)
            out = np.zeros(7, np.float32)
            _capi.check(lib.apg_debug_kernel_times(ctypes.c_void_p(out.ctypes.data)))
            acc += out
        lib.apg_debug_timing(0)
        # (slot 3 is the interval between two back-to-back event records - the cost of the instrumentation itself,
        # ~2.5 us, which every other entry contains once: the loss sum it used to time is part of tq_dyn_kernel now)
        names = ["tq_pack_kernel", "tq_fwd_kernel", "tq_dyn_kernel", "event_pair_overhead", "tq_dx_kernel",
                 "tq_dw_kernel", "apg_reduce4_sgd_kernel" if world == 1 else "apg_reduce4_kernel"]
        kernel_ms = {k: float(v) / reps for k, v in zip(names, acc)}
    t_step = [e[0].elapsed_time(e[3]) for e in ev]
    t_fwd = [e[0].elapsed_time(e[1]) for e in ev]
    t_adj = [e[1].elapsed_time(e[2]) for e in ev]
    ms = sum(t_step) / len(t_step)
    ms_t = torch.tensor([ms, sum(t_fwd) / len(t_fwd), sum(t_adj) / len(t_adj)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms, ms_fwd, ms_adj = [float(x) for x in ms_t.tolist()]
    value = world * n * h / (ms * 1e-3)
    final_loss = float(loss.item())

    # ---- end-to-end: the same train step through the public API with HOST (pinned) inputs every step:
    #      H2D of in_state/cur/in_ref/ref + forward + adjoint (+ allreduce) + SGD + D2H of the loss
    host = {k: (v.cpu().pin_memory() if v is not None else None) for k, v in case.items()}
    e2e_steps = max(3, min(args.steps, 10))
    hargs = (host["in_state"], host["cur"], host.get("in_ref"), host.get("ref"), host.get("h0c0"))
    h2d = sum(v.numel() * 4 for v in hargs if v is not None)
    for _ in range(2):
        float(stepper.step(*hargs).item())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        float(stepper.step(*hargs).item())
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * h / float(e2e_t.item())

    # ---- CPU baseline next to it (rank 0, single GPU runs only): the oracle port on all host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = physical_cores()
        ncpu = min(n, 65536) if w.get("mode", "concurrent") == "concurrent" else min(n, 8192)
        ccase = make_case(w, ncpu, 1234, "cpu")
        cstep = cpu_reference_step(w, ncpu, params, ccase, threads)
        cstep()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 40):
            cstep()
            reps += 1
        cs = (time.perf_counter() - t0) / reps
        cpu_baseline = {"value": ncpu * h / cs, "unit": "drone-steps/s", "cores": threads, "kind": "port",
                        "sample": f"N={ncpu} drones x h={h}, {reps} iterations ({cs * 1e3:.1f} ms each), "
                                  "oracle port of the reference's PyTorch op chain incl. SGD step"}

    # ---- end-to-end from RAW host samples (input side on the device, chunked H2D overlapped with the kernels).
    #      Runs last and guarded: whatever happens here, the line below still carries the measurements above.
    e2e_raw = None
    if not args.no_raw_e2e:
        try:
            e2e_raw = measure_e2e_raw(stepper, w, host, case, e2e_steps, world, dev, barrier)
        except Exception as ex:                                   # noqa: BLE001 - reported in the JSON line
            e2e_raw = {"ok": False, "error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        peak, peak_src, _ = measured_peaks()
        tq = stepper.runner.tcgen05
        mode = w.get("mode", "concurrent")
        sm_max = clocks.get("sm_max_mhz") or 1965.0
        _, _, pk = measured_peaks()
        fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
        if tq:
            fwd_name = "tq_fwd_kernel (tcgen05/TMEM policy chain) + tq_dyn_kernel (thread-per-drone dynamics, loss, reverse sweep)"
            adj_name = "tq_dx_kernel (tcgen05/TMEM dX chain) + tq_dw_kernel (tcgen05/TMEM streaming dW GEMM)"
            # dominant kernel = the longest launch of the step; algorithmic bytes of the adjoint pass (the
            # per-drone inputs read once = half of the fwd+bwd per-step figure) over ITS duration
            if kernel_ms is not None:
                dom = max(("tq_fwd_kernel", "tq_dyn_kernel", "tq_dx_kernel", "tq_dw_kernel"), key=lambda k: kernel_ms[k])
                dom_ms = kernel_ms[dom]
            else:
                dom, dom_ms = "tq_dx_kernel + tq_dw_kernel (adjoint pass)", ms_adj
            rows_f, rows_z = 632, 456                       # stash rows per drone (tq_layout.cuh F_ROWS / Z_ROWS), fp32
            model = {"tq_fwd_kernel": (15 + 90) * 4 + rows_f * 4, "tq_dyn_kernel": (12 + 90 + 40 + 40) * 4,
                     "tq_dx_kernel": (40 + 416 + 416) * 4, "tq_dw_kernel": (rows_f - 40 + rows_z) * 4}
        else:
            fwd_name = adj_name = "tile-engine kernels (mma.sync 3xTF32 / FFMA)"
            dom = {"concurrent": "hutter_adj_kernel", "autoregressive": "rec_adj_kernel", "lstm": "lstm_adj_kernel"}[
                mode] if w["system"] != "cartpole" else "simple_adj_kernel"
            dom_ms, model = ms_adj, {}
        alg_bytes = 0.5 * w["bytes_per_step"] * n * h
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        # DRAM bytes of ONE launch of the dominant kernel: measured now, by a short ncu capture of the same workload in
        # a child process (two metrics, one kernel launch, outside every timed region); if ncu is not there or fails,
        # the figure of the committed `--set full` capture (profiles/ncu_traffic.json), and the line says which
        traffic, traffic_src = None, None
        if world == 1 and not args.no_live_traffic:
            traffic = live_dram_traffic(args.workload, n, dom.split(" ")[0])
            traffic_src = "ncu child process of this run (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        if traffic is None:
            tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
                    traffic_src = "profiles/ncu_traffic.json (committed ncu --set full capture)"
                except Exception:
                    traffic = None
        flops = w["flops_per_step"] * n * h
        if tq:
            tf32_peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1650.0))) / 2.0
            rc = {"bound": "tensor (tcgen05 kind::tf32, 3xTF32 split = 3 tensor instructions per fp32-level product)",
                  "achieved": flops / (ms_fwd + ms_adj) / 1e9, "peak": tf32_peak / 3.0,
                  "unit": "TFLOP/s (fp32-equivalent algorithmic flops)",
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (tf32) / 3 (3xTF32); fp32 FFMA peak "
                                 "for reference: %.1f TFLOP/s" % fp32_peak}
        else:
            mma_tf32_peak = 148 * 476 * 2 * sm_max * 1e6 / 1e12
            rc = {"bound": "tensor (mma.sync tf32, 3xTF32 split = 3 instructions per product)",
                  "achieved": flops / (ms_fwd + ms_adj) / 1e9, "peak": mma_tf32_peak / 3.0,
                  "unit": "TFLOP/s (fp32-equivalent algorithmic flops)",
                  "peak_source": "measured mma.sync m16n8k8 tf32 rate 476 MAC/clk/SM x 148 SM x clocks.max.sm / 3 "
                                 "(tools/micro/rates.cu); fp32 FFMA peak for reference: %.1f TFLOP/s" % fp32_peak}
        rc["frac"] = rc["achieved"] / rc["peak"]
        own = stepper.kernel_launches_per_step
        fused_sgd = tq and world == 1 and stepper.peer is None
        line = {
            "metric": "drone-steps/sec (fwd+bwd)", "value": value, "unit": "drone-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "horizon": h, "n_drones_per_gpu": n, "n_drones_total": n * world,
                       "forward_kernel": fwd_name, "adjoint_kernel": adj_name,
                       "policy_init": "torch default init, seed 0", "optimizer": "SGD lr %g momentum 0.9" % LR[w["system"]]
                       + (" (fused into the gradient reduction kernel)" if fused_sgd else " (two torch element-wise launches)"),
                       "l2": "flushed between timed steps (256 MiB write, untimed)" if flush_buf is not None else "not flushed",
                       "parallelism": f"dp{world} (drone-axis shards, " + (
                           "gradient exchanged by apg_reduce_scatter_p2p_kernel / apg_gather_sgd_p2p_kernel over "
                           "NVLink peer memory)" if stepper.peer is not None else
                           ("one NCCL sum-allreduce of the flat gradient)" if world > 1 else "single device, no collective)"))},
            "ms_forward_kernel": ms_fwd, "ms_adjoint_kernel": ms_adj, "wall_s_timed_region": t_wall,
            "final_loss": final_loss,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "ms_kernel": dom_ms,
                         "note": "algorithmic bytes = the per-drone inputs read once by the adjoint pass (SURVEY 8d: "
                                 "828 B per drone for this workload) over the duration of the dominant launch; "
                                 "traffic = ncu dram bytes of one launch of that kernel (traffic_source); "
                                 "the kernels move more than the algorithmic bytes by design (operand-image stash, "
                                 "see roofline_stash); the path as a whole is compute / latency bound, see "
                                 "roofline_compute"},
            "roofline_compute": rc,
            "e2e": {"value": e2e_value, "unit": "drone-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "api": "apg_trajectory_tracking_b200.train.FusedTrainStep.step(host tensors)"},
            "gpu_launches": own * args.steps,
            "launches_per_step": {"own_kernels": own, "torch_elementwise": 0 if (fused_sgd or stepper.peer is not None) else 2,
                                  "nccl": 1 if (world > 1 and stepper.peer is None) else 0,
                                  "memset_nodes": 1 if tq else 0},
            "clocks": clocks,
        }
        if kernel_ms is not None:
            line["kernel_ms"] = kernel_ms
            # how close each kernel is to the HBM roof on the bytes it moves BY CONSTRUCTION (stash sets it reads /
            # writes + per-drone inputs; model, per drone x N) - the streaming dW GEMM is the HBM-bound one
            line["roofline_stash"] = {k: {"model_bytes": model[k] * n, "GB/s": model[k] * n / (kernel_ms[k] * 1e-3) / 1e9,
                                          "frac_of_hbm_peak": model[k] * n / (kernel_ms[k] * 1e-3) / 1e9 / peak}
                                      for k in model}
        if e2e_raw is not None:
            # the raw-sample path is the headline e2e when it ran, matched the prepared-input path and is faster;
            # the prepared-input measurement is kept next to it
            if e2e_raw.get("ok") and e2e_raw["value"] > e2e_value:
                line["e2e_prepared_inputs"] = line["e2e"]
                line["e2e"] = {k: e2e_raw[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step",
                                                        "steps", "api", "chunk", "chunk_candidates_ms", "check",
                                                        "value_loss_read_every_step",
                                                        "value_loss_read_one_step_late", "value_eager_step_host",
                                                        "value_graph_replay_host", "graph_chunk",
                                                        "graph_chunk_candidates_ms", "graph_error") if k in e2e_raw}
            else:
                line["e2e_raw_samples"] = e2e_raw
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
