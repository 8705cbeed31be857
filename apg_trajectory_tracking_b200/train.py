"""Train step of the reference's trainers on the fused kernels.

``FusedTrainStep`` is the body of one mini-batch iteration of ``TrainBase.run_epoch`` (scripts/train_base.py:188-218)
together with the per-system loop it calls (train_drone.py:113-203, train_fixed_wing.py:90-116,
train_cartpole.py:118-155) and ``optim.SGD(lr, momentum=0.9).step()`` (train_base.py:139-143):

    zero_grad -> policy forward -> sigmoid -> h x dynamics -> *_mpc_loss -> backward -> [allreduce] -> SGD step

as two kernel launches (+ pack / reduce helpers) on a flat parameter vector.  Batches shard over GPUs along the
drone axis; the only collective is one sum-allreduce of the flat gradient (the loss is a sum over drones,
drone_loss.py:22-33, so the summed gradient equals the single-device large-batch gradient).
"""
import os

import torch

from . import _capi, prepare as PR, rollout as R, synthetic as _syn


def chunk_bounds(n, chunk, align=64):
    """[(a, b)] contiguous drone ranges of at most `chunk` rows covering [0, n); every start is a multiple of `align`
    (the 64-drone tile, which also keeps every per-chunk pointer 16-byte aligned)."""
    n, chunk = int(n), int(chunk)
    if n <= 0:
        return []
    if chunk <= 0 or chunk >= n:
        return [(0, n)]
    chunk = max(align, (chunk // align) * align)
    return [(a, min(a + chunk, n)) for a in range(0, n, chunk)]


class HostLoss:
    """loss of one ``step_host_async`` call on its way to the host"""

    def __init__(self, pinned, event):
        self._pinned, self._event = pinned, event

    def item(self):
        self._event.synchronize()
        return float(self._pinned[0])


class FusedTrainStep:
    def __init__(self, params, spec: R.RolloutSpec, n_drones: int, lr: float, momentum: float = 0.9, device=None,
                 process_group=None, distributed=None, peer_exchange=None):
        """params: iterable of tensors in net.parameters() order (or an nn.Module).  ``peer_exchange`` (default: the
        environment variable APG_P2P_GRAD == "1"): with more than one rank, exchange the gradient with this
        package's own kernels over NVLink peer memory (``dist.PeerGradExchange``) instead of an NCCL all-reduce."""
        if isinstance(params, torch.nn.Module):
            params = list(params.parameters())
        self.like = [p.detach() for p in params]
        self.runner = R.Rollout(spec, n_drones, device)
        self.device = self.runner.device
        self.flat = R.flatten_params(self.like).to(self.device).float().contiguous()
        if self.flat.numel() != self.runner.n_params:
            raise ValueError(f"parameter count {self.flat.numel()} does not match the policy described by the spec "
                             f"({self.runner.n_params})")
        self.grad = torch.zeros_like(self.flat)
        self.buf = torch.zeros_like(self.flat)
        self.lr, self.momentum = float(lr), float(momentum)
        self.pg = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(process_group) > 1
        self.distributed = distributed
        # own kernels per iteration: pack, forward, loss-sum, adjoint, grad-reduce (tile-engine kernels) or pack,
        # forward chain, dynamics (+ loss sum), dX chain, dW GEMM, grad-reduce (+ SGD on one device) on the tcgen05 path;
        # elsewhere the SGD update adds two torch element-wise launches, N > 1 one NCCL all-reduce kernel
        self.kernel_launches_per_step = 6 if getattr(self.runner, "tcgen05", False) else 5
        # Gradient exchange with more than one rank.  Measured on 2 / 4 / 8 B200 (profiles/r2_multi_gpu.md): the package's
        # own peer-memory kernels beat NCCL all-reduce + two optimizer launches at every size (8 GPUs: 0.324 vs 0.336 ms
        # per iteration) and keep the replicas bitwise identical, so they are the DEFAULT for the configuration they
        # were measured on (the tcgen05 path, NCCL backend); APG_P2P_GRAD=0 / =1 forces NCCL / the peer kernels
        # everywhere.  If the peer mappings cannot be set up on some rank, all ranks fall back to NCCL together.
        if peer_exchange is None:
            env = os.environ.get("APG_P2P_GRAD")
            if env is not None:
                peer_exchange = env == "1"
            else:
                peer_exchange = bool(getattr(self.runner, "tcgen05", False)) and self.distributed and \
                    torch.distributed.get_backend(process_group) == "nccl"
        self.peer = None
        if peer_exchange and self.distributed:
            from . import dist as D
            try:
                self.peer = D.PeerGradExchange(self.runner.n_params, self.device, process_group)
                ok = 1.0
            except Exception as ex:                                   # noqa: BLE001
                if os.environ.get("APG_P2P_GRAD") == "1":
                    raise
                self.peer, ok, self.peer_error = None, 0.0, f"{type(ex).__name__}: {ex}"[:200]
            flag = torch.tensor([ok], device=self.device)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=process_group)
            if float(flag.item()) < 1.0:
                self.peer = None
            if self.peer is not None:
                self.kernel_launches_per_step += 1  # + the gather / SGD kernel (the exchange rides on the grad-reduce)

    def _dev(self, x):
        if x is None:
            return None
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)      # host inputs: H2D copy is part of the step
        return x

    def value_and_grad(self, in_state, cur, in_ref=None, ref=None, h0c0=None):
        if self.peer is not None:
            loss = self._forward_and_scatter(in_state, cur, in_ref, ref, h0c0)
            self.peer.gather(*self._peer_step, grad_out=self.grad)
            return loss, self.grad
        loss, _ = self.runner.value_and_grad(self.flat, self._dev(in_state), self._dev(cur), self._dev(in_ref),
                                             self._dev(ref), self._dev(h0c0), out=self.grad)
        if self.distributed:
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        return loss, self.grad

    def _forward_and_scatter(self, in_state, cur, in_ref, ref, h0c0):
        loss, _, _ = self.runner.forward(self.flat, self._dev(in_state), self._dev(cur), self._dev(in_ref),
                                         self._dev(ref), self._dev(h0c0))
        self._peer_step = self.peer.next_step()
        self.runner.backward_p2p(self._peer_step[0])
        return loss

    def step(self, in_state, cur, in_ref=None, ref=None, h0c0=None):
        """one full train iteration; returns the (local-shard) loss as a 1-element device tensor"""
        if self.peer is not None:
            # adjoint + reduction + peer scatter, then ONE kernel: wait, rank-ordered sum, SGD(momentum) update
            loss = self._forward_and_scatter(in_state, cur, in_ref, ref, h0c0)
            self.peer.gather(*self._peer_step, grad_out=self.grad, params=self.flat, momentum_buf=self.buf,
                             lr=self.lr, momentum=self.momentum)
            return loss
        if getattr(self.runner, "tcgen05", False) and not self.distributed:
            # single device, tcgen05 path: the optimizer step rides on the gradient reduction (one launch)
            loss, _, _ = self.runner.forward(self.flat, self._dev(in_state), self._dev(cur), self._dev(in_ref),
                                             self._dev(ref), self._dev(h0c0))
            self.runner.backward_sgd(self.flat, self.buf, self.lr, self.momentum, out=self.grad)
            return loss
        loss, grad = self.value_and_grad(in_state, cur, in_ref, ref, h0c0)
        # optim.SGD(momentum=0.9): buf = momentum*buf + g ; p -= lr*buf   (first step: buf = g)
        self.buf.mul_(self.momentum).add_(grad)
        self.flat.add_(self.buf, alpha=-self.lr)
        return loss

    # -----------------------------------------------------------------------------------------------------------
    # CUDA graph of one train iteration (launch-bound small batches: the reference's own cartpole case is 128 drones)
    # -----------------------------------------------------------------------------------------------------------
    def capture(self, in_state, cur, in_ref=None, ref=None, h0c0=None, warmup=3):
        """Capture ``step`` on the given (static, device) tensors into a CUDA graph: pack, forward, loss sum, adjoint,
        gradient reduction and the SGD update become ONE launch.  ``replay()`` reruns it on whatever the static tensors
        hold then (copy a new batch into them first); returns the static 1-element loss tensor."""
        if self.distributed:
            raise _capi.ApgError("capture() is single-process (the gradient all-reduce is not captured)")
        args = tuple(None if x is None else self._dev(x) for x in (in_state, cur, in_ref, ref, h0c0))
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # kernels get their attributes set / modules loaded here
                self.step(*args)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_loss = self.step(*args)
        self._graph_args = args
        return self._graph_loss

    def replay(self):
        self._graph.replay()
        return self._graph_loss

    # -----------------------------------------------------------------------------------------------------------
    # host batches: raw samples in (pinned) host memory -> chunked H2D on a copy stream, overlapped with the
    # kernels of the previous chunk; the policy inputs are derived on the device (prepare.py)
    # -----------------------------------------------------------------------------------------------------------
    def default_chunk(self):
        """drones per chunk of step_host: two waves of 64-drone tiles over the SMs (each persistent CTA gets two
        tiles, so its GEMM / dynamics warps overlap)"""
        return 2 * 64 * max(1, _capi.lib().apg_sm_count())

    def _staging(self, n):
        """device staging of one step's raw samples and derived policy inputs"""
        spec, dev = self.runner.spec, self.device
        rows = spec.horizon if spec.mode == "concurrent" else 2 * spec.horizon
        sg = {"cur": torch.empty(n, spec.state_dim, device=dev), "done": None}
        if spec.system == "quad":
            sg["ref"] = torch.empty(n, rows, 9, device=dev)
            sg["in_ref"] = torch.empty(n, rows, 9, device=dev)
            sg["in_state"] = torch.empty(n, 15, device=dev) if spec.mode == "concurrent" else None
            sg["h0c0"] = torch.empty(2, n, 8, device=dev) if spec.mode.lower() == "lstm" else None
        elif spec.system == "wing":
            sg["target"] = torch.empty(n, 3, device=dev)
            sg["ref"] = torch.empty(n, spec.horizon, 3, device=dev)
            sg["in_ref"] = torch.empty(n, 3, device=dev)
            sg["in_state"] = torch.empty(n, 9, device=dev)
        return sg

    def _host_state(self, n):
        st = getattr(self, "_hs", None)
        if st is not None and st["n"] == n:
            return st
        dev = self.device
        # two staging sets used alternately: the copies of step i+1 may start while the kernels of step i still read
        # theirs (step_host_async); the loss of a step is copied into one of two pinned scalars
        st = {"n": n, "copy_stream": torch.cuda.Stream(device=dev), "runners": {}, "sets": [self._staging(n),
              self._staging(n)], "turn": 0, "grad_tmp": torch.zeros_like(self.grad),
              "loss_ring": torch.zeros(1024, device=dev), "calls": 0,
              "loss_host": [torch.empty(1, pin_memory=torch.cuda.is_available()) for _ in range(2)]}
        self._hs = st
        return st

    def _chunk_runner(self, st, size):
        if size == self.runner.n:
            return self.runner
        if size not in st["runners"]:
            st["runners"][size] = R.Rollout(self.runner.spec, size, self.device)
        return st["runners"][size]

    def step_host(self, cur, ref=None, h0c0=None, target=None, norm=None, chunk=None):
        """One full train iteration from RAW host samples: ``value_and_grad_host`` + the SGD(momentum) update.
        Returns the (local-shard) loss as a 1-element device tensor."""
        loss, grad = self.value_and_grad_host(cur, ref, h0c0, target, norm, chunk)
        self.buf.mul_(self.momentum).add_(grad)
        self.flat.add_(self.buf, alpha=-self.lr)
        return loss

    def step_host_async(self, cur, ref=None, h0c0=None, target=None, norm=None, chunk=None):
        """``step_host`` without the host synchronisation: returns a ``HostLoss`` whose ``item()`` waits for the
        device->host copy of this step's loss.  Calling it for step i+1 BEFORE reading the loss of step i lets the
        H2D copies of step i+1 overlap the last kernels of step i (two staging sets).  Read a handle before the step
        after next is issued (two pinned scalars are used alternately)."""
        st = self._host_state(int(cur.shape[0]))
        turn = st["turn"]
        loss = self.step_host(cur, ref, h0c0, target, norm, chunk)
        host = st["loss_host"][turn]
        host.copy_(loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return HostLoss(host, ev)

    def capture_host(self, cur, ref=None, h0c0=None, target=None, norm=None, chunk=None, warmup=2):
        """Capture ``step_host`` on the given PINNED host tensors into ONE CUDA graph: the chunked H2D copies (a parallel
        branch of the graph that runs ahead of the kernels), prepare / forward / adjoint of every chunk, the SGD update
        and the D2H copy of the loss into a pinned scalar.  Eagerly, every chunk costs ~15 CUDA calls issued from
        Python; beyond two chunks that CPU time, not PCIe or the kernels, sets the step time - captured, a step is one
        launch and the batch can be cut finely enough that little is left to compute once the last copy has landed.
        ``replay_host()`` reruns it on whatever the host tensors hold THEN and returns a ``HostLoss``."""
        if self.distributed:
            raise _capi.ApgError("capture_host() is single-process (the gradient all-reduce is not captured)")
        host = [x for x in (cur, ref, h0c0, target) if x is not None]
        if not all(x.is_pinned() for x in host):
            raise _capi.ApgError("capture_host(): the host tensors must be pinned (asynchronous copies inside a graph)")
        kw = dict(cur=cur, ref=ref, h0c0=h0c0, target=target, norm=norm, chunk=chunk)
        self._graph_state(int(cur.shape[0]))
        main = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # per-chunk runners / workspaces are created here, not in capture
                self._host_step_body(kw, graph=True)
        main.wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._host_step_body(kw, graph=True)
        self._hgraph, self._hgraph_keep = g, kw
        return self

    def _graph_state(self, n):
        st = self._host_state(n)
        if "graph_set" not in st:
            st["graph_set"] = self._staging(n)
            st["graph_loss_host"] = torch.empty(1, pin_memory=torch.cuda.is_available())
        return st

    def _host_step_body(self, kw, graph):
        st = self._hs
        loss, grad = self.value_and_grad_host(kw["cur"], kw["ref"], kw["h0c0"], kw["target"], kw["norm"], kw["chunk"],
                                              _graph=graph)
        self.buf.mul_(self.momentum).add_(grad)
        self.flat.add_(self.buf, alpha=-self.lr)
        st["graph_loss_host"].copy_(loss, non_blocking=True)

    def replay_host(self):
        """one captured ``step_host`` (see ``capture_host``); returns a ``HostLoss`` (``item()`` waits for the step).
        Every replay writes its loss into the same pinned scalar: read a handle before the next replay."""
        self._hgraph.replay()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return HostLoss(self._hs["graph_loss_host"], ev)

    def value_and_grad_host(self, cur, ref=None, h0c0=None, target=None, norm=None, chunk=None, allreduce=True,
                            _graph=False):
        """Loss and flat gradient from RAW host samples (pinned memory for asynchronous copies).

        quad: ``cur`` (N,12), ``ref`` (N,L,9) raw reference rows (L = h concurrent, 2h recurrent) [+ ``h0c0`` (2,N,8)];
        wing: ``cur`` (N,12), ``target`` (N,3), ``norm`` = (mean, std) of WingDataset (default: its fixed statistics);
        cartpole: ``cur`` (N,4).  The batch is cut into chunks of ``chunk`` drones (default ``default_chunk()``); the
        H2D copy of chunk c+1 runs on a copy stream while chunk c goes through prepare -> forward -> adjoint on the
        current stream; loss and gradient are summed over the chunks (the loss is a sum over drones), then the
        gradient is all-reduced as in ``value_and_grad``.  Returns (loss 1-element device tensor, flat gradient)."""
        spec = self.runner.spec
        n = int(cur.shape[0])
        st = self._host_state(n)
        compute = torch.cuda.current_stream(self.device)
        copy = st["copy_stream"]
        if _graph:
            # inside (or warming up for) a stream capture: a staging set of its own, the copy stream forks off the
            # capturing stream here and joins it again below; consecutive replays are ordered by the launching stream
            sg = st["graph_set"]
            copy.wait_stream(compute)
        else:
            sg = st["sets"][st["turn"]]
            st["turn"] ^= 1
            if sg["done"] is None:
                # first use: freshly allocated staging memory may be a recycled block with compute-stream work in flight
                copy.wait_stream(compute)
            else:
                copy.wait_event(sg["done"])          # the step that last used this set (two calls ago) has finished
        bounds = chunk_bounds(n, chunk if chunk is not None else self.default_chunk())
        if spec.system == "wing":
            mean, std = norm if norm is not None else (_syn.WING_MEAN, _syn.WING_STD)
        first = True
        # the summed loss of this call: one element of a ring (like Rollout.forward), so that losses returned by
        # earlier calls keep their values
        k = st["calls"] % st["loss_ring"].numel()
        st["calls"] += 1
        st["loss"] = st["loss_ring"][k:k + 1]
        for a, b in bounds:
            with torch.cuda.stream(copy):
                sg["cur"][a:b].copy_(cur[a:b], non_blocking=True)
                if spec.system == "quad":
                    sg["ref"][a:b].copy_(ref[a:b], non_blocking=True)
                    if sg["h0c0"] is not None:
                        sg["h0c0"][:, a:b].copy_(h0c0[:, a:b], non_blocking=True)
                elif spec.system == "wing":
                    sg["target"][a:b].copy_(target[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            compute.wait_event(ev)
            c_cur = sg["cur"][a:b]
            hc = None
            if spec.system == "quad" and spec.mode == "concurrent" and getattr(self.runner, "tcgen05", False):
                # tcgen05 path: the kernels take the RAW samples (prepare_data runs in their prologue)
                c_ins, c_inr, c_ref = None, None, sg["ref"][a:b]
            elif spec.system == "quad":
                want = ("in_state", "cur", "in_ref", "ref") if sg["in_state"] is not None else ("cur", "in_ref", "ref")
                outs = {"cur": c_cur, "ref": sg["ref"][a:b], "in_ref": sg["in_ref"][a:b]}
                if sg["in_state"] is not None:
                    outs["in_state"] = sg["in_state"][a:b]
                PR.prepare_quad(c_cur, sg["ref"][a:b], want=want, out=outs)
                c_ins, c_inr, c_ref = outs.get("in_state"), outs["in_ref"], outs["ref"]
                if sg["h0c0"] is not None:
                    hc = sg["h0c0"][:, a:b].contiguous() if len(bounds) > 1 else sg["h0c0"]
            elif spec.system == "wing":
                outs = {"cur": c_cur, "ref": sg["ref"][a:b], "in_ref": sg["in_ref"][a:b],
                        "in_state": sg["in_state"][a:b]}
                PR.prepare_wing(c_cur, sg["target"][a:b], mean, std, spec.dt, spec.horizon, out=outs)
                c_ins, c_inr, c_ref = outs["in_state"], outs["in_ref"], outs["ref"]
            else:
                c_ins, c_inr, c_ref = c_cur, None, None
            runner = self._chunk_runner(st, b - a)
            loss, _, _ = runner.forward(self.flat, c_ins, c_cur, c_inr, c_ref, hc)
            if first:
                runner.backward(1.0, out=self.grad)
                st["loss"].copy_(loss)
                first = False
            else:
                runner.backward(1.0, out=st["grad_tmp"])
                self.grad.add_(st["grad_tmp"])
                st["loss"].add_(loss)
        if self.distributed and allreduce:
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        if _graph:
            compute.wait_stream(copy)            # join (every copy is already ordered before its chunk's kernels)
        else:
            sg["done"] = torch.cuda.Event()
            sg["done"].record(compute)
        # launches of this package's kernels: per chunk the rollout's 5 + the prepare kernels (quad / wing: 2)
        raw_mode = spec.system == "quad" and spec.mode == "concurrent" and getattr(self.runner, "tcgen05", False)
        self.host_launches_per_step = len(bounds) * (self.kernel_launches_per_step +
                                                     (0 if (spec.system == "cartpole" or raw_mode) else 2))
        return st["loss"], self.grad

    def parameters(self):
        """current parameters as views shaped like the originals"""
        return R.split_flat(self.flat, self.like)

    def gradients(self):
        return R.split_flat(self.grad, self.like)


# ---------------------------------------------------------------------------------------------------------------
# nn.Module binding: the trainers keep a regular module + torch.optim.SGD (reference plumbing, train_base.py:130-143)
# ---------------------------------------------------------------------------------------------------------------
def spec_for_net(net, system, horizon, dt, train_mode="concurrent", window="cumulative", modified_params=None):
    """RolloutSpec of a policy module of this package for the given system / train mode."""
    mp = modified_params or {}
    mode = {"concurrent": "concurrent", "autoregressive": "autoregressive", "LSTM": "lstm", "lstm": "lstm"}[train_mode]
    if system == "cartpole":
        return R.RolloutSpec.cartpole_concurrent(horizon, dt, mp)
    if system == "wing":
        return R.RolloutSpec.wing_concurrent(horizon, dt, mp)
    if mode == "concurrent":
        return R.RolloutSpec.quad_concurrent(horizon, dt, mp)
    return R.RolloutSpec.quad_recurrent(mode, horizon, dt, window, mp)


class ModuleRollout:
    """Binds a policy nn.Module to the fused rollout: the module's parameters become views into one flat CUDA
    buffer (so optimizers keep working on the module), gradients are written into views of one flat buffer and
    ``p.grad`` of tensors the forward does not use stays ``None`` exactly like under the reference's autograd."""

    def __init__(self, net, spec: R.RolloutSpec, device=None, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ModuleRollout needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.net, self.spec, self.pg = net, spec, process_group
        net.to(self.device)
        plist = list(net.named_parameters())
        self.flat = torch.cat([p.detach().reshape(-1).float() for _, p in plist]).contiguous()
        self.flat_grad = torch.zeros_like(self.flat)
        o = 0
        self._views = []
        used = set(net.used_parameter_names()) if hasattr(net, "used_parameter_names") else {n for n, _ in plist}
        for name, p in plist:
            n = p.numel()
            p.data = self.flat[o:o + n].view_as(p)
            self._views.append((name in used, p, self.flat_grad[o:o + n].view_as(p)))
            o += n
        self._runners = {}

    def runner(self, n):
        if n not in self._runners:
            self._runners[n] = R.Rollout(self.spec, n, self.device)
            if self._runners[n].n_params != self.flat.numel():
                raise ValueError("module does not match the rollout spec")
        return self._runners[n]

    def _dev(self, x):
        return None if x is None else x.to(self.device, non_blocking=True)

    def supports_learnt_dynamics(self, n):
        """True if the rollout of n drones can take its steps through a learnt dynamics model (tcgen05 path)"""
        return self.spec.system == "quad" and self.spec.mode == "concurrent" and getattr(self.runner(n), "tcgen05", False)

    def loss_and_grad(self, in_state, cur, in_ref=None, ref=None, h0c0=None, learnt_params=None):
        """``in_ref=None`` for a quadrotor batch: the policy inputs are derived from the raw ``(cur, ref)`` samples on
        the device (QuadDataset.prepare_data as a kernel, prepare.py) -- only those two tensors cross PCIe.
        ``learnt_params``: flat parameters of a ``LearntDynamics`` the horizon is rolled through (see Rollout.forward)."""
        cur = self._dev(cur)
        runner = self.runner(cur.shape[0])
        in_state, in_ref, ref = self._dev(in_state), self._dev(in_ref), self._dev(ref)
        if self.spec.system == "quad" and in_ref is None and ref is not None and not (
                self.spec.mode == "concurrent" and getattr(runner, "tcgen05", False)):
            # (the tcgen05 kernels take the raw samples as they are: prepare_data runs in their prologue)
            want = ("in_state", "cur", "in_ref", "ref") if self.spec.mode == "concurrent" else ("cur", "in_ref", "ref")
            d = PR.prepare_quad(cur, ref, want=want)
            in_state, cur, in_ref, ref = d.get("in_state"), d["cur"], d["in_ref"], d["ref"]
        loss, _ = runner.value_and_grad(self.flat, in_state, cur, in_ref, ref, self._dev(h0c0), out=self.flat_grad,
                                        learnt_params=learnt_params)
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        for used, p, gview in self._views:
            p.grad = gview if used else None
        return loss.reshape(()).clone()
