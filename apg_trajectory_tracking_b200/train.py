"""Train step of the reference's trainers on the fused kernels.

``FusedTrainStep`` is the body of one mini-batch iteration of ``TrainBase.run_epoch`` (scripts/train_base.py:188-218)
together with the per-system loop it calls (train_drone.py:113-203, train_fixed_wing.py:90-116,
train_cartpole.py:118-155) and ``optim.SGD(lr, momentum=0.9).step()`` (train_base.py:139-143):

    zero_grad -> policy forward -> sigmoid -> h x dynamics -> *_mpc_loss -> backward -> [allreduce] -> SGD step

as two kernel launches (+ pack / reduce helpers) on a flat parameter vector.  Batches shard over GPUs along the
drone axis; the only collective is one sum-allreduce of the flat gradient (the loss is a sum over drones,
drone_loss.py:22-33, so the summed gradient equals the single-device large-batch gradient).
"""
import torch

from . import rollout as R


class FusedTrainStep:
    def __init__(self, params, spec: R.RolloutSpec, n_drones: int, lr: float, momentum: float = 0.9, device=None,
                 process_group=None, distributed=None):
        """params: iterable of tensors in net.parameters() order (or an nn.Module)."""
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
        self.kernel_launches_per_step = 5      # pack, forward, loss-sum, adjoint, grad-reduce

    def _dev(self, x):
        if x is None:
            return None
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)      # host inputs: H2D copy is part of the step
        return x

    def value_and_grad(self, in_state, cur, in_ref=None, ref=None, h0c0=None):
        loss, _ = self.runner.value_and_grad(self.flat, self._dev(in_state), self._dev(cur), self._dev(in_ref),
                                             self._dev(ref), self._dev(h0c0), out=self.grad)
        if self.distributed:
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM, group=self.pg)
        return loss, self.grad

    def step(self, in_state, cur, in_ref=None, ref=None, h0c0=None):
        """one full train iteration; returns the (local-shard) loss as a 1-element device tensor"""
        loss, grad = self.value_and_grad(in_state, cur, in_ref, ref, h0c0)
        # optim.SGD(momentum=0.9): buf = momentum*buf + g ; p -= lr*buf   (first step: buf = g)
        self.buf.mul_(self.momentum).add_(grad)
        self.flat.add_(self.buf, alpha=-self.lr)
        return loss

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

    def loss_and_grad(self, in_state, cur, in_ref=None, ref=None, h0c0=None):
        cur = self._dev(cur)
        runner = self.runner(cur.shape[0])
        loss, _ = runner.value_and_grad(self.flat, self._dev(in_state), cur, self._dev(in_ref), self._dev(ref),
                                        self._dev(h0c0), out=self.flat_grad)
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(self.pg) > 1:
            torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        for used, p, gview in self._views:
            p.grad = gview if used else None
        return loss.reshape(()).clone()
