"""Host-side driver of the fused rollout: specification, workspace, autograd bridge.

The training step of the reference (scripts/train_base.py:188-218 with the per-system loops it calls) maps to ONE
forward launch + ONE adjoint launch here; torch only provides device memory, streams and the optimizer."""
import ctypes
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _capi, params as P

NET_KIND = {"hutter_conv": 0, "hutter_lin": 1, "simple": 2, "lstm": 3}
MODE = {"concurrent": 0, "autoregressive": 1, "lstm": 2, "LSTM": 2}
WINDOW = {"cumulative": 0, "relative": 1}
REF_WIDTH = {"quad": 9, "wing": 3, "cartpole": 0}


@dataclass
class RolloutSpec:
    """Static description of one rollout problem (what apg_config carries)."""
    system: str
    horizon: int
    dt: float
    mode: str = "concurrent"
    net: str = "hutter_conv"
    state_feat: int = 15
    ref_len: int = 10
    ref_dim: int = 9
    out_dim: int = 40
    window: str = "cumulative"
    modified_params: dict = field(default_factory=dict)

    @staticmethod
    def quad_concurrent(horizon=10, dt=0.1, modified_params=None):
        return RolloutSpec("quad", horizon, dt, "concurrent", "hutter_conv", 15, horizon, 9, 4 * horizon,
                           modified_params=modified_params or {})

    @staticmethod
    def wing_concurrent(horizon=10, dt=0.05, modified_params=None):
        return RolloutSpec("wing", horizon, dt, "concurrent", "hutter_lin", 9, 1, 3, 4 * horizon,
                           modified_params=modified_params or {})

    @staticmethod
    def cartpole_concurrent(horizon=10, dt=0.05, modified_params=None):
        return RolloutSpec("cartpole", horizon, dt, "concurrent", "simple", 4, 0, 0, horizon,
                           modified_params=modified_params or {})

    @staticmethod
    def quad_recurrent(mode, horizon=10, dt=0.1, window="cumulative", modified_params=None):
        net = "lstm" if mode.lower() == "lstm" else "hutter_conv"
        return RolloutSpec("quad", horizon, dt, mode, net, 15, horizon, 9, 4, window=window,
                           modified_params=modified_params or {})

    def config(self, n_drones):
        c = _capi.ApgConfig()
        c.system = P.SYSTEM_ID[self.system]
        c.mode = MODE[self.mode]
        c.window = WINDOW[self.window]
        c.net = NET_KIND[self.net]
        c.n_drones = int(n_drones)
        c.horizon = int(self.horizon)
        c.state_feat, c.ref_len, c.ref_dim, c.out_dim = self.state_feat, self.ref_len, self.ref_dim, self.out_dim
        c.dt = float(self.dt)
        phys = P.PHYS[self.system](self.modified_params)
        for i in range(P.MAX_PHYS):
            c.phys[i] = float(phys[i])
        return c

    @property
    def state_dim(self):
        return P.STATE_DIM[self.system]

    @property
    def action_dim(self):
        return P.ACTION_DIM[self.system]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _dev_f32(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise _capi.ApgError(f"{name}: expected a CUDA tensor (the rollout has no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


class Rollout:
    """Owns the workspace for a (spec, N) pair and issues the forward / adjoint launches on the current stream."""

    def __init__(self, spec: RolloutSpec, n_drones: int, device=None):
        if not torch.cuda.is_available():
            raise _capi.ApgError("no CUDA device: the fused rollout only runs on the GPU")
        self.spec, self.n = spec, int(n_drones)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _capi.lib()
        with torch.cuda.device(self.device):
            self.cfg = spec.config(self.n)
            n_params = self.lib.apg_num_params(ctypes.byref(self.cfg))
            if n_params < 0:
                _capi.check(n_params)
            self.n_params = n_params
            # True: the tcgen05 / TMEM kernels serve this configuration (pack, chain, dynamics + loss sum | dX chain,
            # dW GEMM, reduce = 6 launches per iteration); False: the tile-engine kernels (5 launches)
            self.tcgen05 = self.lib.apg_rollout_kernel_path(ctypes.byref(self.cfg)) == 1
            ws = self.lib.apg_workspace_bytes(ctypes.byref(self.cfg))
            self.workspace = torch.empty(ws + 256, dtype=torch.uint8, device=self.device)
            off = (-self.workspace.data_ptr()) % 256
            self._ws_ptr = ctypes.c_void_p(self.workspace.data_ptr() + off)
            # the loss of forward number k lands in element k % LOSS_RING of this buffer, and forward() returns that
            # 1-element view: losses collected over up to LOSS_RING steps stay what they were (the reference returns a
            # fresh tensor per step) without an extra clone launch per step
            self._loss_ring = torch.zeros(self.LOSS_RING, dtype=torch.float32, device=self.device)
            self.loss = self._loss_ring[0:1]
            self.forward_count = 0

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    LOSS_RING = 1024

    def forward(self, params_flat, in_state, cur, in_ref=None, ref=None, h0c0=None, want_states=False,
                want_actions=False, learnt_params=None):
        """One fused rollout; returns (loss, states or None, actions or None).  ``loss`` is a 1-element view into a ring
        of ``LOSS_RING`` device floats: it keeps its value until ``LOSS_RING`` further forwards have run (clone it to
        keep it longer).  ``backward`` differentiates the LAST forward of this runner.

        ``learnt_params`` (flat vector of ``LearntDynamics``, 1891 floats): the h steps are taken by the learnt residual
        model instead of the analytic one (``apg_rollout_forward_learnt``; tcgen05 configurations only) - the
        controller-through-learnt-dynamics phase of the reference (train_drone.py:175-199, 260-278)."""
        s = self.spec
        self.loss = self._loss_ring[self.forward_count % self.LOSS_RING:self.forward_count % self.LOSS_RING + 1]
        self.forward_count += 1
        self._inputs = [_dev_f32(x, n) for x, n in ((params_flat, "params"), (in_state, "in_state"), (cur, "cur"),
                                                     (in_ref, "in_ref"), (ref, "ref"), (h0c0, "h0c0"))]
        states = torch.empty(self.n, s.horizon, s.state_dim, device=self.device) if want_states else None
        actions = torch.empty(self.n, s.horizon, s.action_dim, device=self.device) if want_actions else None
        with torch.cuda.device(self.device):
            if learnt_params is not None:
                lp = _dev_f32(learnt_params, "learnt_params")
                if lp.numel() != self.lib.apg_learnt_num_params(self.cfg.system):
                    raise _capi.ApgError("learnt_params: wrong length for this system's learnt dynamics")
                if h0c0 is not None:
                    raise _capi.ApgError("learnt dynamics inside the rollout: concurrent mode only")
                self._learnt = lp                  # keeps the vector alive until the launch has run
                i = self._inputs
                _capi.check(self.lib.apg_rollout_forward_learnt(ctypes.byref(self.cfg), _ptr(i[0]), _ptr(lp),
                                                                _ptr(i[1]), _ptr(i[2]), _ptr(i[3]), _ptr(i[4]),
                                                                self._ws_ptr, _ptr(self.loss), _ptr(states),
                                                                _ptr(actions), self._stream()))
            else:
                _capi.check(self.lib.apg_rollout_forward(ctypes.byref(self.cfg), *[_ptr(x) for x in self._inputs],
                                                         self._ws_ptr, _ptr(self.loss), _ptr(states), _ptr(actions),
                                                         self._stream()))
        return self.loss, states, actions

    def backward(self, grad_loss=1.0, out=None):
        """Adjoint of the last forward() (same inputs); returns the flat parameter gradient."""
        if out is None:
            out = torch.empty(self.n_params, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.apg_rollout_backward(ctypes.byref(self.cfg), *[_ptr(x) for x in self._inputs],
                                                      self._ws_ptr, ctypes.c_float(float(grad_loss)), _ptr(out),
                                                      self._stream()))
        return out

    def backward_sgd(self, params_flat, momentum_buf, lr, momentum, grad_loss=1.0, out=None):
        """Adjoint of the last forward() with the SGD(momentum) update fused into the gradient reduction (tcgen05 path
        only): ``params_flat`` (the vector the forward read) and ``momentum_buf`` are updated in place."""
        ins = list(self._inputs)
        ins[0] = _dev_f32(params_flat, "params")
        with torch.cuda.device(self.device):
            _capi.check(self.lib.apg_rollout_backward_sgd(ctypes.byref(self.cfg), *[_ptr(x) for x in ins],
                                                          self._ws_ptr, ctypes.c_float(float(grad_loss)), _ptr(out),
                                                          _ptr(momentum_buf), ctypes.c_float(float(lr)),
                                                          ctypes.c_float(float(momentum)), self._stream()))
        return out

    def backward_p2p(self, comm, grad_loss=1.0):
        """Adjoint of the last forward() whose final gradient reduction stores its result into every rank's receive
        slot over peer memory (``dist.PeerGradExchange.next_step()`` gives ``comm``); complete the step with
        ``PeerGradExchange.gather``."""
        with torch.cuda.device(self.device):
            _capi.check(self.lib.apg_rollout_backward_p2p(ctypes.byref(self.cfg), *[_ptr(x) for x in self._inputs],
                                                          self._ws_ptr, ctypes.c_float(float(grad_loss)),
                                                          ctypes.byref(comm), self._stream()))

    def value_and_grad(self, params_flat, in_state, cur, in_ref=None, ref=None, h0c0=None, out=None,
                       learnt_params=None):
        loss, _, _ = self.forward(params_flat, in_state, cur, in_ref, ref, h0c0, learnt_params=learnt_params)
        return loss, self.backward(1.0, out=out)


def value_and_grad_host(spec: RolloutSpec, params_flat, in_state, cur, in_ref=None, ref=None, h0c0=None):
    """The same train-step evaluation through the HOST-buffer C entry point (numpy / CPU tensors in, numpy out)."""
    lib = _capi.lib()

    def host(x):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return np.ascontiguousarray(x, dtype=np.float32)
    arrs = [host(x) for x in (params_flat, in_state, cur, in_ref, ref, h0c0)]
    n = arrs[2].shape[0]
    cfg = spec.config(n)
    n_params = lib.apg_num_params(ctypes.byref(cfg))
    if n_params < 0:
        _capi.check(n_params)
    loss = np.zeros(1, dtype=np.float32)
    grad = np.zeros(n_params, dtype=np.float32)

    def hp(a):
        return None if a is None else ctypes.c_void_p(a.ctypes.data)
    _capi.check(lib.apg_rollout_value_and_grad_host(ctypes.byref(cfg), *[hp(a) for a in arrs], hp(loss), hp(grad)))
    return float(loss[0]), grad


class _FusedRolloutFn(torch.autograd.Function):
    """loss = rollout(params); backward multiplies the analytic gradient by the incoming scalar."""

    @staticmethod
    def forward(ctx, params_flat, runner, in_state, cur, in_ref, ref, h0c0):
        loss, _, _ = runner.forward(params_flat, in_state, cur, in_ref, ref, h0c0)
        ctx.runner = runner
        ctx.forward_count = runner.forward_count
        return loss.clone().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.runner.forward_count != ctx.forward_count:
            raise RuntimeError("fused_rollout_loss: the runner has run another forward since this loss was computed; its "
                               "adjoint would differentiate that one (use one Rollout per live loss, or call backward "
                               "before the next forward)")
        g = ctx.runner.backward(1.0)
        return g * grad_out, None, None, None, None, None, None


def fused_rollout_loss(runner: Rollout, params_flat, in_state, cur, in_ref=None, ref=None, h0c0=None):
    """Differentiable scalar loss of the whole rollout w.r.t. the flat parameter vector."""
    return _FusedRolloutFn.apply(params_flat, runner, in_state, cur, in_ref, ref, h0c0)


def flatten_params(params):
    return torch.cat([p.detach().reshape(-1) for p in params])


def split_flat(flat, like):
    out, o = [], 0
    for p in like:
        n = p.numel()
        out.append(flat[o:o + n].view_as(p))
        o += n
    return out
