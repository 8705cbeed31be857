"""Un-fused CUDA ops with autograd support: one dynamics step and the quadrotor featurizer.

They back the per-call API of the reference (``Dynamics.__call__``, ``state_preprocessing``) for callers that do not
use the fused rollout; forward and backward each run one hand-written kernel (csrc/misc_kernels.cu).  CUDA tensors
only -- there is no CPU implementation in this package."""
import ctypes

import numpy as np
import torch

from . import _capi, params as P


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _capi.ApgError("apg_trajectory_tracking_b200 computes on the GPU only: move the tensors to a CUDA "
                                 "device (there is no CPU fallback; use the reference package for CPU-only runs)")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _DynamicsStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, action, dt, system, phys):
        _require_cuda(state, action)
        lib = _capi.lib()
        s = state.detach().contiguous().float()
        a = action.detach().contiguous().float()
        out = torch.empty_like(s)
        with torch.cuda.device(s.device):
            _capi.check(lib.apg_dynamics_step(system, ctypes.c_void_p(phys.ctypes.data), _p(s), _p(a),
                                              ctypes.c_float(dt), s.shape[0], _p(out), _stream(s)))
        ctx.save_for_backward(s, a)
        ctx.dt, ctx.system, ctx.phys = dt, system, phys
        return out

    @staticmethod
    def backward(ctx, grad_out):
        s, a = ctx.saved_tensors
        g = grad_out.contiguous().float()
        gs, ga = torch.empty_like(s), torch.empty_like(a)
        with torch.cuda.device(s.device):
            _capi.check(_capi.lib().apg_dynamics_step_adjoint(ctx.system, ctypes.c_void_p(ctx.phys.ctypes.data), _p(s),
                                                              _p(a), ctypes.c_float(ctx.dt), s.shape[0], _p(g), _p(gs),
                                                              _p(ga), _stream(s)))
        return gs, ga, None, None, None


def dynamics_step(system, phys, state, action, dt):
    """next_state = f(state, action) for a batch; differentiable w.r.t. state and action."""
    phys = np.ascontiguousarray(phys, dtype=np.float32)
    return _DynamicsStep.apply(state, action, float(dt), P.SYSTEM_ID[system], phys)


class _QuadFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state):
        _require_cuda(state)
        s = state.detach().contiguous().float()
        out = torch.empty(s.shape[0], 15, device=s.device, dtype=torch.float32)
        with torch.cuda.device(s.device):
            _capi.check(_capi.lib().apg_quad_features(_p(s), s.shape[0], _p(out), _stream(s)))
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, grad_feat):
        (s,) = ctx.saved_tensors
        g = grad_feat.contiguous().float()
        gs = torch.empty_like(s)
        with torch.cuda.device(s.device):
            _capi.check(_capi.lib().apg_quad_features_adjoint(_p(s), _p(g), s.shape[0], _p(gs), _stream(s)))
        return gs


def quad_features(state):
    return _QuadFeatures.apply(state)
