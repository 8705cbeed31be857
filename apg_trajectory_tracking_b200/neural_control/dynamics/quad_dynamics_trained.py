"""Learnt residual quadrotor dynamics (reference: ``LearntDynamics`` in
``neural_control/dynamics/quad_dynamics_trained.py:10-69``).

Same parameters (names, shapes, initial values: identity action transform, zero residual MLP, mass / inertia / kinv
vectors) and the same forward, ``simulate_quadrotor(linear_at @ action, state, dt) + state_transformer(state,
linear_at @ action)``, as ONE CUDA kernel; backward is one hand-written adjoint kernel that returns the cotangents of
state, action and of every parameter (csrc/learnt_math.cuh, csrc/learnt_kernels.cu).  Like the reference, the
simulator keeps the construction-time kinv / inertia values (its derived matrices are built once in ``__init__``,
:47-48) while the parameter vectors still receive their gradients."""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from ... import _capi
from ...ops import _p, _require_cuda, _stream
from .quad_dynamics_flightmare import FlightmareDynamics


class _LearntStep(torch.autograd.Function):
    """one learnt-dynamics step (system 0 = quadrotor, 1 = fixed wing): forward kernel + hand-written adjoint kernel
    that returns the cotangents of the flat parameter vector, the state and the action"""

    @staticmethod
    def forward(ctx, flat, state, action, dt, phys, system=0):
        _require_cuda(flat, state, action)
        lib = _capi.lib()
        s = state.detach().contiguous().float()
        a = action.detach().contiguous().float()
        p = flat.detach().contiguous().float()
        out = torch.empty_like(s)
        with torch.cuda.device(s.device):
            _capi.check(lib.apg_learnt_step(system, _p(p), ctypes.c_void_p(phys.ctypes.data), _p(s), _p(a),
                                            ctypes.c_float(dt), s.shape[0], _p(out), _stream(s)))
        ctx.save_for_backward(p, s, a)
        ctx.dt, ctx.phys, ctx.system = dt, phys, system
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, s, a = ctx.saved_tensors
        lib = _capi.lib()
        g = grad_out.contiguous().float()
        gs, ga, gp = torch.empty_like(s), torch.empty_like(a), torch.empty_like(p)
        n = s.shape[0]
        with torch.cuda.device(s.device):
            ws = torch.empty(lib.apg_learnt_workspace_bytes(ctx.system, n), dtype=torch.uint8, device=s.device)
            _capi.check(lib.apg_learnt_step_adjoint(ctx.system, _p(p), ctypes.c_void_p(ctx.phys.ctypes.data), _p(s),
                                                    _p(a), ctypes.c_float(ctx.dt), n, _p(g), _p(gs), _p(ga), _p(gp),
                                                    _p(ws), _stream(s)))
        return gp, gs, ga, None, None, None


class LearntDynamics(nn.Module, FlightmareDynamics):
    def __init__(self, initial_params={}):
        FlightmareDynamics.__init__(self, initial_params)
        nn.Module.__init__(self)
        self.linear_at = nn.Parameter(torch.diag(torch.ones(4)), requires_grad=True)
        self.linear_state_1 = nn.Linear(16, 64)
        nn.init.constant_(self.linear_state_1.weight, 0)
        nn.init.constant_(self.linear_state_1.bias, 0)
        self.linear_state_2 = nn.Linear(64, 12)
        nn.init.constant_(self.linear_state_2.weight, 0)
        nn.init.constant_(self.linear_state_2.bias, 0)
        mass, inertia, kinv = float(self.mass), self.inertia_vector, self.kinv_ang_vel_tau
        self.mass = nn.Parameter(torch.tensor([mass]), requires_grad=True)
        self.torch_inertia_vector = nn.Parameter(torch.from_numpy(np.asarray(inertia)).float(), requires_grad=True)
        self.torch_kinv_vector = nn.Parameter(torch.tensor(np.asarray(kinv)).float(), requires_grad=True)

    def _flat(self):
        """named_parameters() order of the reference class: linear_at, mass, torch_inertia_vector, torch_kinv_vector,
        linear_state_1.{weight,bias}, linear_state_2.{weight,bias}"""
        ps = (self.linear_at, self.mass, self.torch_inertia_vector, self.torch_kinv_vector,
              self.linear_state_1.weight, self.linear_state_1.bias, self.linear_state_2.weight,
              self.linear_state_2.bias)
        return torch.cat([p.reshape(-1) for p in ps])

    def state_transformer(self, state, action):
        x = torch.cat((state, action), dim=1)
        return self.linear_state_2(torch.relu(self.linear_state_1(x)))

    def forward(self, state, action, dt):
        return _LearntStep.apply(self._flat(), state, action, float(dt), self.phys, 0)

    def __call__(self, state, action, dt):
        return nn.Module.__call__(self, state, action, dt)
