"""Differentiable fixed-wing step (reference: ``FixedWingDynamics`` in
``neural_control/dynamics/fixed_wing_dynamics.py:13-267``) on CUDA tensors (csrc/apg_math.cuh ``Wing``), and the
learnt variant ``LearntFixedWingDynamics`` (:270-326): every physical constant a live parameter + a residual MLP, one
fused CUDA kernel forward and one hand-written adjoint kernel backward (csrc/learnt_wing_math.cuh)."""
import math

import numpy as np
import torch

from ... import params as P
from ...ops import dynamics_step

alpha_bound = float(10 / 180 * np.pi)


class FixedWingDynamics:
    def __init__(self, modified_params={}):
        self.cfg = P.wing_cfg(modified_params)
        self.pi = math.pi
        c = self.cfg
        self.I = torch.tensor([[c["I_xx"], 0, -c["I_xz"]], [0, c["I_yy"], 0], [-c["I_xz"], 0, c["I_zz"]]])
        self.phys = P.wing_phys(modified_params)

    def normalize_action(self, thrust, ome_x, ome_y, ome_z):
        """controls from the [0,1] policy outputs (fixed_wing_dynamics.py:41-46)"""
        deg = self.pi / 180
        return thrust * 7, deg * (ome_x * 40 - 20), deg * (ome_y * 5 - 2.5), deg * (ome_z * 40 - 20)

    def __call__(self, state, action, dt):
        return self.simulate_fixed_wing(state, action, dt)

    def simulate_fixed_wing(self, state, action, dt):
        return dynamics_step("wing", self.phys, state, action, dt)


class LearntFixedWingDynamics(torch.nn.Module, FixedWingDynamics):
    """Trainable fixed-wing dynamics: parameters as in the reference (``I`` (3,3), ``cfg.<key>`` one-element parameters
    in a ParameterDict, which orders its keys by sorting; ``linear_state_1`` / ``linear_state_2`` zero-initialised)."""

    def __init__(self, modified_params={}):
        FixedWingDynamics.__init__(self, modified_params)
        torch.nn.Module.__init__(self)
        c = self.cfg
        self.I = torch.nn.Parameter(torch.tensor([[c["I_xx"], 0, -c["I_xz"]], [0, c["I_yy"], 0],
                                                  [-c["I_xz"], 0, c["I_zz"]]]), requires_grad=True)
        self.cfg = torch.nn.ParameterDict({k: torch.nn.Parameter(torch.tensor([float(v)]), requires_grad=True)
                                           for k, v in sorted(c.items()) if "I_" not in k})
        self.linear_state_1 = torch.nn.Linear(16, 64)
        torch.nn.init.constant_(self.linear_state_1.weight, 0)
        torch.nn.init.constant_(self.linear_state_1.bias, 0)
        self.linear_state_2 = torch.nn.Linear(64, 12)
        torch.nn.init.constant_(self.linear_state_2.weight, 0)
        torch.nn.init.constant_(self.linear_state_2.bias, 0)

    def _flat(self):
        ps = [self.I] + [self.cfg[k] for k in sorted(self.cfg.keys())] + \
             [self.linear_state_1.weight, self.linear_state_1.bias, self.linear_state_2.weight,
              self.linear_state_2.bias]
        return torch.cat([p.reshape(-1) for p in ps])

    def state_transformer(self, state, action):
        x = torch.cat((state, action), dim=1)
        return self.linear_state_2(torch.relu(self.linear_state_1(x)))

    def forward(self, state, action, dt):
        from .quad_dynamics_trained import _LearntStep
        return _LearntStep.apply(self._flat(), state, action, float(dt), self.phys, 1)

    def __call__(self, state, action, dt):
        return torch.nn.Module.__call__(self, state, action, dt)
