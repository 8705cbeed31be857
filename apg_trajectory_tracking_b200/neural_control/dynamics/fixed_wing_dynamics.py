"""Differentiable fixed-wing step (reference: ``FixedWingDynamics`` in
``neural_control/dynamics/fixed_wing_dynamics.py:13-267``) on CUDA tensors (csrc/apg_math.cuh ``Wing``)."""
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
