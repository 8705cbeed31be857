"""Quadrotor parameters and rotation helpers (reference: ``neural_control/dynamics/quad_dynamics_base.py``).

Constants come from ``apg_trajectory_tracking_b200.params`` (the values of the reference's config_quad.json),
overridable through the reference's ``modified_params`` dict.  The casadi members of the reference class belong to
its MPC baseline and are not part of this package."""
import numpy as np
import torch

from ... import params as P


class Dynamics:
    def __init__(self, modified_params={}):
        self.cfg = P.quad_cfg(modified_params)
        self.mass = self.cfg["mass"]
        self.arm_length = self.cfg["arm_length"]
        self.kinv_ang_vel_tau = np.array(self.cfg["kinv_ang_vel_tau"])
        self.inertia_vector = self.mass / 12.0 * self.arm_length ** 2 * np.array(self.cfg["frame_inertia"])
        self.torch_translational_drag = torch.tensor(self.cfg["translational_drag"]).float()
        self.torch_gravity = torch.tensor(self.cfg["gravity"])
        self.torch_rotational_drag = torch.tensor(self.cfg["rotational_drag"]).float()
        self.torch_inertia_vector = torch.from_numpy(self.inertia_vector).float()
        self.torch_inertia_J = torch.diag(self.torch_inertia_vector)
        self.torch_inertia_J_inv = torch.diag(1 / self.torch_inertia_vector)
        self.torch_kinv_vector = torch.tensor(self.kinv_ang_vel_tau).float()
        self.torch_kinv_ang_vel_tau = torch.diag(self.torch_kinv_vector)
        self.phys = P.quad_phys(modified_params)

    @staticmethod
    def _trig(attitude):
        r, p, y = attitude[:, 0], attitude[:, 1], attitude[:, 2]
        return torch.cos(r), torch.sin(r), torch.cos(p), torch.sin(p), torch.cos(y), torch.sin(y)

    @staticmethod
    def world_to_body_matrix(attitude):
        """(N,3) roll/pitch/yaw -> (N,3,3) world->body rotation (rows as in quad_dynamics_base.py:59-94)"""
        cr, sr, cp, sp, cy, sy = Dynamics._trig(attitude)
        rows = (torch.stack((cy * cp, sy * cp, -sp), 1),
                torch.stack((cy * sp * sr - cr * sy, cr * cy + sr * sy * sp, cp * sr), 1),
                torch.stack((cy * sp * cr + sr * sy, cr * sy * sp - cy * sr, cr * cp), 1))
        return torch.stack(rows, 1)

    @staticmethod
    def to_euler_matrix(attitude):
        cr, sr, cp, sp, _, _ = Dynamics._trig(attitude)
        one, zero = torch.ones_like(cr), torch.zeros_like(cr)
        rows = (torch.stack((one, zero, -sp), 1), torch.stack((zero, cr, cp * sr), 1),
                torch.stack((zero, -sr, cp * cr), 1))
        return torch.stack(rows, 1)

    @staticmethod
    def euler_rate(attitude, angular_velocity):
        return torch.squeeze(torch.matmul(Dynamics.to_euler_matrix(attitude),
                                          angular_velocity.float().unsqueeze(2)))
