"""Single-quadrotor environment of the evaluation loop (reference: ``neural_control/environments/drone_env.py:32-210``
without the renderer): holds one state, ``step`` applies one action through the CUDA dynamics op (one launch per
step; N drones at once go through ``evaluate.TableEvaluator``).  State layout [pos, euler rpy, vel, body rates]."""
import math

import numpy as np
import torch

from .. import environments as _env


class _State:
    """what the callers use of the reference's ``DynamicsState``: the flat view and the named slices"""

    def __init__(self):
        self._v = np.zeros(12)
        self._last_velocity = np.zeros(3)

    @property
    def as_np(self):
        return self._v.copy()

    def from_np(self, arr):
        self._last_velocity = self._v[6:9].copy()
        self._v = np.array(arr, dtype=np.float64).copy()

    position = property(lambda self: self._v[0:3])
    attitude = property(lambda self: self._v[3:6])
    velocity = property(lambda self: self._v[6:9])
    angular_velocity = property(lambda self: self._v[9:12])


class QuadRotorEnvBase:
    def __init__(self, dynamics, dt):
        self._state = _State()
        self.random_state = np.random.RandomState()
        self.dt, self.dynamics = dt, dynamics
        self.renderer = None

    def seed(self, seed=None):
        self.random_state = np.random.RandomState(seed)
        return [seed]

    @staticmethod
    def get_is_stable(np_state, thresh=.4):
        """only roll and pitch are constrained (:66-74)"""
        return bool(np.all(np.absolute(np_state[3:5]) < thresh))

    def get_acceleration(self):
        return (self._state.velocity - self._state._last_velocity) / self.dt

    def step(self, action, thresh=.4):
        """-> (new state (12,), still stable)   (:83-115)"""
        action = np.clip(np.asarray(action, dtype=np.float64), 0.0, 1.0)
        if action.shape != (4,):
            raise AssertionError(f"action not size 4 but {action.shape}")
        dev = _env.compute_device()
        s = torch.from_numpy(self._state.as_np[None]).to(dev)
        a = torch.from_numpy(action[None]).float().to(dev)
        out = self.dynamics(s, a, dt=self.dt).cpu().numpy()[0]
        self._state.from_np(out)
        return out, self.get_is_stable(out, thresh=thresh)

    def zero_reset(self, position_x=0, position_y=0, position_z=2):
        """zero velocities and attitude at the given position (:129-142)"""
        self._state = _State()
        self._state.from_np(np.array([position_x, position_y, position_z] + [0.0] * 9))
        self._state._last_velocity = np.zeros(3)
        return self._state.as_np

    def reset(self, strength=.8):
        """random state (:151-172): roll / pitch within 3*strength degrees, yaw in [-1.5, 1.5], body rates within
        2*strength (yaw rate halved), position in [-1, 1]^3, velocity within 3 m/s"""
        rs, v = self.random_state, np.zeros(12)
        mpr = 3 * strength * math.pi / 180
        v[3], v[4] = rs.uniform(-mpr, mpr), rs.uniform(-mpr, mpr)
        rs.uniform(-math.pi, math.pi)                               # random_angle's yaw draw, overwritten below
        v[9:12] = rs.uniform(-2.0 * strength, 2.0 * strength, size=3)
        v[5] = rs.uniform(-1.5, 1.5)
        v[0:3] = np.random.rand(3) * 2 - 1
        v[11] *= 0.5
        v[6:9] = rs.uniform(-3, 3, size=3)
        self._state = _State()
        self._state.from_np(v)
        self._state._last_velocity = v[6:9].copy()
        return self._state

    def render_reset(self, strength=.8):
        self.reset(strength=strength)
        self._state.position[2] += 2

    def get_copter_state(self):
        return self._state

    def close(self):
        pass
