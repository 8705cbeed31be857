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


def full_state_training_data(len_data, ref_length=5, dt=0.02, speed_factor=.6, data_dir="data/traj_data_1", device=None,
                             **kwargs):
    """Training samples cut from random trajectory files (reference: ``full_state_training_data``,
    environments/drone_env.py:232-269): every ``2 * ref_length``-th row of a table is a drone state (body rates 0),
    the following ``ref_length`` rows its reference.  The table layout (``load_prepare_trajectory``) and the window
    cutting run on the device (``prepare.reference_table`` / ``prepare.sample_windows``); file choice by
    ``np.random.choice`` like the reference.  Returns numpy (states (len_data,12), ref_states (len_data,ref_length,9))."""
    import os
    from ... import prepare as PR
    from .. import environments as _e
    dev = torch.device(device) if device is not None else _e.compute_device()
    folder = os.path.join(data_dir, "train")
    if not os.path.isdir(folder):
        # the reference's trajectory files (data/traj_data_1, made by its casadi generator) are not part of its
        # checkout: seeded polynomial trajectories in the same table layout instead (BASELINE.md synthetic inputs)
        return synthetic_training_data(len_data, ref_length=ref_length, dt=dt, device=dev)
    names = sorted(os.listdir(folder))
    states, refs, have = [], [], 0
    sample_freq = 2 * ref_length
    while have < len_data:
        raw = np.load(os.path.join(folder, np.random.choice(names)))
        table = PR.reference_table(torch.as_tensor(raw, dtype=torch.float32).to(dev), dt, speed_factor, z_offset=0.0)
        n = len(range(0, table.shape[0] - (ref_length + 1), sample_freq))
        if n <= 0:
            raise ValueError("trajectory file too short for the requested reference length")
        s, r = PR.sample_windows(table, n, ref_length, sample_freq)
        states.append(s.cpu().numpy().astype(np.float64))
        refs.append(r.cpu().numpy().astype(np.float64))
        have += n
    return np.concatenate(states)[:len_data], np.concatenate(refs)[:len_data]


def synthetic_training_data(len_data, ref_length=5, dt=0.02, device=None, seed=None):
    """Stand-in for ``full_state_training_data`` when no trajectory files exist: per sample a degree-5 polynomial per
    axis (coefficients as SURVEY.md 8d: c1 ~ U(-1.5, 1.5), c_i ~ U(-.5, .5) / i!), sampled on the device
    (``prepare.poly_reference``: rows ``[p(t), 0, p'(t)]`` at t = dt .. ref_length * dt); the drone starts on the
    trajectory at t = 0 (random position in [-1, 1]^3 added to both) with small attitude and velocity noise.
    Returns numpy (states (len_data, 12), ref_states (len_data, ref_length, 9)) like the reference function."""
    from ... import prepare as PR
    from .. import environments as _e
    dev = torch.device(device) if device is not None else _e.compute_device()
    rs = np.random.RandomState(seed)
    coef = np.zeros((len_data, 3, 6), dtype=np.float32)
    coef[:, :, 1] = rs.uniform(-1.5, 1.5, size=(len_data, 3))
    fact = 1.0
    for i in range(2, 6):
        fact *= i
        coef[:, :, i] = rs.uniform(-0.5, 0.5, size=(len_data, 3)) / fact
    refs = PR.poly_reference(torch.as_tensor(coef).to(dev), ref_length, dt).cpu().numpy().astype(np.float64)
    start = rs.uniform(-1.0, 1.0, size=(len_data, 3))
    refs[:, :, :3] += start[:, None, :]
    states = np.zeros((len_data, 12))
    states[:, :3] = start
    states[:, 3:6] = rs.uniform(-0.2, 0.2, size=(len_data, 3))
    states[:, 6:9] = coef[:, :, 1] + rs.normal(0.0, 0.3, size=(len_data, 3))
    return states, refs
