"""Single fixed-wing environment of the evaluation loop (reference: ``neural_control/environments/wing_env.py:12-58``
without the renderer).  N flights at once: ``evaluate.WingTargetEvaluator``."""
import numpy as np
import torch

from .. import environments as _env


class SimpleWingEnv:
    def __init__(self, dynamics, dt):
        self.dt, self.dynamics = dt, dynamics
        self.renderer = None

    def zero_reset(self):
        self._state = np.zeros(12)
        self._state[3] = 11.5

    def reset(self):
        """the reference's randomised start (:30-42): a 6-entry longitudinal state [x, z, u ~ 10 +- .5, w +- .5,
        pitch +- 2 deg, pitch rate +- .005] - kept as it is there (``zero_reset`` is what the evaluators use)"""
        u, w = np.random.rand(1) - .5 + 10, np.random.rand(1) - .5
        pitch, rate = np.deg2rad(np.random.rand(1) * 4 - 2), np.random.rand(1) * 0.01 - 0.005
        self._state = np.array([0, 0, u[0], w[0], pitch[0], rate[0]])

    def step(self, action, thresh_stable=.7):
        """-> (new state (12,) float32, |roll| and |pitch| below thresh_stable)   (:44-58)"""
        dev = _env.compute_device()
        a = torch.tensor([np.asarray(action).tolist()]).float().to(dev)
        s = torch.tensor([np.asarray(self._state).tolist()]).float().to(dev)
        self._state = self.dynamics(s, a, self.dt)[0].cpu().numpy()
        return self._state, bool(np.all(np.absolute(self._state[6:8]) < thresh_stable))

    def close(self):
        pass


def run_wing_flight(env, traj_len=1000, render=0, **kwargs):
    """one open-loop flight from ``zero_reset`` with actions around the prior [.25, .5, .5, .5], redrawn every 10
    steps (N(0, .15), clipped to [0, 1]); stops at the first unstable state; returns the visited states (:72-95)"""
    prior = np.array([.25, .5, .5, .5])
    env.zero_reset()
    visited, action = [], prior
    for j in range(traj_len):
        if j % 10 == 0:
            action = np.clip(np.random.normal(scale=.15, size=4) + prior, 0, 1)
        state, stable = env.step(action)
        if not stable:
            break
        visited.append(state)
    return np.array(visited)


def generate_unit_vecs(num_vecs, mean_vec=[1, 0, 0], std=.15):
    """direction vectors normally distributed around ``mean_vec`` (covariance std * I); x components below 0.01 are
    set to 1 (:98-109; not normalised there either)"""
    vecs = np.random.multivariate_normal(mean_vec, np.eye(3) * std, size=num_vecs)
    vecs[vecs[:, 0] < 0.01, 0] = 1
    return vecs


def sample_training_data(num_samples, dt=0.01, take_every=10, traj_len=500, vec_std=.15, **kwargs):
    """(state, target point) pairs from open-loop flights (reference: ``sample_training_data``, environments/
    wing_env.py:112-162): every ``take_every``-th state of a flight (with a random offset below 5) is paired with the
    positions of up to 20 randomly chosen later states of the same flight (at least 10 steps ahead).  The flights run
    on the CUDA dynamics op (``run_wing_flight``).  Like the reference says, the shipped baseline model is trained on
    self-play data only; this seeds the dataset."""
    from ..dynamics.fixed_wing_dynamics import FixedWingDynamics
    use_at_each = 20
    env = SimpleWingEnv(FixedWingDynamics(), dt)
    states, refs = [], []
    leftover = num_samples
    while leftover > 0:
        traj = run_wing_flight(env, traj_len=traj_len, **kwargs)
        n_traj = len(traj)
        for i in range(min(n_traj // take_every, leftover)):
            at = int(i * take_every + np.random.rand() * 5)
            later = np.random.permutation(np.arange(at + 10, n_traj))[:use_at_each]
            for k in later:
                states.append(traj[at])
                refs.append(traj[k, :3])
        leftover = num_samples - len(refs)
    return np.array(states)[:num_samples], np.array(refs)[:num_samples]
