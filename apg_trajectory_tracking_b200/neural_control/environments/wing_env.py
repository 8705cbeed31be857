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

    def step(self, action, thresh_stable=.7):
        """-> (new state (12,) float32, |roll| and |pitch| below thresh_stable)   (:44-58)"""
        dev = _env.compute_device()
        a = torch.tensor([np.asarray(action).tolist()]).float().to(dev)
        s = torch.tensor([np.asarray(self._state).tolist()]).float().to(dev)
        self._state = self.dynamics(s, a, self.dt)[0].cpu().numpy()
        return self._state, bool(np.all(np.absolute(self._state[6:8]) < thresh_stable))

    def close(self):
        pass
