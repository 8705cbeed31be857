"""Single cartpole environment of the evaluation loop (reference: ``neural_control/environments/cartpole_env.py:26-115``
without rendering / image dynamics).  N carts at once: ``evaluate.CartpoleBalanceEvaluator``."""
import numpy as np
import torch

from .. import environments as _env


class CartPoleEnv:
    def __init__(self, dynamics, dt, thresh_div=.21):
        self.dynamics, self.dt, self.thresh_div = dynamics, dt, thresh_div
        self.x_threshold = 2.4
        self.state_limits = np.array([2.4, 7.5, np.pi, 7.5])
        self.viewer = None
        self.state = self._reset()
        self.steps_beyond_done = None

    def is_upright(self):
        return bool(-self.thresh_div < self.state[2] < self.thresh_div)

    def _step(self, action, image=None, state_action_buffer=None, is_torch=True):
        """action: (1,) tensor (or a list with ``is_torch=False``) -> new state (4,) float32, theta in (-pi, pi]"""
        dev = _env.compute_device()
        s = torch.tensor([list(self.state)]).float().to(dev)
        a = (action if is_torch else torch.tensor([action])).float().reshape(1, -1).to(dev)
        self.state = self.dynamics(s, a, dt=self.dt)[0].cpu().numpy()
        theta = self.state[2]
        if theta > np.pi:
            self.state[2] = theta - 2 * np.pi
        if theta <= -np.pi:
            self.state[2] = 2 * np.pi + theta
        return self.state

    def _reset(self):
        self.state = (np.random.rand(4) * 2 - 1) * self.state_limits
        self.steps_beyond_done = None
        return np.array(self.state)

    def _reset_swingup(self):
        self.state = (np.random.rand(4) * 2 - 1) * self.state_limits
        self.state[0] = 0
        self.state[1] *= 0.1
        rand_sign = (-1) if np.random.rand() > .5 else 1
        self.state[2] = rand_sign * (2.8 + np.random.rand() * .3)
        self.state[3] *= 0.1
        return self.state

    def _reset_upright(self):
        self.state = (np.random.rand(4) - .5) * .3
        self.state[2] = (np.random.rand(1) - .5) * .1
        return self.state
