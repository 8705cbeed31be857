"""Single cartpole environment of the evaluation loop (reference: ``neural_control/environments/cartpole_env.py:26-115``
without rendering / image dynamics).  N carts at once: ``evaluate.CartpoleBalanceEvaluator``."""
import math

import numpy as np
import torch

from .. import environments as _env

_LIMITS = (2.4, 7.5, math.pi, 7.5)          # |x|, |x_dot|, |theta|, |theta_dot| of a uniformly drawn state


def _wrap_angle(theta):
    """theta in (-pi, pi] (:76-80)"""
    if theta > math.pi:
        return theta - 2 * math.pi
    if theta <= -math.pi:
        return theta + 2 * math.pi
    return theta


class CartPoleEnv:
    def __init__(self, dynamics, dt, thresh_div=.21):
        self.dynamics, self.dt, self.thresh_div = dynamics, dt, thresh_div
        self.x_threshold = _LIMITS[0]
        self.state_limits = np.array(_LIMITS)
        self.viewer, self.steps_beyond_done = None, None
        self.state = self._reset()

    def is_upright(self):
        return bool(abs(self.state[2]) < self.thresh_div)

    def _step(self, action, image=None, state_action_buffer=None, is_torch=True):
        """action: (1,) tensor (or a list with ``is_torch=False``) -> new state (4,) float32, theta wrapped"""
        dev = _env.compute_device()
        force = action if is_torch else torch.as_tensor(action)
        cart = torch.as_tensor(np.asarray(self.state, dtype=np.float32))[None].to(dev)
        nxt = self.dynamics(cart, force.float().reshape(1, -1).to(dev), dt=self.dt)
        self.state = nxt[0].cpu().numpy()
        self.state[2] = _wrap_angle(self.state[2])
        return self.state

    def _uniform_state(self):
        return (2 * np.random.rand(4) - 1) * self.state_limits

    def _reset(self):
        """anywhere in the state box (:84-93)"""
        self.state, self.steps_beyond_done = self._uniform_state(), None
        return self.state.copy()

    def _reset_swingup(self):
        """cart centred and slow, pole hanging: |theta| in [2.8, 3.1) with a random sign (:95-105)"""
        s = self._uniform_state()
        s[0], s[1] = 0.0, 0.1 * s[1]
        sign = -1.0 if np.random.rand() > 0.5 else 1.0
        s[2] = sign * (2.8 + 0.3 * np.random.rand())
        s[3] = 0.1 * s[3]
        self.state = s
        return self.state

    def _reset_upright(self):
        """close to upright: every component within +-0.15, theta within +-0.05 (:107-115)"""
        s = 0.3 * (np.random.rand(4) - 0.5)
        s[2] = 0.1 * (np.random.rand() - 0.5)
        self.state = s
        return self.state


def construct_states(num_data, dt, save_path=None, thresh_div=.21, **kwargs):
    """Training states of the cartpole trainer (reference: ``construct_states``, environments/cartpole_env.py:178-236):
    80 % from runs of 20 small random pushes out of slowed-down random states, the rest from random pushes near the
    upright position until the pole leaves the threshold.  Every push is one launch of the CUDA dynamics op."""
    from ..dynamics.cartpole_dynamics import CartpoleDynamics
    env = CartPoleEnv(CartpoleDynamics(), dt, thresh_div=thresh_div)
    data = []
    while len(data) < num_data * .8:
        env._reset()
        env.state[1] *= .2
        env.state[3] *= .2
        for _ in range(20):
            data.append(env._step((np.random.rand() - 0.5) * .2, is_torch=False))
        env._reset()
    while len(data) < num_data:
        env.state = (np.random.rand(4) - .5) * .1
        while env.is_upright():
            data.append(env._step(np.random.rand() - 0.5, is_torch=False))
        env._reset()
    return np.array(data)[:num_data]
