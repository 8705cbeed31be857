"""Differentiable cartpole step (reference: ``CartpoleDynamics`` in
``neural_control/dynamics/cartpole_dynamics.py:21-119``) on CUDA tensors (csrc/apg_math.cuh ``Cartpole``)."""
from ... import params as P
from ...ops import dynamics_step

gravity = 9.81
target_state = 0


class CartpoleDynamics:
    def __init__(self, modified_params={}, test_time=0, batch_size=1):
        self.batch_size, self.test_time = batch_size, test_time
        self.cfg = P.cartpole_cfg(modified_params)
        self.timestamp = 0
        self.enforce_contact = -1
        self.phys = P.cartpole_phys(modified_params)

    def __call__(self, state, action, dt):
        return self.simulate_cartpole(state, action, dt)

    def simulate_cartpole(self, state, action, delta_t):
        self.timestamp += .05            # side effect kept from the reference (:57)
        return dynamics_step("cartpole", self.phys, state.reshape(-1, 4), action.reshape(-1, 1), delta_t)
