"""Differentiable quadrotor step (reference: ``FlightmareDynamics`` in
``neural_control/dynamics/quad_dynamics_flightmare.py:7-216``): ``dyn(state, action, dt)`` /
``dyn.simulate_quadrotor(action, state, dt)`` (note the reference's swapped argument order) on CUDA tensors, forward
and backward each one kernel launch (csrc/apg_math.cuh ``Quad``)."""
from ...ops import dynamics_step
from .quad_dynamics_base import Dynamics


class FlightmareDynamics(Dynamics):
    def __init__(self, modified_params={}, simulate_rotors=False):
        super().__init__(modified_params=modified_params)
        if simulate_rotors:
            raise NotImplementedError("the reference's rotor simulation is commented out on its own rollout path "
                                      "(quad_dynamics_flightmare.py:154-161); simulate_rotors must stay False")
        self.simulate_rotors = False

    def __call__(self, state, action, dt):
        return self.simulate_quadrotor(action, state, dt)

    def simulate_quadrotor(self, action, state, dt):
        return dynamics_step("quad", self.phys, state, action, dt)
