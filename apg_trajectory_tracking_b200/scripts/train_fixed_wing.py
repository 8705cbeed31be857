"""Fixed-wing trainer (reference: ``scripts/train_fixed_wing.py`` ``TrainFixedWing``, concurrent only, :33-37, 90-116)."""
import torch

from ..neural_control.drone_loss import fixed_wing_mpc_loss
from ..neural_control.models.hutter_model import Net
from .train_base import TrainBase


class TrainFixedWing(TrainBase):
    def __init__(self, train_dynamics, eval_dynamics, config):
        self.config = config
        super().__init__(train_dynamics, eval_dynamics, **config)
        if self.train_mode != "concurrent":
            raise NotImplementedError("the fixed wing is trained in concurrent mode only (train_fixed_wing.py:33-37)")

    def rollout_dt(self):
        return self.delta_t_train

    def initialize_model(self, base_model=None, state_data=None):
        """Net(12-3, 1, 3, 4h, conv=False) as in train_fixed_wing.py:67-73"""
        self.net = base_model if base_model is not None else Net(self.state_size - self.ref_dim, 1, self.ref_dim,
                                                                 self.action_dim * self.horizon, conv=False)
        self.state_data = state_data
        self.init_optimizer()

    def train_controller_model(self, current_state, action_seq, in_ref_state, ref_states):
        self.optimizer_controller.zero_grad()
        states = []
        for k in range(self.horizon):
            current_state = self.train_dynamics(current_state, action_seq[:, k], dt=self.delta_t_train)
            states.append(current_state)
        loss = fixed_wing_mpc_loss(torch.stack(states, 1), ref_states, action_seq, printout=0)
        loss.backward()
        self.optimizer_controller.step()
        return loss
