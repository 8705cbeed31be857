"""Cartpole trainer (reference: ``scripts/train_cartpole.py`` ``TrainCartpole``: ``make_reference`` :103-110 and the
controller branch of ``run_epoch`` :118-165)."""
import torch

from ..neural_control.models.simple_model import Net
from .train_base import TrainBase


class TrainCartpole(TrainBase):
    def __init__(self, train_dynamics, eval_dynamics, config):
        self.config = config
        config.setdefault("system", "cartpole")
        config.setdefault("state_size", 4)
        config.setdefault("action_dim", 1)
        super().__init__(train_dynamics, eval_dynamics, **config)

    def initialize_model(self, base_model=None, state_data=None):
        self.net = base_model if base_model is not None else Net(self.state_size, self.horizon * self.action_dim)
        self.state_data = state_data
        self.init_optimizer()

    def make_reference(self, current_state):
        """reference fades linearly from the current state to 0 over the horizon, last row 0 (no gradient)"""
        ref = torch.zeros(current_state.size()[0], self.horizon, self.state_size, device=current_state.device)
        for k in range(self.horizon - 1):
            ref[:, k] = current_state.detach() * (1 - 1 / (self.horizon - 1) * k)
        return ref

    def run_epoch(self, train="controller"):
        self.results_dict["trained"].append(train)
        running_loss, i = 0.0, 0
        for i, data in enumerate(self.trainloader, 0):
            in_state, current_state = data
            loss = self.fused_train_step(in_state, current_state, None, None)
            running_loss += loss.item()
        epoch_loss = running_loss / max(i, 1)
        self.results_dict["loss_" + train].append(epoch_loss)
        return epoch_loss
