"""Quadrotor trainer (reference: ``scripts/train_drone.py`` ``TrainDrone``): concurrent (:175-203) and
autoregressive / LSTM (:113-173) train steps."""
import torch

from ..neural_control.dataset import state_preprocessing
from ..neural_control.drone_loss import quad_mpc_loss
from ..neural_control.models.hutter_model import Net
from ..neural_control.models.rnn import LSTM_NEW
from .train_base import TrainBase


class TrainDrone(TrainBase):
    def __init__(self, train_dynamics, eval_dynamics, config):
        self.config = config
        super().__init__(train_dynamics, eval_dynamics, **config)
        if self.sample_in not in ("eval_env", "train_env"):
            raise ValueError("sample in must be one of eval_env, train_env")

    def initialize_model(self, base_model=None, state_data=None, in_state_size=15):
        """Net(15, h, 9, 4h) / Net(15, h, 9, 4) / LSTM_NEW(15, h, 9, 4) as in train_drone.py:81-88"""
        if base_model is not None:
            self.net = base_model
        else:
            cls = LSTM_NEW if self.train_mode == "LSTM" else Net
            self.net = cls(in_state_size, self.horizon, self.ref_dim, self.actions_out_dim, conv=1)
        self.state_data = state_data
        self.init_optimizer()

    def train_controller_model(self, current_state, action_seq, in_ref_states, ref_states):
        """un-fused concurrent step for callers that evaluated the policy themselves (autograd over the per-step
        CUDA ops); ``run_epoch`` uses the fused path instead"""
        self.optimizer_controller.zero_grad()
        states = []
        for k in range(self.horizon):
            current_state = self.train_dynamics(current_state, action_seq[:, k], dt=self.delta_t)
            states.append(current_state)
        loss = quad_mpc_loss(torch.stack(states, 1), ref_states, action_seq, printout=0)
        loss.backward()
        self.optimizer_controller.step()
        return loss

    def train_recurrent_model(self, in_state, current_state, in_ref_states, ref_states):
        """fused autoregressive / LSTM step; window semantics per ``self.window`` (cumulative == the reference's
        forward, whose own backward() raises)"""
        return self.fused_train_step(in_state, current_state, in_ref_states, ref_states)

    def recurrent_features(self, current_state):
        return state_preprocessing(current_state)
