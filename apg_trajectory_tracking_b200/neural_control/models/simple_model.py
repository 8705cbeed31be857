"""Cartpole policy MLP 4 -> 32 -> 64 -> 64 -> 32 -> out, tanh after every layer including the last
(reference: ``neural_control/models/simple_model.py:9-28``; only ``Net`` is on the rollout path)."""
import torch
import torch.nn as nn

from ...ops import _require_cuda


class Net(nn.Module):
    def __init__(self, in_size, out_size):
        super().__init__()
        widths = [in_size, 32, 64, 64, 32]
        self.fc0 = nn.Linear(widths[0], widths[1])
        self.fc1 = nn.Linear(widths[1], widths[2])
        self.fc2 = nn.Linear(widths[2], widths[3])
        self.fc3 = nn.Linear(widths[3], widths[4])
        self.fc_out = nn.Linear(widths[4], out_size)

    def used_parameter_names(self):
        return [n for n, _ in self.named_parameters()]

    def forward(self, x):
        _require_cuda(x)
        x[:, 0] *= 0          # the reference zeroes the cart position IN PLACE on the caller's tensor (:21)
        for fc in (self.fc0, self.fc1, self.fc2, self.fc3, self.fc_out):
            x = torch.tanh(fc(x))
        return x
