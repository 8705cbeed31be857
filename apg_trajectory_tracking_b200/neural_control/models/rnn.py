"""Recurrent policy: conv reference encoder -> LSTMCell(state_dim + 20*(horizon-2), 8) -> Linear(8, actions).

Interface of the reference's ``neural_control/models/rnn.py:7-50`` (class name, constructor, ``reset_hidden_state``,
the ``hidden_state`` / ``cell_state`` attributes, parameter registration order).  The initial state is drawn exactly
like the reference does (two ``torch.randn(batch, 8)`` draws on the CPU generator, hidden first) and handed to the
fused kernels as an input."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...ops import _require_cuda


class LSTM_NEW(nn.Module):
    HIDDEN = 8
    CONV_CHANNELS = 20

    def __init__(self, state_dim, horizon, ref_dim, nr_actions_predict, conv=True):
        super().__init__()
        self.state_dim, self.ref_dim, self.horizon, self.conv = state_dim, ref_dim, horizon, conv
        self.reshape_len = self.CONV_CHANNELS * (horizon - 2) if conv else 64
        self.conv_ref = nn.Conv1d(ref_dim, self.CONV_CHANNELS, kernel_size=3)
        self.ref_in = nn.Linear(horizon * ref_dim, 64)
        self.fc_out = nn.Linear(self.HIDDEN, nr_actions_predict)
        self.lstm = nn.LSTMCell(state_dim + self.reshape_len, self.HIDDEN)
        self.reset_hidden_state(1)

    def used_parameter_names(self):
        skip = "ref_in." if self.conv else "conv_ref."
        return [n for n, _ in self.named_parameters() if not n.startswith(skip)]

    def reset_hidden_state(self, batch_size=1):
        dev = self.fc_out.weight.device
        self.hidden_state = torch.randn(batch_size, self.HIDDEN).to(dev)
        self.cell_state = torch.randn(batch_size, self.HIDDEN).to(dev)

    def forward(self, state, ref):
        _require_cuda(state, ref)
        if self.conv:
            r = F.relu(F.conv1d(ref.transpose(1, 2), self.conv_ref.weight, self.conv_ref.bias))
            r = r.reshape(-1, self.reshape_len)
        else:
            r = torch.tanh(self.ref_in(ref))
        x = torch.cat((state, r), dim=1)
        self.hidden_state, self.cell_state = self.lstm(x, (self.hidden_state.to(x.device),
                                                           self.cell_state.to(x.device)))
        return self.fc_out(self.hidden_state)
