"""Recurrent quadrotor policy: conv reference encoder -> LSTMCell(state_dim + 20*(horizon-2), 8) -> Linear(8, actions).

Interface of the reference's ``neural_control/models/rnn.py:7-50``: class name ``LSTM_NEW``, constructor arguments,
``reset_hidden_state(batch_size)``, the ``hidden_state`` / ``cell_state`` attributes and the parameter registration
order (conv_ref, ref_in, fc_out, lstm) that the fused kernels' flat parameter vector relies on (csrc/layouts.h).
The initial state is random like in the reference (two standard-normal draws of shape (batch, 8) from the global CPU
generator, hidden state first) and is handed to the fused kernels as an explicit input ``h0c0``."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...ops import _require_cuda

_HIDDEN_UNITS = 8
_ENCODER_CHANNELS = 20
_ENCODER_KERNEL = 3


class LSTM_NEW(nn.Module):
    def __init__(self, state_dim, horizon, ref_dim, nr_actions_predict, conv=True):
        super().__init__()
        self.state_dim, self.ref_dim, self.horizon, self.conv = state_dim, ref_dim, horizon, conv
        encoded = _ENCODER_CHANNELS * (horizon - _ENCODER_KERNEL + 1) if conv else 64
        self.reshape_len = encoded
        self.conv_ref = nn.Conv1d(ref_dim, _ENCODER_CHANNELS, kernel_size=_ENCODER_KERNEL)
        self.ref_in = nn.Linear(horizon * ref_dim, 64)
        self.fc_out = nn.Linear(_HIDDEN_UNITS, nr_actions_predict)
        self.lstm = nn.LSTMCell(state_dim + encoded, _HIDDEN_UNITS)
        self.reset_hidden_state(1)

    def used_parameter_names(self):
        """parameters the forward touches; the other branch of the reference encoder keeps ``grad is None``"""
        unused_prefix = "ref_in." if self.conv else "conv_ref."
        return [name for name, _ in self.named_parameters() if not name.startswith(unused_prefix)]

    def _draw_state(self, batch_size):
        return torch.randn(batch_size, _HIDDEN_UNITS).to(self.fc_out.weight.device)

    def reset_hidden_state(self, batch_size=1):
        """new sequence: fresh random hidden / cell state (draw order hidden, then cell)"""
        self.hidden_state = self._draw_state(batch_size)
        self.cell_state = self._draw_state(batch_size)

    def _encode_reference(self, ref):
        if not self.conv:
            return torch.tanh(self.ref_in(ref))
        feats = F.relu(F.conv1d(ref.transpose(1, 2), self.conv_ref.weight, self.conv_ref.bias))
        return feats.reshape(-1, self.reshape_len)          # channel-major, like the hutter net

    def forward(self, state, ref):
        _require_cuda(state, ref)
        cell_in = torch.cat((state, self._encode_reference(ref)), dim=1)
        carry = (self.hidden_state.to(cell_in.device), self.cell_state.to(cell_in.device))
        self.hidden_state, self.cell_state = self.lstm(cell_in, carry)
        return self.fc_out(self.hidden_state)
