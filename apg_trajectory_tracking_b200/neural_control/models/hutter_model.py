"""Policy MLP of the quadrotor / fixed-wing controllers.

Same constructor, parameter names and shapes as the reference's ``neural_control/models/hutter_model.py:12-30`` so
that its pickled models (``torch.save(net)``) load against this class and ``net.parameters()`` has the order the
fused kernels expect (csrc/layouts.h).  ``forward`` is the un-fused evaluation used by inference-style callers; the
train step goes through ``apg_trajectory_tracking_b200.train`` which evaluates the whole rollout in two launches."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...ops import _require_cuda


class Net(nn.Module):
    HIDDEN = 64
    CONV_CHANNELS = 20

    def __init__(self, state_dim, horizon, ref_dim, nr_actions_predict, conv=True):
        super().__init__()
        hid = self.HIDDEN
        self.horizon, self.conv = horizon, conv
        self.reshape_len = self.CONV_CHANNELS * (horizon - 2) if conv else hid
        # registration order == parameter order of the reference
        self.states_in = nn.Linear(state_dim, hid)
        self.conv_ref = nn.Conv1d(ref_dim, self.CONV_CHANNELS, kernel_size=3)
        self.ref_in = nn.Linear(horizon * ref_dim, hid)
        self.fc1 = nn.Linear(hid + self.reshape_len, hid)
        self.fc2 = nn.Linear(hid, hid)
        self.fc3 = nn.Linear(hid, hid)
        self.fc_out = nn.Linear(hid, nr_actions_predict)

    # which tensors the forward actually uses (the others keep grad None, like in the reference)
    def used_parameter_names(self):
        skip = "ref_in." if self.conv else "conv_ref."
        return [n for n, _ in self.named_parameters() if not n.startswith(skip)]

    def forward(self, state, ref):
        _require_cuda(state, ref)
        s = torch.tanh(F.linear(state, self.states_in.weight, self.states_in.bias))
        if self.conv:
            r = F.relu(F.conv1d(ref.transpose(1, 2), self.conv_ref.weight, self.conv_ref.bias))
            r = r.reshape(-1, self.reshape_len)                    # channel-major: c*(h-2)+t
        else:
            r = torch.tanh(F.linear(ref, self.ref_in.weight, self.ref_in.bias))
        x = torch.cat((s, r), dim=1)
        for fc in (self.fc1, self.fc2, self.fc3):
            x = torch.tanh(fc(x))
        return self.fc_out(x)
