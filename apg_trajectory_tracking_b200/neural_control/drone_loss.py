"""Tracking losses of the rollout (reference: ``neural_control/drone_loss.py:12-39, 72-82, 136-145``): plain sums
over batch, horizon and components.  These are the un-fused forms for callers that hold the state / action
tensors; the fused kernels accumulate the same terms per step (csrc/apg_math.cuh ``loss`` / ``loss_grad``).
Unlike the reference this module does NOT switch on ``torch.autograd.set_detect_anomaly`` at import."""
import torch

from ..ops import _require_cuda


def _ssq(x):
    return (x * x).sum()


def quad_mpc_loss(states, ref_states, action_seq, printout=0):
    _require_cuda(states, ref_states, action_seq)
    pos = _ssq(states[:, :, 0:3] - ref_states[:, :, 0:3])
    vel = _ssq(states[:, :, 6:9] - ref_states[:, :, 6:9])
    rates = _ssq(states[:, :, 9:12])
    thrust = _ssq(action_seq[:, :, 0] - 0.5)
    body = _ssq(action_seq[:, :, 1:] - 0.5)
    return 10 * pos + vel + 0.1 * rates + 0.1 * body + 5 * thrust


def fixed_wing_mpc_loss(drone_states, linear_reference, action, printout=0):
    _require_cuda(drone_states, linear_reference, action)
    return 10 * _ssq(drone_states[:, :, :3] - linear_reference) + 0.1 * _ssq(action[:, :, 1:] - 0.5)


def cartpole_loss_mpc(states, ref_states, actions):
    _require_cuda(states, ref_states, actions)
    w = torch.tensor([0.0, 3.0, 10.0, 1.0], device=states.device, dtype=states.dtype)
    return (((states - ref_states) ** 2) * w).sum() + 0.01 * _ssq(actions)
