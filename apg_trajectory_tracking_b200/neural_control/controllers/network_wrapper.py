"""Single-drone policy wrappers of the evaluation loops (reference:
``neural_control/controllers/network_wrapper.py``: ``NetworkWrapper`` :13-68, ``FixedWingNetWrapper`` :71-98,
``CartpoleWrapper`` :101-148): same constructors, counters and ``predict_actions`` signatures.  One call = one
policy evaluation for ONE drone: the sample goes through the dataset's ``get_and_add_eval_data`` (which is also the
self-play feed, every ``take_every_x``-th call) and the policy runs on the GPU.

These are the batch-1 callers next to the hot path; N drones at once go through ``apg_trajectory_tracking_b200.
evaluate`` (``TableEvaluator`` / ``WingTargetEvaluator`` / ``CartpoleBalanceEvaluator``), which replaces the
per-step host loop by one kernel launch and reproduces the same counters (``evaluate.selfplay_samples``)."""
import numpy as np
import torch


def _device_of(net):
    p = next(net.parameters())
    if not p.is_cuda:
        raise RuntimeError("the policy must live on a CUDA device (net.cuda()): there is no CPU path")
    return p.device


class NetworkWrapper:
    def __init__(self, model, dataset, optimizer=None, horizon=10, max_drone_dist=0.1, render=0, dt=0.02,
                 take_every_x=1000, **kwargs):
        self.dataset, self.net, self.optimizer = dataset, model, optimizer
        self.horizon, self.max_drone_dist, self.render, self.dt = horizon, max_drone_dist, render, dt
        self.training_means = None
        self.take_every_x = take_every_x
        self.action_counter = 0
        self.action_dim = 4

    def predict_actions(self, current_np_state, ref_states):
        """(12,) state, (>= horizon, 9) reference rows -> (horizon, 4) action sequence (concurrent nets) or (1, 4)"""
        add = (self.action_counter + 1) % self.take_every_x == 0
        in_state, _, ref, _ = self.dataset.get_and_add_eval_data(np.array(current_np_state, copy=True), ref_states,
                                                                 add_to_dataset=add)
        dev = _device_of(self.net)
        with torch.no_grad():
            act = torch.sigmoid(self.net(in_state.to(dev), ref[:, :self.horizon].to(dev)))
            if act.shape[-1] > self.action_dim:
                act = act.reshape(1, self.horizon, self.action_dim)
        self.action_counter += 1
        return act[0].cpu().numpy()


class FixedWingNetWrapper:
    def __init__(self, model, dataset, horizon=1, take_every_x=1000, **kwargs):
        self.net, self.dataset, self.horizon = model, dataset, horizon
        self.action_dim = 4
        self.action_counter = 0
        self.take_every_x = take_every_x

    def predict_actions(self, state, ref_state):
        """(12,) state, (3,) target -> (horizon, 4) actions (or (4,) for a one-step net)"""
        add = (self.action_counter + 1) % self.take_every_x == 0
        normed_state, _, normed_ref, _ = self.dataset.get_and_add_eval_data(state, ref_state, add_to_dataset=add)
        dev = _device_of(self.net)
        with torch.no_grad():
            act = torch.sigmoid(self.net(normed_state.to(dev), normed_ref.to(dev)))[0]
            if act.shape[-1] > self.action_dim:
                act = act.reshape(self.horizon, self.action_dim)
        self.action_counter += 1
        return act.cpu().numpy()


class CartpoleWrapper:
    def __init__(self, model, horizon=10, action_dim=1, **kwargs):
        self.horizon, self.action_dim, self.net = horizon, action_dim, model

    def raw_states_to_torch(self, states, normalize=False, std=None, mean=None, return_std=False):
        states = np.asarray(states)
        if states.ndim == 1:
            states = states[None]
        if normalize:
            std = np.std(states, axis=0) if std is None else std
            mean = np.mean(states, axis=0) if mean is None else mean
            states = (states - mean) / std
        else:
            std = 1
        out = torch.from_numpy(states).float()
        return (out, mean, std) if return_std else out

    def predict_actions(self, state, ref_state):
        """(4,) state -> (1, horizon, action_dim) tensor (on the policy's device).  Side effect kept from the
        reference: its network zeroes column 0 of the input IN PLACE (models/simple_model.py:21) and a float32 numpy
        ``state`` is aliased by ``torch.from_numpy(state).float()``, so the caller's array - in the evaluation loop the
        environment's own state - loses its cart position; a float64 ``state`` is copied and stays untouched.  The
        closed-loop results of ``evaluate_in_environment`` depend on it (``evaluate.CartpoleBalanceEvaluator``
        reproduces the same loop for N carts in one launch)."""
        x = self.raw_states_to_torch(state).to(_device_of(self.net))
        act = self.net(x)
        if isinstance(state, np.ndarray) and state.dtype == np.float32:
            state.reshape(-1, state.shape[-1])[:, 0] = 0
        if act.shape[-1] > self.action_dim:
            act = act.reshape(-1, self.horizon, self.action_dim)
        return act
