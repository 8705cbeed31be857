"""Trainer base of the rollout hot path (reference: ``scripts/train_base.py`` ``TrainBase``).

Kept: the constructor keywords that select the hot-path variant, ``init_optimizer`` (DataLoader + SGD momentum 0.9,
:130-143), ``run_epoch`` (:188-218), the abstract ``train_controller_model``.  The mini-batch body of ``run_epoch``
is ONE fused rollout (forward + adjoint kernel) followed by the torch optimizer step.  Out of scope here
(SURVEY.md section 2): evaluation, self-play data, curriculum, model saving, TensorBoard."""
from collections import defaultdict

import torch
import torch.optim as optim

from .. import train as T


class _NullWriter:
    def __getattr__(self, name):
        return lambda *a, **k: None


class TrainBase:
    def __init__(self, train_dynamics, eval_dynamics, sample_in="train_env", delta_t=0.05, delta_t_train=0.05,
                 epoch_size=500, batch_size=8, state_size=12, horizon=10, ref_dim=3, action_dim=4,
                 learning_rate_controller=0.0001, learning_rate_dynamics=0.001, train_mode="concurrent",
                 system="quad", window="cumulative", device=None, **kwargs):
        self.sample_in = sample_in
        self.delta_t, self.delta_t_train = delta_t, delta_t_train
        self.epoch_size, self.batch_size = epoch_size, batch_size
        self.state_size, self.horizon, self.ref_dim, self.action_dim = state_size, horizon, ref_dim, action_dim
        self.learning_rate_controller = learning_rate_controller
        self.learning_rate_dynamics = learning_rate_dynamics
        self.train_mode, self.system, self.window = train_mode, system, window
        self.device = device
        self.config = dict(kwargs)
        self.results_dict = defaultdict(list)
        self.results_dict["loss"].append(0)
        self.train_dynamics, self.eval_dynamics = train_dynamics, eval_dynamics
        self.state_data, self.net, self.fused = None, None, None
        self.writer = _NullWriter()
        if self.train_mode in ("autoregressive", "LSTM"):
            self.actions_out_dim, self.ref_length = self.action_dim, self.horizon * 2
        elif self.train_mode == "concurrent":
            self.actions_out_dim, self.ref_length = self.action_dim * self.horizon, self.horizon
        else:
            raise ValueError("Train mode must be one of concurrent, autoregressive, or LSTM")

    # dt the rollout integrates with (the wing trainer uses delta_t_train, train_fixed_wing.py:101-103)
    def rollout_dt(self):
        return self.delta_t

    def modified_params(self):
        """Physical constants of the fused rollout = those of the dynamics object the trainer was GIVEN (the reference
        differentiates through ``self.train_dynamics``, train_drone.py:186-190), not the config's ``modified_params``
        entry (which in the reference's fine-tuning set-ups describes the EVALUATION dynamics)."""
        if isinstance(self.train_dynamics, torch.nn.Module):
            if self.train_mode != "concurrent":
                raise ValueError("learnt dynamics are only supported with train_mode='concurrent' (the recurrent "
                                 "kernels integrate the analytic model)")
            # quadrotor: the fused learnt rollout integrates with the construction-time constants of the learnt object
            # (what its simulator keeps using, quad_dynamics_trained.py:47-48); fixed wing: per-step ops, spec unused
            cfg = getattr(self.train_dynamics, "cfg", None) if self.system == "quad" else None
            return dict(cfg) if isinstance(cfg, dict) else {}
        cfg = getattr(self.train_dynamics, "cfg", None)
        if cfg is not None:
            return dict(cfg)
        return dict(self.config.get("modified_params", {}))

    def init_optimizer(self):
        if self.state_data is not None:
            # pinned batches: the H2D copies of the train step are asynchronous DMA transfers
            on_host = getattr(self.state_data, "device", None) is None        # device-resident datasets: nothing to pin
            self.trainloader = torch.utils.data.DataLoader(self.state_data, batch_size=self.batch_size, shuffle=True,
                                                           num_workers=0,
                                                           pin_memory=torch.cuda.is_available() and on_host)
        spec = T.spec_for_net(self.net, self.system, self.horizon, self.rollout_dt(), self.train_mode, self.window,
                              self.modified_params())
        self.fused = T.ModuleRollout(self.net, spec, self.device)
        self.optimizer_controller = optim.SGD(self.net.parameters(), lr=self.learning_rate_controller, momentum=0.9)

    def init_dynamics_optimizer(self, l2_lambda=0.0):
        """the SGD over the learnt dynamics' parameters of the reference (train_base.py:144-150)"""
        self.l2_lambda = l2_lambda
        self.optimizer_dynamics = optim.SGD(self.train_dynamics.parameters(), lr=self.learning_rate_dynamics,
                                            momentum=0.9)

    def train_dynamics_model(self, current_state, action_seq):
        """One dynamics-fitting step (train_base.py:160-186): squared difference between the learnt step (one fused
        CUDA kernel forward, one adjoint kernel backward) and the evaluation dynamics' step on the first action,
        + l2_lambda * (norms of the residual MLP's tensors)."""
        self.optimizer_dynamics.zero_grad()
        next_d1 = self.train_dynamics(current_state, action_seq[:, 0], dt=self.delta_t)
        with torch.no_grad():
            next_d2 = self.eval_dynamics(current_state, action_seq[:, 0], dt=self.delta_t)
        l2_loss = 0
        if getattr(self, "l2_lambda", 0) > 0:
            d = self.train_dynamics
            l2_loss = (torch.norm(d.linear_state_2.weight) + torch.norm(d.linear_state_2.bias) +
                       torch.norm(d.linear_state_1.weight) + torch.norm(d.linear_state_1.bias))
        loss = torch.sum((next_d1 - next_d2) ** 2) + getattr(self, "l2_lambda", 0) * l2_loss
        loss.backward()
        self.optimizer_dynamics.step()
        self.results_dict["loss_dyn_per_step"].append(loss.item())
        return loss

    def train_controller_model(self, current_state, action_seq, in_ref_state, ref_states):
        """implemented in the sub classes (un-fused path: the caller already evaluated the policy)"""
        raise NotImplementedError

    def fused_train_step(self, in_state, current_state, in_ref_state, ref_states, learnt_params=None):
        """zero_grad -> rollout loss + analytic gradient (two launches) -> optimizer step.  ``learnt_params``: the
        horizon is rolled through the learnt dynamics with these (frozen) parameters."""
        self.optimizer_controller.zero_grad()
        h0c0 = None
        if self.train_mode == "LSTM":
            self.net.reset_hidden_state(current_state.size()[0])
            h0c0 = torch.stack((self.net.hidden_state, self.net.cell_state), 0)
        loss = self.fused.loss_and_grad(None if self.train_mode != "concurrent" else in_state, current_state,
                                        in_ref_state, ref_states, h0c0, learnt_params=learnt_params)
        self.writer.add_scalar("loss/training", loss)
        self.optimizer_controller.step()
        return loss

    def run_epoch(self, train="controller", epoch=0):
        running_loss, i = 0.0, 0
        learnt = isinstance(self.train_dynamics, torch.nn.Module)      # LearntDynamics / LearntFixedWingDynamics
        for i, data in enumerate(self.trainloader, 0):
            in_state, current_state, in_ref_state, ref_states = data
            fused_learnt = (learnt and self.train_mode == "concurrent" and self.system == "quad" and
                            not self.config.get("unfused_learnt_rollout", False) and
                            self.fused.supports_learnt_dynamics(current_state.shape[0]))
            if fused_learnt:
                # the controller is trained THROUGH the learnt dynamics (train_drone.py:262-279) in the fused rollout:
                # the dynamics kernel of the tcgen05 path steps LearntDynamics.forward and its state / action adjoint
                with torch.no_grad():
                    lp = self.train_dynamics._flat().to(self.fused.device)
                loss = self.fused_train_step(in_state, current_state, in_ref_state, ref_states, learnt_params=lp)
            elif learnt and self.train_mode == "concurrent":
                # configurations without the fused variant (fixed wing, other horizons): the reference's own loop,
                # autograd over the per-step CUDA ops (policy forward, learnt step + its adjoint kernel)
                dev = self.fused.device
                in_state, current_state, in_ref_state, ref_states = (x.to(dev) for x in (in_state, current_state,
                                                                                          in_ref_state, ref_states))
                actions = torch.sigmoid(self.net(in_state, in_ref_state))
                action_seq = torch.reshape(actions, (-1, self.horizon, self.action_dim))
                loss = self.train_controller_model(current_state, action_seq, in_ref_state, ref_states)
            else:
                loss = self.fused_train_step(in_state, current_state, in_ref_state, ref_states)
            running_loss += loss.item()
        epoch_loss = running_loss / max(i, 1)      # the reference divides by the last batch index (:213)
        self.results_dict["loss"].append(epoch_loss)
        self.results_dict["trained"].append(train)
        self.writer.add_scalar("Loss/train", epoch_loss, epoch)
        return epoch_loss
