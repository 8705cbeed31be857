"""Device-resident training data (B200: 180 GB of HBM holds far more raw samples than any run of the reference uses).

The reference keeps its dataset in host tensors, re-prepares it on the CPU and hands mini-batches to the train step
through a ``DataLoader`` (scripts/train_base.py:130-137, 188-218; neural_control/dataset.py).  With the rollout itself
at ~0.4 ms per 65k-drone batch that host path (prepare + H2D) is what an epoch would spend its time on.  Here the RAW
samples live in HBM, a shuffled mini-batch is one row gather, and the policy inputs are derived by the prepare kernels
(prepare.py) right before the fused rollout: nothing crosses PCIe during an epoch.

``DeviceQuadDataset.from_trajectory`` cuts the samples out of a trajectory table on the device
(``full_state_training_data`` layout, environments/drone_env.py:232-269), ``from_polynomials`` evaluates polynomial
references on the device."""
import math

import torch

from . import _capi, prepare as PR


class DeviceQuadDataset:
    """raw quadrotor samples on the GPU: ``states`` (N,12), ``ref_states`` (N,L,9) rows [pos, euler, vel]"""

    def __init__(self, states, ref_states, device=None, num_self_play=0):
        """``num_self_play``: the last ``num_self_play`` rows are the ring of self-play slots ``add_self_play``
        overwrites (``DroneDataset``: ``int(self_play * num_sampled_states)`` rows after the sampled ones,
        dataset.py:40-58); until then they hold ordinary samples."""
        if not torch.cuda.is_available():
            raise _capi.ApgError("DeviceQuadDataset needs a CUDA device")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.states = torch.as_tensor(states).to(dev, torch.float32).contiguous()
        self.ref_states = torch.as_tensor(ref_states).to(dev, torch.float32).contiguous()
        if self.states.shape[0] != self.ref_states.shape[0] or self.states.shape[1] != 12 or \
                self.ref_states.shape[2] != 9:
            raise ValueError("expected states (N,12) and ref_states (N,L,9)")
        self.device = dev
        if not 0 <= int(num_self_play) < self.states.shape[0]:
            raise ValueError("num_self_play must leave at least one sampled row")
        self.num_self_play = int(num_self_play)
        self.num_sampled_states = self.states.shape[0] - self.num_self_play
        self.eval_counter = 0

    def get_eval_index(self):
        """slot the next self-play sample goes to (``DroneDataset.get_eval_index``, dataset.py:78-85)"""
        if self.num_self_play > 0:
            return self.eval_counter % self.num_self_play + self.num_sampled_states
        return None

    def add_self_play(self, states, ref_states):
        """write M raw samples (``evaluate.selfplay_samples``) into the ring of self-play slots, one after the other
        like M calls of ``get_and_add_eval_data(..., add_to_dataset=True)`` (dataset.py:103-119)"""
        m = states.shape[0]
        if self.num_self_play == 0 or m == 0:
            return
        if ref_states.shape[1] < self.ref_states.shape[1] or ref_states.shape[0] != m:
            raise ValueError("self-play reference rows do not fit the dataset's")
        keep = min(m, self.num_self_play)                            # older ones would be overwritten anyway
        j = torch.arange(m - keep, m, device=self.device)
        slots = (self.eval_counter + j) % self.num_self_play + self.num_sampled_states
        self.states.index_copy_(0, slots, states[m - keep:].to(self.device, torch.float32))
        self.ref_states.index_copy_(0, slots, ref_states[m - keep:, :self.ref_states.shape[1]]
                                    .to(self.device, torch.float32))
        self.eval_counter += m

    @classmethod
    def from_trajectory(cls, traj, ref_length, sample_freq=None, device=None, num_self_play=0):
        """samples cut from one trajectory table (T, W>=9) like full_state_training_data: every ``sample_freq``-th row
        is a drone state, the following ``ref_length`` rows its reference"""
        traj = torch.as_tensor(traj)
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        traj = traj.to(dev, torch.float32).contiguous()
        sample_freq = sample_freq or 2 * ref_length
        n = len(range(0, traj.shape[0] - (ref_length + 1), sample_freq))
        states, refs = PR.sample_windows(traj, n, ref_length, sample_freq)
        return cls(states, refs, dev, num_self_play)

    @classmethod
    def from_polynomials(cls, n, ref_length, dt, seed=0, device=None, num_self_play=0):
        """the bench's synthetic samples: degree-5 polynomial references, drone near the start of its reference"""
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        g = torch.Generator(device=dev).manual_seed(seed)
        c = torch.zeros(n, 3, 6, device=dev)
        c[:, :, 1] = torch.rand(n, 3, device=dev, generator=g) * 3 - 1.5
        for i in range(2, 6):
            c[:, :, i] = (torch.rand(n, 3, device=dev, generator=g) - 0.5) / math.factorial(i)
        refs = PR.poly_reference(c, ref_length, dt)
        states = torch.zeros(n, 12, device=dev)
        states[:, 3:6] = torch.rand(n, 3, device=dev, generator=g) * 0.4 - 0.2
        states[:, 6:9] = c[:, :, 1] + 0.3 * torch.randn(n, 3, device=dev, generator=g)
        return cls(states, refs, dev, num_self_play)

    def __len__(self):
        return self.states.shape[0]

    def batches(self, batch_size, shuffle=True, generator=None, drop_last=False):
        """yields (states, ref_states) mini-batches as device tensors (a gather when shuffled, views otherwise)"""
        n = len(self)
        order = torch.randperm(n, device=self.device, generator=generator) if shuffle else None
        for lo in range(0, n, batch_size):
            hi = min(lo + batch_size, n)
            if drop_last and hi - lo < batch_size:
                return
            if order is None:
                yield self.states[lo:hi], self.ref_states[lo:hi]
            else:
                idx = order[lo:hi]
                yield self.states.index_select(0, idx), self.ref_states.index_select(0, idx)


def run_epoch_device(module_rollout, optimizer, dataset, batch_size, shuffle=True, generator=None):
    """``TrainBase.run_epoch`` (train_base.py:188-218) over a device-resident dataset: per mini-batch one gather, the
    prepare kernels, ONE fused rollout (forward + adjoint) and the optimizer step.  Returns the mean batch loss with
    the reference's divisor (the last batch index, :213)."""
    total, i = torch.zeros((), device=dataset.device), 0
    for i, (states, refs) in enumerate(dataset.batches(batch_size, shuffle, generator)):
        optimizer.zero_grad()
        loss = module_rollout.loss_and_grad(None, states, None, refs)
        optimizer.step()
        total += loss
    return float(total) / max(i, 1)
