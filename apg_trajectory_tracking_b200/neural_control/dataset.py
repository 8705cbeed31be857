"""Data formats on the input side of the rollout (reference: ``neural_control/dataset.py``).

* ``state_preprocessing`` (:207-220): the per-step featurizer of the recurrent train modes, (N,12) quadrotor state ->
  (N,15) policy features, differentiable, one CUDA kernel forward / one backward (csrc/apg_math.cuh ``Quad::features``).
* ``QuadDataset`` / ``WingDataset`` / ``CartpoleDataset``: the HOST-side containers that turn raw (state, reference)
  samples into the four tensors of a train batch -- ``(in_state, current_state, in_ref_state, ref_states)`` -- with the
  reference's layouts (``prepare_data`` :155-204 and :326-350).  Like in the reference they live in CPU memory by
  default (data preparation, not rollout math; the train step copies each batch to the GPU); with ``device=`` a CUDA
  device their ``prepare_data`` is the device kernels of ``prepare.py`` and the prepared tensors stay in HBM.

Two constructor forms, told apart by the first argument:
  * the REFERENCE's (``dataset.py:46-73, 135-153, 223-240, 261-283``): ``QuadDataset(num_states, self_play, mean=None,
    std=None, **kwargs)``, ``WingDataset(num_states, self_play=0, mean=None, std=None, ref_mean=None, ref_std=None,
    delta_t=0.05, horizon=10, **kwargs)``, ``CartpoleDataset(num_states=1000, thresh_div=.21, dt=0.05, **kwargs)``:
    the samples come from ``sample_data`` -> ``full_state_training_data`` / ``sample_training_data`` /
    ``construct_states`` of the mirrored environments (device kernels: table layout, window cutting, dynamics steps),
    with ``resample_data`` and ``get_and_add_eval_data`` as in the reference - so ``scripts/train_*.py`` of the
    reference construct them unchanged;
  * raw samples: ``QuadDataset(states, ref_states, ...)`` (arrays), used by the batched device pipeline and the tests.
"""
import numpy as np
import torch

from ..ops import quad_features
from .. import prepare as _prep, synthetic as _syn


def state_preprocessing(drone_states):
    return quad_features(drone_states)


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x.detach().float().cpu().clone()
    return torch.from_numpy(np.array(x, dtype=np.float64)).float()


def _is_count(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


class DroneDataset(torch.utils.data.Dataset):
    """common container: holds the prepared tensors, hands out 4-tuples, supports replacing samples (self play)"""

    def __init__(self, states, ref_states=None, mean=None, std=None, self_play=0, device=None, **kwargs):
        """``device``: None (default) keeps the container and its ``prepare_data`` on the host like the reference;
        a CUDA device makes ``prepare_data`` the device kernels of ``prepare.py`` and keeps the four
        prepared tensors in HBM, so that batches need no host->device copy."""
        self.device = None if device is None else torch.device(device)
        self.kwargs = kwargs
        if _is_count(states):
            # the reference's constructor (dataset.py:46-73): (num_states, self_play, mean, std, **kwargs); the second
            # positional argument is the self-play FRACTION there
            num_states = int(states)
            frac = ref_states if ref_states is not None else self_play
            self.num_sampled_states = num_states
            self.num_self_play = int(frac * num_states)
            self.total_dataset_size = self.num_sampled_states + self.num_self_play
            states, ref_states = self.sample_data(self.total_dataset_size)
            states_np = np.asarray(states, dtype=np.float64)
        else:
            states_np = np.asarray(states, dtype=np.float64)
            self.num_sampled_states = int(len(states_np) / (1 + self_play)) if self_play else len(states_np)
            self.num_self_play = len(states_np) - self.num_sampled_states
            self.total_dataset_size = len(states_np)
        if mean is None:
            mean, std = states_np.mean(axis=0), states_np.std(axis=0)
        self.mean = torch.as_tensor(np.asarray(mean)).float()
        self.std = torch.as_tensor(np.asarray(std)).float()
        self.normed_states, self.states, self.in_ref_states, self.ref_states = self.prepare_data(states, ref_states)
        self.eval_counter = 0

    def sample_data(self, num_states):
        raise NotImplementedError

    def resample_data(self):
        """new samples into the sampled part of the dataset (dataset.py:89-102)"""
        states, ref_states = self.sample_data(self.num_sampled_states)
        self.replace_sampled(states, ref_states)

    def prepare_data(self, states, ref_states):
        raise NotImplementedError

    def get_means_stds(self, param_dict):
        param_dict["mean"], param_dict["std"] = self.mean.tolist(), self.std.tolist()
        return param_dict

    def get_eval_index(self):
        if self.num_self_play > 0:
            return (self.eval_counter % self.num_self_play) + self.num_sampled_states

    def replace_sampled(self, states, ref_states):
        """what the reference's ``resample_data`` does once new raw samples exist: overwrite the sampled part"""
        prep = self.prepare_data(states, ref_states)
        n = min(self.num_sampled_states, len(prep[0]))
        for dst, src in zip((self.normed_states, self.states, self.in_ref_states, self.ref_states), prep):
            dst[:n] = src[:n]

    def get_and_add_eval_data(self, states, ref_states, add_to_dataset=False):
        prep = self.prepare_data(states, ref_states)
        if add_to_dataset and self.num_self_play > 0:
            at = self.get_eval_index()
            for dst, src in zip((self.normed_states, self.states, self.in_ref_states, self.ref_states), prep):
                dst[at] = src[0]
            self.eval_counter += 1
        return prep

    def to_torch(self, states):
        return _as_tensor(states)

    def __len__(self):
        return len(self.states)

    def __getitem__(self, index):
        return self.normed_states[index], self.states[index], self.in_ref_states[index], self.ref_states[index]


class QuadDataset(DroneDataset):
    """quadrotor batches: reference positions relative to the drone, drone position zeroed, features =
    [world vel, first two columns of the world->body matrix, body vel, body rates], reference input =
    [rel. position, velocity, velocity - drone velocity] per horizon row"""

    def sample_data(self, num_states):
        """dataset.py:140-144: windows cut from random trajectory tables"""
        from .environments.drone_env import full_state_training_data
        return full_state_training_data(num_states, **self.kwargs)

    @staticmethod
    def rot_world_to_body(state_vector, world_to_body):
        return torch.matmul(world_to_body, state_vector.unsqueeze(2))[:, :, 0]

    def prepare_data(self, states, ref_states):
        cur, ref = _as_tensor(states), _as_tensor(ref_states)
        if cur.dim() == 1:
            cur, ref = cur[None], ref[None]
        if getattr(self, "device", None) is not None:
            out = _prep.prepare_quad(cur.to(self.device), ref.to(self.device))
            return out["in_state"], out["cur"], out["in_ref"], out["ref"]
        ref[:, :, :3] -= cur[:, None, :3]
        cur[:, :3] = 0
        drone_vel = cur[:, None, 6:9]
        in_ref = torch.cat((ref[:, :, :3], ref[:, :, 6:9], ref[:, :, 6:9] - drone_vel), dim=2)
        return _syn.quad_features(cur), cur, in_ref, ref


class WingDataset(DroneDataset):
    """fixed-wing batches: normalised state without position, unit vector towards the target scaled to the last point
    of a 12 m/s straight-line reference, the straight-line reference itself for the loss"""

    def __init__(self, states, ref_states=None, mean=None, std=None, ref_mean=None, ref_std=None, delta_t=0.05,
                 horizon=10, self_play=0, **kwargs):
        """reference form (dataset.py:263-283): ``WingDataset(num_states, self_play=0, mean=None, std=None, ref_mean=
        None, ref_std=None, delta_t=0.05, horizon=10, **kwargs)``; raw form: ``WingDataset(states, targets, ...)``"""
        self.dt, self.horizon = delta_t, horizon
        if _is_count(states):
            frac = ref_states if ref_states is not None else self_play
            if mean is None:
                super().__init__(states, frac, mean=_syn.WING_MEAN.numpy(), std=_syn.WING_STD.numpy(), **kwargs)
                self.set_fixed_mean()                           # dataset.py:280-283: fixed statistics
            else:
                super().__init__(states, frac, mean=mean, std=std, **kwargs)
            return
        if mean is None:
            mean, std = _syn.WING_MEAN.numpy(), _syn.WING_STD.numpy()       # fixed statistics (dataset.py:284-300)
        super().__init__(states, ref_states, mean=mean, std=std, self_play=self_play, **kwargs)

    def set_fixed_mean(self):
        self.mean, self.std = _syn.WING_MEAN.clone().float(), _syn.WING_STD.clone().float()

    def sample_data(self, num_samples):
        """dataset.py:302-307"""
        from .environments.wing_env import sample_training_data
        if num_samples == 0:                                    # the trainer starts from an empty dataset + self play
            return np.zeros((0, 12)), np.zeros((0, 3))
        return sample_training_data(num_samples, **self.kwargs)

    def _compute_target_pos(self, current_state, ref_vector):
        steps = (torch.arange(self.horizon, dtype=torch.float32) + 1)[None, :, None]
        return current_state[:, None, :3] + ref_vector[:, None, :] * (12 * self.dt) * steps

    def prepare_data(self, states, ref_states):
        cur, target = _as_tensor(states), _as_tensor(ref_states)
        if cur.dim() == 1:
            cur, target = cur[None], target[None]
        if getattr(self, "device", None) is not None:
            out = _prep.prepare_wing(cur.to(self.device), target.to(self.device), self.mean, self.std, self.dt,
                                     self.horizon)
            return out["in_state"], out["cur"], out["in_ref"], out["ref"]
        normed = ((cur - self.mean) / self.std)[:, 3:]
        rel = target - cur[:, :3]
        unit = rel / rel.norm(dim=1, keepdim=True)
        lin_ref = self._compute_target_pos(cur, unit)
        return normed, cur, lin_ref[:, -1] - cur[:, :3], lin_ref


class CartpoleDataset(torch.utils.data.Dataset):
    """cartpole batches are (state, state): the policy input is the raw state (dataset.py:223-258)"""

    def __init__(self, states=1000, thresh_div=.21, dt=0.05, **kwargs):
        """reference form (dataset.py:228-230): ``CartpoleDataset(num_states=1000, thresh_div=.21, dt=0.05)``;
        raw form: ``CartpoleDataset(states)``"""
        self.dt = dt
        if _is_count(states):
            self.resample_data(int(states), thresh_div)
        else:
            self.labels = _as_tensor(states)
            self.states = self.labels.clone()

    def resample_data(self, num_states, thresh_div):
        """dataset.py:232-240"""
        from .environments.cartpole_env import construct_states
        self.labels = _as_tensor(construct_states(num_states, self.dt, thresh_div=thresh_div))
        self.states = self.labels.clone()

    def to_torch(self, states):
        return _as_tensor(states)

    def __len__(self):
        return len(self.states)

    def __getitem__(self, index):
        return self.states[index], self.labels[index]

    def add_data(self, new_numpy_data):
        new = _as_tensor(new_numpy_data)
        n = min(len(new), len(self))
        self.labels[:n] = new[:n]
        self.states[:n] = new[:n]
