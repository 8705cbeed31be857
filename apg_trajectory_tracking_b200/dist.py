"""Multi-GPU plumbing: one process per GPU, drones sharded along the batch axis, ONE sum-allreduce of the flat
gradient per train step (the loss is a sum over drones, so the summed gradient equals the single-device gradient of
the concatenated batch; no mean scaling).  NCCL on GPUs, gloo in the CPU tests."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _capi


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_bounds(n_total, rank, world):
    """contiguous split of the drone axis: rank r owns [lo, hi)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(tensor, rank, world):
    if tensor is None:
        return None
    lo, hi = shard_bounds(tensor.shape[0], rank, world)
    return tensor[lo:hi]


def allreduce_sum_(flat_grad, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def sgd_momentum_step_(flat, grad, buf, lr, momentum=0.9):
    """identical update on every rank after the allreduce (optim.SGD semantics, train_base.py:139-143)"""
    buf.mul_(momentum).add_(grad)
    flat.add_(buf, alpha=-lr)


class PeerGradExchange:
    """The gradient all-reduce as this package's own kernels over NVLink peer memory (csrc/p2p_kernels.cu; protocol in
    csrc/p2p_math.cuh): every rank's reduction kernel stores its gradient straight into a slot on every peer and
    raises a flag there; a second kernel waits for the flags, sums the slots in rank order (bitwise identical on all
    ranks) and can apply the SGD-momentum update in the same pass.  OPTIONAL (``FusedTrainStep(..., peer_exchange=True)``
    or ``APG_P2P_GRAD=1``; default: NCCL all-reduce) - it needs all ranks on one NVLink domain and
    ``torch.distributed._symmetric_memory`` for the peer mappings (plumbing only: allocation + pointer exchange).

    One instance per (parameter count, process group); ``next_step()`` hands out the descriptor of the next step -
    every rank must call it the same number of times."""

    def __init__(self, n_params, device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized()):
            raise _capi.ApgError("PeerGradExchange needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.device = int(n_params), torch.device(device)
        self.lib = _capi.lib()
        nbytes = self.lib.apg_grad_comm_bytes(self.world, self.n)
        if nbytes == 0:
            raise _capi.ApgError("bad world size / parameter count for the peer gradient exchange")
        with torch.cuda.device(self.device):
            self.buf = symm.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            self.buf.zero_()
            self.handle = symm.rendezvous(self.buf, self.group)
            peers = [int(p) for p in self.handle.buffer_ptrs]
            if len(peers) != self.world:
                raise _capi.ApgError("symmetric memory rendezvous returned %d peer mappings for %d ranks"
                                     % (len(peers), self.world))
            self._local, slot_tab, flag_tab = [], [], []
            for s in (0, 1):
                so, fo = ctypes.c_size_t(), ctypes.c_size_t()
                _capi.check(self.lib.apg_grad_comm_offsets(self.world, self.n, s, ctypes.byref(so), ctypes.byref(fo)))
                slot_tab.append([p + so.value for p in peers])
                flag_tab.append([p + fo.value for p in peers])
                self._local.append(peers[self.rank] + so.value)
            self.slot_tab = torch.tensor(slot_tab, dtype=torch.int64, device=self.device)
            self.flag_tab = torch.tensor(flag_tab, dtype=torch.int64, device=self.device)
            self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.epoch = 0
            torch.cuda.synchronize(self.device)
        dist.barrier(self.group)              # every rank's sets are zeroed before the first peer store can arrive

    def next_step(self):
        """-> (apg_grad_comm descriptor, address of this rank's receive set) of the next step"""
        self.epoch += 1
        s = self.epoch & 1
        comm = _capi.ApgGradComm(self.rank, self.world, ctypes.c_void_p(self.slot_tab[s].data_ptr()),
                                 ctypes.c_void_p(self.flag_tab[s].data_ptr()), self.epoch & 0xffffffff,
                                 ctypes.c_void_p(self.ticket.data_ptr()))
        return comm, ctypes.c_void_p(self._local[s])

    def gather(self, comm, local_set, grad_out=None, params=None, momentum_buf=None, lr=0.0, momentum=0.0):
        """second half of the step: wait for all ranks' slots, sum them in rank order into ``grad_out``; with
        ``params`` / ``momentum_buf`` also ``buf = momentum*buf + g; params -= lr*buf`` in the same kernel"""
        def ptr(t):
            return None if t is None else ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _capi.check(self.lib.apg_grad_gather_sgd_p2p(ctypes.byref(comm), local_set, self.n, ptr(grad_out),
                                                         ptr(params), ptr(momentum_buf), ctypes.c_float(float(lr)),
                                                         ctypes.c_float(float(momentum)), st))
