"""Data formats on the input side of the rollout, produced ON THE DEVICE.

The reference prepares its train batches on the host (``QuadDataset.prepare_data`` neural_control/dataset.py:155-204,
``WingDataset.prepare_data`` :309-350, ``full_state_training_data`` environments/drone_env.py:232-269).  Here the same
layouts come out of element-wise CUDA kernels (csrc/prep_kernels.cu) so that a train step only needs the raw samples:
for the quadrotor the host->device traffic of a batch drops from 828 B to 408 B per drone.  CUDA tensors only.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from .ops import _p, _require_cuda, _stream


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()


def _f64c(t):
    t = torch.as_tensor(t)
    return t if (t.dtype == torch.float64 and t.is_contiguous()) else t.contiguous().double()


def _opt(t):
    return None if t is None else _p(t)


def prepare_quad(states, ref_states, want=("in_state", "cur", "in_ref", "ref"), out=None, in_place=False):
    """QuadDataset.prepare_data on the device.

    states (N,12), ref_states (N,L,9) raw samples -> dict with the requested tensors of
    ``in_state`` (N,15), ``cur`` (N,12; position zeroed), ``in_ref`` (N,L,9), ``ref`` (N,L,9; positions relative).
    ``in_place=True`` writes ``cur`` / ``ref`` into the inputs (like the reference, dataset.py:170-174);
    ``out`` may carry preallocated result tensors."""
    _require_cuda(states, ref_states)
    s, r = _f32c(states), _f32c(ref_states)
    n, L = s.shape[0], r.shape[1]
    if s.shape[1] != 12 or r.shape[0] != n or r.shape[2] != 9:
        raise ValueError(f"prepare_quad: expected states (N,12) and ref_states (N,L,9), got {tuple(s.shape)} and "
                         f"{tuple(r.shape)}")
    out = dict(out or {})
    shapes = {"in_state": (n, 15), "cur": (n, 12), "in_ref": (n, L, 9), "ref": (n, L, 9)}
    for k in want:
        if k not in shapes:
            raise ValueError(f"prepare_quad: unknown output {k!r}")
        if k not in out:
            if in_place and k == "cur":
                out[k] = s
            elif in_place and k == "ref":
                out[k] = r
            else:
                out[k] = torch.empty(shapes[k], dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _capi.check(_capi.lib().apg_prepare_quad(_p(s), _p(r), n, L, _opt(out.get("in_state")), _opt(out.get("cur")),
                                                 _opt(out.get("in_ref")), _opt(out.get("ref")), _stream(s)))
    return out


def prepare_wing(states, targets, mean, std, dt, horizon, want=("in_state", "cur", "in_ref", "ref"), out=None):
    """WingDataset.prepare_data on the device: states (N,12), targets (N,3) ->
    ``in_state`` (N,9), ``cur`` (N,12), ``in_ref`` (N,3), ``ref`` (N,h,3)."""
    _require_cuda(states, targets)
    s, tg = _f32c(states), _f32c(targets)
    n = s.shape[0]
    if s.shape[1] != 12 or tuple(tg.shape) != (n, 3):
        raise ValueError(f"prepare_wing: expected states (N,12) and targets (N,3), got {tuple(s.shape)} and "
                         f"{tuple(tg.shape)}")
    mean_h = np.ascontiguousarray(torch.as_tensor(mean).detach().cpu().numpy(), dtype=np.float32)
    std_h = np.ascontiguousarray(torch.as_tensor(std).detach().cpu().numpy(), dtype=np.float32)
    if mean_h.shape != (12,) or std_h.shape != (12,):
        raise ValueError("prepare_wing: mean / std must have 12 entries")
    out = dict(out or {})
    shapes = {"in_state": (n, 9), "cur": (n, 12), "in_ref": (n, 3), "ref": (n, int(horizon), 3)}
    for k in want:
        if k not in shapes:
            raise ValueError(f"prepare_wing: unknown output {k!r}")
        if k not in out:
            out[k] = s if k == "cur" else torch.empty(shapes[k], dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _capi.check(_capi.lib().apg_prepare_wing(
            _p(s), _p(tg), ctypes.c_void_p(mean_h.ctypes.data), ctypes.c_void_p(std_h.ctypes.data),
            ctypes.c_float(float(dt)), int(horizon), n, _opt(out.get("in_state")), _opt(out.get("cur")),
            _opt(out.get("in_ref")), _opt(out.get("ref")), _stream(s)))
    return out


def sample_windows(traj, n, ref_rows, stride):
    """``full_state_training_data`` for one trajectory table (T, W>=9) on the device: sample i starts at row
    i*stride; returns states (n,12) = [traj row, 0 0 0] and ref_states (n, ref_rows, 9) = the following rows."""
    _require_cuda(traj)
    t = _f32c(traj)
    states = torch.empty(n, 12, dtype=torch.float32, device=t.device)
    refs = torch.empty(n, ref_rows, 9, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _capi.check(_capi.lib().apg_sample_windows(_p(t), t.shape[0], t.shape[1], int(ref_rows), int(stride), int(n),
                                                   _p(states), _p(refs), _stream(t)))
    return states, refs


def poly_reference(coef, rows, dt, t_first=None):
    """coef (N,3,6): per-axis polynomial coefficients c0..c5 -> (N, rows, 9) reference rows [p(t), 0 0 0, p'(t)] at
    t = t_first + k*dt (default t_first = dt, i.e. the first row is one step ahead of the drone)."""
    _require_cuda(coef)
    c = _f32c(coef)
    if c.dim() != 3 or c.shape[1] != 3 or c.shape[2] != 6:
        raise ValueError(f"poly_reference: expected coef (N,3,6), got {tuple(c.shape)}")
    out = torch.empty(c.shape[0], int(rows), 9, dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _capi.check(_capi.lib().apg_poly_reference(_p(c), c.shape[0], int(rows),
                                                   ctypes.c_float(float(dt if t_first is None else t_first)),
                                                   ctypes.c_float(float(dt)), _p(out), _stream(c)))
    return out


def reference_table(traj, dt, speed_factor, z_offset=3.0):
    """``load_prepare_trajectory`` (neural_control/trajectory/generate_trajectory.py:566-603) on the device, for an
    already loaded raw trajectory (T, W>=10) = [pos, quaternion wxyz, vel, ...] sampled every 0.01 s: every
    ``int(dt / 0.01 * speed_factor)``-th row -> (rows, 9) = [pos, euler * speed_factor, vel * speed_factor * 2], with
    the ``+3`` on z that ``Random.__init__`` applies (random_traj.py:35; pass ``z_offset=0`` for the bare table)."""
    _require_cuda(traj)
    t = _f32c(traj)
    if t.dim() != 2 or t.shape[1] < 10:
        raise ValueError(f"reference_table: expected a raw trajectory (T, W>=10), got {tuple(t.shape)}")
    nth = int(dt / 0.01 * speed_factor)
    if nth < 1 or not np.isclose(nth, dt / 0.01 * speed_factor):         # the reference asserts this too (:585)
        raise ValueError("reference_table: dt / 0.01 * speed_factor must be a positive integer")
    rows = (t.shape[0] + nth - 1) // nth                                  # len(traj[::nth])
    out = torch.empty(rows, 9, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _capi.check(_capi.lib().apg_reference_table(_p(t), t.shape[0], t.shape[1], nth,
                                                    ctypes.c_float(float(speed_factor)),
                                                    ctypes.c_float(float(z_offset)), rows, _p(out), _stream(t)))
    return out


def polynomial_points(coef, rot, start=None, x_range=20.0, max_drone_dist=0.25, horizon=10, hover_steps=50,
                      x_start=1.0, max_rows=None, check=True):
    """The reference rows of ``Polynomial`` (neural_control/trajectory/polynomial.py:8-125, ``random_polynomial`` branch)
    for N trajectories on the device: coef (N, degree+1) as ``np.polyfit`` returns them, rot (N,3,3) rotations,
    start (N,3) drone positions (or None) -> (points (N,max_rows,3), ref_len (N,) int32); trajectory i is
    ``points[i, :ref_len[i]]`` (hover copies, marched points, hover copies).  ``check`` synchronises and raises when
    ``max_rows`` (default: hover padding + 4x the straight-line number of steps) was too small for a trajectory."""
    _require_cuda(coef, rot, start)
    c, r = _f64c(coef), _f64c(rot)                 # double like the numpy fit: x^5 at x ~ 20 amplifies float32 rounding
    n = c.shape[0]
    if c.dim() != 2 or tuple(r.shape) != (n, 3, 3) or not 2 <= c.shape[1] <= 12:
        raise ValueError(f"polynomial_points: expected coef (N,degree+1) and rot (N,3,3), got {tuple(c.shape)} and "
                         f"{tuple(r.shape)}")
    st = None
    if start is not None:
        st = _f64c(start)
        if tuple(st.shape) != (n, 3):
            raise ValueError("polynomial_points: start must be (N,3)")
    dist = float(max_drone_dist) / int(horizon)
    if max_rows is None:
        max_rows = 2 * int(hover_steps) + 4 * int(float(x_range) / dist) + 8
    pts = torch.zeros(n, int(max_rows), 3, dtype=torch.float32, device=c.device)
    ref_len = torch.zeros(n, dtype=torch.int32, device=c.device)
    with torch.cuda.device(c.device):
        _capi.check(_capi.lib().apg_polynomial_points(_p(c), c.shape[1] - 1, _p(r), _opt(st), n,
                                                      ctypes.c_double(float(x_start)), ctypes.c_double(float(x_range)),
                                                      ctypes.c_double(dist), int(hover_steps), int(max_rows), _p(pts),
                                                      _p(ref_len), _stream(c)))
    if check and n and int(ref_len.max()) > max_rows:
        raise ValueError(f"polynomial_points: a trajectory needs {int(ref_len.max())} rows, max_rows = {max_rows}")
    return pts, ref_len
