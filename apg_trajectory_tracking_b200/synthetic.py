"""Seeded synthetic inputs of the shapes the reference's datasets hand to the train step.

quad  : degree-5 polynomial reference trajectories, drone at the origin with small attitude / velocity noise,
        featurised exactly like neural_control/dataset.py:155-204 (QuadDataset.prepare_data)
wing  : level flight around 11.5 m/s towards a target 50 m ahead, linear reference of 12 m/s
        (neural_control/dataset.py:309-350, environments/wing_env.py:26-42)
cartpole: uniform states scaled like environments/cartpole_env.py:178-236
Everything is generated on the CPU with a torch.Generator (reproducible across devices) and returned on `device`.
"""
import math

import torch

WING_MEAN = torch.tensor([0.0, 0.0, 0.0, 11.525899887084961, -0.00016766408225521445, 0.16617104411125183,
                          0.007394296582788229, 0.018172707409, 0.020353179425001144, -0.0005361468647606671,
                          0.01662314310669899, 0.004487641621381044])
WING_STD = torch.tensor([16.626325607299805, 0.8449159860610962, 0.8879243731498718, 0.6243225932121277,
                         0.28072822093963623, 0.29176747798, 0.04499124363064766, 0.10370047390460968,
                         0.049977313727, 0.06449887901544571, 0.27508440613746643, 0.05634994804859])


def _u(gen, *shape, lo, hi):
    return torch.rand(*shape, generator=gen) * (hi - lo) + lo


def quad_world_to_body(att):
    r, p, y = att[:, 0], att[:, 1], att[:, 2]
    cy, sy, cp, sp, cr, sr = torch.cos(y), torch.sin(y), torch.cos(p), torch.sin(p), torch.cos(r), torch.sin(r)
    rows = [torch.stack((cy * cp, sy * cp, -sp), 1),
            torch.stack((cy * sp * sr - cr * sy, cr * cy + sr * sy * sp, cp * sr), 1),
            torch.stack((cy * sp * cr + sr * sy, cr * sy * sp - cy * sr, cr * cp), 1)]
    return torch.stack(rows, 1)


def quad_features(cur):
    """[vel, W00 W01 W10 W11 W20 W21, W.vel, body rates] (dataset.py:207-220)."""
    w = quad_world_to_body(cur[:, 3:6])
    vel = cur[:, 6:9]
    vb = torch.einsum("nij,nj->ni", w, vel)
    return torch.cat((vel, w[:, :, :2].reshape(-1, 6), vb, cur[:, 9:12]), 1)


def quad_case(n, ref_rows, dt=0.1, seed=1234, device="cpu"):
    """returns dict(cur (n,12), ref (n,ref_rows,9), in_ref (n,ref_rows,9), in_state (n,15))."""
    g = torch.Generator().manual_seed(seed)
    c = torch.zeros(n, 3, 6)
    c[:, :, 1] = _u(g, n, 3, lo=-1.5, hi=1.5)
    for i in range(2, 6):
        c[:, :, i] = _u(g, n, 3, lo=-0.5, hi=0.5) / math.factorial(i)
    t = (torch.arange(ref_rows, dtype=torch.float32) + 1) * dt
    pw = torch.stack([t ** i for i in range(6)])
    dpw = torch.stack([torch.zeros_like(t) if i == 0 else i * t ** (i - 1) for i in range(6)])
    ref = torch.zeros(n, ref_rows, 9)
    ref[:, :, 0:3] = torch.einsum("nai,il->nla", c, pw)
    ref[:, :, 6:9] = torch.einsum("nai,il->nla", c, dpw)
    cur = torch.zeros(n, 12)
    cur[:, 3:6] = _u(g, n, 3, lo=-0.2, hi=0.2)
    cur[:, 6:9] = c[:, :, 1] + 0.3 * torch.randn(n, 3, generator=g)
    in_ref = torch.cat((ref[..., 0:3], ref[..., 6:9], ref[..., 6:9] - cur[:, None, 6:9]), 2)
    out = dict(cur=cur, ref=ref, in_ref=in_ref, in_state=quad_features(cur))
    return {k: v.contiguous().to(device) for k, v in out.items()}


def wing_case(n, horizon, dt=0.05, seed=1234, device="cpu"):
    """returns dict(cur (n,12), ref (n,h,3), in_ref (n,3), in_state (n,9), target (n,3) = the raw reference sample)."""
    g = torch.Generator().manual_seed(seed)
    cur = torch.zeros(n, 12)
    cur[:, 3] = 11.5 + _u(g, n, lo=-0.5, hi=0.5)
    cur[:, 5] = _u(g, n, lo=-0.5, hi=0.5)
    cur[:, 7] = _u(g, n, lo=-2, hi=2) * math.pi / 180
    cur[:, 10] = _u(g, n, lo=-0.005, hi=0.005)
    target = torch.stack((torch.full((n,), 50.0), _u(g, n, lo=-5, hi=5), _u(g, n, lo=-5, hi=5)), 1)
    rel = target - cur[:, :3]
    unit = rel / rel.norm(dim=1, keepdim=True)
    steps = (torch.arange(horizon, dtype=torch.float32) + 1)[None, :, None]
    ref = cur[:, None, :3] + unit[:, None, :] * (12 * dt) * steps
    in_ref = ref[:, -1] - cur[:, :3]
    in_state = ((cur - WING_MEAN) / WING_STD)[:, 3:]
    out = dict(cur=cur, ref=ref, in_ref=in_ref, in_state=in_state, target=target)
    return {k: v.contiguous().to(device) for k, v in out.items()}


def cartpole_case(n, seed=1234, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    st = (torch.rand(n, 4, generator=g) * 2 - 1) * torch.tensor([2.4, 7.5, math.pi, 7.5])
    st[:, 1] *= 0.2
    st[:, 3] *= 0.2
    return dict(cur=st.contiguous().to(device), in_state=st.clone().contiguous().to(device))
