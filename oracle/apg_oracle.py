"""CPU oracle for the batched differentiable rollout (policy -> dynamics x horizon -> loss -> gradient).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it,
and there only as the checker / the timed CPU arm -- never as a code path of the product.

It is an independent closed-form restatement, in plain PyTorch (CPU, fp32 or fp64), of the reference
functions on the hot path.  Every function cites the reference file:line it follows
(paths relative to the reference checkout).  Gradients come from ``torch.autograd`` on these restated ops.

Pinning: ``oracle/make_golden.py`` runs the *unmodified reference code* (imported from /root/reference with
stubbed optional deps) on the known-answer inputs KAT-1..6 of SURVEY.md section 8c plus seeded random cases
and freezes its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
those vectors.  The reference's autoregressive / LSTM train steps cannot run ``backward()`` (in-place
aliasing at scripts/train_drone.py:138-142), so for those two modes only the *forward* (loss, actions,
states) is pinned against the reference; the gradient oracle there is autograd on the functional
restatement below whose forward is pinned ("cumulative" window semantics).
"""
import math

import torch

# --------------------------------------------------------------------------------------------
# physical constants (neural_control/dynamics/config_quad.json, config_fixed_wing.json,
# config_cartpole.json + overrides at cartpole_dynamics.py:34-36)
# --------------------------------------------------------------------------------------------
QUAD_CFG = dict(
    mass=0.723, arm_length=0.31, frame_inertia=(4.5, 4.5, 7.0), kinv_ang_vel_tau=(16.6, 16.6, 5.0),
    gravity=(0.0, 0.0, -9.81), rotational_drag=(0.0, 0.0, 0.0), translational_drag=(0.0, 0.0, 0.0),
)
WING_CFG = dict(
    mass=1.01, I_xx=0.04766, I_yy=0.05005, I_zz=0.09558, I_xz=-0.00105, rho=1.225, S=0.276, c=0.185,
    b=1.54, g=9.81, CL0=0.39, CL_alpha=4.5321, CL_q=0.318, CL_del_e=0.527, CD0=0.0765, CD_alpha=0.3346,
    CD_q=0.354, CD_del_e=0.004, CY0=0.0, CY_beta=-0.033, CY_p=-0.1, CY_r=0.039, CY_del_a=0.0,
    CY_del_r=0.225, Cl0=0.0, Cl_beta=-0.081, Cl_p=-0.529, Cl_r=0.159, Cl_del_a=-0.453, Cl_del_r=0.005,
    Cm0=0.02, Cm_alpha=-1.4037, Cm_q=-0.1324, Cm_del_e=-0.4236, Cn0=0.0, Cn_beta=0.189, Cn_p=-0.083,
    Cn_r=-0.948, Cn_del_a=-0.041, Cn_del_r=-0.077, epsilon=0.16534698176788384,
)
CARTPOLE_CFG = dict(masscart=1.0, masspole=0.1, length=0.5, max_force_mag=30.0, friction=0.5)
ALPHA_BOUND = float(10.0 / 180.0 * math.pi)  # fixed_wing_dynamics.py:10

# fixed normalisation of the wing state (neural_control/dataset.py:284-300)
WING_MEAN = (0.0, 0.0, 0.0, 11.525899887084961, -0.00016766408225521445, 0.16617104411125183,
             0.007394296582788229, 0.018172707409, 0.020353179425001144, -0.0005361468647606671,
             0.01662314310669899, 0.004487641621381044)
WING_STD = (16.626325607299805, 0.8449159860610962, 0.8879243731498718, 0.6243225932121277,
            0.28072822093963623, 0.29176747798, 0.04499124363064766, 0.10370047390460968, 0.049977313727,
            0.06449887901544571, 0.27508440613746643, 0.05634994804859)


def _cols(x):
    return [x[:, i] for i in range(x.shape[1])]


# --------------------------------------------------------------------------------------------
# A7  quadrotor step  (dynamics/quad_dynamics_flightmare.py:128-216, quad_dynamics_base.py:59-127)
# --------------------------------------------------------------------------------------------
def quad_inertia(cfg=QUAD_CFG):
    """J = m/12 * l^2 * frame_inertia  (quad_dynamics_base.py:32-35)."""
    s = cfg["mass"] / 12.0 * cfg["arm_length"] ** 2
    return tuple(s * f for f in cfg["frame_inertia"])


def quad_step(state, action, dt, cfg=QUAD_CFG):
    """state (N,12) = [pos, euler rpy, vel world, body rates]; action (N,4) in [0,1]."""
    px, py, pz, roll, pitch, yaw, vx, vy, vz, wx, wy, wz = _cols(state)
    a0, a1, a2, a3 = _cols(action)
    dt_ = float(dt)
    Jx, Jy, Jz = cfg["inertia_vector"] if "inertia_vector" in cfg else quad_inertia(cfg)
    Kx, Ky, Kz = cfg["kinv_ang_vel_tau"]
    gx, gy, gz = cfg["gravity"]
    tdx, tdy, tdz = cfg["translational_drag"]
    rdx, rdy, rdz = cfg["rotational_drag"]
    m = cfg["mass"]

    thrust = a0 * 15 - 7.5 + 9.81                                   # :139
    brx, bry, brz = a1 - 0.5, a2 - 0.5, a3 - 0.5                     # :140
    # c = w x (J w)                                                  # :146-149
    jwx, jwy, jwz = Jx * wx, Jy * wy, Jz * wz
    cx = wy * jwz - wz * jwy
    cy = wz * jwx - wx * jwz
    cz = wx * jwy - wy * jwx
    # desired torque = J K (br - w) + c + rot_drag                   # :95-117
    tx = Jx * (Kx * (brx - wx)) + cx + rdx
    ty = Jy * (Ky * (bry - wy)) + cy + rdy
    tz = Jz * (Kz * (brz - wz)) + cz + rdz
    force = m * thrust                                               # :101
    # acceleration = 1/m * R_body_to_world [0,0,force] + g + drag    # :74-93
    Cy, Sy = torch.cos(yaw), torch.sin(yaw)
    Cp, Sp = torch.cos(pitch), torch.sin(pitch)
    Cr, Sr = torch.cos(roll), torch.sin(roll)
    f_over_m = (1.0 / m) * force
    ax = f_over_m * (Cy * Sp * Cr + Sr * Sy) + gx + tdx
    ay = f_over_m * (Cr * Sy * Sp - Cy * Sr) + gy + tdy
    az = f_over_m * (Cr * Cp) + gz + tdz
    # position uses 0.5*dt*vel (reference quirk, :172-174)
    npx = px + 0.5 * dt_ * dt_ * ax + 0.5 * dt_ * vx
    npy = py + 0.5 * dt_ * dt_ * ay + 0.5 * dt_ * vy
    npz = pz + 0.5 * dt_ * dt_ * az + 0.5 * dt_ * vz
    nvx, nvy, nvz = vx + dt_ * ax, vy + dt_ * ay, vz + dt_ * az     # :175
    # angular acceleration = J^-1 (tau - c)                          # :178-183
    nwx = wx + dt_ * ((tx - cx) / Jx)
    nwy = wy + dt_ * ((ty - cy) / Jy)
    nwz = wz + dt_ * ((tz - cz) / Jz)
    # attitude integrates the OLD body rates through E(att)         # :210, base :96-127
    er = wx - Sp * wz
    ep = Cr * wy + Cp * Sr * wz
    ey = -Sr * wy + Cp * Cr * wz
    nroll, npitch, nyaw = roll + dt_ * er, pitch + dt_ * ep, yaw + dt_ * ey
    return torch.stack((npx, npy, npz, nroll, npitch, nyaw, nvx, nvy, nvz, nwx, nwy, nwz), dim=1)


def world_to_body_rows(att):
    """Rows of the world->body matrix (quad_dynamics_base.py:59-94)."""
    roll, pitch, yaw = att[:, 0], att[:, 1], att[:, 2]
    Cy, Sy = torch.cos(yaw), torch.sin(yaw)
    Cp, Sp = torch.cos(pitch), torch.sin(pitch)
    Cr, Sr = torch.cos(roll), torch.sin(roll)
    r0 = (Cy * Cp, Sy * Cp, -Sp)
    r1 = (Cy * Sp * Sr - Cr * Sy, Cr * Cy + Sr * Sy * Sp, Cp * Sr)
    r2 = (Cy * Sp * Cr + Sr * Sy, Cr * Sy * Sp - Cy * Sr, Cr * Cp)
    return r0, r1, r2


# A4  per-step featurizer (neural_control/dataset.py:207-220, 146-153)
def state_preprocessing(state):
    """(N,12) -> (N,15): [vel(3), W00,W01,W10,W11,W20,W21, W.vel(3), body rates(3)]."""
    vel = state[:, 6:9]
    r0, r1, r2 = world_to_body_rows(state[:, 3:6])
    vb = [r[0] * vel[:, 0] + r[1] * vel[:, 1] + r[2] * vel[:, 2] for r in (r0, r1, r2)]
    feats = [vel[:, 0], vel[:, 1], vel[:, 2], r0[0], r0[1], r1[0], r1[1], r2[0], r2[1],
             vb[0], vb[1], vb[2], state[:, 9], state[:, 10], state[:, 11]]
    return torch.stack(feats, dim=1)


# --------------------------------------------------------------------------------------------
# A8  fixed-wing step (dynamics/fixed_wing_dynamics.py:98-267)
# --------------------------------------------------------------------------------------------
def wing_step(state, action, dt, cfg=WING_CFG):
    """state (N,12) = [pos NED, vel body uvw, euler phi/theta/psi, body rates pqr]."""
    x, y, z, u, v, w, phi, theta, psi, p, q, r = _cols(state)
    a0, a1, a2, a3 = _cols(action)
    pi = math.pi
    T = a0 * 7                                                        # :41-46
    del_e = pi * (a1 * 40 - 20) / 180
    del_a = pi * (a2 * 5 - 2.5) / 180
    del_r = pi * (a3 * 40 - 20) / 180
    g_m = cfg["g"] * cfg["mass"]

    V = torch.sqrt(u ** 2 + v ** 2 + w ** 2)                          # :129
    alpha = torch.clamp(torch.arctan(w / u), -ALPHA_BOUND, ALPHA_BOUND)   # :130-131
    beta = torch.clamp(torch.arctan(v / V), -ALPHA_BOUND, ALPHA_BOUND)    # :132-133
    c2v = cfg["c"] / (2 * V)
    b2v = cfg["b"] / (2 * V)
    CL = cfg["CL0"] + cfg["CL_alpha"] * alpha + cfg["CL_q"] * c2v * q + cfg["CL_del_e"] * del_e
    CD = cfg["CD0"] + cfg["CD_alpha"] * alpha + cfg["CD_q"] * c2v * q + cfg["CD_del_e"] * del_e
    CY = (cfg["CY0"] + cfg["CY_beta"] * beta + cfg["CY_p"] * b2v * p + cfg["CY_r"] * b2v * r
          + cfg["CY_del_a"] * del_a + cfg["CY_del_r"] * del_r)
    Cl = (cfg["Cl0"] + cfg["Cl_beta"] * beta + cfg["Cl_p"] * b2v * p + cfg["Cl_r"] * b2v * r
          + cfg["Cl_del_a"] * del_a + cfg["Cl_del_r"] * del_r)
    Cm = cfg["Cm0"] + cfg["Cm_alpha"] * alpha + cfg["Cm_q"] * c2v * q + cfg["Cm_del_e"] * del_e
    Cn = (cfg["Cn0"] + cfg["Cn_beta"] * beta + cfg["Cn_p"] * b2v * p + cfg["Cn_r"] * b2v * r
          + cfg["Cn_del_a"] * del_a + cfg["Cn_del_r"] * del_r)
    qS = 0.5 * cfg["rho"] * V ** 2 * cfg["S"]
    L, D, Y = qS * CL, qS * CD, qS * CY
    # all three moments are scaled by the chord c (reference quirk, :170-175)
    l_m, m_m, n_m = qS * cfg["c"] * Cl, qS * cfg["c"] * Cm, qS * cfg["c"] * Cn

    sa, ca, sb, cb = torch.sin(alpha), torch.cos(alpha), torch.sin(beta), torch.cos(beta)
    sph, cph, sth, cth = torch.sin(phi), torch.cos(phi), torch.sin(theta), torch.cos(theta)
    sps, cps = torch.sin(psi), torch.cos(psi)
    # f = R_bw [-D, Y, -L] + R(phi,theta,0) [0,0,g m] + [T cos eps, 0, T sin eps]   (:185-204)
    fx = ca * cb * (-D) + (-ca * sb) * Y + (-sa) * (-L) + (-sth) * g_m + T * math.cos(cfg["epsilon"])
    fy = sb * (-D) + cb * Y + (sph * cth) * g_m
    fz = sa * cb * (-D) + (-sa * sb) * Y + ca * (-L) + (cph * cth) * g_m + T * math.sin(cfg["epsilon"])
    # pos_dot = R_ib(phi,theta,psi) uvw   (:213-216, 65-93)
    xd = cth * cps * u + (-cph * sps + sph * sth * cps) * v + (sph * sps + cph * sth * cps) * w
    yd = cth * sps * u + (cph * cps + sph * sth * sps) * v + (-sph * cps + cph * sth * sps) * w
    zd = -sth * u + sph * cth * v + cph * cth * w
    # uvw_dot = f/m - omega x vel   (:220-221)
    inv_m = 1.0 / cfg["mass"]
    ud = inv_m * fx - (q * w - r * v)
    vd = inv_m * fy - (r * u - p * w)
    wd = inv_m * fz - (p * v - q * u)
    # euler-angle kinematics (:225-245)
    tth = torch.tan(theta)
    phid = p + sph * tth * q + cph * tth * r
    thetad = cph * q - sph * r
    psid = (sph / cth) * q + (cph / cth) * r
    # omega_dot = I^-1 (M - omega x I omega), I with -I_xz off-diagonals (:33-39, 250-255)
    Ixx, Iyy, Izz, off = cfg["I_xx"], cfg["I_yy"], cfg["I_zz"], -cfg["I_xz"]
    Iw_x = Ixx * p + off * r
    Iw_y = Iyy * q
    Iw_z = off * p + Izz * r
    rx = l_m - (q * Iw_z - r * Iw_y)
    ry = m_m - (r * Iw_x - p * Iw_z)
    rz = n_m - (p * Iw_y - q * Iw_x)
    det = Ixx * Izz - off * off
    pd = (Izz * rx - off * rz) / det
    qd = ry / Iyy
    rd = (-off * rx + Ixx * rz) / det
    dt_ = float(dt)
    dot = torch.stack((xd, yd, zd, ud, vd, wd, phid, thetad, psid, pd, qd, rd), dim=1)
    return state + dt_ * dot                                           # :265


# --------------------------------------------------------------------------------------------
# A9  cartpole step (dynamics/cartpole_dynamics.py:53-119)
# --------------------------------------------------------------------------------------------
def cartpole_step(state, action, dt, cfg=CARTPOLE_CFG):
    """state (N,4) = [x, xdot, theta, thetadot]; action (N,1) in [-1,1]."""
    gravity = 9.81
    total_mass = cfg["masspole"] + cfg["masscart"]
    pml = cfg["masspole"] * cfg["length"]
    fr = cfg["friction"]
    x, xdot, th, thdot = _cols(state)
    force = action[:, 0] * cfg["max_force_mag"] * 0.5                  # :60
    s, c = torch.sin(th), torch.cos(th)
    xacc = (-2 * pml * thdot ** 2 * s + 3 * cfg["masspole"] * gravity * s * c + 4 * force
            - 4 * fr * xdot) / (4 * total_mass - 3 * cfg["masspole"] * c ** 2)          # :86-97
    thacc = (-3 * pml * thdot ** 2 * s * c + 6 * total_mass * gravity * s
             + 6 * (force - fr * xdot) * c) / (4 * cfg["length"] * total_mass - 3 * pml * c ** 2)  # :99-111
    dt_ = float(dt)
    sd, cd = torch.sin(thdot * dt_), torch.cos(thdot * dt_)           # :113-119
    new_s = s * cd + c * sd
    new_c = c * cd - s * sd
    return torch.stack((x + xdot * dt_, xdot + xacc * dt_, torch.atan2(new_s, new_c),
                        thdot + thacc * dt_), dim=1)


STEP_FN = {"quad": quad_step, "wing": wing_step, "cartpole": cartpole_step}


# --------------------------------------------------------------------------------------------
# A10-A12 losses (neural_control/drone_loss.py:12-39, 72-82, 136-145); sums, not means
# --------------------------------------------------------------------------------------------
def quad_mpc_loss(states, ref, actions):
    pos = ((states[:, :, 0:3] - ref[:, :, 0:3]) ** 2).sum()
    vel = ((states[:, :, 6:9] - ref[:, :, 6:9]) ** 2).sum()
    av = (states[:, :, 9:12] ** 2).sum()
    thr = ((actions[:, :, 0] - 0.5) ** 2).sum()
    rates = ((actions[:, :, 1:4] - 0.5) ** 2).sum()
    return 10 * pos + 1 * vel + 0.1 * av + 0.1 * rates + 5 * thr


def fixed_wing_mpc_loss(states, lin_ref, actions):
    return 10 * ((states[:, :, 0:3] - lin_ref) ** 2).sum() + 0.1 * ((actions[:, :, 1:4] - 0.5) ** 2).sum()


def cartpole_loss_mpc(states, ref, actions):
    wts = torch.tensor([0.0, 3.0, 10.0, 1.0], dtype=states.dtype)
    return (((states - ref) ** 2) * wts).sum() + 0.01 * (actions ** 2).sum()


def cartpole_make_reference(cur, h):
    """scripts/train_cartpole.py:103-110: ref[:,k] = cur*(1-k/(h-1)) for k<h-1, last row 0; no grad."""
    ref = torch.zeros(cur.shape[0], h, cur.shape[1], dtype=cur.dtype)
    for k in range(h - 1):
        ref[:, k] = cur.detach() * (1 - 1 / (h - 1) * k)
    return ref


LOSS_FN = {"quad": quad_mpc_loss, "wing": fixed_wing_mpc_loss, "cartpole": cartpole_loss_mpc}


# --------------------------------------------------------------------------------------------
# A1-A3 policies.  ``params`` is the list of tensors in net.parameters() order, torch layouts:
#   hutter : states_in.{w,b}, conv_ref.{w,b}, ref_in.{w,b}, fc1, fc2, fc3, fc_out      (14 tensors)
#   lstm   : conv_ref.{w,b}, ref_in.{w,b}, fc_out.{w,b}, lstm.{w_ih,w_hh,b_ih,b_hh}    (10 tensors)
#   simple : fc0, fc1, fc2, fc3, fc_out                                               (10 tensors)
# --------------------------------------------------------------------------------------------
def _conv_encoder(ref, w, b):
    """Conv1d(ref_dim->20,k=3,valid) over the horizon axis + relu, flattened channel-major
    (index c*(h-2)+t)  (models/hutter_model.py:36-40)."""
    n, h, d = ref.shape
    cols = torch.stack([ref[:, t:t + 3, :] for t in range(h - 2)], dim=1)   # (N, h-2, 3, d): [j, ch]
    # out[n,t,c] = b[c] + sum_{ch,j} w[c,ch,j] * ref[n,t+j,ch]
    out = torch.einsum("ntjd,cdj->ntc", cols, w) + b
    return torch.relu(out).transpose(1, 2).reshape(n, -1)


def hutter_forward(params, state, ref, conv=True):
    """models/hutter_model.py:32-49 (logits; the caller applies sigmoid)."""
    ws, bs, wc, bc, wr, br, w1, b1, w2, b2, w3, b3, wo, bo = params
    s = torch.tanh(state @ ws.t() + bs)
    if conv:
        r = _conv_encoder(ref, wc, bc)
    else:
        r = torch.tanh(ref.reshape(ref.shape[0], -1) @ wr.t() + br)
    x = torch.cat((s, r), dim=1)
    x = torch.tanh(x @ w1.t() + b1)
    x = torch.tanh(x @ w2.t() + b2)
    x = torch.tanh(x @ w3.t() + b3)
    return x @ wo.t() + bo


def lstm_forward(params, state, ref, hc):
    """models/rnn.py:34-50: conv encoder -> LSTMCell(175,8) (gate order i,f,g,o) -> Linear(8,out)."""
    wc, bc, wr, br, wo, bo, w_ih, w_hh, b_ih, b_hh = params
    h_prev, c_prev = hc
    x = torch.cat((state, _conv_encoder(ref, wc, bc)), dim=1)
    gates = x @ w_ih.t() + b_ih + h_prev @ w_hh.t() + b_hh
    i, f, g, o = gates.chunk(4, dim=1)
    c_new = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new @ wo.t() + bo, (h_new, c_new)


def simple_forward(params, x):
    """models/simple_model.py:20-28: column 0 zeroed, tanh on every layer including the output."""
    x = torch.cat((torch.zeros_like(x[:, :1]), x[:, 1:]), dim=1)
    for i in range(0, 10, 2):
        x = torch.tanh(x @ params[i].t() + params[i + 1])
    return x


# --------------------------------------------------------------------------------------------
# A6 concurrent rollouts (scripts/train_drone.py:175-203, train_fixed_wing.py:90-116,
#    train_cartpole.py:118-155, train_base.py:202-206)
# --------------------------------------------------------------------------------------------
def rollout_concurrent(system, params, in_state, cur, in_ref, ref, h, dt):
    """Returns (loss, states (N,h,S), actions (N,h,A))."""
    if system == "quad":
        act = torch.sigmoid(hutter_forward(params, in_state, in_ref, conv=True)).reshape(-1, h, 4)
    elif system == "wing":
        act = torch.sigmoid(hutter_forward(params, in_state, in_ref, conv=False)).reshape(-1, h, 4)
    elif system == "cartpole":
        act = simple_forward(params, in_state).reshape(-1, h, 1)
        ref = cartpole_make_reference(cur, h)
    else:
        raise ValueError(system)
    step = STEP_FN[system]
    states = []
    s = cur
    for k in range(h):
        s = step(s, act[:, k], dt)
        states.append(s)
    states = torch.stack(states, dim=1)
    return LOSS_FN[system](states, ref, act), states, act


# --------------------------------------------------------------------------------------------
# A5 autoregressive / LSTM rollouts (scripts/train_drone.py:113-173), functional restatement
# --------------------------------------------------------------------------------------------
def recurrent_window(in_ref0, k, h, pos_hist, window):
    """Window of h reference rows seen by the policy at step k.

    cumulative (the reference's actual forward semantics, caused by the in-place write into the shared
    (N,2h,9) buffer at train_drone.py:138-142): row j = k+r holds
        in_ref0[j,:3] - sum_{i=max(0,j-h+1)}^{k} pos_i
    relative (documented intent): in_ref0[j,:3] - pos_k.   Columns 3:9 are never touched.
    ``pos_hist`` = [pos_0 .. pos_k] (each (N,3)).
    """
    win = in_ref0[:, k:k + h]
    if window == "relative":
        sub = pos_hist[k][:, None, :].expand(-1, h, -1)
    elif window == "cumulative":
        rows = []
        for r in range(h):
            lo = max(0, k + r - h + 1)
            acc = pos_hist[lo]
            for i in range(lo + 1, k + 1):
                acc = acc + pos_hist[i]
            rows.append(acc)
        sub = torch.stack(rows, dim=1)
    else:
        raise ValueError(window)
    return torch.cat((win[:, :, :3] - sub, win[:, :, 3:]), dim=2)


def rollout_recurrent(mode, params, cur, in_ref0, ref, h, dt, window="cumulative", hc0=None):
    """Quadrotor only.  mode in {'autoregressive','lstm'}.  in_ref0, ref: (N,2h,9).
    Returns (loss, states (N,h,12), actions (N,h,4))."""
    s = cur
    pos_hist, states, actions = [], [], []
    hc = hc0
    for k in range(h):
        pos_hist.append(s[:, :3])
        win = recurrent_window(in_ref0, k, h, pos_hist, window)
        feat = state_preprocessing(s)
        if mode == "autoregressive":
            logits = hutter_forward(params, feat, win, conv=True)
        elif mode == "lstm":
            logits, hc = lstm_forward(params, feat, win, hc)
        else:
            raise ValueError(mode)
        a = torch.sigmoid(logits)
        s = quad_step(s, a, dt)
        states.append(s)
        actions.append(a)
    states = torch.stack(states, dim=1)
    actions = torch.stack(actions, dim=1)
    return quad_mpc_loss(states, ref[:, :h], actions), states, actions


# --------------------------------------------------------------------------------------------
# value-and-gradient drivers
# --------------------------------------------------------------------------------------------
def value_and_grad(fn, params, *args, **kw):
    """Runs ``fn(params, ...)`` -> (loss, states, actions) and back-propagates the loss.
    Returns (loss, [grad or None per param], states, actions); unused params get None, exactly like
    the reference leaves ``.grad`` None for ``ref_in`` (conv nets) / ``conv_ref`` (wing nets)."""
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    loss, states, actions = fn(ps, *args, **kw)
    grads = torch.autograd.grad(loss, ps, allow_unused=True)
    return loss.detach(), list(grads), states.detach(), actions.detach()


def concurrent_value_and_grad(system, params, in_state, cur, in_ref, ref, h, dt):
    return value_and_grad(lambda ps: rollout_concurrent(system, ps, in_state, cur, in_ref, ref, h, dt), params)


def recurrent_value_and_grad(mode, params, cur, in_ref0, ref, h, dt, window="cumulative", hc0=None):
    return value_and_grad(
        lambda ps: rollout_recurrent(mode, ps, cur, in_ref0, ref, h, dt, window=window, hc0=hc0), params)


def sgd_momentum_step(params, grads, bufs, lr, momentum=0.9):
    """A13: optim.SGD(lr, momentum=0.9) update (scripts/train_base.py:139-143); None grads are skipped."""
    out_p, out_b = [], []
    for p, g, b in zip(params, grads, bufs):
        if g is None:
            out_p.append(p)
            out_b.append(b)
            continue
        nb = g.clone() if b is None else momentum * b + g
        out_p.append(p - lr * nb)
        out_b.append(nb)
    return out_p, out_b


# --------------------------------------------------------------------------------------------
# N2  closed-loop evaluation on table references: QuadEvaluator.follow_trajectory("rand")
#     (scripts/evaluate_drone.py:81-194) with trajectory/random_traj.py:62-96 (Random.get_ref_traj /
#     project_on_ref / get_current_full_state), NetworkWrapper.predict_actions (controllers/network_wrapper.py:41-71),
#     QuadRotorEnvBase.step / get_is_stable (environments/drone_env.py:66-115) and the run_eval statistics
#     (evaluate_drone.py:236-298), restated for N independent drones at once.
# --------------------------------------------------------------------------------------------
def eval_window(tables, ci, h):
    """Random.get_ref_traj for every drone: tables (N,RL,9), ci (N,) current index -> (rows (N,h,9), new ci).
    Regular case: rows ci+1 .. ci+h and the index advances; at the end of the table (ci >= RL-h) the remaining rows
    from ci on, padded with [last position, 0...], and the index stays (random_traj.py:67-81)."""
    n, rl, w = tables.shape
    end = ci >= rl - h
    start = torch.where(end, ci, ci + 1)
    nreal = torch.where(end, rl - ci, torch.full_like(ci, h))
    r = torch.arange(h)[None, :]
    idx = (start[:, None] + r).clamp(max=rl - 1)
    rows = torch.gather(tables, 1, idx[:, :, None].expand(n, h, w))
    pad = torch.zeros(n, h, w, dtype=tables.dtype)
    pad[:, :, :3] = tables[:, -1:, :3]
    rows = torch.where((r < nreal[:, None])[:, :, None], rows, pad)
    return rows, torch.where(end, ci, ci + 1)


def eval_follow_tables(params, tables, init_states, steps, h, dt, thresh_div=1.0, thresh_stable=1.0, test_time=0,
                       cfg=QUAD_CFG, record_policy_inputs=False, hc0=None):
    """Batched restatement of QuadEvaluator.follow_trajectory("rand").

    hc0 = (h0 (N,8), c0 (N,8)): `params` are those of the LSTM policy (models/rnn.py LSTM_NEW, 10 tensors); every
    policy call advances the drone's hidden / cell state (rnn.py:45-48) - resets of the drone do not touch it, the
    reference only re-draws it when an evaluator is constructed (evaluate_drone.py:55-57).  The final state is
    returned as out["hc"].

    params: hutter Net(15,h,9,4h) (concurrent: the first of the h predicted actions is applied,
    evaluate_drone.py:154-155) or Net(15,h,9,4); tables (N,RL,9) reference rows [pos, euler, vel] as
    Random.__init__ leaves them; init_states (N,12).  Per step: window -> QuadDataset.prepare_data -> net ->
    sigmoid -> clip -> dynamics (evaluated in float64 on the float64 env state and rounded to float32,
    drone_env.py:94-102) -> divergence to tables[ci,:3] -> if diverged / unstable: stop (test_time) or reset the
    drone to the reference state.  Returns dict(states (N,steps+1,12), div (N,steps), actions (N,steps,4),
    n_steps (N,) = steps taken before the break; entries after a break are zero); with record_policy_inputs also
    policy_states (N,steps,12) and windows (N,steps,h,9): the raw (state, reference rows) every policy call hands to
    the dataset (network_wrapper.py:47-52)."""
    n, rl, _ = tables.shape
    tables = tables.float()
    s = init_states.float().clone()
    ci = torch.zeros(n, dtype=torch.long)
    alive = torch.ones(n, dtype=torch.bool)
    states = torch.zeros(n, steps + 1, 12)
    states[:, 0] = s
    divs, actions = torch.zeros(n, steps), torch.zeros(n, steps, 4)
    n_steps = torch.zeros(n, dtype=torch.long)
    out_dim = params[-1].shape[0] if hc0 is None else 4
    hc = None if hc0 is None else (hc0[0].float().clone(), hc0[1].float().clone())
    pol_states, windows = torch.zeros(n, steps, 12), torch.zeros(n, steps, h, 9)
    for i in range(steps):
        if not bool(alive.any()):
            break
        rows, ci_new = eval_window(tables, ci, h)
        ci = torch.where(alive, ci_new, ci)
        pol_states[alive, i] = s[alive]
        windows[alive, i] = rows[alive]
        cur = s.clone()
        rel = rows.clone()
        rel[:, :, :3] = rel[:, :, :3] - cur[:, None, :3]                 # dataset.py:170-173
        cur[:, :3] = 0
        in_ref = torch.cat((rel[:, :, :3], rel[:, :, 6:9], rel[:, :, 6:9] - cur[:, None, 6:9]), dim=2)
        with torch.no_grad():
            if hc0 is None:
                act = torch.sigmoid(hutter_forward(params, state_preprocessing(cur), in_ref))
            else:
                logits, (h_new, c_new) = lstm_forward(params, state_preprocessing(cur), in_ref, hc)
                hc = (torch.where(alive[:, None], h_new, hc[0]), torch.where(alive[:, None], c_new, hc[1]))
                act = torch.sigmoid(logits)
        a0 = (act.reshape(n, h, 4)[:, 0] if out_dim == 4 * h else act).clamp(0.0, 1.0)
        nxt = quad_step(s.double(), a0.double(), dt, cfg).float()
        stable = (nxt[:, 3:5].abs() < thresh_stable).all(dim=1)           # drone_env.py:66-74
        on_line = tables[torch.arange(n), ci, :3]                         # random_traj.py:83-87
        div = (on_line.double() - nxt[:, :3].double()).norm(dim=1).float()
        states[alive, i + 1] = nxt[alive]
        divs[alive, i] = div[alive]
        actions[alive, i] = a0[alive]
        n_steps[alive] += 1
        bad = (div > thresh_div) | ~stable
        reset_state = torch.cat((tables[torch.arange(n), ci], torch.zeros(n, 3)), dim=1)   # random_traj.py:89-92
        if test_time:
            s = torch.where(alive[:, None], nxt, s)
            alive = alive & ~bad
        else:
            s = torch.where((alive & bad)[:, None], reset_state, torch.where(alive[:, None], nxt, s))
        alive = alive & (i < rl)                                          # evaluate_drone.py:187-188
    out = dict(states=states, div=divs, actions=actions, n_steps=n_steps)
    if hc is not None:
        out["hc"] = hc
    if record_policy_inputs:
        out.update(policy_states=pol_states, windows=windows)
    return out


def selfplay_kept_calls(n_steps, take_every_x, action_counter=0):
    """Which policy calls of runs taken one after the other NetworkWrapper.predict_actions adds to the dataset
    (network_wrapper.py:47: call number c, 1-based and running on over the runs, is kept when c % take_every_x == 0).
    Returns ([(run, step), ...] in call order, the action counter afterwards)."""
    kept, c = [], int(action_counter)
    for run, ns in enumerate([int(x) for x in n_steps]):
        for step in range(ns):
            c += 1
            if c % take_every_x == 0:
                kept.append((run, step))
    return kept, c


def selfplay_ring_slots(n_kept, num_sampled, num_self_play, eval_counter=0):
    """DroneDataset.get_eval_index (dataset.py:78-85) for n_kept consecutive additions: slot of each, counter after"""
    return [num_sampled + (eval_counter + j) % num_self_play for j in range(n_kept)], eval_counter + n_kept


def eval_statistics(div, n_steps, thresh_div):
    """QuadEvaluator.run_eval (evaluate_drone.py:266-298) over the N runs: per run mean divergence and the number
    of steps below the threshold; returns the 6-tuple of run_eval."""
    import numpy as np
    div, n_steps = div.numpy(), n_steps.numpy()
    d = np.array([div[i, :n_steps[i]].mean() for i in range(len(n_steps))])
    stable = np.array([(div[i, :n_steps[i]] < thresh_div).sum() for i in range(len(n_steps))])
    full = d[stable == n_steps[-1]]              # max_steps_stable = len(reference_traj) of the LAST run
    return stable.mean(), stable.std(), full.mean() if len(full) else float("nan"), \
        full.std() if len(full) else float("nan"), d.mean(), d.std()


# --------------------------------------------------------------------------------------------
# N3  learnt residual dynamics of the quadrotor (neural_control/dynamics/quad_dynamics_trained.py:10-69):
#     next = simulate_quadrotor(linear_at @ action, state, dt) + linear_state_2(relu(linear_state_1([state, at])))
#     with mass / inertia vector / kinv vector as (differentiable) parameters of the simulator.
#     lparams in named_parameters() order: linear_at (4,4), mass (1,), torch_inertia_vector (3,),
#     torch_kinv_vector (3,), linear_state_1.weight (64,16), .bias (64,), linear_state_2.weight (12,64), .bias (12,)
# --------------------------------------------------------------------------------------------
def learnt_quad_step(lparams, state, action, dt, cfg=QUAD_CFG):
    lin_at, mass, jvec, kvec, w1, b1, w2, b2 = lparams
    at = action @ lin_at.t()                                              # :59-61
    c = dict(cfg)
    c["mass"] = mass[0]
    c["inertia_vector"] = (jvec[0], jvec[1], jvec[2])
    c["kinv_ang_vel_tau"] = (kvec[0], kvec[1], kvec[2])
    new_state = quad_step(state, at, dt, c)                               # :63
    x = torch.cat((state, at), dim=1)                                     # :51-56
    added = torch.relu(x @ w1.t() + b1) @ w2.t() + b2
    return new_state + added                                              # :65-66


def rollout_concurrent_learnt(params, lparams, in_state, cur, in_ref, ref, h, dt, cfg=QUAD_CFG):
    """The controller-training phase of run_dynamics (scripts/train_base.py:334-375): TrainDrone.train_controller_model
    (scripts/train_drone.py:175-199) with self.train_dynamics = LearntDynamics - the concurrent quadrotor rollout whose
    h steps are taken by the learnt model.  Returns (loss, states (N,h,12), actions (N,h,4))."""
    act = torch.sigmoid(hutter_forward(params, in_state, in_ref, conv=True)).reshape(-1, h, 4)   # train_base.py:202-206
    states, s = [], cur
    for k in range(h):
        s = learnt_quad_step(lparams, s, act[:, k], dt, cfg)                                      # train_drone.py:186-190
        states.append(s)
    states = torch.stack(states, dim=1)
    return quad_mpc_loss(states, ref, act), states, act


def learnt_dynamics_loss(lparams, state, action, target_next, dt, l2_lambda=0.0, cfg=QUAD_CFG):
    """TrainBase.train_dynamics_model (scripts/train_base.py:160-186): sum of squared differences between the learnt
    step and the target dynamics' step on the first action, + l2_lambda * (norms of the residual MLP tensors)."""
    nxt = learnt_quad_step(lparams, state, action, dt, cfg)
    loss = torch.sum((nxt - target_next) ** 2)
    if l2_lambda > 0:
        w1, b1, w2, b2 = lparams[4:]
        loss = loss + l2_lambda * (torch.norm(w2) + torch.norm(b2) + torch.norm(w1) + torch.norm(b1))
    return loss


# --------------------------------------------------------------------------------------------
# N2 (fixed wing)  closed-loop evaluation: FixedWingEvaluator.fly_to_point (scripts/evaluate_fixed_wing.py:46-130)
#     with FixedWingNetWrapper.predict_actions (controllers/network_wrapper.py:81-98), WingDataset.prepare_data,
#     SimpleWingEnv.step (environments/wing_env.py:44-58) and project_to_line (trajectory/q_funcs.py:6-18),
#     restated for N independent drones at once.
# --------------------------------------------------------------------------------------------
def project_to_line(a, b, p):
    """rows of a, b, p (N,3): projection of p onto the line through a and b; a where a == b (q_funcs.py:11-12)"""
    ab = b - a
    nrm = (ab * ab).sum(dim=1, keepdim=True)
    same = (a == b).all(dim=1, keepdim=True)
    t = ((p - a) * ab).sum(dim=1, keepdim=True)
    return torch.where(same, a, a + ab * t / torch.where(same, torch.ones_like(nrm), nrm))


def eval_fly_to_points(params, targets, init_states, mean, std, steps, h, dt_data, dt_env, thresh_div=10.0,
                       thresh_stable=0.8, test_time=0, des_speed=11.5, cfg=WING_CFG, record_policy_inputs=False):
    """Batched restatement of FixedWingEvaluator.fly_to_point.  params: hutter Net(9,1,3,4h, conv=False);
    targets (N,K,3); init_states (N,12) (zero_reset: zeros with u = 11.5).  Kept quirk: after a reset the policy
    still sees the PRE-reset state for one step (the evaluator's local `state` is not refreshed, :118-129) while the
    environment continues from the reset state.
    Returns dict(states (N,steps+1,12) as returned by env.step, div_linear (N,steps), actions (N,steps,4),
    n_steps (N,), div_target_sum (N,), div_target_cnt (N,)); with record_policy_inputs also policy_states
    (N,steps,12) and target_index (N,steps): what every policy call hands to the dataset (network_wrapper.py:85-87)."""
    n, K, _ = targets.shape
    mean, std = torch.as_tensor(mean).float(), torch.as_tensor(std).float()
    env = init_states.float().clone()               # the environment's state
    obs = env.clone()                               # what the policy is shown
    prev_pos = env[:, :3].clone()
    line_start = env[:, :3].clone()
    ti = torch.zeros(n, dtype=torch.long)
    alive = torch.ones(n, dtype=torch.bool)
    states = torch.zeros(n, steps + 1, 12)
    states[:, 0] = env
    div_lin, actions = torch.zeros(n, steps), torch.zeros(n, steps, 4)
    n_steps = torch.zeros(n, dtype=torch.long)
    dts, dtc = torch.zeros(n), torch.zeros(n)
    vlen = torch.tensor(12 * dt_data, dtype=torch.float32)
    ar = torch.arange(n)
    pol_states, pol_ti = torch.zeros(n, steps, 12), torch.zeros(n, steps, dtype=torch.long)
    for i in range(steps):
        if not bool(alive.any()):
            break
        tgt = targets[ar, ti].float()
        pol_states[alive, i] = obs[alive]
        pol_ti[alive, i] = ti[alive]
        normed = ((obs - mean) / std)[:, 3:]                                      # dataset.py:336
        rel = tgt - obs[:, :3]
        unit = rel / torch.sqrt((rel ** 2).sum(dim=1, keepdim=True))
        last = obs[:, :3] + unit * vlen * h                                        # :311-321, row h-1
        in_ref = last - obs[:, :3]                                                 # :346
        with torch.no_grad():
            act = torch.sigmoid(hutter_forward(params, normed, in_ref[:, None, :], conv=False))
        a0 = act.reshape(n, h, 4)[:, 0]
        nxt = wing_step(env, a0, dt_env, cfg).float()
        stable = (nxt[:, 6:8].abs() < thresh_stable).all(dim=1)                    # wing_env.py:54
        pos = nxt[:, :3]
        on_line = project_to_line(line_start, tgt, pos)
        div = (on_line - pos).norm(dim=1)
        states[alive, i + 1] = nxt[alive]
        div_lin[alive, i] = div[alive]
        actions[alive, i] = a0[alive]
        n_steps[alive] += 1
        passed = alive & (pos[:, 0] > tgt[:, 0])                                   # :93-110
        t_on = project_to_line(prev_pos, pos, tgt)
        dts = dts + torch.where(passed, (t_on - tgt).norm(dim=1), torch.zeros(n))
        dtc = dtc + passed.float()
        more = passed & (ti < K - 1)
        finished = passed & ~more
        line_start = torch.where(more[:, None], pos, line_start)
        ti = torch.where(more, ti + 1, ti)
        bad = alive & ~finished & (~stable | (div > thresh_div))                   # :112-129 (old target)
        if test_time:
            dts = dts + torch.where(bad, (pos - tgt).norm(dim=1), torch.zeros(n))
        else:
            dts = dts + torch.where(bad, torch.full((n,), float(thresh_div)), torch.zeros(n))
        dtc = dtc + bad.float()
        vec = tgt - on_line
        reset = torch.zeros(n, 12)
        reset[:, :3] = on_line
        reset[:, 3:6] = vec / vec.norm(dim=1, keepdim=True) * des_speed
        keep = alive[:, None]
        prev_pos = torch.where(keep, pos, prev_pos)
        obs = torch.where(keep, nxt, obs)
        if test_time:
            env = torch.where(keep, nxt, env)
            alive = alive & ~finished & ~bad
        else:
            env = torch.where((bad)[:, None], reset, torch.where(keep, nxt, env))
            alive = alive & ~finished
    maxed = alive & (n_steps == steps)                                             # :130-132
    dts = dts + torch.where(maxed, torch.full((n,), float(thresh_div)), torch.zeros(n))
    dtc = dtc + maxed.float()
    out = dict(states=states, div_linear=div_lin, actions=actions, n_steps=n_steps, div_target_sum=dts,
               div_target_cnt=dtc)
    if record_policy_inputs:
        out.update(policy_states=pol_states, target_index=pol_ti)
    return out


# --------------------------------------------------------------------------------------------
# N3 (fixed wing)  learnt dynamics (neural_control/dynamics/fixed_wing_dynamics.py:270-326, LearntFixedWingDynamics):
#     next = simulate_fixed_wing(state, action, dt) + linear_state_2(relu(linear_state_1([state, action])))
#     with EVERY physical constant a parameter: the full 3x3 inertia matrix `I` and one 1-element parameter per
#     config key (a ParameterDict, which orders its keys by sorting them), used live by the simulator.
#     lparams in named_parameters() order: I (3,3), cfg.<key> for key in WING_LEARNT_KEYS (1,) each,
#     linear_state_1.weight (64,16), .bias (64,), linear_state_2.weight (12,64), .bias (12,).
#     Quirk kept: gravity enters through `torch.tensor(g_m)` (:197), which DETACHES g * mass -> `g` gets no gradient
#     and `mass` only the one through 1/mass.
# --------------------------------------------------------------------------------------------
WING_LEARNT_KEYS = sorted(k for k in WING_CFG if not k.startswith("I_"))


def wing_step_general(state, action, dt, c, inertia):
    """wing_step with tensor-valued constants c[key] (0-dim or 1-element) and a general 3x3 inertia matrix, the
    inverse and products written out like the reference (torch.inverse / matmul, :250-255)"""
    x, y, z, u, v, w, phi, theta, psi, p, q, r = _cols(state)
    a0, a1, a2, a3 = _cols(action)
    pi = math.pi
    T = a0 * 7
    del_e = pi * (a1 * 40 - 20) / 180
    del_a = pi * (a2 * 5 - 2.5) / 180
    del_r = pi * (a3 * 40 - 20) / 180
    g_m = (c["g"] * c["mass"]).detach()                                 # torch.tensor(g_m): no gradient (:197)
    V = torch.sqrt(u ** 2 + v ** 2 + w ** 2)
    alpha = torch.clamp(torch.arctan(w / u), -ALPHA_BOUND, ALPHA_BOUND)
    beta = torch.clamp(torch.arctan(v / V), -ALPHA_BOUND, ALPHA_BOUND)
    c2v, b2v = c["c"] / (2 * V), c["b"] / (2 * V)
    CL = c["CL0"] + c["CL_alpha"] * alpha + c["CL_q"] * c2v * q + c["CL_del_e"] * del_e
    CD = c["CD0"] + c["CD_alpha"] * alpha + c["CD_q"] * c2v * q + c["CD_del_e"] * del_e
    CY = (c["CY0"] + c["CY_beta"] * beta + c["CY_p"] * b2v * p + c["CY_r"] * b2v * r + c["CY_del_a"] * del_a
          + c["CY_del_r"] * del_r)
    Cl = (c["Cl0"] + c["Cl_beta"] * beta + c["Cl_p"] * b2v * p + c["Cl_r"] * b2v * r + c["Cl_del_a"] * del_a
          + c["Cl_del_r"] * del_r)
    Cm = c["Cm0"] + c["Cm_alpha"] * alpha + c["Cm_q"] * c2v * q + c["Cm_del_e"] * del_e
    Cn = (c["Cn0"] + c["Cn_beta"] * beta + c["Cn_p"] * b2v * p + c["Cn_r"] * b2v * r + c["Cn_del_a"] * del_a
          + c["Cn_del_r"] * del_r)
    qS = 0.5 * c["rho"] * V ** 2 * c["S"]
    L, D, Y = qS * CL, qS * CD, qS * CY
    l_m, m_m, n_m = qS * c["c"] * Cl, qS * c["c"] * Cm, qS * c["c"] * Cn
    sa, ca, sb, cb = torch.sin(alpha), torch.cos(alpha), torch.sin(beta), torch.cos(beta)
    sph, cph, sth, cth = torch.sin(phi), torch.cos(phi), torch.sin(theta), torch.cos(theta)
    sps, cps = torch.sin(psi), torch.cos(psi)
    ce, se = torch.cos(c["epsilon"]), torch.sin(c["epsilon"])
    fx = ca * cb * (-D) + (-ca * sb) * Y + (-sa) * (-L) + (-sth) * g_m + T * ce
    fy = sb * (-D) + cb * Y + (sph * cth) * g_m
    fz = sa * cb * (-D) + (-sa * sb) * Y + ca * (-L) + (cph * cth) * g_m + T * se
    xd = cth * cps * u + (-cph * sps + sph * sth * cps) * v + (sph * sps + cph * sth * cps) * w
    yd = cth * sps * u + (cph * cps + sph * sth * sps) * v + (-sph * cps + cph * sth * sps) * w
    zd = -sth * u + sph * cth * v + cph * cth * w
    inv_m = 1.0 / c["mass"]
    ud = inv_m * fx - (q * w - r * v)
    vd = inv_m * fy - (r * u - p * w)
    wd = inv_m * fz - (p * v - q * u)
    tth = torch.tan(theta)
    phid = p + sph * tth * q + cph * tth * r
    thetad = cph * q - sph * r
    psid = (sph / cth) * q + (cph / cth) * r
    omega = torch.stack((p, q, r), dim=1)
    iw = omega @ inertia.t()                                             # I omega, per row
    mom = torch.stack((l_m, m_m, n_m), dim=1)
    rhs = mom - torch.cross(omega, iw, dim=1)
    od = rhs @ torch.inverse(inertia).t()                               # I^-1 rhs
    dot = torch.stack((xd, yd, zd, ud, vd, wd, phid, thetad, psid, od[:, 0], od[:, 1], od[:, 2]), dim=1)
    return state + float(dt) * dot


def learnt_wing_step(lparams, state, action, dt):
    inertia = lparams[0]
    c = {k: lparams[1 + i].reshape(()) for i, k in enumerate(WING_LEARNT_KEYS)}
    w1, b1, w2, b2 = lparams[1 + len(WING_LEARNT_KEYS):]
    new_state = wing_step_general(state, action, dt, c, inertia)
    x = torch.cat((state, action), dim=1)
    return new_state + torch.relu(x @ w1.t() + b1) @ w2.t() + b2


# --------------------------------------------------------------------------------------------
# N2 (cartpole)  closed-loop balancing evaluation: Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py:
#     81-262, state-based controller) with CartpoleWrapper.predict_actions (controllers/network_wrapper.py:101-148)
#     and CartPoleEnv._step / is_upright (environments/cartpole_env.py:52-84), for N independent runs at once.
#     Kept quirk: simple_model.Net zeroes column 0 of its input IN PLACE (simple_model.py:21); from the second step on
#     the wrapper's tensor shares memory with the environment's float32 state, so the cart POSITION of the
#     environment itself is reset to zero before every policy call but the first.
# --------------------------------------------------------------------------------------------
def eval_cartpole_balance(params, init_states, max_steps, dt, thresh_div=0.21, burn_in_steps=50, cfg=CARTPOLE_CFG):
    """Returns dict(states (N,max_steps,4) as returned by env._step, actions (N,max_steps), success (N,) = index of
    the last step taken, n_steps (N,), mean_angle (N,) = mean |theta| after the burn-in (100 if none),
    vel_sum (N,) = sum of |x_dot| over the steps)."""
    n = init_states.shape[0]
    s = init_states.float().clone()
    alive = torch.ones(n, dtype=torch.bool)
    states = torch.zeros(n, max_steps, 4)
    actions = torch.zeros(n, max_steps)
    n_steps = torch.zeros(n, dtype=torch.long)
    ang_sum, ang_cnt, vel_sum = torch.zeros(n), torch.zeros(n), torch.zeros(n)
    for i in range(max_steps):
        if not bool(alive.any()):
            break
        if i > 0:
            s = s.clone()
            s[:, 0] = 0                                                   # the in-place zeroing reached the env state
        x_in = s.clone()
        with torch.no_grad():
            a0 = simple_forward(params, x_in)[:, 0:1]                     # first of the h predicted actions
        nxt = cartpole_step(s, a0, dt, cfg).float()
        th = nxt[:, 2]
        th = torch.where(th > math.pi, th - 2 * math.pi, torch.where(th <= -math.pi, th + 2 * math.pi, th))
        nxt = torch.cat((nxt[:, :2], th[:, None], nxt[:, 3:]), dim=1)
        states[alive, i] = nxt[alive]
        actions[alive, i] = a0[alive, 0]
        n_steps[alive] += 1
        vel_sum = vel_sum + torch.where(alive, nxt[:, 1].abs(), torch.zeros(n))
        if i > burn_in_steps:
            ang_sum = ang_sum + torch.where(alive, nxt[:, 2].abs(), torch.zeros(n))
            ang_cnt = ang_cnt + alive.float()
        s = torch.where(alive[:, None], nxt, s)
        upright = (nxt[:, 2] > -thresh_div) & (nxt[:, 2] < thresh_div)
        alive = alive & upright
    mean_angle = torch.where(ang_cnt > 0, ang_sum / ang_cnt.clamp(min=1), torch.full((n,), 100.0))
    return dict(states=states, actions=actions, success=n_steps - 1, n_steps=n_steps, mean_angle=mean_angle,
                vel_sum=vel_sum)


# --------------------------------------------------------------------------------------------
# N4  reference tables from raw trajectory files: load_prepare_trajectory
#     (neural_control/trajectory/generate_trajectory.py:566-603), q_funcs.quaternion_to_euler (:38-41).
#     The Euler angles come from pyquaternion (Quaternion.yaw_pitch_roll), which is NOT installed here and not
#     pinned by the reference: its published formula is restated (PARITY UNPINNED for these three columns; the
#     sub-sampling, the column layout and the speed scaling are pinned on the reference's own code run with a
#     pyquaternion stand-in implementing the same formula, tests/golden/ref_table.npz).
# --------------------------------------------------------------------------------------------
def quaternion_to_euler(q):
    """q (..., 4) = w, x, y, z -> (..., 3) = roll, pitch, yaw (float64 numpy)"""
    import numpy as np
    q = np.asarray(q, dtype=np.float64)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    yaw = np.arctan2(2 * (w * z - x * y), 1 - 2 * (y ** 2 + z ** 2))
    pitch = np.arcsin(2 * (w * y + z * x))
    roll = np.arctan2(2 * (w * x - y * z), 1 - 2 * (x ** 2 + y ** 2))
    return np.stack((roll, pitch, yaw), axis=-1)


def reference_table(traj, dt, speed_factor, z_offset=0.0):
    """raw trajectory (T, >=10) -> (ceil(T/nth), 9) = [pos, euler * speed, vel * speed * 2] (+ z_offset on z:
    Random.__init__, random_traj.py:35)"""
    import numpy as np
    nth = int(dt / 0.01 * speed_factor)
    assert np.isclose(nth, dt / 0.01 * speed_factor)
    taken = np.asarray(traj, dtype=np.float64)[::nth]
    out = np.hstack((taken[:, :3], quaternion_to_euler(taken[:, 3:7]) * speed_factor,
                     taken[:, 7:10] * speed_factor * 2))
    out[:, 2] += z_offset
    return out


def polynomial_points(coef, rot, start, x_range=20.0, dist_points=0.025, hover_steps=50, x_start=1.0):
    """Polynomial.__init__ with random_polynomial (neural_control/trajectory/polynomial.py:36-47, 84-125) for GIVEN
    fit coefficients (np.poly1d order) and rotation: march in steps of dist_points of arc length, [x, 0, y] @ rot,
    shift to `start`, hover padding.  float64 numpy, one trajectory."""
    import numpy as np
    coef = np.asarray(coef, dtype=np.float64)
    degree = len(coef) - 1
    poly = np.poly1d(coef)
    grad = lambda x: np.sum([(degree - i) * coef[i] * x ** (degree - i - 1) for i in range(degree)])   # noqa: E731
    x, x_final = float(x_start), float(x_start) + x_range
    pts = [[x, poly(x)]]
    while x < x_final:
        vec = np.array([1, grad(x)])
        x = x + (vec / np.linalg.norm(vec) * dist_points)[0]
        pts.append([x, poly(x)])
    pts = np.array(pts)
    p3 = np.stack((pts[:, 0], np.zeros(len(pts)), pts[:, 1]), axis=1) @ np.asarray(rot, dtype=np.float64)
    if start is not None:
        p3 = p3 - p3[0] + np.asarray(start, dtype=np.float64)
    return np.vstack([np.repeat(p3[:1], hover_steps, 0), p3, np.repeat(p3[-1:], hover_steps, 0)])
