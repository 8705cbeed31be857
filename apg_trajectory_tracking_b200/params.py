"""Physical constants of the three systems and their packing into the flat fp32 array the kernels take.

Values are those of the reference's neural_control/dynamics/config_quad.json, config_fixed_wing.json and
config_cartpole.json (+ the overrides the reference applies in cartpole_dynamics.py:34-36); every dynamics class
accepts the reference's ``modified_params`` dict to override them (dynamics-mismatch experiments).
The index layout must match the enums QuadC / WingC / CartC in csrc/apg_math.cuh.
"""
import numpy as np

MAX_PHYS = 48

QUAD_DEFAULTS = {
    "mass": 0.723, "rotational_drag": [0, 0, 0], "translational_drag": [0, 0, 0], "arm_length": 0.31,
    "frame_inertia": [4.5, 4.5, 7.0], "gravity": [0, 0, -9.81], "kinv_ang_vel_tau": [16.6, 16.6, 5.0],
}
WING_DEFAULTS = {
    "mass": 1.01, "I_xx": 0.04766, "I_yy": 0.05005, "I_zz": 0.09558, "I_xz": -0.00105, "rho": 1.225,
    "S": 0.276, "c": 0.185, "b": 1.54, "g": 9.81, "CL0": 0.39, "CL_alpha": 4.5321, "CL_q": 0.318,
    "CL_del_e": 0.527, "CD0": 0.0765, "CD_alpha": 0.3346, "CD_q": 0.354, "CD_del_e": 0.004, "CY0": 0.0,
    "CY_beta": -0.033, "CY_p": -0.1, "CY_r": 0.039, "CY_del_a": 0.0, "CY_del_r": 0.225, "Cl0": 0.0,
    "Cl_beta": -0.081, "Cl_p": -0.529, "Cl_r": 0.159, "Cl_del_a": -0.453, "Cl_del_r": 0.005, "Cm0": 0.02,
    "Cm_alpha": -1.4037, "Cm_q": -0.1324, "Cm_del_e": -0.4236, "Cn0": 0.0, "Cn_beta": 0.189, "Cn_p": -0.083,
    "Cn_r": -0.948, "Cn_del_a": -0.041, "Cn_del_r": -0.077, "epsilon": 0.16534698176788384,
}
CARTPOLE_DEFAULTS = {
    "masscart": 1.0, "masspole": 0.1, "length": 0.5, "max_force_mag": 30.0, "muc": 0.0005, "mup": 0.000002,
    "wind": 0.0, "vel_drag": 0.0, "contact": 0.0, "delay": 0.0,
}

_WING_ORDER = ["mass", "I_xx", "I_yy", "I_zz", "I_xz", "rho", "S", "c", "b", "g",
               "CL0", "CL_alpha", "CL_q", "CL_del_e", "CD0", "CD_alpha", "CD_q", "CD_del_e",
               "CY0", "CY_beta", "CY_p", "CY_r", "CY_del_a", "CY_del_r",
               "Cl0", "Cl_beta", "Cl_p", "Cl_r", "Cl_del_a", "Cl_del_r",
               "Cm0", "Cm_alpha", "Cm_q", "Cm_del_e",
               "Cn0", "Cn_beta", "Cn_p", "Cn_r", "Cn_del_a", "Cn_del_r", "epsilon"]


def _pad(vals):
    out = np.zeros(MAX_PHYS, dtype=np.float32)
    out[:len(vals)] = np.asarray(vals, dtype=np.float32)
    return out


def quad_cfg(modified_params=None):
    cfg = dict(QUAD_DEFAULTS)
    cfg.update(modified_params or {})
    return cfg


def quad_phys(modified_params=None):
    """[mass, Jx,Jy,Jz, Kx,Ky,Kz, g(3), translational_drag(3), rotational_drag(3)];
    J = mass/12 * arm_length^2 * frame_inertia, rounded to fp32 like the reference's torch tensors
    (quad_dynamics_base.py:32-49)."""
    cfg = quad_cfg(modified_params)
    inertia = (cfg["mass"] / 12.0 * cfg["arm_length"] ** 2 * np.asarray(cfg["frame_inertia"], dtype=np.float64))
    vals = [cfg["mass"], *inertia.astype(np.float32), *cfg["kinv_ang_vel_tau"], *cfg["gravity"],
            *cfg["translational_drag"], *cfg["rotational_drag"]]
    return _pad(vals)


def wing_cfg(modified_params=None):
    cfg = dict(WING_DEFAULTS)
    cfg.update(modified_params or {})
    return cfg


def wing_phys(modified_params=None):
    cfg = wing_cfg(modified_params)
    return _pad([cfg[k] for k in _WING_ORDER])


def cartpole_cfg(modified_params=None):
    cfg = dict(CARTPOLE_DEFAULTS)
    cfg.update(modified_params or {})
    cfg["friction"] = 0.5                                            # cartpole_dynamics.py:34
    cfg["total_mass"] = cfg["masspole"] + cfg["masscart"]
    cfg["polemass_length"] = cfg["masspole"] * cfg["length"]
    return cfg


def cartpole_phys(modified_params=None):
    cfg = cartpole_cfg(modified_params)
    return _pad([cfg["masscart"], cfg["masspole"], cfg["length"], cfg["max_force_mag"], cfg["friction"]])


PHYS = {"quad": quad_phys, "wing": wing_phys, "cartpole": cartpole_phys}
SYSTEM_ID = {"quad": 0, "wing": 1, "cartpole": 2}
STATE_DIM = {"quad": 12, "wing": 12, "cartpole": 4}
ACTION_DIM = {"quad": 4, "wing": 4, "cartpole": 1}
