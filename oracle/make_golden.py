"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference checkout does not exist on the GPU box):

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

The reference is imported from its checkout with its optional third-party deps (casadi, gym, matplotlib,
pyglet, pyquaternion, ruamel) stubbed in ``sys.modules`` (they are only touched at import / construction
time on this path).  Everything saved here is an output of the reference's own functions/methods:

  steps.npz        single dynamics steps: the ``__main__`` vectors of the three dynamics files (KAT-1/2/3,
                   SURVEY.md 8c) + seeded random batches, with vector-Jacobian products from reference autograd
  conc_*.npz       concurrent train steps through TrainDrone / TrainFixedWing.train_controller_model and the
                   TrainCartpole.run_epoch body: loss, actions, states, all parameter gradients; both with the
                   shipped trained_models (KAT-4/5/6) and with seeded default-initialised nets
  rec_*.npz        autoregressive / LSTM forward through TrainDrone.train_recurrent_model (its backward()
                   raises in the reference, so only loss/actions/states exist), plus state_preprocessing
"""
import argparse
import os
import sys
from unittest.mock import MagicMock

import numpy as np


def import_reference(ref_root):
    for m in ['casadi', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation', 'mpl_toolkits',
              'mpl_toolkits.mplot3d', 'gym', 'gym.utils', 'gym.spaces', 'pyglet', 'pyglet.gl',
              'pyquaternion', 'ruamel', 'ruamel.yaml']:
        sys.modules[m] = MagicMock()

    class _Env:
        pass
    sys.modules['gym'].Env = _Env
    sys.path[:0] = [ref_root, os.path.join(ref_root, 'scripts')]


def npify(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        out[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return out


def eval_table(seed, rows, dt, speed):
    """synthetic stand-in for one file of data/traj_data_1 as load_prepare_trajectory returns it:
    [pos, euler (0), vel] of a degree-5 polynomial per axis, sampled every dt"""
    import math
    rng = np.random.default_rng(seed)
    c = np.zeros((3, 6))
    c[:, 1] = rng.uniform(-speed, speed, 3)
    for i in range(2, 6):
        c[:, i] = rng.uniform(-0.3 * speed, 0.3 * speed, 3) / math.factorial(i)
    t = np.arange(rows) * dt
    pos = np.stack([sum(c[a, i] * t ** i for i in range(6)) for a in range(3)], 1)
    vel = np.stack([sum(i * c[a, i] * t ** (i - 1) for i in range(1, 6)) for a in range(3)], 1)
    return np.hstack((pos, np.zeros((rows, 3)), vel))


def make_eval_golden(args):
    """eval_rand.npz: QuadEvaluator.follow_trajectory("rand") of the unmodified reference (scripts/evaluate_drone.py)
    with the shipped model_quad.  The only substitution is the trajectory FILE: data/traj_data_1 is not part of the
    checkout, so load_prepare_trajectory (the file reader) is replaced by a function returning the synthetic tables
    saved next to the outputs; Random.__init__ / get_ref_traj / project_on_ref, NetworkWrapper, QuadDataset,
    QuadRotorEnvBase.step and FlightmareDynamics run as they are."""
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    _np_random = lambda seed=None: (np.random.RandomState(0), 0)      # gym is stubbed: give the env a real RNG
    sys.modules['gym.utils'].seeding.np_random = _np_random
    sys.modules['gym'].utils.seeding.np_random = _np_random
    import neural_control.trajectory.random_traj as RT
    import evaluate_drone as ED
    from neural_control.environments.drone_env import QuadRotorEnvBase
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from neural_control.controllers.network_wrapper import NetworkWrapper
    from neural_control.dataset import QuadDataset
    net = torch.load('trained_models/quad/current_model/model_quad', weights_only=False)
    net.eval()
    h, dt = 10, 0.1
    ds = QuadDataset.__new__(QuadDataset)            # no sampling (needs the absent data files); prepare_data only
    ds.get_and_add_eval_data = lambda st, rf, add_to_dataset=False: QuadDataset.prepare_data(ds, st, rf)
    out = {}
    # (name, table seed, rows, speed, steps, test_time, thresh_div, thresh_stable)
    runs = [("gentle", 1, 120, 0.25, 80, 0, 1.0, 1.0), ("fast_reset", 2, 120, 1.0, 80, 0, 1.0, 1.0),
            ("fast_stop", 2, 120, 1.0, 80, 1, 1.0, 1.0), ("short_table", 3, 30, 0.3, 45, 0, 1.0, 1.0),
            ("tight", 4, 100, 0.6, 60, 0, 0.3, 0.05)]
    for name, seed, rows, speed, steps, test_time, tdiv, tstab in runs:
        table = eval_table(seed, rows, dt, speed)
        RT.load_prepare_trajectory = lambda base_dir, dt_, speed_factor, test=False, _t=table: _t.copy()
        env = QuadRotorEnvBase(FlightmareDynamics(), dt)
        ctrl = NetworkWrapper(net, ds, horizon=h, dt=dt)
        ev = ED.QuadEvaluator(ctrl, env, ref_length=h, dt=dt, test_time=test_time, speed_factor=0.4,
                              train_mode="concurrent")
        ref_traj, drone_traj, div, acts = ev.follow_trajectory("rand", max_nr_steps=steps, thresh_stable=tstab,
                                                               thresh_div=tdiv)
        tab = table.copy()
        tab[:, 2] += 3                                # Random.__init__ (random_traj.py:35)
        out[f"{name}_table"] = tab
        out[f"{name}_cfg"] = np.array([steps, test_time, tdiv, tstab, h, dt], dtype=np.float64)
        out[f"{name}_ref_traj"] = np.asarray(ref_traj)
        out[f"{name}_states"] = np.asarray(drone_traj)
        out[f"{name}_div"] = np.asarray(div)
        out[f"{name}_actions"] = np.asarray(acts)[:, 0]          # the applied action (evaluate_drone.py:154-155)
        print("eval", name, "steps taken", len(div), "mean div %.4f" % np.mean(div), "resets/stops",
              int(np.sum(np.asarray(div) > tdiv)))
    out["run_names"] = np.array([r[0] for r in runs])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_rand.npz"), **out)


def make_eval_lstm_golden(args):
    """eval_rand_lstm.npz: QuadEvaluator.follow_trajectory("rand") of the unmodified reference with an LSTM_NEW policy
    (models/rnn.py; seeded default init - the reference ships no trained LSTM), train_mode="LSTM": the applied action is
    the net's 4 outputs, the hidden / cell state drawn at evaluator construction (evaluate_drone.py:55-57) is carried
    through the run.  Same trajectory-file substitution as make_eval_golden."""
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    _np_random = lambda seed=None: (np.random.RandomState(0), 0)
    sys.modules['gym.utils'].seeding.np_random = _np_random
    sys.modules['gym'].utils.seeding.np_random = _np_random
    import neural_control.trajectory.random_traj as RT
    import evaluate_drone as ED
    from neural_control.environments.drone_env import QuadRotorEnvBase
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from neural_control.controllers.network_wrapper import NetworkWrapper
    from neural_control.dataset import QuadDataset
    from neural_control.models.rnn import LSTM_NEW
    h, dt = 10, 0.1
    torch.manual_seed(7)
    net = LSTM_NEW(15, h, 9, 4)
    with torch.no_grad():
        net.fc_out.bias += torch.tensor([0.3, 0.0, 0.0, 0.0])   # hover-ish thrust so that runs last a few steps
    net.eval()
    ds = QuadDataset.__new__(QuadDataset)
    ds.get_and_add_eval_data = lambda st, rf, add_to_dataset=False: QuadDataset.prepare_data(ds, st, rf)
    out = {}
    for i, (_, p) in enumerate(net.named_parameters()):
        out[f"param_{i}"] = p.detach().numpy().copy()
    out["param_names"] = np.array([n for n, _ in net.named_parameters()])
    runs = [("gentle", 1, 120, 0.25, 60, 0, 1.0, 1.0), ("fast_stop", 2, 120, 1.0, 60, 1, 1.0, 1.0),
            ("loose", 4, 100, 0.6, 50, 0, 3.0, 0.4)]
    for name, seed, rows, speed, steps, test_time, tdiv, tstab in runs:
        table = eval_table(seed, rows, dt, speed)
        RT.load_prepare_trajectory = lambda base_dir, dt_, speed_factor, test=False, _t=table: _t.copy()
        env = QuadRotorEnvBase(FlightmareDynamics(), dt)
        ctrl = NetworkWrapper(net, ds, horizon=h, dt=dt)
        torch.manual_seed(100 + seed)
        ev = ED.QuadEvaluator(ctrl, env, ref_length=h, dt=dt, test_time=test_time, speed_factor=0.4,
                              train_mode="LSTM")
        out[f"{name}_h0"] = net.hidden_state.detach().numpy().copy()
        out[f"{name}_c0"] = net.cell_state.detach().numpy().copy()
        ref_traj, drone_traj, div, acts = ev.follow_trajectory("rand", max_nr_steps=steps, thresh_stable=tstab,
                                                               thresh_div=tdiv)
        tab = table.copy()
        tab[:, 2] += 3
        out[f"{name}_table"] = tab
        out[f"{name}_cfg"] = np.array([steps, test_time, tdiv, tstab, h, dt], dtype=np.float64)
        out[f"{name}_states"] = np.asarray(drone_traj)
        out[f"{name}_div"] = np.asarray(div)
        out[f"{name}_actions"] = np.asarray(acts)                # train_mode LSTM: the action IS the output (:156-157)
        out[f"{name}_h1"] = net.hidden_state.detach().numpy().copy()
        out[f"{name}_c1"] = net.cell_state.detach().numpy().copy()
        print("eval lstm", name, "steps taken", len(div), "mean div %.4f" % np.mean(div), "resets/stops",
              int(np.sum(np.asarray(div) > tdiv)))
    out["run_names"] = np.array([r[0] for r in runs])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_rand_lstm.npz"), **out)


def make_selfplay_golden(args):
    """eval_selfplay.npz: the self-play feed of the evaluation (NetworkWrapper.predict_actions,
    controllers/network_wrapper.py:42-52 -> DroneDataset.get_and_add_eval_data, dataset.py:98-119).  ONE controller
    (its action counter runs on over the runs) evaluates three tables one after the other with take_every_x = 7 into a
    dataset with 3 sampled rows and 5 self-play slots (the ring wraps).  Saved: the raw (state, window) pairs handed to
    the dataset for the kept calls, and the dataset tensors afterwards.  Substitutions as in make_eval_golden (the
    trajectory FILE reader); the dataset is allocated without its sampler (needs the absent data files)."""
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    _np_random = lambda seed=None: (np.random.RandomState(0), 0)
    sys.modules['gym.utils'].seeding.np_random = _np_random
    sys.modules['gym'].utils.seeding.np_random = _np_random
    import neural_control.trajectory.random_traj as RT
    import evaluate_drone as ED
    from neural_control.environments.drone_env import QuadRotorEnvBase
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from neural_control.controllers.network_wrapper import NetworkWrapper
    from neural_control.dataset import QuadDataset
    net = torch.load('trained_models/quad/current_model/model_quad', weights_only=False)
    net.eval()
    h, dt, take, n_sampled, n_slots = 10, 0.1, 7, 3, 5
    ds = QuadDataset.__new__(QuadDataset)
    ds.num_sampled_states, ds.num_self_play, ds.eval_counter = n_sampled, n_slots, 0
    tot = n_sampled + n_slots
    ds.normed_states, ds.states = torch.zeros(tot, 15), torch.zeros(tot, 12)
    ds.in_ref_states, ds.ref_states = torch.zeros(tot, h, 9), torch.zeros(tot, h, 9)
    kept_states, kept_refs = [], []
    real = ds.get_and_add_eval_data

    def recording(st, rf, add_to_dataset=False):
        if add_to_dataset:
            kept_states.append(np.array(st, dtype=np.float64).copy())
            kept_refs.append(np.array(rf, dtype=np.float64).copy())
        return real(st, rf, add_to_dataset=add_to_dataset)
    ds.get_and_add_eval_data = recording
    ctrl = NetworkWrapper(net, ds, horizon=h, dt=dt, take_every_x=take)
    out = {"cfg": np.array([h, dt, take, n_sampled, n_slots], dtype=np.float64)}
    # (name, table seed, rows, speed, steps, thresh_div, thresh_stable): tracking, resets, end-of-table padding
    runs = [("a", 5, 60, 0.3, 32, 1.0, 1.0), ("b", 2, 120, 1.0, 50, 1.0, 1.0), ("c", 3, 30, 0.3, 41, 1.0, 1.0)]
    for name, seed, rows, speed, steps, tdiv, tstab in runs:
        table = eval_table(seed, rows, dt, speed)
        RT.load_prepare_trajectory = lambda base_dir, dt_, speed_factor, test=False, _t=table: _t.copy()
        env = QuadRotorEnvBase(FlightmareDynamics(), dt)
        ev = ED.QuadEvaluator(ctrl, env, ref_length=h, dt=dt, test_time=0, speed_factor=0.4, train_mode="concurrent")
        _, drone_traj, div, _ = ev.follow_trajectory("rand", max_nr_steps=steps, thresh_stable=tstab, thresh_div=tdiv)
        tab = table.copy()
        tab[:, 2] += 3
        out[f"{name}_table"] = tab
        out[f"{name}_cfg"] = np.array([steps, tdiv, tstab], dtype=np.float64)
        out[f"{name}_states"] = np.asarray(drone_traj)
        out[f"{name}_div"] = np.asarray(div)
        print("selfplay", name, "steps", len(div), "resets", int(np.sum(np.asarray(div) > tdiv)), "kept so far",
              len(kept_states))
    out["run_names"] = np.array([r[0] for r in runs])
    out["kept_states"] = np.asarray(kept_states).reshape(len(kept_states), 12)
    out["kept_refs"] = np.asarray(kept_refs).reshape(len(kept_refs), h, 9)
    out["action_counter"] = np.array([ctrl.action_counter])
    out["eval_counter"] = np.array([ds.eval_counter])
    out["ds_normed_states"], out["ds_states"] = ds.normed_states, ds.states
    out["ds_in_ref_states"], out["ds_ref_states"] = ds.in_ref_states, ds.ref_states
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_selfplay.npz"), **npify(out))


def make_ref_table_golden(args):
    """ref_table.npz: load_prepare_trajectory (neural_control/trajectory/generate_trajectory.py:566-603) of the
    unmodified reference on synthetic raw trajectory files written to a temporary data directory (the real
    data/traj_data_1 is not part of the checkout).  pyquaternion is not installed: `pyquaternion.Quaternion` is a
    stand-in with the published yaw_pitch_roll formula (normalise; yaw = atan2(2(wz - xy), 1 - 2(y^2 + z^2)),
    pitch = asin(2(wy + zx)), roll = atan2(2(wx - yz), 1 - 2(x^2 + y^2))), so the Euler columns pin the CALL SITE
    (argument order w x y z, output order roll pitch yaw, scaling), not pyquaternion itself."""
    import tempfile

    class Quaternion:
        def __init__(self, w, x, y, z):
            q = np.array([w, x, y, z], dtype=np.float64)
            self.q = q / np.linalg.norm(q)

        @property
        def yaw_pitch_roll(self):
            w, x, y, z = self.q
            return (np.arctan2(2 * (w * z - x * y), 1 - 2 * (y ** 2 + z ** 2)), np.arcsin(2 * (w * y + z * x)),
                    np.arctan2(2 * (w * x - y * z), 1 - 2 * (x ** 2 + y ** 2)))
    import neural_control.trajectory.q_funcs as QF
    import neural_control.trajectory.generate_trajectory as GT
    QF.pyquaternion.Quaternion = Quaternion
    rng = np.random.default_rng(11)
    out = {}
    # (name, raw rows, dt, speed_factor)
    cases = [("a", 401, 0.1, 0.4), ("b", 203, 0.05, 0.6), ("c", 77, 0.1, 1.0)]
    for name, T, dt, speed in cases:
        raw = np.zeros((T, 12))
        t = np.arange(T) * 0.01
        raw[:, :3] = np.stack((np.sin(t), np.cos(0.7 * t), 1 + 0.3 * t), 1) + rng.normal(0, 0.01, (T, 3))
        q = rng.normal(0, 1, (T, 4)) * np.array([0.2, 0.3, 0.3, 0.2]) + np.array([1.0, 0, 0, 0])
        raw[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True) * rng.uniform(0.9, 1.1, (T, 1))   # not quite unit
        raw[:, 7:10] = rng.normal(0, 1, (T, 3))
        raw[:, 10:] = rng.normal(0, 1, (T, 2))
        raw = raw.astype(np.float32).astype(np.float64)          # exactly representable in the kernel's float32
        with tempfile.TemporaryDirectory() as d:
            os.makedirs(os.path.join(d, "train"))
            np.save(os.path.join(d, "train", "traj_0.npy"), raw)
            table = GT.load_prepare_trajectory(d, dt, speed, test=False)
        out[f"{name}_raw"], out[f"{name}_table"] = raw.astype(np.float32), table
        out[f"{name}_cfg"] = np.array([dt, speed], dtype=np.float64)
        print("ref table", name, raw.shape, "->", table.shape)
    out["case_names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(args.out, "ref_table.npz"), **out)


def make_poly_traj_golden(args):
    """poly_traj.npz: Polynomial(drone_state, ...) of the unmodified reference (neural_control/trajectory/
    polynomial.py) with its random_polynomial branch; the fit coefficients and the random rotation it drew are
    recorded (np.polyfit / special_ortho_group.rvs wrapped) next to the resulting reference points."""
    import neural_control.trajectory.polynomial as PT
    out = {}
    rec = {}
    real_fit, real_rvs = np.polyfit, PT.special_ortho_group.rvs

    def fit(x, y, deg):
        rec["coef"] = real_fit(x, y, deg)
        return rec["coef"]

    def rvs(dim):
        rec["rot"] = real_rvs(dim)
        return rec["rot"]
    PT.np.polyfit = fit
    PT.special_ortho_group.rvs = rvs
    # (name, seed, start, x_range, degree, max_drone_dist, horizon, hover_steps)
    cases = [("a", 0, [0.5, -1.0, 2.0], 20, 5, 0.25, 10, 50), ("b", 1, [0.0, 0.0, 3.0], 6, 5, 0.5, 10, 5),
             ("c", 2, [-2.0, 1.0, 1.0], 10, 3, 0.25, 5, 2)]
    try:
        for name, seed, start, x_range, degree, mdd, h, hover in cases:
            np.random.seed(seed)
            tr = PT.Polynomial(np.array(start + [0.0] * 9), max_drone_dist=mdd, horizon=h, hover_steps=hover,
                               x_range=x_range, degree=degree, dt=0.05)
            out[f"{name}_coef"], out[f"{name}_rot"] = rec["coef"].copy(), rec["rot"].copy()
            out[f"{name}_start"] = np.array(start)
            out[f"{name}_cfg"] = np.array([x_range, degree, mdd, h, hover], dtype=np.float64)
            out[f"{name}_points"] = np.asarray(tr.reference)
            print("poly traj", name, "rows", tr.ref_len)
    finally:
        PT.np.polyfit = real_fit
        PT.special_ortho_group.rvs = real_rvs
    out["case_names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(args.out, "poly_traj.npz"), **out)


def make_wing_eval_golden(args):
    """eval_wing.npz: FixedWingEvaluator.fly_to_point (scripts/evaluate_fixed_wing.py) of the unmodified reference with
    the shipped model_wing and its config (mean / std / horizon / dt): trajectories, divergences and the target
    errors for a few target lists and thresholds (resets, stop at divergence, several targets, step limit)."""
    import json
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    import evaluate_fixed_wing as EF
    from neural_control.environments.wing_env import SimpleWingEnv
    from neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from neural_control.controllers.network_wrapper import FixedWingNetWrapper
    from neural_control.dataset import WingDataset
    net = torch.load('trained_models/wing/current_model/model_wing', weights_only=False)
    net.eval()
    cfg = json.load(open('trained_models/wing/current_model/config.json'))
    ds = WingDataset.__new__(WingDataset)              # no sampling; prepare_data only
    ds.dt, ds.horizon = cfg["delta_t"], cfg["horizon"]
    ds.mean, ds.std = torch.tensor(cfg["mean"]).float(), torch.tensor(cfg["std"]).float()
    ds.get_and_add_eval_data = lambda st, rf, add_to_dataset=False: WingDataset.prepare_data(ds, st, rf)
    out = {"mean": np.array(cfg["mean"]), "std": np.array(cfg["std"]),
           "cfg": np.array([cfg["horizon"], cfg["delta_t"], cfg["dt"]], dtype=np.float64)}
    for i, p in enumerate(net.parameters()):
        out[f"param_{i}"] = p.detach()
    # (name, targets, max_steps, thresh_div, thresh_stable, test_time)
    runs = [("one_target", [[50., -3., 3.]], 300, 4.0, 0.4, 0),
            ("two_targets", [[30., 2., -2.], [60., -4., 1.]], 300, 4.0, 0.4, 0),
            ("tight_reset", [[50., 8., -8.]], 200, 0.25, 0.4, 0),
            ("tight_stop", [[50., 8., -8.]], 200, 0.25, 0.4, 1),
            ("unstable", [[40., 12., 10.]], 150, 10.0, 0.12, 0),
            ("step_limit", [[50., 1., 1.]], 40, 4.0, 0.4, 0)]
    for name, targets, steps, tdiv, tstab, test_time in runs:
        ctrl = FixedWingNetWrapper(net, ds, horizon=cfg["horizon"])
        env = SimpleWingEnv(FixedWingDynamics(), cfg["dt"])
        ev = EF.FixedWingEvaluator(ctrl, env, dt=cfg["dt"], horizon=cfg["horizon"], thresh_div=tdiv,
                                   thresh_stable=tstab, test_time=test_time)
        tg = np.array(targets)
        traj = ev.fly_to_point(tg, max_steps=steps, return_traj=True)
        div_target, div_linear = ev.fly_to_point(tg, max_steps=steps)
        out[f"{name}_targets"] = tg
        out[f"{name}_cfg"] = np.array([steps, test_time, tdiv, tstab], dtype=np.float64)
        out[f"{name}_traj"] = traj                      # rows [state (12), applied action (4)]
        out[f"{name}_div_target"] = np.asarray(div_target, dtype=np.float64)
        out[f"{name}_div_linear"] = np.asarray(div_linear, dtype=np.float64)
        print("wing eval", name, "steps", len(div_linear), "div_target", np.round(div_target, 4),
              "max div_linear %.3f" % np.max(div_linear))
    out["run_names"] = np.array([r[0] for r in runs])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_wing.npz"), **npify(out))


def make_wing_selfplay_golden(args):
    """eval_wing_selfplay.npz: the self-play feed of the fixed-wing evaluation (FixedWingNetWrapper.predict_actions,
    controllers/network_wrapper.py:81-98 -> DroneDataset.get_and_add_eval_data): ONE controller (take_every_x = 11,
    counter running on) flies three target lists one after the other; the raw (state, target) pairs handed to the
    dataset for the kept calls are recorded, and the dataset's slots afterwards (2 sampled rows + 4 slots)."""
    import json
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    import evaluate_fixed_wing as EF
    from neural_control.environments.wing_env import SimpleWingEnv
    from neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from neural_control.controllers.network_wrapper import FixedWingNetWrapper
    from neural_control.dataset import WingDataset
    net = torch.load('trained_models/wing/current_model/model_wing', weights_only=False)
    net.eval()
    cfg = json.load(open('trained_models/wing/current_model/config.json'))
    h = cfg["horizon"]
    take, n_sampled, n_slots = 11, 2, 4
    ds = WingDataset.__new__(WingDataset)
    ds.dt, ds.horizon = cfg["delta_t"], h
    ds.mean, ds.std = torch.tensor(cfg["mean"]).float(), torch.tensor(cfg["std"]).float()
    ds.num_sampled_states, ds.num_self_play, ds.eval_counter = n_sampled, n_slots, 0
    tot = n_sampled + n_slots
    ds.normed_states, ds.states = torch.zeros(tot, 9), torch.zeros(tot, 12)
    ds.in_ref_states, ds.ref_states = torch.zeros(tot, 3), torch.zeros(tot, h, 3)
    kept_states, kept_targets = [], []
    real = ds.get_and_add_eval_data

    def recording(st, rf, add_to_dataset=False):
        if add_to_dataset:
            kept_states.append(np.array(st, dtype=np.float64).copy())
            kept_targets.append(np.array(rf, dtype=np.float64).copy())
        return real(st, rf, add_to_dataset=add_to_dataset)
    ds.get_and_add_eval_data = recording
    ctrl = FixedWingNetWrapper(net, ds, horizon=h, take_every_x=take)
    out = {"cfg": np.array([h, cfg["delta_t"], cfg["dt"], take, n_sampled, n_slots], dtype=np.float64)}
    # (name, targets, max_steps, thresh_div, thresh_stable)
    runs = [("a", [[30., 2., -2.], [60., -4., 1.]], 300, 4.0, 0.4), ("b", [[50., -3., 3.]], 300, 4.0, 0.4),
            ("c", [[50., 8., -8.]], 120, 0.25, 0.4)]
    for name, targets, steps, tdiv, tstab in runs:
        env = SimpleWingEnv(FixedWingDynamics(), cfg["dt"])
        ev = EF.FixedWingEvaluator(ctrl, env, dt=cfg["dt"], horizon=h, thresh_div=tdiv, thresh_stable=tstab,
                                   test_time=0)
        div_target, div_linear = ev.fly_to_point(np.array(targets), max_steps=steps)
        out[f"{name}_targets"] = np.array(targets)
        out[f"{name}_cfg"] = np.array([steps, tdiv, tstab], dtype=np.float64)
        out[f"{name}_n_steps"] = np.array([len(div_linear)])
        print("wing selfplay", name, "steps", len(div_linear), "kept so far", len(kept_states))
    out["run_names"] = np.array([r[0] for r in runs])
    out["kept_states"] = np.asarray(kept_states).reshape(len(kept_states), 12)
    out["kept_targets"] = np.asarray(kept_targets).reshape(len(kept_targets), 3)
    out["action_counter"] = np.array([ctrl.action_counter])
    out["eval_counter"] = np.array([ds.eval_counter])
    out["ds_states"], out["ds_normed_states"] = ds.states, ds.normed_states
    out["ds_in_ref_states"], out["ds_ref_states"] = ds.in_ref_states, ds.ref_states
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_wing_selfplay.npz"), **npify(out))


def make_cartpole_eval_golden(args):
    """eval_cartpole.npz: Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py) of the unmodified reference
    with the shipped model_cartpole: the state after every env._step and the success indices.  The reference always
    starts from the zero state (:101-116); the runs with other start states and thresholds patch ONLY the start
    state the reset leaves behind."""
    import torch
    cwd = os.getcwd()
    os.chdir(args.ref)
    import evaluate_cartpole as EC
    from neural_control.environments.cartpole_env import CartPoleEnv
    from neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from neural_control.controllers.network_wrapper import CartpoleWrapper
    net = torch.load('trained_models/cartpole/current_model/model_cartpole', weights_only=False)
    net.eval()
    out = {}
    for i, p in enumerate(net.parameters()):
        out[f"param_{i}"] = p.detach()
    # (name, start state or None (= the reference's zero start), max_steps, thresh_div, burn_in)
    runs = [("zero_start", None, 80, 0.21, 10), ("tilted", [0.3, 0.2, 0.12, -0.3], 80, 0.21, 10),
            ("falls", [0.4, -1.232, 0.068, 1.425], 60, 0.21, 5), ("tight", [-0.3, 1.13, 0.088, -0.595], 60, 0.12, 50),
            ("falls_at_once", [0.0, 0.5, 0.19, 0.9], 60, 0.21, 5)]
    for name, start, steps, tdiv, burn in runs:
        env = CartPoleEnv(CartpoleDynamics(), 0.05, thresh_div=tdiv)
        ev = EC.Evaluator(CartpoleWrapper(net, horizon=10, action_dim=1), env)
        if start is not None:
            ev.initialize_straight = 0
            env._reset_upright = lambda _e=env, _s=start: setattr(_e, "state", np.array(_s, dtype=np.float64))
        log = []
        orig = env._step

        def logged(*a, _orig=orig, _log=log, **k):
            r = _orig(*a, **k)
            _log.append(np.array(r, dtype=np.float64).copy())
            return r
        env._step = logged
        np.random.seed(0)
        succ, vel = ev.evaluate_in_environment(nr_iters=1, max_steps=steps, render=False, burn_in_steps=burn,
                                               return_success=1)
        out[f"{name}_init"] = np.zeros(4) if start is None else np.array(start, dtype=np.float64)
        out[f"{name}_cfg"] = np.array([steps, tdiv, burn], dtype=np.float64)
        out[f"{name}_states"] = np.array(log)
        out[f"{name}_success"] = np.asarray(succ)
        out[f"{name}_vel"] = np.asarray(vel)
        print("cartpole eval", name, "steps", len(log), "success", succ)
    out["run_names"] = np.array([r[0] for r in runs])
    os.chdir(cwd)
    np.savez_compressed(os.path.join(args.out, "eval_cartpole.npz"), **npify(out))


def make_learnt_golden(args):
    """learnt_dyn.npz: LearntDynamics.forward (neural_control/dynamics/quad_dynamics_trained.py) of the unmodified
    reference with seeded NON-zero parameters (the reference initialises the residual MLP with zeros, which leaves it
    without gradient forever), its vector-Jacobian products w.r.t. state, action and every parameter, and one
    TrainBase.train_dynamics_model loss / gradient against a FlightmareDynamics with modified parameters."""
    import torch
    from neural_control.dynamics.quad_dynamics_trained import LearntDynamics
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    torch.manual_seed(7)
    out = {}
    for tag, drag in (("a", [0.0, 0.0, 0.0]), ("b", [0.02, -0.01, 0.03])):
        d1 = LearntDynamics({"rotational_drag": drag} if tag == "b" else {})
        with torch.no_grad():
            d1.linear_at.add_(0.1 * torch.randn(4, 4))
            d1.linear_state_1.weight.copy_(0.3 * torch.randn(64, 16))
            d1.linear_state_1.bias.copy_(0.1 * torch.randn(64))
            d1.linear_state_2.weight.copy_(0.1 * torch.randn(12, 64))
            d1.linear_state_2.bias.copy_(0.05 * torch.randn(12))
        n, dt = 37, 0.1 if tag == "a" else 0.05
        s = (0.4 * torch.randn(n, 12)).requires_grad_(True)
        a = torch.rand(n, 4).requires_grad_(True)
        cot = torch.randn(n, 12)
        o = d1(s, a, dt)
        names = [k for k, _ in d1.named_parameters()]
        grads = torch.autograd.grad(o, [s, a] + [p for _, p in d1.named_parameters()], cot, allow_unused=True)
        out.update({f"{tag}_state": s, f"{tag}_action": a, f"{tag}_cot": cot, f"{tag}_out": o, f"{tag}_dt": dt,
                    f"{tag}_gstate": grads[0], f"{tag}_gaction": grads[1],
                    f"{tag}_rot_drag": np.array(drag, dtype=np.float64)})
        for i, (k, p) in enumerate(d1.named_parameters()):
            out[f"{tag}_param_{i}"] = p
            out[f"{tag}_gparam_{i}"] = grads[2 + i] if grads[2 + i] is not None else torch.zeros_like(p)
        # train_dynamics_model (train_base.py:160-186) body: loss + gradient, l2_lambda = 0.01
        d2 = FlightmareDynamics(modified_params={"translational_drag": [0.3, 0.3, 0.3]})
        for p in d1.parameters():
            p.grad = None
        n1 = d1(s.detach(), a.detach(), dt=dt)
        n2 = d2(s.detach(), a.detach(), dt=dt)
        l2 = (torch.norm(d1.linear_state_2.weight) + torch.norm(d1.linear_state_2.bias) +
              torch.norm(d1.linear_state_1.weight) + torch.norm(d1.linear_state_1.bias))
        loss = torch.sum((n1 - n2) ** 2) + 0.01 * l2
        loss.backward()
        out[f"{tag}_dyn_target"] = n2
        out[f"{tag}_dyn_loss"] = loss
        for i, (k, p) in enumerate(d1.named_parameters()):
            out[f"{tag}_dyn_gparam_{i}"] = p.grad if p.grad is not None else torch.zeros_like(p)
        print("learnt", tag, "loss %.6f" % float(loss), "max |g kinv| %.4g" % float(out[f"{tag}_gparam_3"].abs().max()),
              "max |g mass| %.3g" % float(out[f"{tag}_gparam_1"].abs().max()),
              "max |g J| %.3g" % float(out[f"{tag}_gparam_2"].abs().max()))
    out["param_names"] = np.array(names)
    # ---- fixed wing: LearntFixedWingDynamics.forward (fixed_wing_dynamics.py:270-326) with seeded non-zero residual
    #      MLP; case "wa": the shipped constants, case "wb": every constant perturbed by a few percent and a general
    #      (non-symmetric, fully populated) inertia matrix, which is what SGD turns `I` into after one step
    from neural_control.dynamics.fixed_wing_dynamics import LearntFixedWingDynamics
    import warnings
    warnings.filterwarnings("ignore")
    for tag in ("wa", "wb"):
        dw_ = LearntFixedWingDynamics()
        with torch.no_grad():
            dw_.linear_state_1.weight.copy_(0.3 * torch.randn(64, 16))
            dw_.linear_state_1.bias.copy_(0.1 * torch.randn(64))
            dw_.linear_state_2.weight.copy_(0.1 * torch.randn(12, 64))
            dw_.linear_state_2.bias.copy_(0.05 * torch.randn(12))
            if tag == "wb":
                dw_.I.add_(0.003 * torch.randn(3, 3))
                for k, p in dw_.cfg.items():
                    p.mul_(1 + 0.05 * float(torch.randn(())))
                    if float(p.abs()) == 0.0:
                        p.add_(0.02 * float(torch.randn(())))
        n, dt = 29, 0.05
        s = torch.zeros(n, 12)
        s[:, :3] = torch.randn(n, 3)
        s[:, 3] = 11.5 + 0.5 * torch.randn(n)
        s[:, 4:6] = 0.4 * torch.randn(n, 2)
        s[:, 6:9] = 0.1 * torch.randn(n, 3)
        s[:, 9:] = 0.1 * torch.randn(n, 3)
        s.requires_grad_(True)
        a = torch.rand(n, 4).requires_grad_(True)
        cot = torch.randn(n, 12)
        o = dw_(s, a, dt)
        wnames = [k for k, _ in dw_.named_parameters()]
        grads = torch.autograd.grad(o, [s, a] + [p for _, p in dw_.named_parameters()], cot, allow_unused=True)
        out.update({f"{tag}_state": s, f"{tag}_action": a, f"{tag}_cot": cot, f"{tag}_out": o, f"{tag}_dt": dt,
                    f"{tag}_gstate": grads[0], f"{tag}_gaction": grads[1]})
        for i, (k, p) in enumerate(dw_.named_parameters()):
            out[f"{tag}_param_{i}"] = p
            out[f"{tag}_gparam_{i}"] = grads[2 + i] if grads[2 + i] is not None else torch.zeros_like(p)
        print("learnt wing", tag, "params", len(wnames), "max |gI| %.3g" % float(out[f"{tag}_gparam_0"].abs().max()))
    out["wing_param_names"] = np.array(wnames)
    np.savez_compressed(os.path.join(args.out, "learnt_dyn.npz"), **npify(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--only-prep", action="store_true", help="only (re)generate prep_data.npz")
    ap.add_argument("--only-eval", action="store_true", help="only (re)generate eval_rand.npz")
    ap.add_argument("--only-wing-eval", action="store_true", help="only (re)generate eval_wing.npz")
    ap.add_argument("--only-cartpole-eval", action="store_true", help="only (re)generate eval_cartpole.npz")
    ap.add_argument("--only-selfplay", action="store_true", help="only (re)generate eval_selfplay.npz")
    ap.add_argument("--only-eval-lstm", action="store_true", help="only (re)generate eval_rand_lstm.npz")
    ap.add_argument("--only-wing-selfplay", action="store_true", help="only (re)generate eval_wing_selfplay.npz")
    ap.add_argument("--only-ref-table", action="store_true", help="only (re)generate ref_table.npz")
    ap.add_argument("--only-poly-traj", action="store_true", help="only (re)generate poly_traj.npz")
    ap.add_argument("--only-learnt", action="store_true", help="only (re)generate learnt_dyn.npz")
    args = ap.parse_args()
    import_reference(args.ref)
    os.makedirs(args.out, exist_ok=True)
    if args.only_eval:
        make_eval_golden(args)
        return
    if args.only_learnt:
        make_learnt_golden(args)
        return
    if args.only_eval_lstm:
        make_eval_lstm_golden(args)
        return
    if args.only_wing_eval:
        make_wing_eval_golden(args)
        return
    if args.only_cartpole_eval:
        make_cartpole_eval_golden(args)
        return
    if args.only_selfplay:
        make_selfplay_golden(args)
        return
    if args.only_ref_table:
        make_ref_table_golden(args)
        return
    if args.only_wing_selfplay:
        make_wing_selfplay_golden(args)
        return
    if args.only_poly_traj:
        make_poly_traj_golden(args)
        return

    import torch
    from neural_control.drone_loss import quad_mpc_loss, fixed_wing_mpc_loss, cartpole_loss_mpc  # noqa: F401
    torch.autograd.set_detect_anomaly(False)
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from neural_control.dynamics.fixed_wing_dynamics import FixedWingDynamics
    from neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from neural_control.dataset import state_preprocessing
    from neural_control.models.hutter_model import Net as HutterNet
    from neural_control.models.rnn import LSTM_NEW
    from neural_control.models.simple_model import Net as SimpleNet
    import train_drone
    import train_fixed_wing
    import train_cartpole

    quad, wing, cart = FlightmareDynamics(), FixedWingDynamics(), CartpoleDynamics()
    f32 = torch.float32

    # ------------------------------------------------------------------ dataset layouts (prepare_data)
    # QuadDataset.prepare_data / WingDataset.prepare_data of the reference on raw numpy (states, ref_states)
    from neural_control.dataset import QuadDataset, WingDataset
    gp = np.random.RandomState(77)
    n_p, h_p = 6, 10
    raw_states = np.concatenate((gp.uniform(-2, 2, (n_p, 3)), gp.uniform(-0.5, 0.5, (n_p, 3)),
                                 gp.uniform(-2, 2, (n_p, 3)), gp.uniform(-0.5, 0.5, (n_p, 3))), axis=1)
    raw_refs = gp.uniform(-3, 3, (n_p, h_p, 9))
    qd = MagicMock()
    qd.to_torch = lambda a: torch.from_numpy(np.array(a)).float()
    qd.rot_world_to_body = QuadDataset.rot_world_to_body
    q_in_state, q_states, q_in_ref, q_ref = QuadDataset.prepare_data(qd, raw_states.copy(), raw_refs.copy())
    wmean_p = torch.tensor([0.0, 0.0, 0.0, 11.5, 0.0, 0.17, 0.007, 0.018, 0.02, 0.0, 0.017, 0.004]).float()
    wstd_p = torch.tensor([16.6, 0.84, 0.89, 0.62, 0.28, 0.29, 0.045, 0.104, 0.05, 0.064, 0.275, 0.056]).float()
    wd = MagicMock()
    wd.mean, wd.std, wd.dt, wd.horizon = wmean_p, wstd_p, 0.05, h_p
    wd.to_torch = qd.to_torch
    wd._compute_target_pos = lambda cs_, rv: WingDataset._compute_target_pos(wd, cs_, rv)
    w_raw_states = np.zeros((n_p, 12))
    w_raw_states[:, :3] = gp.uniform(-1, 1, (n_p, 3))
    w_raw_states[:, 3] = 11.5 + gp.uniform(-1, 1, n_p)
    w_raw_states[:, 4:] = gp.uniform(-0.3, 0.3, (n_p, 8))
    w_targets = np.stack((np.full(n_p, 50.0), gp.uniform(-5, 5, n_p), gp.uniform(-5, 5, n_p)), axis=1)
    w_in_state, w_states, w_in_ref, w_ref = WingDataset.prepare_data(wd, w_raw_states.copy(), w_targets.copy())
    np.savez_compressed(os.path.join(args.out, "prep_data.npz"), **npify(dict(
        quad_raw_states=raw_states, quad_raw_refs=raw_refs, quad_in_state=q_in_state, quad_states=q_states,
        quad_in_ref=q_in_ref, quad_ref=q_ref, wing_raw_states=w_raw_states, wing_targets=w_targets, wing_mean=wmean_p,
        wing_std=wstd_p, wing_dt=np.float64(0.05), wing_h=np.int64(h_p), wing_in_state=w_in_state,
        wing_states=w_states, wing_in_ref=w_in_ref, wing_ref=w_ref)))
    print("prep_data: quad", tuple(q_in_state.shape), tuple(q_in_ref.shape), "wing", tuple(w_in_state.shape),
          tuple(w_ref.shape))
    if args.only_prep:
        return

    # ------------------------------------------------------------------ single steps
    steps = {}
    kat1_state = torch.tensor([[-0.203302, -8.12219, 0.484883, -0.15613, -0.446313, 0.25728, -4.70952,
                                0.627684, -2.506545, -0.039999, -0.200001, 0.1]], dtype=f32)
    steps["kat1_state"] = kat1_state
    steps["kat1_action"] = torch.tensor([[0.45, 0.46, 0.3, 0.6]], dtype=f32)
    steps["kat1_out"] = quad(kat1_state, steps["kat1_action"], 0.05)
    steps["kat1b_action"] = torch.tensor([[0.7, 0.9, 0.2, 0.35]], dtype=f32)
    steps["kat1b_out"] = quad(kat1_state, steps["kat1b_action"], 0.1)
    # fixed_wing_dynamics.py:496-504 __main__ vectors
    kat2_state = torch.tensor([[0.6933, -0.8747, 0.9757, -0.8422, 0.5494, -1.1936, 0.0368, 0.8417,
                                -0.9412, -1.4291, 0.4538, -0.5257]], dtype=f32)
    kat2_action = torch.tensor([[-0.5518, -2.9553, 0.0311, -0.6691]], dtype=f32)
    steps["kat2_state"], steps["kat2_action"] = kat2_state, kat2_action
    steps["kat2_out"] = wing(kat2_state, kat2_action, 0.05)
    kat2b_state = torch.zeros(1, 12, dtype=f32)
    kat2b_state[0, 3], kat2b_state[0, 5], kat2b_state[0, 7], kat2b_state[0, 10] = 11.5, 0.3, 0.05, 0.02
    steps["kat2b_state"] = kat2b_state
    steps["kat2b_action"] = torch.tensor([[0.3, 0.55, 0.45, 0.6]], dtype=f32)
    steps["kat2b_out"] = wing(kat2b_state, steps["kat2b_action"], 0.05)
    steps["kat3_state"] = torch.tensor([[0.5, 1.3, 0.1, 0.4]], dtype=f32)
    steps["kat3_action"] = torch.tensor([[0.4]], dtype=f32)
    steps["kat3_out"] = cart(steps["kat3_state"], steps["kat3_action"], 0.02)

    g = torch.Generator().manual_seed(20260925)
    n = 16

    def rnd(*shape, lo=-1.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=f32) * (hi - lo) + lo

    # random quad states: moderate attitudes/velocities/rates
    qs = torch.cat((rnd(n, 3, lo=-2, hi=2), rnd(n, 3, lo=-0.8, hi=0.8), rnd(n, 3, lo=-3, hi=3),
                    rnd(n, 3, lo=-1, hi=1)), dim=1)
    qa = rnd(n, 4, lo=0.02, hi=0.98)
    ws = torch.zeros(n, 12, dtype=f32)
    ws[:, :3] = rnd(n, 3, lo=-5, hi=5)
    ws[:, 3] = 11.5 + rnd(n, lo=-1.5, hi=1.5)
    ws[:, 4] = rnd(n, lo=-0.8, hi=0.8)
    ws[:, 5] = rnd(n, lo=-1.0, hi=1.0)
    ws[:, 6:9] = rnd(n, 3, lo=-0.3, hi=0.3)
    ws[:, 9:12] = rnd(n, 3, lo=-0.3, hi=0.3)
    # make a few rows exercise the alpha / beta clamps (|atan| > 10 deg)
    ws[0, 5], ws[1, 5], ws[2, 4], ws[3, 4] = 3.5, -3.2, 3.0, -2.8
    wa = rnd(n, 4, lo=0.02, hi=0.98)
    cs = torch.stack((rnd(n, lo=-2.4, hi=2.4), rnd(n, lo=-1.5, hi=1.5), rnd(n, lo=-3.1, hi=3.1),
                      rnd(n, lo=-1.5, hi=1.5)), dim=1)
    ca = rnd(n, 1, lo=-1, hi=1)
    for name, dyn, s0, a0, dt in (("quad", quad, qs, qa, 0.1), ("wing", wing, ws, wa, 0.05),
                                  ("cartpole", cart, cs, ca, 0.05)):
        s = s0.clone().requires_grad_(True)
        a = a0.clone().requires_grad_(True)
        out = dyn(s, a, dt)
        cot = rnd(*out.shape)
        gs, ga = torch.autograd.grad((out * cot).sum(), (s, a))
        steps[f"rand_{name}_state"], steps[f"rand_{name}_action"] = s0, a0
        steps[f"rand_{name}_dt"] = np.float64(dt)
        steps[f"rand_{name}_out"], steps[f"rand_{name}_cot"] = out, cot
        steps[f"rand_{name}_gstate"], steps[f"rand_{name}_gaction"] = gs, ga
    # featurizer
    fs = qs.clone().requires_grad_(True)
    feat = state_preprocessing(fs)
    fcot = rnd(*feat.shape)
    steps["feat_state"], steps["feat_out"], steps["feat_cot"] = qs, feat, fcot
    steps["feat_gstate"] = torch.autograd.grad((feat * fcot).sum(), fs)[0]
    np.savez_compressed(os.path.join(args.out, "steps.npz"), **npify(steps))

    # ------------------------------------------------------------------ concurrent train steps
    class Recorder:
        """wraps a dynamics object, recording the states it returns (detached copies)."""

        def __init__(self, dyn):
            self.dyn, self.states = dyn, []

        def __call__(self, state, action, dt):
            out = self.dyn(state, action, dt)
            self.states.append(out.detach().clone())
            return out

    def grads_of(net):
        return {f"grad_{i}": (p.grad.detach().clone() if p.grad is not None else None)
                for i, (nm, p) in enumerate(net.named_parameters())}

    def params_of(net):
        d = {f"param_{i}": p.detach().clone() for i, (nm, p) in enumerate(net.named_parameters())}
        d["param_names"] = np.array([nm for nm, _ in net.named_parameters()])
        return d

    def quad_inputs(cur, ref):
        in_state = state_preprocessing(cur)
        in_ref = torch.cat((ref[..., :3], ref[..., 6:9], ref[..., 6:9] - cur[:, None, 6:9]), dim=2)
        return in_state, in_ref

    def run_quad_conc(net, cur, ref, h, dt, tag):
        in_state, in_ref = quad_inputs(cur, ref)
        rec = Recorder(quad)
        me = MagicMock()
        me.horizon, me.state_size, me.action_dim, me.delta_t = h, 12, 4, dt
        me.train_dynamics, me.net = rec, net
        for p in net.parameters():
            p.grad = None
        actions = torch.sigmoid(net(in_state, in_ref))                      # train_base.py:202-206
        action_seq = torch.reshape(actions, (-1, h, 4))
        loss = train_drone.TrainDrone.train_controller_model(me, cur, action_seq, in_ref, ref)
        d = dict(in_state=in_state, cur=cur, in_ref=in_ref, ref=ref, h=np.int64(h), dt=np.float64(dt),
                 loss=loss.detach(), actions=action_seq.detach(), states=torch.stack(rec.states, dim=1))
        d.update(params_of(net))
        d.update(grads_of(net))
        np.savez_compressed(os.path.join(args.out, f"conc_quad_{tag}.npz"), **npify(d))
        print(f"conc_quad_{tag}: loss {loss.item():.6f}")

    # KAT-4: shipped model
    net = torch.load(os.path.join(args.ref, "trained_models/quad/current_model/model_quad"), weights_only=False)
    cur = torch.cat((kat1_state, 0.5 * kat1_state), dim=0).clone()
    cur[:, :3] = 0
    kk = torch.arange(1, 11, dtype=f32)[None, :, None]
    bb = torch.arange(1, 3, dtype=f32)[:, None, None]
    ref = 0.01 * kk * bb * torch.arange(1, 10, dtype=f32)[None, None, :]
    run_quad_conc(net, cur, ref, 10, 0.1, "kat4")

    def synth_quad(nn_, L, dt, gen):
        """polynomial references + drone states as in SURVEY.md 8d."""
        def u(*shape, lo, hi):
            return torch.rand(*shape, generator=gen, dtype=f32) * (hi - lo) + lo
        c = torch.zeros(nn_, 3, 6, dtype=f32)
        c[:, :, 1] = u(nn_, 3, lo=-1.5, hi=1.5)
        for i in range(2, 6):
            c[:, :, i] = u(nn_, 3, lo=-0.5, hi=0.5) / math.factorial(i)
        t = (torch.arange(L, dtype=f32) + 1) * dt
        pw = torch.stack([t ** i for i in range(6)], dim=0)                      # (6, L)
        dpw = torch.stack([i * t ** max(i - 1, 0) if i > 0 else torch.zeros_like(t) for i in range(6)], dim=0)
        pos = torch.einsum("nai,il->nla", c, pw)
        vel = torch.einsum("nai,il->nla", c, dpw)
        ref_ = torch.zeros(nn_, L, 9, dtype=f32)
        ref_[:, :, 0:3], ref_[:, :, 6:9] = pos, vel
        cur_ = torch.zeros(nn_, 12, dtype=f32)
        cur_[:, 3:6] = u(nn_, 3, lo=-0.2, hi=0.2)
        cur_[:, 6:9] = c[:, :, 1] + 0.3 * torch.randn(nn_, 3, generator=gen, dtype=f32)
        return cur_, ref_

    import math
    g2 = torch.Generator().manual_seed(1234)
    torch.manual_seed(0)
    net = HutterNet(15, 10, 9, 40)
    cur, ref = synth_quad(8, 10, 0.1, g2)
    run_quad_conc(net, cur, ref, 10, 0.1, "rand")
    torch.manual_seed(1)
    net = HutterNet(15, 6, 9, 24)
    cur, ref = synth_quad(5, 6, 0.1, g2)
    cur[:, 9:12] = 0.3 * torch.randn(5, 3, generator=g2, dtype=f32)     # non-zero body rates too
    run_quad_conc(net, cur, ref, 6, 0.1, "rand_h6")

    # wing
    from neural_control.dataset import WingDataset
    wmean = torch.tensor([0.0, 0.0, 0.0, 11.525899887084961, -0.00016766408225521445, 0.16617104411125183,
                          0.007394296582788229, 0.018172707409, 0.020353179425001144, -0.0005361468647606671,
                          0.01662314310669899, 0.004487641621381044]).float()
    wstd = torch.tensor([16.626325607299805, 0.8449159860610962, 0.8879243731498718, 0.6243225932121277,
                         0.28072822093963623, 0.29176747798, 0.04499124363064766, 0.10370047390460968,
                         0.049977313727, 0.06449887901544571, 0.27508440613746643, 0.05634994804859]).float()

    def wing_inputs(cur, target, h, dt, mean, std):
        """the arithmetic of WingDataset.prepare_data / _compute_target_pos, through the reference methods"""
        ds = MagicMock()
        ds.mean, ds.std, ds.dt, ds.horizon = mean, std, dt, h
        ds._compute_target_pos = lambda cs_, rv: WingDataset._compute_target_pos(ds, cs_, rv)
        return WingDataset.prepare_data(ds, cur.clone(), target.clone())

    def run_wing_conc(net, cur, target, h, dt, mean, std, tag):
        in_state, cur2, in_ref, ref_ = wing_inputs(cur, target, h, dt, mean, std)
        rec = Recorder(wing)
        me = MagicMock()
        me.horizon, me.state_size, me.action_dim, me.delta_t_train = h, 12, 4, dt
        me.train_dynamics, me.net = rec, net
        for p in net.parameters():
            p.grad = None
        actions = torch.sigmoid(net(in_state, in_ref))
        action_seq = torch.reshape(actions, (-1, h, 4))
        loss = train_fixed_wing.TrainFixedWing.train_controller_model(me, cur2, action_seq, in_ref, ref_)
        d = dict(in_state=in_state, cur=cur2, in_ref=in_ref, ref=ref_, h=np.int64(h), dt=np.float64(dt),
                 loss=loss.detach(), actions=action_seq.detach(), states=torch.stack(rec.states, dim=1),
                 target=target, mean=mean, std=std)
        d.update(params_of(net))
        d.update(grads_of(net))
        np.savez_compressed(os.path.join(args.out, f"conc_wing_{tag}.npz"), **npify(d))
        print(f"conc_wing_{tag}: loss {loss.item():.6f}")

    import json
    wcfg = json.load(open(os.path.join(args.ref, "trained_models/wing/current_model/config.json")))
    net = torch.load(os.path.join(args.ref, "trained_models/wing/current_model/model_wing"), weights_only=False)
    sw = kat2b_state[0]
    cur = torch.stack((sw, sw * torch.tensor([1, 1, 1, 1.02, 1, .5, 1, -1, 1, 1, 2, 1], dtype=f32)), dim=0)
    target = torch.tensor([[50.0, 3.0, -2.0], [50.0, -4.0, 1.0]], dtype=f32)
    run_wing_conc(net, cur, target, 10, 0.05, torch.tensor(wcfg["mean"]).float(), torch.tensor(wcfg["std"]).float(),
                  "kat5")

    def synth_wing(nn_, gen):
        def u(*shape, lo, hi):
            return torch.rand(*shape, generator=gen, dtype=f32) * (hi - lo) + lo
        cur_ = torch.zeros(nn_, 12, dtype=f32)
        cur_[:, 3] = 11.5 + u(nn_, lo=-0.5, hi=0.5)
        cur_[:, 5] = u(nn_, lo=-0.5, hi=0.5)
        cur_[:, 7] = u(nn_, lo=-2, hi=2) * math.pi / 180
        cur_[:, 10] = u(nn_, lo=-0.005, hi=0.005)
        tgt = torch.stack((torch.full((nn_,), 50.0), u(nn_, lo=-5, hi=5), u(nn_, lo=-5, hi=5)), dim=1)
        return cur_, tgt

    torch.manual_seed(2)
    net = HutterNet(9, 1, 3, 80, conv=False)
    cur, target = synth_wing(8, g2)
    run_wing_conc(net, cur, target, 20, 0.05, wmean, wstd, "rand_h20")

    # cartpole: body of TrainCartpole.run_epoch (train_cartpole.py:118-165) through the real method
    def run_cart_conc(net, states, h, dt, tag):
        rec = Recorder(cart)
        me = MagicMock()
        me.horizon, me.state_size, me.action_dim, me.delta_t = h, 4, 1, dt
        me.train_dynamics, me.net = rec, net
        me.results_dict = {"trained": []}
        in_state = states.clone()
        cur_ = states.clone()
        # two identical batches so that ``running_loss / i`` (i = last batch index) is defined
        me.trainloader = [(in_state.clone(), cur_.clone()), (in_state.clone(), cur_.clone())]
        me.make_reference = lambda cs_: train_cartpole.TrainCartpole.make_reference(me, cs_)
        for p in net.parameters():
            p.grad = None
        grads, losses = [], []

        class Opt:
            def zero_grad(self_):
                for p in net.parameters():
                    p.grad = None

            def step(self_):
                grads.append([p.grad.detach().clone() for p in net.parameters()])
        me.optimizer_controller = Opt()
        me.loss_logging = lambda epoch_loss, train="controller": losses.append(epoch_loss)
        epoch_loss = train_cartpole.TrainCartpole.run_epoch(me, train="controller")
        # epoch_loss = (l0 + l1) / 1 with l0 == l1
        loss = torch.tensor(epoch_loss / 2.0, dtype=f32)
        act = net(states.clone()).reshape(-1, h, 1).detach()
        d = dict(in_state=states, cur=states, h=np.int64(h), dt=np.float64(dt), loss=loss, actions=act,
                 states=torch.stack(rec.states[:h], dim=1),
                 ref=train_cartpole.TrainCartpole.make_reference(me, states))
        d.update(params_of(net))
        d.update({f"grad_{i}": gr for i, gr in enumerate(grads[0])})
        np.savez_compressed(os.path.join(args.out, f"conc_cartpole_{tag}.npz"), **npify(d))
        print(f"conc_cartpole_{tag}: loss {loss.item():.6f}")

    net = torch.load(os.path.join(args.ref, "trained_models/cartpole/current_model/model_cartpole"),
                     weights_only=False)
    run_cart_conc(net, torch.tensor([[.5, 1.3, .1, .4], [-.2, .1, -.05, .3]], dtype=f32), 10, 0.05, "kat6")
    torch.manual_seed(3)
    net = SimpleNet(4, 5)
    st = (torch.rand(128, 4, generator=g2, dtype=f32) * 2 - 1) * torch.tensor([2.4, 7.5, math.pi, 7.5])
    st[:, 1] *= 0.2
    st[:, 3] *= 0.2
    run_cart_conc(net, st, 5, 0.05, "rand_b128_h5")

    # ------------------------------------------------------------------ recurrent forward (AR / LSTM)
    def run_rec(mode, net, cur, in_ref2, ref2, h, dt, tag):
        rec = Recorder(quad)
        me = MagicMock()
        me.horizon, me.state_size, me.action_dim, me.delta_t = h, 12, 4, dt
        me.train_dynamics, me.net, me.train_mode = rec, net, ("LSTM" if mode == "lstm" else "autoregressive")
        hc0 = {}
        if mode == "lstm":
            orig_reset = net.reset_hidden_state

            def reset(bs=1):
                orig_reset(bs)
                hc0["h0"], hc0["c0"] = net.hidden_state.clone(), net.cell_state.clone()
            net.reset_hidden_state = reset
        acts = []
        orig_forward = net.forward

        def fwd(a, b):
            o = orig_forward(a, b)
            acts.append(torch.sigmoid(o).detach().clone())
            return o
        net.forward = fwd
        in_ref_buf = in_ref2.clone()           # the method mutates this buffer in place
        orig_backward = torch.Tensor.backward
        torch.Tensor.backward = lambda self, *a, **k: None   # reference backward() raises here
        try:
            with torch.no_grad():
                loss = train_drone.TrainDrone.train_recurrent_model(me, None, cur.clone(), in_ref_buf, ref2)
        finally:
            torch.Tensor.backward = orig_backward
            net.forward = orig_forward
        d = dict(cur=cur, in_ref=in_ref2, ref=ref2, h=np.int64(h), dt=np.float64(dt), loss=loss.detach(),
                 actions=torch.stack(acts, dim=1), states=torch.stack(rec.states, dim=1))
        d.update(hc0)
        d.update(params_of(net))
        np.savez_compressed(os.path.join(args.out, f"rec_{tag}.npz"), **npify(d))
        print(f"rec_{tag}: loss {loss.item():.6f}")

    def rec_inputs(nn_, h, dt, gen):
        cur_, ref2 = synth_quad(nn_, 2 * h, dt, gen)
        in_ref2 = torch.cat((ref2[..., :3], ref2[..., 6:9], ref2[..., 6:9] - cur_[:, None, 6:9]), dim=2)
        return cur_, in_ref2, ref2

    torch.manual_seed(4)
    net = HutterNet(15, 10, 9, 4)
    cur, in_ref2, ref2 = rec_inputs(8, 10, 0.1, g2)
    run_rec("autoregressive", net, cur, in_ref2, ref2, 10, 0.1, "ar_rand")
    cur2 = cur.clone()
    cur2[:, :3] = 0.2 * torch.randn(8, 3, generator=g2, dtype=f32)      # non-zero initial position
    run_rec("autoregressive", net, cur2, in_ref2, ref2, 10, 0.1, "ar_rand_pos0")
    torch.manual_seed(5)
    net = LSTM_NEW(15, 10, 9, 4)
    torch.manual_seed(6)
    run_rec("lstm", net, cur, in_ref2, ref2, 10, 0.1, "lstm_rand")


if __name__ == "__main__":
    main()
