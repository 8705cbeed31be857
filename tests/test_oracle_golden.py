"""The oracle (oracle/apg_oracle.py) against the golden vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only.

Tolerances (fp32 oracle vs fp32 reference; both are chains of fp32 ops in different association order):
  single steps      : max abs error <= 2e-6 * scale
  rollout loss      : rel <= 2e-6
  rollout gradients : per-tensor L2 rel <= 2e-5
"""
import numpy as np
import pytest
import torch

from oracle import apg_oracle as O
from tests.helpers import load_golden, golden_params, golden_grads, t, rel_err, max_rel_to_scale

STEP_TOL = 2e-6
LOSS_TOL = 2e-6
GRAD_TOL = 2e-5


def test_kat_single_steps():
    g = load_golden("steps.npz")
    for name, fn, dt in (("kat1", O.quad_step, 0.05), ("kat2", O.wing_step, 0.05), ("kat2b", O.wing_step, 0.05),
                         ("kat3", O.cartpole_step, 0.02)):
        out = fn(t(g[f"{name}_state"]), t(g[f"{name}_action"]), dt)
        assert max_rel_to_scale(out, g[f"{name}_out"]) <= STEP_TOL, name
    out = O.quad_step(t(g["kat1_state"]), t(g["kat1b_action"]), 0.1)
    assert max_rel_to_scale(out, g["kat1b_out"]) <= STEP_TOL
    # the literal values quoted in SURVEY.md 8c
    np.testing.assert_allclose(g["kat1_out"][0, :3], [-0.32615805, -8.10602474, 0.42004830], rtol=0, atol=1e-6)
    np.testing.assert_allclose(g["kat3_out"][0], [0.52600002, 1.40573132, 0.10800000, 0.77437150], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name,fn", [("quad", O.quad_step), ("wing", O.wing_step), ("cartpole", O.cartpole_step)])
def test_random_steps_and_vjp(name, fn):
    g = load_golden("steps.npz")
    dt = float(g[f"rand_{name}_dt"])
    s = t(g[f"rand_{name}_state"]).requires_grad_(True)
    a = t(g[f"rand_{name}_action"]).requires_grad_(True)
    out = fn(s, a, dt)
    assert max_rel_to_scale(out, g[f"rand_{name}_out"]) <= STEP_TOL
    gs, ga = torch.autograd.grad((out * t(g[f"rand_{name}_cot"])).sum(), (s, a))
    assert max_rel_to_scale(gs, g[f"rand_{name}_gstate"]) <= 2e-5
    assert max_rel_to_scale(ga, g[f"rand_{name}_gaction"]) <= 2e-5
    # fp64 oracle agrees with the fp32 reference to fp32 precision as well
    out64 = fn(t(g[f"rand_{name}_state"], torch.float64), t(g[f"rand_{name}_action"], torch.float64), dt)
    assert max_rel_to_scale(out64, g[f"rand_{name}_out"]) <= 5e-6


def test_state_preprocessing():
    g = load_golden("steps.npz")
    s = t(g["feat_state"]).requires_grad_(True)
    feat = O.state_preprocessing(s)
    assert max_rel_to_scale(feat, g["feat_out"]) <= STEP_TOL
    gs = torch.autograd.grad((feat * t(g["feat_cot"])).sum(), s)[0]
    assert max_rel_to_scale(gs, g["feat_gstate"]) <= 1e-5


CONC = [("quad", "conc_quad_kat4.npz"), ("quad", "conc_quad_rand.npz"), ("quad", "conc_quad_rand_h6.npz"),
        ("wing", "conc_wing_kat5.npz"), ("wing", "conc_wing_rand_h20.npz"),
        ("cartpole", "conc_cartpole_kat6.npz"), ("cartpole", "conc_cartpole_rand_b128_h5.npz")]


@pytest.mark.parametrize("system,fname", CONC)
def test_concurrent_rollout(system, fname):
    g = load_golden(fname)
    params = golden_params(g)
    h, dt = int(g["h"]), float(g["dt"])
    in_ref = t(g["in_ref"]) if "in_ref" in g else None
    loss, grads, states, actions = O.concurrent_value_and_grad(
        system, params, t(g["in_state"]), t(g["cur"]), in_ref, t(g["ref"]), h, dt)
    assert abs(float(loss) - float(g["loss"])) <= LOSS_TOL * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    assert max_rel_to_scale(actions, g["actions"]) <= 2e-6
    assert max_rel_to_scale(states, g["states"]) <= 5e-6
    ref_grads = golden_grads(g)
    for i, (go, gr) in enumerate(zip(grads, ref_grads)):
        if gr is None:
            assert go is None, f"param {i} should have no gradient (unused by the reference forward)"
        else:
            assert rel_err(go, gr) <= GRAD_TOL, (i, rel_err(go, gr))


def test_kat4_literal_values():
    """the numbers quoted in SURVEY.md 8c for the shipped quad model"""
    g = load_golden("conc_quad_kat4.npz")
    assert abs(float(g["loss"]) - 1557.710938) < 2e-3
    np.testing.assert_allclose(g["actions"][0, 0], [0.85247970, 0.26715422, 0.99926859, 0.09718276], atol=2e-6)
    assert abs(np.linalg.norm(g["grad_6"]) - 2679.971191) < 0.05     # fc1.weight
    assert "grad_4" not in g and "grad_5" not in g                    # ref_in.* unused -> None


@pytest.mark.parametrize("mode,fname", [("autoregressive", "rec_ar_rand.npz"),
                                        ("autoregressive", "rec_ar_rand_pos0.npz"), ("lstm", "rec_lstm_rand.npz")])
def test_recurrent_forward_cumulative(mode, fname):
    g = load_golden(fname)
    params = golden_params(g)
    h, dt = int(g["h"]), float(g["dt"])
    hc0 = (t(g["h0"]), t(g["c0"])) if mode == "lstm" else None
    loss, states, actions = O.rollout_recurrent(mode, params, t(g["cur"]), t(g["in_ref"]), t(g["ref"]), h, dt,
                                                window="cumulative", hc0=hc0)
    assert abs(float(loss) - float(g["loss"])) <= 5e-6 * abs(float(g["loss"]))
    assert max_rel_to_scale(actions, g["actions"]) <= 1e-5
    assert max_rel_to_scale(states, g["states"]) <= 1e-5
    # the "relative" window is a different function (documented-intent variant)
    loss_rel, _, _ = O.rollout_recurrent(mode, params, t(g["cur"]), t(g["in_ref"]), t(g["ref"]), h, dt,
                                         window="relative", hc0=hc0)
    assert abs(float(loss_rel) - float(g["loss"])) > 2e-5 * abs(float(g["loss"]))


def test_recurrent_gradients_exist_and_match_fp64():
    """No reference gradient exists for AR/LSTM (its backward raises); the gradient oracle is autograd on the
    forward-pinned restatement.  Check fp32 vs fp64 self-consistency."""
    g = load_golden("rec_ar_rand.npz")
    h, dt = int(g["h"]), float(g["dt"])
    l32, g32, _, _ = O.recurrent_value_and_grad("autoregressive", golden_params(g), t(g["cur"]), t(g["in_ref"]),
                                                t(g["ref"]), h, dt)
    d = torch.float64
    l64, g64, _, _ = O.recurrent_value_and_grad("autoregressive", golden_params(g, d), t(g["cur"], d),
                                                t(g["in_ref"], d), t(g["ref"], d), h, dt)
    assert abs(float(l32) - float(l64)) <= 2e-6 * abs(float(l64))
    for a, b in zip(g32, g64):
        if b is None:
            assert a is None
        else:
            assert rel_err(a, b) <= 5e-5


@pytest.mark.parametrize("mode,fname,window", [("autoregressive", "rec_ar_rand.npz", "cumulative"),
                                               ("autoregressive", "rec_ar_rand.npz", "relative"),
                                               ("lstm", "rec_lstm_rand.npz", "cumulative")])
def test_recurrent_gradient_oracle_against_fp64_finite_differences(mode, fname, window):
    """The reference's own backward() raises in the recurrent modes (in-place write into the shared reference buffer,
    train_drone.py:138-142), so no reference gradient exists to pin the oracle's on.  What CAN be pinned: the oracle's
    forward on the reference's forward (the goldens above), and its gradient on that forward - autograd of the
    restatement against CENTRAL FINITE DIFFERENCES of the same fp64 loss, for a random sample of entries of every
    parameter tensor.  A gradient that matches finite differences of a forward that matches the reference is the
    gradient the reference would compute if its backward ran."""
    g = load_golden(fname)
    h, dt = int(g["h"]), float(g["dt"])
    d = torch.float64
    n = 6                                                  # a few drones: the loss is a sum over drones
    params = golden_params(g, d)
    cur, in_ref, ref = t(g["cur"], d)[:n], t(g["in_ref"], d)[:n], t(g["ref"], d)[:n]
    hc0 = (t(g["h0"], d)[:n], t(g["c0"], d)[:n]) if mode == "lstm" else None

    def loss_of(ps):
        return float(O.rollout_recurrent(mode, ps, cur, in_ref, ref, h, dt, window=window, hc0=hc0)[0])

    _, grads, _, _ = O.recurrent_value_and_grad(mode, params, cur, in_ref, ref, h, dt, window=window, hc0=hc0)
    gen = torch.Generator().manual_seed(7)
    checked = 0
    for i, (p, gp) in enumerate(zip(params, grads)):
        if gp is None:
            continue
        flat = p.reshape(-1)
        for j in torch.randperm(flat.numel(), generator=gen)[:4].tolist():
            eps = 1e-6 * max(1.0, abs(float(flat[j])))
            plus = [q.clone() for q in params]
            minus = [q.clone() for q in params]
            plus[i].reshape(-1)[j] += eps
            minus[i].reshape(-1)[j] -= eps
            fd = (loss_of(plus) - loss_of(minus)) / (2 * eps)
            an = float(gp.reshape(-1)[j])
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an), float(gp.abs().max())), (mode, window, i, j, fd, an)
            checked += 1
    assert checked >= 20


def test_sgd_momentum_matches_torch():
    torch.manual_seed(0)
    p = [torch.randn(5, 3), torch.randn(7)]
    mods = [torch.nn.Parameter(x.clone()) for x in p]
    opt = torch.optim.SGD(mods, lr=1e-2, momentum=0.9)
    bufs = [None, None]
    for it in range(3):
        grads = [torch.randn_like(x) for x in p]
        for m, gr in zip(mods, grads):
            m.grad = gr.clone()
        opt.step()
        p, bufs = O.sgd_momentum_step(p, grads, bufs, 1e-2)
        for a, b in zip(p, mods):
            assert torch.allclose(a, b.detach(), atol=1e-7)


# ---------------------------------------------------------------------------------------------------------------
# N2: closed-loop evaluation on table references, pinned on QuadEvaluator.follow_trajectory("rand") of the reference
# ---------------------------------------------------------------------------------------------------------------
def _eval_runs():
    g = load_golden("eval_rand.npz")
    return g, [str(x) for x in g["run_names"]]


@pytest.mark.parametrize("name", ["gentle", "fast_reset", "fast_stop", "short_table", "tight"])
def test_eval_follow_tables_matches_reference_evaluator(name):
    g, names = _eval_runs()
    assert name in names
    params = golden_params(load_golden("conc_quad_kat4.npz"))            # the shipped model_quad
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    steps, test_time, h = int(steps), int(test_time), int(h)
    table = torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None]
    ref_states = g[f"{name}_states"]
    init = torch.tensor(ref_states[0], dtype=torch.float32)[None]
    out = O.eval_follow_tables(params, table, init, steps, h, dt, tdiv, tstab, test_time)
    taken = len(g[f"{name}_div"])
    assert int(out["n_steps"][0]) == taken
    # closed loop, fp32 policy on both sides: measured 4e-6 at most over these horizons
    assert np.abs(out["states"][0, :taken + 1].numpy() - ref_states).max() <= 2e-5
    assert np.abs(out["div"][0, :taken].numpy() - g[f"{name}_div"]).max() <= 2e-5
    assert np.abs(out["actions"][0, :taken].numpy() - g[f"{name}_actions"]).max() <= 2e-5
    # the projected reference points are the table rows at the walking index
    assert float(out["states"][0, taken + 1:].abs().sum()) == 0.0


@pytest.mark.parametrize("name", ["gentle", "fast_stop", "loose"])
def test_eval_follow_tables_lstm_policy_matches_reference_evaluator(name):
    """train_mode "LSTM": LSTM_NEW policy whose hidden / cell state is carried through the run (and through resets)"""
    g = load_golden("eval_rand_lstm.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    steps, test_time, h = int(steps), int(test_time), int(h)
    table = torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None]
    ref_states = g[f"{name}_states"]
    init = torch.tensor(ref_states[0], dtype=torch.float32)[None]
    hc0 = (torch.tensor(g[f"{name}_h0"]), torch.tensor(g[f"{name}_c0"]))
    out = O.eval_follow_tables(params, table, init, steps, h, dt, tdiv, tstab, test_time, hc0=hc0)
    taken = len(g[f"{name}_div"])
    assert int(out["n_steps"][0]) == taken
    assert np.abs(out["states"][0, :taken + 1].numpy() - ref_states).max() <= 2e-5
    assert np.abs(out["div"][0, :taken].numpy() - g[f"{name}_div"]).max() <= 2e-5
    assert np.abs(out["actions"][0, :taken].numpy() - g[f"{name}_actions"]).max() <= 2e-5
    assert np.abs(out["hc"][0].numpy() - g[f"{name}_h1"]).max() <= 2e-5
    assert np.abs(out["hc"][1].numpy() - g[f"{name}_c1"]).max() <= 2e-5


def test_eval_follow_tables_is_batched_consistently():
    """N drones at once == the same drones one by one (different tables, thresholds hit at different steps)"""
    g, names = _eval_runs()
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    tabs = [torch.tensor(g[f"{n}_table"], dtype=torch.float32)[:100] for n in ("gentle", "fast_reset", "tight")]
    tables = torch.stack(tabs)
    init = torch.zeros(3, 12)
    init[:, :3] = tables[:, 0, :3]
    both = O.eval_follow_tables(params, tables, init, 40, 10, 0.1, 0.8, 1.0, 1)
    for i in range(3):
        one = O.eval_follow_tables(params, tables[i:i + 1], init[i:i + 1], 40, 10, 0.1, 0.8, 1.0, 1)
        assert int(one["n_steps"][0]) == int(both["n_steps"][i])
        assert torch.allclose(one["states"][0], both["states"][i], atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# N3: learnt residual quadrotor dynamics, pinned on the reference's LearntDynamics (forward + autograd)
# ---------------------------------------------------------------------------------------------------------------
def _learnt_case(tag):
    g = load_golden("learnt_dyn.npz")
    lparams = [torch.tensor(g[f"{tag}_param_{i}"], requires_grad=True) for i in range(8)]
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g[f"{tag}_rot_drag"]))
    return g, lparams, cfg


@pytest.mark.parametrize("tag", ["a", "b"])
def test_learnt_dynamics_forward_and_vjp_match_reference(tag):
    g, lparams, cfg = _learnt_case(tag)
    s, a = t(g[f"{tag}_state"]).requires_grad_(True), t(g[f"{tag}_action"]).requires_grad_(True)
    out = O.learnt_quad_step(lparams, s, a, float(g[f"{tag}_dt"]), cfg)
    assert max_rel_to_scale(out, g[f"{tag}_out"]) <= 2e-6
    grads = torch.autograd.grad(out, [s, a] + lparams, t(g[f"{tag}_cot"]), allow_unused=True)
    assert max_rel_to_scale(grads[0], g[f"{tag}_gstate"]) <= 1e-5
    assert max_rel_to_scale(grads[1], g[f"{tag}_gaction"]) <= 1e-5
    scale = max(float(np.abs(g[f"{tag}_gparam_{i}"]).max()) for i in range(8))
    for i in range(8):
        want = g[f"{tag}_gparam_{i}"]
        got = grads[2 + i] if grads[2 + i] is not None else torch.zeros_like(lparams[i])
        if i in (1, 2) and float(np.abs(g[f"{tag}_rot_drag"]).max()) == 0.0:
            # mass / inertia cancel analytically; the reference's autograd leaves rounding noise (<= 3e-5 here)
            assert float(got.abs().max()) <= 1e-4 and float(np.abs(want).max()) <= 1e-4
        elif i == 1:
            assert float(got.abs().max()) <= 1e-4 and float(np.abs(want).max()) <= 1e-4
        else:
            assert float((got - t(want)).abs().max()) <= 2e-5 * max(float(np.abs(want).max()), 1e-3 * scale), i


@pytest.mark.parametrize("tag", ["a", "b"])
def test_learnt_dynamics_training_loss_matches_reference(tag):
    g, lparams, cfg = _learnt_case(tag)
    loss = O.learnt_dynamics_loss(lparams, t(g[f"{tag}_state"]), t(g[f"{tag}_action"]), t(g[f"{tag}_dyn_target"]),
                                  float(g[f"{tag}_dt"]), 0.01, cfg)
    assert abs(float(loss) - float(g[f"{tag}_dyn_loss"])) <= 1e-5 * abs(float(g[f"{tag}_dyn_loss"]))
    grads = torch.autograd.grad(loss, lparams, allow_unused=True)
    for i in (0, 3, 4, 5, 6, 7):
        want = g[f"{tag}_dyn_gparam_{i}"]
        assert float((grads[i] - t(want)).abs().max()) <= 2e-5 * max(float(np.abs(want).max()), 1e-3), i


# ---------------------------------------------------------------------------------------------------------------
# N2 (fixed wing): closed-loop evaluation, pinned on FixedWingEvaluator.fly_to_point of the reference
# ---------------------------------------------------------------------------------------------------------------
WING_EVAL_RUNS = ["one_target", "two_targets", "tight_reset", "tight_stop", "unstable", "step_limit"]


def wing_eval_case(g, name):
    params = [torch.tensor(g[f"param_{i}"]) for i in range(14)]
    h, dt_data, dt_env = int(g["cfg"][0]), float(g["cfg"][1]), float(g["cfg"][2])
    steps, test_time, tdiv, tstab = [float(x) for x in g[f"{name}_cfg"]]
    init = torch.zeros(1, 12)
    init[0, 3] = 11.5                                                    # SimpleWingEnv.zero_reset
    targets = torch.tensor(g[f"{name}_targets"], dtype=torch.float32)[None]
    return params, targets, init, h, dt_data, dt_env, int(steps), int(test_time), tdiv, tstab


@pytest.mark.parametrize("name", WING_EVAL_RUNS)
def test_eval_fly_to_points_matches_reference_evaluator(name):
    g = load_golden("eval_wing.npz")
    params, targets, init, h, dt_data, dt_env, steps, test_time, tdiv, tstab = wing_eval_case(g, name)
    out = O.eval_fly_to_points(params, targets, init, g["mean"], g["std"], steps, h, dt_data, dt_env, tdiv, tstab,
                               test_time)
    traj, dl, dtg = g[f"{name}_traj"], g[f"{name}_div_linear"], g[f"{name}_div_target"]
    taken = len(dl)
    assert int(out["n_steps"][0]) == taken
    scale = np.abs(traj[:, :12]).max()
    assert np.abs(out["states"][0, 1:taken + 1].numpy() - traj[:, :12]).max() <= 2e-5 * scale
    assert np.abs(out["actions"][0, :taken].numpy() - traj[:, 12:]).max() <= 2e-5
    assert np.abs(out["div_linear"][0, :taken].numpy() - dl).max() <= 2e-5 * max(dl.max(), 1.0)
    assert int(out["div_target_cnt"][0]) == len(dtg)
    assert abs(float(out["div_target_sum"][0]) - dtg.sum()) <= 1e-4 * max(dtg.sum(), 1.0)


@pytest.mark.parametrize("tag", ["wa", "wb"])
def test_learnt_wing_dynamics_forward_and_vjp_match_reference(tag):
    """LearntFixedWingDynamics (fixed_wing_dynamics.py:270-326): shipped constants (wa) and every constant perturbed
    with a fully populated inertia matrix (wb)"""
    g = load_golden("learnt_dyn.npz")
    assert [str(x) for x in g["wing_param_names"]][1:38] == ["cfg." + k for k in O.WING_LEARNT_KEYS]
    lparams = [torch.tensor(g[f"{tag}_param_{i}"], requires_grad=True) for i in range(42)]
    s, a = t(g[f"{tag}_state"]).requires_grad_(True), t(g[f"{tag}_action"]).requires_grad_(True)
    out = O.learnt_wing_step(lparams, s, a, float(g[f"{tag}_dt"]))
    assert max_rel_to_scale(out, g[f"{tag}_out"]) <= 2e-6
    grads = torch.autograd.grad(out, [s, a] + lparams, t(g[f"{tag}_cot"]), allow_unused=True)
    assert max_rel_to_scale(grads[0], g[f"{tag}_gstate"]) <= 1e-5
    assert max_rel_to_scale(grads[1], g[f"{tag}_gaction"]) <= 1e-5
    for i in range(42):
        want = t(g[f"{tag}_gparam_{i}"])
        got = grads[2 + i] if grads[2 + i] is not None else torch.zeros_like(lparams[i])
        assert float((got - want).abs().max()) <= 1e-5 * max(float(want.abs().max()), 1e-3), i
    # with the shipped constants the general-inertia restatement reduces to the pinned wing_step
    if tag == "wa":
        assert max_rel_to_scale(O.learnt_wing_step([p.detach() for p in lparams[:38]] +
                                                   [torch.zeros(64, 16), torch.zeros(64), torch.zeros(12, 64),
                                                    torch.zeros(12)], s.detach(), a.detach(), 0.05),
                                O.wing_step(s.detach(), a.detach(), 0.05)) <= 1e-6


CARTPOLE_EVAL_RUNS = ["zero_start", "tilted", "falls", "tight", "falls_at_once"]


@pytest.mark.parametrize("name", CARTPOLE_EVAL_RUNS)
def test_eval_cartpole_balance_matches_reference_evaluator(name):
    """Evaluator.evaluate_in_environment (scripts/evaluate_cartpole.py:78-262) with CartpoleWrapper on the shipped
    model_cartpole: the states env._step returned, the success count and the |x_dot| log"""
    g = load_golden("eval_cartpole.npz")
    assert [str(x) for x in g["run_names"]] == CARTPOLE_EVAL_RUNS
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    steps, tdiv, burn = g[f"{name}_cfg"]
    out = O.eval_cartpole_balance(params, torch.tensor(g[f"{name}_init"], dtype=torch.float32)[None], int(steps), 0.05,
                                  float(tdiv), int(burn))
    want = g[f"{name}_states"]
    taken = len(want)
    assert int(out["n_steps"][0]) == taken
    assert int(out["success"][0]) == int(g[f"{name}_success"][0])
    assert np.abs(out["states"][0, :taken].numpy() - want).max() <= 1e-6 * max(np.abs(want).max(), 1.0)
    assert abs(float(out["vel_sum"][0]) - g[f"{name}_vel"].sum()) <= 1e-5 * max(g[f"{name}_vel"].sum(), 1.0)
    if taken - 1 > burn:
        assert abs(float(out["mean_angle"][0]) - np.abs(want[int(burn) + 1:, 2]).mean()) <= 1e-6
    else:
        assert float(out["mean_angle"][0]) == 100.0


def test_eval_cartpole_balance_is_batched_consistently():
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    init = torch.tensor(np.stack([g[f"{n}_init"] for n in CARTPOLE_EVAL_RUNS]), dtype=torch.float32)
    out = O.eval_cartpole_balance(params, init, 60, 0.05, 0.21, 5)
    for k, name in enumerate(CARTPOLE_EVAL_RUNS):
        one = O.eval_cartpole_balance(params, init[k:k + 1], 60, 0.05, 0.21, 5)
        assert int(one["n_steps"][0]) == int(out["n_steps"][k])
        assert float((one["states"][0] - out["states"][k]).abs().max()) <= 1e-5


def test_eval_selfplay_feed_matches_reference_dataset():
    """which policy calls NetworkWrapper.predict_actions adds to the dataset, the raw (state, window) it hands over
    (also right after a reset) and the ring slots (network_wrapper.py:42-52, dataset.py:78-119): three runs with one
    running action counter, take_every_x = 7, 3 sampled rows + 5 self-play slots"""
    g = load_golden("eval_selfplay.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    h, dt, take, n_sampled, n_slots = [float(v) for v in g["cfg"]]
    h, take = int(h), int(take)
    ac, ks, kr = 0, [], []
    for name in [str(v) for v in g["run_names"]]:
        steps, tdiv, tstab = g[f"{name}_cfg"]
        out = O.eval_follow_tables(params, torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None],
                                   torch.tensor(g[f"{name}_states"][:1], dtype=torch.float32), int(steps), h, dt,
                                   tdiv, tstab, 0, record_policy_inputs=True)
        assert int(out["n_steps"][0]) == len(g[f"{name}_div"])
        kept, ac = O.selfplay_kept_calls(out["n_steps"], take, ac)
        ks += [out["policy_states"][r, i].numpy() for r, i in kept]
        kr += [out["windows"][r, i].numpy() for r, i in kept]
    assert ac == int(g["action_counter"][0]) and len(ks) == int(g["eval_counter"][0])
    assert np.abs(np.array(ks) - g["kept_states"]).max() <= 5e-6
    assert np.abs(np.array(kr) - g["kept_refs"]).max() <= 1e-6
    slots, counter = O.selfplay_ring_slots(len(ks), int(n_sampled), int(n_slots))
    final = {}
    for slot, s in zip(slots, ks):
        final[slot] = s
    for slot, s in final.items():                                    # dataset `states`: position zeroed
        assert np.abs(s[3:] - g["ds_states"][slot, 3:]).max() <= 5e-6 and np.abs(g["ds_states"][slot, :3]).max() == 0
    assert counter == int(g["eval_counter"][0]) and sorted(final) == list(range(int(n_sampled), int(n_sampled + n_slots)))


def _wing_selfplay_runs(g):
    params = [torch.tensor(load_golden("eval_wing.npz")[f"param_{i}"]) for i in range(14)]
    h, dt_data, dt_env, take, n_sampled, n_slots = [float(v) for v in g["cfg"]]
    gw = load_golden("eval_wing.npz")
    for name in [str(v) for v in g["run_names"]]:
        steps, tdiv, tstab = g[f"{name}_cfg"]
        init = torch.zeros(1, 12)
        init[0, 3] = 11.5
        targets = torch.tensor(g[f"{name}_targets"], dtype=torch.float32)[None]
        yield (name, params, targets, init, gw["mean"], gw["std"], int(steps), int(h), dt_data, dt_env, float(tdiv),
               float(tstab))


def test_eval_wing_selfplay_feed_matches_reference_dataset():
    """FixedWingNetWrapper.predict_actions -> get_and_add_eval_data(add_to_dataset=True) (network_wrapper.py:81-90):
    three flights with one running action counter, take_every_x = 11; kept (state, target) pairs incl. the target
    switch of the two-target flight and states shown right after a reset"""
    g = load_golden("eval_wing_selfplay.npz")
    take = int(g["cfg"][3])
    ac, ks, kt = 0, [], []
    for name, params, targets, init, mean, std, steps, h, dt_data, dt_env, tdiv, tstab in _wing_selfplay_runs(g):
        out = O.eval_fly_to_points(params, targets, init, mean, std, steps, h, dt_data, dt_env, tdiv, tstab, 0,
                                   record_policy_inputs=True)
        assert int(out["n_steps"][0]) == int(g[f"{name}_n_steps"][0])
        kept, ac = O.selfplay_kept_calls(out["n_steps"], take, ac)
        ks += [out["policy_states"][r, i].numpy() for r, i in kept]
        kt += [targets[r, int(out["target_index"][r, i])].numpy() for r, i in kept]
    assert ac == int(g["action_counter"][0]) and len(ks) == int(g["eval_counter"][0])
    assert np.abs(np.array(kt) - g["kept_targets"]).max() == 0
    assert len({tuple(t) for t in g["kept_targets"][:9]}) == 2                   # both targets of flight "a" occur
    assert np.abs(np.array(ks) - g["kept_states"]).max() <= 2e-5 * np.abs(g["kept_states"]).max()
