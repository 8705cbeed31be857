"""GPU tests of everything built after the round-1 GPU budget was spent: the input side of the path (SURVEY.md 8f N1 /
N4: dataset layouts on the device, chunked host-batch train step, device-resident dataset), the closed-loop evaluation
rollouts (N2, quadrotor + fixed wing), the learnt residual dynamics (N3) and the tcgen05 path against the mma.sync one.
The file name sorts last on purpose: under `pytest -x` these run after every previously verified parity test.

Tolerances (fp32): subtraction-only outputs bit-exact; outputs that go through sin/cos/sqrt/div 2e-6 of scale;
train steps from raw host samples vs the same steps from prepared device tensors: loss 1e-5 relative, gradient L2
1e-4 relative (chunk-wise summation order + last-ulp feature differences)."""
import math
import os

import numpy as np
import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _mods():
    from apg_trajectory_tracking_b200 import prepare as PR, rollout as R, synthetic as SY, train as T
    from apg_trajectory_tracking_b200.neural_control import dataset as DS
    return PR, R, SY, T, DS


def _close(a, b, tol):
    a, b = a.detach().cpu().double(), torch.as_tensor(b).double()
    return float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30)


def test_prepare_quad_golden_and_in_place():
    PR, *_ = _mods()
    g = load_golden("prep_data.npz")
    s = torch.tensor(g["quad_raw_states"], dtype=torch.float32).cuda()
    r = torch.tensor(g["quad_raw_refs"], dtype=torch.float32).cuda()
    out = PR.prepare_quad(s, r)
    assert torch.equal(out["cur"].cpu(), torch.tensor(g["quad_states"]))
    assert torch.equal(out["ref"].cpu(), torch.tensor(g["quad_ref"]))
    assert torch.equal(out["in_ref"].cpu(), torch.tensor(g["quad_in_ref"]))
    assert _close(out["in_state"], g["quad_in_state"], 2e-6)
    s2, r2 = s.clone(), r.clone()
    out2 = PR.prepare_quad(s2, r2, in_place=True)
    assert out2["cur"].data_ptr() == s2.data_ptr() and out2["ref"].data_ptr() == r2.data_ptr()
    assert torch.equal(s2, out["cur"]) and torch.equal(r2, out["ref"]) and torch.equal(out2["in_ref"], out["in_ref"])


@pytest.mark.parametrize("n,L", [(1, 10), (63, 10), (257, 20), (4099, 7)])
def test_prepare_quad_random_vs_host_dataset(n, L):
    PR, R, SY, T, DS = _mods()
    g = torch.Generator().manual_seed(n + L)
    s = torch.randn(n, 12, generator=g)
    s[:, 3:6] *= 0.3
    r = torch.randn(n, L, 9, generator=g) * 2
    want = DS.QuadDataset.prepare_data(None, s.clone(), r.clone())
    out = PR.prepare_quad(s.cuda(), r.cuda())
    assert torch.equal(out["cur"].cpu(), want[1]) and torch.equal(out["ref"].cpu(), want[3])
    assert torch.equal(out["in_ref"].cpu(), want[2])
    assert _close(out["in_state"], want[0], 2e-6)
    only = PR.prepare_quad(s.cuda(), r.cuda(), want=("in_ref",))
    assert set(only) == {"in_ref"} and torch.equal(only["in_ref"], out["in_ref"])


def test_prepare_wing_golden_and_random():
    PR, R, SY, T, DS = _mods()
    g = load_golden("prep_data.npz")
    s = torch.tensor(g["wing_raw_states"], dtype=torch.float32).cuda()
    tg = torch.tensor(g["wing_targets"], dtype=torch.float32).cuda()
    out = PR.prepare_wing(s, tg, g["wing_mean"], g["wing_std"], float(g["wing_dt"]), int(g["wing_h"]))
    assert torch.equal(out["cur"].cpu(), torch.tensor(g["wing_states"]))
    for k, name in (("in_state", "wing_in_state"), ("in_ref", "wing_in_ref"), ("ref", "wing_ref")):
        assert _close(out[k], g[name], 2e-6), k
    c = SY.wing_case(1031, 20, 0.05, seed=3)
    out = PR.prepare_wing(c["cur"].cuda(), c["target"].cuda(), SY.WING_MEAN, SY.WING_STD, 0.05, 20)
    for k in ("in_state", "in_ref", "ref"):
        assert _close(out[k], c[k], 2e-6), k


def test_sample_windows_and_poly_reference():
    PR, R, SY, T, DS = _mods()
    rng = np.random.default_rng(0)
    Tn, W, L = 1001, 12, 10
    traj = torch.tensor(rng.standard_normal((Tn, W)), dtype=torch.float32)
    stride = 2 * L
    n = len(traj[:-(L + 1)][::stride])
    states, refs = PR.sample_windows(traj.cuda(), n, L, stride)
    idx = torch.arange(n) * stride
    assert torch.equal(states.cpu()[:, :9], traj[idx, :9]) and float(states[:, 9:].abs().max()) == 0.0
    for k in range(L):
        assert torch.equal(refs.cpu()[:, k], traj[idx + k + 1, :9])
    with pytest.raises(RuntimeError):
        PR.sample_windows(traj.cuda(), n + 1, L, stride * 2)        # would read past the table
    # polynomial rows against the bench's host generator (same coefficients)
    nq, rows, dt = 300, 20, 0.1
    gen = torch.Generator().manual_seed(9)
    c = torch.zeros(nq, 3, 6)
    c[:, :, 1] = torch.rand(nq, 3, generator=gen) * 3 - 1.5
    for i in range(2, 6):
        c[:, :, i] = (torch.rand(nq, 3, generator=gen) - 0.5) / math.factorial(i)
    t = (torch.arange(rows, dtype=torch.float64) + 1) * dt
    pw = torch.stack([t ** i for i in range(6)])
    dpw = torch.stack([torch.zeros_like(t) if i == 0 else i * t ** (i - 1) for i in range(6)])
    want = torch.zeros(nq, rows, 9, dtype=torch.float64)
    want[:, :, 0:3] = torch.einsum("nai,il->nla", c.double(), pw)
    want[:, :, 6:9] = torch.einsum("nai,il->nla", c.double(), dpw)
    got = PR.poly_reference(c.cuda(), rows, dt)
    assert _close(got, want, 2e-6)


def _params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(*s, generator=g) * 2 - 1) / (s[-1] if len(s) > 1 else 64) ** 0.5 for s in shapes]


CASES = {
    "quad": dict(h=10, dt=0.1, n=1000, chunk=192),
    "wing": dict(h=20, dt=0.05, n=777, chunk=256),
    "cartpole": dict(h=5, dt=0.05, n=300, chunk=128),
    "autoregressive": dict(h=10, dt=0.1, n=200, chunk=64),
    "lstm": dict(h=10, dt=0.1, n=200, chunk=128),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_step_host_from_raw_samples_equals_step_from_prepared_tensors(name):
    """3 SGD iterations: FusedTrainStep.step(prepared CUDA tensors) vs step_host(raw pinned host samples, several
    ragged chunks) -- same losses, same parameters afterwards"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    c = CASES[name]
    h, dt, n = c["h"], c["dt"], c["n"]
    system = "quad" if name in ("quad", "autoregressive", "lstm") else name
    mode = name if name in ("autoregressive", "lstm") else "concurrent"
    w = dict(system=system, mode=mode, h=h, dt=dt)
    params = B.default_init(system, h, seed=4, mode=mode)
    spec = B.make_spec(w)
    case = B.make_case(w, n, 21, "cpu")
    lr = 1e-4
    a = T.FusedTrainStep(params, spec, n, lr=lr, device="cuda:0", distributed=False)
    b = T.FusedTrainStep(params, spec, n, lr=lr, device="cuda:0", distributed=False)
    dev = {k: (v.cuda() if v is not None else None) for k, v in case.items()}
    pin = {k: (v.clone().pin_memory() if v is not None else None) for k, v in case.items()}
    for it in range(3):
        la = a.step(dev.get("in_state"), dev["cur"], dev.get("in_ref"), dev.get("ref"), dev.get("h0c0"))
        lb = b.step_host(pin["cur"], ref=pin.get("ref"), h0c0=pin.get("h0c0"), target=pin.get("target"),
                         chunk=c["chunk"])
        la, lb = float(la.item()), float(lb.item())
        assert abs(la - lb) <= 1e-5 * abs(la), (name, it, la, lb)
        assert rel_err(b.grad, a.grad) <= 1e-4, (name, it)
    assert rel_err(b.flat, a.flat) <= 1e-6
    # the whole batch as one chunk goes through the same code path
    l1, g1 = b.value_and_grad_host(pin["cur"], ref=pin.get("ref"), h0c0=pin.get("h0c0"), target=pin.get("target"),
                                   chunk=0)
    l1, g1 = float(l1.item()), g1.clone()
    l2, g2 = b.value_and_grad_host(pin["cur"], ref=pin.get("ref"), h0c0=pin.get("h0c0"), target=pin.get("target"),
                                   chunk=64)
    assert abs(l1 - float(l2.item())) <= 1e-5 * abs(l1) and rel_err(g2, g1) <= 1e-4


def test_step_host_async_overlapped_steps_equal_step_host():
    """loss read one step late, alternating staging sets: same losses and parameters as the synchronous loop"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    n, h, dt = 1000, 10, 0.1
    w = dict(system="quad", mode="concurrent", h=h, dt=dt)
    params = B.default_init("quad", h, seed=3)
    a = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    b = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    batches = [{k: v.pin_memory() for k, v in B.make_case(w, n, 40 + i, "cpu").items()} for i in range(5)]
    la = [float(a.step_host(c["cur"], ref=c["ref"], chunk=256).item()) for c in batches]
    lb, prev = [], None
    for c in batches:
        hnd = b.step_host_async(c["cur"], ref=c["ref"], chunk=256)
        if prev is not None:
            lb.append(prev.item())
        prev = hnd
    lb.append(prev.item())
    assert la == lb, (la, lb)                        # same kernels, same order: bitwise equal
    assert torch.equal(a.flat, b.flat)


@pytest.mark.parametrize("name", ["quad", "wing"])
def test_step_host_captured_graph_equals_eager_step_host(name):
    """capture_host / replay_host: one CUDA-graph launch per step (chunked H2D branch + kernels + SGD + loss D2H) ==
    the eager step_host on the same chunk schedule, with NEW data written into the same pinned tensors every step"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    c = CASES[name]
    h, dt, n = c["h"], c["dt"], 1000
    w = dict(system=name, mode="concurrent", h=h, dt=dt)
    params = B.default_init(name, h, seed=3)
    a = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    b = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    batches = [B.make_case(w, n, 60 + i, "cpu") for i in range(5)]
    pin = {k: v.clone().pin_memory() for k, v in batches[0].items() if k in ("cur", "ref", "target")}
    kw = dict(ref=pin.get("ref") if name == "quad" else None, target=pin.get("target"), chunk=192)
    b.capture_host(pin["cur"], warmup=2, **kw)          # 2 warm-up steps + the captured one do not count: replays do
    for _ in range(2):
        a.step_host(pin["cur"], **kw)
    la, lb = [], []
    for case in batches:
        for k in pin:
            pin[k].copy_(case[k])
        la.append(float(a.step_host(pin["cur"], **kw).item()))
        lb.append(b.replay_host().item())
    assert la == lb, (la, lb)
    assert torch.equal(a.flat, b.flat)
    with pytest.raises(Exception):
        b.capture_host(batches[0]["cur"], ref=batches[0].get("ref"), target=batches[0].get("target"))   # not pinned


def test_step_host_accepts_absolute_positions():
    """truly raw quad samples (drone not at the origin): the device prepare makes them relative like the dataset"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    n, h, dt = 500, 10, 0.1
    w = dict(system="quad", mode="concurrent", h=h, dt=dt)
    case = B.make_case(w, n, 5, "cpu")
    params = B.default_init("quad", h, seed=1)
    st = T.FusedTrainStep(params, B.make_spec(w), n, lr=0.0, device="cuda:0", distributed=False)
    l0, g0 = st.value_and_grad_host(case["cur"].pin_memory(), ref=case["ref"].pin_memory(), chunk=128)
    l0, g0 = float(l0.item()), g0.clone()
    off = torch.randn(n, 3) * 5
    cur, ref = case["cur"].clone(), case["ref"].clone()
    cur[:, :3] += off
    ref[:, :, :3] += off[:, None, :]
    l1, g1 = st.value_and_grad_host(cur.pin_memory(), ref=ref.pin_memory(), chunk=128)
    assert abs(float(l1.item()) - l0) <= 2e-5 * abs(l0)      # (ref + off) - (cur + off) rounds differently
    assert rel_err(g1, g0) <= 2e-4


# ---------------------------------------------------------------------------------------------------------------
# N2: closed-loop evaluation rollout (csrc/eval_kernels.cu) vs the reference evaluator's golden runs and the oracle
# ---------------------------------------------------------------------------------------------------------------
def _eval_mods():
    from apg_trajectory_tracking_b200 import evaluate as EV, rollout as R
    from oracle import apg_oracle as O
    from tests.helpers import golden_params
    return EV, R, O, golden_params


@pytest.mark.parametrize("name", ["gentle", "fast_reset", "fast_stop", "short_table", "tight"])
def test_eval_rollout_matches_reference_evaluator(name):
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))            # shipped model_quad, Net(15,10,9,40)
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    steps, test_time, h = int(steps), int(test_time), int(h)
    ev = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(h, dt), 1, "cuda:0")
    ref_states = g[f"{name}_states"]
    out = ev.follow(R.flatten_params(params).cuda(), torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None].cuda(),
                    init_states=torch.tensor(ref_states[:1], dtype=torch.float32).cuda(), steps=steps,
                    thresh_div=tdiv, thresh_stable=tstab, test_time=test_time)
    taken = len(g[f"{name}_div"])
    assert int(out["n_steps"][0]) == taken
    # closed loop over up to 80 steps with resets: fp32 (3xTF32 policy) vs the reference's fp32 policy / fp64 dynamics
    assert np.abs(out["states"][0, :taken + 1].cpu().numpy() - ref_states).max() <= 1e-4
    assert np.abs(out["div"][0, :taken].cpu().numpy() - g[f"{name}_div"]).max() <= 1e-4
    assert np.abs(out["actions"][0, :taken].cpu().numpy() - g[f"{name}_actions"]).max() <= 1e-4
    assert float(out["states"][0, taken + 1:].abs().sum()) == 0.0


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_sample_windows_reference_table_matches_load_prepare_trajectory(name):
    from apg_trajectory_tracking_b200 import prepare as PR
    g = load_golden("ref_table.npz")
    dt, speed = [float(v) for v in g[f"{name}_cfg"]]
    tab = PR.reference_table(torch.tensor(g[f"{name}_raw"]).cuda(), dt, speed, z_offset=0.0).cpu().numpy()
    want = g[f"{name}_table"]
    assert tab.shape == want.shape
    assert np.array_equal(tab[:, :3], want[:, :3].astype(np.float32))
    assert np.abs(tab[:, 3:6] - want[:, 3:6]).max() <= 5e-6              # device atan2f / asinf
    assert np.abs(tab[:, 6:] - want[:, 6:]).max() <= 1e-6 * np.abs(want[:, 6:]).max()


def test_sample_windows_polynomial_points_match_reference_polynomial_class():
    from apg_trajectory_tracking_b200 import prepare as PR
    g = load_golden("poly_traj.npz")
    for name in [str(v) for v in g["case_names"]]:
        x_range, degree, mdd, h, hover = g[f"{name}_cfg"]
        pts, ref_len = PR.polynomial_points(torch.tensor(g[f"{name}_coef"])[None].cuda(),
                                            torch.tensor(g[f"{name}_rot"])[None].cuda(),
                                            torch.tensor(g[f"{name}_start"])[None].cuda(), x_range=x_range,
                                            max_drone_dist=mdd, horizon=int(h), hover_steps=int(hover))
        want = g[f"{name}_points"]
        assert int(ref_len[0]) == len(want)
        assert np.abs(pts[0, :len(want)].cpu().numpy() - want).max() <= 2e-6     # double march (FMA-contracted)
    # many trajectories at once: every thread marches its own polynomial
    n = 300
    coef = torch.tensor(g["a_coef"])[None].repeat(n, 1) * (1 + 0.01 * torch.arange(n, dtype=torch.float64)[:, None])
    rot = torch.tensor(g["a_rot"])[None].repeat(n, 1, 1)
    pts, ref_len = PR.polynomial_points(coef.cuda(), rot.cuda(), None, x_range=5, hover_steps=3)
    from oracle import apg_oracle as O
    for i in (0, 17, n - 1):
        want = O.polynomial_points(coef[i].numpy(), rot[i].numpy(), None, 5.0, 0.025, 3)
        assert int(ref_len[i]) == len(want)
        assert np.abs(pts[i, :len(want)].cpu().numpy() - want).max() <= 1e-5 * max(np.abs(want).max(), 1.0)


def test_eval_rollout_selfplay_feed_matches_reference_dataset():
    """evaluation kernel -> evaluate.selfplay_samples -> DeviceQuadDataset ring, against what the reference's
    NetworkWrapper / DroneDataset hold after the same three runs (tests/golden/eval_selfplay.npz)"""
    from apg_trajectory_tracking_b200 import device_data as DD, prepare as PR
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_selfplay.npz")
    params = golden_params(load_golden("conc_quad_kat4.npz"))
    h, dt, take, n_sampled, n_slots = [float(v) for v in g["cfg"]]
    h, take, n_sampled, n_slots = int(h), int(take), int(n_sampled), int(n_slots)
    tot = n_sampled + n_slots
    ds = DD.DeviceQuadDataset(torch.zeros(tot, 12), torch.zeros(tot, h, 9), "cuda:0", num_self_play=n_slots)
    flat = R.flatten_params(params).cuda()
    counter, kept_s, kept_r = 0, [], []
    for name in [str(v) for v in g["run_names"]]:
        steps, tdiv, tstab = g[f"{name}_cfg"]
        tab = torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None].cuda()
        ev = EV.TableEvaluator(R.RolloutSpec.quad_concurrent(h, dt), 1, "cuda:0")
        out = ev.follow(flat, tab, init_states=torch.tensor(g[f"{name}_states"][:1], dtype=torch.float32).cuda(),
                        steps=int(steps), thresh_div=tdiv, thresh_stable=tstab, test_time=0)
        assert int(out["n_steps"][0]) == len(g[f"{name}_div"])
        s, r, counter = EV.selfplay_samples(out, tab, None, h, take, tdiv, tstab, 0, counter)
        ds.add_self_play(s, r)
        kept_s.append(s)
        kept_r.append(r)
    kept_s, kept_r = torch.cat(kept_s).cpu().numpy(), torch.cat(kept_r).cpu().numpy()
    assert counter == int(g["action_counter"][0]) and ds.eval_counter == int(g["eval_counter"][0])
    assert kept_s.shape == g["kept_states"].shape
    assert np.abs(kept_s - g["kept_states"]).max() <= 1e-4 and np.abs(kept_r - g["kept_refs"]).max() <= 1e-6
    prep = PR.prepare_quad(ds.states, ds.ref_states, want=("cur", "ref"))
    assert np.abs(prep["cur"].cpu().numpy() - g["ds_states"]).max() <= 1e-4
    assert np.abs(prep["ref"].cpu().numpy() - g["ds_ref_states"]).max() <= 1e-4


@pytest.mark.parametrize("mode,n", [("concurrent", 333), ("autoregressive", 130)])
def test_eval_rollout_batched_vs_oracle(mode, n):
    """many drones (partial tiles, several tiles per CTA), shared tables through the index, random policy"""
    import bench as B
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand.npz")
    h, dt, steps = 10, 0.1, 40
    params = B.default_init("quad", h, seed=3, mode=mode)
    spec = R.RolloutSpec.quad_concurrent(h, dt) if mode == "concurrent" else R.RolloutSpec.quad_recurrent(mode, h, dt)
    tabs = torch.tensor(np.stack([g["gentle_table"][:100], g["tight_table"][:100], g["fast_reset_table"][:100]]),
                        dtype=torch.float32)
    gen = torch.Generator().manual_seed(n)
    index = torch.randint(0, 3, (n,), generator=gen, dtype=torch.int32)
    init = torch.zeros(n, 12)
    init[:, :3] = tabs[index.long(), 0, :3] + 0.05 * torch.randn(n, 3, generator=gen)
    init[:, 6:9] = 0.1 * torch.randn(n, 3, generator=gen)
    for test_time in (0, 1):
        want = O.eval_follow_tables(params, tabs[index.long()], init, steps, h, dt, 0.5, 0.4, test_time)
        ev = EV.TableEvaluator(spec, n, "cuda:0")
        out = ev.follow(R.flatten_params(params).cuda(), tabs.cuda(), init_states=init.cuda(),
                        table_index=index.cuda(), steps=steps, thresh_div=0.5, thresh_stable=0.4, test_time=test_time)
        # a threshold crossing can flip for a drone that sits within rounding of it: compare the drones whose
        # decisions agree (all but at most a handful) and require the step counts to agree for >= 98 %
        same = (out["n_steps"].cpu() == want["n_steps"].to(torch.int32))
        assert float(same.float().mean()) >= 0.98
        d = (out["states"].cpu() - want["states"]).abs().amax(dim=(1, 2))
        assert float((d[same] <= 2e-3).float().mean()) >= 0.98
        k = int(want["n_steps"].min().clamp(max=5))
        assert float((out["states"].cpu()[:, :k + 1] - want["states"][:, :k + 1]).abs().max()) <= 1e-4
        stats = EV.eval_statistics(out["div"], out["n_steps"], 0.5)
        assert len(stats) == 6 and np.isfinite(stats[4])


@pytest.mark.parametrize("name", ["gentle", "fast_stop", "loose"])
def test_eval_rollout_lstm_policy_matches_reference_evaluator(name):
    """apg_eval_rollout_lstm against QuadEvaluator.follow_trajectory("rand") of the reference with an LSTM_NEW policy
    (train_mode "LSTM"): states, divergences, applied actions and the hidden / cell state the net is left with"""
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand_lstm.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    steps, test_time, tdiv, tstab, h, dt = [float(x) for x in g[f"{name}_cfg"]]
    steps, test_time, h = int(steps), int(test_time), int(h)
    ev = EV.TableEvaluator(R.RolloutSpec.quad_recurrent("lstm", h, dt), 1, "cuda:0")
    h0c0 = torch.tensor(np.stack([g[f"{name}_h0"], g[f"{name}_c0"]])).cuda()
    out = ev.follow(R.flatten_params(params).cuda(), torch.tensor(g[f"{name}_table"], dtype=torch.float32)[None].cuda(),
                    init_states=torch.tensor(g[f"{name}_states"][:1], dtype=torch.float32).cuda(), steps=steps,
                    thresh_div=tdiv, thresh_stable=tstab, test_time=test_time, h0c0=h0c0)
    taken = len(g[f"{name}_div"])
    assert int(out["n_steps"][0]) == taken
    assert np.abs(out["states"][0, :taken + 1].cpu().numpy() - g[f"{name}_states"]).max() <= 1e-4
    assert np.abs(out["div"][0, :taken].cpu().numpy() - g[f"{name}_div"]).max() <= 1e-4
    assert np.abs(out["actions"][0, :taken].cpu().numpy() - g[f"{name}_actions"]).max() <= 1e-4
    assert np.abs(out["hc"][0, 0].cpu().numpy() - g[f"{name}_h1"][0]).max() <= 1e-4
    assert np.abs(out["hc"][1, 0].cpu().numpy() - g[f"{name}_c1"][0]).max() <= 1e-4
    assert float(out["states"][0, taken + 1:].abs().sum()) == 0.0


def test_eval_rollout_lstm_policy_batched_vs_oracle():
    """many drones with their own hidden states (partial tiles, several tiles per CTA), shared tables"""
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand_lstm.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    h, dt, steps, n = 10, 0.1, 40, 333
    spec = R.RolloutSpec.quad_recurrent("lstm", h, dt)
    tabs = torch.tensor(np.stack([g["gentle_table"][:100], g["loose_table"][:100], g["fast_stop_table"][:100]]),
                        dtype=torch.float32)
    gen = torch.Generator().manual_seed(n)
    index = torch.randint(0, 3, (n,), generator=gen, dtype=torch.int32)
    init = torch.zeros(n, 12)
    init[:, :3] = tabs[index.long(), 0, :3] + 0.05 * torch.randn(n, 3, generator=gen)
    init[:, 6:9] = 0.1 * torch.randn(n, 3, generator=gen)
    hc0 = torch.randn(2, n, 8, generator=gen)
    for test_time in (0, 1):
        want = O.eval_follow_tables(params, tabs[index.long()], init, steps, h, dt, 0.5, 0.4, test_time,
                                    hc0=(hc0[0], hc0[1]))
        out = EV.TableEvaluator(spec, n, "cuda:0").follow(
            R.flatten_params(params).cuda(), tabs.cuda(), init_states=init.cuda(), table_index=index.cuda(), steps=steps,
            thresh_div=0.5, thresh_stable=0.4, test_time=test_time, h0c0=hc0.cuda())
        same = (out["n_steps"].cpu() == want["n_steps"].to(torch.int32))
        assert float(same.float().mean()) >= 0.98
        d = (out["states"].cpu() - want["states"]).abs().amax(dim=(1, 2))
        assert float((d[same] <= 2e-3).float().mean()) >= 0.98
        k = int(want["n_steps"].min().clamp(max=5))
        assert float((out["states"].cpu()[:, :k + 1] - want["states"][:, :k + 1]).abs().max()) <= 1e-4
        dh = (out["hc"][0].cpu() - want["hc"][0]).abs().amax(dim=1)
        assert float((dh[same] <= 2e-3).float().mean()) >= 0.98


# ---------------------------------------------------------------------------------------------------------------
# N3: learnt residual quadrotor dynamics (csrc/learnt_kernels.cu) vs the reference's LearntDynamics and the oracle
# ---------------------------------------------------------------------------------------------------------------
def _learnt_module(g, tag):
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_trained import LearntDynamics
    d = LearntDynamics({"rotational_drag": [float(x) for x in g[f"{tag}_rot_drag"]]})
    with torch.no_grad():
        for i, (_, p) in enumerate(d.named_parameters()):
            if i not in (1, 2, 3):                 # mass / inertia / kinv keep their construction-time values
                p.copy_(torch.tensor(g[f"{tag}_param_{i}"]))
    return d.cuda()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_learnt_dynamics_matches_reference(tag):
    g = load_golden("learnt_dyn.npz")
    d = _learnt_module(g, tag)
    s = torch.tensor(g[f"{tag}_state"]).cuda().requires_grad_(True)
    a = torch.tensor(g[f"{tag}_action"]).cuda().requires_grad_(True)
    out = d(s, a, float(g[f"{tag}_dt"]))
    assert _close(out, g[f"{tag}_out"], 5e-6)
    (out * torch.tensor(g[f"{tag}_cot"]).cuda()).sum().backward()
    assert _close(s.grad, g[f"{tag}_gstate"], 2e-5) and _close(a.grad, g[f"{tag}_gaction"], 2e-5)
    scale = max(float(np.abs(g[f"{tag}_gparam_{i}"]).max()) for i in range(8))
    for i, (name, p) in enumerate(d.named_parameters()):
        want = torch.tensor(g[f"{tag}_gparam_{i}"])
        if i == 1 or (i == 2 and float(np.abs(g[f"{tag}_rot_drag"]).max()) == 0.0):
            assert float(p.grad.abs().max()) == 0.0                     # analytic zero (reference: rounding noise)
        else:
            err = float((p.grad.cpu() - want).abs().max())
            assert err <= 5e-5 * max(float(want.abs().max()), 1e-3 * scale), (name, err)


@pytest.mark.parametrize("tag", ["wa", "wb"])
def test_learnt_wing_dynamics_matches_reference(tag):
    """LearntFixedWingDynamics: forward, state / action cotangents and the gradient of all 42 parameter tensors"""
    from apg_trajectory_tracking_b200.neural_control.dynamics.fixed_wing_dynamics import LearntFixedWingDynamics
    g = load_golden("learnt_dyn.npz")
    d = LearntFixedWingDynamics()
    with torch.no_grad():
        for i, (_, p) in enumerate(d.named_parameters()):
            p.copy_(torch.tensor(g[f"{tag}_param_{i}"]))
    d = d.cuda()
    s = torch.tensor(g[f"{tag}_state"]).cuda().requires_grad_(True)
    a = torch.tensor(g[f"{tag}_action"]).cuda().requires_grad_(True)
    out = d(s, a, float(g[f"{tag}_dt"]))
    assert _close(out, g[f"{tag}_out"], 5e-6)
    (out * torch.tensor(g[f"{tag}_cot"]).cuda()).sum().backward()
    assert _close(s.grad, g[f"{tag}_gstate"], 5e-5) and _close(a.grad, g[f"{tag}_gaction"], 5e-5)
    for i, (name, p) in enumerate(d.named_parameters()):
        want = torch.tensor(g[f"{tag}_gparam_{i}"])
        err = float((p.grad.cpu() - want).abs().max())
        assert err <= 2e-4 * max(float(want.abs().max()), 1e-2), (name, err)
    # several tiles per block, ragged: against fp64 autograd of the oracle.  A parameter gradient is a sum over the
    # drones that can cancel by orders of magnitude (cfg.CL_del_e in case wb: 0.012 out of per-drone terms summing to
    # 72 in magnitude; the fp32 oracle itself is 1e-4 off there), so the bar is 2e-4 of the gradient plus 2e-7 (a few
    # fp32 ulps) of the cancellation scale = the sum of |per-chunk gradients| over 50-drone chunks.
    from oracle import apg_oracle as O
    gen = torch.Generator().manual_seed(1)
    n = 3000
    s0 = torch.tensor(g[f"{tag}_state"])
    sn = s0[torch.randint(0, s0.shape[0], (n,), generator=gen)] + 0.02 * torch.randn(n, 12, generator=gen)
    an, cot = torch.rand(n, 4, generator=gen), torch.randn(n, 12, generator=gen)
    lparams = [p.detach().cpu().double().clone().requires_grad_(True) for _, p in d.named_parameters()]
    so, ao = sn.double().requires_grad_(True), an.double().requires_grad_(True)
    want = O.learnt_wing_step(lparams, so, ao, 0.05)
    wg = torch.autograd.grad(want, [so, ao] + lparams, cot.double(), allow_unused=True)
    scale = [torch.zeros_like(p) for p in lparams]
    for c in range(0, n, 50):
        part = O.learnt_wing_step(lparams, sn[c:c + 50].double(), an[c:c + 50].double(), 0.05)
        pg = torch.autograd.grad(part, lparams, cot[c:c + 50].double(), allow_unused=True)
        for acc, x in zip(scale, pg):
            if x is not None:
                acc += x.abs()
    d.zero_grad()
    sc, ac = sn.cuda().requires_grad_(True), an.cuda().requires_grad_(True)
    out = d(sc, ac, 0.05)
    assert _close(out, want.detach(), 5e-6)
    (out * cot.cuda()).sum().backward()
    assert _close(sc.grad, wg[0], 5e-5) and _close(ac.grad, wg[1], 5e-5)
    for i, (name, p) in enumerate(d.named_parameters()):
        if wg[2 + i] is not None:
            err = float((p.grad.cpu().double() - wg[2 + i]).norm())
            bar = 2e-4 * float(wg[2 + i].norm()) + 2e-7 * float(scale[i].norm())
            assert err <= bar, (name, err, bar)


def test_learnt_controller_epoch_through_identity_learnt_dynamics_equals_fused_epoch():
    """TrainDrone.run_epoch with train_dynamics = LearntDynamics at its initial values (identity action transform,
    zero residual) takes the un-fused path through the learnt step and must reproduce the fused epoch"""
    from apg_trajectory_tracking_b200.scripts.train_drone import TrainDrone
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_trained import LearntDynamics
    PR, R, SY, T, DS = _mods()
    n, h, dt = 512, 10, 0.1
    raw = SY.quad_case(n, h, dt, seed=4)
    cfg = dict(delta_t=dt, horizon=h, ref_dim=9, action_dim=4, state_size=12, batch_size=128, system="quad",
               learning_rate_controller=1e-5, train_mode="concurrent", device="cuda:0")
    losses = []
    for dyn in (FlightmareDynamics(), LearntDynamics().cuda()):
        torch.manual_seed(0)
        tr = TrainDrone(dyn, FlightmareDynamics(), dict(cfg))
        tr.initialize_model(state_data=DS.QuadDataset(raw["cur"].numpy(), raw["ref"].numpy()))
        tr.trainloader = torch.utils.data.DataLoader(tr.state_data, batch_size=128, shuffle=False)
        losses.append([tr.run_epoch(epoch=e) for e in range(2)])
    for a, b in zip(*losses):
        assert abs(a - b) <= 1e-4 * abs(a), losses


@pytest.mark.parametrize("n", [100, 1337])
def test_rollout_through_learnt_dynamics_fused_vs_oracle(n):
    """apg_rollout_forward_learnt (tq_dyn_kernel<true>: every step = LearntDynamics.forward, reverse sweep through its
    state / action adjoint) + the ordinary backward: loss, states and policy gradient of the oracle's rollout through
    the learnt model (reference-pinned parameters, golden variant b), prepared AND raw inputs; the per-step loop of
    CUDA ops (learnt step kernel under autograd) gives the same loss"""
    import bench as B
    from oracle import apg_oracle as O
    PR, R, SY, T, DS = _mods()
    g = load_golden("learnt_dyn.npz")
    d = _learnt_module(g, "b")
    lparams = [p.detach().cpu().clone() for _, p in d.named_parameters()]
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g["b_rot_drag"]))
    h, dt = 10, 0.1
    params = B.default_init("quad", h, seed=n)
    case = SY.quad_case(n, h, dt, seed=n)
    want = O.value_and_grad(lambda ps: O.rollout_concurrent_learnt(ps, lparams, case["in_state"], case["cur"],
                                                                   case["in_ref"], case["ref"], h, dt, cfg), params)
    runner = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt, modified_params=dict(d.cfg)), n, "cuda:0")
    assert runner.tcgen05
    flat, lflat = R.flatten_params(params).cuda(), d._flat().detach()
    cu = {k: v.cuda() for k, v in case.items()}
    loss, states, actions = runner.forward(flat, cu["in_state"], cu["cur"], cu["in_ref"], cu["ref"], want_states=True,
                                           want_actions=True, learnt_params=lflat)
    loss = loss.clone()
    grad = runner.backward(1.0).cpu()
    assert abs(float(loss) - float(want[0])) <= 2e-5 * abs(float(want[0]))
    assert _close(states, want[2], 2e-5) and _close(actions, want[3], 2e-5)
    for got, w in zip(R.split_flat(grad, params), want[1]):
        if w is None:
            assert float(got.abs().max()) == 0
        else:
            assert rel_err(got, w) <= 1e-4
    # the un-fused form of the same rollout: h launches of the learnt step under autograd on the kernel's actions
    s = cu["cur"]
    sts = []
    for k in range(h):
        s = d(s, actions[:, k].contiguous(), dt)
        sts.append(s)
    from apg_trajectory_tracking_b200.neural_control.drone_loss import quad_mpc_loss
    l2 = quad_mpc_loss(torch.stack(sts, 1), cu["ref"], actions)
    assert abs(float(l2) - float(loss)) <= 2e-5 * abs(float(loss))
    # raw samples (prepare_data in the kernels' prologue) through the learnt steps
    raw = SY.quad_case(n, h, dt, seed=n + 1)
    gen = torch.Generator().manual_seed(n)
    raw_cur = raw["cur"].clone()
    raw_cur[:, :3] = torch.randn(n, 3, generator=gen)
    raw_ref = raw["ref"].clone()
    raw_ref[:, :, :3] += raw_cur[:, None, :3]
    prep = PR.prepare_quad(raw_cur.cuda(), raw_ref.cuda())
    l_prep, _, _ = runner.forward(flat, prep["in_state"], prep["cur"], prep["in_ref"], prep["ref"], learnt_params=lflat)
    l_prep = l_prep.clone()
    g_prep = runner.backward(1.0).clone()
    l_raw, _, _ = runner.forward(flat, None, raw_cur.cuda(), None, raw_ref.cuda(), learnt_params=lflat)
    g_raw = runner.backward(1.0)
    assert abs(float(l_raw) - float(l_prep)) <= 1e-5 * abs(float(l_prep))
    assert rel_err(g_raw, g_prep) <= 1e-4


def test_rollout_through_learnt_dynamics_full_size_properties():
    """BASELINE size (N = 65536, h = 10) through the learnt steps: run-to-run bitwise determinism, additivity of loss and
    gradient over a partition of the drones, and a 256-drone sample of the SAME batch against the oracle (loss, states,
    actions, policy gradient)"""
    import bench as B
    from oracle import apg_oracle as O
    PR, R, SY, T, DS = _mods()
    g = load_golden("learnt_dyn.npz")
    d = _learnt_module(g, "b")
    lparams = [p.detach().cpu().clone() for _, p in d.named_parameters()]
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g["b_rot_drag"]))
    n, h, dt = 65536, 10, 0.1
    params = B.default_init("quad", h, seed=6)
    case = {k: v.cuda() for k, v in SY.quad_case(n, h, dt, seed=6).items()}
    flat, lflat = R.flatten_params(params).cuda(), d._flat().detach()
    spec = R.RolloutSpec.quad_concurrent(h, dt, modified_params=dict(d.cfg))

    def run(sl):
        r = R.Rollout(spec, case["cur"][sl].shape[0], "cuda:0")
        loss, st, ac = r.forward(flat, case["in_state"][sl].contiguous(), case["cur"][sl].contiguous(),
                                 case["in_ref"][sl].contiguous(), case["ref"][sl].contiguous(), want_states=True,
                                 want_actions=True, learnt_params=lflat)
        loss = float(loss.item())
        return loss, st.cpu(), ac.cpu(), r.backward(1.0).cpu()
    full = run(slice(0, n))
    again = run(slice(0, n))
    assert full[0] == again[0] and torch.equal(full[3], again[3])
    cut = 40000
    a, b = run(slice(0, cut)), run(slice(cut, n))
    assert abs(a[0] + b[0] - full[0]) <= 2e-5 * abs(full[0])
    assert rel_err(a[3] + b[3], full[3]) <= 2e-5
    idx = torch.arange(0, n, n // 256)
    sub = {k: v[idx.cuda()].contiguous() for k, v in case.items()}
    r = R.Rollout(spec, len(idx), "cuda:0")
    ls, st, ac = r.forward(flat, sub["in_state"], sub["cur"], sub["in_ref"], sub["ref"], want_states=True,
                           want_actions=True, learnt_params=lflat)
    ls = float(ls.item())
    gs = r.backward(1.0).cpu()
    want = O.value_and_grad(lambda ps: O.rollout_concurrent_learnt(ps, lparams, sub["in_state"].cpu(), sub["cur"].cpu(),
                                                                   sub["in_ref"].cpu(), sub["ref"].cpu(), h, dt, cfg),
                            params)
    assert abs(ls - float(want[0])) <= 2e-5 * abs(float(want[0]))
    assert _close(st, want[2], 2e-5) and _close(ac, want[3], 2e-5)
    assert torch.equal(st.cpu(), full[1][idx]) and torch.equal(ac.cpu(), full[2][idx])     # same drones, same bits
    for got, w in zip(R.split_flat(gs, params), want[1]):
        if w is not None:
            assert rel_err(got, w) <= 1e-4


def test_learnt_controller_epoch_fused_vs_oracle_training():
    """TrainDrone.run_epoch with train_dynamics = LearntDynamics (golden variant b): the FUSED rollout through the learnt
    steps + SGD, two epochs of four mini-batches, against the same training on the oracle; the un-fused per-step loop
    (config unfused_learnt_rollout) follows the same curve"""
    import bench as B
    from oracle import apg_oracle as O
    from apg_trajectory_tracking_b200.scripts.train_drone import TrainDrone
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    PR, R, SY, T, DS = _mods()
    g = load_golden("learnt_dyn.npz")
    n, h, dt, bs, lr = 512, 10, 0.1, 128, 1e-5
    raw = SY.quad_case(n, h, dt, seed=4)
    cfg = dict(delta_t=dt, horizon=h, ref_dim=9, action_dim=4, state_size=12, batch_size=bs, system="quad",
               learning_rate_controller=lr, train_mode="concurrent", device="cuda:0")
    curves, nets = [], []
    for unfused in (False, True):
        torch.manual_seed(0)
        d = _learnt_module(g, "b")
        tr = TrainDrone(d, FlightmareDynamics(), dict(cfg, unfused_learnt_rollout=unfused))
        tr.initialize_model(state_data=DS.QuadDataset(raw["cur"].numpy(), raw["ref"].numpy()))
        tr.trainloader = torch.utils.data.DataLoader(tr.state_data, batch_size=bs, shuffle=False)
        if not unfused:
            p0 = [p.detach().cpu().clone() for p in tr.net.parameters()]
            assert tr.fused.supports_learnt_dynamics(bs)
        curves.append([tr.run_epoch(epoch=e) for e in range(2)])
        nets.append([p.detach().cpu().clone() for p in tr.net.parameters()])
    # the oracle's training: same batches in the same order, SGD momentum 0.9 (train_base.py:139-143)
    lparams = [p.detach().cpu().clone() for _, p in d.named_parameters()]
    ocfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g["b_rot_drag"]))
    ds = DS.QuadDataset(raw["cur"].numpy(), raw["ref"].numpy())
    ps, bufs, want_curve = [p.clone() for p in p0], [None] * len(p0), []
    for e in range(2):
        run = 0.0
        for b in range(n // bs):
            ins, cur, inr, ref = (torch.stack(x) for x in zip(*[ds[i] for i in range(b * bs, (b + 1) * bs)]))
            loss, grads, _, _ = O.value_and_grad(
                lambda q: O.rollout_concurrent_learnt(q, lparams, ins, cur, inr, ref, h, dt, ocfg), ps)
            ps, bufs = O.sgd_momentum_step(ps, grads, bufs, lr)
            run += float(loss)
        want_curve.append(run / (n // bs - 1))                    # the reference divides by the last batch index
    for got in curves:
        for a, w in zip(got, want_curve):
            assert abs(a - w) <= 1e-4 * abs(w), (curves, want_curve)
    for a, w in zip(nets[0], ps):
        assert rel_err(a, w) <= 1e-4


def test_learnt_dynamics_many_tiles_vs_oracle_and_training_step():
    """N = 5000 (several 128-drone tiles per block, ragged tail) against fp32 autograd of the oracle; then three
    train_dynamics_model steps of the trainer against the same steps on the oracle"""
    from oracle import apg_oracle as O
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.scripts.train_base import TrainBase
    g = load_golden("learnt_dyn.npz")
    d = _learnt_module(g, "b")
    gen = torch.Generator().manual_seed(0)
    n, dt = 5000, 0.05
    s, a, cot = 0.4 * torch.randn(n, 12, generator=gen), torch.rand(n, 4, generator=gen), torch.randn(n, 12, generator=gen)
    lparams = [p.detach().cpu().clone().requires_grad_(True) for _, p in d.named_parameters()]
    cfg = dict(O.QUAD_CFG, rotational_drag=tuple(float(x) for x in g["b_rot_drag"]))
    so, ao = s.clone().requires_grad_(True), a.clone().requires_grad_(True)
    want = O.learnt_quad_step(lparams, so, ao, dt, cfg)
    wg = torch.autograd.grad(want, [so, ao] + lparams, cot, allow_unused=True)
    sc, ac = s.cuda().requires_grad_(True), a.cuda().requires_grad_(True)
    out = d(sc, ac, dt)
    assert _close(out, want.detach(), 5e-6)
    (out * cot.cuda()).sum().backward()
    assert _close(sc.grad, wg[0], 2e-5) and _close(ac.grad, wg[1], 2e-5)
    for i, (name, p) in enumerate(d.named_parameters()):
        if i in (1, 2):
            continue
        assert rel_err(p.grad, wg[2 + i]) <= 1e-4, name
    # trainer step
    tr = TrainBase(d, FlightmareDynamics({"translational_drag": [0.3, 0.3, 0.3]}), delta_t=dt,
                   learning_rate_dynamics=1e-3)
    tr.init_dynamics_optimizer(l2_lambda=0.01)
    d.zero_grad()
    tgt = O.quad_step(s, a, dt, dict(O.QUAD_CFG, translational_drag=(0.3, 0.3, 0.3)))
    bufs = [None] * 8
    for it in range(3):
        loss = tr.train_dynamics_model(s.cuda(), a.cuda()[:, None, :])
        ol = O.learnt_dynamics_loss(lparams, s, a, tgt, dt, 0.01, cfg)
        og = torch.autograd.grad(ol, lparams, allow_unused=True)
        assert abs(float(loss) - float(ol)) <= 2e-5 * abs(float(ol)), (it, float(loss), float(ol))
        with torch.no_grad():
            for i, p in enumerate(lparams):
                gi = og[i] if (og[i] is not None and i not in (1, 2)) else torch.zeros_like(p)
                bufs[i] = gi.clone() if bufs[i] is None else bufs[i] * 0.9 + gi
                if i not in (1, 2, 3):             # the simulator ignores updates of these (construction-time values)
                    p -= 1e-3 * bufs[i]


@pytest.mark.parametrize("workload", ["cartpole_concurrent", "quad_concurrent"])
def test_device_captured_step_graph_equals_eager_steps(workload):
    """FusedTrainStep.capture / replay: the whole train iteration as one CUDA graph launch == the eager steps"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    w = dict(B.WORKLOADS[workload], n=300)
    n, h = w["n"], w["h"]
    case = B.make_case(w, n, 3, "cuda:0")
    params = B.default_init(w["system"], h, seed=1)
    a = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    b = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cuda:0", distributed=False)
    args = (case["in_state"], case["cur"], case.get("in_ref"), case.get("ref"))
    b.capture(*args, warmup=2)                       # 2 warm-up steps + the captured one do not run: replay does
    for _ in range(2):
        a.step(*args)                                # bring `a` to the state `b` is in after its warm-up
    la = [float(a.step(*args).item()) for _ in range(4)]
    lb = [float(b.replay().item()) for _ in range(4)]
    assert la == lb, (la, lb)
    assert torch.equal(a.flat, b.flat)


def test_device_dataset_epoch_equals_host_dataset_epoch():
    """an epoch over the device-resident raw samples == the same epoch through the host QuadDataset + DataLoader path
    (same batches, no shuffling): losses and parameters agree"""
    from apg_trajectory_tracking_b200 import device_data as DD, train as T
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    PR, R, SY, _, DS = _mods()
    n, h, dt, bs = 1000, 10, 0.1, 256
    raw = SY.quad_case(n, h, dt, seed=2)
    off = torch.randn(n, 3)
    states, refs = raw["cur"].clone(), raw["ref"].clone()
    states[:, :3] += off
    refs[:, :, :3] += off[:, None]
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        nets.append(Net(15, h, 9, 4 * h, conv=1))
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    ma, mb = T.ModuleRollout(nets[0], spec, "cuda:0"), T.ModuleRollout(nets[1], spec, "cuda:0")
    oa = torch.optim.SGD(nets[0].parameters(), lr=1e-5, momentum=0.9)
    ob = torch.optim.SGD(nets[1].parameters(), lr=1e-5, momentum=0.9)
    host = DS.QuadDataset(states.numpy(), refs.numpy())
    dev = DD.DeviceQuadDataset(states, refs, "cuda:0")
    tot, i = 0.0, 0
    for i, lo in enumerate(range(0, n, bs)):
        b = [t[lo:lo + bs] for t in (host.normed_states, host.states, host.in_ref_states, host.ref_states)]
        oa.zero_grad()
        tot += float(ma.loss_and_grad(*b))
        oa.step()
    la = tot / max(i, 1)
    lb = DD.run_epoch_device(mb, ob, dev, bs, shuffle=False)
    assert abs(la - lb) <= 2e-5 * abs(la), (la, lb)
    assert rel_err(mb.flat, ma.flat) <= 1e-6
    # the device constructors produce the layouts of the host generator / window cutter
    d2 = DD.DeviceQuadDataset.from_polynomials(512, h, dt, seed=1, device="cuda:0")
    assert d2.ref_states.shape == (512, h, 9) and float(d2.ref_states[:, :, 3:6].abs().max()) == 0.0
    traj = torch.randn(1001, 9)
    d3 = DD.DeviceQuadDataset.from_trajectory(traj, h, device="cuda:0")
    assert len(d3) == len(range(0, 1001 - (h + 1), 2 * h)) and torch.equal(d3.ref_states[3, 0].cpu(), traj[61])
    assert sum(s.shape[0] for s, _ in d3.batches(16)) == len(d3)


# ---------------------------------------------------------------------------------------------------------------
# N2 (fixed wing): eval_wing_kernel vs the reference's FixedWingEvaluator.fly_to_point and the oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["one_target", "two_targets", "tight_reset", "tight_stop", "unstable", "step_limit"])
def test_wing_fly_to_points_matches_reference_evaluator(name):
    from tests.test_oracle_golden import wing_eval_case
    EV, R, O, _ = _eval_mods()
    g = load_golden("eval_wing.npz")
    params, targets, init, h, dt_data, dt_env, steps, test_time, tdiv, tstab = wing_eval_case(g, name)
    ev = EV.WingTargetEvaluator(R.RolloutSpec.wing_concurrent(h, dt_env), 1, g["mean"], g["std"], dt_data, "cuda:0")
    out = ev.fly(R.flatten_params(params).cuda(), targets.cuda(), steps=steps, thresh_div=tdiv, thresh_stable=tstab,
                 test_time=test_time)
    traj, dl, dtg = g[f"{name}_traj"], g[f"{name}_div_linear"], g[f"{name}_div_target"]
    taken = len(dl)
    assert int(out["n_steps"][0]) == taken
    # closed loop with resets over up to ~100 steps, 3xTF32 policy vs the reference's fp32 CPU policy
    scale = np.abs(traj[:, :12]).max()
    assert np.abs(out["states"][0, 1:taken + 1].cpu().numpy() - traj[:, :12]).max() <= 5e-4 * scale
    assert np.abs(out["actions"][0, :taken].cpu().numpy() - traj[:, 12:]).max() <= 5e-4
    assert np.abs(out["div_linear"][0, :taken].cpu().numpy() - dl).max() <= 5e-4 * max(dl.max(), 1.0)
    assert int(out["div_target_cnt"][0]) == len(dtg)
    assert abs(float(out["div_target_sum"][0]) - dtg.sum()) <= 2e-3 * max(dtg.sum(), 1.0)


def test_wing_fly_to_points_batched_vs_oracle():
    EV, R, O, _ = _eval_mods()
    g = load_golden("eval_wing.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(14)]
    h, dt_data, dt_env = int(g["cfg"][0]), float(g["cfg"][1]), float(g["cfg"][2])
    n, K, steps = 200, 2, 120
    gen = torch.Generator().manual_seed(3)
    targets = torch.zeros(n, K, 3)
    targets[:, 0] = torch.tensor([25.0, 0, 0]) + (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([6.0, 8, 8])
    targets[:, 1] = torch.tensor([50.0, 0, 0]) + (torch.rand(n, 3, generator=gen) - 0.5) * torch.tensor([6.0, 10, 10])
    init = torch.zeros(n, 12)
    init[:, 3] = 11.5 + 0.3 * torch.randn(n, generator=gen)
    init[:, 7] = 0.02 * torch.randn(n, generator=gen)
    want = O.eval_fly_to_points(params, targets, init, g["mean"], g["std"], steps, h, dt_data, dt_env, 2.0, 0.4, 0)
    ev = EV.WingTargetEvaluator(R.RolloutSpec.wing_concurrent(h, dt_env), n, g["mean"], g["std"], dt_data, "cuda:0")
    out = ev.fly(R.flatten_params(params).cuda(), targets.cuda(), init_states=init.cuda(), steps=steps, thresh_div=2.0,
                 thresh_stable=0.4, test_time=0)
    same = out["n_steps"].cpu() == want["n_steps"].to(torch.int32)
    assert float(same.float().mean()) >= 0.97          # a threshold within rounding can flip a decision
    ok = same & (out["div_target_cnt"].cpu() == want["div_target_cnt"])
    d = (out["states"].cpu() - want["states"]).abs().amax(dim=(1, 2))
    assert float((d[ok] <= 5e-3).float().mean()) >= 0.97
    assert float((out["states"].cpu()[:, :6] - want["states"][:, :6]).abs().max()) <= 2e-4
    m, sd = EV.wing_eval_statistics(out["div_target_sum"], out["div_target_cnt"])
    assert np.isfinite(m) and np.isfinite(sd)


# ---------------------------------------------------------------------------------------------------------------
# cartpole balance evaluation (apg_eval_cartpole): Evaluator.evaluate_in_environment for N carts in one launch
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["zero_start", "tilted", "falls", "tight", "falls_at_once"])
def test_cartpole_balance_matches_reference_evaluator(name):
    EV, R, O, _ = _eval_mods()
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]                  # the shipped model_cartpole
    steps, tdiv, burn = g[f"{name}_cfg"]
    ev = EV.CartpoleBalanceEvaluator(R.RolloutSpec.cartpole_concurrent(10, 0.05), 1, "cuda:0")
    out = ev.balance(R.flatten_params(params).cuda(), torch.tensor(g[f"{name}_init"], dtype=torch.float32)[None].cuda(),
                     steps=int(steps), thresh_div=float(tdiv), burn_in_steps=int(burn))
    want = g[f"{name}_states"]
    taken = len(want)
    assert int(out["n_steps"][0]) == taken and int(out["success"][0]) == int(g[f"{name}_success"][0])
    # closed loop over up to 80 steps: fp32 (3xTF32 policy) vs the reference's fp32 policy / fp64 dynamics
    assert np.abs(out["states"][0, :taken].cpu().numpy() - want).max() <= 2e-4 * max(np.abs(want).max(), 1.0)
    assert float(out["states"][0, taken:].abs().sum()) == 0.0
    assert abs(float(out["vel_sum"][0]) - g[f"{name}_vel"].sum()) <= 1e-3 * max(g[f"{name}_vel"].sum(), 1.0)
    late = np.abs(want[int(burn) + 1:, 2])
    assert abs(float(out["mean_angle"][0]) - (late.mean() if len(late) else 100.0)) <= 1e-4


def test_cartpole_balance_batched_vs_oracle():
    """partial tiles and several tiles per CTA, runs of different length in one batch"""
    EV, R, O, _ = _eval_mods()
    g = load_golden("eval_cartpole.npz")
    params = [torch.tensor(g[f"param_{i}"]) for i in range(10)]
    n, steps = 20000, 60
    gen = torch.Generator().manual_seed(5)
    init = (torch.rand(n, 4, generator=gen) * 2 - 1) * torch.tensor([0.5, 1.5, 0.12, 1.8])
    want = O.eval_cartpole_balance(params, init, steps, 0.05, 0.21, 5)
    ev = EV.CartpoleBalanceEvaluator(R.RolloutSpec.cartpole_concurrent(10, 0.05), n, "cuda:0")
    out = ev.balance(R.flatten_params(params).cuda(), init.cuda(), steps=steps, thresh_div=0.21, burn_in_steps=5)
    same = out["n_steps"].cpu() == want["n_steps"].to(torch.int32)
    assert float(same.float().mean()) >= 0.995              # a threshold within rounding can flip a decision
    assert 1 < len(set(want["n_steps"].tolist()))
    d = (out["states"].cpu() - want["states"]).abs().amax(dim=(1, 2))
    assert float((d[same] <= 1e-3).float().mean()) >= 0.995
    assert float((out["states"].cpu()[:, :3] - want["states"][:, :3]).abs().max()) <= 1e-4
    assert float((out["mean_angle"].cpu() - want["mean_angle"])[same].abs().max()) <= 1e-3
    st = EV.cartpole_eval_statistics(out["n_steps"], out["vel_sum"])
    assert abs(st["mean_stable"] - float(want["success"].double().mean())) <= 0.05


# ---------------------------------------------------------------------------------------------------------------
# The tcgen05 / TMEM path (tq_kernels.cu, tq_dw_kernels.cu) is the DEFAULT path of the quadrotor concurrent rollout;
# the mma.sync kernels (APG_LEGACY_MMA=1) are an independent second implementation of the same math: both must give
# the same loss / actions / states / gradient, for ragged sizes down to one drone and up to several tiles per SM.
# (Both are also checked against the oracle and the reference goldens in tests/test_gpu_parity.py.)
# ---------------------------------------------------------------------------------------------------------------
def _check_tcgen05_vs_legacy(n):
    import bench as B
    PR, R, SY, T, DS = _mods()
    h, dt = 10, 0.1
    params = B.default_init("quad", h, seed=n % 97)
    case = {k: v.cuda() for k, v in SY.quad_case(n, h, dt, seed=n % 89).items()}
    flat = R.flatten_params(params).cuda()
    spec = R.RolloutSpec.quad_concurrent(h, dt)

    def run():
        r = R.Rollout(spec, n, "cuda:0")
        loss, states, actions = r.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"],
                                          want_states=True, want_actions=True)
        grad = r.backward(1.0)
        grad2 = r.backward(0.5)                      # the adjoint can be repeated and scales with grad_loss
        torch.cuda.synchronize()
        return float(loss.item()), states.cpu(), actions.cpu(), grad.cpu(), grad2.cpu()
    os.environ["APG_LEGACY_MMA"] = "1"
    try:
        l0, s0, a0, g0, _ = run()
    finally:
        os.environ.pop("APG_LEGACY_MMA", None)
    l1, s1, a1, g1, g1h = run()
    # one message with every diagnostic: a failure here is the only feedback a GPU run gives
    like = [torch.empty_like(p) for p in params]
    per_tensor = [round(rel_err(a, b), 7) if float(b.norm()) > 0 else float(a.abs().max())
                  for a, b in zip(R.split_flat(g1, like), R.split_flat(g0, like))]
    diag = dict(n=n, loss_legacy=l0, loss_tcgen05=l1, grad_nan=int(torch.isnan(g1).sum()),
                actions_max_abs=float((a1 - a0).abs().max()), states_max_abs=float((s1 - s0).abs().max()),
                grad_rel=rel_err(torch.nan_to_num(g1), g0), grad_rel_per_tensor=per_tensor,
                half_scale_rel=rel_err(2.0 * torch.nan_to_num(g1h), torch.nan_to_num(g1)))
    assert np.isfinite(l1) and bool(torch.isfinite(g1).all()), f"tcgen05 kernel reported a protocol timeout (NaN): {diag}"
    ok = (abs(l1 - l0) <= 1e-5 * abs(l0) and float((a1 - a0).abs().max()) <= 1e-5 and
          float((s1 - s0).abs().max()) <= 1e-5 * max(float(s0.abs().max()), 1.0) and rel_err(g1, g0) <= 1e-4 and
          rel_err(2.0 * g1h, g1) <= 1e-6)
    for a, b in zip(R.split_flat(g1, like), R.split_flat(g0, like)):
        if float(b.norm()) > 0:
            ok = ok and rel_err(a, b) <= 2e-4
        else:
            ok = ok and float(a.abs().max()) == 0.0          # ref_in.*: unused by the conv net -> exactly zero
    assert ok, str(diag)


# ---------------------------------------------------------------------------------------------------------------
# single-drone mirrors of the evaluation loop (neural_control.controllers / .environments): batch-1 callers on CUDA
# ---------------------------------------------------------------------------------------------------------------
def test_single_drone_quad_mirrors_follow_reference_run():
    from apg_trajectory_tracking_b200.neural_control import dataset as DS
    from apg_trajectory_tracking_b200.neural_control.controllers.network_wrapper import NetworkWrapper
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.drone_env import QuadRotorEnvBase
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand.npz")
    name = "gentle"
    steps, test_time, tdiv, tstab, h, dt = [float(v) for v in g[f"{name}_cfg"]]
    h = int(h)
    net = Net(15, h, 9, 4 * h)
    with torch.no_grad():
        for p, q in zip(net.parameters(), golden_params(load_golden("conc_quad_kat4.npz"))):
            p.copy_(q)
    net.cuda()
    table = g[f"{name}_table"]
    ds = DS.QuadDataset(np.zeros((4, 12)), np.zeros((4, h, 9)), self_play=1.0)
    ctrl = NetworkWrapper(net, ds, horizon=h, dt=dt, take_every_x=4)
    env = QuadRotorEnvBase(FlightmareDynamics(), dt)
    state = env.zero_reset(*table[0, :3])
    ci, want = torch.zeros(1, dtype=torch.long), g[f"{name}_states"]
    for i in range(25):
        rows, ci = O.eval_window(torch.tensor(table)[None], ci, h)
        action = ctrl.predict_actions(state, rows[0].numpy().copy())
        state, stable = env.step(action[0], thresh=tstab)
        assert stable and np.abs(state - want[i + 1]).max() <= 1e-4, i
    assert ctrl.action_counter == 25 and ds.eval_counter == 6


def test_single_drone_cartpole_mirrors_follow_reference_run():
    from apg_trajectory_tracking_b200.neural_control.controllers.network_wrapper import CartpoleWrapper
    from apg_trajectory_tracking_b200.neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.cartpole_env import CartPoleEnv
    from apg_trajectory_tracking_b200.neural_control.models.simple_model import Net
    g = load_golden("eval_cartpole.npz")
    net = Net(4, 10)
    with torch.no_grad():
        for i, p in enumerate(net.parameters()):
            p.copy_(torch.tensor(g[f"param_{i}"]))
    net.cuda()
    env = CartPoleEnv(CartpoleDynamics(), 0.05, thresh_div=0.21)
    ctrl = CartpoleWrapper(net, horizon=10, action_dim=1)
    env.state = np.array(g["falls_init"], dtype=np.float64)
    new_state, log = env.state, []
    for i in range(60):
        with torch.no_grad():
            action_seq = ctrl.predict_actions(new_state, None)
        new_state = env._step(action_seq[:, 0], is_torch=True)
        log.append(new_state.copy())
        if not env.is_upright():
            break
    want = g["falls_states"]
    assert len(log) == len(want) and np.abs(np.array(log) - want).max() <= 1e-4


def test_single_drone_evaluator_classes_run_eval_in_one_launch():
    """scripts.evaluate_drone.QuadEvaluator.run_eval on CUDA: two golden runs of the reference as one batch"""
    from apg_trajectory_tracking_b200.neural_control import dataset as DS
    from apg_trajectory_tracking_b200.neural_control.controllers.network_wrapper import NetworkWrapper
    from apg_trajectory_tracking_b200.neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from apg_trajectory_tracking_b200.neural_control.environments.drone_env import QuadRotorEnvBase
    from apg_trajectory_tracking_b200.neural_control.models.hutter_model import Net
    from apg_trajectory_tracking_b200.scripts import evaluate_drone as ED
    EV, R, O, golden_params = _eval_mods()
    g = load_golden("eval_rand.npz")
    h, dt = 10, 0.1
    net = Net(15, h, 9, 4 * h)
    with torch.no_grad():
        for p, q in zip(net.parameters(), golden_params(load_golden("conc_quad_kat4.npz"))):
            p.copy_(q)
    net.cuda()
    ds = DS.QuadDataset(np.zeros((6, 12)), np.zeros((6, h, 9)), self_play=1.0)
    ctrl = NetworkWrapper(net, ds, horizon=h, dt=dt, take_every_x=9)
    ev = ED.QuadEvaluator(ctrl, QuadRotorEnvBase(FlightmareDynamics(), dt), ref_length=h, dt=dt, speed_factor=0.4)
    tables = torch.tensor(np.stack([g["gentle_table"], g["fast_reset_table"]]), dtype=torch.float32)
    got = ev.run_eval("rand", nr_test=2, max_steps=80, thresh_div=1.0, thresh_stable=1.0, tables=tables)
    divs = [g["gentle_div"], g["fast_reset_div"]]
    per_run = np.array([d.mean() for d in divs])
    stable = np.array([(d < 1.0).sum() for d in divs])
    assert abs(got[0] - stable.mean()) <= 1e-9 and abs(got[4] - per_run.mean()) <= 1e-3
    assert ctrl.action_counter == 160 and ds.eval_counter == 160 // 9


@pytest.mark.parametrize("n", [128, 1, 63, 65, 300, 1000, 9600, 148 * 128 * 3 + 77])
def test_tc1_tcgen05_path_matches_legacy_mma_path(n):
    _check_tcgen05_vs_legacy(n)


@pytest.mark.parametrize("n", [1, 129, 1000, 148 * 128 * 2 + 5])
def test_tc3_raw_samples_in_the_kernel_prologue_equal_prepared_inputs(n):
    """SURVEY 8f N1 as specified: with in_state = in_ref = NULL the tcgen05 kernels take the RAW (state, reference
    rows) samples and run QuadDataset.prepare_data (dataset.py:155-204) in their prologue; loss, states, actions and
    gradient equal the prepared-input call on the tensors the prepare kernels produce from the same samples."""
    import bench as B
    PR, R, SY, T, DS = _mods()
    h, dt = 10, 0.1
    params = B.default_init("quad", h, seed=n % 31)
    case = SY.quad_case(n, h, dt, seed=n % 29)
    gen = torch.Generator().manual_seed(n)
    pos = (torch.rand(n, 3, generator=gen) * 6 - 3)
    cur_raw = case["cur"].clone()
    cur_raw[:, :3] = pos
    ref_raw = case["ref"].clone()
    ref_raw[:, :, :3] += pos[:, None, :]
    flat = R.flatten_params(params).cuda()
    prep = PR.prepare_quad(cur_raw.cuda(), ref_raw.cuda())
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, "cuda:0")
    assert r.tcgen05
    l0, s0, a0 = r.forward(flat, prep["in_state"], prep["cur"], prep["in_ref"], prep["ref"], want_states=True,
                           want_actions=True)
    l0, g0 = float(l0.item()), r.backward(1.0).clone()
    l1, s1, a1 = r.forward(flat, None, cur_raw.cuda(), None, ref_raw.cuda(), want_states=True, want_actions=True)
    l1, g1 = float(l1.item()), r.backward(1.0)
    torch.cuda.synchronize()
    assert abs(l1 - l0) <= 1e-5 * abs(l0), (l0, l1)
    assert float((a1 - a0).abs().max()) <= 1e-5 and float((s1 - s0).abs().max()) <= 2e-5 * max(float(s0.abs().max()), 1.0)
    assert rel_err(g1, g0) <= 1e-4


@pytest.mark.parametrize("n", [1000, 65536])
def test_tc4_tcgen05_gradient_is_bitwise_reproducible(n):
    """every kernel of the tcgen05 path reduces in a fixed order: the same inputs give the same bits, launch after
    launch (a protocol race - an operand read before it was complete, a stage overwritten early - shows up here as
    run-to-run noise long before it is large enough to fail a tolerance)"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    h, dt = 10, 0.1
    params = B.default_init("quad", h, seed=3)
    case = {k: v.cuda() for k, v in SY.quad_case(n, h, dt, seed=5).items()}
    flat = R.flatten_params(params).cuda()
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, "cuda:0")
    assert r.tcgen05
    grads, losses = [], []
    for rep in range(6):
        loss, _, _ = r.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])
        losses.append(float(loss.item()))
        grads.append(r.backward(1.0).clone())
    torch.cuda.synchronize()
    assert len(set(losses)) == 1, losses
    worst = max(float((g - grads[0]).abs().max()) for g in grads[1:])
    rel = max(rel_err(g, grads[0]) for g in grads[1:])
    assert worst == 0.0, f"gradient differs between identical launches: max abs {worst}, rel l2 {rel}"


def test_tc2_backward_after_a_forward_of_the_other_path_poisons_the_gradient():
    """the workspace is stamped by the forward that filled its stash; an adjoint of the tcgen05 path run on a
    workspace whose last forward was the legacy one must not return a silently wrong gradient"""
    import bench as B
    PR, R, SY, T, DS = _mods()
    n, h, dt = 300, 10, 0.1
    params = B.default_init("quad", h, seed=5)
    case = {k: v.cuda() for k, v in SY.quad_case(n, h, dt, seed=5).items()}
    flat = R.flatten_params(params).cuda()
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, "cuda:0")
    r.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])      # tcgen05 forward: stash is valid
    os.environ["APG_LEGACY_MMA"] = "1"
    try:
        r.forward(flat, case["in_state"], case["cur"], case["in_ref"], case["ref"])  # legacy forward re-stamps
    finally:
        os.environ.pop("APG_LEGACY_MMA", None)
    g = r.backward(1.0)                                                              # tcgen05 adjoint on a legacy stamp
    torch.cuda.synchronize()
    assert not bool(torch.isfinite(g).all())


# ---------------------------------------------------------------------------------------------------------------
# the gradient exchange over NVLink peer memory (optional path, needs 2 GPUs; opt-in like the tcgen05 tests until its
# first hardware run):  APG_TEST_P2P=1 python -m pytest tests/test_zz_new_paths_gpu.py -k p2p
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2 or os.environ.get("APG_TEST_P2P") != "1",
                    reason="needs 2 GPUs and APG_TEST_P2P=1 (first hardware run pending)")
def test_tc3_p2p_gradient_exchange_matches_nccl():
    import json
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_p2p_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", script], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["finite"] and res["p2p_params_bitwise_equal_across_ranks"], res
    assert res["grad_rel_diff_vs_nccl"] <= 1e-6 and res["params_rel_diff_vs_nccl"] <= 1e-6, res
    assert res["loss_rel_diff_vs_nccl"] <= 1e-6, res
