"""GPU parity: the CUDA rollout (through the C-ABI) against the golden vectors of the unmodified reference and
against the CPU oracle on seeded random inputs.

Tolerances (fp32 kernels vs fp32 reference / oracle; different summation order, fused FMAs):
    loss                      rel <= 1e-5
    actions / states          max abs err <= 1e-5 * scale
    parameter gradients       per-tensor L2 rel <= 1e-4
    size-independent checks   additivity over drone subsets rel <= 2e-5, run-to-run bitwise equality
"""
import numpy as np
import pytest
import torch

from tests.helpers import load_golden, golden_params, golden_grads, t, rel_err, max_rel_to_scale

pytestmark = pytest.mark.gpu

LOSS_TOL, ACT_TOL, GRAD_TOL = 1e-5, 1e-5, 1e-4


def _imports():
    from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY, params as P, _capi
    from oracle import apg_oracle as O
    return R, SY, P, _capi, O


def _spec_for(R, system, h, dt):
    if system == "quad":
        return R.RolloutSpec.quad_concurrent(h, dt)
    if system == "wing":
        return R.RolloutSpec.wing_concurrent(h, dt)
    return R.RolloutSpec.cartpole_concurrent(h, dt)


def _run_gpu(R, spec, params, in_state, cur, in_ref, ref, h0c0=None):
    dev = "cuda:0"
    n = cur.shape[0]
    runner = R.Rollout(spec, n, dev)
    flat = R.flatten_params(params).to(dev)
    cu = lambda x: None if x is None else x.to(dev).contiguous()
    loss, states, actions = runner.forward(flat, cu(in_state), cu(cur), cu(in_ref), cu(ref), cu(h0c0), True, True)
    grad = runner.backward(1.0)
    torch.cuda.synchronize()
    return float(loss.item()), states.cpu(), actions.cpu(), R.split_flat(grad.cpu(), params), runner


def _check_grads(grads, ref_grads, tol=GRAD_TOL):
    for i, (a, b) in enumerate(zip(grads, ref_grads)):
        if b is None:
            assert float(a.abs().max()) == 0.0, f"param {i}: unused tensor must get a zero gradient"
        else:
            assert rel_err(a, b) <= tol, (i, rel_err(a, b))


def test_library_loads_and_sees_the_gpu():
    R, SY, P, _capi, O = _imports()
    assert _capi.lib().apg_sm_count() > 0


@pytest.mark.parametrize("name", ["quad", "wing", "cartpole"])
def test_single_step_kernels(name):
    import ctypes
    R, SY, P, _capi, O = _imports()
    g = load_golden("steps.npz")
    lib = _capi.lib()
    dev = "cuda:0"
    s, a = t(g[f"rand_{name}_state"]).to(dev), t(g[f"rand_{name}_action"]).to(dev)
    cot = t(g[f"rand_{name}_cot"]).to(dev)
    dt = float(g[f"rand_{name}_dt"])
    phys = np.ascontiguousarray(P.PHYS[name]())
    out, gs, ga = torch.empty_like(s), torch.empty_like(s), torch.empty_like(a)
    pp = lambda x: ctypes.c_void_p(x.data_ptr())
    _capi.check(lib.apg_dynamics_step(P.SYSTEM_ID[name], ctypes.c_void_p(phys.ctypes.data), pp(s), pp(a),
                                      ctypes.c_float(dt), s.shape[0], pp(out), None))
    _capi.check(lib.apg_dynamics_step_adjoint(P.SYSTEM_ID[name], ctypes.c_void_p(phys.ctypes.data), pp(s), pp(a),
                                              ctypes.c_float(dt), s.shape[0], pp(cot), pp(gs), pp(ga), None))
    torch.cuda.synchronize()
    assert max_rel_to_scale(out.cpu(), g[f"rand_{name}_out"]) <= 3e-6
    assert max_rel_to_scale(gs.cpu(), g[f"rand_{name}_gstate"]) <= 3e-5
    assert max_rel_to_scale(ga.cpu(), g[f"rand_{name}_gaction"]) <= 3e-5


CONC = [("quad", "conc_quad_kat4.npz"), ("quad", "conc_quad_rand.npz"), ("quad", "conc_quad_rand_h6.npz"),
        ("wing", "conc_wing_kat5.npz"), ("wing", "conc_wing_rand_h20.npz"),
        ("cartpole", "conc_cartpole_kat6.npz"), ("cartpole", "conc_cartpole_rand_b128_h5.npz")]


@pytest.mark.parametrize("system,fname", CONC)
def test_concurrent_vs_reference_golden(system, fname):
    R, SY, P, _capi, O = _imports()
    g = load_golden(fname)
    params = golden_params(g)
    h, dt = int(g["h"]), float(g["dt"])
    spec = _spec_for(R, system, h, dt)
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, t(g["in_state"]), t(g["cur"]),
                                               t(g["in_ref"]) if "in_ref" in g else None,
                                               t(g["ref"]) if system != "cartpole" else None)
    assert abs(loss - float(g["loss"])) <= LOSS_TOL * abs(float(g["loss"])), (loss, float(g["loss"]))
    assert max_rel_to_scale(actions, g["actions"]) <= ACT_TOL
    assert max_rel_to_scale(states, g["states"]) <= ACT_TOL
    _check_grads(grads, golden_grads(g))


def _random_hutter(system, h, seed):
    torch.manual_seed(seed)
    if system == "quad":
        shapes = [(64, 15), (64,), (20, 9, 3), (20,), (64, 9 * h), (64,), (64, 64 + 20 * (h - 2)), (64,), (64, 64), (64,),
                  (64, 64), (64,), (4 * h, 64), (4 * h,)]
    else:
        shapes = [(64, 9), (64,), (20, 3, 3), (20,), (64, 3), (64,), (64, 128), (64,), (64, 64), (64,), (64, 64), (64,),
                  (4 * h, 64), (4 * h,)]
    out = []
    for s in shapes:
        fan_in = s[1] * (s[2] if len(s) == 3 else 1) if len(s) > 1 else 64
        out.append((torch.rand(*s) * 2 - 1) / fan_in ** 0.5)
    return out


@pytest.mark.parametrize("system,h,n", [("quad", 10, 1), ("quad", 10, 63), ("quad", 10, 64), ("quad", 10, 65),
                                         ("quad", 10, 1000), ("quad", 5, 130), ("wing", 20, 200), ("wing", 10, 777),
                                         ("quad", 10, 9600)])
def test_concurrent_vs_oracle_random(system, h, n):
    """ragged sizes (partial tiles, fewer tiles than SMs, several tiles per CTA) against the CPU oracle"""
    R, SY, P, _capi, O = _imports()
    dt = 0.1 if system == "quad" else 0.05
    case = SY.quad_case(n, h, dt, seed=7 + n) if system == "quad" else SY.wing_case(n, h, dt, seed=7 + n)
    params = _random_hutter(system, h, seed=n)
    spec = _spec_for(R, system, h, dt)
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, case["in_state"], case["cur"], case["in_ref"],
                                               case["ref"])
    ol, og, ost, oact = O.concurrent_value_and_grad(system, params, case["in_state"], case["cur"], case["in_ref"],
                                                    case["ref"], h, dt)
    assert abs(loss - float(ol)) <= LOSS_TOL * abs(float(ol)), (loss, float(ol))
    assert max_rel_to_scale(actions, oact) <= ACT_TOL
    assert max_rel_to_scale(states, ost) <= ACT_TOL
    _check_grads(grads, og)


@pytest.mark.parametrize("h,n", [(5, 128), (10, 1000), (7, 65)])
def test_cartpole_vs_oracle_random(h, n):
    R, SY, P, _capi, O = _imports()
    dt = 0.05
    case = SY.cartpole_case(n, seed=n)
    torch.manual_seed(n)
    shapes = [(32, 4), (32,), (64, 32), (64,), (64, 64), (64,), (32, 64), (32,), (h, 32), (h,)]
    params = [(torch.rand(*s) * 2 - 1) / (s[-1] if len(s) > 1 else 32) ** 0.5 for s in shapes]
    spec = R.RolloutSpec.cartpole_concurrent(h, dt)
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, case["in_state"], case["cur"], None, None)
    ol, og, ost, oact = O.concurrent_value_and_grad("cartpole", params, case["in_state"], case["cur"], None, None, h,
                                                    dt)
    assert abs(loss - float(ol)) <= LOSS_TOL * abs(float(ol)), (loss, float(ol))
    assert max_rel_to_scale(actions, oact) <= ACT_TOL
    assert max_rel_to_scale(states, ost) <= 2e-5      # theta wraps through atan2: abs error relative to pi
    _check_grads(grads, og)


def _random_ar_params(h, seed):
    torch.manual_seed(seed)
    shapes = [(64, 15), (64,), (20, 9, 3), (20,), (64, 9 * h), (64,), (64, 64 + 20 * (h - 2)), (64,), (64, 64), (64,),
              (64, 64), (64,), (4, 64), (4,)]
    out = []
    for s_ in shapes:
        fan_in = s_[1] * (s_[2] if len(s_) == 3 else 1) if len(s_) > 1 else 64
        out.append((torch.rand(*s_) * 2 - 1) / fan_in ** 0.5)
    return out


@pytest.mark.parametrize("fname", ["rec_ar_rand.npz", "rec_ar_rand_pos0.npz"])
def test_autoregressive_forward_vs_reference_golden(fname):
    """the reference's AR train step only has a forward (its backward() raises): loss / actions / states"""
    R, SY, P, _capi, O = _imports()
    g = load_golden(fname)
    params = golden_params(g)
    h, dt = int(g["h"]), float(g["dt"])
    spec = R.RolloutSpec.quad_recurrent("autoregressive", h, dt, "cumulative")
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, None, t(g["cur"]), t(g["in_ref"]), t(g["ref"]))
    assert abs(loss - float(g["loss"])) <= LOSS_TOL * abs(float(g["loss"])), (loss, float(g["loss"]))
    assert max_rel_to_scale(actions, g["actions"]) <= 2e-5
    assert max_rel_to_scale(states, g["states"]) <= 2e-5
    # gradient oracle: autograd on the forward-pinned functional restatement
    ol, og, _, _ = O.recurrent_value_and_grad("autoregressive", params, t(g["cur"]), t(g["in_ref"]), t(g["ref"]), h, dt,
                                              window="cumulative")
    _check_grads(grads, og, tol=2e-4)


@pytest.mark.parametrize("window,h,n,pos0", [("cumulative", 10, 70, False), ("relative", 10, 70, False),
                                             ("cumulative", 10, 333, True), ("relative", 6, 64, True),
                                             ("cumulative", 5, 1, False)])
def test_autoregressive_vs_oracle_random(window, h, n, pos0):
    R, SY, P, _capi, O = _imports()
    dt = 0.1
    case = SY.quad_case(n, 2 * h, dt, seed=100 + n)
    cur = case["cur"].clone()
    if pos0:
        cur[:, :3] = 0.3 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
        cur[:, 9:12] = 0.2 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n + 1))
    params = _random_ar_params(h, seed=n)
    spec = R.RolloutSpec.quad_recurrent("autoregressive", h, dt, window)
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, None, cur, case["in_ref"], case["ref"])
    ol, og, ost, oact = O.recurrent_value_and_grad("autoregressive", params, cur, case["in_ref"], case["ref"], h, dt,
                                                   window=window)
    assert abs(loss - float(ol)) <= LOSS_TOL * abs(float(ol)), (loss, float(ol))
    assert max_rel_to_scale(actions, oact) <= 2e-5
    assert max_rel_to_scale(states, ost) <= 2e-5
    _check_grads(grads, og, tol=2e-4)


def _random_lstm_params(h, seed):
    torch.manual_seed(seed)
    ih = 15 + 20 * (h - 2)
    shapes = [(20, 9, 3), (20,), (64, 9 * h), (64,), (4, 8), (4,), (32, ih), (32, 8), (32,), (32,)]
    fans = [27, 27, 9 * h, 9 * h, 8, 8, 8, 8, 8, 8]
    return [(torch.rand(*s_) * 2 - 1) / f ** 0.5 for s_, f in zip(shapes, fans)]


def test_lstm_forward_vs_reference_golden():
    R, SY, P, _capi, O = _imports()
    g = load_golden("rec_lstm_rand.npz")
    params = golden_params(g)
    h, dt = int(g["h"]), float(g["dt"])
    h0c0 = torch.stack((t(g["h0"]), t(g["c0"])), 0)
    spec = R.RolloutSpec.quad_recurrent("lstm", h, dt, "cumulative")
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, None, t(g["cur"]), t(g["in_ref"]), t(g["ref"]), h0c0)
    assert abs(loss - float(g["loss"])) <= LOSS_TOL * abs(float(g["loss"])), (loss, float(g["loss"]))
    assert max_rel_to_scale(actions, g["actions"]) <= 2e-5
    assert max_rel_to_scale(states, g["states"]) <= 2e-5
    ol, og, _, _ = O.recurrent_value_and_grad("lstm", params, t(g["cur"]), t(g["in_ref"]), t(g["ref"]), h, dt,
                                              window="cumulative", hc0=(t(g["h0"]), t(g["c0"])))
    _check_grads(grads, og, tol=2e-4)


@pytest.mark.parametrize("window,h,n", [("cumulative", 10, 70), ("relative", 10, 130), ("cumulative", 6, 64),
                                        ("cumulative", 10, 1000)])
def test_lstm_vs_oracle_random(window, h, n):
    R, SY, P, _capi, O = _imports()
    dt = 0.1
    case = SY.quad_case(n, 2 * h, dt, seed=200 + n)
    cur = case["cur"].clone()
    cur[:, :3] = 0.2 * torch.randn(n, 3, generator=torch.Generator().manual_seed(n))
    gen = torch.Generator().manual_seed(n + 5)
    h0, c0 = torch.randn(n, 8, generator=gen), torch.randn(n, 8, generator=gen)
    params = _random_lstm_params(h, seed=n)
    spec = R.RolloutSpec.quad_recurrent("lstm", h, dt, window)
    loss, states, actions, grads, _ = _run_gpu(R, spec, params, None, cur, case["in_ref"], case["ref"],
                                               torch.stack((h0, c0), 0))
    ol, og, ost, oact = O.recurrent_value_and_grad("lstm", params, cur, case["in_ref"], case["ref"], h, dt,
                                                   window=window, hc0=(h0, c0))
    assert abs(loss - float(ol)) <= LOSS_TOL * abs(float(ol)), (loss, float(ol))
    assert max_rel_to_scale(actions, oact) <= 2e-5
    assert max_rel_to_scale(states, ost) <= 2e-5
    _check_grads(grads, og, tol=2e-4)


def test_host_buffer_entry_point_matches_device_path():
    R, SY, P, _capi, O = _imports()
    n, h, dt = 300, 10, 0.1
    case = SY.quad_case(n, h, dt, seed=3)
    params = _random_hutter("quad", h, seed=11)
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    loss_d, _, _, grads_d, _ = _run_gpu(R, spec, params, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    loss_h, grad_h = R.value_and_grad_host(spec, R.flatten_params(params), case["in_state"], case["cur"],
                                           case["in_ref"], case["ref"])
    assert loss_h == loss_d
    assert np.array_equal(grad_h, R.flatten_params(grads_d).numpy())


def test_full_size_properties():
    """BASELINE size (N=65536, h=10): run-to-run bitwise determinism and additivity of loss / gradient over a
    partition of the drones (the loss is a plain sum over drones: drone_loss.py:22-33)."""
    R, SY, P, _capi, O = _imports()
    n, h, dt = 65536, 10, 0.1
    case = SY.quad_case(n, h, dt, seed=5)
    params = _random_hutter("quad", h, seed=5)
    spec = R.RolloutSpec.quad_concurrent(h, dt)
    loss, _, _, grads, _ = _run_gpu(R, spec, params, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    loss2, _, _, grads2, _ = _run_gpu(R, spec, params, case["in_state"], case["cur"], case["in_ref"], case["ref"])
    assert loss == loss2
    for a, b in zip(grads, grads2):
        assert torch.equal(a, b)
    cut = 40000
    parts = []
    for sl in (slice(0, cut), slice(cut, n)):
        parts.append(_run_gpu(R, spec, params, case["in_state"][sl], case["cur"][sl], case["in_ref"][sl],
                              case["ref"][sl]))
    assert abs(parts[0][0] + parts[1][0] - loss) <= 2e-5 * abs(loss)
    for gfull, ga, gb in zip(grads, parts[0][3], parts[1][3]):
        if float(gfull.abs().max()) > 0:
            assert rel_err(ga + gb, gfull) <= 2e-5
    # a 512-drone sample of the full batch against the oracle
    idx = torch.arange(0, n, n // 512)
    sub = {k: v[idx] for k, v in case.items()}
    ls, _, acts, _, _ = _run_gpu(R, spec, params, sub["in_state"], sub["cur"], sub["in_ref"], sub["ref"])
    ol, _, _, oact = O.concurrent_value_and_grad("quad", params, sub["in_state"], sub["cur"], sub["in_ref"],
                                                 sub["ref"], h, dt)
    assert abs(ls - float(ol)) <= LOSS_TOL * abs(float(ol))
    assert max_rel_to_scale(acts, oact) <= ACT_TOL


@pytest.mark.parametrize("config", ["wing_131072", "autoregressive_65536", "lstm_32768"])
def test_full_size_properties_of_the_other_baseline_configurations(config):
    """BASELINE.json configs 3-5 at their per-GPU sizes (fixed wing h = 20 N = 131072; quad autoregressive h = 10
    N = 65536; quad LSTM h = 10, 32768 of the 262144 drones per GPU): run-to-run bitwise determinism, additivity of loss
    and gradient over a partition of the drones, and a 256-drone sample of the SAME batch against the oracle - loss,
    actions, states AND the parameter gradient."""
    R, SY, P, _capi, O = _imports()
    import bench as B
    kind, n = config.split("_")[0], int(config.split("_")[1])
    if kind == "wing":
        h, dt = 20, 0.05
        case = SY.wing_case(n, h, dt, seed=3)
        params = B.default_init("wing", h, seed=3)
        spec = R.RolloutSpec.wing_concurrent(h, dt)
        args = lambda sl: (case["in_state"][sl], case["cur"][sl], case["in_ref"][sl], case["ref"][sl], None)   # noqa: E731
        oracle = lambda a: O.concurrent_value_and_grad("wing", params, a[0], a[1], a[2], a[3], h, dt)          # noqa: E731
    else:
        h, dt = 10, 0.1
        mode = "autoregressive" if kind == "autoregressive" else "lstm"
        case = SY.quad_case(n, 2 * h, dt, seed=4)
        params = B.default_init("quad", h, seed=4, mode=mode)
        spec = R.RolloutSpec.quad_recurrent(mode, h, dt, "cumulative")
        gen = torch.Generator().manual_seed(9)
        hc = torch.stack((torch.randn(n, 8, generator=gen), torch.randn(n, 8, generator=gen)), 0) if mode == "lstm" else None
        args = lambda sl: (None, case["cur"][sl], case["in_ref"][sl], case["ref"][sl],                          # noqa: E731
                           None if hc is None else hc[:, sl].contiguous())
        oracle = lambda a: O.recurrent_value_and_grad(mode, params, a[1], a[2], a[3], h, dt, window="cumulative",  # noqa: E731
                                                      hc0=None if a[4] is None else (a[4][0], a[4][1]))
    full = slice(0, n)
    loss, _, _, grads, _ = _run_gpu(R, spec, params, *args(full))
    loss2, _, _, grads2, _ = _run_gpu(R, spec, params, *args(full))
    assert loss == loss2 and all(torch.equal(a, b) for a, b in zip(grads, grads2))
    cut = (n * 5 // 8) // 64 * 64
    parts = [_run_gpu(R, spec, params, *args(sl)) for sl in (slice(0, cut), slice(cut, n))]
    assert abs(parts[0][0] + parts[1][0] - loss) <= 2e-5 * abs(loss)
    for gfull, ga, gb in zip(grads, parts[0][3], parts[1][3]):
        if float(gfull.abs().max()) > 0:
            # the gradient is a sum over drones whose terms cancel: the bar is relative to the size of the two partial
            # sums (what fp32 reassociation can move), not to the norm of their sum
            bar = 2e-5 * float(gfull.norm()) + 2e-6 * float((ga.abs() + gb.abs()).norm())
            assert float((ga + gb - gfull).norm()) <= bar
    idx = torch.arange(0, n, n // 256)
    sub = args(idx)
    ls, sts, acts, gs, _ = _run_gpu(R, spec, params, *sub)
    ol, og, ost, oact = oracle(sub)
    assert abs(ls - float(ol)) <= LOSS_TOL * abs(float(ol)), (ls, float(ol))
    assert max_rel_to_scale(acts, oact) <= 2e-5 and max_rel_to_scale(sts, ost) <= 2e-5
    _check_grads(gs, og, tol=2e-4)
