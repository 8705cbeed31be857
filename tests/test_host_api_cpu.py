"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/apg_b200.h declares,
its host-only entry points (layout arithmetic) answer correctly, and the Python mirror refuses CPU tensors
(no CPU fallback).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from apg_trajectory_tracking_b200 import build as B, _capi
    B.build()
    return _capi


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "apg_b200.h")).read()
    declared = set(re.findall(r"\b(apg_[a-z_0-9]+)\s*\(", hdr))
    assert declared >= {"apg_rollout_forward", "apg_rollout_backward", "apg_rollout_value_and_grad_host",
                        "apg_dynamics_step", "apg_dynamics_step_adjoint", "apg_num_params", "apg_workspace_bytes"}
    raw = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/apg_b200.h but not exported"
    assert set(capi.EXPORTS) == declared
    assert capi.lib().apg_version() >= 1
    assert b"success" in capi.lib().apg_error_string(0)


@pytest.mark.parametrize("fname,spec_args,expected", [
    ("conc_quad_rand.npz", ("quad", "concurrent", 10), 32728),
    ("conc_wing_rand_h20.npz", ("wing", "concurrent", 20), 22872),
    ("conc_cartpole_kat6.npz", ("cartpole", "concurrent", 10), 8842),
    ("rec_ar_rand.npz", ("quad", "autoregressive", 10), 30388),
    ("rec_lstm_rand.npz", ("quad", "lstm", 10), 12340),
])
def test_param_counts_match_reference_models(capi, fname, spec_args, expected):
    """apg_num_params == sum of net.parameters() sizes of the reference nets"""
    from apg_trajectory_tracking_b200 import rollout as R
    g = load_golden(fname)
    n_ref = sum(g[f"param_{i}"].size for i in range(len(g["param_names"])))
    system, mode, h = spec_args
    if mode == "concurrent":
        spec = {"quad": R.RolloutSpec.quad_concurrent, "wing": R.RolloutSpec.wing_concurrent,
                "cartpole": R.RolloutSpec.cartpole_concurrent}[system](h)
    else:
        spec = R.RolloutSpec.quad_recurrent(mode, h)
    cfg = spec.config(128)
    n = capi.lib().apg_num_params(ctypes.byref(cfg))
    assert n == n_ref == expected
    assert capi.lib().apg_workspace_bytes(ctypes.byref(cfg)) > 0


def test_bad_configs_are_rejected(capi):
    from apg_trajectory_tracking_b200 import rollout as R
    lib = capi.lib()
    cfg = R.RolloutSpec.quad_concurrent(10).config(16)
    cfg.out_dim = 39
    assert lib.apg_num_params(ctypes.byref(cfg)) == -1          # APG_ERR_BAD_CONFIG
    cfg = R.RolloutSpec.wing_concurrent(10).config(16)
    cfg.mode = 1
    assert lib.apg_num_params(ctypes.byref(cfg)) == -2          # APG_ERR_UNSUPPORTED
    cfg = R.RolloutSpec.quad_concurrent(10).config(0)
    assert lib.apg_num_params(ctypes.byref(cfg)) == -1
    with pytest.raises(capi.ApgError):
        capi.check(-2)


def test_module_mirror_has_reference_parameter_names_and_shapes():
    import neural_control  # noqa: F401  (top-level alias of the mirror package)
    from neural_control.models.hutter_model import Net
    from neural_control.models.rnn import LSTM_NEW
    from neural_control.models.simple_model import Net as SimpleNet
    for fname, net in (("conc_quad_rand.npz", Net(15, 10, 9, 40)), ("conc_wing_rand_h20.npz", Net(9, 1, 3, 80, conv=False)),
                       ("conc_cartpole_kat6.npz", SimpleNet(4, 10)), ("rec_ar_rand.npz", Net(15, 10, 9, 4)),
                       ("rec_lstm_rand.npz", LSTM_NEW(15, 10, 9, 4))):
        g = load_golden(fname)
        names = [str(x) for x in g["param_names"]]
        mine = list(net.named_parameters())
        assert [n for n, _ in mine] == names
        assert [tuple(p.shape) for _, p in mine] == [g[f"param_{i}"].shape for i in range(len(names))]


def test_no_cpu_fallback():
    from apg_trajectory_tracking_b200 import _capi
    from neural_control.models.hutter_model import Net
    from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics
    from neural_control.dynamics.cartpole_dynamics import CartpoleDynamics
    from neural_control.drone_loss import quad_mpc_loss
    from neural_control.dataset import state_preprocessing
    with pytest.raises(_capi.ApgError):
        Net(15, 10, 9, 40)(torch.zeros(2, 15), torch.zeros(2, 10, 9))
    with pytest.raises(_capi.ApgError):
        FlightmareDynamics()(torch.zeros(2, 12), torch.zeros(2, 4), 0.1)
    with pytest.raises(_capi.ApgError):
        CartpoleDynamics()(torch.zeros(2, 4), torch.zeros(2, 1), 0.05)
    with pytest.raises(_capi.ApgError):
        state_preprocessing(torch.zeros(2, 12))
    with pytest.raises(_capi.ApgError):
        quad_mpc_loss(torch.zeros(2, 10, 12), torch.zeros(2, 10, 9), torch.zeros(2, 10, 4))
    if not torch.cuda.is_available():
        from apg_trajectory_tracking_b200 import rollout as R
        with pytest.raises(_capi.ApgError):
            R.Rollout(R.RolloutSpec.quad_concurrent(10), 8)


def test_physical_constants_match_reference_config():
    from apg_trajectory_tracking_b200 import params as P
    q = P.quad_phys()
    assert abs(q[0] - 0.723) < 1e-7 and abs(q[1] - 0.723 / 12 * 0.31 ** 2 * 4.5) < 1e-8 and q[9] == np.float32(-9.81)
    q2 = P.quad_phys({"mass": 1.0})
    assert q2[0] == 1.0 and q2[1] != q[1]
    w = P.wing_phys()
    assert abs(w[4] + 0.00105) < 1e-9 and abs(w[40] - 0.16534698176788384) < 1e-7
    c = P.cartpole_phys()
    assert list(c[:5]) == [1.0, np.float32(0.1), 0.5, 30.0, 0.5]


def test_dataset_layouts_match_reference_prepare_data():
    """QuadDataset / WingDataset.prepare_data mirrors against the reference's own prepare_data output
    (tests/golden/prep_data.npz, produced by oracle/make_golden.py from the unmodified reference)"""
    from neural_control.dataset import QuadDataset, WingDataset, CartpoleDataset
    g = load_golden("prep_data.npz")
    q = QuadDataset(g["quad_raw_states"], g["quad_raw_refs"])
    for mine, key in ((q.normed_states, "quad_in_state"), (q.states, "quad_states"), (q.in_ref_states, "quad_in_ref"),
                      (q.ref_states, "quad_ref")):
        np.testing.assert_allclose(mine.numpy(), g[key], rtol=0, atol=2e-6)
    a, b, c, d = q[2]
    assert a.shape == (15,) and b.shape == (12,) and c.shape == (10, 9) and d.shape == (10, 9) and len(q) == 6
    w = WingDataset(g["wing_raw_states"], g["wing_targets"], mean=g["wing_mean"], std=g["wing_std"],
                    delta_t=float(g["wing_dt"]), horizon=int(g["wing_h"]))
    for mine, key in ((w.normed_states, "wing_in_state"), (w.states, "wing_states"), (w.in_ref_states, "wing_in_ref"),
                      (w.ref_states, "wing_ref")):
        np.testing.assert_allclose(mine.numpy(), g[key], rtol=0, atol=3e-6)
    # single-sample path + self-play replacement keep the layouts
    one = q.get_and_add_eval_data(g["quad_raw_states"][0], g["quad_raw_refs"][0])
    np.testing.assert_allclose(one[0].numpy(), g["quad_in_state"][:1], atol=2e-6)
    c = CartpoleDataset(np.random.RandomState(0).randn(5, 4))
    s, l = c[1]
    assert torch.equal(s, l) and len(c) == 5


def test_chunk_bounds_cover_the_batch_on_tile_boundaries():
    """host-side logic of FusedTrainStep.step_host: chunks cover [0, n) without overlap, start on 64-drone tiles"""
    from apg_trajectory_tracking_b200.train import chunk_bounds
    for n, chunk in [(65536, 18944), (65536, 9472), (1000, 192), (130, 64), (64, 64), (63, 64), (10, 3), (777, 0),
                     (5, 100000), (200, 100)]:
        b = chunk_bounds(n, chunk)
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(a % 64 == 0 and a < e for a, e in b)
        if 0 < chunk < n:
            assert max(e - a for a, e in b) <= max(64, chunk)
    assert chunk_bounds(0, 64) == []


def test_input_side_ops_refuse_cpu_tensors():
    from apg_trajectory_tracking_b200 import prepare as PR
    from apg_trajectory_tracking_b200._capi import ApgError
    with pytest.raises(ApgError):
        PR.prepare_quad(torch.zeros(4, 12), torch.zeros(4, 10, 9))
    with pytest.raises(ApgError):
        PR.prepare_wing(torch.zeros(4, 12), torch.zeros(4, 3), torch.zeros(12), torch.ones(12), 0.05, 10)
    with pytest.raises(ApgError):
        PR.sample_windows(torch.zeros(100, 9), 3, 10, 20)
    with pytest.raises(ApgError):
        PR.poly_reference(torch.zeros(4, 3, 6), 10, 0.1)


def test_learnt_dynamics_mirror_has_reference_parameters_and_refuses_cpu(capi):
    from neural_control.dynamics.quad_dynamics_trained import LearntDynamics
    from apg_trajectory_tracking_b200._capi import ApgError
    g = load_golden("learnt_dyn.npz")
    d = LearntDynamics()
    names = [n for n, _ in d.named_parameters()]
    assert names == [str(x) for x in g["param_names"]]
    for i, (_, p) in enumerate(d.named_parameters()):
        assert tuple(p.shape) == g[f"a_param_{i}"].shape
    assert sum(p.numel() for p in d.parameters()) == capi.lib().apg_learnt_num_params(0) == 1891
    assert torch.equal(d.linear_at.detach(), torch.eye(4)) and float(d.linear_state_2.weight.abs().max()) == 0.0
    with pytest.raises(ApgError):
        d(torch.zeros(2, 12), torch.zeros(2, 4), 0.1)


def test_new_entry_points_validate_their_arguments_before_touching_the_device(capi):
    """error codes, not crashes: null pointers / bad shapes are rejected on the host; with valid arguments the calls
    stop at 'no CUDA device' here (nothing is computed on the CPU)"""
    import ctypes as C
    from apg_trajectory_tracking_b200 import rollout as R
    lib = capi.lib()
    buf = (C.c_float * 4096)()
    pb = C.c_void_p(C.addressof(buf))
    BAD, UNSUP, NODEV = -1, -2, -4
    assert lib.apg_prepare_quad(None, pb, 4, 10, pb, pb, pb, pb, None) == BAD
    assert lib.apg_prepare_quad(pb, None, 4, 10, None, None, pb, None, None) == BAD
    assert lib.apg_prepare_wing(pb, pb, None, pb, C.c_float(0.05), 10, 4, pb, pb, pb, pb, None) == BAD
    assert lib.apg_prepare_wing(pb, pb, pb, pb, C.c_float(0.05), 0, 4, pb, pb, pb, pb, None) == BAD
    assert lib.apg_sample_windows(pb, 100, 8, 10, 20, 3, pb, pb, None) == BAD          # fewer than 9 columns
    assert lib.apg_sample_windows(pb, 100, 9, 10, 20, 6, pb, pb, None) == BAD          # would read past the table
    assert lib.apg_poly_reference(None, 4, 10, C.c_float(0.1), C.c_float(0.1), pb, None) == BAD
    assert lib.apg_learnt_step(0, None, pb, pb, pb, C.c_float(0.1), 4, pb, None) == BAD
    assert lib.apg_learnt_step(2, pb, pb, pb, pb, C.c_float(0.1), 4, pb, None) == UNSUP          # cartpole
    assert lib.apg_learnt_step_adjoint(1, pb, pb, pb, pb, C.c_float(0.1), 4, pb, pb, pb, pb, None, None) == BAD
    assert lib.apg_learnt_workspace_bytes(0, 1000) >= 1891 * 4 and lib.apg_learnt_workspace_bytes(1, 1000) >= 1914 * 4
    assert lib.apg_learnt_num_params(1) == 1914 and lib.apg_learnt_num_params(2) == UNSUP
    ws = C.c_void_p((C.addressof(buf) + 255) & ~255)
    quad = R.RolloutSpec.quad_concurrent(10, 0.1).config(8)
    wing = R.RolloutSpec.wing_concurrent(10, 0.05).config(8)
    cart = R.RolloutSpec.cartpole_concurrent(5, 0.05).config(8)
    ev = lambda cfg, steps=10, rows=20: lib.apg_eval_rollout(C.byref(cfg), pb, pb, None, 8, rows, pb, steps,   # noqa: E731
                                                             C.c_float(1), C.c_float(1), 0, ws, None, None, None,
                                                             None, None)
    assert ev(cart) == UNSUP and ev(wing) == UNSUP
    assert ev(quad, steps=0) == BAD and ev(quad, rows=0) == BAD
    assert ev(quad) == NODEV
    fly = lambda cfg, k=1: lib.apg_eval_fly_to_points(C.byref(cfg), pb, pb, k, pb, pb, pb, C.c_float(0.05), 10,   # noqa: E731
                                                      C.c_float(4), C.c_float(0.4), 0, ws, None, None, None, None,
                                                      None, None, None)
    assert fly(quad) == UNSUP and fly(wing, k=0) == BAD and fly(wing) == NODEV
    for code in (BAD, UNSUP, -3, NODEV):
        assert lib.apg_error_string(code).startswith(b"apg:")


def test_learnt_wing_mirror_has_reference_parameters_and_refuses_cpu():
    from neural_control.dynamics.fixed_wing_dynamics import LearntFixedWingDynamics
    from apg_trajectory_tracking_b200._capi import ApgError
    g = load_golden("learnt_dyn.npz")
    d = LearntFixedWingDynamics()
    assert [n for n, _ in d.named_parameters()] == [str(x) for x in g["wing_param_names"]]
    for i, (n, p) in enumerate(d.named_parameters()):
        assert tuple(p.shape) == g[f"wa_param_{i}"].shape
        if "linear" not in n:
            assert np.allclose(p.detach().numpy(), g[f"wa_param_{i}"]), n        # the shipped constants
    assert d._flat().numel() == 1914
    with pytest.raises(ApgError):
        d(torch.zeros(2, 12), torch.zeros(2, 4), 0.05)
