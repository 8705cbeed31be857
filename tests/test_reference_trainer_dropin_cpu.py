"""Drop-in check of SURVEY.md 8b on the reference's OWN trainer code (needs the reference checkout, i.e. the build
container; skipped elsewhere).  With this repository first on sys.path and APG_REFERENCE_ROOT set, the reference's
unmodified ``scripts/train_drone.py`` / ``train_base.py`` import (``neural_control.*`` resolves to the mirror, every
module the mirror does not carry to the reference's own file: the overlay of neural_control/__init__.py), and

    TrainDrone(train_dynamics, eval_dynamics, config).initialize_model();  .run_epoch()

run as written: ``QuadDataset(num_states, self_play, **config)`` with the reference's constructor, the policy module,
``FlightmareDynamics.__call__`` per step, ``quad_mpc_loss``, ``loss.backward()``, SGD - every kernel behind it on the
CPU model library (libapg_b200_sim.so; no GPU here).  The epoch's batch losses are then reproduced with the oracle on
the same batches in the same order.  Second test: the reference's shipped pickles load against the mirror classes."""
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import pytest
import torch

from tests.test_capi_sim_host import simlib, simlib_path  # noqa: F401  (fixtures: the library on the CPU models)

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "scripts")),
                                reason="needs the reference checkout (build container only)")
_STUBS = ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d",
          "pyglet", "pyglet.gl", "pyquaternion", "ruamel", "ruamel.yaml", "gym", "gym.utils", "gym.spaces"]
_REF_SCRIPT_MODULES = ["train_drone", "train_base", "evaluate_drone", "evaluate_base", "train_fixed_wing",
                       "evaluate_fixed_wing", "train_cartpole", "evaluate_cartpole"]


@pytest.fixture
def reference_scripts(monkeypatch, tmp_path):
    """the reference's scripts importable against the overlay; everything undone afterwards"""
    monkeypatch.setenv("APG_REFERENCE_ROOT", REF)
    monkeypatch.chdir(tmp_path)                      # the trainer writes trained_models/ and runs/ into the cwd
    saved = {k: v for k, v in sys.modules.items() if k == "neural_control" or k.startswith("neural_control.")}
    for k in saved:
        del sys.modules[k]
    for name in _STUBS:                              # optional third-party packages of the reference, absent here
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, MagicMock())
    sys.modules["gym"].Env = type("Env", (), {})
    monkeypatch.syspath_prepend(os.path.join(REF, "scripts"))
    import neural_control                            # noqa: F401  re-import with the overlay active
    yield
    for k in [k for k in sys.modules if k == "neural_control" or k.startswith("neural_control.") or
              k.startswith("_apg_reference_twin") or k in _REF_SCRIPT_MODULES]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_reference_train_drone_runs_on_the_mirror_and_matches_the_oracle(simlib, reference_scripts, monkeypatch):
    from oracle import apg_oracle as O
    from apg_trajectory_tracking_b200.neural_control import environments as ENV
    monkeypatch.setattr(ENV, "compute_device", lambda: torch.device("cpu"))        # "device" memory of the model library
    import importlib
    for name in ("models.hutter_model", "models.rnn", "models.simple_model", "drone_loss", "dataset",
                 "dynamics.quad_dynamics_flightmare", "dynamics.fixed_wing_dynamics", "dynamics.cartpole_dynamics"):
        mod = importlib.import_module("apg_trajectory_tracking_b200.neural_control." + name)
        monkeypatch.setattr(mod, "_require_cuda", lambda *a, **k: None, raising=False)   # host pointers ARE device pointers
    import train_drone                                                             # the REFERENCE's file
    assert train_drone.__file__.startswith(REF)
    assert train_drone.QuadDataset.__module__.startswith("apg_trajectory_tracking_b200.")
    assert train_drone.Net.__module__.startswith("apg_trajectory_tracking_b200.")
    with open(os.path.join(REF, "configs", "quad_config.json")) as f:
        config = json.load(f)
    config.update(epoch_size=16, self_play=0.5, batch_size=8, save_name="dropin_test", sample_in="train_env")
    mp = config["modified_params"]
    np.random.seed(3)
    torch.manual_seed(3)
    trainer = train_drone.TrainDrone(train_drone.FlightmareDynamics(modified_params=mp),
                                     train_drone.FlightmareDynamics(modified_params=mp), config)
    trainer.initialize_model()                       # QuadDataset(epoch_size, self_play, **config), Net, SGD
    ds = trainer.state_data
    assert len(ds) == 24 and ds.num_sampled_states == 16 and ds.num_self_play == 8
    assert tuple(ds.normed_states.shape) == (24, 15) and tuple(ds.in_ref_states.shape) == (24, 10, 9)
    params0 = [p.detach().clone() for p in trainer.net.parameters()]
    torch.manual_seed(11)
    got = trainer.run_epoch(train="controller")      # the reference's loop, unchanged
    # the same epoch on the oracle: same batches (same DataLoader seed), SGD momentum 0.9, lr from the config
    torch.manual_seed(11)
    loader = torch.utils.data.DataLoader(ds, batch_size=8, shuffle=True, num_workers=0)
    params = [p.clone() for p in params0]
    bufs = [None] * len(params)
    lr, running, i = config["learning_rate_controller"], 0.0, 0
    for i, (in_state, cur, in_ref, ref) in enumerate(loader):
        loss, grads, _, _ = O.concurrent_value_and_grad("quad", params, in_state, cur, in_ref, ref, 10, config["delta_t"])
        running += float(loss)
        for k, g in enumerate(grads):
            if g is None:
                continue
            bufs[k] = g.clone() if bufs[k] is None else bufs[k] * 0.9 + g
            params[k] = params[k] - lr * bufs[k]
    want = running / i
    assert abs(got - want) <= 2e-5 * abs(want), (got, want)
    for p, q in zip(trainer.net.parameters(), params):
        assert float((p.detach() - q).abs().max()) <= 1e-6 * max(float(q.abs().max()), 1.0)
    # resampling through the reference's own entry point
    trainer.state_data.resample_data()
    assert len(trainer.state_data) == 24


def test_reference_pickles_load_against_the_mirror_classes():
    """trained_models/*/current_model/model_* were written with torch.save(net): unpickling resolves
    neural_control.models.* to the mirror, with the reference's parameter names and shapes"""
    import neural_control  # noqa: F401
    want = {"quad": ("model_quad", "hutter_model", 14), "wing": ("model_wing", "hutter_model", 14),
            "cartpole": ("model_cartpole", "simple_model", 10)}
    for system, (fname, module, n_tensors) in want.items():
        path = os.path.join(REF, "trained_models", system, "current_model", fname)
        net = torch.load(path, weights_only=False)
        assert type(net).__module__ == f"apg_trajectory_tracking_b200.neural_control.models.{module}", type(net)
        assert len(list(net.parameters())) == n_tensors
