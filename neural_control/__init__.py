"""Import-path alias so code (and pickles) written against the reference's ``neural_control`` package resolve to
the B200 implementations: ``neural_control.models.hutter_model.Net`` etc."""
import importlib
import sys

_impl = importlib.import_module("apg_trajectory_tracking_b200.neural_control")
for _name in ("models", "models.hutter_model", "models.rnn", "models.simple_model", "dynamics",
              "dynamics.quad_dynamics_base", "dynamics.quad_dynamics_flightmare", "dynamics.quad_dynamics_trained",
              "dynamics.fixed_wing_dynamics",
              "dynamics.cartpole_dynamics", "drone_loss", "dataset", "controllers", "controllers.network_wrapper",
              "environments", "environments.drone_env", "environments.wing_env", "environments.cartpole_env"):
    _m = importlib.import_module("apg_trajectory_tracking_b200.neural_control." + _name)
    sys.modules[__name__ + "." + _name] = _m
    if "." not in _name:
        globals()[_name] = _m
__path__ = _impl.__path__
