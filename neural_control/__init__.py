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
__path__ = list(_impl.__path__)

# Overlay: with APG_REFERENCE_ROOT pointing at a checkout of the reference, every module
# this package does NOT mirror (plotting, rendering, trajectory generators, MPC / PPO baselines ...) resolves to the
# reference's own file, while the mirrored ones above keep resolving to the CUDA implementations - so the reference's
# scripts import and run unchanged.
import os as _os

_ref = _os.environ.get("APG_REFERENCE_ROOT")
if _ref and _os.path.isdir(_os.path.join(_ref, "neural_control")):
    import importlib.util as _ilu

    _ref_pkg = _os.path.join(_ref, "neural_control")
    __path__.append(_ref_pkg)
    for _sub in ("models", "dynamics", "controllers", "environments"):
        _m = sys.modules[__name__ + "." + _sub]
        _d = _os.path.join(_ref_pkg, _sub)
        if _os.path.isdir(_d) and _d not in list(_m.__path__):
            _m.__path__ = list(_m.__path__) + [_d]

    # A mirrored module only carries the hot-path names (e.g. FlightmareDynamics, not the casadi twin
    # FlightmareDynamicsMPC of the MPC baseline).  Names it does not define are looked up in the reference's file of
    # the same module, loaded on first use under a private name (PEP 562 module __getattr__).
    _twins = {}

    def _twin_getattr(_modname):
        def _getattr(name):
            if name.startswith("__"):
                raise AttributeError(name)
            twin = _twins.get(_modname)
            if twin is None:
                path = _os.path.join(_ref_pkg, *_modname.split(".")) + ".py"
                if not _os.path.exists(path):
                    raise AttributeError(f"module neural_control.{_modname} has no attribute {name!r}")
                spec = _ilu.spec_from_file_location("_apg_reference_twin." + _modname, path)
                twin = _ilu.module_from_spec(spec)
                _twins[_modname] = twin
                spec.loader.exec_module(twin)
            try:
                return getattr(twin, name)
            except AttributeError:
                raise AttributeError(f"module neural_control.{_modname} has no attribute {name!r}") from None
        return _getattr

    for _name in list(sys.modules):
        if _name.startswith(__name__ + ".") and not hasattr(sys.modules[_name], "__path__"):
            _mod = sys.modules[_name]
            if "__getattr__" not in vars(_mod):
                _mod.__getattr__ = _twin_getattr(_name[len(__name__) + 1:])
