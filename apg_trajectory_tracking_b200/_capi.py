"""ctypes binding of libapg_b200.so (include/apg_b200.h).  No CPU fallback: if the library is missing or there is
no CUDA device the product path raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("APG_B200_LIB", os.path.join(HERE, "libapg_b200.so"))
MAX_PHYS = 48

c_float_p = ctypes.c_void_p   # device / host pointers are passed as raw addresses


class ApgConfig(ctypes.Structure):
    _fields_ = [("system", ctypes.c_int), ("mode", ctypes.c_int), ("window", ctypes.c_int), ("net", ctypes.c_int),
                ("n_drones", ctypes.c_int), ("horizon", ctypes.c_int), ("state_feat", ctypes.c_int),
                ("ref_len", ctypes.c_int), ("ref_dim", ctypes.c_int), ("out_dim", ctypes.c_int),
                ("dt", ctypes.c_float), ("phys", ctypes.c_float * MAX_PHYS)]


class ApgGradComm(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int), ("world", ctypes.c_int), ("slot_ptrs", ctypes.c_void_p),
                ("flag_ptrs", ctypes.c_void_p), ("epoch", ctypes.c_uint), ("ticket", ctypes.c_void_p)]


EXPORTS = {
    "apg_version": (ctypes.c_int, []),
    "apg_sm_count": (ctypes.c_int, []),
    "apg_rollout_kernel_path": (ctypes.c_int, [ctypes.POINTER(ApgConfig)]),
    "apg_debug_timing": (ctypes.c_int, [ctypes.c_int]),
    "apg_debug_kernel_times": (ctypes.c_int, [ctypes.c_void_p]),
    "apg_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "apg_num_params": (ctypes.c_int, [ctypes.POINTER(ApgConfig)]),
    "apg_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(ApgConfig)]),
    "apg_rollout_forward": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 6 + [ctypes.c_void_p] +
                            [c_float_p] * 3 + [ctypes.c_void_p]),
    "apg_rollout_forward_learnt": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 6 + [ctypes.c_void_p] +
                                   [c_float_p] * 3 + [ctypes.c_void_p]),
    "apg_rollout_backward": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 6 + [ctypes.c_void_p,
                             ctypes.c_float, c_float_p, ctypes.c_void_p]),
    "apg_rollout_backward_sgd": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 6 + [ctypes.c_void_p,
                                 ctypes.c_float, c_float_p, c_float_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    "apg_rollout_value_and_grad_host": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 8),
    "apg_grad_comm_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "apg_grad_comm_offsets": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "apg_rollout_backward_p2p": (ctypes.c_int, [ctypes.POINTER(ApgConfig)] + [c_float_p] * 6 + [ctypes.c_void_p,
                                 ctypes.c_float, ctypes.POINTER(ApgGradComm), ctypes.c_void_p]),
    "apg_grad_gather_sgd_p2p": (ctypes.c_int, [ctypes.POINTER(ApgGradComm), ctypes.c_void_p, ctypes.c_int, c_float_p,
                                               c_float_p, c_float_p, ctypes.c_float, ctypes.c_float,
                                               ctypes.c_void_p]),
    "apg_dynamics_step": (ctypes.c_int, [ctypes.c_int, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.c_int,
                                         c_float_p, ctypes.c_void_p]),
    "apg_dynamics_step_adjoint": (ctypes.c_int, [ctypes.c_int, c_float_p, c_float_p, c_float_p, ctypes.c_float,
                                                 ctypes.c_int, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "apg_quad_features": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "apg_quad_features_adjoint": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "apg_eval_rollout": (ctypes.c_int, [ctypes.POINTER(ApgConfig), c_float_p, c_float_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_int, c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_int, ctypes.c_void_p, c_float_p, c_float_p, c_float_p,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "apg_eval_rollout_lstm": (ctypes.c_int, [ctypes.POINTER(ApgConfig), c_float_p, c_float_p, c_float_p,
                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_int,
                                             ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, c_float_p,
                                             c_float_p, c_float_p, ctypes.c_void_p, c_float_p, ctypes.c_void_p]),
    "apg_eval_fly_to_points": (ctypes.c_int, [ctypes.POINTER(ApgConfig), c_float_p, c_float_p, ctypes.c_int, c_float_p,
                                              c_float_p, c_float_p, ctypes.c_float, ctypes.c_int, ctypes.c_float,
                                              ctypes.c_float, ctypes.c_int, ctypes.c_void_p, c_float_p, c_float_p,
                                              c_float_p, ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "apg_eval_cartpole": (ctypes.c_int, [ctypes.POINTER(ApgConfig), c_float_p, c_float_p, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_int, ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_void_p,
                                         c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "apg_learnt_num_params": (ctypes.c_int, [ctypes.c_int]),
    "apg_learnt_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "apg_learnt_step": (ctypes.c_int, [ctypes.c_int] + [c_float_p] * 4 + [ctypes.c_float, ctypes.c_int, c_float_p,
                                                                        ctypes.c_void_p]),
    "apg_learnt_step_adjoint": (ctypes.c_int, [ctypes.c_int] + [c_float_p] * 4 + [ctypes.c_float, ctypes.c_int] +
                                [c_float_p] * 4 + [ctypes.c_void_p, ctypes.c_void_p]),
    "apg_prepare_quad": (ctypes.c_int, [c_float_p, c_float_p, ctypes.c_int, ctypes.c_int] + [c_float_p] * 4 +
                         [ctypes.c_void_p]),
    "apg_prepare_wing": (ctypes.c_int, [c_float_p] * 4 + [ctypes.c_float, ctypes.c_int, ctypes.c_int] +
                         [c_float_p] * 4 + [ctypes.c_void_p]),
    "apg_sample_windows": (ctypes.c_int, [c_float_p] + [ctypes.c_int] * 5 + [c_float_p, c_float_p, ctypes.c_void_p]),
    "apg_poly_reference": (ctypes.c_int, [c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                          c_float_p, ctypes.c_void_p]),
    "apg_polynomial_points": (ctypes.c_int, [c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_int, ctypes.c_double,
                                             ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_float_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "apg_reference_table": (ctypes.c_int, [c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                           ctypes.c_float, ctypes.c_int, c_float_p, ctypes.c_void_p]),
}

_lib = None


class ApgError(RuntimeError):
    pass


def lib():
    """The loaded shared library (raises if it has not been built: `python -m apg_trajectory_tracking_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ApgError(f"{LIB_PATH} is missing: build the CUDA extension with "
                           "`python -m apg_trajectory_tracking_b200.build` (there is no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code):
    if code != 0:
        msg = lib().apg_error_string(code)
        raise ApgError(f"apg_b200 call failed ({code}): {msg.decode() if msg else '?'}")
