"""ctypes binding of libgensmc.so (include/gen_b200.h). No CPU fallback: if the CUDA library is
missing or no GPU is present, calls fail loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgensmc.so")

MODEL_HMM, MODEL_LGSSM, MODEL_SV, MODEL_BEARINGS, MODEL_REGRESSION, MODEL_NORMAL_NORMAL = 1, 2, 3, 4, 5, 6
MODEL_OUTLIER_REGRESSION, MODEL_UNIFORM_NORMAL = 7, 8
IS_FAMILIES = (MODEL_REGRESSION, MODEL_NORMAL_NORMAL, MODEL_OUTLIER_REGRESSION, MODEL_UNIFORM_NORMAL)
PROPOSAL_DEFAULT, PROPOSAL_CUSTOM = 0, 1
RESAMPLE_MULTINOMIAL, RESAMPLE_RESIDUAL = 0, 1
F64, F32 = 0, 1
E_BADARG, E_CUDA, E_NCCL, E_DEGENERATE, E_UNSUPPORTED, E_NOMEM, E_PEER = -1, -2, -3, -4, -5, -6, -7


class GsmcError(RuntimeError):
    """Raised where the reference would call error(...)."""

    def __init__(self, code, message):
        super().__init__("libgensmc error %d: %s" % (code, message))
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("model_id", C.c_int32), ("dtype", C.c_int32),
                ("resample_scheme", C.c_int32), ("num_particles", C.c_uint64), ("seed", C.c_uint64),
                ("device", C.c_int32), ("keep_history", C.c_int32), ("history_capacity", C.c_int64),
                ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("last_ess", C.c_double), ("last_log_total", C.c_double), ("log_ml_est", C.c_double),
                ("num_steps", C.c_int64), ("num_resamples", C.c_int64), ("kernel_launches", C.c_int64),
                ("ms_propagate", C.c_double), ("ms_propagate_gather", C.c_double), ("ms_finalize", C.c_double),
                ("ms_scan", C.c_double), ("ms_spacings", C.c_double), ("ms_search", C.c_double), ("ms_other", C.c_double),
                ("n_propagate", C.c_int64), ("n_propagate_gather", C.c_int64), ("n_finalize", C.c_int64),
                ("n_scan", C.c_int64), ("n_spacings", C.c_int64), ("n_search", C.c_int64), ("n_other", C.c_int64),
                ("graph_replays", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_H = C.c_void_p

# every symbol include/gen_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gsmc_version": (C.c_char_p, []),
    "gsmc_last_error": (C.c_char_p, [_H]),
    "gsmc_create": (C.c_int, [C.POINTER(Config), _dp, C.c_size_t, C.POINTER(_H)]),
    "gsmc_destroy": (None, [_H]),
    "gsmc_reset": (C.c_int, [_H]),
    "gsmc_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "gsmc_comm_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(_H)]),
    "gsmc_comm_destroy": (None, [_H]),
    "gsmc_comm_attach": (C.c_int, [_H, _H]),
    "gsmc_group_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_H)]),
    "gsmc_group_destroy": (None, [_H]),
    "gsmc_group_attach": (C.c_int, [_H, C.c_int, _H]),
    "gsmc_group_init": (C.c_int, [_H, _dp, C.c_size_t, C.c_int, _dp, C.c_size_t]),
    "gsmc_group_step": (C.c_int, [_H, _dp, C.c_size_t, C.c_int, _dp, C.c_size_t]),
    "gsmc_group_maybe_resample": (C.c_int, [_H, C.c_double, C.POINTER(C.c_int), _dp]),
    "gsmc_group_sample_unweighted": (C.c_int, [_H, C.c_uint64, _ip]),
    "gsmc_set_replay": (C.c_int, [_H, _dp, C.c_size_t, _dp, C.c_size_t]),
    "gsmc_init": (C.c_int, [_H, _dp, C.c_size_t, C.c_int, _dp, C.c_size_t]),
    "gsmc_step": (C.c_int, [_H, _dp, C.c_size_t, C.c_int, _dp, C.c_size_t]),
    "gsmc_maybe_resample": (C.c_int, [_H, C.c_double, C.POINTER(C.c_int), _dp]),
    "gsmc_log_ml_estimate": (C.c_int, [_H, _dp]),
    "gsmc_get_log_weights": (C.c_int, [_H, _dp, C.c_size_t]),
    "gsmc_get_log_weights_device": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "gsmc_get_state": (C.c_int, [_H, C.c_int64, _dp, C.c_size_t]),
    "gsmc_get_observation": (C.c_int, [_H, C.c_int64, _dp, C.c_size_t]),
    "gsmc_get_trajectories": (C.c_int, [_H, _ip, C.c_size_t, _dp, C.c_size_t]),
    "gsmc_get_ancestors": (C.c_int, [_H, _ip, C.c_size_t]),
    "gsmc_sample_unweighted": (C.c_int, [_H, C.c_uint64, _ip]),
    "gsmc_importance_sampling": (C.c_int, [C.POINTER(Config), _dp, C.c_size_t, _dp, C.c_size_t, C.c_int, _dp,
                                           C.c_size_t, _dp, C.POINTER(_H)]),
    "gsmc_run_steps": (C.c_int, [_H, _dp, C.c_size_t, C.c_size_t, C.c_int, _dp, C.c_size_t, C.c_double]),
    "gsmc_save": (C.c_int, [_H, C.c_char_p]),
    "gsmc_restore": (C.c_int, [_H, C.c_char_p]),
    "gsmc_register_model_plugin": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "gsmc_trim": (C.c_int, []),
    "gsmc_local_count": (C.c_int, [_H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gsmc_state_dim": (C.c_int, [_H, C.POINTER(C.c_int)]),
    "gsmc_synchronize": (C.c_int, [_H]),
    "gsmc_get_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "gsmc_set_profiling": (C.c_int, [_H, C.c_int]),
    "gsmc_timer_start": (C.c_int, [_H]),
    "gsmc_timer_stop": (C.c_int, [_H, _dp]),
}

_lib = None


def load():
    """Load libgensmc.so; raises if it has not been built (python -m gen_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsmcError(E_CUDA, "%s is missing: build it with `python -m gen_b200.build` (needs nvcc); "
                                "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != 0:
        raise GsmcError(rc, load().gsmc_last_error(handle).decode())


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)
