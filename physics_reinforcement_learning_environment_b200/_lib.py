"""ctypes binding of libprl.so, the C ABI declared in include/prl.h.

The library is the product: CUDA kernels for sm_100a behind plain-C entry points. There is no
CPU fallback -- if the shared object is missing this module raises, and every compute call
raises PrlError(PRL_ERR_CUDA) when no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRL_LIBRARY: A/B measurements of kernel variants built beside the product library
LIB_PATH = os.environ.get("PRL_LIBRARY") or os.path.join(_HERE, "libprl.so")

PRL_OK = 0
PRL_ERR_INVALID = -1
PRL_ERR_PARSE = -2
PRL_ERR_CAPACITY = -3
PRL_ERR_NO_GOAL = -4
PRL_ERR_CUDA = -5
PRL_ERR_OVERFLOW = -6


class PrlObject(C.Structure):
    """`prl_object` (include/prl.h) = one ObjectAndTransform (reference src/common.rs:26-31)."""
    _fields_ = [("position", C.c_float * 3), ("scale", C.c_float * 2),
                ("rotation", C.c_float), ("kind", C.c_int32)]


class PrlGaConfig(C.Structure):
    """`prl_ga_config` (include/prl.h) = GeneticAlgorithm's fields (reference src/main.rs:18-38)."""
    _fields_ = [("number_of_steps", C.c_int32), ("number_of_agents", C.c_int64),
                ("repeat_move", C.c_int32), ("mutation_rate", C.c_float),
                ("keep_best", C.c_int32), ("seed", C.c_uint64)]


class PrlError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"prl status {status}: {message}")
        self.status = status


_lib = None

_SIGNATURES = {
    "prl_last_error": (C.c_char_p, []),
    "prl_abi_version": (C.c_int32, []),
    "prl_device_count": (C.c_int32, []),
    "prl_debug_bounds_violations": (C.c_int32, []),
    "prl_world_create": (C.c_int, [C.POINTER(C.c_float), C.POINTER(PrlObject), C.c_int32, C.POINTER(C.c_void_p)]),
    "prl_world_from_json": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "prl_world_to_json": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "prl_world_num_objects": (C.c_int32, [C.c_void_p]),
    "prl_world_get_object": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(PrlObject)]),
    "prl_world_player_position": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "prl_world_counts": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
    "prl_world_object_body": (C.c_int32, [C.c_void_p, C.c_int32]),
    "prl_world_destroy": (None, [C.c_void_p]),
    "prl_batch_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "prl_batch_destroy": (None, [C.c_void_p]),
    "prl_batch_num_envs": (C.c_int64, [C.c_void_p]),
    "prl_batch_num_bodies": (C.c_int32, [C.c_void_p]),
    "prl_batch_state_bytes": (C.c_int64, [C.c_void_p]),
    "prl_batch_reset": (C.c_int, [C.c_void_p]),
    "prl_batch_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prl_batch_step_uniform": (C.c_int, [C.c_void_p, C.c_uint32, C.c_int32]),
    "prl_batch_distance_to_goals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prl_batch_won": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prl_batch_get_state": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "prl_batch_rollout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]),
    "prl_batch_rollout_info": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prl_batch_trajectory": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "prl_batch_observation_dim": (C.c_int32, [C.c_void_p]),
    "prl_batch_get_observation": (C.c_int, [C.c_void_p, C.c_void_p]),
    "prl_policy_num_params": (C.c_int32, [C.c_void_p, C.c_int32]),
    "prl_batch_rollout_policy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "prl_batch_sync": (C.c_int, [C.c_void_p]),
    "prl_batch_launch_count": (C.c_int64, [C.c_void_p]),
    "prl_ga_create": (C.c_int, [C.c_void_p, C.POINTER(PrlGaConfig), C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                C.POINTER(C.c_void_p)]),
    "prl_ga_create_multi": (C.c_int, [C.c_void_p, C.POINTER(PrlGaConfig), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "prl_ga_destroy": (None, [C.c_void_p]),
    "prl_ga_num_shards": (C.c_int32, [C.c_void_p]),
    "prl_ga_init": (C.c_int, [C.c_void_p]),
    "prl_ga_buffers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_int32)]),
    "prl_ga_best": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int64)]),
    "prl_ga_next_generation": (C.c_int, [C.c_void_p]),
    "prl_ga_step_generation": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int64)]),
    "prl_ga_read_generation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "prl_ga_generation": (C.c_int64, [C.c_void_p]),
    "prl_ga_launch_count": (C.c_int64, [C.c_void_p]),
    "prl_ga_env_steps": (C.c_int64, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load libprl.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != PRL_OK:
        raise PrlError(status, lib().prl_last_error().decode("utf-8", "replace"))
