"""ctypes binding of the CPU ORACLE (test infrastructure, not a product path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package never does.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcObject(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("scale", C.c_float * 2),
                ("rotation", C.c_float), ("kind", C.c_int32)]


# Sensitivity variants (prl_oracle.h ORC_VARIANT_*): flip one documented deviation of the oracle
VARIANT_REVERSE_ORDER, VARIANT_CREATION_ORDER, VARIANT_GJK, VARIANT_WAKE_ON_ZERO_IMPULSE = 1, 2, 4, 8
VARIANTS = {"D2 descending pair order": 1, "D2 contact-graph edge order": 2, "D1 GJK closest points": 4,
            "D4 zero impulse wakes": 8, "D1+D2+D4 together": 2 | 4 | 8}


def set_default_variant(flags):
    """Applies to environments created afterwards (rollout_batch included). Tests only."""
    lib().orc_set_default_variant(int(flags))


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libprl_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_env_create.restype = C.c_void_p
        L.orc_env_create.argtypes = [C.POINTER(C.c_float), C.POINTER(OrcObject), C.c_int32]
        L.orc_env_destroy.argtypes = [C.c_void_p]
        L.orc_set_default_variant.argtypes = [C.c_int32]
        L.orc_env_set_variant.argtypes = [C.c_void_p, C.c_int32]
        L.orc_env_reset.argtypes = [C.c_void_p]
        L.orc_env_step.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_env_distance_to_goals.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.orc_env_won.argtypes = [C.c_void_p]
        L.orc_env_num_bodies.argtypes = [C.c_void_p]
        L.orc_env_body_state.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        L.orc_env_set_body_state.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        L.orc_env_num_pairs.argtypes = [C.c_void_p]
        L.orc_env_manifold.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        L.orc_env_pair.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.orc_rollout.restype = C.c_float
        L.orc_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_rollout_batch.restype = C.c_int64
        L.orc_rollout_batch.argtypes = [C.POINTER(C.c_float), C.POINTER(OrcObject), C.c_int32, C.c_void_p,
                                        C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p]
        L.orc_env_observation_dim.argtypes = [C.c_void_p]
        L.orc_env_observation.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_policy_move.restype = C.c_uint32
        L.orc_policy_move.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_rollout_policy_batch.restype = C.c_int64
        L.orc_rollout_policy_batch.argtypes = [C.POINTER(C.c_float), C.POINTER(OrcObject), C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def world_to_objects(world):
    """world: dict in the serde schema of common.rs:19-49 (already json-parsed)."""
    objs = (OrcObject * max(1, len(world["objects"])))()
    for i, o in enumerate(world["objects"]):
        ob = o["object"]
        kind = 2 if ob == "Goal" else (0 if ob["Block"]["fixed"] else 1)
        objs[i].position[:] = [np.float32(v) for v in o["position"]]
        objs[i].scale[:] = [np.float32(v) for v in o["scale"]]
        objs[i].rotation = np.float32(o["rotation"])
        objs[i].kind = kind
    pp = (C.c_float * 2)(*[np.float32(v) for v in world["player_position"]])
    return pp, objs, len(world["objects"])


class OracleEnv:
    """Single CPU environment with the reference's Environment API."""

    def __init__(self, world, variant=None):
        if isinstance(world, str):
            world = json.loads(world)
        self._pp, self._objs, self._n = world_to_objects(world)
        self._h = lib().orc_env_create(self._pp, self._objs, self._n)
        if variant is not None:
            lib().orc_env_set_variant(self._h, int(variant))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_env_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_env_reset(self._h)

    def step(self, move):
        lib().orc_env_step(self._h, int(move))

    def distance_to_goals(self):
        out = C.c_float()
        if lib().orc_env_distance_to_goals(self._h, C.byref(out)) != 0:
            return None
        return np.float32(out.value)

    def won(self):
        return bool(lib().orc_env_won(self._h))

    @property
    def num_bodies(self):
        return lib().orc_env_num_bodies(self._h)

    def body_state(self, b):
        out = (C.c_float * 9)()
        lib().orc_env_body_state(self._h, b, out)
        return np.array(out[:], dtype=np.float32)

    def set_body_state(self, b, st):
        arr = (C.c_float * 9)(*[np.float32(v) for v in st])
        lib().orc_env_set_body_state(self._h, b, arr)

    def state(self):
        return np.stack([self.body_state(b) for b in range(self.num_bodies)])

    def observation(self):
        """What Agent::get_move can read between two steps (layout of prl_batch_get_observation)."""
        out = np.zeros(lib().orc_env_observation_dim(self._h), dtype=np.float32)
        lib().orc_env_observation(self._h, out.ctypes.data)
        return out

    @property
    def num_pairs(self):
        return lib().orc_env_num_pairs(self._h)

    def manifold(self, p):
        out = (C.c_float * 16)()
        lib().orc_env_manifold(self._h, p, out)
        return np.array(out[:], dtype=np.float32)

    def pair(self, p):
        out = (C.c_int32 * 2)()
        lib().orc_env_pair(self._h, p, out)
        return out[0], out[1]

    def rollout(self, genome, repeat_move, number_of_steps, early_exit=True):
        g = np.ascontiguousarray(genome, dtype=np.uint8)
        win = C.c_int32()
        run = C.c_int32()
        s = lib().orc_rollout(self._h, g.ctypes.data, len(g), repeat_move, number_of_steps,
                              1 if early_exit else 0, C.byref(win), C.byref(run))
        return np.float32(s), win.value, run.value


def rollout_batch(world, genomes, repeat_move, number_of_steps, early_exit=True, n_threads=1):
    if isinstance(world, str):
        world = json.loads(world)
    pp, objs, n = world_to_objects(world)
    g = np.ascontiguousarray(genomes, dtype=np.uint8)
    n_envs, glen = g.shape
    score = np.empty(n_envs, dtype=np.float32)
    win = np.empty(n_envs, dtype=np.int32)
    steps = lib().orc_rollout_batch(pp, objs, n, g.ctypes.data, n_envs, glen, repeat_move, number_of_steps,
                                    1 if early_exit else 0, n_threads, score.ctypes.data, win.ctypes.data)
    return score, win, int(steps)


def rollout_policy_batch(world, params, hidden, number_of_steps, early_exit=True, n_threads=1):
    """params: (P, n_envs) float32, parameter-major (include/prl.h prl_batch_rollout_policy)."""
    if isinstance(world, str):
        world = json.loads(world)
    pp, objs, n = world_to_objects(world)
    th = np.ascontiguousarray(params, dtype=np.float32)
    n_envs = th.shape[1]
    score = np.empty(n_envs, dtype=np.float32)
    win = np.empty(n_envs, dtype=np.int32)
    steps = lib().orc_rollout_policy_batch(pp, objs, n, th.ctypes.data, n_envs, int(hidden), number_of_steps,
                                           1 if early_exit else 0, n_threads, score.ctypes.data, win.ctypes.data)
    return score, win, int(steps)
