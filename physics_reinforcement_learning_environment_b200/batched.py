"""`BatchedEnvironment`: n independent `Environment`s of one `World` on one B200.

Mirrors the reference's `Environment` API (src/common.rs:162-340) with a leading batch axis:
`from_world`, `step`, `distance_to_goals`, `won`, plus the fused GA rollout of src/main.rs:44-70.
All arithmetic happens in the CUDA kernels of libprl.so; numpy arrays and torch tensors are only
buffers handed across the C ABI.
"""
import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .world import Move, World


def _ptr(buf, dtype, count, what):
    """Raw pointer of a numpy array (host) or a torch CUDA/CPU tensor, after checking dtype/size."""
    if isinstance(buf, np.ndarray):
        if buf.dtype != dtype or not buf.flags["C_CONTIGUOUS"] or buf.size < count:
            raise ValueError(f"{what}: need a C-contiguous {np.dtype(dtype).name} array of >= {count} elements")
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):  # torch tensor
        import torch
        want = {np.uint8: torch.uint8, np.float32: torch.float32, np.int32: torch.int32}[dtype]
        if buf.dtype != want or not buf.is_contiguous() or buf.numel() < count:
            raise ValueError(f"{what}: need a contiguous {want} tensor of >= {count} elements")
        return buf.data_ptr()
    raise TypeError(f"{what}: expected a numpy array or a torch tensor")


class BatchedEnvironment:
    def __init__(self, world: World, n_envs: int, device: int = 0, stream: Optional[int] = None):
        """n_envs x `Environment::from_world(world)` (common.rs:162-172). `stream`: a raw
        cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = default stream."""
        self.world = world
        self.n_envs = int(n_envs)
        self.device = int(device)
        h = C.c_void_p()
        _lib.check(_lib.lib().prl_batch_create(world._h, self.n_envs, self.device, C.c_void_p(stream or 0), C.byref(h)))
        self._h = h
        self.n_bodies = _lib.lib().prl_batch_num_bodies(h)

    @staticmethod
    def from_world(world: World, n_envs: int, device: int = 0, stream: Optional[int] = None):
        env = BatchedEnvironment(world, n_envs, device, stream)
        return env, world.rigid_body_handles()

    def reset(self):
        _lib.check(_lib.lib().prl_batch_reset(self._h))

    def step(self, moves):
        """Environment::step (common.rs:215-340) for every environment. `moves`: n_envs uint8
        (bit0 left, bit1 right, bit2 up), a list of `Move`, or a single `Move`/int for all."""
        L = _lib.lib()
        if isinstance(moves, Move):
            _lib.check(L.prl_batch_step_uniform(self._h, moves.bits(), 1))
            return
        if isinstance(moves, int):
            _lib.check(L.prl_batch_step_uniform(self._h, moves, 1))
            return
        if isinstance(moves, (list, tuple)):
            moves = np.array([m.bits() if isinstance(m, Move) else int(m) for m in moves], dtype=np.uint8)
        _lib.check(L.prl_batch_step(self._h, _ptr(moves, np.uint8, self.n_envs, "moves")))

    def step_uniform(self, move, n_steps: int):
        bits = move.bits() if isinstance(move, Move) else int(move)
        _lib.check(_lib.lib().prl_batch_step_uniform(self._h, bits, int(n_steps)))

    def distance_to_goals(self, out=None):
        """common.rs:175-195 for every environment, editor pixels. Raises PrlError(PRL_ERR_NO_GOAL)
        for a goal-less world (the reference returns None)."""
        if out is None:
            out = np.empty(self.n_envs, dtype=np.float32)
        _lib.check(_lib.lib().prl_batch_distance_to_goals(self._h, _ptr(out, np.float32, self.n_envs, "out")))
        return out

    def won(self, out=None):
        """common.rs:197-199 for every environment (uint8 0/1)."""
        if out is None:
            out = np.empty(self.n_envs, dtype=np.uint8)
        _lib.check(_lib.lib().prl_batch_won(self._h, _ptr(out, np.uint8, self.n_envs, "out")))
        return out

    def state(self, env0: int = 0, n: Optional[int] = None, out=None):
        """Bodies of environments [env0, env0+n) as (n, n_bodies, 9) float32:
        x, y, cos, sin, vx, vy, angvel, sleep timer, sleeping -- the batched form of
        rigid_body_set()/player_handle() (common.rs:201-211), i.e. what `Agent::get_move(&self,
        &Environment)` (algorithm.rs:9) can observe. Body 0 is the player. `out` may be a torch
        CUDA tensor: the observation then stays on the device for a policy that runs there."""
        n = self.n_envs - env0 if n is None else n
        if out is None:
            out = np.empty((n, self.n_bodies, 9), dtype=np.float32)
        _lib.check(_lib.lib().prl_batch_get_state(self._h, env0, n, _ptr(out, np.float32, n * self.n_bodies * 9, "out")))
        return out

    def rollout(self, genomes, repeat_move: int, number_of_steps: int, early_exit: bool = True,
                out_score=None, out_step_of_win=None) -> Tuple[np.ndarray, np.ndarray]:
        """main.rs:44-70 for every agent: genomes is (n_envs, genome_len) uint8 Moves."""
        if isinstance(genomes, np.ndarray):
            genome_len = genomes.shape[1]
        else:
            genome_len = int(genomes.shape[1])
        if out_score is None:
            out_score = np.empty(self.n_envs, dtype=np.float32)
        if out_step_of_win is None:
            out_step_of_win = np.empty(self.n_envs, dtype=np.int32)
        _lib.check(_lib.lib().prl_batch_rollout(
            self._h, _ptr(genomes, np.uint8, self.n_envs * genome_len, "genomes"), genome_len, int(repeat_move),
            int(number_of_steps), 1 if early_exit else 0, _ptr(out_score, np.float32, self.n_envs, "out_score"),
            _ptr(out_step_of_win, np.int32, self.n_envs, "out_step_of_win")))
        return out_score, out_step_of_win

    @property
    def observation_dim(self) -> int:
        return _lib.lib().prl_batch_observation_dim(self._h)

    def observation(self, out=None):
        """What `Agent::get_move(&mut self, &Environment)` (algorithm.rs:9) can read between two
        steps, (n_envs, D = 9 n_bodies + 4) float32: per body x, y, cos, sin, vx, vy, angvel, sleep
        timer, sleeping; on_ground, number of floor contacts, touching a pushable block (the scan of
        common.rs:223-244); distance_to_goals in pixels. `out` may be a torch CUDA tensor."""
        D = self.observation_dim
        if out is None:
            out = np.empty((self.n_envs, D), dtype=np.float32)
        _lib.check(_lib.lib().prl_batch_get_observation(self._h, _ptr(out, np.float32, self.n_envs * D, "out")))
        return out

    def policy_num_params(self, hidden: int) -> int:
        return _lib.lib().prl_policy_num_params(self._h, int(hidden))

    def rollout_policy(self, params, hidden: int, number_of_steps: int, early_exit: bool = True,
                       out_score=None, out_step_of_win=None) -> Tuple[np.ndarray, np.ndarray]:
        """The fitness rollout of main.rs:44-70 with a per-agent two-layer perceptron evaluated inside
        the kernel every step in place of the genome (include/prl.h prl_batch_rollout_policy).
        params: (policy_num_params(hidden), n_envs) float32, parameter-major; numpy or torch CUDA."""
        P = self.policy_num_params(hidden)
        if out_score is None:
            out_score = np.empty(self.n_envs, dtype=np.float32)
        if out_step_of_win is None:
            out_step_of_win = np.empty(self.n_envs, dtype=np.int32)
        _lib.check(_lib.lib().prl_batch_rollout_policy(
            self._h, _ptr(params, np.float32, P * self.n_envs, "params"), int(hidden), int(number_of_steps), 1 if early_exit else 0,
            _ptr(out_score, np.float32, self.n_envs, "out_score"), _ptr(out_step_of_win, np.int32, self.n_envs, "out_step_of_win")))
        return out_score, out_step_of_win

    def trajectory(self, genomes, repeat_move: int, number_of_steps: int):
        """Replay for the viewers (train.rs:132-159): every genome (a GeneticAgent, main.rs:276-284)
        is played on a fresh from_world environment for `number_of_steps` steps in one launch.
        Returns (state[T, n, n_bodies, 9], distance_to_goals[T, n], won[T, n]); the batch's own
        environments are not touched."""
        genomes = np.ascontiguousarray(genomes, dtype=np.uint8)
        n, genome_len = genomes.shape
        T = int(number_of_steps)
        state = np.empty((T, n, self.n_bodies, 9), dtype=np.float32)
        dist = np.empty((T, n), dtype=np.float32)
        won = np.empty((T, n), dtype=np.uint8)
        _lib.check(_lib.lib().prl_batch_trajectory(self._h, n, genomes.ctypes.data, genome_len, int(repeat_move), T,
                                                   state.ctypes.data, dist.ctypes.data, won.ctypes.data))
        return state, dist, won

    def rollout_info(self):
        """(running min distance, step of first win or -1, steps since reset) per environment."""
        s = np.empty(self.n_envs, np.float32)
        w = np.empty(self.n_envs, np.int32)
        t = np.empty(self.n_envs, np.int32)
        _lib.check(_lib.lib().prl_batch_rollout_info(self._h, s.ctypes.data, w.ctypes.data, t.ctypes.data))
        return s, w, t

    def sync(self):
        _lib.check(_lib.lib().prl_batch_sync(self._h))

    @property
    def launch_count(self) -> int:
        return _lib.lib().prl_batch_launch_count(self._h)

    @property
    def state_bytes(self) -> int:
        return _lib.lib().prl_batch_state_bytes(self._h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().prl_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
