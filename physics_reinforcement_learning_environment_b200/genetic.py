"""The reference's bundled genetic algorithm (src/main.rs) on the batched environment.

`GeneticAlgorithm` keeps the reference's fields and defaults (main.rs:18-38) and `train()` keeps
its contract (main.rs:41-163, README.md:84-94): score every agent by its rollout's minimum
distance to the goals, send `(best_score, GeneticAgent)` once per generation, stop when the
receiver goes away. All arithmetic -- rollouts, min/max, weighted parent choice, crossover,
mutation -- runs in the CUDA kernels behind prl_ga_* (csrc/prl_ga.cu); this module only sequences
the calls and, with several GPUs, the per-generation all-gather.
"""
import ctypes as C
import dataclasses
from typing import Callable, List, Optional

import numpy as np

from . import _lib
from .parallel import gather_generation, slice_bounds
from .world import Move, World


@dataclasses.dataclass
class GeneticAgent:
    """main.rs:223-284: a fixed move sequence, each move repeated `repeat_move` times."""
    moves: List[Move]
    curr: int = 0
    repeat_move: int = 20

    def get_move(self, _environment=None) -> Move:
        if self.curr // self.repeat_move < len(self.moves):
            mv = self.moves[self.curr // self.repeat_move]
            self.curr += 1
            return mv
        return Move()

    def genome(self) -> np.ndarray:
        return np.array([m.bits() for m in self.moves], dtype=np.uint8)


class _DeviceArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view a prl_ga device buffer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class GaPopulation:
    """One rank's slice of a GA population: thin owner of a `prl_ga` handle."""

    def __init__(self, world: World, number_of_steps=1000, number_of_agents=1000, repeat_move=20, mutation_rate=0.1,
                 keep_best=False, seed=0, device=0, stream=None, rank=0, n_ranks=1):
        cfg = _lib.PrlGaConfig(int(number_of_steps), int(number_of_agents), int(repeat_move), float(mutation_rate),
                               1 if keep_best else 0, int(seed) & 0xFFFFFFFFFFFFFFFF)
        h = C.c_void_p()
        _lib.check(_lib.lib().prl_ga_create(world._h, C.byref(cfg), int(device), C.c_void_p(stream or 0), int(rank), int(n_ranks),
                                           C.byref(h)))
        self._h = h
        self.world = world
        self.device = int(device)
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        self.number_of_agents = int(number_of_agents)
        self.repeat_move = int(repeat_move)
        ptrs = [C.c_void_p() for _ in range(4)]
        cnt, off, glen = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.lib().prl_ga_buffers(h, *[C.byref(p) for p in ptrs], C.byref(cnt), C.byref(off), C.byref(glen)))
        self._ptrs = [p.value for p in ptrs]
        self.local_count, self.local_offset, self.genome_len = cnt.value, off.value, glen.value
        assert (self.local_offset, self.local_offset + self.local_count) == slice_bounds(self.number_of_agents, rank, n_ranks)

    def init(self):
        """main.rs:72-84: random genomes for the local slice, scored."""
        _lib.check(_lib.lib().prl_ga_init(self._h))

    def tensors(self):
        """(local_scores, local_genomes, global_scores, global_genomes) as torch CUDA tensors
        aliasing the library's buffers."""
        import torch
        dev = torch.device("cuda", self.device)
        ls = torch.as_tensor(_DeviceArray(self._ptrs[0], (self.local_count,), "<f4"), device=dev)
        lg = torch.as_tensor(_DeviceArray(self._ptrs[1], (self.local_count, self.genome_len), "|u1"), device=dev)
        gs = torch.as_tensor(_DeviceArray(self._ptrs[2], (self.number_of_agents,), "<f4"), device=dev)
        gg = torch.as_tensor(_DeviceArray(self._ptrs[3], (self.number_of_agents, self.genome_len), "|u1"), device=dev)
        return ls, lg, gs, gg

    def gather(self, group=None):
        """Per-generation exchange (multi-GPU only): all ranks' slices -> every rank's global buffers."""
        if self.n_ranks > 1:
            import torch
            ls, lg, gs, gg = self.tensors()
            torch.cuda.current_stream(self.device).synchronize()
            gather_generation(ls, lg, gs, gg, group)
            torch.cuda.current_stream(self.device).synchronize()

    def best(self):
        """main.rs:87-124: (best score, best genome, index); the first minimum wins ties."""
        score, idx = C.c_float(), C.c_int64()
        genome = np.empty(self.genome_len, dtype=np.uint8)
        _lib.check(_lib.lib().prl_ga_best(self._h, C.byref(score), genome.ctypes.data, C.byref(idx)))
        return np.float32(score.value), genome, idx.value

    def next_generation(self):
        """main.rs:126-160 for the local slice: elitism, weighted 2-parent choice, crossover,
        mutation, scoring."""
        _lib.check(_lib.lib().prl_ga_next_generation(self._h))

    @property
    def generation(self) -> int:
        return _lib.lib().prl_ga_generation(self._h)

    @property
    def launch_count(self) -> int:
        return _lib.lib().prl_ga_launch_count(self._h)

    def env_steps(self) -> int:
        return _lib.lib().prl_ga_env_steps(self._h)

    def local_scores(self) -> np.ndarray:
        return self.tensors()[0].cpu().numpy()

    def local_genomes(self) -> np.ndarray:
        return self.tensors()[1].cpu().numpy()

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().prl_ga_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclasses.dataclass
class GeneticAlgorithm:
    """main.rs:18-38 (fields and defaults). `seed` replaces the reference's thread_rng (main.rs:42)."""
    number_of_steps: int = 1000
    number_of_agents: int = 1000
    repeat_move: int = 20
    mutation_rate: float = 0.1
    keep_best: bool = False
    seed: int = 0xB200

    def train(self, world: World, sender: Callable, device: int = 0, group=None, max_generations: Optional[int] = None):
        """Algorithm::train (src/algorithm.rs:27, main.rs:41-163). `sender((best_score, agent))`
        is called once per generation on rank 0 and must return False (or raise) when the
        receiver is gone -- the reference returns from train() when `sender.send` fails.
        With `group` (torch.distributed), every rank runs train() on its own GPU and slice."""
        rank, n_ranks = 0, 1
        if group is not None or _dist_initialized():
            import torch.distributed as dist
            rank, n_ranks = dist.get_rank(group), dist.get_world_size(group)
        pop = GaPopulation(world, self.number_of_steps, self.number_of_agents, self.repeat_move, self.mutation_rate,
                           self.keep_best, self.seed, device=device, rank=rank, n_ranks=n_ranks)
        try:
            pop.init()
            while True:
                pop.gather(group)
                score, genome, _ = pop.best()
                go_on = True
                if rank == 0:
                    agent = GeneticAgent([Move.from_bits(int(b)) for b in genome], 0, self.repeat_move)
                    try:
                        go_on = sender((score, agent)) is not False
                    except Exception:
                        go_on = False
                if n_ranks > 1:
                    go_on = _broadcast_flag(go_on, device, group)
                if not go_on or (max_generations is not None and pop.generation + 1 >= max_generations):
                    return pop.generation + 1
                pop.next_generation()
        finally:
            pop.close()


def _dist_initialized() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def _broadcast_flag(flag: bool, device: int, group) -> bool:
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=torch.device("cuda", device))
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(t, src=src, group=group)
    return bool(t.item())
