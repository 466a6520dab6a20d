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
    """A GA population behind one `prl_ga` handle: the whole population on one GPU (default), one
    rank's slice of a multi-process run (`rank`, `n_ranks`: the host all-gathers once per
    generation, see `gather`), or -- `devices=[...]` -- one slice per listed GPU inside this
    process, exchanged by the library over peer memory (prl_ga_create_multi)."""

    def __init__(self, world: World, number_of_steps=1000, number_of_agents=1000, repeat_move=20, mutation_rate=0.1,
                 keep_best=False, seed=0, device=0, stream=None, rank=0, n_ranks=1, devices=None):
        cfg = _lib.PrlGaConfig(int(number_of_steps), int(number_of_agents), int(repeat_move), float(mutation_rate),
                               1 if keep_best else 0, int(seed) & 0xFFFFFFFFFFFFFFFF)
        h = C.c_void_p()
        if devices is not None:
            if n_ranks != 1:
                raise ValueError("devices=[...] (one process, several GPUs) excludes rank / n_ranks (one process per GPU)")
            devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            _lib.check(_lib.lib().prl_ga_create_multi(world._h, C.byref(cfg), devs, len(devices), C.byref(h)))
            device = int(devices[0])
        else:
            _lib.check(_lib.lib().prl_ga_create(world._h, C.byref(cfg), int(device), C.c_void_p(stream or 0), int(rank), int(n_ranks),
                                               C.byref(h)))
        self._h = h
        self.world = world
        self.device = int(device)
        self.stream = int(stream or 0)
        self.devices = None if devices is None else [int(d) for d in devices]
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        self.number_of_agents = int(number_of_agents)
        self.repeat_move = int(repeat_move)
        self.genome_len = int(number_of_steps) // int(repeat_move)
        if devices is None:
            self.local_offset, hi = slice_bounds(self.number_of_agents, rank, n_ranks)
            self.local_count = hi - self.local_offset
            _, cnt, off, glen = self._buffers()
            assert (cnt, off, glen) == (self.local_count, self.local_offset, self.genome_len)
        else:
            self.local_offset, self.local_count = 0, self.number_of_agents

    def _buffers(self):
        # queried every time: with one rank the genome buffers of parents and children swap
        ptrs = [C.c_void_p() for _ in range(4)]
        cnt, off, glen = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.lib().prl_ga_buffers(self._h, *[C.byref(p) for p in ptrs], C.byref(cnt), C.byref(off), C.byref(glen)))
        return [p.value for p in ptrs], cnt.value, off.value, glen.value

    @property
    def num_shards(self) -> int:
        return _lib.lib().prl_ga_num_shards(self._h)

    def init(self):
        """main.rs:72-84: random genomes for the local slice(s), scored."""
        _lib.check(_lib.lib().prl_ga_init(self._h))

    def tensors(self):
        """(local_scores, local_genomes, global_scores, global_genomes) as torch CUDA tensors
        aliasing the library's buffers (not for devices=[...] populations). Valid until the next
        next_generation()."""
        import torch
        dev = torch.device("cuda", self.device)
        ptrs = self._buffers()[0]
        ls = torch.as_tensor(_DeviceArray(ptrs[0], (self.local_count,), "<f4"), device=dev)
        lg = torch.as_tensor(_DeviceArray(ptrs[1], (self.local_count, self.genome_len), "|u1"), device=dev)
        gs = torch.as_tensor(_DeviceArray(ptrs[2], (self.number_of_agents,), "<f4"), device=dev)
        gg = torch.as_tensor(_DeviceArray(ptrs[3], (self.number_of_agents, self.genome_len), "|u1"), device=dev)
        return ls, lg, gs, gg

    def gather(self, group=None):
        """Per-generation exchange of a multi-process run: all ranks' slices -> every rank's global
        buffers (NCCL all-gather). The collective runs on the handle's own stream, after the
        rollouts that produce the local scores, so nothing is synchronised with the host."""
        if self.n_ranks > 1:
            import torch
            ls, lg, gs, gg = self.tensors()
            dev = torch.device("cuda", self.device)
            if gs.is_cuda:
                ext = torch.cuda.ExternalStream(self.stream, device=dev) if self.stream else torch.cuda.default_stream(dev)
                with torch.cuda.stream(ext):
                    gather_generation(ls, lg, gs, gg, group)
            else:
                gather_generation(ls, lg, gs, gg, group)

    def best(self):
        """main.rs:87-124: (best score, best genome, index); the first minimum wins ties."""
        score, idx = C.c_float(), C.c_int64()
        genome = np.empty(self.genome_len, dtype=np.uint8)
        _lib.check(_lib.lib().prl_ga_best(self._h, C.byref(score), genome.ctypes.data, C.byref(idx)))
        return np.float32(score.value), genome, idx.value

    def next_generation(self):
        """main.rs:126-160 for the local slice(s): elitism, weighted 2-parent choice, crossover,
        mutation, scoring."""
        _lib.check(_lib.lib().prl_ga_next_generation(self._h))

    def step_generation(self):
        """One iteration of the loop at main.rs:86-162 (single-process populations): the best
        (score, genome, index) of the current generation, then the next generation is bred and
        scored. The first call initialises the population."""
        score, idx = C.c_float(), C.c_int64()
        genome = np.empty(self.genome_len, dtype=np.uint8)
        _lib.check(_lib.lib().prl_ga_step_generation(self._h, C.byref(score), genome.ctypes.data, C.byref(idx)))
        return np.float32(score.value), genome, idx.value

    def read_generation(self):
        """(scores, genomes) of the whole current generation on the host (single-process populations)."""
        scores = np.empty(self.number_of_agents, dtype=np.float32)
        genomes = np.empty((self.number_of_agents, self.genome_len), dtype=np.uint8)
        _lib.check(_lib.lib().prl_ga_read_generation(self._h, scores.ctypes.data, genomes.ctypes.data))
        return scores, genomes

    @property
    def generation(self) -> int:
        return _lib.lib().prl_ga_generation(self._h)

    @property
    def launch_count(self) -> int:
        return _lib.lib().prl_ga_launch_count(self._h)

    def env_steps(self) -> int:
        return _lib.lib().prl_ga_env_steps(self._h)

    def local_scores(self) -> np.ndarray:
        if self.devices is not None:
            return self.read_generation()[0]
        return self.tensors()[0].cpu().numpy()

    def local_genomes(self) -> np.ndarray:
        if self.devices is not None:
            return self.read_generation()[1]
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

    def train(self, world: World, sender: Callable, device: Optional[int] = None, group=None, max_generations: Optional[int] = None,
              devices: Optional[List[int]] = None):
        """Algorithm::train (src/algorithm.rs:27, main.rs:41-163). `sender((best_score, agent))`
        is called once per generation on rank 0 and must return False (or raise) when the
        receiver is gone -- the reference returns from train() when `sender.send` fails.
        Several GPUs: either `devices=[0, 1, ...]` (this process drives them all, exchange over peer
        memory inside the library), or one process per GPU under torch.distributed (`group`):
        every rank runs train() on its own GPU (default: LOCAL_RANK) and slice."""
        rank, n_ranks = 0, 1
        if devices is None and (group is not None or _dist_initialized()):
            import os
            import torch.distributed as dist
            rank, n_ranks = dist.get_rank(group), dist.get_world_size(group)
            if device is None:
                device = int(os.environ.get("LOCAL_RANK", "0"))
        pop = GaPopulation(world, self.number_of_steps, self.number_of_agents, self.repeat_move, self.mutation_rate,
                           self.keep_best, self.seed, device=device or 0, rank=rank, n_ranks=n_ranks, devices=devices)
        device = pop.device
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
