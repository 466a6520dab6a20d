"""Multi-GPU plumbing of the genetic algorithm: one process per GPU, `torch.distributed` (NCCL over
NVLink on the B200 box, gloo in CPU tests). The physics needs no communication -- environments
are independent -- so the only exchange is once per generation: every rank's slice of scores and
genomes is gathered into every rank's copy of the whole generation, because the reference draws
parents from the WHOLE generation (src/main.rs:134-138)."""
from typing import Optional, Tuple


def slice_bounds(n_agents: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the population owned by `rank`: contiguous, sizes differ by at most one.
    Must agree with prl_ga_create (csrc/prl_ga.cu)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return n_agents * rank // world_size, n_agents * (rank + 1) // world_size


def gather_generation(local_scores, local_genomes, global_scores, global_genomes, group=None) -> None:
    """All-gather of one generation. local_*: this rank's slice (n_local [x genome_len]);
    global_*: the whole generation (n_agents [x genome_len]), filled in place on every rank.
    Equal slices use one all_gather_into_tensor per buffer (NCCL: ncclAllGather); ragged slices
    fall back to one broadcast per rank."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = global_scores.shape[0]
    lo, hi = slice_bounds(n, rank, world)
    if local_scores.shape[0] != hi - lo or local_genomes.shape[0] != hi - lo:
        raise ValueError("local slice does not match slice_bounds()")
    if n % world == 0 and dist.get_backend(group) != "gloo":
        dist.all_gather_into_tensor(global_scores, local_scores, group=group)
        dist.all_gather_into_tensor(global_genomes, local_genomes, group=group)
        return
    if n % world == 0:
        # gloo has no all_gather_into_tensor for every dtype: gather into views of the output
        dist.all_gather(list(global_scores.chunk(world)), local_scores, group=group)
        dist.all_gather(list(global_genomes.chunk(world)), local_genomes, group=group)
        return
    global_scores[lo:hi].copy_(local_scores)
    global_genomes[lo:hi].copy_(local_genomes)
    for r in range(world):
        rlo, rhi = slice_bounds(n, r, world)
        src = dist.get_global_rank(group, r) if group is not None else r
        dist.broadcast(global_scores[rlo:rhi], src=src, group=group)
        dist.broadcast(global_genomes[rlo:rhi], src=src, group=group)
