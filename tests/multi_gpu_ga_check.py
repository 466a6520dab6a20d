#!/usr/bin/env python
"""Run under torchrun with N >= 2 ranks (one GPU each): the GA's per-generation NCCL all-gather on
real GPUs. Every rank breeds its slice; rank 0 also runs the same GA unsharded on its own GPU and
the two must agree bit for bit (population sharding must not change the result)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402


def main():
    rank, local_rank, world_size = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = prl.World.load(os.path.join(ROOT, "worlds", "push.json"))
    kw = dict(number_of_steps=400, number_of_agents=4096 + 3, repeat_move=20, mutation_rate=0.1, keep_best=True, seed=99)
    pop = prl.GaPopulation(world, device=local_rank, rank=rank, n_ranks=world_size, **kw)
    single = prl.GaPopulation(world, device=local_rank, **kw) if rank == 0 else None
    pop.init()
    if single:
        single.init()
    ok = True
    for gen in range(3):
        pop.gather()
        score, genome, idx = pop.best()
        _, _, gs, gg = pop.tensors()
        # every rank holds the same generation
        chk = torch.stack([gs.double().sum(), gg.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo, hi))
        if single:
            s1, g1, i1 = single.best()
            ok &= (np.float32(s1).view(np.uint32) == np.float32(score).view(np.uint32)) and i1 == idx and np.array_equal(g1, genome)
            ok &= np.array_equal(gg.cpu().numpy(), single.local_genomes())
            ok &= np.array_equal(gs.cpu().numpy().view(np.uint32), single.local_scores().view(np.uint32))
            single.next_generation()
        pop.next_generation()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_GA_OK" if flag.item() == 1 else "MULTI_GPU_GA_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
