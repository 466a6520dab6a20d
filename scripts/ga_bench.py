#!/usr/bin/env python
"""Secondary measurements (not the headline bench): (a) the genetic algorithm end to end
(BASELINE config C5 shape, one rank or torchrun), (b) closed-loop stepping, one launch per step."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=1 << 20)
ap.add_argument("--generations", type=int, default=5)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--world", default="push")
args = ap.parse_args()

rank, local_rank, world_size = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
if world_size > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
world = prl.World.load(os.path.join(ROOT, "worlds", f"{args.world}.json"))

pop = prl.GaPopulation(world, number_of_steps=args.steps, number_of_agents=args.agents, repeat_move=20, mutation_rate=0.1,
                       keep_best=False, seed=0xB200, device=local_rank, rank=rank, n_ranks=world_size)
t0 = time.perf_counter()
pop.init()
torch.cuda.synchronize()
steps_total = pop.env_steps()
t_init = time.perf_counter() - t0
times, bests = [], []
for g in range(args.generations):
    t0 = time.perf_counter()
    pop.gather()
    bests.append(float(pop.best()[0]))
    pop.next_generation()
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    steps_total += pop.env_steps()
if rank == 0:
    print(f"GA {args.world}: {args.agents} agents x {args.steps} steps, {world_size} rank(s): init {t_init:.3f} s, "
          f"generation {np.mean(times):.3f} s avg; best per generation {bests}; "
          f"{steps_total * world_size / (t_init + sum(times)):.3e} env-steps/s incl. selection and gather (early exit on)")
    # closed-loop stepping: observation on device -> policy -> one step per launch
    n = 65536
    env = prl.BatchedEnvironment(world, n, device=local_rank, stream=torch.cuda.current_stream().cuda_stream)
    obs = torch.empty((n, env.n_bodies, 9), dtype=torch.float32, device="cuda")
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(200):
            env.state(out=obs)
            moves = (2 + 4 * (obs[:, 0, 4] < 0.01).to(torch.uint8)).contiguous()
            env.step(moves)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"closed loop ({args.world}, {n} envs, observation -> torch policy -> step, one launch per step): {n * 200 / dt:.3e} env-steps/s")
if world_size > 1:
    dist.destroy_process_group()
