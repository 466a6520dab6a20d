#!/usr/bin/env python
"""Rollout throughput on one of the test levels of tests/conftest.py (edge cases, many supports):
python scripts/world_bench.py stack [envs] [steps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402
from conftest import edge_case_worlds, many_supports_world, random_genomes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "stack"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
worlds = dict(edge_case_worlds(), many_supports=many_supports_world(6), many_supports8=many_supports_world(8))
world = prl.World.from_json(json.dumps(worlds[name]))
env = prl.BatchedEnvironment(world, n, stream=torch.cuda.current_stream().cuda_stream)
genomes = torch.from_numpy(random_genomes(n, steps // 20, 3)).cuda()
score = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    env.rollout(genomes, 20, steps, early_exit=False, out_score=score)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    env.rollout(genomes, 20, steps, early_exit=False, out_score=score)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(f"{name}: {n} envs x {steps} steps, {world.counts()[0]} bodies: {ms:.1f} ms per rollout, {n * steps / ms * 1e3:.3e} env-steps/s")
