#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both rollout block shapes
(64-thread blocks and one 448-thread block per SM with regrouping), stepping and the GA kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402

world = prl.World.load(os.path.join(ROOT, "worlds", "push.json"))
rng = np.random.default_rng(3)
for n in (200, 148 * 448 // 2 + 37):  # small blocks / SM-wide blocks with a ragged last block
    env = prl.BatchedEnvironment(world, n)
    genomes = rng.integers(0, 8, size=(n, 8), dtype=np.uint8)
    score, win = env.rollout(genomes, 20, 165, early_exit=True)
    env.sync()
    print(n, float(score.min()), int((win > 0).sum()))
env = prl.BatchedEnvironment(world, 100)
for t in range(30):
    env.step(rng.integers(0, 8, size=100, dtype=np.uint8))
env.distance_to_goals(); env.won(); env.state(); env.sync()
pop = prl.GaPopulation(prl.World.load(os.path.join(ROOT, "worlds", "right.json")), number_of_steps=100, number_of_agents=300, repeat_move=20)
pop.init(); pop.best(); pop.next_generation(); pop.best()
print("sanitize case done")
