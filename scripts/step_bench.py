#!/usr/bin/env python
"""One launch per step (prl_batch_step / step_kernel): the mode closed-loop agents use. Here the
whole state (bodies + manifold slots) streams through HBM every step, so this kernel -- unlike the
fused rollout -- has a meaningful HBM figure: state bytes read + written per launch / launch time."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402

world_name = sys.argv[1] if len(sys.argv) > 1 else "push"
world = prl.World.load(os.path.join(ROOT, "worlds", f"{world_name}.json"))
for n in (65536, 1 << 20):
    env = prl.BatchedEnvironment(world, n, stream=torch.cuda.current_stream().cuda_stream)
    moves = torch.from_numpy(np.random.default_rng(0).integers(0, 8, n, dtype=np.uint8)).cuda()
    for _ in range(50):
        env.step(moves)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    K = 300
    for _ in range(K):
        env.step(moves)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"{world_name} step_kernel n={n}: {ms * 1e3:.1f} us/launch, {n / ms * 1e3:.3e} env-steps/s, "
          f"state {env.state_bytes / 1e6:.1f} MB -> {2 * env.state_bytes / ms / 1e6:.0f} GB/s if the whole state is read and written")
