#!/usr/bin/env python
"""Throughput of observation-driven agents on one B200: the per-agent perceptron
evaluated INSIDE the fused rollout (prl_batch_rollout_policy) next to the closed loop through the
host API (prl_batch_get_observation -> torch policy -> prl_batch_step, one launch per step) and the
open-loop genome rollout, all on push.json, 65536 environments."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import physics_reinforcement_learning_environment_b200 as prl  # noqa: E402


def main():
    n, steps, hidden = 65536, 1000, 16
    world = prl.World.load(os.path.join(ROOT, "worlds", "push.json"))
    stream = torch.cuda.current_stream()
    env = prl.BatchedEnvironment(world, n, stream=stream.cuda_stream)
    D, P = env.observation_dim, env.policy_num_params(hidden)
    params = (torch.randn((P, n), device="cuda") * 1.0).contiguous()
    score = torch.empty(n, dtype=torch.float32, device="cuda")
    win = torch.empty(n, dtype=torch.int32, device="cuda")
    out = {}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    ms = timed(lambda: env.rollout_policy(params, hidden, steps, early_exit=False, out_score=score, out_step_of_win=win))
    out["policy_in_kernel"] = {"env_steps_per_s": n * steps / (ms * 1e-3), "ms": ms, "hidden": hidden, "obs_dim": D,
                               "param_bytes_read_per_step": P * n * 4}
    genomes = torch.randint(0, 8, (n, steps // 20), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: env.rollout(genomes, 20, steps, early_exit=False, out_score=score, out_step_of_win=win))
    out["genome_rollout"] = {"env_steps_per_s": n * steps / (ms * 1e-3), "ms": ms}
    # closed loop through the host API: observation tensor -> the same perceptron in torch -> step
    obs = torch.empty((n, D), dtype=torch.float32, device="cuda")
    W1 = params[:hidden * (D + 1)].reshape(hidden, D + 1, n)
    W2 = params[hidden * (D + 1):].reshape(3, hidden + 1, n)

    def closed_loop(k=100):
        env.reset()
        for _ in range(k):
            env.observation(out=obs)
            h = torch.relu(torch.einsum("jin,ni->nj", W1[:, :D], obs) + W1[:, D].T)
            o = torch.einsum("kjn,nj->nk", W2[:, :hidden], h) + W2[:, hidden].T
            mv = ((o[:, 0] > 0).to(torch.uint8) | ((o[:, 1] > 0).to(torch.uint8) << 1) | ((o[:, 2] > 0).to(torch.uint8) << 2)).contiguous()
            env.step(mv)
    ms = timed(closed_loop, reps=2)
    out["closed_loop_host_api"] = {"env_steps_per_s": n * 100 / (ms * 1e-3), "ms_per_step": ms / 100}
    env.sync()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
