"""Committed golden vectors (tests/golden/*.npz, made by scripts/make_golden.py from the oracle --
the reference ships none and cannot be run here: PARITY UNPINNED) against the oracle (CPU) and the
CUDA path (GPU)."""
import glob
import json
import os

import numpy as np
import pytest

import physics_reinforcement_learning_environment_b200 as prl
from conftest import ROOT, WORLDS, world_dict, world_text
from oracle.binding import OracleEnv


GOLDEN = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))


def golden_world(name, g):
    """(dict, text): the shipped levels by name, the others from the JSON stored in the fixture"""
    if name in WORLDS:
        return world_dict(name), world_text(name)
    text = str(g["world_json"])
    return json.loads(text), text


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_reproduces_golden(name):
    g = golden(name)
    env = OracleEnv(golden_world(name, g)[0])
    for i, genome in enumerate(g["genomes"]):
        sc, win, run = env.rollout(genome, int(g["repeat_move"]), int(g["number_of_steps"]), True)
        assert sc.view(np.uint32) == g["score"][i].view(np.uint32) and win == g["step_of_win"][i] and run == g["steps_run"][i]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_reproduces_golden(name):
    g = golden(name)
    repeat, steps = int(g["repeat_move"]), int(g["number_of_steps"])
    text = golden_world(name, g)[1]
    env = prl.BatchedEnvironment(prl.World.from_json(text), 16)
    score, win = env.rollout(np.ascontiguousarray(g["genomes"]), repeat, steps, early_exit=True)
    _, _, run = env.rollout_info()
    assert np.array_equal(score.view(np.uint32), g["score"].view(np.uint32))
    assert np.array_equal(win, g["step_of_win"]) and np.array_equal(run, g["steps_run"])
    # sampled trajectories (every 10th step), stepping one launch per step through prl_batch_step
    tg = g["traj_genomes"]
    env4 = prl.BatchedEnvironment(prl.World.from_json(text), len(tg))
    for t in range(steps):
        env4.step(np.ascontiguousarray(g["genomes"][tg, t // repeat]))
        if t % 10 == 0:
            assert np.array_equal(env4.state().view(np.uint32), g["traj"][:, t // 10].view(np.uint32)), t
