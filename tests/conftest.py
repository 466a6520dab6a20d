import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORLDS = ("right", "jump", "push")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Builds libprl.so (product), the C oracle and the host build of the device logic (checkers)."""
    import __graft_entry__
    __graft_entry__.build()


def world_text(name):
    return open(os.path.join(ROOT, "worlds", f"{name}.json")).read()


def world_dict(name):
    return json.loads(world_text(name))


def n_bodies_of(world):
    return 1 + sum(1 for o in world["objects"] if o["object"] != "Goal" and not o["object"]["Block"]["fixed"])


def synth_world(n_obstacles=63, seed=64):
    """floor strip + n fixed obstacles (x spacing 150 px, width U[20,60],
    height U[20,50], every 4th rotated U[-0.2,0.2] rad), goal at the far end. Editor pixels."""
    rng = np.random.Generator(np.random.Philox(seed))
    length = 150.0 * (n_obstacles + 2)
    objs = [{"object": {"Block": {"fixed": True}}, "position": [length / 2 - 100.0, -55.0, 1.0],
             "scale": [length + 400.0, 50.0], "rotation": 0.0}]
    for i in range(n_obstacles):
        w = float(rng.uniform(20, 60)); h = float(rng.uniform(20, 50))
        rot = float(rng.uniform(-0.2, 0.2)) if i % 4 == 3 else 0.0
        objs.append({"object": {"Block": {"fixed": True}}, "position": [150.0 * (i + 1), -30.0 + h / 2, float(i + 2)],
                     "scale": [w, h], "rotation": rot})
    objs.append({"object": "Goal", "position": [150.0 * (n_obstacles + 1), 0.0, float(n_obstacles + 2)],
                 "scale": [100.0, 100.0], "rotation": 0.0})
    return {"player_position": [0.0, 0.0], "objects": objs}


def random_genomes(n, genome_len, seed):
    return np.random.Generator(np.random.Philox(seed)).integers(0, 8, size=(n, genome_len), dtype=np.uint8)


def has_cuda():
    try:
        from physics_reinforcement_learning_environment_b200 import _lib
        return _lib.lib().prl_device_count() > 0
    except Exception:
        return False
