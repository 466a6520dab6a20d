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


def many_supports_world(n_posts=6):
    """One pushable plank resting on n fixed posts. Two solver contacts per post plus the
    player's make an environment need more constraint rows than one column of the solver's row
    storage holds, so the warps of the kernels solve in rounds (prl_kernels.cuh,
    solve_in_rounds): 6 posts = 13 candidate pairs, the small-world kernel (9-row columns, 12-14
    rows: two rounds); 8 posts = the general kernel (8-row columns, 16-18 rows: four rounds)."""
    objs = []
    for i in range(n_posts):
        objs.append({"object": {"Block": {"fixed": True}}, "position": [-150.0 + 60.0 * i, -40.0, float(i + 1)],
                     "scale": [30.0, 40.0], "rotation": 0.0})
    objs.append({"object": {"Block": {"fixed": False}}, "position": [-150.0 + 30.0 * (n_posts - 1), -14.0, float(n_posts + 1)],
                 "scale": [60.0 * n_posts + 40.0, 12.0], "rotation": 0.0})
    objs.append({"object": "Goal", "position": [60.0 * n_posts - 100.0, 30.0, float(n_posts + 2)], "scale": [60.0, 60.0], "rotation": 0.0})
    return {"player_position": [-100.0, 24.0], "objects": objs}


def edge_case_worlds():
    """Levels the editor can produce that the shipped examples do not cover: a block
    dragged to zero size, the player spawned deep inside a block (parry's penetration path), three
    goals of which one is rotated and one has a negative scale, and pushable blocks stacked on each
    other (body-body contacts in a pile) with the player dropped on top."""
    fixed = lambda x, y, w, h, r=0.0, z=1.0: {"object": {"Block": {"fixed": True}}, "position": [x, y, z], "scale": [w, h], "rotation": r}
    dyn = lambda x, y, w, h, r=0.0, z=1.0: {"object": {"Block": {"fixed": False}}, "position": [x, y, z], "scale": [w, h], "rotation": r}
    goal = lambda x, y, w, h, r=0.0, z=9.0: {"object": "Goal", "position": [x, y, z], "scale": [w, h], "rotation": r}
    return {
        "zero_size_block": {"player_position": [0.0, 10.0], "objects": [
            fixed(0.0, -50.0, 800.0, 40.0), fixed(60.0, -25.0, 0.0, 30.0), fixed(-60.0, -30.0, 40.0, 0.0), goal(300.0, 0.0, 80.0, 80.0)]},
        "deep_penetration": {"player_position": [0.0, 0.0], "objects": [
            fixed(5.0, -4.0, 200.0, 120.0), fixed(260.0, -80.0, 300.0, 40.0, 0.1), goal(330.0, -20.0, 60.0, 60.0)]},
        "three_goals": {"player_position": [-50.0, 0.0], "objects": [
            goal(-300.0, 10.0, 40.0, 40.0, 1.2), fixed(0.0, -50.0, 900.0, 40.0), goal(200.0, 20.0, -60.0, 90.0), goal(60.0, 120.0, 30.0, 30.0, -0.5)]},
        # six pushable blocks (the reference has no limit, common.rs:127-146; the kernels keep up to 7 in
        # one environment): two piles of two and two loose ones on a floor with a wall, the player dropped on a pile
        "six_pushables": {"player_position": [-20.0, 150.0], "objects": [
            fixed(0.0, -50.0, 1000.0, 40.0), dyn(-30.0, -10.0, 80.0, 40.0), dyn(-25.0, 32.0, 60.0, 40.0), fixed(330.0, 0.0, 40.0, 200.0),
            dyn(120.0, -10.0, 50.0, 40.0), dyn(125.0, 28.0, 40.0, 30.0, 0.05), dyn(220.0, -15.0, 40.0, 30.0), dyn(-150.0, -12.0, 70.0, 36.0, -0.02),
            goal(280.0, 0.0, 60.0, 60.0)]},
        "stack": {"player_position": [0.0, 160.0], "objects": [
            fixed(0.0, -50.0, 700.0, 40.0), dyn(0.0, -10.0, 120.0, 40.0), dyn(10.0, 35.0, 90.0, 40.0), dyn(-5.0, 80.0, 60.0, 40.0, 0.1),
            goal(250.0, 0.0, 70.0, 70.0)]},
    }


def random_genomes(n, genome_len, seed):
    return np.random.Generator(np.random.Philox(seed)).integers(0, 8, size=(n, genome_len), dtype=np.uint8)


def has_cuda():
    try:
        from physics_reinforcement_learning_environment_b200 import _lib
        return _lib.lib().prl_device_count() > 0
    except Exception:
        return False


def random_world(seed):
    """A random level in the editor's schema: a long floor, fixed blocks (some rotated, some with
    negative scales), 0-3 pushable blocks resting or dropped on the floor, 1-2 goals."""
    rng = np.random.Generator(np.random.Philox(1000 + seed))
    objs = [{"object": {"Block": {"fixed": True}}, "position": [300.0, -60.0, 1.0], "scale": [1600.0, 50.0],
             "rotation": float(rng.uniform(-0.03, 0.03))}]
    z = 2.0
    for _ in range(int(rng.integers(1, 6))):
        objs.append({"object": {"Block": {"fixed": True}},
                     "position": [float(rng.uniform(-350, 900)), float(rng.uniform(-40, 30)), z],
                     "scale": [float(rng.uniform(20, 120)) * (1 if rng.random() < 0.8 else -1), float(rng.uniform(15, 70))],
                     "rotation": float(rng.uniform(-0.5, 0.5)) if rng.random() < 0.5 else 0.0})
        z += 1.0
    for _ in range(int(rng.integers(0, 4))):
        objs.append({"object": {"Block": {"fixed": False}},
                     "position": [float(rng.uniform(-250, 600)), float(rng.uniform(-20, 120)), z],
                     "scale": [float(rng.uniform(25, 150)), float(rng.uniform(12, 60))],
                     "rotation": float(rng.uniform(-0.4, 0.4)) if rng.random() < 0.5 else 0.0})
        z += 1.0
    for _ in range(int(rng.integers(1, 3))):
        objs.append({"object": "Goal", "position": [float(rng.uniform(-400, 900)), float(rng.uniform(-30, 80)), z],
                     "scale": [float(rng.uniform(30, 120)), float(rng.uniform(30, 120))], "rotation": float(rng.uniform(-0.6, 0.6))})
        z += 1.0
    order = rng.permutation(len(objs))  # object order decides collider / body indices
    return {"player_position": [float(rng.uniform(-100, 100)), float(rng.uniform(20, 90))], "objects": [objs[i] for i in order]}


def random_world_many_pushables(seed):
    """Like random_world but with 4-7 pushable blocks dropped over the floor: piles, body-body contacts,
    the 8-body kernel variant (5-7 pushables) and the 32-slot manifold lists."""
    rng = np.random.Generator(np.random.Philox(5000 + seed))
    objs = [{"object": {"Block": {"fixed": True}}, "position": [300.0, -60.0, 1.0], "scale": [1800.0, 50.0], "rotation": float(rng.uniform(-0.02, 0.02))}]
    for _ in range(int(rng.integers(0, 5))):
        objs.append({"object": {"Block": {"fixed": True}}, "position": [float(rng.uniform(-350, 900)), float(rng.uniform(-40, 30)), 2.0],
                     "scale": [float(rng.uniform(20, 120)), float(rng.uniform(15, 70))],
                     "rotation": float(rng.uniform(-0.5, 0.5)) if rng.random() < 0.5 else 0.0})
    for _ in range(int(rng.integers(4, 8))):
        objs.append({"object": {"Block": {"fixed": False}}, "position": [float(rng.uniform(-200, 500)), float(rng.uniform(-20, 160)), 3.0],
                     "scale": [float(rng.uniform(25, 120)), float(rng.uniform(12, 50))],
                     "rotation": float(rng.uniform(-0.4, 0.4)) if rng.random() < 0.5 else 0.0})
    objs.append({"object": "Goal", "position": [float(rng.uniform(-400, 900)), 20.0, 9.0], "scale": [80.0, 80.0], "rotation": 0.3})
    order = rng.permutation(len(objs))
    return {"player_position": [float(rng.uniform(-100, 100)), float(rng.uniform(40, 200))], "objects": [objs[i] for i in order]}
