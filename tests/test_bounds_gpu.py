"""Memory safety of the kernels without compute-sanitizer (not available on the GPU pool): the library
built with -DPRL_BOUNDS_CHECK (libprl_debug.so, csrc/Makefile) checks every run-time index into the
solver's row planes, the manifold records, the world tables, the HBM state slots, the genomes and the
regroup tables, and records failures in a device word. A workload that reaches every kernel variant and
every storage shape (ragged batches, rounds, regrouping, sparse warps) must leave it at zero."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKLOAD = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, os.environ["PRL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["PRL_ROOT"], "tests"))
import physics_reinforcement_learning_environment_b200 as prl
from physics_reinforcement_learning_environment_b200 import _lib
from conftest import edge_case_worlds, many_supports_world, random_genomes, random_world, synth_world, world_text
assert _lib.LIB_PATH.endswith("libprl_debug.so")
assert _lib.lib().prl_debug_bounds_violations() == 0
worlds = {"right": world_text("right"), "jump": world_text("jump"), "push": world_text("push"), "synth64": json.dumps(synth_world()),
          "posts6": json.dumps(many_supports_world(6)), "posts8": json.dumps(many_supports_world(8))}
worlds.update({k: json.dumps(v) for k, v in edge_case_worlds().items()})
worlds.update({f"random{s}": json.dumps(random_world(s)) for s in range(4)})
for name, text in worlds.items():
    w = prl.World.from_json(text)
    for n in (1, 33, 700, 4096 + 5):            # sparse warps, ragged 64-thread blocks
        env = prl.BatchedEnvironment(w, n)
        env.rollout(random_genomes(n, 15, 3), 20, 307, early_exit=(n % 2 == 0))
        env.step(np.ascontiguousarray(random_genomes(n, 1, 4)[:, 0]))
        env.step_uniform(6, 40)
        env.observation(); env.state(); env.distance_to_goals(); env.won()
        env.rollout_policy((np.random.default_rng(n).standard_normal((env.policy_num_params(8), n)) * 0.2).astype(np.float32), 8, 120)
        env.trajectory(random_genomes(3, 8, 5), 20, 200)
        try:
            env.sync()
        except prl.PrlError as e:
            assert e.status == _lib.PRL_ERR_OVERFLOW
    if name in ("push", "posts6", "posts8", "stack", "synth64", "six_pushables"):
        n = 148 * 448 // 2 + 77                 # SM-wide blocks: lock step, regrouping, a ragged last block
        env = prl.BatchedEnvironment(w, n)
        env.rollout(np.resize(random_genomes(2048, 15, 7), (n, 15)).copy(), 20, 300, early_exit=(name == "push"))
pop = prl.GaPopulation(prl.World.from_json(worlds["push"]), number_of_steps=200, number_of_agents=3001, devices=[0, 0])
for _ in range(3):
    pop.step_generation()
pop.close()
print("BOUNDS", _lib.lib().prl_debug_bounds_violations())
'''


def test_no_index_leaves_its_array_in_the_bounds_checked_build():
    debug = os.path.join(ROOT, "physics_reinforcement_learning_environment_b200", "libprl_debug.so")
    assert os.path.exists(debug), "make -C physics_reinforcement_learning_environment_b200/csrc builds it"
    env = dict(os.environ, PRL_LIBRARY=debug, PRL_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-c", WORKLOAD], env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "BOUNDS 0" in out.stdout, out.stdout[-500:]


def test_a_failed_check_is_recorded():
    """Negative control: with PRL_BOUNDS_SELFTEST the debug library runs a check that must fail (code 30)."""
    debug = os.path.join(ROOT, "physics_reinforcement_learning_environment_b200", "libprl_debug.so")
    env = dict(os.environ, PRL_LIBRARY=debug, PRL_ROOT=ROOT, PRL_BOUNDS_SELFTEST="1")
    code = ("import os, sys; sys.path.insert(0, os.environ['PRL_ROOT']);"
            "from physics_reinforcement_learning_environment_b200 import _lib; print('BOUNDS', _lib.lib().prl_debug_bounds_violations())")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert f"BOUNDS {1 << 30}" in out.stdout, out.stdout + out.stderr[-2000:]


def test_the_product_library_has_no_checks():
    from physics_reinforcement_learning_environment_b200 import _lib
    assert _lib.lib().prl_debug_bounds_violations() == -1
