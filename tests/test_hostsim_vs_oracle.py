"""The product's per-environment device logic (csrc/prl_device.cuh), compiled for the host by
tests/hostsim, against the oracle: every body's state must agree BIT FOR BIT at every step.
Runs without a GPU; the GPU tests repeat this through the C ABI on the real kernels."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import (ROOT, WORLDS, edge_case_worlds, many_supports_world, n_bodies_of, random_genomes, random_world,
                      random_world_many_pushables, synth_world, world_dict)
from oracle.binding import OracleEnv

_hs = None


def hostsim():
    global _hs
    if _hs is None:
        _hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
        _hs.hs_rollout.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _hs


def host_rollout(world, genome, repeat, steps, early_exit, want_traj=True):
    js = json.dumps(world).encode()
    nb = n_bodies_of(world)
    total = len(genome) * repeat + steps % repeat
    traj = np.zeros((total, nb, 9), np.float32)
    sc, w, st, ov = C.c_float(), C.c_int(), C.c_int(), C.c_int()
    rc = hostsim().hs_rollout(js, len(js), genome.ctypes.data, len(genome), repeat, steps, int(early_exit),
                              traj.ctypes.data if want_traj else None, C.byref(sc), C.byref(w), C.byref(st), C.byref(ov))
    assert rc == 0 and ov.value == 0
    return traj[:st.value], np.float32(sc.value), w.value, st.value


def oracle_traj(world, genome, repeat, steps):
    env = OracleEnv(world)
    out = []
    for g in genome:
        for _ in range(repeat):
            env.step(int(g))
            out.append(env.state())
    for _ in range(steps % repeat):
        env.step(0)
        out.append(env.state())
    return np.stack(out)


@pytest.mark.parametrize("name", WORLDS)
def test_trajectories_bit_exact(name):
    world = world_dict(name)
    genomes = random_genomes(6, 50, 11)
    genomes[0] = 2
    genomes[1] = 6
    for g in genomes:
        t, _, _, _ = host_rollout(world, g, 20, 1000, False)
        o = oracle_traj(world, g, 20, 1000)
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))


@pytest.mark.parametrize("n_posts", [6, 8, 20])
def test_many_constraint_rows(n_posts):
    """12-18 constraint rows per environment (a plank on many posts)."""
    world = many_supports_world(n_posts)
    for g in random_genomes(3, 30, 41 + n_posts):
        t, _, _, _ = host_rollout(world, g, 20, 600, False)
        o = oracle_traj(world, g, 20, 600)
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))


@pytest.mark.parametrize("name", sorted(edge_case_worlds()))
def test_editor_edge_cases(name):
    world = edge_case_worlds()[name]
    for g in random_genomes(4, 30, 7):
        t, _, _, _ = host_rollout(world, g, 20, 600, False)
        o = oracle_traj(world, g, 20, 600)
        assert np.isfinite(o).all()
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))


def test_row_storage_layout_does_not_change_results():
    """The solver's constraint rows live in per-lane columns of shared-memory planes; with 4-row
    columns an environment with more than 4 rows spills into the neighbouring columns (the kernels
    then solve in rounds). The column height must not matter."""
    world = world_dict("push")
    hostsim().hs_set_row_cap(4)
    try:
        for g in random_genomes(3, 50, 12):
            t, _, _, _ = host_rollout(world, g, 20, 1000, False)
            o = oracle_traj(world, g, 20, 1000)
            assert np.array_equal(t.view(np.uint32), o.view(np.uint32))
    finally:
        hostsim().hs_set_row_cap(8)


@pytest.mark.parametrize("name", WORLDS)
def test_rollout_scores_and_early_exit(name):
    world = world_dict(name)
    env = OracleEnv(world)
    for g in random_genomes(8, 25, 5):
        for early in (True, False):
            _, sc, win, run = host_rollout(world, g, 20, 507, early, want_traj=False)
            osc, owin, orun = env.rollout(g, 20, 507, early)
            assert sc.view(np.uint32) == osc.view(np.uint32) and win == owin and run == orun


def test_synthetic_64_obstacle_world():
    world = synth_world()
    for g in random_genomes(3, 60, 3):
        t, _, _, _ = host_rollout(world, g, 20, 1200, False)
        o = oracle_traj(world, g, 20, 1200)
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))


def test_editor_edge_cases_rotations_negative_scales_two_pushables():
    # rotated fixed blocks, negative scales (editor drags, editor.rs:409-423), two pushable blocks
    # stacked (dynamic-dynamic contact), two goals
    world = {"player_position": [0.0, 40.0], "objects": [
        {"object": {"Block": {"fixed": True}}, "position": [0.0, -40.0, 1.0], "scale": [-900.0, 50.0], "rotation": 0.05},
        {"object": {"Block": {"fixed": False}}, "position": [120.0, 10.0, 2.0], "scale": [60.0, -40.0], "rotation": 0.3},
        {"object": {"Block": {"fixed": False}}, "position": [125.0, 70.0, 3.0], "scale": [50.0, 30.0], "rotation": -0.2},
        {"object": {"Block": {"fixed": True}}, "position": [-200.0, 0.0, 4.0], "scale": [40.0, 80.0], "rotation": -0.4},
        {"object": "Goal", "position": [400.0, 0.0, 5.0], "scale": [50.0, 50.0], "rotation": 0.7},
        {"object": "Goal", "position": [-400.0, 0.0, 6.0], "scale": [-50.0, 50.0], "rotation": 0.0}]}
    for g in random_genomes(4, 40, 9):
        t, _, _, _ = host_rollout(world, g, 15, 600, False)
        o = oracle_traj(world, g, 15, 600)
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))
        assert not np.isnan(t).any()


@pytest.mark.parametrize("seed", range(16))
def test_random_worlds_bit_exact(seed):
    """Randomly generated levels (rotated / mirrored fixed blocks, up to 3 pushable blocks in random
    object order, several goals): every body, every step."""
    world = random_world(seed)
    for g in random_genomes(2, 30, 100 + seed):
        t, _, _, _ = host_rollout(world, g, 20, 600, False)
        o = oracle_traj(world, g, 20, 600)
        assert not np.isnan(o).any()
        assert np.array_equal(t.view(np.uint32), o.view(np.uint32))


def host_policy_rollout(world, theta_col, hidden, steps, early_exit):
    """theta_col: (P, n) parameter-major array and a column index -> the host build of the policy rollout."""
    theta, col = theta_col
    js = json.dumps(world).encode()
    hs = hostsim()
    hs.hs_policy_rollout.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    D = 9 * n_bodies_of(world) + 4
    obs = np.zeros((steps, D), np.float32)
    sc, w, st, ov = C.c_float(), C.c_int(), C.c_int(), C.c_int()
    rc = hs.hs_policy_rollout(js, len(js), theta.ctypes.data + 4 * col, theta.shape[1], hidden, steps, int(early_exit), obs.ctypes.data,
                              C.byref(sc), C.byref(w), C.byref(st), C.byref(ov))
    assert rc == 0 and ov.value == 0
    return obs[:st.value], np.float32(sc.value), w.value, st.value


@pytest.mark.parametrize("name", ["right", "jump", "push", "stack"])
def test_policy_rollout_and_observation_bit_exact(name):
    """The device logic of the observation-driven rollout (Stepper::observe + policy_move, host build)
    against the oracle's independent restatement: every observation of every step, the moves that
    follow from them (through the trajectories), scores and steps of win."""
    from oracle import binding as oracle
    world = world_dict(name) if name in WORLDS else edge_case_worlds()[name]
    nb = n_bodies_of(world)
    D, hidden, n, steps = 9 * nb + 4, 8, 3, 300
    P = hidden * (D + 1) + 3 * (hidden + 1)
    theta = np.random.Generator(np.random.Philox(31)).standard_normal((P, n)).astype(np.float32)
    for early in (True, False):
        rs, rw, _ = oracle.rollout_policy_batch(world, theta, hidden, steps, early, n_threads=2)
        for a in range(n):
            obs, sc, win, run = host_policy_rollout(world, (theta, a), hidden, steps, early)
            assert sc.view(np.uint32) == rs[a].view(np.uint32) and win == rw[a]
            if not early:
                env = OracleEnv(world)
                lib = oracle.lib()
                for t in range(run):
                    ref = env.observation()
                    assert np.array_equal(obs[t].view(np.uint32), ref.view(np.uint32)), (a, t)
                    mv = lib.orc_policy_move(theta.ctypes.data + 4 * a, n, hidden, D, ref.ctypes.data)
                    env.step(mv)


@pytest.mark.parametrize("seed", range(6))
def test_random_worlds_with_many_pushable_blocks_bit_exact(seed):
    """4-7 pushable blocks (the 8-body variant, 32 manifold slots). 160 further seeds of this generator and
    480 of random_world were run offline in round 2: no mismatch, no overflow."""
    world = random_world_many_pushables(seed)
    g = random_genomes(1, 20, 900 + seed)[0]
    t, _, _, _ = host_rollout(world, g, 15, 300, False)
    o = oracle_traj(world, g, 15, 300)
    assert np.array_equal(t.view(np.uint32), o.view(np.uint32))
