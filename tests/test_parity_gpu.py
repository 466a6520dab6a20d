"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same worlds and move
sequences. Tolerance: ZERO -- every f32 of every body and every score is compared by bit pattern
(the kernels are built with -fmad=false and evaluate the reference's expressions in order), which
is stricter than the north star's "within an f32 tolerance; bit-exact won()/step-of-win"."""
import json
import os

import numpy as np
import pytest

import physics_reinforcement_learning_environment_b200 as prl
from physics_reinforcement_learning_environment_b200 import _lib
from conftest import (WORLDS, edge_case_worlds, many_supports_world, random_genomes, random_world, random_world_many_pushables,
                      synth_world, world_dict, world_text)
from oracle import binding as oracle
from oracle.binding import OracleEnv

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", WORLDS)
def test_rollout_matches_oracle(name):
    n, L, repeat, steps = 512, 50, 20, 1000
    genomes = random_genomes(n, L, 21)
    genomes[:8] = np.arange(8, dtype=np.uint8)[:, None]
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), n)
    for early in (True, False):
        score, win = env.rollout(genomes, repeat, steps, early_exit=early)
        _, _, run = env.rollout_info()
        rs, rw, rsteps = oracle.rollout_batch(world_dict(name), genomes, repeat, steps, early, n_threads=os.cpu_count())
        assert np.array_equal(bits(score), bits(rs))
        assert np.array_equal(win, rw)
        assert int(run.sum()) == rsteps
    env.sync()


@pytest.mark.parametrize("name", WORLDS)
def test_step_by_step_state_matches_oracle(name):
    """One kernel launch per step (state round-trips through HBM, manifolds included)."""
    n, steps, repeat = 8, 600, 10
    genomes = random_genomes(n, steps // repeat, 33)
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), n)
    oracles = [OracleEnv(world_dict(name)) for _ in range(n)]
    for t in range(steps):
        mv = np.ascontiguousarray(genomes[:, t // repeat])
        env.step(mv)
        for i, o in enumerate(oracles):
            o.step(int(mv[i]))
        if t % 7 == 0 or t == steps - 1:
            st = env.state()
            ref = np.stack([o.state() for o in oracles])
            assert np.array_equal(bits(st), bits(ref)), f"step {t + 1}"
            d = env.distance_to_goals()
            assert np.array_equal(bits(d), bits(np.array([o.distance_to_goals() for o in oracles], np.float32)))
            assert np.array_equal(env.won(), np.array([o.won() for o in oracles], np.uint8))


def test_uniform_moves_readme_example():
    """README.md:96-114 / BASELINE config C1: the 8 constant Moves x 1000 steps on right.json."""
    world = prl.World.from_json(world_text("right"))
    for mv in range(8):
        env = prl.BatchedEnvironment(world, 4)
        env.step_uniform(prl.Move.from_bits(mv), 1000)
        score, win, run = env.rollout_info()
        o = OracleEnv(world_dict("right"))
        best, owin = np.float32(np.inf), -1
        for t in range(1000):
            o.step(mv)
            best = min(best, o.distance_to_goals())
            if o.won() and owin < 0:
                owin = t + 1
        assert np.all(bits(score) == bits(np.float32(best))) and np.all(win == owin) and np.all(run == 1000)
        assert np.array_equal(bits(env.state()[0]), bits(o.state()))


def test_synthetic_64_obstacles_and_edge_case_world():
    for world in (synth_world(),):
        text = json.dumps(world)
        n, L, repeat, steps = 256, 60, 20, 1200
        genomes = random_genomes(n, L, 5)
        genomes[:4] = 2
        genomes[4:8] = 6
        env = prl.BatchedEnvironment(prl.World.from_json(text), n)
        score, win = env.rollout(genomes, repeat, steps, early_exit=False)
        rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
        assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
        env.sync()


def test_two_pushable_blocks_rotations_negative_scales():
    world = {"player_position": [0.0, 40.0], "objects": [
        {"object": {"Block": {"fixed": True}}, "position": [0.0, -40.0, 1.0], "scale": [-900.0, 50.0], "rotation": 0.05},
        {"object": {"Block": {"fixed": False}}, "position": [120.0, 10.0, 2.0], "scale": [60.0, -40.0], "rotation": 0.3},
        {"object": {"Block": {"fixed": False}}, "position": [125.0, 70.0, 3.0], "scale": [50.0, 30.0], "rotation": -0.2},
        {"object": {"Block": {"fixed": True}}, "position": [-200.0, 0.0, 4.0], "scale": [40.0, 80.0], "rotation": -0.4},
        {"object": "Goal", "position": [400.0, 0.0, 5.0], "scale": [50.0, 50.0], "rotation": 0.7},
        {"object": "Goal", "position": [-400.0, 0.0, 6.0], "scale": [-50.0, 50.0], "rotation": 0.0}]}
    n, L, repeat, steps = 128, 40, 15, 600
    genomes = random_genomes(n, L, 9)
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
    # final body states too
    st = env.state(0, 4)
    for i in range(4):
        o = OracleEnv(world)
        for t in range(steps):
            o.step(int(genomes[i][t // repeat]))
        assert np.array_equal(bits(st[i]), bits(o.state()))


@pytest.mark.parametrize("n_posts", [6, 8, 20])
def test_more_rows_than_a_column_holds_solves_in_rounds(n_posts):
    """A plank on many posts needs 12-18 constraint rows per environment: more than a column of
    the warp's row storage, so the lanes of a warp take turns (two / four rounds). Rollouts and
    single steps must still match the oracle bit for bit."""
    world = many_supports_world(n_posts)
    n, L, repeat, steps = 200, 30, 20, 600
    genomes = random_genomes(n, L, 41 + n_posts)
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
    env.reset()
    o = [OracleEnv(world) for _ in range(3)]
    for t in range(120):
        env.step(np.ascontiguousarray(genomes[:, t // repeat]))
        for i in range(3):
            o[i].step(int(genomes[i][t // repeat]))
    st = env.state(0, 3)
    for i in range(3):
        assert np.array_equal(bits(st[i]), bits(o[i].state()))


@pytest.mark.parametrize("name", sorted(edge_case_worlds()))
def test_editor_edge_cases(name):
    """Zero-size blocks, a player spawned deep inside a block, several goals, a pile of pushable blocks."""
    world = edge_case_worlds()[name]
    n, L, repeat, steps = 160, 30, 20, 600
    genomes = random_genomes(n, L, 70)
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert np.isfinite(rs).all()
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
    st = env.state(0, 2)
    for i in range(2):
        o = OracleEnv(world)
        for t in range(steps):
            o.step(int(genomes[i][t // repeat]))
        assert np.array_equal(bits(st[i]), bits(o.state()))


def test_result_is_independent_of_batch_size_and_neighbours():
    """env i's trajectory must not depend on the batch it sits in (sharding invariance)."""
    genomes = random_genomes(1000, 50, 77)
    world = prl.World.from_json(world_text("push"))
    full = prl.BatchedEnvironment(world, 1000)
    s_full, w_full = full.rollout(genomes, 20, 1000, early_exit=True)
    lo, hi = 137, 137 + 301  # an odd-sized slice that straddles warps and blocks
    part = prl.BatchedEnvironment(world, hi - lo)
    s_part, w_part = part.rollout(np.ascontiguousarray(genomes[lo:hi]), 20, 1000, early_exit=True)
    assert np.array_equal(bits(s_full[lo:hi]), bits(s_part)) and np.array_equal(w_full[lo:hi], w_part)


def test_full_size_push_65536x2000_every_env_matches_oracle():
    """BASELINE config C3 at full size, i.e. the kernel variant bench.py times
    (`rollout_kernel<2, DevWorldSmall, 448>`: one 448-thread block per SM, barrier every step,
    regrouping through HBM every 128 steps): ALL 65536 environments are compared with the oracle
    bit for bit (score, step of win), plus the size-independent properties."""
    n, L, repeat, steps = 65536, 100, 20, 2000
    genomes = random_genomes(n, L, 0xB200)
    env = prl.BatchedEnvironment(prl.World.from_json(world_text("push")), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    assert np.isfinite(score).all() and (score >= 0).all()
    assert np.all(score[win > 0] == 0.0) and np.all(win[score > 0] == -1)
    rs, rw, rsteps = oracle.rollout_batch(world_dict("push"), genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert rsteps == n * steps
    bad = np.flatnonzero((bits(score) != bits(rs)) | (win != rw))
    assert bad.size == 0, f"{bad.size} of {n} environments differ from the oracle, first {bad[:8]}"
    # the GA's break-on-win semantics through the same kernel variant
    score, win = env.rollout(genomes, repeat, steps, early_exit=True)
    _, _, run = env.rollout_info()
    env.sync()
    rs, rw, rsteps = oracle.rollout_batch(world_dict("push"), genomes, repeat, steps, True, n_threads=os.cpu_count())
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw) and int(run.sum()) == rsteps


def test_ragged_batch_on_the_448_thread_kernel():
    """The smallest batches that take the 448-thread variant, with a ragged tail (the last block is
    partly empty and its idle lanes still take part in the votes, barriers and regrouping)."""
    sm = 148
    n = sm * 448 // 2 + 17  # 33152 + 17
    L, repeat, steps = 30, 20, 600
    distinct = 4096 + 17
    genomes = random_genomes(distinct, L, 0x448)
    genomes = np.ascontiguousarray(np.resize(genomes, (n, L)))  # rows repeat with period `distinct`
    env = prl.BatchedEnvironment(prl.World.from_json(world_text("push")), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world_dict("push"), np.ascontiguousarray(genomes[:distinct]), repeat, steps, False,
                                     n_threads=os.cpu_count())
    idx = np.arange(n) % distinct
    assert np.array_equal(bits(score), bits(rs[idx])) and np.array_equal(win, rw[idx])


@pytest.mark.parametrize("which", ["stack", "posts8", "posts6", "jump", "synth64"])
def test_448_thread_kernel_variants_with_several_pushable_blocks(which):
    """`rollout_kernel<MAXB, DevWorldSmall, 448>` (a pile of three pushable blocks) and
    `rollout_kernel<MAXB, DevWorld, 448>` (a plank on 8 posts: 17+ candidate pairs, the general
    tables) are only launched for batches of >= 33152 environments; 6 posts is push's own variant
    solving in rounds; jump.json / the synthetic 64-obstacle world run the single-body 64-thread
    variants at the same batch size.
    Every environment is compared with the oracle (4096 distinct genomes)."""
    world = {"stack": edge_case_worlds()["stack"], "posts8": many_supports_world(8), "posts6": many_supports_world(6),
             "jump": world_dict("jump"), "synth64": synth_world()}[which]
    n, L, repeat, steps = 148 * 448 // 2 + 448, 20, 20, 400
    distinct = 4096
    genomes = np.ascontiguousarray(np.resize(random_genomes(distinct, L, 0x99 + len(which)), (n, L)))
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world, np.ascontiguousarray(genomes[:distinct]), repeat, steps, False, n_threads=os.cpu_count())
    idx = np.arange(n) % distinct
    bad = np.flatnonzero((bits(score) != bits(rs[idx])) | (win != rw[idx]))
    assert bad.size == 0, f"{bad.size} of {n} environments differ from the oracle, first {bad[:8]}"


def test_reset_and_device_buffers():
    import torch
    world = prl.World.from_json(world_text("jump"))
    n = 300
    env = prl.BatchedEnvironment(world, n, stream=torch.cuda.current_stream().cuda_stream)
    genomes = random_genomes(n, 30, 4)
    s1, w1 = env.rollout(genomes, 20, 600)
    dg = torch.from_numpy(genomes).cuda()
    ds = torch.empty(n, dtype=torch.float32, device="cuda")
    dw = torch.empty(n, dtype=torch.int32, device="cuda")
    env.rollout(dg, 20, 600, out_score=ds, out_step_of_win=dw)
    torch.cuda.synchronize()
    assert np.array_equal(bits(ds.cpu().numpy()), bits(s1)) and np.array_equal(dw.cpu().numpy(), w1)
    # stepping after reset starts again from from_world
    env.reset()
    a = env.state()
    env.step_uniform(2, 50)
    env.reset()
    assert np.array_equal(bits(env.state()), bits(a))
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    env.distance_to_goals(d)
    torch.cuda.synchronize()
    assert np.array_equal(bits(d.cpu().numpy()), bits(env.distance_to_goals()))


@pytest.mark.parametrize("name", WORLDS)
def test_trajectory_replay_matches_a_cpu_replay(name):
    """The viewer bridge (train.rs:132-159): a GeneticAgent replayed on a fresh environment, every
    body recorded after every step -- also past the end of the genome (Move::default()) and past
    a win (no early exit), with host and device buffers."""
    import torch
    world = world_dict(name)
    n, L, repeat, steps = 5, 12, 20, 300
    genomes = random_genomes(n, L, 123)
    genomes[0] = 2
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), 3)
    before = env.state()
    state, dist, won = env.trajectory(genomes, repeat, steps)
    assert np.array_equal(bits(env.state()), bits(before))  # the batch itself is untouched
    for i in range(n):
        agent = prl.GeneticAgent([prl.Move.from_bits(int(b)) for b in genomes[i]], 0, repeat)
        o = OracleEnv(world)
        for t in range(steps):
            o.step(agent.get_move().bits())
            assert np.array_equal(bits(state[t, i]), bits(o.state())), f"env {i} step {t + 1}"
            assert bits(dist[t, i]) == bits(np.float32(o.distance_to_goals())) and won[t, i] == o.won()
    # device buffers
    dg = torch.from_numpy(genomes).cuda()
    ds = torch.empty((steps, n, env.n_bodies, 9), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().prl_batch_trajectory(env._h, n, dg.data_ptr(), L, repeat, steps, ds.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert np.array_equal(bits(ds.cpu().numpy()), bits(state))


def test_goalless_world_reports_no_goal():
    w = prl.World.from_json('{"player_position":[0.0,0.0],"objects":[]}')
    env = prl.BatchedEnvironment(w, 4)
    env.step_uniform(0, 3)
    assert not env.won().any()
    with pytest.raises(prl.PrlError) as e:
        env.distance_to_goals()
    assert e.value.status == _lib.PRL_ERR_NO_GOAL
    with pytest.raises(prl.PrlError):
        env.rollout(np.zeros((4, 2), np.uint8), 1, 2)


def test_closed_loop_policy_on_device_matches_cpu_agent():
    """Agent::get_move(&mut self, &Environment) (algorithm.rs:9) in batch: observations stay on
    the GPU (torch tensor), a policy computes the moves there, step() consumes them from device
    memory. The same policy driven by the oracle's state must produce the same trajectories."""
    import torch
    name = "jump"
    n, steps = 64, 400
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), n, stream=torch.cuda.current_stream().cuda_stream)
    obs = torch.empty((n, env.n_bodies, 9), dtype=torch.float32, device="cuda")
    thresholds = torch.linspace(0.0, 0.02, n, device="cuda")

    def policy(o, thr):
        # walk right; jump when horizontal speed has stalled below a per-agent threshold
        vx = o[:, 0, 4]
        return (2 + 4 * (vx < thr).to(torch.uint8)).to(torch.uint8)

    oracles = [OracleEnv(world_dict(name)) for _ in range(n)]
    for t in range(steps):
        env.state(out=obs)
        moves = policy(obs, thresholds).contiguous()
        env.step(moves)
        for i, o in enumerate(oracles):
            vx = np.float32(o.state()[0][4])
            o.step(6 if vx < np.float32(thresholds[i].item()) else 2)
    torch.cuda.synchronize()
    ref = np.stack([o.state() for o in oracles])
    assert np.array_equal(bits(env.state()), bits(ref))
    assert np.array_equal(env.won(), np.array([o.won() for o in oracles], np.uint8))


@pytest.mark.parametrize("seed", range(8))
def test_random_worlds_match_oracle(seed):
    world = random_world(seed)
    n, L, repeat, steps = 96, 30, 20, 600
    genomes = random_genomes(n, L, 300 + seed)
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
    st = env.state(0, 3)
    for i in range(3):
        o = OracleEnv(world)
        for t in range(steps):
            o.step(int(genomes[i][t // repeat]))
        assert np.array_equal(bits(st[i]), bits(o.state()))


def policy_params(env, n, hidden, seed, scale=1.0):
    """Random per-agent perceptrons whose moves actually vary with the observation."""
    P, D = env.policy_num_params(hidden), env.observation_dim
    rng = np.random.Generator(np.random.Philox(seed))
    th = (rng.standard_normal((P, n)) * scale).astype(np.float32)
    for j in range(hidden):
        th[j * (D + 1) + D - 1] *= np.float32(0.003)  # the distance input is hundreds of pixels: keep it from saturating the units
    return th


@pytest.mark.parametrize("name", ["right", "jump", "push"])
def test_observation_matches_oracle(name):
    """prl_batch_get_observation after 0, 1 and many steps: body state, on_ground / floor contacts /
    touching-a-pushable-block (the scan of common.rs:223-244), distance to the goals."""
    n, steps, repeat = 24, 240, 8
    genomes = random_genomes(n, steps // repeat, 57)
    genomes[0] = 0
    genomes[1] = 2
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), n)
    oracles = [OracleEnv(world_dict(name)) for _ in range(n)]
    assert env.observation_dim == 9 * env.n_bodies + 4
    seen_ground, seen_touch = False, False
    for t in range(steps + 1):
        if t in (0, 1, 2, 3) or t % 13 == 0 or t == steps:
            obs = env.observation()
            ref = np.stack([o.observation() for o in oracles])
            assert np.array_equal(bits(obs), bits(ref)), f"after {t} steps"
            seen_ground |= bool(obs[:, 9 * env.n_bodies].any())
            seen_touch |= bool(obs[:, 9 * env.n_bodies + 2].any())
            if t == 0:
                assert not obs[:, 9 * env.n_bodies:9 * env.n_bodies + 3].any()  # the narrow phase is empty before the first step
        if t < steps:
            mv = np.ascontiguousarray(genomes[:, t // repeat])
            env.step(mv)
            for i, o in enumerate(oracles):
                o.step(int(mv[i]))
    assert seen_ground and (seen_touch or name != "push")


@pytest.mark.parametrize("name,hidden", [("right", 4), ("jump", 16), ("push", 16), ("push", 32)])
def test_policy_rollout_inside_the_kernel_matches_oracle(name, hidden):
    """Agent::get_move evaluated in the fused rollout: a per-agent perceptron on the observation,
    every step, against the oracle driven by the same parameters -- bit for bit, both early-exit modes."""
    n, steps = 300, 500
    env = prl.BatchedEnvironment(prl.World.from_json(world_text(name)), n)
    params = policy_params(env, n, hidden, 1000 + hidden)
    for early in (True, False):
        score, win = env.rollout_policy(params, hidden, steps, early_exit=early)
        env.sync()
        rs, rw, rsteps = oracle.rollout_policy_batch(world_dict(name), params, hidden, steps, early, n_threads=os.cpu_count())
        assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
        assert int(env.rollout_info()[2].sum()) == rsteps
    # the policies are not trivial: the population plays several different moves
    assert len(np.unique(bits(score))) > 12


def test_policy_rollout_device_buffers_and_edge_world():
    import torch
    world = edge_case_worlds()["stack"]  # three pushable blocks: the general kernel variant, 40 observation values
    n, hidden, steps = 130, 8, 300
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n, stream=torch.cuda.current_stream().cuda_stream)
    assert env.observation_dim == 9 * 4 + 4
    params = policy_params(env, n, hidden, 5)
    dp = torch.from_numpy(params).cuda()
    ds = torch.empty(n, dtype=torch.float32, device="cuda")
    dw = torch.empty(n, dtype=torch.int32, device="cuda")
    env.rollout_policy(dp, hidden, steps, early_exit=False, out_score=ds, out_step_of_win=dw)
    torch.cuda.synchronize()
    rs, rw, _ = oracle.rollout_policy_batch(world, params, hidden, steps, False, n_threads=os.cpu_count())
    assert np.array_equal(bits(ds.cpu().numpy()), bits(rs)) and np.array_equal(dw.cpu().numpy(), rw)
    with pytest.raises(prl.PrlError):
        env.rollout_policy(params, 33, steps)  # hidden > 32


@pytest.mark.parametrize("seed", range(4))
def test_random_worlds_with_many_pushable_blocks_match_oracle(seed):
    world = random_world_many_pushables(seed)
    n, L, repeat, steps = 96, 20, 15, 300
    genomes = random_genomes(n, L, 700 + seed)
    env = prl.BatchedEnvironment(prl.World.from_json(json.dumps(world)), n)
    score, win = env.rollout(genomes, repeat, steps, early_exit=False)
    env.sync()
    rs, rw, _ = oracle.rollout_batch(world, genomes, repeat, steps, False, n_threads=os.cpu_count())
    assert np.array_equal(bits(score), bits(rs)) and np.array_equal(win, rw)
