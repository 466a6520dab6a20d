"""The genetic algorithm (reference src/main.rs:41-163) on the device, through the C ABI:
bit-exact against tests/ga_ref.py (numpy restatement with the same Philox counters) and the
oracle's rollouts, sharding invariance with emulated ranks, statistics of selection / crossover /
mutation against the reference's semantics, and the train() protocol."""
import json
import os

import numpy as np
import pytest

import ga_ref
import physics_reinforcement_learning_environment_b200 as prl
from physics_reinforcement_learning_environment_b200 import _lib
from conftest import world_dict, world_text
from oracle import binding as oracle

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def make(world_name="right", **kw):
    cfg = dict(number_of_steps=200, number_of_agents=96, repeat_move=20, mutation_rate=0.1, keep_best=False, seed=1234)
    cfg.update(kw)
    return prl.GaPopulation(prl.World.from_json(world_text(world_name)), **cfg), cfg


def test_init_matches_philox_reference_and_oracle_scores():
    pop, cfg = make()
    pop.init()
    genomes = pop.local_genomes()
    assert genomes.shape == (96, 10)  # number_of_steps / repeat_move genes (main.rs:75)
    assert np.array_equal(genomes, ga_ref.init_genomes(96, 10, cfg["seed"]))
    rs, _, _ = oracle.rollout_batch(world_dict("right"), genomes, 20, 200, early_exit=True, n_threads=os.cpu_count())
    assert np.array_equal(bits(pop.local_scores()), bits(rs))
    score, genome, idx = pop.best()
    bi, bmin, _ = ga_ref.best(rs)
    assert idx == bi and bits(score) == bits(bmin) and np.array_equal(genome, genomes[bi])


@pytest.mark.parametrize("keep_best", [False, True])
def test_next_generation_children_bit_exact(keep_best):
    pop, cfg = make("push", keep_best=keep_best, number_of_agents=64, number_of_steps=300, mutation_rate=0.25)
    pop.init()
    for gen in (1, 2):
        parents, scores = pop.local_genomes(), pop.local_scores()
        pop.next_generation()
        assert pop.generation == gen
        ref, _ = ga_ref.breed(parents, scores, gen, cfg["seed"], cfg["mutation_rate"], keep_best)
        children = pop.local_genomes()
        for c in range(64):
            assert np.array_equal(children[c], ref[c]), (gen, c)
        rs, _, _ = oracle.rollout_batch(world_dict("push"), children, 20, 300, early_exit=True, n_threads=os.cpu_count())
        assert np.array_equal(bits(pop.local_scores()), bits(rs))
    if keep_best:
        # the elite is carried over unchanged (main.rs:126-131) and its score is its old score
        bi, bmin, _ = ga_ref.best(scores)
        assert np.array_equal(children[0], parents[bi]) and bits(pop.local_scores()[0]) == bits(bmin)


def test_population_sharding_does_not_change_the_result():
    """Three emulated ranks on one GPU, exchanging slices by hand, against one rank."""
    import torch
    world = prl.World.from_json(world_text("jump"))
    kw = dict(number_of_steps=240, number_of_agents=100, repeat_move=20, mutation_rate=0.1, keep_best=True, seed=77)
    single = prl.GaPopulation(world, **kw)
    single.init()
    ranks = [prl.GaPopulation(world, rank=r, n_ranks=3, **kw) for r in range(3)]
    for p in ranks:
        p.init()

    def exchange():
        torch.cuda.synchronize()
        for p in ranks:
            _, _, gs, gg = p.tensors()
            for q in ranks:
                ls, lg, _, _ = q.tensors()
                gs[q.local_offset:q.local_offset + q.local_count].copy_(ls)
                gg[q.local_offset:q.local_offset + q.local_count].copy_(lg)
        torch.cuda.synchronize()

    for gen in range(3):
        exchange()
        s1 = single.best()
        for p in ranks:
            sr = p.best()
            assert bits(sr[0]) == bits(s1[0]) and sr[2] == s1[2] and np.array_equal(sr[1], s1[1])
        whole = np.concatenate([p.local_genomes() for p in ranks])
        assert np.array_equal(whole, single.local_genomes())
        assert np.array_equal(bits(np.concatenate([p.local_scores() for p in ranks])), bits(single.local_scores()))
        single.next_generation()
        for p in ranks:
            p.next_generation()


@pytest.mark.parametrize("keep_best", [False, True])
def test_peer_shards_inside_the_c_abi_equal_the_unsharded_ga(keep_best):
    """prl_ga_create_multi through ctypes only (no torch.distributed, no host exchange): one slice
    and one stream per listed device, scores pushed into the peers' score vectors, parent genomes
    read from the owning slice by the breeding kernel. With one visible GPU the slices all sit on
    device 0 (the same code path: peer copies and peer pointers of a device to itself); with
    more, one slice per GPU over NVLink. Every genome and score of every generation must equal
    the unsharded GA's."""
    n_dev = _lib.lib().prl_device_count()
    devices = list(range(n_dev)) if n_dev >= 2 else [0, 0, 0]
    if len(devices) == 2:
        devices = [0, 1, 0]  # ragged three-way split, two slices on device 0
    world = prl.World.from_json(world_text("push"))
    kw = dict(number_of_steps=300, number_of_agents=1000 + 3, repeat_move=20, mutation_rate=0.1, keep_best=keep_best, seed=4242)
    single = prl.GaPopulation(world, **kw)
    multi = prl.GaPopulation(world, devices=devices, **kw)
    assert multi.num_shards == len(devices) and single.num_shards == 1
    for gen in range(4):
        b1 = single.step_generation()
        bm = multi.step_generation()
        assert bits(b1[0]) == bits(bm[0]) and b1[2] == bm[2] and np.array_equal(b1[1], bm[1]), gen
        assert single.generation == multi.generation == gen + 1
        s1, g1 = single.read_generation()
        sm, gm = multi.read_generation()
        assert np.array_equal(g1, gm), gen
        assert np.array_equal(bits(s1), bits(sm)), gen
    # and both equal the numpy restatement of the breeding (children of the last generation)
    assert multi.env_steps() == single.env_steps() > 0
    with pytest.raises(prl.PrlError):
        multi.tensors()  # no host-visible exchange buffers in this mode


def test_single_rank_buffers_alias_and_swap():
    """include/prl.h: with one rank local_* and global_* are the same memory (no copy per
    generation) and hold the CURRENT generation right after init / next_generation."""
    pop, cfg = make("jump", number_of_agents=50)
    pop.init()
    ls, lg, gs, gg = pop.tensors()
    assert ls.data_ptr() == gs.data_ptr() and lg.data_ptr() == gg.data_ptr()
    g0, s0 = lg.cpu().numpy().copy(), ls.cpu().numpy().copy()
    assert np.array_equal(bits(s0), bits(_scores_of(g0, "jump", 200)))
    assert np.array_equal(g0, ga_ref.init_genomes(50, 10, cfg["seed"]))
    pop.next_generation()
    ls2, lg2, gs2, gg2 = pop.tensors()
    assert lg2.data_ptr() == gg2.data_ptr() and lg2.data_ptr() != lg.data_ptr()
    ref, _ = ga_ref.breed(g0, s0, 1, cfg["seed"], cfg["mutation_rate"], False)
    children = lg2.cpu().numpy()
    assert all(np.array_equal(children[c], ref[c]) for c in range(50))


def _scores_of(genomes, name, steps):
    rs, _, _ = oracle.rollout_batch(world_dict(name), genomes, 20, steps, early_exit=True, n_threads=os.cpu_count())
    return rs


def test_ga_reports_contact_overflow_instead_of_selecting_on_inexact_scores():
    """A world whose environments need more simultaneous contact manifolds than PRL_MAX_CONTACTS:
    the scores are not exact, so prl_ga_best must fail with PRL_ERR_OVERFLOW rather than report a
    best agent (the batch API reports the same through prl_batch_sync)."""
    # a long plank resting on 40 posts: 40 plank-post manifolds alive at once (the kernels hold 32)
    n_posts = 40
    objs = [{"object": {"Block": {"fixed": True}}, "position": [-300.0 + 30.0 * i, -40.0, float(i + 1)], "scale": [14.0, 40.0], "rotation": 0.0}
            for i in range(n_posts)]
    objs.append({"object": {"Block": {"fixed": False}}, "position": [-300.0 + 15.0 * (n_posts - 1), -14.0, 50.0], "scale": [30.0 * n_posts + 20.0, 12.0], "rotation": 0.0})
    objs.append({"object": "Goal", "position": [500.0, 30.0, 31.0], "scale": [60.0, 60.0], "rotation": 0.0})
    world = prl.World.from_json(json.dumps({"player_position": [0.0, 24.0], "objects": objs}))
    env = prl.BatchedEnvironment(world, 8)
    env.step_uniform(0, 5)
    with pytest.raises(prl.PrlError) as e:
        env.sync()
    assert e.value.status == _lib.PRL_ERR_OVERFLOW
    pop = prl.GaPopulation(world, number_of_steps=40, number_of_agents=16, repeat_move=20)
    pop.init()
    with pytest.raises(prl.PrlError) as e:
        pop.best()
    assert e.value.status == _lib.PRL_ERR_OVERFLOW


def test_selection_crossover_mutation_statistics():
    """choose_multiple_weighted(2, w = max + 1 - score) (main.rs:134-138): two DISTINCT parents,
    the first drawn with probability w_i / sum(w); uniform crossover (main.rs:140-147); each of
    the three bits re-drawn with probability mutation_rate (main.rs:148-158), i.e. flipped with
    probability rate / 2."""
    n, L = 64, 10
    pop, cfg = make("right", number_of_agents=n, mutation_rate=0.2, seed=9)
    pop.init()
    parents, scores = pop.local_genomes(), pop.local_scores()
    w = np.array(ga_ref.fixed_weights(scores), dtype=np.float64)
    p_first = w / w.sum()
    counts = np.zeros(n)
    pairs, flips, from_p1, total_genes = 0, 0, 0, 0
    gens = 60
    for gen in range(1, gens + 1):
        _, chosen = ga_ref.breed(parents, scores, gen, cfg["seed"], 0.0, False)  # which parents the device must have used
        no_mut, _ = ga_ref.breed(parents, scores, gen, cfg["seed"], 0.0, False)
        mut, _ = ga_ref.breed(parents, scores, gen, cfg["seed"], 0.2, False)
        for c in range(n):
            p1, p2 = chosen[c]
            assert p1 != p2
            counts[p1] += 1
            pairs += 1
            x = no_mut[c] ^ mut[c]
            flips += int(np.unpackbits(x.astype(np.uint8)).sum())
            differ = parents[p1] != parents[p2]
            from_p1 += int((no_mut[c][differ] == parents[p1][differ]).sum())
            total_genes += int(differ.sum())
    # the device agrees with the restatement on generation 1 (bit-exact test above covers more)
    pop.next_generation()
    ref1, _ = ga_ref.breed(parents, scores, 1, cfg["seed"], 0.2, False)
    assert all(np.array_equal(pop.local_genomes()[c], ref1[c]) for c in range(n))
    # first-parent frequencies follow the weights (chi-square, 63 dof: mean 63, sd 11)
    expected = p_first * pairs
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 63 + 5 * 11.2, chi2
    # crossover is a fair coin on genes where the parents differ
    assert abs(from_p1 / total_genes - 0.5) < 0.02
    # each bit flips with probability rate / 2 = 0.1
    assert abs(flips / (pairs * L * 3) - 0.1) < 0.01


def test_train_protocol_and_learning_on_right_json():
    """train() sends (best score, agent) once per generation until the receiver goes away
    (main.rs:112-124); on right.json the population learns to walk right and wins."""
    world = prl.World.from_json(world_text("right"))
    algo = prl.GeneticAlgorithm(number_of_steps=400, number_of_agents=512, repeat_move=20, mutation_rate=0.1, keep_best=True, seed=5)
    msgs = []

    def sender(msg):
        msgs.append(msg)
        return len(msgs) < 12 and msg[0] > 0.0  # "receiver dropped" after a win or 12 generations

    gens = algo.train(world, sender)
    assert gens == len(msgs) and 1 <= gens <= 12
    scores = [m[0] for m in msgs]
    assert all(scores[i + 1] <= scores[i] for i in range(len(scores) - 1))  # keep_best: never worse
    assert scores[-1] == 0.0
    # replay the winning agent on the CPU environment, as the reference's visualiser does (train.rs:132-159)
    agent = msgs[-1][1]
    env = oracle.OracleEnv(world_dict("right"))
    for _ in range(400):
        env.step(agent.get_move(env).bits())
        if env.won():
            break
    assert env.won()


def test_ga_rejects_bad_configs():
    world = prl.World.from_json(world_text("right"))
    with pytest.raises(prl.PrlError):
        prl.GaPopulation(world, number_of_agents=1)
    with pytest.raises(prl.PrlError):
        prl.GaPopulation(world, number_of_steps=10, repeat_move=20)
    with pytest.raises(prl.PrlError):
        prl.GaPopulation(prl.World.from_json('{"player_position":[0.0,0.0],"objects":[]}'))


def test_nccl_all_gather_on_real_gpus():
    """Needs >= 2 visible GPUs (skipped on the 1-GPU box): tests/multi_gpu_ga_check.py under torchrun."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    from conftest import ROOT
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_ga_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_GA_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
