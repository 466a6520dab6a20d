"""The oracle's documented deviations from rapier/parry (oracle/prl_oracle.c D1, D2, D4) flipped one
at a time: the canonical variant must be untouched by the machinery; on right.json nothing
observable changes; on jump.json and push.json the measured sensitivity is asserted as recorded.
Full numbers: scripts/oracle_sensitivity.py -> profiles/r02_oracle_sensitivity.md."""
import os

import numpy as np
import pytest

from conftest import random_genomes, world_dict
from oracle import binding as oracle


def rollouts(world, genomes, steps, variant):
    oracle.set_default_variant(variant)
    try:
        return oracle.rollout_batch(world, genomes, 20, steps, True, n_threads=os.cpu_count())
    finally:
        oracle.set_default_variant(0)


@pytest.mark.parametrize("name", ["right", "jump", "push"])
def test_variant_zero_is_the_canonical_oracle(name):
    g = random_genomes(64, 50, 3)
    a = oracle.rollout_batch(world_dict(name), g, 20, 1000, True, n_threads=2)
    b = rollouts(world_dict(name), g, 1000, 0)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


def hopping_genomes():
    g = random_genomes(256, 100, 11)
    g[:8] = np.arange(8, dtype=np.uint8)[:, None]
    g[8:136] = np.where(random_genomes(128, 100, 12) < 3, 6, 2)  # walk right, hop now and then: many of these reach the goals
    return g


@pytest.mark.parametrize("label", sorted(oracle.VARIANTS))
def test_deviations_change_nothing_observable_on_right_json(label):
    """One fixed collider: nothing for the order to permute, and GJK agrees with the exact closest
    points to rounding. 2000-step rollouts, 138 of 256 genomes win; who wins and when is identical."""
    world = world_dict("right")
    g = hopping_genomes()
    s0, w0, _ = rollouts(world, g, 2000, 0)
    s1, w1, _ = rollouts(world, g, 2000, oracle.VARIANTS[label])
    assert (w0 > 0).sum() >= 100
    assert np.array_equal(w0, w1)
    assert np.abs(s0 - s1).max() < 0.1  # pixels (scores of non-winners: 0.02 px at most)


@pytest.mark.parametrize("label", sorted(oracle.VARIANTS))
def test_hopping_over_hurdles_is_chaotic_under_every_deviation_but_d4(label):
    """jump.json, agents that hop over the hurdles for 2000 steps: a capsule landing on a box corner
    amplifies rounding-level differences, so flipping D1 (GJK) or D2 (order) changes who wins for
    about a tenth of such genomes and moves steps of win by hundreds of steps -- measured here and
    recorded in DESIGN.md: no restatement that is not bit-identical to rapier (rapier itself on
    another libm included) can promise the same step of win on these rollouts. D4 changes nothing.
    Short or walking rollouts are unaffected (previous test, scripts/oracle_sensitivity.py)."""
    world = world_dict("jump")
    g = hopping_genomes()
    s0, w0, _ = rollouts(world, g, 2000, 0)
    s1, w1, _ = rollouts(world, g, 2000, oracle.VARIANTS[label])
    flips = int(((w0 > 0) != (w1 > 0)).sum())
    assert (w0 > 0).sum() >= 40
    if oracle.VARIANTS[label] == oracle.VARIANT_WAKE_ON_ZERO_IMPULSE:
        assert flips == 0 and np.array_equal(w0, w1)
    else:
        assert flips <= 64, flips          # recorded: 8 .. 34 of 256
        assert np.median(np.abs(s0 - s1)) < 0.01  # pixels: the typical rollout barely moves
    # the first second of every rollout is unaffected by any deviation
    e0 = [oracle.OracleEnv(world) for _ in range(8)]
    e1 = [oracle.OracleEnv(world, variant=oracle.VARIANTS[label]) for _ in range(8)]
    for t in range(60):
        for i in range(8):
            e0[i].step(int(g[8 + i][t // 20])); e1[i].step(int(g[8 + i][t // 20]))
    assert max(float(np.abs(a.state()[:, :2] - b.state()[:, :2]).max()) for a, b in zip(e0, e1)) < 1e-4


def test_push_is_sensitive_to_the_gauss_seidel_order_but_not_to_zero_impulses():
    """push.json couples the plank's contacts through one body, so the constraint order (D2) changes
    trajectories at the solver-tolerance level (recorded, not hidden); waking on a zero impulse (D4)
    changes nothing."""
    world = world_dict("push")
    g = random_genomes(128, 100, 13)
    s0, w0, _ = rollouts(world, g, 2000, 0)
    s4, w4, _ = rollouts(world, g, 2000, oracle.VARIANT_WAKE_ON_ZERO_IMPULSE)
    assert np.array_equal(s0.view(np.uint32), s4.view(np.uint32)) and np.array_equal(w0, w4)
    s2, w2, _ = rollouts(world, g, 2000, oracle.VARIANT_CREATION_ORDER)
    assert np.array_equal(w0 > 0, w2 > 0)
    assert 0.0 < np.median(np.abs(s0 - s2)) < 2.0  # pixels: small but not zero
