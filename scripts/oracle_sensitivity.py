#!/usr/bin/env python
"""How much do the results depend on the oracle's documented deviations from rapier/parry?

The oracle (oracle/prl_oracle.c) cannot be checked against the real crates here (PARITY UNPINNED),
and it knowingly deviates from them in four places (D1 GJK vs exact closest points, D2 constraint /
contact order, D3 broad phase, D4 zero impulses). D1, D2 and D4 can be flipped at run time; this
script runs the same rollouts under each variant and reports what changes: win/lose, step of win,
score, and the free-running trajectory error. The numbers go into DESIGN.md and
profiles/r02_oracle_sensitivity.md.

  python scripts/oracle_sensitivity.py [n_genomes] > profiles/r02_oracle_sensitivity.md
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as oracle  # noqa: E402


def genomes_for(n, genome_len, seed=0xB200):
    g = np.random.Generator(np.random.Philox(seed)).integers(0, 8, size=(n, genome_len), dtype=np.uint8)
    g[:8] = np.arange(8, dtype=np.uint8)[:, None]
    return g


def trajectory(world, genome, repeat, steps, variant):
    env = oracle.OracleEnv(world, variant=variant)
    out = np.zeros((steps, env.num_bodies, 9), np.float32)
    for t in range(steps):
        env.step(int(genome[t // repeat]) if t // repeat < len(genome) else 0)
        out[t] = env.state()
    return out


def evolve(world, n, steps, repeat, generations, seed=7):
    """A plain GA on the canonical oracle (weights max + 1 - score, uniform crossover, 10 % bit
    mutation, elitism -- main.rs:86-162 in spirit, numpy RNG): a population that contains winners
    and near-winners, i.e. the genomes whose step of win a deviation could move."""
    rng = np.random.Generator(np.random.Philox(seed))
    threads = os.cpu_count() or 1
    pop = genomes_for(n, steps // repeat, seed)
    for _ in range(generations):
        score, win, _ = oracle.rollout_batch(world, pop, repeat, steps, True, n_threads=threads)
        if (win > 0).mean() > 0.4:
            break
        w = (score.max() + 1.0 - score).astype(np.float64)
        p1 = rng.choice(n, size=n, p=w / w.sum())
        p2 = rng.choice(n, size=n, p=w / w.sum())
        mask = rng.integers(0, 2, size=pop.shape, dtype=np.uint8).astype(bool)
        child = np.where(mask, pop[p1], pop[p2])
        flip = (rng.random(size=pop.shape + (3,)) < 0.05).astype(np.uint8)
        child ^= (flip[..., 0] | (flip[..., 1] << 1) | (flip[..., 2] << 2)).astype(np.uint8)
        child[0] = pop[int(np.argmin(score))]
        pop = child
    return pop


def study(name, world, n, steps, repeat=20, genomes=None):
    genomes = genomes_for(n, steps // repeat) if genomes is None else genomes
    threads = os.cpu_count() or 1
    oracle.set_default_variant(0)
    base = {e: oracle.rollout_batch(world, genomes, repeat, steps, e, n_threads=threads) for e in (True,)}
    bs, bw, _ = base[True]
    rows = []
    for label, flags in oracle.VARIANTS.items():
        oracle.set_default_variant(flags)
        s, w, _ = oracle.rollout_batch(world, genomes, repeat, steps, True, n_threads=threads)
        oracle.set_default_variant(0)
        flips = int(((bw > 0) != (w > 0)).sum())
        both = (bw > 0) & (w > 0)
        moved = int((bw[both] != w[both]).sum())
        dstep = int(np.abs(bw[both] - w[both]).max()) if both.any() else 0
        bit_same = int((s.view(np.uint32) == bs.view(np.uint32)).sum())
        close = int((np.abs(s - bs) <= 1e-3 * np.maximum(1.0, np.abs(bs))).sum())
        # free-running trajectory error of the first 16 genomes
        errs = []
        for g in genomes[:16]:
            a, b = trajectory(world, g, repeat, steps, 0), trajectory(world, g, repeat, steps, flags)
            errs.append(np.abs(a[:, :, :2] - b[:, :, :2]).max(axis=(1, 2)))
        errs = np.array(errs)  # [16, steps]
        first_div = [int(np.argmax(e > 1e-4)) if (e > 1e-4).any() else steps for e in errs]
        rows.append((label, flips, moved, dstep, bit_same, close, float(np.median(errs[:, min(99, steps - 1)])), float(np.median(errs[:, -1])),
                     int(np.median(first_div)), float(np.median(np.abs(s - bs)))))
    wins = int((bw > 0).sum())
    print(f"### {name}: {n} genomes x {steps} steps (repeat_move {repeat}), {wins} win under the canonical oracle\n")
    print("| variant | win/lose flips | winners whose step of win moves | max shift (steps) | scores bit-identical | scores within 1e-3 rel | median score change (px) | median position error at step 100 / at the end (m) | median first step with error > 1e-4 m |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r[0]} | {r[1]} / {n} | {r[2]} / {wins} | {r[3]} | {r[4]} / {n} | {r[5]} / {n} | {r[9]:.3g} | {r[6]:.1e} / {r[7]:.1e} | {r[8]} |")
    print()
    return rows


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    print("# Sensitivity of the oracle to its documented deviations from rapier2d 0.17.2 / parry2d 0.13.4\n")
    print("Generated by `scripts/oracle_sensitivity.py`. The canonical oracle (variant 0) is what the CUDA path matches bit for bit;")
    print("each row flips ONE deviation (oracle/prl_oracle.h `ORC_VARIANT_*`) and re-runs the same seeded genomes with the")
    print("reference's break-on-win rollout (main.rs:44-70). Position errors are free-running (no teacher forcing), player and blocks.\n")
    for name in ("right", "jump", "push"):
        world = json.load(open(os.path.join(ROOT, "worlds", f"{name}.json")))
        steps = 1000 if name != "push" else 2000
        study(f"{name}, random genomes", world, n, steps)
        if name != "push":
            # agents that walk right and hop now and then, 2000 steps: many reach the goals, and landing on
            # the hurdles' corners (jump.json) makes the rollouts chaotic
            rng = np.random.Generator(np.random.Philox(12))
            hop = np.where(rng.integers(0, 8, size=(min(n, 512), 100), dtype=np.uint8) < 3, 6, 2).astype(np.uint8)
            study(f"{name}, walk-right-and-hop genomes, 2000 steps", world, len(hop), 2000, genomes=hop)
        trained = evolve(world, min(n, 512), steps, 20, 60)
        study(f"{name}, a GA-trained population (winners and near-winners)", world, len(trained), steps, genomes=trained)


if __name__ == "__main__":
    main()
