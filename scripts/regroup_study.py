#!/usr/bin/env python
"""Offline study of the rollout kernel's regrouping (csrc/prl_kernels.cuh): traces the per-step work
of a sample of the bench workload's environments with the host build of the device logic
(tests/hostsim, test infrastructure) and evaluates, under a simple cost model, how many warp
instructions different ways of dealing environments to warps would need.

  python scripts/regroup_study.py [n_blocks]     # n_blocks x 448 environments of push.json x 2000 steps
"""
import ctypes as C
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_genomes  # noqa: E402

STEPS, REPEAT = 2000, 20
WORLD = os.environ.get("PRL_STUDY_WORLD", "push")


def trace_chunk(args):
    lo, hi = args
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    hs.hs_rollout.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                              C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    hs.hs_set_trace.argtypes = [C.c_void_p, C.c_long]
    js = open(os.path.join(ROOT, "worlds", f"{WORLD}.json")).read().encode()
    genomes = make_genomes(65536, STEPS // REPEAT, 0xB200)[lo:hi]
    out = np.zeros((hi - lo, STEPS, 4), np.int32)
    sc, w, st, ov = C.c_float(), C.c_int(), C.c_int(), C.c_int()
    for i, g in enumerate(genomes):
        g = np.ascontiguousarray(g)
        hs.hs_set_trace(out[i].ctypes.data, STEPS)
        hs.hs_rollout(js, len(js), g.ctypes.data, len(g), REPEAT, STEPS, 0, None, C.byref(sc), C.byref(w), C.byref(st), C.byref(ov))
    return out


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = n_blocks * 448
    cache = f"/tmp/regroup_trace_{WORLD}_{n}.npy"
    if os.path.exists(cache):
        tr = np.load(cache)
    else:
        chunks = [(i, min(i + 112, n)) for i in range(0, n, 112)]
        with mp.Pool(os.cpu_count()) as pool:
            tr = np.concatenate(pool.map(trace_chunk, chunks))
        np.save(cache, tr)
    need, cand, full, sleeping = (tr[..., k].astype(np.int64) for k in range(4))
    print("envs", n, "mean rows", need.mean(), "mean pairs visited", cand.mean(), "recompute rate per env-step", full.mean())
    print("rows histogram", np.bincount(need.ravel(), minlength=12)[:12] / need.size)
    print("pairs histogram", np.bincount(cand.ravel(), minlength=10)[:10] / cand.size)
    print("plank asleep", ((sleeping >> 1) & 1).mean(), "player asleep", (sleeping & 1).mean())

    # cost model (warp instructions per step of a warp): every loop runs max-over-lanes trips
    def warp_cost(nd, cd, fl):  # arrays [..., 32]
        rows = nd.max(-1)
        pairs = cd.max(-1)
        anyfull = (fl.max(-1) > 0)
        return 1500 + 14 * 30 * rows + 90 * rows + 260 * pairs + 700 * anyfull

    def evaluate(order_fn, period):
        total = 0
        for b in range(n_blocks):
            sl = slice(b * 448, (b + 1) * 448)
            nd, cd, fl = need[sl], cand[sl], full[sl]
            perm = np.arange(448)
            for t0 in range(0, STEPS, period):
                if t0 > 0:
                    perm = order_fn(nd, cd, t0, perm)
                seg = slice(t0, min(t0 + period, STEPS))
                w = lambda a: a[perm][:, seg].reshape(14, 32, -1).transpose(0, 2, 1)
                total += warp_cost(w(nd), w(cd), w(fl)).sum()
        return total / (n * STEPS)

    ideal = (1500 + 14 * 30 * need + 90 * need + 260 * cand + 700 * (full > 0)).mean() / 32
    print("per-lane ideal (no divergence)", ideal)
    keep = lambda nd, cd, t0, perm: perm
    by_need = lambda nd, cd, t0, perm: np.argsort(-np.minimum(nd[:, t0 - 1], 32), kind="stable")
    def by_avg(window):
        return lambda nd, cd, t0, perm: np.argsort(-nd[:, max(0, t0 - window):t0].sum(1), kind="stable")
    def by_avg_pairs(window):
        return lambda nd, cd, t0, perm: np.argsort(-(nd[:, max(0, t0 - window):t0].sum(1) * 5 + cd[:, max(0, t0 - window):t0].sum(1) * 3), kind="stable")
    oracle_future = lambda period: (lambda nd, cd, t0, perm: np.argsort(-nd[:, t0:t0 + period].sum(1), kind="stable"))
    for name, fn, period in [("never", keep, STEPS), ("need@128 (current)", by_need, 128), ("need@64", by_need, 64), ("need@32", by_need, 32),
                             ("avg32@128", by_avg(32), 128), ("avg128@128", by_avg(128), 128), ("avg32@64", by_avg(32), 64),
                             ("avg16@32", by_avg(16), 32), ("avg8@16", by_avg(8), 16), ("need@16", by_need, 16), ("need@8", by_need, 8),
                             ("avg+pairs32@64", by_avg_pairs(32), 64),
                             ("future@128 (clairvoyant)", oracle_future(128), 128), ("future@32 (clairvoyant)", oracle_future(32), 32)]:
        c = evaluate(fn, period)
        print(f"{name:28s} warp instr per env-step {c / 32:8.1f}   lane efficiency {ideal / (c / 32):.3f}")


if __name__ == "__main__":
    main()
