#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference (Rust + rapier2d 0.17.2 from crates.io) cannot be built or run in this image and
ships no golden vectors, so these fixtures are produced by the CPU ORACLE
(oracle/prl_oracle.c) -- they pin the oracle and the CUDA path against regressions and against
each other, not against rapier itself: PARITY UNPINNED. They are REGRESSION PINS (`source =
"oracle"`). Fixtures from the real crate are produced by rust/examples/dump_golden.rs +
scripts/import_rapier_golden.py (`source = "rapier"`, `--request` below writes their case list)
and are compared in tolerance mode by tests/test_golden.py.

Format, per world: genomes (16 x 50 uint8), repeat_move, number_of_steps, score (f32), step_of_win
(i32), steps_run (i32) with the reference's early exit, and the full body trajectory of the first
4 genomes without early exit: traj[g, step, body, 9].
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.binding import OracleEnv  # noqa: E402
from conftest import edge_case_worlds, many_supports_world  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    if "--request" in sys.argv:
        # the cases rust/examples/dump_golden.rs runs on the REAL reference (see scripts/import_rapier_golden.py)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from import_rapier_golden import request_cases
        json.dump({"cases": request_cases()}, open(os.path.join(out_dir, "request.json"), "w"))
        print("wrote", os.path.join(out_dir, "request.json"))
        return
    worlds = {name: json.load(open(os.path.join(ROOT, "worlds", f"{name}.json"))) for name in ("right", "jump", "push")}
    # levels the editor can produce that the shipped examples do not cover; their JSON travels in the fixture
    worlds.update({f"edge_{k}": v for k, v in edge_case_worlds().items()})
    worlds["edge_many_supports"] = many_supports_world(6)
    for name, world in worlds.items():
        rng = np.random.Generator(np.random.Philox(0xB200))
        genomes = rng.integers(0, 8, size=(16, 50), dtype=np.uint8)
        genomes[:8] = np.arange(8, dtype=np.uint8)[:, None]  # the 8 constant moves of README.md:96-114
        repeat, steps = 20, (1000 if not name.startswith("edge_") else 400)
        env = OracleEnv(world)
        score = np.zeros(16, np.float32); win = np.zeros(16, np.int32); run = np.zeros(16, np.int32)
        for i in range(16):
            score[i], win[i], run[i] = env.rollout(genomes[i], repeat, steps, early_exit=True)
        traj = np.zeros((4, steps, env.num_bodies, 9), np.float32)
        for gi, g in enumerate((2, 6, 8, 9)):
            env.reset()
            for t in range(steps):
                env.step(int(genomes[g][t // repeat]))
                traj[gi, t] = env.state()
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), source=np.array("oracle"), world_json=np.array(json.dumps(world)), genomes=genomes, repeat_move=repeat,
                            number_of_steps=steps, score=score, step_of_win=win, steps_run=run,
                            traj_genomes=np.array((2, 6, 8, 9)), traj=traj[:, ::10].copy())
        print(name, "scores", score[:8], "wins", win)


if __name__ == "__main__":
    main()
