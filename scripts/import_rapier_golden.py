#!/usr/bin/env python
"""tests/golden/rapier_dump.json (written by rust/examples/dump_golden.rs from the REAL reference,
rapier2d 0.17.2) -> tests/golden/rapier_<case>.npz with source = "rapier", the fixtures
tests/test_golden.py consumes in tolerance mode.

  python scripts/make_golden.py --request      # tests/golden/request.json: the cases to dump
  (on a machine with cargo)  cd rust && cargo run --release --features reference --example dump_golden -- ../tests/golden
  python scripts/import_rapier_golden.py [dump.json] [out_dir]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def request_cases():
    """The cases of tests/golden/request.json: the three shipped levels x (the 8 constant moves of
    README.md:96-114 + 16 seeded random genomes) x 2000 steps."""
    cases = []
    for name in ("right", "jump", "push"):
        world = json.load(open(os.path.join(ROOT, "worlds", f"{name}.json")))
        rng = np.random.Generator(np.random.Philox(0xB200))
        genomes = rng.integers(0, 8, size=(24, 100), dtype=np.uint8)
        genomes[:8] = np.arange(8, dtype=np.uint8)[:, None]
        cases.append({"name": name, "world": world, "genomes": genomes.tolist(), "repeat_move": 20, "number_of_steps": 2000})
    return cases


def dump_to_npz(case, dump, path, source):
    """One case of the request + its dump (the JSON object dump_golden.rs writes) -> npz."""
    f32 = lambda bits: np.array(bits, dtype=np.uint32).view(np.float32)
    nb = int(dump["n_bodies"])
    state = f32(dump["state_bits"]).reshape(len(case["genomes"]), case["number_of_steps"], nb, 9)
    np.savez_compressed(
        path, source=np.array(source), world_json=np.array(json.dumps(case["world"])), genomes=np.array(case["genomes"], dtype=np.uint8),
        repeat_move=case["repeat_move"], number_of_steps=case["number_of_steps"], score=f32(dump["score_bits"]),
        step_of_win=np.array(dump["step_of_win"], np.int32), steps_run=np.array(dump["steps_run"], np.int32),
        traj=state, distance=f32(dump["distance_bits"]).reshape(len(case["genomes"]), -1), won=np.array(dump["won"], np.uint8))


def main():
    dump_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "rapier_dump.json")
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "tests", "golden")
    cases = {c["name"]: c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "request.json")))["cases"]}
    for d in json.load(open(dump_path)):
        out = os.path.join(out_dir, f"rapier_{d['name']}.npz")
        dump_to_npz(cases[d["name"]], d, out, "rapier")
        print("wrote", out)


if __name__ == "__main__":
    main()
