"""Committed golden vectors (tests/golden/*.npz) against the oracle (CPU) and the CUDA path (GPU).

Two kinds of fixture, told apart by their `source` field:

* `source = "oracle"` -- REGRESSION PINS made by scripts/make_golden.py from the CPU oracle. The
  reference ships no golden vectors and cannot be run in this image (PARITY UNPINNED), so these pin
  the oracle and the CUDA path against regressions and against each other, bit for bit; they say
  nothing about rapier.
* `source = "rapier"` (or any other source) -- dumps of the REAL reference made by
  rust/examples/dump_golden.rs on a machine with cargo and imported by
  scripts/import_rapier_golden.py. rapier's GJK and contact order make such a dump ulp-different
  from the oracle, so it is compared in TOLERANCE MODE (`compare_with_dump`): exact `won()` /
  step of win / steps run for every genome (the north star's bar), a teacher-forced one-step error
  bound, and a free-running error bound over the first steps. None is committed yet; the mode is
  exercised on dumps made from the oracle's own deviation variants (oracle/prl_oracle.h
  ORC_VARIANT_*), which differ from the canonical oracle the way a rapier dump is expected to.
"""
import glob
import json
import os
import sys

import numpy as np
import pytest

import physics_reinforcement_learning_environment_b200 as prl
from conftest import ROOT, WORLDS, world_dict, world_text
from oracle import binding as oracle
from oracle.binding import OracleEnv

sys.path.insert(0, os.path.join(ROOT, "scripts"))
from import_rapier_golden import dump_to_npz, request_cases  # noqa: E402

ALL = sorted(os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))


def source_of(g):
    return str(g["source"]) if "source" in g.files else "oracle"


PINS = [n for n in ALL if source_of(golden(n)) == "oracle"]
DUMPS = [n for n in ALL if source_of(golden(n)) != "oracle"]

# ---- tolerance mode. Thresholds (f32, metres and metres/second; the worlds are a few metres across,
# speeds are below 2 m/s): what a correct restatement with a different-but-valid GJK / summation
# order stays within, measured on the oracle's own variants (profiles/r02_oracle_sensitivity.md).
TF_POS_TOL = 2.0e-5      # teacher-forced: position / rotation (cos, sin) after ONE step from the dump's state
TF_VEL_TOL = 2.0e-3      # teacher-forced: linear / angular velocity after one step
FREE_STEPS = 60          # free-running: the first second of simulation ...
FREE_POS_TOL = 1.0e-3    # ... stays within a millimetre (contact dynamics are chaotic beyond that)
TF_QUANTILE = 0.99       # share of teacher-forced steps that must be within tolerance: a step on
                         # which a contact manifold is rebuilt differently (persistence vs recompute)
                         # legitimately jumps; the rest must agree


def golden_world(name, g):
    """(dict, text): the shipped levels by name, the others from the JSON stored in the fixture"""
    if name in WORLDS:
        return world_dict(name), world_text(name)
    text = str(g["world_json"])
    return json.loads(text), text


def compare_with_dump(g, make_env):
    """Tolerance-mode comparison of a full-trajectory dump (fields of import_rapier_golden.dump_to_npz)
    with environments made by make_env() (API of OracleEnv). Returns a dict of measured errors;
    raises AssertionError on the first violated bound."""
    world = json.loads(str(g["world_json"]))
    genomes, repeat, steps = g["genomes"], int(g["repeat_move"]), int(g["number_of_steps"])
    traj = g["traj"]  # [genome, step, body, 9]
    report = {"tf_pos": 0.0, "tf_vel": 0.0, "free_pos": 0.0, "tf_outliers": 0, "tf_steps": 0}
    for i, genome in enumerate(genomes):
        env = make_env(world)
        sc, win, run = env.rollout(genome, repeat, steps, True)
        assert win == int(g["step_of_win"][i]), f"genome {i}: step of win {win} != dump {int(g['step_of_win'][i])}"
        assert run == int(g["steps_run"][i]), f"genome {i}: steps run {run} != dump {int(g['steps_run'][i])}"
        assert abs(float(sc) - float(g["score"][i])) <= 1e-3 * max(1.0, abs(float(g["score"][i]))), f"genome {i}: score"
        move = lambda t: int(genome[t // repeat]) if t // repeat < len(genome) else 0
        # free-running prefix
        env.reset()
        for t in range(min(FREE_STEPS, steps)):
            env.step(move(t))
            err = float(np.abs(env.state()[:, :4] - traj[i, t, :, :4]).max())
            report["free_pos"] = max(report["free_pos"], err)
            assert err <= FREE_POS_TOL, f"genome {i} step {t + 1}: free-running position error {err:.3e}"
            assert env.won() == bool(g["won"][i, t])
        # teacher-forced: from the dump's state at step t, one step must land on the dump's state at t+1
        # (the contact manifolds are the environment's own: a dump does not carry them)
        env.reset()
        env.step(move(0))
        for t in range(1, min(steps, 400)):
            for b in range(traj.shape[2]):
                env.set_body_state(b, traj[i, t - 1, b])
            env.step(move(t))
            st = env.state()
            ep = float(np.abs(st[:, :4] - traj[i, t, :, :4]).max())
            ev = float(np.abs(st[:, 4:7] - traj[i, t, :, 4:7]).max())
            report["tf_steps"] += 1
            if ep > TF_POS_TOL or ev > TF_VEL_TOL:
                report["tf_outliers"] += 1
            else:
                report["tf_pos"], report["tf_vel"] = max(report["tf_pos"], ep), max(report["tf_vel"], ev)
    assert report["tf_outliers"] <= (1.0 - TF_QUANTILE) * report["tf_steps"], report
    return report


@pytest.mark.parametrize("name", PINS)
def test_oracle_reproduces_its_regression_pins(name):
    g = golden(name)
    env = OracleEnv(golden_world(name, g)[0])
    for i, genome in enumerate(g["genomes"]):
        sc, win, run = env.rollout(genome, int(g["repeat_move"]), int(g["number_of_steps"]), True)
        assert sc.view(np.uint32) == g["score"][i].view(np.uint32) and win == g["step_of_win"][i] and run == g["steps_run"][i]


@pytest.mark.gpu
@pytest.mark.parametrize("name", PINS)
def test_cuda_reproduces_the_regression_pins(name):
    g = golden(name)
    repeat, steps = int(g["repeat_move"]), int(g["number_of_steps"])
    text = golden_world(name, g)[1]
    env = prl.BatchedEnvironment(prl.World.from_json(text), 16)
    score, win = env.rollout(np.ascontiguousarray(g["genomes"]), repeat, steps, early_exit=True)
    _, _, run = env.rollout_info()
    assert np.array_equal(score.view(np.uint32), g["score"].view(np.uint32))
    assert np.array_equal(win, g["step_of_win"]) and np.array_equal(run, g["steps_run"])
    # sampled trajectories (every 10th step), stepping one launch per step through prl_batch_step
    tg = g["traj_genomes"]
    env4 = prl.BatchedEnvironment(prl.World.from_json(text), len(tg))
    for t in range(steps):
        env4.step(np.ascontiguousarray(g["genomes"][tg, t // repeat]))
        if t % 10 == 0:
            assert np.array_equal(env4.state().view(np.uint32), g["traj"][:, t // 10].view(np.uint32)), t


@pytest.mark.skipif(not DUMPS, reason="no dump of the real reference is committed: run rust/examples/dump_golden.rs on a machine with "
                                      "cargo, then scripts/import_rapier_golden.py (tests/golden/request.json lists the cases)")
@pytest.mark.parametrize("name", DUMPS or ["none"])
def test_oracle_matches_the_reference_dump_within_tolerance(name):
    compare_with_dump(golden(name), lambda world: OracleEnv(world))


def _oracle_dump(case, variant):
    """What rust/examples/dump_golden.rs writes for one case, produced by an oracle variant."""
    bits = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).tolist()
    d = {"name": case["name"], "score_bits": [], "step_of_win": [], "steps_run": [], "state_bits": [], "distance_bits": [], "won": []}
    for genome in case["genomes"]:
        env = OracleEnv(case["world"], variant=variant)
        sc, win, run = env.rollout(np.array(genome, np.uint8), case["repeat_move"], case["number_of_steps"], True)
        d["score_bits"] += bits([sc]); d["step_of_win"].append(win); d["steps_run"].append(run)
        env.reset()
        states, dists, wons = [], [], []
        for t in range(case["number_of_steps"]):
            g = t // case["repeat_move"]
            env.step(genome[g] if g < len(genome) else 0)
            states.append(bits(env.state().reshape(-1))); dists.append(bits([env.distance_to_goals()])[0]); wons.append(int(env.won()))
        d["state_bits"].append(states); d["distance_bits"].append(dists); d["won"].append(wons)
        d["n_bodies"] = env.num_bodies
    return d


@pytest.mark.parametrize("name", ["right", "jump"])
def test_tolerance_mode_on_dumps_of_the_oracles_own_variants(name, tmp_path):
    """End to end through the importer: a dump made with parry-style GJK closest points, the contact
    graph's edge order and waking zero impulses (the three deviations that can be flipped) must pass
    tolerance mode against the canonical oracle; a dump of a WRONG restatement (gravity acting for
    one extra step, emulated by shifting the trajectory by one step) must not."""
    case = dict(next(c for c in request_cases() if c["name"] == name))
    # 8 genomes (the constant moves 2 = right and 6 = right + up among them) x 400 steps
    case["genomes"], case["number_of_steps"] = [g[:20] for g in case["genomes"][2:10]], 400
    variant = oracle.VARIANT_GJK | oracle.VARIANT_CREATION_ORDER | oracle.VARIANT_WAKE_ON_ZERO_IMPULSE
    dump = _oracle_dump(case, variant)
    path = str(tmp_path / "variant.npz")
    dump_to_npz(case, dump, path, "oracle-variant")
    g = np.load(path)
    assert source_of(g) == "oracle-variant" and g["traj"].shape[:2] == (8, 400)
    report = compare_with_dump(g, lambda world: OracleEnv(world))
    assert report["tf_steps"] > 3000 and report["tf_pos"] <= TF_POS_TOL
    # the variant is not bit-identical to the canonical oracle (otherwise this test shows nothing)
    canon = _oracle_dump(case, 0)
    assert canon["state_bits"] != dump["state_bits"] or name == "right"
    wrong = dict(dump)
    wrong["state_bits"] = [s[1:] + s[-1:] for s in dump["state_bits"]]
    dump_to_npz(case, wrong, path, "wrong")
    with pytest.raises(AssertionError):
        compare_with_dump(np.load(path), lambda world: OracleEnv(world))
