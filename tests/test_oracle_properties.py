"""Property tests of the oracle (hypothesis): invariants that hold for any level the editor can produce,
whatever the arithmetic details of rapier are: finite state, non-negative goal distance, a
sticky win latch, sleeping bodies at rest, no tunnelling through a wide floor at the speeds this game
reaches, and determinism (two environments of one world stepped with the same moves stay bit-identical)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import random_genomes, random_world, random_world_many_pushables
from oracle.binding import OracleEnv


def floor_top(world):
    """y (pixels) of the top face at x = 0 of the widest, nearly level fixed block (random_world's floor)"""
    floors = [o for o in world["objects"] if o["object"] != "Goal" and o["object"]["Block"]["fixed"] and abs(o["scale"][0]) >= 1500]
    f = floors[0]
    return f["position"][1] + abs(f["scale"][1]) / 2.0, f


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10**6), many=st.booleans())
def test_invariants_on_random_levels(seed, many):
    world = random_world_many_pushables(seed) if many else random_world(seed)
    genome = random_genomes(1, 20, seed % 9973)[0]
    a, b = OracleEnv(world), OracleEnv(world)
    top, floor = floor_top(world)
    scale = 0.25 / 40.0
    won = False
    for t in range(300):
        mv = int(genome[t // 15])
        a.step(mv)
        b.step(mv)
        s = a.state()
        assert np.isfinite(s).all()
        assert np.array_equal(s.view(np.uint32), b.state().view(np.uint32))          # deterministic
        d = a.distance_to_goals()
        assert d is not None and d >= 0.0 and np.isfinite(d)
        assert (d < 1e-7) <= a.won()                                                  # inside a goal => won
        won = won or a.won()
        assert a.won() == won                                                         # the latch is sticky
        asleep = s[:, 8] > 0
        assert np.all(s[asleep, 4:7] == 0.0)                                          # sleeping bodies are at rest
        assert np.allclose(s[:, 2] ** 2 + s[:, 3] ** 2, 1.0, atol=1e-4)               # rotations stay unit complex
        # while above the floor's span, the player's capsule bottom never sinks through its top face
        # (speeds stay below ~3 m/s = 0.05 m per step, far less than the 0.31 m floor)
        px, py = s[0, 0] / scale, s[0, 1] / scale
        half = abs(floor["scale"][0]) / 2.0 - 40.0
        if abs(px - floor["position"][0]) < half:
            rot = floor["rotation"]
            top_here = floor["position"][1] + (px - floor["position"][0]) * np.tan(rot) + abs(floor["scale"][1]) / 2.0 / np.cos(rot)
            assert py - 30.0 > top_here - 12.0, (t, px, py, top_here)
