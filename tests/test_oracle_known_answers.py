"""Analytic known answers the CPU oracle must satisfy. The reference has no
tests or golden vectors of its own for this path, so these are what pins the oracle."""
import math

import numpy as np

from conftest import world_dict
from oracle.binding import OracleEnv

SCALE = 0.25 / 40.0
DT = 1.0 / 60.0
PLAYER_MASS = 4 * 0.125 * 0.0625 + math.pi * 0.125 ** 2  # capsule area, density 1


def empty_world(goal_at=(0.0, -4000.0)):
    return {"player_position": [0.0, 0.0],
            "objects": [{"object": "Goal", "position": [goal_at[0], goal_at[1], 1.0], "scale": [10.0, 10.0], "rotation": 0.0}]}


def test_free_fall_is_semi_implicit_euler():
    # v_k = -2 dt k ; y_k = y_0 - 2 dt^2 k (k+1) / 2   (rapier: velocity first, then position)
    env = OracleEnv(empty_world())
    for k in range(1, 201):
        env.step(0)
        st = env.state()[0]
        assert abs(st[5] - (-2.0 * DT * k)) < 1e-4
        assert abs(st[1] - (-2.0 * DT * DT * k * (k + 1) / 2)) < 1e-4
        assert st[0] == 0.0 and st[4] == 0.0


def test_moves_do_nothing_in_the_air():
    a, b = OracleEnv(empty_world()), OracleEnv(empty_world())
    for k in range(50):
        a.step(0)
        b.step(7)
    assert np.array_equal(a.state(), b.state())


def test_distance_to_goals_and_win_latch():
    # goal 10 px square centred 100 px below the start: distance is to the rectangle, in pixels
    env = OracleEnv(empty_world(goal_at=(0.0, -100.0)))
    assert not env.won()
    assert abs(env.distance_to_goals() - 95.0) < 1e-3
    won_at = None
    for k in range(1, 400):
        env.step(0)
        y_px = env.state()[0][1] / SCALE
        if -105.0 <= y_px <= -95.0:
            assert env.distance_to_goals() == 0.0 and env.won()
            won_at = won_at or k
        if won_at:
            assert env.won()  # sticky (common.rs:333-339)
    assert won_at is not None
    assert OracleEnv({"player_position": [0, 0], "objects": []}).distance_to_goals() is None


def test_rotated_goal_uses_rotated_axes():
    w = empty_world(goal_at=(100.0, 0.0))
    w["objects"][0]["scale"] = [40.0, 10.0]
    d0 = OracleEnv(w).distance_to_goals()
    w["objects"][0]["rotation"] = math.pi / 2
    d90 = OracleEnv(w).distance_to_goals()
    assert abs(d0 - 80.0) < 1e-3 and abs(d90 - 95.0) < 1e-3


def test_first_step_is_never_grounded_then_jump_impulse():
    # right.json starts ~1e-4 above the floor; contacts exist only after the first pipeline step
    env = OracleEnv(world_dict("right"))
    env.step(4)
    assert env.state()[0][5] < 0.0  # no jump on step 1 (common.rs:224 reads the previous step's contacts)
    vy = []
    for _ in range(30):
        env.step(4)
        vy.append(env.state()[0][5])
    # jump: dv = 0.1 / m = 1.2448 m/s (minus one step of gravity and solver softness)
    assert abs(max(vy) - 0.1 / PLAYER_MASS) < 0.08


def test_jump_apex():
    env = OracleEnv(world_dict("right"))
    y0 = env.state()[0][1]
    ys = []
    for _ in range(120):
        env.step(4 if len(ys) < 3 else 0)
        ys.append(env.state()[0][1])
    apex = max(ys) - y0
    assert 0.33 < apex < 0.42  # v^2 / (2 g) = 1.2448^2 / 4 = 0.387 m


def test_walk_right_wins_right_json():
    env = OracleEnv(world_dict("right"))
    score, win, run = env.rollout(np.full(50, 2, np.uint8), 20, 1000)
    assert score == 0.0 and 170 <= win <= 185 and run == win
    # left walks off the edge and falls for ever; idle never wins
    for mv in (0, 1):
        score, win, run = env.rollout(np.full(50, mv, np.uint8), 20, 1000)
        assert win == -1 and run == 1000


def test_resting_player_stays_put_and_sleeps():
    env = OracleEnv(world_dict("jump"))
    for _ in range(200):
        env.step(0)
    st = env.state()[0]
    assert st[8] == 1.0 and st[4] == 0.0 and st[5] == 0.0  # asleep after 2 s below the energy threshold
    y = st[1]
    for _ in range(100):
        env.step(0)
    assert env.state()[0][1] == y


def test_push_plank_settles_and_sleeps():
    env = OracleEnv(world_dict("push"))
    for _ in range(400):
        env.step(0)
    st = env.state()
    assert st[1][8] == 1.0 and abs(st[1][3]) < 1e-2  # plank asleep, essentially level
    assert not np.isnan(st).any()


def test_early_exit_tail_quirk():
    # number_of_steps % repeat_move tail runs one extra step after a win (main.rs:61-68)
    env = OracleEnv(world_dict("right"))
    score, win, run = env.rollout(np.full(50, 2, np.uint8), 20, 1007)
    assert run == win + 1


def test_walking_accelerates_by_impulse_minus_coulomb_friction_and_slides_to_a_stop():
    """On the flat floor of right.json every grounded step with Move{right} adds 0.003 / m to vx
    (common.rs:289) and the contact's friction row takes away at most mu * g * dt (mu = 0.5 from
    rapier's default collider friction, Average rule; g = 2): net +0.003/m - 1/60 per step. With
    the keys released the player decelerates by exactly mu * g * dt per step until it stops."""
    env = OracleEnv(world_dict("right"))
    vx = []
    for _ in range(60):
        env.step(2)
        vx.append(float(env.state()[0][4]))
    gain = np.diff(vx[5:])
    assert np.allclose(gain, 0.003 / PLAYER_MASS - 0.5 * 2.0 * DT, atol=2e-5)
    vx = []
    for _ in range(40):
        env.step(0)
        vx.append(float(env.state()[0][4]))
    assert np.allclose(np.diff(vx), -0.5 * 2.0 * DT, atol=1e-6)
    for _ in range(100):
        env.step(0)
    st = env.state()[0]
    assert abs(st[4]) < 1e-5 and abs(st[5]) < 1e-3   # at rest on the floor


def test_resting_penetration_is_the_allowed_linear_error():
    """rapier only pushes a penetration out down to allowed_linear_error (1 mm; the bias
    erp / dt * clamp(dist + 0.001, ., 0) vanishes there) and the soft constraint sags a little
    under gravity: a player at rest on the floor sits 1.0-1.3 mm inside it."""
    world = world_dict("right")
    env = OracleEnv(world)
    for _ in range(300):
        env.step(0)
    floor = [o for o in world["objects"] if o["object"] != "Goal"][0]
    top = floor["position"][1] * SCALE + abs(floor["scale"][1]) * SCALE / 2
    bottom = float(env.state()[0][1]) - (0.0625 + 0.125)
    assert 0.0010 <= top - bottom <= 0.0013


def test_pushable_block_in_free_fall():
    """A dynamic cuboid with nothing under it: same semi-implicit Euler as the player, no spin, its
    rotation (kept as a unit complex, re-normalised every step) unchanged to rounding."""
    rot = 0.3
    world = empty_world()
    world["objects"].append({"object": {"Block": {"fixed": False}}, "position": [200.0, 50.0, 2.0], "scale": [60.0, 20.0], "rotation": rot})
    env = OracleEnv(world)
    y0 = 50.0 * SCALE
    for k in range(1, 121):
        env.step(0)
    blk = env.state()[1]
    k = 120
    assert abs(blk[0] - 200.0 * SCALE) < 1e-6 and abs(blk[4]) == 0.0 and blk[6] == 0.0
    assert abs(blk[5] - (-2.0 * DT * k)) < 1e-4
    assert abs(blk[1] - (y0 - 2.0 * DT * DT * k * (k + 1) / 2)) < 1e-4
    assert abs(blk[2] - math.cos(rot)) < 1e-5 and abs(blk[3] - math.sin(rot)) < 1e-5


def test_momentum_of_player_and_free_block_changes_only_by_gravity():
    """The player stands on a pushable block and both fall freely (no fixed collider anywhere).
    Contact impulses act equal and opposite on the two bodies, and so do the move impulses
    (common.rs:268, 292, 305-306: the reaction -J is applied to the block at the contact point):
    whatever the moves -- including the jump that separates the two -- the total linear momentum
    changes by exactly (m_player + m_block) * g * dt per step."""
    world = empty_world()
    world["objects"].append({"object": {"Block": {"fixed": False}}, "position": [5.0, -40.0, 2.0], "scale": [120.0, 20.0], "rotation": 0.0})
    m_block = (120.0 * SCALE) * (20.0 * SCALE)
    env = OracleEnv(world)
    rng = np.random.default_rng(0)
    prev, jumped = None, False
    for t in range(80):
        env.step(int(rng.integers(0, 8)) if t > 5 else 0)
        st = env.state()
        p = np.array([PLAYER_MASS * st[0][4] + m_block * st[1][4], PLAYER_MASS * st[0][5] + m_block * st[1][5]], np.float64)
        if prev is not None:
            assert abs(p[0] - prev[0]) < 1e-7 and abs((p[1] - prev[1]) + (PLAYER_MASS + m_block) * 2.0 * DT) < 1e-7, t
        jumped = jumped or st[0][5] - st[1][5] > 0.5   # the jump impulse 0.1 / m really separated them
        prev = p
    assert jumped
