import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import physics_reinforcement_learning_environment_b200 as prl
w = prl.World.load("worlds/push.json")
for devices in (None, [0, 0]):
    pop = prl.GaPopulation(w, number_of_steps=2000, number_of_agents=1 << 20, repeat_move=20, mutation_rate=0.1, seed=5, devices=devices)
    t0 = time.perf_counter()
    out = [pop.step_generation() for _ in range(3)]
    dt = time.perf_counter() - t0
    print("devices", devices, "best", [float(o[0]) for o in out], "idx", [o[2] for o in out], f"{dt / 3 * 1e3:.1f} ms/gen", "env_steps", pop.env_steps(), flush=True)
    s, g = pop.read_generation()
    print("  scores finite", bool(np.isfinite(s).all()), "min", float(s.min()), "genome checksum", int(g.astype(np.uint64).sum()))
    pop.close()
