#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched `Environment::step` hot path on push.json.

One "step" of this benchmark is one pass of the hot path over one batch: the fused GA-fitness
rollout (reference src/main.rs:44-70 over src/common.rs:215-340) of `--envs` environments for
`--horizon` physics steps, early exit on win DISABLED so the work is input independent.
Workload = BASELINE.json configs[2]: example_worlds/push.json, 65536 envs x 2000 steps per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference's
                                                           # step on the host cores (oracle/)
Under torchrun (N > 1) every rank owns its own slice of the population (weak scaling, no
data-path collective: environments are independent); time is the max over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def workload_config(args, genome_len):
    """`config` of the JSON line: identical for the CUDA arm and the --impl reference arm."""
    return {"workload": f"{args.world}.json, {args.envs} envs x {args.horizon} steps per GPU, genome_len {genome_len}, "
                        f"repeat_move {args.repeat}, early exit off",
            "l2": "CUDA arm: L2 flushed between timed passes (256 MB fill outside the event brackets); CPU arm: not applicable"}


def kernel_source_hash():
    """sha256 over the sources the rollout kernel is compiled from (device code, world tables, libm,
    build flags): keys the committed ncu figures (profiles/rollout_profile.json) to the kernel they
    were measured on."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "physics_reinforcement_learning_environment_b200", "csrc")
    for f in ("prl_device.cuh", "prl_kernels.cuh", "prl_world.h", "prl_libm.h", "Makefile"):
        h.update(f.encode())
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def algorithmic_bytes(n_bodies):
    """algorithmic HBM bytes per env-step: cross-step body state read once + written once"""
    return 2 * (28 + 32 * (n_bodies - 1))


def make_genomes(n_envs, genome_len, seed):
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.integers(0, 8, size=(n_envs, genome_len), dtype=np.uint8)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_rate(world_text, n_envs, genome_len, repeat, horizon, threads, seed):
    """env-steps/s of the oracle (CPU restatement of the reference step) on `threads` host threads."""
    from oracle import binding as oracle
    genomes = make_genomes(n_envs, genome_len, seed)
    t0 = time.perf_counter()
    _, _, steps = oracle.rollout_batch(json.loads(world_text), genomes, repeat, horizon, early_exit=False, n_threads=threads)
    dt = time.perf_counter() - t0
    return steps / dt, steps, dt


def run_reference(args, world_text, rank):
    """--impl reference: the CPU path (oracle port; the Rust reference cannot be built in this image)."""
    if rank != 0:
        return
    from oracle import binding as oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    genome_len = args.horizon // args.repeat
    sample_envs = max(cores * 256, 1024)  # ~1 s of host work per step
    for _ in range(args.warmup):
        cpu_reference_rate(world_text, sample_envs, genome_len, args.repeat, args.horizon, cores, 1)
    tot_steps, tot_time = 0, 0.0
    for k in range(args.steps):
        _, s, dt = cpu_reference_rate(world_text, sample_envs, genome_len, args.repeat, args.horizon, cores, 100 + k)
        tot_steps += s
        tot_time += dt
    value = tot_steps / tot_time
    sample = f"{sample_envs} envs x {args.horizon} steps of {args.world}.json per step (bounded sample of the {args.envs}-env workload), {cores} threads"
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, genome_len),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--world", default="push", help="push | jump | right | synth64 (worlds/*.json)")
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--horizon", type=int, default=2000, help="physics steps per rollout")
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C2 / C4 records")
    ap.add_argument("--no-ga", action="store_true", help="skip the GA (C5-shape) record")
    ap.add_argument("--ga-agents", type=int, default=131072, help="GA agents per GPU")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    world_text = open(os.path.join(ROOT, "worlds", f"{args.world}.json")).read()

    if args.impl == "reference":
        run_reference(args, world_text, rank)
        return

    import torch
    import torch.distributed as dist
    import physics_reinforcement_learning_environment_b200 as prl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)

    world = prl.World.from_json(world_text)
    n_bodies = world.counts()[0]
    genome_len = args.horizon // args.repeat
    stream = torch.cuda.current_stream()
    env = prl.BatchedEnvironment(world, args.envs, device=local_rank, stream=stream.cuda_stream)

    # inputs: synthetic random genomes (Philox, seed 0xB200 + rank), resident in HBM for `value`,
    # in pinned host memory for `e2e`
    host_genomes = torch.from_numpy(make_genomes(args.envs, genome_len, 0xB200 + rank)).pin_memory()
    dev_genomes = host_genomes.to(dev)
    dev_score = torch.empty(args.envs, dtype=torch.float32, device=dev)
    dev_win = torch.empty(args.envs, dtype=torch.int32, device=dev)
    host_score = torch.empty(args.envs, dtype=torch.float32).pin_memory()
    host_win = torch.empty(args.envs, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_pass():
        env.rollout(dev_genomes, args.repeat, args.horizon, early_exit=False, out_score=dev_score, out_step_of_win=dev_win)

    def one_pass_e2e():
        env.rollout(host_genomes.numpy(), args.repeat, args.horizon, early_exit=False, out_score=host_score.numpy(),
                    out_step_of_win=host_win.numpy())

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        flush.fill_(1)
        one_pass()
    barrier()

    # ---- timed region: K passes, each bracketed by CUDA events on the launching stream; L2 is
    # flushed (256 MB memset) between passes, outside the event brackets
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = env.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        starts[k].record(stream)
        one_pass()
        ends[k].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_launches = env.launch_count - launches0
    clocks = sampler.stop()
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))

    # ---- end to end through the public API with HOST buffers: H2D of the genomes and D2H of the
    # scores / win steps inside the timed region (wall clock around synchronous calls)
    one_pass_e2e()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_pass_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0

    env.sync()  # raises PRL_ERR_OVERFLOW if an environment ran out of contact-manifold slots in the timed passes

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    env_steps_per_pass = args.envs * args.horizon
    total_steps = env_steps_per_pass * args.steps * world_size
    value = total_steps / (dev_ms * 1e-3)
    e2e_value = total_steps / (e2e_ms * 1e-3)
    del env, flush

    secondary = None if args.no_secondary else secondary_records(prl, torch, dist, dev, local_rank, world_size, stream)
    ga = None if args.no_ga else ga_record(prl, torch, dist, dev, local_rank, rank, world_size, args)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        balg = algorithmic_bytes(n_bodies)
        launch_s = dev_ms * 1e-3 / args.steps
        achieved = balg * env_steps_per_pass / launch_s / 1e9
        # ncu figures of the dominant kernel come from the committed capture of THIS kernel source
        # (profiles/rollout_profile.json, keyed by the source hash); a stale capture is not reported
        traffic, issue, src_hash = None, None, kernel_source_hash()
        ppath = os.path.join(ROOT, "profiles", "rollout_profile.json")
        if os.path.exists(ppath):
            try:
                pj = json.load(open(ppath))
                same = pj.get("envs") == args.envs and pj.get("horizon") == args.horizon and pj.get("world") == args.world
                if same and pj.get("source_sha256") == src_hash:
                    traffic = pj.get("dram_bytes_per_launch")
                    wi = pj.get("warp_instructions_per_launch")
                    issue = {"warp_inst_per_env_step": wi / env_steps_per_pass if wi else None,
                             "issue_frac": pj.get("issue_active_frac"), "active_lanes": pj.get("active_lanes_per_inst"),
                             "source": pj.get("source")}
                elif same:
                    issue = {"stale": f"profiles/rollout_profile.json was captured on kernel source {pj.get('source_sha256')}, this run is {src_hash}"}
            except Exception:
                traffic, issue = None, None
        cpu = None
        if not args.no_cpu_baseline and world_size == 1:  # the CPU leg is timed on rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            # bounded sample: ~cpu-seconds of host work at ~3 us per env-step per core
            sample_envs = int(max(cores, min(args.envs, args.cpu_seconds * cores / (3.0e-6 * args.horizon))))
            rate, s_, dt = cpu_reference_rate(world_text, sample_envs, genome_len, args.repeat, args.horizon, cores, 0xB200)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample_envs} envs x {args.horizon} steps of the same workload, {cores} threads, {dt:.1f} s"}
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, genome_len),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_env_step": balg,
                         "issue": issue, "kernel_source_sha256": src_hash,
                         "note": "state stays in registers across the fused horizon, so the kernel is not HBM-bound: it is a dependent f32 chain (no FMA) "
                                 "issued at ~6 cycles per instruction per warp, 14 warps per SM in lock step (DESIGN.md, profiles/r02_experiments.md); "
                                 "`issue` gives its instruction count and issue-slot utilisation"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(args.envs * genome_len),
                    "d2h_bytes_per_step": int(args.envs * 8)},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "secondary": secondary,
            "ga": ga,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


def _timed_pass(torch, stream, fn, flush, reps=3):
    """best-of-`reps` device time (ms) of fn() after one warm-up, L2 flushed before every pass"""
    fn()
    best = None
    for k in range(reps):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best


def secondary_records(prl, torch, dist, dev, local_rank, world_size, stream):
    """The other single-GPU configurations of BASELINE.json, one fused rollout each (device time,
    early exit off, inputs resident): C2 = jump.json 4096 x 1000, C4 = synthetic 64-obstacle world,
    131072 envs per GPU x 1000 (1M envs over 8 GPUs). Whole-job env-steps/s, max over ranks."""
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for key, wname, n, horizon in (("C2_jump_4096x1000", "jump", 4096, 1000), ("C4_synth64_131072x1000_per_gpu", "synth64", 131072, 1000)):
        world = prl.World.from_json(open(os.path.join(ROOT, "worlds", f"{wname}.json")).read())
        env = prl.BatchedEnvironment(world, n, device=local_rank, stream=stream.cuda_stream)
        genomes = torch.from_numpy(make_genomes(n, horizon // 20, 0xB200 + local_rank)).to(dev)
        score = torch.empty(n, dtype=torch.float32, device=dev)
        win = torch.empty(n, dtype=torch.int32, device=dev)
        ms = _timed_pass(torch, stream, lambda: env.rollout(genomes, 20, horizon, early_exit=False, out_score=score, out_step_of_win=win), flush)
        env.sync()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        out[key] = {"value": n * horizon * world_size / (ms * 1e-3), "unit": "env-steps/s", "ms": ms, "n_gpus": world_size,
                    "roofline_frac": algorithmic_bytes(world.counts()[0]) * n * horizon / (ms * 1e-3) / 1e9 / 6551.7}
        del env
    return out


def ga_record(prl, torch, dist, dev, local_rank, rank, world_size, args):
    """BASELINE config C5 shape: the full GA (main.rs:41-163) on push.json, 131072 agents per GPU,
    2000 steps, 3 generations. N > 1: one process per GPU, the generation all-gathered with NCCL once
    per generation, and rank 0's gathered generations compared bit for bit with an unsharded run of
    the same seed on its own GPU. N == 1: the same comparison for the in-library peer shards
    (prl_ga_create_multi) against the unsharded population."""
    world = prl.World.from_json(open(os.path.join(ROOT, "worlds", "push.json")).read())
    per_gpu, gens = args.ga_agents, 3
    kw = dict(number_of_steps=2000, number_of_agents=per_gpu * world_size, repeat_move=20, mutation_rate=0.1, keep_best=False, seed=0xC5)
    stream = torch.cuda.current_stream()
    pop = prl.GaPopulation(world, device=local_rank, stream=stream.cuda_stream, rank=rank, n_ranks=world_size, **kw)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * gens)]
    gathered = []
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pop.init()
    env_steps = 0
    for g in range(gens):
        ev[4 * g].record(stream)
        pop.gather()
        ev[4 * g + 1].record(stream)
        best = pop.best()
        if rank == 0:
            _, _, gs, gg = pop.tensors()
            gathered.append((gs.clone(), gg.clone(), best))
        env_steps += pop.env_steps()
        if g + 1 < gens:
            pop.next_generation()
    torch.cuda.synchronize()
    if world_size > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    ag_ms = sum(ev[4 * g].elapsed_time(ev[4 * g + 1]) for g in range(gens)) / gens
    t = torch.tensor([wall * 1e3 / gens, ag_ms, float(env_steps)], dtype=torch.float64, device=dev)
    if world_size > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        gen_ms, ag_ms, steps_total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        gen_ms, ag_ms, steps_total = float(t[0]), float(t[1]), float(t[2])
    launches = pop.launch_count
    pop.close()
    rec = None
    if rank == 0:
        n_total, L = kw["number_of_agents"], 2000 // 20
        equal = True
        if world_size > 1:
            single = prl.GaPopulation(world, device=local_rank, **kw)
            single.init()
            for g in range(gens):
                b1 = single.best()
                s1, g1 = single.read_generation()
                gs, gg, bm = gathered[g]
                equal &= bool(np.array_equal(gs.cpu().numpy().view(np.uint32), s1.view(np.uint32)) and np.array_equal(gg.cpu().numpy(), g1))
                equal &= bool(np.float32(b1[0]).view(np.uint32) == np.float32(bm[0]).view(np.uint32) and b1[2] == bm[2])
                if g + 1 < gens:
                    single.next_generation()
            single.close()
            how = f"rank 0's NCCL-gathered generations 0..{gens - 1} ({n_total} agents) == unsharded run of the same seed on one GPU, scores and genomes bit for bit"
        else:
            small = dict(kw, number_of_agents=16384 + 5)
            single, multi = prl.GaPopulation(world, device=local_rank, **small), prl.GaPopulation(world, devices=[local_rank] * 3, **small)
            for g in range(gens):
                b1, bm = single.step_generation(), multi.step_generation()
                (s1, g1), (s2, g2) = single.read_generation(), multi.read_generation()
                equal &= bool(np.array_equal(s1.view(np.uint32), s2.view(np.uint32)) and np.array_equal(g1, g2) and b1[2] == bm[2])
            single.close(); multi.close()
            how = f"prl_ga_create_multi with 3 peer shards on the one GPU == unsharded, {small['number_of_agents']} agents x {gens} generations, bit for bit"
        rec = {"workload": f"push.json GA, {per_gpu} agents per GPU ({n_total} total) x 2000 steps, repeat_move 20, mutation 0.1, {gens} generations, early exit on win",
               "gen_ms": gen_ms, "allgather_ms": ag_ms if world_size > 1 else 0.0,
               "bytes": int(n_total * (L + 4)) if world_size > 1 else 0,
               "env_steps_per_s": steps_total / (gen_ms * 1e-3 * gens), "n_ranks": world_size,
               "exchange": "ncclAllGather (torch.distributed) of scores + genomes on the GA's stream" if world_size > 1 else "none (one GPU)",
               "sharded_equals_unsharded": bool(equal), "check": how, "gpu_launches": int(launches)}
    return rec


if __name__ == "__main__":
    main()
