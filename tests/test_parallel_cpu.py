"""The N > 1 path on CPU: world_size 2 and 3 `gloo` groups exercising the per-generation exchange
(parallel.gather_generation) with equal and ragged population slices."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from physics_reinforcement_learning_environment_b200.parallel import gather_generation, slice_bounds


def test_slice_bounds_partition_the_population():
    for n in (1, 2, 7, 1000, 65536, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            edges = [slice_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slice_bounds(10, 3, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, genome_len, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # every rank can rebuild the whole generation
        all_scores = rng.random(n).astype(np.float32)
        all_genomes = rng.integers(0, 8, size=(n, genome_len), dtype=np.uint8)
        lo, hi = slice_bounds(n, rank, world)
        gs = torch.full((n,), -1.0)
        gg = torch.zeros((n, genome_len), dtype=torch.uint8)
        gather_generation(torch.from_numpy(all_scores[lo:hi].copy()), torch.from_numpy(all_genomes[lo:hi].copy()), gs, gg)
        ok = np.array_equal(gs.numpy(), all_scores) and np.array_equal(gg.numpy(), all_genomes)
        # wrong slice size must be rejected before any collective is issued
        try:
            gather_generation(torch.zeros(hi - lo + 1), torch.zeros((hi - lo + 1, genome_len), dtype=torch.uint8), gs, gg)
            ok = False
        except ValueError:
            pass
        open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 64), (2, 33), (3, 100)])
def test_gather_generation_gloo(tmp_path, world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, 7, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
