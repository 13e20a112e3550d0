"""world_size-2 gloo test of the multi-GPU host logic on CPU: block sharding of the particle arrays and
the all-reduce of the fixed-point charge grid give the same grid, bit for bit, as one rank depositing
everything (the property the NCCL path of libmag2d_b200 relies on; checked with the CPU oracle)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    from mag2d_b200.sharding import shard_range
    from oracle import Oracle, OrcGrid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    g = OrcGrid.make(65, 49, 1.0e-2, 0.75e-2, selfconsistent=1)
    rng = np.random.default_rng(123)            # same global particle set on every rank
    x = rng.uniform(0, 1.0e-2, n)
    z = rng.uniform(0, 0.75e-2, n)
    lo, hi = shard_range(n, rank, world)
    local, bad = orc.deposit_fixed(g, x[lo:hi].copy(), z[lo:hi].copy())
    assert bad == 0
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)     # int64 sum: what ncclAllReduce(ncclInt64, ncclSum) does
    full, _ = orc.deposit_fixed(g, x, z)
    ok = np.array_equal(t.numpy(), full)
    counts = torch.tensor([hi - lo])
    dist.all_reduce(counts)
    ok = ok and int(counts[0]) == n
    with open(os.path.join(out_dir, "rank%d.txt" % rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_sharded_deposit_allreduce_is_bit_exact(tmp_path):
    world, n = 2, 30001
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("rank%d.txt" % r)).read() == "ok"


def test_shard_ranges_partition_exactly():
    from mag2d_b200.sharding import shard_range, shard_seed
    for n in (0, 1, 7, 1000, 10 ** 9 + 7):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    assert len({shard_seed(1234, r) for r in range(8)}) == 8
