"""torchrun worker for tests/test_multi_gpu.py: particles sharded over the ranks, NCCL all-reduce of the
fixed-point charge grid inside mag2d_step; rank 0 then repeats the run alone with all particles and checks
that the charge grid and the potential are bit-identical."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import disk_particles  # noqa: E402
from mag2d_b200 import decks  # noqa: E402
from mag2d_b200.api import Sim  # noqa: E402
from mag2d_b200.sharding import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tmp = tempfile.mkdtemp(prefix="mgpu_%d_" % rank)
    n = 40000
    d = decks.deck("c4", tmp, n_particles=2 * n, collisions=False, x_sampl=65, z_sampl=65, r_max=6.4e-3, z_max=6.4e-3)
    rng = np.random.default_rng(7)
    ai = disk_particles(rng, n, 3.2e-3, 3.2e-3, 2.8e-3, 300.0)
    ae = disk_particles(rng, n, 3.3e-3, 3.2e-3, 2.8e-3, 6e5)

    def run(sim, lo, hi, steps):
        ii, ie = sim.species_index("ARGON_POS"), sim.species_index("ELECTRON")
        sim.set_particles(ii, ai[lo:hi])
        sim.set_particles(ie, ae[lo:hi])
        sim.advance_init()
        sim.advance(steps)
        return sim.rho_fixed(ii), sim.rho_fixed(ie), sim.get_field("u")
    sim = Sim(d["config"], d["species_conf"], device=local)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    sim.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    lo, hi = shard_range(n, rank, world)
    ri, re, u = run(sim, lo, hi, 4)
    sim.close()
    # the same for the 3-D path (coord = CARTESIAN3D): one species, fused cell sort on, all-reduce of the [M][K][N] grid
    d3 = decks.deck("c5", tmp, n_particles=n, collisions=False, x_sampl=14, y_sampl=12, z_sampl=13, macroparticle_factor=2e6)
    a3 = np.zeros((n, 7))
    a3[:, 0] = rng.uniform(0, 1.3e-3, n)
    a3[:, 1] = rng.uniform(0, 1.1e-3, n)
    a3[:, 2] = rng.uniform(0, 1.2e-3, n)
    a3[:, 3:6] = rng.normal(size=(n, 3)) * 3e5

    def run3(sim3, lo, hi, steps):
        e = sim3.species_index("ELECTRON")
        sim3.set_particles(e, a3[lo:hi])
        sim3.set_sort_interval(2)
        sim3.advance_init()
        sim3.advance(steps)
        return sim3.rho_fixed(e), sim3.get_field("u")
    sim3 = Sim(d3["config"], d3["species_conf"], device=local)
    uid3 = torch.zeros(128, dtype=torch.uint8, device=dev)      # every communicator needs its own id
    if rank == 0:
        uid3.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid3, 0)
    sim3.comm_init(rank, world, bytes(uid3.cpu().numpy().tobytes()))
    r3, u3 = run3(sim3, lo, hi, 4)
    sim3.close()
    # a grid whose x planes do not divide by the rank count: the slab-parallel solve spreads the potential with broadcasts instead of
    # one all-gather
    d3b = decks.deck("c5", tmp + "_odd", n_particles=n, collisions=False, x_sampl=15, y_sampl=12, z_sampl=13, macroparticle_factor=2e6)
    sim3b = Sim(d3b["config"], d3b["species_conf"], device=local)
    uid3b = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid3b.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid3b, 0)
    sim3b.comm_init(rank, world, bytes(uid3b.cpu().numpy().tobytes()))
    r3b, u3b = run3(sim3b, lo, hi, 3)
    sim3b.close()
    ok = True
    if rank == 0:
        with Sim(d["config"], d["species_conf"], device=local) as solo:
            si, se, su = run(solo, 0, n, 4)
        ok = np.array_equal(ri, si) and np.array_equal(re, se) and np.array_equal(u, su)
        with Sim(d3["config"], d3["species_conf"], device=local) as solo3:
            s3, su3 = run3(solo3, 0, n, 4)
        ok3 = np.array_equal(r3, s3) and np.array_equal(u3, su3)
        with Sim(d3b["config"], d3b["species_conf"], device=local) as solo3b:
            s3b, su3b = run3(solo3b, 0, n, 3)
        ok3 = ok3 and np.array_equal(r3b, s3b) and np.array_equal(u3b, su3b)
        print("MGPU_RESULT_3D", "ok" if ok3 else "mismatch", "max|du|", float(np.abs(u3 - su3).max()))
        ok = ok and ok3
        print("MGPU_RESULT", "ok" if ok else "mismatch", "world", world, "max|du|", float(np.abs(u - su).max()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
