"""CPU checks of bench.py's bookkeeping: the committed ncu traffic table is only used for the kernel sources it was captured
from, the N-rank digest comparison (ranks_agree) works over gloo, and the line's static parts keep the contract's keys."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_traffic_table_is_refused_for_other_kernel_sources(tmp_path, monkeypatch):
    table = {"kernel_source_sha1": bench.kernel_source_hash(), "captured": "test", "dram_bytes_per_particle_step": {"c4": 64.0},
             "fp64_flop_per_particle_step": {"c1": 9000.0}}
    f = tmp_path / "traffic.json"
    f.write_text(json.dumps(table))
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(f))
    t, note = bench.measured_traffic("c4", 5e7)
    assert t == 64.0 * 5e7 and "ncu" in note
    assert bench.measured_traffic("c3", 1e7)[0] is None                    # not profiled
    assert bench.measured_flops("c1") == 9000.0
    table["kernel_source_sha1"] = "0" * 40                                 # a capture of other sources
    f.write_text(json.dumps(table))
    t, note = bench.measured_traffic("c4", 5e7)
    assert t is None and "stale" in note
    assert bench.measured_flops("c1") is None
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(tmp_path / "missing.json"))
    assert bench.measured_traffic("c4", 5e7)[0] is None


def test_committed_traffic_table_names_its_sources():
    with open(bench.TRAFFIC_FILE) as f:
        t = json.load(f)
    assert set(t["sources"]) == set(bench.KERNEL_SOURCES) and len(t["kernel_source_sha1"]) == 40
    assert {"c4", "c5"} <= set(t["dram_bytes_per_particle_step"])


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
import hashlib
def digest(a):
    return int.from_bytes(hashlib.blake2b(np.ascontiguousarray(a).tobytes(), digest_size=8).digest(), "little") >> 1
same = np.arange(1000, dtype=np.int64)
differs = same.copy()
if rank == 1 and sys.argv[1] == "break":
    differs[17] += 1
d = torch.tensor([digest(same), digest(differs)], dtype=torch.int64)
lo, hi = d.clone(), d.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("AGREE", bool(torch.equal(lo, hi)))
dist.destroy_process_group()
'''


def test_rank_agreement_digest_over_gloo(tmp_path):
    """the min == max comparison of per-rank digests bench.py prints as ranks_agree, on two CPU ranks"""
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    for mode, want in (("same", "AGREE True"), ("break", "AGREE False")):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29641", str(w), mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1500:]
        assert want in r.stdout, r.stdout
