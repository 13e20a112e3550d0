#!/usr/bin/env python
"""bench.py — throughput of the PIC/MCC hot path (push + MCC + boundary + deposit, with the Poisson solve
and the periodic cell sort inside the timed region) in particle-steps/s, on synthetic particle loads of the
shapes BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|c1] [--particles P]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's own CPU code on the host cores

Default workload: C4 of BASELINE.json — 2-D self-consistent Ar+/e- discharge, 512x512 grid, 1e8 particles
per GPU, Poisson solve every step — the one configuration that exercises every part of the metric
(push + MCC + deposit, solve ms/step).  The other configs are parity-test cases (tests/), selectable here
with --workload for exploration only.

One step = one Pic::advance (reference src/pic.cpp:330-358).  A "particle-step" is one live particle
advanced by one step.  Weak scaling: every rank pushes its own --particles shard and the fixed-point charge
grid is all-reduced (NCCL) before the replicated solve.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_PARTICLE_STEP = 80.0       # 2D3V fp64: x,z,vx,vy,vz read + written once (SURVEY.md §8d / DESIGN.md)
WORKLOADS = {
    "c4": "C4: 2-D self-consistent Ar+/e- discharge, 512x512 grid, Poisson each step, Boris + MCC + deposit",
    "c2": "C2: 22-pole RF ion trap, 200x200, H- in He/H2 buffer gas, Boris + RF gather + Langevin MCC",
    "c3": "C3: cylindrical r-z, 200x100, Bz=0.03 T, electrons, self-consistent, Boris",
    "c1": "C1: e- swarm in He, E=1 kV/m, multi-collision mover (~90 events per particle-step)",
    "c5": "C5: 3-D self-consistent electron cloud in Ar, 256^3 grid, Poisson each step, Boris + MCC + 8-node deposit",
}
DEFAULT_PARTICLES = {"c4": 100_000_000, "c2": 1_000_000, "c3": 10_000_000, "c1": 1_000_000, "c5": 125_000_000}
BYTES = {"c4": 80.0, "c2": 80.0, "c3": 80.0, "c1": 96.0, "c5": 96.0}


KERNEL_SOURCES = ("push.cu", "push3d.cu", "push3d.cuh", "push3d_brick.cu", "mcc.cuh", "common.cuh")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_push_dram.json")


def kernel_source_hash():
    """sha1 over the kernel sources the profiled launches come from: a committed ncu capture only speaks for the code it profiled"""
    import hashlib
    h = hashlib.sha1()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "mag2d_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _traffic_table():
    with open(TRAFFIC_FILE) as f:
        t = json.load(f)
    if t.get("kernel_source_sha1") != kernel_source_hash():
        return None, "stale: %s was captured from other kernel sources (profiles/make_traffic.py refreshes it)" % os.path.basename(TRAFFIC_FILE)
    return t, None


def measured_traffic(workload, n_particles):
    """DRAM bytes per push launch from this round's `ncu --set full` capture (profiles/r2_push_dram.json holds
    dram__bytes_read.sum + dram__bytes_write.sum per particle of the profiled launches, keyed by the sha1 of the kernel sources),
    scaled to this run's particles per launch.  -> (bytes or None, note): a capture of other sources is refused, not reused"""
    try:
        t, why = _traffic_table()
        if t is None:
            return None, why
        v = t["dram_bytes_per_particle_step"].get(workload)
        if v is None:
            return None, "workload not profiled"
        return float(v) * n_particles, "ncu --set full, %s" % t.get("captured", "this round")
    except Exception as e:
        return None, "no capture (%s)" % str(e)[:60]


def measured_flops(workload):
    try:
        t, why = _traffic_table()
        return float(t["fp64_flop_per_particle_step"][workload]) if t else None
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bind_near_gpu(torch, index):
    """pin this process to the CPUs of the GPU's NUMA node before any host buffer is allocated: the e2e leg moves 8 GB per
    step through pinned host memory, and pages first touched on the far socket cost more than half of the PCIe rate.
    Returns a description for the JSON line (PCIe link included); never fails the run."""
    info = {}
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        info["pci"] = bdf
        base = "/sys/bus/pci/devices/" + bdf
        for key in ("numa_node", "current_link_speed", "current_link_width"):
            try:
                with open(os.path.join(base, key)) as f:
                    info[key] = f.read().strip()
            except OSError:
                pass
        with open(os.path.join(base, "local_cpulist")) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                if part:
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound_cpus"] = len(use)
        else:
            info["bound_cpus"] = 0          # one node, or nothing to narrow
    except Exception as e:                  # no sysfs entry (container), no permission: run unbound
        info["error"] = str(e)[:80]
    return info


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            t = [c.strip() for c in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1]))
                mx.append(float(t[2]))
                power.append(float(t[3]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if t[5 + k].lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_deck(workload, particles_per_gpu, world, tmp):
    from mag2d_b200 import decks
    n_global = particles_per_gpu * world
    if workload == "c4":
        # dV = V/cells with V = n_particles_total/density_total (param.cpp:134-136): the macro-particle
        # weight follows the GLOBAL count per species so that the physical density stays 1e15 m^-3
        return decks.deck("c4", tmp, n_particles=particles_per_gpu, n_particles_total=n_global // 2)
    if workload == "c5":
        # the macro-particle volume follows the GLOBAL count (Param derives dy from it, param.cpp:133-137)
        L, dy = 1e-4 * 255, 1e-4
        return decks.deck("c5", tmp, n_particles=particles_per_gpu, n_particles_total=n_global, density_total=n_global / (L * L * dy))
    return decks.deck(workload, tmp, n_particles=particles_per_gpu)


def load_particles(sim, workload, d, n):
    """synthetic load through the device-side Philox loaders (the initscript verbs)"""
    if workload == "c3":
        sim.generate(sim.species_index("ELECTRON"), "cylinder", n, 1.0, 3.75e-2, 4e-3, 2e-2)
    elif workload == "c2":
        # uniform disk of radius 4 mm in the 22-pole trap (field-free core of the trap)
        sim.generate(sim.species_index("H_NEG"), "on_disk", n, 1e-2, 1e-2, 4e-3)
    elif workload == "c5":
        sim.generate(sim.species_index("ELECTRON"), "everywhere", n)
    else:
        sim.run_initscript(d["initscript"])


def state_checksums(sim, part_species):
    """64-bit digests of the potential and of every species' int64 charge grid (what the ranks must agree on bit for bit)"""
    import hashlib

    import numpy as np
    out = []
    for arr in [sim.get_field("u")] + [sim.rho_fixed(s) for s in part_species]:
        h = hashlib.blake2b(np.ascontiguousarray(arr).tobytes(), digest_size=8).digest()
        out.append(int.from_bytes(h, "little") >> 1)          # 63 bits: fits a signed int64 tensor
    return out


def run_workload(wl, args, env, steps, warmup, e2e_steps, with_cpu):
    """one workload on this rank's GPU: load, warm up, time `steps` steps (device events on the launching stream, max over
    ranks), phase times, roofline, solve, rank agreement, optionally the e2e leg and the CPU arm.  -> dict (rank 0 uses it)"""
    import numpy as np

    from mag2d_b200.api import Sim
    torch, dist, world, rank, local_rank, dev, stream = (env[k] for k in ("torch", "dist", "world", "rank", "local_rank", "dev", "stream"))
    n = (args.particles if wl == args.workload else 0) or DEFAULT_PARTICLES[wl]
    tmp = tempfile.mkdtemp(prefix="mag2d_bench_")
    d = make_deck(wl, n, world, tmp)
    sim = Sim(d["config"], d["species_conf"], device=local_rank, stream=stream.cuda_stream, seed=1234 + rank)
    f32 = args.storage == "f32" and wl == args.workload
    if f32:
        sim.set_storage("f32")          # before any particle is loaded
        e2e_steps = 0                   # the streamed step moves the caller's double arrays: fp64 storage only
    bytes_per = BYTES[wl] / 2 if f32 else BYTES[wl]
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        sim.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
    selfconsistent = bool(sim.param["selfconsistent"])
    load_particles(sim, wl, d, n)
    part_species = [sim.species_index(s) for s in d["species"]]
    for s in part_species:
        sim.sort(s)
    sort_interval = args.sort_interval
    sim.set_sort_interval(sort_interval)
    species_sort = {}
    if args.sort_intervals and wl == args.workload:
        for item in args.sort_intervals.split(","):
            name, k = item.split("=")
            species_sort[name] = int(k)
            sim.set_sort_interval(int(k), species=sim.species_index(name))
    if selfconsistent:
        sim.set_solver_kind(args.solver)
        sim.set_solver(cycles_per_step=args.cycles, tol=1e-12, max_cycles=60)
    direct = selfconsistent and sim.solver_is_direct()
    three_d = wl == "c5"
    sim.advance_init()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    sim.advance(warmup)
    barrier()
    if selfconsistent:
        sim.solver_stats()      # reset the residual monitor
    n_live = sum(sim.count(s)[0] for s in part_species)
    launches0 = sim.kernel_launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # ---- timed region: exactly K steps, device-timed on the launching stream, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    sim.advance(steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop() if rank == 0 else None
    monitored_resid = sim.solver_stats()["resid"] if selfconsistent else None
    launches = sim.kernel_launches() - launches0
    n_live_end = sum(sim.count(s)[0] for s in part_species)
    # ---- do the ranks hold the same potential and the same (all-reduced) charge grids, bit for bit?
    ranks_agree = None
    if selfconsistent:
        digests = torch.tensor(state_checksums(sim, part_species), dtype=torch.int64, device=dev)
        if world > 1:
            lo, hi = digests.clone(), digests.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ranks_agree = bool(torch.equal(lo, hi))
        else:
            ranks_agree = True
    if three_d:
        monitored_resid = sim.solve()["resid"]       # measured after the timed region (the 3-D step does not monitor)
    t = torch.tensor([ms, float(n_live), float(n_live_end)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0])
        n_live_all, n_live_end_all = float(tsum[1]), float(tsum[2])
    else:
        n_live_all, n_live_end_all = float(n_live), float(n_live_end)
    # live particles decay slowly through wall losses: use the mean of the two counts
    pstep = 0.5 * (n_live_all + n_live_end_all) * steps
    value = pstep / (ms * 1e-3)

    # ---- per-phase device times (events around each phase; adds one sync per step, so it is a separate pass)
    sim.set_timing(True)
    sim.advance(steps)
    tm = sim.timers()
    sim.set_timing(False)
    push_ms = tm["push"] / steps
    n_now = sum(sim.count(s)[0] for s in part_species)
    n_push_launches = len(part_species) * (2 if sim.param["rf"] else 1)
    peak, peak_src = measured_peaks()
    achieved = bytes_per * n_now / (push_ms * 1e-3) / 1e9 if push_ms > 0 else 0.0
    traffic, traffic_note = measured_traffic(wl + ("_f32" if f32 else ""), n_now / n_push_launches)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                # the same launches on the DRAM bytes ncu saw them move (frac uses the contract's algorithmic bytes)
                "frac_moved": (traffic * n_push_launches / (push_ms * 1e-3) / 1e9 / peak) if traffic and push_ms > 0 else None,
                "algorithmic_bytes_per_launch": bytes_per * n_now / n_push_launches,
                "kernel": "%s (fused gather+push+MCC+boundary+deposit), %d launches/step" % (
                    "k_push3d / k_push3d_brick" if three_d else "k_push_multicoll" if wl == "c1" else "k_push_boris", n_push_launches),
                "algorithmic_bytes_per_particle_step": bytes_per, "push_ms_per_step": push_ms,
                "peak_source": peak_src}
    extra = {}
    if wl == "c1":
        # the multi-collision mover is instruction-bound: report collision events/s and FP64 FLOP/s
        he = [k for k, s in enumerate(sim.species) if s["type"] == 0]
        e = part_species[0]
        sim.set_collision_counting(True)
        sim.collision_counts(e, reset=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record(stream)
        sim.advance(steps)
        c1.record(stream)
        torch.cuda.synchronize()
        cnt = sim.collision_counts(e, reset=True)
        sim.set_collision_counting(False)
        events = float(cnt.sum())
        extra["collisions"] = {"events_per_s": events / (c0.elapsed_time(c1) * 1e-3), "events_per_particle_step": events / (n_now * steps),
                               "real_collisions_per_particle_step": float(sum(cnt[k * 16:k * 16 + 16].sum() for k in he)) / (n_now * steps),
                               "note": "null + real events of the null-collision method, counted in a separate pass (the counters are off in the timed region)"}
        fl = measured_flops(wl)
        if fl:
            extra["fp64"] = {"flop_per_particle_step": fl, "achieved_tflops": fl * value / 1e12, "peak_tflops": 37.0,
                             "note": "FP64 FLOP per particle-step from the committed ncu capture (DFMA = 2, DMUL/DADD = 1) x the measured rate; "
                                     "peak = 148 SMs x 64 FP64 FMA lanes x 2 x 1.965 GHz"}
    solve_info = None
    if direct:
        solve_info = {"kind": ("direct: sine transforms along y and z (FP64 matrix products) x tridiagonal solve per mode along x, "
                               "point electrodes by the capacitance-matrix method; exact to round-off like the reference's LU")
                      if three_d else
                      ("direct: sine transform along z (FP64 matrix product) x tridiagonal solve per mode along x "
                       "(the grid has no internal electrodes); exact to round-off like the reference's LU"),
                      "ms_per_step": tm["solve"] / steps, "vcycles_per_step": 0,
                      "max_resid_over_timed_steps": monitored_resid,
                      "resid_def": "max|r_k/a_kk| / max|u| (largest Jacobi update relative to the potential)"}
    elif selfconsistent:
        # how far from converged is the field after the fixed number of warm-started V-cycles per step?
        info = sim.solve(rf=False, tol=1e-12)
        # and how many cycles does a step need when it iterates to the tolerance (one host sync per cycle)?
        sim.set_solver(cycles_per_step=0, tol=args.solve_tol, max_cycles=60)
        need = []
        for _ in range(4):
            sim.advance(1)
            need.append(sim.solver_stats()["cycles"])
        sim.set_solver(cycles_per_step=args.cycles, tol=1e-12, max_cycles=60)
        solve_info = {"kind": "geometric multigrid (Galerkin coarse operators), fixed V-cycles per step",
                      "ms_per_step": tm["solve"] / steps, "vcycles_per_step": abs(args.cycles),
                      "first_guess": "2u_n - u_(n-1)" if args.cycles < 0 else "u_n",
                      "ms_per_vcycle": tm["solve"] / steps / max(abs(args.cycles), 1),
                      "max_resid_over_timed_steps": monitored_resid,
                      "note": "relative error of u against the converged solve is ~5x resid (calibrated on this deck)",
                      "extra_cycles_to_1e-12": info["cycles"], "resid_after_extra": info["resid"],
                      "cycles_per_step_to_tol": need, "tol": args.solve_tol,
                      "resid_def": "max|r_k/a_kk| / max|u| (largest Jacobi update relative to the potential)"}

    # ---- end-to-end through the C ABI with HOST buffers: per step, particles go host -> device from pinned
    # memory, one Pic::advance runs, particles and the charge grid come back (what a host-resident caller
    # of Species::advance pays when it keeps the reference's host-side particle array)
    e2e = None
    if e2e_steps > 0 and (rank == 0 or world > 1):
        e2e = bench_e2e(sim, part_species, args, torch, stream, world, dist, dev, e2e_steps)
        e2e["host"] = env["host_binding"]

    cpu = None
    if rank == 0 and with_cpu:
        try:
            cpu = cpu_baseline(wl, d, sim, part_species, args)
        except Exception as ex:     # the checker must never take the bench down
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 1, "kind": "unavailable", "sample": repr(ex)}
    rec = {
        "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "dtype": "f32 storage of the particle arrays in HBM, f64 arithmetic" if f32 else "f64",
        "config": {"workload": WORKLOADS[wl], "particles_per_gpu": n, "live_particles": n_live_all,
                   "grid": list(sim.shape),
                   "species": d["species"], "sort_interval": sort_interval, "species_sort_interval": species_sort,
                   "l2": "inputs larger than L2 (%.1f GB of particle state per GPU)" % (n * 40 / 1e9)
                   if n * 40 > 200e6 else "particle state fits L2: flush not applied, see roofline note",
                   "parallelism": "particle shards, %d rank(s), NCCL all-reduce of the int64 charge grid" % world,
                   "poisson": (("direct (2 sine transforms x tridiagonal + capacitance matrix)" if three_d else "direct (sine transform x tridiagonal)") if direct else "multigrid, %d V-cycles/step" % abs(args.cycles))
                   if selfconsistent else "none (vacuum field solved once)"},
        "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "ranks_agree": ranks_agree,
        "phases_ms_per_step": {k: v / steps for k, v in tm.items()}, "solve": solve_info, "e2e": e2e, "cpu_baseline": cpu,
    }
    if three_d:
        rec["store"] = {d["species"][q]: sim.store_stats(s) for q, s in enumerate(part_species)}
    rec.update(extra)
    sim.close()
    return rec


def bench_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_binding = bind_near_gpu(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    # a dedicated (non-default) stream: the context enqueues on it and the timing events are recorded on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    env = dict(torch=torch, dist=dist, world=world, rank=rank, local_rank=local_rank, dev=dev, stream=stream, host_binding=host_binding)
    wl = args.workload
    rec = run_workload(wl, args, env, args.steps, args.warmup, args.e2e_steps, not args.no_cpu_baseline)
    secondary = None
    if wl == "c4" and not args.no_secondary:
        # BASELINE.json configs[4] (the named 8-GPU configuration: 256^3, 1.25e8 particles per GPU) rides on every line so that
        # the driver's 1 -> 8 GPU scaling record carries it too; C4 stays the headline value
        r5 = run_workload("c5", args, env, min(args.steps, args.secondary_steps), args.warmup, 0, False)
        secondary = {"c5": {k: r5[k] for k in ("value", "ms_per_step", "steps", "warmup", "config", "gpu_launches", "roofline",
                                               "ranks_agree", "phases_ms_per_step", "solve", "store")}}
        secondary["c5"].update(unit="particle-steps/s", n_gpus=world, scaling="weak", dtype="f64",
                               allreduce_ms_per_step=r5["phases_ms_per_step"]["allreduce"])
    if rank == 0:
        out = {
            "metric": "particle-steps/sec (push+MCC+deposit, Poisson solve and periodic cell sort inside the step)",
            "value": rec["value"], "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": rec["dtype"], "data": "synthetic (device-side Philox loaders, seed 1234)",
        }
        for k in ("config", "gpu_launches", "clocks", "roofline", "ranks_agree", "phases_ms_per_step", "solve", "e2e", "cpu_baseline",
                  "collisions", "fp64", "store"):
            if k in rec:
                out[k] = rec[k]
        if secondary:
            out["secondary"] = secondary
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_e2e(sim, part_species, args, torch, stream, world, dist, dev, e2e_steps):
    import ctypes as C

    import numpy as np
    steps = max(1, e2e_steps)
    bufs = {}
    h2d = d2h = 0
    if sim.is3d:
        # the device-resident 3-D store is binned by brick, with slack behind every bin (1.5 slots per particle): a caller that keeps
        # the particles on the host holds the live ones only, so the store is compacted (stand-alone cell sort) before it is handed over
        sim.set_store_layout("slots")
        for s in part_species:
            sim.sort(s)
    for s in part_species:
        n_slots = sim.count(s)[1]
        comps = ("x", "y", "z", "vx", "vy", "vz") if sim.is3d else ("x", "z", "vx", "vy", "vz")
        host = {k: torch.empty(n_slots, dtype=torch.float64).pin_memory() for k in comps}
        bufs[s] = (n_slots, host)
        h2d += len(comps) * 8 * n_slots
        d2h += len(comps) * 8 * n_slots
    n_nodes = int(np.prod(sim.shape))
    rho_host = torch.empty(n_nodes, dtype=torch.float64).pin_memory()
    d2h += 8 * n_nodes
    dp = C.POINTER(C.c_double)

    def ptr(t):
        return C.cast(t.data_ptr(), dp)

    def download(s):
        n_slots, host = bufs[s]
        ns = C.c_int64()
        sim._chk(sim.L.mag2d_particles_download_soa(sim.h, s, n_slots, ptr(host["x"]), ptr(host["y"]) if sim.is3d else None, ptr(host["z"]), ptr(host["vx"]),
                                                    ptr(host["vy"]), ptr(host["vz"]), None, None, C.byref(ns)))

    def upload(s):
        n_slots, host = bufs[s]
        sim._chk(sim.L.mag2d_particles_clear(sim.h, s))
        sim._chk(sim.L.mag2d_particles_upload_soa(sim.h, s, n_slots, ptr(host["x"]), ptr(host["y"]) if sim.is3d else None, ptr(host["z"]), ptr(host["vx"]),
                                                  ptr(host["vy"]), ptr(host["vz"]), None))
    for s in part_species:
        download(s)          # the host-side particle arrays the caller owns
    n_live = sum(sim.count(s)[0] for s in part_species)
    streamed = int(sim.param["mover"]) == 0
    if streamed:
        # the device copy is dropped: from here on the particles live in the caller's (pinned) host arrays only
        for s in part_species:
            sim._chk(sim.L.mag2d_particles_clear(sim.h, s))
        pointers = [[bufs[s][1][k].data_ptr() for k in comps] for s in part_species]
        counts = [bufs[s][0] for s in part_species]
    # the host side of these boxes is a shared VM: one disturbed repeat can double the time of a PCIe-bound step, so the
    # leg is timed `repeats` times over `steps` steps each and the median repeat is reported (all are listed)
    def one_step():
        if streamed:
            sim.step_streamed(part_species, counts, pointers, chunk_slots=args.e2e_chunk)
        else:
            for s in part_species:
                upload(s)
            sim.advance(1)
            for s in part_species:
                download(s)
        sim._chk(sim.L.mag2d_rho_download(sim.h, ptr(rho_host)))

    one_step()              # untimed: the staging ring and the copy streams are created by the first streamed call
    if streamed:
        # what one step really copies (the library leaves vy in place when the push does not need it and the array is pinned)
        sim.streamed_bytes(reset=True)
        one_step()
        sb = sim.streamed_bytes(reset=True)
        h2d, d2h = sb[0], sb[1] + 8 * n_nodes
    repeats = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        torch.cuda.synchronize()
        repeats.append(time.perf_counter() - t0)
    dt = sorted(repeats)[len(repeats) // 2]
    t = torch.tensor([dt, float(n_live)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, n_live = float(tmax[0]), float(tsum[1])
    # per-rank link rates of this rank's own median repeat (min / max over the ranks): with N ranks on one host the sum of these is
    # what the host's memory system and PCIe root complexes deliver in aggregate
    own = sorted(repeats)[len(repeats) // 2] / steps
    rate = torch.tensor([h2d / own / 1e9, d2h / own / 1e9], dtype=torch.float64, device=dev)
    lo, hi, tot = rate.clone(), rate.clone(), rate.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    link = {"h2d_gb_per_s_per_rank_min_max": [round(float(lo[0]), 2), round(float(hi[0]), 2)],
            "d2h_gb_per_s_per_rank_min_max": [round(float(lo[1]), 2), round(float(hi[1]), 2)],
            "host_aggregate_gb_per_s_both_directions": round(float(tot[0] + tot[1]), 2)}
    return {"value": n_live * steps / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": dt / steps * 1e3, "link": link,
            "repeats_ms_per_step": [round(r / steps * 1e3, 3) for r in repeats],
            "path": ("mag2d_step_streamed%s: host SoA arrays (pinned) -> chunked H2D / fused step / D2H overlapped on three streams -> host arrays, "
                     "+ mag2d_rho_download; bytes as counted by the library (vy is not copied when the push does not read it: "
                     "the collision pass reaches the pinned array in place)") % ("3" if sim.is3d else "") if streamed else
                    "mag2d_particles_upload_soa (pinned host) -> mag2d_step -> mag2d_particles_download_soa + mag2d_rho_download"}


class _StdoutToStderr:
    """the reference prints progress on C++ stdout (seed, loader echoes); keep bench.py's stdout to ONE JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def reference_run(workload, d, n_cpu, steps, u=None, variant="fast"):
    with _StdoutToStderr():
        return _reference_run(workload, d, n_cpu, steps, u, variant)


def _reference_run(workload, d, n_cpu, steps, u=None, variant="fast"):
    """time the reference's own CPU implementation (oracle/_ref) of the particle phase of Pic::advance
    on a bounded sample of the workload; falls back to the oracle port when oracle/_ref is absent"""
    import numpy as np

    from mag2d_b200 import config as cfg
    from oracle import RefHarness, ref_available
    p = cfg.read_config(d["config"])
    rng = np.random.default_rng(1234)
    sp, _ = cfg.read_species(d["species_conf"])
    names = [s["name"] for s in sp]
    per = max(1, n_cpu // len(d["species"]))

    def sample(name):
        s = sp[names.index(name)]
        vth = np.sqrt(1.380662e-23 * s["temperature"] / s["mass"])
        aos = np.zeros((per, 7))
        if workload == "c2":
            ang = rng.uniform(0, 2 * np.pi, per)
            rad = np.sqrt(rng.uniform(0, 1, per)) * 4e-3
            aos[:, 0] = 1e-2 + rad * np.cos(ang)
            aos[:, 2] = 1e-2 + rad * np.sin(ang)
        elif workload == "c3":
            aos[:, 0] = np.sqrt(rng.uniform(0, 1, per)) * 4e-3
            aos[:, 2] = 3.75e-2 + 2e-2 * (rng.uniform(0, 1, per) - 0.5)
        else:
            aos[:, 0] = rng.uniform(0, p["x_max"], per) * (1 - 1e-12)
            aos[:, 2] = rng.uniform(0, p["z_max"], per) * (1 - 1e-12)
        aos[:, 3:6] = rng.normal(size=(per, 3)) * vth
        # cell-sorted like the GPU store, which is also the friendliest order for the CPU caches
        key = (aos[:, 0] * p["idx"]).astype(np.int64) * int(p["z_sampl"]) + (aos[:, 2] * p["idz"]).astype(np.int64)
        return aos[np.argsort(key, kind="stable")]
    if ref_available(variant):
        # the reference's constructor pre-solves the vacuum field with UMFPACK (pic.cpp:180-187); the shim's
        # banded LU would need 2 GB and minutes at 512x512, so it is skipped on this timing arm (u = 0 is the
        # exact vacuum solution of the grounded box) — see oracle/shims/umfpack_shim.cpp
        os.environ["MAG2D_UMFPACK_SHIM_MAXN"] = "60000"
        with RefHarness(d["config"], d["species_conf"], seed=1234, variant=variant) as ref:
            for name in d["species"]:
                ref.set_particles(ref.species_index(name), sample(name))
            if u is not None:
                ref.set_field("u", u)
            ref.advance_particles(2)       # warm the caches, as the GPU arm's warm-up does
            n_live = sum(ref.species(ref.species_index(name))["n_particles"] for name in d["species"])
            wall, cpu = ref.time_advance(steps, particles_only=True)
            n_end = sum(ref.species(ref.species_index(name))["n_particles"] for name in d["species"])
        return {"value": 0.5 * (n_live + n_end) * steps / wall, "unit": "particle-steps/s", "cores": 1, "kind": "reference",
                "sample": "%d particles x %d steps of the same deck, reference sources compiled -Ofast (oracle/_ref), "
                          "Species::advance of every species + rho sum (pic.cpp:343-354), single thread: the "
                          "reference's simulation loop is serial; wall %.2f s, user-cpu %.2f s" % (n_live, steps, wall, cpu),
                "seconds": wall}
    raise RuntimeError("oracle/_ref is not built on this machine")


def port3d_run(d, n_cpu, steps, u=None):
    """CPU arm of the 3-D workload: the reference's 3-D step does not compile (species3d.cpp is dead code), so the
    restatement oracle/mag3d_oracle.c (pinned to the reference's Field3D / Geometry) is what runs: kind = port"""
    import numpy as np

    from mag2d_b200 import config as cfg
    from oracle import Oracle3, Orc3Grid
    p = cfg.read_config(d["config"])
    sp, _ = cfg.read_species(d["species_conf"])
    s = [q for q in sp if q["name"] == "ELECTRON"][0]
    g = Orc3Grid.make((int(p["x_sampl"]), int(p["y_sampl"]), int(p["z_sampl"])), p["idx"], p["idy"], p["idz"], p["x_max"], p["y_max"],
                      p["z_max"], int(p["boundary"]), p["macroparticle_factor"])
    orc = Oracle3()
    mask, _ = orc.geometry(g)
    rng = np.random.default_rng(1234)
    vth = np.sqrt(1.380662e-23 * s["temperature"] / s["mass"])
    soa = {k: np.ascontiguousarray(rng.uniform(0, hi, n_cpu)) for k, hi in (("x", p["x_max"]), ("y", p["y_max"]), ("z", p["z_max"]))}
    key = (np.floor(soa["x"] * p["idx"]).astype(np.int64) * int(p["y_sampl"]) + np.floor(soa["y"] * p["idy"]).astype(np.int64)) * int(p["z_sampl"]) \
        + np.floor(soa["z"] * p["idz"]).astype(np.int64)
    order = np.argsort(key, kind="stable")         # cell-sorted like the GPU store
    for k in ("x", "y", "z"):
        soa[k] = np.ascontiguousarray(soa[k][order])
    for k in ("vx", "vy", "vz"):
        soa[k] = np.ascontiguousarray(rng.normal(size=n_cpu) * vth)
    alive = np.ones(n_cpu, dtype=np.uint8)
    if u is None:
        u = np.zeros(g.shape)
    fixed = np.zeros(g.shape, dtype=np.int64)
    orc.advance(g, u, mask, s["charge"], s["mass"], s["dt"], (0, 0, 0), soa, alive, None, fixed)     # warm the caches
    n0 = int(alive.sum())
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.advance(g, u, mask, s["charge"], s["mass"], s["dt"], (0, 0, 0), soa, alive, None, fixed)
    wall = time.perf_counter() - t0
    n1 = int(alive.sum())
    return {"value": 0.5 * (n0 + n1) * steps / wall, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": "%d particles x %d steps of the same deck, oracle/mag3d_oracle.c (gcc -O2; the reference's 3-D step "
                      "species3d.cpp does not compile) gather + Boris + boundary + 8-node deposit, collisions and the field solve "
                      "excluded, single thread; wall %.2f s" % (n0, steps, wall),
            "seconds": wall}


def cpu_baseline(workload, d, sim, part_species, args):
    if workload == "c5":
        return port3d_run(d, 2_000_000, 40, sim.get_field("u"))
    u = sim.get_field("u") if sim.param["selfconsistent"] else None
    n_cpu = {"c1": 40000}.get(workload, 4_000_000)      # about 10 s of single-thread work
    steps = {"c1": 25}.get(workload, 100)
    return reference_run(workload, d, n_cpu, steps, u)


def bench_reference(args):
    """--impl reference: the reference's own CPU code on the host cores, same metric/config, rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    n = args.particles or DEFAULT_PARTICLES[wl]
    tmp = tempfile.mkdtemp(prefix="mag2d_bench_ref_")
    d = make_deck(wl, n, 1, tmp)
    n_cpu = {"c1": 20000}.get(wl, 4_000_000)
    try:
        if wl == "c5":
            r = port3d_run(d, 1_000_000, args.steps)
        else:
            for _ in range(max(0, min(args.warmup, 1))):
                reference_run(wl, d, n_cpu, 2)
            r = reference_run(wl, d, n_cpu, args.steps)
    except Exception as ex:
        print(json.dumps({"impl": "reference", "unavailable": repr(ex)[:200]}))
        return
    from mag2d_b200 import config as cfg
    p = cfg.read_config(d["config"])
    out = {
        "impl": "reference",
        "metric": "particle-steps/sec (push+MCC+deposit, Poisson solve and periodic cell sort inside the step)",
        "value": r["value"], "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["seconds"] / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (numpy, seed 1234)",
        "config": {"workload": WORKLOADS[wl], "particles_per_gpu": n, "grid": [int(p["x_sampl"]), int(p["z_sampl"])],
                   "species": d["species"], "note": "each step is a bounded sample of %d particles of this workload; "
                   "the field solve is excluded on this arm (UMFPACK is not available here; SURVEY.md §8c)" % n_cpu},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's named size)")
    ap.add_argument("--sort-interval", type=int, default=-1, help="pushes between cell sorts; -1: per species from its thermal drift")
    ap.add_argument("--sort-intervals", default="", help="per-species overrides, e.g. ARGON_POS=64,ELECTRON=2")
    ap.add_argument("--cycles", type=int, default=-3,
                    help="multigrid V-cycles per step; negative: |n| cycles from the time-extrapolated guess 2u_n - u_(n-1)")
    ap.add_argument("--solver", default="auto", choices=["auto", "multigrid", "direct"],
                    help="Poisson solver of the self-consistent step (auto: direct when the grid separates)")
    ap.add_argument("--solve-tol", type=float, default=1e-10)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="slots per chunk of the streamed e2e step (0: the library default, 4 Mi)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--storage", default="f64", choices=["f64", "f32"],
                    help="element type of the device-resident particle arrays (f32: 2-D Boris workloads; a separate line, never the headline)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C5 record that rides on the C4 line")
    ap.add_argument("--secondary-steps", type=int, default=16)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    # the contract is ONE JSON line on stdout: anything native code prints there (NCCL's version banner, the
    # reference's progress echoes) is sent to stderr; only the final print goes to the real stdout
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w")
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
