"""The files either side of the hot path, against the reference itself:

  * Field2D::load (src/Field2D.cpp:46-130): the C++ host layer's and the Python front end's readers against the compiled
    reference class on the same files (CPU);
  * checkpoints: BaseSpecies::save of the reference -> Species::load of the host layer -> device store -> Species::save ->
    BaseSpecies::load of the reference (src/particles.cpp:32-93), byte for byte (GPU);
  * the output files of a whole run — out.dat, <NAME>.dat, potential.dat, <NAME>_rho.dat, <NAME>_energy_dist.dat
    (src/particles.cpp:367-384, src/pic.cpp:429-461) — of the reference's own plasma2d binary (oracle/_ref/plasma2d) and of
    plasma2d_b200 on a collision-free deck started from the same checkpoint, column by column (GPU);
  * electric_field_from_file (src/pic.cpp:154-177): gather and trajectories on fields read from files (GPU);
  * the energy histogram against histogram.cpp through BaseSpecies::energy_dist_compute (GPU).
"""
import os
import subprocess

import numpy as np
import pytest

from common import disk_particles
from mag2d_b200 import config as cfg
from mag2d_b200 import decks
from oracle.pyref import REF_DIR, RefHarness, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mag2d_b200", "bin")
needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref is not built")
PARTICLE = np.dtype([("x", "f8"), ("y", "f8"), ("z", "f8"), ("vx", "f8"), ("vy", "f8"), ("vz", "f8"), ("time_to_death", "f8"),
                     ("empty", "u1"), ("pad", "u1", (7,))])


@pytest.fixture(scope="module")
def host_bins():
    from mag2d_b200.build import build
    build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mag2d_b200", "csrc", "host")])
    return BIN


def write_field(path, M, N, x_max, z_max, f, order="ascending", blank_rows=True, header=None):
    """a potential in the format Field2D::print writes and Field2D::load reads: "x<TAB>z<TAB>value", x slowest"""
    xs, zs = np.linspace(0.0, x_max, M), np.linspace(0.0, z_max, N)
    data = f(xs[:, None], zs[None, :]) + np.zeros((M, N))
    ii = range(M) if order == "ascending" else range(M - 1, -1, -1)
    with open(path, "w") as out:
        if header:
            out.write(header + "\n")
        for i in ii:
            jj = range(N) if order == "ascending" else range(N - 1, -1, -1)
            for j in jj:
                out.write("%.17g\t%.17g\t%.17g\n" % (xs[i], zs[j], data[i, j]))
            if blank_rows:
                out.write("\n")
    return data


def write_checkpoint(path, aos7):
    """BaseSpecies::save's format (src/particles.cpp:32-59): int particles.size(); int n; n x t_particle (64 bytes)"""
    rec = np.zeros(len(aos7), dtype=PARTICLE)
    for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "time_to_death")):
        rec[k] = aos7[:, c]
    with open(path, "wb") as f:
        f.write(np.array([len(rec), len(rec)], dtype=np.int32).tobytes())
        f.write(rec.tobytes())


def read_checkpoint(path):
    raw = open(path, "rb").read()
    cap, n = np.frombuffer(raw[:8], dtype=np.int32)
    return int(cap), int(n), np.frombuffer(raw[8:], dtype=PARTICLE)


# ------------------------------------------------------------------------------------------------ Field2D::load (CPU)
@needs_ref
@pytest.mark.parametrize("order,blank,header", [("ascending", True, None), ("descending", False, "# x z u"), ("ascending", False, None)])
def test_field2d_load_matches_reference(host_bins, tmp_path, order, blank, header):
    path = str(tmp_path / "field.dat")
    data = write_field(path, 23, 17, 1.1e-2, 0.8e-2, lambda x, z: np.sin(300 * x) * np.cosh(200 * z) + 7 * x, order, blank, header)
    d = decks.deck("c2", str(tmp_path), n_particles=10, x_sampl=23, z_sampl=17)
    with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
        want = ref.field2d_load(path)
    assert (want["M"], want["N"]) == (23, 17) and np.array_equal(want["data"], data)
    r = subprocess.run([os.path.join(host_bins, "host_dump"), "field2d=" + path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    head = lines[0].split()
    assert head[0] == "field2d" and (int(head[1]), int(head[2])) == (23, 17)
    assert [float(v) for v in head[3:7]] == [want["x_min"], want["z_min"], want["x_max"], want["z_max"]]
    assert np.array_equal(np.array(lines[1].split()[1:], dtype=float).reshape(23, 17), want["data"])
    mine = cfg.load_field2d(path)
    assert (mine["M"], mine["N"], mine["x_min"], mine["z_min"]) == (23, 17, want["x_min"], want["z_min"])
    assert np.array_equal(mine["data"], want["data"])


@needs_ref
def test_field2d_load_errors_match_reference(host_bins, tmp_path):
    d = decks.deck("c2", str(tmp_path), n_particles=10, x_sampl=23, z_sampl=17)
    bad = str(tmp_path / "short.dat")
    write_field(bad, 9, 7, 1e-2, 1e-2, lambda x, z: x + z)
    lines = open(bad).read().splitlines()
    open(bad, "w").write("\n".join(lines[:-5]) + "\n")            # rows missing: the grid is not rectangular any more
    with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
        for path, text in ((str(tmp_path / "nope.dat"), "failed opening file"), (bad, "wrong size of input vector")):
            with pytest.raises(RuntimeError, match=text):
                ref.field2d_load(path)
            with pytest.raises(RuntimeError, match=text):
                cfg.load_field2d(path)
            r = subprocess.run([os.path.join(host_bins, "host_dump"), "field2d=" + path], capture_output=True, text=True)
            assert r.returncode != 0 and text in r.stderr


# ------------------------------------------------------------------------------------------------ checkpoints (GPU)
@needs_ref
@pytest.mark.gpu
def test_checkpoint_round_trip_through_the_device_store_is_byte_exact(host_bins, tmp_path):
    d = decks.deck("c4", str(tmp_path), n_particles=6000, collisions=False, x_sampl=33, z_sampl=33, r_max=6.4e-3, z_max=6.4e-3)
    rng = np.random.default_rng(3)
    names = ("ARGON_POS", "ELECTRON")
    start = {n: disk_particles(rng, 3000, 3.2e-3, 3.2e-3, 3.1e-3, v) for n, v in zip(names, (4e2, 6e5))}
    ck1 = tmp_path / "ck_ref"
    ck1.mkdir()
    write_checkpoint(str(ck1 / "particles_ARGON.dat"), np.zeros((0, 7)))       # every species is reloaded (particles.hpp:159)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        for n in names:
            i = ref.species_index(n)
            ref.set_particles(i, start[n])
        ref.advance_init()
        ref.advance(3)                           # some particles leave through the walls: the store has empty slots
        kept = {}
        for n in names:
            i = ref.species_index(n)
            ref.species_save(i, str(ck1 / ("particles_%s.dat" % n)))
            p = ref.get_particles(i)
            kept[n] = p[p[:, 7] > 0][:, :7]
            assert 0 < len(kept[n]) <= 3000
    # host layer: Pic ctor -> Species::load -> mag2d_particles_upload; Pic::save -> mag2d_particles_download -> file
    d2 = decks.deck("c4", str(tmp_path / "reload"), n_particles=6000, collisions=False, x_sampl=33, z_sampl=33, r_max=6.4e-3, z_max=6.4e-3,
                    particle_reload=1, particle_reload_dir=str(ck1))
    out = str(tmp_path / "ck_b200")
    r = subprocess.run([os.path.join(host_bins, "checkpoint_b200"), "config=" + d2["config"], "species_conf=" + d2["species_conf"],
                        "output_dir=" + out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
    for n in names:
        cap1, n1, rec1 = read_checkpoint(str(ck1 / ("particles_%s.dat" % n)))
        cap2, n2, rec2 = read_checkpoint(os.path.join(out, "particles_%s.dat" % n))
        assert n1 == n2 == len(kept[n]) and cap1 >= n1 and cap2 == n2          # the reference's first word counts its empty slots too
        assert rec1.tobytes() == rec2.tobytes(), n                            # every record, padding included
    # ... and the reference reads what the host layer wrote
    with RefHarness(d["config"], d["species_conf"], seed=6) as ref:
        for n in names:
            i = ref.species_index(n)
            ref.species_load(i, os.path.join(out, "particles_%s.dat" % n))
            p = ref.get_particles(i)
            assert np.array_equal(p[p[:, 7] > 0][:, :7], kept[n]), n


# ------------------------------------------------------------------------------------------------ whole-run outputs (GPU)
def run_driver(exe, d, outdir):
    r = subprocess.run([exe, "config=" + d["config"], "species_conf=" + d["species_conf"], "initscript=" + d["initscript"],
                        "output_dir=" + outdir], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def table(path):
    return np.loadtxt(path, ndmin=2)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["selfconsistent", "field_from_file"])
def test_plasma2d_output_files_match_the_reference_binary(host_bins, tmp_path, kind):
    """oracle/_ref/plasma2d (src/test.cpp compiled unmodified) and plasma2d_b200, same decks, same particles (both load the same
    checkpoint), collisions off: every output file agrees column by column.  Tolerances: trajectories 1e-10 after 40 steps
    (means over the ensemble much tighter), averaged charge 1e-9 of its maximum (fixed-point against sequential fp64
    sums), potential 1e-8 (solver), histogram bins exactly the reference's bin rule (a particle within one ulp of a bin edge may
    move: at most 2 counts of 8000 x samples)."""
    ref_exe = os.path.join(REF_DIR, "plasma2d")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/plasma2d is not built")
    rng = np.random.default_rng(11)
    ck = tmp_path / "start"
    ck.mkdir()
    L = 6.4e-3
    common = dict(collisions=False, x_sampl=33, z_sampl=33, r_max=L, z_max=L, niter=40, t_print=10, t_print_dist=0, do_plot=0,
                  particle_reload=1, particle_reload_dir=str(ck))
    if kind == "selfconsistent":
        names, vth = ("ARGON_POS", "ELECTRON"), (4e2, 6e5)
        d = decks.deck("c4", str(tmp_path / "deck"), n_particles=8000, density_total=1e14, **common)
    else:
        # static + RF potentials read from files (what config_haitrap_vtk.txt does), RF drive on, not self-consistent
        names, vth = ("H_NEG",), (2.5e3,)
        fu, frf = str(tmp_path / "u_static.dat"), str(tmp_path / "u_rf.dat")
        write_field(fu, 33, 33, L, L, lambda x, z: 40.0 * ((x - L / 2) ** 2 - (z - L / 2) ** 2) / L ** 2 + 3.0 * x / L)
        write_field(frf, 33, 33, L, L, lambda x, z: np.sin(np.pi * x / L) * np.sin(2 * np.pi * z / L))
        d = decks.deck("c2", str(tmp_path / "deck"), n_particles=4000, geometry="RF_8PT", electric_field_from_file=1,
                       electric_field_static_file=fu, electric_field_rf_file=frf, extern_field=0.0, has_probe=0, **common)
    n_each = 4000
    for n, v in zip(names, vth):
        write_checkpoint(str(ck / ("particles_%s.dat" % n)), disk_particles(rng, n_each, L / 2, L / 2, 0.45 * L, v))
    # the reference flags EVERY species as a particle species (particles.hpp:159), so the neutral gas needs its (empty) file too
    for s_ in cfg.read_species(d["species_conf"])[0]:
        if s_["name"] not in names:
            write_checkpoint(str(ck / ("particles_%s.dat" % s_["name"])), np.zeros((0, 7)))
    open(d["initscript"], "w").write("")
    out_ref, out_b = str(tmp_path / "out_ref"), str(tmp_path / "out_b200")
    log_ref = run_driver(ref_exe, d, out_ref)
    log_b = run_driver(os.path.join(host_bins, "plasma2d_b200"), d, out_b)
    assert "plot 40" in log_ref and "plot 40" in log_b
    # out.dat: iter, cpu time (not comparable), probe current, U_trap, u[0][...]
    a, b = table(os.path.join(out_ref, "out.dat")), table(os.path.join(out_b, "out.dat"))
    assert a.shape == b.shape == (4, 5)
    assert np.array_equal(a[:, 0], b[:, 0])
    assert np.allclose(a[:, 2:], b[:, 2:], rtol=1e-8, atol=1e-12 * max(1.0, np.abs(a[:, 2:]).max()))
    for n in names:
        # <NAME>.dat: niter, n_particles, mean energy of the sampled steps, t
        a, b = table(os.path.join(out_ref, n + ".dat")), table(os.path.join(out_b, n + ".dat"))
        assert a.shape == b.shape and a.shape[0] == 4
        assert np.array_equal(a[:, :2], b[:, :2]), n                  # step counter and particle count: exact
        assert np.allclose(a[:, 2], b[:, 2], rtol=1e-5), n            # printed with 6 digits
        assert np.allclose(a[:, 3], b[:, 3], rtol=1e-5), n
        # energy distribution (normalised histogram of the sampled steps)
        a, b = table(os.path.join(out_ref, n + "_energy_dist.dat")), table(os.path.join(out_b, n + "_energy_dist.dat"))
        assert a.shape == b.shape == (200, 2) and np.array_equal(a[:, 0], b[:, 0])
        norm = np.abs(a[:, 1]).sum()
        assert norm > 0 and np.abs(a[:, 1] - b[:, 1]).sum() <= 1e-5 * norm + 4.0 * a[:, 1].max() / (n_each * 10), n
        # time-averaged charge per node
        a, b = table(os.path.join(out_ref, n + "_rho.dat")), table(os.path.join(out_b, n + "_rho.dat"))
        assert a.shape == b.shape == (33 * 33, 3) and np.array_equal(a[:, :2], b[:, :2])
        if kind == "selfconsistent":
            assert np.abs(a[:, 2]).max() > 0
        assert np.abs(a[:, 2] - b[:, 2]).max() <= 2e-6 * max(np.abs(a[:, 2]).max(), 1e-300), n      # 6 printed digits
    a, b = table(os.path.join(out_ref, "potential.dat")), table(os.path.join(out_b, "potential.dat"))
    assert a.shape == b.shape == (33 * 33, 3) and np.array_equal(a[:, :2], b[:, :2])
    assert np.abs(a[:, 2] - b[:, 2]).max() <= 2e-6 * max(np.abs(a[:, 2]).max(), 1e-300)


# ------------------------------------------------------------------------------------------------ field from file (GPU)
@needs_ref
@pytest.mark.gpu
def test_electric_field_from_file_gather_and_trajectories_vs_reference(tmp_path):
    from mag2d_b200.api import Sim
    L = 8e-3
    fu, frf = str(tmp_path / "u_static.dat"), str(tmp_path / "u_rf.dat")
    write_field(fu, 41, 37, L, L, lambda x, z: 25.0 * np.cos(2 * x / L) * np.exp(z / L) - 11.0 * z / L, order="descending")
    write_field(frf, 41, 37, L, L, lambda x, z: 9.0 * np.sin(3 * x / L + 0.3) * np.sin(2 * z / L + 0.1))
    d = decks.deck("c2", str(tmp_path), n_particles=10, collisions=False, geometry="RF_8PT", x_sampl=41, z_sampl=37, r_max=L, z_max=L,
                   electric_field_from_file=1, electric_field_static_file=fu, electric_field_rf_file=frf, Bz=0.02, Bt=0.01)
    rng = np.random.default_rng(8)
    aos = disk_particles(rng, 3000, L / 2, L / 2, 0.49 * L, 3e3)
    with RefHarness(d["config"], d["species_conf"], seed=2) as ref, Sim(d["config"], d["species_conf"]) as sim:
        assert np.array_equal(sim.get_field("u"), ref.get_field("u")) and np.array_equal(sim.get_field("uRF"), ref.get_field("uRF"))
        x, z = rng.uniform(0, L, 4000), rng.uniform(0, L, 4000)
        for t in (0.0, 1.7e-8):
            ex, ez = sim.field_E(x, z, t)
            rx, rz = ref.field_E(x, z, t)
            scale = max(np.abs(rx).max(), np.abs(rz).max())
            assert np.abs(ex - rx).max() <= 1e-12 * scale and np.abs(ez - rz).max() <= 1e-12 * scale
        h = sim.species_index("H_NEG")
        hr = ref.species_index("H_NEG")
        sim.set_particles(h, aos)
        ref.set_particles(hr, aos)
        for steps, tol in ((1, 1e-12), (49, 1e-10)):
            sim.advance(steps)           # Pic::advance on both sides: the species clock (RF phase) advances with every step
            ref.advance(steps)
            a, b = sim.get_particles(h), ref.get_particles(hr)
            # electrodes do not absorb when the field comes from a file (particles.hpp:395): only the box does
            assert np.array_equal(a[:, 7] > 0, b[:, 7] > 0)
            live = b[:, 7] > 0
            cols = [0, 2, 3, 4, 5]
            err = np.abs(a[live][:, cols] - b[live][:, cols]).max(axis=0) / np.abs(b[live][:, cols]).max(axis=0)
            assert err.max() <= tol, (steps, err)


# ------------------------------------------------------------------------------------------------ histogram (GPU)
@needs_ref
@pytest.mark.gpu
def test_energy_histogram_matches_reference_histogram_class(tmp_path):
    """mag2d_energy_hist against Histogram::add (src/histogram.cpp:22-35) through BaseSpecies::energy_dist_compute
    (src/particles.cpp:408-414): 200 bins over (0, E_max), strict inequalities at both ends, totals over all particles"""
    from mag2d_b200.api import Sim
    d = decks.deck("c4", str(tmp_path), n_particles=100, collisions=False, x_sampl=33, z_sampl=33, r_max=6.4e-3, z_max=6.4e-3)
    rng = np.random.default_rng(21)
    with RefHarness(d["config"], d["species_conf"], seed=2) as ref, Sim(d["config"], d["species_conf"]) as sim:
        for name, vth in (("ELECTRON", 1.1e6), ("ARGON_POS", 9e2)):
            aos = disk_particles(rng, 50000, 3.2e-3, 3.2e-3, 3e-3, vth)
            aos[:5, 3:6] = 0.0                               # energy exactly 0: outside the open interval, counted in the totals
            i, ir = sim.species_index(name), ref.species_index(name)
            sim.set_particles(i, aos)
            ref.set_particles(ir, aos)
            want, st = ref.energy_hist(ir)
            emax = sim.species_get(i, "E_max")
            assert st["min"] == 0.0 and st["max"] == emax
            got, gs = sim.energy_hist(i, 200, emax)
            assert 0 < st["n_val"] < 50000                   # the tail beyond E_max is dropped from the bins
            assert gs["n_in"] == st["n_val"] and gs["n_tot"] == 50000
            assert np.abs(got - want).sum() <= 2             # a particle within an ulp of a bin edge may sit in the neighbouring bin
            assert gs["sum_in"] / gs["n_in"] == pytest.approx(st["mean"], rel=1e-12)
            assert gs["mean_tot"] == pytest.approx(st["mean_tot"], rel=1e-12)
