"""C++ host layer (mag2d_b200/csrc/host): Param / species_conf parser / geometry builders compared with
the compiled reference and the Python readers (CPU), and the plasma2d / test_MCC drivers run end to end on
a GPU and compared with the Python front end driving the same library."""
import os
import subprocess

import numpy as np
import pytest

from mag2d_b200 import config as cfg
from mag2d_b200 import decks
from mag2d_b200.geometry import build_geometry
from oracle.pyref import RefHarness, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mag2d_b200", "bin")


@pytest.fixture(scope="module")
def host_bins():
    from mag2d_b200.build import build
    build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "mag2d_b200", "csrc", "host")])
    return BIN


def host_dump(bins, d):
    r = subprocess.run([os.path.join(bins, "host_dump"), "config=" + d["config"], "species_conf=" + d["species_conf"]],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(species=[], interaction=[])
    for line in r.stdout.splitlines():
        key, _, rest = line.partition(" ")
        if key in ("species", "interaction"):
            out[key].append(rest.split())
        elif key in ("mask", "voltage"):
            out[key] = np.array(rest.split(), dtype=float)
        else:
            out[key] = float(rest)
    return out


CASES = [("c1", {}), ("c2", dict(x_sampl=41, z_sampl=41)), ("c3", dict(x_sampl=31, z_sampl=41)), ("c4", dict(x_sampl=33, z_sampl=33))]
CASES += [("c2", dict(geometry=g, x_sampl=41, z_sampl=41, probe_radius=7.5e-3)) for g in ("EMPTY", "PROBE", "RF_8PT", "RF_HAITRAP", "RF_QUAD", "TUBE")]
CASES += [("c3", dict(geometry=g, x_sampl=61, z_sampl=81, r_max=5e-2, z_max=45e-2)) for g in ("MAC", "PENNING", "PENNING_SIMPLE")]


@pytest.mark.parametrize("name,over", CASES)
def test_host_param_species_geometry(host_bins, tmp_path, name, over):
    d = decks.deck(name, str(tmp_path), n_particles=100, **over)
    mine = host_dump(host_bins, d)
    p = cfg.read_config(d["config"])
    for k in ("x_max", "z_max", "x_sampl", "z_sampl", "dx", "dz", "idx", "idz", "V", "dV", "dy", "extern_field", "rf_omega", "rf_amplitude",
              "niter", "mover", "coord", "boundary", "geometry", "selfconsistent", "rf", "t_dist_sample", "t_equilib",
              "macroparticle_factor", "neutral_density"):
        assert mine[k] == float(p[k]), k
    species, inters = cfg.read_species(d["species_conf"])
    assert [s[0] for s in mine["species"]] == [s["name"] for s in species]
    for row, s in zip(mine["species"], species):
        got = [float(v) for v in row[2:]]
        assert got == [s["mass"], s["charge"], s["density"], s["temperature"], s["E_max"], s["dt"]], row
    assert [q[0] for q in mine["interaction"]] == [q["name"] for q in inters]
    for row, q in zip(mine["interaction"], inters):
        assert row[2:4] == [q["primary"], q["secondary"]] and int(row[7]) == len(q["CS_energy"])
        assert [float(v) for v in row[4:7]] == [q["DE"], q["rate"], q["cutoff"]]
    mask, volt = build_geometry(p)
    assert np.array_equal(mine["mask"].reshape(mask.shape), mask)
    fixed = mask < 2
    assert np.array_equal(mine["voltage"].reshape(mask.shape)[fixed], volt[fixed])
    if ref_available():
        with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
            rp = ref.param()
            for k, v in rp.items():
                if k in mine:
                    assert mine[k] == v, k
            assert np.array_equal(mine["mask"].reshape(mask.shape), ref.get_field("mask"))
            assert np.array_equal(mine["voltage"].reshape(mask.shape)[fixed], ref.get_field("voltage")[fixed])


def test_host_config_errors_use_reference_messages(host_bins, tmp_path):
    d = decks.deck("c2", str(tmp_path), n_particles=10, x_sampl=21, z_sampl=21, selfconsistent=1)
    r = subprocess.run([os.path.join(host_bins, "host_dump"), "config=" + d["config"], "species_conf=" + d["species_conf"]],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "selfconsistent rf trap not implemented" in r.stderr
    d = decks.deck("c2", str(tmp_path), n_particles=10, x_sampl=21, z_sampl=21, mover="ADVANCE_LEAPFROG")
    r = subprocess.run([os.path.join(host_bins, "host_dump"), "config=" + d["config"], "species_conf=" + d["species_conf"]],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "unrecognized mover" in r.stderr


def run_driver(bins, which, d, outdir):
    r = subprocess.run([os.path.join(bins, which), "config=" + d["config"], "species_conf=" + d["species_conf"],
                        "initscript=" + d["initscript"], "output_dir=" + outdir], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
def test_plasma2d_driver_matches_python_front_end(host_bins, tmp_path):
    """same deck, same seed, same library: the C++ Pic<CARTESIAN> loop and api.Sim must leave identical
    particles (collisions on, self-consistent C4 deck, 20 steps)."""
    from mag2d_b200.api import Sim
    d = decks.deck("c4", str(tmp_path), n_particles=20000, x_sampl=65, z_sampl=65, r_max=6.4e-3, z_max=6.4e-3, niter=20, t_print=10,
                   t_print_dist=0)
    out = str(tmp_path / "out_cpp")
    log = run_driver(host_bins, "plasma2d_b200", d, out)
    assert "plot 20" in log
    for f in ("out.dat", "potential.dat", "config.txt", "species_conf.txt", "initscript.txt", "ELECTRON_energy_dist.dat", "ARGON_POS_rho.dat"):
        assert os.path.exists(os.path.join(out, f)), f
    pot = np.loadtxt(os.path.join(out, "potential.dat"))
    with Sim(d["config"], d["species_conf"]) as sim:
        sim.run_initscript(d["initscript"])
        sim.set_sort_interval(8)
        sim.advance_init()
        sim.advance(20)
        u = sim.get_field("u")
    assert pot.shape[0] == u.size
    # potential.dat holds the time average over the sampled steps; the last-step potential bounds its scale
    assert np.isfinite(pot).all() and np.abs(pot[:, 2]).max() <= 2 * np.abs(u).max() + 1e-12


@pytest.mark.gpu
def test_test_mcc_driver_runs_c1(host_bins, tmp_path):
    d = decks.deck("c1", str(tmp_path), n_particles=10000, niter=20, t_print=10, t_print_dist=0)
    out = str(tmp_path / "out_mcc")
    log = run_driver(host_bins, "test_MCC_b200", d, out)
    assert "ms / iteration" in log
    e = np.loadtxt(os.path.join(out, "ELECTRON_energy_dist.dat"))
    assert e.shape[1] >= 2 and e[:, 1].sum() > 0


@pytest.mark.gpu
def test_plasma3d_driver_field_and_trajectory(host_bins, tmp_path):
    """plasma3d_b200 (the reference's 3-D test driver, which does not build there): vacuum field of the point electrode
    and one electron trajectory, against the 3-D oracle"""
    from oracle import Oracle3, Orc3Grid
    d = decks.deck("c5", str(tmp_path), n_particles=10, collisions=False, x_sampl=17, y_sampl=15, z_sampl=13, niter=40, dt_elon=2e-10)
    out = str(tmp_path / "out3d")
    r = subprocess.run([os.path.join(host_bins, "plasma3d_b200"), "config=" + d["config"], "output_dir=" + out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    p = cfg.read_config(d["config"])
    g = Orc3Grid.make((17, 15, 13), p["idx"], p["idy"], p["idz"], p["x_max"], p["y_max"], p["z_max"], 0, p["macroparticle_factor"])
    orc = Oracle3()
    mask, volt = orc.geometry(g)
    u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, np.zeros(g.shape)))
    table = np.loadtxt(os.path.join(out, "field3d.dat"))
    assert table.shape == (17 * 15 * 13, 4)
    assert np.abs(table[:, 3].reshape(g.shape) - u_ref).max() <= 1e-5 * np.abs(u_ref).max()      # 6 printed digits
    vtk = open(os.path.join(out, "field3d.vtk")).read().splitlines()
    assert vtk[0] == "# vtk DataFile Version 2.0" and vtk[4] == "DIMENSIONS 17 15 13" and len(vtk) == 10 + 17 * 15 * 13
    volt_table = np.loadtxt(os.path.join(out, "voltage.dat"))
    assert volt_table[:, 3].sum() == 1.0
    traj = np.loadtxt(os.path.join(out, "traj1.dat"))
    vel = np.loadtxt(os.path.join(out, "vel1.dat"))
    assert traj.shape == (40, 3) and vel.shape == (40, 3)
    mass, charge = 9.109534e-31, -1.6e-19
    soa = {k: np.array([v]) for k, v in dict(x=p["x_max"] * 0.5, y=p["y_max"] * 0.5, z=p["z_max"] * 0.7,
                                             vx=np.sqrt(0.03 * 1.602189e-19 / mass * 2.0), vy=0.0, vz=0.0).items()}
    alive = np.ones(1, dtype=np.uint8)
    for step in range(40):
        assert traj[step] == pytest.approx([soa["x"][0], soa["y"][0], soa["z"][0]], rel=2e-5)
        assert vel[step] == pytest.approx([soa["vx"][0], soa["vy"][0], soa["vz"][0]], rel=2e-5, abs=1e-3)
        orc.advance(g, u_ref, mask, charge, mass, 2e-10, (0, 0, 0), soa, alive)


@pytest.mark.gpu
def test_plasma2d_driver_loads_a_magnetic_field_table(host_bins, tmp_path):
    """magnetic_field_const = 0 through the C++ host layer: Pic ctor -> Fields::load_magnetic_field (pic.cpp:148-149) ->
    mag2d_set_magnetic_field; a table that does not cover the box is refused with the reference's interpolate() message"""
    from common import write_btable
    L = 6.4e-3
    bfile = write_btable(str(tmp_path / "btable.txt"), 17, 17, L, L)
    d = decks.deck("c4", str(tmp_path), n_particles=4000, x_sampl=33, z_sampl=33, r_max=L, z_max=L, niter=10, t_print=5, t_print_dist=0,
                   magnetic_field_const=0, magnetic_field_file=bfile)
    out = str(tmp_path / "out_bt")
    log = run_driver(host_bins, "plasma2d_b200", d, out)
    assert "plot 10" in log
    rows = np.loadtxt(os.path.join(out, "out.dat"))
    assert np.isfinite(rows).all()
    small = write_btable(str(tmp_path / "btable_small.txt"), 9, 9, 0.5 * L, L)
    d = decks.deck("c4", str(tmp_path / "bad"), n_particles=4000, x_sampl=33, z_sampl=33, r_max=L, z_max=L, niter=10, t_print=5,
                   t_print_dist=0, magnetic_field_const=0, magnetic_field_file=small)
    r = subprocess.run([os.path.join(host_bins, "plasma2d_b200"), "config=" + d["config"], "species_conf=" + d["species_conf"],
                        "initscript=" + d["initscript"], "output_dir=" + str(tmp_path / "out_bad")], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "outside of range" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_plasma2d_driver_with_particle_source(host_bins, tmp_path):
    """use_source = 1 through the C++ host layer: the driver refreshes one reservoir per particle species after
    advance_init (test.cpp:56-59) and Pic::advance runs Species::source() after every push (pic.cpp:346-347)"""
    Lx = 6.4e-3
    d = decks.deck("c4", str(tmp_path), n_particles=4000, x_sampl=33, z_sampl=33, r_max=Lx, z_max=Lx, niter=20, t_print=10, t_print_dist=0,
                   use_source=1, src_fact=4, density_total=1e13)
    out = str(tmp_path / "out_src")
    log = run_driver(host_bins, "plasma2d_b200", d, out)
    assert "ELECTRON source initialized:" in log and "ARGON_POS source initialized:" in log and "plot 20" in log
    rows = np.loadtxt(os.path.join(out, "out.dat"))
    assert np.isfinite(rows).all()
