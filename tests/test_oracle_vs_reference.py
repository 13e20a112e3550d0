"""Pin the CPU restatement (oracle/mag2d_oracle.c) against the UNMODIFIED reference compiled into
oracle/_ref (libmag2d_ref_parity.so, built -O2 -ffp-contract=off like the restatement).

Everything deterministic must agree bit for bit; RNG-driven code is compared under the same SHR3 seed
(the restatement reproduces t_random's float/double mix exactly).  The one tolerance is the Langevin
branch, where std::tr1::comp_ellint_1 (libstdc++) is replaced by an AGM evaluation of K(k).
"""
import os

import numpy as np
import pytest

from common import Particles, disk_particles, grid_from_param, model_from, needs_ref, write_btable
from mag2d_b200 import config as cfg
from mag2d_b200 import decks
from oracle import RefHarness

pytestmark = pytest.mark.skipif(not needs_ref, reason="oracle/_ref not built (make -C oracle ref)")


def test_rng_streams_bit_exact(orc, deckdir):
    d = decks.deck("c1", deckdir)
    with RefHarness(d["config"], d["species_conf"], seed=1234) as ref:
        r = orc.rng(1234)
        for what in ("iuni", "uni", "rnor", "rexp", "radius"):
            assert np.array_equal(ref.rng_draw(what, 100000), orc.rng_draw(r, what, 100000)), what
        assert np.array_equal(ref.rng_rot(2.5, 500), orc.rng_rot(r, 2.5, 500))
        v = np.random.default_rng(0).normal(size=(500, 3))
        assert np.array_equal(ref.rng_deflect(0.3, v), orc.rng_deflect(r, 0.3, v))


def test_param_and_species_parsing_match_reference(orc, deckdir):
    for name in ("c1", "c2", "c3", "c4"):
        d = decks.deck(name, deckdir, n_particles=1000, x_sampl=21, z_sampl=21)
        with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
            p = ref.param()
            mine = cfg.read_config(d["config"])
            for k, v in p.items():
                if k in mine:
                    assert float(mine[k]) == v, (name, k)
            m, names = model_from(orc, d["species_conf"])
            assert names == ref.species_names()
            for i in range(len(names)):
                s = ref.species(i)
                assert s["lifetime"] == m.lifetime(i), (name, names[i])
                assert np.array_equal(ref.rates(i), m.rates(i))
                for key in ("mass", "charge", "density", "temperature", "E_max", "dt", "v_max"):
                    assert s[key] == m.get(i, key), (name, names[i], key)


def test_known_answer_lifetime_c1(orc, deckdir):
    # check_params prints lifetime = 1.11353e-10 for species_conf_MCC.txt ELECTRON
    d = decks.deck("c1", deckdir)
    m, names = model_from(orc, d["species_conf"])
    assert abs(m.lifetime(names.index("ELECTRON")) - 1.11353e-10) < 1e-15


@pytest.mark.parametrize("geometry", ["EMPTY", "PROBE", "RF_22PT", "RF_8PT", "RF_HAITRAP", "RF_QUAD", "TUBE"])
def test_geometry_and_presolved_fields_cartesian(orc, deckdir, geometry):
    d = decks.deck("c2", deckdir, n_particles=10, geometry=geometry, x_sampl=41, z_sampl=41, probe_radius=7.5e-3)
    with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
        p = ref.param()
        g = grid_from_param(p)
        mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
        rmask = ref.get_field("mask")
        assert np.array_equal(mask, rmask)
        fixed = rmask < 2
        assert np.array_equal(volt[fixed], ref.get_field("voltage")[fixed])
        # Pic ctor: boundary_solve_rf(); boundary_solve(); reset()   (pic.cpp:180-187)
        zero = np.zeros((g.M, g.N))
        u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=False))
        urf = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=True))
        assert np.abs(u - ref.get_field("u")).max() <= 1e-12 * max(1.0, np.abs(u).max())
        assert np.abs(urf - ref.get_field("uRF")).max() <= 1e-12


@pytest.mark.parametrize("geometry", ["EMPTY", "MAC", "PENNING", "PENNING_SIMPLE"])
def test_geometry_cylindrical(orc, deckdir, geometry):
    d = decks.deck("c3", deckdir, geometry=geometry, x_sampl=61, z_sampl=81, r_max=5e-2, z_max=45e-2)
    with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
        p = ref.param()
        g = grid_from_param(p)
        mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
        rmask = ref.get_field("mask")
        assert np.array_equal(mask, rmask)
        fixed = rmask < 2
        assert np.array_equal(volt[fixed], ref.get_field("voltage")[fixed])


def test_gather_E_bit_exact_including_edges(orc, deckdir):
    d = decks.deck("c2", deckdir, n_particles=10, geometry="RF_8PT", x_sampl=41, z_sampl=41)
    with RefHarness(d["config"], d["species_conf"], seed=1) as ref:
        p = ref.param()
        g = grid_from_param(p)
        u, urf = ref.get_field("u"), ref.get_field("uRF")
        rng = np.random.default_rng(3)
        n = 4000
        x = rng.uniform(0, p["x_max"], n)
        z = rng.uniform(0, p["z_max"], n)
        x[:10] = 0
        z[10:20] = 0
        x[20:30] = p["x_max"]
        z[30:40] = p["z_max"]
        for t in (0.0, 3.3e-8, 1.234e-3):
            ex, ez = ref.field_E(x, z, t)
            ex2, ez2 = orc.field_E(g, u, urf, x, z, t)
            assert np.array_equal(ex, ex2) and np.array_equal(ez, ez2)
        mask = ref.get_field("mask").astype(np.uint8)
        assert np.array_equal(ref.is_free(x[40:], z[40:]), orc.is_free(g, mask, x[40:], z[40:]))


def test_boris_cartesian_trajectory_bit_exact(orc, deckdir):
    d = decks.deck("c2", deckdir, n_particles=10, geometry="RF_8PT", x_sampl=41, z_sampl=41, Bt=0.01, Bz=0.02, Br=0.005)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        u, urf = ref.get_field("u"), ref.get_field("uRF")
        m, names = model_from(orc, d["species_conf"])
        h = names.index("H_NEG")
        aos = disk_particles(np.random.default_rng(3), 1500, 1e-2, 1e-2, 2.5e-3, 1500.0)
        ref.set_particles(h, aos)
        P = Particles.from_aos7(aos)
        ref.species_set(h, "niter", 17)
        ref.advance_position(h, init=True)
        orc.advance_boris_init(g, u, urf, m, h, P, niter=17)
        assert np.array_equal(ref.get_particles(h)[:, :7], P.aos7())
        ref.species_set(h, "lifetime", np.inf)   # collisions off on the reference side
        for step in range(100):
            ref.species_set(h, "niter", 17 + step)
            ref.advance_position(h)
            orc.advance_boris(g, u, urf, m, h, P, niter=17 + step, rng=None)
        assert np.array_equal(ref.get_particles(h)[:, :7], P.aos7())
        mask = ref.get_field("mask").astype(np.uint8)
        ref.advance_boundary(h)
        orc.advance_boundary(g, mask, m.get(h, "charge"), P)
        assert np.array_equal(ref.get_particles(h)[:, 7], P.alive)


def test_boris_with_mcc_same_seed(orc, deckdir):
    d = decks.deck("c2", deckdir, n_particles=10, geometry="RF_8PT", x_sampl=41, z_sampl=41)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        u, urf = ref.get_field("u"), ref.get_field("uRF")
        m, names = model_from(orc, d["species_conf"])
        h = names.index("H_NEG")
        # scatter only (Langevin + tabulated elastic on two targets)
        r = orc.rng(9)
        ref.rng_seed(9)
        v = np.random.default_rng(1).normal(size=(20000, 3)) * 1500
        a = ref.scatter(h, v)
        b, proc, targ = orc.scatter(m, h, r, v)
        assert np.abs(a - b).max() <= 1e-11 * 1500     # K(k): AGM vs libstdc++ comp_ellint_1
        assert np.array_equal(ref.rng_draw("iuni", 5), orc.rng_draw(r, "iuni", 5))   # streams stay in step
        assert (proc >= 0).sum() > 1000
        aos = disk_particles(np.random.default_rng(3), 2000, 1e-2, 1e-2, 2.5e-3, 1500.0)
        ref.set_particles(h, aos)
        P = Particles.from_aos7(aos)
        ref.rng_seed(11)
        orc.rng_seed(r, 11)
        for step in range(30):
            ref.species_set(h, "niter", step)
            ref.advance_position(h)
            orc.advance_boris(g, u, urf, m, h, P, niter=step, rng=r)
        out = ref.get_particles(h)[:, :7]
        assert np.abs(out[:, [0, 2]] - P.aos7()[:, [0, 2]]).max() < 1e-15
        assert np.abs(out[:, 3:6] - P.aos7()[:, 3:6]).max() < 1e-9


def test_multicoll_same_seed_bit_exact(orc, deckdir):
    d = decks.deck("c1", deckdir, n_particles=2000)
    with RefHarness(d["config"], d["species_conf"], seed=1234) as ref:
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        p = ref.param()
        g = grid_from_param(p)
        mask, _ = orc.geometry(g, 0)
        ref.run_initscript(d["initscript"])
        parts = ref.get_particles(e)
        aos = parts[parts[:, 7] > 0, :7].copy()
        assert aos.shape[0] == 2000
        ref.set_particles(e, aos)
        P = Particles.from_aos7(aos)
        ref.rng_seed(99)
        r = orc.rng(99)
        for step in range(3):
            ref.advance_position(e)
            orc.advance_multicoll(0.0, p["extern_field"], m, e, P, r)
            ref.advance_boundary(e)
            orc.advance_boundary(g, mask, m.get(e, "charge"), P)
        out = ref.get_particles(e)
        assert np.array_equal(out[:, :7], P.aos7())
        assert np.array_equal(out[:, 7], P.alive)


def test_selfconsistent_pic_loop_bit_exact(orc, deckdir):
    # two particle species, collisions with a continuum neutral, deposit + solve each step
    d = decks.deck("c4", deckdir, n_particles=20000, x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        m, names = model_from(orc, d["species_conf"])
        ii, ie = names.index("ARGON_POS"), names.index("ELECTRON")
        mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
        rng = np.random.default_rng(7)
        ai = disk_particles(rng, 10000, 1.6e-3, 1.6e-3, 1.4e-3, 300.0)
        ae = disk_particles(rng, 10000, 1.7e-3, 1.6e-3, 1.4e-3, 6e5)
        ref.set_particles(ii, ai)
        ref.set_particles(ie, ae)
        Pi, Pe = Particles.from_aos7(ai), Particles.from_aos7(ae)
        qi, qe = m.get(ii, "charge"), m.get(ie, "charge")
        ref.advance_init()
        rho_i, _ = orc.deposit_fp64(g, qi, Pi.x, Pi.z)
        rho_e, _ = orc.deposit_fp64(g, qe, Pe.x, Pe.z)
        rho = np.zeros_like(rho_i)
        rho += rho_i
        rho += rho_e
        assert np.array_equal(rho, ref.get_field("rho"))
        u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        assert np.array_equal(u, ref.get_field("u"))
        urf = ref.get_field("uRF")
        orc.advance_boris_init(g, u, urf, m, ii, Pi)
        orc.advance_boris_init(g, u, urf, m, ie, Pe)
        r = orc.rng(21)
        ref.rng_seed(21)
        for step in range(5):
            ref.advance()
            u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
            rho_i[:] = 0
            rho_e[:] = 0
            orc.advance_boris(g, u, urf, m, ii, Pi, niter=step, rng=r)
            orc.advance_boundary(g, mask, qi, Pi, rho=rho_i)
            orc.advance_boris(g, u, urf, m, ie, Pe, niter=step, rng=r)
            orc.advance_boundary(g, mask, qe, Pe, rho=rho_e)
            rho = np.zeros_like(rho_i)
            rho += rho_i
            rho += rho_e
        oi, oe = ref.get_particles(ii), ref.get_particles(ie)
        assert np.array_equal(oi[:, :7], Pi.aos7()) and np.array_equal(oi[:, 7], Pi.alive)
        assert np.array_equal(oe[:, :7], Pe.aos7()) and np.array_equal(oe[:, 7], Pe.alive)
        assert np.array_equal(rho, ref.get_field("rho"))
        assert np.array_equal(u, ref.get_field("u"))
        # the build's fixed-point deposit stays within n_contrib * 2^-33 * |q| of the fp64 grid
        rf, _ = orc.deposit_fixed(g, Pe.x, Pe.z, Pe.alive)
        ncontrib, _ = orc.deposit_fixed(g, Pe.x, Pe.z, Pe.alive)
        assert np.abs(rf * 2.0 ** -32 * qe - rho_e).max() <= 10000 * 2.0 ** -33 * abs(qe)


def test_cylindrical_selfconsistent_bit_exact(orc, deckdir):
    d = decks.deck("c3", deckdir, n_particles=5000, x_sampl=41, z_sampl=51)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        assert p["coord"] == 1 and p["selfconsistent"] == 1
        m, names = model_from(orc, d["species_conf"])
        ie = names.index("ELECTRON")
        mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
        rng = np.random.default_rng(11)
        n = 5000
        aos = np.zeros((n, 7))
        aos[:, 0] = np.sqrt(rng.uniform(0, 1, n)) * 4e-3
        aos[:, 2] = 3.75e-2 + 2e-2 * (rng.uniform(0, 1, n) - 0.5)
        aos[:, 3:6] = rng.normal(size=(n, 3)) * 4e5
        aos[:5, 0] = 0.0      # r == 0 guard of the frame rotation (particles.cpp:607-611)
        aos[:5, 3] = 0.0
        aos[:5, 4] = 0.0
        ref.set_particles(ie, aos)
        Pe = Particles.from_aos7(aos)
        qe = m.get(ie, "charge")
        ref.advance_init()
        rho, _ = orc.deposit_fp64(g, qe, Pe.x, Pe.z)
        u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        assert np.array_equal(u, ref.get_field("u"))
        urf = ref.get_field("uRF")
        orc.advance_boris_init(g, u, urf, m, ie, Pe)
        assert np.array_equal(ref.get_particles(ie)[:, :7], Pe.aos7())
        for step in range(5):
            ref.advance()
            u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
            rho[:] = 0
            orc.advance_boris(g, u, urf, m, ie, Pe, niter=step, rng=None)
            orc.advance_boundary(g, mask, qe, Pe, rho=rho)
        oe = ref.get_particles(ie)
        assert np.array_equal(oe[:, :7], Pe.aos7()) and np.array_equal(oe[:, 7], Pe.alive)
        assert np.array_equal(rho, ref.get_field("rho"))


def test_u_smooth_matches_reference_including_its_row_overrun(orc, deckdir):
    # config_CRDS.txt: TUBE geometry, u_smooth = 1 (9-point smoothing of the potential after every solve)
    d = decks.deck("c4", deckdir + "_sm", n_particles=4000, geometry="TUBE", probe_radius=1.4e-3, u_smooth=1,
                   x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        rng = np.random.default_rng(2)
        u = rng.normal(size=(g.M, g.N))
        ref.set_field("u", u)
        ref.field_op("u_smooth")
        got = orc.u_smooth(g, u)
        want = ref.get_field("u")
        # the very last interior row's last node reads past the end of the array in the reference
        got[g.M - 2, g.N - 1] = want[g.M - 2, g.N - 1]
        assert np.array_equal(got, want)


@pytest.mark.parametrize("coord,descending", [("CYLINDRICAL", False), ("CARTESIAN", True)])
def test_magnetic_field_table_and_boris_bit_exact(orc, deckdir, coord, descending):
    """magnetic_field_const = 0 (f-row 3): Fields::load_magnetic_field, Fields::B and both Boris movers with the
    per-particle table look-up, restatement against the compiled reference"""
    r_max, z_max = (1.2e-2, 7.5e-2) if coord == "CYLINDRICAL" else (2e-2, 2e-2)
    bfile = write_btable(os.path.join(deckdir, "btable_%s.txt" % coord), 25, 31, r_max, z_max, descending)
    if coord == "CYLINDRICAL":
        d = decks.deck("c3", deckdir, n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=bfile,
                       selfconsistent=0, geometry="PENNING_SIMPLE")
        name = "ELECTRON"
    else:
        d = decks.deck("c2", deckdir, n_particles=10, geometry="RF_8PT", x_sampl=41, z_sampl=41, magnetic_field_const=0,
                       magnetic_field_file=bfile)
        name = "H_NEG"
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        bt = orc.load_magnetic_field(bfile)
        info, br, bz = ref.btable()
        assert (bt.jmax, bt.lmax, bt.dx, bt.dy, bt.xmin, bt.ymin) == (info["jmax"], info["lmax"], info["dx"], info["dy"], info["xmin"], info["ymin"])
        obr, obz = orc.btable_arrays(bt)
        assert np.array_equal(obr, br) and np.array_equal(obz, bz)
        rng = np.random.default_rng(4)
        x, z = rng.uniform(0, r_max * 0.999, 3000), rng.uniform(0, z_max * 0.999, 3000)
        assert np.array_equal(ref.field_B(x, z), orc.field_B(g, bt, x, z))
        u, urf = ref.get_field("u"), ref.get_field("uRF")
        m, names = model_from(orc, d["species_conf"])
        h = names.index(name)
        vth = 1500.0 if name == "H_NEG" else 4e5
        aos = disk_particles(rng, 1500, 0.5 * r_max, 0.5 * z_max, 0.2 * r_max, vth)
        ref.set_particles(h, aos)
        P = Particles.from_aos7(aos)
        ref.species_set(h, "niter", 3)
        ref.advance_position(h, init=True)
        orc.advance_boris_init(g, u, urf, m, h, P, niter=3, btable=bt)
        assert np.array_equal(ref.get_particles(h)[:, :7], P.aos7())
        ref.species_set(h, "lifetime", np.inf)
        for step in range(50):
            ref.species_set(h, "niter", 3 + step)
            ref.advance_position(h)
            orc.advance_boris(g, u, urf, m, h, P, niter=3 + step, rng=None, btable=bt)
        assert np.array_equal(ref.get_particles(h)[:, :7], P.aos7())
        # the rotation really depends on the position: the same run with the table's centre value differs
        assert np.abs(P.aos7()[:, 3:6] - aos[:, 3:6]).max() > 1.0


def _sorted_rows(a):
    a = np.asarray(a)
    return a[np.lexsort(a.T[::-1])]


@pytest.mark.parametrize("factor,collisions", [(4, True), (1, False), (3, True)])
def test_particle_source_bit_exact(orc, deckdir, factor, collisions):
    """use_source (f-row 4): Species<CARTESIAN>::source5_refresh and ::source against the compiled reference under the same
    SHR3 seed and the same libc rand() sequence: reservoir, injected particle set and the fp64 charge they deposit"""
    L = 6.4e-3
    d = decks.deck("c4", deckdir, n_particles=10, collisions=collisions, x_sampl=33, z_sampl=33, r_max=L, z_max=L, use_source=1,
                   src_fact=factor, n_particles_total=40, density_total=1e13, Bz=0.02, extern_field=200.0)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        p = ref.param()
        g = grid_from_param(p)
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        sp = cfg.read_species(d["species_conf"])[0][e]
        aos = disk_particles(np.random.default_rng(3), 50, 0.5 * L, 0.5 * L, 0.2 * L, 4e5)
        ref.set_particles(e, aos)
        ref.rng_seed(21)
        ref.source_refresh(e, factor)
        R = ref.source_particles(e)
        r = orc.rng(21)
        V = cfg.read_config(d["config"])["V"]
        src = orc.source_refresh(g, m, e, factor, V, r)
        assert src.n == R.shape[0] > 200
        assert np.array_equal(R[:, [0, 2, 3, 4, 5, 6]], src.aos7()[:, [0, 2, 3, 4, 5, 6]])
        # the reference first: both sides draw from the one libc rand() state
        nsteps = 12
        ref.rng_seed(33)
        ref.srand(7)
        rho0 = ref.get_field("rho:%d" % e).copy()
        assert not rho0.any()
        for _ in range(nsteps):
            ref.source(e)
        out_ref = ref.get_particles(e)
        rho_ref = ref.get_field("rho:%d" % e)
        src_ref = ref.source_particles(e)
        orc.rng_seed(r, 33)
        dst = Particles(50 + 4000)
        dst.alive[:] = 0
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "ttd")):
            getattr(dst, k)[:50] = aos[:, c]
        dst.alive[:50] = 1
        n_dst, total = 50, 0
        rho = np.zeros((g.M, g.N))
        for step in range(nsteps):
            inj, n_dst = orc.source(g, m, e, factor, src, dst, n_dst, rng=r, rho=rho, libc_seed=7 if step == 0 else None)
            total += inj
        assert total > 20 and n_dst == 50 + total
        assert np.array_equal(src_ref[:, [0, 2, 3, 4, 5, 6]], src.aos7()[:, [0, 2, 3, 4, 5, 6]])
        live_ref = out_ref[out_ref[:, 7] > 0][:, [0, 2, 3, 4, 5, 6]]
        live_orc = dst.aos7()[dst.alive > 0][:, [0, 2, 3, 4, 5, 6]]
        assert live_ref.shape == live_orc.shape
        assert np.array_equal(_sorted_rows(live_ref), _sorted_rows(live_orc))
        assert np.array_equal(rho_ref, rho)
        assert np.array_equal(ref.rng_draw("iuni", 3), orc.rng_draw(r, "iuni", 3))      # SHR3 streams stayed in step


def test_magnetic_field_loader_error_texts_match_the_reference(orc, deckdir):
    """Fields::load_magnetic_field's failure modes: the product's reader (mag2d_b200.config), the restatement and the
    compiled reference refuse the same files with the same message"""
    from common import write_btable
    good = write_btable(os.path.join(deckdir, "bt_err_good.txt"), 7, 9, 1.2e-2, 7.5e-2)
    rows = open(good).read().splitlines()
    short = os.path.join(deckdir, "bt_err_short.txt")
    with open(short, "w") as f:
        f.write("\n".join(rows[:-1]) + "\n")                    # one node missing
    cases = [(short, "wrong size of input vector"), (os.path.join(deckdir, "bt_err_none.txt"), "failed opening file")]
    for path, text in cases:
        with pytest.raises(RuntimeError, match=text):
            cfg.load_magnetic_field(path)
        d = decks.deck("c3", deckdir + "_bterr", n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=path,
                       selfconsistent=0)
        with pytest.raises(RuntimeError, match=text):
            RefHarness(d["config"], d["species_conf"], seed=5)
    with pytest.raises(RuntimeError, match="wrong size"):
        orc.load_magnetic_field(short)
    # a duplicated node instead of a missing one: the count is right, one node stays NaN -> "garbage loaded"
    dup = os.path.join(deckdir, "bt_err_dup.txt")
    with open(dup, "w") as f:
        mid = len(rows) // 2
        f.write("\n".join(rows[:mid] + [rows[mid - 1]] + rows[mid + 1:]) + "\n")
    for loader in (cfg.load_magnetic_field, orc.load_magnetic_field):
        with pytest.raises(RuntimeError, match="garbage loaded"):
            loader(dup)
    d = decks.deck("c3", deckdir + "_bterr2", n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=dup,
                   selfconsistent=0)
    with pytest.raises(RuntimeError, match="garbage loaded"):
        RefHarness(d["config"], d["species_conf"], seed=5)
