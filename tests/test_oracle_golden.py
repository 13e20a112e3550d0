"""Pin the CPU restatement against tests/golden/reference_v1.npz, the outputs of the unmodified
reference recorded by tests/golden/make_golden.py.  Needs neither /root/reference nor oracle/_ref."""
import os

import numpy as np
import pytest

from common import Particles, grid_from_param, model_from
from mag2d_b200 import config as cfg
from mag2d_b200 import decks

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_v1.npz"))


def test_rng_golden(orc):
    r = orc.rng(1234)
    for what in ("iuni", "uni", "rnor", "rexp", "radius"):
        assert np.array_equal(G["rng_" + what], orc.rng_draw(r, what, 4096)), what
    assert np.array_equal(G["rng_rot"], orc.rng_rot(r, 2.5, 64))
    assert np.array_equal(G["rng_deflect_out"], orc.rng_deflect(r, 0.3, G["rng_deflect_in"]))


def test_langevin_chi_known_answers(orc):
    # reference tests/test_langevin.cpp prints beta, chi, asymptote with 6 significant digits
    rows = G["langevin_chi"]
    for beta, chi, _ in rows:
        if not np.isfinite(chi):
            continue            # beta == 1: K(1) diverges, the reference prints nan
        assert abs(orc.lib.orc_langevin_chi(beta) - chi) <= 6e-6 * max(abs(chi), 1e-3), beta   # 6 printed digits
    # the values quoted in SURVEY.md §4
    for beta, chi in ((1.1, -0.709263), (2.0, -0.0381307), (5.0, -9.43303e-4), (9.9, -6.13247e-5)):
        assert abs(orc.lib.orc_langevin_chi(beta) - chi) <= 6e-6 * abs(chi)


def test_c1_model_scatter_multicoll_golden(orc, deckdir):
    d = decks.deck("c1", deckdir, n_particles=200)
    m, names = model_from(orc, d["species_conf"])
    e, he = names.index("ELECTRON"), names.index("HELIUM")
    assert m.lifetime(e) == G["c1_lifetime"][0]
    assert np.array_equal(m.rates(e), G["c1_rates"])
    sv = np.array([[m.sigma_v(e, he, k, v) for v in G["c1_sigma_v_v"]] for k in range(3)])
    assert np.array_equal(sv, G["c1_sigma_v"])
    r = orc.rng(77)
    out, proc, _ = orc.scatter(m, e, r, G["c1_scatter_in"])
    assert np.array_equal(out, G["c1_scatter_out"])
    p = cfg.read_config(d["config"])
    g = grid_from_param(p)
    mask, _ = orc.geometry(g, 0)
    P = Particles.from_aos7(G["c1_multicoll_in"])
    orc.rng_seed(r, 99)
    for _ in range(2):
        orc.advance_multicoll(0.0, p["extern_field"], m, e, P, r)
        orc.advance_boundary(g, mask, m.get(e, "charge"), P)
    assert np.array_equal(G["c1_multicoll_out"][:, :7], P.aos7())
    assert np.array_equal(G["c1_multicoll_out"][:, 7], P.alive)


@pytest.mark.parametrize("geo", ["RF_8PT", "RF_22PT"])
def test_c2_fields_gather_boris_golden(orc, deckdir, geo):
    d = decks.deck("c2", deckdir, n_particles=10, geometry=geo, x_sampl=41, z_sampl=41, Bt=0.01, Bz=0.02, Br=0.005)
    p = cfg.read_config(d["config"])
    g = grid_from_param(p)
    k = "c2_%s_" % geo
    mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
    assert np.array_equal(mask, G[k + "mask"])
    assert np.array_equal(np.where(mask < 2, volt, 0.0), G[k + "voltage"])
    zero = np.zeros((g.M, g.N))
    u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=False))
    urf = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=True))
    assert np.abs(u - G[k + "u"]).max() <= 1e-12 * max(1.0, np.abs(u).max())
    assert np.abs(urf - G[k + "uRF"]).max() <= 1e-12
    u, urf = G[k + "u"], G[k + "uRF"]
    x, z = G[k + "E_xz"]
    ex, ez = orc.field_E(g, u, urf, x, z, 3.3e-8)
    assert np.array_equal(np.stack([ex, ez]), G[k + "E"])
    assert np.array_equal(orc.is_free(g, mask, x[16:], z[16:]), G[k + "is_free"])
    m, names = model_from(orc, d["species_conf"])
    h = names.index("H_NEG")
    P = Particles.from_aos7(G[k + "boris_in"])
    orc.advance_boris_init(g, u, urf, m, h, P, niter=17)
    assert np.array_equal(P.aos7(), G[k + "boris_init"][:, :7])
    for step in range(100):
        orc.advance_boris(g, u, urf, m, h, P, niter=17 + step, rng=None)
        if step == 0:
            assert np.array_equal(P.aos7(), G[k + "boris_1"][:, :7])
    orc.advance_boundary(g, mask, m.get(h, "charge"), P)
    assert np.array_equal(P.aos7(), G[k + "boris_100"][:, :7])
    assert np.array_equal(P.alive, G[k + "boris_100"][:, 7])


def test_c4_selfconsistent_loop_golden(orc, deckdir):
    d = decks.deck("c4", deckdir, n_particles=1000, x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    p = cfg.read_config(d["config"])
    g = grid_from_param(p)
    m, names = model_from(orc, d["species_conf"])
    ii, ie = names.index("ARGON_POS"), names.index("ELECTRON")
    assert np.array_equal([m.lifetime(ii), m.lifetime(ie)], G["c4_lifetimes"])
    mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
    assert np.array_equal(mask, G["c4_mask"])
    Pi, Pe = Particles.from_aos7(G["c4_in_i"]), Particles.from_aos7(G["c4_in_e"])
    qi, qe = m.get(ii, "charge"), m.get(ie, "charge")
    rho_i, _ = orc.deposit_fp64(g, qi, Pi.x, Pi.z)
    rho_e, _ = orc.deposit_fp64(g, qe, Pe.x, Pe.z)
    rho = np.zeros_like(rho_i)
    rho += rho_i
    rho += rho_e
    assert np.array_equal(rho, G["c4_rho0"])
    u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
    assert np.array_equal(u, G["c4_u0"])
    urf = np.zeros_like(u)
    orc.advance_boris_init(g, u, urf, m, ii, Pi)
    orc.advance_boris_init(g, u, urf, m, ie, Pe)
    r = orc.rng(21)
    for step in range(5):
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
    assert np.array_equal(G["c4_out_i"][:, :7], Pi.aos7()) and np.array_equal(G["c4_out_i"][:, 7], Pi.alive)
    assert np.array_equal(G["c4_out_e"][:, :7], Pe.aos7()) and np.array_equal(G["c4_out_e"][:, 7], Pe.alive)
    assert np.array_equal(rho, G["c4_rho5"])
    assert np.array_equal(u, G["c4_u5"])


def test_c3_cylindrical_loop_golden(orc, deckdir):
    d = decks.deck("c3", deckdir, n_particles=500, x_sampl=41, z_sampl=51)
    p = cfg.read_config(d["config"])
    g = grid_from_param(p)
    m, names = model_from(orc, d["species_conf"])
    ie = names.index("ELECTRON")
    mask, volt = orc.geometry(g, int(p["geometry"]), p["probe_radius"], p["u_probe"])
    Pe = Particles.from_aos7(G["c3_in"])
    qe = m.get(ie, "charge")
    rho, _ = orc.deposit_fp64(g, qe, Pe.x, Pe.z)
    u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
    assert np.array_equal(u, G["c3_u0"])
    urf = np.zeros_like(u)
    orc.advance_boris_init(g, u, urf, m, ie, Pe)
    assert np.array_equal(G["c3_init"][:, :7], Pe.aos7())
    for step in range(5):
        u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        rho[:] = 0
        orc.advance_boris(g, u, urf, m, ie, Pe, niter=step, rng=None)
        orc.advance_boundary(g, mask, qe, Pe, rho=rho)
    assert np.array_equal(G["c3_out"][:, :7], Pe.aos7()) and np.array_equal(G["c3_out"][:, 7], Pe.alive)
    assert np.array_equal(rho, G["c3_rho5"])
    assert np.array_equal(u, G["c3_u5"])


def test_fixed_point_deposit_properties(orc):
    """the build's Q32 rule: bit-exact regardless of order, per-particle charge conserved to 2 ulps"""
    from oracle import OrcGrid
    g = OrcGrid.make(65, 49, 1.0e-2, 0.75e-2, selfconsistent=1)
    rng = np.random.default_rng(5)
    n = 20000
    x = rng.uniform(0, 1.0e-2 * (1 - 1e-12), n)
    z = rng.uniform(0, 0.75e-2 * (1 - 1e-12), n)
    a, bad = orc.deposit_fixed(g, x, z)
    assert bad == 0
    perm = rng.permutation(n)
    b, _ = orc.deposit_fixed(g, x[perm], z[perm])
    assert np.array_equal(a, b)
    assert abs(int(a.sum()) - n * 2 ** 32) <= 2 * n
    f, _ = orc.deposit_fp64(g, 1.0, x, z)
    assert np.abs(a * 2.0 ** -32 - f).max() <= n * 2.0 ** -33


def test_3d_field_code_golden():
    """oracle/mag3d_oracle.c against the fixture recorded from the reference's Field3D / Geometry / Solver
    (tests/golden/make_golden3d.py): deposit, interpolation and gradient bit-exact, geometry identical, solve 1e-11"""
    from oracle import Oracle3, Orc3Grid
    G3 = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference3d_v1.npz"))
    idx, idy, idz, x_max, y_max, z_max, mpf = G3["grid"]
    g = Orc3Grid.make(tuple(int(v) for v in G3["dims"]), idx, idy, idz, x_max, y_max, z_max, 0, mpf)
    orc3 = Oracle3()
    mask, volt = orc3.geometry(g)
    assert np.array_equal(mask, G3["mask"]) and np.array_equal(volt, G3["voltage"])
    x, y, z = (np.ascontiguousarray(G3["pos"][:, c]) for c in range(3))
    assert np.array_equal(orc3.is_free(g, mask, x, y, z), G3["is_free"])
    rho, bad = orc3.accumulate(g, -1.6e-19, x, y, z)
    assert bad == 0 and np.array_equal(rho, G3["rho"])
    assert np.array_equal(orc3.interpolate(g, G3["u_random"], x, y, z), G3["interp"])
    gx, gy, gz = (np.ascontiguousarray(G3["grad_pos"][:, c]) for c in range(3))
    assert np.array_equal(orc3.grad(g, G3["u_random"], gx, gy, gz), G3["grad"])
    b = orc3.rhs(g, mask, volt, rho)
    assert np.array_equal(b, G3["rhs"])
    u = orc3.solve_direct(g, mask, b)
    assert np.abs(u - G3["u_solved"]).max() <= 1e-11 * np.abs(G3["u_solved"]).max()


def test_magnetic_field_table_golden(orc, deckdir):
    """magnetic_field_const = 0: table build, Fields::B and the cylindrical Boris mover against the reference's
    outputs in reference_v2_btable.npz (make_golden_btable.py)"""
    from common import write_btable
    GB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_v2_btable.npz"))
    r_max, z_max = 1.2e-2, 7.5e-2
    bfile = write_btable(os.path.join(deckdir, "btable_golden.txt"), 25, 31, r_max, z_max)
    d = decks.deck("c3", deckdir, n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=bfile,
                   selfconsistent=0, geometry="PENNING_SIMPLE")
    p = cfg.read_config(d["config"])
    g = grid_from_param(p)
    bt = orc.load_magnetic_field(bfile)
    assert np.array_equal(np.array([bt.jmax, bt.lmax, bt.dx, bt.dy, bt.xmin, bt.ymin], dtype=np.float64), GB["bt_info"])
    br, bz = orc.btable_arrays(bt)
    assert np.array_equal(br, GB["bt_Br"]) and np.array_equal(bz, GB["bt_Bz"])
    assert np.array_equal(orc.field_B(g, bt, GB["bt_x"], GB["bt_z"]), GB["bt_B"])
    m, names = model_from(orc, d["species_conf"])
    e = names.index("ELECTRON")
    P = Particles.from_aos7(GB["bt_in"])
    orc.advance_boris_init(g, GB["bt_u"], GB["bt_uRF"], m, e, P, niter=3, btable=bt)
    assert np.array_equal(GB["bt_init"][:, :7], P.aos7())
    for step in range(20):
        orc.advance_boris(g, GB["bt_u"], GB["bt_uRF"], m, e, P, niter=3 + step, rng=None, btable=bt)
    assert np.array_equal(GB["bt_out"][:, :7], P.aos7())
