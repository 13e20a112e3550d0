"""GPU parity tests: the CUDA path (through the C ABI, mag2d_b200.api.Sim) against the CPU oracle
(oracle/mag2d_oracle.c), against the committed golden fixtures recorded from the unmodified reference,
and — when oracle/_ref travelled to this box — against the reference itself.

Tolerances (BASELINE.json north_star / SURVEY.md §8c):
  trajectories, collisions off, fp64:  <= 1e-12 relative after 1 step, <= 1e-10 after 100 steps
  deposited grid:                      bit-exact against the fixed-point restatement
  Poisson solve:                       ||u - u_direct||_inf / ||u_direct||_inf <= 1e-8
"""
import os

import numpy as np
import pytest

from common import Particles, disk_particles, grid_from_param, model_from, needs_ref
from mag2d_b200 import decks

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_v1.npz"))


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def relerr(a, b):
    """max |a-b| per column relative to the column's magnitude"""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return (np.abs(a - b).max(axis=0) / scale).max()


# ----------------------------------------------------------------------------------------- fields
@pytest.mark.parametrize("deck,kw", [
    ("c2", dict(geometry="RF_8PT", x_sampl=41, z_sampl=41)),
    ("c2", dict(geometry="RF_22PT", x_sampl=200, z_sampl=200)),
    ("c2", dict(geometry="RF_QUAD", x_sampl=64, z_sampl=50)),
    ("c2", dict(geometry="PROBE", x_sampl=101, z_sampl=101, probe_radius=2e-3, u_probe=-7.0)),
    ("c3", dict(geometry="PENNING_SIMPLE", x_sampl=61, z_sampl=81, selfconsistent=0)),
    ("c3", dict(geometry="EMPTY", x_sampl=200, z_sampl=100, selfconsistent=0)),
])
def test_vacuum_solve_matches_direct_solver(orc, deckdir, deck, kw):
    d = decks.deck(deck, deckdir, n_particles=10, **kw)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        mask, volt = orc.geometry(g, int(sim.param["geometry"]), sim.param["probe_radius"], sim.param["u_probe"])
        assert np.array_equal(mask, sim.mask)
        zero = np.zeros((g.M, g.N))
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=False))
        urf_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, zero, rf=True))
        u, urf = sim.get_field("u"), sim.get_field("uRF")
        assert np.abs(u - u_ref).max() <= 1e-8 * max(np.abs(u_ref).max(), 1e-30), sim.solve_info
        assert np.abs(urf - urf_ref).max() <= 1e-8 * max(np.abs(urf_ref).max(), 1e-30), sim.solve_info
        assert sim.solve_info["u"]["cycles"] <= 40


@pytest.mark.parametrize("coord_deck,M,N", [("c4", 33, 34), ("c4", 34, 35), ("c4", 21, 99), ("c4", 20, 130), ("c3", 30, 67)])
def test_folded_direct_solver_every_parity_and_padding(orc, deckdir, coord_deck, M, N):
    """the direct solver folds every row into its symmetric and antisymmetric halves (poisson_direct.cu): even and odd interior
    column counts (a middle column that is its own mirror image), half lengths below, at and above one 32-wide tile, Cartesian
    and cylindrical operator, with charge"""
    kw = dict(x_sampl=M, z_sampl=N, r_max=1e-4 * (M - 1), z_max=1e-4 * (N - 1))
    if coord_deck == "c3":
        kw.update(geometry="EMPTY")
    d = decks.deck(coord_deck, deckdir + "_fold%d_%d" % (M, N), n_particles=4000, **kw)
    with _sim(d["config"], d["species_conf"]) as sim:
        assert sim.solver_is_direct()
        g = grid_from_param(sim.param)
        ie = sim.species_index("ELECTRON")
        rng = np.random.default_rng(M * 1000 + N)
        a = np.zeros((4000, 7))
        a[:, 0] = rng.uniform(0.05, 0.95, 4000) * g.x_max
        a[:, 2] = rng.uniform(0.05, 0.95, 4000) * g.z_max
        sim.set_particles(ie, a)
        sim.species_accumulate(ie)
        info = sim.solve(rf=False)
        rho = sim.get_field("rho")
        mask, volt = orc.geometry(g, int(sim.param["geometry"]))
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        assert np.abs(sim.get_field("u") - u_ref).max() <= 1e-8 * np.abs(u_ref).max(), info


def test_solve_with_charge_matches_direct_solver(orc, deckdir):
    d = decks.deck("c4", deckdir, n_particles=40000, x_sampl=65, z_sampl=65, r_max=6.4e-3, z_max=6.4e-3)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        ie = sim.species_index("ELECTRON")
        ii = sim.species_index("ARGON_POS")
        rng = np.random.default_rng(5)
        ae = disk_particles(rng, 20000, 3.2e-3, 3.2e-3, 2.5e-3, 6e5)
        ai = disk_particles(rng, 20000, 3.0e-3, 3.3e-3, 2.0e-3, 300.0)
        sim.set_particles(ie, ae)
        sim.set_particles(ii, ai)
        sim.species_accumulate(ie)
        sim.species_accumulate(ii)
        info = sim.solve(rf=False)
        rho = sim.get_field("rho")
        mask, volt = orc.geometry(g, 0)
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        assert np.abs(sim.get_field("u") - u_ref).max() <= 1e-8 * np.abs(u_ref).max(), info


@pytest.mark.parametrize("deck,kw,species", [
    ("c4", dict(x_sampl=80, z_sampl=65, r_max=7.9e-3, z_max=6.4e-3), ("ELECTRON", "ARGON_POS")),
    ("c3", dict(geometry="EMPTY", x_sampl=97, z_sampl=130), None),
])
def test_direct_and_multigrid_solvers_agree_with_oracle(orc, deckdir, deck, kw, species):
    """grids whose electrodes are whole rows use the sine-transform x tridiagonal direct solver (auto);
    it has to reproduce the reference's LU to round-off, and the multigrid forced on the same system to 1e-8"""
    d = decks.deck(deck, deckdir, n_particles=40000, **kw)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        assert sim.solver_is_direct()
        rng = np.random.default_rng(11)
        names = species or [n for n in d["species"]]
        for q, name in enumerate(names):
            a = disk_particles(rng, 20000, 0.45 * g.x_max, 0.5 * g.z_max, 0.3 * min(g.x_max, g.z_max), 300.0 * (1 + 100 * q))
            i = sim.species_index(name)
            sim.set_particles(i, a)
            sim.species_accumulate(i)
        mask, volt = orc.geometry(g, int(sim.param["geometry"]), sim.param["probe_radius"], sim.param["u_probe"])
        rho = sim.get_field("rho")
        assert np.abs(rho).max() > 0
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        info = sim.solve(rf=False)
        u_direct = sim.get_field("u")
        assert info["cycles"] == 0 and info["resid"] < 1e-10, info
        assert np.abs(u_direct - u_ref).max() <= 1e-9 * np.abs(u_ref).max(), info
        sim.set_solver_kind("multigrid")
        assert not sim.solver_is_direct()
        sim.set_field("u", np.zeros_like(u_ref))
        info = sim.solve(rf=False)
        assert info["cycles"] > 0
        assert np.abs(sim.get_field("u") - u_ref).max() <= 1e-8 * np.abs(u_ref).max(), info
    d = decks.deck("c2", deckdir, n_particles=10, geometry="PROBE", x_sampl=41, z_sampl=41, probe_radius=2e-3)
    with _sim(d["config"], d["species_conf"]) as sim:
        assert not sim.solver_is_direct()
        with pytest.raises(Exception, match="does not separate"):
            sim.set_solver_kind("direct")


def test_gather_matches_oracle_and_golden(orc, deckdir):
    d = decks.deck("c2", deckdir, n_particles=10, geometry="RF_8PT", x_sampl=41, z_sampl=41, Bt=0.01, Bz=0.02, Br=0.005)
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        k = "c2_RF_8PT_"
        sim.set_field("u", G[k + "u"])
        sim.set_field("uRF", G[k + "uRF"])
        x, z = G[k + "E_xz"]
        ex, ez = sim.field_E(x, z, 3.3e-8)
        scale = np.abs(G[k + "E"]).max()
        assert np.abs(np.stack([ex, ez]) - G[k + "E"]).max() <= 1e-12 * scale


# ----------------------------------------------------------------------------------- trajectories
@pytest.mark.parametrize("geo", ["RF_8PT", "RF_22PT"])
def test_boris_cartesian_trajectory_vs_golden_reference(deckdir, geo):
    # the golden run forced lifetime = inf on the reference side; here the gas densities are zero
    d2 = decks.deck("c2", deckdir + "_nocoll", n_particles=10, collisions=False, geometry=geo, x_sampl=41, z_sampl=41,
                    Bt=0.01, Bz=0.02, Br=0.005)
    with _sim(d2["config"], d2["species_conf"], presolve=False) as sim2:
        k = "c2_%s_" % geo
        h = sim2.species_index("H_NEG")
        assert not np.isfinite(sim2.species_get(h, "lifetime"))
        sim2.set_field("u", G[k + "u"])
        sim2.set_field("uRF", G[k + "uRF"])
        # the golden run started at niter = 17 (RF phase): advance the species clock on an empty store
        for _ in range(17):
            sim2.species_advance(h)
        assert sim2.species_get(h, "niter") == 17
        sim2.set_particles(h, G[k + "boris_in"])
        sim2.species_advance_init(h)
        out = sim2.get_particles(h)
        assert relerr(out[:, 3:6], G[k + "boris_init"][:, 3:6]) <= 1e-12
        sim2.species_advance(h)
        out = sim2.get_particles(h)
        ref1 = G[k + "boris_1"]
        # after one step nothing has been removed yet in the reference either (boundary ran only at the end)
        alive = out[:, 7] > 0
        assert relerr(out[alive][:, [0, 2, 3, 4, 5]], ref1[alive][:, [0, 2, 3, 4, 5]]) <= 1e-12
        for _ in range(99):
            sim2.species_advance(h)
        out = sim2.get_particles(h)
        ref100 = G[k + "boris_100"]
        both = (out[:, 7] > 0) & (ref100[:, 7] > 0)
        assert both.sum() >= 0.9 * (ref100[:, 7] > 0).sum()
        assert relerr(out[both][:, [0, 2, 3, 4, 5]], ref100[both][:, [0, 2, 3, 4, 5]]) <= 1e-10
        # removal decisions agree except for particles within rounding of an electrode cell edge
        assert np.sum((out[:, 7] > 0) != (ref100[:, 7] > 0)) <= 2


def test_boris_cartesian_vs_oracle_large(orc, deckdir):
    d = decks.deck("c2", deckdir + "_nc", n_particles=10, collisions=False, geometry="RF_22PT", x_sampl=200, z_sampl=200, Bz=0.05)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        h = names.index("H_NEG")
        u, urf = sim.get_field("u"), sim.get_field("uRF")
        aos = disk_particles(np.random.default_rng(9), 20000, 1e-2, 1e-2, 7.8e-3, 1500.0)   # reaches into the rods
        sim.set_particles(h, aos)
        P = Particles.from_aos7(aos)
        mask = sim.mask
        for step in range(20):
            sim.species_advance(h)
            orc.advance_boris(g, u, urf, m, h, P, niter=step, rng=None)
            orc.advance_boundary(g, mask, m.get(h, "charge"), P)
        out = sim.get_particles(h)
        both = (out[:, 7] > 0) & (P.alive > 0)
        assert np.sum((out[:, 7] > 0) != (P.alive > 0)) <= 3
        assert 0 < both.sum() < 20000          # some ions hit the rods
        assert relerr(out[both][:, [0, 2, 3, 4, 5]], P.aos7()[both][:, [0, 2, 3, 4, 5]]) <= 1e-11


def test_cylindrical_selfconsistent_loop_vs_golden_reference(orc, deckdir):
    d = decks.deck("c3", deckdir, n_particles=500, x_sampl=41, z_sampl=51)
    with _sim(d["config"], d["species_conf"]) as sim:
        ie = sim.species_index("ELECTRON")
        sim.set_particles(ie, G["c3_in"])
        sim.advance_init()
        u0 = sim.get_field("u")
        assert np.abs(u0 - G["c3_u0"]).max() <= 1e-8 * np.abs(G["c3_u0"]).max()
        out = sim.get_particles(ie)
        assert relerr(out[:, 3:6], G["c3_init"][:, 3:6]) <= 1e-9     # E from an iterative solve: 1e-8 * |E| dt q/m
        sim.advance(5)
        out = sim.get_particles(ie)
        ref = G["c3_out"]
        assert np.array_equal(out[:, 7], ref[:, 7])
        assert relerr(out[:, [0, 2, 3, 4, 5]], ref[:, [0, 2, 3, 4, 5]]) <= 1e-8
        assert np.abs(sim.get_field("u") - G["c3_u5"]).max() <= 1e-7 * np.abs(G["c3_u5"]).max()
        # deposit: bit-exact against the fixed-point restatement on the device's own positions
        g = grid_from_param(sim.param)
        fixed, _ = orc.deposit_fixed(g, out[:, 0].copy(), out[:, 2].copy(), out[:, 7].astype(np.uint8))
        assert np.array_equal(sim.rho_fixed(ie), fixed)
        qe = sim.species[ie]["charge"]
        assert np.abs(sim.get_field("rho") - G["c3_rho5"]).max() <= 500 * 2.0 ** -33 * abs(qe) + 1e-6 * abs(qe)


@pytest.mark.parametrize("deposit", ["fixed", "fp64"])
def test_cartesian_selfconsistent_loop_vs_oracle(orc, deckdir, deposit):
    """two particle species, deposit + solve every step, collisions off (gas density 0).

    deposit = "fixed": the oracle loop deposits with the same Q32 fixed-point rule as the product (the bit-exactness contract
    of the charge grid), so the two loops see the same right-hand side and differ only by the round-off of the solve and of the
    push: trajectories within 1e-10 after 5 coupled steps, like the collision-free bar.
    deposit = "fp64": the oracle deposits like the reference (sequential fp64 sums).  Each weight of the product is rounded at
    2^-32 (2.3e-10), the charge of a node is a difference of two species' sums that nearly cancel (quasi-neutral plasma), so the
    right-hand side differs by ~1e-8 of its size, the field by as much, and the trajectories after 5 steps by up to 1e-7 of the
    largest coordinate: that is the price of an order-independent grid, not solver or push error (the "fixed" leg shows it)."""
    d = decks.deck("c4", deckdir + "_nc", n_particles=20000, collisions=False, x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        ii, ie = names.index("ARGON_POS"), names.index("ELECTRON")
        mask, volt = orc.geometry(g, 0)
        rng = np.random.default_rng(7)
        ai = disk_particles(rng, 10000, 1.6e-3, 1.6e-3, 1.4e-3, 300.0)
        ae = disk_particles(rng, 10000, 1.7e-3, 1.6e-3, 1.4e-3, 6e5)
        sim.set_particles(ii, ai)
        sim.set_particles(ie, ae)
        Pi, Pe = Particles.from_aos7(ai), Particles.from_aos7(ae)
        qi, qe = m.get(ii, "charge"), m.get(ie, "charge")
        sim.advance_init()

        def coulombs(q, P):
            if deposit == "fp64":
                return orc.deposit_fp64(g, q, P.x, P.z, P.alive.astype(np.uint8))[0]
            return q * (orc.deposit_fixed(g, P.x, P.z, P.alive.astype(np.uint8))[0].astype(np.float64) * 2.0 ** -32)
        rho = coulombs(qi, Pi) + coulombs(qe, Pe) if deposit == "fp64" else None
        if deposit == "fixed":
            # the product sums q_s W_s 2^-32 over the species in index order (k_rhs): so does this
            rho = coulombs(qi, Pi) + coulombs(qe, Pe) if ii < ie else coulombs(qe, Pe) + coulombs(qi, Pi)
        u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        urf = np.zeros_like(u)
        orc.advance_boris_init(g, u, urf, m, ii, Pi)
        orc.advance_boris_init(g, u, urf, m, ie, Pe)
        for step in range(5):
            sim.advance(1)
            u = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
            orc.advance_boris(g, u, urf, m, ii, Pi, niter=step, rng=None)
            orc.advance_boundary(g, mask, qi, Pi)
            orc.advance_boris(g, u, urf, m, ie, Pe, niter=step, rng=None)
            orc.advance_boundary(g, mask, qe, Pe)
            rho = coulombs(qi, Pi) + coulombs(qe, Pe) if ii < ie else coulombs(qe, Pe) + coulombs(qi, Pi)
        oi, oe = sim.get_particles(ii), sim.get_particles(ie)
        assert np.array_equal(oi[:, 7], Pi.alive) and np.array_equal(oe[:, 7], Pe.alive)
        tol = 1e-10 if deposit == "fixed" else 1e-7
        assert relerr(oe[:, [0, 2, 3, 4, 5]], Pe.aos7()[:, [0, 2, 3, 4, 5]]) <= tol
        assert relerr(oi[:, [0, 2, 3, 4, 5]], Pi.aos7()[:, [0, 2, 3, 4, 5]]) <= tol
        # fixed-point grids are bit-exact for the device's own particle set
        for s, o in ((ii, oi), (ie, oe)):
            fixed, _ = orc.deposit_fixed(g, o[:, 0].copy(), o[:, 2].copy(), o[:, 7].astype(np.uint8))
            assert np.array_equal(sim.rho_fixed(s), fixed)
        if deposit == "fp64":
            assert np.abs(sim.get_field("rho") - rho).max() <= 20000 * 2.0 ** -33 * abs(qe) + 1e-6 * abs(qe)


# ---------------------------------------------------------------------------------------- deposit
def test_deposit_bit_exact_and_order_independent(orc, deckdir):
    d = decks.deck("c4", deckdir, n_particles=200000, x_sampl=129, z_sampl=97, r_max=1.28e-2, z_max=0.96e-2)
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        g = grid_from_param(sim.param)
        ie = sim.species_index("ELECTRON")
        rng = np.random.default_rng(11)
        n = 200000
        x = rng.uniform(0, 1.28e-2, n)
        z = rng.uniform(0, 0.96e-2, n)
        x[:3] = [0.0, 1.28e-2, 1.28e-2]        # corners and the x == x_max edge
        z[:3] = [0.0, 0.96e-2, 0.0]
        v = np.zeros(n)
        sim.add_particles_soa(ie, x, z, v, v, v)
        sim.species_accumulate(ie)
        a = sim.rho_fixed(ie)
        ref, bad = orc.deposit_fixed(g, x, z)
        assert bad == 0 and np.array_equal(a, ref)
        assert abs(int(a.sum()) - n * 2 ** 32) <= 2 * n          # charge conserved to 2 ulps per particle
        # any order / any partition over "ranks": integer sums are associative
        sim.rho_reset()
        sim.L.mag2d_particles_clear(sim.h, ie)
        perm = rng.permutation(n)
        sim.add_particles_soa(ie, x[perm], z[perm], v, v, v)
        sim.species_accumulate(ie)
        assert np.array_equal(sim.rho_fixed(ie), ref)
        halves = [orc.deposit_fixed(g, x[s], z[s])[0] for s in (slice(0, n // 3), slice(n // 3, n))]
        assert np.array_equal(halves[0] + halves[1], ref)
        # distance to the reference's sequential fp64 grid
        f, _ = orc.deposit_fp64(g, 1.0, x, z)
        assert np.abs(a * 2.0 ** -32 - f).max() <= 200 * 2.0 ** -33


# --------------------------------------------------------------------------- boundary / store / sort
def test_periodic_wrap_and_free_removal(orc, deckdir):
    for boundary in ("PERIODIC", "FREE"):
        d = decks.deck("c1", deckdir, n_particles=100, collisions=False, boundary=boundary, mover="ADVANCE_BORIS",
                       x_sampl=5, z_sampl=5, extern_field=0.0)
        with _sim(d["config"], d["species_conf"]) as sim:
            g = grid_from_param(sim.param)
            m, names = model_from(orc, d["species_conf"])
            e = names.index("ELECTRON")
            rng = np.random.default_rng(2)
            n = 4000
            aos = np.zeros((n, 7))
            aos[:, 0] = rng.uniform(0, 2e-2, n)
            aos[:, 2] = rng.uniform(0, 2e-2, n)
            aos[:, 3:6] = rng.normal(size=(n, 3)) * 4e5      # v*dt = 4 mm: many leave the 2 cm box
            sim.set_particles(e, aos)
            P = Particles.from_aos7(aos)
            u = np.zeros((g.M, g.N))
            for step in range(3):
                sim.species_advance(e)
                orc.advance_boris(g, u, u, m, e, P, niter=step, rng=None)
                orc.advance_boundary(g, sim.mask, m.get(e, "charge"), P)
            out = sim.get_particles(e)
            assert np.array_equal(out[:, 7], P.alive)
            alive = P.alive > 0
            assert relerr(out[alive][:, [0, 2, 3, 5]], P.aos7()[alive][:, [0, 2, 3, 5]]) <= 1e-12
            n_alive, n_slots = sim.count(e)
            assert n_alive == alive.sum()
            if boundary == "FREE":
                assert n_alive < n


def test_cell_sort_compacts_and_preserves_the_particle_set(orc, deckdir):
    d = decks.deck("c4", deckdir, n_particles=100000, x_sampl=65, z_sampl=49, r_max=6.4e-3, z_max=4.8e-3)
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        g = grid_from_param(sim.param)
        ie = sim.species_index("ELECTRON")
        rng = np.random.default_rng(3)
        n = 100000
        x = rng.uniform(0, 6.4e-3, n)
        z = rng.uniform(0, 4.8e-3, n)
        x[::7] = np.nan                                  # removed particles keep their slot until the sort
        vx, vy, vz = rng.normal(size=(3, n))
        sim.add_particles_soa(ie, x, z, vx, vy, vz)
        sim.species_accumulate(ie)
        before = sim.rho_fixed(ie)
        sim.sort(ie)
        n_alive, n_slots = sim.count(ie)
        live = ~np.isnan(x)
        assert n_alive == live.sum() and n_slots == n_alive        # compacted
        out = sim.get_particles(ie)
        assert np.all(out[:, 7] == 1)
        key = (np.minimum((out[:, 0] * g.idx).astype(int), g.M - 2) * (g.N - 1)
               + np.minimum((out[:, 2] * g.idz).astype(int), g.N - 2))
        assert np.all(np.diff(key) >= 0)                         # cell-sorted
        a = np.stack([x, z, vx, vy, vz], axis=1)[live]
        b = out[:, [0, 2, 3, 4, 5]]
        assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])   # same multiset, bit for bit
        sim.rho_reset()
        sim.species_accumulate(ie)
        assert np.array_equal(sim.rho_fixed(ie), before)         # deposit does not depend on particle order


# ------------------------------------------------------------------------------- reference, live
@pytest.mark.parametrize("interval", [1, 2, 3])
def test_fused_cell_sort_keeps_the_particle_set_and_the_charge(orc, deckdir, interval):
    """the Boris push carries the cell sort (COUNT step -> tickets, PERMUTE step -> sorted slots of the other slab):
    with collisions off the physics does not depend on the slot order, so after N steps the multiset of particles and
    the fixed-point charge grids must be bit-identical to a run that never sorts; dead slots must be compacted away"""
    d = decks.deck("c4", deckdir, n_particles=30000, collisions=False, x_sampl=33, z_sampl=49, r_max=3.2e-3, z_max=4.8e-3)
    rng = np.random.default_rng(21)
    init = {}
    results = []
    for k in (0, interval):
        with _sim(d["config"], d["species_conf"]) as sim:
            for name, vth in (("ARGON_POS", 4e5), ("ELECTRON", 8e5)):
                i = sim.species_index(name)
                if name not in init:
                    # uniform over the whole box and fast: particles leave through the walls every step, so the
                    # compaction is exercised
                    a = np.zeros((15000, 7))
                    a[:, 0] = rng.uniform(1e-7, 3.2e-3 - 1e-7, 15000)
                    a[:, 2] = rng.uniform(1e-7, 4.8e-3 - 1e-7, 15000)
                    a[:, 3:6] = rng.normal(size=(15000, 3)) * vth
                    init[name] = a
                sim.set_particles(i, init[name])
            sim.set_sort_interval(k)
            sim.advance_init()
            sim.advance(7)
            out = {}
            for name in ("ARGON_POS", "ELECTRON"):
                i = sim.species_index(name)
                p = sim.get_particles(i)
                alive = p[p[:, 7] > 0][:, [0, 2, 3, 4, 5]]
                order = np.lexsort(alive.T[::-1])
                out[name] = (alive[order], sim.rho_fixed(i), sim.count(i), p[:, 7])
            results.append(out)
    for name in ("ARGON_POS", "ELECTRON"):
        a, b = results[0][name], results[1][name]
        assert a[2][0] == b[2][0] and a[2][0] < 15000          # same survivors, and some particles did leave
        assert np.array_equal(a[0], b[0]), name
        assert np.array_equal(a[1], b[1]), name
        # the permuting steps drop dead slots: only the removals since the last permute are still holes
        flags = b[3]
        n_alive = int(flags.sum())
        n_dead = 15000 - n_alive
        assert (flags[:n_alive] == 0).sum() < 0.6 * n_dead
    with _sim(d["config"], d["species_conf"]) as sim:
        i = sim.species_index("ELECTRON")
        sim.set_particles(i, init["ELECTRON"])
        sim.set_sort_interval(interval)
        sim.advance_init()
        sim.advance(2 * interval + 1)
        p = sim.get_particles(i)
        live = p[p[:, 7] > 0]
        g = grid_from_param(sim.param)
        key = np.floor(live[:, 0] * g.idx).astype(np.int64) * (g.N - 1) + np.floor(live[:, 2] * g.idz).astype(np.int64)
        assert (np.diff(key) < 0).mean() < 0.35      # an unsorted store has ~0.5


@pytest.mark.parametrize("interval", [2, 3])
def test_per_species_pushes_between_steps_do_not_corrupt_the_fused_sort(deckdir, interval):
    """a COUNT push leaves per-cell cursors for the next PERMUTE push; a per-species push in between
    (mag2d_species_advance, the host layer's Species::advance) moves particles away from the cells they were counted
    in, so the cursors must be dropped — otherwise the permute overflows cells and loses / duplicates particles.
    Also: the interval switched off and on again between a COUNT and its PERMUTE."""
    d = decks.deck("c4", deckdir, n_particles=30000, collisions=False, x_sampl=33, z_sampl=49, r_max=3.2e-3, z_max=4.8e-3)
    rng = np.random.default_rng(77)
    init = {}
    for name, vth in (("ARGON_POS", 4e5), ("ELECTRON", 8e5)):
        a = np.zeros((15000, 7))
        a[:, 0] = rng.uniform(1e-7, 3.2e-3 - 1e-7, 15000)
        a[:, 2] = rng.uniform(1e-7, 4.8e-3 - 1e-7, 15000)
        a[:, 3:6] = rng.normal(size=(15000, 3)) * vth
        init[name] = a
    results = []
    for k in (0, interval):
        with _sim(d["config"], d["species_conf"]) as sim:
            idx = [sim.species_index(n) for n in ("ARGON_POS", "ELECTRON")]
            for n, i in zip(("ARGON_POS", "ELECTRON"), idx):
                sim.set_particles(i, init[n])
            sim.set_sort_interval(k)
            sim.advance_init()
            for _ in range(4):
                sim.advance(1)                   # COUNT (first step of a period)
                for i in idx:
                    sim.species_advance(i)       # plain push between COUNT and PERMUTE
                sim.advance(1)
            sim.set_sort_interval(0)             # interval off ...
            sim.advance(1)
            sim.set_sort_interval(k)             # ... and on again
            sim.advance(3)
            out = {}
            for n, i in zip(("ARGON_POS", "ELECTRON"), idx):
                p = sim.get_particles(i)
                alive = p[p[:, 7] > 0][:, [0, 2, 3, 4, 5]]
                out[n] = (alive[np.lexsort(alive.T[::-1])], sim.count(i)[0])
            results.append(out)
    for n in ("ARGON_POS", "ELECTRON"):
        assert results[0][n][1] == results[1][n][1] and 0 < results[0][n][1] < 15000
        assert np.array_equal(results[0][n][0], results[1][n][0]), n


def test_streamed_step_with_host_resident_particles_equals_the_resident_step(orc, deckdir):
    """mag2d_step_streamed pushes host SoA arrays through device staging buffers chunk by chunk; with collisions off
    it must leave exactly the particles, charge grids and potential of mag2d_step on a device-resident store"""
    d = decks.deck("c4", deckdir, n_particles=30000, collisions=False, x_sampl=33, z_sampl=49, r_max=3.2e-3, z_max=4.8e-3)
    rng = np.random.default_rng(31)
    init = {}
    for name, vth in (("ARGON_POS", 4e5), ("ELECTRON", 8e5)):
        a = np.zeros((7000, 7))
        a[:, 0] = rng.uniform(1e-7, 3.2e-3 - 1e-7, 7000)
        a[:, 2] = rng.uniform(1e-7, 4.8e-3 - 1e-7, 7000)
        a[:, 3:6] = rng.normal(size=(7000, 3)) * vth
        init[name] = a
    with _sim(d["config"], d["species_conf"]) as ref, _sim(d["config"], d["species_conf"]) as sim:
        idx = [sim.species_index(n) for n in ("ARGON_POS", "ELECTRON")]
        for s_, name in zip(idx, ("ARGON_POS", "ELECTRON")):
            ref.set_particles(s_, init[name])
            sim.set_particles(s_, init[name])
        ref.set_sort_interval(0)
        ref.advance_init()
        sim.advance_init()
        host = {}
        for s_ in idx:
            p = sim.get_particles(s_)
            cols = [np.ascontiguousarray(p[:, c]) for c in (0, 2, 3, 4, 5)]
            cols[0][p[:, 7] == 0] = np.nan
            host[s_] = cols
            sim._chk(sim.L.mag2d_particles_clear(sim.h, s_))
        for _ in range(4):
            ref.advance(1)
            # 2048-slot chunks: 4 chunks per species, so the staging ring wraps
            sim.step_streamed(idx, [7000, 7000], [[c.ctypes.data for c in host[s_]] for s_ in idx], chunk_slots=2048)
        for s_ in idx:
            want = ref.get_particles(s_)
            alive = want[:, 7] > 0
            assert np.array_equal(~np.isnan(host[s_][0]), alive)
            assert 0 < alive.sum() < 7000
            for c, col in zip((0, 2, 3, 4, 5), host[s_]):
                assert np.array_equal(col[alive], want[alive, c])
            assert np.array_equal(sim.rho_fixed(s_), ref.rho_fixed(s_))
        assert np.array_equal(sim.get_field("u"), ref.get_field("u"))


def test_streamed_step_leaves_vy_in_pinned_host_memory(deckdir):
    """2-D Cartesian, B = 0: the push never reads the out-of-plane velocity, so a streamed step does not copy a PINNED vy array at
    all — the collision pass reaches the few elements it needs in place over PCIe.  Checked here: the bytes the step copies
    (four arrays instead of five), vy untouched bit for bit for every particle that did not collide, changed for a fraction
    1 - exp(-dt / lifetime) of them (5 sigma), and the same particles / charge as with the staged copy when collisions are off."""
    import torch
    n = 60000
    rng = np.random.default_rng(77)
    out = {}
    for collisions in (False, True):
        d = decks.deck("c4", deckdir + "_zc%d" % collisions, n_particles=2 * n, collisions=collisions, x_sampl=33, z_sampl=49, r_max=3.2e-3,
                       z_max=4.8e-3)
        for pinned in (True, False):
            with _sim(d["config"], d["species_conf"]) as sim:
                e = sim.species_index("ELECTRON")
                a = np.zeros((n, 5))
                r2 = np.random.default_rng(5)
                a[:, 0] = r2.uniform(1e-7, 3.2e-3 - 1e-7, n)
                a[:, 1] = r2.uniform(1e-7, 4.8e-3 - 1e-7, n)
                a[:, 2:5] = r2.normal(size=(n, 3)) * 8e5
                sim.advance_init()
                cols = [torch.from_numpy(np.ascontiguousarray(a[:, c])) for c in range(5)]
                if pinned:
                    cols = [t.pin_memory() for t in cols]
                vy0 = cols[3].numpy().copy()
                sim.streamed_bytes(reset=True)
                sim.step_streamed([e], [n], [[t.data_ptr() for t in cols]], chunk_slots=16384)
                h2d, d2h = sim.streamed_bytes()
                assert h2d == d2h == (4 if pinned else 5) * 8 * n
                vy1 = cols[3].numpy()
                changed = vy1 != vy0
                if collisions:
                    prob = sim.species_get(e, "prob")
                    assert abs(changed.sum() - n * prob) <= 5 * np.sqrt(n * prob) + 5, (changed.sum(), n * prob)
                else:
                    assert not changed.any()
                out[(collisions, pinned)] = ([t.numpy().copy() for t in cols], sim.rho_fixed(e))
    a, b = out[(False, True)], out[(False, False)]
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y, equal_nan=True)
    assert np.array_equal(a[1], b[1])
    # with collisions the two runs draw the same Philox streams (keyed by slot and chunk): identical, whichever way vy travels
    a, b = out[(True, True)], out[(True, False)]
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.skipif(not needs_ref, reason="oracle/_ref not present on this machine")
def test_live_reference_rf_trap_100_steps(deckdir):
    from oracle import RefHarness
    d = decks.deck("c2", deckdir + "_live", n_particles=10, collisions=False, geometry="RF_22PT", x_sampl=101, z_sampl=101)
    aos = disk_particles(np.random.default_rng(4), 5000, 1e-2, 1e-2, 5e-3, 1500.0)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref, _sim(d["config"], d["species_conf"]) as sim:
        h = ref.species_index("H_NEG")
        assert np.abs(sim.get_field("uRF") - ref.get_field("uRF")).max() <= 1e-8
        sim.set_field("u", ref.get_field("u"))
        sim.set_field("uRF", ref.get_field("uRF"))
        ref.set_particles(h, aos)
        sim.set_particles(h, aos)
        ref.advance_init()
        sim.advance_init()
        ref.advance(100)
        sim.advance(100)
        a, b = sim.get_particles(h), ref.get_particles(h)
        both = (a[:, 7] > 0) & (b[:, 7] > 0)
        assert np.sum((a[:, 7] > 0) != (b[:, 7] > 0)) <= 2
        assert relerr(a[both][:, [0, 2, 3, 4, 5]], b[both][:, [0, 2, 3, 4, 5]]) <= 1e-10


# ------------------------------------------------------------------------------ next rows (§8f)
def test_u_smooth_matches_oracle(orc, deckdir):
    d = decks.deck("c4", deckdir + "_sm", n_particles=4000, geometry="TUBE", probe_radius=1.4e-3, u_smooth=1,
                   x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        u = np.random.default_rng(2).normal(size=(g.M, g.N))
        sim.set_field("u", u)
        sim.u_smooth()
        assert np.abs(sim.get_field("u") - orc.u_smooth(g, u)).max() <= 1e-15
        sim.set_field("u", u)
        sim.u_smooth(symmetry=True, radius=1.0e-3)
        assert np.abs(sim.get_field("u") - orc.u_smooth(g, u, symmetry=True, radius=1.0e-3)).max() <= 1e-15


def test_energy_histogram_matches_reference_histogram_rule(deckdir):
    # Histogram::add (histogram.cpp:21-33): strict bounds, bin = (int)((f-min)*n/(max-min)), totals include outliers
    d = decks.deck("c1", deckdir, n_particles=100)
    with _sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        rng = np.random.default_rng(4)
        n = 50000
        aos = np.zeros((n, 7))
        aos[:, 0] = aos[:, 2] = 1e-2
        aos[:, 3:6] = rng.normal(size=(n, 3)) * 1.2e6
        sim.set_particles(e, aos)
        hist, st = sim.energy_hist(e, nbins=200, emax=30.0)
        f = (aos[:, 3] ** 2 + aos[:, 5] ** 2 + aos[:, 4] ** 2) * 9.11e-31 * 0.5 / 1.602189e-19
        inside = (f < 30.0) & (f > 0.0)
        want = np.bincount((f[inside] * 200 / 30.0).astype(int), minlength=200)
        assert np.array_equal(hist, want)
        assert st["n_tot"] == n and st["n_in"] == inside.sum()
        assert abs(st["sum_tot"] - f.sum()) <= 1e-9 * f.sum()


def test_device_loaders_fill_the_domain_statistically(deckdir):
    d = decks.deck("c4", deckdir, n_particles=200000, x_sampl=65, z_sampl=65, r_max=6.4e-3, z_max=6.4e-3)
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        sim.run_initscript(d["initscript"])
        for name, T, m in (("ARGON_POS", 300.0, 6.68173e-26), ("ELECTRON", 23200.0, 9.11e-31)):
            p = sim.get_particles(sim.species_index(name))
            assert p.shape[0] == 100000 and p[:, 7].all()
            assert abs(p[:, 0].mean() / 6.4e-3 - 0.5) < 0.01 and abs(p[:, 2].mean() / 6.4e-3 - 0.5) < 0.01
            vth = np.sqrt(1.380662e-23 * T / m)              # each component: rnor * v_max / sqrt(2)
            assert abs(p[:, 3:6].std() / vth - 1.0) < 0.01


# --------------------------------------------------------------------- magnetic field from file (f-row 3)
GB = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_v2_btable.npz"))


def test_magnetic_field_table_cylindrical_vs_golden_reference(orc, deckdir):
    """magnetic_field_const = 0: the table travels through config.load_magnetic_field -> mag2d_set_magnetic_field,
    Fields::B and the cylindrical Boris mover (init + 20 steps) against the reference's recorded outputs"""
    from common import write_btable
    bfile = write_btable(os.path.join(deckdir, "btable_gpu.txt"), 25, 31, 1.2e-2, 7.5e-2)
    d = decks.deck("c3", deckdir + "_bt", n_particles=10, collisions=False, x_sampl=41, z_sampl=61, magnetic_field_const=0,
                   magnetic_field_file=bfile, selfconsistent=0, geometry="PENNING_SIMPLE")
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        assert sim.btable["r_sampl"] == 25 and sim.btable["z_sampl"] == 31
        B = sim.field_B(GB["bt_x"], GB["bt_z"])
        assert np.abs(B - GB["bt_B"]).max() <= 1e-15 * np.abs(GB["bt_B"]).max() * 4
        e = sim.species_index("ELECTRON")
        sim.set_field("u", GB["bt_u"])
        sim.set_field("uRF", GB["bt_uRF"])
        for _ in range(3):
            sim.species_advance(e)          # the golden run started at niter = 3
        sim.set_particles(e, GB["bt_in"])
        sim.species_advance_init(e)
        out = sim.get_particles(e)
        assert relerr(out[:, 3:6], GB["bt_init"][:, 3:6]) <= 1e-12
        sim.species_advance(e)
        for _ in range(19):
            sim.species_advance(e)
        out = sim.get_particles(e)
        ref = GB["bt_out"]
        both = out[:, 7] > 0
        assert both.sum() >= 0.9 * len(ref)
        assert relerr(out[both][:, [0, 2, 3, 4, 5]], ref[both][:, [0, 2, 3, 4, 5]]) <= 1e-10


@pytest.mark.parametrize("coord", ["CARTESIAN", "CYLINDRICAL"])
def test_magnetic_field_table_vs_oracle_with_deposit_and_sort(orc, deckdir, coord):
    """the table mode of the fused kernel in the self-consistent configuration (deposit + fused sort variants):
    trajectories against the oracle, charge grid bit-exact for the final particle set"""
    from common import write_btable
    r_max, z_max = (1.2e-2, 7.5e-2) if coord == "CYLINDRICAL" else (6.4e-3, 6.4e-3)
    bfile = write_btable(os.path.join(deckdir, "btable_gpu_%s.txt" % coord), 33, 21, r_max, z_max, descending=True)
    if coord == "CYLINDRICAL":
        d = decks.deck("c3", deckdir + "_bts", n_particles=10, collisions=False, x_sampl=41, z_sampl=61, magnetic_field_const=0,
                       magnetic_field_file=bfile, macroparticle_factor=1e-3)     # light macro-particles: negligible space charge
        name = "ELECTRON"
    else:
        d = decks.deck("c4", deckdir + "_bts", n_particles=10, collisions=False, x_sampl=65, z_sampl=65, r_max=r_max, z_max=z_max,
                       magnetic_field_const=0, magnetic_field_file=bfile, density_total=1e6)
        name = "ELECTRON"
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        bt = orc.load_magnetic_field(bfile)
        m, names = model_from(orc, d["species_conf"])
        e = names.index(name)
        rng = np.random.default_rng(12)
        aos = disk_particles(rng, 6000, 0.5 * r_max, 0.5 * z_max, 0.3 * r_max, 4e5)
        sim.set_particles(e, aos)
        P = Particles.from_aos7(aos)
        u = np.zeros((sim.M, sim.N))
        sim.set_field("u", u)          # field-free: the same array drives both sides
        for step in range(9):
            if step == 8:
                sim.rho_reset(e)       # the grid then holds the deposit of the last step alone
            sim.species_advance(e)
            orc.advance_boris(g, u, u, m, e, P, niter=step, rng=None, btable=bt)
            orc.advance_boundary(g, sim.mask, m.get(e, "charge"), P)
        out = sim.get_particles(e)
        both = (out[:, 7] > 0) & (P.alive > 0)
        assert np.sum((out[:, 7] > 0) != (P.alive > 0)) <= 2 and both.sum() > 3000
        assert relerr(out[both][:, [0, 2, 3, 4, 5]], P.aos7()[both][:, [0, 2, 3, 4, 5]]) <= 1e-11
        fixed, _ = orc.deposit_fixed(g, out[:, 0].copy(), out[:, 2].copy(), out[:, 7].astype(np.uint8))
        assert np.array_equal(sim.rho_fixed(e), fixed)
        # the full step (solve + push with the cell sort fused in: the SORTING instantiations of the table mode)
        n0 = int((out[:, 7] > 0).sum())
        sim.set_sort_interval(2)
        sim.advance_init()
        sim.advance(6)
        out2 = sim.get_particles(e)
        live = out2[:, 7] > 0
        assert 0.9 * n0 <= live.sum() <= n0 and np.isfinite(out2[live][:, [0, 2, 3, 4, 5]]).all()
        fixed2, _ = orc.deposit_fixed(g, out2[:, 0].copy(), out2[:, 2].copy(), out2[:, 7].astype(np.uint8))
        assert np.array_equal(sim.rho_fixed(e), fixed2)


def test_magnetic_field_table_must_cover_the_box(deckdir):
    """the reference throws "Field2D::interpolate() outside of range" when a particle leaves the table; the ABI refuses
    such a table up front, with the same text"""
    from common import write_btable
    from mag2d_b200.api import Mag2dError
    bfile = write_btable(os.path.join(deckdir, "btable_small.txt"), 9, 9, 0.6e-2, 7.5e-2)
    d = decks.deck("c3", deckdir + "_btx", n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=bfile)
    with pytest.raises(Mag2dError, match="outside of range"):
        _sim(d["config"], d["species_conf"], presolve=False)
    d = decks.deck("c3", deckdir + "_bty", n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0,
                   magnetic_field_file=os.path.join(deckdir, "no_such_file.txt"))
    with pytest.raises(RuntimeError, match="failed opening file"):
        _sim(d["config"], d["species_conf"], presolve=False)


def test_magnetic_field_table_can_be_dropped_again(orc, deckdir):
    """mag2d_set_magnetic_field(NULL): the context forgets the table; with magnetic_field_const = 0 the step then refuses
    to run instead of silently using a constant"""
    from common import write_btable
    from mag2d_b200.api import Mag2dError
    bfile = write_btable(os.path.join(deckdir, "btable_drop.txt"), 9, 9, 1.2e-2, 7.5e-2)
    d = decks.deck("c3", deckdir + "_btdrop", n_particles=10, collisions=False, x_sampl=41, z_sampl=61, magnetic_field_const=0,
                   magnetic_field_file=bfile, selfconsistent=0)
    with _sim(d["config"], d["species_conf"], presolve=False) as sim:
        e = sim.species_index("ELECTRON")
        sim.set_particles(e, disk_particles(np.random.default_rng(2), 100, 0.6e-2, 3.7e-2, 2e-3, 4e5))
        sim.species_advance(e)
        B = sim.field_B(np.array([0.3e-2]), np.array([3.75e-2]))
        assert B[1, 0] == pytest.approx(0.03 * (1 - 0.5 * 400.0 * 0.3e-2 ** 2), rel=1e-3) and B[2, 0] == 0.0
        sim.set_magnetic_field(None)
        with pytest.raises(Mag2dError, match="no table"):
            sim.species_advance(e)
