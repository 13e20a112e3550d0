"""GPU parity of the 3-D path (coord = CARTESIAN3D) against oracle/mag3d_oracle.c, which is
itself pinned to the reference's Field3D / Geometry / Solver (tests/test_oracle3d_vs_reference.py).

Tolerances: trajectories (collisions off) 1e-12 relative after 1 step, 1e-10 after 50; deposit bit-exact against the
fixed-point restatement; Poisson solve 1e-8 against the direct solve of the reference's matrix."""
import numpy as np
import pytest

from mag2d_b200 import decks
from oracle import Oracle3, Orc3Grid

pytestmark = pytest.mark.gpu
QE, ME = 1.602189e-19, 9.11e-31


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def grid3(p):
    return Orc3Grid.make((int(p["x_sampl"]), int(p["y_sampl"]), int(p["z_sampl"])), p["idx"], p["idy"], p["idz"], p["x_max"],
                         p["y_max"], p["z_max"], int(p["boundary"]), p["macroparticle_factor"])


def box_particles(rng, n, g, vth, margin=0.0):
    a = np.zeros((n, 7))
    a[:, 0] = rng.uniform(margin, g.x_max - margin, n)
    a[:, 1] = rng.uniform(margin, g.y_max - margin, n)
    a[:, 2] = rng.uniform(margin, g.z_max - margin, n)
    a[:, 3:6] = rng.normal(size=(n, 3)) * vth
    return a


def soa_of(aos):
    return {k: np.ascontiguousarray(aos[:, c]) for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz"))}


def small_deck(deckdir, **kw):
    args = dict(n_particles=1000, collisions=False, x_sampl=17, y_sampl=15, z_sampl=13)
    args.update(kw)
    return decks.deck("c5", deckdir, **args)


def test_geometry_and_vacuum_solve(deckdir):
    orc = Oracle3()
    d = small_deck(deckdir)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        assert sim.param["dy"] == pytest.approx(1e-4, rel=1e-12)
        mask, volt = orc.geometry(g)
        assert np.array_equal(sim.mask.astype(np.int8), mask) and np.array_equal(sim.voltage, volt)
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, np.zeros(g.shape)))
        u = sim.get_field("u")
        assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max(), sim.solve_info
        assert u[8, 7, 6] == pytest.approx(1.0, abs=1e-12)
        assert sim.solve_info["u"]["resid"] < 1e-12


def test_solve_with_charge_and_gather(deckdir):
    orc = Oracle3()
    d = small_deck(deckdir, x_sampl=20, y_sampl=17, z_sampl=23)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        mask, volt = orc.geometry(g)
        rng = np.random.default_rng(7)
        e = sim.species_index("ELECTRON")
        aos = box_particles(rng, 20000, g, 1e5, margin=2e-4)
        sim.set_particles(e, aos)
        sim.species_accumulate(e)
        fixed, bad = orc.deposit_fixed(g, aos[:, 0], aos[:, 1], aos[:, 2])
        assert bad == 0 and np.array_equal(sim.rho_fixed(e), fixed)
        rho = sim.get_field("rho")
        assert np.abs(rho - fixed * 2.0 ** -32 * -QE).max() <= 1e-12 * np.abs(rho).max()
        info = sim.solve()
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        u = sim.get_field("u")
        assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max(), info
        n = 3000
        x = rng.uniform(0, g.x_max, n)
        y = rng.uniform(0, g.y_max, n)
        z = rng.uniform(0, g.z_max, n)
        got = sim.field_E3(x, y, z)
        want = -orc.grad(g, u, x, y, z)
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()


@pytest.mark.parametrize("shape", [(9, 8, 10), (8, 9, 9), (5, 70, 9), (5, 9, 69)])
def test_folded_sine_transforms_every_parity_and_padding(deckdir, shape):
    """the solver folds every grid line into its symmetric and antisymmetric halves (poisson3d.cu): even and odd interior
    lengths along y and z (a middle element that is its own mirror image), half lengths below and above one 32-wide tile"""
    orc = Oracle3()
    d = small_deck(deckdir + "_fold%d_%d_%d" % shape, x_sampl=shape[0], y_sampl=shape[1], z_sampl=shape[2])
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        mask, volt = orc.geometry(g)
        rng = np.random.default_rng(11)
        e = sim.species_index("ELECTRON")
        aos = box_particles(rng, 5000, g, 1e5)
        sim.set_particles(e, aos)
        sim.species_accumulate(e)
        rho = sim.get_field("rho")
        info = sim.solve()
        u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, rho))
        u = sim.get_field("u")
        assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max(), info
        assert info["resid"] < 1e-12


def test_dense_deposit_uses_the_warp_merge_and_stays_bit_exact(deckdir):
    """>= 32 particles per cell switches the REDUX merge on (push3d.cu: warp_deposit3); sorted and unsorted stores,
    one and many particles per warp call must all give the oracle's integer grid"""
    orc = Oracle3()
    d = small_deck(deckdir, x_sampl=7, y_sampl=6, z_sampl=8)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        rng = np.random.default_rng(5)
        e = sim.species_index("ELECTRON")
        n = 30001                                   # ragged: not a multiple of the 512-slot CTA tile
        aos = box_particles(rng, n, g, 1e5)
        assert n / ((g.imax - 1) * (g.jmax - 1) * (g.kmax - 1)) > 32
        sim.set_particles(e, aos)
        for sort_first in (False, True):
            if sort_first:
                sim.sort(e)
            sim.rho_reset()
            sim.species_accumulate(e)
            fixed, bad = orc.deposit_fixed(g, aos[:, 0], aos[:, 1], aos[:, 2])
            assert bad == 0 and np.array_equal(sim.rho_fixed(e), fixed)
        assert fixed.sum() == pytest.approx(n * 2.0 ** 32, abs=8 * n)


@pytest.mark.parametrize("B", [(0.0, 0.0, 0.0), (0.01, -0.02, 0.03)])
def test_trajectories_vs_oracle(deckdir, B):
    orc = Oracle3()
    d = small_deck(deckdir, selfconsistent=0, Br=B[0], Bt=B[1], Bz=B[2])
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        mask, _ = orc.geometry(g)
        # a stronger field than the 1 V point electrode gives: scale the vacuum solution
        u = sim.get_field("u") * 50.0
        sim.set_field("u", u)
        rng = np.random.default_rng(9)
        e = sim.species_index("ELECTRON")
        aos = box_particles(rng, 5000, g, 3e5)
        sim.set_particles(e, aos)
        soa = soa_of(aos)
        alive = np.ones(len(aos), dtype=np.uint8)
        dt = sim.species[e]["dt"]
        for steps, tol in ((1, 1e-12), (49, 1e-10)):
            for _ in range(steps):
                sim.species_advance(e)
                orc.advance(g, u, mask, -QE, ME, dt, B, soa, alive)
            out = sim.get_particles(e)
            assert np.array_equal(out[:, 7] > 0, alive > 0)
            live = alive > 0
            for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz")):
                scale = np.abs(soa[k][live]).max()
                assert np.abs(out[live, c] - soa[k][live]).max() <= tol * scale, (steps, k)
        assert 0 < (alive == 0).sum() < len(alive)


def test_selfconsistent_loop_deposit_bit_exact(deckdir):
    orc = Oracle3()
    d = small_deck(deckdir, x_sampl=14, y_sampl=12, z_sampl=13, macroparticle_factor=2e6)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        mask, volt = orc.geometry(g)
        rng = np.random.default_rng(11)
        e = sim.species_index("ELECTRON")
        aos = box_particles(rng, 8000, g, 2e5)
        sim.set_particles(e, aos)
        soa = soa_of(aos)
        alive = np.ones(len(aos), dtype=np.uint8)
        dt = sim.species[e]["dt"]
        sim.advance_init()
        fixed, _ = orc.deposit_fixed(g, soa["x"], soa["y"], soa["z"])
        for step in range(4):
            sim.advance(1)
            # Pic::advance order: solve with the charge of the previous step, then push + deposit
            u_ref = orc.solve_direct(g, mask, orc.rhs(g, mask, volt, fixed * 2.0 ** -32 * -QE))
            fixed = np.zeros(g.shape, dtype=np.int64)
            orc.advance(g, u_ref, mask, -QE, ME, dt, (0, 0, 0), soa, alive, None, fixed)
            assert np.abs(sim.get_field("u") - u_ref).max() <= 1e-9 * np.abs(u_ref).max()
            out = sim.get_particles(e)
            assert np.array_equal(out[:, 7] > 0, alive > 0)
            live = alive > 0
            for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz")):
                assert np.abs(out[live, c] - soa[k][live]).max() <= 1e-9 * np.abs(soa[k][live]).max(), (step, k)
            # the deposit is an integer sum: bit-exact against the restatement fed the GPU's own positions
            gfix, _ = orc.deposit_fixed(g, out[live, 0], out[live, 1], out[live, 2])
            assert np.array_equal(sim.rho_fixed(e), gfix)


@pytest.mark.parametrize("layout,interval", [("slots", 1), ("slots", 3), ("bricks", 4)])
def test_fused_cell_sort_3d_keeps_the_particle_set_and_the_charge(deckdir, layout, interval):
    """the 3-D step reorders the store — slot order with the fused COUNT / PERMUTE cell sort, or binned by brick —; with
    collisions off the multiset of particles, the integer charge grid and the potential must equal those of a run that
    never sorts (same operations on every particle: bit for bit)"""
    d = small_deck(deckdir, x_sampl=12, y_sampl=11, z_sampl=13, macroparticle_factor=2e6)
    rng = np.random.default_rng(41)
    results = []
    aos = None
    for k in (0, interval):
        with _sim(d["config"], d["species_conf"]) as sim:
            g = grid3(sim.param)
            e = sim.species_index("ELECTRON")
            if aos is None:
                aos = box_particles(rng, 9001, g, 4e5)
            sim.set_particles(e, aos)
            sim.set_sort_interval(k)
            sim.set_store_layout(layout)
            sim.advance_init()
            sim.advance(7)
            p = sim.get_particles(e)
            live = p[p[:, 7] > 0][:, :6]
            results.append((live[np.lexsort(live.T[::-1])], sim.rho_fixed(e), sim.get_field("u"), p[:, 7]))
    a, b = results
    assert 0 < len(a[0]) < 9001 and len(a[0]) == len(b[0])
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2])
    n_alive = int(b[3].sum())
    if layout == "slots":
        assert (b[3][:n_alive] == 0).sum() < 0.6 * (9001 - n_alive)      # dead slots were compacted away


@pytest.mark.parametrize("B,boundary", [((0.0, 0.0, 0.0), "FREE"), ((0.3, -0.2, 0.5), "FREE"), ((0.0, 0.0, 0.0), "PERIODIC")])
def test_brick_layout_equals_the_plain_step_bit_for_bit(deckdir, B, boundary):
    """push3d_brick.cu against the slot-order kernel on a grid whose last bricks are partial (22 x 18 x 17 cells), with fast
    particles (a third of a cell per step: every step a fifth of them changes brick, many leave the box or wrap around),
    a magnetic field and a potential with structure: the same particles, charge grids and potentials, bit for bit"""
    d = small_deck(deckdir, x_sampl=23, y_sampl=19, z_sampl=18, macroparticle_factor=5e6, boundary=boundary,
                   Br=B[0], Bt=B[1], Bz=B[2])
    rng = np.random.default_rng(4242)
    results = []
    aos = None
    for layout in ("plain", "bricks"):
        with _sim(d["config"], d["species_conf"]) as sim:
            g = grid3(sim.param)
            e = sim.species_index("ELECTRON")
            if aos is None:
                aos = box_particles(rng, 60013, g, 2.5e6)
                aos[:7, 0] = g.x_max              # exactly on the far faces
                aos[7:14, 1] = g.y_max
                aos[14:21, 2] = g.z_max
            sim.set_particles(e, aos)
            sim.set_sort_interval(0 if layout == "plain" else 4)
            sim.set_store_layout("bricks")
            sim.advance_init()
            snaps = []
            for _ in range(3):
                sim.advance(4)
                p = sim.get_particles(e)
                live = p[p[:, 7] > 0][:, :6]
                snaps.append((live[np.lexsort(live.T[::-1])], sim.rho_fixed(e), sim.get_field("u")))
            results.append(snaps)
            if layout == "bricks":
                st = sim.store_stats(e)
                assert st["rebinnings"] >= 1 and st["list_overflow"] == 0
    for a, b in zip(*results):
        assert len(a[0]) == len(b[0]) and len(a[0]) > 1000
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1], b[1])
        assert np.array_equal(a[2], b[2])
    if boundary == "FREE":
        assert len(results[0][-1][0]) < 60013
    else:
        assert len(results[0][-1][0]) == 60013


def test_brick_bins_overflow_into_guests_and_rebin(deckdir):
    """all particles start in one corner and stream into empty bricks whose bins only have the minimal slack: arrivals
    that find a bin full are placed in the next bin with room (guests: global-memory path), the store is re-binned with
    more slack, and nothing is lost or duplicated on the way"""
    d = small_deck(deckdir, x_sampl=21, y_sampl=21, z_sampl=21, macroparticle_factor=5e6)
    rng = np.random.default_rng(77)
    results = []
    aos = None
    for layout in ("plain", "bricks"):
        with _sim(d["config"], d["species_conf"]) as sim:
            g = grid3(sim.param)
            e = sim.species_index("ELECTRON")
            if aos is None:
                n = 40000
                aos = np.zeros((n, 7))
                aos[:, 0:3] = rng.uniform(0.02, 0.2, (n, 3)) * np.array([g.x_max, g.y_max, g.z_max])
                aos[:, 3:6] = np.abs(rng.normal(size=(n, 3))) * 1.5e6 + 5e5       # all heading into the box
            sim.set_particles(e, aos)
            sim.set_sort_interval(0 if layout == "plain" else 4)
            sim.advance_init()
            overflowed = 0
            for _ in range(5):
                sim.advance(5)
                sim.sync()                # lets the step adopt the overflow counters of the previous ones
                if layout == "bricks":
                    overflowed += sim.store_stats(e)["full_bins"]
            p = sim.get_particles(e)
            live = p[p[:, 7] > 0][:, :6]
            results.append((live[np.lexsort(live.T[::-1])], sim.rho_fixed(e), sim.count(e)[0]))
            if layout == "bricks":
                st = sim.store_stats(e)
                assert overflowed > 0, st                 # arrivals did find full bins (and were placed elsewhere) ...
                assert st["rebinnings"] >= 2, st          # ... which forced at least one re-binning after the first one
    a, b = results
    assert a[2] == b[2] and a[2] > 1000
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1], b[1])


def test_brick_layout_collision_rate(deckdir):
    """collisions on the binned store: the Bernoulli test of the null-collision method fires with probability
    1 - exp(-dt / lifetime) per particle-step, the collision pass changes velocities only (count conserved inside the box)"""
    d = small_deck(deckdir, x_sampl=21, y_sampl=21, z_sampl=21, collisions=True, boundary="PERIODIC")
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        e = sim.species_index("ELECTRON")
        n, steps = 200000, 20
        sim.set_particles(e, box_particles(np.random.default_rng(5), n, g, 6e5))
        sim.set_collision_counting(True)
        sim.set_sort_interval(-1)
        sim.advance_init()
        e0 = (sim.get_particles(e)[:, 3:6] ** 2).sum()
        sim.advance(steps)
        assert sim.count(e)[0] == n
        c = sim.collision_counts(e)
        events = c.sum()
        prob = sim.species_get(e, "prob")
        expect = n * steps * prob
        assert expect > 1000 and abs(events - expect) <= 5 * np.sqrt(expect), (events, expect)
        assert (sim.get_particles(e)[:, 3:6] ** 2).sum() != e0
        assert sim.store_stats(e)["rebinnings"] >= 1


def test_sort_and_loader_3d(deckdir):
    d = small_deck(deckdir, x_sampl=17, y_sampl=17, z_sampl=17, collisions=True)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid3(sim.param)
        e = sim.species_index("ELECTRON")
        sim.generate(e, "everywhere", 50000)
        p = sim.get_particles(e)
        assert p.shape[0] == 50000 and (p[:, 7] > 0).all()
        for c, hi in ((0, g.x_max), (1, g.y_max), (2, g.z_max)):
            assert 0 <= p[:, c].min() and p[:, c].max() <= hi
            assert abs(p[:, c].mean() / hi - 0.5) < 0.01
        before = p[np.lexsort(p[:, :6].T[::-1])]
        sim.sort(e)
        q = sim.get_particles(e)
        assert np.array_equal(q[np.lexsort(q[:, :6].T[::-1])][:, :6], before[:, :6])
        key = (np.floor(q[:, 0] * g.idx).astype(np.int64) * (g.jmax - 1) + np.floor(q[:, 1] * g.idy).astype(np.int64)) * (g.kmax - 1) \
            + np.floor(q[:, 2] * g.idz).astype(np.int64)
        assert (np.diff(key) >= 0).all()
        # a few self-consistent steps with collisions on: particles are conserved or absorbed, never created
        sim.set_sort_interval(2)
        sim.advance_init()
        sim.advance(5)
        alive, slots = sim.count(e)
        assert 0 < alive <= 50000
        assert np.isfinite(sim.get_field("u")).all()


def test_streamed_step_3d_equals_the_resident_step(deckdir):
    """mag2d_step_streamed3: host-resident six-array store pushed through the staging ring chunk by chunk; with
    collisions off it leaves exactly the particles, the charge grid and the potential of mag2d_step"""
    from mag2d_b200.api import Sim
    d = decks.deck("c5", deckdir + "_str3", n_particles=10, collisions=False, x_sampl=17, y_sampl=15, z_sampl=13)
    rng = np.random.default_rng(41)
    n = 9000
    with Sim(d["config"], d["species_conf"]) as ref, Sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        p = sim.param
        a = np.zeros((n, 7))
        a[:, 0] = rng.uniform(1e-7, p["x_max"] - 1e-7, n)
        a[:, 1] = rng.uniform(1e-7, p["y_max"] - 1e-7, n)
        a[:, 2] = rng.uniform(1e-7, p["z_max"] - 1e-7, n)
        a[:, 3:6] = rng.normal(size=(n, 3)) * 8e5
        ref.set_particles(e, a)
        sim.set_particles(e, a)
        ref.set_sort_interval(0)
        ref.advance_init()
        sim.advance_init()
        got = sim.get_particles(e)
        host = [np.ascontiguousarray(got[:, c]) for c in range(6)]
        host[0][got[:, 7] == 0] = np.nan
        sim._chk(sim.L.mag2d_particles_clear(sim.h, e))
        for _ in range(4):
            ref.advance(1)
            sim.step_streamed([e], [n], [[c.ctypes.data for c in host]], chunk_slots=2048)     # 5 chunks: the ring wraps
        want = ref.get_particles(e)
        alive = want[:, 7] > 0
        assert np.array_equal(~np.isnan(host[0]), alive) and 0 < alive.sum() < n
        for c in range(6):
            assert np.array_equal(host[c][alive], want[alive, c])
        assert np.array_equal(sim.rho_fixed(e), ref.rho_fixed(e))
        assert np.array_equal(sim.get_field("u"), ref.get_field("u"))
