"""Edge cases of the CUDA path through the C ABI: empty and ragged particle stores, everything removed, positions
exactly on the domain boundary, tiny stores through the fused sort and the streamed step, and the error paths.
Checked against the CPU oracle like the other parity tests."""
import ctypes as C

import numpy as np
import pytest

from common import Particles, grid_from_param, model_from
from mag2d_b200 import decks

pytestmark = pytest.mark.gpu


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def c4(deckdir, **kw):
    args = dict(n_particles=1000, collisions=False, x_sampl=17, z_sampl=21, r_max=1.6e-3, z_max=2.0e-3)
    args.update(kw)
    return decks.deck("c4", deckdir, **args)


def uniform(rng, n, p, vth):
    a = np.zeros((n, 7))
    a[:, 0] = rng.uniform(0, p["x_max"], n)
    a[:, 2] = rng.uniform(0, p["z_max"], n)
    a[:, 3:6] = rng.normal(size=(n, 3)) * vth
    return a


def test_empty_species_step_and_solve(orc, deckdir):
    d = c4(deckdir)
    with _sim(d["config"], d["species_conf"]) as sim:
        sim.set_sort_interval(1)
        sim.advance_init()
        sim.advance(3)
        for name in ("ARGON_POS", "ELECTRON"):
            i = sim.species_index(name)
            assert sim.count(i) == (0, 0)
            assert sim.get_particles(i).shape[0] == 0
            assert not sim.rho_fixed(i).any()
            sim.sort(i)
        assert not sim.get_field("u").any()          # grounded box, no charge
        hist, stats = sim.energy_hist(sim.species_index("ELECTRON"))
        assert hist.sum() == 0


@pytest.mark.parametrize("n", [1, 2, 255, 257, 1023, 1025])
def test_ragged_store_sizes_match_the_oracle(orc, deckdir, n):
    """stores that end inside a warp tile, inside a thread's pair, or one slot past a CTA"""
    d = c4(deckdir)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        rng = np.random.default_rng(n)
        aos = uniform(rng, n, sim.param, 3e5)
        sim.set_particles(e, aos)
        sim.set_sort_interval(1)
        sim.advance_init()
        P = Particles.from_aos7(aos)
        u = sim.get_field("u")
        orc.advance_boris_init(g, u, u, m, e, P, niter=0)
        for step in range(3):
            sim.advance(1)
            u = sim.get_field("u")
            orc.advance_boris(g, u, u, m, e, P, niter=step, rng=None)
            orc.advance_boundary(g, sim.mask, m.get(e, "charge"), P)
        out = sim.get_particles(e)
        live = out[:, 7] > 0
        ref = P.aos7()[P.alive > 0]
        assert live.sum() == (P.alive > 0).sum()
        got = out[live]
        # the fused sort permutes the slots: compare as sets ordered by position
        o1, o2 = np.lexsort(got[:, [2, 0]].T), np.lexsort(ref[:, [2, 0]].T)
        scale = np.abs(ref[:, [0, 2, 3, 5]]).max(axis=0) if len(ref) else 1.0
        assert (np.abs(got[o1][:, [0, 2, 3, 5]] - ref[o2][:, [0, 2, 3, 5]]) <= 1e-11 * scale).all()
        fixed, _ = orc.deposit_fixed(g, got[:, 0].copy(), got[:, 2].copy(), np.ones(len(got), dtype=np.uint8))
        assert np.array_equal(sim.rho_fixed(e), fixed)


def test_everything_removed_and_boundary_positions(orc, deckdir):
    d = c4(deckdir)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        e = sim.species_index("ELECTRON")
        p = sim.param
        # exactly on the four edges and corners: inside the domain (<=), deposited into the clamped last cell
        xs = np.array([0.0, p["x_max"], 0.0, p["x_max"], 0.5 * p["x_max"], p["x_max"]])
        zs = np.array([0.0, 0.0, p["z_max"], p["z_max"], p["z_max"], 0.5 * p["z_max"]])
        aos = np.zeros((6, 7))
        aos[:, 0], aos[:, 2] = xs, zs
        sim.set_particles(e, aos)
        sim.species_accumulate(e)
        fixed, bad = orc.deposit_fixed(g, xs, zs, np.ones(6, dtype=np.uint8))
        assert bad == 0 and np.array_equal(sim.rho_fixed(e), fixed)
        assert fixed.sum() == pytest.approx(6 * 2.0 ** 32, abs=48)
        # now everybody leaves: 1e9 m/s outwards
        aos[:, 3] = 1e9
        sim._chk(sim.L.mag2d_particles_clear(sim.h, e))
        sim.set_particles(e, aos)
        sim.set_sort_interval(1)
        sim.advance(2)
        assert sim.count(e)[0] == 0
        assert not sim.rho_fixed(e).any()
        sim.advance(2)                               # permuting an all-dead store
        assert sim.count(e)[0] == 0
        # the live count of a permute travels to the host asynchronously; once it has landed the slot range shrinks
        sim.advance(2)
        assert sim.count(e) == (0, 0)


def test_streamed_step_ragged_and_empty(orc, deckdir):
    d = c4(deckdir)
    rng = np.random.default_rng(3)
    with _sim(d["config"], d["species_conf"]) as ref, _sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        i = sim.species_index("ARGON_POS")
        aos = uniform(rng, 1500, sim.param, 3e5)
        ref.set_particles(e, aos)
        ref.set_sort_interval(0)
        cols = [np.ascontiguousarray(aos[:, c]) for c in (0, 2, 3, 4, 5)]
        empty = [np.zeros(0) for _ in range(5)]
        for _ in range(3):
            ref.advance(1)
            # chunk of 1024 slots: one full chunk + a ragged one; the ion species is empty
            sim.step_streamed([i, e], [0, 1500], [[c.ctypes.data for c in empty], [c.ctypes.data for c in cols]], chunk_slots=1000)
        want = ref.get_particles(e)
        alive = want[:, 7] > 0
        assert np.array_equal(~np.isnan(cols[0]), alive)
        for c, col in zip((0, 2, 3, 4, 5), cols):
            assert np.array_equal(col[alive], want[alive, c])
        assert np.array_equal(sim.rho_fixed(e), ref.rho_fixed(e))


def test_error_paths_return_messages_not_crashes(deckdir):
    from mag2d_b200.api import Mag2dError
    d = c4(deckdir)
    with _sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        with pytest.raises(Mag2dError):
            sim.species_advance(99)
        sim.set_particles(e, uniform(np.random.default_rng(0), 100, sim.param, 1e5))
        small = np.zeros(10, dtype=np.float64)
        ns = C.c_int64()
        dp = C.POINTER(C.c_double)
        ptr = small.ctypes.data_as(dp)
        rc = sim.L.mag2d_particles_download_soa(sim.h, e, 10, ptr, None, ptr, ptr, ptr, ptr, None, None, C.byref(ns))
        assert rc != 0 and b"too small" in sim.L.mag2d_last_error()
        with pytest.raises(Mag2dError, match="unknown solver kind"):
            sim._chk(sim.L.mag2d_set_solver_kind(sim.h, 7))
        with pytest.raises(Mag2dError):
            sim.generate(e, "cylinder", 10, 1.0, 1e-3, 1e-3, 1e-3)     # cylindrical loader on a Cartesian grid
        # the context is still usable afterwards
        sim.advance_init()
        sim.advance(1)
        assert sim.count(e)[0] > 0
