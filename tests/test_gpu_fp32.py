"""fp32 storage mode of the particle arrays (BASELINE.json north_star: "fp64, or a stated tolerance in fp32 mode"; BASELINE.md §3:
40 bytes per 2D3V particle-step).  The arrays in HBM hold floats, the arithmetic is fp64 on the widened values.

Stated tolerances (against the fp64 oracle started from the same float-rounded state):
  * one step:   positions within 1.2e-7 of the box size (half an ulp of a float of that magnitude is 6e-8), velocities 1.2e-7 of max |v|
  * 100 steps:  positions 1e-5 of the box size (rounding errors of the stored state accumulate like a random walk, ~sqrt(100) x 6e-8,
                and grow in the field); velocities 2e-4 of max |v| (a position error of 1e-7 box sizes next to the rods of the RF trap, where
                the field changes over a cell, is a field error of that order applied every step; measured 8.7e-5)
  * deposited grid: bit-exact against the fixed-point deposit of the STORED (float) positions — they are rounded before the boundary
    test and the deposit
"""
import numpy as np
import pytest

from common import Particles, disk_particles, grid_from_param, model_from
from mag2d_b200 import decks

pytestmark = pytest.mark.gpu


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


@pytest.mark.parametrize("deck,kw,name,vth", [
    ("c2", dict(geometry="RF_8PT", x_sampl=41, z_sampl=41, Bz=0.02, Bt=0.01), "H_NEG", 2e3),
    ("c3", dict(geometry="EMPTY", x_sampl=61, z_sampl=81, selfconsistent=0), "ELECTRON", 4e5),
])
def test_fp32_storage_trajectories_within_the_stated_tolerance(orc, deckdir, deck, kw, name, vth):
    d = decks.deck(deck, deckdir + "_f32", n_particles=10, collisions=False, **kw)
    with _sim(d["config"], d["species_conf"]) as sim:
        sim.set_storage("f32")
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        i, io = sim.species_index(name), names.index(name)
        u, urf = sim.get_field("u"), sim.get_field("uRF")
        L = min(g.x_max, g.z_max)
        aos = f32(disk_particles(np.random.default_rng(31), 4000, 0.5 * g.x_max if deck == "c2" else 0.3 * g.x_max, 0.5 * g.z_max, 0.25 * L, vth))
        if deck == "c3":
            aos[:, 0] = np.abs(aos[:, 0])
        sim.set_particles(i, aos)
        start = sim.get_particles(i)
        assert np.array_equal(start[:, :7], aos)                      # float-representable values survive the round trip
        P = Particles.from_aos7(aos)
        cols = [0, 2, 3, 4, 5]
        step = 0
        for steps, tol, vtol in ((1, 1.2e-7, 1.2e-7), (99, 1e-5, 2e-4)):
            for _ in range(steps):
                sim.species_advance(i)
                orc.advance_boris(g, u, urf, m, io, P, niter=step, rng=None)
                orc.advance_boundary(g, sim.mask, m.get(io, "charge"), P)
                step += 1
            out, ref = sim.get_particles(i), P.aos7()
            both = (out[:, 7] > 0) & (P.alive > 0)
            assert both.sum() > 2000 and np.sum((out[:, 7] > 0) != (P.alive > 0)) <= 4
            scale = np.array([g.x_max, g.z_max] + [np.abs(ref[both][:, 3:6]).max()] * 3)
            err = np.abs(out[both][:, cols] - ref[both][:, cols]).max(axis=0) / scale
            assert err[:2].max() <= tol and err[2:].max() <= vtol, (steps, err)
            assert np.array_equal(out[both][:, cols], f32(out[both][:, cols]))       # what comes back is what a float holds


def test_fp32_storage_deposit_is_bit_exact_for_the_stored_positions(orc, deckdir):
    d = decks.deck("c4", deckdir + "_f32d", n_particles=40000, collisions=True, x_sampl=65, z_sampl=49, r_max=6.4e-3, z_max=4.8e-3)
    rng = np.random.default_rng(17)
    results = {}
    for interval in (0, 3):
        with _sim(d["config"], d["species_conf"]) as sim:
            sim.set_storage("f32")
            g = grid_from_param(sim.param)
            for name, vth in (("ARGON_POS", 3e2), ("ELECTRON", 7e5)):
                if name not in results:
                    results[name] = f32(disk_particles(rng, 20000, 3.2e-3, 2.4e-3, 2.2e-3, vth))
                sim.set_particles(sim.species_index(name), results[name])
            sim.set_sort_interval(interval)
            sim.advance_init()
            sim.advance(7)                                   # self-consistent steps with collisions and (interval 3) the fused cell sort
            for name in ("ARGON_POS", "ELECTRON"):
                i = sim.species_index(name)
                p = sim.get_particles_soa(i, ("x", "z"))
                assert np.array_equal(p["x"][p["alive"] > 0], f32(p["x"][p["alive"] > 0]))
                fixed, bad = orc.deposit_fixed(g, p["x"], p["z"], p["alive"])
                assert bad == 0 and np.array_equal(sim.rho_fixed(i), fixed), (interval, name)
                live, slots = sim.count(i)
                assert 0 < live <= 20000
            # the explicit sort compacts without changing a bit of the stored state
            i = sim.species_index("ELECTRON")
            before = sim.get_particles(i)
            sim.sort(i)
            after = sim.get_particles(i)
            a, b = before[before[:, 7] > 0][:, :6], after[after[:, 7] > 0][:, :6]
            assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
            hist, st = sim.energy_hist(i)
            assert st["n_tot"] == len(b)


def test_fp32_storage_is_refused_where_it_is_not_implemented(deckdir):
    from mag2d_b200.api import Mag2dError
    d = decks.deck("c4", deckdir + "_f32e", n_particles=100, x_sampl=17, z_sampl=17)
    with _sim(d["config"], d["species_conf"]) as sim:
        sim.set_particles(sim.species_index("ELECTRON"), np.zeros((4, 7)) + 1e-3)
        with pytest.raises(Mag2dError, match="before any particle"):
            sim.set_storage("f32")
    for deck, kw in (("c1", {}), ("c5", dict(x_sampl=9, y_sampl=9, z_sampl=9))):
        d = decks.deck(deck, deckdir + "_f32e", n_particles=100, **kw)
        with _sim(d["config"], d["species_conf"]) as sim:
            with pytest.raises(Mag2dError, match="2-D Boris movers"):
                sim.set_storage("f32")
