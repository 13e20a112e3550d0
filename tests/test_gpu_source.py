"""GPU parity tests of the particle source (use_source = 1): Species<CARTESIAN>::source5_refresh
and ::source (reference src/particles.cpp:1053-1080, 1158-1226) through the C ABI against the CPU oracle, which is pinned
bit for bit against the compiled reference in tests/test_oracle_vs_reference.py::test_particle_source_bit_exact.

Deterministic legs (factor = 1: rand() % 1 == 0, collisions off) are compared particle by particle (1e-12); the
RNG-driven legs (lateral shifts, reservoir generation, collisions) statistically, because the device draws from Philox.
"""
import numpy as np
import pytest

from common import Particles, disk_particles, grid_from_param, model_from
from mag2d_b200 import decks

pytestmark = pytest.mark.gpu
L = 6.4e-3


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def _deck(deckdir, tag, factor, collisions, **kw):
    return decks.deck("c4", deckdir + tag, n_particles=10, collisions=collisions, x_sampl=33, z_sampl=33, r_max=L, z_max=L,
                      use_source=1, src_fact=factor, n_particles_total=40, density_total=1e13, **kw)


@pytest.mark.parametrize("Bz", [0.0, 0.02])
def test_source_deterministic_leg_vs_oracle(orc, deckdir, Bz):
    d = _deck(deckdir, "_src1", 1, False, Bz=Bz, extern_field=200.0)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        src = orc.source_refresh(g, m, e, 1, sim.param["V"], orc.rng(21))
        assert src.n == 4000
        main = disk_particles(np.random.default_rng(3), 50, 0.5 * L, 0.5 * L, 0.2 * L, 4e5)
        sim.set_particles(e, main)
        sim.set_source_particles(e, 1, src.aos7())
        got = sim.get_source_particles(e)
        assert np.array_equal(got[:, [0, 2, 3, 4, 5, 6]], src.aos7()[:, [0, 2, 3, 4, 5, 6]])
        dst = Particles(50 + 6000)
        dst.alive[:] = 0
        n_dst, total_orc, total_gpu = 50, 0, 0
        sim.rho_reset(e)
        for step in range(30):
            total_gpu += sim.species_source(e)
            inj, n_dst = orc.source(g, m, e, 1, src, dst, n_dst, rng=None)
            total_orc += inj
        assert total_gpu == total_orc > 100
        # reservoir: same particles, slot by slot
        res = sim.get_source_particles(e)
        ref = src.aos7()
        for col in (0, 2, 3, 4, 5):
            assert np.abs(res[:, col] - ref[:, col]).max() <= 1e-12 * np.abs(ref[:, col]).max()
        # injected copies: the same set (slot order differs: the device hands out tail slots atomically)
        out = sim.get_particles(e)
        assert out.shape[0] == 50 + total_gpu and (out[:, 7] > 0).all()
        assert np.array_equal(out[:50, :6], main[:, :6])
        a = out[50:][:, [0, 2, 3, 4, 5]]
        b = dst.aos7()[50:n_dst][:, [0, 2, 3, 4, 5]]
        # by vx (continuous, collision-free), then x: a corner crossing yields two copies with the same velocity
        a, b = a[np.lexsort((a[:, 0], a[:, 2]))], b[np.lexsort((b[:, 0], b[:, 2]))]
        assert (np.abs(a - b).max(axis=0) <= 1e-12 * np.abs(b).max(axis=0)).all()
        assert (a[:, 0] > 0).all() and (a[:, 0] < L).all() and (a[:, 1] > 0).all() and (a[:, 1] < L).all()
        # their charge: bit-exact fixed-point deposit of exactly the injected copies
        inj_only = out[50:]
        fixed, bad = orc.deposit_fixed(g, inj_only[:, 0].copy(), inj_only[:, 2].copy(), np.ones(len(inj_only), dtype=np.uint8))
        assert bad == 0 and np.array_equal(sim.rho_fixed(e), fixed)


def test_source_random_legs_statistics(orc, deckdir):
    """factor = 4, collisions on: device-generated reservoir (Philox) and lateral shifts against the oracle's
    (SHR3 + libc rand) in distribution"""
    factor = 4
    d = _deck(deckdir, "_src4", factor, True, extern_field=0.0)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        sim.set_particles(e, disk_particles(np.random.default_rng(3), 50, 0.5 * L, 0.5 * L, 0.2 * L, 4e5))
        sim.param["V"] *= 40.0                      # a bigger reservoir for the statistics: n = 1e15 * 1.6e-10 / 4
        sim.source_refresh(e, factor)
        res = sim.get_source_particles(e)
        n = res.shape[0]
        assert n == int(1e15 * sim.param["V"] / factor) == 40000
        w = L / factor
        assert res[:, 0].min() >= 0 and res[:, 0].max() <= w and res[:, 2].min() >= 0 and res[:, 2].max() <= w
        for col in (0, 2):                          # uniform: mean w/2 +- 4 sigma, sigma = w / sqrt(12 n)
            assert abs(res[:, col].mean() - 0.5 * w) <= 4 * w / np.sqrt(12 * n)
        vmax = sim.species_get(e, "v_max")
        for col in (3, 4, 5):                       # Maxwellian components: variance v_max^2 / 2
            v = res[:, col]
            assert abs(v.mean()) <= 4 * vmax / np.sqrt(2 * n)
            assert abs(v.var() / (0.5 * vmax ** 2) - 1) <= 4 * np.sqrt(2.0 / n)
        lifetime = sim.species_get(e, "lifetime")
        assert abs(res[:, 6].mean() / lifetime - 1) <= 4 / np.sqrt(n)
        # influx over 40 steps against the oracle with its own reservoir of the same size
        src = orc.source_refresh(g, m, e, factor, sim.param["V"], orc.rng(5))
        assert src.n == n
        r = orc.rng(6)
        dst = Particles(50 + 400000)
        dst.alive[:] = 0
        n_dst, n_orc, n_gpu = 50, 0, 0
        for step in range(40):
            n_gpu += sim.species_source(e)
            inj, n_dst = orc.source(g, m, e, factor, src, dst, n_dst, rng=r, libc_seed=3 if step == 0 else None)
            n_orc += inj
        assert n_gpu > 5000 and abs(n_gpu - n_orc) <= 5 * np.sqrt(n_gpu + n_orc)
        out = sim.get_particles(e)[50:]
        assert out.shape[0] == n_gpu
        # copies enter through the four edges within one step's flight of the wall; the lateral coordinate covers the whole
        # edge: its `factor` reservoir-width bins are equally likely (chi-square, 3 dof: P(> 21) ~ 1e-4)
        reach = 6 * vmax * 1e-11
        near_x = (out[:, 0] < reach) | (out[:, 0] > L - reach)
        near_z = (out[:, 2] < reach) | (out[:, 2] > L - reach)
        assert (near_x | near_z).all()
        lat = np.concatenate([out[near_x & ~near_z][:, 2], out[near_z & ~near_x][:, 0]])
        counts = np.bincount(np.minimum((lat / w).astype(int), factor - 1), minlength=factor)
        chi2 = ((counts - counts.mean()) ** 2 / counts.mean()).sum()
        assert chi2 < 21.0, counts
        # energy of the influx: copies are a flux-weighted sample (mean kinetic energy 2 kT), the oracle's agrees
        ek_gpu = (out[:, 3:6] ** 2).sum(axis=1).mean()
        oo = dst.aos7()[50:n_dst]
        ek_orc = (oo[:, 3:6] ** 2).sum(axis=1).mean()
        assert abs(ek_gpu / ek_orc - 1) <= 0.05


def test_source_inside_the_step_keeps_the_charge_grid_exact(orc, deckdir):
    """mag2d_step with use_source: push + boundary + deposit, then source() per species; the fixed-point grid equals the
    deposit of everything alive at the end of the step, and the stand-alone sort trims the slot range"""
    d = _deck(deckdir, "_srcstep", 4, True, extern_field=0.0, selfconsistent=1)
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        rng = np.random.default_rng(8)
        idx = [sim.species_index(nm) for nm in ("ARGON_POS", "ELECTRON")]
        for i, vth in zip(idx, (300.0, 4e5)):
            sim.set_particles(i, disk_particles(rng, 3000, 0.5 * L, 0.5 * L, 0.45 * L, vth))
        sim.param["V"] *= 10.0
        sim.set_sort_interval(3)
        sim.advance_init()
        sim.source_refresh()
        n_res = [sim.get_source_particles(i).shape[0] for i in idx]
        assert n_res == [10000, 10000]
        sim.advance(12)
        for i in idx:
            o = sim.get_particles(i)
            live = o[:, 7] > 0
            assert live.sum() > 1000
            fixed, bad = orc.deposit_fixed(g, o[:, 0].copy(), o[:, 2].copy(), o[:, 7].astype(np.uint8))
            assert bad == 0 and np.array_equal(sim.rho_fixed(i), fixed)
        # electrons stream in (and out through the FREE walls); the slot range follows the live count after the sort
        e = idx[1]
        n_live, n_slots = sim.count(e)
        assert n_slots <= 1.5 * n_live + 4096
        assert sim.get_source_particles(e).shape[0] == 10000


def test_source_is_cartesian_only(deckdir):
    from mag2d_b200.api import Mag2dError
    d = decks.deck("c3", deckdir + "_srccyl", n_particles=10, x_sampl=41, z_sampl=61, use_source=1)
    with _sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        sim.set_particles(e, disk_particles(np.random.default_rng(1), 10, 5e-3, 3e-2, 1e-3, 4e5))
        with pytest.raises(Mag2dError, match="CARTESIAN"):
            sim.source_refresh(e, 4)


def test_source_edge_cases(orc, deckdir):
    """empty store -> no reservoir (particles.cpp:1063); empty reservoir -> source() is a no-op; download of nothing"""
    d = _deck(deckdir, "_srcedge", 4, False, extern_field=0.0)
    with _sim(d["config"], d["species_conf"]) as sim:
        e = sim.species_index("ELECTRON")
        sim.source_refresh(e, 4)                       # the species has no particles yet
        assert sim.get_source_particles(e).shape[0] == 0
        assert sim.species_source(e) == 0
        sim.set_particles(e, disk_particles(np.random.default_rng(3), 10, 0.5 * L, 0.5 * L, 0.2 * L, 4e5))
        sim.set_source_particles(e, 4, np.zeros((0, 7)))
        assert sim.species_source(e) == 0 and sim.count(e)[0] == 10
        # a reservoir particle that crosses a corner of its box in one step enters twice (x block, then z block)
        w = L / 4
        v = 0.5 * w / 1e-11                            # half a reservoir width per step, towards the corner
        sim.set_source_particles(e, 4, np.array([[w * 0.9, 0, w * 0.9, v, 0.0, v, 0.0]]))
        n = sim.species_source(e)
        out = sim.get_particles(e)[10:]
        assert n == out.shape[0] and n <= 2            # copies whose lateral shift leaves the box are dropped like in the reference
        res = sim.get_source_particles(e)
        assert 0 <= res[0, 0] <= w and 0 <= res[0, 2] <= w
