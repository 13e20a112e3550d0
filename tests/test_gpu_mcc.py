"""Statistical parity of the Monte-Carlo collisions: the CUDA path draws from counter-based Philox
streams, the reference from one shared SHR3/ziggurat stream, so trajectories differ and agreement is
statistical (BASELINE.json north_star):

  * mean energy within 4 standard errors of the oracle's,
  * per-process collision counts within 4 sigma (Poisson) of the oracle's rates,
  * two-sample Kolmogorov-Smirnov test on the energy distribution, p > 1e-3.

The oracle (oracle/mag2d_oracle.c) is the CPU restatement that tests/test_oracle_vs_reference.py pins
bit-for-bit against the reference under a common seed.
"""
import numpy as np
import pytest
from scipy import stats

from common import Particles, grid_from_param, model_from
from mag2d_b200 import decks

pytestmark = pytest.mark.gpu

QE = 1.602189e-19


def _sim(*a, **k):
    from mag2d_b200.api import Sim
    return Sim(*a, **k)


def energies_eV(p, mass):
    return 0.5 * mass * (p[:, 3] ** 2 + p[:, 4] ** 2 + p[:, 5] ** 2) / QE


def maxwellian(rng, n, temperature, mass):
    vth = np.sqrt(1.380662e-23 * temperature / mass)
    return rng.normal(size=(n, 3)) * vth


def assert_means_agree(a, b, what):
    se = np.sqrt(a.var() / a.size + b.var() / b.size)
    assert abs(a.mean() - b.mean()) <= 4 * se, (what, a.mean(), b.mean(), se)


def assert_counts_agree(c_gpu, n_gpu, c_cpu, n_cpu, what, np_gpu, np_cpu):
    """per-process event counts per particle-step agree within 4 sigma.

    sigma has two parts: the Poisson noise of the counts, and the finite-ensemble noise of the mean
    acceptance probability sigma(E)v/rate_max — the counts of one ensemble are correlated through its
    energy distribution, so two ensembles of np particles differ by about spread/sqrt(np) in relative
    terms even with infinitely many steps (spread <= 0.25 for every process used here)."""
    for k in range(c_gpu.size):
        if c_gpu[k] + c_cpu[k] < 50:
            continue
        ra, rb = c_gpu[k] / n_gpu, c_cpu[k] / n_cpu
        var = c_gpu[k] / n_gpu ** 2 + c_cpu[k] / n_cpu ** 2
        var += (0.25 * max(ra, rb)) ** 2 * (1.0 / np_gpu + 1.0 / np_cpu)
        assert abs(ra - rb) <= 4 * np.sqrt(var), (what, k, ra, rb, np.sqrt(var))


def test_c1_electron_swarm_multicoll(orc, deckdir):
    """config_test_MCC: e- in He at 1 kV/m, ~90 null-collision events per particle-step"""
    d = decks.deck("c1", deckdir, n_particles=1000)
    n_gpu, n_cpu, steps = 200000, 15000, 30
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        e = names.index("ELECTRON")
        mass = m.get(e, "mass")
        assert abs(sim.species_get(e, "lifetime") - m.lifetime(e)) <= 1e-15 * m.lifetime(e)
        assert abs(sim.species_get(e, "lifetime") - 1.11353e-10) < 1e-15       # known answer, SURVEY.md §8c
        rng = np.random.default_rng(12)

        def start(n):
            aos = np.zeros((n, 7))
            aos[:, 0] = rng.uniform(0.5e-2, 1.5e-2, n)
            aos[:, 2] = rng.uniform(0.5e-2, 1.5e-2, n)
            aos[:, 3:6] = maxwellian(rng, n, 1e4, mass)
            return aos                                   # time_to_death = 0, as add_particles_on_disk leaves it
        a_gpu, a_cpu = start(n_gpu), start(n_cpu)
        sim.set_collision_counting(True)
        sim.set_particles(e, a_gpu)
        P = Particles.from_aos7(a_cpu)
        r = orc.rng(4321)
        counts = np.zeros(16 * (len(names) + 1), dtype=np.int64)
        mask, _ = orc.geometry(g, 0)
        for step in range(steps):
            sim.species_advance(e)
            orc.advance_multicoll(0.0, sim.param["extern_field"], m, e, P, r, counts)
            orc.advance_boundary(g, mask, m.get(e, "charge"), P)
            if step + 1 in (10, 20, 30):
                out = sim.get_particles(e)
                assert out[:, 7].all()
                assert_means_agree(energies_eV(out, mass), energies_eV(P.aos7(), mass), "mean energy at step %d" % (step + 1))
        Eg, Ec = energies_eV(sim.get_particles(e), mass), energies_eV(P.aos7(), mass)
        assert 2.0 < Eg.mean() < 7.0          # relaxing towards ~5.9 eV at 1 Td (SURVEY.md §8c swarm curve)
        ks = stats.ks_2samp(Eg[:50000], Ec)
        assert ks.pvalue > 1e-3, ks
        cg = sim.collision_counts(e)
        assert_counts_agree(cg, n_gpu * steps, counts, n_cpu * steps, "c1 process counts", n_gpu, n_cpu)
        he = names.index("HELIUM")
        per_step = (cg[he * 16:he * 16 + 16].sum() + cg[len(names) * 16 + he]) / (n_gpu * steps)
        assert abs(per_step - 89.8043) < 0.2     # dt/lifetime events per particle-step (check_params known answer)
        # the species clock advances twice per step in this mover (particles.cpp:857-858 + particles.hpp:347-348)
        assert sim.species_get(e, "niter") == 2 * steps
        # time_to_death stays in (0, few lifetimes)
        ttd = sim.get_particles(e)[:, 6]
        assert (ttd > 0).all() and abs(ttd.mean() / m.lifetime(e) - 1.0) < 0.02


def test_c2_langevin_buffer_gas_boris(orc, deckdir):
    """H- in the 22-pole trap with He + H2 buffer gas: Langevin (Nanbu-Kitatani) + tabulated elastic"""
    d = decks.deck("c2", deckdir, n_particles=10, x_sampl=101, z_sampl=101)
    n_gpu, n_cpu, steps = 300000, 60000, 100
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        h = names.index("H_NEG")
        mass = m.get(h, "mass")
        assert np.allclose(sim.rates(h), m.rates(h), rtol=1e-14)
        u, urf = sim.get_field("u"), sim.get_field("uRF")
        rng = np.random.default_rng(3)

        def start(n):
            aos = np.zeros((n, 7))
            ang = rng.uniform(0, 2 * np.pi, n)
            rad = np.sqrt(rng.uniform(0, 1, n)) * 1e-3
            aos[:, 0] = 1e-2 + rad * np.cos(ang)
            aos[:, 2] = 1e-2 + rad * np.sin(ang)
            aos[:, 3:6] = maxwellian(rng, n, 3000.0, mass)      # hot ions: the buffer gas cools them
            return aos
        a_gpu, a_cpu = start(n_gpu), start(n_cpu)
        sim.set_collision_counting(True)
        sim.set_particles(h, a_gpu)
        P = Particles.from_aos7(a_cpu)
        r = orc.rng(99)
        counts = np.zeros(16 * (len(names) + 1), dtype=np.int64)
        for step in range(steps):
            sim.species_advance(h)
            orc.advance_boris(g, u, urf, m, h, P, niter=step, rng=r, counts=counts)
            orc.advance_boundary(g, sim.mask, m.get(h, "charge"), P)
        out = sim.get_particles(h)
        assert out[:, 7].all() and P.alive.all()
        Eg, Ec = energies_eV(out, mass), energies_eV(P.aos7(), mass)
        assert_means_agree(Eg, Ec, "H- mean energy")
        assert stats.ks_2samp(Eg[:60000], Ec).pvalue > 1e-3
        cg = sim.collision_counts(h)
        assert_counts_agree(cg, n_gpu * steps, counts, n_cpu * steps, "c2 process counts", n_gpu, n_cpu)
        # the number of collision attempts is Binomial(N*steps, 1-exp(-dt/lifetime))
        prob = sim.species_get(h, "prob")
        attempts = cg.sum()
        assert abs(attempts - n_gpu * steps * prob) <= 4 * np.sqrt(n_gpu * steps * prob)
        assert cg[names.index("HELIUM") * 16] > 1000 and cg[names.index("H2") * 16] > 100   # both Langevin channels fire


def test_c4_charge_exchange_and_elastic_ions(orc, deckdir):
    """Ar+ in Ar: CX hands the ion a thermal neutral velocity, elastic scatters isotropically in the CM frame"""
    d = decks.deck("c4", deckdir + "_cx", n_particles=10, selfconsistent=0, x_sampl=17, z_sampl=17, boundary="PERIODIC")
    # a long ion time step (1e-7 s) so that ~6 % of the ions collide per step
    txt = open(d["species_conf"]).read().replace("DT 1e-11", "DT 1e-07", 1).replace("EMAX 1.0", "EMAX 5.0", 1)
    open(d["species_conf"], "w").write(txt)
    n_gpu, n_cpu, steps = 300000, 100000, 40
    with _sim(d["config"], d["species_conf"]) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, d["species_conf"])
        ii = names.index("ARGON_POS")
        mass = m.get(ii, "mass")
        dt_over_tau = m.get(ii, "dt") / m.lifetime(ii)
        assert 0.01 < dt_over_tau < 0.2
        rng = np.random.default_rng(8)

        def start(n):
            aos = np.zeros((n, 7))
            aos[:, 0] = rng.uniform(1e-2, 4e-2, n)
            aos[:, 2] = rng.uniform(1e-2, 4e-2, n)
            aos[:, 3] = 3000.0        # a 1.9 eV beam along x
            return aos
        a_gpu, a_cpu = start(n_gpu), start(n_cpu)
        sim.set_collision_counting(True)
        sim.set_particles(ii, a_gpu)
        P = Particles.from_aos7(a_cpu)
        r = orc.rng(7)
        counts = np.zeros(16 * (len(names) + 1), dtype=np.int64)
        u = np.zeros((g.M, g.N))
        for step in range(steps):
            sim.species_advance(ii)
            orc.advance_boris(g, u, u, m, ii, P, niter=step, rng=r, counts=counts)
            orc.advance_boundary(g, sim.mask, m.get(ii, "charge"), P)
        out = sim.get_particles(ii)
        Eg, Ec = energies_eV(out, mass), energies_eV(P.aos7(), mass)
        assert_means_agree(Eg, Ec, "Ar+ mean energy")
        assert_means_agree(out[:, 3], P.vx, "Ar+ mean drift")
        cg = sim.collision_counts(ii)
        assert_counts_agree(cg, n_gpu * steps, counts, n_cpu * steps, "c4 ion process counts", n_gpu, n_cpu)
        ar = names.index("ARGON")
        assert cg[ar * 16 + 0] > 100 and cg[ar * 16 + 1] > 100      # elastic and CX both occur


def test_particle_partner_superelastic(orc, deckdir, tmp_path):
    """SUPERELASTIC against a *particle* species (the reference's CRR process, species_conf.txt:127-133):
    partners are drawn from the target's live particles (particles.cpp:230-238)"""
    sp = tmp_path / "crr_species.txt"
    sp.write_text(
        "SPECIES\n NAME H3+\n TYPE ION\n MASS 5.02e-27\n CHARGE 1.602189e-19\n DENSITY 1e17\n TEMPERATURE 77.0\n DT 1e-9\n EMAX 0.5\n\n"
        "SPECIES\n NAME ELECTRON\n TYPE ELECTRON\n MASS 9.11e-31\n CHARGE -1.602189e-19\n DENSITY 1e17\n TEMPERATURE 77.0\n DT 1e-9\n EMAX 10.0\n\n"
        "INTERACTION\n NAME CRR\n TYPE SUPERELASTIC\n PRIMARY ELECTRON\n SECONDARY H3+\n DE 0.13\n RATE 2.0e-10\n CUTOFF 0.0\n\n")
    d = decks.deck("c4", deckdir, n_particles=10, selfconsistent=0, x_sampl=9, z_sampl=9, boundary="PERIODIC")
    n_gpu, n_cpu, steps = 200000, 100000, 20
    with _sim(d["config"], str(sp)) as sim:
        g = grid_from_param(sim.param)
        m, names = model_from(orc, str(sp))
        ie, ii = names.index("ELECTRON"), names.index("H3+")
        me = m.get(ie, "mass")
        rng = np.random.default_rng(21)

        def ions(n):
            aos = np.zeros((n, 7))
            aos[:, 0] = rng.uniform(1e-2, 4e-2, n)
            aos[:, 2] = rng.uniform(1e-2, 4e-2, n)
            aos[:, 3:6] = maxwellian(rng, n, 77.0, m.get(ii, "mass"))
            return aos

        def electrons(n):
            aos = ions(n)
            aos[:, 3:6] = maxwellian(rng, n, 300.0, me)
            return aos
        ai, ae_g, ae_c = ions(5000), electrons(n_gpu), electrons(n_cpu)
        sim.set_particles(ii, ai)
        sim.set_particles(ie, ae_g)
        sim.set_collision_counting(True)
        Pi, Pe = Particles.from_aos7(ai), Particles.from_aos7(ae_c)
        m.set_pool(ii, Pi)
        r = orc.rng(5)
        counts = np.zeros(16 * (len(names) + 1), dtype=np.int64)
        u = np.zeros((g.M, g.N))
        for step in range(steps):
            sim.species_advance(ie)
            orc.advance_boris(g, u, u, m, ie, Pe, niter=step, rng=r, counts=counts)
            orc.advance_boundary(g, sim.mask, m.get(ie, "charge"), Pe)
        out = sim.get_particles(ie)
        Eg, Ec = energies_eV(out, me), energies_eV(Pe.aos7(), me)
        assert Eg.mean() > 1.2 * energies_eV(ae_g, me).mean()       # each CRR event releases 0.13 eV
        assert_means_agree(Eg, Ec, "electron mean energy with CRR heating")
        assert_counts_agree(sim.collision_counts(ie), n_gpu * steps, counts, n_cpu * steps, "CRR counts", n_gpu, n_cpu)
