#!/usr/bin/env python
"""Generate tests/golden/reference_v1.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`).  Run in the build container:

    python tests/golden/make_golden.py

The fixture holds inputs and the reference's outputs for every leg of the hot path so that
tests/test_oracle_golden.py can pin the CPU restatement (and, through it, the CUDA path) on machines
where neither /root/reference nor oracle/_ref exists.  Reference build: -O2 -ffp-contract=off.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import disk_particles  # noqa: E402
from mag2d_b200 import decks  # noqa: E402
from oracle import REF_DIR, RefHarness  # noqa: E402


def main():
    out = {}
    tmp = tempfile.mkdtemp(prefix="golden_")
    # --- RNG streams (t_random seeded like test_MCC.cpp:51)
    d = decks.deck("c1", tmp, n_particles=200)
    with RefHarness(d["config"], d["species_conf"], seed=1234) as ref:
        for what in ("iuni", "uni", "rnor", "rexp", "radius"):
            out["rng_" + what] = ref.rng_draw(what, 4096)
        out["rng_rot"] = ref.rng_rot(2.5, 64)
        v = np.random.default_rng(0).normal(size=(64, 3))
        out["rng_deflect_in"] = v
        out["rng_deflect_out"] = ref.rng_deflect(0.3, v)
        # lifetime / rates / sigma_v known answers
        e, he = ref.species_index("ELECTRON"), ref.species_index("HELIUM")
        out["c1_lifetime"] = np.array([ref.species(e)["lifetime"]])
        out["c1_rates"] = ref.rates(e)
        vs = np.geomspace(1e3, 2e7, 64)
        out["c1_sigma_v_v"] = vs
        out["c1_sigma_v"] = np.array([[ref.sigma_v(e, he, k, x) for x in vs] for k in range(3)])
        # scatter under seed 77
        ref.rng_seed(77)
        v = np.random.default_rng(1).normal(size=(512, 3)) * 1.5e6
        out["c1_scatter_in"] = v
        out["c1_scatter_out"] = ref.scatter(e, v)
        # multicoll mover + periodic boundary, seed 99, 2 steps
        ref.run_initscript(d["initscript"])
        parts = ref.get_particles(e)
        aos = parts[parts[:, 7] > 0, :7].copy()
        ref.set_particles(e, aos)
        out["c1_multicoll_in"] = aos
        ref.rng_seed(99)
        for _ in range(2):
            ref.advance_position(e)
            ref.advance_boundary(e)
        out["c1_multicoll_out"] = ref.get_particles(e)
    # --- chi(beta) from the reference's own tests/test_langevin.cpp
    txt = subprocess.check_output([os.path.join(REF_DIR, "test_langevin")]).decode().split("\n")
    rows = [list(map(float, l.split())) for l in txt if l.strip()]
    out["langevin_chi"] = np.array(rows)          # beta, chi, asymptote (6 significant digits)
    # --- C2: RF trap (8-pole geometry as shipped, plus 22-pole), Boris, gather
    for geo in ("RF_8PT", "RF_22PT"):
        d = decks.deck("c2", tmp, n_particles=10, geometry=geo, x_sampl=41, z_sampl=41, Bt=0.01, Bz=0.02, Br=0.005)
        with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
            h = ref.species_index("H_NEG")
            k = "c2_%s_" % geo
            out[k + "mask"] = ref.get_field("mask").astype(np.uint8)
            out[k + "voltage"] = np.where(ref.get_field("mask") < 2, ref.get_field("voltage"), 0.0)
            out[k + "u"] = ref.get_field("u")
            out[k + "uRF"] = ref.get_field("uRF")
            rng = np.random.default_rng(3)
            x = rng.uniform(0, 2e-2, 256)
            z = rng.uniform(0, 2e-2, 256)
            x[:4] = 0
            z[4:8] = 0
            x[8:12] = 2e-2
            z[12:16] = 2e-2
            out[k + "E_xz"] = np.stack([x, z])
            out[k + "E"] = np.stack(ref.field_E(x, z, 3.3e-8))
            out[k + "is_free"] = ref.is_free(x[16:], z[16:])
            aos = disk_particles(np.random.default_rng(3), 256, 1e-2, 1e-2, 2.5e-3, 1500.0)
            out[k + "boris_in"] = aos
            ref.set_particles(h, aos)
            ref.species_set(h, "lifetime", np.inf)
            ref.species_set(h, "niter", 17)
            ref.advance_position(h, init=True)
            out[k + "boris_init"] = ref.get_particles(h)
            for step in range(100):
                ref.species_set(h, "niter", 17 + step)
                ref.advance_position(h)
                if step == 0:
                    out[k + "boris_1"] = ref.get_particles(h)
            ref.advance_boundary(h)
            out[k + "boris_100"] = ref.get_particles(h)
    # --- C4: self-consistent two-species loop with MCC, seed 21
    d = decks.deck("c4", tmp, n_particles=1000, x_sampl=33, z_sampl=33, r_max=3.2e-3, z_max=3.2e-3)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        ii, ie = ref.species_index("ARGON_POS"), ref.species_index("ELECTRON")
        rng = np.random.default_rng(7)
        ai = disk_particles(rng, 500, 1.6e-3, 1.6e-3, 1.4e-3, 300.0)
        ae = disk_particles(rng, 500, 1.7e-3, 1.6e-3, 1.4e-3, 6e5)
        out["c4_in_i"], out["c4_in_e"] = ai, ae
        ref.set_particles(ii, ai)
        ref.set_particles(ie, ae)
        ref.advance_init()
        out["c4_rho0"] = ref.get_field("rho")
        out["c4_u0"] = ref.get_field("u")
        out["c4_mask"] = ref.get_field("mask").astype(np.uint8)
        ref.rng_seed(21)
        ref.advance(5)
        out["c4_out_i"], out["c4_out_e"] = ref.get_particles(ii), ref.get_particles(ie)
        out["c4_rho5"] = ref.get_field("rho")
        out["c4_u5"] = ref.get_field("u")
        out["c4_lifetimes"] = np.array([ref.species(ii)["lifetime"], ref.species(ie)["lifetime"]])
    # --- C3: cylindrical self-consistent, Bz = 0.03 T
    d = decks.deck("c3", tmp, n_particles=500, x_sampl=41, z_sampl=51)
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        ie = ref.species_index("ELECTRON")
        rng = np.random.default_rng(11)
        n = 500
        aos = np.zeros((n, 7))
        aos[:, 0] = np.sqrt(rng.uniform(0, 1, n)) * 4e-3
        aos[:, 2] = 3.75e-2 + 2e-2 * (rng.uniform(0, 1, n) - 0.5)
        aos[:, 3:6] = rng.normal(size=(n, 3)) * 4e5
        aos[:3, 0] = 0.0
        aos[:3, 3] = 0.0
        aos[:3, 4] = 0.0
        out["c3_in"] = aos
        ref.set_particles(ie, aos)
        ref.advance_init()
        out["c3_u0"] = ref.get_field("u")
        out["c3_init"] = ref.get_particles(ie)
        ref.advance(5)
        out["c3_out"] = ref.get_particles(ie)
        out["c3_rho5"] = ref.get_field("rho")
        out["c3_u5"] = ref.get_field("u")
    path = os.path.join(HERE, "reference_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
