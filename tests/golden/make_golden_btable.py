#!/usr/bin/env python
"""Generate tests/golden/reference_v2_btable.npz from the UNMODIFIED reference (oracle/_ref): the magnetic field
table path (magnetic_field_const = 0): Fields::load_magnetic_field, Fields::B, and the cylindrical Boris mover
(init + 20 steps) with the per-particle table look-up.  Run in the build container:

    python tests/golden/make_golden_btable.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import disk_particles, write_btable  # noqa: E402
from mag2d_b200 import decks  # noqa: E402
from oracle import RefHarness  # noqa: E402

R_MAX, Z_MAX = 1.2e-2, 7.5e-2


def main():
    out = {}
    tmp = tempfile.mkdtemp(prefix="golden_bt_")
    bfile = write_btable(os.path.join(tmp, "btable.txt"), 25, 31, R_MAX, Z_MAX)
    d = decks.deck("c3", tmp, n_particles=10, x_sampl=41, z_sampl=61, magnetic_field_const=0, magnetic_field_file=bfile,
                   selfconsistent=0, geometry="PENNING_SIMPLE")
    with RefHarness(d["config"], d["species_conf"], seed=5) as ref:
        info, br, bz = ref.btable()
        out["bt_info"] = np.array([info[k] for k in ("jmax", "lmax", "dx", "dy", "xmin", "ymin")], dtype=np.float64)
        out["bt_Br"], out["bt_Bz"] = br, bz
        rng = np.random.default_rng(4)
        x, z = rng.uniform(0, R_MAX * 0.999, 512), rng.uniform(0, Z_MAX * 0.999, 512)
        out["bt_x"], out["bt_z"], out["bt_B"] = x, z, ref.field_B(x, z)
        out["bt_u"], out["bt_uRF"] = ref.get_field("u"), ref.get_field("uRF")
        e = ref.species_index("ELECTRON")
        aos = disk_particles(rng, 800, 0.5 * R_MAX, 0.5 * Z_MAX, 0.2 * R_MAX, 4e5)
        out["bt_in"] = aos
        ref.set_particles(e, aos)
        ref.species_set(e, "niter", 3)
        ref.advance_position(e, init=True)
        out["bt_init"] = ref.get_particles(e)
        ref.species_set(e, "lifetime", np.inf)
        for step in range(20):
            ref.species_set(e, "niter", 3 + step)
            ref.advance_position(e)
        out["bt_out"] = ref.get_particles(e)
    np.savez_compressed(os.path.join(HERE, "reference_v2_btable.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_v2_btable.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
