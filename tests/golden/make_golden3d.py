#!/usr/bin/env python
"""Generate tests/golden/reference3d_v1.npz from the compilable part of the reference's 3-D code (Field3D, Geometry,
Solver in oracle/_ref/libmag3d_ref.so, built from /root/reference by `make -C oracle ref`):

    python tests/golden/make_golden3d.py

Inputs and the reference's outputs for deposit, interpolation, gradient, geometry, is_free and the solve, so that
tests/test_oracle_golden.py pins oracle/mag3d_oracle.c where neither /root/reference nor oracle/_ref exists.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import Ref3D  # noqa: E402


def write_config(tmp, nx=9, ny=8, nz=7, x_max=8e-3, y_max=7e-3, z_max=6e-3, mpf=50.0):
    dx, dz, dy = x_max / (nx - 1), z_max / (nz - 1), y_max / (ny - 1)
    V = dx * dy * dz * (nx - 1) * (nz - 1)
    path = os.path.join(tmp, "config3d.txt")
    with open(path, "w") as f:
        f.write("coord = CARTESIAN3D\nboundary = FREE\nmover = ADVANCE_BORIS\ngeometry = EMPTY\nselfconsistent = 1\n"
                "x_sampl = %d\ny_sampl = %d\nz_sampl = %d\nr_max = %.17g\ny_max = %.17g\nz_max = %.17g\n"
                "n_particles_total = 1e6\ndensity_total = %.17g\nmacroparticle_factor = %.17g\nrf = 0\n"
                % (nx, ny, nz, x_max, y_max, z_max, 1e6 / V, mpf))
    return path


def main():
    out = {}
    tmp = tempfile.mkdtemp(prefix="golden3d_")
    with Ref3D(write_config(tmp)) as r:
        out["dims"] = np.array(r.shape)
        out["grid"] = np.array([r.idx, r.idy, r.idz, r.x_max, r.y_max, r.z_max, r.macroparticle_factor])
        mask, volt = r.mask()
        out["mask"], out["voltage"] = mask, volt
        rng = np.random.default_rng(2024)
        n = 2000
        x = rng.uniform(0, r.x_max * (1 - 1e-9), n)
        y = rng.uniform(0, r.y_max * (1 - 1e-9), n)
        z = rng.uniform(0, r.z_max * (1 - 1e-9), n)
        out["pos"] = np.stack([x, y, z], axis=1)
        out["is_free"] = r.is_free(x, y, z)
        assert r.accumulate(-1.6e-19, x, y, z) == 0
        out["rho"] = r.get("rho")
        u = rng.normal(size=r.shape)
        r.set("u", u)
        out["u_random"] = u
        out["interp"] = r.interpolate(x, y, z)
        gx = rng.uniform(0, r.x_max - 0.51 * r.dx, n)
        gy = rng.uniform(0, r.y_max - 0.51 * r.dy, n)
        gz = rng.uniform(0, r.z_max - 0.51 * r.dz, n)
        out["grad_pos"] = np.stack([gx, gy, gz], axis=1)
        out["grad"], _ = r.grad(gx, gy, gz)
        out["u_solved"] = r.solve()          # Solver::solve scales rho in place into the right-hand side
        out["rhs"] = r.get("rho")
    path = os.path.join(HERE, "reference3d_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
