"""Electrode geometries: mask (FIXED=0, FIXED_RF=1, FREE=2, BOUNDARY=3) and Dirichlet voltages.

Python host mirror of the reference's ``t_grid`` builders (src/fields.cpp:374-868).  Runs once per
run on the host; the arrays go to the device through ``mag2d_set_grid``.
"""
import math

import numpy as np

from .config import GEOMETRY

FIXED, FIXED_RF, FREE, BOUNDARY = 0, 1, 2, 3


def _circle(mask, volt, dx, dz, rc, zc, radius, v, kind):
    M, N = mask.shape
    r = np.arange(M)[:, None] * dx - rc
    z = np.arange(N)[None, :] * dz - zc
    sel = r * r + z * z <= radius * radius
    mask[sel] = kind
    volt[sel] = v


def _square(mask, volt, dx, dz, rmin, rmax, zmin, zmax, v, kind=FIXED):
    M, N = mask.shape
    r = np.arange(M)[:, None] * dx
    z = np.arange(N)[None, :] * dz
    sel = (r > rmin) & (r < rmax) & (z > zmin) & (z < zmax)
    mask[sel] = kind
    volt[sel] = v


def _mark_boundary(mask):
    """nodes next to a FIXED node become BOUNDARY (only FIXED neighbours count, e.g. fields.cpp:553-564);
    the reference sweeps in place, so a node is tested against the *updated* mask — BOUNDARY != FIXED,
    hence the order of the sweep does not matter."""
    M, N = mask.shape
    f = mask == FIXED
    nb = np.zeros_like(f)
    nb[1:, :] |= f[:-1, :]
    nb[:-1, :] |= f[1:, :]
    nb[:, 1:] |= f[:, :-1]
    nb[:, :-1] |= f[:, 1:]
    inner = np.zeros_like(f)
    inner[2:M - 2, 2:N - 2] = True
    mask[inner & nb & ~f] = BOUNDARY


def _multipole(mask, volt, dx, dz, npoles, r_ring, r_rod):
    for i in range(npoles):
        x = 1e-2 + math.sin(2 * math.pi * (i + 1.0 / 32) / npoles) * r_ring
        y = 1e-2 + math.cos(2 * math.pi * (i + 1.0 / 32) / npoles) * r_ring
        _circle(mask, volt, dx, dz, x, y, r_rod, -1 if i % 2 == 0 else 1, FIXED_RF)


def build_geometry3d(p, electrode=True):
    """Geometry::Geometry of the reference's 3-D code (src/fields3d.cpp:13-37): zero-Dirichlet box frame — the test
    `k == z_sampl` of :28 never fires, so the k = z_sampl-1 face stays FREE — plus the one-node Quadrupole electrode
    (id -1, 1 V; fields3d.hpp:26-36) at the centre node.  -> (mask uint8 [M,K,N] with the electrode id stored as 255,
    voltage float64 [M,K,N]); any mask value other than FREE is a Dirichlet node."""
    M, K, N = int(p["x_sampl"]), int(p["y_sampl"]), int(p["z_sampl"])
    mask = np.full((M, K, N), FREE, dtype=np.uint8)
    mask[0, :, :] = mask[M - 1, :, :] = FIXED
    mask[:, 0, :] = mask[:, K - 1, :] = FIXED
    mask[:, :, 0] = FIXED
    volt = np.zeros((M, K, N))
    if electrode:
        mask[M // 2, K // 2, N // 2] = 255
        volt[M // 2, K // 2, N // 2] = 1.0
    return mask, volt


def build_geometry(p):
    """p: Param dict from config.read_config -> (mask uint8 [M,N], voltage float64 [M,N])"""
    if int(p["coord"]) == 2:
        return build_geometry3d(p)
    M, N = int(p["x_sampl"]), int(p["z_sampl"])
    dx, dz = p["dx"], p["dz"]
    geo = int(p["geometry"])
    name = {v: k for k, v in GEOMETRY.items()}[geo]
    mask = np.full((M, N), FREE, dtype=np.uint8)
    volt = np.zeros((M, N))
    axis_open = name in ("MAC", "PENNING", "PENNING_SIMPLE")
    ramp = name in ("EMPTY", "PROBE", "RF_8PT", "RF_HAITRAP")
    edge = np.zeros((M, N), dtype=bool)
    edge[M - 1, :] = True
    edge[:, 0] = True
    edge[:, N - 1] = True
    if not axis_open:
        edge[0, :] = True
    mask[edge] = FIXED
    if ramp:
        j = np.arange(N)[None, :] * np.ones((M, 1))
        volt[edge] = (-p["extern_field"] * dz * (j - N // 2))[edge]
    if name == "PROBE":
        _circle(mask, volt, dx, dz, (M - 1) * dx / 2, (N - 1) * dz / 2, p["probe_radius"], p["u_probe"], FIXED)
        _mark_boundary(mask)
    elif name == "RF_22PT":
        _multipole(mask, volt, dx, dz, 22, 0.75e-2, 0.05e-2)
        _mark_boundary(mask)
    elif name == "RF_8PT":
        _multipole(mask, volt, dx, dz, 8, 0.3e-2 + 0.1e-2, 0.1e-2)
        _mark_boundary(mask)
    elif name == "RF_HAITRAP":
        _multipole(mask, volt, dx, dz, 8, 0.3e-2 + 0.01e-2, 0.01e-2)
        _mark_boundary(mask)
    elif name == "RF_QUAD":
        _circle(mask, volt, dx, dz, 5e-3, 1e-2, 2e-3, 1.0, FIXED_RF)
        _circle(mask, volt, dx, dz, 15e-3, 1e-2, 2e-3, 1.0, FIXED_RF)
        _circle(mask, volt, dx, dz, 1e-2, 5e-3, 2e-3, -1.0, FIXED_RF)
        _circle(mask, volt, dx, dz, 1e-2, 15e-3, 2e-3, -1.0, FIXED_RF)
        _mark_boundary(mask)
    elif name == "TUBE":
        c = p["x_max"] / 2.0
        r = np.arange(M)[:, None] * dx - c
        z = np.arange(N)[None, :] * dz - c
        sel = r * r + z * z >= p["probe_radius"] ** 2
        mask[sel] = FIXED
        volt[sel] = 0.0
        _mark_boundary(mask)
    elif name == "MAC":
        th, ofs = p["u_probe"], 3e-2
        _square(mask, volt, dx, dz, 5e-3, 4.5e-2, 1e-2, 1.5e-2, -.00)
        _square(mask, volt, dx, dz, 5e-3, 7e-3, 2e-2, 8e-2, 0.0)
        _square(mask, volt, dx, dz, 5e-3, 4.5e-2, 8.5e-2, 9e-2, -.00)
        _square(mask, volt, dx, dz, 3e-2, 3.3e-2, 11e-2 + ofs, 14e-2 + ofs, 0.8 * th)
        _square(mask, volt, dx, dz, 4.5e-2, 4.8e-2, 15e-2 + ofs, 25e-2 + ofs, th)
        _square(mask, volt, dx, dz, 3e-2, 3.3e-2, 26e-2 + ofs, 29e-2 + ofs, 1.0 * th)
        _square(mask, volt, dx, dz, 2.5e-2, 2.8e-2, 29e-2 + ofs, 30.5e-2 + ofs, 1.0 * th)
        _square(mask, volt, dx, dz, 15e-3, 4.5e-2, 35e-2, 35.3e-2, .0)
        _square(mask, volt, dx, dz, 0.0, 4.5e-2, 39.5e-2, 40e-2, 3e3)
        _mark_boundary(mask)
    elif name == "PENNING":
        _square(mask, volt, dx, dz, 1.57e-2 / 2, 1.67e-2 / 2, 1e-3, 25e-3, -5)
        _square(mask, volt, dx, dz, 0, 1.46e-2 / 2, 15e-3, 16e-3, 10)
        _square(mask, volt, dx, dz, 0, 7e-3 / 2, 12e-3, 19e-3, 10)
        _square(mask, volt, dx, dz, 0, 1.9e-3, 1e-3, 12e-3, 10)
        _square(mask, volt, dx, dz, 4e-3, 7e-3, 52e-3, 53e-3, -5)
        _square(mask, volt, dx, dz, 2.5e-3, 7e-3, 46e-3, 47e-3, 0)
        _mark_boundary(mask)
    elif name == "PENNING_SIMPLE":
        ri, ro = 1e-2, 1.1e-2
        _square(mask, volt, dx, dz, ri, ro, 0, 1e-2, -0.5)
        _square(mask, volt, dx, dz, ri, ro, 1.1e-2, 2e-2, 0)
        _square(mask, volt, dx, dz, ri, ro, 2.1e-2, 6e-2, -1.0)
        _square(mask, volt, dx, dz, ri, ro, 6.1e-2, 7.5e-2, -10)
        _mark_boundary(mask)
    return mask, volt
