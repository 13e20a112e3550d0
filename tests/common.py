"""Shared helpers for the parity tests: build the oracle's grid/model from the same input decks
that drive the reference harness and the CUDA path."""
import numpy as np

from mag2d_b200 import config as cfg
from oracle import OrcGrid, ref_available
from oracle.pyoracle import Particles  # noqa: F401

needs_ref = ref_available("parity")


def grid_from_param(p):
    return OrcGrid.make(int(p["x_sampl"]), int(p["z_sampl"]), p["x_max"], p["z_max"], coord=int(p["coord"]),
                        boundary=int(p["boundary"]), selfconsistent=int(p["selfconsistent"]), rf=int(p["rf"]),
                        geometry_empty=int(p["geometry"] == 0), extern_field=p["extern_field"],
                        rf_amplitude=p["rf_amplitude"], rf_U0=p["rf_U0"], rf_omega=p["rf_omega"], Br=p["Br"],
                        Bz=p["Bz"], Bt=p["Bt"], dV=p["dV"], macroparticle_factor=p["macroparticle_factor"],
                        field_from_file=int(p["electric_field_from_file"]))


def model_from(orc, species_path):
    sp, it = cfg.read_species(species_path)
    names = [s["name"] for s in sp]
    m = orc.model(len(sp))
    for i, s in enumerate(sp):
        m.set_species(i, s["type"], s["mass"], s["charge"], s["density"], s["temperature"], s["E_max"], s["dt"])
    for q in it:
        m.add_interaction(q["type"], q["DE"], q["rate"], q["cutoff"], names.index(q["primary"]),
                          names.index(q["secondary"]), q["CS_energy"] or None, q["CS_value"] or None)
    m.lifetime_init()
    return m, names


def disk_particles(rng, n, cx, cz, radius, vth, ttd=0.0):
    """n x 7 AoS (x,y,z,vx,vy,vz,time_to_death): uniform disk, Maxwellian-like velocities"""
    aos = np.zeros((n, 7))
    ang = rng.uniform(0, 2 * np.pi, n)
    rad = np.sqrt(rng.uniform(0, 1, n)) * radius
    aos[:, 0] = cx + rad * np.cos(ang)
    aos[:, 2] = cz + rad * np.sin(ang)
    aos[:, 3:6] = rng.normal(size=(n, 3)) * vth
    aos[:, 6] = ttd
    return aos


def write_btable(path, nr, nz, r_max, z_max, descending=False, header=True):
    """a magnetic field table in the four-column format Fields::load_magnetic_field reads (fields.cpp:882-896):
    r z Br Bz per row, r slowest; a mirror-like field Bz = B0 (1 + a z'^2 - a r^2 / 2), Br = -a B0 r z'"""
    r = np.linspace(0.0, r_max, nr)
    z = np.linspace(0.0, z_max, nz)
    if descending:
        r, z = r[::-1], z[::-1]
    B0, a = 0.03, 400.0
    with open(path, "w") as f:
        if header:
            f.write("# r z Br Bz\n")
        for ri in r:
            for zj in z:
                zp = zj - 0.5 * z_max
                f.write("%.17g %.17g %.17g %.17g\n" % (ri, zj, -a * B0 * ri * zp, B0 * (1 + a * zp * zp - 0.5 * a * ri * ri)))
    return path
