"""Host-side readers for mag2d's three input formats (Python mirror of the C++ host in csrc/host/).

* ``read_config``     flat GetPot ``key = value  # comment`` files -> ``Param``-like dict
                     (reference src/param.cpp:12-143; keys and defaults in SURVEY.md Appendix B)
* ``read_species``    ``SPECIES`` / ``INTERACTION`` / ``CROSS_SECTION`` grammar
                     (reference src/parser.cpp:3-147); unlike the reference every field is
                     zero-initialised instead of left as garbage (parser.cpp:45 TODO)
* ``read_initscript`` the particle-loader mini language (reference src/pic.cpp:241-328)

These only describe a run; all computation happens behind the C ABI (include/mag2d_b200.h).
"""
import math

import numpy as np

EPS0 = 8.854187817e-12   # reference src/param.cpp:8-10 (old CODATA values kept for parity)
KB = 1.380662e-23
QE = 1.602189e-19

COORD = {"CARTESIAN": 0, "CYLINDRICAL": 1, "CARTESIAN3D": 2}
BOUNDARY = {"FREE": 0, "PERIODIC": 1, "MIRROR": 2}
MOVER = {"ADVANCE_BORIS": 0, "ADVANCE_MULTICOLL": 1}
GEOMETRY = {"EMPTY": 0, "PROBE": 1, "RF_22PT": 2, "RF_8PT": 3, "RF_HAITRAP": 4, "RF_QUAD": 5, "MAC": 6,
            "PENNING": 7, "PENNING_SIMPLE": 8, "TUBE": 9}
SPECIES_TYPE = {"NEUTRAL": 0, "ELECTRON": 1, "ION": 2}
COLL_TYPE = {"ELASTIC": 0, "LANGEVIN": 1, "CX": 2, "COULOMB": 3, "SUPERELASTIC": 4}


class ConfigError(RuntimeError):
    pass


def _getpot(path, overrides=None):
    """flat key=value reader with GetPot semantics: '#' comments, [section] prefixes, first token wins"""
    kv = {}
    section = ""
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            if line.startswith("["):
                sec = line[1:line.find("]")].strip() if "]" in line else line[1:].strip()
                section = sec + "/" if sec else ""
                continue
            if "=" not in line:
                continue
            k, v = line.split("=", 1)
            k = k.strip()
            v = v.strip().split()
            if k:
                kv[section + k] = v[0] if v else ""
    for k, v in (overrides or {}).items():
        kv[k] = str(v)
    return kv


def _num(kv, key, dflt):
    try:
        return float(kv[key])
    except (KeyError, ValueError):
        return float(dflt)


def _int(kv, key, dflt):
    return int(_num(kv, key, dflt))


def read_config(path, overrides=None):
    """-> dict with the public fields of the reference's ``Param`` (src/param.hpp:24-79)"""
    kv = _getpot(path, overrides)
    p = {}
    p["x_max"] = _num(kv, "x_max", _num(kv, "r_max", 1e-2))
    p["y_max"] = _num(kv, "y_max", 1e-2)
    p["z_max"] = _num(kv, "z_max", 1e-2)
    p["x_sampl"] = _int(kv, "x_sampl", _int(kv, "r_sampl", 100))
    p["y_sampl"] = _int(kv, "y_sampl", 2)
    p["z_sampl"] = _int(kv, "z_sampl", 100)
    p["x_min"] = p["y_min"] = p["z_min"] = 0.0
    p["n_particles_total"] = _num(kv, "n_particles_total", 1e5)
    p["density_total"] = _num(kv, "density_total", 1e11)
    p["pressure"] = _num(kv, "pressure", 133.0)
    p["neutral_temperature"] = _num(kv, "neutral_temperature", 300.0)
    p["probe_radius"] = _num(kv, "probe_radius", 1e-4)
    p["probe_length"] = _num(kv, "probe_length", 1e-2)
    p["u_probe"] = _num(kv, "u_probe", -10.0)
    p["extern_field"] = _num(kv, "extern_field", 100.0)
    p["electric_field_static_file"] = kv.get("electric_field_static_file", "")
    p["electric_field_rf_file"] = kv.get("electric_field_rf_file", "")
    p["electric_field_from_file"] = bool(_int(kv, "electric_field_from_file", 0))
    p["has_probe"] = bool(_int(kv, "has_probe", 0))
    p["magnetic_field_file"] = kv.get("magnetic_field_file", "")
    p["magnetic_field_const"] = bool(_int(kv, "magnetic_field_const", 1))
    p["Br"] = _num(kv, "Br", 0.0)
    p["Bz"] = _num(kv, "Bz", 0.0)
    p["Bt"] = _num(kv, "Bt", 0.0)
    try:
        p["niter"] = int(kv.get("niter", "100000"))
    except ValueError:
        raise ConfigError("string2: error converting string %s to T\n" % kv.get("niter"))
    p["dt_elon"] = _num(kv, "dt_elon", 1e-11)
    p["selfconsistent"] = bool(_int(kv, "selfconsistent", 1))
    p["use_source"] = bool(_int(kv, "use_source", 0))
    p["u_smooth"] = bool(_int(kv, "u_smooth", 0))
    p["rf"] = bool(_int(kv, "rf", 1))
    p["rf_amplitude"] = _num(kv, "rf_amplitude", 10.0)
    p["rf_U0"] = _num(kv, "rf_U0", 0.0)
    p["rf_omega"] = _num(kv, "rf_omega", 2 * math.pi * _num(kv, "rf_freq", 20e6))
    if p["selfconsistent"] and p["rf"]:
        raise ConfigError("Param: selfconsistent rf trap not implemented\n")
    if p["selfconsistent"] and p["electric_field_from_file"]:
        raise ConfigError("Param: selfconsistent with electric_field_from_file not implemented")
    p["t_print"] = int(kv.get("t_print", "0"))
    p["t_print_dist"] = int(kv.get("t_print_dist", "0"))
    p["t_dist_sample"] = p["t_print"] // 10 if p["t_print"] > 10 else 1
    te = kv.get("t_equilib", "niter+1")
    p["t_equilib"] = p["niter"] + 1 if te == "niter+1" else int(te)
    p["particle_reload"] = bool(_int(kv, "particle_reload", 0))
    p["particle_reload_dir"] = kv.get("particle_reload_dir", ".")
    p["src_fact"] = _int(kv, "src_fact", 20)
    p["neutral_density"] = p["pressure"] / (KB * p["neutral_temperature"])
    p["do_plot"] = bool(_int(kv, "do_plot", 1))

    def enum(key, dflt, table, what):
        s = kv.get(key, dflt)
        if s not in table:
            raise ConfigError("Param: unrecognized %s value %s\n" % (what, s))
        return table[s]

    p["coord"] = enum("coord", "CYLINDRICAL", COORD, "coord")
    p["boundary"] = enum("boundary", "FREE", BOUNDARY, "boundary")
    if p["coord"] == COORD["CYLINDRICAL"] and p["boundary"] != BOUNDARY["FREE"]:
        raise ConfigError("Param: only FREE boundary condition in cylindrical coords is implemented\n")
    if p["boundary"] == BOUNDARY["MIRROR"]:
        raise ConfigError("Param: MIRROR boundary condition not implemented\n")
    p["mover"] = enum("mover", "ADVANCE_BORIS", MOVER, "mover")
    p["geometry"] = enum("geometry", "EMPTY", GEOMETRY, "geometry")
    p["dx"] = p["x_max"] / (p["x_sampl"] - 1)
    p["dz"] = p["z_max"] / (p["z_sampl"] - 1)
    p["idx"] = 1.0 / p["dx"]
    p["idz"] = 1.0 / p["dz"]
    p["macroparticle_factor"] = _num(kv, "macroparticle_factor", 1e4)
    p["V"] = p["n_particles_total"] / p["density_total"]
    p["dV"] = p["V"] / ((p["x_sampl"] - 1) * (p["z_sampl"] - 1))
    p["dy"] = p["dV"] / (p["dx"] * p["dz"])
    if p["coord"] == COORD["CYLINDRICAL"]:
        p["dy"] = 2 * math.pi / p["macroparticle_factor"]
    p["idy"] = 1.0 / p["dy"]
    return p


def _tokens(path):
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.split()
            if t:
                out.append(t)
    return out


def read_species(path):
    """-> (species list, interaction list) of plain dicts, file order preserved"""
    species, inter = [], []
    state = "INIT"
    for line in _tokens(path):
        head = line[0]
        if head == "DEFAULT":
            state = "DEFAULT"
            continue
        if head == "SPECIES":
            species.append(dict(name="", type=0, charge=0.0, mass=0.0, dt=0.0, density=0.0, temperature=0.0,
                                polarizability=0.0, E_max=0.0))
            state = "SPECIES"
            continue
        if head == "INTERACTION":
            inter.append(dict(name="", type=0, DE=0.0, rate=0.0, cutoff=0.0, primary="", secondary="",
                              CS_energy=[], CS_value=[]))
            state = "INTERACTION"
            continue
        if state == "INIT":
            raise ConfigError("config_parse: unrecognized first config block\n")
        if state == "SPECIES":
            s = species[-1]
            if head == "NAME":
                s["name"] = line[1]
            elif head == "TYPE":
                if line[1] not in SPECIES_TYPE:
                    raise ConfigError('config_parse: unrecognized first species type "%s"' % line[1])
                s["type"] = SPECIES_TYPE[line[1]]
            elif head in ("MASS", "CHARGE", "DENSITY", "DT", "TEMPERATURE", "EMAX"):
                key = {"MASS": "mass", "CHARGE": "charge", "DENSITY": "density", "DT": "dt",
                       "TEMPERATURE": "temperature", "EMAX": "E_max"}[head]
                s[key] = float(line[1])
            else:
                raise ConfigError('config_parse: unrecognized species  parameter "%s"' % head)
        elif state == "INTERACTION":
            it = inter[-1]
            if head == "NAME":
                it["name"] = line[1]
            elif head == "TYPE":
                if line[1] not in COLL_TYPE:
                    raise ConfigError('config_parse: unrecognized interaction type "%s"' % line[1])
                it["type"] = COLL_TYPE[line[1]]
            elif head in ("DE", "RATE", "CUTOFF"):
                it[{"DE": "DE", "RATE": "rate", "CUTOFF": "cutoff"}[head]] = float(line[1])
            elif head == "PRIMARY":
                it["primary"] = line[1]
            elif head == "SECONDARY":
                it["secondary"] = line[1]
            elif head == "CROSS_SECTION":
                state = "CROSS_SECTION"
            else:
                raise ConfigError('config_parse: unrecognized species  parameter"%s"' % head)
        elif state == "CROSS_SECTION":
            if head == "END_CROSS_SECTION":
                state = "INTERACTION"
            else:
                inter[-1]["CS_energy"].append(float(line[0]))
                inter[-1]["CS_value"].append(float(line[1]))
    names = [s["name"] for s in species]
    for it in inter:
        for role in ("primary", "secondary"):
            if it[role] not in names:
                raise ConfigError('Speclist::Speclist: unrecognized primary species "%s" of interaction "%s"\n'
                                  % (it[role], it["name"]))
    return species, inter


def read_initscript(path):
    """-> list of (verb, species name, numeric args) tuples"""
    nargs = {"add_particles_bessel": 6, "add_particles_everywhere": 3, "add_particles_on_disk": 6,
             "add_tracked_particle": 7}
    out = []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                continue
            t = line.split()
            if not t or t[0] not in nargs:
                continue
            if len(t) != nargs[t[0]]:
                raise ConfigError("Pic::run_initscript: wrong number of parameters (%d) to %s\n" % (len(t), t[0]))
            out.append((t[0], t[1], [float(v) for v in t[2:]]))
    return out


def _double2int(x, eps=1e-2):
    """util.cpp:22-28"""
    res = int(x + 0.5)
    if abs(res - x) > eps:
        raise RuntimeError("double2int() %r is not integer\n" % x)
    return res


def load_magnetic_field(fname):
    """Fields::load_magnetic_field (reference src/fields.cpp:870-959): rows "r z Br Bz" (anything else is skipped),
    r and z on a regular grid in either order -> dict(r_sampl, z_sampl, dr, dz, r_min, z_min, Br[r_sampl][z_sampl], Bz),
    what mag2d_set_magnetic_field takes.  Same error texts as the reference."""
    try:
        f = open(fname)
    except OSError:
        raise RuntimeError("Fields::load_magnetic_field(): failed opening file\n")
    rows = []
    with f:
        for line in f:
            t = line.split()
            try:
                rows.append((float(t[0]), float(t[1]), float(t[2]), float(t[3])))
            except (ValueError, IndexError):
                continue
    a = np.array(rows, dtype=np.float64).reshape(-1, 4)
    n = a.shape[0]
    if n < 2:
        raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")
    rv, zv = a[:, 0], a[:, 1]

    def axis(v):
        lo, hi = v[0], v[-1]
        nz = np.nonzero(np.diff(v) != 0.0)[0]
        if nz.size == 0:
            raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")
        d = v[nz[0] + 1] - v[nz[0]]
        if d < 0:
            d, lo, hi = -d, v[-1], v[0]
        return d, lo, _double2int((hi - lo) / d + 1)

    dr, r_min, r_sampl = axis(rv)
    dz, z_min, z_sampl = axis(zv)
    if r_sampl * z_sampl != n:
        raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")
    Br = np.full((r_sampl, z_sampl), np.nan)
    Bz = np.full((r_sampl, z_sampl), np.nan)
    for k in range(n):
        ri, zi = _double2int((rv[k] - r_min) / dr), _double2int((zv[k] - z_min) / dz)
        Br[ri, zi] = a[k, 2]
        Bz[ri, zi] = a[k, 3]
    if np.isnan(Br).any() or np.isnan(Bz).any():
        raise RuntimeError("Fields::load_magnetic_field() garbage loaded")
    return dict(r_sampl=r_sampl, z_sampl=z_sampl, dr=float(dr), dz=float(dz), r_min=float(r_min), z_min=float(z_min), Br=Br, Bz=Bz)


def load_field2d(fname):
    """Field2D::load (reference src/Field2D.cpp:46-130): rows "x y value" on a regular grid in either order (other
    lines are skipped) -> dict(M, N, dx, dz, x_min, z_min, data[M][N]); same error texts as the reference"""
    try:
        f = open(fname)
    except OSError:
        raise RuntimeError("Field2D::load(): failed opening file\n")
    rows = []
    with f:
        for line in f:
            t = line.split()
            try:
                rows.append((float(t[0]), float(t[1]), float(t[2])))
            except (ValueError, IndexError):
                continue
    a = np.array(rows, dtype=np.float64).reshape(-1, 3)
    n = a.shape[0]
    if n < 2:
        raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")

    def axis(v):
        lo, hi = v[0], v[-1]
        nz = np.nonzero(np.diff(v) != 0.0)[0]
        if nz.size == 0:
            raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")
        d = v[nz[0] + 1] - v[nz[0]]
        if d < 0:
            d, lo, hi = -d, v[-1], v[0]
        return d, lo, _double2int((hi - lo) / d + 1, 1e-1)

    dx, x_min, M = axis(a[:, 0])
    dz, z_min, N = axis(a[:, 1])
    if M * N != n:
        raise RuntimeError("Fields::load_magnetic_field() wrong size of input vector")
    data = np.full((M, N), np.nan)
    for k in range(n):
        data[_double2int((a[k, 0] - x_min) / dx, 1e-1), _double2int((a[k, 1] - z_min) / dz, 1e-1)] = a[k, 2]
    if np.isnan(data).any():
        raise RuntimeError("Fields::load_magnetic_field() garbage loaded")
    return dict(M=M, N=N, dx=float(dx), dz=float(dz), x_min=float(x_min), z_min=float(z_min), data=data)
