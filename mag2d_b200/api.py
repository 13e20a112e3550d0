"""ctypes binding of libmag2d_b200.so (include/mag2d_b200.h) and ``Sim``, a Python mirror of the
reference's ``Pic<D>`` driver surface (src/pic.cpp:115-538) on top of it.

All computation happens in the CUDA library; this module only describes runs and moves buffers.  The
library is required: importing works without it (so that CPU-only tooling can read configs), but
every call into ``lib()`` raises when the shared object or a CUDA device is missing — there is no
CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import config as cfg
from . import geometry

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAG2D_B200_LIB") or os.path.join(HERE, "libmag2d_b200.so")   # env: tuning builds

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
u8p = C.POINTER(C.c_uint8)


class GridDesc(C.Structure):
    _fields_ = [
        ("coord", C.c_int32), ("boundary", C.c_int32), ("mover", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("x_max", C.c_double), ("z_max", C.c_double), ("y_max", C.c_double),
        ("dx", C.c_double), ("dz", C.c_double), ("dy", C.c_double),
        ("idx", C.c_double), ("idz", C.c_double), ("idy", C.c_double),
        ("selfconsistent", C.c_int32), ("rf", C.c_int32), ("geometry_empty", C.c_int32),
        ("electric_field_from_file", C.c_int32),
        ("extern_field", C.c_double),
        ("rf_amplitude", C.c_double), ("rf_U0", C.c_double), ("rf_omega", C.c_double),
        ("magnetic_field_const", C.c_int32), ("u_smooth", C.c_int32),
        ("Br", C.c_double), ("Bz", C.c_double), ("Bt", C.c_double),
        ("dV", C.c_double), ("macroparticle_factor", C.c_double),
    ]


class SpeciesDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved0", C.c_int32), ("mass", C.c_double), ("charge", C.c_double),
                ("density", C.c_double), ("temperature", C.c_double), ("E_max", C.c_double), ("dt", C.c_double)]


class InteractionDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("primary", C.c_int32), ("secondary", C.c_int32), ("n_table", C.c_int32),
                ("table_offset", C.c_int32), ("reserved0", C.c_int32), ("DE_eV", C.c_double), ("rate", C.c_double),
                ("cutoff", C.c_double)]


PARTICLE_DTYPE = np.dtype([("x", "f8"), ("y", "f8"), ("z", "f8"), ("vx", "f8"), ("vy", "f8"), ("vz", "f8"),
                           ("time_to_death", "f8"), ("empty", "u1"), ("pad", "u1", (7,))])
assert PARTICLE_DTYPE.itemsize == 64

_lib = None


class Mag2dError(RuntimeError):
    pass


def lib():
    """load libmag2d_b200.so; raises when it is missing (build it with `python -m mag2d_b200.build`)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Mag2dError("libmag2d_b200.so is not built (python -m mag2d_b200.build); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.mag2d_last_error.restype = C.c_char_p
    L.mag2d_create.argtypes = [C.c_int, C.POINTER(GridDesc), vp, C.POINTER(vp)]
    L.mag2d_destroy.argtypes = [vp]
    L.mag2d_sync.argtypes = [vp]
    L.mag2d_seed.argtypes = [vp, C.c_uint64]
    L.mag2d_set_grid.argtypes = [vp, u8p, dp]
    L.mag2d_set_potential.argtypes = [vp, C.c_int, dp]
    L.mag2d_get_potential.argtypes = [vp, C.c_int, dp]
    L.mag2d_solve.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), dp]
    L.mag2d_set_solver.argtypes = [vp, C.c_int, C.c_double, C.c_int]
    L.mag2d_set_solver_kind.argtypes = [vp, C.c_int]
    L.mag2d_solver_is_direct.argtypes = [vp]
    L.mag2d_solver_stats.argtypes = [vp, C.POINTER(C.c_int), dp]
    L.mag2d_u_smooth.argtypes = [vp, C.c_int, C.c_double]
    L.mag2d_field_E.argtypes = [vp, C.c_int, dp, dp, C.c_double, dp, dp]
    L.mag2d_field_E3.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp]
    L.mag2d_set_species.argtypes = [vp, C.c_int, C.POINTER(SpeciesDesc), C.c_int, C.POINTER(InteractionDesc), dp, dp, C.c_int]
    L.mag2d_species_get.argtypes = [vp, C.c_int, C.c_int, dp]
    L.mag2d_species_rates.argtypes = [vp, C.c_int, dp]
    L.mag2d_collision_counts.argtypes = [vp, C.c_int, i64p, C.c_int]
    L.mag2d_set_collision_counting.argtypes = [vp, C.c_int]
    L.mag2d_reserve.argtypes = [vp, C.c_int, C.c_int64]
    L.mag2d_set_use_source.argtypes = [vp, C.c_int]
    L.mag2d_source_refresh.argtypes = [vp, C.c_int, C.c_uint32, C.c_double]
    L.mag2d_source_upload.argtypes = [vp, C.c_int, C.c_uint32, vp, C.c_int64]
    L.mag2d_source_download.argtypes = [vp, C.c_int, vp, C.c_int64, i64p]
    L.mag2d_species_source.argtypes = [vp, C.c_int, i64p]
    L.mag2d_set_magnetic_field.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    L.mag2d_field_B.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp]
    L.mag2d_particles_upload.argtypes = [vp, C.c_int, vp, C.c_int64]
    L.mag2d_particles_upload_soa.argtypes = [vp, C.c_int, C.c_int64, dp, dp, dp, dp, dp, dp, dp]
    L.mag2d_particles_download.argtypes = [vp, C.c_int, vp, C.c_int64, i64p]
    L.mag2d_particles_download_soa.argtypes = [vp, C.c_int, C.c_int64, dp, dp, dp, dp, dp, dp, dp, u8p, i64p]
    L.mag2d_particles_clear.argtypes = [vp, C.c_int]
    L.mag2d_count.argtypes = [vp, C.c_int, i64p, i64p]
    L.mag2d_particles_generate.argtypes = [vp, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double]
    L.mag2d_sort.argtypes = [vp, C.c_int]
    L.mag2d_set_sort_interval.argtypes = [vp, C.c_int]
    L.mag2d_step_streamed.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), i64p] + [C.POINTER(dp)] * 5 + [C.c_int64]
    L.mag2d_step_streamed3.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), i64p] + [C.POINTER(dp)] * 6 + [C.c_int64]
    L.mag2d_set_species_sort_interval.argtypes = [vp, C.c_int, C.c_int]
    L.mag2d_streamed_bytes.argtypes = [vp, i64p, i64p, C.c_int]
    L.mag2d_set_store_layout.argtypes = [vp, C.c_int]
    L.mag2d_set_storage.argtypes = [vp, C.c_int]
    L.mag2d_store_stats.argtypes = [vp, C.c_int, i64p]
    L.mag2d_advance_init.argtypes = [vp]
    L.mag2d_step.argtypes = [vp, C.c_int]
    L.mag2d_species_advance.argtypes = [vp, C.c_int]
    L.mag2d_species_advance_init.argtypes = [vp, C.c_int]
    L.mag2d_species_accumulate.argtypes = [vp, C.c_int]
    L.mag2d_rho_reset.argtypes = [vp, C.c_int]
    L.mag2d_rho_fixed_download.argtypes = [vp, C.c_int, i64p]
    L.mag2d_rho_download.argtypes = [vp, dp]
    L.mag2d_rho_upload.argtypes = [vp, C.c_int, i64p]
    L.mag2d_energy_hist.argtypes = [vp, C.c_int, C.c_int, C.c_double, dp, dp]
    L.mag2d_comm_unique_id.argtypes = [vp]
    L.mag2d_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.mag2d_comm_destroy.argtypes = [vp]
    L.mag2d_kernel_launches.argtypes = [vp, i64p]
    L.mag2d_set_timing.argtypes = [vp, C.c_int]
    L.mag2d_timers.argtypes = [vp, dp]
    L.mag2d_device_pointer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    _lib = L
    return L


def _d(a):
    return a.ctypes.data_as(dp)


def grid_desc_from_param(p):
    g = GridDesc()
    g.coord, g.boundary, g.mover = int(p["coord"]), int(p["boundary"]), int(p["mover"])
    g.M, g.N, g.K = int(p["x_sampl"]), int(p["z_sampl"]), int(p["y_sampl"])
    g.x_max, g.z_max, g.y_max = p["x_max"], p["z_max"], p["y_max"]
    g.dx, g.dz, g.dy = p["dx"], p["dz"], p["dy"]
    g.idx, g.idz, g.idy = p["idx"], p["idz"], p["idy"]
    g.selfconsistent, g.rf = int(p["selfconsistent"]), int(p["rf"])
    g.geometry_empty = int(p["geometry"] == cfg.GEOMETRY["EMPTY"])
    g.electric_field_from_file = int(p["electric_field_from_file"])
    g.extern_field = p["extern_field"]
    g.rf_amplitude, g.rf_U0, g.rf_omega = p["rf_amplitude"], p["rf_U0"], p["rf_omega"]
    g.magnetic_field_const, g.u_smooth = int(p["magnetic_field_const"]), int(p["u_smooth"])
    g.Br, g.Bz, g.Bt = p["Br"], p["Bz"], p["Bt"]
    g.dV, g.macroparticle_factor = p["dV"], p["macroparticle_factor"]
    return g


class Sim:
    """One simulation on one GPU: the calls a ``Pic<D>`` user makes (src/test.cpp:50-66)."""

    def __init__(self, config, species_conf, overrides=None, device=0, stream=None, seed=1234, presolve=True,
                 solver_tol=1e-13):
        self.L = lib()
        self.param = cfg.read_config(config, overrides)
        self.species, self.interactions = cfg.read_species(species_conf)
        self.names = [s["name"] for s in self.species]
        p = self.param
        self.M, self.N = int(p["x_sampl"]), int(p["z_sampl"])
        self.is3d = int(p["coord"]) == 2
        self.shape = (self.M, int(p["y_sampl"]), self.N) if self.is3d else (self.M, self.N)
        self.grid = grid_desc_from_param(p)
        h = C.c_void_p()
        # stream: a cudaStream_t handle (int); None -> the context owns a private non-blocking stream.  The
        # legacy default stream (handle 0) cannot be named through this argument.
        self._chk(self.L.mag2d_create(device, C.byref(self.grid), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.L.mag2d_seed(self.h, seed)
        self.mask, self.voltage = geometry.build_geometry(p)
        self._chk(self.L.mag2d_set_grid(self.h, self.mask.ctypes.data_as(u8p), _d(self.voltage)))
        self._set_species()
        self._chk(self.L.mag2d_set_use_source(self.h, int(bool(p.get("use_source", 0)))))
        self.btable = None
        if not self.is3d and not p["magnetic_field_const"]:
            # Pic ctor: if(!param.magnetic_field_const) field.load_magnetic_field(...)  (pic.cpp:148-149)
            self.set_magnetic_field(cfg.load_magnetic_field(p["magnetic_field_file"]))
        self.solver_tol = solver_tol
        self._chk(self.L.mag2d_set_solver(self.h, 0, solver_tol, 100))
        self.solve_info = {}
        if self.is3d:
            # plasma3d.cpp:44-46: the vacuum field of the electrodes is solved once at start-up
            if presolve:
                self.solve_info["u"] = self.solve(rf=False)
        elif p["electric_field_from_file"]:
            # Pic ctor: field.u.load(static file); if(rf) field.uRF.load(rf file)  (pic.cpp:154-177).  The reference adopts the
            # file's own grid; here the file has to carry the grid of config.txt (as the C++ host layer requires too)
            for which, key in (("u", "electric_field_static_file"),) + ((("uRF", "electric_field_rf_file"),) if p["rf"] else ()):
                f = cfg.load_field2d(p[key])
                if (f["M"], f["N"]) != (self.M, self.N) or abs(f["x_min"]) > 1e-9 * p["dx"] or abs(f["z_min"]) > 1e-9 * p["dz"]:
                    raise Mag2dError("Pic: field file grid differs from x_sampl/z_sampl (regridding is not implemented)")
                self.set_field(which, f["data"])
        elif presolve:
            # Pic ctor: boundary_solve_rf(); if(!selfconsistent){ boundary_solve(); reset(); }  (pic.cpp:180-187)
            self.solve_info["uRF"] = self.solve(rf=True)
            if not p["selfconsistent"]:
                self.solve_info["u"] = self.solve(rf=False)

    # ---- plumbing
    def _chk(self, rc):
        if rc:
            raise Mag2dError(self.L.mag2d_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mag2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _set_species(self):
        ns = len(self.species)
        sd = (SpeciesDesc * ns)()
        for i, s in enumerate(self.species):
            sd[i].type, sd[i].mass, sd[i].charge = s["type"], s["mass"], s["charge"]
            sd[i].density, sd[i].temperature, sd[i].E_max, sd[i].dt = s["density"], s["temperature"], s["E_max"], s["dt"]
        ni = len(self.interactions)
        idesc = (InteractionDesc * max(ni, 1))()
        tE, tS = [], []
        for k, q in enumerate(self.interactions):
            idesc[k].type = q["type"]
            idesc[k].primary = self.names.index(q["primary"])
            idesc[k].secondary = self.names.index(q["secondary"])
            idesc[k].n_table = len(q["CS_energy"])
            idesc[k].table_offset = len(tE)
            idesc[k].DE_eV, idesc[k].rate, idesc[k].cutoff = q["DE"], q["rate"], q["cutoff"]
            tE += q["CS_energy"]
            tS += q["CS_value"]
        tEa = np.ascontiguousarray(tE if tE else [0.0], dtype=np.float64)
        tSa = np.ascontiguousarray(tS if tS else [0.0], dtype=np.float64)
        self._chk(self.L.mag2d_set_species(self.h, ns, sd, ni, idesc, _d(tEa), _d(tSa), len(tE)))

    def set_magnetic_field(self, table):
        """table: what config.load_magnetic_field returns; None goes back to the constant (Br, Bz, Bt) of config.txt"""
        if table is None:
            self._chk(self.L.mag2d_set_magnetic_field(self.h, 0, 0, 0.0, 0.0, 0.0, 0.0, None, None))
        else:
            br = np.ascontiguousarray(table["Br"], dtype=np.float64)
            bz = np.ascontiguousarray(table["Bz"], dtype=np.float64)
            self._chk(self.L.mag2d_set_magnetic_field(self.h, table["r_sampl"], table["z_sampl"], table["dr"], table["dz"],
                                                      table["r_min"], table["z_min"], _d(br), _d(bz)))
        self.btable = table

    def field_B(self, x, z):
        """Fields::B at points (fields.hpp:152-177) -> array [3][n]: Br, Bz, Bt"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.zeros((3, x.size))
        self._chk(self.L.mag2d_field_B(self.h, x.size, _d(x), _d(z), _d(out[0]), _d(out[1]), _d(out[2])))
        return out

    # ---- configuration queries
    def species_index(self, name):
        return self.names.index(name)

    def species_get(self, i, what):
        keys = ["lifetime", "v_max", "E_max", "t", "niter", "prob"]
        out = C.c_double()
        self._chk(self.L.mag2d_species_get(self.h, i, keys.index(what), C.byref(out)))
        return out.value

    def rates(self, i):
        out = np.zeros(len(self.species))
        self._chk(self.L.mag2d_species_rates(self.h, i, _d(out)))
        return out

    # ---- particles
    def set_particles(self, i, aos7):
        """replace species i by the rows of aos7 (x,y,z,vx,vy,vz,time_to_death); slot k = row k"""
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        rec = np.zeros(a.shape[0], dtype=PARTICLE_DTYPE)
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "time_to_death")):
            rec[k] = a[:, c]
        self._chk(self.L.mag2d_particles_clear(self.h, i))
        self._chk(self.L.mag2d_particles_upload(self.h, i, rec.ctypes.data_as(C.c_void_p), rec.shape[0]))

    def add_particles_soa(self, i, x, z, vx, vy, vz, y=None, ttd=None):
        arrs = [np.ascontiguousarray(v, dtype=np.float64) if v is not None else None for v in (x, y, z, vx, vy, vz, ttd)]
        ptr = [_d(v) if v is not None else None for v in arrs]
        self._chk(self.L.mag2d_particles_upload_soa(self.h, i, arrs[0].size, *ptr))

    def get_particles(self, i):
        """-> (n_slots, 8): x,y,z,vx,vy,vz,time_to_death,alive in device slot order"""
        n = self.count(i)[1]
        rec = np.zeros(max(n, 1), dtype=PARTICLE_DTYPE)
        ns = C.c_int64()
        self._chk(self.L.mag2d_particles_download(self.h, i, rec.ctypes.data_as(C.c_void_p), rec.shape[0], C.byref(ns)))
        rec = rec[:ns.value]
        out = np.zeros((rec.shape[0], 8))
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "time_to_death")):
            out[:, c] = rec[k]
        out[:, 7] = 1 - rec["empty"]
        return out

    def get_particles_soa(self, i, what=("x", "y", "z", "vx", "vy", "vz")):
        """-> dict of float64 arrays in device slot order (only the requested components are copied) + 'alive' (uint8)"""
        n = self.count(i)[1]
        want = set(what) | {"x"}
        arrs = {k: (np.zeros(max(n, 1)) if k in want else None) for k in ("x", "y", "z", "vx", "vy", "vz", "ttd")}
        alive = np.zeros(max(n, 1), dtype=np.uint8)
        ns = C.c_int64()
        ptr = [_d(arrs[k]) if arrs[k] is not None else None for k in ("x", "y", "z", "vx", "vy", "vz", "ttd")]
        self._chk(self.L.mag2d_particles_download_soa(self.h, i, max(n, 1), *ptr, alive.ctypes.data_as(u8p), C.byref(ns)))
        out = {k: v[:ns.value] for k, v in arrs.items() if v is not None}
        out["alive"] = alive[:ns.value]
        return out

    def count(self, i):
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.L.mag2d_count(self.h, i, C.byref(a), C.byref(b)))
        return a.value, b.value

    def generate(self, i, kind, n, a=0.0, b=0.0, c=0.0, d=0.0):
        kinds = {"everywhere": 0, "on_disk": 1, "cylinder": 2}
        self._chk(self.L.mag2d_particles_generate(self.h, i, kinds[kind], int(n), a, b, c, d))

    def run_initscript(self, path):
        """the loader verbs of src/pic.cpp:257-325, executed by the device-side Philox loaders"""
        for verb, name, args in cfg.read_initscript(path):
            if name not in self.names:
                raise Mag2dError('Pic::run_initscript: unrecognized species type "%s"\n' % name)
            i = self.names.index(name)
            if verb == "add_particles_everywhere":
                self.generate(i, "everywhere", int(args[0]))
            elif verb == "add_particles_on_disk":
                self.generate(i, "on_disk", int(args[0]), args[1], args[2], args[3])
            elif verb == "add_tracked_particle":
                x, y, vx, vy, vz = args
                row = np.array([[x, 0.0, y, vx, vz, vy, 0.0]])   # particles.hpp:297-306: vy <-> vz swap
                a = np.ascontiguousarray(row[:, [0, 2, 3, 4, 5]])
                self.add_particles_soa(i, a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4])
            else:
                raise Mag2dError("initscript verb %s needs the host Bessel sampler (C++ host layer)" % verb)

    def sort(self, i):
        self._chk(self.L.mag2d_sort(self.h, i))

    def set_sort_interval(self, steps, species=None):
        """cell-sort period in pushes (0: never); species=None sets the context-wide value, else a per-species override"""
        if species is None:
            self._chk(self.L.mag2d_set_sort_interval(self.h, steps))
        else:
            self._chk(self.L.mag2d_set_species_sort_interval(self.h, species, steps))

    def set_storage(self, dtype):
        """'f64' (default) or 'f32': element type of the device-resident particle arrays (2-D Boris movers; before any particle is loaded)"""
        self._chk(self.L.mag2d_set_storage(self.h, {"f64": 0, "f32": 1}[dtype]))

    def set_store_layout(self, layout):
        """'auto' / 'bricks': CARTESIAN3D stores are binned by 4 x 4 x 4-cell brick inside advance(); 'slots': slot order + fused cell sort"""
        self._chk(self.L.mag2d_set_store_layout(self.h, {"auto": 0, "slots": 1, "bricks": 2}[layout]))

    def store_stats(self, i):
        out = np.zeros(8, dtype=np.int64)
        self._chk(self.L.mag2d_store_stats(self.h, i, out.ctypes.data_as(i64p)))
        return dict(rebinnings=int(out[0]), full_bins=int(out[1]), list_overflow=int(out[2]), leavers_last_step=int(out[3]),
                    bins=int(out[4]), fullest_bin=int(out[5]), fullest_bin_room=int(out[6]), slots_in_use=int(out[7]))

    # ---- stepping
    def advance_init(self):
        self._chk(self.L.mag2d_advance_init(self.h))

    # ---- particle source (use_source; Species<CARTESIAN>::source5_refresh / source)
    def source_refresh(self, i=None, factor=None):
        """what the driver does after advance_init (test.cpp:56-59): one reservoir per particle species"""
        factor = int(self.param["src_fact"] if factor is None else factor)
        for k in (range(len(self.species)) if i is None else [i]):
            if self.species[k]["type"] != 0:          # speclist[i]->particle: everything but NEUTRAL
                self._chk(self.L.mag2d_source_refresh(self.h, k, factor, float(self.param["V"])))

    def set_source_particles(self, i, factor, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        rec = np.zeros(a.shape[0], dtype=PARTICLE_DTYPE)
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "time_to_death")):
            rec[k] = a[:, c]
        self._chk(self.L.mag2d_source_upload(self.h, i, int(factor), rec.ctypes.data_as(C.c_void_p), rec.shape[0]))

    def get_source_particles(self, i):
        n = C.c_int64()
        self._chk(self.L.mag2d_source_download(self.h, i, None, 0, C.byref(n)))
        rec = np.zeros(max(n.value, 1), dtype=PARTICLE_DTYPE)
        self._chk(self.L.mag2d_source_download(self.h, i, rec.ctypes.data_as(C.c_void_p), rec.shape[0], C.byref(n)))
        rec = rec[:n.value]
        out = np.zeros((rec.shape[0], 8))
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "time_to_death")):
            out[:, c] = rec[k]
        out[:, 7] = 1 - rec["empty"]
        return out

    def species_source(self, i):
        n = C.c_int64()
        self._chk(self.L.mag2d_species_source(self.h, i, C.byref(n)))
        return n.value

    def advance(self, nsteps=1):
        self._chk(self.L.mag2d_step(self.h, nsteps))

    def step_streamed(self, species, n_slots, pointers, chunk_slots=0):
        """one Pic::advance with HOST-resident particles: ``pointers[q]`` = five integer addresses (x, z, vx, vy, vz) of
        species[q]'s host arrays (numpy ``.ctypes.data`` or torch ``.data_ptr()``, pinned for overlap); six
        (x, y, z, vx, vy, vz) for CARTESIAN3D"""
        n = len(species)
        sp = (C.c_int32 * n)(*species)
        ns = (C.c_int64 * n)(*n_slots)
        cols = []
        for a in range(6 if self.is3d else 5):
            cols.append((dp * n)(*[C.cast(pointers[q][a], dp) for q in range(n)]))
        if self.is3d:
            self._chk(self.L.mag2d_step_streamed3(self.h, n, sp, ns, *cols, chunk_slots))
        else:
            self._chk(self.L.mag2d_step_streamed(self.h, n, sp, ns, *cols, chunk_slots))

    def streamed_bytes(self, reset=False):
        """bytes the streamed steps have copied (host -> device, device -> host)"""
        a, b = C.c_int64(), C.c_int64()
        self._chk(self.L.mag2d_streamed_bytes(self.h, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    def species_advance(self, i):
        self._chk(self.L.mag2d_species_advance(self.h, i))

    def species_advance_init(self, i):
        self._chk(self.L.mag2d_species_advance_init(self.h, i))

    def species_accumulate(self, i):
        self._chk(self.L.mag2d_species_accumulate(self.h, i))

    def rho_reset(self, i=-1):
        self._chk(self.L.mag2d_rho_reset(self.h, i))

    def sync(self):
        self._chk(self.L.mag2d_sync(self.h))

    # ---- fields
    def solve(self, rf=False, tol=None, max_cycles=100):
        cyc, res = C.c_int(), C.c_double()
        self._chk(self.L.mag2d_solve(self.h, 1 if rf else 0, self.solver_tol if tol is None else tol, max_cycles,
                                     C.byref(cyc), C.byref(res)))
        return dict(cycles=cyc.value, resid=res.value)

    def set_solver(self, cycles_per_step=0, tol=1e-13, max_cycles=100):
        self._chk(self.L.mag2d_set_solver(self.h, cycles_per_step, tol, max_cycles))

    def set_solver_kind(self, kind):
        """'auto' (direct sine-transform solver when the grid separates, else multigrid), 'multigrid' or 'direct'"""
        self._chk(self.L.mag2d_set_solver_kind(self.h, {"auto": 0, "multigrid": 1, "mg": 1, "direct": 2}[kind]))

    def solver_is_direct(self):
        return bool(self.L.mag2d_solver_is_direct(self.h))

    def solver_stats(self):
        cyc, res = C.c_int(), C.c_double()
        self._chk(self.L.mag2d_solver_stats(self.h, C.byref(cyc), C.byref(res)))
        return dict(cycles=cyc.value, resid=res.value)

    def get_field(self, which):
        if which == "mask":
            return self.mask.copy()
        if which == "voltage":
            return self.voltage.copy()
        out = np.zeros(self.shape)
        if which in ("u", "uRF"):
            self._chk(self.L.mag2d_get_potential(self.h, 0 if which == "u" else 1, _d(out)))
        elif which == "rho":
            self._chk(self.L.mag2d_rho_download(self.h, _d(out)))
        else:
            raise KeyError(which)
        return out

    def set_field(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._chk(self.L.mag2d_set_potential(self.h, 0 if which == "u" else 1, _d(a)))

    def rho_fixed(self, i):
        out = np.zeros(self.shape, dtype=np.int64)
        self._chk(self.L.mag2d_rho_fixed_download(self.h, i, out.ctypes.data_as(i64p)))
        return out

    def rho_upload(self, i, a):
        a = np.ascontiguousarray(a, dtype=np.int64)
        self._chk(self.L.mag2d_rho_upload(self.h, i, a.ctypes.data_as(i64p)))

    def u_smooth(self, symmetry=False, radius=-1.0):
        self._chk(self.L.mag2d_u_smooth(self.h, 1 if symmetry else 0, radius))

    def field_E3(self, x, y, z):
        """ElMag3D::E at points (CARTESIAN3D) -> array [n, 3]"""
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        n = len(x)
        ex, ey, ez = np.zeros(n), np.zeros(n), np.zeros(n)
        self._chk(self.L.mag2d_field_E3(self.h, n, _d(x), _d(y), _d(z), _d(ex), _d(ey), _d(ez)))
        return np.stack([ex, ey, ez], axis=1)

    def field_E(self, x, z, time=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        ex, ez = np.zeros_like(x), np.zeros_like(x)
        self._chk(self.L.mag2d_field_E(self.h, x.size, _d(x), _d(z), float(time), _d(ex), _d(ez)))
        return ex, ez

    # ---- diagnostics / measurement
    def energy_hist(self, i, nbins=200, emax=None):
        emax = self.species_get(i, "E_max") if emax is None else emax
        hist, stats = np.zeros(nbins), np.zeros(4)
        self._chk(self.L.mag2d_energy_hist(self.h, i, nbins, emax, _d(hist), _d(stats)))
        return hist, dict(n_in=stats[0], sum_in=stats[1], n_tot=stats[2], sum_tot=stats[3],
                          mean_tot=stats[3] / stats[2] if stats[2] else float("nan"))

    def set_collision_counting(self, on=True):
        self._chk(self.L.mag2d_set_collision_counting(self.h, 1 if on else 0))

    def collision_counts(self, i, reset=False):
        out = np.zeros((len(self.species) + 1) * 16, dtype=np.int64)
        self._chk(self.L.mag2d_collision_counts(self.h, i, out.ctypes.data_as(i64p), 1 if reset else 0))
        return out

    def kernel_launches(self):
        n = C.c_int64()
        self._chk(self.L.mag2d_kernel_launches(self.h, C.byref(n)))
        return n.value

    def set_timing(self, on=True):
        self._chk(self.L.mag2d_set_timing(self.h, 1 if on else 0))

    def timers(self):
        out = np.zeros(5)
        self._chk(self.L.mag2d_timers(self.h, _d(out)))
        return dict(push=out[0], solve=out[1], sort=out[2], allreduce=out[3], total=out[4])

    def device_pointer(self, what):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._chk(self.L.mag2d_device_pointer(self.h, {"rho_fixed": 0, "u": 1, "uRF": 2}[what], C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        L = lib()
        if L.mag2d_comm_unique_id(buf):
            raise Mag2dError(L.mag2d_last_error().decode())
        return bytes(buf)

    def comm_init(self, rank, nranks, uid):
        buf = C.create_string_buffer(uid, 128)
        self._chk(self.L.mag2d_comm_init(self.h, rank, nranks, buf))
