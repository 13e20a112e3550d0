"""ctypes window onto oracle/_ref/libmag2d_ref_*.so (the UNMODIFIED reference, see ref_harness.cpp).

TEST INFRASTRUCTURE ONLY.  The shared objects are built by ``make -C oracle ref`` in the container
that has /root/reference and then travel to the GPU box inside oracle/_ref/ (git-ignored).
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

PARAM_FIELDS = [
    "x_max", "y_max", "z_max", "x_min", "y_min", "z_min", "x_sampl", "y_sampl", "z_sampl",
    "extern_field", "electric_field_from_file", "magnetic_field_const", "Br", "Bz", "Bt",
    "has_probe", "probe_radius", "probe_length", "u_probe", "n_particles_total", "density_total",
    "dx", "dy", "dz", "V", "dV", "idx", "idy", "idz", "pressure", "neutral_temperature",
    "macroparticle_factor", "dt_elon", "niter", "mover", "coord", "boundary", "geometry", "src_fact",
    "selfconsistent", "use_source", "u_smooth", "rf", "rf_amplitude", "rf_U0", "rf_omega",
    "particle_reload", "t_print", "t_print_dist", "t_dist_sample", "t_equilib", "do_plot",
    "neutral_density",
]
SPECIES_FIELDS = ["type", "mass", "charge", "lifetime", "temperature", "E_max", "density", "v_max",
                  "dt", "t", "niter", "n_particles", "n_slots"]


def ref_available(variant="parity"):
    return os.path.exists(os.path.join(REF_DIR, "libmag2d_ref_%s.so" % variant))


_libs = {}


def _lib(variant):
    if variant in _libs:
        return _libs[variant]
    path = os.path.join(REF_DIR, "libmag2d_ref_%s.so" % variant)
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    lib.ref_last_error.restype = C.c_char_p
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_param_get.argtypes = [C.c_void_p, dp]
    lib.ref_n_species.argtypes = [C.c_void_p]
    lib.ref_species_name.restype = C.c_char_p
    lib.ref_species_name.argtypes = [C.c_void_p, C.c_int]
    lib.ref_species_get.argtypes = [C.c_void_p, C.c_int, dp]
    lib.ref_species_set.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_double]
    lib.ref_lifetime_init.argtypes = [C.c_void_p, C.c_int]
    lib.ref_species_rates.argtypes = [C.c_void_p, C.c_int, dp]
    lib.ref_n_interactions.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_interaction_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
    lib.ref_sigma_v.restype = C.c_double
    lib.ref_sigma_v.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
    lib.ref_run_initscript.argtypes = [C.c_void_p, C.c_char_p]
    lib.ref_set_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    lib.ref_get_particles.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    lib.ref_n_slots.argtypes = [C.c_void_p, C.c_int]
    lib.ref_advance_init.argtypes = [C.c_void_p]
    lib.ref_advance.argtypes = [C.c_void_p, C.c_int]
    lib.ref_advance_particles.argtypes = [C.c_void_p, C.c_int]
    lib.ref_advance_position.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ref_advance_boundary.argtypes = [C.c_void_p, C.c_int]
    lib.ref_species_accumulate.argtypes = [C.c_void_p, C.c_int]
    lib.ref_time_advance.restype = C.c_double
    lib.ref_time_advance.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    lib.ref_grid_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_get_field.argtypes = [C.c_void_p, C.c_char_p, dp]
    lib.ref_set_field.argtypes = [C.c_void_p, C.c_char_p, dp]
    lib.ref_field_op.argtypes = [C.c_void_p, C.c_char_p]
    lib.ref_field_E.argtypes = [C.c_void_p, C.c_int, dp, dp, C.c_double, dp, dp]
    lib.ref_species_save.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.ref_species_load.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    lib.ref_field2d_load.argtypes = [C.c_void_p, C.c_char_p, dp, dp, C.c_int]
    lib.ref_energy_hist.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp]
    lib.ref_source_refresh.argtypes = [C.c_void_p, C.c_int, C.c_uint]
    lib.ref_source.argtypes = [C.c_void_p, C.c_int]
    lib.ref_source_n.argtypes = [C.c_void_p, C.c_int]
    lib.ref_source_get.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    lib.ref_srand.argtypes = [C.c_uint]
    lib.ref_field_B.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp]
    lib.ref_btable_info.argtypes = [C.c_void_p, dp]
    lib.ref_btable_get.argtypes = [C.c_void_p, C.c_int, dp]
    lib.ref_field_accumulate.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_int, dp, dp]
    lib.ref_is_free.argtypes = [C.c_void_p, C.c_int, dp, dp, C.POINTER(C.c_int)]
    lib.ref_scatter.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    lib.ref_rng_seed.argtypes = [C.c_void_p, C.c_uint]
    lib.ref_rng_draw.argtypes = [C.c_void_p, C.c_char_p, C.c_int, dp]
    lib.ref_rng_rot.argtypes = [C.c_void_p, C.c_double, C.c_int, dp]
    lib.ref_rng_deflect.argtypes = [C.c_void_p, C.c_double, C.c_int, dp]
    _libs[variant] = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefHarness:
    """One reference ``Pic<CARTESIAN>`` / ``Pic<CYLINDRICAL>`` object (src/pic.cpp:115-189)."""

    def __init__(self, config, species_conf, overrides=None, seed=1234, variant="parity", output_dir=None):
        self.lib = _lib(variant)
        self._tmp = None
        if output_dir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="mag2d_ref_")
            output_dir = self._tmp.name
        ov = dict(do_plot=0)
        ov.update(overrides or {})
        ovs = ";".join("%s=%s" % (k, v) for k, v in ov.items())
        # the reference prints its seed and warnings on stdout; keep the test log quiet
        self.h = self.lib.ref_create(str(config).encode(), str(species_conf).encode(), output_dir.encode(),
                                     ovs.encode(), seed)
        if not self.h:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())
        M, N = C.c_int(), C.c_int()
        self.lib.ref_grid_dims(self.h, C.byref(M), C.byref(N))
        self.M, self.N = M.value, N.value

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None
        if self._tmp:
            self._tmp.cleanup()
            self._tmp = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())

    # ---- config
    def param(self):
        out = np.zeros(64)
        n = self.lib.ref_param_get(self.h, _dp(out))
        assert n == len(PARAM_FIELDS)
        return dict(zip(PARAM_FIELDS, out[:n].tolist()))

    def n_species(self):
        return self.lib.ref_n_species(self.h)

    def species_names(self):
        return [self.lib.ref_species_name(self.h, i).decode() for i in range(self.n_species())]

    def species_index(self, name):
        return self.species_names().index(name)

    def species(self, i):
        out = np.zeros(16)
        n = self.lib.ref_species_get(self.h, i, _dp(out))
        d = dict(zip(SPECIES_FIELDS, out[:n].tolist()))
        d["name"] = self.lib.ref_species_name(self.h, i).decode()
        return d

    def species_set(self, i, what, value):
        self._chk(self.lib.ref_species_set(self.h, i, what.encode(), float(value)))

    def lifetime_init(self, i):
        self._chk(self.lib.ref_lifetime_init(self.h, i))

    def rates(self, i):
        out = np.zeros(self.n_species())
        self.lib.ref_species_rates(self.h, i, _dp(out))
        return out

    def n_interactions(self, i, target):
        return self.lib.ref_n_interactions(self.h, i, target)

    def interaction(self, i, target, k):
        out = np.zeros(5)
        self.lib.ref_interaction_get(self.h, i, target, k, _dp(out))
        return dict(type=int(out[0]), DE=out[1], rate=out[2], cutoff=out[3])

    def sigma_v(self, i, target, k, v):
        return self.lib.ref_sigma_v(self.h, i, target, k, float(v))

    # ---- particles
    def run_initscript(self, path):
        self._chk(self.lib.ref_run_initscript(self.h, str(path).encode()))

    def set_particles(self, i, aos7):
        a = np.ascontiguousarray(aos7, dtype=np.float64).reshape(-1, 7)
        self._chk(self.lib.ref_set_particles(self.h, i, a.shape[0], _dp(a)))

    def get_particles(self, i):
        """-> (n_slots, 8) array: x,y,z,vx,vy,vz,time_to_death,alive"""
        n = self.lib.ref_n_slots(self.h, i)
        out = np.zeros((max(n, 1), 8))
        self.lib.ref_get_particles(self.h, i, _dp(out), n)
        return out[:n]

    def advance_init(self):
        self._chk(self.lib.ref_advance_init(self.h))

    def advance(self, nsteps=1):
        self._chk(self.lib.ref_advance(self.h, nsteps))

    def advance_particles(self, nsteps=1):
        self._chk(self.lib.ref_advance_particles(self.h, nsteps))

    def advance_position(self, i, init=False):
        self._chk(self.lib.ref_advance_position(self.h, i, 1 if init else 0))

    def advance_boundary(self, i):
        self._chk(self.lib.ref_advance_boundary(self.h, i))

    def species_accumulate(self, i):
        self._chk(self.lib.ref_species_accumulate(self.h, i))

    def time_advance(self, nsteps, particles_only=False):
        cpu = C.c_double()
        wall = self.lib.ref_time_advance(self.h, nsteps, 1 if particles_only else 0, C.byref(cpu))
        if wall < 0:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())
        return wall, cpu.value

    # ---- fields
    def get_field(self, which):
        out = np.zeros((self.M, self.N))
        self._chk(self.lib.ref_get_field(self.h, which.encode(), _dp(out)))
        return out

    def set_field(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self._chk(self.lib.ref_set_field(self.h, which.encode(), _dp(a)))

    def species_save(self, i, path):
        """BaseSpecies::save (src/particles.cpp:32-59)"""
        self._chk(self.lib.ref_species_save(self.h, i, path.encode()))

    def species_load(self, i, path):
        """BaseSpecies::load (src/particles.cpp:61-93)"""
        self._chk(self.lib.ref_species_load(self.h, i, path.encode()))

    def field2d_load(self, path, max_values=1 << 22):
        """Field2D::load (src/Field2D.cpp:46-130) -> dict(M, N, x_min, z_min, x_max, z_max, data)"""
        info = np.zeros(6)
        buf = np.zeros(max_values)
        self._chk(self.lib.ref_field2d_load(self.h, path.encode(), _dp(info), _dp(buf), max_values))
        M, N = int(info[0]), int(info[1])
        return dict(M=M, N=N, x_min=info[2], z_min=info[3], x_max=info[4], z_max=info[5], data=buf[:M * N].reshape(M, N).copy())

    def energy_hist(self, i, n_hist=200):
        """BaseSpecies::energy_dist after reset + energy_dist_compute -> (bins, dict(n_val, mean, mean_tot, norm, min, max))"""
        out, st = np.zeros(n_hist), np.zeros(6)
        self._chk(self.lib.ref_energy_hist(self.h, i, n_hist, _dp(out), _dp(st)))
        return out, dict(n_val=st[0], mean=st[1], mean_tot=st[2], norm=st[3], min=st[4], max=st[5])

    def field_op(self, op):
        self._chk(self.lib.ref_field_op(self.h, op.encode()))

    def field_E(self, x, z, time=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        ex = np.zeros_like(x)
        ez = np.zeros_like(x)
        self._chk(self.lib.ref_field_E(self.h, x.size, _dp(x), _dp(z), float(time), _dp(ex), _dp(ez)))
        return ex, ez

    def field_B(self, x, z):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.zeros((3, len(x)))
        self._chk(self.lib.ref_field_B(self.h, len(x), _dp(x), _dp(z), _dp(out[0]), _dp(out[1]), _dp(out[2])))
        return out

    def btable(self):
        """(info dict, Br, Bz) of the table loaded by Fields::load_magnetic_field"""
        o = np.zeros(6)
        self._chk(self.lib.ref_btable_info(self.h, _dp(o)))
        M, N = int(o[0]), int(o[1])
        br, bz = np.zeros((M, N)), np.zeros((M, N))
        self._chk(self.lib.ref_btable_get(self.h, 0, _dp(br)))
        self._chk(self.lib.ref_btable_get(self.h, 1, _dp(bz)))
        return dict(jmax=M, lmax=N, dx=o[2], dy=o[3], xmin=o[4], ymin=o[5]), br, bz

    # ---- particle source (use_source)
    def source_refresh(self, i, factor):
        self._chk(self.lib.ref_source_refresh(self.h, i, int(factor)))

    def source(self, i):
        self._chk(self.lib.ref_source(self.h, i))

    def source_particles(self, i):
        n = self.lib.ref_source_n(self.h, i)
        out = np.zeros((max(n, 1), 8))
        n = self.lib.ref_source_get(self.h, i, _dp(out), n)
        return out[:n]

    def srand(self, seed):
        self.lib.ref_srand(int(seed))

    def field_accumulate(self, which, charge, x, z):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        self._chk(self.lib.ref_field_accumulate(self.h, which.encode(), float(charge), x.size, _dp(x), _dp(z)))

    def is_free(self, x, z):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.zeros(x.size, dtype=np.int32)
        self.lib.ref_is_free(self.h, x.size, _dp(x), _dp(z), out.ctypes.data_as(C.POINTER(C.c_int)))
        return out.astype(bool)

    # ---- collisions / rng
    def scatter(self, i, v3):
        """v3: (n,3) columns vx,vy,vz (t_particle member order); returns the scattered copy"""
        a = np.ascontiguousarray(v3, dtype=np.float64).copy()
        self._chk(self.lib.ref_scatter(self.h, i, a.shape[0], _dp(a)))
        return a

    def rng_seed(self, seed):
        self.lib.ref_rng_seed(self.h, seed)

    def rng_draw(self, what, n):
        out = np.zeros(n)
        self.lib.ref_rng_draw(self.h, what.encode(), n, _dp(out))
        return out

    def rng_rot(self, length, n):
        out = np.zeros((n, 3))
        self.lib.ref_rng_rot(self.h, float(length), n, _dp(out))
        return out

    def rng_deflect(self, angle, v3):
        a = np.ascontiguousarray(v3, dtype=np.float64).copy()
        self.lib.ref_rng_deflect(self.h, float(angle), a.shape[0], _dp(a))
        return a


# ---------------------------------------------------------------------------------------------- 3-D
class Ref3D:
    """the compilable part of the reference's 3-D code (Field3D, Geometry, Solver) through oracle/ref3d_harness.cpp"""

    def __init__(self, config, with_solver=True):
        path = os.path.join(REF_DIR, "libmag3d_ref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libmag3d_ref.so is not built (make -C oracle ref)")
        L = C.CDLL(path)
        self.lib = L
        dp_ = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        i8p = C.POINTER(C.c_int8)
        L.ref3_create.restype = C.c_void_p
        L.ref3_create.argtypes = [C.c_char_p, C.c_int]
        L.ref3_error.restype = C.c_char_p
        L.ref3_error.argtypes = [C.c_void_p]
        L.ref3_destroy.argtypes = [C.c_void_p]
        L.ref3_dims.argtypes = [C.c_void_p, ip, dp_]
        L.ref3_mask.argtypes = [C.c_void_p, i8p, dp_]
        L.ref3_is_free.argtypes = [C.c_void_p, C.c_int, dp_, dp_, dp_, ip]
        L.ref3_accumulate.argtypes = [C.c_void_p, C.c_int, C.c_double, dp_, dp_, dp_]
        L.ref3_get.argtypes = [C.c_void_p, C.c_int, dp_]
        L.ref3_set.argtypes = [C.c_void_p, C.c_int, dp_]
        L.ref3_grad.argtypes = [C.c_void_p, C.c_int, dp_, dp_, dp_, dp_, dp_, dp_, dp_]
        L.ref3_solve.argtypes = [C.c_void_p]
        L.ref3_interpolate.argtypes = [C.c_void_p, C.c_int, dp_, dp_, dp_, dp_]
        self.h = L.ref3_create(config.encode(), 1 if with_solver else 0)
        err = L.ref3_error(self.h).decode()
        if err:
            raise RuntimeError(err)
        dims = (C.c_int * 3)()
        d = (C.c_double * 10)()
        L.ref3_dims(self.h, dims, d)
        self.shape = tuple(dims)
        (self.dx, self.dy, self.dz, self.idx, self.idy, self.idz, self.x_max, self.y_max, self.z_max,
         self.macroparticle_factor) = list(d)

    def close(self):
        if self.h:
            self.lib.ref3_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def mask(self):
        m = np.zeros(self.shape, dtype=np.int8)
        v = np.zeros(self.shape)
        self.lib.ref3_mask(self.h, m.ctypes.data_as(C.POINTER(C.c_int8)), _dp(v))
        return m, v

    def is_free(self, x, y, z):
        out = np.zeros(len(x), dtype=np.int32)
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        self.lib.ref3_is_free(self.h, len(x), _dp(x), _dp(y), _dp(z), out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    def accumulate(self, charge, x, y, z):
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        return self.lib.ref3_accumulate(self.h, len(x), charge, _dp(x), _dp(y), _dp(z))

    def get(self, which):
        out = np.zeros(self.shape)
        self.lib.ref3_get(self.h, {"u": 0, "rho": 1}[which], _dp(out))
        return out

    def set(self, which, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.shape
        self.lib.ref3_set(self.h, {"u": 0, "rho": 1}[which], _dp(a))

    def grad(self, x, y, z):
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        n = len(x)
        gx, gy, gz, val = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        self.lib.ref3_grad(self.h, n, _dp(x), _dp(y), _dp(z), _dp(gx), _dp(gy), _dp(gz), _dp(val))
        return np.stack([gx, gy, gz], axis=1), val

    def interpolate(self, x, y, z):
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        val = np.zeros(len(x))
        self.lib.ref3_interpolate(self.h, len(x), _dp(x), _dp(y), _dp(z), _dp(val))
        return val

    def solve(self):
        assert self.lib.ref3_solve(self.h) == 0
        return self.get("u")
