"""ctypes binding of oracle/libmag2d_oracle.so (the plain-C restatement, mag2d_oracle.h).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmag2d_oracle.so")

dp = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_ubyte)
i64p = C.POINTER(C.c_int64)


class OrcGrid(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int),
        ("dx", C.c_double), ("dz", C.c_double), ("idx", C.c_double), ("idz", C.c_double),
        ("x_min", C.c_double), ("x_max", C.c_double), ("z_min", C.c_double), ("z_max", C.c_double),
        ("coord", C.c_int), ("boundary", C.c_int),
        ("selfconsistent", C.c_int), ("rf", C.c_int), ("geometry_empty", C.c_int), ("field_from_file", C.c_int),
        ("extern_field", C.c_double),
        ("rf_amplitude", C.c_double), ("rf_U0", C.c_double), ("rf_omega", C.c_double),
        ("Br", C.c_double), ("Bz", C.c_double), ("Bt", C.c_double),
        ("dV", C.c_double), ("macroparticle_factor", C.c_double),
    ]

    @classmethod
    def make(cls, M, N, x_max, z_max, coord=0, boundary=0, selfconsistent=0, rf=0, geometry_empty=0,
             extern_field=0.0, rf_amplitude=0.0, rf_U0=0.0, rf_omega=0.0, Br=0.0, Bz=0.0, Bt=0.0,
             dV=1.0, macroparticle_factor=1.0, field_from_file=0):
        g = cls()
        g.M, g.N = M, N
        g.dx = x_max / (M - 1)          # param.cpp:126-130
        g.dz = z_max / (N - 1)
        g.idx = 1.0 / g.dx
        g.idz = 1.0 / g.dz
        g.x_min = g.z_min = 0.0
        g.x_max, g.z_max = x_max, z_max
        g.coord, g.boundary = coord, boundary
        g.selfconsistent, g.rf, g.geometry_empty, g.field_from_file = selfconsistent, rf, geometry_empty, field_from_file
        g.extern_field = extern_field
        g.rf_amplitude, g.rf_U0, g.rf_omega = rf_amplitude, rf_U0, rf_omega
        g.Br, g.Bz, g.Bt = Br, Bz, Bt
        g.dV, g.macroparticle_factor = dV, macroparticle_factor
        return g


class OrcParticles(C.Structure):
    _fields_ = [("n", C.c_int)] + [(k, dp) for k in ("x", "y", "z", "vx", "vy", "vz", "ttd")] + [("alive", u8p)]


class OrcBTable(C.Structure):
    _fields_ = [("jmax", C.c_int), ("lmax", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("xmin", C.c_double), ("ymin", C.c_double), ("Br", dp), ("Bz", dp)]


class OrcRng(C.Structure):
    _fields_ = [
        ("jz", C.c_uint32), ("jsr", C.c_uint32), ("hz", C.c_int32), ("iz", C.c_uint32),
        ("kn", C.c_uint32 * 128), ("ke", C.c_uint32 * 256),
        ("wn", C.c_float * 128), ("fn", C.c_float * 128), ("we", C.c_float * 256), ("fe", C.c_float * 256),
        ("z", C.c_uint32), ("w", C.c_uint32), ("jcong", C.c_uint32),
        ("nfix_x", C.c_float), ("nfix_y", C.c_float),
    ]


def build_oracle(force=False):
    """Compile the restatement (gcc, -O2 -ffp-contract=off).  Building the checker is not using it."""
    srcs = [os.path.join(HERE, f) for f in ("mag2d_oracle.c", "mag3d_oracle.c")]
    deps = srcs + [os.path.join(HERE, f) for f in ("mag2d_oracle.h", "mag3d_oracle.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(f) for f in deps):
        return LIB
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB] + srcs + ["-lm"])
    return LIB


def _d(a):
    return a.ctypes.data_as(dp)


def _u8(a):
    return a.ctypes.data_as(u8p)


class Particles:
    """SoA particle arrays (numpy, owned here) + the C view handed to the oracle."""

    def __init__(self, n):
        self.n = n
        for k in ("x", "y", "z", "vx", "vy", "vz", "ttd"):
            setattr(self, k, np.zeros(n))
        self.alive = np.ones(n, dtype=np.uint8)

    @classmethod
    def from_aos7(cls, aos):
        aos = np.asarray(aos, dtype=np.float64).reshape(-1, 7)
        p = cls(aos.shape[0])
        for c, k in enumerate(("x", "y", "z", "vx", "vy", "vz", "ttd")):
            getattr(p, k)[:] = aos[:, c]
        return p

    def aos7(self):
        return np.stack([self.x, self.y, self.z, self.vx, self.vy, self.vz, self.ttd], axis=1)

    def copy(self):
        q = Particles(self.n)
        for k in ("x", "y", "z", "vx", "vy", "vz", "ttd", "alive"):
            getattr(q, k)[:] = getattr(self, k)
        return q

    def cview(self):
        v = OrcParticles()
        v.n = self.n
        for k in ("x", "y", "z", "vx", "vy", "vz", "ttd"):
            setattr(v, k, _d(getattr(self, k)))
        v.alive = _u8(self.alive)
        return v


class Model:
    """Species + interaction wiring (orc_model)."""

    def __init__(self, lib, n_species):
        self.lib = lib
        self.ns = n_species
        self.h = lib.orc_model_new(n_species)
        self._keep = []

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_model_free(self.h)
            self.h = None

    def set_species(self, i, type, mass, charge, density, temperature, E_max, dt):
        assert self.lib.orc_model_set_species(self.h, i, type, mass, charge, density, temperature, E_max, dt) == 0

    def add_interaction(self, type, DE_eV, rate, cutoff, primary, secondary, E=None, sigma=None):
        n = 0 if E is None else len(E)
        Ea = np.ascontiguousarray(E if n else [0.0], dtype=np.float64)
        Sa = np.ascontiguousarray(sigma if n else [0.0], dtype=np.float64)
        assert self.lib.orc_model_add_interaction(self.h, type, DE_eV, rate, cutoff, primary, secondary, n, _d(Ea), _d(Sa)) == 0

    def set_pool(self, i, particles):
        self._keep.append(particles)
        self.lib.orc_model_set_pool(self.h, i, particles.n, _d(particles.vx), _d(particles.vy), _d(particles.vz), _u8(particles.alive))

    def lifetime_init(self):
        self.lib.orc_model_lifetime_init(self.h)

    def lifetime(self, i):
        return self.lib.orc_model_lifetime(self.h, i)

    def get(self, i, what):
        keys = ["mass", "charge", "density", "temperature", "E_max", "dt", "v_max", "lifetime"]
        return self.lib.orc_model_get(self.h, i, keys.index(what))

    def rates(self, i):
        out = np.zeros(self.ns)
        self.lib.orc_model_rates(self.h, i, _d(out))
        return out

    def sigma_v(self, primary, target, k, v):
        return self.lib.orc_sigma_v(self.h, primary, target, k, float(v))


class Oracle:
    def __init__(self):
        build_oracle()
        lib = C.CDLL(LIB)
        self.lib = lib
        G = C.POINTER(OrcGrid)
        R = C.POINTER(OrcRng)
        P = C.POINTER(OrcParticles)
        lib.orc_grad.argtypes = [dp, C.c_int, C.c_int] + [C.c_double] * 6 + [dp, dp]
        lib.orc_interpolate.restype = C.c_double
        lib.orc_interpolate.argtypes = [dp, C.c_int, C.c_int] + [C.c_double] * 6
        lib.orc_field_E.argtypes = [G, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
        for name, nptr in (("orc_boris_cart", 5), ("orc_boris_cart_init", 3), ("orc_boris_cyl", 5), ("orc_boris_cyl_init", 3)):
            getattr(lib, name).argtypes = [C.c_double] * 8 + [dp] * nptr
        lib.orc_rng_init.argtypes = [R, C.c_uint32]
        lib.orc_rng_seed.argtypes = [R, C.c_uint32]
        lib.orc_rng_draw.argtypes = [R, C.c_int, C.c_int, dp]
        lib.orc_rng_rot.argtypes = [R, C.c_double, dp, dp, dp]
        lib.orc_rng_deflect.argtypes = [R, C.c_double, dp, dp, dp]
        lib.orc_ellint_K.restype = C.c_double
        lib.orc_ellint_K.argtypes = [C.c_double]
        lib.orc_langevin_chi.restype = C.c_double
        lib.orc_langevin_chi.argtypes = [C.c_double]
        lib.orc_model_new.restype = C.c_void_p
        lib.orc_model_new.argtypes = [C.c_int]
        lib.orc_model_free.argtypes = [C.c_void_p]
        lib.orc_model_set_species.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 6
        lib.orc_model_add_interaction.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp]
        lib.orc_model_set_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp, u8p]
        lib.orc_model_lifetime_init.argtypes = [C.c_void_p]
        lib.orc_model_lifetime.restype = C.c_double
        lib.orc_model_lifetime.argtypes = [C.c_void_p, C.c_int]
        lib.orc_model_get.restype = C.c_double
        lib.orc_model_get.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_model_rates.argtypes = [C.c_void_p, C.c_int, dp]
        lib.orc_model_n_interactions.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_sigma_v.restype = C.c_double
        lib.orc_sigma_v.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        lib.orc_table_lookup.restype = C.c_double
        lib.orc_table_lookup.argtypes = [C.c_int, dp, dp, C.c_double]
        lib.orc_scatter.argtypes = [C.c_void_p, C.c_int, R, dp, dp, dp, C.POINTER(C.c_int)]
        lib.orc_advance_boris.argtypes = [G, dp, dp, C.c_void_p, C.c_int, P, C.c_ulong, R, i64p]
        lib.orc_advance_boris_init.argtypes = [G, dp, dp, C.c_void_p, C.c_int, P, C.c_ulong]
        lib.orc_advance_boris_extern.argtypes = [G, C.c_void_p, C.c_int, P, R, i64p, C.c_int]
        lib.orc_source_size.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint]
        lib.orc_source_size.restype = C.c_uint
        lib.orc_source_refresh.argtypes = [G, C.c_void_p, C.c_int, C.c_uint, R, P]
        lib.orc_source.argtypes = [G, C.c_void_p, C.c_int, C.c_uint, P, R, i64p, P, C.POINTER(C.c_int), C.c_int, dp, i64p, C.c_void_p]
        lib.orc_source.restype = C.c_int
        BT = C.POINTER(OrcBTable)
        lib.orc_btable_build.argtypes = [C.c_int, dp, dp, dp, dp, BT]
        lib.orc_btable_build.restype = C.c_int
        lib.orc_btable_free.argtypes = [BT]
        lib.orc_field_B.argtypes = [G, BT, C.c_double, C.c_double, dp, dp, dp]
        lib.orc_advance_boris_B.argtypes = [G, BT, dp, dp, C.c_void_p, C.c_int, P, C.c_ulong, R, i64p]
        lib.orc_advance_boris_init_B.argtypes = [G, BT, dp, dp, C.c_void_p, C.c_int, P, C.c_ulong]
        lib.orc_advance_multicoll.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int, P, R, i64p]
        lib.orc_advance_boundary.argtypes = [G, u8p, C.c_double, P, dp, i64p]
        lib.orc_deposit_fp64.argtypes = [G, C.c_double, C.c_int, dp, dp, u8p, dp]
        lib.orc_deposit_fixed.argtypes = [G, C.c_int, dp, dp, u8p, i64p]
        lib.orc_is_free.argtypes = [G, u8p, C.c_double, C.c_double]
        lib.orc_rhs.argtypes = [G, u8p, dp, C.c_int, dp]
        lib.orc_apply_operator.argtypes = [G, u8p, dp, dp]
        lib.orc_solve_direct.argtypes = [G, u8p, dp, dp]
        lib.orc_u_smooth.argtypes = [G, C.c_int, C.c_double, dp]
        lib.orc_geometry.argtypes = [G, C.c_int, C.c_double, C.c_double, u8p, dp]

    # ---- helpers
    def model(self, n_species):
        return Model(self.lib, n_species)

    def rng(self, seed):
        r = OrcRng()
        self.lib.orc_rng_init(C.byref(r), seed)
        return r

    def rng_seed(self, r, seed):
        self.lib.orc_rng_seed(C.byref(r), seed)

    def rng_draw(self, r, what, n):
        out = np.zeros(n)
        self.lib.orc_rng_draw(C.byref(r), ["uni", "rnor", "rexp", "iuni", "radius"].index(what), n, _d(out))
        return out

    def rng_rot(self, r, length, n):
        out = np.zeros((n, 3))
        for k in range(n):
            x, y, z = C.c_double(), C.c_double(), C.c_double()
            self.lib.orc_rng_rot(C.byref(r), length, C.byref(x), C.byref(y), C.byref(z))
            out[k] = (x.value, y.value, z.value)
        return out

    def rng_deflect(self, r, angle, v3):
        out = np.array(v3, dtype=np.float64).reshape(-1, 3).copy()
        for k in range(out.shape[0]):
            x, y, z = (C.c_double(out[k, 0]), C.c_double(out[k, 1]), C.c_double(out[k, 2]))
            self.lib.orc_rng_deflect(C.byref(r), angle, C.byref(x), C.byref(y), C.byref(z))
            out[k] = (x.value, y.value, z.value)
        return out

    def field_E(self, g, u, uRF, x, z, time=0.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        ex = np.zeros_like(x)
        ez = np.zeros_like(x)
        a, b = C.c_double(), C.c_double()
        for k in range(x.size):
            a.value, b.value = 0.0, 0.0
            self.lib.orc_field_E(C.byref(g), _d(u), _d(uRF), x[k], z[k], time, C.byref(a), C.byref(b))
            ex[k], ez[k] = a.value, b.value
        return ex, ez

    def scatter(self, model, primary, rng, v3):
        """v3 (n,3) columns vx,vy,vz; returns (scattered copy, process ids, target ids)"""
        out = np.array(v3, dtype=np.float64).reshape(-1, 3).copy()
        proc = np.zeros(out.shape[0], dtype=np.int32)
        targ = np.zeros(out.shape[0], dtype=np.int32)
        t = C.c_int()
        for k in range(out.shape[0]):
            vx, vy, vz = C.c_double(out[k, 0]), C.c_double(out[k, 1]), C.c_double(out[k, 2])
            proc[k] = self.lib.orc_scatter(model.h, primary, C.byref(rng), C.byref(vx), C.byref(vy), C.byref(vz), C.byref(t))
            targ[k] = t.value
            out[k] = (vx.value, vy.value, vz.value)
        return out, proc, targ

    def advance_boris(self, g, u, uRF, model, sp, particles, niter=0, rng=None, counts=None, btable=None):
        v = particles.cview()
        self.lib.orc_advance_boris_B(C.byref(g), C.byref(btable) if btable is not None else None, _d(u), _d(uRF), model.h, sp,
                                     C.byref(v), niter, C.byref(rng) if rng is not None else None,
                                     counts.ctypes.data_as(i64p) if counts is not None else None)

    def advance_boris_init(self, g, u, uRF, model, sp, particles, niter=0, btable=None):
        v = particles.cview()
        self.lib.orc_advance_boris_init_B(C.byref(g), C.byref(btable) if btable is not None else None, _d(u), _d(uRF), model.h, sp,
                                          C.byref(v), niter)

    # ---- particle source (use_source; particles.cpp:1053-1080, 1158-1226)
    def advance_boris_extern(self, g, model, sp, particles, rng=None, counts=None, init=False):
        v = particles.cview()
        self.lib.orc_advance_boris_extern(C.byref(g), model.h, sp, C.byref(v), C.byref(rng) if rng is not None else None,
                                          counts.ctypes.data_as(i64p) if counts is not None else None, int(init))

    def source_refresh(self, g, model, sp, factor, V, rng):
        n = self.lib.orc_source_size(model.h, sp, float(V), int(factor))
        src = Particles(n)
        v = src.cview()
        self.lib.orc_source_refresh(C.byref(g), model.h, sp, int(factor), C.byref(rng), C.byref(v))
        return src

    def source(self, g, model, sp, factor, src, dst, n_dst, rng=None, counts=None, rho=None, rho_fixed=None, libc_seed=None):
        """-> (injected, new n_dst).  dst: Particles with spare slots from n_dst on.  The lateral shifts come from libc
        rand(), as in the reference (srand(libc_seed) first when given)"""
        libc = C.CDLL(None)
        if libc_seed is not None:
            libc.srand(int(libc_seed))
        irand = C.cast(libc.rand, C.c_void_p)
        vs, vd = src.cview(), dst.cview()
        nd = C.c_int(int(n_dst))
        r = self.lib.orc_source(C.byref(g), model.h, sp, int(factor), C.byref(vs), C.byref(rng) if rng is not None else None,
                                counts.ctypes.data_as(i64p) if counts is not None else None, C.byref(vd), C.byref(nd), dst.n,
                                _d(rho) if rho is not None else None,
                                rho_fixed.ctypes.data_as(i64p) if rho_fixed is not None else None, irand)
        if r < 0:
            raise RuntimeError("orc_source: destination store is full")
        return r, nd.value

    # ---- magnetic field table (Fields::load_magnetic_field, fields.cpp:870-959)
    def load_magnetic_field(self, fname):
        """the file loop of fields.cpp:882-896 (rows with four numbers; everything else is skipped) + orc_btable_build"""
        rows = []
        with open(fname) as f:
            for line in f:
                t = line.split()
                try:
                    rows.append([float(t[0]), float(t[1]), float(t[2]), float(t[3])])
                except (ValueError, IndexError):
                    continue
        a = np.ascontiguousarray(np.array(rows, dtype=np.float64).T)
        bt = OrcBTable()
        rc = self.lib.orc_btable_build(a.shape[1], _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]), C.byref(bt))
        if rc:
            raise RuntimeError({1: "Fields::load_magnetic_field() wrong size of input vector",
                                2: "Fields::load_magnetic_field() garbage loaded", 3: "double2int() is not integer"}[rc])
        return bt

    def btable_arrays(self, bt):
        n = bt.jmax * bt.lmax
        return (np.ctypeslib.as_array(bt.Br, shape=(n,)).reshape(bt.jmax, bt.lmax).copy(),
                np.ctypeslib.as_array(bt.Bz, shape=(n,)).reshape(bt.jmax, bt.lmax).copy())

    def field_B(self, g, bt, x, z):
        out = np.zeros((3, len(x)))
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        for k in range(len(x)):
            self.lib.orc_field_B(C.byref(g), C.byref(bt) if bt is not None else None, float(x[k]), float(z[k]), C.byref(a), C.byref(b), C.byref(c))
            out[:, k] = a.value, b.value, c.value
        return out

    def advance_multicoll(self, fx, fz, model, sp, particles, rng, counts=None):
        v = particles.cview()
        self.lib.orc_advance_multicoll(fx, fz, model.h, sp, C.byref(v), C.byref(rng),
                                       counts.ctypes.data_as(i64p) if counts is not None else None)

    def advance_boundary(self, g, mask, charge, particles, rho=None, rho_fixed=None):
        v = particles.cview()
        return self.lib.orc_advance_boundary(C.byref(g), _u8(mask), charge, C.byref(v),
                                             _d(rho) if rho is not None else None,
                                             rho_fixed.ctypes.data_as(i64p) if rho_fixed is not None else None)

    def deposit_fp64(self, g, charge, x, z, alive=None):
        rho = np.zeros((g.M, g.N))
        bad = self.lib.orc_deposit_fp64(C.byref(g), charge, x.size, _d(x), _d(z), _u8(alive) if alive is not None else None, _d(rho))
        return rho, bad

    def deposit_fixed(self, g, x, z, alive=None):
        rho = np.zeros((g.M, g.N), dtype=np.int64)
        bad = self.lib.orc_deposit_fixed(C.byref(g), x.size, _d(x), _d(z), _u8(alive) if alive is not None else None,
                                         rho.ctypes.data_as(i64p))
        return rho, bad

    def is_free(self, g, mask, x, z):
        return np.array([self.lib.orc_is_free(C.byref(g), _u8(mask), float(a), float(b)) for a, b in zip(x, z)], dtype=bool)

    def rhs(self, g, mask, voltage, rho, rf=False):
        b = np.array(rho, dtype=np.float64, copy=True)
        self.lib.orc_rhs(C.byref(g), _u8(mask), _d(voltage), 1 if rf else 0, _d(b))
        return b

    def apply_operator(self, g, mask, u):
        y = np.zeros_like(u)
        self.lib.orc_apply_operator(C.byref(g), _u8(mask), _d(np.ascontiguousarray(u)), _d(y))
        return y

    def solve_direct(self, g, mask, b):
        u = np.zeros_like(b)
        assert self.lib.orc_solve_direct(C.byref(g), _u8(mask), _d(np.ascontiguousarray(b)), _d(u)) == 0
        return u

    def u_smooth(self, g, u, symmetry=False, radius=-1.0):
        v = np.array(u, dtype=np.float64, copy=True)
        self.lib.orc_u_smooth(C.byref(g), 1 if symmetry else 0, radius, _d(v))
        return v

    def geometry(self, g, geometry, probe_radius=1e-4, u_probe=-10.0):
        mask = np.zeros((g.M, g.N), dtype=np.uint8)
        voltage = np.zeros((g.M, g.N))
        self.lib.orc_geometry(C.byref(g), geometry, probe_radius, u_probe, _u8(mask), _d(voltage))
        return mask, voltage


# ---------------------------------------------------------------------------------------------- 3-D
class Orc3Grid(C.Structure):
    _fields_ = [("imax", C.c_int), ("jmax", C.c_int), ("kmax", C.c_int),
                ("idx", C.c_double), ("idy", C.c_double), ("idz", C.c_double),
                ("x_max", C.c_double), ("y_max", C.c_double), ("z_max", C.c_double),
                ("boundary", C.c_int), ("macroparticle_factor", C.c_double)]

    @classmethod
    def make(cls, dims, idx, idy, idz, x_max, y_max, z_max, boundary=0, mpf=1.0):
        g = cls()
        g.imax, g.jmax, g.kmax = (int(v) for v in dims)
        g.idx, g.idy, g.idz = idx, idy, idz
        g.x_max, g.y_max, g.z_max = x_max, y_max, z_max
        g.boundary = boundary
        g.macroparticle_factor = mpf
        return g

    @property
    def shape(self):
        return (self.imax, self.jmax, self.kmax)


class Oracle3:
    """ctypes view of oracle/mag3d_oracle.c (the restatement of the reference's 3-D path)"""

    def __init__(self):
        build_oracle()
        lib = C.CDLL(LIB)
        self.lib = lib
        G = C.POINTER(Orc3Grid)
        i8p = C.POINTER(C.c_int8)
        i64p = C.POINTER(C.c_int64)
        self._i8p, self._i64p = i8p, i64p
        lib.orc3_geometry.argtypes = [G, i8p, dp]
        lib.orc3_is_free.argtypes = [G, i8p, C.c_double, C.c_double, C.c_double]
        lib.orc3_accumulate.argtypes = [G, dp, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.orc3_deposit_fixed.argtypes = [G, C.c_int, dp, dp, dp, u8p, i64p]
        lib.orc3_interpolate.restype = C.c_double
        lib.orc3_interpolate.argtypes = [G, dp, C.c_double, C.c_double, C.c_double]
        lib.orc3_grad.argtypes = [G, dp, C.c_double, C.c_double, C.c_double, dp, dp, dp]
        lib.orc3_rhs.argtypes = [G, i8p, dp, dp]
        lib.orc3_apply_operator.argtypes = [G, i8p, dp, dp]
        lib.orc3_solve_direct.argtypes = [G, i8p, dp, dp]
        lib.orc3_advance.argtypes = [G, dp, i8p] + [C.c_double] * 6 + [C.c_int] + [dp] * 6 + [u8p, dp, i64p]

    def geometry(self, g):
        mask = np.zeros(g.shape, dtype=np.int8)
        volt = np.zeros(g.shape)
        self.lib.orc3_geometry(C.byref(g), mask.ctypes.data_as(self._i8p), _d(volt))
        return mask, volt

    def is_free(self, g, mask, x, y, z):
        m = mask.ctypes.data_as(self._i8p)
        return np.array([self.lib.orc3_is_free(C.byref(g), m, a, b, c) for a, b, c in zip(x, y, z)], dtype=np.int32)

    def accumulate(self, g, charge, x, y, z):
        rho = np.zeros(g.shape)
        bad = sum(self.lib.orc3_accumulate(C.byref(g), _d(rho), charge, a, b, c) != 0 for a, b, c in zip(x, y, z))
        return rho, bad

    def deposit_fixed(self, g, x, y, z, alive=None):
        out = np.zeros(g.shape, dtype=np.int64)
        x, y, z = (np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, z))
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        bad = self.lib.orc3_deposit_fixed(C.byref(g), len(x), _d(x), _d(y), _d(z), None if al is None else _u8(al),
                                          out.ctypes.data_as(self._i64p))
        return out, bad

    def grad(self, g, u, x, y, z):
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.zeros((len(x), 3))
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        for p in range(len(x)):
            self.lib.orc3_grad(C.byref(g), _d(u), x[p], y[p], z[p], C.byref(a), C.byref(b), C.byref(c))
            out[p] = (a.value, b.value, c.value)
        return out

    def interpolate(self, g, u, x, y, z):
        u = np.ascontiguousarray(u, dtype=np.float64)
        return np.array([self.lib.orc3_interpolate(C.byref(g), _d(u), a, b, c) for a, b, c in zip(x, y, z)])

    def rhs(self, g, mask, voltage, rho):
        b = np.array(rho, dtype=np.float64, copy=True)
        self.lib.orc3_rhs(C.byref(g), mask.ctypes.data_as(self._i8p), _d(np.ascontiguousarray(voltage)), _d(b))
        return b

    def apply_operator(self, g, mask, u):
        y = np.zeros(g.shape)
        self.lib.orc3_apply_operator(C.byref(g), mask.ctypes.data_as(self._i8p), _d(np.ascontiguousarray(u)), _d(y))
        return y

    def solve_direct(self, g, mask, b):
        u = np.zeros(g.shape)
        rc = self.lib.orc3_solve_direct(C.byref(g), mask.ctypes.data_as(self._i8p), _d(np.ascontiguousarray(b)), _d(u))
        assert rc == 0
        return u

    def advance(self, g, u, mask, charge, mass, dt, B, soa, alive, rho=None, rho_fixed=None):
        """soa: dict of contiguous float64 arrays x,y,z,vx,vy,vz (updated in place); alive uint8 (in place)"""
        n = len(alive)
        return self.lib.orc3_advance(C.byref(g), _d(np.ascontiguousarray(u)), mask.ctypes.data_as(self._i8p), charge, mass, dt,
                                     B[0], B[1], B[2], n, _d(soa["x"]), _d(soa["y"]), _d(soa["z"]), _d(soa["vx"]),
                                     _d(soa["vy"]), _d(soa["vz"]), _u8(alive), None if rho is None else _d(rho),
                                     None if rho_fixed is None else rho_fixed.ctypes.data_as(self._i64p))
