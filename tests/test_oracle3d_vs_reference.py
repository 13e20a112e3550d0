"""3-D oracle (oracle/mag3d_oracle.c) pinned against the parts of the reference's 3-D code that compile:
Field3D::accumulate / interpolate / grad (src/Field3D.hpp), Geometry (src/fields3d.cpp:13-37, fields3d.hpp:48-57)
and Solver::matrix_init / solve (src/fields3d.cpp:39-95, with the UMFPACK shim).  Species<CARTESIAN3D>::advance is
dead code in the reference (species3d.cpp does not compile), so the step itself is only checked for self-consistency."""
import os

import numpy as np
import pytest

from oracle import Oracle3, Orc3Grid, REF_DIR, Ref3D

needs_ref3 = os.path.exists(os.path.join(REF_DIR, "libmag3d_ref.so"))
pytestmark = pytest.mark.skipif(not needs_ref3, reason="oracle/_ref/libmag3d_ref.so not present on this machine")


def write_config(tmp_path, nx=9, ny=8, nz=7, x_max=8e-3, y_max=7e-3, z_max=6e-3, mpf=50.0):
    # Param derives dy from the macroparticle volume (param.cpp:133-137): pick n_particles_total / density_total
    # so that dy = y_max / (y_sampl - 1)
    dx, dz, dy = x_max / (nx - 1), z_max / (nz - 1), y_max / (ny - 1)
    dV = dx * dy * dz
    V = dV * (nx - 1) * (nz - 1)
    f = tmp_path / "config3d.txt"
    f.write_text("coord = CARTESIAN3D\nboundary = FREE\nmover = ADVANCE_BORIS\ngeometry = EMPTY\nselfconsistent = 1\n"
                 "x_sampl = %d\ny_sampl = %d\nz_sampl = %d\nr_max = %.17g\ny_max = %.17g\nz_max = %.17g\n"
                 "n_particles_total = 1e6\ndensity_total = %.17g\nmacroparticle_factor = %.17g\nrf = 0\n"
                 % (nx, ny, nz, x_max, y_max, z_max, 1e6 / V, mpf))
    return str(f)


@pytest.fixture()
def ref(tmp_path):
    with Ref3D(write_config(tmp_path)) as r:
        yield r


def grid_of(r):
    return Orc3Grid.make(r.shape, r.idx, r.idy, r.idz, r.x_max, r.y_max, r.z_max, 0, r.macroparticle_factor)


def test_param_gives_the_intended_spacing(ref):
    assert ref.shape == (9, 8, 7)
    assert ref.dy == pytest.approx(1e-3, rel=1e-12) and ref.dx == pytest.approx(1e-3) and ref.dz == pytest.approx(1e-3)


def test_geometry_mask_and_is_free(ref):
    orc = Oracle3()
    g = grid_of(ref)
    mask, volt = orc.geometry(g)
    rmask, rvolt = ref.mask()
    assert np.array_equal(mask, rmask)
    assert np.array_equal(volt, rvolt)
    # the k = kmax-1 face is FREE (fields3d.cpp:28 tests k == z_sampl) and the centre node is electrode -1 at 1 V
    assert (mask[1:-1, 1:-1, -1] == 2).all() and mask[4, 4, 3] == -1 and volt[4, 4, 3] == 1.0
    rng = np.random.default_rng(0)
    n = 4000
    x = rng.uniform(0, ref.x_max * (1 - 1e-9), n)
    y = rng.uniform(0, ref.y_max * (1 - 1e-9), n)
    z = rng.uniform(0, ref.z_max * (1 - 1e-9), n)
    assert np.array_equal(orc.is_free(g, mask, x, y, z), ref.is_free(x, y, z))


def test_deposit_and_interpolate_bit_exact(ref):
    orc = Oracle3()
    g = grid_of(ref)
    rng = np.random.default_rng(1)
    n = 3000
    x = rng.uniform(0, ref.x_max * (1 - 1e-9), n)
    y = rng.uniform(0, ref.y_max * (1 - 1e-9), n)
    z = rng.uniform(0, ref.z_max * (1 - 1e-9), n)
    assert ref.accumulate(-1.6e-19, x, y, z) == 0
    rho, bad = orc.accumulate(g, -1.6e-19, x, y, z)
    assert bad == 0 and np.array_equal(rho, ref.get("rho"))
    # fixed-point deposit against the fp64 one: |sum(Q32 w) 2^-32 q - rho| <= n_contrib 2^-33 |q| per node
    fixed, bad = orc.deposit_fixed(g, x, y, z)
    assert bad == 0
    assert np.abs(fixed * 2.0 ** -32 * -1.6e-19 - rho).max() <= n * 2.0 ** -33 * 1.6e-19
    assert fixed.sum() == pytest.approx(n * 2.0 ** 32, abs=8 * n)
    u = rng.normal(size=ref.shape)
    ref.set("u", u)
    assert np.array_equal(orc.interpolate(g, u, x, y, z), ref.interpolate(x, y, z))


def test_grad_matches_field3d(ref):
    orc = Oracle3()
    g = grid_of(ref)
    rng = np.random.default_rng(2)
    u = rng.normal(size=ref.shape)
    ref.set("u", u)
    n = 4000
    # the reference reads one plane past the array in the last half cell of each axis (weight 0): stay below it
    x = rng.uniform(0, ref.x_max - 0.51 * ref.dx, n)
    y = rng.uniform(0, ref.y_max - 0.51 * ref.dy, n)
    z = rng.uniform(0, ref.z_max - 0.51 * ref.dz, n)
    got = orc.grad(g, u, x, y, z)
    want, _ = ref.grad(x, y, z)
    assert np.array_equal(got, want)
    # linear potential: exact gradient everywhere, also in the clamped half cells at the upper faces
    i, j, k = np.meshgrid(*(np.arange(s) for s in ref.shape), indexing="ij")
    lin = 3.0 * i * ref.dx - 2.0 * j * ref.dy + 0.5 * k * ref.dz
    xe = rng.uniform(0, ref.x_max, 500)
    ye = rng.uniform(0, ref.y_max, 500)
    ze = rng.uniform(0, ref.z_max, 500)
    ge = orc.grad(g, lin, xe, ye, ze)
    assert np.allclose(ge, [3.0, -2.0, 0.5], rtol=0, atol=1e-9)


def test_rhs_operator_and_solve(ref):
    orc = Oracle3()
    g = grid_of(ref)
    mask, volt = orc.geometry(g)
    rng = np.random.default_rng(3)
    n = 2000
    x = rng.uniform(1e-3, ref.x_max - 1e-3, n)
    y = rng.uniform(1e-3, ref.y_max - 1e-3, n)
    z = rng.uniform(1e-3, ref.z_max - 1e-3, n)
    ref.accumulate(-1.6e-19, x, y, z)
    rho = ref.get("rho")
    u_ref = ref.solve()                         # scales rho in place into the right-hand side
    b_ref = ref.get("rho")
    b = orc.rhs(g, mask, volt, rho)
    assert np.array_equal(b, b_ref)
    u = orc.solve_direct(g, mask, b)
    assert np.abs(u - u_ref).max() <= 1e-11 * np.abs(u_ref).max()
    assert u[4, 4, 3] == pytest.approx(1.0, abs=1e-12)
    # the restated operator applied to the reference's solution reproduces the right-hand side
    assert np.abs(orc.apply_operator(g, mask, u_ref) - b_ref).max() <= 1e-9 * np.abs(b_ref).max()


def test_advance_restatement_is_self_consistent(ref):
    """species3d.cpp cannot be compiled; check the restated step against its parts: free flight in u = 0 moves
    particles ballistically, removes the ones that leave, and deposits exactly what accumulate() would"""
    orc = Oracle3()
    g = grid_of(ref)
    mask, _ = orc.geometry(g)
    rng = np.random.default_rng(4)
    n = 2000
    soa = {k: np.ascontiguousarray(v) for k, v in dict(
        x=rng.uniform(0, ref.x_max, n), y=rng.uniform(0, ref.y_max, n), z=rng.uniform(0, ref.z_max, n),
        vx=rng.normal(size=n) * 1e5, vy=rng.normal(size=n) * 1e5, vz=rng.normal(size=n) * 1e5).items()}
    start = {k: v.copy() for k, v in soa.items()}
    alive = np.ones(n, dtype=np.uint8)
    rho = np.zeros(ref.shape)
    fixed = np.zeros(ref.shape, dtype=np.int64)
    dt = 2e-9
    removed = orc.advance(g, np.zeros(ref.shape), mask, -1.6e-19, 9.11e-31, dt, (0, 0, 0), soa, alive, rho, fixed)
    assert removed == n - alive.sum() and 0 < removed < n
    live = alive > 0
    for c in "xyz":
        assert np.array_equal(soa[c][live], (start[c] + start["v" + c] * dt)[live])
    rho2, _ = orc.accumulate(g, -1.6e-19, soa["x"][live], soa["y"][live], soa["z"][live])
    assert np.array_equal(rho, rho2)
    fixed2, _ = orc.deposit_fixed(g, soa["x"][live], soa["y"][live], soa["z"][live])
    assert np.array_equal(fixed, fixed2)
