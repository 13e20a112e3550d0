/* oracle/mag3d_oracle.h — TEST INFRASTRUCTURE ONLY (see mag2d_oracle.h).
 *
 * CPU restatement of the reference's 3-D path.  The reference's 3-D step does not compile
 * (species3d.cpp is dead code), so Species<CARTESIAN3D>::advance is restated from src/species3d.cpp:3-93; the
 * field classes it calls DO compile (Field3D.hpp, fields3d.cpp) and the restatement of those is pinned against
 * them through oracle/ref3d_harness.cpp (tests/test_oracle3d_vs_reference.py).
 * Grids are the reference's Array3D layout: a[(i*jmax + j)*kmax + k], i along x, j along y, k along z. */
#ifndef MAG3D_ORACLE_H
#define MAG3D_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    int imax, jmax, kmax;          /* x_sampl, y_sampl, z_sampl */
    double idx, idy, idz;          /* 1/dx, 1/dy, 1/dz as Param computed them */
    double x_max, y_max, z_max;    /* Param::x_max.. (boundary test of species3d.cpp:64-66) */
    int boundary;                  /* 0 FREE, 1 PERIODIC */
    double macroparticle_factor;
} orc3_grid;

/* Geometry::Geometry, src/fields3d.cpp:13-37: zero-Dirichlet box frame (the k == z_sampl test of :28 never fires,
 * so the k = kmax-1 face stays FREE) plus the one-node Quadrupole electrode (id -1, 1 V) at the centre */
void orc3_geometry(const orc3_grid* g, signed char* mask, double* voltage);
/* Geometry::is_free, src/fields3d.hpp:48-57 */
int orc3_is_free(const orc3_grid* g, const signed char* mask, double x, double y, double z);
/* Field3D::accumulate, src/Field3D.hpp:40-65 (sequential fp64); returns -1 where the reference throws */
int orc3_accumulate(const orc3_grid* g, double* rho, double charge, double x, double y, double z);
/* the build's fixed-point rule: the eight fp64 weights of Field3D.hpp:56-64 (charge factored out), each rounded to
 * nearest-even at 2^-32, summed as int64 */
int orc3_deposit_fixed(const orc3_grid* g, int n, const double* x, const double* y, const double* z,
                       const unsigned char* alive, int64_t* rho_fixed);
/* Field3D::interpolate, src/Field3D.hpp:67-95 (NaN where the reference throws) */
double orc3_interpolate(const orc3_grid* g, const double* data, double x, double y, double z);
/* Field3D::grad, src/Field3D.hpp:110-163, with xmax = (imax-1)/idx .. which the reference leaves uninitialised;
 * indices are kept inside the array where the reference reads one plane past it with weight zero */
void orc3_grad(const orc3_grid* g, const double* data, double x, double y, double z, double* gx, double* gy, double* gz);
/* Solver::solve right-hand side, src/fields3d.cpp:83-93 (rho scaled in place) */
void orc3_rhs(const orc3_grid* g, const signed char* mask, const double* voltage, double* rho_inout);
/* y = A^T-row operator of Solver::matrix_init, src/fields3d.cpp:39-73: neighbours by FLAT index m +- 1, +- kmax,
 * +- jmax*kmax exactly as the reference stores them (so free nodes on the k = kmax-1 face wrap to the next row) */
void orc3_apply_operator(const orc3_grid* g, const signed char* mask, const double* u, double* y);
/* direct banded solve of that system (stand-in for umfpack_di_solve, src/fields3d.cpp:94) */
int orc3_solve_direct(const orc3_grid* g, const signed char* mask, const double* b, double* u);
/* Species<CARTESIAN3D>::advance, src/species3d.cpp:3-93, collisions off: E = -grad u (ElMag3D::E), Boris with the
 * constant field (Bx,By,Bz) (the reference hard-wires 0), drift, box boundary, is_free, deposit.  rho (fp64) and
 * rho_fixed may be NULL.  alive[k] = 0 marks removed particles.  Returns the number of removals. */
int orc3_advance(const orc3_grid* g, const double* u, const signed char* mask, double charge, double mass, double dt,
                 double Bx, double By, double Bz, int n, double* x, double* y, double* z, double* vx, double* vy,
                 double* vz, unsigned char* alive, double* rho, int64_t* rho_fixed);

#ifdef __cplusplus
}
#endif
#endif
