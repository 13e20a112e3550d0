/* oracle/mag3d_oracle.c — TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's 3-D path.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may use it; the product never does. */
#include "mag3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FREE3 2
#define FIXED3 0
#define IDX(g, i, j, k) (((size_t)(i) * (g)->jmax + (j)) * (g)->kmax + (k))

static int is_dirichlet(signed char m) { return m == FIXED3 || m < 0; }   /* fields3d.cpp:52, :87 */

void orc3_geometry(const orc3_grid* g, signed char* mask, double* voltage)
{
    for (int i = 0; i < g->imax; i++)
        for (int j = 0; j < g->jmax; j++)
            for (int k = 0; k < g->kmax; k++)
            {
                /* fields3d.cpp:28 compares k with z_sampl (not z_sampl-1): that face stays FREE */
                const int fixed = i == 0 || i == g->imax - 1 || j == 0 || j == g->jmax - 1 || k == 0 || k == g->kmax;
                mask[IDX(g, i, j, k)] = fixed ? FIXED3 : FREE3;
                voltage[IDX(g, i, j, k)] = 0.0;
            }
    /* Quadrupole(-1)::set_mask + Electrode::set_voltage (fields3d.hpp:26-36, fields3d.cpp:5-10), voltage 1.0 */
    const size_t c = IDX(g, g->imax / 2, g->jmax / 2, g->kmax / 2);
    mask[c] = -1;
    voltage[c] = 1.0;
}

int orc3_is_free(const orc3_grid* g, const signed char* mask, double x, double y, double z)
{
    int i = (int)(x * g->idx), j = (int)(y * g->idy), k = (int)(z * g->idz);
    if (i < 0 || i > g->imax - 1 || j < 0 || j > g->jmax - 1 || k < 0 || k > g->kmax - 1) return 0;
    /* x == x_max exactly: the reference reads one plane past the mask; the cell is clamped instead (as in 2-D) */
    if (i > g->imax - 2) i = g->imax - 2;
    if (j > g->jmax - 2) j = g->jmax - 2;
    if (k > g->kmax - 2) k = g->kmax - 2;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            for (int c = 0; c < 2; c++)
                if (mask[IDX(g, i + a, j + b, k + c)] == FREE3) return 1;
    return 0;
}

int orc3_accumulate(const orc3_grid* g, double* rho, double charge, double x, double y, double z)
{
    const int i = (int)(x * g->idx), j = (int)(y * g->idy), k = (int)(z * g->idz);
    const double u = x * g->idx - i, v = y * g->idy - j, w = z * g->idz - k;
    /* the reference's own test lets i = imax-1 through and then writes one plane past the array; here that is -1 too */
    if (i < 0 || i > g->imax - 2 || j < 0 || j > g->jmax - 2 || k < 0 || k > g->kmax - 2) return -1;
    rho[IDX(g, i, j, k)] += (1 - u) * (1 - v) * (1 - w) * charge;
    rho[IDX(g, i + 1, j, k)] += u * (1 - v) * (1 - w) * charge;
    rho[IDX(g, i, j + 1, k)] += (1 - u) * v * (1 - w) * charge;
    rho[IDX(g, i + 1, j + 1, k)] += u * v * (1 - w) * charge;
    rho[IDX(g, i, j, k + 1)] += (1 - u) * (1 - v) * w * charge;
    rho[IDX(g, i + 1, j, k + 1)] += u * (1 - v) * w * charge;
    rho[IDX(g, i, j + 1, k + 1)] += (1 - u) * v * w * charge;
    rho[IDX(g, i + 1, j + 1, k + 1)] += u * v * w * charge;
    return 0;
}

static int64_t q32(double w) { return (int64_t)llrint(w * 4294967296.0); }

int orc3_deposit_fixed(const orc3_grid* g, int n, const double* x, const double* y, const double* z,
                       const unsigned char* alive, int64_t* rho)
{
    int bad = 0;
    for (int p = 0; p < n; p++)
    {
        if (alive && !alive[p]) continue;
        const double X = x[p] * g->idx, Y = y[p] * g->idy, Z = z[p] * g->idz;
        int i = (int)X, j = (int)Y, k = (int)Z;
        if (i < 0 || i > g->imax - 1 || j < 0 || j > g->jmax - 1 || k < 0 || k > g->kmax - 1) { bad++; continue; }
        /* x == x_max exactly lands on the last node: clamp the cell, as the 2-D rule does */
        if (i > g->imax - 2) i = g->imax - 2;
        if (j > g->jmax - 2) j = g->jmax - 2;
        if (k > g->kmax - 2) k = g->kmax - 2;
        const double u = X - i, v = Y - j, w = Z - k;
        const double cu = 1.0 - u, cv = 1.0 - v, cw = 1.0 - w;
        /* products grouped as ((a*b)*c), each rounded separately (no FMA: built with -ffp-contract=off) */
        rho[IDX(g, i, j, k)] += q32(cu * cv * cw);
        rho[IDX(g, i + 1, j, k)] += q32(u * cv * cw);
        rho[IDX(g, i, j + 1, k)] += q32(cu * v * cw);
        rho[IDX(g, i + 1, j + 1, k)] += q32(u * v * cw);
        rho[IDX(g, i, j, k + 1)] += q32(cu * cv * w);
        rho[IDX(g, i + 1, j, k + 1)] += q32(u * cv * w);
        rho[IDX(g, i, j + 1, k + 1)] += q32(cu * v * w);
        rho[IDX(g, i + 1, j + 1, k + 1)] += q32(u * v * w);
    }
    return bad;
}

double orc3_interpolate(const orc3_grid* g, const double* d, double x, double y, double z)
{
    const int i = (int)(x * g->idx), j = (int)(y * g->idy), k = (int)(z * g->idz);
    const double u = x * g->idx - i, v = y * g->idy - j, w = z * g->idz - k;
    if (i < 0 || i > g->imax - 2 || j < 0 || j > g->jmax - 2 || k < 0 || k > g->kmax - 2) return NAN;
    double res = 0;
    res += (1 - u) * (1 - v) * (1 - w) * d[IDX(g, i, j, k)];
    res += u * (1 - v) * (1 - w) * d[IDX(g, i + 1, j, k)];
    res += (1 - u) * v * (1 - w) * d[IDX(g, i, j + 1, k)];
    res += u * v * (1 - w) * d[IDX(g, i + 1, j + 1, k)];
    res += (1 - u) * (1 - v) * w * d[IDX(g, i, j, k + 1)];
    res += u * (1 - v) * w * d[IDX(g, i + 1, j, k + 1)];
    res += (1 - u) * v * w * d[IDX(g, i, j + 1, k + 1)];
    res += u * v * w * d[IDX(g, i + 1, j + 1, k + 1)];
    return res;
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   /* mymath.cpp clamp */

/* Field3D::grad_component, src/Field3D.hpp:110-142 */
static double grad_component(const orc3_grid* g, const double* d, double x, double y, double z, int dirx, int diry, int dirz)
{
    const double xmax = (g->imax - 1) / g->idx, ymax = (g->jmax - 1) / g->idy, zmax = (g->kmax - 1) / g->idz;
    x = clampd(x * g->idx, 0.5 * dirx, xmax * g->idx - 0.5 * dirx);
    y = clampd(y * g->idy, 0.5 * diry, ymax * g->idy - 0.5 * diry);
    z = clampd(z * g->idz, 0.5 * dirz, zmax * g->idz - 0.5 * dirz);
    int i = (int)(x + 0.5 * dirx), j = (int)(y + 0.5 * diry), k = (int)(z + 0.5 * dirz);
    /* at the upper clamp the reference reads plane i+1 = imax with weight exactly 0; use plane i-1..i with weight 1 */
    if (i > g->imax - 2) i = g->imax - 2;
    if (j > g->jmax - 2) j = g->jmax - 2;
    if (k > g->kmax - 2) k = g->kmax - 2;
    const double u = x + 0.5 * dirx - i, v = y + 0.5 * diry - j, w = z + 0.5 * dirz - k;
    double gg[8];
    int q = 0;
    for (int c = 0; c < 2; c++)
        for (int b = 0; b < 2; b++)
            for (int a = 0; a < 2; a++)
                gg[q++] = d[IDX(g, i + a, j + b, k + c)] - d[IDX(g, i + a - dirx, j + b - diry, k + c - dirz)];
    const double li = (1 - u) * (1 - v) * (1 - w) * gg[0] + u * (1 - v) * (1 - w) * gg[1] + (1 - u) * v * (1 - w) * gg[2] +
                      u * v * (1 - w) * gg[3] + (1 - u) * (1 - v) * w * gg[4] + u * (1 - v) * w * gg[5] + (1 - u) * v * w * gg[6] +
                      u * v * w * gg[7];
    return li * (g->idx * dirx + g->idy * diry + g->idz * dirz);
}

void orc3_grad(const orc3_grid* g, const double* d, double x, double y, double z, double* gx, double* gy, double* gz)
{
    *gx = grad_component(g, d, x, y, z, 1, 0, 0);
    *gy = grad_component(g, d, x, y, z, 0, 1, 0);
    *gz = grad_component(g, d, x, y, z, 0, 0, 1);
}

void orc3_rhs(const orc3_grid* g, const signed char* mask, const double* voltage, double* rho)
{
    const double eps_0 = 8.854187817e-12;      /* Param::eps_0 */
    const size_t n = (size_t)g->imax * g->jmax * g->kmax;
    for (size_t m = 0; m < n; m++)
        if (is_dirichlet(mask[m])) rho[m] = voltage[m];
        else rho[m] *= -g->macroparticle_factor / eps_0;
}

void orc3_apply_operator(const orc3_grid* g, const signed char* mask, const double* u, double* y)
{
    const long n = (long)g->imax * g->jmax * g->kmax, sk = 1, sj = g->kmax, si = (long)g->jmax * g->kmax;
    for (long m = 0; m < n; m++)
    {
        if (is_dirichlet(mask[m])) { y[m] = u[m]; continue; }
        y[m] = u[m - si] + u[m - sj] + u[m - sk] - 6.0 * u[m] + u[m + sk] + u[m + sj] + u[m + si];
    }
}

int orc3_solve_direct(const orc3_grid* g, const signed char* mask, const double* b, double* u)
{
    /* banded Gaussian elimination without pivoting (identity rows / weakly dominant stencil rows), band = jmax*kmax */
    const long n = (long)g->imax * g->jmax * g->kmax, bw = (long)g->jmax * g->kmax, w = 2 * bw + 1;
    double* a = (double*)calloc((size_t)n * w, sizeof(double));
    double* r = (double*)malloc((size_t)n * sizeof(double));
    if (!a || !r) { free(a); free(r); return 1; }
#define A(row, col) a[(size_t)(row) * w + ((col) - (row) + bw)]
    for (long m = 0; m < n; m++)
    {
        r[m] = b[m];
        if (is_dirichlet(mask[m])) { A(m, m) = 1.0; continue; }
        A(m, m - bw) = 1.0; A(m, m - g->kmax) = 1.0; A(m, m - 1) = 1.0; A(m, m) = -6.0;
        A(m, m + 1) = 1.0; A(m, m + g->kmax) = 1.0; A(m, m + bw) = 1.0;
    }
    for (long p = 0; p < n; p++)
    {
        const double piv = A(p, p);
        const long last = p + bw < n - 1 ? p + bw : n - 1;
        for (long q = p + 1; q <= last; q++)
        {
            const double f = A(q, p);
            if (f == 0.0) continue;
            const double l = f / piv;
            const long cend = p + bw < n - 1 ? p + bw : n - 1;
            for (long c = p; c <= cend; c++) A(q, c) -= l * A(p, c);
            r[q] -= l * r[p];
        }
    }
    for (long p = n - 1; p >= 0; p--)
    {
        double s = r[p];
        const long cend = p + bw < n - 1 ? p + bw : n - 1;
        for (long c = p + 1; c <= cend; c++) s -= A(p, c) * u[c];
        u[p] = s / A(p, p);
    }
#undef A
    free(a);
    free(r);
    return 0;
}

static double mod_ref(double x, double y)      /* mymath.cpp:71-76 */
{
    if (x >= 0.0 && x <= y) return x;
    return x - y * (int)(x / y) + (x < 0 ? y : 0);
}

int orc3_advance(const orc3_grid* g, const double* u, const signed char* mask, double charge, double mass, double dt,
                 double Bx, double By, double Bz, int n, double* x, double* y, double* z, double* vx, double* vy,
                 double* vz, unsigned char* alive, double* rho, int64_t* rho_fixed)
{
    const double qmdt = charge / mass * dt;
    int removed = 0;
    for (int p = 0; p < n; p++)
    {
        if (!alive[p]) continue;
        double Ex, Ey, Ez;
        orc3_grad(g, u, x[p], y[p], z[p], &Ex, &Ey, &Ez);
        Ex *= -1.0; Ey *= -1.0; Ez *= -1.0;
        vx[p] += Ex * qmdt / 2.0;
        vy[p] += Ey * qmdt / 2.0;
        vz[p] += Ez * qmdt / 2.0;
        double tmp = charge * dt / (2.0 * mass);
        const double tx = Bx * tmp, ty = By * tmp, tz = Bz * tmp;
        const double px = vx[p] - vy[p] * tz + vz[p] * ty;
        const double py = vy[p] - vz[p] * tx + vx[p] * tz;
        const double pz = vz[p] - vx[p] * ty + vy[p] * tx;
        tmp = 2.0 / (1 + tx * tx + ty * ty + tz * tz);
        const double sx = tx * tmp, sy = ty * tmp, sz = tz * tmp;
        const double ox = vx[p], oy = vy[p], oz = vz[p];
        vx[p] = ox - py * sz + pz * sy;
        vy[p] = oy - pz * sx + px * sz;
        vz[p] = oz - px * sy + py * sx;
        vx[p] += Ex * qmdt / 2.0;
        vy[p] += Ey * qmdt / 2.0;
        vz[p] += Ez * qmdt / 2.0;
        x[p] += vx[p] * dt;
        y[p] += vy[p] * dt;
        z[p] += vz[p] * dt;
        if (x[p] > g->x_max || x[p] < 0 || y[p] > g->y_max || y[p] < 0 || z[p] > g->z_max || z[p] < 0)
        {
            if (g->boundary == 0) { alive[p] = 0; removed++; continue; }
            x[p] = mod_ref(x[p], g->x_max);
            y[p] = mod_ref(y[p], g->y_max);
            z[p] = mod_ref(z[p], g->z_max);
        }
        if (!orc3_is_free(g, mask, x[p], y[p], z[p])) { alive[p] = 0; removed++; continue; }
        if (rho) orc3_accumulate(g, rho, charge, x[p], y[p], z[p]);
        if (rho_fixed) orc3_deposit_fixed(g, 1, &x[p], &y[p], &z[p], NULL, rho_fixed);
    }
    return removed;
}
