// poisson.cu — GPU Poisson solve: Galerkin geometric multigrid on the reference's operator.
//
// Replaces Fields::Fields' CSC build + UMFPACK LU (reference src/fields.cpp:133-275) and the
// per-step triangular solves of Fields::boundary_solve / boundary_solve_rf (src/fields.cpp:278-348).
//
// The operator is the reference's: identity rows on FIXED / FIXED_RF nodes, the unit 5-point stencil
// [1,1,-4,1,1] elsewhere in Cartesian coordinates (fields.cpp:229-235), the r-z stencil with
// k1=(i-1/2)/(dx^2 i), k2=1/dz^2, k3=(i+1/2)/(dx^2 i) and the axis row k3=4/dx^2 in cylindrical
// coordinates (fields.cpp:156-208).  Cylindrical rows are multiplied by s_i = i (1/8 on the axis),
// which makes the matrix symmetric without changing the solution.
//
// Hierarchy: vertex-centred coarsening by 2 (semi-coarsening while one direction couples >3x more
// strongly), bilinear interpolation P that skips Dirichlet nodes, restriction R = P^T and Galerkin
// coarse operators A_c = R A P, which stay 9-point.  They are built once per geometry on the host by
// probing R A P with the nine (I mod 3, J mod 3) colourings.  Electrodes of any shape and odd
// interval counts (x_sampl = 512 or 200) are handled by the Galerkin product itself; measured
// convergence is 0.1-0.2 per V(2,2) cycle on every reference geometry (scratch/mg_galerkin_proto.py).
// Smoother: 4-colour Gauss-Seidel (exact for 9-point stencils, deterministic).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace {

// ---------------------------------------------------------------- host side: hierarchy construction
struct HostLevel
{
    int M, N, fx, fz;
    std::vector<double> coef;            // [9][M*N]
    std::vector<unsigned char> freem;
};

inline int cidx(int di, int dj) { return (di + 1) * 3 + (dj + 1); }

void host_apply(const HostLevel& L, const std::vector<double>& x, std::vector<double>& y)
{
    const int M = L.M, N = L.N;
    const size_t n = (size_t)M * N;
    y.assign(n, 0.0);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++)
        {
            const size_t k = (size_t)i * N + j;
            if (!L.freem[k]) continue;
            double s = 0;
            for (int di = -1; di <= 1; di++)
                for (int dj = -1; dj <= 1; dj++)
                {
                    const int ii = i + di, jj = j + dj;
                    if (ii < 0 || ii >= M || jj < 0 || jj >= N) continue;
                    const double a = L.coef[(size_t)cidx(di, dj) * n + k];
                    if (a != 0.0) s += a * x[(size_t)ii * N + jj];
                }
            y[k] = s;
        }
}

// ef = P ec (zero on non-free fine nodes)
void host_prolong(const HostLevel& F, int Mc, int Nc, const std::vector<double>& ec, std::vector<double>& ef)
{
    const int M = F.M, N = F.N;
    ef.assign((size_t)M * N, 0.0);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++)
        {
            const size_t k = (size_t)i * N + j;
            if (!F.freem[k]) continue;
            const int I = i / F.fx, J = j / F.fz;
            const double wi = (F.fx == 2 && (i & 1)) ? 0.5 : 0.0, wj = (F.fz == 2 && (j & 1)) ? 0.5 : 0.0;
            double s = 0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                {
                    const double w = (a ? wi : 1.0 - wi) * (b ? wj : 1.0 - wj);
                    if (w == 0.0) continue;
                    const int II = I + a, JJ = J + b;
                    if (II >= Mc || JJ >= Nc) continue;
                    s += w * ec[(size_t)II * Nc + JJ];
                }
            ef[k] = s;
        }
}

// rc = P^T rf
void host_restrict(const HostLevel& F, int Mc, int Nc, const std::vector<double>& rf, std::vector<double>& rc)
{
    const int M = F.M, N = F.N;
    rc.assign((size_t)Mc * Nc, 0.0);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++)
        {
            const size_t k = (size_t)i * N + j;
            if (!F.freem[k] || rf[k] == 0.0) continue;
            const int I = i / F.fx, J = j / F.fz;
            const double wi = (F.fx == 2 && (i & 1)) ? 0.5 : 0.0, wj = (F.fz == 2 && (j & 1)) ? 0.5 : 0.0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++)
                {
                    const double w = (a ? wi : 1.0 - wi) * (b ? wj : 1.0 - wj);
                    if (w == 0.0) continue;
                    const int II = I + a, JJ = J + b;
                    if (II >= Mc || JJ >= Nc) continue;
                    rc[(size_t)II * Nc + JJ] += w * rf[k];
                }
        }
}

void build_fine_level(const mag2d_ctx* c, HostLevel& L, std::vector<double>& rowscale)
{
    const mag2d_grid_desc& g = c->g;
    const int M = g.M, N = g.N;
    const size_t n = (size_t)M * N;
    L.M = M;
    L.N = N;
    L.fx = L.fz = 1;
    L.coef.assign(9 * n, 0.0);
    L.freem.assign(n, 0);
    rowscale.assign(M, 1.0);
    const bool cyl = g.coord == MAG2D_CYLINDRICAL;
    const double dx = g.dx, dz = g.dz;
    for (int i = 0; i < M; i++)
    {
        if (cyl) rowscale[i] = i == 0 ? 0.125 : (double)i;
        for (int j = 0; j < N; j++)
        {
            const size_t k = (size_t)i * N + j;
            const unsigned char m = c->h_mask[k];
            if (m == MAG2D_FIXED || m == MAG2D_FIXED_RF)
            {
                L.coef[(size_t)cidx(0, 0) * n + k] = 1.0;
                continue;
            }
            L.freem[k] = 1;
            double W, E, S, Nn, C;
            if (cyl)
            {
                const double k2 = 1.0 / (dz * dz);
                if (i == 0)
                {
                    const double k3 = 1.0 / (dx * dx * 0.25);
                    W = 0.0; E = k3; S = k2; Nn = k2; C = -2.0 * k2 - k3;
                }
                else
                {
                    const double k1 = (i - 0.5) / (dx * dx * i), k3 = (i + 0.5) / (dx * dx * i);
                    W = k1; E = k3; S = k2; Nn = k2; C = -2.0 * k2 - k1 - k3;
                }
                const double s = rowscale[i];
                W *= s; E *= s; S *= s; Nn *= s; C *= s;
            }
            else
            {
                W = E = S = Nn = 1.0;
                C = -4.0;
            }
            if (i > 0) L.coef[(size_t)cidx(-1, 0) * n + k] = W;
            if (i < M - 1) L.coef[(size_t)cidx(1, 0) * n + k] = E;
            if (j > 0) L.coef[(size_t)cidx(0, -1) * n + k] = S;
            if (j < N - 1) L.coef[(size_t)cidx(0, 1) * n + k] = Nn;
            L.coef[(size_t)cidx(0, 0) * n + k] = C;
        }
    }
}

void build_coarse_level(const HostLevel& F, HostLevel& Cl)
{
    const int Mc = (F.M - 1) / F.fx + 1, Nc = (F.N - 1) / F.fz + 1;
    const size_t nc = (size_t)Mc * Nc;
    Cl.M = Mc;
    Cl.N = Nc;
    Cl.fx = Cl.fz = 1;
    Cl.coef.assign(9 * nc, 0.0);
    Cl.freem.assign(nc, 0);
    std::vector<double> ec(nc), ef, af, y;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
        {
            std::fill(ec.begin(), ec.end(), 0.0);
            for (int I = a; I < Mc; I += 3)
                for (int J = b; J < Nc; J += 3) ec[(size_t)I * Nc + J] = 1.0;
            host_prolong(F, Mc, Nc, ec, ef);
            host_apply(F, ef, af);
            host_restrict(F, Mc, Nc, af, y);
            for (int I = 0; I < Mc; I++)
                for (int J = 0; J < Nc; J++)
                    for (int di = -1; di <= 1; di++)
                        for (int dj = -1; dj <= 1; dj++)
                        {
                            const int II = I + di, JJ = J + dj;
                            if (II < 0 || II >= Mc || JJ < 0 || JJ >= Nc) continue;
                            if (II % 3 != a || JJ % 3 != b) continue;
                            Cl.coef[(size_t)cidx(di, dj) * nc + (size_t)I * Nc + J] = y[(size_t)I * Nc + J];
                        }
        }
    for (size_t k = 0; k < nc; k++)
    {
        const double d = Cl.coef[(size_t)cidx(0, 0) * nc + k];
        if (std::fabs(d) > 1e-300) Cl.freem[k] = 1;
        else
        {
            for (int q = 0; q < 9; q++) Cl.coef[(size_t)q * nc + k] = 0.0;
            Cl.coef[(size_t)cidx(0, 0) * nc + k] = 1.0;
        }
    }
}

// ---------------------------------------------------------------------------------- device kernels
struct LevelDev
{
    int M, N;
    const double* __restrict__ coef;         // never written on the device
    const unsigned char* __restrict__ freem;
    double* u;     // u and b are written and re-read inside k_mg_coarse: no __restrict__, no read-only path
    double* b;
};

// (A u)(i,j) excluding the diagonal term; returns the diagonal through diag
__device__ __forceinline__ double offdiag_sum(const LevelDev& L, int i, int j, double& diag)
{
    const int M = L.M, N = L.N;
    const size_t n = (size_t)M * N, k = (size_t)i * N + j;
    double s = 0.0;
#pragma unroll
    for (int di = -1; di <= 1; di++)
#pragma unroll
        for (int dj = -1; dj <= 1; dj++)
        {
            const int q = (di + 1) * 3 + (dj + 1);
            const double a = __ldg(L.coef + (size_t)q * n + k);
            if (di == 0 && dj == 0) { diag = a; continue; }
            const int ii = i + di, jj = j + dj;
            if (a != 0.0 && ii >= 0 && ii < M && jj >= 0 && jj < N) s += a * L.u[(size_t)ii * N + jj];
        }
    return s;
}

__device__ __forceinline__ double residual_at(const LevelDev& L, int i, int j)
{
    const size_t k = (size_t)i * L.N + j;
    if (!L.freem[k]) return 0.0;
    double diag;
    const double s = offdiag_sum(L, i, j, diag);
    return L.b[k] - s - diag * L.u[k];
}

// coarse rhs = P^T (b - A u) of the fine level; coarse error initialised to zero
__global__ void k_mg_restrict(const __grid_constant__ LevelDev F, int fx, int fz, int Mc, int Nc, double* __restrict__ bc,
                              double* __restrict__ ec)
{
    const int J = blockIdx.x * blockDim.x + threadIdx.x;
    const int I = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= Mc || J >= Nc) return;
    const int i0 = I * fx, j0 = J * fz;
    double s = 0.0;
    const int ri = fx == 2 ? 1 : 0, rj = fz == 2 ? 1 : 0;
    for (int di = -ri; di <= ri; di++)
        for (int dj = -rj; dj <= rj; dj++)
        {
            const int i = i0 + di, j = j0 + dj;
            if (i < 0 || i >= F.M || j < 0 || j >= F.N) continue;
            const double w = (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0);
            s += w * residual_at(F, i, j);
        }
    bc[(size_t)I * Nc + J] = s;
    ec[(size_t)I * Nc + J] = 0.0;
}

// fine u += P ec on free nodes
__global__ void k_mg_prolong(const __grid_constant__ LevelDev F, int fx, int fz, int Mc, int Nc, const double* __restrict__ ec)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.M || j >= F.N) return;
    const size_t k = (size_t)i * F.N + j;
    if (!F.freem[k]) return;
    const int I = i / fx, J = j / fz;
    const double wi = (fx == 2 && (i & 1)) ? 0.5 : 0.0, wj = (fz == 2 && (j & 1)) ? 0.5 : 0.0;
    double s = (1.0 - wi) * (1.0 - wj) * ec[(size_t)I * Nc + J];
    if (wi != 0.0 && I + 1 < Mc) s += wi * (1.0 - wj) * ec[(size_t)(I + 1) * Nc + J];
    if (wj != 0.0 && J + 1 < Nc) s += (1.0 - wi) * wj * ec[(size_t)I * Nc + J + 1];
    if (wi != 0.0 && wj != 0.0 && I + 1 < Mc && J + 1 < Nc) s += wi * wj * ec[(size_t)(I + 1) * Nc + J + 1];
    F.u[k] += s;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// Convergence measure: max_k |r_k / a_kk| (the size of the Jacobi update, in volts; row scaling cancels)
// and max_k |u_k|.  Rows of the cylindrical operator carry coefficients ~1e9 and the Dirichlet rows
// coefficients of 1, so a plain residual norm against max|b| would mix units.
__global__ void k_mg_residual_norm(const __grid_constant__ LevelDev L, double* __restrict__ out2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    double r = 0.0, um = 0.0;
    if (i < L.M && j < L.N)
    {
        const size_t k = (size_t)i * L.N + j;
        um = fabs(L.u[k]);
        if (L.freem[k])
        {
            double diag;
            const double s = offdiag_sum(L, i, j, diag);
            r = fabs((L.b[k] - s - diag * L.u[k]) / diag);
        }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        r = fmax(r, __shfl_xor_sync(MAG2D_FULL_MASK, r, o));
        um = fmax(um, __shfl_xor_sync(MAG2D_FULL_MASK, um, o));
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0)
    {
        atomic_max_nonneg(out2, r);
        atomic_max_nonneg(out2 + 1, um);
    }
}

// rho (all species, fixed point) -> right-hand side: Fields::boundary_solve / _rf, fields.cpp:278-346.
// Writes b in the reference's scaling (b_ref), the row-scaled copy used by the hierarchy (b_mg) and
// the Dirichlet values into u.
struct RhsArgs
{
    int M, N, coord, rf, n_species;
    double dx, dz, dV, mpf;
    const unsigned char* mask;
    const double* voltage;
    const unsigned long long* rho;   // [n_species][M*N]
    const double* charges;           // [n_species]
    const double* rowscale;
    double* b_ref;
    double* b_mg;
    double* u;
    // direct solver (poisson_direct.cu): interior columns with the Dirichlet end nodes moved to the right-hand side
    double* bp;                      // FOLDED: [i][c] = v_c + v_(n-1-c), [i][hp + c] = v_c - v_(n-1-c), c = j - 1 the interior column
    int ld, hp;
    const unsigned char* rowfree;
    const double* k2row;
};

__device__ __forceinline__ double dirichlet_value(unsigned char m, double voltage, int rf)
{
    if (m == MAG2D_FIXED) return rf ? 0.0 : voltage;
    return rf ? voltage : 0.0;      // MAG2D_FIXED_RF
}

__device__ __forceinline__ double rho_coulomb(const unsigned long long* rho, const double* charges, int ns, size_t n, size_t k)
{
    double q = 0.0;
    for (int s = 0; s < ns; s++)
    {
        const double c = charges[s];
        if (c != 0.0) q += c * ((double)(long long)rho[(size_t)s * n + k] * 2.3283064365386963e-10);
    }
    return q;
}

// right-hand side of node (i, j); returns the value the direct solver sees in interior column j - 1 (Dirichlet end nodes moved over)
__device__ __forceinline__ double rhs_node(const RhsArgs& A, int i, int j)
{
    const size_t n = (size_t)A.M * A.N, k = (size_t)i * A.N + j;
    const unsigned char m = A.mask[k];
    double b;
    if (m == MAG2D_FIXED) b = A.rf ? 0.0 : A.voltage[k];
    else if (m == MAG2D_FIXED_RF) b = A.rf ? A.voltage[k] : 0.0;
    else
    {
        const double rho = rho_coulomb(A.rho, A.charges, A.n_species, n, k);
        if (A.coord == MAG2D_CYLINDRICAL)
        {
            if (i > 0) b = rho * (-1.0 / MAG2D_EPS0 / (M_PI * A.dx * A.dx * 2.0 * i * A.dz) * A.mpf);
            else b = rho * (-1.0 / MAG2D_EPS0 / (M_PI * A.dx * A.dx * 0.25 * A.dz) * A.mpf);
        }
        else
            b = rho * (-(A.dx * A.dx) / MAG2D_EPS0 / A.dV);
    }
    A.b_ref[k] = b;
    double v = b;
    if (A.bp && j >= 1 && j <= A.N - 2 && A.rowfree[i])
    {
        if (j == 1) v -= A.k2row[i] * dirichlet_value(A.mask[k - 1], A.voltage[k - 1], A.rf);
        if (j == A.N - 2) v -= A.k2row[i] * dirichlet_value(A.mask[k + 1], A.voltage[k + 1], A.rf);
    }
    if (m == MAG2D_FIXED || m == MAG2D_FIXED_RF)
    {
        A.b_mg[k] = b;
        A.u[k] = b;
    }
    else
        A.b_mg[k] = b * A.rowscale[i];
    return v;
}

// one thread per pair of nodes (i, j), (i, N-1-j): mirror images of each other among the interior columns, so the direct
// solver's right-hand side leaves the kernel folded (poisson_direct.cu)
__global__ void k_rhs(const __grid_constant__ RhsArgs A)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int jm = A.N - 1 - j;
    if (i >= A.M || j > jm) return;
    const double v = rhs_node(A, i, j);
    const double vm = j < jm ? rhs_node(A, i, jm) : 0.0;
    if (A.bp && j >= 1)
    {
        double* row = A.bp + (size_t)i * A.ld;
        row[j - 1] = v + vm;                                 // the middle column of an odd count is its own mirror image
        row[A.hp + j - 1] = j < jm ? v - vm : 0.0;
    }
}

__global__ void k_rho_total(const unsigned long long* __restrict__ rho, const double* __restrict__ charges, int ns, size_t n,
                            double* __restrict__ out)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = rho_coulomb(rho, charges, ns, n, k);
}

// Fields::u_smooth, fields.cpp:28-113
__global__ void k_symmetrize(double* __restrict__ u, int M)
{
    // one thread per orbit (i <= ic/2, j <= i) of the 8-fold symmetry group; orbits are disjoint
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    const int ic = M - 1;
    if (i > ic / 2 || j > i) return;
    const int N = M;
#define U(a, b) u[(size_t)(a) * N + (b)]
    double sum = 0;
    sum += U(i, j);
    sum += U(ic - i, j);
    sum += U(i, ic - j);
    sum += U(ic - i, ic - j);
    sum += U(ic - j, ic - i);
    sum += U(j, i);
    sum += U(ic - j, i);
    sum += U(j, ic - i);
    sum /= 8.0;
    U(i, j) = sum;
    U(ic - i, j) = sum;
    U(i, ic - j) = sum;
    U(ic - i, ic - j) = sum;
    U(ic - j, ic - i) = sum;
    U(j, i) = sum;
    U(ic - j, i) = sum;
    U(j, ic - i) = sum;
#undef U
}

__global__ void k_smooth9(const double* __restrict__ t, double* __restrict__ u, int M, int N, double radius2)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int i = blockIdx.y * blockDim.y + threadIdx.y + 1;
    // the reference's inner loop runs to j <= lmax-1 and so reads one element past each row end, i.e.
    // the first element of the next row (flat storage); reproduced, with 0 past the end of the array
    if (i >= M - 1 || j > N - 1) return;
    const double icf = (M - 1) / 2.0, jcf = (N - 1) / 2.0;
    const double r = (i - icf) * (i - icf) + (j - jcf) * (j - jcf);
    if (radius2 > 0 && r > radius2) return;
    const long long n = (long long)M * N;
    auto T = [&](int a, int b) -> double {
        const long long k = (long long)a * N + b;
        return k < n ? t[k] : 0.0;
    };
    const double sum = T(i, j) + T(i - 1, j) * 0.5 + T(i + 1, j) * 0.5 + T(i, j - 1) * 0.5 + T(i, j + 1) * 0.5 +
                       T(i - 1, j - 1) * 0.25 + T(i + 1, j - 1) * 0.25 + T(i - 1, j + 1) * 0.25 + T(i + 1, j + 1) * 0.25;
    u[(size_t)i * N + j] = sum / 4.0;
}

// initial guess for the next solve by linear extrapolation in time: u <- 2u - u_prev on the unknowns
__global__ void k_extrapolate(double* __restrict__ u, double* __restrict__ u_prev, const unsigned char* __restrict__ freem, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double cur = u[k];
    if (freem[k]) u[k] = 2.0 * cur - u_prev[k];
    u_prev[k] = cur;
}

__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[k];
}

LevelDev level_view(const MgLevel& L)
{
    LevelDev v;
    v.M = L.M;
    v.N = L.N;
    v.coef = L.coef;
    v.freem = L.freem;
    v.u = L.u;
    v.b = L.b;
    return v;
}

inline dim3 grid2d(int nj, int ni, dim3 block) { return dim3((nj + block.x - 1) / block.x, (ni + block.y - 1) / block.y); }

// ---- tile smoother: SWEEPS full 4-colour Gauss-Seidel sweeps in ONE launch ---------------------------
// Each CTA owns a TILE x TILE block of nodes and loads it with a halo of HALO = 4*SWEEPS nodes into shared
// memory.  Colour step s may update every node at distance >= s+1 from the edge of the loaded block (its
// neighbours are still exact there), so after 4*SWEEPS colour steps the owned tile holds exactly the values
// the global 4-colour sweep would have produced: same iteration, same bits, 4*SWEEPS times fewer launches,
// at the price of recomputing the halo ((TILE+2*HALO)^2 / TILE^2 = 2.25x arithmetic for SWEEPS = 2).
constexpr int MG_TILE = 32;
template <int SWEEPS>
struct MgTileCfg
{
    static constexpr int HALO = 4 * SWEEPS, W = MG_TILE + 2 * HALO, NODES = W * W;
    // shared memory: u, b and the nine coefficient planes of the loaded block, plus the free flags
    static constexpr size_t SMEM = sizeof(double) * 11 * NODES + NODES;
};

template <int SWEEPS>
__global__ void __launch_bounds__(256, 1) k_mg_smooth_tile(const __grid_constant__ LevelDev L)
{
    using Cfg = MgTileCfg<SWEEPS>;
    constexpr int HALO = Cfg::HALO, W = Cfg::W, NODES = Cfg::NODES;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* su = reinterpret_cast<double*>(smem_raw);      // [NODES]
    double* sb = su + NODES;                               // [NODES]
    double* sc = sb + NODES;                               // [9][NODES]
    unsigned char* sf = reinterpret_cast<unsigned char*>(sc + 9 * NODES);
    const int M = L.M, N = L.N;
    const size_t n = (size_t)M * N;
    const int i0 = blockIdx.y * MG_TILE - HALO, j0 = blockIdx.x * MG_TILE - HALO;
    // stage the whole block once (coalesced along j); everything after this runs out of shared memory
    for (int q = threadIdx.x; q < NODES; q += 256)
    {
        const int li = q / W, lj = q - li * W;
        const int gi = i0 + li, gj = j0 + lj;
        const bool in = gi >= 0 && gi < M && gj >= 0 && gj < N;
        const size_t k = in ? (size_t)gi * N + gj : 0;
        su[q] = in ? L.u[k] : 0.0;
        sb[q] = in ? L.b[k] : 0.0;
        sf[q] = in ? L.freem[k] : 0;
#pragma unroll
        for (int c = 0; c < 9; c++) sc[c * NODES + q] = in ? __ldg(L.coef + (size_t)c * n + k) : 0.0;
    }
    __syncthreads();
    for (int step = 0; step < 4 * SWEEPS; step++)
    {
        const int colour = step & 3;
        const int ci = colour >> 1, cj = colour & 1;
        // nodes of this colour inside the still-exact region [lo, hi]^2, enumerated without gaps
        const int lo = step + 1, hi = W - 2 - step;
        const int fi = lo + (((i0 + lo) & 1) != ci), fj = lo + (((j0 + lo) & 1) != cj);   // first local row/col of the colour
        const int ni = fi > hi ? 0 : (hi - fi) / 2 + 1, nj = fj > hi ? 0 : (hi - fj) / 2 + 1;
        for (int q = threadIdx.x; q < ni * nj; q += 256)
        {
            const int li = fi + 2 * (q / nj), lj = fj + 2 * (q % nj);
            const int p = li * W + lj;
            if (!sf[p]) continue;
            const double acc = sc[0 * NODES + p] * su[p - W - 1] + sc[1 * NODES + p] * su[p - W] + sc[2 * NODES + p] * su[p - W + 1] +
                               sc[3 * NODES + p] * su[p - 1] + sc[5 * NODES + p] * su[p + 1] + sc[6 * NODES + p] * su[p + W - 1] +
                               sc[7 * NODES + p] * su[p + W] + sc[8 * NODES + p] * su[p + W + 1];
            su[p] = (sb[p] - acc) / sc[4 * NODES + p];
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < MG_TILE * MG_TILE; q += 256)
    {
        const int li = HALO + q / MG_TILE, lj = HALO + q % MG_TILE;
        const int gi = i0 + li, gj = j0 + lj;
        const int p = li * W + lj;
        if (gi < M && gj < N && sf[p]) L.u[(size_t)gi * N + gj] = su[p];
    }
}

// ---- the coarse part of the V-cycle in ONE single-CTA launch -----------------------------------------
// All levels with at most MG_COARSE_NODES nodes are visited by one thread block that walks down
// (pre-smooth, restrict), solves the coarsest level with many sweeps and walks back up (prolong,
// post-smooth) with __syncthreads() between colour steps.  The arrays stay in L1/L2.
constexpr int MG_COARSE_NODES = 4225;     // 65 x 65
constexpr int MG_MAX_LEVELS = 14;
struct CoarseArgs
{
    LevelDev L[MG_MAX_LEVELS];
    int fx[MG_MAX_LEVELS], fz[MG_MAX_LEVELS];
    int first, count;      // levels [first, first+count)
    int nu1, nu2, nu_coarsest;
    const double* inv;     // dense inverse of the coarsest operator (row-major n x n) or nullptr
};

__device__ __forceinline__ void cta_smooth(const LevelDev& L, int sweeps)
{
    const int M = L.M, N = L.N;
    const int half_n = (N + 1) >> 1, half_m = (M + 1) >> 1;
    for (int s = 0; s < sweeps; s++)
        for (int colour = 0; colour < 4; colour++)
        {
            const int ci = colour >> 1, cj = colour & 1;
            for (int q = threadIdx.x; q < half_m * half_n; q += blockDim.x)
            {
                const int i = 2 * (q / half_n) + ci, j = 2 * (q % half_n) + cj;
                if (i >= M || j >= N) continue;
                const size_t k = (size_t)i * N + j;
                if (!L.freem[k]) continue;
                double diag;
                const double acc = offdiag_sum(L, i, j, diag);
                L.u[k] = (L.b[k] - acc) / diag;
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(1024) k_mg_coarse(const __grid_constant__ CoarseArgs A)
{
    const int last = A.first + A.count - 1;
    for (int l = A.first; l < last; l++)
    {
        const LevelDev& F = A.L[l];
        const LevelDev& Cl = A.L[l + 1];
        cta_smooth(F, A.nu1);
        const int fx = A.fx[l], fz = A.fz[l];
        const int ri = fx == 2 ? 1 : 0, rj = fz == 2 ? 1 : 0;
        for (int q = threadIdx.x; q < Cl.M * Cl.N; q += blockDim.x)
        {
            const int I = q / Cl.N, J = q % Cl.N;
            double s = 0.0;
            for (int di = -ri; di <= ri; di++)
                for (int dj = -rj; dj <= rj; dj++)
                {
                    const int i = I * fx + di, j = J * fz + dj;
                    if (i < 0 || i >= F.M || j < 0 || j >= F.N) continue;
                    s += (di ? 0.5 : 1.0) * (dj ? 0.5 : 1.0) * residual_at(F, i, j);
                }
            Cl.b[q] = s;
            Cl.u[q] = 0.0;
        }
        __syncthreads();
    }
    if (A.inv)
    {
        // exact coarsest solve: e = A^-1 b with the host-computed dense inverse (n <= 256 unknowns)
        const LevelDev& Cz = A.L[last];
        const int nz = Cz.M * Cz.N;
        for (int q = threadIdx.x; q < nz; q += blockDim.x)
        {
            const double* row = A.inv + (size_t)q * nz;
            double acc = 0.0;
            for (int r = 0; r < nz; r++) acc += row[r] * Cz.b[r];
            Cz.u[q] = acc;
        }
        __syncthreads();
    }
    else
        cta_smooth(A.L[last], A.nu_coarsest);
    for (int l = last - 1; l >= A.first; l--)
    {
        const LevelDev& F = A.L[l];
        const LevelDev& Cl = A.L[l + 1];
        const int fx = A.fx[l], fz = A.fz[l];
        for (int q = threadIdx.x; q < F.M * F.N; q += blockDim.x)
        {
            if (!F.freem[q]) continue;
            const int i = q / F.N, j = q % F.N;
            const int I = i / fx, J = j / fz;
            const double wi = (fx == 2 && (i & 1)) ? 0.5 : 0.0, wj = (fz == 2 && (j & 1)) ? 0.5 : 0.0;
            double s = (1.0 - wi) * (1.0 - wj) * Cl.u[(size_t)I * Cl.N + J];
            if (wi != 0.0 && I + 1 < Cl.M) s += wi * (1.0 - wj) * Cl.u[(size_t)(I + 1) * Cl.N + J];
            if (wj != 0.0 && J + 1 < Cl.N) s += (1.0 - wi) * wj * Cl.u[(size_t)I * Cl.N + J + 1];
            if (wi != 0.0 && wj != 0.0 && I + 1 < Cl.M && J + 1 < Cl.N) s += wi * wj * Cl.u[(size_t)(I + 1) * Cl.N + J + 1];
            F.u[q] += s;
        }
        __syncthreads();
        cta_smooth(F, A.nu2);
    }
}

int smooth_level(mag2d_ctx* c, const MgLevel& L)
{
    static bool attr_set = false;
    if (!attr_set)
    {
        CUDA_OK(cudaFuncSetAttribute(k_mg_smooth_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MgTileCfg<2>::SMEM));
        attr_set = true;
    }
    const dim3 grid((L.N + MG_TILE - 1) / MG_TILE, (L.M + MG_TILE - 1) / MG_TILE);
    k_mg_smooth_tile<2><<<grid, 256, MgTileCfg<2>::SMEM, c->stream>>>(level_view(L));
    c->launches++;
    return 0;
}

int coarse_vcycle(mag2d_ctx* c, size_t first)
{
    CoarseArgs A;
    memset(&A, 0, sizeof(A));
    A.first = (int)first;
    A.count = (int)(c->mg.size() - first);
    for (size_t l = first; l < c->mg.size(); l++)
    {
        A.L[l] = level_view(c->mg[l]);
        A.fx[l] = c->mg[l].fx;
        A.fz[l] = c->mg[l].fz;
    }
    A.nu1 = A.nu2 = 2;
    A.nu_coarsest = 30;
    A.inv = c->d_mg_inv;
    k_mg_coarse<<<1, 1024, 0, c->stream>>>(A);
    c->launches++;
    return 0;
}

int vcycle_level(mag2d_ctx* c, size_t l)
{
    const MgLevel& L = c->mg[l];
    if ((size_t)L.M * L.N <= (size_t)MG_COARSE_NODES || l + 1 == c->mg.size()) return coarse_vcycle(c, l);
    const MgLevel& Cl = c->mg[l + 1];
    smooth_level(c, L);
    const dim3 block(32, 8);
    k_mg_restrict<<<grid2d(Cl.N, Cl.M, block), block, 0, c->stream>>>(level_view(L), L.fx, L.fz, Cl.M, Cl.N, Cl.b, Cl.u);
    c->launches++;
    if (vcycle_level(c, l + 1)) return 1;
    k_mg_prolong<<<grid2d(L.N, L.M, block), block, 0, c->stream>>>(level_view(L), L.fx, L.fz, Cl.M, Cl.N, Cl.u);
    c->launches++;
    smooth_level(c, L);
    return 0;
}

}  // namespace

void mg_free(mag2d_ctx* c)
{
    for (size_t l = 0; l < c->mg.size(); l++)
    {
        MgLevel& L = c->mg[l];
        if (L.coef) cudaFree(L.coef);
        if (L.freem) cudaFree(L.freem);
        if (l > 0 && L.u) cudaFree(L.u);
        if (L.b) cudaFree(L.b);
    }
    c->mg.clear();
    if (c->d_rowscale) { cudaFree(c->d_rowscale); c->d_rowscale = nullptr; }
    if (c->d_mg_inv) { cudaFree(c->d_mg_inv); c->d_mg_inv = nullptr; }
}

int mg_setup(mag2d_ctx* c)
{
    mg_free(c);
    const mag2d_grid_desc& g = c->g;
    std::vector<HostLevel> H(1);
    std::vector<double> rowscale;
    build_fine_level(c, H[0], rowscale);
    const bool cyl = g.coord == MAG2D_CYLINDRICAL;
    double hx = cyl ? g.dx : 1.0, hz = cyl ? g.dz : 1.0;
    const int min_size = 5;
    while (H.size() < 14)
    {
        HostLevel& F = H.back();
        const double ax = 1.0 / (hx * hx), az = 1.0 / (hz * hz);
        const bool cx = (F.M - 1) / 2 + 1 >= min_size && ax >= 0.3 * az;
        const bool cz = (F.N - 1) / 2 + 1 >= min_size && az >= 0.3 * ax;
        if (!cx && !cz) break;
        F.fx = cx ? 2 : 1;
        F.fz = cz ? 2 : 1;
        HostLevel Cl;
        build_coarse_level(F, Cl);
        hx *= F.fx;
        hz *= F.fz;
        H.push_back(std::move(Cl));
    }
    c->mg.resize(H.size());
    for (size_t l = 0; l < H.size(); l++)
    {
        MgLevel& L = c->mg[l];
        const HostLevel& h = H[l];
        const size_t n = (size_t)h.M * h.N;
        L.M = h.M;
        L.N = h.N;
        L.fx = h.fx;
        L.fz = h.fz;
        CUDA_OK(cudaMalloc(&L.coef, sizeof(double) * 9 * n));
        CUDA_OK(cudaMalloc(&L.freem, n));
        CUDA_OK(cudaMalloc(&L.b, sizeof(double) * n));
        if (l == 0) L.u = c->d_u;
        else CUDA_OK(cudaMalloc(&L.u, sizeof(double) * n));
        CUDA_OK(cudaMemcpyAsync(L.coef, h.coef.data(), sizeof(double) * 9 * n, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(L.freem, h.freem.data(), n, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemsetAsync(L.b, 0, sizeof(double) * n, c->stream));
        if (l > 0) CUDA_OK(cudaMemsetAsync(L.u, 0, sizeof(double) * n, c->stream));
    }
    // dense inverse of the coarsest operator (Gauss-Jordan with partial pivoting on the host)
    {
        const HostLevel& Z = H.back();
        const int nz = Z.M * Z.N;
        if (nz <= 256 && H.size() > 1)
        {
            std::vector<double> a((size_t)nz * nz, 0.0), inv((size_t)nz * nz, 0.0);
            for (int i = 0; i < Z.M; i++)
                for (int j = 0; j < Z.N; j++)
                {
                    const int k = i * Z.N + j;
                    if (!Z.freem[k]) { a[(size_t)k * nz + k] = 1.0; continue; }
                    for (int di = -1; di <= 1; di++)
                        for (int dj = -1; dj <= 1; dj++)
                        {
                            const int ii = i + di, jj = j + dj;
                            if (ii < 0 || ii >= Z.M || jj < 0 || jj >= Z.N) continue;
                            const int kk = ii * Z.N + jj;
                            if (kk != k && !Z.freem[kk]) continue;      // eliminated unknowns carry zero error
                            a[(size_t)k * nz + kk] = Z.coef[(size_t)cidx(di, dj) * nz + k];
                        }
                }
            for (int k = 0; k < nz; k++) inv[(size_t)k * nz + k] = 1.0;
            bool ok = true;
            for (int col = 0; col < nz && ok; col++)
            {
                int piv = col;
                for (int r = col + 1; r < nz; r++)
                    if (std::fabs(a[(size_t)r * nz + col]) > std::fabs(a[(size_t)piv * nz + col])) piv = r;
                if (std::fabs(a[(size_t)piv * nz + col]) < 1e-300) { ok = false; break; }
                if (piv != col)
                    for (int q = 0; q < nz; q++)
                    {
                        std::swap(a[(size_t)piv * nz + q], a[(size_t)col * nz + q]);
                        std::swap(inv[(size_t)piv * nz + q], inv[(size_t)col * nz + q]);
                    }
                const double d = 1.0 / a[(size_t)col * nz + col];
                for (int q = 0; q < nz; q++) { a[(size_t)col * nz + q] *= d; inv[(size_t)col * nz + q] *= d; }
                for (int r = 0; r < nz; r++)
                {
                    if (r == col) continue;
                    const double f = a[(size_t)r * nz + col];
                    if (f == 0.0) continue;
                    for (int q = 0; q < nz; q++) { a[(size_t)r * nz + q] -= f * a[(size_t)col * nz + q]; inv[(size_t)r * nz + q] -= f * inv[(size_t)col * nz + q]; }
                }
            }
            if (ok)
            {
                CUDA_OK(cudaMalloc(&c->d_mg_inv, sizeof(double) * (size_t)nz * nz));
                CUDA_OK(cudaMemcpyAsync(c->d_mg_inv, inv.data(), sizeof(double) * (size_t)nz * nz, cudaMemcpyHostToDevice, c->stream));
                CUDA_OK(cudaStreamSynchronize(c->stream));
            }
        }
    }
    CUDA_OK(cudaMalloc(&c->d_rowscale, sizeof(double) * g.M));
    CUDA_OK(cudaMemcpyAsync(c->d_rowscale, rowscale.data(), sizeof(double) * g.M, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mg_rhs(mag2d_ctx* c, int rf)
{
    const mag2d_grid_desc& g = c->g;
    RhsArgs A;
    A.M = g.M;
    A.N = g.N;
    A.coord = g.coord;
    A.rf = rf;
    A.n_species = (int)c->sp.size();
    A.dx = g.dx;
    A.dz = g.dz;
    A.dV = g.dV;
    A.mpf = g.macroparticle_factor;
    A.mask = c->d_mask;
    A.voltage = c->d_voltage;
    A.rho = c->d_rho;
    A.charges = c->d_charges;
    A.rowscale = c->d_rowscale;
    A.b_ref = c->d_b;
    A.b_mg = c->mg[0].b;
    const bool direct = c->direct.ok && c->solver_kind != MAG2D_SOLVER_MULTIGRID;
    A.bp = direct ? c->direct.bp : nullptr;
    A.ld = c->direct.ld;
    A.hp = c->direct.hp;
    A.rowfree = c->direct.rowfree;
    A.k2row = c->direct.k2;
    A.u = rf ? c->d_uRF : c->d_u;
    const dim3 block(32, 8);
    k_rhs<<<grid2d((g.N + 1) / 2, g.M, block), block, 0, c->stream>>>(A);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mg_vcycle(mag2d_ctx* c)
{
    if (vcycle_level(c, 0)) return 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mg_residual(mag2d_ctx* c, double* resid_max, double* u_max)
{
    const MgLevel& L = c->mg[0];
    CUDA_OK(cudaMemsetAsync(c->d_scratch, 0, 2 * sizeof(double), c->stream));
    const dim3 block(32, 8);
    k_mg_residual_norm<<<grid2d(L.N, L.M, block), block, 0, c->stream>>>(level_view(L), c->d_scratch);
    c->launches++;
    double h[2];
    CUDA_OK(cudaMemcpyAsync(h, c->d_scratch, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *resid_max = h[0];
    *u_max = h[1];
    return 0;
}

// solve Op(u) = b for u (rf = 0) or uRF (rf = 1).  fixed_cycles > 0: run exactly that many V-cycles and
// do not look at the residual (no host synchronisation); otherwise iterate until the largest Jacobi
// update max|r_k/a_kk| is below tol * max|u|.
int mg_solve(mag2d_ctx* c, int rf, double tol, int max_cycles, int fixed_cycles, int* cycles, double* resid)
{
    if (!c->grid_set || c->mg.empty())
    {
        mag2d_set_error("mag2d_solve: mag2d_set_grid has not been called");
        return 1;
    }
    // the hierarchy works on mg[0].u; solving for uRF swaps the pointer for the duration of the call
    double* saved = c->mg[0].u;
    if (rf) c->mg[0].u = c->d_uRF;
    int rc = mg_rhs(c, rf);
    int done = 0;
    double r = 0.0, bm = 0.0;
    const bool direct = c->direct.ok && c->solver_kind != MAG2D_SOLVER_MULTIGRID;
    if (!rc && direct)
    {
        rc = direct_solve(c, c->mg[0].u);
        if (!rc)
        {
            if (fixed_cycles != 0)
            {
                // in-step solve: the result is exact to round-off, so the residual is only sampled (every 32nd step)
                if (c->direct_calls++ % 32 != 0) goto direct_done;
                const MgLevel& L = c->mg[0];
                const dim3 block(32, 8);
                k_mg_residual_norm<<<grid2d(L.N, L.M, block), block, 0, c->stream>>>(level_view(L), c->d_scratch + 16);
                c->launches++;
                c->monitor_armed = true;
            }
            else
                rc = mg_residual(c, &r, &bm);
        }
    direct_done:;
    }
    else if (!rc && !rf && c->extrapolate)
    {
        // warm start: the potential changes smoothly from step to step (omega_p dt << 1), so 2u_n - u_{n-1}
        // is a better first guess than u_n
        const size_t n = (size_t)c->g.M * c->g.N;
        if (!c->d_u_prev) CUDA_OK(cudaMalloc(&c->d_u_prev, sizeof(double) * n));
        if (c->have_prev) k_extrapolate<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_u, c->d_u_prev, c->mg[0].freem, n);
        else k_copy<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_u_prev, c->d_u, n);
        c->have_prev = true;
        c->launches++;
    }
    if (!rc && !direct)
    {
        if (fixed_cycles > 0)
        {
            for (; done < fixed_cycles && !rc; done++) rc = mg_vcycle(c);
            // residual monitor without a host sync: the running maxima are read by mag2d_solver_stats
            const MgLevel& L = c->mg[0];
            const dim3 block(32, 8);
            k_mg_residual_norm<<<grid2d(L.N, L.M, block), block, 0, c->stream>>>(level_view(L), c->d_scratch + 16);
            c->launches++;
            c->monitor_armed = true;
        }
        else
        {
            while (!rc)
            {
                rc = mg_residual(c, &r, &bm);
                if (rc || r <= tol * bm || done >= max_cycles) break;
                rc = mg_vcycle(c);
                done++;
            }
        }
    }
    c->mg[0].u = saved;
    if (cycles) *cycles = done;
    if (resid) *resid = bm > 0 ? r / bm : r;
    c->last_cycles = done;
    c->last_resid = bm > 0 ? r / bm : r;
    return rc;
}

int launch_rho_total(mag2d_ctx* c, double* d_out)
{
    const size_t n = grid_nodes(c);
    k_rho_total<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_rho, c->d_charges, (int)c->sp.size(), n, d_out);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_u_smooth(mag2d_ctx* c, int symmetry, double radius)
{
    const int M = c->g.M, N = c->g.N;
    const dim3 block(32, 8);
    if (symmetry)
    {
        if (M != N)
        {
            mag2d_set_error("Fields::u_smooth: smoothing of non-square matrix not implemented\n");
            return 1;
        }
        k_symmetrize<<<grid2d(M / 2 + 1, M / 2 + 1, block), block, 0, c->stream>>>(c->d_u, M);
        c->launches++;
    }
    double r2 = radius;
    if (radius > 0) r2 = (radius / c->g.dx) * (radius / c->g.dx);
    // uTmp.assign(u): d_ueff is free scratch whenever the field is not an RF field
    double* tmp = c->d_ueff;
    CUDA_OK(cudaMemcpyAsync(tmp, c->d_u, sizeof(double) * (size_t)M * N, cudaMemcpyDeviceToDevice, c->stream));
    k_smooth9<<<grid2d(N - 1, M - 2, block), block, 0, c->stream>>>(tmp, c->d_u, M, N, r2);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}
