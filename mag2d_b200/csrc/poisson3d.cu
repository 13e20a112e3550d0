// poisson3d.cu — Poisson solve of the 3-D path.
//
// The reference factorises the seven-point operator of Solver::matrix_init (src/fields3d.cpp:39-73: unit stencil
// [1,1,1,-6,1,1,1], identity rows on Dirichlet nodes, neighbours addressed by FLAT index) with UMFPACK and solves
// A^T x = b every step (Solver::solve, src/fields3d.cpp:83-95; right-hand side rho * (-macroparticle_factor/eps_0)).
// Its only geometry (Geometry::Geometry, src/fields3d.cpp:13-37) is a zero-Dirichlet box frame plus point
// electrodes, for which the operator is a constant-coefficient Laplacian on a box: it is diagonalised by sine
// transforms along y and z and tridiagonal (Toeplitz) along x.  The solve is
//     R^  = S_y (R S_z)            two FP64 matrix products with the symmetric sine matrices
//     T u^ = R^                    one Thomas solve along x per (y,z) mode, factors precomputed at set_grid
//     U   = S_y (U^ S_z) * 4/((n_y+1)(n_z+1))
// and interior electrode nodes are imposed exactly afterwards by the capacitance-matrix method (Green's functions
// of the electrode nodes are precomputed with the same solver).  Exact to round-off like the reference's LU.
//
// The sine matrix S[j][k] = sin(pi (j+1)(k+1)/(n+1)) obeys S[j][n-1-k] = (-1)^j S[j][k]: modes with even j only see the
// symmetric part x[k] + x[n-1-k] of a vector, modes with odd j the antisymmetric part x[k] - x[n-1-k].  Every transform is
// therefore done FOLDED: a row of length ld = 2 hp holds the symmetric half [0, hp) and the antisymmetric half [hp, 2 hp),
// the spectrum is kept in the same split order (even j, then odd j), and each of the four transforms becomes two products
// with hp x hp matrices — half the flops of the full products.  k_fold3d folds the right-hand side along y and z at once,
// k_unfold3d undoes both (x[k] = e + o, x[n-1-k] = e - o) while it writes the interior of the potential.
// Geometries with extended internal electrodes need a 3-D multigrid (not built yet): set_grid refuses them.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace {

// ---- batched FP64 matrix product C_b = A_b x B_b (row major), tile 64 x 32, 4 x 4 per thread, cp.async ring -----
constexpr int G3_TM = 64, G3_TN = 32, G3_KC = 32, G3_THREADS = 128, G3_STAGES = 3;
constexpr int G3_LDA = G3_KC + 2;
constexpr int G3_STAGE_DOUBLES = G3_TM * G3_LDA + G3_KC * G3_TN;
constexpr int G3_SMEM = G3_STAGES * G3_STAGE_DOUBLES * (int)sizeof(double);

struct Gemm3Args
{
    // batch entry z = 2 * outer + parity: operand X starts at X + outer * strideX + parity * halfX (the two halves of a folded transform)
    const double* A; long long strideA, halfA; int lda;     // [rows][Kdim]
    const double* B; long long strideB, halfB; int ldb;     // [Kdim][cols]
    double* C; long long strideC, halfC; int ldc;
    int rows, cols, Kdim;                                   // cols and Kdim are multiples of 32
    double scale;
};

__device__ __forceinline__ void cp16(void* smem, const void* gmem)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int W>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(W) : "memory"); }

__global__ void __launch_bounds__(G3_THREADS) k_gemm3(const __grid_constant__ Gemm3Args G)
{
    extern __shared__ __align__(16) double g3_smem[];
    const int t = threadIdx.x;
    const int r0 = blockIdx.y * G3_TM, c0 = blockIdx.x * G3_TN;
    const long long outer = blockIdx.z >> 1, par = blockIdx.z & 1;
    const double* A = G.A + outer * G.strideA + par * G.halfA;
    const double* B = G.B + outer * G.strideB + par * G.halfB;
    double* C = G.C + outer * G.strideC + par * G.halfC;
    const int tr = (t / 8) * 4, tc = (t % 8) * 4;
    const int nk = G.Kdim / G3_KC;
    auto issue = [&](int kb) {
        double* sa = g3_smem + (kb % G3_STAGES) * G3_STAGE_DOUBLES;
        double* sb = sa + G3_TM * G3_LDA;
        const int k0 = kb * G3_KC;
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const int e = t + q * G3_THREADS, r = e >> 4, c2 = (e & 15) * 2;
            const int row = min(r0 + r, G.rows - 1);
            cp16(sa + r * G3_LDA + c2, A + (size_t)row * G.lda + k0 + c2);
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const int e = t + q * G3_THREADS, r = e >> 4, c2 = (e & 15) * 2;
            cp16(sb + r * G3_TN + c2, B + (size_t)(k0 + r) * G.ldb + c0 + c2);
        }
    };
    double acc[4][4] = {};
    for (int s = 0; s < G3_STAGES - 1; s++)
    {
        if (s < nk) issue(s);
        cp_commit();
    }
    for (int kb = 0; kb < nk; kb++)
    {
        cp_wait<G3_STAGES - 2>();
        __syncthreads();
        if (kb + G3_STAGES - 1 < nk) issue(kb + G3_STAGES - 1);
        cp_commit();
        const double* sa = g3_smem + (kb % G3_STAGES) * G3_STAGE_DOUBLES;
        const double* sb = sa + G3_TM * G3_LDA;
#pragma unroll 8
        for (int k = 0; k < G3_KC; k++)
        {
            double a[4], b[4];
#pragma unroll
            for (int p = 0; p < 4; p++) a[p] = sa[(tr + p) * G3_LDA + k];
            const double2 b01 = *reinterpret_cast<const double2*>(sb + k * G3_TN + tc);
            const double2 b23 = *reinterpret_cast<const double2*>(sb + k * G3_TN + tc + 2);
            b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++)
    {
        const int r = r0 + tr + p;
        if (r >= G.rows) continue;
        double* out = C + (size_t)r * G.ldc + c0 + tc;
        *reinterpret_cast<double2*>(out) = make_double2(acc[p][0] * G.scale, acc[p][1] * G.scale);
        *reinterpret_cast<double2*>(out + 2) = make_double2(acc[p][2] * G.scale, acc[p][3] * G.scale);
    }
}

// ---- the same product on the FP64 tensor cores: mma.sync m8n8k4 (DMMA) ---------------------------------------------
// CTA tile 128 x 32 (four warps stacked along the rows, 32 x 32 per warp = 4 x 4 DMMA tiles), K in chunks of 16 through a
// 3-stage cp.async ring.  Per k4-step a warp loads 4 A and 4 B fragments (one double per lane each) for 16 DMMAs:
// 16 FMAs per shared-memory double instead of 2 with the 4 x 4 register tile of k_gemm3, and 8x fewer issue slots.
// Row strides of 20 / 36 doubles make both fragment loads bank-conflict free (lane/4 -> row * 8 banks, lane%4 -> 2 banks).
constexpr int M3_TM = 128, M3_KC = 16, M3_THREADS = 128, M3_STAGES = 3;
constexpr int M3_LDA = M3_KC + 4;
template <int TN>
struct M3Cfg
{
    static constexpr int LDB = TN + 4;
    static constexpr int STAGE_DOUBLES = M3_TM * M3_LDA + M3_KC * LDB;
    static constexpr int SMEM = M3_STAGES * STAGE_DOUBLES * (int)sizeof(double);
};

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// TN = 32 or 64 columns per CTA (32 x TN per warp)
template <int TN>
__global__ void __launch_bounds__(M3_THREADS) k_gemm3_mma(const __grid_constant__ Gemm3Args G)
{
    extern __shared__ __align__(16) double m3_smem[];
    constexpr int LDB = M3Cfg<TN>::LDB, STAGE = M3Cfg<TN>::STAGE_DOUBLES, NQ = TN / 8;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int r0 = blockIdx.y * M3_TM, c0 = blockIdx.x * TN;
    const long long outer = blockIdx.z >> 1, par = blockIdx.z & 1;
    const double* A = G.A + outer * G.strideA + par * G.halfA;
    const double* B = G.B + outer * G.strideB + par * G.halfB;
    double* C = G.C + outer * G.strideC + par * G.halfC;
    const int nk = G.Kdim / M3_KC;
    auto issue = [&](int kb) {
        double* sa = m3_smem + (kb % M3_STAGES) * STAGE;
        double* sb = sa + M3_TM * M3_LDA;
        const int k0 = kb * M3_KC;
        // A tile: 128 rows x 16 doubles = 1024 16-byte pieces (8 per row), 8 per thread
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const int e = t + q * M3_THREADS, r = e >> 3, c2 = (e & 7) * 2;
            const int row = min(r0 + r, G.rows - 1);
            cp16(sa + r * M3_LDA + c2, A + (size_t)row * G.lda + k0 + c2);
        }
        // B tile: 16 k x TN columns = 8 TN pieces, TN / 16 per thread
#pragma unroll
        for (int q = 0; q < TN / 16; q++)
        {
            const int e = t + q * M3_THREADS, r = e / (TN / 2), c2 = (e % (TN / 2)) * 2;
            cp16(sb + r * LDB + c2, B + (size_t)(k0 + r) * G.ldb + c0 + c2);
        }
    };
    double acc[4][NQ][2] = {};
    for (int s = 0; s < M3_STAGES - 1; s++)
    {
        if (s < nk) issue(s);
        cp_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;     // fragment row / column of this lane
    for (int kb = 0; kb < nk; kb++)
    {
        cp_wait<M3_STAGES - 2>();
        __syncthreads();
        if (kb + M3_STAGES - 1 < nk) issue(kb + M3_STAGES - 1);
        cp_commit();
        const double* sa = m3_smem + (kb % M3_STAGES) * STAGE + (warp * 32 + fr) * M3_LDA + fc;
        const double* sb = m3_smem + (kb % M3_STAGES) * STAGE + M3_TM * M3_LDA + fc * LDB + fr;
#pragma unroll
        for (int k4 = 0; k4 < M3_KC / 4; k4++)
        {
            double a[4], b[NQ];
#pragma unroll
            for (int p = 0; p < 4; p++) a[p] = sa[p * 8 * M3_LDA + k4 * 4];            // A[rb*8 + fr][k4*4 + fc]
#pragma unroll
            for (int q = 0; q < NQ; q++) b[q] = sb[k4 * 4 * LDB + q * 8];              // B[k4*4 + fc][cb*8 + fr]
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < NQ; q++) dmma884(acc[p][q], a[p], b[q]);
        }
    }
    // D fragment: row fr, columns 2*fc, 2*fc+1 of every 8 x 8 block
#pragma unroll
    for (int p = 0; p < 4; p++)
    {
        const int r = r0 + warp * 32 + p * 8 + fr;
        if (r >= G.rows) continue;
        double* out = C + (size_t)r * G.ldc + c0 + 2 * fc;
#pragma unroll
        for (int q = 0; q < NQ; q++) *reinterpret_cast<double2*>(out + q * 8) = make_double2(acc[p][q][0] * G.scale, acc[p][q][1] * G.scale);
    }
}

// ---- folding --------------------------------------------------------------------------------------------------------
// natural planes [ldj][ldk] (entries (jl, kl), jl < n_j, kl < n_k) -> folded planes: with s/a = symmetric / antisymmetric part
// along an axis, the quadrants hold [ss | sa ; as | aa] (rows: y part, columns: z part).  The middle element of an odd
// length is its own mirror image: it goes to the symmetric half as it is.  Every entry of the folded plane is written
// (zeros in the padding).  One block per (pair of rows jj / n_j-1-jj, plane), threads along z.
__global__ void __launch_bounds__(128) k_fold3d(int n_j, int n_k, int hpj, int hpk, const double* __restrict__ nat, double* __restrict__ fold)
{
    const int jj = blockIdx.x, jm = n_j - 1 - jj, ldk = 2 * hpk;
    const size_t base = (size_t)blockIdx.y * (2 * hpj) * ldk;
    const double* n0 = nat + base + (size_t)jj * ldk;
    const double* n1 = nat + base + (size_t)max(jm, 0) * ldk;
    double* f0 = fold + base + (size_t)jj * ldk;
    double* f1 = fold + base + (size_t)(hpj + jj) * ldk;
    const bool row_s = jj <= jm, row_a = jj < jm;
    for (int kk = threadIdx.x; kk < hpk; kk += blockDim.x)
    {
        const int km = n_k - 1 - kk;
        const bool col_s = kk <= km, col_a = kk < km;
        double a = 0.0, b = 0.0, c = 0.0, d = 0.0;      // (j, k), (j, k~), (j~, k), (j~, k~)
        if (row_s && col_s)
        {
            a = n0[kk];
            if (col_a) b = n0[km];
            if (row_a)
            {
                c = n1[kk];
                if (col_a) d = n1[km];
            }
        }
        const double s0 = a + b, a0 = col_a ? a - b : 0.0, s1 = c + d, a1 = col_a ? c - d : 0.0;
        f0[kk] = s0 + s1;
        f0[hpk + kk] = a0 + a1;
        f1[kk] = row_a ? s0 - s1 : 0.0;
        f1[hpk + kk] = row_a ? a0 - a1 : 0.0;
    }
}

// folded planes [ee | eo ; oe | oo] of the inverse transforms -> interior nodes of the potential: plane il of the block is plane
// i0 + il + 1 of u.  x[k] = e + o, x[n-1-k] = e - o along both axes.
__global__ void __launch_bounds__(128) k_unfold3d(int n_j, int n_k, int hpj, int hpk, int K, int N, int i0, const double* __restrict__ fold,
                                                  double* __restrict__ u)
{
    const int jj = blockIdx.x, jm = n_j - 1 - jj, ldk = 2 * hpk;
    if (jj > jm) return;
    const size_t base = (size_t)blockIdx.y * (2 * hpj) * ldk;
    const double* f0 = fold + base + (size_t)jj * ldk;
    const double* f1 = fold + base + (size_t)(hpj + jj) * ldk;
    double* up = u + ((size_t)(i0 + (int)blockIdx.y + 1) * K + 1) * N + 1;       // node (i, 1, 1)
    double* u0 = up + (size_t)jj * N;
    double* u1 = up + (size_t)jm * N;
    for (int kk = threadIdx.x; kk < hpk; kk += blockDim.x)
    {
        const int km = n_k - 1 - kk;
        if (kk > km) break;
        const double ee = f0[kk], eo = f0[hpk + kk], oe = f1[kk], oo = f1[hpk + kk];
        const double p = ee + eo, q = oe + oo, r = ee - eo, t = oe - oo;
        u0[kk] = p + q;
        if (kk < km) u0[km] = r + t;
        if (jj < jm)
        {
            u1[kk] = p - q;
            if (kk < km) u1[km] = r - t;
        }
    }
}

constexpr int RHS_MAX_CHARGED = 8;
struct Rhs3Args
{
    int M, K, N, n_i, n_j, n_k, hpj, hpk, n_species;
    int ne;                              // electrode nodes inside the box (0: interior_fixed is all zero and is not read)
    int i_first;                         // first x plane of this launch (blockIdx.y counts from it)
    double factor;                       // -macroparticle_factor / eps_0
    const unsigned char* mask;           // MAG2D_FREE or Dirichlet (anything else)
    const unsigned char* interior_fixed; // 1 on electrode nodes inside the box (capacitance method)
    const double* voltage;
    const unsigned long long* rho;       // [n_species][M*K*N] Q32 counts
    int n_charged;                       // the species that carry charge, in species order
    int charged[RHS_MAX_CHARGED];
    double charge[RHS_MAX_CHARGED];
    double* b;                           // reference right-hand side, or null (only the residual check reads it)
    double* u;                           // Dirichlet nodes receive their voltage
    double* R;                           // [n_i][2 hpj][2 hpk] interior right-hand side with the frame values moved over, FOLDED along y and z
};

// Solver::solve's right-hand side + the reduction to the interior block, which leaves the kernel already FOLDED (k_fold3d's
// layout).  A thread owns the four nodes (j, k), (j, k~), (j~, k), (j~, k~) that are mirror images of each other in the
// interior block (j~ = K-1-j, k~ = n_k+1-k): it forms their right-hand sides and the four folded combinations in registers —
// no shared memory, no barrier.  Block = 128 threads along k, grid = (row pairs, planes).
__global__ void __launch_bounds__(128) k_rhs3d(const __grid_constant__ Rhs3Args A)
{
    const size_t n = (size_t)A.M * A.K * A.N;
    const long long sj = A.N, si = (long long)A.K * A.N;
    const int i = A.i_first + (int)blockIdx.y, j0 = (int)blockIdx.x, j1 = A.K - 1 - j0;
    const bool has_j1 = j1 > j0;                          // the middle row of an odd K is its own mirror image
    const int il = i - 1, jj = j0 - 1;                    // interior plane; interior rows jj and n_j - 1 - jj
    const bool plane_in = il >= 0 && il < A.n_i;
    const int kpairs = (A.n_k + 3) / 2;
    for (int kk = threadIdx.x; kk < kpairs; kk += blockDim.x)
    {
        const int kb = A.n_k + 1 - kk;
        const bool has_kb = kb > kk && kb < A.N;          // a free k = N-1 face (n_k = N-1) leaves node k = 0 without a partner
        size_t m[4];
        bool valid[4];
        unsigned char mk[4];
        double q[4], r[4];
#pragma unroll
        for (int e = 0; e < 4; e++)
        {
            valid[e] = ((e & 1) == 0 || has_kb) && ((e & 2) == 0 || has_j1);
            m[e] = ((size_t)i * A.K + (e & 2 ? j1 : j0)) * A.N + (e & 1 ? kb : kk);
            mk[e] = valid[e] ? A.mask[m[e]] : (unsigned char)MAG2D_FIXED;
            q[e] = 0.0;
            r[e] = 0.0;
        }
        // species outside, nodes inside: the four loads of a species are in flight together
        for (int s = 0; s < A.n_charged; s++)
        {
            const unsigned long long* rho = A.rho + (size_t)A.charged[s] * n;
            unsigned long long w[4];
#pragma unroll
            for (int e = 0; e < 4; e++) w[e] = valid[e] ? rho[m[e]] : 0ULL;
#pragma unroll
            for (int e = 0; e < 4; e++) q[e] += A.charge[s] * ((double)(long long)w[e] * 2.3283064365386963e-10);
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
        {
            if (!valid[e]) continue;
            const int j = e & 2 ? j1 : j0, k = e & 1 ? kb : kk;
            const bool fixed = mk[e] != MAG2D_FREE;
            const double b = fixed ? A.voltage[m[e]] : q[e] * A.factor;
            if (A.b) A.b[m[e]] = b;
            if (fixed) A.u[m[e]] = b;
            const int jl = j - 1, kl = k - 1;
            if (plane_in && jl >= 0 && jl < A.n_j && kl >= 0 && kl < A.n_k && !(A.ne && A.interior_fixed[m[e]]))
            {
                double v = b;
                // neighbours by flat index, exactly as the reference's matrix rows address them (fields3d.cpp:61-67);
                // Dirichlet neighbours on the frame contribute their value to the right-hand side.  Only the outermost
                // shell of the interior block has such neighbours (the flat-index neighbours of every other interior node
                // are interior nodes themselves)
                if (il == 0 || il == A.n_i - 1 || jl == 0 || jl == A.n_j - 1 || kl == 0 || kl == A.n_k - 1)
                {
                    const long long mm = (long long)m[e];
#pragma unroll 1
                    for (int t = 0; t < 6; t++)
                    {
                        const long long nb = mm + (t == 0 ? -si : t == 1 ? -sj : t == 2 ? -1LL : t == 3 ? 1LL : t == 4 ? sj : si);
                        if (A.mask[nb] != MAG2D_FREE && !A.interior_fixed[nb]) v -= A.voltage[nb];
                    }
                }
                r[e] = v;
            }
        }
        if (!plane_in || jj < 0 || kk == 0) continue;
        // interior column kl = kk - 1 and its mirror image n_k - 1 - kl (node kb); the padding of the folded block is never
        // written: it is zero from set_grid on and every later stage of the solve maps zeros to zeros
        const int kl = kk - 1, ldk = 2 * A.hpk;
        double* f0 = A.R + ((size_t)il * (2 * A.hpj) + jj) * ldk;
        double* f1 = f0 + (size_t)A.hpj * ldk;
        const double s0 = r[0] + r[1], a0 = has_kb ? r[0] - r[1] : 0.0, s1 = r[2] + r[3], a1 = has_kb ? r[2] - r[3] : 0.0;
        f0[kl] = s0 + s1;
        f0[A.hpk + kl] = a0 + a1;
        f1[kl] = has_j1 ? s0 - s1 : 0.0;
        f1[A.hpk + kl] = has_j1 ? a0 - a1 : 0.0;
    }
}

// Thomas factors of the Toeplitz systems [1, d, 1], d = -6 + lam_j + lam_k: inv[i][mode] = 1 / (d - inv[i-1][mode])
// (a padding mode gets d = -6: its right-hand side is zero and stays zero)
__global__ void k_thomas_setup(int n_i, int n_j, int n_k, int ldj, int ldk, double* __restrict__ inv)
{
    const int mode = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode >= ldj * ldk) return;
    // folded spectrum: entry l of a row is mode 2 l (l < hp) or mode 2 (l - hp) + 1 (counted from 0); the rest is padding
    const int jl = mode / ldk, kl = mode % ldk, hpj = ldj / 2, hpk = ldk / 2;
    const int mj = jl < hpj ? 2 * jl : 2 * (jl - hpj) + 1, mk = kl < hpk ? 2 * kl : 2 * (kl - hpk) + 1;
    double d = -6.0;
    if (mj < n_j && mk < n_k) d += 2.0 * cospi((double)(mj + 1) / (double)(n_j + 1)) + 2.0 * cospi((double)(mk + 1) / (double)(n_k + 1));
    double c = 0.0;
    for (int i = 0; i < n_i; i++)
    {
        c = 1.0 / (d - c);
        inv[(size_t)i * ldj * ldk + mode] = c;
    }
}

// y_i = (r_i - y_(i-1)) inv_i ; x_i = y_i - inv_i x_(i+1); one thread per (y,z) mode, coalesced across modes.
// A 256^3 grid has only 65k modes — 14 warps per SM — so the recurrence is fed through a register ring: the loads of the
// next TH_U planes are in flight while the current TH_U are consumed (the recurrence itself is two flops per plane).
constexpr int TH_U = 8;
__global__ void __launch_bounds__(64) k_thomas_solve(int n_i, int plane, const double* __restrict__ inv, double* v, double scale)
{
    const int mode = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode >= plane) return;
    double* vp = v + mode;
    const double* cp = inv + mode;
    double ra[TH_U], fa[TH_U], rb[TH_U], fb[TH_U];        // two register buffers, addressed statically
    auto fetch = [&](double (&r)[TH_U], double (&f)[TH_U], int first, int dir) {
#pragma unroll
        for (int q = 0; q < TH_U; q++)
        {
            const int i = first + dir * q;
            if (i >= 0 && i < n_i)
            {
                r[q] = vp[(size_t)i * plane];
                f[q] = __ldg(cp + (size_t)i * plane);
            }
        }
    };
    double y = 0.0;
    auto forward = [&](const double (&r)[TH_U], const double (&f)[TH_U], int first) {
#pragma unroll
        for (int q = 0; q < TH_U; q++)
        {
            const int i = first + q;
            if (i < n_i)
            {
                y = (r[q] - y) * f[q];
                vp[(size_t)i * plane] = y;
            }
        }
    };
    // the planes fetched ahead have not been written yet in the running sweep
    fetch(ra, fa, 0, 1);
    for (int i0 = 0; i0 < n_i; i0 += 2 * TH_U)
    {
        fetch(rb, fb, i0 + TH_U, 1);
        forward(ra, fa, i0);
        fetch(ra, fa, i0 + 2 * TH_U, 1);
        forward(rb, fb, i0 + TH_U);
    }
    // the two inverse transforms carry the factor 4 / ((n_j+1)(n_k+1)): it is folded into the stores of the back
    // substitution (the recurrence itself runs on the unscaled x)
    double x = 0.0;
    auto backward = [&](const double (&r)[TH_U], const double (&f)[TH_U], int first) {
#pragma unroll
        for (int q = 0; q < TH_U; q++)
        {
            const int i = first - q;
            if (i >= 0)
            {
                x = r[q] - f[q] * x;
                vp[(size_t)i * plane] = x * scale;
            }
        }
    };
    fetch(ra, fa, n_i - 1, -1);
    for (int i0 = n_i - 1; i0 >= 0; i0 -= 2 * TH_U)
    {
        fetch(rb, fb, i0 - TH_U, -1);
        backward(ra, fa, i0);
        fetch(ra, fa, i0 - 2 * TH_U, -1);
        backward(rb, fb, i0 - TH_U);
    }
}

// alpha = Cinv (V_E - u0[E]);  one block
__global__ void k_capacitance(int ne, const int* __restrict__ nodes, const double* __restrict__ volts, const double* __restrict__ cinv,
                              const double* __restrict__ u, double* __restrict__ alpha)
{
    __shared__ double d[64];
    const int t = threadIdx.x;
    if (t < ne) d[t] = volts[t] - u[nodes[t]];
    __syncthreads();
    if (t < ne)
    {
        double a = 0.0;
        for (int q = 0; q < ne; q++) a += cinv[t * ne + q] * d[q];
        alpha[t] = a;
    }
}

__global__ void k_add_green(size_t n, int ne, const double* __restrict__ alpha, const double* __restrict__ green, double* __restrict__ u)
{
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x)
    {
        double s = u[m];
        for (int e = 0; e < ne; e++) s += alpha[e] * green[(size_t)e * n + m];
        u[m] = s;
    }
}

__global__ void k_residual3d(int M, int K, int N, const unsigned char* __restrict__ mask, const double* __restrict__ u,
                             const double* __restrict__ b, double* __restrict__ out)
{
    const size_t n = (size_t)M * K * N;
    const long long sj = N, si = (long long)K * N;
    double r = 0.0, um = 0.0;
    for (size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x)
    {
        const double c = u[m];
        um = fmax(um, fabs(c));
        if (mask[m] != MAG2D_FREE) { r = fmax(r, fabs(c - b[m])); continue; }
        const double res = b[m] - (u[m - si] + u[m - sj] + u[m - 1] - 6.0 * c + u[m + 1] + u[m + sj] + u[m + si]);
        r = fmax(r, fabs(res) / 6.0);
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
        um = fmax(um, __shfl_xor_sync(0xffffffffu, um, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(r));
        atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)__double_as_longlong(um));
    }
}

// both halves of `outer` folded products in one launch
int gemm3(mag2d_ctx* c, const Gemm3Args& G, int outer)
{
    static const bool use_fma = getenv("MAG2D_GEMM") && !strcmp(getenv("MAG2D_GEMM"), "fma");
    if (use_fma)
    {
        const dim3 grid(G.cols / G3_TN, (G.rows + G3_TM - 1) / G3_TM, 2 * outer);
        k_gemm3<<<grid, G3_THREADS, G3_SMEM, c->stream>>>(G);
    }
    else
    {
        // 64 columns per CTA (210 registers, two CTAs per SM) measured slower on C5: 6.23 against 6.19 ms per step
        const dim3 grid(G.cols / 32, (G.rows + M3_TM - 1) / M3_TM, 2 * outer);
        k_gemm3_mma<32><<<grid, M3_THREADS, M3Cfg<32>::SMEM, c->stream>>>(G);
    }
    c->launches++;
    return 0;
}

// x planes [a, a + np): folded right-hand side in D.R (natural: in D.T, folded here first) -> transformed along z and y, in D.R
int forward_planes(mag2d_ctx* c, int a, int np, bool natural)
{
    Direct3D& D = c->direct3;
    const long long plane = (long long)D.ldj * D.ldk;
    if (natural)
    {
        k_fold3d<<<dim3((unsigned)D.hpj, (unsigned)np), 128, 0, c->stream>>>(D.n_j, D.n_k, D.hpj, D.hpk, D.T + a * plane, D.R + a * plane);
        c->launches++;
    }
    Gemm3Args G;
    memset(&G, 0, sizeof(G));
    G.scale = 1.0;
    // along z: T = R E_z, all rows of the np planes at once; the halves are the two column blocks
    G.A = D.R + a * plane; G.lda = D.ldk; G.halfA = D.hpk;
    G.B = D.Sz_f; G.ldb = D.hpk; G.halfB = (long long)D.hpk * D.hpk;
    G.C = D.T + a * plane; G.ldc = D.ldk; G.halfC = D.hpk;
    G.rows = np * D.ldj; G.cols = D.hpk; G.Kdim = D.hpk;
    gemm3(c, G, 1);
    // along y: R_i = E_y^T T_i for every plane; the halves are the two row blocks
    G.A = D.Sy_f; G.lda = D.hpj; G.strideA = 0; G.halfA = (long long)D.hpj * D.hpj;
    G.B = D.T + a * plane; G.ldb = D.ldk; G.strideB = plane; G.halfB = (long long)D.hpj * D.ldk;
    G.C = D.R + a * plane; G.ldc = D.ldk; G.strideC = plane; G.halfC = (long long)D.hpj * D.ldk;
    G.rows = D.hpj; G.cols = D.ldk; G.Kdim = D.hpj;
    gemm3(c, G, np);
    return 0;
}

// x planes [a, a + np): folded spectrum in D.R -> inverse transforms along y and z -> interior nodes of those planes of u
int inverse_planes(mag2d_ctx* c, int a, int np, double* u)
{
    Direct3D& D = c->direct3;
    const long long plane = (long long)D.ldj * D.ldk;
    Gemm3Args G;
    memset(&G, 0, sizeof(G));
    G.scale = 1.0;
    G.A = D.Sy_i; G.lda = D.hpj; G.strideA = 0; G.halfA = (long long)D.hpj * D.hpj;
    G.B = D.R + a * plane; G.ldb = D.ldk; G.strideB = plane; G.halfB = (long long)D.hpj * D.ldk;
    G.C = D.T + a * plane; G.ldc = D.ldk; G.strideC = plane; G.halfC = (long long)D.hpj * D.ldk;
    G.rows = D.hpj; G.cols = D.ldk; G.Kdim = D.hpj;
    gemm3(c, G, np);
    G.A = D.T + a * plane; G.lda = D.ldk; G.strideA = 0; G.halfA = D.hpk;
    G.B = D.Sz_i; G.ldb = D.hpk; G.strideB = 0; G.halfB = (long long)D.hpk * D.hpk;
    G.C = D.R + a * plane; G.ldc = D.ldk; G.strideC = 0; G.halfC = D.hpk;
    G.rows = np * D.ldj; G.cols = D.hpk; G.Kdim = D.hpk;
    gemm3(c, G, 1);
    k_unfold3d<<<dim3((unsigned)D.hpj, (unsigned)np), 128, 0, c->stream>>>(D.n_j, D.n_k, D.hpj, D.hpk, c->g.K, c->g.N, a, D.R + a * plane, u);
    c->launches++;
    return 0;
}

// interior right-hand side (folded in D.R as k_rhs3d leaves it, or natural in D.T) -> potential on the interior nodes of u (frame untouched)
int solve_interior(mag2d_ctx* c, double* u, bool natural)
{
    Direct3D& D = c->direct3;
    const int plane = D.ldj * D.ldk;
    if (forward_planes(c, 0, D.n_i, natural)) return 1;
    k_thomas_solve<<<(plane + 63) / 64, 64, 0, c->stream>>>(D.n_i, plane, D.inv, D.R, 4.0 / ((double)(D.n_j + 1) * (double)(D.n_k + 1)));
    c->launches++;
    if (inverse_planes(c, 0, D.n_i, u)) return 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- the same solve on N ranks (SURVEY.md §8e "optional later: reduce-scatter + slab-parallel ... + allgather") -------------
// The sine transforms act inside an x plane, the Thomas recurrences along x inside a (y, z) mode: rank r transforms its slab of
// x planes, the slabs are transposed over NVLink (ncclSend / ncclRecv, every pair exchanges 1/N^2 of the block) so that every
// rank holds ALL planes of its slab of y rows, runs the recurrences of those modes, transposes back, applies the inverse
// transforms to its planes and broadcasts them: every rank ends with the full potential.  Each number is produced by exactly the
// kernels and operation order of the one-rank solve, so the potential is bit-identical for any N.
int slab_setup(mag2d_ctx* c)
{
    Direct3D& D = c->direct3;
    const int nr = c->nranks, me = c->rank;
    D.pi0.assign(nr + 1, 0);
    D.pj0.assign(nr + 1, 0);
    // planes are shared out by their index in the potential (interior plane i is plane i + 1 of u, the two frame planes ride along with
    // the first and the last rank): when M divides by N every rank owns M / N whole planes of u and one all-gather spreads them
    const int pu = (c->g.M + nr - 1) / nr, pj = (D.ldj + nr - 1) / nr;
    for (int q = 0; q <= nr; q++)
    {
        D.pi0[q] = q == nr ? D.n_i : std::min(std::max(q * pu - 1, 0), D.n_i);
        D.pj0[q] = std::min(q * pj, D.ldj);
    }
    const size_t rows_me = (size_t)(D.pj0[me + 1] - D.pj0[me]);
    const size_t slab_x = (size_t)std::max(D.pi0[me + 1] - D.pi0[me], 1) * D.ldj * D.ldk;
    const size_t slab_m = std::max<size_t>((size_t)D.n_i * rows_me * D.ldk, 1);
    CUDA_OK(cudaMalloc(&D.V, sizeof(double) * slab_m));
    CUDA_OK(cudaMalloc(&D.inv_slab, sizeof(double) * slab_m));
    CUDA_OK(cudaMalloc(&D.Sb, sizeof(double) * slab_x));
    CUDA_OK(cudaMalloc(&D.Xb, sizeof(double) * slab_x));
    if (rows_me > 0)
        CUDA_OK(cudaMemcpy2DAsync(D.inv_slab, rows_me * D.ldk * sizeof(double), D.inv + (size_t)D.pj0[me] * D.ldk, (size_t)D.ldj * D.ldk * sizeof(double),
                                  rows_me * D.ldk * sizeof(double), (size_t)D.n_i, cudaMemcpyDeviceToDevice, c->stream));
    D.slab_ok = true;
    return 0;
}

int solve_interior_slab(mag2d_ctx* c, double* u)
{
    Direct3D& D = c->direct3;
    if (!D.slab_ok && slab_setup(c)) return 1;
    const int nr = c->nranks, me = c->rank;
    const int a = D.pi0[me], b = D.pi0[me + 1], np = b - a;
    const size_t plane = (size_t)D.ldj * D.ldk;
    const size_t rows_me = (size_t)(D.pj0[me + 1] - D.pj0[me]);
    const size_t esz = sizeof(double);
    // forward transforms of this rank's planes
    if (np > 0 && forward_planes(c, a, np, false)) return 1;
    // x slabs -> mode slabs: the rows [pj0[q], pj0[q+1]) of my planes go to rank q; their rows of my slab arrive from everybody
    std::vector<size_t> off(nr + 1, 0);
    for (int q = 0; q < nr; q++) off[q + 1] = off[q] + (size_t)np * (D.pj0[q + 1] - D.pj0[q]) * D.ldk;
    for (int q = 0; q < nr && np > 0; q++)
    {
        const size_t rows_q = (size_t)(D.pj0[q + 1] - D.pj0[q]);
        if (rows_q == 0) continue;
        double* dst = q == me ? D.V + (size_t)a * rows_me * D.ldk : D.Sb + off[q];
        CUDA_OK(cudaMemcpy2DAsync(dst, rows_q * D.ldk * esz, D.R + a * plane + (size_t)D.pj0[q] * D.ldk, plane * esz, rows_q * D.ldk * esz, (size_t)np,
                                  cudaMemcpyDeviceToDevice, c->stream));
    }
    if (comm_group_start()) return 1;
    for (int q = 0; q < nr; q++)
    {
        if (q == me) continue;
        const size_t rows_q = (size_t)(D.pj0[q + 1] - D.pj0[q]), np_q = (size_t)(D.pi0[q + 1] - D.pi0[q]);
        if (np > 0 && rows_q > 0 && comm_send(c, D.Sb + off[q], (size_t)np * rows_q * D.ldk, q)) return 1;
        if (np_q > 0 && rows_me > 0 && comm_recv(c, D.V + (size_t)D.pi0[q] * rows_me * D.ldk, np_q * rows_me * D.ldk, q)) return 1;
    }
    if (comm_group_end()) return 1;
    if (rows_me > 0)
    {
        const int plane_me = (int)(rows_me * D.ldk);
        k_thomas_solve<<<(plane_me + 63) / 64, 64, 0, c->stream>>>(D.n_i, plane_me, D.inv_slab, D.V, 4.0 / ((double)(D.n_j + 1) * (double)(D.n_k + 1)));
        c->launches++;
    }
    // mode slabs -> x slabs
    if (comm_group_start()) return 1;
    for (int q = 0; q < nr; q++)
    {
        if (q == me) continue;
        const size_t rows_q = (size_t)(D.pj0[q + 1] - D.pj0[q]), np_q = (size_t)(D.pi0[q + 1] - D.pi0[q]);
        if (np_q > 0 && rows_me > 0 && comm_send(c, D.V + (size_t)D.pi0[q] * rows_me * D.ldk, np_q * rows_me * D.ldk, q)) return 1;
        if (np > 0 && rows_q > 0 && comm_recv(c, D.Xb + off[q], (size_t)np * rows_q * D.ldk, q)) return 1;
    }
    if (comm_group_end()) return 1;
    for (int q = 0; q < nr && np > 0; q++)
    {
        const size_t rows_q = (size_t)(D.pj0[q + 1] - D.pj0[q]);
        if (rows_q == 0) continue;
        const double* src = q == me ? D.V + (size_t)a * rows_me * D.ldk : D.Xb + off[q];
        CUDA_OK(cudaMemcpy2DAsync(D.R + a * plane + (size_t)D.pj0[q] * D.ldk, plane * esz, src, rows_q * D.ldk * esz, rows_q * D.ldk * esz, (size_t)np,
                                  cudaMemcpyDeviceToDevice, c->stream));
    }
    // inverse transforms of this rank's planes, unfolded into its planes of the potential
    if (np > 0 && inverse_planes(c, a, np, u)) return 1;
    // every rank's planes to everybody (the frame planes 0 and M-1 are Dirichlet values every rank has written itself)
    const size_t node_plane = (size_t)c->g.K * c->g.N;
    if (c->g.M % nr == 0)
    {
        if (comm_allgather_inplace(c, u, (size_t)(c->g.M / nr) * node_plane)) return 1;
    }
    else
    {
        if (comm_group_start()) return 1;
        for (int q = 0; q < nr; q++)
        {
            const size_t np_q = (size_t)(D.pi0[q + 1] - D.pi0[q]);
            if (np_q > 0 && comm_broadcast(c, u + (size_t)(D.pi0[q] + 1) * node_plane, np_q * node_plane, q)) return 1;
        }
        if (comm_group_end()) return 1;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// the two half matrices of a folded sine transform of length n: E_b[kk][m] = sin(pi (2m + b + 1)(kk + 1) / (n + 1)) for the
// ceil(n/2) (b = 0) or floor(n/2) (b = 1) input pairs kk and modes m; [2][hp][hp], zero padding.  transposed: [b][m][kk].
std::vector<double> half_sines(int n, int hp, bool transposed)
{
    std::vector<double> E(2 * (size_t)hp * hp, 0.0);
    for (int b = 0; b < 2; b++)
    {
        const int h = b ? n / 2 : (n + 1) / 2;
        for (int kk = 0; kk < h; kk++)
            for (int m = 0; m < h; m++)
            {
                const long long p = (long long)(2 * m + b + 1) * (kk + 1) % (2LL * (n + 1));
                const double v = (double)sinl(M_PIl * (long double)p / (long double)(n + 1));
                E[(size_t)b * hp * hp + (transposed ? (size_t)m * hp + kk : (size_t)kk * hp + m)] = v;
            }
    }
    return E;
}

}  // namespace

void direct3d_free(mag2d_ctx* c)
{
    cudaFree(c->direct3.V); cudaFree(c->direct3.inv_slab); cudaFree(c->direct3.Sb); cudaFree(c->direct3.Xb);
    Direct3D& D = c->direct3;
    cudaFree(D.Sy_f); cudaFree(D.Sy_i); cudaFree(D.Sz_f); cudaFree(D.Sz_i); cudaFree(D.inv); cudaFree(D.R); cudaFree(D.T); cudaFree(D.interior_fixed);
    cudaFree(D.e_nodes); cudaFree(D.e_volts); cudaFree(D.cinv); cudaFree(D.alpha); cudaFree(D.green);
    D = Direct3D();
}

int direct3d_setup(mag2d_ctx* c)
{
    direct3d_free(c);
    Direct3D& D = c->direct3;
    const int M = c->g.M, K = c->g.K, N = c->g.N;
    const size_t n = (size_t)M * K * N;
    const unsigned char* mask = c->h_mask.data();
    auto fixed = [&](int i, int j, int k) { return mask[((size_t)i * K + j) * N + k] != MAG2D_FREE; };
    // the five faces the reference fixes must be Dirichlet; the k = N-1 face is either Dirichlet or (the reference's
    // off-by-one, fields3d.cpp:28) free, in which case its free nodes couple to the k = 0 node of the next row
    bool top_fixed = true, top_free = true;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < K; j++)
        {
            if (!fixed(i, j, 0)) { mag2d_set_error("3-D solver: the k = 0 face must be Dirichlet"); return 1; }
            const bool frame_ij = i == 0 || i == M - 1 || j == 0 || j == K - 1;
            if (frame_ij) continue;
            if (fixed(i, j, N - 1)) top_free = false;
            else top_fixed = false;
        }
    for (int j = 0; j < K; j++)
        for (int k = 0; k < N; k++)
            if (!fixed(0, j, k) || !fixed(M - 1, j, k)) { mag2d_set_error("3-D solver: the x faces must be Dirichlet"); return 1; }
    for (int i = 0; i < M; i++)
        for (int k = 0; k < N; k++)
            if (!fixed(i, 0, k) || !fixed(i, K - 1, k)) { mag2d_set_error("3-D solver: the y faces must be Dirichlet"); return 1; }
    if (!top_fixed && !top_free) { mag2d_set_error("3-D solver: the k = N-1 face must be entirely Dirichlet or entirely free"); return 1; }
    D.n_i = M - 2;
    D.n_j = K - 2;
    D.n_k = top_fixed ? N - 2 : N - 1;
    if (D.n_i < 1 || D.n_j < 1 || D.n_k < 1) { mag2d_set_error("3-D solver: grid too small"); return 1; }
    if (!top_fixed)
        // the wrapped neighbour (i, j+1, 0) of a free top node must carry zero volts for the sine transform to apply
        for (int i = 1; i < M - 1; i++)
            for (int j = 1; j < K; j++)
                if (c->h_voltage[((size_t)i * K + j) * N] != 0.0)
                {
                    mag2d_set_error("3-D solver: a free k = N-1 face needs zero volts on the k = 0 face");
                    return 1;
                }
    D.hpj = ((D.n_j + 1) / 2 + 31) / 32 * 32;
    D.hpk = ((D.n_k + 1) / 2 + 31) / 32 * 32;
    D.ldj = 2 * D.hpj;
    D.ldk = 2 * D.hpk;
    // electrode nodes inside the box
    std::vector<unsigned char> interior_fixed(n, 0);
    std::vector<int> e_nodes;
    std::vector<double> e_volts;
    for (int i = 1; i <= D.n_i; i++)
        for (int j = 1; j <= D.n_j; j++)
            for (int k = 1; k <= D.n_k; k++)
                if (fixed(i, j, k))
                {
                    const size_t m = ((size_t)i * K + j) * N + k;
                    interior_fixed[m] = 1;
                    e_nodes.push_back((int)m);
                    e_volts.push_back(c->h_voltage[m]);
                }
    D.ne = (int)e_nodes.size();
    if (D.ne > 64)
    {
        mag2d_set_error("3-D solver: more than 64 electrode nodes inside the box (extended electrodes need the 3-D multigrid, not built yet)");
        return 1;
    }
    const size_t block = (size_t)D.n_i * D.ldj * D.ldk;
    // y transforms multiply from the left (forward E^T, inverse E), z transforms from the right (forward E, inverse E^T)
    const std::vector<double> Sy_f = half_sines(D.n_j, D.hpj, true), Sy_i = half_sines(D.n_j, D.hpj, false);
    const std::vector<double> Sz_f = half_sines(D.n_k, D.hpk, false), Sz_i = half_sines(D.n_k, D.hpk, true);
    CUDA_OK(cudaMalloc(&D.Sy_f, sizeof(double) * Sy_f.size()));
    CUDA_OK(cudaMalloc(&D.Sy_i, sizeof(double) * Sy_i.size()));
    CUDA_OK(cudaMalloc(&D.Sz_f, sizeof(double) * Sz_f.size()));
    CUDA_OK(cudaMalloc(&D.Sz_i, sizeof(double) * Sz_i.size()));
    CUDA_OK(cudaMalloc(&D.inv, sizeof(double) * block));
    CUDA_OK(cudaMalloc(&D.R, sizeof(double) * block));
    CUDA_OK(cudaMalloc(&D.T, sizeof(double) * block));
    CUDA_OK(cudaMalloc(&D.interior_fixed, n));
    CUDA_OK(cudaMemcpyAsync(D.Sy_f, Sy_f.data(), sizeof(double) * Sy_f.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.Sy_i, Sy_i.data(), sizeof(double) * Sy_i.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.Sz_f, Sz_f.data(), sizeof(double) * Sz_f.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.Sz_i, Sz_i.data(), sizeof(double) * Sz_i.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.interior_fixed, interior_fixed.data(), n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemsetAsync(D.R, 0, sizeof(double) * block, c->stream));
    CUDA_OK(cudaMemsetAsync(D.T, 0, sizeof(double) * block, c->stream));
    CUDA_OK(cudaFuncSetAttribute(k_gemm3, cudaFuncAttributeMaxDynamicSharedMemorySize, G3_SMEM));
    CUDA_OK(cudaFuncSetAttribute(k_gemm3_mma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, M3Cfg<32>::SMEM));
    const int plane = D.ldj * D.ldk;
    k_thomas_setup<<<(plane + 127) / 128, 128, 0, c->stream>>>(D.n_i, D.n_j, D.n_k, D.ldj, D.ldk, D.inv);
    c->launches++;
    CUDA_OK(cudaStreamSynchronize(c->stream));
    D.ok = true;
    if (D.ne > 0)
    {
        // Green's functions of the electrode nodes: unit source at node e, zero frame
        CUDA_OK(cudaMalloc(&D.green, sizeof(double) * n * D.ne));
        CUDA_OK(cudaMemsetAsync(D.green, 0, sizeof(double) * n * D.ne, c->stream));
        std::vector<double> cmat((size_t)D.ne * D.ne);
        for (int e = 0; e < D.ne; e++)
        {
            const int m = e_nodes[e];
            const int k = m % N, j = (m / N) % K, i = m / (K * N);
            CUDA_OK(cudaMemsetAsync(D.T, 0, sizeof(double) * block, c->stream));
            const double one = 1.0;
            CUDA_OK(cudaMemcpyAsync(D.T + ((size_t)(i - 1) * D.ldj + (j - 1)) * D.ldk + (k - 1), &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            if (solve_interior(c, D.green + (size_t)e * n, true)) return 1;
            for (int q = 0; q < D.ne; q++)
                CUDA_OK(cudaMemcpyAsync(&cmat[(size_t)q * D.ne + e], D.green + (size_t)e * n + e_nodes[q], sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaStreamSynchronize(c->stream));
        }
        // invert the capacitance matrix (Gauss-Jordan, partial pivoting)
        const int ne = D.ne;
        std::vector<double> a = cmat, inv((size_t)ne * ne, 0.0);
        for (int q = 0; q < ne; q++) inv[(size_t)q * ne + q] = 1.0;
        for (int col = 0; col < ne; col++)
        {
            int piv = col;
            for (int r = col + 1; r < ne; r++)
                if (std::fabs(a[(size_t)r * ne + col]) > std::fabs(a[(size_t)piv * ne + col])) piv = r;
            if (std::fabs(a[(size_t)piv * ne + col]) < 1e-300) { mag2d_set_error("3-D solver: singular capacitance matrix"); return 1; }
            for (int q = 0; q < ne; q++)
            {
                std::swap(a[(size_t)piv * ne + q], a[(size_t)col * ne + q]);
                std::swap(inv[(size_t)piv * ne + q], inv[(size_t)col * ne + q]);
            }
            const double d = 1.0 / a[(size_t)col * ne + col];
            for (int q = 0; q < ne; q++) { a[(size_t)col * ne + q] *= d; inv[(size_t)col * ne + q] *= d; }
            for (int r = 0; r < ne; r++)
            {
                if (r == col) continue;
                const double f = a[(size_t)r * ne + col];
                if (f == 0.0) continue;
                for (int q = 0; q < ne; q++) { a[(size_t)r * ne + q] -= f * a[(size_t)col * ne + q]; inv[(size_t)r * ne + q] -= f * inv[(size_t)col * ne + q]; }
            }
        }
        CUDA_OK(cudaMalloc(&D.e_nodes, sizeof(int) * ne));
        CUDA_OK(cudaMalloc(&D.e_volts, sizeof(double) * ne));
        CUDA_OK(cudaMalloc(&D.cinv, sizeof(double) * ne * ne));
        CUDA_OK(cudaMalloc(&D.alpha, sizeof(double) * ne));
        CUDA_OK(cudaMemcpyAsync(D.e_nodes, e_nodes.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(D.e_volts, e_volts.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaMemcpyAsync(D.cinv, inv.data(), sizeof(double) * ne * ne, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

static bool slab_solve_active(const mag2d_ctx* c)
{
    // N ranks: the solve itself is shared out from three ranks on (every rank must make this call).  On two ranks the transposes
    // and the all-gather cost more than half a solve saves (C5 at 256^3: 1.16 + 0.20 ms against 1.00 + 0.28 ms for the replicated
    // solve and the all-reduce).  MAG3D_SLAB_SOLVE=1 shares it out on any N > 1, =0 keeps it replicated.
    static const int slab_env = getenv("MAG3D_SLAB_SOLVE") ? atoi(getenv("MAG3D_SLAB_SOLVE")) : -1;
    const bool want = slab_env < 0 ? c->nranks >= 3 : slab_env != 0;
    return c->nccl_comm && c->nranks > 1 && want && comm_has_p2p();
}

// with M divisible by N, rank r owns the planes [r M/N, (r+1) M/N) of the potential (the all-gather path of solve_interior_slab):
// it forms the right-hand side of those planes only, and only their charge has to be summed over the ranks
bool solve3d_reads_own_planes_only(const mag2d_ctx* c) { return slab_solve_active(c) && c->g.M % c->nranks == 0; }

// ElMag3D::solve / Solver::solve (src/fields3d.cpp:83-95, 170-181): rho of every species -> b -> u
int solve3d(mag2d_ctx* c, double* resid_out)
{
    Direct3D& D = c->direct3;
    if (!D.ok) { mag2d_set_error("mag2d_solve: mag2d_set_grid has not been called (3-D)"); return 1; }
    const int M = c->g.M, K = c->g.K, N = c->g.N;
    const size_t n = (size_t)M * K * N;
    Rhs3Args A;
    A.M = M; A.K = K; A.N = N;
    A.n_i = D.n_i; A.n_j = D.n_j; A.n_k = D.n_k; A.hpj = D.hpj; A.hpk = D.hpk;
    A.n_species = (int)c->sp.size();
    A.n_charged = 0;
    for (int s = 0; s < (int)c->sp.size(); s++)
        if (c->sp[s].desc.charge != 0.0)
        {
            if (A.n_charged == RHS_MAX_CHARGED) { mag2d_set_error("3-D solver: more than 8 charged species"); return 1; }
            A.charged[A.n_charged] = s;
            A.charge[A.n_charged++] = c->sp[s].desc.charge;
        }
    A.factor = -c->g.macroparticle_factor / MAG2D_EPS0;
    A.mask = c->d_mask;
    A.interior_fixed = D.interior_fixed;
    A.voltage = c->d_voltage;
    A.rho = c->d_rho;
    A.ne = D.ne;
    A.b = resid_out ? c->d_b : nullptr;
    A.u = c->d_u;
    A.R = D.R;
    // the residual check needs the right-hand side of every plane (and the complete charge grids: mag2d_step's last step and
    // mag2d_advance_init all-reduce)
    const bool own_only = solve3d_reads_own_planes_only(c) && !resid_out;
    const int planes = own_only ? M / c->nranks : M;
    A.i_first = own_only ? c->rank * planes : 0;
    k_rhs3d<<<dim3((unsigned)((K + 1) / 2), (unsigned)planes), 128, 0, c->stream>>>(A);
    c->launches++;
    if (slab_solve_active(c))
    {
        if (solve_interior_slab(c, c->d_u)) return 1;
    }
    else if (solve_interior(c, c->d_u, false)) return 1;
    if (D.ne > 0)
    {
        k_capacitance<<<1, 64, 0, c->stream>>>(D.ne, D.e_nodes, D.e_volts, D.cinv, c->d_u, D.alpha);
        k_add_green<<<148 * 8, 256, 0, c->stream>>>(n, D.ne, D.alpha, D.green, c->d_u);
        c->launches += 2;
    }
    CUDA_OK(cudaGetLastError());
    if (resid_out)
    {
        CUDA_OK(cudaMemsetAsync(c->d_scratch, 0, 2 * sizeof(double), c->stream));
        k_residual3d<<<148 * 8, 256, 0, c->stream>>>(M, K, N, c->d_mask, c->d_u, c->d_b, c->d_scratch);
        c->launches++;
        double h[2];
        CUDA_OK(cudaMemcpyAsync(h, c->d_scratch, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        *resid_out = h[1] > 0 ? h[0] / h[1] : h[0];
    }
    return 0;
}
