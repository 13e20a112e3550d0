// poisson_direct.cu — direct Poisson solve for grids whose electrodes are whole rows.
//
// The reference factorises the five-point operator with UMFPACK (src/fields.cpp:169-263 builds it,
// fields.cpp:278-329 solves every step).  When every grid row i is either an electrode over its whole length or
// free between two Dirichlet end nodes (j = 0 and j = N-1) — geometry EMPTY in both coordinate systems, the
// self-consistent discharge decks — the operator separates: the z-direction part is the constant-coefficient
// second difference, whose eigenvectors are the sine modes sin(pi j k / (N-1)).  The solve is then
//      B^ = B' S            (sine transform of every row, one FP64 matrix product with the symmetric sine matrix S)
//      T_k u^_k = b^_k      (one tridiagonal system in x / r per mode k; Thomas factors precomputed at set_grid)
//      U  = (2/(N-1)) U^ S  (inverse transform)
// which is exact up to round-off like the reference's LU and needs no convergence test.
// The transforms are FOLDED, as in poisson3d.cu: S[j][n-1-k] = (-1)^j S[j][k], so modes with even j only see the symmetric
// part of a row and modes with odd j the antisymmetric part.  k_rhs writes every row as [symmetric half | antisymmetric
// half], the spectrum is kept in the same split order (the Thomas factors are stored by that order), each transform is two
// products with n/2 x n/2 matrices (half the flops, and twice the CTAs per unit of work: these small products are bound by
// the latency of a CTA, not by the FP64 pipe), and k_direct_unfold writes x[k] = e + o, x[n-1-k] = e - o into the potential.
// Grids with internal electrodes keep the multigrid of poisson.cu.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace {

// FP64 matrix product C = A x S on the CUDA cores: CTA tile 64 x 32, GM_RPT x 4 outputs per thread, K in chunks of 32
// streamed through a 3-stage cp.async ring (the operands are L2-resident, the ring hides the L2 latency).
// All operands are padded to ld = a multiple of 32 columns with zeros, so the K loop has no bounds tests.
// Measured on C4 (one CTA per SM at 512^2): 128 threads with 4 x 4 outputs each give a solve of 0.084 ms per step, 256 threads with
// 2 x 4 (two warps per scheduler, -DMAG2D_GM_RPT=2) 0.096 ms: the shared-memory loads per FMA count, not the warps per scheduler.
#ifndef MAG2D_GM_RPT
#define MAG2D_GM_RPT 4
#endif
constexpr int GM_RPT = MAG2D_GM_RPT;                                  // output rows per thread
constexpr int GM_TM = 64, GM_TN = 32, GM_KC = 32, GM_THREADS = 8 * GM_TM / GM_RPT, GM_STAGES = 3;
constexpr int GM_LDA = GM_KC + 2;                                    // smem row stride of the A tile (doubles)
constexpr int GM_STAGE_DOUBLES = GM_TM * GM_LDA + GM_KC * GM_TN;
constexpr int GM_SMEM = GM_STAGES * GM_STAGE_DOUBLES * (int)sizeof(double);

struct GemmArgs
{
    int M, ld, hp;                  // rows, leading dimension 2 hp of A / hat, half length (K and column count of one product)
    const double* A;                // [M][ld]: parity p = blockIdx.z works on the columns [p hp, (p + 1) hp)
    const double* S;                // [2][hp][hp] half-size sine matrices of the two parities, zero padded
    const double* inv;              // FORWARD: [M][ld] 1/den of the Thomas factorisation, folded into the epilogue
    double* C;                      // FORWARD: value slots of the pair array [M][ld][2];  INVERSE: [M][ld]
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool FORWARD>
__global__ void __launch_bounds__(GM_THREADS) k_direct_gemm(const __grid_constant__ GemmArgs G)
{
    extern __shared__ __align__(16) double gm_smem[];
    const int t = threadIdx.x;
    const int i0 = blockIdx.y * GM_TM, j0 = blockIdx.x * GM_TN;
    const int tr = (t / 8) * GM_RPT, tc = (t % 8) * 4;
    const int nk = G.hp / GM_KC;
    const int par = blockIdx.z;
    const double* Ap = G.A + par * G.hp;
    const double* Sp = G.S + (size_t)par * G.hp * G.hp;
    auto issue = [&](int kb) {
        double* sa = gm_smem + (kb % GM_STAGES) * GM_STAGE_DOUBLES;
        double* sb = sa + GM_TM * GM_LDA;
        const int k0 = kb * GM_KC;
        // A tile: 64 rows x 32 doubles = 1024 16-byte pieces (16 pieces per row)
#pragma unroll
        for (int q = 0; q < 1024 / GM_THREADS; q++)
        {
            const int e = t + q * GM_THREADS, r = e >> 4, c2 = (e & 15) * 2;
            const int i = min(i0 + r, G.M - 1);
            cp_async16(sa + r * GM_LDA + c2, Ap + (size_t)i * G.ld + k0 + c2);
        }
        // S tile: 32 k x 32 columns = 512 pieces
#pragma unroll
        for (int q = 0; q < 512 / GM_THREADS; q++)
        {
            const int e = t + q * GM_THREADS, r = e >> 4, c2 = (e & 15) * 2;
            cp_async16(sb + r * GM_TN + c2, Sp + (size_t)(k0 + r) * G.hp + j0 + c2);
        }
    };
    double acc[GM_RPT][4] = {};
    for (int s = 0; s < GM_STAGES - 1; s++)
    {
        if (s < nk) issue(s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; kb++)
    {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();                               // tile kb landed for everyone; tile kb-1's buffer is free
        if (kb + GM_STAGES - 1 < nk) issue(kb + GM_STAGES - 1);
        cp_async_commit();
        const double* sa = gm_smem + (kb % GM_STAGES) * GM_STAGE_DOUBLES;
        const double* sb = sa + GM_TM * GM_LDA;
#pragma unroll 8
        for (int k = 0; k < GM_KC; k++)
        {
            double a[GM_RPT], b[4];
#pragma unroll
            for (int p = 0; p < GM_RPT; p++) a[p] = sa[(tr + p) * GM_LDA + k];
            const double2 b01 = *reinterpret_cast<const double2*>(sb + k * GM_TN + tc);
            const double2 b23 = *reinterpret_cast<const double2*>(sb + k * GM_TN + tc + 2);
            b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
#pragma unroll
            for (int p = 0; p < GM_RPT; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
    }
#pragma unroll
    for (int p = 0; p < GM_RPT; p++)
    {
        const int i = i0 + tr + p;
        if (i >= G.M) continue;
        const size_t e = (size_t)i * G.ld + par * G.hp + j0 + tc;
        if (FORWARD)
        {
            // p = hat / den goes into the value slots of the forward pair array [i][k][2]
            const double2 i01 = *reinterpret_cast<const double2*>(G.inv + e), i23 = *reinterpret_cast<const double2*>(G.inv + e + 2);
            G.C[2 * e] = acc[p][0] * i01.x;
            G.C[2 * e + 2] = acc[p][1] * i01.y;
            G.C[2 * e + 4] = acc[p][2] * i23.x;
            G.C[2 * e + 6] = acc[p][3] * i23.y;
        }
        else
        {
            *reinterpret_cast<double2*>(G.C + e) = make_double2(acc[p][0], acc[p][1]);
            *reinterpret_cast<double2*>(G.C + e + 2) = make_double2(acc[p][2], acc[p][3]);
        }
    }
}

// ---- the same product on the FP64 tensor pipe (mma.sync m8n8k4, DMMA; poisson3d.cu has the large-grid version) ------------
// CTA tile 64 x 32, four warps stacked along the rows (16 x 32 per warp = 2 x 4 DMMA tiles), K in chunks of 16 through a 3-stage
// cp.async ring.  The FMA kernel above spends ~100 cycles per k step on its 16 dependent-load FMAs per thread (one CTA per SM, nothing
// to hide the shared-memory latency with); here a warp issues 8 DMMAs per 6 fragment loads.  Row strides of 20 / 36 doubles keep the
// fragment loads free of bank conflicts.
constexpr int DM_KC = 16, DM_STAGES = 3, DM_LDA = DM_KC + 4, DM_LDB = GM_TN + 4;
constexpr int DM_STAGE_DOUBLES = GM_TM * DM_LDA + DM_KC * DM_LDB;
constexpr int DM_SMEM = DM_STAGES * DM_STAGE_DOUBLES * (int)sizeof(double);

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <bool FORWARD>
__global__ void __launch_bounds__(128) k_direct_gemm_mma(const __grid_constant__ GemmArgs G)
{
    extern __shared__ __align__(16) double dm_smem[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i0 = blockIdx.y * GM_TM, j0 = blockIdx.x * GM_TN;
    const int nk = G.hp / DM_KC;
    const int par = blockIdx.z;
    const double* Ap = G.A + par * G.hp;
    const double* Sp = G.S + (size_t)par * G.hp * G.hp;
    auto issue = [&](int kb) {
        double* sa = dm_smem + (kb % DM_STAGES) * DM_STAGE_DOUBLES;
        double* sb = sa + GM_TM * DM_LDA;
        const int k0 = kb * DM_KC;
        // A tile: 64 rows x 16 doubles = 512 16-byte pieces (8 per row), 4 per thread
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const int e = t + q * 128, r = e >> 3, c2 = (e & 7) * 2;
            const int i = min(i0 + r, G.M - 1);
            cp_async16(sa + r * DM_LDA + c2, Ap + (size_t)i * G.ld + k0 + c2);
        }
        // S tile: 16 k x 32 columns = 256 pieces, 2 per thread
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            const int e = t + q * 128, r = e >> 4, c2 = (e & 15) * 2;
            cp_async16(sb + r * DM_LDB + c2, Sp + (size_t)(k0 + r) * G.hp + j0 + c2);
        }
    };
    double acc[2][4][2] = {};
    for (int s = 0; s < DM_STAGES - 1; s++)
    {
        if (s < nk) issue(s);
        cp_async_commit();
    }
    const int fr = lane >> 2, fc = lane & 3;     // fragment row / column of this lane
    for (int kb = 0; kb < nk; kb++)
    {
        cp_async_wait<DM_STAGES - 2>();
        __syncthreads();
        if (kb + DM_STAGES - 1 < nk) issue(kb + DM_STAGES - 1);
        cp_async_commit();
        const double* sa = dm_smem + (kb % DM_STAGES) * DM_STAGE_DOUBLES + (warp * 16 + fr) * DM_LDA + fc;
        const double* sb = dm_smem + (kb % DM_STAGES) * DM_STAGE_DOUBLES + GM_TM * DM_LDA + fc * DM_LDB + fr;
#pragma unroll
        for (int k4 = 0; k4 < DM_KC / 4; k4++)
        {
            double a[2], b[4];
#pragma unroll
            for (int p = 0; p < 2; p++) a[p] = sa[p * 8 * DM_LDA + k4 * 4];            // A[rb*8 + fr][k4*4 + fc]
#pragma unroll
            for (int q = 0; q < 4; q++) b[q] = sb[k4 * 4 * DM_LDB + q * 8];            // S[k4*4 + fc][cb*8 + fr]
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) dmma884(acc[p][q], a[p], b[q]);
        }
    }
    // D fragment: row fr, columns 2 fc, 2 fc + 1 of every 8 x 8 block
#pragma unroll
    for (int p = 0; p < 2; p++)
    {
        const int i = i0 + warp * 16 + p * 8 + fr;
        if (i >= G.M) continue;
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const size_t e = (size_t)i * G.ld + par * G.hp + j0 + q * 8 + 2 * fc;
            if (FORWARD)
            {
                // p = hat / den goes into the value slots of the forward pair array [i][k][2]
                const double2 iv = *reinterpret_cast<const double2*>(G.inv + e);
                G.C[2 * e] = acc[p][q][0] * iv.x;
                G.C[2 * e + 2] = acc[p][q][1] * iv.y;
            }
            else
                *reinterpret_cast<double2*>(G.C + e) = make_double2(acc[p][q][0], acc[p][q][1]);
        }
    }
}

// the two halves [e | o] of the inverse products -> the interior columns of the potential: x[k] = e + o, x[n-1-k] = e - o.
// Electrode rows keep the voltages k_rhs wrote.
__global__ void k_direct_unfold(int M, int n, int hp, const double* __restrict__ eo, const unsigned char* __restrict__ rowfree, double scale,
                                double* __restrict__ u, int ldc)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    const int km = n - 1 - kk;
    if (i >= M || kk > km || !rowfree[i]) return;
    const double e = eo[(size_t)i * (2 * hp) + kk], o = eo[(size_t)i * (2 * hp) + hp + kk];
    u[(size_t)i * ldc + kk] = (e + o) * scale;
    if (kk < km) u[(size_t)i * ldc + km] = (e - o) * scale;
}

// Thomas sweeps, one sine mode per lane of warp 0: y_i = p_i - lower_i y_(i-1) with p = hat/den already formed by the
// forward product's epilogue, then x_i = y_i - upper_i x_(i+1).  The recurrence is one DFMA per row, so the sweep is
// bound by the issue latency of the single warp that owns a mode: it must execute nothing but LDS.128 / DFMA / STG.
// Warps 1..3 are the producers: they stream the (value, coefficient) pairs, interleaved in memory as [row][mode][2],
// through a cp.async ring of TD_STAGES chunks of TD_ROWS rows (512 contiguous bytes per row).
constexpr int TD_ROWS = 32, TD_STAGES = 4, TD_THREADS = 128;
constexpr int TD_STAGE_DOUBLES = TD_ROWS * 64;
constexpr int TD_SMEM = TD_STAGES * TD_STAGE_DOUBLES * (int)sizeof(double);

// pair: [M][ld][2] (value, coefficient); BACKWARD walks the rows downwards.  OUT_STRIDE 2 writes the result into the
// value slots of the other pair array (forward sweep -> y), 1 into the plain [M][ld] array (backward sweep -> x)
template <bool BACKWARD, int OUT_STRIDE>
__device__ __forceinline__ void tridiag_sweep(int M, int ld, int k0, const double* __restrict__ pair, double* __restrict__ out, double* sm)
{
    const int t = threadIdx.x, lane = t & 31;
    const int nchunks = (M + TD_ROWS - 1) / TD_ROWS;
    auto issue = [&](int chunk) {
        double* st = sm + (chunk % TD_STAGES) * TD_STAGE_DOUBLES;
        // 32 rows x 32 pieces of 16 bytes, spread over the 96 producer threads
        for (int e = t - 32; e < TD_ROWS * 32; e += TD_THREADS - 32)
        {
            const int r = e >> 5, c16 = e & 31;
            const int q = chunk * TD_ROWS + r;
            if (q >= M) break;
            const int i = BACKWARD ? M - 1 - q : q;
            cp_async16(st + r * 64 + c16 * 2, pair + ((size_t)i * ld + k0 + c16) * 2);
        }
    };
    if (t >= 32)
        for (int s = 0; s < TD_STAGES - 1; s++)
        {
            if (s < nchunks) issue(s);
            cp_async_commit();
        }
    double y = 0.0;
    for (int chunk = 0; chunk < nchunks; chunk++)
    {
        if (t >= 32) cp_async_wait<TD_STAGES - 2>();
        __syncthreads();            // chunk landed; everybody is done with chunk - 1, whose buffer is refilled below
        if (t >= 32)
        {
            if (chunk + TD_STAGES - 1 < nchunks) issue(chunk + TD_STAGES - 1);
            cp_async_commit();
        }
        else
        {
            const double2* st = reinterpret_cast<const double2*>(sm + (chunk % TD_STAGES) * TD_STAGE_DOUBLES) + lane;
            const int q0 = chunk * TD_ROWS;
            const int rows = min(TD_ROWS, M - q0);
            const long long step = (BACKWARD ? -(long long)ld : (long long)ld) * OUT_STRIDE;
            double* o = out + ((size_t)(BACKWARD ? M - 1 - q0 : q0) * ld + k0 + lane) * OUT_STRIDE;
            if (rows == TD_ROWS)
            {
#pragma unroll
                for (int r = 0; r < TD_ROWS; r++)
                {
                    const double2 pc = st[r * 32];
                    y = fma(-pc.y, y, pc.x);
                    *o = y;
                    o += step;
                }
            }
            else
                for (int r = 0; r < rows; r++)
                {
                    const double2 pc = st[r * 32];
                    y = fma(-pc.y, y, pc.x);
                    *o = y;
                    o += step;
                }
        }
    }
    if (t >= 32) cp_async_wait<0>();
    __syncthreads();
}

__global__ void __launch_bounds__(TD_THREADS) k_direct_tridiag(int M, int ld, const double* __restrict__ fwd, double* __restrict__ bwd,
                                                               double* __restrict__ x)
{
    extern __shared__ __align__(16) double td_smem[];
    const int k0 = blockIdx.x * 32;
    tridiag_sweep<false, 2>(M, ld, k0, fwd, bwd, td_smem);
    __threadfence();          // the backward sweep re-reads the y values through the async copy path
    __syncthreads();
    tridiag_sweep<true, 1>(M, ld, k0, bwd, x, td_smem);
}

}  // namespace

void direct_free(mag2d_ctx* c)
{
    DirectSolver& D = c->direct;
    cudaFree(D.S);
    cudaFree(D.fwd);
    cudaFree(D.inv);
    cudaFree(D.bwd);
    cudaFree(D.hat);
    cudaFree(D.bp);
    cudaFree(D.rowfree);
    cudaFree(D.k2);
    D = DirectSolver();
}

// decide whether the grid separates and, if so, precompute the sine matrix and the Thomas factors of every mode
int direct_setup(mag2d_ctx* c)
{
    direct_free(c);
    const mag2d_grid_desc& g = c->g;
    const int M = g.M, N = g.N, n = N - 2;
    if (n < 1 || M < 2) return 0;
    std::vector<unsigned char> rowfree(M, 0);
    for (int i = 0; i < M; i++)
    {
        const unsigned char* m = &c->h_mask[(size_t)i * N];
        int nfree = 0;
        for (int j = 0; j < N; j++) nfree += m[j] != MAG2D_FIXED && m[j] != MAG2D_FIXED_RF;
        if (nfree == 0) continue;
        const bool ends_fixed = (m[0] == MAG2D_FIXED || m[0] == MAG2D_FIXED_RF) && (m[N - 1] == MAG2D_FIXED || m[N - 1] == MAG2D_FIXED_RF);
        if (nfree != n || !ends_fixed) return 0;       // an electrode inside a free row: not separable
        rowfree[i] = 1;
    }
    const bool cyl = g.coord == MAG2D_CYLINDRICAL;
    // unscaled five-point rows exactly as the reference assembles them (fields.cpp:195-262)
    std::vector<double> W(M, 0.0), E(M, 0.0), C(M, 1.0), k2(M, 0.0);
    for (int i = 0; i < M; i++)
    {
        if (!rowfree[i]) continue;
        if (cyl)
        {
            k2[i] = 1.0 / (g.dz * g.dz);
            if (i == 0)
            {
                E[i] = 1.0 / (g.dx * g.dx * 0.25);
                C[i] = -2.0 * k2[i] - E[i];
            }
            else
            {
                W[i] = (i - 0.5) / (g.dx * g.dx * i);
                E[i] = (i + 0.5) / (g.dx * g.dx * i);
                C[i] = -2.0 * k2[i] - W[i] - E[i];
            }
        }
        else
        {
            W[i] = E[i] = k2[i] = 1.0;
            C[i] = -4.0;
        }
        if (i == 0) W[i] = 0.0;
        if (i == M - 1) E[i] = 0.0;
    }
    // every [.][ld] array holds two zero-padded halves of hp columns (a multiple of 32): the kernels need no column bounds tests
    const int hp = ((n + 1) / 2 + 31) / 32 * 32, ld = 2 * hp;
    // S: [forward | inverse][parity][hp][hp].  E_b[kk][m] = sin(pi (2m + b + 1)(kk + 1) / (n + 1)) for the ceil(n/2) (b = 0) or
    // floor(n/2) (b = 1) folded columns kk and modes m; forward products take E_b, inverse products its transpose
    std::vector<double> S(4 * (size_t)hp * hp, 0.0), lower((size_t)M * ld, 0.0), inv((size_t)M * ld, 0.0), upper((size_t)M * ld, 0.0);
    for (int b = 0; b < 2; b++)
    {
        const int h = b ? n / 2 : (n + 1) / 2;
        for (int kk = 0; kk < h; kk++)
            for (int m = 0; m < h; m++)
            {
                // reduce the argument exactly before calling sin: (2m+b+1)(kk+1) mod 2(n+1)
                const long long p = (long long)(2 * m + b + 1) * (kk + 1) % (2LL * (n + 1));
                const double s = (double)sinl(M_PIl * (long double)p / (long double)(n + 1));
                S[((size_t)b * hp + kk) * hp + m] = s;
                S[((size_t)(2 + b) * hp + m) * hp + kk] = s;
            }
    }
    for (int k = 0; k < ld; k++)
    {
        // column k of the folded spectrum is mode 2k (k < hp) or 2(k - hp) + 1, counted from 0; beyond n: padding (its
        // right-hand side is zero and stays zero; the factors below are those of a well-posed dummy system)
        const int mode = k < hp ? 2 * k : 2 * (k - hp) + 1;
        const double lam = mode < n ? (double)(2.0L * cosl(M_PIl * (long double)(mode + 1) / (long double)(n + 1))) : 0.0;
        double cp_prev = 0.0;
        for (int i = 0; i < M; i++)
        {
            const double d = rowfree[i] ? C[i] + k2[i] * lam : 1.0;
            const double den = d - W[i] * cp_prev;
            if (!(std::fabs(den) > 1e-300)) return 0;
            const size_t e = (size_t)i * ld + k;
            inv[e] = 1.0 / den;
            lower[e] = W[i] / den;
            upper[e] = E[i] / den;
            cp_prev = upper[e];
        }
    }
    DirectSolver& D = c->direct;
    D.n = n;
    D.hp = hp;
    D.ld = ld;
    CUDA_OK(cudaMalloc(&D.S, sizeof(double) * S.size()));
    // (value, coefficient) pairs of the two sweeps: the coefficient slots are filled once, here
    std::vector<double> fwd(2 * lower.size(), 0.0), bwd(2 * upper.size(), 0.0);
    for (size_t e = 0; e < lower.size(); e++)
    {
        fwd[2 * e + 1] = lower[e];
        bwd[2 * e + 1] = upper[e];
    }
    CUDA_OK(cudaMalloc(&D.fwd, sizeof(double) * fwd.size()));
    CUDA_OK(cudaMalloc(&D.inv, sizeof(double) * inv.size()));
    CUDA_OK(cudaMalloc(&D.bwd, sizeof(double) * bwd.size()));
    CUDA_OK(cudaMalloc(&D.hat, sizeof(double) * lower.size()));
    CUDA_OK(cudaMalloc(&D.bp, sizeof(double) * lower.size()));
    CUDA_OK(cudaMemsetAsync(D.hat, 0, sizeof(double) * lower.size(), c->stream));
    CUDA_OK(cudaMemsetAsync(D.bp, 0, sizeof(double) * lower.size(), c->stream));
    CUDA_OK(cudaFuncSetAttribute(k_direct_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM));
    CUDA_OK(cudaFuncSetAttribute(k_direct_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM));
    CUDA_OK(cudaFuncSetAttribute(k_direct_gemm_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
    CUDA_OK(cudaFuncSetAttribute(k_direct_gemm_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
    CUDA_OK(cudaFuncSetAttribute(k_direct_tridiag, cudaFuncAttributeMaxDynamicSharedMemorySize, TD_SMEM));
    CUDA_OK(cudaMalloc(&D.rowfree, M));
    CUDA_OK(cudaMalloc(&D.k2, sizeof(double) * M));
    CUDA_OK(cudaMemcpyAsync(D.S, S.data(), sizeof(double) * S.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.fwd, fwd.data(), sizeof(double) * fwd.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.inv, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.bwd, bwd.data(), sizeof(double) * bwd.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.rowfree, rowfree.data(), M, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.k2, k2.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    D.ok = true;
    return 0;
}

// u (or uRF) <- solution of the system whose right-hand side k_rhs left in c->direct.bp (B': interior columns, the
// Dirichlet end nodes already moved to the right-hand side; electrode rows hold their voltages)
int direct_solve(mag2d_ctx* c, double* u)
{
    const DirectSolver& D = c->direct;
    GemmArgs G;
    G.M = c->g.M;
    G.ld = D.ld;
    G.hp = D.hp;
    G.inv = D.inv;
    const dim3 grid(D.hp / GM_TN, (G.M + GM_TM - 1) / GM_TM, 2);
    // MAG2D_GEMM=fma: the products on the FP64 FMA pipe instead of the tensor pipe
    static const bool use_fma = getenv("MAG2D_GEMM") && !strcmp(getenv("MAG2D_GEMM"), "fma");
    G.A = D.bp;
    G.S = D.S;
    G.C = D.fwd;
    if (use_fma) k_direct_gemm<true><<<grid, GM_THREADS, GM_SMEM, c->stream>>>(G);
    else k_direct_gemm_mma<true><<<grid, 128, DM_SMEM, c->stream>>>(G);
    k_direct_tridiag<<<D.ld / 32, TD_THREADS, TD_SMEM, c->stream>>>(G.M, D.ld, D.fwd, D.bwd, D.hat);
    G.A = D.hat;
    G.S = D.S + 2 * (size_t)D.hp * D.hp;
    G.C = D.bp;              // the folded right-hand side has been consumed: its array takes the halves [e | o]
    if (use_fma) k_direct_gemm<false><<<grid, GM_THREADS, GM_SMEM, c->stream>>>(G);
    else k_direct_gemm_mma<false><<<grid, 128, DM_SMEM, c->stream>>>(G);
    const dim3 block(32, 8);
    k_direct_unfold<<<dim3((D.hp + 31) / 32, (G.M + 7) / 8), block, 0, c->stream>>>(G.M, D.n, D.hp, D.bp, D.rowfree, 2.0 / (D.n + 1), u + 1, c->g.N);
    c->launches += 4;
    CUDA_OK(cudaGetLastError());
    return 0;
}
