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
// which is exact up to round-off like the reference's LU, costs three kernels, and needs no convergence test.
// Grids with internal electrodes keep the multigrid of poisson.cu.
#include <cmath>
#include <vector>

#include "ctx.hpp"

namespace {

constexpr int GM_TM = 64, GM_TN = 32, GM_TK = 16, GM_THREADS = 128;   // CTA tile 64x32, 4x4 per thread

struct DirectArgs
{
    int M, N, n;              // n = N - 2 interior columns
    const double* b;          // [M][N] right-hand side in the reference's scaling (k_rhs); electrode nodes hold their voltage
    const double* S;          // [n][n] sine matrix
    const unsigned char* rowfree;   // [M]
    const double* k2;         // [M] z-coupling of row i (0 on electrode rows)
    double* hat;              // [M][n] transformed rows
    double* u;                // [M][N] potential
    double scale;             // 2/(n+1)
};

// element (i, jj) of B': the right-hand side of interior column j = jj+1 with the two Dirichlet end nodes moved over
__device__ __forceinline__ double bprime(const DirectArgs& A, int i, int jj)
{
    const double* row = A.b + (size_t)i * A.N;
    double v = row[jj + 1];
    if (A.rowfree[i])
    {
        if (jj == 0) v -= A.k2[i] * row[0];
        if (jj == A.n - 1) v -= A.k2[i] * row[A.N - 1];
    }
    return v;
}

// C = A x S with A either B' (FORWARD) or hat (inverse; result scaled and scattered into u's interior columns)
template <bool FORWARD>
__global__ void __launch_bounds__(GM_THREADS) k_direct_gemm(const __grid_constant__ DirectArgs A)
{
    __shared__ __align__(32) double sa[2][GM_TK][GM_TM + 4];   // A tile stored k-major so that a thread's 4 rows are contiguous
    __shared__ __align__(32) double sb[2][GM_TK][GM_TN];
    const int n = A.n, M = A.M;
    const int i0 = blockIdx.y * GM_TM, j0 = blockIdx.x * GM_TN;
    const int t = threadIdx.x;
    const int tr = (t / 8) * 4, tc = (t % 8) * 4;     // 16 x 8 threads, 4 x 4 outputs each
    double acc[4][4] = {};
    auto load = [&](int buf, int k0) {
        // A tile: 64 rows x 16 k, 1024 elements, 8 per thread; consecutive threads read consecutive k (coalesced rows)
        for (int e = t; e < GM_TM * GM_TK; e += GM_THREADS)
        {
            const int r = e / GM_TK, k = e % GM_TK;
            const int i = i0 + r, kk = k0 + k;
            double v = 0.0;
            if (i < M && kk < n) v = FORWARD ? bprime(A, i, kk) : A.hat[(size_t)i * n + kk];
            sa[buf][k][r] = v;
        }
        for (int e = t; e < GM_TK * GM_TN; e += GM_THREADS)
        {
            const int k = e / GM_TN, cidx = e % GM_TN;
            const int kk = k0 + k, j = j0 + cidx;
            sb[buf][k][cidx] = (kk < n && j < n) ? A.S[(size_t)kk * n + j] : 0.0;
        }
    };
    const int nk = (n + GM_TK - 1) / GM_TK;
    load(0, 0);
    __syncthreads();
    for (int kb = 0; kb < nk; kb++)
    {
        const int buf = kb & 1;
        if (kb + 1 < nk) load(buf ^ 1, (kb + 1) * GM_TK);
#pragma unroll
        for (int k = 0; k < GM_TK; k++)
        {
            const double4 a4 = *reinterpret_cast<const double4*>(&sa[buf][k][tr]);
            const double4 b4 = *reinterpret_cast<const double4*>(&sb[buf][k][tc]);
            const double a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int p = 0; p < 4; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[p][q] = fma(a[p], b[q], acc[p][q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 4; p++)
    {
        const int i = i0 + tr + p;
        if (i >= M) continue;
        if (!FORWARD && !A.rowfree[i]) continue;      // electrode rows keep the voltages k_rhs wrote
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const int j = j0 + tc + q;
            if (j >= n) continue;
            if (FORWARD) A.hat[(size_t)i * n + j] = acc[p][q];
            else A.u[(size_t)i * A.N + j + 1] = acc[p][q] * A.scale;
        }
    }
}

// Thomas sweeps, one thread per sine mode; lower[i][k] = a_i/den, inv = 1/den, upper = c_i/den precomputed on the host
__global__ void k_direct_tridiag(int M, int n, const double* __restrict__ lower, const double* __restrict__ inv,
                                 const double* __restrict__ upper, double* __restrict__ hat)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double y = 0.0;
    constexpr int U = 8;
    int i = 0;
    for (; i + U <= M; i += U)
    {
        double bi[U], li[U];
#pragma unroll
        for (int q = 0; q < U; q++)
        {
            const size_t e = (size_t)(i + q) * n + k;
            bi[q] = hat[e] * inv[e];
            li[q] = lower[e];
        }
#pragma unroll
        for (int q = 0; q < U; q++)
        {
            y = fma(-li[q], y, bi[q]);
            hat[(size_t)(i + q) * n + k] = y;
        }
    }
    for (; i < M; i++)
    {
        const size_t e = (size_t)i * n + k;
        y = fma(-lower[e], y, hat[e] * inv[e]);
        hat[e] = y;
    }
    double x = 0.0;
    i = M - 1;
    for (; i - U + 1 >= 0; i -= U)
    {
        double yi[U], ui[U];
#pragma unroll
        for (int q = 0; q < U; q++)
        {
            const size_t e = (size_t)(i - q) * n + k;
            yi[q] = hat[e];
            ui[q] = upper[e];
        }
#pragma unroll
        for (int q = 0; q < U; q++)
        {
            x = fma(-ui[q], x, yi[q]);
            hat[(size_t)(i - q) * n + k] = x;
        }
    }
    for (; i >= 0; i--)
    {
        const size_t e = (size_t)i * n + k;
        x = fma(-upper[e], x, hat[e]);
        hat[e] = x;
    }
}

}  // namespace

void direct_free(mag2d_ctx* c)
{
    DirectSolver& D = c->direct;
    cudaFree(D.S);
    cudaFree(D.lower);
    cudaFree(D.inv);
    cudaFree(D.upper);
    cudaFree(D.hat);
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
    std::vector<double> S((size_t)n * n), lower((size_t)M * n), inv((size_t)M * n), upper((size_t)M * n);
    for (int j = 0; j < n; j++)
        for (int k = j; k < n; k++)
        {
            // reduce the argument exactly before calling sin: (j+1)(k+1) mod 2(n+1)
            const long long p = (long long)(j + 1) * (k + 1) % (2LL * (n + 1));
            const double s = (double)sinl(M_PIl * (long double)p / (long double)(n + 1));
            S[(size_t)j * n + k] = S[(size_t)k * n + j] = s;
        }
    for (int k = 0; k < n; k++)
    {
        const double lam = (double)(2.0L * cosl(M_PIl * (long double)(k + 1) / (long double)(n + 1)));
        double cp_prev = 0.0;
        for (int i = 0; i < M; i++)
        {
            const double d = rowfree[i] ? C[i] + k2[i] * lam : 1.0;
            const double den = d - W[i] * cp_prev;
            if (!(std::fabs(den) > 1e-300)) return 0;
            const size_t e = (size_t)i * n + k;
            inv[e] = 1.0 / den;
            lower[e] = W[i] / den;
            upper[e] = E[i] / den;
            cp_prev = upper[e];
        }
    }
    DirectSolver& D = c->direct;
    D.n = n;
    CUDA_OK(cudaMalloc(&D.S, sizeof(double) * S.size()));
    CUDA_OK(cudaMalloc(&D.lower, sizeof(double) * lower.size()));
    CUDA_OK(cudaMalloc(&D.inv, sizeof(double) * inv.size()));
    CUDA_OK(cudaMalloc(&D.upper, sizeof(double) * upper.size()));
    CUDA_OK(cudaMalloc(&D.hat, sizeof(double) * lower.size()));
    CUDA_OK(cudaMalloc(&D.rowfree, M));
    CUDA_OK(cudaMalloc(&D.k2, sizeof(double) * M));
    CUDA_OK(cudaMemcpyAsync(D.S, S.data(), sizeof(double) * S.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.lower, lower.data(), sizeof(double) * lower.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.inv, inv.data(), sizeof(double) * inv.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.upper, upper.data(), sizeof(double) * upper.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.rowfree, rowfree.data(), M, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(D.k2, k2.data(), sizeof(double) * M, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    D.ok = true;
    return 0;
}

// u (or uRF) <- solution of the system whose right-hand side k_rhs left in c->d_b
int direct_solve(mag2d_ctx* c, double* u)
{
    const DirectSolver& D = c->direct;
    DirectArgs A;
    A.M = c->g.M;
    A.N = c->g.N;
    A.n = D.n;
    A.b = c->d_b;
    A.S = D.S;
    A.rowfree = D.rowfree;
    A.k2 = D.k2;
    A.hat = D.hat;
    A.u = u;
    A.scale = 2.0 / (D.n + 1);
    const dim3 grid((D.n + GM_TN - 1) / GM_TN, (A.M + GM_TM - 1) / GM_TM);
    k_direct_gemm<true><<<grid, GM_THREADS, 0, c->stream>>>(A);
    k_direct_tridiag<<<(D.n + 31) / 32, 32, 0, c->stream>>>(A.M, D.n, D.lower, D.inv, D.upper, D.hat);
    k_direct_gemm<false><<<grid, GM_THREADS, 0, c->stream>>>(A);
    c->launches += 3;
    CUDA_OK(cudaGetLastError());
    return 0;
}
