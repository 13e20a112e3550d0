// sort.cu — periodic cell sort + compaction of the device-resident SoA particle store.
//
// Replaces the reference's LIFO free list (BaseSpecies::insert/remove, src/particles.hpp:223-247):
// removed particles only get a NaN marker in x and keep their slot until the next sort, which
// (1) counts particles per cell and hands each one its rank inside the cell, (2) scans the counts,
// (3) permutes all phase-space arrays out of place into the other slab.  Dead slots are dropped
// (stream compaction); afterwards particles of one cell are contiguous, which is what keeps the
// field gather and the charge scatter of push.cu cache- and atomics-friendly.
// Extra HBM traffic: 16 B read + 8 B written (keys) and 48+40 B (permute) per particle, every K steps.
#include "ctx.hpp"

namespace {

constexpr unsigned INVALID_KEY = 0xFFFFFFFFu;
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;   // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__global__ void k_sort_count(const double* __restrict__ x, const double* __restrict__ z, long long n, double idx, double idz,
                             int M, int N, unsigned* __restrict__ count, unsigned* __restrict__ key, unsigned* __restrict__ rank,
                             const double* __restrict__ y, double idy, int K)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double px = k < n ? pld<T>(x, k) : dead_marker();
    const bool valid = particle_alive(px);
    unsigned ky = INVALID_KEY, rk = 0;
    if (valid)
    {
        int i = (int)(px * idx), j = (int)(pld<T>(z, k) * idz);
        i = max(min(i, M - 2), 0);
        j = max(min(j, N - 2), 0);
        ky = (unsigned)i * (unsigned)(N - 1) + (unsigned)j;
        if (y)
        {
            // CARTESIAN3D: cell (i, jy, k) in the Array3D order, x slowest
            int jy = (int)(pld<T>(y, k) * idy);
            jy = max(min(jy, K - 2), 0);
            ky = ((unsigned)i * (unsigned)(K - 1) + (unsigned)jy) * (unsigned)(N - 1) + (unsigned)j;
        }
    }
    // warp-aggregated ticket: the store is already almost sorted, so the lanes of a warp share a few cells;
    // one atomic per (warp, cell) instead of one per particle, ranks handed out in lane order
    const unsigned lane = lane_id();
    unsigned remaining = __ballot_sync(MAG2D_FULL_MASK, valid);
    while (remaining)
    {
        const int src = __ffs(remaining) - 1;
        const unsigned k0 = __shfl_sync(MAG2D_FULL_MASK, ky, src);
        const bool mine = valid && ky == k0;
        const unsigned m = __ballot_sync(MAG2D_FULL_MASK, mine);
        unsigned base = 0;
        if ((int)lane == src) base = atomicAdd(&count[k0], (unsigned)__popc(m));
        base = __shfl_sync(MAG2D_FULL_MASK, base, src);
        if (mine) rk = base + __popc(m & ((1u << lane) - 1u));
        remaining &= ~m;
    }
    if (k < n)
    {
        key[k] = ky;
        rank[k] = rk;
    }
}

// exclusive scan, level 1: per-tile scan + tile sums
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const unsigned* __restrict__ in, unsigned* __restrict__ out, unsigned* __restrict__ tile_sums, int n)
{
    __shared__ unsigned warp_sums[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++)
    {
        v[q] = base + q < n ? in[base + q] : 0u;
        s += v[q];
    }
    // inclusive scan of s across the block
    unsigned incl = s;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        unsigned t = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        unsigned w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            unsigned t = __shfl_up_sync(MAG2D_FULL_MASK, w, o);
            if (lane >= o) w += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    unsigned excl = incl - s + (warp ? warp_sums[warp - 1] : 0u);
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++)
    {
        if (base + q < n) out[base + q] = excl;
        excl += v[q];
    }
    if (threadIdx.x == SCAN_THREADS - 1) tile_sums[blockIdx.x] = excl;
}

// level 2: one block scans the tile sums in place (exclusive) and publishes the grand total
__global__ void __launch_bounds__(1024) k_scan_sums(unsigned* __restrict__ sums, int n, unsigned long long* __restrict__ total)
{
    __shared__ unsigned carry;
    __shared__ unsigned warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024)
    {
        const int k = base + threadIdx.x;
        const unsigned v = k < n ? sums[k] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            unsigned t = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                unsigned t = __shfl_up_sync(MAG2D_FULL_MASK, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned excl = incl - v + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (k < n) sums[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(unsigned* __restrict__ out, const unsigned* __restrict__ tile_sums, int n)
{
    const unsigned add = tile_sums[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++)
        if (base + q < n) out[base + q] += add;
}

struct PermArgs
{
    const double* src[N_ARR];
    double* dst[N_ARR];
    int n_arr;
};

template <typename T>
__global__ void k_sort_scatter(const __grid_constant__ PermArgs P, long long n, const unsigned* __restrict__ key,
                               const unsigned* __restrict__ rank, const unsigned* __restrict__ offset)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned ky = key[k];
    if (ky == INVALID_KEY) return;
    const long long d = (long long)offset[ky] + rank[k];
#pragma unroll
    for (int a = 0; a < N_ARR; a++)
        if (a < P.n_arr) reinterpret_cast<T*>(P.dst[a])[d] = reinterpret_cast<const T*>(P.src[a])[k];
}

// slots behind the compacted particles become dead markers
template <typename T>
__global__ void k_fill_dead(double* __restrict__ x, const unsigned long long* __restrict__ total, long long n)
{
    for (long long k = (long long)*total + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        pst<T>(x, k, dead_marker());
}

// the same for a fused permuting step
template <typename T>
__global__ void k_fill_dead_keys(double* __restrict__ x, unsigned* __restrict__ key, const unsigned long long* __restrict__ total, long long n)
{
    for (long long k = (long long)*total + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    {
        pst<T>(x, k, dead_marker());
        if (key) key[k] = INVALID_KEY;
    }
}

}  // namespace

int launch_sort(mag2d_ctx* c, int s, bool trim)
{
    SpeciesStore& S = c->sp[s];
    const long long n = S.n_slots;
    S.tickets_valid = false;
    S.bins_valid = false;
    if (n == 0) return 0;
    const int M = c->g.M, N = c->g.N;
    const bool three_d = is3d(c);
    const int ncells = (M - 1) * (N - 1) * (three_d ? c->g.K - 1 : 1);
    if (!c->d_cell_count)
    {
        const int ntiles = (ncells + SCAN_TILE - 1) / SCAN_TILE;
        CUDA_OK(cudaMalloc(&c->d_cell_count, sizeof(unsigned) * (size_t)ncells));
        CUDA_OK(cudaMalloc(&c->d_cell_offset, sizeof(unsigned) * (size_t)ncells));
        CUDA_OK(cudaMalloc(&c->d_block_sums, sizeof(unsigned) * (size_t)(ntiles + 2) + 16));
    }
    if (ensure_particle_scratch(c, S.capacity)) return 1;
    const int ntiles = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(c->d_block_sums + ((ntiles + 1) / 2 * 2 + 2));
    double* const* cur = S.arr[S.cur];
    double* const* oth = S.arr[S.cur ^ 1];
    CUDA_OK(cudaMemsetAsync(c->d_cell_count, 0, sizeof(unsigned) * (size_t)ncells, c->stream));
    const unsigned pblocks = (unsigned)((n + 255) / 256);
    if (c->store_f32)
        k_sort_count<float><<<pblocks, 256, 0, c->stream>>>(cur[ARR_X], cur[ARR_Z], n, c->g.idx, c->g.idz, M, N, c->d_cell_count, c->d_key, c->d_rank,
                                                            three_d ? cur[ARR_Y] : nullptr, c->g.idy, c->g.K);
    else
        k_sort_count<double><<<pblocks, 256, 0, c->stream>>>(cur[ARR_X], cur[ARR_Z], n, c->g.idx, c->g.idz, M, N, c->d_cell_count, c->d_key, c->d_rank,
                                                             three_d ? cur[ARR_Y] : nullptr, c->g.idy, c->g.K);
    k_scan_tiles<<<ntiles, SCAN_THREADS, 0, c->stream>>>(c->d_cell_count, c->d_cell_offset, c->d_block_sums, ncells);
    k_scan_sums<<<1, 1024, 0, c->stream>>>(c->d_block_sums, ntiles, d_total);
    k_scan_add<<<ntiles, SCAN_THREADS, 0, c->stream>>>(c->d_cell_offset, c->d_block_sums, ncells);
    PermArgs P;
    P.n_arr = 0;
    for (int a = 0; a < N_ARR; a++)
        if (cur[a])
        {
            P.src[P.n_arr] = cur[a];
            P.dst[P.n_arr] = oth[a];
            P.n_arr++;
        }
    for (int a = P.n_arr; a < N_ARR; a++) { P.src[a] = nullptr; P.dst[a] = nullptr; }
    if (c->store_f32)
    {
        k_sort_scatter<float><<<pblocks, 256, 0, c->stream>>>(P, n, c->d_key, c->d_rank, c->d_cell_offset);
        k_fill_dead<float><<<148 * 4, 256, 0, c->stream>>>(oth[ARR_X], d_total, n);
    }
    else
    {
        k_sort_scatter<double><<<pblocks, 256, 0, c->stream>>>(P, n, c->d_key, c->d_rank, c->d_cell_offset);
        k_fill_dead<double><<<148 * 4, 256, 0, c->stream>>>(oth[ARR_X], d_total, n);
    }
    c->launches += 6;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemsetAsync(S.d_removed, 0, sizeof(unsigned long long), c->stream));
    S.cur ^= 1;
    S.steps_since_sort = 0;
    if (trim)
    {
        // explicit mag2d_sort: release the dead tail now (one host sync); sorts issued from inside
        // mag2d_step stay asynchronous and keep the old slot count as an upper bound
        unsigned long long total = 0;
        CUDA_OK(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        S.n_slots = (long long)total;
    }
    return 0;
}

// ---- cell sort fused into the Boris push (push.cu, SORTING kernels) -----------------------------------------------
// Instead of a stand-alone count + scatter pass every K steps (16 + 88 bytes per particle), the push itself does the
// work of both: a COUNT step counts the surviving particles per cell at the position they have just been moved to
// (one RED per warp and cell); the counts are scanned into per-cell cursors; the next PERMUTE step reads the particles
// in slot order as always — each one at exactly the position it was counted at, so it recomputes its cell — draws its
// slot from the cell's cursor (one returning atomic per warp and cell) and writes the particle there in the other
// slab, dropping dead slots on the way.  The store is then sorted by the cell each particle occupied one step earlier,
// which is as good for the gather / scatter locality, with no per-particle ticket arrays and no extra traffic
// beyond the scattered stores themselves.
void sort_fused_free(SpeciesStore& S)
{
    cudaFree(S.d_cell_count);
    cudaFree(S.d_cell_offset);
    cudaFree(S.d_sort_sums);
    if (S.h_total) cudaFreeHost(S.h_total);
    if (S.ev_total) cudaEventDestroy(S.ev_total);
    S.h_total = nullptr;
    S.ev_total = nullptr;
    S.total_pending = false;
    S.d_cell_count = S.d_cell_offset = S.d_sort_sums = nullptr;
    S.tickets_valid = false;
}

int sort_fused_begin(mag2d_ctx* c, int s, bool permute, bool count)
{
    SpeciesStore& S = c->sp[s];
    const int ncells = (c->g.M - 1) * (c->g.N - 1) * (is3d(c) ? c->g.K - 1 : 1);
    const int ntiles = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    if (!S.d_cell_count)
    {
        CUDA_OK(cudaMalloc(&S.d_cell_count, sizeof(unsigned) * (size_t)ncells));
        CUDA_OK(cudaMalloc(&S.d_cell_offset, sizeof(unsigned) * (size_t)ncells));
        CUDA_OK(cudaMalloc(&S.d_sort_sums, sizeof(unsigned) * (size_t)(ntiles + 8) + 16));
    }
    if (count) CUDA_OK(cudaMemsetAsync(S.d_cell_count, 0, sizeof(unsigned) * (size_t)ncells, c->stream));
    // a live count sent home by an earlier permute has landed and nothing was appended since: everything behind it
    // is dead in the current slab, so the slot range shrinks to it (rounded up to the 256-slot allocation unit)
    if (S.total_pending && cudaEventQuery(S.ev_total) == cudaSuccess)
    {
        S.total_pending = false;
        if (S.total_epoch == S.append_epoch)
        {
            const long long live = (long long)*S.h_total;
            if (live < S.n_slots) S.n_slots = live;
        }
    }
    if (permute)
    {
        if (!S.arr[S.cur ^ 1][ARR_X] && store_alloc_slab(c, S, S.cur ^ 1, S.capacity)) return 1;
        // removals are counted per compaction period: the dead slots of the old slab are dropped by this step
        CUDA_OK(cudaMemsetAsync(S.d_removed, 0, sizeof(unsigned long long), c->stream));
    }
    return 0;
}

int sort_fused_end(mag2d_ctx* c, int s, bool permute, bool count)
{
    SpeciesStore& S = c->sp[s];
    const int ncells = (c->g.M - 1) * (c->g.N - 1) * (is3d(c) ? c->g.K - 1 : 1);
    const int ntiles = (ncells + SCAN_TILE - 1) / SCAN_TILE;
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(S.d_sort_sums + ((ntiles + 1) / 2 * 2 + 2));
    if (permute)
    {
        // d_total still holds the number of particles the consumed cursors covered: everything behind is dead
        if (c->store_f32) k_fill_dead_keys<float><<<148 * 4, 256, 0, c->stream>>>(S.arr[S.cur ^ 1][ARR_X], nullptr, d_total, S.n_slots);
        else k_fill_dead_keys<double><<<148 * 4, 256, 0, c->stream>>>(S.arr[S.cur ^ 1][ARR_X], nullptr, d_total, S.n_slots);
        c->launches++;
        if (!S.total_pending)
        {
            if (!S.h_total)
            {
                CUDA_OK(cudaMallocHost(&S.h_total, sizeof(unsigned long long)));
                CUDA_OK(cudaEventCreateWithFlags(&S.ev_total, cudaEventDisableTiming));
            }
            CUDA_OK(cudaMemcpyAsync(S.h_total, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaEventRecord(S.ev_total, c->stream));
            S.total_pending = true;
            S.total_epoch = S.append_epoch;
        }
        S.cur ^= 1;
        S.pushes_since_permute = 0;
        S.steps_since_sort = 0;
        S.tickets_valid = false;
    }
    if (count)
    {
        k_scan_tiles<<<ntiles, SCAN_THREADS, 0, c->stream>>>(S.d_cell_count, S.d_cell_offset, S.d_sort_sums, ncells);
        k_scan_sums<<<1, 1024, 0, c->stream>>>(S.d_sort_sums, ntiles, d_total);
        k_scan_add<<<ntiles, SCAN_THREADS, 0, c->stream>>>(S.d_cell_offset, S.d_sort_sums, ncells);
        c->launches += 3;
        S.tickets_valid = true;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}
