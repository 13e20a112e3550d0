// push3d.cu — the 3-D Cartesian (3D3V) step: field gather + Boris push + null-collision MCC + box boundary /
// electrode absorption + fixed-point CIC deposit on eight nodes, one kernel per species and step.
//
// Reference code replaced (dead code there — species3d.cpp does not compile — restated in oracle/mag3d_oracle.c):
//   Species<CARTESIAN3D>::advance              src/species3d.cpp:3-93
//   ElMag3D::E / Field3D::grad, grad_component  src/fields3d.hpp:95-101, src/Field3D.hpp:110-163
//   Geometry::is_free                           src/fields3d.hpp:48-57
//   Field3D::accumulate                         src/Field3D.hpp:40-65
// Grids use the reference's Array3D layout a[(i*K + j)*N + k] (i along x, j along y, k along z; M, K, N nodes).
// 96 B per particle-step (six fp64 components read and written once).
#include <algorithm>
#include <cstring>

#include "push3d.cuh"

namespace {

constexpr int P3_THREADS = 256;
#ifndef MAG3D_MIN_BLOCKS
#define MAG3D_MIN_BLOCKS 3      // 80 registers, no spills (the weights are formed after the stores); C5: 5.77 ms vs 6.49 ms with 2
#endif

// warp-aggregated scatter: lanes that share a cell are summed with REDUX (two 16/17-bit pieces per weight), lanes
// 0..7 issue one RED.ADD.64 each; after max_runs distinct cells the remaining lanes scatter on their own.  The merge
// only pays when a warp's 64 slots share a few cells: measured on C5 (7.5 particles per cell and GPU) plain scatter
// is 8 % faster (5.72 vs 6.24 ms), so the host enables it from 32 particles per cell upwards.
__device__ __forceinline__ void warp_deposit3(const Grid3Dev& g, bool valid, unsigned node, const unsigned long long (&w)[8], int max_runs)
{
    const unsigned lane = lane_id();
    unsigned remaining = __ballot_sync(MAG2D_FULL_MASK, valid);
    if (remaining == 0) return;
    const unsigned sj = (unsigned)g.N, si = (unsigned)(g.K * g.N);
#pragma unroll 1
    for (int it = 0; remaining && it < max_runs; it++)
    {
        const int src = __ffs(remaining) - 1;
        const unsigned k0 = __shfl_sync(MAG2D_FULL_MASK, node, src);
        const bool mine = valid && node == k0;
        const unsigned m = __ballot_sync(MAG2D_FULL_MASK, mine);
        unsigned long long mysum = 0;
#pragma unroll
        for (int q = 0; q < 8; q++)
        {
            const unsigned lo = __reduce_add_sync(MAG2D_FULL_MASK, mine ? (unsigned)(w[q] & 0xFFFFu) : 0u);
            const unsigned hi = __reduce_add_sync(MAG2D_FULL_MASK, mine ? (unsigned)(w[q] >> 16) : 0u);
            if ((int)lane == q) mysum = ((unsigned long long)hi << 16) + lo;
        }
        if (lane < 8) atomicAdd(g.rho + k0 + (lane & 1 ? si : 0u) + (lane & 2 ? sj : 0u) + (lane >> 2), mysum);
        remaining &= ~m;
    }
    if (valid && ((remaining >> lane) & 1u)) scatter3(g, node, w);
}

// ---- the fused 3-D step -------------------------------------------------------------------------------------------
// Each thread owns two neighbouring slots (128-bit loads and stores, a warp moves 512 B per instruction).  PUSH =
// false is the deposit-only pass of Pic::advance_init.
template <bool PUSH, bool GATHER, bool HASB, bool MCC, bool DEPOSIT, bool SORTING>
__global__ void __launch_bounds__(P3_THREADS, MAG3D_MIN_BLOCKS) k_push3d(const __grid_constant__ Push3Args A)
{
    const bool permute = SORTING && A.permute, count = SORTING && A.count;
    const unsigned lane = lane_id();
    const long long k = 2 * ((long long)blockIdx.x * P3_THREADS + threadIdx.x);
    const long long n = A.p.n;
    if ((k & ~63LL) >= n) return;       // warp-uniform: the slabs are allocated in multiples of 256 slots
    double x[2], y[2], z[2], vx[2], vy[2], vz[2];
    {
        const double2 a = *reinterpret_cast<const double2*>(A.p.x + k);
        const double2 b = *reinterpret_cast<const double2*>(A.p.y + k);
        const double2 c = *reinterpret_cast<const double2*>(A.p.z + k);
        x[0] = a.x; x[1] = a.y; y[0] = b.x; y[1] = b.y; z[0] = c.x; z[1] = c.y;
        if (PUSH)
        {
            const double2 d = *reinterpret_cast<const double2*>(A.p.vx + k);
            const double2 e = *reinterpret_cast<const double2*>(A.p.vy + k);
            const double2 f = *reinterpret_cast<const double2*>(A.p.vz + k);
            vx[0] = d.x; vx[1] = d.y; vy[0] = e.x; vy[1] = e.y; vz[0] = f.x; vz[1] = f.y;
        }
    }
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (MCC)
    {
        Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)k);
        rnd = rng.block();
    }
    long long dest[2] = {k, k + 1};
    if (permute)
    {
        // the particle sits exactly where the COUNT push left it: recompute that cell (same operations as boundary3)
        // and draw the next slot of the cell from the cursor array (the scanned counts)
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            const bool alive = (k + e < n) && particle_alive(x[e]);
            const int ci = max(min((int)__dmul_rn(x[e], A.g.idx), A.g.M - 2), 0), cj = max(min((int)__dmul_rn(y[e], A.g.idy), A.g.K - 2), 0),
                      ck = max(min((int)__dmul_rn(z[e], A.g.idz), A.g.N - 2), 0);
            const unsigned key = ((unsigned)ci * (unsigned)(A.g.K - 1) + (unsigned)cj) * (unsigned)(A.g.N - 1) + (unsigned)ck;
            const unsigned slot = warp_ticket(A.cursor, alive, alive ? key : 0u);
            dest[e] = alive ? (long long)slot : -1;
        }
    }
    const double dt = A.s.dt;
    bool keep[2];
    unsigned node[2], cell[2] = {0, 0}, removed = 0, hit_mask = 0;
#pragma unroll
    for (int e = 0; e < 2; e++)
    {
        const bool live = (k + e < n) && particle_alive(x[e]);
        if (PUSH)
        {
            double Ex = 0.0, Ey = 0.0, Ez = 0.0;
            if (GATHER)
            {
                const double X = x[e] * A.g.idx, Y = y[e] * A.g.idy, Z = z[e] * A.g.idz;
                Ex = -grad_component<0>(A.g, X, Y, Z);
                Ey = -grad_component<1>(A.g, X, Y, Z);
                Ez = -grad_component<2>(A.g, X, Y, Z);
            }
            // half acceleration, rotation (species3d.cpp:39-50, its own sign convention), half acceleration
            vx[e] += Ex * A.s.hq;
            vy[e] += Ey * A.s.hq;
            vz[e] += Ez * A.s.hq;
            if (HASB)
            {
                const double px = vx[e] - vy[e] * A.s.tz + vz[e] * A.s.ty;
                const double py = vy[e] - vz[e] * A.s.tx + vx[e] * A.s.tz;
                const double pz = vz[e] - vx[e] * A.s.ty + vy[e] * A.s.tx;
                const double ox = vx[e], oy = vy[e], oz = vz[e];
                vx[e] = ox - py * A.s.sz + pz * A.s.sy;
                vy[e] = oy - pz * A.s.sx + px * A.s.sz;
                vz[e] = oz - px * A.s.sy + py * A.s.sx;
            }
            vx[e] += Ex * A.s.hq;
            vy[e] += Ey * A.s.hq;
            vz[e] += Ez * A.s.hq;
            x[e] += vx[e] * dt;
            y[e] += vy[e] * dt;
            z[e] += vz[e] * dt;
        }
        const bool inside = boundary3(A.g, x[e], y[e], z[e], node[e], SORTING ? &cell[e] : nullptr);
        keep[e] = live && inside;
        removed += (live && !inside) ? 1u : 0u;
        if (!keep[e]) x[e] = dead_marker();
        if (MCC)
        {
            const unsigned word = e == 0 ? rnd.x : rnd.y;
            if (keep[e] && (unsigned long long)word < A.s.prob_u32) hit_mask |= 1u << e;
        }
    }
    if (PUSH && permute)
    {
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            const long long d = dest[e];
            if (d < 0) continue;
            A.dst.x[d] = x[e]; A.dst.y[d] = y[e]; A.dst.z[d] = z[e];
            A.dst.vx[d] = vx[e]; A.dst.vy[d] = vy[e]; A.dst.vz[d] = vz[e];
        }
    }
    else if (PUSH)
    {
        *reinterpret_cast<double2*>(A.p.x + k) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2*>(A.p.y + k) = make_double2(y[0], y[1]);
        *reinterpret_cast<double2*>(A.p.z + k) = make_double2(z[0], z[1]);
        *reinterpret_cast<double2*>(A.p.vx + k) = make_double2(vx[0], vx[1]);
        *reinterpret_cast<double2*>(A.p.vy + k) = make_double2(vy[0], vy[1]);
        *reinterpret_cast<double2*>(A.p.vz + k) = make_double2(vz[0], vz[1]);
    }
    if (count)
    {
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            const unsigned key = keep[e] ? cell[e] : SORT_INVALID_KEY;
            warp_count(A.count_out, keep[e], key);
        }
    }
    if (DEPOSIT)
    {
        unsigned long long w0[8], w1[8];
        weights3(A.g, x[0], y[0], z[0], w0);
        weights3(A.g, x[1], y[1], z[1], w1);
        if (keep[0] && keep[1] && node[0] == node[1])
        {
#pragma unroll
            for (int c = 0; c < 8; c++) w0[c] += w1[c];
            keep[1] = false;
        }
        if (A.deposit_runs > 0)
        {
            warp_deposit3(A.g, keep[0], node[0], w0, A.deposit_runs);
            warp_deposit3(A.g, keep[1], node[1], w1, A.deposit_runs);
        }
        else
        {
            if (keep[0]) scatter3(A.g, node[0], w0);
            if (keep[1]) scatter3(A.g, node[1], w1);
        }
    }
    if (MCC)
    {
        const unsigned cnt = __popc(hit_mask);
        if (__any_sync(MAG2D_FULL_MASK, cnt != 0))
        {
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned t = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
                if (lane >= (unsigned)o) incl += t;
            }
            const unsigned total = __shfl_sync(MAG2D_FULL_MASK, incl, 31);
            unsigned start = 0;
            if (lane == 31) start = atomicAdd(A.coll_count, total);
            start = __shfl_sync(MAG2D_FULL_MASK, start, 31) + incl - cnt;
            if (hit_mask & 1u) A.coll_list[start++] = (unsigned)dest[0];
            if (hit_mask & 2u) A.coll_list[start++] = (unsigned)dest[1];
        }
    }
    if (PUSH && __any_sync(MAG2D_FULL_MASK, removed != 0))
    {
        unsigned rsum = removed;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(MAG2D_FULL_MASK, rsum, o);
        if (lane == 0) atomicAdd(A.removed, (unsigned long long)rsum);
    }
}

// BaseSpecies::scatter for the slots whose Bernoulli test fired (velocity space only: shared with the 2-D movers)
__global__ void __launch_bounds__(128) k_mcc_collide3d(const __grid_constant__ Push3Args A)
{
    const unsigned n = *A.coll_count;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
    {
        // brick mode: entries with the high bit set are leavers waiting in the outbox (A.dst), see push3d_brick.cu
        const unsigned raw = A.coll_list[q];
        const ParticlesDev& P = (raw & 0x80000000u) ? A.dst : A.p;
        const long long k = raw & 0x7FFFFFFFu;
        double vx = P.vx[k], vy = P.vy[k], vz = P.vz[k];
        Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)raw);
        rng.draw = 1;
        int target;
        const int proc = mcc_scatter(A.mcc, rng, vx, vy, vz, target);
        mcc_count(A.counts, A.mcc->n_targets, target, proc);
        if (proc >= 0)
        {
            P.vx[k] = vx;
            P.vy[k] = vy;
            P.vz[k] = vz;
        }
    }
}

// gx'[i][j][k] = u[ic][jc][kc] - u[ic-1][jc][kc] with ic = clamp(i, 1, M-1), jc = min(j, K-1), kc = min(k, N-1) on the
// ghost-extended index range [0..M][0..K][0..N]; gy', gz' likewise along their own axes.  The eight differences that
// Field3D::grad_component forms per particle become eight loads.
__global__ void k_edge_fields3d(const double* __restrict__ u, int M, int K, int N, double* __restrict__ gx, double* __restrict__ gy,
                                double* __restrict__ gz)
{
    // one block per row (i, j) of the ghost-extended range, threads along k: no integer divisions per element
    const size_t sj = (size_t)N + 1, si = ((size_t)K + 1) * sj;
    const size_t uj = N, ui = (size_t)K * N;
    const int i = (int)(blockIdx.x / (unsigned)(K + 1)), j = (int)(blockIdx.x % (unsigned)(K + 1));
    const int i0 = min(i, M - 1), j0 = min(j, K - 1);
    const int i1 = max(i0, 1), j1 = max(j0, 1);
    for (int k = threadIdx.x; k <= N; k += blockDim.x)
    {
        const size_t m = (size_t)i * si + (size_t)j * sj + k;
        const int k0 = min(k, N - 1);
        const int k1 = max(k0, 1);
        gx[m] = u[i1 * ui + j0 * uj + k0] - u[(i1 - 1) * ui + j0 * uj + k0];
        gy[m] = u[i0 * ui + j1 * uj + k0] - u[i0 * ui + (j1 - 1) * uj + k0];
        gz[m] = u[i0 * ui + j0 * uj + k1] - u[i0 * ui + j0 * uj + k1 - 1];
    }
}

__global__ void k_field_E3d(Grid3Dev g, int n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                            double* __restrict__ Ex, double* __restrict__ Ey, double* __restrict__ Ez)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double X = x[p] * g.idx, Y = y[p] * g.idy, Z = z[p] * g.idz;
    Ex[p] = -grad_component<0>(g, X, Y, Z);
    Ey[p] = -grad_component<1>(g, X, Y, Z);
    Ez[p] = -grad_component<2>(g, X, Y, Z);
}

Grid3Dev grid3_view(const mag2d_ctx* c, int s)
{
    Grid3Dev g;
    memset(&g, 0, sizeof(g));
    const mag2d_grid_desc& d = c->g;
    g.M = d.M; g.K = d.K; g.N = d.N;
    g.boundary = d.boundary;
    g.check_mask = !c->all_cells_free;
    g.deposit = d.selfconsistent;
    g.x_max = d.x_max; g.y_max = d.y_max; g.z_max = d.z_max;
    g.idx = d.idx; g.idy = d.idy; g.idz = d.idz;
    g.gx = c->d_gx; g.gy = c->d_gy; g.gz = c->d_gz;
    g.cfree = c->d_cfree;
    g.rho = s >= 0 && c->d_rho ? c->d_rho + (size_t)s * d.M * d.K * d.N : nullptr;
    return g;
}

SpeciesDev species3_view(const mag2d_ctx* c, int s)
{
    const SpeciesStore& S = c->sp[s];
    const mag2d_grid_desc& d = c->g;
    SpeciesDev v;
    memset(&v, 0, sizeof(v));
    const double charge = S.desc.charge, mass = S.desc.mass, dt = S.desc.dt;
    v.dt = dt;
    v.hq = charge / mass * dt / 2.0;
    // (Bx, By, Bz) = (Br, Bt, Bz) of the grid descriptor; the reference hard-wires zero (species3d.cpp:18)
    double tmp = charge * dt / (2.0 * mass);
    v.tx = d.Br * tmp;
    v.ty = d.Bt * tmp;
    v.tz = d.Bz * tmp;
    tmp = 2.0 / (1 + v.tx * v.tx + v.ty * v.ty + v.tz * v.tz);
    v.sx = v.tx * tmp;
    v.sy = v.ty * tmp;
    v.sz = v.tz * tmp;
    v.has_B = (d.Br != 0.0 || d.Bt != 0.0 || d.Bz != 0.0);
    v.species = s;
    v.prob = 1.0 - exp(-dt / S.lifetime);
    v.prob_u32 = bernoulli_threshold(v.prob);
    v.lifetime = S.lifetime;
    v.qm = charge / mass;
    v.step = S.niter;
    return v;
}

ParticlesDev particles3_view(const SpeciesStore& S)
{
    ParticlesDev p;
    double* const* a = S.arr[S.cur];
    p.x = a[ARR_X]; p.z = a[ARR_Z]; p.vx = a[ARR_VX]; p.vy = a[ARR_VY]; p.vz = a[ARR_VZ]; p.y = a[ARR_Y]; p.ttd = a[ARR_TTD];
    p.n = S.n_slots;
    return p;
}

}  // namespace

int update_edge_fields3d(mag2d_ctx* c)
{
    k_edge_fields3d<<<(unsigned)((c->g.M + 1) * (c->g.K + 1)), std::min(256, (c->g.N + 32) / 32 * 32), 0, c->stream>>>(c->d_u, c->g.M, c->g.K, c->g.N, c->d_gx,
                                                                                                                    c->d_gy, c->d_gz);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Species<CARTESIAN3D>::advance (deposit_only: the accumulate pass of Pic::advance_init); sort_mode as in 2-D
int launch_species_advance3d(mag2d_ctx* c, int s, bool deposit_only, int sort_mode)
{
    SpeciesStore& S = c->sp[s];
    const mag2d_grid_desc& d = c->g;
    // streamed step (abi.cu): the particle arrays are one chunk of a host-resident store staged in device buffers
    const bool chunked = c->chunk_view != nullptr && !deposit_only;
    // a push that does not consume the pending cell cursors invalidates them (see launch_species_advance)
    if (!chunked && !deposit_only && !(sort_mode & 1)) S.tickets_valid = false;
    // brick mode (sort_mode bit 2): the step runs on the brick-binned store; any other push of the resident store moves particles
    // behind the bins' back
    const bool brick = !chunked && !deposit_only && (sort_mode & 4) != 0;
    const int brick_compact_every = std::max(1, sort_mode >> 3);     // bits 3..: steps between compacting (cell-sorting) passes
    if (!chunked && !deposit_only && !brick) S.bins_valid = false;
    if (brick)
    {
        brick_poll_overflow(S);
        if (!S.bins_valid && brick_rebuild(c, s, grid3_view(c, s))) return 1;
        sort_mode = 0;
    }
    const long long n_active = chunked ? c->chunk_view->n : S.n_slots;
    if (n_active > 0)
    {
        Push3Args A;
        A.g = grid3_view(c, s);
        A.s = species3_view(c, s);
        A.p = chunked ? *c->chunk_view : particles3_view(S);
        A.mcc = S.d_blob;
        A.counts = c->count_collisions ? S.d_counts : nullptr;
        A.removed = S.d_removed;
        A.seed = c->seed;
        if (chunked && c->chunk_slot0)
        {
            // every chunk draws from its own Philox key (splitmix64 of its first global slot), as in the 2-D launcher
            unsigned long long zz = (unsigned long long)c->chunk_slot0 + 0x9E3779B97F4A7C15ULL;
            zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ULL;
            zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBULL;
            A.seed ^= zz ^ (zz >> 31);
        }
        A.coll_list = nullptr;
        A.coll_count = nullptr;
        A.permute = A.count = 0;
        A.cursor = A.count_out = nullptr;
        memset(&A.dst, 0, sizeof(A.dst));
        const double per_cell = (double)n_active / ((double)(d.M - 1) * (d.K - 1) * (d.N - 1));
        A.deposit_runs = per_cell >= 32.0 ? 3 : 0;
        const unsigned blocks = (unsigned)((n_active + 2 * P3_THREADS - 1) / (2 * P3_THREADS));
        const bool mcc = !deposit_only && S.h_blob && S.h_blob->has_collisions && std::isfinite(S.lifetime);
        const bool deposit = d.selfconsistent != 0;
        if (deposit_only)
        {
            if (deposit) k_push3d<false, false, false, false, true, false><<<blocks, P3_THREADS, 0, c->stream>>>(A);
        }
        else
        {
            // the potential does not change inside a step: later chunks of a streamed step reuse the edge fields
            if (!(chunked && c->chunk_slot0 > 0) && update_edge_fields3d(c)) return 1;
            if (mcc)
            {
                if (ensure_particle_scratch(c, chunked ? std::max(n_active, S.capacity) : S.capacity)) return 1;
                A.coll_list = c->d_key;
                A.coll_count = c->d_coll_count;
                CUDA_OK(cudaMemsetAsync(c->d_coll_count, 0, sizeof(unsigned), c->stream));
            }
            const bool permute = (sort_mode & 1) != 0, count = (sort_mode & 2) != 0;
            const bool sorting = permute || count;
            if (brick)
            {
                brick_outbox_view(S, A.dst);        // collision-list entries with the high bit set address the outbox
                if (launch_brick_push(c, s, A, mcc, deposit, brick_compact_every)) return 1;
                if (mcc)
                {
                    k_mcc_collide3d<<<148 * 8, 128, 0, c->stream>>>(A);
                    c->launches++;
                }
                if (launch_brick_migrate(c, s, A)) return 1;
                CUDA_OK(cudaGetLastError());
                S.niter++;
                S.t += S.desc.dt;
                S.steps_since_sort++;
                return 0;
            }
            if (sorting)
            {
                if (sort_fused_begin(c, s, permute, count)) return 1;
                A.permute = permute;
                A.count = count;
                A.cursor = S.d_cell_offset;
                A.count_out = S.d_cell_count;
                if (permute)
                {
                    double* const* o = S.arr[S.cur ^ 1];
                    A.dst.x = o[ARR_X]; A.dst.y = o[ARR_Y]; A.dst.z = o[ARR_Z];
                    A.dst.vx = o[ARR_VX]; A.dst.vy = o[ARR_VY]; A.dst.vz = o[ARR_VZ];
                    A.dst.n = S.n_slots;
                }
            }
            const int code = (sorting ? 8 : 0) | (A.s.has_B ? 4 : 0) | (mcc ? 2 : 0) | (deposit ? 1 : 0);
#define L3(B, Mc, D, So) k_push3d<true, true, B, Mc, D, So><<<blocks, P3_THREADS, 0, c->stream>>>(A)
            switch (code)
            {
                case 0: L3(false, false, false, false); break;
                case 1: L3(false, false, true, false); break;
                case 2: L3(false, true, false, false); break;
                case 3: L3(false, true, true, false); break;
                case 4: L3(true, false, false, false); break;
                case 5: L3(true, false, true, false); break;
                case 6: L3(true, true, false, false); break;
                case 7: L3(true, true, true, false); break;
                case 8: L3(false, false, false, true); break;
                case 9: L3(false, false, true, true); break;
                case 10: L3(false, true, false, true); break;
                case 11: L3(false, true, true, true); break;
                case 12: L3(true, false, false, true); break;
                case 13: L3(true, false, true, true); break;
                case 14: L3(true, true, false, true); break;
                default: L3(true, true, true, true); break;
            }
#undef L3
            if (sorting)
            {
                if (sort_fused_end(c, s, permute, count)) return 1;
                if (permute)
                {
                    A.p = particles3_view(S);
                    if (mcc && refresh_pools_all(c)) return 1;      // partner pools follow the slab flip
                }
            }
            if (mcc)
            {
                k_mcc_collide3d<<<148 * 8, 128, 0, c->stream>>>(A);
                c->launches++;
            }
        }
        c->launches++;
        CUDA_OK(cudaGetLastError());
    }
    if (chunked) return 0;          // the streamed step advances the species clock once per step, not per chunk
    if (!deposit_only)
    {
        S.niter++;
        S.t += S.desc.dt;
        S.steps_since_sort++;
        if (S.pushes_since_permute < (1 << 20)) S.pushes_since_permute++;
    }
    return 0;
}

int launch_field_E3d(mag2d_ctx* c, int n, const double* x, const double* y, const double* z, double* Ex, double* Ey, double* Ez)
{
    if (n <= 0) return 0;
    if (update_edge_fields3d(c)) return 1;
    double* d;
    CUDA_OK(cudaMalloc(&d, sizeof(double) * 6 * (size_t)n));
    CUDA_OK(cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(d + n, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(d + 2 * (size_t)n, z, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    k_field_E3d<<<(n + 255) / 256, 256, 0, c->stream>>>(grid3_view(c, -1), n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, d + 4 * (size_t)n,
                                                         d + 5 * (size_t)n);
    c->launches++;
    CUDA_OK(cudaMemcpyAsync(Ex, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(Ey, d + 4 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(Ez, d + 5 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d));
    return 0;
}
