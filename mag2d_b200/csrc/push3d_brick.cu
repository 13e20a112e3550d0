// push3d_brick.cu — the 3-D step on a brick-binned particle store: one CTA owns one 4 x 4 x 4-cell brick.
//
// Same physics and the same arithmetic as k_push3d (push3d.cu) — Species<CARTESIAN3D>::advance, src/species3d.cpp:3-93;
// Field3D::grad / accumulate, src/Field3D.hpp:40-163; Geometry::is_free, src/fields3d.hpp:48-57 — but organised around
// the grid instead of around the slot order, because at C5's 7.5 particles per cell the slot-order kernel is bound by
// the L1 data pipe (24 gather loads and 8 RED.64 lanes per particle), not by HBM:
//
//   * the store is binned by brick: brick b owns the slots [bin_off[b], bin_off[b+1]) of every SoA array, of which the
//     first bin_cnt[b] are in use (live particles, plus holes marked x = NaN); the rest is slack for arrivals;
//   * a CTA stages its bin in shared memory with the TMA engine (cp.async.bulk, one copy per array), and the three
//     ghost-extended edge-difference tiles of its brick (6 x 5 x 5 doubles each) next to it: the 24 gather loads of a
//     particle are shared-memory loads, and since the bin is cell-sorted the lanes of a warp mostly read the same words;
//   * after the push the CTA counting-sorts its particles by their new cell in shared memory (one shared atomic per
//     particle), writes the stayers back compacted and cell-sorted, and forms the charge of every cell with a small
//     team of threads per CELL (registers, no atomics); the 5 x 5 x 5 node sums of the brick are flushed with ONE
//     RED.ADD.64 per node and step instead of eight per particle;
//   * particles that leave the brick stay in their bin for the moment (deposited with plain REDs at their new position)
//     and are listed; k_migrate3d then moves them to the tail of their new brick's bin.  A full bin or a full list only
//     means the particle stays a guest of its old bin (gathered / deposited through the global-memory path) and is
//     offered again next step, and the store is re-binned with more slack — never a lost particle.
//
// Integer sums are associative, so the charge grid is bit-identical to the slot-order kernel's; trajectories are
// bit-identical too (same operations in the same order, only the operands come from shared memory).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "push3d.cuh"

namespace {

constexpr int BR = 4;                    // cells per brick edge
constexpr int BCELLS = BR * BR * BR;     // 64
constexpr int BK_THREADS = 256;
constexpr int BK_PPT = 3;
constexpr int BK_CHUNK = BK_THREADS * BK_PPT;      // slots staged per pass (a bin of C5 holds ~480 + arrivals)
constexpr int TILE_X = (BR + 2) * (BR + 1) * (BR + 1);   // planes along the differenced axis: BR + 2, along the others: BR + 1
constexpr int TILE_LD = 152;             // TILE_X = 150 padded
constexpr int CLS_GUEST = BCELLS;        // alive, but not (or no longer) inside this brick
constexpr unsigned char CLS_DEAD = 255;
constexpr int LIST_CAP = 128;            // leavers / collision hits of one pass that are flushed with a single atomic (more: direct appends)
#ifndef MAG3D_BRICK_MIN_BLOCKS
#define MAG3D_BRICK_MIN_BLOCKS 4
#endif

struct Outbox
{
    double* arr[6];              // x, y, z, vx, vy, vz of this step's leavers
    unsigned* dst;               // destination brick (OUTBOX_VOID: a reserved entry that was not filled)
    unsigned* count;             // [0] entries reserved this step, [1] bins found full by arrivals (since the last re-binning), [2] CTAs whose leavers did not fit
    unsigned cap;
};
constexpr unsigned OUTBOX_VOID = 0xFFFFFFFFu;

struct BrickArgs
{
    Push3Args A;
    const unsigned* bin_off;     // [nb + 1]
    unsigned* bin_cnt;           // [nb]
    int nbx, nby, nbz;
    int compact;                 // this step writes the stayers back compacted and cell-sorted (else: in place, holes stay)
    Outbox ob;
};

struct __align__(128) BrickSmem
{
    double P[6][BK_CHUNK];                       // x, y, z, vx, vy, vz of the staged slots
    double tile[3][TILE_LD];                     // edge differences gx [6][5][5], gy [5][6][5], gz [5][5][6]
    unsigned long long cellsum[8][BCELLS];       // Q32 weight sums, [corner][cell]: neighbouring cells sit in neighbouring banks
    unsigned cnt[BCELLS + 4];                    // particles per class (64 cells + leavers)
    unsigned start[BCELLS + 4];                  // exclusive scan of cnt
    unsigned dstb[BK_CHUNK];                     // destination brick of a leaver
    unsigned short rank[BK_CHUNK];               // rank inside the class
    unsigned short perm[BK_CHUNK];               // sorted position -> staged slot (stayers)
    unsigned char cls[BK_CHUNK];                 // class (cell 0..63, CLS_GUEST) | 128 when the collision test fired; CLS_DEAD
    unsigned short hits[LIST_CAP];               // staged slots whose collision test fired
    unsigned short leave[LIST_CAP];              // staged slots of this pass' leavers, by rank
    unsigned n_hit, hit_base, ob_base, pad;
    unsigned long long bar;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one component of grad u from the brick's tile: grad_component (push3d.cuh) with shared-memory operands.  (li, lj, lk)
// = integer parts of the particle's index-space position relative to the brick's first cell, all in [0, BR).
template <int DIR>
__device__ __forceinline__ double grad_tile(const Grid3Dev& g, const double* __restrict__ t, double X, double Y, double Z, int ci0, int cj0, int ck0)
{
    const double xs = DIR == 0 ? X + 0.5 : X, ys = DIR == 1 ? Y + 0.5 : Y, zs = DIR == 2 ? Z + 0.5 : Z;
    const int i = (int)xs, j = (int)ys, k = (int)zs;
    const double u = xs - i, v = ys - j, w = zs - k;
    // tile extents: BR + 2 planes along the differenced axis, BR + 1 along the others; k fastest
    constexpr int nj = DIR == 1 ? BR + 2 : BR + 1, nk = DIR == 2 ? BR + 2 : BR + 1;
    constexpr int sj = nk, si = nj * nk;
    const double* f = t + ((i - ci0) * si + (j - cj0) * sj + (k - ck0));
    const double g0 = f[0], g1 = f[si], g2 = f[sj], g3 = f[si + sj];
    const double g4 = f[1], g5 = f[si + 1], g6 = f[sj + 1], g7 = f[si + sj + 1];
    return trilerp(u, v, w, g0, g1, g2, g3, g4, g5, g6, g7) * (DIR == 0 ? g.idx : DIR == 1 ? g.idy : g.idz);
}

// boundary3 (push3d.cuh) that also returns the cell's integer coordinates
__device__ __forceinline__ bool boundary3_ijk(const Grid3Dev& g, double& x, double& y, double& z, int& i, int& j, int& k)
{
    i = j = k = 0;
    if (!(x >= 0.0 && x <= g.x_max && y >= 0.0 && y <= g.y_max && z >= 0.0 && z <= g.z_max))
    {
        if (g.boundary == MAG2D_BOUNDARY_FREE || !(x == x && y == y && z == z)) return false;
        x = fmod(x, g.x_max); if (x < 0) x += g.x_max;
        y = fmod(y, g.y_max); if (y < 0) y += g.y_max;
        z = fmod(z, g.z_max); if (z < 0) z += g.z_max;
    }
    const double X = __dmul_rn(x, g.idx), Y = __dmul_rn(y, g.idy), Z = __dmul_rn(z, g.idz);
    i = max(min((int)X, g.M - 2), 0);
    j = max(min((int)Y, g.K - 2), 0);
    k = max(min((int)Z, g.N - 2), 0);
    if (g.check_mask && !g.cfree[((size_t)i * g.K + j) * g.N + k]) return false;
    return true;
}

template <bool HASB, bool MCC, bool DEPOSIT>
__global__ void __launch_bounds__(BK_THREADS, MAG3D_BRICK_MIN_BLOCKS) k_push3d_brick(const __grid_constant__ BrickArgs B)
{
    extern __shared__ __align__(128) unsigned char brick_smem[];
    BrickSmem& S = *reinterpret_cast<BrickSmem*>(brick_smem);
    const Push3Args& A = B.A;
    const unsigned t = threadIdx.x, lane = t & 31u;
    const int bk = (int)blockIdx.x, bj = (int)blockIdx.y, bi = (int)blockIdx.z;      // grid (nbz, nby, nbx): z fastest, like the bins
    const unsigned b = (unsigned)((bi * B.nby + bj) * B.nbz + bk);
    const unsigned n = B.bin_cnt[b];
    if (n == 0) return;
    const unsigned off = B.bin_off[b];
    const bool compact = B.compact != 0;
    const int ci0 = bi * BR, cj0 = bj * BR, ck0 = bk * BR;
    double* const arr[6] = {A.p.x, A.p.y, A.p.z, A.p.vx, A.p.vy, A.p.vz};
    if (t == 0)
    {
        mbar_init(&S.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto chunk_bytes = [&](unsigned in) {
        const unsigned m = min((unsigned)BK_CHUNK, n - in);
        return ((m + 1u) & ~1u) * (unsigned)sizeof(double);      // bins are padded to 32 slots: the extra slot exists
    };
    auto issue_load = [&](unsigned in) {
        const unsigned bytes = chunk_bytes(in);
        mbar_expect_tx(&S.bar, 6 * bytes);
#pragma unroll
        for (int a = 0; a < 6; a++) bulk_load(S.P[a], arr[a] + off + in, bytes, &S.bar);
    };
    if (t == 0) issue_load(0);
    // the brick's edge-difference tiles (ghost-extended arrays [M+1][K+1][N+1]; indices past the end are clamped: those
    // planes are never read, a particle's cell lies inside the grid).  All divisors are compile-time constants.
    {
        const unsigned gsj = (unsigned)A.g.N + 1u, gsi = ((unsigned)A.g.K + 1u) * gsj;
#pragma unroll
        for (int dir = 0; dir < 3; dir++)
        {
            constexpr unsigned NB1 = BR + 1, NB2 = BR + 2;
            const unsigned nj = dir == 1 ? NB2 : NB1, nk = dir == 2 ? NB2 : NB1;
            const double* src = dir == 0 ? A.g.gx : dir == 1 ? A.g.gy : A.g.gz;
            if (t < (unsigned)TILE_X)
            {
                const unsigned a = dir == 0 ? t / (NB1 * NB1) : dir == 1 ? t / (NB2 * NB1) : t / (NB1 * NB2);
                const unsigned rem = t - a * nj * nk;
                const unsigned bb = dir == 2 ? rem / NB2 : rem / NB1, cc = rem - bb * nk;
                const unsigned gi = min((unsigned)ci0 + a, (unsigned)A.g.M), gj = min((unsigned)cj0 + bb, (unsigned)A.g.K), gk = min((unsigned)ck0 + cc, (unsigned)A.g.N);
                S.tile[dir][t] = __ldg(src + (gi * gsi + gj * gsj + gk));
            }
        }
    }
    const double dt = A.s.dt;
    unsigned out = 0, removed = 0, parity = 0;
    for (unsigned in = 0; in < n; in += BK_CHUNK, parity ^= 1u)
    {
        const unsigned m = min((unsigned)BK_CHUNK, n - in);
        if (t < BCELLS + 4) S.cnt[t] = 0;
        if (t == 0) S.n_hit = 0;
        while (!mbar_try_wait(&S.bar, parity)) {}
        __syncthreads();
        // ---- pass 1: gather, push, boundary, class + rank inside the class
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (MCC)
        {
            Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)(off + in + t));
            rnd = rng.block();
        }
#pragma unroll 1
        for (int it = 0; it < BK_PPT; it++)
        {
            const unsigned p = t + (unsigned)it * BK_THREADS;
            if (p >= m) break;
            double x = S.P[0][p], y = S.P[1][p], z = S.P[2][p];
            unsigned char cls = CLS_DEAD;
            if (particle_alive(x))
            {
                double vx = S.P[3][p], vy = S.P[4][p], vz = S.P[5][p];
                const double X = x * A.g.idx, Y = y * A.g.idy, Z = z * A.g.idz;
                const int oi = (int)X - ci0, oj = (int)Y - cj0, ok = (int)Z - ck0;
                double Ex, Ey, Ez;
                if ((unsigned)oi < (unsigned)BR && (unsigned)oj < (unsigned)BR && (unsigned)ok < (unsigned)BR)
                {
                    Ex = -grad_tile<0>(A.g, S.tile[0], X, Y, Z, ci0, cj0, ck0);
                    Ey = -grad_tile<1>(A.g, S.tile[1], X, Y, Z, ci0, cj0, ck0);
                    Ez = -grad_tile<2>(A.g, S.tile[2], X, Y, Z, ci0, cj0, ck0);
                }
                else
                {
                    // a guest (it found its own bin full) or a particle exactly on the far face of the box
                    Ex = -grad_component<0>(A.g, X, Y, Z);
                    Ey = -grad_component<1>(A.g, X, Y, Z);
                    Ez = -grad_component<2>(A.g, X, Y, Z);
                }
                vx += Ex * A.s.hq;
                vy += Ey * A.s.hq;
                vz += Ez * A.s.hq;
                if (HASB)
                {
                    const double px = vx - vy * A.s.tz + vz * A.s.ty;
                    const double py = vy - vz * A.s.tx + vx * A.s.tz;
                    const double pz = vz - vx * A.s.ty + vy * A.s.tx;
                    const double ox = vx, oy = vy, oz = vz;
                    vx = ox - py * A.s.sz + pz * A.s.sy;
                    vy = oy - pz * A.s.sx + px * A.s.sz;
                    vz = oz - px * A.s.sy + py * A.s.sx;
                }
                vx += Ex * A.s.hq;
                vy += Ey * A.s.hq;
                vz += Ez * A.s.hq;
                x += vx * dt;
                y += vy * dt;
                z += vz * dt;
                int i, j, k;
                if (boundary3_ijk(A.g, x, y, z, i, j, k))
                {
                    const int li = i - ci0, lj = j - cj0, lk = k - ck0;
                    if ((unsigned)li < (unsigned)BR && (unsigned)lj < (unsigned)BR && (unsigned)lk < (unsigned)BR)
                        cls = (unsigned char)((li * BR + lj) * BR + lk);
                    else
                    {
                        cls = CLS_GUEST;
                        S.dstb[p] = (unsigned)(((i / BR) * B.nby + (j / BR)) * B.nbz + (k / BR));
                        if (DEPOSIT)
                        {
                            unsigned long long w[8];
                            weights3(A.g, x, y, z, w);
                            scatter3(A.g, (unsigned)(((size_t)i * A.g.K + j) * A.g.N + k), w);
                        }
                    }
                    const unsigned rk = atomicAdd(&S.cnt[cls], 1u);
                    S.rank[p] = (unsigned short)rk;
                    if (cls == CLS_GUEST && rk < (unsigned)LIST_CAP) S.leave[rk] = (unsigned short)p;
                    if (MCC)
                    {
                        const unsigned word = it == 0 ? rnd.x : it == 1 ? rnd.y : rnd.z;
                        if ((unsigned long long)word < A.s.prob_u32)
                        {
                            const unsigned h = atomicAdd(&S.n_hit, 1u);
                            if (h < (unsigned)LIST_CAP) S.hits[h] = (unsigned short)p;
                            else cls |= 128;              // list full: appended one by one below
                        }
                    }
                    S.P[1][p] = y; S.P[2][p] = z;
                    S.P[3][p] = vx; S.P[4][p] = vy; S.P[5][p] = vz;
                }
                else
                {
                    removed++;
                    x = dead_marker();
                }
                S.P[0][p] = x;
            }
            S.cls[p] = cls;
        }
        __syncthreads();
        // ---- exclusive scan of the 65 class counts (warp 0)
        if (t < 32)
        {
            const unsigned a0 = S.cnt[2 * t], a1 = S.cnt[2 * t + 1], a2 = t == 0 ? S.cnt[64] : 0u;
            unsigned incl = a0 + a1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned v = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
                if (lane >= (unsigned)o) incl += v;
            }
            const unsigned excl = incl - a0 - a1;
            S.start[2 * t] = excl;
            S.start[2 * t + 1] = excl + a0;
            const unsigned total = __shfl_sync(MAG2D_FULL_MASK, incl, 31);
            if (t == 0)
            {
                S.start[64] = total;
                S.start[65] = total + a2;
            }
        }
        __syncthreads();
        const unsigned n_stay = S.start[64], n_leave = S.start[65] - S.start[64], n_hit = MCC ? min(S.n_hit, (unsigned)LIST_CAP) : 0u;
        if (MCC && t == 32 && n_hit) S.hit_base = atomicAdd(A.coll_count, n_hit);
        // ---- sorted index of the stayers (perm[sorted position] = staged slot)
#pragma unroll 1
        for (int it = 0; it < BK_PPT; it++)
        {
            const unsigned p = t + (unsigned)it * BK_THREADS;
            if (p >= m) break;
            const unsigned cls = S.cls[p] & 127u;
            if (cls < (unsigned)BCELLS) S.perm[S.start[cls] + S.rank[p]] = (unsigned short)p;
        }
        __syncthreads();
        // room in the outbox for this pass' leavers: one returning atomic per CTA, issued here and consumed after the deposit
        unsigned ob_base = 0;
        if (t == 0 && n_leave) ob_base = atomicAdd(B.ob.count, n_leave);
        // ---- charge of the stayers: four threads per cell walk the cell's particles, registers only
        if (DEPOSIT)
        {
            const unsigned c = t >> 2, q = t & 3u;
            unsigned long long acc[8];
#pragma unroll
            for (int e = 0; e < 8; e++) acc[e] = 0ULL;
            const unsigned s0 = S.start[c], s1 = S.start[c + 1];
            for (unsigned r = s0 + q; r < s1; r += 4)
            {
                const unsigned p = S.perm[r];
                unsigned long long w[8];
                weights3(A.g, S.P[0][p], S.P[1][p], S.P[2][p], w);
#pragma unroll
                for (int e = 0; e < 8; e++) acc[e] += w[e];
            }
#pragma unroll
            for (int e = 0; e < 8; e++)
            {
                acc[e] += __shfl_xor_sync(MAG2D_FULL_MASK, acc[e], 1);
                acc[e] += __shfl_xor_sync(MAG2D_FULL_MASK, acc[e], 2);
            }
            if (q == 0)
            {
#pragma unroll
                for (int e = 0; e < 8; e++) S.cellsum[e][c] = in == 0 ? acc[e] : S.cellsum[e][c] + acc[e];
            }
        }
        if (t == 0 && n_leave) S.ob_base = ob_base;
        __syncthreads();
        // ---- leavers go to the outbox (all of this pass' leavers or none: a CTA that does not get room keeps them as guests)
        const bool leavers_go = n_leave > 0 && S.ob_base + n_leave <= B.ob.cap;
        if (n_leave)
        {
            if (!leavers_go && t == 0) atomicAdd(B.ob.count + 2, 1u);
            // the first LIST_CAP leavers are listed by rank; a pass with more walks its slots for the rest
            for (unsigned r = t; r < n_leave; r += BK_THREADS)
            {
                unsigned p;
                if (r < (unsigned)LIST_CAP) p = S.leave[r];
                else
                {
                    p = 0;
                    for (unsigned k = 0; k < m; k++)
                        if ((S.cls[k] & 127u) == (unsigned)CLS_GUEST && S.rank[k] == r) { p = k; break; }
                }
                const unsigned q = S.ob_base + r;
                if (leavers_go)
                {
#pragma unroll
                    for (int a = 0; a < 6; a++) B.ob.arr[a][q] = S.P[a][p];
                    B.ob.dst[q] = S.dstb[p];
                    if (MCC && (S.cls[p] & 128)) A.coll_list[atomicAdd(A.coll_count, 1u)] = 0x80000000u | q;
                    S.P[0][p] = dead_marker();
                    S.cls[p] = CLS_DEAD;
                }
                else if (q < B.ob.cap) B.ob.dst[q] = OUTBOX_VOID;
            }
        }
        if (!compact)
        {
            // ---- in place: the staged slots go back where they came from (holes included) with one bulk store per array
            if (MCC)
            {
                if (t >= 128 && t - 128 < n_hit)
                {
                    const unsigned p = S.hits[t - 128];
                    // a leaver collides where it is now: in the outbox (high bit: k_mcc_collide3d reads A.dst, the outbox view)
                    A.coll_list[S.hit_base + (t - 128)] = S.cls[p] != CLS_DEAD ? off + in + p : 0x80000000u | (S.ob_base + S.rank[p]);
                }
                if (S.n_hit > (unsigned)LIST_CAP)
                {
                    for (unsigned p = t; p < m; p += BK_THREADS)
                        if (S.cls[p] != CLS_DEAD && (S.cls[p] & 128)) A.coll_list[atomicAdd(A.coll_count, 1u)] = off + in + p;
                }
            }
            fence_proxy_async();
            __syncthreads();
            if (t == 0)
            {
                const unsigned bytes = chunk_bytes(in);
#pragma unroll
                for (int a = 0; a < 6; a++) bulk_store(arr[a] + off + in, S.P[a], bytes);
                bulk_commit();
                bulk_wait_read_all();           // shared memory may be overwritten (next pass) or released (exit)
                if (in + BK_CHUNK < n) issue_load(in + BK_CHUNK);
            }
            out += m;
            // the last pass of a bin: nobody writes the staging arrays any more, so only thread 0 waits for the bulk stores
            // to have read them; everybody else goes on to the node flush
            if (in + BK_CHUNK < n) __syncthreads();
            continue;
        }
        __syncthreads();
        // ---- compacting pass: the stayers (and guests that could not leave) are written back cell-sorted, holes dropped
        const unsigned kept = leavers_go ? n_stay : n_stay + n_leave;
#pragma unroll 1
        for (int it = 0; it < BK_PPT; it++)
        {
            const unsigned p = t + (unsigned)it * BK_THREADS;
            if (p >= m) break;
            const unsigned char cf = S.cls[p];
            if (cf == CLS_DEAD) continue;
            const unsigned dest = S.start[cf & 127u] + S.rank[p];
            const size_t d = (size_t)off + out + dest;
#pragma unroll
            for (int a = 0; a < 6; a++) arr[a][d] = S.P[a][p];
            if (MCC && (cf & 128)) A.coll_list[atomicAdd(A.coll_count, 1u)] = off + out + dest;
        }
        if (MCC && t >= 128 && t - 128 < n_hit)
        {
            const unsigned p = S.hits[t - 128];
            const unsigned char cf = S.cls[p];
            A.coll_list[S.hit_base + (t - 128)] = cf != CLS_DEAD ? off + out + S.start[cf & 127u] + S.rank[p] : 0x80000000u | (S.ob_base + S.rank[p]);
        }
        out += kept;
        fence_proxy_async();
        __syncthreads();
        if (t == 0 && in + BK_CHUNK < n) issue_load(in + BK_CHUNK);
    }
    if (compact)
    {
        // the vacated tail of the bin becomes holes; the new fill level
        for (unsigned k = out + t; k < n; k += BK_THREADS) A.p.x[(size_t)off + k] = dead_marker();
        if (t == 0) B.bin_cnt[b] = out;
    }
    if (DEPOSIT)
    {
        // one RED per node of the brick: node (ni, nj, nk) collects corner (a, b, c) of cell (ni - a, nj - b, nk - c)
        if (t < (BR + 1) * (BR + 1) * (BR + 1))
        {
            const int ni = (int)t / ((BR + 1) * (BR + 1)), nj = ((int)t / (BR + 1)) % (BR + 1), nk = (int)t % (BR + 1);
            unsigned long long sum = 0ULL;
#pragma unroll
            for (int e = 0; e < 8; e++)
            {
                const int li = ni - (e & 1), lj = nj - ((e >> 1) & 1), lk = nk - (e >> 2);
                if ((unsigned)li < (unsigned)BR && (unsigned)lj < (unsigned)BR && (unsigned)lk < (unsigned)BR) sum += S.cellsum[e][(li * BR + lj) * BR + lk];
            }
            if (sum) atomicAdd(A.g.rho + (((size_t)(ci0 + ni) * A.g.K + (cj0 + nj)) * A.g.N + (ck0 + nk)), sum);
        }
    }
    if (__any_sync(MAG2D_FULL_MASK, removed != 0))
    {
        unsigned rsum = removed;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(MAG2D_FULL_MASK, rsum, o);
        if (lane == 0) atomicAdd(A.removed, (unsigned long long)rsum);
    }
}

// the outbox empties into the bins: every leaver takes the next free slot of its new brick's bin.  A full bin sends it on to
// the following bins (any bin will do: a particle outside its bin's brick is a guest and goes through the global-memory path
// until it leaves again), so nothing is ever dropped; the host re-bins with more slack when that happened.
__global__ void __launch_bounds__(256) k_place3d(const __grid_constant__ BrickArgs B)
{
    const unsigned n = min(B.ob.count[0], B.ob.cap);
    const unsigned nb = (unsigned)(B.nbx * B.nby * B.nbz);
    double* const arr[6] = {B.A.p.x, B.A.p.y, B.A.p.z, B.A.p.vx, B.A.p.vy, B.A.p.vz};
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
    {
        unsigned dst = B.ob.dst[q];
        if (dst == OUTBOX_VOID) continue;
        for (unsigned probe = 0; probe < nb; probe++)
        {
            const unsigned lo = B.bin_off[dst], cap = B.bin_off[dst + 1] - lo;
            const unsigned pos = atomicAdd(&B.bin_cnt[dst], 1u);
            if (pos < cap)
            {
                const size_t d = (size_t)lo + pos;
#pragma unroll
                for (int a = 0; a < 6; a++) arr[a][d] = B.ob.arr[a][q];
                break;
            }
            atomicSub(&B.bin_cnt[dst], 1u);
            if (probe == 0) atomicAdd(B.ob.count + 1, 1u);
            dst = dst + 1 < nb ? dst + 1 : 0;
        }
    }
}

// ---- (re)binning of a store in any order ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned brick_of(const Grid3Dev& g, double x, double y, double z, int nby, int nbz)
{
    const int i = max(min((int)__dmul_rn(x, g.idx), g.M - 2), 0), j = max(min((int)__dmul_rn(y, g.idy), g.K - 2), 0),
              k = max(min((int)__dmul_rn(z, g.idz), g.N - 2), 0);
    return (unsigned)(((i / BR) * nby + (j / BR)) * nbz + (k / BR));
}

__global__ void k_brick_count(Grid3Dev g, ParticlesDev p, int nby, int nbz, unsigned* __restrict__ count)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = k < p.n ? p.x[k] : dead_marker();
    const bool alive = particle_alive(x);
    const unsigned b = alive ? brick_of(g, x, p.y[k], p.z[k], nby, nbz) : 0u;
    warp_count(count, alive, b);
}

// capacities with slack and their exclusive scan (one CTA walks the bricks: a re-binning is rare)
// Bin capacity: the fill of a bin fluctuates like a Poisson variable around the LOCAL density (arrivals and departures
// balance on average), so a bin that happens to be low at the re-binning drifts back up: the capacity follows
// m = max(own count, mean of the 3 x 3 x 3 neighbourhood) with room for 6 sqrt(m) + slack m + 32 more, padded to 32 slots.
__global__ void __launch_bounds__(1024) k_brick_layout(const unsigned* __restrict__ count, int nbx, int nby, int nbz, double slack, unsigned* __restrict__ off,
                                                       unsigned* __restrict__ cursor, unsigned long long* __restrict__ total)
{
    const int nb = nbx * nby * nbz;
    __shared__ unsigned carry;
    __shared__ unsigned warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024)
    {
        const int k = base + (int)threadIdx.x;
        unsigned v = 0;
        if (k < nb)
        {
            const unsigned c = count[k];
            const int bk = k % nbz, bj = (k / nbz) % nby, bi = k / (nbz * nby);
            unsigned long long sum = 0;
            int cells = 0;
            for (int a = max(bi - 1, 0); a <= min(bi + 1, nbx - 1); a++)
                for (int b = max(bj - 1, 0); b <= min(bj + 1, nby - 1); b++)
                    for (int d = max(bk - 1, 0); d <= min(bk + 1, nbz - 1); d++, cells++) sum += count[(a * nby + b) * nbz + d];
            const double m = fmax((double)c, (double)sum / cells);
            v = ((unsigned)(m + m * slack + 6.0 * sqrt(m)) + 32u + 31u) & ~31u;
            cursor[k] = 0;
        }
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const unsigned tt = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
            if (lane >= (unsigned)o) incl += tt;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned tt = __shfl_up_sync(MAG2D_FULL_MASK, w, o);
                if (lane >= (unsigned)o) w += tt;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned excl = incl - v + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (k < nb) off[k] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        off[nb] = carry;
        *total = carry;
    }
}

__global__ void k_fill_nan(double* __restrict__ x, long long n)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) x[k] = dead_marker();
}

struct Perm6
{
    const double* src[6];
    double* dst[6];
};

__global__ void k_brick_scatter(Grid3Dev g, long long n, const __grid_constant__ Perm6 P, int nby, int nbz, const unsigned* __restrict__ off, unsigned* __restrict__ cursor)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = k < n ? P.src[0][k] : dead_marker();
    const bool alive = particle_alive(x);
    const unsigned b = alive ? brick_of(g, x, P.src[1][k], P.src[2][k], nby, nbz) : 0u;
    const unsigned pos = warp_ticket(cursor, alive, b);
    if (alive)
    {
        const size_t d = (size_t)off[b] + pos;
#pragma unroll
        for (int a = 0; a < 6; a++) P.dst[a][d] = P.src[a][k];
    }
}

}  // namespace

void brick_free(SpeciesStore& S)
{
    cudaFree(S.d_bin_off);
    cudaFree(S.d_bin_cnt);
    cudaFree(S.d_bin_scratch);
    for (int a = 0; a < 6; a++) { cudaFree(S.d_outbox[a]); S.d_outbox[a] = nullptr; }
    cudaFree(S.d_ob_dst);
    cudaFree(S.d_ob_count);
    if (S.h_bin_flags) cudaFreeHost(S.h_bin_flags);
    if (S.ev_bin) cudaEventDestroy(S.ev_bin);
    S.d_bin_off = S.d_bin_cnt = S.d_bin_scratch = S.d_ob_dst = S.d_ob_count = nullptr;
    S.ob_cap = 0;
    S.h_bin_flags = nullptr;
    S.ev_bin = nullptr;
    S.bins_valid = false;
    S.bin_flags_pending = false;
}

// bin the store of species s by brick into the other slab (with slack behind every bin) and make that slab current
int brick_rebuild(mag2d_ctx* c, int s, const Grid3Dev& g)
{
    SpeciesStore& S = c->sp[s];
    const int nbx = (g.M - 1 + BR - 1) / BR, nby = (g.K - 1 + BR - 1) / BR, nbz = (g.N - 1 + BR - 1) / BR;
    const int nb = nbx * nby * nbz;
    if (S.bin_nb != nb)
    {
        cudaFree(S.d_bin_off); cudaFree(S.d_bin_cnt); cudaFree(S.d_bin_scratch);
        S.d_bin_off = S.d_bin_cnt = S.d_bin_scratch = nullptr;
        CUDA_OK(cudaMalloc(&S.d_bin_off, sizeof(unsigned) * (size_t)(nb + 1)));
        CUDA_OK(cudaMalloc(&S.d_bin_cnt, sizeof(unsigned) * (size_t)nb));
        CUDA_OK(cudaMalloc(&S.d_bin_scratch, sizeof(unsigned) * (size_t)nb + 16));
        S.bin_nb = nb;
    }
    if (!S.d_ob_count)
    {
        CUDA_OK(cudaMalloc(&S.d_ob_count, sizeof(unsigned) * 4));
        CUDA_OK(cudaMallocHost(&S.h_bin_flags, sizeof(unsigned) * 4));
        CUDA_OK(cudaEventCreateWithFlags(&S.ev_bin, cudaEventDisableTiming));
    }
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(S.d_bin_scratch + ((nb + 1) / 2 * 2));
    ParticlesDev p;
    memset(&p, 0, sizeof(p));
    double* const* cur = S.arr[S.cur];
    p.x = cur[ARR_X]; p.y = cur[ARR_Y]; p.z = cur[ARR_Z];
    p.n = S.n_slots;
    const unsigned pblocks = (unsigned)((S.n_slots + 255) / 256);
    CUDA_OK(cudaMemsetAsync(S.d_bin_scratch, 0, sizeof(unsigned) * (size_t)nb, c->stream));
    k_brick_count<<<pblocks, 256, 0, c->stream>>>(g, p, nby, nbz, S.d_bin_scratch);
    k_brick_layout<<<1, 1024, 0, c->stream>>>(S.d_bin_scratch, nbx, nby, nbz, S.bin_slack, S.d_bin_off, S.d_bin_cnt, d_total);
    unsigned long long total = 0;
    CUDA_OK(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (total >= 0x7FFFFFF0ULL) { mag2d_set_error("brick_rebuild: more than 2^31 slots per species and GPU"); return 1; }
    const long long need = (long long)total;
    if (need > S.capacity)
    {
        // grow the current slab (keeps the particles), then give the idle slab the same size
        if (mag2d_reserve(c, s, need)) return 1;
    }
    if (!S.arr[S.cur ^ 1][ARR_X] && store_alloc_slab(c, S, S.cur ^ 1, S.capacity)) return 1;
    cur = S.arr[S.cur];
    double* const* oth = S.arr[S.cur ^ 1];
    Perm6 P;
    const int order[6] = {ARR_X, ARR_Y, ARR_Z, ARR_VX, ARR_VY, ARR_VZ};
    for (int a = 0; a < 6; a++) { P.src[a] = cur[order[a]]; P.dst[a] = oth[order[a]]; }
    k_fill_nan<<<148 * 8, 256, 0, c->stream>>>(oth[ARR_X], need);
    k_brick_scatter<<<pblocks, 256, 0, c->stream>>>(g, S.n_slots, P, nby, nbz, S.d_bin_off, S.d_bin_cnt);
    c->launches += 4;
    CUDA_OK(cudaGetLastError());
    // the outbox takes one step's leavers: an eighth of the slots (a CTA whose leavers do not fit keeps them as guests)
    const long long ob_cap = std::max<long long>(need / 8, 4096);
    if (S.ob_cap < ob_cap)
    {
        for (int a = 0; a < 6; a++)
        {
            cudaFree(S.d_outbox[a]);
            S.d_outbox[a] = nullptr;
            CUDA_OK(cudaMalloc(&S.d_outbox[a], sizeof(double) * (size_t)ob_cap));
        }
        cudaFree(S.d_ob_dst);
        S.d_ob_dst = nullptr;
        CUDA_OK(cudaMalloc(&S.d_ob_dst, sizeof(unsigned) * (size_t)ob_cap));
        S.ob_cap = ob_cap;
    }
    CUDA_OK(cudaMemsetAsync(S.d_ob_count, 0, sizeof(unsigned) * 4, c->stream));
    CUDA_OK(cudaMemsetAsync(S.d_removed, 0, sizeof(unsigned long long), c->stream));
    S.cur ^= 1;
    S.n_slots = need;
    S.bins_valid = true;
    S.bin_flags_pending = false;
    S.bin_overflow_seen = 0;
    S.pushes_since_compact = 0;
    S.tickets_valid = false;
    S.steps_since_sort = 0;
    S.append_epoch++;
    S.rebinnings++;
    if (getenv("MAG3D_BRICK_DEBUG")) fprintf(stderr, "brick: re-binned species %d: %lld slots in %d bins, slack %.2f (#%lld)\n", s, need, nb, S.bin_slack, S.rebinnings);
    return 0;
}

static void brick_args(const mag2d_ctx* c, SpeciesStore& S, const Push3Args& A, BrickArgs& B)
{
    memset(&B, 0, sizeof(B));
    B.A = A;
    B.bin_off = S.d_bin_off;
    B.bin_cnt = S.d_bin_cnt;
    B.nbx = (A.g.M - 1 + BR - 1) / BR;
    B.nby = (A.g.K - 1 + BR - 1) / BR;
    B.nbz = (A.g.N - 1 + BR - 1) / BR;
    for (int a = 0; a < 6; a++) B.ob.arr[a] = S.d_outbox[a];
    B.ob.dst = S.d_ob_dst;
    B.ob.count = S.d_ob_count;
    B.ob.cap = (unsigned)S.ob_cap;
    (void)c;
}

// the outbox as a particle view: what k_mcc_collide3d reads for list entries with the high bit set (Push3Args::dst)
void brick_outbox_view(const SpeciesStore& S, ParticlesDev& v)
{
    memset(&v, 0, sizeof(v));
    v.x = S.d_outbox[0]; v.y = S.d_outbox[1]; v.z = S.d_outbox[2];
    v.vx = S.d_outbox[3]; v.vy = S.d_outbox[4]; v.vz = S.d_outbox[5];
    v.n = S.ob_cap;
}

// one step of species s on its binned store: A is the argument block launch_species_advance3d has prepared.  Every
// compact_every-th step writes the bins back compacted and cell-sorted; the steps in between store them in place.
int launch_brick_push(mag2d_ctx* c, int s, const Push3Args& A, bool mcc, bool deposit, int compact_every)
{
    SpeciesStore& S = c->sp[s];
    BrickArgs B;
    brick_args(c, S, A, B);
    B.compact = S.pushes_since_compact + 1 >= compact_every ? 1 : 0;
    S.pushes_since_compact = B.compact ? 0 : S.pushes_since_compact + 1;
    CUDA_OK(cudaMemsetAsync(S.d_ob_count, 0, sizeof(unsigned), c->stream));
    const int smem = (int)sizeof(BrickSmem);
    static bool attr_set = false;
    if (!attr_set)
    {
#define SETATTR(Bm, Mc, D) CUDA_OK(cudaFuncSetAttribute(k_push3d_brick<Bm, Mc, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
        SETATTR(false, false, false); SETATTR(false, false, true); SETATTR(false, true, false); SETATTR(false, true, true);
        SETATTR(true, false, false); SETATTR(true, false, true); SETATTR(true, true, false); SETATTR(true, true, true);
#undef SETATTR
        attr_set = true;
    }
    const dim3 blocks((unsigned)B.nbz, (unsigned)B.nby, (unsigned)B.nbx);
    const int code = (A.s.has_B ? 4 : 0) | (mcc ? 2 : 0) | (deposit ? 1 : 0);
#define LB(Bm, Mc, D) k_push3d_brick<Bm, Mc, D><<<blocks, BK_THREADS, smem, c->stream>>>(B)
    switch (code)
    {
        case 0: LB(false, false, false); break;
        case 1: LB(false, false, true); break;
        case 2: LB(false, true, false); break;
        case 3: LB(false, true, true); break;
        case 4: LB(true, false, false); break;
        case 5: LB(true, false, true); break;
        case 6: LB(true, true, false); break;
        default: LB(true, true, true); break;
    }
#undef LB
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// after the collision pass: the outbox empties into the bins, and the overflow counters go home (adopted by a later step, no sync)
int launch_brick_migrate(mag2d_ctx* c, int s, const Push3Args& A)
{
    SpeciesStore& S = c->sp[s];
    BrickArgs B;
    brick_args(c, S, A, B);
    k_place3d<<<148 * 8, 256, 0, c->stream>>>(B);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    if (!S.bin_flags_pending)
    {
        CUDA_OK(cudaMemcpyAsync(S.h_bin_flags, S.d_ob_count, sizeof(unsigned) * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaEventRecord(S.ev_bin, c->stream));
        S.bin_flags_pending = true;
    }
    return 0;
}

// a readback of the overflow counters has landed: arrivals that found their bin full (or CTAs whose leavers did not fit the
// outbox) mean the bins are too tight
void brick_poll_overflow(SpeciesStore& S)
{
    if (!S.bin_flags_pending || cudaEventQuery(S.ev_bin) != cudaSuccess) return;
    S.bin_flags_pending = false;
    const unsigned full_bins = S.h_bin_flags[1], no_room = S.h_bin_flags[2];
    if (no_room + full_bins > S.bin_overflow_seen)
    {
        if (getenv("MAG3D_BRICK_DEBUG")) fprintf(stderr, "brick: overflow flags outbox=%u bins=%u (seen %u) -> re-bin with slack %.2f\n", no_room, full_bins, S.bin_overflow_seen, S.bin_slack * 1.5 + 0.1);
        S.bin_overflow_seen = no_room + full_bins;
        S.bin_slack = std::min(S.bin_slack * 1.5 + 0.1, 4.0);
        S.bins_valid = false;            // re-bin with more slack at the next step
    }
}
