// push.cu — fused particle kernels: field gather + Boris push + null-collision MCC + boundary /
// electrode absorption + fixed-point CIC deposit, for Cartesian and cylindrical 2D3V species; the
// multi-collision free-flight mover; the half-step-back initialisation; loaders and diagnostics.
//
// Reference loops replaced (paths relative to the reference checkout):
//   Species<CARTESIAN>::advance_boris      src/particles.cpp:924-995
//   Species<CYLINDRICAL>::advance_boris    src/particles.cpp:539-621
//   Species<CARTESIAN>::advance_multicoll  src/particles.cpp:812-859
//   Species<D>::advance_boris_init         src/particles.cpp:997-1050 / 623-679
//   Species<D>::advance_boundary           src/particles.hpp:370-411
//   Field2D::grad / Fields::E              src/Field2D.hpp:80-168, src/fields.hpp:124-150
//   Field2D::accumulate                    src/Field2D.hpp:45-62
// The reference makes two passes over an AoS array per species and step; here one kernel reads and
// writes each live phase-space component once (80 B per particle-step, 2D3V fp64).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ctx.hpp"
#include "mcc.cuh"

namespace {

constexpr int PUSH_THREADS = 256;
#ifndef MAG2D_DEPOSIT_SMALL
#define MAG2D_DEPOSIT_SMALL 2   // measured on C4: 1.91 ms (2) vs 1.97 ms (off), 1.92 ms (3)
#endif

struct PushArgs
{
    GridDev g;
    SpeciesDev s;
    ParticlesDev p;
    const MccBlob* mcc;
    unsigned long long* counts;    // collision counters or nullptr
    unsigned long long* removed;   // removal counter
    unsigned long long seed;
    unsigned* coll_list;           // slots whose Bernoulli test fired this step (processed by k_mcc_collide)
    unsigned* coll_count;
    // cell sort fused into the step (SORTING kernels; sort.cu describes the pipeline)
    int permute;                   // write every array to its sorted slot of the other slab (keys of an earlier COUNT step)
    int count;                     // count every surviving particle in its new cell for the next permuting step
    int cell_cols;                 // N - 1
    unsigned* cursor;              // [cell] next free sorted slot of the cell (the scanned counts of the last COUNT push)
    ParticlesDev dst;              // the other slab
    unsigned* count_out;           // [cell]
};

// ---- gather: E = -grad(ue), the staggered-difference bilinear form of Field2D::grad ----------------
// The reference differences the potential around every particle; here the differences live in the
// precomputed edge fields gx/gz (same operations, same rounding) and the particle only interpolates:
// 8 loads instead of 12.  Edge cases of the reference (i == 0, i == jmax-1, j == 0, j == lmax-1) are
// folded into the general formula by ghost rows / columns that repeat the first / last difference: in the
// first and last HALF cell of each axis the reference evaluates the one-sided form g (1 - fy), here the
// same value comes out as g cx cy + g fx cy, which agrees to one or two ulp, not bit for bit (everywhere
// else the operations and their order are the reference's).  Float->int conversions are the expensive part on this pipe mix, so each axis is
// converted once: (int)(X + 0.5) is derived from (int)X and the fraction.  The two differ only when X
// lies within one ulp below a half-integer, where both stencils interpolate the same edge value.
__device__ __forceinline__ void gather_E(const GridDev& g, double x, double z, double& Ex, double& Ez)
{
    // gx / gz carry one ghost row and column on every side that the stencil can reach ([M+1][N+1], k_edge_fields):
    // the copies of the first / last difference make the reference's one-sided edge rules (i == 0, i == jmax-1,
    // j == 0, j == lmax-1 of Field2D::grad) fall out of the general formula, so no per-particle edge tests remain.
    // (Tests of a clamped integer against its bound are also what ptxas 12.9 miscompiled on sm_100a, see DESIGN.md.)
    const int M = g.M, N = g.N;
    const unsigned ld = (unsigned)N + 1u;
    const double X = x * g.idx, Y = z * g.idz;
    int ix = (int)X, jy = (int)Y;
    ix = max(min(ix, M - 1), 0);
    jy = max(min(jy, N - 1), 0);
    const double dix = (double)ix, djy = (double)jy;
    const double fxc = X - dix, fyc = Y - djy;
    const bool upx = fxc >= 0.5, upy = fyc >= 0.5;
    {
        // x component: rows (int)(X + 0.5) and the next one, columns (int)Y and the next one
        const int i = min(ix + (upx ? 1 : 0), M - 1);
        const double fx = X - (upx ? dix + 1.0 : dix) + .5;
        const double* r1 = g.gx + ((unsigned)i * ld + (unsigned)jy);
        const double g1 = __ldg(r1), g2 = __ldg(r1 + 1), g4 = __ldg(r1 + ld), g3 = __ldg(r1 + ld + 1);
        const double cx = 1 - fx, cy = 1 - fyc;
        Ex = -(g1 * cx * cy + g2 * cx * fyc + g3 * fx * fyc + g4 * fx * cy);
    }
    {
        // z component: rows (int)X and the next one, columns (int)(Y + 0.5) and the next one
        const int j = min(jy + (upy ? 1 : 0), N - 1);
        const double fy = Y - (upy ? djy + 1.0 : djy) + 0.5;
        const double* r0 = g.gz + ((unsigned)ix * ld + (unsigned)j);
        const double g1 = __ldg(r0), g4 = __ldg(r0 + 1), g2 = __ldg(r0 + ld), g3 = __ldg(r0 + ld + 1);
        const double cx = 1 - fxc, cy = 1 - fy;
        Ez = -(g1 * cx * cy + g2 * cy * fxc + g3 * fxc * fy + g4 * fy * cx);
    }
}

// ---- Boris rotation + half accelerations (velocity part of the movers) ---------------------------
// BMODE: 0 no magnetic field, 1 constant B (t and s precomputed per species), 2 B interpolated per particle from the
// table of Fields::load_magnetic_field (t and s formed per particle with the reference's expression order)
constexpr int B_NONE = 0, B_CONST = 1, B_TABLE = 2;

// Field2D::interpolate (Field2D.hpp:64-78) on the magnetic field table.  The reference throws outside the table;
// mag2d_set_magnetic_field only accepts tables that cover the whole box, so the index clamp never changes a result
// (it only keeps x == table edge from reading one row past the end, which the reference does with weight zero).
__device__ __forceinline__ double table_interpolate(const GridDev& g, const double* __restrict__ data, double x, double z)
{
    x -= g.bxmin;
    z -= g.bzmin;
    const int i = max(min((int)(x * g.bidx), g.bM - 2), 0), j = max(min((int)(z * g.bidz), g.bN - 2), 0);
    const double u = x * g.bidx - i, v = z * g.bidz - j;
    const double* r = data + (unsigned)i * (unsigned)g.bN + (unsigned)j;
    return (1 - u) * (1 - v) * __ldg(r) + u * (1 - v) * __ldg(r + g.bN) + (1 - u) * v * __ldg(r + 1) + u * v * __ldg(r + g.bN + 1);
}

template <int COORD>
__device__ __forceinline__ void boris_rotate(double tx, double ty, double tz, double sx, double sy, double sz, double& vx, double& vy,
                                             double& vz)
{
    if (COORD == MAG2D_CYLINDRICAL)
    {
        // textbook orientation in (r, theta, z), particles.cpp:581-592
        const double pr = vx + vy * tz - vz * ty;
        const double pt = vy + vz * tx - vx * tz;
        const double pz = vz + vx * ty - vy * tx;
        vx = vx + pt * sz - pz * sy;
        vy = vy + pz * sx - pr * sz;
        vz = vz + pr * sy - pt * sx;
    }
    else
    {
        // right-handed (x, z, y) triad: opposite signs, particles.cpp:964-979
        const double pr = vx - vy * tz + vz * ty;
        const double pz = vz - vx * ty + vy * tx;
        const double pt = vy - vz * tx + vx * tz;
        vx = vx - pt * sz + pz * sy;
        vy = vy - pz * sx + pr * sz;
        vz = vz - pr * sy + pt * sx;
    }
}

// rotation about the table's B at (x, z): Fields::B gives (Br, Bz, 0) (fields.hpp:172-175)
template <int COORD>
__device__ __forceinline__ void boris_rotate_table(const GridDev& g, const SpeciesDev& s, double x, double z, double& vx, double& vy,
                                                   double& vz)
{
    const double tx = table_interpolate(g, g.b_r, x, z) * s.tb;
    const double ty = 0.0 * s.tb;
    const double tz = table_interpolate(g, g.b_z, x, z) * s.tb;
    const double tmp = 2.0 / (1 + tx * tx + ty * ty + tz * tz);
    boris_rotate<COORD>(tx, ty, tz, tx * tmp, ty * tmp, tz * tmp, vx, vy, vz);
}

template <int COORD, int BMODE>
__device__ __forceinline__ void boris_velocity(const GridDev& g, const SpeciesDev& s, double x, double z, double Ex, double Ez, double& vx,
                                               double& vy, double& vz)
{
    vx += Ex * s.hq;
    vz += Ez * s.hq;
    if (BMODE == B_CONST) boris_rotate<COORD>(s.tx, s.ty, s.tz, s.sx, s.sy, s.sz, vx, vy, vz);
    if (BMODE == B_TABLE) boris_rotate_table<COORD>(g, s, x, z, vx, vy, vz);
    vx += Ex * s.hq;
    vz += Ez * s.hq;
}

// round-to-nearest-even of w * 2^32 for 0 <= w <= 1 (the Q32 rule of the fixed-point deposit) without the
// slow F2I.S64 path: adding 1.5 * 2^52 leaves the integer in the low mantissa bits; the scaling by 2^32 is
// exact, so this equals llrint(w * 4294967296.0) of the CPU restatement bit for bit
__device__ __forceinline__ unsigned long long q32_rn(double w)
{
    const double magic = 6755399441055744.0;   // 1.5 * 2^52
    const double t = __dadd_rn(__dmul_rn(w, 4294967296.0), magic);
    return (unsigned long long)(__double_as_longlong(t) - __double_as_longlong(magic));
}

// ---- boundary + electrode absorption + fixed-point CIC weights --------------------------------------
// Returns false when the particle is removed.  Positions may be wrapped (PERIODIC).  When DEPOSIT, node
// receives the index of the cell's lower-left node and w the four Q32 weights (Field2D.hpp:57-60 order:
// [i][j], [i+1][j], [i][j+1], [i+1][j+1]).
template <bool DEPOSIT>
__device__ __forceinline__ bool boundary_weights(const GridDev& g, double& x, double& z, unsigned& node, unsigned long long (&w)[4], unsigned* row = nullptr)
{
    node = 0;
    if (!(x >= 0.0 && x <= g.x_max && z >= 0.0 && z <= g.z_max))
    {
        if (g.boundary == MAG2D_BOUNDARY_FREE || !(x == x && z == z)) return false;
        x = fmod(x, g.x_max);
        if (x < 0) x += g.x_max;
        z = fmod(z, g.z_max);
        if (z < 0) z += g.z_max;
    }
    // products rounded separately (no FMA): the CPU restatement must reproduce the weights bit for bit
    const double X = __dmul_rn(x, g.idx), Y = __dmul_rn(z, g.idz);
    int i = (int)X, j = (int)Y;
    // the reference indexes one node past the grid for x == x_max exactly (fields.hpp:96-100); clamp
    i = max(min(i, g.M - 2), 0);
    j = max(min(j, g.N - 2), 0);
    const unsigned k = (unsigned)i * (unsigned)g.N + (unsigned)j;      // grids stay far below 2^32 nodes
    node = k;
    if (row) *row = (unsigned)i;
    if (g.check_mask && !g.cfree[k]) return false;
    if (DEPOSIT)
    {
        const double fu = __dsub_rn(X, (double)i), fv = __dsub_rn(Y, (double)j);
        const double cu = __dsub_rn(1.0, fu), cv = __dsub_rn(1.0, fv);
        w[0] = q32_rn(__dmul_rn(cu, cv));
        w[1] = q32_rn(__dmul_rn(fu, cv));
        w[2] = q32_rn(__dmul_rn(cu, fv));
        w[3] = q32_rn(__dmul_rn(fu, fv));
    }
    return true;
}

// ---- warp-aggregated scatter of the fixed-point weights ----------------------------------------------
// Must be called by all 32 lanes.  Particles are cell-sorted, so the lanes of a warp share a handful of
// cells: for each distinct cell the four weights are summed across its lanes with REDUX (two 32-bit
// pieces per weight: hi = w >> 16 <= 2^16, lo < 2^16, so 32 lanes cannot overflow) and lanes 0..3 issue one
// RED.ADD.64 each.  Integer sums are associative: the grid is bit-identical to the unaggregated scatter.
// After MAX_RUNS distinct cells the remaining lanes fall back to their own four REDs (unsorted input).
template <int MAX_RUNS>
__device__ __forceinline__ void warp_deposit(unsigned long long* __restrict__ rho, int N, bool valid, unsigned node,
                                             const unsigned long long (&w)[4])
{
    const unsigned lane = lane_id();
#if MAG2D_DEPOSIT_SMALL > 0
    if (!__any_sync(MAG2D_FULL_MASK, valid)) return;
    // MATCH.ANY groups the lanes by cell in one instruction: cells that only one or two lanes sit in (particles that
    // drifted out of the warp's home cell) are scattered directly; the REDUX merge is kept for the crowded cells
    const unsigned group = __match_any_sync(MAG2D_FULL_MASK, valid ? node : (0xFFFFFFE0u | lane));
    const bool small = valid && __popc(group) <= MAG2D_DEPOSIT_SMALL;
    if (small)
    {
        unsigned long long* r = rho + node;
        atomicAdd(r, w[0]);
        atomicAdd(r + N, w[1]);
        atomicAdd(r + 1, w[2]);
        atomicAdd(r + N + 1, w[3]);
    }
    valid = valid && !small;
#endif
    unsigned remaining = __ballot_sync(MAG2D_FULL_MASK, valid);
    if (remaining == 0) return;
    // 32-bit pieces once per entry: hi = w >> 16 (<= 2^16), lo = w & 0xffff
    unsigned piece[8];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
        piece[2 * q] = (unsigned)(w[q] & 0xFFFFu);
        piece[2 * q + 1] = (unsigned)(w[q] >> 16);
    }
#pragma unroll 1
    for (int it = 0; remaining && it < MAX_RUNS; it++)
    {
        const int src = __ffs(remaining) - 1;
        const unsigned k0 = __shfl_sync(MAG2D_FULL_MASK, node, src);
        const bool mine = valid && node == k0;
        const unsigned m = __ballot_sync(MAG2D_FULL_MASK, mine);
        unsigned sum[8];
#pragma unroll
        for (int q = 0; q < 8; q++) sum[q] = __reduce_add_sync(MAG2D_FULL_MASK, mine ? piece[q] : 0u);
        if (lane < 4)
        {
            const unsigned lo = lane == 0 ? sum[0] : lane == 1 ? sum[2] : lane == 2 ? sum[4] : sum[6];
            const unsigned hi = lane == 0 ? sum[1] : lane == 1 ? sum[3] : lane == 2 ? sum[5] : sum[7];
            const unsigned off = k0 + (lane & 1 ? (unsigned)N : 0u) + (lane >> 1);
            atomicAdd(rho + off, ((unsigned long long)hi << 16) + lo);
        }
        remaining &= ~m;
    }
    if (valid && ((remaining >> lane) & 1u))
    {
        // left-overs (particles that drifted away from the cells the warp mostly sits in): own REDs
        unsigned long long* r = rho + node;
        atomicAdd(r, w[0]);
        atomicAdd(r + N, w[1]);
        atomicAdd(r + 1, w[2]);
        atomicAdd(r + N + 1, w[3]);
    }
}

__device__ __forceinline__ void count_removed(unsigned long long* counter, bool removed_now)
{
    const unsigned m = __ballot_sync(MAG2D_FULL_MASK, removed_now);
    if (m && lane_id() == 0) atomicAdd(counter, (unsigned long long)__popc(m));
}

// ---- the fused Boris step --------------------------------------------------------------------------
// A warp owns a tile of 128 consecutive slots; lane l handles the pairs (2l, 2l+1) and (64+2l, 64+2l+1)
// of the tile, so that every array is moved with 128-bit loads/stores that a warp issues fully coalesced
// (512 B per instruction).  Four particles per thread amortise the Philox block (one word per particle),
// the address arithmetic and — because neighbouring slots sit in the same cell after the sort — the
// charge scatter: a thread first merges the weights of its own particles that share a cell, then the
// warp merges equal cells across lanes (warp_deposit).
//
// Collisions: only the Bernoulli test of the null-collision method (uni() < 1-exp(-dt/lifetime),
// particles.cpp:990) runs here; the slots that fire are appended to a list and scattered by k_mcc_collide
// afterwards.  That is legal because scatter() only changes the velocity, which neither the boundary test
// nor the deposit reads, and it keeps the rarely-taken, register-hungry collision kinematics out of this
// kernel.
#ifndef MAG2D_PPT
#define MAG2D_PPT 2   // measured on C4: 1.88 ms (2 per thread, no spills) vs 1.91 ms (4 per thread, 32 B of spills)
#endif
constexpr int PPT = MAG2D_PPT;               // particles per thread
constexpr int TILE = 32 * PPT;               // slots per warp
#ifndef MAG2D_DEPOSIT_RUNS
#define MAG2D_DEPOSIT_RUNS 6   // measured on C4: 1.872 ms (6) vs 1.891 ms (4)
#endif
#ifndef MAG2D_PUSH_MIN_BLOCKS
#define MAG2D_PUSH_MIN_BLOCKS 4   // 64 registers: measured 2.00 ms vs 2.16 ms (3 blocks, 80 regs) on C4
#endif
constexpr int DEPOSIT_RUNS = MAG2D_DEPOSIT_RUNS;   // cells per warp call that get the REDUX treatment

#ifndef MAG2D_SORT_MIN_BLOCKS
#define MAG2D_SORT_MIN_BLOCKS MAG2D_PUSH_MIN_BLOCKS
#endif
template <int COORD, bool GATHER, int BMODE, bool MCC, bool DEPOSIT, bool SORTING, typename T>
__global__ void __launch_bounds__(PUSH_THREADS, SORTING ? MAG2D_SORT_MIN_BLOCKS : MAG2D_PUSH_MIN_BLOCKS) k_push_boris(const __grid_constant__ PushArgs A)
{
    const unsigned lane = lane_id();
    const long long warp_id = ((long long)blockIdx.x * PUSH_THREADS + threadIdx.x) >> 5;
    const long long tile0 = warp_id * TILE;
    const long long n = A.p.n;
    if (tile0 >= n) return;                   // warp-uniform; arrays are allocated in multiples of 256 slots
    const long long base = tile0 + 2 * lane;  // slot of this thread's first pair
    constexpr bool need_vy = BMODE != B_NONE || COORD == MAG2D_CYLINDRICAL;
    const bool permute = SORTING && A.permute, count = SORTING && A.count;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (MCC)
    {
        Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)base);
        rnd = rng.block();
    }
    const double dt = A.s.dt;
    unsigned hit_mask = 0, removed = 0;
    long long dest[PPT];                      // slot this step's output of particle q lives in (-1: none)
#pragma unroll
    for (int p = 0; p < PPT / 2; p++)
    {
        const long long k = base + 64 * p;
        double x[2], z[2], vx[2], vz[2], vy[2];
        {
            const double2 a = pld2<T>(A.p.x, k);
            const double2 b = pld2<T>(A.p.z, k);
            const double2 c = pld2<T>(A.p.vx, k);
            const double2 d = pld2<T>(A.p.vz, k);
            x[0] = a.x; x[1] = a.y; z[0] = b.x; z[1] = b.y;
            vx[0] = c.x; vx[1] = c.y; vz[0] = d.x; vz[1] = d.y;
            vy[0] = vy[1] = 0.0;
            if (need_vy || permute)
            {
                const double2 e = pld2<T>(A.p.vy, k);
                vy[0] = e.x; vy[1] = e.y;
            }
        }
        dest[2 * p] = k;
        dest[2 * p + 1] = k + 1;
        if (permute)
        {
            // dead slots and the slots past n have no destination: the permutation compacts
            // the particle sits exactly where the COUNT push left it: recompute that cell (same operations as
            // boundary_weights) and draw the next slot of the cell from the cursor array (the scanned counts)
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const double px = x[e], pz = z[e];
                const bool alive = (k + e < n) && particle_alive(px);
                const int ci = max(min((int)__dmul_rn(px, A.g.idx), A.g.M - 2), 0), cj = max(min((int)__dmul_rn(pz, A.g.idz), A.g.N - 2), 0);
                const unsigned key = (unsigned)ci * (unsigned)(A.g.N - 1) + (unsigned)cj;
                const unsigned slot = warp_ticket(A.cursor, alive, alive ? key : 0u);
                dest[2 * p + e] = alive ? (long long)slot : -1;
            }
        }
        bool keep[2];
        unsigned node[2], row[2] = {0, 0};
        unsigned long long w[2][4];
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            // removed slots (x = NaN) and the slots past n are pushed like everybody else — NaNs and garbage
            // only ever index clamped cells — and masked out at the end: no divergent branch in the hot path
            const bool live = (k + e < n) && particle_alive(x[e]);
            double Ex = 0.0, Ez = A.g.extern_field;
            if (GATHER) gather_E(A.g, x[e], z[e], Ex, Ez);
            boris_velocity<COORD, BMODE>(A.g, A.s, x[e], z[e], Ex, Ez, vx[e], vy[e], vz[e]);
            if (COORD == MAG2D_CYLINDRICAL)
            {
                // Birdsall & Langdon p.338: drift in the local Cartesian frame, rotate back (particles.cpp:599-614)
                const double x2 = x[e] + vx[e] * dt;
                const double y2 = vy[e] * dt;
                x[e] = sqrt(x2 * x2 + y2 * y2);
                z[e] += vz[e] * dt;
                double sa = y2 / x[e], ca = x2 / x[e];
                if (x[e] == 0) { sa = 0; ca = 1; }
                const double t = vx[e];
                vx[e] = ca * vx[e] + sa * vy[e];
                vy[e] = -sa * t + ca * vy[e];
            }
            else
            {
                x[e] += vx[e] * dt;
                z[e] += vz[e] * dt;
            }
            if (sizeof(T) == 4)
            {
                // fp32 storage: the boundary test, the cell and the deposit see the position as it will be stored
                x[e] = stored<T>(x[e]);
                z[e] = stored<T>(z[e]);
            }
            const bool inside = boundary_weights<DEPOSIT>(A.g, x[e], z[e], node[e], w[e], SORTING ? &row[e] : nullptr);
            keep[e] = live && inside;
            removed += (live && !inside) ? 1u : 0u;
            if (!keep[e]) x[e] = dead_marker();
            if (MCC)
            {
                const int q = 2 * p + e;
                const unsigned word = q == 0 ? rnd.x : q == 1 ? rnd.y : q == 2 ? rnd.z : rnd.w;
                if (keep[e] && (unsigned long long)word < A.s.prob_u32) hit_mask |= 1u << q;
            }
        }
        if (permute)
        {
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const long long d = dest[2 * p + e];
                if (d < 0) continue;
                pst<T>(A.dst.x, d, x[e]);
                pst<T>(A.dst.z, d, z[e]);
                pst<T>(A.dst.vx, d, vx[e]);
                pst<T>(A.dst.vz, d, vz[e]);
                pst<T>(A.dst.vy, d, vy[e]);
            }
        }
        else
        {
            pst2<T>(A.p.x, k, x[0], x[1]);
            pst2<T>(A.p.z, k, z[0], z[1]);
            pst2<T>(A.p.vx, k, vx[0], vx[1]);
            pst2<T>(A.p.vz, k, vz[0], vz[1]);
            if (need_vy) pst2<T>(A.p.vy, k, vy[0], vy[1]);
        }
        if (count)
        {
            // cell counts for the next permuting step, keyed by the cell the particle sits in now (node = i*N + j)
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const unsigned key = keep[e] ? node[e] - row[e] : SORT_INVALID_KEY;      // i*N + j - i = i*(N-1) + j
                warp_count(A.count_out, keep[e], key);
            }
        }
        if (DEPOSIT)
        {
            // neighbouring slots share a cell after the sort: merge the pair, then merge across the warp
            if (keep[0] && keep[1] && node[0] == node[1])
            {
#pragma unroll
                for (int c = 0; c < 4; c++) w[0][c] += w[1][c];
                keep[1] = false;
            }
            warp_deposit<DEPOSIT_RUNS>(A.g.rho, A.g.N, keep[0], node[0], w[0]);
            warp_deposit<DEPOSIT_RUNS>(A.g.rho, A.g.N, keep[1], node[1], w[1]);
        }
    }
    if (MCC)
    {
        // warp-aggregated append of the firing slots to the collision list
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
#pragma unroll
            for (int q = 0; q < PPT; q++)
                if (hit_mask & (1u << q)) A.coll_list[start++] = (unsigned)dest[q];
        }
    }
    // removal counter: one atomic per warp, only when something was removed
    if (__any_sync(MAG2D_FULL_MASK, removed != 0))
    {
        unsigned rsum = removed;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(MAG2D_FULL_MASK, rsum, o);
        if (lane == 0) atomicAdd(A.removed, (unsigned long long)rsum);
    }
}

#ifdef MAG2D_WITH_TMA_PUSH
// ---- the same step with TMA-staged particle tiles ----------------------------------------------------
// Persistent CTAs (MAG2D_TMA_CTAS_PER_SM per SM) walk the slot range in tiles of 256 * PPT slots.  One elected
// thread moves every phase-space array of a tile HBM -> shared memory with 1-D bulk copies (cp.async.bulk, the TMA
// engine) that complete on an mbarrier, MAG2D_TMA_STAGES tiles deep, and moves the updated tile shared -> HBM with a
// bulk store; the compute threads only touch shared memory, so the HBM latency of the particle stream is off their
// scoreboards and no registers are spent on staging.  The arithmetic is the code of k_push_boris, shared verbatim.
#ifndef MAG2D_TMA_STAGES
#define MAG2D_TMA_STAGES 3
#endif
#ifndef MAG2D_TMA_CTAS_PER_SM
#define MAG2D_TMA_CTAS_PER_SM 2
#endif
constexpr int TMA_STAGES = MAG2D_TMA_STAGES;
constexpr int TMA_TILE = PUSH_THREADS * PPT;

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
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COORD, bool HASB>
constexpr int tma_narr() { return (HASB || COORD == MAG2D_CYLINDRICAL) ? 5 : 4; }
template <int COORD, bool HASB>
constexpr int tma_smem_bytes() { return TMA_STAGES * tma_narr<COORD, HASB>() * TMA_TILE * (int)sizeof(double) + 64; }

template <int COORD, bool GATHER, bool HASB, bool MCC, bool DEPOSIT>
__global__ void __launch_bounds__(PUSH_THREADS, MAG2D_TMA_CTAS_PER_SM) k_push_boris_tma(const __grid_constant__ PushArgs A)
{
    constexpr bool need_vy = HASB || COORD == MAG2D_CYLINDRICAL;
    constexpr int NARR = need_vy ? 5 : 4;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    double* const buf = reinterpret_cast<double*>(tma_smem);                       // [stage][array][slot]
    unsigned long long* const full = reinterpret_cast<unsigned long long*>(tma_smem + (size_t)TMA_STAGES * NARR * TMA_TILE * sizeof(double));
    const unsigned t = threadIdx.x, lane = t & 31;
    const long long n = A.p.n;
    const long long n_round = (n + 255) / 256 * 256;        // the slabs are allocated in multiples of 256 slots
    const long long ntiles = (n_round + TMA_TILE - 1) / TMA_TILE;
    double* const arr[5] = {A.p.x, A.p.z, A.p.vx, A.p.vz, A.p.vy};
    if (t == 0)
    {
        for (int s = 0; s < TMA_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto tile_bytes = [&](long long tile) { return (unsigned)(min((long long)TMA_TILE, n_round - tile * TMA_TILE) * (long long)sizeof(double)); };
    auto issue_load = [&](long long tile, int stage) {
        const unsigned bytes = tile_bytes(tile);
        mbar_expect_tx(&full[stage], NARR * bytes);
#pragma unroll
        for (int a = 0; a < NARR; a++) bulk_load(buf + ((size_t)stage * NARR + a) * TMA_TILE, arr[a] + tile * TMA_TILE, bytes, &full[stage]);
    };
    if (t == 0)
        for (int s = 0; s < TMA_STAGES - 1; s++)
        {
            const long long tile = blockIdx.x + (long long)s * gridDim.x;
            if (tile < ntiles) issue_load(tile, s);
        }
    const double dt = A.s.dt;
    unsigned removed = 0;
    for (int it = 0;; it++)
    {
        const long long tile = blockIdx.x + (long long)it * gridDim.x;
        if (tile >= ntiles) break;
        const int stage = it % TMA_STAGES;
        const unsigned parity = (unsigned)(it / TMA_STAGES) & 1u;
        while (!mbar_try_wait(&full[stage], parity)) {}
        double* const sx = buf + (size_t)stage * NARR * TMA_TILE;
        double* const sz = sx + TMA_TILE;
        double* const svx = sz + TMA_TILE;
        double* const svz = svx + TMA_TILE;
        double* const svy = svz + TMA_TILE;       // only dereferenced when need_vy
        const long long tile0 = tile * TMA_TILE;
        const long long base = tile0 + 2 * t;     // first slot of this thread: pairs (2t, 2t+1) + 512 p
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (MCC)
        {
            Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)base);
            rnd = rng.block();
        }
        unsigned hit_mask = 0;
#pragma unroll
        for (int p = 0; p < PPT / 2; p++)
        {
            const int slot = 2 * (int)t + 2 * PUSH_THREADS * p;
            const long long k = tile0 + slot;
            double x[2], z[2], vx[2], vz[2], vy[2];
            {
                const double2 a = *reinterpret_cast<const double2*>(sx + slot);
                const double2 b = *reinterpret_cast<const double2*>(sz + slot);
                const double2 c = *reinterpret_cast<const double2*>(svx + slot);
                const double2 d = *reinterpret_cast<const double2*>(svz + slot);
                x[0] = a.x; x[1] = a.y; z[0] = b.x; z[1] = b.y;
                vx[0] = c.x; vx[1] = c.y; vz[0] = d.x; vz[1] = d.y;
                vy[0] = vy[1] = 0.0;
                if (need_vy)
                {
                    const double2 e = *reinterpret_cast<const double2*>(svy + slot);
                    vy[0] = e.x; vy[1] = e.y;
                }
            }
            bool keep[2];
            unsigned node[2];
            unsigned long long w[2][4];
#pragma unroll
            for (int e = 0; e < 2; e++)
            {
                const bool live = (k + e < n) && particle_alive(x[e]);
                double Ex = 0.0, Ez = A.g.extern_field;
                if (GATHER) gather_E(A.g, x[e], z[e], Ex, Ez);
                boris_velocity<COORD, HASB ? B_CONST : B_NONE>(A.g, A.s, x[e], z[e], Ex, Ez, vx[e], vy[e], vz[e]);
                if (COORD == MAG2D_CYLINDRICAL)
                {
                    const double x2 = x[e] + vx[e] * dt;
                    const double y2 = vy[e] * dt;
                    x[e] = sqrt(x2 * x2 + y2 * y2);
                    z[e] += vz[e] * dt;
                    double sa = y2 / x[e], ca = x2 / x[e];
                    if (x[e] == 0) { sa = 0; ca = 1; }
                    const double tv = vx[e];
                    vx[e] = ca * vx[e] + sa * vy[e];
                    vy[e] = -sa * tv + ca * vy[e];
                }
                else
                {
                    x[e] += vx[e] * dt;
                    z[e] += vz[e] * dt;
                }
                const bool inside = boundary_weights<DEPOSIT>(A.g, x[e], z[e], node[e], w[e]);
                keep[e] = live && inside;
                removed += (live && !inside) ? 1u : 0u;
                if (!keep[e]) x[e] = dead_marker();
                if (MCC)
                {
                    const int q = 2 * p + e;
                    const unsigned word = q == 0 ? rnd.x : q == 1 ? rnd.y : q == 2 ? rnd.z : rnd.w;
                    if (keep[e] && (unsigned long long)word < A.s.prob_u32) hit_mask |= 1u << q;
                }
            }
            *reinterpret_cast<double2*>(sx + slot) = make_double2(x[0], x[1]);
            *reinterpret_cast<double2*>(sz + slot) = make_double2(z[0], z[1]);
            *reinterpret_cast<double2*>(svx + slot) = make_double2(vx[0], vx[1]);
            *reinterpret_cast<double2*>(svz + slot) = make_double2(vz[0], vz[1]);
            if (need_vy) *reinterpret_cast<double2*>(svy + slot) = make_double2(vy[0], vy[1]);
            if (DEPOSIT)
            {
                if (keep[0] && keep[1] && node[0] == node[1])
                {
#pragma unroll
                    for (int c = 0; c < 4; c++) w[0][c] += w[1][c];
                    keep[1] = false;
                }
                warp_deposit<DEPOSIT_RUNS>(A.g.rho, A.g.N, keep[0], node[0], w[0]);
                warp_deposit<DEPOSIT_RUNS>(A.g.rho, A.g.N, keep[1], node[1], w[1]);
            }
        }
        // the updated tile goes back with one bulk store per array; make the generic-proxy writes visible to it
        fence_proxy_async();
        __syncthreads();
        if (t == 0)
        {
            const unsigned bytes = tile_bytes(tile);
#pragma unroll
            for (int a = 0; a < NARR; a++) bulk_store(arr[a] + tile0, buf + ((size_t)stage * NARR + a) * TMA_TILE, bytes);
            bulk_commit();
            // refill the stage that was stored one iteration ago: its store must have finished reading shared memory
            const long long next = tile + (long long)(TMA_STAGES - 1) * gridDim.x;
            if (next < ntiles)
            {
                bulk_wait_read<1>();
                issue_load(next, (it + TMA_STAGES - 1) % TMA_STAGES);
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
                    const unsigned tt = __shfl_up_sync(MAG2D_FULL_MASK, incl, o);
                    if (lane >= (unsigned)o) incl += tt;
                }
                const unsigned total = __shfl_sync(MAG2D_FULL_MASK, incl, 31);
                unsigned start = 0;
                if (lane == 31) start = atomicAdd(A.coll_count, total);
                start = __shfl_sync(MAG2D_FULL_MASK, start, 31) + incl - cnt;
#pragma unroll
                for (int q = 0; q < PPT; q++)
                    if (hit_mask & (1u << q)) A.coll_list[start++] = (unsigned)(base + 2 * PUSH_THREADS * (q >> 1) + (q & 1));
            }
        }
    }
    if (t == 0) bulk_wait_read<0>();
    if (__any_sync(MAG2D_FULL_MASK, removed != 0))
    {
        unsigned rsum = removed;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(MAG2D_FULL_MASK, rsum, o);
        if (lane == 0) atomicAdd(A.removed, (unsigned long long)rsum);
    }
}

#endif  // MAG2D_WITH_TMA_PUSH

// second pass of the Boris movers: BaseSpecies::scatter for the slots whose Bernoulli test fired
template <typename T>
__global__ void __launch_bounds__(128) k_mcc_collide(const __grid_constant__ PushArgs A)
{
    const unsigned n = *A.coll_count;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x)
    {
        const long long k = A.coll_list[q];
        double vx = pld<T>(A.p.vx, k), vy = pld<T>(A.p.vy, k), vz = pld<T>(A.p.vz, k);
        Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)k);
        rng.draw = 1;     // block 0 was consumed by the Bernoulli test
        int target;
        const int proc = mcc_scatter(A.mcc, rng, vx, vy, vz, target);
        mcc_count(A.counts, A.mcc->n_targets, target, proc);
        if (proc >= 0)
        {
            pst<T>(A.p.vx, k, vx);
            pst<T>(A.p.vy, k, vy);
            pst<T>(A.p.vz, k, vz);
        }
    }
}

// ---- half step back: Species<D>::advance_boris_init ----------------------------------------------
template <int COORD, bool GATHER, typename T>
__global__ void __launch_bounds__(PUSH_THREADS) k_push_boris_init(const __grid_constant__ PushArgs A)
{
    const long long k = (long long)blockIdx.x * PUSH_THREADS + threadIdx.x;
    if (k >= A.p.n) return;
    const double x = pld<T>(A.p.x, k);
    if (!particle_alive(x)) return;
    const double z = pld<T>(A.p.z, k);
    double vx = pld<T>(A.p.vx, k), vy = pld<T>(A.p.vy, k), vz = pld<T>(A.p.vz, k);
    double Ex = 0.0, Ez = A.g.extern_field;
    if (GATHER) gather_E(A.g, x, z, Ex, Ez);
    // A.s holds the init constants: t = B*(-0.5*q*dt/(2m)), hq = (-q/m*dt)/2  (particles.cpp:1010,1028)
    const SpeciesDev& s = A.s;
    if (A.g.b_r) boris_rotate_table<COORD>(A.g, s, x, z, vx, vy, vz);
    else boris_rotate<COORD>(s.tx, s.ty, s.tz, s.sx, s.sy, s.sz, vx, vy, vz);
    vx += Ex * s.hq;
    vz += Ez * s.hq;
    pst<T>(A.p.vx, k, vx);
    pst<T>(A.p.vy, k, vy);
    pst<T>(A.p.vz, k, vz);
}

// ---- multi-collision free-flight mover: Species<CARTESIAN>::advance_multicoll ---------------------
// Constant field (0, extern_field); each particle flies to its next null-collision time, scatters,
// draws a new time_to_death = lifetime*rexp(), until the step is used up (~dt/lifetime events/step).
// The collision model (a few KB) is staged in shared memory: every event does 2-3 table look-ups.
__global__ void __launch_bounds__(PUSH_THREADS) k_push_multicoll(const __grid_constant__ PushArgs A, int blob_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_blob[];
    {
        const uint4* src = reinterpret_cast<const uint4*>(A.mcc);
        uint4* dst = reinterpret_cast<uint4*>(smem_blob);
        for (int q = threadIdx.x; q < blob_bytes / 16; q += PUSH_THREADS) dst[q] = src[q];
    }
    __syncthreads();
    const MccBlob* B = reinterpret_cast<const MccBlob*>(smem_blob);
    const long long k = (long long)blockIdx.x * PUSH_THREADS + threadIdx.x;
    const bool in_range = k < A.p.n;
    double x = in_range ? A.p.x[k] : dead_marker();
    bool removed_now = false;
    if (particle_alive(x))
    {
        double z = A.p.z[k], vx = A.p.vx[k], vy = A.p.vy[k], vz = A.p.vz[k], ttd = A.p.ttd[k];
        const double ax = 0.0 * A.s.qm, az = A.g.extern_field * A.s.qm;
        const double dt = A.s.dt;
        Rng rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)k);
        double local_time = 0.0;
        while (local_time + ttd < dt)
        {
            vx += ax * ttd;
            vz += az * ttd;
            // the reference advances the position with the already-updated velocity (particles.cpp:841-844)
            x += (vx + 0.5 * ax * ttd) * ttd;
            z += (vz + 0.5 * az * ttd) * ttd;
            local_time += ttd;
            int target;
            const int proc = mcc_scatter(B, rng, vx, vy, vz, target);
            mcc_count(A.counts, B->n_targets, target, proc);
            const uint4 r = rng.block();
            ttd = A.s.lifetime * rexp1(r.x);
        }
        const double rest = dt - local_time;
        vx += ax * rest;
        vz += az * rest;
        x += (vx + 0.5 * ax * rest) * rest;
        z += (vz + 0.5 * az * rest) * rest;
        ttd -= rest;
        unsigned node;
        unsigned long long w[4];
        const bool keep = boundary_weights<false>(A.g, x, z, node, w);
        if (keep)
        {
            A.p.x[k] = x;
            A.p.z[k] = z;
            A.p.vx[k] = vx;
            A.p.vy[k] = vy;
            A.p.vz[k] = vz;
            A.p.ttd[k] = ttd;
        }
        else
        {
            A.p.x[k] = dead_marker();
            removed_now = true;
        }
    }
    count_removed(A.removed, removed_now);
}

// ---- the same mover with persistent warps and lane refill ------------------------------------------------------------
// The number of events a particle goes through in a step is Poisson distributed (~dt / lifetime = 90 in C1), so in the kernel
// above a warp runs until its unluckiest lane is done (max of 32 Poisson(90) ~ 110 events: a fifth of the lane-iterations idle),
// at 16 resident warps per SM.  Here every warp owns a contiguous range of slots and walks it with lane refill: a lane whose
// particle has used up its step stores it and immediately takes the warp's next slot, so all 32 lanes stay inside the event
// loop until the range is exhausted.  Every particle still draws from its own Philox stream (keyed by slot and step), so the
// result does not depend on how lanes and particles are paired.  The one change to the stream is that the exponential variate of
// a new time_to_death takes the next unused word of a block instead of a fresh block each (one Philox block per four events
// saved), so the two kernels agree statistically, not bit for bit.  Not the default: see the launcher.
__global__ void __launch_bounds__(PUSH_THREADS, 2) k_push_multicoll_persistent(const __grid_constant__ PushArgs A, int blob_bytes, long long slots_per_warp)
{
    extern __shared__ __align__(16) unsigned char smem_blob[];
    {
        const uint4* src = reinterpret_cast<const uint4*>(A.mcc);
        uint4* dst = reinterpret_cast<uint4*>(smem_blob);
        for (int q = threadIdx.x; q < blob_bytes / 16; q += PUSH_THREADS) dst[q] = src[q];
    }
    __syncthreads();
    const MccBlob* B = reinterpret_cast<const MccBlob*>(smem_blob);
    const unsigned lane = lane_id();
    const long long warp_id = ((long long)blockIdx.x * PUSH_THREADS + threadIdx.x) >> 5;
    long long next = warp_id * slots_per_warp;
    const long long end = min(next + slots_per_warp, A.p.n);
    const double ax = 0.0 * A.s.qm, az = A.g.extern_field * A.s.qm;
    const double dt = A.s.dt;
    bool active = false;
    long long k = 0;
    double x = 0, z = 0, vx = 0, vy = 0, vz = 0, ttd = 0, local_time = 0;
    Rng rng = make_rng(A.seed, A.s.species, A.s.step, 0ULL);
    uint4 spare = make_uint4(0, 0, 0, 0);
    int n_spare = 0;
    unsigned removed = 0;
    while (true)
    {
        // ---- refill: idle lanes take the next slots of the warp's range (dead slots are skipped on the way)
        while (next < end)
        {
            const unsigned idle = __ballot_sync(MAG2D_FULL_MASK, !active);
            if (idle == 0) break;
            const long long mine = next + __popc(idle & ((1u << lane) - 1u));
            if (!active && mine < end)
            {
                const double px = A.p.x[mine];
                if (particle_alive(px))
                {
                    k = mine;
                    x = px;
                    z = A.p.z[k]; vx = A.p.vx[k]; vy = A.p.vy[k]; vz = A.p.vz[k]; ttd = A.p.ttd[k];
                    local_time = 0.0;
                    rng = make_rng(A.seed, A.s.species, A.s.step, (unsigned long long)k);
                    n_spare = 0;
                    active = true;
                }
            }
            next += __popc(idle);
        }
        if (!__any_sync(MAG2D_FULL_MASK, active)) break;
        if (active)
        {
            if (local_time + ttd < dt)
            {
                vx += ax * ttd;
                vz += az * ttd;
                // the reference advances the position with the already-updated velocity (particles.cpp:841-844)
                x += (vx + 0.5 * ax * ttd) * ttd;
                z += (vz + 0.5 * az * ttd) * ttd;
                local_time += ttd;
                int target;
                const int proc = mcc_scatter(B, rng, vx, vy, vz, target);
                mcc_count(A.counts, B->n_targets, target, proc);
                if (n_spare == 0)
                {
                    spare = rng.block();
                    n_spare = 4;
                }
                const unsigned word = n_spare == 4 ? spare.x : n_spare == 3 ? spare.y : n_spare == 2 ? spare.z : spare.w;
                n_spare--;
                ttd = A.s.lifetime * rexp1(word);
            }
            else
            {
                const double rest = dt - local_time;
                vx += ax * rest;
                vz += az * rest;
                x += (vx + 0.5 * ax * rest) * rest;
                z += (vz + 0.5 * az * rest) * rest;
                ttd -= rest;
                unsigned node;
                unsigned long long w[4];
                if (boundary_weights<false>(A.g, x, z, node, w))
                {
                    A.p.x[k] = x;
                    A.p.z[k] = z;
                    A.p.vx[k] = vx;
                    A.p.vy[k] = vy;
                    A.p.vz[k] = vz;
                    A.p.ttd[k] = ttd;
                }
                else
                {
                    A.p.x[k] = dead_marker();
                    removed++;
                }
                active = false;
            }
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

// ---- Species<D>::accumulate: deposit the current positions ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(PUSH_THREADS) k_accumulate(const __grid_constant__ PushArgs A)
{
    const long long k = (long long)blockIdx.x * PUSH_THREADS + threadIdx.x;
    double x = k < A.p.n ? pld<T>(A.p.x, k) : dead_marker();
    bool valid = particle_alive(x);
    unsigned node = 0;
    unsigned long long w[4] = {0, 0, 0, 0};
    if (valid)
    {
        double z = pld<T>(A.p.z, k);
        GridDev g = A.g;
        g.check_mask = 0;
        g.boundary = MAG2D_BOUNDARY_PERIODIC;   // never drop here: accumulate() deposits every live particle
        valid = boundary_weights<true>(g, x, z, node, w);
    }
    warp_deposit<8>(A.g.rho, A.g.N, valid, node, w);
}

// edge-centred differences of ue = u + phase*uRF: the g1..g4 terms of Field2D::grad (Field2D.hpp:97-100,
// 141-144), once per node and step instead of once per particle
__global__ void k_edge_fields(const double* __restrict__ u, const double* __restrict__ urf, double phase, int rf, int M, int N,
                              double idx, double idz, double* __restrict__ gx, double* __restrict__ gz)
{
    // [M+1][N+1] with ghosts: gx row 0 repeats row 1 and row M repeats row M-1 (the one-sided differences the
    // reference uses next to the walls); gz likewise along j; the ghost column of gx / ghost row of gz is zero
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > M || j > N) return;
    auto ue = [&](int a, int b) { const size_t q = (size_t)a * N + b; return rf ? u[q] + urf[q] * phase : u[q]; };
    const size_t k = (size_t)i * (N + 1) + j;
    const int ii = max(min(i, M - 1), 1), jj = max(min(j, N - 1), 1);
    gx[k] = j < N ? (ue(ii, j) - ue(ii - 1, j)) * idx : 0.0;
    gz[k] = i < M ? (ue(i, jj) - ue(i, jj - 1)) * idz : 0.0;
}

// Fields::E at arbitrary points (diagnostics, parity tests)
__global__ void k_field_E(const __grid_constant__ GridDev g, int n, const double* __restrict__ x, const double* __restrict__ z,
                          double* __restrict__ Ex, double* __restrict__ Ez)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double ex = 0.0, ez = g.extern_field;
    if (!g.const_E) gather_E(g, x[k], z[k], ex, ez);
    Ex[k] = ex;
    Ez[k] = ez;
}

// ---- particle source: Species<CARTESIAN>::source5_refresh / source (particles.cpp:1053-1080, 1158-1226) ------------
// A reservoir of density*V/factor particles lives in the periodic box [0, x_max/factor] x [0, z_max/factor] outside the
// simulated plasma.  Every step it is pushed with the external fields only; whatever leaves it is wrapped back and a copy
// enters the main box through the opposite edge, shifted by a random number of reservoir widths along the other axis —
// a thermal influx without simulating the surrounding plasma.  One thread per reservoir particle; copies take the next
// free tail slot of the species' store (atomic counter) and deposit their charge.
struct SourceArgs
{
    GridDev g;                     // check_mask = 0: source() tests the box only (particles.cpp:1178-1179)
    SpeciesDev s;
    double *x, *z, *vx, *vy, *vz, *ttd;   // reservoir
    long long n;
    ParticlesDev dst;              // the species' store
    long long dst_base, dst_cap;
    unsigned* inject_count;
    const MccBlob* mcc;            // nullptr: no collisions
    unsigned long long* counts;
    unsigned long long seed;
    unsigned factor;
    double src_x_max, src_z_max, v_scale;
};

__device__ __forceinline__ void source_inject(const SourceArgs& A, double x, double z, double vx, double vy, double vz, double ttd)
{
    if (!(z < A.g.z_max && z > 0 && x < A.g.x_max && x > 0)) return;        // the reference inserts and removes again
    const long long d = A.dst_base + atomicAdd(A.inject_count, 1u);
    if (d >= A.dst_cap) return;                                             // the host reads the count and reports it
    A.dst.x[d] = x;
    A.dst.z[d] = z;
    A.dst.vx[d] = vx;
    A.dst.vy[d] = vy;
    A.dst.vz[d] = vz;
    if (A.dst.ttd) A.dst.ttd[d] = ttd;
    if (A.g.deposit)
    {
        unsigned node;
        unsigned long long w[4];
        boundary_weights<true>(A.g, x, z, node, w);
        unsigned long long* r = A.g.rho + node;
        atomicAdd(r, w[0]);
        atomicAdd(r + A.g.N, w[1]);
        atomicAdd(r + 1, w[2]);
        atomicAdd(r + A.g.N + 1, w[3]);
    }
}

// INIT: the half step back of source5_refresh (advance_position_init(source2_particles, true), particles.cpp:1077)
template <bool INIT>
__global__ void __launch_bounds__(128) k_source(const __grid_constant__ SourceArgs A)
{
    const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
    if (k >= A.n) return;
    double x = A.x[k], z = A.z[k], vx = A.vx[k], vy = A.vy[k], vz = A.vz[k];
    // advance_boris(what, extern_fields = true): E = (0, extern_field), B = the constants, nothing is looked up
    const double Ex = 0.0, Ez = A.g.extern_field;
    if (INIT)
    {
        if (A.s.has_B) boris_rotate<MAG2D_CARTESIAN>(A.s.tx, A.s.ty, A.s.tz, A.s.sx, A.s.sy, A.s.sz, vx, vy, vz);
        vx += Ex * A.s.hq;
        vz += Ez * A.s.hq;
        A.vx[k] = vx;
        A.vy[k] = vy;
        A.vz[k] = vz;
        return;
    }
    if (A.s.has_B) boris_velocity<MAG2D_CARTESIAN, B_CONST>(A.g, A.s, x, z, Ex, Ez, vx, vy, vz);
    else boris_velocity<MAG2D_CARTESIAN, B_NONE>(A.g, A.s, x, z, Ex, Ez, vx, vy, vz);
    x += vx * A.s.dt;
    z += vz * A.s.dt;
    Rng rng = make_rng(A.seed ^ 0x9D2C5680A5A5F00DULL, A.s.species, A.s.step, (unsigned long long)k);
    const uint4 r0 = rng.block();
    if (A.mcc && (unsigned long long)r0.x < A.s.prob_u32)
    {
        int target;
        const int proc = mcc_scatter(A.mcc, rng, vx, vy, vz, target);
        mcc_count(A.counts, A.mcc->n_targets, target, proc);
    }
    const double ttd = A.ttd ? A.ttd[k] : 0.0;
    // lateral shifts: rand() % source5_factor of the reference; three words of the first block, then further blocks
    uint4 sh = r0;
    int used = 1;
    auto shift = [&]() -> double {
        if (used == 4)
        {
            sh = rng.block();
            used = 0;
        }
        const unsigned w = used == 0 ? sh.x : used == 1 ? sh.y : used == 2 ? sh.z : sh.w;
        used++;
        return (double)__umulhi(w, A.factor);
    };
    if (x > A.src_x_max)
        while (x > A.src_x_max)
        {
            x -= A.src_x_max;
            source_inject(A, x, z + shift() * A.src_z_max, vx, vy, vz, ttd);
        }
    else if (x < 0)
        while (x < 0)
        {
            source_inject(A, x + A.g.x_max, z + shift() * A.src_z_max, vx, vy, vz, ttd);
            x += A.src_x_max;
        }
    if (z > A.src_z_max)
        while (z > A.src_z_max)
        {
            z -= A.src_z_max;
            source_inject(A, x + shift() * A.src_x_max, z, vx, vy, vz, ttd);
        }
    else if (z < 0)
        while (z < 0)
        {
            source_inject(A, x + shift() * A.src_x_max, z + A.g.z_max, vx, vy, vz, ttd);
            z += A.src_z_max;
        }
    A.x[k] = x;
    A.z[k] = z;
    A.vx[k] = vx;
    A.vy[k] = vy;
    A.vz[k] = vz;
}

// the random part of source5_refresh (particles.cpp:1066-1075): uniform positions in the reservoir box, Maxwellian
// velocities rnor()*v_max/sqrt(2), time_to_death = rexp()*lifetime
__global__ void __launch_bounds__(128) k_source_generate(const __grid_constant__ SourceArgs A)
{
    const long long k = (long long)blockIdx.x * 128 + threadIdx.x;
    if (k >= A.n) return;
    Rng rng = make_rng(A.seed ^ 0x3C6EF372FE94F82BULL, A.s.species, A.s.step, (unsigned long long)k);
    const uint4 a = rng.block(), b = rng.block();
    A.x[k] = A.src_x_max * u01(a.x);
    A.z[k] = A.src_z_max * u01(a.y);
    float n0, n1, n2, n3;
    normal2(a.z, a.w, n0, n1);
    normal2(b.x, b.y, n2, n3);
    A.vx[k] = n0 * A.v_scale;
    A.vz[k] = n1 * A.v_scale;
    A.vy[k] = n2 * A.v_scale;
    A.ttd[k] = rexp1(b.z) * A.s.lifetime;
}

// Fields::B at n points (fields.hpp:152-177): the constants of the grid descriptor or the interpolated table
__global__ void k_field_B(const __grid_constant__ GridDev g, double Br0, double Bz0, double Bt0, int n, const double* __restrict__ x,
                          const double* __restrict__ z, double* __restrict__ Br, double* __restrict__ Bz, double* __restrict__ Bt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Br[k] = g.b_r ? table_interpolate(g, g.b_r, x[k], z[k]) : Br0;
    Bz[k] = g.b_r ? table_interpolate(g, g.b_z, x[k], z[k]) : Bz0;
    Bt[k] = g.b_r ? 0.0 : Bt0;
}

// ---- device-side loaders (Philox), src/particles.cpp:685-749, 485-510 -----------------------------
struct GenArgs
{
    ParticlesDev p;
    long long first, n;
    int kind;
    double a, b, c, d;
    double x_max, z_max, y_max;
    double vth;        // v_max / sqrt(2)
    double lifetime;
    double vmono;      // veV(energy)
    unsigned long long seed;
    int species;
    int has_ttd;
};

template <typename T>
__global__ void __launch_bounds__(PUSH_THREADS) k_generate(const __grid_constant__ GenArgs G)
{
    const long long q = (long long)blockIdx.x * PUSH_THREADS + threadIdx.x;
    if (q >= G.n) return;
    const long long k = G.first + q;
    Rng rng = make_rng(G.seed ^ 0xA5A5A5A55A5A5A5AULL, G.species, 0xFFFFFFFF00000000ULL + (unsigned long long)G.kind, (unsigned long long)k);
    double x, z, vx, vy, vz, y = 0.0;
    if (G.kind == 0)
    {
        const uint4 r = rng.block();
        x = G.x_max * u01(r.x);
        z = G.z_max * u01(r.z);
        y = G.y_max * u01(r.y);      // stored for CARTESIAN3D only
    }
    else
    {
        // rejection sampling of the unit disk, as add_particles_on_disk does
        double px, pz;
        do
        {
            const uint4 r = rng.block();
            px = u01(r.x) - 0.5;
            pz = u01(r.y) - 0.5;
        } while (px * px + pz * pz > 0.25);
        if (G.kind == 1)
        {
            x = px * 2 * G.c + G.a;
            z = pz * 2 * G.c + G.b;
        }
        else
        {
            const uint4 r = rng.block();
            x = sqrt(px * px + pz * pz) * G.c * 2;
            z = G.b + G.d * (u01(r.x) - 0.5);
        }
    }
    if (G.kind == 2)
    {
        const uint4 r = rng.block();
        rot_iso(G.vmono, r.x, r.y, vx, vz, vy);
    }
    else
    {
        const uint4 r = rng.block();
        float n0, n1, n2, n3;
        normal2(r.x, r.y, n0, n1);
        normal2(r.z, r.w, n2, n3);
        vx = (double)n0 * G.vth;
        vy = (double)n1 * G.vth;
        vz = (double)n2 * G.vth;
    }
    x = stored<T>(x);
    z = stored<T>(z);
    const bool inside = x >= 0.0 && x <= G.x_max && z >= 0.0 && z <= G.z_max;
    pst<T>(G.p.x, k, inside ? x : dead_marker());
    pst<T>(G.p.z, k, z);
    pst<T>(G.p.vx, k, vx);
    pst<T>(G.p.vy, k, vy);
    pst<T>(G.p.vz, k, vz);
    if (G.p.y) pst<T>(G.p.y, k, y);
    if (G.has_ttd)
    {
        const uint4 r = rng.block();
        pst<T>(G.p.ttd, k, G.kind == 0 ? G.lifetime * rexp1(r.x) : 0.0);   // on_disk leaves time_to_death = 0
    }
}

// ---- energy histogram: BaseSpecies::energy_dist_compute + Histogram::add (strict bounds) ----------
template <typename T>
__global__ void __launch_bounds__(PUSH_THREADS) k_energy_hist(ParticlesDev p, double half_m_over_qe, int nbins, double emax,
                                                               unsigned long long* __restrict__ hist, double* __restrict__ sums)
{
    extern __shared__ unsigned long long sh[];
    for (int q = threadIdx.x; q < nbins; q += PUSH_THREADS) sh[q] = 0;
    __syncthreads();
    double s_in = 0, s_tot = 0;
    unsigned long long n_in = 0, n_tot = 0;
    for (long long k = (long long)blockIdx.x * PUSH_THREADS + threadIdx.x; k < p.n; k += (long long)gridDim.x * PUSH_THREADS)
    {
        if (!particle_alive(pld<T>(p.x, k))) continue;
        const double vx = pld<T>(p.vx, k), vy = pld<T>(p.vy, k), vz = pld<T>(p.vz, k);
        const double f = (vx * vx + vz * vz + vy * vy) * half_m_over_qe;
        if (f < emax && f > 0.0)
        {
            int j = (int)(f * nbins / emax);
            j = min(j, nbins - 1);
            atomicAdd(&sh[j], 1ULL);
            n_in++;
            s_in += f;
        }
        n_tot++;
        s_tot += f;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < nbins; q += PUSH_THREADS)
        if (sh[q]) atomicAdd(&hist[q], sh[q]);
    // block reduction of the four scalars
    for (int o = 16; o > 0; o >>= 1)
    {
        s_in += __shfl_down_sync(MAG2D_FULL_MASK, s_in, o);
        s_tot += __shfl_down_sync(MAG2D_FULL_MASK, s_tot, o);
        n_in += __shfl_down_sync(MAG2D_FULL_MASK, n_in, o);
        n_tot += __shfl_down_sync(MAG2D_FULL_MASK, n_tot, o);
    }
    if (lane_id() == 0)
    {
        atomicAdd(&sums[0], (double)n_in);
        atomicAdd(&sums[1], s_in);
        atomicAdd(&sums[2], (double)n_tot);
        atomicAdd(&sums[3], s_tot);
    }
}

// ---- AoS (reference t_particle, 64 B) <-> SoA converters ------------------------------------------
template <typename T>
__global__ void k_aos_to_soa(const mag2d_particle* __restrict__ aos, long long n_in, ParticlesDev p, long long first,
                             int has_y, int has_ttd)
{
    // record q lands in slot first+q, so device order == input order; empty records become dead slots
    // (k_mark_dead) that the next sort compacts away
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_in || aos[q].empty) return;
    const long long dst = first + q;
    pst<T>(p.x, dst, aos[q].x);
    pst<T>(p.z, dst, aos[q].z);
    pst<T>(p.vx, dst, aos[q].vx);
    pst<T>(p.vy, dst, aos[q].vy);
    pst<T>(p.vz, dst, aos[q].vz);
    if (has_y) pst<T>(p.y, dst, aos[q].y);
    if (has_ttd) pst<T>(p.ttd, dst, aos[q].time_to_death);
}

template <typename T>
__global__ void k_mark_dead(ParticlesDev p, const mag2d_particle* __restrict__ aos, long long n_in, long long first)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_in && aos[q].empty) pst<T>(p.x, first + q, dead_marker());
}

template <typename T>
__global__ void k_soa_to_aos(ParticlesDev p, mag2d_particle* __restrict__ aos, int has_y, int has_ttd)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n) return;
    mag2d_particle o;
    memset(&o, 0, sizeof(o));
    const double x = pld<T>(p.x, q);
    o.empty = particle_alive(x) ? 0 : 1;
    o.x = x;
    o.z = pld<T>(p.z, q);
    o.vx = pld<T>(p.vx, q);
    o.vy = pld<T>(p.vy, q);
    o.vz = pld<T>(p.vz, q);
    o.y = has_y ? pld<T>(p.y, q) : 0.0;
    o.time_to_death = has_ttd ? pld<T>(p.ttd, q) : 0.0;
    aos[q] = o;
}

// ---- host helpers ---------------------------------------------------------------------------------
ParticlesDev particles_view(const SpeciesStore& S)
{
    ParticlesDev p;
    double* const* a = S.arr[S.cur];
    p.x = a[ARR_X];
    p.z = a[ARR_Z];
    p.vx = a[ARR_VX];
    p.vy = a[ARR_VY];
    p.vz = a[ARR_VZ];
    p.y = a[ARR_Y];
    p.ttd = a[ARR_TTD];
    p.n = S.n_slots;
    return p;
}

GridDev grid_view(const mag2d_ctx* c, int s)
{
    GridDev g;
    memset(&g, 0, sizeof(g));
    const mag2d_grid_desc& d = c->g;
    g.coord = d.coord;
    g.boundary = d.boundary;
    g.M = d.M;
    g.N = d.N;
    g.x_max = d.x_max;
    g.z_max = d.z_max;
    g.idx = d.idx;
    g.idz = d.idz;
    g.const_E = d.geometry_empty && !d.selfconsistent;
    g.check_mask = !d.electric_field_from_file && !c->all_cells_free;
    g.deposit = d.selfconsistent;
    g.extern_field = d.extern_field;
    g.dM1 = (double)(d.M - 1);
    g.dM2 = (double)(d.M - 2);
    g.dN1 = (double)(d.N - 1);
    g.dN2 = (double)(d.N - 2);
    g.gx = c->d_gx;
    g.gz = c->d_gz;
    g.cfree = c->d_cfree;
    g.rho = s >= 0 ? c->d_rho + (size_t)s * d.M * d.N : nullptr;
    // Fields::B reads the table only when magnetic_field_const = 0 (fields.hpp:154)
    g.b_r = c->g.magnetic_field_const ? nullptr : c->d_btab_r;
    g.b_z = c->g.magnetic_field_const ? nullptr : c->d_btab_z;
    g.bM = c->btab_M;
    g.bN = c->btab_N;
    g.bidx = c->btab_idx;
    g.bidz = c->btab_idz;
    g.bxmin = c->btab_xmin;
    g.bzmin = c->btab_zmin;
    return g;
}

// mymath.cpp:71-76
double mod_ref(double x, double y)
{
    if (x >= 0.0 && x <= y) return x;
    return x - y * (int)(x / y) + (x < 0 ? y : 0);
}

// SpeciesDev constants with the reference's expression order (particles.cpp:936-975 / 1010-1040)
SpeciesDev species_view(const mag2d_ctx* c, int s, bool init)
{
    const SpeciesStore& S = c->sp[s];
    const mag2d_grid_desc& d = c->g;
    SpeciesDev v;
    memset(&v, 0, sizeof(v));
    const double charge = S.desc.charge, mass = S.desc.mass, dt = S.desc.dt;
    v.dt = dt;
    const double qmdt = init ? -charge / mass * dt : charge / mass * dt;
    v.hq = qmdt / 2.0;
    double tmp = init ? -0.5 * charge * dt / (2.0 * mass) : charge * dt / (2.0 * mass);
    v.tb = tmp;
    const double Bx = d.Br, By = d.Bt, Bz = d.Bz;
    v.tx = Bx * tmp;
    v.ty = By * tmp;
    v.tz = Bz * tmp;
    tmp = 2.0 / (1 + v.tx * v.tx + v.ty * v.ty + v.tz * v.tz);
    v.sx = v.tx * tmp;
    v.sy = v.ty * tmp;
    v.sz = v.tz * tmp;
    v.has_B = (Bx != 0.0 || By != 0.0 || Bz != 0.0);
    v.species = s;
    v.prob = 1.0 - exp(-dt / S.lifetime);
    v.prob_u32 = bernoulli_threshold(v.prob);
    v.lifetime = S.lifetime;
    v.qm = charge / mass;
    v.step = S.niter;
    return v;
}

// storage type of a species' particle arrays: T = float in the fp32 storage mode (mag2d_set_storage), else double
#define DISPATCH_STORE(f32, ...)                 \
    do {                                         \
        if (f32) { using T = float; __VA_ARGS__; } \
        else { using T = double; __VA_ARGS__; }  \
    } while (0)

template <int COORD, bool SORTING, bool G, int B, typename T>
void launch_boris_gb(mag2d_ctx* c, const PushArgs& A, bool mcc, bool deposit, unsigned blocks)
{
#define LAUNCH(Mc, D) k_push_boris<COORD, G, B, Mc, D, SORTING, T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A)
    if (mcc) { if (deposit) LAUNCH(true, true); else LAUNCH(true, false); }
    else { if (deposit) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
}

template <int COORD, bool SORTING, typename T>
int launch_boris_variant_t(mag2d_ctx* c, const PushArgs& A, bool gather, int bmode, bool mcc, bool deposit, unsigned blocks)
{
    switch ((gather ? 3 : 0) + bmode)
    {
        case 0: launch_boris_gb<COORD, SORTING, false, B_NONE, T>(c, A, mcc, deposit, blocks); break;
        case 1: launch_boris_gb<COORD, SORTING, false, B_CONST, T>(c, A, mcc, deposit, blocks); break;
        case 2: launch_boris_gb<COORD, SORTING, false, B_TABLE, T>(c, A, mcc, deposit, blocks); break;
        case 3: launch_boris_gb<COORD, SORTING, true, B_NONE, T>(c, A, mcc, deposit, blocks); break;
        case 4: launch_boris_gb<COORD, SORTING, true, B_CONST, T>(c, A, mcc, deposit, blocks); break;
        default: launch_boris_gb<COORD, SORTING, true, B_TABLE, T>(c, A, mcc, deposit, blocks); break;
    }
    c->launches++;
    return 0;
}

template <int COORD, bool SORTING>
int launch_boris_variant(mag2d_ctx* c, const PushArgs& A, bool gather, int bmode, bool mcc, bool deposit, unsigned blocks)
{
    if (c->store_f32) return launch_boris_variant_t<COORD, SORTING, float>(c, A, gather, bmode, mcc, deposit, blocks);
    return launch_boris_variant_t<COORD, SORTING, double>(c, A, gather, bmode, mcc, deposit, blocks);
}

}  // namespace

// per-slot scratch shared by the sort (keys, ranks) and the push (collision list)
int ensure_particle_scratch(mag2d_ctx* c, long long capacity)
{
    if (!c->d_coll_count) CUDA_OK(cudaMalloc(&c->d_coll_count, sizeof(unsigned)));
    if (c->rank_capacity >= capacity) return 0;
    if (c->d_rank) cudaFree(c->d_rank);
    if (c->d_key) cudaFree(c->d_key);
    c->d_rank = c->d_key = nullptr;
    CUDA_OK(cudaMalloc(&c->d_rank, sizeof(unsigned) * (size_t)capacity));
    CUDA_OK(cudaMalloc(&c->d_key, sizeof(unsigned) * (size_t)capacity));
    c->rank_capacity = capacity;
    return 0;
}

int update_ueff(mag2d_ctx* c, double phase, bool rf)
{
    const int M = c->g.M, N = c->g.N;
    const dim3 block(32, 8), grid((N + 1 + 31) / 32, (M + 1 + 7) / 8);
    k_edge_fields<<<grid, block, 0, c->stream>>>(c->d_u, c->d_uRF, phase, rf ? 1 : 0, M, N, c->g.idx, c->g.idz, c->d_gx, c->d_gz);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static double rf_phase(const mag2d_ctx* c, const SpeciesStore& S)
{
    // Fields::E: phase = rf_amplitude*cos(mod(rf_omega*time, 1e7*pi)) + rf_U0, time = niter*dt (fields.hpp:140-142)
    const double time = S.niter * S.desc.dt;
    double phase = mod_ref(c->g.rf_omega * time, 10000000 * M_PI);
    return c->g.rf_amplitude * cos(phase) + c->g.rf_U0;
}

// sort_mode: bit 0 = permute (write into the other slab at the sorted slots drawn from the pending cell cursors), bit 1 = count
// (count the particles per cell for the next permuting step); Boris movers only
int launch_species_advance(mag2d_ctx* c, int s, int sort_mode)
{
    SpeciesStore& S = c->sp[s];
    const mag2d_grid_desc& d = c->g;
    // streamed step (abi.cu): the particle arrays are one chunk of a host-resident store staged in device buffers
    const bool chunked = c->chunk_view != nullptr;
    const long long n_active = chunked ? c->chunk_view->n : S.n_slots;
    // a push that does not consume the pending cell cursors moves (or removes) particles away from the cells they were
    // counted in: the cursors are stale from here on (a COUNT push re-validates them in sort_fused_end)
    if (!chunked && !(sort_mode & 1)) S.tickets_valid = false;
    if (n_active > 0)
    {
        if (!d.magnetic_field_const && !c->d_btab_r)
        {
            mag2d_set_error("magnetic_field_const = 0 but no table was loaded (mag2d_set_magnetic_field)");
            return 1;
        }
        PushArgs A;
        A.g = grid_view(c, s);
        A.s = species_view(c, s, false);
        A.p = chunked ? *c->chunk_view : particles_view(S);
        A.mcc = S.d_blob;
        A.counts = c->count_collisions ? S.d_counts : nullptr;
        A.removed = S.d_removed;
        A.seed = c->seed;
        if (chunked)
        {
            // every chunk of a streamed step draws from its own Philox key (splitmix64 of its first global slot);
            // folding the offset into the key costs the kernels nothing, an extra 64-bit add per thread cost 6 %
            unsigned long long zz = (unsigned long long)c->chunk_slot0 + 0x9E3779B97F4A7C15ULL;
            zz = (zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9ULL;
            zz = (zz ^ (zz >> 27)) * 0x94D049BB133111EBULL;
            A.seed ^= c->chunk_slot0 ? (zz ^ (zz >> 31)) : 0ULL;
        }
        A.coll_list = nullptr;
        A.coll_count = nullptr;
        A.permute = A.count = A.cell_cols = 0;
        A.cursor = A.count_out = nullptr;
        memset(&A.dst, 0, sizeof(A.dst));
        const unsigned blocks = (unsigned)((n_active + PUSH_THREADS - 1) / PUSH_THREADS);
        const unsigned tile_blocks = (unsigned)((n_active + PUSH_THREADS * PPT - 1) / (PUSH_THREADS * PPT));
        const bool mcc = S.h_blob && S.h_blob->has_collisions && std::isfinite(S.lifetime);
        if (chunked && d.mover == MAG2D_ADVANCE_MULTICOLL)
        {
            mag2d_set_error("mag2d_step_streamed: Boris movers only");
            return 1;
        }
        if (d.mover == MAG2D_ADVANCE_MULTICOLL)
        {
            if (d.coord != MAG2D_CARTESIAN)
            {
                mag2d_set_error("Species<CYLINDRICAL>: advance method not implemented\n");
                return 1;
            }
            if (d.Br != 0. || d.Bz != 0. || d.Bt != 0.)
            {
                mag2d_set_error("Species<CARTESIAN>::advance_multicoll() implemented for zero B only\n");
                return 1;
            }
            if (d.selfconsistent || !d.geometry_empty)
            {
                mag2d_set_error("Species<CARTESIAN>::advance_multicoll() implemented for const extern field only:\n\tset selfconsistent=0 and geometry=EMPTY\n");
                return 1;
            }
            const int blob_bytes = (int)((offsetof(MccBlob, tab) + sizeof(double) * 3 * (size_t)S.h_blob->n_tab + 15) / 16 * 16);
            static bool attr_set = false;
            if (!attr_set)
            {
                CUDA_OK(cudaFuncSetAttribute(k_push_multicoll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MccBlob)));
                attr_set = true;
            }
            // one thread per particle by default; MAG2D_MULTICOLL=persistent selects the lane-refill variant (measured on C1: warp
            // execution efficiency 23.3 -> 24.7 of 32 lanes, but 118 instead of 80 registers: 4.46 ms against 3.96 ms per 1e6 particles —
            // the idle lanes come from divergence INSIDE an event (process type, table searches), not from the Poisson trip counts)
            static const bool persistent = getenv("MAG2D_MULTICOLL") && !strcmp(getenv("MAG2D_MULTICOLL"), "persistent");
            if (!persistent) k_push_multicoll<<<blocks, PUSH_THREADS, blob_bytes, c->stream>>>(A, blob_bytes);
            else
            {
                // persistent warps: two CTAs per SM, every warp walks its own range of slots with lane refill
                static bool attr2 = false;
                if (!attr2)
                {
                    CUDA_OK(cudaFuncSetAttribute(k_push_multicoll_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MccBlob)));
                    attr2 = true;
                }
                const unsigned pblocks = (unsigned)std::min<long long>(blocks, 148 * 2);
                const long long warps = (long long)pblocks * (PUSH_THREADS / 32);
                const long long per_warp = (n_active + warps - 1) / warps;
                k_push_multicoll_persistent<<<pblocks, PUSH_THREADS, blob_bytes, c->stream>>>(A, blob_bytes, per_warp);
            }
            c->launches++;
        }
        else
        {
            const bool gather = !A.g.const_E;
            const int bmode = A.g.b_r ? B_TABLE : A.s.has_B ? B_CONST : B_NONE;
            if (gather && !(c->edge_fields_fresh && !d.rf))
            {
                if (update_ueff(c, d.rf ? rf_phase(c, S) : 0.0, d.rf != 0)) return 1;
                // mag2d_step arms this flag per iteration (it is false everywhere else): the next species of the iteration skips the pass
                if (c->edge_fields_fresh_armed) c->edge_fields_fresh = true;
            }
            if (mcc)
            {
                if (ensure_particle_scratch(c, chunked ? std::max(n_active, S.capacity) : S.capacity)) return 1;
                A.coll_list = c->d_key;          // the stand-alone sort's key buffer is idle during a push
                A.coll_count = c->d_coll_count;
                CUDA_OK(cudaMemsetAsync(c->d_coll_count, 0, sizeof(unsigned), c->stream));
            }
            const bool permute = (sort_mode & 1) != 0, count = (sort_mode & 2) != 0;
            if (permute || count)
            {
                if (sort_fused_begin(c, s, permute, count)) return 1;
                A.permute = permute;
                A.count = count;
                A.cell_cols = d.N - 1;
                A.cursor = S.d_cell_offset;
                A.count_out = S.d_cell_count;
                if (permute)
                {
                    double* const* o = S.arr[S.cur ^ 1];
                    A.dst.x = o[ARR_X];
                    A.dst.z = o[ARR_Z];
                    A.dst.vx = o[ARR_VX];
                    A.dst.vy = o[ARR_VY];
                    A.dst.vz = o[ARR_VZ];
                    A.dst.n = S.n_slots;
                }
                if (d.coord == MAG2D_CYLINDRICAL)
                    launch_boris_variant<MAG2D_CYLINDRICAL, true>(c, A, gather, bmode, mcc, d.selfconsistent != 0, tile_blocks);
                else
                    launch_boris_variant<MAG2D_CARTESIAN, true>(c, A, gather, bmode, mcc, d.selfconsistent != 0, tile_blocks);
                if (sort_fused_end(c, s, permute, count)) return 1;
                if (permute)
                {
                    A.p = particles_view(S);      // the collision pass works on the new slab
                    // ... and so must its partner pools when this species is its own (or another species') target
                    if (mcc && refresh_pools_all(c)) return 1;
                }
            }
            else if (d.coord == MAG2D_CYLINDRICAL)
                launch_boris_variant<MAG2D_CYLINDRICAL, false>(c, A, gather, bmode, mcc, d.selfconsistent != 0, tile_blocks);
            else
                launch_boris_variant<MAG2D_CARTESIAN, false>(c, A, gather, bmode, mcc, d.selfconsistent != 0, tile_blocks);
            if (mcc)
            {
                DISPATCH_STORE(c->store_f32, k_mcc_collide<T><<<148 * 8, 128, 0, c->stream>>>(A));
                c->launches++;
            }
        }
        CUDA_OK(cudaGetLastError());
    }
    if (chunked) return 0;        // mag2d_step_streamed advances the species clock once per step
    // Species<D>::advance: niter++, t += dt; advance_multicoll advances the clock a second time
    // (particles.cpp:857-858) — kept so that <name>.dat columns match the reference
    S.niter++;
    S.t += S.desc.dt;
    if (d.mover == MAG2D_ADVANCE_MULTICOLL)
    {
        S.niter++;
        S.t += S.desc.dt;
    }
    S.steps_since_sort++;
    if (S.pushes_since_permute < (1 << 20)) S.pushes_since_permute++;
    return 0;
}

int launch_species_advance_init(mag2d_ctx* c, int s)
{
    SpeciesStore& S = c->sp[s];
    const mag2d_grid_desc& d = c->g;
    if (S.n_slots == 0 || d.mover == MAG2D_ADVANCE_MULTICOLL) return 0;   // particles.cpp:805-806
    PushArgs A;
    A.g = grid_view(c, s);
    A.s = species_view(c, s, true);
    A.p = particles_view(S);
    A.mcc = nullptr;
    A.counts = nullptr;
    A.removed = S.d_removed;
    A.seed = c->seed;
    const unsigned blocks = (unsigned)((S.n_slots + PUSH_THREADS - 1) / PUSH_THREADS);
    const bool gather = !A.g.const_E;
    if (gather)
        if (update_ueff(c, d.rf ? rf_phase(c, S) : 0.0, d.rf != 0)) return 1;
    if (d.coord == MAG2D_CYLINDRICAL)
    {
        if (gather) DISPATCH_STORE(c->store_f32, k_push_boris_init<MAG2D_CYLINDRICAL, true, T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A));
        else DISPATCH_STORE(c->store_f32, k_push_boris_init<MAG2D_CYLINDRICAL, false, T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A));
    }
    else
    {
        if (gather) DISPATCH_STORE(c->store_f32, k_push_boris_init<MAG2D_CARTESIAN, true, T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A));
        else DISPATCH_STORE(c->store_f32, k_push_boris_init<MAG2D_CARTESIAN, false, T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A));
    }
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_species_accumulate(mag2d_ctx* c, int s)
{
    SpeciesStore& S = c->sp[s];
    if (S.n_slots == 0) return 0;
    PushArgs A;
    memset(&A, 0, sizeof(A));
    A.g = grid_view(c, s);
    A.p = particles_view(S);
    const unsigned blocks = (unsigned)((S.n_slots + PUSH_THREADS - 1) / PUSH_THREADS);
    DISPATCH_STORE(c->store_f32, k_accumulate<T><<<blocks, PUSH_THREADS, 0, c->stream>>>(A));
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_field_E(mag2d_ctx* c, int n, const double* x, const double* z, double time, double* Ex, double* Ez)
{
    double *dx, *dz, *dex, *dez;
    CUDA_OK(cudaMalloc(&dx, sizeof(double) * 4 * (size_t)n));
    dz = dx + n;
    dex = dz + n;
    dez = dex + n;
    CUDA_OK(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(dz, z, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    GridDev g = grid_view(c, -1);
    if (!g.const_E)
    {
        double phase = mod_ref(c->g.rf_omega * time, 10000000 * M_PI);
        phase = c->g.rf_amplitude * cos(phase) + c->g.rf_U0;
        if (update_ueff(c, phase, c->g.rf != 0)) return 1;
    }
    k_field_E<<<(n + 255) / 256, 256, 0, c->stream>>>(g, n, dx, dz, dex, dez);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(Ex, dex, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(Ez, dez, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(dx));
    return 0;
}

int source_alloc(mag2d_ctx* c, SpeciesStore& S, long long n)
{
    (void)c;
    if (!S.d_src_count) CUDA_OK(cudaMalloc(&S.d_src_count, sizeof(unsigned)));
    if (n > S.src_capacity)
    {
        for (int a = 0; a < 6; a++)
        {
            if (S.src[a]) cudaFree(S.src[a]);
            S.src[a] = nullptr;
            CUDA_OK(cudaMalloc(&S.src[a], sizeof(double) * (size_t)n));
        }
        S.src_capacity = n;
    }
    S.src_n = n;
    return 0;
}

void source_free(SpeciesStore& S)
{
    for (int a = 0; a < 6; a++)
    {
        if (S.src[a]) cudaFree(S.src[a]);
        S.src[a] = nullptr;
    }
    if (S.d_src_count) cudaFree(S.d_src_count);
    S.d_src_count = nullptr;
    S.src_n = S.src_capacity = 0;
}

static SourceArgs source_args(mag2d_ctx* c, int s, bool init)
{
    SpeciesStore& S = c->sp[s];
    SourceArgs A;
    memset(&A, 0, sizeof(A));
    A.g = grid_view(c, s);
    A.g.check_mask = 0;
    A.s = species_view(c, s, init);
    A.s.step = S.src_calls;
    A.x = S.src[0]; A.z = S.src[1]; A.vx = S.src[2]; A.vy = S.src[3]; A.vz = S.src[4]; A.ttd = S.src[5];
    A.n = S.src_n;
    A.seed = c->seed;
    A.factor = S.src_factor;
    const double K = 1.0 / S.src_factor;
    A.src_x_max = K * c->g.x_max;
    A.src_z_max = K * c->g.z_max;
    A.v_scale = S.v_max / M_SQRT2;
    return A;
}

int launch_source_generate(mag2d_ctx* c, int s, unsigned factor, long long n)
{
    SpeciesStore& S = c->sp[s];
    if (source_alloc(c, S, n)) return 1;
    S.src_factor = factor;
    if (n == 0) return 0;
    SourceArgs A = source_args(c, s, true);
    k_source_generate<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(A);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    S.src_calls++;
    return launch_source_init(c, s);
}

int launch_source_init(mag2d_ctx* c, int s)
{
    SpeciesStore& S = c->sp[s];
    if (S.src_n == 0) return 0;
    SourceArgs A = source_args(c, s, true);
    k_source<true><<<(unsigned)((S.src_n + 127) / 128), 128, 0, c->stream>>>(A);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// Species<CARTESIAN>::source (particles.cpp:1158-1226).  One host synchronisation: the new slot count comes back.
int launch_species_source(mag2d_ctx* c, int s, long long* injected)
{
    SpeciesStore& S = c->sp[s];
    if (injected) *injected = 0;
    if (S.src_n == 0) return 0;
    // the caller (abi.cu: species_source) has reserved room for two copies per reservoir particle — a corner crossing;
    // more would need v*dt beyond a reservoir width
    SourceArgs A = source_args(c, s, false);
    A.dst = particles_view(S);
    A.dst_base = S.n_slots;
    A.dst_cap = S.capacity;
    A.inject_count = S.d_src_count;
    const bool mcc = S.h_blob && S.h_blob->has_collisions && std::isfinite(S.lifetime);
    A.mcc = mcc ? S.d_blob : nullptr;
    A.counts = c->count_collisions ? S.d_counts : nullptr;
    CUDA_OK(cudaMemsetAsync(S.d_src_count, 0, sizeof(unsigned), c->stream));
    k_source<false><<<(unsigned)((S.src_n + 127) / 128), 128, 0, c->stream>>>(A);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    unsigned count = 0;
    CUDA_OK(cudaMemcpyAsync(&count, S.d_src_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    S.src_calls++;
    if ((long long)count > S.capacity - S.n_slots)
    {
        S.n_slots = S.capacity;
        mag2d_set_error("mag2d_species_source: more than two copies per reservoir particle in one step (dt too long for the reservoir)");
        return 1;
    }
    S.n_slots += count;
    if (injected) *injected = count;
    return 0;
}

int launch_field_B(mag2d_ctx* c, int n, const double* x, const double* z, double* Br, double* Bz, double* Bt)
{
    if (n <= 0) return 0;
    double* d;
    CUDA_OK(cudaMalloc(&d, sizeof(double) * 5 * (size_t)n));
    CUDA_OK(cudaMemcpyAsync(d, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(d + n, z, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    k_field_B<<<(n + 255) / 256, 256, 0, c->stream>>>(grid_view(c, -1), c->g.Br, c->g.Bz, c->g.Bt, n, d, d + n, d + 2 * (size_t)n,
                                                      d + 3 * (size_t)n, d + 4 * (size_t)n);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(Br, d + 2 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(Bz, d + 3 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(Bt, d + 4 * (size_t)n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d));
    return 0;
}

int launch_generate(mag2d_ctx* c, int s, int kind, long long n, double a, double b, double cc, double d)
{
    SpeciesStore& S = c->sp[s];
    if (n <= 0) return 0;
    GenArgs G;
    memset(&G, 0, sizeof(G));
    G.p = particles_view(S);
    G.first = S.n_slots;
    G.n = n;
    G.kind = kind;
    G.a = a;
    G.b = b;
    G.c = cc;
    G.d = d;
    G.x_max = c->g.x_max;
    G.z_max = c->g.z_max;
    G.y_max = c->g.y_max;
    G.vth = S.v_max / M_SQRT2;
    G.lifetime = std::isfinite(S.lifetime) ? S.lifetime : 0.0;
    G.vmono = sqrt(a * MAG2D_QE / S.desc.mass * 2.0);
    G.seed = c->seed;
    G.species = s;
    G.has_ttd = S.arr[S.cur][ARR_TTD] != nullptr;
    const unsigned blocks = (unsigned)((n + PUSH_THREADS - 1) / PUSH_THREADS);
    DISPATCH_STORE(c->store_f32, k_generate<T><<<blocks, PUSH_THREADS, 0, c->stream>>>(G));
    c->launches++;
    CUDA_OK(cudaGetLastError());
    S.n_slots += n;
    return 0;
}

int launch_energy_hist(mag2d_ctx* c, int s, int nbins, double emax, double* hist, double* stats)
{
    SpeciesStore& S = c->sp[s];
    unsigned long long* dh;
    double* ds;
    CUDA_OK(cudaMalloc(&dh, sizeof(unsigned long long) * nbins + sizeof(double) * 4));
    ds = reinterpret_cast<double*>(dh + nbins);
    CUDA_OK(cudaMemsetAsync(dh, 0, sizeof(unsigned long long) * nbins + sizeof(double) * 4, c->stream));
    if (S.n_slots > 0)
    {
        const unsigned blocks = (unsigned)std::min<long long>((S.n_slots + PUSH_THREADS - 1) / PUSH_THREADS, 148 * 8);
        DISPATCH_STORE(c->store_f32, k_energy_hist<T><<<blocks, PUSH_THREADS, sizeof(unsigned long long) * nbins, c->stream>>>(
            particles_view(S), S.desc.mass * 0.5 / MAG2D_QE, nbins, emax, dh, ds));
        c->launches++;
        CUDA_OK(cudaGetLastError());
    }
    std::vector<unsigned long long> hh(nbins);
    CUDA_OK(cudaMemcpyAsync(hh.data(), dh, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(stats, ds, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < nbins; q++) hist[q] = (double)hh[q];
    CUDA_OK(cudaFree(dh));
    return 0;
}

int launch_aos_to_soa(mag2d_ctx* c, int s, const mag2d_particle* d_aos, long long n_in, long long* n_added)
{
    SpeciesStore& S = c->sp[s];
    ParticlesDev p = particles_view(S);
    const unsigned blocks = (unsigned)((n_in + 255) / 256);
    const int has_y = S.arr[S.cur][ARR_Y] != nullptr, has_ttd = S.arr[S.cur][ARR_TTD] != nullptr;
    DISPATCH_STORE(c->store_f32, k_aos_to_soa<T><<<blocks, 256, 0, c->stream>>>(d_aos, n_in, p, S.n_slots, has_y, has_ttd));
    DISPATCH_STORE(c->store_f32, k_mark_dead<T><<<blocks, 256, 0, c->stream>>>(p, d_aos, n_in, S.n_slots));
    c->launches += 2;
    CUDA_OK(cudaGetLastError());
    *n_added = n_in;
    S.n_slots += n_in;
    return 0;
}

int launch_soa_to_aos(mag2d_ctx* c, int s, mag2d_particle* d_aos)
{
    SpeciesStore& S = c->sp[s];
    if (S.n_slots == 0) return 0;
    const unsigned blocks = (unsigned)((S.n_slots + 255) / 256);
    const int has_y = S.arr[S.cur][ARR_Y] != nullptr, has_ttd = S.arr[S.cur][ARR_TTD] != nullptr;
    DISPATCH_STORE(c->store_f32, k_soa_to_aos<T><<<blocks, 256, 0, c->stream>>>(particles_view(S), d_aos, has_y, has_ttd));
    c->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}
