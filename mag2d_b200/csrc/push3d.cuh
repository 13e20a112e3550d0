// push3d.cuh — device-side pieces shared by the 3-D push kernels (push3d.cu: slot-order kernel, push3d_brick.cu:
// brick-binned kernel): grid view, kernel arguments, the staggered trilinear gather, the box boundary, the Q32
// weights and the eight-node scatter.  Reference code restated: src/Field3D.hpp:40-65,110-163, src/fields3d.hpp:48-57.
#pragma once
#include "ctx.hpp"
#include "mcc.cuh"

struct Grid3Dev
{
    int M, K, N;                    // nodes along x, y, z
    int boundary, check_mask, deposit;
    double x_max, y_max, z_max;
    double idx, idy, idz;
    const double* gx;               // edge differences u[m] - u[m - stride] along x / y / z
    const double* gy;
    const double* gz;
    const unsigned char* cfree;     // per cell (indexed by its lowest node): any corner FREE
    unsigned long long* rho;        // this species' fixed-point charge grid
};

struct Push3Args
{
    Grid3Dev g;
    SpeciesDev s;
    ParticlesDev p;
    const MccBlob* mcc;
    unsigned long long* counts;
    unsigned long long* removed;
    unsigned long long seed;
    unsigned* coll_list;
    unsigned* coll_count;
    int deposit_runs;   // distinct cells per warp call that get the REDUX merge (0: every lane scatters on its own)
    // cell sort fused into the step (sort.cu): COUNT counts per cell, the next PERMUTE step draws slots and stores sorted
    int permute, count;
    unsigned* cursor;   // [cell] next free sorted slot of the cell (the scanned counts of the last COUNT push)
    ParticlesDev dst;
    unsigned* count_out;
};

namespace {

__device__ __forceinline__ unsigned long long q32_rn3(double w)
{
    const double magic = 6755399441055744.0;   // 1.5 * 2^52
    const double t = __dadd_rn(__dmul_rn(w, 4294967296.0), magic);
    return (unsigned long long)(__double_as_longlong(t) - __double_as_longlong(magic));
}

// round-to-nearest-even integer of a value that already carries the 2^32 scale
__device__ __forceinline__ unsigned long long q32_scaled(double w32)
{
    const double magic = 6755399441055744.0;   // 1.5 * 2^52
    const double t = __dadd_rn(w32, magic);
    return (unsigned long long)(__double_as_longlong(t) - __double_as_longlong(magic));
}

// one component of grad u at (X, Y, Z) in index units; DIR selects the differenced axis (0 x, 1 y, 2 z).
// The edge-difference arrays carry ghost planes ([M+1][K+1][N+1], k_edge_fields3d): along the differenced axis plane 0
// repeats plane 1 and plane M repeats plane M-1, which is exactly what Field3D::grad_component's clamp of the
// coordinate to [0.5, xmax*idx - 0.5] produces; along the other axes the plane past the end repeats the last one.
// No floating-point clamps remain in the particle loop (nine per particle before), only integer ones.
// trilinear interpolation of the eight differences g0..g7 (x fastest, then y, then z) with weights u, v, w: the nested form of
// the reference's eight-term sum (Field3D.hpp:133-141) — seven two-point interpolations, 14 FP64 operations instead of 39.  It
// agrees with the expanded sum to a few ulp of max|g| (each form rounds differently); the parity bar for trajectories is 1e-12.
__device__ __forceinline__ double trilerp(double u, double v, double w, double g0, double g1, double g2, double g3, double g4, double g5,
                                          double g6, double g7)
{
    const double cu = 1 - u, cv = 1 - v, cw = 1 - w;
    const double a0 = cu * g0 + u * g1, a1 = cu * g2 + u * g3, a2 = cu * g4 + u * g5, a3 = cu * g6 + u * g7;
    const double b0 = cv * a0 + v * a1, b1 = cv * a2 + v * a3;
    return cw * b0 + w * b1;
}

template <int DIR>
__device__ __forceinline__ double grad_component(const Grid3Dev& g, double X, double Y, double Z)
{
    const double xs = DIR == 0 ? X + 0.5 : X, ys = DIR == 1 ? Y + 0.5 : Y, zs = DIR == 2 ? Z + 0.5 : Z;
    const int i = max(min((int)xs, g.M - 1), 0), j = max(min((int)ys, g.K - 1), 0), k = max(min((int)zs, g.N - 1), 0);
    const double u = xs - i, v = ys - j, w = zs - k;
    const unsigned sj = (unsigned)g.N + 1u, si = ((unsigned)g.K + 1u) * sj;
    const double* f = (DIR == 0 ? g.gx : DIR == 1 ? g.gy : g.gz) + ((unsigned)i * si + (unsigned)j * sj + (unsigned)k);
    const double g0 = __ldg(f), g1 = __ldg(f + si), g2 = __ldg(f + sj), g3 = __ldg(f + si + sj);
    const double g4 = __ldg(f + 1), g5 = __ldg(f + si + 1), g6 = __ldg(f + sj + 1), g7 = __ldg(f + si + sj + 1);
    return trilerp(u, v, w, g0, g1, g2, g3, g4, g5, g6, g7) * (DIR == 0 ? g.idx : DIR == 1 ? g.idy : g.idz);
}

// box boundary and electrode absorption; node = lowest node of the particle's cell
__device__ __forceinline__ bool boundary3(const Grid3Dev& g, double& x, double& y, double& z, unsigned& node, unsigned* cell = nullptr)
{
    node = 0;
    if (!(x >= 0.0 && x <= g.x_max && y >= 0.0 && y <= g.y_max && z >= 0.0 && z <= g.z_max))
    {
        if (g.boundary == MAG2D_BOUNDARY_FREE || !(x == x && y == y && z == z)) return false;
        x = fmod(x, g.x_max); if (x < 0) x += g.x_max;
        y = fmod(y, g.y_max); if (y < 0) y += g.y_max;
        z = fmod(z, g.z_max); if (z < 0) z += g.z_max;
    }
    const double X = __dmul_rn(x, g.idx), Y = __dmul_rn(y, g.idy), Z = __dmul_rn(z, g.idz);
    const int i = max(min((int)X, g.M - 2), 0), j = max(min((int)Y, g.K - 2), 0), k = max(min((int)Z, g.N - 2), 0);
    const size_t m = ((size_t)i * g.K + j) * g.N + k;
    node = (unsigned)m;
    if (cell) *cell = ((unsigned)i * (unsigned)(g.K - 1) + (unsigned)j) * (unsigned)(g.N - 1) + (unsigned)k;
    if (g.check_mask && !g.cfree[m]) return false;
    return true;
}

// the eight Q32 weights of a particle that passed boundary3 (Field3D.hpp:56-64 order).  Computed right before the
// deposit from the stored position, so that the sixteen weight registers are not live across the push.
__device__ __forceinline__ void weights3(const Grid3Dev& g, double x, double y, double z, unsigned long long (&w)[8])
{
    const double X = __dmul_rn(x, g.idx), Y = __dmul_rn(y, g.idy), Z = __dmul_rn(z, g.idz);
    const int i = max(min((int)X, g.M - 2), 0), j = max(min((int)Y, g.K - 2), 0), k = max(min((int)Z, g.N - 2), 0);
    const double u = __dsub_rn(X, (double)i), v = __dsub_rn(Y, (double)j), t = __dsub_rn(Z, (double)k);
    const double cu = __dsub_rn(1.0, u), cv = __dsub_rn(1.0, v), ct = __dsub_rn(1.0, t);
    const double a00 = __dmul_rn(cu, cv), a10 = __dmul_rn(u, cv), a01 = __dmul_rn(cu, v), a11 = __dmul_rn(u, v);
    // q32_rn3(a * c) = rint(fl(a * c) * 2^32): the scaling by 2^32 is exact and commutes with the rounding of the product, so it is
    // folded into the z factors once instead of once per weight (bit-identical, eight DMUL fewer)
    const double ct32 = __dmul_rn(ct, 4294967296.0), t32 = __dmul_rn(t, 4294967296.0);
    w[0] = q32_scaled(__dmul_rn(a00, ct32));
    w[1] = q32_scaled(__dmul_rn(a10, ct32));
    w[2] = q32_scaled(__dmul_rn(a01, ct32));
    w[3] = q32_scaled(__dmul_rn(a11, ct32));
    w[4] = q32_scaled(__dmul_rn(a00, t32));
    w[5] = q32_scaled(__dmul_rn(a10, t32));
    w[6] = q32_scaled(__dmul_rn(a01, t32));
    w[7] = q32_scaled(__dmul_rn(a11, t32));
}

// eight RED.ADD.64 of one lane
__device__ __forceinline__ void scatter3(const Grid3Dev& g, unsigned node, const unsigned long long (&w)[8])
{
    const unsigned sj = (unsigned)g.N, si = (unsigned)(g.K * g.N);
    unsigned long long* r = g.rho + node;
    atomicAdd(r, w[0]);
    atomicAdd(r + si, w[1]);
    atomicAdd(r + sj, w[2]);
    atomicAdd(r + si + sj, w[3]);
    atomicAdd(r + 1, w[4]);
    atomicAdd(r + si + 1, w[5]);
    atomicAdd(r + sj + 1, w[6]);
    atomicAdd(r + si + sj + 1, w[7]);
}


}  // namespace
