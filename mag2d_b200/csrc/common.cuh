// common.cuh — device-side parameter blocks, Philox, small helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MAG2D_FULL_MASK 0xffffffffu

// physical constants of the reference (src/param.cpp:8-10, old CODATA values kept for parity)
#define MAG2D_EPS0 8.854187817e-12
#define MAG2D_KB 1.380662e-23
#define MAG2D_QE 1.602189e-19

#define MCC_MAX_T 16    // target species per primary (== MAG2D_MAX_SPECIES)
#define MCC_MAX_I 32    // interactions per primary species
#define MCC_MAX_TAB 2048 // cross-section table points per primary species

// ------------------------------------------------------------------------------------------------
// Grid / field constants of one species step.  Passed by value (kernel parameter space).
struct GridDev
{
    int coord, boundary, M, N;
    double x_max, z_max;
    double idx, idz;
    int const_E;          // geometry == EMPTY && !selfconsistent: E = (0, extern_field)  (fields.hpp:126-131)
    int check_mask;       // !electric_field_from_file                                   (particles.hpp:395)
    int deposit;          // selfconsistent                                             (particles.hpp:408)
    int pad0;
    double extern_field;
    double dM1, dM2, dN1, dN2;   // (double)(M-1), (M-2), (N-1), (N-2): int->double conversions are slow
    // edge-centred field differences of this step's potential ue = u + phase*uRF (k_edge_fields):
    //   gx[i][j] = (ue[i][j] - ue[i-1][j]) * idx   (row 0 is zero)
    //   gz[i][j] = (ue[i][j] - ue[i][j-1]) * idz   (column 0 is zero)
    // exactly the g1..g4 terms of Field2D::grad, evaluated once per node instead of once per particle
    const double* __restrict__ gx;
    const double* __restrict__ gz;
    // cfree[i*N+j] = 1 when any corner of cell (i,j) is a FREE node: t_grid::is_free (fields.hpp:94-101)
    const unsigned char* __restrict__ cfree;
    unsigned long long* __restrict__ rho;      // fixed-point charge grid of this species, [M*N]
    // magnetic field table of Fields::load_magnetic_field (fields.cpp:870-959), nullptr while magnetic_field_const:
    // Br, Bz on their own bM x bN grid, row-major like Field2D; bilinear Field2D::interpolate per particle (fields.hpp:172-175)
    const double* __restrict__ b_r;
    const double* __restrict__ b_z;
    int bM, bN;
    double bidx, bidz, bxmin, bzmin;
};

// Species constants of one step, precomputed on the host with the reference's own expression
// order (particles.cpp:936-937, 959-975) so that they round identically.
struct SpeciesDev
{
    double dt;
    double hq;            // (charge/mass*dt)/2  — exact halving of qmdt
    double tx, ty, tz;    // Boris t = B * charge*dt/(2 mass)   (x, out-of-plane y, z)
    double sx, sy, sz;    // s = t * 2/(1+|t|^2)
    double tb;            // charge*dt/(2 mass): t = B * tb per particle when B comes from the table
    int has_B;
    int species;          // index, part of the RNG key
    double prob;          // 1 - exp(-dt/lifetime)
    unsigned long long prob_u32;   // the same test on the raw Philox word: u01(w) < prob  <=>  w < prob_u32 (bernoulli_threshold)
    double lifetime;
    double qm;            // charge/mass (multi-collision mover)
    unsigned long long step;   // species step counter (niter), part of the RNG counter
};

// SoA particle arrays of one species (device pointers, capacity elements each)
// u01(w) = (w + 0.5) 2^-32 < prob  <=>  w < ceil(prob 2^32 - 0.5): the Bernoulli test of the null-collision method as one
// integer compare instead of an I2F.F64 + DFMA + DSETP chain per particle
inline unsigned long long bernoulli_threshold(double prob)
{
    if (!(prob > 0.0)) return 0ULL;
    const double t = ceil(prob * 4294967296.0 - 0.5);
    return t >= 4294967296.0 ? 4294967296ULL : (t <= 0.0 ? 0ULL : (unsigned long long)t);
}

struct ParticlesDev
{
    double* __restrict__ x;
    double* __restrict__ z;
    double* __restrict__ vx;
    double* __restrict__ vy;
    double* __restrict__ vz;
    double* __restrict__ y;     // CARTESIAN3D only
    double* __restrict__ ttd;   // ADVANCE_MULTICOLL only (time_to_death)
    long long n;                // slots in use
};

// Element access of the particle arrays in their STORAGE type T: double, or float in the fp32 storage mode
// (mag2d_set_storage; the arithmetic stays fp64, only what lives in HBM is rounded: 40 instead of 80 bytes per 2D3V
// particle-step).  The array pointers keep the type double* everywhere; T says how the bytes behind them are laid out.
template <typename T> __device__ __forceinline__ double pld(const double* base, long long k) { return (double)reinterpret_cast<const T*>(base)[k]; }
template <typename T> __device__ __forceinline__ void pst(double* base, long long k, double v) { reinterpret_cast<T*>(base)[k] = (T)v; }
template <typename T> __device__ __forceinline__ double2 pld2(const double* base, long long k);      // k even: one 128-bit / 64-bit load
template <> __device__ __forceinline__ double2 pld2<double>(const double* base, long long k) { return *reinterpret_cast<const double2*>(base + k); }
template <> __device__ __forceinline__ double2 pld2<float>(const double* base, long long k)
{
    const float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + k);
    return make_double2((double)f.x, (double)f.y);
}
template <typename T> __device__ __forceinline__ void pst2(double* base, long long k, double a, double b);
template <> __device__ __forceinline__ void pst2<double>(double* base, long long k, double a, double b) { *reinterpret_cast<double2*>(base + k) = make_double2(a, b); }
template <> __device__ __forceinline__ void pst2<float>(double* base, long long k, double a, double b)
{
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + k) = make_float2((float)a, (float)b);
}
// what a value becomes once it is stored (identity for double): positions are rounded BEFORE the boundary test and the
// deposit, so that the charge grid is the exact fixed-point deposit of the stored positions
template <typename T> __device__ __forceinline__ double stored(double v) { return (double)(T)v; }

// A removed particle keeps its slot until the next sort; it is marked by x = NaN.
__device__ __forceinline__ bool particle_alive(double x) { return x == x; }
__device__ __forceinline__ double dead_marker() { return __longlong_as_double(0x7ff8000000000000LL); }

// ------------------------------------------------------------------------------------------------
// Collision model of one primary species (BaseSpecies::{speclist, interactions_by_species,
// rates_by_species, lifetime}, src/particles.hpp:109-127), flattened for the device.
struct MccInter
{
    int type, n_table, table_off, pad;
    double DE;        // J
    double rate;      // sigma*v for table-less processes (LANGEVIN already * cutoff^2)
    double cutoff;
    double half_mu;   // 0.5 * m1 m2/(m1+m2)
    double mu;
};
struct MccTarget
{
    double rate_max;  // rates_by_species[k]
    double density;
    double mass;
    double vth;       // v_max * M_SQRT1_2 scale of the Maxwellian partner (particles.hpp:188-193)
    int n_inter, first_inter;
    int pool;         // 1: partners are drawn from the target's particle array (particles.cpp:230-238)
    int pad;
    double inv_M;     // 1 / (primary mass + target mass)
};
struct PoolDev
{
    const double* x;  // alive marker
    const double* vx;
    const double* vy;
    const double* vz;
    long long n;
};
struct MccBlob
{
    int n_targets, n_inter_total, n_tab, has_collisions;
    double lifetime, inv_lifetime, mass, charge;
    MccTarget t[MCC_MAX_T];
    MccInter in[MCC_MAX_I];
    PoolDev pool[MCC_MAX_T];
    double tab[3 * MCC_MAX_TAB];   // energies [0, n_tab), cross sections [n_tab, 2 n_tab), 1 / (E[j+1] - E[j]) [2 n_tab, 3 n_tab)
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (slot, step_lo, step_hi|species, draw), key = seed.
struct PhiloxKey { uint32_t k0, k1; };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++)
    {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += W0;
        k1 += W1;
    }
    return c;
}

struct Rng
{
    uint32_t k0, k1;     // seed
    uint32_t c0, c1, c2; // slot, step_lo, step_hi ^ (species << 24)
    uint32_t draw;       // block counter within this particle-step
    __device__ __forceinline__ uint4 block() { return philox4x32_10(make_uint4(c0, c1, c2, draw++), k0, k1); }
};

__device__ __forceinline__ Rng make_rng(uint64_t seed, int species, unsigned long long step, unsigned long long slot)
{
    Rng r;
    r.k0 = (uint32_t)seed;
    r.k1 = (uint32_t)(seed >> 32);
    r.c0 = (uint32_t)slot;
    r.c1 = (uint32_t)step;
    r.c2 = (uint32_t)(step >> 32) ^ ((uint32_t)species << 24) ^ ((uint32_t)(slot >> 32) << 16);
    r.draw = 0;
    return r;
}

// uniform in (0,1), 32-bit resolution (the reference's uni() is a 24-bit float, random.cpp:42)
__device__ __forceinline__ double u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
// uniform float in (0,1), 23-bit, never rounds to 0 or 1
__device__ __forceinline__ float u01f(uint32_t w) { return ((float)(w >> 9) + 0.5f) * 1.1920928955078125e-7f; }

// two standard normals from two words (Box-Muller in float: the reference's rnor() is a float ziggurat)
__device__ __forceinline__ void normal2(uint32_t a, uint32_t b, float& n0, float& n1)
{
    float r = sqrtf(-2.0f * logf(u01f(a)));
    float s, c;
    sincospif(2.0f * u01f(b), &s, &c);
    n0 = r * c;
    n1 = r * s;
}
// exponential variate (t_random::rexp, random.cpp:54-57)
__device__ __forceinline__ double rexp1(uint32_t a) { return (double)(-logf(u01f(a))); }

// unit-circle point for azimuth 2*pi*u: float sincospi renormalised in double so that the rotated
// vector keeps its length to ~1e-14 (the reference calls double sincos, random.cpp:110)
__device__ __forceinline__ void unit_circle(uint32_t w, double& sp, double& cp)
{
    float sf, cf;
    sincospif(2.0f * u01f(w), &sf, &cf);
    double s = sf, c = cf;
    double corr = 1.5 - 0.5 * (s * s + c * c);
    sp = s * corr;
    cp = c * corr;
}

// isotropic vector of length len: t_random::rot(len, x, y, z), random.cpp:103-115
__device__ __forceinline__ void rot_iso(double len, uint32_t w0, uint32_t w1, double& x, double& y, double& z)
{
    double ct = 1.0 - 2.0 * u01(w0);
    x = len * ct;
    double st = sqrt(fmax(1.0 - ct * ct, 0.0));
    double sp, cp;
    unit_circle(w1, sp, cp);
    y = len * st * sp;
    z = len * st * cp;
}

__device__ __forceinline__ unsigned lane_id()
{
    unsigned l;
    asm("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

constexpr unsigned SORT_INVALID_KEY = 0xFFFFFFFFu;

// warp-aggregated ticket: one atomic per (warp, cell), ranks handed out in lane order.  MATCH.ANY finds the lanes
// that share a cell in one instruction, so all group leaders issue their atomics together: one memory round trip
// per call, not one per distinct cell (the returning atomic is the long pole of a COUNT step).
// The fused cell sort needs no stored tickets: a COUNT push only counts the particles per cell (one RED per warp and cell),
// and the next (PERMUTE) push — which loads every particle at exactly the position it was counted at — recomputes the
// cell and draws the slot from a per-cell cursor (the scanned counts).  Saves 16 bytes of ticket traffic per particle
// and sort and the two key / rank arrays per species (16 bytes per slot).
// count[key] += number of lanes holding key (one non-returning atomic per distinct key of the warp)
__device__ __forceinline__ void warp_count(unsigned* __restrict__ count, bool valid, unsigned key)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned group = __match_any_sync(MAG2D_FULL_MASK, valid ? key : (0xFFFFFFE0u | lane));
    if (valid && (int)lane == __ffs(group) - 1) atomicAdd(&count[key], (unsigned)__popc(group));
}

__device__ __forceinline__ unsigned warp_ticket(unsigned* __restrict__ count, bool valid, unsigned key)
{
    const unsigned lane = lane_id();
    // invalid lanes get distinct pseudo-keys so that they never join a group of live particles
    const unsigned group = __match_any_sync(MAG2D_FULL_MASK, valid ? key : (0xFFFFFFE0u | lane));
    const int leader = __ffs(group) - 1;
    unsigned base = 0;
    if (valid && (int)lane == leader) base = atomicAdd(&count[key], (unsigned)__popc(group));
    base = __shfl_sync(MAG2D_FULL_MASK, base, leader);
    return base + __popc(group & ((1u << lane) - 1u));
}

