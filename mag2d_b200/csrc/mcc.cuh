// mcc.cuh — null-collision Monte-Carlo collisions on the device.
// Restates BaseSpecies::scatter (reference src/particles.cpp:208-365) with a counter-based Philox
// stream per particle-step instead of the shared SHR3/ziggurat t_random (src/random.cpp), so the
// random streams differ from the reference: parity is statistical (tests/test_gpu_mcc.py).
#pragma once
#include "common.cuh"

// vec_interpolate::operator(), src/tabulate.cpp:124-140: clamped ends, binary search, lerp
// (the weight (x - x1) / (x2 - x1) is formed with the precomputed reciprocal interval width: one ulp off the division, which
// only the statistics of the collisions see)
__device__ __forceinline__ double table_lookup(const double* __restrict__ xd, const double* __restrict__ yd, const double* __restrict__ inv_w,
                                               int n, double x)
{
    if (x >= xd[n - 1]) return yd[n - 1];
    if (x <= xd[0]) return yd[0];
    int j1 = 0, j2 = n - 1;
    while (j2 - j1 > 1)
    {
        int j3 = (j1 + j2) >> 1;
        if (x < xd[j3]) j2 = j3;
        else j1 = j3;
    }
    const double w = (x - xd[j1]) * inv_w[j1];
    return yd[j1] * (1.0 - w) + yd[j2] * w;
}

// Interaction::sigma_v, src/particles.hpp:61-70 (COULOMB: src/particles.cpp:19-26 is not in the
// first bar — SURVEY.md "Hard parts"; it falls back to the constant rate)
__device__ __forceinline__ double sigma_v(const MccBlob* B, const MccInter& I, double v)
{
    if (I.n_table > 0)
    {
        double EeV = I.half_mu * v * v * (1.0 / MAG2D_QE);
        return table_lookup(B->tab + I.table_off, B->tab + B->n_tab + I.table_off, B->tab + 2 * B->n_tab + I.table_off, I.n_table, EeV) * v;
    }
    return I.rate;
}

// complete elliptic integral K(k) by the arithmetic-geometric mean (replaces std::tr1::comp_ellint_1,
// src/particles.cpp:346); 6 iterations reach 1e-16 for k <= 0.9999
__device__ __forceinline__ double ellint_K(double k)
{
    double a = 1.0, b = sqrt((1.0 - k) * (1.0 + k));
#pragma unroll 1
    for (int it = 0; it < 12; it++)
    {
        double an = 0.5 * (a + b);
        b = sqrt(a * b);
        a = an;
        if (fabs(a - b) <= 1e-16 * a) break;
    }
    return 1.5707963267948966 / a;
}

// t_random::deflect(angle, x, y, z), src/random.cpp:133-175: rotate by exactly `angle` about a random
// axis perpendicular to the vector (Boris-style rotation with |t| = tan(angle/2))
__device__ __forceinline__ void deflect(double angle, uint32_t w0, uint32_t w1, double& x, double& y, double& z)
{
    double len = tan(0.5 * angle);
    double x1, y1, z1;
    rot_iso(len, w0, w1, x1, y1, z1);
    double tx = y * z1 - z * y1;
    double ty = z * x1 - x * z1;
    double tz = x * y1 - y * x1;
    double tmp = len * rsqrt(tx * tx + ty * ty + tz * tz);
    tx *= tmp;
    ty *= tmp;
    tz *= tmp;
    double xp = x - y * tz + z * ty;
    double yp = y - z * tx + x * tz;
    double zp = z - x * ty + y * tx;
    tmp = 2.0 / (1.0 + len * len);
    x1 = tx * tmp;
    y1 = ty * tmp;
    z1 = tz * tmp;
    x += -yp * z1 + zp * y1;
    y += -zp * x1 + xp * z1;
    z += -xp * y1 + yp * x1;
}

// One null-collision event for the particle (vx, vy, vz).  Returns the process index inside
// interactions_by_species[target] or -1 for a null collision; target receives the species index.
// Consumes up to three Philox blocks.
static __device__ __noinline__ int mcc_scatter(const MccBlob* __restrict__ B, Rng& rng, double& pvx, double& pvy, double& pvz,
                                        int& target)
{
    const uint4 ra = rng.block();
    const double mass = B->mass;
    // target species by cumulative maximal rate; the last species is the fall-through (particles.cpp:214-222)
    double gamma = u01(ra.x) * B->inv_lifetime;
    double acc = 0.0;
    int specid = 0;
    const int nt = B->n_targets;
    for (; specid < nt - 1; specid++)
    {
        acc += B->t[specid].rate_max;
        if (acc > gamma) break;
    }
    target = specid;
    const MccTarget T = B->t[specid];
    double vr2, vz2, vt2;
    if (!T.pool)
    {
        const uint4 rb = rng.block();
        float n0, n1, n2, n3;
        normal2(rb.x, rb.y, n0, n1);
        normal2(rb.z, rb.w, n2, n3);
        vr2 = (double)n0 * T.vth;
        vz2 = (double)n1 * T.vth;
        vt2 = (double)n2 * T.vth;
    }
    else
    {
        // random live particle of the target species (BaseSpecies::random_particle, particles.hpp:207-217);
        // reads may be stale within a step, which the null-collision method tolerates
        const PoolDev P = B->pool[specid];
        uint4 rb = rng.block();
        long long i = (long long)(rb.x % (unsigned long long)P.n);
        int tries = 0;
        while (!particle_alive(P.x[i]) && tries < 64)
        {
            rb = rng.block();
            i = (long long)(rb.x % (unsigned long long)P.n);
            tries++;
        }
        // no live partner found in 65 draws (a nearly emptied store): null collision instead of a dead slot's velocities
        if (!particle_alive(P.x[i])) return -1;
        vr2 = P.vx[i];
        vz2 = P.vz[i];
        vt2 = P.vy[i];
    }
    const double dvx = pvx - vr2, dvz = pvz - vz2, dvy = pvy - vt2;
    const double v_rel = sqrt(dvx * dvx + dvz * dvz + dvy * dvy);
    const double m2 = T.mass;
    // process by cumulative n*sigma*v against the maximal rate; the remainder is a null collision
    gamma = u01(ra.y) * T.rate_max;
    acc = 0.0;
    int intid = 0;
    for (; intid < T.n_inter; intid++)
    {
        acc += sigma_v(B, B->in[T.first_inter + intid], v_rel) * T.density;
        if (acc > gamma) break;
    }
    if (intid == T.n_inter) return -1;
    const MccInter I = B->in[T.first_inter + intid];
    const double inv_M = T.inv_M;
    switch (I.type)
    {
        case 4:  // SUPERELASTIC: E' = E_rel + DE, isotropic in the centre-of-mass frame (particles.cpp:268-288)
        {
            double E = I.half_mu * v_rel * v_rel + I.DE;
            double v2 = sqrt(fmax(2.0 * E / I.mu, 0.0));
            double rx, ry, rz;
            rot_iso(v2, ra.z, ra.w, rx, ry, rz);
            pvx = (rx * m2 + pvx * mass + vr2 * m2) * inv_M;
            pvz = (rz * m2 + pvz * mass + vz2 * m2) * inv_M;
            pvy = (ry * m2 + pvy * mass + vt2 * m2) * inv_M;
            break;
        }
        case 3:  // COULOMB (partner update not in the first bar)
        case 0:  // ELASTIC: isotropic in the centre-of-mass frame (particles.cpp:290-316)
        {
            double cx = (pvx * mass + vr2 * m2) * inv_M;
            double cz = (pvz * mass + vz2 * m2) * inv_M;
            double cy = (pvy * mass + vt2 * m2) * inv_M;
            double rx, ry, rz;
            rot_iso(v_rel, ra.z, ra.w, rx, ry, rz);
            pvx = rx * m2 * inv_M + cx;
            pvz = rz * m2 * inv_M + cz;
            pvy = ry * m2 * inv_M + cy;
            break;
        }
        case 2:  // CX: take the partner's velocity (particles.cpp:318-325)
            pvx = vr2;
            pvz = vz2;
            pvy = vt2;
            break;
        case 1:  // LANGEVIN (Nanbu & Kitatani 1995), particles.cpp:327-360
        {
            double wx = dvx * m2 * inv_M;
            double wy = dvz * m2 * inv_M;
            double wz = dvy * m2 * inv_M;
            const uint4 rc = rng.block();
            double beta = sqrt(u01(rc.x)) * I.cutoff;
            if (beta > 1.0)
            {
                double b2 = beta * beta;
                double t = sqrt(b2 * b2 - 1.0);
                double xi0 = sqrt(b2 - t);
                double xi1 = sqrt(b2 + t);
                double theta = ellint_K(xi0 / xi1) * 1.4142135623730951 * beta / xi1;
                double chi = 3.141592653589793 - 2.0 * theta;
                deflect(chi, ra.z, ra.w, wx, wy, wz);
            }
            else
            {
                double len = sqrt(wx * wx + wy * wy + wz * wz);
                rot_iso(len, ra.z, ra.w, wx, wy, wz);
            }
            pvx = wx + (pvx * mass + vr2 * m2) * inv_M;
            pvz = wy + (pvz * mass + vz2 * m2) * inv_M;
            pvy = wz + (pvy * mass + vt2 * m2) * inv_M;
            break;
        }
        default: break;
    }
    return intid;
}

__device__ __forceinline__ void mcc_count(unsigned long long* counts, int n_targets, int target, int intid)
{
    if (!counts) return;
    if (intid < 0) atomicAdd(&counts[n_targets * 16 + target], 1ULL);
    else if (intid < 16) atomicAdd(&counts[target * 16 + intid], 1ULL);
}
