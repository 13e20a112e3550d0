/* oracle/mag2d_oracle.c — TEST INFRASTRUCTURE ONLY.  See mag2d_oracle.h.
 *
 * Plain-C restatement of the reference's hot path.  Build with -O2 -ffp-contract=off so that every
 * expression rounds exactly like the reference's parity build (oracle/_ref/libmag2d_ref_parity.so);
 * tests/test_oracle_vs_reference.py checks bit-equality of the deterministic pieces and of the
 * RNG-driven pieces under the same SHR3 seed.
 */
#define _GNU_SOURCE
#include "mag2d_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline double sqr(double x) { return x * x; }
/* mymath.cpp:77 */
static inline double norm3(double x, double y, double z) { return sqrt(sqr(x) + sqr(y) + sqr(z)); }
/* mymath.cpp:71-76 */
static inline double mod_ref(double x, double y)
{
    if (x >= 0.0 && x <= y) return x;
    return x - y * (int)(x / y) + (x < 0 ? y : 0);
}
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ gather */

void orc_grad(const double* data, int jmax, int lmax, double idx, double idy, double xmin, double ymin,
              double x, double y, double* grad_x, double* grad_y)
{
#define D(i, j) data[(size_t)(i) * (size_t)lmax + (size_t)(j)]
    int i, j;
    double g1, g2, g3, g4, fx, fy;
    x -= xmin;
    y -= ymin;
    /* x component: differences live on x-edges, so the stencil is shifted by half a cell in x */
    i = (int)(x * idx + 0.5);
    j = (int)(y * idy);
    j = imin(j, lmax - 2);
    if (i > 0 && i < jmax - 1)
    {
        g1 = (D(i, j) - D(i - 1, j)) * idx;
        g2 = (D(i, j + 1) - D(i - 1, j + 1)) * idx;
        g3 = (D(i + 1, j + 1) - D(i, j + 1)) * idx;
        g4 = (D(i + 1, j) - D(i, j)) * idx;
        fx = x * idx - i + .5;
        fy = y * idy - j;
        *grad_x = g1 * (1 - fx) * (1 - fy) + g2 * (1 - fx) * fy + g3 * fx * fy + g4 * fx * (1 - fy);
    }
    else if (i == jmax - 1)
    {
        g1 = (D(i, j) - D(i - 1, j)) * idx;
        g2 = (D(i, j + 1) - D(i - 1, j + 1)) * idx;
        fy = y * idy - j;
        *grad_x = g1 * (1 - fy) + g2 * fy;
    }
    else if (i == 0)
    {
        g3 = (D(i + 1, j + 1) - D(i, j + 1)) * idx;
        g4 = (D(i + 1, j) - D(i, j)) * idx;
        fy = y * idy - j;
        *grad_x = g3 * fy + g4 * (1 - fy);
    }
    /* y component */
    i = (int)(x * idx);
    j = (int)(y * idy + 0.5);
    i = imin(i, jmax - 2);
    if (j > 0 && j < lmax - 1)
    {
        g1 = (D(i, j) - D(i, j - 1)) * idy;
        g2 = (D(i + 1, j) - D(i + 1, j - 1)) * idy;
        g3 = (D(i + 1, j + 1) - D(i + 1, j)) * idy;
        g4 = (D(i, j + 1) - D(i, j)) * idy;
        fx = x * idx - i;
        fy = y * idy - j + 0.5;
        *grad_y = g1 * (1 - fx) * (1 - fy) + g2 * (1 - fy) * fx + g3 * fx * fy + g4 * fy * (1 - fx);
    }
    else if (j == lmax - 1)
    {
        g1 = (D(i, j) - D(i, j - 1)) * idy;
        g2 = (D(i + 1, j) - D(i + 1, j - 1)) * idy;
        fx = x * idx - i;
        *grad_y = g1 * (1 - fx) + g2 * fx;
    }
    else if (j == 0)
    {
        g3 = (D(i + 1, j + 1) - D(i + 1, j)) * idy;
        g4 = (D(i, j + 1) - D(i, j)) * idy;
        fx = x * idx - i;
        *grad_y = g3 * fx + g4 * (1 - fx);
    }
#undef D
}

double orc_interpolate(const double* data, int jmax, int lmax, double idx, double idy, double xmin,
                       double ymin, double x, double y)
{
    x -= xmin;
    y -= ymin;
    int i = (int)(x * idx);
    int j = (int)(y * idy);
    double u = x * idx - i;
    double v = y * idy - j;
    if (i < 0 || i > jmax - 1 || j < 0 || j > lmax - 1) return NAN;
    const double* d = data + (size_t)i * lmax + j;
    return (1 - u) * (1 - v) * d[0] + u * (1 - v) * d[lmax] + (1 - u) * v * d[1] + u * v * d[lmax + 1];
}

void orc_field_E(const orc_grid* g, const double* u, const double* uRF, double x, double y, double time,
                 double* Ex, double* Ez)
{
    if (g->geometry_empty && !g->selfconsistent)
    {
        *Ex = 0.;
        *Ez = g->extern_field;
        return;
    }
    double gx = *Ex, gy = *Ez;
    orc_grad(u, g->M, g->N, g->idx, g->idz, 0.0, 0.0, x, y, &gx, &gy);
    if (g->rf)
    {
        double rx = 0, ry = 0;
        orc_grad(uRF, g->M, g->N, g->idx, g->idz, 0.0, 0.0, x, y, &rx, &ry);
        double phase = mod_ref(g->rf_omega * time, 10000000 * M_PI);
        phase = g->rf_amplitude * cos(phase) + g->rf_U0;
        gx += rx * phase;
        gy += ry * phase;
    }
    *Ex = -gx;
    *Ez = -gy;
}

/* ------------------------------------------------------------------ movers */

void orc_boris_cart(double charge, double mass, double dt, double fx, double fz, double Bx, double Bz,
                    double By, double* x, double* z, double* vx, double* vy, double* vz)
{
    const double qmdt = charge / mass * dt;
    *vx += fx * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
    double tmp = charge * dt / (2.0 * mass);
    double tx = Bx * tmp, ty = By * tmp, tz = Bz * tmp;
    /* cross product in the right-handed (x, z, y) triad: signs flipped w.r.t. the cylindrical mover */
    double vprime_r = *vx - *vy * tz + *vz * ty;
    double vprime_z = *vz - *vx * ty + *vy * tx;
    double vprime_t = *vy - *vz * tx + *vx * tz;
    tmp = 2.0 / (1 + sqr(tx) + sqr(ty) + sqr(tz));
    double sx = tx * tmp, sy = ty * tmp, sz = tz * tmp;
    *vx = *vx - vprime_t * sz + vprime_z * sy;
    *vy = *vy - vprime_z * sx + vprime_r * sz;
    *vz = *vz - vprime_r * sy + vprime_t * sx;
    *vx += fx * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
    *x += *vx * dt;
    *z += *vz * dt;
}

void orc_boris_cart_init(double charge, double mass, double dt, double fx, double fz, double Bx, double Bz,
                         double By, double* vx, double* vy, double* vz)
{
    const double qmdt = -charge / mass * dt;
    double tmp = -0.5 * charge * dt / (2.0 * mass);
    double tx = Bx * tmp, ty = By * tmp, tz = Bz * tmp;
    double vprime_r = *vx - *vy * tz + *vz * ty;
    double vprime_z = *vz - *vx * ty + *vy * tx;
    double vprime_t = *vy - *vz * tx + *vx * tz;
    tmp = 2.0 / (1 + sqr(tx) + sqr(ty) + sqr(tz));
    double sx = tx * tmp, sy = ty * tmp, sz = tz * tmp;
    *vx = *vx - vprime_t * sz + vprime_z * sy;
    *vy = *vy - vprime_z * sx + vprime_r * sz;
    *vz = *vz - vprime_r * sy + vprime_t * sx;
    *vx += fx * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
}

void orc_boris_cyl(double charge, double mass, double dt, double fr, double fz, double Bx, double Bz,
                   double By, double* r, double* z, double* vr, double* vt, double* vz)
{
    const double qmdt = charge / mass * dt;
    *vr += fr * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
    double tmp = charge * dt / (2.0 * mass);
    double tx = Bx * tmp, ty = By * tmp, tz = Bz * tmp;
    double vprime_r = *vr + *vt * tz - *vz * ty;
    double vprime_t = *vt + *vz * tx - *vr * tz;
    double vprime_z = *vz + *vr * ty - *vt * tx;
    tmp = 2.0 / (1 + sqr(tx) + sqr(ty) + sqr(tz));
    double sx = tx * tmp, sy = ty * tmp, sz = tz * tmp;
    *vr = *vr + vprime_t * sz - vprime_z * sy;
    *vt = *vt + vprime_z * sx - vprime_r * sz;
    *vz = *vz + vprime_r * sy - vprime_t * sx;
    *vr += fr * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
    /* drift in the local Cartesian frame, then rotate the frame back onto the new radius */
    double x2 = *r + *vr * dt;
    double y2 = *vt * dt;
    *r = sqrt(sqr(x2) + sqr(y2));
    *z += *vz * dt;
    double sa = y2 / *r;
    double ca = x2 / *r;
    if (*r == 0)
    {
        sa = 0;
        ca = 1;
    }
    tmp = *vr;
    *vr = ca * *vr + sa * *vt;
    *vt = -sa * tmp + ca * *vt;
}

void orc_boris_cyl_init(double charge, double mass, double dt, double fr, double fz, double Bx, double Bz,
                        double By, double* vr, double* vt, double* vz)
{
    const double qmdt = -1.0 * charge / mass * dt;
    double tmp = -0.5 * charge * dt / (2.0 * mass);
    double tx = Bx * tmp, ty = By * tmp, tz = Bz * tmp;
    double vprime_r = *vr + *vt * tz - *vz * ty;
    double vprime_t = *vt + *vz * tx - *vr * tz;
    double vprime_z = *vz + *vr * ty - *vt * tx;
    tmp = 2.0 / (1 + sqr(tx) + sqr(ty) + sqr(tz));
    double sx = tx * tmp, sy = ty * tmp, sz = tz * tmp;
    *vr = *vr + vprime_t * sz - vprime_z * sy;
    *vt = *vt + vprime_z * sx - vprime_r * sz;
    *vz = *vz + vprime_r * sy - vprime_t * sx;
    *vr += fr * qmdt / 2.0;
    *vz += fz * qmdt / 2.0;
}

/* --------------------------------------------------------------------- RNG */
/* The reference mixes float and double arithmetic through C++ overload resolution (std::log(float),
 * std::sqrt(float), int*float ...); the casts below reproduce those types. */

uint32_t orc_rng_iuni(orc_rng* r)
{
    r->jz = r->jsr;
    r->jsr ^= (r->jsr << 13);
    r->jsr ^= (r->jsr >> 17);
    r->jsr ^= (r->jsr << 5);
    return r->jz + r->jsr;
}

float orc_rng_uni(orc_rng* r) { return (float)(.5 + (int32_t)orc_rng_iuni(r) * 2.3283064365386963e-10); }

void orc_rng_seed(orc_rng* r, uint32_t seed)
{
    r->jsr = 123456789u;
    r->jsr ^= seed;
    r->z = orc_rng_iuni(r);
    r->w = orc_rng_iuni(r);
    r->jcong = orc_rng_iuni(r);
}

void orc_rng_init(orc_rng* r, uint32_t seed)
{
    memset(r, 0, sizeof(*r));
    orc_rng_seed(r, seed);
    /* ziggurat layer tables (Marsaglia & Tsang 2000) */
    const double m1 = 2147483648.0, m2 = 4294967296.;
    double dn = 3.442619855899, tn = dn, vn = 9.91256303526217e-3, q;
    double de = 7.697117470131487, te = de, ve = 3.949659822581572e-3;
    int i;
    q = vn / exp(-.5 * dn * dn);
    r->kn[0] = (uint32_t)((dn / q) * m1);
    r->kn[1] = 0;
    r->wn[0] = (float)(q / m1);
    r->wn[127] = (float)(dn / m1);
    r->fn[0] = 1.f;
    r->fn[127] = (float)exp(-.5 * dn * dn);
    for (i = 126; i >= 1; i--)
    {
        dn = sqrt(-2. * log(vn / dn + exp(-.5 * dn * dn)));
        r->kn[i + 1] = (uint32_t)((dn / tn) * m1);
        tn = dn;
        r->fn[i] = (float)exp(-.5 * dn * dn);
        r->wn[i] = (float)(dn / m1);
    }
    q = ve / exp(-de);
    r->ke[0] = (uint32_t)((de / q) * m2);
    r->ke[1] = 0;
    r->we[0] = (float)(q / m2);
    r->we[255] = (float)(de / m2);
    r->fe[0] = 1.f;
    r->fe[255] = (float)exp(-de);
    for (i = 254; i >= 1; i--)
    {
        de = -log(ve / de + exp(-de));
        r->ke[i + 1] = (uint32_t)((de / te) * m2);
        te = de;
        r->fe[i] = (float)exp(-de);
        r->we[i] = (float)(de / m2);
    }
}

static float orc_nfix(orc_rng* r)
{
    const float rr = 3.442620f;
    for (;;)
    {
        r->nfix_x = r->hz * r->wn[r->iz];
        if (r->iz == 0)
        {
            do
            {
                r->nfix_x = (float)(-logf(orc_rng_uni(r)) * 0.2904764);
                r->nfix_y = -logf(orc_rng_uni(r));
            } while (r->nfix_y + r->nfix_y < r->nfix_x * r->nfix_x);
            return (r->hz > 0) ? rr + r->nfix_x : -rr - r->nfix_x;
        }
        {
            /* operand order as in the reference: fn[iz] + uni()*(fn[iz-1]-fn[iz]) */
            float u = orc_rng_uni(r);
            float lhs = r->fn[r->iz] + u * (r->fn[r->iz - 1] - r->fn[r->iz]);
            if (lhs < exp(-.5 * r->nfix_x * r->nfix_x)) return r->nfix_x;
        }
        r->hz = (int32_t)orc_rng_iuni(r);
        r->iz = r->hz & 127;
        if (fabs((double)r->hz) < r->kn[r->iz]) return (r->hz * r->wn[r->iz]);
    }
}

float orc_rng_rnor(orc_rng* r)
{
    r->hz = (int32_t)orc_rng_iuni(r);
    r->iz = r->hz & 127;
    return ((uint32_t)abs(r->hz) < r->kn[r->iz]) ? r->hz * r->wn[r->iz] : orc_nfix(r);
}

static float orc_efix(orc_rng* r)
{
    float x;
    for (;;)
    {
        if (r->iz == 0) return (float)(7.69711 - logf(orc_rng_uni(r)));
        x = r->jz * r->we[r->iz];
        {
            float u = orc_rng_uni(r);
            float lhs = r->fe[r->iz] + u * (r->fe[r->iz - 1] - r->fe[r->iz]);
            if (lhs < expf(-x)) return x;
        }
        r->jz = orc_rng_iuni(r);
        r->iz = (r->jz & 255);
        if (r->jz < r->ke[r->iz]) return (r->jz * r->we[r->iz]);
    }
}

float orc_rng_rexp(orc_rng* r)
{
    r->jz = orc_rng_iuni(r);
    r->iz = r->jz & 255;
    return (r->jz < r->ke[r->iz]) ? r->jz * r->we[r->iz] : orc_efix(r);
}

double orc_rng_radius(orc_rng* r) { return sqrtf(orc_rng_uni(r)); }

void orc_rng_rot(orc_rng* r, double len, double* x, double* y, double* z)
{
    double cs_theta = (1 - 2 * orc_rng_uni(r)); /* float arithmetic, widened on assignment */
    *x = len * cs_theta;
    cs_theta = sqrt(1 - sqr(cs_theta));
    double sp, cp;
    sincos(2 * M_PI * orc_rng_uni(r), &sp, &cp);
    *y = len * cs_theta * sp;
    *z = len * cs_theta * cp;
}

void orc_rng_rot_inplace(orc_rng* r, double* x, double* y, double* z)
{
    double len = sqrt(sqr(*x) + sqr(*y) + sqr(*z));
    orc_rng_rot(r, len, x, y, z);
}

void orc_rng_deflect(orc_rng* r, double angle, double* x, double* y, double* z)
{
    double len = tan(angle / 2.0);
    double x1, y1, z1;
    double tmp = (1 - 2 * orc_rng_uni(r));
    x1 = len * tmp;
    tmp = sqrt(1 - sqr(tmp));
    double sp, cp;
    sincos(2 * M_PI * orc_rng_uni(r), &sp, &cp);
    y1 = len * tmp * sp;
    z1 = len * tmp * cp;
    /* rotation axis = v x (random vector), rescaled to length tan(angle/2) */
    double tx = *y * z1 - *z * y1;
    double ty = *z * x1 - *x * z1;
    double tz = *x * y1 - *y * x1;
    tmp = len / norm3(tx, ty, tz);
    tx *= tmp;
    ty *= tmp;
    tz *= tmp;
    double xprime = *x - *y * tz + *z * ty;
    double yprime = *y - *z * tx + *x * tz;
    double zprime = *z - *x * ty + *y * tx;
    tmp = 2.0 / (1 + len * len);
    x1 = tx * tmp;
    y1 = ty * tmp;
    z1 = tz * tmp;
    *x += -yprime * z1 + zprime * y1;
    *y += -zprime * x1 + xprime * z1;
    *z += -xprime * y1 + yprime * x1;
}

void orc_rng_draw(orc_rng* r, int kind, int n, double* out)
{
    for (int k = 0; k < n; k++)
    {
        switch (kind)
        {
            case 0: out[k] = orc_rng_uni(r); break;
            case 1: out[k] = orc_rng_rnor(r); break;
            case 2: out[k] = orc_rng_rexp(r); break;
            case 3: out[k] = orc_rng_iuni(r); break;
            default: out[k] = orc_rng_radius(r); break;
        }
    }
}

double orc_ellint_K(double k)
{
    /* K(k) = pi / (2 AGM(1, sqrt(1-k^2))) */
    double a = 1.0, b = sqrt((1.0 - k) * (1.0 + k));
    for (int it = 0; it < 40; it++)
    {
        double an = 0.5 * (a + b);
        b = sqrt(a * b);
        a = an;
        if (fabs(a - b) <= 1e-17 * a) break;
    }
    return M_PI / (2.0 * a);
}

double orc_langevin_chi(double beta)
{
    double tmp = sqrt(beta * beta * beta * beta - 1.0);
    double xi0 = sqrt(beta * beta - tmp);
    double xi1 = sqrt(beta * beta + tmp);
    double zeta = xi0 / xi1;
    double theta = orc_ellint_K(zeta) * M_SQRT2 * beta / xi1;
    return M_PI - 2 * theta;
}

/* ------------------------------------------------------------------- model */

typedef struct
{
    int type;
    double DE, rate, cutoff;
    int primary, secondary;
    int n;
    double *E, *sigma;
} orc_inter;

typedef struct
{
    int type;
    double mass, charge, density, temperature, E_max, dt, v_max, lifetime;
    int* n_inter;        /* [n_species] */
    orc_inter*** inter;  /* [n_species][n_inter] */
    double* rates;       /* [n_species] */
    int pool_n;
    const double *pool_vx, *pool_vy, *pool_vz;
    const unsigned char* pool_alive;
} orc_spec;

struct orc_model
{
    int ns;
    orc_spec* s;
};

orc_model* orc_model_new(int n_species)
{
    orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
    m->ns = n_species;
    m->s = (orc_spec*)calloc((size_t)n_species, sizeof(orc_spec));
    for (int i = 0; i < n_species; i++)
    {
        m->s[i].n_inter = (int*)calloc((size_t)n_species, sizeof(int));
        m->s[i].inter = (orc_inter***)calloc((size_t)n_species, sizeof(orc_inter**));
        m->s[i].rates = (double*)calloc((size_t)n_species, sizeof(double));
        m->s[i].lifetime = INFINITY;
    }
    return m;
}

void orc_model_free(orc_model* m)
{
    if (!m) return;
    for (int i = 0; i < m->ns; i++)
    {
        for (int k = 0; k < m->ns; k++)
        {
            for (int q = 0; q < m->s[i].n_inter[k]; q++)
            {
                free(m->s[i].inter[k][q]->E);
                free(m->s[i].inter[k][q]->sigma);
                free(m->s[i].inter[k][q]);
            }
            free(m->s[i].inter[k]);
        }
        free(m->s[i].n_inter);
        free(m->s[i].inter);
        free(m->s[i].rates);
    }
    free(m->s);
    free(m);
}

int orc_model_set_species(orc_model* m, int i, int type, double mass, double charge, double density,
                          double temperature, double E_max, double dt)
{
    if (i < 0 || i >= m->ns) return 1;
    orc_spec* s = &m->s[i];
    s->type = type;
    s->mass = mass;
    s->charge = charge;
    s->density = density;
    s->temperature = temperature;
    s->E_max = E_max > 0. ? E_max : temperature * ORC_KB / ORC_QE * 10.0;
    s->dt = dt;
    s->v_max = sqrt(2.0 * ORC_KB * temperature / mass);
    s->lifetime = INFINITY;
    return 0;
}

int orc_model_add_interaction(orc_model* m, int type, double DE_eV, double rate, double cutoff, int primary,
                              int secondary, int n, const double* E_eV, const double* sigma)
{
    if (primary < 0 || primary >= m->ns || secondary < 0 || secondary >= m->ns) return 1;
    orc_inter* I = (orc_inter*)calloc(1, sizeof(orc_inter));
    I->type = type;
    I->DE = DE_eV * ORC_QE;
    I->rate = rate;
    I->cutoff = cutoff;
    I->primary = primary;
    I->secondary = secondary;
    I->n = n;
    if (n > 0)
    {
        I->E = (double*)malloc(sizeof(double) * (size_t)n);
        I->sigma = (double*)malloc(sizeof(double) * (size_t)n);
        memcpy(I->E, E_eV, sizeof(double) * (size_t)n);
        memcpy(I->sigma, sigma, sizeof(double) * (size_t)n);
    }
    if (type == ORC_LANGEVIN) I->rate *= sqr(cutoff);
    orc_spec* s = &m->s[primary];
    int k = s->n_inter[secondary]++;
    s->inter[secondary] = (orc_inter**)realloc(s->inter[secondary], sizeof(orc_inter*) * (size_t)(k + 1));
    s->inter[secondary][k] = I;
    return 0;
}

void orc_model_set_pool(orc_model* m, int i, int n_slots, const double* vx, const double* vy,
                        const double* vz, const unsigned char* alive)
{
    m->s[i].pool_n = n_slots;
    m->s[i].pool_vx = vx;
    m->s[i].pool_vy = vy;
    m->s[i].pool_vz = vz;
    m->s[i].pool_alive = alive;
}

double orc_table_lookup(int n, const double* xdata, const double* ydata, double x)
{
    if (x >= xdata[n - 1]) return ydata[n - 1];
    if (x <= xdata[0]) return ydata[0];
    int j1 = 0, j2 = n - 1, j3;
    while ((j2 - j1) > 1)
    {
        j3 = (j1 + j2) / 2;
        if (x < xdata[j3]) j2 = j3;
        else j1 = j3;
    }
    double w = (x - xdata[j1]) / (xdata[j2] - xdata[j1]);
    return ydata[j1] * (1 - w) + ydata[j2] * w;
}

static double inter_mu(const orc_model* m, const orc_inter* I)
{
    double m1 = m->s[I->primary].mass, m2 = m->s[I->secondary].mass;
    return m1 * m2 / (m1 + m2);
}
/* Interaction::EeV / E / v_rel, src/particles.hpp:415-433 */
static double inter_EeV(const orc_model* m, const orc_inter* I, double v) { return 0.5 * inter_mu(m, I) * v * v / ORC_QE; }
static double inter_E(const orc_model* m, const orc_inter* I, double v) { return 0.5 * inter_mu(m, I) * v * v; }
static double inter_vrel(const orc_model* m, const orc_inter* I, double E) { return sqrt(2 * E / inter_mu(m, I)); }

/* Interaction::coulomb_sigma, src/particles.cpp:19-26 */
static double coulomb_sigma(const orc_model* m, const orc_inter* I, double E)
{
    const orc_spec* p = &m->s[I->primary];
    const orc_spec* s = &m->s[I->secondary];
    E *= ORC_QE;
    double lambda_D = sqrt(ORC_EPS0 * ORC_KB * p->temperature / (p->density * p->charge * p->charge));
    double Lambda = p->charge * s->charge / (4 * M_PI * ORC_EPS0 * E);
    return M_PI * Lambda * Lambda * log(lambda_D / Lambda);
}

static double inter_sigma_v(const orc_model* m, const orc_inter* I, double v)
{
    if (I->type == ORC_COULOMB) return coulomb_sigma(m, I, inter_EeV(m, I, v)) * v;
    if (I->n > 0) return orc_table_lookup(I->n, I->E, I->sigma, inter_EeV(m, I, v)) * v;
    return I->rate;
}

double orc_sigma_v(const orc_model* m, int primary, int target, int k, double v_rel)
{
    return inter_sigma_v(m, m->s[primary].inter[target][k], v_rel);
}

int orc_model_n_interactions(const orc_model* m, int primary, int target) { return m->s[primary].n_inter[target]; }

/* BaseSpecies::svmax_find, src/particles.cpp:190-206 (v advanced by repeated addition, as there) */
static double svmax_find(const orc_model* m, const orc_spec* s, int target, double vmax, int samples)
{
    double dv = vmax / samples;
    double svmax = 0.0;
    for (double v = 0; v < vmax; v += dv)
    {
        double sv = 0;
        for (int i = 0; i < s->n_inter[target]; i++) sv += inter_sigma_v(m, s->inter[target][i], v);
        if (isnan(sv)) continue;
        if (sv > svmax) svmax = sv;
    }
    return svmax;
}

void orc_model_lifetime_init(orc_model* m)
{
    for (int i = 0; i < m->ns; i++)
    {
        orc_spec* s = &m->s[i];
        double rate = 0;
        double vmax = sqrt(s->E_max * ORC_QE / s->mass * 2.0); /* veV(E_max), particles.hpp:129 */
        for (int k = 0; k < m->ns; k++)
        {
            s->rates[k] = svmax_find(m, s, k, vmax, 1000) * m->s[k].density;
            rate += s->rates[k];
        }
        s->lifetime = rate > 0.0 ? 1.0 / rate : INFINITY;
    }
}

double orc_model_lifetime(const orc_model* m, int i) { return m->s[i].lifetime; }
double orc_model_get(const orc_model* m, int i, int what)
{
    const orc_spec* s = &m->s[i];
    switch (what)
    {
        case 0: return s->mass;
        case 1: return s->charge;
        case 2: return s->density;
        case 3: return s->temperature;
        case 4: return s->E_max;
        case 5: return s->dt;
        case 6: return s->v_max;
        default: return s->lifetime;
    }
}
int orc_model_rates(const orc_model* m, int i, double* rates)
{
    for (int k = 0; k < m->ns; k++) rates[k] = m->s[i].rates[k];
    return m->ns;
}

int orc_scatter(const orc_model* m, int primary, orc_rng* rng, double* pvx, double* pvy, double* pvz,
                int* target_out)
{
    const orc_spec* s = &m->s[primary];
    const double mass = s->mass;
    /* target species by cumulative maximal rate; the last species is the fall-through */
    double gamma = orc_rng_uni(rng) / s->lifetime;
    double tmp = 0.0;
    int specid;
    for (specid = 0; specid < m->ns - 1; specid++)
    {
        tmp += s->rates[specid];
        if (tmp > gamma) break;
    }
    if (target_out) *target_out = specid;
    const orc_spec* t = &m->s[specid];
    /* partner velocity: Maxwellian sample for a continuum species, a random live particle otherwise */
    double vr2, vz2, vt2;
    if (t->pool_n == 0)
    {
        vr2 = orc_rng_rnor(rng) * t->v_max * M_SQRT1_2;
        vz2 = orc_rng_rnor(rng) * t->v_max * M_SQRT1_2;
        vt2 = orc_rng_rnor(rng) * t->v_max * M_SQRT1_2;
    }
    else
    {
        int i = (int)(orc_rng_iuni(rng) % (uint32_t)t->pool_n);
        while (!t->pool_alive[i]) i = (int)(orc_rng_iuni(rng) % (uint32_t)t->pool_n);
        vr2 = t->pool_vx[i];
        vz2 = t->pool_vz[i];
        vt2 = t->pool_vy[i];
    }
    double v_rel = norm3(*pvx - vr2, *pvz - vz2, *pvy - vt2);
    double m2 = t->mass;
    /* process by cumulative n*sigma(E_rel)*v_rel against the maximal rate: the remainder is null */
    gamma = orc_rng_uni(rng) * s->rates[specid];
    tmp = 0.0;
    int intid;
    for (intid = 0; intid < s->n_inter[specid]; intid++)
    {
        tmp += inter_sigma_v(m, s->inter[specid][intid], v_rel) * t->density;
        if (tmp > gamma) break;
    }
    if (intid == s->n_inter[specid]) return -1;
    const orc_inter* I = s->inter[specid][intid];
    tmp = 1.0 / (mass + m2);
    switch (I->type)
    {
        case ORC_SUPERELASTIC:
        {
            double E = inter_E(m, I, v_rel) + I->DE;
            double v_rel2 = inter_vrel(m, I, E);
            double rx, rz, ry;
            orc_rng_rot(rng, v_rel2, &rx, &rz, &ry);
            *pvx = (rx * m2 + *pvx * mass + vr2 * m2) * tmp;
            *pvz = (ry * m2 + *pvz * mass + vz2 * m2) * tmp;
            *pvy = (rz * m2 + *pvy * mass + vt2 * m2) * tmp;
            break;
        }
        case ORC_COULOMB:
        case ORC_ELASTIC:
        {
            double v_cm_x = (*pvx * mass + vr2 * m2) * tmp;
            double v_cm_z = (*pvz * mass + vz2 * m2) * tmp;
            double v_cm_y = (*pvy * mass + vt2 * m2) * tmp;
            double rx, ry, rz;
            orc_rng_rot(rng, v_rel, &rx, &ry, &rz);
            *pvx = rx * m2 * tmp + v_cm_x;
            *pvz = rz * m2 * tmp + v_cm_z;
            *pvy = ry * m2 * tmp + v_cm_y;
            /* COULOMB also rewrites the partner (particles.cpp:309-314); pools are read-only here */
            break;
        }
        case ORC_CX:
            *pvx = vr2;
            *pvz = vz2;
            *pvy = vt2;
            break;
        case ORC_LANGEVIN:
        {
            double cx = (*pvx - vr2) * m2 * tmp;
            double cy = (*pvz - vz2) * m2 * tmp;
            double cz = (*pvy - vt2) * m2 * tmp;
            double beta = orc_rng_radius(rng) * I->cutoff;
            if (beta > 1.0)
            {
                double chi = orc_langevin_chi(beta);
                orc_rng_deflect(rng, chi, &cx, &cy, &cz);
            }
            else
                orc_rng_rot_inplace(rng, &cx, &cy, &cz);
            *pvx = cx + (*pvx * mass + vr2 * m2) * tmp;
            *pvz = cy + (*pvz * mass + vz2 * m2) * tmp;
            *pvy = cz + (*pvy * mass + vt2 * m2) * tmp;
            break;
        }
        default: break;
    }
    return intid;
}

/* ------------------------------------------------------- whole-array movers */

static void count_coll(int64_t* counts, int target, int intid, int ns)
{
    if (!counts) return;
    /* counts[target*16 + process], null collisions at counts[ns*16 + target] */
    if (intid < 0) counts[ns * 16 + target]++;
    else if (intid < 16) counts[target * 16 + intid]++;
}

/* util.cpp:22-28 with the default eps = 1e-2 used by fields.cpp:911,927,948-949 */
static int double2int_ref(double x, int* bad)
{
    int res = (int)(x + 0.5);
    if (fabs(res - x) > 1e-2) *bad = 1;
    return res;
}

/* fields.cpp:898-959 (the file has been read into the four vectors, :882-896) */
int orc_btable_build(int n, const double* rvec, const double* zvec, const double* brvec, const double* bzvec, orc_btable* out)
{
    int bad = 0;
    memset(out, 0, sizeof(*out));
    if (n < 2) return 1;
    double dx = 0, rmin = rvec[0], rmax = rvec[n - 1];
    int i = 1;
    while (i < n && rvec[i] - rvec[i - 1] == 0.0) i++;
    if (i >= n) return 1;
    dx = rvec[i] - rvec[i - 1];
    if (dx < 0)
    {
        dx = -dx;
        rmin = rvec[n - 1];
        rmax = rvec[0];
    }
    const int rsampl = double2int_ref((rmax - rmin) / dx + 1, &bad);
    double dz = 0, zmin = zvec[0], zmax = zvec[n - 1];
    i = 1;
    while (i < n && zvec[i] - zvec[i - 1] == 0.0) i++;
    if (i >= n) return 1;
    dz = zvec[i] - zvec[i - 1];
    if (dz < 0)
    {
        dz = -dz;
        zmin = zvec[n - 1];
        zmax = zvec[0];
    }
    const int zsampl = double2int_ref((zmax - zmin) / dz + 1, &bad);
    if (bad) return 3;
    if ((long long)rsampl * zsampl != n) return 1;
    out->jmax = rsampl;
    out->lmax = zsampl;
    out->dx = dx;
    out->dy = dz;
    out->xmin = rmin;
    out->ymin = zmin;
    out->Br = (double*)malloc(sizeof(double) * (size_t)n);
    out->Bz = (double*)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; k++) out->Br[k] = out->Bz[k] = NAN;
    for (int k = 0; k < n; k++)
    {
        const int ri = double2int_ref((rvec[k] - rmin) / dx, &bad);
        const int zi = double2int_ref((zvec[k] - zmin) / dz, &bad);
        if (bad || ri < 0 || ri >= rsampl || zi < 0 || zi >= zsampl)
        {
            orc_btable_free(out);
            return 3;
        }
        out->Br[(size_t)ri * zsampl + zi] = brvec[k];
        out->Bz[(size_t)ri * zsampl + zi] = bzvec[k];
    }
    for (int k = 0; k < n; k++)
        if (isnan(out->Br[k]) || isnan(out->Bz[k]))
        {
            orc_btable_free(out);
            return 2;
        }
    return 0;
}

void orc_btable_free(orc_btable* t)
{
    free(t->Br);
    free(t->Bz);
    t->Br = t->Bz = NULL;
}

void orc_field_B(const orc_grid* g, const orc_btable* t, double x, double y, double* Br, double* Bz, double* Bt)
{
    if (!t)
    {
        *Br = g->Br;
        *Bz = g->Bz;
        *Bt = g->Bt;
        return;
    }
    /* Field2D::resize sets idx = 1.0/dx (Field2D.cpp:11) */
    *Br = orc_interpolate(t->Br, t->jmax, t->lmax, 1.0 / t->dx, 1.0 / t->dy, t->xmin, t->ymin, x, y);
    *Bz = orc_interpolate(t->Bz, t->jmax, t->lmax, 1.0 / t->dx, 1.0 / t->dy, t->xmin, t->ymin, x, y);
    *Bt = 0.00;
}

void orc_advance_boris_B(const orc_grid* g, const orc_btable* t, const double* u, const double* uRF, const orc_model* m, int sp,
                         orc_particles* p, unsigned long niter, orc_rng* rng, int64_t* coll_counts)
{
    const orc_spec* s = &m->s[sp];
    double fx = 0, fz = g->extern_field;
    double Bx = g->Br, Bz = g->Bz, By = g->Bt;
    const double prob = 1.0 - exp(-s->dt / s->lifetime);
    for (int k = 0; k < p->n; k++)
    {
        if (!p->alive[k]) continue;
        orc_field_E(g, u, uRF, p->x[k], p->z[k], niter * s->dt, &fx, &fz);
        orc_field_B(g, t, p->x[k], p->z[k], &Bx, &Bz, &By);
        if (g->coord == ORC_CYLINDRICAL)
            orc_boris_cyl(s->charge, s->mass, s->dt, fx, fz, Bx, Bz, By, &p->x[k], &p->z[k], &p->vx[k], &p->vy[k], &p->vz[k]);
        else
            orc_boris_cart(s->charge, s->mass, s->dt, fx, fz, Bx, Bz, By, &p->x[k], &p->z[k], &p->vx[k], &p->vy[k], &p->vz[k]);
        if (rng && orc_rng_uni(rng) < prob)
        {
            int target;
            int intid = orc_scatter(m, sp, rng, &p->vx[k], &p->vy[k], &p->vz[k], &target);
            count_coll(coll_counts, target, intid, m->ns);
        }
    }
}

void orc_advance_boris(const orc_grid* g, const double* u, const double* uRF, const orc_model* m, int sp,
                       orc_particles* p, unsigned long niter, orc_rng* rng, int64_t* coll_counts)
{
    orc_advance_boris_B(g, NULL, u, uRF, m, sp, p, niter, rng, coll_counts);
}

void orc_advance_boris_init_B(const orc_grid* g, const orc_btable* t, const double* u, const double* uRF, const orc_model* m,
                              int sp, orc_particles* p, unsigned long niter)
{
    const orc_spec* s = &m->s[sp];
    double fx = 0, fz = g->extern_field;
    double Bx = g->Br, Bz = g->Bz, By = g->Bt;
    for (int k = 0; k < p->n; k++)
    {
        if (!p->alive[k]) continue;
        orc_field_E(g, u, uRF, p->x[k], p->z[k], niter * s->dt, &fx, &fz);
        orc_field_B(g, t, p->x[k], p->z[k], &Bx, &Bz, &By);
        if (g->coord == ORC_CYLINDRICAL)
            orc_boris_cyl_init(s->charge, s->mass, s->dt, fx, fz, Bx, Bz, By, &p->vx[k], &p->vy[k], &p->vz[k]);
        else
            orc_boris_cart_init(s->charge, s->mass, s->dt, fx, fz, Bx, Bz, By, &p->vx[k], &p->vy[k], &p->vz[k]);
    }
}

void orc_advance_boris_init(const orc_grid* g, const double* u, const double* uRF, const orc_model* m,
                            int sp, orc_particles* p, unsigned long niter)
{
    orc_advance_boris_init_B(g, NULL, u, uRF, m, sp, p, niter);
}

void orc_advance_boris_extern(const orc_grid* g, const orc_model* m, int sp, orc_particles* p, orc_rng* rng,
                              int64_t* coll_counts, int init)
{
    const orc_spec* s = &m->s[sp];
    const double fx = 0, fz = g->extern_field;
    const double prob = 1.0 - exp(-s->dt / s->lifetime);
    for (int k = 0; k < p->n; k++)
    {
        if (!p->alive[k]) continue;
        if (init)
        {
            orc_boris_cart_init(s->charge, s->mass, s->dt, fx, fz, g->Br, g->Bz, g->Bt, &p->vx[k], &p->vy[k], &p->vz[k]);
            continue;
        }
        orc_boris_cart(s->charge, s->mass, s->dt, fx, fz, g->Br, g->Bz, g->Bt, &p->x[k], &p->z[k], &p->vx[k], &p->vy[k], &p->vz[k]);
        if (rng && orc_rng_uni(rng) < prob)
        {
            int target;
            int intid = orc_scatter(m, sp, rng, &p->vx[k], &p->vy[k], &p->vz[k], &target);
            count_coll(coll_counts, target, intid, m->ns);
        }
    }
}

unsigned orc_source_size(const orc_model* m, int sp, double V, unsigned factor)
{
    double N = m->s[sp].density * V;
    unsigned int n_particles = N / factor;
    return n_particles;
}

void orc_source_refresh(const orc_grid* g, const orc_model* m, int sp, unsigned factor, orc_rng* rng, orc_particles* src)
{
    const orc_spec* s = &m->s[sp];
    double K = 1.0 / factor;
    for (int i = 0; i < src->n; i++)
    {
        src->alive[i] = 1;
        src->x[i] = K * g->x_max * orc_rng_uni(rng);
        src->z[i] = K * g->z_max * orc_rng_uni(rng);
        src->vx[i] = orc_rng_rnor(rng) * s->v_max / (M_SQRT2);
        src->vz[i] = orc_rng_rnor(rng) * s->v_max / (M_SQRT2);
        src->vy[i] = orc_rng_rnor(rng) * s->v_max / (M_SQRT2);
        src->ttd[i] = orc_rng_rexp(rng) * s->lifetime;
    }
    orc_advance_boris_extern(g, m, sp, src, NULL, NULL, 1);
}

/* one "j = insert(); particles[j] = *I; shift; inside ? accumulate : remove(j)" block of source() */
static int source_inject(const orc_grid* g, double charge, const orc_particles* src, int k, double x, double z,
                         orc_particles* dst, int* n_dst, int dst_cap, double* rho, int64_t* rho_fixed)
{
    if (!(z < g->z_max && z > 0 && x < g->x_max && x > 0)) return 0;      /* inserted and removed again */
    if (*n_dst >= dst_cap) return -1;
    const int j = (*n_dst)++;
    dst->x[j] = x;
    dst->z[j] = z;
    if (dst->y) dst->y[j] = src->y ? src->y[k] : 0.0;
    dst->vx[j] = src->vx[k];
    dst->vy[j] = src->vy[k];
    dst->vz[j] = src->vz[k];
    if (dst->ttd) dst->ttd[j] = src->ttd ? src->ttd[k] : 0.0;
    dst->alive[j] = 1;
    unsigned char one = 1;
    if (rho) orc_deposit_fp64(g, charge, 1, &x, &z, &one, rho);
    if (rho_fixed) orc_deposit_fixed(g, 1, &x, &z, &one, rho_fixed);
    return 1;
}

int orc_source(const orc_grid* g, const orc_model* m, int sp, unsigned factor, orc_particles* src, orc_rng* rng,
               int64_t* coll_counts, orc_particles* dst, int* n_dst, int dst_cap, double* rho, int64_t* rho_fixed,
               int (*irand)(void))
{
    const orc_spec* s = &m->s[sp];
    double K = 1.0 / factor;
    double src_z_max = K * g->z_max;
    double src_x_max = K * g->x_max;
    int injected = 0, r;
    orc_advance_boris_extern(g, m, sp, src, rng, coll_counts, 0);
#define INJECT(X, Z)                                                                                  \
    do {                                                                                              \
        r = source_inject(g, s->charge, src, k, (X), (Z), dst, n_dst, dst_cap, rho, rho_fixed);          \
        if (r < 0) return -1;                                                                         \
        injected += r;                                                                                \
    } while (0)
    for (int k = 0; k < src->n; k++)
    {
        /* the reference does not skip empty reservoir slots; the reservoir never has any */
        if (src->x[k] > src_x_max)
            while (src->x[k] > src_x_max)
            {
                src->x[k] -= src_x_max;
                double pz = src->z[k];
                pz += irand() % factor * src_z_max;
                INJECT(src->x[k], pz);
            }
        else if (src->x[k] < 0)
            while (src->x[k] < 0)
            {
                double pz = src->z[k], px = src->x[k];
                pz += irand() % factor * src_z_max;
                px += g->x_max;
                src->x[k] += src_x_max;
                INJECT(px, pz);
            }
        if (src->z[k] > src_z_max)
            while (src->z[k] > src_z_max)
            {
                src->z[k] -= src_z_max;
                double px = src->x[k];
                px += irand() % factor * src_x_max;
                INJECT(px, src->z[k]);
            }
        else if (src->z[k] < 0)
            while (src->z[k] < 0)
            {
                double px = src->x[k], pz = src->z[k];
                px += irand() % factor * src_x_max;
                pz += g->z_max;
                src->z[k] += src_z_max;
                INJECT(px, pz);
            }
    }
#undef INJECT
    return injected;
}

void orc_advance_multicoll(double fx, double fz, const orc_model* m, int sp, orc_particles* p, orc_rng* rng,
                           int64_t* coll_counts)
{
    const orc_spec* s = &m->s[sp];
    const double qm = s->charge / s->mass;
    const double dt = s->dt;
    for (int k = 0; k < p->n; k++)
    {
        if (!p->alive[k]) continue;
        double local_time = 0.;
        double ttd;
        while (local_time + p->ttd[k] < dt)
        {
            ttd = p->ttd[k];
            p->vx[k] += fx * qm * ttd;
            p->vz[k] += fz * qm * ttd;
            /* position uses the already-updated velocity: a reference quirk kept on purpose */
            p->x[k] += (p->vx[k] + 0.5 * fx * qm * ttd) * ttd;
            p->z[k] += (p->vz[k] + 0.5 * fz * qm * ttd) * ttd;
            local_time += p->ttd[k];
            int target;
            int intid = orc_scatter(m, sp, rng, &p->vx[k], &p->vy[k], &p->vz[k], &target);
            count_coll(coll_counts, target, intid, m->ns);
            p->ttd[k] = s->lifetime * orc_rng_rexp(rng);
        }
        ttd = dt - local_time;
        p->vx[k] += fx * qm * ttd;
        p->vz[k] += fz * qm * ttd;
        p->x[k] += (p->vx[k] + 0.5 * fx * qm * ttd) * ttd;
        p->z[k] += (p->vz[k] + 0.5 * fz * qm * ttd) * ttd;
        p->ttd[k] -= ttd;
    }
}

int orc_is_free(const orc_grid* g, const unsigned char* mask, double r, double z)
{
    int i = (int)(r * g->idx);
    int j = (int)(z * g->idz);
    /* the reference reads one node past the grid when r == x_max exactly; clamp instead */
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i > g->M - 2) i = g->M - 2;
    if (j > g->N - 2) j = g->N - 2;
    const unsigned char* q = mask + (size_t)i * g->N + j;
    return q[0] == ORC_FREE || q[g->N] == ORC_FREE || q[1] == ORC_FREE || q[g->N + 1] == ORC_FREE;
}

static inline void cic_weights(const orc_grid* g, double x, double z, int* i, int* j, double w[4])
{
    *i = (int)(x * g->idx);
    *j = (int)(z * g->idz);
    /* x == x_max exactly: the reference indexes one node past the grid (Field2D.hpp:55-60 lets
     * i == jmax-1 through); the build clamps the cell so that the weight goes to the last node */
    if (*i > g->M - 2) *i = g->M - 2;
    if (*j > g->N - 2) *j = g->N - 2;
    if (*i < 0) *i = 0;
    if (*j < 0) *j = 0;
    double u = x * g->idx - *i;
    double v = z * g->idz - *j;
    w[0] = (1 - u) * (1 - v); /* [i][j]     */
    w[1] = u * (1 - v);       /* [i+1][j]   */
    w[2] = (1 - u) * v;       /* [i][j+1]   */
    w[3] = u * v;             /* [i+1][j+1] */
}

int orc_deposit_fp64(const orc_grid* g, double charge, int n, const double* x, const double* z,
                     const unsigned char* alive, double* rho)
{
    int bad = 0;
    for (int k = 0; k < n; k++)
    {
        if (alive && !alive[k]) continue;
        int i, j;
        double w[4];
        cic_weights(g, x[k], z[k], &i, &j, w);
        if (i < 0 || i > g->M - 2 || j < 0 || j > g->N - 2) { bad++; continue; }
        double* q = rho + (size_t)i * g->N + j;
        q[0] += w[0] * charge;
        q[g->N] += w[1] * charge;
        q[1] += w[2] * charge;
        q[g->N + 1] += w[3] * charge;
    }
    return bad;
}

int orc_deposit_fixed(const orc_grid* g, int n, const double* x, const double* z,
                      const unsigned char* alive, int64_t* rho)
{
    int bad = 0;
    for (int k = 0; k < n; k++)
    {
        if (alive && !alive[k]) continue;
        int i, j;
        double w[4];
        cic_weights(g, x[k], z[k], &i, &j, w);
        if (i < 0 || i > g->M - 2 || j < 0 || j > g->N - 2) { bad++; continue; }
        int64_t* q = rho + (size_t)i * g->N + j;
        q[0] += llrint(w[0] * 4294967296.0);
        q[g->N] += llrint(w[1] * 4294967296.0);
        q[1] += llrint(w[2] * 4294967296.0);
        q[g->N + 1] += llrint(w[3] * 4294967296.0);
    }
    return bad;
}

int orc_advance_boundary(const orc_grid* g, const unsigned char* mask, double charge, orc_particles* p,
                         double* rho, int64_t* rho_fixed)
{
    int removed = 0;
    for (int k = 0; k < p->n; k++)
    {
        if (!p->alive[k]) continue;
        if (p->x[k] < g->x_min || p->x[k] > g->x_max || p->z[k] < g->z_min || p->z[k] > g->z_max)
        {
            if (g->boundary == ORC_BC_FREE)
            {
                p->alive[k] = 0;
                removed++;
                continue;
            }
            else if (g->boundary == ORC_BC_PERIODIC)
            {
                p->x[k] = fmod(p->x[k], g->x_max);
                if (p->x[k] < 0) p->x[k] += g->x_max;
                p->z[k] = fmod(p->z[k], g->z_max);
                if (p->z[k] < 0) p->z[k] += g->z_max;
            }
        }
        if (!g->field_from_file)
        {
            if (!orc_is_free(g, mask, p->x[k], p->z[k]))
            {
                p->alive[k] = 0;
                removed++;
                continue;
            }
        }
        if (g->selfconsistent)
        {
            if (rho) orc_deposit_fp64(g, charge, 1, &p->x[k], &p->z[k], 0, rho);
            if (rho_fixed) orc_deposit_fixed(g, 1, &p->x[k], &p->z[k], 0, rho_fixed);
        }
    }
    return removed;
}

/* ----------------------------------------------------------------- Poisson */

void orc_rhs(const orc_grid* g, const unsigned char* mask, const double* voltage, int rf, double* rho)
{
    const double dx = g->dx, dz = g->dz;
    for (int i = 0; i < g->M; i++)
        for (int j = 0; j < g->N; j++)
        {
            size_t k = (size_t)i * g->N + j;
            if (mask[k] == ORC_FIXED) rho[k] = rf ? 0 : voltage[k];
            else if (mask[k] == ORC_FIXED_RF) rho[k] = rf ? voltage[k] : 0;
            else if (g->coord == ORC_CYLINDRICAL)
            {
                if (i > 0) rho[k] *= -1.0 / ORC_EPS0 / (M_PI * dx * dx * 2.0 * i * dz) * g->macroparticle_factor;
                else rho[k] *= -1.0 / ORC_EPS0 / (M_PI * dx * dx * 0.25 * dz) * g->macroparticle_factor;
            }
            else
                rho[k] *= -sqr(dx) / ORC_EPS0 / g->dV;
        }
}

/* stencil row of node (i,j): returns the number of entries written to cols/vals */
static int op_row(const orc_grid* g, const unsigned char* mask, int i, int j, long cols[5], double vals[5])
{
    const int N = g->N;
    const long k = (long)j + (long)N * i;
    const double dx = g->dx, dz = g->dz;
    if (mask[k] == ORC_FIXED || mask[k] == ORC_FIXED_RF)
    {
        cols[0] = k;
        vals[0] = 1;
        return 1;
    }
    if (g->coord == ORC_CYLINDRICAL)
    {
        if (i == 0 && j > 0 && j < N - 1)
        {
            double k2 = 1.0 / (dz * dz);
            double k3 = 1.0 / (dx * dx * 0.25);
            cols[0] = k - 1; vals[0] = k2;
            cols[1] = k; vals[1] = -2.0 * k2 - k3;
            cols[2] = k + 1; vals[2] = k2;
            cols[3] = k + N; vals[3] = k3;
            return 4;
        }
        double k1 = (i - 0.5) / (dx * dx * i);
        double k2 = 1.0 / (dz * dz);
        double k3 = (i + 0.5) / (dx * dx * i);
        cols[0] = k - N; vals[0] = k1;
        cols[1] = k - 1; vals[1] = k2;
        cols[2] = k; vals[2] = -2.0 * k2 - k1 - k3;
        cols[3] = k + 1; vals[3] = k2;
        cols[4] = k + N; vals[4] = k3;
        return 5;
    }
    cols[0] = k - N; vals[0] = 1.0;
    cols[1] = k - 1; vals[1] = 1.0;
    cols[2] = k; vals[2] = -4;
    cols[3] = k + 1; vals[3] = 1.0;
    cols[4] = k + N; vals[4] = 1.0;
    return 5;
}

void orc_apply_operator(const orc_grid* g, const unsigned char* mask, const double* u, double* y)
{
    const long n = (long)g->M * g->N;
    for (int i = 0; i < g->M; i++)
        for (int j = 0; j < g->N; j++)
        {
            long cols[5];
            double vals[5];
            int c = op_row(g, mask, i, j, cols, vals);
            double s = 0;
            for (int q = 0; q < c; q++)
                if (cols[q] >= 0 && cols[q] < n) s += vals[q] * u[cols[q]];
            y[(long)i * g->N + j] = s;
        }
}

int orc_solve_direct(const orc_grid* g, const unsigned char* mask, const double* b, double* u)
{
    const long n = (long)g->M * g->N;
    const long bw = g->N;
    const long w = 2 * bw + 1;
    double* lu = (double*)calloc((size_t)(n * w), sizeof(double));
    if (!lu) return 1;
    for (int i = 0; i < g->M; i++)
        for (int j = 0; j < g->N; j++)
        {
            long cols[5];
            double vals[5];
            long r = (long)i * g->N + j;
            int c = op_row(g, mask, i, j, cols, vals);
            for (int q = 0; q < c; q++)
                if (cols[q] >= 0 && cols[q] < n) lu[r * w + (cols[q] - r + bw)] += vals[q];
        }
    for (long k = 0; k < n; k++)
    {
        const double piv = lu[k * w + bw];
        const long rmax = k + bw < n - 1 ? k + bw : n - 1;
        for (long r = k + 1; r <= rmax; r++)
        {
            double* lrk = &lu[r * w + (k - r + bw)];
            if (*lrk == 0.0) continue;
            *lrk /= piv;
            const double f = *lrk;
            double* rowr = &lu[r * w + bw - r];
            const double* rowk = &lu[k * w + bw - k];
            for (long c = k + 1; c <= rmax; c++) rowr[c] -= f * rowk[c];
        }
    }
    for (long r = 0; r < n; r++) u[r] = b[r];
    for (long r = 0; r < n; r++)
    {
        long c0 = r - bw > 0 ? r - bw : 0;
        const double* rowr = &lu[r * w + bw - r];
        double s = u[r];
        for (long c = c0; c < r; c++) s -= rowr[c] * u[c];
        u[r] = s;
    }
    for (long r = n - 1; r >= 0; r--)
    {
        long c1 = r + bw < n - 1 ? r + bw : n - 1;
        const double* rowr = &lu[r * w + bw - r];
        double s = u[r];
        for (long c = r + 1; c <= c1; c++) s -= rowr[c] * u[c];
        u[r] = s / rowr[r];
    }
    free(lu);
    return 0;
}

void orc_u_smooth(const orc_grid* g, int symmetry, double radius, double* u)
{
    const int M = g->M, N = g->N;
#define U(i, j) u[(size_t)(i) * N + (j)]
    int ic = M - 1, jc = N - 1;
    if (symmetry && ic == jc)
    {
        for (int i = 0; i <= ic / 2; i++)
            for (int j = 0; j <= i; j++)
            {
                double sum = 0;
                sum += U(i, j);
                sum += U(ic - i, j);
                sum += U(i, jc - j);
                sum += U(ic - i, jc - j);
                sum += U(jc - j, ic - i);
                sum += U(j, i);
                sum += U(jc - j, i);
                sum += U(j, ic - i);
                sum /= 8.0;
                U(i, j) = sum;
                U(ic - i, j) = sum;
                U(i, jc - j) = sum;
                U(ic - i, jc - j) = sum;
                U(jc - j, ic - i) = sum;
                U(j, i) = sum;
                U(jc - j, i) = sum;
                U(j, ic - i) = sum;
            }
    }
    if (radius > 0) radius = sqr(radius / g->dx);
    double* t = (double*)malloc(sizeof(double) * (size_t)M * N);
    memcpy(t, u, sizeof(double) * (size_t)M * N);
#define T(i, j) t[(size_t)(i) * N + (j)]
    double icf = ic / 2.0, jcf = jc / 2.0;
    /* NB the reference's inner bound is j <= lmax-1 (fields.cpp:88): for the last column it reads one
     * element past the row end, i.e. the first element of the next row of the contiguous array, and one
     * past the array end for the very last node.  Reproduced with flat indexing; 0 past the end. */
    const long nn = (long)M * N;
#define TF(i, j) (((long)(i) * N + (j)) < nn ? t[(long)(i) * N + (j)] : 0.0)
    for (int i = 1; i < M - 1; i++)
        for (int j = 1; j <= N - 1; j++)
        {
            double r = sqr(i - icf) + sqr(j - jcf);
            if (radius > 0 && r > radius) continue;
            double sum = TF(i, j) + TF(i - 1, j) * 0.5 + TF(i + 1, j) * 0.5 + TF(i, j - 1) * 0.5 + TF(i, j + 1) * 0.5 +
                         TF(i - 1, j - 1) * 0.25 + TF(i + 1, j - 1) * 0.25 + TF(i - 1, j + 1) * 0.25 +
                         TF(i + 1, j + 1) * 0.25;
            U(i, j) = sum / 4.0;
        }
#undef TF
#undef T
#undef U
    free(t);
}

/* ---------------------------------------------------------------- geometry */

static void circle_electrode(const orc_grid* g, unsigned char* mask, double* voltage, double rc, double zc,
                             double radius, double v, unsigned char type)
{
    double sq = sqr(radius);
    for (int i = 0; i < g->M; i++)
        for (int j = 0; j < g->N; j++)
        {
            double r = i * g->dx - rc;
            double z = j * g->dz - zc;
            if (sqr(r) + sqr(z) <= sq)
            {
                mask[(size_t)i * g->N + j] = type;
                voltage[(size_t)i * g->N + j] = v;
            }
        }
}

static void square_electrode(const orc_grid* g, unsigned char* mask, double* voltage, double rmin, double rmax,
                             double zmin, double zmax, double v, unsigned char type)
{
    for (int i = 0; i < g->M; i++)
        for (int j = 0; j < g->N; j++)
        {
            double r = i * g->dx;
            double z = j * g->dz;
            if (r > rmin && r < rmax && z > zmin && z < zmax)
            {
                mask[(size_t)i * g->N + j] = type;
                voltage[(size_t)i * g->N + j] = v;
            }
        }
}

static void mark_boundary(const orc_grid* g, unsigned char* mask)
{
    const int N = g->N;
    for (int i = 2; i < g->M - 2; i++)
        for (int j = 2; j < N - 2; j++)
        {
            size_t k = (size_t)i * N + j;
            if ((mask[k - N] == ORC_FIXED || mask[k + N] == ORC_FIXED || mask[k - 1] == ORC_FIXED ||
                 mask[k + 1] == ORC_FIXED) &&
                mask[k] != ORC_FIXED)
                mask[k] = ORC_BOUNDARY;
        }
}

static void multipole(const orc_grid* g, unsigned char* mask, double* voltage, int npoles, double r_ring,
                      double r_rod)
{
    const double xc = 1e-2, yc = 1e-2;
    for (int i = 0; i < npoles; i++)
    {
        double x = xc + sin(2 * M_PI * (i + 1.0 / 32) / npoles) * r_ring;
        double y = yc + cos(2 * M_PI * (i + 1.0 / 32) / npoles) * r_ring;
        int sign = i % 2 == 0 ? -1 : 1;
        circle_electrode(g, mask, voltage, x, y, r_rod, sign, ORC_FIXED_RF);
    }
}

void orc_geometry(const orc_grid* g, int geometry, double probe_radius, double u_probe, unsigned char* mask,
                  double* voltage)
{
    const int M = g->M, N = g->N;
    /* frame: which edges are Dirichlet and what they carry (fields.cpp:423-434 and siblings) */
    const int axis_open = (geometry == ORC_GEO_MAC || geometry == ORC_GEO_PENNING || geometry == ORC_GEO_PENNING_SIMPLE);
    const int ramp = (geometry == ORC_GEO_EMPTY || geometry == ORC_GEO_PROBE || geometry == ORC_GEO_RF_8PT ||
                      geometry == ORC_GEO_RF_HAITRAP);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++)
        {
            size_t k = (size_t)i * N + j;
            voltage[k] = 0.0;
            int edge = (i == M - 1 || j == 0 || j == N - 1 || (!axis_open && i == 0));
            if (edge)
            {
                mask[k] = ORC_FIXED;
                voltage[k] = ramp ? -g->extern_field * g->dz * (j - N / 2) : 0.0;
            }
            else
                mask[k] = ORC_FREE;
        }
    switch (geometry)
    {
        case ORC_GEO_PROBE:
            circle_electrode(g, mask, voltage, (M - 1) * g->dx / 2, (N - 1) * g->dz / 2, probe_radius, u_probe, ORC_FIXED);
            mark_boundary(g, mask);
            break;
        case ORC_GEO_RF_22PT:
            multipole(g, mask, voltage, 22, 0.75e-2, 0.05e-2);
            mark_boundary(g, mask);
            break;
        case ORC_GEO_RF_8PT:
        {
            double r_rod = 0.1e-2;
            multipole(g, mask, voltage, 8, 0.3e-2 + r_rod, r_rod);
            mark_boundary(g, mask);
            break;
        }
        case ORC_GEO_RF_HAITRAP:
        {
            double r_rod = 0.01e-2;
            multipole(g, mask, voltage, 8, 0.3e-2 + r_rod, r_rod);
            mark_boundary(g, mask);
            break;
        }
        case ORC_GEO_RF_QUAD:
            circle_electrode(g, mask, voltage, 5e-3, 1e-2, 2e-3, 1.0, ORC_FIXED_RF);
            circle_electrode(g, mask, voltage, 15e-3, 1e-2, 2e-3, 1.0, ORC_FIXED_RF);
            circle_electrode(g, mask, voltage, 1e-2, 5e-3, 2e-3, -1.0, ORC_FIXED_RF);
            circle_electrode(g, mask, voltage, 1e-2, 15e-3, 2e-3, -1.0, ORC_FIXED_RF);
            mark_boundary(g, mask);
            break;
        case ORC_GEO_TUBE:
        {
            double c = g->x_max / 2.0;
            double sq = sqr(probe_radius);
            for (int i = 0; i < M; i++)
                for (int j = 0; j < N; j++)
                {
                    double r = i * g->dx - c, z = j * g->dz - c;
                    if (sqr(r) + sqr(z) >= sq)
                    {
                        mask[(size_t)i * N + j] = ORC_FIXED;
                        voltage[(size_t)i * N + j] = 0.0;
                    }
                }
            mark_boundary(g, mask);
            break;
        }
        case ORC_GEO_MAC:
        {
            square_electrode(g, mask, voltage, 5e-3, 4.5e-2, 1e-2, 1.5e-2, -.00, ORC_FIXED);
            square_electrode(g, mask, voltage, 5e-3, 7e-3, 2e-2, 8e-2, 0.0, ORC_FIXED);
            square_electrode(g, mask, voltage, 5e-3, 4.5e-2, 8.5e-2, 9e-2, -.00, ORC_FIXED);
            double th = u_probe, ofs = 3e-2;
            square_electrode(g, mask, voltage, 3e-2, 3.3e-2, 11e-2 + ofs, 14e-2 + ofs, 0.8 * th, ORC_FIXED);
            square_electrode(g, mask, voltage, 4.5e-2, 4.8e-2, 15e-2 + ofs, 25e-2 + ofs, th, ORC_FIXED);
            square_electrode(g, mask, voltage, 3e-2, 3.3e-2, 26e-2 + ofs, 29e-2 + ofs, 1.0 * th, ORC_FIXED);
            square_electrode(g, mask, voltage, 2.5e-2, 2.8e-2, 29e-2 + ofs, 30.5e-2 + ofs, 1.0 * th, ORC_FIXED);
            square_electrode(g, mask, voltage, 15e-3, 4.5e-2, 35e-2, 35.3e-2, .0, ORC_FIXED);
            square_electrode(g, mask, voltage, 0.0, 4.5e-2, 39.5e-2, 40e-2, 3e3, ORC_FIXED);
            mark_boundary(g, mask);
            break;
        }
        case ORC_GEO_PENNING:
            square_electrode(g, mask, voltage, 1.57e-2 / 2, 1.67e-2 / 2, 1e-3, 25e-3, -5, ORC_FIXED);
            square_electrode(g, mask, voltage, 0, 1.46e-2 / 2, 15e-3, 16e-3, 10, ORC_FIXED);
            square_electrode(g, mask, voltage, 0, 7e-3 / 2, 12e-3, 19e-3, 10, ORC_FIXED);
            square_electrode(g, mask, voltage, 0, 1.9e-3, 1e-3, 12e-3, 10, ORC_FIXED);
            square_electrode(g, mask, voltage, 4e-3, 7e-3, 52e-3, 53e-3, -5, ORC_FIXED);
            square_electrode(g, mask, voltage, 2.5e-3, 7e-3, 46e-3, 47e-3, 0, ORC_FIXED);
            mark_boundary(g, mask);
            break;
        case ORC_GEO_PENNING_SIMPLE:
        {
            double ri = 1e-2, ro = 1.1e-2;
            square_electrode(g, mask, voltage, ri, ro, 0, 1e-2, -0.5, ORC_FIXED);
            square_electrode(g, mask, voltage, ri, ro, 1.1e-2, 2e-2, 0, ORC_FIXED);
            square_electrode(g, mask, voltage, ri, ro, 2.1e-2, 6e-2, -1.0, ORC_FIXED);
            square_electrode(g, mask, voltage, ri, ro, 6.1e-2, 7.5e-2, -10, ORC_FIXED);
            mark_boundary(g, mask);
            break;
        }
        case ORC_GEO_EMPTY:
        default: break;
    }
}
