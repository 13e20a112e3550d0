/* oracle/mag2d_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the particle-in-cell / Monte-Carlo-collision hot path of
 * rouckas/mag2d.  It exists to check the CUDA implementation in mag2d_b200/ and is itself pinned
 * against the unmodified reference compiled into oracle/_ref (tests/test_oracle_vs_reference.py,
 * fixtures under tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product never does.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Conventions follow the reference: a 2-D particle has in-plane position (x, z), in-plane velocity
 * (vx, vz) and out-of-plane / azimuthal velocity vy; grids are row-major data[i*N + j] with i along
 * x (r) and j along z (src/Array.hpp:17-22).
 */
#ifndef MAG2D_ORACLE_H
#define MAG2D_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* src/param.cpp:8-10 (old CODATA values, kept for parity) */
#define ORC_EPS0 8.854187817e-12
#define ORC_KB 1.380662e-23
#define ORC_QE 1.602189e-19

/* src/fields.hpp:20 */
enum { ORC_FIXED = 0, ORC_FIXED_RF = 1, ORC_FREE = 2, ORC_BOUNDARY = 3 };
/* src/parser.hpp:14-15 */
enum { ORC_NEUTRAL = 0, ORC_ELECTRON = 1, ORC_ION = 2 };
enum { ORC_ELASTIC = 0, ORC_LANGEVIN = 1, ORC_CX = 2, ORC_COULOMB = 3, ORC_SUPERELASTIC = 4 };
/* src/param.hpp:7,20-23 */
enum { ORC_CARTESIAN = 0, ORC_CYLINDRICAL = 1, ORC_CARTESIAN3D = 2 };
enum { ORC_BC_FREE = 0, ORC_BC_PERIODIC = 1 };

/* ---- grid + field parameters (the subset of Param the hot path reads) ---- */
typedef struct
{
    int M, N;                 /* x_sampl, z_sampl */
    double dx, dz, idx, idz;  /* param.cpp:126-130 */
    double x_min, x_max, z_min, z_max;
    int coord;                /* ORC_CARTESIAN / ORC_CYLINDRICAL */
    int boundary;             /* ORC_BC_FREE / ORC_BC_PERIODIC */
    int selfconsistent, rf, geometry_empty, field_from_file;
    double extern_field;
    double rf_amplitude, rf_U0, rf_omega;
    double Br, Bz, Bt;        /* constant B (fields.hpp:154-171) */
    double dV, macroparticle_factor; /* RHS scaling (fields.cpp:292-307) */
} orc_grid;

/* ---- field gather ---- */
/* Field2D::grad, src/Field2D.hpp:80-168 */
void orc_grad(const double* data, int jmax, int lmax, double idx, double idy, double xmin, double ymin,
              double x, double y, double* grad_x, double* grad_y);
/* Field2D::interpolate, src/Field2D.hpp:64-78 (returns NaN where the reference throws) */
double orc_interpolate(const double* data, int jmax, int lmax, double idx, double idy, double xmin,
                       double ymin, double x, double y);
/* Fields::E, src/fields.hpp:124-150 */
void orc_field_E(const orc_grid* g, const double* u, const double* uRF, double x, double y, double time,
                 double* Ex, double* Ez);

/* ---- magnetic field table (magnetic_field_const = 0) ---- */
/* Fields::Br / Fields::Bz: two Field2D on their own grid (src/fields.hpp:64) */
typedef struct
{
    int jmax, lmax;           /* rsampl, zsampl */
    double dx, dy, xmin, ymin;
    double *Br, *Bz;          /* [jmax*lmax], row-major like Array2D */
} orc_btable;
/* Fields::load_magnetic_field, src/fields.cpp:870-959, from the n parsed rows (r, z, Br, Bz) of the file.
 * Returns 0, or 1 "wrong size of input vector", 2 "garbage loaded", 3 double2int "is not integer" (src/util.cpp:22-28) */
int orc_btable_build(int n, const double* r, const double* z, const double* br, const double* bz, orc_btable* out);
void orc_btable_free(orc_btable* t);
/* Fields::B, src/fields.hpp:152-177; t == NULL: magnetic_field_const */
void orc_field_B(const orc_grid* g, const orc_btable* t, double x, double y, double* Br, double* Bz, double* Bt);

/* ---- movers (collisions handled separately), src/particles.cpp ---- */
/* one particle, Species<CARTESIAN>::advance_boris body :947-987 */
void orc_boris_cart(double charge, double mass, double dt, double fx, double fz, double Bx, double Bz,
                    double By, double* x, double* z, double* vx, double* vy, double* vz);
/* Species<CARTESIAN>::advance_boris_init body :1020-1048 */
void orc_boris_cart_init(double charge, double mass, double dt, double fx, double fz, double Bx, double Bz,
                         double By, double* vx, double* vy, double* vz);
/* Species<CYLINDRICAL>::advance_boris body :562-614 */
void orc_boris_cyl(double charge, double mass, double dt, double fr, double fz, double Bx, double Bz,
                   double By, double* r, double* z, double* vr, double* vt, double* vz);
/* Species<CYLINDRICAL>::advance_boris_init body :646-674 */
void orc_boris_cyl_init(double charge, double mass, double dt, double fr, double fz, double Bx, double Bz,
                        double By, double* vr, double* vt, double* vz);

/* ---- RNG: Marsaglia SHR3 + Marsaglia-Tsang ziggurat as used by t_random, src/random.cpp:19-336 ---- */
typedef struct
{
    uint32_t jz, jsr;
    int32_t hz;
    uint32_t iz, kn[128], ke[256];
    float wn[128], fn[128], we[256], fe[256];
    uint32_t z, w, jcong;
    float nfix_x, nfix_y; /* function-static floats of nfix(), random.cpp:297 */
} orc_rng;
void orc_rng_init(orc_rng* r, uint32_t seed);  /* initialize_seed :200-208 + initialize_tables :211-243 */
void orc_rng_seed(orc_rng* r, uint32_t seed);  /* initialize_seed only */
uint32_t orc_rng_iuni(orc_rng* r);             /* shr3 :33-36 */
float orc_rng_uni(orc_rng* r);                 /* :42 */
float orc_rng_rnor(orc_rng* r);                /* :49-52 + nfix :294-319 */
float orc_rng_rexp(orc_rng* r);                /* :54-57 + efix :323-336 */
double orc_rng_radius(orc_rng* r);             /* :177 */
void orc_rng_rot(orc_rng* r, double len, double* x, double* y, double* z);     /* :103-115 */
void orc_rng_rot_inplace(orc_rng* r, double* x, double* y, double* z);          /* :117-131 */
void orc_rng_deflect(orc_rng* r, double angle, double* x, double* y, double* z);/* :133-175 */
/* generic: fill out[] with n draws of kind "uni"/"rnor"/"rexp"/"iuni"/"radius" (0..4) */
void orc_rng_draw(orc_rng* r, int kind, int n, double* out);

/* complete elliptic integral of the first kind K(k) by AGM; stands in for std::tr1::comp_ellint_1
 * (libstdc++, called at src/particles.cpp:346) */
double orc_ellint_K(double k);
/* Langevin deflection angle chi(beta), src/particles.cpp:342-347; pinned by tests/test_langevin.cpp */
double orc_langevin_chi(double beta);

/* ---- species / interaction model (Speclist wiring, src/pic.cpp:27-80) ---- */
typedef struct orc_model orc_model;
orc_model* orc_model_new(int n_species);
void orc_model_free(orc_model* m);
/* BaseSpecies ctor, src/particles.hpp:152-183.  E_max <= 0 selects the default of :164 */
int orc_model_set_species(orc_model* m, int i, int type, double mass, double charge, double density,
                          double temperature, double E_max, double dt);
/* Interaction ctor, src/particles.hpp:74-83: DE given in eV, LANGEVIN rate *= cutoff^2.  n = 0: no table */
int orc_model_add_interaction(orc_model* m, int type, double DE_eV, double rate, double cutoff, int primary,
                              int secondary, int n, const double* E_eV, const double* sigma);
/* collision partners are drawn from this pool when it is non-empty (src/particles.cpp:230-238);
 * borrowed pointers, n_slots entries, alive[k] != 0 marks a live particle */
void orc_model_set_pool(orc_model* m, int i, int n_slots, const double* vx, const double* vy,
                        const double* vz, const unsigned char* alive);
/* BaseSpecies::lifetime_init without the time_to_death reseeding, src/particles.cpp:151-161 */
void orc_model_lifetime_init(orc_model* m);
double orc_model_lifetime(const orc_model* m, int i);
double orc_model_get(const orc_model* m, int i, int what); /* 0 mass 1 charge 2 density 3 temperature 4 E_max 5 dt 6 v_max 7 lifetime */
int orc_model_rates(const orc_model* m, int i, double* rates_by_species);
int orc_model_n_interactions(const orc_model* m, int primary, int target);
/* Interaction::sigma_v, src/particles.hpp:61-70 */
double orc_sigma_v(const orc_model* m, int primary, int target, int k, double v_rel);
/* vec_interpolate::operator(), src/tabulate.cpp:124-140 */
double orc_table_lookup(int n, const double* xdata, const double* ydata, double x);
/* BaseSpecies::scatter, src/particles.cpp:208-365.  Returns the index of the chosen process inside
 * interactions_by_species[target] (or -1 for a null collision); *target_out gets the target species */
int orc_scatter(const orc_model* m, int primary, orc_rng* rng, double* vx, double* vy, double* vz,
                int* target_out);

/* ---- whole-array movers with collisions, in reference particle order ---- */
typedef struct
{
    int n;                     /* slots */
    double *x, *y, *z, *vx, *vy, *vz, *ttd;
    unsigned char* alive;
} orc_particles;

/* Species<D>::advance_position for ADVANCE_BORIS (src/particles.cpp:925-995 / 540-621): niter is the
 * species step counter used for the RF phase (time = niter*dt).  rng == NULL disables collisions. */
void orc_advance_boris(const orc_grid* g, const double* u, const double* uRF, const orc_model* m, int sp,
                       orc_particles* p, unsigned long niter, orc_rng* rng, int64_t* coll_counts);
void orc_advance_boris_init(const orc_grid* g, const double* u, const double* uRF, const orc_model* m,
                            int sp, orc_particles* p, unsigned long niter);
/* the same two with field->B(I->x, I->z, Bx, Bz, By) read from a table (src/particles.cpp:563, 648, 948, 1022) */
void orc_advance_boris_B(const orc_grid* g, const orc_btable* t, const double* u, const double* uRF, const orc_model* m, int sp,
                         orc_particles* p, unsigned long niter, orc_rng* rng, int64_t* coll_counts);
void orc_advance_boris_init_B(const orc_grid* g, const orc_btable* t, const double* u, const double* uRF, const orc_model* m,
                              int sp, orc_particles* p, unsigned long niter);
/* Species<CARTESIAN>::advance_multicoll, src/particles.cpp:813-859 (constant field fx,fz) */
void orc_advance_multicoll(double fx, double fz, const orc_model* m, int sp, orc_particles* p, orc_rng* rng,
                           int64_t* coll_counts);
/* Species<D>::advance_boundary, src/particles.hpp:370-411; rho (fp64, may be NULL) and rho_fixed
 * (Q32 fixed point, may be NULL) receive the CIC deposit when g->selfconsistent.  Returns removals. */
int orc_advance_boundary(const orc_grid* g, const unsigned char* mask, double charge, orc_particles* p,
                         double* rho, int64_t* rho_fixed);

/* ---- particle source (use_source = 1), Cartesian only: Species<CYLINDRICAL>::source is never defined ---- */
/* advance_boris(what, extern_fields = true) / advance_boris_init(what, true), src/particles.cpp:925-995, 997-1050:
 * fx = 0, fz = extern_field, B = the constants of config.txt; neither field is looked up (:944-949).  init != 0 selects the
 * half step back (no collisions there) */
void orc_advance_boris_extern(const orc_grid* g, const orc_model* m, int sp, orc_particles* p, orc_rng* rng,
                              int64_t* coll_counts, int init);
/* reservoir size of Species<CARTESIAN>::source5_refresh, src/particles.cpp:1057-1060: (unsigned)(density*V/factor) */
unsigned orc_source_size(const orc_model* m, int sp, double V, unsigned factor);
/* Species<CARTESIAN>::source5_refresh, src/particles.cpp:1053-1080: src->n reservoir particles in the box
 * [0, x_max/factor] x [0, z_max/factor] with Maxwellian velocities, then the half step back */
void orc_source_refresh(const orc_grid* g, const orc_model* m, int sp, unsigned factor, orc_rng* rng, orc_particles* src);
/* Species<CARTESIAN>::source, src/particles.cpp:1158-1226: push the reservoir with the external fields, wrap what left it and
 * inject a copy per crossing at the opposite edge of the main box, shifted by rand() % factor reservoir widths along the
 * other axis.  Copies land in dst from slot *n_dst on (the reference's insert() recycles freed slots instead; the particle
 * SET is the same), dst_cap slots available.  rho (fp64, Field2D::accumulate order) and rho_fixed may be NULL.  irand
 * stands for libc rand().  Returns the number of injected particles, -1 when dst is full. */
int orc_source(const orc_grid* g, const orc_model* m, int sp, unsigned factor, orc_particles* src, orc_rng* rng,
               int64_t* coll_counts, orc_particles* dst, int* n_dst, int dst_cap, double* rho, int64_t* rho_fixed,
               int (*irand)(void));

/* ---- deposition ---- */
/* Field2D::accumulate, src/Field2D.hpp:45-62, sequential fp64 in particle order */
int orc_deposit_fp64(const orc_grid* g, double charge, int n, const double* x, const double* z,
                     const unsigned char* alive, double* rho);
/* the build's fixed-point rule: the four fp64 CIC weights of Field2D.hpp:57-60 (charge factored
 * out), each rounded to nearest-even at 2^-32 and summed as int64 (order independent) */
int orc_deposit_fixed(const orc_grid* g, int n, const double* x, const double* z,
                      const unsigned char* alive, int64_t* rho_fixed);
/* t_grid::is_free, src/fields.hpp:94-101 */
int orc_is_free(const orc_grid* g, const unsigned char* mask, double x, double z);

/* ---- Poisson ---- */
/* RHS of Fields::boundary_solve / boundary_solve_rf, src/fields.cpp:278-310 / 314-346 (rf != 0
 * selects the _rf variant); rho is scaled in place into b */
void orc_rhs(const orc_grid* g, const unsigned char* mask, const double* voltage, int rf, double* rho_inout);
/* y = Op(u), the operator whose rows are built at src/fields.cpp:140-259 */
void orc_apply_operator(const orc_grid* g, const unsigned char* mask, const double* u, double* y);
/* direct banded solve of Op(u) = b (stand-in for umfpack_di_solve, src/fields.cpp:311) */
int orc_solve_direct(const orc_grid* g, const unsigned char* mask, const double* b, double* u);
/* Fields::u_smooth, src/fields.cpp:28-113 */
void orc_u_smooth(const orc_grid* g, int symmetry, double radius, double* u);

/* ---- geometry builders, t_grid, src/fields.cpp:374-868 (geometry codes = Param::Geometry order) ---- */
enum { ORC_GEO_EMPTY = 0, ORC_GEO_PROBE, ORC_GEO_RF_22PT, ORC_GEO_RF_8PT, ORC_GEO_RF_HAITRAP, ORC_GEO_RF_QUAD,
       ORC_GEO_MAC, ORC_GEO_PENNING, ORC_GEO_PENNING_SIMPLE, ORC_GEO_TUBE };
void orc_geometry(const orc_grid* g, int geometry, double probe_radius, double u_probe, unsigned char* mask,
                  double* voltage);

#ifdef __cplusplus
}
#endif
#endif
