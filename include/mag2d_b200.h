/* mag2d_b200.h — C ABI of the B200-native mag2d hot path (libmag2d_b200.so).
 *
 * The reference (rouckas/mag2d) has no FFI: its boundary is the C++ class surface that Pic<D> and
 * the drivers use.  This ABI sits directly underneath that surface; every entry point names the
 * reference interface it replaces (paths relative to the reference checkout).  The C++ host layer in
 * mag2d_b200/csrc/host/ (Param, Fields, Species, Pic, plasma2d, test_MCC) and the ctypes binding in
 * mag2d_b200/api.py are both thin callers of these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer argument is a HOST pointer borrowed for the call
 *     unless its name ends in _dev;
 *   - every function returns 0 on success, non-zero on error (mag2d_last_error() has the message);
 *     no exception crosses the ABI (the reference throws std::runtime_error, e.g. particles.hpp:226);
 *   - one context per GPU; calls on one context must be serialised by the caller; work is enqueued
 *     on the context's stream and is asynchronous unless stated (download / sync / solve block);
 *   - 2-D naming follows the reference: position (x, z), velocity (vx, vz) in plane, vy out of plane
 *     (azimuthal in cylindrical coordinates); grids are row-major a[i*N + j], i along x/r;
 *   - coord = MAG2D_CARTESIAN3D runs the reference's 3-D path (src/species3d.cpp, src/fields3d.cpp, src/Field3D.hpp)
 *     through the same entry points: grids are then a[(i*K + j)*N + k] with i along x (M nodes), j along y (K nodes),
 *     k along z (N nodes), particles carry y, and (Br, Bt, Bz) are the constant (Bx, By, Bz);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MAG2D_B200_H
#define MAG2D_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MAG2D_ABI_VERSION 4
#define MAG2D_MAX_SPECIES 16

typedef struct mag2d_ctx mag2d_ctx;

/* enum values equal the reference's (src/param.hpp:7,20-23; src/parser.hpp:14-15; src/fields.hpp:20) */
enum { MAG2D_CARTESIAN = 0, MAG2D_CYLINDRICAL = 1, MAG2D_CARTESIAN3D = 2 };
enum { MAG2D_BOUNDARY_FREE = 0, MAG2D_BOUNDARY_PERIODIC = 1 };
enum { MAG2D_ADVANCE_BORIS = 0, MAG2D_ADVANCE_MULTICOLL = 1 };
enum { MAG2D_NEUTRAL = 0, MAG2D_ELECTRON = 1, MAG2D_ION = 2 };
enum { MAG2D_ELASTIC = 0, MAG2D_LANGEVIN = 1, MAG2D_CX = 2, MAG2D_COULOMB = 3, MAG2D_SUPERELASTIC = 4 };
enum { MAG2D_FIXED = 0, MAG2D_FIXED_RF = 1, MAG2D_FREE = 2, MAG2D_BOUNDARY = 3 };

/* The subset of the reference's Param (src/param.hpp:24-79) that the hot path reads.  dx, dz, idx,
 * idz, dV are passed as the host computed them (param.cpp:126-136) so both sides round identically. */
typedef struct
{
    int32_t coord, boundary, mover;
    int32_t M, N, K;            /* x_sampl, z_sampl, y_sampl (K only for CARTESIAN3D) */
    double x_max, z_max, y_max; /* domain is [0,x_max] x [0,z_max] (x [0,y_max]) */
    double dx, dz, dy, idx, idz, idy;
    int32_t selfconsistent, rf, geometry_empty, electric_field_from_file;
    double extern_field;
    double rf_amplitude, rf_U0, rf_omega;
    int32_t magnetic_field_const, u_smooth;   /* u_smooth: Fields::u_smooth() after every solve (pic.cpp:336) */
    double Br, Bz, Bt;
    double dV, macroparticle_factor;
} mag2d_grid_desc;

/* SpeciesParams (src/parser.hpp:16-28); E_max <= 0 selects 10 kT/q_e (src/particles.hpp:164) */
typedef struct
{
    int32_t type, reserved0;
    double mass, charge, density, temperature, E_max, dt;
} mag2d_species_desc;

/* InteractionParams (src/parser.hpp:30-43); the cross-section table of interaction k is
 * table_E[table_offset .. table_offset+n_table) / table_sigma[...] (eV, m^2); n_table = 0: constant RATE */
typedef struct
{
    int32_t type, primary, secondary, n_table, table_offset, reserved0;
    double DE_eV, rate, cutoff;
} mag2d_interaction_desc;

/* binary layout of the reference's t_particle (src/particles.hpp:26-33, 64 bytes), used by
 * BaseSpecies::save/load files (src/particles.cpp:32-93) */
typedef struct
{
    double x, y, z, vx, vy, vz, time_to_death;
    uint8_t empty, pad[7];
} mag2d_particle;

/* ---- life cycle ------------------------------------------------------------------------------ */
int mag2d_abi_version(void);
/* thread-local message of the last failing call on this thread */
const char* mag2d_last_error(void);
/* Replaces Pic<D>::Pic / Fields::Fields (src/pic.cpp:127-189, src/fields.cpp:115-276).  `stream` is a
 * cudaStream_t to enqueue on, or NULL for a context-owned non-blocking stream. */
int mag2d_create(int device, const mag2d_grid_desc* grid, void* stream, mag2d_ctx** out);
int mag2d_destroy(mag2d_ctx* ctx);
int mag2d_sync(mag2d_ctx* ctx);
int mag2d_seed(mag2d_ctx* ctx, uint64_t seed); /* t_random::initialize_seed, src/random.cpp:200-208 */

/* ---- geometry and fields ---------------------------------------------------------------------- */
/* t_grid mask / voltage (src/fields.hpp:31-55): mask[M*N] of MAG2D_FIXED.., voltage[M*N].  Builds the
 * multigrid hierarchy that replaces the UMFPACK factorisation of src/fields.cpp:133-275. */
int mag2d_set_grid(mag2d_ctx* ctx, const uint8_t* mask, const double* voltage);
/* which: 0 = u, 1 = uRF (Fields::u / uRF host mirrors, src/fields.hpp:60-62) */
int mag2d_set_potential(mag2d_ctx* ctx, int which, const double* values);
int mag2d_get_potential(mag2d_ctx* ctx, int which, double* values);
/* Fields::boundary_solve (rf = 0) / boundary_solve_rf (rf = 1), src/fields.cpp:278-348: scale rho into
 * the right-hand side, overwrite Dirichlet rows, solve Op(u) = b by multigrid V-cycles until the largest
 * Jacobi update max_k |r_k / a_kk| drops below tol * max_k |u_k| (or max_cycles).  Blocks.  resid_out
 * receives that ratio.  Any out pointer may be NULL. */
int mag2d_solve(mag2d_ctx* ctx, int rf, double tol, int max_cycles, int* cycles_out, double* resid_out);
/* solver knobs for mag2d_step: V-cycles per step (0 = iterate to tol, one host sync per cycle; n > 0 = exactly
 * n cycles, no sync, residual monitored on the device; n < 0 = |n| cycles started from the time-extrapolated
 * guess 2u_n - u_{n-1}) and the tolerance of the iterate-to-tol mode */
int mag2d_set_solver(mag2d_ctx* ctx, int cycles_per_step, double tol, int max_cycles);
/* which Poisson solver stands in for the reference's UMFPACK factorisation (src/fields.cpp:169-329).
 * MAG2D_SOLVER_AUTO: the direct sine-transform x tridiagonal solver when the grid separates (every grid row is an
 * electrode over its whole length or free between two Dirichlet end nodes: geometry EMPTY), multigrid otherwise;
 * MAG2D_SOLVER_MULTIGRID: always multigrid; MAG2D_SOLVER_DIRECT: error unless the grid separates.  The direct
 * solver is exact to round-off, so cycles_per_step / tol do not apply to it (solver_stats reports 0 cycles and
 * the measured residual). */
#define MAG2D_SOLVER_AUTO 0
#define MAG2D_SOLVER_MULTIGRID 1
#define MAG2D_SOLVER_DIRECT 2
int mag2d_set_solver_kind(mag2d_ctx* ctx, int kind);
/* 1 when the direct solver is the one in use for the current grid and kind, else 0 */
int mag2d_solver_is_direct(mag2d_ctx* ctx);
/* V-cycles used and convergence measure reached by the most recent solve (also the one inside mag2d_step
 * when cycles_per_step == 0; with a fixed cycle count the residual is not evaluated and reads 0) */
int mag2d_solver_stats(mag2d_ctx* ctx, int* last_cycles, double* last_resid);
/* Fields::u_smooth, src/fields.cpp:28-113 */
int mag2d_u_smooth(mag2d_ctx* ctx, int symmetry, double radius);
/* ElMag3D::E at n points, src/fields3d.hpp:95-101 (CARTESIAN3D; diagnostics / parity checks) */
int mag2d_field_E3(mag2d_ctx* ctx, int n, const double* x, const double* y, const double* z, double* Ex, double* Ey, double* Ez);
/* Fields::Br / Fields::Bz as filled by Fields::load_magnetic_field (src/fields.cpp:870-959) when
 * magnetic_field_const = 0: two tables on their own r_sampl x z_sampl grid (row-major, r slowest, spacing dr / dz,
 * origin r_min / z_min), interpolated bilinearly at every particle by the Boris movers (Fields::B,
 * src/fields.hpp:172-175; B_theta = 0).  The reference throws "Field2D::interpolate() outside of range" when a
 * particle leaves the table; here the table must cover the simulation box [0, x_max] x [0, z_max] and the call
 * fails with that message otherwise.  Br = Bz = NULL returns to the constant field of the grid descriptor. */
int mag2d_set_magnetic_field(mag2d_ctx* ctx, int r_sampl, int z_sampl, double dr, double dz, double r_min, double z_min,
                             const double* Br, const double* Bz);
/* Fields::B at n points, src/fields.hpp:152-177 (diagnostics / parity checks) */
int mag2d_field_B(mag2d_ctx* ctx, int n, const double* x, const double* z, double* Br, double* Bz, double* Bt);
/* Fields::E at n points, src/fields.hpp:124-150 (diagnostics / parity checks) */
int mag2d_field_E(mag2d_ctx* ctx, int n, const double* x, const double* z, double time, double* Ex, double* Ez);

/* ---- species and collisions ------------------------------------------------------------------- */
/* Speclist<D>::Speclist (src/pic.cpp:27-80) + BaseSpecies::lifetime_init (src/particles.cpp:142-170) */
int mag2d_set_species(mag2d_ctx* ctx, int n_species, const mag2d_species_desc* species, int n_interactions,
                      const mag2d_interaction_desc* interactions, const double* table_E,
                      const double* table_sigma, int n_table_total);
/* what: 0 lifetime, 1 v_max, 2 E_max, 3 t (species clock), 4 niter, 5 prob per step */
int mag2d_species_get(mag2d_ctx* ctx, int species, int what, double* out);
int mag2d_species_rates(mag2d_ctx* ctx, int species, double* rates_by_species /* [n_species] */);
/* per-process collision counters of a primary species since the last reset:
 * counts[target*16 + process], null collisions at counts[n_species*16 + target] */
int mag2d_collision_counts(mag2d_ctx* ctx, int species, int64_t* counts /* [(n_species+1)*16] */, int reset);
/* counting costs one global atomic per collision event; off by default */
int mag2d_set_collision_counting(mag2d_ctx* ctx, int enable);

/* ---- particle store (BaseSpecies::particles / insert / remove, src/particles.hpp:113,223-249) ---- */
int mag2d_reserve(mag2d_ctx* ctx, int species, int64_t capacity);
/* append n particles given as the reference's 64-byte AoS records (empty ones are skipped) */
int mag2d_particles_upload(mag2d_ctx* ctx, int species, const mag2d_particle* aos, int64_t n);
/* append n particles given as SoA host arrays (y and ttd may be NULL) */
int mag2d_particles_upload_soa(mag2d_ctx* ctx, int species, int64_t n, const double* x, const double* y,
                               const double* z, const double* vx, const double* vy, const double* vz,
                               const double* ttd);
/* all slots in device order; removed particles are returned with empty = 1 */
int mag2d_particles_download(mag2d_ctx* ctx, int species, mag2d_particle* aos, int64_t capacity, int64_t* n_slots);
int mag2d_particles_download_soa(mag2d_ctx* ctx, int species, int64_t capacity, double* x, double* y, double* z,
                                 double* vx, double* vy, double* vz, double* ttd, uint8_t* alive,
                                 int64_t* n_slots);
int mag2d_particles_clear(mag2d_ctx* ctx, int species);
int mag2d_count(mag2d_ctx* ctx, int species, int64_t* n_alive, int64_t* n_slots); /* n_particles() */
/* device-side loaders with the context's Philox stream (src/particles.cpp:685-749, 485-510):
 * kind 0 add_particles_everywhere(n), 1 add_particles_on_disk(n, a=cx, b=cz, c=radius),
 * 2 add_monoenergetic_particles_on_cylinder_cylindrical(n, a=energy eV, b=centre z, c=radius, d=height) */
int mag2d_particles_generate(mag2d_ctx* ctx, int species, int kind, int64_t n, double a, double b, double c, double d);
/* ---- particle source (use_source = 1; CARTESIAN + ADVANCE_BORIS, as far as the reference's own code goes) ----- */
/* Param::use_source: mag2d_step then runs Species::source() after every species advance (src/pic.cpp:346-347) and
 * sorts with the stand-alone pass, which also trims the slot range */
int mag2d_set_use_source(mag2d_ctx* ctx, int on);
/* Species<CARTESIAN>::source5_refresh(factor) (src/particles.cpp:1053-1080): (unsigned)(density*V/factor) reservoir
 * particles, uniform in [0, x_max/factor] x [0, z_max/factor], Maxwellian at the species temperature, then the half
 * step back with the external fields.  V is Param::V (n_particles_total / density_total).  A species without particles
 * keeps an empty reservoir (:1063).  Device-side Philox generation.  In a particle-sharded multi-GPU run every rank
 * owns its own reservoir: pass this rank's share V / nranks, so that the ranks together inject the physical flux. */
int mag2d_source_refresh(mag2d_ctx* ctx, int species, uint32_t factor, double V);
/* BaseSpecies::source2_particles (src/particles.hpp:114) as the save/load files carry it (src/particles.cpp:115-138) */
int mag2d_source_upload(mag2d_ctx* ctx, int species, uint32_t factor, const mag2d_particle* aos, int64_t n);
int mag2d_source_download(mag2d_ctx* ctx, int species, mag2d_particle* aos, int64_t capacity, int64_t* n_out);
/* Species<CARTESIAN>::source() (src/particles.cpp:1158-1226): push the reservoir with the external fields (collisions
 * included), wrap it, and for every crossing of a reservoir edge insert a copy at the opposite edge of the simulation box,
 * shifted by a random whole number of reservoir widths along the other axis; copies deposit their charge when the run is
 * self-consistent.  injected (may be NULL) receives the number of particles added.  Blocks (the slot count comes back). */
int mag2d_species_source(mag2d_ctx* ctx, int species, int64_t* injected);

/* cell sort + compaction of removed particles (replaces the free list, src/particles.hpp:223-247) */
int mag2d_sort(mag2d_ctx* ctx, int species);
/* Pic<D>::advance (src/pic.cpp:330-358) for a caller that keeps its particles in HOST memory, as the reference does
 * (BaseSpecies::particles, src/particles.hpp:113): species[q] has n_slots[q] slots in the host SoA arrays x[q], z[q],
 * vx[q], vy[q], vz[q] (pinned memory recommended).  The arrays stream through device staging buffers in chunks of
 * chunk_slots (0 = 4 Mi slots): upload, fused push / MCC / deposit and download of successive chunks overlap, the
 * field solve and the charge all-reduce run as in mag2d_step.  Blocks until the host arrays hold the new state;
 * removed particles come back with x = NaN.  2-D Boris movers. */
int mag2d_step_streamed(mag2d_ctx* ctx, int n_species, const int32_t* species, const int64_t* n_slots, double* const* x,
                        double* const* z, double* const* vx, double* const* vy, double* const* vz, int64_t chunk_slots);
/* bytes the streamed steps have copied host -> device / device -> host so far (reset != 0 clears the counters).  In a 2-D
 * Cartesian run without magnetic field the out-of-plane velocity is not staged when the caller's vy arrays are pinned host memory:
 * only the collision pass touches it, in place over PCIe (MAG2D_STREAM_VY=1 in the environment stages it like the others). */
int mag2d_streamed_bytes(mag2d_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes, int reset);
/* the same for CARTESIAN3D stores (six arrays: Species<CARTESIAN3D>::advance, src/species3d.cpp:3-93, on host-resident particles) */
int mag2d_step_streamed3(mag2d_ctx* ctx, int n_species, const int32_t* species, const int64_t* n_slots, double* const* x,
                         double* const* y, double* const* z, double* const* vx, double* const* vy, double* const* vz,
                         int64_t chunk_slots);
int mag2d_set_sort_interval(mag2d_ctx* ctx, int steps); /* 0 = never sort inside mag2d_step; -1 = per species from its thermal drift (v_th dt K ~ 0.3 cell, 2..64) */
/* per-species override (-1 = use the context-wide interval): slow species (ions) need far fewer sorts than fast
 * ones.  With the Boris movers the sort is carried by the push kernels themselves (a COUNT step takes per-cell
 * counts, the next step draws every particle's sorted slot from them and writes it there). */
int mag2d_set_species_sort_interval(mag2d_ctx* ctx, int species, int steps);
/* Layout of a CARTESIAN3D store inside mag2d_step: MAG2D_LAYOUT_AUTO / MAG2D_LAYOUT_BRICKS bin the particles by 4 x 4 x 4-cell brick (one
 * CTA per brick: field tile and charge tile in shared memory, leavers migrate between bins every step; push3d_brick.cu),
 * MAG2D_LAYOUT_SLOTS keeps the slot-order kernel with the fused COUNT / PERMUTE cell sort.  2-D stores ignore it.  No reference
 * counterpart: the reference's vector<t_particle> has no order (src/particles.hpp:223-247). */
/* Element type of the device-resident particle arrays (north_star: "vectorised coalesced fp64/fp32 loads"; the reference's
 * t_particle is all double, src/particles.hpp:26-33).  MAG2D_STORE_F32 keeps x, z, vx, vy, vz as floats in HBM (40 instead of 80 bytes
 * per 2D3V particle-step); every operation is still carried out in fp64 on the widened values, positions are rounded to float BEFORE the
 * boundary test and the deposit, so the charge grid is the exact fixed-point deposit of the stored positions.  Host-side interfaces
 * (mag2d_particle, the SoA arrays) stay double.  2-D Boris movers; set it before any particle is loaded.  Not available with the
 * multi-collision mover, CARTESIAN3D, the particle source, mag2d_step_streamed and particle-partner collisions. */
#define MAG2D_STORE_F64 0
#define MAG2D_STORE_F32 1
int mag2d_set_storage(mag2d_ctx* ctx, int storage);
#define MAG2D_LAYOUT_AUTO 0
#define MAG2D_LAYOUT_SLOTS 1
#define MAG2D_LAYOUT_BRICKS 2
int mag2d_set_store_layout(mag2d_ctx* ctx, int layout);
/* out8 (species of a CARTESIAN3D context): re-binnings so far; since the last re-binning: guests that found their new bin full,
 * leavers that did not fit the list; leavers listed by the last step; number of bins, index of the fullest bin, its free slots,
 * slots in use over all bins (live particles + holes) */
int mag2d_store_stats(mag2d_ctx* ctx, int species, int64_t* out8);

/* ---- stepping --------------------------------------------------------------------------------- */
/* Pic<D>::advance_init, src/pic.cpp:359-384 */
int mag2d_advance_init(mag2d_ctx* ctx);
/* nsteps x Pic<D>::advance, src/pic.cpp:330-358: [solve] -> per species push+MCC+boundary+deposit ->
 * [rho all-reduce].  Asynchronous.  On N ranks every rank returns holding the complete summed charge grids; between the
 * steps of ONE call a CARTESIAN3D context with the shared-out field solve only sums the planes each rank owns
 * (reduce-scatter), so batching steps into one call is cheaper than calling with nsteps = 1. */
int mag2d_step(mag2d_ctx* ctx, int nsteps);
/* one Species<D>::advance (src/particles.hpp:342-349) of one species: fused advance_position +
 * advance_boundary (+ deposit into that species' rho when selfconsistent) */
int mag2d_species_advance(mag2d_ctx* ctx, int species);
/* Species<D>::advance_init for one species (half step back), src/particles.hpp:351-355 */
int mag2d_species_advance_init(mag2d_ctx* ctx, int species);
/* Species<D>::accumulate, src/particles.hpp:358-367: deposit the current positions */
int mag2d_species_accumulate(mag2d_ctx* ctx, int species);
int mag2d_rho_reset(mag2d_ctx* ctx, int species /* -1 = all */);
/* fixed-point charge grid of one species: sum of Q32 CIC weights (int64 [M*N]) */
int mag2d_rho_fixed_download(mag2d_ctx* ctx, int species, int64_t* rho_fixed);
/* Fields::rho in coulombs: sum over species of charge * weights * 2^-32, fixed species order */
int mag2d_rho_download(mag2d_ctx* ctx, double* rho);
int mag2d_rho_upload(mag2d_ctx* ctx, int species, const int64_t* rho_fixed);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* BaseSpecies::energy_dist_compute (src/particles.cpp:408-414): histogram of kinetic energy in eV,
 * nbins bins on (0, emax); stats = {n_in_range, sum_in_range, n_total, sum_total} (Histogram, histogram.cpp) */
int mag2d_energy_hist(mag2d_ctx* ctx, int species, int nbins, double emax, double* hist, double* stats);

/* ---- multi-GPU (new: the reference is single-process) ------------------------------------------- */
/* fill a 128-byte ncclUniqueId on rank 0; broadcast it out of band; then every rank calls comm_init */
int mag2d_comm_unique_id(void* id128);
int mag2d_comm_init(mag2d_ctx* ctx, int rank, int nranks, const void* id128);
int mag2d_comm_destroy(mag2d_ctx* ctx);

/* ---- measurement ------------------------------------------------------------------------------ */
/* number of kernels this context has launched since creation */
int mag2d_kernel_launches(mag2d_ctx* ctx, int64_t* n);
/* device time in ms of the most recent mag2d_step broken down by phase:
 * out[0] push+MCC+boundary+deposit, out[1] rhs+solve, out[2] sort, out[3] all-reduce, out[4] total.
 * Needs mag2d_set_timing(ctx, 1) before the step; blocks until the step has finished. */
int mag2d_set_timing(mag2d_ctx* ctx, int enable);
int mag2d_timers(mag2d_ctx* ctx, double* out5);
/* raw device pointers for zero-copy interop (torch.distributed, CUDA IPC): what = 0 rho_fixed
 * (all species, int64 [n_species*M*N]), 1 u, 2 uRF */
int mag2d_device_pointer(mag2d_ctx* ctx, int what, void** ptr_dev, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif
