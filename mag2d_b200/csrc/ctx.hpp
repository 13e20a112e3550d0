// ctx.hpp — host-side context behind the opaque mag2d_ctx of include/mag2d_b200.h
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/mag2d_b200.h"
#include "common.cuh"

#define ARR_X 0
#define ARR_Z 1
#define ARR_VX 2
#define ARR_VY 3
#define ARR_VZ 4
#define ARR_Y 5
#define ARR_TTD 6
#define N_ARR 7

struct SpeciesStore
{
    mag2d_species_desc desc{};
    double v_max = 0, E_max = 0, lifetime = INFINITY;
    std::vector<double> rates;        // rates_by_species
    // particle slabs: two generations (the sort writes out of place), N_ARR arrays each
    double* arr[2][N_ARR] = {};
    int cur = 0;
    long long capacity = 0;
    long long n_slots = 0;            // high-water mark of used slots
    unsigned long long* d_removed = nullptr;   // device counter of removals since the last compaction
    unsigned long long* d_counts = nullptr;    // collision counters [(ns+1)*16]
    MccBlob* d_blob = nullptr;
    MccBlob* h_blob = nullptr;
    unsigned long long niter = 0;     // BaseSpecies::niter
    double t = 0;                     // BaseSpecies::t
    int steps_since_sort = 0;
    // cell sort fused into the push (sort.cu): cells counted by a COUNT step, cursors consumed by the next PERMUTE step
    unsigned* d_cell_count = nullptr; // [ncells]
    unsigned* d_cell_offset = nullptr;
    unsigned* d_sort_sums = nullptr;  // scan scratch + grand total
    bool tickets_valid = false;       // a COUNT step has run (cursors are pending) and nothing has disturbed the slots since
    int sort_interval = -1;           // pushes between permuting steps (-1: the context-wide setting)
    int pushes_since_permute = 1 << 20;
    // n_slots is only an upper bound after a fused permute (the live count stays on the device); every permute also
    // sends it to a pinned host word, and a later step adopts it once the copy has landed — no synchronisation
    unsigned long long* h_total = nullptr;   // pinned
    cudaEvent_t ev_total = nullptr;
    bool total_pending = false;
    long long append_epoch = 0, total_epoch = -1;
    // brick-binned layout of a CARTESIAN3D store (push3d_brick.cu): brick b owns the slots [bin_off[b], bin_off[b+1]) of the current
    // slab, the first bin_cnt[b] of them in use.  Anything that moves particles without going through the brick kernel, or appends
    // to the store, drops bins_valid; the next step in brick mode re-bins.
    bool bins_valid = false;
    int bin_nb = 0;
    unsigned* d_bin_off = nullptr;
    unsigned* d_bin_cnt = nullptr;
    unsigned* d_bin_scratch = nullptr;
    double* d_outbox[6] = {};             // this step's leavers (x, y, z, vx, vy, vz), placed into their new bins by k_place3d
    unsigned* d_ob_dst = nullptr;         // their destination bricks
    unsigned* d_ob_count = nullptr;       // [0] outbox entries of this step, [1] arrivals that found a bin full, [2] CTAs whose leavers did not fit (since the last re-binning)
    long long ob_cap = 0;
    int pushes_since_compact = 0;         // in-place steps since the last compacting one
    unsigned* h_bin_flags = nullptr;      // pinned copy of d_mig_count, adopted without a synchronisation
    cudaEvent_t ev_bin = nullptr;
    bool bin_flags_pending = false;
    unsigned bin_overflow_seen = 0;
    double bin_slack = 0.125;             // spare slots behind every bin relative to the local mean fill (on top of 6 sigma + 32)
    long long rebinnings = 0;
    // particle source (use_source): the reservoir BaseSpecies::source2_particles (particles.hpp:114), filled by
    // mag2d_source_refresh / mag2d_source_upload and pushed + sampled by mag2d_species_source.  x, z, vx, vy, vz, ttd
    double* src[6] = {};
    long long src_n = 0, src_capacity = 0;
    unsigned src_factor = 0;              // source5_factor
    unsigned* d_src_count = nullptr;      // particles injected by the current call
    unsigned long long src_calls = 0;     // part of the reservoir's RNG counter
};

// one level of the Galerkin multigrid hierarchy (poisson.cu)
struct MgLevel
{
    int M = 0, N = 0;
    int fx = 1, fz = 1;         // coarsening factor towards the next (coarser) level
    double* coef = nullptr;     // [9][M*N] stencil coefficients, order (di,dj) = (-1,-1),(-1,0),(-1,1),(0,-1),(0,0),(0,1),(1,-1),(1,0),(1,1)
    unsigned char* freem = nullptr;  // 1 = unknown, 0 = Dirichlet / eliminated
    double* u = nullptr;        // solution (level 0: the potential) or error (coarser levels)
    double* b = nullptr;        // right-hand side (scaled)
};

// separable direct solver (poisson_direct.cu): sine transform along z, tridiagonal solves along x / r
struct DirectSolver
{
    bool ok = false;            // the grid separates (every row is an electrode or free between Dirichlet ends)
    int n = 0;                  // N - 2
    int hp = 0;                 // ceil(n / 2) rounded up to a multiple of 32
    int ld = 0;                 // 2 hp: leading dimension of every [.][n] array below; a row is FOLDED: [symmetric half | antisymmetric half]
    double* bp = nullptr;       // [M][ld] folded right-hand side of the interior columns (written by k_rhs); then the inverse products' output
    double* S = nullptr;        // [2][2][hp][hp] half-size sine matrices: forward (parity 0, 1), inverse (parity 0, 1)
    double* fwd = nullptr;      // [M][ld][2] forward sweep pairs (hat/den, lower/den)
    double* inv = nullptr;      // [M][ld] 1/den, applied in the forward product's epilogue
    double* bwd = nullptr;      // [M][ld][2] backward sweep pairs (y, upper/den)
    double* hat = nullptr;      // [M][ld] transformed solution
    unsigned char* rowfree = nullptr;
    double* k2 = nullptr;
};

// direct solver of the 3-D box (poisson3d.cu)
struct Direct3D
{
    bool ok = false;
    int n_i = 0, n_j = 0, n_k = 0;   // interior unknowns along x, y, z
    int ldj = 0, ldk = 0;            // 2 * hpj, 2 * hpk: a folded row holds the symmetric half, then the antisymmetric half
    int hpj = 0, hpk = 0;            // ceil(n / 2) rounded up to a multiple of 32
    int ne = 0;                      // electrode nodes inside the box (capacitance-matrix method)
    // half-size sine matrices of the folded transforms, [2][hp][hp] each (parity 0: odd mode numbers acting on the symmetric
    // half, parity 1: even mode numbers on the antisymmetric half); _f forward, _i inverse (the transposes)
    double* Sy_f = nullptr;
    double* Sy_i = nullptr;
    double* Sz_f = nullptr;
    double* Sz_i = nullptr;
    double* inv = nullptr;           // [n_i][ldj][ldk] Thomas factors
    double* R = nullptr;             // [n_i][ldj][ldk] work arrays
    double* T = nullptr;
    unsigned char* interior_fixed = nullptr;
    int* e_nodes = nullptr;
    double* e_volts = nullptr;
    double* cinv = nullptr;
    double* alpha = nullptr;
    double* green = nullptr;         // [ne][M*K*N] Green's functions of the electrode nodes
    // slab-parallel solve on N ranks (solve_interior_slab): this rank transforms the x planes [pi0[r], pi0[r+1]) and runs the
    // Thomas recurrences of the y rows [pj0[r], pj0[r+1]) of every plane
    bool slab_ok = false;
    std::vector<int> pi0, pj0;       // [nranks + 1]
    double* V = nullptr;             // [n_i][rows of this rank][ldk] transformed right-hand side / solution, mode slab
    double* inv_slab = nullptr;      // the same slab of the Thomas factors
    double* Sb = nullptr;            // send / receive staging, one x slab each
    double* Xb = nullptr;
};

struct mag2d_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    mag2d_grid_desc g{};
    uint64_t seed = 0x9E3779B97F4A7C15ULL;
    long long launches = 0;

    // grid-sized device arrays
    unsigned char* d_mask = nullptr;
    double* d_voltage = nullptr;
    double* d_u = nullptr;      // also mg[0].u
    double* d_uRF = nullptr;
    double* d_ueff = nullptr;   // grid-sized scratch (u_smooth)
    double* d_gx = nullptr;     // edge-centred field differences of the current step (push.cu: k_edge_fields)
    double* d_gz = nullptr;
    double* d_gy = nullptr;     // CARTESIAN3D only
    // magnetic field table (mag2d_set_magnetic_field; Fields::load_magnetic_field, fields.cpp:870-959)
    double* d_btab_r = nullptr;
    double* d_btab_z = nullptr;
    int btab_M = 0, btab_N = 0;
    double btab_idx = 0, btab_idz = 0, btab_xmin = 0, btab_zmin = 0;
    unsigned char* d_cfree = nullptr;   // per-cell "has a FREE corner" flag (t_grid::is_free)
    double* d_b = nullptr;      // RHS in the reference's scaling (for the residual check)
    double* d_rowscale = nullptr;  // symmetrising row scale s_i of the cylindrical operator, [M]
    unsigned long long* d_rho = nullptr;   // [n_species][M*N] fixed-point charge grids
    double* d_scratch = nullptr;           // reductions
    bool grid_set = false;
    bool all_cells_free = false;  // every cell has a FREE corner: t_grid::is_free is true everywhere inside the box
    std::vector<unsigned char> h_mask;
    std::vector<double> h_voltage;

    std::vector<MgLevel> mg;
    double* d_mg_inv = nullptr;   // dense inverse of the coarsest-level operator
    DirectSolver direct;
    Direct3D direct3;
    int solver_kind = MAG2D_SOLVER_AUTO;
    unsigned long long direct_calls = 0;
    int cycles_per_step = 0;
    double solve_tol = 1e-13;
    int max_cycles = 60;
    int last_cycles = 0;
    bool extrapolate = false;     // 2u_n - u_{n-1} as the initial guess of the in-step solve
    bool have_prev = false;
    double* d_u_prev = nullptr;
    bool monitor_armed = false;   // d_scratch[16..17] hold running max |r/a_kk| and max |u| of fixed-cycle solves
    double last_resid = 0;

    std::vector<SpeciesStore> sp;
    double* d_charges = nullptr;  // [n_species]
    int sort_interval = 0;
    bool use_source = false;      // Param::use_source: mag2d_step calls Species::source() after every advance (pic.cpp:346-347)
    int store_layout = MAG2D_LAYOUT_AUTO;   // mag2d_set_store_layout
    bool store_f32 = false;                 // mag2d_set_storage: the particle arrays hold floats (2-D Boris movers; arithmetic stays fp64)
    size_t elem_size() const { return store_f32 ? sizeof(float) : sizeof(double); }
    bool fused_sort = true;       // cell sort carried by the Boris push itself (MAG2D_FUSED_SORT=0: stand-alone passes)
    bool count_collisions = false;

    // sort scratch
    unsigned int* d_cell_count = nullptr;
    unsigned int* d_cell_offset = nullptr;
    unsigned int* d_rank = nullptr;
    unsigned int* d_key = nullptr;
    long long rank_capacity = 0;
    unsigned int* d_block_sums = nullptr;
    unsigned int* d_coll_count = nullptr;

    // streamed step: the particle arrays a push works on are one chunk of a host-resident store (abi.cu)
    const ParticlesDev* chunk_view = nullptr;
    long long chunk_slot0 = 0;
    double* d_chunk[3][6] = {};        // ring of device staging buffers (x, z, vx, vy, vz; y for CARTESIAN3D)
    long long chunk_capacity = 0;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    unsigned long long streamed_h2d = 0, streamed_d2h = 0;   // bytes the streamed steps have staged in each direction (mag2d_streamed_bytes)
    cudaEvent_t ev_h2d[3] = {}, ev_comp[3] = {}, ev_d2h[3] = {};

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t s_comm = nullptr;      // side stream of the per-species charge all-reduce (overlaps the next species' push)
    cudaEvent_t ev_comm_in = nullptr, ev_comm_out = nullptr;
    bool comm_pending = false;

    // timing
    bool timing = false;
    // inside one iteration of mag2d_step, after the solve, nothing writes the potential: without an RF term the edge-difference
    // fields the first species computed serve the following species too
    bool edge_fields_fresh = false, edge_fields_fresh_armed = false;
    cudaEvent_t ev[8] = {};
    double timers[5] = {};
};

// ---- error plumbing (abi.cu)
void mag2d_set_error(const std::string& msg);
#define CUDA_OK(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            mag2d_set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                   \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// ---- store management (abi.cu)
int refresh_pools_all(mag2d_ctx* c);   // partner pools of every species' collision model follow the particle arrays
int store_alloc_slab(mag2d_ctx* c, SpeciesStore& S, int slab, long long capacity);

// ---- launchers implemented in the kernel translation units
// push.cu
int launch_species_advance(mag2d_ctx* c, int s, int sort_mode = 0);
int launch_species_advance_init(mag2d_ctx* c, int s);
int launch_species_accumulate(mag2d_ctx* c, int s);
int launch_generate(mag2d_ctx* c, int s, int kind, long long n, double a, double b, double cc, double d);
int launch_energy_hist(mag2d_ctx* c, int s, int nbins, double emax, double* hist, double* stats);
int launch_field_E(mag2d_ctx* c, int n, const double* x, const double* z, double time, double* Ex, double* Ez);
int source_alloc(mag2d_ctx* c, SpeciesStore& S, long long n);
void source_free(SpeciesStore& S);
int launch_source_generate(mag2d_ctx* c, int s, unsigned factor, long long n);
int launch_source_init(mag2d_ctx* c, int s);
int launch_species_source(mag2d_ctx* c, int s, long long* injected);
int launch_field_B(mag2d_ctx* c, int n, const double* x, const double* z, double* Br, double* Bz, double* Bt);
int launch_aos_to_soa(mag2d_ctx* c, int s, const mag2d_particle* d_aos, long long n_in, long long* n_added);
int launch_soa_to_aos(mag2d_ctx* c, int s, mag2d_particle* d_aos);
int update_ueff(mag2d_ctx* c, double phase, bool rf);
int ensure_particle_scratch(mag2d_ctx* c, long long capacity);
// sort.cu
int launch_sort(mag2d_ctx* c, int s, bool trim);
int sort_fused_begin(mag2d_ctx* c, int s, bool permute, bool count);   // buffers, zeroed counters, other slab
int sort_fused_end(mag2d_ctx* c, int s, bool permute, bool count);     // scan of the new counts, dead tail, slab swap
void sort_fused_free(SpeciesStore& S);
// poisson.cu
int mg_setup(mag2d_ctx* c);
void mg_free(mag2d_ctx* c);
int mg_rhs(mag2d_ctx* c, int rf);          // rho (all species) -> b, Dirichlet rows, scaled copy into mg[0].b
int mg_vcycle(mag2d_ctx* c);
int mg_residual(mag2d_ctx* c, double* resid_max, double* u_max);
int mg_solve(mag2d_ctx* c, int rf, double tol, int max_cycles, int fixed_cycles, int* cycles, double* resid);
int launch_u_smooth(mag2d_ctx* c, int symmetry, double radius);
// poisson_direct.cu
int direct_setup(mag2d_ctx* c);
void direct_free(mag2d_ctx* c);
int direct_solve(mag2d_ctx* c, double* u);
int launch_rho_total(mag2d_ctx* c, double* d_out);
// push3d.cu / poisson3d.cu (coord == MAG2D_CARTESIAN3D)
int launch_species_advance3d(mag2d_ctx* c, int s, bool deposit_only, int sort_mode = 0);
int launch_field_E3d(mag2d_ctx* c, int n, const double* x, const double* y, const double* z, double* Ex, double* Ey, double* Ez);
int update_edge_fields3d(mag2d_ctx* c);
int direct3d_setup(mag2d_ctx* c);
void direct3d_free(mag2d_ctx* c);
int solve3d(mag2d_ctx* c, double* resid_out);
bool solve3d_reads_own_planes_only(const mag2d_ctx* c);   // N ranks, slab-parallel solve, M divisible by N
// push3d_brick.cu
struct Grid3Dev;
struct Push3Args;
int brick_rebuild(mag2d_ctx* c, int s, const Grid3Dev& g);
int launch_brick_push(mag2d_ctx* c, int s, const Push3Args& A, bool mcc, bool deposit, int compact_every);
void brick_outbox_view(const SpeciesStore& S, ParticlesDev& v);
int launch_brick_migrate(mag2d_ctx* c, int s, const Push3Args& A);
void brick_poll_overflow(SpeciesStore& S);
void brick_free(SpeciesStore& S);
inline bool is3d(const mag2d_ctx* c) { return c->g.coord == MAG2D_CARTESIAN3D; }
inline size_t grid_nodes(const mag2d_ctx* c) { return (size_t)c->g.M * c->g.N * (is3d(c) ? (size_t)c->g.K : 1); }
// comm.cu
int comm_allreduce_rho(mag2d_ctx* c);
int comm_allreduce_species_async(mag2d_ctx* c, int s, bool own_slab_only = false);
int comm_allreduce_join(mag2d_ctx* c);
bool comm_has_p2p();
int comm_group_start();
int comm_group_end();
int comm_send(mag2d_ctx* c, const double* buf, size_t count, int peer);
int comm_recv(mag2d_ctx* c, double* buf, size_t count, int peer);
int comm_broadcast(mag2d_ctx* c, double* buf, size_t count, int root);
int comm_allgather_inplace(mag2d_ctx* c, double* buf, size_t count);
