// abi.cu — the C ABI of include/mag2d_b200.h: context, device memory, species/collision model set-up,
// particle store management and the per-step orchestration (Pic<D>::advance, src/pic.cpp:330-358).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ctx.hpp"

static thread_local std::string g_last_error;
void mag2d_set_error(const std::string& msg) { g_last_error = msg; }

#define CHECK_CTX(c)                                   \
    do {                                               \
        if (!(c)) { mag2d_set_error("null context"); return 1; } \
        if (cudaSetDevice((c)->device) != cudaSuccess) { mag2d_set_error("cudaSetDevice failed"); return 1; } \
    } while (0)
#define CHECK_SPECIES(c, s)                                                      \
    do {                                                                         \
        if ((s) < 0 || (s) >= (int)(c)->sp.size()) { mag2d_set_error("species index out of range"); return 1; } \
    } while (0)

namespace {

size_t grid_n(const mag2d_ctx* c) { return grid_nodes(c); }

template <typename T>
__global__ void k_count_alive(const double* __restrict__ x, long long n, unsigned long long* __restrict__ out2)
{
    // out2[0] = live particles, out2[1] = 1 + highest live slot
    unsigned long long cnt = 0, hi = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        if (particle_alive(pld<T>(x, k))) { cnt++; hi = (unsigned long long)k + 1; }
    for (int o = 16; o > 0; o >>= 1)
    {
        cnt += __shfl_xor_sync(MAG2D_FULL_MASK, cnt, o);
        hi = max(hi, __shfl_xor_sync(MAG2D_FULL_MASK, hi, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (cnt) atomicAdd(out2, cnt);
        if (hi) atomicMax(out2 + 1, hi);
    }
}

// ---- host restatement of the collision-rate bookkeeping (BaseSpecies::lifetime_init, svmax_find,
// Interaction::sigma_v; src/particles.cpp:142-206, src/particles.hpp:61-83,415-433) -----------------
struct HostInter
{
    mag2d_interaction_desc d;
    double DE, rate, mu;
    const double* E;
    const double* sigma;
};

double host_table(const double* xd, const double* yd, int n, double x)
{
    if (x >= xd[n - 1]) return yd[n - 1];
    if (x <= xd[0]) return yd[0];
    int j1 = 0, j2 = n - 1;
    while (j2 - j1 > 1)
    {
        const int j3 = (j1 + j2) / 2;
        if (x < xd[j3]) j2 = j3;
        else j1 = j3;
    }
    const double w = (x - xd[j1]) / (xd[j2] - xd[j1]);
    return yd[j1] * (1 - w) + yd[j2] * w;
}

double host_sigma_v(const HostInter& I, const mag2d_species_desc& prim, const mag2d_species_desc& sec, double v)
{
    const double EeV = 0.5 * I.mu * v * v / MAG2D_QE;
    if (I.d.type == MAG2D_COULOMB)
    {
        const double E = EeV * MAG2D_QE;
        const double lambda_D = sqrt(MAG2D_EPS0 * MAG2D_KB * prim.temperature / (prim.density * prim.charge * prim.charge));
        const double Lambda = prim.charge * sec.charge / (4 * M_PI * MAG2D_EPS0 * E);
        return M_PI * Lambda * Lambda * log(lambda_D / Lambda) * v;
    }
    if (I.d.n_table > 0) return host_table(I.E, I.sigma, I.d.n_table, EeV) * v;
    return I.rate;
}

int free_store(SpeciesStore& S)
{
    for (int b = 0; b < 2; b++)
        for (int a = 0; a < N_ARR; a++)
            if (S.arr[b][a]) { cudaFree(S.arr[b][a]); S.arr[b][a] = nullptr; }
    if (S.d_removed) cudaFree(S.d_removed);
    if (S.d_counts) cudaFree(S.d_counts);
    if (S.d_blob) cudaFree(S.d_blob);
    delete S.h_blob;
    sort_fused_free(S);
    brick_free(S);
    source_free(S);
    S = SpeciesStore();
    return 0;
}

bool needs_array(const mag2d_ctx* c, int a)
{
    if (a == ARR_Y) return c->g.coord == MAG2D_CARTESIAN3D;
    if (a == ARR_TTD) return c->g.mover == MAG2D_ADVANCE_MULTICOLL;
    return true;
}

int ensure_capacity(mag2d_ctx* c, SpeciesStore& S, long long need)
{
    S.tickets_valid = false;      // every append goes through here: pending sort tickets do not cover the new slots
    S.bins_valid = false;         // ... and neither do the brick bins
    S.append_epoch++;
    if (need <= S.capacity) return 0;
    long long cap = std::max<long long>(need, (long long)(S.capacity * 1.5) + 1024);
    cap = (cap + 255) / 256 * 256;
    // drop the idle slab first (it is re-created lazily by the next sort)
    for (int a = 0; a < N_ARR; a++)
        if (S.arr[S.cur ^ 1][a]) { cudaFree(S.arr[S.cur ^ 1][a]); S.arr[S.cur ^ 1][a] = nullptr; }
    // the larger slab goes into temporaries and is swapped in only when every array could be allocated: an
    // out-of-memory failure leaves the store as it was
    double* fresh[N_ARR] = {};
    for (int a = 0; a < N_ARR; a++)
    {
        if (!needs_array(c, a)) continue;
        if (cudaMalloc(&fresh[a], c->elem_size() * (size_t)cap) != cudaSuccess)
        {
            cudaGetLastError();
            for (int b = 0; b < N_ARR; b++)
                if (fresh[b]) cudaFree(fresh[b]);
            mag2d_set_error("ensure_capacity: out of device memory growing a particle store");
            return 1;
        }
    }
    for (int a = 0; a < N_ARR; a++)
        if (fresh[a] && S.arr[S.cur][a] && S.n_slots > 0)
            CUDA_OK(cudaMemcpyAsync(fresh[a], S.arr[S.cur][a], c->elem_size() * (size_t)S.n_slots, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int a = 0; a < N_ARR; a++)
    {
        if (S.arr[S.cur][a]) cudaFree(S.arr[S.cur][a]);
        S.arr[S.cur][a] = fresh[a];
    }
    S.capacity = cap;
    return 0;
}

int upload_blob_header(mag2d_ctx* c, SpeciesStore& S)
{
    CUDA_OK(cudaMemcpyAsync(S.d_blob, S.h_blob, offsetof(MccBlob, tab), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// partner pools follow the particle arrays (realloc, sort flips the slab): refresh when stale
int refresh_pools(mag2d_ctx* c, int s)
{
    SpeciesStore& S = c->sp[s];
    if (!S.h_blob || !S.h_blob->has_collisions) return 0;
    bool dirty = false;
    for (int k = 0; k < S.h_blob->n_targets; k++)
    {
        const SpeciesStore& T = c->sp[k];
        PoolDev want;
        memset(&want, 0, sizeof(want));
        const int pool = T.n_slots > 0 ? 1 : 0;   // speclist[k]->particles.size() == 0 -> continuum (particles.cpp:230)
        if (pool && c->store_f32 && S.h_blob->t[k].n_inter > 0)
        {
            mag2d_set_error("fp32 particle storage: collisions with particle partners are not implemented (the partner pools are read as doubles)");
            return 1;
        }
        if (pool)
        {
            want.x = T.arr[T.cur][ARR_X];
            want.vx = T.arr[T.cur][ARR_VX];
            want.vy = T.arr[T.cur][ARR_VY];
            want.vz = T.arr[T.cur][ARR_VZ];
            want.n = T.n_slots;
        }
        if (S.h_blob->t[k].pool != pool || memcmp(&S.h_blob->pool[k], &want, sizeof(want)) != 0)
        {
            S.h_blob->t[k].pool = pool;
            S.h_blob->pool[k] = want;
            dirty = true;
        }
    }
    if (dirty) return upload_blob_header(c, S);
    return 0;
}

// pushes between two sorts of a species.  A context-wide interval of -1 picks it per species from the thermal drift:
// the sort pays off while most particles of a warp still share their cell, i.e. until the thermal displacement
// v_th * dt * K reaches about a third of a cell (measured on the C4 deck: electrons 5, optimum flat from 4 to 6);
// slow species (ions) are re-sorted every 64 pushes, which only serves to compact the removed slots.
int effective_sort_interval(const mag2d_ctx* c, const SpeciesStore& S)
{
    if (S.sort_interval >= 0) return S.sort_interval;
    if (c->sort_interval >= 0) return c->sort_interval;
    if (!(S.desc.mass > 0) || !(S.desc.dt > 0)) return 64;
    const double vth = sqrt(1.380662e-23 * std::max(S.desc.temperature, 0.0) / S.desc.mass);
    const double h = is3d(c) ? std::min(std::min(c->g.dx, c->g.dz), c->g.dy) : std::min(c->g.dx, c->g.dz);
    const double per_step = vth * S.desc.dt / h;
    if (!(per_step > 0)) return 64;
    // three axes to drift along: the same disorder is reached earlier in 3-D (C5: 5.67 ms with 4 pushes vs 5.77 with 6)
    const double k = (is3d(c) ? 0.25 : 0.30) / per_step;
    return k >= 64 ? 64 : k <= 2 ? 2 : (int)(k + 0.5);
}

// which part of the fused cell sort this push of species s takes on (sort.cu): bit 0 permute, bit 1 count.
// With an interval of K pushes the sequence is PERMUTE, K-2 plain pushes, COUNT, PERMUTE, ... (K = 1: both every push)
int fused_sort_mode(const mag2d_ctx* c, const SpeciesStore& S)
{
    if (!c->fused_sort || c->g.mover == MAG2D_ADVANCE_MULTICOLL || S.n_slots == 0) return 0;
    const int K = effective_sort_interval(c, S);
    if (K <= 0) return 0;
    int mode = S.tickets_valid ? 1 : 0;
    const int since = mode ? 0 : S.pushes_since_permute;
    if (since >= K - 1) mode |= 2;
    return mode;
}

int advance_one(mag2d_ctx* c, int s, bool in_step)
{
    if (refresh_pools(c, s)) return 1;
    if (is3d(c))
    {
        // inside mag2d_step a 3-D species runs on the brick-binned store (push3d_brick.cu; mode bit 2) wherever the fused cell sort
        // would be allowed to reorder it; MAG3D_BRICK=0 keeps the slot-order kernel with the fused sort
        static const bool brick_env = !getenv("MAG3D_BRICK") || atoi(getenv("MAG3D_BRICK")) != 0;
        const SpeciesStore& S = c->sp[s];
        const bool bricks = c->store_layout == MAG2D_LAYOUT_BRICKS || (c->store_layout == MAG2D_LAYOUT_AUTO && brick_env);
        if (in_step && bricks && c->fused_sort && !c->use_source && S.n_slots > 0 && effective_sort_interval(c, S) > 0)
            return launch_species_advance3d(c, s, false, 4 | (std::min(effective_sort_interval(c, S), 1 << 20) << 3));
        return launch_species_advance3d(c, s, false, in_step ? fused_sort_mode(c, c->sp[s]) : 0);
    }
    return launch_species_advance(c, s, in_step ? fused_sort_mode(c, c->sp[s]) : 0);
}

// Species<CARTESIAN>::source() of species s (particles.cpp:1158-1226)
int species_source(mag2d_ctx* c, int s, long long* injected)
{
    SpeciesStore& S = c->sp[s];
    if (injected) *injected = 0;
    if (S.src_n == 0) return 0;
    if (c->store_f32) { mag2d_set_error("mag2d_species_source: fp64 particle storage only"); return 1; }
    if (c->g.coord != MAG2D_CARTESIAN || c->g.mover != MAG2D_ADVANCE_BORIS)
    {
        // Species<CYLINDRICAL>::source is declared but never defined in the reference (it does not link)
        mag2d_set_error("mag2d_species_source: CARTESIAN coordinates with the ADVANCE_BORIS mover only");
        return 1;
    }
    if (ensure_capacity(c, S, S.n_slots + 2 * S.src_n + 256)) return 1;
    if (refresh_pools(c, s)) return 1;        // the arrays may have moved
    return launch_species_source(c, s, injected);
}

}  // namespace

int refresh_pools_all(mag2d_ctx* c)
{
    for (size_t s = 0; s < c->sp.size(); s++)
        if (refresh_pools(c, (int)s)) return 1;
    return 0;
}

int store_alloc_slab(mag2d_ctx* c, SpeciesStore& S, int slab, long long capacity)
{
    for (int a = 0; a < N_ARR; a++)
    {
        if (!needs_array(c, a)) continue;
        if (S.arr[slab][a]) { cudaFree(S.arr[slab][a]); S.arr[slab][a] = nullptr; }
        CUDA_OK(cudaMalloc(&S.arr[slab][a], c->elem_size() * (size_t)capacity));
    }
    S.capacity = capacity;
    return 0;
}

extern "C" {

int mag2d_abi_version(void) { return MAG2D_ABI_VERSION; }
const char* mag2d_last_error(void) { return g_last_error.c_str(); }

int mag2d_create(int device, const mag2d_grid_desc* grid, void* stream, mag2d_ctx** out)
{
    if (!grid || !out) { mag2d_set_error("mag2d_create: null argument"); return 1; }
    if (grid->coord != MAG2D_CARTESIAN && grid->coord != MAG2D_CYLINDRICAL && grid->coord != MAG2D_CARTESIAN3D)
    {
        mag2d_set_error("mag2d_create: coord must be CARTESIAN, CYLINDRICAL or CARTESIAN3D");
        return 1;
    }
    if (grid->M < 2 || grid->N < 2) { mag2d_set_error("mag2d_create: grid must be at least 2x2"); return 1; }
    if (grid->coord == MAG2D_CARTESIAN3D)
    {
        if (grid->K < 3 || grid->M < 3 || grid->N < 3) { mag2d_set_error("mag2d_create: a 3-D grid must be at least 3x3x3"); return 1; }
        if (grid->rf || grid->mover != MAG2D_ADVANCE_BORIS) { mag2d_set_error("mag2d_create: CARTESIAN3D supports the Boris mover without RF only"); return 1; }
    }
    // Param's own validation (param.cpp:60-64, 91-95)
    if (grid->selfconsistent && grid->rf) { mag2d_set_error("Param: selfconsistent rf trap not implemented\n"); return 1; }
    if (grid->selfconsistent && grid->electric_field_from_file)
    {
        mag2d_set_error("Param: selfconsistent with electric_field_from_file not implemented");
        return 1;
    }
    if (grid->coord == MAG2D_CYLINDRICAL && grid->boundary != MAG2D_BOUNDARY_FREE)
    {
        mag2d_set_error("Param: only FREE boundary condition in cylindrical coords is implemented\n");
        return 1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        mag2d_set_error("mag2d_create: no CUDA device available (this library has no CPU path)");
        return 1;
    }
    if (device < 0 || device >= ndev) { mag2d_set_error("mag2d_create: device index out of range"); return 1; }
    CUDA_OK(cudaSetDevice(device));
    mag2d_ctx* c = new mag2d_ctx;
    c->device = device;
    c->g = *grid;
    if (const char* e = getenv("MAG2D_FUSED_SORT")) c->fused_sort = atoi(e) != 0;
    if (stream) c->stream = (cudaStream_t)stream;
    else
    {
        CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    const size_t n = grid_n(c);
    CUDA_OK(cudaMalloc(&c->d_mask, n));
    CUDA_OK(cudaMalloc(&c->d_voltage, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&c->d_u, sizeof(double) * n));
    const bool three_d = grid->coord == MAG2D_CARTESIAN3D;
    // uRF / the scratch copy are 2-D features; a 3-D grid only gets one-node stand-ins (the potential alone is 134 MB at 256^3)
    CUDA_OK(cudaMalloc(&c->d_uRF, sizeof(double) * (three_d ? 1 : n)));
    CUDA_OK(cudaMalloc(&c->d_ueff, sizeof(double) * (three_d ? 1 : n)));
    // the edge fields carry one ghost plane per axis (push.cu: gather_E, push3d.cu: grad_component)
    const size_t n_edge = (size_t)(grid->M + 1) * (grid->N + 1) * (three_d ? (size_t)(grid->K + 1) : 1);
    CUDA_OK(cudaMalloc(&c->d_gx, sizeof(double) * n_edge));
    CUDA_OK(cudaMalloc(&c->d_gz, sizeof(double) * n_edge));
    if (three_d)
    {
        CUDA_OK(cudaMalloc(&c->d_gy, sizeof(double) * n_edge));
        CUDA_OK(cudaMemsetAsync(c->d_gy, 0, sizeof(double) * n_edge, c->stream));
    }
    CUDA_OK(cudaMalloc(&c->d_cfree, n));
    CUDA_OK(cudaMemsetAsync(c->d_cfree, 1, n, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_gx, 0, sizeof(double) * n_edge, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_gz, 0, sizeof(double) * n_edge, c->stream));
    CUDA_OK(cudaMalloc(&c->d_b, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&c->d_scratch, sizeof(double) * 64));
    CUDA_OK(cudaMemsetAsync(c->d_scratch, 0, sizeof(double) * 64, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_u, 0, sizeof(double) * n, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_uRF, 0, sizeof(double) * (three_d ? 1 : n), c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_ueff, 0, sizeof(double) * (three_d ? 1 : n), c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_voltage, 0, sizeof(double) * n, c->stream));
    CUDA_OK(cudaMemsetAsync(c->d_mask, MAG2D_FREE, n, c->stream));
    for (int q = 0; q < 8; q++) CUDA_OK(cudaEventCreate(&c->ev[q]));
    *out = c;
    return 0;
}

int mag2d_destroy(mag2d_ctx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    mag2d_comm_destroy(c);
    mg_free(c);
    direct_free(c);
    direct3d_free(c);
    for (auto& S : c->sp) free_store(S);
    cudaFree(c->d_gy);
    for (int b = 0; b < 3; b++)
    {
        for (int a = 0; a < 6; a++) cudaFree(c->d_chunk[b][a]);
        if (c->ev_h2d[b]) { cudaEventDestroy(c->ev_h2d[b]); cudaEventDestroy(c->ev_comp[b]); cudaEventDestroy(c->ev_d2h[b]); }
    }
    if (c->s_h2d) { cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h); }
    if (c->s_comm) { cudaStreamDestroy(c->s_comm); cudaEventDestroy(c->ev_comm_in); cudaEventDestroy(c->ev_comm_out); }
    cudaFree(c->d_mask);
    cudaFree(c->d_voltage);
    cudaFree(c->d_u);
    cudaFree(c->d_uRF);
    cudaFree(c->d_ueff);
    cudaFree(c->d_gx);
    cudaFree(c->d_btab_r);
    cudaFree(c->d_btab_z);
    cudaFree(c->d_gz);
    cudaFree(c->d_cfree);
    cudaFree(c->d_b);
    cudaFree(c->d_scratch);
    if (c->d_u_prev) cudaFree(c->d_u_prev);
    if (c->d_rho) cudaFree(c->d_rho);
    if (c->d_charges) cudaFree(c->d_charges);
    if (c->d_cell_count) cudaFree(c->d_cell_count);
    if (c->d_cell_offset) cudaFree(c->d_cell_offset);
    if (c->d_block_sums) cudaFree(c->d_block_sums);
    if (c->d_coll_count) cudaFree(c->d_coll_count);
    if (c->d_rank) cudaFree(c->d_rank);
    if (c->d_key) cudaFree(c->d_key);
    for (int q = 0; q < 8; q++)
        if (c->ev[q]) cudaEventDestroy(c->ev[q]);
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int mag2d_sync(mag2d_ctx* c)
{
    CHECK_CTX(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_seed(mag2d_ctx* c, uint64_t seed)
{
    CHECK_CTX(c);
    c->seed = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
    return 0;
}

int mag2d_set_grid(mag2d_ctx* c, const uint8_t* mask, const double* voltage)
{
    CHECK_CTX(c);
    const size_t n = grid_n(c);
    c->h_mask.assign(mask, mask + n);
    c->h_voltage.assign(voltage, voltage + n);
    CUDA_OK(cudaMemcpyAsync(c->d_mask, mask, n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_voltage, voltage, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    // t_grid::is_free per cell (fields.hpp:94-101): a particle survives when any corner of its cell is FREE
    std::vector<unsigned char> cfree(n, 0);
    const int M = c->g.M, N = c->g.N;
    if (is3d(c))
    {
        // Geometry::is_free, fields3d.hpp:48-57: eight corners
        const int K = c->g.K;
        const size_t sj = N, si = (size_t)K * N;
        c->all_cells_free = true;       // then k_push3d skips the per-particle flag load, as in 2-D
        for (int i = 0; i + 1 < M; i++)
            for (int j = 0; j + 1 < K; j++)
                for (int k = 0; k + 1 < N; k++)
                {
                    const size_t m = ((size_t)i * K + j) * N + k;
                    bool f = false;
                    for (int q = 0; q < 8; q++) f = f || mask[m + (q & 1 ? si : 0) + (q & 2 ? sj : 0) + (q >> 2)] == MAG2D_FREE;
                    cfree[m] = f;
                    if (!f) c->all_cells_free = false;
                }
        CUDA_OK(cudaMemcpyAsync(c->d_cfree, cfree.data(), n, cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        c->grid_set = true;
        return direct3d_setup(c);
    }
    for (int i = 0; i + 1 < M; i++)
        for (int j = 0; j + 1 < N; j++)
        {
            const size_t k = (size_t)i * N + j;
            cfree[k] = mask[k] == MAG2D_FREE || mask[k + N] == MAG2D_FREE || mask[k + 1] == MAG2D_FREE || mask[k + N + 1] == MAG2D_FREE;
        }
    CUDA_OK(cudaMemcpyAsync(c->d_cfree, cfree.data(), n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    // no electrode blocks a whole cell (geometry EMPTY): the push kernels can skip the per-particle flag load
    c->all_cells_free = true;
    for (int i = 0; i + 1 < M && c->all_cells_free; i++)
        for (int j = 0; j + 1 < N; j++)
            if (!cfree[(size_t)i * N + j]) { c->all_cells_free = false; break; }
    c->grid_set = true;
    if (mg_setup(c)) return 1;
    return direct_setup(c);
}

int mag2d_set_potential(mag2d_ctx* c, int which, const double* values)
{
    CHECK_CTX(c);
    if (is3d(c) && which != 0) { mag2d_set_error("CARTESIAN3D has no RF potential"); return 1; }
    double* dst = which == 0 ? c->d_u : c->d_uRF;
    CUDA_OK(cudaMemcpyAsync(dst, values, sizeof(double) * grid_n(c), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_set_magnetic_field(mag2d_ctx* c, int r_sampl, int z_sampl, double dr, double dz, double r_min, double z_min, const double* Br,
                             const double* Bz)
{
    CHECK_CTX(c);
    if (is3d(c)) { mag2d_set_error("mag2d_set_magnetic_field: CARTESIAN3D uses the constant field of the grid descriptor"); return 1; }
    if (c->d_btab_r) { cudaFree(c->d_btab_r); c->d_btab_r = nullptr; }
    if (c->d_btab_z) { cudaFree(c->d_btab_z); c->d_btab_z = nullptr; }
    c->btab_M = c->btab_N = 0;
    if (!Br && !Bz) return 0;
    if (!Br || !Bz || r_sampl < 2 || z_sampl < 2 || !(dr > 0) || !(dz > 0))
    {
        mag2d_set_error("Fields::load_magnetic_field() wrong size of input vector");
        return 1;
    }
    // Field2D::interpolate accepts (int)((x - xmin)*idx) in [0, jmax-1] (Field2D.hpp:74): the box has to lie inside
    const double r_top = r_min + (r_sampl - 1) * dr, z_top = z_min + (z_sampl - 1) * dz;
    const double er = 1e-9 * dr, ez = 1e-9 * dz;
    if (r_min > er || z_min > ez || r_top < c->g.x_max - er || z_top < c->g.z_max - ez)
    {
        mag2d_set_error("Field2D::interpolate() outside of range");
        return 1;
    }
    const size_t n = (size_t)r_sampl * z_sampl;
    for (size_t k = 0; k < n; k++)
        if (std::isnan(Br[k]) || std::isnan(Bz[k])) { mag2d_set_error("Fields::load_magnetic_field() garbage loaded"); return 1; }
    CUDA_OK(cudaMalloc(&c->d_btab_r, sizeof(double) * n));
    CUDA_OK(cudaMalloc(&c->d_btab_z, sizeof(double) * n));
    CUDA_OK(cudaMemcpyAsync(c->d_btab_r, Br, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->d_btab_z, Bz, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    c->btab_M = r_sampl;
    c->btab_N = z_sampl;
    c->btab_idx = 1.0 / dr;      // Field2D::resize: idx = 1.0/dx (Field2D.hpp)
    c->btab_idz = 1.0 / dz;
    c->btab_xmin = r_min;
    c->btab_zmin = z_min;
    return 0;
}

int mag2d_field_B(mag2d_ctx* c, int n, const double* x, const double* z, double* Br, double* Bz, double* Bt)
{
    CHECK_CTX(c);
    if (is3d(c)) { mag2d_set_error("mag2d_field_B: 2-D grids only"); return 1; }
    return launch_field_B(c, n, x, z, Br, Bz, Bt);
}

int mag2d_get_potential(mag2d_ctx* c, int which, double* values)
{
    CHECK_CTX(c);
    if (is3d(c) && which != 0) { mag2d_set_error("CARTESIAN3D has no RF potential"); return 1; }
    const double* src = which == 0 ? c->d_u : c->d_uRF;
    CUDA_OK(cudaMemcpyAsync(values, src, sizeof(double) * grid_n(c), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_solve(mag2d_ctx* c, int rf, double tol, int max_cycles, int* cycles_out, double* resid_out)
{
    CHECK_CTX(c);
    if (c->sp.empty() && !c->d_rho)
    {
        // no species yet (the Pic constructor pre-solves the vacuum fields, pic.cpp:180-187): zero charge
        CUDA_OK(cudaMalloc(&c->d_rho, sizeof(unsigned long long) * grid_n(c)));
        CUDA_OK(cudaMemsetAsync(c->d_rho, 0, sizeof(unsigned long long) * grid_n(c), c->stream));
    }
    if (is3d(c))
    {
        if (rf) { mag2d_set_error("CARTESIAN3D has no RF potential"); return 1; }
        if (cycles_out) *cycles_out = 0;
        double r = 0.0;
        if (solve3d(c, &r)) return 1;
        if (resid_out) *resid_out = r;
        c->last_cycles = 0;
        c->last_resid = r;
        return 0;
    }
    return mg_solve(c, rf, tol, max_cycles, 0, cycles_out, resid_out);
}

int mag2d_set_solver(mag2d_ctx* c, int cycles_per_step, double tol, int max_cycles)
{
    CHECK_CTX(c);
    // a negative cycle count selects |cycles| V-cycles per step with the time-extrapolated first guess
    c->extrapolate = cycles_per_step < 0;
    c->cycles_per_step = cycles_per_step < 0 ? -cycles_per_step : cycles_per_step;
    c->solve_tol = tol;
    c->max_cycles = max_cycles;
    c->have_prev = false;
    return 0;
}

int mag2d_set_solver_kind(mag2d_ctx* c, int kind)
{
    CHECK_CTX(c);
    if (kind < MAG2D_SOLVER_AUTO || kind > MAG2D_SOLVER_DIRECT)
    {
        mag2d_set_error("mag2d_set_solver_kind: unknown solver kind");
        return 1;
    }
    if (is3d(c))
    {
        if (kind == MAG2D_SOLVER_MULTIGRID) { mag2d_set_error("mag2d_set_solver_kind: CARTESIAN3D has the direct solver only"); return 1; }
        return 0;
    }
    if (kind == MAG2D_SOLVER_DIRECT && !c->direct.ok)
    {
        mag2d_set_error("mag2d_set_solver_kind: the grid does not separate (electrodes inside free rows); use the multigrid solver");
        return 1;
    }
    c->solver_kind = kind;
    c->have_prev = false;
    return 0;
}

int mag2d_solver_is_direct(mag2d_ctx* c)
{
    if (!c) return 0;
    if (is3d(c)) return c->direct3.ok;
    return c->direct.ok && c->solver_kind != MAG2D_SOLVER_MULTIGRID;
}

int mag2d_solver_stats(mag2d_ctx* c, int* last_cycles, double* last_resid)
{
    CHECK_CTX(c);
    if (c->monitor_armed)
    {
        // fixed-cycle solves do not look at the residual themselves: read (and reset) the running maxima
        double h[2];
        CUDA_OK(cudaMemcpyAsync(h, c->d_scratch + 16, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemsetAsync(c->d_scratch + 16, 0, sizeof(h), c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        c->last_resid = h[1] > 0 ? h[0] / h[1] : h[0];
        c->monitor_armed = false;
    }
    if (last_cycles) *last_cycles = c->last_cycles;
    if (last_resid) *last_resid = c->last_resid;
    return 0;
}

int mag2d_u_smooth(mag2d_ctx* c, int symmetry, double radius)
{
    CHECK_CTX(c);
    return launch_u_smooth(c, symmetry, radius);
}

int mag2d_field_E3(mag2d_ctx* c, int n, const double* x, const double* y, const double* z, double* Ex, double* Ey, double* Ez)
{
    CHECK_CTX(c);
    if (!is3d(c)) { mag2d_set_error("mag2d_field_E3: CARTESIAN3D only"); return 1; }
    return launch_field_E3d(c, n, x, y, z, Ex, Ey, Ez);
}

int mag2d_field_E(mag2d_ctx* c, int n, const double* x, const double* z, double time, double* Ex, double* Ez)
{
    CHECK_CTX(c);
    if (is3d(c)) { mag2d_set_error("mag2d_field_E: use mag2d_field_E3 for CARTESIAN3D"); return 1; }
    return launch_field_E(c, n, x, z, time, Ex, Ez);
}

int mag2d_set_species(mag2d_ctx* c, int ns, const mag2d_species_desc* species, int ni,
                      const mag2d_interaction_desc* inter, const double* table_E, const double* table_sigma,
                      int n_table_total)
{
    CHECK_CTX(c);
    if (ns < 1 || ns > MAG2D_MAX_SPECIES) { mag2d_set_error("mag2d_set_species: 1..16 species supported"); return 1; }
    for (auto& S : c->sp) free_store(S);
    c->sp.assign(ns, SpeciesStore());
    const size_t n = grid_n(c);
    if (c->d_rho) cudaFree(c->d_rho);
    if (c->d_charges) cudaFree(c->d_charges);
    CUDA_OK(cudaMalloc(&c->d_rho, sizeof(unsigned long long) * n * ns));
    CUDA_OK(cudaMemsetAsync(c->d_rho, 0, sizeof(unsigned long long) * n * ns, c->stream));
    CUDA_OK(cudaMalloc(&c->d_charges, sizeof(double) * ns));
    std::vector<double> charges(ns);
    for (int i = 0; i < ns; i++)
    {
        SpeciesStore& S = c->sp[i];
        S.desc = species[i];
        // BaseSpecies ctor (particles.hpp:164,180)
        S.E_max = species[i].E_max > 0. ? species[i].E_max : species[i].temperature * MAG2D_KB / MAG2D_QE * 10.0;
        S.v_max = sqrt(2.0 * MAG2D_KB * species[i].temperature / species[i].mass);
        S.rates.assign(ns, 0.0);
        charges[i] = species[i].charge;
        CUDA_OK(cudaMalloc(&S.d_removed, sizeof(unsigned long long)));
        CUDA_OK(cudaMemsetAsync(S.d_removed, 0, sizeof(unsigned long long), c->stream));
        CUDA_OK(cudaMalloc(&S.d_counts, sizeof(unsigned long long) * (ns + 1) * 16));
        CUDA_OK(cudaMemsetAsync(S.d_counts, 0, sizeof(unsigned long long) * (ns + 1) * 16, c->stream));
    }
    CUDA_OK(cudaMemcpyAsync(c->d_charges, charges.data(), sizeof(double) * ns, cudaMemcpyHostToDevice, c->stream));
    // Interaction ctor (particles.hpp:74-83) and the per-primary, per-target lists of Speclist (pic.cpp:46-72)
    std::vector<std::vector<std::vector<HostInter>>> by(ns, std::vector<std::vector<HostInter>>(ns));
    for (int k = 0; k < ni; k++)
    {
        const mag2d_interaction_desc& d = inter[k];
        if (d.primary < 0 || d.primary >= ns || d.secondary < 0 || d.secondary >= ns)
        {
            mag2d_set_error("Speclist::Speclist: unrecognized primary species of interaction\n");
            return 1;
        }
        if (d.n_table > 0 && (d.table_offset < 0 || d.table_offset + d.n_table > n_table_total))
        {
            mag2d_set_error("mag2d_set_species: cross-section table out of range");
            return 1;
        }
        HostInter I;
        I.d = d;
        I.DE = d.DE_eV * MAG2D_QE;
        I.rate = d.rate;
        if (d.type == MAG2D_LANGEVIN) I.rate *= d.cutoff * d.cutoff;
        const double m1 = species[d.primary].mass, m2 = species[d.secondary].mass;
        I.mu = m1 * m2 / (m1 + m2);
        I.E = d.n_table > 0 ? table_E + d.table_offset : nullptr;
        I.sigma = d.n_table > 0 ? table_sigma + d.table_offset : nullptr;
        by[d.primary][d.secondary].push_back(I);
    }
    for (int i = 0; i < ns; i++)
    {
        SpeciesStore& S = c->sp[i];
        // lifetime_init (particles.cpp:151-161) with svmax_find (particles.cpp:190-206): 1000 samples of
        // v in [0, veV(E_max)), v advanced by repeated addition as the reference does
        const double vmax = sqrt(S.E_max * MAG2D_QE / S.desc.mass * 2.0);
        double rate = 0;
        for (int k = 0; k < ns; k++)
        {
            const double dv = vmax / 1000;
            double svmax = 0.0;
            for (double v = 0; v < vmax; v += dv)
            {
                double sv = 0;
                for (const HostInter& I : by[i][k]) sv += host_sigma_v(I, species[i], species[k], v);
                if (std::isnan(sv)) continue;
                if (sv > svmax) svmax = sv;
            }
            S.rates[k] = svmax * species[k].density;
            rate += S.rates[k];
        }
        S.lifetime = rate > 0.0 ? 1.0 / rate : INFINITY;
        // device blob
        MccBlob* B = new MccBlob;
        memset(B, 0, sizeof(MccBlob));
        B->n_targets = ns;
        B->lifetime = S.lifetime;
        B->inv_lifetime = rate;
        B->mass = S.desc.mass;
        B->charge = S.desc.charge;
        int ii = 0, nt = 0;
        std::vector<double> tE, tS;
        for (int k = 0; k < ns; k++)
        {
            MccTarget& T = B->t[k];
            T.rate_max = S.rates[k];
            T.density = species[k].density;
            T.mass = species[k].mass;
            T.vth = c->sp[k].v_max * M_SQRT1_2;
            T.inv_M = 1.0 / (S.desc.mass + species[k].mass);
            T.first_inter = ii;
            T.n_inter = (int)by[i][k].size();
            for (const HostInter& I : by[i][k])
            {
                if (ii >= MCC_MAX_I) { mag2d_set_error("mag2d_set_species: more than 32 interactions for one primary species"); delete B; return 1; }
                MccInter& D = B->in[ii++];
                D.type = I.d.type;
                D.n_table = I.d.n_table;
                D.table_off = nt;
                D.DE = I.DE;
                D.rate = I.rate;
                D.cutoff = I.d.cutoff;
                D.mu = I.mu;
                D.half_mu = 0.5 * I.mu;
                for (int q = 0; q < I.d.n_table; q++) { tE.push_back(I.E[q]); tS.push_back(I.sigma[q]); }
                nt += I.d.n_table;
            }
        }
        if (nt > MCC_MAX_TAB) { mag2d_set_error("mag2d_set_species: cross-section tables exceed 2048 points for one primary species"); delete B; return 1; }
        B->n_inter_total = ii;
        B->n_tab = nt;
        B->has_collisions = rate > 0.0;
        for (int q = 0; q < nt; q++)
        {
            B->tab[q] = tE[q];
            B->tab[nt + q] = tS[q];
            // the interpolation weight of vec_interpolate (tabulate.cpp:124-140) as a product instead of a division per look-up;
            // the last point of a table (and of all tables) has no interval behind it
            const double w = q + 1 < nt ? tE[q + 1] - tE[q] : 0.0;
            B->tab[2 * nt + q] = w > 0.0 ? 1.0 / w : 0.0;
        }
        S.h_blob = B;
        CUDA_OK(cudaMalloc(&S.d_blob, sizeof(MccBlob)));
        CUDA_OK(cudaMemcpyAsync(S.d_blob, B, sizeof(MccBlob), cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_species_get(mag2d_ctx* c, int s, int what, double* out)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    const SpeciesStore& S = c->sp[s];
    switch (what)
    {
        case 0: *out = S.lifetime; break;
        case 1: *out = S.v_max; break;
        case 2: *out = S.E_max; break;
        case 3: *out = S.t; break;
        case 4: *out = (double)S.niter; break;
        case 5: *out = 1.0 - exp(-S.desc.dt / S.lifetime); break;
        default: mag2d_set_error("mag2d_species_get: unknown selector"); return 1;
    }
    return 0;
}

int mag2d_species_rates(mag2d_ctx* c, int s, double* rates)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    for (size_t k = 0; k < c->sp.size(); k++) rates[k] = c->sp[s].rates[k];
    return 0;
}

int mag2d_collision_counts(mag2d_ctx* c, int s, int64_t* counts, int reset)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    const size_t bytes = sizeof(unsigned long long) * (c->sp.size() + 1) * 16;
    CUDA_OK(cudaMemcpyAsync(counts, c->sp[s].d_counts, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (reset) CUDA_OK(cudaMemsetAsync(c->sp[s].d_counts, 0, bytes, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_set_collision_counting(mag2d_ctx* c, int enable)
{
    CHECK_CTX(c);
    c->count_collisions = enable != 0;
    return 0;
}

int mag2d_reserve(mag2d_ctx* c, int s, int64_t capacity)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    return ensure_capacity(c, c->sp[s], capacity);
}

int mag2d_particles_upload(mag2d_ctx* c, int s, const mag2d_particle* aos, int64_t n)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (n <= 0) return 0;
    SpeciesStore& S = c->sp[s];
    if (ensure_capacity(c, S, S.n_slots + n)) return 1;
    mag2d_particle* d_aos;
    CUDA_OK(cudaMalloc(&d_aos, sizeof(mag2d_particle) * (size_t)n));
    CUDA_OK(cudaMemcpyAsync(d_aos, aos, sizeof(mag2d_particle) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    long long added = 0;
    const int rc = launch_aos_to_soa(c, s, d_aos, n, &added);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_aos));
    return rc;
}

int mag2d_particles_upload_soa(mag2d_ctx* c, int s, int64_t n, const double* x, const double* y, const double* z,
                               const double* vx, const double* vy, const double* vz, const double* ttd)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (n <= 0) return 0;
    SpeciesStore& S = c->sp[s];
    if (ensure_capacity(c, S, S.n_slots + n)) return 1;
    const double* src[N_ARR] = {x, z, vx, vy, vz, y, ttd};
    std::vector<float> narrow;
    for (int a = 0; a < N_ARR; a++)
    {
        double* dst = S.arr[S.cur][a];
        if (!dst) continue;
        if (c->store_f32)
        {
            // fp32 storage: the host arrays stay double (the interface of the reference's t_particle), rounded here
            float* dstf = reinterpret_cast<float*>(dst) + S.n_slots;
            if (src[a])
            {
                narrow.resize((size_t)n);
                for (int64_t k = 0; k < n; k++) narrow[(size_t)k] = (float)src[a][k];
                CUDA_OK(cudaMemcpyAsync(dstf, narrow.data(), sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
                CUDA_OK(cudaStreamSynchronize(c->stream));
            }
            else CUDA_OK(cudaMemsetAsync(dstf, 0, sizeof(float) * (size_t)n, c->stream));
            continue;
        }
        if (src[a]) CUDA_OK(cudaMemcpyAsync(dst + S.n_slots, src[a], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
        else CUDA_OK(cudaMemsetAsync(dst + S.n_slots, 0, sizeof(double) * (size_t)n, c->stream));
    }
    S.n_slots += n;
    return 0;
}

int mag2d_particles_download(mag2d_ctx* c, int s, mag2d_particle* aos, int64_t capacity, int64_t* n_slots)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    if (n_slots) *n_slots = S.n_slots;
    if (S.n_slots == 0) return 0;
    if (capacity < S.n_slots) { mag2d_set_error("mag2d_particles_download: buffer too small"); return 1; }
    mag2d_particle* d_aos;
    CUDA_OK(cudaMalloc(&d_aos, sizeof(mag2d_particle) * (size_t)S.n_slots));
    const int rc = launch_soa_to_aos(c, s, d_aos);
    if (!rc) CUDA_OK(cudaMemcpyAsync(aos, d_aos, sizeof(mag2d_particle) * (size_t)S.n_slots, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d_aos));
    return rc;
}

int mag2d_particles_download_soa(mag2d_ctx* c, int s, int64_t capacity, double* x, double* y, double* z, double* vx,
                                 double* vy, double* vz, double* ttd, uint8_t* alive, int64_t* n_slots)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    if (n_slots) *n_slots = S.n_slots;
    if (S.n_slots == 0) return 0;
    if (capacity < S.n_slots) { mag2d_set_error("mag2d_particles_download_soa: buffer too small"); return 1; }
    double* dst[N_ARR] = {x, z, vx, vy, vz, y, ttd};
    std::vector<float> narrow;
    for (int a = 0; a < N_ARR; a++)
    {
        if (!dst[a]) continue;
        const double* src = S.arr[S.cur][a];
        if (src && c->store_f32)
        {
            narrow.resize((size_t)S.n_slots);
            CUDA_OK(cudaMemcpyAsync(narrow.data(), src, sizeof(float) * (size_t)S.n_slots, cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(cudaStreamSynchronize(c->stream));
            for (long long k = 0; k < S.n_slots; k++) dst[a][k] = (double)narrow[(size_t)k];
        }
        else if (src) CUDA_OK(cudaMemcpyAsync(dst[a], src, sizeof(double) * (size_t)S.n_slots, cudaMemcpyDeviceToHost, c->stream));
        else memset(dst[a], 0, sizeof(double) * (size_t)S.n_slots);
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (alive)
    {
        if (!x) { mag2d_set_error("mag2d_particles_download_soa: alive needs x"); return 1; }
        for (long long k = 0; k < S.n_slots; k++) alive[k] = x[k] == x[k] ? 1 : 0;
    }
    return 0;
}

int mag2d_particles_clear(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    c->sp[s].n_slots = 0;
    c->sp[s].tickets_valid = false;
    c->sp[s].bins_valid = false;
    c->sp[s].append_epoch++;
    CUDA_OK(cudaMemsetAsync(c->sp[s].d_removed, 0, sizeof(unsigned long long), c->stream));
    return 0;
}

int mag2d_count(mag2d_ctx* c, int s, int64_t* n_alive, int64_t* n_slots)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    unsigned long long h[2] = {0, 0};
    if (S.n_slots > 0)
    {
        unsigned long long* d = reinterpret_cast<unsigned long long*>(c->d_scratch + 8);
        CUDA_OK(cudaMemsetAsync(d, 0, sizeof(h), c->stream));
        const unsigned blocks = (unsigned)std::min<long long>((S.n_slots + 255) / 256, 148 * 16);
        if (c->store_f32) k_count_alive<float><<<blocks, 256, 0, c->stream>>>(S.arr[S.cur][ARR_X], S.n_slots, d);
        else k_count_alive<double><<<blocks, 256, 0, c->stream>>>(S.arr[S.cur][ARR_X], S.n_slots, d);
        c->launches++;
        CUDA_OK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    if (n_alive) *n_alive = (int64_t)h[0];
    if (n_slots) *n_slots = S.n_slots;
    return 0;
}

int mag2d_particles_generate(mag2d_ctx* c, int s, int kind, int64_t n, double a, double b, double cc, double d)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (kind < 0 || kind > 2) { mag2d_set_error("mag2d_particles_generate: unknown loader"); return 1; }
    if (kind == 2 && c->g.coord != MAG2D_CYLINDRICAL) { mag2d_set_error("mag2d_particles_generate: loader 2 is cylindrical only"); return 1; }
    if (kind == 2 && (b < 0 || b > c->g.z_max)) return 0;   // particles.cpp:489
    if (is3d(c) && kind != 0) { mag2d_set_error("mag2d_particles_generate: CARTESIAN3D has the uniform loader (kind 0) only"); return 1; }
    SpeciesStore& S = c->sp[s];
    if (ensure_capacity(c, S, S.n_slots + n)) return 1;
    return launch_generate(c, s, kind, n, a, b, cc, d);
}

static int sort_species(mag2d_ctx* c, int s, bool trim)
{
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    if (S.n_slots == 0) return 0;
    if (!S.arr[S.cur ^ 1][ARR_X])
    {
        const long long cap = S.capacity;
        if (store_alloc_slab(c, S, S.cur ^ 1, cap)) return 1;
    }
    return launch_sort(c, s, trim);
}

int mag2d_set_use_source(mag2d_ctx* c, int on)
{
    CHECK_CTX(c);
    c->use_source = on != 0;
    return 0;
}

int mag2d_source_refresh(mag2d_ctx* c, int s, uint32_t factor, double V)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (factor == 0) { mag2d_set_error("mag2d_source_refresh: factor must be positive"); return 1; }
    if (c->g.coord != MAG2D_CARTESIAN) { mag2d_set_error("mag2d_source_refresh: CARTESIAN coordinates only"); return 1; }
    SpeciesStore& S = c->sp[s];
    // particles.cpp:1057-1064: N = density*V, n = (unsigned)(N/factor); species without particles keep an empty reservoir
    const double N = S.desc.density * V;
    const unsigned n = (unsigned)(N / factor);
    if (S.n_slots == 0) return 0;
    return launch_source_generate(c, s, factor, (long long)n);
}

int mag2d_source_upload(mag2d_ctx* c, int s, uint32_t factor, const mag2d_particle* aos, int64_t n)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (factor == 0 || n < 0) { mag2d_set_error("mag2d_source_upload: bad arguments"); return 1; }
    SpeciesStore& S = c->sp[s];
    // the copies below are blocking copies on the legacy stream, which does not order against the context's
    // non-blocking stream: let the kernels that still read the old reservoir finish first
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (source_alloc(c, S, n)) return 1;
    S.src_factor = factor;
    std::vector<double> col((size_t)std::max<int64_t>(n, 1));
    for (int a = 0; a < 6; a++)
    {
        for (int64_t k = 0; k < n; k++)
        {
            const mag2d_particle& p = aos[k];
            col[(size_t)k] = a == 0 ? p.x : a == 1 ? p.z : a == 2 ? p.vx : a == 3 ? p.vy : a == 4 ? p.vz : p.time_to_death;
        }
        if (n > 0) CUDA_OK(cudaMemcpy(S.src[a], col.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    }
    return 0;
}

int mag2d_source_download(mag2d_ctx* c, int s, mag2d_particle* aos, int64_t capacity, int64_t* n_out)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    if (n_out) *n_out = S.src_n;
    if (S.src_n == 0 || !aos) return 0;
    if (capacity < S.src_n) { mag2d_set_error("mag2d_source_download: buffer too small"); return 1; }
    CUDA_OK(cudaStreamSynchronize(c->stream));
    std::vector<double> col((size_t)S.src_n);
    for (long long k = 0; k < S.src_n; k++)
    {
        memset(&aos[k], 0, sizeof(mag2d_particle));
        aos[k].empty = 0;
    }
    for (int a = 0; a < 6; a++)
    {
        CUDA_OK(cudaMemcpy(col.data(), S.src[a], sizeof(double) * (size_t)S.src_n, cudaMemcpyDeviceToHost));
        for (long long k = 0; k < S.src_n; k++)
        {
            mag2d_particle& p = aos[k];
            (a == 0 ? p.x : a == 1 ? p.z : a == 2 ? p.vx : a == 3 ? p.vy : a == 4 ? p.vz : p.time_to_death) = col[(size_t)k];
        }
    }
    return 0;
}

int mag2d_species_source(mag2d_ctx* c, int s, int64_t* injected)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    long long n = 0;
    const int rc = species_source(c, s, &n);
    if (injected) *injected = n;
    return rc;
}

int mag2d_sort(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    return sort_species(c, s, true);
}

int mag2d_set_sort_interval(mag2d_ctx* c, int steps)
{
    CHECK_CTX(c);
    c->sort_interval = steps;
    return 0;
}

int mag2d_set_storage(mag2d_ctx* c, int storage)
{
    CHECK_CTX(c);
    if (storage != MAG2D_STORE_F64 && storage != MAG2D_STORE_F32) { mag2d_set_error("mag2d_set_storage: unknown storage type"); return 1; }
    if (storage == MAG2D_STORE_F32 && (is3d(c) || c->g.mover != MAG2D_ADVANCE_BORIS))
    {
        mag2d_set_error("mag2d_set_storage: fp32 particle storage is implemented for the 2-D Boris movers");
        return 1;
    }
    for (const SpeciesStore& S : c->sp)
        if (S.capacity > 0 || S.n_slots > 0) { mag2d_set_error("mag2d_set_storage: set the storage type before any particle is loaded"); return 1; }
    c->store_f32 = storage == MAG2D_STORE_F32;
    return 0;
}

int mag2d_set_store_layout(mag2d_ctx* c, int layout)
{
    CHECK_CTX(c);
    if (layout < MAG2D_LAYOUT_AUTO || layout > MAG2D_LAYOUT_BRICKS) { mag2d_set_error("mag2d_set_store_layout: unknown layout"); return 1; }
    c->store_layout = layout;
    return 0;
}

int mag2d_store_stats(mag2d_ctx* c, int s, int64_t* out8)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    SpeciesStore& S = c->sp[s];
    for (int q = 0; q < 8; q++) out8[q] = 0;
    out8[0] = S.rebinnings;
    if (S.d_ob_count)
    {
        unsigned h[4] = {0, 0, 0, 0};
        CUDA_OK(cudaMemcpyAsync(h, S.d_ob_count, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        out8[1] = h[1];
        out8[2] = h[2];
        out8[3] = h[0];
    }
    if (S.bins_valid && S.bin_nb > 0)
    {
        std::vector<unsigned> off((size_t)S.bin_nb + 1), cnt((size_t)S.bin_nb);
        CUDA_OK(cudaMemcpyAsync(off.data(), S.d_bin_off, sizeof(unsigned) * off.size(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaMemcpyAsync(cnt.data(), S.d_bin_cnt, sizeof(unsigned) * cnt.size(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
        long long worst = -1, room = 1LL << 40, used = 0;
        for (int b = 0; b < S.bin_nb; b++)
        {
            const long long r = (long long)(off[b + 1] - off[b]) - cnt[b];
            used += cnt[b];
            if (r < room) { room = r; worst = b; }
        }
        out8[4] = S.bin_nb;
        out8[5] = worst;
        out8[6] = room;          // free slots of the fullest bin
        out8[7] = used;          // slots in use (live particles + holes)
    }
    return 0;
}

int mag2d_set_species_sort_interval(mag2d_ctx* c, int s, int steps)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    c->sp[s].sort_interval = steps;
    return 0;
}

int mag2d_rho_reset(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    const size_t n = grid_n(c);
    if (!c->d_rho) return 0;
    if (s < 0) CUDA_OK(cudaMemsetAsync(c->d_rho, 0, sizeof(unsigned long long) * n * c->sp.size(), c->stream));
    else
    {
        CHECK_SPECIES(c, s);
        CUDA_OK(cudaMemsetAsync(c->d_rho + (size_t)s * n, 0, sizeof(unsigned long long) * n, c->stream));
    }
    return 0;
}

int mag2d_species_advance(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    c->edge_fields_fresh = c->edge_fields_fresh_armed = false;      // only mag2d_step shares the edge fields between species
    return advance_one(c, s, false);
}

int mag2d_species_advance_init(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    return launch_species_advance_init(c, s);
}

int mag2d_species_accumulate(mag2d_ctx* c, int s)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (is3d(c)) return launch_species_advance3d(c, s, true);
    return launch_species_accumulate(c, s);
}

int mag2d_rho_fixed_download(mag2d_ctx* c, int s, int64_t* rho_fixed)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    const size_t n = grid_n(c);
    CUDA_OK(cudaMemcpyAsync(rho_fixed, c->d_rho + (size_t)s * n, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_rho_upload(mag2d_ctx* c, int s, const int64_t* rho_fixed)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    const size_t n = grid_n(c);
    CUDA_OK(cudaMemcpyAsync(c->d_rho + (size_t)s * n, rho_fixed, sizeof(int64_t) * n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_rho_download(mag2d_ctx* c, double* rho)
{
    CHECK_CTX(c);
    if (c->sp.empty()) { mag2d_set_error("mag2d_rho_download: no species"); return 1; }
    // d_b is scratch between solves
    if (launch_rho_total(c, c->d_b)) return 1;
    CUDA_OK(cudaMemcpyAsync(rho, c->d_b, sizeof(double) * grid_n(c), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// Pic<D>::advance_init, src/pic.cpp:359-384
int mag2d_advance_init(mag2d_ctx* c)
{
    CHECK_CTX(c);
    if (is3d(c))
    {
        // the reference's 3-D species has no half-step-back; only the first charge deposit and field solve apply
        if (c->g.selfconsistent)
        {
            if (mag2d_rho_reset(c, -1)) return 1;
            for (size_t s = 0; s < c->sp.size(); s++)
                if (launch_species_advance3d(c, (int)s, true)) return 1;
            if (comm_allreduce_rho(c)) return 1;
        }
        return solve3d(c, nullptr);
    }
    if (c->g.selfconsistent)
    {
        if (mag2d_rho_reset(c, -1)) return 1;
        for (size_t s = 0; s < c->sp.size(); s++)
            if (launch_species_accumulate(c, (int)s)) return 1;
        if (comm_allreduce_rho(c)) return 1;
        if (mg_solve(c, 0, c->solve_tol, c->max_cycles, 0, nullptr, nullptr)) return 1;
        if (c->g.u_smooth && launch_u_smooth(c, 0, -1.0)) return 1;
        // the species grids are kept: they are the charge the first advance() solves with
    }
    for (size_t s = 0; s < c->sp.size(); s++)
        if (launch_species_advance_init(c, (int)s)) return 1;
    return 0;
}

// Pic<D>::advance, src/pic.cpp:330-358
int mag2d_step(mag2d_ctx* c, int nsteps)
{
    CHECK_CTX(c);
    if (c->sp.empty()) { mag2d_set_error("mag2d_step: no species"); return 1; }
    // N ranks, 3-D: between the steps of one call nobody but the slab-parallel solve reads the summed charge, and it reads only
    // this rank's planes: those steps reduce-scatter; the last step of the call all-reduces, so that every rank returns with
    // the complete grids the ABI promises (mag2d_rho_download, mag2d_rho_fixed_download)
    const bool own_slab = is3d(c) && c->g.selfconsistent && !c->use_source && solve3d_reads_own_planes_only(c);
    for (int it = 0; it < nsteps; it++)
    {
        if (c->timing) CUDA_OK(cudaEventRecord(c->ev[0], c->stream));
        if (is3d(c))
        {
            if (c->g.selfconsistent)
            {
                if (solve3d(c, nullptr)) return 1;
                if (mag2d_rho_reset(c, -1)) return 1;
            }
        }
        else if (c->g.selfconsistent)
        {
            // the direct solver needs no convergence test: never synchronise with the host inside the step
            const bool direct = c->direct.ok && c->solver_kind != MAG2D_SOLVER_MULTIGRID;
            if (mg_solve(c, 0, c->solve_tol, c->max_cycles, direct && !c->cycles_per_step ? 1 : c->cycles_per_step, nullptr, nullptr)) return 1;
            if (c->g.u_smooth && launch_u_smooth(c, 0, -1.0)) return 1;
            if (mag2d_rho_reset(c, -1)) return 1;
        }
        if (c->timing) CUDA_OK(cudaEventRecord(c->ev[1], c->stream));
        c->edge_fields_fresh = false;
        c->edge_fields_fresh_armed = !is3d(c);
        for (size_t s = 0; s < c->sp.size(); s++)
        {
            // the source appends to the store after every push: the pending cell counts of a fused sort would never survive, so these runs
            // use the stand-alone sort below (which also trims the slot range: the influx balances the wall losses)
            if (advance_one(c, (int)s, !c->use_source)) return 1;
            if (c->use_source && species_source(c, (int)s, nullptr)) return 1;      // pic.cpp:346-347
            // this species' charge grid is complete: its all-reduce runs on the side stream under the next species' push
            if (c->g.selfconsistent && comm_allreduce_species_async(c, (int)s, own_slab && it + 1 < nsteps)) return 1;
        }
        c->edge_fields_fresh = c->edge_fields_fresh_armed = false;
        if (c->timing) CUDA_OK(cudaEventRecord(c->ev[2], c->stream));
        if (c->g.selfconsistent && comm_allreduce_join(c)) return 1;
        if (c->timing) CUDA_OK(cudaEventRecord(c->ev[3], c->stream));
        // stand-alone sort: the multi-collision mover, or the fused sort switched off
        if (!c->fused_sort || c->g.mover == MAG2D_ADVANCE_MULTICOLL || c->use_source)
            for (size_t s = 0; s < c->sp.size(); s++)
            {
                const int K = effective_sort_interval(c, c->sp[s]);
                if (K > 0 && c->sp[s].n_slots > 0 && c->sp[s].steps_since_sort >= K)
                    if (sort_species(c, (int)s, c->use_source)) return 1;
            }
        if (c->timing)
        {
            CUDA_OK(cudaEventRecord(c->ev[4], c->stream));
            CUDA_OK(cudaEventSynchronize(c->ev[4]));
            float ms;
            CUDA_OK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->timers[1] += ms;
            CUDA_OK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->timers[0] += ms;
            CUDA_OK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3])); c->timers[3] += ms;
            CUDA_OK(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4])); c->timers[2] += ms;
            CUDA_OK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[4])); c->timers[4] += ms;
        }
    }
    return 0;
}

// Pic<D>::advance for a caller that keeps the particle arrays in HOST memory (the reference's own layout): every
// species' SoA arrays stream through a three-deep ring of device staging buffers, chunk by chunk — H2D copy of chunk
// i+1, the fused push/deposit kernel on chunk i and the D2H copy of chunk i-1 run concurrently on three streams, so
// the step costs max(upload, compute, download) instead of their sum, and the particle set is not limited by HBM.
// Blocks until the host arrays hold the pushed particles.  Removed particles come back with x = NaN.
static int step_streamed_impl(mag2d_ctx* c, int n_sp, const int32_t* species, const int64_t* n_slots, double* const* x, double* const* y,
                              double* const* z, double* const* vx, double* const* vy, double* const* vz, int64_t chunk_slots)
{
    const bool three_d = is3d(c);
    const int n_arr = three_d ? 6 : 5;
    if (c->store_f32) { mag2d_set_error("mag2d_step_streamed: fp64 particle storage only (the host arrays are double)"); return 1; }
    if (chunk_slots <= 0) chunk_slots = 1 << 22;
    chunk_slots = (chunk_slots + 1023) / 1024 * 1024;
    if (!c->s_h2d)
    {
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        for (int b = 0; b < 3; b++)
        {
            CUDA_OK(cudaEventCreateWithFlags(&c->ev_h2d[b], cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&c->ev_comp[b], cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&c->ev_d2h[b], cudaEventDisableTiming));
        }
    }
    if (c->chunk_capacity < chunk_slots)
    {
        CUDA_OK(cudaStreamSynchronize(c->s_d2h));
        for (int b = 0; b < 3; b++)
            for (int a = 0; a < n_arr; a++)
            {
                cudaFree(c->d_chunk[b][a]);
                CUDA_OK(cudaMalloc(&c->d_chunk[b][a], sizeof(double) * (size_t)chunk_slots));
            }
        c->chunk_capacity = chunk_slots;
    }
    if (three_d)
    {
        if (c->g.selfconsistent)
        {
            if (solve3d(c, nullptr)) return 1;
            if (mag2d_rho_reset(c, -1)) return 1;
        }
    }
    else if (c->g.selfconsistent)
    {
        const bool direct = c->direct.ok && c->solver_kind != MAG2D_SOLVER_MULTIGRID;
        if (mg_solve(c, 0, c->solve_tol, c->max_cycles, direct && !c->cycles_per_step ? 1 : c->cycles_per_step, nullptr, nullptr)) return 1;
        if (c->g.u_smooth && launch_u_smooth(c, 0, -1.0)) return 1;
        if (mag2d_rho_reset(c, -1)) return 1;
    }
    long long ring = 0;
    int rc = 0;
    for (int q = 0; q < n_sp && !rc; q++)
    {
        const int s = species[q];
        CHECK_SPECIES(c, s);
        SpeciesStore& S = c->sp[s];
        if (refresh_pools(c, s)) return 1;
        double* const host[6] = {x[q], z[q], vx[q], vy[q], vz[q], three_d ? y[q] : nullptr};      // staging order: y last
        // 2-D Cartesian push without a magnetic field: the push never touches the out-of-plane velocity (k_push_boris: need_vy), only
        // the collision pass does, for the ~1 % of particles whose collision test fired.  When the caller's vy array is pinned (device-
        // accessible) host memory, it is not copied at all: k_mcc_collide reads and writes those few elements in place over PCIe
        // (zero-copy), which takes a fifth off the bytes of the step.  Pageable arrays are staged like the others.
        double* vy_mapped = nullptr;
        if (!three_d && c->g.coord == MAG2D_CARTESIAN && c->g.magnetic_field_const && c->g.Br == 0.0 && c->g.Bz == 0.0 && c->g.Bt == 0.0 && n_slots[q] > 0 &&
            !(getenv("MAG2D_STREAM_VY") && atoi(getenv("MAG2D_STREAM_VY")) != 0))
        {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, vy[q]) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
                vy_mapped = static_cast<double*>(attr.devicePointer);
            else cudaGetLastError();
        }
        for (long long off = 0; off < n_slots[q] && !rc; off += chunk_slots, ring++)
        {
            const int b = (int)(ring % 3);
            const long long cnt = std::min<long long>(chunk_slots, n_slots[q] - off);
            CUDA_OK(cudaStreamWaitEvent(c->s_h2d, c->ev_d2h[b], 0));       // the buffer's previous tenant has left
            for (int a = 0; a < n_arr; a++)
            {
                if (a == 3 && vy_mapped) continue;
                CUDA_OK(cudaMemcpyAsync(c->d_chunk[b][a], host[a] + off, sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, c->s_h2d));
                c->streamed_h2d += sizeof(double) * (size_t)cnt;
            }
            CUDA_OK(cudaEventRecord(c->ev_h2d[b], c->s_h2d));
            CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_h2d[b], 0));
            ParticlesDev view;
            memset(&view, 0, sizeof(view));
            view.x = c->d_chunk[b][0]; view.z = c->d_chunk[b][1]; view.vx = c->d_chunk[b][2]; view.vy = c->d_chunk[b][3]; view.vz = c->d_chunk[b][4];
            if (vy_mapped) view.vy = vy_mapped + off;
            view.y = three_d ? c->d_chunk[b][5] : nullptr;
            view.n = cnt;
            c->chunk_view = &view;
            c->chunk_slot0 = off;
            rc = three_d ? launch_species_advance3d(c, s, false, 0) : launch_species_advance(c, s, 0);
            c->chunk_view = nullptr;
            if (rc) break;
            CUDA_OK(cudaEventRecord(c->ev_comp[b], c->stream));
            CUDA_OK(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[b], 0));
            for (int a = 0; a < n_arr; a++)
            {
                if (a == 3 && vy_mapped) continue;
                CUDA_OK(cudaMemcpyAsync(host[a] + off, c->d_chunk[b][a], sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToHost, c->s_d2h));
                c->streamed_d2h += sizeof(double) * (size_t)cnt;
            }
            CUDA_OK(cudaEventRecord(c->ev_d2h[b], c->s_d2h));
        }
        // Species<D>::advance: niter++, t += dt — once per step, not per chunk
        S.niter++;
        S.t += S.desc.dt;
    }
    if (rc) return rc;
    if (c->g.selfconsistent && comm_allreduce_rho(c)) return 1;
    CUDA_OK(cudaStreamSynchronize(c->s_d2h));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

int mag2d_step_streamed(mag2d_ctx* c, int n_sp, const int32_t* species, const int64_t* n_slots, double* const* x, double* const* z,
                        double* const* vx, double* const* vy, double* const* vz, int64_t chunk_slots)
{
    CHECK_CTX(c);
    if (is3d(c) || c->g.mover != MAG2D_ADVANCE_BORIS) { mag2d_set_error("mag2d_step_streamed: 2-D Boris movers only"); return 1; }
    return step_streamed_impl(c, n_sp, species, n_slots, x, nullptr, z, vx, vy, vz, chunk_slots);
}

int mag2d_step_streamed3(mag2d_ctx* c, int n_sp, const int32_t* species, const int64_t* n_slots, double* const* x, double* const* y,
                         double* const* z, double* const* vx, double* const* vy, double* const* vz, int64_t chunk_slots)
{
    CHECK_CTX(c);
    if (!is3d(c)) { mag2d_set_error("mag2d_step_streamed3: CARTESIAN3D only"); return 1; }
    return step_streamed_impl(c, n_sp, species, n_slots, x, y, z, vx, vy, vz, chunk_slots);
}

int mag2d_streamed_bytes(mag2d_ctx* c, int64_t* h2d, int64_t* d2h, int reset)
{
    if (!c) { mag2d_set_error("null context"); return 1; }
    if (h2d) *h2d = (int64_t)c->streamed_h2d;
    if (d2h) *d2h = (int64_t)c->streamed_d2h;
    if (reset) c->streamed_h2d = c->streamed_d2h = 0;
    return 0;
}

int mag2d_energy_hist(mag2d_ctx* c, int s, int nbins, double emax, double* hist, double* stats)
{
    CHECK_CTX(c);
    CHECK_SPECIES(c, s);
    if (nbins < 1 || nbins > 4096) { mag2d_set_error("mag2d_energy_hist: 1..4096 bins"); return 1; }
    return launch_energy_hist(c, s, nbins, emax, hist, stats);
}

int mag2d_kernel_launches(mag2d_ctx* c, int64_t* n)
{
    if (!c) { mag2d_set_error("null context"); return 1; }
    *n = c->launches;
    return 0;
}

int mag2d_set_timing(mag2d_ctx* c, int enable)
{
    CHECK_CTX(c);
    c->timing = enable != 0;
    for (double& t : c->timers) t = 0;
    return 0;
}

int mag2d_timers(mag2d_ctx* c, double* out5)
{
    CHECK_CTX(c);
    CUDA_OK(cudaStreamSynchronize(c->stream));
    for (int q = 0; q < 5; q++) out5[q] = c->timers[q];
    for (double& t : c->timers) t = 0;
    return 0;
}

int mag2d_device_pointer(mag2d_ctx* c, int what, void** ptr, size_t* bytes)
{
    CHECK_CTX(c);
    const size_t n = grid_n(c);
    switch (what)
    {
        case 0: *ptr = c->d_rho; *bytes = sizeof(unsigned long long) * n * c->sp.size(); break;
        case 1: *ptr = c->d_u; *bytes = sizeof(double) * n; break;
        case 2: *ptr = c->d_uRF; *bytes = sizeof(double) * n; break;
        default: mag2d_set_error("mag2d_device_pointer: unknown selector"); return 1;
    }
    return 0;
}

}  // extern "C"
