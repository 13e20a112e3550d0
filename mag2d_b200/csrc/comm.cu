// comm.cu — multi-GPU plumbing: one process per GPU, particles sharded, charge grid all-reduced.
//
// The reference is a single-process code (no MPI/NCCL anywhere); this is new.  Every rank pushes its
// own shard of each species and deposits into its own fixed-point grid; one ncclAllReduce(int64, sum)
// per step over NVLink/NVSwitch makes every rank hold the global grid, after which the field solve
// is replicated.  Integer addition is associative, so the reduced grid — and everything downstream —
// is bit-identical for 1, 2, 4 or 8 GPUs and for any reduction order.
//
// NCCL is bound at run time with dlopen so that the library loads on machines without NCCL and so
// that, inside a torch process, the already-loaded libnccl.so.2 of the torch wheel is reused.
#include <dlfcn.h>

#include "ctx.hpp"

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_Send)(const void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_Recv)(void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_Group)();
typedef int (*fn_Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef int (*fn_ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi
{
    void* handle = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
    fn_Send Send = nullptr;
    fn_Recv Recv = nullptr;
    fn_Group GroupStart = nullptr, GroupEnd = nullptr;
    fn_Broadcast Broadcast = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_ReduceScatter ReduceScatter = nullptr;
};

NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
    {
        h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // reuse the copy torch already loaded
        if (h) break;
    }
    if (!h)
        for (const char* n : names)
        {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
    if (!h)
    {
        mag2d_set_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        return 1;
    }
    g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (fn_AllReduce)dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (fn_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (fn_GetErrorString)dlsym(h, "ncclGetErrorString");
    g_nccl.Send = (fn_Send)dlsym(h, "ncclSend");
    g_nccl.Recv = (fn_Recv)dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (fn_Group)dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (fn_Group)dlsym(h, "ncclGroupEnd");
    g_nccl.Broadcast = (fn_Broadcast)dlsym(h, "ncclBroadcast");
    g_nccl.AllGather = (fn_AllGather)dlsym(h, "ncclAllGather");
    g_nccl.ReduceScatter = (fn_ReduceScatter)dlsym(h, "ncclReduceScatter");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    {
        mag2d_set_error("libnccl.so.2 lacks the expected symbols");
        return 1;
    }
    g_nccl.handle = h;
    return 0;
}

int nccl_fail(const char* what, int rc)
{
    mag2d_set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error"));
    return 1;
}

}  // namespace

extern "C" int mag2d_comm_unique_id(void* id128)
{
    if (load_nccl()) return 1;
    NcclUniqueId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int mag2d_comm_init(mag2d_ctx* c, int rank, int nranks, const void* id128)
{
    if (load_nccl()) return 1;
    CUDA_OK(cudaSetDevice(c->device));
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm comm = nullptr;
    const int rc = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (rc) return nccl_fail("ncclCommInitRank", rc);
    c->nccl_comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    return 0;
}

extern "C" int mag2d_comm_destroy(mag2d_ctx* c)
{
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((NcclComm)c->nccl_comm);
    c->nccl_comm = nullptr;
    c->nranks = 1;
    c->rank = 0;
    return 0;
}

// Species s has finished depositing (everything enqueued on c->stream so far): sum its grid over the ranks on the side stream,
// so that the transfer runs under the NEXT species' push instead of after the last one.  comm_allreduce_join makes the
// compute stream wait for all of them.  Every rank issues the same sequence of collectives (species order), as NCCL requires.
// own_slab_only: the only reader of the sum will be the slab-parallel 3-D solve, which takes rank r's block of M / nranks x planes
// from rank r: an in-place reduce-scatter (half the traffic of the all-reduce) leaves exactly that block summed.
int comm_allreduce_species_async(mag2d_ctx* c, int s, bool own_slab_only)
{
    if (!c->nccl_comm || c->nranks <= 1 || c->sp[s].desc.charge == 0.0) return 0;
    if (!c->s_comm)
    {
        CUDA_OK(cudaStreamCreateWithFlags(&c->s_comm, cudaStreamNonBlocking));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_comm_in, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&c->ev_comm_out, cudaEventDisableTiming));
    }
    CUDA_OK(cudaEventRecord(c->ev_comm_in, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->s_comm, c->ev_comm_in, 0));
    const size_t n = grid_nodes(c);
    const int ncclInt64 = 4, ncclSum = 0;
    unsigned long long* buf = c->d_rho + (size_t)s * n;
    if (own_slab_only && g_nccl.ReduceScatter && n % (size_t)c->nranks == 0)
    {
        const size_t per = n / (size_t)c->nranks;
        const int rc = g_nccl.ReduceScatter(buf, buf + (size_t)c->rank * per, per, ncclInt64, ncclSum, (NcclComm)c->nccl_comm, c->s_comm);
        if (rc) return nccl_fail("ncclReduceScatter", rc);
    }
    else
    {
        const int rc = g_nccl.AllReduce(buf, buf, n, ncclInt64, ncclSum, (NcclComm)c->nccl_comm, c->s_comm);
        if (rc) return nccl_fail("ncclAllReduce", rc);
    }
    c->comm_pending = true;
    return 0;
}

int comm_allreduce_join(mag2d_ctx* c)
{
    if (!c->comm_pending) return 0;
    CUDA_OK(cudaEventRecord(c->ev_comm_out, c->s_comm));
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_comm_out, 0));
    c->comm_pending = false;
    return 0;
}

// in-place sum of the fixed-point charge grids of all species over all ranks
int comm_allreduce_rho(mag2d_ctx* c)
{
    if (!c->nccl_comm || c->nranks <= 1) return 0;
    // neutral species never deposit: reduce the contiguous range of charged species only
    int first = -1, last = -1;
    for (int s = 0; s < (int)c->sp.size(); s++)
        if (c->sp[s].desc.charge != 0.0)
        {
            if (first < 0) first = s;
            last = s;
        }
    if (first < 0) return 0;
    const size_t n = grid_nodes(c);
    const size_t count = (size_t)(last - first + 1) * n;
    const int ncclInt64 = 4, ncclSum = 0;
    unsigned long long* buf = c->d_rho + (size_t)first * n;
    const int rc = g_nccl.AllReduce(buf, buf, count, ncclInt64, ncclSum, (NcclComm)c->nccl_comm, c->stream);
    if (rc) return nccl_fail("ncclAllReduce", rc);
    return 0;
}

// ---- point-to-point pieces of the slab-parallel 3-D solve (poisson3d.cu); doubles, on the compute stream -----------------
bool comm_has_p2p() { return g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.Broadcast && g_nccl.AllGather; }
int comm_group_start()
{
    const int rc = g_nccl.GroupStart();
    return rc ? nccl_fail("ncclGroupStart", rc) : 0;
}
int comm_group_end()
{
    const int rc = g_nccl.GroupEnd();
    return rc ? nccl_fail("ncclGroupEnd", rc) : 0;
}
int comm_send(mag2d_ctx* c, const double* buf, size_t count, int peer)
{
    const int rc = g_nccl.Send(buf, count, 8 /* ncclFloat64 */, peer, (NcclComm)c->nccl_comm, c->stream);
    return rc ? nccl_fail("ncclSend", rc) : 0;
}
int comm_recv(mag2d_ctx* c, double* buf, size_t count, int peer)
{
    const int rc = g_nccl.Recv(buf, count, 8, peer, (NcclComm)c->nccl_comm, c->stream);
    return rc ? nccl_fail("ncclRecv", rc) : 0;
}
int comm_broadcast(mag2d_ctx* c, double* buf, size_t count, int root)
{
    const int rc = g_nccl.Broadcast(buf, buf, count, 8, root, (NcclComm)c->nccl_comm, c->stream);
    return rc ? nccl_fail("ncclBroadcast", rc) : 0;
}
// in-place all-gather: rank r's block of `count` doubles sits at buf + r * count
int comm_allgather_inplace(mag2d_ctx* c, double* buf, size_t count)
{
    const int rc = g_nccl.AllGather(buf + (size_t)c->rank * count, buf, count, 8, (NcclComm)c->nccl_comm, c->stream);
    return rc ? nccl_fail("ncclAllGather", rc) : 0;
}
