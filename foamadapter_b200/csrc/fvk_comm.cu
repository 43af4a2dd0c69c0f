// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no distributed finite-volume path: its MPI layer
// (src/NeoN/include/NeoN/core/mpi/{operators,halfDuplexCommBuffer,fullDuplexCommBuffer}.hpp,
// src/NeoN/include/NeoN/mesh/unstructured/communicator.hpp:89-143) packs `field[sendMap[rank][i]]`
// per neighbour rank on the HOST, Isend/Irecv's char buffers and unpacks on the host, and is not
// called by any operator. This is its device-resident replacement with the same protocol shape
// (start -> interior work -> finish): ghost cells live directly behind the owned cells of every
// field, a pack kernel gathers the send cells into one device buffer, and a single NCCL group of
// ncclSend/ncclRecv pairs delivers every neighbour's slice straight into the ghost range -- no unpack,
// no host staging. Global reductions are ncclAllReduce on 1-4 doubles that stay on the device.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2") so the library has no link-time dependency:
// inside a torch process this binds to the NCCL torch already loaded, otherwise to the system one.
#include "fvk_device.cuh"

#include <dlfcn.h>

#include <cstring>
#include <vector>

namespace
{
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8;

struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"})
    {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
#define SYM(field, name) *reinterpret_cast<void**>(&api.field) = dlsym(api.handle, name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv
             && api.GroupStart && api.GroupEnd;
    return api;
}

#define FVK_NCCL(call)                                                                             \
    do                                                                                             \
    {                                                                                              \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != 0)                                                                               \
            return fvk_fail(FVK_ENCCL, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
                            nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error");     \
    } while (0)

template <int NC>
__global__ void __launch_bounds__(256)
k_pack(int n, const int* __restrict__ cells, const double* __restrict__ field, double* __restrict__ buf)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int64_t c = cells[i];
#pragma unroll
        for (int k = 0; k < NC; ++k) buf[int64_t(NC) * i + k] = field[NC * c + k];
    }
}
} // namespace

struct fvk_comm
{
    int rank = 0, nRanks = 1;
    ncclComm_t nc = nullptr;
    // halo plan
    int32_t nOwned = 0;
    std::vector<int32_t> nbrRank, sendOff, recvOff; // host
    int32_t* sendCells = nullptr;                   // device [sendOff.back()]
    double* sendBuf = nullptr;                      // device [3 * sendOff.back()]
};

extern "C" int fvk_comm_unique_id(void* id128)
{
    if (!id128) return fvk_fail(FVK_EINVAL, "fvk_comm_unique_id: null");
    if (!nccl().ok) return fvk_fail(FVK_ENCCL, "fvk_comm_unique_id: NCCL (libnccl.so.2) not available");
    ncclUniqueId id;
    FVK_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return FVK_OK;
}

extern "C" int fvk_comm_create(int rank, int nRanks, const void* id128, fvk_comm** out)
{
    if (!out || nRanks < 1 || rank < 0 || rank >= nRanks || (nRanks > 1 && !id128))
        return fvk_fail(FVK_EINVAL, "fvk_comm_create: bad argument");
    *out = nullptr;
    fvk_comm* c = new fvk_comm;
    c->rank = rank; c->nRanks = nRanks;
    if (nRanks > 1)
    {
        if (!nccl().ok) { delete c; return fvk_fail(FVK_ENCCL, "fvk_comm_create: NCCL (libnccl.so.2) not available"); }
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        ncclResult_t r = nccl().CommInitRank(&c->nc, nRanks, id, rank);
        if (r != 0) { delete c; return fvk_fail(FVK_ENCCL, "ncclCommInitRank failed (%d)", r); }
    }
    *out = c;
    return FVK_OK;
}

extern "C" int fvk_comm_destroy(fvk_comm* c)
{
    if (!c) return FVK_OK;
    if (c->sendCells) cudaFree(c->sendCells);
    if (c->sendBuf) cudaFree(c->sendBuf);
    if (c->nc) nccl().CommDestroy(c->nc);
    delete c;
    return FVK_OK;
}

extern "C" int fvk_comm_rank(const fvk_comm* c, int* rank, int* nRanks)
{
    if (!c) return fvk_fail(FVK_EINVAL, "fvk_comm_rank: null");
    if (rank) *rank = c->rank;
    if (nRanks) *nRanks = c->nRanks;
    return FVK_OK;
}

extern "C" int fvk_comm_set_halo(fvk_comm* c, int32_t nOwned, int32_t nNeighbours, const int32_t* neighbourRanks_h,
                                 const int32_t* sendOffsets_h, const int32_t* sendCells_h, const int32_t* recvOffsets_h)
{
    if (!c || nOwned < 0 || nNeighbours < 0 || (nNeighbours && (!neighbourRanks_h || !sendOffsets_h || !sendCells_h || !recvOffsets_h)))
        return fvk_fail(FVK_EINVAL, "fvk_comm_set_halo: bad argument");
    c->nOwned = nOwned;
    c->nbrRank.assign(neighbourRanks_h, neighbourRanks_h + nNeighbours);
    c->sendOff.assign(1, 0); c->recvOff.assign(1, 0);
    if (nNeighbours)
    {
        c->sendOff.assign(sendOffsets_h, sendOffsets_h + nNeighbours + 1);
        c->recvOff.assign(recvOffsets_h, recvOffsets_h + nNeighbours + 1);
    }
    for (int k = 0; k < nNeighbours; ++k)
        if (c->nbrRank[k] < 0 || c->nbrRank[k] >= c->nRanks || c->nbrRank[k] == c->rank || c->sendOff[k + 1] < c->sendOff[k]
            || c->recvOff[k + 1] < c->recvOff[k])
            return fvk_fail(FVK_EINVAL, "fvk_comm_set_halo: bad neighbour %d", k);
    if (c->sendCells) { cudaFree(c->sendCells); c->sendCells = nullptr; }
    if (c->sendBuf) { cudaFree(c->sendBuf); c->sendBuf = nullptr; }
    const size_t nSend = size_t(c->sendOff.back());
    for (size_t i = 0; i < nSend; ++i)
        if (sendCells_h[i] < 0 || sendCells_h[i] >= nOwned) return fvk_fail(FVK_EINVAL, "fvk_comm_set_halo: send cell %zu not owned", i);
    if (nSend)
    {
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->sendCells), sizeof(int32_t) * nSend));
        FVK_CUDA(cudaMemcpy(c->sendCells, sendCells_h, sizeof(int32_t) * nSend, cudaMemcpyHostToDevice));
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->sendBuf), sizeof(double) * 3 * nSend));
    }
    return FVK_OK;
}

int fvk_comm_halo_exchange_impl(fvk_comm* c, double* field, int ncomp, cudaStream_t st)
{
    if (!c || !field || (ncomp != 1 && ncomp != 3)) return fvk_fail(FVK_EINVAL, "fvk_comm_halo_exchange: bad argument");
    if (c->nRanks == 1 || c->nbrRank.empty()) return FVK_OK;
    const int nSend = c->sendOff.back();
    if (nSend)
    {
        const int grid = (nSend + 255) / 256 < 1184 ? (nSend + 255) / 256 : 1184;
        if (ncomp == 1) k_pack<1><<<grid, 256, 0, st>>>(nSend, c->sendCells, field, c->sendBuf);
        else k_pack<3><<<grid, 256, 0, st>>>(nSend, c->sendCells, field, c->sendBuf);
        FVK_LAUNCH_CHECK();
    }
    FVK_NCCL(nccl().GroupStart());
    for (size_t k = 0; k < c->nbrRank.size(); ++k)
    {
        const size_t ns = size_t(c->sendOff[k + 1] - c->sendOff[k]) * ncomp;
        const size_t nr = size_t(c->recvOff[k + 1] - c->recvOff[k]) * ncomp;
        if (ns) FVK_NCCL(nccl().Send(c->sendBuf + size_t(c->sendOff[k]) * ncomp, ns, kNcclFloat64, c->nbrRank[k], c->nc, st));
        if (nr) FVK_NCCL(nccl().Recv(field + (size_t(c->nOwned) + c->recvOff[k]) * ncomp, nr, kNcclFloat64, c->nbrRank[k], c->nc, st));
    }
    FVK_NCCL(nccl().GroupEnd());
    return FVK_OK;
}

static int allreduce(fvk_comm* c, double* data, int count, int op, cudaStream_t st)
{
    if (!c || !data || count < 1) return fvk_fail(FVK_EINVAL, "fvk_comm_allreduce: bad argument");
    if (c->nRanks == 1) return FVK_OK;
    FVK_NCCL(nccl().AllReduce(data, data, size_t(count), kNcclFloat64, op, c->nc, st));
    return FVK_OK;
}
int fvk_comm_allreduce_sum_impl(fvk_comm* c, double* data, int count, cudaStream_t st) { return allreduce(c, data, count, kNcclSum, st); }

extern "C" int fvk_comm_halo_exchange(fvk_comm* c, double* field, int ncomp, fvk_stream s)
{
    return fvk_comm_halo_exchange_impl(c, field, ncomp, fvk_cu(s));
}
extern "C" int fvk_comm_allreduce_sum(fvk_comm* c, double* data_d, int count, fvk_stream s)
{
    return allreduce(c, data_d, count, kNcclSum, fvk_cu(s));
}
extern "C" int fvk_comm_allreduce_max(fvk_comm* c, double* data_d, int count, fvk_stream s)
{
    return allreduce(c, data_d, count, kNcclMax, fvk_cu(s));
}
