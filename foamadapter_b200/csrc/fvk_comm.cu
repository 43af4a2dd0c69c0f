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
// Peer-memory path (fvk_comm_p2p_export / fvk_comm_p2p_connect): every rank owns a WINDOW (cudaMalloc + CUDA IPC) that
// its peers map; the halo exchange becomes "gather my send cells and store them straight into the neighbour's window
// over NVLink, then raise a flag there" + "wait for my neighbours' flags, copy the window into the ghost range", and the
// all-reduce of the CG scalars is a mailbox exchange executed by the LAST BLOCK of the kernel that produced the partial
// sums (fvk_la.cu) -- no NCCL kernel, no separate launch, no host involvement, bit-identical sums on every rank (fixed
// rank order). NCCL stays the transport when the windows are not connected.
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

// ---- peer-memory kernels ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// gather the send cells and store them into the neighbours' windows; the last block to finish raises the flags
template <int NC>
__global__ void __launch_bounds__(256)
k_halo_push(const FvkP2PCtx* __restrict__ ctxp, const int* __restrict__ cells, const double* __restrict__ field)
{
    const FvkP2PCtx& ctx = *ctxp;
    const unsigned long long seq = ctx.state->haloSeq + 1; // only the last block advances it, after everyone read it
    const int nSend = ctx.sendOff[ctx.nNbr];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSend; i += gridDim.x * blockDim.x)
    {
        int k = 0;
        while (i >= ctx.sendOff[k + 1]) ++k;
        double* dst = reinterpret_cast<double*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_HALO_OFF)
                      + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.peerGhost[k] + size_t(NC) * (ctx.peerRecvOff[k] + (i - ctx.sendOff[k]));
        const int64_t c = cells[i];
#pragma unroll
        for (int q = 0; q < NC; ++q) dst[q] = field[NC * c + q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(&ctx.state->pushCounter, 1u);
        if (t == gridDim.x - 1)
        {
            __threadfence();
            ctx.state->pushCounter = 0u;
            for (int k = 0; k < ctx.nNbr; ++k)
                st_release_sys(reinterpret_cast<unsigned long long*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_HALOFLAG_OFF) + ctx.rank, seq);
            ctx.state->haloSeq = seq;
        }
    }
}

// wait for every neighbour's flag of the current exchange, then copy the window into the ghost range
template <int NC>
__global__ void __launch_bounds__(256)
k_halo_wait_unpack(const FvkP2PCtx* __restrict__ ctxp, double* __restrict__ field)
{
    const FvkP2PCtx& ctx = *ctxp;
    const unsigned long long seq = ctx.state->haloSeq;
    const unsigned long long t0 = (blockIdx.x == 0 && threadIdx.x == 0) ? fvk_gtime() : 0ull;
    if (threadIdx.x < ctx.nNbr)
    {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_HALOFLAG_OFF) + ctx.nbrRank[threadIdx.x];
        while (ld_acquire_sys(flag) < seq) __nanosleep(40);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctx.state->dbg[6] += fvk_gtime() - t0; ctx.state->dbg[7] += 1; } // time spent waiting for the neighbours
    const double* src = reinterpret_cast<const double*>(ctx.win[ctx.rank] + FVK_P2P_HALO_OFF) + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.nGhost;
    double* dst = field + size_t(NC) * ctx.nOwned;
    const int n = NC * ctx.nGhost;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldcg(src + i); // L1 may hold the previous exchange
}


// several cell fields in ONE exchange (one push, one flag, one wait): ghost cell g of neighbour k carries the fields'
// components back to back, TOT doubles per ghost
struct HaloFields
{
    int n, tot;
    double* f[4];
    int nc[4], off[4];
};
__global__ void __launch_bounds__(256)
k_halo_push_multi(const FvkP2PCtx* __restrict__ ctxp, const int* __restrict__ cells, HaloFields hf)
{
    const FvkP2PCtx& ctx = *ctxp;
    const unsigned long long seq = ctx.state->haloSeq + 1;
    const int nSend = ctx.sendOff[ctx.nNbr];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nSend; i += gridDim.x * blockDim.x)
    {
        int k = 0;
        while (i >= ctx.sendOff[k + 1]) ++k;
        double* dst = reinterpret_cast<double*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_HALO_OFF)
                      + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.peerGhost[k] + size_t(hf.tot) * (ctx.peerRecvOff[k] + (i - ctx.sendOff[k]));
        const int64_t c = cells[i];
        for (int a = 0; a < hf.n; ++a)
            for (int q = 0; q < hf.nc[a]; ++q) dst[hf.off[a] + q] = hf.f[a][hf.nc[a] * c + q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(&ctx.state->pushCounter, 1u);
        if (t == gridDim.x - 1)
        {
            __threadfence();
            ctx.state->pushCounter = 0u;
            for (int k = 0; k < ctx.nNbr; ++k)
                st_release_sys(reinterpret_cast<unsigned long long*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_HALOFLAG_OFF) + ctx.rank, seq);
            ctx.state->haloSeq = seq;
        }
    }
}
__global__ void __launch_bounds__(256)
k_halo_wait_unpack_multi(const FvkP2PCtx* __restrict__ ctxp, HaloFields hf)
{
    const FvkP2PCtx& ctx = *ctxp;
    const unsigned long long seq = ctx.state->haloSeq;
    const unsigned long long t0 = (blockIdx.x == 0 && threadIdx.x == 0) ? fvk_gtime() : 0ull;
    if (threadIdx.x < ctx.nNbr)
    {
        const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_HALOFLAG_OFF) + ctx.nbrRank[threadIdx.x];
        while (ld_acquire_sys(flag) < seq) __nanosleep(40);
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctx.state->dbg[6] += fvk_gtime() - t0; ctx.state->dbg[7] += 1; } // time spent waiting for the neighbours
    const double* src = reinterpret_cast<const double*>(ctx.win[ctx.rank] + FVK_P2P_HALO_OFF) + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.nGhost;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ctx.nGhost; g += gridDim.x * blockDim.x)
        for (int a = 0; a < hf.n; ++a)
            for (int q = 0; q < hf.nc[a]; ++q) hf.f[a][size_t(hf.nc[a]) * (ctx.nOwned + g) + q] = __ldcg(src + size_t(hf.tot) * g + hf.off[a] + q);
}

__global__ void k_allreduce_p2p(const FvkP2PCtx* __restrict__ ctx, double* data, int n)
{
    __shared__ double sh[4];
    if (threadIdx.x < n) sh[threadIdx.x] = data[threadIdx.x];
    __syncthreads();
    fvk_p2p_allreduce_sum(*ctx, sh, n);
    __syncthreads();
    if (threadIdx.x < n) data[threadIdx.x] = sh[threadIdx.x];
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
    // peer-memory windows
    char* window = nullptr;        // this rank's window (cudaMalloc)
    size_t windowBytes = 0;
    std::vector<char*> peerWin;    // [nRanks] mapped windows (own pointer at [rank]); empty until connected
    FvkP2PCtx* ctx_d = nullptr;    // device copy of the context the kernels read
    FvkP2PState* state_d = nullptr;
    bool p2p = false;
};

const FvkP2PCtx* fvk_comm_p2p_ctx(const fvk_comm* c) { return (c && c->p2p) ? c->ctx_d : nullptr; }

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
    for (size_t r = 0; r < c->peerWin.size(); ++r)
        if (c->peerWin[r] && int(r) != c->rank) cudaIpcCloseMemHandle(c->peerWin[r]);
    if (c->window) cudaFree(c->window);
    if (c->ctx_d) cudaFree(c->ctx_d);
    if (c->state_d) cudaFree(c->state_d);
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
    c->p2p = false; // the windows are laid out for one halo plan: export + connect again
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
    if (c->p2p)
    {
        const int nS = c->sendOff.back(), nG = c->recvOff.back();
        const int gp = nS > 0 ? ((nS + 255) / 256 < 64 ? (nS + 255) / 256 : 64) : 1;
        const int gw = nG > 0 ? ((nG * ncomp + 255) / 256 < 64 ? (nG * ncomp + 255) / 256 : 64) : 1;
        if (ncomp == 1)
        {
            k_halo_push<1><<<gp, 256, 0, st>>>(c->ctx_d, c->sendCells, field);
            k_halo_wait_unpack<1><<<gw, 256, 0, st>>>(c->ctx_d, field);
        }
        else
        {
            k_halo_push<3><<<gp, 256, 0, st>>>(c->ctx_d, c->sendCells, field);
            k_halo_wait_unpack<3><<<gw, 256, 0, st>>>(c->ctx_d, field);
        }
        FVK_LAUNCH_CHECK();
        return FVK_OK;
    }
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
    if (c->p2p && op == kNcclSum && count <= 4)
    {
        k_allreduce_p2p<<<1, 32, 0, st>>>(c->ctx_d, data, count);
        FVK_LAUNCH_CHECK();
        return FVK_OK;
    }
    FVK_NCCL(nccl().AllReduce(data, data, size_t(count), kNcclFloat64, op, c->nc, st));
    return FVK_OK;
}
int fvk_comm_allreduce_sum_impl(fvk_comm* c, double* data, int count, cudaStream_t st) { return allreduce(c, data, count, kNcclSum, st); }

extern "C" int fvk_comm_halo_exchange(fvk_comm* c, double* field, int ncomp, fvk_stream s)
{
    return fvk_comm_halo_exchange_impl(c, field, ncomp, fvk_cu(s));
}
extern "C" int fvk_comm_halo_exchange_multi(fvk_comm* c, int nFields, const fvk_halo_field* fields_h, fvk_stream s)
{
    if (!c || nFields < 1 || nFields > 4 || !fields_h) return fvk_fail(FVK_EINVAL, "fvk_comm_halo_exchange_multi: bad argument");
    HaloFields hf;
    hf.n = nFields; hf.tot = 0;
    for (int a = 0; a < nFields; ++a)
    {
        if (!fields_h[a].field || (fields_h[a].ncomp != 1 && fields_h[a].ncomp != 3)) return fvk_fail(FVK_EINVAL, "fvk_comm_halo_exchange_multi: bad field %d", a);
        hf.f[a] = fields_h[a].field; hf.nc[a] = fields_h[a].ncomp; hf.off[a] = hf.tot; hf.tot += fields_h[a].ncomp;
    }
    if (hf.tot > FVK_P2P_HALO_COMPS) return fvk_fail(FVK_EINVAL, "fvk_comm_halo_exchange_multi: more than %d components", FVK_P2P_HALO_COMPS);
    if (c->nRanks == 1 || c->nbrRank.empty()) return FVK_OK;
    cudaStream_t st = fvk_cu(s);
    if (!c->p2p || nFields == 1)
    { // NCCL transport: one send/recv group per field
        for (int a = 0; a < nFields; ++a)
            if (int rc = fvk_comm_halo_exchange_impl(c, fields_h[a].field, fields_h[a].ncomp, st)) return rc;
        return FVK_OK;
    }
    const int nS = c->sendOff.back(), nG = c->recvOff.back();
    const int gp = nS > 0 ? ((nS + 255) / 256 < 64 ? (nS + 255) / 256 : 64) : 1;
    const int gw = nG > 0 ? ((nG + 255) / 256 < 64 ? (nG + 255) / 256 : 64) : 1;
    k_halo_push_multi<<<gp, 256, 0, st>>>(c->ctx_d, c->sendCells, hf);
    k_halo_wait_unpack_multi<<<gw, 256, 0, st>>>(c->ctx_d, hf);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_comm_allreduce_sum(fvk_comm* c, double* data_d, int count, fvk_stream s)
{
    return allreduce(c, data_d, count, kNcclSum, fvk_cu(s));
}
extern "C" int fvk_comm_allreduce_max(fvk_comm* c, double* data_d, int count, fvk_stream s)
{
    return allreduce(c, data_d, count, kNcclMax, fvk_cu(s));
}

// ---- peer-memory windows ---------------------------------------------------------------------------------------------
namespace
{
struct P2PBlob // what a rank publishes (FVK_P2P_BLOB_BYTES)
{
    cudaIpcMemHandle_t handle; // 64 bytes
    int32_t nGhost;
    int32_t nOwned;
    int32_t nNbr;
    int32_t nbrRank[FVK_P2P_MAX_NBR];
    int32_t recvOff[FVK_P2P_MAX_NBR]; // offset (cells) of neighbour k's data in this rank's ghost range
};
static_assert(sizeof(P2PBlob) <= FVK_P2P_BLOB_BYTES, "blob too large");
} // namespace

extern "C" int fvk_comm_p2p_export(fvk_comm* c, void* blob)
{
    if (!c || !blob) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_export: null");
    if (c->nRanks > FVK_P2P_MAX_RANKS || int(c->nbrRank.size()) > FVK_P2P_MAX_NBR)
        return fvk_fail(FVK_EUNSUPPORTED, "fvk_comm_p2p_export: more than %d ranks / %d neighbours", FVK_P2P_MAX_RANKS, FVK_P2P_MAX_NBR);
    c->p2p = false;
    for (size_t r = 0; r < c->peerWin.size(); ++r)
        if (c->peerWin[r] && int(r) != c->rank) cudaIpcCloseMemHandle(c->peerWin[r]);
    c->peerWin.clear();
    if (c->window) { cudaFree(c->window); c->window = nullptr; }
    const size_t nGhost = size_t(c->recvOff.empty() ? 0 : c->recvOff.back());
    c->windowBytes = FVK_P2P_Z_OFF(nGhost) + sizeof(double) * 2 * (size_t(c->nOwned) + nGhost);
    FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->window), c->windowBytes));
    FVK_CUDA(cudaMemset(c->window, 0, c->windowBytes));
    FVK_CUDA(cudaDeviceSynchronize());
    P2PBlob b;
    std::memset(&b, 0, sizeof(b));
    FVK_CUDA(cudaIpcGetMemHandle(&b.handle, c->window));
    b.nGhost = int32_t(nGhost);
    b.nOwned = c->nOwned;
    b.nNbr = int32_t(c->nbrRank.size());
    for (int k = 0; k < b.nNbr; ++k) { b.nbrRank[k] = c->nbrRank[k]; b.recvOff[k] = c->recvOff[k]; }
    std::memset(blob, 0, FVK_P2P_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof(b));
    return FVK_OK;
}

extern "C" int fvk_comm_p2p_connect(fvk_comm* c, const void* blobs)
{
    if (!c || !blobs) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_connect: null");
    if (!c->window) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_connect: call fvk_comm_p2p_export first");
    const char* bl = static_cast<const char*>(blobs);
    c->peerWin.assign(size_t(c->nRanks), nullptr);
    FvkP2PCtx ctx;
    std::memset(&ctx, 0, sizeof(ctx));
    ctx.rank = c->rank; ctx.nRanks = c->nRanks; ctx.nNbr = int(c->nbrRank.size());
    ctx.nOwned = c->nOwned; ctx.nGhost = c->recvOff.empty() ? 0 : c->recvOff.back();
    for (int r = 0; r < c->nRanks; ++r)
    {
        P2PBlob b;
        std::memcpy(&b, bl + size_t(r) * FVK_P2P_BLOB_BYTES, sizeof(b));
        if (r == c->rank)
            c->peerWin[r] = c->window;
        else
        {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, b.handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                cudaGetLastError();
                return fvk_fail(FVK_ECUDA, "fvk_comm_p2p_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            }
            c->peerWin[r] = static_cast<char*>(p);
        }
        ctx.win[r] = c->peerWin[r];
    }
    for (int k = 0; k < ctx.nNbr; ++k)
    {
        const int r = c->nbrRank[k];
        P2PBlob b;
        std::memcpy(&b, bl + size_t(r) * FVK_P2P_BLOB_BYTES, sizeof(b));
        int kk = -1;
        for (int j = 0; j < b.nNbr; ++j)
            if (b.nbrRank[j] == c->rank) kk = j;
        if (kk < 0) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_connect: rank %d does not list rank %d as a neighbour", r, c->rank);
        ctx.nbrRank[k] = r; ctx.peerRecvOff[k] = b.recvOff[kk]; ctx.peerGhost[k] = b.nGhost; ctx.peerOwned[k] = b.nOwned;
        ctx.sendOff[k] = c->sendOff[k];
    }
    ctx.sendOff[ctx.nNbr] = c->sendOff.empty() ? 0 : c->sendOff.back();
    if (!c->state_d)
    {
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->state_d), sizeof(FvkP2PState)));
        FVK_CUDA(cudaMemset(c->state_d, 0, sizeof(FvkP2PState)));
    }
    ctx.state = c->state_d;
    ctx.sendCells = c->sendCells;
    ctx.zwin = reinterpret_cast<double*>(c->window + FVK_P2P_Z_OFF(ctx.nGhost));
    if (!c->ctx_d) FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->ctx_d), sizeof(FvkP2PCtx)));
    FVK_CUDA(cudaMemcpy(c->ctx_d, &ctx, sizeof(ctx), cudaMemcpyHostToDevice));
    FVK_CUDA(cudaDeviceSynchronize());
    c->p2p = true;
    return FVK_OK;
}

extern "C" int fvk_comm_p2p_enabled(const fvk_comm* c) { return (c && c->p2p) ? 1 : 0; }
/* back to the NCCL transport (e.g. when another rank could not map the windows: the choice must be collective) */
extern "C" int fvk_comm_p2p_disable(fvk_comm* c)
{
    if (!c) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_disable: null");
    c->p2p = false;
    return FVK_OK;
}

/* accumulated nanoseconds of the in-kernel communication phases (diagnostics): [0] flag raise + system fence, [1] all-reduce
 * (r.z, r.r), [2] wait for the halo flags, [3] count; [4] all-reduce p.q, [5] count; [6] field exchanges: wait for the neighbours'
 * flags (block 0), [7] count */
extern "C" int fvk_comm_p2p_debug(const fvk_comm* c, uint64_t* out8)
{
    if (!c || !out8 || !c->state_d) return fvk_fail(FVK_EINVAL, "fvk_comm_p2p_debug: not connected");
    FvkP2PState s;
    FVK_CUDA(cudaMemcpy(&s, c->state_d, sizeof(s), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out8[k] = s.dbg[k];
    return FVK_OK;
}
