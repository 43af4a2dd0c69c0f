// Device-side helpers shared by the kernel translation units.
#pragma once
#include "fvk_internal.hpp"

#include <cuda_runtime.h>

#define FVK_CUDA(call)                                                                             \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fvk_fail(                                                                       \
                (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? FVK_ENODEVICE     \
                                                                               : FVK_ECUDA,        \
                "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));             \
    } while (0)

#define FVK_LAUNCH_CHECK() FVK_CUDA(cudaGetLastError())

static inline cudaStream_t fvk_cu(fvk_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached); B200 = 148
int fvk_sm_count();
// current experiment variant (fvk_set_variant)
int fvk_variant();
bool fvk_brick_config(int cfg[3]); // true: {cells per thread, threads per block, resident blocks} override is set

struct Vec3d
{
    double x, y, z;
};

__device__ __forceinline__ Vec3d ld3(const double* __restrict__ p, int64_t i)
{
    return Vec3d {p[3 * i], p[3 * i + 1], p[3 * i + 2]};
}
__device__ __forceinline__ void st3(double* __restrict__ p, int64_t i, Vec3d v)
{
    p[3 * i] = v.x;
    p[3 * i + 1] = v.y;
    p[3 * i + 2] = v.z;
}

// streaming (read-once) loads: bypass L1 allocation so gathered data keeps the L1
__device__ __forceinline__ double ld_stream(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- peer-memory windows (fvk_comm.cu; used by the CG kernels of fvk_la.cu) ----------------------------------------
// Window layout (bytes): [0, 512) halo flags (u64 per sender rank) | [1024, ...) all-reduce flags u64[2][64] |
// [4096, ...) all-reduce values double[2][64][4] | [FVK_P2P_HALO_OFF, ...) halo data double[2][3 * nGhost]
#define FVK_P2P_MAX_RANKS 64
#define FVK_P2P_MAX_NBR 32
#define FVK_P2P_HALOFLAG_OFF 0
#define FVK_P2P_ARFLAG_OFF 1024
#define FVK_P2P_ARVAL_OFF 4096
#define FVK_P2P_HALO_OFF 16384
struct FvkP2PState // device memory of the owning rank only
{
    unsigned long long haloSeq, arSeq;
    unsigned pushCounter, pad;
};
struct FvkP2PCtx
{
    int rank, nRanks, nNbr, nOwned, nGhost;
    char* win[FVK_P2P_MAX_RANKS];         // every rank's window as mapped into this process
    int nbrRank[FVK_P2P_MAX_NBR], sendOff[FVK_P2P_MAX_NBR + 1];
    int peerRecvOff[FVK_P2P_MAX_NBR];      // where my cells start in neighbour k's ghost range
    int peerGhost[FVK_P2P_MAX_NBR];        // neighbour k's ghost count (parity stride of its halo area)
    FvkP2PState* state;
};
struct fvk_comm;
const FvkP2PCtx* fvk_comm_p2p_ctx(const fvk_comm* c); // nullptr: windows not connected (NCCL transport)

#ifdef __CUDACC__
// Sum of n <= 4 doubles over all ranks, executed by (at least) the first warp of ONE block per rank; vals (shared or
// global memory visible to the block) is replaced by the total. Every rank adds the contributions in rank order, so
// the result is bit-identical everywhere. Mailboxes are double-buffered by the parity of the sequence number: a rank
// can only be one all-reduce ahead of the slowest one.
__device__ __forceinline__ void fvk_p2p_allreduce_sum(const FvkP2PCtx& ctx, double* vals, int n)
{
    if (threadIdx.x >= 32) return;
    const unsigned long long seq = ctx.state->arSeq + 1;
    const int par = int(seq & 1);
    double mine[4];
    for (int q = 0; q < n; ++q) mine[q] = vals[q];
    for (int r = threadIdx.x; r < ctx.nRanks; r += 32)
    {
        double* dv = reinterpret_cast<double*>(ctx.win[r] + FVK_P2P_ARVAL_OFF) + (size_t(par) * FVK_P2P_MAX_RANKS + ctx.rank) * 4;
        for (int q = 0; q < n; ++q) dv[q] = mine[q];
        __threadfence_system();
        unsigned long long* df = reinterpret_cast<unsigned long long*>(ctx.win[r] + FVK_P2P_ARFLAG_OFF) + size_t(par) * FVK_P2P_MAX_RANKS + ctx.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(df), "l"(seq) : "memory");
    }
    for (int r = threadIdx.x; r < ctx.nRanks; r += 32)
    {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_ARFLAG_OFF) + size_t(par) * FVK_P2P_MAX_RANKS + r;
        unsigned long long v;
        do
        {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        } while (v < seq);
    }
    __syncwarp();
    if (threadIdx.x == 0)
    {
        const double* mv = reinterpret_cast<const double*>(ctx.win[ctx.rank] + FVK_P2P_ARVAL_OFF) + size_t(par) * FVK_P2P_MAX_RANKS * 4;
        for (int q = 0; q < n; ++q)
        {
            double s = 0.0;
            for (int r = 0; r < ctx.nRanks; ++r) s += __ldcg(mv + r * 4 + q);
            vals[q] = s;
        }
        ctx.state->arSeq = seq;
    }
    __syncwarp();
}
#endif
