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
bool fvk_no_affine(); // true: the affine interior kernel is switched off (A/B measurements)
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
// [4096, 12288) all-reduce words u64[2][64][8] ({32 data bits | 32-bit sequence} pairs) | [FVK_P2P_HALO_OFF, ...) halo data double[2][FVK_P2P_HALO_COMPS * nGhost]
#define FVK_P2P_MAX_RANKS 64
#define FVK_P2P_MAX_NBR 32
#define FVK_P2P_HALOFLAG_OFF 0
#define FVK_P2P_ARFLAG_OFF 1024
#define FVK_P2P_ARVAL_OFF 4096
#define FVK_P2P_HALO_OFF 16384
#define FVK_P2P_HALO_COMPS 8 // doubles per ghost cell one exchange can carry (e.g. rAU + HbyA + U = 7)
// byte offset of the CG work area z[2][nOwned + nGhost] of a rank with nGhost ghost cells: behind its halo area
#define FVK_P2P_Z_OFF(nGhost) (FVK_P2P_HALO_OFF + ((sizeof(double) * 2 * FVK_P2P_HALO_COMPS * (size_t(nGhost) + 1) + 255) & ~size_t(255)))
struct FvkP2PState // device memory of the owning rank only
{
    unsigned long long haloSeq, arSeq;
    unsigned pushCounter, pad;
    unsigned long long dbg[8]; // accumulated ns of the last-block phases of k_cg_update / k_spmv (fvk_comm_p2p_debug)
};
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long fvk_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
struct FvkP2PCtx
{
    int rank, nRanks, nNbr, nOwned, nGhost;
    char* win[FVK_P2P_MAX_RANKS];         // every rank's window as mapped into this process
    int nbrRank[FVK_P2P_MAX_NBR], sendOff[FVK_P2P_MAX_NBR + 1];
    int peerRecvOff[FVK_P2P_MAX_NBR];      // where my cells start in neighbour k's ghost range
    int peerGhost[FVK_P2P_MAX_NBR];        // neighbour k's ghost count (parity stride of its halo area)
    int peerOwned[FVK_P2P_MAX_NBR];        // neighbour k's owned-cell count
    double* zwin;                          // this rank's CG work area in its window: z[2][nOwned + nGhost]
    const int* sendCells;                  // [sendOff[nNbr]] owned cells whose values the neighbours need
    FvkP2PState* state;
};
struct fvk_comm;
const FvkP2PCtx* fvk_comm_p2p_ctx(const fvk_comm* c); // nullptr: windows not connected (NCCL transport)

#ifdef __CUDACC__
// Sum of n <= 4 doubles over all ranks, executed by (at least) the first warp of ONE block per rank; vals (shared or
// global memory visible to the block) is replaced by the total. Low-latency mailbox protocol: every double travels as
// two 8-byte stores {32 data bits | 32-bit sequence number}, so data and "flag" arrive together -- no fence, no
// separate flag store; the receiver polls each word until it carries the current sequence number (one NVLink
// traversal on the critical path). Every rank adds the contributions in rank order: bit-identical everywhere.
// Mailboxes are double-buffered by the parity of the sequence number: a rank can only be one all-reduce ahead.
__device__ __forceinline__ void fvk_p2p_allreduce_sum(const FvkP2PCtx& ctx, double* vals, int n)
{
    __shared__ double ar_sh[FVK_P2P_MAX_RANKS][4];
    if (threadIdx.x >= 32) return;
    const unsigned long long seq = ctx.state->arSeq + 1;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int par = int(seq & 1);
    unsigned long long bits[4];
    for (int q = 0; q < n; ++q) bits[q] = (unsigned long long) __double_as_longlong(vals[q]);
    for (int r = threadIdx.x; r < ctx.nRanks; r += 32)
    {
        unsigned long long* slot = reinterpret_cast<unsigned long long*>(ctx.win[r] + FVK_P2P_ARVAL_OFF) + (size_t(par) * FVK_P2P_MAX_RANKS + ctx.rank) * 8;
        for (int q = 0; q < n; ++q)
        {
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + 2 * q), "l"((bits[q] & 0xffffffffull) | tag) : "memory");
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + 2 * q + 1), "l"((bits[q] >> 32) | tag) : "memory");
        }
    }
    for (int r = threadIdx.x; r < ctx.nRanks; r += 32)
    {
        const unsigned long long* slot = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_ARVAL_OFF) + (size_t(par) * FVK_P2P_MAX_RANKS + r) * 8;
        for (int q = 0; q < n; ++q)
        {
            unsigned long long lo, hi;
            do
            {
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(lo) : "l"(slot + 2 * q) : "memory");
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(hi) : "l"(slot + 2 * q + 1) : "memory");
            } while ((lo & 0xffffffff00000000ull) != tag || (hi & 0xffffffff00000000ull) != tag);
            ar_sh[r][q] = __longlong_as_double((long long) ((lo & 0xffffffffull) | (hi << 32)));
        }
    }
    __syncwarp();
    if (threadIdx.x == 0)
    {
        for (int q = 0; q < n; ++q)
        {
            double s = 0.0;
            for (int r = 0; r < ctx.nRanks; ++r) s += ar_sh[r][q];
            vals[q] = s;
        }
        ctx.state->arSeq = seq;
    }
    __syncwarp();
}
#endif

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Halo exchange of a scalar cell field executed by ONE block (the last block of the kernel that produced the field):
// store my send cells into the neighbours' windows, raise the flags; then (second call) wait for the neighbours'
// flags and copy my window into the ghost range. Returns / takes the sequence number of the exchange.
__device__ __forceinline__ unsigned long long fvk_p2p_halo_push_block(const FvkP2PCtx& ctx, const double* field)
{
    const unsigned long long seq = ctx.state->haloSeq + 1;
    const int nSend = ctx.sendOff[ctx.nNbr];
    for (int i = threadIdx.x; i < nSend; i += blockDim.x)
    {
        int k = 0;
        while (i >= ctx.sendOff[k + 1]) ++k;
        double* dst = reinterpret_cast<double*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_HALO_OFF) + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.peerGhost[k]
                      + (ctx.peerRecvOff[k] + (i - ctx.sendOff[k]));
        *dst = __ldcg(field + ctx.sendCells[i]); // written by other blocks of this kernel: read through L2
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < ctx.nNbr)
    {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(ctx.win[ctx.nbrRank[threadIdx.x]] + FVK_P2P_HALOFLAG_OFF) + ctx.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    }
    return seq;
}
__device__ __forceinline__ void fvk_p2p_halo_wait_unpack_block(const FvkP2PCtx& ctx, double* field, unsigned long long seq)
{
    if (threadIdx.x < ctx.nNbr)
    {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_HALOFLAG_OFF) + ctx.nbrRank[threadIdx.x];
        unsigned long long v;
        do
        {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        } while (v < seq);
    }
    __syncthreads();
    const double* src = reinterpret_cast<const double*>(ctx.win[ctx.rank] + FVK_P2P_HALO_OFF) + (seq & 1) * size_t(FVK_P2P_HALO_COMPS) * ctx.nGhost;
    for (int i = threadIdx.x; i < ctx.nGhost; i += blockDim.x) field[ctx.nOwned + i] = __ldcg(src + i);
    if (threadIdx.x == 0) ctx.state->haloSeq = seq;
}
#endif
