// Linear algebra of the pressure solve: CSR SpMV, BLAS-1, and the Jacobi-preconditioned CG that
// NeoN::la::Solver delegates to Ginkgo (src/NeoN/include/NeoN/linearAlgebra/ginkgo.hpp:116-155,
// src/NeoN/src/linearAlgebra/utilities.cpp:11-35, src/NeoN/src/core/vector/vectorFreeFunctions.cpp).
//
// SpMV: a persistent grid walks tiles of 256 rows. A tile's values/colIdxs range is contiguous in
// CSR, so the block streams it with fully coalesced loads, multiplies by the gathered x and parks the
// products in shared memory; then thread r adds up row r's products in ascending entry order -- the
// summation order of the reference's row loop, so y is bit-identical to the Serial executor.
// CG: two kernels per iteration, every scalar (rho, alpha, beta, norms, stop flag) lives in a device
// struct, reductions are two-stage (warp shuffle -> block partial -> last block folds the partials in
// a fixed order), so a solve is bit-reproducible run to run and needs no host round trip per iteration.
#include "fvk_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace
{
constexpr int TB = 256;          // threads per block
constexpr int SPMV_ROWS = 256;   // rows per tile
constexpr int SPMV_CAP = 2304;   // products staged per pass (18 KB)
constexpr int MAX_GRID = 148 * 8;

struct PcgState
{
    double rho, rhoPrev, alpha, beta, pq, rr, normB, normR;
    double sums[4]; // local (pre-allreduce) partial results: [0] r.z  [1] r.r  [2] p.q  [3] b.b
    double relTol, absTol;
    int iter, done, nHist, maxHist, maxIter, half; // half (BiCGStab): stopped at the half step, finalize pending
    double gamma, omega, tmp; // BiCGStab scalars; tmp = (rho/rhoPrev)(alpha/omega) of step 1
    int useP, pad2;           // BiCGStab step 1: 1 = p = r + tmp (p - omega v), 0 = p = r
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh /* [NV*8] */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sh[k * 8 + wid] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < NV; ++k)
        {
            double s = sh[k * 8];
            for (int w = 1; w < TB / 32; ++w) s += sh[k * 8 + w];
            v[k] = s;
        }
    }
}

// Grid-wide deterministic sum: every block deposits its NV partials; the last block to arrive folds
// all of them in index order (independent of arrival order) and returns true in ALL its threads with
// the totals in out[] (thread 0 only).
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double* __restrict__ partial, unsigned* __restrict__ counter,
                                         double (&out)[NV], bool sysFence = false)
{
    __shared__ double sh[NV * 8];
    __shared__ bool last;
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < NV; ++k) partial[size_t(k) * MAX_GRID + blockIdx.x] = v[k];
        if (sysFence) __threadfence_system(); // the block's stores into peer windows are performed before it checks in
        else __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
        acc[k] = 0.0;
        for (int i = threadIdx.x; i < int(gridDim.x); i += TB) acc[k] += __ldcg(&partial[size_t(k) * MAX_GRID + i]);
    }
    __syncthreads(); // sh reuse
    block_sum<NV>(acc, sh);
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = acc[k];
        *counter = 0u;
    }
    return true;
}

// ---- stopping rule + scalar updates (Ginkgo CG) ----------------------------------
__device__ __forceinline__ void decide_after_update(PcgState* st, double* __restrict__ hist)
{
    // sums[0] = r.z, sums[1] = r.r (global)
    st->rho = st->sums[0];
    st->rr = st->sums[1];
    const double normR = sqrt(st->sums[1]);
    st->normR = normR;
    if (hist && st->nHist < st->maxHist) hist[st->nHist++] = normR;
    if (st->iter >= st->maxIter || normR <= st->relTol * st->normB || normR <= st->absTol)
        st->done = 1;
    else
        st->beta = st->rho / st->rhoPrev;
}
__device__ __forceinline__ void decide_after_spmv(PcgState* st)
{
    st->pq = st->sums[2];
    st->alpha = st->rho / st->sums[2];
    st->rhoPrev = st->rho;
    st->iter += 1;
}
__global__ void k_decide_after_update(PcgState* st, double* hist)
{
    if (st->done) return;
    decide_after_update(st, hist);
}
__global__ void k_decide_after_spmv(PcgState* st)
{
    if (st->done) return;
    decide_after_spmv(st);
}
__global__ void k_set_normB(PcgState* st) { st->normB = sqrt(st->sums[3]); }

// ---- K1: x += alpha p; r -= alpha q; z = M^-1 r; r.z; r.r ------------------------------------------
template <bool FIRST, int JACOBI /* 0 none, 1 scalar Jacobi, 2 external (DIC): z and r.z are produced by the kernels that follow */>
__global__ void __launch_bounds__(TB)
k_cg_update(int n, PcgState* __restrict__ st, double* __restrict__ x, const double* __restrict__ p,
            const double* __restrict__ q, const double* rIn, double* rOut, const double* __restrict__ dinv,
            double* z, double* __restrict__ partial, unsigned* __restrict__ counter,
            double* __restrict__ hist, int distributed, const FvkP2PCtx* __restrict__ p2p)
{
    __shared__ double shv[4];
    if (st->done) return;
    const double alpha = FIRST ? 0.0 : st->alpha;
    unsigned long long hseq = 0;
    if (distributed == 2)
    {
        // z of this iteration lives in the window half selected by the exchange's parity (zero-copy halo)
        z = p2p->zwin + ((p2p->state->haloSeq + 1) & 1) * size_t(p2p->nOwned + p2p->nGhost);
        // peer-memory halo of z, spread over the whole grid: z of my send cells is formed from the kernel's INPUTS (rIn is
        // not written here: r is double-buffered in this mode) with the same expression as below and stored straight
        // into the neighbours' windows over NVLink while the main loop streams
        const FvkP2PCtx& ctx = *p2p;
        hseq = ctx.state->haloSeq + 1; // advanced by the last block only, after every block has read it
        const int nSend = ctx.sendOff[ctx.nNbr];
        for (int i = blockIdx.x * TB + threadIdx.x; i < nSend; i += gridDim.x * TB)
        {
            int k = 0;
            while (i >= ctx.sendOff[k + 1]) ++k;
            const int c = ctx.sendCells[i];
            double ri = rIn[c];
            if (!FIRST) ri = ri - alpha * q[c];
            const double zi = JACOBI == 1 ? ri * dinv[c] : ri;
            // the neighbour's ghost entry of ITS z (same parity): it reads it like any other column
            double* dst = reinterpret_cast<double*>(ctx.win[ctx.nbrRank[k]] + FVK_P2P_Z_OFF(ctx.peerGhost[k]))
                          + (hseq & 1) * size_t(ctx.peerOwned[k] + ctx.peerGhost[k]) + ctx.peerOwned[k] + ctx.peerRecvOff[k] + (i - ctx.sendOff[k]);
            *dst = zi;
        }
        // the ghost entries of x follow the same update (p's ghost entries were formed by the previous SpMV from the exchanged
        // z, alpha is global): the solution leaves the solver with current ghosts, no exchange of x afterwards
        if (!FIRST)
            for (int g = n + blockIdx.x * TB + threadIdx.x; g < ctx.nOwned + ctx.nGhost; g += gridDim.x * TB) x[g] = x[g] + alpha * p[g];
    }
    double acc[2] = {0.0, 0.0};
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB)
    {
        double ri = rIn[i];
        if (!FIRST)
        {
            x[i] = x[i] + alpha * p[i];
            ri = ri - alpha * q[i];
            rOut[i] = ri;
        }
        const double zi = JACOBI == 1 ? ri * dinv[i] : ri;
        if (JACOBI != 2)
        {
            z[i] = zi;
            acc[0] += ri * zi;
        }
        acc[1] += ri * ri;
    }
    double tot[2];
    if (distributed == 2)
    {
        // only the blocks that stored into a peer window pay for a system-scope fence
        if (!grid_sum<2>(acc, partial, counter, tot, blockIdx.x * TB < p2p->sendOff[p2p->nNbr])) return;
        // last block: every block's window stores are done. Raise the halo flags, all-reduce (r.z, r.r) through the
        // mailboxes, wait for the neighbours' flags (the next kernel reads the ghost z from the window), scalar update.
        const FvkP2PCtx& ctx = *p2p;
        const unsigned long long t0 = fvk_gtime();
        __threadfence(); // every block fenced its window stores at system scope before checking in (grid_sum, sysFence)
        if (threadIdx.x < ctx.nNbr)
            st_release_sys_u64(reinterpret_cast<unsigned long long*>(ctx.win[ctx.nbrRank[threadIdx.x]] + FVK_P2P_HALOFLAG_OFF) + ctx.rank, hseq);
        if (threadIdx.x == 0) { shv[0] = tot[0]; shv[1] = tot[1]; if (FIRST) shv[2] = st->sums[3]; }
        __syncthreads();
        const unsigned long long t1 = fvk_gtime();
        fvk_p2p_allreduce_sum(ctx, shv, FIRST ? 3 : 2); // start-up: ||b||^2 (local part left by the r0 SpMV) rides along
        const unsigned long long t2 = fvk_gtime();
        if (threadIdx.x < ctx.nNbr)
        {
            const unsigned long long* f = reinterpret_cast<const unsigned long long*>(ctx.win[ctx.rank] + FVK_P2P_HALOFLAG_OFF) + ctx.nbrRank[threadIdx.x];
            while (ld_acquire_sys_u64(f) < hseq) {}
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            const unsigned long long t3 = fvk_gtime();
            ctx.state->dbg[0] += t1 - t0; ctx.state->dbg[1] += t2 - t1; ctx.state->dbg[2] += t3 - t2; ctx.state->dbg[3] += 1;
            ctx.state->haloSeq = hseq;
            st->sums[0] = shv[0];
            st->sums[1] = shv[1];
            if (FIRST) { st->sums[3] = shv[2]; st->normB = sqrt(shv[2]); }
            decide_after_update(st, hist);
        }
        return;
    }
    if (grid_sum<2>(acc, partial, counter, tot) && threadIdx.x == 0)
    {
        st->sums[0] = tot[0];
        st->sums[1] = tot[1];
        if (!distributed && JACOBI != 2) decide_after_update(st, hist);
    }
}

// ---- multicolour DIC (diagonal incomplete Cholesky, OpenFOAM's DIC recurrences applied in COLOUR order) ---------------------
// Extension (the reference maps DIC to scalar Jacobi, fvSolution.cpp:51-55): M = (D* + L) D*^-1 (D* + U) with L / U
// the couplings to cells of lower / higher colour and D* chosen so that diag(M) = diag(A). Cells of one colour are not
// coupled, so every step is one fully parallel kernel over that colour's cells (a hex block needs 2 colours: apply = 3 kernels).
__global__ void __launch_bounds__(TB)
k_dic_diag(int n0, int n1, const int* __restrict__ cells, const int* __restrict__ rowOffs, const int* __restrict__ colIdxs,
           const double* __restrict__ values, const uint8_t* __restrict__ color, int c, int nRows, double* __restrict__ dinvStar)
{
    const int idx = n0 + blockIdx.x * TB + threadIdx.x;
    if (idx >= n1) return;
    const int i = cells[idx];
    double d = 0.0;
    for (int k = rowOffs[i]; k < rowOffs[i + 1]; ++k)
    {
        const int j = colIdxs[k];
        const double a = values[k];
        if (j == i) d += a;
        else if (j < nRows && color[j] < c) d -= a * a * dinvStar[j];
    }
    dinvStar[i] = 1.0 / d;
}
// forward substitution of colour c: z_i = (r_i - sum_{colour(j) < c} a_ij z_j) / D*_i
__global__ void __launch_bounds__(TB)
k_dic_fwd(int n0, int n1, const PcgState* __restrict__ st, const int* __restrict__ cells, const int* __restrict__ rowOffs,
          const int* __restrict__ colIdxs, const double* __restrict__ values, const uint8_t* __restrict__ color, int c, int nRows,
          const double* __restrict__ dinvStar, const double* __restrict__ r, double* __restrict__ z)
{
    if (st->done) return;
    const int idx = n0 + blockIdx.x * TB + threadIdx.x;
    if (idx >= n1) return;
    const int i = cells[idx];
    double s = r[i];
    if (c > 0)
        for (int k = rowOffs[i]; k < rowOffs[i + 1]; ++k)
        {
            const int j = colIdxs[k];
            if (j != i && j < nRows && color[j] < c) s -= values[k] * z[j];
        }
    z[i] = s * dinvStar[i];
}
// backward substitution of colour c: z_i -= (sum_{colour(j) > c} a_ij z_j) / D*_i
__global__ void __launch_bounds__(TB)
k_dic_bwd(int n0, int n1, const PcgState* __restrict__ st, const int* __restrict__ cells, const int* __restrict__ rowOffs,
          const int* __restrict__ colIdxs, const double* __restrict__ values, const uint8_t* __restrict__ color, int c, int nRows,
          const double* __restrict__ dinvStar, double* __restrict__ z)
{
    if (st->done) return;
    const int idx = n0 + blockIdx.x * TB + threadIdx.x;
    if (idx >= n1) return;
    const int i = cells[idx];
    double s = 0.0;
    for (int k = rowOffs[i]; k < rowOffs[i + 1]; ++k)
    {
        const int j = colIdxs[k];
        if (j != i && j < nRows && color[j] > c) s += values[k] * z[j];
    }
    z[i] = z[i] - dinvStar[i] * s;
}
// r.z after an external preconditioner, then the stopping test / beta (the tail of k_cg_update)
__global__ void __launch_bounds__(TB)
k_cg_rz_decide(int n, PcgState* __restrict__ st, const double* __restrict__ r, const double* __restrict__ z, double* __restrict__ partial,
               unsigned* __restrict__ counter, double* __restrict__ hist)
{
    if (st->done) return;
    double acc[1] = {0.0};
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) acc[0] += r[i] * z[i];
    double tot[1];
    if (grid_sum<1>(acc, partial, counter, tot) && threadIdx.x == 0)
    {
        st->sums[0] = tot[0];
        decide_after_update(st, hist);
    }
}

// p = z + beta p on the owned rows (distributed path: followed by a halo exchange of p)
__global__ void __launch_bounds__(TB)
k_cg_pupdate(int n, const PcgState* __restrict__ st, const double* __restrict__ z, double* __restrict__ p)
{
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) p[i] = z[i] + beta * p[i];
}


// ---- BiCGStab (Ginkgo 1.10 solver::Bicgstab, restated in oracle/fvo.cpp fvo_bicgstab) -------------------------------
// Five kernels per iteration: {p, y} | {v = A y, rr.v} | {s, z, s.s} | {t = A z, s.t, t.t} | {x, r, rr.r, r.r}; the scalar
// updates and both stopping checks run in the last block of the kernel that finishes the reduction they depend on.
__device__ __forceinline__ void bicg_decide_alpha(PcgState* st)
{
    st->beta = st->sums[2];
    st->alpha = st->sums[2] != 0.0 ? st->rho / st->sums[2] : 0.0;
}
__device__ __forceinline__ void bicg_decide_half(PcgState* st, double* __restrict__ hist)
{
    const double normS = sqrt(st->sums[1]);
    st->normR = normS;
    if (hist && st->nHist < st->maxHist) hist[st->nHist++] = normS;
    if (st->iter >= st->maxIter || normS <= st->relTol * st->normB || normS <= st->absTol) st->half = 1;
}
__device__ __forceinline__ void bicg_decide_omega(PcgState* st)
{
    st->gamma = st->sums[2];
    st->omega = st->sums[3] != 0.0 ? st->sums[2] / st->sums[3] : 0.0;
}
// after step 3 (or the start-up): sums[0] = rr.r, sums[1] = r.r
template <bool FIRST>
__device__ __forceinline__ void bicg_decide_full(PcgState* st, double* __restrict__ hist)
{
    if (st->half)
    { // the half-step check stopped the solve; x was finalised by this kernel
        st->done = 1;
        return;
    }
    if (!FIRST)
    {
        st->rhoPrev = st->rho; // swap(prev_rho, rho) followed by the next iteration's rho = rr.r
        st->iter += 1;
    }
    st->rho = st->sums[0];
    const double normR = sqrt(st->sums[1]);
    st->normR = normR;
    if (hist && st->nHist < st->maxHist) hist[st->nHist++] = normR;
    if (st->iter >= st->maxIter || normR <= st->relTol * st->normB || normR <= st->absTol)
    {
        st->done = 1;
        return;
    }
    if (st->rhoPrev * st->omega != 0.0)
    {
        st->tmp = (st->rho / st->rhoPrev) * (st->alpha / st->omega);
        st->useP = 1;
    }
    else
        st->useP = 0;
}
template <int STAGE>
__global__ void k_bicg_decide(PcgState* st, double* hist)
{
    if (st->done) return;
    if (STAGE == 0) bicg_decide_full<true>(st, hist);
    if (STAGE == 1) bicg_decide_alpha(st);
    if (STAGE == 2) bicg_decide_half(st, hist);
    if (STAGE == 3 && !st->half) bicg_decide_omega(st);
    if (STAGE == 4) bicg_decide_full<false>(st, hist);
}

// step 1 + preconditioner: p = r + tmp (p - omega v); y = M^-1 p
template <bool JACOBI>
__global__ void __launch_bounds__(TB)
k_bicg_step1(int n, const PcgState* __restrict__ st, const double* __restrict__ r, double* __restrict__ p,
             const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ y)
{
    if (st->done) return;
    const double tmp = st->tmp, omega = st->omega;
    const bool useP = st->useP != 0;
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB)
    {
        const double pi = useP ? r[i] + tmp * (p[i] - omega * v[i]) : r[i];
        p[i] = pi;
        y[i] = JACOBI ? pi * dinv[i] : pi;
    }
}
// step 2 + preconditioner + ||s||^2: s = r - alpha v; z = M^-1 s
template <bool JACOBI>
__global__ void __launch_bounds__(TB)
k_bicg_step2(int n, PcgState* __restrict__ st, const double* __restrict__ r, const double* __restrict__ v,
             const double* __restrict__ dinv, double* __restrict__ sv, double* __restrict__ z, double* __restrict__ partial,
             unsigned* __restrict__ counter, double* __restrict__ hist, int distributed)
{
    if (st->done) return;
    const double alpha = st->alpha;
    const bool plain = st->beta == 0.0; // step_2 of the reference kernel: beta == 0 -> alpha = 0, s = r
    double acc[1] = {0.0};
    for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB)
    {
        const double si = plain ? r[i] : r[i] - alpha * v[i];
        sv[i] = si;
        z[i] = JACOBI ? si * dinv[i] : si;
        acc[0] += si * si;
    }
    double tot[1];
    if (grid_sum<1>(acc, partial, counter, tot) && threadIdx.x == 0)
    {
        st->sums[1] = tot[0];
        if (!distributed) bicg_decide_half(st, hist);
    }
}
// step 3 (+ finalize after a half-step stop) + rho = rr.r + ||r||^2; FIRST: start-up (rr = r, the first check)
template <bool FIRST>
__global__ void __launch_bounds__(TB)
k_bicg_step3(int n, PcgState* __restrict__ st, double* __restrict__ x, const double* __restrict__ y,
             const double* __restrict__ z, const double* __restrict__ sv, const double* __restrict__ t,
             double* __restrict__ r, double* __restrict__ rr, double* __restrict__ partial, unsigned* __restrict__ counter,
             double* __restrict__ hist, int distributed)
{
    if (st->done) return;
    double acc[2] = {0.0, 0.0};
    if (FIRST)
    {
        for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB)
        {
            const double ri = r[i];
            rr[i] = ri;
            acc[0] += ri * ri;
        }
        acc[1] = acc[0];
    }
    else if (st->half)
    {
        const double alpha = st->alpha;
        for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB) x[i] = x[i] + alpha * y[i];
    }
    else
    {
        const double alpha = st->alpha, omega = st->omega;
        for (int i = blockIdx.x * TB + threadIdx.x; i < n; i += gridDim.x * TB)
        {
            x[i] = x[i] + (alpha * y[i] + omega * z[i]);
            const double ri = sv[i] - omega * t[i];
            r[i] = ri;
            acc[0] += rr[i] * ri;
            acc[1] += ri * ri;
        }
    }
    double tot[2];
    if (grid_sum<2>(acc, partial, counter, tot) && threadIdx.x == 0)
    {
        st->sums[0] = tot[0];
        st->sums[1] = tot[1];
        if (!distributed) bicg_decide_full<FIRST>(st, hist);
    }
}

// Vec3 systems (identical components, A.3): component matrix / right-hand side / solution (un)packing
__global__ void __launch_bounds__(TB) k_take_component(int64_t n, int comp, const double* __restrict__ v3, double* __restrict__ out)
{
    for (int64_t i = int64_t(blockIdx.x) * TB + threadIdx.x; i < n; i += int64_t(gridDim.x) * TB) out[i] = v3[3 * i + comp];
}
__global__ void __launch_bounds__(TB) k_put_component(int64_t n, int comp, const double* __restrict__ in, double* __restrict__ v3)
{
    for (int64_t i = int64_t(blockIdx.x) * TB + threadIdx.x; i < n; i += int64_t(gridDim.x) * TB) v3[3 * i + comp] = in[i];
}

// ---- tiled CSR SpMV ----------------------------------------------------------------------------
// MODE 0: y = A x            (fvk_spmv)
// MODE 1: y = A x - b        (computeResidual)
// MODE 2: y = b - A x        (CG start-up r0)
// MODE 3: CG: q = A p, dot p.q, scalar update; p read from `x`
// MODE 4: CG fused: pNew = z + beta pOld evaluated on the fly for the gathered columns, q = A pNew
// MODE 5: BiCGStab: v = A y, dot rr.v (rr passed as `b`), alpha update
// MODE 6: BiCGStab: t = A z, dots s.t and t.t (s passed as `b`), omega update
// Block-structured sparsity (the mesh plan proved the topology, FvkBrickGeom::affine): the row of a REGULAR cell c is
// [c-nx*ny, c-nx, c-1 | c | c+1, c+nx, c+nx*ny]; its column indices are not read (28 of the 104 bytes a row moves).
struct SpmvAffine
{
    int on, nx, ny, nz;
};

template <int MODE>
__global__ void __launch_bounds__(TB)
k_spmv(int nRows, const int* __restrict__ rowOffs, const int* __restrict__ colIdxs, const double* __restrict__ values,
       const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y, PcgState* __restrict__ st,
       const double* __restrict__ z, double* __restrict__ pNew, double* __restrict__ partial,
       unsigned* __restrict__ counter, int distributed, const FvkP2PCtx* __restrict__ p2p = nullptr, int nCols = 0,
       SpmvAffine aff = SpmvAffine {0, 0, 0, 0}, const uint8_t* __restrict__ diagOffs = nullptr, double* __restrict__ dinvOut = nullptr)
{
    __shared__ double prod[SPMV_CAP];
    __shared__ int ro[SPMV_ROWS + 1];
    double beta = 0.0;
    if (MODE >= 3)
    {
        if (st->done) return;
        if (MODE == 6 && st->half) return;
        beta = st->beta;
    }
    constexpr int NV = MODE == 6 ? 2 : 1;
    double acc[NV] = {0.0};
    // peer-memory CG: z (owned + ghost entries, the latter stored by the neighbours' update kernels, flags already
    // awaited) lies in the window half of the last exchange's parity; p of the ghost columns follows p = z + beta p
    const double* zz = z; // read with __ldg below: the non-coherent path the __restrict__ parameter would get
    if (MODE == 4 && distributed == 2)
    {
        zz = p2p->zwin + (p2p->state->haloSeq & 1) * size_t(nCols);
        for (int g = nRows + blockIdx.x * TB + threadIdx.x; g < nCols; g += gridDim.x * TB) pNew[g] = __ldg(zz + g) + beta * x[g];
    }
    const int nTiles = (nRows + SPMV_ROWS - 1) / SPMV_ROWS;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x)
    {
        const int r0 = tile * SPMV_ROWS;
        const int nr = min(SPMV_ROWS, nRows - r0);
        __syncthreads(); // previous tile done with ro/prod
        for (int i = threadIdx.x; i <= nr; i += TB) ro[i] = rowOffs[r0 + i];
        __syncthreads();
        const int base = ro[0], end = ro[nr];
        const bool hasRow = threadIdx.x < nr;
        int k = hasRow ? ro[threadIdx.x] : 0;
        const int kend = hasRow ? ro[threadIdx.x + 1] : 0;
        double sum = 0.0;
        if (aff.on && end - base <= SPMV_CAP)
        {
            // structured tile: stream the VALUES only (coalesced), then every row multiplies its own entries in entry
            // order -- the same products in the same order as below, columns from arithmetic for regular rows
            for (int e = base + threadIdx.x; e < end; e += TB) prod[e - base] = ld_stream(values + e);
            __syncthreads();
            if (hasRow)
            {
                const int c = r0 + threadIdx.x;
                const int i = c % aff.nx, q = c / aff.nx, j = q % aff.ny, kz = q / aff.ny;
                const bool reg = i > 0 && i < aff.nx - 1 && j > 0 && j < aff.ny - 1 && kz > 0 && kz < aff.nz - 1 && kend - k == 7;
                if (reg)
                {
                    const int nxy = aff.nx * aff.ny;
                    const int col[7] = {c - nxy, c - aff.nx, c - 1, c, c + 1, c + aff.nx, c + nxy};
                    double xv[7];
#pragma unroll
                    for (int t = 0; t < 7; ++t) xv[t] = (MODE == 4) ? (__ldg(zz + col[t]) + beta * x[col[t]]) : x[col[t]];
#pragma unroll
                    for (int t = 0; t < 7; ++t) sum += prod[k - base + t] * xv[t];
                }
                else
                    for (; k < kend; ++k)
                    {
                        const int jc = colIdxs[k];
                        const double xv = (MODE == 4) ? (__ldg(zz + jc) + beta * x[jc]) : x[jc];
                        sum += prod[k - base] * xv;
                    }
            }
        }
        else
        for (int cb = base; cb < end; cb += SPMV_CAP)
        {
            const int ce = min(end, cb + SPMV_CAP);
            if (cb != base) __syncthreads();
#pragma unroll 4
            for (int e = cb + threadIdx.x; e < ce; e += TB)
            {
                const int j = colIdxs[e];
                const double xv = (MODE == 4) ? (__ldg(zz + j) + beta * x[j]) : x[j];
                prod[e - cb] = ld_stream(values + e) * xv;
            }
            __syncthreads();
            const int ke = min(kend, ce);
            for (; k < ke; ++k) sum += prod[k - cb];
        }
        if (hasRow)
        {
            const int r = r0 + threadIdx.x;
            if (MODE == 0) y[r] = sum;
            if (MODE == 1) y[r] = sum - b[r];
            if (MODE == 2)
            {
                const double br = b[r];
                y[r] = br - sum;
                acc[0] += br * br;
                // solver start-up: the scalar Jacobi preconditioner's 1 / a_rr from the row just streamed (k_extract_dinv_offs)
                if (dinvOut) dinvOut[r] = 1.0 / values[ro[threadIdx.x] + diagOffs[r]];
            }
            if (MODE == 3)
            {
                y[r] = sum;
                acc[0] += x[r] * sum;
            }
            if (MODE == 4)
            {
                const double pr = __ldg(zz + r) + beta * x[r];
                pNew[r] = pr;
                y[r] = sum;
                acc[0] += pr * sum;
            }
            if (MODE == 5)
            {
                y[r] = sum;
                acc[0] += b[r] * sum;
            }
            if (MODE == 6)
            {
                y[r] = sum;
                acc[0] += b[r] * sum;
                acc[NV - 1] += sum * sum;
            }
        }
    }
    if (MODE == 2)
    { // solver start-up: ||b||^2 next to r0 = b - A x (st == nullptr: plain residual)
        if (!st) return;
        double tot[NV];
        if (grid_sum<NV>(acc, partial, counter, tot) && threadIdx.x == 0)
        {
            st->sums[3] = tot[0];
            if (!distributed) st->normB = sqrt(tot[0]);
        }
        return;
    }
    if (MODE == 5 || MODE == 6)
    {
        double tot[NV];
        if (grid_sum<NV>(acc, partial, counter, tot) && threadIdx.x == 0)
        {
            st->sums[2] = tot[0];
            if (MODE == 6) st->sums[3] = tot[NV - 1];
            if (!distributed)
            {
                if (MODE == 5) bicg_decide_alpha(st);
                else bicg_decide_omega(st);
            }
        }
        return;
    }
    if (MODE == 3 || MODE == 4)
    {
        double tot[NV];
        if (grid_sum<NV>(acc, partial, counter, tot))
        {
            if (distributed == 2)
            { // peer-memory all-reduce of p.q inside the SpMV's last block
                __syncthreads();
                if (threadIdx.x == 0) prod[0] = tot[0];
                __syncthreads();
                const unsigned long long t0 = fvk_gtime();
                fvk_p2p_allreduce_sum(*p2p, prod, 1);
                __syncthreads();
                if (threadIdx.x == 0)
                {
                    p2p->state->dbg[4] += fvk_gtime() - t0; p2p->state->dbg[5] += 1;
                    st->sums[2] = prod[0];
                    decide_after_spmv(st);
                }
            }
            else if (threadIdx.x == 0)
            {
                st->sums[2] = tot[0];
                if (!distributed) decide_after_spmv(st);
            }
        }
    }
}

// diagonal through the SparsityPattern's diagOffset (the attached mesh's own pattern): one entry per row instead of the row scan
__global__ void __launch_bounds__(TB)
k_extract_dinv_offs(int n, const int* __restrict__ rowOffs, const uint8_t* __restrict__ diagOffs, const double* __restrict__ values,
                    double* __restrict__ dinv)
{
    const int r = blockIdx.x * TB + threadIdx.x;
    if (r >= n) return;
    dinv[r] = 1.0 / values[rowOffs[r] + diagOffs[r]];
}
__global__ void __launch_bounds__(TB)
k_extract_dinv(int n, const int* __restrict__ rowOffs, const int* __restrict__ colIdxs, const double* __restrict__ values,
               double* __restrict__ dinv)
{
    const int r = blockIdx.x * TB + threadIdx.x;
    if (r >= n) return;
    double d = 1.0;
    for (int k = rowOffs[r]; k < rowOffs[r + 1]; ++k)
        if (colIdxs[k] == r) d = 1.0 / values[k];
    dinv[r] = d;
}

// ---- BLAS-1 --------------------------------------------------------------------------------------
enum { OP_FILL, OP_SCALE, OP_ADD, OP_SUB, OP_MUL, OP_AXPBY, OP_SCALED_COPY };
// w = a x + b y (three-array form of axpby: forwardEuler's solution = old - source * dt without the copy of `old`)
__global__ void __launch_bounds__(TB) k_waxpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y, double* __restrict__ w)
{
    for (int64_t i = int64_t(blockIdx.x) * TB + threadIdx.x; i < n; i += int64_t(gridDim.x) * TB) w[i] = a * x[i] + b * y[i];
}
template <int OP>
__global__ void __launch_bounds__(TB)
k_vec(int64_t n, double a, double bb, double* __restrict__ x, const double* __restrict__ y)
{
    for (int64_t i = int64_t(blockIdx.x) * TB + threadIdx.x; i < n; i += int64_t(gridDim.x) * TB)
    {
        if (OP == OP_FILL) x[i] = a;
        if (OP == OP_SCALE) x[i] = x[i] * a;
        if (OP == OP_ADD) x[i] = x[i] + y[i];
        if (OP == OP_SUB) x[i] = x[i] - y[i];
        if (OP == OP_MUL) x[i] = x[i] * y[i];
        if (OP == OP_AXPBY) x[i] = a * y[i] + bb * x[i]; // here x is the output `y` of fvk_vec_axpby
        if (OP == OP_SCALED_COPY) x[i] = y[i] * a;
    }
}

template <bool NORM>
__global__ void __launch_bounds__(TB)
k_dot(int64_t n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ result,
      double* __restrict__ partial, unsigned* __restrict__ counter)
{
    double acc[1] = {0.0};
    for (int64_t i = int64_t(blockIdx.x) * TB + threadIdx.x; i < n; i += int64_t(gridDim.x) * TB) acc[0] += x[i] * y[i];
    double tot[1];
    if (grid_sum<1>(acc, partial, counter, tot) && threadIdx.x == 0) result[0] = NORM ? sqrt(tot[0]) : tot[0];
}

int stream_grid(int64_t n, int perThread = 4)
{
    int64_t g = (n + int64_t(TB) * perThread - 1) / (int64_t(TB) * perThread);
    const int cap = min(MAX_GRID, fvk_sm_count() * 8);
    if (g > cap) g = cap;
    return g < 1 ? 1 : int(g);
}
int spmv_grid(int nRows)
{
    const int tiles = (nRows + SPMV_ROWS - 1) / SPMV_ROWS;
    const int cap = min(MAX_GRID, fvk_sm_count() * 8);
    return tiles < 1 ? 1 : (tiles < cap ? tiles : cap);
}

// ---- bandwidth probe (diagnostics): out[i] = sum_k in_k[i] over NS concurrent read streams + one write stream ------
// Same launch shape as the operator kernels (one 8-byte element per thread and stream, 256-thread blocks): measures
// what HBM delivers for NS interleaved streams, the practical ceiling of kernels that read 6-8 arrays at once.
struct ProbePtrs { const double* p[8]; };
template <int NS>
__global__ void __launch_bounds__(256)
k_probe_streams(ProbePtrs in, int64_t n, double* __restrict__ out)
{
    const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (i >= n) return;
    double v[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = in.p[k][i];
    double s = v[0];
#pragma unroll
    for (int k = 1; k < NS; ++k) s += v[k];
    out[i] = s;
}

// per-device scratch of the stand-alone reductions (fvk_dot / fvk_norm2)
struct RedScratch
{
    double* partial = nullptr;
    unsigned* counter = nullptr;
};
int get_scratch(RedScratch** out)
{
    static RedScratch tab[64];
    int dev = 0;
    FVK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fvk_fail(FVK_EUNSUPPORTED, "device index %d", dev);
    RedScratch& s = tab[dev];
    if (!s.partial)
    {
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.partial), sizeof(double) * 4 * MAX_GRID));
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.counter), sizeof(unsigned)));
        FVK_CUDA(cudaMemset(s.counter, 0, sizeof(unsigned)));
    }
    *out = &s;
    return FVK_OK;
}
} // namespace

// comm hooks implemented in fvk_comm.cu
int fvk_comm_allreduce_sum_impl(fvk_comm* comm, double* data_d, int count, cudaStream_t st);
int fvk_comm_halo_exchange_impl(fvk_comm* comm, double* field_d, int ncomp, cudaStream_t st);

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int fvk_spmv(int32_t nRows, const int32_t* rowOffs, const int32_t* colIdxs, const double* values,
                        const double* x, double* y, fvk_stream s)
{
    if (nRows < 0 || !rowOffs || (nRows && (!colIdxs || !values || !x || !y))) return fvk_fail(FVK_EINVAL, "fvk_spmv: bad argument");
    if (nRows == 0) return FVK_OK;
    k_spmv<0><<<spmv_grid(nRows), TB, 0, fvk_cu(s)>>>(nRows, rowOffs, colIdxs, values, x, nullptr, y, nullptr, nullptr,
                                                      nullptr, nullptr, nullptr, 0);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
static SpmvAffine mesh_affine(const fvk_mesh* m)
{
    static const bool off = [] { const char* e = std::getenv("FVK_SPMV_NO_AFFINE"); return e && *e == '1'; }();
    const FvkBrickGeom& g = m->bp.geom;
    // rows [lower faces ascending | diag | upper faces ascending] + the plan's proof give the regular rows' columns
    if (off || !g.affine || m->bp.nTiles == 0 || int64_t(g.dims[0]) * g.dims[1] * g.dims[2] != m->nOwned) return SpmvAffine {0, 0, 0, 0};
    return SpmvAffine {1, g.dims[0], g.dims[1], g.dims[2]};
}
extern "C" int fvk_spmv_structured(const fvk_mesh* m, const double* values, const double* x, double* y, fvk_stream s)
{
    if (!m || !values || !x || !y) return fvk_fail(FVK_EINVAL, "fvk_spmv_structured: null argument");
    const int nRows = m->nOwned;
    if (nRows == 0) return FVK_OK;
    k_spmv<0><<<spmv_grid(nRows), TB, 0, fvk_cu(s)>>>(nRows, m->rowOffs, m->colIdxs, values, x, nullptr, y, nullptr, nullptr, nullptr, nullptr,
                                                      nullptr, 0, nullptr, 0, mesh_affine(m));
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_residual(int32_t nRows, const int32_t* rowOffs, const int32_t* colIdxs, const double* values,
                            const double* b, const double* x, double* res, fvk_stream s)
{
    if (nRows < 0 || !rowOffs || (nRows && (!colIdxs || !values || !x || !b || !res)))
        return fvk_fail(FVK_EINVAL, "fvk_residual: bad argument");
    if (nRows == 0) return FVK_OK;
    k_spmv<1><<<spmv_grid(nRows), TB, 0, fvk_cu(s)>>>(nRows, rowOffs, colIdxs, values, x, b, res, nullptr, nullptr,
                                                      nullptr, nullptr, nullptr, 0);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_probe_streams(int nStreams, const double* const* in_h, int64_t n, double* out, fvk_stream s)
{
    if (nStreams < 1 || nStreams > 8 || !in_h || !out || n < 0) return fvk_fail(FVK_EINVAL, "fvk_probe_streams: bad argument");
    if (n == 0) return FVK_OK;
    ProbePtrs pp;
    for (int k = 0; k < 8; ++k) pp.p[k] = in_h[k < nStreams ? k : 0];
    const unsigned grid = unsigned((n + 255) / 256);
    switch (nStreams)
    {
        case 1: k_probe_streams<1><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 2: k_probe_streams<2><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 3: k_probe_streams<3><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 4: k_probe_streams<4><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 5: k_probe_streams<5><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 6: k_probe_streams<6><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        case 7: k_probe_streams<7><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
        default: k_probe_streams<8><<<grid, 256, 0, fvk_cu(s)>>>(pp, n, out); break;
    }
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

#define VEC_OP(OP, a, b, x, y)                                                                     \
    do                                                                                             \
    {                                                                                              \
        if (n < 0 || (n && !(x))) return fvk_fail(FVK_EINVAL, "%s: bad argument", __func__);       \
        if (n == 0) return FVK_OK;                                                                 \
        k_vec<OP><<<stream_grid(n), TB, 0, fvk_cu(s)>>>(n, a, b, x, y);                            \
        FVK_LAUNCH_CHECK();                                                                        \
        return FVK_OK;                                                                             \
    } while (0)

extern "C" int fvk_vec_fill(int64_t n, double v, double* x, fvk_stream s) { VEC_OP(OP_FILL, v, 0.0, x, nullptr); }
extern "C" int fvk_vec_scale(int64_t n, double a, double* x, fvk_stream s) { VEC_OP(OP_SCALE, a, 0.0, x, nullptr); }
extern "C" int fvk_vec_add(int64_t n, double* x, const double* y, fvk_stream s)
{
    if (n && !y) return fvk_fail(FVK_EINVAL, "fvk_vec_add: null");
    VEC_OP(OP_ADD, 0.0, 0.0, x, y);
}
extern "C" int fvk_vec_sub(int64_t n, double* x, const double* y, fvk_stream s)
{
    if (n && !y) return fvk_fail(FVK_EINVAL, "fvk_vec_sub: null");
    VEC_OP(OP_SUB, 0.0, 0.0, x, y);
}
extern "C" int fvk_vec_mul(int64_t n, double* x, const double* y, fvk_stream s)
{
    if (n && !y) return fvk_fail(FVK_EINVAL, "fvk_vec_mul: null");
    VEC_OP(OP_MUL, 0.0, 0.0, x, y);
}
extern "C" int fvk_vec_axpby(int64_t n, double a, const double* x, double b, double* y, fvk_stream s)
{
    if (n && !x) return fvk_fail(FVK_EINVAL, "fvk_vec_axpby: null");
    VEC_OP(OP_AXPBY, a, b, y, x);
}

extern "C" int fvk_vec_waxpby(int64_t n, double a, const double* x, double b, const double* y, double* w, fvk_stream s)
{
    if (n < 0 || (n && (!x || !y || !w))) return fvk_fail(FVK_EINVAL, "fvk_vec_waxpby: bad argument");
    if (n == 0) return FVK_OK;
    k_waxpby<<<stream_grid(n), TB, 0, fvk_cu(s)>>>(n, a, x, b, y, w);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_vec_scaled_copy(int64_t n, double a, const double* x, double* out, fvk_stream s)
{
    if (n && !x) return fvk_fail(FVK_EINVAL, "fvk_vec_scaled_copy: null");
    VEC_OP(OP_SCALED_COPY, a, 0.0, out, x);
}

extern "C" int fvk_dot(int64_t n, const double* x, const double* y, double* result_d, fvk_stream s)
{
    if (n < 0 || !result_d || (n && (!x || !y))) return fvk_fail(FVK_EINVAL, "fvk_dot: bad argument");
    RedScratch* sc = nullptr;
    if (int rc = get_scratch(&sc)) return rc;
    k_dot<false><<<stream_grid(n), TB, 0, fvk_cu(s)>>>(n, x, y, result_d, sc->partial, sc->counter);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_norm2(int64_t n, const double* x, double* result_d, fvk_stream s)
{
    if (n < 0 || !result_d || (n && !x)) return fvk_fail(FVK_EINVAL, "fvk_norm2: bad argument");
    RedScratch* sc = nullptr;
    if (int rc = get_scratch(&sc)) return rc;
    k_dot<true><<<stream_grid(n), TB, 0, fvk_cu(s)>>>(n, x, x, result_d, sc->partial, sc->counter);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

// ---- solver ---------------------------------------------------------------------------------------
struct fvk_solver
{
    int32_t nRows = 0, nCols = 0;
    fvk_solver_config cfg {};
    fvk_comm* comm = nullptr;
    bool guessGhostsCurrent = false; // fvk_solver_set_ghosts_current: skip the exchange of the initial guess
    SpmvAffine aff {0, 0, 0, 0}; // set by fvk_solver_attach_mesh
    const int32_t *affRowOffs = nullptr, *affColIdxs = nullptr; // the attached mesh's pattern: aff applies to these arrays only
    const uint8_t* affDiagOffs = nullptr;                       // its diagOffset (Jacobi diagonal without a row scan)
    const fvk_mesh* attached = nullptr;                         // the attached mesh (DIC needs its colouring)
    double *r2 = nullptr; // second residual buffer of the peer-memory mode
    double *rr = nullptr, *sB = nullptr, *tB = nullptr; // BiCGStab: shadow residual, s, t
    double *vals0 = nullptr, *bC = nullptr, *xC = nullptr; // Vec3 solves: component matrix / rhs / solution (lazy)
    int64_t vals0Cap = 0;
    double *r = nullptr, *z = nullptr, *p0 = nullptr, *p1 = nullptr, *q = nullptr, *dinv = nullptr;
    double *partial = nullptr, *hist = nullptr;
    unsigned* counter = nullptr;
    PcgState* state = nullptr;
    PcgState* state_h = nullptr; // pinned, [2]: double-buffered stop checks
    // stream-capture mode (fvk_solver_solve on a capturing stream): the iteration is a conditional WHILE node of the graph
    // being captured -- no host round trip; every captured solve copies its final state into its own pinned slot
    PcgState* init_h = nullptr;  // pinned start state the graph's memcpy node reads
    PcgState* cap_h = nullptr;   // pinned [FVK_MAX_CAPTURED_SOLVES]
    int nCaptured = 0;
    cudaStream_t bodyStream = nullptr;
    // device-side log of the captured solves' iteration counts (one entry per executed solve, in execution order), so that
    // back-to-back graph replays need no host synchronisation to keep their statistics
    int32_t* capLog = nullptr;   // [FVK_CAPTURE_LOG]
    unsigned* capLogCount = nullptr;
    cudaEvent_t checkEv[2] = {nullptr, nullptr};
    int32_t histCap = 0;
};

extern "C" int fvk_solver_destroy(fvk_solver* sv)
{
    if (!sv) return FVK_OK;
    for (void* ptr : {(void*) sv->rr, (void*) sv->sB, (void*) sv->tB, (void*) sv->vals0, (void*) sv->bC, (void*) sv->xC})
        if (ptr) cudaFree(ptr);
    for (void* ptr : {(void*) sv->r, (void*) sv->r2, (void*) sv->z, (void*) sv->p0, (void*) sv->p1, (void*) sv->q, (void*) sv->dinv,
                      (void*) sv->partial, (void*) sv->hist, (void*) sv->counter, (void*) sv->state})
        if (ptr) cudaFree(ptr);
    if (sv->state_h) cudaFreeHost(sv->state_h);
    if (sv->init_h) cudaFreeHost(sv->init_h);
    if (sv->cap_h) cudaFreeHost(sv->cap_h);
    if (sv->bodyStream) cudaStreamDestroy(sv->bodyStream);
    if (sv->capLog) cudaFree(sv->capLog);
    if (sv->capLogCount) cudaFree(sv->capLogCount);
    for (auto& e : sv->checkEv)
        if (e) cudaEventDestroy(e);
    delete sv;
    return FVK_OK;
}

extern "C" int fvk_solver_create(int32_t nRows, int32_t nCols, const fvk_solver_config* cfg, fvk_comm* comm, fvk_solver** out)
{
    if (!cfg || !out || nRows <= 0 || nCols < nRows) return fvk_fail(FVK_EINVAL, "fvk_solver_create: bad argument");
    if (cfg->maxIter < 0 || cfg->checkEvery < 1 || cfg->preconditioner < FVK_PRECOND_NONE || cfg->preconditioner > FVK_PRECOND_DIC
        || (cfg->preconditioner == FVK_PRECOND_DIC && (cfg->solverType != FVK_SOLVER_CG || comm))
        || (cfg->solverType != FVK_SOLVER_CG && cfg->solverType != FVK_SOLVER_BICGSTAB))
        return fvk_fail(FVK_EINVAL, "fvk_solver_create: bad configuration");
    *out = nullptr;
    fvk_solver* sv = new fvk_solver;
    sv->nRows = nRows; sv->nCols = nCols; sv->cfg = *cfg; sv->comm = comm;
    sv->histCap = (cfg->solverType == FVK_SOLVER_BICGSTAB ? 2 : 1) * cfg->maxIter + 2;
    cudaError_t e = cudaSuccess;
    auto A = [&](double** ptr, size_t n) { if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(ptr), sizeof(double) * n); };
    A(&sv->r, nRows); if (comm) A(&sv->r2, nRows); A(&sv->z, nCols); A(&sv->p0, nCols); A(&sv->p1, nCols); A(&sv->q, nRows); A(&sv->dinv, nRows);
    if (cfg->solverType == FVK_SOLVER_BICGSTAB) { A(&sv->rr, nRows); A(&sv->sB, nRows); A(&sv->tB, nRows); }
    A(&sv->partial, 4 * MAX_GRID); A(&sv->hist, sv->histCap);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sv->counter), sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(sv->counter, 0, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sv->state), sizeof(PcgState));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&sv->state_h), 2 * sizeof(PcgState));
    // stream-capture mode needs these before a capture starts (allocations are not allowed while a stream captures)
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&sv->init_h), sizeof(PcgState));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&sv->cap_h), sizeof(PcgState) * 64);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sv->bodyStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sv->capLog), sizeof(int32_t) * 8192);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sv->capLogCount), sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(sv->capLogCount, 0, sizeof(unsigned));
    for (auto& ev : sv->checkEv)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess)
    {
        fvk_solver_destroy(sv);
        return fvk_fail(e == cudaErrorNoDevice ? FVK_ENODEVICE : FVK_ECUDA, "fvk_solver_create: %s", cudaGetErrorString(e));
    }
    *out = sv;
    return FVK_OK;
}

extern "C" int fvk_solver_attach_mesh(fvk_solver* sv, const fvk_mesh* m)
{
    if (!sv) return fvk_fail(FVK_EINVAL, "fvk_solver_attach_mesh: null solver");
    sv->aff = (m && m->nOwned == sv->nRows && m->nCells == sv->nCols) ? mesh_affine(m) : SpmvAffine {0, 0, 0, 0};
    sv->affRowOffs = m ? m->rowOffs : nullptr;
    sv->affColIdxs = m ? m->colIdxs : nullptr;
    sv->affDiagOffs = (m && m->nOwned == sv->nRows && m->nCells == sv->nCols) ? m->diagOffset : nullptr;
    sv->attached = (m && m->nOwned == sv->nRows && m->nCells == sv->nCols) ? m : nullptr;
    return FVK_OK;
}


// the CG start-up SpMV (r0 = b - A x) can write 1 / a_rr itself when the diagonal's slot in every row is known
static bool dinv_fusable(const fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs)
{
    return sv->affDiagOffs && rowOffs == sv->affRowOffs && colIdxs == sv->affColIdxs;
}
static int launch_dinv(fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs, const double* values, cudaStream_t st)
{
    const int n = sv->nRows;
    if (sv->affDiagOffs && rowOffs == sv->affRowOffs && colIdxs == sv->affColIdxs)
        k_extract_dinv_offs<<<(n + TB - 1) / TB, TB, 0, st>>>(n, rowOffs, sv->affDiagOffs, values, sv->dinv);
    else
        k_extract_dinv<<<(n + TB - 1) / TB, TB, 0, st>>>(n, rowOffs, colIdxs, values, sv->dinv);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

// Pipelined stop polling shared by the solvers: check after 1, 2, 4, ... rounds up to `every`, then every `every`; the
// state copy of check k is awaited only after the rounds up to check k+1 have been queued.
struct StopPoll
{
    fvk_solver* sv;
    cudaStream_t st;
    int every, nextCheck = 1, pending = -1, finalBuf = 0;
    // returns 1 when the solve has stopped, 0 to continue, <0 on error (-(code))
    int after_round(int it, int lastRound)
    {
        if (it + 1 != nextCheck && it != lastRound) return 0;
        if (pending >= 0)
        {
            if (cudaEventSynchronize(sv->checkEv[pending]) != cudaSuccess) return -FVK_ECUDA;
            if (sv->state_h[pending].done) { finalBuf = pending; return 1; }
        }
        pending = pending < 0 ? 0 : 1 - pending;
        if (cudaMemcpyAsync(&sv->state_h[pending], sv->state, sizeof(PcgState), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -FVK_ECUDA;
        if (cudaEventRecord(sv->checkEv[pending], st) != cudaSuccess) return -FVK_ECUDA;
        if (it == lastRound)
        {
            if (cudaEventSynchronize(sv->checkEv[pending]) != cudaSuccess) return -FVK_ECUDA;
            finalBuf = pending;
            return 1;
        }
        nextCheck = nextCheck < every ? (2 * nextCheck < every ? 2 * nextCheck : every) : nextCheck + every;
        return 0;
    }
};

static int bicgstab_solve(fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs, const double* values, const double* b,
                          double* x, fvk_solver_stats* stats_h, double* history_h, int32_t maxHistory, fvk_stream s)
{
    cudaStream_t st = fvk_cu(s);
    const int n = sv->nRows;
    const bool dist = sv->comm != nullptr;
    const int dm = dist ? 1 : 0;
    const bool jacobi = sv->cfg.preconditioner == FVK_PRECOND_JACOBI;
    const int gV = stream_grid(n), gS = spmv_grid(n);
    const int wantHist = (history_h && maxHistory > 0) ? (maxHistory < sv->histCap ? maxHistory : sv->histCap) : 0;
    double *r = sv->r, *rr = sv->rr, *p = sv->p1, *v = sv->q, *sB = sv->sB, *tB = sv->tB, *y = sv->p0, *z = sv->z;

    PcgState init;
    std::memset(&init, 0, sizeof(init));
    init.rho = init.rhoPrev = init.alpha = init.beta = init.gamma = init.omega = 1.0;
    init.relTol = sv->cfg.relTol; init.absTol = sv->cfg.absTol;
    init.maxIter = sv->cfg.maxIter; init.maxHist = wantHist;
    *sv->state_h = init;
    FVK_CUDA(cudaMemcpyAsync(sv->state, sv->state_h, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    FVK_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * sv->nCols, st));
    FVK_CUDA(cudaMemsetAsync(v, 0, sizeof(double) * n, st));
    if (jacobi)
        if (int rc = launch_dinv(sv, rowOffs, colIdxs, values, st)) return rc;
    if (dist)
        if (int rc = fvk_comm_halo_exchange_impl(sv->comm, x, 1, st)) return rc;
    // r = b - A x and ||b||^2 in one pass
    k_spmv<2><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, x, b, r, sv->state, nullptr, nullptr, sv->partial, sv->counter, dm, nullptr, 0, sv->aff);
    FVK_LAUNCH_CHECK();
    if (dist)
    {
        if (int rc = fvk_comm_allreduce_sum_impl(sv->comm, &sv->state->sums[3], 1, st)) return rc;
        k_set_normB<<<1, 1, 0, st>>>(sv->state);
    }
    auto reduce = [&](int first, int count) -> int { return dist ? fvk_comm_allreduce_sum_impl(sv->comm, &sv->state->sums[first], count, st) : FVK_OK; };
    k_bicg_step3<true><<<gV, TB, 0, st>>>(n, sv->state, x, y, z, sB, tB, r, rr, sv->partial, sv->counter, sv->hist, dm);
    FVK_LAUNCH_CHECK();
    if (dist)
    {
        if (int rc = reduce(0, 2)) return rc;
        k_bicg_decide<0><<<1, 1, 0, st>>>(sv->state, sv->hist);
    }
    StopPoll poll {sv, st, sv->cfg.checkEvery};
    // round k runs the body of iteration k and the first stopping check of iteration k + 1
    const int lastRound = sv->cfg.maxIter > 0 ? sv->cfg.maxIter - 1 : 0;
    for (int it = 0; it <= lastRound; ++it)
    {
        if (jacobi) k_bicg_step1<true><<<gV, TB, 0, st>>>(n, sv->state, r, p, v, sv->dinv, y);
        else k_bicg_step1<false><<<gV, TB, 0, st>>>(n, sv->state, r, p, v, sv->dinv, y);
        FVK_LAUNCH_CHECK();
        if (dist) if (int rc = fvk_comm_halo_exchange_impl(sv->comm, y, 1, st)) return rc;
        k_spmv<5><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, y, rr, v, sv->state, nullptr, nullptr, sv->partial, sv->counter, dm, nullptr, 0, sv->aff);
        FVK_LAUNCH_CHECK();
        if (dist)
        {
            if (int rc = reduce(2, 1)) return rc;
            k_bicg_decide<1><<<1, 1, 0, st>>>(sv->state, sv->hist);
        }
        if (jacobi) k_bicg_step2<true><<<gV, TB, 0, st>>>(n, sv->state, r, v, sv->dinv, sB, z, sv->partial, sv->counter, sv->hist, dm);
        else k_bicg_step2<false><<<gV, TB, 0, st>>>(n, sv->state, r, v, sv->dinv, sB, z, sv->partial, sv->counter, sv->hist, dm);
        FVK_LAUNCH_CHECK();
        if (dist)
        {
            if (int rc = reduce(1, 1)) return rc;
            k_bicg_decide<2><<<1, 1, 0, st>>>(sv->state, sv->hist);
            if (int rc = fvk_comm_halo_exchange_impl(sv->comm, z, 1, st)) return rc;
        }
        k_spmv<6><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, z, sB, tB, sv->state, nullptr, nullptr, sv->partial, sv->counter, dm, nullptr, 0, sv->aff);
        FVK_LAUNCH_CHECK();
        if (dist)
        {
            if (int rc = reduce(2, 2)) return rc;
            k_bicg_decide<3><<<1, 1, 0, st>>>(sv->state, sv->hist);
        }
        k_bicg_step3<false><<<gV, TB, 0, st>>>(n, sv->state, x, y, z, sB, tB, r, rr, sv->partial, sv->counter, sv->hist, dm);
        FVK_LAUNCH_CHECK();
        if (dist)
        {
            if (int rc = reduce(0, 2)) return rc;
            k_bicg_decide<4><<<1, 1, 0, st>>>(sv->state, sv->hist);
        }
        const int pr = poll.after_round(it, lastRound);
        if (pr < 0) return fvk_fail(FVK_ECUDA, "fvk_solver_solve: stop polling failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (pr == 1) break;
    }
    const PcgState& fin = sv->state_h[poll.finalBuf];
    if (!fin.done) return fvk_fail(FVK_ECUDA, "fvk_solver_solve: stop flag not raised after maxIter rounds");
    stats_h->numIter = fin.iter;
    stats_h->initResNorm = fin.normB;
    stats_h->finalResNorm = fin.normR;
    stats_h->nHistory = fin.nHist;
    if (wantHist && fin.nHist > 0)
    {
        FVK_CUDA(cudaMemcpyAsync(history_h, sv->hist, sizeof(double) * fin.nHist, cudaMemcpyDeviceToHost, st));
        FVK_CUDA(cudaStreamSynchronize(st));
    }
    return FVK_OK;
}


// ---- CG inside a stream capture: one conditional WHILE node, zero host round trips ---------------------------------------
constexpr int FVK_MAX_CAPTURED_SOLVES = 64; // = the pinned slots allocated by fvk_solver_create
__global__ void k_loop_cond(const PcgState* __restrict__ st, cudaGraphConditionalHandle h)
{
    if (st->done) cudaGraphSetConditional(h, 0);
}
constexpr unsigned FVK_CAPTURE_LOG = 8192;
__global__ void k_log_captured(const PcgState* __restrict__ st, int32_t* __restrict__ log, unsigned* __restrict__ count)
{
    const unsigned i = (*count)++;
    if (i < FVK_CAPTURE_LOG) log[i] = st->iter;
}

static int cg_solve_captured(fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs, const double* values, const double* b,
                             double* x, fvk_solver_stats* stats_h, cudaStream_t st)
{
    const int n = sv->nRows;
    const bool dist = sv->comm != nullptr;
    const FvkP2PCtx* p2p = fvk_comm_p2p_ctx(sv->comm);
    if (dist && !p2p) return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: stream capture needs the peer-memory transport (or one GPU)");
    if (sv->nCaptured >= FVK_MAX_CAPTURED_SOLVES) return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: more than %d captured solves", FVK_MAX_CAPTURED_SOLVES);
    const int dmode = dist ? 2 : 0;
    const bool jacobi = sv->cfg.preconditioner == FVK_PRECOND_JACOBI;
    const int gV = stream_grid(n), gS = spmv_grid(n);
    if (!sv->init_h || !sv->cap_h || !sv->bodyStream) return fvk_fail(FVK_ECUDA, "fvk_solver_solve: capture buffers missing");
    PcgState init;
    std::memset(&init, 0, sizeof(init));
    init.rhoPrev = 1.0;
    init.relTol = sv->cfg.relTol; init.absTol = sv->cfg.absTol; init.maxIter = sv->cfg.maxIter;
    *sv->init_h = init; // identical for every captured solve of this solver
    // the graph being captured, for the conditional handle
    cudaStreamCaptureStatus status;
    cudaGraph_t graph = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t nDeps = 0;
    FVK_CUDA(cudaStreamGetCaptureInfo(st, &status, nullptr, &graph, &deps, &nDeps));
    cudaGraphConditionalHandle handle;
    FVK_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
    // ---- start-up (same kernels as the eager path)
    FVK_CUDA(cudaMemcpyAsync(sv->state, sv->init_h, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    FVK_CUDA(cudaMemsetAsync(sv->p0, 0, sizeof(double) * sv->nCols, st));
    const bool fuseDinv = jacobi && dinv_fusable(sv, rowOffs, colIdxs);
    if (jacobi && !fuseDinv)
        if (int rc = launch_dinv(sv, rowOffs, colIdxs, values, st)) return rc;
    if (dist && !sv->guessGhostsCurrent)
        if (int rc = fvk_comm_halo_exchange_impl(sv->comm, x, 1, st)) return rc;
    // distributed: ||b||^2 is all-reduced together with (r.z, r.r) inside the first update kernel
    k_spmv<2><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, x, b, sv->r, sv->state, nullptr, nullptr, sv->partial, sv->counter, dist ? 1 : 0, nullptr, 0, sv->aff,
                                 fuseDinv ? sv->affDiagOffs : nullptr, fuseDinv ? sv->dinv : nullptr);
    FVK_LAUNCH_CHECK();
    double *rA = sv->r, *rB = dmode == 2 ? sv->r2 : sv->r;
    auto K1 = [&](cudaStream_t q, bool first, double* rIn, double* rOut, double* pCur) {
        if (first)
        {
            if (jacobi) k_cg_update<true, true><<<gV, TB, 0, q>>>(n, sv->state, x, pCur, sv->q, rIn, rIn, sv->dinv, sv->z, sv->partial, sv->counter, nullptr, dmode, p2p);
            else k_cg_update<true, false><<<gV, TB, 0, q>>>(n, sv->state, x, pCur, sv->q, rIn, rIn, sv->dinv, sv->z, sv->partial, sv->counter, nullptr, dmode, p2p);
        }
        else
        {
            if (jacobi) k_cg_update<false, true><<<gV, TB, 0, q>>>(n, sv->state, x, pCur, sv->q, rIn, rOut, sv->dinv, sv->z, sv->partial, sv->counter, nullptr, dmode, p2p);
            else k_cg_update<false, false><<<gV, TB, 0, q>>>(n, sv->state, x, pCur, sv->q, rIn, rOut, sv->dinv, sv->z, sv->partial, sv->counter, nullptr, dmode, p2p);
        }
    };
    auto K2 = [&](cudaStream_t q, double* pCur, double* pNext) {
        k_spmv<4><<<gS, TB, 0, q>>>(n, rowOffs, colIdxs, values, pCur, nullptr, sv->q, sv->state, sv->z, pNext, sv->partial, sv->counter, dmode, p2p, sv->nCols, sv->aff);
    };
    K1(st, true, rA, rA, sv->p0);
    K2(st, sv->p0, sv->p1);
    k_loop_cond<<<1, 1, 0, st>>>(sv->state, handle);
    FVK_LAUNCH_CHECK();
    // ---- WHILE (not done): two iterations per pass, so the double-buffered p (and r) are back where they started
    FVK_CUDA(cudaStreamGetCaptureInfo(st, &status, nullptr, &graph, &deps, &nDeps));
    cudaGraphNodeParams cp = {};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t node;
    FVK_CUDA(cudaGraphAddNode(&node, graph, deps, nDeps, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    FVK_CUDA(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
    FVK_CUDA(cudaStreamBeginCaptureToGraph(sv->bodyStream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    // iterations per pass: an even number (the double-buffered p and r are back where they started); once the stop flag is up
    // the rest of a pass is no-op kernels, so longer passes trade fewer loop turns against a few idle launches at the end
    static const int passPairs = [] { const char* e = std::getenv("FVK_CG_PASS_ITERS"); const int v = e ? std::atoi(e) : 2; return v >= 2 && v <= 16 ? v / 2 : 1; }();
    for (int pp = 0; pp < passPairs; ++pp)
    {
        K1(sv->bodyStream, false, rA, rB, sv->p1);
        K2(sv->bodyStream, sv->p1, sv->p0);
        K1(sv->bodyStream, false, rB, rA, sv->p0);
        K2(sv->bodyStream, sv->p0, sv->p1);
    }
    k_loop_cond<<<1, 1, 0, sv->bodyStream>>>(sv->state, handle);
    cudaError_t le = cudaGetLastError();
    cudaGraph_t bodyOut = nullptr;
    cudaError_t ee = cudaStreamEndCapture(sv->bodyStream, &bodyOut);
    if (le != cudaSuccess || ee != cudaSuccess)
        return fvk_fail(FVK_ECUDA, "fvk_solver_solve: capturing the iteration body failed: %s", cudaGetErrorString(le != cudaSuccess ? le : ee));
    // ---- iteration count into the device log, final state into this solve's pinned slot
    k_log_captured<<<1, 1, 0, st>>>(sv->state, sv->capLog, sv->capLogCount);
    FVK_LAUNCH_CHECK();
    const int slot = sv->nCaptured++;
    FVK_CUDA(cudaMemcpyAsync(&sv->cap_h[slot], sv->state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
    stats_h->numIter = -(slot + 1); // not known at capture time: fvk_solver_captured_stats(slot) after a replay
    stats_h->initResNorm = stats_h->finalResNorm = 0.0;
    stats_h->nHistory = 0;
    return FVK_OK;
}

extern "C" int fvk_solver_solve(fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs, const double* values,
                                const double* b, double* x, fvk_solver_stats* stats_h, double* history_h,
                                int32_t maxHistory, fvk_stream s)
{
    if (!sv || !rowOffs || !colIdxs || !values || !b || !x || !stats_h) return fvk_fail(FVK_EINVAL, "fvk_solver_solve: null argument");
    // the structured SpMV computes columns from the attached mesh's dimensions: only valid for that mesh's own pattern
    struct AffGuard { fvk_solver* s; SpmvAffine saved; ~AffGuard() { s->aff = saved; } } guard {sv, sv->aff};
    if (rowOffs != sv->affRowOffs || colIdxs != sv->affColIdxs) sv->aff = SpmvAffine {0, 0, 0, 0};
    cudaStream_t st = fvk_cu(s);
    {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (st != nullptr && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive)
        {
            if (sv->cfg.solverType != FVK_SOLVER_CG) return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: only solver::Cg can be captured into a CUDA graph");
            if (history_h) return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: residual histories are not available inside a stream capture");
            if (sv->cfg.preconditioner == FVK_PRECOND_DIC) return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: the DIC preconditioner cannot be captured yet");
            return cg_solve_captured(sv, rowOffs, colIdxs, values, b, x, stats_h, st);
        }
    }
    if (sv->cfg.solverType == FVK_SOLVER_BICGSTAB) return bicgstab_solve(sv, rowOffs, colIdxs, values, b, x, stats_h, history_h, maxHistory, s);
    const int n = sv->nRows;
    const bool dist = sv->comm != nullptr;
    const FvkP2PCtx* p2p = fvk_comm_p2p_ctx(sv->comm);
    const int dmode = !dist ? 0 : (p2p ? 2 : 1); // 2: all-reduces fused into the kernels over peer memory
    const bool jacobi = sv->cfg.preconditioner == FVK_PRECOND_JACOBI;
    const int gV = stream_grid(n), gS = spmv_grid(n);
    const int wantHist = (history_h && maxHistory > 0) ? (maxHistory < sv->histCap ? maxHistory : sv->histCap) : 0;

    PcgState init;
    std::memset(&init, 0, sizeof(init));
    init.rhoPrev = 1.0;
    init.relTol = sv->cfg.relTol; init.absTol = sv->cfg.absTol;
    init.maxIter = sv->cfg.maxIter; init.maxHist = wantHist;
    *sv->state_h = init;
    FVK_CUDA(cudaMemcpyAsync(sv->state, sv->state_h, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    // p of "iteration -1" is zero (p = z + beta p); p1 and q are fully written before they are read
    FVK_CUDA(cudaMemsetAsync(sv->p0, 0, sizeof(double) * sv->nCols, st));
    const bool fuseDinv = jacobi && dinv_fusable(sv, rowOffs, colIdxs);
    if (jacobi && !fuseDinv)
        if (int rc = launch_dinv(sv, rowOffs, colIdxs, values, st)) return rc;
    const bool dic = sv->cfg.preconditioner == FVK_PRECOND_DIC;
    const fvk_mesh* cm = sv->attached;
    if (dic)
    { // D* colour by colour (sv->dinv holds 1 / D*)
        if (!cm || rowOffs != sv->affRowOffs || colIdxs != sv->affColIdxs)
            return fvk_fail(FVK_EUNSUPPORTED, "fvk_solver_solve: the DIC preconditioner needs fvk_solver_attach_mesh and that mesh's own sparsity pattern");
        if (int rc = fvk_mesh_ensure_colors(cm)) return rc;
        for (int c = 0; c < cm->dicNColors; ++c)
        {
            const int n0 = cm->dicOff[c], n1 = cm->dicOff[c + 1];
            if (n1 > n0) k_dic_diag<<<(n1 - n0 + TB - 1) / TB, TB, 0, st>>>(n0, n1, cm->dicCells, rowOffs, colIdxs, values, cm->dicColor, c, n, sv->dinv);
        }
        FVK_LAUNCH_CHECK();
    }
    auto dic_apply = [&](const double* rr) -> int {
        for (int c = 0; c < cm->dicNColors; ++c)
        {
            const int n0 = cm->dicOff[c], n1 = cm->dicOff[c + 1];
            if (n1 > n0) k_dic_fwd<<<(n1 - n0 + TB - 1) / TB, TB, 0, st>>>(n0, n1, sv->state, cm->dicCells, rowOffs, colIdxs, values, cm->dicColor, c, n, sv->dinv, rr, sv->z);
        }
        for (int c = cm->dicNColors - 2; c >= 0; --c) // the highest colour has no higher neighbours
        {
            const int n0 = cm->dicOff[c], n1 = cm->dicOff[c + 1];
            if (n1 > n0) k_dic_bwd<<<(n1 - n0 + TB - 1) / TB, TB, 0, st>>>(n0, n1, sv->state, cm->dicCells, rowOffs, colIdxs, values, cm->dicColor, c, n, sv->dinv, sv->z);
        }
        k_cg_rz_decide<<<gV, TB, 0, st>>>(n, sv->state, rr, sv->z, sv->partial, sv->counter, sv->hist);
        FVK_LAUNCH_CHECK();
        return FVK_OK;
    };
    if (dist && !sv->guessGhostsCurrent)
        if (int rc = fvk_comm_halo_exchange_impl(sv->comm, x, 1, st)) return rc;
    // r = b - A x, fused with ||b||^2 (the reference's "initial residual" is ||b||, ginkgo.hpp:143-144)
    k_spmv<2><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, x, b, sv->r, sv->state, nullptr, nullptr, sv->partial, sv->counter, dist ? 1 : 0, nullptr, 0, sv->aff,
                                 fuseDinv ? sv->affDiagOffs : nullptr, fuseDinv ? sv->dinv : nullptr);
    FVK_LAUNCH_CHECK();
    if (dmode == 1)
    { // NCCL transport; the peer-memory transport all-reduces ||b||^2 inside the first update kernel
        if (int rc = fvk_comm_allreduce_sum_impl(sv->comm, &sv->state->sums[3], 1, st)) return rc;
        k_set_normB<<<1, 1, 0, st>>>(sv->state);
    }

    double* pCur = sv->p0;  // p of the previous iteration
    double* pNext = sv->p1;
    double* rIn = sv->r;    // peer-memory mode double-buffers r (k_cg_update re-reads its input for the halo cells)
    double* rOut = dmode == 2 ? sv->r2 : sv->r;
    const int every = sv->cfg.checkEvery;
    int nextCheck = 1, pending = -1, finalBuf = 0;
    // FVK_CG_TIMING=1: CUDA events around the two kernels of the first 64 iterations, averages to stderr
    static const bool timing = [] { const char* e = std::getenv("FVK_CG_TIMING"); return e && *e == '1'; }();
    constexpr int NT = 64;
    cudaEvent_t tev[NT][3];
    if (timing)
        for (auto& e3 : tev)
            for (auto& e : e3) cudaEventCreate(&e);
    int timed = 0;
    for (int it = 0; it <= sv->cfg.maxIter; ++it)
    {
        if (timing && it >= 1 && it <= NT) cudaEventRecord(tev[it - 1][0], st);
        // K1
        if (dic)
        {
            if (it == 0) k_cg_update<true, 2><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rIn, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
            else k_cg_update<false, 2><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rOut, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
            FVK_LAUNCH_CHECK();
            if (int rc = dic_apply(it == 0 ? rIn : rOut)) return rc;
        }
        else if (it == 0)
        {
            if (jacobi) k_cg_update<true, true><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rIn, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
            else k_cg_update<true, false><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rIn, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
        }
        else
        {
            if (jacobi) k_cg_update<false, true><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rOut, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
            else k_cg_update<false, false><<<gV, TB, 0, st>>>(n, sv->state, x, pCur, sv->q, rIn, rOut, sv->dinv, sv->z, sv->partial, sv->counter, sv->hist, dmode, p2p);
        }
        FVK_LAUNCH_CHECK();
        if (timing && it >= 1 && it <= NT) cudaEventRecord(tev[it - 1][1], st);
        if (it > 0 && dmode == 2) { double* t = rIn; rIn = rOut; rOut = t; }
        if (dmode == 2)
        {
            // peer-memory CG: the same two kernels as on one GPU. K1's last block exchanged the halo of z and all-reduced
            // (r.z, r.r); K2 forms p = z + beta p on the fly for owned AND ghost columns and all-reduces p.q.
            k_spmv<4><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, pCur, nullptr, sv->q, sv->state, sv->z, pNext, sv->partial, sv->counter, 2, p2p, sv->nCols, sv->aff);
            FVK_LAUNCH_CHECK();
            double* t = pCur; pCur = pNext; pNext = t;
        }
        else if (dist)
        {
            if (dmode == 1)
            {
                if (int rc = fvk_comm_allreduce_sum_impl(sv->comm, &sv->state->sums[0], 2, st)) return rc;
                k_decide_after_update<<<1, 1, 0, st>>>(sv->state, sv->hist);
            }
            k_cg_pupdate<<<gV, TB, 0, st>>>(n, sv->state, sv->z, pCur);
            if (int rc = fvk_comm_halo_exchange_impl(sv->comm, pCur, 1, st)) return rc;
            k_spmv<3><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, pCur, nullptr, sv->q, sv->state, nullptr, nullptr, sv->partial, sv->counter, dmode, p2p, 0, sv->aff);
            FVK_LAUNCH_CHECK();
            if (dmode == 1)
            {
                if (int rc = fvk_comm_allreduce_sum_impl(sv->comm, &sv->state->sums[2], 1, st)) return rc;
                k_decide_after_spmv<<<1, 1, 0, st>>>(sv->state);
            }
        }
        else
        {
            // K2 (fused): pNext = z + beta pCur, q = A pNext, p.q, alpha
            k_spmv<4><<<gS, TB, 0, st>>>(n, rowOffs, colIdxs, values, pCur, nullptr, sv->q, sv->state, sv->z, pNext, sv->partial, sv->counter, 0, nullptr, 0, sv->aff);
            FVK_LAUNCH_CHECK();
            double* t = pCur; pCur = pNext; pNext = t;
        }
        if (timing && it >= 1 && it <= NT) { cudaEventRecord(tev[it - 1][2], st); timed = it; }
        // stop check after 1, 2, 4, ... iterations up to `every`, then every `every`: a solve that converges at once (the
        // second PISO corrector often needs 0-3 iterations) does not queue `every` pairs of no-op kernels first
        // The check is pipelined: the state copy of batch k is awaited only after batch k+1 has been queued, so the GPU
        // never waits for the host; at most one batch of (no-op) kernels is queued beyond the stop.
        if (it + 1 == nextCheck || it == sv->cfg.maxIter)
        {
            if (pending >= 0)
            {
                FVK_CUDA(cudaEventSynchronize(sv->checkEv[pending]));
                if (sv->state_h[pending].done) { finalBuf = pending; break; }
            }
            pending = pending < 0 ? 0 : 1 - pending;
            FVK_CUDA(cudaMemcpyAsync(&sv->state_h[pending], sv->state, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
            FVK_CUDA(cudaEventRecord(sv->checkEv[pending], st));
            if (it == sv->cfg.maxIter)
            {
                FVK_CUDA(cudaEventSynchronize(sv->checkEv[pending]));
                finalBuf = pending;
                break;
            }
            nextCheck = nextCheck < every ? (2 * nextCheck < every ? 2 * nextCheck : every) : nextCheck + every;
        }
    }
    if (timing)
    {
        cudaStreamSynchronize(st);
        double a = 0, b = 0;
        for (int i = 0; i < timed; ++i)
        {
            float x1 = 0, x2 = 0;
            cudaEventElapsedTime(&x1, tev[i][0], tev[i][1]);
            cudaEventElapsedTime(&x2, tev[i][1], tev[i][2]);
            a += x1; b += x2;
        }
        if (timed) std::fprintf(stderr, "[fvk cg timing] mode %d: update phase %.1f us, spmv phase %.1f us (avg of %d iterations)\n", dmode, a / timed * 1e3, b / timed * 1e3, timed);
        for (auto& e3 : tev)
            for (auto& e : e3) cudaEventDestroy(e);
    }
    const PcgState& fin = sv->state_h[finalBuf];
    if (!fin.done) return fvk_fail(FVK_ECUDA, "fvk_solver_solve: stop flag not raised after maxIter+1 checks");
    stats_h->numIter = fin.iter;
    stats_h->initResNorm = fin.normB;
    stats_h->finalResNorm = fin.normR;
    stats_h->nHistory = fin.nHist;
    if (wantHist && fin.nHist > 0)
    {
        FVK_CUDA(cudaMemcpyAsync(history_h, sv->hist, sizeof(double) * fin.nHist, cudaMemcpyDeviceToHost, st));
        FVK_CUDA(cudaStreamSynchronize(st));
    }
    return FVK_OK;
}

// Vec3 LinearSystem (values Vec3[nnz] with identical components, rhs / x Vec3): three scalar solves over the component
// matrix, like OpenFOAM's segregated vector solve. NeoN's la::Solver has no Vec3 overload (solver.hpp:52 is commented out);
// this is the solve `momentumPredictor yes` (neoIcoFoam.cpp:100-103) needs.
extern "C" int fvk_solver_solve_vec3(fvk_solver* sv, int64_t nnz, const int32_t* rowOffs, const int32_t* colIdxs, const double* valuesV,
                                     const double* bV, double* xV, fvk_solver_stats* stats3_h, fvk_stream s)
{
    if (!sv || nnz <= 0 || !rowOffs || !colIdxs || !valuesV || !bV || !xV || !stats3_h) return fvk_fail(FVK_EINVAL, "fvk_solver_solve_vec3: bad argument");
    cudaStream_t st = fvk_cu(s);
    if (sv->vals0Cap < nnz)
    {
        if (sv->vals0) cudaFree(sv->vals0);
        sv->vals0 = nullptr; sv->vals0Cap = 0;
        FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&sv->vals0), sizeof(double) * nnz));
        sv->vals0Cap = nnz;
    }
    k_take_component<<<stream_grid(nnz), TB, 0, st>>>(nnz, 0, valuesV, sv->vals0);
    FVK_LAUNCH_CHECK();
    return fvk_solver_solve_vec3c(sv, rowOffs, colIdxs, sv->vals0, bV, xV, stats3_h, s);
}

extern "C" int fvk_solver_solve_vec3c(fvk_solver* sv, const int32_t* rowOffs, const int32_t* colIdxs, const double* valuesCompact,
                                      const double* bV, double* xV, fvk_solver_stats* stats3_h, fvk_stream s)
{
    if (!sv || !rowOffs || !colIdxs || !valuesCompact || !bV || !xV || !stats3_h) return fvk_fail(FVK_EINVAL, "fvk_solver_solve_vec3c: bad argument");
    cudaStream_t st = fvk_cu(s);
    if (!sv->bC) FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&sv->bC), sizeof(double) * sv->nRows));
    if (!sv->xC) FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&sv->xC), sizeof(double) * sv->nCols));
    for (int c = 0; c < 3; ++c)
    {
        k_take_component<<<stream_grid(sv->nRows), TB, 0, st>>>(sv->nRows, c, bV, sv->bC);
        k_take_component<<<stream_grid(sv->nCols), TB, 0, st>>>(sv->nCols, c, xV, sv->xC);
        FVK_LAUNCH_CHECK();
        if (int rc = fvk_solver_solve(sv, rowOffs, colIdxs, valuesCompact, sv->bC, sv->xC, &stats3_h[c], nullptr, 0, s)) return rc;
        k_put_component<<<stream_grid(sv->nRows), TB, 0, st>>>(sv->nRows, c, sv->xC, xV);
        FVK_LAUNCH_CHECK();
    }
    return FVK_OK;
}

// Stream-capture mode: statistics of captured solve `slot` (the order of the fvk_solver_solve calls during the capture, as
// returned in stats.numIter = -(slot + 1)), valid after a replay of the graph has completed. fvk_solver_reset_captures
// forgets the slots (before capturing again).
extern "C" int fvk_solver_captured_stats(const fvk_solver* sv, int32_t slot, fvk_solver_stats* stats_h)
{
    if (!sv || !stats_h || slot < 0 || slot >= sv->nCaptured || !sv->cap_h) return fvk_fail(FVK_EINVAL, "fvk_solver_captured_stats: bad slot");
    const PcgState& fin = sv->cap_h[slot];
    stats_h->numIter = fin.iter;
    stats_h->initResNorm = fin.normB;
    stats_h->finalResNorm = fin.normR;
    stats_h->nHistory = 0;
    return fin.done ? FVK_OK : fvk_fail(FVK_ECUDA, "fvk_solver_captured_stats: the solve has not finished (replay not complete?)");
}
// iteration counts of the captured solves executed since the last call, in execution order (synchronises the device)
extern "C" int fvk_solver_captured_log(fvk_solver* sv, int32_t* out_h, int32_t capacity, int32_t* n_h)
{
    if (!sv || !n_h || capacity < 0 || (capacity && !out_h)) return fvk_fail(FVK_EINVAL, "fvk_solver_captured_log: bad argument");
    FVK_CUDA(cudaDeviceSynchronize());
    unsigned cnt = 0;
    FVK_CUDA(cudaMemcpy(&cnt, sv->capLogCount, sizeof(unsigned), cudaMemcpyDeviceToHost));
    const unsigned take = std::min<unsigned>(std::min<unsigned>(cnt, FVK_CAPTURE_LOG), unsigned(capacity));
    if (take) FVK_CUDA(cudaMemcpy(out_h, sv->capLog, sizeof(int32_t) * take, cudaMemcpyDeviceToHost));
    FVK_CUDA(cudaMemset(sv->capLogCount, 0, sizeof(unsigned)));
    *n_h = int32_t(take);
    return FVK_OK;
}
extern "C" int fvk_solver_set_ghosts_current(fvk_solver* sv, int32_t on)
{
    if (!sv) return fvk_fail(FVK_EINVAL, "fvk_solver_set_ghosts_current: null");
    sv->guessGhostsCurrent = on != 0;
    return FVK_OK;
}
extern "C" int fvk_solver_keeps_ghosts(const fvk_solver* sv, int32_t* out_h)
{
    if (!sv || !out_h) return fvk_fail(FVK_EINVAL, "fvk_solver_keeps_ghosts: null");
    *out_h = (!sv->comm || (sv->cfg.solverType == FVK_SOLVER_CG && fvk_comm_p2p_ctx(sv->comm))) ? 1 : 0;
    return FVK_OK;
}
extern "C" int fvk_solver_reset_captures(fvk_solver* sv)
{
    if (!sv) return fvk_fail(FVK_EINVAL, "fvk_solver_reset_captures: null");
    sv->nCaptured = 0;
    return FVK_OK;
}
