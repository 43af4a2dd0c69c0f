// Explicit finite-volume operators: fused, deterministic, cell-centric gather kernels for sm_100a.
//
// The reference computes each operator as (1) a face loop writing a face-sized temporary phi_f,
// (2) a face loop scattering flux to owner/neighbour with atomics, (3) a cell loop scaling by
// coeff/V (src/NeoN/src/finiteVolume/cellCentred/operators/gaussGreen{Div,Grad,Laplacian}.cpp,
// surfaceIntegrate.cpp, interpolation/{linear,upwind}.cpp, faceNormalGradient/uncorrected.cpp).
// Here one kernel per operator walks, for every cell, its faces in ascending face id (the
// summation order of the reference's SerialExecutor), recomputes the face value on the fly, and
// writes the scaled result once. No atomics, no temporaries, bit-reproducible.
//
// Per-face arithmetic is written exactly as in the reference and the library is compiled with
// -fmad=false, so results are bit-identical to the Serial executor built without FP contraction.
#include "fvk_device.cuh"

#include <cstdio>
#include <cstdlib>

namespace
{

// ---- value-type helpers ------------------------------------------------------------------------
struct S1 // scalar field
{
    using T = double;
    static constexpr int NC = 1;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T ld(const double* __restrict__ p, int64_t i) { return p[i]; }
    static __device__ __forceinline__ void st(double* __restrict__ p, int64_t i, T v) { p[i] = v; }
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T mul(double s, T a) { return s * a; } // scalar * value
};
struct S3 // Vec3 field, AoS
{
    using T = Vec3d;
    static constexpr int NC = 3;
    static __device__ __forceinline__ T zero() { return Vec3d {0.0, 0.0, 0.0}; }
    static __device__ __forceinline__ T ld(const double* __restrict__ p, int64_t i) { return ld3(p, i); }
    static __device__ __forceinline__ void st(double* __restrict__ p, int64_t i, T v) { st3(p, i, v); }
    static __device__ __forceinline__ T add(T a, T b) { return Vec3d {a.x + b.x, a.y + b.y, a.z + b.z}; }
    static __device__ __forceinline__ T sub(T a, T b) { return Vec3d {a.x - b.x, a.y - b.y, a.z - b.z}; }
    // NeoN: operator*(scalar, Vec3) is rhs *= s  ->  component * s (vec3.hpp)
    static __device__ __forceinline__ T mul(double s, T a) { return Vec3d {a.x * s, a.y * s, a.z * s}; }
};

// ---- per-face flux functors --------------------------------------------------------------------
// internal(f, own, nei) is the value the reference adds to res[own] and subtracts from res[nei];
// boundary(f, b, own) is the value added to res[own] for boundary face f = nI + b.

// Split interface used by the batched kernel: cell(c) loads what the thread's own cell contributes,
// load(f, other) issues the raw loads of one internal face (nothing else), flux(face, cell, side) is the
// signed value to ADD to the cell's sum: side 0 = the cell owns the face (+flux with own = cell,
// nei = other), side 1 = the cell is the neighbour (-flux with own = other, nei = cell).
template <class VT, int SCHEME>
struct DivOp // gaussGreenDiv.cpp:46-67 with linear.cpp:30-45 / upwind.cpp:32-55
{
    using V = VT;
    using T = typename VT::T;
    const double* __restrict__ faceFlux;
    const double* __restrict__ w;
    const double* __restrict__ phi;
    const double* __restrict__ phiB;
    // staged-stream interface of the tile kernel: stream 0 = faceFlux, stream 1 = weights (linear only)
    using CV = VT;
    static constexpr int W0 = 1, W1 = (SCHEME == FVK_LINEAR) ? 1 : 0;
    static constexpr bool NEEDS_BLEND = true; // w*own + (1-w)*nei: more live registers than a difference
    __device__ __forceinline__ bool ownsStreams() const { return false; } // stream 0 is the caller's faceFlux
    __host__ __device__ __forceinline__ const double* s0() const { return faceFlux; }
    __host__ __device__ __forceinline__ const double* s1() const { return w; }
    __host__ __device__ __forceinline__ const double* cells() const { return phi; }
    __device__ __forceinline__ T fluxv(const double* a, const double* b, const T& pOwn, const T& pNei) const
    {
        const double F = a[0];
        T phif;
        if (SCHEME == FVK_LINEAR)
        {
            const double wf = b[0];
            phif = VT::add(VT::mul(wf, pOwn), VT::mul(1 - wf, pNei));
        }
        else
            phif = (F >= 0) ? pOwn : pNei;
        return VT::mul(F, phif);
    }
    struct Face { double F, w; T pO; };
    __device__ __forceinline__ T cell(int c) const { return VT::ld(phi, c); }
    __device__ __forceinline__ Face load(int f, int other) const
    {
        Face r;
        r.F = faceFlux[f];
        r.w = (SCHEME == FVK_LINEAR) ? w[f] : 0.0;
        r.pO = VT::ld(phi, other);
        return r;
    }
    __device__ __forceinline__ T flux(const Face& fd, const T& pc, bool side) const
    {
        const T pOwn = side ? fd.pO : pc, pNei = side ? pc : fd.pO;
        T phif;
        if (SCHEME == FVK_LINEAR)
            phif = VT::add(VT::mul(fd.w, pOwn), VT::mul(1 - fd.w, pNei));
        else
            phif = (fd.F >= 0) ? pOwn : pNei;
        const T fl = VT::mul(fd.F, phif);
        return side ? VT::sub(VT::zero(), fl) : fl;
    }
    __device__ __forceinline__ T internal(int f, int own, int nei) const
    {
        const double F = faceFlux[f];
        T phif;
        if (SCHEME == FVK_LINEAR)
        {
            const double wf = w[f];
            phif = VT::add(VT::mul(wf, VT::ld(phi, own)), VT::mul(1 - wf, VT::ld(phi, nei)));
        }
        else
        {
            phif = (F >= 0) ? VT::ld(phi, own) : VT::ld(phi, nei);
        }
        return VT::mul(F, phif);
    }
    __device__ __forceinline__ T boundary(int f, int b, int) const
    {
        // phif = weights[f] * bvalue[b]; the geometric boundary weight is exactly 1
        return VT::mul(faceFlux[f], VT::ld(phiB, b));
    }
};

struct GradOp // gaussGreenGrad.cpp:46-64 with linear.cpp:30-45
{
    using V = S3;
    using T = Vec3d;
    const double* __restrict__ Sf;
    const double* __restrict__ w;
    const double* __restrict__ phi;
    const double* __restrict__ phiB;
    using CV = S1;
    static constexpr int W0 = 3, W1 = 1; // stream 0 = Sf (Vec3), stream 1 = weights
    static constexpr bool NEEDS_BLEND = true;
    __device__ __forceinline__ bool ownsStreams() const { return true; }
    __host__ __device__ __forceinline__ const double* s0() const { return Sf; }
    __host__ __device__ __forceinline__ const double* s1() const { return w; }
    __host__ __device__ __forceinline__ const double* cells() const { return phi; }
    __device__ __forceinline__ T fluxv(const double* a, const double* b, double pOwn, double pNei) const
    {
        const double wf = b[0];
        const double phif = wf * pOwn + (1 - wf) * pNei;
        return Vec3d {a[0] * phif, a[1] * phif, a[2] * phif};
    }
    struct Face { Vec3d s; double w, pO; };
    __device__ __forceinline__ double cell(int c) const { return phi[c]; }
    __device__ __forceinline__ Face load(int f, int other) const { return Face {ld3(Sf, f), w[f], phi[other]}; }
    __device__ __forceinline__ T flux(const Face& fd, double pc, bool side) const
    {
        const double pOwn = side ? fd.pO : pc, pNei = side ? pc : fd.pO;
        const double phif = fd.w * pOwn + (1 - fd.w) * pNei;
        const Vec3d fl {fd.s.x * phif, fd.s.y * phif, fd.s.z * phif};
        return side ? Vec3d {0.0 - fl.x, 0.0 - fl.y, 0.0 - fl.z} : fl;
    }
    __device__ __forceinline__ T internal(int f, int own, int nei) const
    {
        const double wf = w[f];
        const double phif = wf * phi[own] + (1 - wf) * phi[nei];
        const Vec3d s = ld3(Sf, f);
        return Vec3d {s.x * phif, s.y * phif, s.z * phif};
    }
    __device__ __forceinline__ T boundary(int f, int b, int) const
    {
        const double phif = phiB[b];
        const Vec3d s = ld3(Sf, f);
        return Vec3d {s.x * phif, s.y * phif, s.z * phif};
    }
};

template <class VT>
struct LaplacianOp // gaussGreenLaplacian.cpp:34-52 with uncorrected.cpp:34-51
{
    using V = VT;
    using T = typename VT::T;
    const double* __restrict__ magSf;
    const double* __restrict__ dc; // nonOrthDeltaCoeffs
    const double* __restrict__ phi;
    const double* __restrict__ phiB;
    using CV = VT;
    static constexpr int W0 = 1, W1 = 1; // stream 0 = magSf, stream 1 = nonOrthDeltaCoeffs
    static constexpr bool NEEDS_BLEND = false;
    __device__ __forceinline__ bool ownsStreams() const { return true; }
    __host__ __device__ __forceinline__ const double* s0() const { return magSf; }
    __host__ __device__ __forceinline__ const double* s1() const { return dc; }
    __host__ __device__ __forceinline__ const double* cells() const { return phi; }
    __device__ __forceinline__ T fluxv(const double* a, const double* b, const T& pOwn, const T& pNei) const
    {
        return VT::mul(a[0], VT::mul(b[0], VT::sub(pNei, pOwn)));
    }
    struct Face { double a, d; T pO; };
    __device__ __forceinline__ T cell(int c) const { return VT::ld(phi, c); }
    __device__ __forceinline__ Face load(int f, int other) const { return Face {magSf[f], dc[f], VT::ld(phi, other)}; }
    __device__ __forceinline__ T flux(const Face& fd, const T& pc, bool side) const
    {
        const T pOwn = side ? fd.pO : pc, pNei = side ? pc : fd.pO;
        const T fl = VT::mul(fd.a, VT::mul(fd.d, VT::sub(pNei, pOwn)));
        return side ? VT::sub(VT::zero(), fl) : fl;
    }
    __device__ __forceinline__ T internal(int f, int own, int nei) const
    {
        const T sn = VT::mul(dc[f], VT::sub(VT::ld(phi, nei), VT::ld(phi, own)));
        return VT::mul(magSf[f], sn);
    }
    __device__ __forceinline__ T boundary(int f, int b, int own) const
    {
        const T sn = VT::mul(dc[f], VT::sub(VT::ld(phiB, b), VT::ld(phi, own)));
        return VT::mul(magSf[f], sn);
    }
};

template <class VT>
struct SurfIntOp // surfaceIntegrate.cpp:26-41
{
    using V = VT;
    using T = typename VT::T;
    const double* __restrict__ flux_;
    using CV = void; // no cell field
    static constexpr int W0 = VT::NC, W1 = 0; // stream 0 = the face flux itself
    static constexpr bool NEEDS_BLEND = false;
    __device__ __forceinline__ bool ownsStreams() const { return false; }
    __host__ __device__ __forceinline__ const double* s0() const { return flux_; }
    __host__ __device__ __forceinline__ const double* s1() const { return nullptr; }
    __host__ __device__ __forceinline__ const double* cells() const { return nullptr; }
    __device__ __forceinline__ T fluxv(const double* a, const double*, int, int) const { return VT::ld(a, 0); }
    struct Face { T v; };
    __device__ __forceinline__ int cell(int) const { return 0; }
    __device__ __forceinline__ Face load(int f, int) const { return Face {VT::ld(flux_, f)}; }
    __device__ __forceinline__ T flux(const Face& fd, int, bool side) const
    {
        return side ? VT::sub(VT::zero(), fd.v) : fd.v;
    }
    __device__ __forceinline__ T internal(int f, int, int) const { return VT::ld(flux_, f); }
    __device__ __forceinline__ T boundary(int f, int, int) const { return VT::ld(flux_, f); }
};

// ---- scaling + store ---------------------------------------------------------------------------
// internal epilogue modes (not part of the C ABI): the operator's result feeds the next cell-wise step without a round
// trip through memory. FVK_EPI_UPDATE_VELOCITY (grad): out = epiA - (s * sum) * epiB   = updateVelocity's U = HbyA - rAU gradP
// (pressureVelocityCoupling.cpp:199-213); FVK_EPI_RHS_SUB (surfaceIntegrate / div / laplacian): out -= (0 + s * sum) * V, i.e.
// Operator::explicitOperation into a zeroed source followed by dsl::solve's rhs -= source * V (dsl/solver.hpp:73-77).
enum { FVK_EPI_UPDATE_VELOCITY = 16, FVK_EPI_RHS_SUB = 17, FVK_EPI_AXPY = 18 /* out = epiScale * (s * sum) + epiA: forwardEuler */ };
struct Scaling
{
    const double* __restrict__ V;    // cell volumes
    const double* __restrict__ view; // Coeff view or nullptr
    double coeff;
    bool invVolOnly; // grad: res *= 1 / V
    const double* epiA = nullptr;    // FVK_EPI_UPDATE_VELOCITY: HbyA (Vec3)
    const double* epiB = nullptr;    // FVK_EPI_UPDATE_VELOCITY: rAU
    double epiScale = 0.0;           // FVK_EPI_AXPY: -dt (epiA = the old-time field)
    __device__ __forceinline__ double at(int c) const
    {
        if (invVolOnly) return 1 / V[c];
        const double os = view ? view[c] * coeff : coeff; // dsl/coeff.hpp:35
        return os / V[c];
    }
};

template <class VT>
__device__ __forceinline__ void finish(double* __restrict__ out, int c, typename VT::T acc, double s, int mode, const Scaling& sc)
{
    if (mode == FVK_ADD)
        VT::st(out, c, VT::add(VT::ld(out, c), VT::mul(s, acc)));
    else if (mode == FVK_EPI_UPDATE_VELOCITY)
        VT::st(out, c, VT::sub(VT::ld(sc.epiA, c), VT::mul(sc.epiB[c], VT::mul(s, acc))));
    else if (mode == FVK_EPI_RHS_SUB)
        VT::st(out, c, VT::sub(VT::ld(out, c), VT::mul(sc.V[c], VT::add(VT::zero(), VT::mul(s, acc)))));
    else if (mode == FVK_EPI_AXPY) // a * source + 1 * old, the products and the sum of fvk_vec_waxpby(-dt, source, 1, old)
        VT::st(out, c, VT::add(VT::mul(sc.epiScale, VT::mul(s, acc)), VT::ld(sc.epiA, c)));
    else
        VT::st(out, c, VT::mul(s, acc));
}
// NOTE: res[c] *= s in the reference is (res * s); mul(s, acc) is the same product (commutative).

// ---- variant 0 (default): unified sorted stencil, one face in flight per thread ------------------------------------------------------------
// one cell: walk its stencil in ascending face id (the Serial executor's accumulation order)
template <class Op>
__device__ __forceinline__ void gather_cell(const Op& op, const Scaling& sc, int c, int nI, const int* __restrict__ seg,
                                            const int* __restrict__ ent, const int* __restrict__ owner,
                                            const int* __restrict__ neighbour, double* __restrict__ out, int mode)
{
    using VT = typename Op::V;
    typename VT::T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, c) : VT::zero();
    const int e1 = seg[c + 1];
    for (int e = seg[c]; e < e1; ++e)
    {
        const int code = ent[e];
        const int f = code >> 1;
        if (f < nI)
        {
            if (code & 1)
                acc = VT::sub(acc, op.internal(f, owner[f], c));
            else
                acc = VT::add(acc, op.internal(f, c, neighbour[f]));
        }
        else
        {
            acc = VT::add(acc, op.boundary(f, f - nI, c));
        }
    }
    finish<VT>(out, c, acc, sc.at(c), mode, sc);
}

template <class Op>
__global__ void __launch_bounds__(256)
k_gather_stencil(Op op, Scaling sc, int nC, int nI, const int* __restrict__ seg,
                 const int* __restrict__ ent, const int* __restrict__ owner,
                 const int* __restrict__ neighbour, double* __restrict__ out, int mode, const int* __restrict__ cellList = nullptr)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nC) return;
    const int c = cellList ? cellList[idx] : idx; // list: the irregular cells next to k_gather_affine (nC = list length)
    gather_cell(op, sc, c, nI, seg, ent, owner, neighbour, out, mode);
}

// ---- variants 1-4: packed plan, U faces in flight per thread -----------------------------------
// plan entry (8 B, one coalescable load): x = (faceId << 1) | side for internal faces (side 1: the cell is
// the face's neighbour), x = -(b + 1) for boundary face b; y = the other cell. Each iteration fetches U
// entries, then issues all their face/cell loads back to back (branch-free, padded lanes re-read the last
// valid entry), then folds the signed fluxes in ascending face order. Compared with the plain loop this
// removes one dependent load level (no owner[]/neighbour[] lookup) and multiplies the bytes in flight per
// thread by U, which is what a latency-bound gather needs to approach the HBM roofline.
template <class Op, int U>
__global__ void __launch_bounds__(256, (U >= 6 ? 3 : 4))
k_gather_plan(Op op, Scaling sc, int nC, int nI, const int* __restrict__ seg, const int2* __restrict__ plan,
              double* __restrict__ out, int mode)
{
    using VT = typename Op::V;
    using T = typename VT::T;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, c) : VT::zero();
    const int e0 = seg[c], e1 = seg[c + 1];
    const double s = sc.at(c);
    const auto cd = op.cell(c);
    for (int e = e0; e < e1; e += U)
    {
        int2 pe[U];
#pragma unroll
        for (int k = 0; k < U; ++k) pe[k] = plan[min(e + k, e1 - 1)];
        typename Op::Face fd[U];
#pragma unroll
        for (int k = 0; k < U; ++k) fd[k] = op.load(pe[k].x < 0 ? 0 : (pe[k].x >> 1), pe[k].y);
        T v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = op.flux(fd[k], cd, (pe[k].x & 1) != 0);
#pragma unroll
        for (int k = 0; k < U; ++k)
        {
            if (pe[k].x < 0)
            {
                const int b = -pe[k].x - 1;
                v[k] = op.boundary(nI + b, b, c);
            }
            if (e + k < e1) acc = VT::add(acc, v[k]);
        }
    }
    finish<VT>(out, c, acc, s, mode, sc);
}

// ---- variant 6 (opt-in, owner-sorted meshes): TMA-staged tile kernel, optionally double-buffered -------------
// A persistent grid (a few blocks per SM) walks the tiles of the mesh's FvkTilePlan. Everything a tile
// streams -- the face streams of its own faces (e.g. faceFlux + weights), the cells' phi, and ONE blob with
// all mesh-static data (volumes, neighbour labels, slot codes, segments, cross / boundary face lists) -- is
// a contiguous range, so one elected thread stages it with 3-4 cp.async.bulk copies (TMA, completion on an
// mbarrier) into one of two shared-memory stages while the block computes on the other: no registers, no
// dependent index loads, every byte through L2 once, loads never wait for compute. Per tile:
//   A1  thread per cell: flux of the cell's own faces from the staged streams (phi of an in-tile neighbour
//       from shared memory, otherwise one global gather)                                  -> slot [0, nf)
//   A2  thread per cross face (owner outside the tile): evaluated from global memory      -> slot [nf, nf+nx)
//   A3  thread per boundary face                                                          -> slot [nf+nx, ..)
//   B   thread per cell: adds the cell's slots in the reference's order (owner side +, neighbour side -),
//       scales by coeff/V and writes the result once.
// Bit-identical to the per-cell gather (same per-face arithmetic, same summation order).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct Stager
{
    uint32_t bar;   // mbarrier (shared address)
    uint32_t bytes; // bytes in flight
    template <class E>
    static __device__ __forceinline__ int lead(int64_t start) { return int(start & (16 / int(sizeof(E)) - 1)); }
    // Executed by ONE thread. Stage elements [start, start+count) of the 16-byte aligned global array `g` (which
    // holds `len` readable elements) so that element start+i lands at sm[lead + i]: ONE cp.async.bulk over the
    // enclosing 16-byte aligned range (over-fetching < 16 bytes on either side). Only a range that would run past
    // `len` is cut at the last aligned boundary and finished with plain loads (at most one tile per array).
    template <class E>
    __device__ __forceinline__ void stage(E* sm, const E* g, int64_t start, int count, int64_t len)
    {
        constexpr int PER = 16 / int(sizeof(E));
        if (count <= 0) return;
        const int64_t a0 = start & ~int64_t(PER - 1);
        int64_t a1 = (start + count + PER - 1) & ~int64_t(PER - 1);
        if (a1 > len)
        {
            a1 = (start + count) & ~int64_t(PER - 1);
            for (int64_t i = (a1 > a0 ? a1 : start); i < start + count; ++i) sm[i - a0] = g[i];
        }
        if (a1 > a0)
        {
            const uint32_t nbytes = uint32_t((a1 - a0) * sizeof(E));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(sm)), "l"(g + a0), "r"(nbytes), "r"(bar) : "memory");
            bytes += nbytes;
        }
    }
};

template <class CV>
struct CellLd
{
    using T = typename CV::T;
    static constexpr int NC = CV::NC;
    static __device__ __forceinline__ T ld(const double* p, int64_t i) { return CV::ld(p, i); }
};
template <>
struct CellLd<void>
{
    using T = int;
    static constexpr int NC = 0;
    static __device__ __forceinline__ int ld(const double*, int64_t) { return 0; }
};

__host__ __device__ inline size_t al16(size_t n) { return (n + 15) & ~size_t(15); }
// shared-memory layout: [2 mbarriers | 2 tile headers | sflux | stage 0 | stage 1], stage = S0 | S1 | phi | blob
template <class Op, int NSTAGE>
struct TileSmem
{
    using VT = typename Op::V;
    static constexpr int CN = CellLd<typename Op::CV>::NC;
    size_t s0, s1, phi, blob, stage, flux, total;
    __host__ __device__ explicit TileSmem(const FvkTilePlan& tp)
    {
        s0 = 0;
        s1 = s0 + al16(sizeof(double) * (size_t(Op::W0) * tp.maxF + 4));
        phi = s1 + al16(sizeof(double) * (size_t(Op::W1) * tp.maxF + 4));
        blob = phi + al16(sizeof(double) * (size_t(CN) * tp.maxC + 4));
        stage = blob + al16(size_t(tp.maxBlob));
        flux = 16 + al16(2 * sizeof(FvkTileHdr));
        total = flux + al16(sizeof(typename VT::T) * (size_t(tp.maxF) + tp.maxX + tp.maxB + 1)) + NSTAGE * stage;
    }
};

template <class Op, int NSTAGE>
__global__ void __launch_bounds__(256)
k_gather_tile(Op op, Scaling sc, FvkTilePlan tp, int nI, double* __restrict__ out, int mode)
{
    using VT = typename Op::V;
    using T = typename VT::T;
    using CL = CellLd<typename Op::CV>;
    using CT = typename CL::T;
    constexpr int CN = CL::NC;
    extern __shared__ __align__(16) unsigned char smem[];
    const TileSmem<Op, NSTAGE> lay(tp);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    FvkTileHdr* sHdr = reinterpret_cast<FvkTileHdr*>(smem + 16);
    T* sflux = reinterpret_cast<T*>(smem + lay.flux);
    unsigned char* stage0 = smem + lay.total - NSTAGE * lay.stage;

    const int G = gridDim.x;
    int t = blockIdx.x;
    if (t >= tp.nTiles) return;
    const bool producer = threadIdx.x == 0;
    const int64_t nF = int64_t(nI) + tp.nB, big = int64_t(1) << 60; // mesh/plan arrays carry 16 bytes of slack
    const int64_t lenS0 = op.ownsStreams() ? big : int64_t(Op::W0) * nF, lenPhi = int64_t(CN) * tp.nCells;

    auto issue = [&](int stg, const FvkTileHdr& h) {
        unsigned char* base = stage0 + size_t(stg) * lay.stage;
        Stager st {smem_u32(bars + stg), 0u};
        st.stage(base + lay.blob, tp.blob, h.blobOff, h.blobBytes, big);
        st.stage(reinterpret_cast<double*>(base + lay.s0), op.s0(), int64_t(Op::W0) * h.f0, Op::W0 * h.nf, lenS0);
        if (Op::W1) st.stage(reinterpret_cast<double*>(base + lay.s1), op.s1(), int64_t(Op::W1) * h.f0, Op::W1 * h.nf, big);
        if (CN) st.stage(reinterpret_cast<double*>(base + lay.phi), op.cells(), int64_t(CN) * h.c0, CN * h.nc, lenPhi);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st.bar), "r"(st.bytes) : "memory");
        sHdr[stg] = h;
    };

    FvkTileHdr hn; // producer only: header of the next tile to issue
    if (producer)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        hn = tp.hdr[t];
        issue(0, hn);
        if (t + G < tp.nTiles) hn = tp.hdr[t + G];
    }
    const double* cellsG = op.cells();
    for (int k = 0; t < tp.nTiles; ++k, t += G)
    {
        const int cur = k & (NSTAGE - 1);
        if (producer)
        {
            if (NSTAGE == 2 && t + G < tp.nTiles)
            {
                // stage cur^1 was released by the barrier that ended the previous iteration
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(cur ^ 1, hn);
                if (t + 2 * G < tp.nTiles) hn = tp.hdr[t + 2 * G];
            }
            const uint32_t parity = (NSTAGE == 2 ? (k >> 1) : k) & 1;
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(smem_u32(bars + cur)), "r"(parity) : "memory");
        }
        __syncthreads(); // stage `cur` has landed (and sHdr[cur] is visible)
        const FvkTileHdr h = sHdr[cur];
        const unsigned char* base = stage0 + size_t(cur) * lay.stage;
        const double* sS0 = reinterpret_cast<const double*>(base + lay.s0) + Stager::lead<double>(int64_t(Op::W0) * h.f0);
        const double* sS1 = reinterpret_cast<const double*>(base + lay.s1) + Stager::lead<double>(int64_t(Op::W1) * h.f0);
        const double* sPhi = reinterpret_cast<const double*>(base + lay.phi) + Stager::lead<double>(int64_t(CN) * h.c0);
        const unsigned char* sBlob = base + lay.blob;
        const FvkBlobLayout L = fvk_blob_layout(h.nc, h.nf, h.nx, h.nb, h.ne);
        const double* sV = reinterpret_cast<const double*>(sBlob);
        const int* sNei = reinterpret_cast<const int*>(sBlob + L.nei);
        const int* sXF = reinterpret_cast<const int*>(sBlob + L.xFace);
        const int* sXO = reinterpret_cast<const int*>(sBlob + L.xOwner);
        const int* sBF = reinterpret_cast<const int*>(sBlob + L.bFace);
        const uint16_t* sSeg = reinterpret_cast<const uint16_t*>(sBlob + L.seg);
        const uint16_t* sOseg = reinterpret_cast<const uint16_t*>(sBlob + L.oseg);
        const uint16_t* sCode = reinterpret_cast<const uint16_t*>(sBlob + L.code);
        const uint16_t* sXC = reinterpret_cast<const uint16_t*>(sBlob + L.xCell);
        const uint16_t* sBC = reinterpret_cast<const uint16_t*>(sBlob + L.bCell);

        // ---- A2 operands of this thread's first cross face are requested before A1 so both gather rounds overlap
        const bool hx = int(threadIdx.x) < h.nx;
        double xa[Op::W0], xb[Op::W1 ? Op::W1 : 1];
        CT xpo = CL::ld(sPhi, 0);
        if (hx)
        {
            const int f = sXF[threadIdx.x];
            xpo = CL::ld(cellsG, sXO[threadIdx.x]);
#pragma unroll
            for (int q = 0; q < Op::W0; ++q) xa[q] = op.s0()[int64_t(Op::W0) * f + q];
            if (Op::W1) xb[0] = op.s1()[f];
        }
        // ---- A1: own faces, thread per cell
        for (int i = threadIdx.x; i < h.nc; i += 256)
        {
            const CT pc = CL::ld(sPhi, i);
            const int j1 = sOseg[i + 1];
            for (int j = sOseg[i]; j < j1; ++j)
            {
                const int n = sNei[j];
                const unsigned nl = unsigned(n - h.c0);
                const CT pn = (nl < unsigned(h.nc)) ? CL::ld(sPhi, nl) : CL::ld(cellsG, n);
                sflux[j] = op.fluxv(sS0 + Op::W0 * j, sS1 + Op::W1 * j, pc, pn);
            }
        }
        // ---- A2: cross faces (this tile holds the neighbour cell), thread per face
        if (hx) sflux[h.nf + threadIdx.x] = op.fluxv(xa, xb, xpo, CL::ld(sPhi, sXC[threadIdx.x]));
        for (int i = threadIdx.x + 256; i < h.nx; i += 256)
        {
            const int f = sXF[i];
            const CT po = CL::ld(cellsG, sXO[i]);
            const CT pn = CL::ld(sPhi, sXC[i]);
            double a[Op::W0], b[Op::W1 ? Op::W1 : 1];
#pragma unroll
            for (int q = 0; q < Op::W0; ++q) a[q] = op.s0()[int64_t(Op::W0) * f + q];
            if (Op::W1) b[0] = op.s1()[f];
            sflux[h.nf + i] = op.fluxv(a, b, po, pn);
        }
        // ---- A3: boundary faces
        for (int i = threadIdx.x; i < h.nb; i += 256)
        {
            const int f = sBF[i];
            sflux[h.nf + h.nx + i] = op.boundary(f, f - nI, h.c0 + sBC[i]);
        }
        __syncthreads();
        // ---- B: per-cell accumulation in the reference's order
        for (int i = threadIdx.x; i < h.nc; i += 256)
        {
            const int c = h.c0 + i;
            T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, c) : VT::zero();
            const int e1 = sSeg[i + 1];
            for (int e = sSeg[i]; e < e1; ++e)
            {
                const unsigned code = sCode[e];
                const T v = sflux[code >> 1];
                acc = (code & 1u) ? VT::sub(acc, v) : VT::add(acc, v);
            }
            const double vol = sV[i];
            double s;
            if (sc.invVolOnly) s = 1 / vol;
            else s = (sc.view ? sc.view[c] * sc.coeff : sc.coeff) / vol;
            finish<VT>(out, c, acc, s, mode, sc);
        }
        __syncthreads(); // stage `cur` and sflux are free again
        if (NSTAGE == 1 && producer && t + G < tp.nTiles)
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(0, hn);
            if (t + 2 * G < tp.nTiles) hn = tp.hdr[t + 2 * G];
        }
    }
}

template <class Op, int NSTAGE>
int launch_tile_n(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, cudaStream_t st)
{
    static int blocksPerSm[64] = {0};
    static size_t cachedBytes[64] = {0};
    static bool optedIn[64] = {false};
    const TileSmem<Op, NSTAGE> lay(m->tp);
    const size_t bytes = lay.total;
    if (bytes > 220 * 1024) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -1;
    if (!optedIn[dev])
    {
        // opt in to large dynamic shared memory once per device and instantiation
        FVK_CUDA(cudaFuncSetAttribute(k_gather_tile<Op, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        optedIn[dev] = true;
    }
    if (cachedBytes[dev] != bytes)
    {
        int nb = 0;
        FVK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gather_tile<Op, NSTAGE>, 256, bytes));
        blocksPerSm[dev] = nb > 0 ? nb : 1;
        cachedBytes[dev] = bytes;
    }
    // NSTAGE 2: persistent grid, loads of tile k+1 overlap compute of tile k inside a block.
    // NSTAGE 1: one tile per block; overlap comes from the 2x more blocks resident per SM.
    int grid = (NSTAGE == 2) ? fvk_sm_count() * blocksPerSm[dev] : m->tp.nTiles;
    if (grid > m->tp.nTiles) grid = m->tp.nTiles;
    k_gather_tile<Op, NSTAGE><<<grid, 256, bytes, st>>>(op, sc, m->tp, m->nInternalFaces, out, mode);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
template <class Op>
int launch_tile(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, cudaStream_t st)
{
    static const int stages = [] { const char* e = std::getenv("FVK_TILE_STAGES"); return e ? std::atoi(e) : 1; }();
    return stages == 2 ? launch_tile_n<Op, 2>(m, op, sc, out, mode, st) : launch_tile_n<Op, 1>(m, op, sc, out, mode, st);
}

// ---- brick kernel (default when the mesh has a brick plan) ---------------------------------------------------------
// One block per tile of the FvkBrickPlan (<= K*TB cells: an x-run x y x z brick of a block-structured numbering, or a
// run of consecutive cells otherwise), TB threads, K cells per thread. OpenFOAM orders faces by owner, so the faces a
// cell owns are consecutive: the thread reads them straight from the face streams (no index array), evaluates each
// flux ONCE and parks it in a shared-memory slot; the few lower faces whose owner lies outside the tile ("cross"
// faces, ~0.5 per cell for a brick instead of 2-3 for a run of consecutive cells) and the boundary faces are
// evaluated by a thread each. After one barrier every cell adds its slots in the reference's order
// [lower | owned | boundary] (gaussGreenDiv.cpp:46-67: ascending face id), scales by coeff/V and writes once.
// All loads of a phase are issued before the first use (K cells x 3 faces x 3-5 operands in flight per thread), the
// only mesh-static data read besides neighbour[] are 8 B/cell of records and 8 B/cell of slot codes (4 uint16, one
// aligned load): DRAM traffic ~1.09x the algorithmic bytes (ncu, profiles/), the owner[] array is never read.
// The load chain hdr -> record -> face operands -> neighbour phi is 4 deep, so the kernel lives on resident blocks:
// K = 1 keeps it under 48 registers (5-6 blocks of 256 threads per SM).
// Load levels (each waits for the previous one): (1) record, slot codes, V, phi of the own cell, the tile's cross /
// boundary list bases -- all addressed arithmetically from blockIdx/threadIdx (FvkBrickGeom), no header; (2) the
// owned faces' neighbour labels and face operands, the cross face's labels; (3) phi of the neighbours, the cross
// face's operands. PHASE: 0 = all tiles, 1 = tiles that read no ghost cell, 2 = tiles that do (halo overlap).
template <class Op, int TB, int MINB, bool XDEFER>
__global__ void __launch_bounds__(TB, MINB)
k_gather_brick(Op op, Scaling sc, FvkBrickPlan bp, int nI, const int* __restrict__ neighbour, double* __restrict__ out, int mode,
               int phase, const int* __restrict__ tileList)
{
    using VT = typename Op::V;
    using T = typename VT::T;
    using CL = CellLd<typename Op::CV>;
    using CT = typename CL::T;
    constexpr int MO = 3; // owned faces handled in the unrolled path
    constexpr int W0 = Op::W0, W1 = Op::W1 ? Op::W1 : 1;
    extern __shared__ __align__(16) unsigned char smem[];
    T* sflux = reinterpret_cast<T*>(smem);
    const double* __restrict__ S0 = op.s0();
    const double* __restrict__ S1 = op.s1();
    const double* __restrict__ cellsG = op.cells();
    const int tid = threadIdx.x, t = tileList ? tileList[blockIdx.x] : int(blockIdx.x); // list: the tiles outside the affine box

    // ---- level 1
    const int4 ti = bp.tileInfo[t]; // xBase, nx | nb << 16, bBase, nOwnSlots | touchesGhost << 30
    if (phase && ((ti.w >> 30) & 1) != phase - 1) return;
    int nc;
    int cell = fvk_brick_cell(bp.geom, t, tid, nc);
    const bool valid = tid < nc;
    if (!valid) cell = fvk_brick_cell(bp.geom, t, 0, nc);
    uint2 r0 = make_uint2(0u, 0u), r1 = make_uint2(0u, 0u), cw = make_uint2(0xffffffffu, 0xffffffffu);
    if (valid)
    {
        const uint2* rp = reinterpret_cast<const uint2*>(bp.recF) + size_t(t) * (TB + 1) + tid;
        r0 = rp[0]; r1 = rp[1];
        cw = bp.codes4[size_t(t) * TB + tid];
    }
    const double vol = sc.V[cell];
    const CT pc = CL::ld(cellsG, cell);
    const int xBase = ti.x, nx = ti.y & 0xffff, nb = ti.y >> 16, bBase = ti.z, nOwnSlots = ti.w & 0xffff;
    const bool hx = tid < nx;
    int xf = 0, xo = 0, xn = 0;
    if (hx) { xf = bp.xFace[xBase + tid]; xo = bp.xOwner[xBase + tid]; xn = bp.xNei[xBase + tid]; }
    // ---- level 2
    const int fs = int(r0.x), slotBase = int(r0.y & 0xffffu), nOwn = int(r1.y & 0xffffu) - slotBase;
    int nbr[MO];
#pragma unroll
    for (int k = 0; k < MO; ++k) nbr[k] = (k < nOwn) ? neighbour[fs + k] : cell;
    double fa[MO][W0], fb[MO][W1];
#pragma unroll
    for (int k = 0; k < MO; ++k)
    {
        const bool p = k < nOwn;
        const int64_t f = fs + k;
#pragma unroll
        for (int i = 0; i < W0; ++i) fa[k][i] = p ? S0[int64_t(W0) * f + i] : 0.0;
        fb[k][0] = (Op::W1 && p) ? S1[f] : 0.0;
    }
    // ---- level 3 (XDEFER: the cross face's operands are fetched after the owned faces are done -- one more exposed
    // round trip for the first nx threads, eight fewer live registers at the peak)
    double xa[W0], xb[W1];
    CT xpo = pc, xpn = pc;
    if (!XDEFER && hx)
    {
#pragma unroll
        for (int i = 0; i < W0; ++i) xa[i] = S0[int64_t(W0) * xf + i];
        if (Op::W1) xb[0] = S1[xf];
        xpo = CL::ld(cellsG, xo);
        xpn = CL::ld(cellsG, xn);
    }
    CT pn[MO];
#pragma unroll
    for (int k = 0; k < MO; ++k) pn[k] = CL::ld(cellsG, nbr[k]);
#pragma unroll
    for (int k = 0; k < MO; ++k)
        if (k < nOwn) sflux[slotBase + k] = op.fluxv(fa[k], fb[k], pc, pn[k]);
    for (int k = MO; k < nOwn; ++k) // polyhedral cells owning more than MO faces
    {
        const int64_t f = fs + k;
        double a[W0], b[W1];
#pragma unroll
        for (int i = 0; i < W0; ++i) a[i] = S0[int64_t(W0) * f + i];
        if (Op::W1) b[0] = S1[f];
        sflux[slotBase + k] = op.fluxv(a, b, pc, CL::ld(cellsG, neighbour[f]));
    }
    // ---- cross faces beyond the first TB, boundary faces: thread per face
    if (XDEFER && hx)
    {
#pragma unroll
        for (int i = 0; i < W0; ++i) xa[i] = S0[int64_t(W0) * xf + i];
        if (Op::W1) xb[0] = S1[xf];
        xpo = CL::ld(cellsG, xo);
        xpn = CL::ld(cellsG, xn);
    }
    if (hx) sflux[nOwnSlots + tid] = op.fluxv(xa, xb, xpo, xpn);
    for (int i = tid + TB; i < nx; i += TB)
    {
        const int f = bp.xFace[xBase + i];
        const CT po = CL::ld(cellsG, bp.xOwner[xBase + i]);
        const CT pnn = CL::ld(cellsG, bp.xNei[xBase + i]);
        double a[W0], b[W1];
#pragma unroll
        for (int q = 0; q < W0; ++q) a[q] = S0[int64_t(W0) * f + q];
        if (Op::W1) b[0] = S1[f];
        sflux[nOwnSlots + i] = op.fluxv(a, b, po, pnn);
    }
    for (int i = tid; i < nb; i += TB)
    {
        const int f = bp.bFace[bBase + i];
        sflux[nOwnSlots + nx + i] = op.boundary(f, f - nI, bp.bCell[bBase + i]);
    }
    __syncthreads();
    // ---- per-cell accumulation in the reference's order
    if (!valid) return;
    T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, cell) : VT::zero();
    const unsigned cd[4] = {cw.x & 0xffffu, cw.x >> 16, cw.y & 0xffffu, cw.y >> 16};
    const int listBase = int(r0.y >> 16), nList = int(r1.y >> 16) - listBase;
    int nLow = 0;
    bool more = true;
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const bool low = more && cd[j] != 0xffffu && (cd[j] & 1u);
        if (low) { acc = VT::sub(acc, sflux[cd[j] >> 1]); ++nLow; }
        more = low;
    }
    if (more && nList > 4) // more than 4 lower faces: the rest of the list comes from the tile's code array
    {
        const unsigned short* gcodes = bp.codes + bp.hdr[t].codeBase + listBase;
        for (; nLow < nList; ++nLow)
        {
            const unsigned code = gcodes[nLow];
            if (code == 0xffffu || !(code & 1u)) break;
            acc = VT::sub(acc, sflux[code >> 1]);
        }
    }
    for (int k = 0; k < nOwn; ++k) acc = VT::add(acc, sflux[slotBase + k]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j >= nLow && cd[j] != 0xffffu) acc = VT::add(acc, sflux[cd[j] >> 1]);
    if (nList > 4)
    {
        const unsigned short* gcodes = bp.codes + bp.hdr[t].codeBase + listBase;
        for (int j = (nLow > 4 ? nLow : 4); j < nList; ++j)
        {
            const unsigned code = gcodes[j];
            if (code == 0xffffu) break;
            acc = VT::add(acc, sflux[code >> 1]);
        }
    }
    double s;
    if (sc.invVolOnly) s = 1 / vol;
    else s = (sc.view ? sc.view[cell] * sc.coeff : sc.coeff) / vol;
    finish<VT>(out, cell, acc, s, mode, sc);
}

// ---- affine kernel: block-structured meshes whose topology the plan proved (FvkBrickGeom) ----------------------------
// NO index array is read: cell ids, face ids, neighbours and slots are arithmetic, so every load of the kernel is issued
// in the first instruction window (one memory round trip per block instead of three dependent ones) and the owner /
// neighbour labels (24 B per face of the algorithmic traffic) stay in DRAM. One block per tile, one cell per thread.
// Every cell that owns all three upper faces evaluates them once into shared-memory slots; the faces on the tile's three
// lower sides ("cross" faces) are evaluated by a thread each; REGULAR cells (not in the outermost layer of the block) then
// add their six slots in the reference's order [zL, yL, xL | x, y, z] -- arithmetic and results are those of the generic
// kernels, bit for bit. The irregular cells (boundary / processor-cut layers, a few per cent) are written by
// k_gather_stencil over the plan's cell list. The regular cells never read a ghost cell: this kernel IS the interior
// phase of the halo overlap.
struct AffineTail
{
    int nBlocks, nIrr, nI; // nBlocks = 0: tiles only
    const int *irrCells, *seg, *ent, *owner, *neighbour;
};
template <class Op, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB)
k_gather_affine(Op op, Scaling sc, FvkBrickGeom g, double* __restrict__ out, int mode, AffineTail tail)
{
    if (int(blockIdx.x) < tail.nBlocks)
    { // the first blocks of the grid: the irregular cells, per-cell gather (same launch; latency-bound, so they start
      // first and run beside the streaming tiles)
        const int idx = blockIdx.x * TB + threadIdx.x;
        if (idx < tail.nIrr) gather_cell(op, sc, tail.irrCells[idx], tail.nI, tail.seg, tail.ent, tail.owner, tail.neighbour, out, mode);
        return;
    }
    const int tileId = blockIdx.x - tail.nBlocks;
    using VT = typename Op::V;
    using T = typename VT::T;
    using CL = CellLd<typename Op::CV>;
    using CT = typename CL::T;
    constexpr int W0 = Op::W0, W1 = Op::W1 ? Op::W1 : 1;
    extern __shared__ __align__(16) unsigned char smem[];
    T* sflux = reinterpret_cast<T*>(smem);
    const double* __restrict__ S0 = op.s0();
    const double* __restrict__ S1 = op.s1();
    const double* __restrict__ cellsG = op.cells();
    const int tid = threadIdx.x;
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int64_t nxy = int64_t(nx) * ny;
    const int tx = g.tUp[0], ty = g.tUp[1];
    // tile -> origin and extent (edge tiles may be ragged)
    const int ix = tileId % g.tdim[0], q = tileId / g.tdim[0], iy = q % g.tdim[1], iz = q / g.tdim[1];
    const int x0 = ix * g.brick[0], y0 = iy * g.brick[1], z0 = iz * g.brick[2];
    const int rl = min(g.brick[0], nx - x0), ry = min(g.brick[1], ny - y0), rz = min(g.brick[2], nz - z0);
    const bool full = rl == g.brick[0] && ry == g.brick[1] && g.shiftL >= 0 && g.shiftBy >= 0;
    const int nc = rl * ry * rz;
    int off, a, b;
    if (full) { off = tid & (rl - 1); const int r = tid >> g.shiftL; a = r & (ry - 1); b = r >> g.shiftBy; }
    else { off = tid % rl; const int r = tid / rl; a = r % ry; b = r / ry; }
    const bool valid = tid < nc;
    const int i = x0 + off, j = y0 + a, k = z0 + b;
    // owns three faces with owned neighbours / is a regular cell
    const bool upper = valid && i < nx - 1 && j < ny - 1 && k < nz - 1;
    const bool regular = upper && i > 0 && j > 0 && k > 0;
    const int64_t cell = i + int64_t(nx) * j + nxy * k;
    const int64_t fs = 3 * cell - int64_t(tx) * (j + int64_t(ny) * k) - int64_t(ty) * k * nx;
    // ---- every load of the owned faces
    double fa[3][W0], fb[3][W1];
    CT pc = CL::ld(cellsG, 0), pn0 = pc, pn1 = pc, pn2 = pc;
    double vol = 1.0;
    if (upper)
    {
#pragma unroll
        for (int f = 0; f < 3; ++f)
        {
#pragma unroll
            for (int c = 0; c < W0; ++c) fa[f][c] = S0[int64_t(W0) * (fs + f) + c];
            fb[f][0] = Op::W1 ? S1[fs + f] : 0.0;
        }
        pc = CL::ld(cellsG, cell);
        pn0 = CL::ld(cellsG, cell + 1); pn1 = CL::ld(cellsG, cell + nx); pn2 = CL::ld(cellsG, cell + nxy);
        vol = sc.V[cell];
    }
    // ---- cross faces: e < nZ: z side (b = 0) | < nZ + nY: y side (a = 0) | x side (off = 0); only for regular consumers
    const int nZ = rl * ry, nY = rl * rz, nX = ry * rz, nCross = nZ + nY + nX;
    const int XB = 3 * TB;
    for (int e = tid; e < nCross; e += TB)
    {
        int co, ca, cb, e1;
        int64_t dOwner, dFace;
        if (e < nZ) { e1 = e; cb = 0; dOwner = nxy; dFace = -3 * nxy + int64_t(tx) * ny + int64_t(ty) * nx + 2; }
        else if (e < nZ + nY) { e1 = e - nZ; ca = 0; dOwner = nx; dFace = -3 * int64_t(nx) + tx + 1; }
        else { e1 = e - nZ - nY; co = 0; dOwner = 1; dFace = -3; }
        if (e < nZ) { if (full) { co = e1 & (rl - 1); ca = e1 >> g.shiftL; } else { co = e1 % rl; ca = e1 / rl; } }
        else if (e < nZ + nY) { if (full) { co = e1 & (rl - 1); cb = e1 >> g.shiftL; } else { co = e1 % rl; cb = e1 / rl; } }
        else { if (full) { ca = e1 & (ry - 1); cb = e1 >> g.shiftBy; } else { ca = e1 % ry; cb = e1 / ry; } }
        const int ci = x0 + co, cj = y0 + ca, ck = z0 + cb;
        if (!(ci > 0 && ci < nx - 1 && cj > 0 && cj < ny - 1 && ck > 0 && ck < nz - 1)) continue; // consumer not regular
        const int64_t cc = ci + int64_t(nx) * cj + nxy * ck;
        const int64_t xf = 3 * cc - int64_t(tx) * (cj + int64_t(ny) * ck) - int64_t(ty) * ck * nx + dFace;
        double xa[W0], xb[W1];
#pragma unroll
        for (int c = 0; c < W0; ++c) xa[c] = S0[int64_t(W0) * xf + c];
        xb[0] = Op::W1 ? S1[xf] : 0.0;
        const CT po = CL::ld(cellsG, cc - dOwner), pnn = CL::ld(cellsG, cc);
        sflux[XB + e] = op.fluxv(xa, xb, po, pnn);
    }
    if (upper)
    {
        sflux[3 * tid + 0] = op.fluxv(fa[0], fb[0], pc, pn0);
        sflux[3 * tid + 1] = op.fluxv(fa[1], fb[1], pc, pn1);
        sflux[3 * tid + 2] = op.fluxv(fa[2], fb[2], pc, pn2);
    }
    __syncthreads();
    if (!regular) return;
    T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, cell) : VT::zero();
    acc = VT::sub(acc, b > 0 ? sflux[3 * (tid - rl * ry) + 2] : sflux[XB + off + rl * a]);
    acc = VT::sub(acc, a > 0 ? sflux[3 * (tid - rl) + 1] : sflux[XB + nZ + off + rl * b]);
    acc = VT::sub(acc, off > 0 ? sflux[3 * (tid - 1)] : sflux[XB + nZ + nY + a + ry * b]);
    acc = VT::add(acc, sflux[3 * tid + 0]);
    acc = VT::add(acc, sflux[3 * tid + 1]);
    acc = VT::add(acc, sflux[3 * tid + 2]);
    double s;
    if (sc.invVolOnly) s = 1 / vol;
    else s = (sc.view ? sc.view[cell] * sc.coeff : sc.coeff) / vol;
    finish<VT>(out, cell, acc, s, mode, sc);
}

// The same kernel with CPT cells per thread: the tile is CPT bricks stacked in z (thread t owns the cells t, t + TB, ...), every
// load of all its cells is issued before the first use. For the operators with ONE face operand (surfaceIntegrate, upwind div:
// 40-60 registers) a thread of the kernel above has too few bytes in flight to keep HBM busy (ncu: 40 % of DRAM peak at 75 %
// occupancy, profiles/r2z_ncu_explicit_256.csv); two cells per thread raise that by half at the register budget of the others.
template <class Op, int TB, int MINB, int CPT>
__global__ void __launch_bounds__(TB, MINB)
k_gather_affine_n(Op op, Scaling sc, FvkBrickGeom g, double* __restrict__ out, int mode, AffineTail tail)
{
    if (int(blockIdx.x) < tail.nBlocks)
    {
        const int idx = blockIdx.x * TB + threadIdx.x;
        if (idx < tail.nIrr) gather_cell(op, sc, tail.irrCells[idx], tail.nI, tail.seg, tail.ent, tail.owner, tail.neighbour, out, mode);
        return;
    }
    const int tileId = blockIdx.x - tail.nBlocks;
    using VT = typename Op::V;
    using T = typename VT::T;
    using CL = CellLd<typename Op::CV>;
    using CT = typename CL::T;
    constexpr int W0 = Op::W0, W1 = Op::W1 ? Op::W1 : 1;
    extern __shared__ __align__(16) unsigned char smem[];
    T* sflux = reinterpret_cast<T*>(smem);
    const double* __restrict__ S0 = op.s0();
    const double* __restrict__ S1 = op.s1();
    const double* __restrict__ cellsG = op.cells();
    const int tid = threadIdx.x;
    const int nx = g.dims[0], ny = g.dims[1], nz = g.dims[2];
    const int64_t nxy = int64_t(nx) * ny;
    const int tx = g.tUp[0], ty = g.tUp[1];
    const int bz = g.brick[2] * CPT;
    const int ix = tileId % g.tdim[0], q = tileId / g.tdim[0], iy = q % g.tdim[1], iz = q / g.tdim[1];
    const int x0 = ix * g.brick[0], y0 = iy * g.brick[1], z0 = iz * bz;
    const int rl = min(g.brick[0], nx - x0), ry = min(g.brick[1], ny - y0), rz = min(bz, nz - z0);
    const bool full = rl == g.brick[0] && ry == g.brick[1] && g.shiftL >= 0 && g.shiftBy >= 0;
    const int nc = rl * ry * rz;
    int off[CPT], a[CPT], b[CPT];
    bool upper[CPT], regular[CPT];
    int64_t cell[CPT];
    double fa[CPT][3][W0], fb[CPT][3][W1], vol[CPT];
    CT pc[CPT], pn[CPT][3];
#pragma unroll
    for (int u = 0; u < CPT; ++u)
    {
        const int lc = tid + u * TB;
        if (full) { off[u] = lc & (rl - 1); const int r = lc >> g.shiftL; a[u] = r & (ry - 1); b[u] = r >> g.shiftBy; }
        else { off[u] = lc % rl; const int r = lc / rl; a[u] = r % ry; b[u] = r / ry; }
        const int i = x0 + off[u], j = y0 + a[u], k = z0 + b[u];
        upper[u] = lc < nc && i < nx - 1 && j < ny - 1 && k < nz - 1;
        regular[u] = upper[u] && i > 0 && j > 0 && k > 0;
        cell[u] = i + int64_t(nx) * j + nxy * k;
        const int64_t fs = 3 * cell[u] - int64_t(tx) * (j + int64_t(ny) * k) - int64_t(ty) * k * nx;
        pc[u] = CL::ld(cellsG, 0); pn[u][0] = pn[u][1] = pn[u][2] = pc[u];
        vol[u] = 1.0;
        if (upper[u])
        {
#pragma unroll
            for (int f = 0; f < 3; ++f)
            {
#pragma unroll
                for (int c = 0; c < W0; ++c) fa[u][f][c] = S0[int64_t(W0) * (fs + f) + c];
                fb[u][f][0] = Op::W1 ? S1[fs + f] : 0.0;
            }
            pc[u] = CL::ld(cellsG, cell[u]);
            pn[u][0] = CL::ld(cellsG, cell[u] + 1); pn[u][1] = CL::ld(cellsG, cell[u] + nx); pn[u][2] = CL::ld(cellsG, cell[u] + nxy);
            vol[u] = sc.V[cell[u]];
        }
    }
    // ---- cross faces: e < nZ: z side (b = 0) | < nZ + nY: y side (a = 0) | x side (off = 0); only for regular consumers
    const int nZ = rl * ry, nY = rl * rz, nX = ry * rz, nCross = nZ + nY + nX;
    constexpr int XB = 3 * TB * CPT;
    for (int e = tid; e < nCross; e += TB)
    {
        int co, ca, cb, e1;
        int64_t dOwner, dFace;
        if (e < nZ) { e1 = e; cb = 0; dOwner = nxy; dFace = -3 * nxy + int64_t(tx) * ny + int64_t(ty) * nx + 2; }
        else if (e < nZ + nY) { e1 = e - nZ; ca = 0; dOwner = nx; dFace = -3 * int64_t(nx) + tx + 1; }
        else { e1 = e - nZ - nY; co = 0; dOwner = 1; dFace = -3; }
        if (e < nZ) { if (full) { co = e1 & (rl - 1); ca = e1 >> g.shiftL; } else { co = e1 % rl; ca = e1 / rl; } }
        else if (e < nZ + nY) { if (full) { co = e1 & (rl - 1); cb = e1 >> g.shiftL; } else { co = e1 % rl; cb = e1 / rl; } }
        else { if (full) { ca = e1 & (ry - 1); cb = e1 >> g.shiftBy; } else { ca = e1 % ry; cb = e1 / ry; } }
        const int ci = x0 + co, cj = y0 + ca, ck = z0 + cb;
        if (!(ci > 0 && ci < nx - 1 && cj > 0 && cj < ny - 1 && ck > 0 && ck < nz - 1)) continue; // consumer not regular
        const int64_t cc = ci + int64_t(nx) * cj + nxy * ck;
        const int64_t xf = 3 * cc - int64_t(tx) * (cj + int64_t(ny) * ck) - int64_t(ty) * ck * nx + dFace;
        double xa[W0], xb[W1];
#pragma unroll
        for (int c = 0; c < W0; ++c) xa[c] = S0[int64_t(W0) * xf + c];
        xb[0] = Op::W1 ? S1[xf] : 0.0;
        const CT po = CL::ld(cellsG, cc - dOwner), pnn = CL::ld(cellsG, cc);
        sflux[XB + e] = op.fluxv(xa, xb, po, pnn);
    }
#pragma unroll
    for (int u = 0; u < CPT; ++u)
        if (upper[u])
        {
            const int lc = tid + u * TB;
            sflux[3 * lc + 0] = op.fluxv(fa[u][0], fb[u][0], pc[u], pn[u][0]);
            sflux[3 * lc + 1] = op.fluxv(fa[u][1], fb[u][1], pc[u], pn[u][1]);
            sflux[3 * lc + 2] = op.fluxv(fa[u][2], fb[u][2], pc[u], pn[u][2]);
        }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < CPT; ++u)
    {
        if (!regular[u]) continue;
        const int lc = tid + u * TB;
        T acc = (mode == FVK_ACC_SCALE) ? VT::ld(out, cell[u]) : VT::zero();
        acc = VT::sub(acc, b[u] > 0 ? sflux[3 * (lc - rl * ry) + 2] : sflux[XB + off[u] + rl * a[u]]);
        acc = VT::sub(acc, a[u] > 0 ? sflux[3 * (lc - rl) + 1] : sflux[XB + nZ + off[u] + rl * b[u]]);
        acc = VT::sub(acc, off[u] > 0 ? sflux[3 * (lc - 1)] : sflux[XB + nZ + nY + a[u] + ry * b[u]]);
        acc = VT::add(acc, sflux[3 * lc + 0]);
        acc = VT::add(acc, sflux[3 * lc + 1]);
        acc = VT::add(acc, sflux[3 * lc + 2]);
        double s;
        if (sc.invVolOnly) s = 1 / vol[u];
        else s = (sc.view ? sc.view[cell[u]] * sc.coeff : sc.coeff) / vol[u];
        finish<VT>(out, cell[u], acc, s, mode, sc);
    }
}

template <class Op, int TB, int MINB, bool XDEFER>
int launch_brick_n(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, cudaStream_t st)
{
    using T = typename Op::V::T;
    static bool optedIn[64] = {false};
    const FvkBrickGeom& g = m->bp.geom;
    const size_t bytes = size_t(m->bp.maxSlots) * sizeof(T);
    if (bytes > 200 * 1024 || g.cap != TB) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -1;
    if (!optedIn[dev])
    {
        FVK_CUDA(cudaFuncSetAttribute(k_gather_brick<Op, TB, MINB, XDEFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        optedIn[dev] = true;
    }
    k_gather_brick<Op, TB, MINB, XDEFER><<<m->bp.nTiles, TB, bytes, st>>>(op, sc, m->bp, m->nInternalFaces, m->neighbour, out, mode, m->tilePhase, nullptr);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
// block-structured mesh with proven topology: the affine kernel writes the regular cells (= the interior phase of the halo
// overlap), the per-cell gather over the plan's list the irregular boundary / cut layers (= the halo phase)
template <class Op, int TB, int MINB>
int launch_affine_n(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, cudaStream_t st)
{
    using T = typename Op::V::T;
    const FvkBrickGeom& g = m->bp.geom;
    const size_t ab = (size_t(3) * TB + size_t(g.maxCross)) * sizeof(T);
    if (ab > 48 * 1024) return -1;
    const int listBlocks = m->tilePhase == 0 ? (m->bp.nIrr + TB - 1) / TB : 0;
    AffineTail tail {listBlocks, m->bp.nIrr, m->nInternalFaces, m->bp.irrCells, m->stencilSeg, m->gatherEnt, m->owner, m->neighbour};
    // one-operand scalar operators: two cells per thread (k_gather_affine_n); FVK_AFFINE_CPT="cpt,minb" overrides for sweeps
    constexpr bool oneOperand = sizeof(T) == 8 && Op::W0 == 1 && Op::W1 == 0;
    int cpt = oneOperand ? 2 : 1, minb2 = 8;
    {
        static const char* env = std::getenv("FVK_AFFINE_CPT");
        int a1 = 0, a2 = 0;
        if (env && std::sscanf(env, "%d,%d", &a1, &a2) == 2) { cpt = oneOperand ? a1 : 1; minb2 = a2; }
    }
    if constexpr (oneOperand && TB == 128)
    if (m->tilePhase != 2 && cpt == 2)
    {
        const int bz = g.brick[2] * 2, tz = (g.dims[2] + bz - 1) / bz;
        const int nTiles2 = g.tdim[0] * g.tdim[1] * tz;
        const size_t ab2 = (size_t(3) * TB * 2 + size_t(g.brick[0]) * g.brick[1] + size_t(g.brick[0]) * bz + size_t(g.brick[1]) * bz) * sizeof(T);
        if (ab2 <= 48 * 1024)
        {
            if (minb2 >= 8) k_gather_affine_n<Op, 128, 8, 2><<<listBlocks + nTiles2, 128, ab2, st>>>(op, sc, g, out, mode, tail);
            else k_gather_affine_n<Op, 128, 6, 2><<<listBlocks + nTiles2, 128, ab2, st>>>(op, sc, g, out, mode, tail);
            FVK_LAUNCH_CHECK();
            return FVK_OK;
        }
    }
    if (m->tilePhase != 2) // phase 0: one launch, list blocks + tiles
        k_gather_affine<Op, TB, MINB><<<listBlocks + m->bp.nTiles, TB, ab, st>>>(op, sc, g, out, mode, tail);
    else if (m->bp.nIrr > 0)
        k_gather_stencil<Op><<<(m->bp.nIrr + 255) / 256, 256, 0, st>>>(op, sc, m->bp.nIrr, m->nInternalFaces, m->stencilSeg, m->gatherEnt, m->owner,
                                                                        m->neighbour, out, mode, m->bp.irrCells);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

// kernel configuration: threads per block TB (= max cells per tile), resident blocks the register allocation aims at.
// fvk_set_brick_config / FVK_BRICK_CFG="1,TB,MINB" override for sweeps (instantiated combinations only).
template <class Op>
int launch_brick(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, cudaStream_t st)
{
    // resident blocks aimed at = the highest occupancy ptxas reaches WITHOUT spilling (spills cost more than the extra
    // warps bring, r2 sweeps): 40 registers only fit the two-operand scalar laplacian and surfaceIntegrate
    constexpr bool light = sizeof(typename Op::V::T) == 8 && Op::W0 == 1 && !Op::NEEDS_BLEND;
    constexpr bool wide = sizeof(typename Op::V::T) > 8 || Op::W0 > 1;
    const int TB = m->bp.geom.cap; // threads per block = the stride of the plan's per-cell arrays
    int MINB = (light ? 6 : 4) * 256 / TB;
    int cfg[3];
    if (fvk_brick_config(cfg) && cfg[1] == TB) MINB = cfg[2];
    if (m->bp.geom.affine && !fvk_no_affine())
    {
        // the affine kernel keeps all operands of a cell in flight: spill-free register budgets (ptxas -v): 40 for the
        // scalar surfaceIntegrate, 64 for the scalar operators, 80 for Vec3 operands
        constexpr bool tiny = sizeof(typename Op::V::T) == 8 && Op::W0 == 1 && Op::W1 == 0 && !Op::NEEDS_BLEND;
        int MA = (tiny ? 6 : (wide ? 3 : 4)) * 256 / TB;
        if (fvk_brick_config(cfg) && cfg[1] == TB && cfg[0] == 3) MA = cfg[2]; // config[0] == 3: override for the affine kernel
#define FVK_AFFINE_CASE(tb, mb) if (TB == tb && MA == mb) { const int rc = launch_affine_n<Op, tb, mb>(m, op, sc, out, mode, st); if (rc >= 0) return rc; }
        FVK_AFFINE_CASE(128, 6) FVK_AFFINE_CASE(128, 8) FVK_AFFINE_CASE(128, 12)
        FVK_AFFINE_CASE(256, 3) FVK_AFFINE_CASE(256, 4) FVK_AFFINE_CASE(256, 6)
        FVK_AFFINE_CASE(512, 2) FVK_AFFINE_CASE(512, 3)
#undef FVK_AFFINE_CASE
    }
#define FVK_BRICK_CASE(tb, mb) if (TB == tb && MINB == mb) return launch_brick_n<Op, tb, mb, false>(m, op, sc, out, mode, st)
    FVK_BRICK_CASE(128, 8); FVK_BRICK_CASE(128, 12); FVK_BRICK_CASE(256, 4); FVK_BRICK_CASE(256, 6);
    FVK_BRICK_CASE(512, 2); FVK_BRICK_CASE(512, 3);
#undef FVK_BRICK_CASE
    return -1;
}

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Op>
int launch_gather(const fvk_mesh* m, Op op, Scaling sc, double* out, int mode, fvk_stream stream)
{
    if (mode != FVK_SET && mode != FVK_ACC_SCALE && mode != FVK_ADD && mode != FVK_EPI_UPDATE_VELOCITY && mode != FVK_EPI_RHS_SUB && mode != FVK_EPI_AXPY)
        return fvk_fail(FVK_EINVAL, "bad mode %d", mode);
    const int nC = m->nOwned;
    const int nI = m->nInternalFaces;
    const int grid = (nC + 255) / 256;
    const int variant = (nI == 0) ? 0 : fvk_variant(); // the plan kernels read face 0 on padded lanes
    if (nC == 0) return FVK_OK;
    // variant 0 (default): the brick kernel when the mesh has a brick plan, else the per-cell gather (also variant 5)
    if ((fvk_variant() == 0 || fvk_variant() == 7) && m->bp.nTiles > 0)
    {
        const int rc = launch_brick(m, op, sc, out, mode, fvk_cu(stream));
        if (rc >= 0) return rc; // -1: slots do not fit in shared memory -> per-cell gather below
    }
    // the other kernels have no interior / halo split: the interior phase does nothing, the halo phase everything
    if (m->tilePhase == 1) return FVK_OK;
    // variant 6: TMA-staged tile kernel. Parity-green but (r1 measurements, profiles/r1_tile_kernel_sweep.md) not yet
    // faster than the per-cell gather, so it is opt-in until the staging pipeline is tuned.
    if (variant == 6 && m->tp.nTiles > 0 && aligned16(op.s0()) && aligned16(op.s1()) && aligned16(op.cells()))
    {
        const int rc = launch_tile(m, op, sc, out, mode, fvk_cu(stream));
        if (rc >= 0) return rc; // -1: tile does not fit in shared memory -> per-cell gather below
    }
    const int2* plan = reinterpret_cast<const int2*>(m->gatherPlan);
    if (variant >= 1 && variant <= 4 && !plan)
        return fvk_fail(FVK_EUNSUPPORTED, "the packed-plan experiment kernels need a mesh created with FVK_EXPERIMENT_PLANS=1");
    cudaStream_t st = fvk_cu(stream);
    switch (variant)
    {
        case 1: k_gather_plan<Op, 3><<<grid, 256, 0, st>>>(op, sc, nC, nI, m->stencilSeg, plan, out, mode); break;
        case 2: k_gather_plan<Op, 2><<<grid, 256, 0, st>>>(op, sc, nC, nI, m->stencilSeg, plan, out, mode); break;
        case 3: k_gather_plan<Op, 6><<<grid, 256, 0, st>>>(op, sc, nC, nI, m->stencilSeg, plan, out, mode); break;
        case 4: k_gather_plan<Op, 1><<<grid, 256, 0, st>>>(op, sc, nC, nI, m->stencilSeg, plan, out, mode); break;
        default: // variant 0 (default): per-cell gather over the unified sorted stencil
            k_gather_stencil<Op><<<grid, 256, 0, st>>>(op, sc, nC, nI, m->stencilSeg, m->gatherEnt, m->owner, m->neighbour, out, mode);
            break;
    }
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

#define REQUIRE(cond, msg)                                                                         \
    do { if (!(cond)) return fvk_fail(FVK_EINVAL, "%s: %s", __func__, msg); } while (0)

// ---- face kernels (interpolate / snGrad / weights) -----------------------------------------------
template <class VT, int SCHEME>
__global__ void __launch_bounds__(256)
k_interpolate(int nI, int nF, const int* __restrict__ owner, const int* __restrict__ neighbour,
              const double* __restrict__ w, const double* __restrict__ faceFlux,
              const double* __restrict__ phi, const double* __restrict__ phiB, double* __restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    typename VT::T r;
    if (f < nI)
    {
        if (SCHEME == FVK_LINEAR)
        {
            const double wf = w[f];
            r = VT::add(VT::mul(wf, VT::ld(phi, owner[f])), VT::mul(1 - wf, VT::ld(phi, neighbour[f])));
        }
        else
            r = (faceFlux[f] >= 0) ? VT::ld(phi, owner[f]) : VT::ld(phi, neighbour[f]);
    }
    else
        r = VT::mul(w[f], VT::ld(phiB, f - nI));
    VT::st(out, f, r);
}

template <class VT>
__global__ void __launch_bounds__(256)
k_sngrad(int nI, int nF, const int* __restrict__ owner, const int* __restrict__ neighbour,
         const double* __restrict__ dc, const double* __restrict__ phi, const double* __restrict__ phiB,
         double* __restrict__ out)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    const typename VT::T a = (f < nI) ? VT::ld(phi, neighbour[f]) : VT::ld(phiB, f - nI);
    VT::st(out, f, VT::mul(dc[f], VT::sub(a, VT::ld(phi, owner[f]))));
}

__global__ void __launch_bounds__(256)
k_weights(int scheme, int nI, int nF, const double* __restrict__ geomW, const double* __restrict__ faceFlux,
          double* __restrict__ wFace, double* __restrict__ wB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    if (f < nI)
        wFace[f] = (scheme == FVK_LINEAR) ? geomW[f] : (faceFlux[f] >= 0 ? 1.0 : 0.0);
    else
    {
        const double v = (scheme == FVK_LINEAR) ? geomW[f] : 1.0;
        wFace[f] = v;
        if (wB) wB[f - nI] = v;
    }
}

// ---- CoNum -------------------------------------------------------------------------------------
// stage 1: per cell sumPhi = sum_f |F_f| (own and nei both +), block partials of
// max(sumPhi/V), sum(sumPhi), sum(V); stage 2: one block folds the partials in fixed order.
struct CoNumOp
{
    const double* __restrict__ faceFlux;
    // coNum.cpp:43,53 takes sqrt(F * F). In binary floating point with correctly rounded * and sqrt that is exactly |F| whenever
    // F * F neither underflows nor overflows (Boldo 2015), so the (slow, multi-instruction) fp64 sqrt only runs outside that range
    static __device__ __forceinline__ bool abs_is_exact(double F)
    {
        const double a = fabs(F);
        return a < 1e150 && (a > 1e-150 || a == 0.0); // (a fluid at rest has F == 0 on most faces)
    }
    static __device__ __forceinline__ double mag(double F) { return abs_is_exact(F) ? fabs(F) : sqrt(F * F); }
    __device__ __forceinline__ double at(int f) const { return mag(faceFlux[f]); }
};

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// aff.on: block topology proven by the plan -- a REGULAR cell's six faces are [zL, yL, xL | x, y, z] with arithmetic ids (same
// order as its stencil), so no index array is read for it
struct CoNumAffine
{
    int on, nx, ny, nz, tx, ty;
    int nIrr;            // on: the irregular cells (boundary / cut layers) are taken from this list, spread over ALL threads
    const int* irrCells;
};
// one cell through its stencil: the face ids of up to eight entries, then their fluxes, then the sums -- three dependent round
// trips instead of one per face
__device__ __forceinline__ double conum_cell_generic(const CoNumOp& op, int c, const int* __restrict__ seg, const int* __restrict__ ent)
{
    double acc = 0.0;
    const int e0 = seg[c], e1 = seg[c + 1];
    int ids[8];
    double F[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) ids[q] = (e0 + q < e1) ? (ent[e0 + q] >> 1) : -1;
#pragma unroll
    for (int q = 0; q < 8; ++q) F[q] = ids[q] >= 0 ? op.faceFlux[ids[q]] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (e0 + q < e1) acc += CoNumOp::mag(F[q]);
    for (int e = e0 + 8; e < e1; ++e) acc += op.at(ent[e] >> 1);
    return acc;
}
__device__ __forceinline__ void conum_block_partials(double lmax, double lphi, double lvol, double* __restrict__ partial)
{
    __shared__ double sMax[8], sPhi[8], sVol[8];
    lmax = warp_max(lmax); lphi = warp_sum(lphi); lvol = warp_sum(lvol);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sMax[wid] = lmax; sPhi[wid] = lphi; sVol[wid] = lvol; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; ++i) { lmax = fmax(lmax, sMax[i]); lphi += sPhi[i]; lvol += sVol[i]; }
        partial[3 * blockIdx.x] = lmax; partial[3 * blockIdx.x + 1] = lphi; partial[3 * blockIdx.x + 2] = lvol;
    }
}
// stage 1a, block topology proven: the REGULAR cells, index-free, two cells per pass of the grid-stride loop (fourteen
// independent loads in flight per thread). The irregular cells are skipped here: they cost three dependent round trips and would
// hold up their warp in EVERY pass (the grid stride is a multiple of nx: the same lanes meet the block's outer layer each time).
__global__ void __launch_bounds__(256, 4)
k_conum_regular(CoNumOp op, int nC, const double* __restrict__ V, double* __restrict__ partial, CoNumAffine aff)
{
    double lmax = -1.7976931348623157e308, lphi = 0.0, lvol = 0.0;
    const int stride = gridDim.x * blockDim.x;
    const int64_t nxy = int64_t(aff.nx) * aff.ny;
    for (int c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < nC; c0 += 2 * stride)
    {
        double F[2][6], v[2];
        bool reg[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            const int c = c0 + u * stride;
            reg[u] = false;
            v[u] = 1.0;
            if (c < nC)
            {
                const int i = c % aff.nx, q = c / aff.nx, j = q % aff.ny, k = q / aff.ny;
                reg[u] = i > 0 && i < aff.nx - 1 && j > 0 && j < aff.ny - 1 && k > 0 && k < aff.nz - 1;
                if (reg[u])
                {
                    const int64_t fs = 3 * int64_t(c) - int64_t(aff.tx) * (j + int64_t(aff.ny) * k) - int64_t(aff.ty) * k * aff.nx;
                    const int64_t f[6] = {fs - 3 * nxy + int64_t(aff.tx) * aff.ny + int64_t(aff.ty) * aff.nx + 2, fs - 3 * int64_t(aff.nx) + aff.tx + 1,
                                          fs - 3, fs, fs + 1, fs + 2};
                    // all loads first (a branch on a loaded value between them would serialise the DRAM round trips)
#pragma unroll
                    for (int e = 0; e < 6; ++e) F[u][e] = op.faceFlux[f[e]];
                    v[u] = V[c];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            if (!reg[u]) continue;
            // one test for the exact-|F| range; the sqrt(F * F) form only runs for a cell that has a face outside it
            bool exact = true;
#pragma unroll
            for (int e = 0; e < 6; ++e) exact = exact && CoNumOp::abs_is_exact(F[u][e]);
            double acc = 0.0;
            if (exact)
            {
#pragma unroll
                for (int e = 0; e < 6; ++e) acc += fabs(F[u][e]);
            }
            else
            {
#pragma unroll
                for (int e = 0; e < 6; ++e) acc += CoNumOp::mag(F[u][e]); // (unrolled: a runtime index would put F in local memory)
            }
            lmax = fmax(lmax, acc / v[u]);
            lphi += acc;
            lvol += v[u];
        }
    }
    conum_block_partials(lmax, lphi, lvol, partial);
}
// stage 1b: cells through their stencil -- the plan's irregular-cell list (spread over all threads) next to k_conum_regular, or
// every cell of a mesh without the affine proof (list == nullptr)
__global__ void __launch_bounds__(256)
k_conum_list(CoNumOp op, int n, const int* __restrict__ list, const int* __restrict__ seg, const int* __restrict__ ent,
             const double* __restrict__ V, double* __restrict__ partial)
{
    double lmax = -1.7976931348623157e308, lphi = 0.0, lvol = 0.0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x)
    {
        const int c = list ? list[idx] : idx;
        const double acc = conum_cell_generic(op, c, seg, ent);
        const double v = V[c];
        lmax = fmax(lmax, acc / v);
        lphi += acc;
        lvol += v;
    }
    conum_block_partials(lmax, lphi, lvol, partial);
}

__global__ void __launch_bounds__(256)
k_conum_stage2(int nPartial, const double* __restrict__ partial, double dt, double* __restrict__ result)
{
    __shared__ double sMax[8], sPhi[8], sVol[8];
    double lmax = -1.7976931348623157e308, lphi = 0.0, lvol = 0.0;
    for (int i = threadIdx.x; i < nPartial; i += blockDim.x)
    {
        lmax = fmax(lmax, partial[3 * i]); lphi += partial[3 * i + 1]; lvol += partial[3 * i + 2];
    }
    lmax = warp_max(lmax); lphi = warp_sum(lphi); lvol = warp_sum(lvol);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sMax[wid] = lmax; sPhi[wid] = lphi; sVol[wid] = lvol; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int i = 1; i < 8; ++i) { lmax = fmax(lmax, sMax[i]); lphi += sPhi[i]; lvol += sVol[i]; }
        result[0] = lmax * 0.5 * dt;            // coNum.cpp:89
        result[1] = 0.5 * (lphi / lvol) * dt;   // coNum.cpp:90
    }
}

// ---- boundary conditions -------------------------------------------------------------------------
struct BcPlan
{
    int nPatches;
    int offsets[FVK_MAX_PATCHES + 1];
    int kind[FVK_MAX_PATCHES];
    double value[FVK_MAX_PATCHES * 3];
};

template <class VT, int NC>
__global__ void __launch_bounds__(256)
k_correct_bcs(BcPlan plan, int nB, const int* __restrict__ faceCells, const double* __restrict__ bDeltaCoeffs,
              const double* __restrict__ internal, double* __restrict__ value, double* __restrict__ refValue,
              double* __restrict__ valueFraction, double* __restrict__ refGrad)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nB) return;
    int p = 0;
    while (p + 1 < plan.nPatches && b >= plan.offsets[p + 1]) ++p;
    const int kind = plan.kind[p];
    const typename VT::T cst = VT::ld(plan.value, p);
    if (kind == FVK_BC_FIXED_VALUE)
    { // fixedValue.hpp:33-42
        VT::st(refValue, b, cst); VT::st(value, b, cst); valueFraction[b] = 1.0; VT::st(refGrad, b, cst);
    }
    else if (kind == FVK_BC_FIXED_GRADIENT)
    { // fixedGradient.hpp:42-53: value = internal + grad * (1/deltaCoeffs)
        VT::st(refGrad, b, cst);
        const double inv = 1 / bDeltaCoeffs[b];
        VT::st(value, b, VT::add(VT::ld(internal, faceCells[b]), VT::mul(inv, cst)));
        valueFraction[b] = 0.0; VT::st(refValue, b, VT::zero());
    }
    else if (kind == FVK_BC_EXTRAPOLATED)
    { // extrapolated.hpp:40-52
        const typename VT::T v = VT::ld(internal, faceCells[b]);
        VT::st(value, b, v); valueFraction[b] = 1.0; VT::st(refValue, b, v); VT::st(refGrad, b, VT::zero());
    }
}

} // namespace

// ================================================================================================
// C ABI
// ================================================================================================
template <class VT>
static int div_impl(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                    const double* phiB, double coeff, const double* view, double* out, int mode, fvk_stream s)
{
    if (!m || !faceFlux || !phi || !out || (m->nBoundaryFaces && !phiB))
        return fvk_fail(FVK_EINVAL, "fvk_div: null argument");
    if (mode != FVK_SET && mode != FVK_ACC_SCALE && mode != FVK_ADD) return fvk_fail(FVK_EINVAL, "fvk_div: bad mode %d", mode);
    Scaling sc {m->V, view, coeff, false};
    if (scheme == FVK_LINEAR)
        return launch_gather(m, DivOp<VT, FVK_LINEAR> {faceFlux, m->weights, phi, phiB}, sc, out, mode, s);
    if (scheme == FVK_UPWIND)
        return launch_gather(m, DivOp<VT, FVK_UPWIND> {faceFlux, m->weights, phi, phiB}, sc, out, mode, s);
    return fvk_fail(FVK_EINVAL, "fvk_div: unknown scheme %d", scheme);
}

extern "C" int fvk_div_s(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                         const double* phiB, double coeff, const double* view, double* out, int mode, fvk_stream s)
{
    return div_impl<S1>(m, scheme, faceFlux, phi, phiB, coeff, view, out, mode, s);
}
extern "C" int fvk_div_v(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                         const double* phiB, double coeff, const double* view, double* out, int mode, fvk_stream s)
{
    return div_impl<S3>(m, scheme, faceFlux, phi, phiB, coeff, view, out, mode, s);
}

// forwardEuler (timeIntegration/forwardEuler.hpp:38-56) of `ddt(phi) + div(faceFlux, phi) = 0` in ONE pass: out = phiOld - dt * div, the
// source vector never goes to memory. phiOld must not alias out (neighbours' old values are read while out is written).
extern "C" int fvk_div_forward_euler_s(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phiOld, const double* phiB,
                                       double coeff, const double* view, double dt, double* out, fvk_stream s)
{
    if (!m || !faceFlux || !phiOld || !out || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_div_forward_euler_s: null argument");
    if (phiOld == out) return fvk_fail(FVK_EINVAL, "fvk_div_forward_euler_s: phiOld and out must be different arrays");
    Scaling sc {m->V, view, coeff, false};
    sc.epiA = phiOld; sc.epiScale = -dt;
    if (scheme == FVK_LINEAR) return launch_gather(m, DivOp<S1, FVK_LINEAR> {faceFlux, m->weights, phiOld, phiB}, sc, out, FVK_EPI_AXPY, s);
    if (scheme == FVK_UPWIND) return launch_gather(m, DivOp<S1, FVK_UPWIND> {faceFlux, m->weights, phiOld, phiB}, sc, out, FVK_EPI_AXPY, s);
    return fvk_fail(FVK_EINVAL, "fvk_div_forward_euler_s: unknown scheme %d", scheme);
}

// updateVelocity (pressureVelocityCoupling.cpp:199-213) fused with the gradient it consumes: U = HbyA - rAU * grad(p), no gradP
// round trip through memory; same arithmetic as fvk_grad_s followed by fvk_update_velocity
extern "C" int fvk_update_velocity_grad(const fvk_mesh* m, const double* HbyA, const double* rAU, const double* p, const double* pB,
                                        double* U, fvk_stream s)
{
    if (!m || !HbyA || !rAU || !p || !U || (m->nBoundaryFaces && !pB)) return fvk_fail(FVK_EINVAL, "fvk_update_velocity_grad: null argument");
    Scaling sc {m->V, nullptr, 1.0, true};
    sc.epiA = HbyA; sc.epiB = rAU;
    return launch_gather(m, GradOp {m->Sf, m->weights, p, pB}, sc, U, FVK_EPI_UPDATE_VELOCITY, s);
}
// dsl::solve's explicit source of ONE surfaceIntegrate operator folded into the right-hand side: rhs -= (0 + op) * V
// (Operator::explicitOperation into a zeroed source, then dsl/solver.hpp:73-77), no source vector in memory
extern "C" int fvk_rhs_sub_surface_integrate_s(const fvk_mesh* m, const double* flux, double coeff, const double* view, double* rhs,
                                               fvk_stream s)
{
    if (!m || !flux || !rhs) return fvk_fail(FVK_EINVAL, "fvk_rhs_sub_surface_integrate_s: null argument");
    Scaling sc {m->V, view, coeff, false};
    return launch_gather(m, SurfIntOp<S1> {flux}, sc, rhs, FVK_EPI_RHS_SUB, s);
}

#define FVK_PUBLIC_MODE(mode) do { if ((mode) != FVK_SET && (mode) != FVK_ACC_SCALE && (mode) != FVK_ADD) return fvk_fail(FVK_EINVAL, "%s: bad mode %d", __func__, (mode)); } while (0)
extern "C" int fvk_grad_s(const fvk_mesh* m, const double* phi, const double* phiB, double* out, int mode, fvk_stream s)
{
    if (!m || !phi || !out || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_grad_s: null argument");
    FVK_PUBLIC_MODE(mode);
    if (mode == FVK_ADD) return fvk_fail(FVK_EINVAL, "fvk_grad_s: mode FVK_ADD not defined for grad");
    Scaling sc {m->V, nullptr, 1.0, true};
    return launch_gather(m, GradOp {m->Sf, m->weights, phi, phiB}, sc, out, mode, s);
}

extern "C" int fvk_laplacian_s(const fvk_mesh* m, const double* phi, const double* phiB, double coeff,
                               const double* view, double* out, int mode, fvk_stream s)
{
    if (!m || !phi || !out || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_laplacian_s: null argument");
    FVK_PUBLIC_MODE(mode);
    Scaling sc {m->V, view, coeff, false};
    return launch_gather(m, LaplacianOp<S1> {m->magSf, m->nonOrthDeltaCoeffs, phi, phiB}, sc, out, mode, s);
}
extern "C" int fvk_laplacian_v(const fvk_mesh* m, const double* phi, const double* phiB, double coeff,
                               const double* view, double* out, int mode, fvk_stream s)
{
    if (!m || !phi || !out || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_laplacian_v: null argument");
    FVK_PUBLIC_MODE(mode);
    Scaling sc {m->V, view, coeff, false};
    return launch_gather(m, LaplacianOp<S3> {m->magSf, m->nonOrthDeltaCoeffs, phi, phiB}, sc, out, mode, s);
}

extern "C" int fvk_surface_integrate_s(const fvk_mesh* m, const double* flux, double coeff, const double* view,
                                       double* out, int mode, fvk_stream s)
{
    if (!m || !flux || !out) return fvk_fail(FVK_EINVAL, "fvk_surface_integrate_s: null argument");
    FVK_PUBLIC_MODE(mode);
    Scaling sc {m->V, view, coeff, false};
    return launch_gather(m, SurfIntOp<S1> {flux}, sc, out, mode, s);
}
extern "C" int fvk_surface_integrate_v(const fvk_mesh* m, const double* flux, double coeff, const double* view,
                                       double* out, int mode, fvk_stream s)
{
    if (!m || !flux || !out) return fvk_fail(FVK_EINVAL, "fvk_surface_integrate_v: null argument");
    FVK_PUBLIC_MODE(mode);
    Scaling sc {m->V, view, coeff, false};
    return launch_gather(m, SurfIntOp<S3> {flux}, sc, out, mode, s);
}

template <class VT>
static int interpolate_impl(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                            const double* phiB, double* outFace, fvk_stream s)
{
    if (!m || !phi || !outFace || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_interpolate: null argument");
    const int nF = m->nInternalFaces + m->nBoundaryFaces;
    const int grid = (nF + 255) / 256;
    if (grid == 0) return FVK_OK;
    if (scheme == FVK_LINEAR)
        k_interpolate<VT, FVK_LINEAR><<<grid, 256, 0, fvk_cu(s)>>>(m->nInternalFaces, nF, m->owner, m->neighbour,
                                                                   m->weights, faceFlux, phi, phiB, outFace);
    else if (scheme == FVK_UPWIND)
    {
        if (!faceFlux) return fvk_fail(FVK_EINVAL, "fvk_interpolate: upwind requires a faceFlux"); // upwind.hpp:66-72
        k_interpolate<VT, FVK_UPWIND><<<grid, 256, 0, fvk_cu(s)>>>(m->nInternalFaces, nF, m->owner, m->neighbour,
                                                                   m->weights, faceFlux, phi, phiB, outFace);
    }
    else
        return fvk_fail(FVK_EINVAL, "fvk_interpolate: unknown scheme %d", scheme);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_interpolate_s(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                                 const double* phiB, double* outFace, fvk_stream s)
{
    return interpolate_impl<S1>(m, scheme, faceFlux, phi, phiB, outFace, s);
}
extern "C" int fvk_interpolate_v(const fvk_mesh* m, int scheme, const double* faceFlux, const double* phi,
                                 const double* phiB, double* outFace, fvk_stream s)
{
    return interpolate_impl<S3>(m, scheme, faceFlux, phi, phiB, outFace, s);
}

extern "C" int fvk_interpolation_weights(const fvk_mesh* m, int scheme, const double* faceFlux, double* wFace,
                                         double* wB, fvk_stream s)
{
    if (!m || !wFace) return fvk_fail(FVK_EINVAL, "fvk_interpolation_weights: null argument");
    if (scheme != FVK_LINEAR && scheme != FVK_UPWIND) return fvk_fail(FVK_EINVAL, "fvk_interpolation_weights: unknown scheme");
    if (scheme == FVK_UPWIND && !faceFlux) return fvk_fail(FVK_EINVAL, "fvk_interpolation_weights: upwind requires a faceFlux");
    const int nF = m->nInternalFaces + m->nBoundaryFaces;
    const int grid = (nF + 255) / 256;
    if (grid == 0) return FVK_OK;
    k_weights<<<grid, 256, 0, fvk_cu(s)>>>(scheme, m->nInternalFaces, nF, m->weights, faceFlux, wFace, wB);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

template <class VT>
static int sngrad_impl(const fvk_mesh* m, const double* phi, const double* phiB, double* outFace, fvk_stream s)
{
    if (!m || !phi || !outFace || (m->nBoundaryFaces && !phiB)) return fvk_fail(FVK_EINVAL, "fvk_face_normal_grad: null argument");
    const int nF = m->nInternalFaces + m->nBoundaryFaces;
    const int grid = (nF + 255) / 256;
    if (grid == 0) return FVK_OK;
    k_sngrad<VT><<<grid, 256, 0, fvk_cu(s)>>>(m->nInternalFaces, nF, m->owner, m->neighbour, m->nonOrthDeltaCoeffs,
                                              phi, phiB, outFace);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_face_normal_grad_s(const fvk_mesh* m, const double* phi, const double* phiB, double* outFace, fvk_stream s)
{
    return sngrad_impl<S1>(m, phi, phiB, outFace, s);
}
extern "C" int fvk_face_normal_grad_v(const fvk_mesh* m, const double* phi, const double* phiB, double* outFace, fvk_stream s)
{
    return sngrad_impl<S3>(m, phi, phiB, outFace, s);
}

static int conum_grid(int n, int perSm)
{
    const int want = (n + 255) / 256;
    const int cap = fvk_sm_count() * perSm;
    return want < cap ? (want > 0 ? want : 1) : cap;
}
extern "C" size_t fvk_conum_scratch_bytes(const fvk_mesh* m)
{
    if (!m) return 0;
    return sizeof(double) * 3 * size_t(148 * 16 > 2 * conum_grid(m->nOwned, 8) ? 148 * 16 : 2 * conum_grid(m->nOwned, 8));
}
extern "C" int fvk_conum(const fvk_mesh* m, const double* faceFlux, double dt, double* result_d, void* scratch_d, fvk_stream s)
{
    if (!m || !faceFlux || !result_d || !scratch_d) return fvk_fail(FVK_EINVAL, "fvk_conum: null argument");
    double* partial = static_cast<double*>(scratch_d);
    const FvkBrickGeom& bg = m->bp.geom;
    const bool affine = !fvk_no_affine() && m->bp.nTiles > 0 && bg.affine && int64_t(bg.dims[0]) * bg.dims[1] * bg.dims[2] == m->nOwned;
    const CoNumOp op {faceFlux};
    int nPartial = 0;
    if (affine)
    {
        const CoNumAffine ca {1, bg.dims[0], bg.dims[1], bg.dims[2], bg.tUp[0], bg.tUp[1], m->bp.nIrr, m->bp.irrCells};
        const int gA = conum_grid((m->nOwned + 1) / 2, 4); // two cells per thread and pass, 4 resident blocks per SM
        k_conum_regular<<<gA, 256, 0, fvk_cu(s)>>>(op, m->nOwned, m->V, partial, ca);
        FVK_LAUNCH_CHECK();
        nPartial = gA;
        if (m->bp.nIrr > 0)
        {
            const int gB = conum_grid(m->bp.nIrr, 4);
            k_conum_list<<<gB, 256, 0, fvk_cu(s)>>>(op, m->bp.nIrr, m->bp.irrCells, m->stencilSeg, m->gatherEnt, m->V, partial + 3 * size_t(gA));
            FVK_LAUNCH_CHECK();
            nPartial += gB;
        }
    }
    else
    {
        const int g = conum_grid(m->nOwned, 8);
        k_conum_list<<<g, 256, 0, fvk_cu(s)>>>(op, m->nOwned, nullptr, m->stencilSeg, m->gatherEnt, m->V, partial);
        FVK_LAUNCH_CHECK();
        nPartial = g;
    }
    k_conum_stage2<<<1, 256, 0, fvk_cu(s)>>>(nPartial, partial, dt, result_d);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_correct_boundary_conditions(const fvk_mesh* m, int ncomp, const int32_t* kind_h, const double* value_h,
                                               const double* internal, double* bValue, double* bRefValue,
                                               double* bValueFraction, double* bRefGrad, fvk_stream s)
{
    if (!m || !kind_h || !value_h || !internal || !bValue || !bRefValue || !bValueFraction || !bRefGrad)
        return fvk_fail(FVK_EINVAL, "fvk_correct_boundary_conditions: null argument");
    if (ncomp != 1 && ncomp != 3) return fvk_fail(FVK_EINVAL, "fvk_correct_boundary_conditions: ncomp must be 1 or 3");
    const int nB = m->nBoundaryFaces;
    if (nB == 0) return FVK_OK;
    BcPlan plan;
    plan.nPatches = m->nPatches;
    for (int p = 0; p <= m->nPatches; ++p) plan.offsets[p] = m->patchOffsets[p];
    for (int p = 0; p < m->nPatches; ++p)
    {
        plan.kind[p] = kind_h[p];
        if (kind_h[p] < FVK_BC_CALCULATED || kind_h[p] > FVK_BC_EMPTY)
            return fvk_fail(FVK_EINVAL, "fvk_correct_boundary_conditions: unknown kind %d on patch %d", kind_h[p], p);
        if (kind_h[p] == FVK_BC_FIXED_GRADIENT && !m->bDeltaCoeffs)
            return fvk_fail(FVK_EINVAL, "fvk_correct_boundary_conditions: mesh has no boundary deltaCoeffs");
        for (int k = 0; k < ncomp; ++k) plan.value[ncomp * p + k] = value_h[ncomp * p + k];
    }
    const int grid = (nB + 255) / 256;
    if (ncomp == 1)
        k_correct_bcs<S1, 1><<<grid, 256, 0, fvk_cu(s)>>>(plan, nB, m->faceCells, m->bDeltaCoeffs, internal, bValue, bRefValue, bValueFraction, bRefGrad);
    else
        k_correct_bcs<S3, 3><<<grid, 256, 0, fvk_cu(s)>>>(plan, nB, m->faceCells, m->bDeltaCoeffs, internal, bValue, bRefValue, bValueFraction, bRefGrad);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
