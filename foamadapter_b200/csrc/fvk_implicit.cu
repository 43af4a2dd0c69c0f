// Implicit operator assembly into the CSR LinearSystem over the precomputed sparsity pattern.
//
// Reference: face loops with scattered read-modify-writes on `values` (two plain += on the
// off-diagonals, two atomic_sub on the diagonals per face), a boundary loop with atomics on diag/rhs,
// after a zero-fill of the whole system (gaussGreenDiv.cpp:155-262, gaussGreenLaplacian.cpp:76-177,
// ddtOperator.cpp:38-60, sourceTerm.cpp:37-55, linearSystem.hpp:140-186).
//
// Here ONE kernel assembles any ordered list of terms (ddt / div / laplacian / source), cell-centric:
// thread c owns row c and writes every entry of it exactly once -- no atomics, no zero-fill, and in
// fused mode no read of `values` at all. Accumulation order per entry equals the order obtained by
// applying the reference operators one after another with its SerialExecutor (terms in list order;
// inside a term, faces in ascending id, then boundary faces), so the result is bit-identical.
#include "fvk_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace
{
struct S1
{
    using T = double;
    static constexpr int NC = 1;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T splat(double s) { return s; } // s * one<T>()
    static __device__ __forceinline__ T splatRaw(double s) { return s; }
    static __device__ __forceinline__ double first(const T& v) { return v; }
    static __device__ __forceinline__ T ld(const double* __restrict__ p, int64_t i) { return p[i]; }
    static __device__ __forceinline__ void st(double* __restrict__ p, int64_t i, T v) { p[i] = v; }
    static __device__ __forceinline__ T add(T a, T b) { return a + b; }
    static __device__ __forceinline__ T sub(T a, T b) { return a - b; }
    static __device__ __forceinline__ T mul(double s, T a) { return s * a; }
};
struct S3
{
    using T = Vec3d;
    static constexpr int NC = 3;
    static __device__ __forceinline__ T zero() { return Vec3d {0.0, 0.0, 0.0}; }
    static __device__ __forceinline__ T splat(double s) { return Vec3d {1.0 * s, 1.0 * s, 1.0 * s}; }
    static __device__ __forceinline__ T splatRaw(double s) { return Vec3d {s, s, s}; }
    static __device__ __forceinline__ double first(const T& v) { return v.x; }
    static __device__ __forceinline__ T ld(const double* __restrict__ p, int64_t i) { return ld3(p, i); }
    static __device__ __forceinline__ void st(double* __restrict__ p, int64_t i, T v) { st3(p, i, v); }
    static __device__ __forceinline__ T add(T a, T b) { return Vec3d {a.x + b.x, a.y + b.y, a.z + b.z}; }
    static __device__ __forceinline__ T sub(T a, T b) { return Vec3d {a.x - b.x, a.y - b.y, a.z - b.z}; }
    static __device__ __forceinline__ T mul(double s, T a) { return Vec3d {a.x * s, a.y * s, a.z * s}; }
};

struct Terms
{
    int n;
    fvk_term t[FVK_MAX_TERMS];
};

struct AsmMesh
{
    int nC, nI;
    const int* __restrict__ seg;
    const int* __restrict__ ent;
    const int* __restrict__ rowOffs;
    const uint8_t* __restrict__ diagOffs;
    const uint8_t* __restrict__ ownOffs;
    const uint8_t* __restrict__ neiOffs;
    const double* __restrict__ V;
    const double* __restrict__ w;     // geometric weights
    const double* __restrict__ nodc;  // nonOrthDeltaCoeffs
    const double* __restrict__ magSf;
    const double* __restrict__ bDeltaCoeffs;
    const int* __restrict__ owner;     // only read by laplacian terms whose gamma is interpolated on the fly (gammaCell)
    const int* __restrict__ neighbour;
};

// gamma of a laplacian term at internal face f (owner own, neighbour nei): the given face field, or -- gammaCell set -- the
// linear interpolate of a cell field evaluated on the fly with computeLinearInterpolation's arithmetic
// (interpolation/linear.cpp:30-45: w phi_P + (1 - w) phi_N), so `laplacian(interpolate(rAU), p)` needs no face-sized temporary
__device__ __forceinline__ double lap_gamma(const fvk_term& t, const AsmMesh& m, int f, int own, int nei)
{
    if (!t.gammaCell) return t.faceField[f];
    const double wf = m.w[f];
    return wf * t.gammaCell[own] + (1 - wf) * t.gammaCell[nei];
}
__device__ __forceinline__ double lap_gamma_boundary(const fvk_term& t, const AsmMesh& m, int f, int b)
{
    return t.gammaCell ? m.w[f] * t.gammaBoundary[b] : t.faceField[f]; // linear.cpp: boundary weight (= 1) x boundary value
}

__device__ __forceinline__ double term_scaling(const fvk_term& t, int c)
{
    return t.coeffView ? t.coeffView[c] * t.coeff : t.coeff; // dsl/coeff.hpp:35
}

// coefficient a face contributes to the off-diagonal entry of row c (side 0: c owns f -> upper entry
// A[c][nei]; side 1: c is the neighbour -> lower entry A[c][own]) and, negated, to the diagonal of the
// OTHER role. div: value1 = -w F (lower, and subtracted from the owner's diag), value2 = F (1 - w)
// (upper, and subtracted from the neighbour's diag); laplacian: flux for all four.
__device__ __forceinline__ void face_coeffs(const fvk_term& t, const AsmMesh& m, int f, int c, bool side, double& lowerAndOwnDiag,
                                            double& upperAndNeiDiag)
{
    if (t.kind == FVK_TERM_DIV)
    {
        const double F = t.faceField[f];
        const double wf = (t.scheme == FVK_LINEAR) ? m.w[f] : (F >= 0 ? 1.0 : 0.0);
        lowerAndOwnDiag = -wf * F;
        upperAndNeiDiag = F * (1 - wf);
    }
    else
    {
        const double g = t.gammaCell ? lap_gamma(t, m, f, side ? m.owner[f] : c, side ? c : m.neighbour[f]) : t.faceField[f];
        const double flux = m.nodc[f] * g * m.magSf[f];
        lowerAndOwnDiag = flux;
        upperAndNeiDiag = flux;
    }
}

// matrix value access: COMPACT (Vec3 systems only) stores the identical components of an entry once (values double[nnz])
template <class VT, bool COMPACT>
__device__ __forceinline__ typename VT::T ldval(const double* __restrict__ values, int64_t i)
{
    if (COMPACT) return VT::splatRaw(values[i]);
    return VT::ld(values, i);
}
template <class VT, bool COMPACT>
__device__ __forceinline__ void stval(double* __restrict__ values, int64_t i, const typename VT::T& v)
{
    if (COMPACT) values[i] = VT::first(v);
    else VT::st(values, i, v);
}

template <class VT, bool COMPACT>
__global__ void __launch_bounds__(256)
k_assemble(Terms terms, AsmMesh m, fvk_bfield bd, double* __restrict__ values, double* __restrict__ rhs,
           double* __restrict__ bcMatrix, double* __restrict__ bcRhs, int accumulate)
{
    using T = typename VT::T;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.nC) return;
    const int r0 = m.rowOffs[c];
    const int s0 = m.seg[c], s1 = m.seg[c + 1];
    // ---- off-diagonal entries: one walk over the faces, terms innermost ---------------------------
    for (int e = s0; e < s1; ++e)
    {
        const int code = m.ent[e];
        const int f = code >> 1;
        if (f >= m.nI) break; // boundary faces come last
        const int side = code & 1;
        const int slot = r0 + (side ? m.neiOffs[f] : m.ownOffs[f]);
        T v = accumulate ? ldval<VT, COMPACT>(values, slot) : VT::zero();
        for (int k = 0; k < terms.n; ++k)
        {
            const fvk_term& t = terms.t[k];
            if (t.kind != FVK_TERM_DIV && t.kind != FVK_TERM_LAPLACIAN) continue;
            double lo, up;
            face_coeffs(t, m, f, c, side, lo, up);
            v = VT::add(v, VT::mul(term_scaling(t, c), VT::splat(side ? lo : up)));
        }
        stval<VT, COMPACT>(values, slot, v);
    }
    // ---- diagonal and rhs: term-major, faces ascending inside a term ------------------------------
    const int dslot = r0 + m.diagOffs[c];
    T d = accumulate ? ldval<VT, COMPACT>(values, dslot) : VT::zero();
    T r = accumulate ? VT::ld(rhs, c) : VT::zero();
    for (int k = 0; k < terms.n; ++k)
    {
        const fvk_term& t = terms.t[k];
        const double os = term_scaling(t, c);
        if (t.kind == FVK_TERM_DIV || t.kind == FVK_TERM_LAPLACIAN)
        {
            for (int e = s0; e < s1; ++e)
            {
                const int code = m.ent[e];
                const int f = code >> 1;
                if (f < m.nI)
                {
                    double lo, up;
                    face_coeffs(t, m, f, c, code & 1, lo, up);
                    // owner's diag -= value1 * os ; neighbour's diag -= value2 * os
                    d = VT::sub(d, VT::mul(os, VT::splat((code & 1) ? up : lo)));
                }
                else
                {
                    const int b = f - m.nI;
                    const double vf1 = bd.valueFraction[b];
                    T valueMat, valueRhs;
                    if (t.kind == FVK_TERM_DIV)
                    { // gaussGreenDiv.cpp:237-260 (boundary weight of both schemes is 1)
                        const double flux = 1.0 * t.faceField[f];
                        const double vf2 = 1.0 - vf1;
                        valueMat = VT::splat(flux * os * vf2);
                        d = VT::add(d, valueMat);
                        valueRhs = VT::add(VT::mul(flux * os, VT::mul(vf1, VT::ld(bd.refValue, b))),
                                           VT::mul(1 / m.bDeltaCoeffs[b], VT::mul(vf2, VT::ld(bd.refGrad, b))));
                    }
                    else
                    { // gaussGreenLaplacian.cpp:156-175
                        const double flux = lap_gamma_boundary(t, m, f, b) * m.magSf[f];
                        const double dcf = m.nodc[f];
                        valueMat = VT::splat(flux * os * vf1 * dcf);
                        d = VT::sub(d, valueMat);
                        valueRhs = VT::mul(flux * os, VT::add(VT::mul(vf1 * dcf, VT::ld(bd.refValue, b)),
                                                              VT::mul(1.0 - vf1, VT::ld(bd.refGrad, b))));
                    }
                    r = VT::sub(r, valueRhs);
                    VT::st(bcMatrix, b, valueMat);
                    VT::st(bcRhs, b, valueRhs);
                }
            }
        }
        else if (t.kind == FVK_TERM_DDT)
        { // ddtOperator.cpp:50-59
            const double dtInver = 1.0 / t.dt;
            const double commonCoef = os * m.V[c] * dtInver;
            d = VT::add(d, VT::splat(commonCoef));
            r = VT::add(r, VT::mul(commonCoef, VT::ld(t.cellField, c)));
        }
        else if (t.kind == FVK_TERM_SOURCE)
        { // sourceTerm.cpp:46-54
            d = VT::add(d, VT::splat(os * t.cellField[c] * m.V[c]));
        }
    }
    stval<VT, COMPACT>(values, dslot, d);
    VT::st(rhs, c, r);
}

// ---- fast path: fresh system, rows laid out like the stencil, at most two face terms -----------------------------
// Same arithmetic and accumulation order as k_assemble, organised for the memory system:
//  * the first E internal entries of the cell->face stencil are fetched up front and the face loop is fully unrolled
//    with the kinds of the (<= 2) face terms as template parameters, so all face operands of a row are in flight
//    together instead of one dependent load chain per face and term;
//  * the coefficient each face contributes to the diagonal is kept in registers: the term-major diagonal pass reads
//    no memory for the first E faces;
//  * rows of a mesh in OpenFOAM face order are [lower | diag | upper] in stencil order, so an entry's slot is its
//    position: ownerOffset/neighbourOffset are not read;
//  * a warp's 32 rows are contiguous in CSR: entries are parked in shared memory and written with coalesced stores
//    (the per-thread 56-byte row stride costs 4x the L2 write transactions).
constexpr int ASM_E = 6;      // stencil entries handled in registers
constexpr int ASM_CAPW = 288; // staged entries per warp (32 rows x 9)

template <int KIND>
__device__ __forceinline__ void face_coeffs_k(const fvk_term& t, const AsmMesh& m, int f, int own, int nei, double& lowerAndOwnDiag,
                                              double& upperAndNeiDiag)
{
    if (KIND == FVK_TERM_DIV)
    {
        const double F = t.faceField[f];
        const double wf = (t.scheme == FVK_LINEAR) ? m.w[f] : (F >= 0 ? 1.0 : 0.0);
        lowerAndOwnDiag = -wf * F;
        upperAndNeiDiag = F * (1 - wf);
    }
    else
    {
        const double flux = m.nodc[f] * lap_gamma(t, m, f, own, nei) * m.magSf[f];
        lowerAndOwnDiag = flux;
        upperAndNeiDiag = flux;
    }
}
// own / nei of stencil entry (f, side) of cell c; the arrays are only read when a term interpolates its gamma on the fly
__device__ __forceinline__ void face_cells(const fvk_term& t0, const fvk_term& t1, const AsmMesh& m, int f, int c, bool side, int& own, int& nei)
{
    own = nei = c;
    if (t0.gammaCell || t1.gammaCell)
    {
        if (side) own = m.owner[f];
        else nei = m.neighbour[f];
    }
}

template <class VT, int K0, int K1, bool COMPACT>
__global__ void __launch_bounds__(256)
k_assemble_fast(Terms terms, AsmMesh m, fvk_bfield bd, int ft0, int ft1, double* __restrict__ values,
                double* __restrict__ rhs, double* __restrict__ bcMatrix, double* __restrict__ bcRhs,
                const int* __restrict__ cellList = nullptr, int nList = 0, int tailFirst = 0, int nTail = 0)
{
    using T = typename VT::T;
    constexpr int NC = COMPACT ? 1 : VT::NC; // doubles per stored matrix entry
    constexpr bool HAS1 = K1 != 0;
    extern __shared__ double stageAll[];
    const int lane = threadIdx.x & 31;
    double* stage = stageAll + size_t(threadIdx.x >> 5) * ASM_CAPW * NC;
    // cellList mode (the rows k_assemble_affine leaves out): entry idx < nList is cellList[idx], the following nTail
    // entries are the rows tailFirst.. (ghost rows of a decomposed mesh); rows are not contiguous -> direct stores
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = cellList ? idx < nList + nTail : idx < m.nC;
    const int c = !cellList ? idx : (idx < nList ? cellList[active ? idx : 0] : tailFirst + (idx - nList));
    const int r0 = active ? m.rowOffs[c] : 0, r1 = active ? m.rowOffs[c + 1] : 0;
    const int s0 = active ? m.seg[c] : 0, s1 = active ? m.seg[c + 1] : 0;
    const int nInt = r1 - r0 - 1; // internal faces of the cell (-1 on inactive lanes)
    const int wbase = __shfl_sync(0xffffffffu, r0, 0);
    const int wend = __reduce_max_sync(0xffffffffu, r1);
    const bool staged = !cellList && (wend - wbase) <= ASM_CAPW;
    int code[ASM_E];
#pragma unroll
    for (int k = 0; k < ASM_E; ++k) code[k] = (k < nInt) ? m.ent[s0 + k] : 0;
    const int nLower = active ? int(m.diagOffs[c]) : 0;
    const fvk_term& t0 = terms.t[ft0];
    const fvk_term& t1 = terms.t[HAS1 ? ft1 : ft0];
    const double os0 = active ? term_scaling(t0, c) : 0.0;
    const double os1 = (HAS1 && active) ? term_scaling(t1, c) : 0.0;

    auto put = [&](int pos, const T& v) {
        if (staged)
        {
            double* q = stage + size_t(r0 - wbase + pos) * NC;
            if (NC == 1) q[0] = VT::first(v);
            else { const double* pv = reinterpret_cast<const double*>(&v); q[0] = pv[0]; q[1] = pv[1]; q[2] = pv[2]; }
        }
        else
            stval<VT, COMPACT>(values, r0 + pos, v);
    };

    // ---- internal faces: coefficients once, off-diagonal entries out, diagonal contributions kept
    double lo0[ASM_E], up0[ASM_E], lo1[ASM_E], up1[ASM_E];
#pragma unroll
    for (int k = 0; k < ASM_E; ++k)
    {
        const int f = code[k] >> 1; // face 0 on padded lanes: valid address, value unused
        int own, nei;
        face_cells(t0, t1, m, f, active ? c : 0, code[k] & 1, own, nei);
        face_coeffs_k<K0>(t0, m, f, own, nei, lo0[k], up0[k]);
        if (HAS1) face_coeffs_k<K1>(t1, m, f, own, nei, lo1[k], up1[k]);
    }
    double dc0[ASM_E], dc1[ASM_E];
#pragma unroll
    for (int k = 0; k < ASM_E; ++k)
    {
        const bool side = code[k] & 1;
        T v = VT::zero();
        v = VT::add(v, VT::mul(os0, VT::splat(side ? lo0[k] : up0[k])));
        if (HAS1) v = VT::add(v, VT::mul(os1, VT::splat(side ? lo1[k] : up1[k])));
        dc0[k] = side ? up0[k] : lo0[k];
        dc1[k] = HAS1 ? (side ? up1[k] : lo1[k]) : 0.0;
        if (k < nInt) put(side ? k : k + 1, v);
    }
    for (int k = ASM_E; k < nInt; ++k) // polyhedral cells with more than ASM_E internal faces
    {
        const int cd = m.ent[s0 + k], f = cd >> 1;
        const bool side = cd & 1;
        double lo, up;
        T v = VT::zero();
        int own, nei;
        face_cells(t0, t1, m, f, c, side, own, nei);
        face_coeffs_k<K0>(t0, m, f, own, nei, lo, up);
        v = VT::add(v, VT::mul(os0, VT::splat(side ? lo : up)));
        if (HAS1)
        {
            face_coeffs_k<K1>(t1, m, f, own, nei, lo, up);
            v = VT::add(v, VT::mul(os1, VT::splat(side ? lo : up)));
        }
        put(side ? k : k + 1, v);
    }
    // ---- diagonal and rhs: term-major, faces ascending inside a term (identical order to k_assemble)
    T d = VT::zero(), r = VT::zero();
    if (active)
        for (int kt = 0; kt < terms.n; ++kt)
        {
            const fvk_term& t = terms.t[kt];
            if (t.kind == FVK_TERM_DIV || t.kind == FVK_TERM_LAPLACIAN)
            {
                const bool second = HAS1 && kt == ft1;
                const double os = second ? os1 : os0;
#pragma unroll
                for (int k = 0; k < ASM_E; ++k)
                    if (k < nInt) d = VT::sub(d, VT::mul(os, VT::splat(second ? dc1[k] : dc0[k])));
                for (int k = ASM_E; k < nInt; ++k)
                {
                    const int cd = m.ent[s0 + k];
                    double lo, up;
                    face_coeffs(t, m, cd >> 1, c, cd & 1, lo, up);
                    d = VT::sub(d, VT::mul(os, VT::splat((cd & 1) ? up : lo)));
                }
                for (int e = s0 + nInt; e < s1; ++e)
                {
                    const int f = m.ent[e] >> 1;
                    const int b = f - m.nI;
                    const double vf1 = bd.valueFraction[b];
                    T valueMat, valueRhs;
                    if (t.kind == FVK_TERM_DIV)
                    { // gaussGreenDiv.cpp:237-260 (boundary weight of both schemes is 1)
                        const double flux = 1.0 * t.faceField[f];
                        const double vf2 = 1.0 - vf1;
                        valueMat = VT::splat(flux * os * vf2);
                        d = VT::add(d, valueMat);
                        valueRhs = VT::add(VT::mul(flux * os, VT::mul(vf1, VT::ld(bd.refValue, b))),
                                           VT::mul(1 / m.bDeltaCoeffs[b], VT::mul(vf2, VT::ld(bd.refGrad, b))));
                    }
                    else
                    { // gaussGreenLaplacian.cpp:156-175
                        const double flux = lap_gamma_boundary(t, m, f, b) * m.magSf[f];
                        const double dcf = m.nodc[f];
                        valueMat = VT::splat(flux * os * vf1 * dcf);
                        d = VT::sub(d, valueMat);
                        valueRhs = VT::mul(flux * os, VT::add(VT::mul(vf1 * dcf, VT::ld(bd.refValue, b)),
                                                              VT::mul(1.0 - vf1, VT::ld(bd.refGrad, b))));
                    }
                    r = VT::sub(r, valueRhs);
                    VT::st(bcMatrix, b, valueMat);
                    VT::st(bcRhs, b, valueRhs);
                }
            }
            else if (t.kind == FVK_TERM_DDT)
            { // ddtOperator.cpp:50-59
                const double os = term_scaling(t, c);
                const double dtInver = 1.0 / t.dt;
                const double commonCoef = os * m.V[c] * dtInver;
                d = VT::add(d, VT::splat(commonCoef));
                r = VT::add(r, VT::mul(commonCoef, VT::ld(t.cellField, c)));
            }
            else if (t.kind == FVK_TERM_SOURCE)
            { // sourceTerm.cpp:46-54
                const double os = term_scaling(t, c);
                d = VT::add(d, VT::splat(os * t.cellField[c] * m.V[c]));
            }
        }
    if (active)
    {
        put(nLower, d);
        VT::st(rhs, c, r);
    }
    if (staged)
    {
        __syncwarp();
        const int n = (wend - wbase) * NC;
        double* __restrict__ dst = values + size_t(wbase) * NC;
        for (int i = lane; i < n; i += 32) dst[i] = stage[i];
    }
}


// ---- block-structured meshes: index-free assembly of the regular rows ---------------------------------------------------
// When the mesh plan proved the block topology (FvkBrickGeom::affine, fvk_brickplan.cpp), a REGULAR cell c = i + nx (j + ny k)
// (not in the outermost layer) has the stencil [zL, yL, xL | x, y, z] with arithmetic face ids (fs(c) .. fs(c)+2 owned,
// fs(c-1), fs(c-nx)+1, fs(c-nx ny)+2 lower) and its CSR row is [3 lower | diag | 3 upper]: no stencil, offset or diagOffset
// array is read. A block owns a 32 x BY x BZ brick (one warp = one x-run of 32 cells): every cell evaluates the
// coefficients of the three faces it owns ONCE into shared memory (the faces on the brick's three lower sides by spare
// threads), so every face operand is read from DRAM once and all loads of a thread are independent; after one barrier a
// regular cell folds its six faces in the reference's order (same arithmetic as k_assemble / k_assemble_fast: bit-identical)
// and the warp writes its 32 rows (contiguous in CSR) with coalesced stores. Vec3 systems have identical components
// (SURVEY A.3): the row is computed once and written replicated. Irregular cells (boundary / cut layers) and ghost rows go
// through k_assemble_fast's cell-list mode.
struct AsmAffine
{
    int nx, ny, nz, tx, ty;
    int tdx, tdy; // tiles along x, y
};
constexpr int AFF_LX = 32;

template <class VT, int K0, int K1, int BY, int BZ, int MINB, bool COMPACT>
__global__ void __launch_bounds__(AFF_LX * BY * BZ, MINB)
k_assemble_affine(Terms terms, AsmMesh m, AsmAffine g, int ft0, int ft1, double* __restrict__ values, double* __restrict__ rhs)
{
    using T = typename VT::T;
    constexpr int NC = COMPACT ? 1 : VT::NC; // doubles per stored matrix entry
    constexpr bool HAS1 = K1 != 0;
    constexpr int NCO = HAS1 ? 4 : 2; // doubles per face slot: {lo0, up0[, lo1, up1]}
    constexpr int TB = AFF_LX * BY * BZ;
    constexpr int XB = 3 * TB;
    extern __shared__ __align__(16) double smemA[];
    double* coef = smemA;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int64_t nxy = int64_t(nx) * ny;
    const int ix = blockIdx.x % g.tdx, q = blockIdx.x / g.tdx, iy = q % g.tdy, iz = q / g.tdy;
    const int x0 = ix * AFF_LX, y0 = iy * BY, z0 = iz * BZ;
    const int rl = min(AFF_LX, nx - x0), ry = min(BY, ny - y0), rz = min(BZ, nz - z0);
    const int off = lane, r = tid >> 5;
    const int a = (ry == BY) ? (r % BY) : (r % ry), b = (ry == BY) ? (r / BY) : (r / ry);
    const bool valid = off < rl && b < rz;
    const int i = x0 + off, j = y0 + a, k = z0 + b;
    const bool upper = valid && i < nx - 1 && j < ny - 1 && k < nz - 1;
    const bool regular = upper && i > 0 && j > 0 && k > 0;
    const int64_t cell = i + int64_t(nx) * j + nxy * k;
    const int64_t fs = 3 * cell - int64_t(g.tx) * (j + int64_t(ny) * k) - int64_t(g.ty) * k * nx;
    const fvk_term& t0 = terms.t[ft0];
    const fvk_term& t1 = terms.t[HAS1 ? ft1 : ft0];
    // ---- the three faces this cell owns: all loads independent, coefficients into shared memory
    if (upper)
    {
        double lo0[3], up0[3], lo1[3], up1[3];
#pragma unroll
        for (int f = 0; f < 3; ++f)
        {
            const int nb = int(cell) + (f == 0 ? 1 : (f == 1 ? nx : int(nxy)));
            face_coeffs_k<K0>(t0, m, int(fs) + f, int(cell), nb, lo0[f], up0[f]);
            if (HAS1) face_coeffs_k<K1>(t1, m, int(fs) + f, int(cell), nb, lo1[f], up1[f]);
        }
#pragma unroll
        for (int f = 0; f < 3; ++f)
        {
            double* s = coef + size_t(3 * tid + f) * NCO;
            s[0] = lo0[f]; s[1] = up0[f];
            if (HAS1) { s[2] = lo1[f]; s[3] = up1[f]; }
        }
    }
    // ---- per-cell operands of a regular row, issued before the barrier
    int r0 = 0;
    double os0 = 0.0, os1 = 0.0;
    if (regular)
    {
        r0 = m.rowOffs[cell];
        os0 = term_scaling(t0, int(cell));
        if (HAS1) os1 = term_scaling(t1, int(cell));
    }
    // ---- cross faces: e < nZ: z side (b = 0) | < nZ + nY: y side (a = 0) | x side (off = 0); only for regular consumers
    const int nZ = rl * ry, nY = rl * rz, nX = ry * rz, nCross = nZ + nY + nX;
    for (int e = tid; e < nCross; e += TB)
    {
        int co, ca, cb;
        int64_t dFace, dOwner;
        if (e < nZ) { co = e % rl; ca = e / rl; cb = 0; dFace = -3 * nxy + int64_t(g.tx) * ny + int64_t(g.ty) * nx + 2; dOwner = nxy; }
        else if (e < nZ + nY) { const int e1 = e - nZ; co = e1 % rl; cb = e1 / rl; ca = 0; dFace = -3 * int64_t(nx) + g.tx + 1; dOwner = nx; }
        else { const int e1 = e - nZ - nY; ca = e1 % ry; cb = e1 / ry; co = 0; dFace = -3; dOwner = 1; }
        const int ci = x0 + co, cj = y0 + ca, ck = z0 + cb;
        if (!(ci > 0 && ci < nx - 1 && cj > 0 && cj < ny - 1 && ck > 0 && ck < nz - 1)) continue; // consumer not regular
        const int64_t cc = ci + int64_t(nx) * cj + nxy * ck;
        const int64_t xf = 3 * cc - int64_t(g.tx) * (cj + int64_t(ny) * ck) - int64_t(g.ty) * ck * nx + dFace;
        double* s = coef + size_t(XB + e) * NCO;
        face_coeffs_k<K0>(t0, m, int(xf), int(cc - dOwner), int(cc), s[0], s[1]);
        if (HAS1) face_coeffs_k<K1>(t1, m, int(xf), int(cc - dOwner), int(cc), s[2], s[3]);
    }
    __syncthreads();
    // ---- regular rows: [zL, yL, xL | diag | x, y, z]
    double row[7];
    T rr = VT::zero();
    if (regular)
    {
        const int slot[6] = {b > 0 ? 3 * (tid - 32 * ry) + 2 : XB + off + rl * a,
                             a > 0 ? 3 * (tid - 32) + 1 : XB + nZ + off + rl * b,
                             off > 0 ? 3 * (tid - 1) : XB + nZ + nY + a + ry * b,
                             3 * tid, 3 * tid + 1, 3 * tid + 2};
        double dc0[6], dc1[6];
#pragma unroll
        for (int e = 0; e < 6; ++e)
        {
            const double* s = coef + size_t(slot[e]) * NCO;
            const bool side = e < 3; // the cell is the face's neighbour
            double v = 0.0 + os0 * (1.0 * (side ? s[0] : s[1]));
            if (HAS1) v = v + os1 * (1.0 * (side ? s[2] : s[3]));
            dc0[e] = side ? s[1] : s[0];
            dc1[e] = HAS1 ? (side ? s[3] : s[2]) : 0.0;
            row[side ? e : e + 1] = v;
        }
        // diagonal and rhs: term-major, faces ascending inside a term (identical order to k_assemble)
        double d = 0.0;
        for (int kt = 0; kt < terms.n; ++kt)
        {
            const fvk_term& t = terms.t[kt];
            if (t.kind == FVK_TERM_DIV || t.kind == FVK_TERM_LAPLACIAN)
            {
                const bool second = HAS1 && kt == ft1;
                const double os = second ? os1 : os0;
#pragma unroll
                for (int e = 0; e < 6; ++e) d = d - os * (1.0 * (second ? dc1[e] : dc0[e]));
            }
            else if (t.kind == FVK_TERM_DDT)
            { // ddtOperator.cpp:50-59
                const double os = term_scaling(t, int(cell));
                const double dtInver = 1.0 / t.dt;
                const double commonCoef = os * m.V[cell] * dtInver;
                d = d + 1.0 * commonCoef;
                rr = VT::add(rr, VT::mul(commonCoef, VT::ld(t.cellField, cell)));
            }
            else if (t.kind == FVK_TERM_SOURCE)
            { // sourceTerm.cpp:46-54
                const double os = term_scaling(t, int(cell));
                d = d + 1.0 * (os * t.cellField[cell] * m.V[cell]);
            }
        }
        row[3] = d;
        VT::st(rhs, cell, rr);
    }
    // ---- coalesced row store: the regular lanes of a warp are consecutive cells = consecutive CSR rows of 7 entries
    __syncthreads(); // every thread is done with the coefficient slots: the staging area reuses them
    double* stage = smemA + size_t(r) * (32 * 7);
    const unsigned regMask = __ballot_sync(0xffffffffu, regular);
    if (regMask == 0) return;
    const int first = __ffs(regMask) - 1;
    const int nReg = __popc(regMask);
    const int base = __shfl_sync(0xffffffffu, r0, first);
    if (regular)
    {
#pragma unroll
        for (int e = 0; e < 7; ++e) stage[(lane - first) * 7 + e] = row[e];
    }
    __syncwarp();
    const int n = nReg * 7 * NC;
    double* __restrict__ dst = values + size_t(base) * NC;
    for (int e = lane; e < n; e += 32) dst[e] = stage[NC == 1 ? e : e / 3];
}

// createEmptyLinearSystem's BoundaryCoefficients index arrays (linearSystem.hpp:163-174):
// matrixIdxs[b] = celli + diagOffset[celli] (sic), rhsIdxs[b] = celli
__global__ void __launch_bounds__(256)
k_bc_indices(int nB, const int* __restrict__ faceCells, const uint8_t* __restrict__ diagOffs,
             int* __restrict__ matrixIdxs, int* __restrict__ rhsIdxs)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nB) return;
    const int c = faceCells[b];
    matrixIdxs[b] = c + diagOffs[c];
    rhsIdxs[b] = c;
}

// explicit cell-wise terms: ddt (ddtOperator.cpp:29-35), source (sourceTerm.cpp:28-34),
// rhs -= src * V (dsl/solver.hpp:73-77)
template <class VT>
__global__ void __launch_bounds__(256)
k_ddt_exp(int nC, const double* __restrict__ V, const double* __restrict__ field, const double* __restrict__ old,
          double dt, double* __restrict__ source)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const double dtInver = 1.0 / dt;
    // source += dtInver * (field - old) * V
    const typename VT::T v = VT::mul(V[c], VT::mul(dtInver, VT::sub(VT::ld(field, c), VT::ld(old, c))));
    VT::st(source, c, VT::add(VT::ld(source, c), v));
}
template <class VT>
__global__ void __launch_bounds__(256)
k_source_exp(int nC, const double* __restrict__ k, const double* __restrict__ field, double coeff,
             const double* __restrict__ view, double* __restrict__ source)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const double os = view ? view[c] * coeff : coeff;
    VT::st(source, c, VT::add(VT::ld(source, c), VT::mul(os * k[c], VT::ld(field, c))));
}
template <class VT>
__global__ void __launch_bounds__(256)
k_rhs_sub_source(int nC, const double* __restrict__ V, const double* __restrict__ src, double* __restrict__ rhs)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    VT::st(rhs, c, VT::sub(VT::ld(rhs, c), VT::mul(V[c], VT::ld(src, c))));
}

// per-device auxiliary stream + events for the forked irregular-row pass (FVK_ASM_NO_FORK=1: same stream)
bool side_stream(cudaStream_t* st, cudaEvent_t* evFork, cudaEvent_t* evJoin)
{
    static const bool off = [] { const char* e = std::getenv("FVK_ASM_NO_FORK"); return e && *e == '1'; }();
    if (off) return false;
    struct Side { cudaStream_t st = nullptr; cudaEvent_t a = nullptr, b = nullptr; bool ok = false, tried = false; };
    static Side tab[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    Side& sd = tab[dev];
    if (!sd.tried)
    {
        sd.tried = true;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        sd.ok = cudaStreamCreateWithPriority(&sd.st, cudaStreamNonBlocking, hi) == cudaSuccess
                && cudaEventCreateWithFlags(&sd.a, cudaEventDisableTiming) == cudaSuccess
                && cudaEventCreateWithFlags(&sd.b, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!sd.ok) return false;
    *st = sd.st; *evFork = sd.a; *evJoin = sd.b;
    return true;
}

template <class VT, bool COMPACT = false>
int assemble_impl(const fvk_mesh* m, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd, double* values,
                  double* rhs, double* bcMatrix, double* bcRhs, int accumulate, fvk_stream s)
{
    if (!m || !terms_h || !values || !rhs) return fvk_fail(FVK_EINVAL, "fvk_assemble: null argument");
    if (nTerms < 1 || nTerms > FVK_MAX_TERMS) return fvk_fail(FVK_EINVAL, "fvk_assemble: nTerms must be 1..%d", FVK_MAX_TERMS);
    Terms T;
    T.n = nTerms;
    bool needsBoundary = false;
    for (int k = 0; k < nTerms; ++k)
    {
        const fvk_term& t = terms_h[k];
        switch (t.kind)
        {
            case FVK_TERM_DIV:
                if (t.scheme != FVK_LINEAR && t.scheme != FVK_UPWIND) return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: unknown scheme", k);
                if (!t.faceField) return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: div needs a faceFlux", k);
                needsBoundary = true;
                break;
            case FVK_TERM_LAPLACIAN:
                if (!t.faceField && !(t.gammaCell && (t.gammaBoundary || m->nBoundaryFaces == 0)))
                    return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: laplacian needs gamma (a face field, or gammaCell + gammaBoundary)", k);
                needsBoundary = true;
                break;
            case FVK_TERM_DDT:
                if (!t.cellField || !(t.dt != 0.0)) return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: ddt needs the old field and dt != 0", k);
                break;
            case FVK_TERM_SOURCE:
                if (!t.cellField) return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: source needs coefficients", k);
                break;
            default: return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: unknown kind %d", k, t.kind);
        }
        T.t[k] = t;
    }
    fvk_bfield b {nullptr, nullptr, nullptr, nullptr};
    if (needsBoundary && m->nBoundaryFaces > 0)
    {
        if (!bd || !bd->valueFraction || !bd->refValue || !bd->refGrad || !bcMatrix || !bcRhs)
            return fvk_fail(FVK_EINVAL, "fvk_assemble: div/laplacian need the field's boundary data and bcCoeffs arrays");
        if (!m->bDeltaCoeffs) return fvk_fail(FVK_EINVAL, "fvk_assemble: mesh has no boundary deltaCoeffs");
        b = *bd;
    }
    // all rows incl. ghost rows: a ghost row's off-diagonals over the cut faces are read by updateFaceVelocity
    AsmMesh am {m->nCells, m->nInternalFaces, m->stencilSeg, m->gatherEnt, m->rowOffs, m->diagOffset, m->ownerOffset,
                m->neighbourOffset, m->V, m->weights, m->nonOrthDeltaCoeffs, m->magSf, m->bDeltaCoeffs, m->owner, m->neighbour};
    const int grid = (m->nCells + 255) / 256;
    // fast path (k_assemble_fast): fresh system, rows in stencil order, one or two face terms
    int ft[2] = {-1, -1}, nFace = 0;
    for (int k = 0; k < nTerms; ++k)
        if (terms_h[k].kind == FVK_TERM_DIV || terms_h[k].kind == FVK_TERM_LAPLACIAN)
        {
            if (nFace < 2) ft[nFace] = k;
            ++nFace;
        }
    static const bool noFast = [] { const char* e = std::getenv("FVK_ASM_GENERIC"); return e && *e == '1'; }();
    if (!noFast && !accumulate && m->rowsInStencilOrder && nFace >= 1 && nFace <= 2)
    {
        const size_t shm = sizeof(double) * 8 * ASM_CAPW * (COMPACT ? 1 : VT::NC);
        const int k0 = terms_h[ft[0]].kind, k1 = nFace == 2 ? terms_h[ft[1]].kind : 0;
        // block-structured mesh with proven topology: index-free kernel for the regular rows + cell-list pass for the rest
        static const bool noAffine = [] { const char* e = std::getenv("FVK_ASM_NO_AFFINE"); return e && *e == '1'; }();
        const FvkBrickGeom& bg = m->bp.geom;
        if (!noAffine && !fvk_no_affine() && bg.affine && m->bp.nTiles > 0 && int64_t(bg.dims[0]) * bg.dims[1] * bg.dims[2] == m->nOwned && bg.dims[0] >= 3
            && bg.dims[1] >= 3 && bg.dims[2] >= 3)
        {
            // brick 32 x BY x BZ and the resident blocks the register allocation aims at; FVK_ASM_TILE="by,bz,minb" selects
            // another instantiated combination (roofline sweeps)
            // B200 sweep at 256^3 (profiles/r2c_sweep_asm_256.jsonl): 32x2x2 bricks; the two-face-term scalar kernel wants its
            // spill-free 6 resident blocks, the lighter ones 8
            int cfg[3] = {2, 2, (VT::NC == 1 && nFace == 2) ? 6 : 8};
            if (const char* e = std::getenv("FVK_ASM_TILE")) // read per call: sweeps change it in-process
            {
                int a_ = 0, b_ = 0, c_ = 0;
                if (std::sscanf(e, "%d,%d,%d", &a_, &b_, &c_) == 3) { cfg[0] = a_; cfg[1] = b_; cfg[2] = c_; }
            }
            const int nTail = m->nCells - m->nOwned, nListed = m->bp.nIrr + nTail;
#define FVK_ASMA_LAUNCH(a, bb, BY, BZ, MINB)                                                                            \
    if (cfg[0] == BY && cfg[1] == BZ && cfg[2] == MINB)                                                                 \
    {                                                                                                                   \
        constexpr int TB = AFF_LX * BY * BZ;                                                                            \
        AsmAffine ag {bg.dims[0], bg.dims[1], bg.dims[2], bg.tUp[0], bg.tUp[1], (bg.dims[0] + AFF_LX - 1) / AFF_LX, (bg.dims[1] + BY - 1) / BY}; \
        const int nTilesA = ag.tdx * ag.tdy * ((bg.dims[2] + BZ - 1) / BZ);                                             \
        const size_t shmA = sizeof(double) * std::max(size_t(3 * TB + AFF_LX * BY + AFF_LX * BZ + BY * BZ) * (nFace == 2 ? 4 : 2), size_t(TB) * 7); \
        if (shmA > 48 * 1024)                                                                                           \
            FVK_CUDA(cudaFuncSetAttribute(k_assemble_affine<VT, a, bb, BY, BZ, MINB, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shmA))); \
        /* the irregular rows (boundary / cut layers, ghost rows: latency-bound per-row walk) run BESIDE the streaming tiles on a \
           forked stream (events: also valid inside a stream capture) */                                               \
        cudaStream_t side = nullptr; cudaEvent_t evFork = nullptr, evJoin = nullptr;                                    \
        const bool forked = nListed > 0 && side_stream(&side, &evFork, &evJoin);                                        \
        if (forked)                                                                                                     \
        {                                                                                                               \
            FVK_CUDA(cudaEventRecord(evFork, fvk_cu(s)));                                                               \
            FVK_CUDA(cudaStreamWaitEvent(side, evFork, 0));                                                             \
        }                                                                                                               \
        if (nListed > 0)                                                                                                \
            k_assemble_fast<VT, a, bb, COMPACT><<<(nListed + 255) / 256, 256, 0, forked ? side : fvk_cu(s)>>>(T, am, b, ft[0], ft[1], values, rhs, bcMatrix, bcRhs, \
                                                                                     m->bp.irrCells, m->bp.nIrr, m->nOwned, nTail); \
        FVK_LAUNCH_CHECK();                                                                                             \
        k_assemble_affine<VT, a, bb, BY, BZ, MINB, COMPACT><<<nTilesA, TB, shmA, fvk_cu(s)>>>(T, am, ag, ft[0], ft[1], values, rhs); \
        FVK_LAUNCH_CHECK();                                                                                             \
        if (forked)                                                                                                     \
        {                                                                                                               \
            FVK_CUDA(cudaEventRecord(evJoin, side));                                                                    \
            FVK_CUDA(cudaStreamWaitEvent(fvk_cu(s), evJoin, 0));                                                        \
        }                                                                                                               \
        return FVK_OK;                                                                                                  \
    }
#define FVK_ASMA_CASE(a, bb)                                                                                            \
    if (k0 == a && k1 == bb)                                                                                            \
    {                                                                                                                   \
        FVK_ASMA_LAUNCH(a, bb, 4, 2, 4) FVK_ASMA_LAUNCH(a, bb, 4, 2, 3) FVK_ASMA_LAUNCH(a, bb, 2, 2, 8) FVK_ASMA_LAUNCH(a, bb, 2, 2, 6) \
        FVK_ASMA_LAUNCH(a, bb, 4, 4, 2) FVK_ASMA_LAUNCH(a, bb, 8, 2, 2) FVK_ASMA_LAUNCH(a, bb, 4, 1, 8)                  \
    }
            FVK_ASMA_CASE(FVK_TERM_DIV, 0) FVK_ASMA_CASE(FVK_TERM_LAPLACIAN, 0)
            FVK_ASMA_CASE(FVK_TERM_DIV, FVK_TERM_LAPLACIAN) FVK_ASMA_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_DIV)
            FVK_ASMA_CASE(FVK_TERM_DIV, FVK_TERM_DIV) FVK_ASMA_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_LAPLACIAN)
#undef FVK_ASMA_LAUNCH
#undef FVK_ASMA_CASE
        }
#define FVK_ASM_CASE(a, bb)                                                                                             \
    if (k0 == a && k1 == bb)                                                                                            \
    {                                                                                                                   \
        if (shm > 48 * 1024)                                                                                            \
            FVK_CUDA(cudaFuncSetAttribute(k_assemble_fast<VT, a, bb, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shm))); \
        k_assemble_fast<VT, a, bb, COMPACT><<<grid, 256, shm, fvk_cu(s)>>>(T, am, b, ft[0], ft[1], values, rhs, bcMatrix, bcRhs); \
        FVK_LAUNCH_CHECK();                                                                                             \
        return FVK_OK;                                                                                                  \
    }
        FVK_ASM_CASE(FVK_TERM_DIV, 0) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, 0)
        FVK_ASM_CASE(FVK_TERM_DIV, FVK_TERM_LAPLACIAN) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_DIV)
        FVK_ASM_CASE(FVK_TERM_DIV, FVK_TERM_DIV) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_LAPLACIAN)
#undef FVK_ASM_CASE
    }
    k_assemble<VT, COMPACT><<<grid, 256, 0, fvk_cu(s)>>>(T, am, b, values, rhs, bcMatrix, bcRhs, accumulate ? 1 : 0);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
} // namespace

extern "C" int fvk_assemble_s(const fvk_mesh* m, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                              double* values, double* rhs, double* bcMatrix, double* bcRhs, int accumulate, fvk_stream s)
{
    return assemble_impl<S1>(m, nTerms, terms_h, bd, values, rhs, bcMatrix, bcRhs, accumulate, s);
}
extern "C" int fvk_assemble_v(const fvk_mesh* m, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                              double* values, double* rhs, double* bcMatrix, double* bcRhs, int accumulate, fvk_stream s)
{
    return assemble_impl<S3>(m, nTerms, terms_h, bd, values, rhs, bcMatrix, bcRhs, accumulate, s);
}

// Vec3 system whose matrix entries have identical components (every implicit operator multiplies by one<Vec3>()):
// values double[nnz] holds each entry ONCE (a third of the HBM traffic of the Vec3 layout); rhs / bcMatrix / bcRhs stay Vec3.
extern "C" int fvk_assemble_vc(const fvk_mesh* m, int nTerms, const fvk_term* terms_h, const fvk_bfield* bd,
                               double* valuesCompact, double* rhs, double* bcMatrix, double* bcRhs, int accumulate, fvk_stream s)
{
    return assemble_impl<S3, true>(m, nTerms, terms_h, bd, valuesCompact, rhs, bcMatrix, bcRhs, accumulate, s);
}
namespace
{
__global__ void __launch_bounds__(256) k_expand_vec3(int64_t n, const double* __restrict__ in, double* __restrict__ out)
{
    for (int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x; i < 3 * n; i += int64_t(gridDim.x) * 256) out[i] = in[i / 3];
}
} // namespace
extern "C" int fvk_expand_vec3(int64_t n, const double* compact, double* outV, fvk_stream s)
{
    if (n < 0 || (n && (!compact || !outV))) return fvk_fail(FVK_EINVAL, "fvk_expand_vec3: bad argument");
    if (n == 0) return FVK_OK;
    const int64_t g = (3 * n + 255) / 256;
    k_expand_vec3<<<unsigned(g < 148 * 16 ? g : 148 * 16), 256, 0, fvk_cu(s)>>>(n, compact, outV);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_bc_coeff_indices(const fvk_mesh* m, int32_t* matrixIdxs, int32_t* rhsIdxs, fvk_stream s)
{
    if (!m || !matrixIdxs || !rhsIdxs) return fvk_fail(FVK_EINVAL, "fvk_bc_coeff_indices: null argument");
    if (m->nBoundaryFaces == 0) return FVK_OK;
    k_bc_indices<<<(m->nBoundaryFaces + 255) / 256, 256, 0, fvk_cu(s)>>>(m->nBoundaryFaces, m->faceCells, m->diagOffset,
                                                                         matrixIdxs, rhsIdxs);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

#define CELL_GRID(m) ((m)->nOwned + 255) / 256, 256, 0, fvk_cu(s)
extern "C" int fvk_ddt_explicit(const fvk_mesh* m, int ncomp, const double* field, const double* oldField, double dt,
                                double* source, fvk_stream s)
{
    if (!m || !field || !oldField || !source || (ncomp != 1 && ncomp != 3) || !(dt != 0.0))
        return fvk_fail(FVK_EINVAL, "fvk_ddt_explicit: bad argument");
    if (ncomp == 1) k_ddt_exp<S1><<<CELL_GRID(m)>>>(m->nOwned, m->V, field, oldField, dt, source);
    else k_ddt_exp<S3><<<CELL_GRID(m)>>>(m->nOwned, m->V, field, oldField, dt, source);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_source_explicit(const fvk_mesh* m, int ncomp, const double* k, const double* field, double coeff,
                                   const double* coeffView, double* source, fvk_stream s)
{
    if (!m || !k || !field || !source || (ncomp != 1 && ncomp != 3)) return fvk_fail(FVK_EINVAL, "fvk_source_explicit: bad argument");
    if (ncomp == 1) k_source_exp<S1><<<CELL_GRID(m)>>>(m->nOwned, k, field, coeff, coeffView, source);
    else k_source_exp<S3><<<CELL_GRID(m)>>>(m->nOwned, k, field, coeff, coeffView, source);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
extern "C" int fvk_rhs_sub_source(const fvk_mesh* m, int ncomp, const double* src, double* rhs, fvk_stream s)
{
    if (!m || !src || !rhs || (ncomp != 1 && ncomp != 3)) return fvk_fail(FVK_EINVAL, "fvk_rhs_sub_source: bad argument");
    if (ncomp == 1) k_rhs_sub_source<S1><<<CELL_GRID(m)>>>(m->nOwned, m->V, src, rhs);
    else k_rhs_sub_source<S3><<<CELL_GRID(m)>>>(m->nOwned, m->V, src, rhs);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
