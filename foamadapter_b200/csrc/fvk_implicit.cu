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

#include <cstdlib>

namespace
{
struct S1
{
    using T = double;
    static constexpr int NC = 1;
    static __device__ __forceinline__ T zero() { return 0.0; }
    static __device__ __forceinline__ T splat(double s) { return s; } // s * one<T>()
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
};

__device__ __forceinline__ double term_scaling(const fvk_term& t, int c)
{
    return t.coeffView ? t.coeffView[c] * t.coeff : t.coeff; // dsl/coeff.hpp:35
}

// coefficient a face contributes to the off-diagonal entry of row c (side 0: c owns f -> upper entry
// A[c][nei]; side 1: c is the neighbour -> lower entry A[c][own]) and, negated, to the diagonal of the
// OTHER role. div: value1 = -w F (lower, and subtracted from the owner's diag), value2 = F (1 - w)
// (upper, and subtracted from the neighbour's diag); laplacian: flux for all four.
__device__ __forceinline__ void face_coeffs(const fvk_term& t, const AsmMesh& m, int f, double& lowerAndOwnDiag,
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
        const double flux = m.nodc[f] * t.faceField[f] * m.magSf[f];
        lowerAndOwnDiag = flux;
        upperAndNeiDiag = flux;
    }
}

template <class VT>
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
        T v = accumulate ? VT::ld(values, slot) : VT::zero();
        for (int k = 0; k < terms.n; ++k)
        {
            const fvk_term& t = terms.t[k];
            if (t.kind != FVK_TERM_DIV && t.kind != FVK_TERM_LAPLACIAN) continue;
            double lo, up;
            face_coeffs(t, m, f, lo, up);
            v = VT::add(v, VT::mul(term_scaling(t, c), VT::splat(side ? lo : up)));
        }
        VT::st(values, slot, v);
    }
    // ---- diagonal and rhs: term-major, faces ascending inside a term ------------------------------
    const int dslot = r0 + m.diagOffs[c];
    T d = accumulate ? VT::ld(values, dslot) : VT::zero();
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
                    face_coeffs(t, m, f, lo, up);
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
                        const double flux = t.faceField[f] * m.magSf[f];
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
    VT::st(values, dslot, d);
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
__device__ __forceinline__ void face_coeffs_k(const fvk_term& t, const AsmMesh& m, int f, double& lowerAndOwnDiag,
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
        const double flux = m.nodc[f] * t.faceField[f] * m.magSf[f];
        lowerAndOwnDiag = flux;
        upperAndNeiDiag = flux;
    }
}

template <class VT, int K0, int K1>
__global__ void __launch_bounds__(256)
k_assemble_fast(Terms terms, AsmMesh m, fvk_bfield bd, int ft0, int ft1, double* __restrict__ values,
                double* __restrict__ rhs, double* __restrict__ bcMatrix, double* __restrict__ bcRhs)
{
    using T = typename VT::T;
    constexpr int NC = VT::NC;
    constexpr bool HAS1 = K1 != 0;
    extern __shared__ double stageAll[];
    const int lane = threadIdx.x & 31;
    double* stage = stageAll + size_t(threadIdx.x >> 5) * ASM_CAPW * NC;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = c < m.nC;
    const int r0 = active ? m.rowOffs[c] : 0, r1 = active ? m.rowOffs[c + 1] : 0;
    const int s0 = active ? m.seg[c] : 0, s1 = active ? m.seg[c + 1] : 0;
    const int nInt = r1 - r0 - 1; // internal faces of the cell (-1 on inactive lanes)
    const int wbase = __shfl_sync(0xffffffffu, r0, 0);
    const int wend = __reduce_max_sync(0xffffffffu, r1);
    const bool staged = (wend - wbase) <= ASM_CAPW;
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
            if (NC == 1) q[0] = reinterpret_cast<const double&>(v);
            else { const double* pv = reinterpret_cast<const double*>(&v); q[0] = pv[0]; q[1] = pv[1]; q[2] = pv[2]; }
        }
        else
            VT::st(values, r0 + pos, v);
    };

    // ---- internal faces: coefficients once, off-diagonal entries out, diagonal contributions kept
    double lo0[ASM_E], up0[ASM_E], lo1[ASM_E], up1[ASM_E];
#pragma unroll
    for (int k = 0; k < ASM_E; ++k)
    {
        const int f = code[k] >> 1; // face 0 on padded lanes: valid address, value unused
        face_coeffs_k<K0>(t0, m, f, lo0[k], up0[k]);
        if (HAS1) face_coeffs_k<K1>(t1, m, f, lo1[k], up1[k]);
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
        face_coeffs_k<K0>(t0, m, f, lo, up);
        v = VT::add(v, VT::mul(os0, VT::splat(side ? lo : up)));
        if (HAS1)
        {
            face_coeffs_k<K1>(t1, m, f, lo, up);
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
                    face_coeffs(t, m, cd >> 1, lo, up);
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
                        const double flux = t.faceField[f] * m.magSf[f];
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

template <class VT>
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
                if (!t.faceField) return fvk_fail(FVK_EINVAL, "fvk_assemble: term %d: laplacian needs gamma", k);
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
                m->neighbourOffset, m->V, m->weights, m->nonOrthDeltaCoeffs, m->magSf, m->bDeltaCoeffs};
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
        const size_t shm = sizeof(double) * 8 * ASM_CAPW * VT::NC;
        const int k0 = terms_h[ft[0]].kind, k1 = nFace == 2 ? terms_h[ft[1]].kind : 0;
#define FVK_ASM_CASE(a, bb)                                                                                             \
    if (k0 == a && k1 == bb)                                                                                            \
    {                                                                                                                   \
        if (shm > 48 * 1024)                                                                                            \
            FVK_CUDA(cudaFuncSetAttribute(k_assemble_fast<VT, a, bb>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(shm))); \
        k_assemble_fast<VT, a, bb><<<grid, 256, shm, fvk_cu(s)>>>(T, am, b, ft[0], ft[1], values, rhs, bcMatrix, bcRhs); \
        FVK_LAUNCH_CHECK();                                                                                             \
        return FVK_OK;                                                                                                  \
    }
        FVK_ASM_CASE(FVK_TERM_DIV, 0) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, 0)
        FVK_ASM_CASE(FVK_TERM_DIV, FVK_TERM_LAPLACIAN) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_DIV)
        FVK_ASM_CASE(FVK_TERM_DIV, FVK_TERM_DIV) FVK_ASM_CASE(FVK_TERM_LAPLACIAN, FVK_TERM_LAPLACIAN)
#undef FVK_ASM_CASE
    }
    k_assemble<VT><<<grid, 256, 0, fvk_cu(s)>>>(T, am, b, values, rhs, bcMatrix, bcRhs, accumulate ? 1 : 0);
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
