// PISO pressure-velocity coupling kernels (FoamAdapter src/algorithms/pressureVelocityCoupling.cpp,
// include/FoamAdapter/datastructures/expression.hpp:86-112,181-199).
//
// The reference scatters the off-diagonal products of HbyA with 6 scalar atomics per face and
// allocates rAU / HbyA / flux temporaries per call. Here rAU and HbyA come from ONE cell-centric
// kernel (row c of the Vec3 momentum matrix is walked in ascending face order = the Serial
// executor's accumulation order, no atomics), and the face kernels are single fused passes.
#include "fvk_device.cuh"

namespace
{
__device__ __forceinline__ Vec3d vsub(Vec3d a, Vec3d b) { return Vec3d {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3d vadd(Vec3d a, Vec3d b) { return Vec3d {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec3d vscale(Vec3d a, double s) { return Vec3d {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double vdot(Vec3d a, Vec3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// computeRAU (:38-63) + computeRAUandHByA (:65-128), internal part
__global__ void __launch_bounds__(256)
k_rAU_HbyA(int nC, int nI, const int* __restrict__ seg, const int* __restrict__ ent, const int* __restrict__ owner,
           const int* __restrict__ neighbour, const int* __restrict__ rowOffs, const uint8_t* __restrict__ diagOffs,
           const uint8_t* __restrict__ ownOffs, const uint8_t* __restrict__ neiOffs, const double* __restrict__ V,
           const double* __restrict__ valuesV, const double* __restrict__ rhsV, const double* __restrict__ U,
           double* __restrict__ rAU, double* __restrict__ HbyA)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int r0 = rowOffs[c];
    const double vol = V[c];
    const double ra = vol / valuesV[3 * int64_t(r0 + diagOffs[c])]; // component [0] of the Vec3 diagonal
    rAU[c] = ra;
    if (!HbyA) return;
    Vec3d h {0.0, 0.0, 0.0};
    const int e1 = seg[c + 1];
    for (int e = seg[c]; e < e1; ++e)
    {
        const int code = ent[e];
        const int f = code >> 1;
        if (f >= nI) break;
        // side 1: c is the neighbour -> lower = A[c][own]; side 0: c is the owner -> upper = A[c][nei]
        const int slot = r0 + ((code & 1) ? neiOffs[f] : ownOffs[f]);
        const int other = (code & 1) ? owner[f] : neighbour[f];
        h = vsub(h, vscale(ld3(U, other), valuesV[3 * int64_t(slot)]));
    }
    h = vadd(h, ld3(rhsV, c));
    st3(HbyA, c, vscale(h, ra / vol));
}


// Row-walk variant for meshes whose CSR rows are laid out like the cell's stencil ([lower faces ascending | diag | upper faces
// ascending], fvk_mesh::rowsInStencilOrder): "off-diagonals subtracted in ascending face order" is then simply the row in
// entry order, so no stencil / owner / neighbour / offset array is read -- only the row itself. VS = doubles between matrix
// entries: 3 for the reference's Vec3 layout (component [0] is read), 1 for the compact layout (fvk_assemble_vc). Block-
// structured meshes with proven topology get the column indices of the regular rows from arithmetic (as k_spmv does).
struct RowsAffine
{
    int on, nx, ny, nz;
};
template <int VS>
__global__ void __launch_bounds__(256)
k_rAU_HbyA_rows(int nC, const int* __restrict__ rowOffs, const int* __restrict__ colIdxs, const uint8_t* __restrict__ diagOffs,
                const double* __restrict__ V, const double* __restrict__ values, const double* __restrict__ rhsV,
                const double* __restrict__ U, double* __restrict__ rAU, double* __restrict__ HbyA, RowsAffine aff)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int r0 = rowOffs[c], r1 = rowOffs[c + 1];
    const double vol = V[c];
    bool reg = false;
    int i = 0, j = 0, kz = 0;
    if (aff.on)
    {
        i = c % aff.nx; const int q = c / aff.nx; j = q % aff.ny; kz = q / aff.ny;
        reg = i > 0 && i < aff.nx - 1 && j > 0 && j < aff.ny - 1 && kz > 0 && kz < aff.nz - 1 && r1 - r0 == 7;
    }
    if (reg)
    {
        const int nxy = aff.nx * aff.ny;
        const int col[6] = {c - nxy, c - aff.nx, c - 1, c + 1, c + aff.nx, c + nxy};
        double a[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) a[k] = values[int64_t(VS) * (r0 + k)];
        const double ra = vol / a[3];
        rAU[c] = ra;
        if (!HbyA) return;
        Vec3d u[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) u[k] = ld3(U, col[k]);
        Vec3d h {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 6; ++k) h = vsub(h, vscale(u[k], a[k < 3 ? k : k + 1]));
        h = vadd(h, ld3(rhsV, c));
        st3(HbyA, c, vscale(h, ra / vol));
        return;
    }
    const int dpos = r0 + diagOffs[c];
    const double ra = vol / values[int64_t(VS) * dpos];
    rAU[c] = ra;
    if (!HbyA) return;
    Vec3d h {0.0, 0.0, 0.0};
    for (int k = r0; k < r1; ++k)
        if (k != dpos) h = vsub(h, vscale(ld3(U, colIdxs[k]), values[int64_t(VS) * k]));
    h = vadd(h, ld3(rhsV, c));
    st3(HbyA, c, vscale(h, ra / vol));
}

// flux (:215-267)
__global__ void __launch_bounds__(256)
k_flux(int nI, int nF, const int* __restrict__ owner, const int* __restrict__ neighbour, const double* __restrict__ w,
       const double* __restrict__ Sf, const double* __restrict__ bSf, const double* __restrict__ U,
       const double* __restrict__ Ub, double* __restrict__ outFace, double* __restrict__ outB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    if (f < nI)
    {
        const Vec3d uo = ld3(U, owner[f]), un = ld3(U, neighbour[f]);
        outFace[f] = vdot(ld3(Sf, f), vadd(vscale(vsub(uo, un), w[f]), un));
    }
    else
    {
        const int b = f - nI;
        const double v = vdot(ld3(bSf, b), ld3(Ub, b));
        outFace[f] = v;
        if (outB) outB[b] = v;
    }
}

// updateFaceVelocity (:131-197)
__global__ void __launch_bounds__(256)
k_update_face_velocity(int nI, int nF, const int* __restrict__ owner, const int* __restrict__ neighbour,
                       const int* __restrict__ faceCells, const int* __restrict__ rowOffs,
                       const uint8_t* __restrict__ ownOffs, const uint8_t* __restrict__ neiOffs,
                       const double* __restrict__ values, const double* __restrict__ bcMatrix,
                       const double* __restrict__ bcRhs, const double* __restrict__ p, const double* __restrict__ predPhi,
                       const double* __restrict__ predPhiB, double* __restrict__ phi, double* __restrict__ phiB)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    if (f < nI)
    {
        const int o = owner[f], n = neighbour[f];
        const double upper = values[rowOffs[n] + neiOffs[f]]; // the reference's naming
        const double lower = values[rowOffs[o] + ownOffs[f]];
        phi[f] = predPhi[f] - (upper * p[n] - lower * p[o]);
    }
    else
    {
        const int b = f - nI;
        const double bflux = bcRhs[b] - bcMatrix[b] * p[faceCells[b]];
        phi[f] = predPhi[f] - bflux;
        if (phiB) phiB[b] = predPhiB[b] - bflux;
    }
}

// updateVelocity (:199-213)
__global__ void __launch_bounds__(256)
k_update_velocity(int nC, const double* __restrict__ HbyA, const double* __restrict__ rAU, const double* __restrict__ gradP,
                  double* __restrict__ U)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    st3(U, c, vsub(ld3(HbyA, c), vscale(ld3(gradP, c), rAU[c])));
}

// PDESolver::SetReference (expression.hpp:86-112)
__global__ void k_set_reference(int refCell, double refValue, const int* __restrict__ rowOffs,
                                const uint8_t* __restrict__ diagOffs, double* __restrict__ values, double* __restrict__ rhs)
{
    const int idx = rowOffs[refCell] + diagOffs[refCell];
    const double d = values[idx];
    rhs[refCell] += d * refValue;
    values[idx] += d;
}

// diag (expression.hpp:181-199)
template <int NC>
__global__ void __launch_bounds__(256)
k_diag(int nC, const int* __restrict__ rowOffs, const uint8_t* __restrict__ diagOffs, const double* __restrict__ values,
       double* __restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nC) return;
    const int64_t s = rowOffs[c] + diagOffs[c];
#pragma unroll
    for (int k = 0; k < NC; ++k) out[NC * int64_t(c) + k] = values[NC * s + k];
}

// constrainHbyA (:14-36): dst_b = src_b on the selected patches
struct PatchMask
{
    int nPatches;
    int offsets[FVK_MAX_PATCHES + 1];
    int on[FVK_MAX_PATCHES];
};
template <int NC>
__global__ void __launch_bounds__(256)
k_copy_patches(PatchMask pm, int nB, const double* __restrict__ src, double* __restrict__ dst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nB) return;
    int p = 0;
    while (p + 1 < pm.nPatches && b >= pm.offsets[p + 1]) ++p;
    if (!pm.on[p]) return;
#pragma unroll
    for (int k = 0; k < NC; ++k) dst[NC * int64_t(b) + k] = src[NC * int64_t(b) + k];
}
} // namespace

#define GRID(n) ((n) + 255) / 256, 256, 0, fvk_cu(s)

static RowsAffine rows_affine(const fvk_mesh* m)
{
    const FvkBrickGeom& g = m->bp.geom;
    if (fvk_no_affine() || !g.affine || m->bp.nTiles == 0 || int64_t(g.dims[0]) * g.dims[1] * g.dims[2] != m->nOwned) return RowsAffine {0, 0, 0, 0};
    return RowsAffine {1, g.dims[0], g.dims[1], g.dims[2]};
}

extern "C" int fvk_rAU_HbyA_c(const fvk_mesh* m, const double* valuesCompact, const double* rhsV, const double* U, double* rAU,
                              double* HbyA, fvk_stream s)
{
    if (!m || !valuesCompact || !rAU || (HbyA && (!rhsV || !U))) return fvk_fail(FVK_EINVAL, "fvk_rAU_HbyA_c: null argument");
    if (!m->rowsInStencilOrder) return fvk_fail(FVK_EUNSUPPORTED, "fvk_rAU_HbyA_c: the mesh's CSR rows are not in stencil order");
    k_rAU_HbyA_rows<1><<<GRID(m->nOwned)>>>(m->nOwned, m->rowOffs, m->colIdxs, m->diagOffset, m->V, valuesCompact, rhsV, U, rAU, HbyA, rows_affine(m));
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_rAU_HbyA(const fvk_mesh* m, const double* valuesV, const double* rhsV, const double* U, double* rAU,
                            double* HbyA, fvk_stream s)
{
    if (!m || !valuesV || !rAU || (HbyA && (!rhsV || !U))) return fvk_fail(FVK_EINVAL, "fvk_rAU_HbyA: null argument");
    if (m->rowsInStencilOrder && fvk_variant() == 0)
    {
        k_rAU_HbyA_rows<3><<<GRID(m->nOwned)>>>(m->nOwned, m->rowOffs, m->colIdxs, m->diagOffset, m->V, valuesV, rhsV, U, rAU, HbyA, rows_affine(m));
        FVK_LAUNCH_CHECK();
        return FVK_OK;
    }
    k_rAU_HbyA<<<GRID(m->nOwned)>>>(m->nOwned, m->nInternalFaces, m->stencilSeg, m->gatherEnt, m->owner, m->neighbour, m->rowOffs,
                                    m->diagOffset, m->ownerOffset, m->neighbourOffset, m->V, valuesV, rhsV, U, rAU, HbyA);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_flux(const fvk_mesh* m, const double* U, const double* Ub, double* outFace, double* outB, fvk_stream s)
{
    if (!m || !U || !outFace || (m->nBoundaryFaces && (!Ub || !m->bSf))) return fvk_fail(FVK_EINVAL, "fvk_flux: null argument");
    const int nF = m->nInternalFaces + m->nBoundaryFaces;
    if (nF == 0) return FVK_OK;
    k_flux<<<GRID(nF)>>>(m->nInternalFaces, nF, m->owner, m->neighbour, m->weights, m->Sf, m->bSf, U, Ub, outFace, outB);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_update_face_velocity(const fvk_mesh* m, const double* values, const double* bcMatrix, const double* bcRhs,
                                        const double* p, const double* predPhi, const double* predPhiB, double* phi,
                                        double* phiB, fvk_stream s)
{
    if (!m || !values || !p || !predPhi || !phi || (m->nBoundaryFaces && (!bcMatrix || !bcRhs || (phiB && !predPhiB))))
        return fvk_fail(FVK_EINVAL, "fvk_update_face_velocity: null argument");
    const int nF = m->nInternalFaces + m->nBoundaryFaces;
    if (nF == 0) return FVK_OK;
    k_update_face_velocity<<<GRID(nF)>>>(m->nInternalFaces, nF, m->owner, m->neighbour, m->faceCells, m->rowOffs, m->ownerOffset,
                                         m->neighbourOffset, values, bcMatrix, bcRhs, p, predPhi, predPhiB, phi, phiB);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_update_velocity(const fvk_mesh* m, const double* HbyA, const double* rAU, const double* gradP, double* U,
                                   fvk_stream s)
{
    if (!m || !HbyA || !rAU || !gradP || !U) return fvk_fail(FVK_EINVAL, "fvk_update_velocity: null argument");
    k_update_velocity<<<GRID(m->nOwned)>>>(m->nOwned, HbyA, rAU, gradP, U);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_set_reference(const fvk_mesh* m, int32_t refCell, double refValue, double* values, double* rhs, fvk_stream s)
{
    if (!m || !values || !rhs || refCell < 0 || refCell >= m->nOwned) return fvk_fail(FVK_EINVAL, "fvk_set_reference: bad argument");
    k_set_reference<<<1, 1, 0, fvk_cu(s)>>>(refCell, refValue, m->rowOffs, m->diagOffset, values, rhs);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_diag(const fvk_mesh* m, int ncomp, const double* values, double* out, fvk_stream s)
{
    if (!m || !values || !out || (ncomp != 1 && ncomp != 3)) return fvk_fail(FVK_EINVAL, "fvk_diag: bad argument");
    if (ncomp == 1) k_diag<1><<<GRID(m->nOwned)>>>(m->nOwned, m->rowOffs, m->diagOffset, values, out);
    else k_diag<3><<<GRID(m->nOwned)>>>(m->nOwned, m->rowOffs, m->diagOffset, values, out);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}

extern "C" int fvk_copy_patches(const fvk_mesh* m, int ncomp, const int32_t* patchMask_h, const double* srcB, double* dstB,
                                fvk_stream s)
{
    if (!m || !patchMask_h || (ncomp != 1 && ncomp != 3) || (m->nBoundaryFaces && (!srcB || !dstB)))
        return fvk_fail(FVK_EINVAL, "fvk_copy_patches: bad argument");
    if (m->nBoundaryFaces == 0) return FVK_OK;
    PatchMask pm;
    pm.nPatches = m->nPatches;
    for (int p = 0; p <= m->nPatches; ++p) pm.offsets[p] = m->patchOffsets[p];
    for (int p = 0; p < m->nPatches; ++p) pm.on[p] = patchMask_h[p] != 0;
    if (ncomp == 1) k_copy_patches<1><<<GRID(m->nBoundaryFaces)>>>(pm, m->nBoundaryFaces, srcB, dstB);
    else k_copy_patches<3><<<GRID(m->nBoundaryFaces)>>>(pm, m->nBoundaryFaces, srcB, dstB);
    FVK_LAUNCH_CHECK();
    return FVK_OK;
}
