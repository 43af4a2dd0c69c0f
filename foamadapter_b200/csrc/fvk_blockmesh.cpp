// Host-side synthetic block-hex mesh in OpenFOAM blockMesh ordering.
//
// Stand-in for `blockMesh` + FoamAdapter::readOpenFOAMMesh (reference
// src/datastructures/meshAdapter.cpp:59-136), which need OpenFOAM. Topology/ordering follows the
// polyMesh fixtures the reference commits (test/setup_*/constant/polyMesh);
// geometry follows OpenFOAM's primitiveMesh (triangle-fan face centres/areas, pyramid cell
// centres/volumes) evaluated on points+faces, so it also holds for non-uniform point sets.
#include "fvk_internal.hpp"

#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include <utility>
#include <vector>

namespace
{
template <class T> using RawVec = FvkRawVec<T>; // (fvk_internal.hpp: no serial zero-fill; every element below is written by the parallel loops)

struct BlockMeshStore
{
    fvk_mesh_desc desc {};
    RawVec<double> points, V, C, Sf, Cf, magSf;
    RawVec<int32_t> owner, neighbour, faceCells;
    std::vector<int32_t> patchOffsets;
    RawVec<double> bCf, bCn, bSf, bMagSf, bNf, bDelta, bWeights, bDeltaCoeffs;
    // poly (all faces incl. empty patches)
    RawVec<int32_t> polyFaces, polyOwner;
    int32_t nPolyFaces = 0;
};

inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// quad face centre and area vector, OpenFOAM primitiveMeshFaceCentresAndAreas.C
inline void quadGeometry(const double* p, const int32_t* f, double* cf, double* sf)
{
    double fc[3];
    for (int d = 0; d < 3; ++d)
    {
        double s = p[3 * f[0] + d];
        for (int k = 1; k < 4; ++k) s += p[3 * f[k] + d];
        fc[d] = s / 4;
    }
    double sumN[3] = {0, 0, 0}, sumAc[3] = {0, 0, 0}, sumA = 0.0;
    for (int k = 0; k < 4; ++k)
    {
        const double* cur = p + 3 * f[k];
        const double* nxt = p + 3 * f[(k + 1) & 3];
        double e1[3], e2[3], n[3], c[3];
        for (int d = 0; d < 3; ++d)
        {
            c[d] = cur[d] + nxt[d] + fc[d];
            e1[d] = nxt[d] - cur[d];
            e2[d] = fc[d] - cur[d];
        }
        cross3(e1, e2, n);
        const double a = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        sumA += a;
        for (int d = 0; d < 3; ++d)
        {
            sumN[d] += n[d];
            sumAc[d] += a * c[d];
        }
    }
    if (sumA < 1e-150)
    {
        for (int d = 0; d < 3; ++d) { cf[d] = fc[d]; sf[d] = 0.0; }
    }
    else
    {
        for (int d = 0; d < 3; ++d)
        {
            cf[d] = (1.0 / 3.0) * sumAc[d] / sumA;
            sf[d] = 0.5 * sumN[d];
        }
    }
}

} // namespace

extern "C" int fvk_blockmesh_create(int32_t nx, int32_t ny, int32_t nz, double lx, double ly,
                                    double lz, int32_t nPatches, const int32_t* patchNSides,
                                    const int32_t* patchSides, const int32_t* patchIsEmpty,
                                    int32_t withPoints, fvk_mesh_desc** out)
{
    if (!out || nx < 1 || ny < 1 || nz < 1 || nPatches < 0 || (nPatches && (!patchNSides || !patchSides)))
        return fvk_fail(FVK_EINVAL, "fvk_blockmesh_create: bad arguments");
    const int64_t nC64 = int64_t(nx) * ny * nz;
    const int64_t nI64 = int64_t(nx - 1) * ny * nz + int64_t(nx) * (ny - 1) * nz + int64_t(nx) * ny * (nz - 1);
    const int64_t sideSize[6] = {int64_t(ny) * nz, int64_t(ny) * nz, int64_t(nx) * nz,
                                 int64_t(nx) * nz, int64_t(nx) * ny, int64_t(nx) * ny};
    int64_t nPolyB = 0, nB64 = 0;
    {
        int s = 0;
        bool used[6] = {false, false, false, false, false, false};
        for (int p = 0; p < nPatches; ++p)
            for (int q = 0; q < patchNSides[p]; ++q, ++s)
            {
                const int side = patchSides[s];
                if (side < 0 || side > 5 || used[side])
                    return fvk_fail(FVK_EINVAL, "fvk_blockmesh_create: bad/duplicate side");
                used[side] = true;
                nPolyB += sideSize[side];
                if (!(patchIsEmpty && patchIsEmpty[p])) nB64 += sideSize[side];
            }
        for (int side = 0; side < 6; ++side)
            if (!used[side]) return fvk_fail(FVK_EINVAL, "fvk_blockmesh_create: every block side needs a patch");
    }
    if (nI64 + nPolyB >= (int64_t(1) << 30) || nC64 >= (int64_t(1) << 30))
        return fvk_fail(FVK_EUNSUPPORTED, "fvk_blockmesh_create: mesh too large for int32 labels");

    BlockMeshStore* st = new (std::nothrow) BlockMeshStore;
    if (!st) return fvk_fail(FVK_ENOMEM, "fvk_blockmesh_create: out of memory");
    try
    {
        const int32_t nC = int32_t(nC64), nI = int32_t(nI64), nB = int32_t(nB64);
        const int32_t nPoly = int32_t(nI64 + nPolyB);
        const int32_t nF = nI + nB;
        const int64_t npx = nx + 1, npy = ny + 1, npz = nz + 1;
        const int64_t nP = npx * npy * npz;
        auto vtx = [=](int64_t i, int64_t j, int64_t k) { return int32_t(i + npx * (j + npy * k)); };
        auto cell = [=](int64_t i, int64_t j, int64_t k) { return int32_t(i + int64_t(nx) * (j + int64_t(ny) * k)); };

        st->points.resize(3 * nP);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < npz; ++k)
            for (int64_t j = 0; j < npy; ++j)
                for (int64_t i = 0; i < npx; ++i)
                {
                    double* p = &st->points[3 * (i + npx * (j + npy * k))];
                    p[0] = lx * (double(i) / nx);
                    p[1] = ly * (double(j) / ny);
                    p[2] = lz * (double(k) / nz);
                }

        // ---- faces ------------------------------------------------------------------------
        // Everything below is CELL-CENTRIC and parallel over cells (or over a side's faces): the block's topology is closed
        // form, so a cell finds its own faces and its lower neighbours' faces by arithmetic. The per-cell sums keep the order
        // of OpenFOAM's face loops (all faces the cell owns in ascending poly id, then the faces it is neighbour of in
        // ascending id), so the geometry is bit-identical to the serial face-loop formulation (and to oracle/blockmesh.cpp).
        st->neighbour.resize(nI);
        st->nPolyFaces = withPoints ? nPoly : 0;
        if (withPoints)
        {
            st->polyFaces.resize(4 * int64_t(nPoly));
            st->polyOwner.resize(nPoly);
        }
        // internal faces: owner ascending, then +x, +y, +z. First face id of each cell by scan.
        RawVec<int32_t> ownStart(size_t(nC) + 1);
        {
            // faces owned by cell (i,j,k) = (i<nx-1)+(j<ny-1)+(k<nz-1); closed form per row
            int64_t run = 0;
            for (int64_t k = 0; k < nz; ++k)
                for (int64_t j = 0; j < ny; ++j)
                {
                    const int base = (j < ny - 1) + (k < nz - 1);
                    int32_t* o = &ownStart[cell(0, j, k)];
                    for (int64_t i = 0; i < nx; ++i)
                    {
                        o[i] = int32_t(run);
                        run += base + (i < nx - 1);
                    }
                }
            ownStart[nC] = int32_t(run);
        }
        // boundary faces: patches in dictionary order, sides in listed order (blockMesh block::createBoundary loop nests);
        // sideStart = poly id of the side's first face, -> its index in the kept (non-empty-type) boundary list or -1
        int64_t sideStart[6] = {0, 0, 0, 0, 0, 0}, sideKept[6] = {-1, -1, -1, -1, -1, -1};
        st->patchOffsets.assign(1, 0);
        {
            int64_t f = nI, kept = 0;
            int s = 0;
            for (int p = 0; p < nPatches; ++p)
            {
                const bool isEmpty = patchIsEmpty && patchIsEmpty[p];
                int32_t patchSize = 0;
                for (int q = 0; q < patchNSides[p]; ++q, ++s)
                {
                    const int side = patchSides[s];
                    sideStart[side] = f;
                    f += sideSize[side];
                    if (!isEmpty)
                    {
                        sideKept[side] = kept;
                        kept += sideSize[side];
                        patchSize += int32_t(sideSize[side]);
                    }
                }
                if (!isEmpty) st->patchOffsets.push_back(st->patchOffsets.back() + patchSize);
            }
        }
        const int32_t nKeptPatches = int32_t(st->patchOffsets.size()) - 1;
        // index of a cell's face inside a side's loop nest, and the inverse
        auto sideIndex = [=](int side, int64_t i, int64_t j, int64_t k) -> int64_t {
            return side < 2 ? k * ny + j : (side < 4 ? i * nz + k : i * ny + j);
        };
        const int64_t nPolyBnd = nPolyB;
        RawVec<double> bndCf(3 * size_t(nPolyBnd)), bndSf(3 * size_t(nPolyBnd)); // geometry of ALL boundary poly faces (incl. empty patches)
        st->owner.resize(nF);
        st->Sf.resize(3 * size_t(nF));
        st->Cf.resize(3 * size_t(nF));
        st->magSf.resize(nF);
        const double* pts = st->points.data();
        for (int side = 0; side < 6; ++side)
        {
            const int64_t n = sideSize[side], f0 = sideStart[side];
#pragma omp parallel for schedule(static)
            for (int64_t idx = 0; idx < n; ++idx)
            {
                int64_t i, j, k;
                int32_t q[4];
                switch (side)
                {
                    case 0: k = idx / ny; j = idx % ny; i = 0;
                        q[0] = vtx(0, j, k); q[1] = vtx(0, j, k + 1); q[2] = vtx(0, j + 1, k + 1); q[3] = vtx(0, j + 1, k); break;
                    case 1: k = idx / ny; j = idx % ny; i = nx - 1;
                        q[0] = vtx(nx, j, k); q[1] = vtx(nx, j + 1, k); q[2] = vtx(nx, j + 1, k + 1); q[3] = vtx(nx, j, k + 1); break;
                    case 2: i = idx / nz; k = idx % nz; j = 0;
                        q[0] = vtx(i, 0, k); q[1] = vtx(i + 1, 0, k); q[2] = vtx(i + 1, 0, k + 1); q[3] = vtx(i, 0, k + 1); break;
                    case 3: i = idx / nz; k = idx % nz; j = ny - 1;
                        q[0] = vtx(i, ny, k); q[1] = vtx(i, ny, k + 1); q[2] = vtx(i + 1, ny, k + 1); q[3] = vtx(i + 1, ny, k); break;
                    case 4: i = idx / ny; j = idx % ny; k = 0;
                        q[0] = vtx(i, j, 0); q[1] = vtx(i, j + 1, 0); q[2] = vtx(i + 1, j + 1, 0); q[3] = vtx(i + 1, j, 0); break;
                    default: i = idx / ny; j = idx % ny; k = nz - 1;
                        q[0] = vtx(i, j, nz); q[1] = vtx(i + 1, j, nz); q[2] = vtx(i + 1, j + 1, nz); q[3] = vtx(i, j + 1, nz); break;
                }
                const int64_t f = f0 + idx, fb = f - nI;
                quadGeometry(pts, q, &bndCf[3 * fb], &bndSf[3 * fb]);
                if (withPoints)
                {
                    for (int d = 0; d < 4; ++d) st->polyFaces[4 * f + d] = q[d];
                    st->polyOwner[f] = cell(i, j, k);
                }
                if (sideKept[side] >= 0)
                { // NeoN view: the kept boundary faces follow the internal ones
                    const int64_t g = nI + sideKept[side] + idx;
                    st->owner[g] = cell(i, j, k);
                    for (int d = 0; d < 3; ++d) { st->Cf[3 * g + d] = bndCf[3 * fb + d]; st->Sf[3 * g + d] = bndSf[3 * fb + d]; }
                }
            }
        }
        // internal faces: labels and geometry straight into the NeoN arrays (poly id == NeoN id for them)
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < nz; ++k)
            for (int64_t j = 0; j < ny; ++j)
                for (int64_t i = 0; i < nx; ++i)
                {
                    const int32_t c = cell(i, j, k);
                    int64_t f = ownStart[c];
                    int32_t q[4];
                    auto emit = [&](int32_t nei) {
                        quadGeometry(pts, q, &st->Cf[3 * f], &st->Sf[3 * f]);
                        st->owner[f] = c; st->neighbour[f] = nei;
                        if (withPoints)
                        {
                            for (int d = 0; d < 4; ++d) st->polyFaces[4 * f + d] = q[d];
                            st->polyOwner[f] = c;
                        }
                        ++f;
                    };
                    if (i < nx - 1)
                    {
                        q[0] = vtx(i + 1, j, k); q[1] = vtx(i + 1, j + 1, k); q[2] = vtx(i + 1, j + 1, k + 1); q[3] = vtx(i + 1, j, k + 1);
                        emit(cell(i + 1, j, k));
                    }
                    if (j < ny - 1)
                    {
                        q[0] = vtx(i, j + 1, k); q[1] = vtx(i, j + 1, k + 1); q[2] = vtx(i + 1, j + 1, k + 1); q[3] = vtx(i + 1, j + 1, k);
                        emit(cell(i, j + 1, k));
                    }
                    if (k < nz - 1)
                    {
                        q[0] = vtx(i, j, k + 1); q[1] = vtx(i + 1, j, k + 1); q[2] = vtx(i + 1, j + 1, k + 1); q[3] = vtx(i, j + 1, k + 1);
                        emit(cell(i, j, k + 1));
                    }
                }

        // ---- cell centres and volumes (primitiveMeshCellCentresAndVols.C), per cell ---------------------------------
        // a cell's faces: owned = its internal faces [ownStart[c], ownStart[c+1]) then its boundary faces in ascending poly id;
        // neighbour-of = the z, y, x faces of the cells below / in front / left (ascending ids)
        st->C.resize(3 * size_t(nC));
        st->V.resize(nC);
        const int64_t nxy = int64_t(nx) * ny;
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < nz; ++k)
            for (int64_t j = 0; j < ny; ++j)
                for (int64_t i = 0; i < nx; ++i)
                {
                    const int64_t c = cell(i, j, k);
                    const double* own[9];
                    const double* ownS[9];
                    int nOwn = 0;
                    for (int64_t f = ownStart[c]; f < ownStart[c + 1]; ++f) { own[nOwn] = &st->Cf[3 * f]; ownS[nOwn] = &st->Sf[3 * f]; ++nOwn; }
                    int64_t bf[6];
                    int nb = 0;
                    if (i == 0) bf[nb++] = sideStart[0] + sideIndex(0, i, j, k);
                    if (i == nx - 1) bf[nb++] = sideStart[1] + sideIndex(1, i, j, k);
                    if (j == 0) bf[nb++] = sideStart[2] + sideIndex(2, i, j, k);
                    if (j == ny - 1) bf[nb++] = sideStart[3] + sideIndex(3, i, j, k);
                    if (k == 0) bf[nb++] = sideStart[4] + sideIndex(4, i, j, k);
                    if (k == nz - 1) bf[nb++] = sideStart[5] + sideIndex(5, i, j, k);
                    for (int x = 1; x < nb; ++x) // ascending poly id (at most six)
                        for (int y = x; y > 0 && bf[y] < bf[y - 1]; --y) { const int64_t t = bf[y]; bf[y] = bf[y - 1]; bf[y - 1] = t; }
                    for (int x = 0; x < nb; ++x) { own[nOwn] = &bndCf[3 * (bf[x] - nI)]; ownS[nOwn] = &bndSf[3 * (bf[x] - nI)]; ++nOwn; }
                    const double* low[3];
                    const double* lowS[3];
                    int nLow = 0;
                    if (k > 0) { const int64_t f = ownStart[c - nxy] + (i < nx - 1) + (j < ny - 1); low[nLow] = &st->Cf[3 * f]; lowS[nLow] = &st->Sf[3 * f]; ++nLow; }
                    if (j > 0) { const int64_t f = ownStart[c - nx] + (i < nx - 1); low[nLow] = &st->Cf[3 * f]; lowS[nLow] = &st->Sf[3 * f]; ++nLow; }
                    if (i > 0) { const int64_t f = ownStart[c - 1]; low[nLow] = &st->Cf[3 * f]; lowS[nLow] = &st->Sf[3 * f]; ++nLow; }
                    // cell centre estimate = mean of face centres (own faces then nei faces, OpenFOAM order)
                    double est[3] = {0.0, 0.0, 0.0};
                    for (int x = 0; x < nOwn; ++x)
                        for (int d = 0; d < 3; ++d) est[d] += own[x][d];
                    for (int x = 0; x < nLow; ++x)
                        for (int d = 0; d < 3; ++d) est[d] += low[x][d];
                    const int cnt = nOwn + nLow;
                    for (int d = 0; d < 3; ++d) est[d] /= cnt;
                    double Cc[3] = {0.0, 0.0, 0.0}, Vc = 0.0;
                    for (int x = 0; x < nOwn; ++x)
                    {
                        double pyr3 = 0.0;
                        for (int d = 0; d < 3; ++d) pyr3 += ownS[x][d] * (own[x][d] - est[d]);
                        for (int d = 0; d < 3; ++d) Cc[d] += pyr3 * ((3.0 / 4.0) * own[x][d] + (1.0 / 4.0) * est[d]);
                        Vc += pyr3;
                    }
                    for (int x = 0; x < nLow; ++x)
                    {
                        double pyr3 = 0.0;
                        for (int d = 0; d < 3; ++d) pyr3 += lowS[x][d] * (est[d] - low[x][d]);
                        for (int d = 0; d < 3; ++d) Cc[d] += pyr3 * ((3.0 / 4.0) * low[x][d] + (1.0 / 4.0) * est[d]);
                        Vc += pyr3;
                    }
                    for (int d = 0; d < 3; ++d) st->C[3 * c + d] = Cc[d] / Vc;
                    st->V[c] = Vc * (1.0 / 3.0);
                }
        RawVec<double>().swap(bndCf);
        RawVec<double>().swap(bndSf);
        RawVec<int32_t>().swap(ownStart);
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < nF; ++f)
        {
            const double* s = &st->Sf[3 * f];
            st->magSf[f] = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        }
        // BoundaryMesh arrays (fvPatch: Cf, Cn, Sf, magSf, nf = Sf/magSf, delta = Cf - Cn,
        // weights = 1 and deltaCoeffs = 1/|delta| on non-coupled patches)
        st->faceCells.resize(nB);
        st->bCf.resize(3 * size_t(nB)); st->bCn.resize(3 * size_t(nB)); st->bSf.resize(3 * size_t(nB));
        st->bNf.resize(3 * size_t(nB)); st->bDelta.resize(3 * size_t(nB));
        st->bMagSf.resize(nB); st->bWeights.resize(nB); st->bDeltaCoeffs.resize(nB);
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < nB; ++b)
        {
            const int64_t f = nI + b;
            const int32_t o = st->owner[f];
            st->faceCells[b] = o;
            double d2 = 0.0;
            for (int d = 0; d < 3; ++d)
            {
                st->bCf[3 * b + d] = st->Cf[3 * f + d];
                st->bCn[3 * b + d] = st->C[3 * size_t(o) + d];
                st->bSf[3 * b + d] = st->Sf[3 * f + d];
                st->bNf[3 * b + d] = st->Sf[3 * f + d] / st->magSf[f];
                const double dl = st->Cf[3 * f + d] - st->C[3 * size_t(o) + d];
                st->bDelta[3 * b + d] = dl;
                d2 += dl * dl;
            }
            st->bMagSf[b] = st->magSf[f];
            st->bWeights[b] = 1.0;
            st->bDeltaCoeffs[b] = 1.0 / std::sqrt(d2);
        }
        if (!withPoints) RawVec<double>().swap(st->points);

        fvk_mesh_desc& d = st->desc;
        d.nCells = nC; d.nInternalFaces = nI; d.nBoundaryFaces = nB; d.nPatches = nKeptPatches;
        d.nPoints = withPoints ? int32_t(nP) : 0;
        d.points = withPoints ? st->points.data() : nullptr;
        d.cellVolumes = st->V.data(); d.cellCentres = st->C.data();
        d.faceAreas = st->Sf.data(); d.faceCentres = st->Cf.data(); d.magFaceAreas = st->magSf.data();
        d.faceOwner = st->owner.data(); d.faceNeighbour = st->neighbour.data();
        d.faceCells = st->faceCells.data();
        d.bCf = st->bCf.data(); d.bCn = st->bCn.data(); d.bSf = st->bSf.data();
        d.bMagSf = st->bMagSf.data(); d.bNf = st->bNf.data(); d.bDelta = st->bDelta.data();
        d.bWeights = st->bWeights.data(); d.bDeltaCoeffs = st->bDeltaCoeffs.data();
        d.patchOffsets = st->patchOffsets.data();
    }
    catch (const std::bad_alloc&)
    {
        delete st;
        return fvk_fail(FVK_ENOMEM, "fvk_blockmesh_create: out of memory");
    }
    static_assert(offsetof(BlockMeshStore, desc) == 0, "desc must be first");
    *out = &st->desc;
    return FVK_OK;
}

extern "C" int fvk_blockmesh_destroy(fvk_mesh_desc* desc)
{
    if (desc) delete reinterpret_cast<BlockMeshStore*>(desc);
    return FVK_OK;
}

extern "C" int fvk_blockmesh_poly(const fvk_mesh_desc* desc, int32_t* nPolyFaces,
                                  const int32_t** facePoints, const int32_t** polyOwner)
{
    if (!desc) return fvk_fail(FVK_EINVAL, "fvk_blockmesh_poly: null mesh");
    const BlockMeshStore* st = reinterpret_cast<const BlockMeshStore*>(desc);
    if (st->nPolyFaces == 0) return fvk_fail(FVK_EINVAL, "fvk_blockmesh_poly: created without points");
    if (nPolyFaces) *nPolyFaces = st->nPolyFaces;
    if (facePoints) *facePoints = st->polyFaces.data();
    if (polyOwner) *polyOwner = st->polyOwner.data();
    return FVK_OK;
}
