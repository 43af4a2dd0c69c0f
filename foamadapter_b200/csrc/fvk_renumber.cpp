// Host-side cell renumbering for coalescing (BASELINE.json north_star: "a renumbered (RCM / space-filling) cell order").
//
// The reference has no renumbering of its own: OpenFOAM cases are renumbered with the `renumberMesh` utility before the
// solver starts, and FoamAdapter reads whatever order the polyMesh has (src/datastructures/meshAdapter.cpp:59-136). This is
// that utility for mesh descriptions: a permutation of the cells (reverse Cuthill-McKee on the cell-cell graph, or the Morton
// order of the cell centres), then the internal faces re-sorted into OpenFOAM's upper-triangular order (owner < neighbour,
// by owner then neighbour; a face whose owner and neighbour swap gets -Sf). Boundary faces keep their patch order. The
// renumbering is applied at MESH level, like OpenFOAM does: fields are caller-owned arrays, so a permutation hidden inside
// the kernels would turn every field access into a scattered load -- instead the application lives in the new order and the
// returned maps take fields across. On the renumbered mesh every kernel is bit-identical to the reference's Serial executor
// on the same mesh (the parity tests run on renumbered meshes too).
#include "fvk_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <numeric>
#include <queue>
#include <vector>

namespace
{
struct RenumStore
{
    fvk_mesh_desc desc {};
    std::vector<double> points, V, C, Sf, Cf, magSf, bCf, bCn, bSf, bMagSf, bNf, bDelta, bWeights, bDeltaCoeffs;
    std::vector<int32_t> owner, neighbour, faceCells, patchOffsets;
    std::vector<int32_t> faceOldToNew;
    std::vector<uint8_t> flipped;
    uint32_t magic = 0x52454e55u; // 'RENU'
};

inline uint64_t spread21(uint64_t v)
{ // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

void morton_order(const fvk_mesh_desc* d, int32_t* oldToNew)
{
    const int32_t nC = d->nCells;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int32_t c = 0; c < nC; ++c)
        for (int k = 0; k < 3; ++k)
        {
            lo[k] = std::min(lo[k], d->cellCentres[3 * size_t(c) + k]);
            hi[k] = std::max(hi[k], d->cellCentres[3 * size_t(c) + k]);
        }
    // one quantum = the smallest extent that still separates neighbouring cells: scale all axes alike so the curve's cells are cubes
    double ext = 0.0;
    for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
    const double scale = ext > 0 ? double((1u << 21) - 1) / ext : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> key;
    key.resize(size_t(nC));
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < nC; ++c)
    {
        uint64_t q[3];
        for (int k = 0; k < 3; ++k) q[k] = uint64_t(std::llround((d->cellCentres[3 * size_t(c) + k] - lo[k]) * scale));
        key[c] = {spread21(q[0]) | spread21(q[1]) << 1 | spread21(q[2]) << 2, c};
    }
    std::sort(key.begin(), key.end());
    for (int32_t i = 0; i < nC; ++i) oldToNew[key[i].second] = i;
}

void rcm_order(const fvk_mesh_desc* d, int32_t* oldToNew)
{
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    std::vector<int32_t> off(size_t(nC) + 1, 0), adj(2 * size_t(nI));
    for (int32_t f = 0; f < nI; ++f) { ++off[size_t(d->faceOwner[f]) + 1]; ++off[size_t(d->faceNeighbour[f]) + 1]; }
    for (int32_t c = 0; c < nC; ++c) off[size_t(c) + 1] += off[c];
    std::vector<int32_t> pos(off.begin(), off.end() - 1);
    for (int32_t f = 0; f < nI; ++f)
    {
        adj[pos[d->faceOwner[f]]++] = d->faceNeighbour[f];
        adj[pos[d->faceNeighbour[f]]++] = d->faceOwner[f];
    }
    auto degree = [&](int32_t c) { return off[size_t(c) + 1] - off[c]; };
    std::vector<int32_t> order;
    order.reserve(nC);
    std::vector<uint8_t> seen(nC, 0);
    // components in ascending order of their minimum-degree cell (ties: lowest id): Cuthill-McKee breadth-first search,
    // neighbours appended by ascending degree (ties: lowest id)
    std::vector<int32_t> byDegree(nC);
    std::iota(byDegree.begin(), byDegree.end(), 0);
    std::stable_sort(byDegree.begin(), byDegree.end(), [&](int32_t a, int32_t b) { return degree(a) < degree(b); });
    std::vector<int32_t> nb;
    for (int32_t start : byDegree)
    {
        if (seen[start]) continue;
        size_t head = order.size();
        order.push_back(start);
        seen[start] = 1;
        while (head < order.size())
        {
            const int32_t c = order[head++];
            nb.clear();
            for (int32_t k = off[c]; k < off[size_t(c) + 1]; ++k)
                if (!seen[adj[k]]) { seen[adj[k]] = 1; nb.push_back(adj[k]); }
            std::sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return degree(a) != degree(b) ? degree(a) < degree(b) : a < b; });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    for (int32_t i = 0; i < nC; ++i) oldToNew[order[size_t(nC) - 1 - i]] = i; // reversed
}
} // namespace

extern "C" int fvk_renumber_order(const fvk_mesh_desc* d, int method, int32_t* cellOldToNew)
{
    if (!d || !cellOldToNew || d->nCells <= 0 || !d->faceOwner || (d->nInternalFaces && !d->faceNeighbour) || !d->cellCentres)
        return fvk_fail(FVK_EINVAL, "fvk_renumber_order: bad argument");
    if (d->nOwnedCells > 0 && d->nOwnedCells != d->nCells) return fvk_fail(FVK_EUNSUPPORTED, "fvk_renumber_order: renumber before decomposing");
    if (method == FVK_RENUMBER_RCM) rcm_order(d, cellOldToNew);
    else if (method == FVK_RENUMBER_MORTON) morton_order(d, cellOldToNew);
    else return fvk_fail(FVK_EINVAL, "fvk_renumber_order: unknown method %d", method);
    return FVK_OK;
}

extern "C" int fvk_renumber_apply(const fvk_mesh_desc* d, const int32_t* map, fvk_mesh_desc** out)
{
    if (!d || !map || !out) return fvk_fail(FVK_EINVAL, "fvk_renumber_apply: null argument");
    if (d->nOwnedCells > 0 && d->nOwnedCells != d->nCells) return fvk_fail(FVK_EUNSUPPORTED, "fvk_renumber_apply: renumber before decomposing");
    *out = nullptr;
    const int32_t nC = d->nCells, nI = d->nInternalFaces, nB = d->nBoundaryFaces;
    const int64_t nF = int64_t(nI) + nB;
    {
        std::vector<uint8_t> hit(size_t(nC), 0);
        for (int32_t c = 0; c < nC; ++c)
        {
            if (map[c] < 0 || map[c] >= nC || hit[map[c]]) return fvk_fail(FVK_EINVAL, "fvk_renumber_apply: cellOldToNew is not a permutation (cell %d)", c);
            hit[map[c]] = 1;
        }
    }
    RenumStore* st = new (std::nothrow) RenumStore;
    if (!st) return fvk_fail(FVK_ENOMEM, "fvk_renumber_apply: out of memory");
    try
    {
        st->V.resize(nC); st->C.resize(3 * size_t(nC));
#pragma omp parallel for schedule(static)
        for (int32_t c = 0; c < nC; ++c)
        {
            const size_t n = size_t(map[c]);
            st->V[n] = d->cellVolumes[c];
            for (int k = 0; k < 3; ++k) st->C[3 * n + k] = d->cellCentres[3 * size_t(c) + k];
        }
        // internal faces: new (owner, neighbour), flipped where they swap, sorted into upper-triangular order
        std::vector<int32_t> idx(nI);
        std::iota(idx.begin(), idx.end(), 0);
        std::vector<int32_t> lo(nI), hi(nI);
        st->flipped.assign(size_t(nF), 0);
        for (int32_t f = 0; f < nI; ++f)
        {
            const int32_t o = map[d->faceOwner[f]], n = map[d->faceNeighbour[f]];
            lo[f] = std::min(o, n); hi[f] = std::max(o, n);
        }
        std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return lo[a] != lo[b] ? lo[a] < lo[b] : hi[a] < hi[b]; });
        st->faceOldToNew.resize(size_t(nF));
        st->owner.resize(size_t(nF)); st->neighbour.resize(nI);
        st->Sf.resize(3 * size_t(nF)); st->Cf.resize(3 * size_t(nF)); st->magSf.resize(size_t(nF));
#pragma omp parallel for schedule(static)
        for (int32_t i = 0; i < nI; ++i)
        {
            const int32_t f = idx[i];
            const bool flip = map[d->faceOwner[f]] > map[d->faceNeighbour[f]];
            st->faceOldToNew[f] = i;
            st->flipped[f] = flip;
            st->owner[i] = lo[f]; st->neighbour[i] = hi[f];
            for (int k = 0; k < 3; ++k)
            {
                const double s = d->faceAreas[3 * size_t(f) + k];
                st->Sf[3 * size_t(i) + k] = flip ? -s : s;
                st->Cf[3 * size_t(i) + k] = d->faceCentres[3 * size_t(f) + k];
            }
            st->magSf[i] = d->magFaceAreas[f];
        }
        st->faceCells.resize(nB);
        for (int32_t b = 0; b < nB; ++b)
        {
            const size_t f = size_t(nI) + b;
            st->faceOldToNew[f] = int32_t(f);
            st->faceCells[b] = map[d->faceCells[b]];
            st->owner[f] = st->faceCells[b];
            for (int k = 0; k < 3; ++k) { st->Sf[3 * f + k] = d->faceAreas[3 * f + k]; st->Cf[3 * f + k] = d->faceCentres[3 * f + k]; }
            st->magSf[f] = d->magFaceAreas[f];
        }
        auto copyB = [&](std::vector<double>& dst, const double* src, size_t n) { if (src) dst.assign(src, src + n); };
        copyB(st->bCf, d->bCf, 3 * size_t(nB)); copyB(st->bCn, d->bCn, 3 * size_t(nB)); copyB(st->bSf, d->bSf, 3 * size_t(nB));
        copyB(st->bMagSf, d->bMagSf, nB); copyB(st->bNf, d->bNf, 3 * size_t(nB)); copyB(st->bDelta, d->bDelta, 3 * size_t(nB));
        copyB(st->bWeights, d->bWeights, nB); copyB(st->bDeltaCoeffs, d->bDeltaCoeffs, nB);
        if (d->points && d->nPoints > 0) st->points.assign(d->points, d->points + 3 * size_t(d->nPoints));
        st->patchOffsets.assign(d->patchOffsets ? d->patchOffsets : nullptr, d->patchOffsets ? d->patchOffsets + d->nPatches + 1 : nullptr);
        if (st->patchOffsets.empty()) st->patchOffsets.assign(size_t(d->nPatches) + 1, 0);
        auto P = [](std::vector<double>& v) -> const double* { return v.empty() ? nullptr : v.data(); };
        fvk_mesh_desc& o = st->desc;
        o.nCells = nC; o.nInternalFaces = nI; o.nBoundaryFaces = nB; o.nPatches = d->nPatches;
        o.nPoints = st->points.empty() ? 0 : d->nPoints; o.points = P(st->points);
        o.cellVolumes = st->V.data(); o.cellCentres = st->C.data(); o.faceAreas = st->Sf.data(); o.faceCentres = st->Cf.data();
        o.magFaceAreas = st->magSf.data(); o.faceOwner = st->owner.data(); o.faceNeighbour = nI ? st->neighbour.data() : nullptr;
        o.faceCells = nB ? st->faceCells.data() : nullptr;
        o.bCf = P(st->bCf); o.bCn = P(st->bCn); o.bSf = P(st->bSf); o.bMagSf = P(st->bMagSf); o.bNf = P(st->bNf); o.bDelta = P(st->bDelta);
        o.bWeights = P(st->bWeights); o.bDeltaCoeffs = P(st->bDeltaCoeffs); o.patchOffsets = st->patchOffsets.data();
        o.nOwnedCells = 0; o.faceOrder = nullptr;
    }
    catch (const std::bad_alloc&)
    {
        delete st;
        return fvk_fail(FVK_ENOMEM, "fvk_renumber_apply: out of memory");
    }
    static_assert(offsetof(RenumStore, desc) == 0, "desc must be first");
    *out = &st->desc;
    return FVK_OK;
}

extern "C" int fvk_renumber_maps(const fvk_mesh_desc* renumbered, const int32_t** faceOldToNew, const uint8_t** faceFlipped)
{
    const RenumStore* st = reinterpret_cast<const RenumStore*>(renumbered);
    if (!st || st->magic != 0x52454e55u) return fvk_fail(FVK_EINVAL, "fvk_renumber_maps: not a mesh from fvk_renumber_apply");
    if (faceOldToNew) *faceOldToNew = st->faceOldToNew.data();
    if (faceFlipped) *faceFlipped = st->flipped.data();
    return FVK_OK;
}

extern "C" int fvk_renumber_destroy(fvk_mesh_desc* renumbered)
{
    RenumStore* st = reinterpret_cast<RenumStore*>(renumbered);
    if (st && st->magic == 0x52454e55u) delete st;
    return FVK_OK;
}
