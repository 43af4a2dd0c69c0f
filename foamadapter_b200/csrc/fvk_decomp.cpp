// Domain decomposition (host): global mesh description + cell->rank map -> this rank's sub-mesh with ghost cells.
//
// The reference has no decomposed finite-volume path; its tutorials carry OpenFOAM
// decomposeParDict files (tutorials/cavity/system/decomposeParDict:17-24: hierarchical / simple, n (px py pz))
// that no solver uses. The layout defined here replaces OpenFOAM's processor patches with GHOST CELLS so every
// kernel of the single-domain path runs unchanged on a sub-domain:
//   cells   [0, nOwned)            this rank's cells, ascending global id
//           [nOwned, nOwned+nGhost) copies of other ranks' cells that share a face with an owned cell, grouped by
//                                  owning rank, ascending global id inside a group (= the receive layout)
//   faces   internal: every global internal face with at least one owned cell. Faces whose owner is owned come
//           first in global order (so they stay sorted by owner); faces owned by a ghost follow, sorted by ghost.
//           faceOrder = global face id, so per-cell accumulation follows the undecomposed mesh's order and
//           sub-domain results are bit-identical to single-domain results for the explicit operators.
//           boundary: the global boundary faces of owned cells, patch by patch, global order.
// The halo plan needs no communication to build: both sides sort the shared cells by global id.
#include "fvk_internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>
#include <parallel/algorithm>
#include <vector>

struct fvk_decomp
{
    fvk_mesh_desc local {};
    int32_t nOwned = 0, nGhost = 0;
    std::vector<double> V, C, Sf, Cf, magSf, bCf, bCn, bSf, bMagSf, bNf, bDelta, bWeights, bDeltaCoeffs;
    std::vector<int32_t> owner, neighbour, faceCells, patchOffsets, faceOrder;
    std::vector<int32_t> cellGlobal, faceGlobal;
    std::vector<int32_t> nbrRanks, sendOff, sendCells, recvOff;
};

// OpenFOAM `simple` geometric decomposition restated as a closed-form rule (parity unpinned: the reference holds no
// decomposed case): along each axis the cells are stably sorted by centre coordinate and cut into p equal-count
// groups; rank = bx + px * (by + py * bz). On a uniform N^3 block mesh this is cell (i,j,k) -> (i*px/N, j*py/N, k*pz/N).
extern "C" int fvk_decomp_simple_map(const fvk_mesh_desc* g, int px, int py, int pz, int32_t* cellRank)
{
    if (!g || !cellRank || px < 1 || py < 1 || pz < 1 || !g->cellCentres) return fvk_fail(FVK_EINVAL, "fvk_decomp_simple_map: bad argument");
    const int32_t nC = g->nCells;
    if (nC <= 0) return FVK_OK;
    std::vector<int32_t> idx(nC), grp(nC);
    std::vector<int64_t> key(nC);
    std::fill(cellRank, cellRank + nC, 0);
    const int p[3] = {px, py, pz};
    int stride = 1;
    for (int axis = 0; axis < 3; ++axis)
    {
        std::iota(idx.begin(), idx.end(), 0);
        const double* C = g->cellCentres;
        // quantise to suppress last-bit noise in equal coordinates, ties broken by cell id (stable)
        double lo = C[axis], hi = C[axis];
#pragma omp parallel for schedule(static) reduction(min : lo) reduction(max : hi)
        for (int32_t c = 0; c < nC; ++c) { lo = std::min(lo, C[3 * size_t(c) + axis]); hi = std::max(hi, C[3 * size_t(c) + axis]); }
        // integer keys: a tolerance comparison (xa < xb - eps) is not a strict weak ordering
        const double eps = (hi - lo) * 1e-9 + 1e-300;
#pragma omp parallel for schedule(static)
        for (int32_t c = 0; c < nC; ++c) key[c] = std::llround((C[3 * size_t(c) + axis] - lo) / eps);
        // (multi-threaded merge sort; stable, so the result is the serial std::stable_sort's)
        __gnu_parallel::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
#pragma omp parallel for schedule(static)
        for (int32_t k = 0; k < nC; ++k) grp[idx[k]] = int32_t((int64_t(k) * p[axis]) / nC);
#pragma omp parallel for schedule(static)
        for (int32_t c = 0; c < nC; ++c) cellRank[c] += stride * grp[c];
        stride *= p[axis];
    }
    return FVK_OK;
}

extern "C" int fvk_decomp_destroy(fvk_decomp* d)
{
    delete d;
    return FVK_OK;
}

extern "C" int fvk_decompose(const fvk_mesh_desc* g, const int32_t* cellRank, int nRanks, int rank, fvk_decomp** out)
{
    if (!g || !cellRank || !out || nRanks < 1 || rank < 0 || rank >= nRanks) return fvk_fail(FVK_EINVAL, "fvk_decompose: bad argument");
    *out = nullptr;
    const int32_t nC = g->nCells, nI = g->nInternalFaces, nB = g->nBoundaryFaces;
    for (int32_t c = 0; c < nC; ++c)
        if (cellRank[c] < 0 || cellRank[c] >= nRanks) return fvk_fail(FVK_EINVAL, "fvk_decompose: cell %d has rank %d", c, cellRank[c]);
    if (!g->faceOwner || (nI && !g->faceNeighbour) || (nB && (!g->faceCells || !g->patchOffsets))) return fvk_fail(FVK_EINVAL, "fvk_decompose: missing array");
    for (int32_t f = 0; f < nI; ++f)
        if (g->faceOwner[f] < 0 || g->faceOwner[f] >= nC || g->faceNeighbour[f] < 0 || g->faceNeighbour[f] >= nC)
            return fvk_fail(FVK_EINVAL, "fvk_decompose: face %d has bad owner/neighbour", f);
    for (int32_t b = 0; b < nB; ++b)
        if (g->faceCells[b] < 0 || g->faceCells[b] >= nC) return fvk_fail(FVK_EINVAL, "fvk_decompose: boundary face %d has bad faceCell", b);
    if (nB && (g->patchOffsets[0] != 0 || g->patchOffsets[g->nPatches] != nB)) return fvk_fail(FVK_EINVAL, "fvk_decompose: patchOffsets do not cover the boundary faces");
    fvk_decomp* d = new fvk_decomp;
    std::vector<int32_t> g2l(nC, -1);
    for (int32_t c = 0; c < nC; ++c)
        if (cellRank[c] == rank) { g2l[c] = int32_t(d->cellGlobal.size()); d->cellGlobal.push_back(c); }
    d->nOwned = int32_t(d->cellGlobal.size());
    // ghosts: (owning rank, global id), unique, sorted
    std::vector<std::pair<int32_t, int32_t>> ghosts;
    for (int32_t f = 0; f < nI; ++f)
    {
        const int32_t o = g->faceOwner[f], n = g->faceNeighbour[f];
        const bool oo = cellRank[o] == rank, nn = cellRank[n] == rank;
        if (oo && !nn) ghosts.emplace_back(cellRank[n], n);
        if (nn && !oo) ghosts.emplace_back(cellRank[o], o);
    }
    std::sort(ghosts.begin(), ghosts.end());
    ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
    d->nGhost = int32_t(ghosts.size());
    d->recvOff.assign(1, 0);
    for (size_t i = 0; i < ghosts.size(); ++i)
    {
        g2l[ghosts[i].second] = d->nOwned + int32_t(i);
        d->cellGlobal.push_back(ghosts[i].second);
        if (d->nbrRanks.empty() || d->nbrRanks.back() != ghosts[i].first)
        {
            if (!d->nbrRanks.empty()) d->recvOff.push_back(int32_t(i));
            d->nbrRanks.push_back(ghosts[i].first);
        }
    }
    if (!d->nbrRanks.empty()) d->recvOff.push_back(d->nGhost);
    // send lists: my cells that neighbour k holds as ghosts = owned cells on a cut face with rank k, by global id
    {
        std::map<int32_t, std::vector<int32_t>> send;
        for (int32_t f = 0; f < nI; ++f)
        {
            const int32_t o = g->faceOwner[f], n = g->faceNeighbour[f];
            const bool oo = cellRank[o] == rank, nn = cellRank[n] == rank;
            if (oo && !nn) send[cellRank[n]].push_back(o);
            if (nn && !oo) send[cellRank[o]].push_back(n);
        }
        d->sendOff.assign(1, 0);
        for (int32_t k : d->nbrRanks)
        {
            auto& v = send[k];
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
            for (int32_t c : v) d->sendCells.push_back(g2l[c]);
            d->sendOff.push_back(int32_t(d->sendCells.size()));
        }
    }
    // internal faces: owned-owner faces in global order, then ghost-owned faces sorted by (local owner, global id)
    std::vector<int32_t> first, second;
    for (int32_t f = 0; f < nI; ++f)
    {
        const int32_t o = g->faceOwner[f], n = g->faceNeighbour[f];
        const bool oo = cellRank[o] == rank, nn = cellRank[n] == rank;
        if (oo) first.push_back(f);
        else if (nn) second.push_back(f);
    }
    std::stable_sort(second.begin(), second.end(), [&](int32_t a, int32_t b) { return g2l[g->faceOwner[a]] < g2l[g->faceOwner[b]]; });
    std::vector<int32_t> faces(first);
    faces.insert(faces.end(), second.begin(), second.end());
    const int32_t lI = int32_t(faces.size());
    // boundary faces of owned cells, patch by patch
    std::vector<int32_t> bfaces;
    d->patchOffsets.assign(1, 0);
    for (int32_t p = 0; p < g->nPatches; ++p)
    {
        for (int32_t b = g->patchOffsets[p]; b < g->patchOffsets[p + 1]; ++b)
            if (cellRank[g->faceCells[b]] == rank) bfaces.push_back(b);
        d->patchOffsets.push_back(int32_t(bfaces.size()));
    }
    const int32_t lB = int32_t(bfaces.size()), lC = d->nOwned + d->nGhost;
    auto take3 = [](std::vector<double>& dst, const double* src, const std::vector<int32_t>& ids, int64_t off) {
        if (!src) return;
        for (int32_t i : ids)
            for (int k = 0; k < 3; ++k) dst.push_back(src[3 * (off + i) + k]);
    };
    auto take1 = [](std::vector<double>& dst, const double* src, const std::vector<int32_t>& ids, int64_t off) {
        if (!src) return;
        for (int32_t i : ids) dst.push_back(src[off + i]);
    };
    take1(d->V, g->cellVolumes, d->cellGlobal, 0);
    take3(d->C, g->cellCentres, d->cellGlobal, 0);
    take3(d->Sf, g->faceAreas, faces, 0); take3(d->Sf, g->faceAreas, bfaces, nI);
    take3(d->Cf, g->faceCentres, faces, 0); take3(d->Cf, g->faceCentres, bfaces, nI);
    take1(d->magSf, g->magFaceAreas, faces, 0); take1(d->magSf, g->magFaceAreas, bfaces, nI);
    for (int32_t f : faces)
    {
        d->owner.push_back(g2l[g->faceOwner[f]]);
        d->neighbour.push_back(g2l[g->faceNeighbour[f]]);
        d->faceOrder.push_back(f);
        d->faceGlobal.push_back(f);
    }
    for (int32_t b : bfaces)
    {
        d->faceCells.push_back(g2l[g->faceCells[b]]);
        d->owner.push_back(g2l[g->faceCells[b]]);
        d->faceGlobal.push_back(nI + b);
    }
    take3(d->bCf, g->bCf, bfaces, 0); take3(d->bCn, g->bCn, bfaces, 0); take3(d->bSf, g->bSf, bfaces, 0);
    take1(d->bMagSf, g->bMagSf, bfaces, 0); take3(d->bNf, g->bNf, bfaces, 0); take3(d->bDelta, g->bDelta, bfaces, 0);
    take1(d->bWeights, g->bWeights, bfaces, 0); take1(d->bDeltaCoeffs, g->bDeltaCoeffs, bfaces, 0);
    (void) nB;
    fvk_mesh_desc& L = d->local;
    L.nCells = lC; L.nInternalFaces = lI; L.nBoundaryFaces = lB; L.nPatches = g->nPatches; L.nPoints = 0; L.points = nullptr;
    auto ptr = [](std::vector<double>& v) -> const double* { return v.empty() ? nullptr : v.data(); };
    L.cellVolumes = ptr(d->V); L.cellCentres = ptr(d->C); L.faceAreas = ptr(d->Sf); L.faceCentres = ptr(d->Cf);
    L.magFaceAreas = ptr(d->magSf); L.faceOwner = d->owner.data(); L.faceNeighbour = d->neighbour.data();
    L.faceCells = d->faceCells.data(); L.bCf = ptr(d->bCf); L.bCn = ptr(d->bCn); L.bSf = ptr(d->bSf); L.bMagSf = ptr(d->bMagSf);
    L.bNf = ptr(d->bNf); L.bDelta = ptr(d->bDelta); L.bWeights = ptr(d->bWeights); L.bDeltaCoeffs = ptr(d->bDeltaCoeffs);
    L.patchOffsets = d->patchOffsets.data();
    L.nOwnedCells = d->nOwned; L.faceOrder = d->faceOrder.data();
    if (d->nOwned == 0) { delete d; return fvk_fail(FVK_EINVAL, "fvk_decompose: rank %d owns no cells", rank); }
    *out = d;
    return FVK_OK;
}

extern "C" const fvk_mesh_desc* fvk_decomp_mesh(const fvk_decomp* d) { return d ? &d->local : nullptr; }

extern "C" int fvk_decomp_info(const fvk_decomp* d, int32_t* nOwned, int32_t* nGhost, int32_t* nNeighbours)
{
    if (!d) return fvk_fail(FVK_EINVAL, "fvk_decomp_info: null");
    if (nOwned) *nOwned = d->nOwned;
    if (nGhost) *nGhost = d->nGhost;
    if (nNeighbours) *nNeighbours = int32_t(d->nbrRanks.size());
    return FVK_OK;
}

extern "C" int fvk_decomp_maps(const fvk_decomp* d, const int32_t** cellGlobal, const int32_t** faceGlobal)
{
    if (!d) return fvk_fail(FVK_EINVAL, "fvk_decomp_maps: null");
    if (cellGlobal) *cellGlobal = d->cellGlobal.data();
    if (faceGlobal) *faceGlobal = d->faceGlobal.data();
    return FVK_OK;
}

extern "C" int fvk_decomp_halo(const fvk_decomp* d, const int32_t** nbrRanks, const int32_t** sendOff, const int32_t** sendCells,
                               const int32_t** recvOff)
{
    if (!d) return fvk_fail(FVK_EINVAL, "fvk_decomp_halo: null");
    if (nbrRanks) *nbrRanks = d->nbrRanks.data();
    if (sendOff) *sendOff = d->sendOff.data();
    if (sendCells) *sendCells = d->sendCells.data();
    if (recvOff) *recvOff = d->recvOff.data();
    return FVK_OK;
}

extern "C" int fvk_comm_set_halo_from_decomp(fvk_comm* comm, const fvk_decomp* d)
{
    if (!comm || !d) return fvk_fail(FVK_EINVAL, "fvk_comm_set_halo_from_decomp: null");
    return fvk_comm_set_halo(comm, d->nOwned, int32_t(d->nbrRanks.size()), d->nbrRanks.data(), d->sendOff.data(), d->sendCells.data(),
                             d->recvOff.data());
}
