// Internal declarations shared by the host (.cpp) and device (.cu) translation units of libfvk.
#pragma once
#include "fvk.h"

#include <cstddef>
#include <cstdint>

// record an error message (thread-local) and return `code`
int fvk_fail(int code, const char* fmt, ...);

// Device-side mesh. All arrays are device pointers in reference order.
struct fvk_mesh
{
    int32_t nCells = 0, nInternalFaces = 0, nBoundaryFaces = 0, nPatches = 0;
    int32_t nOwned = 0; // cells [0, nOwned) are computed; [nOwned, nCells) are ghosts of a decomposed mesh
    int64_t nnz = 0;
    int device = 0;
    // UnstructuredMesh
    double *V = nullptr, *C = nullptr, *Sf = nullptr, *Cf = nullptr, *magSf = nullptr;
    int32_t *owner = nullptr, *neighbour = nullptr;
    // BoundaryMesh
    int32_t* faceCells = nullptr;
    double *bCf = nullptr, *bCn = nullptr, *bSf = nullptr, *bMagSf = nullptr, *bNf = nullptr,
           *bDelta = nullptr, *bWeights = nullptr, *bDeltaCoeffs = nullptr;
    int32_t patchOffsets[FVK_MAX_PATCHES + 1] = {0};
    // BasicGeometryScheme
    double *weights = nullptr, *deltaCoeffs = nullptr, *nonOrthDeltaCoeffs = nullptr;
    // CellToFaceStencil (sorted face ids per cell)
    int32_t *stencilSeg = nullptr, *stencilVal = nullptr;
    // gather plan: entry = (faceId << 1) | (cell is the face's neighbour), same order as stencilVal
    int32_t* gatherEnt = nullptr;
    // same order, 8 bytes per entry: {(faceId << 1) | side, other cell}; boundary face b: {-(b + 1), own cell}
    int32_t* gatherPlan = nullptr;
    // SparsityPattern
    int32_t *rowOffs = nullptr, *colIdxs = nullptr;
    uint8_t *ownerOffset = nullptr, *neighbourOffset = nullptr, *diagOffset = nullptr;
    // faces sorted by owner (OpenFOAM upper-triangular order)? then the faces owned by cell c are
    // [ownStart[c], ownStart[c+1])
    bool ownerSorted = false;
    int32_t* ownStart = nullptr; // [nCells+1], only if ownerSorted
    // lower faces (cell is the neighbour) per cell, ascending
    int32_t *lowSeg = nullptr, *lowFace = nullptr, *lowOwner = nullptr;
    // cells with boundary faces: sorted unique cell ids, segments into bndFace (boundary face ids)
    int32_t nBndCells = 0;
    int32_t *bndCell = nullptr, *bndSeg = nullptr, *bndFace = nullptr;
    uint32_t* hasBnd = nullptr; // bitmask [ceil(nCells/32)]
};
