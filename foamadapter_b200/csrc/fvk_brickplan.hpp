// Host-side builder of the brick plan (FvkBrickPlan, fvk_internal.hpp) and of the per-cell face stencil.
#pragma once
#include "fvk_internal.hpp"

#include <vector>

// cell -> faces in the reference's accumulation order (ascending face id, or ascending faceOrder key over the
// internal faces of a decomposed mesh; boundary faces last): CellToFaceStencil,
// src/NeoN/src/finiteVolume/cellCentred/stencil/cellToFaceStencil.cpp:14-96.
struct FvkStencilHost
{
    std::vector<int32_t> seg;  // [nC+1]
    FvkRawVec<int32_t> val;  // ascending local face id per cell (the reference's stencil values)
    FvkRawVec<int32_t> ent;  // (face << 1) | (cell is the face's neighbour), accumulation order
    FvkRawVec<int32_t> plan; // 2 per entry: {ent, other cell}; boundary face b: {-(b + 1), own cell} (only withPlan)
};
void fvk_build_stencil(const fvk_mesh_desc* d, FvkStencilHost& st, bool withPlan = false);

// SparsityPattern on the host (sparsityPattern.cpp:21-143): row = [lower (face order) | diag | upper (face order)], built row by row in
// parallel from the cell's stencil. A decomposed mesh (faceOrder key) orders each half by the GLOBAL face id -- see fvk_build_sparsity.
struct FvkSparsityHost
{
    std::vector<int32_t> rowOffs;            // [nC + 1]
    FvkRawVec<int32_t> col;                  // [nnz]
    FvkRawVec<uint8_t> ownOff, neiOff, diagOff; // [nI], [nI], [nC]
    bool rowsInStencilOrder = false;         // every row's entries sit at their stencil positions (+1 behind the diagonal)
    int32_t tooLongCell = -1;                // a row with > 255 entries (uint8 offsets): pattern not built
};
bool fvk_build_sparsity(const fvk_mesh_desc* d, const FvkStencilHost& st, FvkSparsityHost& out);

struct FvkBrickPlanHost
{
    std::vector<FvkBrickHdr> hdr;
    FvkRawVec<FvkBrickRec> rec;      // (raw: every entry is written by the parallel fill passes)
    FvkRawVec<uint16_t> codes;
    FvkRawVec<int32_t> xFace, xOwner, xNei, bFace, bCell;
    int32_t maxSlots = 0, maxCells = 0;
    // direct-indexed copies + tiling geometry (see FvkBrickPlan)
    FvkBrickGeom geom;
    FvkRawVec<FvkBrickRec> recF;
    FvkRawVec<uint2> codes4;
    std::vector<int4> tileInfo;
    std::vector<int32_t> irrCells;
    int32_t dims[3] = {0, 0, 0};  // detected block-structured numbering (0,0,0: none -> runs of consecutive cells)
    int32_t brick[3] = {0, 0, 0}; // brick shape used
};
// false: the mesh does not have the [lower | owned (consecutive ids) | boundary] per-cell order, or a tile exceeds the
// 16-bit slot range -> the caller keeps the per-cell gather kernel. `reason` (may be NULL) gets a short text.
bool fvk_build_brick_plan(const fvk_mesh_desc* d, const FvkStencilHost& st, FvkBrickPlanHost& out, const char** reason);
// replay the plan on the host exactly as k_gather_brick reads it and compare, cell by cell, the (face, sign) sequence
// with the stencil; returns the number of mismatching cells (0 = the plan reproduces the reference order)
int64_t fvk_verify_brick_plan(const fvk_mesh_desc* d, const FvkStencilHost& st, const FvkBrickPlanHost& bp);
