// Internal declarations shared by the host (.cpp) and device (.cu) translation units of libfvk.
#pragma once
#include "fvk.h"

#include <cstddef>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

// Host work arrays that are FULLY written by parallel loops: std::vector::resize would value-initialise them on one thread first
// (gigabytes of zero pages at 256^3); with this allocator resize leaves them default-initialised and the pages are first touched
// by the threads that fill them.
template <class T>
struct FvkRawAllocator : std::allocator<T>
{
    template <class U> struct rebind { using other = FvkRawAllocator<U>; };
    FvkRawAllocator() = default;
    template <class U> FvkRawAllocator(const FvkRawAllocator<U>&) {}
    template <class U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using FvkRawVec = std::vector<T, FvkRawAllocator<T>>;


// record an error message (thread-local) and return `code`
int fvk_fail(int code, const char* fmt, ...);

// Tile plan of the explicit gather kernels (built once per mesh on the host, fvk_mesh.cu).
// Tile t owns the consecutive cells [c0, c0+nc) and the faces those cells own [f0, f0+nf) -- contiguous because
// OpenFOAM orders faces by owner. Every face value a tile needs gets a SLOT in shared memory:
//   [0, nf)            the tile's own faces (flux evaluated once, used by owner AND in-tile neighbour)
//   [nf, nf+nx)        "cross" faces: owned by another tile (or a ghost cell) with their neighbour in this tile
//   [nf+nx, nf+nx+nb)  boundary faces of the tile's cells
// Everything mesh-static a tile needs lives in ONE contiguous, 16-byte aligned blob (one TMA bulk copy):
//   V[nc] f64 | neighbour[nf] i32 | xFace[nx] i32 | xOwner[nx] i32 | bFace[nb] i32 |
//   seg[nc+1] u16 | oseg[nc+1] u16 | code[ne] u16 | xCell[nx] u16 | bCell[nb] u16     (each section 16-byte aligned)
// code[e] = (slot << 1) | side lists, per cell, the slots in the reference's accumulation order (side 1 = the
// cell is the face's neighbour -> subtract); seg / oseg are tile-local offsets into code / the own-face slots.
struct FvkTileHdr
{
    int32_t c0, nc, f0, nf, nx, nb, ne, blobBytes;
    int64_t blobOff;
};
struct FvkTilePlan
{
    int32_t nTiles = 0;
    int32_t nCells = 0, nB = 0; // array lengths (owned + ghost cells, boundary faces) for the staging bounds
    int32_t maxC = 0, maxF = 0, maxX = 0, maxB = 0, maxE = 0, maxBlob = 0;
    FvkTileHdr* hdr = nullptr;
    unsigned char* blob = nullptr;
};
// section offsets inside a tile blob
struct FvkBlobLayout
{
    int32_t nei, xFace, xOwner, bFace, seg, oseg, code, xCell, bCell, total;
};
#if defined(__CUDACC__)
__host__ __device__
#endif
inline FvkBlobLayout fvk_blob_layout(int32_t nc, int32_t nf, int32_t nx, int32_t nb, int32_t ne)
{
    auto al = [](int32_t n) { return (n + 15) & ~15; };
    FvkBlobLayout L;
    int32_t o = al(8 * nc);
    L.nei = o; o += al(4 * nf);
    L.xFace = o; o += al(4 * nx);
    L.xOwner = o; o += al(4 * nx);
    L.bFace = o; o += al(4 * nb);
    L.seg = o; o += al(2 * (nc + 1));
    L.oseg = o; o += al(2 * (nc + 1));
    L.code = o; o += al(2 * ne);
    L.xCell = o; o += al(2 * nx);
    L.bCell = o; o += al(2 * nb);
    L.total = o;
    return L;
}

// Brick plan of the explicit gather kernels (k_gather_brick, built once per mesh on the host, fvk_brickplan.cpp).
// A tile is a small strided set of cell RUNS {c0 + a*sy + b*sz + [0, runLen) : a < by, b < nRuns/by}: an (x,y,z)
// brick of a block-structured numbering, or one run of consecutive cells for any other numbering. Faces are in
// OpenFOAM order (sorted by owner), so the faces a cell owns are the consecutive ids [faceStart, faceStart+nOwn).
// Every face value a tile needs has a SLOT in shared memory:
//   [0, nOwnSlots)                 faces owned by the tile's cells, cell after cell (evaluated once, used by the owner
//                                  and by an in-tile neighbour)
//   [nOwnSlots, +nx)               "cross" faces: a tile cell is the neighbour, the owner lies outside the tile
//   [nOwnSlots+nx, +nb)            boundary faces of the tile's cells
// rec[recBase + lc] = {faceStart, slotBase | listBase << 16} for local cell lc (one extra record closes the tile);
// codes[codeBase + listBase ...] lists, in the reference's accumulation order, the slots the cell does not own:
// (slot << 1) | 1 = lower face (subtract), (slot << 1) | 0 = boundary face (add). The reference order per cell is
// [lower faces | owned faces | boundary faces] (checked at build time; meshes that violate it get no brick plan).
struct FvkBrickHdr // 64 bytes
{
    int32_t c0, runLen, by, nRuns;
    int32_t sy, sz;
    int32_t shiftL, shiftBy; // log2 of runLen / by when they are powers of two, else -1
    int32_t recBase, codeBase;
    int32_t xBase, nx;
    int32_t bBase, nb;
    int32_t nOwnSlots, nc;
};
struct FvkBrickRec
{
    int32_t faceStart;
    uint32_t bases; // slotBase | listBase << 16
};
// Geometry of the tiling, enough to compute a tile's cells WITHOUT loading anything (the kernel's first load level
// then already fetches per-cell data): structured = bricks of a block-structured numbering c = i + nx*(j + ny*k),
// tile t = (tx, ty, tz) in x-fastest order; otherwise tile t = cells [t*cap, t*cap + cap).
struct FvkBrickGeom
{
    int32_t structured = 0;
    int32_t dims[3] = {0, 0, 0};  // nx, ny, nz of the owned block
    int32_t brick[3] = {0, 0, 0}; // lx, by, bz
    int32_t tdim[2] = {0, 0};     // tiles along x, y
    int32_t shiftL = -1, shiftBy = -1; // log2(lx), log2(by) when powers of two (full tiles use shifts)
    int32_t cap = 0;              // cells per tile the per-cell arrays are strided with (= threads per block)
    int32_t nOwned = 0;
    // Block-structured topology PROVEN by the plan (fvk_brickplan.cpp checks every cell): cell c = i + nx(j + ny k) owns
    // the consecutive faces fs(c).. towards +x, +y, +z (those that exist), fs(c) = 3c - tx(j + ny k) - ty k nx for every
    // cell with j < ny-1 and k < nz-1 (t* = 1 where the block's upper side is a true boundary, 0 where it is a processor
    // cut), neighbours c+1, c+nx, c+nx*ny, and a REGULAR cell (0 < i < nx-1, ...) has exactly the stencil
    // [zL, yL, xL | x, y, z] with xL = fs(c-1), yL = fs(c-nx)+1, zL = fs(c-nx*ny)+2. k_gather_affine computes the regular
    // cells without reading any index array; the irregular ones (boundary / cut layers, irrCells) use the per-cell gather.
    int32_t affine = 0;
    int32_t tUp[3] = {1, 1, 1};
    int32_t tdimZ = 0;
    int32_t maxCross = 0; // cross faces of a full tile: lx*by + lx*bz + by*bz
};
#ifdef __CUDACC__
#define FVK_HD __host__ __device__ __forceinline__
#else
#define FVK_HD inline
#endif
// local cell lc of tile t -> global cell id; nc = number of cells of the tile. Same enumeration as the host plan.
FVK_HD int32_t fvk_brick_cell(const FvkBrickGeom& g, int32_t t, int32_t lc, int32_t& nc)
{
    if (!g.structured)
    {
        const int32_t c0 = t * g.cap;
        nc = g.nOwned - c0 < g.cap ? g.nOwned - c0 : g.cap;
        return c0 + lc;
    }
    const int32_t ix = t % g.tdim[0], q = t / g.tdim[0], iy = q % g.tdim[1], iz = q / g.tdim[1];
    const int32_t x0 = ix * g.brick[0], y0 = iy * g.brick[1], z0 = iz * g.brick[2];
    const int32_t rl = g.dims[0] - x0 < g.brick[0] ? g.dims[0] - x0 : g.brick[0];
    const int32_t ry = g.dims[1] - y0 < g.brick[1] ? g.dims[1] - y0 : g.brick[1];
    const int32_t rz = g.dims[2] - z0 < g.brick[2] ? g.dims[2] - z0 : g.brick[2];
    nc = rl * ry * rz;
    int32_t r, off, a, b;
    if (rl == g.brick[0] && ry == g.brick[1] && g.shiftL >= 0 && g.shiftBy >= 0)
    {
        r = lc >> g.shiftL; off = lc & (rl - 1); b = r >> g.shiftBy; a = r & (ry - 1);
    }
    else
    {
        r = lc / rl; off = lc - r * rl; b = r / ry; a = r - b * ry;
    }
    return (x0 + off) + g.dims[0] * ((y0 + a) + g.dims[1] * (z0 + b));
}

struct FvkBrickPlan
{
    int32_t nTiles = 0, maxSlots = 0, maxCells = 0;
    FvkBrickHdr* hdr = nullptr;
    FvkBrickRec* rec = nullptr;
    uint16_t* codes = nullptr;
    int32_t *xFace = nullptr, *xOwner = nullptr, *xNei = nullptr;
    int32_t *bFace = nullptr, *bCell = nullptr;
    // direct-indexed copies (no header needed): recF[t*(cap+1) + lc], codes4[t*cap + lc] = first 4 codes of the cell's
    // list, tileInfo[t] = {xBase, nx | nb << 16, bBase, nOwnSlots | touchesGhost << 30}
    FvkBrickGeom geom;
    FvkBrickRec* recF = nullptr;
    uint2* codes4 = nullptr;
    int4* tileInfo = nullptr;
    int32_t* irrCells = nullptr; // owned cells that are not regular (only when geom.affine), ascending
    int32_t nIrr = 0;
};

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
    // every row is [lower | diag | upper] in the order of the cell's stencil entries (true for OpenFOAM face order
    // without a faceOrder key): an internal entry's slot is its stencil position (+1 behind the diagonal)
    bool rowsInStencilOrder = false;
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
    FvkTilePlan tp; // tile plan of the explicit gather kernels (nTiles == 0: faces not sorted by owner)
    FvkBrickPlan bp; // brick plan (nTiles == 0: not available for this mesh)
    // halo overlap (fvk_mesh_set_tile_phase): 0 = operators compute every cell, 1 = only tiles that read no ghost
    // cell, 2 = only tiles that do
    int tilePhase = 0;
    // multicolour ordering of the cells for the DIC preconditioner (fvk_mesh_ensure_colors, built on first use): color [nCells]
    // (device), the owned cells sorted by colour (device) and the colour segments (host)
    mutable uint8_t* dicColor = nullptr;
    mutable int32_t* dicCells = nullptr;
    mutable int32_t dicOff[66] = {0};
    mutable int32_t dicNColors = 0;
};
// greedy colouring of the owned rows of the mesh's own sparsity pattern in natural cell order (no two coupled cells share a
// colour); cached in the handle. Returns an FVK_* code.
int fvk_mesh_ensure_colors(const fvk_mesh* m);
