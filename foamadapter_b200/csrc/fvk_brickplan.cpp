// Host-side plans of the explicit gather kernels: the cell->face stencil in the reference's accumulation order and
// the brick plan of k_gather_brick (fvk_explicit.cu). Nothing here is reference code: the reference scatters face
// fluxes with atomics (src/NeoN/src/finiteVolume/cellCentred/operators/gaussGreenDiv.cpp:69-101); the plan exists so
// that a cell-centric gather can reproduce the SerialExecutor's summation order (gaussGreenDiv.cpp:46-67) while
// streaming every face array once.
#include "fvk_brickplan.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <utility>

void fvk_build_stencil(const fvk_mesh_desc* d, FvkStencilHost& st, bool withPlan)
{
    const int32_t nC = d->nCells, nI = d->nInternalFaces, nB = d->nBoundaryFaces;
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    std::vector<int32_t>& seg = st.seg;
    seg.assign(size_t(nC) + 1, 0);
    // counting sort of the (cell, face) incidences, every pass parallel: count, scan, scatter in any order, then each cell
    // sorts its handful of entries by face id -- the order a serial visit of the faces in ascending id would append them in
    // (cellToFaceStencil.cpp:82-93 sorts too). A boundary face's key (nI + b) << 1 puts it behind the internal faces.
#pragma omp parallel for schedule(static)
    for (int32_t f = 0; f < nI; ++f)
    {
#pragma omp atomic
        ++seg[size_t(own[f]) + 1];
#pragma omp atomic
        ++seg[size_t(nei[f]) + 1];
    }
#pragma omp parallel for schedule(static)
    for (int32_t b = 0; b < nB; ++b)
    {
#pragma omp atomic
        ++seg[size_t(d->faceCells[b]) + 1];
    }
    for (int32_t c = 0; c < nC; ++c) seg[size_t(c) + 1] += seg[c];
    const size_t nEnt = size_t(seg[nC]);
    FvkRawVec<int32_t>&val = st.val, &ent = st.ent, &plan = st.plan;
    val.resize(nEnt); ent.resize(nEnt);
    if (withPlan) plan.resize(2 * nEnt); else plan.clear();
    {
        std::vector<int32_t> pos(seg.begin(), seg.end() - 1);
#pragma omp parallel for schedule(static)
        for (int32_t f = 0; f < nI; ++f)
        {
            int32_t k;
#pragma omp atomic capture
            k = pos[own[f]]++;
            ent[k] = f << 1;
#pragma omp atomic capture
            k = pos[nei[f]]++;
            ent[k] = (f << 1) | 1;
        }
#pragma omp parallel for schedule(static)
        for (int32_t b = 0; b < nB; ++b)
        {
            int32_t k;
#pragma omp atomic capture
            k = pos[d->faceCells[b]]++;
            ent[k] = (nI + b) << 1;
        }
    }
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < nC; ++c)
    {
        const int32_t b0 = seg[c], b1 = seg[size_t(c) + 1];
        for (int32_t x = b0 + 1; x < b1; ++x) // insertion sort: a cell has a handful of faces
        {
            const int32_t v = ent[x];
            int32_t y = x;
            for (; y > b0 && ent[y - 1] > v; --y) ent[y] = ent[y - 1];
            ent[y] = v;
        }
        for (int32_t k = b0; k < b1; ++k)
        {
            const int32_t f = ent[k] >> 1;
            val[k] = f;
            if (withPlan)
            {
                if (f < nI) { plan[2 * size_t(k)] = ent[k]; plan[2 * size_t(k) + 1] = (ent[k] & 1) ? own[f] : nei[f]; }
                else { plan[2 * size_t(k)] = -(f - nI + 1); plan[2 * size_t(k) + 1] = d->faceCells[f - nI]; }
            }
        }
    }
    if (d->faceOrder)
    {
        // per-cell gather order = ascending key over the internal faces (boundary faces stay last);
        // stencil values keep the ascending local id
        const int32_t* key = d->faceOrder;
#pragma omp parallel
        {
            std::vector<std::pair<int32_t, int32_t>> tmp;
            std::vector<int32_t> e2, p2;
#pragma omp for schedule(static)
            for (int32_t c = 0; c < nC; ++c)
            {
                int32_t b0 = seg[c], b1 = seg[size_t(c) + 1];
                while (b1 > b0 && (ent[b1 - 1] >> 1) >= nI) --b1;
                tmp.clear();
                for (int32_t k = b0; k < b1; ++k) tmp.emplace_back(key[ent[k] >> 1], k);
                if (std::is_sorted(tmp.begin(), tmp.end())) continue;
                std::stable_sort(tmp.begin(), tmp.end());
                e2.resize(tmp.size()); p2.resize(2 * tmp.size());
                for (size_t i = 0; i < tmp.size(); ++i)
                {
                    e2[i] = ent[tmp[i].second];
                    if (withPlan) { p2[2 * i] = plan[2 * size_t(tmp[i].second)]; p2[2 * i + 1] = plan[2 * size_t(tmp[i].second) + 1]; }
                }
                for (size_t i = 0; i < tmp.size(); ++i)
                {
                    ent[b0 + i] = e2[i];
                    if (withPlan) { plan[2 * (size_t(b0) + i)] = p2[2 * i]; plan[2 * (size_t(b0) + i) + 1] = p2[2 * i + 1]; }
                }
            }
        }
    }
}

namespace
{
constexpr int32_t kMaxTileCells = 512; // k_gather_brick: up to 2 cells per thread x 256 threads, or 512 threads
constexpr uint16_t kPad = 0xffffu;     // pads a cell's code list to a multiple of 4 (one aligned 8-byte load)
// cells per tile = threads per block (one cell per thread): FVK_BRICK_CELLS (64..512) overrides. Default 128 cells as a
// 16x4x2 brick: best of the r2 sweeps on B200 (profiles/r2_brick_sweep.md) -- small blocks retire and refill fastest.
int32_t tile_cells()
{
    int32_t v = 128;
    if (const char* e = std::getenv("FVK_BRICK_CELLS")) v = std::atoi(e);
    return std::max(64, std::min(kMaxTileCells, v));
}

struct TileShape
{
    int32_t c0, runLen, by, nRuns, sy, sz;
};

int32_t log2_exact(int32_t v)
{
    if (v <= 0 || (v & (v - 1))) return -1;
    int32_t s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

// Block-structured numbering c = i + nx*(j + ny*k) of the computed cells, read off the owner->neighbour id strides.
// Purely a tiling hint: any partition of the cells gives a correct plan.
bool detect_dims(int32_t nOwned, int32_t nI, const int32_t* own, const int32_t* nei, int32_t dims[3])
{
    // at most three distinct owner->neighbour strides (kept in a 3-entry table: this loop visits every face)
    int32_t known[3] = {0, 0, 0};
    int nKnown = 0;
    int tooMany = 0, nonPositive = 0;
#pragma omp parallel
    {
        int32_t mine[3] = {0, 0, 0};
        int nMine = 0, bad = 0, neg = 0;
#pragma omp for schedule(static) nowait
        for (int32_t f = 0; f < nI; ++f)
        {
            if (own[f] >= nOwned || nei[f] >= nOwned) continue;
            const int32_t dlt = nei[f] - own[f];
            if (dlt <= 0) { neg = 1; continue; }
            if ((nMine > 0 && dlt == mine[0]) || (nMine > 1 && dlt == mine[1]) || (nMine > 2 && dlt == mine[2])) continue;
            if (nMine == 3) { bad = 1; continue; }
            mine[nMine++] = dlt;
        }
#pragma omp critical
        {
            tooMany |= bad; nonPositive |= neg;
            for (int q = 0; q < nMine; ++q)
            {
                bool have = false;
                for (int r = 0; r < nKnown; ++r) have = have || known[r] == mine[q];
                if (!have) { if (nKnown == 3) tooMany = 1; else known[nKnown++] = mine[q]; }
            }
        }
    }
    if (tooMany || nonPositive) return false;
    std::map<int32_t, int64_t> diffs;
    for (int r = 0; r < nKnown; ++r) diffs[known[r]] = 1;
    if (diffs.empty()) { dims[0] = nOwned; dims[1] = dims[2] = 1; return true; }
    std::vector<int32_t> s;
    for (auto& kv : diffs) s.push_back(kv.first);
    if (s[0] != 1) return false;
    if (s.size() == 1) { dims[0] = nOwned; dims[1] = dims[2] = 1; return true; }
    if (nOwned % s[1]) return false;
    if (s.size() == 2) { dims[0] = s[1]; dims[1] = nOwned / s[1]; dims[2] = 1; return true; }
    if (s[2] % s[1] || nOwned % s[2]) return false;
    dims[0] = s[1]; dims[1] = s[2] / s[1]; dims[2] = nOwned / s[2];
    return true;
}

void make_tiles(int32_t nOwned, const int32_t dims[3], bool structured, int32_t brick[3], std::vector<TileShape>& tiles)
{
    int32_t kTileCells = tile_cells();
    tiles.clear();
    if (!structured)
    {
        kTileCells = kTileCells <= 128 ? 128 : (kTileCells <= 256 ? 256 : 512); // = FvkBrickGeom::cap
        brick[0] = kTileCells; brick[1] = brick[2] = 1;
        for (int32_t c = 0; c < nOwned; c += kTileCells)
            tiles.push_back(TileShape {c, std::min(kTileCells, nOwned - c), 1, 1, 0, 0});
        return;
    }
    int32_t L = kTileCells >= 256 ? 32 : 16, BY = 4, BZ = kTileCells >= 512 ? 4 : 2;
    bool fixed = false; // FVK_BRICK="lx,by,bz": use exactly this shape (clipped to the mesh and to 512 cells)
    if (const char* e = std::getenv("FVK_BRICK"))
    {
        int a = 0, b = 0, c = 0;
        if (std::sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0) { L = a; BY = b; BZ = c; fixed = true; kTileCells = kMaxTileCells; }
    }
    const int32_t nx = dims[0], ny = dims[1], nz = dims[2];
    int32_t lx = std::min(std::min(L, nx), kTileCells), bz = std::min(BZ, nz);
    if (lx * bz > kTileCells) bz = std::max(1, kTileCells / lx);
    int32_t by = std::min(ny, BY);
    if (!fixed) by = std::min(ny, std::max(BY, kTileCells / (lx * bz))); // flat meshes: more rows per tile
    while (by > 1 && lx * by * bz > kTileCells) --by;
    if (!fixed)
    {
        if (lx * by * bz < kTileCells && bz < nz) bz = std::min(nz, kTileCells / (lx * by)); // thin in y: stack more planes
        if (lx * by * bz < kTileCells && lx < nx) lx = std::min(nx, kTileCells / (by * bz));  // 1-D / thin meshes: longer runs
    }
    brick[0] = lx; brick[1] = by; brick[2] = bz;
    for (int32_t z0 = 0; z0 < nz; z0 += bz)
        for (int32_t y0 = 0; y0 < ny; y0 += by)
            for (int32_t x0 = 0; x0 < nx; x0 += lx)
            {
                const int32_t rl = std::min(lx, nx - x0), ry = std::min(by, ny - y0), rz = std::min(bz, nz - z0);
                tiles.push_back(TileShape {x0 + nx * (y0 + ny * z0), rl, ry, ry * rz, nx, nx * ny});
            }
}

template <class F>
inline void for_tile_cells(const TileShape& t, F&& fn)
{
    int32_t lc = 0;
    for (int32_t r = 0; r < t.nRuns; ++r)
    {
        const int32_t a = r % t.by, b = r / t.by;
        const int32_t base = t.c0 + a * t.sy + b * t.sz;
        for (int32_t o = 0; o < t.runLen; ++o, ++lc) fn(base + o, lc);
    }
}

struct CellInfo
{
    int32_t fs = 0, nOwn = 0, nLow = 0, nBnd = 0;
    bool ok = true;
};
// the per-cell order the kernel assumes: [lower faces | owned faces, consecutive ascending ids | boundary faces]
inline CellInfo analyse(int32_t c, int32_t nI, const int32_t* seg, const int32_t* ent)
{
    CellInfo ci;
    int phase = 0;
    for (int32_t e = seg[c]; e < seg[size_t(c) + 1]; ++e)
    {
        const int32_t f = ent[e] >> 1, side = ent[e] & 1;
        if (f >= nI)
        {
            if (side) { ci.ok = false; return ci; }
            phase = 2; ++ci.nBnd;
        }
        else if (side)
        {
            if (phase != 0) { ci.ok = false; return ci; }
            ++ci.nLow;
        }
        else
        {
            if (phase == 2) { ci.ok = false; return ci; }
            if (phase == 0) { phase = 1; ci.fs = f; }
            if (f != ci.fs + ci.nOwn) { ci.ok = false; return ci; }
            ++ci.nOwn;
        }
    }
    return ci;
}
} // namespace

bool fvk_build_brick_plan(const fvk_mesh_desc* d, const FvkStencilHost& st, FvkBrickPlanHost& out, const char** reason)
{
    const char* dummy;
    if (!reason) reason = &dummy;
    *reason = "";
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    const int32_t nOwned = (d->nOwnedCells > 0 && d->nOwnedCells <= nC) ? d->nOwnedCells : nC;
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    const int32_t* seg = st.seg.data();
    const int32_t* ent = st.ent.data();
    out = FvkBrickPlanHost {};
    if (nOwned <= 0) { *reason = "no cells"; return false; }

    const bool structured = detect_dims(nOwned, nI, own, nei, out.dims)
                            && int64_t(out.dims[0]) * out.dims[1] * out.dims[2] == nOwned;
    if (!structured) out.dims[0] = out.dims[1] = out.dims[2] = 0;
    std::vector<TileShape> tiles;
    make_tiles(nOwned, out.dims, structured, out.brick, tiles);
    const int32_t nT = int32_t(tiles.size());

    FvkRawVec<int32_t> cellTile, cellFaceStart, cellSlotBase; // (filled in parallel below)
    cellTile.resize(size_t(nC)); cellFaceStart.resize(size_t(nC)); cellSlotBase.resize(size_t(nC));
#pragma omp parallel for schedule(static)
    for (int32_t c = 0; c < nC; ++c) { cellTile[c] = -1; cellFaceStart[c] = 0; cellSlotBase[c] = 0; }
    std::vector<int32_t> tSlots(nT, 0), tCodes(nT, 0), tX(nT, 0), tB(nT, 0), tC(nT, 0);
    int bad = 0;
    // pass 1: per-cell order check, own-slot numbering
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int32_t t = 0; t < nT; ++t)
    {
        int32_t slots = 0, codes = 0, nb = 0, nc = 0;
        for_tile_cells(tiles[t], [&](int32_t c, int32_t) {
            if (c < 0 || c >= nOwned) { bad |= 1; return; }
            const CellInfo ci = analyse(c, nI, seg, ent);
            if (!ci.ok) { bad |= 2; return; }
            cellTile[c] = t; cellFaceStart[c] = ci.fs; cellSlotBase[c] = slots;
            slots += ci.nOwn; codes += (ci.nLow + ci.nBnd + 3) & ~3; nb += ci.nBnd; ++nc;
        });
        tSlots[t] = slots; tCodes[t] = codes; tB[t] = nb; tC[t] = nc;
    }
    if (bad & 1) { *reason = "tiling left the cell range"; return false; }
    if (bad & 2) { *reason = "per-cell face order is not [lower | owned consecutive | boundary]"; return false; }
    for (int32_t c = 0; c < nOwned; ++c)
        if (cellTile[c] < 0) { *reason = "tiling does not cover every cell"; return false; }
    // pass 1b: cross faces (lower faces whose owner is not a cell of the same tile)
#pragma omp parallel for schedule(static)
    for (int32_t t = 0; t < nT; ++t)
    {
        int32_t nx = 0;
        for_tile_cells(tiles[t], [&](int32_t c, int32_t) {
            for (int32_t e = seg[c]; e < seg[size_t(c) + 1]; ++e)
            {
                const int32_t f = ent[e] >> 1;
                if (f < nI && (ent[e] & 1))
                {
                    const int32_t o = own[f];
                    if (!(o < nOwned && cellTile[o] == t)) ++nx;
                }
            }
        });
        tX[t] = nx;
    }
    out.hdr.resize(nT);
    int64_t recBase = 0, codeBase = 0, xBase = 0, bBase = 0;
    for (int32_t t = 0; t < nT; ++t)
    {
        const TileShape& s = tiles[t];
        if (tSlots[t] + tX[t] + tB[t] >= 32767 || tCodes[t] >= 65535 || tC[t] > kMaxTileCells)
        {
            *reason = "tile exceeds the 16-bit slot range";
            return false;
        }
        if (recBase + tC[t] + 1 >= (int64_t(1) << 31) || codeBase + tCodes[t] >= (int64_t(1) << 31))
        {
            *reason = "plan exceeds 2^31 entries";
            return false;
        }
        FvkBrickHdr& h = out.hdr[t];
        h.c0 = s.c0; h.runLen = s.runLen; h.by = s.by; h.nRuns = s.nRuns; h.sy = s.sy; h.sz = s.sz;
        h.shiftL = log2_exact(s.runLen); h.shiftBy = log2_exact(s.by);
        h.recBase = int32_t(recBase); h.codeBase = int32_t(codeBase);
        h.xBase = int32_t(xBase); h.nx = tX[t]; h.bBase = int32_t(bBase); h.nb = tB[t];
        h.nOwnSlots = tSlots[t]; h.nc = tC[t];
        recBase += tC[t] + 1; codeBase += tCodes[t]; xBase += tX[t]; bBase += tB[t];
        out.maxSlots = std::max(out.maxSlots, tSlots[t] + tX[t] + tB[t]);
        out.maxCells = std::max(out.maxCells, tC[t]);
    }
    out.rec.resize(size_t(recBase));
    out.codes.resize(size_t(codeBase) + 8);       // every list is a multiple of 4 codes (codeBase too), padded with kPad by pass 2
    for (int q = 0; q < 8; ++q) out.codes[size_t(codeBase) + q] = kPad;
    out.xFace.resize(size_t(xBase)); out.xOwner.resize(size_t(xBase)); out.xNei.resize(size_t(xBase));
    out.bFace.resize(size_t(bBase)); out.bCell.resize(size_t(bBase));
    // pass 2: fill
#pragma omp parallel for schedule(static)
    for (int32_t t = 0; t < nT; ++t)
    {
        const FvkBrickHdr& h = out.hdr[t];
        int32_t list = 0, ix = 0, ib = 0, lastLc = -1;
        for_tile_cells(tiles[t], [&](int32_t c, int32_t lc) {
            FvkBrickRec& r = out.rec[size_t(h.recBase) + lc];
            r.faceStart = cellFaceStart[c];
            r.bases = uint32_t(cellSlotBase[c]) | (uint32_t(list) << 16);
            for (int32_t e = seg[c]; e < seg[size_t(c) + 1]; ++e)
            {
                const int32_t f = ent[e] >> 1, side = ent[e] & 1;
                if (f >= nI)
                {
                    const int32_t slot = h.nOwnSlots + h.nx + ib;
                    out.bFace[size_t(h.bBase) + ib] = f; out.bCell[size_t(h.bBase) + ib] = c;
                    ++ib;
                    out.codes[size_t(h.codeBase) + list++] = uint16_t(slot << 1);
                }
                else if (side)
                {
                    const int32_t o = own[f];
                    int32_t slot;
                    if (o < nOwned && cellTile[o] == t)
                        slot = cellSlotBase[o] + (f - cellFaceStart[o]);
                    else
                    {
                        slot = h.nOwnSlots + ix;
                        out.xFace[size_t(h.xBase) + ix] = f; out.xOwner[size_t(h.xBase) + ix] = o; out.xNei[size_t(h.xBase) + ix] = c;
                        ++ix;
                    }
                    out.codes[size_t(h.codeBase) + list++] = uint16_t((slot << 1) | 1);
                }
            }
            while (list & 3) out.codes[size_t(h.codeBase) + list++] = kPad;
            lastLc = lc;
        });
        FvkBrickRec& r = out.rec[size_t(h.recBase) + lastLc + 1]; // closing record
        r.faceStart = 0;
        r.bases = uint32_t(h.nOwnSlots) | (uint32_t(list) << 16);
    }
    // ---- direct-indexed copies: the kernel's first load level needs no header
    FvkBrickGeom& g = out.geom;
    g.structured = structured ? 1 : 0;
    g.nOwned = nOwned;
    g.cap = out.maxCells <= 128 ? 128 : (out.maxCells <= 256 ? 256 : 512);
    for (int k = 0; k < 3; ++k) { g.dims[k] = out.dims[k]; g.brick[k] = out.brick[k]; }
    if (structured)
    {
        g.tdim[0] = (g.dims[0] + g.brick[0] - 1) / g.brick[0];
        g.tdim[1] = (g.dims[1] + g.brick[1] - 1) / g.brick[1];
        g.shiftL = log2_exact(g.brick[0]); g.shiftBy = log2_exact(g.brick[1]);
    }
    else
        g.brick[0] = g.cap; // tiles of `cap` consecutive cells (make_tiles used the same size)
    out.recF.resize(size_t(nT) * (g.cap + 1));    // the entries behind a ragged tile's cells are padded in the loop below
    out.codes4.resize(size_t(nT) * g.cap);
    out.tileInfo.resize(nT);
    int geomBad = 0;
#pragma omp parallel for schedule(static) reduction(| : geomBad)
    for (int32_t t = 0; t < nT; ++t)
    {
        const FvkBrickHdr& h = out.hdr[t];
        bool ghost = false;
        for (int32_t lc = 0; lc <= h.nc; ++lc) out.recF[size_t(t) * (g.cap + 1) + lc] = out.rec[size_t(h.recBase) + lc];
        for (int32_t lc = h.nc + 1; lc <= g.cap; ++lc) out.recF[size_t(t) * (g.cap + 1) + lc] = FvkBrickRec {0, 0};
        for (int32_t lc = h.nc; lc < g.cap; ++lc) out.codes4[size_t(t) * g.cap + lc] = make_uint2(0xffffffffu, 0xffffffffu);
        for_tile_cells(tiles[t], [&](int32_t c, int32_t lc) {
            int32_t nc = 0;
            if (fvk_brick_cell(g, t, lc, nc) != c || nc != h.nc) geomBad |= 1;
            const FvkBrickRec r0 = out.rec[size_t(h.recBase) + lc], r1 = out.rec[size_t(h.recBase) + lc + 1];
            const int32_t listBase = int32_t(r0.bases >> 16), nList = int32_t(r1.bases >> 16) - listBase;
            uint16_t cd[4] = {kPad, kPad, kPad, kPad};
            for (int32_t j = 0; j < 4 && j < nList; ++j) cd[j] = out.codes[size_t(h.codeBase) + listBase + j];
            out.codes4[size_t(t) * g.cap + lc] = make_uint2(uint32_t(cd[0]) | (uint32_t(cd[1]) << 16), uint32_t(cd[2]) | (uint32_t(cd[3]) << 16));
            const int32_t nOwn = int32_t(r1.bases & 0xffffu) - int32_t(r0.bases & 0xffffu);
            for (int32_t k = 0; k < nOwn; ++k)
                if (nei[r0.faceStart + k] >= nOwned) ghost = true;
        });
        for (int32_t i = 0; i < h.nx; ++i) // a lower face whose owner is a ghost cell (the ghost side keeps the global orientation)
            if (out.xOwner[size_t(h.xBase) + i] >= nOwned) ghost = true;
        out.tileInfo[t] = make_int4(h.xBase, h.nx | (h.nb << 16), h.bBase, h.nOwnSlots | (ghost ? (1 << 30) : 0));
        if (h.nx >= 65536 || h.nb >= 32768) geomBad |= 2;
    }
    if (geomBad) { *reason = "tile geometry does not reproduce the tiling"; return false; }
    // ---- affine topology: prove, cell by cell, the closed form k_gather_affine assumes (see FvkBrickGeom)
    static const bool noAffine = [] { const char* e = std::getenv("FVK_NO_AFFINE"); return e && *e == '1'; }();
    g.tdimZ = structured ? (g.dims[2] + g.brick[2] - 1) / g.brick[2] : 0;
    g.maxCross = g.brick[0] * g.brick[1] + g.brick[0] * g.brick[2] + g.brick[1] * g.brick[2];
    if (!noAffine && structured && g.dims[0] >= 3 && g.dims[1] >= 3 && g.dims[2] >= 3 && int64_t(3) * nOwned + 8 < (int64_t(1) << 31))
    {
        const int64_t nx = g.dims[0], ny = g.dims[1], nz = g.dims[2], nxy = nx * ny;
        for (int cq = 7; cq >= 0 && !g.affine; --cq) // (1,1,1) -- a whole block, the common case -- first
        {
            const int combo = cq;
            const int64_t tx = combo & 1, ty = (combo >> 1) & 1, tz = (combo >> 2) & 1;
            auto fsOf = [&](int64_t c) {
                const int64_t i = c % nx, j = (c / nx) % ny, k = c / nxy;
                return 3 * c - tx * (j + ny * k) - ty * (k * nx + (j == ny - 1 ? i : 0)) - tz * (k == nz - 1 ? i + nx * j : 0);
            };
            int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
            for (int32_t c = 0; c < nOwned; ++c)
            {
                if (bad) continue;
                const int64_t i = c % nx, j = (c / nx) % ny, k = c / nxy;
                const bool hasX = !(tx && i == nx - 1), hasY = !(ty && j == ny - 1), hasZ = !(tz && k == nz - 1);
                const int64_t fs = fsOf(c), nOwn = int64_t(hasX) + hasY + hasZ;
                if (fs < 0 || fs + nOwn > nI) { bad |= 1; continue; }
                for (int64_t q = 0; q < nOwn; ++q)
                    if (own[fs + q] != c) bad |= 1;
                if ((fs + nOwn < nI && own[fs + nOwn] == c) || (fs > 0 && own[fs - 1] == c)) bad |= 1;
                int64_t q = 0;
                if (hasX) { const int32_t n = nei[fs + q++]; if (i < nx - 1 ? n != c + 1 : n < nOwned) bad |= 1; }
                if (hasY) { const int32_t n = nei[fs + q++]; if (j < ny - 1 ? n != c + nx : n < nOwned) bad |= 1; }
                if (hasZ) { const int32_t n = nei[fs + q++]; if (k < nz - 1 ? n != c + nxy : n < nOwned) bad |= 1; }
                if (bad) continue;
                if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && k > 0 && k < nz - 1)
                {
                    const int64_t want[6] = {((fsOf(c - nxy) + 2) << 1) | 1, ((fsOf(c - nx) + 1) << 1) | 1, (fsOf(c - 1) << 1) | 1,
                                             fs << 1, (fs + 1) << 1, (fs + 2) << 1};
                    if (seg[size_t(c) + 1] - seg[c] != 6) { bad |= 1; continue; }
                    for (int e = 0; e < 6; ++e)
                        if (ent[seg[c] + e] != want[e]) bad |= 1;
                    // the simplified forms the kernel uses
                    const int64_t f0 = 3 * int64_t(c) - tx * (j + ny * k) - ty * k * nx;
                    if (f0 != fs || fsOf(c - 1) != f0 - 3 || fsOf(c - nx) + 1 != f0 - 3 * nx + tx + 1
                        || fsOf(c - nxy) + 2 != f0 - 3 * nxy + tx * ny + ty * nx + 2)
                        bad |= 1;
                }
            }
            if (!bad)
            {
                g.affine = 1;
                g.tUp[0] = int32_t(tx); g.tUp[1] = int32_t(ty); g.tUp[2] = int32_t(tz);
            }
        }
        if (g.affine)
        { // the irregular cells in ascending id: count per z-plane, scan, fill (parallel)
            std::vector<int64_t> planeOff(size_t(nz) + 1, 0);
            for (int64_t k = 0; k < nz; ++k)
                planeOff[size_t(k) + 1] = planeOff[k] + ((k == 0 || k == nz - 1) ? nxy : nxy - std::max<int64_t>(nx - 2, 0) * std::max<int64_t>(ny - 2, 0));
            out.irrCells.resize(size_t(planeOff[nz]));
#pragma omp parallel for schedule(static)
            for (int64_t k = 0; k < nz; ++k)
            {
                int64_t w = planeOff[k];
                for (int64_t j = 0; j < ny; ++j)
                    for (int64_t i = 0; i < nx; ++i)
                        if (!(i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && k > 0 && k < nz - 1)) out.irrCells[size_t(w++)] = int32_t(i + nx * j + nxy * k);
            }
        }
    }
    return true;
}

int64_t fvk_verify_brick_plan(const fvk_mesh_desc* d, const FvkStencilHost& st, const FvkBrickPlanHost& bp)
{
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    const int32_t nOwned = (d->nOwnedCells > 0 && d->nOwnedCells <= nC) ? d->nOwnedCells : nC;
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    std::vector<uint8_t> seen(size_t(nOwned), 0);
    int64_t badCells = 0;
    for (size_t t = 0; t < bp.hdr.size(); ++t)
    {
        const FvkBrickHdr& h = bp.hdr[t];
        const int32_t nSlots = h.nOwnSlots + h.nx + h.nb;
        std::vector<int64_t> slotFace(size_t(nSlots), -1);
        std::vector<int32_t> cellOf(size_t(h.nc), -1);
        bool tileOk = nSlots <= bp.maxSlots && h.nc <= bp.maxCells && h.nc == h.runLen * h.nRuns;
        if (h.shiftL >= 0 && (1 << h.shiftL) != h.runLen) tileOk = false;
        if (h.shiftBy >= 0 && (1 << h.shiftBy) != h.by) tileOk = false;
        // phase A1 as the kernel does it
        for (int32_t lc = 0; lc < h.nc; ++lc)
        {
            const int32_t r = lc / h.runLen, off = lc - r * h.runLen, b = r / h.by, a = r - b * h.by;
            const int32_t c = h.c0 + a * h.sy + b * h.sz + off;
            cellOf[lc] = c;
            if (c < 0 || c >= nOwned || seen[c]) { tileOk = false; continue; }
            seen[c] = 1;
            const FvkBrickRec r0 = bp.rec[size_t(h.recBase) + lc], r1 = bp.rec[size_t(h.recBase) + lc + 1];
            const int32_t slotBase = int32_t(r0.bases & 0xffffu), nOwn = int32_t(r1.bases & 0xffffu) - slotBase;
            if (nOwn < 0 || slotBase + nOwn > h.nOwnSlots) { tileOk = false; continue; }
            for (int32_t k = 0; k < nOwn; ++k)
            {
                const int32_t f = r0.faceStart + k;
                if (f < 0 || f >= nI || own[f] != c) { tileOk = false; continue; }
                slotFace[size_t(slotBase) + k] = f;
            }
        }
        for (int32_t i = 0; i < h.nx; ++i)
        {
            const int32_t f = bp.xFace[size_t(h.xBase) + i];
            if (f < 0 || f >= nI || own[f] != bp.xOwner[size_t(h.xBase) + i] || nei[f] != bp.xNei[size_t(h.xBase) + i]) { tileOk = false; continue; }
            slotFace[size_t(h.nOwnSlots) + i] = f;
        }
        for (int32_t i = 0; i < h.nb; ++i)
        {
            const int32_t f = bp.bFace[size_t(h.bBase) + i];
            if (f < nI || f >= nI + d->nBoundaryFaces || d->faceCells[f - nI] != bp.bCell[size_t(h.bBase) + i]) { tileOk = false; continue; }
            slotFace[size_t(h.nOwnSlots) + h.nx + i] = f;
        }
        // phase B
        for (int32_t lc = 0; lc < h.nc; ++lc)
        {
            const int32_t c = cellOf[lc];
            if (c < 0 || c >= nOwned) { ++badCells; continue; }
            const FvkBrickRec r0 = bp.rec[size_t(h.recBase) + lc], r1 = bp.rec[size_t(h.recBase) + lc + 1];
            const int32_t slotBase = int32_t(r0.bases & 0xffffu), nOwn = int32_t(r1.bases & 0xffffu) - slotBase;
            const int32_t listBase = int32_t(r0.bases >> 16), nList = int32_t(r1.bases >> 16) - listBase;
            std::vector<int32_t> seq;
            bool ok = tileOk && nList >= 0;
            int32_t i = 0;
            auto pull = [&](uint16_t code, int32_t wantSide) {
                const int32_t slot = code >> 1;
                if (slot >= nSlots || slotFace[slot] < 0) { ok = false; return; }
                const int64_t f = slotFace[slot];
                if (wantSide) { if (f >= nI || nei[f] != c) ok = false; }
                else if (f < nI || d->faceCells[f - nI] != c) ok = false;
                seq.push_back(int32_t((f << 1) | wantSide));
            };
            if ((listBase & 3) || (nList & 3) || (h.codeBase & 3)) ok = false; // aligned 8-byte code loads
            auto codeAt = [&](int32_t k) { return bp.codes[size_t(h.codeBase) + listBase + k]; };
            for (; ok && i < nList && codeAt(i) != kPad && (codeAt(i) & 1); ++i) pull(codeAt(i), 1);
            for (int32_t k = 0; ok && k < nOwn; ++k) seq.push_back((r0.faceStart + k) << 1);
            for (; ok && i < nList; ++i)
            {
                const uint16_t code = codeAt(i);
                if (code == kPad) break; // padding only at the end of a list
                if (code & 1) { ok = false; break; }
                pull(code, 0);
            }
            for (; ok && i < nList; ++i)
                if (codeAt(i) != kPad) ok = false;
            const int32_t e0 = st.seg[c], e1 = st.seg[size_t(c) + 1];
            if (ok && int32_t(seq.size()) == e1 - e0)
            {
                for (int32_t e = e0; e < e1; ++e)
                    if (seq[size_t(e - e0)] != st.ent[e]) ok = false;
            }
            else
                ok = false;
            if (!ok) ++badCells;
        }
    }
    for (int32_t c = 0; c < nOwned; ++c)
        if (!seen[c]) ++badCells;
    return badCells;
}

// ---- sparsity pattern: row = [lower (face order) | diag | upper (face order)] (sparsityPattern.cpp:21-143). The reference's three
// serial passes over the faces place, in row r, the lower entries in ascending face id, the diagonal, then the upper entries in
// ascending face id; built here row by row in parallel from the cell's stencil (sorted by LOCAL face id, which is what the passes
// visit). A decomposed mesh (faceOrder key; the reference has no such path) sorts each half by the GLOBAL face id instead: the
// ghost-owned faces, which the sub-domain numbers last, then sit where the undecomposed mesh has them, the rows stay in stencil
// order and the index-free kernels serve every rank's sub-domain.
bool fvk_build_sparsity(const fvk_mesh_desc* d, const FvkStencilHost& sth, FvkSparsityHost& out)
{
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    const int32_t* fkey = d->faceOrder;
    std::vector<int32_t>& rowOffs = out.rowOffs;
    rowOffs.assign(size_t(nC) + 1, 0);
    int tooLong = -1;
#pragma omp parallel for schedule(static) reduction(max : tooLong)
    for (int32_t c = 0; c < nC; ++c)
    {
        int32_t n = 1;
        for (int32_t e = sth.seg[c]; e < sth.seg[size_t(c) + 1]; ++e) n += (sth.ent[e] >> 1) < nI;
        if (n > 255) tooLong = std::max(tooLong, c);
        rowOffs[size_t(c) + 1] = n;
    }
    out.tooLongCell = tooLong;
    if (tooLong >= 0) return false;
    for (int32_t c = 0; c < nC; ++c) rowOffs[size_t(c) + 1] += rowOffs[c];
    FvkRawVec<int32_t>& col = out.col;          // every entry / offset below is written by the per-row loop
    FvkRawVec<uint8_t>&ownOff = out.ownOff, &neiOff = out.neiOff, &diagOff = out.diagOff;
    col.resize(size_t(rowOffs[nC]));
    ownOff.resize(size_t(nI)); neiOff.resize(size_t(nI)); diagOff.resize(size_t(nC));
#pragma omp parallel
    {
        std::vector<int32_t> lower, upper;
#pragma omp for schedule(static)
        for (int32_t c = 0; c < nC; ++c)
        {
            lower.clear(); upper.clear();
            for (int32_t e = sth.seg[c]; e < sth.seg[size_t(c) + 1]; ++e)
            {
                const int32_t f = sth.ent[e] >> 1;
                if (f >= nI) continue;
                ((sth.ent[e] & 1) ? lower : upper).push_back(f);
            }
            if (fkey)
            {
                auto byKey = [fkey](int32_t a, int32_t b) { return fkey[a] < fkey[b]; };
                std::sort(lower.begin(), lower.end(), byKey);
                std::sort(upper.begin(), upper.end(), byKey);
            }
            else
            {
                std::sort(lower.begin(), lower.end());
                std::sort(upper.begin(), upper.end());
            }
            const size_t r0 = size_t(rowOffs[c]);
            int32_t k = 0;
            for (int32_t f : lower) { neiOff[f] = uint8_t(k); col[r0 + k] = own[f]; ++k; }
            diagOff[c] = uint8_t(k);
            col[r0 + k] = c;
            ++k;
            for (int32_t f : upper) { ownOff[f] = uint8_t(k); col[r0 + k] = nei[f]; ++k; }
        }
    }
    // rows in stencil order? (k_assemble_fast / k_rAU_HbyA_rows derive slots from stencil positions)
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int32_t c = 0; c < nC; ++c)
    {
        int32_t k = 0;
        for (int32_t e = sth.seg[c]; e < sth.seg[size_t(c) + 1]; ++e, ++k)
        {
            const int32_t f = sth.ent[e] >> 1;
            if (f >= nI) break;
            const bool side = sth.ent[e] & 1;
            if (side ? (neiOff[f] != k || k >= diagOff[c]) : (ownOff[f] != k + 1 || k < diagOff[c])) bad |= 1;
        }
    }
    out.rowsInStencilOrder = !bad;
    return true;
}

// diagnostics (host only): the SparsityPattern exactly as fvk_mesh_create builds it. result[4] = {rows, nnz, every row in stencil
// order (0/1), rows that are NOT [lower columns | own id | upper columns] with the diagonal where diagOffset says (must be 0)}
extern "C" int fvk_sparsity_selftest(const fvk_mesh_desc* d, int64_t* result /* [4] */)
{
    if (!d || !result) return fvk_fail(FVK_EINVAL, "fvk_sparsity_selftest: null");
    FvkStencilHost st;
    fvk_build_stencil(d, st);
    FvkSparsityHost sp;
    if (!fvk_build_sparsity(d, st, sp)) return fvk_fail(FVK_EUNSUPPORTED, "fvk_sparsity_selftest: cell %d has > 255 row entries", sp.tooLongCell);
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    int64_t badRows = 0;
    for (int32_t c = 0; c < nC; ++c)
        if (sp.col[size_t(sp.rowOffs[c]) + sp.diagOff[c]] != c) ++badRows;
    for (int32_t f = 0; f < nI; ++f)
        if (sp.col[size_t(sp.rowOffs[d->faceOwner[f]]) + sp.ownOff[f]] != d->faceNeighbour[f]
            || sp.col[size_t(sp.rowOffs[d->faceNeighbour[f]]) + sp.neiOff[f]] != d->faceOwner[f])
            ++badRows;
    if (!d->faceOrder)
    { // single-domain mesh: the reference's own three serial passes over the faces (sparsityPattern.cpp:21-143) must give the same
      // column array as the row-parallel builder
        std::vector<int64_t> ro(size_t(nC) + 1, 0);
        for (int32_t c = 0; c < nC; ++c) ro[size_t(c) + 1] = 1;
        for (int32_t f = 0; f < nI; ++f) { ++ro[size_t(d->faceOwner[f]) + 1]; ++ro[size_t(d->faceNeighbour[f]) + 1]; }
        for (int32_t c = 0; c < nC; ++c) ro[size_t(c) + 1] += ro[c];
        std::vector<int32_t> col, cnt;
        col.resize(size_t(ro[nC])); cnt.assign(size_t(nC), 0);
        for (int32_t f = 0; f < nI; ++f) { const int32_t n = d->faceNeighbour[f]; col[size_t(ro[n]) + cnt[n]++] = d->faceOwner[f]; }
        for (int32_t c = 0; c < nC; ++c) col[size_t(ro[c]) + cnt[c]++] = c;
        for (int32_t f = 0; f < nI; ++f) { const int32_t o = d->faceOwner[f]; col[size_t(ro[o]) + cnt[o]++] = d->faceNeighbour[f]; }
        if (ro[nC] != sp.rowOffs[nC]) ++badRows;
        else
            for (int32_t c = 0; c < nC; ++c)
            {
                bool same = ro[c] == sp.rowOffs[c];
                for (int64_t k = ro[c]; same && k < ro[size_t(c) + 1]; ++k) same = col[size_t(k)] == sp.col[size_t(k)];
                if (!same) ++badRows;
            }
    }
    result[0] = nC; result[1] = sp.rowOffs[nC]; result[2] = sp.rowsInStencilOrder ? 1 : 0; result[3] = badRows;
    return FVK_OK;
}

// diagnostics entry (host only, no device needed): build stencil + brick plan for a mesh description and replay it
extern "C" int fvk_brick_plan_selftest(const fvk_mesh_desc* d, int32_t* info /* [8] */, int64_t* badCells)
{
    if (!d || !info || !badCells) return fvk_fail(FVK_EINVAL, "fvk_brick_plan_selftest: null");
    FvkStencilHost st;
    fvk_build_stencil(d, st);
    FvkBrickPlanHost bp;
    const char* why = "";
    std::memset(info, 0, sizeof(int32_t) * 8);
    *badCells = -1;
    if (!fvk_build_brick_plan(d, st, bp, &why)) return fvk_fail(FVK_EUNSUPPORTED, "brick plan: %s", why);
    info[0] = int32_t(bp.hdr.size());
    info[1] = bp.dims[0]; info[2] = bp.dims[1]; info[3] = bp.dims[2];
    info[4] = bp.brick[0]; info[5] = bp.brick[1]; info[6] = bp.brick[2];
    info[7] = bp.maxSlots;
    *badCells = fvk_verify_brick_plan(d, st, bp);
    return FVK_OK;
}

// diagnostics (host only): did the plan prove the affine topology? info[5] = {affine, tUp x, y, z, irregular cells}
extern "C" int fvk_brick_plan_affine_info(const fvk_mesh_desc* d, int32_t* info /* [5] */)
{
    if (!d || !info) return fvk_fail(FVK_EINVAL, "fvk_brick_plan_affine_info: null");
    FvkStencilHost st;
    fvk_build_stencil(d, st);
    FvkBrickPlanHost bp;
    const char* why = "";
    if (!fvk_build_brick_plan(d, st, bp, &why)) return fvk_fail(FVK_EUNSUPPORTED, "brick plan: %s", why);
    info[0] = bp.geom.affine; info[1] = bp.geom.tUp[0]; info[2] = bp.geom.tUp[1]; info[3] = bp.geom.tUp[2];
    info[4] = int32_t(bp.irrCells.size());
    return FVK_OK;
}

// diagnostics (host only): the assumption behind the structured SpMV (fvk_spmv_structured / fvk_solver_attach_mesh). With
// the SparsityPattern built as in fvk_mesh_create (row = [lower faces ascending | diag | upper faces ascending]) and a
// proven affine topology, every REGULAR row must be exactly [c-nx*ny, c-nx, c-1, c, c+1, c+nx, c+nx*ny]. result[0] = 1 when
// the topology is affine and all regular rows check, 0 when not affine, -1 when a row violates it; result[1] = rows checked.
extern "C" int fvk_brick_plan_structured_rows(const fvk_mesh_desc* d, int64_t* result /* [2] */)
{
    if (!d || !result) return fvk_fail(FVK_EINVAL, "fvk_brick_plan_structured_rows: null");
    FvkStencilHost st;
    fvk_build_stencil(d, st);
    FvkBrickPlanHost bp;
    const char* why = "";
    result[0] = 0; result[1] = 0;
    if (!fvk_build_brick_plan(d, st, bp, &why) || !bp.geom.affine) return FVK_OK;
    const int32_t nC = d->nCells, nI = d->nInternalFaces;
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    std::vector<int64_t> rowOffs(size_t(nC) + 1, 0);
    for (int32_t c = 0; c < nC; ++c) rowOffs[size_t(c) + 1] = 1;
    for (int32_t f = 0; f < nI; ++f) { ++rowOffs[size_t(own[f]) + 1]; ++rowOffs[size_t(nei[f]) + 1]; }
    for (int32_t c = 0; c < nC; ++c) rowOffs[size_t(c) + 1] += rowOffs[c];
    std::vector<int32_t> col, cnt;
    col.resize(size_t(rowOffs[nC]));
    cnt.assign(size_t(nC), 0);
    for (int32_t f = 0; f < nI; ++f) col[size_t(rowOffs[nei[f]]) + cnt[nei[f]]++] = own[f];
    for (int32_t c = 0; c < nC; ++c) col[size_t(rowOffs[c]) + cnt[c]++] = c;
    for (int32_t f = 0; f < nI; ++f) col[size_t(rowOffs[own[f]]) + cnt[own[f]]++] = nei[f];
    const int64_t nx = bp.geom.dims[0], ny = bp.geom.dims[1], nz = bp.geom.dims[2], nxy = nx * ny;
    int64_t checked = 0;
    bool ok = true;
    for (int64_t c = 0; c < bp.geom.nOwned && ok; ++c)
    {
        const int64_t i = c % nx, j = (c / nx) % ny, k = c / nxy;
        if (!(i > 0 && i < nx - 1 && j > 0 && j < ny - 1 && k > 0 && k < nz - 1)) continue;
        const int64_t want[7] = {c - nxy, c - nx, c - 1, c, c + 1, c + nx, c + nxy};
        if (rowOffs[c + 1] - rowOffs[c] != 7) { ok = false; break; }
        for (int t = 0; t < 7; ++t)
            if (col[size_t(rowOffs[c]) + t] != want[t]) ok = false;
        ++checked;
    }
    result[0] = ok ? 1 : -1;
    result[1] = checked;
    return FVK_OK;
}
