// TEST INFRASTRUCTURE -- block-hex mesh for the CPU oracle (part of oracle/libfvo.so). NOT product code.
//
// Stand-in for OpenFOAM `blockMesh` + FoamAdapter::readOpenFOAMMesh (reference src/datastructures/meshAdapter.cpp:12-45,
// 59-136), which need OpenFOAM: one hex block (nx ny nz) of a box lx x ly x lz, simpleGrading 1, in the ordering of the
// polyMesh fixtures the reference commits with OpenFOAM's primitiveMesh geometry
// (primitiveMeshFaceCentresAndAreas.C: triangle fan about the vertex average; primitiveMeshCellCentresAndVols.C: pyramids
// about the face-centre average, owner faces in ascending id, then neighbour faces). Written cell-centric (every cell
// gathers its faces; geometry of a face is recomputed from the point formula where needed), i.e. independently of the
// product's generator (foamadapter_b200/csrc/fvk_blockmesh.cpp); tests/test_blockmesh.py pins both against the reference's
// fixtures and against each other, bit for bit. Lets bench.py's CPU arms build their mesh without loading libfvk.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace
{
struct Block
{
    int64_t nx, ny, nz;
    double lx, ly, lz;
    void point(int64_t i, int64_t j, int64_t k, double* p) const
    {
        p[0] = lx * (double(i) / double(nx));
        p[1] = ly * (double(j) / double(ny));
        p[2] = lz * (double(k) / double(nz));
    }
};
struct Quad
{
    double p[4][3];
};
// face kinds: 0..5 = boundary side x-min, x-max, y-min, y-max, z-min, z-max of cell (i,j,k); 6, 7, 8 = the internal face
// the cell owns towards +x, +y, +z. Vertex order as blockMesh emits it (pinned by the fixtures' `faces` files).
Quad quad(const Block& b, int kind, int64_t i, int64_t j, int64_t k)
{
    static const int V[9][4][3] = {
        {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}}, // x-min
        {{1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}}, // x-max
        {{0, 0, 0}, {1, 0, 0}, {1, 0, 1}, {0, 0, 1}}, // y-min
        {{0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}}, // y-max
        {{0, 0, 0}, {0, 1, 0}, {1, 1, 0}, {1, 0, 0}}, // z-min
        {{0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}}, // z-max
        {{1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}}, // internal +x
        {{0, 1, 0}, {0, 1, 1}, {1, 1, 1}, {1, 1, 0}}, // internal +y
        {{0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}}, // internal +z
    };
    Quad q;
    for (int v = 0; v < 4; ++v) b.point(i + V[kind][v][0], j + V[kind][v][1], k + V[kind][v][2], q.p[v]);
    return q;
}
// primitiveMeshFaceCentresAndAreas.C for a 4-point face
void faceGeometry(const Quad& q, double* cf, double* sf)
{
    double fc[3];
    for (int d = 0; d < 3; ++d) fc[d] = (((q.p[0][d] + q.p[1][d]) + q.p[2][d]) + q.p[3][d]) / 4;
    double sumN[3] = {0, 0, 0}, sumAc[3] = {0, 0, 0}, sumA = 0.0;
    for (int v = 0; v < 4; ++v)
    {
        const double* cur = q.p[v];
        const double* nxt = q.p[(v + 1) % 4];
        double e1[3], e2[3], c[3];
        for (int d = 0; d < 3; ++d) { c[d] = cur[d] + nxt[d] + fc[d]; e1[d] = nxt[d] - cur[d]; e2[d] = fc[d] - cur[d]; }
        const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double a = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        sumA += a;
        for (int d = 0; d < 3; ++d) { sumN[d] += n[d]; sumAc[d] += a * c[d]; }
    }
    for (int d = 0; d < 3; ++d)
    {
        cf[d] = sumA < 1e-150 ? fc[d] : (1.0 / 3.0) * sumAc[d] / sumA;
        sf[d] = sumA < 1e-150 ? 0.0 : 0.5 * sumN[d];
    }
}
} // namespace

extern "C" {
// sizes of the NeoN view: out[0] = nCells, [1] = nInternalFaces, [2] = nBoundaryFaces (empty patches dropped), [3] = kept patches
void fvo_blockmesh_sizes(int32_t nx, int32_t ny, int32_t nz, int32_t nPatches, const int32_t* patchNSides, const int32_t* patchSides,
                         const int32_t* patchIsEmpty, int64_t* out)
{
    const int64_t X = nx, Y = ny, Z = nz;
    const int64_t sideSize[6] = {Y * Z, Y * Z, X * Z, X * Z, X * Y, X * Y};
    int64_t nB = 0, kept = 0;
    for (int p = 0, s = 0; p < nPatches; ++p)
    {
        for (int q = 0; q < patchNSides[p]; ++q, ++s)
            if (!patchIsEmpty[p]) nB += sideSize[patchSides[s]];
        kept += !patchIsEmpty[p];
    }
    out[0] = X * Y * Z;
    out[1] = (X - 1) * Y * Z + X * (Y - 1) * Z + X * Y * (Z - 1);
    out[2] = nB;
    out[3] = kept;
}

// fills the arrays oracle.cpu.Mesh takes (caller-allocated with the sizes above): owner [nI+nB], neighbour [nI],
// faceCells [nB], V [nC], C [nC*3], Sf / Cf [(nI+nB)*3], magSf [nI+nB], bSf [nB*3], bDeltaCoeffs / bWeights [nB],
// patchOffsets [kept+1]
void fvo_blockmesh(int32_t nx_, int32_t ny_, int32_t nz_, double lx, double ly, double lz, int32_t nPatches,
                   const int32_t* patchNSides, const int32_t* patchSides, const int32_t* patchIsEmpty, int32_t* owner,
                   int32_t* neighbour, int32_t* faceCells, double* V, double* C, double* Sf, double* Cf, double* magSf, double* bSf,
                   double* bDeltaCoeffs, double* bWeights, int32_t* patchOffsets)
{
    const Block blk {nx_, ny_, nz_, lx, ly, lz};
    const int64_t nx = nx_, ny = ny_, nz = nz_, nC = nx * ny * nz;
    const int64_t nI = (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1);
    auto cellId = [=](int64_t i, int64_t j, int64_t k) { return i + nx * (j + ny * k); };
    // first internal face of cell (i,j,k): faces are sorted by owner, a cell owns its +x, +y, +z faces in that order
    auto faceStart = [=](int64_t i, int64_t j, int64_t k) {
        const int64_t a = k < nz - 1, rowFull = nx * (1 + a) + nx - 1;
        const int64_t plane = (ny - 1) * (2 * nx + nx - 1) + nx + nx - 1; // a plane below the top one
        return k * plane + j * rowFull + i * ((j < ny - 1) + a) + i;
    };
    // order of the block sides in the (full, empty patches included) boundary listing, and where each KEPT side starts
    int sidePos[6] = {0, 0, 0, 0, 0, 0};
    int64_t sideStart[6] = {-1, -1, -1, -1, -1, -1};
    const int64_t sideSize[6] = {ny * nz, ny * nz, nx * nz, nx * nz, nx * ny, nx * ny};
    {
        int64_t b = 0;
        int kept = 0;
        patchOffsets[0] = 0;
        for (int p = 0, s = 0; p < nPatches; ++p)
        {
            for (int q = 0; q < patchNSides[p]; ++q, ++s)
            {
                const int side = patchSides[s];
                sidePos[side] = s;
                if (!patchIsEmpty[p]) { sideStart[side] = b; b += sideSize[side]; }
            }
            if (!patchIsEmpty[p]) patchOffsets[++kept] = int32_t(b);
        }
    }
    // position of cell (i,j,k)'s face inside its side (blockMesh loop nests: x sides j fastest then k; y sides k fastest
    // then i; z sides j fastest then i)
    auto sideIndex = [=](int side, int64_t i, int64_t j, int64_t k) {
        return side < 2 ? j + ny * k : (side < 4 ? k + nz * i : j + ny * i);
    };
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nz; ++k)
        for (int64_t j = 0; j < ny; ++j)
            for (int64_t i = 0; i < nx; ++i)
            {
                const int64_t c = cellId(i, j, k);
                // ---- the cell's faces in OpenFOAM's accumulation order: owned internal (+x +y +z), owned boundary (listing
                //      order), then the faces it is the neighbour of (owners c - nx ny, c - nx, c - 1 ascending)
                struct F { double cf[3], sf[3]; bool own; };
                F fl[12];
                int n = 0;
                auto add = [&](int kind, int64_t fi, int64_t fj, int64_t fk, bool own) {
                    faceGeometry(quad(blk, kind, fi, fj, fk), fl[n].cf, fl[n].sf);
                    fl[n].own = own;
                    return n++;
                };
                int64_t f = faceStart(i, j, k);
                const bool has[3] = {i < nx - 1, j < ny - 1, k < nz - 1};
                const int64_t nb[3] = {c + 1, c + nx, c + nx * ny};
                for (int dir = 0; dir < 3; ++dir)
                    if (has[dir])
                    {
                        const int q = add(6 + dir, i, j, k, true);
                        owner[f] = int32_t(c); neighbour[f] = int32_t(nb[dir]);
                        for (int d = 0; d < 3; ++d) { Cf[3 * f + d] = fl[q].cf[d]; Sf[3 * f + d] = fl[q].sf[d]; }
                        magSf[f] = std::sqrt(fl[q].sf[0] * fl[q].sf[0] + fl[q].sf[1] * fl[q].sf[1] + fl[q].sf[2] * fl[q].sf[2]);
                        ++f;
                    }
                int sides[6], ns = 0;
                if (i == 0) sides[ns++] = 0;
                if (i == nx - 1) sides[ns++] = 1;
                if (j == 0) sides[ns++] = 2;
                if (j == ny - 1) sides[ns++] = 3;
                if (k == 0) sides[ns++] = 4;
                if (k == nz - 1) sides[ns++] = 5;
                int bsides[6], nbs = 0;
                for (int s = 0; s < ns; ++s) bsides[nbs++] = sides[s];
                std::sort(bsides, bsides + nbs, [&](int a, int b2) { return sidePos[a] < sidePos[b2]; });
                int bq[6];
                for (int s = 0; s < nbs; ++s) bq[s] = add(bsides[s], i, j, k, true);
                if (k > 0) add(8, i, j, k - 1, false);
                if (j > 0) add(7, i, j - 1, k, false);
                if (i > 0) add(6, i - 1, j, k, false);
                // ---- primitiveMeshCellCentresAndVols.C
                double cEst[3] = {0, 0, 0};
                for (int q = 0; q < n; ++q)
                    for (int d = 0; d < 3; ++d) cEst[d] += fl[q].cf[d];
                for (int d = 0; d < 3; ++d) cEst[d] /= n;
                double cc[3] = {0, 0, 0}, vol = 0.0;
                for (int q = 0; q < n; ++q)
                {
                    double pyr3 = 0.0;
                    for (int d = 0; d < 3; ++d) pyr3 += fl[q].sf[d] * (fl[q].own ? fl[q].cf[d] - cEst[d] : cEst[d] - fl[q].cf[d]);
                    for (int d = 0; d < 3; ++d) cc[d] += pyr3 * ((3.0 / 4.0) * fl[q].cf[d] + (1.0 / 4.0) * cEst[d]);
                    vol += pyr3;
                }
                for (int d = 0; d < 3; ++d) C[3 * c + d] = cc[d] / vol;
                V[c] = vol * (1.0 / 3.0);
                // ---- kept boundary faces of this cell (fvPatch view: Sf, faceCells, weights 1, deltaCoeffs 1/|Cf - C|)
                for (int s = 0; s < nbs; ++s)
                {
                    const int side = bsides[s];
                    if (sideStart[side] < 0) continue; // empty patch: no faces in the NeoN mesh
                    const int64_t b = sideStart[side] + sideIndex(side, i, j, k), fb = nI + b;
                    const F& q = fl[bq[s]];
                    double d2 = 0.0;
                    for (int d = 0; d < 3; ++d)
                    {
                        Cf[3 * fb + d] = q.cf[d]; Sf[3 * fb + d] = q.sf[d]; bSf[3 * b + d] = q.sf[d];
                        const double dl = q.cf[d] - C[3 * c + d];
                        d2 += dl * dl;
                    }
                    magSf[fb] = std::sqrt(q.sf[0] * q.sf[0] + q.sf[1] * q.sf[1] + q.sf[2] * q.sf[2]);
                    owner[fb] = int32_t(c); faceCells[b] = int32_t(c);
                    bWeights[b] = 1.0;
                    bDeltaCoeffs[b] = 1.0 / std::sqrt(d2);
                }
            }
}
}
