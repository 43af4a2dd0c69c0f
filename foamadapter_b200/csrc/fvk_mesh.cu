// Mesh handle: upload, geometry scheme (device kernels), cell->face stencil, sparsity pattern.
//
// Reference behaviour restated here (not copied): NeoN::UnstructuredMesh / BoundaryMesh layout
// (src/NeoN/include/NeoN/mesh/unstructured/*.hpp), BasicGeometryScheme
// (src/NeoN/src/finiteVolume/cellCentred/stencil/basicGeometryScheme.cpp:15-136),
// CellToFaceStencil (stencil/cellToFaceStencil.cpp:14-96), SparsityPattern
// (src/NeoN/src/linearAlgebra/sparsityPattern.cpp:21-143; like the reference the pattern is
// assembled once per mesh on the host and uploaded).
#include "fvk_device.cuh"
#include "fvk_brickplan.hpp"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------
// errors, device management
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int g_variant = 0;

int fvk_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" int fvk_version(void) { return FVK_VERSION; }
extern "C" const char* fvk_last_error(void) { return g_err; }
extern "C" int fvk_set_variant(int v)
{
    g_variant = v;
    return FVK_OK;
}
int fvk_variant() { return g_variant; }

// kernel configuration override {kernel (1 brick, 3 affine), threads per block, resident blocks aimed at}; {0,0,0}: defaults
static int g_brickCfg[3] = {0, 0, 0};
static bool g_brickCfgEnv = false;
extern "C" int fvk_mesh_set_tile_phase(fvk_mesh* m, int phase)
{
    if (!m || phase < 0 || phase > 2) return fvk_fail(FVK_EINVAL, "fvk_mesh_set_tile_phase: bad argument");
    m->tilePhase = phase;
    return FVK_OK;
}
// run-time switch for A/B measurements: 1 = keep the generic brick kernel for the whole mesh
static int g_noAffine = -1;
extern "C" int fvk_set_affine(int enabled)
{
    g_noAffine = enabled ? 0 : 1;
    return FVK_OK;
}
bool fvk_no_affine()
{
    if (g_noAffine < 0)
    {
        const char* e = std::getenv("FVK_NO_AFFINE");
        g_noAffine = (e && *e == '1') ? 1 : 0;
    }
    return g_noAffine == 1;
}
extern "C" int fvk_set_brick_config(int kernel, int threads, int minBlocks)
{
    g_brickCfg[0] = kernel; g_brickCfg[1] = threads; g_brickCfg[2] = minBlocks;
    g_brickCfgEnv = true; // an explicit call wins over the environment
    return FVK_OK;
}
bool fvk_brick_config(int cfg[3])
{
    if (!g_brickCfgEnv)
    {
        g_brickCfgEnv = true;
        if (const char* e = std::getenv("FVK_BRICK_CFG"))
            if (std::sscanf(e, "%d,%d,%d", &g_brickCfg[0], &g_brickCfg[1], &g_brickCfg[2]) != 3)
                g_brickCfg[0] = g_brickCfg[1] = g_brickCfg[2] = 0;
    }
    cfg[0] = g_brickCfg[0]; cfg[1] = g_brickCfg[1]; cfg[2] = g_brickCfg[2];
    return cfg[0] > 0;
}

int fvk_sm_count()
{
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0)
    {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

extern "C" int fvk_device_count(int* n)
{
    if (!n) return fvk_fail(FVK_EINVAL, "fvk_device_count: null");
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess)
    {
        *n = 0;
        return fvk_fail(FVK_ENODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return FVK_OK;
}
extern "C" int fvk_set_device(int device)
{
    FVK_CUDA(cudaSetDevice(device));
    return FVK_OK;
}
extern "C" int fvk_malloc(void** dptr, size_t bytes)
{
    if (!dptr) return fvk_fail(FVK_EINVAL, "fvk_malloc: null");
    *dptr = nullptr;
    if (bytes == 0) return FVK_OK;
    FVK_CUDA(cudaMalloc(dptr, bytes));
    return FVK_OK;
}
extern "C" int fvk_free(void* dptr)
{
    if (dptr) FVK_CUDA(cudaFree(dptr));
    return FVK_OK;
}
extern "C" int fvk_malloc_host(void** hptr, size_t bytes)
{
    if (!hptr) return fvk_fail(FVK_EINVAL, "fvk_malloc_host: null");
    *hptr = nullptr;
    if (bytes == 0) return FVK_OK;
    FVK_CUDA(cudaMallocHost(hptr, bytes));
    return FVK_OK;
}
extern "C" int fvk_free_host(void* hptr)
{
    if (hptr) FVK_CUDA(cudaFreeHost(hptr));
    return FVK_OK;
}
extern "C" int fvk_memcpy_h2d(void* dst, const void* src, size_t bytes, fvk_stream s)
{
    if (bytes) FVK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, fvk_cu(s)));
    return FVK_OK;
}
extern "C" int fvk_memcpy_d2h(void* dst, const void* src, size_t bytes, fvk_stream s)
{
    if (bytes) FVK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, fvk_cu(s)));
    return FVK_OK;
}
extern "C" int fvk_memcpy_d2d(void* dst, const void* src, size_t bytes, fvk_stream s)
{
    if (bytes) FVK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, fvk_cu(s)));
    return FVK_OK;
}
extern "C" int fvk_memset(void* dst, int byte, size_t bytes, fvk_stream s)
{
    if (bytes) FVK_CUDA(cudaMemsetAsync(dst, byte, bytes, fvk_cu(s)));
    return FVK_OK;
}
extern "C" int fvk_stream_create(fvk_stream* s)
{
    if (!s) return fvk_fail(FVK_EINVAL, "fvk_stream_create: null");
    cudaStream_t st;
    FVK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *s = st;
    return FVK_OK;
}
extern "C" int fvk_stream_destroy(fvk_stream s)
{
    if (s) FVK_CUDA(cudaStreamDestroy(fvk_cu(s)));
    return FVK_OK;
}
extern "C" int fvk_stream_sync(fvk_stream s)
{
    FVK_CUDA(cudaStreamSynchronize(fvk_cu(s)));
    return FVK_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry scheme kernels (K31-K34)
// ------------------------------------------------------------------------------------------------
namespace
{
constexpr double ROOTVSMALL = 1e-18; // src/NeoN/include/NeoN/core/primitives/scalar.hpp

__global__ void __launch_bounds__(256)
k_geometry_scheme(int nI, int nF, const int* __restrict__ owner, const int* __restrict__ neighbour,
                  const double* __restrict__ C, const double* __restrict__ Cf,
                  const double* __restrict__ Sf, const double* __restrict__ magSf,
                  double* __restrict__ w, double* __restrict__ dc, double* __restrict__ nodc)
{
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x)
    {
        const Vec3d sf = ld3(Sf, f), cf = ld3(Cf, f);
        const Vec3d cp = ld3(C, owner[f]);
        Vec3d d; // cell-to-cell (internal) or cell-to-face (boundary) distance
        if (f < nI)
        {
            const Vec3d cn = ld3(C, neighbour[f]);
            // basicGeometryScheme.cpp:27-43
            const double sfdOwn = fabs(sf.x * (cf.x - cp.x) + sf.y * (cf.y - cp.y) + sf.z * (cf.z - cp.z));
            const double sfdNei = fabs(sf.x * (cn.x - cf.x) + sf.y * (cn.y - cf.y) + sf.z * (cn.z - cf.z));
            w[f] = (fabs(sfdOwn + sfdNei) > ROOTVSMALL) ? sfdNei / (sfdOwn + sfdNei) : 0.5;
            d = Vec3d {cn.x - cp.x, cn.y - cp.y, cn.z - cp.z};
        }
        else
        {
            w[f] = 1.0; // :45-53
            d = Vec3d {cf.x - cp.x, cf.y - cp.y, cf.z - cp.z};
        }
        const double magD = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
        dc[f] = 1.0 / magD; // :68-88
        // :108-135  faceNormal = (1/|Sf|) * Sf ; orthoDist = faceNormal & d
        const double inv = 1 / magSf[f];
        const double ortho = (inv * sf.x) * d.x + (inv * sf.y) * d.y + (inv * sf.z) * d.z;
        nodc[f] = 1.0 / fmax(ortho, 0.05 * magD);
    }
}

// Host -> device copy of a (pageable) setup array. cudaMemcpy from pageable memory stages through the driver's own bounce buffer on
// one thread (~7 GB/s on this host: 0.5 s for the 3.3 GB of a 256^3 mesh); large arrays instead go through two pinned 32 MB
// buffers filled by all host threads while the previous chunk is on the wire.
struct Stager
{
    static constexpr size_t CHUNK = size_t(32) << 20;
    unsigned char* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    bool ok = false, tried = false;
    int device = -1;
    std::mutex mtx;
    bool init()
    {
        int dev = -1;
        cudaGetDevice(&dev);
        if (tried) return ok && dev == device; // buffers, stream and events belong to the device of the first mesh
        tried = true;
        device = dev;
        ok = cudaMallocHost(reinterpret_cast<void**>(&buf[0]), CHUNK) == cudaSuccess && cudaMallocHost(reinterpret_cast<void**>(&buf[1]), CHUNK) == cudaSuccess
             && cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess
             && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        if (!ok) cudaGetLastError();
        return ok;
    }
    cudaError_t copy(void* dst, const void* src, size_t bytes)
    {
        std::lock_guard<std::mutex> lock(mtx);
        const unsigned char* s = static_cast<const unsigned char*>(src);
        unsigned char* d = static_cast<unsigned char*>(dst);
        int b = 0;
        for (size_t off = 0; off < bytes; off += CHUNK, b ^= 1)
        {
            const size_t n = bytes - off < CHUNK ? bytes - off : CHUNK;
            cudaError_t e = cudaEventSynchronize(ev[b]); // the copy that last used this buffer is done
            if (e != cudaSuccess) return e;
            const int64_t nBlocks = int64_t((n + 262143) / 262144);
#pragma omp parallel for schedule(static)
            for (int64_t q = 0; q < nBlocks; ++q)
            {
                const size_t o = size_t(q) * 262144, len = n - o < 262144 ? n - o : 262144;
                std::memcpy(buf[b] + o, s + off + o, len);
            }
            e = cudaMemcpyAsync(d + off, buf[b], n, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(ev[b], st);
            if (e != cudaSuccess) return e;
        }
        return cudaStreamSynchronize(st);
    }
};
Stager& stager()
{
    static Stager s;
    return s;
}

template <class T>
int upload(T** dst, const T* src, size_t n)
{
    *dst = nullptr;
    if (n == 0) return FVK_OK;
    FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), n * sizeof(T) + 16)); // 16 bytes of slack: bulk copies over-fetch
    static const bool noStage = [] { const char* e = std::getenv("FVK_NO_STAGED_UPLOAD"); return e && *e == '1'; }();
    if (!noStage && n * sizeof(T) >= (size_t(8) << 20) && stager().init())
        FVK_CUDA(stager().copy(*dst, src, n * sizeof(T)));
    else
        FVK_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return FVK_OK;
}
#define UP(field, src, n)                                                                          \
    do                                                                                             \
    {                                                                                              \
        int rc_ = upload(&m->field, src, size_t(n));                                               \
        if (rc_) { fvk_mesh_destroy(m); return rc_; }                                              \
    } while (0)
} // namespace

extern "C" int fvk_mesh_destroy(fvk_mesh* m)
{
    if (!m) return FVK_OK;
    void* ptrs[] = {m->V, m->C, m->Sf, m->Cf, m->magSf, m->owner, m->neighbour, m->faceCells, m->bCf,
                    m->bCn, m->bSf, m->bMagSf, m->bNf, m->bDelta, m->bWeights, m->bDeltaCoeffs,
                    m->weights, m->deltaCoeffs, m->nonOrthDeltaCoeffs, m->stencilSeg, m->stencilVal,
                    m->gatherEnt, m->gatherPlan, m->rowOffs, m->colIdxs, m->ownerOffset, m->neighbourOffset,
                    m->diagOffset, m->ownStart, m->lowSeg, m->lowFace, m->lowOwner, m->bndCell,
                    m->bndSeg, m->bndFace, m->hasBnd, m->tp.hdr, m->tp.blob, m->bp.hdr, m->bp.rec, m->bp.codes,
                    m->bp.xFace, m->bp.xOwner, m->bp.xNei, m->bp.bFace, m->bp.bCell, m->bp.recF, m->bp.codes4, m->bp.tileInfo, m->bp.irrCells,
                    m->dicColor, m->dicCells};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete m;
    return FVK_OK;
}

namespace
{
// FVK_SETUP_TIMING=1: wall-clock of the phases of fvk_mesh_create to stderr
struct SetupTimer
{
    bool on;
    std::chrono::steady_clock::time_point t0;
    SetupTimer() : on([] { const char* e = std::getenv("FVK_SETUP_TIMING"); return e && *e == '1'; }()), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[fvk setup] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
bool experiment_plans()
{
    static const bool on = [] { const char* e = std::getenv("FVK_EXPERIMENT_PLANS"); return e && *e == '1'; }();
    return on;
}
} // namespace

extern "C" int fvk_mesh_create(const fvk_mesh_desc* d, fvk_mesh** out)
{
    if (!d || !out) return fvk_fail(FVK_EINVAL, "fvk_mesh_create: null argument");
    *out = nullptr;
    SetupTimer tm;
    const int32_t nC = d->nCells, nI = d->nInternalFaces, nB = d->nBoundaryFaces;
    if (nC <= 0 || nI < 0 || nB < 0 || d->nPatches < 0 || d->nPatches > FVK_MAX_PATCHES)
        return fvk_fail(FVK_EINVAL, "fvk_mesh_create: bad sizes (nCells=%d nI=%d nB=%d nPatches=%d)",
                        nC, nI, nB, d->nPatches);
    if (!d->cellVolumes || !d->cellCentres || !d->faceAreas || !d->faceCentres || !d->magFaceAreas
        || !d->faceOwner || (nI && !d->faceNeighbour) || (nB && (!d->faceCells || !d->patchOffsets)))
        return fvk_fail(FVK_EINVAL, "fvk_mesh_create: missing array");
    const int64_t nF = int64_t(nI) + nB;
    if (nF >= (int64_t(1) << 30)) return fvk_fail(FVK_EUNSUPPORTED, "fvk_mesh_create: > 2^30 faces");
    {
        int n = 0;
        int rc = fvk_device_count(&n);
        if (rc || n == 0) return fvk_fail(FVK_ENODEVICE, "fvk_mesh_create: no CUDA device (no CPU fallback)");
    }
    {
        int32_t badFace = -1, badB = -1;
#pragma omp parallel for schedule(static) reduction(max : badFace)
        for (int32_t f = 0; f < nI; ++f)
        {
            const int32_t o = d->faceOwner[f], n = d->faceNeighbour[f];
            if (o < 0 || o >= nC || n < 0 || n >= nC || o == n) badFace = std::max(badFace, f);
        }
        if (badFace >= 0) return fvk_fail(FVK_EINVAL, "fvk_mesh_create: face %d has bad owner/neighbour", badFace);
#pragma omp parallel for schedule(static) reduction(max : badB)
        for (int32_t b = 0; b < nB; ++b)
            if (d->faceCells[b] < 0 || d->faceCells[b] >= nC) badB = std::max(badB, b);
        if (badB >= 0) return fvk_fail(FVK_EINVAL, "fvk_mesh_create: boundary face %d has bad faceCell", badB);
    }
    tm.lap("validate");

    fvk_mesh* m = new fvk_mesh;
    m->nCells = nC; m->nInternalFaces = nI; m->nBoundaryFaces = nB; m->nPatches = d->nPatches;
    m->nnz = int64_t(nC) + 2 * int64_t(nI);
    m->nOwned = (d->nOwnedCells > 0 && d->nOwnedCells <= nC) ? d->nOwnedCells : nC;
    cudaGetDevice(&m->device);
    for (int p = 0; p <= d->nPatches; ++p) m->patchOffsets[p] = nB ? d->patchOffsets[p] : 0;

    UP(V, d->cellVolumes, nC);
    UP(C, d->cellCentres, 3 * size_t(nC));
    UP(Sf, d->faceAreas, 3 * size_t(nF));
    UP(Cf, d->faceCentres, 3 * size_t(nF));
    UP(magSf, d->magFaceAreas, nF);
    {
        // faceOwner over all nF faces: boundary part = faceCells
        std::vector<int32_t> own(static_cast<size_t>(nF));
        std::memcpy(own.data(), d->faceOwner, sizeof(int32_t) * size_t(nI));
        for (int32_t b = 0; b < nB; ++b) own[size_t(nI) + b] = d->faceCells[b];
        UP(owner, own.data(), nF);
    }
    UP(neighbour, d->faceNeighbour, nI);
    UP(faceCells, d->faceCells, nB);
    if (d->bCf) UP(bCf, d->bCf, 3 * size_t(nB));
    if (d->bCn) UP(bCn, d->bCn, 3 * size_t(nB));
    if (d->bSf) UP(bSf, d->bSf, 3 * size_t(nB));
    if (d->bMagSf) UP(bMagSf, d->bMagSf, nB);
    if (d->bNf) UP(bNf, d->bNf, 3 * size_t(nB));
    if (d->bDelta) UP(bDelta, d->bDelta, 3 * size_t(nB));
    if (d->bWeights) UP(bWeights, d->bWeights, nB);
    if (d->bDeltaCoeffs) UP(bDeltaCoeffs, d->bDeltaCoeffs, nB);
    tm.lap("upload mesh arrays");

    // ---- cell->face stencil + gather plan: visiting faces in ascending id appends ascending ids
    const int32_t* own = d->faceOwner;
    const int32_t* nei = d->faceNeighbour;
    FvkStencilHost sth;
    fvk_build_stencil(d, sth, experiment_plans());
    tm.lap("cell->face stencil");
    {
        auto &seg = sth.seg; auto &val = sth.val; auto &ent = sth.ent; auto &plan = sth.plan;
        const size_t nEnt = ent.size();
        // ---- brick plan of k_gather_brick (default explicit-operator kernel when available)
        {
            FvkBrickPlanHost bph;
            const char* why = "";
            static const bool off = [] { const char* e = std::getenv("FVK_NO_BRICK"); return e && *e == '1'; }();
            if (!off && fvk_build_brick_plan(d, sth, bph, &why))
            {
                FvkBrickPlan& bp = m->bp;
                UP(bp.hdr, bph.hdr.data(), bph.hdr.size());
                UP(bp.rec, bph.rec.data(), bph.rec.size());
                UP(bp.codes, bph.codes.data(), bph.codes.size());
                UP(bp.xFace, bph.xFace.data(), bph.xFace.size());
                UP(bp.xOwner, bph.xOwner.data(), bph.xOwner.size());
                UP(bp.xNei, bph.xNei.data(), bph.xNei.size());
                UP(bp.bFace, bph.bFace.data(), bph.bFace.size());
                UP(bp.bCell, bph.bCell.data(), bph.bCell.size());
                UP(bp.recF, bph.recF.data(), bph.recF.size());
                UP(bp.codes4, bph.codes4.data(), bph.codes4.size());
                UP(bp.tileInfo, bph.tileInfo.data(), bph.tileInfo.size());
                bp.geom = bph.geom;
                if (!bph.irrCells.empty()) UP(bp.irrCells, bph.irrCells.data(), bph.irrCells.size());
                bp.nIrr = int32_t(bph.irrCells.size());
                bp.nTiles = int32_t(bph.hdr.size()); bp.maxSlots = bph.maxSlots; bp.maxCells = bph.maxCells;
            }
        }
        tm.lap("brick plan + upload");
        // ---- tile plan (see FvkTilePlan): consecutive owned cells, bounded slot / entry counts. Only the opt-in experiment
        // kernel (fvk_set_variant 6) reads it: built on request (FVK_EXPERIMENT_PLANS=1; the parity tests set it)
        if (experiment_plans())
        {
            bool sorted = nI > 0;
            for (int32_t f = 1; f < nI && sorted; ++f) sorted = own[f - 1] <= own[f];
            int tileCells = 256;
            if (const char* e = std::getenv("FVK_TILE_CELLS")) tileCells = std::atoi(e);
            if (tileCells < 32) tileCells = 32;
            if (tileCells > 2048) tileCells = 2048;
            const int32_t maxSlots = 12 * tileCells, maxEnt = 16 * tileCells; // uint16 codes: slot < 32768
            if (sorted)
            {
                const int32_t nOwned = m->nOwned;
                std::vector<int32_t> os(size_t(nC) + 1, 0);
                for (int32_t f = 0; f < nI; ++f) ++os[size_t(own[f]) + 1];
                for (int32_t c = 0; c < nC; ++c) os[size_t(c) + 1] += os[c];
                // greedy tiling; a cell's slot demand is bounded by its entry count
                std::vector<int32_t> tc(1, 0), cellTile(nC, -1);
                int32_t cells = 0, ents = 0;
                bool ok = true;
                for (int32_t c = 0; c < nOwned; ++c)
                {
                    const int32_t k = seg[size_t(c) + 1] - seg[c];
                    if (k > maxSlots) { ok = false; break; }
                    if (cells == tileCells || ents + k > maxSlots)
                    {
                        tc.push_back(c);
                        cells = 0; ents = 0;
                    }
                    cellTile[c] = int32_t(tc.size()) - 1;
                    ++cells; ents += k;
                }
                if (ok && nOwned > 0)
                {
                    tc.push_back(nOwned);
                    const int32_t nT = int32_t(tc.size()) - 1;
                    std::vector<FvkTileHdr> hdr(nT);
                    std::vector<unsigned char> blob;
                    std::vector<uint16_t> tseg, oseg, code, xCell, bCell;
                    std::vector<int32_t> xFace, xOwner, bFace;
                    FvkTilePlan& tp = m->tp;
                    for (int32_t t = 0; t < nT && ok; ++t)
                    {
                        FvkTileHdr& h = hdr[t];
                        h.c0 = tc[t]; h.nc = tc[size_t(t) + 1] - tc[t];
                        h.f0 = os[h.c0]; h.nf = os[h.c0 + h.nc] - h.f0;
                        tseg.clear(); oseg.clear(); code.clear(); xCell.clear(); bCell.clear();
                        xFace.clear(); xOwner.clear(); bFace.clear();
                        int32_t nx = 0, nb = 0, ne = 0;
                        // first pass: count cross and boundary faces so slots can be numbered
                        for (int32_t c = h.c0; c < h.c0 + h.nc; ++c)
                            for (int32_t k = seg[c]; k < seg[size_t(c) + 1]; ++k)
                            {
                                const int32_t f = ent[k] >> 1;
                                if (f >= nI) ++nb;
                                else if ((ent[k] & 1) && !(own[f] < nOwned && cellTile[own[f]] == t)) ++nx;
                            }
                        int32_t ix = 0, ib = 0;
                        for (int32_t c = h.c0; c < h.c0 + h.nc; ++c)
                        {
                            tseg.push_back(uint16_t(ne));
                            oseg.push_back(uint16_t(os[c] - h.f0));
                            for (int32_t k = seg[c]; k < seg[size_t(c) + 1]; ++k, ++ne)
                            {
                                const int32_t f = ent[k] >> 1, side = ent[k] & 1;
                                int32_t slot;
                                if (f >= nI)
                                {
                                    slot = h.nf + nx + ib++;
                                    bFace.push_back(f); bCell.push_back(uint16_t(c - h.c0));
                                }
                                else if (!side || (own[f] < nOwned && cellTile[own[f]] == t))
                                    slot = f - h.f0; // own face, or lower face owned inside the tile
                                else
                                {
                                    slot = h.nf + ix++;
                                    xFace.push_back(f); xOwner.push_back(own[f]); xCell.push_back(uint16_t(c - h.c0));
                                }
                                code.push_back(uint16_t((slot << 1) | side));
                            }
                        }
                        tseg.push_back(uint16_t(ne));
                        oseg.push_back(uint16_t(h.nf));
                        h.nx = nx; h.nb = nb; h.ne = ne;
                        if (h.nf + nx + nb >= 32768 || ne >= 65536) { ok = false; break; }
                        const FvkBlobLayout L = fvk_blob_layout(h.nc, h.nf, nx, nb, ne);
                        h.blobOff = int64_t(blob.size()); h.blobBytes = L.total;
                        blob.resize(blob.size() + size_t(L.total), 0);
                        unsigned char* bp = blob.data() + h.blobOff;
                        std::memcpy(bp, d->cellVolumes + h.c0, sizeof(double) * size_t(h.nc));
                        if (h.nf) std::memcpy(bp + L.nei, nei + h.f0, sizeof(int32_t) * size_t(h.nf));
                        if (nx) std::memcpy(bp + L.xFace, xFace.data(), sizeof(int32_t) * size_t(nx));
                        if (nx) std::memcpy(bp + L.xOwner, xOwner.data(), sizeof(int32_t) * size_t(nx));
                        if (nb) std::memcpy(bp + L.bFace, bFace.data(), sizeof(int32_t) * size_t(nb));
                        std::memcpy(bp + L.seg, tseg.data(), sizeof(uint16_t) * tseg.size());
                        std::memcpy(bp + L.oseg, oseg.data(), sizeof(uint16_t) * oseg.size());
                        if (ne) std::memcpy(bp + L.code, code.data(), sizeof(uint16_t) * size_t(ne));
                        if (nx) std::memcpy(bp + L.xCell, xCell.data(), sizeof(uint16_t) * size_t(nx));
                        if (nb) std::memcpy(bp + L.bCell, bCell.data(), sizeof(uint16_t) * size_t(nb));
                        tp.maxC = std::max(tp.maxC, h.nc); tp.maxF = std::max(tp.maxF, h.nf);
                        tp.maxX = std::max(tp.maxX, nx); tp.maxB = std::max(tp.maxB, nb); tp.maxE = std::max(tp.maxE, ne);
                        tp.maxBlob = std::max(tp.maxBlob, L.total);
                    }
                    (void) maxEnt;
                    if (ok)
                    {
                        UP(tp.hdr, hdr.data(), hdr.size());
                        UP(tp.blob, blob.data(), blob.size());
                        tp.nTiles = nT; tp.nCells = nC; tp.nB = nB;
                    }
                    else
                        tp = FvkTilePlan {};
                }
            }
        }
        tm.lap("tile plan (experiments)");
        UP(stencilSeg, seg.data(), seg.size());
        UP(stencilVal, val.data(), nEnt);
        UP(gatherEnt, ent.data(), nEnt);
        if (experiment_plans()) UP(gatherPlan, plan.data(), 2 * nEnt); // packed-plan gathers (fvk_set_variant 1-4)
        tm.lap("upload stencil");
    }
    // ---- sparsity pattern (host builder shared with the CPU self-test: fvk_build_sparsity, fvk_brickplan.cpp)
    {
        FvkSparsityHost sph;
        if (!fvk_build_sparsity(d, sth, sph))
        {
            const int32_t cell = sph.tooLongCell;
            fvk_mesh_destroy(m);
            return fvk_fail(FVK_EUNSUPPORTED, "fvk_mesh_create: cell %d has > 255 row entries (uint8 offsets)", cell);
        }
        m->rowsInStencilOrder = sph.rowsInStencilOrder;
        UP(rowOffs, sph.rowOffs.data(), sph.rowOffs.size());
        UP(colIdxs, sph.col.data(), sph.col.size());
        UP(ownerOffset, sph.ownOff.data(), nI);
        UP(neighbourOffset, sph.neiOff.data(), nI);
        UP(diagOffset, sph.diagOff.data(), nC);
        tm.lap("sparsity pattern + upload");
    }
    {
        bool sorted = true;
        for (int32_t f = 1; f < nI && sorted; ++f) sorted = own[f - 1] <= own[f];
        for (int32_t f = 0; f < nI && sorted; ++f) sorted = own[f] < nei[f];
        m->ownerSorted = sorted; // OpenFOAM upper-triangular face order
    }
    // ---- geometry scheme on device
    {
        cudaError_t e = cudaSuccess;
        for (double** p : {&m->weights, &m->deltaCoeffs, &m->nonOrthDeltaCoeffs})
            if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(p), sizeof(double) * size_t(nF) + 16);
        if (e == cudaSuccess)
        {
            const int grid = int((nF + 255) / 256 < 148 * 16 ? (nF + 255) / 256 : 148 * 16);
            k_geometry_scheme<<<grid > 0 ? grid : 1, 256>>>(nI, int(nF), m->owner, m->neighbour, m->C, m->Cf, m->Sf,
                                                            m->magSf, m->weights, m->deltaCoeffs, m->nonOrthDeltaCoeffs);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess)
        {
            fvk_mesh_destroy(m);
            return fvk_fail(FVK_ECUDA, "fvk_mesh_create: geometry scheme: %s", cudaGetErrorString(e));
        }
    }
    tm.lap("geometry scheme");
    *out = m;
    return FVK_OK;
}

int fvk_mesh_ensure_colors(const fvk_mesh* m)
{
    if (!m) return fvk_fail(FVK_EINVAL, "fvk_mesh_ensure_colors: null mesh");
    if (m->dicNColors > 0) return FVK_OK;
    const int32_t n = m->nOwned;
    std::vector<int32_t> ro(size_t(m->nCells) + 1);
    std::vector<int32_t> col;
    col.resize(size_t(m->nnz));
    FVK_CUDA(cudaMemcpy(ro.data(), m->rowOffs, sizeof(int32_t) * ro.size(), cudaMemcpyDeviceToHost));
    FVK_CUDA(cudaMemcpy(col.data(), m->colIdxs, sizeof(int32_t) * col.size(), cudaMemcpyDeviceToHost));
    std::vector<uint8_t> color(size_t(m->nCells), 255);
    int32_t nColors = 0;
    for (int32_t c = 0; c < n; ++c)
    {
        uint64_t used = 0;
        for (int32_t k = ro[c]; k < ro[size_t(c) + 1]; ++k)
        {
            const int32_t j = col[k];
            if (j != c && j < n && color[j] != 255) used |= uint64_t(1) << color[j];
        }
        int32_t cc = 0;
        while (cc < 64 && (used >> cc & 1)) ++cc;
        if (cc >= 64) return fvk_fail(FVK_EUNSUPPORTED, "fvk_mesh_ensure_colors: more than 64 colours");
        color[c] = uint8_t(cc);
        nColors = std::max(nColors, cc + 1);
    }
    std::vector<int32_t> off(size_t(nColors) + 1, 0);
    std::vector<int32_t> cells;
    cells.resize(size_t(n));
    for (int32_t c = 0; c < n; ++c) ++off[size_t(color[c]) + 1];
    for (int32_t k = 0; k < nColors; ++k) off[size_t(k) + 1] += off[k];
    std::vector<int32_t> pos(off.begin(), off.end() - 1);
    for (int32_t c = 0; c < n; ++c) cells[pos[color[c]]++] = c;
    FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->dicColor), color.size()));
    FVK_CUDA(cudaMemcpy(m->dicColor, color.data(), color.size(), cudaMemcpyHostToDevice));
    FVK_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->dicCells), sizeof(int32_t) * std::max<size_t>(cells.size(), 1)));
    FVK_CUDA(cudaMemcpy(m->dicCells, cells.data(), sizeof(int32_t) * cells.size(), cudaMemcpyHostToDevice));
    for (int32_t k = 0; k <= nColors; ++k) m->dicOff[k] = off[k];
    m->dicNColors = nColors;
    return FVK_OK;
}

extern "C" int fvk_mesh_size(const fvk_mesh* m, int field, int64_t* value)
{
    if (!m || !value) return fvk_fail(FVK_EINVAL, "fvk_mesh_size: null");
    switch (field)
    {
        case FVK_N_CELLS: *value = m->nCells; break;
        case FVK_N_INTERNAL_FACES: *value = m->nInternalFaces; break;
        case FVK_N_BOUNDARY_FACES: *value = m->nBoundaryFaces; break;
        case FVK_N_PATCHES: *value = m->nPatches; break;
        case FVK_NNZ: *value = m->nnz; break;
        case FVK_N_OWNED_CELLS: *value = m->nOwned; break;
        case FVK_ROWS_IN_STENCIL_ORDER: *value = m->rowsInStencilOrder ? 1 : 0; break;
        case FVK_AFFINE_TOPOLOGY: *value = (m->bp.nTiles > 0 && m->bp.geom.affine) ? 1 : 0; break;
        default: return fvk_fail(FVK_EINVAL, "fvk_mesh_size: unknown field %d", field);
    }
    return FVK_OK;
}

extern "C" int fvk_mesh_array(const fvk_mesh* m, int field, const void** dptr, int64_t* count)
{
    if (!m || !dptr || !count) return fvk_fail(FVK_EINVAL, "fvk_mesh_array: null");
    const int64_t nC = m->nCells, nI = m->nInternalFaces, nB = m->nBoundaryFaces, nF = nI + nB;
    const void* p = nullptr;
    int64_t n = 0;
    switch (field)
    {
        case FVK_CELL_VOLUMES: p = m->V; n = nC; break;
        case FVK_CELL_CENTRES: p = m->C; n = 3 * nC; break;
        case FVK_FACE_AREAS: p = m->Sf; n = 3 * nF; break;
        case FVK_FACE_CENTRES: p = m->Cf; n = 3 * nF; break;
        case FVK_MAG_FACE_AREAS: p = m->magSf; n = nF; break;
        case FVK_FACE_OWNER: p = m->owner; n = nF; break;
        case FVK_FACE_NEIGHBOUR: p = m->neighbour; n = nI; break;
        case FVK_FACE_CELLS: p = m->faceCells; n = nB; break;
        case FVK_B_CF: p = m->bCf; n = 3 * nB; break;
        case FVK_B_CN: p = m->bCn; n = 3 * nB; break;
        case FVK_B_SF: p = m->bSf; n = 3 * nB; break;
        case FVK_B_MAGSF: p = m->bMagSf; n = nB; break;
        case FVK_B_NF: p = m->bNf; n = 3 * nB; break;
        case FVK_B_DELTA: p = m->bDelta; n = 3 * nB; break;
        case FVK_B_WEIGHTS: p = m->bWeights; n = nB; break;
        case FVK_B_DELTACOEFFS: p = m->bDeltaCoeffs; n = nB; break;
        case FVK_WEIGHTS: p = m->weights; n = nF; break;
        case FVK_DELTACOEFFS: p = m->deltaCoeffs; n = nF; break;
        case FVK_NONORTH_DELTACOEFFS: p = m->nonOrthDeltaCoeffs; n = nF; break;
        case FVK_STENCIL_SEGMENTS: p = m->stencilSeg; n = nC + 1; break;
        case FVK_STENCIL_VALUES: p = m->stencilVal; n = 2 * nI + nB; break;
        case FVK_ROW_OFFS: p = m->rowOffs; n = nC + 1; break;
        case FVK_COL_IDXS: p = m->colIdxs; n = m->nnz; break;
        case FVK_OWNER_OFFSET: p = m->ownerOffset; n = nI; break;
        case FVK_NEIGHBOUR_OFFSET: p = m->neighbourOffset; n = nI; break;
        case FVK_DIAG_OFFSET: p = m->diagOffset; n = nC; break;
        default: return fvk_fail(FVK_EINVAL, "fvk_mesh_array: unknown field %d", field);
    }
    *dptr = p;
    *count = p ? n : 0;
    return FVK_OK;
}

extern "C" int fvk_mesh_patch_offsets(const fvk_mesh* m, int32_t* offsets_h)
{
    if (!m || !offsets_h) return fvk_fail(FVK_EINVAL, "fvk_mesh_patch_offsets: null");
    for (int p = 0; p <= m->nPatches; ++p) offsets_h[p] = m->patchOffsets[p];
    return FVK_OK;
}
