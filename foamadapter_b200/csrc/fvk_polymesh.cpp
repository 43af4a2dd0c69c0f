// Host-side reader / writer for OpenFOAM ASCII polyMesh directories (points, faces, owner, neighbour, boundary).
//
// Stand-in for Foam::polyMesh + FoamAdapter::readOpenFOAMMesh (reference src/datastructures/meshAdapter.cpp:12-45,59-136),
// which need OpenFOAM: the files are parsed here, the geometry follows OpenFOAM's primitiveMesh (triangle-fan face
// centres / areas about the vertex average, pyramid cell centres / volumes about the face-centre average -- the same
// arithmetic as fvk_blockmesh.cpp, for arbitrary polygons / polyhedra), and the boundary is flattened like the reference's
// converter: patches concatenated in file order, `empty` patches contribute no faces (fvPatch::size() == 0). NOTE (deviation
// from meshAdapter.cpp:33-45,63, which keeps an empty patch as a zero-size entry that still occupies a patch index): here an
// `empty` patch is dropped from the patch list altogether, so patch indices are those of the KEPT patches in file order;
// fvk_polymesh_patch gives each kept patch's name and type, fvk_polymesh_patch_file_index its index in the boundary file.
// The result is an fvk_mesh_desc that fvk_mesh_create / fvk_decompose take like a generated block mesh.
#include "fvk_internal.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <new>
#include <sstream>
#include <string>
#include <vector>

namespace
{
struct PolyStore
{
    fvk_mesh_desc desc {};
    char magic[8]; // right behind desc: storeOf() may be handed a desc that is not ours
    std::vector<double> points, V, C, Sf, Cf, magSf;
    std::vector<int32_t> owner, neighbour, faceCells, patchOffsets;
    std::vector<double> bCf, bCn, bSf, bMagSf, bNf, bDelta, bWeights, bDeltaCoeffs;
    std::vector<std::string> patchName, patchType; // kept (non-empty) patches
    // the polyMesh as read
    std::vector<int32_t> faceOff, facePts, polyOwner;
    std::vector<std::string> allName, allType;
    std::vector<int32_t> allStart, allSize;
};
constexpr char kMagic[8] = {'F', 'V', 'K', 'P', 'O', 'L', 'Y', 0};

// ---- tokenizer over a file with comments and the FoamFile header removed --------------------------------------------
struct Text
{
    std::string s;
    size_t i = 0;
    void skipWs() { while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i; }
    bool eof() { skipWs(); return i >= s.size(); }
    char peek() { skipWs(); return i < s.size() ? s[i] : '\0'; }
    bool take(char c) { if (peek() == c) { ++i; return true; } return false; }
    // a word / number token: up to whitespace or one of ( ) { } ;
    std::string word()
    {
        skipWs();
        const size_t b = i;
        while (i < s.size() && !std::isspace(static_cast<unsigned char>(s[i])) && !std::strchr("(){};", s[i])) ++i;
        return s.substr(b, i - b);
    }
};

bool load(const std::string& path, Text& t, std::string& err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = "cannot open " + path; return false; }
    std::stringstream ss;
    ss << in.rdbuf();
    std::string raw = ss.str(), out;
    out.reserve(raw.size());
    for (size_t k = 0; k < raw.size();)
    {
        if (raw.compare(k, 2, "/*") == 0)
        {
            const size_t e = raw.find("*/", k + 2);
            k = (e == std::string::npos) ? raw.size() : e + 2;
            out.push_back(' ');
        }
        else if (raw.compare(k, 2, "//") == 0)
        {
            const size_t e = raw.find('\n', k);
            k = (e == std::string::npos) ? raw.size() : e;
        }
        else
            out.push_back(raw[k++]);
    }
    // FoamFile { ... } header: note the format, then drop it
    const size_t h = out.find("FoamFile");
    if (h != std::string::npos)
    {
        const size_t b = out.find('{', h), e = out.find('}', h);
        if (b == std::string::npos || e == std::string::npos) { err = path + ": malformed FoamFile header"; return false; }
        const std::string hdr = out.substr(b, e - b);
        const size_t f = hdr.find("format");
        if (f != std::string::npos && hdr.find("binary", f) != std::string::npos && hdr.find("binary", f) < hdr.find(';', f))
        {
            err = path + ": binary format is not supported (write the case with writeFormat ascii)";
            return false;
        }
        out.erase(h, e + 1 - h);
    }
    t.s.swap(out);
    t.i = 0;
    return true;
}

bool listHeader(Text& t, int64_t& n, const std::string& path, std::string& err)
{
    const std::string w = t.word();
    char* end = nullptr;
    n = std::strtoll(w.c_str(), &end, 10);
    if (w.empty() || *end || n < 0) { err = path + ": expected a list size, got '" + w + "'"; return false; }
    if (!t.take('(')) { err = path + ": expected '(' after the list size (uniform '{}' lists are not supported)"; return false; }
    return true;
}

bool readLabels(const std::string& path, std::vector<int32_t>& out, std::string& err)
{
    Text t;
    if (!load(path, t, err)) return false;
    int64_t n;
    if (!listHeader(t, n, path, err)) return false;
    out.resize(size_t(n));
    for (int64_t k = 0; k < n; ++k)
    {
        const std::string w = t.word();
        char* end = nullptr;
        const long v = std::strtol(w.c_str(), &end, 10);
        if (w.empty() || *end) { err = path + ": bad label '" + w + "'"; return false; }
        out[size_t(k)] = int32_t(v);
    }
    if (!t.take(')')) { err = path + ": list longer than its size"; return false; }
    return true;
}

bool readPoints(const std::string& path, std::vector<double>& out, std::string& err)
{
    Text t;
    if (!load(path, t, err)) return false;
    int64_t n;
    if (!listHeader(t, n, path, err)) return false;
    out.resize(3 * size_t(n));
    for (int64_t k = 0; k < n; ++k)
    {
        if (!t.take('(')) { err = path + ": expected '(' of a point"; return false; }
        for (int d = 0; d < 3; ++d)
        {
            const std::string w = t.word();
            char* end = nullptr;
            out[3 * size_t(k) + d] = std::strtod(w.c_str(), &end);
            if (w.empty() || *end) { err = path + ": bad coordinate '" + w + "'"; return false; }
        }
        if (!t.take(')')) { err = path + ": expected ')' of a point"; return false; }
    }
    if (!t.take(')')) { err = path + ": list longer than its size"; return false; }
    return true;
}

bool readFaces(const std::string& path, std::vector<int32_t>& off, std::vector<int32_t>& pts, std::string& err)
{
    Text t;
    if (!load(path, t, err)) return false;
    int64_t n;
    if (!listHeader(t, n, path, err)) return false;
    off.assign(1, 0);
    off.reserve(size_t(n) + 1);
    pts.reserve(4 * size_t(n));
    for (int64_t k = 0; k < n; ++k)
    {
        const std::string w = t.word();
        char* end = nullptr;
        const long m = std::strtol(w.c_str(), &end, 10);
        if (w.empty() || *end || m < 3) { err = path + ": bad face size '" + w + "'"; return false; }
        if (!t.take('(')) { err = path + ": expected '(' of a face"; return false; }
        for (long q = 0; q < m; ++q)
        {
            const std::string v = t.word();
            char* e2 = nullptr;
            const long p = std::strtol(v.c_str(), &e2, 10);
            if (v.empty() || *e2 || p < 0) { err = path + ": bad point label '" + v + "'"; return false; }
            pts.push_back(int32_t(p));
        }
        if (!t.take(')')) { err = path + ": expected ')' of a face"; return false; }
        off.push_back(int32_t(pts.size()));
    }
    if (!t.take(')')) { err = path + ": list longer than its size"; return false; }
    return true;
}

bool readBoundary(const std::string& path, PolyStore& st, std::string& err)
{
    Text t;
    if (!load(path, t, err)) return false;
    int64_t n;
    if (!listHeader(t, n, path, err)) return false;
    for (int64_t p = 0; p < n; ++p)
    {
        const std::string name = t.word();
        if (name.empty() || !t.take('{')) { err = path + ": expected 'patchName {'"; return false; }
        std::string type;
        long nFaces = -1, startFace = -1;
        while (!t.take('}'))
        {
            if (t.eof()) { err = path + ": unterminated patch dictionary"; return false; }
            const std::string key = t.word();
            // value: everything up to ';' (may contain lists such as `inGroups 1(wall)`)
            std::string val;
            int depth = 0;
            for (;;)
            {
                if (t.i >= t.s.size()) { err = path + ": missing ';'"; return false; }
                const char c = t.s[t.i++];
                if (c == '(') ++depth;
                if (c == ')') --depth;
                if (c == ';' && depth == 0) break;
                val.push_back(c);
            }
            size_t b = 0, e = val.size();
            while (b < e && std::isspace(static_cast<unsigned char>(val[b]))) ++b;
            while (e > b && std::isspace(static_cast<unsigned char>(val[e - 1]))) --e;
            val = val.substr(b, e - b);
            if (key == "type") type = val;
            else if (key == "nFaces") nFaces = std::strtol(val.c_str(), nullptr, 10);
            else if (key == "startFace") startFace = std::strtol(val.c_str(), nullptr, 10);
        }
        if (type.empty() || nFaces < 0 || startFace < 0) { err = path + ": patch '" + name + "' lacks type / nFaces / startFace"; return false; }
        st.allName.push_back(name); st.allType.push_back(type);
        st.allStart.push_back(int32_t(startFace)); st.allSize.push_back(int32_t(nFaces));
    }
    if (!t.take(')')) { err = path + ": list longer than its size"; return false; }
    return true;
}

inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// OpenFOAM primitiveMeshFaceCentresAndAreas.C: triangles directly, other polygons as a fan about the vertex average
inline void polygonGeometry(const double* p, const int32_t* f, int n, double* cf, double* sf)
{
    if (n == 3)
    {
        const double *a = p + 3 * f[0], *b = p + 3 * f[1], *c = p + 3 * f[2];
        double e1[3], e2[3], nn[3];
        for (int d = 0; d < 3; ++d)
        {
            cf[d] = (1.0 / 3.0) * (a[d] + b[d] + c[d]);
            e1[d] = b[d] - a[d];
            e2[d] = c[d] - a[d];
        }
        cross3(e1, e2, nn);
        for (int d = 0; d < 3; ++d) sf[d] = 0.5 * nn[d];
        return;
    }
    double fc[3];
    for (int d = 0; d < 3; ++d)
    {
        double s = p[3 * f[0] + d];
        for (int k = 1; k < n; ++k) s += p[3 * f[k] + d];
        fc[d] = s / n;
    }
    double sumN[3] = {0, 0, 0}, sumAc[3] = {0, 0, 0}, sumA = 0.0;
    for (int k = 0; k < n; ++k)
    {
        const double* cur = p + 3 * f[k];
        const double* nxt = p + 3 * f[(k + 1) % n];
        double e1[3], e2[3], nn[3], c[3];
        for (int d = 0; d < 3; ++d)
        {
            c[d] = cur[d] + nxt[d] + fc[d];
            e1[d] = nxt[d] - cur[d];
            e2[d] = fc[d] - cur[d];
        }
        cross3(e1, e2, nn);
        const double a = std::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
        sumA += a;
        for (int d = 0; d < 3; ++d)
        {
            sumN[d] += nn[d];
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

PolyStore* storeOf(const fvk_mesh_desc* d)
{
    PolyStore* st = const_cast<PolyStore*>(reinterpret_cast<const PolyStore*>(d));
    return (st && std::memcmp(st->magic, kMagic, sizeof(kMagic)) == 0) ? st : nullptr;
}
} // namespace

extern "C" int fvk_polymesh_read(const char* polyMeshDir, fvk_mesh_desc** out)
{
    if (!polyMeshDir || !out) return fvk_fail(FVK_EINVAL, "fvk_polymesh_read: null argument");
    *out = nullptr;
    PolyStore* st = new (std::nothrow) PolyStore;
    if (!st) return fvk_fail(FVK_ENOMEM, "fvk_polymesh_read: out of memory");
    std::memcpy(st->magic, kMagic, sizeof(kMagic));
    const std::string dir(polyMeshDir);
    std::string err;
    try
    {
        if (!readPoints(dir + "/points", st->points, err) || !readFaces(dir + "/faces", st->faceOff, st->facePts, err)
            || !readLabels(dir + "/owner", st->polyOwner, err) || !readLabels(dir + "/neighbour", st->neighbour, err)
            || !readBoundary(dir + "/boundary", *st, err))
        {
            delete st;
            return fvk_fail(FVK_EINVAL, "fvk_polymesh_read: %s", err.c_str());
        }
        const int64_t nP = int64_t(st->points.size() / 3), nPoly = int64_t(st->faceOff.size()) - 1, nI = int64_t(st->neighbour.size());
        auto bad = [&](const std::string& m) { delete st; return fvk_fail(FVK_EINVAL, "fvk_polymesh_read: %s: %s", polyMeshDir, m.c_str()); };
        if (int64_t(st->polyOwner.size()) != nPoly) return bad("owner and faces differ in length");
        if (nI > nPoly) return bad("more neighbours than faces");
        if (nPoly >= (int64_t(1) << 30)) { delete st; return fvk_fail(FVK_EUNSUPPORTED, "fvk_polymesh_read: > 2^30 faces"); }
        int32_t nC = 0;
        for (int64_t f = 0; f < nPoly; ++f)
        {
            if (st->polyOwner[f] < 0) return bad("negative owner label");
            nC = std::max(nC, st->polyOwner[f] + 1);
        }
        for (int64_t f = 0; f < nI; ++f)
        {
            if (st->neighbour[f] < 0) return bad("negative neighbour label");
            nC = std::max(nC, st->neighbour[f] + 1);
        }
        for (int32_t p : st->facePts)
            if (p >= nP) return bad("face refers to a point that does not exist");
        // patches must tile [nI, nPoly) in order
        int64_t expect = nI;
        for (size_t p = 0; p < st->allName.size(); ++p)
        {
            if (st->allStart[p] != expect) return bad("patch '" + st->allName[p] + "' does not start where the previous one ends");
            expect += st->allSize[p];
        }
        if (expect != nPoly) return bad("patches do not cover the boundary faces");

        // ---- geometry on all poly faces (OpenFOAM order of operations, as fvk_blockmesh.cpp)
        std::vector<double> pCf(3 * size_t(nPoly)), pSf(3 * size_t(nPoly));
#pragma omp parallel for schedule(static)
        for (int64_t f = 0; f < nPoly; ++f)
            polygonGeometry(st->points.data(), &st->facePts[st->faceOff[f]], st->faceOff[f + 1] - st->faceOff[f], &pCf[3 * f], &pSf[3 * f]);
        std::vector<double> cEst(3 * size_t(nC), 0.0);
        std::vector<int32_t> cnt(size_t(nC), 0);
        for (int64_t f = 0; f < nPoly; ++f)
        {
            const int32_t o = st->polyOwner[f];
            for (int d = 0; d < 3; ++d) cEst[3 * size_t(o) + d] += pCf[3 * f + d];
            ++cnt[o];
        }
        for (int64_t f = 0; f < nI; ++f)
        {
            const int32_t n = st->neighbour[f];
            for (int d = 0; d < 3; ++d) cEst[3 * size_t(n) + d] += pCf[3 * f + d];
            ++cnt[n];
        }
        for (int32_t c = 0; c < nC; ++c)
        {
            if (cnt[c] == 0) return bad("a cell has no faces");
            for (int d = 0; d < 3; ++d) cEst[3 * size_t(c) + d] /= cnt[c];
        }
        st->C.assign(3 * size_t(nC), 0.0);
        st->V.assign(size_t(nC), 0.0);
        for (int64_t f = 0; f < nPoly; ++f)
        {
            const size_t o = size_t(st->polyOwner[f]);
            double pyr3 = 0.0;
            for (int d = 0; d < 3; ++d) pyr3 += pSf[3 * f + d] * (pCf[3 * f + d] - cEst[3 * o + d]);
            for (int d = 0; d < 3; ++d) st->C[3 * o + d] += pyr3 * ((3.0 / 4.0) * pCf[3 * f + d] + (1.0 / 4.0) * cEst[3 * o + d]);
            st->V[o] += pyr3;
        }
        for (int64_t f = 0; f < nI; ++f)
        {
            const size_t n = size_t(st->neighbour[f]);
            double pyr3 = 0.0;
            for (int d = 0; d < 3; ++d) pyr3 += pSf[3 * f + d] * (cEst[3 * n + d] - pCf[3 * f + d]);
            for (int d = 0; d < 3; ++d) st->C[3 * n + d] += pyr3 * ((3.0 / 4.0) * pCf[3 * f + d] + (1.0 / 4.0) * cEst[3 * n + d]);
            st->V[n] += pyr3;
        }
        for (int32_t c = 0; c < nC; ++c)
        {
            if (std::fabs(st->V[c]) > 1e-300)
                for (int d = 0; d < 3; ++d) st->C[3 * size_t(c) + d] /= st->V[c];
            else
                for (int d = 0; d < 3; ++d) st->C[3 * size_t(c) + d] = cEst[3 * size_t(c) + d];
            st->V[c] *= (1.0 / 3.0);
        }

        // ---- NeoN view: internal faces + the faces of every non-empty patch, patch after patch (meshAdapter.cpp:12-45)
        st->patchOffsets.assign(1, 0);
        std::vector<int32_t> keep;
        for (size_t p = 0; p < st->allName.size(); ++p)
        {
            if (st->allType[p] == "empty") continue;
            for (int32_t q = 0; q < st->allSize[p]; ++q) keep.push_back(st->allStart[p] + q);
            st->patchOffsets.push_back(int32_t(keep.size()));
            st->patchName.push_back(st->allName[p]); st->patchType.push_back(st->allType[p]);
        }
        const int64_t nB = int64_t(keep.size()), nF = nI + nB;
        if (int64_t(st->patchName.size()) > FVK_MAX_PATCHES) { delete st; return fvk_fail(FVK_EUNSUPPORTED, "fvk_polymesh_read: more than %d patches", FVK_MAX_PATCHES); }
        st->owner.resize(size_t(nF)); st->Sf.resize(3 * size_t(nF)); st->Cf.resize(3 * size_t(nF)); st->magSf.resize(size_t(nF));
        for (int64_t f = 0; f < nF; ++f)
        {
            const int64_t src = f < nI ? f : keep[size_t(f - nI)];
            st->owner[f] = st->polyOwner[src];
            for (int d = 0; d < 3; ++d) { st->Sf[3 * f + d] = pSf[3 * src + d]; st->Cf[3 * f + d] = pCf[3 * src + d]; }
            const double* s = &st->Sf[3 * f];
            st->magSf[f] = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        }
        st->faceCells.resize(size_t(nB));
        st->bCf.resize(3 * size_t(nB)); st->bCn.resize(3 * size_t(nB)); st->bSf.resize(3 * size_t(nB));
        st->bNf.resize(3 * size_t(nB)); st->bDelta.resize(3 * size_t(nB));
        st->bMagSf.resize(size_t(nB)); st->bWeights.resize(size_t(nB)); st->bDeltaCoeffs.resize(size_t(nB));
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
        fvk_mesh_desc& d = st->desc;
        d.nCells = nC; d.nInternalFaces = int32_t(nI); d.nBoundaryFaces = int32_t(nB); d.nPatches = int32_t(st->patchName.size());
        d.nPoints = int32_t(nP); d.points = st->points.data();
        d.cellVolumes = st->V.data(); d.cellCentres = st->C.data();
        d.faceAreas = st->Sf.data(); d.faceCentres = st->Cf.data(); d.magFaceAreas = st->magSf.data();
        d.faceOwner = st->owner.data(); d.faceNeighbour = st->neighbour.data();
        d.faceCells = st->faceCells.data();
        d.bCf = st->bCf.data(); d.bCn = st->bCn.data(); d.bSf = st->bSf.data();
        d.bMagSf = st->bMagSf.data(); d.bNf = st->bNf.data(); d.bDelta = st->bDelta.data();
        d.bWeights = st->bWeights.data(); d.bDeltaCoeffs = st->bDeltaCoeffs.data();
        d.patchOffsets = st->patchOffsets.data();
        d.nOwnedCells = 0; d.faceOrder = nullptr;
    }
    catch (const std::bad_alloc&)
    {
        delete st;
        return fvk_fail(FVK_ENOMEM, "fvk_polymesh_read: out of memory");
    }
    static_assert(offsetof(PolyStore, desc) == 0, "desc must be first");
    *out = &st->desc;
    return FVK_OK;
}

extern "C" int fvk_polymesh_destroy(fvk_mesh_desc* desc)
{
    if (!desc) return FVK_OK;
    PolyStore* st = storeOf(desc);
    if (!st) return fvk_fail(FVK_EINVAL, "fvk_polymesh_destroy: not a mesh from fvk_polymesh_read");
    st->magic[0] = 0;
    delete st;
    return FVK_OK;
}

// index of kept patch `patch` in the case's `boundary` file (= OpenFOAM's patch id, which counts the dropped `empty` patches)
extern "C" int fvk_polymesh_patch_file_index(const fvk_mesh_desc* desc, int32_t patch, int32_t* fileIndex)
{
    const PolyStore* st = storeOf(desc);
    if (!st || !fileIndex) return fvk_fail(FVK_EINVAL, "fvk_polymesh_patch_file_index: not a mesh from fvk_polymesh_read");
    if (patch < 0 || patch >= int32_t(st->patchName.size())) return fvk_fail(FVK_EINVAL, "fvk_polymesh_patch_file_index: patch %d out of range", patch);
    int32_t kept = -1;
    for (size_t i = 0; i < st->allName.size(); ++i)
    {
        if (st->allType[i] != "empty") ++kept;
        if (kept == patch) { *fileIndex = int32_t(i); return FVK_OK; }
    }
    return fvk_fail(FVK_EINVAL, "fvk_polymesh_patch_file_index: patch %d not found", patch);
}

extern "C" int fvk_polymesh_patch(const fvk_mesh_desc* desc, int32_t patch, char* name, int32_t nameCap, char* type, int32_t typeCap)
{
    const PolyStore* st = storeOf(desc);
    if (!st) return fvk_fail(FVK_EINVAL, "fvk_polymesh_patch: not a mesh from fvk_polymesh_read");
    if (patch < 0 || patch >= int32_t(st->patchName.size())) return fvk_fail(FVK_EINVAL, "fvk_polymesh_patch: patch %d out of range", patch);
    if (name && nameCap > 0) std::snprintf(name, size_t(nameCap), "%s", st->patchName[size_t(patch)].c_str());
    if (type && typeCap > 0) std::snprintf(type, size_t(typeCap), "%s", st->patchType[size_t(patch)].c_str());
    return FVK_OK;
}

extern "C" int fvk_polymesh_write(const char* polyMeshDir, int32_t nPoints, const double* points, int32_t nFaces,
                                  const int32_t* faceOffsets, const int32_t* facePoints, const int32_t* owner,
                                  int32_t nInternalFaces, const int32_t* neighbour, int32_t nPatches,
                                  const char* const* patchNames, const char* const* patchTypes, const int32_t* patchSizes)
{
    if (!polyMeshDir || nPoints < 0 || nFaces < 0 || nInternalFaces < 0 || nInternalFaces > nFaces || nPatches < 0 || (nPoints && !points)
        || (nFaces && (!faceOffsets || !facePoints || !owner)) || (nInternalFaces && !neighbour)
        || (nPatches && (!patchNames || !patchTypes || !patchSizes)))
        return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: bad argument");
    int64_t cover = nInternalFaces;
    for (int32_t p = 0; p < nPatches; ++p) cover += patchSizes[p];
    if (cover != nFaces) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: patches do not cover the boundary faces");
    int32_t nC = 0;
    for (int32_t f = 0; f < nFaces; ++f) nC = std::max(nC, owner[f] + 1);
    for (int32_t f = 0; f < nInternalFaces; ++f) nC = std::max(nC, neighbour[f] + 1);
    const std::string dir(polyMeshDir);
    auto open = [&](const char* file, const char* cls, const std::string& note) -> FILE* {
        FILE* fp = std::fopen((dir + "/" + file).c_str(), "w");
        if (!fp) return nullptr;
        std::fprintf(fp, "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n", cls);
        if (!note.empty()) std::fprintf(fp, "    note        \"%s\";\n", note.c_str());
        std::fprintf(fp, "    location    \"constant/polyMesh\";\n    object      %s;\n}\n\n", file);
        return fp;
    };
    char nb[160];
    std::snprintf(nb, sizeof nb, "nPoints:%d  nCells:%d  nFaces:%d  nInternalFaces:%d", nPoints, nC, nFaces, nInternalFaces);
    FILE* fp = open("points", "vectorField", "");
    if (!fp) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: cannot write into %s", polyMeshDir);
    std::fprintf(fp, "%d\n(\n", nPoints);
    for (int32_t p = 0; p < nPoints; ++p) std::fprintf(fp, "(%.17g %.17g %.17g)\n", points[3 * p], points[3 * p + 1], points[3 * p + 2]);
    std::fprintf(fp, ")\n");
    std::fclose(fp);
    if (!(fp = open("faces", "faceList", ""))) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: cannot write faces");
    std::fprintf(fp, "%d\n(\n", nFaces);
    for (int32_t f = 0; f < nFaces; ++f)
    {
        std::fprintf(fp, "%d(", faceOffsets[f + 1] - faceOffsets[f]);
        for (int32_t q = faceOffsets[f]; q < faceOffsets[f + 1]; ++q) std::fprintf(fp, q + 1 < faceOffsets[f + 1] ? "%d " : "%d", facePoints[q]);
        std::fprintf(fp, ")\n");
    }
    std::fprintf(fp, ")\n");
    std::fclose(fp);
    if (!(fp = open("owner", "labelList", nb))) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: cannot write owner");
    std::fprintf(fp, "%d\n(\n", nFaces);
    for (int32_t f = 0; f < nFaces; ++f) std::fprintf(fp, "%d\n", owner[f]);
    std::fprintf(fp, ")\n");
    std::fclose(fp);
    if (!(fp = open("neighbour", "labelList", nb))) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: cannot write neighbour");
    std::fprintf(fp, "%d\n(\n", nInternalFaces);
    for (int32_t f = 0; f < nInternalFaces; ++f) std::fprintf(fp, "%d\n", neighbour[f]);
    std::fprintf(fp, ")\n");
    std::fclose(fp);
    if (!(fp = open("boundary", "polyBoundaryMesh", ""))) return fvk_fail(FVK_EINVAL, "fvk_polymesh_write: cannot write boundary");
    std::fprintf(fp, "%d\n(\n", nPatches);
    int32_t start = nInternalFaces;
    for (int32_t p = 0; p < nPatches; ++p)
    {
        std::fprintf(fp, "    %s\n    {\n        type            %s;\n", patchNames[p], patchTypes[p]);
        if (!std::strcmp(patchTypes[p], "wall")) std::fprintf(fp, "        inGroups        1(wall);\n");
        std::fprintf(fp, "        nFaces          %d;\n        startFace       %d;\n    }\n", patchSizes[p], start);
        start += patchSizes[p];
    }
    std::fprintf(fp, ")\n");
    std::fclose(fp);
    return FVK_OK;
}

// labelList files, e.g. constant/cellDecomposition written by `decomposePar -cellDist` (any decomposition method:
// scotch, hierarchical, ...): the cell -> rank map fvk_decompose takes.
extern "C" int fvk_labellist_read(const char* path, int32_t nExpected, int32_t* out)
{
    if (!path || nExpected < 0 || (nExpected && !out)) return fvk_fail(FVK_EINVAL, "fvk_labellist_read: bad argument");
    std::vector<int32_t> v;
    std::string err;
    if (!readLabels(path, v, err)) return fvk_fail(FVK_EINVAL, "fvk_labellist_read: %s", err.c_str());
    if (int64_t(v.size()) != nExpected)
        return fvk_fail(FVK_EINVAL, "fvk_labellist_read: %s has %zu entries, expected %d", path, v.size(), nExpected);
    if (nExpected) std::memcpy(out, v.data(), sizeof(int32_t) * v.size());
    return FVK_OK;
}
