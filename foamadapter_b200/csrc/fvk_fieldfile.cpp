// Host-side reader for OpenFOAM ASCII field files (0/T, 0/U, 0/phi, ...): the internal field and the per-patch
// boundary dictionaries, without OpenFOAM. Stand-in for the Foam::GeometricField constructor the reference's converters
// start from (src/datastructures/fieldAdapter / include/FoamAdapter/auxiliary/readers.hpp in the reference read an
// existing Foam field and copy internalField / boundaryField into NeoN containers). Formats as in the reference's
// fixtures test/setup_operator/0/*: `internalField uniform v;`, `internalField nonuniform List<scalar|vector> n ( ... );`,
// `boundaryField { patch { type T; value uniform v; } ... }`.
#include "fvk_internal.hpp"

#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace
{
bool slurp(const char* path, std::string& out, std::string& err)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) { err = std::string("cannot open ") + path; return false; }
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string raw = ss.str();
    out.clear();
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
    const size_t h = out.find("FoamFile");
    if (h != std::string::npos)
    {
        const size_t b = out.find('{', h), e = out.find('}', h);
        if (b == std::string::npos || e == std::string::npos) { err = std::string(path) + ": malformed FoamFile header"; return false; }
        const std::string hdr = out.substr(b, e - b);
        const size_t f = hdr.find("format");
        if (f != std::string::npos && hdr.find("binary", f) != std::string::npos && hdr.find("binary", f) < hdr.find(';', f))
        {
            err = std::string(path) + ": binary format is not supported";
            return false;
        }
        out.erase(h, e + 1 - h);
    }
    return true;
}

// position just behind the whole word `key` at brace depth `depth` (0 = top level), or npos
size_t findKey(const std::string& s, const std::string& key, size_t from, size_t to, int depthWanted)
{
    int depth = 0;
    for (size_t i = from; i < to; ++i)
    {
        const char c = s[i];
        if (c == '{') ++depth;
        else if (c == '}') --depth;
        else if (depth == depthWanted && s.compare(i, key.size(), key) == 0
                 && (i == 0 || !(std::isalnum(static_cast<unsigned char>(s[i - 1])) || s[i - 1] == '_'))
                 && (i + key.size() >= s.size() || !(std::isalnum(static_cast<unsigned char>(s[i + key.size()])) || s[i + key.size()] == '_')))
            return i + key.size();
    }
    return std::string::npos;
}

// parse `uniform v;` / `uniform (a b c);` / `nonuniform List<T> n ( ... );` starting at `pos`
int parseValue(const std::string& s, size_t pos, const char* what, int64_t nExpected, int32_t* ncomp, double* out, int64_t cap, double uni[3],
               int* isUniform)
{
    auto skip = [&](size_t& p) { while (p < s.size() && std::isspace(static_cast<unsigned char>(s[p]))) ++p; };
    auto word = [&](size_t& p) { skip(p); const size_t b = p; while (p < s.size() && !std::isspace(static_cast<unsigned char>(s[p])) && !std::strchr("();{}", s[p])) ++p; return s.substr(b, p - b); };
    auto number = [&](size_t& p, double& v) { skip(p); char* e = nullptr; v = std::strtod(s.c_str() + p, &e); if (e == s.c_str() + p) return false; p = size_t(e - s.c_str()); return true; };
    size_t p = pos;
    const std::string kind = word(p);
    if (kind == "uniform")
    {
        skip(p);
        int nc = 1;
        double v[3] = {0, 0, 0};
        if (p < s.size() && s[p] == '(')
        {
            ++p; nc = 3;
            for (int d = 0; d < 3; ++d)
                if (!number(p, v[d])) return fvk_fail(FVK_EINVAL, "%s: bad uniform vector", what);
        }
        else if (!number(p, v[0])) return fvk_fail(FVK_EINVAL, "%s: bad uniform value", what);
        if (ncomp) *ncomp = nc;
        if (isUniform) *isUniform = 1;
        if (uni) { uni[0] = v[0]; uni[1] = v[1]; uni[2] = v[2]; }
        if (out && nExpected > 0)
        {
            if (nExpected * nc > cap) return fvk_fail(FVK_EINVAL, "%s: output buffer too small", what);
            for (int64_t i = 0; i < nExpected; ++i)
                for (int d = 0; d < nc; ++d) out[i * nc + d] = v[d];
        }
        return FVK_OK;
    }
    if (kind != "nonuniform") return fvk_fail(FVK_EINVAL, "%s: expected 'uniform' or 'nonuniform', got '%s'", what, kind.c_str());
    const std::string list = word(p);
    int nc = 0;
    if (list == "List<scalar>") nc = 1;
    else if (list == "List<vector>") nc = 3;
    else return fvk_fail(FVK_EUNSUPPORTED, "%s: unsupported list type '%s'", what, list.c_str());
    double nd;
    if (!number(p, nd) || nd < 0) return fvk_fail(FVK_EINVAL, "%s: missing list size", what);
    const int64_t n = int64_t(nd);
    skip(p);
    if (p >= s.size() || s[p] != '(') return fvk_fail(FVK_EINVAL, "%s: expected '(' of the list", what);
    ++p;
    if (nExpected >= 0 && n != nExpected) return fvk_fail(FVK_EINVAL, "%s: list has %lld entries, expected %lld", what, (long long) n, (long long) nExpected);
    if (out && n * nc > cap) return fvk_fail(FVK_EINVAL, "%s: output buffer too small", what);
    for (int64_t i = 0; i < n; ++i)
    {
        skip(p);
        if (nc == 3) { if (p >= s.size() || s[p] != '(') return fvk_fail(FVK_EINVAL, "%s: expected '(' of a vector", what); ++p; }
        for (int d = 0; d < nc; ++d)
        {
            double v;
            if (!number(p, v)) return fvk_fail(FVK_EINVAL, "%s: bad number in entry %lld", what, (long long) i);
            if (out) out[i * nc + d] = v;
        }
        if (nc == 3) { skip(p); if (p >= s.size() || s[p] != ')') return fvk_fail(FVK_EINVAL, "%s: expected ')' of a vector", what); ++p; }
    }
    if (ncomp) *ncomp = nc;
    if (isUniform) *isUniform = 0;
    return FVK_OK;
}
} // namespace

extern "C" int fvk_fieldfile_read_internal(const char* path, int32_t nCells, int32_t* ncomp, double* out, int64_t outCapacity)
{
    if (!path || nCells < 0) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_internal: bad argument");
    std::string s, err;
    if (!slurp(path, s, err)) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_internal: %s", err.c_str());
    const size_t pos = findKey(s, "internalField", 0, s.size(), 0);
    if (pos == std::string::npos) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_internal: %s has no internalField", path);
    return parseValue(s, pos, path, nCells, ncomp, out, outCapacity, nullptr, nullptr);
}

extern "C" int fvk_fieldfile_read_patch(const char* path, const char* patchName, char* type, int32_t typeCap, int32_t nPatchFaces,
                                        int32_t* hasValue, int32_t* ncomp, double* out, int64_t outCapacity)
{
    if (!path || !patchName) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: bad argument");
    std::string s, err;
    if (!slurp(path, s, err)) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: %s", err.c_str());
    size_t bf = findKey(s, "boundaryField", 0, s.size(), 0);
    if (bf == std::string::npos) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: %s has no boundaryField", path);
    const size_t open = s.find('{', bf);
    if (open == std::string::npos) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: malformed boundaryField");
    int depth = 0;
    size_t close = open;
    for (; close < s.size(); ++close)
    {
        if (s[close] == '{') ++depth;
        if (s[close] == '}' && --depth == 0) break;
    }
    const size_t pk = findKey(s, patchName, open + 1, close, 0);
    if (pk == std::string::npos) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: %s has no patch '%s'", path, patchName);
    const size_t po = s.find('{', pk);
    if (po == std::string::npos || po >= close) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: patch '%s' has no dictionary", patchName);
    size_t pc = po;
    for (depth = 0; pc < close; ++pc)
    {
        if (s[pc] == '{') ++depth;
        if (s[pc] == '}' && --depth == 0) break;
    }
    const size_t tk = findKey(s, "type", po + 1, pc, 0);
    if (tk == std::string::npos) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_read_patch: patch '%s' has no type", patchName);
    size_t b = tk;
    while (b < pc && std::isspace(static_cast<unsigned char>(s[b]))) ++b;
    size_t e = b;
    while (e < pc && s[e] != ';' && !std::isspace(static_cast<unsigned char>(s[e]))) ++e;
    if (type && typeCap > 0) std::snprintf(type, size_t(typeCap), "%s", s.substr(b, e - b).c_str());
    if (hasValue) *hasValue = 0;
    const size_t vk = findKey(s, "value", po + 1, pc, 0);
    if (vk == std::string::npos) return FVK_OK;
    if (hasValue) *hasValue = 1;
    return parseValue(s, vk, path, nPatchFaces, ncomp, out, outCapacity, nullptr, nullptr);
}

// Writer: a vol<Scalar|Vector>Field file with a nonuniform internal field and one dictionary per patch
// (`type`, and `value uniform v` where patchHasValue[p] != 0). 17 significant digits: values round-trip exactly.
extern "C" int fvk_fieldfile_write(const char* path, const char* objectName, int32_t ncomp, int32_t nCells, const double* internal,
                                   int32_t nPatches, const char* const* patchNames, const char* const* patchTypes,
                                   const int32_t* patchHasValue, const double* patchValues /* [nPatches*ncomp] */)
{
    if (!path || !objectName || (ncomp != 1 && ncomp != 3) || nCells < 0 || (nCells && !internal) || nPatches < 0
        || (nPatches && (!patchNames || !patchTypes)))
        return fvk_fail(FVK_EINVAL, "fvk_fieldfile_write: bad argument");
    FILE* fp = std::fopen(path, "w");
    if (!fp) return fvk_fail(FVK_EINVAL, "fvk_fieldfile_write: cannot open %s", path);
    std::fprintf(fp, "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n    object      %s;\n}\n\n",
                 ncomp == 1 ? "volScalarField" : "volVectorField", objectName);
    std::fprintf(fp, "dimensions      [0 0 0 0 0 0 0];\n\ninternalField   nonuniform List<%s> %d\n(\n", ncomp == 1 ? "scalar" : "vector", nCells);
    for (int32_t c = 0; c < nCells; ++c)
    {
        if (ncomp == 1) std::fprintf(fp, "%.17g\n", internal[c]);
        else std::fprintf(fp, "(%.17g %.17g %.17g)\n", internal[3 * c], internal[3 * c + 1], internal[3 * c + 2]);
    }
    std::fprintf(fp, ")\n;\n\nboundaryField\n{\n");
    for (int32_t p = 0; p < nPatches; ++p)
    {
        std::fprintf(fp, "    %s\n    {\n        type            %s;\n", patchNames[p], patchTypes[p]);
        if (patchHasValue && patchHasValue[p] && patchValues)
        {
            if (ncomp == 1) std::fprintf(fp, "        value           uniform %.17g;\n", patchValues[p]);
            else std::fprintf(fp, "        value           uniform (%.17g %.17g %.17g);\n", patchValues[3 * p], patchValues[3 * p + 1], patchValues[3 * p + 2]);
        }
        std::fprintf(fp, "    }\n");
    }
    std::fprintf(fp, "}\n");
    std::fclose(fp);
    return FVK_OK;
}
