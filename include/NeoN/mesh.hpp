// NeoN::UnstructuredMesh / BoundaryMesh on the device (src/NeoN/include/NeoN/mesh/unstructured/{unstructuredMesh,
// boundaryMesh}.hpp) + the cached per-mesh data the reference keeps in its StencilDataBase: BasicGeometryScheme,
// CellToFaceStencil and la::SparsityPattern (all built once by fvk_mesh_create). Arrays are exposed as device Views
// in reference order. Mesh sources: the synthetic block mesh that stands in for blockMesh + readOpenFOAMMesh
// (FoamAdapter src/datastructures/meshAdapter.cpp:59-136), create1DUniformMesh, or a caller-filled fvk_mesh_desc.
#pragma once

#include "NeoN/core.hpp"

namespace NeoN
{
struct BlockPatch
{
    std::string name;
    std::vector<int> sides; // 0 x-min, 1 x-max, 2 y-min, 3 y-max, 4 z-min, 5 z-max
    bool isEmpty = false;
};

class UnstructuredMesh
{
    struct Impl
    {
        fvk_mesh* h = nullptr;
        fvk_mesh_desc* blockDesc = nullptr; // owned when created by the block generator
        fvk_mesh_desc* polyDesc = nullptr;  // owned when read from a polyMesh directory
        ~Impl()
        {
            if (h) fvk_mesh_destroy(h);
            if (blockDesc) fvk_blockmesh_destroy(blockDesc);
            if (polyDesc) fvk_polymesh_destroy(polyDesc);
        }
    };

public:
    // upload a host description (what readOpenFOAMMesh produces)
    UnstructuredMesh(const Executor& exec, const fvk_mesh_desc& desc, std::vector<std::string> patchNames = {})
        : exec_(exec), impl_(std::make_shared<Impl>()), patchNames_(std::move(patchNames))
    {
        check(fvk_mesh_create(&desc, &impl_->h));
        init();
    }
    // single hex block in OpenFOAM blockMesh ordering (see fvk_blockmesh_create)
    static UnstructuredMesh createBlockMesh(const Executor& exec, int nx, int ny, int nz, double lx, double ly, double lz,
                                            const std::vector<BlockPatch>& patches)
    {
        std::vector<int32_t> nSides, sides, empty;
        std::vector<std::string> names;
        for (const auto& p : patches)
        {
            nSides.push_back(int32_t(p.sides.size()));
            for (int s : p.sides) sides.push_back(s);
            empty.push_back(p.isEmpty ? 1 : 0);
            if (!p.isEmpty) names.push_back(p.name);
        }
        fvk_mesh_desc* d = nullptr;
        check(fvk_blockmesh_create(nx, ny, nz, lx, ly, lz, int32_t(patches.size()), nSides.data(), sides.data(), empty.data(), 0, &d));
        UnstructuredMesh m(exec, *d, names);
        m.impl_->blockDesc = d;
        return m;
    }

    // FoamAdapter::readOpenFOAMMesh (src/datastructures/meshAdapter.cpp:59-136) without OpenFOAM: an ASCII
    // constant/polyMesh directory (points, faces, owner, neighbour, boundary); `empty` patches are dropped
    static UnstructuredMesh readPolyMesh(const Executor& exec, const std::string& polyMeshDir)
    {
        fvk_mesh_desc* d = nullptr;
        check(fvk_polymesh_read(polyMeshDir.c_str(), &d));
        std::vector<std::string> names;
        char name[256], type[256];
        for (int32_t p = 0; p < d->nPatches; ++p)
        {
            check(fvk_polymesh_patch(d, p, name, sizeof name, type, sizeof type));
            names.emplace_back(name);
        }
        UnstructuredMesh m(exec, *d, names);
        m.impl_->polyDesc = d;
        return m;
    }

    const Executor& exec() const { return exec_; }
    const fvk_mesh* handle() const { return impl_->h; }
    localIdx nCells() const { return nCells_; }
    localIdx nOwnedCells() const { return nOwned_; }
    localIdx nInternalFaces() const { return nI_; }
    localIdx nBoundaryFaces() const { return nB_; }
    localIdx nFaces() const { return nI_ + nB_; }
    localIdx nBoundaries() const { return localIdx(offsets_.size()) - 1; }
    int64_t nnz() const { return nnz_; }
    const std::vector<localIdx>& boundaryOffsets() const { return offsets_; } // BoundaryMesh::offset()
    const std::vector<std::string>& patchNames() const { return patchNames_; }

    View<const scalar> cellVolumes() const { return arr<scalar>(FVK_CELL_VOLUMES); }
    View<const Vec3> cellCentres() const { return arr3(FVK_CELL_CENTRES); }
    View<const Vec3> faceAreas() const { return arr3(FVK_FACE_AREAS); }
    View<const Vec3> faceCentres() const { return arr3(FVK_FACE_CENTRES); }
    View<const scalar> magFaceAreas() const { return arr<scalar>(FVK_MAG_FACE_AREAS); }
    View<const localIdx> faceOwner() const { return arr<localIdx>(FVK_FACE_OWNER); }
    View<const localIdx> faceNeighbour() const { return arr<localIdx>(FVK_FACE_NEIGHBOUR); }
    View<const localIdx> faceCells() const { return arr<localIdx>(FVK_FACE_CELLS); }
    // BasicGeometryScheme (stencil/basicGeometryScheme.cpp)
    View<const scalar> weights() const { return arr<scalar>(FVK_WEIGHTS); }
    View<const scalar> deltaCoeffs() const { return arr<scalar>(FVK_DELTACOEFFS); }
    View<const scalar> nonOrthDeltaCoeffs() const { return arr<scalar>(FVK_NONORTH_DELTACOEFFS); }

    template<typename T> View<const T> arr(int field) const
    {
        const void* p = nullptr;
        int64_t n = 0;
        check(fvk_mesh_array(impl_->h, field, &p, &n));
        return {static_cast<const T*>(p), size_t(n)};
    }
    View<const Vec3> arr3(int field) const
    {
        auto v = arr<scalar>(field);
        return {reinterpret_cast<const Vec3*>(v.ptr), v.n / 3};
    }

private:
    void init()
    {
        int64_t v = 0;
        check(fvk_mesh_size(impl_->h, FVK_N_CELLS, &v)); nCells_ = localIdx(v);
        check(fvk_mesh_size(impl_->h, FVK_N_OWNED_CELLS, &v)); nOwned_ = localIdx(v);
        check(fvk_mesh_size(impl_->h, FVK_N_INTERNAL_FACES, &v)); nI_ = localIdx(v);
        check(fvk_mesh_size(impl_->h, FVK_N_BOUNDARY_FACES, &v)); nB_ = localIdx(v);
        check(fvk_mesh_size(impl_->h, FVK_NNZ, &v)); nnz_ = v;
        check(fvk_mesh_size(impl_->h, FVK_N_PATCHES, &v));
        offsets_.assign(size_t(v) + 1, 0);
        check(fvk_mesh_patch_offsets(impl_->h, offsets_.data()));
        while (patchNames_.size() < size_t(v)) patchNames_.push_back("patch" + std::to_string(patchNames_.size()));
    }
    Executor exec_;
    std::shared_ptr<Impl> impl_;
    std::vector<std::string> patchNames_;
    std::vector<localIdx> offsets_;
    localIdx nCells_ = 0, nOwned_ = 0, nI_ = 0, nB_ = 0;
    int64_t nnz_ = 0;
};

// NeoN::create1DUniformMesh (src/NeoN/src/mesh/unstructured/unstructuredMesh.cpp:112-220)
inline UnstructuredMesh create1DUniformMesh(const Executor& exec, localIdx nCells)
{
    const localIdx n = nCells;
    const scalar h = (1.0 - 0.0) / scalar(n);
    std::vector<double> V(n, h), C(3 * n, 0.0), Sf(3 * (n + 1), 0.0), Cf(3 * (n + 1), 0.0), magSf(n + 1, 1.0);
    std::vector<int32_t> owner(n + 1), nei(n > 1 ? n - 1 : 0), faceCells {0, n - 1}, off {0, 1, 2};
    for (localIdx i = 0; i < n; ++i) C[3 * i] = 0.5 * h + h * scalar(i);
    for (localIdx i = 0; i < n - 1; ++i) { Cf[3 * i] = 0.0 + (scalar(i) + 1.0) * h; Sf[3 * i] = 1.0; owner[i] = i; nei[i] = i + 1; }
    Cf[3 * (n - 1)] = 0.0; Sf[3 * (n - 1)] = -1.0; owner[n - 1] = 0;
    Cf[3 * n] = 1.0; Sf[3 * n] = 1.0; owner[n] = n - 1;
    std::vector<double> bCf(Cf.begin() + 3 * (n - 1), Cf.end()), bSf(Sf.begin() + 3 * (n - 1), Sf.end()), bCn {C[0], 0, 0, C[3 * (n - 1)], 0, 0};
    std::vector<double> bDelta {0.0 - C[0], 0, 0, 1.0 - C[3 * (n - 1)], 0, 0}, ones {1.0, 1.0};
    std::vector<double> bDc {1.0 / std::abs(bDelta[0]), 1.0 / std::abs(bDelta[3])};
    fvk_mesh_desc d {};
    d.nCells = n; d.nInternalFaces = n - 1; d.nBoundaryFaces = 2; d.nPatches = 2;
    d.cellVolumes = V.data(); d.cellCentres = C.data(); d.faceAreas = Sf.data(); d.faceCentres = Cf.data(); d.magFaceAreas = magSf.data();
    d.faceOwner = owner.data(); d.faceNeighbour = nei.data(); d.faceCells = faceCells.data();
    d.bCf = bCf.data(); d.bCn = bCn.data(); d.bSf = bSf.data(); d.bMagSf = ones.data(); d.bNf = bSf.data(); d.bDelta = bDelta.data();
    d.bWeights = ones.data(); d.bDeltaCoeffs = bDc.data(); d.patchOffsets = off.data();
    return UnstructuredMesh(exec, d, {"left", "right"});
}

} // namespace NeoN
