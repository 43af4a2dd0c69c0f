// NeoN::finiteVolume::cellCentred for the B200 build: fields with boundary data, boundary conditions, surface
// interpolation, face-normal gradient and the Gauss-Green operators -- the reference's class and method names
// (src/NeoN/include/NeoN/finiteVolume/cellCentred/{fields,boundary,interpolation,faceNormalGradient,operators}/*.hpp,
// src/NeoN/include/NeoN/fields/boundaryData.hpp) with bodies that call the fused CUDA kernels of libfvk.
#pragma once

#include "NeoN/core.hpp"
#include "NeoN/mesh.hpp"

namespace NeoN
{
namespace dsl
{
// dsl::Coeff (dsl/coeff.hpp:21-54): scalar x optional per-cell view, evaluated inside the kernels
class Coeff
{
public:
    Coeff() : coeff_(1.0) {}
    Coeff(scalar v) : coeff_(v) {}
    Coeff(scalar c, const Vector<scalar>& field) : coeff_(c), view_(field.data()), hasView_(true) {}
    explicit Coeff(const Vector<scalar>& field) : coeff_(1.0), view_(field.data()), hasView_(true) {}
    bool hasView() const { return hasView_; }
    const scalar* view() const { return hasView_ ? view_ : nullptr; }
    scalar value() const { return coeff_; }
    Coeff& operator*=(scalar r) { coeff_ *= r; return *this; }
    Coeff& operator*=(const Coeff& r)
    {
        if (hasView_ && r.hasView_) NF_ERROR_EXIT("Not implemented");
        if (!hasView_ && r.hasView_) { view_ = r.view_; hasView_ = true; }
        coeff_ *= r.coeff_;
        return *this;
    }
private:
    scalar coeff_;
    const scalar* view_ = nullptr;
    bool hasView_ = false;
};
inline Coeff operator*(const Coeff& l, const Coeff& r) { Coeff c = l; c *= r; return c; }
} // namespace dsl

// fields/boundaryData.hpp:32-215
template<typename T>
class BoundaryData
{
public:
    BoundaryData(const Executor& exec, const std::vector<localIdx>& offsets)
        : value_(exec, size_t(offsets.back()), zero<T>()), refValue_(exec, size_t(offsets.back()), zero<T>()),
          valueFraction_(exec, size_t(offsets.back()), 0.0), refGrad_(exec, size_t(offsets.back()), zero<T>()), offset_(offsets) {}
    Vector<T>& value() { return value_; }
    const Vector<T>& value() const { return value_; }
    Vector<T>& refValue() { return refValue_; }
    const Vector<T>& refValue() const { return refValue_; }
    Vector<scalar>& valueFraction() { return valueFraction_; }
    const Vector<scalar>& valueFraction() const { return valueFraction_; }
    Vector<T>& refGrad() { return refGrad_; }
    const Vector<T>& refGrad() const { return refGrad_; }
    const std::vector<localIdx>& offset() const { return offset_; }
    localIdx nBoundaries() const { return localIdx(offset_.size()) - 1; }
    localIdx nBoundaryFaces() const { return offset_.back(); }
    std::pair<localIdx, localIdx> range(localIdx patch) const { return {offset_[patch], offset_[patch + 1]}; }
    fvk_bfield c() const { return {value_.raw(), refValue_.raw(), valueFraction_.data(), refGrad_.raw()}; }
private:
    Vector<T> value_, refValue_;
    Vector<scalar> valueFraction_;
    Vector<T> refGrad_;
    std::vector<localIdx> offset_;
};

namespace finiteVolume::cellCentred
{
// VolumeBoundary<T> (boundary/volumeBoundaryFactory.hpp + boundary/volume/*.hpp): type by name like the
// RuntimeSelectionFactory keys of the reference: fixedValue | fixedGradient | calculated | extrapolated | empty;
// zeroGradient and noSlip are mapped as FoamAdapter's reader does (include/FoamAdapter/auxiliary/readers.hpp:43-95).
template<typename T>
struct VolumeBoundary
{
    std::string type;
    T constant;
    VolumeBoundary(std::string t = "calculated", T c = zero<T>()) : type(std::move(t)), constant(c)
    {
        if (type == "zeroGradient") { type = "fixedGradient"; constant = zero<T>(); }
        if (type == "noSlip") { type = "fixedValue"; constant = zero<T>(); }
        if (kind() < 0) NF_ERROR_EXIT("unknown boundary condition type: " + type); // keyExistsOrError
    }
    int kind() const
    {
        if (type == "calculated") return FVK_BC_CALCULATED;
        if (type == "fixedValue") return FVK_BC_FIXED_VALUE;
        if (type == "fixedGradient") return FVK_BC_FIXED_GRADIENT;
        if (type == "extrapolated") return FVK_BC_EXTRAPOLATED;
        if (type == "empty") return FVK_BC_EMPTY;
        return -1;
    }
    bool assignable() const { return kind() != FVK_BC_FIXED_VALUE; } // fixedValue.hpp:56
};
template<typename T> std::vector<VolumeBoundary<T>> createCalculatedBCs(const UnstructuredMesh& m)
{
    return std::vector<VolumeBoundary<T>>(size_t(m.nBoundaries()), VolumeBoundary<T>("calculated"));
}
template<typename T> std::vector<VolumeBoundary<T>> createExtrapolatedBCs(const UnstructuredMesh& m)
{
    return std::vector<VolumeBoundary<T>>(size_t(m.nBoundaries()), VolumeBoundary<T>("extrapolated"));
}

// fields/volumeField.hpp
template<typename T>
class VolumeField
{
public:
    using ElementType = T;
    VolumeField(const Executor& exec, std::string fieldName, const UnstructuredMesh& mesh, std::vector<VolumeBoundary<T>> bcs)
        : name(std::move(fieldName)), exec_(exec), mesh_(mesh), internal_(exec, size_t(mesh.nCells()), zero<T>()),
          boundary_(exec, mesh.boundaryOffsets()), bcs_(std::move(bcs))
    {
        if (localIdx(bcs_.size()) != mesh.nBoundaries()) NF_ERROR_EXIT("VolumeField " + name + ": one boundary condition per patch required");
    }
    std::string name;
    const Executor& exec() const { return exec_; }
    const UnstructuredMesh& mesh() const { return mesh_; }
    Vector<T>& internalVector() { return internal_; }
    const Vector<T>& internalVector() const { return internal_; }
    BoundaryData<T>& boundaryData() { return boundary_; }
    const BoundaryData<T>& boundaryData() const { return boundary_; }
    const std::vector<VolumeBoundary<T>>& boundaryConditions() const { return bcs_; }
    // one launch for all patches (the reference launches one kernel per patch)
    void correctBoundaryConditions()
    {
        std::vector<int32_t> kinds;
        std::vector<double> consts;
        for (const auto& b : bcs_)
        {
            kinds.push_back(b.kind());
            if constexpr (std::is_same_v<T, Vec3>) { consts.push_back(b.constant[0]); consts.push_back(b.constant[1]); consts.push_back(b.constant[2]); }
            else consts.push_back(b.constant);
        }
        check(fvk_correct_boundary_conditions(mesh_.handle(), nComponents<T>(), kinds.data(), consts.data(), internal_.raw(),
                                              boundary_.value().raw(), boundary_.refValue().raw(), boundary_.valueFraction().data(),
                                              boundary_.refGrad().raw(), exec_.stream()));
    }
    // fvcc::oldTime(field) (core/database/oldTimeCollection.hpp:146-152): a registered copy created on first use
    VolumeField& oldTime()
    {
        if (!old_) { old_ = std::make_shared<VolumeField>(exec_, name + "_0", mesh_, bcs_); old_->internal_ = internal_; }
        return *old_;
    }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
    Vector<T> internal_;
    BoundaryData<T> boundary_;
    std::vector<VolumeBoundary<T>> bcs_;
    std::shared_ptr<VolumeField> old_;
};
template<typename T> VolumeField<T>& oldTime(VolumeField<T>& f) { return f.oldTime(); }

// fields/surfaceField.hpp:38-53: internalVector holds nInternalFaces + nBoundaryFaces values
template<typename T>
class SurfaceField
{
public:
    using ElementType = T;
    SurfaceField(const Executor& exec, std::string fieldName, const UnstructuredMesh& mesh)
        : name(std::move(fieldName)), exec_(exec), mesh_(mesh), internal_(exec, size_t(mesh.nFaces()), zero<T>()),
          boundary_(exec, mesh.boundaryOffsets()) {}
    std::string name;
    const Executor& exec() const { return exec_; }
    const UnstructuredMesh& mesh() const { return mesh_; }
    Vector<T>& internalVector() { return internal_; }
    const Vector<T>& internalVector() const { return internal_; }
    BoundaryData<T>& boundaryData() { return boundary_; }
    const BoundaryData<T>& boundaryData() const { return boundary_; }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
    Vector<T> internal_;
    BoundaryData<T> boundary_;
};

namespace detail
{
inline int scheme(const std::string& name)
{
    if (name == "linear") return FVK_LINEAR;
    if (name == "upwind") return FVK_UPWIND;
    NF_ERROR_EXIT("unknown interpolation scheme: " + name); // RuntimeSelectionFactory::keyExistsOrError
    return -1;
}
}

// interpolation/surfaceInterpolation.hpp:54-69 (keys "linear" | "upwind")
template<typename T>
class SurfaceInterpolation
{
public:
    SurfaceInterpolation(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
        : exec_(exec), mesh_(mesh), scheme_(detail::scheme(input[0])) {}
    int scheme() const { return scheme_; }
    void interpolate(const VolumeField<T>& src, SurfaceField<T>& dst) const
    {
        if (scheme_ == FVK_UPWIND) NF_ERROR_EXIT("limited scheme require a faceFlux"); // upwind.hpp:66-72
        run(nullptr, src, dst);
    }
    void interpolate(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<T>& dst) const { run(flux.internalVector().data(), src, dst); }
    SurfaceField<T> interpolate(const VolumeField<T>& src) const
    {
        SurfaceField<T> dst(exec_, "interpolated_" + src.name, mesh_);
        interpolate(src, dst);
        return dst;
    }
    void weight(const VolumeField<T>&, SurfaceField<scalar>& w) const
    {
        if (scheme_ == FVK_UPWIND) NF_ERROR_EXIT("limited scheme require a faceFlux");
        check(fvk_interpolation_weights(mesh_.handle(), scheme_, nullptr, w.internalVector().data(), w.boundaryData().value().data(), exec_.stream()));
    }
    void weight(const SurfaceField<scalar>& flux, const VolumeField<T>&, SurfaceField<scalar>& w) const
    {
        check(fvk_interpolation_weights(mesh_.handle(), scheme_, flux.internalVector().data(), w.internalVector().data(),
                                        w.boundaryData().value().data(), exec_.stream()));
    }
private:
    void run(const scalar* flux, const VolumeField<T>& src, SurfaceField<T>& dst) const
    {
        if constexpr (std::is_same_v<T, Vec3>)
            check(fvk_interpolate_v(mesh_.handle(), scheme_, flux, src.internalVector().raw(), src.boundaryData().value().raw(), dst.internalVector().raw(), exec_.stream()));
        else
            check(fvk_interpolate_s(mesh_.handle(), scheme_, flux, src.internalVector().raw(), src.boundaryData().value().raw(), dst.internalVector().raw(), exec_.stream()));
    }
    Executor exec_;
    UnstructuredMesh mesh_;
    int scheme_;
};

// faceNormalGradient/faceNormalGradient.hpp:50-54 (key "uncorrected")
template<typename T>
class FaceNormalGradient
{
public:
    FaceNormalGradient(const Executor& exec, const UnstructuredMesh& mesh, const Input& input) : exec_(exec), mesh_(mesh)
    {
        if (input[0] != "uncorrected") NF_ERROR_EXIT("unknown faceNormalGradient scheme: " + input[0]);
    }
    void faceNormalGrad(const VolumeField<T>& phi, SurfaceField<T>& out) const
    {
        if constexpr (std::is_same_v<T, Vec3>)
            check(fvk_face_normal_grad_v(mesh_.handle(), phi.internalVector().raw(), phi.boundaryData().value().raw(), out.internalVector().raw(), exec_.stream()));
        else
            check(fvk_face_normal_grad_s(mesh_.handle(), phi.internalVector().raw(), phi.boundaryData().value().raw(), out.internalVector().raw(), exec_.stream()));
    }
    View<const scalar> deltaCoeffs() const { return mesh_.nonOrthDeltaCoeffs(); } // uncorrected.hpp:55-58
private:
    Executor exec_;
    UnstructuredMesh mesh_;
};

// operators/gaussGreenDiv.hpp:75-81 -- accumulate into divPhi, then scale all of it by coeff/V (computeDiv)
template<typename T>
class GaussGreenDiv
{
public:
    GaussGreenDiv(const Executor& exec, const UnstructuredMesh& mesh, const Input& input) : exec_(exec), mesh_(mesh), scheme_(detail::scheme(input[0])) {}
    int scheme() const { return scheme_; }
    void div(Vector<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os, int mode = FVK_ACC_SCALE) const
    {
        auto fn = std::is_same_v<T, Vec3> ? fvk_div_v : fvk_div_s;
        check(fn(mesh_.handle(), scheme_, faceFlux.internalVector().data(), phi.internalVector().raw(), phi.boundaryData().value().raw(),
                 os.value(), os.view(), divPhi.raw(), mode, exec_.stream()));
    }
    void div(VolumeField<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const
    {
        div(divPhi.internalVector(), faceFlux, phi, os);
    }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
    int scheme_;
};

// operators/gaussGreenGrad.hpp -- always linear interpolation, scale 1/V
class GaussGreenGrad
{
public:
    GaussGreenGrad(const Executor& exec, const UnstructuredMesh& mesh) : exec_(exec), mesh_(mesh) {}
    void grad(const VolumeField<scalar>& phi, const dsl::Coeff, Vector<Vec3>& gradPhi) const { run(phi, gradPhi, FVK_ACC_SCALE); }
    void grad(const VolumeField<scalar>& phi, VolumeField<Vec3>& gradPhi) const { run(phi, gradPhi.internalVector(), FVK_ACC_SCALE); }
    VolumeField<Vec3> grad(const VolumeField<scalar>& phi) const
    {
        VolumeField<Vec3> g(exec_, "grad_" + phi.name, mesh_, createCalculatedBCs<Vec3>(mesh_));
        run(phi, g.internalVector(), FVK_SET);
        return g;
    }
    void run(const VolumeField<scalar>& phi, Vector<Vec3>& out, int mode) const
    {
        check(fvk_grad_s(mesh_.handle(), phi.internalVector().data(), phi.boundaryData().value().data(), out.raw(), mode, exec_.stream()));
    }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
};

// operators/gaussGreenLaplacian.hpp -- explicit: gamma is ignored like the reference (gaussGreenLaplacian.cpp:14)
template<typename T>
class GaussGreenLaplacian
{
public:
    GaussGreenLaplacian(const Executor& exec, const UnstructuredMesh& mesh, const Input& input) : exec_(exec), mesh_(mesh)
    {
        // "linear uncorrected": interpolation of gamma, face-normal gradient scheme
        const std::string fng = input.size() > 1 ? input[1] : input[0];
        if (fng != "uncorrected") NF_ERROR_EXIT("unknown faceNormalGradient scheme: " + fng);
    }
    void laplacian(Vector<T>& lapPhi, const SurfaceField<scalar>&, const VolumeField<T>& phi, const dsl::Coeff os, int mode = FVK_ACC_SCALE) const
    {
        auto fn = std::is_same_v<T, Vec3> ? fvk_laplacian_v : fvk_laplacian_s;
        check(fn(mesh_.handle(), phi.internalVector().raw(), phi.boundaryData().value().raw(), os.value(), os.view(), lapPhi.raw(), mode, exec_.stream()));
    }
    void laplacian(VolumeField<T>& lapPhi, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const
    {
        laplacian(lapPhi.internalVector(), gamma, phi, os);
    }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
};

// auxiliary/coNum.cpp:18-96 -> maxCoNum (device -> host scalar like the reference)
inline scalar computeCoNum(const SurfaceField<scalar>& faceFlux, scalar dt)
{
    const auto& m = faceFlux.mesh();
    Vector<scalar> res(faceFlux.exec(), 2);
    Vector<scalar> scratch(faceFlux.exec(), fvk_conum_scratch_bytes(m.handle()) / sizeof(double));
    check(fvk_conum(m.handle(), faceFlux.internalVector().data(), dt, res.data(), scratch.data(), faceFlux.exec().stream()));
    auto h = res.copyToHost();
    std::cout << "Courant Number mean: " << h[1] << " max: " << h[0] << std::endl; // coNum.cpp:91-93
    return h[0];
}

} // namespace finiteVolume::cellCentred
} // namespace NeoN
