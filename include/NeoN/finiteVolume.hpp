// NeoN::finiteVolume::cellCentred for the B200 build: fields with boundary data, boundary conditions, surface
// interpolation, face-normal gradient and the Gauss-Green operators -- the reference's class and method names
// (src/NeoN/include/NeoN/finiteVolume/cellCentred/{fields,boundary,interpolation,faceNormalGradient,operators}/*.hpp,
// src/NeoN/include/NeoN/fields/boundaryData.hpp) with bodies that call the fused CUDA kernels of libfvk.
#pragma once

#include "NeoN/core.hpp"
#include "NeoN/mesh.hpp"
#include "NeoN/linearAlgebra.hpp"

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
inline Input tail(const Input& in)
{ // the tokens after the selection key (TokenList::popFront of the reference's create())
    std::vector<std::string> t;
    for (size_t i = 1; i < in.size(); ++i) t.push_back(in[i]);
    return Input(t);
}
}

// ---- interpolation/surfaceInterpolation.hpp:27-140: strategies selected BY NAME ("linear", "upwind", or a plug-in) ------------
template<typename T>
class SurfaceInterpolationFactory
    : public RuntimeSelectionFactory<SurfaceInterpolationFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>
{
public:
    using Factory = RuntimeSelectionFactory<SurfaceInterpolationFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>;
    static std::unique_ptr<SurfaceInterpolationFactory<T>> create(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
    {
        if (input.empty()) NF_ERROR_EXIT("SurfaceInterpolation: empty token list");
        return Factory::create(input[0], exec, mesh, detail::tail(input));
    }
    static std::string name() { return "SurfaceInterpolationFactory"; }
    SurfaceInterpolationFactory(const Executor& exec, const UnstructuredMesh& mesh) : exec_(exec), mesh_(mesh) {}
    virtual ~SurfaceInterpolationFactory() = default;
    virtual void interpolate(const VolumeField<T>& src, SurfaceField<T>& dst) const = 0;
    virtual void interpolate(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<T>& dst) const = 0;
    virtual void weight(const VolumeField<T>& src, SurfaceField<scalar>& w) const = 0;
    virtual void weight(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<scalar>& w) const = 0;
    virtual std::unique_ptr<SurfaceInterpolationFactory<T>> clone() const = 0;
    // hot-path hook (new here): the scheme code the fused operator kernels understand (FVK_LINEAR / FVK_UPWIND); a plug-in
    // scheme returns -1 and the operators fall back to interpolate() + surfaceIntegrate
    virtual int fusedScheme() const { return -1; }
protected:
    Executor exec_;
    UnstructuredMesh mesh_;
};

namespace detail
{
template<typename T>
void runInterpolate(const Executor& exec, const UnstructuredMesh& mesh, int scheme, const scalar* flux, const VolumeField<T>& src, SurfaceField<T>& dst)
{
    auto fn = std::is_same_v<T, Vec3> ? fvk_interpolate_v : fvk_interpolate_s;
    check(fn(mesh.handle(), scheme, flux, src.internalVector().raw(), src.boundaryData().value().raw(), dst.internalVector().raw(), exec.stream()));
}
}

// interpolation/linear.hpp
template<typename T>
class Linear : public SurfaceInterpolationFactory<T>::template Register<Linear<T>>
{
    using Base = typename SurfaceInterpolationFactory<T>::template Register<Linear<T>>;
public:
    Linear(const Executor& exec, const UnstructuredMesh& mesh, const Input&) : Base(exec, mesh) {}
    static std::string name() { return "linear"; }
    static std::string doc() { return "linear interpolation"; }
    static std::string schema() { return "none"; }
    int fusedScheme() const override { return FVK_LINEAR; }
    void interpolate(const VolumeField<T>& src, SurfaceField<T>& dst) const override { detail::runInterpolate(this->exec_, this->mesh_, FVK_LINEAR, nullptr, src, dst); }
    void interpolate(const SurfaceField<scalar>&, const VolumeField<T>& src, SurfaceField<T>& dst) const override { interpolate(src, dst); }
    void weight(const VolumeField<T>&, SurfaceField<scalar>& w) const override
    {
        check(fvk_interpolation_weights(this->mesh_.handle(), FVK_LINEAR, nullptr, w.internalVector().data(), w.boundaryData().value().data(), this->exec_.stream()));
    }
    void weight(const SurfaceField<scalar>&, const VolumeField<T>& src, SurfaceField<scalar>& w) const override { weight(src, w); }
    std::unique_ptr<SurfaceInterpolationFactory<T>> clone() const override { return std::make_unique<Linear<T>>(*this); }
};
// interpolation/upwind.hpp
template<typename T>
class Upwind : public SurfaceInterpolationFactory<T>::template Register<Upwind<T>>
{
    using Base = typename SurfaceInterpolationFactory<T>::template Register<Upwind<T>>;
public:
    Upwind(const Executor& exec, const UnstructuredMesh& mesh, const Input&) : Base(exec, mesh) {}
    static std::string name() { return "upwind"; }
    static std::string doc() { return "upwind interpolation"; }
    static std::string schema() { return "none"; }
    int fusedScheme() const override { return FVK_UPWIND; }
    void interpolate(const VolumeField<T>&, SurfaceField<T>&) const override { NF_ERROR_EXIT("limited scheme require a faceFlux"); } // upwind.hpp:66-72
    void interpolate(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<T>& dst) const override
    {
        detail::runInterpolate(this->exec_, this->mesh_, FVK_UPWIND, flux.internalVector().data(), src, dst);
    }
    void weight(const VolumeField<T>&, SurfaceField<scalar>&) const override { NF_ERROR_EXIT("limited scheme require a faceFlux"); }
    void weight(const SurfaceField<scalar>& flux, const VolumeField<T>&, SurfaceField<scalar>& w) const override
    {
        check(fvk_interpolation_weights(this->mesh_.handle(), FVK_UPWIND, flux.internalVector().data(), w.internalVector().data(),
                                        w.boundaryData().value().data(), this->exec_.stream()));
    }
    std::unique_ptr<SurfaceInterpolationFactory<T>> clone() const override { return std::make_unique<Upwind<T>>(*this); }
};
NF_REGISTER((SurfaceInterpolationFactory<scalar>), (Linear<scalar>));
NF_REGISTER((SurfaceInterpolationFactory<Vec3>), (Linear<Vec3>));
NF_REGISTER((SurfaceInterpolationFactory<scalar>), (Upwind<scalar>));
NF_REGISTER((SurfaceInterpolationFactory<Vec3>), (Upwind<Vec3>));

// interpolation/surfaceInterpolation.hpp:74-140: the value-semantic front end owning a cloned strategy
template<typename T>
class SurfaceInterpolation
{
public:
    SurfaceInterpolation(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
        : exec_(exec), mesh_(mesh), strategy_(SurfaceInterpolationFactory<T>::create(exec, mesh, input)) {}
    SurfaceInterpolation(const SurfaceInterpolation& o) : exec_(o.exec_), mesh_(o.mesh_), strategy_(o.strategy_->clone()) {}
    SurfaceInterpolation(SurfaceInterpolation&&) = default;
    int scheme() const { return strategy_->fusedScheme(); }
    void interpolate(const VolumeField<T>& src, SurfaceField<T>& dst) const { strategy_->interpolate(src, dst); }
    void interpolate(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<T>& dst) const { strategy_->interpolate(flux, src, dst); }
    SurfaceField<T> interpolate(const VolumeField<T>& src) const
    {
        SurfaceField<T> dst(exec_, "interpolated_" + src.name, mesh_);
        interpolate(src, dst);
        return dst;
    }
    void weight(const VolumeField<T>& src, SurfaceField<scalar>& w) const { strategy_->weight(src, w); }
    void weight(const SurfaceField<scalar>& flux, const VolumeField<T>& src, SurfaceField<scalar>& w) const { strategy_->weight(flux, src, w); }
private:
    Executor exec_;
    UnstructuredMesh mesh_;
    std::unique_ptr<SurfaceInterpolationFactory<T>> strategy_;
};

// ---- faceNormalGradient/faceNormalGradient.hpp:27-100 (key "uncorrected") ----------------------------------------------------
template<typename T>
class FaceNormalGradientFactory : public RuntimeSelectionFactory<FaceNormalGradientFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>
{
public:
    using Factory = RuntimeSelectionFactory<FaceNormalGradientFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>;
    static std::unique_ptr<FaceNormalGradientFactory<T>> create(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
    {
        if (input.empty()) NF_ERROR_EXIT("FaceNormalGradient: empty token list");
        return Factory::create(input[0], exec, mesh, detail::tail(input));
    }
    static std::string name() { return "FaceNormalGradientFactory"; }
    FaceNormalGradientFactory(const Executor& exec, const UnstructuredMesh& mesh) : exec_(exec), mesh_(mesh) {}
    virtual ~FaceNormalGradientFactory() = default;
    virtual void faceNormalGrad(const VolumeField<T>& phi, SurfaceField<T>& out) const = 0;
    virtual View<const scalar> deltaCoeffs() const = 0;
    virtual std::unique_ptr<FaceNormalGradientFactory<T>> clone() const = 0;
    virtual bool fused() const { return false; } // true: the fused laplacian kernels implement this scheme
protected:
    Executor exec_;
    UnstructuredMesh mesh_;
};
template<typename T>
class Uncorrected : public FaceNormalGradientFactory<T>::template Register<Uncorrected<T>>
{
    using Base = typename FaceNormalGradientFactory<T>::template Register<Uncorrected<T>>;
public:
    Uncorrected(const Executor& exec, const UnstructuredMesh& mesh, const Input&) : Base(exec, mesh) {}
    static std::string name() { return "uncorrected"; }
    static std::string doc() { return "Uncorrected interpolation"; }
    static std::string schema() { return "none"; }
    bool fused() const override { return true; }
    void faceNormalGrad(const VolumeField<T>& phi, SurfaceField<T>& out) const override
    {
        auto fn = std::is_same_v<T, Vec3> ? fvk_face_normal_grad_v : fvk_face_normal_grad_s;
        check(fn(this->mesh_.handle(), phi.internalVector().raw(), phi.boundaryData().value().raw(), out.internalVector().raw(), this->exec_.stream()));
    }
    View<const scalar> deltaCoeffs() const override { return this->mesh_.nonOrthDeltaCoeffs(); } // uncorrected.hpp:55-58
    std::unique_ptr<FaceNormalGradientFactory<T>> clone() const override { return std::make_unique<Uncorrected<T>>(*this); }
};
NF_REGISTER((FaceNormalGradientFactory<scalar>), (Uncorrected<scalar>));
NF_REGISTER((FaceNormalGradientFactory<Vec3>), (Uncorrected<Vec3>));

template<typename T>
class FaceNormalGradient
{
public:
    FaceNormalGradient(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
        : strategy_(FaceNormalGradientFactory<T>::create(exec, mesh, input)) {}
    FaceNormalGradient(const FaceNormalGradient& o) : strategy_(o.strategy_->clone()) {}
    void faceNormalGrad(const VolumeField<T>& phi, SurfaceField<T>& out) const { strategy_->faceNormalGrad(phi, out); }
    View<const scalar> deltaCoeffs() const { return strategy_->deltaCoeffs(); }
    bool fused() const { return strategy_->fused(); }
private:
    std::unique_ptr<FaceNormalGradientFactory<T>> strategy_;
};

// ---- operators/divOperator.hpp:27-100: DivOperatorFactory, key "Gauss" --------------------------------------------------------
template<typename T>
class DivOperatorFactory : public RuntimeSelectionFactory<DivOperatorFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>
{
public:
    using Factory = RuntimeSelectionFactory<DivOperatorFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>;
    static std::unique_ptr<DivOperatorFactory<T>> create(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
    {
        if (input.empty()) NF_ERROR_EXIT("DivOperator: empty token list");
        return Factory::create(input[0], exec, mesh, detail::tail(input));
    }
    static std::string name() { return "DivOperatorFactory"; }
    DivOperatorFactory(const Executor& exec, const UnstructuredMesh& mesh) : exec_(exec), mesh_(mesh) {}
    virtual ~DivOperatorFactory() = default;
    virtual void div(Vector<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    virtual void div(VolumeField<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    // implicit: adds the operator's coefficients to an existing system (divOperator.hpp:56-61)
    virtual void div(la::LinearSystem<T, localIdx>& ls, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    virtual std::unique_ptr<DivOperatorFactory<T>> clone() const = 0;
    // hot-path hooks (new here): explicit operator as ONE ADD-mode launch; implicit operator as a term of the fused assembly
    virtual bool addTo(Vector<T>&, const SurfaceField<scalar>&, const VolumeField<T>&, const dsl::Coeff) const { return false; }
    virtual bool fusedTerm(fvk_term&, const SurfaceField<scalar>&, const dsl::Coeff) const { return false; }
protected:
    Executor exec_;
    UnstructuredMesh mesh_;
};

// operators/gaussGreenDiv.hpp:75-81 -- accumulate into divPhi, then scale all of it by coeff/V (computeDiv)
template<typename T>
class GaussGreenDiv : public DivOperatorFactory<T>::template Register<GaussGreenDiv<T>>
{
    using Base = typename DivOperatorFactory<T>::template Register<GaussGreenDiv<T>>;
public:
    GaussGreenDiv(const Executor& exec, const UnstructuredMesh& mesh, const Input& input) : Base(exec, mesh), interp_(exec, mesh, input) {}
    static std::string name() { return "Gauss"; }
    static std::string doc() { return "Gauss-Green Divergence"; }
    static std::string schema() { return "none"; }
    int scheme() const { return interp_.scheme(); }
    void div(Vector<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os, int mode) const
    {
        if (scheme() < 0)
        { // plug-in interpolation scheme: the reference's three steps (interpolate, multiply by the flux, surfaceIntegrate)
            SurfaceField<T> phif(this->exec_, "phif", this->mesh_);
            interp_.interpolate(faceFlux, phi, phif);
            if constexpr (std::is_same_v<T, scalar>)
            {
                check(fvk_vec_mul(int64_t(phif.internalVector().size()), phif.internalVector().raw(), faceFlux.internalVector().data(), this->exec_.stream()));
                check(fvk_surface_integrate_s(this->mesh_.handle(), phif.internalVector().raw(), os.value(), os.view(), divPhi.raw(), mode, this->exec_.stream()));
            }
            else NF_ERROR_EXIT("plug-in interpolation schemes are supported for scalar fields");
            return;
        }
        auto fn = std::is_same_v<T, Vec3> ? fvk_div_v : fvk_div_s;
        check(fn(this->mesh_.handle(), scheme(), faceFlux.internalVector().data(), phi.internalVector().raw(), phi.boundaryData().value().raw(),
                 os.value(), os.view(), divPhi.raw(), mode, this->exec_.stream()));
    }
    void div(Vector<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        div(divPhi, faceFlux, phi, os, FVK_ACC_SCALE);
    }
    void div(VolumeField<T>& divPhi, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        div(divPhi.internalVector(), faceFlux, phi, os, FVK_ACC_SCALE);
    }
    void div(la::LinearSystem<T, localIdx>& ls, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        fvk_term t {};
        if (!fusedTerm(t, faceFlux, os)) NF_ERROR_EXIT("implicit div needs one of the built-in interpolation schemes (linear | upwind)");
        const fvk_bfield bd = phi.boundaryData().c();
        auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
        check(fn(this->mesh_.handle(), 1, &t, &bd, ls.values().raw(), ls.rhs().raw(), ls.boundaryCoefficients().matrixValues.raw(),
                 ls.boundaryCoefficients().rhsValues.raw(), 1, this->exec_.stream()));
    }
    bool addTo(Vector<T>& source, const SurfaceField<scalar>& faceFlux, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        div(source, faceFlux, phi, os, FVK_ADD);
        return true;
    }
    bool fusedTerm(fvk_term& t, const SurfaceField<scalar>& faceFlux, const dsl::Coeff os) const override
    {
        if (scheme() < 0) return false;
        t = fvk_term {};
        t.kind = FVK_TERM_DIV; t.scheme = scheme(); t.coeff = os.value(); t.coeffView = os.view(); t.faceField = faceFlux.internalVector().data();
        return true;
    }
    std::unique_ptr<DivOperatorFactory<T>> clone() const override { return std::make_unique<GaussGreenDiv<T>>(*this); }
private:
    SurfaceInterpolation<T> interp_;
};
NF_REGISTER((DivOperatorFactory<scalar>), (GaussGreenDiv<scalar>));
NF_REGISTER((DivOperatorFactory<Vec3>), (GaussGreenDiv<Vec3>));

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

// ---- operators/laplacianOperator.hpp:27-110: LaplacianOperatorFactory, key "Gauss" --------------------------------------------
template<typename T>
class LaplacianOperatorFactory : public RuntimeSelectionFactory<LaplacianOperatorFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>
{
public:
    using Factory = RuntimeSelectionFactory<LaplacianOperatorFactory<T>, Parameters<const Executor&, const UnstructuredMesh&, const Input&>>;
    static std::unique_ptr<LaplacianOperatorFactory<T>> create(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
    {
        if (input.empty()) NF_ERROR_EXIT("LaplacianOperator: empty token list");
        return Factory::create(input[0], exec, mesh, detail::tail(input));
    }
    static std::string name() { return "LaplacianOperatorFactory"; }
    LaplacianOperatorFactory(const Executor& exec, const UnstructuredMesh& mesh) : exec_(exec), mesh_(mesh) {}
    virtual ~LaplacianOperatorFactory() = default;
    virtual void laplacian(Vector<T>& lapPhi, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    virtual void laplacian(VolumeField<T>& lapPhi, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    virtual void laplacian(la::LinearSystem<T, localIdx>& ls, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const = 0;
    virtual std::unique_ptr<LaplacianOperatorFactory<T>> clone() const = 0;
    virtual bool addTo(Vector<T>&, const SurfaceField<scalar>&, const VolumeField<T>&, const dsl::Coeff) const { return false; }
    virtual bool fusedTerm(fvk_term&, const SurfaceField<scalar>&, const dsl::Coeff) const { return false; }
protected:
    Executor exec_;
    UnstructuredMesh mesh_;
};

// operators/gaussGreenLaplacian.hpp -- explicit: gamma is ignored like the reference (gaussGreenLaplacian.cpp:14)
template<typename T>
class GaussGreenLaplacian : public LaplacianOperatorFactory<T>::template Register<GaussGreenLaplacian<T>>
{
    using Base = typename LaplacianOperatorFactory<T>::template Register<GaussGreenLaplacian<T>>;
    static Input gammaTokens(const Input& in) { return Input({in.size() > 0 ? in[0] : std::string("linear")}); }
    static Input fngTokens(const Input& in) { return Input({in.size() > 1 ? in[1] : (in.size() > 0 ? in[0] : std::string("uncorrected"))}); }
public:
    // tokens after "Gauss": "<interpolation of gamma> <faceNormalGradient scheme>", e.g. "linear uncorrected"
    GaussGreenLaplacian(const Executor& exec, const UnstructuredMesh& mesh, const Input& input)
        : Base(exec, mesh), fng_(exec, mesh, fngTokens(input))
    {
        if (!fng_.fused()) NF_ERROR_EXIT("the fused laplacian kernels implement the uncorrected faceNormalGradient scheme only");
    }
    static std::string name() { return "Gauss"; }
    static std::string doc() { return "Gauss-Green Laplacian"; }
    static std::string schema() { return "none"; }
    void laplacian(Vector<T>& lapPhi, const SurfaceField<scalar>&, const VolumeField<T>& phi, const dsl::Coeff os, int mode) const
    {
        auto fn = std::is_same_v<T, Vec3> ? fvk_laplacian_v : fvk_laplacian_s;
        check(fn(this->mesh_.handle(), phi.internalVector().raw(), phi.boundaryData().value().raw(), os.value(), os.view(), lapPhi.raw(), mode, this->exec_.stream()));
    }
    void laplacian(Vector<T>& lapPhi, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        laplacian(lapPhi, gamma, phi, os, FVK_ACC_SCALE);
    }
    void laplacian(VolumeField<T>& lapPhi, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        laplacian(lapPhi.internalVector(), gamma, phi, os, FVK_ACC_SCALE);
    }
    void laplacian(la::LinearSystem<T, localIdx>& ls, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        fvk_term t {};
        fusedTerm(t, gamma, os);
        const fvk_bfield bd = phi.boundaryData().c();
        auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
        check(fn(this->mesh_.handle(), 1, &t, &bd, ls.values().raw(), ls.rhs().raw(), ls.boundaryCoefficients().matrixValues.raw(),
                 ls.boundaryCoefficients().rhsValues.raw(), 1, this->exec_.stream()));
    }
    bool addTo(Vector<T>& source, const SurfaceField<scalar>& gamma, const VolumeField<T>& phi, const dsl::Coeff os) const override
    {
        laplacian(source, gamma, phi, os, FVK_ADD);
        return true;
    }
    bool fusedTerm(fvk_term& t, const SurfaceField<scalar>& gamma, const dsl::Coeff os) const override
    {
        t = fvk_term {};
        t.kind = FVK_TERM_LAPLACIAN; t.coeff = os.value(); t.coeffView = os.view(); t.faceField = gamma.internalVector().data();
        return true;
    }
    std::unique_ptr<LaplacianOperatorFactory<T>> clone() const override { return std::make_unique<GaussGreenLaplacian<T>>(*this); }
private:
    FaceNormalGradient<T> fng_;
};
NF_REGISTER((LaplacianOperatorFactory<scalar>), (GaussGreenLaplacian<scalar>));
NF_REGISTER((LaplacianOperatorFactory<Vec3>), (GaussGreenLaplacian<Vec3>));

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
