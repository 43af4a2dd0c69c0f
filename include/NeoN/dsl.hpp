// NeoN::dsl for the B200 build: Operator / SpatialOperator / TemporalOperator, Expression, the imp:: and exp:: factories and
// dsl::solve (src/NeoN/include/NeoN/dsl/{operator,spatialOperator,temporalOperator,expression,implicit,explicit,solver}.hpp),
// plus the fvcc operators they wrap (finiteVolume/cellCentred/operators/{divOperator,laplacianOperator,ddtOperator,
// sourceTerm,surfaceIntegrate}.hpp). Same factory names, operator arithmetic and evaluation order; the implicit part of an
// Expression is assembled by ONE fused kernel (fvk_assemble_*), each explicit operator by one fused gather kernel.
#pragma once

#include "NeoN/finiteVolume.hpp"
#include "NeoN/linearAlgebra.hpp"

namespace NeoN::dsl
{
namespace fvcc = NeoN::finiteVolume::cellCentred;

class Operator
{
public:
    enum class Type { Implicit, Explicit };
};

// One term of an expression. The reference type-erases arbitrary operator classes (spatialOperator.hpp:21-124); the hot
// path has a closed set, so a tagged value is enough and lets Expression::assemble hand the whole list to one kernel.
template<typename T>
class SpatialOperator
{
public:
    enum class Kind { Ddt, Div, Laplacian, Source, SurfaceIntegrate };
    SpatialOperator(Kind k, Operator::Type t, fvcc::VolumeField<T>* field, const fvcc::SurfaceField<scalar>* faceScalar,
                    const fvcc::SurfaceField<T>* faceT, const Vector<scalar>* cellCoeff)
        : kind(k), type(t), field_(field), faceScalar_(faceScalar), faceT_(faceT), cellCoeff_(cellCoeff) {}
    Kind kind;
    Operator::Type type;
    Operator::Type getType() const { return type; }
    Coeff& getCoefficient() { return coeffs_; }
    const Coeff& getCoefficient() const { return coeffs_; }
    std::string getName() const
    {
        switch (kind) { case Kind::Ddt: return "DdtOperator"; case Kind::Div: return "DivOperator"; case Kind::Laplacian: return "LaplacianOperator";
                        case Kind::Source: return "sourceTerm"; default: return "SurfaceIntegrate"; }
    }
    fvcc::VolumeField<T>* field() const { return field_; }
    const fvcc::SurfaceField<scalar>* faceField() const { return faceScalar_; }

    // DivOperator::read / LaplacianOperator::read (divOperator.hpp:173-190, laplacianOperator.hpp:181-198)
    void read(const Dictionary& fvSchemes)
    {
        if (kind == Kind::Div && field_)
        {
            const std::string key = "div(" + faceScalar_->name + "," + field_->name + ")";
            auto toks = TokenList::split(fvSchemes.contains("divSchemes") && fvSchemes.subDict("divSchemes").contains(key)
                                             ? fvSchemes.subDict("divSchemes").get<std::string>(key) : std::string("Gauss linear"));
            if (toks[0] != "Gauss") NF_ERROR_EXIT("unknown div scheme: " + toks[0]);
            scheme_ = fvcc::detail::scheme(toks[1]);
        }
        if (kind == Kind::Laplacian)
        {
            const std::string key = "laplacian(" + faceScalar_->name + "," + field_->name + ")";
            auto toks = TokenList::split(fvSchemes.contains("laplacianSchemes") && fvSchemes.subDict("laplacianSchemes").contains(key)
                                             ? fvSchemes.subDict("laplacianSchemes").get<std::string>(key) : std::string("Gauss linear uncorrected"));
            if (toks[0] != "Gauss" || (toks.size() > 2 && toks[2] != "uncorrected")) NF_ERROR_EXIT("unknown laplacian scheme");
        }
    }

    // the term as the fused assembly kernel takes it
    fvk_term term(scalar dt) const
    {
        fvk_term t {};
        t.coeff = coeffs_.value(); t.coeffView = coeffs_.view(); t.scheme = scheme_; t.dt = dt;
        switch (kind)
        {
            case Kind::Ddt: t.kind = FVK_TERM_DDT; t.cellField = field_->oldTime().internalVector().raw(); break;
            case Kind::Div: t.kind = FVK_TERM_DIV; t.faceField = faceScalar_->internalVector().data(); break;
            case Kind::Laplacian: t.kind = FVK_TERM_LAPLACIAN; t.faceField = faceScalar_->internalVector().data(); break;
            case Kind::Source: t.kind = FVK_TERM_SOURCE; t.cellField = cellCoeff_->data(); break;
            default: NF_ERROR_EXIT("SurfaceIntegrate has no implicit form");
        }
        return t;
    }

    // Operator::implicitOperation(ls): applied to an existing system (accumulate)
    void implicitOperation(la::LinearSystem<T, localIdx>& ls, scalar dt = 1.0) const
    {
        const fvk_term t = term(dt);
        const fvk_bfield bd = field_->boundaryData().c();
        auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
        check(fn(ls.mesh().handle(), 1, &t, &bd, ls.values().raw(), ls.rhs().raw(), ls.boundaryCoefficients().matrixValues.raw(),
                 ls.boundaryCoefficients().rhsValues.raw(), 1, ls.exec().stream()));
    }

    // Operator::explicitOperation(source): source += op (divOperator.hpp:133-140 etc.; no temporary, one kernel)
    void explicitOperation(Vector<T>& source, scalar dt = 1.0) const
    {
        const UnstructuredMesh& mesh = field_ ? field_->mesh() : faceT_->mesh();
        const fvk_stream s = source.exec().stream();
        const scalar c = coeffs_.value();
        const scalar* v = coeffs_.view();
        constexpr bool vec = std::is_same_v<T, Vec3>;
        switch (kind)
        {
            case Kind::SurfaceIntegrate:
                check((vec ? fvk_surface_integrate_v : fvk_surface_integrate_s)(mesh.handle(), faceT_->internalVector().raw(), c, v, source.raw(), FVK_ADD, s));
                break;
            case Kind::Div:
                check((vec ? fvk_div_v : fvk_div_s)(mesh.handle(), scheme_, faceScalar_->internalVector().data(), field_->internalVector().raw(),
                                                    field_->boundaryData().value().raw(), c, v, source.raw(), FVK_ADD, s));
                break;
            case Kind::Laplacian:
                check((vec ? fvk_laplacian_v : fvk_laplacian_s)(mesh.handle(), field_->internalVector().raw(), field_->boundaryData().value().raw(), c, v,
                                                                source.raw(), FVK_ADD, s));
                break;
            case Kind::Source:
                check(fvk_source_explicit(mesh.handle(), nComponents<T>(), cellCoeff_->data(), field_->internalVector().raw(), c, v, source.raw(), s));
                break;
            case Kind::Ddt:
                if (coeffs_.hasView() || c != 1.0) NF_ERROR_EXIT("explicit ddt with a coefficient is not implemented"); // as the reference
                check(fvk_ddt_explicit(mesh.handle(), nComponents<T>(), field_->internalVector().raw(), field_->oldTime().internalVector().raw(), dt, source.raw(), s));
                break;
        }
    }
private:
    Coeff coeffs_;
    fvcc::VolumeField<T>* field_;
    const fvcc::SurfaceField<scalar>* faceScalar_;
    const fvcc::SurfaceField<T>* faceT_;
    const Vector<scalar>* cellCoeff_;
    int scheme_ = FVK_LINEAR;
};
template<typename T> using TemporalOperator = SpatialOperator<T>;

template<typename T> SpatialOperator<T> operator*(scalar s, SpatialOperator<T> rhs) { rhs.getCoefficient() *= s; return rhs; }
template<typename T> SpatialOperator<T> operator*(const Vector<scalar>& f, SpatialOperator<T> rhs) { rhs.getCoefficient() *= Coeff(f); return rhs; }
template<typename T> SpatialOperator<T> operator*(const Coeff& c, SpatialOperator<T> rhs) { rhs.getCoefficient() *= c; return rhs; }

// dsl/expression.hpp:47-224
template<typename T>
class Expression
{
public:
    Expression() = default;
    void addOperator(const SpatialOperator<T>& op)
    {
        (op.kind == SpatialOperator<T>::Kind::Ddt ? temporal_ : spatial_).push_back(op);
    }
    void addExpression(const Expression& e)
    {
        for (const auto& o : e.temporal_) temporal_.push_back(o);
        for (const auto& o : e.spatial_) spatial_.push_back(o);
    }
    std::vector<SpatialOperator<T>>& temporalOperators() { return temporal_; }
    std::vector<SpatialOperator<T>>& spatialOperators() { return spatial_; }
    size_t size() const { return temporal_.size() + spatial_.size(); }
    void read(const Dictionary& fvSchemes)
    {
        for (auto& o : temporal_) o.read(fvSchemes);
        for (auto& o : spatial_) o.read(fvSchemes);
    }
    // Expression::explicitOperation(nCells) (expression.hpp:69-78)
    Vector<T> explicitOperation(const Executor& exec, size_t nCells, scalar dt = 1.0) const
    {
        Vector<T> source(exec, nCells, zero<T>());
        for (const auto& o : spatial_) if (o.type == Operator::Type::Explicit) o.explicitOperation(source, dt);
        for (const auto& o : temporal_) if (o.type == Operator::Type::Explicit) o.explicitOperation(source, dt);
        return source;
    }
    bool hasExplicit() const
    {
        for (const auto& o : spatial_) if (o.type == Operator::Type::Explicit) return true;
        for (const auto& o : temporal_) if (o.type == Operator::Type::Explicit) return true;
        return false;
    }
    // Expression::implicitOperation(ls) then (ls, t, dt) (expression.hpp:80-101) as ONE fused launch that writes a
    // fresh system (no zero-fill needed): spatial operators in insertion order, then the temporal ones.
    void assemble(scalar, scalar dt, const la::SparsityPattern&, la::LinearSystem<T, localIdx>& ls, const fvcc::VolumeField<T>& psi) const
    {
        std::vector<fvk_term> terms;
        for (const auto& o : spatial_) if (o.type == Operator::Type::Implicit) terms.push_back(o.term(dt));
        for (const auto& o : temporal_) if (o.type == Operator::Type::Implicit) terms.push_back(o.term(dt));
        if (terms.empty()) { ls.reset(); return; }
        const fvk_bfield bd = psi.boundaryData().c();
        auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
        check(fn(ls.mesh().handle(), int(terms.size()), terms.data(), &bd, ls.values().raw(), ls.rhs().raw(),
                 ls.boundaryCoefficients().matrixValues.raw(), ls.boundaryCoefficients().rhsValues.raw(), 0, ls.exec().stream()));
    }
private:
    std::vector<SpatialOperator<T>> temporal_, spatial_;
};

template<typename T> Expression<T> operator+(const SpatialOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(r); return e; }
template<typename T> Expression<T> operator-(const SpatialOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(-1.0 * r); return e; }
template<typename T> Expression<T> operator+(Expression<T> l, const SpatialOperator<T>& r) { l.addOperator(r); return l; }
template<typename T> Expression<T> operator-(Expression<T> l, const SpatialOperator<T>& r) { l.addOperator(-1.0 * r); return l; }
template<typename T> Expression<T> operator+(Expression<T> l, const Expression<T>& r) { l.addExpression(r); return l; }

// dsl/implicit.hpp
namespace imp
{
template<typename T> SpatialOperator<T> ddt(fvcc::VolumeField<T>& phi) { return {SpatialOperator<T>::Kind::Ddt, Operator::Type::Implicit, &phi, nullptr, nullptr, nullptr}; }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<scalar>& faceFlux, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Div, Operator::Type::Implicit, &phi, &faceFlux, nullptr, nullptr};
}
template<typename T> SpatialOperator<T> laplacian(const fvcc::SurfaceField<scalar>& gamma, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Laplacian, Operator::Type::Implicit, &phi, &gamma, nullptr, nullptr};
}
template<typename T> SpatialOperator<T> source(const fvcc::VolumeField<scalar>& coeff, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Source, Operator::Type::Implicit, &phi, nullptr, nullptr, &coeff.internalVector()};
}
}
// dsl/explicit.hpp: exp::div(flux) is SurfaceIntegrate
namespace exp
{
template<typename T> SpatialOperator<T> ddt(fvcc::VolumeField<T>& phi) { return {SpatialOperator<T>::Kind::Ddt, Operator::Type::Explicit, &phi, nullptr, nullptr, nullptr}; }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<T>& flux) { return {SpatialOperator<T>::Kind::SurfaceIntegrate, Operator::Type::Explicit, nullptr, nullptr, &flux, nullptr}; }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<scalar>& faceFlux, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Div, Operator::Type::Explicit, &phi, &faceFlux, nullptr, nullptr};
}
template<typename T> SpatialOperator<T> laplacian(const fvcc::SurfaceField<scalar>& gamma, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Laplacian, Operator::Type::Explicit, &phi, &gamma, nullptr, nullptr};
}
template<typename T> SpatialOperator<T> source(const fvcc::VolumeField<scalar>& coeff, fvcc::VolumeField<T>& phi)
{
    return {SpatialOperator<T>::Kind::Source, Operator::Type::Explicit, &phi, nullptr, nullptr, &coeff.internalVector()};
}
}

namespace detail
{
// The solve sequence of dsl::solve's steady branch (dsl/solver.hpp:60-80), which FoamAdapter's PDESolver calls as
// iterativeSolveImpl: implicit assembly, rhs -= explicit * V, post-assembly functors, la::Solver.
template<typename T, typename PostAssembly>
la::SolverStats iterativeSolveImpl(Expression<T>& expr, const la::SparsityPattern& sp, la::LinearSystem<T, localIdx>& ls,
                                   fvcc::VolumeField<T>& psi, scalar t, scalar dt, const la::Solver& solver, PostAssembly&& post)
{
    static_assert(std::is_same_v<T, scalar>, "only the scalar solve is on the hot path");
    expr.assemble(t, dt, sp, ls, psi);
    if (expr.hasExplicit())
    {
        auto expTmp = expr.explicitOperation(psi.exec(), size_t(psi.mesh().nCells()), dt);
        check(fvk_rhs_sub_source(psi.mesh().handle(), 1, expTmp.raw(), ls.rhs().raw(), psi.exec().stream()));
    }
    post(sp, ls);
    return solver.solve(ls, psi.internalVector());
}
}

// dsl::solve (dsl/solver.hpp:35-82) for expressions without temporal terms, and backwardEuler / forwardEuler for those with
// (timeIntegration/{backwardEuler,forwardEuler}.hpp): fvSchemes.ddtSchemes.type selects the integrator.
inline la::SolverStats solve(Expression<scalar>& expr, fvcc::VolumeField<scalar>& solution, scalar t, scalar dt,
                             const Dictionary& fvSchemes, const Dictionary& fvSolution)
{
    if (expr.size() == 0) NF_ERROR_EXIT("No temporal or implicit terms to solve.");
    expr.read(fvSchemes);
    const auto& mesh = solution.mesh();
    if (!expr.temporalOperators().empty())
    {
        const auto type = fvSchemes.subDict("ddtSchemes").get<std::string>("type");
        if (type == "forwardEuler")
        { // forwardEuler.hpp:44-49: phi = old - source*dt
            Expression<scalar> rhsOnly;
            for (auto& o : expr.spatialOperators()) rhsOnly.addOperator(o);
            auto source = rhsOnly.explicitOperation(solution.exec(), size_t(mesh.nCells()), dt);
            auto& old = solution.oldTime().internalVector();
            solution.internalVector() = old;
            check(fvk_vec_axpby(int64_t(mesh.nCells()), -dt, source.data(), 1.0, solution.internalVector().data(), solution.exec().stream()));
            solution.correctBoundaryConditions();
            solution.exec().sync();
            return {0, 0.0, 0.0, {}};
        }
        if (type != "backwardEuler") NF_ERROR_EXIT("time integrator " + type + " is out of scope (forwardEuler | backwardEuler)");
    }
    auto sp = la::SparsityPattern::readOrCreate(mesh);
    la::LinearSystem<scalar, localIdx> ls(mesh, sp, false);
    la::Solver solver(solution.exec(), fvSolution);
    return detail::iterativeSolveImpl(expr, sp, ls, solution, t, dt, solver, [](const la::SparsityPattern&, la::LinearSystem<scalar, localIdx>&) {});
}

} // namespace NeoN::dsl
