// NeoN::dsl for the B200 build: Operator / OperatorMixin, the TYPE-ERASED SpatialOperator / TemporalOperator, Expression, the
// imp:: and exp:: factories, the time integrators and dsl::solve (src/NeoN/include/NeoN/dsl/{operator,spatialOperator,
// temporalOperator,expression,implicit,explicit,solver}.hpp, timeIntegration/*.hpp), plus the fvcc operator classes they wrap
// (finiteVolume/cellCentred/operators/{divOperator,laplacianOperator,ddtOperator,sourceTerm,surfaceIntegrate}.hpp).
//
// Open like the reference: ANY class with `void explicitOperation(Vector<T>&) const` and / or `void implicitOperation(
// la::LinearSystem<T, localIdx>&) const` (spatialOperator.hpp:21-37; e.g. the Dummy of src/NeoN/test/dsl/common.hpp:53-106)
// can be added to an Expression, and div / laplacian / interpolation / faceNormalGradient strategies, linear solvers and time
// integrators are selected by NAME through RuntimeSelectionFactory tables a plug-in can register into.
// Fast like the hardware wants: when every implicit operator of an expression is a built-in one, Expression::assemble hands
// the whole term list to ONE fused kernel (fvk_assemble_*) that writes a fresh system; otherwise the system is zeroed and each
// operator adds its coefficients, exactly the reference's sequence.
#pragma once

#include "NeoN/finiteVolume.hpp"
#include "NeoN/linearAlgebra.hpp"

namespace NeoN::dsl
{
namespace fvcc = NeoN::finiteVolume::cellCentred;

// dsl/operator.hpp:17-64
class Operator
{
public:
    enum class Type { Implicit, Explicit };
};

template<typename VectorType>
class OperatorMixin
{
public:
    OperatorMixin(const Executor exec, const Coeff& coeffs, VectorType& field, Operator::Type type) : exec_(exec), coeffs_(coeffs), field_(field), type_(type) {}
    virtual ~OperatorMixin() = default;
    Operator::Type getType() const { return type_; }
    virtual const Executor& exec() const final { return exec_; }
    Coeff& getCoefficient() { return coeffs_; }
    const Coeff& getCoefficient() const { return coeffs_; }
    VectorType& getVector() { return field_; }
    const VectorType& getVector() const { return field_; }
    void read(const Dictionary&) {}
protected:
    const Executor exec_;
    Coeff coeffs_;
    VectorType& field_;
    Operator::Type type_;
};

namespace detail
{
// the reference's concepts (spatialOperator.hpp:21-37, temporalOperator.hpp:18-41) in C++17 detection-idiom form
template<typename T, typename V, typename = void> struct HasExplicit : std::false_type {};
template<typename T, typename V>
struct HasExplicit<T, V, std::void_t<decltype(std::declval<const T&>().explicitOperation(std::declval<Vector<V>&>()))>> : std::true_type {};
template<typename T, typename V, typename = void> struct HasImplicit : std::false_type {};
template<typename T, typename V>
struct HasImplicit<T, V, std::void_t<decltype(std::declval<const T&>().implicitOperation(std::declval<la::LinearSystem<V, localIdx>&>()))>> : std::true_type {};
template<typename T, typename V, typename = void> struct HasTemporalExplicit : std::false_type {};
template<typename T, typename V>
struct HasTemporalExplicit<T, V, std::void_t<decltype(std::declval<const T&>().explicitOperation(std::declval<Vector<V>&>(), scalar(), scalar()))>> : std::true_type {};
template<typename T, typename V, typename = void> struct HasTemporalImplicit : std::false_type {};
template<typename T, typename V>
struct HasTemporalImplicit<T, V, std::void_t<decltype(std::declval<const T&>().implicitOperation(std::declval<la::LinearSystem<V, localIdx>&>(), scalar(), scalar()))>> : std::true_type {};
// optional hot-path hooks of the built-in operators
template<typename T, typename = void> struct HasFusedTerm : std::false_type {};
template<typename T>
struct HasFusedTerm<T, std::void_t<decltype(std::declval<const T&>().fusedTerm(std::declval<fvk_term&>(), scalar()))>> : std::true_type {};
template<typename T, typename = void> struct HasRead : std::false_type {};
template<typename T> struct HasRead<T, std::void_t<decltype(std::declval<T&>().read(std::declval<const Dictionary&>()))>> : std::true_type {};
}

// dsl/spatialOperator.hpp:40-200 -- type erasure: concrete operators of any class live in one vector
template<typename ValueType>
class SpatialOperator
{
public:
    using VectorValueType = ValueType;
    template<typename T, typename = std::enable_if_t<!std::is_same_v<std::decay_t<T>, SpatialOperator>
                                                     && (detail::HasExplicit<std::decay_t<T>, ValueType>::value || detail::HasImplicit<std::decay_t<T>, ValueType>::value)>>
    SpatialOperator(T cls) : model_(std::make_unique<OperatorModel<std::decay_t<T>>>(std::move(cls))) {}
    SpatialOperator(const SpatialOperator& o) : model_(o.model_->clone()) {}
    SpatialOperator(SpatialOperator&&) = default;
    SpatialOperator& operator=(const SpatialOperator& o) { model_ = o.model_->clone(); return *this; }
    SpatialOperator& operator=(SpatialOperator&&) = default;

    void explicitOperation(Vector<ValueType>& source) const { model_->explicitOperation(source); }
    void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls) const { model_->implicitOperation(ls); }
    Operator::Type getType() const { return model_->getType(); }
    std::string getName() const { return model_->getName(); }
    Coeff& getCoefficient() { return model_->getCoefficient(); }
    Coeff getCoefficient() const { return model_->getCoefficientCopy(); }
    void read(const Dictionary& input) { model_->read(input); }
    const Executor& exec() const { return model_->exec(); }
    // hot path: this operator as a term of the fused assembly kernel (false: not a built-in operator)
    bool fusedTerm(fvk_term& t, scalar dt) const { return model_->fusedTerm(t, dt); }

private:
    struct OperatorConcept
    {
        virtual ~OperatorConcept() = default;
        virtual void explicitOperation(Vector<ValueType>& source) const = 0;
        virtual void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls) const = 0;
        virtual void read(const Dictionary& input) = 0;
        virtual std::string getName() const = 0;
        virtual Operator::Type getType() const = 0;
        virtual Coeff& getCoefficient() = 0;
        virtual Coeff getCoefficientCopy() const = 0;
        virtual const Executor& exec() const = 0;
        virtual bool fusedTerm(fvk_term&, scalar) const = 0;
        virtual std::unique_ptr<OperatorConcept> clone() const = 0;
    };
    template<typename ConcreteOperatorType>
    struct OperatorModel : OperatorConcept
    {
        OperatorModel(ConcreteOperatorType c) : concreteOp_(std::move(c)) {}
        void explicitOperation(Vector<ValueType>& source) const override
        {
            if constexpr (detail::HasExplicit<ConcreteOperatorType, ValueType>::value) concreteOp_.explicitOperation(source);
        }
        void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls) const override
        {
            if constexpr (detail::HasImplicit<ConcreteOperatorType, ValueType>::value) concreteOp_.implicitOperation(ls);
        }
        void read(const Dictionary& input) override
        {
            if constexpr (detail::HasRead<ConcreteOperatorType>::value) concreteOp_.read(input);
        }
        std::string getName() const override { return concreteOp_.getName(); }
        Operator::Type getType() const override { return concreteOp_.getType(); }
        Coeff& getCoefficient() override { return concreteOp_.getCoefficient(); }
        Coeff getCoefficientCopy() const override { return concreteOp_.getCoefficient(); }
        const Executor& exec() const override { return concreteOp_.exec(); }
        bool fusedTerm(fvk_term& t, scalar dt) const override
        {
            if constexpr (detail::HasFusedTerm<ConcreteOperatorType>::value) return concreteOp_.fusedTerm(t, dt);
            else { (void) t; (void) dt; return false; }
        }
        std::unique_ptr<OperatorConcept> clone() const override { return std::make_unique<OperatorModel>(*this); }
        ConcreteOperatorType concreteOp_;
    };
    std::unique_ptr<OperatorConcept> model_;
};

// dsl/temporalOperator.hpp:44-190
template<typename ValueType>
class TemporalOperator
{
public:
    using VectorValueType = ValueType;
    template<typename T, typename = std::enable_if_t<!std::is_same_v<std::decay_t<T>, TemporalOperator>
                                                     && (detail::HasTemporalExplicit<std::decay_t<T>, ValueType>::value
                                                         || detail::HasTemporalImplicit<std::decay_t<T>, ValueType>::value)>>
    TemporalOperator(T cls) : model_(std::make_unique<Model<std::decay_t<T>>>(std::move(cls))) {}
    TemporalOperator(const TemporalOperator& o) : model_(o.model_->clone()) {}
    TemporalOperator(TemporalOperator&&) = default;
    TemporalOperator& operator=(const TemporalOperator& o) { model_ = o.model_->clone(); return *this; }
    TemporalOperator& operator=(TemporalOperator&&) = default;

    void explicitOperation(Vector<ValueType>& source, scalar t, scalar dt) const { model_->explicitOperation(source, t, dt); }
    void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls, scalar t, scalar dt) const { model_->implicitOperation(ls, t, dt); }
    Operator::Type getType() const { return model_->getType(); }
    std::string getName() const { return model_->getName(); }
    Coeff& getCoefficient() { return model_->getCoefficient(); }
    Coeff getCoefficient() const { return model_->getCoefficientCopy(); }
    void read(const Dictionary& input) { model_->read(input); }
    const Executor& exec() const { return model_->exec(); }
    bool fusedTerm(fvk_term& t, scalar dt) const { return model_->fusedTerm(t, dt); }

private:
    struct Concept
    {
        virtual ~Concept() = default;
        virtual void explicitOperation(Vector<ValueType>&, scalar, scalar) const = 0;
        virtual void implicitOperation(la::LinearSystem<ValueType, localIdx>&, scalar, scalar) const = 0;
        virtual void read(const Dictionary&) = 0;
        virtual std::string getName() const = 0;
        virtual Operator::Type getType() const = 0;
        virtual Coeff& getCoefficient() = 0;
        virtual Coeff getCoefficientCopy() const = 0;
        virtual const Executor& exec() const = 0;
        virtual bool fusedTerm(fvk_term&, scalar) const = 0;
        virtual std::unique_ptr<Concept> clone() const = 0;
    };
    template<typename C>
    struct Model : Concept
    {
        Model(C c) : op_(std::move(c)) {}
        void explicitOperation(Vector<ValueType>& s, scalar t, scalar dt) const override
        {
            if constexpr (detail::HasTemporalExplicit<C, ValueType>::value) op_.explicitOperation(s, t, dt);
        }
        void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls, scalar t, scalar dt) const override
        {
            if constexpr (detail::HasTemporalImplicit<C, ValueType>::value) op_.implicitOperation(ls, t, dt);
        }
        void read(const Dictionary& input) override
        {
            if constexpr (detail::HasRead<C>::value) op_.read(input);
        }
        std::string getName() const override { return op_.getName(); }
        Operator::Type getType() const override { return op_.getType(); }
        Coeff& getCoefficient() override { return op_.getCoefficient(); }
        Coeff getCoefficientCopy() const override { return op_.getCoefficient(); }
        const Executor& exec() const override { return op_.exec(); }
        bool fusedTerm(fvk_term& t, scalar dt) const override
        {
            if constexpr (detail::HasFusedTerm<C>::value) return op_.fusedTerm(t, dt);
            else { (void) t; (void) dt; return false; }
        }
        std::unique_ptr<Concept> clone() const override { return std::make_unique<Model>(*this); }
        C op_;
    };
    std::unique_ptr<Concept> model_;
};

template<typename T> SpatialOperator<T> operator*(scalar s, SpatialOperator<T> rhs) { rhs.getCoefficient() *= s; return rhs; }
template<typename T> SpatialOperator<T> operator*(const Vector<scalar>& f, SpatialOperator<T> rhs) { rhs.getCoefficient() *= Coeff(f); return rhs; }
template<typename T> SpatialOperator<T> operator*(const Coeff& c, SpatialOperator<T> rhs) { rhs.getCoefficient() *= c; return rhs; }
template<typename T> TemporalOperator<T> operator*(scalar s, TemporalOperator<T> rhs) { rhs.getCoefficient() *= s; return rhs; }
template<typename T> TemporalOperator<T> operator*(const Vector<scalar>& f, TemporalOperator<T> rhs) { rhs.getCoefficient() *= Coeff(f); return rhs; }
} // namespace NeoN::dsl

// ---- the built-in operators (finiteVolume/cellCentred/operators/*.hpp) ------------------------------------------------------
namespace NeoN::finiteVolume::cellCentred
{
namespace detail
{
inline Input schemeTokens(const Dictionary& fvSchemes, const std::string& dict, const std::string& key)
{
    if (!fvSchemes.contains(dict) || !fvSchemes.subDict(dict).contains(key)) NF_ERROR_EXIT("Key " + key + " not found in dictionary " + dict);
    const auto& d = fvSchemes.subDict(dict);
    try { return d.get<TokenList>(key); } catch (...) { return TokenList::split(d.get<std::string>(key)); }
}
template<typename T>
void assembleOne(const fvk_term& t, const VolumeField<T>& phi, la::LinearSystem<T, localIdx>& ls)
{
    const fvk_bfield bd = phi.boundaryData().c();
    auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
    check(fn(ls.mesh().handle(), 1, &t, &bd, ls.values().raw(), ls.rhs().raw(), ls.boundaryCoefficients().matrixValues.raw(),
             ls.boundaryCoefficients().rhsValues.raw(), 1, ls.exec().stream()));
}
}

// operators/divOperator.hpp:103-200
template<typename T>
class DivOperator : public dsl::OperatorMixin<VolumeField<T>>
{
public:
    using VectorValueType = T;
    DivOperator(dsl::Operator::Type termType, const SurfaceField<scalar>& faceFlux, VolumeField<T>& phi)
        : dsl::OperatorMixin<VolumeField<T>>(phi.exec(), dsl::Coeff(1.0), phi, termType), faceFlux_(faceFlux) {}
    DivOperator(dsl::Operator::Type termType, const SurfaceField<scalar>& faceFlux, VolumeField<T>& phi, const Input& input)
        : dsl::OperatorMixin<VolumeField<T>>(phi.exec(), dsl::Coeff(1.0), phi, termType), faceFlux_(faceFlux),
          strategy_(DivOperatorFactory<T>::create(phi.exec(), phi.mesh(), input)) {}
    DivOperator(const DivOperator& o)
        : dsl::OperatorMixin<VolumeField<T>>(o.exec_, o.coeffs_, o.field_, o.type_), faceFlux_(o.faceFlux_), strategy_(o.strategy_ ? o.strategy_->clone() : nullptr) {}
    void explicitOperation(Vector<T>& source) const
    { // divOperator.hpp:133-140: tmp = 0; div(tmp); source += tmp -- one ADD-mode launch when the strategy offers it
        NF_ASSERT(strategy_, "DivOperatorStrategy not initialized");
        if (strategy_->addTo(source, faceFlux_, this->getVector(), this->getCoefficient())) return;
        Vector<T> tmp(source.exec(), source.size(), zero<T>());
        strategy_->div(tmp, faceFlux_, this->getVector(), this->getCoefficient());
        source += tmp;
    }
    void implicitOperation(la::LinearSystem<T, localIdx>& ls) const
    {
        NF_ASSERT(strategy_, "DivOperatorStrategy not initialized");
        strategy_->div(ls, faceFlux_, this->getVector(), this->getCoefficient());
    }
    void div(Vector<T>& divPhi) const { strategy_->div(divPhi, faceFlux_, this->getVector(), this->getCoefficient()); }
    void div(VolumeField<T>& divPhi) const { strategy_->div(divPhi, faceFlux_, this->getVector(), this->getCoefficient()); }
    void read(const Dictionary& fvSchemes)
    { // :173-190
        strategy_ = DivOperatorFactory<T>::create(this->exec(), this->getVector().mesh(),
                                                  detail::schemeTokens(fvSchemes, "divSchemes", "div(" + faceFlux_.name + "," + this->getVector().name + ")"));
    }
    bool fusedTerm(fvk_term& t, scalar) const { return strategy_ && strategy_->fusedTerm(t, faceFlux_, this->getCoefficient()); }
    std::string getName() const { return "DivOperator"; }
private:
    const SurfaceField<scalar>& faceFlux_;
    std::unique_ptr<DivOperatorFactory<T>> strategy_;
};

// operators/laplacianOperator.hpp:113-210
template<typename T>
class LaplacianOperator : public dsl::OperatorMixin<VolumeField<T>>
{
public:
    using VectorValueType = T;
    LaplacianOperator(dsl::Operator::Type termType, const SurfaceField<scalar>& gamma, VolumeField<T>& phi)
        : dsl::OperatorMixin<VolumeField<T>>(phi.exec(), dsl::Coeff(1.0), phi, termType), gamma_(gamma) {}
    LaplacianOperator(dsl::Operator::Type termType, const SurfaceField<scalar>& gamma, VolumeField<T>& phi, const Input& input)
        : dsl::OperatorMixin<VolumeField<T>>(phi.exec(), dsl::Coeff(1.0), phi, termType), gamma_(gamma),
          strategy_(LaplacianOperatorFactory<T>::create(phi.exec(), phi.mesh(), input)) {}
    LaplacianOperator(const LaplacianOperator& o)
        : dsl::OperatorMixin<VolumeField<T>>(o.exec_, o.coeffs_, o.field_, o.type_), gamma_(o.gamma_), strategy_(o.strategy_ ? o.strategy_->clone() : nullptr) {}
    void explicitOperation(Vector<T>& source) const
    {
        NF_ASSERT(strategy_, "LaplacianOperatorStrategy not initialized");
        if (strategy_->addTo(source, gamma_, this->getVector(), this->getCoefficient())) return;
        Vector<T> tmp(source.exec(), source.size(), zero<T>());
        strategy_->laplacian(tmp, gamma_, this->getVector(), this->getCoefficient());
        source += tmp;
    }
    void implicitOperation(la::LinearSystem<T, localIdx>& ls) const
    {
        NF_ASSERT(strategy_, "LaplacianOperatorStrategy not initialized");
        strategy_->laplacian(ls, gamma_, this->getVector(), this->getCoefficient());
    }
    void laplacian(Vector<T>& lapPhi) const { strategy_->laplacian(lapPhi, gamma_, this->getVector(), this->getCoefficient()); }
    void laplacian(VolumeField<T>& lapPhi) const { strategy_->laplacian(lapPhi, gamma_, this->getVector(), this->getCoefficient()); }
    void read(const Dictionary& fvSchemes)
    { // :181-198
        strategy_ = LaplacianOperatorFactory<T>::create(this->exec(), this->getVector().mesh(),
                                                        detail::schemeTokens(fvSchemes, "laplacianSchemes", "laplacian(" + gamma_.name + "," + this->getVector().name + ")"));
    }
    bool fusedTerm(fvk_term& t, scalar) const { return strategy_ && strategy_->fusedTerm(t, gamma_, this->getCoefficient()); }
    std::string getName() const { return "LaplacianOperator"; }
private:
    const SurfaceField<scalar>& gamma_;
    std::unique_ptr<LaplacianOperatorFactory<T>> strategy_;
};

// operators/ddtOperator.hpp + ddtOperator.cpp:21-60 (a TEMPORAL operator)
template<typename FieldT>
class DdtOperator : public dsl::OperatorMixin<FieldT>
{
public:
    using VectorValueType = typename FieldT::ElementType;
    using T = VectorValueType;
    DdtOperator(dsl::Operator::Type termType, FieldT& field) : dsl::OperatorMixin<FieldT>(field.exec(), dsl::Coeff(1.0), field, termType) {}
    void explicitOperation(Vector<T>& source, scalar, scalar dt) const
    {
        if (this->getCoefficient().hasView() || this->getCoefficient().value() != 1.0) NF_ERROR_EXIT("Not implemented"); // ddtOperator.cpp:28-30: coefficients ignored
        auto& f = const_cast<FieldT&>(this->getVector());
        check(fvk_ddt_explicit(f.mesh().handle(), nComponents<T>(), f.internalVector().raw(), f.oldTime().internalVector().raw(), dt, source.raw(), source.exec().stream()));
    }
    void implicitOperation(la::LinearSystem<T, localIdx>& ls, scalar, scalar dt) const
    {
        fvk_term t {};
        fusedTerm(t, dt);
        detail::assembleOne<T>(t, this->getVector(), ls);
    }
    bool fusedTerm(fvk_term& t, scalar dt) const
    {
        t = fvk_term {};
        t.kind = FVK_TERM_DDT; t.coeff = this->getCoefficient().value(); t.coeffView = this->getCoefficient().view(); t.dt = dt;
        t.cellField = const_cast<FieldT&>(this->getVector()).oldTime().internalVector().raw();
        return true;
    }
    std::string getName() const { return "DdtOperator"; }
};

// operators/sourceTerm.hpp + sourceTerm.cpp:22-55
template<typename T>
class SourceTerm : public dsl::OperatorMixin<VolumeField<T>>
{
public:
    using VectorValueType = T;
    SourceTerm(dsl::Operator::Type termType, const VolumeField<scalar>& coefficients, VolumeField<T>& field)
        : dsl::OperatorMixin<VolumeField<T>>(field.exec(), dsl::Coeff(1.0), field, termType), coefficients_(coefficients) {}
    void explicitOperation(Vector<T>& source) const
    {
        const auto& f = this->getVector();
        check(fvk_source_explicit(f.mesh().handle(), nComponents<T>(), coefficients_.internalVector().data(), f.internalVector().raw(), this->getCoefficient().value(),
                                  this->getCoefficient().view(), source.raw(), source.exec().stream()));
    }
    void implicitOperation(la::LinearSystem<T, localIdx>& ls) const
    {
        fvk_term t {};
        fusedTerm(t, 1.0);
        detail::assembleOne<T>(t, this->getVector(), ls);
    }
    bool fusedTerm(fvk_term& t, scalar) const
    {
        t = fvk_term {};
        t.kind = FVK_TERM_SOURCE; t.coeff = this->getCoefficient().value(); t.coeffView = this->getCoefficient().view();
        t.cellField = coefficients_.internalVector().data();
        return true;
    }
    std::string getName() const { return "sourceTerm"; }
private:
    const VolumeField<scalar>& coefficients_;
};

// operators/surfaceIntegrate.hpp (exp::div(flux)): source += surfaceIntegrate(flux) in one ADD-mode launch
template<typename T>
class SurfaceIntegrate : public dsl::OperatorMixin<SurfaceField<T>>
{
public:
    using VectorValueType = T;
    SurfaceIntegrate(const SurfaceField<T>& flux)
        : dsl::OperatorMixin<SurfaceField<T>>(flux.exec(), dsl::Coeff(1.0), const_cast<SurfaceField<T>&>(flux), dsl::Operator::Type::Explicit) {}
    void explicitOperation(Vector<T>& source) const
    {
        const auto& flux = this->getVector();
        auto fn = std::is_same_v<T, Vec3> ? fvk_surface_integrate_v : fvk_surface_integrate_s;
        check(fn(flux.mesh().handle(), flux.internalVector().raw(), this->getCoefficient().value(), this->getCoefficient().view(), source.raw(), FVK_ADD,
                 source.exec().stream()));
    }
    std::string getName() const { return "SurfaceIntegrate"; }
};
} // namespace NeoN::finiteVolume::cellCentred

namespace NeoN::dsl
{
// dsl/expression.hpp:25-224
template<typename T>
class Expression
{
public:
    Expression() = default;
    explicit Expression(const Executor&) {}
    void addOperator(const SpatialOperator<T>& op) { spatial_.push_back(op); }
    void addOperator(const TemporalOperator<T>& op) { temporal_.push_back(op); }
    void addExpression(const Expression& e)
    {
        for (const auto& o : e.temporal_) temporal_.push_back(o);
        for (const auto& o : e.spatial_) spatial_.push_back(o);
    }
    std::vector<TemporalOperator<T>>& temporalOperators() { return temporal_; }
    const std::vector<TemporalOperator<T>>& temporalOperators() const { return temporal_; }
    std::vector<SpatialOperator<T>>& spatialOperators() { return spatial_; }
    const std::vector<SpatialOperator<T>>& spatialOperators() const { return spatial_; }
    size_t size() const { return temporal_.size() + spatial_.size(); }
    void read(const Dictionary& input)
    {
        for (auto& o : temporal_) o.read(input);
        for (auto& o : spatial_) o.read(input);
    }
    // expression.hpp:48-65: the explicit SPATIAL operators
    Vector<T> explicitOperation(const Executor& exec, size_t nCells) const
    {
        Vector<T> source(exec, nCells, zero<T>());
        explicitOperation(source);
        return source;
    }
    void explicitOperation(Vector<T>& source) const
    {
        for (const auto& o : spatial_) if (o.getType() == Operator::Type::Explicit) o.explicitOperation(source);
    }
    // :67-78 the explicit TEMPORAL operators
    void explicitOperation(Vector<T>& source, scalar t, scalar dt) const
    {
        for (const auto& o : temporal_) if (o.getType() == Operator::Type::Explicit) o.explicitOperation(source, t, dt);
    }
    bool hasExplicit() const
    {
        for (const auto& o : spatial_) if (o.getType() == Operator::Type::Explicit) return true;
        return false;
    }
    // :80-101
    void implicitOperation(la::LinearSystem<T, localIdx>& ls) const
    {
        for (const auto& o : spatial_) if (o.getType() == Operator::Type::Implicit) o.implicitOperation(ls);
    }
    void implicitOperation(la::LinearSystem<T, localIdx>& ls, scalar t, scalar dt) const
    {
        for (const auto& o : temporal_) if (o.getType() == Operator::Type::Implicit) o.implicitOperation(ls, t, dt);
    }
    // Expression::implicitOperation(ls) then (ls, t, dt) on a fresh system. All implicit operators built-in: ONE fused launch that
    // writes every entry (no zero-fill); otherwise zero the system and let every operator add its coefficients, in the same order.
    void assemble(scalar t, scalar dt, const la::SparsityPattern&, la::LinearSystem<T, localIdx>& ls, const fvcc::VolumeField<T>& psi) const
    {
        std::vector<fvk_term> terms;
        bool fused = true;
        for (const auto& o : spatial_)
            if (o.getType() == Operator::Type::Implicit) { fvk_term tm {}; if (o.fusedTerm(tm, dt)) terms.push_back(tm); else fused = false; }
        for (const auto& o : temporal_)
            if (o.getType() == Operator::Type::Implicit) { fvk_term tm {}; if (o.fusedTerm(tm, dt)) terms.push_back(tm); else fused = false; }
        if (fused && !terms.empty() && terms.size() <= FVK_MAX_TERMS)
        {
            const fvk_bfield bd = psi.boundaryData().c();
            auto fn = std::is_same_v<T, Vec3> ? fvk_assemble_v : fvk_assemble_s;
            check(fn(ls.mesh().handle(), int(terms.size()), terms.data(), &bd, ls.values().raw(), ls.rhs().raw(),
                     ls.boundaryCoefficients().matrixValues.raw(), ls.boundaryCoefficients().rhsValues.raw(), 0, ls.exec().stream()));
            return;
        }
        ls.reset();
        implicitOperation(ls);
        implicitOperation(ls, t, dt);
    }
private:
    std::vector<TemporalOperator<T>> temporal_;
    std::vector<SpatialOperator<T>> spatial_;
};

// expression.hpp:226-330: operator arithmetic; operator- multiplies the right-hand side's Coeff by -1
template<typename T> Expression<T> operator+(const SpatialOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(r); return e; }
template<typename T> Expression<T> operator-(const SpatialOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(-1.0 * r); return e; }
template<typename T> Expression<T> operator+(const TemporalOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(r); return e; }
template<typename T> Expression<T> operator-(const TemporalOperator<T>& l, const SpatialOperator<T>& r) { Expression<T> e; e.addOperator(l); e.addOperator(-1.0 * r); return e; }
template<typename T> Expression<T> operator+(Expression<T> l, const SpatialOperator<T>& r) { l.addOperator(r); return l; }
template<typename T> Expression<T> operator-(Expression<T> l, const SpatialOperator<T>& r) { l.addOperator(-1.0 * r); return l; }
template<typename T> Expression<T> operator+(Expression<T> l, const TemporalOperator<T>& r) { l.addOperator(r); return l; }
template<typename T> Expression<T> operator+(Expression<T> l, const Expression<T>& r) { l.addExpression(r); return l; }

// dsl/implicit.hpp
namespace imp
{
template<typename T> TemporalOperator<T> ddt(fvcc::VolumeField<T>& phi) { return TemporalOperator<T>(fvcc::DdtOperator<fvcc::VolumeField<T>>(Operator::Type::Implicit, phi)); }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<scalar>& faceFlux, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::DivOperator<T>(Operator::Type::Implicit, faceFlux, phi));
}
template<typename T> SpatialOperator<T> laplacian(const fvcc::SurfaceField<scalar>& gamma, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::LaplacianOperator<T>(Operator::Type::Implicit, gamma, phi));
}
template<typename T> SpatialOperator<T> source(const fvcc::VolumeField<scalar>& coeff, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::SourceTerm<T>(Operator::Type::Implicit, coeff, phi));
}
}
// dsl/explicit.hpp: exp::div(flux) is SurfaceIntegrate
namespace exp
{
template<typename T> TemporalOperator<T> ddt(fvcc::VolumeField<T>& phi) { return TemporalOperator<T>(fvcc::DdtOperator<fvcc::VolumeField<T>>(Operator::Type::Explicit, phi)); }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<T>& flux) { return SpatialOperator<T>(fvcc::SurfaceIntegrate<T>(flux)); }
template<typename T> SpatialOperator<T> div(const fvcc::SurfaceField<scalar>& faceFlux, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::DivOperator<T>(Operator::Type::Explicit, faceFlux, phi));
}
template<typename T> SpatialOperator<T> laplacian(const fvcc::SurfaceField<scalar>& gamma, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::LaplacianOperator<T>(Operator::Type::Explicit, gamma, phi));
}
template<typename T> SpatialOperator<T> source(const fvcc::VolumeField<scalar>& coeff, fvcc::VolumeField<T>& phi)
{
    return SpatialOperator<T>(fvcc::SourceTerm<T>(Operator::Type::Explicit, coeff, phi));
}
}

namespace detail
{
// The solve sequence of dsl::solve's steady branch (dsl/solver.hpp:60-80), which FoamAdapter's PDESolver calls as
// iterativeSolveImpl: implicit assembly, rhs -= explicit * V, post-assembly functors, la::Solver.
template<typename T, typename PostAssembly>
auto iterativeSolveImpl(Expression<T>& expr, const la::SparsityPattern& sp, la::LinearSystem<T, localIdx>& ls,
                        fvcc::VolumeField<T>& psi, scalar t, scalar dt, const la::Solver& solver, PostAssembly&& post)
{
    expr.assemble(t, dt, sp, ls, psi);
    if (expr.hasExplicit())
    {
        auto expTmp = expr.explicitOperation(psi.exec(), size_t(psi.mesh().nCells()));
        check(fvk_rhs_sub_source(psi.mesh().handle(), nComponents<T>(), expTmp.raw(), ls.rhs().raw(), psi.exec().stream()));
    }
    post(sp, ls);
    return solver.solve(ls, psi.internalVector());
}
}
} // namespace NeoN::dsl

// ---- timeIntegration/{timeIntegration,forwardEuler,backwardEuler,rungeKutta}.hpp: integrators selected by ddtSchemes.type -----
namespace NeoN::timeIntegration
{
template<typename SolutionVectorType>
class TimeIntegratorBase : public RuntimeSelectionFactory<TimeIntegratorBase<SolutionVectorType>, Parameters<const Dictionary&, const Dictionary&>>
{
public:
    using ValueType = typename SolutionVectorType::ElementType;
    using Expression = dsl::Expression<ValueType>;
    static std::string name() { return "timeIntegrationFactory"; }
    TimeIntegratorBase(const Dictionary& schemeDict, const Dictionary& solutionDict) : schemeDict_(schemeDict), solutionDict_(solutionDict) {}
    virtual ~TimeIntegratorBase() = default;
    virtual void solve(Expression& eqn, SolutionVectorType& sol, scalar t, scalar dt) = 0;
    virtual std::unique_ptr<TimeIntegratorBase> clone() const = 0;
    la::SolverStats lastStats {0, 0.0, 0.0, {}};
protected:
    const Dictionary schemeDict_;
    const Dictionary solutionDict_;
};

// forwardEuler.hpp:38-56
template<typename SolutionVectorType>
class ForwardEuler : public TimeIntegratorBase<SolutionVectorType>::template Register<ForwardEuler<SolutionVectorType>>
{
    using Base = typename TimeIntegratorBase<SolutionVectorType>::template Register<ForwardEuler<SolutionVectorType>>;
    using ValueType = typename SolutionVectorType::ElementType;
public:
    ForwardEuler(const Dictionary& schemeDict, const Dictionary& solutionDict) : Base(schemeDict, solutionDict) {}
    static std::string name() { return "forwardEuler"; }
    static std::string doc() { return "first order time integration method"; }
    static std::string schema() { return "none"; }
    void solve(dsl::Expression<ValueType>& eqn, SolutionVectorType& sol, scalar, scalar dt) override
    {
        auto source = eqn.explicitOperation(sol.exec(), size_t(sol.mesh().nCells()));
        auto& old = sol.oldTime();
        sol.internalVector() = old.internalVector();   // solution = old - source * dt
        check(fvk_vec_axpby(int64_t(source.size()) * nComponents<ValueType>(), -dt, source.raw(), 1.0, sol.internalVector().raw(), sol.exec().stream()));
        sol.correctBoundaryConditions();
        sol.exec().sync();
    }
    std::unique_ptr<TimeIntegratorBase<SolutionVectorType>> clone() const override { return std::make_unique<ForwardEuler>(*this); }
};

// backwardEuler.hpp:41-60: the explicit source is evaluated and dropped (sic, :45); fresh system <- implicit spatial, then temporal
template<typename SolutionVectorType>
class BackwardEuler : public TimeIntegratorBase<SolutionVectorType>::template Register<BackwardEuler<SolutionVectorType>>
{
    using Base = typename TimeIntegratorBase<SolutionVectorType>::template Register<BackwardEuler<SolutionVectorType>>;
    using ValueType = typename SolutionVectorType::ElementType;
public:
    BackwardEuler(const Dictionary& schemeDict, const Dictionary& solutionDict) : Base(schemeDict, solutionDict) {}
    static std::string name() { return "backwardEuler"; }
    static std::string doc() { return "first order time integration method"; }
    static std::string schema() { return "none"; }
    void solve(dsl::Expression<ValueType>& eqn, SolutionVectorType& sol, scalar t, scalar dt) override
    {
        if (eqn.hasExplicit()) (void) eqn.explicitOperation(sol.exec(), size_t(sol.mesh().nCells()));
        auto sp = la::SparsityPattern::readOrCreate(sol.mesh());
        la::LinearSystem<ValueType, localIdx> ls(sol.mesh(), sp, false);
        eqn.assemble(t, dt, sp, ls, sol);
        la::Solver solver(sol.exec(), this->solutionDict_);
        if constexpr (std::is_same_v<ValueType, scalar>) this->lastStats = solver.solve(ls, sol.internalVector());
        else this->lastStats = solver.solve(ls, sol.internalVector())[0];
        sol.exec().sync();
    }
    std::unique_ptr<TimeIntegratorBase<SolutionVectorType>> clone() const override { return std::make_unique<BackwardEuler>(*this); }
};

// rungeKutta.hpp / rungeKutta.cpp:34-57 over SUNDIALS ERKStep with a fixed step. Only the 1-stage Forward-Euler table is usable in
// the reference (sundials.hpp:59-78: Heun / Midpoint exit with "Currently unsupported ..."); one ERK step with it is y + dt f(t, y),
// f = -explicitOperation (sundials.hpp:196-215): forwardEuler without the boundary correction, then old = new (:55-56).
template<typename SolutionVectorType>
class RungeKutta : public TimeIntegratorBase<SolutionVectorType>::template Register<RungeKutta<SolutionVectorType>>
{
    using Base = typename TimeIntegratorBase<SolutionVectorType>::template Register<RungeKutta<SolutionVectorType>>;
    using ValueType = typename SolutionVectorType::ElementType;
public:
    RungeKutta(const Dictionary& schemeDict, const Dictionary& solutionDict) : Base(schemeDict, solutionDict)
    {
        const auto method = schemeDict.get<std::string>("Runge-Kutta-Method");
        if (method == "Heun" || method == "Midpoint") NF_ERROR_EXIT("Currently unsupported until field time step-stage indexing resolved.");
        if (method != "Forward-Euler")
            NF_ERROR_EXIT("Unsupported Runge-Kutta time integration method selectied: " + method + ".\nSupported methods are: Forward-Euler, Heun, Midpoint.");
    }
    static std::string name() { return "Runge-Kutta"; }
    static std::string doc() { return "Explicit time integration using the Runge-Kutta method."; }
    static std::string schema() { return "none"; }
    void solve(dsl::Expression<ValueType>& eqn, SolutionVectorType& sol, scalar, scalar dt) override
    {
        auto& old = sol.oldTime();
        auto source = eqn.explicitOperation(sol.exec(), size_t(sol.mesh().nCells()));
        sol.internalVector() = old.internalVector();
        check(fvk_vec_axpby(int64_t(source.size()) * nComponents<ValueType>(), -dt, source.raw(), 1.0, sol.internalVector().raw(), sol.exec().stream()));
        old.internalVector() = sol.internalVector();
        sol.exec().sync();
    }
    std::unique_ptr<TimeIntegratorBase<SolutionVectorType>> clone() const override { return std::make_unique<RungeKutta>(*this); }
};
NF_REGISTER((TimeIntegratorBase<finiteVolume::cellCentred::VolumeField<scalar>>), (ForwardEuler<finiteVolume::cellCentred::VolumeField<scalar>>));
NF_REGISTER((TimeIntegratorBase<finiteVolume::cellCentred::VolumeField<scalar>>), (BackwardEuler<finiteVolume::cellCentred::VolumeField<scalar>>));
NF_REGISTER((TimeIntegratorBase<finiteVolume::cellCentred::VolumeField<scalar>>), (RungeKutta<finiteVolume::cellCentred::VolumeField<scalar>>));
NF_REGISTER((TimeIntegratorBase<finiteVolume::cellCentred::VolumeField<Vec3>>), (ForwardEuler<finiteVolume::cellCentred::VolumeField<Vec3>>));
NF_REGISTER((TimeIntegratorBase<finiteVolume::cellCentred::VolumeField<Vec3>>), (BackwardEuler<finiteVolume::cellCentred::VolumeField<Vec3>>));

// timeIntegration.hpp:58-95
template<typename SolutionVectorType>
class TimeIntegration
{
public:
    using ValueType = typename SolutionVectorType::ElementType;
    TimeIntegration(const Dictionary& schemeDict, const Dictionary& solutionDict)
        : strategy_(TimeIntegratorBase<SolutionVectorType>::create(schemeDict.get<std::string>("type"), schemeDict, solutionDict)) {}
    TimeIntegration(const TimeIntegration& o) : strategy_(o.strategy_->clone()) {}
    void solve(dsl::Expression<ValueType>& eqn, SolutionVectorType& sol, scalar t, scalar dt) { strategy_->solve(eqn, sol, t, dt); }
    const la::SolverStats& stats() const { return strategy_->lastStats; }
private:
    std::unique_ptr<TimeIntegratorBase<SolutionVectorType>> strategy_;
};
} // namespace NeoN::timeIntegration

namespace NeoN::dsl
{
// dsl::solve (dsl/solver.hpp:35-82)
template<typename VectorType>
la::SolverStats solve(Expression<typename VectorType::ElementType>& exp, VectorType& solution, scalar t, scalar dt, const Dictionary& fvSchemes,
                      const Dictionary& fvSolution)
{
    using ValueType = typename VectorType::ElementType;
    if (exp.temporalOperators().size() == 0 && exp.spatialOperators().size() == 0) NF_ERROR_EXIT("No temporal or implicit terms to solve.");
    exp.read(fvSchemes);
    if (exp.temporalOperators().size() > 0)
    {
        timeIntegration::TimeIntegration<VectorType> timeIntegrator(fvSchemes.subDict("ddtSchemes"), fvSolution);
        timeIntegrator.solve(exp, solution, t, dt);
        return timeIntegrator.stats();
    }
    const auto& mesh = solution.mesh();
    auto sp = la::SparsityPattern::readOrCreate(mesh);
    la::LinearSystem<ValueType, localIdx> ls(mesh, sp, false);
    la::Solver solver(solution.exec(), fvSolution);
    if constexpr (std::is_same_v<ValueType, scalar>)
        return detail::iterativeSolveImpl(exp, sp, ls, solution, t, dt, solver, [](const la::SparsityPattern&, la::LinearSystem<ValueType, localIdx>&) {});
    else
        return detail::iterativeSolveImpl(exp, sp, ls, solution, t, dt, solver, [](const la::SparsityPattern&, la::LinearSystem<ValueType, localIdx>&) {})[0];
}
} // namespace NeoN::dsl
