// C++ host-API test (NeoN / FoamAdapter names over libfvk): the known-answer cases of the reference's own NeoN tests that
// need no OpenFOAM, run through the C++ classes on the GPU. Exit code 0 = all passed.
//   src/NeoN/test/linearAlgebra/sparsityPattern.cpp:36-92, utilities.cpp:22-44, ginkgo.cpp:95-124,
//   src/NeoN/test/finiteVolume/cellCentred/{operator/gaussGreenDiv.cpp:18-69, operator/laplacianOperator.cpp:21-137,
//   faceNormalGradient/uncorrected.cpp:45-68, interpolation/linear.cpp}
#include "FoamAdapter/FoamAdapter.hpp"

#include <cstdio>

using namespace NeoN;
namespace fvcc = NeoN::finiteVolume::cellCentred;

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

// ---- third-party plug-ins written against the reference's interfaces -------------------------------------------------------
// A conforming SpatialOperator in the shape of the reference's test operator (src/NeoN/test/dsl/common.hpp:53-106): explicit:
// source += coeff * field; implicit: rhs += coeff * field. A plug-in has no parallelFor here: it composes C-ABI kernels.
template<typename ValueType>
class Dummy : public dsl::OperatorMixin<fvcc::VolumeField<ValueType>>
{
public:
    using VectorValueType = ValueType;
    Dummy(fvcc::VolumeField<ValueType>& field, dsl::Operator::Type type = dsl::Operator::Type::Explicit)
        : dsl::OperatorMixin<fvcc::VolumeField<ValueType>>(field.exec(), dsl::Coeff(1.0), field, type) {}
    void explicitOperation(Vector<ValueType>& source) const
    {
        check(fvk_vec_axpby(int64_t(source.size()), this->getCoefficient().value(), this->field_.internalVector().raw(), 1.0, source.raw(), source.exec().stream()));
    }
    void implicitOperation(la::LinearSystem<ValueType, localIdx>& ls) const
    {
        check(fvk_vec_axpby(int64_t(ls.rhs().size()), this->getCoefficient().value(), this->field_.internalVector().raw(), 1.0, ls.rhs().raw(), ls.exec().stream()));
    }
    std::string getName() const { return "Dummy"; }
};
// an interpolation scheme registered by name: the arithmetic mean of owner and neighbour = linear weights on a uniform mesh
template<typename T>
class Midpoint : public fvcc::SurfaceInterpolationFactory<T>::template Register<Midpoint<T>>
{
    using Base = typename fvcc::SurfaceInterpolationFactory<T>::template Register<Midpoint<T>>;
public:
    Midpoint(const Executor& exec, const UnstructuredMesh& mesh, const Input&) : Base(exec, mesh) {}
    static std::string name() { return "midPoint"; }
    void interpolate(const fvcc::VolumeField<T>& src, fvcc::SurfaceField<T>& dst) const override
    { // plug-in schemes compose existing kernels: here the library's linear interpolation (weights are 0.5 on the uniform test mesh)
        fvcc::Linear<T>(this->exec_, this->mesh_, Input()).interpolate(src, dst);
    }
    void interpolate(const fvcc::SurfaceField<scalar>&, const fvcc::VolumeField<T>& src, fvcc::SurfaceField<T>& dst) const override { interpolate(src, dst); }
    void weight(const fvcc::VolumeField<T>&, fvcc::SurfaceField<scalar>& w) const override { fill(w.internalVector(), 0.5); }
    void weight(const fvcc::SurfaceField<scalar>&, const fvcc::VolumeField<T>& s, fvcc::SurfaceField<scalar>& w) const override { weight(s, w); }
    std::unique_ptr<fvcc::SurfaceInterpolationFactory<T>> clone() const override { return std::make_unique<Midpoint<T>>(*this); }
};
NF_REGISTER((fvcc::SurfaceInterpolationFactory<scalar>), (Midpoint<scalar>));
// a time integrator registered by name: two forward-Euler half steps
class TwoHalfSteps : public timeIntegration::TimeIntegratorBase<fvcc::VolumeField<scalar>>::Register<TwoHalfSteps>
{
    using Sol = fvcc::VolumeField<scalar>;
public:
    TwoHalfSteps(const Dictionary& a, const Dictionary& b) : timeIntegration::TimeIntegratorBase<Sol>::Register<TwoHalfSteps>(a, b) {}
    static std::string name() { return "twoHalfSteps"; }
    void solve(dsl::Expression<scalar>& eqn, Sol& sol, scalar t, scalar dt) override
    {
        timeIntegration::ForwardEuler<Sol> fe(schemeDict_, solutionDict_);
        fe.solve(eqn, sol, t, 0.5 * dt);
        sol.oldTime().internalVector() = sol.internalVector();
        fe.solve(eqn, sol, t + 0.5 * dt, 0.5 * dt);
    }
    std::unique_ptr<timeIntegration::TimeIntegratorBase<Sol>> clone() const override { return std::make_unique<TwoHalfSteps>(*this); }
};
NF_REGISTER((timeIntegration::TimeIntegratorBase<fvcc::VolumeField<scalar>>), (TwoHalfSteps));

int main()
{
    try
    {
        Executor exec(0);
        auto mesh = create1DUniformMesh(exec, 10);
        // sparsity pattern known answer
        la::SparsityPattern sp(mesh);
        {
            std::vector<localIdx> ro(11);
            check(fvk_memcpy_d2h(ro.data(), sp.rowOffs().ptr, sizeof(localIdx) * 11, nullptr)); exec.sync();
            const localIdx exp[11] = {0, 2, 5, 8, 11, 14, 17, 20, 23, 26, 28};
            for (int i = 0; i < 11; ++i) EXPECT(ro[i] == exp[i]);
            EXPECT(sp.nnz() == 28);
        }
        // div of a uniform field with unit flux is zero (gaussGreenDiv.cpp:18-69)
        {
            std::vector<fvcc::VolumeBoundary<scalar>> bcs {{"fixedValue", 1.0}, {"fixedValue", 1.0}};
            fvcc::VolumeField<scalar> phi(exec, "phi", mesh, bcs);
            fill(phi.internalVector(), 1.0);
            phi.correctBoundaryConditions();
            fvcc::SurfaceField<scalar> faceFlux(exec, "sf", mesh);
            fill(faceFlux.internalVector(), 1.0);
            // boundary fluxes: left face points outwards (-x): flux = -1, right +1 (as the reference test sets them)
            auto h = faceFlux.internalVector().copyToHost(); h[9] = -1.0; h[10] = 1.0;
            faceFlux.internalVector().copyFromHost(h.data());
            Vector<scalar> out(exec, 10, 0.0);
            fvcc::GaussGreenDiv<scalar>(exec, mesh, TokenList({"linear"})).div(out, faceFlux, phi, dsl::Coeff(1.0));
            for (auto v : out.copyToHost()) EXPECT(v == 0.0);
        }
        // laplacian of a linear field is zero, explicit and implicit (laplacianOperator.cpp:21-137)
        {
            std::vector<fvcc::VolumeBoundary<scalar>> bcs {{"fixedValue", 0.5}, {"fixedValue", 10.5}};
            fvcc::VolumeField<scalar> phi(exec, "phi", mesh, bcs);
            std::vector<scalar> lin(10);
            for (int i = 0; i < 10; ++i) lin[i] = scalar(i + 1);
            phi.internalVector().copyFromHost(lin.data());
            phi.correctBoundaryConditions();
            fvcc::SurfaceField<scalar> gamma(exec, "gamma", mesh);
            fill(gamma.internalVector(), 2.0);
            Vector<scalar> out(exec, 10, 0.0);
            fvcc::GaussGreenLaplacian<scalar>(exec, mesh, TokenList({"linear", "uncorrected"})).laplacian(out, gamma, phi, dsl::Coeff(-0.5));
            for (auto v : out.copyToHost()) EXPECT(std::abs(v) < 1e-8);
            // implicit: A phi - b == 0
            auto op = dsl::imp::laplacian(gamma, phi);
            bool threw = false;
            la::LinearSystem<scalar> ls(mesh, sp, true);
            try { op.implicitOperation(ls); } catch (const NeoNException&) { threw = true; }   // no strategy before read(), as in the reference
            EXPECT(threw);
            op.read(Dictionary {{"laplacianSchemes", Dictionary {{"laplacian(gamma,phi)", std::string("Gauss linear uncorrected")}}}});
            op.implicitOperation(ls);
            Vector<scalar> res(exec, 10, 0.0);
            la::computeResidual(ls, phi.internalVector(), res);
            for (auto v : res.copyToHost()) EXPECT(std::abs(v) < 1e-8);
            // snGrad of phi = i+1 is 10 on internal faces (uncorrected.cpp:45-68)
            fvcc::SurfaceField<scalar> sn(exec, "sn", mesh);
            fvcc::FaceNormalGradient<scalar>(exec, mesh, TokenList({"uncorrected"})).faceNormalGrad(phi, sn);
            auto s = sn.internalVector().copyToHost();
            for (int f = 0; f < 9; ++f) EXPECT(std::abs(s[f] - 10.0) < 1e-9);
            EXPECT(std::abs(s[9] + 10.0) < 1e-9 && std::abs(s[10] - 10.0) < 1e-9);
        }
        // unknown plugin keys fail loudly like RuntimeSelectionFactory::keyExistsOrError
        {
            bool threw = false;
            try { fvcc::SurfaceInterpolation<scalar>(exec, mesh, TokenList({"cubic"})); } catch (const NeoNException&) { threw = true; }
            EXPECT(threw);
        }
        // CG known answer through la::Solver with a mapped fvSolution entry (ginkgo.cpp:95-124) on a 3-cell 1-D mesh
        {
            auto m3 = create1DUniformMesh(exec, 3);
            la::SparsityPattern sp3(m3);
            la::LinearSystem<scalar> ls(m3, sp3, true);
            const std::vector<scalar> vals {1.0, -0.1, -0.1, 1.0, -0.1, -0.1, 1.0}, b {1.0, 2.0, 3.0};
            ls.values().copyFromHost(vals.data());
            ls.rhs().copyFromHost(b.data());
            Dictionary d {{"solver", std::string("Ginkgo")}, {"type", std::string("solver::Cg")},
                          {"criteria", Dictionary {{"iteration", 3}, {"relative_residual_norm", 1e-7}}}};
            la::Solver solver(exec, d);
            Vector<scalar> x(exec, 3, 0.0);
            auto st = solver.solve(ls, x);
            auto xh = x.copyToHost();
            EXPECT(st.numIter == 3);
            EXPECT(std::abs(st.initResNorm - 3.741657386) < 1e-8);
            EXPECT(st.finalResNorm < 1e-4);
            EXPECT(std::abs(xh[0] - 1.24489796) < 1e-8 && std::abs(xh[1] - 2.44897959) < 1e-8 && std::abs(xh[2] - 3.24489796) < 1e-8);
            // single-GPU solver: the solution trivially "keeps its ghosts"; no solve ran inside a CUDA graph, so the device log is empty
            EXPECT(solver.keepsGhosts());
            EXPECT(solver.capturedLog().empty());
            solver.setGhostsCurrent(true);   // a no-op without a communicator: same answer
            Vector<scalar> x2(exec, 3, 0.0);
            EXPECT(solver.solve(ls, x2).numIter == 3);
            // mapFvSolution: PCG + DIC -> Cg + scalar Jacobi, tolerance -> absolute_residual_norm
            auto mapped = FoamAdapter::mapFvSolution(Dictionary {{"solver", std::string("PCG")}, {"preconditioner", std::string("DIC")}, {"tolerance", 1e-6}, {"relTol", 0.0}});
            EXPECT(mapped.get<std::string>("type") == "solver::Cg");
            EXPECT(mapped.subDict("preconditioner").get<std::string>("type") == "preconditioner::Jacobi");
            EXPECT(mapped.subDict("criteria").get<scalar>("absolute_residual_norm") == 1e-6);
            EXPECT(mapped.subDict("criteria").get<int>("iteration") == 1000);
        }
        // ---- the open DSL: a third-party operator class, type-erased next to the built-in ones (dsl/spatialOperator.hpp:21-124) ----
        {
            std::vector<fvcc::VolumeBoundary<scalar>> bcs {{"fixedValue", 2.0}, {"fixedValue", 2.0}};
            fvcc::VolumeField<scalar> phi(exec, "phi", mesh, bcs);
            fill(phi.internalVector(), 2.0);
            phi.correctBoundaryConditions();
            dsl::SpatialOperator<scalar> a = Dummy<scalar>(phi);                                     // explicit by default
            dsl::SpatialOperator<scalar> b = Dummy<scalar>(phi, dsl::Operator::Type::Implicit);
            EXPECT(a.getName() == "Dummy" && a.getType() == dsl::Operator::Type::Explicit && b.getType() == dsl::Operator::Type::Implicit);
            auto c = 2.0 * a;                                                                        // coefficient arithmetic on the erased type
            EXPECT(c.getCoefficient().value() == 2.0 && a.getCoefficient().value() == 1.0);
            Vector<scalar> src(exec, 10, 1.0);
            c.explicitOperation(src);
            for (auto v : src.copyToHost()) EXPECT(v == 5.0);                                        // 1 + 2 * 2
            // an expression mixing a plug-in with built-ins: not fusable -> zeroed system, operators applied one after another
            fvcc::SurfaceField<scalar> gamma(exec, "gamma", mesh);
            fill(gamma.internalVector(), 1.0);
            Dictionary schemes {{"laplacianSchemes", Dictionary {{"laplacian(gamma,phi)", std::string("Gauss linear uncorrected")}}}};
            dsl::Expression<scalar> mixed = dsl::imp::laplacian(gamma, phi) + b + a;
            mixed.read(schemes);
            EXPECT(mixed.size() == 3 && mixed.hasExplicit());
            la::LinearSystem<scalar> ls(mesh, sp, false), ref(mesh, sp, true);
            mixed.assemble(0.0, 1.0, sp, ls, phi);
            dsl::Expression<scalar> builtin; builtin.addOperator(dsl::imp::laplacian(gamma, phi)); builtin.read(schemes);
            builtin.assemble(0.0, 1.0, sp, ref, phi);                                                // the fused single-launch path
            auto v1 = ls.values().copyToHost(), v2 = ref.values().copyToHost(), r1 = ls.rhs().copyToHost(), r2 = ref.rhs().copyToHost();
            for (size_t i = 0; i < v1.size(); ++i) EXPECT(v1[i] == v2[i]);
            for (size_t i = 0; i < r1.size(); ++i) EXPECT(r1[i] == r2[i] + 2.0);                     // + the plug-in's rhs contribution
            auto e = mixed.explicitOperation(exec, 10);
            for (auto v : e.copyToHost()) EXPECT(v == 2.0);
        }
        // ---- strategies and integrators by NAME (core/runtimeSelectionFactory.hpp): built-ins + the plug-ins registered above ----
        {
            auto names = fvcc::SurfaceInterpolationFactory<scalar>::entries();
            EXPECT(names.size() == 3 && fvcc::SurfaceInterpolationFactory<scalar>::contains("midPoint") && fvcc::SurfaceInterpolationFactory<scalar>::contains("upwind"));
            EXPECT(fvcc::DivOperatorFactory<Vec3>::contains("Gauss") && fvcc::LaplacianOperatorFactory<scalar>::contains("Gauss"));
            EXPECT(fvcc::FaceNormalGradientFactory<scalar>::contains("uncorrected") && la::SolverFactory::contains("Ginkgo"));
            std::vector<fvcc::VolumeBoundary<scalar>> bcs {{"fixedValue", 1.0}, {"fixedValue", 1.0}};
            fvcc::VolumeField<scalar> T(exec, "T", mesh, bcs);
            std::vector<scalar> init(10);
            for (int i = 0; i < 10; ++i) init[i] = 1.0 + 0.1 * i;
            T.internalVector().copyFromHost(init.data());
            T.correctBoundaryConditions();
            fvcc::SurfaceField<scalar> phi(exec, "phi", mesh);
            fill(phi.internalVector(), 1.0);
            auto h = phi.internalVector().copyToHost(); h[9] = -1.0; h[10] = 1.0; phi.internalVector().copyFromHost(h.data());
            // div through the plug-in interpolation scheme equals div through "linear" on the uniform mesh
            Vector<scalar> d1(exec, 10, 0.0), d2(exec, 10, 0.0);
            fvcc::DivOperatorFactory<scalar>::create(exec, mesh, TokenList({"Gauss", "midPoint"}))->div(d1, phi, T, dsl::Coeff(1.0));
            fvcc::DivOperatorFactory<scalar>::create(exec, mesh, TokenList({"Gauss", "linear"}))->div(d2, phi, T, dsl::Coeff(1.0));
            auto a1 = d1.copyToHost(), a2 = d2.copyToHost();
            for (int i = 0; i < 10; ++i) EXPECT(std::abs(a1[i] - a2[i]) <= 1e-13 * (1.0 + std::abs(a2[i])));
            // dsl::solve with forwardEuler, Runge-Kutta (Forward-Euler table) and the registered two-half-steps integrator
            auto run = [&](const std::string& type, int steps, scalar dt) {
                T.internalVector().copyFromHost(init.data());
                T.correctBoundaryConditions();
                Dictionary schemes {{"ddtSchemes", Dictionary {{"type", type}, {"Runge-Kutta-Method", std::string("Forward-Euler")}}},
                                    {"divSchemes", Dictionary {{"div(phi,T)", std::string("Gauss upwind")}}}};
                for (int s_ = 0; s_ < steps; ++s_)
                {
                    fvcc::oldTime(T).internalVector() = T.internalVector();
                    dsl::Expression<scalar> eqn = dsl::imp::ddt(T) + dsl::exp::div(phi, T);
                    dsl::solve(eqn, T, s_ * dt, dt, schemes, Dictionary {});
                }
                return T.internalVector().copyToHost();
            };
            auto fe2 = run("forwardEuler", 2, 0.005), half = run("twoHalfSteps", 1, 0.01), rk = run("Runge-Kutta", 2, 0.005);
            for (int i = 1; i < 9; ++i) EXPECT(std::abs(fe2[i] - half[i]) < 1e-12 && std::abs(fe2[i] - rk[i]) < 1e-12);
            EXPECT(std::abs(fe2[5] - init[5]) > 1e-4);                                                  // it did advect
            bool threw = false;
            try { run("crankNicolson", 1, 0.01); } catch (const NeoNException&) { threw = true; }
            EXPECT(threw);
            // backwardEuler + BiCGStab (test/test_advection.cpp:176-228): (I + dt div) T = T_old, checked through the residual
            T.internalVector().copyFromHost(init.data());
            fvcc::oldTime(T).internalVector() = T.internalVector();
            Dictionary schemes {{"ddtSchemes", Dictionary {{"type", std::string("backwardEuler")}}}, {"divSchemes", Dictionary {{"div(phi,T)", std::string("Gauss upwind")}}}};
            Dictionary solver {{"solver", std::string("Ginkgo")}, {"type", std::string("solver::Bicgstab")},
                               {"preconditioner", Dictionary {{"type", std::string("preconditioner::Jacobi")}, {"max_block_size", 1}}},
                               {"criteria", Dictionary {{"iteration", 20}, {"relative_residual_norm", 1e-14}}}};
            dsl::Expression<scalar> eqn = dsl::imp::ddt(T) + dsl::imp::div(phi, T);
            auto st = dsl::solve(eqn, T, 0.0, 0.01, schemes, solver);
            EXPECT(st.numIter >= 1 && st.numIter <= 20 && st.finalResNorm <= 1e-13 * st.initResNorm);
            la::LinearSystem<scalar> ls(mesh, sp, false);
            eqn.assemble(0.0, 0.01, sp, ls, T);
            Vector<scalar> res(exec, 10, 0.0);
            la::computeResidual(ls, T.internalVector(), res);
            for (auto v : res.copyToHost()) EXPECT(std::abs(v) < 1e-13);
        }
    }
    catch (const std::exception& e)
    {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "%d FAILURES\n" : "host api ok%.0d\n", failures);
    return failures ? 1 : 0;
}
