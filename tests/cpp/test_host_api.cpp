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
            la::LinearSystem<scalar> ls(mesh, sp, true);
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
            // mapFvSolution: PCG + DIC -> Cg + scalar Jacobi, tolerance -> absolute_residual_norm
            auto mapped = FoamAdapter::mapFvSolution(Dictionary {{"solver", std::string("PCG")}, {"preconditioner", std::string("DIC")}, {"tolerance", 1e-6}, {"relTol", 0.0}});
            EXPECT(mapped.get<std::string>("type") == "solver::Cg");
            EXPECT(mapped.subDict("preconditioner").get<std::string>("type") == "preconditioner::Jacobi");
            EXPECT(mapped.subDict("criteria").get<scalar>("absolute_residual_norm") == 1e-6);
            EXPECT(mapped.subDict("criteria").get<int>("iteration") == 1000);
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
