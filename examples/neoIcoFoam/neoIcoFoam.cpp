// neoIcoFoam on the B200 build: the time loop of the reference's examples/neoIcoFoam/neoIcoFoam.cpp:80-180, line for line in
// terms of the FoamAdapter / NeoN API, on the lid-driven cavity of tutorials/cavity generated synthetically
// (N x N x 1 with empty front/back, or N^3 with --3d). Everything inside the loop runs on the GPU through libfvk.
//   build: g++ -std=c++17 -O2 -Iinclude examples/neoIcoFoam/neoIcoFoam.cpp -Lfoamadapter_b200/lib -lfvk -Wl,-rpath,$PWD/foamadapter_b200/lib -o neoIcoFoam
//   run:   ./neoIcoFoam [N=20] [steps=5] [--3d]
#include "FoamAdapter/FoamAdapter.hpp"

#include <cstdlib>
#include <cstring>

namespace nf = FoamAdapter;
namespace fvcc = NeoN::finiteVolume::cellCentred;
namespace dsl = NeoN::dsl;

int main(int argc, char* argv[])
{
    int N = 20, steps = 5;
    bool threeD = false;
    int pos = 0;
    for (int i = 1; i < argc; ++i)
    {
        if (!std::strcmp(argv[i], "--3d")) threeD = true;
        else if (pos++ == 0) N = std::atoi(argv[i]);
        else steps = std::atoi(argv[i]);
    }
    try
    {
        NeoN::Executor exec(0);
        // tutorials/cavity/system/blockMeshDict: scale 0.1, movingWall (y-max), fixedWalls, frontAndBack empty
        std::vector<NeoN::BlockPatch> patches {{"movingWall", {3}, false}};
        if (threeD) patches.push_back({"fixedWalls", {0, 1, 2, 4, 5}, false});
        else { patches.push_back({"fixedWalls", {0, 1, 2}, false}); patches.push_back({"frontAndBack", {4, 5}, true}); }
        auto mesh = NeoN::UnstructuredMesh::createBlockMesh(exec, N, N, threeD ? N : 1, 0.1, 0.1, threeD ? 0.1 : 0.01, patches);

        nf::RunTime rt {exec, mesh};
        rt.dt = 1e-4 * 20.0 / N; // tutorials/cavity/system/controlDict deltaT at N=20... scaled with the mesh
        // tutorials/cavity/system/{fvSchemes,fvSolution}
        rt.fvSchemesDict.insert("ddtSchemes", NeoN::Dictionary {{"type", std::string("backwardEuler")}});
        rt.fvSchemesDict.insert("divSchemes", NeoN::Dictionary {{"div(phi,U)", std::string("Gauss linear")}});
        rt.fvSchemesDict.insert("laplacianSchemes", NeoN::Dictionary {{"laplacian(nu,U)", std::string("Gauss linear uncorrected")},
                                                                      {"laplacian(rAUf,p)", std::string("Gauss linear uncorrected")}});
        NeoN::Dictionary pSolver {{"solver", std::string("PCG")}, {"preconditioner", std::string("DIC")}, {"tolerance", 1e-6}, {"relTol", 0.0}};
        rt.fvSolutionDict.insert("solvers", NeoN::Dictionary {{"p", nf::mapFvSolution(pSolver)}});
        const int nCorrectors = 2, nNonOrthCorr = 0;
        const bool momentumPredictor = false;
        const NeoN::localIdx pRefCell = 0;
        const NeoN::scalar pRefValue = 0.0, viscosity = 0.01;

        std::vector<fvcc::VolumeBoundary<NeoN::Vec3>> UBCs {{"fixedValue", NeoN::Vec3(1.0, 0.0, 0.0)}, {"noSlip", NeoN::zero<NeoN::Vec3>()}};
        std::vector<fvcc::VolumeBoundary<NeoN::scalar>> pBCs {{"zeroGradient", 0.0}, {"zeroGradient", 0.0}};
        fvcc::VolumeField<NeoN::scalar> p(exec, "p", mesh, pBCs);
        fvcc::VolumeField<NeoN::Vec3> U(exec, "U", mesh, UBCs);
        p.correctBoundaryConditions();
        U.correctBoundaryConditions();
        fvcc::SurfaceField<NeoN::scalar> nu(exec, "nu", mesh);
        NeoN::fill(nu.internalVector(), viscosity);
        NeoN::fill(nu.boundaryData().value(), viscosity);
        auto phi = nf::flux(U);
        phi.name = "phi";

        std::cout << "\nStarting time loop\n" << std::endl;
        for (int step = 0; step < steps; ++step)
        {
            rt.t += rt.dt;
            std::cout << "Time = " << rt.t << "\n" << std::endl;
            auto& oldU = fvcc::oldTime(U);
            oldU.internalVector() = U.internalVector();
            fvcc::computeCoNum(phi, rt.dt);

            // Momentum predictor
            nf::PDESolver<NeoN::Vec3> UEqn(dsl::imp::ddt(U) + dsl::imp::div(phi, U) - dsl::imp::laplacian(nu, U), U, rt);
            if (momentumPredictor) NF_ERROR_EXIT("momentumPredictor yes is outside the hot path");
            UEqn.assemble();

            // --- PISO loop
            for (int corr = 0; corr < nCorrectors; ++corr)
            {
                std::cout << "PISO loop" << std::endl;
                auto [crAU, hByA] = nf::computeRAUandHByA(UEqn);
                nf::constrainHbyA(U, p, hByA);
                fvcc::SurfaceField<NeoN::scalar> rAU =
                    fvcc::SurfaceInterpolation<NeoN::scalar>(exec, mesh, NeoN::TokenList({std::string("linear")})).interpolate(crAU);
                rAU.name = "rAUf";
                auto phiHbyA = nf::flux(hByA);

                // Non-orthogonal pressure corrector loop
                for (int nonOrth = 0; nonOrth <= nNonOrthCorr; ++nonOrth)
                {
                    nf::PDESolver<NeoN::scalar> pEqn(dsl::imp::laplacian(rAU, p) - dsl::exp::div(phiHbyA), p, rt);
                    if (pRefCell >= 0) pEqn.setReference(pRefCell, pRefValue);
                    auto stats = pEqn.solve();
                    (void) stats;
                    p.correctBoundaryConditions();
                    if (nonOrth == nNonOrthCorr) nf::updateFaceVelocity(phiHbyA, pEqn, phi);
                }
                nf::updateVelocity(hByA, crAU, p, U);
                U.correctBoundaryConditions();
            }
        }
        auto Uh = U.internalVector().copyToHost();
        auto ph = p.internalVector().copyToHost();
        double umax = 0, pmin = 1e300, pmax = -1e300;
        for (auto& u : Uh) umax = std::max(umax, NeoN::mag(u));
        for (auto& v : ph) { pmin = std::min(pmin, v); pmax = std::max(pmax, v); }
        std::cout.precision(17);
        std::cout << "End: cells " << mesh.nCells() << " max|U| " << umax << " p range [" << pmin << ", " << pmax << "]" << std::endl;
        // machine-readable dump for the parity test
        if (const char* out = std::getenv("NEOICOFOAM_DUMP"))
        {
            FILE* f = std::fopen(out, "wb");
            if (f)
            {
                std::fwrite(Uh.data(), sizeof(NeoN::Vec3), Uh.size(), f);
                std::fwrite(ph.data(), sizeof(double), ph.size(), f);
                std::fclose(f);
            }
        }
    }
    catch (const std::exception& e)
    {
        std::cerr << "neoIcoFoam: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
