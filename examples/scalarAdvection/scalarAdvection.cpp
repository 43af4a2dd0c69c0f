// scalarAdvection on the B200 kernels: the reference's time loop (examples/scalarAdvection/scalarAdvection.cpp:52-95, the loop of
// test/test_advection.cpp:118-166,185-227) written against include/NeoN + include/FoamAdapter. OpenFOAM's part of the example
// (case reading, createFields.H) is replaced by the block-mesh generator and the same initial fields evaluated on the host:
//   U = (-sin(2 pi y) sin^2(pi x), sin(2 pi x) sin^2(pi y), 0), T = exp(-0.5 (((x-0.5)/s)^2 + ((y-0.75)/s)^2)), s = 0.05
//   (createFields.H:27-50), phi = linearInterpolate(U) & Sf, zeroGradient walls (tutorials/scalarAdvection/0.orig).
// usage: scalarAdvection [N=50] [steps=100] [forwardEuler|backwardEuler|Runge-Kutta] [--3d]      (dt = 0.1/N, endTime 3)
#include "FoamAdapter/FoamAdapter.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dsl = NeoN::dsl;
namespace fvcc = NeoN::finiteVolume::cellCentred;
namespace nf = FoamAdapter;

int main(int argc, char* argv[])
{
    try
    {
        int N = 50, steps = 100;
        std::string integrator = "forwardEuler";
        bool threeD = false;
        int pos = 0;
        for (int i = 1; i < argc; ++i)
        {
            if (std::strcmp(argv[i], "--3d") == 0) threeD = true;
            else if (pos == 0) { N = std::atoi(argv[i]); ++pos; }
            else if (pos == 1) { steps = std::atoi(argv[i]); ++pos; }
            else integrator = argv[i];
        }
        NeoN::Executor exec(0);
        // tutorials/scalarAdvection/system/blockMeshDict: fixedWalls = y-max, x-min, x-max, y-min; frontAndBack empty
        std::vector<NeoN::BlockPatch> patches;
        if (threeD) patches.push_back({"fixedWalls", {3, 0, 1, 2, 4, 5}, false});
        else { patches.push_back({"fixedWalls", {3, 0, 1, 2}, false}); patches.push_back({"frontAndBack", {4, 5}, true}); }
        auto mesh = NeoN::UnstructuredMesh::createBlockMesh(exec, N, N, threeD ? N : 1, 1.0, 1.0, threeD ? 1.0 : 0.1, patches);

        nf::RunTime rt {exec, mesh};
        rt.dt = 0.1 / N;
        const NeoN::scalar endTime = 3.0;
        // tutorials/scalarAdvection/system/fvSchemes (div(phi,nfT) Gauss upwind) + the integrators of test/test_advection.cpp
        rt.fvSchemesDict.insert("ddtSchemes", NeoN::Dictionary {{"type", integrator}, {"Runge-Kutta-Method", std::string("Forward-Euler")}});
        rt.fvSchemesDict.insert("divSchemes", NeoN::Dictionary {{"div(phi,nfT)", std::string("Gauss upwind")}});
        // test_advection.cpp:176-183
        NeoN::Dictionary fvSolution {{"solver", std::string("Ginkgo")}, {"type", std::string("solver::Bicgstab")},
                                     {"preconditioner", NeoN::Dictionary {{"type", std::string("preconditioner::Jacobi")}, {"max_block_size", 1}}},
                                     {"criteria", NeoN::Dictionary {{"iteration", 20}, {"relative_residual_norm", 1e-14}}}};

        // createFields.H:27-50 on the host
        const auto nC = size_t(mesh.nCells());
        std::vector<NeoN::Vec3> C(nC);
        NeoN::check(fvk_memcpy_d2h(C.data(), mesh.cellCentres().ptr, nC * sizeof(NeoN::Vec3), nullptr));
        exec.sync();
        std::vector<NeoN::Vec3> Uh(nC);
        std::vector<NeoN::scalar> Th(nC);
        const double spread = 0.05, pi = M_PI;
        const size_t nxy = size_t(N) * N;
        for (size_t c = 0; c < nC; ++c)
        {
            const size_t col = c % nxy; // the fields depend on (x, y) only: every layer repeats the first one
            if (c >= nxy) { Uh[c] = Uh[col]; Th[c] = Th[col]; continue; }
            const double x = C[c][0], y = C[c][1];
            Uh[c] = NeoN::Vec3(-std::sin(2.0 * pi * y) * std::pow(std::sin(pi * x), 2.0), std::sin(2.0 * pi * x) * std::pow(std::sin(pi * y), 2.0), 0.0);
            Th[c] = std::exp(-0.5 * (std::pow((x - 0.5) / spread, 2.0) + std::pow((y - 0.75) / spread, 2.0)));
        }
        std::vector<fvcc::VolumeBoundary<NeoN::scalar>> TBCs(size_t(mesh.nBoundaries()), fvcc::VolumeBoundary<NeoN::scalar>("zeroGradient", 0.0));
        std::vector<fvcc::VolumeBoundary<NeoN::Vec3>> UBCs(size_t(mesh.nBoundaries()), fvcc::VolumeBoundary<NeoN::Vec3>("zeroGradient", NeoN::zero<NeoN::Vec3>()));
        fvcc::VolumeField<NeoN::scalar> nfT(exec, "nfT", mesh, TBCs);
        fvcc::VolumeField<NeoN::Vec3> U(exec, "U", mesh, UBCs);
        nfT.internalVector().copyFromHost(Th.data());
        U.internalVector().copyFromHost(Uh.data());
        nfT.correctBoundaryConditions();
        U.correctBoundaryConditions();
        auto nfPhi0 = nf::flux(U); // createFields.H:42-49
        nfPhi0.name = "phi";
        fvcc::SurfaceField<NeoN::scalar> nfPhi(exec, "phi", mesh);

        std::cout << "\nStarting time loop\n" << std::endl;
        NeoN::scalar t = 0.0;
        NeoN::la::SolverStats last {0, 0.0, 0.0, {}};
        for (int step = 0; step < steps; ++step)
        {
            const NeoN::scalar dt = rt.dt;
            auto& nfOldT = fvcc::oldTime(nfT);
            nfOldT.internalVector() = nfT.internalVector(); // scalarAdvection.cpp:57-58
            // :66-67 nfPhi = nfPhi0 * cos(pi (t + dt/2) / endTime)
            NeoN::check(fvk_vec_scaled_copy(int64_t(nfPhi.internalVector().size()), std::cos(pi * (t + 0.5 * dt) / endTime), nfPhi0.internalVector().data(),
                                            nfPhi.internalVector().data(), exec.stream()));
            const auto coNum = fvcc::computeCoNum(nfPhi, dt); // :70
            (void) coNum;
            t += dt;
            std::cout << "Time = " << t << std::endl;
            if (integrator == "backwardEuler")
            {
                dsl::Expression<NeoN::scalar> eqnSys(dsl::imp::ddt(nfT) + dsl::imp::div(nfPhi, nfT)); // :80
                last = dsl::solve(eqnSys, nfT, t - dt, dt, rt.fvSchemesDict, fvSolution);
            }
            else
            {
                dsl::Expression<NeoN::scalar> eqnSys(dsl::imp::ddt(nfT) + dsl::exp::div(nfPhi, nfT)); // test_advection.cpp:150-153
                last = dsl::solve(eqnSys, nfT, t - dt, dt, rt.fvSchemesDict, fvSolution);
            }
        }
        auto out = nfT.internalVector().copyToHost();
        double tmin = 1e300, tmax = -1e300;
        for (auto v : out) { tmin = std::min(tmin, v); tmax = std::max(tmax, v); }
        std::cout.precision(17);
        std::cout << "End: cells " << mesh.nCells() << " T range [" << tmin << ", " << tmax << "] last solve iterations " << last.numIter << std::endl;
        if (const char* dump = std::getenv("SCALARADVECTION_DUMP"))
        {
            FILE* f = std::fopen(dump, "wb");
            if (f) { std::fwrite(out.data(), sizeof(double), out.size(), f); std::fclose(f); }
        }
    }
    catch (const std::exception& e)
    {
        std::cerr << "scalarAdvection: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
