// heatTransfer on the B200 kernels: the reference's loop (examples/heatTransfer/heatTransfer.cpp:55-90) written against
// include/NeoN: ddt(T) - laplacian(kappa, T) = 0 advanced with dsl::solve (backwardEuler + the linear solver of fvSolution.solvers.nfT).
// OpenFOAM's case reading is replaced by the block-mesh generator: unit cube (or N x N x 1 slab), T = 300 on the first patch
// (y-max) and 273 elsewhere / initially, kappa uniform.
// usage: heatTransfer [N=20] [steps=10] [--3d]
#include "FoamAdapter/FoamAdapter.hpp"

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
        int N = 20, steps = 10, pos = 0;
        bool threeD = false;
        for (int i = 1; i < argc; ++i)
        {
            if (std::strcmp(argv[i], "--3d") == 0) threeD = true;
            else if (pos == 0) { N = std::atoi(argv[i]); ++pos; }
            else steps = std::atoi(argv[i]);
        }
        NeoN::Executor exec(0);
        std::vector<NeoN::BlockPatch> patches {{"hot", {3}, false}};
        if (threeD) patches.push_back({"cold", {0, 1, 2, 4, 5}, false});
        else { patches.push_back({"cold", {0, 1, 2}, false}); patches.push_back({"frontAndBack", {4, 5}, true}); }
        auto mesh = NeoN::UnstructuredMesh::createBlockMesh(exec, N, N, threeD ? N : 1, 1.0, 1.0, threeD ? 1.0 : 0.1, patches);
        nf::RunTime rt {exec, mesh};
        rt.dt = 1e-3;
        rt.fvSchemesDict.insert("ddtSchemes", NeoN::Dictionary {{"type", std::string("backwardEuler")}});
        rt.fvSchemesDict.insert("laplacianSchemes", NeoN::Dictionary {{"laplacian(kappa,nfT)", std::string("Gauss linear uncorrected")}});
        NeoN::Dictionary nfTSolver = nf::mapFvSolution(NeoN::Dictionary {{"solver", std::string("PCG")}, {"preconditioner", std::string("DIC")},
                                                                          {"tolerance", 1e-10}, {"relTol", 0.0}});
        std::vector<fvcc::VolumeBoundary<NeoN::scalar>> bcs {{"fixedValue", 300.0}, {"fixedValue", 273.0}};
        fvcc::VolumeField<NeoN::scalar> nfT(exec, "nfT", mesh, bcs);
        NeoN::fill(nfT.internalVector(), 273.0);
        nfT.correctBoundaryConditions();
        auto& nfTOld = fvcc::oldTime(nfT);
        fvcc::SurfaceField<NeoN::scalar> nfKappa(exec, "kappa", mesh);
        NeoN::fill(nfKappa.internalVector(), 0.5);
        NeoN::fill(nfKappa.boundaryData().value(), 0.5);

        std::cout << "\nStarting time loop\n" << std::endl;
        for (int step = 0; step < steps; ++step)
        {
            rt.t += rt.dt;
            std::cout << "Time = " << rt.t << "\n" << std::endl;
            nfTOld.internalVector() = nfT.internalVector();
            dsl::Expression<NeoN::scalar> nfTEqn(dsl::imp::ddt(nfT) - dsl::imp::laplacian(nfKappa, nfT)); // heatTransfer.cpp:67
            auto st = dsl::solve(nfTEqn, nfT, rt.t - rt.dt, rt.dt, rt.fvSchemesDict, nfTSolver);           // :69-76
            nfT.correctBoundaryConditions();
            std::cout << "[NeoN] Solving for nfT: Final residual: " << st.finalResNorm << " No Iterations: " << st.numIter << std::endl;
        }
        auto out = nfT.internalVector().copyToHost();
        double tmin = 1e300, tmax = -1e300;
        for (auto v : out) { tmin = std::min(tmin, v); tmax = std::max(tmax, v); }
        std::cout.precision(17);
        std::cout << "End: cells " << mesh.nCells() << " T range [" << tmin << ", " << tmax << "]" << std::endl;
        if (const char* dump = std::getenv("HEATTRANSFER_DUMP"))
        {
            FILE* f = std::fopen(dump, "wb");
            if (f) { std::fwrite(out.data(), sizeof(double), out.size(), f); std::fclose(f); }
        }
    }
    catch (const std::exception& e)
    {
        std::cerr << "heatTransfer: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
