// FoamAdapter glue for the B200 build: RunTime, mapFvSolution, PDESolver<T> and the PISO pressure-velocity coupling
// helpers -- the names and call sequence of the reference (include/FoamAdapter/datastructures/{runTime,expression}.hpp,
// src/compatibility/fvSolution.cpp, src/algorithms/pressureVelocityCoupling.cpp). OpenFOAM itself (mesh reading,
// dictionaries, Time) is outside the hot path: the mesh comes from NeoN::UnstructuredMesh::createBlockMesh or any
// fvk_mesh_desc, the dictionaries are NeoN::Dictionary objects.
#pragma once

#include "NeoN/NeoN.hpp"

namespace FoamAdapter
{
namespace nnfvcc = NeoN::finiteVolume::cellCentred;
namespace dsl = NeoN::dsl;
namespace la = NeoN::la;
using NeoN::scalar;
using NeoN::Vec3;

// datastructures/runTime.hpp
struct RunTime
{
    NeoN::Executor exec;
    NeoN::UnstructuredMesh nfMesh;
    NeoN::scalar t = 0.0, dt = 1.0;
    NeoN::Dictionary fvSchemesDict, fvSolutionDict;
    fvk_comm* comm = nullptr; // multi-GPU runs: halo exchange + all-reduced dots (nullptr = single GPU)
};

// FoamAdapter::mapFvSolution (src/compatibility/fvSolution.cpp:142-157), step for step: updateSolver (:19-46),
// updatePreconditioner (:48-105; a MISSING preconditioner becomes DIC -> scalar Jacobi, `smoother` is dropped, DILU -> Ilu /
// ParIlu, which la::Solver then rejects), updateCriteria (:107-139; iteration 1000 unless maxIter, the norms only when relTol /
// tolerance are present). A dictionary with `configFile` is returned unchanged.
inline NeoN::Dictionary mapFvSolution(const NeoN::Dictionary& in)
{
    if (in.contains("configFile")) return in;
    NeoN::Dictionary out = in;
    static const std::map<std::string, std::string> solverMap = {{"PCG", "solver::Cg"}, {"PBiCG", "solver::Bicg"}, {"PBiCGStab", "solver::Bicgstab"},
                                                                 {"smoothSolver", "solver::Bicgstab"}, {"GAMG", "solver::Multigrid"}};
    if (out.contains("solver"))
    {
        auto it = solverMap.find(out.get<std::string>("solver"));
        if (it != solverMap.end())
        {
            out.insert("solver", std::string("Ginkgo"));
            out.insert("type", it->second);
        }
    }
    auto jacobi = [] {
        NeoN::Dictionary p;
        p.insert("type", std::string("preconditioner::Jacobi"));
        p.insert("max_block_size", 1);
        return p;
    };
    if (!out.contains("preconditioner")) out.insert("preconditioner", jacobi());
    if (out.contains("smoother")) out.remove("smoother");
    if (out.isDict("preconditioner"))
    {
        if (out.subDict("preconditioner").isDict("type"))
            NF_ERROR_EXIT("GAMG is not supported in FoamAdapter, please use a different preconditioner.");
    }
    else
    {
        const auto pre = out.get<std::string>("preconditioner");
        if (pre == "DIC") out.insert("preconditioner", jacobi());
        else if (pre == "DILU")
        {
            NeoN::Dictionary p, f;
            f.insert("type", std::string("factorization::ParIlu"));
            p.insert("type", std::string("preconditioner::Ilu"));
            p.insert("reverse_apply", false);
            p.insert("factorization", f);
            out.insert("preconditioner", p);
        }
    }
    if (!out.contains("criteria")) out.insert("criteria", NeoN::Dictionary());
    auto& crit = out.subDict("criteria");
    crit.insert("iteration", 1000);
    auto num = [&](const std::string& key) { // scalar entries may have been written as integers (relTol 0)
        try { return out.get<NeoN::scalar>(key); } catch (...) { return NeoN::scalar(out.get<int>(key)); }
    };
    if (out.contains("relTol")) { crit.insert("relative_residual_norm", num("relTol")); out.remove("relTol"); }
    if (out.contains("maxIter")) { crit.insert("iteration", out.get<int>("maxIter")); out.remove("maxIter"); }
    if (out.contains("tolerance")) { crit.insert("absolute_residual_norm", num("tolerance")); out.remove("tolerance"); }
    return out;
}

// datastructures/expression.hpp:23-179
template<typename ValueType, typename IndexType = NeoN::localIdx>
class PDESolver
{
    using VolumeField = nnfvcc::VolumeField<ValueType>;

public:
    PDESolver(dsl::Expression<ValueType> expr, VolumeField& psi, const RunTime& runTime)
        : psi_(psi), expr_(std::move(expr)), runTime_(runTime), sparsityPattern_(la::SparsityPattern::readOrCreate(psi.mesh())),
          ls_(psi.mesh(), sparsityPattern_, false) // the fused assembly writes every entry: no zero-fill
    {
        expr_.read(runTime_.fvSchemesDict);
    }
    VolumeField& getField() { return psi_; }
    const VolumeField& getField() const { return psi_; }
    const la::SparsityPattern& sparsityPattern() const { return sparsityPattern_; }
    la::LinearSystem<ValueType, IndexType>& linearSystem() { return ls_; }
    const la::LinearSystem<ValueType, IndexType>& linearSystem() const { return ls_; }
    const NeoN::Executor& exec() const { return ls_.exec(); }

    la::LinearSystem<ValueType, IndexType>& assemble()
    {
        expr_.assemble(runTime_.t, runTime_.dt, sparsityPattern_, ls_, psi_);
        return ls_;
    }
    void setReference(NeoN::localIdx pRefCell, scalar pRefValue) { needReference_ = true; pRefCell_ = pRefCell; pRefValue_ = pRefValue; }

    // scalar systems: SolverStats; Vec3 systems (momentumPredictor, neoIcoFoam.cpp:100-103): one SolverStats per component
    auto solve()
    {
        const auto& solverDict = runTime_.fvSolutionDict.subDict("solvers").subDict(psi_.name);
        if (!solver_) solver_ = std::make_shared<la::Solver>(exec(), solverDict, runTime_.comm);
        auto post = [&](const la::SparsityPattern&, la::LinearSystem<ValueType, IndexType>& ls)
        { // SetReference (expression.hpp:86-112), scalar systems only (:120-132)
            if constexpr (std::is_same_v<ValueType, scalar>)
            {
                if (needReference_) NeoN::check(fvk_set_reference(psi_.mesh().handle(), pRefCell_, pRefValue_, ls.values().data(), ls.rhs().data(), exec().stream()));
            }
            else (void) ls;
        };
        auto stats = dsl::detail::iterativeSolveImpl(expr_, sparsityPattern_, ls_, psi_, runTime_.t, runTime_.dt, *solver_, post);
        if constexpr (std::is_same_v<ValueType, scalar>)
            std::cout << "[NeoN] Solving for " << psi_.name << ":" << " Initial residual: " << stats.initResNorm
                      << " Final residual: " << stats.finalResNorm << " No Iterations: " << stats.numIter << std::endl;
        else
            for (int c = 0; c < 3; ++c)
                std::cout << "[NeoN] Solving for " << psi_.name << "[" << c << "]:" << " Initial residual: " << stats[c].initResNorm
                          << " Final residual: " << stats[c].finalResNorm << " No Iterations: " << stats[c].numIter << std::endl;
        return stats;
    }
    // expression.hpp:163-167
    auto solve(dsl::SpatialOperator<ValueType>&& rhs)
    {
        expr_.addOperator(-1.0 * rhs);
        return solve();
    }
    void useSolver(std::shared_ptr<la::Solver> s) { solver_ = std::move(s); }

private:
    VolumeField& psi_;
    dsl::Expression<ValueType> expr_;
    const RunTime& runTime_;
    la::SparsityPattern sparsityPattern_;
    la::LinearSystem<ValueType, IndexType> ls_;
    bool needReference_ = false;
    NeoN::localIdx pRefCell_ = 0;
    scalar pRefValue_ = 0.0;
    std::shared_ptr<la::Solver> solver_;
};

// diag(ls, sparsityPattern) (expression.hpp:181-199)
template<typename T>
NeoN::Vector<T> diag(const la::LinearSystem<T>& ls, const la::SparsityPattern&)
{
    NeoN::Vector<T> d(ls.exec(), size_t(ls.mesh().nCells()), NeoN::zero<T>());
    NeoN::check(fvk_diag(ls.mesh().handle(), NeoN::nComponents<T>(), ls.values().raw(), d.raw(), ls.exec().stream()));
    return d;
}

// ---- src/algorithms/pressureVelocityCoupling.cpp ------------------------------------------------------------------
inline void constrainHbyA(const nnfvcc::VolumeField<Vec3>& u, const nnfvcc::VolumeField<scalar>&, nnfvcc::VolumeField<Vec3>& hByA)
{ // :14-36
    std::vector<int32_t> mask;
    bool any = false;
    for (const auto& bc : u.boundaryConditions()) { mask.push_back(bc.assignable() ? 0 : 1); any = any || !bc.assignable(); }
    if (any) NeoN::check(fvk_copy_patches(u.mesh().handle(), 3, mask.data(), u.boundaryData().value().raw(), hByA.boundaryData().value().raw(), u.exec().stream()));
}

inline nnfvcc::VolumeField<scalar> computeRAU(const PDESolver<Vec3>& expr)
{ // :38-63
    const auto& mesh = expr.getField().mesh();
    nnfvcc::VolumeField<scalar> rAU(expr.exec(), "rAU", mesh, nnfvcc::createExtrapolatedBCs<scalar>(mesh));
    NeoN::check(fvk_rAU_HbyA(mesh.handle(), expr.linearSystem().values().raw(), nullptr, nullptr, rAU.internalVector().data(), nullptr, expr.exec().stream()));
    return rAU;
}

inline std::tuple<nnfvcc::VolumeField<scalar>, nnfvcc::VolumeField<Vec3>> computeRAUandHByA(const PDESolver<Vec3>& expr)
{ // :65-128, one fused kernel
    const auto& u = expr.getField();
    const auto& mesh = u.mesh();
    nnfvcc::VolumeField<scalar> rAU(expr.exec(), "rAU", mesh, nnfvcc::createExtrapolatedBCs<scalar>(mesh));
    nnfvcc::VolumeField<Vec3> hByA(expr.exec(), "HbyA", mesh, nnfvcc::createExtrapolatedBCs<Vec3>(mesh));
    const auto& ls = expr.linearSystem();
    NeoN::check(fvk_rAU_HbyA(mesh.handle(), ls.values().raw(), ls.rhs().raw(), u.internalVector().raw(), rAU.internalVector().data(),
                             hByA.internalVector().raw(), expr.exec().stream()));
    hByA.correctBoundaryConditions();
    rAU.correctBoundaryConditions();
    return {std::move(rAU), std::move(hByA)};
}

inline nnfvcc::SurfaceField<scalar> flux(const nnfvcc::VolumeField<Vec3>& volField)
{ // :215-267
    nnfvcc::SurfaceField<scalar> faceFlux(volField.exec(), "out", volField.mesh());
    NeoN::check(fvk_flux(volField.mesh().handle(), volField.internalVector().raw(), volField.boundaryData().value().raw(),
                         faceFlux.internalVector().data(), faceFlux.boundaryData().value().data(), volField.exec().stream()));
    return faceFlux;
}

inline void updateFaceVelocity(const nnfvcc::SurfaceField<scalar>& predictedPhi, const PDESolver<scalar>& expr, nnfvcc::SurfaceField<scalar>& phi)
{ // :131-197
    const auto& ls = expr.linearSystem();
    const auto& bc = ls.boundaryCoefficients();
    NeoN::check(fvk_update_face_velocity(phi.mesh().handle(), ls.values().data(), bc.matrixValues.data(), bc.rhsValues.data(),
                                         expr.getField().internalVector().data(), predictedPhi.internalVector().data(),
                                         predictedPhi.boundaryData().value().data(), phi.internalVector().data(),
                                         phi.boundaryData().value().data(), phi.exec().stream()));
}

inline void updateVelocity(const nnfvcc::VolumeField<Vec3>& hByA, const nnfvcc::VolumeField<scalar>& rAU, const nnfvcc::VolumeField<scalar>& p,
                           nnfvcc::VolumeField<Vec3>& u)
{ // :199-213
    auto gradP = nnfvcc::GaussGreenGrad(p.exec(), p.mesh()).grad(p);
    NeoN::check(fvk_update_velocity(u.mesh().handle(), hByA.internalVector().raw(), rAU.internalVector().data(), gradP.internalVector().raw(),
                                    u.internalVector().raw(), u.exec().stream()));
}

} // namespace FoamAdapter
