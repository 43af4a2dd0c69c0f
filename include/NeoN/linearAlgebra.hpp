// NeoN::la for the B200 build: SparsityPattern, CSRMatrix views, LinearSystem, BoundaryCoefficients, computeResidual and
// the Solver front end (src/NeoN/include/NeoN/linearAlgebra/{sparsityPattern,CSRMatrix,linearSystem,utilities,solver,
// ginkgo}.hpp). The sparsity pattern is built once per mesh by fvk_mesh_create (bit-exact with
// src/NeoN/src/linearAlgebra/sparsityPattern.cpp:21-143); the solver is libfvk's device-resident Jacobi-CG instead of Ginkgo.
#pragma once

#include "NeoN/core.hpp"
#include "NeoN/mesh.hpp"

namespace NeoN::la
{
class SparsityPattern
{
public:
    explicit SparsityPattern(const UnstructuredMesh& mesh) : mesh_(mesh) {}
    // cached in the mesh handle like SparsityPattern::readOrCreate (sparsityPattern.cpp:11-19)
    static SparsityPattern readOrCreate(const UnstructuredMesh& mesh) { return SparsityPattern(mesh); }
    const UnstructuredMesh& mesh() const { return mesh_; }
    View<const localIdx> rowOffs() const { return mesh_.arr<localIdx>(FVK_ROW_OFFS); }
    View<const localIdx> colIdxs() const { return mesh_.arr<localIdx>(FVK_COL_IDXS); }
    View<const uint8_t> ownerOffset() const { return mesh_.arr<uint8_t>(FVK_OWNER_OFFSET); }
    View<const uint8_t> neighbourOffset() const { return mesh_.arr<uint8_t>(FVK_NEIGHBOUR_OFFSET); }
    View<const uint8_t> diagOffset() const { return mesh_.arr<uint8_t>(FVK_DIAG_OFFSET); }
    localIdx rows() const { return mesh_.nCells(); }
    int64_t nnz() const { return mesh_.nnz(); }
private:
    UnstructuredMesh mesh_;
};

// linearSystem.hpp:36-43
template<typename T>
struct BoundaryCoefficients
{
    BoundaryCoefficients(const Executor& exec, size_t nB) : matrixValues(exec, nB, zero<T>()), rhsValues(exec, nB, zero<T>()), matrixIdxs(exec, nB), rhsIdxs(exec, nB) {}
    Vector<T> matrixValues, rhsValues;
    Vector<localIdx> matrixIdxs, rhsIdxs;
};

template<typename T>
struct CSRMatrixView
{
    T* values;
    const localIdx* colIdxs;
    const localIdx* rowOffs;
};

// LinearSystem<ValueType, localIdx> (linearSystem.hpp:53-121): values T[nnz], rhs T[nCells] over the mesh's pattern
template<typename T, typename IndexType = localIdx>
class LinearSystem
{
public:
    LinearSystem(const UnstructuredMesh& mesh, const SparsityPattern& sp, bool zeroFill = true)
        : mesh_(mesh), sp_(sp), values_(mesh.exec(), size_t(sp.nnz())), rhs_(mesh.exec(), size_t(mesh.nCells())),
          bc_(mesh.exec(), size_t(mesh.nBoundaryFaces()))
    {
        if (zeroFill) reset();
        check(fvk_bc_coeff_indices(mesh.handle(), bc_.matrixIdxs.data(), bc_.rhsIdxs.data(), mesh.exec().stream()));
    }
    const Executor& exec() const { return mesh_.exec(); }
    const UnstructuredMesh& mesh() const { return mesh_; }
    const SparsityPattern& sparsityPattern() const { return sp_; }
    Vector<T>& values() { return values_; }
    const Vector<T>& values() const { return values_; }
    Vector<T>& rhs() { return rhs_; }
    const Vector<T>& rhs() const { return rhs_; }
    BoundaryCoefficients<T>& boundaryCoefficients() { return bc_; }
    const BoundaryCoefficients<T>& boundaryCoefficients() const { return bc_; }
    CSRMatrixView<T> view() { return {values_.data(), sp_.colIdxs().ptr, sp_.rowOffs().ptr}; }
    void reset() { values_.fillWith(zero<T>()); rhs_.fillWith(zero<T>()); }
private:
    UnstructuredMesh mesh_;
    SparsityPattern sp_;
    Vector<T> values_, rhs_;
    BoundaryCoefficients<T> bc_;
};

// createEmptyLinearSystem (linearSystem.hpp:140-186)
template<typename T, typename IndexType = localIdx>
LinearSystem<T, IndexType> createEmptyLinearSystem(const UnstructuredMesh& mesh, const SparsityPattern& sp)
{
    return LinearSystem<T, IndexType>(mesh, sp, true);
}

// computeResidual (src/NeoN/src/linearAlgebra/utilities.cpp:11-35): res = A x - b
inline void computeResidual(const LinearSystem<scalar>& ls, const Vector<scalar>& x, Vector<scalar>& res)
{
    const auto& sp = ls.sparsityPattern();
    check(fvk_residual(ls.mesh().nOwnedCells(), sp.rowOffs().ptr, sp.colIdxs().ptr, ls.values().data(), ls.rhs().data(), x.data(),
                       res.data(), ls.exec().stream()));
}
inline void spmv(const LinearSystem<scalar>& ls, const Vector<scalar>& x, Vector<scalar>& y)
{
    const auto& sp = ls.sparsityPattern();
    check(fvk_spmv(ls.mesh().nOwnedCells(), sp.rowOffs().ptr, sp.colIdxs().ptr, ls.values().data(), x.data(), y.data(), ls.exec().stream()));
}

// solver.hpp:14-27
struct SolverStats
{
    int numIter;
    scalar initResNorm;
    scalar finalResNorm;
    std::vector<scalar> residualHistory; // ||r|| at every stopping check (extension: the reference logs only the final value)
    void print(std::string solverName) const
    {
        std::cout << "Solver: " << solverName << " , Initial residual = " << initResNorm << " , Final residual = " << finalResNorm
                  << " , No Iterations = " << numIter << std::endl;
    }
};

// la::SolverFactory (solver.hpp:29-60): linear solvers registered by name; the dictionary's `solver` entry selects one. The
// reference registers exactly one, "Ginkgo" (ginkgo.hpp:95-169); here that name is bound to libfvk's device-resident CG /
// BiCGStab (class Solver below doubles as the registered implementation and as the front end la::Solver of solver.hpp:63-91).
class SolverFactory : public RuntimeSelectionFactory<SolverFactory, Parameters<const Executor&, const Dictionary&>>
{
public:
    static std::string name() { return "SolverFactory"; }
    virtual ~SolverFactory() = default;
};

// la::Solver(exec, dict) (solver.hpp:63-91) for Ginkgo-style dictionaries (ginkgo.hpp:95-108) as mapFvSolution emits them
// (FoamAdapter src/compatibility/fvSolution.cpp:19-159) or as the reference's tests pass them (test/test_advection.cpp:125-131):
// {solver Ginkgo; type solver::Cg | solver::Bicgstab; preconditioner{type preconditioner::Jacobi; max_block_size 1};
// criteria{iteration; relative_residual_norm; absolute_residual_norm}}. Anything else is an error (no silent downgrade).
class Solver
{
public:
    Solver(const Executor& exec, const Dictionary& dict, fvk_comm* comm = nullptr, int checkEvery = 8, bool history = false)
        : exec_(exec), comm_(comm), history_(history)
    {
        const auto name = dict.getOr<std::string>("solver", "Ginkgo");
        SolverFactory::keyExistsOrError(name); // a third-party solver registered under another name constructs itself via SolverFactory::create
        if (name != "Ginkgo") NF_ERROR_EXIT("la::Solver implements the solver registered as Ginkgo; use SolverFactory::create for " + name);
        const auto type = dict.getOr<std::string>("type", "solver::Cg");
        if (type == "solver::Cg") cfg_.solverType = FVK_SOLVER_CG;
        else if (type == "solver::Bicgstab") cfg_.solverType = FVK_SOLVER_BICGSTAB;
        else NF_ERROR_EXIT("solver type " + type + " is not on the hot path (solver::Cg, solver::Bicgstab)");
        cfg_.maxIter = 1000; cfg_.relTol = 0.0; cfg_.absTol = 0.0; cfg_.preconditioner = FVK_PRECOND_NONE; cfg_.checkEvery = checkEvery;
        if (dict.contains("criteria"))
        {
            const auto& c = dict.subDict("criteria");
            cfg_.maxIter = c.getOr<int>("iteration", 1000);
            cfg_.relTol = c.getOr<scalar>("relative_residual_norm", 0.0);
            cfg_.absTol = c.getOr<scalar>("absolute_residual_norm", 0.0);
        }
        if (dict.contains("preconditioner"))
        {
            if (!dict.isDict("preconditioner")) NF_ERROR_EXIT("preconditioner " + dict.get<std::string>("preconditioner") + " not supported");
            const auto& pd = dict.subDict("preconditioner");
            const auto ptype = pd.getOr<std::string>("type", "");
            if (ptype == "preconditioner::Jacobi" && pd.getOr<int>("max_block_size", 1) == 1) cfg_.preconditioner = FVK_PRECOND_JACOBI;
            else if (ptype == "preconditioner::Ic") cfg_.preconditioner = FVK_PRECOND_DIC; // extension: multicolour DIC (fvk.h)
            else NF_ERROR_EXIT("preconditioner " + ptype + " not supported (scalar preconditioner::Jacobi only)");
        }
    }
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;
    ~Solver() { if (h_) fvk_solver_destroy(h_); }

    SolverStats solve(const LinearSystem<scalar, localIdx>& ls, Vector<scalar>& x) const
    {
        const auto& m = ls.mesh();
        if (!h_ || rows_ != m.nOwnedCells() || cols_ != m.nCells())
        {
            if (h_) fvk_solver_destroy(h_);
            h_ = nullptr;
            check(fvk_solver_create(m.nOwnedCells(), m.nCells(), &cfg_, comm_, &h_));
            if (ghostsCurrent_) check(fvk_solver_set_ghosts_current(h_, 1));
            rows_ = m.nOwnedCells(); cols_ = m.nCells();
        }
        // structured SpMV inside the solver when the mesh plan proved a block topology; re-attached on every solve (host-only,
        // cheap): a mesh destroyed and re-created at the same address must not leave stale dimensions behind
        check(fvk_solver_attach_mesh(h_, m.handle()));
        const auto& sp = ls.sparsityPattern();
        fvk_solver_stats st {};
        std::vector<scalar> hist(history_ ? size_t(cfg_.solverType == FVK_SOLVER_BICGSTAB ? 2 : 1) * size_t(cfg_.maxIter) + 2 : 0);
        check(fvk_solver_solve(h_, sp.rowOffs().ptr, sp.colIdxs().ptr, ls.values().data(), ls.rhs().data(), x.data(), &st,
                               history_ ? hist.data() : nullptr, int32_t(hist.size()), exec_.stream()));
        hist.resize(history_ ? size_t(st.nHistory) : 0);
        return {st.numIter, st.initResNorm, st.finalResNorm, hist};
    }
    // Vec3 system with identical components (momentum equation): one scalar solve per component over the component matrix
    std::array<SolverStats, 3> solve(const LinearSystem<Vec3, localIdx>& ls, Vector<Vec3>& x) const
    {
        const auto& m = ls.mesh();
        if (!h_ || rows_ != m.nOwnedCells() || cols_ != m.nCells())
        {
            if (h_) fvk_solver_destroy(h_);
            h_ = nullptr;
            check(fvk_solver_create(m.nOwnedCells(), m.nCells(), &cfg_, comm_, &h_));
            if (ghostsCurrent_) check(fvk_solver_set_ghosts_current(h_, 1));
            rows_ = m.nOwnedCells(); cols_ = m.nCells();
        }
        check(fvk_solver_attach_mesh(h_, m.handle()));
        const auto& sp = ls.sparsityPattern();
        fvk_solver_stats st[3] {};
        check(fvk_solver_solve_vec3(h_, int64_t(ls.values().size()), sp.rowOffs().ptr, sp.colIdxs().ptr, ls.values().raw(), ls.rhs().raw(), x.raw(), st,
                                    exec_.stream()));
        return {SolverStats {st[0].numIter, st[0].initResNorm, st[0].finalResNorm, {}}, SolverStats {st[1].numIter, st[1].initResNorm, st[1].finalResNorm, {}},
                SolverStats {st[2].numIter, st[2].initResNorm, st[2].finalResNorm, {}}};
    }
    // ---- distributed solves: ghost entries of the unknown (fvk.h "Ghost entries around a distributed solve") -----------------
    // the initial guesses passed from now on already have current ghost entries (skip the start-up exchange)
    void setGhostsCurrent(bool on) const
    {
        ghostsCurrent_ = on;
        if (h_) check(fvk_solver_set_ghosts_current(h_, on ? 1 : 0));
    }
    // does the solution come back with current ghost entries (no exchange of x needed after solve)?
    bool keepsGhosts() const
    {
        if (!h_) return comm_ == nullptr || (cfg_.solverType == FVK_SOLVER_CG && fvk_comm_p2p_enabled(comm_) == 1);
        int32_t yes = 0;
        check(fvk_solver_keeps_ghosts(h_, &yes));
        return yes != 0;
    }
    // iteration counts of the solves that ran inside replayed CUDA graphs since the last call (device-side log)
    std::vector<int32_t> capturedLog() const
    {
        std::vector<int32_t> out(8192);
        int32_t n = 0;
        if (h_) check(fvk_solver_captured_log(h_, out.data(), int32_t(out.size()), &n));
        out.resize(size_t(n));
        return out;
    }
private:
    Executor exec_;
    fvk_comm* comm_;
    bool history_;
    mutable bool ghostsCurrent_ = false;
    fvk_solver_config cfg_ {};
    mutable fvk_solver* h_ = nullptr;
    mutable localIdx rows_ = 0, cols_ = 0;
};

// the name the reference's solver dictionaries carry
class GinkgoSolver : public SolverFactory::Register<GinkgoSolver>
{
public:
    GinkgoSolver(const Executor& exec, const Dictionary& dict) : solver_(exec, dict) {}
    static std::string name() { return "Ginkgo"; }
    static std::string doc() { return "libfvk device-resident solver::Cg / solver::Bicgstab with scalar Jacobi"; }
    static std::string schema() { return "none"; }
    SolverStats solve(const LinearSystem<scalar, localIdx>& sys, Vector<scalar>& x) const { return solver_.solve(sys, x); }
private:
    Solver solver_;
};
NF_REGISTER((SolverFactory), (GinkgoSolver));

} // namespace NeoN::la
