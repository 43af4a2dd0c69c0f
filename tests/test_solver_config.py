"""CPU: mapFvSolution against the reference's mapping table (src/compatibility/fvSolution.cpp:19-159) and the oracle's
BiCGStab restatement (Ginkgo 1.10 solver::Bicgstab) against known answers / its defining identities."""
import numpy as np
import pytest

from foamadapter_b200.la import mapFvSolution
from oracle import cpu as ocpu

JACOBI = {"type": "preconditioner::Jacobi", "max_block_size": 1}


def test_map_cavity_pressure_entry():
    # tutorials/cavity/system/fvSolution:19-25 -> SURVEY A.5
    out = mapFvSolution({"solver": "PCG", "preconditioner": "DIC", "tolerance": 1e-6, "relTol": 0})
    assert out == {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": JACOBI,
                   "criteria": {"iteration": 1000, "relative_residual_norm": 0.0, "absolute_residual_norm": 1e-6}}


@pytest.mark.parametrize("name,typ", [("PCG", "solver::Cg"), ("PBiCG", "solver::Bicg"), ("PBiCGStab", "solver::Bicgstab"),
                                      ("smoothSolver", "solver::Bicgstab"), ("GAMG", "solver::Multigrid")])
def test_map_solver_table(name, typ):
    out = mapFvSolution({"solver": name})  # fvSolution.cpp:22-28
    assert out["solver"] == "Ginkgo" and out["type"] == typ


def test_map_defaults_follow_the_reference():
    out = mapFvSolution({"solver": "PCG"})
    assert out["preconditioner"] == JACOBI                    # :65-68 missing preconditioner -> DIC -> Jacobi
    assert out["criteria"] == {"iteration": 1000}             # :117-138 no tolerance -> no absolute_residual_norm
    out = mapFvSolution({"solver": "smoothSolver", "smoother": "symGaussSeidel", "maxIter": 50, "relTol": 0.1})
    assert "smoother" not in out and "maxIter" not in out and "relTol" not in out
    assert out["criteria"] == {"iteration": 50, "relative_residual_norm": 0.1}
    out = mapFvSolution({"solver": "PBiCGStab", "preconditioner": "DILU"})
    assert out["preconditioner"] == {"type": "preconditioner::Ilu", "reverse_apply": False, "factorization": {"type": "factorization::ParIlu"}}
    out = mapFvSolution({"solver": "diagonal"})                # tutorials/scalarAdvection/system/fvSolution: not in the map
    assert out["solver"] == "diagonal" and "type" not in out
    cfg = {"configFile": "gko.json", "solver": "PCG"}
    assert mapFvSolution(cfg) is cfg                          # :146


def test_map_rejects_dictionary_typed_preconditioner():
    with pytest.raises(RuntimeError):
        mapFvSolution({"solver": "PCG", "preconditioner": {"type": {"x": 1}}})


def _tridiag(n, lo, di, up):
    ro, ci, va = [0], [], []
    for i in range(n):
        if i > 0: ci.append(i - 1); va.append(lo)
        ci.append(i); va.append(di)
        if i < n - 1: ci.append(i + 1); va.append(up)
        ro.append(len(ci))
    return np.array(ro, np.int32), np.array(ci, np.int32), np.array(va)


@pytest.mark.parametrize("jacobi", [True, False])
def test_oracle_bicgstab_known_answer_3x3(jacobi):
    # the system of src/NeoN/test/linearAlgebra/ginkgo.cpp:95-124
    ro, ci, va = _tridiag(3, -0.1, 1.0, -0.1)
    x, st, h = ocpu.bicgstab(ro, ci, va, np.array([1.0, 2.0, 3.0]), np.zeros(3), jacobi=jacobi, max_iter=50, rel_tol=1e-14, max_hist=200)
    assert np.allclose(x, [1.24489796, 2.44897959, 3.24489796], atol=1e-8)
    assert abs(st["initResNorm"] - 3.741657386) < 1e-8 and st["numIter"] <= 3 and st["finalResNorm"] < 1e-13
    assert len(h) in (2 * st["numIter"] + 1, 2 * st["numIter"] + 2)  # two checks per full iteration


def test_oracle_bicgstab_nonsymmetric_and_iteration_cap():
    ro, ci, va = _tridiag(200, -1.3, 2.5, -0.4)
    rng = np.random.default_rng(5)
    b = rng.uniform(-1, 1, 200)
    x, st, h = ocpu.bicgstab(ro, ci, va, b, np.zeros(200), jacobi=True, max_iter=500, rel_tol=1e-12, max_hist=2000)
    A = np.zeros((200, 200))
    for i in range(200):
        for k in range(ro[i], ro[i + 1]):
            A[i, ci[k]] = va[k]
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b)
    assert h[-1] == st["finalResNorm"] and h[0] == pytest.approx(np.linalg.norm(b))
    x2, st2, _ = ocpu.bicgstab(ro, ci, va, b, np.zeros(200), jacobi=True, max_iter=3, rel_tol=0.0)
    assert st2["numIter"] == 3   # criteria.iteration stops at the top of iteration 3
    # already converged start: zero iterations, x untouched
    x3, st3, _ = ocpu.bicgstab(ro, ci, va, b, x, jacobi=True, max_iter=10, rel_tol=1e-9)
    assert st3["numIter"] == 0 and np.array_equal(x3, x)
