#!/usr/bin/env python
"""Multi-GPU parity check over both transports, NCCL and peer-memory windows (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Checks, against the single-domain CPU oracle on the same global mesh: halo exchange through fvk_comm, explicit
operators (bit-exact on owned cells), distributed Jacobi-CG (iteration count +-1, solution), and two neoIcoFoam steps."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]  # this script lives in tests/: it is the multi-GPU parity test (oracle = checker)
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import advection as adv, la, ops, piso  # noqa: E402
from foamadapter_b200.decomp import Comm, Decomposition  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402
from oracle.cpu import Mesh as OMesh, cg as oracle_cg  # noqa: E402
from oracle.piso import IcoFoamOracle  # noqa: E402
from oracle.advection import ScalarAdvectionOracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for p2p in (False, True):
        check(rank, world, p2p)
    if "--perf" in sys.argv:
        perf(rank, world)
    dist.destroy_process_group()


def perf(rank, world):
    """Latency of one halo exchange and time of one distributed Jacobi-CG iteration, NCCL vs peer-memory windows."""
    import json
    from foamadapter_b200.decomp import default_split
    n = 128
    px, py, pz = default_split(world)
    G = MeshDesc.block(n * px, n * py, n * pz, 0.1 * px, 0.1 * py, 0.1 * pz)
    d = Decomposition(G, world, rank, n=(px, py, pz))
    lm = UnstructuredMesh(d.desc)
    out = {"ranks": world, "cells_per_rank": d.nOwned, "ghosts": d.nGhost}
    for p2p in (False, True):
        comm = Comm.from_torch()
        comm.set_halo(d, p2p=p2p)
        f = torch.rand(lm.nCells, dtype=torch.float64, device="cuda")
        for _ in range(20):
            comm.halo_exchange(f)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(200):
            comm.halo_exchange(f)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 200 * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        key = "p2p" if p2p else "nccl"
        out[f"halo_us_{key}"] = float(t.item())
        ls = la.LinearSystem(lm, 1, zero=False)
        nB = lm.nBoundaryFaces

        class BD:
            value = torch.zeros(nB, dtype=torch.float64, device="cuda"); refValue = value; refGrad = value
            valueFraction = torch.ones(nB, dtype=torch.float64, device="cuda")
        ops.assemble(lm, [dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=torch.ones(lm.nFaces, dtype=torch.float64, device="cuda"))],
                     BD, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        ls.rhs[: d.nOwned] = torch.rand(d.nOwned, dtype=torch.float64, device="cuda") - 0.5
        iters = 100
        cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
               "criteria": {"iteration": iters, "relative_residual_norm": 0.0, "absolute_residual_norm": 0.0}}
        solver = la.Solver(cfg, comm=comm, check_every=iters + 1)
        x = torch.zeros(lm.nCells, dtype=torch.float64, device="cuda")
        solver.solve(ls, x)
        ts = []
        for _ in range(3):
            x.zero_(); dist.barrier(); torch.cuda.synchronize()
            e0.record(); st = solver.solve(ls, x); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / iters * 1e3)
        t = torch.tensor([float(np.median(ts))], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"cg_iter_us_{key}"] = float(t.item())
        out[f"cg_final_res_{key}"] = st.finalResNorm
        if p2p:
            import ctypes as C
            from foamadapter_b200._capi import lib
            dbg = (C.c_uint64 * 8)()
            lib().fvk_comm_p2p_debug(comm.handle, dbg)
            d_ = list(dbg)
            out["p2p_phase_us"] = {"flag_raise": d_[0] / max(d_[3], 1) / 1e3, "allreduce_rz_rr": d_[1] / max(d_[3], 1) / 1e3,
                                   "halo_wait": d_[2] / max(d_[3], 1) / 1e3, "allreduce_pq": d_[4] / max(d_[5], 1) / 1e3}
        del solver
        comm.close()
    if rank == 0:
        print("MGPU PERF " + json.dumps(out), flush=True)


def check(rank, world, p2p):
    comm = Comm.from_torch()
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    g = MeshDesc.block(24, 20, 16, 1.2, 1.0, 0.8)
    om = OMesh.from_desc(g)
    d = Decomposition(g, world, rank)
    lm = UnstructuredMesh(d.desc)
    comm.set_halo(d, p2p=p2p)
    assert comm.p2p == p2p
    rng = np.random.default_rng(7)
    phi, phib, flux = rng.uniform(1, 2, om.nC), rng.uniform(1, 2, om.nB), rng.uniform(-1, 1, om.nF)
    # 1. halo exchange
    f = torch.zeros(lm.nCells, dtype=torch.float64, device="cuda")
    f[: d.nOwned] = dev(phi[d.cellGlobal[: d.nOwned]])
    comm.halo_exchange(f)
    torch.cuda.synchronize()
    assert np.array_equal(f.cpu().numpy(), phi[d.cellGlobal]), "scalar halo"
    U = rng.uniform(-1, 1, (om.nC, 3))
    fU = torch.zeros((lm.nCells, 3), dtype=torch.float64, device="cuda")
    fU[: d.nOwned] = dev(U[d.cellGlobal[: d.nOwned]])
    comm.halo_exchange(fU)
    torch.cuda.synchronize()
    assert np.array_equal(fU.cpu().numpy(), U[d.cellGlobal]), "Vec3 halo"
    # 2. explicit operators: bit-exact on owned cells
    gid = d.cellGlobal[: d.nOwned]
    o = torch.zeros(lm.nCells, dtype=torch.float64, device="cuda")
    lflux, lphib = dev(d.scatter_faces(flux)), dev(d.scatter_boundary(phib, om.nI))
    assert np.array_equal(ops.div(lm, lflux, f, lphib, o).cpu().numpy()[: d.nOwned], om.div(flux, phi, phib, 0)[gid]), "div"
    assert np.array_equal(ops.laplacian(lm, f, lphib, o).cpu().numpy()[: d.nOwned], om.laplacian(phi, phib)[gid]), "laplacian"
    o3 = torch.zeros((lm.nCells, 3), dtype=torch.float64, device="cuda")
    assert np.array_equal(ops.grad(lm, f, lphib, o3).cpu().numpy()[: d.nOwned], om.grad(phi, phib)[gid]), "grad"
    # 3. distributed CG vs the oracle on the global system
    bd = dict(value=np.zeros(om.nB), refValue=np.zeros(om.nB), valueFraction=np.ones(om.nB), refGrad=np.zeros(om.nB))
    gls = om.empty_system(False)
    om.laplacian_imp(gls, np.ones(om.nF), bd, -1.0, None)
    b = rng.uniform(-1, 1, om.nC)
    tol = 1e-9 * np.linalg.norm(b)
    xo, so, ho = oracle_cg(om.rowOffs, om.colIdxs, gls["values"], b, np.zeros(om.nC), jacobi=True, max_iter=500, rel_tol=0.0, abs_tol=tol, max_hist=600)

    class BD:
        def __init__(s, bb): s.value, s.refValue, s.valueFraction, s.refGrad = (dev(bb[k]) for k in ("value", "refValue", "valueFraction", "refGrad"))
    ls = la.LinearSystem(lm, 1, zero=False)
    lbd = {k: d.scatter_boundary(v, om.nI) for k, v in bd.items()}
    ops.assemble(lm, [dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=torch.ones(lm.nFaces, dtype=torch.float64, device="cuda"))],
                 BD(lbd), ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
    ls.rhs[: d.nOwned] = dev(b[gid])
    x = torch.zeros(lm.nCells, dtype=torch.float64, device="cuda")
    cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
           "criteria": {"iteration": 500, "relative_residual_norm": 0.0, "absolute_residual_norm": tol}}
    cg = la.Solver(cfg, comm=comm, check_every=4, history=True)
    st = cg.solve(ls, x)
    assert abs(st.numIter - so["numIter"]) <= 1, (st.numIter, so["numIter"])
    assert cg.keeps_ghosts == p2p
    if p2p:  # the solution's ghost entries followed the owners' updates bit for bit: an exchange changes nothing
        x2 = x.clone()
        comm.halo_exchange(x2)
        torch.cuda.synchronize()
        assert torch.equal(x2, x), "CG keeps the ghost entries of x current"
        cg.set_ghosts_current(True)  # and a second solve from that guess needs no start-up exchange
        st2 = cg.solve(ls, x2)
        assert st2.numIter <= 1, st2.numIter
        cg.set_ghosts_current(False)
    assert abs(st.initResNorm - so["initResNorm"]) <= 1e-12 * so["initResNorm"]
    n = min(len(st.history), len(ho))
    sig = ho[:n] > 1e-8 * ho[0]
    assert np.allclose(st.history[:n][sig], ho[:n][sig], rtol=1e-7), "CG residual history"
    assert np.abs(x.cpu().numpy()[: d.nOwned] - xo[gid]).max() <= 1e-7 * np.abs(xo).max(), "CG solution"
    # 4. two neoIcoFoam steps on a decomposed 3-D cavity vs the single-domain oracle
    gc = piso.cavity_desc(12, True)
    oc = OMesh.from_desc(gc)
    dc = Decomposition(gc, world, rank)
    lc = UnstructuredMesh(dc.desc)
    comm.set_halo(dc)
    app = piso.IcoFoam(lc, nu=0.01, dt=5e-4, comm=comm, check_every=4)
    ref = IcoFoamOracle(oc, nu=0.01, dt=5e-4)
    for _ in range(2):
        sts = app.step()
        outs = ref.step()
        for s, (so2, _) in zip(sts, outs):
            assert abs(s.numIter - so2["numIter"]) <= 1, (s.numIter, so2["numIter"])
    gidc = dc.cellGlobal[: dc.nOwned]
    Ul, pl = app.U.internal.cpu().numpy()[: dc.nOwned], app.p.internal.cpu().numpy()[: dc.nOwned]
    assert np.abs(Ul - ref.U[gidc]).max() <= 1e-8 * np.abs(ref.U).max(), "PISO U"
    assert np.abs(pl - ref.p[gidc]).max() <= 1e-7 * np.abs(ref.p).max(), "PISO p"
    # 5. scalarAdvection (BASELINE configs[3]) on a decomposed 3-D block: forward Euler bit-identical to the single-domain
    #    oracle, backward Euler (distributed BiCGStab) within the reference's 1e-8
    ga = adv.advection_desc(16, True)
    oa = OMesh.from_desc(ga)
    da = Decomposition(ga, world, rank)
    la_ = UnstructuredMesh(da.desc)
    comm.set_halo(da)
    Ug, Tg = adv.init_fields_columns(oa.C.reshape(-1, 3), 16 * 16)
    gida = da.cellGlobal[: da.nOwned]
    for ddt, steps in (("forwardEuler", 6), ("backwardEuler", 3)):
        sch = {"ddtSchemes": {"type": ddt}, "divSchemes": {"div(phi,nfT)": "Gauss upwind"}}
        appA = adv.ScalarAdvection(la_, 2e-3, 0.1, fvSchemes=sch, comm=comm, U=Ug[da.cellGlobal], T=Tg[da.cellGlobal], check_every=4)
        refA = ScalarAdvectionOracle(oa, 2e-3, 0.1, scheme=1, ddt=ddt, U=Ug, T=Tg)
        for _ in range(steps):
            appA.step(); refA.step()
        Tl = appA.T.internal.cpu().numpy()[: da.nOwned]
        if ddt == "forwardEuler":
            assert np.array_equal(Tl, refA.T[gida]), "advection forward Euler"
        else:
            assert np.abs(Tl - refA.T[gida]).max() <= 1e-8 * np.abs(refA.T).max(), "advection backward Euler"
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print(f"MGPU CHECK OK on {world} ranks ({'peer-memory windows' if p2p else 'NCCL'}): halo, explicit ops bit-exact, CG iters {st.numIter} (oracle {so['numIter']}), PISO 2 steps, scalarAdvection fwd (bit-exact) / bwd Euler", flush=True)
    comm.close()


if __name__ == "__main__":
    main()
