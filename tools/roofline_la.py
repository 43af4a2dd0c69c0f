#!/usr/bin/env python
"""Roofline harness (GPU box) for the implicit path: fused assembly, CSR SpMV, one Jacobi-CG iteration, and the PISO step
of the 3-D cavity. CUDA events, L2 flushed between timed launches. Usage: python tools/roofline_la.py [--mesh 128 256] [--piso 64 128]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import mesh_counts, peaks  # noqa: E402
from foamadapter_b200 import fvcc, la, ops, piso  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402
from tools.roofline import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs="*", default=[128, 256])
    ap.add_argument("--piso", type=int, nargs="*", default=[64, 128])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "roofline_la.jsonl"))
    args = ap.parse_args()
    peak, kind = peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    rows = []

    def emit(r):
        rows.append(r)
        print(json.dumps(r), flush=True)

    for n in args.mesh:
        gm = UnstructuredMesh(MeshDesc.block(n, n, n, 0.1, 0.1, 0.01))
        nC, nI, nB = mesh_counts(n)
        nnz = nC + 2 * nI
        rng = np.random.default_rng(42)
        T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
        T.internal.copy_(torch.from_numpy(rng.uniform(1, 2, nC)))
        T.correctBoundaryConditions()
        flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
        gamma = torch.ones(nI + nB, dtype=torch.float64, device="cuda")
        old = T.internal - 1.0
        ls = la.LinearSystem(gm, 1, zero=False)
        terms = [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=gamma),
                 dict(kind=ops.TERM_DDT, coeff=1.0, cellField=old, dt=1.0)]
        bytes_asm = 50 * nI + 85 * nC + 52 * nB
        med, best = timeit(lambda: ops.assemble(gm, terms, T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs), args.reps, flush)
        emit({"mesh": n, "kernel": "assemble_ddt_div_lap_s", "ms": med, "best_ms": best, "alg_bytes": bytes_asm, "gbs": bytes_asm / med / 1e6,
              "frac_of_" + kind: bytes_asm / med / 1e6 / peak, "face_ops_per_s": (nI + nB) / (med * 1e-3)})
        sp = la.SparsityPattern.readOrCreate(gm)
        x = torch.from_numpy(rng.uniform(-1, 1, nC)).cuda()
        y = torch.empty_like(x)
        bytes_spmv = nnz * 12 + nC * 20
        med, best = timeit(lambda: la.spmv(sp, ls.values, x, y), args.reps, flush)
        emit({"mesh": n, "kernel": "spmv", "ms": med, "best_ms": best, "alg_bytes": bytes_spmv, "gbs": bytes_spmv / med / 1e6,
              "frac_of_" + kind: bytes_spmv / med / 1e6 / peak})
        med, best = timeit(lambda: la.spmv_structured(gm, ls.values, x, y), args.reps, flush)
        emit({"mesh": n, "kernel": "spmv_structured", "ms": med, "best_ms": best, "alg_bytes": bytes_spmv, "gbs": bytes_spmv / med / 1e6,
              "frac_of_" + kind: bytes_spmv / med / 1e6 / peak})
        # CG iterations on the SPD system -laplacian + ddt (fixed iteration count, no early stop)
        ops.assemble(gm, [dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=gamma), dict(kind=ops.TERM_DDT, coeff=1.0, cellField=old, dt=1.0)],
                     T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        iters = 50
        cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
               "criteria": {"iteration": iters, "relative_residual_norm": 0.0, "absolute_residual_norm": 0.0}}
        solver = la.Solver(cfg, check_every=iters + 1)
        xs = torch.zeros(nC, dtype=torch.float64, device="cuda")
        def solve():
            xs.zero_()
            return solver.solve(ls, xs)
        solve()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = solve()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        ms_it = float(np.median(ts)) / iters
        bytes_it = bytes_spmv + 72 * nC
        emit({"mesh": n, "kernel": "pcg_jacobi_iteration", "ms": ms_it, "iters": st.numIter, "alg_bytes": bytes_it, "gbs": bytes_it / ms_it / 1e6,
              "frac_of_" + kind: bytes_it / ms_it / 1e6 / peak, "note": "wall clock of a 50-iteration solve / 50 (includes setup kernels)"})
        del gm, ls, T, flux, gamma, old, x, y, xs, solver
        torch.cuda.empty_cache()
    for n in args.piso:
        gm = UnstructuredMesh(piso.cavity_desc(n, True))
        app = piso.IcoFoam(gm, nu=0.01, dt=1e-4 * 20 / n, check_every=16)
        for _ in range(2):
            app.step()
        torch.cuda.synchronize()
        ts, its = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            st = app.step()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
            its.append([s.numIter for s in st])
        emit({"mesh": n, "kernel": "piso_step_cavity3d", "ms": float(np.median(ts)), "ms_all": ts, "cg_iters_per_corrector": its,
              "cells": n ** 3})
        del gm, app
        torch.cuda.empty_cache()
    Path(args.out).parent.mkdir(exist_ok=True, parents=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
