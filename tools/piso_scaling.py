#!/usr/bin/env python
"""PISO step time of the 3-D lid-driven cavity (BASELINE configs[4]) on 1..8 GPUs, strong scaling: the SAME n^3 mesh is
decomposed over the ranks. Run plain (1 GPU) or under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/piso_scaling.py --size 256
Prints one JSON line on rank 0: ms per step (max over ranks, CUDA events), CG iterations, transport."""
import os as _os
if "LOCAL_WORLD_SIZE" in _os.environ and _os.environ.get("OMP_NUM_THREADS") == "1":
    # torchrun pins every rank to ONE OpenMP thread unless told otherwise; the once-per-mesh host setup (block generator, stencil,
    # plans, decomposition) is multi-threaded: give each rank its share of the host cores
    _os.environ["OMP_NUM_THREADS"] = str(max(1, (_os.cpu_count() or 1) // int(_os.environ["LOCAL_WORLD_SIZE"])))


import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import piso  # noqa: E402
from foamadapter_b200.decomp import Comm, Decomposition, default_split  # noqa: E402
from foamadapter_b200.mesh import UnstructuredMesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--check-every", type=int, default=16)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--no-comm", action="store_true", help="diagnostic: every rank steps its sub-domain WITHOUT a communicator (stale ghosts, "
                    "pressure solve capped at 3 iterations): the kernel-only time of a rank's share, to separate it from the communication cost")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    comm = None
    t0 = time.perf_counter()
    g = piso.cavity_desc(args.size, True)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dec = Decomposition(g, world, rank, n=default_split(world))
        mesh = UnstructuredMesh(dec.desc)
        if not args.no_comm:
            comm = Comm.from_torch()
            comm.set_halo(dec, p2p=args.transport == "p2p")
        del g
    else:
        mesh = UnstructuredMesh(g)
    setup_s = time.perf_counter() - t0
    sol = None
    if args.no_comm and world > 1:
        import copy
        sol = copy.deepcopy(piso.CAVITY_FVSOLUTION)
        sol["solvers"]["p"] = {"solver": "PCG", "preconditioner": "DIC", "tolerance": 0.0, "relTol": 0.0, "maxIter": 3}
    app = piso.IcoFoam(mesh, nu=0.01, dt=1e-4 * 20 / args.size, comm=comm, check_every=args.check_every, graphs=bool(args.graphs), fvSolution=sol)
    for _ in range(args.warmup):
        app.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its, ms = [], []
    for _ in range(args.steps):
        e0.record()
        st = app.step()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        its.append([s.numIter for s in st])
    b2b = None
    if app._whole is not None:   # sustained rate: steps back to back, iteration counts from the solver's device-side log
        nb = 2 * args.steps
        app.solver.captured_log()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        d0 = comm.p2p_debug() if (comm and comm.p2p) else None
        e0.record()
        for _ in range(nb):
            app.step()
        e1.record(); torch.cuda.synchronize()
        tb = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        log = app.solver.captured_log()
        b2b = {"steps": nb, "ms_per_step": float(tb.item()) / nb, "cg_iterations_total": int(sum(log))}
        if d0 is not None:  # time this rank's kernels spent WAITING for neighbours (exchange flags, in-kernel all-reduces), per step
            d1 = comm.p2p_debug()
            dd = [b - a for a, b in zip(d0, d1)]
            mine = torch.tensor([dd[6] / 1e6 / nb, dd[7] / nb, (dd[1] + dd[2] + dd[4]) / 1e6 / nb, (dd[3] + dd[5]) / nb], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            b2b["comm_wait_per_rank"] = [{"exchange_wait_ms": round(float(t[0]), 4), "exchanges": round(float(t[1]), 1),
                                          "cg_sync_ms": round(float(t[2]), 4), "cg_syncs": round(float(t[3]), 1)} for t in allr]
    t = torch.tensor(ms, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.cpu().numpy()
    if rank == 0:
        tot_it = [int(sum(i)) for i in its]
        print("PISO " + json.dumps({"n": args.size, "cells": args.size ** 3, "n_gpus": world, "transport": ("peer-memory windows" if (comm and comm.p2p) else ("NCCL" if comm else ("none (diagnostic: no communicator)" if world > 1 else "none"))),
                                    "ms_per_step": [round(float(x), 3) for x in ms], "median_ms": float(np.median(ms)), "cg_iterations": its, "cuda_graphs": ("whole step" if app._whole is not None else ("segments" if app._captured else "none")),
                                    "ms_per_cg_iteration_upper_bound": float(np.median(ms / np.maximum(tot_it, 1))), "setup_s": round(setup_s, 1), "back_to_back": b2b}), flush=True)
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
