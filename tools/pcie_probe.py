#!/usr/bin/env python
"""Host<->device copy ceiling of the box, for reading the e2e numbers: pinned 256 MB buffers, H2D alone, D2H alone, both directions
at once (two streams); first on rank 0 alone, then on all ranks simultaneously. Run plain or under torchrun; rank 0 prints JSON lines."""
import json
import os

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 32 * 1024 * 1024
h_in, h_out = torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_in, d_out = torch.empty(n, dtype=torch.float64, device="cuda"), torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 8


def run(mode, active):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    if active:
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for _ in range(REPS):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = REPS * n * 8 * (2 if mode == "both" else 1)
    return nbytes / (float(t.item()) * 1e-3) / 1e9


for who in (("rank 0 alone", lambda: rank == 0), (f"all {world} ranks at once", lambda: True)):
    for mode in ("h2d", "d2h", "both"):
        run(mode, who[1]())  # warm-up
        g = run(mode, who[1]())
        nact = 1 if who[0].startswith("rank 0") else world
        if rank == 0:
            print(json.dumps({"who": who[0], "mode": mode, "GBs_per_gpu": round(g, 1), "GBs_aggregate": round(g * nact, 1)}), flush=True)
    if world == 1:
        break
if world > 1:
    dist.destroy_process_group()
