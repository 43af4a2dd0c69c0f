#!/usr/bin/env python
"""Wall clock of the once-per-mesh setup at N^3 (GPU box): block-mesh generator, fvk_mesh_create (phases to stderr with
FVK_SETUP_TIMING=1), and the 8-way decomposition of rank 0."""
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("FVK_SETUP_TIMING", "1")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from foamadapter_b200.decomp import Decomposition  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.cuda.init(); torch.zeros(1, device="cuda")
t = time.perf_counter(); d = MeshDesc.block(n, n, n, 0.1, 0.1, 0.01); t1 = time.perf_counter()
print(f"[setup] MeshDesc.block {n}^3: {t1 - t:.2f} s", flush=True)
gm = UnstructuredMesh(d); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"[setup] UnstructuredMesh (fvk_mesh_create): {t2 - t1:.2f} s", flush=True)
print(f"[setup] device memory in use: {(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30:.2f} GiB", flush=True)
del gm
t3 = time.perf_counter(); dec = Decomposition(d, 8, 0, n=(2, 2, 2)); t4 = time.perf_counter()
print(f"[setup] Decomposition 2x2x2 rank 0: {t4 - t3:.2f} s", flush=True)
lm = UnstructuredMesh(dec.desc); torch.cuda.synchronize(); t5 = time.perf_counter()
print(f"[setup] sub-domain UnstructuredMesh: {t5 - t4:.2f} s", flush=True)
