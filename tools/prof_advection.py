#!/usr/bin/env python
"""Driver for ncu launch lists: a few forward-Euler scalarAdvection steps on the n^3 unit cube (BASELINE configs[3])."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import advection as adv, mesh as M  # noqa: E402
from foamadapter_b200.mesh import UnstructuredMesh  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh = UnstructuredMesh(adv.advection_desc(n, True))
C = mesh.to_host(M.CELL_CENTRES).reshape(-1, 3)
U, T = adv.init_fields_columns(C, n * n)
app = adv.ScalarAdvection(mesh, 0.1 / n, 3.0, U=U, T=T)
for _ in range(steps):
    app.step()
torch.cuda.synchronize()
print("done", float(app.T.internal.max()))
