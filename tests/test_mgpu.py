"""Multi-GPU parity under torchrun (needs >= 2 GPUs on the box, skipped otherwise): tests/mgpu_check.py runs the halo
exchange, the explicit operators (bit-exact on owned cells), the distributed Jacobi-CG and two neoIcoFoam steps against the
single-domain CPU oracle, once over NCCL and once over the peer-memory windows."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_parity_on_both_transports():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(ROOT / "tests" / "mgpu_check.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "MGPU CHECK OK on 2 ranks (NCCL)" in r.stdout and "MGPU CHECK OK on 2 ranks (peer-memory windows)" in r.stdout
