"""The C-ABI library loads and exports every symbol include/*.h declares; without a GPU the compute
entry points fail loudly (no CPU fallback). CPU only, no compute calls."""
import ctypes as C

import pytest
import torch

from foamadapter_b200 import _capi
from foamadapter_b200.mesh import MeshDesc


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    names = _capi.declared_symbols()
    assert len(names) > 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.fvk_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    d = MeshDesc.uniform_1d(10)
    h = C.c_void_p()
    rc = _capi.lib().fvk_mesh_create(C.byref(d.c), C.byref(h))
    assert rc == 2  # FVK_ENODEVICE
    assert b"no CUDA device" in _capi.lib().fvk_last_error() or b"cuda" in _capi.lib().fvk_last_error().lower()


def test_new_entry_points_validate_arguments():
    """Argument checks of the r1e additions run before any device work, so they hold on a CPU-only box."""
    lib = _capi.lib()
    assert lib.fvk_mesh_set_tile_phase(None, C.c_int(1)) == 1          # FVK_EINVAL
    assert lib.fvk_comm_p2p_export(None, None) == 1
    assert lib.fvk_comm_p2p_connect(None, None) == 1
    assert lib.fvk_comm_p2p_enabled(None) == 0
    assert lib.fvk_probe_streams(C.c_int(0), None, C.c_int64(0), None, None) == 1
    assert lib.fvk_set_brick_config(C.c_int(0), C.c_int(0), C.c_int(0)) == 0
    h = C.c_void_p()
    assert lib.fvk_comm_create(C.c_int(0), C.c_int(1), None, C.byref(h)) == 0  # a 1-rank comm needs neither NCCL nor a device
    blob = (C.c_char * 512)()
    if not torch.cuda.is_available():
        assert lib.fvk_comm_p2p_export(h, blob) in (2, 3)  # FVK_ENODEVICE / FVK_ECUDA: no CPU stand-in for the window
    assert lib.fvk_comm_p2p_enabled(h) == 0
    lib.fvk_comm_destroy(h)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py contract: stdout carries ONE JSON line (anything a native library prints goes to stderr); the reference
    arm runs without a GPU. Tiny mesh so the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--mesh", "12", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fp64_face_ops_per_s" and d["unit"] == "face-ops/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True
