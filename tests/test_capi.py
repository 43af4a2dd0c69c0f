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
