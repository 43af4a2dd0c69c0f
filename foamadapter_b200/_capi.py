"""ctypes binding of the C ABI in include/fvk.h (libfvk.so). Thin: no logic, no fallbacks."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

from .build import LIB, ROOT, build_lib

c_int32_p = C.POINTER(C.c_int32)
c_double_p = C.POINTER(C.c_double)


class MeshDesc(C.Structure):
    _fields_ = [
        ("nCells", C.c_int32), ("nInternalFaces", C.c_int32), ("nBoundaryFaces", C.c_int32),
        ("nPatches", C.c_int32), ("nPoints", C.c_int32),
        ("points", c_double_p), ("cellVolumes", c_double_p), ("cellCentres", c_double_p),
        ("faceAreas", c_double_p), ("faceCentres", c_double_p), ("magFaceAreas", c_double_p),
        ("faceOwner", c_int32_p), ("faceNeighbour", c_int32_p), ("faceCells", c_int32_p),
        ("bCf", c_double_p), ("bCn", c_double_p), ("bSf", c_double_p), ("bMagSf", c_double_p),
        ("bNf", c_double_p), ("bDelta", c_double_p), ("bWeights", c_double_p),
        ("bDeltaCoeffs", c_double_p), ("patchOffsets", c_int32_p),
        ("nOwnedCells", C.c_int32), ("faceOrder", c_int32_p),
    ]


class FvkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fvk error {code}: {msg}")
        self.code = code


_lib = None


def declared_symbols() -> list[str]:
    """Every function name declared in include/*.h (used by the export test)."""
    names = []
    for h in sorted((ROOT / "include").glob("*.h")):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names += re.findall(r"\b(fvk_\w+)\s*\(", text)
    return sorted(set(names))


def lib() -> C.CDLL:
    """Load (building if needed) libfvk.so. Fails loudly when it cannot be built/loaded."""
    global _lib
    if _lib is None:
        if not LIB.exists():
            build_lib()
        _lib = C.CDLL(str(LIB))
        _lib.fvk_last_error.restype = C.c_char_p
        _lib.fvk_conum_scratch_bytes.restype = C.c_size_t
        if hasattr(_lib, "fvk_pcg_scratch_bytes"):
            _lib.fvk_pcg_scratch_bytes.restype = C.c_size_t
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise FvkError(rc, lib().fvk_last_error().decode())


def ptr(x):
    """Device/host pointer argument: accepts None, int, torch tensor, numpy array."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return C.c_void_p(x.ctypes.data)
    raise TypeError(type(x))
