"""Python mirror of NeoN::finiteVolume::cellCentred (the reference's operator/plugin interface for
the hot path), with the same class and method names and argument meaning; bodies call the C ABI.

Reference: src/NeoN/include/NeoN/finiteVolume/cellCentred/{fields,operators,interpolation,
faceNormalGradient,boundary}/*.hpp. The C++ twin lives in include/NeoN/.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import ops
from .mesh import UnstructuredMesh


@dataclass
class Coeff:
    """dsl::Coeff (src/NeoN/include/NeoN/dsl/coeff.hpp:21-54): scalar x optional per-cell view."""
    value: float = 1.0
    view: torch.Tensor | None = None

    def __mul__(self, rhs):
        if isinstance(rhs, Coeff):
            if rhs.view is not None and self.view is not None:
                return Coeff(self.value * rhs.value, self.view * rhs.view)
            return Coeff(self.value * rhs.value, self.view if self.view is not None else rhs.view)
        return Coeff(self.value * float(rhs), self.view)

    __rmul__ = __mul__


class BoundaryData:
    """fields/boundaryData.hpp:32-215: value/refValue/refGrad T[nB], valueFraction double[nB]."""

    def __init__(self, mesh, ncomp, device):
        shp = (mesh.nBoundaryFaces, 3) if ncomp == 3 else (mesh.nBoundaryFaces,)
        z = lambda s: torch.zeros(s, dtype=torch.float64, device=device)
        self.value, self.refValue, self.refGrad = z(shp), z(shp), z(shp)
        self.valueFraction = z((mesh.nBoundaryFaces,))


_BC_KINDS = {"calculated": ops.BC_CALCULATED, "fixedValue": ops.BC_FIXED_VALUE, "fixedGradient": ops.BC_FIXED_GRADIENT,
             "extrapolated": ops.BC_EXTRAPOLATED, "empty": ops.BC_EMPTY,
             # readers.hpp:43-95 maps zeroGradient -> fixedGradient 0 and noSlip -> fixedValue 0
             "zeroGradient": ops.BC_FIXED_GRADIENT, "noSlip": ops.BC_FIXED_VALUE}


class VolumeField:
    """VolumeField<scalar|Vec3> (fields/volumeField.hpp). bcs: one (type, constant) per patch."""

    def __init__(self, mesh: UnstructuredMesh, name: str, ncomp: int = 1, bcs=None, device="cuda"):
        self.mesh, self.name, self.ncomp = mesh, name, ncomp
        shp = (mesh.nCells, 3) if ncomp == 3 else (mesh.nCells,)
        self.internal = torch.zeros(shp, dtype=torch.float64, device=device)
        self.boundary = BoundaryData(mesh, ncomp, device)
        bcs = bcs if bcs is not None else [("calculated", 0.0)] * mesh.nPatches
        if len(bcs) != mesh.nPatches:
            raise ValueError(f"{name}: {len(bcs)} boundary conditions for {mesh.nPatches} patches")
        self.bcs = []
        for kind, cst in bcs:
            if kind not in _BC_KINDS:
                raise KeyError(f"unknown boundary condition '{kind}'")  # RuntimeSelectionFactory::keyExistsOrError
            if kind in ("zeroGradient", "noSlip", "calculated", "extrapolated", "empty"):
                cst = (0.0, 0.0, 0.0) if ncomp == 3 else 0.0
            self.bcs.append((kind, cst))

    def internalVector(self):
        return self.internal

    def boundaryData(self):
        return self.boundary

    def oldTime(self):
        """fvcc::oldTime(field) (core/database/oldTimeCollection.hpp:146-152): a registered copy, created on first use."""
        if getattr(self, "_old", None) is None:
            self._old = VolumeField(self.mesh, self.name + "_0", self.ncomp, list(self.bcs), self.internal.device)
            self._old.internal.copy_(self.internal)
        return self._old

    def assignable(self, patch: int) -> bool:
        # fixedValue is the only non-assignable BC (fixedValue.hpp:56 vs fixedGradient/extrapolated)
        return _BC_KINDS[self.bcs[patch][0]] != ops.BC_FIXED_VALUE

    def correctBoundaryConditions(self):
        b = self.boundary
        ops.correct_boundary_conditions(self.mesh, [_BC_KINDS[k] for k, _ in self.bcs], [c for _, c in self.bcs],
                                        self.internal, b.value, b.refValue, b.valueFraction, b.refGrad)


class SurfaceField:
    """SurfaceField<T>: internalVector has nInternalFaces + nBoundaryFaces entries
    (fields/surfaceField.hpp:38-53); boundary value duplicates the boundary slice."""

    def __init__(self, mesh, name, ncomp=1, device="cuda"):
        self.mesh, self.name, self.ncomp = mesh, name, ncomp
        shp = (mesh.nFaces, 3) if ncomp == 3 else (mesh.nFaces,)
        self.internal = torch.zeros(shp, dtype=torch.float64, device=device)
        self.bvalue = torch.zeros((mesh.nBoundaryFaces,) + shp[1:], dtype=torch.float64, device=device)

    def internalVector(self):
        return self.internal


class InterpolatedSurfaceField:
    """A surface field DEFINED as the linear interpolate of a volume field and never materialised:
    `imp.laplacian(InterpolatedSurfaceField(rAU, "rAUf"), p)` makes the assembly kernel evaluate gamma_f = w rAU_P + (1 - w) rAU_N
    on the fly with computeLinearInterpolation's arithmetic (interpolation/linear.cpp:30-45) -- same bits as
    SurfaceInterpolation("linear").interpolate(rAU) followed by the laplacian, without the face-sized temporary."""

    def __init__(self, src: "VolumeField", name: str):
        if src.ncomp != 1:
            raise ValueError("only scalar fields can be used as an interpolated diffusivity")
        self.src, self.name, self.mesh, self.ncomp = src, name, src.mesh, 1

    def materialise(self):
        return SurfaceInterpolation(self.mesh, "linear").interpolate(self.src)


class SurfaceInterpolation:
    """interpolation/surfaceInterpolation.hpp:54-69; keys "linear" | "upwind"."""

    def __init__(self, mesh, scheme: str):
        if scheme not in ops.SCHEMES:
            raise KeyError(f"unknown interpolation scheme '{scheme}'")
        self.mesh, self.scheme = mesh, ops.SCHEMES[scheme]

    def interpolate(self, src: VolumeField, dst: SurfaceField | None = None, faceFlux: SurfaceField | None = None):
        if self.scheme == ops.UPWIND and faceFlux is None:
            raise ValueError("limited scheme require a faceFlux")  # upwind.hpp:66-72
        dst = dst or SurfaceField(self.mesh, "phif", src.ncomp, src.internal.device)
        ops.interpolate(self.mesh, src.internal, src.boundary.value, dst.internal, self.scheme,
                        None if faceFlux is None else faceFlux.internal)
        return dst

    def weight(self, faceFlux: SurfaceField | None = None, dst: SurfaceField | None = None):
        dst = dst or SurfaceField(self.mesh, "weight", 1)
        ops.interpolation_weights(self.mesh, dst.internal, dst.bvalue, self.scheme,
                                  None if faceFlux is None else faceFlux.internal)
        return dst


class FaceNormalGradient:
    """faceNormalGradient/faceNormalGradient.hpp:50-54; key "uncorrected"."""

    def __init__(self, mesh, scheme="uncorrected"):
        if scheme != "uncorrected":
            raise KeyError(f"unknown faceNormalGradient scheme '{scheme}'")
        self.mesh = mesh

    def faceNormalGrad(self, phi: VolumeField, dst: SurfaceField | None = None):
        dst = dst or SurfaceField(self.mesh, "snGrad", phi.ncomp, phi.internal.device)
        ops.face_normal_grad(self.mesh, phi.internal, phi.boundary.value, dst.internal)
        return dst


class GaussGreenDiv:
    """operators/gaussGreenDiv.hpp:75-81: div(divPhi, faceFlux, phi, operatorScaling)."""

    def __init__(self, mesh, scheme="linear"):
        self.mesh, self.interp = mesh, SurfaceInterpolation(mesh, scheme)

    def div(self, divPhi: torch.Tensor, faceFlux: SurfaceField, phi: VolumeField, operatorScaling=Coeff(), mode=ops.ACC_SCALE):
        """Default mode reproduces computeDiv: accumulate into divPhi, then scale all of it."""
        return ops.div(self.mesh, faceFlux.internal, phi.internal, phi.boundary.value, divPhi, self.interp.scheme,
                       operatorScaling.value, operatorScaling.view, mode)


class GaussGreenGrad:
    """operators/gaussGreenGrad.hpp: grad(phi, gradPhi); always linear."""

    def __init__(self, mesh):
        self.mesh = mesh

    def grad(self, phi: VolumeField, gradPhi: torch.Tensor | None = None):
        if gradPhi is None:
            gradPhi = torch.empty((self.mesh.nCells, 3), dtype=torch.float64, device=phi.internal.device)
            return ops.grad(self.mesh, phi.internal, phi.boundary.value, gradPhi, ops.SET)
        return ops.grad(self.mesh, phi.internal, phi.boundary.value, gradPhi, ops.ACC_SCALE)


class GaussGreenLaplacian:
    """operators/gaussGreenLaplacian.hpp: laplacian(lapPhi, gamma, phi, operatorScaling); gamma is
    ignored by the explicit reference kernel (gaussGreenLaplacian.cpp:14)."""

    def __init__(self, mesh, scheme="uncorrected"):
        self.mesh, self.fng = mesh, FaceNormalGradient(mesh, scheme)

    def laplacian(self, lapPhi, gamma, phi: VolumeField, operatorScaling=Coeff(), mode=ops.ACC_SCALE):
        return ops.laplacian(self.mesh, phi.internal, phi.boundary.value, lapPhi, operatorScaling.value,
                             operatorScaling.view, mode)


def computeCoNum(faceFlux: SurfaceField, dt: float) -> float:
    """auxiliary/coNum.cpp:18-96; returns maxCoNum (device->host scalar, like the reference)."""
    res = ops.conum(faceFlux.mesh, faceFlux.internal, dt)
    return float(res[0].item())


def read_field_file(path, nCells: int):
    """internalField of an OpenFOAM ASCII field file -> numpy array [nCells] or [nCells, 3] (fvk_fieldfile_read_internal)."""
    import ctypes as C
    import numpy as np
    from ._capi import check, lib
    nc = C.c_int32(0)
    out = np.zeros(3 * max(nCells, 1))
    check(lib().fvk_fieldfile_read_internal(str(path).encode(), C.c_int32(nCells), C.byref(nc), out.ctypes.data_as(C.c_void_p), C.c_int64(out.size)))
    return out[:nCells].copy() if nc.value == 1 else out[: 3 * nCells].reshape(nCells, 3).copy()


def read_patch_conditions(path, patch_names, patch_sizes=None):
    """[(type, value)] per patch from the boundaryField of an OpenFOAM ASCII field file (readers.hpp:43-95 of the reference
    maps the same dictionaries). value: None (no `value` entry), a constant (uniform, or a nonuniform list whose entries are all
    equal), or -- with patch_sizes given -- the per-face array [n] / [n, 3] of a `value nonuniform List<...>` entry, which is
    what OpenFOAM itself writes into time directories."""
    import ctypes as C
    import numpy as np
    from ._capi import check, lib
    bcs = []
    for i, name in enumerate(patch_names):
        n = int(patch_sizes[i]) if patch_sizes is not None else 1
        typ, has, nc = C.create_string_buffer(128), C.c_int32(0), C.c_int32(0)
        val = np.zeros(3 * max(n, 1))
        check(lib().fvk_fieldfile_read_patch(str(path).encode(), name.encode(), typ, C.c_int32(128), C.c_int32(n if patch_sizes is not None else 1),
                                             C.byref(has), C.byref(nc), val.ctypes.data_as(C.c_void_p), C.c_int64(val.size)))
        v = None
        if has.value and n > 0:
            a = val[:n].copy() if nc.value == 1 else val[: 3 * n].reshape(n, 3).copy()
            if (a == a[0]).all():
                v = float(a[0]) if nc.value == 1 else tuple(float(x) for x in a[0])
            else:
                v = a
        bcs.append((typ.value.decode(), v))
    return bcs


def read_volume_field(mesh, path, name=None, device="cuda"):
    """VolumeField from an OpenFOAM ASCII field file on a mesh whose patches carry names (e.g. MeshDesc.from_polymesh).
    Patches whose `value` is a genuinely nonuniform list keep their per-face values: the patch becomes `calculated` (the BC
    kernel leaves it alone) with value / refValue = the list and, for fixedValue, valueFraction = 1 (fixedValue.hpp:21-43)."""
    import os
    import numpy as np
    internal = read_field_file(path, mesh.nOwned)
    ncomp = 3 if internal.ndim == 2 else 1
    zero = (0.0, 0.0, 0.0) if ncomp == 3 else 0.0
    off = mesh.patch_offsets
    raw = read_patch_conditions(path, mesh.patch_names, [off[i + 1] - off[i] for i in range(mesh.nPatches)])
    bcs, perFace = [], []
    for i, (t, v) in enumerate(raw):
        if isinstance(v, np.ndarray):
            perFace.append((i, t, v))
            bcs.append(("calculated", zero))
        else:
            bcs.append((t, v if v is not None else zero))
    f = VolumeField(mesh, name or os.path.basename(str(path)), ncomp, bcs, device)
    f.internal[: mesh.nOwned].copy_(torch.from_numpy(internal))
    f.correctBoundaryConditions()
    for i, t, v in perFace:
        sl = slice(off[i], off[i + 1])
        tv = torch.from_numpy(np.ascontiguousarray(v)).to(f.internal.device)
        f.boundary.value[sl] = tv
        if t in ("fixedValue", "noSlip"):
            f.boundary.refValue[sl] = tv
            f.boundary.valueFraction[sl] = 1.0
    return f


def write_field_file(path, internal, patch_names, bcs, name=None):
    """Write an OpenFOAM ASCII vol<Scalar|Vector>Field file (fvk_fieldfile_write). bcs: [(type, value or None)] per patch."""
    import ctypes as C
    import os
    import numpy as np
    from ._capi import check, lib
    a = np.ascontiguousarray(internal, dtype=np.float64)
    ncomp = 3 if a.ndim == 2 else 1
    n = a.shape[0]
    names = (C.c_char_p * len(patch_names))(*[p.encode() for p in patch_names])
    types = (C.c_char_p * len(bcs))(*[t.encode() for t, _ in bcs])
    has = (C.c_int32 * len(bcs))(*[0 if v is None else 1 for _, v in bcs])
    vals = np.zeros(max(len(bcs), 1) * ncomp)
    for i, (_, v) in enumerate(bcs):
        if v is not None:
            vals[i * ncomp:(i + 1) * ncomp] = v
    check(lib().fvk_fieldfile_write(str(path).encode(), (name or os.path.basename(str(path))).encode(), C.c_int32(ncomp), C.c_int32(n),
                                    a.ctypes.data_as(C.c_void_p), C.c_int32(len(bcs)), names, types, has, vals.ctypes.data_as(C.c_void_p)))
    return path
