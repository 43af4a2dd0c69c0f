"""Typed Python wrappers over the C ABI (include/fvk.h). Tensors are torch CUDA tensors (fp64 /
int32) used purely as device-memory handles; every call goes to libfvk.so on torch's current stream.
There is no fallback: without the library or a GPU these raise."""
from __future__ import annotations

import ctypes as C

import torch

from ._capi import check, lib, ptr
from .mesh import UnstructuredMesh

SET, ACC_SCALE, ADD = 0, 1, 2
LINEAR, UPWIND = 0, 1
SCHEMES = {"linear": LINEAR, "upwind": UPWIND}
BC_CALCULATED, BC_FIXED_VALUE, BC_FIXED_GRADIENT, BC_EXTRAPOLATED, BC_EMPTY = range(5)

gpu_launches = 0  # kernels launched through this module (bench.py reports it)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, n, name, comps=None):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError(f"{name}: need a contiguous CUDA float64 tensor")
    if t.shape[0] != n or (comps == 3 and (t.ndim != 2 or t.shape[1] != 3)) or (comps == 1 and t.ndim != 1):
        raise ValueError(f"{name}: bad shape {tuple(t.shape)}, expected leading {n}, comps {comps}")


def _count(k=1):
    global gpu_launches
    gpu_launches += k


def div(mesh: UnstructuredMesh, faceFlux, phi, phiB, out, scheme=LINEAR, coeff=1.0, coeffView=None, mode=SET):
    vec = phi.ndim == 2
    _chk(faceFlux, mesh.nFaces, "faceFlux", 1); _chk(phi, mesh.nCells, "phi", 3 if vec else 1)
    _chk(phiB, mesh.nBoundaryFaces, "phiB", 3 if vec else 1); _chk(out, mesh.nCells, "out", 3 if vec else 1)
    _chk(coeffView, mesh.nCells, "coeffView", 1)
    fn = lib().fvk_div_v if vec else lib().fvk_div_s
    check(fn(mesh.handle, C.c_int(scheme), ptr(faceFlux), ptr(phi), ptr(phiB), C.c_double(coeff), ptr(coeffView),
             ptr(out), C.c_int(mode), _stream()))
    _count()
    return out


def div_forward_euler(mesh: UnstructuredMesh, faceFlux, phiOld, phiB, dt, out, scheme=LINEAR, coeff=1.0, coeffView=None):
    """fvk_div_forward_euler_s: out = phiOld - dt * div(faceFlux, phiOld) in one pass (forwardEuler of ddt + div, scalar)"""
    _chk(faceFlux, mesh.nFaces, "faceFlux", 1); _chk(phiOld, mesh.nCells, "phiOld", 1)
    _chk(phiB, mesh.nBoundaryFaces, "phiB", 1); _chk(out, mesh.nCells, "out", 1); _chk(coeffView, mesh.nCells, "coeffView", 1)
    check(lib().fvk_div_forward_euler_s(mesh.handle, C.c_int(scheme), ptr(faceFlux), ptr(phiOld), ptr(phiB), C.c_double(coeff), ptr(coeffView),
                                        C.c_double(dt), ptr(out), _stream()))
    _count()
    return out


def grad(mesh, phi, phiB, out, mode=SET):
    _chk(phi, mesh.nCells, "phi", 1); _chk(phiB, mesh.nBoundaryFaces, "phiB", 1); _chk(out, mesh.nCells, "out", 3)
    check(lib().fvk_grad_s(mesh.handle, ptr(phi), ptr(phiB), ptr(out), C.c_int(mode), _stream()))
    _count()
    return out


def laplacian(mesh, phi, phiB, out, coeff=1.0, coeffView=None, mode=SET):
    vec = phi.ndim == 2
    _chk(phi, mesh.nCells, "phi", 3 if vec else 1); _chk(phiB, mesh.nBoundaryFaces, "phiB", 3 if vec else 1)
    _chk(out, mesh.nCells, "out", 3 if vec else 1); _chk(coeffView, mesh.nCells, "coeffView", 1)
    fn = lib().fvk_laplacian_v if vec else lib().fvk_laplacian_s
    check(fn(mesh.handle, ptr(phi), ptr(phiB), C.c_double(coeff), ptr(coeffView), ptr(out), C.c_int(mode), _stream()))
    _count()
    return out


def surface_integrate(mesh, flux, out, coeff=1.0, coeffView=None, mode=SET):
    vec = flux.ndim == 2
    _chk(flux, mesh.nFaces, "flux", 3 if vec else 1); _chk(out, mesh.nCells, "out", 3 if vec else 1)
    _chk(coeffView, mesh.nCells, "coeffView", 1)
    fn = lib().fvk_surface_integrate_v if vec else lib().fvk_surface_integrate_s
    check(fn(mesh.handle, ptr(flux), C.c_double(coeff), ptr(coeffView), ptr(out), C.c_int(mode), _stream()))
    _count()
    return out


def interpolate(mesh, phi, phiB, outFace, scheme=LINEAR, faceFlux=None):
    vec = phi.ndim == 2
    _chk(phi, mesh.nCells, "phi", 3 if vec else 1); _chk(phiB, mesh.nBoundaryFaces, "phiB", 3 if vec else 1)
    _chk(outFace, mesh.nFaces, "outFace", 3 if vec else 1); _chk(faceFlux, mesh.nFaces, "faceFlux", 1)
    fn = lib().fvk_interpolate_v if vec else lib().fvk_interpolate_s
    check(fn(mesh.handle, C.c_int(scheme), ptr(faceFlux), ptr(phi), ptr(phiB), ptr(outFace), _stream()))
    _count()
    return outFace


def interpolation_weights(mesh, wFace, wBoundary=None, scheme=LINEAR, faceFlux=None):
    _chk(wFace, mesh.nFaces, "wFace", 1); _chk(wBoundary, mesh.nBoundaryFaces, "wBoundary", 1)
    _chk(faceFlux, mesh.nFaces, "faceFlux", 1)
    check(lib().fvk_interpolation_weights(mesh.handle, C.c_int(scheme), ptr(faceFlux), ptr(wFace), ptr(wBoundary), _stream()))
    _count()
    return wFace


def face_normal_grad(mesh, phi, phiB, outFace):
    vec = phi.ndim == 2
    _chk(phi, mesh.nCells, "phi", 3 if vec else 1); _chk(phiB, mesh.nBoundaryFaces, "phiB", 3 if vec else 1)
    _chk(outFace, mesh.nFaces, "outFace", 3 if vec else 1)
    fn = lib().fvk_face_normal_grad_v if vec else lib().fvk_face_normal_grad_s
    check(fn(mesh.handle, ptr(phi), ptr(phiB), ptr(outFace), _stream()))
    _count()
    return outFace


def conum(mesh, faceFlux, dt, result=None, scratch=None):
    """Returns a device tensor [maxCoNum, meanCoNum]; no host sync."""
    _chk(faceFlux, mesh.nFaces, "faceFlux", 1)
    if result is None:
        result = torch.empty(2, dtype=torch.float64, device=faceFlux.device)
    if scratch is None:
        scratch = torch.empty(lib().fvk_conum_scratch_bytes(mesh.handle) // 8, dtype=torch.float64, device=faceFlux.device)
    check(lib().fvk_conum(mesh.handle, ptr(faceFlux), C.c_double(dt), ptr(result), ptr(scratch), _stream()))
    if getattr(mesh, "_conum_launches", None) is None:  # block topology proven: regular-cell kernel + irregular-cell list + fold
        from . import mesh as _m
        mesh._conum_launches = 3 if mesh.size(_m.AFFINE_TOPOLOGY) else 2
    _count(mesh._conum_launches)
    return result


def correct_boundary_conditions(mesh, kinds, consts, internal, value, refValue, valueFraction, refGrad):
    ncomp = 3 if internal.ndim == 2 else 1
    n = mesh.nPatches
    if len(kinds) != n:
        raise ValueError(f"need one boundary condition per patch ({n}), got {len(kinds)}")
    k = (C.c_int32 * n)(*[int(x) for x in kinds])
    flat = []
    for c in consts:
        flat += [float(x) for x in (c if ncomp == 3 else [c])]
    v = (C.c_double * (n * ncomp))(*flat)
    check(lib().fvk_correct_boundary_conditions(mesh.handle, C.c_int(ncomp), k, v, ptr(internal), ptr(value),
                                                ptr(refValue), ptr(valueFraction), ptr(refGrad), _stream()))
    _count()


# ---- implicit assembly ---------------------------------------------------------------------------
TERM_DDT, TERM_DIV, TERM_LAPLACIAN, TERM_SOURCE = range(4)


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("scheme", C.c_int32), ("coeff", C.c_double), ("coeffView", C.c_void_p),
                ("faceField", C.c_void_p), ("cellField", C.c_void_p), ("dt", C.c_double), ("gammaCell", C.c_void_p), ("gammaBoundary", C.c_void_p)]


class _BField(C.Structure):
    _fields_ = [("value", C.c_void_p), ("refValue", C.c_void_p), ("valueFraction", C.c_void_p), ("refGrad", C.c_void_p)]


def _p(t):
    return None if t is None else t.data_ptr()


def assemble(mesh, terms, boundary, values, rhs, bcMatrix=None, bcRhs=None, accumulate=False):
    """terms: list of dicts {kind, scheme?, coeff?, coeffView?, faceField?, cellField?, dt?} applied in order
    (Expression::implicitOperation). boundary: BoundaryData of the unknown field (or None). A Vec3 right-hand side with
    1-D `values` selects the compact Vec3 system (la.LinearSystem(compact=True), fvk_assemble_vc)."""
    vec = rhs.ndim == 2
    compact = vec and values.ndim == 1
    arr = (_Term * len(terms))()
    keep = []
    for i, t in enumerate(terms):
        keep += [t.get("coeffView"), t.get("faceField"), t.get("cellField")]
        arr[i] = _Term(int(t["kind"]), int(t.get("scheme", 0)), float(t.get("coeff", 1.0)), _p(t.get("coeffView")),
                       _p(t.get("faceField")), _p(t.get("cellField")), float(t.get("dt", 0.0)), _p(t.get("gammaCell")), _p(t.get("gammaBoundary")))
    bf = None
    if boundary is not None:
        bf = C.byref(_BField(_p(boundary.value), _p(boundary.refValue), _p(boundary.valueFraction), _p(boundary.refGrad)))
    fn = (lib().fvk_assemble_vc if compact else lib().fvk_assemble_v) if vec else lib().fvk_assemble_s
    check(fn(mesh.handle, C.c_int(len(terms)), arr, bf, ptr(values), ptr(rhs), ptr(bcMatrix), ptr(bcRhs),
             C.c_int(1 if accumulate else 0), _stream()))
    _count()


def ddt_explicit(mesh, field, oldField, dt, source):
    check(lib().fvk_ddt_explicit(mesh.handle, C.c_int(3 if field.ndim == 2 else 1), ptr(field), ptr(oldField), C.c_double(dt),
                                 ptr(source), _stream()))
    _count()


def source_explicit(mesh, k, field, source, coeff=1.0, coeffView=None):
    check(lib().fvk_source_explicit(mesh.handle, C.c_int(3 if field.ndim == 2 else 1), ptr(k), ptr(field), C.c_double(coeff),
                                    ptr(coeffView), ptr(source), _stream()))
    _count()


def rhs_sub_source(mesh, src, rhs):
    check(lib().fvk_rhs_sub_source(mesh.handle, C.c_int(3 if rhs.ndim == 2 else 1), ptr(src), ptr(rhs), _stream()))
    _count()


def bc_coeff_indices(mesh, matrixIdxs, rhsIdxs):
    check(lib().fvk_bc_coeff_indices(mesh.handle, ptr(matrixIdxs), ptr(rhsIdxs), _stream()))
    _count()


# ---- PISO glue -----------------------------------------------------------------------------------
def rAU_HbyA(mesh, valuesV, rhsV, U, rAU, HbyA=None):
    """valuesV: Vec3[nnz] (reference layout) or double[nnz] (compact momentum matrix)"""
    fn = lib().fvk_rAU_HbyA_c if valuesV.ndim == 1 else lib().fvk_rAU_HbyA
    check(fn(mesh.handle, ptr(valuesV), ptr(rhsV), ptr(U), ptr(rAU), ptr(HbyA), _stream()))
    _count()


def copy_patches(mesh, mask, srcB, dstB):
    m = (C.c_int32 * mesh.nPatches)(*[int(bool(x)) for x in mask])
    check(lib().fvk_copy_patches(mesh.handle, C.c_int(3 if srcB.ndim == 2 else 1), m, ptr(srcB), ptr(dstB), _stream()))
    _count()


def flux(mesh, U, Ub, outFace, outB=None):
    check(lib().fvk_flux(mesh.handle, ptr(U), ptr(Ub), ptr(outFace), ptr(outB), _stream()))
    _count()


def update_face_velocity(mesh, values, bcMatrix, bcRhs, p, predPhi, predPhiB, phi, phiB):
    check(lib().fvk_update_face_velocity(mesh.handle, ptr(values), ptr(bcMatrix), ptr(bcRhs), ptr(p), ptr(predPhi),
                                         ptr(predPhiB), ptr(phi), ptr(phiB), _stream()))
    _count()


def update_velocity(mesh, HbyA, rAU, gradP, U):
    check(lib().fvk_update_velocity(mesh.handle, ptr(HbyA), ptr(rAU), ptr(gradP), ptr(U), _stream()))
    _count()


def update_velocity_grad(mesh, HbyA, rAU, p, pB, U):
    """U = HbyA - rAU * grad(p), fused (fvk_update_velocity_grad)"""
    check(lib().fvk_update_velocity_grad(mesh.handle, ptr(HbyA), ptr(rAU), ptr(p), ptr(pB), ptr(U), _stream()))
    _count()


def rhs_sub_surface_integrate(mesh, flux, rhs, coeff=1.0, coeffView=None):
    """rhs -= surfaceIntegrate(flux, coeff) * V, fused (fvk_rhs_sub_surface_integrate_s)"""
    check(lib().fvk_rhs_sub_surface_integrate_s(mesh.handle, ptr(flux), C.c_double(coeff), ptr(coeffView), ptr(rhs), _stream()))
    _count()


def set_reference(mesh, refCell, refValue, values, rhs):
    check(lib().fvk_set_reference(mesh.handle, C.c_int32(refCell), C.c_double(refValue), ptr(values), ptr(rhs), _stream()))
    _count()


def diag(mesh, values, out):
    check(lib().fvk_diag(mesh.handle, C.c_int(3 if values.ndim == 2 else 1), ptr(values), ptr(out), _stream()))
    _count()
