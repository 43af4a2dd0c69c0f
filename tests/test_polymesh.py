"""CPU tests of the product-side OpenFOAM polyMesh reader / writer (fvk_polymesh_read / _write, host code of libfvk):
round trips of generated block meshes are bit-exact in every array of the mesh description (topology, ordering,
primitiveMesh geometry, boundary flattening with `empty` patches), the reader agrees with the golden fixtures extracted
from the reference's committed polyMesh directories and with the oracle's numpy restatement of primitiveMesh on a
polyhedral (prism) mesh, and malformed input fails loudly."""
import numpy as np
import pytest

from foamadapter_b200._capi import FvkError
from foamadapter_b200.mesh import PATCHES_3DCUBE, PATCHES_CAVITY2D, PATCHES_CAVITY3D, MeshDesc
from oracle import polymesh as opoly

ARRAYS = ["cellVolumes", "cellCentres", "faceAreas", "faceCentres", "magFaceAreas", "faceOwner", "faceNeighbour", "faceCells",
          "bCf", "bCn", "bSf", "bMagSf", "bNf", "bDelta", "bWeights", "bDeltaCoeffs", "patchOffsets", "points"]


@pytest.mark.parametrize("dims,box,patches", [((5, 5, 1), (0.1, 0.1, 0.01), PATCHES_CAVITY2D), ((4, 3, 2), (1.0, 0.5, 0.25), PATCHES_3DCUBE),
                                              ((7, 5, 4), (0.7, 0.3, 0.9), PATCHES_CAVITY3D), ((1, 1, 1), (1.0, 1.0, 1.0), PATCHES_3DCUBE)])
def test_block_mesh_round_trip_is_bit_exact(tmp_path, dims, box, patches):
    g = MeshDesc.block(*dims, *box, patches=patches, with_points=True)
    d = g.write_polymesh(tmp_path / "polyMesh", patches)
    r = MeshDesc.from_polymesh(d)
    assert (r.nCells, r.nInternalFaces, r.nBoundaryFaces, r.nPatches) == (g.nCells, g.nInternalFaces, g.nBoundaryFaces, g.nPatches)
    assert r.patch_names == g.patch_names
    assert r.patch_types == ["wall"] * g.nPatches      # `empty` patches are dropped like fvPatch::size() == 0
    for name in ARRAYS:
        assert np.array_equal(r.array(name), g.array(name)), name


def _write_raw(d, pts, faces, owner, nei, names, types, sizes):
    import ctypes as C
    from foamadapter_b200._capi import check, lib
    d.mkdir(parents=True, exist_ok=True)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    off = np.cumsum([0] + [len(f) for f in faces]).astype(np.int32)
    fp = np.array([p for f in faces for p in f], dtype=np.int32)
    ow, ne = np.ascontiguousarray(owner, dtype=np.int32), np.ascontiguousarray(nei, dtype=np.int32)
    cn = (C.c_char_p * len(names))(*[str(n).encode() for n in names])
    ct = (C.c_char_p * len(types))(*[str(t).encode() for t in types])
    check(lib().fvk_polymesh_write(str(d).encode(), C.c_int32(len(pts)), pts.ctypes.data_as(C.c_void_p), C.c_int32(len(faces)),
                                   off.ctypes.data_as(C.c_void_p), fp.ctypes.data_as(C.c_void_p), ow.ctypes.data_as(C.c_void_p),
                                   C.c_int32(len(ne)), ne.ctypes.data_as(C.c_void_p), C.c_int32(len(names)), cn, ct,
                                   (C.c_int32 * len(sizes))(*[int(x) for x in sizes])))
    return d


@pytest.mark.parametrize("fixture", ["setup_operator", "setup_stencil3D", "setup_unstructuredMesh", "setup_advection",
                                     "setup_pressureVelocityCoupling", "setup_compatibility"])
def test_reader_on_the_reference_fixture_meshes(tmp_path, fixture):
    """tests/golden/*.npz hold the raw content of the polyMesh directories the reference commits
    (test/setup_*/constant/polyMesh, extracted by tests/golden/make_golden.py). Written back as ASCII files and read by the
    product reader, they must give the reference's topology, the oracle's primitiveMesh geometry and the flattened
    boundary of readOpenFOAMMesh (empty patches dropped)."""
    from pathlib import Path
    z = np.load(Path(__file__).parent / "golden" / f"{fixture}.npz")
    if "faces" not in z.files:
        pytest.skip("fixture without a mesh")
    faces = [list(f) for f in z["faces"]]
    d = _write_raw(tmp_path / "polyMesh", z["points"], faces, z["owner"], z["neighbour"], z["patch_names"], z["patch_types"], z["patch_size"])
    r = MeshDesc.from_polymesh(d)
    nI = len(z["neighbour"])
    keep = [i for i, t in enumerate(z["patch_types"]) if str(t) != "empty"]
    bsel = np.concatenate([np.arange(z["patch_start"][i], z["patch_start"][i] + z["patch_size"][i]) for i in keep]) if keep else np.zeros(0, int)
    sel = np.concatenate([np.arange(nI), bsel]).astype(int)
    assert r.nInternalFaces == nI and r.nBoundaryFaces == len(bsel) and r.nCells == int(max(z["owner"].max(), z["neighbour"].max())) + 1
    assert r.patch_names == [str(z["patch_names"][i]) for i in keep]
    assert np.array_equal(r.array("faceOwner"), z["owner"][sel]) and np.array_equal(r.array("faceNeighbour"), z["neighbour"])
    assert np.array_equal(r.array("patchOffsets"), np.concatenate([[0], np.cumsum([z["patch_size"][i] for i in keep])]))
    Cf, Sf = opoly.face_geometry(z["points"], [np.array(f) for f in faces])
    Cc, V = opoly.cell_geometry(Cf, Sf, z["owner"], z["neighbour"], r.nCells)
    assert np.allclose(r.array("faceCentres"), Cf[sel], rtol=1e-14, atol=1e-18) and np.allclose(r.array("faceAreas"), Sf[sel], rtol=1e-14, atol=1e-18)
    assert np.allclose(r.array("cellVolumes"), V, rtol=1e-14) and np.allclose(r.array("cellCentres"), Cc, rtol=1e-14, atol=1e-18)
    # the block generator reproduces the same mesh (what tests/test_blockmesh.py pins): same topology, geometry to the
    # ~3e-16 absolute noise the fixture's points carry
    from tests.helpers import FIXTURE_BLOCKS
    if fixture in FIXTURE_BLOCKS:
        g = MeshDesc.block(*FIXTURE_BLOCKS[fixture][0], *FIXTURE_BLOCKS[fixture][1], patches=FIXTURE_BLOCKS[fixture][2])
        for name in ("faceOwner", "faceNeighbour", "faceCells", "patchOffsets"):
            assert np.array_equal(r.array(name), g.array(name)), name
        for name in ("cellVolumes", "cellCentres", "faceAreas", "faceCentres", "magFaceAreas", "bCf", "bCn", "bSf", "bMagSf", "bNf",
                     "bDelta", "bWeights", "bDeltaCoeffs"):
            x, y = r.array(name), g.array(name)
            assert np.abs(x - y).max() <= 1e-12 * np.abs(y).max(), name


def _write_prisms(tmp_path):
    """Unit cube cut along the x = y diagonal into two triangular prisms: triangles, quads, one internal face."""
    pts = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.float64)
    faces = [[0, 4, 6, 2],                      # internal: the diagonal plane, owner 0 (normal towards cell 1)
             [0, 2, 1], [4, 5, 6], [0, 1, 5, 4], [1, 2, 6, 5],      # cell 0: bottom, top, y = 0, x = 1 (outward normals)
             [0, 3, 2], [4, 6, 7], [2, 3, 7, 6], [3, 0, 4, 7]]      # cell 1: bottom, top, y = 1, x = 0
    owner = [0, 0, 0, 0, 0, 1, 1, 1, 1]
    nei = [1]
    import ctypes as C
    from foamadapter_b200._capi import check, lib
    d = tmp_path / "polyMesh"
    d.mkdir()
    off = np.cumsum([0] + [len(f) for f in faces]).astype(np.int32)
    fp = np.array([p for f in faces for p in f], dtype=np.int32)
    ow, ne = np.array(owner, dtype=np.int32), np.array(nei, dtype=np.int32)
    names = (C.c_char_p * 2)(b"walls0", b"walls1")
    types = (C.c_char_p * 2)(b"wall", b"patch")
    check(lib().fvk_polymesh_write(str(d).encode(), C.c_int32(8), pts.ctypes.data_as(C.c_void_p), C.c_int32(9), off.ctypes.data_as(C.c_void_p),
                                   fp.ctypes.data_as(C.c_void_p), ow.ctypes.data_as(C.c_void_p), C.c_int32(1), ne.ctypes.data_as(C.c_void_p),
                                   C.c_int32(2), names, types, (C.c_int32 * 2)(4, 4)))
    return d, pts, faces, ow, ne


def test_polyhedral_mesh_matches_the_oracle_geometry(tmp_path):
    d, pts, faces, ow, ne = _write_prisms(tmp_path)
    r = MeshDesc.from_polymesh(d)
    assert (r.nCells, r.nInternalFaces, r.nBoundaryFaces, r.nPatches) == (2, 1, 8, 2)
    assert r.patch_names == ["walls0", "walls1"] and r.patch_types == ["wall", "patch"]
    Cf, Sf = opoly.face_geometry(pts, [np.array(f) for f in faces])
    Cc, V = opoly.cell_geometry(Cf, Sf, ow, ne, 2)
    assert np.allclose(r.array("faceCentres"), Cf, rtol=1e-15, atol=1e-16) and np.allclose(r.array("faceAreas"), Sf, rtol=1e-15, atol=1e-16)
    assert np.allclose(r.array("cellVolumes"), V, rtol=1e-15) and np.allclose(r.array("cellCentres"), Cc, rtol=1e-15)
    assert np.allclose(r.array("cellVolumes"), [0.5, 0.5], rtol=1e-15)
    # closed cells: the outward face area vectors of a cell sum to zero
    s = np.zeros((2, 3))
    own, Sfr = r.array("faceOwner"), r.array("faceAreas")
    for f in range(r.nFaces):
        s[own[f]] += Sfr[f]
    s[r.array("faceNeighbour")[0]] -= Sfr[0]
    assert np.abs(s).max() < 1e-15
    # and the oracle's own file reader sees the same mesh in the files we wrote
    assert np.array_equal(opoly.read_label_list(d / "owner"), ow) and np.allclose(opoly.read_points(d / "points"), pts)


def test_malformed_input_fails_loudly(tmp_path):
    with pytest.raises(FvkError):
        MeshDesc.from_polymesh(tmp_path / "nowhere")
    d, *_ = _write_prisms(tmp_path)
    (d / "neighbour").write_text((d / "neighbour").read_text().replace("format      ascii", "format      binary"))
    with pytest.raises(FvkError) as e:
        MeshDesc.from_polymesh(d)
    assert "binary" in str(e.value)
    d2 = tmp_path / "b"
    d2.mkdir()
    for f in ("points", "faces", "owner", "neighbour"):
        (d2 / f).write_text((d / f).read_text().replace("binary", "ascii"))
    (d2 / "boundary").write_text((d / "boundary").read_text().replace("startFace       5", "startFace       6"))
    with pytest.raises(FvkError) as e:
        MeshDesc.from_polymesh(d2)
    assert "does not start" in str(e.value)


@pytest.mark.gpu
def test_polyhedral_mesh_through_the_kernels(tmp_path):
    """A mesh read from polyMesh files (triangular and quadrilateral faces, cells with 5 faces, 4 boundary faces per cell:
    the overflow path of the brick plan's code lists) through the explicit operators, the assembly and SpMV: bit-identical
    to the Serial oracle on the same description."""
    import torch
    from foamadapter_b200 import la, ops
    from foamadapter_b200.mesh import UnstructuredMesh
    from oracle.cpu import Mesh as OMesh
    d, *_ = _write_prisms(tmp_path)
    desc = MeshDesc.from_polymesh(d)
    gm, om = UnstructuredMesh(desc), OMesh.from_desc(desc)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    rng = np.random.default_rng(3)
    phi, phib, flux = rng.uniform(1, 2, om.nC), rng.uniform(1, 2, om.nB), rng.uniform(-1, 1, om.nF)
    out = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    assert np.array_equal(ops.div(gm, dev(flux), dev(phi), dev(phib), out).cpu().numpy(), om.div(flux, phi, phib, 0))
    assert np.array_equal(ops.laplacian(gm, dev(phi), dev(phib), out).cpu().numpy(), om.laplacian(phi, phib))
    o3 = torch.zeros((om.nC, 3), dtype=torch.float64, device="cuda")
    assert np.array_equal(ops.grad(gm, dev(phi), dev(phib), o3).cpu().numpy(), om.grad(phi, phib))
    vals, x = rng.uniform(-1, 1, om.nnz), rng.uniform(-1, 1, om.nC)
    sp = la.SparsityPattern.readOrCreate(gm)
    assert np.array_equal(la.spmv(sp, dev(vals), dev(x)).cpu().numpy(), om.spmv(vals, x))
    assert np.array_equal(la.spmv_structured(gm, dev(vals), dev(x)).cpu().numpy(), om.spmv(vals, x))


def _field_file(path, cls, internal, patches):
    txt = ["FoamFile\n{\n    version 2.0;\n    format ascii;\n    class " + cls + ";\n    object " + path.name + ";\n}\n",
           "// a comment\ndimensions      [0 0 0 1 0 0 0];\n", "internalField   " + internal + ";\n", "boundaryField\n{\n"]
    for name, body in patches:
        txt.append(f"    {name}\n    {{\n{body}    }}\n")
    txt.append("}\n")
    path.write_text("\n".join(txt))
    return path


def test_field_file_reader(tmp_path):
    """fvk_fieldfile_read_*: the formats of the reference's test/setup_operator/0/ files."""
    from foamadapter_b200 import fvcc
    rng = np.random.default_rng(5)
    T = rng.uniform(1, 2, 25)
    p = _field_file(tmp_path / "T", "volScalarField", "nonuniform List<scalar> 25\n(\n" + "\n".join(repr(float(v)) for v in T) + "\n)",
                    [("movingWall", "        type            fixedValue;\n        value           uniform 10.5;\n"),
                     ("fixedWalls", "        type            zeroGradient;\n"), ("frontAndBack", "        type            empty;\n")])
    assert np.array_equal(fvcc.read_field_file(p, 25), T)            # repr round-trips doubles exactly
    assert fvcc.read_patch_conditions(p, ["movingWall", "fixedWalls", "frontAndBack"]) == [("fixedValue", 10.5), ("zeroGradient", None), ("empty", None)]
    U = rng.uniform(-1, 1, (25, 3))
    u = _field_file(tmp_path / "U", "volVectorField", "nonuniform List<vector> 25\n(\n" + "\n".join("(%r %r %r)" % tuple(map(float, v)) for v in U) + "\n)",
                    [("movingWall", "        type            fixedValue;\n        value           uniform (1 0 0);\n"),
                     ("fixedWalls", "        type            noSlip;\n")])
    assert np.array_equal(fvcc.read_field_file(u, 25), U)
    assert fvcc.read_patch_conditions(u, ["movingWall", "fixedWalls"]) == [("fixedValue", (1.0, 0.0, 0.0)), ("noSlip", None)]
    q = _field_file(tmp_path / "p", "volScalarField", "uniform 0", [("movingWall", "        type            zeroGradient;\n")])
    assert np.array_equal(fvcc.read_field_file(q, 25), np.zeros(25))
    v = _field_file(tmp_path / "U0", "volVectorField", "uniform (0 0.5 -2)", [])
    assert np.array_equal(fvcc.read_field_file(v, 4), np.tile([0.0, 0.5, -2.0], (4, 1)))
    with pytest.raises(FvkError):
        fvcc.read_field_file(p, 24)                                   # wrong cell count
    with pytest.raises(FvkError):
        fvcc.read_patch_conditions(p, ["noSuchPatch"])
    with pytest.raises(FvkError):
        fvcc.read_field_file(tmp_path / "missing", 25)


@pytest.mark.gpu
def test_reference_case_from_files_through_the_gpu(tmp_path):
    """The reference's test/setup_operator case rebuilt as FILES (polyMesh + 0/T written from the golden vectors), loaded
    by the product readers, through the GPU operators: the reference's own 16-digit divT / gradT come out."""
    import torch
    from pathlib import Path
    from foamadapter_b200 import fvcc
    from foamadapter_b200.mesh import UnstructuredMesh
    z = np.load(Path(__file__).parent / "golden" / "setup_operator.npz")
    d = _write_raw(tmp_path / "constant" / "polyMesh", z["points"], [list(f) for f in z["faces"]], z["owner"], z["neighbour"],
                   z["patch_names"], z["patch_types"], z["patch_size"])
    desc = MeshDesc.from_polymesh(d)
    gm = UnstructuredMesh(desc)
    assert gm.patch_names == desc.patch_names and (gm.nCells, gm.nInternalFaces, gm.nBoundaryFaces) == (25, 40, 20)
    (tmp_path / "0").mkdir()
    bodies = [(str(n), "        type            empty;\n" if str(t) == "empty" else "        type            zeroGradient;\n")
              for n, t in zip(z["patch_names"], z["patch_types"])]
    tf = _field_file(tmp_path / "0" / "T", "volScalarField",
                     "nonuniform List<scalar> 25\n(\n" + "\n".join(repr(float(v)) for v in z["field_T"]) + "\n)", bodies)
    T = fvcc.read_volume_field(gm, tf)
    assert T.ncomp == 1 and [k for k, _ in T.bcs] == ["zeroGradient"] * gm.nPatches
    phi = fvcc.SurfaceField(gm, "phi")
    phi.internal[: gm.nInternalFaces] = torch.as_tensor(z["field_phi"], device="cuda")
    div = torch.zeros(gm.nCells, dtype=torch.float64, device="cuda")
    fvcc.GaussGreenDiv(gm, "linear").div(div, phi, T, fvcc.Coeff(1.0))
    np.testing.assert_allclose(div.cpu().numpy(), z["field_divT_Serial"], rtol=5e-15, atol=0)
    grad = fvcc.GaussGreenGrad(gm).grad(T).cpu().numpy()
    np.testing.assert_allclose(grad[:, :2], z["field_gradT_Serial"][:, :2], rtol=0, atol=1e-12)


def test_field_file_round_trip(tmp_path):
    from foamadapter_b200 import fvcc
    rng = np.random.default_rng(8)
    T, U = rng.uniform(-1e3, 1e3, 37), rng.uniform(-1, 1, (37, 3)) * 10.0 ** rng.integers(-12, 12, (37, 3))
    names = ["inlet", "walls"]
    fvcc.write_field_file(tmp_path / "T", T, names, [("fixedValue", 0.1), ("zeroGradient", None)])
    fvcc.write_field_file(tmp_path / "U", U, names, [("fixedValue", (1.0, -2.5, 1e-9)), ("noSlip", None)])
    assert np.array_equal(fvcc.read_field_file(tmp_path / "T", 37), T) and np.array_equal(fvcc.read_field_file(tmp_path / "U", 37), U)
    assert fvcc.read_patch_conditions(tmp_path / "T", names) == [("fixedValue", 0.1), ("zeroGradient", None)]
    assert fvcc.read_patch_conditions(tmp_path / "U", names) == [("fixedValue", (1.0, -2.5, 1e-9)), ("noSlip", None)]
    assert np.array_equal(opoly.read_internal_field(tmp_path / "U"), U)   # the oracle's reader sees the same file


def test_reader_tolerates_formatting_variants(tmp_path):
    """Same mesh, three spellings: as written by fvk_polymesh_write, everything on one line with block comments, and with the
    `inGroups List<word> 1(wall);` / extra-key boundary dictionaries real OpenFOAM cases carry."""
    d, pts, faces, ow, ne = _write_prisms(tmp_path)
    ref = MeshDesc.from_polymesh(d)
    d2 = tmp_path / "compact"
    d2.mkdir()
    hdr = "FoamFile { version 2.0; format ascii; class x; object y; } /* block\ncomment */ "
    (d2 / "points").write_text(hdr + f"{len(pts)}(" + " ".join("(%r %r %r)" % tuple(map(float, p)) for p in pts) + ")")
    (d2 / "faces").write_text(hdr + f"{len(faces)} ( " + "  ".join(f"{len(f)}(" + " ".join(map(str, f)) + ")" for f in faces) + " )  // tail")
    (d2 / "owner").write_text(hdr + f"{len(ow)}(" + " ".join(map(str, ow)) + ")")
    (d2 / "neighbour").write_text(hdr + "1\n(\n1 // the only internal face\n)\n")
    (d2 / "boundary").write_text(hdr + """2
(
    walls0 { type wall; inGroups List<word> 1(wall); physicalType wall; nFaces 4; startFace 1; }
    walls1
    {
        type            patch;
        startFace       5;   // keys in any order
        nFaces          4;
    }
)
""")
    r = MeshDesc.from_polymesh(d2)
    assert r.patch_names == ref.patch_names and r.patch_types == ref.patch_types
    for name in ARRAYS:
        assert np.array_equal(r.array(name), ref.array(name)), name


def test_nonuniform_patch_values_are_read_per_face(tmp_path):
    """OpenFOAM writes `value nonuniform List<...>` into time directories; read_patch_conditions must take the real
    patch size and return the per-face values (a list whose entries are all equal collapses to a constant)."""
    import numpy as np
    from foamadapter_b200 import fvcc
    f = tmp_path / "T"
    f.write_text("""FoamFile { version 2.0; format ascii; class volScalarField; object T; }
dimensions [0 0 0 1 0 0 0];
internalField uniform 1;
boundaryField
{
    inlet { type fixedValue; value nonuniform List<scalar> 3 ( 1.5 2.5 3.5 ); }
    wall { type calculated; value nonuniform List<scalar> 2 ( 4 4 ); }
    outlet { type zeroGradient; }
}
""")
    bcs = fvcc.read_patch_conditions(f, ["inlet", "wall", "outlet"], [3, 2, 5])
    assert bcs[0][0] == "fixedValue" and np.array_equal(bcs[0][1], [1.5, 2.5, 3.5])
    assert bcs[1] == ("calculated", 4.0) and bcs[2] == ("zeroGradient", None)
    u = tmp_path / "U"
    u.write_text("""FoamFile { version 2.0; format ascii; class volVectorField; object U; }
internalField uniform (0 0 0);
boundaryField { inlet { type fixedValue; value nonuniform List<vector> 2 ( (1 0 0) (2 0 0.5) ); } }
""")
    (t, v), = fvcc.read_patch_conditions(u, ["inlet"], [2])
    assert t == "fixedValue" and np.array_equal(v, [[1, 0, 0], [2, 0, 0.5]])
    from foamadapter_b200._capi import FvkError
    with pytest.raises(FvkError):
        fvcc.read_patch_conditions(f, ["inlet"], [4])      # wrong patch size: loud
