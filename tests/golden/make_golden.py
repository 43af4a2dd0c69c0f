"""Regenerate tests/golden/*.npz from the fixtures committed in the reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The .npz files hold DATA (mesh topology, points, 16-digit golden fields), no reference source.
Sources:
  test/setup_*/constant/polyMesh/{points,faces,owner,neighbour,boundary}      (blockMesh output)
  test/setup_operator/0/{T,phi,divT_Serial,divT_OpenMP,gradT_Serial,gradT_OpenMP,ofDivT,ofGradT}
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import polymesh  # noqa: E402

REF = Path("/root/reference/test")
OUT = Path(__file__).resolve().parent

# (case, (nx,ny,nz), (lx,ly,lz) incl. `scale`, patches as (name, sides, isEmpty)) read off each
# case's system/blockMeshDict + simulationParameters
X0, X1, Y0, Y1, Z0, Z1 = range(6)
CASES = {
    "setup_operator": ((5, 5, 1), (0.1, 0.1, 0.01),
                       [("fixedWalls", [Y1, X0, X1, Y0], 0), ("frontAndBack", [Z0, Z1], 1)]),
    "setup_stencil3D": ((3, 3, 3), None, [("fixedWalls", [Y1, Y0, Z0, Z1], 0), ("inlet", [X1], 0), ("outlet", [X0], 0)]),
    "setup_pressureVelocityCoupling": ((3, 3, 3), None, None),
    "setup_advection": ((50, 50, 1), None, None),
    "setup_unstructuredMesh": ((3, 3, 1), None, None),
    "setup_compatibility": ((3, 3, 1), None, None),
}


def main():
    for case, (dims, _, _) in CASES.items():
        pm = REF / case / "constant" / "polyMesh"
        points = polymesh.read_points(pm / "points")
        faces = polymesh.read_faces(pm / "faces")
        owner = polymesh.read_label_list(pm / "owner")
        neighbour = polymesh.read_label_list(pm / "neighbour")
        patches = polymesh.read_boundary(pm / "boundary")
        assert all(len(f) == 4 for f in faces)
        out = dict(
            dims=np.array(dims, dtype=np.int32), points=points,
            faces=np.array(faces, dtype=np.int32), owner=owner, neighbour=neighbour,
            patch_names=np.array([p["name"] for p in patches]),
            patch_types=np.array([p["type"] for p in patches]),
            patch_start=np.array([p["startFace"] for p in patches], dtype=np.int32),
            patch_size=np.array([p["nFaces"] for p in patches], dtype=np.int32),
        )
        if case == "setup_operator":
            for name in ("T", "phi", "divT_Serial", "divT_OpenMP", "gradT_Serial", "gradT_OpenMP",
                         "ofDivT", "ofGradT"):
                out["field_" + name] = polymesh.read_internal_field(REF / case / "0" / name)
        np.savez_compressed(OUT / f"{case}.npz", **out)
        print(case, dims, "cells", owner.max() + 1, "faces", len(faces), "nI", len(neighbour),
              [(p["name"], p["type"], p["nFaces"]) for p in patches])


if __name__ == "__main__":
    main()
