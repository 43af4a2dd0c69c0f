"""The C++ host layer (include/NeoN, include/FoamAdapter -- the reference's class names over the C ABI): it must compile
with a plain g++ against libfvk.so (CPU test), fail loudly without a GPU, and on the GPU pass the reference's known-answer
cases and reproduce the oracle's neoIcoFoam run from examples/neoIcoFoam/neoIcoFoam.cpp."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from foamadapter_b200.build import LIB, build_lib

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "tests" / "cpp" / "build"


def _compile(src: Path, name: str) -> Path:
    build_lib()
    OUT.mkdir(parents=True, exist_ok=True)
    exe = OUT / name
    deps = [src] + list((ROOT / "include").rglob("*.h*")) + [LIB]
    if not exe.exists() or any(d.stat().st_mtime > exe.stat().st_mtime for d in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), f"-L{LIB.parent}", "-lfvk",
               f"-Wl,-rpath,{LIB.parent}", "-o", str(exe)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_layer_compiles_and_has_no_cpu_fallback():
    api = _compile(ROOT / "tests" / "cpp" / "test_host_api.cpp", "test_host_api")
    ico = _compile(ROOT / "examples" / "neoIcoFoam" / "neoIcoFoam.cpp", "neoIcoFoam")
    if not torch.cuda.is_available():
        for exe in (api, ico):
            r = subprocess.run([str(exe)], capture_output=True, text=True)
            assert r.returncode != 0 and "libfvk" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_cpp_known_answers_on_gpu():
    exe = _compile(ROOT / "tests" / "cpp" / "test_host_api.cpp", "test_host_api")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "host api ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("three_d", [False, True])
def test_cpp_neoicofoam_matches_oracle(tmp_path, three_d):
    from foamadapter_b200 import piso
    from oracle.cpu import Mesh as OMesh
    from oracle.piso import IcoFoamOracle
    exe = _compile(ROOT / "examples" / "neoIcoFoam" / "neoIcoFoam.cpp", "neoIcoFoam")
    n, steps = (8, 3) if three_d else (12, 3)
    dump = tmp_path / "fields.bin"
    r = subprocess.run([str(exe), str(n), str(steps)] + (["--3d"] if three_d else []), capture_output=True, text=True,
                       env=dict(os.environ, NEOICOFOAM_DUMP=str(dump)))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("[NeoN] Solving for p") == 2 * steps
    om = OMesh.from_desc(piso.cavity_desc(n, three_d))
    o = IcoFoamOracle(om, nu=0.01, dt=1e-4 * 20.0 / n)
    for _ in range(steps):
        o.step()
    raw = np.fromfile(dump, dtype=np.float64)
    U, p = raw[: 3 * om.nC].reshape(-1, 3), raw[3 * om.nC:]
    assert np.abs(U - o.U).max() <= 1e-8 * np.abs(o.U).max()
    assert np.abs(p - o.p).max() <= 1e-7 * np.abs(o.p).max()
