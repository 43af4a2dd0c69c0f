"""Build libfvk.so (sm_100a) in-tree with nvcc. No JIT cache, no torch extension machinery: the
library is a plain C-ABI shared object (include/fvk.h) that travels with the source tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libfvk.so"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    # bit-reproducibility vs. the reference's Serial executor: no FMA contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fopenmp,-O3,-ffp-contract=off",
    f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.hpp")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))
    me = Path(__file__)
    jobs = []
    objs = []
    for src in sources():
        obj = objdir / (src.name + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, me] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-Xcompiler", "-fopenmp", "-lgomp",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
