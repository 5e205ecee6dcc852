"""In-tree build of the two shared libraries with plain nvcc/g++:

  librheo_b200.so   the product: the sm_100a CUDA stress step behind include/rheo_gpu.h (csrc/gpu/*.cu) — needs a CUDA device
  librheo_host.so   host services (csrc/host/*.cpp): blockMesh-lite mesh generator, decomposePar-style decomposition,
                    synthetic fields, OpenFOAM file formats (include/rheo_mesh.h, include/rheo_io.h) — no CUDA in it, so the
                    CPU reference arm of bench.py and the oracle-only tests never map the CUDA library

Both are built next to this file so that they travel to the GPU box with the repo snapshot (git-ignored, not
gpurun-ignored).  No JIT cache, no torch.utils.cpp_extension: the C-ABI has no torch types.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "rheotool_b200"
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "librheo_b200.so"
HOST_LIB = PKG / "librheo_host.so"

NVCC = os.environ.get("RHEO_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"   # the image's default CXX (/opt/gcc) has no OpenMP spec; system g++ 13 is complete

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX,
] + (["-Xptxas", "-v"] if os.environ.get("RHEO_PTXAS_V") else []) + os.environ.get("RHEO_NVCC_EXTRA", "").split()
CXX_FLAGS = ["-O2", "-fPIC", "-std=c++17", "-Wall"]


def _newer(src: Path, dst: Path, deps: list[Path]) -> bool:
    if not dst.exists():
        return True
    t = dst.stat().st_mtime
    return any(p.stat().st_mtime > t for p in [src, *deps] if p.exists())


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {cmd[0]} exit {r.returncode}")
    if os.environ.get("RHEO_PTXAS_V"):
        sys.stderr.write(r.stderr)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every source under csrc/ and link librheo_b200.so and librheo_host.so.  Returns the path of the CUDA library."""
    if not Path(NVCC).exists():
        raise RuntimeError(f"nvcc not found at {NVCC}; the B200 stress-step library cannot be built")
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = list((ROOT / "include").glob("*.h")) + list(CSRC.rglob("*.hpp")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.inl"))
    inc = ["-I", str(ROOT / "include"), "-I", str(CSRC / "host"), "-I", str(CSRC / "gpu")]
    objs = []
    for src in sorted(CSRC.rglob("*.cpp")) + sorted(CSRC.rglob("*.cu")):
        if "plugin" in src.parts and src.name.startswith("of90_"):
            continue   # OpenFOAM-9 shim: only compilable inside an OpenFOAM environment (wmake)
        obj = OBJ / (src.relative_to(CSRC).as_posix().replace("/", "__") + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            if src.suffix == ".cu":
                cmd = [NVCC, *NVCC_FLAGS, *inc, "-c", str(src), "-o", str(obj)]
            else:
                cmd = [HOST_CXX, *CXX_FLAGS, *inc, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd))
            _run(cmd)
    gpu_objs = [o for o in objs if o.name.startswith("gpu__")]
    host_objs = [o for o in objs if not o.name.startswith("gpu__")]
    if force or not HOST_LIB.exists() or any(o.stat().st_mtime > HOST_LIB.stat().st_mtime for o in host_objs):
        cmd = [HOST_CXX, "-shared", "-o", str(HOST_LIB), *map(str, host_objs), "-lpthread", "-lz"]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    if force or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in gpu_objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX, "-o", str(LIB), *map(str, gpu_objs), "-lcudart", "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
