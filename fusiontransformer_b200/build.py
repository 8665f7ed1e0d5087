"""In-tree build of libft3d.so (sm_100a only).

``python -m fusiontransformer_b200.build`` or ``build_library()``: every ``csrc/*.cu`` is compiled with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` into ``build/*.o`` and linked into
``fusiontransformer_b200/libft3d.so`` (static cudart, no torch types anywhere in the ABI).  nvcc
cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "ft3d"
LIB = PKG / "libft3d.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-DFT3D_BUILD",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(path: Path, extra: str) -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    h.update(path.read_bytes())
    for dep in sorted(list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(dep.read_bytes())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha")
    dig = _digest(src, " ".join(NVCC_FLAGS))
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src.name, r.stdout, r.stderr))
    if verbose and r.stderr.strip():
        print(r.stderr, file=sys.stderr)
    stamp.write_text(dig)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile and link libft3d.so; returns its path.  Incremental unless ``force``."""
    OBJ.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError("no CUDA sources under %s" % CSRC)
    if force:
        for f in OBJ.glob("*.sha"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-o", str(LIB), *map(str, objs), "-cudart", "static", "-ldl", "-lpthread", "-lrt"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
