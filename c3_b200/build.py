"""Build libc3b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

One object file per kernel family (c3_b200/csrc/*.cu), compiled in parallel and linked into
c3_b200/libc3b200.so; objects are cached under c3_b200/csrc/_obj (git-ignored) and rebuilt when the
source or any header changes."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libc3b200.so")
PUBLIC_HEADER = os.path.normpath(os.path.join(HERE, "..", "include", "c3b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libc3b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [PUBLIC_HEADER]


def _obj_path(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale_objects(force: bool):
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    out = []
    for src in sources():
        obj = _obj_path(src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            out.append(src)
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + _headers())


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile c3_b200/csrc/*.cu into c3_b200/libc3b200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", _obj_path(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    todo = _stale_objects(force)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        for src, res in pool.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB + ".tmp"] + [_obj_path(s) for s in sources()]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
