"""In-tree build of the CUDA shared library (sm_100a only).

Every csrc/*.cu is compiled to its own object (in parallel, cached by a content hash of the
file, the headers and the flags) and the objects are linked into sssp_b200/_lib/libmrmp_b200.so.
No relocatable device code: kernels never call across translation units."""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "_lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libmrmp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # the reference never fuses a*b+c; parity needs un-fused FP64
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libmrmp_b200.so")


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(_PKG, "..", "include", "*.h")))


def _hash_files(paths, extra=()) -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for e in extra:
        h.update(str(e).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _digest(extra=()) -> str:
    """Content hash of everything the library is built from (mtimes do not survive the
    snapshot to the GPU box, content does)."""
    return _hash_files(sources() + _headers(), extra)


STAMP_PATH = os.path.join(LIB_DIR, "build.stamp")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _digest()


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile csrc/*.cu into sssp_b200/_lib/libmrmp_b200.so; returns the path.  `defines`
    (-DNAME=VALUE strings) and `out` build an experimental variant beside the product library."""
    out = out or LIB_PATH
    if not force and not defines and out == LIB_PATH and not _stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = _headers()
    env = dict(os.environ)

    def one(src):
        key = _hash_files([src] + hdrs, defines)[:24]
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + "." + key + ".o")
        if force and not defines and os.path.exists(obj):
            os.remove(obj)
        if not os.path.exists(obj):
            cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", "-o", obj + ".tmp", src]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
            subprocess.check_call(cmd, env=env)
            os.replace(obj + ".tmp", obj)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, sources()))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                           "-o", out + ".tmp", *objs], env=env)
    os.replace(out + ".tmp", out)
    # drop objects of earlier source versions
    keep = set(objs)
    for p in glob.glob(os.path.join(OBJ_DIR, "*.o")):
        if p not in keep and not defines:
            os.remove(p)
    if out == LIB_PATH:
        with open(STAMP_PATH, "w") as f:
            f.write(_digest())
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
