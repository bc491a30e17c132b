"""Builds routeformer_b200/_lib/librouteformer_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree).

`python -m routeformer_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "librouteformer_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return cand


def _digest(paths) -> str:
    """Content hash of the sources and flags.  File NAMES only, never absolute paths: the tree is copied to other locations
    (the GPU box runs it from a scratch directory) and must not look stale there."""
    h = hashlib.sha256()
    for p in sorted(paths, key=os.path.basename):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = (sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.inc"))) +
               [os.path.join(HERE, "..", "include", "routeformer_b200.h")])
    os.makedirs(os.path.join(OUT_DIR, "obj"), exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.sha256")
    digest = _digest(sources + headers)

    def fresh() -> bool:
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest

    if not force and fresh():
        return LIB
    # one builder at a time (torchrun starts one process per GPU): the others wait on the lock and find a fresh library
    import fcntl

    with open(os.path.join(OUT_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return LIB
            return _build_locked(sources, digest, stamp, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(sources, digest: str, stamp: str, verbose: bool) -> str:
    nvcc = _nvcc()
    pid = os.getpid()

    def compile_one(src):
        obj = os.path.join(OUT_DIR, "obj", f"{os.path.basename(src)[:-3]}.{pid}.o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp_lib = f"{LIB}.{pid}.tmp"
    r = subprocess.run([nvcc, "-shared", "-o", tmp_lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    for o in objs:
        os.remove(o)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp_lib, LIB)  # atomic: a process that is loading the old file keeps its mapping
    with open(stamp + ".tmp", "w") as f:
        f.write(digest)
    os.replace(stamp + ".tmp", stamp)
    if verbose:
        print(f"built {LIB} from {len(sources)} sources", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
