"""Build recipe for the in-tree CUDA library (sm_100a only).

`python -m distillclip_b200.build` or `__graft_entry__.build()` compiles every `.cu` under
`distillclip_b200/csrc/` into `distillclip_b200/csrc/libdistillclip_b200.so` with nvcc.  The library
exports the C ABI declared in `include/distillclip_b200.h`; Python reaches it with ctypes
(`distillclip_b200/_lib.py`).  The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(CSRC, "libdistillclip_b200.so")
OBJ = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"),
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths) -> str:
    """Content hash of the sources and flags, independent of where the tree lives (the GPU box unpacks it elsewhere)."""
    h = hashlib.sha256()
    for p in sorted(paths, key=os.path.basename):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode() + b"\0" + f.read())
    h.update(" ".join(x for x in NVCC_FLAGS if ROOT not in x).encode() + b" cudart=shared")
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed. Returns the library path."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "distillclip_b200.h"))
    stamp = os.path.join(OBJ, "stamp.txt")
    digest = _digest(deps)

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest
    if not force and fresh():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    # one builder at a time (torchrun starts every rank at once); the others wait on the lock and find the result
    import fcntl
    with open(os.path.join(OBJ, "build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return LIB
        return _build_locked(srcs, stamp, digest, verbose)


def _build_locked(srcs, stamp, digest, verbose) -> str:
    exe = nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [exe, *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-6000:]}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB + f".tmp{os.getpid()}"
    # shared cudart: the process already holds torch's libcudart.so.12 (same SONAME); the rpath covers a bare ctypes load.
    # A static cudart would drag every runtime entry point's name into the shipped .so.
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(exe)), "lib64")
    cmd = [exe, "-shared", "--cudart", "shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-Xlinker", f"-rpath={cuda_lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    os.replace(tmp, LIB)                    # atomic: a concurrent loader sees the old or the new library, never half of one
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
