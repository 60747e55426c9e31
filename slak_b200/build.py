"""In-tree build of libslak_b200.so (nvcc, sm_100a only) and of the oracle checkers.

`python -m slak_b200.build` or `__graft_entry__.build()` call `build_all()`.
nvcc cross-compiles without a GPU; the .so files are git-ignored but travel to the
GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "slak_b200", "csrc")
LIB = os.path.join(ROOT, "slak_b200", "libslak_b200.so")
OBJDIR = os.path.join(ROOT, "build", "obj")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "slak_b200.h"))
    return hdrs


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ and link libslak_b200.so. Returns the .so path."""
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    stamp_file = os.path.join(OBJDIR, "stamp.txt")
    stamp = _stamp(srcs + _deps())
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file):
        if open(stamp_file).read().strip() == stamp:
            return LIB
    hdr_stamp = _stamp(_deps())

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        ostamp = obj + ".stamp"
        s = _stamp([src]) + hdr_stamp
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == s:
            return obj
        cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(ostamp, "w") as f:
            f.write(s)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


def build_oracle() -> None:
    """Build the C restatement (oracle/liboracle.so) and, when the reference checkout is
    present, the reference's own host code (oracle/_ref/libslak_ref.so)."""
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force=force, verbose=verbose)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
