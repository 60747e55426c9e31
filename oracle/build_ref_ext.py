"""oracle/build_ref_ext.py -- TEST / BASELINE INFRASTRUCTURE (never imported by slak_b200/).

Builds the reference's OWN CUDA operator -- the vendored MegEngine-CUTLASS example-19 torch
extension `_depthwise_conv2d_implicit_gemm_C` -- for sm_100a, from its sources WHERE THEY LIE under
/root/reference (nothing is copied), into oracle/_ref/ (git-ignored, travels to the GPU box with
the gpurun snapshot).  It is north_star's stated comparison target ("the reference CUTLASS-ext
build on the same B200 box") and a second checker for the GPU parity tests.

Recipe = the reference's setup.py (cutlass/examples/19_large_depthwise_conv2d_torch_extension/
setup.py:9-35: the seven sources, the five include dirs, '-g') with the arch pinned to sm_100a.
Each .cu takes 3.5-5 minutes of nvcc; ninja runs them in parallel.

    python oracle/build_ref_ext.py            # build if missing
    python oracle/build_ref_ext.py --force
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SLAK_REFERENCE", "/root/reference")
EXT_DIR = os.path.join(REF, "cutlass", "examples", "19_large_depthwise_conv2d_torch_extension")
CUTLASS_ROOT = os.path.join(REF, "cutlass")
OUT = os.path.join(HERE, "_ref", "ext")
NAME = "_depthwise_conv2d_implicit_gemm_C"
SOURCES = ["frontend.cpp", "forward_fp32.cu", "backward_data_fp32.cu", "backward_filter_fp32.cu",
           "forward_fp16.cu", "backward_data_fp16.cu", "backward_filter_fp16.cu"]


def so_path() -> str:
    return os.path.join(OUT, NAME + ".so")


def build(force: bool = False, verbose: bool = False) -> str | None:
    """Returns the path of the built extension, or None when the reference checkout is absent
    (GPU box: the prebuilt file shipped with the snapshot is used as is)."""
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.isdir(EXT_DIR):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "7")
    from torch.utils import cpp_extension
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(EXT_DIR, s) for s in SOURCES],
        extra_include_paths=[EXT_DIR, os.path.join(CUTLASS_ROOT, "include"),
                             os.path.join(CUTLASS_ROOT, "tools", "library", "include"),
                             os.path.join(CUTLASS_ROOT, "tools", "util", "include"),
                             os.path.join(CUTLASS_ROOT, "examples", "common")],
        extra_cflags=["-g"], extra_cuda_cflags=["-g"],
        build_directory=OUT, verbose=verbose, is_python_module=False)
    return so_path() if os.path.exists(so_path()) else None


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
