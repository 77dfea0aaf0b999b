"""Builds libhuffb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m huff_encoding_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libhuffb200.so")
SOURCES = ["hb_api.cu", "hb_tree.cpp"]
INCLUDE = os.path.join("..", "..", "include", "huffb200.h")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--use_fast_math", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-shared", "-cudart", "static"]


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    # every source and header under csrc/ (a stale .so must never ship because a header was left off a list)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp"))]
    deps += [os.path.join(CSRC, INCLUDE), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: str | None = None) -> str:
    """defines / out: experiment builds (e.g. defines=["HB_LOOKBACK_BITS=128"], out="libhuffb200_lb128.so"),
    selected at run time with HUFFB200_SO=<path>."""
    target = os.path.join(HERE, out) if out else SO
    if not force and not defines and not out and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + [f"-D{d}" for d in (defines or [])] + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", target] + [os.path.join(CSRC, f) for f in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libhuffb200.so")
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
