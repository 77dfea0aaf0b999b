"""TEST INFRASTRUCTURE.  Builds tests/emu/_build/libhuffb200_emu.so: the product's sources (translated by translate.py)
compiled with g++ against the CPU execution model.  Same C ABI as libhuffb200.so; loaded only by the tests
(HUFFB200_SO=...), never by the product.

    python tests/emu/build.py [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import translate  # noqa: E402

OUT = os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libhuffb200_emu.so")
FLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-DHB_EMU=1", "-Wall", "-Wno-unknown-pragmas",
         "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-sign-compare",
         "-fno-strict-aliasing"]


def build(force: bool = False) -> str:
    files = translate.translate_tree(OUT)
    deps = files + [os.path.join(HERE, "hb_emu.cpp"), os.path.join(HERE, "include", "hb_emu.h"),
                    os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "include", "nccl.h"), os.path.join(HERE, "nccl_emu.cpp"),
                    os.path.join(ROOT, "include", "huffb200.h"), os.path.abspath(__file__),
                    os.path.join(HERE, "translate.py")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    nccl = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-I", os.path.join(HERE, "include"),
                           "-o", os.path.join(OUT, "libnccl_emu.so"), os.path.join(HERE, "nccl_emu.cpp")],
                          capture_output=True, text=True)
    if nccl.returncode:
        sys.stderr.write(nccl.stdout + nccl.stderr)
        raise RuntimeError("g++ failed building libnccl_emu.so")
    srcs = [f for f in files if f.endswith(".cpp")] + [os.path.join(HERE, "hb_emu.cpp")]
    cmd = ["g++"] + FLAGS + ["-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "include"), "-I", OUT,
                             "-o", SO] + srcs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building libhuffb200_emu.so")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
