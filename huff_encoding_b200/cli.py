"""`huff`-compatible file front end: python -m huff_encoding_b200.cli [-d] [-t] [-r] [-n]
[-b SIZE] SRC_FILE [DST_FILE]

Mirrors the reference binary's interface and file format (paths relative to /root/reference/huff):
  flags and defaults           res/cli.yml:1-39
  path / extension rules       src/cli.rs:24-77   (".hff" is appended on compress, required on decompress)
  block-size units             src/cli.rs:79-114
  .hff layout                  src/comp.rs:32-74  = [(tree_pad << 4) + data_pad][BE u32 tree bytes][tree][data],
                               i.e. exactly CompressData::to_bytes (huff_coding/src/comp.rs:279-300)
  tree from the file's bytes   src/comp.rs:161-172: ByteWeights::threaded_from_bytes(block, 12) folded with `+=`
                               (reproduced, including that iterator's wrap-around double count of byte 0)

Deliberate deviation: a file larger than the block size is still written as ONE gap-free stream.  The reference
stitches blocks with a wrong shift (src/comp.rs:199, src/utils.rs:8), so its multi-block
output is not a parity target; single-block files (the default block is 2 GB) are byte-identical in layout.
All counting / packing / decoding runs on the GPU through libhuffb200 (no CPU fallback).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

EXTENSION = "hff"
_UNITS = {"": 1, "k": 1_000, "m": 1_000_000, "g": 1_000_000_000, "ki": 1024, "mi": 1_048_576, "gi": 1_073_741_824}


def parse_block_size(text: str) -> int:
    """src/cli.rs:79-114"""
    digits = "".join(c for c in text if c.isdigit())
    unit = text[len(digits):].lower()
    if not digits or unit not in _UNITS:
        raise SystemExit("Invalid block size")
    return int(digits) * _UNITS[unit]


def _paths(src: str, dst: str, decompress: bool) -> tuple[str, str]:
    """src/cli.rs:24-77"""
    if dst == "./SRC_FILE.hff":
        dst = os.path.join(".", os.path.basename(src))
    if os.path.isdir(src):
        raise SystemExit(f"{src!r} is a directory")
    if not decompress:
        return src, dst + "." + EXTENSION
    if not src.endswith("." + EXTENSION):
        raise SystemExit(f"Unrecognized file format, expected {EXTENSION}")
    if os.path.normpath(dst) == os.path.normpath(os.path.join(".", src)):
        dst = dst[: -len(EXTENSION) - 1]
    if os.path.isdir(dst):
        raise SystemExit(f"Destination {dst!r} is a directory")
    return src, dst


def _read_pinned(path: str):
    """The whole file in page-locked memory (hb_host_alloc), read in 64 MiB pieces straight into it: the H2D copy that
    follows runs at PCIe speed without a second host copy."""
    from . import api
    n = os.path.getsize(path)
    buf = api.PinnedBuffer(n)
    view = memoryview(buf.array) if n else memoryview(b"")
    with open(path, "rb", buffering=0) as f:
        got = 0
        while got < n:
            k = f.readinto(view[got: min(n, got + (64 << 20))])
            if not k:
                break
            got += k
    if got != n:
        raise SystemExit(f"short read on {path!r}")
    return buf


def compress_file(src: str, dst: str, block_size: int) -> None:
    """src/comp.rs:32-74.  Source and stream are staged in pinned host memory; the container header is written first and the
    stream follows from the pinned buffer without being copied together with it."""
    from . import api
    with _read_pinned(src) as inp:
        data = inp.array
        if data.size == 0:
            raise SystemExit("provided empty weights")               # tree_inner.rs:283-285
        bw = api.ByteWeights()
        for s in range(0, data.size, block_size):                    # src/comp.rs:161-172
            bw += api.ByteWeights.threaded_from_bytes(data[s:s + block_size], 12)
        tree = api.HuffTree.from_weights(bw)
        with api.PinnedBuffer(data.size + data.size // 4 + 4096) as outp:
            cd = api.compress_with_tree(data, tree, out=outp.array)
            header = api.CompressData(np.zeros(1, np.uint8), cd.padding_bits(), tree).to_bytes()[:-1]
            with open(dst, "wb") as f:
                f.write(header)                                          # (tree_pad << 4) + data_pad, BE u32 tree bytes, tree
                f.write(memoryview(cd.comp_bytes()))                     # the stream, straight from pinned memory


def decompress_file(src: str, dst: str) -> None:
    """src/comp.rs:79-157"""
    from . import api
    with _read_pinned(src) as inp:
        blob = inp.array
        if blob.size < 5:
            raise SystemExit("MissingHeaderInfo")
        cd = api.CompressData.try_from_bytes(blob, copy=False)       # comp_bytes stay a view of the pinned file image
        bound = cd.comp_bytes().size * 8 // max(1, min(len(c) for c in cd.huff_tree().read_codes().values())) + 64
        with api.PinnedBuffer(bound) as outp:
            back = api.decompress(cd, out=outp.array)
            with open(dst, "wb") as f:
                f.write(memoryview(back))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="huff", description="Compress/decompress SRC_FILE into DST_FILE.hff (compress by default)")
    ap.add_argument("-d", "--decompress", action="store_true", help="Decompresses the hff SRC_FILE into DST_FILE")
    ap.add_argument("-t", "--time", action="store_true", help="Prints how long it took to finish")
    ap.add_argument("-r", "--replace", action="store_true", help="Deletes SRC_FILE upon completion")
    ap.add_argument("-n", "--noask", action="store_true", help="Omits asking if existing DST_FILE should be replaced")
    ap.add_argument("-b", "--block-size", default="2G", metavar="SIZE")
    ap.add_argument("SRC_FILE")
    ap.add_argument("DST_FILE", nargs="?", default="./SRC_FILE.hff")
    a = ap.parse_args(argv)
    start = time.perf_counter()
    block = parse_block_size(a.block_size)
    src, dst = _paths(a.SRC_FILE, a.DST_FILE, a.decompress)
    if os.path.exists(dst) and not a.noask:                       # src/cli.rs:116-130
        ans = input(f"{dst!r} already exists, do you want to replace it? [Y/N]: ")
        if not ans.startswith("y"):
            return 0
    if a.decompress:
        decompress_file(src, dst)
    else:
        compress_file(src, dst, block)
    if a.replace:
        os.remove(src)
    if a.time:
        print(f"{time.perf_counter() - start:.6f}s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
