"""TEST INFRASTRUCTURE.  Memory-safety contracts of the device-buffer API, checked with GUARD PAGES under the CPU model
(compute-sanitizer is not available on the GPU pool): every buffer handed to the library ends exactly where the header
says the library may stop (include/huffb200.h), and the next page is PROT_NONE -- one byte too far is a SIGSEGV.

    letters   readable [d_data, d_data + n)                       (d_data 16-byte aligned)
    stream    written  [d_out, d_out + ceil((start_bit+bits)/8) rounded up to 4)
    stream    readable [d_comp, d_comp + comp_len rounded up to 16)
    letters   written  [d_out, d_out + out_cap), any alignment

usage: HB_EMU=1 HUFFB200_SO=.../libhuffb200_emu.so python tests/emu/guard_check.py"""
import ctypes
import mmap
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from huff_encoding_b200 import datagen as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.emu.model_engine import ModelEngine  # noqa: E402

PAGE = 4096
libc = ctypes.CDLL(None, use_errno=True)
_keep = []


def guarded(nbytes: int, end_round: int = 1, start_align: int = 1, fill: int = 0x77) -> torch.Tensor:
    """uint8 tensor of `nbytes` whose (rounded-up-to-end_round) end is the start of a PROT_NONE page."""
    span = (nbytes + end_round + start_align + PAGE - 1) // PAGE * PAGE + PAGE
    m = mmap.mmap(-1, span + PAGE)
    base = ctypes.addressof(ctypes.c_char.from_buffer(m))
    if libc.mprotect(ctypes.c_void_p(base + span), PAGE, 0) != 0:
        raise OSError(ctypes.get_errno(), "mprotect")
    end = base + span
    start = end - (nbytes + end_round - 1) // end_round * end_round
    assert start % start_align == 0, (start, start_align, "choose sizes so that the start is aligned")
    arr = np.frombuffer(m, dtype=np.uint8)
    arr[:span] = fill
    _keep.append((m, arr))
    t = torch.from_numpy(arr[start - base: start - base + nbytes])
    assert t.data_ptr() == start
    return t


def check(eng, data: np.ndarray, start_bit: int, out_misalign: int):
    n = data.size
    comp, pad, otree = O.compress(data)
    bits = comp.size * 8 - pad
    tree = eng.tree_from_weights(np.bincount(data, minlength=256))
    # letters: exactly n readable bytes (start 16-byte aligned => n is a multiple of 16 here)
    d = guarded(n, 1, 16)
    d.copy_(torch.from_numpy(data))
    h = eng.histogram(d).numpy()
    assert np.array_equal(h, np.bincount(data, minlength=256))
    # stream out: ceil((start_bit + bits) / 8) rounded up to 4, not a byte more
    clen = (start_bit + bits + 7) // 8
    out = guarded((clen + 3) // 4 * 4, 1, 4)
    eng.histogram(d)
    eng.encode(d, tree, out, start_bit=start_bit)
    got = np.unpackbits(out[:clen].numpy())[start_bit: start_bit + bits]
    assert np.array_equal(got, np.unpackbits(comp)[:bits])
    # stream in: readable up to the length rounded up to 16; letters out: exactly n bytes at an odd address
    src = guarded(clen, 16, 16)
    src.copy_(out[:clen])
    dst_all = guarded(n + out_misalign, 1, 1)
    dst = dst_all[out_misalign:]
    e, x, cnt = eng.decode_shard(src, start_bit + bits, start_bit, start_bit + bits, 0, tree, start_bit, dst)
    assert cnt == n and np.array_equal(dst.numpy(), data), "decode_shard"
    path = eng.ctx.last_decode_path()[0]
    # the two-pass decoder on the same buffers (count + write)
    e, x, cnt = eng.decode_count(src, start_bit + bits, start_bit, start_bit + bits, 0, tree, entry_bit=start_bit)
    dst.zero_()
    eng.decode_write(dst)
    assert cnt == n and np.array_equal(dst.numpy(), data), "decode_count + decode_write"
    if start_bit == 0:
        # whole-stream entry points
        o2 = guarded((comp.size + 3) // 4 * 4, 1, 4)
        _, clen2, pad2, t2 = eng.compress(d, out=o2)
        assert clen2 == comp.size and pad2 == pad and np.array_equal(o2[:clen2].numpy(), comp)
        src2 = guarded((comp.size + 15) // 16 * 16, 1, 16)
        src2[: comp.size].copy_(torch.from_numpy(comp))
        dst.zero_()
        _, m = eng.decompress(src2, comp.size, pad, t2, out=dst)
        assert m == n and np.array_equal(dst.numpy(), data), "decompress"
    return path


if __name__ == "__main__":
    eng = ModelEngine()
    rng = np.random.default_rng(3)
    cases = []
    for gen, n in (("zipf", 1 << 19), ("english", 3 * 33792 * 8 // 4 // 16 * 16), ("uniform", 1 << 17), ("zipf", 4096), ("english", 16)):
        cases.append((gen, getattr(G, gen)(n, seed=n)))
    w = G.fibonacci_weights(n_fib=24, n_ones=40)
    fib = G.from_weights_runs(w)
    cases.append(("fibonacci", fib[: fib.size // 16 * 16][:600_000 // 16 * 16]))
    cases.append(("two letters", rng.choice(np.array([65, 66], np.uint8), size=1 << 16)))
    paths = {}
    for name, data in cases:
        for sb in (0, 5, 31):
            for mis in (0, 1, 17):
                p = check(eng, data, sb, mis)
                paths[p] = paths.get(p, 0) + 1
    assert paths.get(1, 0) > 0 and paths.get(0, 0) > 0, paths       # both the fused and the two-pass / fixed decoders were hit
    print("guard pages: ok", len(cases), "inputs x 3 start bits x 3 output alignments; decoder paths", paths)
