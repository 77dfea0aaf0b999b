"""TEST INFRASTRUCTURE.  A rehearsal of bench.py's control flow (N = 1) without a GPU: the library is the CPU model
(HUFFB200_SO), the few torch.cuda entry points bench.py touches are stood in for, sizes are shrunk.  What it checks is that
the script runs from argument parsing to the one JSON line -- every key access, every branch of the secondary configs -- not
a single number: times measured here are meaningless and the line says so ("shrink").  Under torchrun (N > 1) the process
group is gloo and the library's communicator is the NCCL model of tests/emu (shared memory between the rank processes).
usage: HB_EMU=1 HUFFB200_SO=.../libhuffb200_emu.so python tests/emu/bench_dryrun.py [bench.py arguments]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

assert os.environ.get("HB_EMU") == "1", "rehearsal only: needs the CPU model of the library"


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _StreamCtx:
    def __init__(self, stream):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Stream:
    def synchronize(self):
        pass


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
torch.cuda.Event = _Event
torch.cuda.stream = _StreamCtx
_empty = torch.empty
torch.empty = lambda *a, pin_memory=False, **k: _empty(*a, **k)

import huff_encoding_b200.engine as engine_mod  # noqa: E402
from tests.emu.model_engine import ModelEngine  # noqa: E402


class _Engine(ModelEngine):
    def __init__(self, device=None):
        super().__init__()
        self._stream = _Stream()

    def _event(self):
        return _Event()


engine_mod.Engine = _Engine

import huff_encoding_b200.sharded as sharded_mod  # noqa: E402


def _codec_event(self):
    ev = _Event()
    ev.record()
    return ev


sharded_mod.ShardedCodec._event = _codec_event

# fault injection (tests of bench.py's containment): HB_DRYRUN_FAIL="<rank>:<letters>[:before]" makes that rank's first round
# trip over an input of that many letters raise AFTER its collective, like a status returned by the decoder -- or, with
# ":before", before it, which leaves the other ranks waiting in a collective nobody will complete
_fail = os.environ.get("HB_DRYRUN_FAIL")
if _fail:
    _fail_rank, _fail_size = (int(x) for x in _fail.split(":")[:2])
    _fail_before = _fail.endswith(":before")     # raise BEFORE the collective: the other ranks are left waiting (watchdog)
    _real_round_trip = sharded_mod.ShardedCodec.round_trip
    _fired = []

    def _round_trip(self, data, comp_buf, out_buf, want_events=False):
        hit = self.rank == _fail_rank and data.numel() == _fail_size and not _fired
        if hit and _fail_before:
            _fired.append(1)
            raise RuntimeError("libhuffb200: CUDA error [injected before the collective]")
        r = _real_round_trip(self, data, comp_buf, out_buf, want_events)
        if hit:
            _fired.append(1)
            raise RuntimeError("libhuffb200: buffer too small (status 5) [injected]")
        return r

    sharded_mod.ShardedCodec.round_trip = _round_trip

import torch.distributed as dist  # noqa: E402

_init_pg = dist.init_process_group
dist.init_process_group = lambda backend=None, **kw: _init_pg("gloo")          # N > 1: gloo instead of NCCL

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
