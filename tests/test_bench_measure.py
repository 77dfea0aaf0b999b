"""bench.py's fault containment, checked on the CPU with stand-ins for the CUDA pieces (bench.measure takes `torch`, the
codec and the engine as arguments): an error on one rank must reach every rank as ConfigFailed at the next checkpoint
with the schedule of collectives intact, and a stalled run must end by itself with the JSON line it has."""
import json
import os
import socket
import subprocess
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _Event:
    clock = 0.0

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        _Event.clock += 1.0
        self.t = _Event.clock

    def elapsed_time(self, other):
        return other.t - self.t


class _Stream:
    def synchronize(self):
        pass


class _StreamCtx:
    def __init__(self, stream):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def fake_torch():
    cuda = types.SimpleNamespace(stream=_StreamCtx, synchronize=lambda: None, Event=_Event)
    return types.SimpleNamespace(cuda=cuda, equal=torch.equal, tensor=torch.tensor, float64=torch.float64, int32=torch.int32)


class FakeCodec:
    """round_trip = one collective (like hb_compress_shard_dev's all-gather) + a local copy; fails on request AFTER the
    collective, like a status returned by the decoder."""

    def __init__(self, d, fail_calls=()):
        self.d, self.fail_calls, self.calls, self.last_info = d, set(fail_calls), 0, None

    def round_trip(self, data, comp_buf, out_buf, want_events=False):
        self.calls += 1
        if self.d is not None:
            t = torch.ones(1)
            self.d.all_reduce(t)
        if self.calls in self.fail_calls:
            raise RuntimeError("libhuffb200: buffer too small (status 5)")
        out_buf[: data.numel()] = data
        self.last_info = {"n_letters": data.numel(), "comp_len": data.numel() // 2, "fixed_len": 0}
        if want_events:
            marks = {}
            for ph in bench.PHASES:
                a, b = _Event(), _Event()
                a.record()
                b.record()
                marks[ph] = (a, b)
            return marks
        return None


class FakeEngine:
    stream = _Stream()
    ctx = types.SimpleNamespace(last_decode_path=lambda: (1, 0))


def _measure(codec, d, world, ticks):
    datas = [torch.arange(100, dtype=torch.uint8), torch.arange(100, dtype=torch.uint8) + 1]
    comp, out = torch.zeros(200, dtype=torch.uint8), torch.zeros(164, dtype=torch.uint8)
    return bench.measure(codec, FakeEngine(), datas, comp, out, 4, 2, fake_torch(), d, world, 6538.3, lambda: ticks.append(1))


def test_measure_single_rank_result_and_failure():
    ticks = []
    r = _measure(FakeCodec(None), None, 1, ticks)
    for key in ("bytes_per_gpu", "comp_bytes_per_gpu", "ms_per_step", "round_trip_gbs", "compress_gbs", "decompress_gbs",
                "phase_ms", "frac", "decoder"):
        assert key in r, key
    assert r["bytes_per_gpu"] == 100 and r["decoder"] == "fused one-pass" and len(ticks) == 3
    with pytest.raises(bench.ConfigFailed, match="buffer too small"):
        _measure(FakeCodec(None, fail_calls={2}), None, 1, [])          # in the warm-up
    with pytest.raises(bench.ConfigFailed, match="timed steps"):
        _measure(FakeCodec(None, fail_calls={5}), None, 1, [])          # in a timed step


def test_measure_reports_a_round_trip_mismatch():
    class Corrupting(FakeCodec):
        def round_trip(self, data, comp_buf, out_buf, want_events=False):
            r = super().round_trip(data, comp_buf, out_buf, want_events)
            out_buf[3] ^= 1
            return r
    with pytest.raises(bench.ConfigFailed, match="round trip mismatch"):
        _measure(Corrupting(None), None, 1, [])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 1 fails in its 4th round trip (a timed step); rank 0 never fails by itself
        codec = FakeCodec(dist, fail_calls={4} if rank == 1 else ())
        try:
            _measure(codec, dist, world, [])
            first = "no error"
        except bench.ConfigFailed as e:
            first = str(e)
        # the schedule of collectives is intact: the next config runs to the end on both ranks
        r = _measure(FakeCodec(dist), dist, world, [])
        q.put((rank, first, r["bytes_per_gpu"], codec.calls))
    finally:
        dist.destroy_process_group()


def test_an_error_on_one_rank_fails_the_config_on_every_rank_and_the_next_config_still_runs():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, e0, n0, c0), (r1, e1, n1, c1) = got
    assert e0 == e1 and "rank 1: RuntimeError" in e0 and "buffer too small" in e0      # every rank learns what failed where
    assert n0 == n1 == 100
    assert c0 == c1 == 6                 # both ranks kept the schedule up to the checkpoint: 2 warm-up + 4 timed trips


def test_watchdog_prints_the_line_it_has_and_exits():
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "wd = bench.Watchdog(0, 1.0); wd.line = {'metric': 'm', 'value': 1.0}; wd.headline_done = True\n"
        "time.sleep(30)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0
    d = json.loads(r.stdout.strip())
    assert d["value"] == 1.0 and "no progress" in d["aborted"]
    # before the headline exists there is nothing to print: non-zero exit
    code2 = ("import sys, time; sys.path.insert(0, %r); import bench\nwd = bench.Watchdog(0, 1.0)\ntime.sleep(30)\n" % ROOT)
    r2 = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=60)
    assert r2.returncode == 1 and r2.stdout.strip() == ""
