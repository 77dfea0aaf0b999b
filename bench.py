#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: compress + decompress GB/s of input, and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload uniform|zipf|english|zipf15|zipf20]
                    [--size BYTES]

One "step" = one pass of the hot path over one batch of synthetic input resident in HBM:
    compress   = histogram kernel -> (N>1: all-gather of the G shard histograms inside the library) -> host tree -> encode kernel
    decompress = the one-pass fused decoder (or, by tree: fixed-length translation / the two-pass count + write kernels)
At N=1 the workload is BASELINE.json configs[1] (1 GiB uniform random bytes), two inputs rotated step by step so that every
step meets a tree the context did not see in the previous one.  At N>1 every rank holds one 1 GiB contiguous shard of an
N GiB input (weak scaling); shards exchange only the 2 KiB histograms.
`value` = input bytes of all ranks / max-over-ranks device time of the K steps (round trip: compress + decompress).
`e2e`   = the same round trip through the host-buffer C ABI (hb_compress_u8_into / hb_decompress_u8_into) with pinned HOST
          buffers, host<->device copies inside the timed region.
`configs` = the other BASELINE shapes on the same GPUs (Zipf 1 GiB / 4 GiB weak + strong, text, Fibonacci, Zipf(1.5)): the
          variable-length kernels; a failure there is contained (all ranks skip the config together) and a stalled run
          prints the line it has (Watchdog).
The oracle (oracle/) is executed only for `cpu_baseline` and for `--impl reference`; never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "compress+decompress throughput of input (round trip)"
UNIT = "GB/s"


# ---------------------------------------------------------------- fault containment
class ConfigFailed(Exception):
    """A secondary config failed on at least one rank (every rank raises it together, see all_ranks_ok)."""


class Watchdog:
    """A stalled run must not hold the GPUs: when nothing ticks for `limit_s` seconds (one rank raised and left a
    collective unmatched, a kernel hangs ...), rank 0 prints the JSON line it has -- the headline is measured first --
    and every rank leaves with os._exit; the other ranks wait a little longer so that rank 0 prints first."""

    def __init__(self, rank: int, limit_s: float):
        self.rank, self.limit_s = rank, limit_s + (0.0 if rank == 0 else 15.0)
        self.last = time.monotonic()
        self.line = None                   # rank 0: the JSON line as far as it is known
        self.headline_done = False         # every rank: the headline is measured (rank 0 holds a printable line)
        self.done = False
        self._thread = threading.Thread(target=self._watch, daemon=True)
        self._thread.start()

    def tick(self):
        self.last = time.monotonic()

    def _watch(self):
        while not self.done:
            time.sleep(1.0)
            if time.monotonic() - self.last > self.limit_s and not self.done:
                self.abort(f"no progress for {self.limit_s:.0f} s")

    def abort(self, why: str):
        sys.stderr.write(f"bench.py rank {self.rank}: {why}; leaving\n")
        sys.stderr.flush()
        if self.rank == 0 and self.line is not None:
            self.line["aborted"] = why
            print(json.dumps(self.line), flush=True)
        os._exit(0 if self.headline_done else 1)


def all_ranks_ok(ok: bool, torch, dist, dev) -> bool:
    """True when `ok` holds on every rank (one tiny all-reduce; every rank gets the same answer)."""
    if dist is None:
        return ok
    t = torch.tensor([0 if ok else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item()) == 0


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------- clocks sampler (NVML polled during the timed region)
class ClockSampler:
    """Polls SM clock + clock-event reasons every ~2 ms from a thread (NVML); falls back to `nvidia-smi -lms`."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, uuid: str | None = None):
        self.index, self.uuid = index, uuid
        self.sm, self.mask, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._smi = None
        self._smi_lines = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if hasattr(self.uuid, "encode") else self.uuid)
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._nvml = (pynvml, h, reasons_fn)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:
            self._nvml = None
        try:
            self._smi = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._thread = threading.Thread(target=self._read_smi, daemon=True)
            self._thread.start()
        except Exception:
            self._smi = None

    def _poll(self):
        pynvml, h, reasons_fn = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(reasons_fn(h))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self._smi.stdout:
            self._smi_lines.append(line.strip())

    def clear(self):
        self.sm.clear()
        self.mask = 0
        self._smi_lines.clear()

    def stop(self) -> dict:
        self._stop.set()
        if self._smi is not None:
            self._smi.terminate()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = set()
            for s in self._smi_lines:
                parts = [x.strip() for x in s.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    self.sm.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                except ValueError:
                    continue
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvidia-smi"}
        if self._nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        sm = sorted(self.sm)
        reasons = sorted(nm for bit, nm in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml"}


# ---------------------------------------------------------------- workloads
WORKLOADS = ("uniform", "zipf", "english", "zipf15", "zipf20", "fibonacci")
CONFIG_NAME = {"uniform": "BASELINE.json configs[1]: uniform random bytes", "zipf": "BASELINE.json configs[2]: Zipf(1.2) bytes",
               "english": "BASELINE.json configs[3]: English-like text", "fibonacci": "BASELINE.json configs[4]: Fibonacci-256 runs",
               "zipf15": "Zipf(1.5) bytes (codes up to 14 bits)", "zipf20": "Zipf(2.0) bytes"}


def make_workload(name: str, n: int, offset: int, device, seed_shift: int = 0):
    """seed_shift != 0 gives a second input of the same distribution with different counts (hence a different tree)."""
    from huff_encoding_b200 import datagen as G
    if name == "fibonacci":
        return G.from_weights_runs(G.fibonacci_weights(), device=device)
    base = {"uniform": G.SEED_BASE + 1, "zipf": G.SEED_BASE + 2, "english": G.SEED_BASE + 0,
            "zipf15": G.SEED_BASE + 2, "zipf20": G.SEED_BASE + 2}[name]
    kw = {"seed": base + 1000 * seed_shift, "offset": offset, "device": device}
    if name == "zipf15":
        return G.zipf(n, s=(15, 10), **kw)
    if name == "zipf20":
        return G.zipf(n, s=(20, 10), **kw)
    return {"uniform": G.uniform, "zipf": G.zipf, "english": G.english}[name](n, **kw)


def make_config(args, world: int) -> dict:
    """The SAME dict in both arms (ours / reference): what the metric is quoted on."""
    return {"workload": f"{CONFIG_NAME[args.workload]}, {args.size} B per GPU, weak scaling over contiguous shards; "
                        "the CPU reference arm times a bounded prefix of the same generator per step (cpu_baseline.sample)",
            "bytes_per_gpu": args.size, "inputs_rotated_per_step": 2,
            "l2": "input per step >> 126 MB L2", "parallelism": f"contiguous shards x{world}" if world > 1 else "single GPU"}


def cpu_port_round_trip(name: str, sample_bytes: int, repeats: int = 1):
    """The oracle's C port of the reference algorithm (single thread, like the reference's own loops,
    comp.rs:424-444 and :513-516) on a bounded sample of the workload: compress (count + tree + per-bit packing) and
    decompress (ONE bit-serial walk).  Returns (GB/s of input for compress + decompress, seconds, split)."""
    import numpy as np
    from oracle import oracle as O
    data = make_workload(name, sample_bytes, 0, None)
    O.lib()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        comp, pad, tree = O.compress(data)
        t1 = time.perf_counter()
        out = O.decompress(comp, pad, tree)
        t2 = time.perf_counter()
        assert out.size == data.size and np.array_equal(out[:4096], data[:4096])
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
    return data.size / best[0] / 1e9, best[0], {"compress_gbs": data.size / best[1] / 1e9, "decompress_gbs": data.size / best[2] / 1e9}


# ---------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0
    # bounded sample: calibrate on 4 MiB, then size the per-step sample so that K steps take ~100 s in total
    rate, _, _ = cpu_port_round_trip(args.workload, 4 << 20)                     # GB/s
    budget = int(rate * 1e9 * 100.0 / max(args.steps, 1))
    sample = max(1 << 20, min(args.size, 64 << 20, budget)) // (1 << 20) * (1 << 20)
    for _ in range(args.warmup):
        cpu_port_round_trip(args.workload, min(sample, 2 << 20))
    t_all, split = 0.0, None
    for _ in range(args.steps):
        _, dt, split = cpu_port_round_trip(args.workload, sample)
        t_all += dt
    value = sample * args.steps / t_all / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": make_config(args, world),
        "sample_bytes_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {sample} B of the workload per step; C port of the reference algorithm "
                                   "(oracle/huff_oracle.c), single thread like the reference's compress/decompress "
                                   "loops, one bit-serial decode walk; the Rust crate itself cannot be built here "
                                   "(no rustc/cargo)", "split_last_step": split},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm
PHASES = ("hist", "encode", "decode")


def _frac(bytes_, ms, peak):
    return round(bytes_ / (ms * 1e-3) / 1e9 / peak, 4) if ms and ms > 0 else None


def measure(codec, eng, datas, comp_buf, out_buf, steps, warmup, torch, dist, world, peak, tick=lambda: None):
    """Device-resident round trips over the inputs in `datas`, rotated step by step (so every step meets a tree the
    context did not see in the previous step).  Returns timing + per-phase CUDA-event breakdown + parity check.
    An error on one rank is noted, the schedule of collectives is kept, and all ranks raise ConfigFailed together at
    the next checkpoint: no rank is left waiting in a collective for one that has gone."""
    stream = eng.stream
    dev = comp_buf.device
    errors = []

    def trip(i, **kw):
        try:
            return codec.round_trip(datas[i % len(datas)], comp_buf, out_buf, **kw)
        except Exception as e:                                 # noqa: BLE001 -- reported through the checkpoint
            if not errors:
                errors.append(f"{type(e).__name__}: {e}")
            return None

    def checkpoint(what):
        tick()
        if not all_ranks_ok(not errors, torch, dist, dev):
            msg = errors[0] if errors else None
            if dist is not None:                                # every rank is here: tell all of them what happened where
                try:
                    msgs = [None] * world
                    dist.all_gather_object(msgs, msg)
                    msg = next((f"rank {r}: {m}" for r, m in enumerate(msgs) if m), msg)
                except Exception:                              # noqa: BLE001 -- the message is a nicety, the agreement is not
                    pass
            raise ConfigFailed(f"{what}: " + (msg or "failed on another rank"))

    with torch.cuda.stream(stream):
        for i in range(warmup):
            trip(i)
        stream.synchronize()
        checkpoint("warm-up")
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(steps):
            trip(i)
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total_ms = ev0.elapsed_time(ev1)
        last = datas[(steps - 1) % len(datas)]
        n = last.numel()
        if not errors and not (codec.last_info["n_letters"] == n and torch.equal(out_buf[:n], last)):
            errors.append("round trip mismatch")
        checkpoint("timed steps")
        # per-phase breakdown: extra steps with CUDA events between the phases (not part of the timing above)
        k = min(max(steps, 2), 6)
        marks = [trip(i, want_events=True) for i in range(k)]
        stream.synchronize()
        if not errors and any(m is None for m in marks):
            errors.append("phase breakdown: no events")
        checkpoint("phase breakdown")
    phase = {ph: sum(m[ph][0].elapsed_time(m[ph][1]) for m in marks) / len(marks) for ph in PHASES}
    t = torch.tensor([total_ms] + [phase[ph] for ph in PHASES], dtype=torch.float64, device=last.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    phase = {ph: float(t[1 + i]) for i, ph in enumerate(PHASES)}
    info = codec.last_info
    c = info["comp_len"]
    ms = total_ms / steps
    comp_ms, dec_ms = phase["hist"] + phase["encode"], phase["decode"]
    res = {"bytes_per_gpu": n, "comp_bytes_per_gpu": c, "ms_per_step": round(ms, 4),
           "round_trip_gbs": round(n * world / (ms * 1e-3) / 1e9, 1),
           "compress_gbs": round(n * world / (comp_ms * 1e-3) / 1e9, 1),
           "decompress_gbs": round(n * world / (dec_ms * 1e-3) / 1e9, 1),
           "phase_ms": {ph: round(v, 4) for ph, v in phase.items()},
           "frac": {"hist": _frac(n, phase["hist"], peak), "encode": _frac(n + c, phase["encode"], peak),
                    "decode": _frac(n + c, phase["decode"], peak), "compress": _frac(2 * n + c, comp_ms, peak)},
           "decoder": {0: "fixed-length translation" if info.get("fixed_len") else "two-pass (count + write)",
                       1: "fused one-pass", 2: "fused refuted -> two-pass"}.get(eng.ctx.last_decode_path()[0], "?")}
    return res


def e2e_round_trip(api, eng, data, steps, torch, dist, world, np):
    """The host-buffer C ABI (hb_compress_u8_into + hb_decompress_u8_into) with pinned HOST buffers: every step copies
    the letters H2D, the stream D2H, the stream H2D and the letters D2H inside the timed region."""
    n = data.numel()
    host_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data)
    host_comp = torch.empty(n + n // 8 + 4096, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(n + 64, dtype=torch.uint8, pin_memory=True)
    h_np, c_np, o_np = host_in.numpy(), host_comp.numpy(), host_out.numpy()
    cd = api.compress(h_np, ctx=eng.ctx, out=c_np)          # warm-up (device staging buffers)
    _ = api.decompress(cd, ctx=eng.ctx, out=o_np)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        cd = api.compress(h_np, ctx=eng.ctx, out=c_np)
        back = api.decompress(cd, ctx=eng.ctx, out=o_np)
    dt = (time.perf_counter() - t0) / steps
    assert back.size == n and np.array_equal(back[:65536], h_np[:65536]) and np.array_equal(back[-4096:], h_np[-4096:])
    clen = int(cd.comp_bytes().size)
    t = torch.tensor([dt], dtype=torch.float64, device=data.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"value": n * world / float(t.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": n + clen,
            "d2h_bytes_per_step": clen + n, "steps": steps,
            "api": "hb_compress_u8_into + hb_decompress_u8_into (pinned host buffers in and out)"}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to the GPU, so that the pinned host buffers of the e2e
    measurement are allocated (first touch) on the NUMA node the GPU's PCIe root hangs off.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)          # pinned host buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from huff_encoding_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from huff_encoding_b200 import api
    from huff_encoding_b200.engine import Engine
    from huff_encoding_b200.sharded import ShardedCodec

    peak, peak_src = measured_peak_gbs()
    wd = Watchdog(rank, float(os.environ.get("HB_BENCH_STALL_S", "240")))
    run_ours.watchdog = wd
    eng = Engine(local_rank)
    dev = eng.device                                  # cuda:<local_rank>
    codec = ShardedCodec(eng, world, rank, dist if world > 1 else None)
    if world > 1:
        # NCCL communicator inside libhuffb200: the timed steps are C calls only.  (NCCL announces its version on the
        # C-level stdout at the first init; stdout must carry the one JSON line only, so it goes to stderr meanwhile.)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            codec.init_library_comm()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    n = args.size
    d = dist if world > 1 else None

    def buffers(nbytes):
        return (torch.empty(nbytes + nbytes // 4 + 4096, dtype=torch.uint8, device=dev),
                torch.empty(nbytes + 64, dtype=torch.uint8, device=dev))

    # ---- headline: BASELINE.json configs[1] (or --workload), two inputs of the same shape rotated step by step: every
    #      compress() builds a tree from a fresh histogram and every decompress() meets a tree it did not see last step
    datas = [make_workload(args.workload, n, rank * n, dev, seed_shift=k) for k in range(2)]
    comp_buf, out_buf = buffers(max(x.numel() for x in datas))
    torch.cuda.synchronize()
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()
    time.sleep(0.05)
    # warm caches first (same input every step), then the timed, rotating run
    warm = measure(codec, eng, datas[:1], comp_buf, out_buf, max(3, args.steps // 2), args.warmup, torch, d, world, peak, wd.tick)
    sampler.clear()
    launches0 = eng.kernel_launches()
    head = measure(codec, eng, datas, comp_buf, out_buf, args.steps, args.warmup, torch, d, world, peak, wd.tick)
    clocks = sampler.stop()
    # measure() runs warm-up + timed + up to 6 event-instrumented steps: count the launches of the timed steps only
    per_step = (eng.kernel_launches() - launches0) / (args.warmup + args.steps + min(max(args.steps, 2), 6))
    launches = int(round(per_step * args.steps))
    wd.tick()
    e2e = e2e_round_trip(api, eng, datas[0], max(1, min(args.steps, 4)), torch, d, world, np)
    wd.tick()
    cpu = None
    if world == 1 and rank == 0:
        sample = min(n, 128 << 20)
        cpu_v, cpu_dt, cpu_split = cpu_port_round_trip(args.workload, sample)
        cpu = {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {sample} B of the workload, compress+decompress once ({cpu_dt:.1f} s); C port of the "
                         "reference algorithm, single thread, one bit-serial decode walk", "split": cpu_split}
    del datas
    torch.cuda.empty_cache()

    # ---- the JSON line as far as the headline goes (rank 0); the secondary configs are appended as they complete, and the
    #      watchdog prints what there is if a later stage stalls
    line = None
    if rank == 0:
        phase = head["phase_ms"]
        c_bytes = head["comp_bytes_per_gpu"]
        algo = {"hist": n, "encode": n + c_bytes, "decode": c_bytes + n}
        kern = {k: {"ms": phase[k], "algorithmic_bytes": algo[k],
                    "achieved_gbs": round(algo[k] / (phase[k] * 1e-3) / 1e9, 1) if phase[k] > 0 else None,
                    "frac": _frac(algo[k], phase[k], peak)} for k in PHASES}
        dom = max(PHASES, key=lambda k: phase[k])
        traffic, traffic_src = None, None
        try:                                              # measured DRAM bytes per launch from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{args.workload}:{n}", {}).get(dom)
            traffic_src = tj.get("_source") if traffic is not None else None
        except Exception:
            pass
        comp_ms = phase["hist"] + phase["encode"]
        roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernels": kern,
                "compress": {"ms": round(comp_ms, 4), "algorithmic_bytes": 2 * n + c_bytes, "frac": _frac(2 * n + c_bytes, comp_ms, peak)},
                "decompress": {"ms": phase["decode"], "algorithmic_bytes": n + c_bytes, "frac": _frac(n + c_bytes, phase["decode"], peak)}}
        line = {
            "metric": METRIC, "value": n * world / (head["ms_per_step"] * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": make_config(args, world),
            "comp_bytes_per_gpu": c_bytes,
            "cold_tree_ms_per_step": head["ms_per_step"], "warm_tree_ms_per_step": warm["ms_per_step"],
            "compress_gbs": head["compress_gbs"], "decompress_gbs": head["decompress_gbs"], "decoder": head["decoder"],
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "host_affinity": (f"{len(numa)} CPUs local to the GPU (NVML)" if numa else "not bound"),
        }
        if args.shrink:
            line["shrink"] = args.shrink
        wd.line = line
    wd.headline_done = True
    wd.tick()

    if world > 1:
        # bit-exactness of the sharded path (outside every timed region): the gathered shard streams must equal the
        # stream ONE GPU produces for the concatenated input
        m = 32 << 20
        part = make_workload("english", m, rank * m, dev)
        cb, ob = buffers(m)
        torch.cuda.synchronize()
        codec.round_trip(part, cb, ob)                      # the library path (hb_compress_shard_dev over NCCL)
        same = bool(torch.equal(ob[:m], part))
        got = codec.gather_stream(cb, codec.last_info)
        parts = [torch.empty_like(part) for _ in range(world)] if rank == 0 else None
        dist.gather(part, parts, dst=0)
        if rank == 0:
            whole = torch.cat(parts)
            out1, clen1, pad1, _ = Engine.compress(eng, whole)
            same = same and got[1] == pad1 and got[0].size == clen1 and np.array_equal(got[0], out1[:clen1].cpu().numpy())
            del whole, out1
            line["sharded_parity"] = ("gathered shard streams == single-GPU stream of the concatenated input "
                                      "(32 MiB per rank): " + ("ok" if same else "MISMATCH"))
        del part, cb, ob, parts
        torch.cuda.empty_cache()
        wd.tick()

    # ---- the other BASELINE configs on the same GPUs: the variable-length (real Huffman) kernels
    configs = []
    if rank == 0:
        line["configs"] = configs
    if not args.no_general:
        def run_config(label, workload, nbytes, scaling, steps=3, with_e2e=False):
            wd.tick()
            r = {"config": label, "workload": workload, "n_gpus": world, "scaling": scaling}
            ds = cb = ob = None
            try:
                ds = [make_workload(workload, nbytes, rank * nbytes, dev, seed_shift=k)
                      for k in range(1 if workload == "fibonacci" else 2)]
                cb, ob = buffers(max(x.numel() for x in ds))
                # the inputs are generated on torch's stream; the library's stream does not wait for it by itself
                torch.cuda.synchronize()
                r.update(measure(codec, eng, ds, cb, ob, steps, 2, torch, d, world, peak, wd.tick))
                if with_e2e:
                    r["e2e"] = e2e_round_trip(api, eng, ds[0], 2, torch, d, world, np)
            except ConfigFailed as e:                       # raised by every rank together: carry on with the next config
                r["error"] = str(e)
                sys.stderr.write(f"bench.py rank {rank}: config '{label}' failed: {e}\n")
            try:                                            # measured DRAM bytes per launch, where an ncu capture of this shape exists
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                if f"{workload}:{nbytes}" in tj and "error" not in r:
                    r["traffic_bytes_per_launch"] = tj[f"{workload}:{nbytes}"]
                    r["algorithmic_bytes"] = {"hist": r["bytes_per_gpu"], "encode": r["bytes_per_gpu"] + r["comp_bytes_per_gpu"],
                                              "decode": r["bytes_per_gpu"] + r["comp_bytes_per_gpu"]}
            except Exception:
                pass
            configs.append(r)
            if rank == 0 and "general_frac" not in line and "error" not in r and workload == "zipf" and nbytes == gib:
                line["general_frac"] = {"encode": r["frac"]["encode"], "decode": r["frac"]["decode"],
                                        "compress": r["frac"]["compress"], "workload": label}
            ds = cb = ob = None
            torch.cuda.empty_cache()
            wd.tick()

        gib = (1 << 30) >> args.shrink                      # --shrink: control-flow rehearsals only (tests/emu)
        run_config("1 GiB Zipf(1.2) per GPU (north_star: 1-GPU encode and decode on 1 GiB inputs)", "zipf", gib, "weak",
                   steps=5, with_e2e=True)
        if world == 1:
            run_config("configs[2]: 4 GiB Zipf(1.2) on one GPU", "zipf", 4 * gib, "weak")
            if not args.shrink:
                run_config("configs[4]: Fibonacci-256, 1 836 311 750 B, 40-bit codes", "fibonacci", 0, "weak", steps=2)
            run_config("Zipf(1.5), 1 GiB: 1.5 % of the letters have codes of 13 and 14 bits", "zipf15", gib, "weak")
        else:
            run_config(f"configs[2] weak: 4 GiB Zipf(1.2) per GPU x{world}", "zipf", 4 * gib, "weak")
            run_config(f"configs[2] strong: 4 GiB Zipf(1.2) in total over {world} GPUs", "zipf", 4 * gib // world, "strong")
        run_config(f"configs[3]: English-like text, 2 GiB contiguous shard per GPU x{world}", "english", 2 * gib, "weak")

    wd.done = True
    if rank == 0:
        if not configs:
            del line["configs"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uniform", choices=list(WORKLOADS))
    ap.add_argument("--size", type=int, default=1 << 30, help="bytes per GPU")
    ap.add_argument("--no-general", action="store_true", help="skip the secondary configs (zipf / text / fibonacci)")
    ap.add_argument("--shrink", type=int, default=0,
                    help="divide the secondary configs' sizes by 2^K and skip Fibonacci: rehearsals of the control flow only "
                         "(the line then carries \"shrink\": K and is no measurement of the named configs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    try:
        return run_ours(args)
    except BaseException as e:                                  # noqa: BLE001
        if isinstance(e, SystemExit):
            raise
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        wd = getattr(run_ours, "watchdog", None)
        rank = int(os.environ.get("RANK", "0"))
        if wd is not None and wd.headline_done and rank != 0:
            # the other ranks are (or will be) waiting for this one in a collective: let rank 0's watchdog print the line
            # it has, then leave through this rank's own watchdog
            while True:
                time.sleep(1.0)
        if wd is not None:
            wd.abort(f"{type(e).__name__}: {e}")
        # a plain `raise` would run the interpreter's teardown, which can wait for ever on a half-finished collective
        os._exit(1)


if __name__ == "__main__":
    sys.exit(main())
