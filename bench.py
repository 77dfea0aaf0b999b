#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: compress + decompress GB/s of input, and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload uniform|zipf|english|zipf15|zipf20]
                    [--size BYTES]

One "step" = one pass of the hot path over one batch of synthetic input resident in HBM:
    compress   = histogram kernel -> (N>1: allreduce of the 256 bins) -> host tree -> encode kernel
    decompress = count pass -> (N>1: neighbour entry check) -> write pass
At N=1 the workload is BASELINE.json configs[1] (1 GiB uniform random bytes).  At N>1 every rank holds one 1 GiB
contiguous shard of an N GiB input (weak scaling); shards exchange only the 2 KiB histogram and 8-byte bit totals.
`value` = input bytes of all ranks / max-over-ranks device time of the K steps (round trip: compress + decompress).
`e2e`   = the same round trip through the host-buffer C ABI (hb_compress_u8 / hb_decompress_u8) with pinned HOST
          buffers, host<->device copies inside the timed region.
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
def make_workload(name: str, n: int, offset: int, device):
    from huff_encoding_b200 import datagen as G
    gen = {"uniform": G.uniform, "zipf": G.zipf, "english": G.english,
           "zipf15": lambda *a, **k: G.zipf(*a, s=(15, 10), **k), "zipf20": lambda *a, **k: G.zipf(*a, s=(20, 10), **k)}[name]
    return gen(n, offset=offset, device=device)


def cpu_port_round_trip(name: str, sample_bytes: int, repeats: int = 1):
    """The oracle's C port of the reference algorithm (single thread, like the reference's own loops,
    comp.rs:424-444 and :513-516) on a bounded sample of the workload.  Returns GB/s of input for
    compress + decompress."""
    import numpy as np
    from oracle import oracle as O
    data = make_workload(name, sample_bytes, 0, None)
    O.lib()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        comp, pad, tree = O.compress(data)
        out = O.decompress(comp, pad, tree)
        dt = time.perf_counter() - t0
        assert out.size == data.size and np.array_equal(out[:4096], data[:4096])
        best = dt if best is None else min(best, dt)
    return sample_bytes / best / 1e9, best


# ---------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: calibrate on 4 MiB, then size the per-step sample so that K steps take ~100 s in total
    rate, _ = cpu_port_round_trip(args.workload, 4 << 20)                     # GB/s
    budget = int(rate * 1e9 * 100.0 / max(args.steps, 1))
    sample = max(1 << 20, min(args.size, 64 << 20, budget)) // (1 << 20) * (1 << 20)
    vals = []
    for _ in range(args.warmup):
        cpu_port_round_trip(args.workload, min(sample, 2 << 20))
    t_all = 0.0
    for _ in range(args.steps):
        v, dt = cpu_port_round_trip(args.workload, sample)
        vals.append(v)
        t_all += dt
    value = sample * args.steps / t_all / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_all / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload} random bytes, {args.size} B per GPU (BASELINE.json configs[1])",
                   "bytes_per_gpu": args.size},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {sample} B of the workload per step; C port of the reference algorithm "
                                   "(oracle/huff_oracle.c), single thread like the reference's compress/decompress "
                                   "loops; the Rust crate itself cannot be built here (no rustc/cargo)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from huff_encoding_b200 import build
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from huff_encoding_b200.engine import Engine
    from huff_encoding_b200.sharded import ShardedCodec

    eng = Engine(local_rank)
    codec = ShardedCodec(eng, world, rank, dist if world > 1 else None)
    n = args.size
    data = make_workload(args.workload, n, rank * n, dev)
    torch.cuda.synchronize()

    comp_buf = torch.empty(n + n // 4 + 4096, dtype=torch.uint8, device=dev)
    out_buf = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    stream = eng.stream
    phases = ["hist", "encode", "dec_count", "dec_write"]
    phase_ms = {k: 0.0 for k in phases}

    def step(record: bool):
        marks = codec.round_trip(data, comp_buf, out_buf, want_events=record)
        return marks

    with torch.cuda.stream(stream):
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        sampler = ClockSampler(local_rank, uuid)
        sampler.start()
        time.sleep(0.05)
        for _ in range(args.warmup):
            step(False)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.clear()                      # keep only samples taken during the timed region
        launches0 = eng.kernel_launches()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step(False)                       # the timed region: no per-phase events, nothing but the product path
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop()
        launches = eng.kernel_launches() - launches0
        # per-kernel breakdown: a few extra steps with CUDA events between the phases (not part of `value`)
        all_marks = [codec.round_trip(data, comp_buf, out_buf, want_events=True) for _ in range(min(args.steps, 10))]
        stream.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    for marks in all_marks:
        for k in phases:
            a, b = marks[k]
            phase_ms[k] += a.elapsed_time(b)
    for k in phases:
        phase_ms[k] /= len(all_marks)
    info = codec.last_info

    # correctness of what was timed (not in the timed region): round trip restores the shard
    assert info["n_letters"] == n and torch.equal(out_buf[:n], data), "round trip mismatch"

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e9

    # ---- secondary measurement: the GENERAL path (variable-length codes) on Zipf(1.2) bytes of the same size.
    # configs[1] (uniform) yields a fixed-length code set and takes the table-translation fast path; this shows what
    # the scan / self-synchronising kernels do.  Not part of `value`.
    general = None
    if args.workload == "uniform" and not args.no_general:
        zdata = make_workload("zipf", n, rank * n, dev)
        gphase = {k: 0.0 for k in phases}
        gsteps = 3
        with torch.cuda.stream(stream):
            for _ in range(2):
                codec.round_trip(zdata, comp_buf, out_buf)
            stream.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            gm = [codec.round_trip(zdata, comp_buf, out_buf, want_events=True) for _ in range(gsteps)]
            g1.record(stream)
            stream.synchronize()
        assert torch.equal(out_buf[:n], zdata), "general-path round trip mismatch"
        for m in gm:
            for k in phases:
                gphase[k] += m[k][0].elapsed_time(m[k][1]) / gsteps
        zc = codec.last_info["comp_len"]
        general = {"workload": f"zipf(1.2) bytes, {n} B per GPU", "ms_per_step": g0.elapsed_time(g1) / gsteps,
                   "comp_bytes": zc, "phase_ms": {k: round(v, 4) for k, v in gphase.items()},
                   "alg_bytes": {"hist": n, "encode": n + zc, "dec_count": zc, "dec_write": zc + n}}
        del zdata
        # a long-tailed input (Zipf(1.5): codes up to 14 bits, 1.5 % of the letters beyond the 12-bit decode table):
        # the decoder instances for trees with long codes
        ldata = make_workload("zipf15", n, rank * n, dev)
        lphase = {k: 0.0 for k in phases}
        with torch.cuda.stream(stream):
            for _ in range(2):
                codec.round_trip(ldata, comp_buf, out_buf)
            stream.synchronize()
            lm = [codec.round_trip(ldata, comp_buf, out_buf, want_events=True) for _ in range(gsteps)]
            stream.synchronize()
        assert torch.equal(out_buf[:n], ldata), "long-code round trip mismatch"
        for m in lm:
            for k in phases:
                lphase[k] += m[k][0].elapsed_time(m[k][1]) / gsteps
        general["zipf15_long_codes_phase_ms"] = {k: round(v, 4) for k, v in lphase.items()}
        del ldata
        # the same uniform input forced through the general kernels (fast path off): what configs[1] costs without it
        os.environ["HB_NO_FASTPATH"] = "1"
        try:
            eng2 = Engine(local_rank)
        finally:
            del os.environ["HB_NO_FASTPATH"]
        codec2 = ShardedCodec(eng2, world, rank, dist if world > 1 else None)
        uphase = {k: 0.0 for k in phases}
        with torch.cuda.stream(eng2.stream):
            for _ in range(2):
                codec2.round_trip(data, comp_buf, out_buf, want_events=True)
            eng2.stream.synchronize()
            um = [codec2.round_trip(data, comp_buf, out_buf, want_events=True) for _ in range(gsteps)]
            eng2.stream.synchronize()
        assert torch.equal(out_buf[:n], data), "forced general-path round trip mismatch"
        for m in um:
            for k in phases:
                uphase[k] += m[k][0].elapsed_time(m[k][1]) / gsteps
        general["uniform_forced_general_phase_ms"] = {k: round(v, 4) for k, v in uphase.items()}
        del codec2, eng2
        codec.round_trip(data, comp_buf, out_buf)      # restore last_info for the headline workload
        info = codec.last_info

    # ---- e2e through the host-buffer C ABI with pinned host memory
    from huff_encoding_b200 import api
    e2e_n = n
    host_in = torch.empty(e2e_n, dtype=torch.uint8, pin_memory=True)
    host_in.copy_(data[:e2e_n])
    host_comp = torch.empty(e2e_n + e2e_n // 8 + 4096, dtype=torch.uint8, pin_memory=True)
    host_out = torch.empty(e2e_n + 64, dtype=torch.uint8, pin_memory=True)
    h_np, c_np, o_np = host_in.numpy(), host_comp.numpy(), host_out.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    cd = api.compress(h_np, ctx=eng.ctx, out=c_np)          # warm-up (device staging buffers, tables)
    _ = api.decompress(cd, ctx=eng.ctx, out=o_np)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cd = api.compress(h_np, ctx=eng.ctx, out=c_np)      # H2D of the letters + D2H of the stream inside
        back = api.decompress(cd, ctx=eng.ctx, out=o_np)    # H2D of the stream + D2H of the letters inside
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    assert back.size == e2e_n and np.array_equal(back[:65536], h_np[:65536]) and np.array_equal(back[-4096:], h_np[-4096:])
    clen = int(cd.comp_bytes().size)
    t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_n * world / float(t.item()) / 1e9

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        c_bytes = info["comp_len"]
        algo = {"hist": n, "encode": n + c_bytes, "dec_count": c_bytes, "dec_write": c_bytes + n}
        kern = {}
        fixed = info.get("fixed_len", 0)
        for k in phases:
            ms = phase_ms[k]
            host_only = (k == "dec_count" and fixed)    # fixed-length code set: the count is host arithmetic, no kernel
            kern[k] = {"ms": round(ms, 4), "algorithmic_bytes": 0 if host_only else algo[k],
                       "achieved_gbs": round(algo[k] / (ms * 1e-3) / 1e9, 1) if ms > 0 and not host_only else None,
                       "frac": round(algo[k] / (ms * 1e-3) / 1e9 / peak, 4) if ms > 0 and not host_only else None}
            if host_only:
                kern[k]["note"] = "no kernel: fixed-length code set, letter count = bits / L on the host"
        # the dominant kernel = the largest share of the step; decode is reported as count+write against C+N
        dom = max([k for k in phases if kern[k]["frac"] is not None], key=lambda k: phase_ms[k])
        dec_ms = phase_ms["dec_count"] + phase_ms["dec_write"]
        comp_ms = phase_ms["hist"] + phase_ms["encode"]
        traffic, traffic_src = None, None
        try:                                              # measured DRAM bytes per launch from the committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
            traffic = tj.get(f"{args.workload}:{n}", {}).get(dom)
            traffic_src = tj.get("_source") if traffic is not None else None
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernels": kern,
                "compress": {"ms": round(comp_ms, 4), "algorithmic_bytes": 2 * n + c_bytes,
                             "frac": round((2 * n + c_bytes) / (comp_ms * 1e-3) / 1e9 / peak, 4)},
                "decompress": {"ms": round(dec_ms, 4), "algorithmic_bytes": n + c_bytes,
                               "frac": round((n + c_bytes) / (dec_ms * 1e-3) / 1e9 / peak, 4)}}
        sample = min(n, 128 << 20)
        cpu_v, cpu_dt = cpu_port_round_trip(args.workload, sample) if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload} random bytes, {n} B per GPU (BASELINE.json configs[1])",
                       "bytes_per_gpu": n, "comp_bytes_per_gpu": c_bytes, "l2": "input per step >> 126 MB L2",
                       "parallelism": f"contiguous shards x{world}" if world > 1 else "single GPU"},
            "compress_gbs": n * world / (comp_ms * 1e-3) / 1e9, "decompress_gbs": n * world / (dec_ms * 1e-3) / 1e9,
            "roofline": roof,
            "cpu_baseline": ({"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                              "sample": f"first {sample} B of the workload, compress+decompress once "
                                        f"({cpu_dt:.1f} s); C port of the reference algorithm, single thread"}
                             if cpu_v is not None else None),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_n + clen,
                    "d2h_bytes_per_step": clen + e2e_n, "steps": e2e_steps,
                    "api": "hb_compress_u8_into + hb_decompress_u8_into (pinned host buffers in and out)"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if general is not None:
            gp = general["phase_ms"]
            ga = general["alg_bytes"]
            general["frac"] = {k: round(ga[k] / (gp[k] * 1e-3) / 1e9 / peak, 4) if gp[k] > 0 else None for k in gp}
            cms, dms = gp["hist"] + gp["encode"], gp["dec_count"] + gp["dec_write"]
            general["compress_gbs"] = n * world / (cms * 1e-3) / 1e9
            general["decompress_gbs"] = n * world / (dms * 1e-3) / 1e9
            general["compress_frac"] = round((2 * n + general["comp_bytes"]) / (cms * 1e-3) / 1e9 / peak, 4)
            general["decompress_frac"] = round((n + general["comp_bytes"]) / (dms * 1e-3) / 1e9 / peak, 4)
            line["general_path"] = general
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uniform", choices=["uniform", "zipf", "english", "zipf15", "zipf20"])
    ap.add_argument("--size", type=int, default=1 << 30, help="bytes per GPU")
    ap.add_argument("--no-general", action="store_true", help="skip the secondary general-path (zipf) measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
