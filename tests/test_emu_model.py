"""The `-m gpu` parity tests, run WITHOUT a GPU against the CPU execution model of the library (tests/emu).

tests/emu compiles the product's own kernel and host sources (huff_encoding_b200/csrc, untouched; see translate.py) with g++
against a small model of the CUDA execution model: a fiber per CUDA thread, real barriers and warp collectives, one
bounds-checked shared-memory arena per CTA, guarded device allocations, the PTX alignment rules of vector and bulk
accesses.  What runs is the same code path the B200 runs, kernel for kernel and launch for launch -- so indexing, halo,
barrier, look-back-protocol and host-logic errors show up here, on every CPU run, and not only at the next GPU session.
The model is test infrastructure: the product never loads it, and nothing measured ever comes from it."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

SUITES = ["tests/test_gpu_parity.py", "tests/test_gpu_fused.py", "tests/test_gpu_encode_warps.py", "tests/test_gpu_fuzz.py",
          "tests/test_gpu_fuzz_dev.py"]
# full-size property tests (1 GiB and more): the B200's job
DESELECT = ["tests/test_gpu_parity.py::test_config1_one_gib_uniform_properties"]


@pytest.fixture(scope="module")
def model_so():
    import build as emu_build
    return emu_build.build()


class _Result:
    def __init__(self, returncode, stdout, stderr):
        self.returncode, self.stdout, self.stderr = returncode, stdout, stderr


class _Jobs:
    """The model runs are independent subprocesses: they all start at once (the machine has the cores) and every test
    collects the result of its own."""

    def __init__(self, model_so):
        self.model_so, self.procs, self.done = model_so, {}, {}
        import tempfile
        self.tmp = tempfile.mkdtemp(prefix="hb_emu_jobs_")

    def start(self, name, cmd, extra_env=None):
        env = dict(os.environ, HB_EMU="1", HUFFB200_SO=self.model_so)
        env.update(extra_env or {})
        out = open(os.path.join(self.tmp, name + ".out"), "w+")
        err = open(os.path.join(self.tmp, name + ".err"), "w+")
        self.procs[name] = (subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=out, stderr=err, text=True), out, err)

    def result(self, name, timeout=3000):
        if name not in self.done:
            p, out, err = self.procs[name]
            try:
                p.wait(timeout=timeout)
            except subprocess.TimeoutExpired:
                p.kill()
                p.wait()
            out.seek(0)
            err.seek(0)
            self.done[name] = _Result(p.returncode, out.read(), err.read())
            out.close()
            err.close()
        return self.done[name]

    def close(self):
        for name, (p, out, err) in self.procs.items():
            if p.poll() is None:
                p.kill()


def _pytest_cmd(args):
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + args
    for d in DESELECT:
        cmd += ["--deselect", d]
    return cmd


def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.fixture(scope="module")
def jobs(model_so):
    j = _Jobs(model_so)
    emu = os.path.join(ROOT, "tests", "emu")
    j.start("suites", _pytest_cmd(SUITES + ["-n", "4"]), {"HB_EMU_WORKERS": "2"})
    j.start("single_rank", _pytest_cmd(["tests/test_gpu_multi.py", "-k", "single_rank"]))
    j.start("eight_ranks", [sys.executable, os.path.join(emu, "multirank_check.py"), "8"])
    j.start("bench_n1", [sys.executable, os.path.join(emu, "bench_dryrun.py"), "--size", str(4 << 20), "--steps", "4",
                         "--shrink", "8"])
    j.start("bench_n2", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                         "127.0.0.1", "--master-port", str(_free_port()), os.path.join(emu, "bench_dryrun.py"),
                         "--gpus", "2", "--size", str(2 << 20), "--steps", "3", "--shrink", "9"],
            {"HB_EMU_WORKERS": "3", "HB_DRYRUN_FAIL": "1:4194304"})
    j.start("bench_stall", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                            "127.0.0.1", "--master-port", str(_free_port()), os.path.join(emu, "bench_dryrun.py"),
                            "--gpus", "2", "--size", str(1 << 20), "--steps", "3", "--shrink", "10"],
            {"HB_EMU_WORKERS": "2", "HB_DRYRUN_FAIL": "1:2097152:before", "HB_BENCH_STALL_S": "20"})
    j.start("gloo", [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "tests/test_sharded_gloo.py"],
            {"HB_EMU_WORKERS": "2"})
    j.start("guard", [sys.executable, os.path.join(emu, "guard_check.py")])
    j.start("sharded_fuzz", [sys.executable, os.path.join(emu, "sharded_fuzz.py"), "30"], {"HB_EMU_WORKERS": "2"})
    j.start("few_sms", _pytest_cmd(["tests/test_gpu_fused.py", "-k",
                                    "fused_path_is_taken or matches_two_pass or wide_table or unaligned"]),
            {"HB_EMU_SMS": "3", "HB_EMU_WORKERS": "1", "HB_EMU_ORDER": "down", "HB_EMU_BULK": "lazy"})
    yield j
    j.close()


def test_the_model_exports_the_whole_c_abi(model_so):
    import ctypes as C
    from huff_encoding_b200 import _lib as L
    lib = C.CDLL(model_so)
    for name, _, _ in L.SYMBOLS:
        assert hasattr(lib, name), name


def test_gpu_parity_suites_pass_under_the_cpu_model(jobs):
    r = jobs.result("suites")
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 140, tail
    assert "skipped" not in r.stdout.splitlines()[-1], tail          # nothing may be skipped silently


def test_single_rank_communicator_under_the_cpu_model(jobs):
    # hb_comm_init(1 rank) + hb_compress_shard_dev / hb_decompress_shard_dev (no NCCL involved with one rank)
    r = jobs.result("single_rank")
    assert r.returncode == 0 and "1 passed" in r.stdout, (r.stdout + r.stderr)[-4000:]


def test_eight_ranks_inside_the_library_under_the_cpu_model(jobs):
    # hb_comm_init + hb_compress_shard_dev (all-gather of the shard histograms) + hb_decompress_shard_dev with 8 ranks as
    # threads: concatenated shard streams == the oracle's stream of the concatenated input
    r = jobs.result("eight_ranks")
    assert r.returncode == 0 and r.stdout.count("ok:") == 6, (r.stdout + r.stderr)[-4000:]


def test_bench_py_rehearsal_prints_one_json_line_with_every_key(jobs):
    # bench.py from argument parsing to its JSON line (N = 1), library = CPU model, sizes shrunk 256 x: a rehearsal of the
    # control flow the driver runs at round end -- no number in that line means anything
    import json
    r = jobs.result("bench_n1")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks",
                "cold_tree_ms_per_step", "warm_tree_ms_per_step", "general_frac", "configs", "shrink"):
        assert key in d, key
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["roofline"]["traffic"] is None   # (no capture at this size)
    assert len(d["configs"]) == 4 and not any("error" in c for c in d["configs"]), d["configs"]
    assert [c["decoder"] for c in d["configs"]] == ["fused one-pass"] * 4
    assert "e2e" in d["configs"][0]


def test_sharded_codec_over_gloo_with_the_real_kernels_under_the_cpu_model(jobs):
    # tests/test_sharded_gloo.py (2-4 processes, gloo) with the model engine instead of the oracle-backed one: the Python
    # orchestration (all-gather of histograms, shard plan, start-bit encode, byte-sharded speculative decode with the
    # neighbour check) over the library's own kernels
    r = jobs.result("gloo")
    assert r.returncode == 0 and "9 passed" in r.stdout, (r.stdout + r.stderr)[-4000:]


def test_bench_py_rehearsal_two_ranks_with_an_injected_failure(jobs):
    # bench.py under torchrun, 2 ranks (gloo; the library's communicator over the NCCL model), sizes shrunk 512 x.  Rank 1's
    # first round trip of the strong-scaling config raises after its collective (what happened on the 8-GPU box): both ranks
    # must drop that config together, run the next one, and rank 0 must still print the one JSON line
    import json
    r = jobs.result("bench_n2")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["sharded_parity"].endswith(": ok") and "aborted" not in d
    cfgs = d["configs"]
    assert [("error" in c) for c in cfgs] == [False, False, True, False], cfgs
    assert "strong" in cfgs[2]["config"] and cfgs[2]["scaling"] == "strong"
    assert cfgs[3]["decoder"] == "fused one-pass" and "general_frac" in d


def test_device_api_stays_inside_its_buffers_guard_pages(jobs, model_so):
    # every buffer ends where include/huffb200.h says the library may stop, followed by a PROT_NONE page: a byte too far
    # is a SIGSEGV (what compute-sanitizer would report on the GPU; the pool does not offer it)
    env = dict(os.environ, HB_EMU="1", HUFFB200_SO=model_so)
    r = jobs.result("guard")
    assert r.returncode == 0 and "guard pages: ok" in r.stdout, (r.returncode, (r.stdout + r.stderr)[-3000:])
    # negative control: a size that runs 16 bytes into the guard page must kill the process
    neg = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
           "import guard_check as g\nfrom tests.emu.model_engine import ModelEngine\nfrom huff_encoding_b200 import _lib as L\n"
           "eng = ModelEngine(); d = g.guarded(1 << 16, 1, 16)\n"
           "L.load().hb_histogram_u8_dev(eng.ctx.handle, d.data_ptr(), d.numel() + 16, eng._hist.data_ptr())\nprint('not caught')\n"
           % (ROOT, os.path.join(ROOT, "tests", "emu")))
    r = subprocess.run([sys.executable, "-c", neg], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode < 0 and "not caught" not in r.stdout, (r.returncode, r.stdout, r.stderr[-500:])


def test_bench_py_rehearsal_a_stalled_collective_ends_with_the_partial_line(jobs):
    # rank 1 raises BEFORE the collective of its first strong-scaling round trip: rank 0 is left waiting in the library's
    # all-gather.  Nothing progresses; after HB_BENCH_STALL_S rank 0 prints the line it has ("aborted") and both ranks leave
    import json
    r = jobs.result("bench_stall")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert "no progress" in d["aborted"] and d["n_gpus"] == 2 and d["value"] > 0
    assert len(d["configs"]) == 2 and not any("error" in c for c in d["configs"])       # the two configs before the stall


def test_sharded_codec_fuzz_with_thread_ranks_under_the_cpu_model(jobs):
    # ShardedCodec (compress / gather_stream / shard decode / byte-sharded decode with the neighbour check) over the real
    # kernels' code, 1..8 ranks as threads, shards from 0 letters up, streams down to a few bytes
    r = jobs.result("sharded_fuzz")
    assert r.returncode == 0 and "sharded fuzz: ok" in r.stdout, (r.stdout + r.stderr)[-3000:]


def test_cpp_mirror_of_the_reference_tests_under_the_cpu_model(model_so, tmp_path):
    # include/huff_coding.hpp (the C++ mirror of the reference API) with the reference's own test cases, linked against the model
    exe = str(tmp_path / "test_huff_coding_model")
    d = os.path.dirname(model_so)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_huff_coding.cpp"),
                           "-L" + d, "-lhuffb200_emu", "-Wl,-rpath," + d])
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_decoder_tests_with_few_sms_and_one_resident_cta(jobs):
    # other interleavings: 3 SMs, CTAs strictly one after the other, and the threads of a CTA taking their turns in
    # DESCENDING order (code that leans on the ascending order -- a missing barrier or __syncwarp -- gives other results),
    # bulk copies landing only when their mbarrier is waited on (a read of the window before the wait would see stale bytes)
    r = jobs.result("few_sms")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]


def test_the_product_refuses_to_load_the_model_outside_a_test_run(model_so):
    env = dict(os.environ, HUFFB200_SO=model_so)
    env.pop("HB_EMU", None)
    r = subprocess.run([sys.executable, "-c", "import huff_encoding_b200 as hb; hb.compress(b'abc')"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "not the library" in r.stderr
