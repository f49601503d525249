#!/usr/bin/env python
"""bench.py — chunks/s of the B200-native BirdNET v2.4 hot path (driver contract).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one batch of 3 s chunks:
  N = 1 : BASELINE config[1] — soundscape.wav, 3 s window / 1.5 s overlap (79 chunks) tiled to batch = 256.
  N > 1 : the same batch per GPU (weak scaling, independent chunks, no data-path collective); the only
          collective is the NCCL all-gather of the per-chunk top-10 (80 B/chunk), inside the timed region.
`value`  : chunks/s with the PCM already resident in HBM (device pointers through the C ABI).
`e2e`    : chunks/s through the host-buffer C ABI: pinned host float32 PCM -> H2D -> kernels -> sigmoid/top-10 -> D2H,
           every step, ONE caller using bnb_analyze_batch_submit / bnb_wait (two batches in flight, so the copy-in of step
           i+1 overlaps the kernels of step i).  Extra keys beside it: `value_synchronous_call` (bnb_analyze_batch, nothing
           overlapped), `value_int16_pcm` (int16 PCM as the reference's queue holds it), `value_two_callers` (two host
           threads with a handle each), top-level `latency_batch1_ms` (bnb_predict on one chunk from a CUDA graph, median of
           30), `config5_*` (8 realtime windows coalesced into one call: p50 / p99) and `config4_*` (bat backbone, embeddings).
`roofline`: the tcgen05 kernels (fused expand+depthwise and the other 1x1 GEMMs): algorithmic FLOPs / their CUDA-event
           time measured on ONE lane (no overlap) vs the measured bf16 dense peak; `traffic` from the committed ncu list.
`cpu_baseline` (N = 1 only): the oracle port on the host cores, bounded sample.  `clocks`: nvidia-smi samples taken
           during the timed regions.  `gpu_launches`: kernels launched by the library inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "birdnet-go_b200"))

METRIC = "3s-chunks/sec"
UNIT = "chunks/s"
BATCH = 256
N_SAMPLES = 144000
N_SPECIES = 6522
TOP_K = 10
# SURVEY.md §8(d) / BASELINE.md §4: algorithmic work per chunk
FLOP_PER_CHUNK = 2 * 328305984
FLOP_PW_PER_CHUNK = 2 * 279244800            # the dense 1x1 ("pointwise") layers: expand + project (+ pool-mix 1x1 in stem)
MIN_HBM_BYTES_PER_CHUNK = 576000 + 26088


def _peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------ workloads
def read_wav_int16(path):
    import struct
    b = open(path, "rb").read()
    p, fmt, data = 12, None, None
    while p + 8 <= len(b):
        cid, sz = b[p:p + 4], struct.unpack_from("<I", b, p + 4)[0]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", b, p + 8)
        elif cid == b"data":
            data = b[p + 8:p + 8 + sz]
        p += 8 + sz + (sz & 1)
    assert fmt[5] == 16 and fmt[1] == 1 and fmt[2] == 48000
    return np.frombuffer(data, "<i2")


def soundscape_batch(batch=BATCH):
    """config[1]: window 144000 / step 72000 over soundscape.wav (79 chunks), tiled to `batch`; float32 = int16/32768
    (internal/analysis/process.go:491-494)."""
    pcm = read_wav_int16(os.path.join(REPO, "assets", "soundscape.wav"))
    n = (len(pcm) - N_SAMPLES) // 72000 + 1
    idx = np.arange(batch) % n
    out = np.empty((batch, N_SAMPLES), np.float32)
    for i, j in enumerate(idx):
        out[i] = pcm[j * 72000:j * 72000 + N_SAMPLES].astype(np.float32) / np.float32(32768.0)
    return out


def synth_chunks(n, seed0=1234, fs=48000):
    """config[2] generator: per chunk seed = seed0 + id; pink noise (RMS 0.05) + linear chirp
    1->10 kHz at amplitude 0.2*U(0.1,1), clipped, quantised to int16, / 32768."""
    out = np.empty((n, N_SAMPLES), np.float32)
    t = np.arange(N_SAMPLES) / fs
    f = np.fft.rfftfreq(N_SAMPLES, 1.0 / fs)
    shape = np.zeros_like(f)
    shape[1:] = 1.0 / np.sqrt(f[1:])
    for i in range(n):
        rng = np.random.default_rng(seed0 + i)
        spec = (rng.standard_normal(len(f)) + 1j * rng.standard_normal(len(f))) * shape
        pink = np.fft.irfft(spec, N_SAMPLES)
        pink *= 0.05 / np.sqrt(np.mean(pink ** 2))
        amp = 0.2 * rng.uniform(0.1, 1.0)
        chirp = amp * np.sin(2 * np.pi * (1000.0 * t + 0.5 * (9000.0 / 3.0) * t * t))
        x = np.clip(pink + chirp, -1.0, 1.0)
        out[i] = np.round(x * 32767.0).astype(np.int16).astype(np.float32) / np.float32(32768.0)
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_info():
    try:
        return [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        return ""


def _cgroup_cpu_limit():
    """CPU quota of this container in cores (cgroup v2 cpu.max, then v1 cfs quota), or None when unlimited."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            return float(q) / float(per)
    except Exception:
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return q / per
    except Exception:
        pass
    return None


def _host_cores():
    """Cores this process can really use: the affinity mask, capped by the cgroup CPU quota (sched_getaffinity alone ignores
    the quota of a shared box — VERDICT r1 weak #9: the same '128 cores' gave 165 and 1483 chunks/s)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    lim = _cgroup_cpu_limit()
    if lim is not None:
        n = max(1, min(n, int(lim)))
    return n


def _host_state():
    try:
        la = [float(v) for v in open("/proc/loadavg").read().split()[:3]]
    except Exception:
        la = None
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count()
    return {"affinity_cores": aff, "cgroup_cpu_limit": _cgroup_cpu_limit(), "effective_cores": _host_cores(), "loadavg": la, "cpu": _cpu_info()}


def probe_reference_runtimes():
    """Is the reference's own arithmetic available on this box?  (libtensorflowlite_c as linked by go-tflite, or the Python
    runtimes that wrap the same kernels.)  BASELINE.md §5.2 / SURVEY §8(d).  Returns {name: path-or-version}."""
    import ctypes.util
    import glob
    import importlib.util
    found = {}
    pats = []
    for d in ["/usr/lib", "/usr/local/lib", "/usr/lib/x86_64-linux-gnu", "/opt", os.path.expanduser("~/.local/lib")] + os.environ.get("LD_LIBRARY_PATH", "").split(":"):
        if d:
            pats += [os.path.join(d, "libtensorflowlite_c*.so*"), os.path.join(d, "**", "libtensorflowlite_c*.so*"), os.path.join(d, "libonnxruntime.so*")]
    for pat in pats:
        try:
            for f in glob.glob(pat, recursive=True)[:1]:
                found["libtensorflowlite_c" if "tensorflowlite" in f else "libonnxruntime"] = f
        except Exception:
            pass
    n = ctypes.util.find_library("tensorflowlite_c")
    if n:
        found.setdefault("libtensorflowlite_c", n)
    for mod in ("tflite_runtime", "tensorflow", "ai_edge_litert", "onnxruntime"):
        try:
            if importlib.util.find_spec(mod) is not None:
                found[mod] = "python module"
        except Exception:
            pass
    return found


class TFLiteC:
    """The reference's backend itself through the TFLite C API, with the reference's thread rule
    (/root/reference/internal/inference/tflite/classifier.go:44-61: XNNPACK delegate with threads-1 workers, interpreter
    pinned to 1 thread; plain interpreter with `threads` when the delegate is unavailable)."""

    def __init__(self, lib_path, model_bytes, threads):
        import ctypes as C
        self.C = C
        L = self.L = C.CDLL(lib_path)
        L.TfLiteModelCreate.restype = C.c_void_p; L.TfLiteModelCreate.argtypes = [C.c_void_p, C.c_size_t]
        L.TfLiteInterpreterOptionsCreate.restype = C.c_void_p
        L.TfLiteInterpreterOptionsSetNumThreads.argtypes = [C.c_void_p, C.c_int32]
        L.TfLiteInterpreterCreate.restype = C.c_void_p; L.TfLiteInterpreterCreate.argtypes = [C.c_void_p, C.c_void_p]
        L.TfLiteInterpreterAllocateTensors.argtypes = [C.c_void_p]
        L.TfLiteInterpreterResizeInputTensor.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int), C.c_int32]
        L.TfLiteInterpreterGetInputTensor.restype = C.c_void_p; L.TfLiteInterpreterGetInputTensor.argtypes = [C.c_void_p, C.c_int32]
        L.TfLiteInterpreterGetOutputTensor.restype = C.c_void_p; L.TfLiteInterpreterGetOutputTensor.argtypes = [C.c_void_p, C.c_int32]
        L.TfLiteTensorCopyFromBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.TfLiteTensorCopyToBuffer.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.TfLiteInterpreterInvoke.argtypes = [C.c_void_p]
        self._model_bytes = model_bytes
        self.model = L.TfLiteModelCreate(model_bytes, len(model_bytes))
        opts = L.TfLiteInterpreterOptionsCreate()
        self.delegate = "none"
        try:                                             # XNNPACK delegate entry points are optional exports of the C library
            class XOpts(C.Structure):
                _fields_ = [("num_threads", C.c_int32), ("flags", C.c_uint32), ("weights_cache", C.c_void_p), ("handle_variable_ops", C.c_bool),
                            ("experimental_weight_cache_file_path", C.c_char_p), ("experimental_adaptive_avx_optimization", C.c_bool)]
            L.TfLiteXNNPackDelegateOptionsDefault.restype = XOpts
            L.TfLiteXNNPackDelegateCreate.restype = C.c_void_p; L.TfLiteXNNPackDelegateCreate.argtypes = [C.POINTER(XOpts)]
            L.TfLiteInterpreterOptionsAddDelegate.argtypes = [C.c_void_p, C.c_void_p]
            xo = L.TfLiteXNNPackDelegateOptionsDefault()
            xo.num_threads = max(1, threads - 1)
            d = L.TfLiteXNNPackDelegateCreate(C.byref(xo))
            if d:
                L.TfLiteInterpreterOptionsAddDelegate(opts, d)
                L.TfLiteInterpreterOptionsSetNumThreads(opts, 1)
                self.delegate = "xnnpack(%d)" % max(1, threads - 1)
        except AttributeError:
            pass
        if self.delegate == "none":
            L.TfLiteInterpreterOptionsSetNumThreads(opts, threads)
        self.it = L.TfLiteInterpreterCreate(self.model, opts)
        dims = (C.c_int * 2)(1, N_SAMPLES)
        L.TfLiteInterpreterResizeInputTensor(self.it, 0, dims, 2)
        if L.TfLiteInterpreterAllocateTensors(self.it) != 0:
            raise RuntimeError("TfLiteInterpreterAllocateTensors failed")
        self.out = np.empty(N_SPECIES, np.float32)

    def predict(self, x):
        C, L = self.C, self.L
        x = np.ascontiguousarray(x, np.float32)
        L.TfLiteTensorCopyFromBuffer(L.TfLiteInterpreterGetInputTensor(self.it, 0), x.ctypes.data_as(C.c_void_p), x.nbytes)
        if L.TfLiteInterpreterInvoke(self.it) != 0:
            raise RuntimeError("TfLiteInterpreterInvoke failed")
        L.TfLiteTensorCopyToBuffer(L.TfLiteInterpreterGetOutputTensor(self.it, 0), self.out.ctypes.data_as(C.c_void_p), self.out.nbytes)
        return self.out


def _cpu_worker(args):
    """One worker process of the CPU arm: `threads` intra-op threads, its own share of the chunks.  All workers start each
    timed repeat together (barrier) and report their own wall time of it."""
    threads, n_chunks, steps, repeats, barrier, tfl_path = args
    x = soundscape_batch(max(n_chunks, 1))
    if tfl_path:
        model = open(os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite"), "rb").read()
        net = TFLiteC(tfl_path, model, threads)
        run = lambda: [net.predict(x[i]) for i in range(len(x))]
    else:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import torch
        torch.set_num_threads(threads)
        import birdnet_oracle as bo
        o = bo.Oracle(dtype=torch.float32)
        run = lambda: o.predict_batch(x, batch=min(8, len(x)))
        o.predict_batch(x[:1], batch=1)
    run()                                              # warm-up: weights paged in, thread pools up
    out = []
    for _ in range(repeats):
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            run()
        out.append(time.perf_counter() - t0)
    return out


def cpu_reference_shaped(seconds=6.0, threads=None, tfl_path=None):
    """What the reference itself does (cmd/benchmark/benchmark.go:91-136): ONE interpreter, batch 1, intra-op threads
    only, inference serialized (orchestrator.go:531 global inferenceMu)."""
    threads = threads or min(16, _host_cores())
    x = soundscape_batch(4)
    if tfl_path:
        net = TFLiteC(tfl_path, open(os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite"), "rb").read(), threads)
        one = lambda i: net.predict(x[i % 4])
    else:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import torch
        import birdnet_oracle as bo
        torch.set_num_threads(threads)
        o = bo.Oracle(dtype=torch.float32)
        one = lambda i: o.predict_batch(x[i % 4:i % 4 + 1], batch=1)
    one(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        one(n); n += 1
    return n / (time.perf_counter() - t0), threads


def cpu_reference_run(steps, warmup, chunks_per_worker=4, threads_per_worker=2, repeats=3):
    """The reference's CPU implementation of the path on ALL host cores this container may use.
    kind = "reference" when libtensorflowlite_c is present on the box (the reference's own arithmetic, driven through the C
    API with its thread rule), else "port" (oracle/: torch-CPU fp32 restatement of the same .tflite; no Go toolchain and no
    TFLite library in this image — DESIGN.md).  Chunks are independent, so the whole-host number is data-parallel:
    effective_cores / threads_per_worker processes x threads_per_worker threads (the reference itself runs batch 1 under a
    global mutex; that shape is reported beside it).  Every repeat starts all workers together; a repeat's throughput =
    all chunks / slowest worker; reported: median (value), min, max, per-core, plus the host state that explains it."""
    import multiprocessing as mp
    state = _host_state()
    cores = state["effective_cores"]
    workers = max(1, cores // threads_per_worker)
    found = probe_reference_runtimes()
    tfl = found.get("libtensorflowlite_c")
    if tfl:
        try:
            TFLiteC(tfl, open(os.path.join(REPO, "assets", "BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite"), "rb").read(), 2)
        except Exception:
            tfl = None
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    barrier = mgr.Barrier(workers)
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        per = pool.map(_cpu_worker, [(threads_per_worker, chunks_per_worker, steps, repeats, barrier, tfl)] * workers)
    wall = time.perf_counter() - t0
    n = workers * chunks_per_worker * steps
    rep_s = [max(w[r] for w in per) for r in range(repeats)]            # slowest worker of each synchronised repeat
    rep_cps = sorted(n / t for t in rep_s)
    cps = float(np.median(rep_cps))
    dt = n / cps
    shaped, sthreads = cpu_reference_shaped(tfl_path=tfl)
    kind = "reference" if tfl else "port"
    what = ("libtensorflowlite_c (%s) through the C API, reference thread rule" % tfl) if tfl else "torch-CPU fp32 restatement of the .tflite graph (oracle/)"
    return cps, dt, {"value": cps, "unit": UNIT, "cores": workers * threads_per_worker, "kind": kind,
                     "value_min": rep_cps[0], "value_max": rep_cps[-1], "per_core": cps / (workers * threads_per_worker), "repeats": repeats,
                     "host": state, "runtimes_found": found,
                     "sample": "%d chunks per repeat = %d processes x %d threads x %d steps x %d soundscape chunks, %s, %d synchronised repeats (median), "
                               "%.1f s timed per repeat (%.1f s wall incl. start-up), %s" %
                               (n, workers, threads_per_worker, steps, chunks_per_worker, what, repeats, dt, wall, state["cpu"]),
                     "reference_shaped": {"value": shaped, "unit": UNIT, "threads": sthreads,
                                          "what": "one interpreter, batch 1, serialized (cmd/benchmark shape), same " + kind}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--micro-batch", type=int, default=128)
    ap.add_argument("--lanes", type=int, default=2)
    ap.add_argument("--precision", default="default", choices=["default", "f32", "f16x3"])
    ap.add_argument("--workload", default="config1", choices=["config1", "config2"],
                    help="config1 = BASELINE configs[1] (soundscape.wav, batch 256 per GPU: the headline, every N); config2 = configs[2] "
                         "(synthetic pink noise + chirp, batch 1024 per GPU-step, chunk ids sharded contiguously over the ranks with birdnet_b200.dist.shard_range)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-two-callers", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.workload == "config2" and a.batch == BATCH:
        a.batch = 1024
    config = {"workload": "soundscape.wav 3s window / 1.5s overlap (79 chunks) tiled to batch=%d per GPU, BirdNET v2.4 fp32 weights" % a.batch,
              "batch_per_gpu": a.batch, "global_batch": a.batch * world, "micro_batch": a.micro_batch, "lanes": a.lanes,
              "l2": "device inputs rotate over 2 distinct 147 MB buffers (> 126 MB L2)"}

    if a.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(a.steps, 4))          # each step = a bounded sample of the workload on every worker; 3 synchronised repeats
        cps, dt, cb = cpu_reference_run(steps, min(a.warmup, 1))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cps, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 1),
                          "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "soundscape.wav (reference repo) tiled", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": cps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import birdnet_b200 as bb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to fd 1 when the communicator comes up (NCCL_DEBUG=VERSION on the boxes): keep
        # stdout to the one JSON line by pointing fd 1 at stderr while the communicator is created
        # the only collective is an 80 B-per-chunk all-gather: one channel is plenty, and every extra NCCL CTA takes an SM away
        # from the 148-CTA persistent kernels it overlaps with (r02: 2.7 % per step at N = 2 and N = 4 alike)
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            t = torch.zeros(1, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    prec = {"default": bb.PRECISION_DEFAULT, "f32": bb.PRECISION_F32, "f16x3": bb.PRECISION_F16X3}[a.precision]
    clf = bb.B200Classifier(device=local, max_batch=a.batch, micro_batch=a.micro_batch, precision=prec, lanes=a.lanes)
    B = a.batch
    if a.workload == "config2":
        # BASELINE configs[2]: one GPU-step = 1024 synthetic chunks per rank; the global chunk index space [0, world * B) of a
        # step is cut into contiguous shards exactly like the 100 k-chunk job would be (seed = 1234 + chunk id)
        from birdnet_b200 import dist as bdist0
        lo, hi = bdist0.shard_range(world * B, rank, world)
        host = synth_chunks(hi - lo, seed0=1234 + lo)
        config["workload"] = "synthetic 48 kHz pink noise + chirp (seed 1234 + chunk id), %d chunks per GPU-step, contiguous shard [%d, %d) of %d, BirdNET v2.4 fp32 weights" % (B, lo, hi, world * B)
    else:
        host = soundscape_batch(B)
        if world > 1:
            host = np.roll(host, rank * 7, axis=0)     # each rank analyses its own shard of the stream
    # device-resident inputs: two distinct copies so consecutive steps never find their input in L2
    d_in = [torch.from_numpy(host).cuda(), torch.from_numpy(host[::-1].copy()).cuda()]
    d_logits = torch.empty((B, N_SPECIES), dtype=torch.float32, device="cuda")
    d_idx = torch.empty((B, TOP_K), dtype=torch.int32, device="cuda")
    d_conf = torch.empty((B, TOP_K), dtype=torch.float32, device="cuda")
    from birdnet_b200 import dist as bdist
    d_pack = torch.empty((B, 2 * TOP_K), dtype=torch.int32, device="cuda") if world > 1 else None
    g_pack = torch.empty((world * B, 2 * TOP_K), dtype=torch.int32, device="cuda") if world > 1 else None
    stream = torch.cuda.Stream()          # an explicit stream: the library enqueues on it, the CUDA events below time it
    torch.cuda.set_stream(stream)

    # N > 1: the only collective on the path is ONE packed all-gather of the per-chunk top-10 per step (birdnet_b200.dist, 80 B per
    # chunk).  It runs on a side stream ("issued right after the top-k kernel so it overlaps the next batch's
    # frontend"): the pack kernel stays on the compute stream, the gather of step i overlaps the kernels of step i + 1; the packed
    # buffers are double-buffered and the timed region ends only after the last gather (stream.wait_stream(side) below).
    side = torch.cuda.Stream() if world > 1 else None
    if world > 1:
        d_pack = [d_pack, torch.empty_like(d_pack)]; g_pack = [g_pack, torch.empty_like(g_pack)]
        ev_packed = [torch.cuda.Event(), torch.cuda.Event()]; ev_gathered = [torch.cuda.Event(), torch.cuda.Event()]
        for e in ev_gathered:
            e.record(side)

    def step(i):
        clf.analyze_batch_device(d_in[i & 1].data_ptr(), bb.PCM_F32, B, 1.0, TOP_K, d_idx.data_ptr(), d_conf.data_ptr(), d_logits.data_ptr(), stream.cuda_stream)
        if world > 1:
            j = i & 1
            stream.wait_event(ev_gathered[j])                  # the gather that last read this packed buffer has finished
            bdist.pack_topk(d_idx, d_conf, d_pack[j])
            ev_packed[j].record(stream)
            side.wait_event(ev_packed[j])
            with torch.cuda.stream(side):
                bdist.gather_topk_packed(d_pack[j], g_pack[j])
                ev_gathered[j].record(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, a.warmup)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = clf.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(a.steps):
        step(i)
    if side is not None:
        stream.wait_stream(side)                               # the last gathers are inside the timed region
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = clf.kernel_launches() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    value = world * B * a.steps / (ms * 1e-3)

    # per-kernel-category split inside the same kind of region (events around every launch)
    # ... on ONE lane: with two lanes kernels of different micro-batches overlap and event-bracketed durations would count
    # the overlap twice (the timed region above keeps the configured lanes)
    clf_p = clf if a.lanes == 1 else bb.B200Classifier(device=local, max_batch=a.batch, micro_batch=a.micro_batch, precision=prec, lanes=1)

    def pstep(i):
        clf_p.analyze_batch_device(d_in[i & 1].data_ptr(), bb.PCM_F32, B, 1.0, TOP_K, d_idx.data_ptr(), d_conf.data_ptr(), d_logits.data_ptr(), stream.cuda_stream)
    for i in range(3):
        pstep(i)
    torch.cuda.synchronize()
    clf_p.profile_begin()
    for i in range(a.steps):
        pstep(i)
    prof = clf_p.profile_end()
    if clf_p is not clf:
        clf_p.close()

    # end-to-end through the host-buffer C-ABI call, pinned host memory, H2D + D2H inside the timed region
    pin = [torch.from_numpy(host).pin_memory(), torch.from_numpy(host[::-1].copy()).pin_memory()]
    idx_h = np.empty((B, TOP_K), np.int32); conf_h = np.empty((B, TOP_K), np.float32)
    import ctypes as C

    def e2e_step(i):
        rc = clf._lib.bnb_analyze_batch(clf._h, C.c_void_p(pin[i & 1].data_ptr()), bb.PCM_F32, B, C.c_float(1.0), TOP_K,
                                        idx_h.ctypes.data_as(C.c_void_p), conf_h.ctypes.data_as(C.c_void_p), None)
        if rc != 0:
            raise RuntimeError(bb.last_error())

    # raw pinned H2D bandwidth of this box (context for e2e)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        d_in[0].copy_(pin[0], non_blocking=True)
    torch.cuda.synchronize(); h2d_gbs = 3 * B * N_SAMPLES * 4 / (time.perf_counter() - t0) / 1e9
    # The headline e2e: ONE caller keeps two batches in flight through bnb_analyze_batch_submit / bnb_wait, so the H2D copy of
    # step i+1 overlaps the kernels of step i (every step still pays its full H2D and D2H inside the timed region).
    outs2 = [(np.empty((B, TOP_K), np.int32), np.empty((B, TOP_K), np.float32)) for _ in range(2)]

    def e2e_submit(i):
        t = C.c_int32()
        rc = clf._lib.bnb_analyze_batch_submit(clf._h, C.c_void_p(pin[i & 1].data_ptr()), bb.PCM_F32, B, C.c_float(1.0), TOP_K,
                                               outs2[i & 1][0].ctypes.data_as(C.c_void_p), outs2[i & 1][1].ctypes.data_as(C.c_void_p), None, C.byref(t))
        if rc != 0:
            raise RuntimeError(bb.last_error())
        return t.value

    def e2e_wait(t):
        if clf._lib.bnb_wait(clf._h, t) != 0:
            raise RuntimeError(bb.last_error())

    def e2e_pipeline(n):
        prev = None
        for i in range(n):
            t = e2e_submit(i)
            if prev is not None:
                e2e_wait(prev)
            prev = t
        e2e_wait(prev)

    for i in range(3):
        e2e_step(i)
    e2e_pipeline(3)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        e2e_step(i)
    barrier()
    e2e_sync_s = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    e2e_pipeline(a.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    e2e_value = world * B * a.steps / e2e_s
    e2e_sync_value = world * B * a.steps / e2e_sync_s
    # the same call fed with the int16 PCM the reference's analysis queue actually holds (process.go:479-497 converts on
    # the host right before Predict; here /32768 is fused into the frontend load): half the H2D bytes
    host16 = np.clip(np.round(host * 32768.0), -32768, 32767).astype(np.int16)
    pin16 = [torch.from_numpy(host16).pin_memory(), torch.from_numpy(host16[::-1].copy()).pin_memory()]

    def e2e16_step(i):
        rc = clf._lib.bnb_analyze_batch(clf._h, C.c_void_p(pin16[i & 1].data_ptr()), bb.PCM_S16, B, C.c_float(1.0), TOP_K,
                                        idx_h.ctypes.data_as(C.c_void_p), conf_h.ctypes.data_as(C.c_void_p), None)
        if rc != 0:
            raise RuntimeError(bb.last_error())
    for i in range(3):
        e2e16_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        e2e16_step(i)
    barrier()
    e2e16_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e16_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e16_s = float(t.item())
    e2e16_value = world * B * a.steps / e2e16_s
    # two host threads, each with its own classifier handle and pinned buffers, calling the same synchronous entry point
    # (the shape of birdnet-go's concurrent analysis workers): one caller's H2D overlaps the other's kernels
    e2e2_value = None
    if not a.no_two_callers:
        clf2 = bb.B200Classifier(device=local, max_batch=a.batch, micro_batch=a.micro_batch, precision=prec, lanes=a.lanes)
        outs = [(idx_h, conf_h), (np.empty_like(idx_h), np.empty_like(conf_h))]

        def caller(c, k, n):
            for i in range(n):
                rc = c._lib.bnb_analyze_batch(c._h, C.c_void_p(pin[(i + k) & 1].data_ptr()), bb.PCM_F32, B, C.c_float(1.0), TOP_K,
                                              outs[k][0].ctypes.data_as(C.c_void_p), outs[k][1].ctypes.data_as(C.c_void_p), None)
                if rc != 0:
                    raise RuntimeError(bb.last_error())
        for rep in range(2):                                # first pass warms the second handle up
            n_each = 2 if rep == 0 else max(1, a.steps // 2)
            ths = [threading.Thread(target=caller, args=(c, k, n_each)) for k, c in enumerate((clf, clf2))]
            barrier()
            t0 = time.perf_counter()
            for t in ths: t.start()
            for t in ths: t.join()
            barrier()
            e2e2_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e2_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e2_s = float(t.item())
        e2e2_value = world * B * 2 * n_each / e2e2_s
        clf2.close()
    # BASELINE config 1 shape: one 3 s chunk through the drop-in bnb_predict (host float32 in, 6522 logits out), median of 30
    # after 5 warm-ups like cmd/perch-benchmark/main.go:29-33
    one = np.ascontiguousarray(host[0])
    clf1 = bb.B200Classifier(device=local, max_batch=8, micro_batch=8, precision=prec, use_graphs=1)     # the drop-in shape: small handle, chain replayed from a CUDA graph
    lat, lat_plain = [], []
    for i in range(35):
        t0 = time.perf_counter()
        clf1.predict(one)
        if i >= 5:
            lat.append(1e3 * (time.perf_counter() - t0))
    for i in range(35):
        t0 = time.perf_counter()
        clf.predict(one)
        if i >= 5:
            lat_plain.append(1e3 * (time.perf_counter() - t0))
    # BASELINE config 5 shape (realtime, one GPU): 8 streams whose 3 s windows become ready in the same 100 ms monitor tick are
    # coalesced into ONE int16 call (birdnet_b200.realtime.RealtimeCoalescer does this); latency = window ready -> top-10 on the host
    extra = {}
    if world == 1 and not a.no_two_callers:
        w8 = np.clip(np.round(host[:8] * 32768.0), -32768, 32767).astype(np.int16)
        l8 = []
        for i in range(110):
            t0 = time.perf_counter()
            clf1.analyze_batch(w8, 1.0, TOP_K)
            if i >= 10:
                l8.append(1e3 * (time.perf_counter() - t0))
        extra["config5_realtime_8_streams_coalesced"] = {"p50_ms": float(np.percentile(l8, 50)), "p99_ms": float(np.percentile(l8, 99)),
                                                         "windows_per_call": 8, "sustained_chunks_per_s": 8e3 / float(np.mean(l8)),
                                                         "note": "8 int16 windows ready in one tick -> one bnb_analyze_batch (CUDA graph); the reference's realtime path needs 8 serialized Predict calls under inferenceMu"}
        # BASELINE config 4 shape (bat backbone): 144000 samples captured at 256 kHz through the same backbone, embedding out
        # (bat_onnx.go:252); synthetic pink noise + 20-80 kHz sweeps; e2e through bnb_predict_batch with embeddings
        xb = synth_chunks(B, seed0=99000, fs=256000)
        tb = []
        for i in range(5):
            t0 = time.perf_counter()
            clf.predict_batch(xb, with_embeddings=True)
            tb.append(time.perf_counter() - t0)
        extra["config4_bat_backbone_embeddings"] = {"value": B / float(np.median(tb[1:])), "unit": UNIT,
                                                     "note": "the config-3 generator at fs = 256000; synchronous bnb_predict_batch, pageable float32 host PCM in, logits + 1024-d embeddings out (27 KB/chunk D2H): copy-bound"}
    clf1.close()
    lat_ms = float(np.median(lat))
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks, which = _peaks()
        pw_ms = prof["pw_expand"][0] + prof["pw_project"][0]
        pw_launches = prof["pw_expand"][1] + prof["pw_project"][1]
        pw_flops = FLOP_PW_PER_CHUNK * B * a.steps - 2 * 7077888 * B * a.steps      # minus the pool-mix 1x1 (fused in the stem kernel)
        tf = pw_flops / (pw_ms * 1e-3) / 1e12 if pw_ms > 0 else 0.0
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1373.2)))
        total_ms = sum(v[0] for v in prof.values())
        # DRAM traffic of the same kernels from the committed ncu launch list (tools/ncu_traffic.py), per launch like `achieved`
        traffic, tsrc = None, None
        for tname, knames in (("r02_dram_traffic.json", ("mbconv2_kernel", "pw2_kernel")), ("r01_dram_traffic.json", ("mbconv_tc_kernel", "pw_tc_kernel"))):
            tpath = os.path.join(REPO, "profiles", tname)
            if os.path.exists(tpath):
                tk = json.load(open(tpath))["kernels"]
                grp = [tk[k] for k in knames if k in tk]
                if grp:
                    traffic = sum(g["dram_bytes_per_step"] for g in grp) / max(1.0, sum(g["launches_per_step"] for g in grp))
                    tsrc = "profiles/%s (ncu dram__bytes_read+write, mean per launch of %s)" % (tname, " + ".join(knames))
                    break
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if clf.runtime_info()[2] == "FP32" else "f16x3(tcgen05 3-term split, f32 accumulate)+f32",
            "data": "soundscape.wav (reference repo fixture) tiled; weights = reference BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite",
            "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * N_SAMPLES * 4), "d2h_bytes_per_step": int(B * TOP_K * 8),
                    "api": "bnb_analyze_batch_submit / bnb_wait, one caller, two batches in flight (float32 PCM in pinned host memory) -> top-10 (idx, conf)",
                    "value_synchronous_call": e2e_sync_value, "h2d_gbs_measured": h2d_gbs,
                    "value_int16_pcm": e2e16_value, "value_two_callers": e2e2_value, "h2d_bytes_per_step_int16_pcm": int(B * N_SAMPLES * 2)},
            "roofline": {"bound": "tensor", "kernel": "pointwise 1x1 conv GEMMs (expand + project, %d launches/step)" % (pw_launches // max(1, a.steps)),
                         "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": traffic,
                         "traffic_source": tsrc,
                         "issued_tflops": 3 * tf, "note": "achieved = algorithmic 1x1-conv FLOPs / time of the tcgen05 kernels (mbconv2_kernel also does the SiLUs, the depthwise conv and the SE sums in that time; pw2_kernel the SE gating, residual and the hi/lo split of its output); every product is issued as 3 fp16 MMAs (hi*hi + lo*hi + hi*lo)",
                         "peak_source": which + " bf16 dense (sustained)", "share_of_step": pw_ms / total_ms if total_ms else None},
            "kernel_ms_per_step": {k: v[0] / a.steps for k, v in prof.items()},
            "latency_batch1_ms": lat_ms, "latency_batch1_ms_no_graph": float(np.median(lat_plain)),
            "kernel_ms_note": "per-category CUDA-event sums of one step run on a single lane (no inter-kernel overlap); sum > ms_per_step because the timed region overlaps two lanes",
            "flop_roofline_frac": (value / world) * FLOP_PER_CHUNK / (peak_tf * 1e12),
            "hbm_floor_frac": (value / world) * MIN_HBM_BYTES_PER_CHUNK / (float(peaks["hbm_gbs"]) * 1e9),
            "precision": clf.runtime_info()[2],
        }
        line.update(extra)
        if not a.no_cpu_baseline and world == 1:
            _, _, cb = cpu_reference_run(3, 1)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
