"""Realtime shape of the hot path (BASELINE config 5): N PCM sources -> per-source AnalysisBuffer -> one coalesced
batched call per tick instead of N serialized batch-1 Predict calls under the reference's global `inferenceMu`.

Restates, in the host language available here, the reference's chunk queue:
  * BufferConsumer.Write: 16-bit LE mono PCM frames fan out into a per-source ring
    (/root/reference/internal/analysis/buffer_consumer.go:163-230)
  * AnalysisBuffer: overwrite-mode ring + overlap prefix; Read() -> [overlap | fresh] window or None
    (/root/reference/internal/audiocore/buffer/analysis.go:152-251; BirdNET geometry 144000 B prefix + 144000 B fresh,
    /root/reference/internal/classifier/model.go:41-46)
  * analysisBufferMonitor: poll every tick, dispatch ready windows (/root/reference/internal/analysis/buffer_manager.go:390-500)
  * ProcessData: convert, infer, enqueue Results (process.go:253-422), non-blocking queue with drop accounting (:405-420)

What changes on the B200 backend is only the dispatch: every window that became ready in a tick goes into ONE
`bnb_analyze_batch` call with the int16 bytes as they sit in the ring (the conversion /32768 happens on the device).
"""
from __future__ import annotations

import collections
import time

import numpy as np

WINDOW_BYTES = 288000          # 3 s of 48 kHz int16 mono
DEFAULT_QUEUE_SIZE = 100       # classifier.DefaultQueueSize (queue.go:22)


class AnalysisBuffer:
    """Python mirror of buffer.AnalysisBuffer (same constructor checks, same Read/Write semantics)."""

    def __init__(self, capacity, overlap_size, read_size, source_id):
        if capacity <= 0:
            raise ValueError("invalid analysis buffer capacity: %d, must be greater than 0" % capacity)
        if overlap_size < 0:
            raise ValueError("invalid overlap size: %d, must be >= 0" % overlap_size)
        if read_size <= 0:
            raise ValueError("invalid read size: %d, must be greater than 0" % read_size)
        if read_size < overlap_size:
            raise ValueError("read size %d must be >= overlap size %d" % (read_size, overlap_size))
        if capacity < read_size:
            raise ValueError("capacity %d must be >= read size %d" % (capacity, read_size))
        if not source_id:
            raise ValueError("source ID must not be empty")
        self.ring = np.zeros(capacity, np.uint8)
        self.head = 0
        self.length = 0
        self.prev = None
        self.overlap, self.read_size, self.source_id = overlap_size, read_size, source_id
        self.overwrites = 0

    def write(self, data: bytes):
        d = np.frombuffer(data, np.uint8)
        cap = len(self.ring)
        if len(d) > cap - self.length:
            self.overwrites += 1
        if len(d) >= cap:
            d = d[-cap:]
            self.head, self.length = 0, 0
        pos = (self.head + self.length) % cap
        first = min(len(d), cap - pos)
        self.ring[pos:pos + first] = d[:first]
        self.ring[:len(d) - first] = d[first:]
        over = max(0, self.length + len(d) - cap)
        self.head = (self.head + over) % cap
        self.length = min(cap, self.length + len(d))

    def read(self):
        """None = "try again later"; else uint8 array [overlap prefix | read_size fresh bytes]."""
        if self.length < self.read_size:
            return None
        cap = len(self.ring)
        window = np.zeros(self.overlap + self.read_size, np.uint8)
        if self.overlap and self.prev is not None:
            window[:self.overlap] = self.prev
        idx = (self.head + np.arange(self.read_size)) % cap
        window[self.overlap:] = self.ring[idx]
        self.head = (self.head + self.read_size) % cap
        self.length -= self.read_size
        if self.overlap:
            self.prev = window[-self.overlap:].copy()
        return window


class Results:
    """classifier.Results (queue.go:10-19): what ProcessData enqueues for the detection processor."""
    __slots__ = ("source_id", "start_time", "pcm", "results", "elapsed")

    def __init__(self, source_id, start_time, pcm, results, elapsed):
        self.source_id, self.start_time, self.pcm, self.results, self.elapsed = source_id, start_time, pcm, results, elapsed


BUFFER_OVERRUN_REPORT_COOLDOWN_S = 3600.0     # process.go:36  bufferOverrunReportCooldown
BUFFER_OVERRUN_MIN_COUNT = 10                 # process.go:39  bufferOverrunMinCount
DEFAULT_BUFFER_INTERVAL_S = 1.5               # process.go:353 fallback: BirdNET v2.4 (3 s window, 50 % overlap)


class BufferOverrunTracker:
    """bufferOverrunTracker + recordBufferOverrun (process.go:44-215): tumbling window per (source, model); when the window
    expires with >= BUFFER_OVERRUN_MIN_COUNT overruns ONE report is emitted (the reference sends a Sentry event)."""

    def __init__(self, source, model_id, report=None):
        self.source, self.model_id = source, model_id
        self.overrun_count = 0
        self.window_start = None
        self.max_elapsed = 0.0
        self.buffer_length = 0.0
        self.reports = []
        self._report = report

    def record(self, elapsed, buffer_len, now):
        if self.window_start is None:
            self.window_start = now
        if now - self.window_start >= BUFFER_OVERRUN_REPORT_COOLDOWN_S:
            if self.overrun_count >= BUFFER_OVERRUN_MIN_COUNT:
                rep = {"source": self.source, "model_id": self.model_id, "overrun_count": self.overrun_count,
                       "max_elapsed_ms": int(self.max_elapsed * 1e3), "buffer_length_ms": int(self.buffer_length * 1e3),
                       "reporting_window_minutes": int((now - self.window_start) / 60.0)}
                self.reports.append(rep)
                if self._report:
                    self._report(rep)
            self.overrun_count, self.max_elapsed, self.window_start = 0, 0.0, now
        self.overrun_count += 1
        if elapsed > self.max_elapsed:
            self.max_elapsed, self.buffer_length = elapsed, buffer_len


class RealtimeCoalescer:
    """One AnalysisBuffer per source; `tick()` plays analysisBufferMonitor for all of them and issues ONE batched call."""

    def __init__(self, analyze_batch, labels, sources, overlap_bytes=144000, read_bytes=144000, capacity_bytes=3 * WINDOW_BYTES,
                 sensitivity=1.0, top_k=10, queue_size=DEFAULT_QUEUE_SIZE, model_id="BirdNET_V2.4"):
        # analyze_batch(pcm_int16 [B,144000], sensitivity, k) -> (idx [B,k], conf [B,k]); e.g. B200Classifier.analyze_batch
        self.analyze_batch, self.labels = analyze_batch, labels
        self.sensitivity, self.top_k = sensitivity, top_k
        if overlap_bytes + read_bytes != WINDOW_BYTES:
            raise ValueError("window must be %d bytes (3 s of 48 kHz int16)" % WINDOW_BYTES)
        self.buffers = {s: AnalysisBuffer(capacity_bytes, overlap_bytes, read_bytes, s) for s in sources}
        self.queue = collections.deque()
        self.queue_size = queue_size
        self.dropped = 0
        self.batches = 0
        self.windows = 0
        self.first_window = {s: True for s in sources}
        self.buffer_interval = DEFAULT_BUFFER_INTERVAL_S * (read_bytes / 144000.0)   # spec.BufferInterval(): the fresh part of a window
        self.model_id = model_id
        self.overruns = {}                                    # "source:modelID" -> BufferOverrunTracker (getOverrunTracker, process.go:66-76)
        self.overrun_total = 0

    def write(self, source_id, frame: bytes):
        """BufferConsumer.Write for one source: frame = int16 LE mono PCM at the model rate."""
        self.buffers[source_id].write(frame)

    def tick(self, now=None):
        """Collect every ready window, run them as one batch, enqueue Results (drop when the queue is full)."""
        ready = []
        for s, b in self.buffers.items():
            w = b.read()
            if w is not None:
                if self.first_window[s] and b.overlap:       # first window has a zero overlap prefix (analysis.go:205-211)
                    self.first_window[s] = False
                ready.append((s, w))
        if not ready:
            return 0
        t0 = time.perf_counter()
        pcm = np.stack([w.view("<i2") for _, w in ready])
        idx, conf = self.analyze_batch(pcm, self.sensitivity, self.top_k)
        dt = time.perf_counter() - t0
        self.batches += 1
        self.windows += len(ready)
        # process.go:351-370: inference slower than the buffer interval means the pipeline falls behind real time; every window of
        # the batch waited for the same call, so the overrun is recorded per source
        if dt > self.buffer_interval:
            t_now = time.monotonic() if now is None else now
            for s, _ in ready:
                key = s + ":" + self.model_id
                tr = self.overruns.get(key)
                if tr is None:
                    tr = self.overruns[key] = BufferOverrunTracker(s, self.model_id)
                tr.record(dt, self.buffer_interval, t_now)
                self.overrun_total += 1
        for (s, w), ri, rc in zip(ready, idx, conf):
            res = [(self.labels[i], float(c)) for i, c in zip(ri, rc)]
            if len(self.queue) >= self.queue_size:           # process.go:405-420: non-blocking send, count the drop
                self.dropped += 1
                continue
            self.queue.append(Results(s, now, w, res, dt))
        return len(ready)
