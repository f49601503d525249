"""Host mirror of the reference's orchestrator-level predict surface plus the batched form the B200 path needs.

Reference: `Orchestrator.PredictModel` /root/reference/internal/classifier/orchestrator.go:507-572 — three-level locking (a
read lock on the models map to fetch the entry, the global `inferenceMu` that serializes inference across ALL models, the entry's
own mutex for instance lifecycle), `globalInferenceCounters.RecordInvoke / RecordError`
(/root/reference/internal/classifier/inferencestats/counters.go:29-213).

`predict_model_batch` is the additive surface INTEGRATION.md section 4 describes: ONE pass through the same three locks and ONE
counter update for a whole batch of windows, so that the global inference lock is taken once per batch instead of once per window
(the reference's multi-stream bottleneck, SURVEY 8(a) a6 "lock wait under multi-stream load").

Pure host logic: works with any object offering `predict(sample)` / `predict_batch(windows)` (tests use fakes exactly like the
reference's own orchestrator tests; the GPU tests pass a `birdnet_b200.BirdNET`)."""
from __future__ import annotations

import math
import re
import threading
import time

LATENCY_WINDOW = 1024            # counters.go:13  latencyWindowSize
HEALTH_PERCENTILE = 0.95         # counters.go:18  healthLatencyPercentile


class Counters:
    """inferencestats.Counters (counters.go:29-106): invoke count / total / max (reset on snapshot) / lifetime max / errors and a
    ring of the last 1024 durations for the nearest-rank percentile."""

    def __init__(self):
        self._mu = threading.Lock()
        self.invoke_count = 0
        self.invoke_total_us = 0
        self.invoke_max_us = 0
        self.invoke_max_us_lifetime = 0
        self.invoke_errors = 0
        self.batch_windows = 0           # extension: windows served through predict_model_batch
        self._ring = [0] * LATENCY_WINDOW
        self._pos = 0
        self._len = 0

    def record_invoke(self, duration_us: int, windows: int = 1):
        with self._mu:
            self.invoke_count += 1
            self.invoke_total_us += duration_us
            self.invoke_max_us = max(self.invoke_max_us, duration_us)
            self.invoke_max_us_lifetime = max(self.invoke_max_us_lifetime, duration_us)
            self.batch_windows += windows
            self._ring[self._pos] = duration_us
            self._pos = (self._pos + 1) % LATENCY_WINDOW
            self._len = min(self._len + 1, LATENCY_WINDOW)

    def record_error(self):
        with self._mu:
            self.invoke_errors += 1

    def recent_percentile_us(self, p: float) -> int:
        """counters.go:60-78: nearest rank over the ring, idx = ceil(p n) - 1 clamped to [0, n - 1]; 0 when empty."""
        with self._mu:
            n = self._len
            if n == 0:
                return 0
            samples = sorted(self._ring[:n])
        idx = min(max(int(math.ceil(p * n)) - 1, 0), n - 1)
        return samples[idx]

    def snapshot(self) -> dict:
        """counters.go:96-106: the per-interval max is RESET by the read; errors and counts are cumulative."""
        with self._mu:
            s = dict(invoke_count=self.invoke_count, invoke_total_us=self.invoke_total_us, invoke_max_us=self.invoke_max_us,
                     invoke_errors=self.invoke_errors, collected_at=time.time())
            self.invoke_max_us = 0
        return s

    def peek(self) -> dict:
        """counters.go:216-246 PeekSnapshot: non-destructive."""
        with self._mu:
            s = dict(invoke_count=self.invoke_count, invoke_total_us=self.invoke_total_us,
                     invoke_max_us_lifetime=self.invoke_max_us_lifetime, invoke_errors=self.invoke_errors,
                     batch_windows=self.batch_windows)
        s["recent_p95_us"] = self.recent_percentile_us(HEALTH_PERCENTILE)
        return s


def sanitize_model_id(model_id: str) -> str:           # counters.go:123-130
    return re.sub(r"[^A-Za-z0-9_]", "_", model_id)


def metric_key(model_id: str) -> str:                  # counters.go:133-135
    return "inference." + sanitize_model_id(model_id) + ".avg_ms"


def rtf_metric_key(model_id: str) -> str:
    return "inference." + sanitize_model_id(model_id) + ".rtf"


def throughput_metric_key(model_id: str) -> str:
    return "inference." + sanitize_model_id(model_id) + ".throughput"


def error_rate_metric_key(model_id: str) -> str:
    return "inference." + sanitize_model_id(model_id) + ".error_rate"


class CounterMap:
    """inferencestats.CounterMap (counters.go:160-251)."""

    def __init__(self):
        self._mu = threading.Lock()
        self._models: dict[str, Counters] = {}

    def _get(self, model_id: str) -> Counters:
        with self._mu:
            c = self._models.get(model_id)
            if c is None:
                c = self._models[model_id] = Counters()
            return c

    def record_invoke(self, model_id: str, duration_us: int, windows: int = 1):
        self._get(model_id).record_invoke(duration_us, windows)

    def record_error(self, model_id: str):
        self._get(model_id).record_error()

    def snapshot_all(self) -> dict:
        with self._mu:
            items = list(self._models.items())
        return {k: c.snapshot() for k, c in items}

    def peek_all(self) -> dict:
        with self._mu:
            items = list(self._models.items())
        return {k: c.peek() for k, c in items}

    def delete(self, model_id: str):
        with self._mu:
            self._models.pop(model_id, None)


class OrchestratorError(Exception):
    pass


class _Entry:
    def __init__(self, instance):
        self.mu = threading.Lock()
        self.instance = instance


class Orchestrator:
    """The predict surface of classifier.Orchestrator: a registry of model instances, `predict` (primary model), `predict_model`
    and the batched `predict_model_batch`."""

    def __init__(self, counters: CounterMap | None = None):
        self._mu = threading.RLock()           # o.mu: guards the models map and the primary id
        self._inference_mu = threading.Lock()  # o.inferenceMu: one model runs at a time
        self._models: dict[str, _Entry] = {}
        self.primary_id: str | None = None
        self.counters = counters or CounterMap()

    def register(self, model_id: str, instance, primary: bool = False):
        with self._mu:
            self._models[model_id] = _Entry(instance)
            if primary or self.primary_id is None:
                self.primary_id = model_id

    def close_model(self, model_id: str):
        """Instance lifecycle under the entry lock (ReloadModel / Delete, orchestrator.go:1451-1533): predictions that arrive
        afterwards fail with 'has been closed' instead of touching freed native memory."""
        with self._mu:
            entry = self._models.get(model_id)
        if entry is None:
            return
        with entry.mu:
            inst, entry.instance = entry.instance, None
        if inst is not None and hasattr(inst, "close"):
            inst.close()

    def delete_model(self, model_id: str):
        self.close_model(model_id)
        with self._mu:
            self._models.pop(model_id, None)
        self.counters.delete(model_id)

    def predict(self, sample):
        with self._mu:
            model_id = self.primary_id
        return self.predict_model(model_id, sample)

    def _run(self, model_id: str, call, windows: int):
        with self._mu:                                         # map lock released before the inference / model locks
            entry = self._models.get(model_id)
        if entry is None:
            raise OrchestratorError("unknown model: %s" % model_id)
        with self._inference_mu:
            with entry.mu:
                if entry.instance is None:
                    raise OrchestratorError("model %s has been closed" % model_id)
                t0 = time.perf_counter()
                try:
                    out = call(entry.instance)
                except Exception:
                    self.counters.record_error(model_id)
                    raise
                self.counters.record_invoke(model_id, int((time.perf_counter() - t0) * 1e6), windows)
                return out

    def predict_model(self, model_id: str, sample):
        """orchestrator.go:514-572: `sample` is [][]float32 with ONE chunk (sample[0]); returns the model's top-k Results."""
        return self._run(model_id, lambda inst: inst.predict(sample), 1)

    def predict_model_batch(self, model_id: str, windows, **kw):
        """Batched surface: every window of `windows` (an [n, 144000] array or a list of chunks) in one pass through the locks;
        returns one Results list per window (BirdNET.predict_batch)."""
        n = len(windows)
        return self._run(model_id, lambda inst: inst.predict_batch(windows, **kw), n)
