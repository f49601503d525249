"""the orchestrator-level predict surface and its inference counters, host logic only (fakes as backends, like the
reference's own tests).  The counter cases restate /root/reference/internal/classifier/inferencestats/counters_test.go."""
import threading
import time

import pytest

from birdnet_b200.orchestrator import (CounterMap, Counters, Orchestrator, OrchestratorError, error_rate_metric_key, metric_key,
                                       rtf_metric_key, sanitize_model_id, throughput_metric_key)


def test_counters_record_and_snapshot_resets_only_the_interval_max():       # counters_test.go:14-55, 337-359
    c = Counters()
    c.record_invoke(100); c.record_invoke(300); c.record_invoke(200); c.record_error()
    s = c.snapshot()
    assert (s["invoke_count"], s["invoke_total_us"], s["invoke_max_us"], s["invoke_errors"]) == (3, 600, 300, 1)
    assert abs(s["collected_at"] - time.time()) < 5
    s2 = c.snapshot()
    assert s2["invoke_max_us"] == 0 and s2["invoke_count"] == 3 and s2["invoke_errors"] == 1          # max reset, the rest cumulative
    assert c.peek()["invoke_max_us_lifetime"] == 300                                                    # :458-470 lifetime max survives


@pytest.mark.parametrize("samples,p,want", [([], 0.95, 0), ([42], 0.95, 42), ([10, 20, 30, 40, 50, 60, 70], 0.95, 70),
                                            ([1, 2, 3, 4], 1.0, 4), ([10, 20, 30, 40, 50], 0.5, 30), ([5, 6, 7], 0.0, 5)])
def test_recent_percentile_boundaries(samples, p, want):                     # counters_test.go:235-262
    c = Counters()
    for s in samples:
        c.record_invoke(s)
    assert c.recent_percentile_us(p) == want


def test_recent_p95_ignores_outliers_and_evicts_old_samples():               # counters_test.go:202-233
    c = Counters()
    for _ in range(99):
        c.record_invoke(10)
    c.record_invoke(100000)
    assert c.recent_percentile_us(0.95) == 10
    c2 = Counters()
    for _ in range(1024):
        c2.record_invoke(9000)
    for _ in range(1024):
        c2.record_invoke(10)
    assert c2.recent_percentile_us(0.95) == 10                              # the slow samples fell out of the 1024-entry ring


def test_counter_map_and_metric_keys():                                      # counters_test.go:78-145, 304-428
    m = CounterMap()
    assert m.snapshot_all() == {} and m.peek_all() == {}
    m.record_invoke("birdnet", 500); m.record_invoke("birdnet", 700); m.record_invoke("bat", 50); m.record_error("new")
    snap = m.snapshot_all()
    assert snap["birdnet"]["invoke_count"] == 2 and snap["birdnet"]["invoke_max_us"] == 700 and snap["bat"]["invoke_total_us"] == 50
    assert snap["new"]["invoke_errors"] == 1 and snap["new"]["invoke_count"] == 0
    assert m.snapshot_all()["birdnet"]["invoke_max_us"] == 0
    peek = m.peek_all()
    assert peek["birdnet"]["invoke_max_us_lifetime"] == 700 and peek["birdnet"]["recent_p95_us"] == 700
    m.delete("bat")
    assert "bat" not in m.peek_all()
    assert sanitize_model_id("BirdNET_V2.4-fp32") == "BirdNET_V2_4_fp32"
    assert metric_key("a.b") == "inference.a_b.avg_ms" and rtf_metric_key("a") == "inference.a.rtf"
    assert throughput_metric_key("a") == "inference.a.throughput" and error_rate_metric_key("a") == "inference.a.error_rate"


def test_counters_are_consistent_under_concurrent_writers():                 # counters_test.go:57-76, 113-138
    m = CounterMap()

    def w(k):
        for i in range(500):
            m.record_invoke("m%d" % (k % 2), i)
            if i % 50 == 0:
                m.record_error("m%d" % (k % 2))
    ts = [threading.Thread(target=w, args=(k,)) for k in range(8)]
    for t in ts: t.start()
    for t in ts: t.join()
    p = m.peek_all()
    assert p["m0"]["invoke_count"] == p["m1"]["invoke_count"] == 2000 and p["m0"]["invoke_errors"] == 40
    assert p["m0"]["invoke_total_us"] == 4 * sum(range(500))


class _Fake:
    def __init__(self, delay=0.0, fail=False):
        self.delay, self.fail, self.active, self.max_active, self.closed = delay, fail, 0, 0, False
        self._mu = threading.Lock()

    def _enter(self):
        with self._mu:
            self.active += 1; self.max_active = max(self.max_active, self.active)
        time.sleep(self.delay)
        with self._mu:
            self.active -= 1
        if self.fail:
            raise RuntimeError("backend failed")

    def predict(self, sample):
        self._enter()
        return [("species", float(len(sample[0])))]

    def predict_batch(self, windows, k=10):
        self._enter()
        return [[("species", float(i))] * k for i in range(len(windows))]

    def close(self):
        self.closed = True


def test_predict_model_errors_and_counters():                                # orchestrator.go:514-572
    o = Orchestrator()
    a, b = _Fake(), _Fake(fail=True)
    o.register("birdnet", a); o.register("bad", b)
    assert o.predict([[0.0] * 5]) == [("species", 5.0)]                      # Predict = PredictModel(primary)
    with pytest.raises(OrchestratorError, match="unknown model: nope"):
        o.predict_model("nope", [[0.0]])
    with pytest.raises(RuntimeError):
        o.predict_model("bad", [[0.0]])
    out = o.predict_model_batch("birdnet", [[0.0]] * 7, k=3)
    assert len(out) == 7 and len(out[0]) == 3
    p = o.counters.peek_all()
    assert p["birdnet"]["invoke_count"] == 2 and p["birdnet"]["batch_windows"] == 8       # one invoke for the whole batch
    assert p["bad"]["invoke_errors"] == 1 and p["bad"]["invoke_count"] == 0
    o.close_model("birdnet")
    assert a.closed
    with pytest.raises(OrchestratorError, match="model birdnet has been closed"):
        o.predict_model("birdnet", [[0.0]])
    o.delete_model("birdnet")
    assert "birdnet" not in o.counters.peek_all()
    with pytest.raises(OrchestratorError, match="unknown model"):
        o.predict_model("birdnet", [[0.0]])


def test_inference_is_serialized_across_models_and_close_waits_for_it():     # inferenceMu + entry.mu (birdnet_backend_lifecycle_race_test.go)
    o = Orchestrator()
    a, b = _Fake(delay=0.02), _Fake(delay=0.02)
    o.register("a", a); o.register("b", b)
    shared = {"active": 0, "max": 0}
    mu = threading.Lock()
    orig = _Fake._enter

    def tracked(self):
        with mu:
            shared["active"] += 1; shared["max"] = max(shared["max"], shared["active"])
        try:
            orig(self)
        finally:
            with mu:
                shared["active"] -= 1
    _Fake._enter = tracked
    refused = []

    def call(mid):
        try:
            o.predict_model(mid, [[0.0]])
        except OrchestratorError as e:                                         # calls that arrive after the close are refused
            refused.append(str(e))
    try:
        ts = [threading.Thread(target=call, args=("a" if i % 2 else "b",)) for i in range(8)]
        for t in ts: t.start()
        closer = threading.Thread(target=o.close_model, args=("a",))
        closer.start()
        for t in ts:
            t.join()
        closer.join()
    finally:
        _Fake._enter = orig
    assert all("has been closed" in r for r in refused)
    assert shared["max"] == 1                                                 # never two models (or two calls) at once
    assert a.closed and a.active == 0                                          # close happened outside any running predict
