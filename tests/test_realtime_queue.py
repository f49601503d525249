"""Chunk-queue host logic (BASELINE config 5): AnalysisBuffer semantics as the reference tests them
(internal/audiocore/buffer/analysis_test.go:23-300) and the coalescing dispatcher; GPU part: 8 synthetic int16 streams."""
import numpy as np
import pytest

from birdnet_b200.realtime import AnalysisBuffer, RealtimeCoalescer, WINDOW_BYTES


def test_analysis_buffer_overlap_and_content_parity():
    ab = AnalysisBuffer(64, 4, 8, "src")
    assert ab.read() is None                                   # try again later
    ab.write(bytes(range(1, 17)))
    w1, w2 = ab.read(), ab.read()
    assert list(w1) == [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8]    # zero prefix on the first window
    assert list(w2) == [5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    assert ab.read() is None
    big = AnalysisBuffer(3 * WINDOW_BYTES, 144000, 144000, "mic")
    pcm = (np.arange(WINDOW_BYTES) * 7 % 251).astype(np.uint8)
    big.write(pcm.tobytes())
    a, b = big.read(), big.read()
    assert len(a) == WINDOW_BYTES and np.array_equal(b[:144000], a[144000:])
    small = AnalysisBuffer(8, 0, 4, "s")
    small.write(bytes(range(6))); small.write(bytes(range(6, 12)))
    assert small.overwrites >= 1 and list(small.read()) == [4, 5, 6, 7]
    for args in ((10, 8, 4, "s"), (2, 0, 4, "s"), (10, 0, 4, ""), (0, 0, 4, "s"), (10, -1, 4, "s"), (10, 0, 0, "s")):
        with pytest.raises(ValueError):
            AnalysisBuffer(*args)


def test_coalescer_batches_ready_windows_and_counts_drops():
    calls = []

    def fake_analyze(pcm, sens, k):
        calls.append(pcm.shape)
        assert pcm.dtype == np.dtype("<i2") and pcm.shape[1] == 144000
        idx = np.tile(np.arange(k, dtype=np.int32), (len(pcm), 1)) + pcm[:, :1].astype(np.int32) % 7
        return idx, np.linspace(0.9, 0.1, k, dtype=np.float32)[None].repeat(len(pcm), 0)

    labels = ["sp%d" % i for i in range(6522)]
    rc = RealtimeCoalescer(fake_analyze, labels, ["s%d" % i for i in range(8)], queue_size=10)
    frame = np.zeros(960, "<i2")                                 # 20 ms frames
    for t in range(75):                                          # 1.5 s: one fresh half-window per source
        for i in range(8):
            frame[:] = i + 1
            rc.write("s%d" % i, frame.tobytes())
    assert rc.tick() == 8 and calls == [(8, 144000)]            # ONE batched call for the eight sources
    assert rc.tick() == 0
    for t in range(75):
        rc.write("s3", frame.tobytes())
    assert rc.tick() == 1 and calls[-1] == (1, 144000)
    assert len(rc.queue) == 9 and rc.dropped == 0
    for t in range(150):
        for i in range(8):
            rc.write("s%d" % i, frame.tobytes())
    rc.tick(); rc.tick()
    assert len(rc.queue) == 10 and rc.dropped > 0                # non-blocking enqueue with drop accounting
    r = rc.queue[0]
    assert r.source_id == "s0" and len(r.pcm) == WINDOW_BYTES and r.results[0][1] == pytest.approx(0.9)


@pytest.mark.gpu
def test_eight_streams_through_the_gpu(lib_path):
    import birdnet_b200 as bb
    import birdnet_oracle as bo
    from bench import synth_chunks
    clf = bb.B200Classifier(max_batch=16)
    labels = bo.read_labels()
    rc = RealtimeCoalescer(lambda p, s, k: clf.analyze_batch(p, s, k), labels, ["rtsp%d" % i for i in range(8)])
    streams = (synth_chunks(16, seed0=777).reshape(8, -1) * 32768.0).astype("<i2")     # 8 streams x 6 s
    for t in range(0, streams.shape[1], 960):
        for i in range(8):
            rc.write("rtsp%d" % i, streams[i, t:t + 960].tobytes())
        if (t // 960) % 5 == 4:                                   # 100 ms monitor tick (buffer_manager.go:392)
            rc.tick()
    assert rc.windows == 8 * 4 and rc.batches == 4               # 6 s at 1.5 s hops -> 4 windows per stream, coalesced 8-wide
    # last window of stream 0 == samples [72000*2, 72000*2+144000) of that stream, analysed alone
    w = streams[0, 72000 * 2:72000 * 2 + 144000][None]
    idx, conf = clf.analyze_batch(w, 1.0, 10)
    last = [r for r in rc.queue if r.source_id == "rtsp0"][-1]
    assert [labels[i] for i in idx[0]] == [s for s, _ in last.results]
    assert np.allclose([c for _, c in last.results], conf[0], atol=1e-6)
    clf.close()


def test_buffer_overrun_tracking_tumbling_window():
    """process.go:351-370 + :177-215: an inference slower than the buffer interval counts as an overrun per source; a report is
    emitted once per expired window when at least 10 overruns accumulated."""
    import numpy as np
    from birdnet_b200 import realtime as rt

    slow = {"dt": 0.0}

    def analyze(pcm, sens, k):
        return np.zeros((len(pcm), k), np.int32), np.zeros((len(pcm), k), np.float32)

    co = rt.RealtimeCoalescer(analyze, ["x"] * 6522, ["a", "b"])
    # fast inference: no overrun
    for s in ("a", "b"):
        co.write(s, b"\0" * 144000)
    co.tick(now=0.0)
    assert co.overrun_total == 0 and not co.overruns
    # force the measured inference time above the interval by patching the clock the coalescer uses
    real = rt.time.perf_counter
    ticks = iter([0.0, 10.0] * 1000)
    rt.time.perf_counter = lambda: next(ticks)
    try:
        t = 1.0
        for _ in range(12):
            co.write("a", b"\0" * 144000)
            co.tick(now=t); t += 1.5
        tr = co.overruns["a:BirdNET_V2.4"]
        assert tr.overrun_count == 12 and not tr.reports and abs(tr.max_elapsed - 10.0) < 1e-9 and abs(tr.buffer_length - 1.5) < 1e-9
        co.write("a", b"\0" * 144000)
        co.tick(now=1.0 + rt.BUFFER_OVERRUN_REPORT_COOLDOWN_S + 1)         # window expired with >= 10 overruns: one report, counters reset
        assert len(tr.reports) == 1 and tr.reports[0]["overrun_count"] == 12 and tr.overrun_count == 1
        assert "b:BirdNET_V2.4" not in co.overruns
    finally:
        rt.time.perf_counter = real
