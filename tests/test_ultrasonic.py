"""N2: ultrasonic frame-power CV filter.  CPU part: the oracle against the reference's own test cases
(/root/reference/internal/audiocore/ultrasonic/filter_test.go, restated one to one).  GPU part: bnb_ultrasonic_cv_batch
(float64 STFT kernels) against the oracle, tolerance 1e-9 relative on the CV."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import ultrasonic_oracle as uo

RATE, N = 256000, 144000


def _tone(freq=40000.0, amp=0.01):                       # filter_test.go:23-42
    t = np.arange(N, dtype=np.float64) / RATE
    return amp * np.sin(2 * math.pi * freq * t)


def _burst(freq=45000.0, amp=0.5):                       # filter_test.go:44-65
    x = np.zeros(N)
    i = np.arange(N // 3, 2 * N // 3)
    x[i] = amp * np.sin(2 * math.pi * freq * (i / RATE))
    return x


def test_flat_tone_low_cv_and_burst_high_cv():
    cv, ok = uo.compute_us_frame_cv(_tone(), RATE)
    assert ok and cv < 0.15
    cv, ok = uo.compute_us_frame_cv(_burst(), RATE)
    assert ok and cv > 0.15


def test_not_ok_cases():
    assert uo.compute_us_frame_cv(np.zeros(100), RATE) == (0.0, False)                                   # :67-73 insufficient data
    assert uo.compute_us_frame_cv(np.zeros(N), 30000, frequency_split_hz=20000) == (0.0, False)         # :75-83 split >= nyquist
    assert uo.compute_us_frame_cv(np.zeros(N), RATE, fft_size=1000) == (0.0, False)                     # :85-92 not a power of two
    assert uo.compute_us_frame_cv(np.zeros(8192 + 100), RATE) == (0.0, False)                            # :94-105 one frame only
    assert uo.is_unlikely(0.1) and not uo.is_unlikely(0.2) and not uo.is_unlikely(0.15)                  # :107-115


def test_window_fft_cv_building_blocks():
    w = uo.hanning_window(8)                                                                             # :117-135
    assert abs(w[0]) < 1e-10 and abs(w[7]) < 1e-10 and np.allclose(w, w[::-1], atol=1e-12) and w.max() <= 1.0
    n, k = 64, 10                                                                                        # :137-162
    x = np.sin(2 * math.pi * k * np.arange(n) / n)
    z = uo.fft_reference(x)
    assert int(np.argmax(np.abs(z[: n // 2]))) == k
    rng = np.random.default_rng(0)
    y = rng.standard_normal(1024)
    assert np.abs(uo.fft_reference(y) - np.fft.fft(y)).max() < 1e-10                                      # the literal FFT is an FFT
    assert abs(uo.coefficient_of_variation([5, 5, 5, 5])) < 1e-10                                         # :164-178
    assert abs(uo.coefficient_of_variation([1, 2, 3]) - math.sqrt(2.0 / 3.0) / 2.0) < 1e-10
    assert uo.coefficient_of_variation([1]) == 0.0 and uo.coefficient_of_variation([]) == 0.0


def test_scale_invariance_and_literal_fft_agree():
    b = _burst()
    cv1, _ = uo.compute_us_frame_cv(b, RATE)                                                              # :180-206
    cv2, _ = uo.compute_us_frame_cv(0.01 * b, RATE)
    assert abs(cv1 - cv2) < 0.01
    short = b[40000: 40000 + 3 * 2048]
    a, ok_a = uo.compute_us_frame_cv(short, RATE, fft_size=2048, hop_size=1024, literal_fft=True)
    f, ok_f = uo.compute_us_frame_cv(short, RATE, fft_size=2048, hop_size=1024)
    assert ok_a and ok_f and abs(a - f) <= 1e-9 * max(1.0, abs(f))


def test_pcm16_conversion():
    raw = np.array([0, 1, -1, 32767, -32768], dtype="<i2").tobytes() + b"\x7f"                            # trailing odd byte ignored
    assert np.array_equal(uo.bytes_to_float64_pcm16(raw), np.array([0, 1, -1, 32767, -32768]) / 32768.0)


@pytest.mark.gpu
def test_gpu_filter_matches_oracle(lib_path):
    import birdnet_b200 as bb
    rng = np.random.default_rng(7)
    chunks = [_tone(), _burst(), _burst(60000.0, 0.2) + 0.01 * rng.standard_normal(N), 0.05 * rng.standard_normal(N), np.zeros(N)]
    x16 = np.stack([np.clip(np.round(c * 32768.0), -32768, 32767).astype(np.int16) for c in chunks])
    cv, ok = bb.ultrasonic_cv_batch(x16, RATE)
    assert ok.all()
    for b in range(len(chunks)):
        ref, ok_ref = uo.compute_us_frame_cv(x16[b].astype(np.float64) / 32768.0, RATE)
        assert ok_ref and abs(cv[b] - ref) <= 1e-9 * max(1.0, abs(ref)), (b, cv[b], ref)
    assert cv[0] < 0.15 < cv[1] and cv[4] == 0.0                                     # tone: unlikely; burst: kept; silence: mean 0 -> CV 0
    # float32 input, other STFT geometry
    xf = np.stack(chunks[:3]).astype(np.float32)
    cv2, ok2 = bb.ultrasonic_cv_batch(xf, RATE, fft_size=2048, hop_size=512, frequency_split_hz=30000)
    for b in range(3):
        ref, _ = uo.compute_us_frame_cv(xf[b].astype(np.float64), RATE, fft_size=2048, hop_size=512, frequency_split_hz=30000)
        assert ok2[b] and abs(cv2[b] - ref) <= 1e-9 * max(1.0, abs(ref))
    # the reference's (0, false) cases come back as ok = False, cv = 0, status OK
    for kw in (dict(fft_size=1000), dict(frequency_split_hz=128000), dict(hop_size=0)):
        c, o = bb.ultrasonic_cv_batch(x16[:2], RATE, **kw)
        assert not o.any() and (c == 0).all()
    c, o = bb.ultrasonic_cv_batch(x16[:1, :100], RATE)
    assert not o.any()


@pytest.mark.gpu
def test_gpu_dense_head_matches_numpy(lib_path):
    import birdnet_b200 as bb
    rng = np.random.default_rng(3)
    e = rng.standard_normal((37, 1024)).astype(np.float32)
    w = (0.05 * rng.standard_normal((19, 1024))).astype(np.float32)
    b = rng.standard_normal(19).astype(np.float32)
    got = bb.dense_head_batch(e, w, b)
    ref = 1.0 / (1.0 + np.exp(-(e.astype(np.float64) @ w.astype(np.float64).T + b)))
    assert np.abs(got - ref).max() <= 2e-6
