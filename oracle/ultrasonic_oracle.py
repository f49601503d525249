"""ultrasonic_oracle.py — CPU restatement of the bat pipeline's ultrasonic validation filter (TEST INFRASTRUCTURE ONLY: imported
by tests/ and nothing else; the product path is birdnet-go_b200/csrc/ultrasonic.cu).

Follows /root/reference/internal/audiocore/ultrasonic/filter.go line by line:
  compute_us_frame_cv   ComputeUSFrameCV            filter.go:20-66
  is_unlikely           IsUnlikely                  filter.go:71-73
  coefficient_of_variation                          filter.go:76-97 (sequential float64 sums, population variance)
  fft_reference         fft (iterative radix-2 with the RUNNING twiddle w *= wn)   filter.go:101-135
  hanning_window        hanningWindow (symmetric: 0.5 (1 - cos(2 pi i / (n - 1))))  filter.go:139-145
  bytes_to_float64_pcm16  convert.BytesToFloat64PCM16 (int16 little endian / 32768)   internal/audiocore/convert/pcm.go:92-113
Parity pin: the reference holds no numeric vectors for this function — its tests (filter_test.go:23-206) are properties (flat
tone below / burst above the 0.15 threshold, the ok = false cases, window end points, FFT peak bin, CV closed forms, scale
invariance); tests/test_ultrasonic.py restates every one of them against this file, and checks fft_reference against numpy.fft.
"""
from __future__ import annotations

import math

import numpy as np

DEFAULTS = dict(cv_threshold=0.15, fft_size=8192, hop_size=4096, frequency_split_hz=20000)   # filter_test.go:13-21, conf/defaults.go:108-112


def bytes_to_float64_pcm16(b: bytes) -> np.ndarray:
    even = len(b) & ~1
    return np.frombuffer(b[:even], dtype="<i2").astype(np.float64) / 32768.0


def hanning_window(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * math.pi * i / float(n - 1)))


def fft_reference(data: np.ndarray) -> np.ndarray:
    """The reference's in-place FFT, including its running-twiddle recurrence (vectorised over the independent blocks of a stage)."""
    x = np.array(data, dtype=np.complex128)
    n = x.shape[0]
    if n <= 1:
        return x
    bits = n.bit_length() - 1
    rev = np.zeros(n, dtype=np.int64)
    for b in range(bits):
        rev |= ((np.arange(n) >> b) & 1) << (bits - 1 - b)
    x = x[rev]
    size = 2
    while size <= n:
        half = size >> 1
        wn = complex(math.cos(-2.0 * math.pi / size), math.sin(-2.0 * math.pi / size))     # cmplx.Rect(1, -2 pi / size)
        blocks = x.reshape(n // size, size)
        w = complex(1.0, 0.0)
        for k in range(half):
            u = blocks[:, k].copy()
            v = w * blocks[:, k + half]
            blocks[:, k] = u + v
            blocks[:, k + half] = u - v
            w *= wn
        size <<= 1
    return x


def coefficient_of_variation(values) -> float:
    n = float(len(values))
    if n < 2:
        return 0.0
    s = 0.0
    for v in values:
        s += float(v)
    mean = s / n
    if mean <= 0:
        return 0.0
    sq = 0.0
    for v in values:
        d = float(v) - mean
        sq += d * d
    return math.sqrt(sq / n) / mean


def compute_us_frame_cv(samples: np.ndarray, sample_rate: int, fft_size=8192, hop_size=4096, frequency_split_hz=20000, literal_fft=False):
    """-> (cv, ok).  literal_fft=True uses fft_reference (slow); the default numpy.fft differs from it by ~1e-13 relative."""
    samples = np.asarray(samples, dtype=np.float64)
    if len(samples) < fft_size or sample_rate <= 0 or fft_size < 2 or hop_size <= 0:
        return 0.0, False
    if fft_size & (fft_size - 1):
        return 0.0, False
    if frequency_split_hz < 0 or frequency_split_hz >= sample_rate // 2:
        return 0.0, False
    window = hanning_window(fft_size)
    n_frames = 1 + (len(samples) - fft_size) // hop_size
    if n_frames < 2:
        return 0.0, False
    bin_width = float(sample_rate) / float(fft_size)
    split_bin = int(float(frequency_split_hz) / bin_width)
    nyq = fft_size // 2
    powers = []
    for f in range(n_frames):
        seg = samples[f * hop_size: f * hop_size + fft_size] * window
        z = fft_reference(seg) if literal_fft else np.fft.fft(seg)
        p = (z.real ** 2 + z.imag ** 2)[split_bin: nyq + 1].copy()
        lo = max(split_bin, 1)
        p[lo - split_bin: nyq - split_bin] *= 2.0            # bins with 0 < bin < nyquist count twice
        power = 0.0
        for v in p:                                           # sequential float64 sum like the Go loop
            power += float(v)
        powers.append(power)
    return coefficient_of_variation(powers), True


def is_unlikely(cv: float, cv_threshold: float = 0.15) -> bool:
    return cv < cv_threshold
