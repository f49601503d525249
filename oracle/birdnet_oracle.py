"""ORACLE (test infrastructure, never shipped): BirdNET v2.4 hot path on the CPU.

Restates, around the graph interpretation in tflite_interp.py, the reference's pre/post steps:

* int16 LE PCM -> float32 / 32768             /root/reference/internal/analysis/process.go:479-497
* window slicing (3 s window, step 3-overlap)  /root/reference/internal/audiocore/buffer/analysis.go:187-251,
                                               /root/reference/internal/classifier/model.go:35-53,
                                               /root/reference/doc/wiki/file-analysis.md:1-13
* inference.Classifier.Predict -> raw logits   /root/reference/internal/inference/backend.go:8-19,
                                               /root/reference/internal/inference/tflite/classifier.go:95-119
* sigmoid(sensitivity * x) in float64          /root/reference/internal/classifier/analyze.go:113-115,197-208
* top-k (k = 10), descending confidence        /root/reference/internal/classifier/analyze.go:220-253

PARITY PINNING: the reference's tests hold no numeric golden for this path (every test uses
fake backends); the only published known-answer data is the 25-row detection table in
/root/reference/doc/wiki/file-analysis.md:20-44 produced by the reference TFLite FP32 path.
GOLDEN_TABLE below restates it; tests/test_oracle_golden.py checks the oracle against all rows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import this.
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tflite_interp import Interpreter  # noqa: E402
from tflite_reader import load  # noqa: E402

SAMPLE_RATE = 48000            # internal/conf/consts.go:14-24
NUM_SAMPLES = 144000           # 3 s window
NUM_SPECIES = 6522
EMBEDDING_DIM = 1024
TOP_K = 10                     # internal/classifier/tracing.go:59 (defaultTopKResults)

# tensor indices in the v2.4 graph
T_FRONTEND_OUT = 265           # [B,96,511,2] after the folded-BN affine
T_STEM_OUT = 266
T_EMBEDDING = 545              # model/GLOBAL_AVG_POOL/Mean
T_LOGITS = 546

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(REPO, "assets")
MODEL_PATH = os.path.join(ASSETS, "BirdNET_GLOBAL_6K_V2.4_Model_FP32.tflite")
LABELS_PATH = os.path.join(ASSETS, "BirdNET_GLOBAL_6K_V2.4_Labels_en_us.txt")

# doc/wiki/file-analysis.md:20-44 — (begin time s, common name, confidence); threshold 0.1,
# sensitivity 1.5 (confidence = sigmoid(1.5 * logit) of the top-1 species of that 3 s chunk).
GOLDEN_TABLE = [
    (0.0, "Black-capped Chickadee", 0.9016), (3.0, "Black-capped Chickadee", 0.2293),
    (9.0, "House Finch", 0.7025), (18.0, "Blue Jay", 0.4036), (21.0, "Blue Jay", 0.2557),
    (27.0, "Merlin", 0.1787), (33.0, "Dark-eyed Junco", 0.4388), (36.0, "Dark-eyed Junco", 0.2882),
    (39.0, "House Finch", 0.1649), (42.0, "Dark-eyed Junco", 0.8249), (51.0, "House Finch", 0.1684),
    (54.0, "House Finch", 0.6576), (60.0, "Dark-eyed Junco", 0.5821), (69.0, "House Finch", 0.2088),
    (72.0, "House Finch", 0.5951), (78.0, "House Finch", 0.1680), (81.0, "Hawfinch", 0.1956),
    (84.0, "House Finch", 0.1953), (90.0, "American Goldfinch", 0.3799), (93.0, "House Finch", 0.2400),
    (96.0, "American Goldfinch", 0.4648), (102.0, "House Finch", 0.4192), (111.0, "Engine", 0.5252),
    (114.0, "Engine", 0.1538), (117.0, "American Goldfinch", 0.3417),
]
GOLDEN_SENSITIVITY = 1.5


def read_labels(path=LABELS_PATH):
    with open(path, encoding="utf-8") as f:
        return [ln.rstrip("\n") for ln in f if ln.strip()]


def read_wav(path):
    """PCM WAV (16/24/32-bit int, incl. WAVE_FORMAT_EXTENSIBLE) -> (float32 mono [-1,1), rate).

    16-bit follows process.go:491-494 (int16 / 32768 in float32); 24/32-bit follow the same
    rule with 2^23 / 2^31 (internal/audiocore/convert/pcm.go:206-270)."""
    with open(path, "rb") as f:
        b = f.read()
    if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    p, fmt, data = 12, None, None
    while p + 8 <= len(b):
        cid, sz = b[p:p + 4], struct.unpack_from("<I", b, p + 4)[0]
        if cid == b"fmt ":
            fmt = struct.unpack_from("<HHIIHH", b, p + 8)
        elif cid == b"data":
            data = b[p + 8:p + 8 + sz]
        p += 8 + sz + (sz & 1)
    tag, ch, rate, _, _, bits = fmt
    if bits == 16:
        x = np.frombuffer(data, "<i2").astype(np.float32) / np.float32(32768.0)
    elif bits == 24:
        r = np.frombuffer(data, np.uint8).reshape(-1, 3).astype(np.int32)
        v = r[:, 0] | (r[:, 1] << 8) | (r[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = v.astype(np.float32) / np.float32(8388608.0)
    elif bits == 32:
        x = np.frombuffer(data, "<i4").astype(np.float32) / np.float32(2147483648.0)
    else:
        raise ValueError("unsupported bit depth %d" % bits)
    if ch > 1:
        x = x.reshape(-1, ch)[:, 0].copy()
    return x, rate


def slice_chunks(x, step, window=NUM_SAMPLES):
    """All full windows advanced by `step` samples (72000 = 1.5 s overlap, 144000 = none)."""
    n = 0 if len(x) < window else (len(x) - window) // step + 1
    if n == 0:
        return np.zeros((0, window), np.float32)
    return np.stack([x[i * step:i * step + window] for i in range(n)]).astype(np.float32)


def sigmoid_sensitivity(logits, sensitivity=1.0):
    """analyze.go:113-115,197-208: float32 -> float64 math.Exp -> float32."""
    z = np.asarray(logits, np.float32).astype(np.float64)
    return (1.0 / (1.0 + np.exp(-float(sensitivity) * z))).astype(np.float32)


def top_k(conf, k=TOP_K):
    """analyze.go:220-253: k highest confidences, descending.  (The reference's quick-select is
    unstable among exact ties; ties are broken here by lower label index.)"""
    conf = np.asarray(conf)
    order = np.lexsort((np.arange(conf.shape[-1]), -conf.astype(np.float64)))[:k] if conf.ndim == 1 else None
    if conf.ndim == 1:
        return order.astype(np.int32), conf[order]
    idx = np.stack([top_k(c, k)[0] for c in conf])
    return idx, np.take_along_axis(conf, idx.astype(np.int64), axis=1)


class Oracle:
    """inference.Classifier-shaped CPU restatement: Predict(samples) -> raw logits."""

    def __init__(self, model_path=MODEL_PATH, dtype=torch.float64, threads=None):
        self.graph = load(model_path)
        self.interp = Interpreter(self.graph, dtype)
        self.threads = threads

    def num_species(self):
        return NUM_SPECIES

    def run(self, chunks, fetch=(), batch=8):
        """chunks [B,144000] float32 -> dict {tensor_index: np.ndarray} (float64 or float32)."""
        if self.threads:
            torch.set_num_threads(self.threads)
        chunks = np.ascontiguousarray(chunks, np.float32)
        if chunks.ndim != 2 or chunks.shape[1] != NUM_SAMPLES:
            raise ValueError("input size mismatch: expected %d samples, got %s" % (NUM_SAMPLES, chunks.shape[1:]))
        want = list(dict.fromkeys(list(fetch) + [T_LOGITS]))
        acc = {k: [] for k in want}
        for i in range(0, len(chunks), batch):
            r = self.interp.run(torch.from_numpy(chunks[i:i + batch]), fetch=fetch)
            for k in want:
                acc[k].append(r[k].numpy())
        return {k: np.concatenate(v) if v else np.zeros((0,)) for k, v in acc.items()}

    def predict(self, samples):
        """Batch-1 drop-in semantics of tflite/classifier.go:95-119 (raw logits, float32)."""
        samples = np.asarray(samples, np.float32)
        if samples.shape != (NUM_SAMPLES,):
            raise ValueError("input size mismatch: expected %d samples, got %d" % (NUM_SAMPLES, samples.size))
        return self.run(samples[None])[T_LOGITS][0].astype(np.float32)

    def predict_batch(self, chunks, with_embeddings=False, batch=8):
        r = self.run(chunks, fetch=(T_EMBEDDING,) if with_embeddings else (), batch=batch)
        if with_embeddings:
            return r[T_LOGITS], r[T_EMBEDDING].reshape(len(chunks), -1)
        return r[T_LOGITS]


# ---------------------------------------------------------------------------------------------------------------------
# Range filter ("meta" model): restates what the reference's tfliteRangeFilter does
# (/root/reference/internal/inference/tflite/rangefilter.go:64-94): write (latitude, longitude, week) into the input
# tensor, Invoke, read NumSpecies scores.  The arithmetic is the .tflite graph itself, interpreted op by op.
RANGE_MODEL = os.path.join(ASSETS, "BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite")


class RangeOracle:
    def __init__(self, path=RANGE_MODEL, dtype=None):
        import torch
        from tflite_interp import Interpreter, load
        self.g = load(path)
        self.it = Interpreter(self.g, dtype=dtype or torch.float64)

    def predict_batch(self, inputs):
        """inputs [B,3] (lat, lon, week) -> [B, n_species] float64 scores."""
        import torch
        x = torch.as_tensor(np.asarray(inputs, np.float32).reshape(-1, 3))
        return self.it.run(x)[self.g.outputs[0]].numpy()
