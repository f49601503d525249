"""Generate the committed golden fixtures from the float64 oracle (run in the build container).

    python tests/golden/make_golden.py

Writes tests/golden/birdnet_v24_golden.npz:
  soundscape_logits  [79, 6522] float32  — oracle-fp64 logits, soundscape.wav, window 144000 / step 72000
  soundscape_emb     [79, 1024] float32
  tawnyowl_logits    [5, 6522]  float32  — tawnyowl.wav, non-overlapping chunks
  tawnyowl_emb       [5, 1024]  float32
  zeros_logits       [1, 6522]  float32  — silent chunk (what cmd/benchmark feeds, benchmark.go:100-101)
  fe_tawny0          [96,511,2] float32  — frontend output (tensor 265) of tawnyowl chunk 0
The oracle itself is pinned to the reference's published detections in tests/test_oracle_golden.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import birdnet_oracle as bo  # noqa: E402

o = bo.Oracle(dtype=torch.float64)
x, sr = bo.read_wav(os.path.join(bo.ASSETS, "soundscape.wav"))
assert sr == 48000
sc = bo.slice_chunks(x, 72000)
y, sr = bo.read_wav(os.path.join(bo.ASSETS, "tawnyowl.wav"))
to = bo.slice_chunks(y, 144000)
sl, se = o.predict_batch(sc, with_embeddings=True)
tl, te = o.predict_batch(to, with_embeddings=True)
zl = o.predict_batch(np.zeros((1, bo.NUM_SAMPLES), np.float32))
fe = o.run(to[:1], fetch=(bo.T_FRONTEND_OUT,))[bo.T_FRONTEND_OUT][0]
np.savez_compressed(os.path.join(HERE, "birdnet_v24_golden.npz"),
                    soundscape_logits=sl.astype(np.float32), soundscape_emb=se.astype(np.float32),
                    tawnyowl_logits=tl.astype(np.float32), tawnyowl_emb=te.astype(np.float32),
                    zeros_logits=zl.astype(np.float32), fe_tawny0=fe.astype(np.float32))
print("soundscape", sl.shape, "tawnyowl", tl.shape, "top1 tawny", tl.argmax(1), tl.max(1))
print("zeros top1", zl.argmax(1), zl.max(1), "sum", zl.sum())
print("fe stats", fe.mean(), fe.std(), fe.min(), fe.max())
