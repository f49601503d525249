"""Golden vectors for BASELINE config 3 (synthetic 48 kHz pink noise + chirp): the float64 oracle's ten best (index, confidence)
per chunk for the fixed 512-chunk subset seed0 = 1234.   python tests/golden/make_synth_golden.py
Writes tests/golden/synth512_golden.npz {top10_idx [512,10] int32, top10_conf [512,10] float32 (sigmoid, sensitivity 1.0)}."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import birdnet_oracle as bo  # noqa: E402
from bench import synth_chunks  # noqa: E402

torch.set_num_threads(max(1, os.cpu_count() or 1))
x = synth_chunks(512, seed0=1234)
o = bo.Oracle(dtype=torch.float64)
logits = np.concatenate([o.predict_batch(x[i:i + 16], batch=16) for i in range(0, 512, 16)])
conf = bo.sigmoid_sensitivity(logits, 1.0)
idx, top = bo.top_k(conf, 10)
np.savez_compressed(os.path.join(HERE, "synth512_golden.npz"), top10_idx=idx.astype(np.int32), top10_conf=top.astype(np.float32))
print(idx.shape, top[:, 0].min(), top[:, 0].max())
