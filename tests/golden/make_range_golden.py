"""Golden vectors of the range-filter meta-model from the float64 oracle (run in the build container).

    python tests/golden/make_range_golden.py   ->  tests/golden/range_filter_golden.npz {inputs [N,3], scores [N,6522] float32}

Inputs: a fixed grid of (latitude, longitude, week) covering both hemispheres, the date line, the week mask edges
(week <= 0 and week >= 49 switch the week features off, as the reference does for "no week") and fractional weeks."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import birdnet_oracle as bo  # noqa: E402

pts = [(60.17, 24.94, 20.0), (-33.9, 151.2, 48.0), (0.0, 0.0, 1.0), (45.0, -75.0, -1.0), (89.9, 179.9, 49.0), (-89.9, -179.9, 0.0),
       (35.68, 139.69, 26.5), (51.5, -0.12, 48.99), (-1.29, 36.82, 13.0), (64.15, -21.94, 50.0), (19.43, -99.13, 7.25), (-54.8, -68.3, 40.0)]
inputs = np.asarray(pts, np.float32)
scores = bo.RangeOracle().predict_batch(inputs)
np.savez_compressed(os.path.join(HERE, "range_filter_golden.npz"), inputs=inputs, scores=scores.astype(np.float32))
print(scores.shape, scores.max(1), (scores > 0.03).sum(1))
