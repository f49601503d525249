"""Range-filter meta-model: inference.RangeFilter / BatchRangeFilter
(/root/reference/internal/inference/backend.go:55-76) on the GPU, against the op-by-op oracle of
BirdNET_GLOBAL_6K_V2.4_MData_Model_V2_FP16.tflite and the committed golden vectors (tests/golden/make_range_golden.py).

The reference's tests hold no numeric vector for this model (fakes only): the oracle is the same interpreter that reproduces
the 25 published detections of the main model, extended by DEQUANTIZE / SIN / GREATER / LESS / SELECT_V2."""
import os

import numpy as np
import pytest
import torch

import birdnet_b200 as bb
import birdnet_oracle as bo

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden", "range_filter_golden.npz")


def test_oracle_reproduces_the_committed_vectors():
    g = np.load(GOLD)
    got = bo.RangeOracle().predict_batch(g["inputs"])
    assert got.shape == (12, 6522)
    assert np.abs(got - g["scores"]).max() <= 1e-6
    # week mask: weeks outside (0, 49) switch the week features off -> same scores whatever the week value
    o = bo.RangeOracle()
    a = o.predict_batch([[45.0, -75.0, -1.0], [45.0, -75.0, 0.0], [45.0, -75.0, 55.0]])
    assert np.abs(a[0] - a[1]).max() <= 1e-12 and np.abs(a[0] - a[2]).max() <= 1e-12   # (-0.0 vs +0.0 features: batched BLAS rounding)


def test_graph_matcher_accepts_the_meta_model_and_rejects_others(lib_path):
    """No GPU here: the structural match runs before any device is touched, so the right model ends in NO_DEVICE (or OK on
    a GPU box) and a wrong graph in UNSUPPORTED_MODEL."""
    try:
        rf = bb.B200RangeFilter()
        assert rf.num_species() == 6522
        rf.close()
    except bb.B200Error as e:
        assert e.status == bb.ERR_NO_DEVICE, str(e)
    with pytest.raises(bb.B200Error) as e:
        bb.B200RangeFilter(model_data=open(bb.DEFAULT_MODEL, "rb").read())
    assert e.value.status == bb.ERR_UNSUPPORTED_MODEL
    with pytest.raises(bb.B200Error) as e:
        bb.B200RangeFilter(model_data=b"not a flatbuffer at all")
    assert e.value.status in (bb.ERR_UNSUPPORTED_MODEL, bb.ERR_INVALID_ARGUMENT)


@pytest.mark.gpu
def test_gpu_range_filter_matches_oracle_and_golden(lib_path):
    g = np.load(GOLD)
    rf = bb.B200RangeFilter()
    assert rf.num_species() == 6522
    got = rf.predict_batch(g["inputs"])
    assert np.abs(got - g["scores"]).max() <= 2e-5
    one = rf.predict(*g["inputs"][0])
    assert np.array_equal(one, got[0])                        # Predict == PredictBatch row (backend.go:58-76)
    # a heat-map sized batch (orchestrator.go:1846-1882 walks a lat/lon grid): compared with the oracle live
    rng = np.random.default_rng(3)
    x = np.stack([rng.uniform(-90, 90, 300), rng.uniform(-180, 180, 300), rng.integers(-1, 51, 300)], axis=1).astype(np.float32)
    ref = bo.RangeOracle(dtype=torch.float64).predict_batch(x)
    got = rf.predict_batch(x)
    assert np.abs(got - ref).max() <= 2e-5
    assert rf.predict_batch(np.zeros((0, 3), np.float32)).shape == (0, 6522)
    with pytest.raises(bb.B200Error):
        rf.predict_batch(np.zeros(7, np.float32))
    rf.close(); rf.close()
