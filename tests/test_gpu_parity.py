"""GPU: parity of the CUDA path (through the C ABI) against the float64 oracle and the golden fixtures.

Tolerance (BASELINE.json north_star): max |sigmoid(gpu) - sigmoid(oracle)| <= 1e-3 and equal
top-1; raw-logit differences are reported (two correct fp32 implementations differ by ~1.5e-3 there)."""
import json
import os

import numpy as np
import pytest
import torch

import birdnet_b200 as bb
import birdnet_oracle as bo

pytestmark = pytest.mark.gpu
SIG_TOL = 1e-3
PRECISIONS = [bb.PRECISION_F32, bb.PRECISION_F16X3]


def _sig(x):
    return bo.sigmoid_sensitivity(x, 1.0).astype(np.float64)


@pytest.fixture(scope="module", params=PRECISIONS, ids=["f32", "f16x3"])
def clf(request, lib_path):
    c = bb.B200Classifier(max_batch=128, micro_batch=16, precision=request.param)
    yield c
    c.close()


def test_runtime_info(clf):
    dev, backend, prec = clf.runtime_info()
    assert dev.startswith("CUDA:") and "B200" in dev
    assert clf.num_species() == 6522 and clf.n_samples == 144000 and clf.emb_dim == 1024


def test_soundscape_sliding_window_parity(clf, golden, audio):
    """BASELINE config 2: soundscape.wav, 3 s window / 1.5 s overlap -> 79 chunks."""
    chunks = bo.slice_chunks(audio["soundscape"], 72000)
    assert chunks.shape[0] == 79
    logits, emb = clf.predict_batch(chunks, with_embeddings=True)
    ref = golden["soundscape_logits"].astype(np.float64)
    ds = np.abs(_sig(logits) - _sig(ref)).max()
    dl = np.abs(logits - ref).max()
    print("soundscape: max|dsigmoid|=%.3e max|dlogit|=%.3e" % (ds, dl))
    assert ds <= SIG_TOL
    assert (logits.argmax(1) == ref.argmax(1)).all()
    assert np.abs(emb - golden["soundscape_emb"]).max() <= 2e-2 * np.abs(golden["soundscape_emb"]).max()


def test_published_detections_through_the_gpu(clf, audio):
    """The reference's own table (doc/wiki/file-analysis.md:20-44) reproduced by the CUDA path."""
    labels = bo.read_labels()
    chunks = bo.slice_chunks(audio["soundscape"], 144000)
    idx, conf = clf.analyze_batch(chunks, sensitivity=bo.GOLDEN_SENSITIVITY, k=10)
    for t0, name, c in bo.GOLDEN_TABLE:
        i = int(round(t0 / 3.0))
        assert labels[idx[i, 0]].split("_", 1)[1] == name
        assert abs(float(conf[i, 0]) - c) <= 1e-3, (t0, conf[i, 0], c)


def test_tawnyowl_batch1_dropin(clf, golden, audio):
    """BASELINE config 1: batch-1 Predict() on tawnyowl.wav chunk by chunk."""
    chunks = bo.slice_chunks(audio["tawnyowl"], 144000)
    for i in range(len(chunks)):
        lg, emb = clf.predict_with_embeddings(chunks[i])
        ref = golden["tawnyowl_logits"][i].astype(np.float64)
        assert np.abs(_sig(lg) - _sig(ref)).max() <= SIG_TOL
        assert int(lg.argmax()) == int(ref.argmax())
    assert int(clf.predict(chunks[0]).argmax()) == 5760      # Strix aluco_Tawny Owl


def test_silence_and_edge_inputs(clf, golden):
    z = clf.predict(np.zeros(144000, np.float32))
    assert np.abs(_sig(z) - _sig(golden["zeros_logits"][0])).max() <= SIG_TOL and int(z.argmax()) == 2143
    # constant (max == min), full-scale square wave, single impulse: compared with the oracle live
    rng = np.random.default_rng(5)
    x = np.zeros((4, 144000), np.float32)
    x[0] = 0.25
    x[1] = np.where(np.arange(144000) % 96 < 48, 1.0, -1.0)
    x[2, 70000] = 1.0
    x[3] = np.clip(rng.standard_normal(144000) * 0.5, -1, 1)
    ref = bo.Oracle(dtype=torch.float64).predict_batch(x)
    got = clf.predict_batch(x)
    assert np.isfinite(got).all()
    assert np.abs(_sig(got) - _sig(ref)).max() <= SIG_TOL


def test_int16_ingest_equals_float_path(clf, audio):
    chunks = bo.slice_chunks(audio["soundscape"], 144000)[:6]
    pcm16 = np.round(chunks * 32768.0).astype(np.int16)        # soundscape.wav is 16-bit: exact round trip
    assert np.array_equal(pcm16.astype(np.float32) / np.float32(32768), chunks)
    a = clf.predict_batch(chunks)
    b = clf.predict_batch(pcm16)
    assert np.array_equal(a, b)


def test_batch_composition_is_irrelevant(clf, audio):
    """Chunks are independent: micro-batch boundaries, batch size and order must not change a single bit."""
    chunks = bo.slice_chunks(audio["soundscape"], 72000)[:37]
    full = clf.predict_batch(chunks)
    one = np.stack([clf.predict(c) for c in chunks[:3]])
    assert np.array_equal(full[:3], one)
    perm = np.random.default_rng(1).permutation(len(chunks))
    assert np.array_equal(clf.predict_batch(chunks[perm]), full[perm])
    assert clf.predict_batch(chunks[:0]).shape == (0, 6522)


def test_topk_matches_reference_postprocessing(clf, audio):
    chunks = bo.slice_chunks(audio["soundscape"], 144000)[:8]
    idx, conf, logits = clf.analyze_batch(chunks, sensitivity=1.5, k=10, want_logits=True)
    want_conf = bo.sigmoid_sensitivity(logits, 1.5)
    widx, wconf = bo.top_k(want_conf, 10)
    assert np.array_equal(idx, widx)
    assert np.abs(conf - wconf).max() <= 1e-7
    assert (np.diff(conf, axis=1) <= 0).all()


def test_error_behaviour(clf):
    with pytest.raises(bb.B200Error) as e:
        clf.predict(np.zeros(1000, np.float32))
    assert e.value.status == bb.ERR_INVALID_ARGUMENT and "input size mismatch" in str(e.value)
    with pytest.raises(bb.B200Error):
        clf.predict_batch(np.zeros((clf.max_batch + 1, 144000), np.float32))
    c2 = bb.B200Classifier(max_batch=2, precision=bb.PRECISION_F32)
    c2.close(); c2.close()                                   # idempotent Close()
    with pytest.raises(bb.B200Error) as e:
        c2.predict(np.zeros(144000, np.float32))
    assert e.value.status == bb.ERR_CLOSED
    with pytest.raises(bb.B200Error) as e:
        bb.B200Classifier(model_data=open(bb.DEFAULT_MODEL, "rb").read()[:1 << 20])
    assert e.value.status == bb.ERR_UNSUPPORTED_MODEL


def test_intermediate_tensors_match_oracle(lib_path, audio):
    """Frontend output and every block output vs the float64 oracle (fp32 truth path)."""
    c = bb.B200Classifier(max_batch=4, micro_batch=4, precision=bb.PRECISION_F32)
    c.keep_intermediates(True)
    chunks = np.stack([audio["tawnyowl"][:144000], audio["soundscape"][:144000]])
    plan = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    ids = [plan["frontend_out_tensor"], plan["stem"]["out_tensor"], plan["mix"]["out_tensor"]] + [b["tensors"]["out"] for b in plan["blocks"]] + [545]
    ref = bo.Oracle(dtype=torch.float64).run(chunks, fetch=tuple(ids), batch=2)
    c.predict_batch(chunks)
    for t in ids:
        r = np.asarray(ref[t], np.float64).reshape(2, -1)
        g = c.read_tensor(t).reshape(2, -1)
        rel = np.abs(g - r).max() / np.abs(r).max()
        # frontend: x^0.19..0.23 compression amplifies fp32 rounding at low mel energies (SURVEY.md §0.4 measured
        # 6.2e-3 abs / range 13.5 for an fp32 contraction); everything downstream inherits that noise floor.
        assert rel < 1.5e-3, (t, rel)
    c.close()


def test_intermediate_tensors_of_the_tensor_core_path(lib_path, audio):
    """The production F16X3 path (fp16 hi/lo planes, mbconv2 + pw2 kernels): depthwise outputs, SE gates and block outputs of
    all 16 blocks + the embedding vs the float64 oracle."""
    c = bb.B200Classifier(max_batch=4, micro_batch=4, precision=bb.PRECISION_F16X3)
    c.keep_intermediates(True)
    chunks = np.stack([audio["tawnyowl"][:144000], audio["soundscape"][:144000]])
    plan = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    ids = [plan["frontend_out_tensor"], plan["mix"]["out_tensor"]]
    for b in plan["blocks"]:
        ids += [b["tensors"]["dw"], b["tensors"]["out"]] + ([b["tensors"]["gate"]] if b["tensors"].get("gate", -1) >= 0 else [])
    ids += [545]
    ref = bo.Oracle(dtype=torch.float64).run(chunks, fetch=tuple(ids), batch=2)
    got_logits = c.predict_batch(chunks)
    worst = 0.0
    for t in ids:
        r = np.asarray(ref[t], np.float64).reshape(2, -1)
        g = c.read_tensor(t).reshape(2, -1)
        rel = np.abs(g - r).max() / np.abs(r).max()
        worst = max(worst, rel)
        assert rel < 1.5e-3, (t, rel)
    print("worst relative error over %d tensors: %.2e" % (len(ids), worst))
    c.keep_intermediates(False)
    assert np.array_equal(c.predict_batch(chunks), got_logits)          # keep mode and production mode: same bits
    c.close()


def test_full_size_property_linearity_of_batching(clf):
    """BASELINE config 3 shape (synthetic pink noise + chirp): duplicated chunks give identical rows, and
    gain-invariance of the min/max normalisation: logits(x) == logits(0.5 x) up to fp32 rounding."""
    from bench import synth_chunks
    x = synth_chunks(24, seed0=1234)
    y = clf.predict_batch(np.concatenate([x, x[:8]]))
    assert np.array_equal(y[:8], y[24:])
    z = clf.predict_batch((x[:8] * np.float32(0.5)))
    assert np.abs(_sig(z) - _sig(y[:8])).max() <= SIG_TOL
    # parity of the synthetic workload itself (broadband noise + a loud tone: the hardest case for the split GEMMs)
    ref = bo.Oracle(dtype=torch.float64).predict_batch(x[:6])
    ds = np.abs(_sig(y[:6]) - _sig(ref)).max()
    print("synthetic chirp+noise: max|dsigmoid|=%.3e max|dlogit|=%.3e" % (ds, np.abs(y[:6] - ref).max()))
    assert ds <= SIG_TOL
    top2 = np.sort(ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-2                      # top-1 must agree wherever the oracle's margin is not a tie
    assert (y[:6].argmax(1) == ref.argmax(1))[clear].all()


def test_bat_pipeline_backbone_embedding(clf):
    """BASELINE config 4 (BattyBirdNET): 144000 raw samples captured at 256 kHz go through the SAME v2.4 backbone and the
    [1024] embedding feeds a small custom head (bat_onnx.go:252, custom_classifier.go:147-173).  The regional head files are
    downloaded at run time in the reference and are not available here, so: backbone-to-embedding parity on synthetic
    20-80 kHz FM sweeps + noise, and a seeded synthetic dense+sigmoid head applied to both embeddings."""
    fs = 256000.0
    t = np.arange(144000) / fs
    rng = np.random.default_rng(256)
    x = np.zeros((6, 144000), np.float32)
    for i in range(6):
        f0, f1 = rng.uniform(60e3, 80e3), rng.uniform(20e3, 35e3)
        dur = rng.uniform(0.004, 0.02)
        sig = 0.02 * rng.standard_normal(144000)
        for start in np.arange(0.02, 0.55, rng.uniform(0.06, 0.12)):
            m = (t >= start) & (t < start + dur)
            tt = t[m] - start
            sig[m] += 0.4 * np.sin(2 * np.pi * (f0 * tt + 0.5 * (f1 - f0) / dur * tt * tt)) * np.hanning(m.sum())
        x[i] = np.round(np.clip(sig, -1, 1) * 32767).astype(np.int16).astype(np.float32) / np.float32(32768)
    logits, emb = clf.predict_batch(x, with_embeddings=True)
    rl, remb = bo.Oracle(dtype=torch.float64).predict_batch(x, with_embeddings=True)
    assert np.abs(emb - remb).max() <= 2e-3 * max(1.0, np.abs(remb).max())
    assert np.abs(_sig(logits) - _sig(rl)).max() <= SIG_TOL
    w = rng.standard_normal((1024, 12)).astype(np.float64) * 0.05     # synthetic 12-class regional head
    head = lambda e: 1.0 / (1.0 + np.exp(-(e.astype(np.float64) @ w)))
    assert np.abs(head(emb) - head(remb)).max() <= 1e-3


def test_full_batch_256_invariants(lib_path):
    """BASELINE config 2/3 sizes (batch 256, micro-batches, two phases): independence of chunks and of batch position."""
    from bench import soundscape_batch
    c = bb.B200Classifier(max_batch=256)
    x = soundscape_batch(256)                                    # 79 distinct chunks tiled
    y = c.predict_batch(x)
    assert np.isfinite(y).all()
    for k in range(79, 256):
        assert np.array_equal(y[k], y[k % 79]), k                # same chunk in another micro-batch / phase slot: identical bits
    idx, conf = c.analyze_batch(x, 1.5, 10)
    assert np.array_equal(idx[:79], idx[79:158]) and (np.diff(conf, axis=1) <= 0).all()
    c.close()


def test_fused_and_unfused_chains_agree(lib_path, golden, audio, monkeypatch):
    """The fused expand+depthwise kernel (default for the large-map blocks, forced for all blocks with BNB_FUSED=2)
    and the two-kernel chain (BNB_FUSED=0) are two schedules of the same arithmetic: same oracle, same tolerance."""
    chunks = bo.slice_chunks(audio["soundscape"], 72000)[:24]
    ref = golden["soundscape_logits"][:24].astype(np.float64)
    outs = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("BNB_FUSED", mode)
        c = bb.B200Classifier(max_batch=24, micro_batch=8)
        outs[mode] = c.predict_batch(chunks).copy()
        c.close()
        err = np.abs(_sig(outs[mode]) - _sig(ref)).max()
        assert err <= 1e-3, (mode, err)
        assert (outs[mode].argmax(1) == ref.argmax(1)).all()
    assert np.abs(outs["0"] - outs["2"]).max() < 5e-3


# ---------------------------------------------------------------------------------------------------------------------
# round 2: the benched geometry, the async / graph entry points, hostile inputs, several devices in one process
@pytest.mark.parametrize("micro", [128, 64])
def test_benched_geometry_matches_golden(lib_path, golden, micro):
    """bench.py's exact classifier (max_batch 256, micro-batch 128 — the library default — and the 64 of earlier runs, 2 lanes):
    all 79 golden soundscape rows, wherever they sit in the tiled batch of 256, against the float64 oracle (VERDICT r1 weak #1:
    the benched configuration had no oracle test)."""
    from bench import soundscape_batch
    c = bb.B200Classifier(max_batch=256, micro_batch=micro, lanes=2)
    x = soundscape_batch(256)
    y = c.predict_batch(x)
    ref = golden["soundscape_logits"].astype(np.float64)
    for k in range(256):
        r = ref[k % 79]
        assert np.abs(_sig(y[k]) - _sig(r)).max() <= SIG_TOL, k
        assert int(y[k].argmax()) == int(r.argmax()), k
    print("benched geometry: max|dsigmoid| = %.3e" % np.abs(_sig(y[:79]) - _sig(ref)).max())
    idx, conf = c.analyze_batch(x, 1.0, 10)
    assert (idx[:, 0] == ref.argmax(1)[np.arange(256) % 79]).all()
    c.close()


def test_config3_shape_max_batch_1024(lib_path):
    """BASELINE config 3 shape: batch 1024 of synthetic pink noise + chirp through one call; a fixed 512-chunk subset is compared
    with the committed oracle vectors (tests/golden/make_synth_golden.py), the rest through batch-position invariance."""
    from bench import synth_chunks
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "synth512_golden.npz"))
    x = synth_chunks(512, seed0=1234)
    c = bb.B200Classifier(max_batch=1024, lanes=2)                            # library-default micro-batch (128)
    big = np.concatenate([x, x[::-1]])
    idx, conf, logits = c.analyze_batch(big, 1.0, 10, want_logits=True)
    assert np.array_equal(logits[:512], logits[1023:511:-1])                 # same chunk, another slot of the batch: same bits
    sig = bo.sigmoid_sensitivity(logits[:512], 1.0).astype(np.float64)
    top_ref, conf_ref = g["top10_idx"], g["top10_conf"].astype(np.float64)
    got_at_ref = np.take_along_axis(sig, top_ref.astype(np.int64), axis=1)
    assert np.abs(got_at_ref - conf_ref).max() <= SIG_TOL                     # the oracle's ten best per chunk
    assert np.abs(sig.max(1) - conf_ref[:, 0]).max() <= SIG_TOL
    clear = (conf_ref[:, 0] - conf_ref[:, 1]) > 2e-3
    assert (idx[:512, 0] == top_ref[:, 0])[clear].all()
    c.close()


def test_async_submit_wait_and_graphs_give_the_same_bits(lib_path, audio):
    chunks = bo.slice_chunks(audio["soundscape"], 72000)
    c = bb.B200Classifier(max_batch=79, micro_batch=16, use_graphs=1)
    idx0, conf0, lg0 = c.analyze_batch(chunks, 1.5, 10, want_logits=True)
    # two batches in flight, different sizes, results must equal the synchronous call bit for bit
    t1 = c.analyze_batch_submit(chunks[:40], 1.5, 10, want_logits=True)
    t2 = c.analyze_batch_submit(chunks[40:], 1.5, 10, want_logits=True)
    with pytest.raises(bb.B200Error):
        c.analyze_batch_submit(chunks[:1], 1.5, 10)                          # a third outstanding ticket is refused
    i1, c1, l1 = t1.wait()
    i2, c2, l2 = t2.wait()
    assert np.array_equal(np.concatenate([l1, l2]), lg0) and np.array_equal(np.concatenate([i1, i2]), idx0)
    assert np.array_equal(np.concatenate([c1, c2]), conf0)
    with pytest.raises(bb.B200Error):
        t1.wait()                                                             # a ticket completes once
    # small batches replay from a CUDA graph (B <= micro-batch): first call captures, later calls replay
    for n in (1, 3, 16, 1, 3):
        i3, c3, l3 = c.analyze_batch(chunks[:n], 1.5, 10, want_logits=True)
        assert np.array_equal(l3, lg0[:n]) and np.array_equal(i3, idx0[:n])
    one = c.predict(chunks[5])
    assert np.array_equal(one, lg0[5])
    n0 = c.kernel_launches()
    c.predict(chunks[5])
    assert c.kernel_launches() > n0                                           # graph replays still count their kernels
    c.close()


def test_nan_and_inf_samples_do_not_poison_the_process(lib_path, audio):
    """A NaN / Inf PCM sample must not fault the top-k kernel (an all-NaN chunk indexed shared memory out of bounds)
    nor touch the other chunks of the batch."""
    chunks = bo.slice_chunks(audio["soundscape"], 144000)[:4].copy()
    good = chunks.copy()
    chunks[1, 1000] = np.nan
    chunks[2, 77] = np.inf
    c = bb.B200Classifier(max_batch=4)
    idx, conf, logits = c.analyze_batch(chunks, 1.0, 10, want_logits=True)
    ref_idx, ref_conf, ref_logits = c.analyze_batch(good, 1.0, 10, want_logits=True)
    assert np.array_equal(logits[[0, 3]], ref_logits[[0, 3]]) and np.array_equal(idx[[0, 3]], ref_idx[[0, 3]])
    # what a NaN sample does to ITS chunk is implementation-defined (the stem's ReLU is max(x, 0), which drops NaN on most
    # hardware, XNNPACK included): all that is required is a well-formed answer — indices in range, confidences NaN or in [0, 1]
    for r in (1, 2):
        assert ((idx[r] >= 0) & (idx[r] < 6522)).all()
        assert (np.isnan(conf[r]) | ((conf[r] >= 0) & (conf[r] <= 1))).all()
    assert c.predict(good[0]).shape == (6522,)                                # the handle is still healthy
    c.close()


def test_two_devices_in_one_process(lib_path, audio):
    """A Go process drives every GPU of the box from one address space (stream_id mod R): handles on different
    devices must not share per-device kernel attributes or scratch state."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    chunks = bo.slice_chunks(audio["soundscape"], 144000)[:6]
    a = bb.B200Classifier(device=0, max_batch=8)
    b = bb.B200Classifier(device=1, max_batch=8)
    ya, yb = a.predict_batch(chunks), b.predict_batch(chunks)
    assert np.array_equal(ya, yb)
    import threading
    out = {}
    ths = [threading.Thread(target=lambda k, c: out.__setitem__(k, c.predict_batch(chunks)), args=(k, c)) for k, c in (("a", a), ("b", b))]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert np.array_equal(out["a"], ya) and np.array_equal(out["b"], ya)
    a.close(); b.close()


def test_detections_on_device_match_host_filter(lib_path, golden):
    """N1: bnb_analyze_batch_detections (sigmoid + threshold + compaction on the device) = thresholding the analyze_batch top-10
    on the host, in chunk-then-confidence order; the published soundscape table (threshold 0.1, sensitivity 1.5: 25 rows of
    /root/reference/doc/wiki/file-analysis.md:20-44) comes back as the top-1 entries of its chunks."""
    from bench import soundscape_batch
    x = soundscape_batch(160)
    c = bb.B200Classifier(max_batch=256)
    idx, conf = c.analyze_batch(x, 1.5, 10)
    for thr in (0.1, 0.5, 0.0001):
        ch, sp, cf, counts = c.analyze_batch_detections(x, 1.5, thr, 10)
        keep = conf >= thr
        assert np.array_equal(counts, keep.sum(1))
        rows = np.nonzero(keep)
        assert np.array_equal(ch, rows[0]) and np.array_equal(sp, idx[keep]) and np.array_equal(cf, conf[keep])
    ch, sp, cf, counts = c.analyze_batch_detections(x, 1.5, 0.1, 10, max_det=7)            # truncated list, full counts
    assert len(ch) == 7 and counts.sum() == (conf >= 0.1).sum()
    x16 = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)                     # int16 ingest through the same entry
    ch16, sp16, cf16, _ = c.analyze_batch_detections(x16, 1.5, 0.1, 10)
    i16, c16 = c.analyze_batch(x16, 1.5, 10)
    assert np.array_equal(sp16, i16[c16 >= 0.1])
    c.close()
