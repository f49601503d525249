"""The C++ host mirror (birdnet-go_b200/host/birdnet_host.hpp): the reference's Go-side logic for this path restated
in C++ above the C ABI (no Go toolchain in the image).  CPU part = the reference's own unit-test cases for
AnalysisBuffer / sigmoid / top-k / PCM conversion; GPU part = the drop-in path end to end against the published table."""
import os
import subprocess

import pytest

import birdnet_b200 as bb
import birdnet_oracle as bo

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(REPO, "birdnet-go_b200", "host")


@pytest.fixture(scope="module")
def host_test(lib_path):
    exe = os.path.join(HOST, "host_test")
    cmd = ["g++", "-std=c++17", "-O2", os.path.join(HOST, "host_test.cc"), "-I" + os.path.join(REPO, "include"),
           "-L" + os.path.dirname(lib_path), "-lbirdnet_b200", "-Wl,-rpath," + os.path.dirname(lib_path), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_host_logic_unit_cases(host_test):
    r = subprocess.run([host_test], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr


def test_host_fails_closed_without_gpu(host_test):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([host_test, "analyze", bo.MODEL_PATH, bo.LABELS_PATH, os.path.join(bo.ASSETS, "soundscape.wav")], capture_output=True, text=True)
    assert r.returncode == 3 and "UNAVAILABLE" in r.stdout      # ErrB200Unavailable -> caller falls back to TFLite


@pytest.mark.gpu
def test_dropin_predict_reproduces_published_table(host_test):
    """BirdNET::Predict (batch-1, sample[0], sigmoid(1.5 x), top-10) through bnb_predict == doc/wiki/file-analysis.md:20-44."""
    r = subprocess.run([host_test, "analyze", bo.MODEL_PATH, bo.LABELS_PATH, os.path.join(bo.ASSETS, "soundscape.wav"), "1.5", "0.1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [ln.split("\t") for ln in r.stdout.splitlines() if "\t" in ln]
    # two listings: "batched" rows = AnalyzeFileBatched (one int16 call, threshold + compaction on the device, best row per window),
    # plain rows = the drop-in loop (BirdNET::Predict per chunk); both must be the published table
    for kind in ("batched", "dropin"):
        sel = [r3[1:] for r3 in rows if r3[0] == "batched"] if kind == "batched" else [r3 for r3 in rows if r3[0] != "batched"]
        got = [(float(t), sp.split("_", 1)[1], float(c)) for t, sp, c in sel]
        assert len(got) == len(bo.GOLDEN_TABLE), kind
        for (t, name, c), (gt, gname, gc) in zip(got, bo.GOLDEN_TABLE):
            assert t == gt and name == gname and abs(c - gc) <= 1e-3, kind
    assert "size-mismatch-error ok" in r.stdout


@pytest.mark.gpu
def test_batched_offline_file_analysis_reproduces_the_published_table(lib_path):
    """BASELINE config 2 driver (doc/wiki/file-analysis.md:1-13): soundscape.wav, no overlap, threshold 0.1, sensitivity 1.5 through
    ONE batched int16 call must list exactly the reference's 25 detections (species and confidence to the table's 4 decimals)."""
    import bench
    pcm = bench.read_wav_int16(os.path.join(bo.ASSETS, "soundscape.wav"))
    bn = bb.BirdNET(sensitivity=bo.GOLDEN_SENSITIVITY, max_batch=64)
    rows = bn.analyze_file(pcm, overlap_s=0.0, threshold=0.1)
    top = {}
    for b0, e0, sp, cf in rows:
        top.setdefault(b0, (sp, cf))                       # first row of a window = its best result
    assert len(top) == len(bo.GOLDEN_TABLE)
    for t0, name, c in bo.GOLDEN_TABLE:
        sp, cf = top[float(t0)]
        assert sp.split("_", 1)[1] == name and abs(cf - c) <= 1e-3, (t0, sp, cf)
    # overlap 1.5 s: 79 windows, every window's rows are sorted and above the threshold
    rows = bn.analyze_file(pcm, overlap_s=1.5, threshold=0.1)
    assert all(r[3] >= 0.1 for r in rows) and len({r[0] for r in rows}) <= 79
    bn.close()
