"""Per-tensor parity report: CUDA path vs float64 oracle for every intermediate the engine materialises.

    python tools/layer_report.py [--precision f32|f16x3] [--out gpurun_out/layer_report.txt]

Debug aid (uses oracle/ as the checker).  Runs a handful of chunks with keep_intermediates on."""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "birdnet-go_b200"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import birdnet_b200 as bb  # noqa: E402
import birdnet_oracle as bo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="f32")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "layer_report.txt"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    prec = {"f32": bb.PRECISION_F32, "f16x3": bb.PRECISION_F16X3}[a.precision]
    x, _ = bo.read_wav(os.path.join(bo.ASSETS, "tawnyowl.wav"))
    y, _ = bo.read_wav(os.path.join(bo.ASSETS, "soundscape.wav"))
    rng = np.random.default_rng(0)
    chunks = np.stack([x[:144000], y[:144000], np.zeros(144000, np.float32),
                       (0.1 * rng.standard_normal(144000)).astype(np.float32)])
    plan = json.loads(bb.describe_model(open(bb.DEFAULT_MODEL, "rb").read()))
    ids = [plan["frontend_out_tensor"], plan["stem"]["out_tensor"], plan["mix"]["out_tensor"]]
    for b in plan["blocks"]:
        t = b["tensors"]
        ids += [t["exp"], t["dw"]] + ([t["gate"]] if t["gate"] >= 0 else []) + [t["out"]]
    ids += [544, plan["post"]["emb_tensor"], plan["head"]["out_tensor"]]
    ref = bo.Oracle(dtype=torch.float64).run(chunks, fetch=tuple(ids), batch=4)
    clf = bb.B200Classifier(max_batch=8, micro_batch=8, precision=prec)
    logits = None
    clf.keep_intermediates(True)
    logits = clf.predict_batch(chunks)
    lines = ["precision=%s device=%s" % (a.precision, clf.runtime_info())]
    worst = 0.0
    for t in ids:
        r = np.asarray(ref[t], np.float64).reshape(len(chunks), -1)
        try:
            g = clf.read_tensor(t).reshape(len(chunks), -1).astype(np.float64)
        except bb.B200Error as e:
            lines.append("tensor %4d: not materialised (%s)" % (t, e)); continue
        if g.shape != r.shape:
            lines.append("tensor %4d: SHAPE MISMATCH got %s want %s" % (t, g.shape, r.shape)); continue
        err = np.abs(g - r)
        scale = np.abs(r).max() + 1e-30
        per_chunk = err.max(1)
        lines.append("tensor %4d: n=%8d absmax_ref=%10.4f max_err=%.3e rel=%.3e per-chunk=%s nan=%d" %
                     (t, r.shape[1], scale, err.max(), err.max() / scale, np.array2string(per_chunk, precision=2), int(np.isnan(g).sum())))
        worst = max(worst, err.max() / scale)
    sg, sr = bo.sigmoid_sensitivity(logits), bo.sigmoid_sensitivity(ref[bo.T_LOGITS])
    lines.append("logits: max|dlogit|=%.3e max|dsigmoid|=%.3e top1 equal=%s" %
                 (np.abs(logits - ref[bo.T_LOGITS]).max(), np.abs(sg.astype(np.float64) - sr).max(),
                  (logits.argmax(1) == ref[bo.T_LOGITS].argmax(1)).tolist()))
    lines.append("launches=%d worst_rel=%.3e" % (clf.kernel_launches(), worst))
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
