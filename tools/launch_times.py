"""Per-launch device times of one step (events around every launch, real clocks, warm caches).
    python tools/launch_times.py [--batch 256] [--micro-batch 64] [--lanes 1] [--precision f16x3]"""
import argparse, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "birdnet-go_b200")); sys.path.insert(0, REPO)
import torch
import birdnet_b200 as bb
from bench import soundscape_batch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256); ap.add_argument("--micro-batch", type=int, default=64)
ap.add_argument("--lanes", type=int, default=1); ap.add_argument("--precision", default="f16x3")
a = ap.parse_args()
prec = {"f32": bb.PRECISION_F32, "f16x3": bb.PRECISION_F16X3}[a.precision]
clf = bb.B200Classifier(max_batch=a.batch, micro_batch=a.micro_batch, precision=prec, lanes=a.lanes)
x = torch.from_numpy(soundscape_batch(a.batch)).cuda()
lg = torch.empty((a.batch, 6522), device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(3):
    clf.predict_batch_device(x.data_ptr(), bb.PCM_F32, a.batch, lg.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
clf.profile_begin()
clf.predict_batch_device(x.data_ptr(), bb.PCM_F32, a.batch, lg.data_ptr(), 0, st.cuda_stream)
tot = clf.profile_end()
rows = clf.profile_launches()
print("batch=%d micro=%d lanes=%d precision=%s  sum=%.3f ms" % (a.batch, a.micro_batch, a.lanes, a.precision, sum(m for _, m in rows)))
n_front = len(rows)
for i, (c, m) in enumerate(rows):
    print("%4d %-12s %8.1f us" % (i, c, m * 1e3))
