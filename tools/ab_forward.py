"""Interleaved A/B of the whole network forward for one environment switch that DetectorEngine (or the plan
creation behind it) reads at construction:   python tools/ab_forward.py VAR v0 v1 [B] [rounds]
Two engines on the same frames, timed alternately (CUDA events on the engine stream, 10 forwards per sample), so box /
clock drift hits both arms alike.  Prints the per-arm median and the ratio; checks both arms return identical rows."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import numpy as np, torch
from vdesens_b200 import synth
from vdesens_b200.engine import DetectorEngine

var, v0, v1 = sys.argv[1], sys.argv[2], sys.argv[3]
B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 7
H, W = 1080, 1920
sd = synth.synthetic_state_dict(0, face_rate="r001")
frames = torch.from_numpy(np.stack([synth.synthetic_frame(H, W, 1000 + i) for i in range(B)])).cuda()
engs = []
for v in (v0, v1):
    os.environ[var] = v
    e = DetectorEngine(sd, B, H, W)
    e.set_frames(frames)
    engs.append(e)
times = [[], []]
res = []
for e in engs:
    for _ in range(3):
        e.forward()
    res.append(tuple(a.copy() for a in e.fetch_results()))
for r in range(rounds):
    for i, e in enumerate(engs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(e.stream):
            e0.record(e.stream)
            for _ in range(10):
                e.forward()
            e1.record(e.stream)
        e.stream.synchronize()
        times[i].append(e0.elapsed_time(e1) / 10)
m0, m1 = statistics.median(times[0]), statistics.median(times[1])
same = all(np.array_equal(a, b) for a, b in zip(res[0], res[1]))
print(f"{var}={v0}: median {m0:.4f} ms (min {min(times[0]):.4f})   {var}={v1}: median {m1:.4f} ms (min {min(times[1]):.4f})   "
      f"ratio {m1 / m0:.4f}   identical results: {same}")
