"""Runs one conv layer of the 1080p engine a few times (for ncu).  usage: one_layer.py <layer name> [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import ctypes as C
import torch
from vdesens_b200 import synth, _lib
from vdesens_b200.engine import DetectorEngine
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = DetectorEngine(synth.synthetic_state_dict(0), B, 1080, 1920)
for k, t in eng.buf.items():
    t.zero_()
h = next(h for kind, h, n in eng.plans if n == name)
s = C.c_void_p(eng.stream.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.check(eng.lib.vd_conv_run(h, s)); eng.stream.synchronize()
e0.record(eng.stream)
for _ in range(reps):
    _lib.check(eng.lib.vd_conv_run(h, s))
e1.record(eng.stream); eng.stream.synchronize()
print(name, "ms", e0.elapsed_time(e1) / reps, [p for p in eng.plan_info() if p[0] == name])
