"""Runs one op of the MobileNet0.25 engine a few times (for ncu).  usage: one_mnet_op.py <op name> [B] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import torch
from vdesens_b200 import synth
from vdesens_b200.engine import DetectorEngine
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = DetectorEngine(synth.synthetic_state_dict(0, face_rate="r01", backbone="mobilenet"), B, 1080, 1920, backbone="mobilenet")
eng.use_graph = False                # per-op timing: eager launches (a graph would be captured inside the timed loops)
frames = torch.stack([torch.from_numpy(synth.synthetic_frame(1080, 1920, 1000 + i)) for i in range(B)]).cuda()
eng.set_frames(frames)
eng.forward()                       # real activations in every buffer
eng.stream.synchronize()
eng.plans = [p for p in eng.plans if p[2] == name]
assert len(eng.plans) == 1, name
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.stream):
    eng.enqueue_network()
    e0.record(eng.stream)
    for _ in range(reps):
        eng.enqueue_network()
    e1.record(eng.stream)
eng.stream.synchronize()
print(name, "ms", e0.elapsed_time(e1) / reps)
