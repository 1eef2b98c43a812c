"""Does running two half-batch engines on two streams overlap memory-bound and tensor-bound layers?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import numpy as np, torch
from vdesens_b200 import synth
from vdesens_b200.engine import DetectorEngine
H, W = 1080, 1920
sd = synth.synthetic_state_dict(0, face_rate="r001")
frames = torch.from_numpy(np.stack([synth.synthetic_frame(H, W, 1000 + i) for i in range(16)])).cuda()
def timeit(fn, sync, n=10):
    for _ in range(3): fn()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    sync(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
e16 = DetectorEngine(sd, 16, H, W); e16.set_frames(frames)
t16 = timeit(lambda: e16.forward(), lambda: torch.cuda.current_stream().wait_stream(e16.stream))
print(f"one engine  B=16: {t16:.3f} ms/batch -> {16 / t16 * 1e3:.0f} frames/s")
del e16; torch.cuda.empty_cache()
ea = DetectorEngine(sd, 8, H, W); eb = DetectorEngine(sd, 8, H, W)
ea.set_frames(frames[:8]); eb.set_frames(frames[8:])
def both():
    ea.forward(); eb.forward()
def sync2():
    torch.cuda.current_stream().wait_stream(ea.stream); torch.cuda.current_stream().wait_stream(eb.stream)
t2 = timeit(both, sync2)
print(f"two engines B=8+8 on two streams: {t2:.3f} ms per 16 frames -> {16 / t2 * 1e3:.0f} frames/s")
ta = timeit(lambda: ea.forward(), lambda: torch.cuda.current_stream().wait_stream(ea.stream))
print(f"one engine  B=8: {ta:.3f} ms/batch -> {8 / ta * 1e3:.0f} frames/s")
