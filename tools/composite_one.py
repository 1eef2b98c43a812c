"""One compositing case for ncu (argv[1]: all|small|big).  Where does the compositing time go?  Times vd_composite_rois on the bench workload's boxes, split by box size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import numpy as np, torch
from vdesens_b200 import synth, ops
from vdesens_b200.engine import DetectorEngine
H, W, B = 1080, 1920, 16
sd = synth.synthetic_state_dict(0, face_rate="r001")
eng = DetectorEngine(sd, B, H, W)
frames_np = [synth.synthetic_frame(H, W, 1000 + i) for i in range(B)]
frames = torch.from_numpy(np.stack(frames_np)).cuda()
eng.set_frames(frames); eng.forward()
rows, counts = eng.fetch_results()
boxes = []
for i in range(B):
    r = rows[i, :counts[i], :4].copy() * np.array([W, H, W, H], np.float32)
    bl = [tuple(int(v) for v in b) for b in r.tolist()] + synth.synthetic_boxes(4, H, W, seed=50 + i)
    boxes.append(bl)
def area(b):
    x1, y1, x2, y2 = max(0, b[0]), max(0, b[1]), min(W, b[2]), min(H, b[3])
    return max(0, x2 - x1) * max(0, y2 - y1)
areas = np.array([area(b) for bl in boxes for b in bl])
print("boxes/frame", np.mean([len(b) for b in boxes]), "total area/frame Mpx", areas.sum() / B / 1e6)
print("area quantiles", np.quantile(areas, [0.1, 0.5, 0.9, 0.99]).astype(int), "n>6144:", int((areas > 6144).sum()) / B, "area share of big", areas[areas > 6144].sum() / areas.sum())
eng2 = DetectorEngine(sd, B, H, W, max_ext=512)
eng2.set_frames(frames)
def run(sel, label, mode="gauss", ksize=31, nframes=B):
    bl = [[b for b in boxes[i] if sel(area(b))] if i < nframes else [] for i in range(B)]
    eng2.set_ext_boxes(bl)
    for _ in range(2):
        eng2.enqueue_composite(mode, ksize=ksize, use_faces=False)
    eng2.stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(eng2.stream)
    for _ in range(5):
        eng2.enqueue_composite(mode, ksize=ksize, use_faces=False)
    e1.record(eng2.stream); eng2.stream.synchronize()
    eng2.fetch_results()
    a = sum(area(b) for l in bl for b in l)
    print(f"{label:34s} frames {nframes:2d} boxes {sum(len(l) for l in bl):5d} area {a/1e6:7.2f} Mpx  {e0.elapsed_time(e1)/5:7.3f} ms")

case = sys.argv[1] if len(sys.argv) > 1 else "all"
sel = {"all": lambda a: True, "small": lambda a: a <= 6144, "big": lambda a: a > 6144}[case]
run(sel, case)
