"""Per-op device times of the MobileNet0.25 variant (CUDA events, each op repeated) with the algorithmic HBM
bytes of each op (input read once + output written once) and the bandwidth that corresponds to.
usage: python tools/layer_times_mnet.py [B] [H] [W] [reps] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import torch
from vdesens_b200 import synth, arch_mnet
from vdesens_b200.engine import DetectorEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
precision = sys.argv[5] if len(sys.argv) > 5 else "bf16"

sd = synth.synthetic_state_dict(0, face_rate="r01", backbone="mobilenet")
eng = DetectorEngine(sd, B, H, W, backbone="mobilenet", precision=precision)
eng.use_graph = False                # per-op timing: eager launches (a graph would be captured inside the timed loops)
frames = torch.stack([torch.from_numpy(synth.synthetic_frame(H, W, 1000 + i)) for i in range(B)]).cuda()
eng.set_frames(frames)
eng.forward()
eng.stream.synchronize()
ops, _ = arch_mnet.build_graph(fuse_dwpw=eng.fuse_dwpw)
byname = {op.name: op for op in ops}


def nbytes(name):
    t = eng.buf[name]
    if name == "x0":
        return B * H * W * 8          # 4 channels x 2 bytes per pixel of the ingest layout
    return t.numel() * t.element_size()


tot = 0.0
rows = []
for i, (kind, h, name) in enumerate(eng.plans):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    saved = eng.plans
    eng.plans = [saved[i]]
    with torch.cuda.stream(eng.stream):
        e0.record(eng.stream)
        for _ in range(reps):
            eng.enqueue_network()
        e1.record(eng.stream)
    eng.stream.synchronize()
    eng.plans = saved
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    op = byname[name]
    by = nbytes(op.src)
    if op.kind == "heads":
        by += nbytes(op.src2) + eng.heads[op.level].numel() * 4
    else:
        by += nbytes(op.dst)
        if op.upsample_add:
            by += nbytes(op.upsample_add)
    rows.append((kind, name, ms, by))
print(f"mobilenet0.25 {precision} B={B} {H}x{W}: sum of per-op times {tot:.3f} ms ({B / tot * 1e3:.0f} frames/s network-only), "
      f"algorithmic bytes {sum(r[3] for r in rows) / 1e6:.1f} MB/batch")
for kind, name, ms, by in rows:
    print(f"{kind:5s} {name:24s} {ms:8.4f} ms  {by / 1e6:9.2f} MB  {by / ms / 1e6:8.1f} GB/s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.stream):
    e0.record(eng.stream)
    for _ in range(5):
        eng.forward()
    e1.record(eng.stream)
eng.stream.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"forward (ingest+net+decode+nms): {ms:.3f} ms/batch -> {B / ms * 1e3:.1f} frames/s")
_, counts = eng.fetch_results()
print("kept per frame:", counts.tolist())
