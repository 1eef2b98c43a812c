"""Per-layer device times of the detector on the current GPU (CUDA events, each op repeated).
usage: python tools/layer_times.py [B] [H] [W] [reps]"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import ctypes as C
import torch
from vdesens_b200 import synth, arch, _lib
from vdesens_b200.engine import DetectorEngine, _ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5

sd = synth.synthetic_state_dict(0, face_rate="r01")
eng = DetectorEngine(sd, B, H, W, precision=os.environ.get("VD_PRECISION", "bf16"))
frames = torch.stack([torch.from_numpy(synth.synthetic_frame(H, W, 1000 + i)) for i in range(B)]).cuda()
eng.set_frames(frames)
eng.forward()
eng.stream.synchronize()
convs, _, _ = arch.build_graph()
flops = {c.name: 2.0 * c.cout * (c.cin * c.k * c.k + c.cin2) for c in convs}
info = {p[0]: p[1:] for p in eng.plan_info()}
s = C.c_void_p(eng.stream.cuda_stream)
rows = []
tot = 0.0
with torch.cuda.stream(eng.stream):
    for kind, h, name in eng.plans:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for _ in range(reps):
            if kind == "conv":
                _lib.check(eng.lib.vd_conv_run(h, s), "vd_conv_run", eng.lib)
            else:
                src, dst = h
                hh, ww, cc = eng.buf_hw[src]
                _lib.check(eng.lib.vd_maxpool3x3s2(_ptr(eng.buf[src]), _ptr(eng.buf[dst]), B, hh, ww, cc, s))
        e1.record(eng.stream)
        eng.stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot += ms
        if kind == "conv" and name in flops:
            c = next(c for c in convs if c.name == name)
            ho, wo = arch.feature_hw(H, W, 1)
            # output pixels from the plan's destination buffer
            for op_name, hw in eng.buf_hw.items():
                pass
        rows.append((name, ms))
# flops per layer from output extents
out_hw = {}
_, ops_, _ = arch.build_graph(fuse_stem_pool=eng.fuse_stem_pool, fuse_chain=eng.fuse_chain, fuse_siblings=eng.fuse_siblings)
chained = {}
for op in ops_:
    if op.kind == "conv":
        hh, ww, _ = eng.buf_hw[op.dst]
        out_hw[op.conv.name] = hh * ww
        if op.chain is not None:          # the kernel also runs the next 1x1 convolution
            flops[op.conv.name] += flops[op.chain.name]
            chained[op.conv.name] = op.chain.name
        if op.sibling is not None:        # one GEMM over the concatenated filters of two sibling layers
            flops[op.conv.name] += flops[op.sibling.name]
            chained[op.conv.name] = op.sibling.name
print(f"B={B} {H}x{W}  sum of per-layer times {tot:.3f} ms  ({B / tot * 1e3:.1f} frames/s network-only)")
for name, ms in rows:
    if name in flops and name in out_hw:
        gf = flops[name] * out_hw[name] * B / 1e9
        tw, th, nb, bn, nt, grid = info[name]
        extra = f"  (+ {chained[name]})" if name in chained else ""
        print(f"{name:28s} {ms:8.4f} ms  {gf / ms:9.1f} TFLOP/s-eq  tile {tw}x{th}x{nb} BN{bn} tiles {nt} grid {grid}{extra}")
    else:
        print(f"{name:28s} {ms:8.4f} ms")
# whole forward
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(eng.stream):
    e0.record(eng.stream)
    for _ in range(3):
        eng.forward()
    e1.record(eng.stream)
eng.stream.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"forward (ingest+net+decode+nms): {ms:.3f} ms/batch -> {B / ms * 1e3:.1f} frames/s")
rows_, counts = eng.fetch_results()
print("kept per frame:", counts.tolist())
