"""Times the generic (letterbox) ingest kernel: 16 x 4K -> 1080p and 16 x 1080p -> 1280x1280 (the reference's default input_shape).
usage: python tools/letterbox_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import torch, bench
dev = torch.device("cuda:0")
print(bench.letterbox_ingest_extra(dev, "fp16"))
import ctypes as C
import numpy as np
from vdesens_b200 import _lib, ops, synth
lib = _lib.lib("fp16")
for (ih, iw, H, W) in ((1080, 1920, 1280, 1280), (720, 1280, 1080, 1920)):
    B = 16
    frames = torch.from_numpy(np.stack([synth.synthetic_frame(ih, iw, 9100 + i) for i in range(2)])).to(dev)
    desc = ops.frame_descriptors([frames[i % 2] for i in range(B)])
    x0 = torch.zeros((B, H, ops.x0_pitch(W), 4), dtype=torch.float16, device=dev)
    mean = (C.c_float * 3)(104.0, 117.0, 123.0)
    st = torch.cuda.current_stream(dev)
    s = C.c_void_p(st.cuda_stream)
    for _ in range(3):
        _lib.check(lib.vd_ingest_u8hwc(C.c_void_p(desc.data_ptr()), B, C.c_void_p(x0.data_ptr()), H, W, mean, s), "vd_ingest_u8hwc", lib)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        _lib.check(lib.vd_ingest_u8hwc(C.c_void_p(desc.data_ptr()), B, C.c_void_p(x0.data_ptr()), H, W, mean, s), "vd_ingest_u8hwc", lib)
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nb = B * (ih * iw * 3 + H * W * 8)
    print(f"{ih}x{iw} -> {H}x{W}: {ms:.4f} ms, source read once + x0 written = {nb/1e6:.0f} MB -> {nb/ms/1e6:.0f} GB/s")
