"""Device JPEG codec timings at the frame loop's size: decode (file bytes -> RGB frames) and encode (RGB frames -> file
bytes) of one batch, GPU time from CUDA events and wall time of the public calls.  usage: python tools/jpeg_times.py [B] [H] [W]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import cv2, numpy as np, torch
from vdesens_b200 import synth, ops
from vdesens_b200.jpeg import JpegDecoder, JpegEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
imgs = [synth.synthetic_frame(H, W, 8000 + i) for i in range(B)]
files = [cv2.imencode(".jpg", cv2.cvtColor(im, cv2.COLOR_RGB2BGR))[1].tobytes() for im in imgs]
res = {"B": B, "frame": [H, W], "file_MB": sum(len(f) for f in files) / B / 1e6}
for sub in (256, 512, 1024, 2048):
    dec = JpegDecoder(B, sub_bytes=sub)
    dec.decode(files); dec.decode(files)
    t0 = time.perf_counter()
    for _ in range(5):
        frames, cpu = dec.decode(files)
    torch.cuda.synchronize()
    res[f"decode_wall_ms_sub{sub}"] = (time.perf_counter() - t0) / 5 * 1e3
    res[f"on_cpu_sub{sub}"] = cpu
    res[f"sync_rounds_sub{sub}"] = dec.sync_rounds()
    # GPU time of the kernels alone: re-enqueue on the already staged bytes
    import ctypes as C
    from vdesens_b200 import _lib
    g0 = __import__("vdesens_b200.jpeg", fromlist=["x"]).prepare_device(files[0])
    desc = ops.frame_descriptors([frames[i] for i in range(B)])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(dec.stream):
        e0.record(dec.stream)
        for _ in range(5):
            _lib.check(dec.lib.vd_jpeg_decode_batch(C.c_void_p(dec.stage_dev.data_ptr()), C.byref(g0), C.c_void_p(dec.stage_dev.data_ptr()), B,
                                                    dec.max_file, C.c_void_p(desc.data_ptr()), C.c_void_p(dec.result.data_ptr()), sub,
                                                    C.c_void_p(dec._ws.data_ptr()), dec._ws_bytes, C.c_void_p(dec.stream.cuda_stream)), "dec", dec.lib)
        e1.record(dec.stream)
    dec.stream.synchronize()
    res[f"decode_gpu_ms_sub{sub}"] = e0.elapsed_time(e1) / 5
assert np.array_equal(frames[0].cpu().numpy(), cv2.cvtColor(cv2.imdecode(np.frombuffer(files[0], np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB))
enc = JpegEncoder(B, H, W)
out = enc.encode(frames); out = enc.encode(frames)
t0 = time.perf_counter()
for _ in range(5):
    out = enc.encode(frames)
res["encode_wall_ms"] = (time.perf_counter() - t0) / 5 * 1e3
desc = ops.frame_descriptors([frames[i] for i in range(B)])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(enc.stream)
for _ in range(5):
    enc.enqueue(desc)
e1.record(enc.stream)
enc.stream.synchronize()
res["encode_gpu_ms"] = e0.elapsed_time(e1) / 5
t0 = time.perf_counter()
for f in files[:4]:
    cv2.cvtColor(cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
res["cv2_decode_ms_per_frame_1core"] = (time.perf_counter() - t0) / 4 * 1e3
t0 = time.perf_counter()
for im in imgs[:4]:
    cv2.imencode(".jpg", cv2.cvtColor(im, cv2.COLOR_RGB2BGR))
res["cv2_encode_ms_per_frame_1core"] = (time.perf_counter() - t0) / 4 * 1e3
res["decode_gpu_ms_per_frame"] = res["decode_gpu_ms_sub1024"] / B
res["encode_gpu_ms_per_frame"] = res["encode_gpu_ms"] / B
print(json.dumps(res))
