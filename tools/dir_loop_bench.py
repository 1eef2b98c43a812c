"""Directory frame loop (SURVEY §8f rows N1 + N2): JPEG files in -> detect + blur -> JPEG files out, with the codec on the
GPU (io="device") against the same loop with cv2.imread / cv2.imwrite thread pools (io="host"), and against the
reference's own schedule (6 loader + 6 saver threads, load -> detect -> composite -> save one batch after the other,
combine_detect.py:183-277) restated around the same detector.  Output files are compared byte for byte.
usage: python tools/dir_loop_bench.py [n_frames] [H] [W]"""
import json
import os
import shutil
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))
import cv2
import numpy as np
import torch
from vdesens_b200 import synth
from detect_face.face import Retinaface
import combine_detect as cd

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
B = 16
tmp = tempfile.mkdtemp(prefix="vdloop_")
src = os.path.join(tmp, "in")
os.makedirs(src)
uniq = [synth.synthetic_frame(H, W, 7000 + i) for i in range(16)]
for i in range(N):
    cv2.imwrite(os.path.join(src, f"frame_{i:06d}.jpg"), cv2.cvtColor(uniq[i % 16], cv2.COLOR_RGB2BGR))
size_mb = sum(os.path.getsize(os.path.join(src, f)) for f in os.listdir(src)) / 1e6
det = Retinaface(state_dict=synth.synthetic_state_dict(0, face_rate="r001"), input_shape=[H, W, 3], confidence=0.5, nms_iou=0.4,
                 letterbox_image=True, cuda=True, batch_size=B, precision=os.environ.get("VD_PRECISION", "fp16"))


def reference_schedule(input_dir, output_dir):
    """the reference's loop shape (combine_detect.py:183-277): 6 loader threads per batch, detect, composite, 6 saver threads,
    one batch after the other — with THIS repo's detector and compositing in the middle (the reference's own CPU/GPU
    detector would only make it slower)"""
    paths = sorted(os.path.join(input_dir, f) for f in os.listdir(input_dir))
    os.makedirs(output_dir, exist_ok=True)
    saver, futs = ThreadPoolExecutor(max_workers=6), []
    for i in range(0, len(paths), B):
        files = paths[i:i + B]
        with ThreadPoolExecutor(max_workers=6) as loader:
            frames = list(loader.map(cd.load_image_rgb, files))
        out, _, _ = cd.batch_process_frames(frames, det, None, batch_size=B, mode="gauss", ksize=31)
        for p, img in zip(files, out):
            futs.append(saver.submit(cd.save_output_image, img, os.path.join(output_dir, "processed_" + os.path.basename(p))))
    for f in futs:
        f.result()
    saver.shutdown()
    return len(paths)


res = {"frames": N, "frame": [H, W], "input_MB": size_mb, "host_cpus": os.cpu_count(), "io_workers": cd._io_workers()}
# warm-up (engine build, cuDNN-free: nothing to autotune, but the first batch pays allocations)
cd.batch_process_images(src, os.path.join(tmp, "warm"), det, None, batch_size=B, mode="gauss", ksize=31, io="device")
for name, fn in (("device", lambda d: cd.batch_process_images(src, d, det, None, batch_size=B, mode="gauss", ksize=31, io="device")[0]),
                 ("host", lambda d: cd.batch_process_images(src, d, det, None, batch_size=B, mode="gauss", ksize=31, io="host")[0]),
                 ("reference_schedule", lambda d: reference_schedule(src, d))):
    dst = os.path.join(tmp, name)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = fn(dst)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[name] = {"frames_per_s": n / dt, "seconds": dt}
same = all(open(os.path.join(tmp, "device", f), "rb").read() == open(os.path.join(tmp, "host", f), "rb").read()
           for f in sorted(os.listdir(os.path.join(tmp, "host"))))
res["device_files_identical_to_host_files"] = bool(same)
res["device_over_host"] = res["device"]["frames_per_s"] / res["host"]["frames_per_s"]
res["device_over_reference_schedule"] = res["device"]["frames_per_s"] / res["reference_schedule"]["frames_per_s"]


# ---- video stage: the same frames as a Motion-JPEG AVI ------------------------------------------
def video_stage():
    from vdesens_b200 import avi
    video = os.path.join(tmp, "in.avi")
    with avi.AviWriter(video, W, H, 60) as wr:
        for f in sorted(os.listdir(src)):
            wr.write(open(os.path.join(src, f), "rb").read())


    def reference_video_chain(path, out_path):
        """reference :597-646 restated around this repo's detector: video -> JPEG files (here a plain demux; the reference
        transcodes through an ffmpeg subprocess) -> its loop schedule over the files -> JPEG files -> cv2.VideoWriter
        (its `.avi` codec list, XVID first) re-encoding every file"""
        d_in, d_out = os.path.join(tmp, "v_frames"), os.path.join(tmp, "v_processed")
        n = cd.convert_video_to_frames(path, d_in)
        reference_schedule(d_in, d_out)
        files = sorted(os.listdir(d_out))
        wr, _ = cd._open_cv2_writer(out_path, 60, W, H)
        for f in files:
            wr.write(cv2.imread(os.path.join(d_out, f)))
        wr.release()
        return n


    cd.process_video_pipeline(video, os.path.join(tmp, "warm.avi"), det, None, batch_size=B, mode="gauss", ksize=31)
    vres = {}
    for name, fn in (("resident", lambda o: cd.process_video_pipeline(video, o, det, None, batch_size=B, mode="gauss", ksize=31)[1]),
                     ("reference_chain", lambda o: reference_video_chain(video, o))):
        out = os.path.join(tmp, name + ".avi")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = fn(out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        vres[name] = {"frames_per_s": n / dt, "seconds": dt, "output_MB": os.path.getsize(out) / 1e6}
    with avi.AviReader(os.path.join(tmp, "resident.avi")) as rd:
        names = sorted(os.listdir(os.path.join(tmp, "device")))
        vres["frames_identical_to_directory_loop_files"] = bool(
            len(rd) == len(names) and all(rd.frame(i) == open(os.path.join(tmp, "device", f), "rb").read() for i, f in enumerate(names)))
    vres["resident_over_reference_chain"] = vres["resident"]["frames_per_s"] / vres["reference_chain"]["frames_per_s"]
    res["video"] = vres


try:
    video_stage()
except Exception as e:                        # noqa: BLE001: keep the directory numbers
    res["video"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
print(json.dumps(res))
shutil.rmtree(tmp, ignore_errors=True)
