"""Drop-in for the hot-path surface of the reference's combine_detect.py, backed by libvdesens:

  mosaic_rectangle_region_single(img, x1, y1, x2, y2, mosaic_level=8)   reference :138-161
  batch_process_images(input_dir, output_dir, face_detector, plate_detector, batch_size=16)  :183-277
plus what BASELINE.json's north_star adds:
  gaussian_rectangle_region_single(...)   the cv2.GaussianBlur ROI idiom, integer-exact
  composite_frames(...)                   batched sequential compositing on the device
  batch_process_frames(...)               the frame loop on in-memory frames (no JPEG hand-off)

Everything outside the per-frame arithmetic (record unpacking, ffmpeg extraction, muxing, config,
logging: reference :21-89, :279-990) is out of scope.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from vdesens_b200 import ops as _ops
from vdesens_b200.pipeline import Anonymizer

_DEFAULT_DEVICE = None


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("combine_detect (B200 build) needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if _DEFAULT_DEVICE is None else _DEFAULT_DEVICE


def composite_frames(frames, boxes_per_frame, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0):
    """frames: list of uint8 HxWx3 arrays (sizes may differ); boxes_per_frame: list of lists of
    (x1,y1,x2,y2) — applied one after another, int()-truncated, clipped, empty ones skipped
    (reference :242-249, :145-150).  Returns NEW arrays; inputs are not modified."""
    dev = _device()
    dfr = [torch.from_numpy(np.ascontiguousarray(f)).to(dev) for f in frames]
    ib = [[tuple(int(v) for v in b) for b in bl] for bl in boxes_per_frame]
    if any(len(b) for b in ib):
        _ops.composite(dfr, ib, mode=mode, mosaic_level=mosaic_level, ksize=ksize, sigma=sigma)
    return [t.cpu().numpy() for t in dfr]


def mosaic_rectangle_region_single(img, x1, y1, x2, y2, mosaic_level=8):
    """reference combine_detect.py:138-161: returns a new array, input untouched."""
    return composite_frames([img], [[(x1, y1, x2, y2)]], "mosaic", mosaic_level)[0]


def gaussian_rectangle_region_single(img, x1, y1, x2, y2, ksize=31, sigma=0.0):
    """img[y1:y2, x1:x2] = cv2.GaussianBlur(img[y1:y2, x1:x2], (ksize, ksize), sigma) on a copy."""
    return composite_frames([img], [[(x1, y1, x2, y2)]], "gauss", ksize=ksize, sigma=sigma)[0]


def _plate_boxes(plate_results, j):
    # reference :239 — only tuple results are honoured: (anything, [[x1,y1,x2,y2], ...])
    if plate_results is None:
        return []
    r = plate_results[j]
    return list(r[1]) if isinstance(r, tuple) else []


def batch_process_frames(frames, face_detector, plate_detector=None, batch_size=16, mode="mosaic", mosaic_level=8,
                         ksize=31, sigma=0.0):
    """The reference frame loop (:204-256) on in-memory RGB frames.  Returns (processed frames list,
    total_faces, total_plates).  Frames of one batch must share a size (as video frames do)."""
    processed, total_faces, total_plates = [], 0, 0
    anon = None
    for i in range(0, len(frames), batch_size):
        batch = frames[i:i + batch_size]
        plate_results = plate_detector(list(batch), verbose=False, conf=0.5) if plate_detector is not None else None
        plates = [[tuple(int(v) for v in b) for b in _plate_boxes(plate_results, j)] for j in range(len(batch))]
        if len({f.shape[:2] for f in batch}) > 1:
            # mixed frame sizes in one batch (the reference letterboxes each frame, face.py:73-80)
            res = face_detector.detect_images(list(batch))
            allb = [[tuple(int(v) for v in b) for b in r[1]] + plates[j] for j, r in enumerate(res)]
            processed.extend(composite_frames(batch, allb, mode, mosaic_level, ksize, sigma))
            total_faces += sum(len(r[1]) for r in res)
            total_plates += sum(len(p) for p in plates)
            continue
        eng = face_detector._engine_for(min(batch_size, len(frames)))
        if anon is None or anon.eng is not eng:
            anon = Anonymizer(eng, mode, mosaic_level, ksize, sigma)
        for s in range(0, len(batch), eng.B):
            out, boxes = anon.process_batch(batch[s:s + eng.B], plates[s:s + eng.B])
            processed.extend(list(out))
            total_faces += sum(len(b) for b in boxes)
        total_plates += sum(len(p) for p in plates)
    return processed, total_faces, total_plates


def load_image_rgb(image_path):
    import cv2
    img = cv2.imread(image_path)
    if img is None:
        raise ValueError(f"cannot read image: {image_path}")
    return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def save_output_image(image_array, output_path):
    import cv2
    cv2.imwrite(output_path, cv2.cvtColor(image_array, cv2.COLOR_RGB2BGR))


def batch_process_images(input_dir, output_dir, face_detector, plate_detector, batch_size=16, mode="mosaic"):
    """Same signature and return value as the reference (:183-277): (n_frames, n_faces, n_plates)."""
    paths = [os.path.join(input_dir, f) for f in os.listdir(input_dir) if f.lower().endswith((".png", ".jpg", ".jpeg"))]
    os.makedirs(output_dir, exist_ok=True)
    total, faces, plates = 0, 0, 0
    with ThreadPoolExecutor(max_workers=6) as pool:
        saves = []
        for i in range(0, len(paths), batch_size):
            files = paths[i:i + batch_size]
            batch = list(pool.map(load_image_rgb, files))
            out, nf, npl = batch_process_frames(batch, face_detector, plate_detector, batch_size, mode)
            for pth, img in zip(files, out):
                saves.append(pool.submit(save_output_image, img, os.path.join(output_dir, "processed_" + os.path.basename(pth))))
            total += len(files)
            faces += nf
            plates += npl
        for s in saves:
            s.result()
    return total, faces, plates
