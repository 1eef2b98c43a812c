"""Device-resident detect + anonymise pipeline for one GPU: H2D of the uint8 frames once, ingest ->
network -> decode/NMS -> compositing on the same resident frames, D2H of the processed frames and
the boxes once.  This is what the drop-in frame loop (combine_detect.batch_process_frames) and
bench.py's end-to-end leg call; it replaces the body of the reference loop
combine_detect.py:204-256 (detect_images, box merge with int() truncation, sequential mosaic)."""
import numpy as np
import torch

from .engine import DetectorEngine


class Anonymizer:
    def __init__(self, engine: DetectorEngine, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0):
        self.eng = engine
        self.mode, self.mosaic_level, self.ksize, self.sigma = mode, int(mosaic_level), int(ksize), float(sigma)
        self._pin_in = self._pin_out = self._dev = None

    def _buffers(self, h, w):
        B = self.eng.B
        if self._dev is None or tuple(self._dev.shape[1:3]) != (h, w):
            self._pin_in = torch.empty((B, h, w, 3), dtype=torch.uint8).pin_memory()
            self._pin_out = torch.empty((B, h, w, 3), dtype=torch.uint8).pin_memory()
            self._dev = torch.empty((B, h, w, 3), dtype=torch.uint8, device=self.eng.device)
        return self._pin_in, self._pin_out, self._dev

    def process_batch(self, frames, ext_boxes=None, composite=True):
        """frames: list of <= B uint8 HxWx3 arrays of ONE size.  ext_boxes: per-frame lists of
        (x1,y1,x2,y2) applied after the detected faces.  Returns (processed frames (n,h,w,3) uint8
        numpy — a fresh array, inputs untouched —, face boxes per frame as lists of [x1,y1,x2,y2]
        floats in original pixels, descending score)."""
        eng = self.eng
        n = len(frames)
        assert 0 < n <= eng.B
        h, w = frames[0].shape[:2]
        pin_in, pin_out, dev = self._buffers(h, w)
        pi = pin_in.numpy()
        for i in range(eng.B):
            pi[i] = frames[min(i, n - 1)]               # tail batch: repeat the last frame
        ext = [list(b) for b in (ext_boxes or [[]] * n)]
        ext += [[] for _ in range(eng.B - n)]
        with torch.cuda.stream(eng.stream):
            dev.copy_(pin_in, non_blocking=True)        # H2D, inside the caller's timed region
        eng.set_frames(dev)
        if composite:
            eng.set_ext_boxes(ext)
        eng.forward()
        if composite:
            eng.enqueue_composite(self.mode, self.mosaic_level, self.ksize, self.sigma)
        with torch.cuda.stream(eng.stream):
            pin_out.copy_(dev, non_blocking=True)       # D2H of the processed frames
        rows, counts = eng.fetch_results()              # D2H of the boxes + sync + status check
        out = pin_out.numpy()[:n].copy()
        boxes = []
        scale = np.array([w, h, w, h], np.float32)
        for i in range(n):
            r = rows[i, :counts[i], :4].copy()
            r *= scale
            boxes.append(r.tolist())
        return out, boxes
