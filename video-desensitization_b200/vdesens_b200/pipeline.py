"""Device-resident detect + anonymise pipeline for one GPU: H2D of the uint8 frames once, ingest ->
network -> decode/NMS -> compositing on the same resident frames, D2H of the processed frames and
the boxes once.  This is what the drop-in frame loop (combine_detect.batch_process_frames) and
bench.py's end-to-end leg call; it replaces the body of the reference loop
combine_detect.py:204-256 (detect_images, box merge with int() truncation, sequential mosaic)."""
import numpy as np
import torch

from . import _lib
from .engine import DetectorEngine


class Anonymizer:
    def __init__(self, engine: DetectorEngine, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0):
        self.eng = engine
        _lib.mode_id(mode)                     # ValueError on anything but "mosaic" / "gauss"
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


class _Slot:
    """one in-flight batch of the pipelined path: pinned staging, resident frames, result mirrors"""

    def __init__(self, eng, h, w):
        B, dev = eng.B, eng.device
        self.pin_in = torch.empty((B, h, w, 3), dtype=torch.uint8).pin_memory()
        self.pin_out = torch.empty((B, h, w, 3), dtype=torch.uint8).pin_memory()
        self.dev = torch.empty((B, h, w, 3), dtype=torch.uint8, device=dev)
        self.desc, self.unlb, self.identity = eng.frame_records(self.dev)
        self.eb = torch.zeros((B, eng.max_ext, 4), dtype=torch.int32).pin_memory()
        self.ec = torch.zeros((B,), dtype=torch.int32).pin_memory()
        self.rows = torch.zeros((B, eng.max_det, 16), dtype=torch.float32).pin_memory()
        self.counts = torch.zeros((B,), dtype=torch.int32).pin_memory()
        self.status = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self.ev_in0 = torch.cuda.Event(enable_timing=True)     # H2D start / end, compute end, D2H start / end
        self.ev_in = torch.cuda.Event(enable_timing=True)
        self.ev_done = torch.cuda.Event(enable_timing=True)
        self.ev_out0 = torch.cuda.Event(enable_timing=True)
        self.ev_out = torch.cuda.Event(enable_timing=True)
        self.n = 0
        self.ext = None
        self.busy = False
        self.src = None                                        # pinned tensor the batch was copied from
        self.host_desc = {}                                    # data_ptr -> device (B,6) int32 records of the HOST frames
        # d2h="dirty": the boxes of THIS batch, kept where the next batch's forward cannot overwrite them
        self.rows_dev = torch.zeros((B, eng.max_det, 16), dtype=torch.float32, device=dev)
        self.count_dev = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.eb_dev = torch.zeros((B, eng.max_ext, 4), dtype=torch.int32, device=dev)
        self.ec_dev = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.dirty_bytes = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.dirty_bytes_host = torch.zeros((1,), dtype=torch.int64).pin_memory()


class StreamingAnonymizer:
    """Pipelined detect + anonymise over a stream of equally sized batches on one GPU: the H2D copy of
    batch i+1 and the D2H copy of batch i-1 run on their own streams while batch i computes, so the
    end-to-end rate approaches the device rate (PCIe moves 2 x 100 MB per 16-frame 1080p batch).
    Results are yielded strictly in input order."""

    def __init__(self, engine: DetectorEngine, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0, depth=2, d2h="full"):
        """d2h="full": the processed frames come back as whole frames in a pinned buffer of the pipeline (inputs untouched).
        d2h="dirty": only the 64 x 16-pixel tiles a box touched travel back, written by the GPU straight INTO the pinned
        host frames the batch came from (vd_copy_dirty_tiles): the yielded frames are that buffer, patched in place —
        ~44 % of the bytes on the bench workload.  Needs pinned inputs (a pinned tensor, or the pipeline's own staging
        buffer when frames are given as numpy arrays)."""
        self.eng = engine
        _lib.mode_id(mode)                     # ValueError on anything but "mosaic" / "gauss"
        if d2h not in ("full", "dirty"):
            raise ValueError(f"d2h must be 'full' or 'dirty', got {d2h!r}")
        self.d2h = d2h
        self.mode, self.mosaic_level, self.ksize, self.sigma = mode, int(mosaic_level), int(ksize), float(sigma)
        self.depth = max(2, int(depth))
        self.slots = None
        # host seconds per stage + device-timed copy durations (CUDA events on the copy streams) + yield timestamps
        self.stats = {"submit_s": 0.0, "collect_wait_s": 0.0, "collect_host_s": 0.0, "batches": 0,
                      "h2d_ms": 0.0, "d2h_ms": 0.0, "h2d_bytes": 0, "d2h_bytes": 0, "yield_t": []}
        self.s_in = torch.cuda.Stream(device=engine.device)
        self.s_out = torch.cuda.Stream(device=engine.device)

    def _ensure(self, h, w):
        if self.slots is None or tuple(self.slots[0].dev.shape[1:3]) != (h, w):
            self.slots = [_Slot(self.eng, h, w) for _ in range(self.depth)]

    def _submit(self, slot, batch, ext):
        import time
        t0 = time.perf_counter()
        self._submit_impl(slot, batch, ext)
        self.stats["submit_s"] += time.perf_counter() - t0
        self.stats["batches"] += 1

    def _submit_impl(self, slot, batch, ext):
        eng = self.eng
        if isinstance(batch, torch.Tensor):                       # (n,h,w,3) uint8, ideally pinned
            n = batch.shape[0]
            src = batch if (n == eng.B and (self.d2h == "full" or batch.is_pinned())) else None
            if src is None:
                slot.pin_in[:n].copy_(batch)
                slot.pin_in[n:] = slot.pin_in[n - 1]
                src = slot.pin_in
        else:                                                     # list of numpy frames
            n = len(batch)
            pi = slot.pin_in.numpy()
            for i in range(eng.B):
                pi[i] = batch[min(i, n - 1)]
            src = slot.pin_in
        slot.n = n
        slot.src = src
        with torch.cuda.stream(self.s_in):
            slot.ev_in0.record(self.s_in)
            slot.dev.copy_(src, non_blocking=True)                # H2D
            slot.ev_in.record(self.s_in)
        eng.stream.wait_event(slot.ev_in)
        eng.bind_frames_async(slot.desc, slot.unlb, slot.identity)
        eng.forward()
        # The external (plate) boxes are only needed by the compositing pass, so a provider that is still running —
        # a Future, as combine_detect hands over: the reference runs the plate model on a sibling thread of the face
        # detector (combine_detect.py:214-217) — is waited for AFTER the detector has been enqueued.
        if hasattr(ext, "result"):
            ext = ext.result()
        eb, ec = slot.eb.numpy(), slot.ec.numpy()
        ec[:] = 0
        for i, bl in enumerate(ext or []):
            if len(bl) > eng.max_ext:
                raise RuntimeError(f"{len(bl)} external boxes exceed max_ext={eng.max_ext}")
            ec[i] = len(bl)
            for j, b in enumerate(bl):
                eb[i, j] = [int(v) for v in b]
        slot.ext = ext
        eng.upload_ext_async(slot.eb, slot.ec)
        eng.enqueue_composite(self.mode, self.mosaic_level, self.ksize, self.sigma)
        eng.download_results_async(slot.rows, slot.counts, slot.status)
        if self.d2h == "dirty":
            with torch.cuda.stream(eng.stream):
                slot.rows_dev.copy_(eng.out_rows, non_blocking=True)
                slot.count_dev.copy_(eng.out_count, non_blocking=True)
                slot.eb_dev.copy_(eng.ext_boxes, non_blocking=True)
                slot.ec_dev.copy_(eng.ext_count, non_blocking=True)
        slot.ev_done.record(eng.stream)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot.ev_done)
            slot.ev_out0.record(self.s_out)
            if self.d2h == "dirty":
                self._copy_dirty(slot)                            # changed tiles -> the pinned frames they came from
            else:
                slot.pin_out.copy_(slot.dev, non_blocking=True)   # D2H of the processed frames
            slot.ev_out.record(self.s_out)
        # (the slot is only refilled after _collect() has synchronised on ev_out, so the next H2D into it
        #  cannot overtake this D2H; no wait is queued on s_in — it would also stall the OTHER slot's H2D)
        slot.busy = True

    def _copy_dirty(self, slot):
        """enqueue vd_copy_dirty_tiles on the current (= copy-out) stream: device frames -> the slot's pinned source"""
        import ctypes as C
        import numpy as np
        from .engine import FRAME_DTYPE
        eng, src = self.eng, slot.src
        B, h, w = eng.B, int(src.shape[1]), int(src.shape[2])
        key = src.data_ptr()
        if key not in slot.host_desc:
            rec = np.zeros((B,), FRAME_DTYPE)
            for i in range(B):
                rec[i] = (key + i * h * w * 3, h, w, w * 3, 0)
            if len(slot.host_desc) > 16:
                slot.host_desc.clear()
            slot.host_desc[key] = torch.from_numpy(rec.view(np.int32).reshape(B, 6)).to(eng.device)
        s = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        slot.dirty_bytes.zero_()
        _lib.check(eng.lib.vd_copy_dirty_tiles(C.c_void_p(slot.desc.data_ptr()), C.c_void_p(slot.host_desc[key].data_ptr()), B, h, w,
                                               C.c_void_p(slot.rows_dev.data_ptr()), C.c_void_p(slot.count_dev.data_ptr()), eng.max_det,
                                               C.c_void_p(slot.eb_dev.data_ptr()), C.c_void_p(slot.ec_dev.data_ptr()), eng.max_ext,
                                               C.c_void_p(slot.dirty_bytes.data_ptr()), s), "vd_copy_dirty_tiles", eng.lib)
        slot.dirty_bytes_host.copy_(slot.dirty_bytes, non_blocking=True)

    def _collect(self, slot, copy_out):
        import time
        t0 = time.perf_counter()
        slot.ev_out.synchronize()
        t1 = time.perf_counter()
        out = self._collect_impl(slot, copy_out)
        st = self.stats
        st["collect_wait_s"] += t1 - t0
        st["collect_host_s"] += time.perf_counter() - t1
        st["h2d_ms"] += slot.ev_in0.elapsed_time(slot.ev_in)
        st["d2h_ms"] += slot.ev_out0.elapsed_time(slot.ev_out)
        st["h2d_bytes"] += slot.dev.numel()
        st["d2h_bytes"] += int(slot.dirty_bytes_host[0]) if self.d2h == "dirty" else slot.dev.numel()
        st["yield_t"].append(time.perf_counter())
        return out

    def _collect_impl(self, slot, copy_out):
        slot.busy = False
        self.eng.check_status(int(slot.status[0]))
        n = slot.n
        h, w = slot.dev.shape[1:3]
        rows, counts = slot.rows.numpy(), slot.counts.numpy()
        scale = np.array([w, h, w, h], np.float32)
        boxes = []
        for i in range(n):
            r = rows[i, :counts[i], :4].copy()
            r *= scale
            boxes.append(r.tolist())
        out = (slot.src if self.d2h == "dirty" else slot.pin_out)[:n]
        return (out.numpy().copy() if copy_out else out), boxes

    def run(self, batches, ext_boxes=None, copy_out=False):
        """batches: iterable of batches — a pinned uint8 tensor (n,h,w,3) (zero extra host copies) or a
        list of numpy frames.  ext_boxes: optional iterable aligned with `batches` of per-frame box lists.
        Yields (processed frames, face boxes per frame) in input order.  With copy_out=False the frames
        are a view of a pinned buffer that stays valid until `depth` more batches have been yielded."""
        ext_iter = iter(ext_boxes) if ext_boxes is not None else None
        pending = []
        k = 0
        for batch in batches:
            h, w = (batch.shape[1:3] if isinstance(batch, torch.Tensor) else batch[0].shape[:2])
            self._ensure(int(h), int(w))
            slot = self.slots[k % self.depth]
            if slot.busy:                                         # oldest in-flight batch owns this slot
                yield self._collect(pending.pop(0), copy_out)
            self._submit(slot, batch, next(ext_iter) if ext_iter is not None else None)
            pending.append(slot)
            k += 1
        while pending:
            yield self._collect(pending.pop(0), copy_out)
