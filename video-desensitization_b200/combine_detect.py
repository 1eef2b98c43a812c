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
import queue
import threading
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
    from vdesens_b200 import _lib
    _lib.mode_id(mode)
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


_warned_plate_type = False


def _plate_boxes(plate_results, j):
    """Boxes of frame j from whatever the plate provider returned.
    reference :239 honours only tuples `(anything, [[x1,y1,x2,y2], ...])` — and therefore silently drops every
    plate when the provider is the Ultralytics YOLO model its own __main__ builds (:868-874), whose per-frame result
    is a `Results` object.  For an anonymiser that is a privacy hole, not a behaviour to reproduce: `Results`
    (anything with `.boxes.xyxy`) is read explicitly, and any OTHER non-tuple result warns once instead of vanishing."""
    global _warned_plate_type
    if plate_results is None:
        return []
    r = plate_results[j]
    if isinstance(r, tuple):
        return list(r[1])
    xyxy = getattr(getattr(r, "boxes", None), "xyxy", None)
    if xyxy is not None:
        if hasattr(xyxy, "detach"):
            xyxy = xyxy.detach().cpu().numpy()
        return [list(map(float, b[:4])) for b in np.asarray(xyxy).reshape(-1, 4)]
    if r is not None and not _warned_plate_type:
        import warnings
        warnings.warn(f"plate_detector returned {type(r).__name__} per frame; expected a tuple (img, boxes) or an "
                      f"Ultralytics Results object — its boxes are NOT applied", RuntimeWarning, stacklevel=2)
        _warned_plate_type = True
    return []


def _plates_of(plate_detector, frames):
    """per-frame integer plate boxes of one batch (reference :217, :239, :243)"""
    res = plate_detector(list(frames), verbose=False, conf=0.5) if plate_detector is not None else None
    return [[tuple(int(v) for v in b) for b in _plate_boxes(res, j)] for j in range(len(frames))]


def batch_process_frames(frames, face_detector, plate_detector=None, batch_size=16, mode="mosaic", mosaic_level=8,
                         ksize=31, sigma=0.0):
    """The reference frame loop (:204-256) on in-memory RGB frames.  Returns (processed frames list,
    total_faces, total_plates).  Frames of one batch must share a size (as video frames do)."""
    from vdesens_b200 import _lib
    _lib.mode_id(mode)
    processed, total_faces, total_plates = [], 0, 0
    anon = None
    for i in range(0, len(frames), batch_size):
        batch = frames[i:i + batch_size]
        plates = _plates_of(plate_detector, batch)
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
    """reference :167-180 (cv2.imread + BGR -> RGB).  OpenCV >= 4.10 can hand the decoder's output over in RGB order
    directly (same bytes, one pass over the frame less); older builds take the reference's two steps."""
    import cv2
    flag = getattr(cv2, "IMREAD_COLOR_RGB", None)
    img = cv2.imread(image_path, flag) if flag is not None else cv2.imread(image_path)
    if img is None:
        raise ValueError(f"cannot read image: {image_path}")
    return img if flag is not None else cv2.cvtColor(img, cv2.COLOR_BGR2RGB)


def save_output_image(image_array, output_path):
    import cv2
    cv2.imwrite(output_path, cv2.cvtColor(image_array, cv2.COLOR_RGB2BGR))


class _Lookahead:
    """iterator with a one-item peek (the frame loop decides per batch which path takes it)"""
    _END = object()

    def __init__(self, it):
        self._it, self._head = iter(it), self._END
        self._advance()

    def _advance(self):
        self._head = next(self._it, self._END)

    def more(self):
        return self._head is not self._END

    def peek(self):
        return self._head

    def pop(self):
        h = self._head
        self._advance()
        return h


def _io_workers():
    """threads for JPEG decode / encode around the device path.  The reference uses 6 of each (:200, :259); at
    ~20 ms per 1080p decode that feeds < 300 frames/s while one GPU takes > 1 500, so use half the host's cores
    each (never fewer than the reference's 6, never more than 32) — cv2 releases the GIL in both calls."""
    return max(6, min(32, (os.cpu_count() or 12) // 2))


def _decoded_batches(groups, workers=None):
    """yields (files, RGB frames) per group; group i+1 is decoded by a thread pool while group i is consumed"""
    workers = workers or _io_workers()
    with ThreadPoolExecutor(max_workers=workers) as dec, ThreadPoolExecutor(max_workers=1) as ahead:
        def load(files):
            return list(dec.map(load_image_rgb, files))
        fut = ahead.submit(load, groups[0]) if groups else None
        for i, files in enumerate(groups):
            frames = fut.result()
            if i + 1 < len(groups):
                fut = ahead.submit(load, groups[i + 1])
            yield files, frames


def _one_shape(frames):
    shapes = {f.shape[:2] for f in frames}
    return next(iter(shapes)) if len(shapes) == 1 else None


def _read_bytes(path):
    with open(path, "rb") as f:
        return f.read()


def _write_bytes(path, blob):
    with open(path, "wb") as f:
        f.write(blob)


def _device_jpeg_loop(groups, output_dir, face_detector, plate_detector, batch_size, mode, mosaic_level, ksize, sigma,
                      quality, failed):
    """The frame loop with the JPEG codec on the device: file bytes -> GPU (Huffman decode, IDCT,
    up-sampling, colour) -> detect + composite on the resident frames -> GPU (DCT, quantisation, Huffman encode) -> file
    bytes.  Pixels equal cv2.imread's and files equal cv2.imwrite's byte for byte, so the output directory is identical
    to the host loop's.  Three batches are in flight: a decode thread (files -> resident frames, plate provider), the
    caller's thread (enqueues ingest / network / NMS / compositing and never synchronises: batch k+1 queues right behind
    batch k on the detector's stream) and an encode thread (its stream waits for the batch's event), each stage on its own
    CUDA stream.  Returns (frames, faces, plates) or None when the first file is not something the device decoder takes
    (the caller then runs the host loop)."""
    from vdesens_b200 import jpeg as vjpeg
    first = _read_bytes(groups[0][0])
    try:
        hdr = vjpeg.prepare_device(first)
    except RuntimeError:
        return None
    H, W = hdr.height, hdr.width
    n_total = sum(len(g) for g in groups)
    eng = face_detector._engine_for(min(batch_size, n_total))
    if (eng.H, eng.W) != (H, W):
        return None                      # letterboxed input: the host path handles it (generic ingest)
    B, dev = eng.B, eng.device
    groups = [g[s:s + B] for g in groups for s in range(0, len(g), B)]
    dec = vjpeg.JpegDecoder(B, device=dev, max_file_bytes=max(4 << 20, 2 * len(first)))
    enc = vjpeg.JpegEncoder(B, H, W, device=dev, quality=quality)
    DEPTH = 3

    class Slot:                                   # one in-flight batch: resident frames + everything that refers to them
        def __init__(self):
            self.frames = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
            self.desc, self.unlb, self.identity = eng.frame_records(self.frames)
            self.eb = torch.zeros((B, eng.max_ext, 4), dtype=torch.int32).pin_memory()
            self.ec = torch.zeros((B,), dtype=torch.int32).pin_memory()
            self.rows = torch.zeros((B, eng.max_det, 16), dtype=torch.float32).pin_memory()
            self.counts = torch.zeros((B,), dtype=torch.int32).pin_memory()
            self.status = torch.zeros((1,), dtype=torch.int32).pin_memory()
            self.ev_done = torch.cuda.Event()
            self.free = threading.Event()
            self.free.set()

    with torch.cuda.device(dev):
        slots = [Slot() for _ in range(DEPTH)]
    totals = [0, 0, 0]                            # frames, faces, plates
    errors = []

    def host_batch(files):
        """one batch through the host codec (a file the device decoder cannot take alongside the others, an oversized
        file, ...): same results, cv2 does the JPEG work"""
        frames = [load_image_rgb(p) for p in files]
        plates = _plates_of(plate_detector, frames)
        res = face_detector.detect_images(frames)
        allb = [[tuple(int(v) for v in b) for b in r[1]] + plates[j] for j, r in enumerate(res)]
        for pth, img in zip(files, composite_frames(frames, allb, mode, mosaic_level, ksize, sigma)):
            save_output_image(img, os.path.join(output_dir, "processed_" + os.path.basename(pth)))
        return len(files), sum(len(r[1]) for r in res), sum(len(p) for p in plates)

    q_in, q_out = queue.Queue(maxsize=DEPTH), queue.Queue(maxsize=DEPTH)
    with ThreadPoolExecutor(max_workers=_io_workers()) as fpool:

        def decode_stage():
            """files -> resident frames (+ the plate boxes: the provider wants HOST frames, so it gets one D2H of the decoded
            batch here, off the detector's thread)"""
            try:
                for k, files in enumerate(groups):
                    slot = slots[k % DEPTH]
                    slot.free.wait()
                    slot.free.clear()
                    item = (k, None, None)
                    try:
                        blobs = list(fpool.map(_read_bytes, files))
                        with torch.cuda.device(dev):
                            frames, _ = dec.decode(blobs, out=slot.frames)
                            plates = [[] for _ in files]
                            if plate_detector is not None:
                                plates = _plates_of(plate_detector, list(frames.cpu().numpy()))
                        item = (k, plates, None)
                    except Exception as e:        # noqa: BLE001: this batch goes through the host codec
                        item = (k, None, e)
                    q_in.put(item)
            finally:
                q_in.put(None)

        def encode_stage():
            """resident, composited frames -> files; nothing here waits for the detector's stream on the host: the encoder's
            stream waits for the batch's event"""
            while True:
                item = q_out.get()
                if item is None:
                    return
                k, n, n_plates = item
                slot, files = slots[k % DEPTH], groups[k]
                try:
                    with torch.cuda.device(dev):
                        enc.stream.wait_event(slot.ev_done)
                        enc.enqueue(slot.desc)
                        blobs = enc.fetch(n)
                        eng.check_status(int(slot.status[0]))       # ev_done precedes the encoder's work: the rows are here
                    list(fpool.map(_write_bytes, [os.path.join(output_dir, "processed_" + os.path.basename(p)) for p in files], blobs))
                    totals[0] += n
                    totals[1] += int(slot.counts[:n].sum())
                    totals[2] += n_plates
                except Exception as e:            # noqa: BLE001 (reference :226)
                    errors.append((f"batch of {n} frames", e))
                finally:
                    slot.free.set()

        td, te = threading.Thread(target=decode_stage, daemon=True), threading.Thread(target=encode_stage, daemon=True)
        td.start()
        te.start()
        try:
            while True:
                item = q_in.get()
                if item is None:
                    break
                k, plates, err = item
                slot, files = slots[k % DEPTH], groups[k]
                n = len(files)
                if err is not None:
                    # the host path reuses the engine: let the batches already queued on it finish first
                    try:
                        eng.stream.synchronize()
                        a, b, c = host_batch(files)
                        totals[0] += a
                        totals[1] += b
                        totals[2] += c
                    except Exception as e:        # noqa: BLE001 (reference :226)
                        errors.append((f"batch of {n} frames", e))
                    slot.free.set()
                    continue
                # detector: enqueue only — no host synchronisation on this thread, batch k+1 queues right behind batch k
                with torch.cuda.device(dev):
                    eb, ec = slot.eb.numpy(), slot.ec.numpy()
                    ec[:] = 0
                    for i, bl in enumerate(plates):
                        if len(bl) > eng.max_ext:
                            raise RuntimeError(f"{len(bl)} external boxes exceed max_ext={eng.max_ext}")
                        ec[i] = len(bl)
                        for j, bx in enumerate(bl):
                            eb[i, j] = bx
                    eng.bind_frames_async(slot.desc, slot.unlb, slot.identity)
                    eng.upload_ext_async(slot.eb, slot.ec)
                    eng.forward()
                    eng.enqueue_composite(mode, mosaic_level, ksize, sigma)
                    eng.download_results_async(slot.rows, slot.counts, slot.status)
                    slot.ev_done.record(eng.stream)
                q_out.put((k, n, sum(len(p) for p in plates)))
        finally:
            q_out.put(None)
            te.join()
            td.join(timeout=5.0)
    for what, e in errors:
        failed(what, e)
    return tuple(totals)


def batch_process_images(input_dir, output_dir, face_detector, plate_detector, batch_size=16, mode="mosaic",
                         mosaic_level=8, ksize=31, sigma=0.0, sort=True, on_error="raise", io="auto", jpeg_quality=95):
    """Same signature and return value as the reference (:183-277): (n_frames, n_faces, n_plates).

    What changes is the schedule: the reference loads a batch, waits for both detectors,
    mosaics on the CPU and only then loads the next batch.  Here batch i+1 is decoded by a thread pool while the
    GPU processes batch i through the pipelined StreamingAnonymizer (H2D / compute / D2H overlapped) and the
    saver pool encodes batch i-1.  The plate provider runs on its own worker thread, concurrently with the face
    detector of the same batch, as in the reference (:214-217); its boxes are only waited for when the compositing
    pass is enqueued.

    Two documented departures, each with a switch back to the reference's behaviour:
      sort=True         files are processed in sorted name order (the reference takes os.listdir order, :186-189,
                        which is arbitrary); sort=False restores that.
      on_error="raise"  a failing batch raises (the reference logs and silently DROPS the batch, :226-228, so frames
                        go missing from the output video); on_error="skip" restores that: the batches in flight when
                        the error surfaced are dropped with a warning and the loop continues with the next one.
    io: where the JPEG codec runs.  "device": files are decoded and re-encoded on the GPU
    (vdesens_b200.jpeg: pixels identical to cv2.imread's, files byte-identical to cv2.imwrite's at `jpeg_quality`, cv2's
    default 95) — what crosses PCIe is the file, not the pixels; "host": cv2 thread pools as in the reference; "auto"
    (default): "device" when every input is a .jpg / .jpeg file and the first one is something the device decoder
    takes at the network's own size, else "host"."""
    from collections import deque
    from vdesens_b200 import _lib
    from vdesens_b200.pipeline import StreamingAnonymizer
    _lib.mode_id(mode)
    if on_error not in ("raise", "skip"):
        raise ValueError(f"on_error must be 'raise' or 'skip', got {on_error!r}")
    if io not in ("auto", "device", "host"):
        raise ValueError(f"io must be 'auto', 'device' or 'host', got {io!r}")
    names = [f for f in os.listdir(input_dir) if f.lower().endswith((".png", ".jpg", ".jpeg"))]
    if sort:
        names.sort()
    paths = [os.path.join(input_dir, f) for f in names]
    os.makedirs(output_dir, exist_ok=True)
    groups = [paths[i:i + batch_size] for i in range(0, len(paths), batch_size)]
    n_frames = n_faces = n_plates = 0

    def failed(what, exc):
        if on_error == "raise":
            raise exc
        import warnings
        warnings.warn(f"batch_process_images: {what} dropped after {type(exc).__name__}: {exc}", RuntimeWarning)

    if io != "host" and groups and all(p.lower().endswith((".jpg", ".jpeg")) for p in paths):
        done = _device_jpeg_loop(groups, output_dir, face_detector, plate_detector, batch_size, mode, mosaic_level,
                                 ksize, sigma, jpeg_quality, failed)
        if done is not None:
            return done
        if io == "device":
            raise RuntimeError("io='device': the first file is not a baseline JPEG of the network's input size")
    elif io == "device":
        raise RuntimeError("io='device' needs a directory of .jpg / .jpeg files")

    with ThreadPoolExecutor(max_workers=_io_workers()) as savers, ThreadPoolExecutor(max_workers=1) as plate_pool:
        saves = []

        def emit(files, out):
            for pth, img in zip(files, out):
                saves.append(savers.submit(save_output_image, img,
                                           os.path.join(output_dir, "processed_" + os.path.basename(pth))))

        src = _Lookahead(_decoded_batches(groups))
        while src.more():
            shape = _one_shape(src.peek()[1])
            if shape is None:
                # mixed frame sizes inside one batch: every frame is letterboxed on its own (face.py:73-80)
                files, frames = src.pop()
                try:
                    pf = plate_pool.submit(_plates_of, plate_detector, frames)      # sibling thread, as the reference
                    res = face_detector.detect_images(frames)
                    plates = pf.result()
                    allb = [[tuple(int(v) for v in b) for b in r[1]] + plates[j] for j, r in enumerate(res)]
                    emit(files, composite_frames(frames, allb, mode, mosaic_level, ksize, sigma))
                except Exception as e:                                               # noqa: BLE001 (reference :226)
                    failed(f"batch of {len(files)} frames", e)
                    continue
                n_frames += len(files)
                n_faces += sum(len(r[1]) for r in res)
                n_plates += sum(len(p) for p in plates)
                continue
            # a run of consecutive batches of one frame size goes through the pipelined path
            eng = face_detector._engine_for(min(batch_size, len(paths)))
            sa = StreamingAnonymizer(eng, mode, mosaic_level, ksize, sigma)
            inflight = deque()                       # (files, Future of the plate boxes) per submitted device batch

            def device_batches():
                while src.more() and _one_shape(src.peek()[1]) == shape:
                    files, frames = src.pop()
                    for s in range(0, len(frames), eng.B):
                        part = frames[s:s + eng.B]
                        inflight.append((files[s:s + eng.B], plate_pool.submit(_plates_of, plate_detector, part)))
                        yield part

            def plate_boxes():                       # run() asks for these right after it took the batch
                while True:
                    yield inflight[-1][1]

            try:
                for out, boxes in sa.run(device_batches(), plate_boxes(), copy_out=True):
                    files, pf = inflight.popleft()
                    plates = pf.result()
                    emit(files, list(out))
                    n_frames += len(files)
                    n_faces += sum(len(b) for b in boxes)
                    n_plates += sum(len(p) for p in plates)
            except Exception as e:                                                   # noqa: BLE001 (reference :226)
                failed(f"{len(inflight)} batch(es) in flight", e)
                inflight.clear()
        for f in saves:
            f.result()
    return n_frames, n_faces, n_plates
