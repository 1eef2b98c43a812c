"""Drop-in for the hot-path surface of the reference's combine_detect.py, backed by libvdesens:

  mosaic_rectangle_region_single(img, x1, y1, x2, y2, mosaic_level=8)   reference :138-161
  batch_process_images(input_dir, output_dir, face_detector, plate_detector, batch_size=16)  :183-277
plus what BASELINE.json's north_star adds:
  gaussian_rectangle_region_single(...)   the cv2.GaussianBlur ROI idiom, integer-exact
  composite_frames(...)                   batched sequential compositing on the device
  batch_process_frames(...)               the frame loop on in-memory frames (no JPEG hand-off)

and the video stage either side of it, same signatures and return values as the reference:
  process_video_pipeline(input, output, face_detector, plate_detector, temp_dir, fps, batch_size)   :597-646
  process_single_video(...) :648-699, convert_video_to_frames(...) :279-476, create_video(...) :479-595
  Motion-JPEG AVI runs demux -> device decode -> detect + composite -> device encode -> mux with the frames resident in
  HBM (vdesens_b200/avi.py + vdesens_b200/jpeg.py); other codecs go through OpenCV's libavcodec at that end.

Record unpacking, config, logging and the CLI (reference :21-89, :701-990) are out of scope.
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


class _JpegSource:
    """keys -> JPEG file bytes -> device decoder (pixels identical to cv2.imread's)"""

    def __init__(self, read, dec):
        self.read, self.dec = read, dec

    def fill(self, keys, slot):
        """decodes the batch into slot.frames; returns the host copy of the frames if it has one (else None)"""
        self.dec.decode(self.read(keys), out=slot.frames)
        return None

    def host_frames(self, keys):
        return [self.dec._fallback(b) for b in self.read(keys)]


class _FrameSource:
    """keys -> host RGB frames (a codec library decoded them) -> pinned staging -> resident frames"""

    def __init__(self, read, B, H, W, dev):
        self.read = read
        self.stage = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)

    def fill(self, keys, slot):
        frames = self.read(keys)
        n = len(frames)
        st = self.stage.numpy()
        for i, f in enumerate(frames):
            st[i] = f
        with torch.cuda.stream(self.stream):
            slot.frames[:n].copy_(self.stage[:n], non_blocking=True)
        self.stream.synchronize()
        return frames

    def host_frames(self, keys):
        return self.read(keys)


class _JpegSink:
    """resident frames -> device encoder -> write(keys, files' bytes) (bytes identical to cv2.imwrite's)"""

    def __init__(self, write, write_frames, enc):
        self.write, self.host_emit, self.enc = write, write_frames, enc

    def emit(self, keys, slot, n):
        enc = self.enc
        enc.stream.wait_event(slot.ev_done)
        enc.enqueue(slot.desc)
        self.write(keys, enc.fetch(n))


class _FrameSink:
    """resident frames -> pinned host frames -> write_frames(keys, RGB arrays) (a codec library encodes them)"""

    def __init__(self, write_frames, B, H, W, dev):
        self.host_emit = write_frames
        self.stage = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)

    def emit(self, keys, slot, n):
        self.stream.wait_event(slot.ev_done)
        with torch.cuda.stream(self.stream):
            self.stage[:n].copy_(slot.frames[:n], non_blocking=True)
        self.stream.synchronize()
        self.host_emit(keys, [self.stage[i].numpy() for i in range(n)])


def _resident_loop(groups, source, sink, eng, frame_hw, face_detector, plate_detector, mode, mosaic_level, ksize, sigma,
                   failed):
    """The frame loop on frames that stay in HBM from the decoder to the encoder.  `groups`
    is any iterable of key lists (<= eng.B keys each: file names, frame numbers, ...); `source` puts a group's frames
    into a slot's resident buffer, `sink` takes them out again after detect + composite.  Three batches are in flight: a
    decode thread (source, plate provider), the caller's thread (enqueues ingest / network / NMS / compositing and
    never synchronises: batch k+1 queues right behind batch k on the detector's stream) and an encode thread (its
    stream waits for the batch's event), each stage on its own CUDA stream.  The sink sees the batches in order.  A
    batch whose source fails goes through the host codec (source.host_frames -> detect_images -> sink.host_emit), in
    order as well.  Returns (frames, faces, plates)."""
    H, W = frame_hw
    B, dev = eng.B, eng.device
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

    def host_batch(keys):
        """one batch through the host codec (a file the device decoder cannot take alongside the others, an oversized
        file, ...): same results, cv2 does the JPEG work.  -> (processed RGB frames, faces, plates)"""
        frames = source.host_frames(keys)
        plates = _plates_of(plate_detector, frames)
        res = face_detector.detect_images(frames)
        allb = [[tuple(int(v) for v in b) for b in r[1]] + plates[j] for j, r in enumerate(res)]
        return (composite_frames(frames, allb, mode, mosaic_level, ksize, sigma), sum(len(r[1]) for r in res),
                sum(len(p) for p in plates))

    q_in, q_out = queue.Queue(maxsize=DEPTH), queue.Queue(maxsize=DEPTH)

    def decode_stage():
        """keys -> resident frames (+ the plate boxes: the provider wants HOST frames, so it gets one D2H of the decoded
        batch here — unless the source had them on the host anyway — off the detector's thread)"""
        try:
            for k, keys in enumerate(groups):
                slot = slots[k % DEPTH]
                slot.free.wait()
                slot.free.clear()
                item = (k, keys, None, None)
                try:
                    with torch.cuda.device(dev):
                        host = source.fill(keys, slot)
                        plates = [[] for _ in keys]
                        if plate_detector is not None:
                            if host is None:
                                host = list(slot.frames[:len(keys)].cpu().numpy())
                            plates = _plates_of(plate_detector, host)
                    item = (k, keys, plates, None)
                except Exception as e:            # noqa: BLE001: this batch goes through the host codec
                    item = (k, keys, None, e)
                q_in.put(item)
        except Exception as e:                    # noqa: BLE001: the group iterator itself failed (a broken container)
            errors.append(("frame source", e))
        finally:
            q_in.put(None)

    def encode_stage():
        """resident, composited frames -> sink; nothing here waits for the detector's stream on the host: the sink's
        stream waits for the batch's event"""
        while True:
            item = q_out.get()
            if item is None:
                return
            k, keys, n, n_faces, n_plates, host_frames = item
            slot = slots[k % DEPTH]
            try:
                if host_frames is not None:       # the batch took the host path; only its place in the order is kept here
                    sink.host_emit(keys, host_frames)
                else:
                    with torch.cuda.device(dev):
                        sink.emit(keys, slot, n)
                        eng.check_status(int(slot.status[0]))       # ev_done precedes the sink's work: the rows are here
                    n_faces = int(slot.counts[:n].sum())
                totals[0] += n
                totals[1] += n_faces
                totals[2] += n_plates
            except Exception as e:                # noqa: BLE001 (reference :226)
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
            k, keys, plates, err = item
            slot = slots[k % DEPTH]
            n = len(keys)
            if err is not None:
                # the host path reuses the engine: let the batches already queued on it finish first
                try:
                    eng.stream.synchronize()
                    out, n_faces, n_plates = host_batch(keys)
                    q_out.put((k, keys, n, n_faces, n_plates, out))
                except Exception as e:            # noqa: BLE001 (reference :226)
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
            q_out.put((k, keys, n, 0, sum(len(p) for p in plates), None))
    finally:
        q_out.put(None)
        te.join()
        td.join(timeout=5.0)
    for what, e in errors:
        failed(what, e)
    return tuple(totals)


def _device_jpeg_loop(groups, output_dir, face_detector, plate_detector, batch_size, mode, mosaic_level, ksize, sigma,
                      quality, failed):
    """The directory loop with the JPEG codec on the device: file bytes -> GPU (Huffman decode, IDCT,
    up-sampling, colour) -> detect + composite on the resident frames -> GPU (DCT, quantisation, Huffman encode) -> file
    bytes.  Pixels equal cv2.imread's and files equal cv2.imwrite's byte for byte, so the output directory is identical
    to the host loop's.  Returns (frames, faces, plates) or None when the first file is not something the device decoder
    takes (the caller then runs the host loop)."""
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

    def out_path(p):
        return os.path.join(output_dir, "processed_" + os.path.basename(p))

    with ThreadPoolExecutor(max_workers=_io_workers()) as fpool:
        source = _JpegSource(lambda files: list(fpool.map(_read_bytes, files)), dec)
        source.host_frames = lambda files: [load_image_rgb(p) for p in files]
        sink = _JpegSink(lambda files, blobs: list(fpool.map(_write_bytes, [out_path(p) for p in files], blobs)),
                         lambda files, frames: [save_output_image(img, out_path(p)) for p, img in zip(files, frames)], enc)
        return _resident_loop(groups, source, sink, eng, (H, W), face_detector, plate_detector, mode, mosaic_level, ksize,
                              sigma, failed)


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


# ----------------------------------------------------------------------------------------------------------------
# Video stage (reference :279-476 convert_video_to_frames, :479-595 create_video, :597-699)
# ----------------------------------------------------------------------------------------------------------------
# reference :505-512: container extension -> fourcc candidates, first one cv2.VideoWriter opens wins
_CODEC_PRIORITY = {
    ".mp4": ["mp4v", "avc1", "X264"],
    ".m4v": ["mp4v", "avc1"],
    ".avi": ["XVID", "MJPG"],
    ".h265": ["HEVC", "avc1"],
    ".hevc": ["HEVC", "avc1"],
    ".mkv": ["mp4v", "avc1", "XVID"],
}


def _open_mjpeg_avi(path):
    """AviReader on `path` if it is a Motion-JPEG AVI with at least one frame, else None"""
    from vdesens_b200 import avi
    try:
        rd = avi.AviReader(path)
    except (OSError, ValueError):
        return None
    if rd.is_mjpeg and len(rd) and rd.frame(0)[:2] == b"\xff\xd8":
        return rd
    rd.close()
    return None


def _open_cv2_writer(output_path, fps, width, height):
    """cv2.VideoWriter by the reference's rules (:497-571): codec candidates by extension, unknown extension -> .mp4,
    last resort mp4v into '<name>_backup.mp4'.  -> (writer, path actually written) or (None, None)"""
    import cv2
    base, ext = os.path.splitext(output_path)
    ext = ext.lower()
    if ext not in _CODEC_PRIORITY:
        ext, output_path = ".mp4", base + ".mp4"
    for cc in _CODEC_PRIORITY[ext]:
        wr = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*cc), fps, (width, height))
        if wr.isOpened():
            return wr, output_path
        wr.release()
    backup = base + "_backup.mp4"
    wr = cv2.VideoWriter(backup, cv2.VideoWriter_fourcc(*"mp4v"), fps, (width, height))
    if wr.isOpened():
        return wr, backup
    wr.release()
    for pth in (output_path, backup):
        if os.path.exists(pth) and os.path.getsize(pth) == 0:
            os.remove(pth)
    return None, None


def process_video_pipeline(input_video_path, output_video_path, face_detector, plate_detector,
                           temp_dir="temp_processing", fps=60, batch_size=16, mode="mosaic", mosaic_level=8, ksize=31,
                           sigma=0.0, jpeg_quality=95, on_error="raise"):
    """Same signature and return value as the reference (:597-646): (success, frames, faces, plates).

    The reference goes video -> ffmpeg -> JPEG files -> frame loop -> JPEG files -> cv2.VideoWriter -> video, through
    `temp_dir`.  Here the frames stay in HBM between the two ends (`_resident_loop`) and nothing is written to `temp_dir`
    (accepted and ignored):
      * a Motion-JPEG AVI input is demuxed in place (vdesens_b200.avi.AviReader) and its frames — complete JPEG files —
        are decoded on the device; any other input is decoded by cv2.VideoCapture (the same libavcodec the reference's
        ffmpeg subprocess uses) and uploaded from pinned memory;
      * an `.avi` output is written as Motion-JPEG: the device encoder's bytes go straight into the container
        (vdesens_b200.avi.AviWriter; MJPG is the second entry of the reference's `.avi` list, :505).  Frame i of that
        file IS, byte for byte, the `processed_frame_%06d.jpg` the reference's loop writes for the same pixels
        (cv2.imwrite at quality 95), so decoding the video returns exactly the arrays the reference hands to
        `VideoWriter.write` (:575-583) — without the second lossy generation that call adds.  Other extensions are
        encoded by cv2.VideoWriter with the reference's codec list (frames come back through pinned memory).
    `fps` is the OUTPUT rate as in the reference (which always passes 60, :676); fps=None keeps the source's rate.
    Frames whose size is not the detector's input_shape are letterboxed on the device (reference face.py:73-80)."""
    import cv2
    from vdesens_b200 import _lib, avi
    from vdesens_b200 import jpeg as vjpeg
    _lib.mode_id(mode)
    if on_error not in ("raise", "skip"):
        raise ValueError(f"on_error must be 'raise' or 'skip', got {on_error!r}")
    if not getattr(face_detector, "letterbox_image", True):
        raise ValueError("Batch inference requires letterbox_image=True for shape alignment.")     # reference face.py:79

    def failed(what, exc):
        if on_error == "raise":
            raise exc
        import warnings
        warnings.warn(f"process_video_pipeline: {what} dropped after {type(exc).__name__}: {exc}", RuntimeWarning)

    rd = cap = wr = avi_out = None
    try:
        # ---- source ---------------------------------------------------------------------------------------------
        rd = _open_mjpeg_avi(input_video_path)
        if rd is not None:
            try:
                hdr = vjpeg.prepare_device(rd.frame(0))
                H, W, n_frames, src_fps = hdr.height, hdr.width, len(rd), rd.fps
            except RuntimeError:                  # progressive / restart markers / odd sampling: libavcodec takes it
                rd.close()
                rd = None
        if rd is None:
            cap = cv2.VideoCapture(input_video_path)
            if not cap.isOpened():
                return False, 0, 0, 0             # reference :614-616 (extraction produced no frames)
            W, H = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
            n_frames, src_fps = int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), cap.get(cv2.CAP_PROP_FPS)
            if W <= 0 or H <= 0:
                return False, 0, 0, 0
        out_fps = float(fps) if fps else float(src_fps or 60)
        eng = face_detector._engine_for(min(batch_size, n_frames) if n_frames > 0 else batch_size)
        B, dev = eng.B, eng.device

        # ---- sink -----------------------------------------------------------------------------------------------
        os.makedirs(os.path.dirname(os.path.abspath(output_video_path)), exist_ok=True)
        if os.path.splitext(output_video_path)[1].lower() == ".avi":
            avi_out = avi.AviWriter(output_video_path, W, H, out_fps)
            enc = vjpeg.JpegEncoder(B, H, W, device=dev, quality=jpeg_quality)

            def write_blobs(_keys, blobs):
                for b in blobs:
                    avi_out.write(b)

            def write_frames_as_jpeg(_keys, frames):
                for f in frames:
                    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(f, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, jpeg_quality])
                    if not ok:
                        raise RuntimeError("cv2.imencode failed")
                    avi_out.write(buf.tobytes())

            sink = _JpegSink(write_blobs, write_frames_as_jpeg, enc)
        else:
            wr, _written = _open_cv2_writer(output_video_path, out_fps, W, H)
            if wr is None:
                return False, 0, 0, 0             # reference :569-571

            def write_frames(_keys, frames):
                for f in frames:
                    wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))

            sink = _FrameSink(write_frames, B, H, W, dev)

        # ---- loop -----------------------------------------------------------------------------------------------
        if rd is not None:
            groups = [list(range(s, min(s + B, n_frames))) for s in range(0, n_frames, B)]
            dec = vjpeg.JpegDecoder(B, device=dev, max_file_bytes=max(4 << 20, 2 * max(len(rd.frame(i)) for i in range(min(n_frames, 64)))))
            source = _JpegSource(rd.frames, dec)
        else:
            def groups_from_capture():
                frames = []
                while True:
                    ok, bgr = cap.read()
                    if not ok:
                        break
                    if bgr.shape[:2] != (H, W):   # reference :580-581
                        bgr = cv2.resize(bgr, (W, H))
                    frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
                    if len(frames) == B:
                        yield frames
                        frames = []
                if frames:
                    yield frames

            groups = groups_from_capture()
            source = _FrameSource(lambda frames: frames, B, H, W, dev)
        n, n_faces, n_plates = _resident_loop(groups, source, sink, eng, (H, W), face_detector, plate_detector, mode,
                                              mosaic_level, ksize, sigma, failed)
        return n > 0, n, n_faces, n_plates
    finally:
        if avi_out is not None:
            avi_out.close()
        if wr is not None:
            wr.release()
        if cap is not None:
            cap.release()
        if rd is not None:
            rd.close()


def process_single_video(video_path, output_videos_dir, face_detector, plate_detector, temp_base_dir=None, cleanup=True,
                         batch_size=16, **kwargs):
    """reference :648-699: '<name>_processed.<ext>' in `output_videos_dir`; True on success, False (never raises) on
    failure.  `temp_base_dir` / `cleanup` are accepted for compatibility: there is no scratch directory to clean."""
    name, ext = os.path.splitext(os.path.basename(video_path))
    if not ext.lstrip("."):
        return False                              # reference :662-664: unknown format is left to the copy step
    out = os.path.join(output_videos_dir, f"{name}_processed.{ext.lstrip('.').lower()}")
    try:
        return bool(process_video_pipeline(video_path, out, face_detector, plate_detector, fps=60, batch_size=batch_size,
                                           **kwargs)[0])
    except Exception as e:                        # noqa: BLE001 (reference :687-689 logs and returns False)
        import logging
        logging.getLogger("VideoProcessor.process_single_video").error("processing %s failed: %s", video_path, e)
        return False


def convert_video_to_frames(video_path, output_dir, interval=1):
    """reference :279-476: video -> `output_dir/frame_%06d.jpg` (numbered from 1 like ffmpeg), returns the number of
    files written (0 on failure).  A Motion-JPEG AVI is demuxed — each frame's bytes become the file unchanged, no
    transcode; anything else is decoded by cv2.VideoCapture and written with cv2.imwrite (the reference runs an ffmpeg
    subprocess with -qscale:v 2).  `interval` = keep every interval-th frame (the reference's select expression uses
    mod(n, interval-1), :362, which keeps every frame for interval=2 — the documented intent is implemented here).
    Only needed by callers that want the files: process_video_pipeline does not go through a directory."""
    import glob
    os.makedirs(output_dir, exist_ok=True)
    for f in glob.glob(os.path.join(output_dir, "*.jpg")):     # reference :290-291
        os.remove(f)
    step = max(1, int(interval))
    saved = 0
    rd = _open_mjpeg_avi(video_path)
    if rd is not None:
        with rd:
            for i in range(0, len(rd), step):
                saved += 1
                _write_bytes(os.path.join(output_dir, "frame_%06d.jpg" % saved), rd.frame(i))
        return saved
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        return 0
    try:
        i = 0
        while True:
            ok, bgr = cap.read()
            if not ok:
                break
            if i % step == 0:
                saved += 1
                cv2.imwrite(os.path.join(output_dir, "frame_%06d.jpg" % saved), bgr)
            i += 1
    finally:
        cap.release()
    return saved


def create_video(frame_dir, output_path, fps=60):
    """reference :479-595: the sorted `*.jpg` files of `frame_dir` -> one video; True on success.  `.avi`: when every file
    is a JPEG of the first file's size the files are muxed as they are (Motion-JPEG, no re-encode: the video's frames
    are the files, bit for bit); otherwise, and for every other extension, cv2.VideoWriter as in the reference."""
    import glob
    import cv2
    from vdesens_b200 import avi
    from vdesens_b200 import jpeg as vjpeg
    files = sorted(glob.glob(os.path.join(frame_dir, "*.jpg")))
    if not files:
        return False
    if os.path.splitext(output_path)[1].lower() == ".avi":
        try:
            first = vjpeg.read_header(_read_bytes(files[0]))
            os.makedirs(os.path.dirname(os.path.abspath(output_path)), exist_ok=True)
            with avi.AviWriter(output_path, first.width, first.height, fps) as out:
                for p in files:
                    blob = _read_bytes(p)
                    h = vjpeg.read_header(blob)
                    if (h.width, h.height) != (first.width, first.height):
                        raise RuntimeError("frame size changes")
                    out.write(blob)
            return True
        except RuntimeError:
            if os.path.exists(output_path):
                os.remove(output_path)
    sample = cv2.imread(files[0])
    if sample is None:
        return False
    height, width = sample.shape[:2]
    wr, _written = _open_cv2_writer(output_path, fps, width, height)
    if wr is None:
        return False
    try:
        for p in files:
            frame = cv2.imread(p)
            if frame is None:
                continue                          # reference :576-578
            if frame.shape[:2] != (height, width):
                frame = cv2.resize(frame, (width, height))
            wr.write(frame)
    finally:
        wr.release()
    return True
