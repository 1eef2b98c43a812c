"""JPEG codec on the device (csrc/jpeg_dec.cu, csrc/jpeg_enc.cu, csrc/jpeg_host.cu behind the C-ABI).

  JpegDecoder   file bytes -> resident RGB frames, pixels identical to cv2.imread + BGR->RGB (reference
                combine_detect.py:167-173): host header parse, parallel self-synchronising Huffman decode, IDCT, fancy
                up-sampling and colour conversion on the GPU; files the device path does not take fall back to cv2
  JpegEncoder   resident RGB frames -> file bytes identical to cv2.imwrite's (reference :176-180)
  read_header / decode_coefficients / reconstruct_from_coefficients   the host entropy decoder and the test hooks
Used by combine_detect.batch_process_images(io="device") and by the video stage (process_video_pipeline, row N3), where
the "files" are the chunks of a Motion-JPEG AVI (vdesens_b200/avi.py)."""
import ctypes as C

import numpy as np

from . import _lib


def read_header(data: bytes) -> _lib.VdJpegInfo:
    info = _lib.VdJpegInfo()
    _lib.check(_lib.lib().vd_jpeg_read_header(data, len(data), C.byref(info)), "vd_jpeg_read_header")
    return info


def decode_coefficients(data: bytes):
    """-> (info, [int16 (blocks_y, blocks_x, 64) per component], [uint16 (64,) quantisation table per component])"""
    info = read_header(data)
    coefs = [np.empty((info.blocks_y[c], info.blocks_x[c], 64), np.int16) for c in range(info.ncomp)]
    ptrs = [c.ctypes.data_as(C.c_void_p) for c in coefs] + [None] * (3 - info.ncomp)
    _lib.check(_lib.lib().vd_jpeg_decode_coefficients(data, len(data), *ptrs), "vd_jpeg_decode_coefficients")
    quant = [np.ctypeslib.as_array(info.quant[c]).copy() for c in range(info.ncomp)]
    return info, coefs, quant


class JpegEncoder:
    """Device-side `cv2.imwrite(path, bgr)` for frames that are already in HBM (reference combine_detect.py:176-180):
    `encode(frames)` returns, per frame, the bytes of the JPEG file cv2 would have written from the same RGB pixels
    (quality 95, 4:2:0 — cv2's defaults), byte for byte.  One instance per (device, batch, H, W); the compressed bytes
    (~0.9 MB per 1080p frame instead of 6.2 MB of pixels) are what crosses PCIe.  CUDA only: there is no CPU fallback."""

    def __init__(self, batch, height, width, device="cuda:0", quality=95, max_bytes_per_frame=None, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("JpegEncoder needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.B, self.H, self.W, self.quality = int(batch), int(height), int(width), int(quality)
        if not 1 <= self.quality <= 100:
            raise ValueError("quality must be in [1, 100]")
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self.stream = stream or torch.cuda.Stream(device=self.device)
        # per-frame output slot: 1 byte per pixel is > 2x what a photographic 1080p frame needs at quality 95
        self.pitch = int(max_bytes_per_frame or (self.H * self.W + 4096))
        with torch.cuda.device(self.device):
            nbytes = int(self.lib.vd_jpeg_encode_workspace_bytes(self.B, self.H, self.W))
            self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
            self.out = torch.empty((self.B, self.pitch), dtype=torch.uint8, device=self.device)
            self.sizes = torch.zeros((self.B,), dtype=torch.int32, device=self.device)
            self.status = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.out_host = torch.empty((self.B, self.pitch), dtype=torch.uint8).pin_memory()
        self.sizes_host = torch.zeros((self.B,), dtype=torch.int32).pin_memory()
        self.status_host = torch.zeros((1,), dtype=torch.int32).pin_memory()
        self._ws_bytes = nbytes

    def enqueue(self, frame_desc, wait_stream=None):
        """frame_desc: (B,6) int32 device tensor of vd_frame_t records (ops.frame_descriptors / engine.frame_desc) of
        B frames of H x W.  Enqueues the five encoder kernels on the encoder's stream (after `wait_stream`'s work)."""
        torch = self.torch
        with torch.cuda.device(self.device):
            if wait_stream is not None:
                self.stream.wait_stream(wait_stream)
            _lib.check(self.lib.vd_jpeg_encode_batch(C.c_void_p(frame_desc.data_ptr()), self.B, self.H, self.W, self.quality,
                                                     C.c_void_p(self.out.data_ptr()), self.pitch,
                                                     C.c_void_p(self.sizes.data_ptr()), C.c_void_p(self.status.data_ptr()),
                                                     C.c_void_p(self.ws.data_ptr()), self._ws_bytes,
                                                     C.c_void_p(self.stream.cuda_stream)), "vd_jpeg_encode_batch", self.lib)

    def fetch(self, n=None):
        """-> list of `bytes`, one JPEG file per frame (the first n).  Two-step D2H: the sizes, then only the bytes used."""
        torch = self.torch
        n = self.B if n is None else int(n)
        with torch.cuda.stream(self.stream):
            self.sizes_host.copy_(self.sizes, non_blocking=True)
            self.status_host.copy_(self.status, non_blocking=True)
        self.stream.synchronize()
        if int(self.status_host[0]) & _lib.VD_STAT_JPEG_OVERFLOW or int(self.sizes_host[:n].min()) < 0:
            with torch.cuda.stream(self.stream):
                self.status.zero_()
            raise RuntimeError("JPEG encoder: a frame did not fit max_bytes_per_frame / the stream workspace")
        sizes = [int(v) for v in self.sizes_host[:n]]
        with torch.cuda.stream(self.stream):
            for i, s in enumerate(sizes):
                self.out_host[i, :s].copy_(self.out[i, :s], non_blocking=True)
        self.stream.synchronize()
        return [self.out_host[i, :s].numpy().tobytes() for i, s in enumerate(sizes)]

    def encode(self, frames):
        """frames: list / tensor of B device uint8 (H,W,3) RGB frames -> list of JPEG files as bytes"""
        from . import ops
        fl = list(frames) if not isinstance(frames, self.torch.Tensor) else [frames[i] for i in range(frames.shape[0])]
        n = len(fl)
        assert 0 < n <= self.B and all(tuple(f.shape) == (self.H, self.W, 3) for f in fl)
        fl = fl + [fl[-1]] * (self.B - n)
        desc = ops.frame_descriptors(fl)
        self.enqueue(desc, wait_stream=self.torch.cuda.current_stream(self.device))
        return self.fetch(n)


def prepare_device(data: bytes) -> _lib.VdJpegDevHdr:
    """host parse of one file into the record the device decoder reads; RuntimeError for files the device path does not
    take (progressive, restart markers, unusual sampling): decode those on the CPU"""
    hdr = _lib.VdJpegDevHdr()
    _lib.check(_lib.lib().vd_jpeg_prepare_device(data, len(data), C.byref(hdr)), "vd_jpeg_prepare_device")
    return hdr


_GEOM_FIELDS = ("width", "height", "ncomp", "mcux", "mcuy", "blocks_per_mcu")


def _geometry_key(h):
    return tuple(getattr(h, f) for f in _GEOM_FIELDS) + tuple(h.h) + tuple(h.v)


class JpegDecoder:
    """Device-side `cv2.cvtColor(cv2.imread(path), COLOR_BGR2RGB)` (reference combine_detect.py:167-173) for a batch of
    JPEG files of one geometry: the file bytes are what crosses PCIe, and Huffman decoding, IDCT, up-sampling and colour
    conversion run on the GPU with pixels identical to cv2's.  `decode(files)` returns a device uint8 (n,H,W,3) tensor.
    Files the device path does not take — or (never seen so far) whose parallel entropy decode does not converge — are
    decoded by `fallback` (default cv2.imdecode) and uploaded.  CUDA only for the device path itself."""

    def __init__(self, batch, device="cuda:0", max_file_bytes=4 << 20, sub_bytes=0, fallback=None, stream=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("JpegDecoder needs a CUDA device")
        self.torch = torch
        self.B, self.max_file = int(batch), int(max_file_bytes)
        self.device = torch.device(device)
        self.lib = _lib.lib()
        self.sub_bytes = int(sub_bytes)
        self.stream = stream or torch.cuda.Stream(device=self.device)
        self.fallback = fallback
        self.hdr_bytes = C.sizeof(_lib.VdJpegDevHdr)
        self.stage_bytes = self.B * self.hdr_bytes + self.B * (self.max_file + 16)
        self.stage_host = torch.empty((self.stage_bytes,), dtype=torch.uint8).pin_memory()
        self.stage_dev = torch.empty((self.stage_bytes,), dtype=torch.uint8, device=self.device)
        self.result = torch.zeros((self.B,), dtype=torch.int32, device=self.device)
        self.result_host = torch.zeros((self.B,), dtype=torch.int32).pin_memory()
        self._geom, self._ws, self._ws_bytes = None, None, 0
        self.stats = {"files": 0, "device": 0, "fallback": 0}

    def _fallback(self, data):
        if self.fallback is not None:
            return self.fallback(data)
        import cv2
        import numpy as np
        img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        if img is None:
            raise ValueError("cannot decode image")
        return cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

    def sync_rounds(self):
        return _sync_rounds(self)

    def _workspace(self, hdr):
        key = _geometry_key(hdr)
        if self._geom != key:
            n = int(self.lib.vd_jpeg_decode_workspace_bytes(self.B, C.byref(hdr), self.max_file))
            if n == 0:
                raise RuntimeError("vd_jpeg_decode_workspace_bytes: bad geometry")
            self._ws = self.torch.empty((n,), dtype=torch_uint8(self.torch), device=self.device)
            self._ws_bytes, self._geom = n, key
        return self._ws

    def decode(self, files, out=None):
        """files: list of <= B `bytes` objects (JPEG files of one size).  out: optional device uint8 (B,H,W,3) tensor to
        decode into.  Returns (frames (n,H,W,3) device tensor, list of indices that took the CPU fallback)."""
        import numpy as np
        from . import ops
        torch = self.torch
        n = len(files)
        assert 0 < n <= self.B
        hdrs, on_cpu = [], []
        for i, data in enumerate(files):
            try:
                if len(data) > self.max_file:
                    raise RuntimeError("file larger than max_file_bytes")
                hdrs.append(prepare_device(data))
            except RuntimeError:
                hdrs.append(None)
                on_cpu.append(i)
        good = [i for i in range(n) if hdrs[i] is not None]
        if good:
            key = _geometry_key(hdrs[good[0]])
            for i in good[1:]:
                if _geometry_key(hdrs[i]) != key:      # another size / sampling than the batch's: CPU for that one
                    hdrs[i] = None
                    on_cpu.append(i)
            good = [i for i in good if hdrs[i] is not None]
        if not good:
            imgs = [self._fallback(d) for d in files]
            self.stats["files"] += n
            self.stats["fallback"] += n
            t = torch.from_numpy(np.stack(imgs)).to(self.device)
            if out is not None:
                out[:n].copy_(t)
                t = out[:n]
            return t, sorted(on_cpu)
        g0 = hdrs[good[0]]
        H, W = g0.height, g0.width
        if out is None:
            out = torch.empty((self.B, H, W, 3), dtype=torch.uint8, device=self.device)
        assert tuple(out.shape[1:]) == (H, W, 3) and out.shape[0] >= self.B
        # staging: [B headers][file bytes ...]; slots of files that go to the CPU repeat the first good file
        stage = self.stage_host.numpy()
        pos = self.B * self.hdr_bytes
        max_scan = 0
        for slot in range(self.B):
            i = slot if (slot < n and hdrs[slot] is not None) else good[0]
            data, h = files[i], hdrs[i]
            hh = _lib.VdJpegDevHdr.from_buffer_copy(bytes(h))
            stage[pos:pos + len(data)] = np.frombuffer(data, np.uint8)
            hh.scan_offset = pos + h.scan_offset
            stage[slot * self.hdr_bytes:(slot + 1) * self.hdr_bytes] = np.frombuffer(bytes(hh), np.uint8)
            max_scan = max(max_scan, h.scan_bytes)
            pos += (len(data) + 15) & ~15
        ws = self._workspace(g0)
        desc = ops.frame_descriptors([out[i] for i in range(self.B)])
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            self.stage_dev[:pos].copy_(self.stage_host[:pos], non_blocking=True)
            _lib.check(self.lib.vd_jpeg_decode_batch(C.c_void_p(self.stage_dev.data_ptr()), C.byref(g0),
                                                     C.c_void_p(self.stage_dev.data_ptr()), self.B, self.max_file,
                                                     C.c_void_p(desc.data_ptr()), C.c_void_p(self.result.data_ptr()),
                                                     self.sub_bytes, C.c_void_p(ws.data_ptr()), self._ws_bytes,
                                                     C.c_void_p(self.stream.cuda_stream)), "vd_jpeg_decode_batch", self.lib)
            self.result_host.copy_(self.result, non_blocking=True)
        self.stream.synchronize()
        for i in good:
            if int(self.result_host[i]) & 0xFF:
                on_cpu.append(i)
        for i in sorted(set(on_cpu)):
            out[i].copy_(torch.from_numpy(self._fallback(files[i])).to(self.device))
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        self.stats["files"] += n
        self.stats["fallback"] += len(set(on_cpu))
        self.stats["device"] += n - len(set(on_cpu))
        return out[:n], sorted(set(on_cpu))


def _sync_rounds(self):
    """diagnostic: per file of the last batch, the first synchronisation round in which no exit state changed"""
    return [(int(v) >> 8) & 0xFF for v in self.result_host]


def torch_uint8(torch):
    return torch.uint8


def reconstruct_from_coefficients(data: bytes, device="cuda:0"):
    """test hook: entropy decoding on the HOST (vd_jpeg_decode_coefficients), everything after it on the device
    (vd_jpeg_reconstruct_batch) -> uint8 (H,W,3) RGB numpy array"""
    import numpy as np
    import torch
    from . import ops
    lib = _lib.lib()
    hdr = prepare_device(data)
    info, coefs, _ = decode_coefficients(data)
    dev = torch.device(device)
    cd = [torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in coefs] + [None] * (3 - info.ncomp)
    hd = torch.from_numpy(np.frombuffer(bytes(hdr), np.uint8).copy()).to(dev)
    out = torch.empty((1, hdr.height, hdr.width, 3), dtype=torch.uint8, device=dev)
    n = int(lib.vd_jpeg_decode_workspace_bytes(1, C.byref(hdr), 4096))
    ws = torch.empty((n,), dtype=torch.uint8, device=dev)
    desc = ops.frame_descriptors([out[0]])
    with torch.cuda.device(dev):
        _lib.check(lib.vd_jpeg_reconstruct_batch(C.c_void_p(hd.data_ptr()), C.byref(hdr), 1, C.c_void_p(cd[0].data_ptr()),
                                                 C.c_void_p(cd[1].data_ptr()) if cd[1] is not None else None,
                                                 C.c_void_p(cd[2].data_ptr()) if cd[2] is not None else None,
                                                 C.c_void_p(desc.data_ptr()), C.c_void_p(ws.data_ptr()), n,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "vd_jpeg_reconstruct_batch", lib)
        torch.cuda.synchronize()
    return out[0].cpu().numpy()
