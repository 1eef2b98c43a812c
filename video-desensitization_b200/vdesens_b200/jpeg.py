"""Host half of the device-resident JPEG ingest: libvdesens' entropy decoder behind numpy arrays.
`decode_coefficients(data)` returns the quantised DCT coefficients of a baseline JPEG as int16 planes; the device half
(IDCT, up-sampling, colour conversion — bit-exact with the libjpeg-turbo behind cv2.imread, reference
combine_detect.py:167-180) is not written yet, so nothing in the frame loop uses this module so far."""
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
