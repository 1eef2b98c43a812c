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
