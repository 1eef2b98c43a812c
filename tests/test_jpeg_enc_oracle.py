"""CPU: oracle/jpeg_enc_oracle.py (libjpeg's compression arithmetic restated) against `cv2.imencode('.jpg')` of this
image's OpenCV build, byte for byte — the pin of the checker the device JPEG encoder (reference
combine_detect.py:176-180 `cv2.imwrite`) is tested against."""
import numpy as np
import pytest
import cv2

from oracle import jpeg_enc_oracle as je
from oracle import jpeg_oracle as jo
from vdesens_b200 import synth

S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
SF = {(2, 2): cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, (2, 1): cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      (1, 1): cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444}


def _cv2(rgb, quality=None, sampling=None):
    params = []
    if quality is not None:
        params += [cv2.IMWRITE_JPEG_QUALITY, quality]
    if sampling is not None:
        params += [S, SF[sampling]]
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), params)
    assert ok
    return buf.tobytes()


def _image(h, w, kind, seed):
    if kind == "smooth":
        return synth.synthetic_frame(max(h, 8), max(w, 8), seed)[:h, :w].copy()
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("hw", [(16, 16), (8, 8), (17, 23), (48, 64), (45, 77), (135, 241), (1, 1), (9, 33), (24, 40), (40, 24),
                                (100, 100), (270, 480), (2, 300), (300, 2)])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_default_parameters_match_cv2_byte_for_byte(hw, kind):
    """cv2's defaults (what cv2.imwrite(path, img) uses): quality 95, 4:2:0.  Sizes cover: whole MCUs, dummy blocks at the
    right / bottom edge, odd and EVEN heights that are not multiples of 16 (the bottom rows are copies of the last
    down-sampled row, not of the last pixel row), one-pixel images."""
    rgb = _image(hw[0], hw[1], kind, hw[0] * 7 + hw[1])
    assert je.encode_rgb(rgb) == _cv2(rgb)


@pytest.mark.parametrize("quality,sampling", [(30, (2, 2)), (50, (2, 2)), (75, (2, 2)), (100, (2, 2)), (1, (2, 2)),
                                              (95, (2, 1)), (95, (1, 1)), (60, (1, 1))])
def test_other_qualities_and_samplings(quality, sampling):
    for kind in ("smooth", "noise"):
        rgb = _image(45, 77, kind, quality)
        assert je.encode_rgb(rgb, quality, sampling) == _cv2(rgb, quality, sampling)


def test_full_hd_coefficients_and_header():
    """one 1080p frame (the size the frame loop writes): header bytes identical and every quantised coefficient — dummy
    blocks of the 68th MCU row included — equal to what cv2's file decodes to; the Python entropy coder is too slow to
    repeat 48 960 blocks here, the byte-level check at this size is the GPU test of the device encoder"""
    rgb = synth.synthetic_frame(1080, 1920, 1000)
    want = _cv2(rgb)
    qt, coefs = je.image_coefficients(rgb)
    hdr = je.header(1920, 1080, (2, 2), qt)
    assert want[:len(hdr)] == hdr
    j = jo.parse(want)
    ref = jo.decode_coefficients(j)
    for c in range(3):
        assert ref[c].shape == coefs[c].shape and np.array_equal(ref[c], coefs[c]), c
