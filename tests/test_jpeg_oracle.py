"""Pins oracle/jpeg_oracle.py — the checker for a future device-side JPEG decode — against the
decoder the reference's frame loop actually uses: `cv2.imread` / `cv2.imdecode` (reference combine_detect.py:167-180),
i.e. OpenCV's bundled libjpeg-turbo with default settings.  Bit-exact or nothing: a frame that differs by one LSB after
decoding would move detections and blurred bytes downstream."""
import numpy as np
import pytest
import cv2

from oracle import jpeg_oracle as jo
from vdesens_b200 import synth

S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
SF = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}


def _roundtrip(rgb, params):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR) if rgb.ndim == 3 else rgb, params)
    assert ok
    want = cv2.cvtColor(cv2.imdecode(buf, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)     # what load_image_rgb hands to the path
    return buf.tobytes(), want


@pytest.mark.parametrize("sampling", ["420", "422", "444"])
@pytest.mark.parametrize("hw,quality", [((48, 64), 95), ((45, 77), 95), ((135, 241), 60), ((67, 93), 100), ((33, 47), 30)])
def test_decode_bit_exact(sampling, hw, quality):
    rng = np.random.default_rng(hw[0] * 1000 + quality)
    for img in (synth.synthetic_frame(hw[0], hw[1], quality), rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)):
        data, want = _roundtrip(img, [S, SF[sampling], cv2.IMWRITE_JPEG_QUALITY, quality])
        j = jo.parse(data)
        assert [(c.h, c.v) for c in j["comps"]][0] == {"420": (2, 2), "422": (2, 1), "444": (1, 1)}[sampling]
        assert np.array_equal(jo.decode_rgb(data), want)


def test_cv2_default_is_420_like_the_reference_writes_it():
    """cv2.imwrite(path, img) with no parameters (reference combine_detect.py:205-206): quality 95, 4:2:0"""
    data, want = _roundtrip(synth.synthetic_frame(96, 128, 3), [])
    j = jo.parse(data)
    assert [(c.h, c.v) for c in j["comps"]] == [(2, 2), (1, 1), (1, 1)] and j["restart"] == 0
    assert np.array_equal(jo.decode_rgb(data), want)


@pytest.mark.parametrize("hw", [(1, 1), (2, 2), (8, 3), (3, 4), (5, 5), (16, 4), (9, 6), (31, 2), (7, 33), (17, 5)])
def test_tiny_and_narrow_images(hw):
    """down-sampled widths <= 2 take libjpeg's replication path instead of the triangle filter; MCU padding is cropped"""
    rng = np.random.default_rng(hw[0] * 37 + hw[1])
    for sampling in SF.values():
        data, want = _roundtrip(rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8), [S, sampling])
        assert np.array_equal(jo.decode_rgb(data), want), (hw, sampling)


def test_greyscale_and_restart_intervals():
    g = synth.synthetic_frame(64, 80, 6)[..., 0].copy()
    data, want = _roundtrip(g, [])
    assert len(jo.parse(data)["comps"]) == 1 and np.array_equal(jo.decode_rgb(data), want)
    for rst in (1, 4, 7):
        data, want = _roundtrip(synth.synthetic_frame(100, 150, 7), [cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
        assert jo.parse(data)["restart"] > 0 and np.array_equal(jo.decode_rgb(data), want)


def test_full_hd_frame_as_the_frame_loop_reads_it():
    """one 1080p frame at cv2's default settings (the bench workload's frame size)"""
    data, want = _roundtrip(synth.synthetic_frame(1080, 1920, 1000), [])
    assert np.array_equal(jo.decode_rgb(data), want)


def test_stages_in_isolation():
    """the IDCT of a DC-only block is flat, the colour tables reproduce cv2.cvtColor's inverse only where JPEG says so
    (full-range BT.601), and the up-samplers keep constant planes constant"""
    coef = np.zeros((1, 64), np.int32)
    coef[0, 0] = 10
    q = np.full(64, 16, np.int32)
    assert np.all(jo.idct_islow(coef, q) == 128 + (10 * 16 + 4) // 8)
    flat = np.full((6, 9), 77, np.uint8)
    assert np.all(jo.h2v2_fancy(flat, 9, 6) == 77) and np.all(jo.h2v1_fancy(flat, 9) == 77)
    y = np.arange(256, dtype=np.uint8).reshape(16, 16)
    n = np.full((16, 16), 128, np.uint8)
    assert np.array_equal(jo.ycc_to_rgb(y, n, n), np.stack([y, y, y], -1))            # neutral chroma: grey
