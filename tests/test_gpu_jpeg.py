"""GPU: the device JPEG encoder (csrc/jpeg_enc.cu, SURVEY §8f row N1, reference combine_detect.py:176-180 cv2.imwrite)
through the C-ABI, byte for byte against cv2.imencode and against oracle/jpeg_enc_oracle.py (itself pinned on cv2)."""
import numpy as np
import pytest
import torch
import cv2

from oracle import jpeg_enc_oracle as je
from vdesens_b200 import synth
from vdesens_b200.jpeg import JpegEncoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cv2(rgb, quality=95):
    ok, buf = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, quality])
    assert ok
    return buf.tobytes()


def _image(h, w, kind, seed):
    if kind == "smooth":
        return synth.synthetic_frame(max(h, 8), max(w, 8), seed)[:h, :w].copy()
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _first_diff(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), n)


@pytest.mark.parametrize("hw", [(16, 16), (8, 8), (17, 23), (48, 64), (45, 77), (135, 241), (1, 1), (9, 33), (24, 40), (40, 24),
                                (100, 100), (270, 480), (2, 300), (300, 2), (136, 130)])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_encoder_is_byte_identical_to_cv2(hw, kind):
    h, w = hw
    imgs = [_image(h, w, kind, h * 7 + w + i) for i in range(3)]
    enc = JpegEncoder(3, h, w, device=DEV, max_bytes_per_frame=h * w * 4 + 4096)
    got = enc.encode([torch.from_numpy(im).to(DEV) for im in imgs])
    for im, g in zip(imgs, got):
        want = _cv2(im)
        assert g == want, f"{hw} {kind}: first difference at byte {_first_diff(g, want)} of {len(g)}/{len(want)}"
        assert g == je.encode_rgb(im)


@pytest.mark.parametrize("quality", [1, 30, 50, 75, 100])
def test_encoder_other_qualities(quality):
    imgs = [_image(45, 77, k, quality) for k in ("smooth", "noise")]
    enc = JpegEncoder(2, 45, 77, device=DEV, quality=quality, max_bytes_per_frame=45 * 77 * 6 + 4096)
    for im, g in zip(imgs, enc.encode([torch.from_numpy(im).to(DEV) for im in imgs])):
        want = _cv2(im, quality)
        assert g == want, f"q={quality}: first difference at byte {_first_diff(g, want)} of {len(g)}/{len(want)}"


def test_encoder_full_hd_batch_and_views():
    """the frame loop's size: a 1080p batch (68th MCU row = dummy luma blocks), frames given as views of one batch tensor;
    decoding the device's file gives exactly the pixels decoding cv2's file gives (same bytes) and the file sizes match"""
    B, H, W = 4, 1080, 1920
    frames = np.stack([synth.synthetic_frame(H, W, 2000 + i) for i in range(B)])
    dev = torch.from_numpy(frames).to(DEV)
    enc = JpegEncoder(B, H, W, device=DEV)
    got = enc.encode(dev)
    for i in range(B):
        want = _cv2(frames[i])
        assert got[i] == want, f"frame {i}: first difference at byte {_first_diff(got[i], want)} of {len(got[i])}/{len(want)}"
    assert np.array_equal(cv2.imdecode(np.frombuffer(got[0], np.uint8), cv2.IMREAD_COLOR),
                          cv2.imdecode(np.frombuffer(_cv2(frames[0]), np.uint8), cv2.IMREAD_COLOR))
    # a second call on the same encoder (workspace re-zeroed) and a short batch
    again = enc.encode(dev[:2])
    assert again == got[:2]


def test_encoder_reports_a_frame_that_does_not_fit():
    im = _image(64, 64, "noise", 5)
    enc = JpegEncoder(1, 64, 64, device=DEV, max_bytes_per_frame=1024)
    with pytest.raises(RuntimeError, match="did not fit"):
        enc.encode([torch.from_numpy(im).to(DEV)])
    enc2 = JpegEncoder(1, 64, 64, device=DEV, max_bytes_per_frame=64 * 64 * 4)
    assert enc2.encode([torch.from_numpy(im).to(DEV)])[0] == _cv2(im)
