"""Build container only: the restatements in oracle/ against the UNMODIFIED reference imported from
/root/reference (skipped on the GPU box, where the tree does not exist)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import cv_oracle as co
from oracle import retinaface_oracle as ro
from oracle.ref_shims import load_reference
from vdesens_b200 import synth

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_anchors_bit_exact(ref):
    face, _ = ref
    cfg = face.cfg_re50
    for h, w in [(96, 128), (135, 241), (640, 640)]:
        a = face.Anchors(cfg, image_size=[h, w]).get_anchors()
        assert torch.equal(a, ro.anchors(h, w))


def test_detector_end_to_end_same_boxes(ref):
    face, _ = ref
    sd = synth.synthetic_state_dict(0, face_rate="r01")
    pth = os.path.join(tempfile.mkdtemp(), "w.pth")
    synth.save_state_dict_pth(sd, pth)
    frames = [synth.synthetic_frame(150, 200, 11), synth.synthetic_frame(200, 150, 12)]
    det = face.Retinaface(model_path=pth, backbone="resnet50", input_shape=[160, 160, 3], confidence=0.5,
                          nms_iou=0.4, letterbox_image=True, cuda=False)
    want = det.detect_images(frames)
    got = ro.detect_images(ro.to_torch_sd(sd), frames, (160, 160), 0.5, 0.4)
    for (wi, wb), (gi, gb) in zip(want, got):
        assert np.array_equal(wi, gi)
        assert len(wb) == len(gb)
        if wb:
            np.testing.assert_allclose(np.array(gb), np.array(wb), atol=0.02)


def test_mosaic_random_boxes(ref):
    _, combine = ref
    rng = np.random.default_rng(0)
    for _ in range(200):
        h, w = (int(v) for v in rng.integers(20, 200, 2))
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        x1, y1 = int(rng.integers(-30, w)), int(rng.integers(-30, h))
        x2, y2 = x1 + int(rng.integers(-5, 150)), y1 + int(rng.integers(-5, 150))
        lvl = int(rng.integers(2, 17))
        assert np.array_equal(combine.mosaic_rectangle_region_single(img, x1, y1, x2, y2, lvl),
                              co.mosaic_box(img, x1, y1, x2, y2, lvl))


def test_letterbox(ref):
    face, _ = ref
    for ih, iw, W, H in [(480, 640, 640, 640), (1080, 1920, 640, 640), (270, 480, 480, 270), (333, 517, 320, 256),
                         (480, 640, 1280, 1280), (120, 160, 320, 256), (200, 100, 256, 300)]:      # the last three up-scale
        img = synth.synthetic_frame(ih, iw, ih + iw)
        assert np.array_equal(face.letterbox_image(img, [W, H]), co.letterbox(img, W, H).astype(np.float64))
