"""CPU: the oracle restatements against the committed outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py) and against the cv2 / torchvision builds
in this image."""
import os

import numpy as np
import pytest
import torch

from oracle import cv_oracle as co
from oracle import retinaface_oracle as ro
from vdesens_b200 import synth


@pytest.fixture(scope="module")
def det_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_detect_golden.npz"))


@pytest.fixture(scope="module")
def sd():
    return ro.to_torch_sd(synth.synthetic_state_dict(0, face_rate="r01"))


def _frames(g, name):
    return [synth.synthetic_frame(int(h), int(w), int(s)) for h, w, s in g[name + ".frames"]]


@pytest.mark.parametrize("case", ["native", "letterbox", "odd"])
def test_preprocess_and_heads_match_reference(det_gold, sd, case):
    frames = _frames(det_gold, case)
    ish = det_gold[case + ".input_shape"]
    x, _ = ro.preprocess(frames, (int(ish[0]), int(ish[1])))
    assert np.array_equal(x.numpy(), det_gold[case + ".x"])          # ingest is exact integers
    loc, conf, landms = ro.forward(sd, x)
    # same torch build + same ops -> equal up to thread-partition rounding of the conv
    np.testing.assert_allclose(loc.numpy(), det_gold[case + ".loc"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(conf.numpy(), det_gold[case + ".conf"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(landms.numpy(), det_gold[case + ".landms"], atol=2e-4, rtol=0)


@pytest.mark.parametrize("case", ["native", "letterbox", "odd"])
def test_postprocess_on_reference_heads_is_exact(det_gold, case):
    """stage-isolated: identical head tensors in -> identical rows out (decode, threshold, NMS,
    un-letterbox), bit for bit."""
    frames = det_gold[case + ".frames"]
    ish = det_gold[case + ".input_shape"]
    shapes = torch.tensor([[float(h), float(w)] for h, w, _ in frames])
    rows = ro.postprocess(torch.from_numpy(det_gold[case + ".loc"]), torch.from_numpy(det_gold[case + ".conf"]),
                          torch.from_numpy(det_gold[case + ".landms"]), shapes, (int(ish[0]), int(ish[1])),
                          0.5, 0.4)
    for i, r in enumerate(rows):
        assert np.array_equal(r.numpy().reshape(-1, 15), det_gold[f"{case}.rows{i}"])


def test_detect_images_c1(det_gold, sd):
    frames = _frames(det_gold, "c1")
    res = ro.detect_images(sd, frames, (640, 640), 0.5, 0.4)
    for i, (img, boxes) in enumerate(res):
        ref = det_gold[f"c1.boxes{i}"]
        got = np.array(boxes).reshape(-1, 4)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, atol=0.05)             # pixels; conv thread-partition noise


def test_anchor_counts():
    from vdesens_b200 import arch
    assert ro.anchors(640, 640).shape[0] == 16800 == arch.num_anchors(640, 640)
    assert arch.num_anchors(480, 640) == 12600
    assert arch.num_anchors(1080, 1920) == 85200
    assert arch.num_anchors(2160, 3840) == 340320


def test_nms_matches_torchvision_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_golden.npz"))
    for ci in range(5):
        keep = ro.nms_keep(g[f"c{ci}.boxes"], g[f"c{ci}.scores"], 0.4)
        assert np.array_equal(keep, g[f"c{ci}.keep"])


def test_cv_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cv_golden.npz"))
    for gi in range(int(g["g.count"])):
        k, sg = g[f"g{gi}.par"]
        assert np.array_equal(co.gaussian_blur_u8(g[f"g{gi}.roi"], int(k), float(sg)), g[f"g{gi}.out"]), gi
    boxes = [tuple(int(v) for v in b) for b in g["mosaic.boxes"]]
    assert np.array_equal(co.composite(g["mosaic.img"], boxes, "mosaic", 8), g["mosaic.out"])
    assert np.array_equal(co.composite(g["mosaic.img"], boxes, "gauss", ksize=15), g["gauss.out"])
    for li in range(int(g["lb.count"])):
        ih, iw, W, H, seed = (int(v) for v in g[f"lb{li}.par"])
        assert np.array_equal(co.letterbox(synth.synthetic_frame(ih, iw, seed), W, H), g[f"lb{li}.out"])


def test_cv_live_against_this_image():
    """the arithmetic's true home is OpenCV: check the restatement against the cv2 in this image"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for k, sg, h, w in [(31, 0, 77, 53), (99, 0, 40, 300), (5, 0, 2, 2), (21, 4.0, 128, 128), (31, 0, 10, 200)]:
        roi = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(co.gaussian_blur_u8(roi, k, sg), cv2.GaussianBlur(roi, (k, k), sg))
    for dst, src in [(7, 60), (60, 7), (13, 13), (1, 5), (5, 1)]:
        a = rng.integers(0, 256, size=(src, src, 3), dtype=np.uint8)
        ref = cv2.resize(a, (dst, dst), interpolation=cv2.INTER_NEAREST)
        m = co.nearest_map(dst, src)
        assert np.array_equal(a[m][:, m], ref)
    # INTER_LINEAR (the letterbox's resize, utils.py:12): down-scaling, the exact 2x reduction OpenCV hands to its fast
    # area path, and UP-scaling, where the first / last output rows blend the border row with itself (VERDICT r1 row a3:
    # "up-scale still unpinned" — the y axis keeps its weights and clips the row indices, the x axis clamps the weights)
    for (h, w), (dh, dw) in [((40, 60), (20, 30)), ((40, 60), (80, 120)), ((40, 60), (57, 83)), ((33, 47), (66, 94)),
                             ((45, 77), (90, 100)), ((30, 30), (31, 31)), ((64, 64), (16, 16)), ((48, 64), (128, 171)),
                             ((7, 9), (50, 64)), ((1, 1), (5, 5)), ((2, 3), (7, 9)), ((37, 91), (36, 200)),
                             ((91, 37), (200, 36)), ((480, 640), (960, 1280))]:
        a = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(co.resize_linear_u8(a, dw, dh), cv2.resize(a, (dw, dh), interpolation=cv2.INTER_LINEAR)), \
            ((h, w), (dh, dw))


def test_oracle_on_default_init_weights_matches_the_executed_reference(golden_dir):
    """The un-conditioned weights (init="default": every BN gamma = 1, beta = 0 — SURVEY §8d to the letter; VERDICT r1
    item 1): the oracle reproduces the heads and the kept rows of the UNMODIFIED reference executed on them
    (tests/golden/ref_detect_default_golden.npz, written by `oracle/make_golden.py --default-only`)."""
    g = np.load(os.path.join(golden_dir, "ref_detect_default_golden.npz"))
    sd = ro.to_torch_sd(synth.synthetic_state_dict(0, face_rate="r01", init="default"))
    assert float(sd["body.layer1.0.bn3.weight"].min()) == 1.0 and float(sd["body.layer4.2.bn3.bias"].abs().max()) == 0.0
    frames = _frames(g, "native")
    x, _ = ro.preprocess(frames, (96, 128))
    loc, conf, landms = ro.forward(sd, x)
    # this network amplifies even the thread-partition rounding of the CPU convolution ~100x: 2e-3 instead of 2e-4
    np.testing.assert_allclose(loc.numpy(), g["native.loc"], atol=2e-3, rtol=0)
    np.testing.assert_allclose(conf.numpy(), g["native.conf"], atol=2e-3, rtol=0)
    np.testing.assert_allclose(landms.numpy(), g["native.landms"], atol=2e-3, rtol=0)
    shapes = torch.tensor([[96.0, 128.0]] * 2)
    rows = ro.postprocess(torch.from_numpy(g["native.loc"]), torch.from_numpy(g["native.conf"]),
                          torch.from_numpy(g["native.landms"]), shapes, (96, 128), 0.5, 0.4)
    for i, r in enumerate(rows):
        assert np.array_equal(r.numpy().reshape(-1, 15), g[f"native.rows{i}"])
