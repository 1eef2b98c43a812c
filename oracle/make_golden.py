"""TEST INFRASTRUCTURE ONLY — generates the committed fixtures by EXECUTING THE UNMODIFIED REFERENCE
(imported read-only from /root/reference through oracle/ref_shims.py) in the build container.

    python oracle/make_golden.py            # writes  video-desensitization_b200/vdesens_b200/data/bn_calib{,_mnet}_seed0.npz
                                            #         tests/golden/{ref_detect,ref_detect_mnet,nms,cv}_golden.npz
    python oracle/make_golden.py --mnet-only   # only the MobileNet0.25 files
    python oracle/make_golden.py --default-only  # only the un-conditioned default-init files (bn_calib_default_seed0.npz,
                                               #         tests/golden/ref_detect_default_golden.npz)

The GPU box has no /root/reference; tests there read only the committed .npz files.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-desensitization_b200"))

from oracle import retinaface_oracle as ro          # noqa: E402
from oracle import cv_oracle as co                  # noqa: E402
from oracle.ref_shims import load_reference         # noqa: E402
from vdesens_b200 import synth, arch                # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RATES = {"r01": 0.01, "dense": 0.0025, "r001": 0.001}


def make_bn_calib(init="conditioned"):
    """BN statistics of the seed-0 random init on a 4-frame synthetic batch + class-bias shifts.
    init="default": the un-conditioned default initialisation (gamma = 1, beta = 0 everywhere)."""
    sd = ro.to_torch_sd(synth._raw_state_dict(0, init=init))
    frames = synth.synthetic_clip(4, 480, 640, base_seed=7000)
    x, _ = ro.preprocess(frames, (480, 640))
    ro.forward(sd, x, calibrate=True)
    out = {k: v.numpy() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    # class-bias shifts: quantiles of the face-vs-background logit difference per pyramid level, estimated on
    # eight 1080p frames (the coarse level has only 4080 anchors per frame, so small samples are too noisy)
    ds = [[], [], []]
    offs = arch.level_offsets(1080, 1920) + [arch.num_anchors(1080, 1920)]
    for i in range(8):
        xf, _ = ro.preprocess([synth.synthetic_frame(1080, 1920, 7100 + i)], (1080, 1920))
        _, cls, _ = ro.forward(sd, xf, softmax=False)
        d = (cls[..., 1] - cls[..., 0])[0]
        for l in range(3):
            ds[l].append(d[offs[l]:offs[l + 1]])
    for name, rate in RATES.items():
        sh = [-float(np.quantile(torch.cat(ds[l]).numpy().astype(np.float64), 1 - rate)) for l in range(3)]
        out["shift." + name] = np.array(sh, np.float32)
    path = synth.BN_CALIB_FILE if init == "conditioned" else synth.BN_CALIB_FILE_DEFAULT
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def make_ref_detect_default(ref_face):
    """the UNMODIFIED reference on the un-conditioned default-init weights (one small case: pins the oracle on them)"""
    sd = synth.synthetic_state_dict(0, face_rate="r01", init="default")
    tmp = tempfile.mkdtemp()
    pth = os.path.join(tmp, "w_default.pth")
    synth.save_state_dict_pth(sd, pth)
    frames = [synth.synthetic_frame(96, 128, 9000), synth.synthetic_frame(96, 128, 9001)]
    det = ref_face.Retinaface(model_path=pth, backbone="resnet50", input_shape=[96, 128, 3], confidence=0.5, nms_iou=0.4,
                              letterbox_image=True, cuda=False)
    x, shapes = det.preprocess(frames)
    with torch.no_grad():
        loc, conf, landms = det.net(x)
        rows = det.postprocess(loc, conf, landms, shapes, len(frames))
    out = {"native.frames": np.array([(96, 128, 9000), (96, 128, 9001)], np.int64),
           "native.input_shape": np.array([96, 128, 3], np.int64),
           "native.loc": loc.numpy(), "native.conf": conf.numpy(), "native.landms": landms.numpy()}
    for i, r in enumerate(rows):
        out[f"native.rows{i}"] = r.numpy().astype(np.float32).reshape(-1, 15)
    print("default-init native kept per frame:", [len(r) for r in rows])
    np.savez_compressed(os.path.join(GOLD, "ref_detect_default_golden.npz"), **out)


def make_ref_detect(ref_face):
    sd = synth.synthetic_state_dict(0, face_rate="r01")
    tmp = tempfile.mkdtemp()
    pth = os.path.join(tmp, "w.pth")
    synth.save_state_dict_pth(sd, pth)
    cases = {
        "native": dict(frames=[(96, 128, 9000), (96, 128, 9001)], input_shape=[96, 128, 3]),
        "letterbox": dict(frames=[(150, 200, 9100), (260, 180, 9101)], input_shape=[128, 128, 3]),
        "c1": dict(frames=[(480, 640, 1000), (480, 640, 1001)], input_shape=[640, 640, 3]),
        "odd": dict(frames=[(135, 241, 9200)], input_shape=[135, 241, 3]),
    }
    out = {}
    for name, cs in cases.items():
        frames = [synth.synthetic_frame(h, w, s) for h, w, s in cs["frames"]]
        det = ref_face.Retinaface(model_path=pth, backbone="resnet50", input_shape=cs["input_shape"],
                                  confidence=0.5, nms_iou=0.4, letterbox_image=True, cuda=False)
        x, shapes = det.preprocess(frames)
        with torch.no_grad():
            loc, conf, landms = det.net(x)
            rows = det.postprocess(loc, conf, landms, shapes, len(frames))
        res = det.detect_images(frames)
        out[name + ".frames"] = np.array(cs["frames"], np.int64)
        out[name + ".input_shape"] = np.array(cs["input_shape"], np.int64)
        if name != "c1":
            out[name + ".x"] = x.numpy().astype(np.float32)
            out[name + ".loc"] = loc.numpy()
            out[name + ".conf"] = conf.numpy()
            out[name + ".landms"] = landms.numpy()
        for i, (r, (img, boxes)) in enumerate(zip(rows, res)):
            out[f"{name}.rows{i}"] = r.numpy().astype(np.float32).reshape(-1, 15)
            out[f"{name}.boxes{i}"] = np.array(boxes, np.float64).reshape(-1, 4)
            assert np.array_equal(img, frames[i])
        print(name, "kept per frame:", [len(r) for r in rows])
    np.savez_compressed(os.path.join(GOLD, "ref_detect_golden.npz"), **out)


def make_bn_calib_mnet():
    """same recipe for the cfg_mnet network."""
    from vdesens_b200 import arch_mnet as am
    sd = ro.to_torch_sd(synth._raw_state_dict(0, "mobilenet"))
    frames = synth.synthetic_clip(4, 480, 640, base_seed=7000)
    x, _ = ro.preprocess(frames, (480, 640))
    ro.forward_mnet(sd, x, calibrate=True)
    out = {k: v.numpy() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")}
    ds = [[], [], []]
    offs = am.level_offsets(1080, 1920) + [am.num_anchors(1080, 1920)]
    for i in range(8):
        xf, _ = ro.preprocess([synth.synthetic_frame(1080, 1920, 7100 + i)], (1080, 1920))
        _, cls, _ = ro.forward_mnet(sd, xf, softmax=False)
        d = (cls[..., 1] - cls[..., 0])[0]
        for l in range(3):
            ds[l].append(d[offs[l]:offs[l + 1]])
    for name, rate in RATES.items():
        sh = [-float(np.quantile(torch.cat(ds[l]).numpy().astype(np.float64), 1 - rate)) for l in range(3)]
        out["shift." + name] = np.array(sh, np.float32)
    np.savez_compressed(synth.BN_CALIB_FILE_MNET, **out)
    print("wrote", synth.BN_CALIB_FILE_MNET, os.path.getsize(synth.BN_CALIB_FILE_MNET), "bytes")


def make_ref_detect_mnet(ref_face):
    """the unmodified reference with backbone='mobilenet' (face.py:35) on the synthetic cfg_mnet weights"""
    sd = synth.synthetic_state_dict(0, face_rate="r01", backbone="mobilenet")
    tmp = tempfile.mkdtemp()
    pth = os.path.join(tmp, "w_mnet.pth")
    synth.save_state_dict_pth(sd, pth)
    cases = {
        "native": dict(frames=[(96, 128, 9000), (96, 128, 9001)], input_shape=[96, 128, 3]),
        "letterbox": dict(frames=[(150, 200, 9100), (260, 180, 9101)], input_shape=[128, 128, 3]),
        "odd": dict(frames=[(135, 241, 9200)], input_shape=[135, 241, 3]),
        "vga": dict(frames=[(480, 640, 1000)], input_shape=[480, 640, 3]),
    }
    out = {}
    for name, cs in cases.items():
        frames = [synth.synthetic_frame(h, w, s) for h, w, s in cs["frames"]]
        det = ref_face.Retinaface(model_path=pth, backbone="mobilenet", input_shape=cs["input_shape"],
                                  confidence=0.5, nms_iou=0.4, letterbox_image=True, cuda=False)
        x, shapes = det.preprocess(frames)
        with torch.no_grad():
            loc, conf, landms = det.net(x)
            rows = det.postprocess(loc, conf, landms, shapes, len(frames))
        res = det.detect_images(frames)
        out[name + ".frames"] = np.array(cs["frames"], np.int64)
        out[name + ".input_shape"] = np.array(cs["input_shape"], np.int64)
        out[name + ".loc"] = loc.numpy()
        out[name + ".conf"] = conf.numpy()
        out[name + ".landms"] = landms.numpy()
        for i, (r, (img, boxes)) in enumerate(zip(rows, res)):
            out[f"{name}.rows{i}"] = r.numpy().astype(np.float32).reshape(-1, 15)
            out[f"{name}.boxes{i}"] = np.array(boxes, np.float64).reshape(-1, 4)
        print("mnet", name, "kept per frame:", [len(r) for r in rows])
    np.savez_compressed(os.path.join(GOLD, "ref_detect_mnet_golden.npz"), **out)


def make_nms():
    from torchvision.ops import nms
    rng = np.random.default_rng(4242)
    out = {}
    for ci, (n, ncl) in enumerate([(1, 1), (7, 2), (300, 12), (1500, 40), (4000, 60)]):
        centers = rng.uniform(0.1, 0.9, size=(ncl, 2))
        which = rng.integers(0, ncl, size=n)
        c = centers[which] + rng.normal(0, 0.004, size=(n, 2))
        wh = rng.uniform(0.02, 0.06, size=(ncl, 2))[which] * rng.uniform(0.8, 1.25, size=(n, 2))
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        scores = rng.permutation(np.linspace(0.5, 0.999, n)).astype(np.float32)   # distinct: no ties
        keep = nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.4).numpy()
        out[f"c{ci}.boxes"], out[f"c{ci}.scores"], out[f"c{ci}.keep"] = boxes, scores, keep
        print("nms case", ci, n, "->", len(keep))
    np.savez_compressed(os.path.join(GOLD, "nms_golden.npz"), **out)


def make_cv(ref_face, ref_combine):
    import cv2
    rng = np.random.default_rng(777)
    out = {}
    gi = 0
    for k, sg in [(3, 0), (5, 0), (7, 0), (9, 0), (15, 2.5), (31, 0), (51, 0), (99, 0), (31, 30.0)]:
        for (h, w) in [(1, 1), (3, 50), (40, 37), (96, 130)]:
            roi = synth.synthetic_frame(max(h, 8), max(w, 8), 500 + gi)[:h, :w].copy()
            out[f"g{gi}.roi"] = roi
            out[f"g{gi}.par"] = np.array([k, sg], np.float64)
            out[f"g{gi}.out"] = cv2.GaussianBlur(roi, (k, k), sg)
            out[f"g{gi}.taps"] = co.gaussian_taps_q8(k, sg)
            gi += 1
    out["g.count"] = np.array(gi)
    # sequential mosaic via the reference function (combine_detect.py:138-161, 247-249)
    img = synth.synthetic_frame(120, 160, 31)
    boxes = synth.synthetic_boxes(12, 120, 160, seed=5, min_size=3, max_size=70)
    m = img.copy()
    for b in boxes:
        m = ref_combine.mosaic_rectangle_region_single(m, *b, mosaic_level=8)
    out["mosaic.img"], out["mosaic.boxes"], out["mosaic.out"] = img, np.array(boxes, np.int64), m
    gimg = img.copy()
    for (x1, y1, x2, y2) in boxes:
        x1c, y1c, x2c, y2c = max(0, x1), max(0, y1), min(160, x2), min(120, y2)
        if x2c > x1c and y2c > y1c:
            gimg[y1c:y2c, x1c:x2c] = cv2.GaussianBlur(gimg[y1c:y2c, x1c:x2c], (15, 15), 0)
    out["gauss.out"] = gimg
    # letterbox via the reference function (detect_face/utils/utils.py:8-18)
    for li, (ih, iw, W, H) in enumerate([(150, 200, 128, 128), (260, 180, 128, 128), (96, 128, 128, 96),
                                          (270, 480, 160, 160)]):
        f = synth.synthetic_frame(ih, iw, 800 + li)
        out[f"lb{li}.par"] = np.array([ih, iw, W, H, 800 + li], np.int64)
        out[f"lb{li}.out"] = ref_face.letterbox_image(f, [W, H]).astype(np.uint8)
    out["lb.count"] = np.array(4)
    np.savez_compressed(os.path.join(GOLD, "cv_golden.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLD, exist_ok=True)
    ref_face, ref_combine = load_reference()
    if "--default-only" in sys.argv:
        make_bn_calib("default")
        make_ref_detect_default(ref_face)
        sys.exit(0)
    if "--mnet-only" in sys.argv:
        make_bn_calib_mnet()
        make_ref_detect_mnet(ref_face)
        sys.exit(0)
    make_bn_calib()
    make_ref_detect(ref_face)
    make_bn_calib("default")
    make_ref_detect_default(ref_face)
    make_bn_calib_mnet()
    make_ref_detect_mnet(ref_face)
    make_nms()
    make_cv(ref_face, ref_combine)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
