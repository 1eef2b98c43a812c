"""Seeded synthetic inputs for benches, smoke and tests: frames, plate boxes and a random-init
RetinaFace-ResNet50 state_dict in the reference's `.pth` key layout (detect_face/face.py:52-56:
keys carry no `module.` prefix).

Everything is generated with numpy's PCG64 `default_rng`, which is bit-identical on every host,
so the container that produced tests/golden/* and the GPU box see the same bytes.  The BN running
statistics come from a committed calibration file (data/bn_calib_seed0.npz, written by
oracle/make_golden.py): plain random-init weights with identity BN statistics saturate every score
at 0/1 and overflow exp() in the box decode, so they are unusable for parity.
"""
import os
import numpy as np

from . import arch

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
BN_CALIB_FILE = os.path.join(_DATA_DIR, "bn_calib_seed0.npz")
BN_CALIB_FILE_MNET = os.path.join(_DATA_DIR, "bn_calib_mnet_seed0.npz")
BN_CALIB_FILE_DEFAULT = os.path.join(_DATA_DIR, "bn_calib_default_seed0.npz")


def synthetic_frame(h: int, w: int, seed: int) -> np.ndarray:
    """uint8 (h,w,3) frame with low-frequency content (so blur/mosaic are visible) plus mild
    per-pixel noise: a (h/8+2, w/8+2, 3) uniform grid, bilinearly upscaled."""
    rng = np.random.default_rng(seed)
    gh, gw = h // 8 + 2, w // 8 + 2
    grid = rng.uniform(0.0, 255.0, size=(gh, gw, 3)).astype(np.float32)
    ys = (np.arange(h, dtype=np.float32) + 0.5) / 8.0
    xs = (np.arange(w, dtype=np.float32) + 0.5) / 8.0
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    g00 = grid[y0][:, x0]
    g01 = grid[y0][:, x0 + 1]
    g10 = grid[y0 + 1][:, x0]
    g11 = grid[y0 + 1][:, x0 + 1]
    img = (g00 * (1 - fy) * (1 - fx) + g01 * (1 - fy) * fx + g10 * fy * (1 - fx) + g11 * fy * fx)
    img += rng.uniform(-6.0, 6.0, size=img.shape).astype(np.float32)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def synthetic_clip(n: int, h: int, w: int, base_seed: int = 1000):
    return [synthetic_frame(h, w, base_seed + i) for i in range(n)]


def synthetic_boxes(n: int, h: int, w: int, seed: int, min_size: int = 32, max_size: int = 256):
    """n integer boxes (x1,y1,x2,y2); may overlap and may cross the frame border (exercises the
    clip rules of reference combine_detect.py:145-150)."""
    rng = np.random.default_rng(seed)
    bw = rng.integers(min_size, max_size + 1, size=n)
    bh = rng.integers(min_size, max_size + 1, size=n)
    x1 = rng.integers(-max_size // 4, w - min_size // 2, size=n)
    y1 = rng.integers(-max_size // 4, h - min_size // 2, size=n)
    return [(int(a), int(b), int(a + c), int(b + d)) for a, b, c, d in zip(x1, y1, bw, bh)]


def _raw_state_dict(seed: int, backbone: str = "resnet50", init: str = "conditioned"):
    """numpy state dict with identity BN statistics.
    init="conditioned" (what every fixture of round 1 uses): random BN gamma/beta with SMALL residual gammas (bn3),
    which keeps an untrained 50-layer net from amplifying rounding noise — see the comment at bn3 below.
    init="default" (SURVEY §8d to the letter): the reference constructor's own initialisation — every BatchNorm
    gamma = 1, beta = 0, nothing shrunk.  That network is in the chaotic regime of random deep ReLU nets: a relative
    perturbation of 2^-12 anywhere is amplified ~100x by the heads; the parity test on these weights
    (tests/test_gpu_kernels.py::test_default_init_weights_1080p) states what each precision mode achieves on them."""
    assert init in ("conditioned", "default")
    rng = np.random.default_rng(seed + (0 if backbone == "resnet50" else 7919))
    sd = {}
    if backbone == "resnet50":
        spec = arch.state_dict_spec()
    else:
        from . import arch_mnet
        spec = arch_mnet.state_dict_spec()
    for key, shape, kind in spec:
        if kind == "conv_backbone":      # kaiming-normal, fan_out (torchvision resnet init)
            fan_out = shape[0] * shape[2] * shape[3]
            sd[key] = rng.normal(0.0, np.sqrt(2.0 / fan_out), size=shape).astype(np.float32)
        elif kind == "conv_dw":
            # depthwise 3x3: positive taps (a local weighted mean).  Thirteen random-SIGN depthwise filters in a row
            # on smooth images are thirteen near-cancelling sums, each re-normalised by its BN: an untrained net
            # built that way amplifies fp16 rounding to 5 % and bf16 rounding to 27 % of the head outputs, which no
            # reduced-precision execution can track (same reasoning as bn3 below)
            sd[key] = rng.uniform(0.0, 0.25, size=shape).astype(np.float32)
        elif kind in ("conv_neck", "head_weight"):   # torch Conv2d default: U(+-1/sqrt(fan_in))
            fan_in = shape[1] * shape[2] * shape[3]
            b = 1.0 / np.sqrt(fan_in)
            sd[key] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif kind == "head_bias":
            b = 1.0 / np.sqrt(256.0 if backbone == "resnet50" else 64.0)
            sd[key] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif kind == "bn_weight" and init == "default":
            sd[key] = np.ones(shape, np.float32)
        elif kind == "bn_bias" and init == "default":
            sd[key] = np.zeros(shape, np.float32)
        elif kind == "bn_weight":        # non-trivial gamma/beta so BN folding is really exercised
            if key.endswith(".bn3.weight"):
                # small residual branches (cf. torchvision's zero_init_residual recipe): with gamma ~ 1 on
                # every bottleneck output an untrained ResNet-50 amplifies rounding noise ~3.5x per stage
                # (32 % feature error at layer4 in bf16), which no reduced-precision path can track
                sd[key] = rng.uniform(0.05, 0.2, size=shape).astype(np.float32)
            else:
                sd[key] = rng.uniform(0.6, 1.4, size=shape).astype(np.float32)
        elif kind == "bn_bias":
            sd[key] = rng.uniform(-0.25, 0.25, size=shape).astype(np.float32)
        elif kind == "bn_mean":
            sd[key] = np.zeros(shape, np.float32)
        elif kind == "bn_var":
            sd[key] = np.ones(shape, np.float32)
        elif kind == "bn_count":
            sd[key] = np.array(1, np.int64)
        else:
            raise AssertionError(kind)
    return sd


def synthetic_state_dict(seed: int = 0, calibrated: bool = True, face_rate: str = "r01", backbone: str = "resnet50",
                         init: str = "conditioned"):
    """Returns {key: np.ndarray}.  `calibrated` loads the committed BN statistics for seed 0;
    `face_rate` picks the class-bias shift stored with them ("r01" ~ 1 % of anchors >= 0.5,
    "r001" ~ 0.1 %, "dense" for the dense-ROI stress case, "" for no shift).
    backbone "mobilenet" gives the cfg_mnet network (reference detect_face/utils/config.py:1-16).
    init: see _raw_state_dict ("default" = un-conditioned default initialisation, resnet50 only)."""
    if init == "default" and backbone != "resnet50":
        raise ValueError('init="default" is provided for the resnet50 backbone only')
    sd = _raw_state_dict(seed, backbone, init)
    if calibrated:
        if seed != 0:
            raise ValueError("calibration statistics are committed for seed 0 only")
        calib_file = BN_CALIB_FILE if backbone == "resnet50" else BN_CALIB_FILE_MNET
        if init == "default":
            calib_file = BN_CALIB_FILE_DEFAULT
        if not os.path.isfile(calib_file):
            raise FileNotFoundError(calib_file + " (run oracle/make_golden.py in the build container)")
        cal = np.load(calib_file)
        for key in cal.files:
            if key.startswith("shift."):
                continue
            assert key in sd and sd[key].shape == cal[key].shape, key
            sd[key] = cal[key].astype(np.float32)
        if face_rate:
            shift = cal["shift." + face_rate].astype(np.float32)   # (3,) one per pyramid level
            for lvl in range(3):
                k = f"ClassHead.{lvl}.conv1x1.bias"
                b = sd[k].copy()
                b[1::2] += shift[lvl]      # face logit of both anchors (rows: a0 bg, a0 face, a1 bg, a1 face)
                sd[k] = b
    return sd


def save_state_dict_pth(sd, path: str):
    """Writes the reference's weight format: torch.save of an un-prefixed state_dict."""
    import torch
    torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, path)
