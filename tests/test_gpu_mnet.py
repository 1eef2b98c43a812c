"""GPU parity tests of the MobileNet0.25 variant (Retinaface(backbone="mobilenet")): the three
streaming kernels and the slope epilogue of the tensor-core kernel against torch on the CPU, the whole network against
the emulated op list (same packed weights, same rounding points) and against the committed outputs of the executed
reference, and the façade end to end."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import retinaface_oracle as ro
from vdesens_b200 import arch_mnet as am, synth, weights

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from vdesens_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_detect_mnet_golden.npz"))


@pytest.fixture(scope="module")
def sd_np():
    return synth.synthetic_state_dict(0, face_rate="r01", backbone="mobilenet")


def _act(y, slope):
    return torch.where(y > 0, y, y * slope.view(1, -1, 1, 1))


def _close_to_rounded(got, want_f32, dtype):
    """`got` (device result, stored in dtype) against fp32 math rounded once: at most one rounding step apart."""
    want = want_f32.to(dtype).float()
    ulp = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)
    err = (got.float() - want).abs()
    tol = want_f32.abs() * (2 * ulp) + 1e-5
    assert bool((err <= tol).all()), f"max err {float(err.max()):.3e}, worst excess {float((err - tol).max()):.3e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("hw", [(16, 24), (45, 77), (135, 241)])
def test_mnet_stem(ops, hw, dtype):
    H, W = hw
    B = 2
    g = torch.Generator().manual_seed(5)
    frames = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, generator=g) for _ in range(B)]
    x0 = ops.ingest([f.to(DEV) for f in frames], H, W, precision="bf16" if dtype == torch.bfloat16 else "fp16")
    x = torch.stack([f.float() for f in frames]).permute(0, 3, 1, 2) - torch.tensor([104.0, 117.0, 123.0]).view(1, 3, 1, 1)
    w = (torch.randn(8, 3, 3, 3, generator=g) * 0.02).to(dtype)
    bias = torch.randn(8, generator=g)
    slope = torch.tensor([0.1, 0.0, 1.0, 0.1, 0.1, 0.0, 0.1, 0.5])
    wp = torch.zeros(3, 3, 4, 8)
    wp[:, :, :3, :] = w.float().permute(2, 3, 1, 0)
    y = ops.mnet_stem(x0, wp.reshape(36, 8).to(dtype).to(DEV), bias.to(DEV), slope.to(DEV), H, W)
    want = _act(F.conv2d(x.to(dtype).float(), w.float(), bias, 2, 1), slope)
    assert tuple(y.shape) == (B, (H + 1) // 2, (W + 1) // 2, 8)
    _close_to_rounded(y.cpu().permute(0, 3, 1, 2), want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,stride,hw", [(8, 1, (23, 31)), (16, 2, (23, 31)), (32, 1, (40, 64)), (64, 2, (33, 17)),
                                         (128, 1, (9, 12)), (256, 2, (7, 5))])
def test_mnet_depthwise(ops, C, stride, hw, dtype):
    H, W = hw
    g = torch.Generator().manual_seed(C + stride)
    x = torch.randn(2, C, H, W, generator=g).to(dtype)
    w = (torch.randn(C, 1, 3, 3, generator=g) * 0.3).to(dtype)
    bias = torch.randn(C, generator=g)
    slope = torch.rand(C, generator=g)
    y = ops.mnet_dw3x3(x.permute(0, 2, 3, 1).contiguous().to(DEV), w[:, 0].permute(1, 2, 0).reshape(9, C).contiguous().to(DEV),
                       bias.to(DEV), slope.to(DEV), stride)
    want = _act(F.conv2d(x.float(), w.float(), bias, stride, 1, 1, C), slope)
    assert tuple(y.shape) == (2, want.shape[2], want.shape[3], C)
    _close_to_rounded(y.cpu().permute(0, 3, 1, 2), want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout", [(8, 16), (16, 32), (32, 32), (32, 64)])
def test_mnet_pointwise(ops, cin, cout, dtype):
    g = torch.Generator().manual_seed(cin * cout)
    x = torch.randn(2, cin, 19, 27, generator=g).to(dtype)
    w = (torch.randn(cout, cin, generator=g) * 0.2).to(dtype)
    bias = torch.randn(cout, generator=g)
    slope = torch.rand(cout, generator=g)
    y = ops.mnet_pw(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.t().contiguous().to(DEV), bias.to(DEV), slope.to(DEV))
    want = _act(F.conv2d(x.float(), w.float().view(cout, cin, 1, 1), bias), slope)
    # fp32 accumulation over <= 32 terms in a different order than torch: a rounding step may flip
    _close_to_rounded(y.cpu().permute(0, 3, 1, 2), want, dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout,stride,shape", [(8, 16, 1, (2, 23, 31)), (16, 32, 2, (2, 23, 31)), (32, 32, 1, (1, 40, 64)),
                                                   (32, 64, 2, (3, 33, 17)), (32, 64, 2, (2, 67, 121)), (8, 16, 1, (1, 8, 32)),
                                                   # more tiles than resident CTAs: the persistent kernel's tile loop, its
                                                   # prefetch across tiles and the division-free tile walk
                                                   (8, 16, 1, (4, 270, 480)), (16, 32, 2, (4, 270, 480)),
                                                   (32, 32, 1, (2, 135, 241)), (32, 64, 2, (3, 135, 239))])
def test_mnet_fused_dwpw_is_bit_identical_to_two_kernels(ops, cin, cout, stride, shape, dtype):
    B, H, W = shape
    g = torch.Generator().manual_seed(cin * cout + stride + H)
    x = torch.randn(B, H, W, cin, generator=g).to(dtype).to(DEV)
    wdw = (torch.randn(9, cin, generator=g) * 0.3).to(dtype).to(DEV)
    bdw, sdw = torch.randn(cin, generator=g).to(DEV), torch.rand(cin, generator=g).to(DEV)
    wpw = (torch.randn(cin, cout, generator=g) * 0.2).to(dtype).to(DEV)
    bpw, spw = torch.randn(cout, generator=g).to(DEV), torch.rand(cout, generator=g).to(DEV)
    want = ops.mnet_pw(ops.mnet_dw3x3(x, wdw, bdw, sdw, stride), wpw, bpw, spw)
    got = ops.mnet_dwpw(x, wdw, bdw, sdw, wpw, bpw, spw, stride)
    assert got.shape == want.shape and torch.equal(got, want), float((got.float() - want.float()).abs().max())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cin,cout,stride,shape", [(8, 16, 1, (3, 200, 333)), (32, 64, 2, (3, 135, 239))])
def test_mnet_fused_dwpw_against_torch(ops, cin, cout, stride, shape, dtype):
    """the fused conv_dw block at a size with several tiles per persistent CTA against torch (fp32 math, the depthwise
    result rounded to the storage type where the kernel rounds it)."""
    B, H, W = shape
    g = torch.Generator().manual_seed(11 * cin + cout + W)
    x = torch.randn(B, cin, H, W, generator=g).to(dtype)
    wdw = (torch.randn(cin, 1, 3, 3, generator=g) * 0.3).to(dtype)
    bdw, sdw = torch.randn(cin, generator=g), torch.rand(cin, generator=g)
    wpw = (torch.randn(cout, cin, generator=g) * 0.2).to(dtype)
    bpw, spw = torch.randn(cout, generator=g), torch.rand(cout, generator=g)
    got = ops.mnet_dwpw(x.permute(0, 2, 3, 1).contiguous().to(DEV), wdw[:, 0].permute(1, 2, 0).reshape(9, cin).contiguous().to(DEV),
                        bdw.to(DEV), sdw.to(DEV), wpw.t().contiguous().to(DEV), bpw.to(DEV), spw.to(DEV), stride)
    d = _act(F.conv2d(x.float(), wdw.float(), bdw, stride, 1, 1, cin), sdw).to(dtype).float()
    want = _act(F.conv2d(d, wpw.float().view(cout, cin, 1, 1), bpw), spw)
    got = got.cpu().permute(0, 3, 1, 2).float()
    assert got.shape == want.shape
    step = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)
    err = (got - want).abs()
    # one rounding step of the output plus what a flipped rounding of a depthwise value moves it (|w| <= ~0.8, |d| step)
    tol = want.abs() * (2 * step) + d.abs().max() * step * 1.0 + 1e-5
    assert bool((err <= tol).all()), f"max err {float(err.max()):.3e}, worst excess {float((err - tol).max()):.3e}"
    assert float((err.pow(2).mean().sqrt() / want.pow(2).mean().sqrt())) <= 2 * step


def test_mnet_kernel_argument_errors(ops):
    x = torch.zeros((1, 4, 4, 12), dtype=torch.bfloat16, device=DEV)
    w = torch.zeros((9, 12), dtype=torch.bfloat16, device=DEV)
    v = torch.zeros((12,), dtype=torch.float32, device=DEV)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.mnet_dw3x3(x, w, v, v, 1)
    x = torch.zeros((1, 4, 4, 128), dtype=torch.bfloat16, device=DEV)
    w = torch.zeros((128, 64), dtype=torch.bfloat16, device=DEV)
    v = torch.zeros((64,), dtype=torch.float32, device=DEV)
    with pytest.raises(RuntimeError, match="4096"):
        ops.mnet_pw(x, w, v, v)


@pytest.mark.parametrize("k,cin,cout,hw", [(1, 64, 64, (20, 28)), (3, 64, 64, (17, 23)), (1, 128, 64, (12, 9)),
                                           (1, 256, 256, (5, 7)), (3, 64, 64, (68, 120))])
def test_conv_slope_epilogue(ops, k, cin, cout, hw):
    """per-channel negative-side slope (ReLU / LeakyReLU(0.1) / identity mixed in one tile) on the tcgen05 kernel."""
    H, W = hw
    g = torch.Generator().manual_seed(k * cin + cout)
    x = torch.randn(2, cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, generator=g) * (1.0 / (k * cin ** 0.5))).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g) * 0.1
    slope = torch.tensor([0.0, 0.1, 1.0, 0.1])[torch.randint(0, 4, (cout,), generator=g)]
    y = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().to(DEV), w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(DEV),
                   bias.to(DEV), k, 1, k // 2, slope=slope.to(DEV))
    want = _act(F.conv2d(x.float(), w.float(), bias, 1, k // 2), slope)
    got = y.cpu().permute(0, 3, 1, 2).float()
    err = (got - want).abs()
    assert float(err.max()) <= 0.02 * float(want.abs().max()), float(err.max())
    # the sign-dependent branch is taken on the fp32 accumulator: no channel may show the wrong activation
    neg = want < -0.05
    assert bool(neg.any()) and float((got - want).abs()[neg].max()) <= 0.02 * float(want.abs().max())


def _engine(sd_np, frames, H, W, precision="bf16", fuse_dwpw=True):
    from vdesens_b200.engine import DetectorEngine
    eng = DetectorEngine(sd_np, len(frames), H, W, device=DEV, confidence=0.5, nms_iou=0.4, precision=precision,
                         backbone="mobilenet", fuse_dwpw=fuse_dwpw)
    eng.set_frames([torch.from_numpy(f).to(DEV) for f in frames])
    eng.forward()
    return eng


def _frames(g, name):
    return [synth.synthetic_frame(int(h), int(w), int(s)) for h, w, s in g[name + ".frames"]]


@pytest.mark.parametrize("precision,dtype", [("bf16", torch.bfloat16), ("fp16", torch.float16)])
@pytest.mark.parametrize("case", ["native", "odd", "vga"])
def test_mnet_network_vs_emulated_op_list(gold, sd_np, case, precision, dtype):
    """device network against the SAME op list on the CPU (same packed weights, same rounding points, fp32 math):
    only the accumulation order differs.  Early layers are tight; flips of a rounding step propagate through the 13
    re-normalising depthwise stages of this untrained network like any other storage noise (bounds stated below)."""
    frames = _frames(gold, case)
    H, W = (int(v) for v in gold[case + ".input_shape"][:2])
    eng = _engine(sd_np, frames, H, W, precision, fuse_dwpw=False)       # layer-by-layer form: every tensor exists
    loc, cls, ldm = (t.cpu() for t in eng.head_tensors())
    # the engine's default form (conv_dw blocks with Cin <= 32 as one kernel) gives the same bits
    engf = _engine(sd_np, frames, H, W, precision)
    assert sum(1 for k, _, _ in engf.plans if k == "dwpw") == 4 and "body.stage1.1.dw" not in engf.buf
    for a, b in zip(engf.head_tensors(), eng.head_tensors()):
        assert torch.equal(a, b)
    assert torch.equal(engf.buf["body.stage1.4.pw"], eng.buf["body.stage1.4.pw"])
    x, _ = ro.preprocess(frames, (H, W))
    cap = {}
    eloc, ecls, eldm = ro.forward_mnet_emulated(weights.pack_mnet(sd_np, dtype), x, capture=cap, dtype=dtype)
    report = []
    for name in ["m.stem", "body.stage1.1.dw", "body.stage1.1.pw", "body.stage1.5.pw", "body.stage2.5.pw",
                 "body.stage3.1.pw", "fpn.o3", "fpn.o2", "fpn.o1", "ssh1.S1", "ssh1.S3", "ssh3.S3"]:
        a, b = eng.activation(name).cpu(), cap[name]
        assert a.shape == b.shape, (name, a.shape, b.shape)
        mx = float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
        rms = float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-6))
        report.append((name, mx, rms))
    msg = " ".join(f"{n}:{m:.4f}/{r:.4f}" for n, m, r in report)
    print(msg)
    step = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)
    assert report[0][1] <= 2 * step and report[1][1] <= 4 * step, msg           # stem, first depthwise: one rounding step
    lim_max, lim_rms = (0.10, 0.03) if dtype == torch.bfloat16 else (0.02, 0.005)
    assert all(m <= lim_max and r <= lim_rms for _, m, r in report), msg
    lim = 0.08 if dtype == torch.bfloat16 else 0.012
    assert float((loc - eloc).abs().max()) <= lim, msg
    assert float((cls - ecls).abs().max()) <= lim, msg
    assert float((ldm - eldm).abs().max()) <= lim, msg
    # dead lanes of the padded SSH tensors are exactly zero (ReLU of a zero accumulator)
    for lvl in (1, 2, 3):
        assert float(eng.activation(f"ssh{lvl}.S1")[:, 48:].abs().max()) == 0.0
        assert float(eng.activation(f"ssh{lvl}.S3")[:, 32:].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["letterbox", "odd", "vga"])
def test_mnet_fp16_heads_vs_reference_fp32(gold, sd_np, case):
    """fp16-storage network vs the EXECUTED fp32 reference (committed golden tensors of backbone='mobilenet').
    Stated tolerance (CPU emulation bound, tests/test_mnet_cpu.py): |head err| <= 0.06; boxes IoU >= 0.97 on every
    anchor and >= 0.99 on at least 90 % of them."""
    frames = _frames(gold, case)
    H, W = (int(v) for v in gold[case + ".input_shape"][:2])
    eng = _engine(sd_np, frames, H, W, "fp16")
    loc, cls, ldm = (t.cpu() for t in eng.head_tensors())
    conf = torch.softmax(cls, -1).numpy()
    dl = np.abs(loc.numpy() - gold[case + ".loc"]).max()
    dc = np.abs(conf - gold[case + ".conf"]).max()
    dm = np.abs(ldm.numpy() - gold[case + ".landms"]).max()
    pri = ro.anchors(H, W).unsqueeze(0)
    b1, b2 = ro.decode(loc, pri).numpy(), ro.decode(torch.from_numpy(gold[case + ".loc"]), pri).numpy()
    x1, y1 = np.maximum(b1[..., 0], b2[..., 0]), np.maximum(b1[..., 1], b2[..., 1])
    x2, y2 = np.minimum(b1[..., 2], b2[..., 2]), np.minimum(b1[..., 3], b2[..., 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    iou = inter / ((b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1]) + (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1]) - inter)
    print(f"loc {dl:.4f} conf {dc:.4f} landm {dm:.4f} IoU min {iou.min():.4f} frac>=0.99 {(iou >= 0.99).mean():.3f}")
    assert dl <= 0.06 and dc <= 0.02 and dm <= 0.06
    assert iou.min() >= 0.97 and (iou >= 0.99).mean() >= 0.90


def test_mnet_detect_rows_vs_oracle_postprocess_of_device_heads(gold, sd_np):
    """decode / threshold / NMS / un-letterbox of the mobilenet engine: the device rows equal the oracle's
    post-processing applied to the device's own head tensors (kept sets identical, values within 1e-3)."""
    frames = _frames(gold, "letterbox")
    H, W = (int(v) for v in gold["letterbox.input_shape"][:2])
    eng = _engine(sd_np, frames, H, W, "fp16")
    rows, counts = eng.fetch_results()
    loc, cls, ldm = (t.cpu() for t in eng.head_tensors())
    shapes = torch.tensor([[f.shape[0], f.shape[1]] for f in frames], dtype=torch.float32)
    want = ro.postprocess(loc, torch.softmax(cls, -1), ldm, shapes, (H, W), 0.5, 0.4)
    for i in range(len(frames)):
        w = want[i].numpy()
        assert counts[i] == len(w) and counts[i] > 0
        np.testing.assert_allclose(rows[i, :counts[i], :15], w, atol=1e-3, rtol=0)


def test_mnet_facade_end_to_end(gold, sd_np):
    """Retinaface(backbone='mobilenet') through detect_images (fp16 storage) against the executed reference's rows:
    boxes away from the threshold are found with IoU >= 0.95."""
    from detect_face.face import Retinaface
    frames = _frames(gold, "vga")
    det = Retinaface(state_dict=sd_np, backbone="mobilenet", input_shape=[480, 640, 3], confidence=0.5, nms_iou=0.4,
                     letterbox_image=True, cuda=True, precision="fp16")
    assert det.cfg["name"] == "mobilenet0.25" and det.anchors.shape == (am.num_anchors(480, 640), 4)
    res = det.detect_images(frames)
    want = gold["vga.rows0"]
    got = np.asarray(res[0][1], np.float32).reshape(-1, 4)
    scale = np.array([640, 480, 640, 480], np.float32)
    strong = want[np.abs(want[:, 4] - 0.5) > 0.05]

    def iou(a, b):
        x1, y1 = np.maximum(a[0], b[:, 0]), np.maximum(a[1], b[:, 1])
        x2, y2 = np.minimum(a[2], b[:, 2]), np.minimum(a[3], b[:, 3])
        inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
        return inter / ((a[2] - a[0]) * (a[3] - a[1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter)

    best = np.array([iou(r[:4] * scale, got).max(initial=0) for r in strong])
    print(f"ref {len(want)} got {len(got)} strong {len(strong)} matched95 {(best >= 0.95).sum()}")
    assert len(got) > 0 and (best < 0.95).sum() <= max(2, len(strong) // 10)
    # the network handle mirrors RetinaFace.forward (reference retinaface.py:114-148)
    x, _ = det.preprocess(frames)
    loc, conf, ldm = det.net(x)
    assert float((conf.cpu() - torch.from_numpy(gold["vga.conf"])).abs().max()) <= 0.02
