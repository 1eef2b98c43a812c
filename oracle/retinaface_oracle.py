"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32 / numpy) of the reference's detector path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this file; the product path (video-desensitization_b200/) never does.

Parity status: the reference ships no tests or golden vectors.  This restatement is
pinned instead against the *unmodified reference executed in the build container*:
tests/test_oracle_vs_reference.py (bit-exact on CPU, same torch build) and the committed outputs in
tests/golden/ written by oracle/make_golden.py.  The arithmetic that lives in third-party code
(ATen conv/BN, torchvision.ops.nms) is "parity unpinned by the reference" and is pinned by those
executed outputs.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

VARIANCE = (0.1, 0.2)                       # detect_face/utils/config.py:22
MIN_SIZES = ((16, 32), (64, 128), (256, 512))   # config.py:20
STEPS = (8, 16, 32)                         # config.py:21
BN_EPS = 1e-5
MEAN = (104.0, 117.0, 123.0)                # detect_face/utils/utils.py:28


def to_torch_sd(sd):
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)))
            for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# network: detect_face/retinaface.py:114-148, detect_face/nets/layers.py:10-114, torchvision resnet50
# ----------------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, sd, calibrate=False, capture=None):
        self.sd, self.calibrate, self.capture = sd, calibrate, capture

    def conv_bn(self, x, wkey, bnkey, stride, pad, relu):
        y = F.conv2d(x, self.sd[wkey], None, stride, pad)
        if self.calibrate:
            # eval-mode statistics := statistics of this layer's own pre-BN output (float64)
            yd = y.double()
            self.sd[bnkey + ".running_mean"] = yd.mean(dim=(0, 2, 3)).float()
            self.sd[bnkey + ".running_var"] = yd.var(dim=(0, 2, 3), unbiased=False).float().clamp_min(1e-6)
        y = F.batch_norm(y, self.sd[bnkey + ".running_mean"], self.sd[bnkey + ".running_var"],
                         self.sd[bnkey + ".weight"], self.sd[bnkey + ".bias"], False, 0.0, BN_EPS)
        if relu:
            y = F.relu(y)
        return y


def _bottleneck(c, x, p, stride, has_ds):
    # torchvision Bottleneck.forward (v1.5: stride on conv2)
    out = c.conv_bn(x, p + ".conv1.weight", p + ".bn1", 1, 0, True)
    out = c.conv_bn(out, p + ".conv2.weight", p + ".bn2", stride, 1, True)
    out = c.conv_bn(out, p + ".conv3.weight", p + ".bn3", 1, 0, False)
    idn = c.conv_bn(x, p + ".downsample.0.weight", p + ".downsample.1", stride, 0, False) if has_ds else x
    return F.relu(out + idn)


def _ssh(c, x, p):
    # detect_face/nets/layers.py:54-66
    a = c.conv_bn(x, p + ".conv3X3.0.weight", p + ".conv3X3.1", 1, 1, False)
    t = c.conv_bn(x, p + ".conv5X5_1.0.weight", p + ".conv5X5_1.1", 1, 1, True)
    b = c.conv_bn(t, p + ".conv5X5_2.0.weight", p + ".conv5X5_2.1", 1, 1, False)
    u = c.conv_bn(t, p + ".conv7X7_2.0.weight", p + ".conv7X7_2.1", 1, 1, True)
    d = c.conv_bn(u, p + ".conv7x7_3.0.weight", p + ".conv7x7_3.1", 1, 1, False)
    return F.relu(torch.cat([a, b, d], dim=1))


def forward(sd, x, calibrate=False, capture=None, softmax=True):
    """x: (B,3,H,W) fp32 mean-subtracted.  Returns (loc (B,A,4), conf (B,A,2), landms (B,A,10)).
    With `capture` (a dict) intermediate NCHW tensors are stored under the engine's buffer names."""
    c = _Ctx(sd, calibrate, capture)
    cap = capture if capture is not None else {}
    with torch.no_grad():
        y = c.conv_bn(x, "body.conv1.weight", "body.bn1", 2, 3, True)
        cap["c1"] = y
        y = F.max_pool2d(y, 3, 2, 1)
        cap["p1"] = y
        feats = []
        for lname, nblocks, lstride in (("layer1", 3, 1), ("layer2", 4, 2), ("layer3", 6, 2), ("layer4", 3, 2)):
            for b in range(nblocks):
                y = _bottleneck(c, y, f"body.{lname}.{b}", lstride if b == 0 else 1, b == 0)
                cap[f"{lname}.{b}.out"] = y
            if lname != "layer1":
                feats.append(y)
        # FPN, layers.py:79-114
        o1 = c.conv_bn(feats[0], "fpn.output1.0.weight", "fpn.output1.1", 1, 0, True)
        o2 = c.conv_bn(feats[1], "fpn.output2.0.weight", "fpn.output2.1", 1, 0, True)
        o3 = c.conv_bn(feats[2], "fpn.output3.0.weight", "fpn.output3.1", 1, 0, True)
        o2 = o2 + F.interpolate(o3, size=[o2.size(2), o2.size(3)], mode="nearest")
        cap["fpn.o2s"] = o2
        o2 = c.conv_bn(o2, "fpn.merge2.0.weight", "fpn.merge2.1", 1, 1, True)
        o1 = o1 + F.interpolate(o2, size=[o1.size(2), o1.size(3)], mode="nearest")
        cap["fpn.o1s"] = o1
        o1 = c.conv_bn(o1, "fpn.merge1.0.weight", "fpn.merge1.1", 1, 1, True)
        cap["fpn.o1"], cap["fpn.o2"], cap["fpn.o3"] = o1, o2, o3
        fs = [_ssh(c, o1, "ssh1"), _ssh(c, o2, "ssh2"), _ssh(c, o3, "ssh3")]
        for i, f in enumerate(fs):
            cap[f"ssh{i + 1}.out"] = f

        def head(prefix, n):
            outs = []
            for i, f in enumerate(fs):
                o = F.conv2d(f, sd[f"{prefix}.{i}.conv1x1.weight"], sd[f"{prefix}.{i}.conv1x1.bias"])
                outs.append(o.permute(0, 2, 3, 1).contiguous().view(o.shape[0], -1, n))
            return torch.cat(outs, dim=1)

        loc, cls, ldm = head("BboxHead", 4), head("ClassHead", 2), head("LandmarkHead", 10)
        if softmax:
            cls = F.softmax(cls, dim=-1)       # retinaface.py:147
    return loc, cls, ldm


# ----------------------------------------------------------------------------------------------
# MobileNet0.25 variant: detect_face/nets/mobilenet025.py:3-49 (conv_bn / conv_dw, LeakyReLU 0.1),
# FPN / SSH with leaky = 0.1 because out_channel = 64 <= 64 (layers.py:41,71), heads on 64 channels
# ----------------------------------------------------------------------------------------------
_MNET_BLOCKS = (("stage1", 1, 1), ("stage1", 2, 2), ("stage1", 3, 1), ("stage1", 4, 2), ("stage1", 5, 1),
                ("stage2", 0, 2), ("stage2", 1, 1), ("stage2", 2, 1), ("stage2", 3, 1), ("stage2", 4, 1), ("stage2", 5, 1),
                ("stage3", 0, 2), ("stage3", 1, 1))


def forward_mnet(sd, x, calibrate=False, capture=None, softmax=True):
    """x: (B,3,H,W) fp32 mean-subtracted.  Returns (loc (B,A,4), conf (B,A,2), landms (B,A,10))."""
    cap = capture if capture is not None else {}

    def cba(x, wkey, bnkey, stride, pad, slope, groups=1):
        y = F.conv2d(x, sd[wkey], None, stride, pad, 1, groups)
        if calibrate:
            yd = y.double()
            sd[bnkey + ".running_mean"] = yd.mean(dim=(0, 2, 3)).float()
            sd[bnkey + ".running_var"] = yd.var(dim=(0, 2, 3), unbiased=False).float().clamp_min(1e-6)
        y = F.batch_norm(y, sd[bnkey + ".running_mean"], sd[bnkey + ".running_var"], sd[bnkey + ".weight"],
                         sd[bnkey + ".bias"], False, 0.0, BN_EPS)
        return y if slope is None else F.leaky_relu(y, slope)

    with torch.no_grad():
        y = cba(x, "body.stage1.0.0.weight", "body.stage1.0.1", 2, 1, 0.1)                 # mobilenet025.py:3-8
        cap["m.stem"] = y
        feats = {}
        for st, i, stride in _MNET_BLOCKS:                                                 # conv_dw, :10-19
            p = f"body.{st}.{i}"
            y = cba(y, p + ".0.weight", p + ".1", stride, 1, 0.1, groups=y.shape[1])
            cap[p + ".dw"] = y
            y = cba(y, p + ".3.weight", p + ".4", 1, 0, 0.1)
            cap[p + ".pw"] = y
            feats[st] = y
        o1 = cba(feats["stage1"], "fpn.output1.0.weight", "fpn.output1.1", 1, 0, 0.1)     # layers.py:94-96
        o2 = cba(feats["stage2"], "fpn.output2.0.weight", "fpn.output2.1", 1, 0, 0.1)
        o3 = cba(feats["stage3"], "fpn.output3.0.weight", "fpn.output3.1", 1, 0, 0.1)
        o2 = o2 + F.interpolate(o3, size=[o2.size(2), o2.size(3)], mode="nearest")
        o2 = cba(o2, "fpn.merge2.0.weight", "fpn.merge2.1", 1, 1, 0.1)
        o1 = o1 + F.interpolate(o2, size=[o1.size(2), o1.size(3)], mode="nearest")
        o1 = cba(o1, "fpn.merge1.0.weight", "fpn.merge1.1", 1, 1, 0.1)
        cap["fpn.o1"], cap["fpn.o2"], cap["fpn.o3"] = o1, o2, o3
        fs = []
        for lvl, o in enumerate((o1, o2, o3)):                                             # layers.py:54-66
            p = f"ssh{lvl + 1}"
            a = cba(o, p + ".conv3X3.0.weight", p + ".conv3X3.1", 1, 1, None)
            t = cba(o, p + ".conv5X5_1.0.weight", p + ".conv5X5_1.1", 1, 1, 0.1)
            b = cba(t, p + ".conv5X5_2.0.weight", p + ".conv5X5_2.1", 1, 1, None)
            u = cba(t, p + ".conv7X7_2.0.weight", p + ".conv7X7_2.1", 1, 1, 0.1)
            d = cba(u, p + ".conv7x7_3.0.weight", p + ".conv7x7_3.1", 1, 1, None)
            fs.append(F.relu(torch.cat([a, b, d], dim=1)))
            cap[p + ".out"] = fs[-1]

        def head(prefix, n):
            outs = []
            for i, f in enumerate(fs):
                o = F.conv2d(f, sd[f"{prefix}.{i}.conv1x1.weight"], sd[f"{prefix}.{i}.conv1x1.bias"])
                outs.append(o.permute(0, 2, 3, 1).contiguous().view(o.shape[0], -1, n))
            return torch.cat(outs, dim=1)

        loc, cls, ldm = head("BboxHead", 4), head("ClassHead", 2), head("LandmarkHead", 10)
        if softmax:
            cls = F.softmax(cls, dim=-1)
    return loc, cls, ldm


def forward_mnet_emulated(packed, x, capture=None, dtype=torch.bfloat16):
    """The device op list of vdesens_b200.arch_mnet executed on the CPU with the SAME packed weights and the same
    rounding points (activations rounded to `dtype` where the device stores them, fp32 math in between)."""
    from vdesens_b200 import arch_mnet as am
    ops, _ = am.build_graph()
    bufs = {"x0": x}
    heads = [None, None, None]

    def bf(t):
        return t.to(dtype).float()

    def act(y, slope):
        return torch.where(y > 0, y, y * slope.view(1, -1, 1, 1))

    with torch.no_grad():
        for op in ops:
            w, b, slope = packed[op.name]
            w = w.float()
            xin = bufs[op.src]
            if op.kind == "stem":
                wt = w.view(3, 3, 4, 8)[:, :, :3, :].permute(3, 2, 0, 1).contiguous()
                y = act(F.conv2d(xin, wt, b, 2, 1), slope)
            elif op.kind == "dw":
                wt = w.view(3, 3, op.cin).permute(2, 0, 1).unsqueeze(1).contiguous()
                y = act(F.conv2d(xin, wt, b, op.stride, 1, 1, op.cin), slope)
            elif op.kind == "pw":
                y = act(F.conv2d(xin, w.t().contiguous().view(op.cout, op.cin, 1, 1), b), slope)
            elif op.kind == "conv":
                wt = w.view(op.cout, op.k, op.k, op.cin).permute(0, 3, 1, 2).contiguous()
                y = act(F.conv2d(xin, wt, b, op.stride, op.pad), slope)
                if op.upsample_add:
                    y = y + F.interpolate(bufs[op.upsample_add], size=y.shape[2:], mode="nearest")
            elif op.kind == "heads":
                xx = torch.cat([xin, bufs[op.src2]], dim=1)
                heads[op.level] = F.conv2d(xx, w.view(32, 128, 1, 1), b)
                continue
            bufs[op.dst] = bf(y)
        if capture is not None:
            capture.update(bufs)
        locs, clss, ldms = [], [], []
        for h in heads:
            hh = h.permute(0, 2, 3, 1).contiguous()
            B = hh.shape[0]
            locs.append(hh[..., 0:8].reshape(B, -1, 4))
            clss.append(hh[..., 8:12].reshape(B, -1, 2))
            ldms.append(hh[..., 12:32].reshape(B, -1, 10))
    return torch.cat(locs, 1), torch.cat(clss, 1), torch.cat(ldms, 1)


# ----------------------------------------------------------------------------------------------
# anchors: detect_face/utils/anchors.py:7-41  (python doubles, rounded once to fp32 by torch.Tensor)
# ----------------------------------------------------------------------------------------------
def anchors(h, w):
    out = []
    for k, step in enumerate(STEPS):
        fh, fw = math.ceil(h / step), math.ceil(w / step)
        ii, jj = np.meshgrid(np.arange(fh), np.arange(fw), indexing="ij")
        cx = (jj.astype(np.float64) + 0.5) * step / w
        cy = (ii.astype(np.float64) + 0.5) * step / h
        per = []
        for ms in MIN_SIZES[k]:
            per.append(np.stack([cx, cy, np.full_like(cx, ms / w), np.full_like(cy, ms / h)], axis=-1))
        out.append(np.stack(per, axis=2).reshape(-1, 4))
    return torch.from_numpy(np.concatenate(out, 0).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# decode: detect_face/utils/utils_bbox.py:49-79 (evaluation order kept: x2 = w + x1)
# ----------------------------------------------------------------------------------------------
def decode(loc, priors):
    cxcy = priors[..., :2] + loc[..., :2] * VARIANCE[0] * priors[..., 2:]
    wh = priors[..., 2:] * torch.exp(loc[..., 2:] * VARIANCE[1])
    x1y1 = cxcy - wh / 2
    x2y2 = wh + x1y1
    return torch.cat([x1y1, x2y2], dim=-1)


def decode_landm(landm, priors):
    parts = [priors[..., :2] + landm[..., 2 * k:2 * k + 2] * VARIANCE[0] * priors[..., 2:] for k in range(5)]
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------------------------
# NMS: torchvision.ops.nms CPU kernel semantics (utils_bbox.py:103-130 calls batched_nms with one
# class): stable descending sort by score, greedy, suppress when IoU > thr, areas without +1.
# ----------------------------------------------------------------------------------------------
def nms_keep(boxes, scores, iou_thr):
    """boxes (N,4) fp32 numpy, scores (N,) fp32 -> kept indices (into the input), score-descending."""
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(len(boxes), bool)
    keep = []
    thr = np.float32(iou_thr)
    for _i, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(int(i))
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1); hgt = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * hgt
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def non_max_suppression(detection, conf_thres, nms_thres):
    """detection (B,A,15) torch -> list of (M_i,15); threshold is >= (utils_bbox.py:116)."""
    out = []
    for b in range(detection.shape[0]):
        d = detection[b]
        d = d[d[:, 4] >= conf_thres]
        if len(d) == 0:
            out.append(torch.empty((0, 15)))
            continue
        keep = nms_keep(d[:, :4].numpy(), d[:, 4].numpy(), nms_thres)
        out.append(d[torch.from_numpy(keep)])
    return out


# ----------------------------------------------------------------------------------------------
# un-letterbox: detect_face/utils/utils_bbox.py:12-43 (fp32 tensor arithmetic, un-truncated floats)
# ----------------------------------------------------------------------------------------------
def correct_boxes(results, input_hw, image_shapes):
    inp = torch.tensor([float(input_hw[0]), float(input_hw[1])], dtype=torch.float32)
    out = []
    for i, r in enumerate(results):
        if len(r) == 0:
            out.append(r)
            continue
        ish = torch.tensor([float(image_shapes[i][0]), float(image_shapes[i][1])], dtype=torch.float32)
        new_shape = ish * torch.min(inp / ish)
        offset = (inp - new_shape) / 2. / inp
        scale = inp / new_shape
        sb = torch.stack([scale[1], scale[0], scale[1], scale[0]])
        ob = torch.stack([offset[1], offset[0], offset[1], offset[0]])
        r = r.clone()
        r[:, :4] = (r[:, :4] - ob) * sb
        r[:, 5:] = (r[:, 5:] - ob[:2].repeat(5)) * sb[:2].repeat(5)
        out.append(r)
    return out


def unletterbox_params(input_hw, image_hw):
    """(offx, offy, scalex, scaley) exactly as fp32 tensor arithmetic of utils_bbox.py:26-31."""
    inp = np.array([input_hw[0], input_hw[1]], np.float32)
    ish = np.array([image_hw[0], image_hw[1]], np.float32)
    new_shape = (ish * np.min(inp / ish)).astype(np.float32)
    offset = ((inp - new_shape) / np.float32(2.) / inp).astype(np.float32)
    scale = (inp / new_shape).astype(np.float32)
    return offset[1], offset[0], scale[1], scale[0]


# ----------------------------------------------------------------------------------------------
# ingest: detect_face/utils/utils.py:8-29 + detect_face/face.py:65-88
# ----------------------------------------------------------------------------------------------
def preprocess(images, input_hw):
    """list of HxWx3 uint8 -> (B,3,H,W) fp32, (B,2) [h,w] fp32.  Letterbox uses the integer-exact
    restatement in oracle/cv_oracle.py (cv2.resize INTER_LINEAR semantics)."""
    from . import cv_oracle
    H, W = input_hw
    outs, shapes = [], []
    for img in images:
        ih, iw = img.shape[:2]
        shapes.append([ih, iw])
        canvas = cv_oracle.letterbox(img, W, H).astype(np.float64)
        canvas -= np.array(MEAN, np.float32)
        outs.append(np.transpose(canvas, (2, 0, 1)))
    return torch.from_numpy(np.stack(outs, 0)).float(), torch.tensor(shapes, dtype=torch.float32)


def postprocess(loc, conf, landms, image_shapes, input_hw, confidence, nms_iou):
    """detect_face/face.py:93-115"""
    pri = anchors(input_hw[0], input_hw[1]).unsqueeze(0).expand(loc.shape[0], -1, -1)
    boxes = decode(loc, pri)
    lm = decode_landm(landms, pri)
    det = torch.cat([boxes, conf[:, :, 1:2], lm], dim=-1)
    res = non_max_suppression(det, confidence, nms_iou)
    return correct_boxes(res, input_hw, image_shapes)


def detect_images(sd, images, input_hw, confidence=0.5, nms_iou=0.45, return_rows=False):
    """detect_face/face.py:120-150.  Returns list of (image copy, [[x1,y1,x2,y2],...]) in original
    pixels, descending score; with return_rows also the (M,15) rows (normalised to the image)."""
    x, shapes = preprocess(images, input_hw)
    loc, conf, landms = forward(sd, x)
    results = postprocess(loc, conf, landms, shapes, input_hw, confidence, nms_iou)
    out = []
    for i, r in enumerate(results):
        boxes = []
        if len(r) > 0:
            rr = r.numpy().copy()
            h, w = shapes[i].numpy()
            rr[:, :4] *= np.array([w, h, w, h])
            boxes = rr[:, :4].tolist()
        out.append((images[i].copy(), boxes))
    return (out, results) if return_rows else out


# ----------------------------------------------------------------------------------------------
# bf16 emulation of the device network (same packed weights, same rounding points, fp32 math):
# separates "the CUDA pipeline computes the right function" from "bf16 differs from fp32".
# ----------------------------------------------------------------------------------------------
def forward_bf16_emulated(packed, x, capture=None, dtype=torch.bfloat16):
    """packed: vdesens_b200.weights.pack_all(sd) (BN-folded bf16 GEMM operands).  x: (B,3,H,W) fp32
    mean-subtracted (exact integers).  Activations are rounded to bf16 where the device stores them."""
    from vdesens_b200 import arch
    convs, ops, _ = arch.build_graph()
    cap = capture if capture is not None else {}

    def bf(t):
        return t.to(dtype).float()

    bufs = {"x0": x}
    heads = [None, None, None]
    with torch.no_grad():
        for op in ops:
            if op.kind == "maxpool":
                bufs[op.dst] = F.max_pool2d(bufs[op.src], 3, 2, 1)
            elif op.kind == "conv":
                c = op.conv
                wp, b = packed[c.name]
                if c.name == "body.conv1":
                    w = wp.float().view(c.cout, 7, 8, 4)[:, :, :7, :3].permute(0, 3, 1, 2).contiguous()
                    y = F.conv2d(bufs[op.src], w, b, c.stride, c.pad)
                elif c.wkey2:       # fused projection: both GEMMs share the fp32 accumulator on the device
                    k1 = c.k * c.k * c.cin
                    w = wp.float()[:, :k1].reshape(c.cout, c.k, c.k, c.cin).permute(0, 3, 1, 2).contiguous()
                    w2 = wp.float()[:, k1:].reshape(c.cout, c.cin2, 1, 1)
                    y = F.conv2d(bufs[op.src], w, b, c.stride, c.pad) + F.conv2d(bufs[op.src2], w2, None, c.stride2, 0)
                else:
                    w = wp.float().view(c.cout, c.k, c.k, c.cin).permute(0, 3, 1, 2).contiguous()
                    y = F.conv2d(bufs[op.src], w, b, c.stride, c.pad)
                if op.residual:
                    y = y + bufs[op.residual]
                if c.relu:
                    y = F.relu(y)
                if op.upsample_add:
                    y = y + F.interpolate(bufs[op.upsample_add], size=y.shape[2:], mode="nearest")
                y = bf(y)
                if op.dst in bufs and (op.dst_coff or bufs[op.dst].shape[1] != c.cout):
                    bufs[op.dst][:, op.dst_coff:op.dst_coff + c.cout] = y
                elif op.dst.endswith(".out") and op.dst.startswith("ssh"):
                    full = torch.zeros((y.shape[0], 256, y.shape[2], y.shape[3]))
                    full[:, op.dst_coff:op.dst_coff + c.cout] = y
                    bufs[op.dst] = full
                else:
                    bufs[op.dst] = y
            elif op.kind == "heads":
                wp, b = packed[f"heads.{op.level}"]
                heads[op.level] = F.conv2d(bufs[op.src], wp.float().view(32, 256, 1, 1), b)
        cap.update(bufs)
        locs, clss, ldms = [], [], []
        for h in heads:
            hh = h.permute(0, 2, 3, 1).contiguous()
            B = hh.shape[0]
            locs.append(hh[..., 0:8].reshape(B, -1, 4))
            clss.append(hh[..., 8:12].reshape(B, -1, 2))
            ldms.append(hh[..., 12:32].reshape(B, -1, 10))
    return torch.cat(locs, 1), torch.cat(clss, 1), torch.cat(ldms, 1)
