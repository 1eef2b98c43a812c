"""`torch.library` registration of the stateless C-ABI entry points (namespace ``vdesens``).

north_star: "a thin C-ABI layer exposed as torch custom ops".  Each op below is ONE libvdesens call
(include/vdesens.h) enqueued on torch's current CUDA stream; the implementations live in `ops.py` (ctypes →
`extern "C"`).  They are registered for the CUDA dispatch key only: a CPU tensor raises NotImplementedError
from the dispatcher — there is no CPU kernel to fall back to.  Fake (meta) kernels give the output shapes so
the ops can be traced / shape-checked without a GPU.

    import vdesens_b200.torch_ops            # registers torch.ops.vdesens.*
    x0 = torch.ops.vdesens.ingest(frames, 1080, 1920)
    y  = torch.ops.vdesens.conv2d(x, w_packed, bias, 3, 1, 1, True, None)

The engine (`engine.py`) drives the same entry points through pre-built plans on its private stream; these ops are
the composable surface for callers that hold torch tensors.
"""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops

NAMESPACE = "vdesens"
OP_NAMES = ("ingest", "ingest_nchw_f32", "conv2d", "maxpool3x3s2", "decode_threshold", "nms", "composite_")


# ---- K1: frame ingest (reference Retinaface.preprocess, detect_face/face.py:65-88) ---------------------------
@torch.library.custom_op(f"{NAMESPACE}::ingest", mutates_args=(), device_types="cuda")
def ingest(frames: List[Tensor], H: int, W: int) -> Tensor:
    """uint8 HWC frames -> mean-subtracted x0 (B,H,Wp,4) bf16 (letterboxed to H×W when the sizes differ)."""
    return ops.ingest(list(frames), H, W)


@ingest.register_fake
def _(frames, H, W):
    wp = _x0_pitch_fake(W)
    return frames[0].new_empty((len(frames), H, wp, 4), dtype=torch.bfloat16)


def _x0_pitch_fake(W: int) -> int:
    # the pitch rule is the library's (vd_x0_pitch, a host-only function: no GPU needed to ask)
    return int(_lib.lib().vd_x0_pitch(int(W)))


@torch.library.custom_op(f"{NAMESPACE}::ingest_nchw_f32", mutates_args=(), device_types="cuda")
def ingest_nchw_f32(frames: List[Tensor], H: int, W: int) -> Tensor:
    """the tensor Retinaface.preprocess returns (face.py:87): fp32 (B,3,H,W)."""
    return ops.ingest_nchw_f32(list(frames), H, W)


@ingest_nchw_f32.register_fake
def _(frames, H, W):
    return frames[0].new_empty((len(frames), 3, H, W), dtype=torch.float32)


# ---- K2: implicit-GEMM convolution (reference nets/layers.py:10-33, retinaface.py:13-51) ---------------------
@torch.library.custom_op(f"{NAMESPACE}::conv2d", mutates_args=(), device_types="cuda")
def conv2d(x: Tensor, w_packed: Tensor, bias: Tensor, k: int, stride: int, pad: int, relu: bool,
           residual: Optional[Tensor]) -> Tensor:
    """NHWC activation × packed (Cout, k·k·Cin) weights + fp32 bias (+ residual) (+ ReLU) on tcgen05."""
    return ops.conv2d(x, w_packed, bias, k, stride, pad, relu=relu, residual=residual)


@conv2d.register_fake
def _(x, w_packed, bias, k, stride, pad, relu, residual):
    B, H, W, _c = x.shape
    return x.new_empty((B, (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1, w_packed.shape[0]))


@torch.library.custom_op(f"{NAMESPACE}::maxpool3x3s2", mutates_args=(), device_types="cuda")
def maxpool3x3s2(x: Tensor) -> Tensor:
    """torchvision resnet50.maxpool (3×3, stride 2, pad 1) on NHWC."""
    return ops.maxpool3x3s2(x)


@maxpool3x3s2.register_fake
def _(x):
    B, H, W, Cc = x.shape
    return x.new_empty((B, (H + 1) // 2, (W + 1) // 2, Cc))


# ---- K3: priors + decode + threshold (reference utils/anchors.py:22-41, utils_bbox.py:49-79,112-116) ---------
@torch.library.custom_op(f"{NAMESPACE}::decode_threshold", mutates_args=(), device_types="cuda")
def decode_threshold(loc: Tensor, conf: Tensor, landms: Tensor, H: int, W: int, conf_thr: float,
                     cap: int) -> Tuple[Tensor, Tensor, Tensor]:
    """reference-layout heads (B,A,4)/(B,A,2)/(B,A,10) -> candidates (B,cap,16), counts (B,), status (1,)."""
    return ops.decode_threshold_lcl(loc, conf, landms, H, W, conf_thr, cap)


@decode_threshold.register_fake
def _(loc, conf, landms, H, W, conf_thr, cap):
    B = loc.shape[0]
    return (loc.new_empty((B, cap, 16), dtype=torch.float32), loc.new_empty((B,), dtype=torch.int32),
            loc.new_empty((1,), dtype=torch.int32))


# ---- K4: batched NMS + un-letterbox (reference utils_bbox.py:103-130, 12-43) ---------------------------------
@torch.library.custom_op(f"{NAMESPACE}::nms", mutates_args=(), device_types="cuda")
def nms(cand: Tensor, count: Tensor, iou_thr: float, unlb: Tensor, max_det: int) -> Tuple[Tensor, Tensor, Tensor]:
    """candidates -> kept rows in descending score order (B,max_det,16), counts (B,), status (1,).
    unlb: (B,4) fp32 [offx, offy, scalex, scaley] (ops.unletterbox_params)."""
    return ops.nms(cand, count, iou_thr, unlb.detach().cpu().numpy(), max_det)


@nms.register_fake
def _(cand, count, iou_thr, unlb, max_det):
    B = cand.shape[0]
    return (cand.new_empty((B, max_det, 16)), cand.new_empty((B,), dtype=torch.int32),
            cand.new_empty((1,), dtype=torch.int32))


# ---- K5: in-place compositing (reference combine_detect.py:138-161, 237-249) ---------------------------------
@torch.library.custom_op(f"{NAMESPACE}::composite_", mutates_args=("frames",), device_types="cuda")
def composite_(frames: List[Tensor], ext_boxes: Tensor, ext_count: Tensor, mode: int, mosaic_level: int,
               ksize: int, sigma: float, det_rows: Optional[Tensor], det_count: Optional[Tensor]) -> None:
    """frames: device uint8 (h,w,3), modified in place.  ext_boxes (B,E,4) int32 / ext_count (B,) int32 on any
    device (host copies are read); mode 0 = mosaic, 1 = Gaussian blur.  Detections (normalised rows from `nms`)
    are applied first, then the external boxes, each in list order — the reference's sequential semantics."""
    eb = ext_boxes.detach().cpu().numpy()
    ec = ext_count.detach().cpu().numpy()
    boxes = [[tuple(int(v) for v in eb[i, j]) for j in range(int(ec[i]))] for i in range(len(frames))]
    ops.composite(list(frames), boxes, mode="mosaic" if mode == _lib.VD_MODE_MOSAIC else "gauss",
                  mosaic_level=mosaic_level, ksize=ksize, sigma=sigma, det_rows=det_rows, det_count=det_count)


@composite_.register_fake
def _(frames, ext_boxes, ext_count, mode, mosaic_level, ksize, sigma, det_rows, det_count):
    return None


def registered():
    """names under torch.ops.vdesens that resolve (used by the CPU test: registration needs no GPU)."""
    return tuple(n for n in OP_NAMES if hasattr(getattr(torch.ops, NAMESPACE), n))
