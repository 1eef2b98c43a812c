"""Functional torch-tensor front end of the C-ABI (one call = one libvdesens entry point, enqueued on
torch's current CUDA stream).  The engine drives the same entry points through pre-built plans; this
module is the stateless surface used by the façade for one-off calls and by the parity tests.
All functions raise RuntimeError on failure and require CUDA tensors — there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import FRAME_DTYPE, MEAN


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_arg_device(fn):
    """run with the CUDA device of the first tensor argument current: the C-ABI launches on the current device and
    on its current stream, and a process may hold tensors on several GPUs"""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        first = args[0] if args else next(iter(kwargs.values()), None)
        if isinstance(first, (list, tuple)):
            first = first[0] if first else None
        if not isinstance(first, torch.Tensor):
            return fn(*args, **kwargs)              # nothing to take a device from: the function's own checks apply
        assert first.is_cuda, "libvdesens ops need CUDA tensors (no CPU fallback)"
        with torch.cuda.device(first.device):
            return fn(*args, **kwargs)
    return wrapped


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "libvdesens ops need CUDA tensors (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def frame_descriptors(frames):
    """list of device uint8 (h,w,3) tensors -> device int32 (B,6) tensor of vd_frame_t records."""
    host = np.zeros((len(frames),), FRAME_DTYPE)
    for i, f in enumerate(frames):
        assert f.is_cuda and f.dtype == torch.uint8 and f.dim() == 3 and f.shape[2] == 3
        assert f.stride(2) == 1 and f.stride(1) == 3
        host[i] = (f.data_ptr(), f.shape[0], f.shape[1], f.stride(0), 0)
    return torch.from_numpy(host.view(np.int32).reshape(len(frames), 6)).to(frames[0].device)


def x0_pitch(W):
    return int(_lib.lib().vd_x0_pitch(int(W)))


@_on_arg_device
def ingest(frames, H, W, identity=None, precision="bf16", copy_to=None):
    """-> x0 (B,H,Wp,4) in the library's activation type.  reference: Retinaface.preprocess (face.py:65-88).
    copy_to: list of device uint8 (H,W,3) tensors that receive a copy of the frames in the same pass."""
    lib = _lib.lib(precision)
    B = len(frames)
    desc = frame_descriptors(frames)
    x0 = torch.zeros((B, H, x0_pitch(W), 4), dtype=_lib.torch_dtype(precision), device=frames[0].device)
    if identity is None:
        identity = all(f.shape[0] == H and f.shape[1] == W for f in frames)
    if copy_to is not None:
        assert identity and len(copy_to) == B
        x0.fill_(float("nan"))            # the kernel writes every column, pad columns included
        _lib.check(lib.vd_ingest_u8hwc_identity_copy(_p(desc), _p(frame_descriptors(copy_to)), B, _p(x0), H, W,
                                                     (C.c_float * 3)(*MEAN), _stream()), "vd_ingest_u8hwc_identity_copy", lib)
        return x0
    fn = lib.vd_ingest_u8hwc_identity if identity else lib.vd_ingest_u8hwc
    _lib.check(fn(_p(desc), B, _p(x0), H, W, (C.c_float * 3)(*MEAN), _stream()), "vd_ingest_u8hwc", lib)
    return x0


@_on_arg_device
def ingest_nchw_f32(frames, H, W):
    """-> fp32 (B,3,H,W), the tensor Retinaface.preprocess returns (face.py:87)."""
    lib = _lib.lib()
    B = len(frames)
    desc = frame_descriptors(frames)
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=frames[0].device)
    _lib.check(lib.vd_ingest_u8hwc_nchw_f32(_p(desc), B, _p(out), H, W, (C.c_float * 3)(*MEAN), _stream()), "vd_ingest_u8hwc_nchw_f32", lib)
    return out


@_on_arg_device
def conv2d(x, w_packed, bias, k, stride, pad, relu=False, residual=None, upadd=None, out=None, dst_coff=0,
           out_f32=False, x0_hw=None, x2=None, stride2=1, split=False, pool=False, slope=None, chain=None, reverse=False,
           acc_scale=0.0, two_dst=None):
    """x: bf16 NHWC (B,H,W,Cin) — or the ingest tensor with x0_hw=(H,W) for conv1.
    w_packed: same dtype (Cout, k*k*Cin) K order (r,s,cin) (weights.pack_all); bias fp32 (Cout).
    The activation dtype (torch.bfloat16 / torch.float16) selects libvdesens.so / libvdesens_fp16.so."""
    lib = _lib.lib(_lib.precision_of(x.dtype))
    assert w_packed.dtype == x.dtype
    B = x.shape[0]
    cout = w_packed.shape[0]
    d = _lib.VdConvDesc()
    flags = 0
    if x0_hw is not None:
        hin, win, cin = x0_hw[0], x0_hw[1], 3
        flags |= _lib.VD_CONV_IN_X0
    else:
        hin, win, cin = x.shape[1], x.shape[2], x.shape[3]
    hout = (hin + 2 * pad - k) // stride + 1
    wout = (win + 2 * pad - k) // stride + 1
    if out is None:
        oc = cout if (out_f32 or not split) else 3 * cout      # split tensors carry [hi | lo | hi] planes
        oh, ow = ((hout + 1) // 2, (wout + 1) // 2) if pool else (hout, wout)   # pool: conv1 + maxpool 3x3/2 fused
        out = torch.empty((B, oh, ow, oc), dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
    if pool:
        flags |= _lib.VD_CONV_POOL
    if split:
        flags |= _lib.VD_CONV_SPLIT
    d.B, d.Hin, d.Win, d.Cin = B, hin, win, cin
    d.Hout, d.Wout, d.Cout = hout, wout, cout
    d.R = d.S = k
    d.stride, d.pad = stride, pad
    d.dst_C, d.dst_coff = out.shape[3], dst_coff
    if relu:
        flags |= _lib.VD_CONV_RELU
    if reverse:               # scheduling hint only: tiles are walked from the last to the first (same results)
        flags |= _lib.VD_CONV_REVERSE
    if residual is not None:
        flags |= _lib.VD_CONV_RESIDUAL
    if upadd is not None:
        flags |= _lib.VD_CONV_UPADD
        d.up_H, d.up_W = upadd.shape[1], upadd.shape[2]
    if out_f32:
        flags |= _lib.VD_CONV_OUT_F32
    if x2 is not None:        # fused 1x1 projection of a second tensor (weights concatenated along K)
        flags |= _lib.VD_CONV_DUAL
        d.x2 = x2.data_ptr()
        d.Hin2, d.Win2, d.Cin2, d.stride2 = x2.shape[1], x2.shape[2], x2.shape[3], stride2
    if slope is not None:     # per-channel negative-side slope instead of relu (0 ReLU, 0.1 LeakyReLU, 1 identity)
        flags |= _lib.VD_CONV_SLOPE
        d.slope = slope.data_ptr()
    out2 = None
    if chain is not None:     # (w2 [N2][Cout], bias2 [N2]): also returns relu(conv1x1(out) + bias2) from the same kernel
        w2, b2 = chain
        out2 = torch.empty((B, hout, wout, w2.shape[0]), dtype=x.dtype, device=x.device)
        flags |= _lib.VD_CONV_CHAIN
        d.chain_w, d.chain_bias, d.chain_y, d.chain_Cout = w2.data_ptr(), b2.data_ptr(), out2.data_ptr(), w2.shape[0]
    if two_dst is not None:   # (cout1, out_b, dst_b_coff): channels >= cout1 are written to out_b[..., dst_b_coff:]
        cout1, out_b, coff_b = two_dst
        flags |= _lib.VD_CONV_TWO_DST
        d.Cout1, d.y_b, d.dst_b_C, d.dst_b_coff = int(cout1), out_b.data_ptr(), out_b.shape[3], int(coff_b)
    d.acc_scale = float(acc_scale)       # split mode: accumulators * acc_scale before the bias (weights packed as w / acc_scale)
    d.flags = flags
    d.x, d.w, d.bias = x.data_ptr(), w_packed.data_ptr(), bias.data_ptr()
    d.residual = residual.data_ptr() if residual is not None else None
    d.upadd = upadd.data_ptr() if upadd is not None else None
    d.y = out.data_ptr()
    h = C.c_void_p()
    _lib.check(lib.vd_conv_plan_create(C.byref(d), C.byref(h)), "vd_conv_plan_create", lib)
    try:
        _lib.check(lib.vd_conv_run(h, _stream()), "vd_conv_run", lib)
    finally:
        lib.vd_conv_plan_destroy(h)
    return out if out2 is None else (out, out2)


@_on_arg_device
def mnet_stem(x0, w, bias, slope, H, W):
    """conv_bn(3, 8, stride 2) of the MobileNet0.25 variant on the ingest tensor (reference mobilenet025.py:3-8)."""
    lib = _lib.lib(_lib.precision_of(x0.dtype))
    y = torch.empty((x0.shape[0], (H + 1) // 2, (W + 1) // 2, 8), dtype=x0.dtype, device=x0.device)
    _lib.check(lib.vd_mnet_stem(_p(x0), _p(w), _p(bias), _p(slope), _p(y), x0.shape[0], H, W, _stream()), "vd_mnet_stem", lib)
    return y


@_on_arg_device
def mnet_dw3x3(x, w, bias, slope, stride):
    """depthwise 3x3 + folded BN + slope activation (reference mobilenet025.py:10-14); w is [9][C]."""
    lib = _lib.lib(_lib.precision_of(x.dtype))
    B, H, W, Cc = x.shape
    y = torch.empty((B, (H + stride - 1) // stride, (W + stride - 1) // stride, Cc), dtype=x.dtype, device=x.device)
    _lib.check(lib.vd_mnet_dw3x3(_p(x), _p(w), _p(bias), _p(slope), _p(y), B, H, W, Cc, stride, _stream()), "vd_mnet_dw3x3", lib)
    return y


@_on_arg_device
def mnet_dwpw(x, wdw, bdw, sdw, wpw, bpw, spw, stride):
    """conv_dw block (depthwise 3x3 + pointwise 1x1, reference mobilenet025.py:10-19) as one kernel."""
    lib = _lib.lib(_lib.precision_of(x.dtype))
    B, H, W, Cin = x.shape
    Cout = wpw.shape[1]
    y = torch.empty((B, (H + stride - 1) // stride, (W + stride - 1) // stride, Cout), dtype=x.dtype, device=x.device)
    _lib.check(lib.vd_mnet_dwpw(_p(x), _p(wdw), _p(bdw), _p(sdw), _p(wpw), _p(bpw), _p(spw), _p(y), B, H, W, Cin, Cout,
                                stride, _stream()), "vd_mnet_dwpw", lib)
    return y


@_on_arg_device
def mnet_pw(x, w, bias, slope):
    """pointwise 1x1 for Cin < 64 (reference mobilenet025.py:16-18); w is [Cin][Cout]."""
    lib = _lib.lib(_lib.precision_of(x.dtype))
    B, H, W, Cin = x.shape
    Cout = w.shape[1]
    y = torch.empty((B, H, W, Cout), dtype=x.dtype, device=x.device)
    _lib.check(lib.vd_mnet_pw(_p(x), _p(w), _p(bias), _p(slope), _p(y), B * H * W, Cin, Cout, _stream()), "vd_mnet_pw", lib)
    return y


@_on_arg_device
def maxpool3x3s2(x):
    lib = _lib.lib(_lib.precision_of(x.dtype))
    B, H, W, Cc = x.shape
    y = torch.empty((B, (H + 1) // 2, (W + 1) // 2, Cc), dtype=x.dtype, device=x.device)
    _lib.check(lib.vd_maxpool3x3s2(_p(x), _p(y), B, H, W, Cc, _stream()), "vd_maxpool3x3s2", lib)
    return y


@_on_arg_device
def decode_threshold_lcl(loc, conf, landms, H, W, conf_thr, cap=4096):
    """reference-layout head tensors -> (cand (B,cap,16), count (B,), status (1,))"""
    lib = _lib.lib()
    B = loc.shape[0]
    cand = torch.zeros((B, cap, 16), dtype=torch.float32, device=loc.device)
    count = torch.zeros((B,), dtype=torch.int32, device=loc.device)
    status = torch.zeros((1,), dtype=torch.int32, device=loc.device)
    _lib.check(lib.vd_decode_threshold_lcl(_p(loc.contiguous()), _p(conf.contiguous()), _p(landms.contiguous()), B,
                                           H, W, float(conf_thr), cap, _p(cand), _p(count), _p(status), _stream()), "vd_decode_threshold_lcl", lib)
    return cand, count, status


@_on_arg_device
def decode_threshold_heads(heads, H, W, conf_thr, cap=4096):
    lib = _lib.lib()
    B = heads[0].shape[0]
    dev = heads[0].device
    cand = torch.zeros((B, cap, 16), dtype=torch.float32, device=dev)
    count = torch.zeros((B,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    _lib.check(lib.vd_decode_threshold(_p(heads[0]), _p(heads[1]), _p(heads[2]), B, H, W, float(conf_thr), cap,
                                       _p(cand), _p(count), _p(status), _stream()), "vd_decode_threshold", lib)
    return cand, count, status


def unletterbox_params(input_hw, image_shapes):
    """(B,4) fp32 [offx, offy, scalex, scaley], fp32 arithmetic of utils_bbox.py:26-31."""
    inp = np.array([input_hw[0], input_hw[1]], np.float32)
    out = np.zeros((len(image_shapes), 4), np.float32)
    for i, (h, w) in enumerate(image_shapes):
        ish = np.array([h, w], np.float32)
        new_shape = (ish * np.min(inp / ish)).astype(np.float32)
        offset = ((inp - new_shape) / np.float32(2.0) / inp).astype(np.float32)
        scale = (inp / new_shape).astype(np.float32)
        out[i] = (offset[1], offset[0], scale[1], scale[0])
    return out


@_on_arg_device
def nms(cand, count, iou_thr, unlb, max_det=1024):
    """-> (rows (B,max_det,16), out_count (B,), status (1,))"""
    lib = _lib.lib()
    B, cap, _ = cand.shape
    dev = cand.device
    rows = torch.zeros((B, max_det, 16), dtype=torch.float32, device=dev)
    out_count = torch.zeros((B,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    nbytes = int(lib.vd_nms_workspace_bytes(B, cap))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    unlb_t = torch.as_tensor(np.asarray(unlb, np.float32)).to(dev)
    _lib.check(lib.vd_nms_batched(_p(cand), _p(count), B, cap, float(iou_thr), _p(unlb_t), max_det, _p(rows),
                                  _p(out_count), _p(status), _p(ws), nbytes, _stream()), "vd_nms_batched", lib)
    return rows, out_count, status


@_on_arg_device
def composite(frames, ext_boxes, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0, det_rows=None,
              det_count=None, scratch_frames=8.0):
    """In-place sequential compositing on device frames.  frames: list of device uint8 (h,w,3);
    ext_boxes: list (per frame) of (x1,y1,x2,y2) int boxes applied after the optional device-side
    detections (det_rows (B,max_det,16) normalised rows + det_count (B,))."""
    lib = _lib.lib()
    B = len(frames)
    dev = frames[0].device
    desc = frame_descriptors(frames)
    max_ext = max(1, max((len(b) for b in ext_boxes), default=1))
    eb = np.zeros((B, max_ext, 4), np.int32)
    ec = np.zeros((B,), np.int32)
    for i, bl in enumerate(ext_boxes):
        ec[i] = len(bl)
        for j, b in enumerate(bl):
            eb[i, j] = [int(v) for v in b]
    eb_t, ec_t = torch.from_numpy(eb).to(dev), torch.from_numpy(ec).to(dev)
    max_det = det_rows.shape[1] if det_rows is not None else 0
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    px = int(scratch_frames * max(f.shape[0] * f.shape[1] for f in frames))
    nbytes = int(lib.vd_composite_workspace_bytes(B, max_det + max_ext, px))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    m = _lib.mode_id(mode)
    taps = None
    if m == _lib.VD_MODE_GAUSS:
        taps = (C.c_uint16 * 256)()
        _lib.check(lib.vd_gaussian_taps_q8(int(ksize), float(sigma), taps), "vd_gaussian_taps_q8", lib)
    _lib.check(lib.vd_composite_rois(_p(desc), B, _p(det_rows), _p(det_count), max_det, _p(eb_t), _p(ec_t), max_ext, m,
                                     int(mosaic_level), taps, int(ksize), _p(status), _p(ws), nbytes, _stream()), "vd_composite_rois", lib)
    torch.cuda.current_stream().synchronize()       # ws / desc are freed on return
    st = int(status.item())
    if st:
        raise RuntimeError(f"libvdesens composite device status {st}")
    return frames
