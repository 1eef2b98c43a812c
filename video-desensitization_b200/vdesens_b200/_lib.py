"""ctypes binding of libvdesens.so (include/vdesens.h).  There is NO fallback: if the CUDA
library is missing or a call fails, a RuntimeError is raised — the product path never computes on
the CPU."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdesens.so")
LIB_PATHS = {"bf16": LIB_PATH, "fp16": os.path.join(_HERE, "libvdesens_fp16.so")}

VD_CONV_RELU, VD_CONV_RESIDUAL, VD_CONV_UPADD, VD_CONV_OUT_F32 = 1, 2, 4, 8
VD_CONV_IN_X0, VD_CONV_DUAL, VD_CONV_SPLIT, VD_CONV_POOL, VD_CONV_SLOPE, VD_CONV_CHAIN = 16, 32, 64, 128, 256, 512
VD_CONV_REVERSE = 1024
VD_CONV_TWO_DST = 2048
VD_MODE_MOSAIC, VD_MODE_GAUSS = 0, 1
VD_STAT_CAND_OVERFLOW, VD_STAT_DET_OVERFLOW, VD_STAT_WORK_OVERFLOW, VD_STAT_WATCHDOG = 1, 2, 4, 8
VD_STAT_JPEG_OVERFLOW = 16
VD_NMS_MAX_CAND = 8192
VD_MAX_BOXES_PER_FRAME = 4096


class VdFrame(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("pitch", C.c_int32),
                ("reserved", C.c_int32)]


class VdConvDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Cin", C.c_int32),
                ("Hout", C.c_int32), ("Wout", C.c_int32), ("Cout", C.c_int32),
                ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("dst_C", C.c_int32), ("dst_coff", C.c_int32), ("up_H", C.c_int32), ("up_W", C.c_int32),
                ("flags", C.c_uint32),
                ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("upadd", C.c_void_p), ("y", C.c_void_p),
                ("x2", C.c_void_p), ("Cin2", C.c_int32), ("Hin2", C.c_int32), ("Win2", C.c_int32),
                ("stride2", C.c_int32), ("slope", C.c_void_p),
                ("chain_w", C.c_void_p), ("chain_bias", C.c_void_p), ("chain_y", C.c_void_p), ("chain_Cout", C.c_int32),
                ("acc_scale", C.c_float), ("y_b", C.c_void_p), ("Cout1", C.c_int32), ("dst_b_C", C.c_int32),
                ("dst_b_coff", C.c_int32)]


class VdJpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32),
                ("h", C.c_int32 * 3), ("v", C.c_int32 * 3), ("blocks_x", C.c_int32 * 3), ("blocks_y", C.c_int32 * 3),
                ("restart_interval", C.c_int32), ("quant", (C.c_uint16 * 64) * 3)]


class VdJpegHuffDev(C.Structure):
    _fields_ = [("fast", C.c_uint16 * 512), ("maxcode", C.c_int32 * 18), ("valoff", C.c_int32 * 18), ("symbols", C.c_uint8 * 256)]


class VdJpegDevHdr(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32),
                ("h", C.c_int32 * 3), ("v", C.c_int32 * 3), ("blocks_x", C.c_int32 * 3), ("blocks_y", C.c_int32 * 3),
                ("mcux", C.c_int32), ("mcuy", C.c_int32), ("blocks_per_mcu", C.c_int32),
                ("blk_comp", C.c_uint8 * 12), ("blk_bx", C.c_uint8 * 12), ("blk_by", C.c_uint8 * 12),
                ("td", C.c_uint8 * 3), ("ta", C.c_uint8 * 3), ("pad_", C.c_uint8 * 2),
                ("quant", (C.c_uint16 * 64) * 3), ("scan_offset", C.c_uint32), ("scan_bytes", C.c_uint32),
                ("dc", VdJpegHuffDev * 2), ("ac", VdJpegHuffDev * 2)]


_SIGS = {
    "vd_abi_version": (C.c_int, []),
    "vd_act_is_fp16": (C.c_int, []),
    "vd_build_info": (C.c_char_p, []),
    "vd_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "vd_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "vd_x0_pitch": (C.c_int, [C.c_int]),
    "vd_ingest_u8hwc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "vd_ingest_u8hwc_identity": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "vd_ingest_u8hwc_identity_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "vd_ingest_u8hwc_nchw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "vd_conv_plan_create": (C.c_int, [C.POINTER(VdConvDesc), C.POINTER(C.c_void_p)]),
    "vd_conv_plan_destroy": (C.c_int, [C.c_void_p]),
    "vd_conv_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vd_conv_plan_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 6),
    "vd_maxpool3x3s2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vd_maxpool3x3s2_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vd_mnet_stem": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vd_mnet_dw3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_void_p]),
    "vd_mnet_dwpw": (C.c_int, [C.c_void_p] * 8 + [C.c_int] * 6 + [C.c_void_p]),
    "vd_mnet_pw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                             C.c_void_p]),
    "vd_decode_threshold": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vd_decode_threshold_lcl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vd_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vd_nms_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vd_gaussian_taps_q8": (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_uint16)]),
    "vd_composite_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_size_t]),
    "vd_composite_rois": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint16), C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "vd_copy_dirty_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vd_jpeg_read_header": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(VdJpegInfo)]),
    "vd_jpeg_decode_coefficients": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vd_jpeg_prepare_device": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(VdJpegDevHdr)]),
    "vd_jpeg_decode_workspace_bytes": (C.c_size_t, [C.c_int, C.POINTER(VdJpegDevHdr), C.c_size_t]),
    "vd_jpeg_decode_batch": (C.c_int, [C.c_void_p, C.POINTER(VdJpegDevHdr), C.c_void_p, C.c_int, C.c_size_t, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vd_jpeg_reconstruct_batch": (C.c_int, [C.c_void_p, C.POINTER(VdJpegDevHdr), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vd_jpeg_encode_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vd_jpeg_encode_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

EXPORTED = tuple(_SIGS)

_libs = {}


def lib(precision: str = "bf16"):
    """Loads libvdesens.so (bf16 activations, BASELINE's dtype) or libvdesens_fp16.so (same sources, fp16
    storage) once.  Raises if it has not been built (see __graft_entry__.build()); there is no fallback."""
    if precision not in LIB_PATHS:
        raise ValueError(f"precision must be 'bf16' or 'fp16', got {precision!r}")
    if precision not in _libs:
        path = LIB_PATHS[precision]
        if not os.path.isfile(path):
            raise RuntimeError(
                f"{path} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                f"g.build()' or make -C video-desensitization_b200/csrc). There is no CPU fallback.")
        l = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)       # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        assert l.vd_act_is_fp16() == (1 if precision == "fp16" else 0)
        _libs[precision] = l
    return _libs[precision]


def torch_dtype(precision: str):
    import torch
    return torch.float16 if precision == "fp16" else torch.bfloat16


def precision_of(dtype) -> str:
    import torch
    return "fp16" if dtype == torch.float16 else "bf16"


def last_error(l=None) -> str:
    """message of the last failure on this thread IN LIBRARY `l` (each .so keeps its own thread-local string,
    so the caller passes the library whose entry point returned the status; default = the bf16 library)"""
    buf = C.create_string_buffer(512)
    (l if l is not None else lib()).vd_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str = "", l=None):
    """reference behaviour at the boundary: Python exceptions, never an abort (combine_detect.py:226-228
    catches Exception and skips the batch).  `l` = the library that produced `rc`."""
    if rc != 0:
        raise RuntimeError(f"libvdesens {what} failed (status {rc}): {last_error(l)}")


def mode_id(mode: str) -> int:
    """"mosaic" (reference combine_detect.py:138-161) or "gauss" (north_star's cv2.GaussianBlur idiom); anything
    else is an error — a typo must not silently change the anonymisation method"""
    if mode == "mosaic":
        return VD_MODE_MOSAIC
    if mode == "gauss":
        return VD_MODE_GAUSS
    raise ValueError(f"mode must be 'mosaic' or 'gauss', got {mode!r}")


def gaussian_taps_q8(ksize: int, sigma: float = 0.0):
    arr = (C.c_uint16 * 256)()
    check(lib().vd_gaussian_taps_q8(int(ksize), float(sigma), arr), "vd_gaussian_taps_q8")
    return [int(arr[i]) for i in range(ksize)]
