/* vdesens.h — C-ABI of the B200-native face-anonymisation hot path (libvdesens.so).
 *
 * The reference (xdu-Liu-learn/Video-desensitization) is pure Python and has no FFI of its own
 * (zero C/CUDA sources).  This header is therefore the thin native layer that sits
 * directly beneath the reference's Python call surface; each entry point names the reference
 * interface it replaces (paths relative to the reference tree):
 *
 *   vd_ingest_u8hwc          <- Retinaface.preprocess          detect_face/face.py:65-88
 *                               letterbox_image/preprocess_input detect_face/utils/utils.py:8-29
 *   vd_conv_* / vd_maxpool   <- RetinaFace.forward             detect_face/retinaface.py:114-148
 *                               conv_bn/SSH/FPN                 detect_face/nets/layers.py:10-114
 *   vd_decode_threshold      <- decode/decode_landm/conf>=thr  detect_face/utils/utils_bbox.py:49-79,116
 *                               Anchors.get_anchors            detect_face/utils/anchors.py:22-41
 *   vd_nms_batched           <- non_max_suppression + retinaface_correct_boxes
 *                                                              detect_face/utils/utils_bbox.py:103-130,12-43
 *   vd_composite_rois        <- mosaic_rectangle_region_single + the per-frame box loop
 *                                                              combine_detect.py:138-161,241-249
 *                               (Gaussian mode: the cv2.GaussianBlur ROI idiom named by BASELINE.json)
 *
 * Conventions: plain pointers and sizes only (no torch types); every function returns 0 on success
 * or a vd_status code, never aborts, never allocates device memory and never synchronises the
 * device; all device work is enqueued on the caller's stream.  Device pointers are marked [dev],
 * host pointers [host].  vd_last_error() returns the message of the last failure on this thread.
 */
#ifndef VDESENS_H
#define VDESENS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VD_ABI_VERSION 1

typedef void* vd_stream_t; /* cudaStream_t */

enum vd_status {
  VD_OK = 0,
  VD_ERR_INVALID = 1,      /* bad argument */
  VD_ERR_CUDA = 2,         /* a CUDA runtime / driver call failed */
  VD_ERR_UNSUPPORTED = 3,  /* shape outside what the kernels implement */
  VD_ERR_NO_DEVICE = 4     /* no sm_100 device / driver entry point missing */
};

/* device-side status word bits written by kernels into `status` arguments */
#define VD_STAT_CAND_OVERFLOW 1u   /* more candidates than `cap` in some frame */
#define VD_STAT_DET_OVERFLOW 2u    /* more kept boxes than `max_det` */
#define VD_STAT_WORK_OVERFLOW 4u   /* composite workspace too small */
#define VD_STAT_WATCHDOG 8u        /* a device-side wait timed out */
#define VD_STAT_JPEG_OVERFLOW 16u  /* an encoded frame did not fit its stream / output buffer (its size is reported as -1) */

int vd_abi_version(void);
/* storage type of activations and weights this build computes in: 0 = bf16 (libvdesens.so),
 * 1 = fp16 (libvdesens_fp16.so, same sources built with -DVD_ACT_FP16) */
int vd_act_is_fp16(void);
const char* vd_build_info(void);
int vd_last_error(char* buf, size_t n);
/* fills SM count and compute capability of the current device */
int vd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------ */
/* K1 ingest: uint8 HWC frames -> mean-subtracted bf16 network input                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct vd_frame {
  const uint8_t* ptr; /* [dev] first byte of the frame, HWC, 3 channels */
  int32_t h, w;       /* frame size in pixels */
  int32_t pitch;      /* bytes per row */
  int32_t reserved;
} vd_frame_t;

/* Layout of the ingested tensor "x0": [B][H][Wp][4] bf16, channels 0..2 = pixel - mean,
 * channel 3 = 0, pixel column x stored at column x+3, columns outside [3, W+3) are zero.
 * Wp = vd_x0_pitch(W).  conv1 reads it through a sliding-window tensor map (8 taps x 4 ch = 64 B). */
int vd_x0_pitch(int W);

/* Letterboxes each frame to HxW (cv2 INTER_LINEAR uint8 fixed-point arithmetic, grey 128 pad,
 * integer-truncated target size and floor-divided placement exactly as utils.py:8-18), subtracts
 * mean[c] from channel c in caller order, writes x0.  frames: [dev] array of B descriptors.
 * The pad columns of x0 must have been zeroed once by the caller (cudaMemset at allocation). */
int vd_ingest_u8hwc(const vd_frame_t* frames, int B, void* x0, int H, int W,
                    const float mean[3], vd_stream_t stream);

/* Fast path for frames that already have the network's size (h == H, w == W for every frame:
 * the letterbox is the identity, BASELINE configs 2-5).  Same output as vd_ingest_u8hwc. */
int vd_ingest_u8hwc_identity(const vd_frame_t* frames, int B, void* x0, int H, int W,
                             const float mean[3], vd_stream_t stream);

/* vd_ingest_u8hwc_identity that also copies frame i into frames_out[i] (same h and w; pitches may differ) while it
 * has the bytes on chip: the reference's `img_with_mosaic = img.copy()` (combine_detect.py:247) fused into the
 * ingest pass, so that compositing can run in place on the copy and the caller's frames stay untouched. */
int vd_ingest_u8hwc_identity_copy(const vd_frame_t* frames, const vd_frame_t* frames_out, int B, void* x0,
                                  int H, int W, const float mean[3], vd_stream_t stream);

/* Optional fp32 NCHW output identical to Retinaface.preprocess()'s tensor (face.py:87). */
int vd_ingest_u8hwc_nchw_f32(const vd_frame_t* frames, int B, float* out, int H, int W,
                             const float mean[3], vd_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K2 convolution: implicit GEMM on tcgen05 (TMEM accumulators, TMA operand feed)              */
/* ------------------------------------------------------------------------------------------ */
#define VD_CONV_RELU 1u        /* ReLU in the epilogue (after the residual add)            */
#define VD_CONV_RESIDUAL 2u    /* + residual[pixel][co] before ReLU                         */
#define VD_CONV_UPADD 4u       /* + nearest-upsampled half-resolution tensor AFTER ReLU     */
#define VD_CONV_OUT_F32 8u     /* store fp32 instead of bf16 (head GEMM)                    */
#define VD_CONV_IN_X0 16u      /* input is the ingest layout (conv1: 7x7 s2 p3, 3->64)       */
#define VD_CONV_SPLIT 64u      /* split-precision tensors ("fp32-class" mode): every activation tensor has 3 planes of
                                  C channels [hi | lo | hi] with value = hi + lo; Cin counts all three planes, weights
                                  are [w_hi | w_hi | w_lo] per tap, so x_hi*w_hi + x_lo*w_hi + x_hi*w_lo accumulates
                                  in fp32.  dst_C, the residual and the upadd tensor have 3*C channels; the epilogue
                                  writes hi, lo and the hi copy.  Not combinable with VD_CONV_DUAL. */
#define VD_CONV_POOL 128u      /* with VD_CONV_IN_X0 | VD_CONV_RELU: torchvision's maxpool (3x3, stride 2, pad 1) is applied
                                  to the convolution's output inside the kernel; Hout/Wout stay conv1's extents and y is
                                  the POOLED tensor [B][(Hout+1)/2][(Wout+1)/2][64] (reference retinaface.py:71-73: the
                                  conv1 -> bn1 -> relu -> maxpool head of the ResNet body) */
#define VD_CONV_SLOPE 256u     /* activation = per-channel negative-side slope (desc.slope) instead of VD_CONV_RELU: the
                                  LeakyReLU(0.1) layers of the MobileNet0.25 variant (reference nets/layers.py:41,71,
                                  nets/mobilenet025.py:3-19) and tiles that mix ReLU / leaky / pass-through channels */
#define VD_CONV_CHAIN 512u     /* the kernel also runs the NEXT layer, a 1x1 convolution Cout -> chain_Cout (+bias, ReLU) of
                                  its own output tile while that tile is still in shared memory (bottleneck conv3 followed
                                  by the next bottleneck's conv1, reference torchvision resnet50 Bottleneck.forward): the
                                  wide tensor is written once and not read back.  Needs R = S = 1, Cout = 256,
                                  chain_Cout in {64, 128}; results are bit-identical to the two separate launches */
#define VD_CONV_REVERSE 1024u   /* scheduling hint, no effect on results: the persistent CTAs walk the output tiles from the
                                  last to the first.  A layer that reads what its predecessor has JUST written finds the
                                  tail of that tensor still in the 126 MB L2 if it starts there; the executor alternates
                                  the direction along the layer chain.  Ignored with VD_CONV_CHAIN / VD_CONV_POOL and by the row-pair
                                  kernel (3x3 64 -> 64 layers, see vd_conv_plan_create) */
#define VD_CONV_TWO_DST 2048u  /* two sibling layers that read the same tensor run as ONE convolution over the concatenated
                                  filters (Cout = Cout1 + Cout2 in {128, 192}): output channels [0, Cout1) go to y
                                  (dst_C, dst_coff), channels [Cout1, Cout) to y_b (dst_b_C, dst_b_coff).  The SSH module's
                                  conv3X3 + conv5X5_1 and conv5X5_2 + conv7X7_2 (reference nets/layers.py:54-66);
                                  bit-identical to the two launches */
#define VD_CONV_DUAL 32u       /* second K segment: + 1x1 conv (stride2) of tensor x2 accumulated into
                                  the same output (bottleneck conv3 + downsample projection fused) */

typedef struct vd_conv_desc {
  int32_t B;                 /* frames */
  int32_t Hin, Win, Cin;     /* input NHWC (Cin multiple of 64; for IN_X0: H, W of the frame) */
  int32_t Hout, Wout, Cout;  /* output extent, Cout multiple of 32 */
  int32_t R, S;              /* filter height / width */
  int32_t stride, pad;
  int32_t dst_C, dst_coff;   /* channels of the destination buffer and first channel written */
  int32_t up_H, up_W;        /* extent of the UPADD source (half resolution) */
  uint32_t flags;
  const void* x;             /* [dev] input activations, bf16 NHWC (or x0) */
  const void* w;             /* [dev] packed weights bf16 [Cout][R][S][Cin] (BN folded); IN_X0: [64][7][8 taps][4 ch] */
  const float* bias;         /* [dev] fp32 [Cout] (BN folded / conv bias) */
  const void* residual;      /* [dev] bf16 [B][Hout][Wout][Cout] or NULL */
  const void* upadd;         /* [dev] bf16 [B][up_H][up_W][Cout] or NULL */
  void* y;                   /* [dev] output */
  /* VD_CONV_DUAL: y += conv1x1_stride2(x2); weights are [Cout][R*S*Cin + Cin2], bias is the sum */
  const void* x2;            /* [dev] bf16 NHWC [B][Hin2][Win2][Cin2] or NULL */
  int32_t Cin2, Hin2, Win2, stride2;
  /* VD_CONV_SLOPE: fp32 [Cout], y = v > 0 ? v : v * slope[c] (0 = ReLU, 0.1 = LeakyReLU(0.1), 1 = identity) */
  const float* slope;
  /* VD_CONV_CHAIN: y2 = relu(conv1x1(y) + chain_bias); chain_w is [chain_Cout][Cout] (K-major), chain_y is
     [B][Hout][Wout][chain_Cout] in the activation type */
  const void* chain_w;
  const float* chain_bias;
  void* chain_y;
  int32_t chain_Cout;
  /* VD_CONV_SPLIT: the packed weights are w * 2^k (k per layer, so that the lo plane stays a normal fp16 number);
     the accumulators are multiplied by acc_scale = 2^-k before the bias is added.  0 means 1.0 */
  float acc_scale;
  /* VD_CONV_TWO_DST */
  void* y_b;                 /* [dev] second destination, [B][Hout][Wout][dst_b_C] in the activation type */
  int32_t Cout1;             /* channels of the first layer */
  int32_t dst_b_C, dst_b_coff;
} vd_conv_desc_t;

typedef struct vd_conv_plan vd_conv_plan_t;

/* Encodes the TMA tensor maps and picks the kernel and its tiling; host-only work, no device access.  Besides the generic
 * implicit-GEMM kernel there are two specialised ones behind the same three calls: the fused stem (VD_CONV_POOL) and, for
 * 3x3 / stride 1 / pad 1 / Cin = Cout = 64 layers with no flag other than VD_CONV_RELU whose width fills its 128-column
 * strips to >= 87 %, a kernel that computes two output rows per accumulator (plan_info reports TW = 128, TH = 2, BN = 128);
 * its results are bit-identical to the generic kernel's, VD_CONV_ROWS=0 in the environment disables it. */
int vd_conv_plan_create(const vd_conv_desc_t* desc, vd_conv_plan_t** plan);
int vd_conv_plan_destroy(vd_conv_plan_t* plan);
int vd_conv_run(const vd_conv_plan_t* plan, vd_stream_t stream);
/* tiling chosen for the plan: tile = TW x TH pixels x NB frames (<= 128 rows), BN output channels */
int vd_conv_plan_info(const vd_conv_plan_t* plan, int* TW, int* TH, int* NB, int* BN,
                      int* num_tiles, int* grid);

/* max-pool 3x3 stride 2 pad 1 on bf16 NHWC (torchvision resnet stem) */
int vd_maxpool3x3s2(const void* x, void* y, int B, int H, int W, int C, vd_stream_t stream);
/* same on split-precision tensors: [B][H][W][3*C] planes hi | lo | hi, maximum of hi + lo (VD_CONV_SPLIT) */
int vd_maxpool3x3s2_split(const void* x, void* y, int B, int H, int W, int C, vd_stream_t stream);

/* ---- MobileNet0.25 variant (Retinaface(backbone="mobilenet"), reference detect_face/nets/mobilenet025.py) ----------
 * The layers whose channel count is below the tcgen05 kernel's 64-channel granule are HBM-bound streaming kernels
 * (bf16/fp16 NHWC in and out, fp32 accumulate, folded BN bias, per-channel slope): depthwise taps as mixed-precision
 * FMAs on the CUDA cores, the stem and the pointwise layers with Cin in {8, 16, 32} as warp-level mma.sync GEMMs.
 * Every tensor, weight, bias and slope pointer must be 16-byte aligned; tensors up to 2^31 16-byte vectors.
 *   vd_mnet_stem   conv_bn(3, 8, stride 2) on the ingest layout x0            (mobilenet025.py:3-8, :26)
 *   vd_mnet_dw3x3  depthwise 3x3, stride 1 or 2, pad 1, C % 8 == 0            (conv_dw first half, :10-14)
 *   vd_mnet_pw     pointwise 1x1, Cin % 8 == 0, Cout % 8 == 0, Cin*Cout <= 4096 (conv_dw second half, :16-18)
 * Weights are in the activation type: stem [3][3][4][8] (r, s, channel incl. the zero 4th, cout), dw [9][C],
 * pw [Cin][Cout]. */
int vd_mnet_stem(const void* x0, const void* w, const float* bias, const float* slope, void* y, int B, int H, int W,
                 vd_stream_t stream);                        /* y: [B][(H+1)/2][(W+1)/2][8] */
int vd_mnet_dw3x3(const void* x, const void* w, const float* bias, const float* slope, void* y, int B, int H, int W,
                  int C, int stride, vd_stream_t stream);    /* y: [B][ceil(H/stride)][ceil(W/stride)][C] */
int vd_mnet_pw(const void* x, const void* w, const float* bias, const float* slope, void* y, long long npix, int Cin,
               int Cout, vd_stream_t stream);
/* one whole conv_dw block (mobilenet025.py:10-19) = vd_mnet_dw3x3 followed by vd_mnet_pw in one kernel, the depthwise
 * result staying in shared memory; (Cin, Cout) in {(8,16), (16,32), (32,32), (32,64)}; bit-identical to the two calls */
int vd_mnet_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw, const float* bpw,
                 const float* spw, void* y, int B, int H, int W, int Cin, int Cout, int stride, vd_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K3 decode + threshold + compaction                                                          */
/* ------------------------------------------------------------------------------------------ */
/* heads[l]: [dev] fp32 [B][h_l*w_l][32] from the fused head GEMM of level l; per cell:
 *   [0:8) bbox (2 anchors x 4), [8:12) class logits (2 anchors x (bg, face)), [12:32) landmarks.
 * Writes candidates with softmax face score >= conf_thr: cand [B][cap][16] fp32 =
 *   x1,y1,x2,y2,score,10 landmark coords (normalised letterbox space), anchor index (int bits);
 * cand_count [B] is zeroed by this call.  Order inside a frame is unspecified (NMS sorts). */
int vd_decode_threshold(const float* heads_l0, const float* heads_l1, const float* heads_l2,
                        int B, int H, int W, float conf_thr, int cap, float* cand,
                        int32_t* cand_count, uint32_t* status, vd_stream_t stream);

/* Same decode for separately laid-out reference tensors loc [B][A][4], conf [B][A][2] (already
 * softmaxed), landms [B][A][10] (the signature of Retinaface.postprocess, face.py:93). */
int vd_decode_threshold_lcl(const float* loc, const float* conf, const float* landms, int B, int H,
                            int W, float conf_thr, int cap, float* cand, int32_t* cand_count,
                            uint32_t* status, vd_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K4 batched NMS + un-letterbox                                                               */
/* ------------------------------------------------------------------------------------------ */
#define VD_NMS_MAX_CAND 8192
size_t vd_nms_workspace_bytes(int B, int cap);
/* Greedy NMS per frame, torchvision semantics (descending score, ties by lower anchor index,
 * suppress when IoU > iou_thr, areas without +1).  unletterbox [B][4] = offx, offy, scalex, scaley
 * (fp32, computed as utils_bbox.py:26-31); out_rows [B][max_det][16]: 15 corrected values +
 * anchor index, in descending score; out_count [B]. */
int vd_nms_batched(const float* cand, const int32_t* cand_count, int B, int cap, float iou_thr,
                   const float* unletterbox, int max_det, float* out_rows, int32_t* out_count,
                   uint32_t* status, void* workspace, size_t workspace_bytes, vd_stream_t stream);

/* ------------------------------------------------------------------------------------------ */
/* K5 compositing: sequential mosaic / Gaussian blur over boxes, in place                      */
/* ------------------------------------------------------------------------------------------ */
#define VD_MODE_MOSAIC 0
#define VD_MODE_GAUSS 1
#define VD_MAX_KSIZE 255
#define VD_MAX_BOXES_PER_FRAME 4096

/* cv::getGaussianKernel + 8.8 fixed point with error diffusion (cv2.GaussianBlur uint8 path);
 * taps_q8 [host] receives ksize values summing to 256. */
int vd_gaussian_taps_q8(int ksize, double sigma, uint16_t* taps_q8);

size_t vd_composite_workspace_bytes(int B, int max_boxes, size_t scratch_pixels_per_frame);

/* frames [dev] B descriptors (ptr is written through: const is cast away by the library);
 * det_rows/det_count: [dev] output of vd_nms_batched (may be NULL): face boxes, applied first in
 * row order, coordinates = int(row[0..3] * (w,h,w,h)) truncated toward zero (combine_detect.py:243);
 * ext_boxes [dev] [B][max_ext][4] int32 + ext_count [dev] [B]: externally supplied (plate) boxes
 * applied after the faces (combine_detect.py:244).  Boxes are clipped to the frame and skipped if
 * empty (combine_detect.py:145-150).  Each box sees the result of the previous ones. */
int vd_composite_rois(const vd_frame_t* frames, int B, const float* det_rows,
                      const int32_t* det_count, int max_det, const int32_t* ext_boxes,
                      const int32_t* ext_count, int max_ext, int mode, int mosaic_level,
                      const uint16_t* taps_q8, int ksize, uint32_t* status, void* workspace,
                      size_t workspace_bytes, vd_stream_t stream);

/* ---- host half of a device-resident JPEG ingest (replaces the entropy-decoding part of cv2.imread,
 * reference combine_detect.py:167-180).  HOST functions: no CUDA call, no allocation.  Baseline / sequential Huffman,
 * 8-bit, 1 or 3 components with sampling factors <= 2, one interleaved scan, restart intervals — what cv2.imwrite
 * produces.  Anything else returns VD_ERR_UNSUPPORTED and the caller keeps using its CPU decoder. ---- */
typedef struct vd_jpeg_info {
  int32_t width, height, ncomp;       /* ncomp = 1 (grey) or 3 (YCbCr) */
  int32_t h[3], v[3];                 /* sampling factors per component */
  int32_t blocks_x[3], blocks_y[3];   /* 8x8-block grid per component (whole MCUs) */
  int32_t restart_interval;           /* MCUs between RSTn markers, 0 = none */
  uint16_t quant[3][64];              /* quantisation table per component, natural (row-major) order */
} vd_jpeg_info_t;

int vd_jpeg_read_header(const uint8_t* data, size_t size, vd_jpeg_info_t* info);

/* coefK [host]: blocks_y[K] * blocks_x[K] * 64 int16, blocks in raster order, coefficients in natural order, still
 * quantised (multiply by info.quant[K] to de-quantise).  coef1 / coef2 may be NULL for a grey image. */
int vd_jpeg_decode_coefficients(const uint8_t* data, size_t size, int16_t* coef0, int16_t* coef1, int16_t* coef2);

/* After vd_composite_rois only the pixels under a box differ from the frame the caller still holds.  Copies every
 * 64 x 16-pixel tile that a (truncated, clipped — same rules as the compositing pass) box touches from frames[i] to
 * dst_frames[i] (same sizes).  dst_frames may describe PINNED HOST memory (device-accessible under unified addressing):
 * the tiles are then written over PCIe by the kernel itself, and the device->host traffic of a batch shrinks from the
 * whole frames to the changed tiles.  bytes_copied: [dev] counter, incremented by the bytes written. */
int vd_copy_dirty_tiles(const vd_frame_t* frames, const vd_frame_t* dst_frames, int B, int max_h, int max_w,
                        const float* det_rows, const int32_t* det_count, int max_det,
                        const int32_t* ext_boxes, const int32_t* ext_count, int max_ext,
                        unsigned long long* bytes_copied, vd_stream_t stream);

/* ---- N1, near end: baseline JPEG decode on the device (csrc/jpeg_dec.cu) ----
 * Replaces cv2.imread + BGR->RGB of the reference frame loop (combine_detect.py:167-173): the FILE bytes go to the GPU
 * (~0.9 MB per 1080p frame instead of 6.2 MB of pixels), and Huffman decoding (parallel, self-synchronising),
 * de-quantisation, jpeg_idct_islow, fancy up-sampling and YCbCr->RGB all run there.  The pixels are IDENTICAL to
 * cv2.imread's (libjpeg-turbo defaults) — checked against oracle/jpeg_oracle.py, which is pinned on cv2.imdecode.
 * vd_jpeg_prepare_device (host, ~0.1 ms) parses one file into the record the kernels read. */
typedef struct vd_jpeg_huff_dev {
  uint16_t fast[512];          /* 9-bit look-ahead: (length << 8) | symbol, 0 = longer code */
  int32_t maxcode[18];         /* largest code of each length 1..16 (-1 = none), [17] = INT32_MAX */
  int32_t valoff[18];          /* symbol index = valoff[length] + code */
  uint8_t symbols[256];
} vd_jpeg_huff_dev_t;

typedef struct vd_jpeg_dev_hdr {
  int32_t width, height, ncomp;            /* ncomp 1 (grey) or 3 (YCbCr) */
  int32_t h[3], v[3];                      /* sampling factors: luma 1x1, 2x1 or 2x2, chroma 1x1 */
  int32_t blocks_x[3], blocks_y[3];        /* block grid per component (whole MCUs) */
  int32_t mcux, mcuy, blocks_per_mcu;
  uint8_t blk_comp[12], blk_bx[12], blk_by[12];   /* per block of an MCU, in scan order: component, position in its footprint */
  uint8_t td[3], ta[3], pad_[2];           /* Huffman table of each component: dc[td], ac[ta] (0 or 1) */
  uint16_t quant[3][64];                   /* natural order */
  uint32_t scan_offset;                    /* where the entropy-coded segment starts inside the caller's byte buffer */
  uint32_t scan_bytes;                     /* its length, up to (not including) the marker that ends it */
  vd_jpeg_huff_dev_t dc[2], ac[2];
} vd_jpeg_dev_hdr_t;

/* Parses one file.  VD_ERR_UNSUPPORTED for anything the device path does not take (progressive, restart markers, more
 * than two table ids, other sampling patterns): the caller falls back to its CPU decoder for that file.
 * hdr->scan_offset is relative to `data`; the caller adds the file's offset inside its staging buffer. */
int vd_jpeg_prepare_device(const uint8_t* data, size_t size, vd_jpeg_dev_hdr_t* hdr);

/* Decodes B files of ONE geometry (same width, height, components and sampling; Huffman and quantisation tables may
 * differ per file) into B RGB frames.  hdrs: [dev] B records; bytes: [dev] the buffer their scan_offset refers to;
 * frames_out: [dev] B descriptors of width x height uint8 HWC frames.  result: [dev] int32[B]; bits 0..7: 0 = ok, else a
 * bit mask (1 = the parallel entropy decoder did not converge, 2 = block count mismatch, 4 = state mismatch while
 * writing: decode that file on the CPU); bits 8..15: the synchronisation round after which nothing changed (diagnostic).  sub_bytes: bytes of entropy data per decoder thread (0 = default 512).  Never allocates, never synchronises. */
size_t vd_jpeg_decode_workspace_bytes(int B, const vd_jpeg_dev_hdr_t* geometry, size_t max_scan_bytes);
int vd_jpeg_decode_batch(const vd_jpeg_dev_hdr_t* hdrs, const vd_jpeg_dev_hdr_t* geometry_host, const uint8_t* bytes,
                         int B, size_t max_scan_bytes, const vd_frame_t* frames_out, int32_t* result, int sub_bytes,
                         void* workspace, size_t workspace_bytes, vd_stream_t stream);
/* test hook: the same de-quantisation / IDCT / up-sampling / colour kernels fed with coefficient planes produced
 * elsewhere (vd_jpeg_decode_coefficients): coef c is [dev] int16 [B][blocks_y[c]*blocks_x[c]*64], natural order */
int vd_jpeg_reconstruct_batch(const vd_jpeg_dev_hdr_t* hdrs, const vd_jpeg_dev_hdr_t* geometry_host, int B,
                              const int16_t* coef0, const int16_t* coef1, const int16_t* coef2,
                              const vd_frame_t* frames_out, void* workspace, size_t workspace_bytes, vd_stream_t stream);

/* ---- N1, far end: baseline JPEG encode of resident frames (csrc/jpeg_enc.cu) ----
 * Replaces cv2.imwrite(path, bgr) of the reference frame loop (combine_detect.py:176-180) for frames that are already
 * in HBM.  The files are BYTE-IDENTICAL to cv2.imwrite's (libjpeg-turbo: 4:2:0, the Annex K Huffman tables, no restart
 * markers, JFIF 1.01 header) for the same pixels and quality; cv2's default quality is 95.
 * frames: [dev] B descriptors of RGB (caller channel order = the order cv2 would be given after its BGR swap) frames,
 * all H x W.  out: [dev] B * out_pitch bytes; frame i's file starts at out + i * out_pitch and is out_sizes[i] bytes long
 * (out_sizes: [dev] int32[B]; -1 and status bit VD_STAT_JPEG_OVERFLOW if it did not fit out_pitch or the workspace).
 * Never allocates, never synchronises. */
size_t vd_jpeg_encode_workspace_bytes(int B, int H, int W);
int vd_jpeg_encode_batch(const vd_frame_t* frames, int B, int H, int W, int quality, uint8_t* out, size_t out_pitch,
                         int32_t* out_sizes, uint32_t* status, void* workspace, size_t workspace_bytes,
                         vd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VDESENS_H */
