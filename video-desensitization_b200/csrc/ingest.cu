// ingest.cu — K1: uint8 HWC frames -> letterboxed, mean-subtracted network input.
//
// Replaces Retinaface.preprocess (reference detect_face/face.py:65-88) = letterbox_image
// (detect_face/utils/utils.py:8-18: cv2.resize INTER_LINEAR on uint8, grey-128 canvas, floor-divided
// placement, int()-truncated target size) + preprocess_input (utils.py:27-29: subtract (104,117,123)
// from channels in caller order) + HWC->CHW + float conversion + H2D.  Frames are already resident
// on the device as uint8, so the 4 B/px/channel host tensor and its H2D copy disappear.
//
// The bilinear resize restates cv2's fixed-point path bit-for-bit: coefficients
// rint(f*2048), horizontal sums in int32, vertical (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.
// HBM-bound: 3 B read + 8 B written per output pixel (x0 layout), coalesced vector stores.
#include "common.cuh"

namespace {

struct LetterGeom {
  int nw, nh, dx, dy;      // resized size and placement inside the HxW canvas
  double sx, sy;           // src/dst scale of cv2.resize
};

__device__ __forceinline__ LetterGeom letter_geom(int ih, int iw, int H, int W) {
  LetterGeom g;
  // python: scale = min(w/iw, h/ih); nw = int(iw*scale); nh = int(ih*scale)   (float64)
  const double scale = fmin(static_cast<double>(W) / iw, static_cast<double>(H) / ih);
  g.nw = static_cast<int>(iw * scale);
  g.nh = static_cast<int>(ih * scale);
  g.dx = (W - g.nw) / 2;     // operands are non-negative: C division == python floor division
  g.dy = (H - g.nh) / 2;
  g.sx = static_cast<double>(iw) / g.nw;   // cv2: scale_x = 1. / (dsize.width / (double)ssize.width)
  g.sy = static_cast<double>(ih) / g.nh;
  return g;
}

// cv2 resize.cpp treats the two axes differently at the borders (only reachable when UP-scaling): along x a tap pair
// that starts left of the row, or on its last pixel, becomes that pixel with weight 1 (CLAMP = true); along y the weights
// are kept and the two ROW INDICES are clipped — the first / last output rows blend the border row with itself, and the
// two separately truncated products can come out one below the full-weight product.  lin_coeff returns the unclipped s0
// for CLAMP = false; the caller clips s0 and s0 + 1 into [0, n-1].
template <bool CLAMP>
__device__ __forceinline__ void lin_coeff(int d, double scale, int n, int& s0, int& a0, int& a1) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= s;
  if (CLAMP) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
  }
  s0 = s;
  a0 = __float2int_rn((1.f - f) * 2048.f);   // cvRound: half to even
  a1 = __float2int_rn(f * 2048.f);
}

// one letterboxed pixel (3 channels, uint8 values 0..255)
__device__ __forceinline__ void letter_pixel(const vd_frame_t& fr, const LetterGeom& g, int y, int x, int v[3]) {
  const int yy = y - g.dy, xx = x - g.dx;
  if (yy < 0 || yy >= g.nh || xx < 0 || xx >= g.nw) {
    v[0] = v[1] = v[2] = 128;
    return;
  }
  if (g.nw == fr.w && g.nh == fr.h) {       // cv2.resize to the same size is a copy
    const uint8_t* p = fr.ptr + static_cast<size_t>(yy) * fr.pitch + xx * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
    return;
  }
  int x0, xa0, xa1, ys, yb0, yb1;
  lin_coeff<true>(xx, g.sx, fr.w, x0, xa0, xa1);
  lin_coeff<false>(yy, g.sy, fr.h, ys, yb0, yb1);
  const int x1 = min(x0 + 1, fr.w - 1);
  const int y0 = min(max(ys, 0), fr.h - 1), y1 = min(max(ys + 1, 0), fr.h - 1);
  const uint8_t* r0 = fr.ptr + static_cast<size_t>(y0) * fr.pitch;
  const uint8_t* r1 = fr.ptr + static_cast<size_t>(y1) * fr.pitch;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = r0[x0 * 3 + c] * xa0 + r0[x1 * 3 + c] * xa1;
    const int h1 = r1[x0 * 3 + c] * xa0 + r1[x1 * 3 + c] * xa1;
    v[c] = (((yb0 * (h0 >> 4)) >> 16) + ((yb1 * (h1 >> 4)) >> 16) + 2) >> 2;
  }
}

// x0 layout: [B][H][Wp][4] activation type, pixel x at column x+3.
// Generic (resizing) path, one block = one output row: the two source rows a bilinear output row
// blends are staged in shared memory with 16-byte vector loads (every load of the block issued before the first store;
// a chunk that would cross the frame buffer's ends is fetched byte by byte), then one thread = one output pixel reads its
// 12 source bytes from shared memory.  (Round 1's form — one thread = one output pixel, the 12 bytes as single-byte global
// loads — measured 0.302 ms for 16 x 4K -> 1080p, 0.201 ms for 16 x 1080p -> 1280 x 1280 and 0.296 ms for 720p -> 1080p;
// this one 0.209 / 0.111 / 0.157 ms on the same box.)  Arithmetic is letter_pixel's, bit for bit.
constexpr int LB_THREADS = 256;
constexpr int LB_MAX_SMEM = 47 * 1024;

__global__ void __launch_bounds__(LB_THREADS)
ingest_x0_rows_kernel(const vd_frame_t* __restrict__ frames, uint2* __restrict__ x0, int H, int W, int Wp, float m0,
                      float m1, float m2, int row_pitch) {
  extern __shared__ __align__(16) uint8_t s_rows[];        // [2][row_pitch]
  const int b = blockIdx.z, y = blockIdx.y, tid = threadIdx.x;
  const vd_frame_t fr = frames[b];
  const LetterGeom g = letter_geom(fr.h, fr.w, H, W);
  uint2* row = x0 + (static_cast<size_t>(b) * H + y) * Wp;
  const float grey0 = 128.f - m0, grey1 = 128.f - m1, grey2 = 128.f - m2;
  const uint2 grey = make_uint2(pack_bf16x2(grey0, grey1), pack_bf16x2(grey2, 0.f));
  const int yy = y - g.dy;
  const bool same = g.nw == fr.w && g.nh == fr.h;          // cv2.resize to the same size is a copy
  const int nbytes = 3 * fr.w;
  if (yy < 0 || yy >= g.nh || 2 * (nbytes + 32) > 2 * row_pitch) {
    // canvas rows above / below the image; (a frame wider than the staging buffer: the per-pixel form)
    for (int x = tid; x < W; x += LB_THREADS) {
      uint2 o = grey;
      if (yy >= 0 && yy < g.nh) {
        int v[3];
        letter_pixel(fr, g, y, x, v);
        o = make_uint2(pack_bf16x2(static_cast<float>(v[0]) - m0, static_cast<float>(v[1]) - m1),
                       pack_bf16x2(static_cast<float>(v[2]) - m2, 0.f));
      }
      row[x + 3] = o;
    }
    return;
  }
  int ys = yy, yb0 = 2048, yb1 = 0;
  if (!same) lin_coeff<false>(yy, g.sy, fr.h, ys, yb0, yb1);
  const int y0 = min(max(ys, 0), fr.h - 1), y1 = min(max(ys + 1, 0), fr.h - 1);
  // ---- stage source rows y0 and y1: s_rows[r][phase_r + i] = byte i of the row ----
  const uint8_t* fr_end = fr.ptr + static_cast<size_t>(fr.h - 1) * fr.pitch + nbytes;
  const uint8_t* src[2] = {fr.ptr + static_cast<size_t>(y0) * fr.pitch, fr.ptr + static_cast<size_t>(y1) * fr.pitch};
  int phase[2];
  phase[0] = static_cast<int>(reinterpret_cast<uintptr_t>(src[0]) & 15);
  phase[1] = static_cast<int>(reinterpret_cast<uintptr_t>(src[1]) & 15);
  const int nch0 = (phase[0] + nbytes + 15) >> 4, nch1 = (phase[1] + nbytes + 15) >> 4;
  const int nch = nch0 + nch1;                              // chunks of both rows, one index space
  for (int base = 0; base < nch; base += 4 * LB_THREADS) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * LB_THREADS + tid;
      v[u] = make_uint4(0u, 0u, 0u, 0u);
      if (i < nch) {
        const int r = i >= nch0, ch = r ? i - nch0 : i;
        const uint8_t* pc = src[r] - phase[r] + 16 * ch;
        if (pc >= fr.ptr && pc + 16 <= fr_end) {
          v[u] = __ldg(reinterpret_cast<const uint4*>(pc));
        } else {
          uint8_t* vb = reinterpret_cast<uint8_t*>(&v[u]);
          for (int k = 0; k < 16; ++k)
            if (pc + k >= fr.ptr && pc + k < fr_end) vb[k] = __ldg(pc + k);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * LB_THREADS + tid;
      if (i < nch) {
        const int r = i >= nch0, ch = r ? i - nch0 : i;
        *reinterpret_cast<uint4*>(s_rows + r * row_pitch + 16 * ch) = v[u];
      }
    }
  }
  __syncthreads();
  const uint8_t* r0 = s_rows + phase[0];
  const uint8_t* r1 = s_rows + row_pitch + phase[1];
  for (int x = tid; x < W; x += LB_THREADS) {
    const int xx = x - g.dx;
    uint2 o = grey;
    if (xx >= 0 && xx < g.nw) {
      int v[3];
      if (same) {
        v[0] = r0[xx * 3]; v[1] = r0[xx * 3 + 1]; v[2] = r0[xx * 3 + 2];
      } else {
        int xs, xa0, xa1;
        lin_coeff<true>(xx, g.sx, fr.w, xs, xa0, xa1);
        const int x1 = min(xs + 1, fr.w - 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = r0[xs * 3 + c] * xa0 + r0[x1 * 3 + c] * xa1;
          const int h1 = r1[xs * 3 + c] * xa0 + r1[x1 * 3 + c] * xa1;
          v[c] = (((yb0 * (h0 >> 4)) >> 16) + ((yb1 * (h1 >> 4)) >> 16) + 2) >> 2;
        }
      }
      o = make_uint2(pack_bf16x2(static_cast<float>(v[0]) - m0, static_cast<float>(v[1]) - m1),
                     pack_bf16x2(static_cast<float>(v[2]) - m2, 0.f));
    }
    row[x + 3] = o;
  }
}

// fast path: frame size == network input (BASELINE configs 2-5): no resize, no pad.
//
// One block = one run of ID_PAIRS_PER_BLOCK column pairs of one row.  Phase 1 moves the raw bytes those pairs need
// into shared memory with 16-byte vector loads (`ld.global.nc.v4`, lane-contiguous: a warp load = 512 contiguous
// bytes; 48 bytes = 16 pixels per three of them) — or, when the row segment is not 16-byte aligned (odd pitch,
// sliced tensors), with byte loads.  With `copy_dst` the same staged bytes are also written to a second frame set
// (the reference's `img.copy()`, combine_detect.py:247, fused: the frame is read from HBM once).  Phase 2: one thread
// = one 16-byte store = two x0 columns (2j, 2j+1), i.e. pixels 2j-3 and 2j-2, whose six bytes start at byte 6j-9 of
// the row: three aligned 32-bit shared-memory loads and two funnel shifts.  Every store instruction of a warp writes
// 512 contiguous bytes.  The pad columns (0..2 and W+3..Wp-1) are written as zero, so the kernel does not depend on a
// pre-zeroed tensor.
constexpr int ID_THREADS = 256;
constexpr int ID_PAIRS = 4;                                  // pairs per thread
constexpr int ID_PAIRS_PER_BLOCK = ID_THREADS * ID_PAIRS;    // 1024 pairs = 2048 columns = 6144 source bytes
constexpr int ID_SMEM_BYTES = 6 * ID_PAIRS_PER_BLOCK + 64;   // + alignment slack on both sides

// ROWS rows per CTA: the (at most two per thread and row) 16-byte loads of all rows are issued before the first store, so a
// CTA keeps ROWS x 5.7 KB in flight instead of one row's worth.
template <int ROWS>
__global__ void __launch_bounds__(ID_THREADS)
ingest_x0_identity_kernel(const vd_frame_t* __restrict__ frames, const vd_frame_t* __restrict__ copy_dst,
                          uint4* __restrict__ x0, int H, int W, int Wp, float m0, float m1, float m2) {
  __shared__ __align__(16) uint8_t s_raw_all[ROWS][ID_SMEM_BYTES];
  const int b = blockIdx.z;
  const vd_frame_t fr = frames[b];
  const int npairs = Wp >> 1;
  const int j0 = blockIdx.x * ID_PAIRS_PER_BLOCK;
  const int j1 = min(j0 + ID_PAIRS_PER_BLOCK, npairs);
  const int row_bytes = 3 * W;
  // bytes [lo, hi) of the row feed pairs [j0, j1): pixels 2*j0 - 3 .. 2*j1 - 4 (clipped to the row); consecutive
  // blocks' ranges tile the row, so with copy_dst every byte of it is copied
  const int lo = max(0, 6 * j0 - 9), hi = min(row_bytes, 6 * j1 - 9);
  vd_frame_t fd;
  fd.ptr = nullptr; fd.pitch = 0;
  if (copy_dst) fd = copy_dst[b];
  int s_base_r[ROWS];
  uint4 v[ROWS][2];
  // ---- phase 1a: request the aligned 16-byte vectors of every row ----
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int y = blockIdx.y * ROWS + rr;
    s_base_r[rr] = 0;
    if (y >= H || hi <= lo) continue;
    const uint8_t* __restrict__ src = fr.ptr + static_cast<size_t>(y) * fr.pitch;
    // shared-memory byte i holds row byte s_base + i; s_base keeps the row's own 16-byte phase so that the vector
    // loads are aligned on both sides
    const int phase = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15);
    const int s_base = ((lo + phase) & ~15) - phase;        // <= lo, and (src + s_base) is 16-byte aligned
    s_base_r[rr] = s_base;
    const int v_lo = s_base < 0 ? s_base + 16 : s_base;     // first 16-byte vector fully inside the row
    const int v_hi = s_base + ((hi - s_base) & ~15);        // end of the last one
#pragma unroll
    for (int q = 0; q < 2; ++q) {                           // (6144 + 64) / 16 <= 2 * ID_THREADS vectors per row chunk
      const int o = v_lo + 16 * (static_cast<int>(threadIdx.x) + q * ID_THREADS);
      if (o < v_hi) v[rr][q] = __ldg(reinterpret_cast<const uint4*>(src + o));
    }
  }
  // ---- phase 1b: into shared memory (and the frame copy), ragged ends byte by byte ----
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int y = blockIdx.y * ROWS + rr;
    if (y >= H || hi <= lo) continue;
    const uint8_t* __restrict__ src = fr.ptr + static_cast<size_t>(y) * fr.pitch;
    uint8_t* s_raw = s_raw_all[rr];
    const int s_base = s_base_r[rr];
    uint8_t* dst_row = fd.ptr ? const_cast<uint8_t*>(fd.ptr) + static_cast<size_t>(y) * fd.pitch : nullptr;
    const int v_lo = s_base < 0 ? s_base + 16 : s_base;
    const int v_hi = s_base + ((hi - s_base) & ~15);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int o = v_lo + 16 * (static_cast<int>(threadIdx.x) + q * ID_THREADS);
      if (o < v_hi) {
        *reinterpret_cast<uint4*>(s_raw + (o - s_base)) = v[rr][q];
        if (dst_row) {
          if ((reinterpret_cast<uintptr_t>(dst_row + o) & 15) == 0) {
            *reinterpret_cast<uint4*>(dst_row + o) = v[rr][q];
          } else {
            const uint8_t* vb = reinterpret_cast<const uint8_t*>(&v[rr][q]);
            for (int k = 0; k < 16; ++k) dst_row[o + k] = vb[k];
          }
        }
      }
    }
    // ragged head and tail of the segment (and everything, if v_hi <= v_lo)
    const int head_end = min(max(v_lo, lo), hi), tail_begin = max(min(v_hi, hi), head_end);
    for (int o = lo + static_cast<int>(threadIdx.x); o < head_end; o += ID_THREADS) {
      const uint8_t bv = __ldg(src + o);
      s_raw[o - s_base] = bv;
      if (dst_row) dst_row[o] = bv;
    }
    for (int o = tail_begin + static_cast<int>(threadIdx.x); o < hi; o += ID_THREADS) {
      const uint8_t bv = __ldg(src + o);
      s_raw[o - s_base] = bv;
      if (dst_row) dst_row[o] = bv;
    }
  }
  __syncthreads();
  // ---- phase 2: x0 rows ----
#pragma unroll
  for (int rr = 0; rr < ROWS; ++rr) {
    const int y = blockIdx.y * ROWS + rr;
    if (y >= H) continue;
    const int s_base = s_base_r[rr];
    uint4* row = x0 + (static_cast<size_t>(b) * H + y) * npairs;
    const uint32_t* s_w = reinterpret_cast<const uint32_t*>(s_raw_all[rr]);
#pragma unroll
    for (int p = 0; p < ID_PAIRS; ++p) {
      const int j = j0 + p * ID_THREADS + static_cast<int>(threadIdx.x);
      if (j >= j1) continue;
      const int xa = 2 * j - 3;                       // pixel of column 2j; xa + 1 is the pixel of column 2j + 1
      const bool va = xa >= 0 && xa < W, vb = xa + 1 >= 0 && xa + 1 < W;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (va || vb) {
        const int off = 3 * xa - s_base;              // shared-memory byte of the pair's first byte (may be < 0 if !va)
        const int w = off >> 2;                       // arithmetic shift: floor
        const uint32_t sh = static_cast<uint32_t>(off & 3) * 8u;
        // words outside the staged range are never selected into a VALID pixel's bytes
        const uint32_t w0 = w >= 0 ? s_w[w] : 0u, w1 = s_w[w + 1], w2 = (4 * (w + 2) < ID_SMEM_BYTES) ? s_w[w + 2] : 0u;
        const uint32_t lo4 = __funnelshift_r(w0, w1, sh);     // bytes 0..3 of the pair
        const uint32_t hi4 = __funnelshift_r(w1, w2, sh);     // bytes 4..7 (4, 5 used)
        if (va) {
          o.x = pack_bf16x2(static_cast<float>(lo4 & 0xFFu) - m0, static_cast<float>((lo4 >> 8) & 0xFFu) - m1);
          o.y = pack_bf16x2(static_cast<float>((lo4 >> 16) & 0xFFu) - m2, 0.f);
        }
        if (vb) {
          o.z = pack_bf16x2(static_cast<float>(lo4 >> 24) - m0, static_cast<float>(hi4 & 0xFFu) - m1);
          o.w = pack_bf16x2(static_cast<float>((hi4 >> 8) & 0xFFu) - m2, 0.f);
        }
      }
      row[j] = o;
    }
  }
}

// fp32 NCHW output, byte-identical to the tensor Retinaface.preprocess returns (face.py:87)
__global__ void __launch_bounds__(256)
ingest_nchw_kernel(const vd_frame_t* __restrict__ frames, float* __restrict__ out, int H, int W, float m0, float m1,
                   float m2) {
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const vd_frame_t fr = frames[b];
  const LetterGeom g = letter_geom(fr.h, fr.w, H, W);
  const size_t plane = static_cast<size_t>(H) * W;
  float* o = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(y) * W;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
    int v[3];
    letter_pixel(fr, g, y, x, v);
    o[x] = static_cast<float>(v[0]) - m0;
    o[plane + x] = static_cast<float>(v[1]) - m1;
    o[2 * plane + x] = static_cast<float>(v[2]) - m2;
  }
}

}  // namespace

extern "C" int vd_ingest_u8hwc(const vd_frame_t* frames, int B, void* x0, int H, int W, const float mean[3],
                               vd_stream_t stream) {
  VD_REQUIRE(frames && x0 && mean, "vd_ingest_u8hwc: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 65535 && B <= 65535, "vd_ingest_u8hwc: bad extent");
  const int Wp = vd_x0_pitch(W);
  // staging buffer for two source rows of the widest frame that fits 47 KB (frames are on the device: their widths are
  // not known here, so the kernel falls back to the per-pixel form, row by row, for a frame that does not fit)
  const int row_pitch = (LB_MAX_SMEM / 2) & ~15;
  dim3 grid(1, H, B);
  ingest_x0_rows_kernel<<<grid, LB_THREADS, 2 * row_pitch, static_cast<cudaStream_t>(stream)>>>(
      frames, static_cast<uint2*>(x0), H, W, Wp, mean[0], mean[1], mean[2], row_pitch);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

// identity fast path, selected by the host when every frame already has the network's size
static int ingest_identity(const vd_frame_t* frames, const vd_frame_t* copy_dst, int B, void* x0, int H, int W,
                           const float mean[3], vd_stream_t stream) {
  VD_REQUIRE(frames && x0 && mean, "vd_ingest_u8hwc_identity: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 65535 && B <= 65535, "vd_ingest_u8hwc_identity: bad extent");
  const int Wp = vd_x0_pitch(W);
  // two rows per CTA (measured on one box, ingest + frame copy of 16 x 1080p: 1 row 0.0810 ms, 2 rows 0.0803, 4 rows 0.0822)
  dim3 grid(vd_cdiv(Wp / 2, ID_PAIRS_PER_BLOCK), vd_cdiv(H, 2), B);      // 1080p: one block per row pair (968 column pairs)
  ingest_x0_identity_kernel<2><<<grid, ID_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      frames, copy_dst, static_cast<uint4*>(x0), H, W, Wp, mean[0], mean[1], mean[2]);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

extern "C" int vd_ingest_u8hwc_identity(const vd_frame_t* frames, int B, void* x0, int H, int W,
                                        const float mean[3], vd_stream_t stream) {
  return ingest_identity(frames, nullptr, B, x0, H, W, mean, stream);
}

// identity ingest that ALSO copies every frame into `frames_out` (same sizes; pitches may differ): the reference's
// `img_with_mosaic = img.copy()` (combine_detect.py:247) without a second pass over the frame.
extern "C" int vd_ingest_u8hwc_identity_copy(const vd_frame_t* frames, const vd_frame_t* frames_out, int B, void* x0,
                                             int H, int W, const float mean[3], vd_stream_t stream) {
  VD_REQUIRE(frames_out, "vd_ingest_u8hwc_identity_copy: null argument");
  return ingest_identity(frames, frames_out, B, x0, H, W, mean, stream);
}

extern "C" int vd_ingest_u8hwc_nchw_f32(const vd_frame_t* frames, int B, float* out, int H, int W,
                                        const float mean[3], vd_stream_t stream) {
  VD_REQUIRE(frames && out && mean, "vd_ingest_u8hwc_nchw_f32: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 65535 && B <= 65535, "vd_ingest_u8hwc_nchw_f32: bad extent");
  dim3 grid(vd_cdiv(W, 256), H, B);
  ingest_nchw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(frames, out, H, W, mean[0], mean[1],
                                                                           mean[2]);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
