// composite.cu — K5: sequential mosaic / Gaussian-blur compositing over boxes, in place, all
// frames of a batch in two launches and without a host round trip.
//
// Replaces the per-frame box loop of the reference frame loop
//   all_boxes = int(face boxes) + int(plate boxes); for box: img = mosaic_rectangle_region_single(img, *box)
//   (reference combine_detect.py:241-249, 138-161)
// and, for the Gaussian mode that BASELINE.json's north_star adds, the idiom
//   img[y1:y2, x1:x2] = cv2.GaussianBlur(img[y1:y2, x1:x2], (k,k), sigma)
// with cv2's uint8 fixed-point arithmetic (8.8 taps, u16 horizontal pass, u32 vertical pass,
// (v + 32768) >> 16, BORDER_REFLECT_101 inside the ROI).
//
// Semantics kept: boxes are truncated toward zero, clipped to the frame, skipped when empty, and
// applied ONE AFTER ANOTHER — a later box reads the pixels earlier boxes wrote.  Parallelism comes
// from a ticket-ordered persistent kernel: every box contributes "phase A" tiles (frame -> scratch:
// mosaic = gather the down-sampled image, Gaussian = horizontal pass) and "phase B" tiles (scratch
// -> frame: mosaic = expand, Gaussian = vertical pass).  Tickets are handed out by an atomic in
// (frame, box, phase, tile) order; a tile waits until the phase-B tiles of every EARLIER OVERLAPPING
// box (phase A) or the phase-A tiles of its own box (phase B) are done.  Dependencies only point to
// smaller tickets, which are always already running or finished, so the wait cannot deadlock; it is
// also bounded (watchdog bit in `status`).  Frame and scratch reads use ld.global.cg because other
// SMs rewrite those bytes during the kernel (L1 is not coherent).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int MA_PIX = 256;               // mosaic phase A: small pixels per tile
constexpr int MB_W = 64, MB_H = 16;       // mosaic phase B tile
constexpr int GA_W = 128, GA_H = 8;       // Gaussian phase A tile (horizontal pass)
constexpr int GB_W = 32, GB_H = 32;       // Gaussian phase B tile (vertical pass)

struct BoxMeta {
  int x1, y1, x2, y2;      // clipped; empty boxes have tA = tB = 0
  int tA, tB;
  int sw, sh;              // mosaic: down-sampled size
  unsigned int item_base;  // first ticket of this box inside its frame
  unsigned int pad;
  unsigned long long scratch_off;   // in u16 elements inside the frame's scratch region
};

struct Header {
  unsigned int ticket;
  unsigned int total_items;
  unsigned int frame_prefix[1];     // [B+1] follows
};

struct Ws {
  Header* hdr;
  BoxMeta* meta;       // [B][maxb]
  int* doneA;          // [B][maxb]
  int* doneB;          // [B][maxb]
  uint16_t* scratch;   // [B][scratch_elems]
  size_t scratch_elems;
};

struct Taps {
  uint16_t q[256];
};

__host__ __device__ inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ size_t ws_bytes(int B, int maxb, size_t scratch_px) {
  return al256(sizeof(unsigned int) * (B + 4)) + al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb) +
         2 * al256(sizeof(int) * static_cast<size_t>(B) * maxb) + al256(static_cast<size_t>(B) * scratch_px * 3 * 2) +
         256;
}

__host__ Ws carve(void* p0, int B, int maxb, size_t total_bytes) {
  Ws w;
  uint8_t* p = static_cast<uint8_t*>(p0);
  w.hdr = reinterpret_cast<Header*>(p);
  p += al256(sizeof(unsigned int) * (B + 4));
  w.meta = reinterpret_cast<BoxMeta*>(p);
  p += al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb);
  w.doneA = reinterpret_cast<int*>(p);
  p += al256(sizeof(int) * static_cast<size_t>(B) * maxb);
  w.doneB = reinterpret_cast<int*>(p);
  p += al256(sizeof(int) * static_cast<size_t>(B) * maxb);
  w.scratch = reinterpret_cast<uint16_t*>(p);
  const size_t used = static_cast<size_t>(p - static_cast<uint8_t*>(p0));
  w.scratch_elems = ((total_bytes - used) / 2 / B) & ~static_cast<size_t>(7);
  return w;
}

// ---------------------------------------------------------------------------------------------
// plan: clip boxes, count tiles, prefix sums.  One CTA; frames are walked sequentially.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
composite_plan_kernel(const vd_frame_t* __restrict__ frames, int B, const float* __restrict__ det_rows,
                      const int* __restrict__ det_count, int max_det, const int* __restrict__ ext_boxes,
                      const int* __restrict__ ext_count, int max_ext, int maxb, int mode, int mosaic_level,
                      uint32_t* __restrict__ status, Ws ws) {
  __shared__ unsigned int s_items[1024];
  __shared__ unsigned long long s_scr[1024];
  __shared__ unsigned int s_total;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_total = 0;
    ws.hdr->ticket = 0;
    ws.hdr->frame_prefix[0] = 0;
  }
  __syncthreads();
  for (int f = 0; f < B; ++f) {
    const vd_frame_t fr = frames[f];
    const int nd = det_rows ? min(det_count[f], max_det) : 0;
    const int ne = ext_boxes ? min(ext_count[f], max_ext) : 0;
    const int nbx = min(nd + ne, maxb);
    unsigned int item_carry = 0;
    unsigned long long scr_carry = 0;
    // pass 1: metadata + prefix sums, 1024 boxes per chunk
    for (int base = 0; base < maxb; base += 1024) {
      const int k = base + tid;
      BoxMeta m;
      memset(&m, 0, sizeof(m));
      unsigned int items = 0;
      unsigned long long scr = 0;
      if (k < nbx) {
        int x1, y1, x2, y2;
        if (k < nd) {
          // face.py:144-146: result[:, :4] *= [w, h, w, h] in float32; combine_detect.py:243: int() truncation
          const float* r = det_rows + (static_cast<size_t>(f) * max_det + k) * 16;
          x1 = static_cast<int>(__fmul_rn(r[0], static_cast<float>(fr.w)));
          y1 = static_cast<int>(__fmul_rn(r[1], static_cast<float>(fr.h)));
          x2 = static_cast<int>(__fmul_rn(r[2], static_cast<float>(fr.w)));
          y2 = static_cast<int>(__fmul_rn(r[3], static_cast<float>(fr.h)));
        } else {
          const int* e = ext_boxes + (static_cast<size_t>(f) * max_ext + (k - nd)) * 4;
          x1 = e[0]; y1 = e[1]; x2 = e[2]; y2 = e[3];
        }
        x1 = max(0, x1); y1 = max(0, y1); x2 = min(fr.w, x2); y2 = min(fr.h, y2);   // combine_detect.py:145-148
        m.x1 = x1; m.y1 = y1; m.x2 = x2; m.y2 = y2;
        if (x2 > x1 && y2 > y1) {                                                   // :150
          const int bw = x2 - x1, bh = y2 - y1;
          if (mode == VD_MODE_MOSAIC) {
            m.sw = max(1, bw / mosaic_level);
            m.sh = max(1, bh / mosaic_level);
            m.tA = (m.sw * m.sh + MA_PIX - 1) / MA_PIX;
            m.tB = ((bw + MB_W - 1) / MB_W) * ((bh + MB_H - 1) / MB_H);
            scr = (static_cast<unsigned long long>(m.sw) * m.sh * 3 + 1) / 2;
          } else {
            m.tA = ((bw + GA_W - 1) / GA_W) * ((bh + GA_H - 1) / GA_H);
            m.tB = ((bw + GB_W - 1) / GB_W) * ((bh + GB_H - 1) / GB_H);
            scr = static_cast<unsigned long long>(bw) * bh * 3;
          }
          scr = (scr + 7) & ~7ull;
          items = static_cast<unsigned int>(m.tA + m.tB);
        }
      }
      // inclusive scan over the chunk (Hillis-Steele in shared memory)
      s_items[tid] = items;
      s_scr[tid] = scr;
      __syncthreads();
      for (int off = 1; off < 1024; off <<= 1) {
        unsigned int a = 0;
        unsigned long long c = 0;
        if (tid >= off) { a = s_items[tid - off]; c = s_scr[tid - off]; }
        __syncthreads();
        s_items[tid] += a;
        s_scr[tid] += c;
        __syncthreads();
      }
      if (k < maxb) {
        m.item_base = item_carry + s_items[tid] - items;
        m.scratch_off = scr_carry + s_scr[tid] - scr;
        ws.meta[static_cast<size_t>(f) * maxb + k] = m;
        ws.doneA[static_cast<size_t>(f) * maxb + k] = 0;
        ws.doneB[static_cast<size_t>(f) * maxb + k] = 0;
      }
      item_carry += s_items[1023];
      scr_carry += s_scr[1023];
      __syncthreads();
    }
    // pass 2: a frame whose boxes do not fit the scratch region is skipped entirely (status bit)
    const bool overflow = scr_carry > ws.scratch_elems;
    if (overflow) {
      for (int k = tid; k < maxb; k += 1024) {
        BoxMeta* mp = ws.meta + static_cast<size_t>(f) * maxb + k;
        mp->tA = 0; mp->tB = 0; mp->item_base = 0;
      }
    }
    if (tid == 0) {
      if (overflow || nd + ne > maxb) atomicOr(status, VD_STAT_WORK_OVERFLOW);
      s_total += overflow ? 0u : item_carry;
      ws.hdr->frame_prefix[f + 1] = s_total;
    }
    __syncthreads();
  }
  if (tid == 0) ws.hdr->total_items = s_total;
}

// ---------------------------------------------------------------------------------------------
// run
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
  return p;
}

// cv2 INTER_NEAREST source index: min(floor(dst * (1.0 / (dst_n / src_n))), src_n - 1) in double
__device__ __forceinline__ int nearest_src(int d, int dst_n, int src_n) {
  const double inv = 1.0 / (static_cast<double>(dst_n) / static_cast<double>(src_n));
  const int s = static_cast<int>(floor(d * inv));
  return min(s, src_n - 1);
}

__device__ __forceinline__ int ld_volatile(const int* p) {
  int v;
  asm volatile("ld.global.acquire.gpu.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void spin_until(const int* ctr, int target, uint32_t* status) {
  uint32_t spins = 0;
  while (ld_volatile(ctr) < target) {
    __nanosleep(64);
    if (++spins > (1u << 24)) {
      atomicOr(status, VD_STAT_WATCHDOG);
      break;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
composite_run_kernel(const vd_frame_t* __restrict__ frames, int B, int maxb, int mode, int ksize, Taps taps,
                     uint32_t* __restrict__ status, Ws ws) {
  extern __shared__ uint8_t dsm[];
  __shared__ unsigned int s_ticket;
  __shared__ unsigned int s_fprefix[130];
  __shared__ uint16_t s_q[256];
  const int tid = threadIdx.x;
  for (int i = tid; i <= B; i += kThreads) s_fprefix[i] = ws.hdr->frame_prefix[i];
  s_q[tid] = taps.q[tid];
  __syncthreads();
  // tickets interleave the frames (ticket t -> frame t % B, item t / B) so the sequential box chains
  // of all frames advance concurrently; inside a frame the (box, phase, tile) order is kept
  unsigned int max_items = 0;
  for (int i = 0; i < B; ++i) max_items = max(max_items, s_fprefix[i + 1] - s_fprefix[i]);
  const unsigned int total = max_items * static_cast<unsigned int>(B);
  const int r = ksize >> 1;

  while (true) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&ws.hdr->ticket, 1u);
    __syncthreads();
    const unsigned int t = s_ticket;
    if (t >= total) break;
    const int f = static_cast<int>(t % static_cast<unsigned int>(B));
    const unsigned int local = t / static_cast<unsigned int>(B);
    if (local >= s_fprefix[f + 1] - s_fprefix[f]) continue;      // this frame has fewer items
    const BoxMeta* fm = ws.meta + static_cast<size_t>(f) * maxb;
    // binary search: last box with item_base <= local and a non-empty item range containing local
    int lo = 0, hi = maxb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (fm[mid].item_base <= local) lo = mid; else hi = mid - 1;
    }
    // boxes with zero items share item_base with their successor: step back to the owner
    int k = lo;
    while (k > 0 && !(local < fm[k].item_base + static_cast<unsigned int>(fm[k].tA + fm[k].tB))) --k;
    const BoxMeta m = fm[k];
    const int rel = static_cast<int>(local - m.item_base);
    const bool phaseA = rel < m.tA;
    const int tile = phaseA ? rel : rel - m.tA;
    int* doneA = ws.doneA + static_cast<size_t>(f) * maxb;
    int* doneB = ws.doneB + static_cast<size_t>(f) * maxb;

    // ---- dependencies ----
    if (phaseA) {
      for (int j = tid; j < k; j += kThreads) {
        const BoxMeta o = fm[j];
        if (o.tB > 0 && o.x1 < m.x2 && m.x1 < o.x2 && o.y1 < m.y2 && m.y1 < o.y2) spin_until(doneB + j, o.tB, status);
      }
    } else if (tid == 0) {
      spin_until(doneA + k, m.tA, status);
    }
    __threadfence();
    __syncthreads();

    const vd_frame_t fr = frames[f];
    uint8_t* img = const_cast<uint8_t*>(fr.ptr);
    const int bw = m.x2 - m.x1, bh = m.y2 - m.y1;
    uint16_t* scr16 = ws.scratch + static_cast<size_t>(f) * ws.scratch_elems + m.scratch_off;

    if (mode == VD_MODE_MOSAIC) {
      uint8_t* small = reinterpret_cast<uint8_t*>(scr16);
      if (phaseA) {
        const int idx = tile * MA_PIX + tid;
        if (idx < m.sw * m.sh) {
          const int sy = idx / m.sw, sx = idx - sy * m.sw;
          const int yy = m.y1 + nearest_src(sy, m.sh, bh), xx = m.x1 + nearest_src(sx, m.sw, bw);
          const uint8_t* p = img + static_cast<size_t>(yy) * fr.pitch + xx * 3;
          small[idx * 3 + 0] = __ldcg(p);
          small[idx * 3 + 1] = __ldcg(p + 1);
          small[idx * 3 + 2] = __ldcg(p + 2);
        }
      } else {
        const int tiles_x = (bw + MB_W - 1) / MB_W;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        for (int i = tid; i < MB_W * MB_H; i += kThreads) {
          const int ly = i / MB_W, lx = i - ly * MB_W;
          const int y = ty * MB_H + ly, x = tx * MB_W + lx;
          if (y < bh && x < bw) {
            const int sy = nearest_src(y, bh, m.sh), sx = nearest_src(x, bw, m.sw);
            const uint8_t* s = small + (sy * m.sw + sx) * 3;
            uint8_t* d = img + static_cast<size_t>(m.y1 + y) * fr.pitch + (m.x1 + x) * 3;
            d[0] = __ldcg(s);
            d[1] = __ldcg(s + 1);
            d[2] = __ldcg(s + 2);
          }
        }
      }
    } else {
      if (phaseA) {
        // horizontal pass: tile GA_H rows x GA_W cols; stage (GA_W + 2r) columns with REFLECT_101
        const int tiles_x = (bw + GA_W - 1) / GA_W;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x0 = tx * GA_W, y0 = ty * GA_H;
        const int span = GA_W + 2 * r;                 // staged columns
        uint8_t* st = dsm;                             // [GA_H][span*3]
        const int rows = min(GA_H, bh - y0);
        for (int i = tid; i < rows * span; i += kThreads) {
          const int ly = i / span, c = i - ly * span;
          const int xs = reflect101(x0 - r + c, bw);
          const uint8_t* p = img + static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + (m.x1 + xs) * 3;
          uint8_t* d = st + (ly * span + c) * 3;
          d[0] = __ldcg(p); d[1] = __ldcg(p + 1); d[2] = __ldcg(p + 2);
        }
        __syncthreads();
        const int cols = min(GA_W, bw - x0);
        for (int i = tid; i < rows * cols * 3; i += kThreads) {
          const int ly = i / (cols * 3), e = i - ly * cols * 3;      // e = lx*3 + ch
          const uint8_t* s = st + ly * span * 3 + e;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * s[tt * 3];
          scr16[(static_cast<size_t>(y0 + ly) * bw + x0) * 3 + e] = static_cast<uint16_t>(min(acc, 65535u));
        }
      } else {
        // vertical pass: tile GB_H rows x GB_W cols (x3 channels); stage (GB_H + 2r) rows of u16
        const int tiles_x = (bw + GB_W - 1) / GB_W;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x0 = tx * GB_W, y0 = ty * GB_H;
        const int cols3 = min(GB_W, bw - x0) * 3;
        const int rows = min(GB_H, bh - y0);
        const int span = rows + 2 * r;
        uint16_t* st = reinterpret_cast<uint16_t*>(dsm);   // [span][GB_W*3]
        for (int i = tid; i < span * cols3; i += kThreads) {
          const int ly = i / cols3, e = i - ly * cols3;
          const int ys = reflect101(y0 - r + ly, bh);
          st[ly * (GB_W * 3) + e] = __ldcg(scr16 + (static_cast<size_t>(ys) * bw + x0) * 3 + e);
        }
        __syncthreads();
        for (int i = tid; i < rows * cols3; i += kThreads) {
          const int ly = i / cols3, e = i - ly * cols3;
          const uint16_t* s = st + ly * (GB_W * 3) + e;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * s[tt * (GB_W * 3)];
          img[static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + (m.x1 + x0) * 3 + e] =
              static_cast<uint8_t>((acc + 32768u) >> 16);
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(phaseA ? (doneA + k) : (doneB + k), 1);
  }
}

}  // namespace

// cv::getGaussianKernel (float64) + fixed-point quantisation with error diffusion
extern "C" int vd_gaussian_taps_q8(int ksize, double sigma, uint16_t* taps) {
  VD_REQUIRE(taps, "vd_gaussian_taps_q8: null output");
  VD_REQUIRE(ksize >= 1 && ksize <= VD_MAX_KSIZE && (ksize & 1), "vd_gaussian_taps_q8: ksize must be odd in [1,255]");
  double g[VD_MAX_KSIZE];
  static const double small[4][7] = {{1.0},
                                     {0.25, 0.5, 0.25},
                                     {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                     {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
  if (ksize <= 7 && sigma <= 0) {
    for (int i = 0; i < ksize; ++i) g[i] = small[ksize >> 1][i];
  } else {
    const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
      const double x = i - (ksize - 1) * 0.5;
      g[i] = exp(scale2x * x * x);
      sum += g[i];
    }
    for (int i = 0; i < ksize; ++i) g[i] /= sum;
  }
  double err = 0;
  int s = 0;
  const int half = ksize / 2;
  for (int i = 0; i < half; ++i) {
    const double adj = g[i] * 256.0 + err;
    const double q = nearbyint(adj);            // default rounding mode: half to even (cvRound)
    err = adj - q;
    taps[i] = taps[ksize - 1 - i] = static_cast<uint16_t>(q);
    s += 2 * static_cast<int>(q);
  }
  taps[half] = static_cast<uint16_t>(256 - s);
  return VD_OK;
}

extern "C" size_t vd_composite_workspace_bytes(int B, int max_boxes, size_t scratch_pixels_per_frame) {
  if (B <= 0 || max_boxes <= 0) return 0;
  return ws_bytes(B, max_boxes, scratch_pixels_per_frame);
}

extern "C" int vd_composite_rois(const vd_frame_t* frames, int B, const float* det_rows, const int32_t* det_count,
                                 int max_det, const int32_t* ext_boxes, const int32_t* ext_count, int max_ext,
                                 int mode, int mosaic_level, const uint16_t* taps_q8, int ksize, uint32_t* status,
                                 void* workspace, size_t workspace_bytes, vd_stream_t stream) {
  VD_REQUIRE(frames && status && workspace, "vd_composite_rois: null argument");
  VD_REQUIRE(B > 0 && B <= 128, "vd_composite_rois: B=%d must be in [1,128]", B);
  VD_REQUIRE(mode == VD_MODE_MOSAIC || mode == VD_MODE_GAUSS, "vd_composite_rois: bad mode");
  VD_REQUIRE((det_rows == nullptr) == (det_count == nullptr), "vd_composite_rois: det_rows/det_count mismatch");
  VD_REQUIRE((ext_boxes == nullptr) == (ext_count == nullptr), "vd_composite_rois: ext_boxes/ext_count mismatch");
  int maxb = (det_rows ? max_det : 0) + (ext_boxes ? max_ext : 0);
  VD_REQUIRE(maxb > 0, "vd_composite_rois: no box source");
  if (maxb > VD_MAX_BOXES_PER_FRAME) maxb = VD_MAX_BOXES_PER_FRAME;
  VD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vd_composite_rois: workspace must be 256-B aligned");
  VD_REQUIRE(workspace_bytes >= ws_bytes(B, maxb, 1024), "vd_composite_rois: workspace too small");
  Taps taps;
  memset(&taps, 0, sizeof(taps));
  if (mode == VD_MODE_MOSAIC) {
    VD_REQUIRE(mosaic_level >= 1, "vd_composite_rois: mosaic_level must be >= 1");
    ksize = 1;
  } else {
    VD_REQUIRE(taps_q8 && ksize >= 1 && ksize <= VD_MAX_KSIZE && (ksize & 1), "vd_composite_rois: bad Gaussian taps");
    for (int i = 0; i < ksize; ++i) taps.q[i] = taps_q8[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Ws ws = carve(workspace, B, maxb, workspace_bytes);
  composite_plan_kernel<<<1, 1024, 0, st>>>(frames, B, det_rows, det_count, max_det, ext_boxes, ext_count, max_ext,
                                            maxb, mode, mosaic_level, status, ws);
  VD_LAUNCH_CHECK();
  const int r = ksize >> 1;
  size_t smem = 0;
  if (mode == VD_MODE_GAUSS) {
    const size_t a = static_cast<size_t>(GA_H) * (GA_W + 2 * r) * 3;
    const size_t b = static_cast<size_t>(GB_H + 2 * r) * GB_W * 3 * 2;
    smem = a > b ? a : b;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VD_CHECK_CUDA(cudaFuncSetAttribute(composite_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  composite_run_kernel<<<sms * 4, kThreads, smem, st>>>(frames, B, maxb, mode, ksize, taps, status, ws);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
