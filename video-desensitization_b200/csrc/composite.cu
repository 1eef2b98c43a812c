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
// applied ONE AFTER ANOTHER — a later box reads the pixels earlier boxes wrote.
//
// Parallel schedule.  Every box contributes work items in a fixed order: "A" tiles (frame -> scratch:
// Gaussian horizontal pass / mosaic down-sample gather), then "B" tiles (scratch -> frame: vertical
// pass / mosaic expand); a Gaussian box of at most FUSE_MAX_PX pixels is ONE fused item that runs
// both passes in shared memory.  A persistent grid takes tickets from an atomic counter; ticket t
// is item t / B of frame t % B, so the box chains of all frames advance together.  Dependencies are
// tracked PER TILE with one flag per item:
//   * an A (or fused) tile waits for the B/fused tiles of EARLIER boxes whose write rectangle
//     intersects the rectangle it reads;
//   * a B tile waits for the A tiles of its own box that produce its scratch rows (+-radius) and for
//     the A tiles that still have to read the frame pixels it is about to overwrite.
// Every dependency is an item with a smaller index in the same frame, hence a smaller ticket, hence
// already running or finished: the waits cannot deadlock (and are bounded by a watchdog bit).
// Frame and scratch reads use ld.global.cg because other SMs rewrite those bytes during the kernel.
#include "common.cuh"

namespace {

#ifndef VD_COMP_MINB
#define VD_COMP_MINB 4            // resident CTAs per SM the run kernel is compiled for (64 registers; A/B on B200: 3 -> 0.676 ms, 4 -> 0.645, 5 -> 0.632 on the bench boxes, single-frame chains 0.046 / 0.055 / 0.058)
#endif
constexpr int kThreads = 256;
constexpr int MA_PIX = 256;               // mosaic phase A: small pixels per tile
constexpr int MB_W = 64, MB_H = 16;       // mosaic phase B tile
constexpr int GA_W = 128, GA_H = 8;       // Gaussian phase A tile (horizontal pass), pixels
constexpr int GB_E = 64, GB_H = 64;       // Gaussian phase B tile (vertical pass): 64 flat u16 elements x 64 rows
constexpr int FUSE_MAX_PX = 2048;         // fused Gaussian box: 3 B (u8) + 6 B (u16) per pixel in smem (18 KB)
constexpr int kMaxDynSmem = 64 * 1024;
constexpr int KW_MAX = 32;                // dp4a fast path: ceil((ksize + 3) / 4) <= KW_MAX  (ksize <= 125)

struct BoxMeta {
  int x1, y1, x2, y2;      // clipped; empty boxes have tA = tB = 0
  int tA, tB;              // fused box: tA = 1, tB = 0
  int sw, sh;              // mosaic: down-sampled size
  unsigned int item_base;  // first item of this box inside its frame
  int fused;
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
  int* done;           // [B][items_cap] one flag per item
  uint16_t* scratch;   // [B][scratch_elems]
  size_t scratch_elems;
  unsigned int items_cap;
};

struct Taps {
  uint16_t q[256];
  // dp4a fast path: taps shifted by a = 0..3 bytes and packed 4 per word: qa[a][j] byte b = q[4j + b - a]
  uint32_t qa[4][KW_MAX];
  int kw;                  // words per shifted tap vector, 0 = fast path not applicable (a tap > 255 or ksize > 125)
};

__host__ __device__ inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ inline size_t items_cap_for(int maxb, size_t scratch_px) { return 4 * static_cast<size_t>(maxb) + scratch_px / 128 + 1024; }

__host__ size_t ws_bytes(int B, int maxb, size_t scratch_px) {
  return al256(sizeof(unsigned int) * (B + 4)) + al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb) +
         al256(sizeof(int) * static_cast<size_t>(B) * items_cap_for(maxb, scratch_px)) +
         al256(static_cast<size_t>(B) * scratch_px * 3 * 2) + 256;
}

__host__ Ws carve(void* p0, int B, int maxb, size_t total_bytes, size_t* flag_bytes) {
  Ws w;
  uint8_t* p = static_cast<uint8_t*>(p0);
  w.hdr = reinterpret_cast<Header*>(p);
  p += al256(sizeof(unsigned int) * (B + 4));
  w.meta = reinterpret_cast<BoxMeta*>(p);
  p += al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb);
  // split what is left: flags get 1 int per 128 scratch pixels (+ 4 per box), the rest is scratch
  const size_t left = total_bytes - static_cast<size_t>(p - static_cast<uint8_t*>(p0)) - 256;
  const size_t fixed = sizeof(int) * static_cast<size_t>(B) * (4 * static_cast<size_t>(maxb) + 1024);
  // scratch_px * (6 + 4/128) * B + fixed = left
  const size_t scratch_px = static_cast<size_t>((left > fixed ? left - fixed : 0) / (static_cast<double>(B) * (6.0 + 4.0 / 128.0)));
  w.items_cap = static_cast<unsigned int>(4 * static_cast<size_t>(maxb) + scratch_px / 128 + 1024 - 8);
  w.done = reinterpret_cast<int*>(p);
  *flag_bytes = sizeof(int) * static_cast<size_t>(B) * w.items_cap;
  p += al256(*flag_bytes);
  w.scratch = reinterpret_cast<uint16_t*>(p);
  const size_t used = static_cast<size_t>(p - static_cast<uint8_t*>(p0));
  w.scratch_elems = ((total_bytes - used) / 2 / B) & ~static_cast<size_t>(7);
  return w;
}

// ---------------------------------------------------------------------------------------------
// plan: clip boxes, count tiles, prefix sums.  One CTA per frame; frame_prefix[f+1] holds the item COUNT of
// frame f (the run kernel only needs per-frame counts).
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
    if (blockIdx.x == 0) {
      ws.hdr->ticket = 0;
      ws.hdr->frame_prefix[0] = 0;
    }
  }
  __syncthreads();
  {
    const int f = blockIdx.x;
    const vd_frame_t fr = frames[f];
    const int nd = det_rows ? min(det_count[f], max_det) : 0;
    const int ne = ext_boxes ? min(ext_count[f], max_ext) : 0;
    const int nbx = min(nd + ne, maxb);
    unsigned int item_carry = 0;
    unsigned long long scr_carry = 0;
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
          } else if (static_cast<long long>(bw) * bh <= FUSE_MAX_PX) {
            m.fused = 1; m.tA = 1; m.tB = 0;
          } else {
            m.tA = ((bw + GA_W - 1) / GA_W) * ((bh + GA_H - 1) / GA_H);
            m.tB = ((bw * 3 + GB_E - 1) / GB_E) * ((bh + GB_H - 1) / GB_H);
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
      }
      item_carry += s_items[1023];
      scr_carry += s_scr[1023];
      __syncthreads();
    }
    // a frame whose boxes do not fit the workspace is skipped entirely (status bit)
    const bool overflow = scr_carry > ws.scratch_elems || item_carry > ws.items_cap;
    if (overflow) {
      for (int k = tid; k < maxb; k += 1024) {
        BoxMeta* mp = ws.meta + static_cast<size_t>(f) * maxb + k;
        mp->tA = 0; mp->tB = 0; mp->item_base = 0;
      }
    }
    if (tid == 0) {
      if (overflow || nd + ne > maxb) atomicOr(status, VD_STAT_WORK_OVERFLOW);
      ws.hdr->frame_prefix[f + 1] = overflow ? 0u : item_carry;       // item count of this frame
    }
  }
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

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.global.acquire.gpu.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.global.release.gpu.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void spin_flag(const int* flag, uint32_t* status) {
  uint32_t spins = 0;
  while (ld_acquire(flag) == 0) {
    if (++spins > 64) __nanosleep(32);
    if (spins > (1u << 24)) {
      atomicOr(status, VD_STAT_WATCHDOG);
      break;
    }
  }
}

// wait until every B / fused tile of an earlier box that writes into [rx1,rx2) x [ry1,ry2) is done
__device__ __forceinline__ void wait_writers(const BoxMeta* fm, const int* done, int k, int rx1, int ry1, int rx2,
                                             int ry2, int mode, uint32_t* status) {
  for (int j = threadIdx.x; j < k; j += kThreads) {
    const BoxMeta o = fm[j];
    if (o.tA + o.tB == 0) continue;
    const int ix1 = max(rx1, o.x1), iy1 = max(ry1, o.y1), ix2 = min(rx2, o.x2), iy2 = min(ry2, o.y2);
    if (ix1 >= ix2 || iy1 >= iy2) continue;
    if (o.fused) {
      spin_flag(done + o.item_base, status);
      continue;
    }
    int tiles_x, tx0, tx1, th;
    if (mode == VD_MODE_MOSAIC) {
      th = MB_H;
      tiles_x = (o.x2 - o.x1 + MB_W - 1) / MB_W;
      tx0 = (ix1 - o.x1) / MB_W; tx1 = (ix2 - 1 - o.x1) / MB_W;
    } else {                      // Gaussian B tiles span GB_E flat (pixel*3 + channel) elements
      th = GB_H;
      tiles_x = ((o.x2 - o.x1) * 3 + GB_E - 1) / GB_E;
      tx0 = ((ix1 - o.x1) * 3) / GB_E; tx1 = ((ix2 - o.x1) * 3 - 1) / GB_E;
    }
    const int ty0 = (iy1 - o.y1) / th, ty1 = (iy2 - 1 - o.y1) / th;
    const int* base = done + o.item_base + o.tA;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) spin_flag(base + ty * tiles_x + tx, status);
  }
}

// ---- dp4a building blocks --------------------------------------------------------------------
// out[4m + a] = sum_j dp4a(S[m + j], QA_a[j]) for M consecutive m; S = byte row packed 4 per word.
template <int M>
__device__ __forceinline__ void dp4a_run(const uint32_t* __restrict__ S, const uint32_t* __restrict__ qa, int kw,
                                         uint32_t (&acc)[M]) {
  // register window of M + 3 words, advanced four tap-words per iteration (the tap vectors are zero padded
  // up to KW_MAX, so running kw up to the next multiple of four is harmless)
  uint32_t win[M + 3];
#pragma unroll
  for (int i = 0; i < M; ++i) acc[i] = 0;
#pragma unroll
  for (int i = 0; i < M + 3; ++i) win[i] = S[i];
  for (int j = 0; j < kw; j += 4) {
    const uint32_t q0 = qa[j], q1 = qa[j + 1], q2 = qa[j + 2], q3 = qa[j + 3];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      acc[i] = __dp4a(win[i], q0, acc[i]);
      acc[i] = __dp4a(win[i + 1], q1, acc[i]);
      acc[i] = __dp4a(win[i + 2], q2, acc[i]);
      acc[i] = __dp4a(win[i + 3], q3, acc[i]);
    }
#pragma unroll
    for (int i = 0; i + 4 < M + 3; ++i) win[i] = win[i + 4];
#pragma unroll
    for (int i = 0; i < 4; ++i) win[M - 1 + i] = S[M + 3 + j + i];
  }
}

// Both passes of a fused (whole box in shared memory) Gaussian item with the taps in registers: KT >= ksize taps,
// the ones beyond ksize are zero and read table entry 0.  out = (sum_j q[j] * (sum_i q[i] * src[refl]) + 32768) >> 16.
template <int KT>
__device__ __forceinline__ void fused_passes(const uint8_t* __restrict__ src, uint16_t* __restrict__ mid,
                                             const uint16_t* __restrict__ xtab, const uint16_t* __restrict__ ytab,
                                             const uint16_t* __restrict__ s_q, int bw, int bh, uint8_t* img, int pitch,
                                             int x1, int y1) {
  const int tid = threadIdx.x;
  const int row3 = bw * 3, n = bh * row3;
  uint32_t q[KT];
#pragma unroll
  for (int t = 0; t < KT; ++t) q[t] = s_q[t];
  for (int i = tid; i < n; i += kThreads) {
    const int y = i / row3, e = i - y * row3;
    const int x = e / 3, c = e - x * 3;
    const uint8_t* srow = src + y * row3 + c;
    const uint16_t* tb = xtab + x;
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < KT; ++t) acc += q[t] * srow[tb[t]];
    mid[i] = static_cast<uint16_t>(min(acc, 65535u));
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    const int y = i / row3, e = i - y * row3;
    const uint16_t* tb = ytab + y;
    const uint16_t* col = mid + e;
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < KT; ++t) acc += q[t] * col[tb[t] * row3];
    img[static_cast<size_t>(y1 + y) * pitch + x1 * 3 + e] = static_cast<uint8_t>((acc + 32768u) >> 16);
  }
}

__global__ void __launch_bounds__(kThreads, VD_COMP_MINB)
composite_run_kernel(const vd_frame_t* __restrict__ frames, int B, int maxb, int mode, int ksize, Taps taps,
                     uint32_t* __restrict__ status, Ws ws) {
  extern __shared__ __align__(16) uint8_t dsm[];
  __shared__ unsigned int s_ticket;
  __shared__ int s_box;
  __shared__ unsigned int s_fprefix[130];
  __shared__ uint16_t s_q[256];
  __shared__ uint32_t s_qa[4][KW_MAX];
  const int tid = threadIdx.x;
  for (int i = tid; i < B; i += kThreads) s_fprefix[i] = ws.hdr->frame_prefix[i + 1];   // items of frame i
  s_q[tid] = taps.q[tid];
  if (tid < 4 * KW_MAX) s_qa[tid / KW_MAX][tid % KW_MAX] = taps.qa[tid / KW_MAX][tid % KW_MAX];
  __syncthreads();
  unsigned int max_items = 0;
  for (int i = 0; i < B; ++i) max_items = max(max_items, s_fprefix[i]);
  const unsigned int total = max_items * static_cast<unsigned int>(B);
  const int r = ksize >> 1;
  const int kw = taps.kw;

  while (true) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&ws.hdr->ticket, 1u);
    __syncthreads();
    const unsigned int t = s_ticket;
    if (t >= total) break;
    const int f = static_cast<int>(t % static_cast<unsigned int>(B));
    const unsigned int local = t / static_cast<unsigned int>(B);
    if (local >= s_fprefix[f]) continue;                           // this frame has fewer items
    const BoxMeta* fm = ws.meta + static_cast<size_t>(f) * maxb;
    int* done = ws.done + static_cast<size_t>(f) * ws.items_cap;
    // owner box of the item: warp-parallel search over the (non-decreasing) item_base array — the last
    // box with item_base <= local that actually owns items
    if (tid < 32) {
      int lo = 0, hi = maxb;                       // answer in [lo, hi)
      while (hi - lo > 32) {
        const int step = (hi - lo + 31) / 32;
        const int idx = min(lo + tid * step, hi - 1);
        const bool le = fm[idx].item_base <= local;
        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, le);
        const int last = 31 - __clz(bal);          // lane 0 always qualifies (item_base[lo] <= local)
        const int nlo = min(lo + last * step, hi - 1);
        hi = min(nlo + step, hi);
        lo = nlo;
      }
      const int idx = lo + tid;
      bool own = false;
      if (idx < hi) {
        const BoxMeta c = fm[idx];
        own = c.item_base <= local && local < c.item_base + static_cast<unsigned int>(c.tA + c.tB);
      }
      const unsigned int bal = __ballot_sync(0xFFFFFFFFu, own);
      if (tid == 0) s_box = bal ? lo + (__ffs(bal) - 1) : -1;
    }
    __syncthreads();
    int k = s_box;
    if (k < 0) {                                   // owner sits before `lo` (run of empty boxes): walk back
      k = 0;
      int lo2 = 0, hi2 = maxb - 1;
      while (lo2 < hi2) {
        const int mid = (lo2 + hi2 + 1) >> 1;
        if (fm[mid].item_base <= local) lo2 = mid; else hi2 = mid - 1;
      }
      k = lo2;
      while (k > 0 && !(local < fm[k].item_base + static_cast<unsigned int>(fm[k].tA + fm[k].tB))) --k;
    }
    const BoxMeta m = fm[k];
    const int rel = static_cast<int>(local - m.item_base);
    const bool phaseA = rel < m.tA;
    const int tile = phaseA ? rel : rel - m.tA;
    const int bw = m.x2 - m.x1, bh = m.y2 - m.y1;
    const int row3 = bw * 3;

    // ---- tile geometry (box-local) and dependencies ----
    int tx = 0, ty = 0;
    if (mode == VD_MODE_MOSAIC) {
      if (phaseA) {
        wait_writers(fm, done, k, m.x1, m.y1, m.x2, m.y2, mode, status);
      } else {
        const int tiles_x = (bw + MB_W - 1) / MB_W;
        ty = tile / tiles_x; tx = tile - ty * tiles_x;
        for (int a = tid; a < m.tA; a += kThreads) spin_flag(done + m.item_base + a, status);
      }
    } else if (m.fused) {
      wait_writers(fm, done, k, m.x1, m.y1, m.x2, m.y2, mode, status);
    } else if (phaseA) {
      const int tiles_x = (bw + GA_W - 1) / GA_W;
      ty = tile / tiles_x; tx = tile - ty * tiles_x;
      const int cx1 = max(0, tx * GA_W - r), cx2 = min(bw, tx * GA_W + GA_W + r);
      const int cy1 = ty * GA_H, cy2 = min(bh, ty * GA_H + GA_H);
      wait_writers(fm, done, k, m.x1 + cx1, m.y1 + cy1, m.x1 + cx2, m.y1 + cy2, mode, status);
    } else {
      const int tiles_x = (row3 + GB_E - 1) / GB_E;
      ty = tile / tiles_x; tx = tile - ty * tiles_x;
      // A tiles of this box with rows in (tile rows +- r) and pixel columns in (tile columns +- r)
      const int px1 = (tx * GB_E) / 3, px2 = min(bw, (tx * GB_E + GB_E + 2) / 3);
      const int cx1 = max(0, px1 - r), cx2 = min(bw, px2 + r);
      const int cy1 = max(0, ty * GB_H - r), cy2 = min(bh, ty * GB_H + GB_H + r);
      const int atx = (bw + GA_W - 1) / GA_W;
      const int ax0 = cx1 / GA_W, ax1 = (cx2 - 1) / GA_W, ay0 = cy1 / GA_H, ay1 = (cy2 - 1) / GA_H;
      const int nax = ax1 - ax0 + 1, n = nax * (ay1 - ay0 + 1);
      for (int i = tid; i < n; i += kThreads)
        spin_flag(done + m.item_base + (ay0 + i / nax) * atx + ax0 + i % nax, status);
    }
    // flag acquire (per thread) -> bar.sync -> ld.global.cg by every thread of the CTA: the barrier extends the
    // acquiring thread's view to the whole CTA, no fence (and no L1 invalidation) needed
    __syncthreads();

    const vd_frame_t fr = frames[f];
    uint8_t* img = const_cast<uint8_t*>(fr.ptr);
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
    } else if (m.fused) {
      // whole box in shared memory: u8 source [bh][bw*3], u16 horizontal result [bh][bw*3], and the
      // REFLECT_101 index tables of both axes (a small box is all border: bw, bh < ksize is the common case)
      const int n = bh * row3;
      uint8_t* src = dsm;
      uint16_t* mid = reinterpret_cast<uint16_t*>(dsm + ((n + 15) & ~15));
      uint16_t* xtab = mid + ((n + 7) & ~7);                 // [bw + 2r + 32]: 3 * reflect(i - r), then zeros
      uint16_t* ytab = xtab + ((bw + 2 * r + 32 + 7) & ~7);  // [bh + 2r + 32]:     reflect(i - r), then zeros
      // loads are issued in batches of 8 per thread before any of them is consumed: the item is latency-bound
      for (int base = 0; base < n; base += 8 * kThreads) {
        uint8_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * kThreads + tid;
          if (i < n) {
            const int y = i / row3, e = i - y * row3;
            v[u] = __ldcg(img + static_cast<size_t>(m.y1 + y) * fr.pitch + m.x1 * 3 + e);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * kThreads + tid;
          if (i < n) src[i] = v[u];
        }
      }
      for (int i = tid; i < bw + 2 * r + 32; i += kThreads)
        xtab[i] = static_cast<uint16_t>(i < bw + 2 * r ? reflect101(i - r, bw) * 3 : 0);
      for (int i = tid; i < bh + 2 * r + 32; i += kThreads)
        ytab[i] = static_cast<uint16_t>(i < bh + 2 * r ? reflect101(i - r, bh) : 0);
      __syncthreads();
      if (ksize <= 8) fused_passes<8>(src, mid, xtab, ytab, s_q, bw, bh, img, fr.pitch, m.x1, m.y1);
      else if (ksize <= 16) fused_passes<16>(src, mid, xtab, ytab, s_q, bw, bh, img, fr.pitch, m.x1, m.y1);
      else if (ksize <= 32) fused_passes<32>(src, mid, xtab, ytab, s_q, bw, bh, img, fr.pitch, m.x1, m.y1);
      else {
        for (int i = tid; i < n; i += kThreads) {
          const int y = i / row3, e = i - y * row3;
          const int x = e / 3, c = e - x * 3;
          const uint8_t* srow = src + y * row3 + c;
          const uint16_t* tb = xtab + x;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * srow[tb[tt]];
          mid[i] = static_cast<uint16_t>(min(acc, 65535u));
        }
        __syncthreads();
        for (int i = tid; i < n; i += kThreads) {
          const int y = i / row3, e = i - y * row3;
          const uint16_t* tb = ytab + y;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * mid[tb[tt] * row3 + e];
          img[static_cast<size_t>(m.y1 + y) * fr.pitch + m.x1 * 3 + e] = static_cast<uint8_t>((acc + 32768u) >> 16);
        }
      }
    } else if (phaseA) {
      // ---------------- horizontal pass: GA_H rows x GA_W pixels ----------------
      const int x0 = tx * GA_W, y0 = ty * GA_H;
      const int rows = min(GA_H, bh - y0);
      const int cols = min(GA_W, bw - x0);
      if (kw > 0) {
        // planar staging: plane c, row ly: bytes s[i] = pixel(x0 - r + i) (REFLECT_101), word-aligned rows
        const int span = GA_W + 2 * r;                       // bytes used per plane row
        const int spanw = (span + 3 + 4 * KW_MAX) / 4;       // words per row incl. slack read by the last window
        uint32_t* planes = reinterpret_cast<uint32_t*>(dsm); // [3][GA_H][spanw]
        uint8_t* pb = dsm;
        const int total = rows * span;
        for (int base = 0; base < total; base += 6 * kThreads) {       // 6 x 3 loads in flight per thread
          uint8_t v[6][3];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int i = base + u * kThreads + tid;
            if (i < total) {
              const int ly = i / span, c = i - ly * span;
              int xs = x0 - r + c;
              if (xs < 0 || xs >= bw) xs = reflect101(xs, bw);
              const uint8_t* p = img + static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + (m.x1 + xs) * 3;
              v[u][0] = __ldcg(p);
              v[u][1] = __ldcg(p + 1);
              v[u][2] = __ldcg(p + 2);
            }
          }
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int i = base + u * kThreads + tid;
            if (i < total) {
              const int ly = i / span, c = i - ly * span;
              pb[(0 * GA_H + ly) * spanw * 4 + c] = v[u][0];
              pb[(1 * GA_H + ly) * spanw * 4 + c] = v[u][1];
              pb[(2 * GA_H + ly) * spanw * 4 + c] = v[u][2];
            }
          }
        }
        __syncthreads();
        // thread -> (channel, a, row, half): 16 outputs x = 64*half + 4m + a
        constexpr int M = 16;
        if (tid < 3 * 4 * GA_H * 2) {
          const int c = tid % 3, a = (tid / 3) & 3, rest = tid / 12;
          const int half = rest & 1, ly = rest >> 1;
          if (ly < rows) {
            const uint32_t* S = planes + (c * GA_H + ly) * spanw + half * (GA_W / 8);
            uint32_t acc[M];
            dp4a_run<M>(S, s_qa[a], kw, acc);
            uint16_t* orow = scr16 + (static_cast<size_t>(y0 + ly) * bw + x0) * 3 + c;
#pragma unroll
            for (int i = 0; i < M; ++i) {
              const int lx = half * (GA_W / 2) + 4 * i + a;
              if (lx < cols) orow[lx * 3] = static_cast<uint16_t>(acc[i]);
            }
          }
        }
      } else {
        const int span = GA_W + 2 * r;
        uint8_t* st = dsm;                             // [GA_H][span*3]
        for (int i = tid; i < rows * span; i += kThreads) {
          const int ly = i / span, c = i - ly * span;
          const int xs = reflect101(x0 - r + c, bw);
          const uint8_t* p = img + static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + (m.x1 + xs) * 3;
          uint8_t* d = st + (ly * span + c) * 3;
          d[0] = __ldcg(p); d[1] = __ldcg(p + 1); d[2] = __ldcg(p + 2);
        }
        __syncthreads();
        for (int i = tid; i < rows * cols * 3; i += kThreads) {
          const int ly = i / (cols * 3), e = i - ly * cols * 3;      // e = lx*3 + ch
          const uint8_t* s = st + ly * span * 3 + e;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * s[tt * 3];
          scr16[(static_cast<size_t>(y0 + ly) * bw + x0) * 3 + e] = static_cast<uint16_t>(min(acc, 65535u));
        }
      }
    } else {
      // ---------------- vertical pass: GB_H rows x GB_E flat elements ----------------
      const int e0 = tx * GB_E, y0 = ty * GB_H;
      const int ne = min(GB_E, row3 - e0);
      const int rows = min(GB_H, bh - y0);
      const int span = rows + 2 * r;                       // staged rows
      if (kw > 0) {
        // transposed byte planes: hi[e][i], lo[e][i] with i = staged row (REFLECT_101); word-aligned columns;
        // an odd word pitch keeps the byte stores of consecutive e on different banks
        int colw = (GB_H + 2 * r + 3 + 4 * KW_MAX) / 4;
        colw |= 1;
        uint8_t* hi = dsm;                                  // [GB_E][colw*4]
        uint8_t* lo = dsm + GB_E * colw * 4;
        {
          const int e = tid & (GB_E - 1);
          if (e < ne) {
            uint8_t* hcol = hi + e * colw * 4;
            uint8_t* lcol = lo + e * colw * 4;
            constexpr int LSTEP = kThreads / GB_E;
            for (int l0 = tid / GB_E; l0 < span; l0 += 8 * LSTEP) {     // 8 loads in flight per thread
              uint16_t v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int ly = l0 + u * LSTEP;
                if (ly < span) {
                  int ys = y0 - r + ly;
                  if (ys < 0 || ys >= bh) ys = reflect101(ys, bh);
                  v[u] = __ldcg(scr16 + static_cast<size_t>(ys) * row3 + e0 + e);
                }
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int ly = l0 + u * LSTEP;
                if (ly < span) {
                  hcol[ly] = static_cast<uint8_t>(v[u] >> 8);
                  lcol[ly] = static_cast<uint8_t>(v[u] & 0xFF);
                }
              }
            }
          }
        }
        __syncthreads();
        // thread -> (e, a): 16 outputs y = 4m + a
        constexpr int M = GB_H / 4;
        const int e = tid & (GB_E - 1), a = tid / GB_E;
        if (e < ne) {
          uint32_t ah[M], al[M];
          dp4a_run<M>(reinterpret_cast<const uint32_t*>(hi) + e * colw, s_qa[a], kw, ah);
          dp4a_run<M>(reinterpret_cast<const uint32_t*>(lo) + e * colw, s_qa[a], kw, al);
          uint8_t* ocol = img + static_cast<size_t>(m.y1 + y0) * fr.pitch + m.x1 * 3 + e0 + e;
#pragma unroll
          for (int i = 0; i < M; ++i) {
            const int ly = 4 * i + a;
            if (ly < rows) ocol[static_cast<size_t>(ly) * fr.pitch] = static_cast<uint8_t>(((ah[i] << 8) + al[i] + 32768u) >> 16);
          }
        }
      } else {
        uint16_t* st = reinterpret_cast<uint16_t*>(dsm);   // [span][GB_E]
        for (int i = tid; i < span * ne; i += kThreads) {
          const int ly = i / ne, e = i - ly * ne;
          const int ys = reflect101(y0 - r + ly, bh);
          st[ly * GB_E + e] = __ldcg(scr16 + static_cast<size_t>(ys) * row3 + e0 + e);
        }
        __syncthreads();
        for (int i = tid; i < rows * ne; i += kThreads) {
          const int ly = i / ne, e = i - ly * ne;
          const uint16_t* s = st + ly * GB_E + e;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * s[tt * GB_E];
          img[static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + m.x1 * 3 + e0 + e] =
              static_cast<uint8_t>((acc + 32768u) >> 16);
        }
      }
    }
    // every thread's stores -> bar.sync -> st.release.gpu by one thread (release is cumulative over what the
    // barrier ordered before it): no fence needed
    __syncthreads();
    if (tid == 0) st_release(done + local, 1);
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
  if (max_boxes > VD_MAX_BOXES_PER_FRAME) max_boxes = VD_MAX_BOXES_PER_FRAME;
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
    // dp4a fast path: every tap must fit a byte and the shifted vectors KW_MAX words
    const int kwords = (ksize + 3 + 3) / 4;
    bool fits = kwords <= KW_MAX;
    for (int i = 0; i < ksize; ++i) fits = fits && taps_q8[i] <= 255;
    if (fits) {
      taps.kw = kwords;
      for (int a = 0; a < 4; ++a)
        for (int j = 0; j < KW_MAX; ++j) {
          uint32_t w = 0;
          for (int b = 0; b < 4; ++b) {
            const int ti = 4 * j + b - a;
            if (ti >= 0 && ti < ksize) w |= static_cast<uint32_t>(taps_q8[ti]) << (8 * b);
          }
          taps.qa[a][j] = w;
        }
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t flag_bytes = 0;
  const Ws ws = carve(workspace, B, maxb, workspace_bytes, &flag_bytes);
  VD_CHECK_CUDA(cudaMemsetAsync(ws.done, 0, flag_bytes, st));
  composite_plan_kernel<<<B, 1024, 0, st>>>(frames, B, det_rows, det_count, max_det, ext_boxes, ext_count, max_ext,
                                            maxb, mode, mosaic_level, status, ws);
  VD_LAUNCH_CHECK();
  const int r = ksize >> 1;
  size_t smem = 0;
  if (mode == VD_MODE_GAUSS) {
    size_t a = static_cast<size_t>(GA_H) * (GA_W + 2 * r) * 3;
    size_t b = static_cast<size_t>(GB_H + 2 * r) * GB_E * 2;
    if (taps.kw > 0) {
      a = static_cast<size_t>(3) * GA_H * ((GA_W + 2 * r + 3 + 4 * KW_MAX) / 4) * 4;
      b = static_cast<size_t>(2) * GB_E * ((((GB_H + 2 * r + 3 + 4 * KW_MAX) / 4) | 1)) * 4;
    }
    // fused item: u8 source + u16 intermediate + the two reflect tables (bw + bh <= FUSE_MAX_PX + 1)
    const size_t fz = static_cast<size_t>(FUSE_MAX_PX) * 9 + 64 + 2 * (static_cast<size_t>(FUSE_MAX_PX) + 1 + 4 * r + 64 + 16);
    smem = a > b ? a : b;
    if (fz > smem) smem = fz;
  }
  VD_REQUIRE(smem <= static_cast<size_t>(kMaxDynSmem), "vd_composite_rois: kernel size too large for shared memory");
  VD_SET_FUNC_ATTR_ONCE_PER_DEVICE(composite_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  int per_sm = 4;
  if (smem > 0) {
    const int fit = static_cast<int>((220 * 1024) / (smem + 2048));
    const int cap = VD_COMP_MINB > 4 ? VD_COMP_MINB : 4;
    per_sm = fit < 1 ? 1 : (fit > cap ? cap : fit);
  }
  composite_run_kernel<<<sms * per_sm, kThreads, smem, st>>>(frames, B, maxb, mode, ksize, taps, status, ws);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
