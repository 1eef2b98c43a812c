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
// The plan kernel fixes the ticket order of a frame's items (BoxMeta::posA / posB and the item map): the A tiles of
// the boxes in box order, the B tiles of box j held back behind the A tiles of the next boxes (about one grid's worth
// of tickets, so that they find their A tiles finished) but never past a later box that intersects box j.  Every
// dependency therefore has a smaller ticket than its waiter, is already running or finished, and the waits cannot
// deadlock (they are also bounded by a watchdog bit).
// Frame and scratch reads use ld.global.cg because other SMs rewrite those bytes during the kernel.
#include "common.cuh"

namespace {

#ifndef VD_COMP_MINB
#define VD_COMP_MINB 4            // resident CTAs per SM the run kernel is compiled for (64 registers; A/B on B200: 3 -> 0.676 ms, 4 -> 0.645, 5 -> 0.632 on the bench boxes, single-frame chains 0.046 / 0.055 / 0.058)
#endif
constexpr int kThreads = 256;
constexpr int MA_PIX = 256;               // mosaic phase A: small pixels per tile
constexpr int MB_W = 64, MB_H = 16;       // mosaic phase B tile
// Gaussian, dp4a path (every tap <= 255, ksize <= 125 — all of BASELINE's kernels).  The intermediate of the separable
// blur is stored as BYTE PLANES PACKED ALONG y: word (c, yg, x) of the HI (LO) plane holds the high (low) bytes of the
// u16 horizontal sums of rows 4yg..4yg+3 at column x, channel c — i.e. exactly the operand word the vertical pass
// feeds to dp4a, so that pass stages plain words instead of transposing bytes.  Rows are VIRTUAL rows v = y + r4
// (r4 = radius rounded up to 4): the horizontal pass also produces the REFLECT_101 halo rows above and below the box
// (source row reflect101(v - r4)), so the vertical pass never reflects and never leaves word alignment.
constexpr int GA_W = 128, GA_H = 32;      // phase A tile (horizontal pass): 128 pixels x 32 virtual rows
constexpr int GB_E = 128, GB_H = 60;      // phase B tile (vertical pass): 128 flat (pixel*3 + channel) elements x 60 rows
                                          // (60 rows + the 0..3 rows of word alignment = 16 packed words = 4 groups of 4)
// Gaussian, generic path (a tap > 255 or ksize > 125): u16 scratch [y][x*3+c], scalar taps
constexpr int GA_W0 = 128, GA_H0 = 8;
constexpr int GB_E0 = 64, GB_H0 = 64;
constexpr int FUSE_MAX_PX = 2048;         // fused Gaussian box: 3 B (u8) + 6 B (u16) per pixel in smem (18 KB)
constexpr int kMaxDynSmem = 64 * 1024;
constexpr int KW_MAX = 32;                // dp4a fast path: ceil((ksize + 3) / 4) <= KW_MAX  (ksize <= 125)

struct BoxMeta {
  int x1, y1, x2, y2;      // clipped; empty boxes have tA = tB = 0
  int tA, tB;              // fused box: tA = 1, tB = 0
  int sw, sh;              // mosaic: down-sampled size.  Gaussian dp4a path: sw = packed row groups (YG), sh = padded width
  unsigned int item_base;  // first done-flag of this box inside its frame: A tiles, then B tiles
  int fused;
  unsigned int posA, posB; // first TICKET-ORDER item of the A / B tiles inside the frame (plan kernel, lagged schedule)
  int has_dep;             // 0: no earlier box of the frame intersects this one (its tiles have no writer to wait for)
  int pad_;
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
  unsigned int* omap;  // [B][items_cap] ticket-order item -> (box << 1) | phase (0 = A, 1 = B)
  uint16_t* scratch;   // [B][scratch_elems]
  size_t scratch_elems;
  unsigned int items_cap;
};

struct Taps {
  uint16_t q[256];
  // dp4a fast path: taps shifted by a = 0..3 bytes and packed 4 per word: qa[a][j] byte b = q[4j + b - a], so that
  // out[4m + a] = sum_j dp4a(S[m + j], qa[a][j]) for a byte sequence S packed 4 per word
  uint32_t qa[4][KW_MAX];
  int kw;                  // words per shifted tap vector, 0 = fast path not applicable (a tap > 255 or ksize > 125)
};

__host__ __device__ inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ inline size_t items_cap_for(int maxb, size_t scratch_px) { return 4 * static_cast<size_t>(maxb) + scratch_px / 128 + 1024; }

__host__ size_t ws_bytes(int B, int maxb, size_t scratch_px) {
  return al256(sizeof(unsigned int) * (B + 4)) + al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb) +
         2 * al256(sizeof(int) * static_cast<size_t>(B) * items_cap_for(maxb, scratch_px)) +
         al256(static_cast<size_t>(B) * scratch_px * 3 * 2) + 256;
}

__host__ Ws carve(void* p0, int B, int maxb, size_t total_bytes, size_t* flag_bytes) {
  Ws w;
  uint8_t* p = static_cast<uint8_t*>(p0);
  w.hdr = reinterpret_cast<Header*>(p);
  p += al256(sizeof(unsigned int) * (B + 4));
  w.meta = reinterpret_cast<BoxMeta*>(p);
  p += al256(sizeof(BoxMeta) * static_cast<size_t>(B) * maxb);
  // split what is left: flags and the item map get 1 int each per 128 scratch pixels (+ 4 per box), the rest is scratch
  const size_t left = total_bytes - static_cast<size_t>(p - static_cast<uint8_t*>(p0)) - 256;
  const size_t fixed = 2 * sizeof(int) * static_cast<size_t>(B) * (4 * static_cast<size_t>(maxb) + 1024);
  // scratch_px * (6 + 8/128) * B + fixed = left
  const size_t scratch_px = static_cast<size_t>((left > fixed ? left - fixed : 0) / (static_cast<double>(B) * (6.0 + 8.0 / 128.0)));
  w.items_cap = static_cast<unsigned int>(4 * static_cast<size_t>(maxb) + scratch_px / 128 + 1024 - 8);
  w.done = reinterpret_cast<int*>(p);
  *flag_bytes = sizeof(int) * static_cast<size_t>(B) * w.items_cap;
  p += al256(*flag_bytes);
  w.omap = reinterpret_cast<unsigned int*>(p);
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
                      int ksize, int fast, unsigned int lag_items, uint32_t* __restrict__ status, Ws ws) {
  __shared__ unsigned int s_items[1024];
  __shared__ unsigned long long s_scr[1024];
  __shared__ unsigned int s_total;
  __shared__ int4 s_rect[1024];
  __shared__ unsigned short s_p[1024];
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
          } else if (fast) {
            const int r4 = ((ksize >> 1) + 3) & ~3;
            const int YG = (bh + 2 * r4 + 3) >> 2, bwp = (bw + 3) & ~3;
            m.sw = YG; m.sh = bwp;
            m.tA = ((bw + GA_W - 1) / GA_W) * ((4 * YG + GA_H - 1) / GA_H);
            m.tB = ((bw * 3 + GB_E - 1) / GB_E) * ((bh + GB_H - 1) / GB_H);
            scr = 12ull * YG * bwp;            // u16 units: 2 byte planes x 3 channels x YG x bwp words
          } else {
            m.tA = ((bw + GA_W0 - 1) / GA_W0) * ((bh + GA_H0 - 1) / GA_H0);
            m.tB = ((bw * 3 + GB_E0 - 1) / GB_E0) * ((bh + GB_H0 - 1) / GB_H0);
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
    if (overflow) return;
    __syncthreads();                                                   // this CTA's meta rows are visible to all its threads

    // ---- ticket order of the frame's items (posA / posB + the item map) ----
    // The run kernel hands items out in ticket order, and a B tile needs A tiles of its own box: scheduled right behind
    // them it would find them still running and the CTA would sit in the flag wait.  So the B tiles of box j are held back
    // behind the A tiles of the following boxes until at least `lag_items` of those separate them — but never past a later
    // box that overlaps box j, whose A tiles wait for exactly these B tiles (every wait must point at a smaller ticket:
    // that is what keeps the flag waits deadlock-free).  Group k = [B tiles placed before box k][A tiles of box k];
    // group nbx = the B tiles left over at the end.
    BoxMeta* fmw = ws.meta + static_cast<size_t>(f) * maxb;
    if (nbx <= 1024 && lag_items > 0) {
      unsigned int* s_tA = reinterpret_cast<unsigned int*>(s_scr);
      unsigned int* s_tB = s_tA + 1024;
      const int j = tid;
      unsigned int tA = 0, tB = 0;
      int4 rc = make_int4(0, 0, 0, 0);
      if (j < nbx) {
        const BoxMeta mj = fmw[j];
        tA = static_cast<unsigned int>(mj.tA); tB = static_cast<unsigned int>(mj.tB);
        rc = make_int4(mj.x1, mj.y1, mj.x2, mj.y2);
      }
      s_rect[j] = rc; s_tA[j] = tA; s_tB[j] = tB;
      __syncthreads();
      int p = nbx;                                     // group the B tiles of box j go to
      if (tB > 0) {
        unsigned int between = 0;
        int k = j + 1;
        for (; k < nbx; ++k) {
          if (s_tA[k] + s_tB[k] == 0) continue;
          const int4 o = s_rect[k];
          if (max(rc.x, o.x) < min(rc.z, o.z) && max(rc.y, o.y) < min(rc.w, o.w)) break;   // box k reads what these tiles write
          if (between >= lag_items) break;
          between += s_tA[k];
        }
        p = k;
      }
      s_p[j] = static_cast<unsigned short>(p);
      s_items[j] = tA;
      __syncthreads();
      if (tB > 0 && p < nbx) atomicAdd(&s_items[p], tB);
      __syncthreads();
      const unsigned int G = s_items[j];               // items of group j
      for (int off = 1; off < 1024; off <<= 1) {
        unsigned int a = 0;
        if (tid >= off) a = s_items[tid - off];
        __syncthreads();
        s_items[tid] += a;
        __syncthreads();
      }
      const unsigned int incl = s_items[j];
      __syncthreads();
      s_items[j] = incl - G;                           // first item of group j
      if (j == 1023) s_total = incl;
      __syncthreads();
      if (j < nbx) {
        int dep = 0;                                   // does an earlier box write into this one?
        for (int q = 0; q < j && !dep; ++q) {
          const int4 o = s_rect[q];
          dep = (s_tA[q] + s_tB[q]) != 0 && max(rc.x, o.x) < min(rc.z, o.z) && max(rc.y, o.y) < min(rc.w, o.w);
        }
        fmw[j].has_dep = dep;
        unsigned int posB = 0;
        if (tB > 0) {
          unsigned int before = 0;                     // B tiles of earlier boxes in the same group
          for (int q = 0; q < j; ++q)
            if (s_p[q] == p) before += s_tB[q];
          posB = (p < nbx ? s_items[p] : s_total) + before;      // group nbx starts where the groups 0 .. nbx-1 end
        }
        fmw[j].posA = s_items[j] + (G - tA);
        fmw[j].posB = posB;
      }
    } else {
      for (int k = tid; k < nbx; k += 1024) {
        fmw[k].posA = fmw[k].item_base;
        fmw[k].posB = fmw[k].item_base + static_cast<unsigned int>(fmw[k].tA);
        fmw[k].has_dep = 1;                            // not analysed: the run kernel scans the earlier boxes
      }
    }
    __syncthreads();
    unsigned int* omap = ws.omap + static_cast<size_t>(f) * ws.items_cap;
    for (int k = tid >> 5; k < nbx; k += 32) {         // one warp per box
      const BoxMeta mk = fmw[k];
      for (int i = tid & 31; i < mk.tA; i += 32) omap[mk.posA + i] = static_cast<unsigned int>(k) << 1;
      for (int i = tid & 31; i < mk.tB; i += 32) omap[mk.posB + i] = (static_cast<unsigned int>(k) << 1) | 1u;
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
                                             int ry2, int mode, int gb_e, int gb_h, uint32_t* status) {
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
    } else {                      // Gaussian B tiles span gb_e flat (pixel*3 + channel) elements
      th = gb_h;
      tiles_x = ((o.x2 - o.x1) * 3 + gb_e - 1) / gb_e;
      tx0 = ((ix1 - o.x1) * 3) / gb_e; tx1 = ((ix2 - o.x1) * 3 - 1) / gb_e;
    }
    const int ty0 = (iy1 - o.y1) / th, ty1 = (iy2 - 1 - o.y1) / th;
    const int* base = done + o.item_base + o.tA;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) spin_flag(base + ty * tiles_x + tx, status);
  }
}

// Both passes of a fused (whole box in shared memory) Gaussian item.  The taps are symmetric (q[t] == q[ksize-1-t],
// vd_gaussian_taps_q8), so only the first half lives in registers and every multiply serves two samples:
// KT / 2 >= ksize / 2 registers, the ones beyond ksize / 2 are zero.  out = (sum_j q[j] * (sum_i q[i] * src[refl]) + 32768) >> 16.
template <int KT>
__device__ __forceinline__ void fused_passes(const uint8_t* __restrict__ src, uint16_t* __restrict__ mid,
                                             const uint16_t* __restrict__ xtab, const uint16_t* __restrict__ ytab,
                                             const uint16_t* __restrict__ s_q, int ksize, int bw, int bh, uint8_t* img,
                                             int pitch, int x1, int y1) {
  const int tid = threadIdx.x;
  const int row3 = bw * 3, n = bh * row3;
  const int hk = ksize >> 1;
  uint32_t qh[KT / 2];
#pragma unroll
  for (int t = 0; t < KT / 2; ++t) qh[t] = t < hk ? s_q[t] : 0u;
  const uint32_t qc = s_q[hk];
  for (int i = tid; i < n; i += kThreads) {
    const int y = i / row3, e = i - y * row3;
    const int x = e / 3, c = e - x * 3;
    const uint8_t* srow = src + y * row3 + c;
    const uint16_t* tb = xtab + x;
    uint32_t acc = qc * srow[tb[hk]];
#pragma unroll
    for (int t = 0; t < KT / 2; ++t) {
      const int t2 = max(ksize - 1 - t, 0);              // mirror sample (a zero tap for t >= hk: any valid index will do)
      acc += qh[t] * (static_cast<uint32_t>(srow[tb[t]]) + srow[tb[t2]]);
    }
    mid[i] = static_cast<uint16_t>(min(acc, 65535u));
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    const int y = i / row3, e = i - y * row3;
    const uint16_t* tb = ytab + y;
    const uint16_t* col = mid + e;
    uint32_t acc = qc * col[tb[hk] * row3];
#pragma unroll
    for (int t = 0; t < KT / 2; ++t) {
      const int t2 = max(ksize - 1 - t, 0);
      acc += qh[t] * (static_cast<uint32_t>(col[tb[t] * row3]) + col[tb[t2] * row3]);
    }
    img[static_cast<size_t>(y1 + y) * pitch + x1 * 3 + e] = static_cast<uint8_t>((acc + 32768u) >> 16);
  }
}

__global__ void __launch_bounds__(kThreads, VD_COMP_MINB)
composite_run_kernel(const vd_frame_t* __restrict__ frames, int B, int maxb, int mode, int ksize, Taps taps,
                     uint32_t* __restrict__ status, Ws ws) {
  extern __shared__ __align__(16) uint8_t dsm[];
  __shared__ unsigned int s_tk;
  __shared__ unsigned int s_fprefix[130];
  __shared__ uint16_t s_q[256];
  __shared__ uint4 s_qt[KW_MAX];                  // s_qt[j] = (qa[0][j], qa[1][j], qa[2][j], qa[3][j]): one LDS.128 per tap word
  __shared__ int s_mx[MB_W], s_my[MB_H];          // mosaic phase B: nearest-neighbour source index of the tile's columns / rows
  const int tid = threadIdx.x;
  for (int i = tid; i < B; i += kThreads) s_fprefix[i] = ws.hdr->frame_prefix[i + 1];   // items of frame i
  s_q[tid] = taps.q[tid];
  if (tid < KW_MAX) s_qt[tid] = make_uint4(taps.qa[0][tid], taps.qa[1][tid], taps.qa[2][tid], taps.qa[3][tid]);
  __syncthreads();
  unsigned int max_items = 0;
  for (int i = 0; i < B; ++i) max_items = max(max_items, s_fprefix[i]);
  const unsigned int total = max_items * static_cast<unsigned int>(B);
  const int r = ksize >> 1;
  const int kw = taps.kw;
  const int r4 = (r + 3) & ~3;                               // dp4a path: halo rows above / below the box
  const int gb_e = kw > 0 ? GB_E : GB_E0, gb_h = kw > 0 ? GB_H : GB_H0;   // B tile geometry of the active Gaussian path

  // The top barrier doubles as the end-of-item barrier: every thread's stores -> bar.sync -> st.release.gpu of the item's
  // flag by one thread (release is cumulative over what the barrier ordered before it).  The ticket is taken only when the
  // CTA is free to start the item: a ticket fetched ahead would sit idle for one item's duration while its dependants wait
  // (measured: 0.38 -> 0.53 ms on the bench boxes).
  int* my_flag = nullptr;
  while (true) {
    __syncthreads();
    if (tid == 0) {
      if (my_flag) st_release(my_flag, 1);
      s_tk = atomicAdd(&ws.hdr->ticket, 1u);
    }
    my_flag = nullptr;
    __syncthreads();
    const unsigned int t = s_tk;
    if (t >= total) break;
    const int f = static_cast<int>(t % static_cast<unsigned int>(B));
    const unsigned int local = t / static_cast<unsigned int>(B);
    if (local >= s_fprefix[f]) continue;                           // this frame has fewer items
    const vd_frame_t fr = frames[f];                                // (issued early: in flight behind the map / box loads)
    const BoxMeta* fm = ws.meta + static_cast<size_t>(f) * maxb;
    int* done = ws.done + static_cast<size_t>(f) * ws.items_cap;
    // owner of the item: the plan kernel's map, ticket-order item -> (box, phase)
    const unsigned int ent = __ldg(ws.omap + static_cast<size_t>(f) * ws.items_cap + local);
    const int k = static_cast<int>(ent >> 1);
    const BoxMeta m = fm[k];
    const int has_dep = m.has_dep;
    const bool phaseA = (ent & 1u) == 0;
    const int tile = static_cast<int>(local - (phaseA ? m.posA : m.posB));
    my_flag = done + m.item_base + (phaseA ? tile : m.tA + tile);
    const int bw = m.x2 - m.x1, bh = m.y2 - m.y1;
    const int row3 = bw * 3;

    // ---- tile geometry (box-local) and dependencies ----
    int tx = 0, ty = 0;
    if (mode == VD_MODE_MOSAIC) {
      if (phaseA) {
        if (has_dep) wait_writers(fm, done, k, m.x1, m.y1, m.x2, m.y2, mode, gb_e, gb_h, status);
      } else {
        const int tiles_x = (bw + MB_W - 1) / MB_W;
        ty = tile / tiles_x; tx = tile - ty * tiles_x;
        for (int a = tid; a < m.tA; a += kThreads) spin_flag(done + m.item_base + a, status);
      }
    } else if (m.fused) {
      if (has_dep) wait_writers(fm, done, k, m.x1, m.y1, m.x2, m.y2, mode, gb_e, gb_h, status);
    } else if (kw > 0) {
      if (phaseA) {
        // reads frame pixels: columns x0 - r .. x0 + cols + r (reflected ones stay inside that range), source rows of
        // its virtual rows — the whole box height for a tile that contains halo (reflected) rows
        const int tiles_x = (bw + GA_W - 1) / GA_W;
        ty = tile / tiles_x; tx = tile - ty * tiles_x;
        const int x0 = tx * GA_W, v0 = ty * GA_H;
        const int cx1 = max(0, x0 - r), cx2 = min(bw, x0 + GA_W + r);
        int cy1 = v0 - r4, cy2 = v0 + GA_H - r4;
        if (cy1 < 0 || cy2 > bh) { cy1 = 0; cy2 = bh; }
        if (has_dep) wait_writers(fm, done, k, m.x1 + cx1, m.y1 + cy1, m.x1 + cx2, m.y1 + cy2, mode, gb_e, gb_h, status);
      } else {
        // a B tile (rows [y0, y0 + rows), pixels [px1, px2)) waits for the A tiles of its own box that
        //   (RAW) produce the packed rows it consumes: virtual rows [y0 + d, y0 + rows + d + 2r), same pixel columns;
        //   (WAR) still have to READ frame pixels it is about to overwrite: A tiles whose source rows (reflected ones
        //         included) meet [y0, y0 + rows) and whose columns +- r meet [px1, px2).
        const int tiles_x = (row3 + GB_E - 1) / GB_E;
        ty = tile / tiles_x; tx = tile - ty * tiles_x;
        const int y0 = ty * GB_H, rows = min(GB_H, bh - y0), d = r4 - r;
        const int px1 = (tx * GB_E) / 3, px2 = min(bw, (tx * GB_E + GB_E + 2) / 3);
        const int atx = (bw + GA_W - 1) / GA_W, aty = (4 * m.sw + GA_H - 1) / GA_H;
        const int ax0 = max(0, (px1 - r - (GA_W - 1) + GA_W * 4) / GA_W - 4), ax1 = min(atx - 1, (px2 - 1 + r) / GA_W);
        const int nax = ax1 - ax0 + 1;
        for (int i = tid; i < aty * nax; i += kThreads) {
          const int ay = i / nax, ax = ax0 + (i - ay * nax);
          const int v0 = ay * GA_H, v1 = v0 + GA_H;
          bool need = v0 < y0 + rows + d + 2 * r && v1 > y0 + d;                // RAW
          if (!need) {                                                          // WAR: conservative source-row interval
            int lo = v0 - r4, hi = v1 - r4;
            if (bh <= 2 * r4 + GA_H) { lo = 0; hi = bh; }                       // short boxes: reflections may land anywhere
            else if (lo < 0) { hi = max(hi, 1 - lo); lo = 0; }                  // rows -lo .. 1 mirrored at the top
            else if (hi > bh) { lo = min(lo, 2 * (bh - 1) - (hi - 1)); hi = bh; }
            need = lo < y0 + rows && hi > y0;
          }
          if (need) spin_flag(done + m.item_base + ay * atx + ax, status);
        }
      }
    } else if (phaseA) {
      const int tiles_x = (bw + GA_W0 - 1) / GA_W0;
      ty = tile / tiles_x; tx = tile - ty * tiles_x;
      const int cx1 = max(0, tx * GA_W0 - r), cx2 = min(bw, tx * GA_W0 + GA_W0 + r);
      const int cy1 = ty * GA_H0, cy2 = min(bh, ty * GA_H0 + GA_H0);
      if (has_dep) wait_writers(fm, done, k, m.x1 + cx1, m.y1 + cy1, m.x1 + cx2, m.y1 + cy2, mode, gb_e, gb_h, status);
    } else {
      const int tiles_x = (row3 + GB_E0 - 1) / GB_E0;
      ty = tile / tiles_x; tx = tile - ty * tiles_x;
      // A tiles of this box with rows in (tile rows +- r) and pixel columns in (tile columns +- r)
      const int px1 = (tx * GB_E0) / 3, px2 = min(bw, (tx * GB_E0 + GB_E0 + 2) / 3);
      const int cx1 = max(0, px1 - r), cx2 = min(bw, px2 + r);
      const int cy1 = max(0, ty * GB_H0 - r), cy2 = min(bh, ty * GB_H0 + GB_H0 + r);
      const int atx = (bw + GA_W0 - 1) / GA_W0;
      const int ax0 = cx1 / GA_W0, ax1 = (cx2 - 1) / GA_W0, ay0 = cy1 / GA_H0, ay1 = (cy2 - 1) / GA_H0;
      const int nax = ax1 - ax0 + 1, n = nax * (ay1 - ay0 + 1);
      for (int i = tid; i < n; i += kThreads)
        spin_flag(done + m.item_base + (ay0 + i / nax) * atx + ax0 + i % nax, status);
    }
    // flag acquire (per thread) -> bar.sync -> ld.global.cg by every thread of the CTA: the barrier extends the
    // acquiring thread's view to the whole CTA, no fence (and no L1 invalidation) needed
    __syncthreads();

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
        // cv2's INTER_NEAREST map (double arithmetic) once per tile column / row, not per pixel
        if (tid < MB_W) s_mx[tid] = nearest_src(min(tx * MB_W + tid, bw - 1), bw, m.sw);
        else if (tid < MB_W + MB_H) s_my[tid - MB_W] = nearest_src(min(ty * MB_H + (tid - MB_W), bh - 1), bh, m.sh);
        __syncthreads();
        for (int i = tid; i < MB_W * MB_H; i += kThreads) {
          const int ly = i / MB_W, lx = i - ly * MB_W;
          const int y = ty * MB_H + ly, x = tx * MB_W + lx;
          if (y < bh && x < bw) {
            const uint8_t* s = small + (s_my[ly] * m.sw + s_mx[lx]) * 3;
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
      if (ksize <= 8) fused_passes<8>(src, mid, xtab, ytab, s_q, ksize, bw, bh, img, fr.pitch, m.x1, m.y1);
      else if (ksize <= 16) fused_passes<16>(src, mid, xtab, ytab, s_q, ksize, bw, bh, img, fr.pitch, m.x1, m.y1);
      else if (ksize <= 32) fused_passes<32>(src, mid, xtab, ytab, s_q, ksize, bw, bh, img, fr.pitch, m.x1, m.y1);
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
    } else if (kw > 0 && phaseA) {
      // ---------------- horizontal pass (dp4a): GA_W pixels x GA_H virtual rows -> packed byte planes ----------------
      const int YG = m.sw, bwp = m.sh;
      const int x0 = tx * GA_W, v0 = ty * GA_H;
      const int cols = min(GA_W, bw - x0);
      const int vrows = min(GA_H, 4 * YG - v0);              // a multiple of 4
      const int spanw = GA_W / 4 + kw;                       // words per plane row: byte i = pixel x0 - r + i
      uint32_t* planes = reinterpret_cast<uint32_t*>(dsm);   // [3][GA_H][spanw]
      const int ngroups = (cols + 2 * r + 3) >> 2;           // 4-pixel groups per row that feed valid outputs
      // Stage 1: the raw bytes of every source row, 16 bytes per load.  (The first packed-plane version fetched every
      // 4-pixel group with four 4-byte ld.global.cg: 48 L2 sector requests per 384 useful bytes, the SMs spent more time
      // waiting for those than computing.)  Row ly of the tile reads frame bytes [b0, b1) of source row ys(ly); they land
      // in s_raw[ly] at the row's own 16-byte phase, so every vector load is aligned on both sides.
      const int px_lo = max(0, x0 - r), px_hi = min(bw, x0 + cols + r);     // box-local pixel range the tile reads
      const int raw_pitch = ((GA_W + 2 * r) * 3 + 15 + 15) & ~15;            // bytes per staged row
      uint8_t* s_raw = dsm + 3 * GA_H * spanw * 4;                           // [GA_H][raw_pitch]
      const uint8_t* fr_end = fr.ptr + static_cast<size_t>(fr.h - 1) * fr.pitch + static_cast<size_t>(fr.w) * 3;
      const int row_len = (px_hi - px_lo) * 3;
      const int chunks = (row_len + 15 + 15) >> 4;                           // 16-byte chunks per row incl. both ragged ends
      // one warp per row (rows warp, warp + 8, ...), lanes over the row's chunks: the per-row quantities (source row,
      // alignment phase) are computed once per row and nothing divides by a run-time value
      const int warp = tid >> 5, lane = tid & 31;
      for (int ch = lane; ch < chunks; ch += 32) {          // one trip for ksize <= 33 (chunks <= 32)
        uint4 v[GA_H / 8];
#pragma unroll
        for (int u = 0; u < GA_H / 8; ++u) {
          const int ly = warp + 8 * u;
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          if (ly < vrows) {
            int ys = v0 + ly - r4;
            if (ys < 0 || ys >= bh) ys = reflect101(ys, bh);
            const uint8_t* p0 = img + static_cast<size_t>(m.y1 + ys) * fr.pitch + (m.x1 + px_lo) * 3;
            const uint8_t* pc = p0 - (reinterpret_cast<uintptr_t>(p0) & 15) + 16 * ch;   // aligned chunk of the row segment
            if (pc >= fr.ptr && pc + 16 <= fr_end) {
              v[u] = __ldcg(reinterpret_cast<const uint4*>(pc));
            } else {                                                                  // chunk straddles the frame buffer's end
              uint8_t* vb = reinterpret_cast<uint8_t*>(&v[u]);
              for (int k2 = 0; k2 < 16; ++k2)
                if (pc + k2 >= fr.ptr && pc + k2 < fr_end) vb[k2] = __ldcg(pc + k2);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < GA_H / 8; ++u) {
          const int ly = warp + 8 * u;
          if (ly < vrows) *reinterpret_cast<uint4*>(s_raw + ly * raw_pitch + 16 * ch) = v[u];
        }
      }
      __syncthreads();
      // Stage 2: de-interleave into byte planes.  Interior groups: 12 contiguous bytes = four shared-memory words, a funnel
      // shift and six PRMT; groups that touch the box border: REFLECT_101 per pixel (reflected pixels stay inside [px_lo, px_hi)).
      for (int ly = warp; ly < vrows; ly += 8) {
        int ys = v0 + ly - r4;
        if (ys < 0 || ys >= bh) ys = reflect101(ys, bh);
        const uint8_t* p0 = img + static_cast<size_t>(m.y1 + ys) * fr.pitch + (m.x1 + px_lo) * 3;
        const int phase = static_cast<int>(reinterpret_cast<uintptr_t>(p0) & 15);          // s_raw[ly][phase] = pixel px_lo
        const uint8_t* rrow = s_raw + ly * raw_pitch + phase;
        uint32_t* prow0 = planes + ly * spanw;
        for (int g = lane; g < ngroups; g += 32) {
          const int px = x0 - r + 4 * g;                     // box-local pixel of the group's first byte
          uint32_t q0, q1, q2;
          if (px >= px_lo && px + 3 < px_hi) {
            const int off = (px - px_lo) * 3 + phase;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_raw + ly * raw_pitch + (off & ~3));
            const int sh8 = (off & 3) * 8;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
            const uint32_t a = __funnelshift_r(w0, w1, sh8), b = __funnelshift_r(w1, w2, sh8), c = __funnelshift_r(w2, w3, sh8);
            q0 = __byte_perm(__byte_perm(a, b, 0x0630), c, 0x5210);     // bytes 0, 3, 6, 9 of the 12-byte group
            q1 = __byte_perm(__byte_perm(a, b, 0x0741), c, 0x6210);     // bytes 1, 4, 7, 10
            q2 = __byte_perm(__byte_perm(a, b, 0x0052), c, 0x7410);     // bytes 2, 5, 8, 11
          } else {
            q0 = q1 = q2 = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              int xs = px + q;
              if (xs < 0 || xs >= bw) xs = reflect101(xs, bw);
              xs = min(max(xs, px_lo), px_hi - 1);           // (columns beyond cols + 2r feed no valid output)
              const uint8_t* pp = rrow + (xs - px_lo) * 3;
              q0 |= static_cast<uint32_t>(pp[0]) << (8 * q);
              q1 |= static_cast<uint32_t>(pp[1]) << (8 * q);
              q2 |= static_cast<uint32_t>(pp[2]) << (8 * q);
            }
          }
          prow0[g] = q0;
          prow0[GA_H * spanw + g] = q1;
          prow0[2 * GA_H * spanw + g] = q2;
        }
      }
      __syncthreads();
      // task = (channel, group of 4 virtual rows, word of 4 pixels): 16 sums h[rr][a] = sum_t q[t] * pixel(x0 + 4mm + a - r + t)
      uint32_t* HI = reinterpret_cast<uint32_t*>(scr16);
      uint32_t* LO = HI + static_cast<size_t>(3) * YG * bwp;
      for (int t = tid; t < 3 * (GA_H / 4) * (GA_W / 4); t += kThreads) {
        const int c = t >> 8, ygl = (t >> 5) & 7, mm = t & 31;
        if (4 * ygl >= vrows || 4 * mm >= cols) continue;
        uint32_t acc[4][4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int a = 0; a < 4; ++a) acc[rr][a] = 0u;
        const uint32_t* prow = planes + (c * GA_H + 4 * ygl) * spanw + mm;
        for (int j = 0; j < kw; ++j) {
          const uint4 q4 = s_qt[j];
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const uint32_t wv = prow[rr * spanw + j];
            acc[rr][0] = __dp4a(wv, q4.x, acc[rr][0]);
            acc[rr][1] = __dp4a(wv, q4.y, acc[rr][1]);
            acc[rr][2] = __dp4a(wv, q4.z, acc[rr][2]);
            acc[rr][3] = __dp4a(wv, q4.w, acc[rr][3]);
          }
        }
        // (taps sum to 256: a sum never exceeds 65280, cv2's saturation to 65535 is a no-op)
        uint4 hi, lo;
        uint32_t* hp = &hi.x;
        uint32_t* lp = &lo.x;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const uint32_t h01 = __byte_perm(acc[0][a], acc[1][a], 0x0051), h23 = __byte_perm(acc[2][a], acc[3][a], 0x0051);
          const uint32_t l01 = __byte_perm(acc[0][a], acc[1][a], 0x0040), l23 = __byte_perm(acc[2][a], acc[3][a], 0x0040);
          hp[a] = __byte_perm(h01, h23, 0x5410);
          lp[a] = __byte_perm(l01, l23, 0x5410);
        }
        const size_t wofs = (static_cast<size_t>(c) * YG + (v0 >> 2) + ygl) * bwp + x0 + 4 * mm;
        *reinterpret_cast<uint4*>(HI + wofs) = hi;
        *reinterpret_cast<uint4*>(LO + wofs) = lo;
      }
    } else if (kw > 0) {
      // ---------------- vertical pass (dp4a): GB_E flat elements x GB_H rows, operands = the packed planes ----------------
      const int YG = m.sw, bwp = m.sh;
      const int d = r4 - r;                                  // output row y sits at packed-sequence index y + d
      const int e0 = tx * GB_E, y0 = ty * GB_H;
      const int ne = min(GB_E, row3 - e0);
      const int rows = min(GB_H, bh - y0);
      const int W0 = y0 >> 2;                                // first staged word row (y0 is a multiple of 4)
      const int NW = GB_H / 4 + 1 + kw;                      // staged word rows
      uint32_t* Shi = reinterpret_cast<uint32_t*>(dsm);      // [NW][GB_E]
      uint32_t* Slo = Shi + NW * GB_E;
      const uint32_t* HI = reinterpret_cast<const uint32_t*>(scr16);
      const uint32_t* LO = HI + static_cast<size_t>(3) * YG * bwp;
      const int e = tid & (GB_E - 1);
      if (e < ne) {
        const int ea = e0 + e, x = ea / 3, c = ea - 3 * x;
        const size_t col = static_cast<size_t>(c) * YG * bwp + x;
        for (int wr0 = tid >> 7; wr0 < NW; wr0 += 12) {      // 6 + 6 loads in flight per thread
          uint32_t vh[6], vl[6];
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int wr = wr0 + 2 * u;
            if (wr < NW) {
              const size_t o = col + static_cast<size_t>(min(W0 + wr, YG - 1)) * bwp;
              vh[u] = __ldcg(HI + o);
              vl[u] = __ldcg(LO + o);
            }
          }
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int wr = wr0 + 2 * u;
            if (wr < NW) {
              Shi[wr * GB_E + e] = vh[u];
              Slo[wr * GB_E + e] = vl[u];
            }
          }
        }
      }
      __syncthreads();
      if (e < ne) {
        // task = (element, group of 4 packed output words = 16 rows): the four words' operand windows overlap, so every
        // staged word is loaded once per task and used four times (registers w0..w3 slide down the column); the HI plane's
        // sums are shifted into place before the LO plane's are added: 16 accumulators, sum = sum_t q[t] * (256 hi + lo)
        uint8_t* ocol = img + static_cast<size_t>(m.y1 + y0) * fr.pitch + m.x1 * 3 + e0 + e;
        for (int grp = tid >> 7; grp < (GB_H / 4 + 1) / 4; grp += 2) {
          uint32_t acc[4][4];
#pragma unroll
          for (int mm = 0; mm < 4; ++mm)
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[mm][a] = 0u;
#pragma unroll
          for (int plane = 0; plane < 2; ++plane) {
            const uint32_t* col = (plane ? Slo : Shi) + (4 * grp) * GB_E + e;
            uint32_t w0 = col[0], w1 = col[GB_E], w2 = col[2 * GB_E], w3;
            for (int j = 0; j < kw; ++j) {
              const uint4 q4 = s_qt[j];
              w3 = col[(j + 3) * GB_E];
              acc[0][0] = __dp4a(w0, q4.x, acc[0][0]); acc[0][1] = __dp4a(w0, q4.y, acc[0][1]);
              acc[0][2] = __dp4a(w0, q4.z, acc[0][2]); acc[0][3] = __dp4a(w0, q4.w, acc[0][3]);
              acc[1][0] = __dp4a(w1, q4.x, acc[1][0]); acc[1][1] = __dp4a(w1, q4.y, acc[1][1]);
              acc[1][2] = __dp4a(w1, q4.z, acc[1][2]); acc[1][3] = __dp4a(w1, q4.w, acc[1][3]);
              acc[2][0] = __dp4a(w2, q4.x, acc[2][0]); acc[2][1] = __dp4a(w2, q4.y, acc[2][1]);
              acc[2][2] = __dp4a(w2, q4.z, acc[2][2]); acc[2][3] = __dp4a(w2, q4.w, acc[2][3]);
              acc[3][0] = __dp4a(w3, q4.x, acc[3][0]); acc[3][1] = __dp4a(w3, q4.y, acc[3][1]);
              acc[3][2] = __dp4a(w3, q4.z, acc[3][2]); acc[3][3] = __dp4a(w3, q4.w, acc[3][3]);
              w0 = w1; w1 = w2; w2 = w3;
            }
            if (plane == 0) {
#pragma unroll
              for (int mm = 0; mm < 4; ++mm)
#pragma unroll
                for (int a = 0; a < 4; ++a) acc[mm][a] <<= 8;
            }
          }
#pragma unroll
          for (int mm = 0; mm < 4; ++mm)
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const int ly = 16 * grp + 4 * mm + a - d;
              if (ly >= 0 && ly < rows)
                ocol[static_cast<size_t>(ly) * fr.pitch] = static_cast<uint8_t>((acc[mm][a] + 32768u) >> 16);
            }
        }
      }
    } else if (phaseA) {
      // ---------------- horizontal pass: GA_H0 rows x GA_W0 pixels ----------------
      const int x0 = tx * GA_W0, y0 = ty * GA_H0;
      const int rows = min(GA_H0, bh - y0);
      const int cols = min(GA_W0, bw - x0);
      {
        // (generic path: scalar taps; the dp4a path above handles every kernel whose taps fit a byte)
        const int span = GA_W0 + 2 * r;
        uint8_t* st = dsm;                             // [GA_H0][span*3]
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
      // ---------------- vertical pass: GB_H0 rows x GB_E0 flat elements ----------------
      const int e0 = tx * GB_E0, y0 = ty * GB_H0;
      const int ne = min(GB_E0, row3 - e0);
      const int rows = min(GB_H0, bh - y0);
      const int span = rows + 2 * r;                       // staged rows
      {
        uint16_t* st = reinterpret_cast<uint16_t*>(dsm);   // [span][GB_E0]
        for (int i = tid; i < span * ne; i += kThreads) {
          const int ly = i / ne, e = i - ly * ne;
          const int ys = reflect101(y0 - r + ly, bh);
          st[ly * GB_E0 + e] = __ldcg(scr16 + static_cast<size_t>(ys) * row3 + e0 + e);
        }
        __syncthreads();
        for (int i = tid; i < rows * ne; i += kThreads) {
          const int ly = i / ne, e = i - ly * ne;
          const uint16_t* s = st + ly * GB_E0 + e;
          uint32_t acc = 0;
          for (int tt = 0; tt < ksize; ++tt) acc += static_cast<uint32_t>(s_q[tt]) * s[tt * GB_E0];
          img[static_cast<size_t>(m.y1 + y0 + ly) * fr.pitch + m.x1 * 3 + e0 + e] =
              static_cast<uint8_t>((acc + 32768u) >> 16);
        }
      }
    }
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
  const int r = ksize >> 1;
  size_t smem = 0;
  if (mode == VD_MODE_GAUSS) {
    size_t a = static_cast<size_t>(GA_H0) * (GA_W0 + 2 * r) * 3;
    size_t b = static_cast<size_t>(GB_H0 + 2 * r) * GB_E0 * 2;
    if (taps.kw > 0) {
      a = static_cast<size_t>(3) * GA_H * (GA_W / 4 + taps.kw) * 4 +         // byte planes of the A tile
          static_cast<size_t>(GA_H) * ((((GA_W + 2 * r) * 3 + 15 + 15) & ~15) + 16);   // + its raw source rows
      b = static_cast<size_t>(2) * (GB_H / 4 + 1 + taps.kw) * GB_E * 4;      // HI / LO word rows of the B tile
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
  // B tiles trail their box's A tiles by about one grid's worth of tickets (tickets interleave the B frames)
  static const int lag_env = [] { const char* e = getenv("VD_COMP_LAG"); return e ? atoi(e) : -1; }();
  const unsigned int lag_items = lag_env >= 0 ? static_cast<unsigned int>(lag_env)
                                              : static_cast<unsigned int>((sms * per_sm + B - 1) / B);
  composite_plan_kernel<<<B, 1024, 0, st>>>(frames, B, det_rows, det_count, max_det, ext_boxes, ext_count, max_ext,
                                            maxb, mode, mosaic_level, ksize, taps.kw > 0 ? 1 : 0, lag_items, status, ws);
  VD_LAUNCH_CHECK();
  composite_run_kernel<<<sms * per_sm, kThreads, smem, st>>>(frames, B, maxb, mode, ksize, taps, status, ws);
  VD_LAUNCH_CHECK();
  return VD_OK;
}


// ---------------------------------------------------------------------------------------------
// dirty tiles: after compositing, only the tiles a box touched differ from the frame the caller still holds
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int DT_W = 64, DT_H = 16;       // tile: 64 x 16 pixels = 3 KB, 192 contiguous bytes per row

// One CTA per tile: does any clipped box of the frame intersect it (same int() truncation and clipping as the
// compositing pass, reference combine_detect.py:243, 145-150)?  If so its rows are copied from the device frame into the
// destination frame — the caller's PINNED HOST frame, written over PCIe by the SMs (zero copy), so the host never touches
// the unchanged ~56 % of the pixels and no per-box copy is issued.
__global__ void __launch_bounds__(128)
dirty_tiles_kernel(const vd_frame_t* __restrict__ frames, const vd_frame_t* __restrict__ dst_frames,
                   const float* __restrict__ det_rows, const int* __restrict__ det_count, int max_det,
                   const int* __restrict__ ext_boxes, const int* __restrict__ ext_count, int max_ext,
                   unsigned long long* __restrict__ bytes_copied) {
  __shared__ int s_hit;
  const int f = blockIdx.z, tid = threadIdx.x;
  const vd_frame_t fr = frames[f];
  const int tx0 = blockIdx.x * DT_W, ty0 = blockIdx.y * DT_H;
  if (tx0 >= fr.w || ty0 >= fr.h) return;
  const int tx1 = min(tx0 + DT_W, fr.w), ty1 = min(ty0 + DT_H, fr.h);
  if (tid == 0) s_hit = 0;
  __syncthreads();
  const int nd = det_rows ? min(det_count[f], max_det) : 0;
  const int ne = ext_boxes ? min(ext_count[f], max_ext) : 0;
  for (int k = tid; k < nd + ne && !s_hit; k += 128) {
    int x1, y1, x2, y2;
    if (k < nd) {
      const float* r = det_rows + (static_cast<size_t>(f) * max_det + k) * 16;
      x1 = static_cast<int>(__fmul_rn(r[0], static_cast<float>(fr.w)));
      y1 = static_cast<int>(__fmul_rn(r[1], static_cast<float>(fr.h)));
      x2 = static_cast<int>(__fmul_rn(r[2], static_cast<float>(fr.w)));
      y2 = static_cast<int>(__fmul_rn(r[3], static_cast<float>(fr.h)));
    } else {
      const int* e = ext_boxes + (static_cast<size_t>(f) * max_ext + (k - nd)) * 4;
      x1 = e[0]; y1 = e[1]; x2 = e[2]; y2 = e[3];
    }
    x1 = max(0, x1); y1 = max(0, y1); x2 = min(fr.w, x2); y2 = min(fr.h, y2);
    if (x2 > x1 && y2 > y1 && x1 < tx1 && x2 > tx0 && y1 < ty1 && y2 > ty0) s_hit = 1;
  }
  __syncthreads();
  if (!s_hit) return;
  const vd_frame_t fd = dst_frames[f];
  const int row_bytes = (tx1 - tx0) * 3;
  for (int ly = 0; ly < ty1 - ty0; ++ly) {
    const uint8_t* src = fr.ptr + static_cast<size_t>(ty0 + ly) * fr.pitch + tx0 * 3;
    uint8_t* dst = const_cast<uint8_t*>(fd.ptr) + static_cast<size_t>(ty0 + ly) * fd.pitch + tx0 * 3;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (row_bytes & 15) == 0) {
      if (tid * 16 < row_bytes) *reinterpret_cast<uint4*>(dst + tid * 16) = __ldcg(reinterpret_cast<const uint4*>(src + tid * 16));
    } else {
      for (int i = tid; i < row_bytes; i += 128) dst[i] = __ldcg(src + i);
    }
  }
  if (tid == 0) atomicAdd(bytes_copied, static_cast<unsigned long long>(row_bytes) * (ty1 - ty0));
}

}  // namespace

extern "C" int vd_copy_dirty_tiles(const vd_frame_t* frames, const vd_frame_t* dst_frames, int B, int max_h, int max_w,
                                   const float* det_rows, const int32_t* det_count, int max_det,
                                   const int32_t* ext_boxes, const int32_t* ext_count, int max_ext,
                                   unsigned long long* bytes_copied, vd_stream_t stream) {
  VD_REQUIRE(frames && dst_frames && bytes_copied, "vd_copy_dirty_tiles: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && max_h > 0 && max_w > 0, "vd_copy_dirty_tiles: bad extent");
  VD_REQUIRE((det_rows == nullptr) == (det_count == nullptr), "vd_copy_dirty_tiles: det_rows/det_count mismatch");
  VD_REQUIRE((ext_boxes == nullptr) == (ext_count == nullptr), "vd_copy_dirty_tiles: ext_boxes/ext_count mismatch");
  VD_REQUIRE(vd_cdiv(max_h, DT_H) <= 65535, "vd_copy_dirty_tiles: frame too tall");
  dim3 grid(vd_cdiv(max_w, DT_W), vd_cdiv(max_h, DT_H), B);
  dirty_tiles_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(frames, dst_frames, det_rows, det_count, max_det,
                                                                            ext_boxes, ext_count, max_ext, bytes_copied);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
