// conv3x3_rows.cu — K2r: 3x3 stride-1 pad-1 convolution 64 -> 64 channels (+ folded BN bias, optional ReLU), the
// conv2 of the three layer1 bottlenecks and the SSH conv7x7_3 layers (reference detect_face/nets/resnet via
// retinaface.py:71-73 — torchvision Bottleneck.conv2 — and nets/layers.py:54-62).
//
// In the generic kernel (conv_tcgen05.cu, halo-patch mode) these layers are N = 64 GEMMs: every 128x64x16 MMA re-reads
// 4 KB of A for 2 KB of B from shared memory — 48 clk of operand fetch for 32 clk of math, tensor pipe 45-63 % busy.
// Here two output rows share their A reads, exactly like the two c1 rows of stem_pool.cu:
//
//   * M = 128 consecutive output pixels of ONE image row; a "group" = output rows y and y+1 of a 128-column strip,
//     accumulator columns 0..63 = row y, 64..127 = row y+1;
//   * input row t of a group (t = 0..3: image rows y-1 .. y+2) feeds filter row t of the upper output row and filter
//     row t-1 of the lower one, so for t = 1, 2 ONE MMA with B = [W_{t,s} ; W_{t-1,s}] (N = 128) reads the A window once
//     for both rows: per group 24 N = 128 + 24 N = 64 MMAs instead of 72 N = 64 (operand fetch 2 688 instead of 3 456 clk
//     for 2 304 clk of math);
//   * the A operand of tap s is the input row in shared memory (130 pixels x 128 B, SWIZZLE_128B) read through a
//     descriptor whose start is shifted by s pixels (the swizzle is a function of absolute address bits, the same
//     property conv_tcgen05.cu's halo-patch mode relies on);
//   * input rows travel through a ring, each fetched ONCE per run of rows by one TMA box (out-of-range rows / columns
//     are zero-filled = the padding); the nine filter taps (72 KB) stay resident, stored so that the row pairs are
//     contiguous: position 3s + (2 - r) holds W_{r,s}; position 9 is a second copy of W_{0,0} followed by 8 KB of zeros:
//     the group's first MMA uses [W_{0,0} ; 0] to clear the lower row's accumulator columns;
//   * every output row accumulates its taps in the generic kernel's order (r, s, 16-channel step), so the results are
//     bit-identical to it;
//   * work split: the (frame, strip, row pair) sequence is cut into gridDim.x equal contiguous ranges — a CTA pays the
//     2 extra input rows of a run at most a few times instead of once per fixed-size unit.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (TMEM -> +bias ->
// ReLU -> 16-bit -> global, 64 contiguous bytes per thread).
#include <cuda.h>
#include <new>

#include "common.cuh"
#include "stem_pool.h"

namespace {

#ifndef CR_TMA_STORE
#define CR_TMA_STORE 0               // epilogue variant: stage each output row in shared memory, one TMA store per row
#endif
#ifndef CR_STG256
#define CR_STG256 0                  // 32-byte stores (STG.E.ENL2.256): measured 0.136 instead of 0.140 ms per layer1 launch, off
#endif
#ifndef CR_QUAD_STORES
#define CR_QUAD_STORES 0              // measured alternative for the epilogue's stores, see the epilogue
#endif
constexpr int CR_COLS = 128;                      // output columns per strip (M)
constexpr int CR_IN_PX = CR_COLS + 2;             // input pixels per row incl. the halo
constexpr int CR_ROW_BYTES = 136 * 128;           // ring slot: 130 pixels x 128 B, padded to the 1 KB swizzle atom
constexpr int CR_RING = 7;
constexpr int CR_WBLK = 64 * 128;                 // one filter tap: [64 cout][64 cin] = 8 KB
constexpr int CR_W_BYTES = 11 * CR_WBLK;
constexpr int CR_A_BYTES = CR_RING * CR_ROW_BYTES;
constexpr int CR_STAGE_BYTES = CR_TMA_STORE ? 128 * 128 : 0;   // CR_TMA_STORE: one output row of the strip (128 pixels x 64 channels)
constexpr int CR_SMEM = CR_A_BYTES + CR_W_BYTES + CR_STAGE_BYTES + 1024 /*bias + barriers*/ + 1024 /*alignment*/;
constexpr int CR_THREADS = 320;
constexpr int CR_TMEM_COLS = 256;                 // two accumulators of 2 rows x 64 columns

struct RowsParams {
  int B, H, W, n_strips, n_pairs;                 // n_pairs = ceil(H / 2)
  long total;                                     // B * n_strips * n_pairs groups
  int dst_C, dst_coff, relu;
  const float* bias;
  uint8_t* y;
};

struct Run { int b, t, k0, k1; };

// the CTA's range [pos, end) of the group sequence is processed as runs of consecutive row pairs of one strip
__device__ __forceinline__ Run next_run(const RowsParams& p, long pos, long end) {
  Run r;
  const long per_frame = static_cast<long>(p.n_strips) * p.n_pairs;
  r.b = static_cast<int>(pos / per_frame);
  const int rem = static_cast<int>(pos - r.b * per_frame);
  r.t = rem / p.n_pairs;
  r.k0 = rem - r.t * p.n_pairs;
  const long left = end - pos;
  r.k1 = static_cast<int>(min(static_cast<long>(p.n_pairs), r.k0 + left));
  return r;
}

// one 32-byte store (STG.E.ENL2.256, sm_100): half as many store instructions — and line look-ups — as two 16-byte ones
__device__ __forceinline__ void stg256(uint8_t* ptr, const uint4& a, const uint4& b) {
#if CR_STG256
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
#else
  *reinterpret_cast<uint4*>(ptr) = a;
  *reinterpret_cast<uint4*>(ptr + 16) = b;
#endif
}

__global__ void __launch_bounds__(CR_THREADS, 1)
conv3x3_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const RowsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + CR_A_BYTES;
  uint8_t* sStage = sB + CR_W_BYTES;                // 1 KB aligned: CR_A_BYTES and CR_W_BYTES are multiples of 1 KB
  float* sBias = reinterpret_cast<float*>(sStage + CR_STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + 64);
  uint64_t* empty = full + CR_RING;
  uint64_t* tfull = empty + CR_RING;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long pos0 = p.total * blockIdx.x / gridDim.x;
  const long pos1 = p.total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x < 64) sBias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  // the zero half of the [W_00 ; 0] block
  for (int i = threadIdx.x; i < CR_WBLK / 16; i += CR_THREADS)
    sts128(smem_u32(sB) + 10 * CR_WBLK + i * 16, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < CR_RING; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 8);
      }
      mbar_init(wfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, CR_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, 10 * CR_WBLK);
      for (int s = 0; s < 3; ++s)
        for (int r = 0; r < 3; ++r) tma_load_2d(sB + (3 * s + (2 - r)) * CR_WBLK, &tmB, wfull, (r * 3 + s) * 64, 0);
      tma_load_2d(sB + 9 * CR_WBLK, &tmB, wfull, 0, 0);
    }
    __syncwarp();
    uint32_t g = 0;                                   // running input-row counter (ring position)
    for (long pos = pos0; pos < pos1;) {
      const Run run = next_run(p, pos, pos1);
      const int nrows = 2 * (run.k1 - run.k0) + 2;
      int iy = 2 * run.k0 - 1;
      for (int j = 0; j < nrows; ++j, ++g, ++iy) {
        const uint32_t slot = g % CR_RING;
        mbar_wait(&empty[slot], ((g / CR_RING) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[slot], CR_IN_PX * 128);
          tma_load_4d(sA + slot * CR_ROW_BYTES, &tmA, &full[slot], 0, CR_COLS * run.t - 1, iy, run.b);
        }
        __syncwarp();
      }
      pos += run.k1 - run.k0;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t kFmt = VD_ACT_IS_FP16 ? 0u : 1u;
    constexpr uint32_t idesc64 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(64 >> 3) << 17) |
                                 (static_cast<uint32_t>(128 >> 4) << 24);
    constexpr uint32_t idesc128 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(128 >> 3) << 17) |
                                  (static_cast<uint32_t>(128 >> 4) << 24);
    constexpr uint32_t hi_std = (1024u >> 4) | (1u << 14) | (2u << 29);         // SBO 8 x 128 B, SWIZZLE_128B
    constexpr uint32_t WB = CR_WBLK >> 4;
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    mbar_wait(wfull, 0);
    uint32_t g = 0;                                   // ring counter of the run's first input row
    uint32_t grp = 0;
    for (long pos = pos0; pos < pos1;) {
      const Run run = next_run(p, pos, pos1);
      const int ngroups = run.k1 - run.k0;
      for (int k = 0; k < ngroups; ++k, ++grp) {
        const uint32_t r0 = g + static_cast<uint32_t>(2 * k);          // ring counter of the group's input row t = 0
        for (int t = (k == 0 ? 0 : 2); t < 4; ++t) mbar_wait(&full[(r0 + t) % CR_RING], ((r0 + t) / CR_RING) & 1u);
        const uint32_t acc = grp & 1u;
        mbar_wait(&tempty[acc], ((grp >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d_lo = tmem_base + acc * 128u;                // row y: columns 0..63, row y+1: 64..127
          const uint32_t d_hi = d_lo + 64u;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t a_row = a_lo0 + static_cast<uint32_t>(((r0 + t) % CR_RING) * (CR_ROW_BYTES >> 4));
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const uint32_t al = a_row + static_cast<uint32_t>(s * 8);          // shifted by s pixels of 128 B
              if (t == 0) {
                // upper row alone: W_{0,s}; the very first MMA also clears the lower row's columns ([W_00 ; 0])
                const uint32_t bl = b_lo0 + (s == 0 ? 9u * WB : static_cast<uint32_t>(3 * s + 2) * WB);
                if (s == 0) umma_bf16_words(d_lo, al, hi_std, bl, hi_std, idesc128, 0u);
                else        umma_bf16_words(d_lo, al, hi_std, bl, hi_std, idesc64, 1u);
                umma_bf16_words(d_lo, al + 2, hi_std, bl + 2, hi_std, idesc64, 1u);
                umma_bf16_words(d_lo, al + 4, hi_std, bl + 4, hi_std, idesc64, 1u);
                umma_bf16_words(d_lo, al + 6, hi_std, bl + 6, hi_std, idesc64, 1u);
              } else if (t < 3) {
                // [W_{t,s} ; W_{t-1,s}] starts at position 3s + (2 - t)
                const uint32_t bl = b_lo0 + static_cast<uint32_t>(3 * s + 2 - t) * WB;
                umma_bf16_words(d_lo, al, hi_std, bl, hi_std, idesc128, 1u);
                umma_bf16_words(d_lo, al + 2, hi_std, bl + 2, hi_std, idesc128, 1u);
                umma_bf16_words(d_lo, al + 4, hi_std, bl + 4, hi_std, idesc128, 1u);
                umma_bf16_words(d_lo, al + 6, hi_std, bl + 6, hi_std, idesc128, 1u);
              } else {
                const uint32_t bl = b_lo0 + static_cast<uint32_t>(3 * s) * WB;   // lower row alone: W_{2,s}
                umma_bf16_words(d_hi, al, hi_std, bl, hi_std, idesc64, 1u);
                umma_bf16_words(d_hi, al + 2, hi_std, bl + 2, hi_std, idesc64, 1u);
                umma_bf16_words(d_hi, al + 4, hi_std, bl + 4, hi_std, idesc64, 1u);
                umma_bf16_words(d_hi, al + 6, hi_std, bl + 6, hi_std, idesc64, 1u);
              }
            }
          }
          umma_commit(&empty[r0 % CR_RING]);                           // input rows no later group needs
          umma_commit(&empty[(r0 + 1) % CR_RING]);
          if (k == ngroups - 1) {
            umma_commit(&empty[(r0 + 2) % CR_RING]);
            umma_commit(&empty[(r0 + 3) % CR_RING]);
          }
          umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
      g += static_cast<uint32_t>(2 * ngroups + 2);
      pos += ngroups;
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                           // TMEM lane quarter
    const int half = (warp - 2) >> 2;                 // 32-channel half
    const int col = q * 32 + lane;                    // output column inside the strip
    uint32_t grp = 0;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = sBias[half * 32 + j];
    const size_t pix_bytes = static_cast<size_t>(p.dst_C) * 2;
    for (long pos = pos0; pos < pos1;) {
      const Run run = next_run(p, pos, pos1);
      const int ngroups = run.k1 - run.k0;
      const int X = CR_COLS * run.t + col;
      const bool xvalid = X < p.W;
      for (int k = 0; k < ngroups; ++k, ++grp) {
        const uint32_t acc = grp & 1u;
        mbar_wait(&tfull[acc], (grp >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128u + static_cast<uint32_t>(half * 32);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int Y = 2 * (run.k0 + k) + h;
          uint32_t v[32];
          tmem_ld_32x32(tcol + static_cast<uint32_t>(h * 64), v);
          tmem_ld_wait();
          if (h == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          uint4 o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
#pragma unroll
            for (int e = 0; e < 8; e += 2) fadd2(f[e], f[e + 1], bias_r[8 * j + e], bias_r[8 * j + e + 1]);
            if (p.relu) {
              o[j].x = pack_bf16x2_relu(f[0], f[1]);
              o[j].y = pack_bf16x2_relu(f[2], f[3]);
              o[j].z = pack_bf16x2_relu(f[4], f[5]);
              o[j].w = pack_bf16x2_relu(f[6], f[7]);
            } else {
              o[j].x = pack_bf16x2(f[0], f[1]);
              o[j].y = pack_bf16x2(f[2], f[3]);
              o[j].z = pack_bf16x2(f[4], f[5]);
              o[j].w = pack_bf16x2(f[6], f[7]);
            }
          }
          // What the epilogue's stores cost was measured with tools/probes/umma_mix.cu (profiles/r02_umma_mix_probe.log): this
          // kernel's MMA stream alone runs at 2 770 clk per group (ideal 2 688), with the ring's bulk loads, the tcgen05.ld
          // readers and the accumulator hand-shake at 2 990 — and at 3 980 once the readers also store 64 B per thread with
          // the lanes 128 B apart, as below (a warp instruction touches 32 lines with 16 bytes each); lane-contiguous
          // stores of the same bytes: 3 005.  In the kernel itself three remedies all land within 4 % of the plain stores
          // (0.140 ms per layer1 launch): CR_QUAD_STORES=1 (12 shuffles per row and thread, 8 half-lines per instruction)
          // 0.134, CR_TMA_STORE=1 (staging tile + one TMA store per row) 0.134, CR_STG256=1 (32-byte stores) 0.136 — the
          // launch is paced by the SM clock the 1 kW power cap allows under tensor load, not by these cycles.  All three
          // are kept as compile-time options; the default is the build the committed ncu captures were taken on.
#if CR_TMA_STORE
          // the row goes through a 16 KB staging tile (SWIZZLE_128B: chunk c of pixel i at i*128 + ((c ^ (i & 7)) << 4)) and
          // leaves with ONE TMA store, which also clips columns >= W and rows >= H
          {
            if (threadIdx.x == 64) tma_store_wait_read<0>();          // the previous row's store has read the tile
            named_bar_sync(1, 256);
            const uint32_t srow = smem_u32(sStage) + col * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) sts128(srow + (((half * 4 + j) ^ (col & 7)) << 4), o[j]);
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (threadIdx.x == 64) {
              tma_store_4d(&tmY, sStage, p.dst_coff, CR_COLS * run.t, Y, run.b);
              tma_store_commit();
            }
          }
#elif CR_QUAD_STORES
          // 4 x 4 transpose of the 16-byte chunks inside each lane quad: lane a of a quad ends up with chunk a of the
          // quad's four pixels, so that the quad writes the 64 contiguous bytes of ONE pixel per instruction (a warp
          // instruction touches 8 half-lines instead of 32 lines with 16 bytes each)
          {
            const int a = lane & 3;
            const uint4 own = a == 0 ? o[0] : a == 1 ? o[1] : a == 2 ? o[2] : o[3];
            uint4 tq[4] = {own, own, own, own};                       // tq[a] = this pixel's chunk a; the others arrive below
#pragma unroll
            for (int r = 1; r < 4; ++r) {
              const int si = (a + 4 - r) & 3;                         // chunk this lane sends: the receiver's quad index
              const uint4 snd = si == 0 ? o[0] : si == 1 ? o[1] : si == 2 ? o[2] : o[3];
              const int src = (lane & ~3) | ((a + r) & 3);
              uint4 rcv;
              rcv.x = __shfl_sync(0xffffffffu, snd.x, src);
              rcv.y = __shfl_sync(0xffffffffu, snd.y, src);
              rcv.z = __shfl_sync(0xffffffffu, snd.z, src);
              rcv.w = __shfl_sync(0xffffffffu, snd.w, src);
              const int di = (a + r) & 3;                             // pixel (inside the quad) the chunk belongs to
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j == di) tq[j] = rcv;
            }
            if (Y < p.H) {
              const int X0 = X - a;
              uint8_t* base = p.y + ((static_cast<size_t>(run.b) * p.H + Y) * p.W + X0) * pix_bytes +
                              static_cast<size_t>(p.dst_coff + half * 32) * 2 + a * 16;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (X0 + j < p.W) *reinterpret_cast<uint4*>(base + j * pix_bytes) = tq[j];
            }
          }
#else
          if (xvalid && Y < p.H) {
            uint8_t* dst = p.y + ((static_cast<size_t>(run.b) * p.H + Y) * p.W + X) * pix_bytes +
                           static_cast<size_t>(p.dst_coff + half * 32) * 2;
            stg256(dst, o[0], o[1]);
            stg256(dst + 32, o[2], o[3]);
          }
#endif
        }
      }
      pos += ngroups;
    }
#if CR_TMA_STORE
    if (threadIdx.x == 64) tma_store_wait_all<0>();
#endif
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, CR_TMEM_COLS);
  }
}

}  // namespace

struct vd_conv_rows_plan {
  CUtensorMap tmA, tmB, tmY;
  RowsParams kp;
  int grid;
};

bool vd_conv_rows_eligible(const vd_conv_desc_t* d) {
  const char* e = getenv("VD_CONV_ROWS");     // read per plan: tests build both variants in one process
  if (e && atoi(e) == 0) return false;
  const unsigned allowed = VD_CONV_RELU | VD_CONV_REVERSE;
  if (d->flags & ~allowed) return false;
  if (d->R != 3 || d->S != 3 || d->stride != 1 || d->pad != 1 || d->Cin != 64 || d->Cout != 64) return false;
  if (d->Hin != d->Hout || d->Win != d->Wout || d->dst_C % 16 || d->dst_coff % 16) return false;
  if (reinterpret_cast<uintptr_t>(d->y) & 31) return false;          // keeps the CR_STG256 build valid for every plan
  // the last strip of a row computes columns beyond the image: accept up to 15 % of wasted columns
  const int strips = vd_cdiv(d->Wout, CR_COLS);
  return static_cast<long>(strips) * CR_COLS * 100 <= static_cast<long>(d->Wout) * 115;
}

int vd_conv_rows_create(const vd_conv_desc_t* d, vd_conv_rows_plan** out) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(vd_tensor_map_encoder());
  if (!enc) {
    vd_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VD_ERR_NO_DEVICE;
  }
  vd_conv_rows_plan* pl = new (std::nothrow) vd_conv_rows_plan();
  VD_REQUIRE(pl, "out of host memory");
  RowsParams& k = pl->kp;
  k.B = d->B; k.H = d->Hout; k.W = d->Wout;
  k.n_strips = vd_cdiv(k.W, CR_COLS);
  k.n_pairs = vd_cdiv(k.H, 2);
  k.total = static_cast<long>(k.B) * k.n_strips * k.n_pairs;
  k.dst_C = d->dst_C; k.dst_coff = d->dst_coff;
  k.relu = (d->flags & VD_CONV_RELU) ? 1 : 0;
  k.bias = d->bias;
  k.y = static_cast<uint8_t*>(d->y);
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  pl->grid = static_cast<int>(k.total < sms ? k.total : sms);

  const CUtensorMapDataType dt = VD_ACT_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t gdim[4] = {64, static_cast<cuuint64_t>(d->Win), static_cast<cuuint64_t>(d->Hin), static_cast<cuuint64_t>(d->B)};
  cuuint64_t gstr[3] = {128, static_cast<cuuint64_t>(128) * d->Win, static_cast<cuuint64_t>(128) * d->Win * d->Hin};
  cuuint32_t box[4] = {64, CR_IN_PX, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&pl->tmA, dt, 4, const_cast<void*>(d->x), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(rows A) failed: CUresult %d", static_cast<int>(r));
    delete pl;
    return VD_ERR_CUDA;
  }
  cuuint64_t wdim[2] = {576, 64};
  cuuint64_t wstr[1] = {1152};
  cuuint32_t wbox[2] = {64, 64};
  cuuint32_t wes[2] = {1, 1};
  r = enc(&pl->tmB, dt, 2, const_cast<void*>(d->w), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(rows B) failed: CUresult %d", static_cast<int>(r));
    delete pl;
    return VD_ERR_CUDA;
  }
  cuuint64_t ydim[4] = {static_cast<cuuint64_t>(d->dst_C), static_cast<cuuint64_t>(d->Wout), static_cast<cuuint64_t>(d->Hout),
                        static_cast<cuuint64_t>(d->B)};
  cuuint64_t ystr[3] = {static_cast<cuuint64_t>(d->dst_C) * 2, static_cast<cuuint64_t>(d->dst_C) * 2 * d->Wout,
                        static_cast<cuuint64_t>(d->dst_C) * 2 * d->Wout * d->Hout};
  cuuint32_t ybox[4] = {64, CR_COLS, 1, 1};
  r = enc(&pl->tmY, dt, 4, d->y, ydim, ystr, ybox, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(rows Y) failed: CUresult %d", static_cast<int>(r));
    delete pl;
    return VD_ERR_CUDA;
  }
  *out = pl;
  return VD_OK;
}

void vd_conv_rows_destroy(vd_conv_rows_plan* pl) { delete pl; }

void vd_conv_rows_info(const vd_conv_rows_plan* pl, int* groups, int* grid) {
  if (groups) *groups = static_cast<int>(pl->kp.total);
  if (grid) *grid = pl->grid;
}

int vd_conv_rows_run(const vd_conv_rows_plan* pl, cudaStream_t st) {
  VD_SET_FUNC_ATTR_ONCE_PER_DEVICE(conv3x3_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM);
  conv3x3_rows_kernel<<<pl->grid, CR_THREADS, CR_SMEM, st>>>(pl->tmA, pl->tmB, pl->tmY, pl->kp);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
