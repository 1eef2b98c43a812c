// stem_pool.cu — K2s: conv1 7x7/2 + folded BN + ReLU + max-pool 3x3/2 in ONE kernel.
//
// Replaces the first three modules of the torchvision ResNet-50 body the reference runs
// (detect_face/retinaface.py:71-73,121: conv1 -> bn1 -> relu -> maxpool).  Unfused, the 64-channel
// half-resolution tensor c1 (66 MB per 1080p frame) is written by the convolution and read back by the pool:
// it is 4x the size of either neighbour and both kernels sit on the SM<->L2 fabric limit because of it.  Here
// c1 never leaves the SM:
//
//   * a CTA owns a "unit" = (frame, strip of 120 c1 columns, run of `seg_rows` pooled rows) and walks DOWN the
//     strip one c1 row at a time; M = 128 c1 pixels per row (the strip + 8 columns of left halo, of which the
//     pool needs one — 8 keeps the first input pixel on a 128-byte chunk);
//   * the A operand is the RAW padded input row in shared memory, read as a sliding window through a
//     no-swizzle K-major descriptor (LBO = 16 B, SBO = 128 B: row i, chunk c at 16*i + 16*c — output pixel i's
//     window starts two input pixels after pixel i-1's), exactly as in conv_tcgen05.cu's stem path;
//   * consecutive c1 rows share 5 of their 7 input rows: input rows travel through a ring of row PAIRS, each
//     fetched once per unit by one TMA box, so the fabric carries ~4.3 KB per c1 row instead of 14 KB;
//   * TWO c1 rows per accumulator (a "group" = c1 rows 2k-1 and 2k of the unit).  What bounds this kernel is the
//     shared-memory pipe (ncu: UTCHMMA issue stalled on MIO half of the time, the pool's LDS waiting behind it), and
//     an N = 64 MMA re-reads 4 KB of A for 2 KB of B.  Input row t of a group (t = 0..8) feeds filter row t of the
//     odd c1 row and filter row t-2 of the even one, so for t = 2..6 ONE MMA with B = [W_t ; W_{t-2}] (N = 128,
//     accumulator columns 0..63 = odd row, 64..127 = even row) reads the A window once for both: 11 N = 128 + 7
//     N = 64 MMAs per two rows instead of 28 N = 64 (operand fetch 520 instead of 672 clk per row).  Every row still
//     accumulates its filter rows in the order 0..6, so results stay bit-identical; the even row's columns are
//     zeroed by the group's first MMA, whose B is [W_0 ; 0];
//   * epilogue: TMEM -> +bias -> ReLU -> bf16.  An epilogue thread owns the SAME c1 column in every row, so the vertical
//     half of the 3x3/2 max stays in registers (the packed values of the previous two rows); only after every second
//     row the column-wise maximum of the last three rows goes to shared memory (two alternating 16 KB buffers), and
//     the horizontal max over three columns is taken from there and stored (32 B per thread, 1 KB contiguous per
//     warp): a third of the shared-memory reads and half of the writes of keeping every c1 row there (the first
//     version; shared-memory bandwidth is what bounds this kernel).  Positions outside c1 count as 0, which is
//     neutral for a max over ReLU outputs (PyTorch pads with -inf: every window holds at least one real value >= 0).
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue + pool.
#include <cuda.h>
#include <new>

#include "common.cuh"
#include "stem_pool.h"

namespace {

constexpr int SP_STRIP = 120;                    // c1 columns a strip owns
constexpr int SP_PCOLS = SP_STRIP / 2;           // pooled columns per strip
constexpr int SP_HALO = 8;                       // c1 halo columns on the left (M = 128)
constexpr int SP_CHUNKS = 17;                    // 2*128 + 6 = 262 input pixels * 8 B = 2096 B -> 17 x 128 B
constexpr int SP_ROW_BYTES = SP_CHUNKS * 128;    // 2176
constexpr int SP_PAIR_BYTES = 2 * SP_ROW_BYTES;  // 4352
constexpr int SP_RING = 8;                       // input row pairs in flight
constexpr int SP_WBLK = 64 * 64;                 // one filter row: [64 cout][8 taps x 4 ch] = 4 KB
// resident weights as six 8 KB row pairs: block 0 = [W_0 ; 0], block t-1 = [W_t ; W_{t-2}] for t = 2..6
constexpr int SP_W_BYTES = 12 * SP_WBLK;
constexpr int SP_C1_BYTES = 128 * 128;           // one c1 row of the strip: 128 pixels x 64 channels
constexpr int SP_C1_RING = 2;                    // column-max rows (one per pooled row), alternating
constexpr int SP_A_BYTES = SP_RING * SP_PAIR_BYTES;                    // 34816
constexpr int SP_SMEM = SP_A_BYTES + SP_W_BYTES + SP_C1_RING * SP_C1_BYTES + 1024 /*bias + barriers*/ + 1024;
constexpr int SP_THREADS = 320;
constexpr int SP_TMEM_COLS = 256;                // two accumulators of 2 c1 rows x 64 columns

struct StemPoolParams {
  int B, Hin, Hc, Wc, Hp, Wpo;
  int n_strips, n_segs, seg_rows, n_units;
  const float* bias;
  uint4* y;                                      // [B][Hp][Wpo][64]
};

// column-max row layout: 16-byte chunk c of strip pixel r lives at r*128 + ((c ^ c1_swz(r)) << 4), c1_swz = (r & 7) with
// bits 0 and 1 swapped.  Writes (8 consecutive pixels, one chunk) hit 8 distinct 16-byte slots; pool reads (pixels
// r and r+2, four channel groups x one chunk parity) do too, because r -> r+2 always flips bit 0 of c1_swz.
__device__ __forceinline__ int c1_swz(int r) { return ((r >> 1) & 1) | ((r & 1) << 1) | (r & 4); }

__global__ void __launch_bounds__(SP_THREADS, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const StemPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + SP_A_BYTES;
  uint8_t* sC = sB + SP_W_BYTES;
  float* sBias = reinterpret_cast<float*>(sC + SP_C1_RING * SP_C1_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sBias + 64);
  uint64_t* empty = full + SP_RING;
  uint64_t* tfull = empty + SP_RING;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x < 64) sBias[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  if (threadIdx.x < SP_WBLK / 16) sts128(smem_u32(sB) + SP_WBLK + threadIdx.x * 16, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();                           // the zero half of block 0 is read by the tensor core
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < SP_RING; ++i) {
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
    tmem_alloc(tmem_slot, SP_TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(wfull, 11 * SP_WBLK);
      tma_load_2d(sB, &tmB, wfull, 0, 0);                                        // block 0 = [W_0 ; 0]
      for (int t = 2; t < 7; ++t) {                                              // block t-1 = [W_t ; W_{t-2}]
        tma_load_2d(sB + (t - 1) * 2 * SP_WBLK, &tmB, wfull, t * 32, 0);
        tma_load_2d(sB + (t - 1) * 2 * SP_WBLK + SP_WBLK, &tmB, wfull, (t - 2) * 32, 0);
      }
    }
    __syncwarp();
    uint32_t g = 0;                                   // running pair counter (ring position)
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u % p.n_segs;
      const int t = (u / p.n_segs) % p.n_strips;
      const int b = u / (p.n_segs * p.n_strips);
      const int py0 = seg * p.seg_rows;
      const int py1 = min(py0 + p.seg_rows, p.Hp);
      const int npairs = 2 * (py1 - py0) + 1 + 3;     // c1 rows of the unit + 3
      const int chunk0 = (SP_STRIP * t - SP_HALO) >> 3;                       // (2 * first c1 column) / 16
      int iy = 4 * py0 - 6;                           // first input row of pair 0
      for (int k = 0; k < npairs; ++k, ++g, iy += 2) {
        const uint32_t slot = g & (SP_RING - 1);
        mbar_wait(&empty[slot], ((g >> 3) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[slot], SP_PAIR_BYTES);
          tma_load_4d(sA + slot * SP_PAIR_BYTES, &tmA, &full[slot], 0, chunk0, iy, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t kFmt = VD_ACT_IS_FP16 ? 0u : 1u;
    constexpr uint32_t idesc64 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(64 >> 3) << 17) |
                                 (static_cast<uint32_t>(128 >> 4) << 24);
    constexpr uint32_t idesc128 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(128 >> 3) << 17) |
                                  (static_cast<uint32_t>(128 >> 4) << 24);
    constexpr uint32_t hi_raw = (128u >> 4) | (1u << 14);                       // SBO 128 B, no swizzle
    constexpr uint32_t hi_w = (512u >> 4) | (1u << 14) | (4u << 29);            // SBO 8 x 64 B, SWIZZLE_64B
    constexpr uint32_t WB = SP_WBLK >> 4;                                       // one filter row in descriptor units
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);       // LBO 16 B
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    mbar_wait(wfull, 0);
    uint32_t g = 0;                                   // ring position of the unit's first input row pair
    uint32_t grp = 0;                                 // running group counter (accumulator buffer, phases)
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u % p.n_segs;
      const int py0 = seg * p.seg_rows;
      const int py1 = min(py0 + p.seg_rows, p.Hp);
      const int ngroups = (py1 - py0) + 1;            // group k = c1 rows 2k-1 and 2k; group 0 has the even row only
      for (int k = 0; k < 4; ++k) mbar_wait(&full[(g + k) & (SP_RING - 1)], ((g + k) >> 3) & 1u);
      for (int k = 0; k < ngroups; ++k, ++grp) {
        // input row t of the group (t = 0..8) lives in pair P0 + ((t + 1) >> 1), sub-row (t + 1) & 1; P0 = 2k - 1
        const uint32_t P0 = g + static_cast<uint32_t>(2 * k) - 1u;
        if (k > 0) {
          mbar_wait(&full[(P0 + 3) & (SP_RING - 1)], ((P0 + 3) >> 3) & 1u);
          mbar_wait(&full[(P0 + 4) & (SP_RING - 1)], ((P0 + 4) >> 3) & 1u);
        }
        const uint32_t acc = grp & 1u;
        mbar_wait(&tempty[acc], ((grp >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d_lo = tmem_base + acc * 128u;           // odd row: columns 0..63, even row: 64..127
          const uint32_t d_hi = d_lo + 64u;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t slot = (P0 + ((t + 1) >> 1)) & (SP_RING - 1);
            const uint32_t al = a_lo0 + ((slot * SP_PAIR_BYTES + ((t + 1) & 1) * SP_ROW_BYTES) >> 4);
            if (k == 0) {
              // the even row alone: filter row t - 2 (W_0 = block 0, W_1 = upper half of block 2, W_r = block r-1)
              if (t < 2) continue;
              const int r = t - 2;
              const uint32_t bl = b_lo0 + (r == 0 ? 0u : r == 1 ? 5u * WB : static_cast<uint32_t>(2 * (r - 1)) * WB);
              umma_bf16_words(d_hi, al, hi_raw, bl, hi_w, idesc64, static_cast<uint32_t>(r != 0));
              umma_bf16_words(d_hi, al + 2, hi_raw, bl + 2, hi_w, idesc64, 1u);
            } else if (t == 0) {
              umma_bf16_words(d_lo, al, hi_raw, b_lo0, hi_w, idesc128, 0u);      // [W_0 ; 0]: zeroes the even row
              umma_bf16_words(d_lo, al + 2, hi_raw, b_lo0 + 2, hi_w, idesc64, 1u);
            } else if (t == 1) {
              umma_bf16_words(d_lo, al, hi_raw, b_lo0 + 5u * WB, hi_w, idesc64, 1u);
              umma_bf16_words(d_lo, al + 2, hi_raw, b_lo0 + 5u * WB + 2, hi_w, idesc64, 1u);
            } else if (t < 7) {
              const uint32_t bl = b_lo0 + static_cast<uint32_t>(2 * (t - 1)) * WB;     // [W_t ; W_{t-2}]
              umma_bf16_words(d_lo, al, hi_raw, bl, hi_w, idesc128, 1u);
              umma_bf16_words(d_lo, al + 2, hi_raw, bl + 2, hi_w, idesc128, 1u);
            } else {
              const uint32_t bl = b_lo0 + static_cast<uint32_t>(2 * (t - 3)) * WB;     // W_{t-2} = lower half of block t-3
              umma_bf16_words(d_hi, al, hi_raw, bl, hi_w, idesc64, 1u);
              umma_bf16_words(d_hi, al + 2, hi_raw, bl + 2, hi_w, idesc64, 1u);
            }
          }
          // input pairs no later group needs: group 0 -> pair 0; group k -> pairs 2k-1 and 2k
          if (k == 0) {
            umma_commit(&empty[g & (SP_RING - 1)]);
          } else {
            umma_commit(&empty[P0 & (SP_RING - 1)]);
            umma_commit(&empty[(P0 + 1) & (SP_RING - 1)]);
          }
          if (k == ngroups - 1) {                                  // end of the unit: release its last three pairs
            umma_commit(&empty[(P0 + 2) & (SP_RING - 1)]);
            umma_commit(&empty[(P0 + 3) & (SP_RING - 1)]);
            umma_commit(&empty[(P0 + 4) & (SP_RING - 1)]);
          }
          umma_commit(&tfull[acc]);
        }
        __syncwarp();
      }
      g += static_cast<uint32_t>(2 * ngroups + 2);     // = c1 rows + 3 pairs
    }
  } else {
    // ===================== epilogue + pool (warps 2..9) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;                    // c1 column inside the strip (0..127)
    const int etid = threadIdx.x - 64;
    const int ppx = etid >> 2;                        // pooled column inside the strip (pool phase, etid < 240)
    const int cg = etid & 3;                          // 16-channel group
    uint32_t grp = 0;                                 // running group counter (accumulator buffer, phases)
    uint32_t vctr = 0;                                // pooled rows produced so far (selects the column-max buffer)
    float bias_r[32];                                 // this thread's 32 output channels
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = sBias[half * 32 + j];
    const uint32_t sC_u32 = smem_u32(sC);
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int seg = u % p.n_segs;
      const int t = (u / p.n_segs) % p.n_strips;
      const int b = u / (p.n_segs * p.n_strips);
      const int py0 = seg * p.seg_rows;
      const int py1 = min(py0 + p.seg_rows, p.Hp);
      const int X = SP_STRIP * t - SP_HALO + row;
      const bool xvalid = (X >= 0) && (X < p.Wc);
      const int opx = SP_PCOLS * t + ppx;
      const bool pool_thread = (etid < 4 * SP_PCOLS) && (opx < p.Wpo);
      const int ngroups = (py1 - py0) + 1;
      uint4 he[4], ho[4];                             // this column's packed values of the last even / odd c1 row
#pragma unroll
      for (int j = 0; j < 4; ++j) he[j] = ho[j] = make_uint4(0u, 0u, 0u, 0u);
      for (int k = 0; k < ngroups; ++k, ++grp) {
        const uint32_t acc = grp & 1u;
        mbar_wait(&tfull[acc], (grp >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128u + static_cast<uint32_t>(half * 32);
        uint4 o[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {                 // h = 0: odd c1 row 2k-1 (not in group 0), h = 1: even row 2k
          if (h == 0 && k == 0) continue;
          const int Y = 2 * py0 - 1 + 2 * k - 1 + h;
          const bool valid = xvalid && (Y >= 0) && (Y < p.Hc);
          uint32_t v[32];
          tmem_ld_32x32(tcol + static_cast<uint32_t>(h * 64), v);
          tmem_ld_wait();
          if (h == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]); // accumulator is free for the group after next
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
#pragma unroll
            for (int e = 0; e < 8; e += 2) fadd2(f[e], f[e + 1], bias_r[8 * j + e], bias_r[8 * j + e + 1]);   // ReLU: in the conversion
            o[j].x = pack_bf16x2_relu(f[0], f[1]);
            o[j].y = pack_bf16x2_relu(f[2], f[3]);
            o[j].z = pack_bf16x2_relu(f[4], f[5]);
            o[j].w = pack_bf16x2_relu(f[6], f[7]);
            if (!valid) o[j] = make_uint4(0u, 0u, 0u, 0u);
          }
          if (h == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ho[j] = o[j];
          }
        }
        if (k > 0) {
          // column-wise max over c1 rows 2k-2, 2k-1, 2k -> shared memory; then pooled row py0 + k - 1 = the max over
          // strip columns 2*ppx + 7 .. + 9 of that row
          const uint32_t vbuf = sC_u32 + (vctr & 1u) * SP_C1_BYTES;
          const uint32_t crow = vbuf + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 m;
            m.x = act2_max(act2_max(o[j].x, ho[j].x), he[j].x);
            m.y = act2_max(act2_max(o[j].y, ho[j].y), he[j].y);
            m.z = act2_max(act2_max(o[j].z, ho[j].z), he[j].z);
            m.w = act2_max(act2_max(o[j].w, ho[j].w), he[j].w);
            sts128(crow + (((half * 4 + j) ^ c1_swz(row)) << 4), m);
          }
          named_bar_sync(1, 256);
          if (pool_thread) {
            uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0;        // 0 is neutral: every value is >= 0
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
              const int lc = 2 * ppx + (SP_HALO - 1) + dc;
              const uint32_t px = vbuf + lc * 128;
              const uint4 a0 = lds128(px + (((2 * cg) ^ c1_swz(lc)) << 4));
              const uint4 a1 = lds128(px + (((2 * cg + 1) ^ c1_swz(lc)) << 4));
              m0.x = act2_max(m0.x, a0.x); m0.y = act2_max(m0.y, a0.y);
              m0.z = act2_max(m0.z, a0.z); m0.w = act2_max(m0.w, a0.w);
              m1.x = act2_max(m1.x, a1.x); m1.y = act2_max(m1.y, a1.y);
              m1.z = act2_max(m1.z, a1.z); m1.w = act2_max(m1.w, a1.w);
            }
            const int py = py0 + k - 1;
            uint4* dst = p.y + ((static_cast<size_t>(b) * p.Hp + py) * p.Wpo + opx) * 8 + cg * 2;
            dst[0] = m0;
            dst[1] = m1;
          }
          ++vctr;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) he[j] = o[j];
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, SP_TMEM_COLS);
  }
}

}  // namespace

struct vd_stem_pool_plan {
  CUtensorMap tmA, tmB;
  StemPoolParams kp;
  int grid;
};

int vd_stem_pool_create(const vd_conv_desc_t* d, vd_stem_pool_plan** out) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(vd_tensor_map_encoder());
  if (!enc) {
    vd_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VD_ERR_NO_DEVICE;
  }
  VD_REQUIRE((d->flags & VD_CONV_IN_X0) && (d->flags & VD_CONV_RELU) && d->R == 7 && d->S == 7 && d->stride == 2 &&
                 d->pad == 3 && d->Cout == 64,
             "VD_CONV_POOL needs the 7x7/2 p3 64-channel ReLU stem on the ingest layout");
  VD_REQUIRE(!(d->flags & (VD_CONV_SPLIT | VD_CONV_RESIDUAL | VD_CONV_UPADD | VD_CONV_OUT_F32 | VD_CONV_DUAL)),
             "VD_CONV_POOL cannot be combined with split / residual / upadd / fp32 / dual");
  VD_REQUIRE(d->dst_C == 64 && d->dst_coff == 0, "VD_CONV_POOL writes a dense 64-channel tensor");
  VD_REQUIRE(d->Hout == (d->Hin + 1) / 2 && d->Wout == (d->Win + 1) / 2, "VD_CONV_POOL: Hout/Wout are conv1's extents");
  vd_stem_pool_plan* pl = new (std::nothrow) vd_stem_pool_plan();
  VD_REQUIRE(pl, "out of host memory");
  StemPoolParams& k = pl->kp;
  k.B = d->B; k.Hin = d->Hin; k.Hc = d->Hout; k.Wc = d->Wout;
  k.Hp = (k.Hc + 1) / 2; k.Wpo = (k.Wc + 1) / 2;
  k.n_strips = vd_cdiv(k.Wpo, SP_PCOLS);
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  // run length: long runs amortise the 6 extra input rows / 1 extra c1 row of a unit, short runs balance the
  // SMs; 15 pooled rows = 31 c1 rows costs 3 % and gives > 15 units per SM at the benchmark size
  k.seg_rows = 15;
  while (k.seg_rows > 4 && static_cast<long>(k.B) * k.n_strips * vd_cdiv(k.Hp, k.seg_rows) < 4L * sms) --k.seg_rows;
  k.n_segs = vd_cdiv(k.Hp, k.seg_rows);
  k.n_units = k.B * k.n_strips * k.n_segs;
  k.bias = d->bias;
  k.y = static_cast<uint4*>(d->y);
  pl->grid = k.n_units < sms ? k.n_units : sms;

  const CUtensorMapDataType dt = VD_ACT_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int Wp = vd_x0_pitch(d->Win);
  cuuint64_t gdim[4] = {64, static_cast<cuuint64_t>(Wp / 16), static_cast<cuuint64_t>(d->Hin),
                        static_cast<cuuint64_t>(d->B)};
  cuuint64_t gstr[3] = {128, static_cast<cuuint64_t>(Wp) * 8, static_cast<cuuint64_t>(Wp) * 8 * d->Hin};
  cuuint32_t box[4] = {64, SP_CHUNKS, 2, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&pl->tmA, dt, 4, const_cast<void*>(d->x), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(stem A) failed: CUresult %d", static_cast<int>(r));
    delete pl;
    return VD_ERR_CUDA;
  }
  cuuint64_t wdim[2] = {224, 64};
  cuuint64_t wstr[1] = {448};
  cuuint32_t wbox[2] = {32, 64};
  cuuint32_t wes[2] = {1, 1};
  r = enc(&pl->tmB, dt, 2, const_cast<void*>(d->w), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(stem B) failed: CUresult %d", static_cast<int>(r));
    delete pl;
    return VD_ERR_CUDA;
  }
  *out = pl;
  return VD_OK;
}

void vd_stem_pool_destroy(vd_stem_pool_plan* pl) { delete pl; }

void vd_stem_pool_info(const vd_stem_pool_plan* pl, int* units, int* grid) {
  if (units) *units = pl->kp.n_units;
  if (grid) *grid = pl->grid;
}

int vd_stem_pool_run(const vd_stem_pool_plan* pl, cudaStream_t st) {
  VD_SET_FUNC_ATTR_ONCE_PER_DEVICE(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM);
  stem_pool_kernel<<<pl->grid, SP_THREADS, SP_SMEM, st>>>(pl->tmA, pl->tmB, pl->kp);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
