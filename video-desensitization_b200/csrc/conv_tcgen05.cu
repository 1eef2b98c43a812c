// conv_tcgen05.cu — K2: convolution as implicit GEMM on the 5th-gen tensor cores.
//
// Replaces every conv+BN(+ReLU)(+residual)(+FPN top-down add) of RetinaFace.forward
// (reference detect_face/retinaface.py:114-148, detect_face/nets/layers.py:10-114, torchvision
// resnet50 Bottleneck) with ONE persistent, warp-specialised kernel:
//
//   D[pixel, co] = sum_{r,s,ci} X[n, y*sh-ph+r, x*sw-pw+s, ci] * W[co, r, s, ci]
//
//   * activations are bf16 NHWC; an M-tile is a TW x TH x NB box of output pixels (<= 128 rows);
//     for every filter tap (r,s) and 64-channel slice the A operand is fetched by ONE 4-D TMA
//     tile load at shifted coordinates — TMA's out-of-bounds zero fill implements the padding and
//     `elementStrides` implements the convolution stride, so there is no im2col buffer anywhere;
//   * each pixel's 64 channels are one 128-byte row of a SWIZZLE_128B K-major UMMA operand;
//   * weights [Cout][R*S*Cin] bf16 (BN folded) are the K-major B operand, 2-D TMA;
//   * tcgen05.mma (M=128, N=BN, K=16) accumulates fp32 in TMEM; two accumulator buffers so the
//     epilogue of tile i overlaps the MMAs of tile i+1;
//   * epilogue (8 warps, two per TMEM lane quarter): tcgen05.ld -> +bias -> (+residual, prefetched
//     from global while the MMAs still run) -> ReLU -> (+nearest-upsampled FPN tensor) -> bf16 (or
//     fp32 for the head GEMM) -> SWIZZLE_128B staging tile in shared memory -> TMA tile store
//     (full 128-byte lines per pixel, clipped by the tensor map at the image borders).
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
#include <cuda.h>
#include <cstdlib>
#include <new>

#include "common.cuh"
#include "stem_pool.h"

namespace {

struct ConvKParams {
  int TW, TH, NB;
  int tiles_x, tiles_y, tiles_b, n_tiles, num_tiles;
  uint32_t dv_n_mul, dv_n_shr, dv_x_mul, dv_x_shr, dv_y_mul, dv_y_shr;   // tile index -> (n, x, y, b) without IDIV
  int rows_used;          // TW*TH*NB
  int a_bytes;            // rows_used * 128
  int B, Hout, Wout, Cout;
  int R, S, cin_blocks;
  int stride_h, stride_w, pad_h, pad_w;
  int dst_C, dst_coff;
  int up_H, up_W;
  int kb2;                // DUAL: K blocks (Cin2 / 64) of the second input
  int kb1;                // K blocks of the main convolution; blocks >= kb1 come from the second input (DUAL)
  int stride2;            // DUAL: spatial stride of the 1x1 projection of x2
  int patch;              // 1: 3x3 stride-1 "halo patch" mode — the (TH+2)x(TW+2) input patch of a 64-channel
                          //    slice is loaded ONCE and the nine taps read it through row-shifted descriptors
  int PW, PH;             // patch extent in pixels
  int a_slots, a_slot_bytes;
  int kps;                // K blocks (generic) / filter taps (patch mode) fetched and issued per pipeline stage:
                          //    one mbarrier round trip costs the single issuing warp ~450 clk, a 64-wide MMA
                          //    K block only 128 clk, so narrow layers batch several blocks per stage
  int stages;             // operand pipeline depth (<= kMaxStages), chosen on the host from the smem budget
  int wres;               // 1: the whole weight matrix of the (single) N tile stays resident in shared memory
  int w_bytes;            // wres: KB * BN * 128
  int reverse;            // VD_CONV_REVERSE: tiles are visited from the last to the first
  int relu_pack;          // 1: ReLU folded into the bf16 conversion (cvt.rn.relu) where the layer allows it
  uint32_t flags;
  const float* bias;
  const float* slope;     // VD_CONV_SLOPE: per-channel negative-side factor
  const float* bias2;     // VD_CONV_CHAIN: folded-BN bias of the chained 1x1 convolution
  int N2;                 // VD_CONV_CHAIN: its output channels (64 or 128)
  int w2_bytes;           // VD_CONV_CHAIN: N2 * Cout * 2 (its whole weight matrix, resident in shared memory)
  const __nv_bfloat16* residual;
  const __nv_bfloat16* upadd;
  void* y;
  float acc_scale;        // VD_CONV_SPLIT: accumulators are multiplied by this (a power of two) before the bias is added
  int cout1;              // VD_CONV_TWO_DST: output channels >= cout1 go to the second destination (tmY2) ...
  int dst_b_coff;         //                  ... starting at this channel of it
};

// RES = the layer adds a residual tensor: it is prefetched by a dedicated TMA warp into a ring of
// RES_SLOTS 16 KB tiles (one per 64-channel chunk), paid for with fewer operand stages.
constexpr CUtensorMapDataType kActTmapType =
    VD_ACT_IS_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 232448;        // 227 KB per CTA on sm_100
constexpr int kWresMaxBytes = 72 * 1024;   // resident-weight limit (3x3 64->64 = 72 KB)

// ROWB = bytes of one operand row (one K block): 128 (64 bf16, SWIZZLE_128B) everywhere except the stem conv1,
// whose K block is one filter row of 8 taps x 4 channels = 64 bytes (SWIZZLE_64B).
template <int BN, bool RES, int ROWB = 128>
struct ConvCfg {
  // operand slot of one K block; the stem's slot holds the seven raw row segments of one tile (each <= 17 chunks of
  // 128 B, see hi_raw), fetched by ONE tensor-map box
  static constexpr int A_BYTES = (ROWB == 64) ? 7 * 17 * 128 + 128 : 128 * ROWB;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int BAR_BYTES = 384;
  static constexpr int OUT_BYTES = 2 * 128 * 128;                             // two staging tiles for the TMA store
  static constexpr int RES_SLOTS = 3;
  static constexpr int RES_BYTES = RES ? RES_SLOTS * 128 * 128 : 0;
  static constexpr int BIAS_BYTES = 1024;
  static constexpr int FIXED_BYTES = OUT_BYTES + RES_BYTES + BIAS_BYTES + BAR_BYTES + 1024;  // +1024: alignment
  static constexpr int THREADS = RES ? 352 : 320;                             // +1 warp: residual producer
};

constexpr int kEpiThreads = 256;

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows*128B)
// | version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
// The hardware applies the 128-byte swizzle to ABSOLUTE shared-memory address bits (csrc/probe_umma.cu,
// measured on B200): a descriptor may start at any 128-byte row of a TMA-written tile and use any
// multiple of 128 as group stride, with base_offset = 0.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// q = n / d for n < 2^31 with a host-made (mul, shr) pair (d == 1 is encoded as mul = 0)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t mul, uint32_t shr) {
  return mul ? (__umulhi(n, mul) >> shr) : n;
}

struct TileCoord { int n_tile, tx, ty, tb; };

__device__ __forceinline__ TileCoord tile_coord(const ConvKParams& p, int tile) {
  TileCoord c;
  uint32_t m = fast_div(static_cast<uint32_t>(tile), p.dv_n_mul, p.dv_n_shr);
  c.n_tile = tile - static_cast<int>(m) * p.n_tiles;
  uint32_t m2 = fast_div(m, p.dv_x_mul, p.dv_x_shr);
  c.tx = static_cast<int>(m) - static_cast<int>(m2) * p.tiles_x;
  uint32_t m3 = fast_div(m2, p.dv_y_mul, p.dv_y_shr);
  c.ty = static_cast<int>(m2) - static_cast<int>(m3) * p.tiles_y;
  c.tb = static_cast<int>(m3);
  return c;
}

// tile visited by this CTA in iteration `it`.  CHAIN: the N tiles of one M tile run back to back on the same CTA (the
// chained convolution accumulates over all of them), M tiles are dealt round-robin.
template <bool CHAIN>
__device__ __forceinline__ int tile_at(const ConvKParams& p, int it) {
  if (CHAIN) return (static_cast<int>(blockIdx.x) + (it >> 1) * static_cast<int>(gridDim.x)) * 2 + (it & 1);
  const int idx = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
  return (p.reverse && idx < p.num_tiles) ? p.num_tiles - 1 - idx : idx;
}

// CHAIN (VD_CONV_CHAIN, BN = 128, Cout = 256): the kernel also computes the NEXT layer, a 1x1 convolution
// Cout -> N2 (+bias, ReLU) of its own output — bottleneck conv3 followed by the next bottleneck's conv1.  Every
// 64-channel output chunk already sits in shared memory as a SWIZZLE_128B 128-row tile for the TMA store; that tile IS
// the K-major A operand of one K block of the chained GEMM, so the MMA warp accumulates it (4 MMAs, B = resident
// W2 block) into a second TMEM accumulator while the store drains, and the wide tensor is never read back from HBM.
template <int BN, bool RES, int ROWB, bool CHAIN>
__global__ void __launch_bounds__((ConvCfg<BN, RES, ROWB>::THREADS), (BN <= 128 && !RES && !CHAIN) ? 2 : 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const __grid_constant__ CUtensorMap tmY2, const ConvKParams p) {
  using Cfg = ConvCfg<BN, RES, ROWB>;
  static_assert(!CHAIN || (BN == 128 && ROWB == 128), "CHAIN: two 128-wide N tiles per M tile");
  constexpr int kTmemCols = CHAIN ? 512 : Cfg::TMEM_COLS;      // CHAIN: 2 x 128 (main) + 2 x N2 (chained), N2 <= 128
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  uint8_t* sA = smem;                                     // A stages (or the ring of halo patches)
  const int KPS = p.kps;
  uint8_t* sB = smem + (p.patch ? p.a_slots * p.a_slot_bytes : STAGES * KPS * Cfg::A_BYTES);   // B stages / resident W
  uint8_t* sOut = sB + (p.wres ? p.w_bytes : STAGES * KPS * Cfg::B_BYTES);
  uint8_t* sRes = sOut + Cfg::OUT_BYTES;
  uint8_t* sW2 = sRes + Cfg::RES_BYTES;                   // CHAIN: resident weights of the chained convolution
  uint8_t* sOut2 = sW2 + (CHAIN ? p.w2_bytes : 0);        // CHAIN: staging tile of the chained output
  float* sBias2 = reinterpret_cast<float*>(sOut2 + (CHAIN ? 128 * 128 : 0));      // CHAIN: bias of the chained conv (512 B)
  float* sBias = sBias2 + (CHAIN ? 128 : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tfull = empty + kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint64_t* rempty = rfull + Cfg::RES_SLOTS;
  uint64_t* wfull = rempty + Cfg::RES_SLOTS;
  uint64_t* afull = wfull + 1;                            // patch ring (<= 4 slots)
  uint64_t* aempty = afull + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 4);
  uint64_t* ystaged = full + 38;                          // CHAIN: output chunk staged in sOut[k] (epilogue -> MMA warp)
  uint64_t* yfree = ystaged + 2;                          //        chained MMAs have read sOut[k]
  uint64_t* t1full = yfree + 2;                           //        chained accumulator complete
  uint64_t* t1empty = t1full + 2;                         //        chained accumulator drained by the epilogue
  uint64_t* w2full = t1empty + 2;                         //        W2 landed
  float* sSlope = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + Cfg::BAR_BYTES);   // VD_CONV_SLOPE only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (RES) tma_prefetch_desc(&tmR);
    if (p.flags & VD_CONV_DUAL) tma_prefetch_desc(&tmA2);
    if (CHAIN) {
      tma_prefetch_desc(&tmB2);
      tma_prefetch_desc(&tmY2);
    }
    if (!CHAIN && p.cout1 > 0) tma_prefetch_desc(&tmY2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(wfull, 1);
      for (int i = 0; i < 4; ++i) {
        mbar_init(&afull[i], 1);
        mbar_init(&aempty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 8);
      }
      for (int i = 0; i < Cfg::RES_SLOTS; ++i) {
        mbar_init(&rfull[i], 1);
        mbar_init(&rempty[i], 8);
      }
      if (CHAIN) {
        for (int i = 0; i < 2; ++i) {
          mbar_init(&ystaged[i], 1);
          mbar_init(&yfree[i], 1);
          mbar_init(&t1full[i], 1);
          mbar_init(&t1empty[i], 8);
        }
        mbar_init(w2full, 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // Programmatic dependent launch: the next kernel of the stream may start its own prologue on SMs this grid has
  // already left; everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of
  // the previous kernel, and no global memory is touched before the wait below.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int KB = p.patch ? p.R * p.S * p.cin_blocks : p.kb1 + ((p.flags & VD_CONV_DUAL) ? p.kb2 : 0);

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) ==========
    int stage = 0;
    uint32_t phase = 0;
    int aslot = 0;
    uint32_t aphase = 0;
    if (p.wres && elect_one()) {                      // weights: loaded once, reused by every tile
      mbar_expect_tx(wfull, static_cast<uint32_t>(p.w_bytes));
      // several N tiles: the grid is a multiple of n_tiles (plan_smem), so this CTA only ever sees one of them
      // (num_tiles = m_tiles * n_tiles, so this also holds for the reversed walk)
      const int t0 = tile_at<CHAIN>(p, 0);
      const int wrow = (p.n_tiles > 1 && t0 < p.num_tiles) ? tile_coord(p, t0).n_tile * BN : 0;
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + kb * Cfg::B_BYTES, &tmB, wfull, kb * (ROWB / 2), wrow);
    }
    if (CHAIN && elect_one()) {                       // chained convolution's weights: [N2][Cout], one box per K block
      mbar_expect_tx(w2full, static_cast<uint32_t>(p.w2_bytes));
      for (int kb = 0; kb < p.Cout / 64; ++kb) tma_load_2d(sW2 + kb * p.N2 * 128, &tmB2, w2full, kb * 64, 0);
    }
    __syncwarp();
    const uint32_t stage_tx = static_cast<uint32_t>(p.a_bytes + (p.wres ? 0 : Cfg::B_BYTES));
    for (int it = 0, tile; (tile = tile_at<CHAIN>(p, it)) < p.num_tiles; ++it) {
      const TileCoord tc = tile_coord(p, tile);
      const int n_tile = tc.n_tile, tx = tc.tx, ty = tc.ty, tb = tc.tb;
      const int xin0 = tx * p.TW * p.stride_w - p.pad_w;
      const int yin0 = ty * p.TH * p.stride_h - p.pad_h;
      const int b0 = tb * p.NB;
      const int ncol = n_tile * BN;
      if (p.patch) {
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait(&aempty[aslot], aphase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&afull[aslot], static_cast<uint32_t>(p.PW * p.PH * 128));
            tma_load_4d(sA + aslot * p.a_slot_bytes, &tmA, &afull[aslot], cb * 64, xin0, yin0, b0);
          }
          __syncwarp();
          if (++aslot == p.a_slots) {
            aslot = 0;
            aphase ^= 1u;
          }
          for (int tap = 0; tap < 9 && !p.wres; tap += KPS) {
            mbar_wait(&empty[stage], phase ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(&full[stage], static_cast<uint32_t>(KPS * Cfg::B_BYTES));
              for (int j = 0; j < KPS; ++j)
                tma_load_2d(sB + (stage * KPS + j) * Cfg::B_BYTES, &tmB, &full[stage],
                            ((tap + j) * p.cin_blocks + cb) * 64, ncol);
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        continue;
      }
      if (ROWB == 64) {
        // stem: the RAW padded input rows 2y-3 .. 2y+3, pixels [2*X0, 2*X0 + 2*TW + 6) (8 B each) as whole 128-byte
        // chunks — one box per tile; the MMAs read it as a sliding window (see hi_raw below)
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], static_cast<uint32_t>(7 * p.a_bytes));
          tma_load_4d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], 0, (tx * p.TW) >> 3, yin0, b0);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
        continue;
      }
      int r = 0, sx = 0, cb = 0;
      for (int kb = 0; kb < KB; kb += KPS) {
        const int n = min(KPS, KB - kb);
        mbar_wait(&empty[stage], phase ^ 1u);
        const bool leader = elect_one() != 0;
        if (leader) mbar_expect_tx(&full[stage], static_cast<uint32_t>(n) * stage_tx);
        for (int j = 0; j < n; ++j) {
          if (leader) {
            if (kb + j < p.kb1)
              tma_load_4d(sA + (stage * KPS + j) * Cfg::A_BYTES, &tmA, &full[stage], cb * 64, xin0 + sx, yin0 + r, b0);
            else        // second K segment: 1x1 projection of x2 at (x*stride2, y*stride2)
              tma_load_4d(sA + (stage * KPS + j) * Cfg::A_BYTES, &tmA2, &full[stage], (kb + j - p.kb1) * 64,
                          tx * p.TW * p.stride2, ty * p.TH * p.stride2, b0);
            if (!p.wres)
              tma_load_2d(sB + (stage * KPS + j) * Cfg::B_BYTES, &tmB, &full[stage], (kb + j) * (ROWB / 2), ncol);
          }
          if (++cb == p.cin_blocks) {            // K order (r, s, cin block): advance for every lane alike
            cb = 0;
            if (++sx == p.S) {
              sx = 0;
              ++r;
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =============
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13),
    // a/b K-major, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t kFmt = VD_ACT_IS_FP16 ? 0u : 1u;      // a/b format: 0 = F16, 1 = BF16
    constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                               (static_cast<uint32_t>(128 >> 4) << 24);
    // descriptor words: lo = start>>4 | LBO(1)<<16 ; hi = SBO>>4 | version(1)<<14 | SWIZZLE_128B(2)<<29
    constexpr uint32_t hi_std = (static_cast<uint32_t>(8 * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
    // stem (ROWB == 64): A is the raw input row in shared memory, NO swizzle, read through the canonical K-major
    // layout with LBO = 16 B (next 16-byte chunk along K) and SBO = 128 B (next 8 rows): row i,
    // chunk c sits at 16*i + 16*c — output pixel i's window starts 2 input pixels (16 B) after pixel i-1's, so the
    // overlapping windows are never materialised (one 2 KB row segment instead of a 128 x 64 B operand tile).
    constexpr uint32_t hi_raw = (128u >> 4) | (1u << 14);
    constexpr uint32_t hi_a = (ROWB == 64) ? hi_raw : hi_std;
    constexpr int MMAS = ROWB / 32;             // K = 16 bf16 = 32 bytes per tcgen05.mma
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int aslot = 0;
    uint32_t aphase = 0;
    if (p.wres) mbar_wait(wfull, 0);
    if (CHAIN) mbar_wait(w2full, 0);      // also keeps a CTA without tiles alive until its W2 copy has landed
    // CHAIN: K blocks of the chained GEMM, one per staged 64-channel output chunk, issued one iteration behind the
    // main MMAs (the epilogue of iteration it-1 stages them while the MMAs of iteration `it` already run)
    int acc1 = 0;
    uint32_t acc1_phase = 0;
    uint32_t ychunk = 0;
    const uint32_t idesc2 = (1u << 4) | (kFmt << 7) | (kFmt << 10) | (static_cast<uint32_t>(p.N2 >> 3) << 17) |
                            (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t o_lo0 = ((smem_u32(sOut) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t w2_lo0 = ((smem_u32(sW2) & 0x3FFFFu) >> 4) | (1u << 16);
    auto chain_issue = [&](int it_prev) {
      constexpr int CPT = BN / 64;                       // chunks per tile
      const int n = it_prev & 1;
      for (int c = 0; c < CPT; ++c, ++ychunk) {
        const int g = n * CPT + c;                       // K block of the chained GEMM
        if (g == 0) mbar_wait(&t1empty[acc1], acc1_phase ^ 1u);
        const uint32_t buf = ychunk & 1u;
        mbar_wait(&ystaged[buf], (ychunk >> 1) & 1u);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t d2 = tmem_base + static_cast<uint32_t>(2 * BN + acc1 * p.N2);
          const uint32_t al = o_lo0 + buf * static_cast<uint32_t>((128 * 128) >> 4);
          const uint32_t bl = w2_lo0 + static_cast<uint32_t>(g * ((p.N2 * 128) >> 4));
          umma_bf16_words(d2, al, hi_std, bl, hi_std, idesc2, static_cast<uint32_t>(g != 0));
          umma_bf16_words(d2, al + 2, hi_std, bl + 2, hi_std, idesc2, 1u);
          umma_bf16_words(d2, al + 4, hi_std, bl + 4, hi_std, idesc2, 1u);
          umma_bf16_words(d2, al + 6, hi_std, bl + 6, hi_std, idesc2, 1u);
          umma_commit(&yfree[buf]);
          if (g == 2 * CPT - 1) umma_commit(&t1full[acc1]);
        }
        __syncwarp();
        if (g == 2 * CPT - 1) {
          acc1 ^= 1;
          if (acc1 == 0) acc1_phase ^= 1u;
        }
      }
    };
    int it_done = 0;
    for (int it = 0, tile; (tile = tile_at<CHAIN>(p, it)) < p.num_tiles; ++it) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      if (p.patch) {
        const uint32_t hi_patch = (static_cast<uint32_t>(p.PW * 128) >> 4) | (1u << 14) | (2u << 29);
        for (int cb = 0; cb < p.cin_blocks; ++cb) {
          mbar_wait(&afull[aslot], aphase);
          const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(aslot * (p.a_slot_bytes >> 4));
          if (p.wres) {
            // resident weights: the nine taps of this 64-channel slice issue back to back
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const int r = tap / 3, sx = tap - 3 * r;
                const uint32_t al = a_lo + static_cast<uint32_t>((r * p.PW + sx) * 8);
                const uint32_t bl = b_lo0 + static_cast<uint32_t>((tap * p.cin_blocks + cb) * (Cfg::B_BYTES >> 4));
                umma_bf16_words(d_tmem, al, hi_patch, bl, hi_std, idesc, static_cast<uint32_t>((cb | tap) != 0));
                umma_bf16_words(d_tmem, al + 2, hi_patch, bl + 2, hi_std, idesc, 1u);
                umma_bf16_words(d_tmem, al + 4, hi_patch, bl + 4, hi_std, idesc, 1u);
                umma_bf16_words(d_tmem, al + 6, hi_patch, bl + 6, hi_std, idesc, 1u);
              }
              umma_commit(&aempty[aslot]);
              if (cb == p.cin_blocks - 1) umma_commit(&tfull[acc]);
            }
            __syncwarp();
          }
          for (int tap0 = 0; tap0 < 9 && !p.wres; tap0 += KPS) {
            mbar_wait(&full[stage], phase);
            tcgen05_fence_after();
            if (elect_one()) {
              for (int j = 0; j < KPS; ++j) {
                const int tap = tap0 + j;
                const int r = tap / 3, sx = tap - 3 * r;
                const uint32_t al = a_lo + static_cast<uint32_t>((r * p.PW + sx) * 8);     // 128 B rows, >>4
                const uint32_t bl = b_lo0 + static_cast<uint32_t>((stage * KPS + j) * (Cfg::B_BYTES >> 4));
                umma_bf16_words(d_tmem, al, hi_patch, bl, hi_std, idesc, static_cast<uint32_t>((cb | tap) != 0));
                umma_bf16_words(d_tmem, al + 2, hi_patch, bl + 2, hi_std, idesc, 1u);
                umma_bf16_words(d_tmem, al + 4, hi_patch, bl + 4, hi_std, idesc, 1u);
                umma_bf16_words(d_tmem, al + 6, hi_patch, bl + 6, hi_std, idesc, 1u);
              }
              umma_commit(&empty[stage]);
              if (tap0 + KPS >= 9) umma_commit(&aempty[aslot]);   // patch slot free once all nine taps have read it
              if (tap0 + KPS >= 9 && cb == p.cin_blocks - 1) umma_commit(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if (++aslot == p.a_slots) {
            aslot = 0;
            aphase ^= 1u;
          }
        }
      } else if (ROWB == 64) {
        // stem: one stage = one tile; K block kb reads raw row kb / cin_blocks (split mode: [w_hi | w_lo] per row)
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a_row = static_cast<uint32_t>(p.a_bytes >> 4);
          uint32_t al = a_lo0 + static_cast<uint32_t>(stage * (Cfg::A_BYTES >> 4));
          uint32_t bl = b_lo0;
          for (int r = 0; r < 7; ++r, al += a_row) {
            for (int cb = 0; cb < p.cin_blocks; ++cb, bl += (Cfg::B_BYTES >> 4)) {
              umma_bf16_words(d_tmem, al, hi_a, bl, hi_std, idesc, static_cast<uint32_t>((r | cb) != 0));
              umma_bf16_words(d_tmem, al + 2, hi_a, bl + 2, hi_std, idesc, 1u);
            }
          }
          umma_commit(&empty[stage]);
          umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      } else {
        for (int kb = 0; kb < KB; kb += KPS) {
          const int n = min(KPS, KB - kb);
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            for (int j = 0; j < n; ++j) {
              const uint32_t al = a_lo0 + static_cast<uint32_t>((stage * KPS + j) * (Cfg::A_BYTES >> 4));
              const uint32_t bl =
                  b_lo0 + static_cast<uint32_t>((p.wres ? kb + j : stage * KPS + j) * (Cfg::B_BYTES >> 4));
              // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in 16-byte units
              umma_bf16_words(d_tmem, al, hi_a, bl, hi_std, idesc, static_cast<uint32_t>((kb + j) != 0));
#pragma unroll
              for (int q = 1; q < MMAS; ++q)
                umma_bf16_words(d_tmem, al + 2 * q, hi_a, bl + 2 * q, hi_std, idesc, 1u);
            }
            umma_commit(&empty[stage]);   // frees the smem slot once these MMAs have read it
            if (kb + KPS >= KB) umma_commit(&tfull[acc]);       // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      if (CHAIN) {
        if (it > 0) chain_issue(it - 1);
        it_done = it + 1;
      }
    }
    if (CHAIN && it_done > 0) chain_issue(it_done - 1);
  } else if (RES && warp == 10) {
    // ===================== residual producer =====================
    if (lane == 0) {
      constexpr int NCHUNK_R = (BN >= 64) ? BN / 64 : 1;
      int slot = 0;
      uint32_t rphase = 0;
      for (int it = 0, tile; (tile = tile_at<CHAIN>(p, it)) < p.num_tiles; ++it) {
        const TileCoord tc = tile_coord(p, tile);
        const int n_tile = tc.n_tile, tx = tc.tx, ty = tc.ty, tb = tc.tb;
        const int nplanes = (p.flags & VD_CONV_SPLIT) ? 2 : 1;      // split residual = hi plane + lo plane
        for (int ch = 0; ch < NCHUNK_R; ++ch) {
          for (int pl = 0; pl < nplanes; ++pl) {
            mbar_wait(&rempty[slot], rphase ^ 1u);
            mbar_expect_tx(&rfull[slot], static_cast<uint32_t>(p.rows_used * 128));
            tma_load_4d(sRes + slot * (128 * 128), &tmR, &rfull[slot], pl * p.Cout + n_tile * BN + ch * 64, tx * p.TW,
                        ty * p.TH, tb * p.NB);
            if (++slot == Cfg::RES_SLOTS) {
              slot = 0;
              rphase ^= 1u;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // which 32 of the chunk's 64 columns this warp converts
    const int row = q * 32 + lane;
    const int tx_ = row % p.TW;
    const int ty_ = (row / p.TW) % p.TH;
    const int nb_ = row / (p.TW * p.TH);
    const bool issuer = (threadIdx.x == 64);        // lane 0 of warp 2 issues the TMA stores
    const bool f32out = (p.flags & VD_CONV_OUT_F32) != 0;
    const bool split = (p.flags & VD_CONV_SPLIT) != 0;
    const int up_C = split ? 3 * p.Cout : p.Cout;        // channels of the upadd tensor
    // plain ReLU layers with a bf16/fp16 output: the ReLU rides on the conversion instruction (cvt.rn.relu)
    const bool relu_in_pack = (p.flags & VD_CONV_RELU) && !(p.flags & (VD_CONV_SLOPE | VD_CONV_UPADD)) && !f32out && !split &&
                              p.relu_pack;
    constexpr int NCHUNK = (BN >= 64) ? BN / 64 : 1;
    const bool active = (BN >= 64) || (half == 0);  // BN == 32 (head GEMM): one 32-column chunk
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    int rslot = 0;
    uint32_t rphase = 0;
    int bias_tile = -1;
    const int etid = threadIdx.x - 64;               // 0..255 inside the epilogue group
    int acc1e = 0;                                   // CHAIN: chained accumulator being drained
    uint32_t acc1e_phase = 0;
    int pend_tx = 0, pend_ty = 0, pend_tb = 0;       // CHAIN: M tile whose chained epilogue is still owed
    bool pending = false;
    if (CHAIN) {                                     // both N tiles' biases and the chained bias, once
      sBias[etid] = __ldg(p.bias + etid);
      if (etid < p.N2) sBias2[etid] = __ldg(p.bias2 + etid);
      named_bar_sync(1, kEpiThreads);
    }
    // ---- chained convolution's epilogue for one M tile: all its K blocks are in -> +bias2, ReLU, store ----
    auto chain_epilogue = [&](int ptx, int pty, int ptb) {
      mbar_wait(&t1full[acc1e], acc1e_phase);
      tcgen05_fence_after();
      for (int c2 = 0; c2 < p.N2 / 64; ++c2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                          static_cast<uint32_t>(2 * BN + acc1e * p.N2 + c2 * 64 + half * 32), v);
        tmem_ld_wait();
        float f[32];
        const uint32_t bp = smem_u32(sBias2) + static_cast<uint32_t>((c2 * 64 + half * 32) * 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 bv = lds128(bp + j * 16);
          f[4 * j + 0] = __uint_as_float(v[4 * j + 0]);
          f[4 * j + 1] = __uint_as_float(v[4 * j + 1]);
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]);
          f[4 * j + 3] = __uint_as_float(v[4 * j + 3]);
          fadd2(f[4 * j + 0], f[4 * j + 1], __uint_as_float(bv.x), __uint_as_float(bv.y));   // ReLU: in the conversion below
          fadd2(f[4 * j + 2], f[4 * j + 3], __uint_as_float(bv.z), __uint_as_float(bv.w));
        }
        if (issuer) {                                       // previous store out of sOut2 has read it (the newest
          if (c2 == 0) tma_store_wait_read<1>();            // group is the main output chunk staged just before)
          else tma_store_wait_read<0>();
        }
        named_bar_sync(1, kEpiThreads);
        const uint32_t srow = smem_u32(sOut2) + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2_relu(f[8 * j + 0], f[8 * j + 1]);
          o.y = pack_bf16x2_relu(f[8 * j + 2], f[8 * j + 3]);
          o.z = pack_bf16x2_relu(f[8 * j + 4], f[8 * j + 5]);
          o.w = pack_bf16x2_relu(f[8 * j + 6], f[8 * j + 7]);
          sts128(srow + (((half * 4 + j) ^ (row & 7)) << 4), o);
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpiThreads);
        if (issuer) {
          tma_store_4d(&tmY2, sOut2, c2 * 64, ptx * p.TW, pty * p.TH, ptb * p.NB);
          tma_store_commit();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t1empty[acc1e]);
      acc1e ^= 1;
      if (acc1e == 0) acc1e_phase ^= 1u;
    };
    for (int it = 0, tile; (tile = tile_at<CHAIN>(p, it)) < p.num_tiles; ++it) {
      const TileCoord tc = tile_coord(p, tile);
      const int n_tile = tc.n_tile, tx = tc.tx, ty = tc.ty, tb = tc.tb;
      const int x = tx * p.TW + tx_;
      const int y = ty * p.TH + ty_;
      const int b = tb * p.NB + nb_;
      const bool valid = (row < p.rows_used) && (x < p.Wout) && (y < p.Hout) && (b < p.B);
      size_t up_pix = 0;
      if (p.flags & VD_CONV_UPADD) {
        const int uy = min(y >> 1, p.up_H - 1), ux = min(x >> 1, p.up_W - 1);
        up_pix = (static_cast<size_t>(b) * p.up_H + uy) * p.up_W + ux;
      }
      const int ncol0 = n_tile * BN + half * 32;       // first output channel of this thread in chunk 0
      if (!CHAIN && n_tile != bias_tile) {             // folded-BN bias of this N tile -> shared memory
        named_bar_sync(1, kEpiThreads);
        if (etid < BN) {
          sBias[etid] = __ldg(p.bias + n_tile * BN + etid);
          if (p.flags & VD_CONV_SLOPE) sSlope[etid] = __ldg(p.slope + n_tile * BN + etid);
        }
        named_bar_sync(1, kEpiThreads);
        bias_tile = n_tile;
      }

      mbar_wait(&tfull[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < NCHUNK; ++ch, ++chunk_ctr) {
        const int n0 = ncol0 + ch * 64;
        uint32_t v[32];
        if (active) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                            static_cast<uint32_t>(acc * BN + ch * 64 + half * 32), v);
        }
        // FPN top-down term (nearest-upsampled coarser level): the four 16-byte loads are issued here, behind the
        // TMEM load, so that their L2 latency overlaps the accumulator read instead of following it
        uint4 upv[4];
        // (only in the BN = 256 one-CTA-per-SM variant the FPN laterals use: the 96-register two-CTA variants would spill)
        constexpr bool UP_EARLY = (BN == 256) && !RES && !CHAIN;
        const bool up_early = UP_EARLY && (p.flags & VD_CONV_UPADD) && !split && active && valid;
        if (up_early) {
          const uint4* up = reinterpret_cast<const uint4*>(p.upadd + up_pix * up_C + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) upv[j] = __ldg(up + j);
        }
        uint4 rcur[4], rlo[4];
        if (RES) {
          // residual chunk staged by the residual-producer warp (SWIZZLE_128B tile, same geometry as the output);
          // split mode: a second tile carries the lo plane
          const int nplanes = split ? 2 : 1;
          for (int pl = 0; pl < nplanes; ++pl) {
            mbar_wait(&rfull[rslot], rphase);
            const uint32_t rrow = smem_u32(sRes) + rslot * (128 * 128) + row * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 t = lds128(rrow + (((half * 4 + j) ^ (row & 7)) << 4));
              if (pl == 0) rcur[j] = t; else rlo[j] = t;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&rempty[rslot]);
            if (++rslot == Cfg::RES_SLOTS) {
              rslot = 0;
              rphase ^= 1u;
            }
          }
        }
        float f[32];
        if (active) {
          tmem_ld_wait();
          if (split) {             // weights were packed as w * 2^k so that their lo plane stays a NORMAL fp16 number
            const float sc = p.acc_scale;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * sc);
          }
          const uint32_t bp = smem_u32(sBias) + static_cast<uint32_t>(((CHAIN ? n_tile * BN : 0) + ch * 64 + half * 32) * 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 bv = lds128(bp + j * 16);
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]);
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]);
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]);
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]);
            fadd2(f[4 * j + 0], f[4 * j + 1], __uint_as_float(bv.x), __uint_as_float(bv.y));
            fadd2(f[4 * j + 2], f[4 * j + 3], __uint_as_float(bv.z), __uint_as_float(bv.w));
          }
          if (RES) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 rv = rcur[j];
              fadd2(f[8 * j + 0], f[8 * j + 1], bf16lo_to_f32(rv.x), bf16hi_to_f32(rv.x));
              fadd2(f[8 * j + 2], f[8 * j + 3], bf16lo_to_f32(rv.y), bf16hi_to_f32(rv.y));
              fadd2(f[8 * j + 4], f[8 * j + 5], bf16lo_to_f32(rv.z), bf16hi_to_f32(rv.z));
              fadd2(f[8 * j + 6], f[8 * j + 7], bf16lo_to_f32(rv.w), bf16hi_to_f32(rv.w));
            }
            if (split) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 rv = rlo[j];
                f[8 * j + 0] += bf16lo_to_f32(rv.x); f[8 * j + 1] += bf16hi_to_f32(rv.x);
                f[8 * j + 2] += bf16lo_to_f32(rv.y); f[8 * j + 3] += bf16hi_to_f32(rv.y);
                f[8 * j + 4] += bf16lo_to_f32(rv.z); f[8 * j + 5] += bf16hi_to_f32(rv.z);
                f[8 * j + 6] += bf16lo_to_f32(rv.w); f[8 * j + 7] += bf16hi_to_f32(rv.w);
              }
            }
          }
          if (p.flags & VD_CONV_SLOPE) {
            const uint32_t sp = smem_u32(sSlope) + static_cast<uint32_t>((ch * 64 + half * 32) * 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 sv = lds128(sp + j * 16);
              f[4 * j + 0] = f[4 * j + 0] > 0.f ? f[4 * j + 0] : f[4 * j + 0] * __uint_as_float(sv.x);
              f[4 * j + 1] = f[4 * j + 1] > 0.f ? f[4 * j + 1] : f[4 * j + 1] * __uint_as_float(sv.y);
              f[4 * j + 2] = f[4 * j + 2] > 0.f ? f[4 * j + 2] : f[4 * j + 2] * __uint_as_float(sv.z);
              f[4 * j + 3] = f[4 * j + 3] > 0.f ? f[4 * j + 3] : f[4 * j + 3] * __uint_as_float(sv.w);
            }
          } else if ((p.flags & VD_CONV_RELU) && !relu_in_pack) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (up_early) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 rv = upv[j];
              f[8 * j + 0] += bf16lo_to_f32(rv.x); f[8 * j + 1] += bf16hi_to_f32(rv.x);
              f[8 * j + 2] += bf16lo_to_f32(rv.y); f[8 * j + 3] += bf16hi_to_f32(rv.y);
              f[8 * j + 4] += bf16lo_to_f32(rv.z); f[8 * j + 5] += bf16hi_to_f32(rv.z);
              f[8 * j + 6] += bf16lo_to_f32(rv.w); f[8 * j + 7] += bf16hi_to_f32(rv.w);
            }
          } else if ((p.flags & VD_CONV_UPADD) && valid) {
            for (int pl = 0; pl < (split ? 2 : 1); ++pl) {
              const uint4* up = reinterpret_cast<const uint4*>(p.upadd + up_pix * up_C + pl * p.Cout + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 rv = __ldg(up + j);
                f[8 * j + 0] += bf16lo_to_f32(rv.x); f[8 * j + 1] += bf16hi_to_f32(rv.x);
                f[8 * j + 2] += bf16lo_to_f32(rv.y); f[8 * j + 3] += bf16hi_to_f32(rv.y);
                f[8 * j + 4] += bf16lo_to_f32(rv.z); f[8 * j + 5] += bf16hi_to_f32(rv.z);
                f[8 * j + 6] += bf16lo_to_f32(rv.w); f[8 * j + 7] += bf16hi_to_f32(rv.w);
              }
            }
          }
        }
        if (split && !f32out) {
          // split-precision output: hi = round(v), lo = round(v - hi); planes [hi | lo | hi] of dst_C / 3 channels.
          // Both staging tiles are used per chunk (hi, lo), so the previous chunk's stores must have read them.
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(1, kEpiThreads);
          if (active) {
            const uint32_t hrow = smem_u32(sOut) + row * 128;
            const uint32_t lrow = hrow + 128 * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 h, l;
              h.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              h.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              h.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              h.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              l.x = pack_bf16x2(f[8 * j + 0] - bf16lo_to_f32(h.x), f[8 * j + 1] - bf16hi_to_f32(h.x));
              l.y = pack_bf16x2(f[8 * j + 2] - bf16lo_to_f32(h.y), f[8 * j + 3] - bf16hi_to_f32(h.y));
              l.z = pack_bf16x2(f[8 * j + 4] - bf16lo_to_f32(h.z), f[8 * j + 5] - bf16hi_to_f32(h.z));
              l.w = pack_bf16x2(f[8 * j + 6] - bf16lo_to_f32(h.w), f[8 * j + 7] - bf16hi_to_f32(h.w));
              const int off = (((half * 4 + j) ^ (row & 7)) << 4);
              sts128(hrow + off, h);
              sts128(lrow + off, l);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (issuer) {
            const int plane = p.dst_C / 3;
            const int cch = p.dst_coff + n_tile * BN + ch * 64;      // (TWO_DST is not available in split mode)
            tma_store_4d(&tmY, sOut, cch, tx * p.TW, ty * p.TH, tb * p.NB);
            tma_store_4d(&tmY, sOut + 128 * 128, plane + cch, tx * p.TW, ty * p.TH, tb * p.NB);
            tma_store_4d(&tmY, sOut, 2 * plane + cch, tx * p.TW, ty * p.TH, tb * p.NB);
            tma_store_commit();
          }
          continue;
        }
        // staging tile `sbuf` is free once the TMA store issued two chunks ago has read it.  ONE_BAR (every variant but
        // CHAIN): the issuer checks that BEFORE the barrier that publishes the previous chunk's tile — it waits for the
        // store issued one chunk ago, i.e. the one reading the tile the NEXT chunk overwrites — so a single CTA-wide
        // barrier per chunk both publishes this tile and frees the next one.
#ifdef VD_TWO_BAR
        constexpr bool ONE_BAR = false;      // A/B build: the two-barrier protocol everywhere
#else
        constexpr bool ONE_BAR = !CHAIN;
#endif
        uint8_t* sbuf = sOut + (chunk_ctr & 1u) * (128 * 128);
        if (!ONE_BAR) {
          if (issuer) tma_store_wait_read<1>();
          if (CHAIN) mbar_wait(&yfree[chunk_ctr & 1u], ((chunk_ctr >> 1) & 1u) ^ 1u);   // chained MMAs have read it too
          named_bar_sync(1, kEpiThreads);
        }
        if (active) {
          // SWIZZLE_128B staging: 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)
          const uint32_t srow = smem_u32(sbuf) + row * 128;
          if (f32out) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(srow + ((j ^ (row & 7)) << 4),
                     make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                                __float_as_uint(f[4 * j + 3])));
          } else if (relu_in_pack) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2_relu(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16x2_relu(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16x2_relu(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16x2_relu(f[8 * j + 6], f[8 * j + 7]);
              sts128(srow + (((half * 4 + j) ^ (row & 7)) << 4), o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              sts128(srow + (((half * 4 + j) ^ (row & 7)) << 4), o);
            }
          }
        }
        fence_proxy_async_smem();              // generic-proxy smem writes -> visible to the TMA engine
        if (ONE_BAR && issuer) tma_store_wait_read<0>();      // the other staging tile's store has read it
        named_bar_sync(1, kEpiThreads);
        if (issuer) {
          const int oc = n_tile * BN + ch * 64;                  // first output channel of this chunk
          if (!CHAIN && p.cout1 > 0 && oc >= p.cout1) {          // VD_CONV_TWO_DST: the sibling layer's tensor
            tma_store_4d(&tmY2, sbuf, p.dst_b_coff + oc - p.cout1, tx * p.TW, ty * p.TH, tb * p.NB);
          } else {
            const int cch = f32out ? 0 : (p.dst_coff + oc);
            tma_store_4d(&tmY, sbuf, cch, tx * p.TW, ty * p.TH, tb * p.NB);
          }
          tma_store_commit();
          if (CHAIN) mbar_arrive(&ystaged[chunk_ctr & 1u]);    // the staged tile is also a K block of the chained GEMM
        }
        if (CHAIN && pending && ch == 0) {       // by now the previous M tile's last chained MMAs have long completed
          chain_epilogue(pend_tx, pend_ty, pend_tb);
          pending = false;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      if (CHAIN && (it & 1)) {           // chained epilogue of this M tile: owed, paid after the next tile's first chunk
        pend_tx = tx; pend_ty = ty; pend_tb = tb;
        pending = true;
      }
    }
    if (CHAIN && pending) chain_epilogue(pend_tx, pend_ty, pend_tb);
    if (issuer) tma_store_wait_all<0>();       // all output bytes written before the CTA retires
  }

  // teardown: everyone done with TMEM before the allocating warp frees it
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

}  // namespace

void* vd_tensor_map_encoder() { return reinterpret_cast<void*>(get_encode_fn()); }

struct vd_conv_plan {
  vd_stem_pool_plan* fused = nullptr;      // VD_CONV_POOL: the plan lives in stem_pool.cu
  vd_conv_rows_plan* rows = nullptr;       // 3x3 64 -> 64: the plan lives in conv3x3_rows.cu
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmY;
  CUtensorMap tmR;
  CUtensorMap tmA2;
  CUtensorMap tmB2;        // VD_CONV_CHAIN: weights of the chained convolution
  CUtensorMap tmY2;        // VD_CONV_CHAIN: its output
  ConvKParams kp;
  int BN;
  int grid;
  int smem_bytes;
  vd_conv_desc_t desc;
};

// Chooses the shared-memory layout of a plan.  occ = CTAs meant to co-reside on one SM: narrow layers are bound
// by the latency of the single issuing warp, so two half-size CTAs per SM (two producers, two MMA issuers,
// sixteen epilogue warps; TMEM 2 x 2*BN columns <= 512) beat one big one.  Returns false if the budget of
// `occ` CTAs per SM leaves fewer than `min_stages` pipeline stages.
template <int BN, bool RES, int ROWB = 128>
static bool plan_smem(vd_conv_plan* pl, int occ, int min_stages) {
  using Cfg = ConvCfg<BN, RES, ROWB>;
  ConvKParams& k = pl->kp;
  const int KB = k.R * k.S * k.cin_blocks + k.kb2;
  const int budget = occ > 1 ? (kSmemBudget / occ - 1024) : kSmemBudget;
  // per-channel slope vector behind the barriers; CHAIN: resident W2 + the chained output's staging tile
  const int extra =
      ((k.flags & VD_CONV_SLOPE) ? 1024 : 0) + ((k.flags & VD_CONV_CHAIN) ? k.w2_bytes + 128 * 128 + 512 : 0);
  const int avail = budget - Cfg::FIXED_BYTES - extra;
  const char* env = getenv("VD_CONV_KPS");
  const int force_kps = env ? atoi(env) : 0;
  if (k.patch) {
    k.a_slot_bytes = (k.PW * k.PH * 128 + 1023) & ~1023;
    {
      // narrow 3x3 layers whose whole filter bank fits (64 -> 64: 72 KB) keep it resident: streaming it costs
      // 72 KB of L2 reads per 128-pixel tile, three times the patch itself
      const char* we = getenv("VD_CONV_PATCH_WRES");
      const int want = we ? atoi(we) : 1;
      const int wb = KB * Cfg::B_BYTES;
      if (want && occ == 1 && k.n_tiles == 1 && wb <= kWresMaxBytes && k.num_tiles > pl->grid) {
        int slots = (avail - wb) / k.a_slot_bytes;
        if (slots > 4) slots = 4;
        if (slots >= 2) {
          k.wres = 1;
          k.w_bytes = wb;
          k.a_slots = slots;
          k.kps = 1;
          k.stages = 1;
          pl->smem_bytes = extra + Cfg::FIXED_BYTES + k.a_slots * k.a_slot_bytes + wb;
          return true;
        }
      }
    }
    k.wres = 0;
    k.w_bytes = 0;
    k.a_slots = 2;
    k.kps = (BN <= 128) ? 3 : 1;                         // taps per stage (must divide 9)
    if (force_kps == 1 || force_kps == 3) k.kps = force_kps;
    int st = (avail - k.a_slots * k.a_slot_bytes) / (k.kps * Cfg::B_BYTES);
    if (st < min_stages && k.kps == 3 && !force_kps) {   // tight budget (two CTAs per SM): one tap per stage
      k.kps = 1;
      st = (avail - k.a_slots * k.a_slot_bytes) / Cfg::B_BYTES;
    }
    if (st > 3 && k.cin_blocks >= 2 && (avail - 3 * k.a_slot_bytes) / (k.kps * Cfg::B_BYTES) >= 3) {
      k.a_slots = 3;
      st = (avail - k.a_slots * k.a_slot_bytes) / (k.kps * Cfg::B_BYTES);
    }
    k.stages = st > kMaxStages ? kMaxStages : st;
    pl->smem_bytes = extra + Cfg::FIXED_BYTES + k.a_slots * k.a_slot_bytes + k.stages * k.kps * Cfg::B_BYTES;
    return k.stages >= min_stages;
  }
  if (ROWB == 64) {
    // stem: weights resident, one pipeline stage = the raw rows of one tile
    k.wres = 1;
    k.w_bytes = KB * Cfg::B_BYTES;
    k.kps = 1;
    const int st = (avail - k.w_bytes) / Cfg::A_BYTES;
    k.stages = st > kMaxStages ? kMaxStages : st;
    pl->smem_bytes = extra + Cfg::FIXED_BYTES + k.w_bytes + k.stages * Cfg::A_BYTES;
    return k.stages >= min_stages;
  }
  const int w_bytes = KB * Cfg::B_BYTES;
  k.wres = (k.n_tiles == 1 && w_bytes <= kWresMaxBytes && w_bytes * 2 <= avail && k.num_tiles > pl->grid &&
            BN <= 64 && KB >= 4) ? 1 : 0;
  // Short-K expansions with several 256-wide N tiles (bottleneck conv3 of layer2: K = 128, 64 KB per N tile): streamed,
  // the weights are 64 KB of L2 -> SM traffic per tile next to 32 KB of activations.  With a grid that is a multiple of
  // n_tiles, tile = blockIdx.x + it * gridDim.x keeps n_tile = blockIdx.x % n_tiles fixed per CTA, so its slice of the
  // weights can stay resident.
  {
    const char* we = getenv("VD_CONV_WRES_NT");
    const int want = we ? atoi(we) : 1;
    if (want && !k.wres && BN == 256 && k.n_tiles > 1 && pl->grid % k.n_tiles == 0 && !(k.flags & VD_CONV_CHAIN) &&
        w_bytes <= 64 * 1024 && (avail - w_bytes) / Cfg::A_BYTES >= 4 && k.num_tiles > pl->grid)
      k.wres = 1;
  }
  k.w_bytes = k.wres ? w_bytes : 0;
  // K blocks per stage: narrow N tiles finish a 64-deep block in 128 (BN=64) / 256 (BN=128) clk, far less
  // than one barrier round trip of the issuing warp
  int kps = 1;
  if (BN <= 64) kps = k.wres ? 4 : 2;
  else if (BN == 128) kps = 2;
  if (kps > KB) kps = KB;
  if (force_kps > 0 && force_kps <= 8) kps = force_kps > KB ? KB : force_kps;
  const int per_block = Cfg::A_BYTES + (k.wres ? 0 : Cfg::B_BYTES);
  int st = (avail - k.w_bytes) / (kps * per_block);
  while (st < min_stages && kps > 1) {
    --kps;
    st = (avail - k.w_bytes) / (kps * per_block);
  }
  k.kps = kps;
  k.stages = st > kMaxStages ? kMaxStages : st;
  pl->smem_bytes = extra + Cfg::FIXED_BYTES + k.w_bytes + k.stages * kps * per_block;
  return k.stages >= min_stages;
}

// occupancy policy: variants compiled with two CTAs per SM (see __launch_bounds__) try occ = 2 first
template <int BN, bool RES, int ROWB = 128>
static bool plan_occ(vd_conv_plan* pl, int sms) {
  const bool kTwo = (BN <= 128 && !RES) && !(pl->kp.flags & VD_CONV_CHAIN);
  const char* env = getenv("VD_CONV_OCC");
  const int want = env ? atoi(env) : 2;
  if (pl->kp.flags & VD_CONV_CHAIN) {        // a CTA walks whole M tiles (tile_at<true>): one CTA per M tile at most
    const int m_tiles = pl->kp.num_tiles / pl->kp.n_tiles;
    pl->grid = m_tiles < sms ? m_tiles : sms;
    return plan_smem<BN, RES, ROWB>(pl, 1, 2);
  }
  if (pl->kp.patch) {                       // narrow 3x3: one CTA per SM with the filter bank resident, if it fits
    pl->grid = pl->kp.num_tiles < sms ? pl->kp.num_tiles : sms;
    if (plan_smem<BN, RES, ROWB>(pl, 1, 2) && pl->kp.wres) return true;
  }
  if (kTwo && want >= 2 && pl->kp.num_tiles > sms) {
    pl->grid = pl->kp.num_tiles < 2 * sms ? pl->kp.num_tiles : 2 * sms;
    if (plan_smem<BN, RES, ROWB>(pl, 2, 3)) return true;
  }
  pl->grid = pl->kp.num_tiles < sms ? pl->kp.num_tiles : sms;
  return plan_smem<BN, RES, ROWB>(pl, 1, 2);
}

static void choose_tile(int Wout, int Hout, int B, int sw, int sh, int* TW, int* TH, int* NB) {
  long best_tiles = -1;
  int bw = 1, bh = 1, bn = 1;
  for (int tw = 1; tw <= Wout && tw <= 128; ++tw) {
    if (tw * sw > 256) break;
    for (int th = 1; th <= Hout && tw * th <= 128; ++th) {
      if (th * sh > 256) break;
      int nb = 128 / (tw * th);
      if (nb > B) nb = B;
      if (nb < 1) nb = 1;
      long tiles = static_cast<long>(vd_cdiv(Wout, tw)) * vd_cdiv(Hout, th) * vd_cdiv(B, nb);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && tw > bw)) {
        best_tiles = tiles;
        bw = tw; bh = th; bn = nb;
      }
    }
  }
  *TW = bw; *TH = bh; *NB = bn;
}

extern "C" int vd_conv_plan_create(const vd_conv_desc_t* d, vd_conv_plan_t** out) {
  VD_REQUIRE(d && out, "vd_conv_plan_create: null argument");
  const bool x0 = (d->flags & VD_CONV_IN_X0) != 0;
  VD_REQUIRE(d->B > 0 && d->Hout > 0 && d->Wout > 0, "vd_conv_plan_create: empty extent");
  if (d->flags & VD_CONV_POOL) {
    VD_REQUIRE(d->x && d->w && d->bias && d->y, "vd_conv_plan_create: null tensor pointer");
    VD_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->y) & 15) == 0,
               "vd_conv_plan_create: tensors must be 16-byte aligned");
    vd_conv_plan* pl = new (std::nothrow) vd_conv_plan();
    VD_REQUIRE(pl, "out of host memory");
    pl->desc = *d;
    memset(&pl->kp, 0, sizeof(pl->kp));
    pl->kp.flags = d->flags;
    pl->BN = 64;
    const int rc = vd_stem_pool_create(d, &pl->fused);
    if (rc != VD_OK) {
      delete pl;
      return rc;
    }
    vd_stem_pool_info(pl->fused, &pl->kp.num_tiles, &pl->grid);
    pl->kp.TW = 120; pl->kp.TH = 1; pl->kp.NB = 1;
    *out = pl;
    return VD_OK;
  }
  VD_REQUIRE(d->Cout % 32 == 0 && d->Cout >= 32, "vd_conv_plan_create: Cout=%d must be a multiple of 32", d->Cout);
  VD_REQUIRE(x0 || (d->Cin % 64 == 0), "vd_conv_plan_create: Cin=%d must be a multiple of 64", d->Cin);
  VD_REQUIRE(d->x && d->w && d->bias && d->y, "vd_conv_plan_create: null tensor pointer");
  VD_REQUIRE(!(d->flags & VD_CONV_RESIDUAL) || d->residual, "vd_conv_plan_create: residual flag without tensor");
  VD_REQUIRE(!(d->flags & VD_CONV_SLOPE) || (d->slope && !(d->flags & VD_CONV_SPLIT)),
             "vd_conv_plan_create: VD_CONV_SLOPE needs desc.slope (and is not available in split mode)");
  VD_REQUIRE(!(d->flags & VD_CONV_UPADD) || (d->upadd && d->up_H > 0 && d->up_W > 0),
             "vd_conv_plan_create: upadd flag without tensor");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->y) & 15) == 0,
             "vd_conv_plan_create: tensors must be 16-byte aligned");
  const bool split_mode = (d->flags & VD_CONV_SPLIT) != 0;
  VD_REQUIRE(d->dst_coff % 32 == 0 && d->dst_C % 8 == 0 &&
                 d->dst_coff + ((d->flags & VD_CONV_TWO_DST) ? d->Cout1 : d->Cout) <=
                     ((split_mode && !(d->flags & VD_CONV_OUT_F32)) ? d->dst_C / 3 : d->dst_C),
             "vd_conv_plan_create: bad destination slice");
  VD_REQUIRE(!split_mode || !(d->flags & VD_CONV_DUAL), "vd_conv_plan_create: SPLIT and DUAL cannot be combined");
  VD_REQUIRE(!split_mode || (d->flags & VD_CONV_OUT_F32) || d->dst_C % 3 == 0,
             "vd_conv_plan_create: split destination must have 3 planes");
  if (d->flags & VD_CONV_UPADD) {
    // the FPN's F.interpolate(mode="nearest") source index floor(dst*in/out) must equal dst>>1
    // (reference layers.py:102,110; true for every ceil(n/2) pyramid) — checked, not assumed
    for (int yy = 0; yy < d->Hout; ++yy) {
      int ref = static_cast<int>(floorf(yy * (static_cast<float>(d->up_H) / d->Hout)));
      if (ref > d->up_H - 1) ref = d->up_H - 1;
      int got = yy >> 1; if (got > d->up_H - 1) got = d->up_H - 1;
      VD_REQUIRE(ref == got, "vd_conv_plan_create: nearest-upsample row map is not y>>1 (%d<-%d)", d->Hout, d->up_H);
    }
    for (int xx = 0; xx < d->Wout; ++xx) {
      int ref = static_cast<int>(floorf(xx * (static_cast<float>(d->up_W) / d->Wout)));
      if (ref > d->up_W - 1) ref = d->up_W - 1;
      int got = xx >> 1; if (got > d->up_W - 1) got = d->up_W - 1;
      VD_REQUIRE(ref == got, "vd_conv_plan_create: nearest-upsample col map is not x>>1 (%d<-%d)", d->Wout, d->up_W);
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    vd_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return VD_ERR_NO_DEVICE;
  }

  vd_conv_plan* pl = new (std::nothrow) vd_conv_plan();
  VD_REQUIRE(pl, "out of host memory");
  pl->desc = *d;
  if (vd_conv_rows_eligible(d)) {           // 3x3 64 -> 64: two output rows per accumulator (conv3x3_rows.cu)
    memset(&pl->kp, 0, sizeof(pl->kp));
    pl->kp.flags = d->flags;
    pl->BN = 128;
    const int rc = vd_conv_rows_create(d, &pl->rows);
    if (rc != VD_OK) {
      delete pl;
      return rc;
    }
    vd_conv_rows_info(pl->rows, &pl->kp.num_tiles, &pl->grid);
    pl->kp.TW = 128; pl->kp.TH = 2; pl->kp.NB = 1;
    *out = pl;
    return VD_OK;
  }
  ConvKParams& k = pl->kp;
  memset(&k, 0, sizeof(k));
  int BN = d->Cout >= 256 ? 256 : d->Cout;
  const bool chain = (d->flags & VD_CONV_CHAIN) != 0;
  if (chain) {
    if (x0 || split_mode || d->R != 1 || d->S != 1 || d->Cout != 256 || (d->chain_Cout != 64 && d->chain_Cout != 128) ||
        !d->chain_w || !d->chain_bias || !d->chain_y || (d->flags & (VD_CONV_OUT_F32 | VD_CONV_UPADD | VD_CONV_SLOPE)) ||
        (reinterpret_cast<uintptr_t>(d->chain_w) & 15) != 0 || (reinterpret_cast<uintptr_t>(d->chain_y) & 15) != 0) {
      delete pl;
      vd_set_error("vd_conv_plan_create: VD_CONV_CHAIN needs a 1x1 convolution with Cout = 256, chain_Cout 64/128 and "
                   "16-byte aligned chain tensors");
      return VD_ERR_UNSUPPORTED;
    }
    BN = 128;                    // two N tiles per M tile: 2 x 128 + 2 x chain_Cout TMEM columns
  }
  const bool two_dst = (d->flags & VD_CONV_TWO_DST) != 0;
  if (two_dst) {
    // two sibling layers that read the same tensor, as ONE GEMM over the concatenated filters: Cout = Cout1 + Cout2
    // in one N tile (128 or 192), so the A operand is fetched and streamed through the MMA once for both
    if (chain || x0 || split_mode || (d->flags & (VD_CONV_OUT_F32 | VD_CONV_RESIDUAL | VD_CONV_UPADD | VD_CONV_DUAL | VD_CONV_POOL)) ||
        (d->Cout != 128 && d->Cout != 192) || d->Cout1 <= 0 || d->Cout1 >= d->Cout || d->Cout1 % 64 != 0 || !d->y_b ||
        (reinterpret_cast<uintptr_t>(d->y_b) & 15) != 0 || d->dst_b_coff % 64 != 0 || d->dst_b_C % 8 != 0 ||
        d->dst_b_coff + (d->Cout - d->Cout1) > d->dst_b_C) {
      delete pl;
      vd_set_error("vd_conv_plan_create: VD_CONV_TWO_DST needs Cout in {128, 192}, Cout1 a multiple of 64 inside it, "
                   "a 16-byte aligned y_b and a 64-aligned slice of it; no residual / upadd / dual / chain / fp32 out");
      return VD_ERR_UNSUPPORTED;
    }
  }
  if (BN != 32 && BN != 64 && BN != 128 && BN != 256 && !(BN == 192 && two_dst)) {
    delete pl;
    vd_set_error("vd_conv_plan_create: unsupported Cout=%d", d->Cout);
    return VD_ERR_UNSUPPORTED;
  }
  pl->BN = BN;

  int sw, sh, pw, ph, R, S, cin_blocks;
  if (x0) {
    // conv1 7x7/2 p3 on the ingest layout: each filter row is one K block of 8 taps x 4 channels (64 bytes);
    // the x stride/pad are baked into the sliding-window tensor map.
    R = 7; S = 1; cin_blocks = split_mode ? 2 : 1; sh = 2; ph = 3; sw = 1; pw = 0;   // split: [w_hi | w_lo] per row
    VD_REQUIRE(d->R == 7 && d->S == 7 && d->stride == 2 && d->pad == 3, "IN_X0 conv must be 7x7 s2 p3");
  } else {
    R = d->R; S = d->S; cin_blocks = d->Cin / 64; sh = sw = d->stride; ph = pw = d->pad;
  }
  choose_tile(d->Wout, d->Hout, d->B, sw, sh, &k.TW, &k.TH, &k.NB);
  if (x0) {
    // stem tile = one run of TW output pixels of one row; TW is a multiple of 8 so that the run's first input
    // pixel (2*X0) starts a 128-byte chunk; fewest tiles first, then the narrowest
    int best = -1;
    for (int tw = 8; tw <= 128; tw += 8)
      if (best < 0 || vd_cdiv(d->Wout, tw) < vd_cdiv(d->Wout, best)) best = tw;
    k.TW = best; k.TH = 1; k.NB = 1;
  }
  {
    // halo-patch mode: 3x3 stride-1 convs whose N tile is narrow are bound by the 9x re-fetch of the A tile
    // from L2 (~25 B/clk/SM of unique data); an 8x16 pixel tile + halo is fetched once per 64-channel slice.
    const char* env = getenv("VD_CONV_PATCH");
    const int mode = env ? atoi(env) : 1;              // 0 = off, 1 = auto, 2 = force where legal
    const bool legal = !x0 && d->R == 3 && d->S == 3 && d->stride == 1 && d->pad == 1;
    const long generic_tiles = static_cast<long>(vd_cdiv(d->Wout, k.TW)) * vd_cdiv(d->Hout, k.TH) * vd_cdiv(d->B, k.NB);
    const long patch_tiles = static_cast<long>(vd_cdiv(d->Wout, 8)) * vd_cdiv(d->Hout, 16) * d->B;
    // (BN = 192 exists only for two sibling layers (TWO_DST), each of which runs in patch mode on its own: the fused
    //  GEMM keeps their K order — cin block outer, taps inner — and so stays bit-identical to the two launches)
    const bool worth = BN <= 192 && patch_tiles * 4 <= generic_tiles * 5;      // <= 25 % more tiles
    if (legal && (mode == 2 || (mode == 1 && worth))) {
      k.patch = 1;
      k.TW = 8; k.TH = 16; k.NB = 1;
      k.PW = 10; k.PH = 18;
    }
  }
  k.tiles_x = vd_cdiv(d->Wout, k.TW);
  k.tiles_y = vd_cdiv(d->Hout, k.TH);
  k.tiles_b = vd_cdiv(d->B, k.NB);
  k.n_tiles = d->Cout / BN;
  k.num_tiles = k.tiles_x * k.tiles_y * k.tiles_b * k.n_tiles;
  {
    auto magic = [](int d, uint32_t* mul, uint32_t* shr) {
      if (d <= 1) { *mul = 0; *shr = 0; return; }
      int lg = 0;
      while ((1u << lg) < static_cast<uint32_t>(d)) ++lg;
      const int pw = 31 + lg;
      *mul = static_cast<uint32_t>(((1ull << pw) + static_cast<unsigned long long>(d) - 1) / d);
      *shr = static_cast<uint32_t>(pw - 32);
    };
    magic(k.n_tiles, &k.dv_n_mul, &k.dv_n_shr);
    magic(k.tiles_x, &k.dv_x_mul, &k.dv_x_shr);
    magic(k.tiles_y, &k.dv_y_mul, &k.dv_y_shr);
  }
  k.rows_used = k.TW * k.TH * k.NB;
  const int x0_chunks = (2 * k.TW + 6 + 15) / 16;          // 128-byte chunks of one stem row segment
  k.a_bytes = x0 ? x0_chunks * 128 : k.rows_used * 128;
  k.B = d->B; k.Hout = d->Hout; k.Wout = d->Wout; k.Cout = d->Cout;
  k.R = R; k.S = S; k.cin_blocks = cin_blocks;
  k.kb1 = R * S * cin_blocks;
  k.kb2 = 0;
  k.stride2 = 1;
  const bool dual = (d->flags & VD_CONV_DUAL) != 0;
  if (dual) {
    if (x0 || !d->x2 || d->Cin2 <= 0 || d->Cin2 % 64 != 0 || d->stride2 < 1 || d->R != 1 || d->S != 1 ||
        (reinterpret_cast<uintptr_t>(d->x2) & 15) != 0 || (d->Hin2 + d->stride2 - 1) / d->stride2 != d->Hout ||
        (d->Win2 + d->stride2 - 1) / d->stride2 != d->Wout) {
      vd_set_error("vd_conv_plan_create: bad second input for VD_CONV_DUAL");
      delete pl;
      return VD_ERR_INVALID;
    }
    k.kb2 = d->Cin2 / 64;
    k.stride2 = d->stride2;
  }
  k.stride_h = sh; k.stride_w = sw; k.pad_h = ph; k.pad_w = pw;
  k.dst_C = d->dst_C; k.dst_coff = d->dst_coff;
  k.up_H = d->up_H; k.up_W = d->up_W;
  k.flags = d->flags;
  k.reverse = ((d->flags & VD_CONV_REVERSE) && !chain) ? 1 : 0;
  {
    const char* re = getenv("VD_CONV_RELU_PACK");
    k.relu_pack = re ? (atoi(re) != 0) : 1;
  }
  k.bias = d->bias;
  k.slope = d->slope;
  k.bias2 = d->chain_bias;
  k.N2 = chain ? d->chain_Cout : 0;
  k.w2_bytes = chain ? d->chain_Cout * d->Cout * 2 : 0;
  k.residual = static_cast<const __nv_bfloat16*>(d->residual);
  k.upadd = static_cast<const __nv_bfloat16*>(d->upadd);
  k.y = d->y;
  k.acc_scale = (split_mode && d->acc_scale > 0.f) ? d->acc_scale : 1.0f;
  k.cout1 = two_dst ? d->Cout1 : 0;
  k.dst_b_coff = two_dst ? d->dst_b_coff : 0;

  // ---- A: activations ----
  cuuint64_t gdim[4], gstr[3];
  cuuint32_t box[4], estr[4];
  if (x0) {
    const int Wp = vd_x0_pitch(d->Win);                   // multiple of 16 pixels
    gdim[0] = 64; gdim[1] = static_cast<cuuint64_t>(Wp / 16); gdim[2] = static_cast<cuuint64_t>(d->Hin);
    gdim[3] = static_cast<cuuint64_t>(d->B);
    gstr[0] = 128;                                        // one chunk = 16 pixels * 8 B
    gstr[1] = static_cast<cuuint64_t>(Wp) * 8;
    gstr[2] = static_cast<cuuint64_t>(Wp) * 8 * d->Hin;
  } else {
    gdim[0] = static_cast<cuuint64_t>(d->Cin); gdim[1] = static_cast<cuuint64_t>(d->Win);
    gdim[2] = static_cast<cuuint64_t>(d->Hin); gdim[3] = static_cast<cuuint64_t>(d->B);
    gstr[0] = static_cast<cuuint64_t>(d->Cin) * 2;
    gstr[1] = gstr[0] * d->Win;
    gstr[2] = gstr[1] * d->Hin;
  }
  box[0] = 64; box[1] = static_cast<cuuint32_t>(k.TW * sw); box[2] = static_cast<cuuint32_t>(k.TH * sh);
  box[3] = static_cast<cuuint32_t>(k.NB);
  if (k.patch) { box[1] = static_cast<cuuint32_t>(k.PW); box[2] = static_cast<cuuint32_t>(k.PH); }
  estr[0] = 1; estr[1] = static_cast<cuuint32_t>(sw); estr[2] = static_cast<cuuint32_t>(sh); estr[3] = 1;
  if (x0) { box[1] = static_cast<cuuint32_t>(x0_chunks); box[2] = 7; estr[1] = estr[2] = 1; }
  CUresult r = enc(&pl->tmA, kActTmapType, 4, const_cast<void*>(d->x), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, x0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(A) failed: CUresult %d (dims %llu %llu %llu %llu, box %u %u %u %u)", (int)r,
                 (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                 (unsigned long long)gdim[3], box[0], box[1], box[2], box[3]);
    delete pl;
    return VD_ERR_CUDA;
  }
  pl->tmA2 = pl->tmA;
  if (dual) {
    cuuint64_t g2[4] = {static_cast<cuuint64_t>(d->Cin2), static_cast<cuuint64_t>(d->Win2),
                        static_cast<cuuint64_t>(d->Hin2), static_cast<cuuint64_t>(d->B)};
    cuuint64_t s2[3] = {static_cast<cuuint64_t>(d->Cin2) * 2, static_cast<cuuint64_t>(d->Cin2) * 2 * d->Win2,
                        static_cast<cuuint64_t>(d->Cin2) * 2 * d->Win2 * d->Hin2};
    cuuint32_t b2[4] = {64, static_cast<cuuint32_t>(k.TW * d->stride2), static_cast<cuuint32_t>(k.TH * d->stride2),
                        static_cast<cuuint32_t>(k.NB)};
    cuuint32_t e2[4] = {1, static_cast<cuuint32_t>(d->stride2), static_cast<cuuint32_t>(d->stride2), 1};
    r = enc(&pl->tmA2, kActTmapType, 4, const_cast<void*>(d->x2), g2, s2, b2, e2,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vd_set_error("cuTensorMapEncodeTiled(A2) failed: CUresult %d", (int)r);
      delete pl;
      return VD_ERR_CUDA;
    }
  }
  // ---- B: weights [Cout][Ktot] ----
  const cuuint64_t kelems = x0 ? 32 : 64;               // elements per K block
  const cuuint64_t Ktot = (static_cast<cuuint64_t>(R) * S * cin_blocks + static_cast<cuuint64_t>(k.kb2)) * kelems;
  cuuint64_t wdim[2] = {Ktot, static_cast<cuuint64_t>(d->Cout)};
  cuuint64_t wstr[1] = {Ktot * 2};
  cuuint32_t wbox[2] = {static_cast<cuuint32_t>(kelems), static_cast<cuuint32_t>(BN)};
  cuuint32_t wes[2] = {1, 1};
  r = enc(&pl->tmB, kActTmapType, 2, const_cast<void*>(d->w), wdim, wstr, wbox, wes,
          CU_TENSOR_MAP_INTERLEAVE_NONE, x0 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vd_set_error("cuTensorMapEncodeTiled(B) failed: CUresult %d", (int)r);
    delete pl;
    return VD_ERR_CUDA;
  }
  // ---- Y: output tile store (box = one 128-byte line per pixel of the M tile) ----
  {
    const bool f32 = (d->flags & VD_CONV_OUT_F32) != 0;
    const int esz = f32 ? 4 : 2;
    if (f32 && (d->dst_C != 32 || d->dst_coff != 0 || BN != 32)) {
      vd_set_error("vd_conv_plan_create: fp32 output is implemented for the 32-channel head GEMM only");
      delete pl;
      return VD_ERR_UNSUPPORTED;
    }
    if (!f32 && (BN < 64 || d->dst_coff % 64 != 0)) {
      vd_set_error("vd_conv_plan_create: bf16 output needs Cout >= 64 and a 64-aligned channel offset");
      delete pl;
      return VD_ERR_UNSUPPORTED;
    }
    cuuint64_t ydim[4] = {static_cast<cuuint64_t>(d->dst_C), static_cast<cuuint64_t>(d->Wout),
                          static_cast<cuuint64_t>(d->Hout), static_cast<cuuint64_t>(d->B)};
    cuuint64_t ystr[3] = {static_cast<cuuint64_t>(d->dst_C) * esz, static_cast<cuuint64_t>(d->dst_C) * esz * d->Wout,
                          static_cast<cuuint64_t>(d->dst_C) * esz * d->Wout * d->Hout};
    cuuint32_t ybox[4] = {static_cast<cuuint32_t>(f32 ? 32 : 64), static_cast<cuuint32_t>(k.TW),
                          static_cast<cuuint32_t>(k.TH), static_cast<cuuint32_t>(k.NB)};
    cuuint32_t yes[4] = {1, 1, 1, 1};
    r = enc(&pl->tmY, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : kActTmapType, 4, d->y, ydim, ystr,
            ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vd_set_error("cuTensorMapEncodeTiled(Y) failed: CUresult %d", (int)r);
      delete pl;
      return VD_ERR_CUDA;
    }
  }
  // ---- R: residual tile load (same tile geometry as the output) ----
  pl->tmR = pl->tmY;
  if (d->flags & VD_CONV_RESIDUAL) {
    if ((reinterpret_cast<uintptr_t>(d->residual) & 15) != 0 || (d->flags & VD_CONV_OUT_F32)) {
      vd_set_error("vd_conv_plan_create: residual must be 16-byte aligned bf16");
      delete pl;
      return VD_ERR_INVALID;
    }
    const cuuint64_t rC = static_cast<cuuint64_t>(d->Cout) * (split_mode ? 3 : 1);
    cuuint64_t rdim[4] = {rC, static_cast<cuuint64_t>(d->Wout), static_cast<cuuint64_t>(d->Hout),
                          static_cast<cuuint64_t>(d->B)};
    cuuint64_t rstr[3] = {rC * 2, rC * 2 * d->Wout, rC * 2 * d->Wout * d->Hout};
    cuuint32_t rbox[4] = {64, static_cast<cuuint32_t>(k.TW), static_cast<cuuint32_t>(k.TH),
                          static_cast<cuuint32_t>(k.NB)};
    cuuint32_t res1[4] = {1, 1, 1, 1};
    r = enc(&pl->tmR, kActTmapType, 4, const_cast<void*>(d->residual), rdim, rstr, rbox, res1,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vd_set_error("cuTensorMapEncodeTiled(R) failed: CUresult %d", (int)r);
      delete pl;
      return VD_ERR_CUDA;
    }
  }
  pl->tmB2 = pl->tmB;
  pl->tmY2 = pl->tmY;
  if (two_dst) {
    const cuuint64_t c2 = static_cast<cuuint64_t>(d->dst_b_C);
    cuuint64_t ydim2[4] = {c2, static_cast<cuuint64_t>(d->Wout), static_cast<cuuint64_t>(d->Hout),
                           static_cast<cuuint64_t>(d->B)};
    cuuint64_t ystr2[3] = {c2 * 2, c2 * 2 * d->Wout, c2 * 2 * d->Wout * d->Hout};
    cuuint32_t ybox2[4] = {64, static_cast<cuuint32_t>(k.TW), static_cast<cuuint32_t>(k.TH),
                           static_cast<cuuint32_t>(k.NB)};
    cuuint32_t one4[4] = {1, 1, 1, 1};
    r = enc(&pl->tmY2, kActTmapType, 4, d->y_b, ydim2, ystr2, ybox2, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      vd_set_error("cuTensorMapEncodeTiled(second destination) failed: CUresult %d", (int)r);
      delete pl;
      return VD_ERR_CUDA;
    }
  }
  if (chain) {
    cuuint64_t wdim2[2] = {static_cast<cuuint64_t>(d->Cout), static_cast<cuuint64_t>(d->chain_Cout)};
    cuuint64_t wstr2[1] = {static_cast<cuuint64_t>(d->Cout) * 2};
    cuuint32_t wbox2[2] = {64, static_cast<cuuint32_t>(d->chain_Cout)};
    cuuint32_t one2[2] = {1, 1};
    r = enc(&pl->tmB2, kActTmapType, 2, const_cast<void*>(d->chain_w), wdim2, wstr2, wbox2, one2,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS) {
      const cuuint64_t c2 = static_cast<cuuint64_t>(d->chain_Cout);
      cuuint64_t ydim2[4] = {c2, static_cast<cuuint64_t>(d->Wout), static_cast<cuuint64_t>(d->Hout),
                             static_cast<cuuint64_t>(d->B)};
      cuuint64_t ystr2[3] = {c2 * 2, c2 * 2 * d->Wout, c2 * 2 * d->Wout * d->Hout};
      cuuint32_t ybox2[4] = {64, static_cast<cuuint32_t>(k.TW), static_cast<cuuint32_t>(k.TH),
                             static_cast<cuuint32_t>(k.NB)};
      cuuint32_t one4[4] = {1, 1, 1, 1};
      r = enc(&pl->tmY2, kActTmapType, 4, d->chain_y, ydim2, ystr2, ybox2, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      vd_set_error("cuTensorMapEncodeTiled(chain) failed: CUresult %d", (int)r);
      delete pl;
      return VD_ERR_CUDA;
    }
  }
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  pl->grid = k.num_tiles < sms ? k.num_tiles : sms;
  {
    const bool res = (d->flags & VD_CONV_RESIDUAL) != 0;
    bool ok;
    if (x0) ok = plan_occ<64, false, 64>(pl, sms);
    else switch (BN) {
      case 32: ok = plan_occ<32, false>(pl, sms); break;
      case 64: ok = res ? plan_occ<64, true>(pl, sms) : plan_occ<64, false>(pl, sms); break;
      case 128: ok = res ? plan_occ<128, true>(pl, sms) : plan_occ<128, false>(pl, sms); break;
      case 192: ok = plan_occ<192, false>(pl, sms); break;
      default: ok = res ? plan_occ<256, true>(pl, sms) : plan_occ<256, false>(pl, sms); break;
    }
    if (!ok) {
      vd_set_error("vd_conv_plan_create: shared memory budget too small");
      delete pl;
      return VD_ERR_UNSUPPORTED;
    }
  }
  *out = pl;
  return VD_OK;
}

extern "C" int vd_conv_plan_destroy(vd_conv_plan_t* plan) {
  if (plan && plan->fused) vd_stem_pool_destroy(plan->fused);
  if (plan && plan->rows) vd_conv_rows_destroy(plan->rows);
  delete plan;
  return VD_OK;
}

extern "C" int vd_conv_plan_info(const vd_conv_plan_t* plan, int* TW, int* TH, int* NB, int* BN, int* num_tiles,
                                 int* grid) {
  VD_REQUIRE(plan, "vd_conv_plan_info: null plan");
  if (TW) *TW = plan->kp.TW;
  if (TH) *TH = plan->kp.TH;
  if (NB) *NB = plan->kp.NB;
  if (BN) *BN = plan->BN;
  if (num_tiles) *num_tiles = plan->kp.num_tiles;
  if (grid) *grid = plan->grid;
  return VD_OK;
}

// Programmatic dependent launch (prologue of kernel k+1 — barrier init, TMEM allocation, descriptor prefetch — overlaps the
// tail of kernel k; no global memory is touched before griddepcontrol.wait).  Round 1, eager launches: no difference
// within run-to-run noise (1351 vs 1357 frames/s), left off.  Round 2, with the network replayed from a CUDA graph (the
// launch latency that used to hide the effect is gone): network segment 8.72 -> 8.69 and 8.77 -> 8.69 ms on one box, so it
// is ON by default; VD_CONV_PDL=0 disables it.
static const bool g_conv_pdl = [] {
  const char* e = getenv("VD_CONV_PDL");
  return !(e && atoi(e) == 0);
}();


template <int BN, bool RES, int ROWB = 128, bool CHAIN = false>
static int launch_conv(const vd_conv_plan* pl, cudaStream_t st) {
  using Cfg = ConvCfg<BN, RES, ROWB>;
  VD_SET_FUNC_ATTR_ONCE_PER_DEVICE((conv_igemm_kernel<BN, RES, ROWB, CHAIN>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSmemBudget);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pl->grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(pl->smem_bytes);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_conv_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, RES, ROWB, CHAIN>, pl->tmA, pl->tmB, pl->tmY, pl->tmR,
                                   pl->tmA2, pl->tmB2, pl->tmY2, pl->kp));
  return VD_OK;
}

extern "C" int vd_conv_run(const vd_conv_plan_t* plan, vd_stream_t stream) {
  VD_REQUIRE(plan, "vd_conv_run: null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan->fused) return vd_stem_pool_run(plan->fused, st);
  if (plan->rows) return vd_conv_rows_run(plan->rows, st);
  const bool res = (plan->kp.flags & VD_CONV_RESIDUAL) != 0;
  if (plan->kp.flags & VD_CONV_IN_X0) return launch_conv<64, false, 64>(plan, st);
  if (plan->kp.flags & VD_CONV_CHAIN)
    return res ? launch_conv<128, true, 128, true>(plan, st) : launch_conv<128, false, 128, true>(plan, st);
  switch (plan->BN) {
    case 32: return launch_conv<32, false>(plan, st);
    case 64: return res ? launch_conv<64, true>(plan, st) : launch_conv<64, false>(plan, st);
    case 128: return res ? launch_conv<128, true>(plan, st) : launch_conv<128, false>(plan, st);
    case 192: return launch_conv<192, false>(plan, st);
    case 256: return res ? launch_conv<256, true>(plan, st) : launch_conv<256, false>(plan, st);
  }
  vd_set_error("vd_conv_run: bad BN");
  return VD_ERR_UNSUPPORTED;
}
