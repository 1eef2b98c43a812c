// mnet.cu — the sub-64-channel layers of the MobileNet0.25 variant (Retinaface(backbone="mobilenet")).
//
// Reference: detect_face/nets/mobilenet025.py:3-19 — conv_bn(3, 8, stride 2) and the conv_dw blocks
// (depthwise 3x3 + BN + LeakyReLU(0.1), pointwise 1x1 + BN + LeakyReLU(0.1)).  These layers move 8..32-channel
// tensors at 1/2 and 1/4 resolution: HBM-bound by their bytes, but only if the arithmetic stays out of the way —
//   * depthwise: FHFMA (fma.rn.f32.bf16 / .f16: activation-type operands with .H0/.H1 selectors, fp32 accumulate; measured
//     at the full FFMA issue rate on B200, tools/probes/fma_rate.cu) — no conversion instructions at all;
//   * the dense parts (stem 3 -> 8, pointwise Cin <= 32) on warp-level mma.sync: they ARE small GEMMs (K = 8..48,
//     N = 8..64), 27 resp. Cin FMAs per output value on the CUDA cores would cost 2-3x their HBM time.
// Pointwise layers with Cin % 64 == 0 and everything after the backbone run on the tcgen05 kernel (conv_tcgen05.cu,
// VD_CONV_SLOPE).  All kernels: activation-type NHWC in/out, fp32 accumulate starting from the folded-BN bias,
// y = v > 0 ? v : v * slope[c].
#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 finish8(float (&a)[8], const float* __restrict__ bias, const float* __restrict__ slope) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = a[i] + __ldg(bias + i);
    a[i] = v > 0.f ? v : v * __ldg(slope + i);
  }
  uint4 o;
  o.x = pack_bf16x2(a[0], a[1]);
  o.y = pack_bf16x2(a[2], a[3]);
  o.z = pack_bf16x2(a[4], a[5]);
  o.w = pack_bf16x2(a[6], a[7]);
  return o;
}


// ---- pointwise 1x1 with Cin in {8, 16, 32} on the tensor cores (warp-level mma.sync; K = 8..32, N = 16..64 is far too
// small a GEMM for the tcgen05 kernel's 128 x N tiles, and on the CUDA cores it costs Cin FMAs per output value: the
// 32 -> 32 layer alone would need 125 us of FMA issue at 1080p, three times its HBM time) ----
#if VD_ACT_IS_FP16
#define VD_MMA_T_ "f16"
#else
#define VD_MMA_T_ "bf16"
#endif
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." VD_MMA_T_ "." VD_MMA_T_ ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32." VD_MMA_T_ "." VD_MMA_T_ ".f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
}

// conv_bn(3, 8, stride 2), pad 1, on the ingest layout x0 = [B][H][Wp][4] (pixel x at column x + 3, zero margins), as a
// GEMM on the tensor cores: M = 16 adjacent output pixels of one row per warp step, N = 8, K = 3 filter rows x 16, where
// the 16 values of filter row r are the FOUR padded pixels 2wo+2 .. 2wo+5 of input row 2ho-1+r exactly as they lie in
// memory (32 contiguous bytes; the fourth pixel and every fourth channel meet zero weights).  Output pixel wo+1 starts
// 16 bytes further on, so each register of the A fragment is one coalesced 128-byte load straight from global memory
// (12 per 16 pixels, L1 serves the overlap) and each of the two output registers one coalesced 128-byte store.
// On the CUDA cores this layer costs 27 FMAs per output value — 105 us of FMA issue at 1080p against 60 us of HBM time.
constexpr int STEM_RY = 4, STEM_NB = 4;
__global__ void __launch_bounds__(256)
mnet_stem_kernel(const uint8_t* __restrict__ x0, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ slope, uint32_t* __restrict__ y, int B, int H, int W, int Wp, int Ho, int Wo) {
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  // B fragments: k = s * 4 + c (s < 3 taps, 4 channels), rows 12..15 (the fourth pixel) are zero; w = [(r, s, c4)][8]
  uint32_t bf[3][2];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int k = tig * 2;                                 // k and k + 1: same tap, channels (2 tig) % 4, + 1
    bf[r][0] = static_cast<uint32_t>(w[(r * 12 + k) * 8 + gid]) | (static_cast<uint32_t>(w[(r * 12 + k + 1) * 8 + gid]) << 16);
    bf[r][1] = tig < 2 ? static_cast<uint32_t>(w[(r * 12 + k + 8) * 8 + gid]) | (static_cast<uint32_t>(w[(r * 12 + k + 9) * 8 + gid]) << 16) : 0u;
  }
  const float2 bs = *reinterpret_cast<const float2*>(bias + tig * 2), sl = *reinterpret_cast<const float2*>(slope + tig * 2);
  // one warp step = 16 output columns x STEM_RY output rows: the 2 RY + 1 input rows it needs are all requested before the
  // first mma (an odd input row is filter row 2 of one output row and filter row 0 of the next: loaded once).  A warp
  // takes STEM_NB consecutive column blocks, so the fragment / bias set-up above is paid once per 64 x 4 outputs.
  const int ho0 = blockIdx.y * STEM_RY, b = blockIdx.z;
  const int roww = Wp * 2;                                 // words per padded input row (32-bit offsets: host checks the extent)
  const uint32_t* xrow = reinterpret_cast<const uint32_t*>(x0) + ((b * H + 2 * ho0 - 1) * roww + 4 + tig);
  uint32_t* yrow = y + (b * Ho + ho0) * Wo * 4 + tig;      // 4 words per output pixel
#pragma unroll 1
  for (int nb = 0; nb < STEM_NB; ++nb) {
    const int wo0 = ((blockIdx.x * 8 + (threadIdx.x >> 5)) * STEM_NB + nb) * 16;
    if (wo0 >= Wo) return;
    // rows of the fragment: output pixels wo0 + gid and wo0 + gid + 8 (clamped for the address, masked at the store)
    const int c0 = min(wo0 + gid, Wo - 1), c1 = min(wo0 + gid + 8, Wo - 1);
    const uint32_t* p0 = xrow + 4 * c0;
    const uint32_t* p1 = xrow + 4 * c1;
    uint32_t a[2 * STEM_RY + 1][4];
#pragma unroll
    for (int i = 0; i < 2 * STEM_RY + 1; ++i) {
      const bool ok = static_cast<unsigned int>(2 * ho0 - 1 + i) < static_cast<unsigned int>(H);   // outside the image: zero operand
      a[i][0] = ok ? __ldg(p0) : 0u;
      a[i][1] = ok ? __ldg(p1) : 0u;
      a[i][2] = ok ? __ldg(p0 + 4) : 0u;
      a[i][3] = ok ? __ldg(p1 + 4) : 0u;
      p0 += roww;
      p1 += roww;
    }
    uint32_t* out = yrow + wo0 * 4;
#pragma unroll
    for (int j = 0; j < STEM_RY; ++j) {
      if (ho0 + j >= Ho) break;
      float acc[4] = {bs.x, bs.y, bs.x, bs.y};             // accumulators start at the bias of their two columns
#pragma unroll
      for (int r = 0; r < 3; ++r) mma_k16(acc, a[2 * j + r], bf[r][0], bf[r][1]);
      const float v0 = acc[0] > 0.f ? acc[0] : acc[0] * sl.x, v1 = acc[1] > 0.f ? acc[1] : acc[1] * sl.y;
      const float v2 = acc[2] > 0.f ? acc[2] : acc[2] * sl.x, v3 = acc[3] > 0.f ? acc[3] : acc[3] * sl.y;
      if (wo0 + gid < Wo) out[gid * 4] = pack_bf16x2(v0, v1);
      if (wo0 + gid + 8 < Wo) out[(gid + 8) * 4] = pack_bf16x2(v2, v3);
      out += Wo * 4;
    }
  }
}

// ---- depthwise 3x3 building block ----
// A thread owns CH (= 4: one 8-byte vector) channels of one output column and walks a vertical run of NY outputs.  ALL the input rows the
// run touches ((NY - 1) * STRIDE + 3 rows x left / centre / right) are requested before the first multiply-add — the
// kernels are bound by the latency of these loads, not by their count — and at stride 1 every row feeds up to three
// outputs.  Multiply-adds are FHFMA (activation-type operands, fp32 accumulate, no conversions); per output the taps
// accumulate in (r, s) order; taps outside the image are skipped.
template <int CH> struct ActVec;
template <> struct ActVec<4> { using T = uint2; };
__device__ __forceinline__ void act_fmav(float (&acc)[4], const uint2& x, const uint2& w) {
  acc[0] = act_fma<0, 0>(x.x, w.x, acc[0]); acc[1] = act_fma<1, 1>(x.x, w.x, acc[1]);
  acc[2] = act_fma<0, 0>(x.y, w.y, acc[2]); acc[3] = act_fma<1, 1>(x.y, w.y, acc[3]);
}
__device__ __forceinline__ uint2 vzero(const uint2*) { return make_uint2(0u, 0u); }
// x: tensor as vectors of CH channels, CV vectors per pixel, rs vectors per row; xoff: vector index of (batch, row 0, centre
// column, this thread's channels); yi0: first input row of the run; wtap(t): weight vector of tap t = r * 3 + s.
template <int STRIDE, int NY, int CH>
__device__ __forceinline__ void dw_load(const typename ActVec<CH>::T* __restrict__ x, int xoff, int rs, int CV, int H,
                                        bool has_l, bool has_r, int yi0,
                                        typename ActVec<CH>::T (&in)[(NY - 1) * STRIDE + 3][3]) {
  using V = typename ActVec<CH>::T;
  const V* row = x + (xoff + yi0 * rs);
#pragma unroll
  for (int ir = 0; ir < (NY - 1) * STRIDE + 3; ++ir) {
    const bool ok = static_cast<unsigned int>(yi0 + ir) < static_cast<unsigned int>(H);
    const V z = vzero(static_cast<const V*>(nullptr));
    in[ir][0] = ok && has_l ? __ldg(row - CV) : z;         // taps outside the image: zero operands (their FMAs add +0)
    in[ir][1] = ok ? __ldg(row) : z;
    in[ir][2] = ok && has_r ? __ldg(row + CV) : z;
    row += rs;
  }
}
// acc[o][*] = bias + sum over the 9 taps, (r, s) order.  No branches: rows / columns outside the image hold zeros.
template <int STRIDE, int NY, int CH, typename WF>
__device__ __forceinline__ void dw_compute(const typename ActVec<CH>::T (&in)[(NY - 1) * STRIDE + 3][3], const float (&bias)[CH],
                                           WF wtap, float (&acc)[NY][CH]) {
#pragma unroll
  for (int o = 0; o < NY; ++o)
#pragma unroll
    for (int k = 0; k < CH; ++k) acc[o][k] = bias[k];
#pragma unroll
  for (int ir = 0; ir < (NY - 1) * STRIDE + 3; ++ir) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // input row ir is tap row r of output o when ir = o * STRIDE + r
      if ((ir - r) % STRIDE != 0) continue;
      const int o = (ir - r) / STRIDE;
      if (ir - r < 0 || o >= NY) continue;
      act_fmav(acc[o], in[ir][0], wtap(r * 3 + 0));
      act_fmav(acc[o], in[ir][1], wtap(r * 3 + 1));
      act_fmav(acc[o], in[ir][2], wtap(r * 3 + 2));
    }
  }
}
// slope + rounding of CH biased sums -> one packed vector
__device__ __forceinline__ uint2 leaky_pack(float (&a)[4], const float (&slope)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = a[i] > 0.f ? a[i] : a[i] * slope[i];
  return make_uint2(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]));
}
template <int CH>
__device__ __forceinline__ void load_bias_slope(const float* bias, const float* slope, int g, float (&bs)[CH], float (&sl)[CH]) {
#pragma unroll
  for (int q = 0; q < CH / 4; ++q) {
    const float4 b4 = *(reinterpret_cast<const float4*>(bias) + (CH / 4) * g + q);
    const float4 s4 = *(reinterpret_cast<const float4*>(slope) + (CH / 4) * g + q);
    bs[4 * q] = b4.x; bs[4 * q + 1] = b4.y; bs[4 * q + 2] = b4.z; bs[4 * q + 3] = b4.w;
    sl[4 * q] = s4.x; sl[4 * q + 1] = s4.y; sl[4 * q + 2] = s4.z; sl[4 * q + 3] = s4.w;
  }
}

// depthwise 3x3 (+ folded BN + slope), pad 1, stride 1 or 2, for C >= 64: dw_load + dw_compute with the nine weight vectors
// in registers.  (A rolling-window form — one thread walks a column segment with a ring of six input rows, 3 loads per output
// instead of 4.5 — was measured on the same box at segment lengths 12 / 17 / 23 / whole column: 0.055-0.089 ms against
// 0.043 ms for this kernel on the 64-channel layer and 0.025-0.029 against 0.0225 ms on the 128-channel ones; many
// short-lived threads hide the load latency better than few long-lived ones.)  Adjacent lanes hold adjacent channel vectors, then adjacent columns: every load and store of a warp is one
// contiguous run.  grid = (ceil(Wo * CV / 256), ceil(Ho / NY), B).
template <int STRIDE, int NY, int CH>
__global__ void __launch_bounds__(256)
mnet_dw3x3_kernel(const typename ActVec<CH>::T* __restrict__ x, const typename ActVec<CH>::T* __restrict__ w,
                  const float* __restrict__ bias, const float* __restrict__ slope, typename ActVec<CH>::T* __restrict__ y,
                  int H, int W, int CV, int cv_shift, int Ho, int Wo) {
  using V = typename ActVec<CH>::T;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Wo * CV) return;
  int g, wo;
  if (cv_shift >= 0) { g = i & (CV - 1); wo = i >> cv_shift; } else { wo = i / CV; g = i - wo * CV; }
  const int ho0 = blockIdx.y * NY, b = blockIdx.z;
  V wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = __ldg(w + t * CV + g);
  float bs[CH], sl[CH];
  load_bias_slope<CH>(bias, slope, g, bs, sl);
  float acc[NY][CH];
  const int xc = wo * STRIDE;                              // centre column; left = xc - 1, right = xc + 1
  const int rs = W * CV;                                   // vectors per input row (32-bit offsets: host checks the extent)
  V in[(NY - 1) * STRIDE + 3][3];
  dw_load<STRIDE, NY, CH>(x, b * H * rs + xc * CV + g, rs, CV, H, xc >= 1, xc + 1 < W, ho0 * STRIDE - 1, in);
  dw_compute<STRIDE, NY, CH>(in, bs, [&](int t) { return wv[t]; }, acc);
  const int yoff = (b * Ho * Wo + wo) * CV + g, yrs = Wo * CV;
#pragma unroll
  for (int o = 0; o < NY; ++o)
    if (ho0 + o < Ho) y[yoff + (ho0 + o) * yrs] = leaky_pack(acc[o], sl);
}

// Shared-memory layout of a pixel tile: one row of CIN (COUT) activation-type values per pixel; the row pitch keeps the
// eight 16-byte rows of an ldmatrix 8x8 block (and the eight rows a warp's epilogue writes) on distinct banks.
template <int C> struct PwPitch { static constexpr int value = C * 2 + (C == 8 ? 0 : 16); };

// B fragments of the whole [CIN][COUT] weight matrix, pre-packed per lane in shared memory: word ((ks * NT + nt) * NW + h) * 32
// + lane is the lane's register h of the (k-step ks, column block nt) fragment — (CIN / 16) k-steps x (COUT / 8) blocks x 2
// words (CIN = 8: one m16n8k8 step, 1 word).  Read back with conflict-free 4-byte loads at every use: 16..32 registers
// that a persistent CTA would otherwise carry through its depthwise phase.
template <int CIN, int COUT>
struct PwFrag {
  static constexpr int KS = CIN >= 16 ? CIN / 16 : 1, NT = COUT / 8, NW = CIN >= 16 ? 2 : 1;
  static constexpr int WORDS = KS * NT * NW * 32;
  __device__ static void pack(uint32_t* tab, const uint16_t* w, int tid, int nthreads) {      // w: [CIN][COUT]
    for (int e = tid; e < WORDS; e += nthreads) {
      const int lane = e & 31, f = e >> 5;
      const int h = f % NW, nt = (f / NW) % NT, ks = f / (NW * NT);
      const int k = ks * 16 + h * 8 + (lane & 3) * 2, n = nt * 8 + (lane >> 2);
      tab[e] = static_cast<uint32_t>(w[k * COUT + n]) | (static_cast<uint32_t>(w[(k + 1) * COUT + n]) << 16);
    }
  }
};

// One warp: 16 pixels (rows of `in`, pitch PwPitch<CIN>) x COUT outputs = act(in * W + bias), written as packed rows into the
// warp's staging area `st` (16 rows, pitch PwPitch<COUT>).  bs / sl: bias and slope in shared memory.
template <int CIN, int COUT>
__device__ __forceinline__ void pw_tile16(const uint8_t* in, const uint32_t* bfrag, const float* bs, const float* sl,
                                          uint8_t* st, int lane) {
  using F = PwFrag<CIN, COUT>;
  constexpr int PI = PwPitch<CIN>::value, PO = PwPitch<COUT>::value, NT = COUT / 8;
  const int gid = lane >> 2, tig = lane & 3;
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {                       // accumulators start at the bias of their two columns
    const float2 b2 = *reinterpret_cast<const float2*>(bs + nt * 8 + tig * 2);
    acc[nt][0] = acc[nt][2] = b2.x;
    acc[nt][1] = acc[nt][3] = b2.y;
  }
  if constexpr (CIN == 8) {
    uint32_t a[4];
    ldsm_x2(a, in + (lane & 15) * PI);                                  // rows 0-7 | rows 8-15, k 0-7
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_k8(acc[nt], a, bfrag[nt * 32 + lane]);
  } else {
#pragma unroll
    for (int ks = 0; ks < CIN / 16; ++ks) {
      uint32_t a[4];
      // lanes 0-7: rows 0-7 k 0-7 | 8-15: rows 8-15 k 0-7 | 16-23: rows 0-7 k 8-15 | 24-31: rows 8-15 k 8-15
      ldsm_x4(a, in + (lane & 15) * PI + ks * 32 + (lane >> 4) * 16);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        mma_k16(acc[nt], a, bfrag[((ks * NT + nt) * 2) * 32 + lane], bfrag[((ks * NT + nt) * 2 + 1) * 32 + lane]);
    }
  }
  (void)F::WORDS;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = nt * 8 + tig * 2;
    const float2 s2 = *reinterpret_cast<const float2*>(sl + col);
    float v0 = acc[nt][0], v1 = acc[nt][1], v2 = acc[nt][2], v3 = acc[nt][3];
    v0 = v0 > 0.f ? v0 : v0 * s2.x; v1 = v1 > 0.f ? v1 : v1 * s2.y;
    v2 = v2 > 0.f ? v2 : v2 * s2.x; v3 = v3 > 0.f ? v3 : v3 * s2.y;
    *reinterpret_cast<uint32_t*>(st + gid * PO + col * 2) = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(st + (gid + 8) * PO + col * 2) = pack_bf16x2(v2, v3);
  }
}

// (For the 8 -> 16 block a variant that keeps whole pixels in one thread — depthwise result in registers, 8 x 16 FHFMA per
// pixel against broadcast weights, no barrier, no staging — was measured against this kernel on the same box: 0.146 vs
// 0.141 ms.  That block writes twice what it reads (265 vs 133 MB) and runs at 2.8 TB/s either way; dropped.)
// conv_dw block = depthwise 3x3 (+BN+slope) followed by pointwise 1x1 (+BN+slope) in ONE kernel for Cin <= 32: the
// depthwise result of a TX x TY pixel tile goes to shared memory (rounded to the activation type exactly where the
// two-kernel form stores it) and never to HBM.  Persistent CTAs walk the tiles (weights, fragment table, biases are set
// up once); per tile: phase 1 = dw_compute on input rows that were requested BEFORE the previous tile's phase 2 (the
// kernel is bound by the latency of those loads: a first version that loaded, computed and stored tile by tile spent 39 %
// of its warp samples waiting on them), phase 2 = pw_tile16 per 16 pixels of a tile row, staged and written back as
// 16-byte vectors.  One thread of phase 1 = CH channels x NY vertically adjacent outputs; TX * TY * (CIN / CH) = 256 * NY.
template <int CIN, int COUT, int STRIDE, int CH>
struct DwpwCfg {
  static constexpr int GV = CIN / CH, NY = STRIDE == 1 ? 4 : 2;
  static constexpr int NPX = 256 * NY / GV;
  static constexpr int TX = NPX >= 1024 ? 64 : (NPX >= 256 ? 32 : 16), TY = NPX / TX;
};

template <int CIN, int COUT, int STRIDE, int CH>
__global__ void __launch_bounds__(256, 2)
mnet_dwpw_kernel(const typename ActVec<CH>::T* __restrict__ x, const typename ActVec<CH>::T* __restrict__ wdw,
                 const float* __restrict__ bdw, const float* __restrict__ sdw, const uint16_t* __restrict__ wpw,
                 const float* __restrict__ bpw, const float* __restrict__ spw, uint4* __restrict__ y, int B, int H, int W,
                 int Ho, int Wo) {
  using Cfg = DwpwCfg<CIN, COUT, STRIDE, CH>;
  using V = typename ActVec<CH>::T;
  using F = PwFrag<CIN, COUT>;
  constexpr int GV = Cfg::GV, NY = Cfg::NY, TX = Cfg::TX, TY = Cfg::TY, NPX = Cfg::NPX;
  constexpr int PI = PwPitch<CIN>::value, PO = PwPitch<COUT>::value;
  static_assert(TX * GV * (TY / NY) == 256 && TY % NY == 0 && TX % 16 == 0 && NPX % 16 == 0, "tile / thread mapping");
  __shared__ __align__(16) uint8_t sD[NPX * PI];        // depthwise result, one row per pixel
  __shared__ __align__(16) uint8_t sO[8 * 16 * PO];     // per-warp output staging
  __shared__ __align__(16) uint32_t s_bf[F::WORDS];     // pointwise B fragments, per lane
  __shared__ __align__(16) float s_b[2 * COUT];         // bpw | spw
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  F::pack(s_bf, wpw, tid, 256);
  for (int i = tid; i < COUT; i += 256) {
    s_b[i] = bpw[i];
    s_b[COUT + i] = spw[i];
  }
  const int ntx = (Wo + TX - 1) / TX, nty = (Ho + TY - 1) / TY;
  const int total = B * nty * ntx;
  // this thread's part of a tile's depthwise phase
  const int cg = tid % (TX * GV), run = tid / (TX * GV);
  const int g = cg % GV, tx = cg / GV, ty0 = run * NY;
  const int rs = W * GV;
  V wv[9];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) wv[tp] = __ldg(wdw + tp * GV + g);
  float bsd[CH], sld[CH];
  load_bias_slope<CH>(bdw, sdw, g, bsd, sld);
  V in[(NY - 1) * STRIDE + 3][3];
  // tiles t, t + gridDim.x, ... as (batch, tile row, tile column), advanced without divisions
  int t = blockIdx.x;
  int txi = t % ntx, tyi = (t / ntx) % nty, tb = t / (ntx * nty);
  const int dxi = gridDim.x % ntx, dr = gridDim.x / ntx, dyi = dr % nty, db = dr / nty;
  bool live = false;
  auto load_tile = [&]() {                                 // request the input rows of tile (tb, tyi, txi)
    const int ox = txi * TX + tx, xc = ox * STRIDE, oy = tyi * TY + ty0;
    live = ox < Wo && oy < Ho;
    if (live) dw_load<STRIDE, NY, CH>(x, tb * H * rs + xc * GV + g, rs, GV, H, xc >= 1, xc + 1 < W, oy * STRIDE - 1, in);
  };
  if (t < total) load_tile();
  __syncthreads();                                       // the shared tables are complete
  uint8_t* st = sO + warp * 16 * PO;
  while (t < total) {
    const int cb = tb, coy0 = tyi * TY, cox0 = txi * TX;  // tile t's place
    // ---- phase 1: depthwise of tile t (rows already requested) into shared memory ----
    {
      float acc[NY][CH];
      if (live) dw_compute<STRIDE, NY, CH>(in, bsd, [&](int tap) { return wv[tap]; }, acc);
#pragma unroll
      for (int o = 0; o < NY; ++o) {
        V out = vzero(static_cast<const V*>(nullptr));
        if (live && coy0 + ty0 + o < Ho) out = leaky_pack(acc[o], sld);
        *reinterpret_cast<V*>(sD + ((ty0 + o) * TX + tx) * PI + g * CH * 2) = out;
      }
    }
    __syncthreads();
    // ---- request the next tile's input rows; they arrive while phase 2 runs ----
    const int tn = t + gridDim.x;
    txi += dxi; tyi += dyi; tb += db;
    if (txi >= ntx) { txi -= ntx; ++tyi; }
    if (tyi >= nty) { tyi -= nty; ++tb; }
    if (tn < total) load_tile();
    // ---- phase 2: pointwise on the tensor cores ----
    for (int mt = warp; mt < NPX / 16; mt += 8) {
      const int p0 = mt * 16, ty = p0 / TX, tx0 = p0 % TX;
      const int oy = coy0 + ty;
      if (oy >= Ho || cox0 + tx0 >= Wo) continue;           // (warp-uniform)
      pw_tile16<CIN, COUT>(sD + p0 * PI, s_bf, s_b, s_b + COUT, st, lane);
      __syncwarp();
      uint4* yrow = y + (static_cast<size_t>(cb) * Ho + oy) * Wo * (COUT / 8) + static_cast<size_t>(cox0 + tx0) * (COUT / 8);
      const int npx = min(16, Wo - (cox0 + tx0));
      for (int i = lane; i < npx * (COUT / 8); i += 32)
        yrow[i] = *reinterpret_cast<const uint4*>(st + (i / (COUT / 8)) * PO + (i % (COUT / 8)) * 16);
      __syncwarp();
    }
    __syncthreads();                                     // every warp is done with sD before the next tile overwrites it
    t = tn;
  }
}

// pointwise 1x1 alone, for the (Cin, Cout) pairs of the fused kernel: the same pw_tile16 on 256 consecutive pixels per
// CTA, so the two-kernel form of a conv_dw block gives the same bits as the fused one.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
mnet_pw_mma_kernel(const uint4* __restrict__ x, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ slope, uint4* __restrict__ y, long long npix) {
  constexpr int G = CIN / 8, PI = PwPitch<CIN>::value, PO = PwPitch<COUT>::value, NPX = 256;
  __shared__ __align__(16) uint8_t sD[NPX * PI];
  __shared__ __align__(16) uint8_t sO[8 * 16 * PO];
  __shared__ __align__(16) uint32_t s_bf[PwFrag<CIN, COUT>::WORDS];
  __shared__ __align__(16) float s_b[2 * COUT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  PwFrag<CIN, COUT>::pack(s_bf, w, tid, 256);
  for (int i = tid; i < COUT; i += 256) {
    s_b[i] = bias[i];
    s_b[COUT + i] = slope[i];
  }
  const long long p0 = static_cast<long long>(blockIdx.x) * NPX;
  const int n = static_cast<int>(min(static_cast<long long>(NPX), npix - p0));
  for (int i = tid; i < NPX * G; i += 256) {
    const int px = i / G, g = i % G;
    *reinterpret_cast<uint4*>(sD + px * PI + g * 16) = px < n ? __ldg(x + p0 * G + i) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  uint8_t* st = sO + warp * 16 * PO;
  for (int mt = warp; mt < NPX / 16; mt += 8) {
    if (mt * 16 >= n) continue;
    pw_tile16<CIN, COUT>(sD + mt * 16 * PI, s_bf, s_b, s_b + COUT, st, lane);
    __syncwarp();
    uint4* yrow = y + (p0 + mt * 16) * (COUT / 8);
    const int npx = min(16, n - mt * 16);
    for (int i = lane; i < npx * (COUT / 8); i += 32)
      yrow[i] = *reinterpret_cast<const uint4*>(st + (i / (COUT / 8)) * PO + (i % (COUT / 8)) * 16);
    __syncwarp();
  }
}

// pointwise 1x1, Cin * Cout <= 4096: weights [Cin][Cout] in shared memory as fp32, one thread = one pixel x 8 outputs.
__global__ void __launch_bounds__(256)
mnet_pw_kernel(const uint4* __restrict__ x, const uint16_t* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ slope, uint4* __restrict__ y, long long npix, int Cin, int Cout) {
  __shared__ uint4 sw[4096 / 8];                        // [Cin][Cout] activation-type, as stored
  for (int i = threadIdx.x; i < Cin * Cout / 8; i += blockDim.x) sw[i] = reinterpret_cast<const uint4*>(w)[i];
  __syncthreads();
  const int Co8 = Cout >> 3, Ci8 = Cin >> 3;
  const long long total = npix * Co8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % Co8);
    const long long p = idx / Co8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c8 = 0; c8 < Ci8; ++c8) {
      const uint4 xv = __ldg(x + p * Ci8 + c8);
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const uint32_t xw = j2 == 0 ? xv.x : j2 == 1 ? xv.y : j2 == 2 ? xv.z : xv.w;
        act_fma8s<0>(acc, xw, sw[(c8 * 8 + 2 * j2) * Co8 + g]);
        act_fma8s<1>(acc, xw, sw[(c8 * 8 + 2 * j2 + 1) * Co8 + g]);
      }
    }
    y[idx] = finish8(acc, bias + g * 8, slope + g * 8);
  }
}

int grid_for(long long total) {
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  const long long want = (total + 255) / 256;
  const long long cap = static_cast<long long>(sms) * 16;
  return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" int vd_mnet_stem(const void* x0, const void* w, const float* bias, const float* slope, void* y, int B, int H,
                            int W, vd_stream_t stream) {
  VD_REQUIRE(x0 && w && bias && slope && y, "vd_mnet_stem: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0, "vd_mnet_stem: bad extent");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "vd_mnet_stem: weights must be 16-byte aligned");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  VD_REQUIRE(B <= 65535 && (Ho + STEM_RY - 1) / STEM_RY <= 65535, "vd_mnet_stem: batch / height beyond the grid limits");
  VD_REQUIRE(static_cast<long long>(B) * H * vd_x0_pitch(W) * 2 < (1ll << 31), "vd_mnet_stem: input beyond the kernel's 32-bit indexing");
  const dim3 grid(((Wo + 15) / 16 + 8 * STEM_NB - 1) / (8 * STEM_NB), (Ho + STEM_RY - 1) / STEM_RY, B);
  mnet_stem_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(x0), static_cast<const uint16_t*>(w), bias, slope, static_cast<uint32_t*>(y), B, H, W,
      vd_x0_pitch(W), Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

extern "C" int vd_mnet_dw3x3(const void* x, const void* w, const float* bias, const float* slope, void* y, int B, int H,
                             int W, int C, int stride, vd_stream_t stream) {
  VD_REQUIRE(x && w && bias && slope && y, "vd_mnet_dw3x3: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
             "vd_mnet_dw3x3: C=%d must be a multiple of 8, stride 1 or 2", C);
  VD_REQUIRE(B <= 65535 && H <= 65535 && static_cast<long long>(B) * H * W * (C / 4) < (1ll << 31),
             "vd_mnet_dw3x3: batch / height / tensor size beyond the kernel's 32-bit indexing");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "vd_mnet_dw3x3: tensors must be 16-byte aligned");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(slope)) & 15) == 0,
             "vd_mnet_dw3x3: bias / slope must be 16-byte aligned");
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 4 channels per thread (18 weight + 36 input registers at NY = 4) keeps three CTAs per SM with every load in flight;
  // 8 channels per thread measured 0.065 vs 0.045 ms on the 64-channel layer, runs of 2 instead of 4 outputs 0.053
  const int CV = C / 4;
  int sh = -1;
  if ((CV & (CV - 1)) == 0) { sh = 0; while ((1 << sh) < CV) ++sh; }
  const int ny = stride == 2 ? 2 : 4;
  const dim3 grid((Wo * CV + 255) / 256, (Ho + ny - 1) / ny, B);
  if (stride == 2)
    mnet_dw3x3_kernel<2, 2, 4><<<grid, 256, 0, st>>>(static_cast<const uint2*>(x), static_cast<const uint2*>(w), bias, slope,
                                                    static_cast<uint2*>(y), H, W, CV, sh, Ho, Wo);
  else
    mnet_dw3x3_kernel<1, 4, 4><<<grid, 256, 0, st>>>(static_cast<const uint2*>(x), static_cast<const uint2*>(w), bias, slope,
                                                    static_cast<uint2*>(y), H, W, CV, sh, Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

template <int CIN, int COUT, int STRIDE, int CH>
static void launch_dwpw2(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                         const float* bpw, const float* spw, void* y, int B, int H, int W, int Ho, int Wo, cudaStream_t st) {
  using Cfg = DwpwCfg<CIN, COUT, STRIDE, CH>;
  using V = typename ActVec<CH>::T;
  static const int per_sm = [] {                          // resident CTAs per SM of this instantiation
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, mnet_dwpw_kernel<CIN, COUT, STRIDE, CH>, 256, 0) != cudaSuccess || n < 1) n = 2;
    return n;
  }();
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  const long long total = static_cast<long long>(B) * ((Ho + Cfg::TY - 1) / Cfg::TY) * ((Wo + Cfg::TX - 1) / Cfg::TX);
  const int grid = static_cast<int>(total < static_cast<long long>(sms) * per_sm ? total : static_cast<long long>(sms) * per_sm);
  mnet_dwpw_kernel<CIN, COUT, STRIDE, CH><<<grid, 256, 0, st>>>(static_cast<const V*>(x), static_cast<const V*>(wdw), bdw, sdw,
                                                               static_cast<const uint16_t*>(wpw), bpw, spw,
                                                               static_cast<uint4*>(y), B, H, W, Ho, Wo);
}

template <int CIN, int COUT>
static void launch_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                        const float* bpw, const float* spw, void* y, int B, int H, int W, int stride, int Ho, int Wo,
                        cudaStream_t st) {
  // 4 channels per depthwise thread: 8 measured 0.212 vs 0.160 ms on the 32 -> 32 block (144-168 registers)
  if (stride == 1) launch_dwpw2<CIN, COUT, 1, 4>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, Ho, Wo, st);
  else launch_dwpw2<CIN, COUT, 2, 4>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, Ho, Wo, st);
}

extern "C" int vd_mnet_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                            const float* bpw, const float* spw, void* y, int B, int H, int W, int Cin, int Cout,
                            int stride, vd_stream_t stream) {
  VD_REQUIRE(x && wdw && bdw && sdw && wpw && bpw && spw && y, "vd_mnet_dwpw: null argument");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(wdw) | reinterpret_cast<uintptr_t>(wpw) | reinterpret_cast<uintptr_t>(x) |
               reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bdw) | reinterpret_cast<uintptr_t>(sdw)) & 15) == 0,
             "vd_mnet_dwpw: tensors, weights and depthwise bias / slope must be 16-byte aligned");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (stride == 1 || stride == 2), "vd_mnet_dwpw: bad extent / stride");
  VD_REQUIRE(Cin > 0 && static_cast<long long>(B) * H * W * (Cin / 4) < (1ll << 31), "vd_mnet_dwpw: tensor beyond the kernel's 32-bit indexing");
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  VD_REQUIRE(static_cast<long long>(B) * ((Ho + 3) / 4) * ((Wo + 15) / 16) < (1ll << 31), "vd_mnet_dwpw: too many tiles");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin == 8 && Cout == 16) launch_dwpw<8, 16>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, stride, Ho, Wo, st);
  else if (Cin == 16 && Cout == 32) launch_dwpw<16, 32>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, stride, Ho, Wo, st);
  else if (Cin == 32 && Cout == 32) launch_dwpw<32, 32>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, stride, Ho, Wo, st);
  else if (Cin == 32 && Cout == 64) launch_dwpw<32, 64>(x, wdw, bdw, sdw, wpw, bpw, spw, y, B, H, W, stride, Ho, Wo, st);
  else {
    vd_set_error("vd_mnet_dwpw: (Cin, Cout) = (%d, %d) is not one of (8,16) (16,32) (32,32) (32,64)", Cin, Cout);
    return VD_ERR_UNSUPPORTED;
  }
  VD_LAUNCH_CHECK();
  return VD_OK;
}

template <int CIN, int COUT>
static void launch_pw_mma(const void* x, const void* w, const float* bias, const float* slope, void* y, long long npix,
                          cudaStream_t st) {
  mnet_pw_mma_kernel<CIN, COUT><<<static_cast<unsigned int>((npix + 255) / 256), 256, 0, st>>>(
      static_cast<const uint4*>(x), static_cast<const uint16_t*>(w), bias, slope, static_cast<uint4*>(y), npix);
}

extern "C" int vd_mnet_pw(const void* x, const void* w, const float* bias, const float* slope, void* y, long long npix,
                          int Cin, int Cout, vd_stream_t stream) {
  VD_REQUIRE(x && w && bias && slope && y, "vd_mnet_pw: null argument");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "vd_mnet_pw: weights must be 16-byte aligned");
  VD_REQUIRE(npix > 0 && Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0 && Cin * Cout <= 4096,
             "vd_mnet_pw: Cin=%d Cout=%d must be multiples of 8 with Cin*Cout <= 4096", Cin, Cout);
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "vd_mnet_pw: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the pairs of the fused conv_dw kernel run the same tensor-core tile code (same bits as vd_mnet_dwpw)
  if (npix + 255 < (1ll << 31) * 256 && Cin == 8 && Cout == 16) launch_pw_mma<8, 16>(x, w, bias, slope, y, npix, st);
  else if (npix + 255 < (1ll << 31) * 256 && Cin == 16 && Cout == 32) launch_pw_mma<16, 32>(x, w, bias, slope, y, npix, st);
  else if (npix + 255 < (1ll << 31) * 256 && Cin == 32 && Cout == 32) launch_pw_mma<32, 32>(x, w, bias, slope, y, npix, st);
  else if (npix + 255 < (1ll << 31) * 256 && Cin == 32 && Cout == 64) launch_pw_mma<32, 64>(x, w, bias, slope, y, npix, st);
  else
    mnet_pw_kernel<<<grid_for(npix * (Cout / 8)), 256, 0, st>>>(
        static_cast<const uint4*>(x), static_cast<const uint16_t*>(w), bias, slope, static_cast<uint4*>(y), npix, Cin, Cout);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
