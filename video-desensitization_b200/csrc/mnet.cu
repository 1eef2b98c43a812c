// mnet.cu — the sub-64-channel layers of the MobileNet0.25 variant (Retinaface(backbone="mobilenet")).
//
// Reference: detect_face/nets/mobilenet025.py:3-19 — conv_bn(3, 8, stride 2) and the conv_dw blocks
// (depthwise 3x3 + BN + LeakyReLU(0.1), pointwise 1x1 + BN + LeakyReLU(0.1)).  These layers move 8..32-channel
// tensors at 1/2 and 1/4 resolution: a few MACs per byte, HBM-bound — streaming kernels on the CUDA cores with
// 16-byte accesses (the rule for this tier: do not reshape byte-bound work into GEMMs).  Pointwise layers with
// Cin % 64 == 0 and everything after the backbone run on the tcgen05 kernel (conv_tcgen05.cu, VD_CONV_SLOPE).
// All kernels: activation-type NHWC in/out, fp32 accumulate, folded-BN bias, y = v > 0 ? v : v * slope[c].
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

// same with bias / slope in shared memory (plain loads)
__device__ __forceinline__ uint4 finish8s(float (&a)[8], const float* bias, const float* slope) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = a[i] + bias[i];
    a[i] = v > 0.f ? v : v * slope[i];
  }
  uint4 o;
  o.x = pack_bf16x2(a[0], a[1]);
  o.y = pack_bf16x2(a[2], a[3]);
  o.z = pack_bf16x2(a[4], a[5]);
  o.w = pack_bf16x2(a[6], a[7]);
  return o;
}

// conv_bn(3, 8, stride 2), pad 1, on the ingest layout x0 = [B][H][Wp][4] (pixel x at column x + 3, zero margins):
// one thread = TWO adjacent output pixels = 3 rows x 5 input pixels (40 contiguous bytes per row), 2 x 16 bytes out;
// the 216 weights sit in shared memory as stored (eight output channels per uint4) and every weight load (same address
// for the whole warp) feeds both pixels.  All multiply-adds are FHFMA (activation-type operands, fp32 accumulate): no
// conversion instructions.
__global__ void __launch_bounds__(256)
mnet_stem_kernel(const uint2* __restrict__ x0, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ slope, uint4* __restrict__ y, int B, int H, int W, int Wp, int Ho, int Wo) {
  __shared__ uint4 sw[36];                            // [(r, s, c4)][8 cout] activation-type, as stored
  for (int i = threadIdx.x; i < 36; i += blockDim.x) sw[i] = reinterpret_cast<const uint4*>(w)[i];
  __syncthreads();
  const int Wo2 = (Wo + 1) >> 1;
  const long long total = static_cast<long long>(B) * Ho * Wo2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(idx % Wo2);
    const int ho = static_cast<int>((idx / Wo2) % Ho);
    const int b = static_cast<int>(idx / (static_cast<long long>(Wo2) * Ho));
    const int wo = 2 * wp;
    float a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yi = 2 * ho - 1 + r;
      if (yi < 0 || yi >= H) continue;
      // input columns 2wo-1 .. 2wo+3 are padded columns 2wo+2 .. 2wo+6 (inside the row: Wp >= 2*Wo + 6); 2wo+2 is even,
      // i.e. 16-byte aligned
      const uint2* row = x0 + (static_cast<size_t>(b) * H + yi) * Wp + (2 * wo + 2);
      const uint4 p01 = __ldg(reinterpret_cast<const uint4*>(row));
      const uint4 p23 = __ldg(reinterpret_cast<const uint4*>(row) + 1);
      const uint2 p4 = __ldg(row + 4);
      const uint32_t px[5][2] = {{p01.x, p01.y}, {p01.z, p01.w}, {p23.x, p23.y}, {p23.z, p23.w}, {p4.x, p4.y}};
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) {
        const uint4 w0 = sw[(r * 3 + sx) * 4], w1 = sw[(r * 3 + sx) * 4 + 1], w2 = sw[(r * 3 + sx) * 4 + 2];
        act_fma8s<0>(a0, px[sx][0], w0); act_fma8s<1>(a0, px[sx][0], w1); act_fma8s<0>(a0, px[sx][1], w2);
        act_fma8s<0>(a1, px[sx + 2][0], w0); act_fma8s<1>(a1, px[sx + 2][0], w1); act_fma8s<0>(a1, px[sx + 2][1], w2);
      }
    }
    uint4* out = y + (static_cast<size_t>(b) * Ho + ho) * Wo + wo;
    out[0] = finish8(a0, bias, slope);
    if (wo + 1 < Wo) out[1] = finish8(a1, bias, slope);
  }
}

// The depthwise run of one thread (see mnet_dw3x3_kernel): acc[o][0..7] for the NY outputs (ho0 + o, wo), channels 8g..8g+7.
template <int STRIDE, int NY>
__device__ __forceinline__ void dw_run(const uint4* __restrict__ x, int xoff, int rs, int C8, int H, bool has_l, bool has_r,
                                       int yi0, const uint4 (&wv)[9], float (&acc)[NY][8]) {
#pragma unroll
  for (int o = 0; o < NY; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;
  constexpr int NIN = (NY - 1) * STRIDE + 3;               // input rows the run touches
#pragma unroll
  for (int ir = 0; ir < NIN; ++ir) {
    const int yi = yi0 + ir;
    if (yi < 0 || yi >= H) continue;
    const uint4* row = x + (xoff + yi * rs);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint4 vl = has_l ? __ldg(row - C8) : z, vc = __ldg(row), vr = has_r ? __ldg(row + C8) : z;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // input row ir is tap row r of output o when ir = o * STRIDE + r
      if ((ir - r) % STRIDE != 0) continue;
      const int o = (ir - r) / STRIDE;
      if (ir - r < 0 || o >= NY) continue;
      if (has_l) act_fma8(acc[o], vl, wv[r * 3 + 0]);      // (a zero tap adds +0: skipping it keeps the bits)
      act_fma8(acc[o], vc, wv[r * 3 + 1]);
      if (has_r) act_fma8(acc[o], vr, wv[r * 3 + 2]);
    }
  }
}

// depthwise 3x3, pad 1, stride 1 or 2.  One thread = 8 channels x a vertical run of NY output pixels of one column:
// the nine weight vectors stay packed in registers for the whole run, and at stride 1 every input row that is loaded
// (three 16-byte loads: left, centre, right) feeds up to three output rows — (NY + 2) * 3 loads per NY outputs instead
// of 9 * NY, plus 9 weight loads per run instead of per output.  Adjacent lanes hold adjacent channel groups, then
// adjacent columns: every load and store of a warp is one contiguous 512-byte run.  All multiply-adds are FHFMA
// (act_fma8: activation-type operands, fp32 accumulate, no conversions); per output the taps accumulate in (r, s) order.
// grid = (ceil(Wo * C8 / 256), ceil(Ho / NY), B).
template <int STRIDE, int NY>
__global__ void __launch_bounds__(256)
mnet_dw3x3_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ slope, uint4* __restrict__ y, int H, int W, int C8, int c8_shift, int Ho, int Wo) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= Wo * C8) return;
  int g, wo;
  if (c8_shift >= 0) { g = i & (C8 - 1); wo = i >> c8_shift; } else { wo = i / C8; g = i - wo * C8; }
  const int ho0 = blockIdx.y * NY, b = blockIdx.z;
  uint4 wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = __ldg(w + t * C8 + g);
  float acc[NY][8];
  const int xc = wo * STRIDE;                              // centre column; left = xc - 1, right = xc + 1
  const int rs = W * C8;                                   // uint4 per input row (32-bit offsets: host checks the extent)
  dw_run<STRIDE, NY>(x, b * H * rs + xc * C8 + g, rs, C8, H, xc >= 1, xc + 1 < W, ho0 * STRIDE - 1, wv, acc);
  float bs[8], sl[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * g), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * g + 1);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(slope) + 2 * g), s1 = __ldg(reinterpret_cast<const float4*>(slope) + 2 * g + 1);
    bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
    sl[0] = s0.x; sl[1] = s0.y; sl[2] = s0.z; sl[3] = s0.w; sl[4] = s1.x; sl[5] = s1.y; sl[6] = s1.z; sl[7] = s1.w;
  }
  const int yoff = (b * Ho * Wo + wo) * C8 + g, yrs = Wo * C8;
#pragma unroll
  for (int o = 0; o < NY; ++o)
    if (ho0 + o < Ho) y[yoff + (ho0 + o) * yrs] = finish8s(acc[o], bs, sl);
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

// Shared-memory layout of a pixel tile: one row of CIN (COUT) activation-type values per pixel; the row pitch keeps the
// eight 16-byte rows of an ldmatrix 8x8 block (and the eight rows a warp's epilogue writes) on distinct banks.
template <int C> struct PwPitch { static constexpr int value = C * 2 + (C == 8 ? 0 : 16); };

// B fragments of the whole [CIN][COUT] weight matrix, per thread: (CIN / 16) k-steps x (COUT / 8) column blocks x 2 words
// (CIN = 8: one m16n8k8 step, 1 word).
template <int CIN, int COUT>
struct PwWeights {
  static constexpr int KS = CIN >= 16 ? CIN / 16 : 1, NT = COUT / 8, NW = CIN >= 16 ? 2 : 1;
  uint32_t b[KS][NT][NW];
  __device__ __forceinline__ void load(const uint16_t* w, int lane) {       // w: [CIN][COUT] (shared or global memory)
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int h = 0; h < NW; ++h) {
          const int k = ks * 16 + h * 8 + tig * 2, n = nt * 8 + gid;
          b[ks][nt][h] = static_cast<uint32_t>(w[k * COUT + n]) | (static_cast<uint32_t>(w[(k + 1) * COUT + n]) << 16);
        }
  }
};

// One warp: 16 pixels (rows of `in`, pitch PwPitch<CIN>) x COUT outputs = act(in * W + bias), written as packed rows into the
// warp's staging area `st` (16 rows, pitch PwPitch<COUT>).  bs / sl: bias and slope in shared memory.
template <int CIN, int COUT>
__device__ __forceinline__ void pw_tile16(const uint8_t* in, const PwWeights<CIN, COUT>& wf, const float* bs, const float* sl,
                                          uint8_t* st, int lane) {
  constexpr int PI = PwPitch<CIN>::value, PO = PwPitch<COUT>::value, NT = COUT / 8;
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  if constexpr (CIN == 8) {
    uint32_t a[4];
    ldsm_x2(a, in + (lane & 15) * PI);                                  // rows 0-7 | rows 8-15, k 0-7
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_k8(acc[nt], a, wf.b[0][nt][0]);
  } else {
#pragma unroll
    for (int ks = 0; ks < CIN / 16; ++ks) {
      uint32_t a[4];
      // lanes 0-7: rows 0-7 k 0-7 | 8-15: rows 8-15 k 0-7 | 16-23: rows 0-7 k 8-15 | 24-31: rows 8-15 k 8-15
      ldsm_x4(a, in + (lane & 15) * PI + ks * 32 + (lane >> 4) * 16);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) mma_k16(acc[nt], a, wf.b[ks][nt][0], wf.b[ks][nt][1]);
    }
  }
  const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = nt * 8 + tig * 2;
    const float2 b2 = *reinterpret_cast<const float2*>(bs + col), s2 = *reinterpret_cast<const float2*>(sl + col);
    float v0 = acc[nt][0] + b2.x, v1 = acc[nt][1] + b2.y, v2 = acc[nt][2] + b2.x, v3 = acc[nt][3] + b2.y;
    v0 = v0 > 0.f ? v0 : v0 * s2.x; v1 = v1 > 0.f ? v1 : v1 * s2.y;
    v2 = v2 > 0.f ? v2 : v2 * s2.x; v3 = v3 > 0.f ? v3 : v3 * s2.y;
    *reinterpret_cast<uint32_t*>(st + gid * PO + col * 2) = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(st + (gid + 8) * PO + col * 2) = pack_bf16x2(v2, v3);
  }
}

// conv_dw block = depthwise 3x3 (+BN+slope) followed by pointwise 1x1 (+BN+slope) in ONE kernel for Cin <= 32: the
// depthwise result of a TX x TY pixel tile goes to shared memory (rounded to the activation type exactly where the
// two-kernel form stores it) and never to HBM.  Phase 1 = dw_run (one thread: 8 channels x NY vertically adjacent
// outputs); phase 2 = pw_tile16 per 16 pixels of a tile row, staged and written back as 16-byte vectors.
// TX * TY * (CIN / 8) = 256 * NY;  grid = (ceil(Wo / TX), ceil(Ho / TY), B).
template <int CIN, int COUT, int STRIDE>
struct DwpwCfg {
  static constexpr int G = CIN / 8, NY = STRIDE == 1 ? 4 : 2;
  static constexpr int NPX = 256 * NY / G;
  static constexpr int TX = NPX >= 1024 ? 64 : (NPX >= 256 ? 32 : 16), TY = NPX / TX;
};

template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(256)
mnet_dwpw_kernel(const uint4* __restrict__ x, const uint4* __restrict__ wdw, const float* __restrict__ bdw,
                 const float* __restrict__ sdw, const uint16_t* __restrict__ wpw, const float* __restrict__ bpw,
                 const float* __restrict__ spw, uint4* __restrict__ y, int H, int W, int Ho, int Wo) {
  using Cfg = DwpwCfg<CIN, COUT, STRIDE>;
  constexpr int G = Cfg::G, NY = Cfg::NY, TX = Cfg::TX, TY = Cfg::TY, NPX = Cfg::NPX;
  constexpr int PI = PwPitch<CIN>::value, PO = PwPitch<COUT>::value;
  static_assert(TX * G * (TY / NY) == 256 && TY % NY == 0 && TX % 16 == 0, "tile / thread mapping");
  __shared__ __align__(16) uint8_t sD[NPX * PI];        // depthwise result, one row per pixel
  __shared__ __align__(16) uint8_t sO[8 * 16 * PO];     // per-warp output staging
  __shared__ __align__(16) uint16_t s_wpw[CIN * COUT];
  __shared__ __align__(16) float s_b[2 * CIN + 2 * COUT];   // bdw | sdw | bpw | spw
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < CIN * COUT / 8; i += 256) reinterpret_cast<uint4*>(s_wpw)[i] = reinterpret_cast<const uint4*>(wpw)[i];
  for (int i = tid; i < CIN; i += 256) {
    s_b[i] = bdw[i];
    s_b[CIN + i] = sdw[i];
  }
  for (int i = tid; i < COUT; i += 256) {
    s_b[2 * CIN + i] = bpw[i];
    s_b[2 * CIN + COUT + i] = spw[i];
  }
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
  // ---- phase 1: depthwise into shared memory ----
  {
    const int cg = tid % (TX * G), run = tid / (TX * G);
    const int g = cg % G, tx = cg / G;
    const int ox = ox0 + tx, ty0 = run * NY;
    uint4 wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = __ldg(wdw + t * G + g);
    float acc[NY][8];
    const bool colok = ox < Wo;
    if (colok && oy0 + ty0 < Ho) {
      const int xc = ox * STRIDE, rs = W * G;
      dw_run<STRIDE, NY>(x, b * H * rs + xc * G + g, rs, G, H, xc >= 1, xc + 1 < W, (oy0 + ty0) * STRIDE - 1, wv, acc);
    }
    __syncthreads();                                      // s_b is loaded
#pragma unroll
    for (int o = 0; o < NY; ++o) {
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (colok && oy0 + ty0 + o < Ho) out = finish8s(acc[o], s_b + g * 8, s_b + CIN + g * 8);
      *reinterpret_cast<uint4*>(sD + ((ty0 + o) * TX + tx) * PI + g * 16) = out;
    }
  }
  __syncthreads();
  // ---- phase 2: pointwise on the tensor cores ----
  PwWeights<CIN, COUT> wf;
  wf.load(s_wpw, lane);
  uint8_t* st = sO + warp * 16 * PO;
  for (int mt = warp; mt < NPX / 16; mt += 8) {
    const int p0 = mt * 16, ty = p0 / TX, tx0 = p0 % TX;
    const int oy = oy0 + ty;
    if (oy >= Ho || ox0 + tx0 >= Wo) continue;            // (warp-uniform)
    pw_tile16<CIN, COUT>(sD + p0 * PI, wf, s_b + 2 * CIN, s_b + 2 * CIN + COUT, st, lane);
    __syncwarp();
    uint4* yrow = y + (static_cast<size_t>(b) * Ho + oy) * Wo * (COUT / 8) + static_cast<size_t>(ox0 + tx0) * (COUT / 8);
    const int npx = min(16, Wo - (ox0 + tx0));
    for (int i = lane; i < npx * (COUT / 8); i += 32)
      yrow[i] = *reinterpret_cast<const uint4*>(st + (i / (COUT / 8)) * PO + (i % (COUT / 8)) * 16);
    __syncwarp();
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
  __shared__ __align__(16) uint16_t s_w[CIN * COUT];
  __shared__ __align__(16) float s_b[2 * COUT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < CIN * COUT / 8; i += 256) reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(w)[i];
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
  PwWeights<CIN, COUT> wf;
  wf.load(s_w, lane);
  uint8_t* st = sO + warp * 16 * PO;
  for (int mt = warp; mt < NPX / 16; mt += 8) {
    if (mt * 16 >= n) continue;
    pw_tile16<CIN, COUT>(sD + mt * 16 * PI, wf, s_b, s_b + COUT, st, lane);
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
  const long long total = static_cast<long long>(B) * Ho * ((Wo + 1) / 2);
  mnet_stem_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint2*>(x0), static_cast<const uint16_t*>(w), bias, slope, static_cast<uint4*>(y), B, H, W,
      vd_x0_pitch(W), Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

extern "C" int vd_mnet_dw3x3(const void* x, const void* w, const float* bias, const float* slope, void* y, int B, int H,
                             int W, int C, int stride, vd_stream_t stream) {
  VD_REQUIRE(x && w && bias && slope && y, "vd_mnet_dw3x3: null argument");
  VD_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
             "vd_mnet_dw3x3: C=%d must be a multiple of 8, stride 1 or 2", C);
  VD_REQUIRE(B <= 65535 && H <= 65535 && static_cast<long long>(B) * H * W * (C / 8) < (1ll << 31),
             "vd_mnet_dw3x3: batch / height / tensor size beyond the kernel's 32-bit indexing");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
             "vd_mnet_dw3x3: tensors must be 16-byte aligned");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(slope)) & 15) == 0,
             "vd_mnet_dw3x3: bias / slope must be 16-byte aligned");
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C8 = C / 8;
  int sh = -1;
  if ((C8 & (C8 - 1)) == 0) { sh = 0; while ((1 << sh) < C8) ++sh; }
  static const int ny_env = [] { const char* e = getenv("VD_MNET_DW_NY"); return e ? atoi(e) : 0; }();
  const int ny = stride == 2 ? (ny_env == 1 ? 1 : 2) : (ny_env > 0 ? ny_env : 4);
  const dim3 grid((Wo * C8 + 255) / 256, (Ho + ny - 1) / ny, B);
  const uint4* xp = static_cast<const uint4*>(x);
  const uint4* wp = static_cast<const uint4*>(w);
  uint4* yp = static_cast<uint4*>(y);
  if (stride == 1 && ny == 1) mnet_dw3x3_kernel<1, 1><<<grid, 256, 0, st>>>(xp, wp, bias, slope, yp, H, W, C8, sh, Ho, Wo);
  else if (stride == 1 && ny == 2) mnet_dw3x3_kernel<1, 2><<<grid, 256, 0, st>>>(xp, wp, bias, slope, yp, H, W, C8, sh, Ho, Wo);
  else if (stride == 1) mnet_dw3x3_kernel<1, 4><<<grid, 256, 0, st>>>(xp, wp, bias, slope, yp, H, W, C8, sh, Ho, Wo);
  else if (ny == 1) mnet_dw3x3_kernel<2, 1><<<grid, 256, 0, st>>>(xp, wp, bias, slope, yp, H, W, C8, sh, Ho, Wo);
  else mnet_dw3x3_kernel<2, 2><<<grid, 256, 0, st>>>(xp, wp, bias, slope, yp, H, W, C8, sh, Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

template <int CIN, int COUT>
static void launch_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                        const float* bpw, const float* spw, void* y, int B, int H, int W, int stride, int Ho, int Wo,
                        cudaStream_t st) {
  const uint4* xp = static_cast<const uint4*>(x);
  const uint4* wd = static_cast<const uint4*>(wdw);
  const uint16_t* wp = static_cast<const uint16_t*>(wpw);
  uint4* yp = static_cast<uint4*>(y);
  if (stride == 1) {
    using Cfg = DwpwCfg<CIN, COUT, 1>;
    const dim3 grid((Wo + Cfg::TX - 1) / Cfg::TX, (Ho + Cfg::TY - 1) / Cfg::TY, B);
    mnet_dwpw_kernel<CIN, COUT, 1><<<grid, 256, 0, st>>>(xp, wd, bdw, sdw, wp, bpw, spw, yp, H, W, Ho, Wo);
  } else {
    using Cfg = DwpwCfg<CIN, COUT, 2>;
    const dim3 grid((Wo + Cfg::TX - 1) / Cfg::TX, (Ho + Cfg::TY - 1) / Cfg::TY, B);
    mnet_dwpw_kernel<CIN, COUT, 2><<<grid, 256, 0, st>>>(xp, wd, bdw, sdw, wp, bpw, spw, yp, H, W, Ho, Wo);
  }
}

extern "C" int vd_mnet_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                            const float* bpw, const float* spw, void* y, int B, int H, int W, int Cin, int Cout,
                            int stride, vd_stream_t stream) {
  VD_REQUIRE(x && wdw && bdw && sdw && wpw && bpw && spw && y, "vd_mnet_dwpw: null argument");
  VD_REQUIRE(((reinterpret_cast<uintptr_t>(wdw) | reinterpret_cast<uintptr_t>(wpw) | reinterpret_cast<uintptr_t>(x) |
               reinterpret_cast<uintptr_t>(y)) & 15) == 0, "vd_mnet_dwpw: tensors and weights must be 16-byte aligned");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (stride == 1 || stride == 2), "vd_mnet_dwpw: bad extent / stride");
  VD_REQUIRE(Cin > 0 && static_cast<long long>(B) * H * W * (Cin / 8) < (1ll << 31), "vd_mnet_dwpw: tensor beyond the kernel's 32-bit indexing");
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  VD_REQUIRE((Ho + 7) / 8 <= 65535, "vd_mnet_dwpw: image too tall");
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
