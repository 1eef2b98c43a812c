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

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo_to_f32(v.x); f[1] = bf16hi_to_f32(v.x);
  f[2] = bf16lo_to_f32(v.y); f[3] = bf16hi_to_f32(v.y);
  f[4] = bf16lo_to_f32(v.z); f[5] = bf16hi_to_f32(v.z);
  f[6] = bf16lo_to_f32(v.w); f[7] = bf16hi_to_f32(v.w);
}

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
// the 216 fp32 weights sit in shared memory as float4 and every weight load (same address for the whole warp) feeds
// both pixels.
__global__ void __launch_bounds__(256)
mnet_stem_kernel(const uint2* __restrict__ x0, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ slope, uint4* __restrict__ y, int B, int H, int W, int Wp, int Ho, int Wo) {
  __shared__ float4 sw[36 * 2];                       // [(r, s, c4)][8 cout] -> two float4 per (tap, channel)
  float* swf = reinterpret_cast<float*>(sw);
  for (int i = threadIdx.x; i < 36 * 8; i += blockDim.x) {
    const uint32_t u = w[i];
    swf[i] = bf16lo_to_f32(u);
  }
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
        const float c00 = bf16lo_to_f32(px[sx][0]), c01 = bf16hi_to_f32(px[sx][0]), c02 = bf16lo_to_f32(px[sx][1]);
        const float c10 = bf16lo_to_f32(px[sx + 2][0]), c11 = bf16hi_to_f32(px[sx + 2][0]),
                    c12 = bf16lo_to_f32(px[sx + 2][1]);
        const float4* wq = sw + (r * 3 + sx) * 8;        // 4 channels x 2 float4
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const float4 w0 = wq[hlf], w1 = wq[2 + hlf], w2 = wq[4 + hlf];
          float* o0 = a0 + 4 * hlf;
          float* o1 = a1 + 4 * hlf;
          o0[0] += c00 * w0.x + c01 * w1.x + c02 * w2.x; o0[1] += c00 * w0.y + c01 * w1.y + c02 * w2.y;
          o0[2] += c00 * w0.z + c01 * w1.z + c02 * w2.z; o0[3] += c00 * w0.w + c01 * w1.w + c02 * w2.w;
          o1[0] += c10 * w0.x + c11 * w1.x + c12 * w2.x; o1[1] += c10 * w0.y + c11 * w1.y + c12 * w2.y;
          o1[2] += c10 * w0.z + c11 * w1.z + c12 * w2.z; o1[3] += c10 * w0.w + c11 * w1.w + c12 * w2.w;
        }
      }
    }
    uint4* out = y + (static_cast<size_t>(b) * Ho + ho) * Wo + wo;
    out[0] = finish8(a0, bias, slope);
    if (wo + 1 < Wo) out[1] = finish8(a1, bias, slope);
  }
}

// depthwise 3x3, pad 1, stride 1 or 2: one thread = one output pixel x 8 channels.  A variant in which a thread walks a
// run of 4 / 2 vertically adjacent outputs with the nine weight vectors packed in registers (18 loads per 4 outputs
// instead of 36, 115 registers instead of 76) was tried: over the nine C >= 64 layers at B = 16, 1080p it measured
// 0.459 ms against 0.433 and 0.456 ms for THIS form on two other boxes — inside the box-to-box spread of the same
// kernel, faster on the stride-1 128-channel layers and slower on the stride-2 / 64- / 256-channel ones.  The two
// were never timed on one box, so no ranking is claimed; the simpler form stays.
template <int STRIDE>
__global__ void __launch_bounds__(256)
mnet_dw3x3_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ slope, uint4* __restrict__ y, int B, int H, int W, int C8, int Ho, int Wo) {
  const long long total = static_cast<long long>(B) * Ho * Wo * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % C8);
    long long p = idx / C8;
    const int wo = static_cast<int>(p % Wo);
    p /= Wo;
    const int ho = static_cast<int>(p % Ho);
    const int b = static_cast<int>(p / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yi = ho * STRIDE - 1 + r;
      if (yi < 0 || yi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xi = wo * STRIDE - 1 + s;
        if (xi < 0 || xi >= W) continue;
        float xv[8], wv[8];
        unpack8(__ldg(x + ((static_cast<size_t>(b) * H + yi) * W + xi) * C8 + g), xv);
        unpack8(__ldg(w + (r * 3 + s) * C8 + g), wv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += xv[i] * wv[i];
      }
    }
    y[idx] = finish8(acc, bias + g * 8, slope + g * 8);
  }
}

// conv_dw block = depthwise 3x3 (+BN+slope) followed by pointwise 1x1 (+BN+slope) in ONE kernel for Cin <= 32: the
// depthwise result of a 32 x 8 pixel tile goes to shared memory (rounded to the activation type exactly where the
// two-kernel form stores it) and never to HBM.  Phase 2: one thread = 4 adjacent pixels x 8 output channels, so a
// weight vector read from shared memory feeds 32 FMAs.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
mnet_dwpw_kernel(const uint4* __restrict__ x, const uint4* __restrict__ wdw, const float* __restrict__ bdw,
                 const float* __restrict__ sdw, const uint16_t* __restrict__ wpw, const float* __restrict__ bpw,
                 const float* __restrict__ spw, uint4* __restrict__ y, int H, int W, int stride, int Ho, int Wo) {
  constexpr int G = CIN / 8, C8 = COUT / 8;
  constexpr int TX = 32, TY = 8, NPX = TX * TY;
  constexpr int DROW = G + 1;                           // uint4 per pixel in sD (+1: spreads the quads over banks)
  __shared__ float4 s_wpw4[CIN * COUT / 4];
  __shared__ uint4 sD[NPX * DROW];
  __shared__ uint4 s_wdw[9 * G];
  __shared__ float s_b[2 * CIN + 2 * COUT];             // bdw | sdw | bpw | spw
  float* s_wpw = reinterpret_cast<float*>(s_wpw4);
  for (int i = threadIdx.x; i < CIN * COUT; i += 256) s_wpw[i] = bf16lo_to_f32(wpw[i]);
  for (int i = threadIdx.x; i < 9 * G; i += 256) s_wdw[i] = wdw[i];
  for (int i = threadIdx.x; i < CIN; i += 256) {
    s_b[i] = bdw[i];
    s_b[CIN + i] = sdw[i];
  }
  for (int i = threadIdx.x; i < COUT; i += 256) {
    s_b[2 * CIN + i] = bpw[i];
    s_b[2 * CIN + COUT + i] = spw[i];
  }
  __syncthreads();
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * TX, oy0 = blockIdx.y * TY;
  // ---- phase 1: depthwise into shared memory ----
  for (int i = threadIdx.x; i < NPX * G; i += 256) {
    const int g = i % G, px = i / G;
    const int ox = ox0 + (px & (TX - 1)), oy = oy0 + (px / TX);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint4 out = make_uint4(0, 0, 0, 0);
    if (ox < Wo && oy < Ho) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int yi = oy * stride - 1 + r;
        if (yi < 0 || yi >= H) continue;
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          const int xi = ox * stride - 1 + sx;
          if (xi < 0 || xi >= W) continue;
          float xv[8], wv[8];
          unpack8(__ldg(x + ((static_cast<size_t>(b) * H + yi) * W + xi) * G + g), xv);
          unpack8(s_wdw[(r * 3 + sx) * G + g], wv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += xv[k] * wv[k];
        }
      }
      out = finish8s(acc, s_b + g * 8, s_b + CIN + g * 8);
    }
    sD[px * DROW + g] = out;
  }
  __syncthreads();
  // ---- phase 2: pointwise from shared memory ----
  for (int it = threadIdx.x; it < (NPX / 4) * C8; it += 256) {
    const int c = it % C8, q = it / C8;
    float acc[4][8];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[pp][k] = 0.f;
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      float d[4][8];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) unpack8(sD[(4 * q + pp) * DROW + g], d[pp]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(s_wpw + (g * 8 + j) * COUT + c * 8);
        const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const float dv = d[pp][j];
          acc[pp][0] += dv * w0.x; acc[pp][1] += dv * w0.y; acc[pp][2] += dv * w0.z; acc[pp][3] += dv * w0.w;
          acc[pp][4] += dv * w1.x; acc[pp][5] += dv * w1.y; acc[pp][6] += dv * w1.z; acc[pp][7] += dv * w1.w;
        }
      }
    }
    const int px0 = 4 * q;
    const int oy = oy0 + px0 / TX;
    if (oy < Ho) {
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const int ox = ox0 + (px0 & (TX - 1)) + pp;
        if (ox < Wo)
          y[((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * C8 + c] =
              finish8s(acc[pp], s_b + 2 * CIN + c * 8, s_b + 2 * CIN + COUT + c * 8);
      }
    }
  }
}

// pointwise 1x1, Cin * Cout <= 4096: weights [Cin][Cout] in shared memory as fp32, one thread = one pixel x 8 outputs.
__global__ void __launch_bounds__(256)
mnet_pw_kernel(const uint4* __restrict__ x, const uint16_t* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ slope, uint4* __restrict__ y, long long npix, int Cin, int Cout) {
  __shared__ float sw[4096];
  for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) {
    const uint32_t u = w[i];
    sw[i] = bf16lo_to_f32(u);
  }
  __syncthreads();
  const int Co8 = Cout >> 3, Ci8 = Cin >> 3;
  const long long total = npix * Co8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % Co8);
    const long long p = idx / Co8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c8 = 0; c8 < Ci8; ++c8) {
      float xv[8];
      unpack8(__ldg(x + p * Ci8 + c8), xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(sw + (c8 * 8 + j) * Cout + g * 8);
        const float4 w0 = wr[0], w1 = wr[1];
        acc[0] += xv[j] * w0.x; acc[1] += xv[j] * w0.y; acc[2] += xv[j] * w0.z; acc[3] += xv[j] * w0.w;
        acc[4] += xv[j] * w1.x; acc[5] += xv[j] * w1.y; acc[6] += xv[j] * w1.z; acc[7] += xv[j] * w1.w;
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
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  if (stride == 1)
    mnet_dw3x3_kernel<1><<<grid_for(total), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(w), bias,
                                                         slope, static_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
  else
    mnet_dw3x3_kernel<2><<<grid_for(total), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(w), bias,
                                                         slope, static_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

template <int CIN, int COUT>
static void launch_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                        const float* bpw, const float* spw, void* y, int B, int H, int W, int stride, int Ho, int Wo,
                        cudaStream_t st) {
  const dim3 grid((Wo + 31) / 32, (Ho + 7) / 8, B);
  mnet_dwpw_kernel<CIN, COUT><<<grid, 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(wdw), bdw, sdw,
                                                   static_cast<const uint16_t*>(wpw), bpw, spw, static_cast<uint4*>(y), H,
                                                   W, stride, Ho, Wo);
}

extern "C" int vd_mnet_dwpw(const void* x, const void* wdw, const float* bdw, const float* sdw, const void* wpw,
                            const float* bpw, const float* spw, void* y, int B, int H, int W, int Cin, int Cout,
                            int stride, vd_stream_t stream) {
  VD_REQUIRE(x && wdw && bdw && sdw && wpw && bpw && spw && y, "vd_mnet_dwpw: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (stride == 1 || stride == 2), "vd_mnet_dwpw: bad extent / stride");
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

extern "C" int vd_mnet_pw(const void* x, const void* w, const float* bias, const float* slope, void* y, long long npix,
                          int Cin, int Cout, vd_stream_t stream) {
  VD_REQUIRE(x && w && bias && slope && y, "vd_mnet_pw: null argument");
  VD_REQUIRE(npix > 0 && Cin > 0 && Cout > 0 && Cin % 8 == 0 && Cout % 8 == 0 && Cin * Cout <= 4096,
             "vd_mnet_pw: Cin=%d Cout=%d must be multiples of 8 with Cin*Cout <= 4096", Cin, Cout);
  mnet_pw_kernel<<<grid_for(npix * (Cout / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint16_t*>(w), bias, slope, static_cast<uint4*>(y), npix, Cin, Cout);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
