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

// conv_bn(3, 8, stride 2), pad 1, on the ingest layout x0 = [B][H][Wp][4] (pixel x at column x + 3, zero margins):
// one thread = one output pixel = 3 rows x 24 contiguous bytes in, 16 bytes out.
__global__ void __launch_bounds__(256)
mnet_stem_kernel(const uint2* __restrict__ x0, const uint16_t* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ slope, uint4* __restrict__ y, int B, int H, int W, int Wp, int Ho, int Wo) {
  __shared__ float sw[36 * 8];
  for (int i = threadIdx.x; i < 36 * 8; i += blockDim.x) {
    const uint32_t u = w[i];
    sw[i] = bf16lo_to_f32(u);
  }
  __syncthreads();
  const long long total = static_cast<long long>(B) * Ho * Wo;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % Wo);
    const int ho = static_cast<int>((idx / Wo) % Ho);
    const int b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yi = 2 * ho - 1 + r;
      if (yi < 0 || yi >= H) continue;
      // input columns 2wo-1 .. 2wo+1 are padded columns 2wo+2 .. 2wo+4 (always inside the row)
      const uint2* row = x0 + (static_cast<size_t>(b) * H + yi) * Wp + (2 * wo + 2);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const uint2 px = __ldg(row + s);
        const float c0 = bf16lo_to_f32(px.x), c1 = bf16hi_to_f32(px.x), c2 = bf16lo_to_f32(px.y);
        const float* wp = sw + (r * 3 + s) * 32;
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] += c0 * wp[o] + c1 * wp[8 + o] + c2 * wp[16 + o];
      }
    }
    y[idx] = finish8(acc, bias, slope);
  }
}

// depthwise 3x3, pad 1, stride 1 or 2: one thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(256)
mnet_dw3x3_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ slope, uint4* __restrict__ y, int B, int H, int W, int C8, int stride,
                  int Ho, int Wo) {
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
      const int yi = ho * stride - 1 + r;
      if (yi < 0 || yi >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xi = wo * stride - 1 + s;
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
  const long long total = static_cast<long long>(B) * Ho * Wo;
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
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  mnet_dw3x3_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(w), bias, slope, static_cast<uint4*>(y), B, H, W, C / 8,
      stride, Ho, Wo);
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
