// pool.cu — max-pool 3x3 stride 2 pad 1 on bf16 NHWC (torchvision resnet50 stem, reached through
// reference detect_face/retinaface.py:71-73,121).  HBM-bound: one thread = 8 channels (16 B) of one
// output pixel, nine 16-byte loads (L2-served overlaps), one 16-byte store.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8, int Ho, int Wo) {
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C8);
    size_t t = i / C8;
    const int xo = static_cast<int>(t % Wo);
    t /= Wo;
    const int yo = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    uint4 m = make_uint4(VD_ACT2_NEG_INF, VD_ACT2_NEG_INF, VD_ACT2_NEG_INF, VD_ACT2_NEG_INF);   // -inf pairs
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yi = yo * 2 - 1 + dy;
      if (yi < 0 || yi >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xi = xo * 2 - 1 + dx;
        if (xi < 0 || xi >= W) continue;
        const uint4 v = __ldg(x + ((static_cast<size_t>(b) * H + yi) * W + xi) * C8 + c);
        m.x = act2_max(m.x, v.x);
        m.y = act2_max(m.y, v.y);
        m.z = act2_max(m.z, v.z);
        m.w = act2_max(m.w, v.w);
      }
    }
    y[i] = m;
  }
}
// split-precision tensors: planes [hi | lo | hi] of C channels, value = hi + lo.  The window maximum is taken on
// the values; the (hi, lo) pair of the winning pixel is copied.
__global__ void __launch_bounds__(256)
maxpool3x3s2_split_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8, int Ho,
                          int Wo) {
  const size_t total = static_cast<size_t>(B) * Ho * Wo * C8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C8);
    size_t t = i / C8;
    const int xo = static_cast<int>(t % Wo);
    t /= Wo;
    const int yo = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    float best[8];
    uint32_t bh[4] = {0, 0, 0, 0}, bl[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) best[k] = -3.0e38f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yi = yo * 2 - 1 + dy;
      if (yi < 0 || yi >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xi = xo * 2 - 1 + dx;
        if (xi < 0 || xi >= W) continue;
        const uint4* px = x + ((static_cast<size_t>(b) * H + yi) * W + xi) * (3 * C8);
        const uint4 h = __ldg(px + c), l = __ldg(px + C8 + c);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float v0 = bf16lo_to_f32(hw[k]) + bf16lo_to_f32(lw[k]);
          const float v1 = bf16hi_to_f32(hw[k]) + bf16hi_to_f32(lw[k]);
          if (v0 > best[2 * k]) {
            best[2 * k] = v0;
            bh[k] = (bh[k] & 0xFFFF0000u) | (hw[k] & 0xFFFFu);
            bl[k] = (bl[k] & 0xFFFF0000u) | (lw[k] & 0xFFFFu);
          }
          if (v1 > best[2 * k + 1]) {
            best[2 * k + 1] = v1;
            bh[k] = (bh[k] & 0xFFFFu) | (hw[k] & 0xFFFF0000u);
            bl[k] = (bl[k] & 0xFFFFu) | (lw[k] & 0xFFFF0000u);
          }
        }
      }
    }
    uint4* po = y + ((static_cast<size_t>(b) * Ho + yo) * Wo + xo) * (3 * C8);
    const uint4 oh = make_uint4(bh[0], bh[1], bh[2], bh[3]), ol = make_uint4(bl[0], bl[1], bl[2], bl[3]);
    po[c] = oh;
    po[C8 + c] = ol;
    po[2 * C8 + c] = oh;
  }
}
}  // namespace

// max-pool on split-precision tensors ([hi | lo | hi] planes of C channels each; see VD_CONV_SPLIT)
extern "C" int vd_maxpool3x3s2_split(const void* x, void* y, int B, int H, int W, int C, vd_stream_t stream) {
  VD_REQUIRE(x && y, "vd_maxpool3x3s2_split: null tensor");
  VD_REQUIRE(C % 8 == 0 && B > 0 && H > 0 && W > 0, "vd_maxpool3x3s2_split: C must be a multiple of 8");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * (C / 8);
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(sms) * 16;
  if (blocks > cap) blocks = cap;
  maxpool3x3s2_split_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

extern "C" int vd_maxpool3x3s2(const void* x, void* y, int B, int H, int W, int C, vd_stream_t stream) {
  VD_REQUIRE(x && y, "vd_maxpool3x3s2: null tensor");
  VD_REQUIRE(C % 8 == 0 && B > 0 && H > 0 && W > 0, "vd_maxpool3x3s2: C must be a multiple of 8");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * (C / 8);
  int sms = vd_num_sms();
  if (sms <= 0) sms = 148;
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(sms) * 16;
  if (blocks > cap) blocks = cap;
  maxpool3x3s2_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), B, H, W, C / 8, Ho, Wo);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
