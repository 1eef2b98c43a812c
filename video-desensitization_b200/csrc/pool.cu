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
}  // namespace

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
