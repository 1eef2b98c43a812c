// decode.cu — K3: PriorBox generation + box/landmark decode + confidence threshold + compaction.
//
// Replaces, in one pass over the head outputs:
//   Anchors.get_anchors          reference detect_face/utils/anchors.py:22-41   (priors recomputed from the
//                                anchor index with the same double -> float rounding; no prior table in HBM)
//   F.softmax(cls, -1)           reference detect_face/retinaface.py:147
//   decode / decode_landm        reference detect_face/utils/utils_bbox.py:49-79 (evaluation order kept:
//                                x1 = cx - w/2, x2 = w + x1; no FMA contraction)
//   mask = score >= conf_thres   reference detect_face/utils/utils_bbox.py:116   (>=)
// Only anchors that pass the threshold are decoded; survivors are appended to a per-frame candidate
// list with a warp-aggregated atomic.  HBM-bound: reads 1 sector (class logits) per cell + 128 B per
// surviving cell; writes 64 B per candidate.
#include "common.cuh"

namespace {

struct Levels {
  int fh[3], fw[3];
  int cell_off[4];     // prefix sum of cells per level
};

__device__ __forceinline__ void prior_of(int level, int i, int j, int a, int H, int W, float pr[4]) {
  const int step = 8 << level;
  const int min_size = (16 << (2 * level)) << a;    // [[16,32],[64,128],[256,512]]
  pr[0] = static_cast<float>(((j + 0.5) * step) / W);
  pr[1] = static_cast<float>(((i + 0.5) * step) / H);
  pr[2] = static_cast<float>(static_cast<double>(min_size) / W);
  pr[3] = static_cast<float>(static_cast<double>(min_size) / H);
}

__device__ __forceinline__ void decode_write(const float loc[4], float score, const float lm[10], const float pr[4],
                                             int anchor, float* __restrict__ dst) {
  const float v0 = 0.1f, v1 = 0.2f;
  const float cx = __fadd_rn(pr[0], __fmul_rn(__fmul_rn(loc[0], v0), pr[2]));
  const float cy = __fadd_rn(pr[1], __fmul_rn(__fmul_rn(loc[1], v0), pr[3]));
  const float w = __fmul_rn(pr[2], expf(__fmul_rn(loc[2], v1)));
  const float h = __fmul_rn(pr[3], expf(__fmul_rn(loc[3], v1)));
  const float x1 = __fsub_rn(cx, __fdiv_rn(w, 2.0f));
  const float y1 = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
  float o[16];
  o[0] = x1;
  o[1] = y1;
  o[2] = __fadd_rn(w, x1);
  o[3] = __fadd_rn(h, y1);
  o[4] = score;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    o[5 + 2 * k] = __fadd_rn(pr[0], __fmul_rn(__fmul_rn(lm[2 * k], v0), pr[2]));
    o[6 + 2 * k] = __fadd_rn(pr[1], __fmul_rn(__fmul_rn(lm[2 * k + 1], v0), pr[3]));
  }
  o[15] = __int_as_float(anchor);
  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
  for (int k = 0; k < 4; ++k) d4[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

__device__ __forceinline__ void locate(const Levels& L, int cell, int& level, int& i, int& j) {
  level = (cell >= L.cell_off[2]) ? 2 : (cell >= L.cell_off[1] ? 1 : 0);
  const int local = cell - L.cell_off[level];
  i = local / L.fw[level];
  j = local - i * L.fw[level];
}

// fused-head layout: heads[level][b][cell][32]
__global__ void __launch_bounds__(256)
decode_heads_kernel(const float* __restrict__ h0, const float* __restrict__ h1, const float* __restrict__ h2,
                    Levels L, int H, int W, float thr, int cap, float* __restrict__ cand,
                    int* __restrict__ cand_count, uint32_t* __restrict__ status) {
  const int b = blockIdx.y;
  const int total = L.cell_off[3] * 2;
  const int a_global = blockIdx.x * blockDim.x + threadIdx.x;      // anchor index inside the frame
  if (a_global >= total) return;
  const int cell = a_global >> 1, a = a_global & 1;
  int level, i, j;
  locate(L, cell, level, i, j);
  const float* hp = (level == 0 ? h0 : (level == 1 ? h1 : h2));
  const int cells_l = L.fh[level] * L.fw[level];
  const float* row = hp + (static_cast<size_t>(b) * cells_l + (cell - L.cell_off[level])) * 32;
  const float2 lg = __ldg(reinterpret_cast<const float2*>(row + 8 + 2 * a));   // (bg, face)
  const float m = fmaxf(lg.x, lg.y);
  const float e0 = expf(lg.x - m), e1 = expf(lg.y - m);
  const float score = __fdiv_rn(e1, __fadd_rn(e0, e1));
  if (!(score >= thr)) return;
  const int slot = atomicAdd(&cand_count[b], 1);
  if (slot >= cap) {
    atomicOr(status, VD_STAT_CAND_OVERFLOW);
    return;
  }
  float loc[4], lm[10], pr[4];
  const float4 l4 = __ldg(reinterpret_cast<const float4*>(row + 4 * a));
  loc[0] = l4.x; loc[1] = l4.y; loc[2] = l4.z; loc[3] = l4.w;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(row + 12 + 10 * a + 2 * k));
    lm[2 * k] = t.x;
    lm[2 * k + 1] = t.y;
  }
  prior_of(level, i, j, a, H, W, pr);
  decode_write(loc, score, lm, pr, a_global, cand + (static_cast<size_t>(b) * cap + slot) * 16);
}

// reference layout: loc [B][A][4], conf [B][A][2] (softmaxed), landms [B][A][10]
__global__ void __launch_bounds__(256)
decode_lcl_kernel(const float* __restrict__ loc_, const float* __restrict__ conf_, const float* __restrict__ lm_,
                  Levels L, int H, int W, float thr, int cap, float* __restrict__ cand, int* __restrict__ cand_count,
                  uint32_t* __restrict__ status) {
  const int b = blockIdx.y;
  const int total = L.cell_off[3] * 2;
  const int a_global = blockIdx.x * blockDim.x + threadIdx.x;
  if (a_global >= total) return;
  const size_t idx = static_cast<size_t>(b) * total + a_global;
  const float score = __ldg(conf_ + idx * 2 + 1);
  if (!(score >= thr)) return;
  const int slot = atomicAdd(&cand_count[b], 1);
  if (slot >= cap) {
    atomicOr(status, VD_STAT_CAND_OVERFLOW);
    return;
  }
  const int cell = a_global >> 1, a = a_global & 1;
  int level, i, j;
  locate(L, cell, level, i, j);
  float loc[4], lm[10], pr[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) loc[k] = __ldg(loc_ + idx * 4 + k);
#pragma unroll
  for (int k = 0; k < 10; ++k) lm[k] = __ldg(lm_ + idx * 10 + k);
  prior_of(level, i, j, a, H, W, pr);
  decode_write(loc, score, lm, pr, a_global, cand + (static_cast<size_t>(b) * cap + slot) * 16);
}

Levels make_levels(int H, int W) {
  Levels L;
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    int h = H, w = W;
    for (int s = 8 << l; s > 1; s >>= 1) { h = (h + 1) / 2; w = (w + 1) / 2; }
    L.fh[l] = h; L.fw[l] = w;
    L.cell_off[l] = off;
    off += h * w;
  }
  L.cell_off[3] = off;
  return L;
}
}  // namespace

extern "C" int vd_decode_threshold(const float* h0, const float* h1, const float* h2, int B, int H, int W,
                                   float conf_thr, int cap, float* cand, int32_t* cand_count, uint32_t* status,
                                   vd_stream_t stream) {
  VD_REQUIRE(h0 && h1 && h2 && cand && cand_count && status, "vd_decode_threshold: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && cap > 0, "vd_decode_threshold: bad extent");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Levels L = make_levels(H, W);
  VD_CHECK_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * B, st));
  dim3 grid(vd_cdiv(L.cell_off[3] * 2, 256), B);
  decode_heads_kernel<<<grid, 256, 0, st>>>(h0, h1, h2, L, H, W, conf_thr, cap, cand, cand_count, status);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

extern "C" int vd_decode_threshold_lcl(const float* loc, const float* conf, const float* landms, int B, int H, int W,
                                       float conf_thr, int cap, float* cand, int32_t* cand_count, uint32_t* status,
                                       vd_stream_t stream) {
  VD_REQUIRE(loc && conf && landms && cand && cand_count && status, "vd_decode_threshold_lcl: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && cap > 0, "vd_decode_threshold_lcl: bad extent");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Levels L = make_levels(H, W);
  VD_CHECK_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int32_t) * B, st));
  dim3 grid(vd_cdiv(L.cell_off[3] * 2, 256), B);
  decode_lcl_kernel<<<grid, 256, 0, st>>>(loc, conf, landms, L, H, W, conf_thr, cap, cand, cand_count, status);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
