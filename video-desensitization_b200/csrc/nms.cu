// nms.cu — K4: batched greedy NMS + un-letterbox, all frames in three launches, no host sync.
//
// Replaces non_max_suppression -> torchvision.ops.batched_nms (single class) and
// retinaface_correct_boxes (reference detect_face/utils/utils_bbox.py:103-130, 12-43), which loop over
// frames in Python with one device->host sync per frame and ~10 more per frame for the correction.
//
//   1. sort   : one CTA per frame, bitonic sort in shared memory of (score, ~anchor) keys —
//               descending score, ties broken by lower anchor index (= torchvision's stable sort on
//               anchor-ordered input);
//   2. mask   : upper-triangular 64x64 blocks of the suppression matrix, IoU > thr with
//               torchvision's arithmetic (inter / (Sa + Sb - inter), areas without +1), one u64 word
//               per (row, column block);
//   3. reduce : one warp per frame walks the sorted list 64 candidates at a time, resolves the
//               diagonal block serially from shared memory, ORs the kept rows into the running
//               "removed" words with all lanes, and emits corrected rows in score order.
#include "common.cuh"

namespace {

constexpr int kSortThreads = 1024;

struct NmsWs {
  float4* sboxes;            // [B][cap] sorted boxes
  int* sslot;                // [B][cap] candidate slot of sorted position
  unsigned long long* mask;  // [B][cap][cap/64]
};

__host__ __device__ inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ NmsWs carve(void* ws, int B, int cap) {
  NmsWs w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  w.sboxes = reinterpret_cast<float4*>(p);
  p += align256(sizeof(float4) * static_cast<size_t>(B) * cap);
  w.sslot = reinterpret_cast<int*>(p);
  p += align256(sizeof(int) * static_cast<size_t>(B) * cap);
  w.mask = reinterpret_cast<unsigned long long*>(p);
  return w;
}

__global__ void __launch_bounds__(kSortThreads)
nms_sort_kernel(const float* __restrict__ cand, const int* __restrict__ cand_count, int cap, NmsWs ws) {
  extern __shared__ unsigned long long skey[];     // P keys, then P u16 slots
  const int b = blockIdx.x;
  const int n = min(cand_count[b], cap);
  int P = 64;
  while (P < n) P <<= 1;
  unsigned short* sval = reinterpret_cast<unsigned short*>(skey + P);
  const float* cb = cand + static_cast<size_t>(b) * cap * 16;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    unsigned long long k = 0ull;
    if (i < n) {
      const uint32_t sbits = __float_as_uint(cb[i * 16 + 4]);          // score > 0: uint order == float order
      const uint32_t anchor = __float_as_uint(cb[i * 16 + 15]);
      k = (static_cast<unsigned long long>(sbits) << 32) | (0xFFFFFFFFu - anchor);
    }
    skey[i] = k;
    sval[i] = static_cast<unsigned short>(i);
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = skey[lo], c = skey[hi];
        if ((a < c) == desc) {
          skey[lo] = c; skey[hi] = a;
          const unsigned short va = sval[lo];
          sval[lo] = sval[hi]; sval[hi] = va;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int slot = sval[i];
    ws.sslot[static_cast<size_t>(b) * cap + i] = slot;
    ws.sboxes[static_cast<size_t>(b) * cap + i] = *reinterpret_cast<const float4*>(cb + slot * 16);
  }
}

__device__ __forceinline__ bool iou_gt(const float4& a, const float4& c, float thr) {
  const float left = fmaxf(a.x, c.x), right = fminf(a.z, c.z);
  const float top = fmaxf(a.y, c.y), bottom = fminf(a.w, c.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(c.z, c.x), __fsub_rn(c.w, c.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(const int* __restrict__ cand_count, int cap, float thr, NmsWs ws) {
  __shared__ float4 cbox[64];
  const int b = blockIdx.y;
  const int n = min(cand_count[b], cap);
  const int nb = (n + 63) >> 6;
  const int words = cap >> 6;
  const int npairs = nb * (nb + 1) / 2;
  const float4* boxes = ws.sboxes + static_cast<size_t>(b) * cap;
  unsigned long long* mask = ws.mask + static_cast<size_t>(b) * cap * words;
  for (int pidx = blockIdx.x; pidx < npairs; pidx += gridDim.x) {
    // pair index -> (rb, cb) with rb <= cb, row-major over the upper triangle
    int rb = 0, rem = pidx;
    while (rem >= nb - rb) { rem -= nb - rb; ++rb; }
    const int cbk = rb + rem;
    const int cj = cbk * 64 + threadIdx.x;
    __syncthreads();
    if (cj < n) cbox[threadIdx.x] = boxes[cj];
    __syncthreads();
    const int ri = rb * 64 + threadIdx.x;
    if (ri < n) {
      const float4 me = boxes[ri];
      const int ncol = min(64, n - cbk * 64);
      unsigned long long bits = 0ull;
      const int start = (rb == cbk) ? threadIdx.x + 1 : 0;
      for (int j = start; j < ncol; ++j)
        if (iou_gt(me, cbox[j], thr)) bits |= 1ull << j;
      mask[static_cast<size_t>(ri) * words + cbk] = bits;
    }
  }
}

// Greedy pass over the suppression bitmask, one warp per frame.  Everything with a global-memory round trip is issued
// lane-parallel so the warp pays ONE latency per 64-candidate block instead of one per kept box: (1) the block's 64
// diagonal words, (2) the propagation of the kept rows into the following column words (lanes own kept ROWS and OR
// their words into shared memory — the first version walked the kept rows serially per lane), (3) the output rows,
// which are only listed here and written after the loop, two rows per step with independent loads.
__global__ void __launch_bounds__(32)
nms_reduce_kernel(const float* __restrict__ cand, const int* __restrict__ cand_count, int cap,
                  const float* __restrict__ unlb, int max_det, float* __restrict__ out_rows,
                  int* __restrict__ out_count, uint32_t* __restrict__ status, NmsWs ws) {
  extern __shared__ unsigned long long remv[];      // cap/64 words, 64 diagonal words, kept word, kept list
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const int n = min(cand_count[b], cap);
  const int nb = (n + 63) >> 6;
  const int words = cap >> 6;
  unsigned long long* diag = remv + words;
  unsigned long long* keptw = diag + 64;
  int* klist = reinterpret_cast<int*>(keptw + 1);   // sorted positions of the kept boxes, first min(max_det, cap)
  const unsigned long long* mask = ws.mask + static_cast<size_t>(b) * cap * words;
  const int* sslot = ws.sslot + static_cast<size_t>(b) * cap;
  const float* cb = cand + static_cast<size_t>(b) * cap * 16;
  const float offx = unlb[b * 4 + 0], offy = unlb[b * 4 + 1], scx = unlb[b * 4 + 2], scy = unlb[b * 4 + 3];
  float* orow = out_rows + static_cast<size_t>(b) * max_det * 16;

  for (int w = lane; w < words; w += 32) remv[w] = 0ull;
  __syncwarp();
  int nkept = 0;
  for (int blk = 0; blk < nb; ++blk) {
    const int base = blk * 64;
    const int cnt = min(64, n - base);
    for (int t = lane; t < 64; t += 32) diag[t] = (t < cnt) ? mask[static_cast<size_t>(base + t) * words + blk] : 0ull;
    __syncwarp();
    if (lane == 0) {
      unsigned long long cur = remv[blk], kept = 0ull;
      for (int t = 0; t < cnt; ++t) {
        if (!((cur >> t) & 1ull)) {
          kept |= 1ull << t;
          cur |= diag[t];
        }
      }
      *keptw = kept;
    }
    __syncwarp();
    const unsigned long long kept = *keptw;
    // propagate: lane owns rows t = lane, lane + 32 of the block; its loads over w are independent
    if (blk + 1 < nb) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
        if ((kept >> t) & 1ull) {
          const unsigned long long* mrow = mask + static_cast<size_t>(base + t) * words;
          for (int w = blk + 1; w < nb; ++w) {
            const unsigned long long m = mrow[w];
            if (m) atomicOr(&remv[w], m);
          }
        }
      }
    }
    // list the kept boxes in order (rank of bit t among the kept bits)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = lane + 32 * h;
      if ((kept >> t) & 1ull) {
        const int pos = nkept + __popcll(kept & ((1ull << t) - 1ull));
        if (pos < max_det) klist[pos] = base + t;
      }
    }
    nkept += __popcll(kept);
    __syncwarp();
  }
  // emit: half a warp per row, rows in kept order
  const int nout = min(nkept, max_det);
  const int col = lane & 15;
#pragma unroll 4
  for (int i = lane >> 4; i < nout; i += 2) {
    const int slot = sslot[klist[i]];
    float v = cb[slot * 16 + col];
    if (col != 4 && col != 15) {
      // utils_bbox.py:39-40: (coord - offset) * scale, x for even coordinate slots, y for odd
      const int c = (col < 4) ? col : (col - 5);
      const bool is_x = (c & 1) == 0;
      v = __fmul_rn(__fsub_rn(v, is_x ? offx : offy), is_x ? scx : scy);
    }
    orow[static_cast<size_t>(i) * 16 + col] = v;
  }
  if (lane == 0) {
    if (nkept > max_det) {
      atomicOr(status, VD_STAT_DET_OVERFLOW);
      nkept = max_det;
    }
    out_count[b] = nkept;
  }
}
}  // namespace

extern "C" size_t vd_nms_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  const size_t capz = static_cast<size_t>(cap);
  return align256(sizeof(float4) * B * capz) + align256(sizeof(int) * B * capz) +
         align256(sizeof(unsigned long long) * B * capz * (capz / 64)) + 256;
}

extern "C" int vd_nms_batched(const float* cand, const int32_t* cand_count, int B, int cap, float iou_thr,
                              const float* unletterbox, int max_det, float* out_rows, int32_t* out_count,
                              uint32_t* status, void* workspace, size_t workspace_bytes, vd_stream_t stream) {
  VD_REQUIRE(cand && cand_count && unletterbox && out_rows && out_count && status && workspace,
             "vd_nms_batched: null argument");
  VD_REQUIRE(B > 0 && max_det > 0, "vd_nms_batched: bad extent");
  VD_REQUIRE(cap >= 64 && cap <= VD_NMS_MAX_CAND && (cap & (cap - 1)) == 0,
             "vd_nms_batched: cap=%d must be a power of two in [64, %d]", cap, VD_NMS_MAX_CAND);
  VD_REQUIRE(workspace_bytes >= vd_nms_workspace_bytes(B, cap), "vd_nms_batched: workspace too small");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vd_nms_batched: workspace must be 256-B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NmsWs ws = carve(workspace, B, cap);

  const size_t sort_smem = static_cast<size_t>(cap) * (sizeof(unsigned long long) + sizeof(unsigned short));
  VD_SET_FUNC_ATTR_ONCE_PER_DEVICE(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VD_NMS_MAX_CAND * 10);
  nms_sort_kernel<<<B, kSortThreads, sort_smem, st>>>(cand, cand_count, cap, ws);
  VD_LAUNCH_CHECK();
  dim3 mgrid(64, B);
  nms_mask_kernel<<<mgrid, 64, 0, st>>>(cand_count, cap, iou_thr, ws);
  VD_LAUNCH_CHECK();
  const size_t red_smem = sizeof(unsigned long long) * (cap / 64 + 64 + 1) + sizeof(int) * static_cast<size_t>(max_det < cap ? max_det : cap);
  nms_reduce_kernel<<<B, 32, red_smem, st>>>(cand, cand_count, cap, unletterbox, max_det, out_rows, out_count, status,
                                             ws);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
