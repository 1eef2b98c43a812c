// probe_tma.cu — standalone probe of the TMA behaviours conv_tcgen05.cu relies on (run on the B200):
//   (1) out-of-bounds zero fill with negative start coordinates,
//   (2) elementStrides semantics (which source pixels a strided box fetches),
//   (3) overlapping global strides (conv1's sliding-window tensor map).
// Prints, for every 128-byte smem row, the (x,y) pixel id it received.  Not part of libvdesens.so.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

void vd_set_error(const char*, ...) {}
int vd_num_sms() { return 148; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int bytes,
                             uint16_t* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xEEEE;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes);
    tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  // de-swizzle (SWIZZLE_128B: 16-byte chunk index ^= row % 8) while copying out
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int row = i / 64, e = i % 64;
    const int chunk = e / 8, within = e % 8;
    const int pchunk = chunk ^ (row & 7);
    out[i] = reinterpret_cast<uint16_t*>(smem)[row * 64 + pchunk * 8 + within];
  }
}

static int run(const char* name, EncodeTiledFn enc, void* base, int rank4, const cuuint64_t* gdim, const cuuint64_t* gstr,
               const cuuint32_t* box, const cuuint32_t* estr, int c0, int c1, int c2, int c3, int rows, int W) {
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("== %s: encode CUresult=%d\n", name, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  uint16_t* dout;
  cudaMalloc(&dout, 128 * 64 * 2);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
  probe_kernel<<<1, 128, 18000>>>(tm, c0, c1, c2, c3, rows * 128, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("   kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<uint16_t> h(128 * 64);
  cudaMemcpy(h.data(), dout, h.size() * 2, cudaMemcpyDeviceToHost);
  for (int row = 0; row < rows; ++row) {
    const uint16_t v = h[row * 64];
    bool uniform = true;
    for (int e2 = 0; e2 < 64; ++e2) uniform &= (h[row * 64 + e2] == (v ? v + (e2 >= 32 ? 0 : 0) : 0));
    if (v == 0) printf("   row %3d: ZERO%s\n", row, uniform ? "" : " (mixed!)");
    else printf("   row %3d: pixel id %5d -> y=%d x=%d%s\n", row, v - 1, (v - 1) / W, (v - 1) % W, uniform ? "" : " (mixed!)");
  }
  cudaFree(dout);
  return 0;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  // tensor [N=2][H=7][W=10][C=64] u16, every channel of pixel (y,x) holds y*W+x+1
  const int N = 2, H = 7, W = 10, C = 64;
  std::vector<uint16_t> h(static_cast<size_t>(N) * H * W * C);
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < C; ++c) h[((static_cast<size_t>(n) * H + y) * W + x) * C + c] = static_cast<uint16_t>(y * W + x + 1);
  uint16_t* d;
  cudaMalloc(&d, h.size() * 2 + 4096);
  cudaMemset(d, 0, h.size() * 2 + 4096);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  {
    cuuint64_t gdim[4] = {C, W, H, N};
    cuuint64_t gstr[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 4, 3, 1}, es[4] = {1, 1, 1, 1};
    run("stride1 box 4x3 at (x=-1,y=-1)", enc, d, 4, gdim, gstr, box, es, 0, -1, -1, 0, 12, W);
    run("stride1 box 4x3 at (x=8,y=5) (right/bottom OOB)", enc, d, 4, gdim, gstr, box, es, 0, 8, 5, 0, 12, W);
    cuuint32_t box2[4] = {64, 8, 6, 1}, es2[4] = {1, 2, 2, 1};
    run("elementStrides 2, box 8x6 (4x3 loaded) at (x=-1,y=-1)", enc, d, 4, gdim, gstr, box2, es2, 0, -1, -1, 0, 12, W);
    run("elementStrides 2, box 8x6 at (x=0,y=0)", enc, d, 4, gdim, gstr, box2, es2, 0, 0, 0, 0, 12, W);
    run("elementStrides 2, box 8x6 at (x=1,y=2)", enc, d, 4, gdim, gstr, box2, es2, 0, 1, 2, 0, 12, W);
    cuuint32_t box3[4] = {64, 4, 2, 2};
    run("stride1 box 4x2x2 (two images) at (x=0,y=0,n=0)", enc, d, 4, gdim, gstr, box3, es, 0, 0, 0, 0, 16, W);
  }
  {
    // overlapping strides: treat rows as flat [W*C] u16; window w starts at element w*16 (32 B) and is 64 long
    cuuint64_t gdim[4] = {64, 5, H, N};
    cuuint64_t gstr[3] = {32, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, 4, 4, 1}, es[4] = {1, 1, 2, 1};
    printf("-- overlapping: expect row = window; first element id: window w starts at element 16*w of the row -> x = 0 for w<4\n");
    run("overlapping stride 32B, box 4 windows x (4 rows, estride 2) at (w=0,y=-1)", enc, d, 4, gdim, gstr, box, es, 0, 0, -1,
        0, 8, W);
  }
  return 0;
}
