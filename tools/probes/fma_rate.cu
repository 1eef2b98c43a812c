// Issue rate of FFMA / FFMA2 / FHFMA.BF16 on one SM sub-partition (B200): warps of independent accumulator chains.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_rate fma_rate.cu ; run: ./fma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 4096
template <int MODE>
__global__ void k(float* out, uint32_t xa, uint32_t xb, long long* cyc) {
  float a[16];
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = (static_cast<unsigned long long>(__float_as_uint(a[2 * i + 1])) << 32) | __float_as_uint(a[2 * i]);
  const float fx = __uint_as_float(xa), fy = __uint_as_float(xb);
  const unsigned long long px = (static_cast<unsigned long long>(xa) << 32) | xa, py = (static_cast<unsigned long long>(xb) << 32) | xb;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < N_IT; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(fx), "f"(fy));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(px), "l"(py));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("{ .reg .b16 l, h, l2, h2;\n mov.b32 {l, h}, %1;\n mov.b32 {l2, h2}, %2;\n fma.rn.f32.bf16 %0, l, h2, %0;\n}" : "+f"(a[i]) : "r"(xa), "r"(xb));
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float(static_cast<uint32_t>(p[i])) + __uint_as_float(static_cast<uint32_t>(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
  const char* names[3] = {"FFMA (16 chains)", "FFMA2 (8 chains x 2)", "FHFMA.BF16 (16 chains)"};
  for (int warps = 1; warps <= 8; warps *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 128>>>(out, 0x3f800000u, 0x3f000000u, cyc);
        if (mode == 1) k<1><<<1, warps * 128>>>(out, 0x3f800000u, 0x3f000000u, cyc);
        if (mode == 2) k<2><<<1, warps * 128>>>(out, 0x3f803f80u, 0x3f003f00u, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double inst = (mode == 1 ? 8.0 : 16.0) * N_IT * warps;       // warp instructions per sub-partition
      printf("%-24s warps/SMSP %2d: %8lld cycles, %.3f warp-inst/clk/SMSP, %.1f FMA lanes/clk/SM\n", names[mode], warps, h,
             inst / h, inst / h * 32 * 4 * (mode == 1 ? 2 : 1));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
