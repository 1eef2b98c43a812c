// umma_rate.cu — what does ONE tcgen05.mma (M = 128, K = 16, bf16, fp32 accumulate, cta_group::1) cost as a function of N,
// with the A operand read from shared memory (SS) or from tensor memory (TS)?  One thread per CTA issues REPS instructions
// back to back into one accumulator (operands cycle over several SWIZZLE_128B tiles: 4 K steps x NT tiles), commits, and
// the CTA measures clock64() around issue + completion.  Grid = one CTA per SM (all SMs loaded) or a single CTA.
// Build:  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I video-desensitization_b200/csrc -o umma_rate tools/probes/umma_rate.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "common.cuh"
void vd_set_error(const char*, ...) {}
int vd_num_sms() { return 148; }

constexpr int NT = 4;                       // operand tiles cycled through
constexpr int REPS = 4096;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int reps = REPS) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                        // NT tiles of 128 rows x 128 B
  uint8_t* sB = smem + NT * 16384;           // NT tiles of N rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (NT * 16384 + NT * N * 128) / 4; i += 128) {
    uint32_t v = 0x3c003c00u;                 // reps > REPS (the sustained run): random signs and mantissas, |x| in [0.5, 2)
    if (reps > REPS) {
      uint32_t x = (i + 1) * 2654435761u + blockIdx.x * 40503u;
      x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
      v = (x & 0x80ff80ffu) | 0x3f003f00u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tslot;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                               (static_cast<uint32_t>(128 >> 4) << 24);
    constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
    const unsigned long long g0 = gtimer();
    t0 = clock64();
    for (int i = 0; i < reps; i += 4 * NT) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t bl = b0 + static_cast<uint32_t>(t * (N * 128 >> 4) + 2 * k);
          if (TS) umma_ts(tm, tm + 256u + static_cast<uint32_t>((t * 4 + k) * 8), bl, hi, idesc, 1u);   // A: 8 columns per K step
          else umma_bf16_words(tm, a0 + static_cast<uint32_t>(t * (16384 >> 4) + 2 * k), hi, bl, hi, idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    out[512 + blockIdx.x] = static_cast<long long>(gtimer() - g0);
  }
  __syncthreads();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, bool TS>
static void run(const char* name, int grid, long long* d) {
  const int smem = NT * 16384 + NT * N * 128 + 2048;
  cudaFuncSetAttribute(rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(grid);
  double best = 1e30, med = 0;
  for (int rep = 0; rep < 3; ++rep) {
    rate<N, TS><<<grid, 128, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d grid=%d: %s\n", name, N, grid, cudaGetErrorString(e)); return; }
    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    med = double(h[grid / 2]) / REPS;
    best = std::min(best, med);
  }
  printf("%s  N=%3d  grid=%3d : %.1f clk per MMA (median CTA, best of 3; last %.1f)   math floor %d clk\n", name, N, grid, best, med, N / 2);
}

int main() {
  long long* d;
  cudaMalloc(&d, 1024 * sizeof(long long));
  for (int grid : {1, 148}) {
    run<64, false>("SS", grid, d);
    run<128, false>("SS", grid, d);
    run<192, false>("SS", grid, d);
    run<256, false>("SS", grid, d);
    run<64, true>("TS", grid, d);
    run<128, true>("TS", grid, d);
    run<256, true>("TS", grid, d);
  }
  // sustained: 2^19 N = 256 instructions per CTA on every SM (~35 ms) on random operands: SM cycles against wall-clock
  // nanoseconds = the clock the power cap allows under tensor load
  {
    const int smem = NT * 16384 + NT * 256 * 128 + 2048;
    std::vector<long long> h(1024);
    for (int rep = 0; rep < 3; ++rep) {
      rate<256, false><<<148, 128, smem>>>(d, 1 << 19);
      cudaDeviceSynchronize();
      cudaMemcpy(h.data(), d, 1024 * sizeof(long long), cudaMemcpyDeviceToHost);
      double clk = 0, ns = 0;
      for (int i = 0; i < 148; ++i) { clk += double(h[i]); ns += double(h[512 + i]); }
      printf("sustained SS N=256 on 148 SMs: %.1f clk per MMA, %.2f ms, effective SM clock %.0f MHz, %.0f TFLOP/s\n",
             clk / 148 / (1 << 19), ns / 148 / 1e6, clk / ns * 1e3, 148.0 * (1 << 19) * 2.0 * 128 * 256 * 16 / (ns / 148) / 1e3);
    }
  }
  return 0;
}
