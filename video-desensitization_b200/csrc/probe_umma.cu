// probe_umma.cu — does a tcgen05.mma A operand read correctly from a SWIZZLE_128B tile whose start address is
// shifted by whole 128-byte rows (conv halo patch reuse)?  Tries start shifts, group strides (SBO) and the
// descriptor base_offset field, and prints which patch row every M row actually received.  Not part of libvdesens.
#include <cstdio>
#include <vector>
#include "common.cuh"
void vd_set_error(const char*, ...) {}
int vd_num_sms() { return 148; }

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(int shift_rows, int sbo, int base_off, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(smem);            // 320 rows x 64
  __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(smem + 320 * 128);   // B: 64 x 64 identity
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid / 32;
  // patch row n, column c: value n (c even) or c (c odd); stored TMA-style: chunk ^= (row address >> 7) & 7
  for (int i = tid; i < 320 * 64; i += 128) {
    const int n = i / 64, c = i % 64;
    const int chunk = c / 8, within = c % 8;
    const int pch = chunk ^ (n & 7);
    patch[n * 64 + pch * 8 + within] = __float2bfloat16((c & 1) ? float(c) : float(n % 250));
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i / 64, c = i % 64;
    const int pch = (c / 8) ^ (n & 7);
    bt[n * 64 + pch * 8 + (c % 8)] = __float2bfloat16(n == c ? 1.f : 0.f);
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 64); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tslot;
  if (tid == 0) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(64 >> 3) << 17) |
                               (static_cast<uint32_t>(128 >> 4) << 24);
    const uint64_t ad = mk_desc(smem_u32(smem) + shift_rows * 128, sbo, base_off);
    const uint64_t bd = mk_desc(smem_u32(bt), 1024, 0);
    for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bd + 2 * k, idesc, k != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (tid & 31)) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int cfgs[][3] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {2, 1024, 2}, {2, 1024, 0}, {1, 2048, 0}, {1, 2048, 1},
                         {0, 1280, 0}, {1, 1280, 0}, {1, 1280, 1}, {17, 2048, 1}, {17, 2048, 0}, {3, 1152, 0}, {3, 1152, 3}};
  for (auto& c : cfgs) {
    probe<<<1, 128, 60 * 1024>>>(c[0], c[1], c[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(128 * 64);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int ok = 0, colok = 0;
    for (int m = 0; m < 128; ++m) {
      const int expect = c[0] + (m / 8) * (c[1] / 128) + (m % 8);
      bool rowok = true, cok = true;
      for (int cc = 0; cc < 64; ++cc) {
        if (cc & 1) cok &= (h[m * 64 + cc] == float(cc));
        else rowok &= (h[m * 64 + cc] == float(expect % 250));
      }
      ok += rowok; colok += cok;
    }
    printf("shift %2d SBO %4d base_off %d : %s rows ok %3d/128 cols ok %3d/128 | m0..3 -> %g %g %g %g ; m8 -> %g ; odd cols m1: %g %g %g\n",
           c[0], c[1], c[2], cudaGetErrorString(e), ok, colok, h[0], h[64], h[128], h[192], h[8 * 64], h[64 + 1], h[64 + 3],
           h[64 + 9]);
  }
  return 0;
}
