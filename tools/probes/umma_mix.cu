// umma_mix.cu — where do the cycles of conv3x3_rows' MMA stream go?  The isolated rate of a 128xNx16 tcgen05.mma is
// 48 / 64 clk for N = 64 / 128 (umma_rate.cu); inside the kernel the same stream runs ~30 % slower.  This probe issues the
// kernel's exact 48-instruction group pattern (t = 0: N = 64 into columns 0..63, t = 1, 2: N = 128, t = 3: N = 64 into
// columns 64..127; A descriptors shifted by s pixels of 128 B; two alternating accumulators) from one thread per CTA and adds
// the kernel's other activities one at a time:
//   bit 0  A descriptor starts shifted by s * 128 B (off the 1 KB swizzle atom), as the taps s = 1, 2 are
//   bit 1  eight warps read the OTHER accumulator with tcgen05.ld (the epilogue)
//   bit 2  a producer warp streams 16 KB bulk copies global -> shared (the input-row ring), 2 per group
//   bit 3  the issuing thread commits after every group and the readers wait for it (tfull / tempty hand-shake)
//   bit 4  the readers store what they loaded like the kernel's epilogue: 64 B per thread, lanes 128 B apart (4 x STG.128)
//   bit 5  same bytes, but lane-contiguous (a warp instruction writes 512 contiguous bytes)
//   bit 6  same bytes after a 4 x 4 transpose inside lane quads: a quad writes the 64 contiguous bytes of one pixel (8 pixels per
//          instruction), with the 12 shuffles the transpose costs
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I video-desensitization_b200/csrc -o umma_mix tools/probes/umma_mix.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "common.cuh"
void vd_set_error(const char*, ...) {}
int vd_num_sms() { return 148; }

constexpr int GROUPS = 96;
constexpr int ROW_BYTES = 136 * 128;
constexpr int RING = 7;
constexpr int WBLK = 8192;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(320, 1) mix(int flags, const uint8_t* __restrict__ gsrc, long long* out, float* sink, uint8_t* gdst) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + RING * ROW_BYTES;
  __shared__ uint64_t bar_done, tfull[2], tempty[2], ld_bar[RING];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid & 31;
  for (int i = tid; i < (RING * ROW_BYTES + 11 * WBLK) / 4; i += 320) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    for (int i = 0; i < RING; ++i) mbar_init(&ld_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tslot, 256); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tslot;
  const bool shifted = flags & 1, readers = flags & 2, loads = flags & 4, shake = flags & 8, st_kernel = flags & 16, st_coal = flags & 32, st_quad = flags & 64;
  if (warp == 0) {
    if (loads && lane == 0) {
      const uint8_t* src = gsrc + static_cast<size_t>(blockIdx.x) * (1 << 20);
      for (int i = 0; i < 2 * GROUPS; ++i) {
        const int slot = i % RING;
        if (i >= RING) mbar_wait(&ld_bar[slot], ((i / RING) - 1) & 1u);       // previous copy into this slot has landed
        mbar_expect_tx(&ld_bar[slot], 130 * 128);
        bulk_g2s(sA + slot * ROW_BYTES, src + (i % 60) * (130 * 128), 130 * 128, &ld_bar[slot]);
      }
      for (int i = 2 * GROUPS; i < 2 * GROUPS + RING; ++i) {                   // drain
        const int slot = i % RING;
        if (i >= RING) mbar_wait(&ld_bar[slot], ((i / RING) - 1) & 1u);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(64 >> 3) << 17) | (8u << 24);
      constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 3) << 17) | (8u << 24);
      constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t WB = WBLK >> 4;
      const uint32_t a0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      const long long t0 = clock64();
      for (int g = 0; g < GROUPS; ++g) {
        const uint32_t acc = g & 1u;
        if (shake) { mbar_wait(&tempty[acc], ((g >> 1) & 1u) ^ 1u); tcgen05_fence_after(); }
        const uint32_t d_lo = tm + acc * 128u, d_hi = d_lo + 64u;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t a_row = a0 + static_cast<uint32_t>(((2 * g + t) % RING) * (ROW_BYTES >> 4));
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t al = a_row + (shifted ? static_cast<uint32_t>(s * 8) : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (t == 0) {
                const uint32_t bl = b0 + (s == 0 ? 9u * WB : static_cast<uint32_t>(3 * s + 2) * WB) + 2 * k;
                if (s == 0 && k == 0) umma_bf16_words(d_lo, al, hi, bl, hi, idesc128, 0u);
                else umma_bf16_words(d_lo, al + 2 * k, hi, bl, hi, idesc64, 1u);
              } else if (t < 3) {
                umma_bf16_words(d_lo, al + 2 * k, hi, b0 + static_cast<uint32_t>(3 * s + 2 - t) * WB + 2 * k, hi, idesc128, 1u);
              } else {
                umma_bf16_words(d_hi, al + 2 * k, hi, b0 + static_cast<uint32_t>(3 * s) * WB + 2 * k, hi, idesc64, 1u);
              }
            }
          }
        }
        if (shake) umma_commit(&tfull[acc]);
      }
      umma_commit(&bar_done);
      mbar_wait(&bar_done, 0);
      out[blockIdx.x] = clock64() - t0;
    }
  } else if (readers || shake || st_kernel || st_coal || st_quad) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    float accu = 0.f;
    for (int g = 0; g < GROUPS; ++g) {
      const uint32_t acc = shake ? (g & 1u) : ((g & 1u) ^ 1u);
      if (shake) { mbar_wait(&tfull[acc], (g >> 1) & 1u); tcgen05_fence_after(); }
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(tm + (static_cast<uint32_t>(q * 32) << 16) + acc * 128u + h * 64u + half * 32u, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) accu += __uint_as_float(v[j]);
        if (st_kernel || st_coal || st_quad) {
          // one output row of the strip: 128 pixels x 128 B; this warp owns pixels q*32.. and channel half `half`
          uint8_t* rowp = gdst + (static_cast<size_t>(blockIdx.x) * 512 + ((2 * g + h) & 511)) * 16384;
          const uint4 a = make_uint4(v[0], v[1], v[2], v[3]), b = make_uint4(v[4], v[5], v[6], v[7]);
          if (st_kernel) {
            uint8_t* dst = rowp + (q * 32 + lane) * 128 + half * 64;
            reinterpret_cast<uint4*>(dst)[0] = a; reinterpret_cast<uint4*>(dst)[1] = b;
            reinterpret_cast<uint4*>(dst)[2] = a; reinterpret_cast<uint4*>(dst)[3] = b;
          } else if (st_quad) {
            uint4 c[4] = {a, b, a, b};
#pragma unroll
            for (int r = 1; r < 4; ++r) {              // 4 x 4 transpose of 16-byte chunks inside the quad: 3 rounds x 4 words
              const int src = (lane & ~3) | ((lane + r) & 3);
              uint4 snd = c[(lane - r) & 3 & 3];
              snd = c[((lane & 3) + 4 - r) & 3];
              uint4 rcv;
              rcv.x = __shfl_sync(0xffffffffu, snd.x, src); rcv.y = __shfl_sync(0xffffffffu, snd.y, src);
              rcv.z = __shfl_sync(0xffffffffu, snd.z, src); rcv.w = __shfl_sync(0xffffffffu, snd.w, src);
              c[(lane + r) & 3] = rcv;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint8_t* dst = rowp + (q * 32 + (lane & ~3) + j) * 128 + half * 64 + (lane & 3) * 16;
              *reinterpret_cast<uint4*>(dst) = c[j];
            }
          } else {
            uint8_t* dst = rowp + (q * 2 + half) * 2048 + lane * 16;
            reinterpret_cast<uint4*>(dst)[0] = a; reinterpret_cast<uint4*>(dst + 512)[0] = b;
            reinterpret_cast<uint4*>(dst + 1024)[0] = a; reinterpret_cast<uint4*>(dst + 1536)[0] = b;
          }
        }
      }
      if (shake) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      } else if (!shake) {
        __nanosleep(400);                  // free-running readers: roughly the kernel's two loads per group time
      }
    }
    if (accu == 12345.f) sink[tid] = accu;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 256);
}

int main() {
  long long* d;
  float* sink;
  uint8_t* g;
  uint8_t* gd;
  cudaMalloc(&d, 1024 * sizeof(long long));
  cudaMalloc(&sink, 4096);
  cudaMalloc(&g, 160u << 20);
  cudaMemset(g, 0x3c, 160u << 20);
  cudaMalloc(&gd, static_cast<size_t>(148) * 512 * 16384);
  const int smem = RING * ROW_BYTES + 11 * WBLK + 2048;
  cudaFuncSetAttribute(mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148;
  const char* names[] = {"pattern only", "+ shifted A", "+ tcgen05.ld readers", "+ shifted + readers", "+ bulk loads", "+ shifted + loads",
                         "+ readers + loads", "+ shifted + readers + loads", "+ hand-shake", "+ shifted + hand-shake", "", "",
                         "+ hand-shake + loads", "+ shifted + hand-shake + loads"};
  for (int flags : {0, 1, 2, 4, 7, 8, 13, 13 + 16, 13 + 32, 13 + 64, 8 + 16, 8 + 32, 8 + 64, 2 + 16, 2 + 32}) {
    std::vector<long long> h(grid);
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      mix<<<grid, 320, smem>>>(flags, g, d, sink, gd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("flags %d: %s\n", flags, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      best = std::min(best, double(h[grid / 2]) / GROUPS);
    }
    printf("flags %2d  %-32s%s : %7.0f clk per 48-MMA group (ideal 2688 = 24 x 48 + 24 x 64)\n", flags, names[flags & 15],
           (flags & 16) ? " + stores as in the kernel" : (flags & 32) ? " + lane-contiguous stores" : (flags & 64) ? " + quad-transposed stores" : "", best);
  }
  return 0;
}
