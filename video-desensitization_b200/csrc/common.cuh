// common.cuh — shared host/device helpers for libvdesens (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>

#include "../../include/vdesens.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvdesens is written for sm_100a only"
#endif

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void vd_set_error(const char* fmt, ...);

#define VD_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      vd_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VD_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define VD_REQUIRE(cond, ...)                                                            \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      vd_set_error(__VA_ARGS__);                                                         \
      return VD_ERR_INVALID;                                                             \
    }                                                                                    \
  } while (0)

// Function attributes (opt-in dynamic shared memory) belong to a DEVICE: a process that drives several GPUs — one
// Retinaface per GPU from one Python process, like the reference's single-process DataParallel set-up — must set them on
// each.  One bit per device ordinal, relaxed atomics (setting an attribute twice is harmless).
struct VdPerDeviceOnce {
  unsigned long long mask = 0ull;
  // returns the current device ordinal if the attribute still has to be set there, -1 if already done, -2 on error
  int pending() const {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -2;
    if (dev < 0 || dev >= 64) return dev < 0 ? -2 : dev;          // ordinals >= 64: set every time
    return (__atomic_load_n(&mask, __ATOMIC_RELAXED) >> dev) & 1ull ? -1 : dev;
  }
  void done(int dev) {
    if (dev >= 0 && dev < 64) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELAXED);
  }
};

#define VD_SET_FUNC_ATTR_ONCE_PER_DEVICE(kernel, attr, value)                            \
  do {                                                                                   \
    static VdPerDeviceOnce _once;                                                        \
    const int _dev = _once.pending();                                                    \
    VD_REQUIRE(_dev != -2, "cudaGetDevice failed");                                      \
    if (_dev >= 0) {                                                                     \
      VD_CHECK_CUDA(cudaFuncSetAttribute(kernel, attr, value));                          \
      _once.done(_dev);                                                                  \
    }                                                                                    \
  } while (0)

#define VD_LAUNCH_CHECK()                                                                \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      vd_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VD_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

static inline int vd_cdiv(int a, int b) { return (a + b - 1) / b; }
int vd_num_sms();

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a trap (-> CUDA error on the host), never as a
// hung GPU.  try_wait suspends the thread in hardware for up to a few microseconds per poll, so 2^24 polls is
// tens of seconds — far beyond any legitimate wait in these kernels (< 1 ms).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// two IEEE fp32 additions in one issue slot (Blackwell FADD2): (a0, a1) += (b0, b1), same results as two add.rn.f32
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
#ifdef VD_NO_FADD2
  a0 = __fadd_rn(a0, b0);
  a1 = __fadd_rn(a1, b1);
  return;
#endif
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}

// TMA store smem -> global (bulk async group), clipped at the tensor bounds
__device__ __forceinline__ void tma_store_4d(const void* desc, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- explicit shared-space 128-bit accesses (a generic LD/ST on a shared pointer pays the address-space check) ----
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (lo, hi) 32-bit words (cheap to advance on the uniform datapath)
__device__ __forceinline__ void umma_bf16_words(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns of fp32: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- misc ----
// Activation / weight storage type: bf16 (default, BASELINE's dtype) or, when the library is built with
// -DVD_ACT_FP16 (libvdesens_fp16.so), IEEE half — same tensor-core rate, 3 more mantissa bits.
#ifdef VD_ACT_FP16
#define VD_ACT_NAME "fp16"
#define VD_ACT_IS_FP16 1
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);               // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}
// conversion with the ReLU folded in (cvt ... .relu: negative results become +0): same bits as fmaxf(x, 0) followed by the
// plain conversion for every finite input, one instruction per pair instead of three
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFFu))); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t v) { return __half2float(__ushort_as_half(static_cast<unsigned short>(v >> 16))); }
__device__ __forceinline__ uint32_t act2_max(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
#define VD_ACT2_NEG_INF 0xFC00FC00u
#else
#define VD_ACT_NAME "bf16"
#define VD_ACT_IS_FP16 0
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t act2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
#define VD_ACT2_NEG_INF 0xFF80FF80u
#endif
// c + a * b with a, b one half each of two packed activation-type words and c fp32: ONE instruction (FHFMA with .H0 / .H1
// operand selectors), no unpacking.  The product of two bf16 (fp16) values is exact in fp32 and the sum is rounded once,
// i.e. the same bits as fmaf(float(a), float(b), c).  AH / BH = 0: low half, 1: high half.
template <int AH, int BH>
__device__ __forceinline__ float act_fma(uint32_t a, uint32_t b, float c) {
  float d;
#if VD_ACT_IS_FP16
#define VD_FHFMA_ "fma.rn.f32.f16"
#else
#define VD_FHFMA_ "fma.rn.f32.bf16"
#endif
  if constexpr (AH == 0 && BH == 0)
    asm("{ .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n " VD_FHFMA_ " %0, al, bl, %3;\n}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else if constexpr (AH == 0 && BH == 1)
    asm("{ .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n " VD_FHFMA_ " %0, al, bh, %3;\n}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else if constexpr (AH == 1 && BH == 0)
    asm("{ .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n " VD_FHFMA_ " %0, ah, bl, %3;\n}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else
    asm("{ .reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n " VD_FHFMA_ " %0, ah, bh, %3;\n}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
#undef VD_FHFMA_
  return d;
}
// acc[0..7] += xs * w[0..7] with xs = half XH of the word x (a broadcast scalar)
template <int XH>
__device__ __forceinline__ void act_fma8s(float (&acc)[8], uint32_t x, const uint4& w) {
  acc[0] = act_fma<XH, 0>(x, w.x, acc[0]); acc[1] = act_fma<XH, 1>(x, w.x, acc[1]);
  acc[2] = act_fma<XH, 0>(x, w.y, acc[2]); acc[3] = act_fma<XH, 1>(x, w.y, acc[3]);
  acc[4] = act_fma<XH, 0>(x, w.z, acc[4]); acc[5] = act_fma<XH, 1>(x, w.z, acc[5]);
  acc[6] = act_fma<XH, 0>(x, w.w, acc[6]); acc[7] = act_fma<XH, 1>(x, w.w, acc[7]);
}
#endif  // __CUDACC__
