// api.cu — error plumbing, device query and small shared entry points of libvdesens.
#include "common.cuh"

static thread_local char g_err[512] = "";

void vd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int vd_num_sms() {
  // per device ordinal (a process may drive several GPUs); plain ints: a racing first call stores the same value twice
  static int sms[64];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (dev >= 0 && dev < 64 && sms[dev] > 0) return sms[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (dev >= 0 && dev < 64) sms[dev] = n;
  return n;
}

extern "C" int vd_abi_version(void) { return VD_ABI_VERSION; }

// 0 = bf16 activations/weights (libvdesens.so), 1 = fp16 (libvdesens_fp16.so)
extern "C" int vd_act_is_fp16(void) { return VD_ACT_IS_FP16; }

extern "C" const char* vd_build_info(void) {
  return "libvdesens sm_100a (tcgen05/TMEM/TMA), activations " VD_ACT_NAME ", built " __DATE__ " " __TIME__;
}

extern "C" int vd_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return VD_ERR_INVALID;
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return VD_OK;
}

extern "C" int vd_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    vd_set_error("no CUDA device: %s", cudaGetErrorString(e));
    return VD_ERR_NO_DEVICE;
  }
  int sm = 0, mj = 0, mn = 0;
  VD_CHECK_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  VD_CHECK_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev));
  VD_CHECK_CUDA(cudaDeviceGetAttribute(&mn, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = mj;
  if (cc_minor) *cc_minor = mn;
  return VD_OK;
}

// conv1 reads windows [2*xo, 2*xo+8) of the padded row; pixel x sits at column x+3.  The pitch is a multiple of
// 16 pixels (128 bytes) so that a row is a whole number of the 128-byte chunks conv1's tensor map is made of.
extern "C" int vd_x0_pitch(int W) {
  const int Wout = (W + 1) / 2;
  const int need = 2 * Wout + 6;
  return (need + 15) & ~15;
}
