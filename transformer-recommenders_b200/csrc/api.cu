// Error plumbing, device queries and ABI bookkeeping for libxfmr_b200.so.
#include "common.cuh"

#include <cstring>

namespace xr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return XR_E_CUDA;
}

static int g_reserved_sms = 0;   // SMs left free for concurrent kernels (xr_reserve_sms)

int sm_count_max() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

// what the persistent kernels size their grids with: the physical SM count minus the reservation
// (kept even: CTA pairs); workspaces are always sized with sm_count_max()
int sm_count() {
  int n = sm_count_max() - g_reserved_sms;
  if (n < 2) n = 2;
  return n & ~1;
}

}  // namespace xr

extern "C" const char* xr_last_error(void) { return xr::g_err; }
extern "C" int xr_abi_version(void) { return XR_ABI_VERSION; }

extern "C" int xr_reserve_sms(int n_reserved) {
  const int prev = xr::g_reserved_sms;
  if (n_reserved >= 0) xr::g_reserved_sms = n_reserved;
  return prev;
}

extern "C" int xr_device_info(int* sm_count, int* cc_major, int* cc_minor, int* has_tcgen05) {
  int dev = 0;
  XR_CUDA(cudaGetDevice(&dev));
  int n = 0, major = 0, minor = 0;
  XR_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  XR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  XR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = n;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  if (has_tcgen05) *has_tcgen05 = (major == 10) ? 1 : 0;
  return XR_OK;
}
