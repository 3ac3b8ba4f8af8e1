// Placeholder entry points used only until the tcgen05 translation units are compiled in.
#include "common.cuh"
extern "C" int xr_fused_available(void) { return 0; }
extern "C" size_t xr_fused_pool_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int xr_fused_pool_loss(const void*, const void*, const void*, int64_t, int64_t, int64_t,
                                  int, const xr_loss_config*, const float*, float, float*, double*,
                                  float*, void*, size_t, void*) {
  xr::set_error("xr_fused_pool_loss: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" size_t xr_score_topk_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int xr_score_topk(const void*, int64_t, const void*, int64_t, int64_t, const float*,
                             const float*, int64_t, int64_t, const int64_t*, const int64_t*, float*,
                             int64_t*, void*, size_t, void*) {
  xr::set_error("xr_score_topk: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" int xr_fused_profile(int) { return XR_E_UNSUPPORTED; }
extern "C" int xr_fused_profile_read(float*, int) { return XR_E_UNSUPPORTED; }
extern "C" int xr_score_groupmax(const void*, int64_t, const void*, int64_t, int64_t, float*, int64_t,
                                 void*) {
  xr::set_error("xr_score_groupmax: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" int xr_fused_wait_stats(int, unsigned long long*) { return XR_E_UNSUPPORTED; }
extern "C" int xr_fused_timeline(long long*) { return XR_E_UNSUPPORTED; }
extern "C" size_t xr_pool_step_workspace_bytes(int64_t, int64_t) { return 0; }
extern "C" int xr_pool_step(const int64_t*, const int64_t*, const int64_t*, int64_t, const void*, int,
                            const void*, const uint8_t*, int64_t, int64_t, int, const xr_loss_config*,
                            float, void*, int, double*, int64_t*, int32_t*, void*, size_t, void*) {
  xr::set_error("xr_pool_step: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" int xr_score_groupmax_layout(int64_t, int64_t) { return 0; }
extern "C" int64_t xr_score_groupmax_ld(int64_t, int64_t n) { return 4 * ((n + 63) / 64); }
extern "C" size_t xr_fused_pool_all_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
extern "C" int xr_fused_pool_all(const void*, const void*, const void*, int64_t, int64_t, int64_t, int,
                                 const xr_loss_config*, double*, double*, void*, size_t, void*) {
  xr::set_error("xr_fused_pool_all: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" size_t xr_pool_step_monitor_workspace_bytes(int64_t, int64_t) { return 0; }
extern "C" int xr_pool_step_monitor(int64_t, int64_t, const xr_loss_config*, double*, double*, double*, void*,
                                    size_t, void*) {
  xr::set_error("xr_pool_step_monitor: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" int xr_pool_step_ingest(const int64_t*, const int64_t*, const int64_t*, int64_t, const void*, int,
                                   const void*, const uint8_t*, int64_t, int64_t, int64_t*, int32_t*, void*,
                                   size_t, void*) {
  xr::set_error("xr_pool_step_ingest: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
extern "C" int xr_pool_step_compute(int64_t, int64_t, int, const xr_loss_config*, float, void*, int, double*,
                                    void*, size_t, void*) {
  xr::set_error("xr_pool_step_compute: tcgen05 kernels not compiled into this build");
  return XR_E_UNSUPPORTED;
}
