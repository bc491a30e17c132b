// Error plumbing and ABI version of the C boundary.
#include <cstdarg>

#include "common.cuh"

namespace rf {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace rf

extern "C" int rf_abi_version(void) { return RF_ABI_VERSION; }
extern "C" const char* rf_last_error(void) { return rf::g_error; }
extern "C" int rf_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(RfFovCropParams);
    case 1: return sizeof(RfGemmParams);
    case 2: return sizeof(RfConv3AssembleParams);
    case 3: return sizeof(RfConv3AssembleBwdParams);
    case 4: return sizeof(RfAttnParams);
    case 5: return sizeof(RfAttnBwdParams);
    case 6: return sizeof(RfDistilParams);
    case 7: return sizeof(RfDistilBwdParams);
    case 8: return sizeof(RfAreaResizeParams);
    default: return -1;
  }
}

extern "C" int rf_stage_frames_h2d(void* dst_dev, const void* src_host, int B, int T, const int* times, int n_sel, long long frame_bytes,
                                   void* stream) {
  using namespace rf;
  RF_CHECK_ARG(dst_dev && src_host && times && B > 0 && T > 0 && n_sel > 0 && frame_bytes > 0, "rf_stage_frames_h2d: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int j = 0; j < n_sel; ++j) {
    RF_CHECK_ARG(times[j] >= 0 && times[j] < T, "rf_stage_frames_h2d: frame time %d outside [0,%d)", times[j], T);
    RF_CUDA_OK(cudaMemcpy2DAsync(static_cast<char*>(dst_dev) + static_cast<long long>(j) * frame_bytes, static_cast<size_t>(n_sel) * frame_bytes,
                                 static_cast<const char*>(src_host) + static_cast<long long>(times[j]) * frame_bytes,
                                 static_cast<size_t>(T) * frame_bytes, static_cast<size_t>(frame_bytes), static_cast<size_t>(B),
                                 cudaMemcpyHostToDevice, s));
  }
  return RF_OK;
}
