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
    default: return -1;
  }
}
