// Library-level entry points: version, error string, device check.
#include <stdarg.h>
#include "common.cuh"

namespace mmb {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace mmb

extern "C" int mmb_version(void) { return 100; }
extern "C" const char* mmb_last_error(void) { return mmb::g_error; }
extern "C" int mmb_device_supported(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    mmb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
  }
  if (prop.major != 10) {
    mmb::set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    return 0;
  }
  return 1;
}
