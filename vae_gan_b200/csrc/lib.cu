// Library-level state: error string, launch counter, device init.
#include <stdarg.h>
#include <string.h>
#include "vg_common.cuh"

namespace vg {
std::atomic<unsigned long long> g_launches{0};
int g_num_sms = 0;
int g_force_simt = 0;
static thread_local char t_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
}  // namespace vg

extern "C" int vg_version(void) { return 100; }

extern "C" const char* vg_last_error(void) { return vg::t_err; }

extern "C" unsigned long long vg_launch_count(void) { return vg::g_launches.load(); }

extern "C" int vg_set_force_simt(int on) {
  int prev = vg::g_force_simt;
  vg::g_force_simt = on ? 1 : 0;
  return prev;
}

extern "C" int vg_init(int device) {
  cudaDeviceProp prop;
  VG_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    vg::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return VG_EARCH;
  }
  VG_CUDA(cudaSetDevice(device));
  vg::g_num_sms = prop.multiProcessorCount;
  const char* e = getenv("VG_FORCE_SIMT");
  if (e && atoi(e)) vg::g_force_simt = 1;
  return VG_OK;
}
