// Library-level state: error string, launch counter, device init.
#include <stdarg.h>
#include <string.h>
#include "vg_common.cuh"

namespace vg {
std::atomic<unsigned long long> g_launches{0};
int g_num_sms = 0;
int g_force_simt = 0;
int g_pdl = 1;
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
  e = getenv("VG_PDL");
  if (e) vg::g_pdl = atoi(e) ? 1 : 0;
  return VG_OK;
}

extern "C" int vg_enable_peer_access(int peer_device) {
  int cur = 0;
  VG_CUDA(cudaGetDevice(&cur));
  if (peer_device == cur) return VG_OK;
  int can = 0;
  VG_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) {
    vg::set_error("device %d cannot access peer %d", cur, peer_device);
    return VG_EUNSUPPORTED;
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return VG_OK;
  }
  VG_CUDA(e);
  return VG_OK;
}

// ---- setup-time helpers for the NVLink peer exchange buffers (not on the hot path) ----------------
// The exchange buffers must be plain cudaMalloc allocations so that their IPC handles can be opened by
// the other ranks of the node; these four calls are the only place the library touches the allocator.
extern "C" int vg_peer_alloc(size_t bytes, void** ptr) {
  VG_CHECK_ARG(ptr && bytes > 0, "bad args");
  VG_CUDA(cudaMalloc(ptr, bytes));
  VG_CUDA(cudaMemset(*ptr, 0, bytes));
  VG_CUDA(cudaDeviceSynchronize());
  return VG_OK;
}
extern "C" int vg_peer_free(void* ptr) {
  if (ptr) VG_CUDA(cudaFree(ptr));
  return VG_OK;
}
extern "C" int vg_peer_get_handle(void* ptr, void* handle64) {
  VG_CHECK_ARG(ptr && handle64, "null pointer");
  cudaIpcMemHandle_t h;
  VG_CUDA(cudaIpcGetMemHandle(&h, ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return VG_OK;
}
extern "C" int vg_peer_open_handle(const void* handle64, void** ptr) {
  VG_CHECK_ARG(ptr && handle64, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  VG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VG_OK;
}
extern "C" int vg_peer_close_handle(void* ptr) {
  if (ptr) VG_CUDA(cudaIpcCloseMemHandle(ptr));
  return VG_OK;
}
