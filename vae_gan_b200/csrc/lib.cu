// Library-level state: error string, launch counter, device init.
#include <stdarg.h>
#include <string.h>
#include "vg_common.cuh"

namespace vg {
std::atomic<unsigned long long> g_launches{0};
int g_num_sms = 0;
int g_force_simt = 0;
int g_pdl = 1;
DetState g_det = {0, nullptr, 0, nullptr, 0};

void* det_scratch(size_t bytes) {
  if (bytes > g_det.scratch_bytes) {
    set_error("deterministic mode: %zu bytes of partial sums exceed the %zu-byte scratch given to vg_set_deterministic", bytes,
              g_det.scratch_bytes);
    return nullptr;
  }
  return g_det.scratch;
}
int* det_locks(long long n) {
  if (n > g_det.n_locks) {
    set_error("deterministic mode: %lld output tiles exceed the %d turn counters given to vg_set_deterministic", n, g_det.n_locks);
    return nullptr;
  }
  return g_det.locks;
}

// 32 values x 8 block slices per CTA; every slice sums its blocks in order, the slices are combined in order
template <typename T>
__global__ void __launch_bounds__(256) ordered_reduce_kernel(const T* __restrict__ partials, int nblocks, long long nvals,
                                                             T* __restrict__ out) {
  vg::pdl_entry();
  __shared__ T sm[8][33];
  const int x = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + x;
  const int per = (nblocks + 7) / 8;
  const int b0 = slice * per, b1 = min(nblocks, b0 + per);
  T s = 0;
  if (i < nvals)
    for (int b = b0; b < b1; ++b) s += partials[(long long)b * nvals + i];
  sm[slice][x] = s;
  __syncthreads();
  if (slice == 0 && i < nvals) {
    T t = sm[0][x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += sm[k][x];
    out[i] += t;
  }
}
int ordered_reduce_f64(const double* partials, int nblocks, long long nvals, double* out, cudaStream_t s) {
  vg::Launch((unsigned)cdiv(nvals, 32), 256, 0, s)(ordered_reduce_kernel<double>, partials, nblocks, nvals, out);
  VG_LAUNCHED();
  return VG_OK;
}
int ordered_reduce_f32(const float* partials, int nblocks, long long nvals, float* out, cudaStream_t s) {
  vg::Launch((unsigned)cdiv(nvals, 32), 256, 0, s)(ordered_reduce_kernel<float>, partials, nblocks, nvals, out);
  VG_LAUNCHED();
  return VG_OK;
}
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

extern "C" int vg_set_deterministic(int on, void* scratch, size_t scratch_bytes, void* locks, int n_locks) {
  if (!on) {
    vg::g_det = vg::DetState{0, nullptr, 0, nullptr, 0};
    return VG_OK;
  }
  VG_CHECK_ARG(scratch && scratch_bytes >= (1u << 20) && locks && n_locks >= 1024,
               "deterministic mode needs a scratch buffer (>= 1 MiB) and >= 1024 zeroed int32 turn counters");
  VG_CHECK_ARG(((uintptr_t)scratch & 15) == 0, "scratch must be 16-byte aligned");
  vg::g_det = vg::DetState{1, (unsigned char*)scratch, scratch_bytes, (int*)locks, n_locks};
  return VG_OK;
}
extern "C" int vg_get_deterministic(void) { return vg::g_det.on; }

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
