// skeleton of the persistent conv main loop: producer warp <-> MMA warp over STAGES full/empty barriers,
// no TMA (producer just arrives), NM MMAs (N=128) per k-block.  Reports cycles per k-block.
#include <cstdio>
#include "../../vae_gan_b200/csrc/sm100_ptx.cuh"
using namespace vg;
__device__ __forceinline__ void spin_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
  }
}
#define WAIT(bar, par) do { if (MODE & 16) spin_test_wait(bar, par); else ptx::mbar_wait(bar, par); } while (0)
// MODE bit0: producer waits on empty (real protocol) ; bit1: MMA thread does tc_fence_after ; bit2: whole-warp waits (else lane 0 only code)
template <int NM, int MODE, int STAGES>
__global__ void __launch_bounds__(384, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  __shared__ uint64_t full[STAGES], empty[STAGES], done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(&done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(&slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && !(MODE & 64)) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (MODE & 1) WAIT(&empty[stage], phase ^ 1u);
      if (ptx::elect_one()) ptx::mbar_arrive(&full[stage]);
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t kHi = ptx::smem_desc_hi(1024);
    const uint32_t a0 = ptx::smem_desc_lo(ptx::smem_u32(smem), 16), b0 = ptx::smem_desc_lo(ptx::smem_u32(smem) + 4 * 32768, 16);
    uint32_t a_lo = a0, b_lo = b0;
    int stage = 0; uint32_t phase = 0;
    long long t0 = clock64();
    bool ready = false;
    for (int i = 0; i < iters; ++i) {
      if (MODE & 64) { if (!(MODE & 128)) WAIT(&done, 1u); } else if (!ready) WAIT(&full[stage], phase);
      if (MODE & 2) ptx::tc_fence_after();
      if (MODE & 32) {   // poll the NEXT stage's barrier now; the answer is consumed after this k-block's MMAs were issued
        const int ns = (stage + 1 == STAGES) ? 0 : stage + 1;
        const uint32_t np = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
        ready = ptx::mbar_try_wait(&full[ns], np);
      }
      if (ptx::elect_one()) {
#pragma unroll
        for (int j = 0; j < NM; ++j)
          ptx::mma_bf16_ss_lohi(tmem + (j / 4) * 128, a_lo + (uint32_t)(((j / 4) * 16384 + (j % 4) * 32) >> 4), b_lo + (uint32_t)(((j % 4) * 32) >> 4), kHi, idesc, true);
        ptx::mma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1u; a_lo = a0; b_lo = b0; }
      else { a_lo += 32768 >> 4; b_lo += 16384 >> 4; }
    }
    long long t1 = clock64();
    if (ptx::elect_one()) ptx::mma_commit(&done);
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp >= 4 && (MODE & 8)) {
    // spinning bystanders on the same schedulers (like epilogue warps waiting for an accumulator)
    ptx::mbar_wait(&done, 0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem);
}
template <int NM, int MODE, int STAGES>
void run(const char* name, long long* d) {
  const int iters = 4000;
  cudaFuncSetAttribute(k<NM, MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<NM, MODE, STAGES><<<148, 384, 200 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-52s NM=%d STAGES=%d: %.1f cyc/k-block (ideal %d)\n", name, NM, STAGES, h[1] / (double)iters, NM * 64);
}
int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<8, 64 + 128 + 2, 4>("no wait at all, commit to cycling barriers", d);
  run<8, 64 + 2, 4>("always-true try_wait, no producer", d);
  run<8, 3, 4>("real protocol", d);
  run<4, 64 + 128 + 2, 4>("no wait at all, commit to cycling barriers", d);
  run<4, 64 + 2, 4>("always-true try_wait, no producer", d);
  run<4, 3, 4>("real protocol", d);
  return 0;
}
