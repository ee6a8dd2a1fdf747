// microbenchmark: cycles per tcgen05.mma as a function of N and of how the instruction stream looks
#include <cstdio>
#include "../../vae_gan_b200/csrc/sm100_ptx.cuh"
using namespace vg;
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 2) ptx::tmem_alloc<512>(&slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = ptx::smem_u32(smem);
    const uint32_t b_addr = a_addr + 4 * 16384;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (ptx::elect_one()) {
        if (MODE == 0) {          // 16 MMAs: 4 tiles x 4 k (same accumulator 4 times in a row)
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = ptx::make_smem_desc(a_addr + m * 16384 + kk * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc(b_addr + kk * 32, 16, 1024);
              ptx::mma_bf16_ss(tmem + (m * N) % 512, ad, bd, idesc, true);
            }
        } else if (MODE == 1) {   // 16 MMAs, k outer, tile inner (accumulator changes every MMA)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const uint64_t ad = ptx::make_smem_desc(a_addr + m * 16384 + kk * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc(b_addr + kk * 32, 16, 1024);
              ptx::mma_bf16_ss(tmem + (m * N) % 512, ad, bd, idesc, true);
            }
        } else if (MODE == 5 || MODE == 6) {   // 16 MMAs, commit after every 2 (5) or every 1 (6)
          const uint64_t ad = ptx::make_smem_desc(a_addr, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc(b_addr, 16, 1024);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ptx::mma_bf16_ss(tmem, ad, bd, idesc, true);
            if (MODE == 6 || (j & 1)) ptx::mma_commit(&bar);
          }
        } else {                   // 16 MMAs with the same operands and accumulator
          const uint64_t ad = ptx::make_smem_desc(a_addr, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc(b_addr, 16, 1024);
#pragma unroll
          for (int j = 0; j < 16; ++j) ptx::mma_bf16_ss(tmem, ad, bd, idesc, true);
        }
        if (MODE != 3 && MODE != 5 && MODE != 6) ptx::mma_commit(&bar);
      }
      __syncwarp();
      if (MODE == 4) ptx::mbar_wait(&bar, i & 1);
    }
    long long t1 = clock64();
    if (ptx::elect_one()) ptx::mma_commit(&bar);
    __syncwarp();
    if (MODE == 3) ptx::mbar_wait(&bar, 0);
    else if (MODE != 4) ptx::mbar_wait(&bar, iters & 1);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<512>(tmem);
}
template <int N, int MODE>
void run(const char* name, long long* d, int grid) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    k<N, MODE><<<grid, 128, 200 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-46s grid %3d N=%3d: issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA (ideal %d)\n", name, grid, N, h[0] / (16.0 * iters),
         h[1] / (16.0 * iters), N / 2);
}
int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int grid : {148}) {
    run<8, 2>("same operands, commit/16", d, grid);
    run<128, 5>("same operands, commit/2", d, grid);
    run<128, 6>("same operands, commit/1", d, grid);
    run<256, 6>("same operands, commit/1", d, grid);
    run<64, 6>("same operands, commit/1", d, grid);
    run<16, 0>("tile-outer k-inner, commit/16", d, grid);
    run<16, 2>("same operands, commit/16", d, grid);
    run<32, 0>("tile-outer k-inner, commit/16", d, grid);
    run<32, 2>("same operands, commit/16", d, grid);
    run<64, 0>("tile-outer k-inner, commit/16", d, grid);
    run<64, 1>("k-outer tile-inner, commit/16", d, grid);
    run<64, 2>("same operands, commit/16", d, grid);
    run<64, 3>("same operands, no commit", d, grid);
    run<128, 0>("tile-outer k-inner, commit/16", d, grid);
    run<128, 2>("same operands, commit/16", d, grid);
    run<256, 0>("tile-outer k-inner (2 acc), commit/16", d, grid);
    run<256, 2>("same operands, commit/16", d, grid);
  }
  return 0;
}
