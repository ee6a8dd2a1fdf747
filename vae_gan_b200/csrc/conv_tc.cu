// tcgen05 / TMEM / TMA implicit-GEMM convolution kernels for sm_100a (bf16 in, fp32 accumulate).
//
// One kernel family covers every dense contraction of the reference's hot path (SURVEY.md K1-K3,
// K5): Conv2d 3x3 / 1x1 at stride 1 and 2, ConvTranspose2d 4x4 stride 2, forward and dgrad
// (tc_conv_kernel, "tap table" form) and all their weight gradients (tc_wgrad_kernel).
//
// tc_conv_kernel: D[128 pixels][BN channels] = sum over taps t, 64-channel chunks kc of
//     A_t,kc[128 pixels][64]  .  W_t,kc[BN][64]^T
//   * A tile = ONE TMA box {64 ch, TW, TH, TN} of the NHWC activation at the tap's (dx,dy) shift;
//     out-of-bounds coordinates are zero-filled by TMA, which is exactly the convolution padding.
//     A box lands as 128 rows x 128 B with the 128B swizzle = the canonical K-major UMMA layout.
//   * stride-2 gathers read one of four "parity" tensor maps (even/odd rows x even/odd columns
//     of the fine tensor), so every load is still a dense unit-stride box;
//   * stride-2 scatters (ConvTranspose2d forward, Conv2d-s2 dgrad) are decomposed into the four
//     output sub-pixel phases (blockIdx.z), each a small stride-1 convolution - no zero insertion.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
//     warps 4.. = epilogue (tcgen05.ld -> bias / Dropout2d column scale -> bf16|f32 -> global).
//   Three forms, chosen per layer by tc_conv_run: one-shot CTAs (tc_conv_kernel, small problems and split-K),
//   persistent CTAs with a double-buffered TMEM accumulator (tc_conv_persist_kernel), and persistent CTA
//   pairs issuing M=256 cta_group::2 MMAs (tc_conv_pair_kernel, 128/256-wide tiles).
//
// tc_wgrad_kernel: dW[128 = 2 x 64 (tap, c_s chunk)][BN = c_u tile] += S^T . U over a range of
//   64-pixel boxes (split-K across CTAs, fp32 atomics into the torch-layout gradient).  Both
//   operands are MN-major (channels contiguous, pixels along K) straight from the same TMA boxes.
#include <algorithm>
#include <array>
#include <map>
#include <mutex>
#include <vector>
#include "vg_common.cuh"
#include "sm100_ptx.cuh"

namespace vg {

// ------------------------------------------------------------------------------------------
// parameter blocks (passed by value as __grid_constant__)
// ------------------------------------------------------------------------------------------
struct TcTap {
  int8_t map, dy, dx, pad_;
  int32_t wrow;   // row offset of this tap's slab in the weight tensor map (= tap * n_out)
};
struct TcPhase {
  int32_t ntaps, oy_off, ox_off, pad_;
  TcTap taps[16];
};
struct alignas(64) TcConvParams {
  CUtensorMap in_maps[4];
  CUtensorMap w_map;
  TcPhase phases[4];
  int tiles_x, tiles_y, tiles_n;
  int tw_log2, th_log2, TW, TH, TN;
  int gw, gh, gn;        // GEMM pixel grid of one phase
  int k_chunks;          // reduction channels / 64
  int n_out;             // output channels
  int n_store;           // columns actually stored per N tile (== BN except for single-channel outputs)
  int OH, OW, os;        // output tensor spatial dims and phase stride
  int out_f32;
  int ksplit, k_per_split;   // split-K over blockIdx.z (only with a single phase, fp32 atomics)
  void* out;
  const float* bias;
  const float* colscale;
  const float* sigma;    // nullable: the accumulator is multiplied by 1 / sigma[group] (spectral norm applied in the epilogue,
  int sigma_group_n;     //   so the packed weights stay valid across forwards); group = sample / sigma_group_n (0: one group)
  double* stats;         // nullable: [2*n_out] per-channel sum / sum of squares of the stored output
  // inference epilogue (BatchNorm-folded sampling path, bf16 outputs only; see VgConvEpilogue in the header)
  float act_slope;       // LeakyReLU slope of the stored output (1 = none)
  const void* residual;  // nullable: same shape/dtype as out, added (after rounding the accumulator to bf16) before the activation
  void* out2;            // nullable second output: leaky_relu(post_scale[c] * out + post_shift[c], post_slope)
  const float* post_scale;
  const float* post_shift;
  float post_slope;
  CUtensorMap out_maps[4];  // bf16 output as {32 channels, bx, by, bn} boxes (64-byte swizzle = the staging buffer's XOR
                            // pattern); one per output phase: a stride-2 scatter writes four interleaved sub-grids
  int tma_store;         // 1: the staged 32 x 32 output chunks leave through cp.async.bulk.tensor stores (out_map)
  int* det_locks;        // deterministic mode, split-K: one turn counter per output tile (the splits add in split order)
  long long det_split_stride;  // deterministic mode, split-K: != 0 -> split z STORES its partial tile at out + z * stride (scratch);
                               //   ordered_reduce_f32 adds the splits in order afterwards (no serialisation, no atomics)
};

struct alignas(64) TcWgradParams {
  CUtensorMap s_maps[4];
  CUtensorMap u_map;
  TcTap taps[16];
  int ntaps, cs_chunks, cs, cu;
  int tiles_x, tiles_y, tiles_n, TW, TH, TN;
  int n_boxes, boxes_per_split;
  int packed;    // 1: dw is the packed scratch [tap][c_s][c_u] (vector reductions); 0: torch layout [c_u][c_s][tap]
  float* dw;
  int* det_locks;  // deterministic mode: one turn counter per (slab pair, c_u tile); the pixel splits add in split order
  long long det_split_stride;  // deterministic mode: != 0 -> split z STORES its partial sums at dw + z * stride (scratch)
};

constexpr int kTcThreads = 256;
constexpr int kABytes = 128 * 128;   // 128 rows x 64 bf16
constexpr int kStageBytes = 32 * 64;  // per epilogue warp: 32 rows x 32 bf16 of one output chunk (coalescing transpose)

template <int BN, int MT, int STAGES>
struct ConvSmem {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kBytes = STAGES * (MT * kABytes + kBBytes) + (2 * STAGES + 1) * 8 + 16 + 1024 + 4 * kStageBytes + 16 + 512;
};


// Column sums of a 32 x 32 block held one ROW per lane (v[j] = column j): recursive halving over the lane bits - in step H
// the lanes with bit H set keep the upper H columns of what they hold and receive the lower lanes' share of them (and vice
// versa), so after the steps 16, 8, 4, 2, 1 lane l holds the sum of column l over all 32 rows.  31 shuffles, all register
// indices static.  (Round 1 transposed through shared memory instead: 1.5 k shared accesses per chunk, measured +35..60 %.)
template <int H>
__device__ __forceinline__ void fold_cols(float (&v)[32], bool up) {
#pragma unroll
  for (int k = 0; k < H; ++k) {
    const float send = up ? v[k] : v[k + H];
    const float keep = up ? v[k + H] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, H);
  }
}
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
  fold_cols<16>(v, (lane & 16) != 0);
  fold_cols<8>(v, (lane & 8) != 0);
  fold_cols<4>(v, (lane & 4) != 0);
  fold_cols<2>(v, (lane & 2) != 0);
  fold_cols<1>(v, (lane & 1) != 0);
  return v[0];
}

// Finishes one 32-column chunk of one accumulator row per thread: 1/sigma, bias, Dropout2d column scale, store
// (bf16 / fp32 / split-K vector reductions / single channel) and - with `do_stats` - adds this warp's per-column sum
// and sum of squares of the values AS STORED to (st_sum, st_sq): lane l owns column l.  Those feed the BatchNorm
// that follows the convolution (README.md:192), saving a full read of y by a separate statistics kernel.
//
// bf16 outputs go through `stage` (a 2 KB per-warp shared-memory buffer) when it is given: lane = pixel row holds 64 B of
// consecutive channels, so a direct 16-byte store per lane touches 32 different 128-byte lines per instruction - the
// store path, not the issue rate, bounded the epilogue (8 epilogue warps instead of 4 changed nothing on the 64-wide
// layers).  The warp writes its 32 x 64 B block into shared memory (XOR-swizzled 16-byte units, conflict-free both ways),
// reads it back transposed and stores with FOUR consecutive lanes covering one row's 64 bytes: 8 lines per instruction.
template <bool INF>
__device__ __forceinline__ void epilogue_chunk(const TcConvParams& p, const uint32_t (&r)[32], bool valid, bool add_bias,
                                               long long opix, int gn, int ncol, bool first_chunk, bool do_stats, int lane,
                                               float& st_sum, float& st_sq, uint8_t* stage = nullptr, int sx = 0, int sy = 0,
                                               int sn = 0, int out_phase = 0) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.sigma != nullptr && valid) {
    const float inv = 1.0f / __ldg(p.sigma + (p.sigma_group_n > 0 ? gn / p.sigma_group_n : 0));
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= inv;
  }
  if (p.bias != nullptr && add_bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (ncol + j < p.n_out) v[j] += __ldg(p.bias + ncol + j);
  }
  if (p.colscale != nullptr && valid) {
    const float* cs = p.colscale + (long long)gn * p.n_out + ncol;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= __ldg(cs + j);
  }
  const bool staged = stage != nullptr && !p.out_f32 && p.ksplit <= 1 && p.n_store != 1;
  if (staged) {
    if (INF && p.act_slope != 1.0f && p.residual == nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.act_slope;
    }
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    const bool tma_out = !INF && p.tma_store != 0;
    if (tma_out) {
      // the previous chunk's bulk store must have READ the staging buffer before it is overwritten
      if (lane == 0) ptx::bulk_wait_group_read0();
      __syncwarp();
    }
    uint4* srow = reinterpret_cast<uint4*>(stage + lane * 64);
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) srow[q ^ sw] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    if (tma_out) {
      // TMA-store epilogue: the warp's 32 pixels are a {bx, by, bn} sub-box of the tile starting at (sx, sy, sn); pixels
      // outside the tensor are clipped by the TMA unit (= the `valid` mask of the manual path)
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_4d(&p.out_maps[out_phase], stage, ncol, sx, sy, sn);
        ptx::bulk_commit_group();
      }
      __syncwarp();
    } else {
    __syncwarp();
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out);
    const bool post = INF && (p.residual != nullptr || p.out2 != nullptr);       // kernel-uniform; compiled out of the training kernels
    float ps[8], pt[8];
    if (INF && p.out2 != nullptr) {
      const int c0 = ncol + (lane & 3) * 8;       // this lane's 8 channels are the same for all four row groups
#pragma unroll
      for (int j = 0; j < 8; ++j) { ps[j] = __ldg(p.post_scale + c0 + j); pt[j] = __ldg(p.post_shift + c0 + j); }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + (lane >> 2), q = lane & 3;
      uint4 val = *reinterpret_cast<const uint4*>(stage + row * 64 + ((q ^ ((row >> 1) & 3)) << 4));
      const long long op = __shfl_sync(0xffffffffu, opix, row);
      const int ok = __shfl_sync(0xffffffffu, (int)valid, row);
      if (ok) {
        const long long e = op * p.n_out + ncol + q * 8;
        if (INF && post) {
          float f[8];
          const uint32_t wv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(wv[i] << 16); f[2 * i + 1] = __uint_as_float(wv[i] & 0xffff0000u); }
          if (p.residual != nullptr) {
            const uint4 rr = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + e);
            const uint32_t rv[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) { f[2 * i] += __uint_as_float(rv[i] << 16); f[2 * i + 1] += __uint_as_float(rv[i] & 0xffff0000u); }
            if (p.act_slope != 1.0f) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * p.act_slope;
            }
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              o[i] = *reinterpret_cast<uint32_t*>(&h);
              // the second output is computed from the value AS STORED (what a separate elementwise pass would read)
              f[2 * i] = __uint_as_float(o[i] << 16);
              f[2 * i + 1] = __uint_as_float(o[i] & 0xffff0000u);
            }
            val = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if (p.out2 != nullptr) {
            uint32_t o2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float a0 = fmaf(ps[2 * i], f[2 * i], pt[2 * i]), a1 = fmaf(ps[2 * i + 1], f[2 * i + 1], pt[2 * i + 1]);
              a0 = a0 > 0.f ? a0 : a0 * p.post_slope;
              a1 = a1 > 0.f ? a1 : a1 * p.post_slope;
              __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
              o2[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + e) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
          }
        }
        *reinterpret_cast<uint4*>(ob + e) = val;
      }
    }
    __syncwarp();                      // the buffer is rewritten by this warp's next chunk
    }
  } else if (valid) {
    if (p.n_store == 1) {
      // single-channel output (Conv2d C->1 forward / Conv2d 1->C dgrad): only column 0 is real
      // (static register index - a dynamic one would push v[] into local memory)
      if (first_chunk) {
        if (p.out_f32) reinterpret_cast<float*>(p.out)[opix] = v[0];
        else reinterpret_cast<__nv_bfloat16*>(p.out)[opix] = __float2bfloat16_rn(v[0]);
      }
    } else if (p.ksplit > 1) {
      // split-K partial sums: 16-byte vector reductions (a quarter of the RED instructions of scalar atomics - the
      // Linear layers' one-shot CTAs spent most of their time issuing them; n_out % 64 == 0 keeps them aligned)
      float* o = reinterpret_cast<float*>(p.out) + opix * p.n_out + ncol;
      if (p.det_split_stride != 0) {
        // deterministic mode: this split's partial tile goes to its own scratch slab (one-shot kernel: split = blockIdx.z)
        o += (long long)blockIdx.z * p.det_split_stride;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
      }
    } else if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + opix * p.n_out + ncol;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + opix * p.n_out + ncol;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * i], v[j + 2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(o + j) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  if (do_stats) {
    float q[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = valid ? (p.out_f32 ? v[j] : __bfloat162float(__float2bfloat16_rn(v[j]))) : 0.f;
      v[j] = a;
      q[j] = a * a;
    }
    st_sq += colsum32(q, lane);
    st_sum += colsum32(v, lane);
  }
}

// ------------------------------------------------------------------------------------------
// forward / dgrad implicit GEMM
// ------------------------------------------------------------------------------------------
// MT pixel tiles (128 rows each) per CTA share every weight tile: B traffic per FLOP drops by MT
// (the implicit GEMM re-reads its operands from L2 for every tap, so it is L2-bandwidth bound
// unless each staged byte feeds enough MMAs).
template <int BN, int MT, int STAGES, bool INF>
__global__ void __launch_bounds__(kTcThreads) tc_conv_kernel(const __grid_constant__ TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kBBytes = BN * 128;
  constexpr int kAStage = MT * kABytes;
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * kBBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  // 512-byte aligned: the staging buffers' XOR pattern is then exactly TMA's 64-byte swizzle (absolute address bits)
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 511) & ~(uintptr_t)511);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : blockIdx.z];
  const int ntaps = ph.ntaps;
  const int kb0 = p.ksplit > 1 ? (int)blockIdx.z * p.k_per_split : 0;
  const int kb1 = p.ksplit > 1 ? min(ntaps * p.k_chunks, kb0 + p.k_per_split) : ntaps * p.k_chunks;
  const int num_k = max(0, kb1 - kb0);
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  const int tile0 = blockIdx.x * MT;
  const int nvalid = min(MT, total_tiles - tile0);
  int x0s[MT], y0s[MT], n0s[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    int t = min(tile0 + i, total_tiles - 1);
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int tn = t / p.tiles_y;
    x0s[i] = tx * p.TW; y0s[i] = ty * p.TH; n0s[i] = tn * p.TN;
  }
  const int ncol0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.in_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  constexpr int kTmemCols = (MT * BN < 32) ? 32 : MT * BN;
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  vg::pdl_entry();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; nothing above touches global memory
  const uint32_t tmem_base = *tmem_slot;

  if (num_k > 0) {
    // both role warps loop warp-uniformly and elect a lane only around the TMA / MMA issue (a divergent
    // `lane == 0` branch makes nvcc wrap every uniform-datapath instruction in an elect-and-retry loop)
    if (warp == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int tp = kb0 / p.k_chunks, kc = kb0 % p.k_chunks;
      for (int i = kb0; i < kb1; ++i) {
        const TcTap tap = ph.taps[tp];
        const CUtensorMap* im = &p.in_maps[tap.map];
        ptx::mbar_wait(&empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[stage], nvalid * kABytes + kBBytes);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            if (m < nvalid)
              ptx::tma_load_4d(sA + stage * kAStage + m * kABytes, im, &full[stage], kc * 64, x0s[m] + tap.dx, y0s[m] + tap.dy, n0s[m]);
          ptx::tma_load_2d(sB + stage * kBBytes, &p.w_map, &full[stage], kc * 64, tap.wrow + ncol0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (++kc == p.k_chunks) { kc = 0; ++tp; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t kDescHi = ptx::smem_desc_hi(1024);
      const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA), 16), b_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sB), 16);
      uint32_t a_lo = a_lo0, b_lo = b_lo0;      // descriptor low words of the current stage
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_k; ++i) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            if (m < nvalid) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::mma_bf16_ss_lohi(tmem_base + (uint32_t)(m * BN), a_lo + (uint32_t)((m * kABytes + k * 32) >> 4),
                                      b_lo + (uint32_t)((k * 32) >> 4), kDescHi, idesc, (i | k) != 0);
            }
          }
          ptx::mma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; a_lo = a_lo0; b_lo = b_lo0; }
        else { a_lo += (uint32_t)(kAStage >> 4); b_lo += (uint32_t)(kBBytes >> 4); }
      }
      if (ptx::elect_one()) ptx::mma_commit(tmem_full);
      __syncwarp();
      vg::pdl_tail_trigger();
    }
  }

  if (warp >= 4) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int xl = row & (p.TW - 1);
    const int yl = (row >> p.tw_log2) & (p.TH - 1);
    const int nl = row >> (p.tw_log2 + p.th_log2);
    // first pixel of this warp's 32-row quarter inside the tile: origin of its TMA-store sub-box
    const int row0 = q * 32;
    const int xl0 = row0 & (p.TW - 1), yl0 = (row0 >> p.tw_log2) & (p.TH - 1), nl0 = row0 >> (p.tw_log2 + p.th_log2);
    if (num_k > 0) {
      ptx::mbar_wait(tmem_full, 0);
      ptx::tc_fence_after();
    }
    const bool do_stats = p.stats != nullptr && p.ksplit <= 1 && num_k > 0;
    float st_s[BN / 32], st_q[BN / 32];
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) st_s[c] = st_q[c] = 0.f;
    // deterministic split-K: this split's reductions go after those of split z - 1 of the same output tile
    int* const det_lock = (p.det_locks != nullptr && p.ksplit > 1) ? p.det_locks + (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    if (det_lock) {
      if (lane == 0) det_wait_turn(det_lock, (int)blockIdx.z);
      __syncwarp();
    }
#pragma unroll 1
    for (int m = 0; m < nvalid; ++m) {
      const int gx = x0s[m] + xl, gy = y0s[m] + yl, gn = n0s[m] + nl;
      const bool valid = gx < p.gw && gy < p.gh && gn < p.gn && !(p.ksplit > 1 && num_k == 0 && p.det_split_stride == 0);
      const long long opix = ((long long)gn * p.OH + (long long)gy * p.os + ph.oy_off) * p.OW + (long long)gx * p.os + ph.ox_off;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        const int c0 = c * 32;
        uint32_t r[32];
        if (num_k > 0) {
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * BN + c0), r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        epilogue_chunk<INF>(p, r, valid, p.ksplit <= 1 || blockIdx.z == 0, opix, gn, ncol0 + c0, c == 0, do_stats, lane, st_s[c], st_q[c],
                       stage_base + q * kStageBytes, x0s[m] + xl0, y0s[m] + yl0, n0s[m] + nl0, (int)blockIdx.z);
      }
    }
    if (do_stats) {
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        const int col = ncol0 + c * 32 + lane;
        if (col < p.n_out) {
          atomicAdd(p.stats + col, (double)st_s[c]);
          atomicAdd(p.stats + p.n_out + col, (double)st_q[c]);
        }
      }
    }
    if (det_lock) {
      __threadfence();
      asm volatile("bar.sync 2, 128;" ::: "memory");      // the four epilogue warps
      if (warp == 4 && lane == 0) det_publish_turn(det_lock, (int)blockIdx.z + 1 == p.ksplit ? 0 : (int)blockIdx.z + 1);
    }
    if (!INF && p.tma_store && lane == 0) ptx::bulk_wait_group0();      // this warp's bulk stores have landed
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// NOTES (measured on B200, round 1; microbenchmarks in scripts/microbench/, numbers in DESIGN.md section 3):
//  * a tcgen05.mma (M=128, K=16, SS mode) takes max(N/2, 32 + N/4) cycles - the operands are re-read from
//    shared memory at 128 B/cycle - so N=64 tiles cap at 67 % of the tensor peak, N>=128 can reach it;
//  * the tensor pipe queues only ~1-2 instructions and every mbarrier wait stalls the issuing thread for
//    ~105 cycles: one wait per k-block costs a ~135-cycle bubble unless the queued MMAs are N=256 wide;
//  * under a divergent `lane == 0` branch nvcc wraps each UTCHMMA/UTMALDG in an elect-and-retry loop and
//    rebuilding 64-bit descriptors costs ~14 uniform instructions per MMA: the role warps therefore run
//    warp-uniformly, elect only around the issue, and advance 32-bit descriptor halves by adds;
//  * the chip is power-capped under tensor load (SM clock 1.3-1.5 GHz), so removing loads, stores or waits
//    buys cycles but little time; what pays is fewer shared-memory/L2 bytes per FLOP (CTA pairs below).
//
// persistent variant: one CTA per SM loops over work items (phase|k-split, N tile, group of MT
// pixel tiles); the TMEM accumulator is double-buffered so the epilogue of item i overlaps the
// main loop of item i+1, and the MT pixel tiles of an item share every staged weight tile.
// ------------------------------------------------------------------------------------------
template <int BN, int MT, int STAGES, int EW>
struct ConvPersistSmem {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kBytes = STAGES * (MT * kABytes + kBBytes) + (2 * STAGES + 4) * 8 + 16 + 1024 + EW * kStageBytes + 16 + 512;
};

// EW = number of epilogue warps (4 or 8).  Warp w may only touch TMEM lanes 32*(w%4)..+31, so with
// eight warps two warps share each lane quarter and split the 32-column chunks between them; that
// doubles the epilogue throughput, which is what makes the fused BatchNorm statistics free.
template <int BN, int MT, int STAGES, int EW, int G, bool INF>
__global__ void __launch_bounds__(128 + 32 * EW, 1) tc_conv_persist_kernel(const __grid_constant__ TcConvParams p, int n_ntiles,
                                                                            int n_groups, int n_z, int n_work) {
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  static_assert(STAGES % G == 0 && STAGES / G >= 2, "at least two barrier groups");
  constexpr int kChunkGroups = EW / 4;
  extern __shared__ uint8_t smem_raw[];
  constexpr int kBBytes = BN * 128;
  constexpr int kAStage = MT * kABytes;
  constexpr int kAccCols = MT * BN;
  constexpr int kTmemCols = 2 * kAccCols;
  static_assert(kTmemCols <= 512 && kTmemCols >= 32, "TMEM budget");
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * kBBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;       // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // 512-byte aligned: the staging buffers' XOR pattern is then exactly TMA's 64-byte swizzle (absolute address bits)
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 511) & ~(uintptr_t)511);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.in_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], EW);     // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  vg::pdl_entry();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; nothing above touches global memory
  const uint32_t tmem_base = *tmem_slot;

  // work item -> (z, n tile, pixel-tile group); groups vary fastest so that concurrently running
  // CTAs stream the same weight tile (L2 reuse) over neighbouring pixels
  auto decode = [&](int w, int& z, int& nt, int& g) {
    const int per_z = n_ntiles * n_groups;
    z = w / per_z;
    const int r = w - z * per_z;
    nt = r / n_groups;
    g = r - nt * n_groups;
  };
  auto k_range = [&](int z, const TcPhase& ph, int& kb0, int& kb1) {
    const int total = ph.ntaps * p.k_chunks;
    if (p.ksplit > 1) { kb0 = z * p.k_per_split; kb1 = min(total, kb0 + p.k_per_split); }
    else { kb0 = 0; kb1 = total; }
    if (kb1 < kb0) kb1 = kb0;
  };
  auto tile_origin = [&](int t, int& x0, int& y0, int& n0) {
    t = min(t, total_tiles - 1);
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int tn = t / p.tiles_y;
    x0 = tx * p.TW; y0 = ty * p.TH; n0 = tn * p.TN;
  };

  // Producer and MMA warps run their loops WARP-UNIFORMLY and elect one lane only around the TMA / MMA
  // issue (under a divergent `lane == 0` branch nvcc wraps every uniform-datapath instruction - UTMALDG,
  // UTCHMMA, UTCBAR - in an elect-and-retry loop).
  // G consecutive k-blocks share ONE full/empty barrier pair: measured on B200 (scratch microbenchmarks,
  // DESIGN.md section 3), every mbarrier wait in the MMA-issuing thread stalls it for ~105 cycles while
  // the tensor pipe only queues ~1-2 instructions, so each wait is a ~135-cycle bubble unless the MMAs
  // are N=256 wide (128 cycles each).  Grouping halves the number of waits per MMA.
  constexpr int kGroups = STAGES / G;
  if (warp == 0) {
    int grp = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int z, nt, g;
      decode(w, z, nt, g);
      const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
      int kb0, kb1;
      k_range(z, ph, kb0, kb1);
      if (kb1 == kb0) continue;
      const int tile0 = g * MT;
      const int nvalid = min(MT, total_tiles - tile0);
      int x0s[MT], y0s[MT], n0s[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) tile_origin(tile0 + m, x0s[m], y0s[m], n0s[m]);
      const int ncol0 = nt * BN;
      int tp = kb0 / p.k_chunks, kc = kb0 % p.k_chunks;
      for (int i = kb0; i < kb1; i += G) {
        const int nk = min(G, kb1 - i);
        ptx::mbar_wait(&empty[grp], phase ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[grp], nk * (nvalid * kABytes + kBBytes));
#pragma unroll
          for (int j = 0; j < G; ++j) {
            if (j < nk) {
              const TcTap tap = ph.taps[tp];
              const CUtensorMap* im = &p.in_maps[tap.map];
              const int stage = grp * G + j;
#pragma unroll
              for (int m = 0; m < MT; ++m)
                if (m < nvalid)
                  ptx::tma_load_4d(sA + stage * kAStage + m * kABytes, im, &full[grp], kc * 64, x0s[m] + tap.dx, y0s[m] + tap.dy, n0s[m]);
              ptx::tma_load_2d(sB + stage * kBBytes, &p.w_map, &full[grp], kc * 64, tap.wrow + ncol0);
              if (++kc == p.k_chunks) { kc = 0; ++tp; }
            }
          }
        }
        __syncwarp();
        // every lane tracks (tp, kc): only the elected lane advanced them above
        {
          const int adv = (i - kb0) + nk + kb0;        // absolute k-block index after this group
          tp = adv / p.k_chunks; kc = adv - tp * p.k_chunks;
        }
        if (++grp == kGroups) { grp = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t kDescHi = ptx::smem_desc_hi(1024);
    const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA), 16), b_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sB), 16);
    uint32_t a_lo = a_lo0, b_lo = b_lo0;      // descriptor low words of the current group's first stage
    int grp = 0;
    uint32_t phase = 0;
    int it = 0;                       // number of accumulator uses so far
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int z, nt, g;
      decode(w, z, nt, g);
      const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
      int kb0, kb1;
      k_range(z, ph, kb0, kb1);
      if (kb1 == kb0) continue;
      const int nvalid = min(MT, total_tiles - g * MT);
      const int as = it & 1;
      ptx::mbar_wait(&acc_empty[as], (uint32_t)(((it >> 1) & 1) ^ 1));
      ptx::tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(as * kAccCols);
      for (int i = kb0; i < kb1; i += G) {
        const int nk = min(G, kb1 - i);
        ptx::mbar_wait(&full[grp], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < G; ++j) {
            if (j < nk) {
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                if (m < nvalid) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    ptx::mma_bf16_ss_lohi(acc + (uint32_t)(m * BN), a_lo + (uint32_t)((j * kAStage + m * kABytes + k * 32) >> 4),
                                          b_lo + (uint32_t)((j * kBBytes + k * 32) >> 4), kDescHi, idesc, (i + j > kb0) || (k > 0));
                }
              }
            }
          }
          ptx::mma_commit(&empty[grp]);
        }
        __syncwarp();
        if (++grp == kGroups) { grp = 0; phase ^= 1u; a_lo = a_lo0; b_lo = b_lo0; }
        else { a_lo += (uint32_t)((G * kAStage) >> 4); b_lo += (uint32_t)((G * kBBytes) >> 4); }
      }
      if (ptx::elect_one()) ptx::mma_commit(&acc_full[as]);
      __syncwarp();
      ++it;
    }
    vg::pdl_tail_trigger();
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = ew & 3;                 // TMEM lane quarter this warp may access
    const int cgrp = ew >> 2;             // which share of the 32-column chunks it handles
    const int row = q * 32 + lane;
    const int xl = row & (p.TW - 1);
    const int yl = (row >> p.tw_log2) & (p.TH - 1);
    const int nl = row >> (p.tw_log2 + p.th_log2);
    // first pixel of this warp's 32-row quarter inside the tile: origin of its TMA-store sub-box
    const int row0 = q * 32;
    const int xl0 = row0 & (p.TW - 1), yl0 = (row0 >> p.tw_log2) & (p.TH - 1), nl0 = row0 >> (p.tw_log2 + p.th_log2);
    int it = 0;
    const bool do_stats = p.stats != nullptr && p.ksplit <= 1;
    float st_s[BN / 32], st_q[BN / 32];
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) st_s[c] = st_q[c] = 0.f;
    int st_nt = -1;                      // N tile the register statistics belong to
    auto flush_stats = [&]() {
      if (st_nt < 0) return;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        if ((c % kChunkGroups) != cgrp) continue;
        const int col = st_nt * BN + c * 32 + lane;
        if (col < p.n_out) {
          atomicAdd(p.stats + col, (double)st_s[c]);
          atomicAdd(p.stats + p.n_out + col, (double)st_q[c]);
        }
        st_s[c] = st_q[c] = 0.f;
      }
    };
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      int z, nt, g;
      decode(w, z, nt, g);
      const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
      int kb0, kb1;
      k_range(z, ph, kb0, kb1);
      const bool has_k = kb1 > kb0;
      if (!has_k && p.ksplit > 1) continue;          // empty k-split: contributes nothing
      const int tile0 = g * MT;
      const int nvalid = min(MT, total_tiles - tile0);
      const int ncol0 = nt * BN;
      const int as = it & 1;
      if (do_stats && nt != st_nt) {
        flush_stats();
        st_nt = nt;
      }
      if (has_k) {
        ptx::mbar_wait(&acc_full[as], (uint32_t)((it >> 1) & 1));
        ptx::tc_fence_after();
      }
      const uint32_t acc = tmem_base + (uint32_t)(as * kAccCols) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int m = 0; m < nvalid; ++m) {
        int x0, y0, n0;
        tile_origin(tile0 + m, x0, y0, n0);
        const int gx = x0 + xl, gy = y0 + yl, gn = n0 + nl;
        const bool valid = gx < p.gw && gy < p.gh && gn < p.gn;
        const long long opix = ((long long)gn * p.OH + (long long)gy * p.os + ph.oy_off) * p.OW + (long long)gx * p.os + ph.ox_off;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if ((c % kChunkGroups) != cgrp) continue;
          const int c0 = c * 32;
          uint32_t r[32];
          if (has_k) {
            ptx::tmem_ld_32x32(acc + (uint32_t)(m * BN + c0), r);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          epilogue_chunk<INF>(p, r, valid, p.ksplit <= 1 || z == 0, opix, gn, ncol0 + c0, c == 0, do_stats && has_k, lane, st_s[c], st_q[c],
                         stage_base + ew * kStageBytes, x0 + xl0, y0 + yl0, n0 + nl0, z);
        }
      }
      if (has_k) {
        // this warp is done reading accumulator stage `as`: hand it back to the MMA warp
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
        ++it;
      }
    }
    if (do_stats) flush_stats();
    if (!INF && p.tma_store && lane == 0) ptx::bulk_wait_group0();      // this warp's bulk stores have landed
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster - two SMs - issue ONE M=256 MMA.  Each CTA
// stages its own MT pixel tiles (A, 128 rows each) and HALF of the weight tile (B, BN/2 rows); the
// tensor core of each SM reads its own A and both halves of B, so the per-SM shared-memory reads per
// MMA drop from (4 KB + BN*32 B) to (4 KB + BN*16 B) and the staged weight bytes halve as well.
// Protocol (leader = cluster rank 0):
//   * producer thread in BOTH CTAs: waits its own empty[s] (multicast commit), issues its TMA loads
//     with the completion bytes counted on the LEADER's full[s]; the leader arms full[s] with the
//     bytes of both CTAs;
//   * MMA thread in the leader only; tcgen05.commit multicasts to empty[s] / acc_full[a] of both CTAs;
//   * epilogue warps in both CTAs drain their own TMEM and arrive (cluster scope) on the leader's
//     acc_empty[a], which therefore counts 2*EW arrivals.
// Requires total pixel tiles % (2*MT) == 0 (the host falls back to the single-CTA kernel otherwise).
// ------------------------------------------------------------------------------------------
template <int BN, int MT, int STAGES, int EW>
struct ConvPairSmem {
  static constexpr int kBBytes = (BN / 2) * 128;
  static constexpr int kBytes = STAGES * (MT * kABytes + kBBytes) + (2 * STAGES + 4) * 8 + 16 + 1024 + EW * kStageBytes + 16 + 512;
};

template <int BN, int MT, int STAGES, int EW, bool INF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
    tc_conv_pair_kernel(const __grid_constant__ TcConvParams p, int n_ntiles, int n_groups, int n_z, int n_work) {
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  constexpr int kChunkGroups = EW / 4;
  extern __shared__ uint8_t smem_raw[];
  constexpr int kBBytes = (BN / 2) * 128;
  constexpr int kAStage = MT * kABytes;
  constexpr int kAccCols = MT * BN;
  constexpr int kTmemCols = 2 * kAccCols;
  static_assert(kTmemCols <= 512 && kTmemCols >= 32, "TMEM budget");
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kAStage;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * kBBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;       // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  // 512-byte aligned: the staging buffers' XOR pattern is then exactly TMA's 64-byte swizzle (absolute address bits)
  uint8_t* stage_base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 511) & ~(uintptr_t)511);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.w_map);
    ptx::prefetch_tmap(&p.in_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 2 * EW);   // every epilogue warp of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc_pair<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();                         // barriers of BOTH CTAs exist before anyone signals them
  ptx::tc_fence_after();
  vg::pdl_entry();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; nothing above touches global memory
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int w, int& z, int& nt, int& g) {
    const int per_z = n_ntiles * n_groups;
    z = w / per_z;
    const int r = w - z * per_z;
    nt = r / n_groups;
    g = r - nt * n_groups;
  };
  auto k_range = [&](int z, const TcPhase& ph, int& kb0, int& kb1) {
    const int total = ph.ntaps * p.k_chunks;
    if (p.ksplit > 1) { kb0 = z * p.k_per_split; kb1 = min(total, kb0 + p.k_per_split); }
    else { kb0 = 0; kb1 = total; }
    if (kb1 < kb0) kb1 = kb0;
  };
  auto tile_origin = [&](int t, int& x0, int& y0, int& n0) {
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int tn = t / p.tiles_y;
    x0 = tx * p.TW; y0 = ty * p.TH; n0 = tn * p.TN;
  };

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        int z, nt, g;
        decode(w, z, nt, g);
        const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
        int kb0, kb1;
        k_range(z, ph, kb0, kb1);
        if (kb1 == kb0) continue;
        const int tile0 = (g * 2 + (int)rank) * MT;      // this CTA's pixel tiles of the pair's group
        int x0s[MT], y0s[MT], n0s[MT];
#pragma unroll
        for (int m = 0; m < MT; ++m) tile_origin(tile0 + m, x0s[m], y0s[m], n0s[m]);
        const int nrow0 = nt * BN + (int)rank * (BN / 2);  // this CTA's half of the weight tile
        int tp = kb0 / p.k_chunks, kc = kb0 % p.k_chunks;
        for (int i = kb0; i < kb1; ++i) {
          const TcTap tap = ph.taps[tp];
          const CUtensorMap* im = &p.in_maps[tap.map];
          ptx::mbar_wait(&empty[stage], phase ^ 1u);
          if (ptx::elect_one()) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * (kAStage + kBBytes));
            const uint32_t bar = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
#pragma unroll
            for (int m = 0; m < MT; ++m)
              ptx::tma_load_4d_pair(sA + stage * kAStage + m * kABytes, im, bar, kc * 64, x0s[m] + tap.dx, y0s[m] + tap.dy, n0s[m]);
            ptx::tma_load_2d_pair(sB + stage * kBBytes, &p.w_map, bar, kc * 64, tap.wrow + nrow0);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          if (++kc == p.k_chunks) { kc = 0; ++tp; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, BN, 0, 0);
      constexpr uint32_t kDescHi = ptx::smem_desc_hi(1024);
      const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sA), 16), b_lo0 = ptx::smem_desc_lo(ptx::smem_u32(sB), 16);
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        int z, nt, g;
        decode(w, z, nt, g);
        const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
        int kb0, kb1;
        k_range(z, ph, kb0, kb1);
        if (kb1 == kb0) continue;
        const int as = it & 1;
        ptx::mbar_wait(&acc_empty[as], (uint32_t)(((it >> 1) & 1) ^ 1));
        ptx::tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(as * kAccCols);
        for (int i = kb0; i < kb1; ++i) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::mma_bf16_ss_pair_lohi(acc + (uint32_t)(m * BN), a_lo + (uint32_t)((m * kABytes + k * 32) >> 4),
                                           b_lo + (uint32_t)((k * 32) >> 4), kDescHi, idesc, (i > kb0) || (k > 0));
            }
            ptx::mma_commit_pair(&empty[stage], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; a_lo = a_lo0; b_lo = b_lo0; }
          else { a_lo += (uint32_t)(kAStage >> 4); b_lo += (uint32_t)(kBBytes >> 4); }
        }
        if (ptx::elect_one()) ptx::mma_commit_pair(&acc_full[as], 3);
        __syncwarp();
        ++it;
      }
    }
    vg::pdl_tail_trigger();
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = ew & 3;
    const int cgrp = ew >> 2;
    const int row = q * 32 + lane;
    const int xl = row & (p.TW - 1);
    const int yl = (row >> p.tw_log2) & (p.TH - 1);
    const int nl = row >> (p.tw_log2 + p.th_log2);
    // first pixel of this warp's 32-row quarter inside the tile: origin of its TMA-store sub-box
    const int row0 = q * 32;
    const int xl0 = row0 & (p.TW - 1), yl0 = (row0 >> p.tw_log2) & (p.TH - 1), nl0 = row0 >> (p.tw_log2 + p.th_log2);
    int it = 0;
    // BatchNorm statistics of the stored output, accumulated per lane (= column of the chunk) across this CTA's work
    // items and flushed with fp64 atomics whenever the N tile changes
    const bool do_stats = p.stats != nullptr && p.ksplit <= 1;
    float st_s[BN / 32], st_q[BN / 32];
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) st_s[c] = st_q[c] = 0.f;
    int st_nt = -1;
    auto flush_stats = [&]() {
      if (st_nt < 0) return;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        if ((c % kChunkGroups) != cgrp) continue;
        const int col = st_nt * BN + c * 32 + lane;
        if (col < p.n_out) {
          atomicAdd(p.stats + col, (double)st_s[c]);
          atomicAdd(p.stats + p.n_out + col, (double)st_q[c]);
        }
        st_s[c] = st_q[c] = 0.f;
      }
    };
    for (int w = pair; w < n_work; w += n_pairs) {
      int z, nt, g;
      decode(w, z, nt, g);
      const TcPhase& ph = p.phases[p.ksplit > 1 ? 0 : z];
      int kb0, kb1;
      k_range(z, ph, kb0, kb1);
      const bool has_k = kb1 > kb0;
      if (!has_k && p.ksplit > 1) continue;
      const int tile0 = (g * 2 + (int)rank) * MT;
      const int ncol0 = nt * BN;
      const int as = it & 1;
      if (do_stats && nt != st_nt) {
        flush_stats();
        st_nt = nt;
      }
      if (has_k) {
        ptx::mbar_wait(&acc_full[as], (uint32_t)((it >> 1) & 1));
        ptx::tc_fence_after();
      }
      const uint32_t acc = tmem_base + (uint32_t)(as * kAccCols) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        int x0, y0, n0;
        tile_origin(tile0 + m, x0, y0, n0);
        const int gx = x0 + xl, gy = y0 + yl, gn = n0 + nl;
        const bool valid = gx < p.gw && gy < p.gh && gn < p.gn;
        const long long opix = ((long long)gn * p.OH + (long long)gy * p.os + ph.oy_off) * p.OW + (long long)gx * p.os + ph.ox_off;
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          if ((c % kChunkGroups) != cgrp) continue;
          const int c0 = c * 32;
          uint32_t r[32];
          if (has_k) {
            ptx::tmem_ld_32x32(acc + (uint32_t)(m * BN + c0), r);
            ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          epilogue_chunk<INF>(p, r, valid, p.ksplit <= 1 || z == 0, opix, gn, ncol0 + c0, c == 0, do_stats && has_k, lane, st_s[c], st_q[c],
                         stage_base + ew * kStageBytes, x0 + xl0, y0 + yl0, n0 + nl0, z);
        }
      }
      if (has_k) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&acc_empty[as]), 0));
        ++it;
      }
    }
    if (do_stats) flush_stats();
    if (!INF && p.tma_store && lane == 0) ptx::bulk_wait_group0();      // this warp's bulk stores have landed
  }
  // neither CTA may retire while the other can still signal its barriers or read its shared memory
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc_pair<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------
template <int NB, int STAGES>
struct WgradSmem {
  static constexpr int kStageBytes = (2 + NB) * 8192;
  static constexpr int kBytes = STAGES * kStageBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int NB, int STAGES>
__global__ void __launch_bounds__(kTcThreads) tc_wgrad_kernel(const __grid_constant__ TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int BN = NB * 64;
  constexpr int kStageBytes = (2 + NB) * 8192;
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slabs = p.ntaps * p.cs_chunks;
  const int slab0 = blockIdx.x * 2;
  const bool has2 = slab0 + 1 < n_slabs;
  const int slab1 = has2 ? slab0 + 1 : slab0;
  const int box_begin = blockIdx.z * p.boxes_per_split;
  const int box_end = min(p.n_boxes, box_begin + p.boxes_per_split);
  const int num_k = max(0, box_end - box_begin);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.u_map);
    ptx::prefetch_tmap(&p.s_maps[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  vg::pdl_entry();   // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; nothing above touches global memory
  const uint32_t tmem_base = *tmem_slot;

  if (num_k > 0) {
    // role warps loop warp-uniformly; one elected lane issues the TMA / MMA instructions (see tc_conv_kernel)
    if (warp == 0) {
      const int tapA = slab0 / p.cs_chunks, chA = slab0 % p.cs_chunks;
      const int tapB = slab1 / p.cs_chunks, chB = slab1 % p.cs_chunks;
      const TcTap ta = p.taps[tapA], tb = p.taps[tapB];
      int stage = 0;
      uint32_t phase = 0;
      for (int b = box_begin; b < box_end; ++b) {
        int t = b;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int tn = t / p.tiles_y;
        const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = tn * p.TN;
        ptx::mbar_wait(&empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full[stage], kStageBytes);
          uint8_t* st = smem + stage * kStageBytes;
          ptx::tma_load_4d(st, &p.s_maps[ta.map], &full[stage], chA * 64, x0 + ta.dx, y0 + ta.dy, n0);
          ptx::tma_load_4d(st + 8192, &p.s_maps[tb.map], &full[stage], chB * 64, x0 + tb.dx, y0 + tb.dy, n0);
#pragma unroll
          for (int j = 0; j < NB; ++j)
            ptx::tma_load_4d(st + (2 + j) * 8192, &p.u_map, &full[stage], (blockIdx.y * NB + j) * 64, x0, y0, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128, BN, 1, 1);
      constexpr uint32_t kDescHi = ptx::smem_desc_hi(1024);
      const uint32_t a_lo0 = ptx::smem_desc_lo(ptx::smem_u32(smem), 8192);   // MN-major: LBO = 64-channel slab stride
      uint32_t a_lo = a_lo0;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_k; ++i) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)     // 16 pixels (two 8-row atoms) per MMA
            ptx::mma_bf16_ss_lohi(tmem_base, a_lo + (uint32_t)((k * 2048) >> 4), a_lo + (uint32_t)((2 * 8192 + k * 2048) >> 4), kDescHi,
                                  idesc, (i | k) != 0);
          ptx::mma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; a_lo = a_lo0; }
        else a_lo += (uint32_t)(kStageBytes >> 4);
      }
      if (ptx::elect_one()) ptx::mma_commit(tmem_full);
      __syncwarp();
      vg::pdl_tail_trigger();
    }
  }

  int* const det_lock = p.det_locks != nullptr ? p.det_locks + (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (warp >= 4 && det_lock) {
    if (lane == 0) det_wait_turn(det_lock, (int)blockIdx.z);
    __syncwarp();
  }
  if (warp >= 4 && num_k > 0) {
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int slab = (row < 64) ? slab0 : slab1;
    const bool row_valid = (row < 64) || has2;
    const int tap = slab / p.cs_chunks, chunk = slab % p.cs_chunks;
    const int cs = chunk * 64 + (row & 63);
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (row_valid) {
        const int cu0 = blockIdx.y * BN + c0;
        if (p.packed) {
          // 32 consecutive c_u of one (tap, c_s) row: eight 16-byte vector reductions
          float* dst = p.dw + (long long)blockIdx.z * p.det_split_stride + ((long long)tap * p.cs + cs) * p.cu + cu0;
          if (p.det_split_stride != 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                                __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j])),
                           "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                           : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float* dst = p.dw + (long long)blockIdx.z * p.det_split_stride + ((long long)(cu0 + j) * p.cs + cs) * p.ntaps + tap;
            if (p.det_split_stride != 0) *dst = __uint_as_float(r[j]);
            else atomicAdd(dst, __uint_as_float(r[j]));
          }
        }
      }
    }
  }
  if (warp >= 4 && det_lock) {
    __threadfence();
    asm volatile("bar.sync 2, 128;" ::: "memory");        // the four epilogue warps
    if (warp == 4 && lane == 0) det_publish_turn(det_lock, blockIdx.z + 1 == gridDim.z ? 0 : (int)blockIdx.z + 1);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
}

// dw[cu][cs][tap] += ws[tap][cs][cu]  (32x32 smem-tile transpose between cu and r = cs*taps+tap)
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const float* __restrict__ ws, int cu_n, int cs_n, int taps, float* __restrict__ dw) {
  vg::pdl_entry();
  __shared__ float tile[32][33];
  const int R = cs_n * taps;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, cu = c0 + tx;
    float v = 0.f;
    if (r < R && cu < cu_n) {
      const int cs = r / taps, tap = r - cs * taps;
      v = ws[((long long)tap * cs_n + cs) * cu_n + cu];
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int cu = c0 + i, r = r0 + tx;
    if (cu < cu_n && r < R) dw[(long long)cu * R + r] += tile[tx][i];
  }
}

// ==========================================================================================
// host side: tensor maps, tap tables, launch
// ==========================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

// cuTensorMapEncodeTiled is a DRIVER entry point and needs a current context on the calling
// thread.  torch's autograd worker threads only select a runtime device; a (cheap, once per
// thread) runtime call binds the primary context before the first encode.
static void bind_context_once() {
  static thread_local bool done = false;
  if (!done) {
    cudaFree(0);
    done = true;
  }
}

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  return g_encode;
}

// NHWC bf16 activation [n][h][w][c]; parity (py,px) with step `s` selects rows py, py+s, ... and
// columns px, px+s, ...  Box = {64, TW, TH, TN}.
static int make_act_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int s, int py, int px, int TW, int TH,
                        int TN) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return VG_ECUDA;
  }
  bind_context_once();
  const int hs = (h - py + s - 1) / s, ws = (w - px + s - 1) / s;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)std::max(ws, 1), (cuuint64_t)std::max(hs, 1), (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)s * c * 2, (cuuint64_t)s * w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  void* addr = (void*)((const char*)base + ((size_t)py * w + px) * c * 2);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, addr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation) failed: %d (n=%d h=%d w=%d c=%d s=%d box=%d,%d,%d)", (int)r, n, h, w, c, s, TW, TH, TN);
    return VG_ECUDA;
  }
  return VG_OK;
}

// bf16 output tensor (N, OH, OW, C) as {32 channels, bx, by, bn} boxes with the 64-byte swizzle (TMA-store epilogue)
// (s, py, px): the sub-grid of output pixels (py + s*i, px + s*j) a scatter phase writes (s = 1: the whole tensor)
static int make_out_map(CUtensorMap* m, void* base, int n, int h, int w, int c, int bx, int by, int bn, int s = 1, int py = 0,
                        int px = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return VG_ECUDA;
  }
  bind_context_once();
  const int hs = (h - py + s - 1) / s, ws = (w - px + s - 1) / s;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)std::max(ws, 1), (cuuint64_t)std::max(hs, 1), (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)s * c * 2, (cuuint64_t)s * w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  base = (void*)((char*)base + ((size_t)py * w + px) * c * 2);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(output) failed: %d (n=%d h=%d w=%d c=%d box=%d,%d,%d)", (int)r, n, h, w, c, bx, by, bn);
    return VG_ECUDA;
  }
  return VG_OK;
}

static int make_weight_map(CUtensorMap* m, const void* base, long long rows, int k, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return VG_ECUDA;
  }
  bind_context_once();
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights) failed: %d (rows=%lld k=%d box_rows=%d)", (int)r, rows, k, box_rows);
    return VG_ECUDA;
  }
  return VG_OK;
}

static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// choose a {TW,TH,TN} box of `pixels` (power of two) GEMM rows minimising padded work
static void choose_box(int gw, int gh, int gn, int pixels, int* TW, int* TH, int* TN) {
  double best = 1e30;
  int bw = 1, bh = 1;
  for (int tw = 1; tw <= pixels; tw *= 2)
    for (int th = 1; tw * th <= pixels; th *= 2) {
      int tn = pixels / (tw * th);
      if (tw > 256 || th > 256 || tn > 256) continue;
      double padded = (double)cdiv(gw, tw) * tw * (double)cdiv(gh, th) * th * (double)cdiv(gn, tn) * tn;
      double score = padded - 1e-3 * tw - 1e-6 * th;   // ties: prefer wide boxes (longer contiguous runs)
      if (score < best) { best = score; bw = tw; bh = th; }
    }
  *TW = bw; *TH = bh; *TN = pixels / (bw * bh);
}

bool tc_conv_supported(const VgConvDesc* d, bool dgrad) {
  if (g_force_simt) return false;
  if (d->act_dtype != VG_BF16) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  if (d->kh != d->kw || d->kh * d->kw > 16) return false;
  const int c_red = dgrad ? d->c_out : d->c_in;
  const int n_out = dgrad ? d->c_in : d->c_out;
  // n_out == 1: the N tile is filled with the neighbouring taps' weight rows (finite garbage) and only
  // column 0 is stored - the reduction over 64+ input channels still runs on the tensor cores
  if (c_red % 64 != 0 || (n_out % 64 != 0 && n_out != 1)) return false;
  if (d->stride == 2 && ((d->transposed ? d->h_out : d->h_in) % 2 != 0 || (d->transposed ? d->w_out : d->w_in) % 2 != 0)) return false;
  return get_encode() != nullptr;
}

bool tc_wgrad_supported(const VgConvDesc* d) {
  if (g_force_simt) return false;
  if (d->act_dtype != VG_BF16) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  if (d->kh != d->kw || d->kh * d->kw > 16) return false;
  if (d->c_in % 64 != 0 || d->c_out % 64 != 0) return false;
  if (d->stride == 2 && ((d->transposed ? d->h_out : d->h_in) % 2 != 0 || (d->transposed ? d->w_out : d->w_in) % 2 != 0)) return false;
  return get_encode() != nullptr;
}

template <int BN, int MT, int STAGES, bool INF>
static int launch_conv_t(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  static std::once_flag once;          // autograd worker threads call in concurrently
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_conv_kernel<BN, MT, STAGES, INF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 ConvSmem<BN, MT, STAGES>::kBytes); });
  VG_CUDA(attr_err);
  grid.x = (unsigned)cdiv(grid.x, MT);
  vg::Launch(grid, kTcThreads, ConvSmem<BN, MT, STAGES>::kBytes, s)(tc_conv_kernel<BN, MT, STAGES, INF>, p);
  VG_LAUNCHED();
  return VG_OK;
}

template <int BN, int MT, int STAGES, int EW, int G, bool INF>
static int launch_conv_persist_t(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  static std::once_flag once;          // autograd worker threads call in concurrently
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_conv_persist_kernel<BN, MT, STAGES, EW, G, INF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 ConvPersistSmem<BN, MT, STAGES, EW>::kBytes); });
  VG_CUDA(attr_err);
  const int n_groups = (int)cdiv(grid.x, MT);
  const int n_ntiles = (int)grid.y;
  const int n_z = (int)grid.z;
  const long long n_work = (long long)n_groups * n_ntiles * n_z;
  const int ctas = (int)std::min<long long>(n_work, num_sms());
  vg::Launch(ctas, 128 + 32 * EW, ConvPersistSmem<BN, MT, STAGES, EW>::kBytes, s)(tc_conv_persist_kernel<BN, MT, STAGES, EW, G, INF>, 
      p, n_ntiles, n_groups, n_z, (int)n_work);
  VG_LAUNCHED();
  return VG_OK;
}

template <int BN, int MT, int STAGES, int EW, bool INF>
static int launch_conv_pair_t(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  static std::once_flag once;          // autograd worker threads call in concurrently
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_conv_pair_kernel<BN, MT, STAGES, EW, INF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 ConvPairSmem<BN, MT, STAGES, EW>::kBytes); });
  VG_CUDA(attr_err);
  const int n_groups = (int)(grid.x / (2 * MT));      // the caller checked divisibility
  const int n_ntiles = (int)grid.y;
  const int n_z = (int)grid.z;
  const long long n_work = (long long)n_groups * n_ntiles * n_z;
  const int pairs = (int)std::min<long long>(n_work, num_sms() / 2);
  vg::Launch(2 * pairs, 128 + 32 * EW, ConvPairSmem<BN, MT, STAGES, EW>::kBytes, s)(tc_conv_pair_kernel<BN, MT, STAGES, EW, INF>, 
      p, n_ntiles, n_groups, n_z, (int)n_work);
  VG_LAUNCHED();
  return VG_OK;
}

// the inference epilogue (activation / residual / second output) lives in its OWN instantiations: compiled into the
// training kernels it cost them 12 % (b256 step 59.8 -> 67.3 ms: registers at the launch-bound cap, spills in the epilogue)
template <int BN, int MT, int STAGES>
static int launch_conv(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  const bool inf = p.act_slope != 1.0f || p.residual != nullptr || p.out2 != nullptr;
  return inf ? launch_conv_t<BN, MT, STAGES, true>(p, grid, s) : launch_conv_t<BN, MT, STAGES, false>(p, grid, s);
}
template <int BN, int MT, int STAGES, int EW, int G>
static int launch_conv_persist(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  const bool inf = p.act_slope != 1.0f || p.residual != nullptr || p.out2 != nullptr;
  return inf ? launch_conv_persist_t<BN, MT, STAGES, EW, G, true>(p, grid, s) : launch_conv_persist_t<BN, MT, STAGES, EW, G, false>(p, grid, s);
}
template <int BN, int MT, int STAGES, int EW>
static int launch_conv_pair(const TcConvParams& p, dim3 grid, cudaStream_t s) {
  const bool inf = p.act_slope != 1.0f || p.residual != nullptr || p.out2 != nullptr;
  return inf ? launch_conv_pair_t<BN, MT, STAGES, EW, true>(p, grid, s) : launch_conv_pair_t<BN, MT, STAGES, EW, false>(p, grid, s);
}

template <int NB, int STAGES>
static int launch_wgrad(const TcWgradParams& p, dim3 grid, cudaStream_t s) {
  static std::once_flag once;          // autograd worker threads call in concurrently
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(tc_wgrad_kernel<NB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<NB, STAGES>::kBytes); });
  VG_CUDA(attr_err);
  vg::Launch(grid, kTcThreads, WgradSmem<NB, STAGES>::kBytes, s)(tc_wgrad_kernel<NB, STAGES>, p);
  VG_LAUNCHED();
  return VG_OK;
}

// Build the tap table(s).  `gather` : out grid coarse (or same), input fine via parity maps when
// stride 2.  `scatter`: out grid fine, decomposed into s*s phases over the coarse input grid.
// flip=false: weight slab index = ky*kw+kx.
static void build_gather_taps(TcPhase* ph, int k, int stride, int pad, int n_out) {
  ph->ntaps = 0; ph->oy_off = 0; ph->ox_off = 0;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      TcTap& t = ph->taps[ph->ntaps++];
      const int ty = ky - pad, tx = kx - pad;
      if (stride == 1) {
        t.map = 0; t.dy = (int8_t)ty; t.dx = (int8_t)tx;
      } else {
        const int py = ((ty % 2) + 2) % 2, px = ((tx % 2) + 2) % 2;
        t.map = (int8_t)(py * 2 + px);
        t.dy = (int8_t)((ty - py) / 2);
        t.dx = (int8_t)((tx - px) / 2);
      }
      t.pad_ = 0;
      t.wrow = (ky * k + kx) * n_out;
    }
}
static void build_scatter_phase(TcPhase* ph, int k, int stride, int pad, int n_out, int phy, int phx) {
  ph->ntaps = 0; ph->oy_off = phy; ph->ox_off = phx;
  for (int ky = 0; ky < k; ++ky) {
    if (((phy + pad - ky) % stride + stride) % stride != 0) continue;
    for (int kx = 0; kx < k; ++kx) {
      if (((phx + pad - kx) % stride + stride) % stride != 0) continue;
      TcTap& t = ph->taps[ph->ntaps++];
      t.map = 0;
      // input index = q + (ph + pad - k)/stride (exact division, may be negative)
      t.dy = (int8_t)((phy + pad - ky) / stride);
      t.dx = (int8_t)((phx + pad - kx) / stride);
      t.pad_ = 0;
      t.wrow = (ky * k + kx) * n_out;
    }
  }
}

// ---- tile table: per (layer shape, batch, direction) the N tile and the kernel form that an
// autotune run measured fastest (vae_gan_b200/tune.py fills it through vg_conv_tune_set; vae_gan_b200/tile_table.json is the
// table measured on B200 for the BASELINE configurations).  Shapes without an entry use the heuristics below.
enum TuneForm { kFormAuto = 0, kFormOneShot = 1, kFormPersist = 2, kFormPair = 3 };
using TuneKey = std::array<int, 10>;   // dgrad, n, h_in, w_in, c_in, c_out, k, stride, pad, transposed
struct TuneVal { int bn, form; };
static std::map<TuneKey, TuneVal> g_tune;
static std::vector<TuneKey> g_tune_seen;
static bool g_tune_record = false;
static std::mutex g_tune_mu;
static TuneKey tune_key(const VgConvDesc* d, bool dgrad) {
  return TuneKey{dgrad ? 1 : 0, d->n, d->h_in, d->w_in, d->c_in, d->c_out, d->kh, d->stride, d->pad, d->transposed ? 1 : 0};
}
static TuneVal tune_lookup(const VgConvDesc* d, bool dgrad) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  const TuneKey k = tune_key(d, dgrad);
  if (g_tune_record && std::find(g_tune_seen.begin(), g_tune_seen.end(), k) == g_tune_seen.end()) g_tune_seen.push_back(k);
  auto it = g_tune.find(k);
  return it == g_tune.end() ? TuneVal{0, kFormAuto} : it->second;
}
int tc_tune_set(const VgConvDesc* d, int dgrad, int bn, int form) {
  VG_CHECK_ARG(d != nullptr, "null descriptor");
  VG_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "N tile must be 0 (heuristic), 64, 128 or 256 (got %d)", bn);
  VG_CHECK_ARG(form >= kFormAuto && form <= kFormPair, "kernel form must be 0..3 (got %d)", form);
  std::lock_guard<std::mutex> lk(g_tune_mu);
  if (bn == 0 && form == kFormAuto) g_tune.erase(tune_key(d, dgrad != 0));
  else g_tune[tune_key(d, dgrad != 0)] = TuneVal{bn, form};
  return VG_OK;
}
void tc_tune_clear() {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  g_tune.clear();
}
void tc_tune_record(int on) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  g_tune_record = on != 0;
  if (on) g_tune_seen.clear();
}
int tc_tune_seen(int* keys, int max_keys) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  const int n = (int)std::min<size_t>(g_tune_seen.size(), (size_t)std::max(0, max_keys));
  for (int i = 0; i < n && keys != nullptr; ++i)
    for (int j = 0; j < 10; ++j) keys[i * 10 + j] = g_tune_seen[i][j];
  return (int)g_tune_seen.size();
}

static int pick_bn(int n_out) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("VG_TC_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced && n_out % forced == 0) return forced;
  if (n_out % 256 == 0) return 256;
  if (n_out % 128 == 0) return 128;
  return 64;
}

// dgrad=false: y = conv(x).  dgrad=true: dx from dy.  `in` is the tensor being read.
int tc_conv_run(const VgConvDesc* d, bool dgrad, const void* in, const void* wpack, const float* bias, const float* colscale,
                const float* sigma, int sigma_group_n, void* out, int out_dtype, double* stats, bool* stats_fused, cudaStream_t s,
                const VgConvEpilogue* ep) {
  if (stats_fused) *stats_fused = false;
  TcConvParams p;
  memset(&p, 0, sizeof(p));
  // pattern: which side is the coarse grid
  //   Conv2d fwd           : gather  (in = x fine, out = y coarse)
  //   Conv2d dgrad         : scatter (in = dy coarse, out = dx fine)
  //   ConvTranspose2d fwd  : scatter (in = x coarse, out = y fine)
  //   ConvTranspose2d dgrad: gather  (in = dy fine, out = dx coarse)
  const bool gather = (d->transposed != 0) == dgrad;
  const int in_h = dgrad ? d->h_out : d->h_in, in_w = dgrad ? d->w_out : d->w_in, in_c = dgrad ? d->c_out : d->c_in;
  const int out_h = dgrad ? d->h_in : d->h_out, out_w = dgrad ? d->w_in : d->w_out, n_out = dgrad ? d->c_in : d->c_out;
  const int k = d->kh, st = d->stride, pad = d->pad;
  p.k_chunks = in_c / 64;
  p.n_out = n_out;
  p.OH = out_h; p.OW = out_w;
  p.out = out; p.out_f32 = out_dtype == VG_F32;
  p.bias = bias; p.colscale = colscale;
  p.sigma = sigma; p.sigma_group_n = sigma_group_n;
  p.act_slope = 1.0f; p.post_slope = 1.0f;
  if (ep != nullptr) {
    p.act_slope = ep->act_slope; p.residual = ep->residual; p.out2 = ep->y2;
    p.post_scale = ep->post_scale; p.post_shift = ep->post_shift; p.post_slope = ep->post_slope;
  }
  p.stats = nullptr;
  int nphase = 1;
  if (gather) {
    p.gw = out_w; p.gh = out_h; p.gn = d->n; p.os = 1;
    build_gather_taps(&p.phases[0], k, st, pad, n_out);
  } else {
    p.os = st;
    nphase = st * st;
    p.gw = (out_w + st - 1) / st; p.gh = (out_h + st - 1) / st; p.gn = d->n;
    for (int phy = 0; phy < st; ++phy)
      for (int phx = 0; phx < st; ++phx) build_scatter_phase(&p.phases[phy * st + phx], k, st, pad, n_out, phy, phx);
  }
  choose_box(p.gw, p.gh, p.gn, 128, &p.TW, &p.TH, &p.TN);
  p.tw_log2 = ilog2(p.TW); p.th_log2 = ilog2(p.TH);
  p.tiles_x = (int)cdiv(p.gw, p.TW); p.tiles_y = (int)cdiv(p.gh, p.TH); p.tiles_n = (int)cdiv(p.gn, p.TN);
  int rc;
  if (gather && st == 2) {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if ((rc = make_act_map(&p.in_maps[py * 2 + px], in, d->n, in_h, in_w, in_c, 2, py, px, p.TW, p.TH, p.TN))) return rc;
  } else {
    if ((rc = make_act_map(&p.in_maps[0], in, d->n, in_h, in_w, in_c, 1, 0, 0, p.TW, p.TH, p.TN))) return rc;
    p.in_maps[1] = p.in_maps[2] = p.in_maps[3] = p.in_maps[0];
  }
  const TuneVal tuned = tune_lookup(d, dgrad);
  const int BN = (n_out == 1) ? 64 : ((tuned.bn != 0 && n_out % tuned.bn == 0) ? tuned.bn : pick_bn(n_out));
  p.n_store = (n_out == 1) ? 1 : BN;
  if ((rc = make_weight_map(&p.w_map, wpack, (long long)k * k * n_out, in_c, BN))) return rc;
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * p.tiles_n), (unsigned)std::max(1, n_out / BN), (unsigned)nphase);
  if (grid.x == 0) return VG_OK;
  if (ep != nullptr && (ep->act_slope != 1.0f || ep->residual != nullptr || ep->y2 != nullptr) && (p.out_f32 || p.n_store == 1)) {
    set_error("the activation / residual / second-output epilogue needs a bf16 output with c_out %% 64 == 0");
    return VG_EUNSUPPORTED;
  }
  p.ksplit = 1;
  p.k_per_split = p.phases[0].ntaps * p.k_chunks;
  {
    // few output tiles but a long reduction (the discriminator's Linear layers run here as 1x1
    // convolutions on [B,1,1,C] tensors): split K over blockIdx.z, fp32 atomics into a zeroed output
    const long long ctas = (long long)grid.x * grid.y;
    const int total_kb = p.phases[0].ntaps * p.k_chunks;
    if (nphase == 1 && p.out_f32 && colscale == nullptr && ctas * 2 <= num_sms() && total_kb >= 16) {
      int ks = (int)std::min<long long>(cdiv(2LL * num_sms(), ctas), total_kb / 4);
      if (ks > 1) {
        p.k_per_split = (int)cdiv(total_kb, ks);
        p.ksplit = (int)cdiv(total_kb, p.k_per_split);
        grid.z = (unsigned)p.ksplit;
        const size_t out_elems = (size_t)d->n * out_h * out_w * n_out;
        VG_CUDA(cudaMemsetAsync(out, 0, out_elems * sizeof(float), s));
        if (g_det.on) {
          // deterministic mode: every split stores its partial tile into its own scratch slab and ordered_reduce_f32 adds
          // the slabs in split order (below); if the caller's scratch is too small the splits take turns instead
          if ((size_t)p.ksplit * out_elems * sizeof(float) <= g_det.scratch_bytes) {
            p.out = g_det.scratch;
            p.det_split_stride = (long long)out_elems;
          } else if (!(p.det_locks = det_locks(ctas))) {
            return VG_EINVAL;
          }
        }
      }
    }
  }
  // kernel variant: persistent CTAs (double-buffered accumulator, MT pixel tiles share each staged weight
  // tile) when there is enough work; one-shot CTAs (two resident per SM) for small problems and split-K.
  // Measured on B200 (scripts/sweep_conv.py): a persistent single-tile CTA loses to two co-resident
  // one-shot CTAs, and multi-tile CTAs without the double-buffered accumulator are slower still
  // (128->128 @96: 897 -> 769 TFLOP/s), so those stay opt-in (VG_TC_MT=1).
  // TMA-store epilogue (VG_TC_TMA_STORE, default 2): bf16 outputs without split-K and without the inference epilogue leave
  // the staging buffers as cp.async.bulk.tensor stores - 1: gather-type launches only (output on the GEMM's own pixel grid),
  // 2: also the scatter launches (one strided output map per phase), 0: the manual coalesced stores.  Measured on B200
  // (profiles/r2_tma_store_ab.txt): 60.2 -> 59.2 ms/step at batch 256, 10.58 -> 10.36 at 32; 64->64 @96 74 -> 67 us,
  // 64->128 @96 110 -> 97 us, 1x1 stride-2 51 -> 43 us; the scatter phases add another 0.3 %.
  static int tma_store = -1;
  if (tma_store < 0) { const char* e = getenv("VG_TC_TMA_STORE"); tma_store = e ? atoi(e) : 2; }
  if (tma_store && !p.out_f32 && p.ksplit == 1 && p.n_store != 1 && (p.os == 1 || tma_store >= 2) &&
      p.act_slope == 1.0f && p.residual == nullptr && p.out2 == nullptr) {
    const int bx = std::min(p.TW, 32), by = std::min(p.TH, 32 / bx), bn = 32 / (bx * by);
    // a scatter launch (ConvTranspose2d forward, stride-2 dgrad) writes st x st interleaved sub-grids, one per phase (grid.z)
    for (int ph = 0; ph < nphase; ++ph)
      if ((rc = make_out_map(&p.out_maps[ph], out, d->n, out_h, out_w, n_out, bx, by, bn, p.os, p.phases[ph].oy_off, p.phases[ph].ox_off)))
        return rc;
    p.tma_store = 1;
  }
  static int mt_on = -1, persist = -1, mt_min = -1, fuse_stats = -1;
  if (mt_on < 0) { const char* e = getenv("VG_TC_MT"); mt_on = (e && atoi(e)) ? 1 : 0; }
  if (persist < 0) { const char* e = getenv("VG_TC_PERSIST"); persist = e ? atoi(e) : 1; }
  if (mt_min < 0) { const char* e = getenv("VG_TC_MT_MIN"); mt_min = e ? atoi(e) : 6; }
  if (fuse_stats < 0) { const char* e = getenv("VG_TC_FUSE_STATS"); fuse_stats = e ? atoi(e) : 0; }
  const long long ctas1 = (long long)grid.x * grid.y * grid.z;
  const bool big = mt_on && p.ksplit == 1 && ctas1 >= 6LL * num_sms();
  // 256-wide tiles go persistent as soon as there is more than one work item per SM (288 tiles of the 24x24 layers at
  // batch 64 ran as one-shot CTAs without epilogue overlap: 38-85 us for 31 us of tensor work)
  const bool heur_persist = persist && !(g_det.on && p.ksplit > 1) &&       // deterministic split-K: the one-shot kernel takes turns
                           ((BN == 256 && ctas1 > (long long)num_sms()) ||
                                       // 128-wide: the CTA-pair kernel (1000+ TFLOP/s) from two work items per pair on; below 6 x SMs tiles
                                       // these layers ran as one-shot CTAs (G 128->128 @48 at 32 images: 496 TFLOP/s)
                                       (BN == 128 && ctas1 >= 2LL * num_sms()) ||
                                       (BN == 64 && ctas1 >= 2LL * num_sms() && ctas1 >= (long long)mt_min * num_sms()));
  // a tile-table entry overrides the form (split-K launches keep the heuristic: only the one-shot kernel was tuned for them)
  const int form = p.ksplit > 1 ? kFormAuto : tuned.form;
  const bool use_persist = form == kFormAuto ? heur_persist : (form != kFormOneShot);
  // BatchNorm statistics of the output can be accumulated by the epilogue warps of the 8-warp persistent
  // kernels (parity-tested), but it is OPT-IN (VG_TC_FUSE_STATS=1): measured on B200 the per-chunk smem
  // transposes cost the L2/smem-bound main loop more (+35 % per launch with 8 epilogue warps, +60 % with
  // 4) than the separate 20 us statistics kernel they replace.
  // Round 2 replaced the shared-memory transpose by a 31-shuffle recursive halving per 32 x 32 chunk and enabled it in
  // every kernel form - measured on B200 (scripts/sweep_conv.py, batch 64) it is STILL a loss: 64->64 @96 74.9 -> 163 us,
  // 128->128 @96 151 -> 217 us, 512->512 @24 110 -> 125 us, whole step 63.7 -> 68.2 ms, against the 17 us streaming
  // statistics kernel it replaces: the epilogue warps are the bottleneck of the short-reduction layers and ~350 extra
  // instructions per chunk double their work.  Kept opt-in (VG_TC_FUSE_STATS=1) and parity-tested.
  if (fuse_stats && !g_det.on && stats != nullptr && p.ksplit == 1 && n_out % 32 == 0 && p.n_store != 1) {
    p.stats = stats;
    if (stats_fused) *stats_fused = true;
  }
  // CTA pairs (cta_group::2) halve the per-SM shared-memory operand reads that bound the narrow tiles
  static int pair_on = -1;
  if (pair_on < 0) { const char* e = getenv("VG_TC_PAIR"); pair_on = e ? atoi(e) : 6; }   // bit0: BN=64, bit1: BN=128, bit2: BN=256
  if (use_persist && form != kFormPersist && (pair_on || form == kFormPair) && p.n_store != 1) {
    const int mt = BN == 64 ? 4 : (BN == 128 ? 2 : 1);
    if (grid.x % (2 * mt) == 0 && (form == kFormPair || (pair_on & (BN == 64 ? 1 : (BN == 128 ? 2 : 4))))) {
      if ((rc = make_weight_map(&p.w_map, wpack, (long long)k * k * n_out, in_c, BN / 2))) return rc;
      switch (BN) {
        case 64: return launch_conv_pair<64, 4, 3, 4>(p, grid, s);
        case 128: return launch_conv_pair<128, 2, 4, 8>(p, grid, s);
        case 256: return launch_conv_pair<256, 1, 6, 8>(p, grid, s);
        default: break;
      }
    }
  }
  static int grp2 = -1;
  if (grp2 < 0) { const char* e = getenv("VG_TC_GROUP"); grp2 = e ? (atoi(e) == 2) : 0; }
  // 64-wide tiles are EPILOGUE-bound (9 k-blocks of N=64 MMAs per 128 x 64 outputs).  Eight epilogue warps instead of four
  // (VG_TC_EW64=82: <64,2,4,8>) measured NO gain on B200 (64->64 @96: 74.4 vs 74.9 us): the bound was the store path
  // (32 lines per store instruction), addressed by the staged, coalesced stores of epilogue_chunk.
  static int ew64 = -1;
  if (ew64 < 0) { const char* e = getenv("VG_TC_EW64"); ew64 = e ? atoi(e) : 4; }
  if (use_persist) {
    switch (BN) {
      case 64:
        if (ew64 == 82) return launch_conv_persist<64, 2, 4, 8, 1>(p, grid, s);
        return launch_conv_persist<64, 4, 3, 4, 1>(p, grid, s);
      case 128: return grp2 ? launch_conv_persist<128, 2, 4, 8, 2>(p, grid, s) : launch_conv_persist<128, 2, 4, 8, 1>(p, grid, s);
      case 256: return grp2 ? launch_conv_persist<256, 1, 4, 8, 2>(p, grid, s) : launch_conv_persist<256, 1, 4, 8, 1>(p, grid, s);
      default: break;
    }
  }
  switch (BN) {
    case 64: rc = big ? launch_conv<64, 2, 4>(p, grid, s) : launch_conv<64, 1, 4>(p, grid, s); break;
    case 128: rc = big ? launch_conv<128, 2, 4>(p, grid, s) : launch_conv<128, 1, 3>(p, grid, s); break;
    case 256: rc = launch_conv<256, 1, 4>(p, grid, s); break;
    default: set_error("unsupported BN %d", BN); return VG_EUNSUPPORTED;
  }
  if (rc == VG_OK && p.det_split_stride != 0)
    rc = ordered_reduce_f32((const float*)p.out, p.ksplit, p.det_split_stride, (float*)out, s);
  return rc;
}

// acc_only: leave the result in `workspace` - packed [tap][c_s][c_u] for k > 1, torch layout [c_u][c_s] for k = 1 (the
// caller zeroed it and consumes it itself: vg_conv_wgrad_sn) - instead of adding it into dw
int tc_wgrad_run(const VgConvDesc* d, const void* x, const void* dy, float* dw, float* workspace, cudaStream_t s, bool acc_only) {
  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  // dw[cu][cs][tap] += sum_q U[q][cu] * S[q*stride - pad + k][cs]
  const void *U, *S;
  int hu, wu, cu, hs, ws, cs;
  if (!d->transposed) { U = dy; hu = d->h_out; wu = d->w_out; cu = d->c_out; S = x; hs = d->h_in; ws = d->w_in; cs = d->c_in; }
  else                { U = x;  hu = d->h_in;  wu = d->w_in;  cu = d->c_in;  S = dy; hs = d->h_out; ws = d->w_out; cs = d->c_out; }
  const int k = d->kh, st = d->stride, pad = d->pad;
  TcPhase tmp;
  build_gather_taps(&tmp, k, st, pad, 0);
  p.ntaps = tmp.ntaps;
  for (int i = 0; i < tmp.ntaps; ++i) p.taps[i] = tmp.taps[i];
  p.cs_chunks = cs / 64; p.cs = cs; p.cu = cu;
  choose_box(wu, hu, d->n, 64, &p.TW, &p.TH, &p.TN);
  p.tiles_x = (int)cdiv(wu, p.TW); p.tiles_y = (int)cdiv(hu, p.TH); p.tiles_n = (int)cdiv(d->n, p.TN);
  p.n_boxes = p.tiles_x * p.tiles_y * p.tiles_n;
  // 1x1 convolutions / Linear layers: torch layout [c_u][c_s] already has c_s contiguous, so the epilogue's per-lane
  // reductions are coalesced (32 lanes = 32 consecutive c_s) and the scratch + memset + unpack round trip is skipped
  static int direct11 = -1;
  if (direct11 < 0) { const char* e = getenv("VG_WGRAD_DIRECT_1X1"); direct11 = e ? atoi(e) : 1; }
  p.packed = workspace != nullptr && (tmp.ntaps > 1 || (!direct11 && !acc_only));
  p.dw = (p.packed || acc_only) ? workspace : dw;
  const size_t welems = (size_t)k * k * cs * cu;
  if (p.packed && !acc_only) VG_CUDA(cudaMemsetAsync(workspace, 0, welems * sizeof(float), s));
  int rc;
  if (st == 2) {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if ((rc = make_act_map(&p.s_maps[py * 2 + px], S, d->n, hs, ws, cs, 2, py, px, p.TW, p.TH, p.TN))) return rc;
  } else {
    if ((rc = make_act_map(&p.s_maps[0], S, d->n, hs, ws, cs, 1, 0, 0, p.TW, p.TH, p.TN))) return rc;
    p.s_maps[1] = p.s_maps[2] = p.s_maps[3] = p.s_maps[0];
  }
  if ((rc = make_act_map(&p.u_map, U, d->n, hu, wu, cu, 1, 0, 0, p.TW, p.TH, p.TN))) return rc;
  const int NB = (cu % 256 == 0) ? 4 : ((cu % 128 == 0) ? 2 : 1);
  const int m_tiles = (p.ntaps * p.cs_chunks + 1) / 2;
  const int n_tiles = cu / (NB * 64);
  if (p.n_boxes == 0) return VG_OK;
  // one CTA per SM (160-192 KB of smem): pick the split count that fills whole waves
  const long long tiles = (long long)m_tiles * n_tiles;
  long long splits = 1;
  {
    double best = -1.0;
    const long long max_splits = std::max<long long>(1, p.n_boxes / 4);
    for (long long waves = 1; waves <= 4; ++waves) {
      long long sp = std::max<long long>(1, std::min<long long>((waves * num_sms()) / tiles, max_splits));
      long long bps = cdiv(p.n_boxes, sp);
      sp = cdiv(p.n_boxes, bps);
      const long long ctas = tiles * sp;
      const double eff = (double)ctas / (double)(cdiv(ctas, num_sms()) * num_sms());   // wave quantisation
      const double score = eff - 0.02 * (double)sp / (double)max_splits - (ctas < num_sms() ? 0.5 : 0.0);
      if (score > best) { best = score; splits = sp; }
    }
  }
  p.boxes_per_split = (int)cdiv(p.n_boxes, splits);
  splits = cdiv(p.n_boxes, p.boxes_per_split);
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  float* const wgrad_dst = p.dw;
  if (g_det.on && splits > 1) {
    // deterministic mode: per-split scratch slabs + ordered reduce (see the forward's split-K), turn-taking as the fallback
    if ((size_t)splits * welems * sizeof(float) <= g_det.scratch_bytes) {
      p.dw = (float*)g_det.scratch;
      p.det_split_stride = (long long)welems;
    } else if (!(p.det_locks = det_locks(tiles))) {
      return VG_EINVAL;
    }
  }
  switch (NB) {
    case 1: rc = launch_wgrad<1, 6>(p, grid, s); break;
    case 2: rc = launch_wgrad<2, 5>(p, grid, s); break;
    default: rc = launch_wgrad<4, 4>(p, grid, s); break;
  }
  if (rc == VG_OK && p.det_split_stride != 0) rc = ordered_reduce_f32(p.dw, (int)splits, p.det_split_stride, wgrad_dst, s);
  if (rc || !p.packed || acc_only) return rc;
  dim3 ug((unsigned)cdiv((long long)cs * p.ntaps, 32), (unsigned)cdiv(cu, 32));
  vg::Launch(ug, 256, 0, s)(wgrad_unpack_kernel, workspace, cu, cs, p.ntaps, dw);
  VG_LAUNCHED();
  return VG_OK;
}

}  // namespace vg
