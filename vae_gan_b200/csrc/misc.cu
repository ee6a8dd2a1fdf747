// The remaining (memory- or latency-bound) pieces of the VAE-GAN step: Linear layers of the
// discriminator head, avg_pool2d+flatten, spectral-norm power iteration, reparameterisation,
// the fused loss+gradient kernels, the fused Adam/RMSprop update, and layout/dtype helpers.
#include <algorithm>
#include <string.h>
#include "vg_common.cuh"

namespace vg {

// ------------------------------------------------------------------------------------------
// generic small GEMM  C[i][j] (+)= sum_l A(i,l) * B(l,j)   (fp32 accumulate)
//   A(i,l) = A[i*sai + l*sal], B(l,j) = B[l*sbl + j*sbj].  32x32 tile, 256 threads, 2x2 micro.
//   mode 0: C = acc ; mode 1: C += acc (single writer) ; mode 2: atomicAdd (split-K)
// ------------------------------------------------------------------------------------------
template <typename TA, typename TB>
__global__ void __launch_bounds__(256) gemm_small_kernel(const TA* __restrict__ A, long long sai, long long sal,
                                                          const TB* __restrict__ B, long long sbl, long long sbj, int I, int J,
                                                          int L, int l_per_split, float* __restrict__ C, int mode,
                                                          int* __restrict__ det_locks) {
  vg::pdl_entry();
  __shared__ float sa[32][33];   // [l][i]
  __shared__ float sb[32][33];   // [l][j]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int lbeg = blockIdx.z * l_per_split, lend = min(L, lbeg + l_per_split);
  const int tid = threadIdx.x;
  const int ti = tid / 16, tj = tid % 16;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // loader index: fast index along whichever dimension is contiguous
  const bool a_l_fast = (sal == 1), b_l_fast = (sbl == 1);
  for (int lb = lbeg; lb < lend; lb += 32) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      int f = tid % 32, sl = tid / 32 + pass * 8;
      {
        int l = a_l_fast ? f : sl, i = a_l_fast ? sl : f;
        float v = 0.f;
        if (lb + l < lend && i0 + i < I) v = to_f32(A[(long long)(i0 + i) * sai + (long long)(lb + l) * sal]);
        sa[l][i] = v;
      }
      {
        int l = b_l_fast ? f : sl, j = b_l_fast ? sl : f;
        float v = 0.f;
        if (lb + l < lend && j0 + j < J) v = to_f32(B[(long long)(lb + l) * sbl + (long long)(j0 + j) * sbj]);
        sb[l][j] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      float a0 = sa[l][ti * 2], a1 = sa[l][ti * 2 + 1];
      float b0 = sb[l][tj * 2], b1 = sb[l][tj * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
  // deterministic mode: the reduction splits (blockIdx.z) of one output tile add in split order
  int* const det_lock = (det_locks && mode == 2) ? det_locks + (blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
  if (det_lock) {
    if (tid == 0) det_wait_turn(det_lock, (int)blockIdx.z);
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int i = i0 + ti * 2 + a, j = j0 + tj * 2 + b;
      if (i < I && j < J) {
        float* c = C + (long long)i * J + j;
        if (mode == 0) *c = acc[a][b];
        else if (mode == 1) *c += acc[a][b];
        else atomicAdd(c, acc[a][b]);
      }
    }
  if (det_lock) {
    __threadfence();
    __syncthreads();
    if (tid == 0) det_publish_turn(det_lock, blockIdx.z + 1 == gridDim.z ? 0 : (int)blockIdx.z + 1);
  }
}

__global__ void bias_lrelu_kernel(float* __restrict__ y, const float* __restrict__ bias, long long total, int n, float slope) {
  vg::pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = y[i] + (bias ? bias[i % n] : 0.f);
    y[i] = v > 0.f ? v : v * slope;
  }
}
__global__ void colsum_f32_kernel(const float* __restrict__ x, int rows, int c, float* __restrict__ out) {
  vg::pdl_entry();
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += x[(long long)r * c + ch];
  out[ch] += s;
}

template <typename TA, typename TB>
static int launch_gemm(const TA* A, long long sai, long long sal, const TB* B, long long sbl, long long sbj, int I, int J, int L,
                       float* C, int mode, int splits, cudaStream_t s) {
  int lps = (int)cdiv(cdiv(L, splits), 32) * 32;
  if (lps < 32) lps = 32;
  splits = (int)cdiv(L, lps);
  dim3 grid((unsigned)cdiv(J, 32), (unsigned)cdiv(I, 32), (unsigned)splits);
  int* locks = nullptr;
  if (g_det.on && mode == 2 && splits > 1 && !(locks = det_locks((long long)grid.x * grid.y))) return VG_EINVAL;
  vg::Launch(grid, 256, 0, s)(gemm_small_kernel<TA, TB>, A, sai, sal, B, sbl, sbj, I, J, L, lps, C, mode, locks);
  VG_LAUNCHED();
  return VG_OK;
}

// ------------------------------------------------------------------------------------------
// avg_pool2d(k) + flatten (NCHW order): out[n][ch*(ph*pw) + py*pw + px]  (fp32)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void avgpool_flatten_fwd_kernel(const T* __restrict__ x, int n, int h, int w, int c, int k, float* __restrict__ out) {
  vg::pdl_entry();
  const int ph = h / k, pw = w / k;
  const long long total = (long long)n * ph * pw * c;
  const float inv = 1.0f / (float)(k * k);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long long t = i / c;
    int px = (int)(t % pw); t /= pw;
    int py = (int)(t % ph);
    int nn = (int)(t / ph);
    float s = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) s += to_f32(x[(((long long)nn * h + py * k + dy) * w + px * k + dx) * c + ch]);
    out[(long long)nn * c * ph * pw + (long long)ch * ph * pw + py * pw + px] = s * inv;
  }
}
// one thread per 8 consecutive channels of an input pixel: 16-byte stores, 32-bit index math
template <typename T>
__global__ void avgpool_flatten_bwd_kernel(const float* __restrict__ dout, int n, int h, int w, int c, int k, T* __restrict__ dx) {
  vg::pdl_entry();
  const int ph = h / k, pw = w / k;
  const float inv = 1.0f / (float)(k * k);
  if ((c & 7) == 0) {
    const unsigned cg = (unsigned)c / 8;
    const unsigned total = (unsigned)n * h * w * cg;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
      const unsigned g = i % cg;
      unsigned t = i / cg;
      const int x = (int)(t % (unsigned)w); t /= (unsigned)w;
      const int y = (int)(t % (unsigned)h);
      const int nn = (int)(t / (unsigned)h);
      const int py = y / k, px = x / k;
      Vec8<T> o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = 0.f;
        if (py < ph && px < pw) v = dout[(long long)nn * c * ph * pw + (long long)(g * 8 + j) * ph * pw + py * pw + px] * inv;
        o.v[j] = v;
      }
      o.store(dx + (long long)i * 8);
    }
    return;
  }
  const long long total = (long long)n * h * w * c;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long long t = i / c;
    int x = (int)(t % w); t /= w;
    int y = (int)(t % h);
    int nn = (int)(t / h);
    int py = y / k, px = x / k;
    float v = 0.f;
    if (py < ph && px < pw) v = dout[(long long)nn * c * ph * pw + (long long)ch * ph * pw + py * pw + px] * inv;
    dx[i] = from_f32<T>(v);
  }
}

// Fast path (h % k == 0, w % k == 0, c % 8 == 0): one block per (image, pooled row).  The block gathers the row's pooled
// gradients dout[n][ch][py][0..pw) ONCE into shared memory as [px][ch] (the per-pixel kernel above re-gathered every value
// 16 times with stride-36 scalar loads: 80 us for a 38 MB output) and then writes its k input rows with 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(256) avgpool_flatten_bwd_rows_kernel(const float* __restrict__ dout, int h, int w, int c, int k,
                                                                       T* __restrict__ dx) {
  vg::pdl_entry();
  extern __shared__ float tile[];      // [pw][c]
  const int ph = h / k, pw = w / k;
  const int n = blockIdx.x / ph, py = blockIdx.x % ph;
  const float inv = 1.0f / (float)(k * k);
  const float* src = dout + (long long)n * c * ph * pw + py * pw;
  for (int idx = threadIdx.x; idx < c * pw; idx += blockDim.x) {
    const int ch = idx / pw, px = idx - ch * pw;
    tile[px * c + ch] = src[(long long)ch * ph * pw + px] * inv;
  }
  __syncthreads();
  const int cg = c / 8;
  const int total = k * w * cg;
  T* dst = dx + ((long long)n * h + (long long)py * k) * w * c;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int g = idx % cg;
    const int x = (idx / cg) % w;
    const float4* t4 = reinterpret_cast<const float4*>(tile + (x / k) * c + g * 8);
    const float4 a = t4[0], b = t4[1];
    Vec8<T> o;
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
    o.store(dst + (long long)idx * 8);
  }
}

// ------------------------------------------------------------------------------------------
// spectral norm
// ------------------------------------------------------------------------------------------
// t[j] += sum_{i in slice} W[i][j] u[i]
__global__ void sn_wt_u_kernel(const float* __restrict__ W, const float* __restrict__ u, int rows, int cols, int rows_per_slice,
                               float* __restrict__ t) {
  vg::pdl_entry();
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  int r0 = blockIdx.y * rows_per_slice, r1 = min(rows, r0 + rows_per_slice);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = r0;
  for (; i + 4 <= r1; i += 4) {      // four independent loads in flight per thread
    const float* wp = W + (long long)i * cols + j;
    s0 = fmaf(wp[0], u[i], s0);
    s1 = fmaf(wp[cols], u[i + 1], s1);
    s2 = fmaf(wp[2LL * cols], u[i + 2], s2);
    s3 = fmaf(wp[3LL * cols], u[i + 3], s3);
  }
  for (; i < r1; ++i) s0 = fmaf(W[(long long)i * cols + j], u[i], s0);
  atomicAdd(&t[j], (s0 + s1) + (s2 + s3));
}
// one 128-thread block per row: s[i] = (W[i] . t) / max(||t||, eps)   (normalize=1), or W[i] . t (normalize=0);
// the block of row 0 also writes v = t / max(||t||, eps)
__global__ void __launch_bounds__(128) sn_w_v_kernel(const float* __restrict__ W, const float* __restrict__ t, int rows, int cols,
                                                     int normalize, float eps, float* __restrict__ v_out, float* __restrict__ s_out) {
  vg::pdl_entry();
  const int row = blockIdx.x;
  float dot = 0.f, nt = 0.f;
  const float* wr = W + (long long)row * cols;
  for (int j = threadIdx.x; j < cols; j += 128) {
    float tv = t[j];
    dot = fmaf(wr[j], tv, dot);
    nt = fmaf(tv, tv, nt);
  }
  __shared__ float red[2][4];
  __shared__ float inv_s;
  dot = warp_sum(dot);
  nt = warp_sum(nt);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = nt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float d = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    const float n = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    const float inv = normalize ? 1.0f / fmaxf(sqrtf(n), eps) : 1.0f;
    s_out[row] = d * inv;
    inv_s = inv;
  }
  if (normalize && row == 0 && v_out != nullptr) {
    __syncthreads();
    const float inv = inv_s;
    for (int j = threadIdx.x; j < cols; j += 128) v_out[j] = t[j] * inv;
  }
}
// single block: training: u = s / max(||s||, eps); sigma = sum u*s.  eval: sigma = sum u*s.
__global__ void sn_finish_kernel(const float* __restrict__ s, float* __restrict__ u, int rows, int training, float eps,
                                 float* __restrict__ sigma) {
  vg::pdl_entry();
  __shared__ float red[32];
  __shared__ float bc;
  int tid = threadIdx.x;
  float part = 0.f;
  if (training) {
    for (int i = tid; i < rows; i += blockDim.x) part = fmaf(s[i], s[i], part);
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      bc = 1.0f / fmaxf(sqrtf(tot), eps);
    }
    __syncthreads();
    float inv = bc;
    for (int i = tid; i < rows; i += blockDim.x) u[i] = s[i] * inv;
    __syncthreads();
  }
  part = 0.f;
  for (int i = tid; i < rows; i += blockDim.x) part = fmaf(u[i], s[i], part);
  part = warp_sum(part);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    *sigma = tot;
  }
}

// ---- batched power iteration: every spectral-normed weight of the discriminator in three launches ----------------
// (per-weight it was memset + 3 kernels + 2 clone kernels, x 9 convolutions x 3 forwards per training iteration)
struct SnTable {
  VgSnItem it[VG_SN_MAX];
  int t_off[VG_SN_MAX];      // workspace offsets (floats): t[cols] then s[rows] per item
  int s_off[VG_SN_MAX];
  int n;
};
__global__ void sn_wt_u_batched_kernel(const __grid_constant__ SnTable tb, float* __restrict__ ws, int one_slice) {
  vg::pdl_entry();
  const VgSnItem& it = tb.it[blockIdx.z];
  const int rows = it.rows, cols = it.cols;
  const int slices = one_slice ? 1 : max(1, min(rows / 8, 64));     // deterministic mode: a single contributor per t[j]
  const int rps = (rows + slices - 1) / slices;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rps;
  if (j >= cols || r0 >= rows) return;
  const int r1 = min(rows, r0 + rps);
  const float* W = it.w;
  const float* u = it.u;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = r0;
  for (; i + 4 <= r1; i += 4) {
    const float* wp = W + (long long)i * cols + j;
    s0 = fmaf(wp[0], u[i], s0);
    s1 = fmaf(wp[cols], u[i + 1], s1);
    s2 = fmaf(wp[2LL * cols], u[i + 2], s2);
    s3 = fmaf(wp[3LL * cols], u[i + 3], s3);
  }
  for (; i < r1; ++i) s0 = fmaf(W[(long long)i * cols + j], u[i], s0);
  atomicAdd(&ws[tb.t_off[blockIdx.z] + j], (s0 + s1) + (s2 + s3));
}
// block (row, item): s[row] = (W[row] . t) / max(||t||, eps) (training; t = W^T u) or W[row] . v (eval);
// row 0 also writes v = t / max(||t||, eps) into the module buffer and into the saved copy
__global__ void __launch_bounds__(128) sn_w_v_batched_kernel(const __grid_constant__ SnTable tb, float* __restrict__ ws, int training, float eps) {
  vg::pdl_entry();
  const VgSnItem& it = tb.it[blockIdx.y];
  const int row = blockIdx.x;
  if (row >= it.rows) return;
  const int cols = it.cols;
  const float* t = training ? ws + tb.t_off[blockIdx.y] : it.v;
  float dot = 0.f, nt = 0.f;
  const float* wr = it.w + (long long)row * cols;
  for (int j = threadIdx.x; j < cols; j += 128) {
    const float tv = t[j];
    dot = fmaf(wr[j], tv, dot);
    nt = fmaf(tv, tv, nt);
  }
  __shared__ float red[2][4];
  __shared__ float inv_s;
  dot = warp_sum(dot);
  nt = warp_sum(nt);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = nt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float d = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    const float n = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    const float inv = training ? 1.0f / fmaxf(sqrtf(n), eps) : 1.0f;
    ws[tb.s_off[blockIdx.y] + row] = d * inv;
    inv_s = inv;
  }
  if (row == 0) {
    __syncthreads();
    const float inv = inv_s;
    for (int j = threadIdx.x; j < cols; j += 128) {
      const float vv = t[j] * inv;                  // eval: inv = 1, v unchanged
      if (training) it.v[j] = vv;
      if (it.v_out != nullptr) it.v_out[j] = vv;
    }
  }
}
// one block per item: training: u = s / max(||s||, eps); sigma = u . s; the saved copy of u
__global__ void sn_finish_batched_kernel(const __grid_constant__ SnTable tb, const float* __restrict__ ws, int training, float eps) {
  vg::pdl_entry();
  const VgSnItem& it = tb.it[blockIdx.x];
  const float* s = ws + tb.s_off[blockIdx.x];
  const int rows = it.rows;
  __shared__ float red[32];
  __shared__ float bc;
  const int tid = threadIdx.x;
  float part = 0.f;
  if (training) {
    for (int i = tid; i < rows; i += blockDim.x) part = fmaf(s[i], s[i], part);
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
      bc = 1.0f / fmaxf(sqrtf(tot), eps);
    }
    __syncthreads();
    const float inv = bc;
    for (int i = tid; i < rows; i += blockDim.x) it.u[i] = s[i] * inv;
    __syncthreads();
  }
  part = 0.f;
  for (int i = tid; i < rows; i += blockDim.x) {
    const float uu = it.u[i];
    if (it.u_out != nullptr) it.u_out[i] = uu;
    part = fmaf(uu, s[i], part);
  }
  part = warp_sum(part);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    *it.sigma = tot;
  }
}
__global__ void sn_bwd_dot_kernel(const float* __restrict__ dwh, const float* __restrict__ w, long long n, float* __restrict__ acc,
                                  float* __restrict__ partials) {
  vg::pdl_entry();
  float part = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    part = fmaf(dwh[i], w[i], part);
  part = warp_sum(part);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += red[k];
    if (partials) partials[blockIdx.x] = tot;
    else atomicAdd(acc, tot);
  }
}
__global__ void sn_bwd_apply_kernel(const float* __restrict__ dwh, const float* __restrict__ u, const float* __restrict__ v,
                                    const float* __restrict__ sigma, const float* __restrict__ dot, int rows, int cols,
                                    float* __restrict__ dw) {
  vg::pdl_entry();
  const long long n = (long long)rows * cols;
  const float inv = 1.0f / *sigma;
  const float c = *dot * inv;   // <dw_hat, W> / sigma
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int r = (int)(i / cols), cc = (int)(i % cols);
    dw[i] += (dwh[i] - c * u[r] * v[cc]) * inv;
  }
}

// The same two steps reading the weight gradient in the tensor-core kernel's PACKED layout acc[tap][c_s][c_u] (what
// tc_wgrad_kernel reduces into), so that no unpack pass is needed: dw_orig[c_u][c_s][tap] is torch's layout,
// rows = c_u (u), cols = c_s * taps + tap (v).
__global__ void sn_bwd_dot_packed_kernel(const float* __restrict__ acc, const float* __restrict__ w, int cu, int cs, int taps,
                                         float* __restrict__ out, float* __restrict__ partials) {
  vg::pdl_entry();
  const long long n = (long long)taps * cs * cu;
  float part = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c_u = (int)(i % cu);
    const long long t = i / cu;
    const int c_s = (int)(t % cs), tap = (int)(t / cs);
    part = fmaf(acc[i], w[((long long)c_u * cs + c_s) * taps + tap], part);      // w: a few MB, L2 resident
  }
  part = warp_sum(part);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += red[k];
    if (partials) partials[blockIdx.x] = tot;
    else atomicAdd(out, tot);
  }
}
// 32 x 32 shared-memory transpose between c_u and r = c_s * taps + tap (as wgrad_unpack_kernel), with the spectral-norm
// correction applied on the way: dw[c_u][r] += (acc - c * u[c_u] * v[r]) / sigma
__global__ void __launch_bounds__(256) sn_bwd_apply_packed_kernel(const float* __restrict__ acc, const float* __restrict__ u,
                                                                  const float* __restrict__ v, const float* __restrict__ sigma,
                                                                  const float* __restrict__ dot, int cu_n, int cs_n, int taps,
                                                                  float* __restrict__ dw) {
  vg::pdl_entry();
  __shared__ float tile[32][33];
  const int R = cs_n * taps;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv = 1.0f / *sigma;
  const float c = *dot * inv;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, cu = c0 + tx;
    float val = 0.f;
    if (r < R && cu < cu_n) {
      const int cs = r / taps, tap = r - cs * taps;
      val = acc[((long long)tap * cs_n + cs) * cu_n + cu];
    }
    tile[i][tx] = val;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int cu = c0 + i, r = r0 + tx;
    if (cu < cu_n && r < R) dw[(long long)cu * R + r] += (tile[tx][i] - c * u[cu] * v[r]) * inv;
  }
}

// ------------------------------------------------------------------------------------------
// reparameterisation
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv_raw, const float* __restrict__ eps,
                                   long long n, int training, T* __restrict__ z, float* __restrict__ lv) {
  vg::pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float l = fminf(fmaxf(lv_raw[i], -50.f), 50.f);
    lv[i] = l;
    float zz = mu[i];
    if (training) zz = fmaf(expf(0.5f * l), eps[i], zz);
    z[i] = from_f32<T>(zz);
  }
}
template <typename T>
__global__ void reparam_bwd_kernel(const T* __restrict__ dz, const float* __restrict__ lv_raw, const float* __restrict__ eps,
                                   const float* __restrict__ dlv_in, long long n, int training, float* __restrict__ d_mu,
                                   float* __restrict__ d_lv) {
  vg::pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = to_f32(dz[i]);
    d_mu[i] = g;
    float raw = lv_raw[i];
    float r = 0.f;
    if (raw >= -50.f && raw <= 50.f) {
      if (training) r = g * 0.5f * expf(0.5f * raw) * eps[i];
      if (dlv_in != nullptr) r += dlv_in[i];
    }
    d_lv[i] = r;
  }
}

// ------------------------------------------------------------------------------------------
// fused generator loss + gradients
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_atomic_add(double v, double* dst, double* partial = nullptr) {
  __shared__ double red[32];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    if (partial) *partial = t;          // deterministic mode: per-block partial, added in block order afterwards
    else atomicAdd(dst, t);
  }
}
__device__ __forceinline__ float softplus_f(float x) { return x > 0.f ? x + log1pf(expf(-x)) : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename T>
__global__ void __launch_bounds__(256) generator_loss_kernel(const T* __restrict__ xhat, const float* __restrict__ x,
                                                              const float* __restrict__ mu, const float* __restrict__ lv,
                                                              const float* __restrict__ logits, VgLossDesc d, T* __restrict__ d_xhat,
                                                              float* __restrict__ d_mu, float* __restrict__ d_lv,
                                                              float* __restrict__ d_logits, double* __restrict__ losses,
                                                              double* __restrict__ partials) {
  vg::pdl_entry();
  const long long gsz = (long long)gridDim.x * blockDim.x;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // reconstruction: mean|d| + mean d^2 over the GLOBAL pixel count
  double recon = 0.0;
  const float inv_pix = 1.0f / (float)d.n_pix_global;
  for (long long i = gid; i < d.n_pix; i += gsz) {
    float df = to_f32(xhat[i]) - x[i];
    recon += (double)(fabsf(df) + df * df);
    float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
    d_xhat[i] = from_f32<T>(d.w_recon * inv_pix * (sg + 2.f * df));
  }
  // KL = -0.5 * sum(1 + lv - mu^2 - exp(lv))
  double kl = 0.0;
  for (long long i = gid; i < d.n_lat; i += gsz) {
    float m = mu[i], l = lv[i], e = expf(l);
    kl += (double)(-0.5f * (1.f + l - m * m - e));
    d_mu[i] = d.w_kl * m;
    d_lv[i] = d.w_kl * 0.5f * (e - 1.f);
  }
  // adversarial term on D(xhat) logits
  double adv = 0.0;
  const float inv_b = 1.0f / (float)d.n_logits_global;
  for (long long i = gid; i < d.n_logits; i += gsz) {
    float z = logits[i];
    if (d.adv_mode == 0) {          // BCE-with-logits, target 1: softplus(-z)
      adv += (double)softplus_f(-z);
      d_logits[i] = d.w_adv * inv_b * (sigmoid_f(z) - 1.f);
    } else {                         // -mean(D)
      adv += (double)(-z);
      d_logits[i] = -d.w_adv * inv_b;
    }
  }
  recon *= (double)inv_pix;
  adv *= (double)inv_b;
  double* pp = partials ? partials + (size_t)blockIdx.x * 4 : nullptr;
  block_atomic_add(recon, &losses[1], pp ? pp + 1 : nullptr);
  block_atomic_add(kl, &losses[2], pp ? pp + 2 : nullptr);
  block_atomic_add(adv, &losses[3], pp ? pp + 3 : nullptr);
  block_atomic_add((double)d.w_recon * recon + (double)d.w_kl * kl + (double)d.w_adv * adv, &losses[0], pp);
}

__global__ void discriminator_loss_kernel(const float* __restrict__ d_real, const float* __restrict__ d_fake, int n, int n_global,
                                          int adv_mode, float* __restrict__ g_real, float* __restrict__ g_fake,
                                          double* __restrict__ losses) {
  vg::pdl_entry();
  double lr = 0.0, lf = 0.0;
  const float inv = 1.0f / (float)n_global;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float r = d_real[i], f = d_fake[i];
    if (adv_mode == 0) {
      lr += (double)softplus_f(-r);            // BCE(real, 1)
      lf += (double)softplus_f(f);             // BCE(fake, 0)
      g_real[i] = inv * (sigmoid_f(r) - 1.f);
      g_fake[i] = inv * sigmoid_f(f);
    } else {
      lr += (double)(-r);                      // -mean(D(real))   README.md:792
      lf += (double)f;                         //  mean(D(fake))   README.md:793
      g_real[i] = -inv;
      g_fake[i] = inv;
    }
  }
  lr *= (double)inv;
  lf *= (double)inv;
  block_atomic_add(lr, &losses[1]);
  block_atomic_add(lf, &losses[2]);
  block_atomic_add(lr + lf, &losses[0]);
}

// ------------------------------------------------------------------------------------------
// fused optimizer
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) optimizer_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, long long n4, long long n, VgOptDesc d,
                                                         const unsigned long long* __restrict__ step_ptr) {
  vg::pdl_entry();
  if (step_ptr != nullptr && d.kind == 0) {
    const float t = (float)(*step_ptr);
    d.bias_corr1 = 1.f - powf(d.beta1, t);
    d.bias_corr2 = 1.f - powf(d.beta2, t);
  }
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= d.grad_scale;
    if (d.weight_decay != 0.f) gg = fmaf(d.weight_decay, pp, gg);
    if (d.kind == 0) {
      mm = d.beta1 * mm + (1.f - d.beta1) * gg;
      vv = d.beta2 * vv + (1.f - d.beta2) * gg * gg;
      float denom = sqrtf(vv) / sqrtf(d.bias_corr2) + d.eps;
      pp -= (d.lr / d.bias_corr1) * (mm / denom);
    } else {
      vv = d.alpha * vv + (1.f - d.alpha) * gg * gg;
      pp -= d.lr * gg / (sqrtf(vv) + d.eps);
    }
    if (d.clamp > 0.f) pp = fminf(fmaxf(pp, -d.clamp), d.clamp);
  };
  const long long gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gsz) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 mm = d.kind == 0 ? reinterpret_cast<float4*>(m)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (d.kind == 0) reinterpret_cast<float4*>(m)[i] = mm;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gsz) {
    float mm = d.kind == 0 ? m[i] : 0.f;
    upd(p[i], g[i], mm, v[i]);
    if (d.kind == 0) m[i] = mm;
  }
}

// ------------------------------------------------------------------------------------------
// one-shot all-reduce of a small fp64 vector over NVLink peer memory (SyncBN statistics)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(double* __restrict__ vec, int n, VgPeerDesc pd, int slot,
                                                                 const unsigned long long* __restrict__ epoch_ptr) {
  vg::pdl_entry();
  const unsigned long long epoch = *epoch_ptr;
  const int tid = threadIdx.x;
  const size_t slot_off = ((size_t)slot * pd.world + pd.rank) * VG_PEER_MAX_N;
  // 1. my contribution into my sub-slot of every rank's buffer (NVLink stores for r != rank)
  for (int r = 0; r < pd.world; ++r) {
    double* dst = reinterpret_cast<double*>(pd.peer_data[r]) + slot_off;
    for (int i = tid; i < n; i += blockDim.x) dst[i] = vec[i];
  }
  __threadfence_system();          // every writer: its stores are visible system-wide ...
  __syncthreads();                 // ... before any thread of the block publishes
  // 2. publish: flag[slot][rank] = epoch on every rank.  The publishing threads did not write all the data
  //    themselves: the fence after the barrier orders the block's (already system-visible) stores before the flag.
  if (tid < pd.world) {
    __threadfence_system();
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(pd.peer_flags[tid]) + (size_t)slot * pd.world + pd.rank;
    *f = epoch;
  }
  // 3. wait until every rank has published THIS epoch into MY flags (equality: a stale or a future value never passes)
  if (tid < pd.world) {
    volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(pd.peer_flags[pd.rank]) + (size_t)slot * pd.world + tid;
    const long long t0 = clock64();
    while (*f != epoch) {
      if (clock64() - t0 > 20000000000LL) {     // ~10 s: a peer never arrived - fail loudly instead of hanging
        printf("vaegan_b200: peer all-reduce timed out (rank %d waiting for rank %d, slot %d, epoch %llu, flag %llu)\n", pd.rank, tid, slot,
               epoch, (unsigned long long)*f);
        __trap();
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  // 4. sum the sub-slots in rank order (identical result on every rank); bypass L1 for peer-written data
  const double* mine = reinterpret_cast<const double*>(pd.peer_data[pd.rank]) + (size_t)slot * pd.world * VG_PEER_MAX_N;
  for (int i = tid; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < pd.world; ++r) s += __ldcv(mine + (size_t)r * VG_PEER_MAX_N + i);
    vec[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// layout / dtype
// ------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long n) {
  vg::pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<TD>(to_f32(src[i]));
}
__global__ void fold_bn_into_conv_kernel(const float* __restrict__ w, const float* __restrict__ bias_in, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                         int c_out, int c_in, int inner, int transposed, float* __restrict__ w_out,
                                         float* __restrict__ bias_out) {
  vg::pdl_entry();
  const long long total = (long long)c_out * c_in * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pair = i / inner;
    const int co = transposed ? (int)(pair % c_out) : (int)(pair / c_in);
    const float sc = (float)((double)gamma[co] / sqrt((double)rv[co] + (double)eps));
    w_out[i] = w[i] * sc;
  }
  for (int co = blockIdx.x * blockDim.x + threadIdx.x; co < c_out; co += gridDim.x * blockDim.x) {
    const float sc = (float)((double)gamma[co] / sqrt((double)rv[co] + (double)eps));
    bias_out[co] = beta[co] - rm[co] * sc + (bias_in ? bias_in[co] * sc : 0.f);
  }
}
__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                                      const float* __restrict__ rv, float eps, int c, float* __restrict__ scale, float* __restrict__ shift) {
  vg::pdl_entry();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float sc = (float)((double)gamma[ch] / sqrt((double)rv[ch] + (double)eps));
  scale[ch] = sc;
  shift[ch] = beta[ch] - rm[ch] * sc;
}

template <typename T>
__global__ void scale_kernel(const T* __restrict__ src, const float* __restrict__ scale, long long n, T* __restrict__ dst) {
  vg::pdl_entry();
  const float sc = *scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(to_f32(src[i]) * sc);
}
// ------------------------------------------------------------------------------------------
// input pipeline: per-image min-max normalisation to [0, 1] - the dataset's
// `img = (img - img.min()) / (img.max() - img.min())` (README.md:87, float64 arithmetic) - fused with the
// float64 -> float32 cast of README.md:785 (`imgs.type(Tensor)`) and, optionally, the bf16 copy the convolutions
// read.  One block per image: pass 1 reduces min / max, pass 2 re-reads the image (L2-resident) and writes.
// ------------------------------------------------------------------------------------------
template <typename TR> __device__ __forceinline__ double raw_to_f64(TR v) { return (double)v; }
template <typename TR>
__global__ void __launch_bounds__(256) normalize_images_kernel(const TR* __restrict__ raw, long long pixels, float* __restrict__ out_f32,
                                                               __nv_bfloat16* __restrict__ out_bf16) {
  vg::pdl_entry();
  const TR* img = raw + (long long)blockIdx.x * pixels;
  double lo = 1e300, hi = -1e300;
  bool has_nan = false;
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const double v = raw_to_f64(img[i]);
    if (v != v) has_nan = true;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  __shared__ double s_lo[8], s_hi[8];
  __shared__ int s_nan;
  if (threadIdx.x == 0) s_nan = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  if (has_nan) s_nan = 1;
  __syncthreads();
  lo = s_lo[0]; hi = s_hi[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
  // numpy semantics: a NaN pixel makes min and max NaN, hence the whole image; a constant image divides 0 by 0
  const double range = s_nan ? __longlong_as_double(0x7ff8000000000000LL) : (hi - lo);
  for (long long i = threadIdx.x; i < pixels; i += blockDim.x) {
    const float r = (float)((raw_to_f64(img[i]) - lo) / range);
    const long long o = (long long)blockIdx.x * pixels + i;
    if (out_f32 != nullptr) out_f32[o] = r;
    if (out_bf16 != nullptr) out_bf16[o] = __float2bfloat16_rn(r);
  }
}

// tiled transpose of [c][hw] <-> [hw][c] per image
template <typename TD>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int c, int hw, TD* __restrict__ dst) {
  vg::pdl_entry();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* s = src + (long long)n * c * hw;
  TD* d = dst + (long long)n * c * hw;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int cc = c0 + r, pp = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (cc < c && pp < hw) ? s[(long long)cc * hw + pp] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int pp = p0 + r, cc = c0 + threadIdx.x;
    if (pp < hw && cc < c) d[(long long)pp * c + cc] = from_f32<TD>(tile[threadIdx.x][r]);
  }
}
template <typename TS>
__global__ void nhwc_to_nchw_kernel(const TS* __restrict__ src, int c, int hw, float* __restrict__ dst) {
  vg::pdl_entry();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const TS* s = src + (long long)n * c * hw;
  float* d = dst + (long long)n * c * hw;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int pp = p0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (pp < hw && cc < c) ? to_f32(s[(long long)pp * c + cc]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int cc = c0 + r, pp = p0 + threadIdx.x;
    if (cc < c && pp < hw) d[(long long)cc * hw + pp] = tile[threadIdx.x][r];
  }
}

}  // namespace vg

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace vg;

static inline int ew_grid(long long n, int per_thread = 4) {
  return (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256LL * per_thread), (long long)num_sms() * 8));
}

extern "C" int vg_linear_forward(const void* x, const void* w, const float* bias, int m, int n, int k, int dtype, float slope,
                                 void* y, vg_stream_t stream) {
  VG_CHECK_ARG(x && w && y && m >= 0 && n > 0 && k > 0, "bad args");
  if (m == 0) return VG_OK;
  cudaStream_t s = as_stream(stream);
  // x, y are fp32 (the head's activations are tiny); w is `dtype`.  y = x w^T: I=m, J=n, L=k
  int tiles = (int)(cdiv(n, 32) * cdiv(m, 32));
  int splits = (int)std::max<long long>(1, std::min<long long>(cdiv(2LL * num_sms(), tiles), cdiv(k, 256)));
  int mode = 0;
  if (splits > 1) {
    VG_CUDA(cudaMemsetAsync(y, 0, (size_t)m * n * sizeof(float), s));
    mode = 2;
  }
  int rc;
  if (dtype == VG_BF16)
    rc = launch_gemm<float, __nv_bfloat16>((const float*)x, k, 1, (const __nv_bfloat16*)w, 1, k, m, n, k, (float*)y, mode, splits, s);
  else
    rc = launch_gemm<float, float>((const float*)x, k, 1, (const float*)w, 1, k, m, n, k, (float*)y, mode, splits, s);
  if (rc) return rc;
  if (bias != nullptr || slope != 1.0f) {
    long long total = (long long)m * n;
    vg::Launch(ew_grid(total, 1), 256, 0, s)(bias_lrelu_kernel, (float*)y, bias, total, n, slope);
    VG_LAUNCHED();
  }
  return VG_OK;
}

extern "C" int vg_lrelu_forward(const void* x, long long n, int dtype, float slope, void* y, vg_stream_t stream) {
  VG_CHECK_ARG(x && y && n >= 0, "bad args");
  VG_CHECK_ARG(dtype == VG_F32, "vg_lrelu_forward: fp32 only (head activations)");
  if (n == 0) return VG_OK;
  if (x != y) VG_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, as_stream(stream)));
  vg::Launch(ew_grid(n, 1), 256, 0, as_stream(stream))(bias_lrelu_kernel, (float*)y, nullptr, n, 1, slope);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_linear_dgrad(const void* dy, const void* w, int m, int n, int k, int dtype, void* dx, vg_stream_t stream) {
  VG_CHECK_ARG(dy && w && dx && m >= 0 && n > 0 && k > 0, "bad args");
  if (m == 0) return VG_OK;
  cudaStream_t s = as_stream(stream);
  // dx[m][k] = dy[m][n] w[n][k]: I=m, J=k, L=n
  if (dtype == VG_BF16)
    return launch_gemm<float, __nv_bfloat16>((const float*)dy, n, 1, (const __nv_bfloat16*)w, k, 1, m, k, n, (float*)dx, 0, 1, s);
  return launch_gemm<float, float>((const float*)dy, n, 1, (const float*)w, k, 1, m, k, n, (float*)dx, 0, 1, s);
}

extern "C" int vg_linear_wgrad(const void* x, const void* dy, int m, int n, int k, int dtype, float* dw, float* dbias,
                               vg_stream_t stream) {
  (void)dtype;
  VG_CHECK_ARG(x && dy && dw && m >= 0 && n > 0 && k > 0, "bad args");
  if (m == 0) return VG_OK;
  cudaStream_t s = as_stream(stream);
  // dw[n][k] += dy^T x : I=n, J=k, L=m ; A(i,l) = dy[l*n + i], B(l,j) = x[l*k + j]
  int rc = launch_gemm<float, float>((const float*)dy, 1, n, (const float*)x, k, 1, n, k, m, dw, 1, 1, s);
  if (rc) return rc;
  if (dbias != nullptr) {
    vg::Launch((n + 127) / 128, 128, 0, s)(colsum_f32_kernel, (const float*)dy, m, n, dbias);
    VG_LAUNCHED();
  }
  return VG_OK;
}

extern "C" int vg_avgpool_flatten_forward(const void* x, int n, int h, int w, int c, int k, int dtype, void* out, vg_stream_t stream) {
  VG_CHECK_ARG(x && out && n >= 0 && h > 0 && w > 0 && c > 0 && k > 0, "bad args");
  long long total = (long long)n * (h / k) * (w / k) * c;
  if (total == 0) return VG_OK;
  if (dtype == VG_BF16)
    vg::Launch(ew_grid(total, 1), 256, 0, as_stream(stream))(avgpool_flatten_fwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)x, n, h, w, c, k, (float*)out);
  else
    vg::Launch(ew_grid(total, 1), 256, 0, as_stream(stream))(avgpool_flatten_fwd_kernel<float>, (const float*)x, n, h, w, c, k, (float*)out);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_avgpool_flatten_backward(const void* dout, int n, int h, int w, int c, int k, int dtype, void* dx, vg_stream_t stream) {
  VG_CHECK_ARG(dout && dx && n >= 0 && h > 0 && w > 0 && c > 0 && k > 0, "bad args");
  long long total = (long long)n * h * w * c;
  if (total == 0) return VG_OK;
  const size_t tile_bytes = (size_t)(w / k) * c * sizeof(float);
  if (h % k == 0 && w % k == 0 && c % 8 == 0 && tile_bytes <= 48 * 1024 && (long long)n * (h / k) < (1LL << 31)) {
    const unsigned blocks = (unsigned)(n * (h / k));
    if (dtype == VG_BF16)
      vg::Launch(blocks, 256, tile_bytes, as_stream(stream))(avgpool_flatten_bwd_rows_kernel<__nv_bfloat16>, (const float*)dout, h, w, c, k, (__nv_bfloat16*)dx);
    else
      vg::Launch(blocks, 256, tile_bytes, as_stream(stream))(avgpool_flatten_bwd_rows_kernel<float>, (const float*)dout, h, w, c, k, (float*)dx);
    VG_LAUNCHED();
    return VG_OK;
  }
  if (dtype == VG_BF16)
    vg::Launch(ew_grid(total), 256, 0, as_stream(stream))(avgpool_flatten_bwd_kernel<__nv_bfloat16>, (const float*)dout, n, h, w, c, k, (__nv_bfloat16*)dx);
  else
    vg::Launch(ew_grid(total), 256, 0, as_stream(stream))(avgpool_flatten_bwd_kernel<float>, (const float*)dout, n, h, w, c, k, (float*)dx);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_spectral_norm_sigma(const float* w_orig, int rows, int cols, float* u, float* v, int training, float eps,
                                      float* sigma, float* workspace, vg_stream_t stream) {
  VG_CHECK_ARG(w_orig && u && v && sigma && workspace && rows > 0 && cols > 0, "bad args");
  cudaStream_t s = as_stream(stream);
  float* t = workspace;            // [cols]
  float* sv = workspace + cols;    // [rows]
  if (training) {
    VG_CUDA(cudaMemsetAsync(t, 0, (size_t)cols * sizeof(float), s));
    int slices = g_det.on ? 1 : std::max(1, std::min(rows / 8, 64));   // deterministic mode: one slice, no cross-block sum
    int rps = (int)cdiv(rows, slices);
    dim3 g1((unsigned)cdiv(cols, 128), (unsigned)cdiv(rows, rps));
    vg::Launch(g1, 128, 0, s)(sn_wt_u_kernel, w_orig, u, rows, cols, rps, t);
    VG_LAUNCHED();
    vg::Launch((unsigned)rows, 128, 0, s)(sn_w_v_kernel, w_orig, t, rows, cols, 1, eps, v, sv);
    VG_LAUNCHED();
  } else {
    vg::Launch((unsigned)rows, 128, 0, s)(sn_w_v_kernel, w_orig, v, rows, cols, 0, eps, nullptr, sv);
    VG_LAUNCHED();
  }
  vg::Launch(1, 256, 0, s)(sn_finish_kernel, sv, u, rows, training, eps, sigma);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_spectral_norm_sigma_batched(const VgSnItem* items, int n_items, int training, float eps, float* workspace,
                                              size_t workspace_floats, vg_stream_t stream) {
  VG_CHECK_ARG(items && n_items >= 0 && workspace, "bad args");
  cudaStream_t s = as_stream(stream);
  for (int base = 0; base < n_items; base += VG_SN_MAX) {
    SnTable tb;
    memset(&tb, 0, sizeof(tb));
    tb.n = std::min(VG_SN_MAX, n_items - base);
    size_t off = 0;
    int max_rows = 0, max_cols = 0;
    for (int i = 0; i < tb.n; ++i) {
      const VgSnItem& it = items[base + i];
      VG_CHECK_ARG(it.w && it.u && it.v && it.sigma && it.rows > 0 && it.cols > 0, "bad spectral-norm item %d", base + i);
      tb.it[i] = it;
      tb.t_off[i] = (int)off; off += (size_t)it.cols;
      tb.s_off[i] = (int)off; off += (size_t)it.rows;
      max_rows = std::max(max_rows, it.rows);
      max_cols = std::max(max_cols, it.cols);
    }
    VG_CHECK_ARG(off <= workspace_floats, "workspace too small: need %zu floats, have %zu", off, workspace_floats);
    if (training) {
      VG_CUDA(cudaMemsetAsync(workspace, 0, off * sizeof(float), s));
      const int max_slices = g_det.on ? 1 : std::max(1, std::min(max_rows / 8, 64));
      dim3 g1((unsigned)cdiv(max_cols, 128), (unsigned)max_slices, (unsigned)tb.n);
      vg::Launch(g1, 128, 0, s)(sn_wt_u_batched_kernel, tb, workspace, g_det.on);
      VG_LAUNCHED();
    }
    dim3 g2((unsigned)max_rows, (unsigned)tb.n);
    vg::Launch(g2, 128, 0, s)(sn_w_v_batched_kernel, tb, workspace, training, eps);
    VG_LAUNCHED();
    vg::Launch(tb.n, 256, 0, s)(sn_finish_batched_kernel, tb, workspace, training, eps);
    VG_LAUNCHED();
  }
  return VG_OK;
}

namespace vg {
int sn_backward_run(const float* dw_hat, const float* w_orig, const float* u, const float* v, const float* sigma, int rows, int cols,
                    float* dw_orig, float* workspace, cudaStream_t s);
}
extern "C" int vg_spectral_norm_backward(const float* dw_hat, const float* w_orig, const float* u, const float* v,
                                         const float* sigma, int rows, int cols, float* dw_orig, float* workspace,
                                         vg_stream_t stream) {
  VG_CHECK_ARG(dw_hat && w_orig && u && v && sigma && dw_orig && workspace && rows > 0 && cols > 0, "bad args");
  cudaStream_t s = as_stream(stream);
  VG_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float), s));
  return vg::sn_backward_run(dw_hat, w_orig, u, v, sigma, rows, cols, dw_orig, workspace, s);
}

namespace vg {
// dot_ws: one float, ZEROED by the caller
int sn_backward_packed_run(const float* acc, const float* w_orig, const float* u, const float* v, const float* sigma, int cu, int cs,
                           int taps, float* dw_orig, float* dot_ws, cudaStream_t s) {
  const long long n = (long long)taps * cs * cu;
  float* part = nullptr;
  const int dot_grid = ew_grid(n);
  if (g_det.on && !(part = (float*)det_scratch((size_t)dot_grid * sizeof(float)))) return VG_EINVAL;
  vg::Launch(dot_grid, 256, 0, s)(sn_bwd_dot_packed_kernel, acc, w_orig, cu, cs, taps, dot_ws, part);
  VG_LAUNCHED();
  if (part) {
    int rc = ordered_reduce_f32(part, dot_grid, 1, dot_ws, s);
    if (rc) return rc;
  }
  dim3 grid((unsigned)cdiv((long long)cs * taps, 32), (unsigned)cdiv(cu, 32));
  vg::Launch(grid, 256, 0, s)(sn_bwd_apply_packed_kernel, acc, u, v, sigma, dot_ws, cu, cs, taps, dw_orig);
  VG_LAUNCHED();
  return VG_OK;
}

int sn_backward_run(const float* dw_hat, const float* w_orig, const float* u, const float* v, const float* sigma, int rows, int cols,
                    float* dw_orig, float* workspace, cudaStream_t s) {
  long long n = (long long)rows * cols;
  float* part = nullptr;
  const int dot_grid = ew_grid(n);
  if (g_det.on && !(part = (float*)det_scratch((size_t)dot_grid * sizeof(float)))) return VG_EINVAL;
  vg::Launch(dot_grid, 256, 0, s)(sn_bwd_dot_kernel, dw_hat, w_orig, n, workspace, part);
  VG_LAUNCHED();
  if (part) {
    int rc = ordered_reduce_f32(part, dot_grid, 1, workspace, s);
    if (rc) return rc;
  }
  vg::Launch(ew_grid(n), 256, 0, s)(sn_bwd_apply_kernel, dw_hat, u, v, sigma, workspace, rows, cols, dw_orig);
  VG_LAUNCHED();
  return VG_OK;
}
}  // namespace vg

extern "C" int vg_reparam_forward(const float* mu, const float* lv_raw, const float* eps, long long n, int training, int z_dtype,
                                  void* z, float* lv_clamped, vg_stream_t stream) {
  VG_CHECK_ARG(mu && lv_raw && z && lv_clamped && n >= 0 && (!training || eps), "bad args");
  if (n == 0) return VG_OK;
  if (z_dtype == VG_BF16)
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(reparam_fwd_kernel<__nv_bfloat16>, mu, lv_raw, eps, n, training, (__nv_bfloat16*)z, lv_clamped);
  else
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(reparam_fwd_kernel<float>, mu, lv_raw, eps, n, training, (float*)z, lv_clamped);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_reparam_backward(const void* dz, const float* lv_raw, const float* eps, const float* dlv_in, long long n,
                                   int training, int z_dtype, float* d_mu, float* d_lv_raw, vg_stream_t stream) {
  VG_CHECK_ARG(dz && lv_raw && d_mu && d_lv_raw && n >= 0 && (!training || eps), "bad args");
  if (n == 0) return VG_OK;
  if (z_dtype == VG_BF16)
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(reparam_bwd_kernel<__nv_bfloat16>, (const __nv_bfloat16*)dz, lv_raw, eps, dlv_in, n, training, d_mu, d_lv_raw);
  else
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(reparam_bwd_kernel<float>, (const float*)dz, lv_raw, eps, dlv_in, n, training, d_mu, d_lv_raw);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_generator_loss(const void* xhat, const float* x, const float* mu, const float* lv, const float* logits,
                                 const VgLossDesc* d, void* d_xhat, float* d_mu, float* d_lv, float* d_logits, double* losses,
                                 vg_stream_t stream) {
  VG_CHECK_ARG(d && xhat && x && mu && lv && d_xhat && d_mu && d_lv && losses, "null pointer");
  VG_CHECK_ARG(d->n_logits == 0 || (logits && d_logits), "logits missing");
  VG_CHECK_ARG(d->n_pix_global > 0 && (d->n_logits == 0 || d->n_logits_global > 0), "global counts must be positive");
  long long work = std::max(d->n_pix, d->n_lat);
  int grid = ew_grid(work, 8);
  VgLossDesc dd = *d;
  if (dd.n_logits_global <= 0) dd.n_logits_global = 1;
  double* part = nullptr;
  if (g_det.on && !(part = (double*)det_scratch((size_t)grid * 4 * sizeof(double)))) return VG_EINVAL;
  if (d->xhat_dtype == VG_BF16)
    vg::Launch(grid, 256, 0, as_stream(stream))(generator_loss_kernel<__nv_bfloat16>, (const __nv_bfloat16*)xhat, x, mu, lv, logits, dd,
                                                (__nv_bfloat16*)d_xhat, d_mu, d_lv, d_logits, losses, part);
  else
    vg::Launch(grid, 256, 0, as_stream(stream))(generator_loss_kernel<float>, (const float*)xhat, x, mu, lv, logits, dd, (float*)d_xhat, d_mu, d_lv,
                                                d_logits, losses, part);
  VG_LAUNCHED();
  if (part) return ordered_reduce_f64(part, grid, 4, losses, as_stream(stream));
  return VG_OK;
}

extern "C" int vg_discriminator_loss(const float* d_real, const float* d_fake, int n, int n_global, int adv_mode, float* g_real,
                                     float* g_fake, double* losses, vg_stream_t stream) {
  VG_CHECK_ARG(d_real && d_fake && g_real && g_fake && losses && n >= 0 && n_global > 0, "bad args");
  vg::Launch(1, 256, 0, as_stream(stream))(discriminator_loss_kernel, d_real, d_fake, n, n_global, adv_mode, g_real, g_fake, losses);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_optimizer_step(float* p, const float* g, float* m, float* v, long long n, const VgOptDesc* d,
                                 const unsigned long long* step_ptr, vg_stream_t stream) {
  VG_CHECK_ARG(p && g && v && d && n >= 0, "null pointer");
  VG_CHECK_ARG(d->kind == 0 || d->kind == 1, "unknown optimizer kind %d", d->kind);
  VG_CHECK_ARG(d->kind != 0 || (m != nullptr && (step_ptr != nullptr || (d->bias_corr1 > 0.f && d->bias_corr2 > 0.f))), "Adam needs m and bias corrections");
  if (n == 0) return VG_OK;
  bool al = ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)v % 16 == 0) && (!m || (uintptr_t)m % 16 == 0);
  long long n4 = al ? n / 4 : 0;
  vg::Launch(ew_grid(n, 8), 256, 0, as_stream(stream))(optimizer_kernel, p, g, m, v, n4, n, *d, step_ptr);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_peer_allreduce_f64(double* vec, int n, const VgPeerDesc* pd, int slot, const unsigned long long* epoch_ptr,
                                     vg_stream_t stream) {
  VG_CHECK_ARG(vec && pd && epoch_ptr, "null pointer");
  VG_CHECK_ARG(n > 0 && n <= VG_PEER_MAX_N, "n must be in 1..%d", VG_PEER_MAX_N);
  VG_CHECK_ARG(pd->world >= 1 && pd->world <= VG_PEER_MAX_WORLD && pd->rank >= 0 && pd->rank < pd->world, "bad rank/world");
  VG_CHECK_ARG(slot >= 0 && slot < pd->n_slots, "slot %d out of range (n_slots %d)", slot, pd->n_slots);
  for (int r = 0; r < pd->world; ++r) VG_CHECK_ARG(pd->peer_data[r] && pd->peer_flags[r], "peer pointer %d is null", r);
  vg::Launch(1, 256, 0, as_stream(stream))(peer_allreduce_f64_kernel, vec, n, *pd, slot, epoch_ptr);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, vg_stream_t stream) {
  VG_CHECK_ARG(src && dst && n >= 0, "bad args");
  if (n == 0) return VG_OK;
  cudaStream_t s = as_stream(stream);
  int grid = ew_grid(n);
  if (src_dtype == VG_F32 && dst_dtype == VG_BF16) vg::Launch(grid, 256, 0, s)(cast_kernel<float, __nv_bfloat16>, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == VG_BF16 && dst_dtype == VG_F32) vg::Launch(grid, 256, 0, s)(cast_kernel<__nv_bfloat16, float>, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == VG_F32 && dst_dtype == VG_F32) vg::Launch(grid, 256, 0, s)(cast_kernel<float, float>, (const float*)src, (float*)dst, n);
  else if (src_dtype == VG_BF16 && dst_dtype == VG_BF16) vg::Launch(grid, 256, 0, s)(cast_kernel<__nv_bfloat16, __nv_bfloat16>, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else { set_error("bad dtypes %d -> %d", src_dtype, dst_dtype); return VG_EINVAL; }
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_nchw_to_nhwc(const float* src, int n, int c, int h, int w, int dst_dtype, void* dst, vg_stream_t stream) {
  VG_CHECK_ARG(src && dst && n >= 0 && c > 0 && h > 0 && w > 0, "bad args");
  if (n == 0) return VG_OK;
  VG_CHECK_ARG(n <= 65535, "n too large for this helper");
  dim3 grid((unsigned)cdiv(h * w, 32), (unsigned)cdiv(c, 32), (unsigned)n), block(32, 8);
  if (dst_dtype == VG_BF16) vg::Launch(grid, block, 0, as_stream(stream))(nchw_to_nhwc_kernel<__nv_bfloat16>, src, c, h * w, (__nv_bfloat16*)dst);
  else vg::Launch(grid, block, 0, as_stream(stream))(nchw_to_nhwc_kernel<float>, src, c, h * w, (float*)dst);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_nhwc_to_nchw(const void* src, int src_dtype, int n, int c, int h, int w, float* dst, vg_stream_t stream) {
  VG_CHECK_ARG(src && dst && n >= 0 && c > 0 && h > 0 && w > 0, "bad args");
  if (n == 0) return VG_OK;
  VG_CHECK_ARG(n <= 65535, "n too large for this helper");
  dim3 grid((unsigned)cdiv(h * w, 32), (unsigned)cdiv(c, 32), (unsigned)n), block(32, 8);
  if (src_dtype == VG_BF16) vg::Launch(grid, block, 0, as_stream(stream))(nhwc_to_nchw_kernel<__nv_bfloat16>, (const __nv_bfloat16*)src, c, h * w, dst);
  else vg::Launch(grid, block, 0, as_stream(stream))(nhwc_to_nchw_kernel<float>, (const float*)src, c, h * w, dst);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_fold_bn_into_conv(const float* w, const float* bias_in, const float* gamma, const float* beta, const float* running_mean,
                                    const float* running_var, float eps, int c_out, int c_in, int inner, int transposed, float* w_out,
                                    float* bias_out, vg_stream_t stream) {
  VG_CHECK_ARG(w && gamma && beta && running_mean && running_var && w_out && bias_out && c_out > 0 && c_in > 0 && inner > 0, "bad args");
  const long long total = (long long)c_out * c_in * inner;
  vg::Launch(ew_grid(total), 256, 0, as_stream(stream))(fold_bn_into_conv_kernel, w, bias_in, gamma, beta, running_mean, running_var, eps, c_out, c_in,
                                                                          inner, transposed, w_out, bias_out);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, int c,
                                 float* scale, float* shift, vg_stream_t stream) {
  VG_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && c > 0, "bad args");
  vg::Launch((c + 127) / 128, 128, 0, as_stream(stream))(bn_eval_affine_kernel, gamma, beta, running_mean, running_var, eps, c, scale, shift);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_scale(const void* src, const float* scale, long long n, int dtype, void* dst, vg_stream_t stream) {
  VG_CHECK_ARG(src && scale && dst && n >= 0, "bad args");
  VG_CHECK_ARG(dtype == VG_F32 || dtype == VG_BF16, "bad dtype %d", dtype);
  if (n == 0) return VG_OK;
  if (dtype == VG_BF16)
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(scale_kernel<__nv_bfloat16>, (const __nv_bfloat16*)src, scale, n, (__nv_bfloat16*)dst);
  else
    vg::Launch(ew_grid(n), 256, 0, as_stream(stream))(scale_kernel<float>, (const float*)src, scale, n, (float*)dst);
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_normalize_images(const void* raw, int raw_dtype, int n, long long pixels_per_image, float* out_f32, void* out_bf16,
                                   vg_stream_t stream) {
  VG_CHECK_ARG(raw && n >= 0 && pixels_per_image > 0 && (out_f32 || out_bf16), "bad args");
  if (n == 0) return VG_OK;
  cudaStream_t s = as_stream(stream);
  __nv_bfloat16* ob = (__nv_bfloat16*)out_bf16;
  switch (raw_dtype) {
    case VG_RAW_U8: vg::Launch(n, 256, 0, s)(normalize_images_kernel<uint8_t>, (const uint8_t*)raw, pixels_per_image, out_f32, ob); break;
    case VG_RAW_U16: vg::Launch(n, 256, 0, s)(normalize_images_kernel<uint16_t>, (const uint16_t*)raw, pixels_per_image, out_f32, ob); break;
    case VG_RAW_I16: vg::Launch(n, 256, 0, s)(normalize_images_kernel<int16_t>, (const int16_t*)raw, pixels_per_image, out_f32, ob); break;
    case VG_RAW_F32: vg::Launch(n, 256, 0, s)(normalize_images_kernel<float>, (const float*)raw, pixels_per_image, out_f32, ob); break;
    case VG_RAW_F64: vg::Launch(n, 256, 0, s)(normalize_images_kernel<double>, (const double*)raw, pixels_per_image, out_f32, ob); break;
    default: set_error("unknown raw dtype %d", raw_dtype); return VG_EINVAL;
  }
  VG_LAUNCHED();
  return VG_OK;
}

extern "C" int vg_fill_zero(void* p, size_t bytes, vg_stream_t stream) {
  VG_CHECK_ARG(p || bytes == 0, "null pointer");
  if (bytes == 0) return VG_OK;
  VG_CUDA(cudaMemsetAsync(p, 0, bytes, as_stream(stream)));
  return VG_OK;
}
