// fp32 dense layers of the generator's mapping/style path (styleganv1.py:471-495 `FC`: F.linear with equalised-lr
// multipliers followed by leaky_relu(0.2)) and the emotion head (model.py:41,121-122).  The batch is tiny (<= 64 rows)
// and the path is 0.02 GF/sample but numerically sensitive, so these stay in fp32 FFMA with warp-shuffle
// reductions instead of going to the tensor cores.
#include "host_util.h"
#include "ptx.cuh"

namespace irfd {

constexpr int kMaxRows = 64;

// y[b, n] = act( wmul * sum_k x[b,k] * W[n,k] + bmul * bias[n] )
// One block (4 warps) per output column n: the K range is interleaved over the 128 threads in float4 steps, so a
// 512-wide layer is ONE load round trip per thread (the first version gave a whole column to one warp and walked K
// serially: ~30 us of load latency per layer on a path of 16 dependent layers per generator call).  The four warps'
// partial dot products are combined through shared memory in a fixed order.
constexpr int kFcWarps = 4;
template <int ROWS>
__global__ void __launch_bounds__(32 * kFcWarps)
linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ y, int B, int N, int K, float wmul, float bmul, int lrelu) {
  __shared__ float part[kFcWarps][ROWS];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < B; b0 += ROWS) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    for (int k = threadIdx.x * 4; k < K; k += 128 * kFcWarps) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k));
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (b0 + r < B) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b0 + r) * K + k));
          acc[r] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float s = warp_sum(acc[r]);
      if (lane == 0) part[warp][r] = s;
    }
    __syncthreads();
    if (threadIdx.x < ROWS && b0 + threadIdx.x < B) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kFcWarps; ++w) s += part[w][threadIdx.x];
      float v = s * wmul + (bias != nullptr ? bias[n] * bmul : 0.f);
      if (lrelu) v = v > 0.f ? v : 0.2f * v;
      y[(size_t)(b0 + threadIdx.x) * N + n] = v;
    }
    __syncthreads();
  }
}

// Long-K variant (the 6144 -> 512 first mapping layer): COLS output columns per block, so the [B, K] activations go
// through L2 N / COLS times instead of N times (168 us -> the x traffic of 805 MB drops 4x).  Per output the k partition
// per thread and the reduction order are those of linear_fwd_kernel: bit-identical results.  The ROWS-row passes over
// the batch are blockIdx.y (the first version looped over them inside 128 four-warp blocks: less than one block per SM,
// twelve dependent L2 round trips per pass, 190 us on the critical path between the encoders and the synthesis).
template <int ROWS, int COLS>
__global__ void __launch_bounds__(32 * kFcWarps)
linear_fwd_cols_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                       float* __restrict__ y, int B, int N, int K, float wmul, float bmul, int lrelu) {
  __shared__ float part[kFcWarps][ROWS][COLS];
  const int n0 = blockIdx.x * COLS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const int b0 = blockIdx.y * ROWS;
    float acc[ROWS][COLS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[r][c] = 0.f;
    for (int k = threadIdx.x * 4; k < K; k += 128 * kFcWarps) {
      float4 wv[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c)
        wv[c] = n0 + c < N ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + c) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (b0 + r < B) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)(b0 + r) * K + k));
#pragma unroll
          for (int c = 0; c < COLS; ++c) acc[r][c] += xv.x * wv[c].x + xv.y * wv[c].y + xv.z * wv[c].z + xv.w * wv[c].w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const float s = warp_sum(acc[r][c]);
        if (lane == 0) part[warp][r][c] = s;
      }
    __syncthreads();
    if (threadIdx.x < ROWS * COLS) {
      const int r = threadIdx.x / COLS, c = threadIdx.x % COLS;
      if (b0 + r < B && n0 + c < N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFcWarps; ++w) s += part[w][r][c];
        float v = s * wmul + (bias != nullptr ? bias[n0 + c] * bmul : 0.f);
        if (lrelu) v = v > 0.f ? v : 0.2f * v;
        y[(size_t)(b0 + r) * N + n0 + c] = v;
      }
    }
  }
}

// dz = dy * (y > 0 ? 1 : 0.2)   (y is the post-activation output; lrelu preserves sign)
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                                 size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = dy[i] * (y[i] > 0.f ? 1.f : 0.2f);
}

// dx[b,k] = beta*dx[b,k] + wmul * sum_n dz[b,n] * W[n,k]
// grid (K/32, B/8), 8 warps per block: the block owns an [8 rows x 32 k] output tile, each warp reduces one eighth of
// n (lanes over k: coalesced 128-byte reads of W rows; the 8 x N slice of dz is staged in shared memory), then the
// eight partial tiles are summed through shared memory in a fixed order.
constexpr int kDxRows = 8;
__global__ void __launch_bounds__(256)
linear_dx_kernel(const float* __restrict__ dz, const float* __restrict__ W, float* __restrict__ dx, int B, int N,
                 int K, float wmul, float beta) {
  extern __shared__ float sdz[];  // [kDxRows][N]
  __shared__ float red[8][kDxRows][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int b0 = blockIdx.y * kDxRows;
  for (int i = threadIdx.x; i < kDxRows * N; i += 256) {
    const int r = i / N, n = i - r * N;
    sdz[i] = (b0 + r < B) ? dz[(size_t)(b0 + r) * N + n] : 0.f;
  }
  __syncthreads();
  float acc[kDxRows];
#pragma unroll
  for (int r = 0; r < kDxRows; ++r) acc[r] = 0.f;
  const int per = (N + 7) / 8;
  const int n0 = warp * per;
  const int n1 = n0 + per < N ? n0 + per : N;
  if (k < K) {
#pragma unroll 4
    for (int n = n0; n < n1; ++n) {
      const float wv = W[(size_t)n * K + k];
#pragma unroll
      for (int r = 0; r < kDxRows; ++r) acc[r] += sdz[r * N + n] * wv;
    }
  }
#pragma unroll
  for (int r = 0; r < kDxRows; ++r) red[warp][r][lane] = acc[r];
  __syncthreads();
  // 256 threads finish the 8 x 32 tile: thread -> (row, lane)
  const int r = threadIdx.x >> 5;
  const int kk = blockIdx.x * 32 + lane;
  if (kk < K && b0 + r < B) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][r][lane];
    float* d = dx + (size_t)(b0 + r) * K + kk;
    *d = (beta != 0.f ? beta * (*d) : 0.f) + wmul * s;
  }
}

// dW[n,k] = beta*dW + wmul * sum_b dz[b,n]*x[b,k] ;  db[n] = beta*db + bmul * sum_b dz[b,n]
// A thread owns one k and kDwCols consecutive n: x[b][k] is loaded once per b for all of them (the first version gave every
// n its own block row, so the 64 x K slice of x went through L2 N times: 805 MB and 176 us for the 6144 -> 512 layer).
// Same summation order per output (b ascending): results are bit-identical to the one-output-per-thread version.
constexpr int kDwCols = 8;
__global__ void __launch_bounds__(128)
linear_dw_kernel(const float* __restrict__ dz, const float* __restrict__ x, float* __restrict__ dW,
                 float* __restrict__ db, int B, int N, int K, float wmul, float bmul, float beta) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = blockIdx.y * kDwCols;
  if (k < K) {
    float s[kDwCols];
#pragma unroll
    for (int j = 0; j < kDwCols; ++j) s[j] = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) {
      const float xv = __ldg(x + (size_t)b * K + k);
#pragma unroll
      for (int j = 0; j < kDwCols; ++j)
        if (n0 + j < N) s[j] += __ldg(dz + (size_t)b * N + n0 + j) * xv;
    }
#pragma unroll
    for (int j = 0; j < kDwCols; ++j)
      if (n0 + j < N) {
        float* d = dW + (size_t)(n0 + j) * K + k;
        *d = (beta != 0.f ? beta * (*d) : 0.f) + wmul * s[j];
      }
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < kDwCols && n0 + (int)threadIdx.x < N) {
    const int n = n0 + threadIdx.x;
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dz[(size_t)b * N + n];
    db[n] = (beta != 0.f ? beta * db[n] : 0.f) + bmul * t;
  }
}

// row-wise softmax for the emotion head (N = 8)
__global__ void softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = -INFINITY;
  for (int i = 0; i < N; ++i) m = fmaxf(m, x[(size_t)b * N + i]);
  float s = 0.f;
  for (int i = 0; i < N; ++i) s += expf(x[(size_t)b * N + i] - m);
  for (int i = 0; i < N; ++i) y[(size_t)b * N + i] = expf(x[(size_t)b * N + i] - m) / s;
}

// w_rows: out[b][k] = scale * (use_second ? w2[b][k] : w[b][k])  — helper for style mixing rows (styleganv1.py:536-553)
__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, float scale, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = scale * src[i];
}

// sp1 = style[:, :C] + 1 ; s1 = style[:, C:]   (ApplyStyle view(-1, 2, C), styleganv1.py:464-467)
__global__ void split_style_kernel(const float* __restrict__ style, float* __restrict__ sp1, float* __restrict__ s1,
                                   int B, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C) return;
  const int b = i / C, c = i % C;
  sp1[i] = style[(size_t)b * 2 * C + c] + 1.f;
  s1[i] = style[(size_t)b * 2 * C + C + c];
}
__global__ void merge_style_grad_kernel(const float* __restrict__ dsp1, const float* __restrict__ ds1,
                                        float* __restrict__ dstyle, int B, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C) return;
  const int b = i / C, c = i % C;
  dstyle[(size_t)b * 2 * C + c] = dsp1[i];
  dstyle[(size_t)b * 2 * C + C + c] = ds1[i];
}

}  // namespace irfd

using namespace irfd;

extern "C" int irfd_linear_fwd(const float* x, const float* w, const float* bias, float* y, int b, int n, int k,
                               float wmul, float bmul, int lrelu, cudaStream_t stream) {
  IRFD_CHECK_ARG(x && w && y && b > 0 && n > 0 && k > 0 && k % 4 == 0, "linear_fwd: bad argument (K %% 4 == 0)");
  const dim3 grid(n), block(32 * kFcWarps);
  if (k >= 2048 && b > 8) {  // long K: four columns per block (x traffic through L2 / 4)
    linear_fwd_cols_kernel<16, 4><<<dim3((n + 3) / 4, (b + 15) / 16), block, 0, stream>>>(x, w, bias, y, b, n, k, wmul, bmul,
                                                                                   lrelu);
    IRFD_CHECK_LAUNCH();
    return IRFD_OK;
  }
  // rows per pass: each pass streams the weight row once, so cover the whole batch in as few passes as possible
  if (b > 16)
    linear_fwd_kernel<32><<<grid, block, 0, stream>>>(x, w, bias, y, b, n, k, wmul, bmul, lrelu);
  else if (b > 8)
    linear_fwd_kernel<16><<<grid, block, 0, stream>>>(x, w, bias, y, b, n, k, wmul, bmul, lrelu);
  else
    linear_fwd_kernel<8><<<grid, block, 0, stream>>>(x, w, bias, y, b, n, k, wmul, bmul, lrelu);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_lrelu_bwd(const float* dy, const float* y, float* dz, long long n, cudaStream_t stream) {
  IRFD_CHECK_ARG(dy && y && dz && n > 0, "lrelu_bwd: bad argument");
  lrelu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dy, y, dz, (size_t)n);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_linear_bwd(const float* dz, const float* x, const float* w, float* dx, float dx_beta, float* dw,
                               float* db, float dw_beta, int b, int n, int k, float wmul, float bmul,
                               cudaStream_t stream) {
  IRFD_CHECK_ARG(dz && b > 0 && b <= kMaxRows && n > 0 && k > 0, "linear_bwd: bad argument (batch <= 64)");
  if (dx != nullptr) {
    IRFD_CHECK_ARG(w != nullptr, "linear_bwd: dx needs w");
    IRFD_CHECK_ARG((size_t)kDxRows * n * sizeof(float) <= 40 * 1024, "linear_bwd: N too large for the dz tile");
    linear_dx_kernel<<<dim3((k + 31) / 32, (b + kDxRows - 1) / kDxRows), 256, kDxRows * n * sizeof(float), stream>>>(
        dz, w, dx, b, n, k, wmul, dx_beta);
    IRFD_CHECK_LAUNCH();
  }
  if (dw != nullptr) {
    IRFD_CHECK_ARG(x != nullptr, "linear_bwd: dw needs x");
    linear_dw_kernel<<<dim3((k + 127) / 128, (n + kDwCols - 1) / kDwCols), 128, 0, stream>>>(dz, x, dw, db, b, n, k, wmul, bmul,
                                                                                      dw_beta);
    IRFD_CHECK_LAUNCH();
  }
  return IRFD_OK;
}

extern "C" int irfd_softmax_rows(const float* x, float* y, int b, int n, cudaStream_t stream) {
  IRFD_CHECK_ARG(x && y && b > 0 && n > 0, "softmax_rows: bad argument");
  softmax_rows_kernel<<<(b + 63) / 64, 64, 0, stream>>>(x, y, b, n);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_scale_copy(const float* src, float* dst, float scale, long long n, cudaStream_t stream) {
  IRFD_CHECK_ARG(src && dst && n > 0, "scale_copy: bad argument");
  scale_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, scale, (size_t)n);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_split_style(const float* style, float* sp1, float* s1, int b, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(style && sp1 && s1, "split_style: null pointer");
  split_style_kernel<<<(unsigned)(((size_t)b * c + 255) / 256), 256, 0, stream>>>(style, sp1, s1, b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_merge_style_grad(const float* dsp1, const float* ds1, float* dstyle, int b, int c,
                                     cudaStream_t stream) {
  IRFD_CHECK_ARG(dsp1 && ds1 && dstyle, "merge_style_grad: null pointer");
  merge_style_grad_kernel<<<(unsigned)(((size_t)b * c + 255) / 256), 256, 0, stream>>>(dsp1, ds1, dstyle, b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
