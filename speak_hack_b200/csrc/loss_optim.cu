// MSE losses (model.py:356-372: nn.MSELoss means, the only differentiable IRFD losses) as warp-shuffle reductions,
// the gradient-norm reduction behind clip_grad_norm_ (train.py:207-208) and a fused Adam step (train.py:346).
#include "host_util.h"
#include "ptx.cuh"

namespace irfd {

// partial[blk] = sum over the block's slice of (a-b)^2, accumulated in double (values reach 1e4^2)
__global__ void mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ partial,
                                   size_t n) {
  __shared__ double s[32];
  double acc = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 x = *reinterpret_cast<const float4*>(a + i);
      const float4 y = *reinterpret_cast<const float4*>(b + i);
      const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
      acc += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
    } else {
      for (size_t j = i; j < n; ++j) {
        const float d = a[j] - b[j];
        acc += (double)d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void mse_finalize_kernel(const double* __restrict__ partial, int nblk, double inv_n, float* __restrict__ out,
                                    float out_beta) {
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += partial[i];
  const float v = (float)(t * inv_n);
  out[0] = (out_beta != 0.f ? out_beta * out[0] : 0.f) + v;
}

// da = gscale[0] * 2 (a - b) / n  (and db = -da when requested); optional accumulate
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                               const float* __restrict__ gscale, float coef, float* __restrict__ da,
                               float* __restrict__ db, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = gscale[0] * coef * (a[i] - b[i]);
  if (da != nullptr) da[i] = g;
  if (db != nullptr) db[i] = -g;
}

// sum of squares of a flat fp32 buffer -> out[0] (double partials, fixed order)
__global__ void sumsq_partial_kernel(const float* __restrict__ g, double* __restrict__ partial, size_t n) {
  __shared__ double s[32];
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += (double)v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_finalize_kernel(const double* __restrict__ partial, int nblk, float* __restrict__ out,
                                      float out_beta) {
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += partial[i];
  out[0] = (out_beta != 0.f ? out_beta * out[0] : 0.f) + (float)t;
}

// torch.optim.Adam semantics (no amsgrad, no weight decay), bias correction from `step` (1-based).
// clip: if total_sumsq != null, grads are scaled by min(1, max_norm / (sqrt(total_sumsq[0]) + 1e-6)) first
// (torch.nn.utils.clip_grad_norm_ semantics).
__global__ void increment_kernel(int* counter) { counter[0] += 1; }

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, const float* __restrict__ total_sumsq, float max_norm,
                            const int* __restrict__ step_dev) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (step_dev != nullptr) {  // static-graph mode: the step number lives on the device
    const float t = (float)step_dev[0];
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  float scale = 1.f;
  if (total_sumsq != nullptr) {
    const float nrm = sqrtf(total_sumsq[0]);
    scale = fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  const float gi = g[i] * scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace irfd

using namespace irfd;

static int reduce_blocks(long long n) {
  long long b = (n + 1023) / 1024;
  const long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" long long irfd_reduce_workspace_bytes(void) { return (long long)num_sms() * 8 * 8; }

extern "C" int irfd_mse_fwd(const float* a, const float* b, long long n, float* out, float out_beta, void* workspace,
                            long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(a && b && out && workspace && n > 0, "mse_fwd: bad argument");
  const int nblk = reduce_blocks(n);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)nblk * 8, "mse_fwd: workspace too small");
  double* partial = reinterpret_cast<double*>(workspace);
  mse_partial_kernel<<<nblk, 256, 0, stream>>>(a, b, partial, (size_t)n);
  IRFD_CHECK_LAUNCH();
  mse_finalize_kernel<<<1, 1, 0, stream>>>(partial, nblk, 1.0 / (double)n, out, out_beta);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_mse_bwd(const float* a, const float* b, long long n, const float* gscale, float* da, float* db,
                            cudaStream_t stream) {
  IRFD_CHECK_ARG(a && b && gscale && n > 0 && (da || db), "mse_bwd: bad argument");
  mse_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(a, b, gscale, 2.f / (float)n, da, db, (size_t)n);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_sumsq(const float* g, long long n, float* out, float out_beta, void* workspace,
                          long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(g && out && workspace && n > 0, "sumsq: bad argument");
  const int nblk = reduce_blocks(n);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)nblk * 8, "sumsq: workspace too small");
  double* partial = reinterpret_cast<double*>(workspace);
  sumsq_partial_kernel<<<nblk, 256, 0, stream>>>(g, partial, (size_t)n);
  IRFD_CHECK_LAUNCH();
  sumsq_finalize_kernel<<<1, 1, 0, stream>>>(partial, nblk, out, out_beta);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                              float beta2, float eps, int step, int* step_dev, const float* total_sumsq,
                              float max_norm, cudaStream_t stream) {
  IRFD_CHECK_ARG(p && g && m && v && n > 0 && (step >= 1 || step_dev), "adam_step: bad argument");
  if (step_dev != nullptr) {
    increment_kernel<<<1, 1, 0, stream>>>(step_dev);
    IRFD_CHECK_LAUNCH();
  }
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, (size_t)n, lr, beta1, beta2, eps, bc1,
                                                                sqrtf(bc2), total_sumsq, max_norm, step_dev);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
