// BatchNorm2d (train + eval) around the conv GEMMs, NHWC bf16 activations, fp32 statistics.
// Reference semantics: torch.nn.BatchNorm2d inside torchvision Bottleneck (torchvision/models/resnet.py:143-164):
// batch statistics in train mode (biased variance for normalisation, unbiased for the running buffer, momentum 0.1,
// eps 1e-5), running statistics in eval mode; ReLU and the residual add are fused into the apply pass.
#include "host_util.h"
#include "rowvec.cuh"

namespace irfd {

// ---------------------------------------------------------------------------------------------------------------
// Statistics finalize: per-tile partial sums (from the conv epilogue) -> mean / rstd, running-buffer update.
// One block per 32 channels, 8 tile-lanes; accumulation in double.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxGroups = 4;

__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int tiles, int C, double count,
                   float eps, float momentum, float* __restrict__ mean, float* __restrict__ rstd, float* running_mean,
                   float* running_var, int running_updates, int groups) {
  // block = 32 channels (lanes, coalesced 128-byte reads) x 32 warps striding over the tiles; statistic groups (e.g.
  // the source and the target half of a paired encoder pass) are processed one after the other so the running
  // buffers can be updated in call order by the same thread.
  __shared__ double s_sum[32][33];
  __shared__ double s_sq[32][33];
  const int cl = threadIdx.x & 31;
  const int tl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double gm[kMaxGroups], gv[kMaxGroups];
  for (int g = 0; g < groups; ++g) {
    const float* ps = psum + (size_t)g * tiles * C;
    const float* pq = psq + (size_t)g * tiles * C;
    double a = 0.0, b = 0.0;
    if (c < C) {
      int t = tl;
      for (; t + 96 < tiles; t += 128) {  // 4 independent loads in flight per stream
        const float a0 = ps[(size_t)t * C + c], a1 = ps[(size_t)(t + 32) * C + c];
        const float a2 = ps[(size_t)(t + 64) * C + c], a3 = ps[(size_t)(t + 96) * C + c];
        const float b0 = pq[(size_t)t * C + c], b1 = pq[(size_t)(t + 32) * C + c];
        const float b2 = pq[(size_t)(t + 64) * C + c], b3 = pq[(size_t)(t + 96) * C + c];
        a += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
        b += ((double)b0 + (double)b1) + ((double)b2 + (double)b3);
      }
      for (; t < tiles; t += 32) {
        a += (double)ps[(size_t)t * C + c];
        b += (double)pq[(size_t)t * C + c];
      }
    }
    __syncthreads();
    s_sum[tl][cl] = a;
    s_sq[tl][cl] = b;
    __syncthreads();
    if (tl == 0 && c < C) {
      for (int i = 1; i < 32; ++i) {
        a += s_sum[i][cl];
        b += s_sq[i][cl];
      }
      const double m = a / count;
      double var = b / count - m * m;
      if (var < 0.0) var = 0.0;
      mean[(size_t)g * C + c] = (float)m;
      rstd[(size_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
      gm[g] = m;
      gv[g] = count > 1.0 ? var * count / (count - 1.0) : var;  // unbiased, for the running buffer
    }
  }
  if (tl == 0 && c < C && running_mean != nullptr) {
    // update 1: the forward calls in order (group 0 first).  update 2 (optional): what the reference's reentrant
    // checkpoint adds when it re-runs each forward during backward, i.e. the same statistics in reverse call order
    // (model.py:84-90).
    float rm = running_mean[c], rv = running_var[c];
    for (int u = 0; u < running_updates; ++u)
      for (int i = 0; i < groups; ++i) {
        const int g = (u & 1) ? groups - 1 - i : i;
        rm = (1.f - momentum) * rm + momentum * (float)gm[g];
        rv = (1.f - momentum) * rv + momentum * (float)gv[g];
      }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
}

// Re-apply the momentum update from saved batch statistics (second update of the reference's checkpoint recompute).
__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, float eps,
                                         double count, float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double r = (double)rstd[c];
  double var = 1.0 / (r * r) - (double)eps;
  if (var < 0.0) var = 0.0;
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
  running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
}

// eval mode folded into a conv epilogue: scale = gamma/sqrt(var+eps), shift = beta - mean*scale
__global__ void bn_eval_affine_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

// eval mode: rstd from the running variance
__global__ void bn_eval_rstd_kernel(const float* __restrict__ running_var, float eps, float* __restrict__ rstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) rstd[c] = rsqrtf(running_var[c] + eps);
}

// ---------------------------------------------------------------------------------------------------------------
// Apply: out = relu( (z - mean)*rstd*gamma + beta  [+ res | + BN2(z2)] )
// ---------------------------------------------------------------------------------------------------------------
struct BnAffine {
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* beta;
};

template <int RES_MODE>  // 0 none, 1 identity tensor, 2 second BN branch
__global__ void __launch_bounds__(kRvThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ z, BnAffine p1, const __nv_bfloat16* __restrict__ res, BnAffine p2,
                __nv_bfloat16* __restrict__ out, long long rows, int C, int rows_per_blk, int relu) {
  RowVec rv(C);
  if (!rv.active) return;
  {  // statistic group = blockIdx.y: `rows` rows each, own mean/rstd (gamma/beta shared)
    const size_t goff = (size_t)blockIdx.y * rows * C;
    z += goff;
    out += goff;
    if (RES_MODE != 0) res += goff;
    p1.mean += (size_t)blockIdx.y * C;
    p1.rstd += (size_t)blockIdx.y * C;
    if (RES_MODE == 2) {
      p2.mean += (size_t)blockIdx.y * C;
      p2.rstd += (size_t)blockIdx.y * C;
    }
  }
  float sc[8], sh[8], sc2[8], sh2[8];
  {
    float m[8], r[8], g[8], b[8];
    loadf8(p1.mean + rv.cv * 8, m);
    loadf8(p1.rstd + rv.cv * 8, r);
    loadf8(p1.gamma + rv.cv * 8, g);
    loadf8(p1.beta + rv.cv * 8, b);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      sc[t] = g[t] * r[t];
      sh[t] = b[t] - m[t] * sc[t];
    }
    if (RES_MODE == 2) {
      loadf8(p2.mean + rv.cv * 8, m);
      loadf8(p2.rstd + rv.cv * 8, r);
      loadf8(p2.gamma + rv.cv * 8, g);
      loadf8(p2.beta + rv.cv * 8, b);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        sc2[t] = g[t] * r[t];
        sh2[t] = b[t] - m[t] * sc2[t];
      }
    }
  }
  const long long r0 = (long long)blockIdx.x * rows_per_blk;
  long long r1 = r0 + rows_per_blk;
  if (r1 > rows) r1 = rows;
  for (long long r = r0 + rv.row_lane; r < r1; r += rv.rows_par) {
    const size_t off = (size_t)r * C + rv.cv * 8;
    float v[8], o[8];
    load8(z + off, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = v[t] * sc[t] + sh[t];
    if (RES_MODE != 0) {
      float w[8];
      load8(res + off, w);
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] += (RES_MODE == 2) ? (w[t] * sc2[t] + sh2[t]) : w[t];
    }
    if (relu) {
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
    }
    store8(out + off, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  g = (g1 [+ g2]) * (act > 0) where `act` is the post-ReLU tensor of the forward pass.
//   reduce  : per-block partials of  sum(g), sum(g * xhat)            xhat = (z - mean) * rstd
//   finalize: dgamma = sum(g*xhat), dbeta = sum(g); c1 = dbeta/n, c2 = dgamma/n
//   apply   : dz = gamma*rstd * (g - c1 - xhat*c2)     (+ optional copy of the masked g for the identity shortcut)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRvThreads)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g1, const __nv_bfloat16* __restrict__ g2,
                     const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ partial, long long rows, int C,
                     int rows_per_blk) {
  extern __shared__ float red_smem[];
  RowVec rv(C);
  {
    const size_t goff = (size_t)blockIdx.y * rows * C;
    g1 += goff;
    if (g2 != nullptr) g2 += goff;
    if (act != nullptr) act += goff;
    z += goff;
    mean += (size_t)blockIdx.y * C;
    rstd += (size_t)blockIdx.y * C;
    partial += (size_t)blockIdx.y * gridDim.x * 2 * C;
  }
  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = acc[1][t] = 0.f;
  if (rv.active) {
    float m[8], rs[8], ga[8], be[8];
    loadf8(mean + rv.cv * 8, m);
    loadf8(rstd + rv.cv * 8, rs);
    if (beta != nullptr) {  // ReLU mask recomputed from z: relu(gamma*xhat + beta) > 0  (no residual on this BN)
      loadf8(gamma + rv.cv * 8, ga);
      loadf8(beta + rv.cv * 8, be);
    }
    const long long r0 = (long long)blockIdx.x * rows_per_blk;
    long long r1 = r0 + rows_per_blk;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + rv.row_lane; r < r1; r += rv.rows_par) {
      const size_t off = (size_t)r * C + rv.cv * 8;
      float g[8], a[8], zz[8];
      load8(g1 + off, g);
      if (g2 != nullptr) {
        float h[8];
        load8(g2 + off, h);
#pragma unroll
        for (int t = 0; t < 8; ++t) g[t] += h[t];
      }
      load8(z + off, zz);
      if (act != nullptr) {
        load8(act + off, a);
#pragma unroll
        for (int t = 0; t < 8; ++t) g[t] = a[t] > 0.f ? g[t] : 0.f;
      } else if (beta != nullptr) {
#pragma unroll
        for (int t = 0; t < 8; ++t) g[t] = (ga[t] * ((zz[t] - m[t]) * rs[t]) + be[t]) > 0.f ? g[t] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        acc[0][t] += g[t];
        acc[1][t] += g[t] * ((zz[t] - m[t]) * rs[t]);
      }
    }
  }
  block_reduce_rows<2>(rv, C, acc, red_smem, partial + (size_t)blockIdx.x * 2 * C, (size_t)C);
}

__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, double count, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float beta_acc, float* __restrict__ c1, float* __restrict__ c2,
                       int batch_stats, int groups) {
  __shared__ double s0s[32][33];
  __shared__ double s1s[32][33];
  const int cl = threadIdx.x & 31;
  const int tl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double tot0 = 0.0, tot1 = 0.0;
  for (int g = 0; g < groups; ++g) {
    const float* pg = partial + (size_t)g * nblk * 2 * C;
    double s0 = 0.0, s1 = 0.0;
    if (c < C) {
      for (int b = tl; b < nblk; b += 32) {
        s0 += (double)pg[((size_t)b * 2 + 0) * C + c];
        s1 += (double)pg[((size_t)b * 2 + 1) * C + c];
      }
    }
    __syncthreads();
    s0s[tl][cl] = s0;
    s1s[tl][cl] = s1;
    __syncthreads();
    if (tl == 0 && c < C) {
      for (int i = 1; i < 32; ++i) {
        s0 += s0s[i][cl];
        s1 += s1s[i][cl];
      }
      // each group normalised with its own statistics; eval mode (constants) has no batch-statistic terms in dz
      c1[(size_t)g * C + c] = batch_stats ? (float)(s0 / count) : 0.f;
      c2[(size_t)g * C + c] = batch_stats ? (float)(s1 / count) : 0.f;
      tot0 += s0;
      tot1 += s1;
    }
  }
  if (tl == 0 && c < C) {
    dbeta[c] = (beta_acc != 0.f ? beta_acc * dbeta[c] : 0.f) + (float)tot0;
    dgamma[c] = (beta_acc != 0.f ? beta_acc * dgamma[c] : 0.f) + (float)tot1;
  }
}

__global__ void __launch_bounds__(kRvThreads)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g1, const __nv_bfloat16* __restrict__ g2,
                    const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ c1, const float* __restrict__ c2,
                    __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ g_out, long long rows, int C,
                    int rows_per_blk) {
  RowVec rv(C);
  if (!rv.active) return;
  {
    const size_t goff = (size_t)blockIdx.y * rows * C;
    g1 += goff;
    if (g2 != nullptr) g2 += goff;
    if (act != nullptr) act += goff;
    z += goff;
    dz += goff;
    if (g_out != nullptr) g_out += goff;
    mean += (size_t)blockIdx.y * C;
    rstd += (size_t)blockIdx.y * C;
    c1 += (size_t)blockIdx.y * C;
    c2 += (size_t)blockIdx.y * C;
  }
  float m[8], rs[8], k0[8], k1[8], k2[8], ga[8], be[8];
  if (beta != nullptr) loadf8(beta + rv.cv * 8, be);
  {
    float a1[8], a2[8];
    loadf8(mean + rv.cv * 8, m);
    loadf8(rstd + rv.cv * 8, rs);
    loadf8(gamma + rv.cv * 8, ga);
    loadf8(c1 + rv.cv * 8, a1);
    loadf8(c2 + rv.cv * 8, a2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      k0[t] = ga[t] * rs[t];
      k1[t] = a1[t];
      k2[t] = a2[t];
    }
  }
  const long long r0 = (long long)blockIdx.x * rows_per_blk;
  long long r1 = r0 + rows_per_blk;
  if (r1 > rows) r1 = rows;
  for (long long r = r0 + rv.row_lane; r < r1; r += rv.rows_par) {
    const size_t off = (size_t)r * C + rv.cv * 8;
    float g[8], zz[8], o[8];
    load8(g1 + off, g);
    if (g2 != nullptr) {
      float h[8];
      load8(g2 + off, h);
#pragma unroll
      for (int t = 0; t < 8; ++t) g[t] += h[t];
    }
    load8(z + off, zz);
    if (act != nullptr) {
      float a[8];
      load8(act + off, a);
#pragma unroll
      for (int t = 0; t < 8; ++t) g[t] = a[t] > 0.f ? g[t] : 0.f;
    } else if (beta != nullptr) {
#pragma unroll
      for (int t = 0; t < 8; ++t) g[t] = (ga[t] * ((zz[t] - m[t]) * rs[t]) + be[t]) > 0.f ? g[t] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = k0[t] * (g[t] - k1[t] - (zz[t] - m[t]) * rs[t] * k2[t]);
    store8(dz + off, o);
    if (g_out != nullptr) store8(g_out + off, g);
  }
}

}  // namespace irfd

using namespace irfd;

// tiles / count are PER GROUP; mean/rstd are [groups][c].  The running buffers get `running_updates` rounds of momentum
// updates: round 1 walks the groups in call order, round 2 (the reference's checkpoint recompute) in reverse order.
extern "C" int irfd_bn_finalize(const float* psum, const float* psq, int tiles, int c, long long count, float eps,
                                float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                                int running_updates, int groups, cudaStream_t stream) {
  IRFD_CHECK_ARG(psum && psq && mean && rstd && tiles > 0 && c > 0 && count > 0, "bn_finalize: bad argument");
  IRFD_CHECK_ARG(groups >= 1 && groups <= kMaxGroups, "bn_finalize: 1..4 statistic groups");
  bn_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>(psum, psq, tiles, c, (double)count, eps, momentum, mean, rstd,
                                                         running_mean, running_var, running_updates, groups);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_running_update(const float* mean, const float* rstd, float eps, long long count, float momentum,
                                      float* running_mean, float* running_var, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(mean && rstd && running_mean && running_var && c > 0 && count > 0, "bn_running_update: bad argument");
  bn_running_update_kernel<<<(c + 255) / 256, 256, 0, stream>>>(mean, rstd, eps, (double)count, momentum, running_mean,
                                                                 running_var, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_eval_affine(const float* running_mean, const float* running_var, const float* gamma,
                                   const float* beta, float eps, float* scale, float* shift, int c,
                                   cudaStream_t stream) {
  IRFD_CHECK_ARG(running_mean && running_var && gamma && beta && scale && shift && c > 0, "bn_eval_affine: bad argument");
  bn_eval_affine_kernel<<<(c + 255) / 256, 256, 0, stream>>>(running_mean, running_var, gamma, beta, eps, scale, shift, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_eval_rstd(const float* running_var, float eps, float* rstd, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(running_var && rstd && c > 0, "bn_eval_rstd: bad argument");
  bn_eval_rstd_kernel<<<(c + 255) / 256, 256, 0, stream>>>(running_var, eps, rstd, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

// rows = TOTAL rows (all groups, groups stacked along the row axis); mean/rstd (and mean2/rstd2) are [groups][c].
extern "C" int irfd_bn_apply(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                             const void* res, const float* mean2, const float* rstd2, const float* gamma2,
                             const float* beta2, void* out, long long rows, int c, int relu, int groups,
                             cudaStream_t stream) {
  IRFD_CHECK_ARG(z && mean && rstd && gamma && beta && out, "bn_apply: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048 && rows > 0, "bn_apply: C must be a multiple of 8 and <= 2048");
  IRFD_CHECK_ARG(groups >= 1 && rows % groups == 0, "bn_apply: rows must split evenly into groups");
  const long long grows = rows / groups;
  int nblk, rpb;
  plan_row_blocks(grows, c, num_sms(), &nblk, &rpb);
  BnAffine p1{mean, rstd, gamma, beta}, p2{mean2, rstd2, gamma2, beta2};
  auto zz = reinterpret_cast<const __nv_bfloat16*>(z);
  auto rr = reinterpret_cast<const __nv_bfloat16*>(res);
  auto oo = reinterpret_cast<__nv_bfloat16*>(out);
  const dim3 grid(nblk, groups);
  if (res == nullptr)
    bn_apply_kernel<0><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, grows, c, rpb, relu);
  else if (mean2 == nullptr)
    bn_apply_kernel<1><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, grows, c, rpb, relu);
  else
    bn_apply_kernel<2><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, grows, c, rpb, relu);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" long long irfd_bn_bwd_workspace_bytes(long long rows, int c, int groups) {
  if (groups < 1) groups = 1;
  int nblk, rpb;
  plan_row_blocks(rows / groups, c, num_sms(), &nblk, &rpb);
  return (long long)(nblk * 2 + 2) * groups * c * 4;
}

// Full BN backward (reduce -> finalize -> apply).  workspace: [groups][nblk][2][C] partials, c1[groups][C], c2[...].
extern "C" int irfd_bn_backward(const void* g1, const void* g2, const void* act, const void* z, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, void* dz, void* g_out,
                                float* dgamma, float* dbeta, float grad_beta, int batch_stats, long long rows, int c,
                                int groups, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(g1 && z && mean && rstd && gamma && dz && dgamma && dbeta && workspace, "bn_backward: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048 && rows > 0, "bn_backward: C must be a multiple of 8 and <= 2048");
  IRFD_CHECK_ARG(groups >= 1 && rows % groups == 0, "bn_backward: rows must split evenly into groups");
  const long long grows = rows / groups;
  int nblk, rpb;
  plan_row_blocks(grows, c, num_sms(), &nblk, &rpb);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)(nblk * 2 + 2) * groups * c * 4, "bn_backward: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  float* c1 = partial + (size_t)groups * nblk * 2 * c;
  float* c2 = c1 + (size_t)groups * c;
  int rows_par = kRvThreads / (c / 8);
  if (rows_par < 1) rows_par = 1;
  const size_t smem = (size_t)2 * rows_par * c * sizeof(float);
  auto G1 = reinterpret_cast<const __nv_bfloat16*>(g1);
  auto G2 = reinterpret_cast<const __nv_bfloat16*>(g2);
  auto A = reinterpret_cast<const __nv_bfloat16*>(act);
  auto Z = reinterpret_cast<const __nv_bfloat16*>(z);
  const dim3 grid(nblk, groups);
  bn_bwd_reduce_kernel<<<grid, kRvThreads, smem, stream>>>(G1, G2, A, Z, mean, rstd, gamma, beta, partial, grows, c,
                                                           rpb);
  IRFD_CHECK_LAUNCH();
  bn_bwd_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>(partial, nblk, c, (double)grows, dgamma, dbeta, grad_beta,
                                                               c1, c2, batch_stats, groups);
  IRFD_CHECK_LAUNCH();
  bn_bwd_apply_kernel<<<grid, kRvThreads, 0, stream>>>(G1, G2, A, Z, mean, rstd, gamma, beta, c1, c2,
                                                        reinterpret_cast<__nv_bfloat16*>(dz),
                                                        reinterpret_cast<__nv_bfloat16*>(g_out), grows, c, rpb);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
