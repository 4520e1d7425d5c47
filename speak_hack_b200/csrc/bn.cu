// BatchNorm2d (train + eval) around the conv GEMMs, NHWC bf16 activations, fp32 statistics.
// Reference semantics: torch.nn.BatchNorm2d inside torchvision Bottleneck (torchvision/models/resnet.py:143-164):
// batch statistics in train mode (biased variance for normalisation, unbiased for the running buffer, momentum 0.1,
// eps 1e-5), running statistics in eval mode; ReLU and the residual add are fused into the apply pass.
#include "host_util.h"
#include "rowvec.cuh"

namespace irfd {

// ---------------------------------------------------------------------------------------------------------------
// Statistics finalize: per-tile partial sums (from the conv epilogue) -> mean / rstd, running-buffer update.
// One block per 8 channels (one 32-byte sector per tile row) x 128 tile lanes, 8 independent loads per stream in
// flight per thread; accumulation in double, fixed summation order (bit-deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxGroups = 4;
// Parameter SETS: several BatchNorm layers of identical shape normalised by one launch sequence (the same layer of the
// three IRFD encoders, model.py:33-35).  Statistic group y (blockIdx.y) belongs to set y / groups_per_set; per-channel
// parameters, running buffers and parameter gradients are per set (pointer arrays passed by value), batch statistics
// per group.
constexpr int kMaxSets = 4;
struct FSet {
  const float* p[kMaxSets];
};
struct FSetM {
  float* p[kMaxSets];
};
constexpr int kFinCh = 8;       // channels per block
constexpr int kFinLanes = 64;   // partial-row lanes per block
constexpr int kFinThreads = kFinCh * kFinLanes;  // 512
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kFinBatch = 8;    // loads in flight per thread per stream

// Sum `n` partial rows of two [n][C] fp32 arrays for channel c, this thread taking rows lane, lane+128, ...
__device__ __forceinline__ void fin_partial_sums(const float* __restrict__ p0, const float* __restrict__ p1, size_t stride0,
                                                 size_t stride1, int n, int lane, double& a, double& b) {
  a = 0.0;
  b = 0.0;
  for (int t = lane; t < n; t += kFinLanes * kFinBatch) {
    float x[kFinBatch], y[kFinBatch];
#pragma unroll
    for (int u = 0; u < kFinBatch; ++u) {
      const int tt = t + u * kFinLanes;
      x[u] = tt < n ? __ldg(p0 + (size_t)tt * stride0) : 0.f;
      y[u] = tt < n ? __ldg(p1 + (size_t)tt * stride1) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kFinBatch; ++u) {
      a += (double)x[u];
      b += (double)y[u];
    }
  }
}

// Block-wide sum over the lanes of one channel: lanes of a warp by shuffle, the warps through shared memory.
// Valid in the threads with lane == 0 (threadIdx.x < kFinCh).
__device__ __forceinline__ void fin_block_sums(double& a, double& b, double (*sa)[kFinCh], double (*sb)[kFinCh]) {
#pragma unroll
  for (int o = kFinCh; o < 32; o <<= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // previous use of sa/sb has been consumed
  if (l < kFinCh) {
    sa[w][l] = a;
    sb[w][l] = b;
  }
  __syncthreads();
  if (threadIdx.x < kFinCh) {
    a = 0.0;
    b = 0.0;
#pragma unroll 8
    for (int i = 0; i < kFinWarps; ++i) {
      a += sa[i][threadIdx.x];
      b += sb[i][threadIdx.x];
    }
  }
}

// 512-thread blocks, two per SM by the launch bound (<= 64 registers): the first version's 1024-thread block at 63
// registers needed the whole register file of an SM, so inside the step's graph it had to wait for a persistent
// weight-gradient CTA of the side stream to retire before it could start (CUPTI: 35 us per backward finalize against
// 19 us alone).
__global__ void __launch_bounds__(kFinThreads, 2)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int tiles, int C, double count,
                   float eps, float momentum, float* __restrict__ mean, float* __restrict__ rstd, FSetM running_mean_s,
                   FSetM running_var_s, int running_updates, int groups) {
  // statistic groups (e.g. the source and the target half of a paired encoder pass) are processed one after the
  // other so the running buffers can be updated in call order by the same thread.
  __shared__ double s_sum[kFinWarps][kFinCh];
  __shared__ double s_sq[kFinWarps][kFinCh];
  const int cl = threadIdx.x & (kFinCh - 1);
  const int lane = threadIdx.x / kFinCh;
  const int c = blockIdx.x * kFinCh + cl;  // C % 8 == 0 on this path
  // parameter set = blockIdx.y: its `groups` statistic groups are consecutive
  psum += (size_t)blockIdx.y * groups * tiles * C;
  psq += (size_t)blockIdx.y * groups * tiles * C;
  mean += (size_t)blockIdx.y * groups * C;
  rstd += (size_t)blockIdx.y * groups * C;
  float* running_mean = running_mean_s.p[blockIdx.y];
  float* running_var = running_var_s.p[blockIdx.y];
  double gm[kMaxGroups], gv[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    if (g < groups) {
      double a, b;
      fin_partial_sums(psum + (size_t)g * tiles * C + c, psq + (size_t)g * tiles * C + c, C, C, tiles, lane, a, b);
      fin_block_sums(a, b, s_sum, s_sq);
      if (threadIdx.x < kFinCh) {
        const double m = a / count;
        double var = b / count - m * m;
        if (var < 0.0) var = 0.0;
        mean[(size_t)g * C + c] = (float)m;
        rstd[(size_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps));
        gm[g] = m;
        gv[g] = count > 1.0 ? var * count / (count - 1.0) : var;  // unbiased, for the running buffer
      }
    }
  }
  if (threadIdx.x < kFinCh && running_mean != nullptr) {
    // update 1: the forward calls in order (group 0 first).  update 2 (optional): what the reference's reentrant
    // checkpoint adds when it re-runs each forward during backward, i.e. the same statistics in reverse call order
    // (model.py:84-90).
    float rm = running_mean[c], rv = running_var[c];
    for (int u = 0; u < running_updates; ++u)
#pragma unroll
      for (int i = 0; i < kMaxGroups; ++i) {
        if (i < groups) {
          const int g = (u & 1) ? groups - 1 - i : i;
          double mg = gm[0], vg = gv[0];
#pragma unroll
          for (int k = 1; k < kMaxGroups; ++k)
            if (k == g) {
              mg = gm[k];
              vg = gv[k];
            }
          rm = (1.f - momentum) * rm + momentum * (float)mg;
          rv = (1.f - momentum) * rv + momentum * (float)vg;
        }
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

// the same for `gridDim.y` parameter sets: scale/shift are [nsets][C]
__global__ void bn_eval_affine_sets_kernel(FSet running_mean, FSet running_var, FSet gamma, FSet beta, float eps,
                                           float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int s = blockIdx.y;
  const float sc = gamma.p[s][c] * rsqrtf(running_var.p[s][c] + eps);
  scale[(size_t)s * C + c] = sc;
  shift[(size_t)s * C + c] = beta.p[s][c] - running_mean.p[s][c] * sc;
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
  FSet gamma;
  FSet beta;
};

template <int RES_MODE>  // 0 none, 1 identity tensor, 2 second BN branch
__global__ void __launch_bounds__(kRvThreads)
bn_apply_kernel(const __nv_bfloat16* __restrict__ z, BnAffine p1, const __nv_bfloat16* __restrict__ res, BnAffine p2,
                __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ bits, long long rows, int C, int rows_per_blk,
                int relu, int gps, int reverse) {
  constexpr int RB = 8;  // rows per thread in flight: 8 (z only) or 16 (z + residual) 16-byte loads
  RowVec rv(C);
  if (!rv.active) return;
  // reverse: the first blocks to be scheduled take the LAST rows — the part of z its producer (the conv epilogue) wrote
  // last and that is still in the 126 MB L2 — and the rows written last here are the ones the next conv reads first
  const unsigned bx = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned by = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  {  // statistic group = by: `rows` rows each, own mean/rstd (gamma/beta shared)
    const size_t goff = (size_t)by * rows * C;
    z += goff;
    out += goff;
    if (bits != nullptr) bits += goff >> 3;
    if (RES_MODE != 0) res += goff;
    p1.mean += (size_t)by * C;
    p1.rstd += (size_t)by * C;
    if (RES_MODE == 2) {
      p2.mean += (size_t)by * C;
      p2.rstd += (size_t)by * C;
    }
  }
  float sc[8], sh[8], sc2[8], sh2[8];
  {
    float m[8], r[8], g[8], b[8];
    loadf8(p1.mean + rv.cv * 8, m);
    loadf8(p1.rstd + rv.cv * 8, r);
    const int set = by / gps;
    loadf8(p1.gamma.p[set] + rv.cv * 8, g);
    loadf8(p1.beta.p[set] + rv.cv * 8, b);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      sc[t] = g[t] * r[t];
      sh[t] = b[t] - m[t] * sc[t];
    }
    if (RES_MODE == 2) {
      loadf8(p2.mean + rv.cv * 8, m);
      loadf8(p2.rstd + rv.cv * 8, r);
      loadf8(p2.gamma.p[set] + rv.cv * 8, g);
      loadf8(p2.beta.p[set] + rv.cv * 8, b);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        sc2[t] = g[t] * r[t];
        sh2[t] = b[t] - m[t] * sc2[t];
      }
    }
  }
  const long long r0 = (long long)bx * rows_per_blk;
  long long r1 = r0 + rows_per_blk;
  if (r1 > rows) r1 = rows;
  // RB rows per thread per iteration, all loads issued before the first use (memory-level parallelism)
  for (long long r = r0 + rv.row_lane; r < r1; r += (long long)rv.rows_par * RB) {
    uint4 qz[RB], qr[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const long long rr = r + (long long)u * rv.rows_par;
      if (rr < r1) {
        const size_t off = (size_t)rr * C + rv.cv * 8;
        qz[u] = ldg16(z + off);
        if (RES_MODE != 0) qr[u] = ldg16(res + off);
      }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const long long rr = r + (long long)u * rv.rows_par;
      if (rr < r1) {
        const size_t off = (size_t)rr * C + rv.cv * 8;
        float v[8], o[8];
        unpack8(qz[u], v);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = v[t] * sc[t] + sh[t];
        if (RES_MODE != 0) {
          float w[8];
          unpack8(qr[u], w);
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] += (RES_MODE == 2) ? (w[t] * sc2[t] + sh2[t]) : w[t];
        }
        if (relu) {
#pragma unroll
          for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
        }
        uint4 q;
        q.x = pack_bf16x2(o[0], o[1]);
        q.y = pack_bf16x2(o[2], o[3]);
        q.z = pack_bf16x2(o[4], o[5]);
        q.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(out + off) = q;
        if (bits != nullptr) {  // bit t = (bf16 output of channel cv*8+t) > 0: the ReLU mask the backward pass needs
          auto pos2 = [](uint32_t w) -> uint32_t {  // two bf16 lanes -> two bits (value > 0)
            return (((w & 0x7fffu) != 0u && (w & 0x8000u) == 0u) ? 1u : 0u) |
                   (((w & 0x7fff0000u) != 0u && (w & 0x80000000u) == 0u) ? 2u : 0u);
          };
          const uint32_t b = pos2(q.x) | (pos2(q.y) << 2) | (pos2(q.z) << 4) | (pos2(q.w) << 6);
          bits[(size_t)rr * (C >> 3) + rv.cv] = (uint8_t)b;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward.  g = (g1 [+ g2]) * (act > 0) where `act` is the post-ReLU tensor of the forward pass.
//   reduce  : per-block partials of  sum(g), sum(g * xhat)            xhat = (z - mean) * rstd
//   finalize: dgamma = sum(g*xhat), dbeta = sum(g); c1 = dbeta/n, c2 = dgamma/n
//   apply   : dz = gamma*rstd * (g - c1 - xhat*c2)
// When the caller wants the masked g (the identity-shortcut gradient of a Bottleneck), the REDUCE pass stores it and the
// apply pass reads that one tensor back instead of g1, g2 and act again: 4R+1W + 2R+1W tensor passes instead of
// 4R + 4R+2W on the widest activations of the network.
// ---------------------------------------------------------------------------------------------------------------
// MASK: 0 = no ReLU on this BN output, 1 = mask from the saved post-ReLU tensor `act`, 2 = mask recomputed from z,
// 3 = mask from the bit plane bn_apply wrote (one byte per 8 channels; `qa.x` carries the byte).
template <bool HAS_G2, int MASK>
__device__ __forceinline__ void bn_bwd_masked_grad(const uint4& qg, const uint4& qh, const uint4& qa, const float (&zz)[8],
                                                   const float (&m)[8], const float (&rs)[8], const float (&ga)[8],
                                                   const float (&be)[8], float (&g)[8]) {
  unpack8(qg, g);
  if (HAS_G2) {
    float h[8];
    unpack8(qh, h);
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] += h[t];
  }
  if (MASK == 1) {
    float a[8];
    unpack8(qa, a);
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] = a[t] > 0.f ? g[t] : 0.f;
  } else if (MASK == 2) {
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] = (ga[t] * ((zz[t] - m[t]) * rs[t]) + be[t]) > 0.f ? g[t] : 0.f;
  } else if (MASK == 3) {
#pragma unroll
    for (int t = 0; t < 8; ++t) g[t] = ((qa.x >> t) & 1u) ? g[t] : 0.f;
  }
}

template <bool HAS_G2, int MASK, bool G_OUT>
__global__ void __launch_bounds__(kRvThreads, 2)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ g1, const __nv_bfloat16* __restrict__ g2,
                     const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ z,
                     const float* __restrict__ mean, const float* __restrict__ rstd, FSet gamma_s, FSet beta_s,
                     float* __restrict__ partial, __nv_bfloat16* __restrict__ g_out, long long rows, int C,
                     int rows_per_blk, int gps, int reverse) {
  constexpr int RB = kRowBatch;  // measured: 8 rows for the two-tensor instances is slower (register pressure)
  // reverse: start with the rows the producer of g (a dgrad GEMM, ascending) wrote last: they are still in L2; the
  // apply pass then runs ascending and starts with what this pass touched last.  Partials stay indexed by row range.
  const unsigned bx = reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const unsigned by = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const float* __restrict__ gamma = gamma_s.p[by / gps];
  const float* __restrict__ beta = beta_s.p[by / gps];
  extern __shared__ float red_smem[];
  RowVec rv(C);
  {
    const size_t goff = (size_t)by * rows * C;
    g1 += goff;
    if (HAS_G2) g2 += goff;
    if (MASK == 1) act += goff;
    if (MASK == 3) act = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const uint8_t*>(act) + (goff >> 3));
    z += goff;
    if (G_OUT) g_out += goff;
    mean += (size_t)by * C;
    rstd += (size_t)by * C;
    partial += (size_t)by * gridDim.x * 2 * C;
  }
  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = acc[1][t] = 0.f;
  if (rv.active) {
    float m[8], rs[8], ga[8], be[8];
    loadf8(mean + rv.cv * 8, m);
    loadf8(rstd + rv.cv * 8, rs);
    if (MASK == 2) {  // ReLU mask recomputed from z: relu(gamma*xhat + beta) > 0  (no residual on this BN)
      loadf8(gamma + rv.cv * 8, ga);
      loadf8(beta + rv.cv * 8, be);
    }
    const long long r0 = (long long)bx * rows_per_blk;
    long long r1 = r0 + rows_per_blk;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + rv.row_lane; r < r1; r += (long long)rv.rows_par * RB) {
      uint4 qg[RB], qh[RB], qa[RB], qz[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          const size_t off = (size_t)rr * C + rv.cv * 8;
          qg[u] = ldg16(g1 + off);
          if (HAS_G2) qh[u] = ldg16(g2 + off);
          qz[u] = ldg16(z + off);
          if (MASK == 1) qa[u] = ldg16(act + off);
          if (MASK == 3) qa[u].x = __ldg(reinterpret_cast<const uint8_t*>(act) + (off >> 3));
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          float g[8], zz[8];
          unpack8(qz[u], zz);
          bn_bwd_masked_grad<HAS_G2, MASK>(qg[u], qh[u], qa[u], zz, m, rs, ga, be, g);
          if (G_OUT) store8(g_out + (size_t)rr * C + rv.cv * 8, g);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            acc[0][t] += g[t];
            acc[1][t] += g[t] * ((zz[t] - m[t]) * rs[t]);
          }
        }
      }
    }
  }
  block_reduce_rows<2>(rv, C, acc, red_smem, partial + (size_t)bx * 2 * C, (size_t)C);
}

__global__ void __launch_bounds__(kFinThreads, 2)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, double count, FSetM dgamma_s, FSetM dbeta_s,
                       float beta_acc, float* __restrict__ c1, float* __restrict__ c2, int batch_stats, int groups) {
  __shared__ double s0s[kFinWarps][kFinCh];
  __shared__ double s1s[kFinWarps][kFinCh];
  const int cl = threadIdx.x & (kFinCh - 1);
  const int lane = threadIdx.x / kFinCh;
  const int c = blockIdx.x * kFinCh + cl;
  // parameter set = blockIdx.y (its `groups` statistic groups are consecutive)
  partial += (size_t)blockIdx.y * groups * nblk * 2 * C;
  c1 += (size_t)blockIdx.y * groups * C;
  c2 += (size_t)blockIdx.y * groups * C;
  float* __restrict__ dgamma = dgamma_s.p[blockIdx.y];
  float* __restrict__ dbeta = dbeta_s.p[blockIdx.y];
  double tot0 = 0.0, tot1 = 0.0;
  for (int g = 0; g < groups; ++g) {
    const float* pg = partial + (size_t)g * nblk * 2 * C + c;
    double s0, s1;
    fin_partial_sums(pg, pg + C, (size_t)2 * C, (size_t)2 * C, nblk, lane, s0, s1);
    fin_block_sums(s0, s1, s0s, s1s);
    if (threadIdx.x < kFinCh) {
      // each group normalised with its own statistics; eval mode (constants) has no batch-statistic terms in dz
      c1[(size_t)g * C + c] = batch_stats ? (float)(s0 / count) : 0.f;
      c2[(size_t)g * C + c] = batch_stats ? (float)(s1 / count) : 0.f;
      tot0 += s0;
      tot1 += s1;
    }
  }
  if (threadIdx.x < kFinCh) {
    dbeta[c] = (beta_acc != 0.f ? beta_acc * dbeta[c] : 0.f) + (float)tot0;
    dgamma[c] = (beta_acc != 0.f ? beta_acc * dgamma[c] : 0.f) + (float)tot1;
  }
}

template <bool HAS_G2, int MASK, bool G_OUT>
__global__ void __launch_bounds__(kRvThreads, 2)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ g1, const __nv_bfloat16* __restrict__ g2,
                    const __nv_bfloat16* __restrict__ act, const __nv_bfloat16* __restrict__ z,
                    const float* __restrict__ mean, const float* __restrict__ rstd, FSet gamma_s, FSet beta_s,
                    const float* __restrict__ c1, const float* __restrict__ c2, __nv_bfloat16* __restrict__ dz,
                    __nv_bfloat16* __restrict__ g_out, long long rows, int C, int rows_per_blk, int gps) {
  constexpr int RB = kRowBatch;
  RowVec rv(C);
  if (!rv.active) return;
  const float* __restrict__ gamma = gamma_s.p[blockIdx.y / gps];
  const float* __restrict__ beta = beta_s.p[blockIdx.y / gps];
  {
    const size_t goff = (size_t)blockIdx.y * rows * C;
    g1 += goff;
    if (HAS_G2) g2 += goff;
    if (MASK == 1) act += goff;
    if (MASK == 3) act = reinterpret_cast<const __nv_bfloat16*>(reinterpret_cast<const uint8_t*>(act) + (goff >> 3));
    z += goff;
    dz += goff;
    if (G_OUT) g_out += goff;
    mean += (size_t)blockIdx.y * C;
    rstd += (size_t)blockIdx.y * C;
    c1 += (size_t)blockIdx.y * C;
    c2 += (size_t)blockIdx.y * C;
  }
  float m[8], rs[8], k0[8], k1[8], k2[8], ga[8], be[8];
  if (MASK == 2) loadf8(beta + rv.cv * 8, be);
  loadf8(mean + rv.cv * 8, m);
  loadf8(rstd + rv.cv * 8, rs);
  loadf8(gamma + rv.cv * 8, ga);
  loadf8(c1 + rv.cv * 8, k1);
  loadf8(c2 + rv.cv * 8, k2);
#pragma unroll
  for (int t = 0; t < 8; ++t) k0[t] = ga[t] * rs[t];
  const long long r0 = (long long)blockIdx.x * rows_per_blk;
  long long r1 = r0 + rows_per_blk;
  if (r1 > rows) r1 = rows;
  for (long long r = r0 + rv.row_lane; r < r1; r += (long long)rv.rows_par * RB) {
    uint4 qg[RB], qh[RB], qa[RB], qz[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const long long rr = r + (long long)u * rv.rows_par;
      if (rr < r1) {
        const size_t off = (size_t)rr * C + rv.cv * 8;
        qg[u] = ldg16(g1 + off);
        if (HAS_G2) qh[u] = ldg16(g2 + off);
        qz[u] = ldg16(z + off);
        if (MASK == 1) qa[u] = ldg16(act + off);
        if (MASK == 3) qa[u].x = __ldg(reinterpret_cast<const uint8_t*>(act) + (off >> 3));
      }
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const long long rr = r + (long long)u * rv.rows_par;
      if (rr < r1) {
        const size_t off = (size_t)rr * C + rv.cv * 8;
        float g[8], zz[8], o[8];
        unpack8(qz[u], zz);
        bn_bwd_masked_grad<HAS_G2, MASK>(qg[u], qh[u], qa[u], zz, m, rs, ga, be, g);
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = k0[t] * (g[t] - k1[t] - (zz[t] - m[t]) * rs[t] * k2[t]);
        store8(dz + off, o);
        if (G_OUT) store8(g_out + off, g);
      }
    }
  }
}

// IRFD_ZIGZAG (default 1): reversed traversal of bn_apply / bn_bwd_reduce (L2 reuse across consecutive kernels)
static int zigzag() {
  static int v = [] {
    const char* e = getenv("IRFD_ZIGZAG");
    return e ? atoi(e) : 1;
  }();
  return v;
}

// Runtime (g2, mask, g_out) -> template instance: reduce -> finalize -> apply on one stream.
struct BnBwdLaunch {
  dim3 grid;
  size_t smem;
  cudaStream_t stream;
  const __nv_bfloat16 *g1, *g2, *act, *z;
  const float *mean, *rstd;
  FSet gamma, beta;
  float *partial, *c1, *c2;
  __nv_bfloat16 *dz, *g_out;
  FSetM dgamma, dbeta;
  float grad_beta;
  int batch_stats;
  long long grows;
  int c, rpb, nblk, groups;  // groups = TOTAL statistic groups (grid.y)
  int nsets, gps;            // parameter sets, groups per set
};

template <bool HAS_G2, int MASK>
static void launch_bn_bwd(const BnBwdLaunch& L) {
  // co-residency with the side-stream weight-gradient GEMMs (see host_util.h)
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_reduce_kernel<HAS_G2, MASK, true>));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_reduce_kernel<HAS_G2, MASK, false>));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_apply_kernel<HAS_G2, MASK, false>));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_apply_kernel<false, 0, false>));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_finalize_kernel));
  if (L.g_out != nullptr)
    bn_bwd_reduce_kernel<HAS_G2, MASK, true><<<L.grid, kRvThreads, L.smem, L.stream>>>(
        L.g1, L.g2, L.act, L.z, L.mean, L.rstd, L.gamma, L.beta, L.partial, L.g_out, L.grows, L.c, L.rpb, L.gps,
        zigzag());
  else
    bn_bwd_reduce_kernel<HAS_G2, MASK, false><<<L.grid, kRvThreads, L.smem, L.stream>>>(
        L.g1, L.g2, L.act, L.z, L.mean, L.rstd, L.gamma, L.beta, L.partial, nullptr, L.grows, L.c, L.rpb, L.gps,
        zigzag());
  bn_bwd_finalize_kernel<<<dim3(L.c / kFinCh, L.nsets), kFinThreads, 0, L.stream>>>(L.partial, L.nblk, L.c, (double)L.grows,
                                                                             L.dgamma, L.dbeta, L.grad_beta, L.c1, L.c2,
                                                                             L.batch_stats, L.gps);
  if (L.g_out != nullptr)  // the masked gradient is already in g_out (bf16, the value its other consumers see)
    bn_bwd_apply_kernel<false, 0, false><<<L.grid, kRvThreads, 0, L.stream>>>(
        L.g_out, nullptr, nullptr, L.z, L.mean, L.rstd, L.gamma, L.beta, L.c1, L.c2, L.dz, nullptr, L.grows, L.c, L.rpb,
        L.gps);
  else
    bn_bwd_apply_kernel<HAS_G2, MASK, false><<<L.grid, kRvThreads, 0, L.stream>>>(
        L.g1, L.g2, L.act, L.z, L.mean, L.rstd, L.gamma, L.beta, L.c1, L.c2, L.dz, nullptr, L.grows, L.c, L.rpb,
        L.gps);
}

}  // namespace irfd

using namespace irfd;

static FSet make_fset(const float* const* ptrs, int nsets) {
  FSet f;
  for (int i = 0; i < kMaxSets; ++i) f.p[i] = (ptrs != nullptr && i < nsets) ? ptrs[i] : nullptr;
  return f;
}
static FSetM make_fsetm(float* const* ptrs, int nsets) {
  FSetM f;
  for (int i = 0; i < kMaxSets; ++i) f.p[i] = (ptrs != nullptr && i < nsets) ? ptrs[i] : nullptr;
  return f;
}

// tiles / count are PER GROUP; mean/rstd are [nsets * groups][c] (`groups` statistic groups per parameter set).  Each
// set's running buffers get `running_updates` rounds of momentum updates: round 1 walks the set's groups in call order,
// round 2 (the reference's checkpoint recompute) in reverse order.  running_mean / running_var: host arrays of `nsets`
// device pointers (or NULL).
extern "C" int irfd_bn_finalize_sets(const float* psum, const float* psq, int tiles, int c, long long count, float eps,
                                     float momentum, float* mean, float* rstd, float* const* running_mean,
                                     float* const* running_var, int running_updates, int groups, int nsets,
                                     cudaStream_t stream) {
  IRFD_CHECK_ARG(psum && psq && mean && rstd && tiles > 0 && c > 0 && count > 0, "bn_finalize: bad argument");
  IRFD_CHECK_ARG(groups >= 1 && groups <= kMaxGroups, "bn_finalize: 1..4 statistic groups per set");
  IRFD_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets, "bn_finalize: 1..4 parameter sets");
  IRFD_CHECK_ARG(c % kFinCh == 0, "bn_finalize: C must be a multiple of 8 (got %d)", c);
  IRFD_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running buffers come in pairs");
  bn_finalize_kernel<<<dim3(c / kFinCh, nsets), kFinThreads, 0, stream>>>(psum, psq, tiles, c, (double)count, eps, momentum,
                                                                     mean, rstd, make_fsetm(running_mean, nsets),
                                                                     make_fsetm(running_var, nsets), running_updates,
                                                                     groups);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_finalize(const float* psum, const float* psq, int tiles, int c, long long count, float eps,
                                float momentum, float* mean, float* rstd, float* running_mean, float* running_var,
                                int running_updates, int groups, cudaStream_t stream) {
  float* rm[1] = {running_mean};
  float* rv[1] = {running_var};
  return irfd_bn_finalize_sets(psum, psq, tiles, c, count, eps, momentum, mean, rstd, running_mean ? rm : nullptr,
                               running_var ? rv : nullptr, running_updates, groups, 1, stream);
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

extern "C" int irfd_bn_eval_affine_sets(const float* const* running_mean, const float* const* running_var,
                                        const float* const* gamma, const float* const* beta, float eps, float* scale,
                                        float* shift, int c, int nsets, cudaStream_t stream) {
  IRFD_CHECK_ARG(running_mean && running_var && gamma && beta && scale && shift && c > 0, "bn_eval_affine_sets: bad argument");
  IRFD_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets, "bn_eval_affine_sets: 1..4 parameter sets");
  bn_eval_affine_sets_kernel<<<dim3((c + 255) / 256, nsets), 256, 0, stream>>>(
      make_fset(running_mean, nsets), make_fset(running_var, nsets), make_fset(gamma, nsets), make_fset(beta, nsets), eps,
      scale, shift, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_eval_rstd(const float* running_var, float eps, float* rstd, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(running_var && rstd && c > 0, "bn_eval_rstd: bad argument");
  bn_eval_rstd_kernel<<<(c + 255) / 256, 256, 0, stream>>>(running_var, eps, rstd, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

// rows = TOTAL rows (all groups, groups stacked along the row axis); mean/rstd (and mean2/rstd2) are [groups][c];
// gamma/beta (and gamma2/beta2): host arrays of `nsets` device pointers, set = group / (groups / nsets).
extern "C" int irfd_bn_apply_sets(const void* z, const float* mean, const float* rstd, const float* const* gamma,
                                  const float* const* beta, const void* res, const float* mean2, const float* rstd2,
                                  const float* const* gamma2, const float* const* beta2, void* out, void* mask_bits,
                                  long long rows, int c, int relu, int groups, int nsets, cudaStream_t stream) {
  IRFD_CHECK_ARG(z && mean && rstd && gamma && beta && out, "bn_apply: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048 && rows > 0, "bn_apply: C must be a multiple of 8 and <= 2048");
  IRFD_CHECK_ARG(groups >= 1 && rows % groups == 0, "bn_apply: rows must split evenly into groups");
  IRFD_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets && groups % nsets == 0, "bn_apply: groups must split evenly into sets");
  const long long grows = rows / groups;
  int nblk, rpb;
  plan_row_blocks(grows, c, num_sms(), &nblk, &rpb);
  BnAffine p1{mean, rstd, make_fset(gamma, nsets), make_fset(beta, nsets)};
  BnAffine p2{mean2, rstd2, make_fset(gamma2, nsets), make_fset(beta2, nsets)};
  auto zz = reinterpret_cast<const __nv_bfloat16*>(z);
  auto rr = reinterpret_cast<const __nv_bfloat16*>(res);
  auto oo = reinterpret_cast<__nv_bfloat16*>(out);
  auto mb = reinterpret_cast<uint8_t*>(mask_bits);  // optional [rows][c/8] ReLU mask plane for irfd_bn_backward_sets
  IRFD_CHECK_ARG(mb == nullptr || relu, "bn_apply: a mask plane needs relu");
  const dim3 grid(nblk, groups);
  const int gps = groups / nsets;
  if (res == nullptr)
    bn_apply_kernel<0><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, mb, grows, c, rpb, relu, gps, zigzag());
  else if (mean2 == nullptr)
    bn_apply_kernel<1><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, mb, grows, c, rpb, relu, gps, zigzag());
  else
    bn_apply_kernel<2><<<grid, kRvThreads, 0, stream>>>(zz, p1, rr, p2, oo, mb, grows, c, rpb, relu, gps, zigzag());
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_apply(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                             const void* res, const float* mean2, const float* rstd2, const float* gamma2,
                             const float* beta2, void* out, long long rows, int c, int relu, int groups,
                             cudaStream_t stream) {
  const float* g1[1] = {gamma};
  const float* b1[1] = {beta};
  const float* g2[1] = {gamma2};
  const float* b2[1] = {beta2};
  return irfd_bn_apply_sets(z, mean, rstd, g1, b1, res, mean2, rstd2, g2, b2, out, nullptr, rows, c, relu, groups, 1,
                            stream);
}

extern "C" long long irfd_bn_bwd_workspace_bytes(long long rows, int c, int groups) {
  if (groups < 1) groups = 1;
  int nblk, rpb;
  plan_row_blocks(rows / groups, c, num_sms(), &nblk, &rpb);
  return (long long)(nblk * 2 + 2) * groups * c * 4;
}

// Full BN backward (reduce -> finalize -> apply).  workspace: [groups][nblk][2][C] partials, c1[groups][C], c2[...].
// groups = TOTAL statistic groups; gamma/beta/dgamma/dbeta: host arrays of `nsets` device pointers (beta may be NULL =
// no ReLU-from-z mask); each set's dgamma/dbeta sums its groups / nsets groups.
// act_is_bits: `act` is the [rows][c/8] mask plane irfd_bn_apply_sets wrote, not the post-ReLU tensor.
extern "C" int irfd_bn_backward_sets(const void* g1, const void* g2, const void* act, int act_is_bits, const void* z,
                                     const float* mean, const float* rstd, const float* const* gamma,
                                     const float* const* beta, void* dz, void* g_out, float* const* dgamma,
                                     float* const* dbeta, float grad_beta, int batch_stats, long long rows, int c,
                                     int groups, int nsets, void* workspace, long long workspace_bytes,
                                     cudaStream_t stream) {
  IRFD_CHECK_ARG(g1 && z && mean && rstd && gamma && dz && dgamma && dbeta && workspace, "bn_backward: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048 && rows > 0, "bn_backward: C must be a multiple of 8 and <= 2048");
  IRFD_CHECK_ARG(groups >= 1 && rows % groups == 0, "bn_backward: rows must split evenly into groups");
  IRFD_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets && groups % nsets == 0,
                 "bn_backward: groups must split evenly into 1..4 sets");
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
  BnBwdLaunch L;
  L.grid = dim3(nblk, groups);
  L.smem = smem;
  L.stream = stream;
  L.g1 = G1; L.g2 = G2; L.act = A; L.z = Z;
  L.mean = mean; L.rstd = rstd;
  L.gamma = make_fset(gamma, nsets);
  L.beta = make_fset(beta, nsets);
  L.partial = partial; L.c1 = c1; L.c2 = c2;
  L.dz = reinterpret_cast<__nv_bfloat16*>(dz);
  L.g_out = reinterpret_cast<__nv_bfloat16*>(g_out);
  L.dgamma = make_fsetm(dgamma, nsets);
  L.dbeta = make_fsetm(dbeta, nsets);
  L.grad_beta = grad_beta; L.batch_stats = batch_stats;
  L.grows = grows; L.c = c; L.rpb = rpb; L.nblk = nblk; L.groups = groups;
  L.nsets = nsets; L.gps = groups / nsets;
  const bool has_beta = beta != nullptr && beta[0] != nullptr;
  const int mask = A != nullptr ? (act_is_bits ? 3 : 1) : (has_beta ? 2 : 0);
  if (G2 != nullptr) {
    if (mask == 0) launch_bn_bwd<true, 0>(L);
    else if (mask == 1) launch_bn_bwd<true, 1>(L);
    else if (mask == 2) launch_bn_bwd<true, 2>(L);
    else launch_bn_bwd<true, 3>(L);
  } else {
    if (mask == 0) launch_bn_bwd<false, 0>(L);
    else if (mask == 1) launch_bn_bwd<false, 1>(L);
    else if (mask == 2) launch_bn_bwd<false, 2>(L);
    else launch_bn_bwd<false, 3>(L);
  }
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

// Second half of a BatchNorm backward whose reduce pass ran inside the producing dgrad GEMM's epilogue
// (irfd_conv_gemm_bnbwd_grouped): g is the MASKED activation gradient, partial [groups][tiles][2][c] the per-tile sums
// of g and g*xhat.  finalize (dgamma, dbeta, c1, c2) + apply (dz from g and z).  workspace: 2 * groups * c floats.
extern "C" int irfd_bn_backward_finish_sets(const void* g, const void* z, const float* mean, const float* rstd,
                                            const float* const* gamma, void* dz, float* const* dgamma,
                                            float* const* dbeta, float grad_beta, int batch_stats, long long rows,
                                            int c, int groups, int nsets, const float* partial, int tiles,
                                            void* workspace, long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(g && z && mean && rstd && gamma && dz && dgamma && dbeta && partial && workspace,
                 "bn_backward_finish: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048 && rows > 0 && tiles > 0, "bn_backward_finish: bad shape");
  IRFD_CHECK_ARG(groups >= 1 && rows % groups == 0, "bn_backward_finish: rows must split evenly into groups");
  IRFD_CHECK_ARG(nsets >= 1 && nsets <= kMaxSets && groups % nsets == 0,
                 "bn_backward_finish: groups must split evenly into 1..4 sets");
  IRFD_CHECK_ARG(workspace_bytes >= (long long)2 * groups * c * 4, "bn_backward_finish: workspace too small");
  const long long grows = rows / groups;
  int nblk, rpb;
  plan_row_blocks(grows, c, num_sms(), &nblk, &rpb);
  float* c1 = reinterpret_cast<float*>(workspace);
  float* c2 = c1 + (size_t)groups * c;
  const FSet ga = make_fset(gamma, nsets);
  const FSet be = make_fset(nullptr, nsets);
  const int gps = groups / nsets;
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_apply_kernel<false, 0, false>));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&bn_bwd_finalize_kernel));
  bn_bwd_finalize_kernel<<<dim3(c / kFinCh, nsets), kFinThreads, 0, stream>>>(partial, tiles, c, (double)grows,
                                                                        make_fsetm(dgamma, nsets),
                                                                        make_fsetm(dbeta, nsets), grad_beta, c1, c2,
                                                                        batch_stats, gps);
  bn_bwd_apply_kernel<false, 0, false><<<dim3(nblk, groups), kRvThreads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(g), nullptr, nullptr, reinterpret_cast<const __nv_bfloat16*>(z), mean, rstd,
      ga, be, c1, c2, reinterpret_cast<__nv_bfloat16*>(dz), nullptr, grows, c, rpb, gps);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_bn_backward(const void* g1, const void* g2, const void* act, const void* z, const float* mean,
                                const float* rstd, const float* gamma, const float* beta, void* dz, void* g_out,
                                float* dgamma, float* dbeta, float grad_beta, int batch_stats, long long rows, int c,
                                int groups, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  const float* ga[1] = {gamma};
  const float* be[1] = {beta};
  float* dg[1] = {dgamma};
  float* db[1] = {dbeta};
  return irfd_bn_backward_sets(g1, g2, act, 0, z, mean, rstd, ga, be, dz, g_out, dg, db, grad_beta, batch_stats, rows,
                               c, groups, 1, workspace, workspace_bytes, stream);
}
