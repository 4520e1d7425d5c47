// Memory-bound pieces of the spectral-norm discriminator (styleganv1.py:637-695) that are not conv epilogues:
//   from_rgb      : 1x1 conv 3 -> C on the NCHW fp32 image + bias + leaky_relu(0.2) -> NHWC bf16 (styleganv1.py:643,662)
//   bias_lrelu_bwd: backward of `y = leaky_relu(conv + bias)`:  dz = g * (y > 0 ? 1 : 0.2),  dbias = sum dz
// The 3x3 convs themselves run on the tcgen05 GEMM kernels (irfd_conv_gemm_affine with the leaky epilogue).
#include "host_util.h"
#include "rowvec.cuh"

namespace irfd {

// one thread per (pixel, 8-channel vector); the three image planes of a pixel are read once per thread (L1 serves the
// C/8 threads of a pixel), the write is one coalesced 16-byte store.
__global__ void __launch_bounds__(256)
from_rgb_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out, int B, int HW, int C, int lrelu) {
  extern __shared__ float sw[];  // [C][3] weights + [C] bias
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[3 * C + i] = bias != nullptr ? bias[i] : 0.f;
  __syncthreads();
  const unsigned vc = C >> 3;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;  // host checks B*HW*vc < 2^32
  if (idx >= (unsigned)B * HW * vc) return;
  const unsigned pix = idx / vc, v = idx - pix * vc;
  const unsigned b = pix / HW, p = pix - b * HW;
  const float* xp = x + (size_t)b * 3 * HW + p;
  const float r = __ldg(xp), g = __ldg(xp + HW), bl = __ldg(xp + 2 * (size_t)HW);
  float o[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int c = v * 8 + t;
    const float z = sw[c * 3] * r + sw[c * 3 + 1] * g + sw[c * 3 + 2] * bl + sw[3 * C + c];
    o[t] = (lrelu == 0 || z > 0.f) ? z : 0.2f * z;
  }
  store8(out + (size_t)pix * C + v * 8, o);
}

__global__ void __launch_bounds__(kRvThreads)
bias_lrelu_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                      __nv_bfloat16* __restrict__ dz, float* __restrict__ partial, long long rows, int C,
                      int rows_per_blk) {
  constexpr int RB = 8;
  extern __shared__ float red_smem[];
  RowVec rv(C);
  float acc[1][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = 0.f;
  if (rv.active) {
    const long long r0 = (long long)blockIdx.x * rows_per_blk;
    long long r1 = r0 + rows_per_blk;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + rv.row_lane; r < r1; r += (long long)rv.rows_par * RB) {
      uint4 qg[RB], qy[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          const size_t off = (size_t)rr * C + rv.cv * 8;
          qg[u] = ldg16(g + off);
          qy[u] = ldg16(y + off);
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          float gv[8], yv[8], o[8];
          unpack8(qg[u], gv);
          unpack8(qy[u], yv);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            o[t] = gv[t] * (yv[t] > 0.f ? 1.f : 0.2f);
            acc[0][t] += o[t];
          }
          store8(dz + (size_t)rr * C + rv.cv * 8, o);
        }
      }
    }
  }
  block_reduce_rows<1>(rv, C, acc, red_smem, partial + (size_t)blockIdx.x * C, 0);
}

__global__ void bias_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = 0;
  for (; b + 3 < nblk; b += 4) {
    s0 += partial[(size_t)b * C + c];
    s1 += partial[(size_t)(b + 1) * C + c];
    s2 += partial[(size_t)(b + 2) * C + c];
    s3 += partial[(size_t)(b + 3) * C + c];
  }
  for (; b < nblk; ++b) s0 += partial[(size_t)b * C + c];
  dbias[c] = (s0 + s1) + (s2 + s3);
}

}  // namespace irfd

using namespace irfd;

extern "C" int irfd_from_rgb_fwd(const float* x, const float* w, const float* bias, void* out, int b, int hw, int c,
                                 int lrelu, cudaStream_t stream) {
  IRFD_CHECK_ARG(x && w && out && b > 0 && hw > 0 && c % 8 == 0 && c <= 2048, "from_rgb_fwd: bad argument");
  const size_t total = (size_t)b * hw * (c / 8);
  IRFD_CHECK_ARG(total < ((size_t)1 << 32) - 256, "from_rgb_fwd: tensor too large for 32-bit indexing");
  from_rgb_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 4 * c * sizeof(float), stream>>>(
      x, w, bias, reinterpret_cast<__nv_bfloat16*>(out), b, hw, c, lrelu);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" long long irfd_bias_lrelu_bwd_workspace_bytes(long long rows, int c) {
  int nblk, rpb;
  plan_row_blocks(rows, c, num_sms(), &nblk, &rpb);
  return (long long)nblk * c * 4;
}

extern "C" int irfd_bias_lrelu_bwd(const void* g, const void* y, void* dz, float* dbias, long long rows, int c,
                                   void* workspace, long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(g && y && dz && dbias && workspace && rows > 0 && c % 8 == 0 && c <= 2048,
                 "bias_lrelu_bwd: bad argument");
  int nblk, rpb;
  plan_row_blocks(rows, c, num_sms(), &nblk, &rpb);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)nblk * c * 4, "bias_lrelu_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  bias_lrelu_bwd_kernel<<<nblk, kRvThreads, 2048 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(g), reinterpret_cast<const __nv_bfloat16*>(y),
      reinterpret_cast<__nv_bfloat16*>(dz), partial, rows, c, rpb);
  IRFD_CHECK_LAUNCH();
  bias_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, stream>>>(partial, nblk, c, dbias);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
