// Memory-bound pieces of the StyleGAN-v1-style synthesis network (styleganv1.py:569-635) that are not conv epilogues:
// constant-input stage, bilinear x2 upsample (+adjoint), backward of the fused noise/lrelu/style epilogue with its
// warp/block reductions, and the 1x1 to_rgb (+backward).  NHWC bf16 activations, fp32 parameters and reductions.
#include "host_util.h"
#include "rowvec.cuh"

namespace irfd {

// ---------------------------------------------------------------------------------------------------------------
// Constant input: a0 = const[c,h,w] + bias[c] + nw[c]*noise[b,h,w];  y0 = a0*sp1[b,c] + s1[b,c]
// (styleganv1.py:596-599 — no leaky_relu on this stage)
// ---------------------------------------------------------------------------------------------------------------
__global__ void const_input_fwd_kernel(const float* __restrict__ cst, const float* __restrict__ bias,
                                       const float* __restrict__ nw, const float* __restrict__ noise,
                                       const float* __restrict__ sp1, const float* __restrict__ s1,
                                       __nv_bfloat16* __restrict__ a0, __nv_bfloat16* __restrict__ y0,
                                       __nv_bfloat16* __restrict__ y0_lo, int B, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [B][16][C]
  if (idx >= (size_t)B * 16 * C) return;
  const int c = idx % C;
  const int hw = (idx / C) % 16;
  const int b = idx / ((size_t)C * 16);
  const float a = cst[c * 16 + hw] + bias[c] + nw[c] * noise[b * 16 + hw];
  a0[idx] = __float2bfloat16(a);
  const float y = a * sp1[(size_t)b * C + c] + s1[(size_t)b * C + c];
  const __nv_bfloat16 yh = __float2bfloat16(y);
  y0[idx] = yh;
  if (y0_lo != nullptr) y0_lo[idx] = __float2bfloat16(y - __bfloat162float(yh));  // split-bf16 source of the upsample
}

// block = 32 channels (lanes, coalesced) x 8 warps striding over the batch; per-image sums (dsp1, ds1) are final in
// their warp, per-channel ones (dconst[16], dbias, dnw) are combined across the warps through shared memory in a fixed
// order.  (The first version gave a whole channel to ONE thread: 1024 dependent iterations, 206 us on the critical path.)
constexpr int kCiWarps = 8;
__global__ void __launch_bounds__(32 * kCiWarps)
const_input_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ a0,
                       const float* __restrict__ noise, const float* __restrict__ sp1, float* __restrict__ dsp1,
                       float* __restrict__ ds1, float* __restrict__ dconst, float* __restrict__ dbias,
                       float* __restrict__ dnw, int B, int C) {
  __shared__ float red[kCiWarps][18][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float dc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) dc[i] = 0.f;
  float db = 0.f, dn = 0.f;
  if (c < C) {
    for (int b = warp; b < B; b += kCiWarps) {
      float t0 = 0.f, t1 = 0.f;
      const float sp = sp1[(size_t)b * C + c];
#pragma unroll
      for (int hw = 0; hw < 16; ++hw) {
        const size_t i = ((size_t)b * 16 + hw) * C + c;
        const float g = __bfloat162float(dy[i]);
        t0 += g * __bfloat162float(a0[i]);
        t1 += g;
        const float dz = g * sp;
        dc[hw] += dz;
        db += dz;
        dn += dz * noise[b * 16 + hw];
      }
      dsp1[(size_t)b * C + c] = t0;
      ds1[(size_t)b * C + c] = t1;
    }
  }
#pragma unroll
  for (int hw = 0; hw < 16; ++hw) red[warp][hw][lane] = dc[hw];
  red[warp][16][lane] = db;
  red[warp][17][lane] = dn;
  __syncthreads();
  if (c < C) {
    for (int j = warp; j < 18; j += kCiWarps) {  // each warp finishes a few of the 18 per-channel sums
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kCiWarps; ++w) s += red[w][j][lane];
      if (j < 16) dconst[c * 16 + j] = s;
      else if (j == 16) dbias[c] = s;
      else dnw[c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Bilinear x2, align_corners=False (nn.Upsample, styleganv1.py:621,624), ATen index arithmetic restated.
// ---------------------------------------------------------------------------------------------------------------
// For scale 2 the source position of output o is o/2 - 0.25 (clamped at 0), so every weight is exactly 0.25, 0.75 or
// (at the two borders) 1.0 and the neighbour indices follow from the parity of o: no float->int conversions and no
// integer divisions in the kernels (the first version spent its time on 64-bit div/mod index arithmetic, not on HBM).
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_src(int o, int in_size) {
  Lerp r;
  const int i = o >> 1;
  if (o & 1) {  // src = i + 0.25
    r.i0 = i;
    r.i1 = i + (i < in_size - 1 ? 1 : 0);
    r.l1 = 0.25f;
    r.l0 = 0.75f;
  } else if (o == 0) {  // src clamped to 0
    r.i0 = 0;
    r.i1 = in_size > 1 ? 1 : 0;
    r.l1 = 0.f;
    r.l0 = 1.f;
  } else {  // src = (i - 1) + 0.75
    r.i0 = i - 1;
    r.i1 = i;
    r.l1 = 0.75f;
    r.l0 = 0.25f;
  }
  return r;
}

// One thread per INPUT pixel vector: it loads the 3x3 input neighbourhood once (9 x 16 B) and writes the 2x2 output
// block (4 x 16 B) — 2.25 loads per store instead of 4, a quarter of the CTAs.  Per output the arithmetic is exactly
// ATen's:  ly.l0 * (lx.l0 * v00 + lx.l1 * v01) + ly.l1 * (lx.l0 * v10 + lx.l1 * v11).
// grid = (B*H input rows, segments of W*C/8 vectors): row decode is per block, the column decode is 32-bit.
// LO: the input is split bf16, value = in + in_lo (the small layers, see conv_gemm STYLE / irfd_conv_gemm_style_split).
template <bool LO>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ in_lo,
                      __nv_bfloat16* __restrict__ out, int B, int H, int W, int C) {
  const unsigned vc = C >> 3, Wo = 2 * W;
  const unsigned i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= (unsigned)W * vc) return;
  const int w = i / vc, v = i - w * vc;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int hs[3] = {h > 0 ? h - 1 : 0, h, h < H - 1 ? h + 1 : h};
  const int ws[3] = {w > 0 ? w - 1 : 0, w, w < W - 1 ? w + 1 : w};
  const __nv_bfloat16* base = in + (size_t)b * H * W * C + v * 8;
  uint4 q[3][3], ql[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      q[r][c] = ldg16(base + ((size_t)hs[r] * W + ws[c]) * C);
      if constexpr (LO) ql[r][c] = ldg16(in_lo + (size_t)b * H * W * C + v * 8 + ((size_t)hs[r] * W + ws[c]) * C);
    }
  auto fetch = [&](int r, int c, float (&f)[8]) {  // r, c are compile-time after unrolling
    unpack8(q[r][c], f);
    if constexpr (LO) {
      float l[8];
      unpack8(ql[r][c], l);
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] += l[t];
    }
  };
  // Output column 2w + dx interpolates neighbourhood columns (dx, dx + 1) = (left, centre) / (centre, right) and output
  // row 2h + dy rows (dy, dy + 1), with ATen's weights: (0.25, 0.75) for the even output, (0.75, 0.25) for the odd one,
  // (l0, l1) = (1, 0) on the clamped first output (whose i0 is the centre: the neighbourhood's clamped left column holds
  // the same pixel, so 0 * left + 1 * centre is ATen's 1 * centre + 0 * next).  Per-thread weights, fixed operands: no
  // per-value selects (ncu, round 2: the select-heavy version was issue-bound at 76 % SM throughput, 2.6 TB/s).
  const float ax[2] = {w > 0 ? 0.25f : 0.f, 0.75f}, bx[2] = {w > 0 ? 0.75f : 1.f, 0.25f};
  const float ay[2] = {h > 0 ? 0.25f : 0.f, 0.75f}, by[2] = {h > 0 ? 0.75f : 1.f, 0.25f};
  float fq[3][3][8];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) fetch(r, c, fq[r][c]);
  // horizontal pass: hl[r][dx] = lx.l0 * row[lx.i0] + lx.l1 * row[lx.i1] for the two output columns 2w, 2w + 1
  float hl[3][2][8];
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 8; ++t) hl[r][dx][t] = ax[dx] * fq[r][dx][t] + bx[dx] * fq[r][dx + 1][t];
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) o[t] = ay[dy] * hl[dy][dx][t] + by[dy] * hl[dy + 1][dx][t];
      store8(out + (((size_t)b * 2 * H + 2 * h + dy) * Wo + 2 * w + dx) * C + v * 8, o);
    }
  }
}

// Adjoint, gather form: input index i receives from outputs 2i-1 .. 2i+2 with weights (0.25, 0.75, 0.75, 0.25); at
// the borders the clamped neighbour folds onto i (0.75 -> 1.0) and the out-of-range output does not exist (-> 0).
__device__ __forceinline__ void lerp_adjoint_weights(int i, int in_size, float (&w)[4]) {
  w[0] = i > 0 ? 0.25f : 0.f;
  w[1] = i > 0 ? 0.75f : 1.f;
  w[2] = i < in_size - 1 ? 0.75f : 1.f;
  w[3] = i < in_size - 1 ? 0.25f : 0.f;
}

// One thread per 2x2 block of INPUT pixels (one 8-channel vector): the four adjoint stencils (4x4 output gradients each)
// overlap in a 6x6 window, so 36 loads serve 4 results — 9 per result instead of 16 (ncu, round 2: the one-pixel-per-
// thread version pulled every gradient 4x through L2, 42 % hit rate, 2.3 TB/s).  Row by row: six gradients are combined
// horizontally into the two input columns, then added into the two input rows with the vertical weights.
// grid = (B*ceil(H/2) input row pairs, segments of ceil(W/2)*C/8 vectors); odd sizes leave the last pair half empty.
__global__ void __launch_bounds__(256, 2)
upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ din, int B, int H, int W,
                      int C) {
  const unsigned vc = C >> 3, Wo = 2 * W, Ho = 2 * H, W2 = (W + 1) >> 1, H2 = (H + 1) >> 1;
  const unsigned i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= W2 * vc) return;
  const int wb = i / vc, v = i - wb * vc;
  const int b = blockIdx.x / H2, hb = blockIdx.x - b * H2;
  const int h0 = 2 * hb, w0 = 2 * wb;
  float wy[2][4], wx[2][4];
  lerp_adjoint_weights(h0, H, wy[0]);
  lerp_adjoint_weights(h0 + 1, H, wy[1]);
  lerp_adjoint_weights(w0, W, wx[0]);
  lerp_adjoint_weights(w0 + 1, W, wx[1]);
  const __nv_bfloat16* base = dout + (size_t)b * Ho * Wo * C + v * 8;
  float acc[2][2][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[a][c][t] = 0.f;
  int ows[6];
#pragma unroll
  for (int dx = 0; dx < 6; ++dx) {  // clamp the coordinates of the non-existent (zero-weight) outputs into range
    int ow = 2 * w0 - 1 + dx;
    ows[dx] = ow < 0 ? 0 : (ow > (int)Wo - 1 ? (int)Wo - 1 : ow);
  }
#pragma unroll
  for (int dy = 0; dy < 6; ++dy) {
    int oh = 2 * h0 - 1 + dy;
    oh = oh < 0 ? 0 : (oh > (int)Ho - 1 ? (int)Ho - 1 : oh);
    uint4 q[6];
#pragma unroll
    for (int dx = 0; dx < 6; ++dx) q[dx] = ldg16(base + ((size_t)oh * Wo + ows[dx]) * C);
    float cw[2][8];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int t = 0; t < 8; ++t) cw[c][t] = 0.f;
#pragma unroll
    for (int dx = 0; dx < 6; ++dx) {
      float g[8];
      unpack8(q[dx], g);
      // input column c (= w0 + c) receives output columns 2(w0+c)-1 .. 2(w0+c)+2 = window columns 2c .. 2c+3
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int k = dx - 2 * c;
        if (k >= 0 && k < 4) {
          const float ww = wx[c][k];
          if (ww != 0.f) {  // a clamped (non-existent) output must not contribute even when it holds inf/nan
#pragma unroll
            for (int t = 0; t < 8; ++t) cw[c][t] += ww * g[t];
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {  // input row a (= h0 + a) receives window rows 2a .. 2a+3
      const int k = dy - 2 * a;
      if (k >= 0 && k < 4) {
        const float ww = wy[a][k];
        if (ww != 0.f) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[a][c][t] += ww * cw[c][t];
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (h0 + a < H && w0 + c < W) store8(din + (((size_t)b * H + h0 + a) * W + w0 + c) * C + v * 8, acc[a][c]);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the fused conv epilogue  y = lrelu(z)*sp1 + s1,  z = conv + bias + nw*noise,  a = lrelu(z) saved.
//   dz = dy * sp1[b,c] * (a > 0 ? 1 : 0.2)
//   per (b,c): ds1 = sum_hw dy, dsp1 = sum_hw dy*a ;  per c: dbias = sum dz, dnw = sum dz*noise[b,hw]
// grid = (chunks, B); partial[b][chunk][4][C]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRvThreads, 2)
style_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ a,
                 const float* __restrict__ noise, const float* __restrict__ sp1, __nv_bfloat16* __restrict__ dz,
                 float* __restrict__ partial, int HW, int C, int rows_per_blk) {
  constexpr int RB = 4;  // two streamed tensors: 8 loads of 16 bytes in flight per thread, two blocks per SM
  extern __shared__ float red_smem[];
  RowVec rv(C);
  const int b = blockIdx.y;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[k][t] = 0.f;
  if (rv.active) {
    float sp[8];
    loadf8(sp1 + (size_t)b * C + rv.cv * 8, sp);
    const int r0 = blockIdx.x * rows_per_blk;
    int r1 = r0 + rows_per_blk;
    if (r1 > HW) r1 = HW;
    for (int r = r0 + rv.row_lane; r < r1; r += rv.rows_par * RB) {
      uint4 qg[RB], qa[RB];
      float nz[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int rr = r + u * rv.rows_par;
        if (rr < r1) {
          const size_t off = ((size_t)b * HW + rr) * C + rv.cv * 8;
          qg[u] = ldg16(dy + off);
          qa[u] = ldg16(a + off);
          nz[u] = __ldg(noise + (size_t)b * HW + rr);
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int rr = r + u * rv.rows_par;
        if (rr < r1) {
          const size_t off = ((size_t)b * HW + rr) * C + rv.cv * 8;
          float g[8], av[8], o[8];
          unpack8(qg[u], g);
          unpack8(qa[u], av);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float d = g[t] * sp[t] * (av[t] > 0.f ? 1.f : 0.2f);
            o[t] = d;
            acc[0][t] += g[t];
            acc[1][t] += g[t] * av[t];
            acc[2][t] += d;
            acc[3][t] += d * nz[u];
          }
          store8(dz + off, o);
        }
      }
    }
  }
  // two rounds of two sums: 16 KB of shared memory instead of 32, so a block fits beside a weight-gradient GEMM CTA
  float* dst = partial + ((size_t)b * gridDim.x + blockIdx.x) * 4 * C;
  block_reduce_rows<2>(rv, C, *reinterpret_cast<float(*)[2][8]>(&acc[0]), red_smem, dst, (size_t)C);
  __syncthreads();
  block_reduce_rows<2>(rv, C, *reinterpret_cast<float(*)[2][8]>(&acc[2]), red_smem, dst + 2 * (size_t)C, (size_t)C);
}

__global__ void __launch_bounds__(1024)
style_bwd_finalize_kernel(const float* __restrict__ partial, int B, int chunks, int C, float* __restrict__ ds1,
                          float* __restrict__ dsp1, float* __restrict__ dbias, float* __restrict__ dnw) {
  // block = 32 channels (lanes) x 32 warps striding over images; per-image sums are final, per-channel ones are
  // combined across warps through shared memory in a fixed order.
  __shared__ float sb[32][33];
  __shared__ float sn[32][33];
  const int cl = threadIdx.x & 31;
  const int wl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float tb = 0.f, tn = 0.f;
  if (c < C) {
    for (int b = wl; b < B; b += 32) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
      for (int k = 0; k < chunks; ++k) {
        const float* p = partial + ((size_t)b * chunks + k) * 4 * C + c;
        t0 += p[0];
        t1 += p[C];
        t2 += p[2 * C];
        t3 += p[3 * C];
      }
      ds1[(size_t)b * C + c] = t0;
      dsp1[(size_t)b * C + c] = t1;
      tb += t2;
      tn += t3;
    }
  }
  sb[wl][cl] = tb;
  sn[wl][cl] = tn;
  __syncthreads();
  if (wl == 0 && c < C) {
    for (int i = 1; i < 32; ++i) {
      tb += sb[i][cl];
      tn += sn[i][cl];
    }
    dbias[c] = tb;
    dnw[c] = tn;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// to_rgb: 1x1 conv C -> 3 with bias, output NCHW fp32 (the reference's output layout) — styleganv1.py:588,607
// ---------------------------------------------------------------------------------------------------------------
__global__ void to_rgb_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ out, int B, int HW, int C) {
  extern __shared__ float sw[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)B * HW) return;
  const int b = pix / HW, p = pix % HW;
  float r0 = bias[0], r1 = bias[1], r2 = bias[2];
  for (int v = 0; v < C / 8; ++v) {
    float f[8];
    load8(y + pix * C + v * 8, f);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      r0 += f[t] * sw[v * 8 + t];
      r1 += f[t] * sw[C + v * 8 + t];
      r2 += f[t] * sw[2 * C + v * 8 + t];
    }
  }
  float* o = out + (size_t)b * 3 * HW + p;
  o[0] = r0;
  o[HW] = r1;
  o[2 * HW] = r2;
}

// dy[pix][c] = sum_k drgb[k][pix]*w[k][c];  partial[blk][4][C]: k<3 -> sum_pix drgb[k]*y[c]; row 3 (ch 0..2) -> dbias
__global__ void __launch_bounds__(kRvThreads)
to_rgb_bwd_kernel(const float* __restrict__ drgb, const __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                  __nv_bfloat16* __restrict__ dy, float* __restrict__ partial, int B, int HW, int C,
                  int rows_per_blk) {
  constexpr int RB = 8;  // (measured: 4 rows in flight and two blocks per SM is no faster: 375 vs 360 us)
  extern __shared__ float red_smem[];
  RowVec rv(C);
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[k][t] = 0.f;
  if (rv.active) {
    float w0[8], w1[8], w2[8];
    loadf8(w + rv.cv * 8, w0);
    loadf8(w + C + rv.cv * 8, w1);
    loadf8(w + 2 * C + rv.cv * 8, w2);
    const long long rows = (long long)B * HW;
    const long long r0 = (long long)blockIdx.x * rows_per_blk;
    long long r1 = r0 + rows_per_blk;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + rv.row_lane; r < r1; r += (long long)rv.rows_par * RB) {
      uint4 qy[RB];
      float g0[RB], g1[RB], g2[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          // 32-bit decode (rows < 2^31, host-checked): the 64-bit division here made the kernel issue-bound
          const unsigned b = (unsigned)rr / (unsigned)HW, p = (unsigned)rr - b * (unsigned)HW;
          const float* g = drgb + (size_t)b * 3 * HW + p;
          g0[u] = __ldg(g);
          g1[u] = __ldg(g + HW);
          g2[u] = __ldg(g + 2 * (size_t)HW);
          qy[u] = ldg16(y + (size_t)rr * C + rv.cv * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const long long rr = r + (long long)u * rv.rows_par;
        if (rr < r1) {
          float f[8], o[8];
          unpack8(qy[u], f);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            o[t] = g0[u] * w0[t] + g1[u] * w1[t] + g2[u] * w2[t];
            acc[0][t] += g0[u] * f[t];
            acc[1][t] += g1[u] * f[t];
            acc[2][t] += g2[u] * f[t];
          }
          if (rv.cv == 0) {
            acc[3][0] += g0[u];
            acc[3][1] += g1[u];
            acc[3][2] += g2[u];
          }
          store8(dy + (size_t)rr * C + rv.cv * 8, o);
        }
      }
    }
  }
  block_reduce_rows<4>(rv, C, acc, red_smem, partial + (size_t)blockIdx.x * 4 * C, (size_t)C);
}

__global__ void to_rgb_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dw,
                                           float* __restrict__ dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 3*C (+3)
  if (i < 3 * C) {
    const int k = i / C, c = i % C;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[((size_t)b * 4 + k) * C + c];
    dw[i] = s;
  } else if (i < 3 * C + 3) {
    const int k = i - 3 * C;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[((size_t)b * 4 + 3) * C + k];
    dbias[k] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Standalone ApplyNoise / ApplyStyle on the reference's own layout (NCHW fp32, styleganv1.py:453-468): the fused
// path applies both inside the conv epilogue; these serve direct calls of the two modules.
//   noise: out = x + w[c] * noise[b, hw]          style: out = x * (style[b, c] + 1) + style[b, C + c]
// ---------------------------------------------------------------------------------------------------------------
__global__ void apply_noise_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                        const float* __restrict__ noise, float* __restrict__ out, int C, int HW,
                                        size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hw = (int)(i % HW);
  const size_t bc = i / HW;
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  out[i] = x[i] + w[c] * noise[b * HW + hw];
}

__global__ void apply_style_nchw_kernel(const float* __restrict__ x, const float* __restrict__ style,
                                        float* __restrict__ out, int C, int HW, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t bc = i / HW;
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  out[i] = x[i] * (style[b * 2 * C + c] + 1.f) + style[b * 2 * C + C + c];
}

}  // namespace irfd

using namespace irfd;

#define GRID1D(total) (unsigned)(((total) + 255) / 256), 256, 0, stream
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int irfd_const_input_split_fwd(const float* cst, const float* bias, const float* nw, const float* noise,
                                          const float* sp1, const float* s1, void* a0, void* y0, void* y0_lo, int b,
                                          int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(cst && bias && nw && noise && sp1 && s1 && a0 && y0, "const_input_fwd: null pointer");
  const_input_fwd_kernel<<<GRID1D((size_t)b * 16 * c)>>>(cst, bias, nw, noise, sp1, s1, BF(a0), BF(y0), BF(y0_lo), b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_const_input_fwd(const float* cst, const float* bias, const float* nw, const float* noise,
                                    const float* sp1, const float* s1, void* a0, void* y0, int b, int c,
                                    cudaStream_t stream) {
  return irfd_const_input_split_fwd(cst, bias, nw, noise, sp1, s1, a0, y0, nullptr, b, c, stream);
}

extern "C" int irfd_const_input_bwd(const void* dy, const void* a0, const float* noise, const float* sp1, float* dsp1,
                                    float* ds1, float* dconst, float* dbias, float* dnw, int b, int c,
                                    cudaStream_t stream) {
  IRFD_CHECK_ARG(dy && a0 && noise && sp1 && dsp1 && ds1 && dconst && dbias && dnw, "const_input_bwd: null pointer");
  const_input_bwd_kernel<<<(c + 31) / 32, 32 * kCiWarps, 0, stream>>>(CBF(dy), CBF(a0), noise, sp1, dsp1, ds1, dconst,
                                                                      dbias, dnw, b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_upsample2x_split_fwd(const void* in, const void* in_lo, void* out, int b, int h, int w, int c,
                                         cudaStream_t stream) {
  IRFD_CHECK_ARG(in && out && c % 8 == 0, "upsample2x_fwd: bad argument");
  IRFD_CHECK_ARG(b > 0 && h > 0 && w > 0 && (long long)w * c < (1ll << 24), "upsample2x_fwd: bad shape");
  const dim3 grid((unsigned)(b * h), (unsigned)((w * (c / 8) + 255) / 256));
  if (in_lo != nullptr)
    upsample2x_fwd_kernel<true><<<grid, 256, 0, stream>>>(CBF(in), CBF(in_lo), BF(out), b, h, w, c);
  else
    upsample2x_fwd_kernel<false><<<grid, 256, 0, stream>>>(CBF(in), nullptr, BF(out), b, h, w, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_upsample2x_fwd(const void* in, void* out, int b, int h, int w, int c, cudaStream_t stream) {
  return irfd_upsample2x_split_fwd(in, nullptr, out, b, h, w, c, stream);
}

extern "C" int irfd_upsample2x_bwd(const void* dout, void* din, int b, int h, int w, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(dout && din && c % 8 == 0, "upsample2x_bwd: bad argument");
  IRFD_CHECK_ARG(b > 0 && h > 0 && w > 0 && (long long)w * c < (1ll << 24), "upsample2x_bwd: bad shape");
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&upsample2x_bwd_kernel));
  upsample2x_bwd_kernel<<<dim3((unsigned)(b * ((h + 1) / 2)), (unsigned)((((w + 1) / 2) * (c / 8) + 255) / 256)), 256, 0,
                          stream>>>(
      CBF(dout), BF(din), b, h, w, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

static void style_plan(int hw, int c, int b, int* chunks, int* rpb) {
  int rows_par = kRvThreads / (c / 8);
  if (rows_par < 1) rows_par = 1;
  // grid = chunks x images, two blocks resident per SM: as many chunks per image as still fit ONE wave (ncu, round 2:
  // ceil(SMs / b) chunks gave 192 blocks at one block per SM = 1.3 waves, the second almost empty)
  long long want = (2ll * num_sms()) / b;  // chunks per image
  if (want < 1) want = 1;
  long long r = (hw + want - 1) / want;
  r = ((r + rows_par - 1) / rows_par) * rows_par;
  if (r < rows_par * 4) r = rows_par * 4;
  *rpb = (int)r;
  *chunks = (int)((hw + r - 1) / r);
}

extern "C" long long irfd_style_bwd_workspace_bytes(int b, int hw, int c) {
  int chunks, rpb;
  style_plan(hw, c, b, &chunks, &rpb);
  return (long long)b * chunks * 4 * c * 4;
}

extern "C" int irfd_style_bwd(const void* dy, const void* a, const float* noise, const float* sp1, void* dz, float* ds1,
                              float* dsp1, float* dbias, float* dnw, int b, int hw, int c, void* workspace,
                              long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(dy && a && noise && sp1 && dz && ds1 && dsp1 && dbias && dnw && workspace, "style_bwd: null pointer");
  IRFD_CHECK_ARG(c % 8 == 0 && c <= 2048, "style_bwd: C must be a multiple of 8 and <= 2048");
  int chunks, rpb;
  style_plan(hw, c, b, &chunks, &rpb);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)b * chunks * 4 * c * 4, "style_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&style_bwd_kernel));
  prefer_max_shared_carveout(reinterpret_cast<const void*>(&style_bwd_finalize_kernel));
  style_bwd_kernel<<<dim3(chunks, b), kRvThreads, 2 * 2048 * sizeof(float), stream>>>(CBF(dy), CBF(a), noise, sp1,
                                                                                     BF(dz), partial, hw, c, rpb);
  IRFD_CHECK_LAUNCH();
  style_bwd_finalize_kernel<<<(c + 31) / 32, 1024, 0, stream>>>(partial, b, chunks, c, ds1, dsp1, dbias, dnw);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_to_rgb_fwd(const void* y, const float* w, const float* bias, float* out, int b, int hw, int c,
                               cudaStream_t stream) {
  IRFD_CHECK_ARG(y && w && bias && out && c % 8 == 0, "to_rgb_fwd: bad argument");
  const size_t pixels = (size_t)b * hw;
  to_rgb_fwd_kernel<<<(unsigned)((pixels + 255) / 256), 256, 3 * c * sizeof(float), stream>>>(CBF(y), w, bias, out, b,
                                                                                             hw, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" long long irfd_to_rgb_bwd_workspace_bytes(int b, int hw, int c) {
  int nblk, rpb;
  plan_row_blocks((long long)b * hw, c, num_sms(), &nblk, &rpb);
  return (long long)nblk * 4 * c * 4;
}

extern "C" int irfd_to_rgb_bwd(const float* drgb, const void* y, const float* w, void* dy, float* dw, float* dbias,
                               int b, int hw, int c, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(drgb && y && w && dy && dw && dbias && workspace && c % 8 == 0 && c <= 2048, "to_rgb_bwd: bad arg");
  IRFD_CHECK_ARG((long long)b * hw < (1ll << 31), "to_rgb_bwd: too many pixels");
  int nblk, rpb;
  plan_row_blocks((long long)b * hw, c, num_sms(), &nblk, &rpb);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)nblk * 4 * c * 4, "to_rgb_bwd: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  to_rgb_bwd_kernel<<<nblk, kRvThreads, 4 * 2048 * sizeof(float), stream>>>(drgb, CBF(y), w, BF(dy), partial, b, hw, c,
                                                                           rpb);
  IRFD_CHECK_LAUNCH();
  to_rgb_bwd_finalize_kernel<<<(3 * c + 3 + 127) / 128, 128, 0, stream>>>(partial, nblk, c, dw, dbias);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_apply_noise_nchw(const float* x, const float* w, const float* noise, float* out, int b, int c, int hw,
                                     cudaStream_t stream) {
  IRFD_CHECK_ARG(x && w && noise && out && b > 0 && c > 0 && hw > 0, "apply_noise_nchw: bad argument");
  const size_t total = (size_t)b * c * hw;
  apply_noise_nchw_kernel<<<GRID1D(total)>>>(x, w, noise, out, c, hw, total);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_apply_style_nchw(const float* x, const float* style, float* out, int b, int c, int hw,
                                     cudaStream_t stream) {
  IRFD_CHECK_ARG(x && style && out && b > 0 && c > 0 && hw > 0, "apply_style_nchw: bad argument");
  const size_t total = (size_t)b * c * hw;
  apply_style_nchw_kernel<<<GRID1D(total)>>>(x, style, out, c, hw, total);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
