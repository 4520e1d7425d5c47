// Device-side versions of the IRFD path's data-dependent routing, so that a whole training step is ONE static CUDA
// graph: the decisions still come from the CPU generator in the reference's order (model.py:98 swap draw,
// styleganv1.py:548-552 style-mixing draws) but reach the kernels through a small int32 control tensor instead of
// Python control flow.
//   ctrl[0] = swap_type in {0,1,2};  ctrl[1 + g] = first mixed row ("cut") of generator call g, == L when not mixing.
#include "host_util.h"
#include "ptx.cuh"

namespace irfd {

// rows_t[l][b][k] = l >= cut ? w2[b][k] : coef(l) * w[b][k],  coef(l) = psi for l < cutoff else 1
// (styleganv1.py:536-553: repeat -> truncation coefficients on w ONLY -> rows >= mix_layer overwritten with the
// UNtruncated rows of w2)
// split_bk: batch rows are two generator calls stacked (source pairs, then target pairs): elements j >= split_bk of a
// row belong to the second call and read its cut from ctrl[ctrl_idx + 1] (split_bk == BK: one call).
__global__ void style_rows_fwd_kernel(const float* __restrict__ w, const float* __restrict__ w2,
                                      const int* __restrict__ ctrl, int ctrl_idx, float psi, int cutoff,
                                      float* __restrict__ rows_t, int L, int BK, int split_bk) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)L * BK) return;
  const int l = i / BK;
  const int j = i - (size_t)l * BK;
  const int cut = ctrl[ctrl_idx + (j >= split_bk ? 1 : 0)];
  const float coef = l < cutoff ? psi : 1.f;
  rows_t[i] = l >= cut ? w2[j] : coef * w[j];
}

// dw[b][k] = sum_l coef(l) * drows_t[l][b][k] over ALL rows: the reference overwrites the mixed rows under no_grad, so
// autograd still routes their gradient into the mapping output (SURVEY/DESIGN note on styleganv1.py:549-553).
__global__ void style_rows_bwd_kernel(const float* __restrict__ drows_t, float psi, int cutoff, float* __restrict__ dw,
                                      int L, int BK) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (size_t)BK) return;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) acc += (l < cutoff ? psi : 1.f) * drows_t[(size_t)l * BK + j];
  dw[j] = acc;
}

// S<->T swap of one code type + concatenation [identity | emotion | pose] (model.py:97-108).  Pure copies: bit-exact.
__global__ void swap_cat_fwd_kernel(const float* __restrict__ fi_s, const float* __restrict__ fe_s,
                                    const float* __restrict__ fp_s, const float* __restrict__ fi_t,
                                    const float* __restrict__ fe_t, const float* __restrict__ fp_t,
                                    const int* __restrict__ ctrl, float* __restrict__ gen_s, float* __restrict__ gen_t,
                                    int B, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [B][3][C]
  if (i >= (size_t)B * 3 * C) return;
  const int c = i % C;
  const int part = (i / C) % 3;
  const int b = i / ((size_t)3 * C);
  const float* s = part == 0 ? fi_s : (part == 1 ? fe_s : fp_s);
  const float* t = part == 0 ? fi_t : (part == 1 ? fe_t : fp_t);
  const bool swap = ctrl[0] == part;
  const float vs = s[(size_t)b * C + c], vt = t[(size_t)b * C + c];
  gen_s[i] = swap ? vt : vs;
  gen_t[i] = swap ? vs : vt;
}

__global__ void swap_cat_bwd_kernel(const float* __restrict__ dgen_s, const float* __restrict__ dgen_t,
                                    const int* __restrict__ ctrl, float* __restrict__ dfi_s, float* __restrict__ dfe_s,
                                    float* __restrict__ dfp_s, float* __restrict__ dfi_t, float* __restrict__ dfe_t,
                                    float* __restrict__ dfp_t, int B, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * 3 * C) return;
  const int c = i % C;
  const int part = (i / C) % 3;
  const int b = i / ((size_t)3 * C);
  float* s = part == 0 ? dfi_s : (part == 1 ? dfe_s : dfp_s);
  float* t = part == 0 ? dfi_t : (part == 1 ? dfe_t : dfp_t);
  const bool swap = ctrl[0] == part;
  const float gs = dgen_s[i], gt = dgen_t[i];
  s[(size_t)b * C + c] = swap ? gt : gs;
  t[(size_t)b * C + c] = swap ? gs : gt;
}

}  // namespace irfd

using namespace irfd;

extern "C" int irfd_style_rows_pair_fwd(const float* w, const float* w2, const int* ctrl, int ctrl_idx, float psi,
                                        int cutoff, float* rows_t, int l, int b, int k, int b_first,
                                        cudaStream_t stream) {
  IRFD_CHECK_ARG(w && w2 && ctrl && rows_t && l > 0 && b > 0 && k > 0, "style_rows_fwd: bad argument");
  IRFD_CHECK_ARG(b_first > 0 && b_first <= b, "style_rows_fwd: b_first must be in 1..b");
  const size_t total = (size_t)l * b * k;
  style_rows_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(w, w2, ctrl, ctrl_idx, psi, cutoff, rows_t,
                                                                            l, b * k, b_first * k);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_style_rows_fwd(const float* w, const float* w2, const int* ctrl, int ctrl_idx, float psi,
                                   int cutoff, float* rows_t, int l, int b, int k, cudaStream_t stream) {
  return irfd_style_rows_pair_fwd(w, w2, ctrl, ctrl_idx, psi, cutoff, rows_t, l, b, k, b, stream);
}

extern "C" int irfd_style_rows_bwd(const float* drows_t, float psi, int cutoff, float* dw, int l, int b, int k,
                                   cudaStream_t stream) {
  IRFD_CHECK_ARG(drows_t && dw && l > 0 && b > 0 && k > 0, "style_rows_bwd: bad argument");
  const size_t total = (size_t)b * k;
  style_rows_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(drows_t, psi, cutoff, dw, l, b * k);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_swap_cat_fwd(const float* fi_s, const float* fe_s, const float* fp_s, const float* fi_t,
                                 const float* fe_t, const float* fp_t, const int* ctrl, float* gen_s, float* gen_t,
                                 int b, int c, cudaStream_t stream) {
  IRFD_CHECK_ARG(fi_s && fe_s && fp_s && fi_t && fe_t && fp_t && ctrl && gen_s && gen_t, "swap_cat_fwd: null pointer");
  const size_t total = (size_t)b * 3 * c;
  swap_cat_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fi_s, fe_s, fp_s, fi_t, fe_t, fp_t, ctrl,
                                                                          gen_s, gen_t, b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

extern "C" int irfd_swap_cat_bwd(const float* dgen_s, const float* dgen_t, const int* ctrl, float* dfi_s, float* dfe_s,
                                 float* dfp_s, float* dfi_t, float* dfe_t, float* dfp_t, int b, int c,
                                 cudaStream_t stream) {
  IRFD_CHECK_ARG(dgen_s && dgen_t && ctrl && dfi_s && dfe_s && dfp_s && dfi_t && dfe_t && dfp_t,
                 "swap_cat_bwd: null pointer");
  const size_t total = (size_t)b * 3 * c;
  swap_cat_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dgen_s, dgen_t, ctrl, dfi_s, dfe_s, dfp_s,
                                                                          dfi_t, dfe_t, dfp_t, b, c);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
