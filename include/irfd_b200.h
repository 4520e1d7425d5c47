/* irfd_b200.h — C ABI of libirfd_b200.so: the sm_100a kernels behind the IRFD hot path.
 *
 * The reference (johndpope/SPEAK-hack) has no FFI layer: its hot path is Python `nn.Module` code calling ATen.
 * Each entry point below therefore names the reference call site(s) whose arithmetic it replaces.  The Python host
 * (speak_hack_b200/*.py) keeps the reference's module surface (IRFD, StyleGenerator, ... — model.py:28-126,
 * styleganv1.py:448-635) and calls these functions through ctypes with raw device pointers and the current stream.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - activations are NHWC bf16 ("pixels x channels"); parameters/gradients handed back to PyTorch are fp32 in the
 *     reference's own layouts (OIHW conv weights, [out,in] linear weights);
 *   - functions enqueue work on `stream` and return immediately; they never allocate, never synchronise and keep no
 *     global state beyond immutable lazily-created function attributes;
 *   - return value: IRFD_OK (0) or a negative error code; `irfd_last_error()` gives the message (thread-local).
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every launch returns IRFD_ERR_CUDA.
 */
#ifndef IRFD_B200_H_
#define IRFD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRFD_ABI_VERSION 1

#define IRFD_OK 0
#define IRFD_ERR_INVALID_ARGUMENT (-1)
#define IRFD_ERR_CUDA (-2)

/* Opaque CUDA stream handle (cudaStream_t). */
typedef struct CUstream_st* irfd_stream_t;

int irfd_abi_version(void);
const char* irfd_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, stride 1, "same" padding, ksize in {1, 3}  (tcgen05 + TMEM + TMA).
 * Replaces: nn.Conv2d 3x3 of SynthesisBlock (styleganv1.py:615-616, 625, 630) with its ApplyNoise / leaky_relu /
 *           ApplyStyle tail (styleganv1.py:453-456, 463-468, 626-633) fused as the epilogue (mode 2);
 *           the stride-1 convs of torchvision Bottleneck (torchvision/models/resnet.py:143-155) with the batch
 *           statistics of the following BatchNorm2d gathered in the epilogue (mode 1);
 *           every dgrad of those convs (mode 0, weights pre-flipped by irfd_pack_conv_weight).
 *   x      [n,h,w,cin]  bf16 NHWC        wk  [cout][ksize*ksize][cin] bf16 (tap-major K)
 *   out    [n,h,w,cout] bf16             out2 same shape (mode 2 only: out = a = lrelu(z), out2 = styled y)
 *   mode 0: out = acc (+ bias[cout] if bias != NULL)
 *   mode 1: out = acc; stat_sum/stat_sq [m_tiles][cout] fp32 per-128-pixel-tile partial sums (of the stored bf16)
 *   mode 2: z = acc + bias[c] + nw[c]*noise[pixel]; a = lrelu(z, 0.2); y = a*sp1[b,c] + s1[b,c]
 *           noise [n*h*w] fp32, sp1/s1 [n][cout] fp32 (sp1 = style0 + 1)
 *   force_block_n: 0 = heuristic, else 64/128/256.
 * Shape rules: cin % 64 == 0, cout % 64 == 0; for ksize 3: w a power of two (<=128) or a multiple of 128, and
 * 128 pixels must form whole rows / whole images.
 */
int irfd_conv_gemm_m_tiles(int n, int h, int w);
int irfd_conv_gemm(const void* x, int n, int h, int w, int cin, const void* wk, int cout, int ksize, void* out,
                   void* out2, int mode, const float* bias, const float* nw, const float* noise, const float* sp1,
                   const float* s1, float* stat_sum, float* stat_sq, int force_block_n, irfd_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Weight gradient of the same convolutions (tcgen05, both operands MN-major, deterministic split-K over pixels).
 * Replaces: autograd's convolution_backward(weight) for styleganv1.py:615-616 and torchvision resnet.py:133-141.
 *   x  [n,h,w,cin] bf16 (the conv input)   dy [n,h,w,cout] bf16 (gradient w.r.t. the raw conv output)
 *   dw [cout][cin][ksize][ksize] fp32 (OIHW, the reference parameter layout):  dw = beta*dw + grad
 *   workspace: >= irfd_wgrad_workspace_bytes(...) bytes of device scratch (fp32 split-K partials).
 */
long long irfd_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize);
int irfd_conv_wgrad(const void* x, const void* dy, int n, int h, int w, int cin, int cout, int ksize, float* dw,
                    float beta, void* workspace, long long workspace_bytes, irfd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IRFD_B200_H_ */
