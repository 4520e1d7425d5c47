// Weight-gradient GEMM for stride-1 "same" convolutions (ksize 1 or 3) on tcgen05 tensor cores.
//
//   dWt[k, n] = sum_pixels  X[pixel + shift(tap), c] * dY[pixel, n]        k = tap*Cin + c,  n = out channel
//
// The reduction runs over PIXELS, which is the strided dimension of NHWC, so both operands are fed to the tensor core
// MN-major (tcgen05 transposes in the datapath; no transposed copies are ever materialised):
//   A tile (M = 128 = two 64-channel atoms of (tap, c-chunk))  <- two 4-D TMA boxes [64 pixels x 64 ch] of X, shifted
//   B tile (N = BLOCK_N out channels = BLOCK_N/64 atoms)        <- 2-D TMA boxes [64 pixels x 64 ch] of dY
// Split-K over pixel ranges: each CTA owns (k-pair, n-tile, pixel-range) and writes an fp32 partial
// [split][Ktot][Cout]; irfd_wgrad_reduce sums the splits in a fixed order (deterministic) into OIHW fp32.
#include "host_util.h"
#include "ptx.cuh"

namespace irfd {

struct WgradArgs {
  int M_total, N_total, K_total;
  int num_pairs, num_n_tiles, splits;
  int num_kb, kb_per_split;
  int taps, kw, pad, cin_chunks, num_atoms;
  int H, W;
  float* partial;  // [splits][K_total][N_total]
};

constexpr int kWgThreads = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kWgWaves = 1;      // split-K target: CTAs per SM over the launch (1 measured faster than 2: fewer partials)
constexpr int kWgMinKb = 16;    // shorter splits only multiply prologues and fp32 partial traffic

template <int BLOCK_N>
struct WgCfg {
  static constexpr int A_BYTES = 2 * 8192;
  static constexpr int B_BYTES = (BLOCK_N / 64) * 8192;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RAW_STAGES = (232448 - 1024 - 1024) / STAGE_BYTES;
  static constexpr int STAGES = RAW_STAGES > 8 ? 8 : RAW_STAGES;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 1024;
  static constexpr uint32_t TMEM_COLS = BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256);
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                  const WgradArgs p) {
  using Cfg = WgCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  int item = blockIdx.x;
  const int split = item % p.splits;
  item /= p.splits;
  const int n_tile = item % p.num_n_tiles;
  const int pair = item / p.num_n_tiles;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.num_kb) kb_end = p.num_kb;
  const int my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_x);
      tma_prefetch_desc(&map_dy);
      int atom[2];
      atom[0] = pair * 2;
      atom[1] = pair * 2 + 1 < p.num_atoms ? pair * 2 + 1 : p.num_atoms - 1;  // odd tail: duplicate, discarded later
      int dh[2], dw[2], c0[2];
      for (int i = 0; i < 2; ++i) {
        const int tap = atom[i] / p.cin_chunks;
        c0[i] = (atom[i] - tap * p.cin_chunks) * 64;
        dh[i] = tap / p.kw - p.pad;
        dw[i] = tap % p.kw - p.pad;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int p0 = kb * 64;
        const int w0 = p0 % p.W;
        const int h0 = (p0 / p.W) % p.H;
        const int n0 = p0 / (p.W * p.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_4d(sa, &map_x, &full_bar[stage], c0[0], w0 + dw[0], h0 + dh[0], n0);
        tma_load_4d(sa + 8192, &map_x, &full_bar[stage], c0[1], w0 + dw[1], h0 + dh[1], n0);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_2d(sb + j * 8192, &map_dy, &full_bar[stage], n_tile * BLOCK_N + j * 64, p0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp converged, one elected lane issues (keeps descriptors in uniform registers)
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);  // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < my_kb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
      const uint64_t adesc0 = make_smem_desc_sw128(a_addr, 8192, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 8192, 1024);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 16 pixels (= 16 rows of 128 B = 2048 B = 128 address units) per MMA
          umma_bf16(tmem_base, adesc0 + 128 * k, bdesc0 + 128 * k, idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // epilogue: 4 warps, thread <-> TMEM lane <-> k row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int atom_i = r >> 6;
    const bool valid = (pair * 2 + atom_i) < p.num_atoms;
    const size_t krow = (size_t)pair * 128 + r;
    float* dst = p.partial + ((size_t)split * p.K_total + krow) * p.N_total + (size_t)n_tile * BLOCK_N;
    if (my_kb > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      uint32_t v[32];
      if (my_kb > 0) {
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + chunk * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = 0u;
      }
      if (valid) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          float4 o = make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]),
                                 __uint_as_float(v[4 * t + 2]), __uint_as_float(v[4 * t + 3]));
          *reinterpret_cast<float4*>(dst + chunk * 32 + 4 * t) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// dW[o][c][tap] (OIHW, fp32) = beta * dW + sum_s partial[s][tap*Cin + c][o]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int K_total,
                                    int N_total, int cin, int taps, float beta) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [K_total][N_total], n fastest
  if (idx >= (size_t)K_total * N_total) return;
  const int n = idx % N_total;
  const int k = idx / N_total;
  if (k >= cin * taps) return;  // zero-padded K tail (stem)
  float acc = 0.f;
  const size_t stride = (size_t)K_total * N_total;
  for (int s = 0; s < splits; ++s) acc += partial[s * stride + idx];
  const int tap = k / cin, c = k - tap * cin;
  float* d = dw + ((size_t)n * cin + c) * taps + tap;
  *d = (beta != 0.f) ? beta * (*d) + acc : acc;
}

template <int BLOCK_N>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdy, const WgradArgs& a, cudaStream_t stream) {
  using Cfg = WgCfg<BLOCK_N>;
  static bool configured = false;
  auto kern = wgrad_gemm_kernel<BLOCK_N>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return IRFD_ERR_CUDA;
    }
    configured = true;
  }
  const int grid = a.num_pairs * a.num_n_tiles * a.splits;
  kern<<<grid, kWgThreads, Cfg::SMEM_BYTES, stream>>>(mx, mdy, a);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}

static void wgrad_plan(int n, int h, int w, int cin, int cout, int ksize, WgradArgs* a, int* block_n) {
  const long long m_total = (long long)n * h * w;
  a->M_total = (int)m_total;
  a->N_total = cout;
  a->taps = ksize * ksize;
  a->kw = ksize;
  a->pad = ksize / 2;
  a->cin_chunks = cin / 64;
  a->K_total = a->taps * cin;
  a->num_atoms = a->taps * a->cin_chunks;
  a->num_pairs = (a->num_atoms + 1) / 2;
  *block_n = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
  a->num_n_tiles = cout / *block_n;
  a->num_kb = (int)((m_total + 63) / 64);
  // One CTA per SM is resident (200 KB of pipeline stages), so the launch must fit ONE wave: rounding the split count
  // up left a second wave of a handful of CTAs (e.g. 153 = 148 + 5) that doubled the kernel time.
  const int base = a->num_pairs * a->num_n_tiles;
  int splits = (kWgWaves * num_sms()) / base;
  const int max_splits = (a->num_kb + kWgMinKb - 1) / kWgMinKb;  // at least kWgMinKb k-blocks (64 pixels each) per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a->kb_per_split = (a->num_kb + splits - 1) / splits;
  a->splits = (a->num_kb + a->kb_per_split - 1) / a->kb_per_split;
}

}  // namespace irfd

using namespace irfd;

extern "C" long long irfd_wgrad_workspace_bytes(int n, int h, int w, int cin, int cout, int ksize) {
  WgradArgs a;
  int bn;
  wgrad_plan(n, h, w, cin, cout, ksize, &a, &bn);
  return (long long)a.splits * a.K_total * a.N_total * 4;
}

extern "C" int irfd_conv_wgrad(const void* x, const void* dy, int n, int h, int w, int cin, int cout, int ksize,
                               float* dw, float beta, int reduce_cin, int reduce_taps, void* workspace,
                               long long workspace_bytes, cudaStream_t stream) {
  IRFD_CHECK_ARG(x && dy && dw && workspace, "conv_wgrad: null pointer");
  IRFD_CHECK_ARG(ksize == 1 || ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  IRFD_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0, "conv_wgrad: channels must be multiples of 64");
  WgradArgs a;
  int block_n;
  wgrad_plan(n, h, w, cin, cout, ksize, &a, &block_n);
  IRFD_CHECK_ARG(workspace_bytes >= (long long)a.splits * a.K_total * a.N_total * 4, "conv_wgrad: workspace too small");
  a.partial = reinterpret_cast<float*>(workspace);

  int H = h, W = w, NB = n;
  if (ksize == 1) {
    H = 1; W = a.M_total; NB = 1;
  }
  int tw, th, tn;
  if (W >= 64) {
    IRFD_CHECK_ARG(ksize == 1 || W % 64 == 0, "conv_wgrad: W=%d must be a multiple of 64", W);
    tw = 64; th = 1; tn = 1;
  } else {
    IRFD_CHECK_ARG(64 % W == 0, "conv_wgrad: W=%d must divide 64", W);
    tw = W;
    const int rows = 64 / W;
    if (H >= rows) {
      IRFD_CHECK_ARG(H % rows == 0, "conv_wgrad: H=%d must be a multiple of %d", H, rows);
      th = rows; tn = 1;
    } else {
      IRFD_CHECK_ARG(rows % H == 0, "conv_wgrad: H=%d must divide %d", H, rows);
      th = H; tn = rows / H;
    }
  }
  a.H = H; a.W = W;

  CUtensorMap mx, mdy;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    const uint64_t str[3] = {(uint64_t)cin * 2, (uint64_t)W * cin * 2, (uint64_t)H * W * cin * 2};
    const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
    int rc = make_tmap_bf16(&mx, x, 4, dims, str, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)cout, (uint64_t)a.M_total};
    const uint64_t str[1] = {(uint64_t)cout * 2};
    const uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&mdy, dy, 2, dims, str, box, true);
    if (rc) return rc;
  }
  int rc;
  switch (block_n) {
    case 64: rc = launch_wgrad<64>(mx, mdy, a, stream); break;
    case 128: rc = launch_wgrad<128>(mx, mdy, a, stream); break;
    default: rc = launch_wgrad<256>(mx, mdy, a, stream); break;
  }
  if (rc) return rc;
  const size_t total = (size_t)a.K_total * a.N_total;
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a.partial, dw, a.splits, a.K_total,
                                                                          a.N_total, reduce_cin > 0 ? reduce_cin : cin,
                                                                          reduce_cin > 0 ? reduce_taps : a.taps, beta);
  IRFD_CHECK_LAUNCH();
  return IRFD_OK;
}
