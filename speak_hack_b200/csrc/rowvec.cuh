// Shared thread mapping for memory-bound kernels over NHWC bf16 activations viewed as a [rows, C] matrix.
//
// A 256-thread block covers `rows_par = 256 / (C/8)` rows per iteration; every thread owns ONE fixed 8-channel vector
// (16-byte loads/stores, fully coalesced: a row of C bf16 is contiguous) so per-channel parameters live in registers.
// Requires C % 8 == 0 and C/8 <= 256 (C <= 2048), true for every tensor on the IRFD path.
#pragma once
#include <stdlib.h>
#include "ptx.cuh"

namespace irfd {

constexpr int kRvThreads = 256;

struct RowVec {
  int cv;        // channel-vector index of this thread (channels cv*8 .. cv*8+7)
  int row_lane;  // which of the rows_par parallel rows
  int rows_par;
  bool active;
  __device__ RowVec(int C) {
    const int vec_per_row = C >> 3;
    rows_par = kRvThreads / vec_per_row;
    if (rows_par < 1) rows_par = 1;
    cv = threadIdx.x % vec_per_row;
    row_lane = threadIdx.x / vec_per_row;
    active = row_lane < rows_par;
  }
};

// Rows a thread keeps in flight per loop iteration in the streaming kernels: with one row per iteration a B200 SM only
// has ~20 KB of loads outstanding (measured 2.4 TB/s); four rows cover the HBM latency-bandwidth product.
constexpr int kRowBatch = 4;

// 16-byte read-only load of 8 bf16, kept packed so that a batch of loads costs 4 registers each until it is consumed.
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack8(u, f);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void loadf8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Block-level reduction of K per-thread 8-channel accumulators across the rows_par row lanes.
// smem must hold K * rows_par * C floats (<= K * 2048 floats).  Result for (k, channel) is written by the threads of
// row lane 0 to dst[k * dst_stride + channel].
template <int K>
__device__ __forceinline__ void block_reduce_rows(const RowVec& rv, int C, float (&acc)[K][8], float* smem, float* dst,
                                                  size_t dst_stride) {
  if (rv.active) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int t = 0; t < 8; ++t) smem[((size_t)k * rv.rows_par + rv.row_lane) * C + rv.cv * 8 + t] = acc[k][t];
  }
  __syncthreads();
  if (rv.active && rv.row_lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float s = 0.f;
        for (int l = 0; l < rv.rows_par; ++l) s += smem[((size_t)k * rv.rows_par + l) * C + rv.cv * 8 + t];
        dst[k * dst_stride + rv.cv * 8 + t] = s;
      }
  }
}

// Blocks per SM the row-streaming kernels are sized for (IRFD_ROW_WAVES overrides it for experiments).
inline int row_block_waves() {
  static int waves = [] {
    const char* e = getenv("IRFD_ROW_WAVES");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 1;  // measured on B200: 1 block per SM 710 pairs/s, 2: 697, 4: 686, 8: 667 (per-block prologue + reduction dominate)
  }();
  return waves;
}

// How many row blocks to launch for a [rows, C] tensor: enough to fill the machine, not so many that partial
// buffers explode.  Also returns rows per block (multiple of rows_par).
inline void plan_row_blocks(long long rows, int C, int sms, int* nblk, int* rows_per_blk) {
  int rows_par = kRvThreads / (C / 8);
  if (rows_par < 1) rows_par = 1;
  long long target = (long long)sms * row_block_waves();
  long long rpb = (rows + target - 1) / target;
  rpb = ((rpb + rows_par - 1) / rows_par) * rows_par;
  if (rpb < rows_par * 4) rpb = rows_par * 4;
  *rows_per_blk = (int)rpb;
  *nblk = (int)((rows + rpb - 1) / rpb);
}

}  // namespace irfd
