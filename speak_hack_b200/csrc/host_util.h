// Host-side helpers shared by the C-ABI translation units: error codes, driver entry point for
// cuTensorMapEncodeTiled (resolved at run time so the library does not link libcuda), device queries.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/irfd_b200.h"

namespace irfd {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled();
int num_sms();
void set_last_error(const char* fmt, ...);

// bf16 tensor map of rank `rank`; dims[0] is the contiguous dimension; strides_bytes has rank-1 entries.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128);

// Ask for the maximum shared-memory carveout for a (memory-bound, small-smem) kernel.  An SM can only run kernels of
// ONE L1/shared split at a time, and the persistent GEMM kernels configure the SMs for ~227 KB of shared memory; a
// row-streaming kernel left on the default split cannot become co-resident with a weight-gradient GEMM running on the
// side stream (ops.SideStream) until the SM drains.  These kernels stream with 16-byte loads and do not rely on L1.
void prefer_max_shared_carveout(const void* kernel);

#define IRFD_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      irfd::set_last_error(__VA_ARGS__);     \
      return IRFD_ERR_INVALID_ARGUMENT;      \
    }                                        \
  } while (0)

#define IRFD_CHECK_LAUNCH()                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      irfd::set_last_error("%s:%d CUDA launch: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return IRFD_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

}  // namespace irfd
