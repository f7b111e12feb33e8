// Host-side helpers shared by the launchers: error reporting, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/adp_b200.h"

namespace adp {

int set_error(const char* fmt, ...);  // returns non-zero, records the message (thread-local)

#define ADP_CHECK(cond, ...)                         \
  do {                                               \
    if (!(cond)) return ::adp::set_error(__VA_ARGS__); \
  } while (0)

#define ADP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return ::adp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                              __FILE__, __LINE__);                                       \
  } while (0)

#define ADP_LAUNCH_CHECK() ADP_CUDA(cudaGetLastError())

// bf16 tensor map, up to 4 dims.  dims[0] is the contiguous dimension; strides_bytes[i] is
// the byte stride of dims[i+1].  swizzle_bytes in {0, 32, 64, 128}.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

inline cudaStream_t as_stream(adp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace adp
