// Shared host-side helpers of the C-ABI implementation files.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/msig.h"

namespace msig {

int set_error(int code, const char* fmt, ...);
void count_launch(int n);
int sm_count();
bool context_ready();

// cuTensorMapEncodeTiled resolved at msig_init (no link-time dependency on libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled();

#define MSIG_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess)                                                           \
      return msig::set_error(MSIG_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define MSIG_CHECK_LAUNCH()                                                              \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return msig::set_error(MSIG_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
  } while (0)

#define MSIG_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return msig::set_error(MSIG_ERR_ARG, __VA_ARGS__);    \
  } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Rows of a packed weight matrix for `k` logical output channels.
inline int pad_rows(int k) { return k <= 16 ? 16 : static_cast<int>(round_up(k, 64)); }
inline int pick_block_n(int k_pad) {
  if (k_pad == 16) return 16;
  if (k_pad % 256 == 0) return 256;
  if (k_pad % 128 == 0) return 128;
  return 64;
}

}  // namespace msig
