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

// ---- launches; programmatic dependent launch (PDL) as an opt-in experiment ---------------------------------
// Every kernel of this library is launched through launch_kernel() and starts with pdl_entry()
// ("griddepcontrol.wait": a no-op unless the launch carries the programmatic-stream-serialization attribute,
// in which case it holds every thread until the predecessor grid has completed and its memory is visible; no
// kernel touches global memory before it). With MSIG_PDL=1 / msig_debug_set_pdl(1) the attribute is set and the
// next kernel's grid is set up while the blocks of its predecessor drain (stream capture records programmatic
// graph edges). MEASURED (profiles/probe/pdl_r2.txt, B=32 train step, same box A/B): no gain -- 79.95 / 80.05 ms
// against 79.92 / 79.73 ms for plain launches (CUPTI: the launch gaps of the replayed step total ~0.6 ms, and the
// persistent tcgen05 kernels own a whole SM, so a dependent block becomes resident only when its predecessor's
// block exits); with an additional early "launch_dependents" trigger at kernel entry the step was 0.6 ms SLOWER
// (81.0 vs 80.4 ms), so the early trigger is not built and PDL stays off by default.
bool pdl_enabled();
void set_pdl(bool on);

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_entry() {
#ifndef MSIG_NO_GRIDDEP
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#ifdef MSIG_LAUNCH_CHEVRON     // probe builds: the classic launch syntax
#define MSIG_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#else
#define MSIG_LAUNCH(kern, grid, block, smem, stream, ...) \
  (void)msig::launch_kernel(kern, grid, block, smem, stream, __VA_ARGS__)
#endif
#endif

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
