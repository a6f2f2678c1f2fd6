// Loss reductions (+ their gradients) and the optimizer-side multi-tensor kernels.
// Reference call sites: trainer.py:50-52 (MSELoss / L1Loss), :99,103,108,116-117,142-147;
// losses.py:80-98 (L1 of Gram matrices, L1 of features); trainer.py:127-134,152-153 and
// utils.py:80-91 (clip_grad_norm_ + Adam + EMA).
#include "common.h"

#include <algorithm>

namespace msig {

__device__ __forceinline__ float warp_sum_l(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Deterministic grid-wide sum (no floating-point atomics): every block writes its partial sum to
// `partial[block]`; the last block to arrive (integer ticket, reset for the next launch) adds all partials
// in a fixed order and writes out = (accumulate ? out : 0) + scale * total. The summation order depends
// on the grid size only, so an eager step and its CUDA-graph replay produce identical bits.
// Call from ALL threads of a 1-D block of 256 threads (tid = linear thread id, nblocks = grid size).
__device__ __forceinline__ float block_sum_256(float v, int tid) {
  __shared__ float ws[8];
  __syncthreads();                       // (ws may still be read by a previous call)
  v = warp_sum_l(v);
  if ((tid & 31) == 0) ws[tid >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += ws[i];
  return t;                              // every thread holds the block total
}
__device__ __forceinline__ void grid_sum_ordered(float v, float scale, float* out, int accumulate,
                                                 float* partial, unsigned int* ticket, int tid, int block,
                                                 int nblocks) {
  __shared__ bool last;
  const float t = block_sum_256(v, tid);
  if (tid == 0) {
    partial[block] = t;
    __threadfence();
    const unsigned int k = atomicAdd(ticket, 1u);
    last = (k == static_cast<unsigned int>(nblocks) - 1u);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float s = 0.f;
  for (int i = tid; i < nblocks; i += 256) s += __ldcg(partial + i);
  s = block_sum_256(s, tid);
  if (tid == 0) *out = (accumulate ? *out : 0.f) + s * scale;
}
__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

static inline int lgrid(int64_t work, int threads) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), 148 * 8)));
}

__global__ void __launch_bounds__(256) l1_f32_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         int64_t n, float inv_n, float* __restrict__ loss,
                                                         float* __restrict__ partial, unsigned int* ticket) {
  pdl_entry();
  float s = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    s += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += fabsf(a[i] - b[i]);
  grid_sum_ordered(s, inv_n, loss, 0, partial, ticket, threadIdx.x, blockIdx.x, gridDim.x);
}
__global__ void l1_f32_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                  float inv_n, const float* __restrict__ gscale, float* __restrict__ g) {
  pdl_entry();
  const float k = inv_n * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    g[i] = sgn(a[i] - b[i]) * k;
}
__global__ void __launch_bounds__(256) l1_bf16_fwd_kernel(const __nv_bfloat16* __restrict__ a,
                                                          const __nv_bfloat16* __restrict__ b, int64_t groups,
                                                          float inv_n, float* __restrict__ loss,
                                                          float* __restrict__ partial, unsigned int* ticket) {
  pdl_entry();
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < groups; i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i];
    const uint4 ub = reinterpret_cast<const uint4*>(b)[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(ha[j]), y = __bfloat1622float2(hb[j]);
      s += fabsf(x.x - y.x) + fabsf(x.y - y.y);
    }
  }
  grid_sum_ordered(s, inv_n, loss, 0, partial, ticket, threadIdx.x, blockIdx.x, gridDim.x);
}
__global__ void l1_bf16_bwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                   int64_t groups, float inv_n, const float* __restrict__ gscale,
                                   const __nv_bfloat16* __restrict__ aux, __nv_bfloat16* __restrict__ g) {
  pdl_entry();
  const float k = inv_n * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < groups; i += int64_t(gridDim.x) * blockDim.x) {
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i];
    const uint4 ub = reinterpret_cast<const uint4*>(b)[i];
    uint4 ux = make_uint4(0, 0, 0, 0);
    if (aux) ux = reinterpret_cast<const uint4*>(aux)[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
    const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
    const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&ux);
    uint4 uo;
    __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&uo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = __bfloat1622float2(ha[j]), y = __bfloat1622float2(hb[j]), z = __bfloat1622float2(hx[j]);
      ho[j] = __floats2bfloat162_rn(sgn(x.x - y.x) * k + z.x, sgn(x.y - y.y) * k + z.y);
    }
    reinterpret_cast<uint4*>(g)[i] = uo;
  }
}
// mean (a - target)^2; the target is the constant `t` (LSGAN all-ones / all-zeros) or the tensor `b`
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      float t, int64_t n, float inv_n, float* __restrict__ loss,
                                                      float* __restrict__ partial, unsigned int* ticket) {
  pdl_entry();
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = a[i] - (b ? b[i] : t);
    s += d * d;
  }
  grid_sum_ordered(s, inv_n, loss, 0, partial, ticket, threadIdx.x, blockIdx.x, gridDim.x);
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float t, int64_t n,
                               float inv_n, const float* __restrict__ gscale, float* __restrict__ g) {
  pdl_entry();
  const float k = 2.f * inv_n * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    g[i] = (a[i] - (b ? b[i] : t)) * k;
}

// loss = mean |Ga - Gb| over the full symmetric matrix; ssym = 2*sign(D) (= sign(D) + sign(D)^T).
// Only 32x32 tiles on/above the diagonal are read (the Gram kernel skips tiles below it); a block
// below the diagonal reads its mirror tile and transposes it through shared memory (coalesced).
__global__ void __launch_bounds__(256) gram_l1_kernel(const float* __restrict__ ga, const float* __restrict__ gb,
                                                      int dim, float inv_n, float* __restrict__ loss, int accumulate,
                                                      __nv_bfloat16* __restrict__ ssym, float* __restrict__ partial,
                                                      unsigned int* ticket) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // this block: rows by.., cols bx..
  const bool upper = blockIdx.x >= blockIdx.y;
  float s = 0.f;
  if (!upper) {
    for (int r = ty; r < 32; r += 8) {
      const int i = bx + r, j = by + tx;                  // mirror tile (above the diagonal)
      float d = 0.f;
      if (i < dim && j < dim) d = ga[int64_t(i) * dim + j] - gb[int64_t(i) * dim + j];
      tile[r][tx] = d;
    }
    __syncthreads();
  }
  for (int r = ty; r < 32; r += 8) {
    const int i = by + r, j = bx + tx;
    if (i < dim && j < dim) {
      const float d = upper ? ga[int64_t(i) * dim + j] - gb[int64_t(i) * dim + j] : tile[tx][r];
      s += fabsf(d);
      ssym[int64_t(i) * dim + j] = __float2bfloat16(2.f * sgn(d));
    }
  }
  grid_sum_ordered(s, inv_n, loss, accumulate, partial, ticket, ty * 32 + tx, blockIdx.y * gridDim.x + blockIdx.x,
                   gridDim.x * gridDim.y);
}

__global__ void colsum_f32_kernel(const float* __restrict__ x, int64_t rows, int c, int64_t ld,
                                  float* __restrict__ out, int accumulate) {
  pdl_entry();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int64_t r = 0; r < rows; ++r) s += x[r * ld + ch];
  out[ch] = accumulate ? out[ch] + s : s;
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out,
                                                    int accumulate, float* __restrict__ partial,
                                                    unsigned int* ticket) {
  pdl_entry();
  float s = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
  grid_sum_ordered(s, 1.f, out, accumulate, partial, ticket, threadIdx.x, blockIdx.x, gridDim.x);
}

__global__ void counter_inc_kernel(int32_t* c) {
  pdl_entry(); *c += 1; }

// clip_grad_norm_(max_norm) + Adam (torch defaults: no weight decay, no amsgrad) + EMA, one pass.
// step_dev != nullptr: the step number (bias corrections) comes from device memory, so the launch can
// be replayed from a CUDA graph.
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, int64_t n,
                                 const float* __restrict__ grad_sumsq, float max_norm, float grad_scale, float lr,
                                 float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float ema_beta,
                                 const int32_t* __restrict__ step_dev) {
  pdl_entry();
  if (step_dev != nullptr) {
    const double st = double(__ldg(step_dev));
    bc1 = float(1.0 - pow(double(beta1), st));
    bc2_sqrt = float(sqrt(1.0 - pow(double(beta2), st)));
  }
  float coef = grad_scale;
  if (grad_sumsq != nullptr && max_norm > 0.f) {
    const float total = sqrtf(__ldg(grad_sumsq)) * grad_scale;
    const float cc = max_norm / (total + 1e-6f);
    coef *= cc < 1.f ? cc : 1.f;
  }
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - step_size * (mi / denom);
    p[i] = pi;
    if (ema != nullptr) ema[i] = ema[i] * ema_beta + (1.f - ema_beta) * pi;
  }
}

}  // namespace msig

using namespace msig;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

// Reduction scratch of the loss kernels: [blocks] fp32 partial sums followed by one u32 ticket.
static inline float* ws_partial(void* ws) { return reinterpret_cast<float*>(ws); }
static inline unsigned int* ws_ticket(void* ws, int64_t blocks) {
  return reinterpret_cast<unsigned int*>(reinterpret_cast<float*>(ws) + blocks);
}
#define REDUCE_WS_OK(blocks) (workspace != nullptr && workspace_bytes >= (size_t(blocks) + 1) * sizeof(float))

size_t msig_reduce_workspace(void) { return (size_t(148) * 8 + 1) * sizeof(float); }
size_t msig_gram_l1_workspace(int32_t dim) {
  const size_t t = static_cast<size_t>(ceil_div(dim, 32));
  return (t * t + 1) * sizeof(float);
}

int msig_l1_loss_f32_fwd(const float* a, const float* b, int64_t numel, float* loss, void* workspace,
                         size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(a && b && loss && numel > 0, "msig_l1_loss_f32_fwd: bad argument");
  const int blocks = lgrid(numel / 4 + 1, 256);
  MSIG_REQUIRE(REDUCE_WS_OK(blocks), "msig_l1_loss_f32_fwd: workspace too small");
  MSIG_CHECK_CUDA(cudaMemsetAsync(ws_ticket(workspace, blocks), 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((l1_f32_fwd_kernel), blocks, 256, 0, ST(stream), a, b, numel, 1.f / numel, loss, ws_partial(workspace),
                                                   ws_ticket(workspace, blocks));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_f32_bwd(const float* a, const float* b, int64_t numel, const float* gscale, float* grad_a,
                         void* stream) {
  MSIG_REQUIRE(a && b && grad_a && numel > 0, "msig_l1_loss_f32_bwd: bad argument");
  MSIG_LAUNCH((l1_f32_bwd_kernel), lgrid(numel, 256), 256, 0, ST(stream), a, b, numel, 1.f / numel, gscale, grad_a);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_bf16_fwd(const void* a, const void* b, int64_t numel, float* loss, void* workspace,
                          size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(a && b && loss && numel > 0 && numel % 8 == 0, "msig_l1_loss_bf16_fwd: bad argument");
  const int blocks = lgrid(numel / 8, 256);
  MSIG_REQUIRE(REDUCE_WS_OK(blocks), "msig_l1_loss_bf16_fwd: workspace too small");
  MSIG_CHECK_CUDA(cudaMemsetAsync(ws_ticket(workspace, blocks), 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((l1_bf16_fwd_kernel), blocks, 256, 0, ST(stream), CBF(a), CBF(b), numel / 8, 1.f / numel, loss,
                                                    ws_partial(workspace), ws_ticket(workspace, blocks));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_bf16_bwd(const void* a, const void* b, int64_t numel, const float* gscale, const void* aux,
                          void* grad_a, void* stream) {
  MSIG_REQUIRE(a && b && grad_a && numel > 0 && numel % 8 == 0, "msig_l1_loss_bf16_bwd: bad argument");
  MSIG_LAUNCH((l1_bf16_bwd_kernel), lgrid(numel / 8, 256), 256, 0, ST(stream), CBF(a), CBF(b), numel / 8, 1.f / numel, gscale,
                                                                   CBF(aux), BF(grad_a));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
static int mse_fwd(const float* a, const float* b, float target, int64_t numel, float* loss, void* workspace,
                   size_t workspace_bytes, void* stream, const char* what) {
  MSIG_REQUIRE(a && loss && numel > 0, "%s: bad argument", what);
  const int blocks = lgrid(numel, 256);
  MSIG_REQUIRE(REDUCE_WS_OK(blocks), "%s: workspace too small", what);
  MSIG_CHECK_CUDA(cudaMemsetAsync(ws_ticket(workspace, blocks), 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((mse_fwd_kernel), blocks, 256, 0, ST(stream), a, b, target, numel, 1.f / numel, loss, ws_partial(workspace),
                                                ws_ticket(workspace, blocks));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_mse_const_fwd(const float* a, float target, int64_t numel, float* loss, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return mse_fwd(a, nullptr, target, numel, loss, workspace, workspace_bytes, stream, "msig_mse_const_fwd");
}
int msig_mse_const_bwd(const float* a, float target, int64_t numel, const float* gscale, float* grad_a,
                       void* stream) {
  MSIG_REQUIRE(a && grad_a && numel > 0, "msig_mse_const_bwd: bad argument");
  MSIG_LAUNCH((mse_bwd_kernel), lgrid(numel, 256), 256, 0, ST(stream), a, nullptr, target, numel, 1.f / numel, gscale, grad_a);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_mse_loss_fwd(const float* a, const float* target, int64_t numel, float* loss, void* workspace,
                      size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(target != nullptr, "msig_mse_loss_fwd: null target");
  return mse_fwd(a, target, 0.f, numel, loss, workspace, workspace_bytes, stream, "msig_mse_loss_fwd");
}
int msig_mse_loss_bwd(const float* a, const float* target, int64_t numel, const float* gscale, float* grad_a,
                      void* stream) {
  MSIG_REQUIRE(a && target && grad_a && numel > 0, "msig_mse_loss_bwd: bad argument");
  MSIG_LAUNCH((mse_bwd_kernel), lgrid(numel, 256), 256, 0, ST(stream), a, target, 0.f, numel, 1.f / numel, gscale, grad_a);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_gram_l1(const float* ga, const float* gb, int32_t dim, float* loss, int accumulate, void* ssym,
                 void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(ga && gb && loss && ssym && dim > 0, "msig_gram_l1: bad argument");
  const unsigned t = static_cast<unsigned>(ceil_div(dim, 32));
  const int64_t blocks = int64_t(t) * t;
  MSIG_REQUIRE(REDUCE_WS_OK(blocks), "msig_gram_l1: workspace too small");
  MSIG_CHECK_CUDA(cudaMemsetAsync(ws_ticket(workspace, blocks), 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((gram_l1_kernel), dim3(t, t), dim3(32, 8), 0, ST(stream), ga, gb, dim, 1.f / (float(dim) * float(dim)), loss,
                                                             accumulate, BF(ssym), ws_partial(workspace),
                                                             ws_ticket(workspace, blocks));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_colsum_f32(const float* x, int64_t rows, int32_t c, int64_t ld, float* out, int accumulate,
                    void* stream) {
  MSIG_REQUIRE(x && out && c > 0, "msig_colsum_f32: bad argument");
  MSIG_LAUNCH((colsum_f32_kernel), static_cast<unsigned>(ceil_div(c, 128)), 128, 0, ST(stream), x, rows, c, ld, out, accumulate);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_sumsq(const float* x, int64_t numel, float* out, int accumulate, void* workspace, size_t workspace_bytes,
               void* stream) {
  MSIG_REQUIRE(x && out && numel > 0, "msig_sumsq: bad argument");
  const int blocks = lgrid(numel / 4 + 1, 256);
  MSIG_REQUIRE(REDUCE_WS_OK(blocks), "msig_sumsq: workspace too small");
  MSIG_CHECK_CUDA(cudaMemsetAsync(ws_ticket(workspace, blocks), 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((sumsq_kernel), blocks, 256, 0, ST(stream), x, numel, out, accumulate, ws_partial(workspace),
                                              ws_ticket(workspace, blocks));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t numel,
                   const float* grad_sumsq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                   float eps, int32_t step, float ema_beta, void* stream) {
  MSIG_REQUIRE(param && grad && exp_avg && exp_avg_sq && numel > 0 && step >= 1, "msig_adam_step: bad argument");
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  MSIG_LAUNCH((adam_step_kernel), lgrid(numel, 256), 256, 0, ST(stream), param, grad, exp_avg, exp_avg_sq, ema, numel,
                                                             grad_sumsq, max_norm, grad_scale, lr, beta1, beta2, eps,
                                                             float(bc1), float(sqrt(bc2)), ema_beta, nullptr);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t numel,
                       const float* grad_sumsq, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                       float eps, int32_t* step_counter, float ema_beta, void* stream) {
  MSIG_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_counter && numel > 0, "msig_adam_step_dev: bad argument");
  MSIG_LAUNCH((counter_inc_kernel), 1, 1, 0, ST(stream), step_counter);
  MSIG_CHECK_LAUNCH();
  MSIG_LAUNCH((adam_step_kernel), lgrid(numel, 256), 256, 0, ST(stream), param, grad, exp_avg, exp_avg_sq, ema, numel,
                                                             grad_sumsq, max_norm, grad_scale, lr, beta1, beta2, eps,
                                                             1.f, 1.f, ema_beta, step_counter);
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

}  // extern "C"
