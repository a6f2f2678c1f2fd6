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
// block-wide sum -> one atomicAdd (scaled) per block
__device__ __forceinline__ void block_atomic_add(float v, float scale, float* out) {
  __shared__ float ws[32];
  v = warp_sum_l(v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    t = warp_sum_l(t);
    if (threadIdx.x == 0) atomicAdd(out, t * scale);
  }
}
__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

static inline int lgrid(int64_t work, int threads) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), 148 * 8)));
}

__global__ void l1_f32_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                  float inv_n, float* __restrict__ loss) {
  float s = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    s += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += fabsf(a[i] - b[i]);
  block_atomic_add(s, inv_n, loss);
}
__global__ void l1_f32_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                  float inv_n, const float* __restrict__ gscale, float* __restrict__ g) {
  const float k = inv_n * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    g[i] = sgn(a[i] - b[i]) * k;
}
__global__ void l1_bf16_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                   int64_t groups, float inv_n, float* __restrict__ loss) {
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
  block_atomic_add(s, inv_n, loss);
}
__global__ void l1_bf16_bwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                   int64_t groups, float inv_n, const float* __restrict__ gscale,
                                   const __nv_bfloat16* __restrict__ aux, __nv_bfloat16* __restrict__ g) {
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
__global__ void mse_const_fwd_kernel(const float* __restrict__ a, float t, int64_t n, float inv_n,
                                     float* __restrict__ loss) {
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = a[i] - t;
    s += d * d;
  }
  block_atomic_add(s, inv_n, loss);
}
__global__ void mse_const_bwd_kernel(const float* __restrict__ a, float t, int64_t n, float inv_n,
                                     const float* __restrict__ gscale, float* __restrict__ g) {
  const float k = 2.f * inv_n * (gscale ? __ldg(gscale) : 1.f);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    g[i] = (a[i] - t) * k;
}

// loss = mean |Ga - Gb| over the full symmetric matrix; ssym = 2*sign(D) (= sign(D) + sign(D)^T).
// Only 32x32 tiles on/above the diagonal are read (the Gram kernel skips tiles below it); a block
// below the diagonal reads its mirror tile and transposes it through shared memory (coalesced).
__global__ void gram_l1_kernel(const float* __restrict__ ga, const float* __restrict__ gb, int dim,
                               float inv_n, float* __restrict__ loss, __nv_bfloat16* __restrict__ ssym) {
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
  __shared__ float ws[8];
  s = warp_sum_l(s);
  const int tid = ty * 32 + tx;
  if ((tid & 31) == 0) ws[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += ws[k];
    atomicAdd(loss, t * inv_n);
  }
}

__global__ void colsum_f32_kernel(const float* __restrict__ x, int64_t rows, int c, int64_t ld,
                                  float* __restrict__ out, int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int64_t r = 0; r < rows; ++r) s += x[r * ld + ch];
  out[ch] = accumulate ? out[ch] + s : s;
}

__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float s = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
  block_atomic_add(s, 1.f, out);
}

__global__ void counter_inc_kernel(int32_t* c) { *c += 1; }

// clip_grad_norm_(max_norm) + Adam (torch defaults: no weight decay, no amsgrad) + EMA, one pass.
// step_dev != nullptr: the step number (bias corrections) comes from device memory, so the launch can
// be replayed from a CUDA graph.
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, int64_t n,
                                 const float* __restrict__ grad_sumsq, float max_norm, float grad_scale, float lr,
                                 float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float ema_beta,
                                 const int32_t* __restrict__ step_dev) {
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

int msig_l1_loss_f32_fwd(const float* a, const float* b, int64_t numel, float* loss, void* stream) {
  MSIG_REQUIRE(a && b && loss && numel > 0, "msig_l1_loss_f32_fwd: bad argument");
  MSIG_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  l1_f32_fwd_kernel<<<lgrid(numel / 4 + 1, 256), 256, 0, ST(stream)>>>(a, b, numel, 1.f / numel, loss);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_f32_bwd(const float* a, const float* b, int64_t numel, const float* gscale, float* grad_a,
                         void* stream) {
  MSIG_REQUIRE(a && b && grad_a && numel > 0, "msig_l1_loss_f32_bwd: bad argument");
  l1_f32_bwd_kernel<<<lgrid(numel, 256), 256, 0, ST(stream)>>>(a, b, numel, 1.f / numel, gscale, grad_a);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_bf16_fwd(const void* a, const void* b, int64_t numel, float* loss, void* stream) {
  MSIG_REQUIRE(a && b && loss && numel > 0 && numel % 8 == 0, "msig_l1_loss_bf16_fwd: bad argument");
  MSIG_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  l1_bf16_fwd_kernel<<<lgrid(numel / 8, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), numel / 8, 1.f / numel, loss);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_l1_loss_bf16_bwd(const void* a, const void* b, int64_t numel, const float* gscale, const void* aux,
                          void* grad_a, void* stream) {
  MSIG_REQUIRE(a && b && grad_a && numel > 0 && numel % 8 == 0, "msig_l1_loss_bf16_bwd: bad argument");
  l1_bf16_bwd_kernel<<<lgrid(numel / 8, 256), 256, 0, ST(stream)>>>(CBF(a), CBF(b), numel / 8, 1.f / numel, gscale,
                                                                   CBF(aux), BF(grad_a));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_mse_const_fwd(const float* a, float target, int64_t numel, float* loss, void* stream) {
  MSIG_REQUIRE(a && loss && numel > 0, "msig_mse_const_fwd: bad argument");
  MSIG_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  mse_const_fwd_kernel<<<lgrid(numel, 256), 256, 0, ST(stream)>>>(a, target, numel, 1.f / numel, loss);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_mse_const_bwd(const float* a, float target, int64_t numel, const float* gscale, float* grad_a,
                       void* stream) {
  MSIG_REQUIRE(a && grad_a && numel > 0, "msig_mse_const_bwd: bad argument");
  mse_const_bwd_kernel<<<lgrid(numel, 256), 256, 0, ST(stream)>>>(a, target, numel, 1.f / numel, gscale, grad_a);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_gram_l1(const float* ga, const float* gb, int32_t dim, float* loss, int accumulate, void* ssym,
                 void* stream) {
  MSIG_REQUIRE(ga && gb && loss && ssym && dim > 0, "msig_gram_l1: bad argument");
  if (!accumulate) MSIG_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), ST(stream)));
  const unsigned t = static_cast<unsigned>(ceil_div(dim, 32));
  gram_l1_kernel<<<dim3(t, t), dim3(32, 8), 0, ST(stream)>>>(ga, gb, dim, 1.f / (float(dim) * float(dim)), loss,
                                                             BF(ssym));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_colsum_f32(const float* x, int64_t rows, int32_t c, int64_t ld, float* out, int accumulate,
                    void* stream) {
  MSIG_REQUIRE(x && out && c > 0, "msig_colsum_f32: bad argument");
  colsum_f32_kernel<<<static_cast<unsigned>(ceil_div(c, 128)), 128, 0, ST(stream)>>>(x, rows, c, ld, out, accumulate);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_sumsq(const float* x, int64_t numel, float* out, int accumulate, void* stream) {
  MSIG_REQUIRE(x && out && numel > 0, "msig_sumsq: bad argument");
  if (!accumulate) MSIG_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ST(stream)));
  sumsq_kernel<<<lgrid(numel / 4 + 1, 256), 256, 0, ST(stream)>>>(x, numel, out);
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
  adam_step_kernel<<<lgrid(numel, 256), 256, 0, ST(stream)>>>(param, grad, exp_avg, exp_avg_sq, ema, numel,
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
  counter_inc_kernel<<<1, 1, 0, ST(stream)>>>(step_counter);
  MSIG_CHECK_LAUNCH();
  adam_step_kernel<<<lgrid(numel, 256), 256, 0, ST(stream)>>>(param, grad, exp_avg, exp_avg_sq, ema, numel,
                                                             grad_sumsq, max_norm, grad_scale, lr, beta1, beta2, eps,
                                                             1.f, 1.f, ema_beta, step_counter);
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

}  // extern "C"
