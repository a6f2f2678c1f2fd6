// Bandwidth-bound kernels: gathered patches for the narrow convs, reflect padding, InstanceNorm /
// AdaIN statistics + fused apply (+activation, +residual) and their backward, pooling, head
// selection, dtype converters. All activations are bf16 NHWC, 16-byte vector accesses, one thread
// per 8 channels, warp-shuffle / shared-memory reductions, fp32 statistics.
//
// Reference call sites: model.py:16,20-36 (AdaIN), :53-55 (ReLU, residual), :131-133,139-140,167
// (InstanceNorm2d + ReLU/LeakyReLU), :76 (AdaptiveAvgPool2d), :112-116,208-212 (head selection),
// :131,141 (reflect padding); losses.py:49-56 (VGG renorm), pool_2/pool_4 (MaxPool2d).
#include "common.h"

#include <algorithm>

namespace msig {

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == MSIG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == MSIG_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}
__device__ __forceinline__ float act_grad(float u, int act, float slope) {
  if (act == MSIG_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == MSIG_ACT_LRELU) return u > 0.f ? 1.f : slope;
  return 1.f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

static inline int grid_for(int64_t work, int threads, int cap = 148 * 16) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, threads), cap)));
}

// ---------------------------------------------------------------- gathered patches
// One block walks 64-output-pixel tiles of one output row. Per tile the input window
// (r rows x ((64-1)*stride + s) columns x c channels, padding already resolved) is staged in shared
// memory; the k -> (channel, r, s) decode is a shared-memory table of window offsets built once per
// block, so the hot loop is two LDS + a convert per element and 16-byte stores contiguous along k.
constexpr int kGatherTile = 64;
__global__ void __launch_bounds__(256) patch_gather_kernel(msig_patch_geom g, const float* __restrict__ src,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           __nv_bfloat16* __restrict__ out, int tiles_w,
                                                           int64_t total_tiles, int win_cols) {
  pdl_entry();
  extern __shared__ uint32_t gsm[];
  const int kgs = g.kpad / 8;
  uint32_t* tab = gsm;                                           // [8][kgs] window offsets (transposed)
  float* win = reinterpret_cast<float*>(gsm + g.kpad);            // [c][r][win_cols]
  const int kvalid = g.r * g.s * g.c;
  for (int k = threadIdx.x; k < g.kpad; k += blockDim.x) {
    uint32_t e = 0xFFFFFFFFu;
    if (k < kvalid) {
      const int t = k / g.c, ch = k - t * g.c;
      const int r = t / g.s, s2 = t - r * g.s;
      e = uint32_t((ch * g.r + r) * win_cols + s2);
    }
    tab[(k & 7) * kgs + (k >> 3)] = e;
  }
  const int items = kGatherTile * kgs;
  const int win_elems = g.c * g.r * win_cols;
  const int64_t plane = int64_t(g.h) * g.w;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int tw = static_cast<int>(tile % tiles_w);
    const int64_t rem = tile / tiles_w;
    const int oh = static_cast<int>(rem % g.oh);
    const int n = static_cast<int>(rem / g.oh);
    const float* sn = src + int64_t(n) * g.c * plane;
    const int ih0 = oh * g.stride - g.pad_t;
    const int iw0 = tw * kGatherTile * g.stride - g.pad_l;
    __syncthreads();                                              // previous tile's readers are done
    for (int i = threadIdx.x; i < win_elems; i += blockDim.x) {
      const int col = i % win_cols;
      const int cr = i / win_cols;
      const int r = cr % g.r, ch = cr / g.r;
      int ih = ih0 + r, iw = iw0 + col;
      bool ok = true;
      if (g.reflect) {
        ih = reflect_idx(ih, g.h);
        iw = reflect_idx(iw, g.w);
        ok = (iw >= 0) && (iw < g.w) && (ih >= 0) && (ih < g.h);   // columns past the last tile's window
      } else {
        ok = (ih >= 0) && (ih < g.h) && (iw >= 0) && (iw < g.w);
      }
      float v = 0.f;
      if (ok) {
        v = __ldg(sn + ch * plane + int64_t(ih) * g.w + iw);
        if (scale != nullptr) v = v * __ldg(scale + ch) + __ldg(shift + ch);
      }
      win[i] = v;
    }
    __syncthreads();
    __nv_bfloat16* orow = out + (int64_t(n) * g.oh + oh) * g.ow * g.kpad;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int p = it / kgs;
      const int kg = it - p * kgs;
      const int ow = tw * kGatherTile + p;
      if (ow >= g.ow) continue;
      const int pbase = p * g.stride;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t e = tab[j * kgs + kg];
        f[j] = (e != 0xFFFFFFFFu) ? win[e + pbase] : 0.f;
      }
      store8(orow + int64_t(ow) * g.kpad + kg * 8, f);
    }
  }
}

// Adjoint of the gather: one thread per source element.
__global__ void patch_scatter_kernel(msig_patch_geom g, const __nv_bfloat16* __restrict__ dp,
                                     const float* __restrict__ scale, float* __restrict__ dsrc,
                                     int accumulate, int64_t total) {
  pdl_entry();
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int iw = static_cast<int>(idx % g.w);
    int64_t rem = idx / g.w;
    const int ih = static_cast<int>(rem % g.h);
    rem /= g.h;
    const int ch = static_cast<int>(rem % g.c);
    const int n = static_cast<int>(rem / g.c);
    // padded-domain positions that read this element
    int ph[3], pw[3];
    int nph = 0, npw = 0;
    ph[nph++] = ih;
    pw[npw++] = iw;
    if (g.reflect) {
      if (ih >= 1 && ih <= g.pad_t) ph[nph++] = -ih;
      const int hi_h = 2 * (g.h - 1) - ih;   // mirrored position beyond the bottom edge
      if (ih <= g.h - 2 && hi_h - (g.h - 1) <= (g.oh - 1) * g.stride + g.r - 1 - g.pad_t - (g.h - 1) &&
          hi_h >= g.h)
        ph[nph++] = hi_h;
      if (iw >= 1 && iw <= g.pad_l) pw[npw++] = -iw;
      const int hi_w = 2 * (g.w - 1) - iw;
      if (iw <= g.w - 2 && hi_w - (g.w - 1) <= (g.ow - 1) * g.stride + g.s - 1 - g.pad_l - (g.w - 1) &&
          hi_w >= g.w)
        pw[npw++] = hi_w;
    }
    float acc = 0.f;
    for (int a = 0; a < nph; ++a)
      for (int r = 0; r < g.r; ++r) {
        const int num_h = ph[a] + g.pad_t - r;
        if (num_h < 0 || (num_h % g.stride) != 0) continue;
        const int oh = num_h / g.stride;
        if (oh >= g.oh) continue;
        for (int b = 0; b < npw; ++b)
          for (int s = 0; s < g.s; ++s) {
            const int num_w = pw[b] + g.pad_l - s;
            if (num_w < 0 || (num_w % g.stride) != 0) continue;
            const int ow = num_w / g.stride;
            if (ow >= g.ow) continue;
            const int64_t m = (int64_t(n) * g.oh + oh) * g.ow + ow;
            acc += __bfloat162float(dp[m * g.kpad + (r * g.s + s) * g.c + ch]);
          }
      }
    if (scale != nullptr) acc *= scale[ch];
    dsrc[idx] = accumulate ? dsrc[idx] + acc : acc;
  }
}

// ---------------------------------------------------------------- padded image copies
// fp32 NCHW [n,c,h,w] (c <= 8) -> bf16 [n][h+2p][w+2p+2][8], reflect or zero padding; channels >= c and
// the two slack columns are zero. One thread per padded pixel (one 16-byte store).
constexpr int kPad8Slack = 8;
__global__ void img_pad8_kernel(const float* __restrict__ src, int c, int h, int w, int pad, int reflect,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                __nv_bfloat16* __restrict__ dst, int64_t total) {
  pdl_entry();
  const int Hp = h + 2 * pad, Wp = w + 2 * pad + 2;
  const int64_t plane = int64_t(h) * w;
  // `total` pixels plus kPad8Slack zero pixels behind the last row: the 8-pixel window of the last output
  // columns of the last padded row reaches up to 7 - s pixels past the row (see msig_img_pad8 in msig.h)
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total + kPad8Slack;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(idx % Wp);
    int64_t rem = idx / Wp;
    const int py = static_cast<int>(rem % Hp);
    const int64_t img = rem / Hp;
    int ih = py - pad, iw = px - pad;
    bool ok = px < w + 2 * pad && idx < total;
    if (reflect) {
      ih = reflect_idx(ih, h);
      iw = reflect_idx(iw, w);
    }
    ok = ok && ih >= 0 && ih < h && iw >= 0 && iw < w;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) {
      const float* sp = src + img * c * plane + int64_t(ih) * w + iw;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        if (ch < c) {
          f[ch] = __ldg(sp + ch * plane);
          if (scale != nullptr) f[ch] = f[ch] * __ldg(scale + ch) + __ldg(shift + ch);
        }
    }
    store8(dst + idx * 8, f);
  }
}

__global__ void reflect_fold_nchw_kernel(const float* __restrict__ dy, int h, int w, int pad,
                                         float* __restrict__ dx, int64_t total) {
  pdl_entry();
  const int H2 = h + 2 * pad, W2 = w + 2 * pad;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int iw = static_cast<int>(idx % w);
    int64_t rem = idx / w;
    const int ih = static_cast<int>(rem % h);
    const int64_t plane = rem / h;                 // (n, c)
    int ph[3], pw[3], nph = 0, npw = 0;
    ph[nph++] = ih + pad;
    pw[npw++] = iw + pad;
    if (ih >= 1 && ih <= pad) ph[nph++] = pad - ih;
    if (ih <= h - 2 && ih >= h - 1 - pad) ph[nph++] = pad + 2 * (h - 1) - ih;
    if (iw >= 1 && iw <= pad) pw[npw++] = pad - iw;
    if (iw <= w - 2 && iw >= w - 1 - pad) pw[npw++] = pad + 2 * (w - 1) - iw;
    const float* src = dy + plane * H2 * W2;
    float acc = 0.f;
    for (int a = 0; a < nph; ++a)
      for (int b = 0; b < npw; ++b) acc += __ldg(src + int64_t(ph[a]) * W2 + pw[b]);
    dx[idx] = acc;
  }
}

// ---------------------------------------------------------------- per-(n,c) reductions
// Block = 256 threads = (c/8) channel groups x (2048/c) pixel lanes; grid = (chunks, n). Every thread
// keeps kUnroll independent 16-byte loads in flight per operand (HBM latency x bandwidth needs
// ~36 KB in flight per SM). The last block of an image to finish (ticket counter in the workspace,
// reset by that block) folds the per-chunk partials into the per-(n,c) results, in chunk order, so
// the result is deterministic and no separate finalize launch is needed.
// MODE 0: sum x, sum x^2 -> mean/rstd/scale/shift (statistics)
// MODE 1: sum g, sum g*xhat -> coef, dgamma, dbeta (norm backward), g = dy*act'(u)
constexpr int kUnroll = 4;

__device__ __forceinline__ uint4 ldg_stream(const __nv_bfloat16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

struct NcFinal {
  // MODE 0
  float eps;
  const float* gamma;
  const float* beta;
  int64_t gb_stride;
  float* mean_out;
  float* rstd_out;
  float* scale_out;
  float* shift_out;
  // MODE 1
  float* coef;
  float* dgamma;
  float* dbeta;
  int64_t dgb_stride;
  int accumulate;
};

// Reflect padding fused into the norm kernels of the layer in front of the generator's final 7x7 conv
// (model.py:140-141): the forward apply writes its output straight into the reflect-padded buffer
// (interior + mirrored border copies), and the backward reads its dy THROUGH the fold (the sum of the
// padded gradient over the positions that mirror onto a pixel). W == 0: plain, unpadded tensors.
struct PadGeom {
  int W, H, pad;
  float inv_w;      // 1 / W: row of pixel p = int((p + 0.5) * inv_w), exact for p < 2^22
};
__device__ __forceinline__ void pixel_hw(const PadGeom& pg, int p, int& h, int& w) {
  h = __float2int_rz((static_cast<float>(p) + 0.5f) * pg.inv_w);
  w = p - h * pg.W;
}
__device__ __forceinline__ bool pad_interior(const PadGeom& pg, int h, int w) {   // no mirrored copy
  return (h > pg.pad) & (h < pg.H - 1 - pg.pad) & (w > pg.pad) & (w < pg.W - 1 - pg.pad);
}
__device__ __forceinline__ void mirror_positions(int i, int n, int pad, int (&pos)[3], int& cnt) {
  cnt = 0;
  pos[cnt++] = i + pad;
  if (i >= 1 && i <= pad) pos[cnt++] = pad - i;
  if (i <= n - 2 && i >= n - 1 - pad) pos[cnt++] = pad + 2 * (n - 1) - i;
}
// dy of pixel p = sum of the padded gradient over its mirror positions. The position (h+pad, w+pad) is
// always one of them: its load is issued with the batched loads (fold_direct_offset); the mirrored
// border copies (a few % of the pixels) are added afterwards (fold_add_mirrors).
__device__ __forceinline__ int64_t fold_direct_offset(const PadGeom& pg, int img, int h, int w, int c, int tx) {
  const int W2 = pg.W + 2 * pg.pad, H2 = pg.H + 2 * pg.pad;
  return ((int64_t(img) * H2 + h + pg.pad) * W2 + w + pg.pad) * c + tx * 8;
}
__device__ __forceinline__ void fold_add_mirrors(const __nv_bfloat16* __restrict__ dyp, int img, int h, int w,
                                                 const PadGeom& pg, int c, int tx, float (&f)[8]) {
  if (pad_interior(pg, h, w)) return;
  const int W2 = pg.W + 2 * pg.pad, H2 = pg.H + 2 * pg.pad;
  int ph[3], pw[3], nph, npw;
  mirror_positions(h, pg.H, pg.pad, ph, nph);
  mirror_positions(w, pg.W, pg.pad, pw, npw);
  for (int a = 0; a < nph; ++a)
    for (int b = 0; b < npw; ++b) {
      if (a == 0 && b == 0) continue;               // the direct position is already in f
      float t[8];
      load8(dyp + ((int64_t(img) * H2 + ph[a]) * W2 + pw[b]) * c + tx * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += t[j];
    }
}

template <int MODE, bool FOLD = false>
__global__ void __launch_bounds__(256, FOLD ? 2 : 1) nc_reduce_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
    const float* __restrict__ shift, int act, float slope, int hw, int c, int pix_per_block,
    float* __restrict__ partial, unsigned int* __restrict__ tickets, NcFinal fin, PadGeom pg = PadGeom{0, 0, 0, 0.f}) {
  pdl_entry();
  __shared__ float red[16][256 + 1];
  __shared__ bool is_last;
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int chunks = gridDim.x;
  const int p0 = chunk * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  float mu[8], rs[8], sc[8], sh[8];
  if (MODE == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = img * c + tx * 8 + j;
      mu[j] = mean[ch]; rs[j] = rstd[ch]; sc[j] = scale[ch]; sh[j] = shift[ch];
    }
  }
  const int64_t base = int64_t(img) * hw * c + tx * 8;
  for (int p = p0 + ty; p < p1; p += kUnroll * lanes) {
    uint4 xv[kUnroll], dv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        xv[u] = ldg_stream(x + base + int64_t(pp) * c);
        if (MODE == 1) {
          if constexpr (FOLD) {                          // dy lives in the reflect-padded buffer: (h, w) per pixel
            int h, w;
            pixel_hw(pg, pp, h, w);
            dv[u] = ldg_stream(dy + fold_direct_offset(pg, img, h, w, c, tx));
          } else {
            dv[u] = ldg_stream(dy + base + int64_t(pp) * c);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (p + u * lanes < p1) {
        float xf[8];
        unpack8(xv[u], xf);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a[j] += xf[j];
            b[j] += xf[j] * xf[j];
          }
        } else {
          float df[8];
          unpack8(dv[u], df);
          if constexpr (FOLD) {
            int h, w;
            pixel_hw(pg, p + u * lanes, h, w);
            fold_add_mirrors(dy, img, h, w, pg, c, tx, df);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float uu = xf[j] * sc[j] + sh[j];
            const float gq = df[j] * act_grad(uu, act, slope);
            a[j] += gq;
            b[j] += gq * (xf[j] - mu[j]) * rs[j];
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[j][threadIdx.x] = a[j];
    red[8 + j][threadIdx.x] = b[j];
  }
  __syncthreads();
  // thread t < 2*c : quantity q = t / c (0: a, 1: b), channel ch = t % c
  for (int t = threadIdx.x; t < 2 * c; t += 256) {
    const int q = t / c, ch = t % c;
    const int gx = ch / 8, j = ch % 8;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[q * 8 + j][l * cg + gx];
    partial[((int64_t(img) * chunks + chunk) * 2 + q) * c + ch] = s;
  }
  // ---- last block of this image finalizes
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[img], 1u);
    is_last = (t == static_cast<unsigned int>(chunks) - 1u);
    if (is_last) tickets[img] = 0u;       // self-cleaning for the next call on this workspace
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    double s1 = 0.0, s2 = 0.0;
    const float* pp = partial + int64_t(img) * chunks * 2 * c + ch;
    for (int k = 0; k < chunks; ++k) {
      s1 += __ldcg(pp + (int64_t(k) * 2 + 0) * c);
      s2 += __ldcg(pp + (int64_t(k) * 2 + 1) * c);
    }
    if (MODE == 0) {
      const double m = s1 / hw;
      double var = s2 / hw - m * m;
      if (var < 0.0) var = 0.0;
      const float r = static_cast<float>(1.0 / sqrt(var + double(fin.eps)));
      const float gm = fin.gamma ? fin.gamma[img * fin.gb_stride + ch] : 1.f;
      const float bt = fin.beta ? fin.beta[img * fin.gb_stride + ch] : 0.f;
      const int o = img * c + ch;
      fin.mean_out[o] = static_cast<float>(m);
      fin.rstd_out[o] = r;
      fin.scale_out[o] = gm * r;
      fin.shift_out[o] = bt - static_cast<float>(m) * gm * r;
    } else {
      fin.coef[(int64_t(img) * 2 + 0) * c + ch] = static_cast<float>(s1 / hw);
      fin.coef[(int64_t(img) * 2 + 1) * c + ch] = static_cast<float>(s2 / hw);
      if (fin.dgamma != nullptr) {
        const int64_t o = img * fin.dgb_stride + ch;
        fin.dgamma[o] = (fin.accumulate ? fin.dgamma[o] : 0.f) + static_cast<float>(s2);
        fin.dbeta[o] = (fin.accumulate ? fin.dbeta[o] : 0.f) + static_cast<float>(s1);
      }
    }
  }
}

// Folds the partial sums written by the implicit-GEMM epilogue (FpropParams::stat_out:
// [img*rows + r][2][ld]) into the per-(image, channel) results. Block = 8 row lanes x 32 channels.
// MODE 0: sum v, sum v^2 -> mean/rstd/scale/shift.  MODE 1: sum g, sum g*z -> coef, dgamma, dbeta.
template <int MODE>
__global__ void __launch_bounds__(1024) epi_stats_finalize_kernel(const float* __restrict__ partial, int rows,
                                                                  int ld, int hw, int c,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, NcFinal fin) {
  pdl_entry();
  __shared__ double red[2][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 channels x 32 row lanes
  const int img = blockIdx.y;
  const int ch = blockIdx.x * 32 + tx;
  double s1 = 0.0, s2 = 0.0;
  if (ch < c) {
    const float* pp = partial + int64_t(img) * rows * 2 * ld + ch;
    for (int r0 = ty; r0 < rows; r0 += 32 * 8) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {            // 16 independent loads in flight per thread: unconditional
        const int r = min(r0 + 32 * u, rows - 1);   // (clamped) -- guarded loads compile to one branch each
        a[u] = __ldg(pp + int64_t(r) * 2 * ld);     // with the use right behind it, i.e. serialised
        b[u] = __ldg(pp + int64_t(r) * 2 * ld + ld);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool live = r0 + 32 * u < rows;
        s1 += live ? double(a[u]) : 0.0;
        s2 += live ? double(b[u]) : 0.0;
      }
    }
  }
  red[0][ty][tx] = s1;
  red[1][ty][tx] = s2;
  __syncthreads();
  if (ty != 0 || ch >= c) return;
  for (int l = 1; l < 32; ++l) {
    s1 += red[0][l][tx];
    s2 += red[1][l][tx];
  }
  if (MODE == 0) {
    const double m = s1 / hw;
    double var = s2 / hw - m * m;
    if (var < 0.0) var = 0.0;
    const float r = static_cast<float>(1.0 / sqrt(var + double(fin.eps)));
    const float gm = fin.gamma ? fin.gamma[img * fin.gb_stride + ch] : 1.f;
    const float bt = fin.beta ? fin.beta[img * fin.gb_stride + ch] : 0.f;
    const int o = img * c + ch;
    fin.mean_out[o] = static_cast<float>(m);
    fin.rstd_out[o] = r;
    fin.scale_out[o] = gm * r;
    fin.shift_out[o] = bt - static_cast<float>(m) * gm * r;
  } else {
    const int o = img * c + ch;
    const double sgx = double(rstd[o]) * (s2 - double(mean[o]) * s1);   // sum g*xhat
    fin.coef[(int64_t(img) * 2 + 0) * c + ch] = static_cast<float>(s1 / hw);
    fin.coef[(int64_t(img) * 2 + 1) * c + ch] = static_cast<float>(sgx / hw);
    if (fin.dgamma != nullptr) {
      const int64_t oo = img * fin.dgb_stride + ch;
      fin.dgamma[oo] = (fin.accumulate ? fin.dgamma[oo] : 0.f) + static_cast<float>(sgx);
      fin.dbeta[oo] = (fin.accumulate ? fin.dbeta[oo] : 0.f) + static_cast<float>(s1);
    }
  }
}

// y = act(x*scale + shift) (+ residual); PADOUT: y is the reflect-padded buffer [n][H+2p][W+2p][c]
template <bool PADOUT = false, int U = kUnroll>
__global__ void __launch_bounds__(256) norm_act_fwd_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
    const __nv_bfloat16* __restrict__ res, int act, float slope, int hw, int c, int pix_per_block,
    __nv_bfloat16* __restrict__ y, PadGeom pg = PadGeom{0, 0, 0, 0.f}) {
  pdl_entry();
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[img * c + tx * 8 + j];
    sh[j] = shift[img * c + tx * 8 + j];
  }
  const int64_t base = int64_t(img) * hw * c + tx * 8;
  for (int p = p0 + ty; p < p1; p += U * lanes) {
    uint4 xv[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        xv[u] = ldg_stream(x + base + int64_t(pp) * c);
        if (res != nullptr) rv[u] = ldg_stream(res + base + int64_t(pp) * c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        float f[8];
        unpack8(xv[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd(f[j] * sc[j] + sh[j], act, slope);
        if (res != nullptr) {
          float rf[8];
          unpack8(rv[u], rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += rf[j];
        }
        if constexpr (PADOUT) {
          int h, w;
          pixel_hw(pg, pp, h, w);
          const int W2 = pg.W + 2 * pg.pad, H2 = pg.H + 2 * pg.pad;
          if (pad_interior(pg, h, w)) {
            store8(y + ((int64_t(img) * H2 + h + pg.pad) * W2 + w + pg.pad) * c + tx * 8, f);
          } else {
            int ph[3], pw[3], nph, npw;
            mirror_positions(h, pg.H, pg.pad, ph, nph);
            mirror_positions(w, pg.W, pg.pad, pw, npw);
            for (int a = 0; a < nph; ++a)
              for (int b = 0; b < npw; ++b)
                store8(y + ((int64_t(img) * H2 + ph[a]) * W2 + pw[b]) * c + tx * 8, f);
          }
        } else {
          store8(y + base + int64_t(pp) * c, f);
        }
      }
    }
  }
}

// dx = scale * (g - c1 - xhat*c2),  g = dy*act'(x*scale+shift); FOLD: dy is read through the reflect fold
template <bool FOLD = false, int U = kUnroll>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
    const float* __restrict__ shift, const float* __restrict__ coef, int act, float slope, int hw, int c,
    int pix_per_block, __nv_bfloat16* __restrict__ dx, PadGeom pg = PadGeom{0, 0, 0, 0.f}) {
  pdl_entry();
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  // dx = k0*g + k1*x + k2 with k0 = scale, k1 = -scale*rstd*c2, k2 = -scale*(c1 - mean*rstd*c2)
  float sc[8], sh[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = img * c + tx * 8 + j;
    const float mu = mean[ch], rs = rstd[ch];
    sc[j] = scale[ch]; sh[j] = shift[ch];
    const float c1 = coef[(int64_t(img) * 2 + 0) * c + tx * 8 + j];
    const float c2 = coef[(int64_t(img) * 2 + 1) * c + tx * 8 + j];
    k1[j] = -sc[j] * rs * c2;
    k2[j] = -sc[j] * (c1 - mu * rs * c2);
  }
  const int64_t base = int64_t(img) * hw * c + tx * 8;
  for (int p = p0 + ty; p < p1; p += U * lanes) {
    uint4 xv[U], dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        xv[u] = ldg_stream(x + base + int64_t(pp) * c);
        if constexpr (FOLD) {
          int h, w;
          pixel_hw(pg, pp, h, w);
          dv[u] = ldg_stream(dy + fold_direct_offset(pg, img, h, w, c, tx));
        } else {
          dv[u] = ldg_stream(dy + base + int64_t(pp) * c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        float xf[8], df[8];
        unpack8(xv[u], xf);
        unpack8(dv[u], df);
        if constexpr (FOLD) {
          int h, w;
          pixel_hw(pg, pp, h, w);
          fold_add_mirrors(dy, img, h, w, pg, c, tx, df);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float uu = xf[j] * sc[j] + sh[j];
          const float gq = df[j] * act_grad(uu, act, slope);
          df[j] = sc[j] * gq + (k1[j] * xf[j] + k2[j]);
        }
        store8(dx + base + int64_t(pp) * c, df);
      }
    }
  }
}

// ---- norm apply kernels that finish the epilogue partial sums themselves (no finalize launch) ----
// For the 64x64 planes of the residual trunk an image has 32 partial rows, so every block can afford to fold
// the [rows][2][c] partials of its image in its prologue (thread = channel, 32 loads in flight, double
// accumulation in row order, so every block of an image derives bit-identical coefficients); block 0 of the
// image also writes the per-(n,c) results the backward pass needs. The first batch of activation loads is
// issued ahead of the fold so HBM latency overlaps it. Replaces epi_stats_finalize_kernel + the plain apply
// (215 launches of ~7 us per training step).
__device__ __forceinline__ void fold_partials(const float* __restrict__ pp, int rows, int ld, int r_first, int r_step,
                                              double& s1, double& s2) {
  s1 = 0.0;
  s2 = 0.0;
  for (int r0 = r_first; r0 < rows; r0 += 16 * r_step) {
    float a[16], b[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int r = min(r0 + u * r_step, rows - 1);   // clamped, unconditional: see epi_stats_finalize_kernel
      a[u] = __ldg(pp + int64_t(r) * 2 * ld);
      b[u] = __ldg(pp + int64_t(r) * 2 * ld + ld);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const bool live = r0 + u * r_step < rows;
      s1 += live ? double(a[u]) : 0.0;
      s2 += live ? double(b[u]) : 0.0;
    }
  }
}

// Per-(image, channel) totals of the partial rows for a 256-thread block: 256 / c row lanes per channel (1 at 256
// channels, 4 at 64), folded in lane order through shared memory, so the result does not depend on the block.
// Calls fin(ch, s1, s2) on the thread that owns channel ch. All threads of the block must call it.
template <typename F>
__device__ __forceinline__ void fold_image_partials(const float* __restrict__ img_partial, int rows, int ld, int c,
                                                    double (&s_p)[2][256], F fin) {
  for (int c0 = 0; c0 < c; c0 += 256) {
    const int cw = min(256, c - c0);
    const int lanes = 256 / cw;
    const int cl = threadIdx.x % cw, ln = threadIdx.x / cw;
    double a = 0.0, b = 0.0;
    if (ln < lanes) fold_partials(img_partial + c0 + cl, rows, ld, ln, lanes, a, b);
    s_p[0][threadIdx.x] = a;
    s_p[1][threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x < cw) {
      double s1 = 0.0, s2 = 0.0;
      for (int l = 0; l < lanes; ++l) {
        s1 += s_p[0][l * cw + threadIdx.x];
        s2 += s_p[1][l * cw + threadIdx.x];
      }
      fin(c0 + threadIdx.x, s1, s2);
    }
    __syncthreads();
  }
}

template <int U>
__global__ void __launch_bounds__(256) norm_act_fwd_fin_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ partial, int rows, int ld, NcFinal fin,
    const __nv_bfloat16* __restrict__ res, int act, float slope, int hw, int c, int pix_per_block,
    __nv_bfloat16* __restrict__ y) {
  pdl_entry();
  __shared__ float s_sc[512], s_sh[512];
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const int64_t base = int64_t(img) * hw * c + tx * 8;
  uint4 xv[U], rv[U];
  int p = p0 + ty;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pp = p + u * lanes;
    if (pp < p1) {
      xv[u] = ldg_stream(x + base + int64_t(pp) * c);
      if (res != nullptr) rv[u] = ldg_stream(res + base + int64_t(pp) * c);
    }
  }
  __shared__ double s_p[2][256];
  fold_image_partials(partial + int64_t(img) * rows * 2 * ld, rows, ld, c, s_p, [&](int ch, double s1, double s2) {
    const double m = s1 / hw;
    double var = s2 / hw - m * m;
    if (var < 0.0) var = 0.0;
    const float r = static_cast<float>(1.0 / sqrt(var + double(fin.eps)));
    const float gm = fin.gamma ? fin.gamma[img * fin.gb_stride + ch] : 1.f;
    const float bt = fin.beta ? fin.beta[img * fin.gb_stride + ch] : 0.f;
    const float scv = gm * r, shv = bt - static_cast<float>(m) * gm * r;
    s_sc[ch] = scv;
    s_sh[ch] = shv;
    if (blockIdx.x == 0) {
      const int o = img * c + ch;
      fin.mean_out[o] = static_cast<float>(m);
      fin.rstd_out[o] = r;
      fin.scale_out[o] = scv;
      fin.shift_out[o] = shv;
    }
  });
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[tx * 8 + j];
    sh[j] = s_sh[tx * 8 + j];
  }
  while (p < p1) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        float f[8];
        unpack8(xv[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd(f[j] * sc[j] + sh[j], act, slope);
        if (res != nullptr) {
          float rf[8];
          unpack8(rv[u], rf);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += rf[j];
        }
        store8(y + base + int64_t(pp) * c, f);
      }
    }
    p += U * lanes;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        xv[u] = ldg_stream(x + base + int64_t(pp) * c);
        if (res != nullptr) rv[u] = ldg_stream(res + base + int64_t(pp) * c);
      }
    }
  }
}

// dx = scale*(g - c1 - xhat*c2) with c1, c2 folded from the dgrad epilogue's partials (sum g, sum g*x)
template <int U>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_fin_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ partial,
    int rows, int ld, const float* __restrict__ mean, const float* __restrict__ rstd,
    const float* __restrict__ scale, const float* __restrict__ shift, NcFinal fin, int act, float slope, int hw,
    int c, int pix_per_block, __nv_bfloat16* __restrict__ dx) {
  pdl_entry();
  __shared__ float s_sc[512], s_sh[512], s_k1[512], s_k2[512];
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const int64_t base = int64_t(img) * hw * c + tx * 8;
  uint4 xv[U], dv[U];
  int p = p0 + ty;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int pp = p + u * lanes;
    if (pp < p1) {
      xv[u] = ldg_stream(x + base + int64_t(pp) * c);
      dv[u] = ldg_stream(dy + base + int64_t(pp) * c);
    }
  }
  __shared__ double s_p[2][256];
  fold_image_partials(partial + int64_t(img) * rows * 2 * ld, rows, ld, c, s_p, [&](int ch, double s1, double s2) {
    const int o = img * c + ch;
    const float mu = mean[o], rs = rstd[o], scv = scale[o];
    const double sgx = double(rs) * (s2 - double(mu) * s1);   // sum g*xhat
    const float c1 = static_cast<float>(s1 / hw), c2 = static_cast<float>(sgx / hw);
    s_sc[ch] = scv;
    s_sh[ch] = shift[o];
    s_k1[ch] = -scv * rs * c2;
    s_k2[ch] = -scv * (c1 - mu * rs * c2);
    if (blockIdx.x == 0) {
      fin.coef[(int64_t(img) * 2 + 0) * c + ch] = c1;
      fin.coef[(int64_t(img) * 2 + 1) * c + ch] = c2;
      if (fin.dgamma != nullptr) {
        const int64_t oo = img * fin.dgb_stride + ch;
        fin.dgamma[oo] = (fin.accumulate ? fin.dgamma[oo] : 0.f) + static_cast<float>(sgx);
        fin.dbeta[oo] = (fin.accumulate ? fin.dbeta[oo] : 0.f) + static_cast<float>(s1);
      }
    }
  });
  float sc[8], sh[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[tx * 8 + j];
    sh[j] = s_sh[tx * 8 + j];
    k1[j] = s_k1[tx * 8 + j];
    k2[j] = s_k2[tx * 8 + j];
  }
  while (p < p1) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        float xf[8], df[8];
        unpack8(xv[u], xf);
        unpack8(dv[u], df);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float uu = xf[j] * sc[j] + sh[j];
          const float gq = df[j] * act_grad(uu, act, slope);
          df[j] = sc[j] * gq + (k1[j] * xf[j] + k2[j]);
        }
        store8(dx + base + int64_t(pp) * c, df);
      }
    }
    p += U * lanes;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = p + u * lanes;
      if (pp < p1) {
        xv[u] = ldg_stream(x + base + int64_t(pp) * c);
        dv[u] = ldg_stream(dy + base + int64_t(pp) * c);
      }
    }
  }
}

// Pixels per block of the per-(n,c) kernels (grid = (chunks, n)). `bps` = resident 256-thread blocks per
// SM of the kernel at hand (register-limited: 4 for the statistics, 3 for the forward apply, 2 for the
// backward kernels): when the batch allows it the grid is ONE full wave of resident blocks, so no
// partially filled last wave trails behind (1024 blocks on 296 slots were 3.46 waves). bps = 0: legacy
// power-of-two chunking (>= 4 blocks per SM, >= 64 pixels per block).
static int pick_pix_per_block(int n, int hw, int bps = 0) {
  if (bps > 0) {
    const int64_t cap = int64_t(sm_count() > 0 ? sm_count() : 148) * bps;
    const int64_t chunks = cap / std::max(n, 1);
    if (chunks >= 1 && ceil_div(hw, chunks) >= 64) return static_cast<int>(ceil_div(hw, chunks));
  }
  int ppb = 256;
  while (ppb > 64 && int64_t(n) * ceil_div(hw, ppb) < 148 * 4) ppb /= 2;
  return std::min(ppb, std::max(hw, 1));
}

// ---------------------------------------------------------------- simple elementwise kernels
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                               int act, float slope, int64_t groups, __nv_bfloat16* __restrict__ dz) {
  pdl_entry();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < groups;
       i += int64_t(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    load8(dy + i * 8, a);
    load8(y + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= act_grad(b[j], act, slope);
    store8(dz + i * 8, a);
  }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ y) {
  pdl_entry();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, float* __restrict__ y) {
  pdl_entry();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n,
                                float* __restrict__ dz) {
  pdl_entry();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    dz[i] = dy[i] * (1.f - y[i] * y[i]);
}

// db[c] (+)= sum_rows dy[rows][c]; block = 256 threads = (c/8 groups) x lanes. Deterministic: every block
// writes its per-channel partial sums, the last block to arrive (integer ticket) adds them in block order
// (no floating-point atomics: an eager step and its CUDA-graph replay give identical bias gradients).
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dy, int64_t rows,
                                                     int c, int rows_per_block, float* __restrict__ db,
                                                     int accumulate, float* __restrict__ partial,
                                                     unsigned int* ticket) {
  pdl_entry();
  __shared__ float red[8][256 + 1];
  __shared__ bool is_last;
  const int cg = c / 8;
  const int lanes = 256 / cg;
  const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
  const int64_t r0 = int64_t(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, rows);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (threadIdx.x < cg * lanes) {
    for (int64_t r = r0 + ty; r < r1; r += lanes) {
      float f[8];
      load8(dy + r * c + tx * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j][threadIdx.x] = a[j];
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    const int gx = ch / 8, j = ch % 8;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[j][l * cg + gx];
    partial[int64_t(blockIdx.x) * c + ch] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1u);
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < c; ch += 256) {
    float s = 0.f;
    for (unsigned int k = 0; k < gridDim.x; ++k) s += __ldcg(partial + int64_t(k) * c + ch);
    db[ch] = (accumulate ? db[ch] : 0.f) + s;
  }
}

// out[c] (+)= sum over (n, hw) of an fp32 NCHW tensor; grid (blocks, c), same two-stage scheme per channel.
__global__ void __launch_bounds__(256) nchw_chansum_kernel(const float* __restrict__ x, int n, int c, int64_t hw,
                                                           int64_t img_stride, float* __restrict__ out,
                                                           int accumulate, float* __restrict__ partial,
                                                           unsigned int* tickets) {
  pdl_entry();
  const int ch = blockIdx.y;
  const int64_t total = int64_t(n) * hw;
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t img = i / hw, p = i - img * hw;
    s += x[img * img_stride + ch * hw + p];
  }
  s = warp_sum(s);
  __shared__ float ws[8];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    partial[int64_t(ch) * gridDim.x + blockIdx.x] = t;
    __threadfence();
    const unsigned int k = atomicAdd(&tickets[ch], 1u);
    is_last = (k == gridDim.x - 1u);
    if (is_last) tickets[ch] = 0u;
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  float t = 0.f;
  for (unsigned int k = 0; k < gridDim.x; ++k) t += __ldcg(partial + int64_t(ch) * gridDim.x + k);
  out[ch] = (accumulate ? out[ch] : 0.f) + t;
}

// ---------------------------------------------------------------- pooling
__global__ void maxpool2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c,
                                    __nv_bfloat16* __restrict__ y, int64_t groups) {
  pdl_entry();
  const int cg = c / 8, oh = h / 2, ow = w / 2;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < groups;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int g8 = static_cast<int>(idx % cg);
    int64_t rem = idx / cg;
    const int x0 = static_cast<int>(rem % ow);
    rem /= ow;
    const int y0 = static_cast<int>(rem % oh);
    const int img = static_cast<int>(rem / oh);
    const __nv_bfloat16* p = x + ((int64_t(img) * h + 2 * y0) * w + 2 * x0) * c + g8 * 8;
    float a[8], b[8];
    load8(p, a);
    load8(p + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], b[j]);
    load8(p + int64_t(w) * c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], b[j]);
    load8(p + int64_t(w) * c + c, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], b[j]);
    store8(y + idx * 8, a);
  }
}
// dx = dy routed to the first max position in window order (torch's tie-break for equal values)
__global__ void maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                    int n, int h, int w, int c, __nv_bfloat16* __restrict__ dx,
                                    int64_t groups) {
  pdl_entry();
  const int cg = c / 8, oh = h / 2, ow = w / 2;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < groups;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int g8 = static_cast<int>(idx % cg);
    int64_t rem = idx / cg;
    const int x0 = static_cast<int>(rem % ow);
    rem /= ow;
    const int y0 = static_cast<int>(rem % oh);
    const int img = static_cast<int>(rem / oh);
    const int64_t off = ((int64_t(img) * h + 2 * y0) * w + 2 * x0) * c + g8 * 8;
    float v[4][8], d[8], o[4][8];
    load8(x + off, v[0]);
    load8(x + off + c, v[1]);
    load8(x + off + int64_t(w) * c, v[2]);
    load8(x + off + int64_t(w) * c + c, v[3]);
    load8(dy + idx * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float bv = v[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][j] > bv) { bv = v[k][j]; best = k; }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k][j] = (k == best) ? d[j] : 0.f;
    }
    store8(dx + off, o[0]);
    store8(dx + off + c, o[1]);
    store8(dx + off + int64_t(w) * c, o[2]);
    store8(dx + off + int64_t(w) * c + c, o[3]);
  }
}

// [n,hw,c] -> [n,c] mean; one block per image, (c/8) x lanes threads
// grid = (n, c / 64): one block per image and 64-channel chunk, 8 channel groups x 32 pixel lanes, every lane's
// loads independent (the first version gave an image ONE block whose threads walked hw / 4 pixels one dependent
// L2 round trip at a time: 20 us for a 4 MB tensor at batch 16)
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int hw, int c,
                                                          __nv_bfloat16* __restrict__ y) {
  pdl_entry();
  __shared__ float red[8][256 + 1];
  const int img = blockIdx.x;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x % 8, ty = threadIdx.x / 8;      // channel group, pixel lane (32)
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const __nv_bfloat16* px = x + int64_t(img) * hw * c + c0 + tx * 8;
  for (int p = ty; p < hw; p += 32 * 4) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pp = min(p + 32 * u, hw - 1);                 // clamped, unconditional loads: all four in flight
      load8(px + int64_t(pp) * c, f[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + 32 * u < hw) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += f[u][j];
      }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j][threadIdx.x] = a[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int gx = threadIdx.x / 8, j = threadIdx.x % 8;
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += red[j][l * 8 + gx];
    y[int64_t(img) * c + c0 + threadIdx.x] = __float2bfloat16(s / hw);
  }
}
__global__ void avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int hw, int c,
                                   __nv_bfloat16* __restrict__ dx, int64_t groups) {
  pdl_entry();
  const int cg = c / 8;
  const float inv = 1.f / hw;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < groups;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int g8 = static_cast<int>(idx % cg);
    const int64_t img = idx / (int64_t(cg) * hw);
    float f[8];
    load8(dy + img * c + g8 * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    store8(dx + idx * 8, f);
  }
}

// ---------------------------------------------------------------- head selection
// head_major = 1: all is [n][heads_ld*per_head], head k occupies [k*per_head, (k+1)*per_head) (SE).
// head_major = 0: all is [n][pix][heads_ld], head k is channel k of every pixel (D, per_head = 1).
// Index semantics of torch advanced indexing (model.py:112-116): a negative index counts from the end;
// an index outside [-heads, heads) is an error there -- here the selected output is poisoned with NaN
// (loud in every downstream loss) instead of reading out of bounds. Host-side callers that hold the
// indices on the CPU raise IndexError before the launch (model.py of this package).
__device__ __forceinline__ int wrap_head(long long k, int heads) {
  if (k < 0) k += heads;
  return (k >= 0 && k < heads) ? static_cast<int>(k) : -1;
}
__global__ void head_gather_kernel(const float* __restrict__ all, const int64_t* __restrict__ idx, int n,
                                   int pix, int heads_ld, int heads, int per_head, int head_major,
                                   float* __restrict__ out) {
  pdl_entry();
  const int64_t total = int64_t(n) * pix * per_head;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % per_head);
    const int64_t rem = i / per_head;
    const int p = static_cast<int>(rem % pix);
    const int b = static_cast<int>(rem / pix);
    const int k = idx ? wrap_head(idx[b], heads) : 0;
    if (k < 0) {
      out[i] = __int_as_float(0x7fc00000);
      continue;
    }
    const int64_t src = head_major ? (int64_t(b) * pix + p) * heads_ld * per_head + int64_t(k) * per_head + e
                                   : (int64_t(b) * pix + p) * heads_ld + k;
    out[i] = all[src];
  }
}
__global__ void head_scatter_kernel(const float* __restrict__ dout, const int64_t* __restrict__ idx, int n,
                                    int pix, int heads_ld, int heads, int per_head, int head_major,
                                    float* __restrict__ dall) {
  pdl_entry();
  const int64_t total = int64_t(n) * pix * heads_ld * per_head;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    int b, p, k, e;
    if (head_major) {
      e = static_cast<int>(i % per_head);
      int64_t rem = i / per_head;
      k = static_cast<int>(rem % heads_ld);
      rem /= heads_ld;
      p = static_cast<int>(rem % pix);
      b = static_cast<int>(rem / pix);
    } else {
      e = 0;
      k = static_cast<int>(i % heads_ld);
      const int64_t rem = i / heads_ld;
      p = static_cast<int>(rem % pix);
      b = static_cast<int>(rem / pix);
    }
    const int sel = idx ? wrap_head(idx[b], heads) : 0;
    dall[i] = (k == sel) ? dout[(int64_t(b) * pix + p) * per_head + e] : 0.f;
  }
}

}  // namespace msig

using namespace msig;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

// 16-byte loads in flight per operand per thread in the plain norm apply kernels: measured in situ at
// [32,64,64,256] (profiles/probe/norm_unroll_r2.txt): forward 4 (27.7 us; 6: 30.7), backward 6 (46.1 us; 4: 50.0).
template <typename... Args>
static void launch_norm_fwd(dim3 grid, cudaStream_t st, Args... args) {
  MSIG_LAUNCH((norm_act_fwd_kernel<false, 4>), grid, 256, 0, st, args...);
}
template <typename... Args>
static void launch_norm_bwd(dim3 grid, cudaStream_t st, Args... args) {
  MSIG_LAUNCH((norm_act_bwd_kernel<false, 6>), grid, 256, 0, st, args...);
}


extern "C" {

int msig_patch_gather(const msig_patch_geom* g, const float* src, const float* scale, const float* shift,
                      void* patches, void* stream) {
  MSIG_REQUIRE(g && src && patches, "msig_patch_gather: null argument");
  MSIG_REQUIRE(g->kpad % 64 == 0 && g->kpad >= g->r * g->s * g->c, "msig_patch_gather: bad kpad %d", g->kpad);
  MSIG_REQUIRE((scale == nullptr) == (shift == nullptr), "msig_patch_gather: scale and shift go together");
  MSIG_REQUIRE(g->c <= 255 && g->r <= 255 && g->s <= 255, "msig_patch_gather: channel / filter extent too large");
  const int tiles_w = static_cast<int>(ceil_div(g->ow, kGatherTile));
  const int64_t tiles = int64_t(g->n) * g->oh * tiles_w;
  const int win_cols = (kGatherTile - 1) * g->stride + g->s;
  const size_t smem = (size_t(g->kpad) + size_t(g->c) * g->r * win_cols) * 4;
  MSIG_REQUIRE(smem <= 48 * 1024, "msig_patch_gather: window too large for shared memory (%zu B)", smem);
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(tiles, 148 * 6)));
  MSIG_LAUNCH((patch_gather_kernel), blocks, 256, smem, ST(stream), *g, src, scale, shift, BF(patches), tiles_w, tiles,
                                                         win_cols);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_patch_scatter(const msig_patch_geom* g, const void* dpatches, const float* scale, float* dsrc,
                       int accumulate, void* stream) {
  MSIG_REQUIRE(g && dpatches && dsrc, "msig_patch_scatter: null argument");
  const int64_t total = int64_t(g->n) * g->c * g->h * g->w;
  MSIG_LAUNCH((patch_scatter_kernel), grid_for(total, 256, 148 * 32), 256, 0, ST(stream), *g, CBF(dpatches), scale, dsrc,
                                                                              accumulate, total);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_img_pad8(const float* src_nchw, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad, int32_t reflect,
                  const float* scale, const float* shift, void* dst, void* stream) {
  MSIG_REQUIRE(src_nchw && dst && c >= 1 && c <= 8 && pad >= 0 && (!reflect || (pad < h && pad < w)),
               "msig_img_pad8: bad argument");
  MSIG_REQUIRE((scale == nullptr) == (shift == nullptr), "msig_img_pad8: scale and shift go together");
  const int64_t total = int64_t(n) * (h + 2 * pad) * (w + 2 * pad + 2);
  MSIG_LAUNCH((img_pad8_kernel), grid_for(total + kPad8Slack, 256, 148 * 32), 256, 0, ST(stream), src_nchw, c, h, w, pad, reflect, scale, shift,
                                                                          BF(dst), total);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_reflect_fold_nchw(const float* dy_padded, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad, float* dx,
                           void* stream) {
  MSIG_REQUIRE(dy_padded && dx && pad < h && pad < w, "msig_reflect_fold_nchw: bad argument");
  const int64_t total = int64_t(n) * c * h * w;
  MSIG_LAUNCH((reflect_fold_nchw_kernel), grid_for(total, 256, 148 * 32), 256, 0, ST(stream), dy_padded, h, w, pad, dx, total);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

static bool norm_c_ok(int c) { return c == 64 || c == 128 || c == 256 || c == 512; }

// workspace: partial [n][chunks][2][c] f32 | coef [n][2][c] f32 | tickets [n] u32
size_t msig_in_stats_workspace(int32_t n, int32_t hw, int32_t c) {
  const int ppb = pick_pix_per_block(n, hw);
  const size_t chunks = static_cast<size_t>(ceil_div(hw, ppb));
  return (size_t(n) * chunks * 2 * c + size_t(n) * 2 * c + size_t(n)) * sizeof(float);
}

int msig_in_stats(const void* x, int32_t n, int32_t hw, int32_t c, float eps, const float* gamma,
                  const float* beta, int64_t gb_stride, float* mean, float* rstd, float* scale, float* shift,
                  void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(x && mean && rstd && scale && shift && workspace, "msig_in_stats: null argument");
  MSIG_REQUIRE(norm_c_ok(c), "msig_in_stats: channels %d unsupported (64/128/256/512)", c);
  MSIG_REQUIRE(workspace_bytes >= msig_in_stats_workspace(n, hw, c), "msig_in_stats: workspace too small");
  const int ppb = pick_pix_per_block(n, hw, 4);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  float* partial = reinterpret_cast<float*>(workspace);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(partial + size_t(n) * chunks * 2 * c + size_t(n) * 2 * c);
  MSIG_CHECK_CUDA(cudaMemsetAsync(tickets, 0, size_t(n) * sizeof(unsigned int), ST(stream)));
  NcFinal fin{};
  fin.eps = eps; fin.gamma = gamma; fin.beta = beta; fin.gb_stride = gb_stride;
  fin.mean_out = mean; fin.rstd_out = rstd; fin.scale_out = scale; fin.shift_out = shift;
  MSIG_LAUNCH((nc_reduce_kernel<0>), dim3(chunks, n), 256, 0, ST(stream), CBF(x), nullptr, nullptr, nullptr, nullptr,
                                                              nullptr, 0, 0.f, hw, c, ppb, partial, tickets, fin, PadGeom{0, 0, 0, 0.f});
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_norm_act_fwd(const void* x, const float* scale, const float* shift, const void* residual, int32_t act,
                      float slope, int32_t n, int32_t hw, int32_t c, void* y, void* stream) {
  MSIG_REQUIRE(x && scale && shift && y, "msig_norm_act_fwd: null argument");
  MSIG_REQUIRE(norm_c_ok(c), "msig_norm_act_fwd: channels %d unsupported", c);
  const int ppb = pick_pix_per_block(n, hw, 3);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  launch_norm_fwd(dim3(chunks, n), ST(stream), CBF(x), scale, shift, CBF(residual), act, slope, hw, c, ppb, BF(y),
                  PadGeom{0, 0, 0, 0.f});
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_norm_act_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* scale,
                      const float* shift, const float* gamma, int64_t gb_stride, int32_t act, float slope,
                      int32_t n, int32_t hw, int32_t c, void* dx, float* dgamma, float* dbeta,
                      int64_t dgb_stride, int accumulate_dgb, void* workspace, size_t workspace_bytes,
                      void* stream) {
  (void)gamma; (void)gb_stride;   // gamma*rstd is already folded into `scale`
  MSIG_REQUIRE(dy && x && mean && rstd && scale && shift && dx && workspace, "msig_norm_act_bwd: null argument");
  MSIG_REQUIRE(norm_c_ok(c), "msig_norm_act_bwd: channels %d unsupported", c);
  MSIG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "msig_norm_act_bwd: dgamma/dbeta go together");
  MSIG_REQUIRE(workspace_bytes >= msig_in_stats_workspace(n, hw, c), "msig_norm_act_bwd: workspace too small");
  const int ppb = pick_pix_per_block(n, hw, 2);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  float* partial = reinterpret_cast<float*>(workspace);
  float* coef = partial + size_t(n) * chunks * 2 * c;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(coef + size_t(n) * 2 * c);
  MSIG_CHECK_CUDA(cudaMemsetAsync(tickets, 0, size_t(n) * sizeof(unsigned int), ST(stream)));
  NcFinal fin{};
  fin.coef = coef; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.dgb_stride = dgb_stride; fin.accumulate = accumulate_dgb;
  MSIG_LAUNCH((nc_reduce_kernel<1>), dim3(chunks, n), 256, 0, ST(stream), CBF(x), CBF(dy), mean, rstd, scale, shift, act,
                                                              slope, hw, c, ppb, partial, tickets, fin, PadGeom{0, 0, 0, 0.f});
  MSIG_CHECK_LAUNCH();
  launch_norm_bwd(dim3(chunks, n), ST(stream), CBF(dy), CBF(x), mean, rstd, scale, shift,
                  static_cast<const float*>(coef), act, slope, hw, c, ppb, BF(dx), PadGeom{0, 0, 0, 0.f});
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

// Pad-fused norm kernels: one full wave of resident blocks (2 per SM) like the plain kernels. The first version
// gave every image row its own 256-thread block (8192 blocks of 2 loop iterations each at [32,256,256,64]):
// forward 142.9 -> 132.4 us, backward (reduce + apply) 403.2 -> 340.9 us (profiles/probe/padnorm_r2.txt).
static int padnorm_ppb(int n, int hw) { return pick_pix_per_block(n, hw, 2); }

int msig_norm_act_fwd_pad(const void* x, const float* scale, const float* shift, int32_t act, float slope, int32_t n,
                          int32_t h, int32_t w, int32_t c, int32_t pad, void* y_padded, void* stream) {
  MSIG_REQUIRE(x && scale && shift && y_padded, "msig_norm_act_fwd_pad: null argument");
  MSIG_REQUIRE(norm_c_ok(c) && pad >= 1 && pad < h && pad < w && int64_t(h) * w < (1 << 22),
               "msig_norm_act_fwd_pad: bad shape");
  const int hw = h * w;
  const int ppb = padnorm_ppb(n, hw);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  MSIG_LAUNCH((norm_act_fwd_kernel<true>), dim3(chunks, n), 256, 0, ST(stream), CBF(x), scale, shift, nullptr, act, slope, hw, c,
                                                                    ppb, BF(y_padded), PadGeom{w, h, pad, 1.f / w});
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

size_t msig_norm_act_bwd_pad_workspace(int32_t n, int32_t h, int32_t w, int32_t c) {
  const int hw = h * w;
  (void)w;
  const size_t chunks = size_t(ceil_div(hw, padnorm_ppb(n, hw)));
  return (size_t(n) * chunks * 2 * c + size_t(n) * 2 * c + size_t(n)) * sizeof(float);
}

int msig_norm_act_bwd_pad(const void* dy_padded, const void* x, const float* mean, const float* rstd,
                          const float* scale, const float* shift, int32_t act, float slope, int32_t n, int32_t h,
                          int32_t w, int32_t c, int32_t pad, void* dx, void* workspace, size_t workspace_bytes,
                          void* stream) {
  MSIG_REQUIRE(dy_padded && x && mean && rstd && scale && shift && dx && workspace,
               "msig_norm_act_bwd_pad: null argument");
  MSIG_REQUIRE(norm_c_ok(c) && pad >= 1 && pad < h && pad < w && int64_t(h) * w < (1 << 22),
               "msig_norm_act_bwd_pad: bad shape");
  const int hw = h * w;
  const int ppb = padnorm_ppb(n, hw);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  MSIG_REQUIRE(workspace_bytes >= (size_t(n) * chunks * 2 * c + size_t(n) * 2 * c + size_t(n)) * sizeof(float),
               "msig_norm_act_bwd_pad: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  float* coef = partial + size_t(n) * chunks * 2 * c;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(coef + size_t(n) * 2 * c);
  MSIG_CHECK_CUDA(cudaMemsetAsync(tickets, 0, size_t(n) * sizeof(unsigned int), ST(stream)));
  NcFinal fin{};
  fin.coef = coef;
  const PadGeom pg{w, h, pad, 1.f / w};
  MSIG_LAUNCH((nc_reduce_kernel<1, true>), dim3(chunks, n), 256, 0, ST(stream), CBF(x), CBF(dy_padded), mean, rstd, scale, shift,
                                                                    act, slope, hw, c, ppb, partial, tickets, fin, pg);
  MSIG_CHECK_LAUNCH();
  MSIG_LAUNCH((norm_act_bwd_kernel<true>), dim3(chunks, n), 256, 0, ST(stream), CBF(dy_padded), CBF(x), mean, rstd, scale, shift,
                                                                    coef, act, slope, hw, c, ppb, BF(dx), pg);
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_in_stats_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld, int32_t hw,
                                int32_t c, float eps, const float* gamma, const float* beta, int64_t gb_stride,
                                float* mean, float* rstd, float* scale, float* shift, void* stream) {
  MSIG_REQUIRE(partial && mean && rstd && scale && shift, "msig_in_stats_from_partials: null argument");
  MSIG_REQUIRE(rows_per_img > 0 && ld >= c && c % 32 == 0, "msig_in_stats_from_partials: bad shape");
  NcFinal fin{};
  fin.eps = eps; fin.gamma = gamma; fin.beta = beta; fin.gb_stride = gb_stride;
  fin.mean_out = mean; fin.rstd_out = rstd; fin.scale_out = scale; fin.shift_out = shift;
  MSIG_LAUNCH((epi_stats_finalize_kernel<0>), dim3(c / 32, n), 1024, 0, ST(stream), partial, rows_per_img, ld, hw, c, nullptr,
                                                                        nullptr, fin);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_norm_bwd_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld, const void* g,
                                const void* x, const float* mean, const float* rstd, const float* scale,
                                const float* shift, int32_t hw, int32_t c, void* dx, float* dgamma, float* dbeta,
                                int64_t dgb_stride, int accumulate_dgb, float* coef, void* stream) {
  MSIG_REQUIRE(partial && g && x && mean && rstd && scale && shift && dx && coef,
               "msig_norm_bwd_from_partials: null argument");
  MSIG_REQUIRE(norm_c_ok(c) && rows_per_img > 0 && ld >= c, "msig_norm_bwd_from_partials: bad shape");
  MSIG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "msig_norm_bwd_from_partials: dgamma/dbeta go together");
  NcFinal fin{};
  fin.coef = coef; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.dgb_stride = dgb_stride; fin.accumulate = accumulate_dgb;
  MSIG_LAUNCH((epi_stats_finalize_kernel<1>), dim3(c / 32, n), 1024, 0, ST(stream), partial, rows_per_img, ld, hw, c, mean, rstd,
                                                                        fin);
  MSIG_CHECK_LAUNCH();
  const int ppb = pick_pix_per_block(n, hw, 2);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  launch_norm_bwd(dim3(chunks, n), ST(stream), CBF(g), CBF(x), mean, rstd, scale, shift,
                  static_cast<const float*>(coef), static_cast<int>(MSIG_ACT_NONE), 0.f, hw, c, ppb, BF(dx),
                  PadGeom{0, 0, 0, 0.f});
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_norm_act_fwd_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld, int32_t hw,
                                    int32_t c, float eps, const float* gamma, const float* beta, int64_t gb_stride,
                                    float* mean, float* rstd, float* scale, float* shift, const void* x,
                                    const void* residual, int32_t act, float slope, void* y, void* stream) {
  MSIG_REQUIRE(partial && mean && rstd && scale && shift && x && y, "msig_norm_act_fwd_from_partials: null argument");
  MSIG_REQUIRE(norm_c_ok(c) && rows_per_img > 0 && ld >= c, "msig_norm_act_fwd_from_partials: bad shape");
  NcFinal fin{};
  fin.eps = eps; fin.gamma = gamma; fin.beta = beta; fin.gb_stride = gb_stride;
  fin.mean_out = mean; fin.rstd_out = rstd; fin.scale_out = scale; fin.shift_out = shift;
  const int ppb = pick_pix_per_block(n, hw, 3);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  MSIG_LAUNCH((norm_act_fwd_fin_kernel<4>), dim3(chunks, n), 256, 0, ST(stream), CBF(x), partial, rows_per_img, ld, fin,
                                                                     CBF(residual), act, slope, hw, c, ppb, BF(y));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_norm_bwd_from_partials_fused(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld, const void* g,
                                      const void* x, const float* mean, const float* rstd, const float* scale,
                                      const float* shift, int32_t hw, int32_t c, void* dx, float* dgamma,
                                      float* dbeta, int64_t dgb_stride, int accumulate_dgb, float* coef,
                                      void* stream) {
  MSIG_REQUIRE(partial && g && x && mean && rstd && scale && shift && dx && coef,
               "msig_norm_bwd_from_partials_fused: null argument");
  MSIG_REQUIRE(norm_c_ok(c) && rows_per_img > 0 && ld >= c, "msig_norm_bwd_from_partials_fused: bad shape");
  MSIG_REQUIRE((dgamma == nullptr) == (dbeta == nullptr),
               "msig_norm_bwd_from_partials_fused: dgamma/dbeta go together");
  NcFinal fin{};
  fin.coef = coef; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.dgb_stride = dgb_stride; fin.accumulate = accumulate_dgb;
  const int ppb = pick_pix_per_block(n, hw, 2);
  const int chunks = static_cast<int>(ceil_div(hw, ppb));
  MSIG_LAUNCH((norm_act_bwd_fin_kernel<6>), dim3(chunks, n), 256, 0, ST(stream), 
      CBF(g), CBF(x), partial, rows_per_img, ld, mean, rstd, scale, shift, fin, static_cast<int>(MSIG_ACT_NONE), 0.f,
      hw, c, ppb, BF(dx));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_act_bwd(const void* dy, const void* y, int32_t act, float slope, int64_t numel, void* dz, void* stream) {
  MSIG_REQUIRE(dy && y && dz && numel % 8 == 0, "msig_act_bwd: bad argument");
  MSIG_LAUNCH((act_bwd_kernel), grid_for(numel / 8, 256), 256, 0, ST(stream), CBF(dy), CBF(y), act, slope, numel / 8, BF(dz));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
static int colsum_rows_per_block(int64_t rows) {
  return static_cast<int>(std::max<int64_t>(512, ceil_div(rows, 148 * 2)));
}
size_t msig_colsum_workspace(int64_t rows, int32_t c) {
  const int64_t blocks = std::max<int64_t>(1, ceil_div(rows, colsum_rows_per_block(rows)));
  return (size_t(blocks) * c + 1) * sizeof(float);
}
int msig_colsum(const void* dy, int64_t rows, int32_t c, float* db, int accumulate, void* workspace,
                size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(dy && db && workspace && c % 8 == 0 && c / 8 <= 256 && 256 % (c / 8) == 0, "msig_colsum: bad argument");
  MSIG_REQUIRE(workspace_bytes >= msig_colsum_workspace(rows, c), "msig_colsum: workspace too small");
  const int rpb = colsum_rows_per_block(rows);
  const int blocks = static_cast<int>(std::max<int64_t>(1, ceil_div(rows, rpb)));
  float* partial = reinterpret_cast<float*>(workspace);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + size_t(blocks) * c);
  MSIG_CHECK_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((colsum_kernel), blocks, 256, 0, ST(stream), CBF(dy), rows, c, rpb, db, accumulate, partial, ticket);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
static int chansum_blocks(int32_t n, int64_t hw) { return grid_for(int64_t(n) * hw, 256, 256); }
size_t msig_nchw_chansum_workspace(int32_t n, int32_t c, int64_t hw) {
  return (size_t(c) * chansum_blocks(n, hw) + size_t(c)) * sizeof(float);
}
int msig_nchw_chansum(const float* x, int32_t n, int32_t c, int64_t hw, int64_t img_stride, float* out,
                      int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(x && out && workspace && c >= 1, "msig_nchw_chansum: null argument");
  MSIG_REQUIRE(workspace_bytes >= msig_nchw_chansum_workspace(n, c, hw), "msig_nchw_chansum: workspace too small");
  const int blocks = chansum_blocks(n, hw);
  float* partial = reinterpret_cast<float*>(workspace);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(partial + size_t(c) * blocks);
  MSIG_CHECK_CUDA(cudaMemsetAsync(tickets, 0, size_t(c) * sizeof(unsigned int), ST(stream)));
  MSIG_LAUNCH((nchw_chansum_kernel), dim3(blocks, c), 256, 0, ST(stream), x, n, c, hw, img_stride, out, accumulate, partial,
                                                              tickets);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_maxpool2_fwd(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, void* stream) {
  MSIG_REQUIRE(x && y && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "msig_maxpool2_fwd: bad argument");
  const int64_t groups = int64_t(n) * (h / 2) * (w / 2) * (c / 8);
  MSIG_LAUNCH((maxpool2_fwd_kernel), grid_for(groups, 256, 148 * 32), 256, 0, ST(stream), CBF(x), n, h, w, c, BF(y), groups);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_maxpool2_bwd(const void* dy, const void* x, const void* y, int32_t n, int32_t h, int32_t w, int32_t c,
                      void* dx, void* stream) {
  (void)y;
  MSIG_REQUIRE(dy && x && dx && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "msig_maxpool2_bwd: bad argument");
  const int64_t groups = int64_t(n) * (h / 2) * (w / 2) * (c / 8);
  MSIG_LAUNCH((maxpool2_bwd_kernel), grid_for(groups, 256, 148 * 32), 256, 0, ST(stream), CBF(dy), CBF(x), n, h, w, c, BF(dx),
                                                                              groups);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_avgpool_fwd(const void* x, int32_t n, int32_t hw, int32_t c, void* y, void* stream) {
  MSIG_REQUIRE(x && y && c % 64 == 0, "msig_avgpool_fwd: bad argument");
  MSIG_LAUNCH((avgpool_fwd_kernel), dim3(n, c / 64), 256, 0, ST(stream), CBF(x), hw, c, BF(y));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_avgpool_bwd(const void* dy, int32_t n, int32_t hw, int32_t c, void* dx, void* stream) {
  MSIG_REQUIRE(dy && dx && c % 8 == 0, "msig_avgpool_bwd: bad argument");
  const int64_t groups = int64_t(n) * hw * (c / 8);
  MSIG_LAUNCH((avgpool_bwd_kernel), grid_for(groups, 256), 256, 0, ST(stream), CBF(dy), hw, c, BF(dx), groups);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_head_gather(const float* all, const int64_t* idx, int32_t n, int32_t pix, int32_t heads_ld,
                     int32_t heads, int32_t per_head, int32_t head_major, float* out, void* stream) {
  MSIG_REQUIRE(all && out && heads >= 1 && heads <= heads_ld, "msig_head_gather: bad argument");
  const int64_t total = int64_t(n) * pix * per_head;
  MSIG_LAUNCH((head_gather_kernel), grid_for(total, 256), 256, 0, ST(stream), all, idx, n, pix, heads_ld, heads, per_head,
                                                                  head_major, out);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_head_scatter(const float* dout, const int64_t* idx, int32_t n, int32_t pix, int32_t heads_ld,
                      int32_t heads, int32_t per_head, int32_t head_major, float* dall, void* stream) {
  MSIG_REQUIRE(dout && dall && heads >= 1 && heads <= heads_ld, "msig_head_scatter: bad argument");
  const int64_t total = int64_t(n) * pix * heads_ld * per_head;
  MSIG_LAUNCH((head_scatter_kernel), grid_for(total, 256), 256, 0, ST(stream), dout, idx, n, pix, heads_ld, heads, per_head,
                                                                   head_major, dall);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_f32_to_bf16(const float* x, int64_t numel, void* y, void* stream) {
  MSIG_REQUIRE(x && y, "msig_f32_to_bf16: null argument");
  MSIG_LAUNCH((f32_to_bf16_kernel), grid_for(numel, 256), 256, 0, ST(stream), x, numel, BF(y));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_bf16_to_f32(const void* x, int64_t numel, float* y, void* stream) {
  MSIG_REQUIRE(x && y, "msig_bf16_to_f32: null argument");
  MSIG_LAUNCH((bf16_to_f32_kernel), grid_for(numel, 256), 256, 0, ST(stream), CBF(x), numel, y);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
int msig_tanh_bwd(const float* dy, const float* y, int64_t numel, float* dz, void* stream) {
  MSIG_REQUIRE(dy && y && dz, "msig_tanh_bwd: null argument");
  MSIG_LAUNCH((tanh_bwd_kernel), grid_for(numel, 256), 256, 0, ST(stream), dy, y, numel, dz);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

}  // extern "C"
