// Host side of the convolution entry points: geometry -> TMA descriptors + tap tables -> the two
// tcgen05 implicit-GEMM kernels; weight packing and the split-K reduce / un-pack of weight grads.
#include "common.h"
#include "igemm.cuh"

#include <algorithm>
#include <cstring>

namespace msig {


// ------------------------------------------------------------------------ phase tables
// k=4, s=2, p=1: output row 2i+py of a transposed conv (or input row of a stride-2 conv's dgrad)
// reads source row i+d through filter row r, for two (r, d) pairs per parity.
__host__ __device__ inline int ph_r(int py, int ty) { return py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 0 : 2); }
__host__ __device__ inline int ph_d(int py, int ty) { return py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0); }
// inverse: filter row r -> (parity, tap-in-phase)
__host__ __device__ inline void r_to_ph(int r, int& py, int& ty) {
  py = (r == 1 || r == 3) ? 0 : 1;
  ty = (r == 1 || r == 0) ? 0 : 1;
}

// ------------------------------------------------------------------------ weight pack / unpack
struct PackGeom {
  int kind, O, I, R, S, OC, OOFF;
  int RS, Ipad, Opad, Kpad;
  int partT;   // wgrad partials stored [I][RS][O] instead of [O][RS][I]
};

static PackGeom make_pack_geom(const msig_wpack_desc* d, int oc, int ooff) {
  PackGeom g;
  g.kind = d->kind; g.O = d->o; g.I = d->i; g.R = d->r; g.S = d->s;
  g.OC = oc > 0 ? oc : d->o;
  g.OOFF = ooff;
  g.partT = 0;
  g.RS = g.R * g.S;
  g.Ipad = pad_rows(g.I);
  g.Opad = pad_rows(g.OC);
  if (g.kind == MSIG_WPACK_IM2COL_FLIP) g.Kpad = static_cast<int>(round_up(int64_t(g.RS) * g.OC, 64));
  else g.Kpad = static_cast<int>(round_up(int64_t(g.RS) * g.I, 64));
  return g;
}

static int64_t pack_elems(const PackGeom& g) {
  switch (g.kind) {
    case MSIG_WPACK_FWD: return int64_t(g.Opad) * g.RS * g.I;
    case MSIG_WPACK_DGRAD_S1: return int64_t(g.Ipad) * g.RS * g.OC;
    case MSIG_WPACK_DGRAD_S2: return int64_t(4) * g.Ipad * 4 * g.OC;
    case MSIG_WPACK_CONVT_FWD: return int64_t(4) * g.Opad * 4 * g.I;
    case MSIG_WPACK_CONVT_DGRAD: return int64_t(g.Ipad) * 16 * g.OC;
    case MSIG_WPACK_IM2COL: return int64_t(g.Opad) * g.Kpad;
    case MSIG_WPACK_IM2COL_DGRAD: return int64_t(g.Kpad) * g.OC;
    case MSIG_WPACK_IM2COL_FLIP: return int64_t(g.Ipad) * g.Kpad;
    case MSIG_WPACK_ROWPATCH: return (g.I <= 8 && g.S <= 8) ? int64_t(g.Opad) * g.R * 64 : -1;
    case MSIG_WPACK_ROWPATCH_FLIP: return (g.O <= 8 && g.S <= 8) ? int64_t(g.Ipad) * g.R * 64 : -1;
    case MSIG_WPACK_ROWFOLD: return (g.O <= 4 && g.S <= 8 && g.I == 64) ? int64_t(g.R) * 32 * g.I : -1;
    case MSIG_WPACK_ROWFOLD_DGRAD: return (g.I <= 4 && g.S <= 8 && g.O == 64) ? int64_t(g.R) * 32 * g.O : -1;
    default: return -1;
  }
}

// (o, i, t) of the master weight -> linear index in the master tensor.
__device__ __forceinline__ void decode_src(const PackGeom& g, int64_t idx, int& o, int& i, int& t) {
  if (g.kind == MSIG_WPACK_CONVT_FWD || g.kind == MSIG_WPACK_CONVT_DGRAD) {  // [I][O][4][4]
    t = static_cast<int>(idx % 16);
    o = static_cast<int>((idx / 16) % g.O);
    i = static_cast<int>(idx / (int64_t(16) * g.O));
  } else {  // [O][I][R][S]
    t = static_cast<int>(idx % g.RS);
    i = static_cast<int>((idx / g.RS) % g.I);
    o = static_cast<int>(idx / (int64_t(g.RS) * g.I));
  }
}

// (o, i, t) -> offset in the packed bf16 matrix
__device__ __forceinline__ int64_t packed_offset(const PackGeom& g, int o, int i, int t) {
  const int oo = o + g.OOFF;
  switch (g.kind) {
    case MSIG_WPACK_FWD: return (int64_t(oo) * g.RS + t) * g.I + i;
    case MSIG_WPACK_DGRAD_S1: return (int64_t(i) * g.RS + (g.RS - 1 - t)) * g.OC + oo;
    case MSIG_WPACK_DGRAD_S2: {
      int py, ty, px, tx;
      r_to_ph(t / 4, py, ty);
      r_to_ph(t % 4, px, tx);
      return ((int64_t(py * 2 + px) * g.Ipad + i) * 4 + (ty * 2 + tx)) * g.OC + oo;
    }
    case MSIG_WPACK_CONVT_FWD: {
      int py, ty, px, tx;
      r_to_ph(t / 4, py, ty);
      r_to_ph(t % 4, px, tx);
      return ((int64_t(py * 2 + px) * g.Opad + oo) * 4 + (ty * 2 + tx)) * g.I + i;
    }
    case MSIG_WPACK_CONVT_DGRAD: return (int64_t(i) * 16 + t) * g.OC + oo;
    case MSIG_WPACK_IM2COL: return int64_t(oo) * g.Kpad + t * g.I + i;
    case MSIG_WPACK_IM2COL_DGRAD: return (int64_t(t) * g.I + i) * g.OC + oo;
    case MSIG_WPACK_IM2COL_FLIP: return int64_t(i) * g.Kpad + (g.RS - 1 - t) * g.OC + oo;
    case MSIG_WPACK_ROWPATCH: {       // [o][r][s*8 + i]
      const int r = t / g.S, s2 = t % g.S;
      return (int64_t(oo) * g.R + r) * 64 + s2 * 8 + i;
    }
    case MSIG_WPACK_ROWPATCH_FLIP: {  // [i][R-1-r][(S-1-s)*8 + o]
      const int r = t / g.S, s2 = t % g.S;
      return (int64_t(i) * g.R + (g.R - 1 - r)) * 64 + (g.S - 1 - s2) * 8 + oo;
    }
    case MSIG_WPACK_ROWFOLD: {        // [r][s*4 + o][i]
      const int r = t / g.S, s2 = t % g.S;
      return (int64_t(r) * 32 + s2 * 4 + oo) * g.I + i;
    }
    case MSIG_WPACK_ROWFOLD_DGRAD: {  // [R-1-r][(S-1-s)*4 + i][o]
      const int r = t / g.S, s2 = t % g.S;
      return (int64_t(g.R - 1 - r) * 32 + (g.S - 1 - s2) * 4 + i) * g.O + oo;
    }
  }
  return 0;
}

__global__ void wpack_kernel(PackGeom g, const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                             int64_t numel) {
  pdl_entry();
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < numel;
       idx += int64_t(gridDim.x) * blockDim.x) {
    int o, i, t;
    decode_src(g, idx, o, i, t);
    out[packed_offset(g, o, i, t)] = __float2bfloat16(w[idx]);
  }
}

// All weight packs of one network in ONE launch: a device-resident table of jobs (built once per
// network by msig_wpack_table_build) indexed by a prefix sum over their element counts.
struct PackJob {
  PackGeom g;            // g.kind < 0: plain fp32 copy of `numel` elements (bias tables)
  const float* w;
  void* out;
  int64_t start, numel;  // element range in the generic kernel's index space (numel = 0: a tiled job)
  int64_t tile_start;    // tile range in the tiled kernel's index space
  int32_t tiles, tile_mode;   // tile_mode: 0 none, 1 = FWD (one tile per o), 2 = DGRAD_S1 (64 o x 64 (i,t) tiles)
};

// The big packs (conv / Linear weights in the FWD and DGRAD_S1 layouts: 95 % of a generator's bytes) go
// through shared memory so that BOTH sides are coalesced: the generic kernel below reads fp32 linearly but
// scatters 2-byte stores (one 32-byte sector per element).
//   mode 1, FWD     [O][I][RS] -> [o + off][RS][I]: the I*RS elements of one o stay one contiguous range,
//                   permuted (i, t) -> (t, i): one block per o, transposed in shared memory.
//   mode 2, DGRAD_S1 [O][I][RS] -> [i][RS-1-t][o + off]: o becomes the fastest index: 64 (o) x 64 (i, t)
//                   tiles, read along (i, t), written along o (128 bytes per warp store).
constexpr int kPackTile = 64;
constexpr int kPackFwdMax = 8192;     // I * RS elements of one output channel held in shared memory

__global__ void __launch_bounds__(256) wpack_multi_tiled_kernel(const PackJob* __restrict__ jobs, int n_jobs,
                                                                int64_t total_tiles) {
  pdl_entry();
  __shared__ __align__(16) unsigned char smem_raw[kPackTile * (kPackTile + 1) * 4 > (kPackFwdMax + 64) * 2
                                                      ? kPackTile * (kPackTile + 1) * 4
                                                      : (kPackFwdMax + 64) * 2];
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int lo = 0, hi = n_jobs - 1;                      // last job with tile_start <= tile (and tiles > 0)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].tile_start <= tile) lo = mid; else hi = mid - 1;
    }
    const PackJob& j = jobs[lo];
    const PackGeom& g = j.g;
    const int lt = static_cast<int>(tile - j.tile_start);
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(j.out);
    __syncthreads();                                  // the previous tile's readers are done with smem
    if (j.tile_mode == 1) {
      const int o = lt, n = g.I * g.RS, ld = g.I + 2;  // padded rows: conflict-free transposed writes
      __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(smem_raw);
      const float* src = j.w + int64_t(o) * n;
      for (int e = threadIdx.x; e < n; e += 256) {
        const int i = e / g.RS, t = e - i * g.RS;
        sm[t * ld + i] = __float2bfloat16(__ldg(src + e));
      }
      __syncthreads();
      __nv_bfloat16* dst = out + int64_t(o + g.OOFF) * n;
      const int half = g.I / 2;                       // I is even (eligibility): bf16x2 stores
      for (int e = threadIdx.x; e < g.RS * half; e += 256) {
        const int t = e / half, i2 = e - t * half;
        *reinterpret_cast<uint32_t*>(dst + t * g.I + 2 * i2) = *reinterpret_cast<const uint32_t*>(sm + t * ld + 2 * i2);
      }
    } else {
      float* sm = reinterpret_cast<float*>(smem_raw);  // [64 o][65]
      const int n = g.I * g.RS;
      const int e_tiles = (n + kPackTile - 1) / kPackTile;
      const int o0 = (lt / e_tiles) * kPackTile, e0 = (lt % e_tiles) * kPackTile;
      for (int q = threadIdx.x; q < kPackTile * kPackTile; q += 256) {
        const int ol = q / kPackTile, x = q % kPackTile;
        sm[ol * (kPackTile + 1) + x] = (e0 + x < n) ? __ldg(j.w + int64_t(o0 + ol) * n + e0 + x) : 0.f;
      }
      __syncthreads();
      for (int q = threadIdx.x; q < kPackTile * (kPackTile / 2); q += 256) {
        const int x = q / (kPackTile / 2), o2 = q % (kPackTile / 2);
        const int e = e0 + x;
        if (e >= n) continue;
        const int i = e / g.RS, t = e - i * g.RS;
        const int64_t row = int64_t(i) * g.RS + (g.RS - 1 - t);
        const __nv_bfloat162 v = __floats2bfloat162_rn(sm[(2 * o2) * (kPackTile + 1) + x],
                                                       sm[(2 * o2 + 1) * (kPackTile + 1) + x]);
        *reinterpret_cast<__nv_bfloat162*>(out + row * g.OC + o0 + g.OOFF + 2 * o2) = v;
      }
    }
  }
}

__global__ void wpack_multi_kernel(const PackJob* __restrict__ jobs, int n_jobs, int64_t total) {
  pdl_entry();
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    int lo = 0, hi = n_jobs - 1;                      // last job with start <= idx
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].start <= idx) lo = mid; else hi = mid - 1;
    }
    const PackJob& j = jobs[lo];
    const int64_t li = idx - j.start;
    if (j.g.kind < 0) {
      reinterpret_cast<float*>(j.out)[li] = j.w[li];
    } else {
      int o, i, t;
      decode_src(j.g, li, o, i, t);
      reinterpret_cast<__nv_bfloat16*>(j.out)[packed_offset(j.g, o, i, t)] = __float2bfloat16(j.w[li]);
    }
  }
}

// Layout of the fp32 split-K partials written by the wgrad kernel, per kind:
//   FWD          [O][RS][I]            CONVT_FWD     [O][16 = phase*4+tap][I]
//   IM2COL       [O][Kpad]             IM2COL_FLIP   [Kpad][I]
__device__ __forceinline__ int64_t partial_offset(const PackGeom& g, int o, int i, int t) {
  switch (g.kind) {
    case MSIG_WPACK_FWD:
      return g.partT ? (int64_t(i) * g.RS + t) * g.O + o : (int64_t(o) * g.RS + t) * g.I + i;
    case MSIG_WPACK_CONVT_FWD: {
      int py, ty, px, tx;
      r_to_ph(t / 4, py, ty);
      r_to_ph(t % 4, px, tx);
      return (int64_t(o) * 16 + (py * 2 + px) * 4 + (ty * 2 + tx)) * g.I + i;
    }
    case MSIG_WPACK_IM2COL: return int64_t(o) * g.Kpad + t * g.I + i;
    case MSIG_WPACK_IM2COL_FLIP: return (int64_t(g.RS - 1 - t) * g.OC + o + g.OOFF) * g.I + i;
    case MSIG_WPACK_ROWPATCH: {       // [r / 4][o (64 rows)][(r % 4) * 64 + s*8 + i]
      const int r = t / g.S, s2 = t % g.S;
      return (int64_t(r / 4) * 64 + o) * 256 + (r % 4) * 64 + s2 * 8 + i;
    }
    case MSIG_WPACK_ROWPATCH_FLIP: {  // r' = R-1-r, s' = S-1-s:  [r' / 4][i (64 rows)][(r' % 4) * 64 + s'*8 + o]
      const int r = g.R - 1 - t / g.S, s2 = g.S - 1 - t % g.S;
      return (int64_t(r / 4) * 64 + i) * 256 + (r % 4) * 64 + s2 * 8 + o;
    }
  }
  return 0;
}

__global__ void wgrad_reduce_kernel(PackGeom g, const float* __restrict__ partial, int splits,
                                    int64_t split_stride, float* __restrict__ dw, int accumulate,
                                    int64_t numel) {
  pdl_entry();
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < numel;
       idx += int64_t(gridDim.x) * blockDim.x) {
    int o, i, t;
    decode_src(g, idx, o, i, t);
    const int64_t off = partial_offset(g, o, i, t);
    float acc = accumulate ? dw[idx] : 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[s * split_stride + off];
    dw[idx] = acc;
  }
}

// Same reduction for the conv layouts ([O][RS][I] or, with swapped operand roles, [I][RS][O]), walking
// the PARTIALS linearly: the (splits x) reads are coalesced, only the single write per element strides.
__global__ void wgrad_reduce_fwd_kernel(PackGeom g, const float* __restrict__ partial, int splits,
                                        int64_t split_stride, float* __restrict__ dw, int accumulate,
                                        int64_t numel) {
  pdl_entry();
  const int inner = g.partT ? g.O : g.I;
  for (int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; q < numel;
       q += int64_t(gridDim.x) * blockDim.x) {
    const int a = static_cast<int>(q % inner);
    const int64_t r2 = q / inner;
    const int t = static_cast<int>(r2 % g.RS);
    const int b = static_cast<int>(r2 / g.RS);
    const int o = g.partT ? a : b, i = g.partT ? b : a;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[s * split_stride + q];
    const int64_t idx = (int64_t(o) * g.I + i) * g.RS + t;
    dw[idx] = accumulate ? dw[idx] + acc : acc;
  }
}

// Swapped-role partials [it = i*RS + t][o] -> dw[o][it]: a tiled transpose, so that both the (splits x)
// reads (along o) and the single read-modify-write per element (along it) are 128-byte coalesced.
// Tile = 32 (it) x 32 (o); block = 32 x 8 threads.
__global__ void __launch_bounds__(256) wgrad_reduce_t_kernel(const float* __restrict__ partial, int splits,
                                                             int64_t split_stride, float* __restrict__ dw,
                                                             int accumulate, int IT, int O) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int it0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int it = it0 + r, o = o0 + threadIdx.x;
    float acc = 0.f;
    if (it < IT && o < O) {
      const float* pp = partial + int64_t(it) * O + o;
      for (int s = 0; s < splits; ++s) acc += pp[s * split_stride];
    }
    tile[r][threadIdx.x] = acc;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int o = o0 + r, it = it0 + threadIdx.x;
    if (it < IT && o < O) {
      float* d = dw + int64_t(o) * IT + it;
      *d = accumulate ? *d + tile[threadIdx.x][r] : tile[threadIdx.x][r];
    }
  }
}

static int launch_wgrad_reduce(const PackGeom& g, const float* partial, int splits,
                               int64_t split_stride, float* dw, int accumulate, cudaStream_t st) {
  const int64_t numel = int64_t(g.O) * g.I * g.RS;
  if (g.kind == MSIG_WPACK_FWD && g.OOFF == 0 && g.OC == g.O && g.partT) {
    const int IT = g.I * g.RS;
    MSIG_LAUNCH((wgrad_reduce_t_kernel), dim3(static_cast<unsigned>(ceil_div(IT, 32)), static_cast<unsigned>(ceil_div(g.O, 32))), dim3(32, 8), 0, st, partial, splits, split_stride, dw, accumulate, IT, g.O);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
    return MSIG_OK;
  }
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(numel, threads), 4096));
  if (g.kind == MSIG_WPACK_FWD && g.OOFF == 0 && g.OC == g.O) {
    MSIG_LAUNCH((wgrad_reduce_fwd_kernel), blocks, threads, 0, st, g, partial, splits, split_stride, dw, accumulate, numel);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
    return MSIG_OK;
  }
  MSIG_LAUNCH((wgrad_reduce_kernel), blocks, threads, 0, st, g, partial, splits, split_stride, dw, accumulate,
                                                  numel);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

// ------------------------------------------------------------------------ TMA descriptors
struct ActView {
  const void* base;
  int64_t C, W, H, N;      // extents
  int64_t sW, sH, sN;      // element strides (channel stride is 1)
};

static int make_act_map(CUtensorMap* m, const ActView& v, int boxW, int boxH) {
  if ((reinterpret_cast<uintptr_t>(v.base) & 15) != 0)
    return set_error(MSIG_ERR_ARG, "activation base %p is not 16-byte aligned", v.base);
  cuuint64_t dims[4] = {cuuint64_t(v.C), cuuint64_t(v.W), cuuint64_t(v.H), cuuint64_t(v.N)};
  cuuint64_t strides[3] = {cuuint64_t(v.sW) * 2, cuuint64_t(v.sH) * 2, cuuint64_t(v.sN) * 2};
  cuuint32_t box[4] = {64, cuuint32_t(boxW), cuuint32_t(boxH), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0) return set_error(MSIG_ERR_ARG, "TMA stride %d not a multiple of 16 B", i);
  CUresult r = encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MSIG_ERR_CUDA,
                     "cuTensorMapEncodeTiled(act) failed: %d (dims %lld,%lld,%lld,%lld box %d,%d)", int(r),
                     (long long)v.C, (long long)v.W, (long long)v.H, (long long)v.N, boxW, boxH);
  return MSIG_OK;
}

static int make_w_map(CUtensorMap* m, const void* w, int64_t rows, int64_t K, int box_rows) {
  if ((reinterpret_cast<uintptr_t>(w) & 15) != 0)
    return set_error(MSIG_ERR_ARG, "packed weight base %p is not 16-byte aligned", w);
  if (K % 64 != 0) return set_error(MSIG_ERR_ARG, "packed weight K=%lld is not a multiple of 64", (long long)K);
  cuuint64_t dims[2] = {cuuint64_t(K), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(K) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(MSIG_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d (rows %lld K %lld box %d)",
                     int(r), (long long)rows, (long long)K, box_rows);
  return MSIG_OK;
}

// 128-pixel output tile: TW x TH.
static void pick_tile(int OW, int& TW, int& TH) {
  if (OW == 64) { TW = 64; TH = 2; }
  else if (OW == 32) { TW = 32; TH = 4; }
  else if (OW == 16) { TW = 16; TH = 8; }
  else { TW = 128; TH = 1; }
}
// 64-pixel K block of the wgrad kernel: PW x PH.
static void pick_kblock(int OW, int& PW, int& PH) {
  if (OW == 32) { PW = 32; PH = 2; }
  else if (OW == 16) { PW = 16; PH = 4; }
  else { PW = 64; PH = 1; }
}

struct OutView {
  void* base;
  int64_t sN, sH, sW, sC;
  int f32;
};

// Output / aux views for a [n, OH, OW, k] result in the requested layout.
static OutView make_out_view(void* y, int layout, int64_t OH, int64_t OW, int64_t k) {
  OutView v;
  v.base = y;
  if (layout == MSIG_OUT_F32_NCHW) {
    v.sN = k * OH * OW; v.sC = OH * OW; v.sH = OW; v.sW = 1; v.f32 = 1;
  } else {
    v.sN = OH * OW * k; v.sH = OW * k; v.sW = k; v.sC = 1; v.f32 = (layout == MSIG_OUT_F32_NHWC);
  }
  return v;
}

static int fill_epilogue(FpropParams& p, const msig_epilogue* e, const OutView& ov, int n_valid) {
  p.out = ov.base;
  p.o_sn = ov.sN; p.o_sh = ov.sH; p.o_sw = ov.sW; p.o_sc = ov.sC;
  p.out_f32 = ov.f32;
  p.n_valid = n_valid;
  p.bias = e ? e->bias : nullptr;
  p.alpha = e ? e->alpha : 1.f;
  p.alpha_ptr = e ? e->alpha_ptr : nullptr;
  p.act = e ? e->act : ACT_NONE;
  p.slope = e ? e->slope : 0.f;
  p.aux = e ? reinterpret_cast<const __nv_bfloat16*>(e->aux) : nullptr;
  p.aux_mode = (e && e->aux) ? e->aux_mode : AUX_NONE;
  if (e && (e->mask_scale != nullptr || e->mask_shift != nullptr)) {
    // activation mask recomputed from the norm input (stats_z) and its per-(image, channel) scale / shift
    if (e->mask_scale == nullptr || e->mask_shift == nullptr || e->stats_z == nullptr ||
        (e->aux_mode != AUX_RELU_MASK && e->aux_mode != AUX_LRELU_MASK) || p.fold_c != 0 || p.tap_is_image ||
        (n_valid % 16) != 0 || ((reinterpret_cast<uintptr_t>(e->mask_scale) | reinterpret_cast<uintptr_t>(e->mask_shift)) & 15) != 0)
      return set_error(MSIG_ERR_ARG, "epilogue mask_scale/mask_shift need both pointers (16-byte aligned), stats_z, "
                                     "a ReLU / LeakyReLU mask mode and k %% 16 == 0");
    p.aux = nullptr;
    p.aux_mode = e->aux_mode;
    p.mask_scale = e->mask_scale; p.mask_shift = e->mask_shift; p.mask_ld = n_valid;
  }
  p.a_sn = ov.sN; p.a_sh = ov.sH; p.a_sw = ov.sW;   // aux is congruent with a bf16 NHWC output
  if (p.aux_mode != AUX_NONE && (ov.f32 || ov.sC != 1))
    return set_error(MSIG_ERR_UNSUPPORTED, "epilogue aux needs a bf16 NHWC output");
  if (!ov.f32 && ov.sC == 1 && (n_valid % 8) != 0)
    return set_error(MSIG_ERR_UNSUPPORTED, "bf16 NHWC output needs k %% 8 == 0 (k=%d)", n_valid);
  if ((p.bias != nullptr) && !ov.f32 && (reinterpret_cast<uintptr_t>(p.bias) & 15) != 0)
    return set_error(MSIG_ERR_ARG, "bias pointer must be 16-byte aligned");
  p.stat_out = e ? e->stats_partial : nullptr;
  p.stat_z = e ? reinterpret_cast<const __nv_bfloat16*>(e->stats_z) : nullptr;
  p.stat_ld = static_cast<int>(round_up(n_valid, 64));
  p.ring_item_stats = (e && e->stats_partial && e->stats_rows > 0) ? e->stats_rows : 0;   // validated by the ring launcher
  if (e && e->ch_scale != nullptr)
    return set_error(MSIG_ERR_UNSUPPORTED, "epilogue ch_scale is only implemented by msig_conv_narrow_fwd");
  if (p.stat_out != nullptr && (ov.f32 || ov.sC != 1 || n_valid < 64 || p.fold_c != 0 || p.tap_is_image))
    return set_error(MSIG_ERR_UNSUPPORTED, "epilogue statistics need a bf16 NHWC output with k >= 64");
  return MSIG_OK;
}

static int g_ring_mode = 3;    // test hook: 0 = generic per-tap kernel for the 64-channel stride-1 layers; bit 1 =
                               // the four phases of a transposed conv in one ring launch (else one launch each)
static int g_wgrad_mode = 3;   // test hook: bit 0 = M-stacked row-patch weight gradients, bit 1 = tap-grouped convT ones
#define g_rowpatch_stack (g_wgrad_mode & 1)
#define g_convt_group (g_wgrad_mode & 2)

// Configure `p` (epilogue already filled, n_img / OH / OW set) for the N = 64 ring kernel.
static void set_ring(FpropParams& p, int R, int S, int org_h, int org_w, int nph = 1) {
  p.TW = 128; p.TH = 1;
  p.tiles_h = p.OH;
  p.tiles_w = static_cast<int>(ceil_div(p.OW, 128));
  p.n_blocks = 1;
  p.strip_r = R; p.strip_s = S;
  p.org_h = org_h; p.org_w = org_w;
  p.ring_cb = 1;
  for (int t = 0; t < R * S && t < 16; ++t) p.ring_tap[t] = static_cast<int8_t>(t);
  const int sms = (sm_count() > 0 ? sm_count() : 148) / nph;      // CTAs that share one phase's items
  int rows = 64;
  while (rows > 8 && int64_t(p.n_img) * p.tiles_w * ceil_div(p.OH, rows) < int64_t(6) * sms) rows /= 2;
  p.ring_rows = rows;
  p.ring_chunks = static_cast<int>(ceil_div(p.OH, rows));
}

static void init_fprop(FpropParams& p) {
  memset(&p, 0, sizeof(p));
  p.phases = 1;
  p.alpha = 1.f;
}

// Plain (stride 1 or 2) convolution of `in` [n,h,w,c] with a [rows][taps*c] packed matrix.
static int run_conv(const void* in, int n, int h, int w, int c, int k, int R, int S, int stride,
                    int pad_t, int pad_l, int OH, int OW, const void* wpk, const msig_epilogue* e,
                    void* out, cudaStream_t st) {
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(c % 64 == 0, "conv: input channels (%d) must be a multiple of 64", c);
  MSIG_REQUIRE(stride == 1 || stride == 2, "conv: stride %d unsupported", stride);
  MSIG_REQUIRE(R * S <= kMaxTaps, "conv: %dx%d filter has too many taps", R, S);
  FpropParams p;
  init_fprop(p);
  const int k_pad = pad_rows(k);
  const int block_n = pick_block_n(k_pad);
  pick_tile(OW, p.TW, p.TH);
  p.OH = OH; p.OW = OW;
  p.tiles_h = static_cast<int>(ceil_div(OH, p.TH));
  p.tiles_w = static_cast<int>(ceil_div(OW, p.TW));
  p.n_img = n;
  p.n_blocks = k_pad / block_n;
  p.taps = R * S;
  p.cblocks = c / 64;
  p.m2 = fprop_uses_m2(p, block_n) ? 1 : 0;            // two m-tiles per CTA: the A box is twice as tall
  const int boxH = p.m2 ? 2 * p.TH : p.TH;
  int rc;
  if (stride == 1) {
    ActView v{in, c, w, h, n, c, int64_t(w) * c, int64_t(h) * w * c};
    if ((rc = make_act_map(&p.tmA[0], v, p.TW, boxH)) != MSIG_OK) return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) p.tap[r * S + s] = Tap{int8_t(r - pad_t), int8_t(s - pad_l), 0, 0};
  } else {
    MSIG_REQUIRE(h % 2 == 0 && w % 2 == 0, "stride-2 conv needs even input dims (%d x %d)", h, w);
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(in) + (int64_t(ph) * w + pw) * c;
        ActView v{base, c, w / 2, h / 2, n, int64_t(2) * c, int64_t(2) * w * c, int64_t(h) * w * c};
        if ((rc = make_act_map(&p.tmA[ph * 2 + pw], v, p.TW, boxH)) != MSIG_OK) return rc;
      }
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        const int rr = r - pad_t, ss = s - pad_l;
        const int ph = ((rr % 2) + 2) % 2, pw = ((ss % 2) + 2) % 2;
        p.tap[r * S + s] = Tap{int8_t((rr - ph) / 2), int8_t((ss - pw) / 2), int8_t(ph * 2 + pw), 0};
      }
  }
  if ((rc = make_w_map(&p.tmB, wpk, k_pad, int64_t(R) * S * c, fprop_uses_pairs(p, block_n) ? 128 : block_n)) != MSIG_OK)
    return rc;
  const OutView ov = make_out_view(out, e ? e->out_layout : MSIG_OUT_BF16_NHWC, OH, OW, k);
  if ((rc = fill_epilogue(p, e, ov, k)) != MSIG_OK) return rc;
  // 64 -> <=64 channel stride-1 layers on wide planes (VGG conv 1_2 fwd / dgrad): resident filter + strip ring.
  if (g_ring_mode != 0 && stride == 1 && c == 64 && block_n == 64 && k_pad == 64 && R * S >= 2 && R * S <= 9 && R <= 7 &&
      S <= 8 && OW >= 128) {
    set_ring(p, R, S, -pad_t, -pad_l);
    ActView v{in, c, w, h, n, c, int64_t(w) * c, int64_t(h) * w * c};
    if ((rc = make_act_map(&p.tmA[1], v, 128 + S - 1, 1)) != MSIG_OK) return rc;
    cudaError_t ce = launch_fprop_ring64(p, sm_count(), st);
    if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(ring) launch: %s", cudaGetErrorString(ce));
    return MSIG_OK;
  }
  cudaError_t ce = launch_fprop(p, block_n, sm_count(), st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop launch: %s", cudaGetErrorString(ce));
  return MSIG_OK;
}

// Phase-decomposed k4 s2 p1 transposed structure: `in` [n,h,w,c] -> out [n,2h,2w,k]; packed
// weights are 4 slabs of [k_pad][4*c].
static int run_phased(const void* in, int n, int h, int w, int c, int k, const void* wpk,
                      const msig_epilogue* e, void* out, cudaStream_t st) {
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(c % 64 == 0, "transposed conv: input channels (%d) must be a multiple of 64", c);
  MSIG_REQUIRE(k % 8 == 0, "transposed conv: output channels (%d) must be a multiple of 8", k);
  const int layout = e ? e->out_layout : MSIG_OUT_BF16_NHWC;
  MSIG_REQUIRE(layout == MSIG_OUT_BF16_NHWC, "transposed conv writes bf16 NHWC only");
  FpropParams p;
  init_fprop(p);
  const int k_pad = pad_rows(k);
  const int block_n = pick_block_n(k_pad);
  pick_tile(w, p.TW, p.TH);
  p.OH = h; p.OW = w;                      // per-phase output plane
  p.tiles_h = static_cast<int>(ceil_div(h, p.TH));
  p.tiles_w = static_cast<int>(ceil_div(w, p.TW));
  p.n_img = n;
  p.n_blocks = k_pad / block_n;
  p.taps = 4;
  p.cblocks = c / 64;
  p.phases = 4;
  p.b_row_per_phase = k_pad;
  p.m2 = fprop_uses_m2(p, block_n) ? 1 : 0;
  int rc;
  ActView v{in, c, w, h, n, c, int64_t(w) * c, int64_t(h) * w * c};
  if ((rc = make_act_map(&p.tmA[0], v, p.TW, p.m2 ? 2 * p.TH : p.TH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  const int64_t OW2 = 2 * int64_t(w);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int ph = py * 2 + px;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx)
          p.tap[ph * 4 + ty * 2 + tx] = Tap{int8_t(ph_d(py, ty)), int8_t(ph_d(px, tx)), 0, 0};
      p.o_ph[ph] = (py * OW2 + px) * k;
      p.a_ph[ph] = p.o_ph[ph];
    }
  if ((rc = make_w_map(&p.tmB, wpk, int64_t(4) * k_pad, int64_t(4) * c, block_n)) != MSIG_OK) return rc;
  OutView ov;
  ov.base = out; ov.f32 = 0; ov.sC = 1;
  ov.sN = int64_t(4) * h * w * k; ov.sH = 2 * OW2 * k; ov.sW = 2 * int64_t(k);
  if ((rc = fill_epilogue(p, e, ov, k)) != MSIG_OK) return rc;
  // 128 -> 64 channel layers on wide planes (model.py:140 forward, the dgrad of model.py:132): every phase is
  // a 2x2 stride-1 conv; run each through the strip-ring kernel (resident 64 KiB filter slab, input rows
  // shared by consecutive output rows) instead of re-fetching 24 KiB per K block.
  if (g_ring_mode != 0 && c == 128 && block_n == 64 && k_pad == 64 && w >= 128 &&
      (p.stat_out == nullptr || (p.ring_item_stats != 0 && (g_ring_mode & 2) != 0))) {
    // input rows / columns of a phase in ascending order: ring position r <-> tap 1 - r
    auto ring_taps = [](FpropParams& q) {
      q.ring_cb = 2;
      for (int r = 0; r < 2; ++r)
        for (int s2 = 0; s2 < 2; ++s2) q.ring_tap[r * 2 + s2] = static_cast<int8_t>((1 - r) * 2 + (1 - s2));
    };
    if (g_ring_mode & 2) {
      // all four phases in ONE launch (CTA b -> phase b % 4): the input is read from HBM once
      set_ring(p, 2, 2, 0, 0, 4);
      ring_taps(p);
      p.ring_phases = 4;
      for (int ph = 0; ph < 4; ++ph) {
        p.ring_org_h[ph] = (ph >> 1) == 0 ? -1 : 0;
        p.ring_org_w[ph] = (ph & 1) == 0 ? -1 : 0;
      }
      if ((rc = make_act_map(&p.tmA[1], v, 128 + 1, 1)) != MSIG_OK) return rc;
      if ((rc = make_w_map(&p.tmB, wpk, int64_t(4) * k_pad, int64_t(4) * c, 64)) != MSIG_OK) return rc;
      cudaError_t ce = launch_fprop_ring64(p, sm_count(), st);
      if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(4-phase ring) launch: %s", cudaGetErrorString(ce));
      return MSIG_OK;
    }
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int ph = py * 2 + px;
        FpropParams q = p;
        q.phases = 1;
        q.out = reinterpret_cast<__nv_bfloat16*>(out) + p.o_ph[ph];
        if (q.aux != nullptr) q.aux = q.aux + p.a_ph[ph];
        set_ring(q, 2, 2, py == 0 ? -1 : 0, px == 0 ? -1 : 0);
        ring_taps(q);
        if ((rc = make_act_map(&q.tmA[1], v, 128 + 1, 1)) != MSIG_OK) return rc;
        const __nv_bfloat16* wph = reinterpret_cast<const __nv_bfloat16*>(wpk) + int64_t(ph) * k_pad * 4 * c;
        if ((rc = make_w_map(&q.tmB, wph, k_pad, int64_t(4) * c, 64)) != MSIG_OK) return rc;
        cudaError_t ce = launch_fprop_ring64(q, sm_count(), st);
        if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(phased ring) launch: %s", cudaGetErrorString(ce));
      }
    return MSIG_OK;
  }
  cudaError_t ce = launch_fprop(p, block_n, sm_count(), st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(phased) launch: %s", cudaGetErrorString(ce));
  return MSIG_OK;
}

// ------------------------------------------------------------------------ wgrad driver
struct WgradPlan {
  int block_n, m_blocks, n_blocks, splits, kb_per_split, kb_total;
};

static WgradPlan plan_wgrad(int M, int N, int taps, int64_t kb_total, bool upper_only = false) {
  WgradPlan pl;
  pl.block_n = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  pl.m_blocks = static_cast<int>(ceil_div(M, 128));
  pl.n_blocks = static_cast<int>(ceil_div(N, pl.block_n));
  int64_t base = int64_t(pl.m_blocks) * pl.n_blocks * taps;
  if (upper_only) {   // only tiles on/above the diagonal do work
    base = 0;
    for (int m = 0; m < pl.m_blocks; ++m)
      for (int n = 0; n < pl.n_blocks; ++n)
        if ((n + 1) * pl.block_n > m * 128) ++base;
  }
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // one CTA per SM is resident (192 KiB of smem): fill ONE wave as completely as possible
  int64_t splits = std::max<int64_t>(1, int64_t(sms) / base);
  splits = std::max<int64_t>(1, std::min<int64_t>(splits, ceil_div(kb_total, 4)));
  pl.kb_per_split = static_cast<int>(ceil_div(kb_total, splits));
  pl.splits = static_cast<int>(ceil_div(kb_total, pl.kb_per_split));
  pl.kb_total = static_cast<int>(kb_total);
  return pl;
}


__global__ void sum_splits_kernel(const float* __restrict__ partial, int splits, int64_t stride,
                                  float* __restrict__ out, int64_t numel) {
  pdl_entry();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < numel;
       i += int64_t(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[s * stride + i];
    out[i] = acc;
  }
}

}  // namespace msig

using namespace msig;

extern "C" {

int32_t msig_epilogue_stats_rows(int32_t oh, int32_t ow, int32_t phases) {
  int TW, TH;
  pick_tile(ow, TW, TH);
  return static_cast<int32_t>(ceil_div(oh, TH) * ceil_div(ow, TW) * phases * 4);
}

// Rows PER IMAGE of msig_epilogue.stats_partial when the strip-ring kernel writes one partial row per (work item,
// phase, accumulator quadrant) (msig_epilogue.stats_rows); 0: this layer does not run on the ring kernel (use the
// per-tile rows of msig_epilogue_stats_rows, or a separate statistics pass). kind 0: msig_conv_rowpatch_fwd,
// kind 1: msig_convT2d_fwd, kind 2: msig_conv2d_fwd (64 -> 64 channels, stride 1), kind 3: msig_conv2d_dgrad of a
// 4x4 stride-2 conv 64 -> 128 (with stats_z + mask_scale / mask_shift: the norm-backward reductions).
int32_t msig_ring_stats_rows(int32_t kind, const msig_conv_geom* g) {
  if (!g || g_ring_mode == 0 || !context_ready()) return 0;
  FpropParams p;
  init_fprop(p);
  p.n_img = g->n;
  int nph = 1;
  if (kind == 0) {
    if (pad_rows(g->k) != 64 || g->k != 64 || g->r > 7 || g->ow < 128 || g->stride != 1) return 0;
    p.OH = g->oh; p.OW = g->ow;
  } else if (kind == 1) {
    if (g->c != 128 || g->k != 64 || g->w < 128 || (g_ring_mode & 2) == 0) return 0;
    p.OH = g->h; p.OW = g->w;
    nph = 4;
  } else if (kind == 3) {       // msig_conv2d_dgrad of a 4x4 stride-2 conv 64 -> 128: a phased 128 -> 64 structure over dy
    if (g->k != 128 || g->c != 64 || g->ow < 128 || g->stride != 2 || g->r != 4 || g->s != 4 || (g_ring_mode & 2) == 0)
      return 0;
    p.OH = g->oh; p.OW = g->ow;
    nph = 4;
  } else if (kind == 2) {
    if (g->stride != 1 || g->c != 64 || g->k != 64 || g->r * g->s < 2 || g->r * g->s > 9 || g->r > 7 || g->s > 8 ||
        g->ow < 128)
      return 0;
    p.OH = g->oh; p.OW = g->ow;
  } else {
    return 0;
  }
  set_ring(p, 1, 1, 0, 0, nph);
  return static_cast<int32_t>(p.tiles_w * p.ring_chunks * nph * 4);
}

// Test hook: ring kernel for the 64-channel stride-1 layers. Bit 0: on; bit 1: the four phases of a transposed
// conv share one launch; bit 2: legacy issue order (one N = 64 MMA chain per OUTPUT row instead of the row-stacked
// N = 64 x R MMAs per input row); bits 8..15: cap on the ring depth (0 = whatever fits). Default 3.
int msig_debug_set_ring_mode(int mode) {
  g_ring_mode = mode & 3;
  set_ring_legacy((mode & 4) != 0);
  set_ring_slots_cap((mode >> 8) & 0xff);
  return MSIG_OK;
}

// Test hook: weight-gradient operand plans (bit 0: M-stacked row-patch, bit 1: tap-grouped convT); default 3.
int msig_debug_set_wgrad_mode(int mask) {
  g_wgrad_mode = mask;
  return MSIG_OK;
}

// Test hook: CTA-pair (cta_group::2) kernel for 256-wide tiles on (default) / off.
int msig_debug_set_m2_mode(int on) {
  set_m2_mode(on != 0);
  return MSIG_OK;
}

int msig_debug_set_pair_mode(int on) {
  set_pair_mode(on != 0);
  return MSIG_OK;
}

// Test hook: selects the kernel variant of the narrow-output 7x7 conv (see run_conv).
size_t msig_wpack_elems(const msig_wpack_desc* d) {
  if (!d) return 0;
  PackGeom g = make_pack_geom(d, 0, 0);
  const int64_t n = pack_elems(g);
  return n < 0 ? 0 : static_cast<size_t>(n);
}

// Composite variant: this weight provides output channels [o_off, o_off + d->o) of a packed matrix
// with `oc` output channels in total (per-domain heads, model.py:84,183). Padding rows/cols are
// never written: the caller zero-initialises the packed buffer once.
int msig_wpack_part(const msig_wpack_desc* d, int32_t oc, int32_t o_off, const float* w, void* packed,
                    void* stream) {
  MSIG_REQUIRE(d && w && packed, "msig_wpack: null argument");
  PackGeom g = make_pack_geom(d, oc, o_off);
  MSIG_REQUIRE(pack_elems(g) > 0, "msig_wpack: unknown kind %d", d->kind);
  if (g.kind == MSIG_WPACK_DGRAD_S2 || g.kind == MSIG_WPACK_CONVT_FWD || g.kind == MSIG_WPACK_CONVT_DGRAD)
    MSIG_REQUIRE(g.R == 4 && g.S == 4, "msig_wpack: kind %d needs a 4x4 filter", d->kind);
  const int64_t numel = int64_t(g.O) * g.I * g.RS;
  const int threads = 256;
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(numel, threads), 4096));
  MSIG_LAUNCH((wpack_kernel), blocks, threads, 0, static_cast<cudaStream_t>(stream), 
      g, w, reinterpret_cast<__nv_bfloat16*>(packed), numel);
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}
size_t msig_wpack_part_elems(const msig_wpack_desc* d, int32_t oc) {
  if (!d) return 0;
  PackGeom g = make_pack_geom(d, oc, 0);
  const int64_t n = pack_elems(g);
  return n < 0 ? 0 : static_cast<size_t>(n);
}

size_t msig_wpack_table_bytes(int32_t n_jobs) { return size_t(n_jobs) * sizeof(PackJob); }

int msig_wpack_table_build(const msig_wpack_job* jobs, int32_t n_jobs, void* table_host, int64_t* total_out,
                           int64_t* total_tiles_out) {
  MSIG_REQUIRE(jobs && table_host && total_out && total_tiles_out && n_jobs > 0, "msig_wpack_table_build: bad argument");
  PackJob* t = reinterpret_cast<PackJob*>(table_host);
  int64_t start = 0, tile_start = 0;
  for (int k = 0; k < n_jobs; ++k) {
    const msig_wpack_job& j = jobs[k];
    MSIG_REQUIRE(j.src && j.dst, "msig_wpack_table_build: job %d has a null pointer", k);
    memset(&t[k], 0, sizeof(PackJob));
    if (j.d.kind < 0) {
      t[k].g.kind = -1;
      t[k].numel = j.copy_numel;
    } else {
      t[k].g = make_pack_geom(&j.d, j.oc, j.o_off);
      MSIG_REQUIRE(pack_elems(t[k].g) > 0, "msig_wpack_table_build: job %d: unknown kind %d", k, j.d.kind);
      t[k].numel = int64_t(t[k].g.O) * t[k].g.I * t[k].g.RS;
      // big FWD / DGRAD_S1 packs go through the shared-memory tiled kernel (coalesced on both sides)
      const PackGeom& g = t[k].g;
      const int n = g.I * g.RS;
      const bool aligned = (reinterpret_cast<uintptr_t>(j.dst) & 3) == 0;
      if (aligned && g.kind == MSIG_WPACK_FWD && g.I % 2 == 0 && n >= 256 && n + 2 * g.RS <= kPackFwdMax + 64 &&
          ((int64_t(g.OOFF) * n) % 2) == 0) {
        t[k].tile_mode = 1;
        t[k].tiles = g.O;
      } else if (aligned && g.kind == MSIG_WPACK_DGRAD_S1 && g.O % kPackTile == 0 && g.OC % 2 == 0 && g.OOFF % 2 == 0 &&
                 n >= 64) {
        t[k].tile_mode = 2;
        t[k].tiles = (g.O / kPackTile) * static_cast<int>(ceil_div(n, kPackTile));
      }
      if (t[k].tile_mode != 0) t[k].numel = 0;
    }
    t[k].w = j.src;
    t[k].out = j.dst;
    t[k].start = start;
    start += t[k].numel;
    t[k].tile_start = tile_start;
    tile_start += t[k].tiles;
  }
  *total_out = start;
  *total_tiles_out = tile_start;
  return MSIG_OK;
}

int msig_wpack_multi(const void* table_dev, int32_t n_jobs, int64_t total, int64_t total_tiles, void* stream) {
  MSIG_REQUIRE(table_dev && n_jobs > 0 && total >= 0 && total_tiles >= 0 && total + total_tiles > 0,
               "msig_wpack_multi: bad argument");
  if (total > 0) {
    const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(total, 256), 148 * 16));
    MSIG_LAUNCH((wpack_multi_kernel), blocks, 256, 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const PackJob*>(table_dev), n_jobs, total);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
  }
  if (total_tiles > 0) {
    const int blocks = static_cast<int>(std::min<int64_t>(total_tiles, 148 * 8));
    MSIG_LAUNCH((wpack_multi_tiled_kernel), blocks, 256, 0, static_cast<cudaStream_t>(stream), 
        reinterpret_cast<const PackJob*>(table_dev), n_jobs, total_tiles);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
  }
  return MSIG_OK;
}

int msig_wpack(const msig_wpack_desc* d, const float* w, void* packed, void* stream) {
  return msig_wpack_part(d, 0, 0, w, packed, stream);
}

int msig_conv2d_fwd(const msig_conv_geom* g, const void* x, const void* w_fwd, const msig_epilogue* e,
                    void* y, void* stream) {
  MSIG_REQUIRE(g && x && w_fwd && y, "msig_conv2d_fwd: null argument");
  return run_conv(x, g->n, g->h, g->w, g->c, g->k, g->r, g->s, g->stride, g->pad_t, g->pad_l, g->oh,
                  g->ow, w_fwd, e, y, static_cast<cudaStream_t>(stream));
}

static size_t wgrad_ws_bytes(int M, int N, int taps, int64_t kb_total);

// ---------------------------------------------------------------- row-patch convs (few-channel images)
// A few-channel image (c <= 8) is stored zero/reflect-padded as bf16 [n][h+2p][w+2p+2][8] ("pad8").
// The 8 consecutive pixels x 8 channels starting at pixel (y, x) are 64 contiguous bf16 = exactly one
// 128-byte K row, so a TMA map whose W stride is ONE pixel (16 B, overlapping rows) delivers the
// im2col row of filter row r for 128 output pixels as an ordinary K-major operand tile: the
// implicit GEMM runs with taps = R, one 64-wide K block per tap, and no patch matrix ever exists.
static void pick_tile_any(int OH, int OW, int& TW, int& TH) {
  const int cand[4][2] = {{128, 1}, {64, 2}, {32, 4}, {16, 8}};
  int64_t best = -1;
  for (auto& cd : cand) {
    const int64_t area = ceil_div(OW, cd[0]) * cd[0] * ceil_div(OH, cd[1]) * cd[1];
    if (best < 0 || area < best) { best = area; TW = cd[0]; TH = cd[1]; }
  }
}

static ActView pad8_view(const void* x_pad8, const msig_conv_geom* g) {
  const int64_t Hp = g->h + 2 * g->pad_t, Wp = g->w + 2 * g->pad_l + 2;
  // extents: 64 elements per window, one window per output column, every padded row
  return ActView{x_pad8, 64, g->ow, Hp, g->n, 8, Wp * 8, Hp * Wp * 8};
}

static int rowpatch_check(const msig_conv_geom* g, const char* what) {
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(g->stride == 1 && g->s >= 1 && g->s <= 8 && g->r >= 1 && g->r <= kMaxTaps && g->pad_t >= 0 &&
                   g->pad_l >= 0 && g->oh == g->h + 2 * g->pad_t - g->r + 1 && g->ow == g->w + 2 * g->pad_l - g->s + 1,
               "%s: needs stride 1, s <= 8 and oh/ow = h/w + 2*pad - r/s + 1", what);
  return MSIG_OK;
}

int msig_conv_rowpatch_fwd(const msig_conv_geom* g, const void* x_pad8, const void* w_rowpatch,
                           const msig_epilogue* e, void* y, void* stream) {
  MSIG_REQUIRE(g && x_pad8 && w_rowpatch && y, "msig_conv_rowpatch_fwd: null argument");
  int rc;
  if ((rc = rowpatch_check(g, "msig_conv_rowpatch_fwd")) != MSIG_OK) return rc;
  FpropParams p;
  init_fprop(p);
  const int k_pad = pad_rows(g->k);
  const int block_n = pick_block_n(k_pad);
  pick_tile_any(g->oh, g->ow, p.TW, p.TH);
  p.OH = g->oh; p.OW = g->ow;
  p.tiles_h = static_cast<int>(ceil_div(g->oh, p.TH));
  p.tiles_w = static_cast<int>(ceil_div(g->ow, p.TW));
  p.n_img = g->n;
  p.n_blocks = k_pad / block_n;
  p.taps = g->r;
  p.cblocks = 1;
  if ((rc = make_act_map(&p.tmA[0], pad8_view(x_pad8, g), p.TW, p.TH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  for (int r = 0; r < g->r; ++r) p.tap[r] = Tap{int8_t(r), 0, 0, 0};
  if ((rc = make_w_map(&p.tmB, w_rowpatch, k_pad, int64_t(g->r) * 64, fprop_uses_pairs(p, block_n) ? 128 : block_n)) !=
      MSIG_OK)
    return rc;
  const OutView ov = make_out_view(y, e ? e->out_layout : MSIG_OUT_BF16_NHWC, g->oh, g->ow, g->k);
  if ((rc = fill_epilogue(p, e, ov, g->k)) != MSIG_OK) return rc;
  if (g_ring_mode != 0 && k_pad == 64 && g->r <= 7 && g->ow >= 128) {
    // resident filter + one overlapped-window strip per padded image row, shared by R output rows
    set_ring(p, g->r, 1, 0, 0);
    if ((rc = make_act_map(&p.tmA[1], pad8_view(x_pad8, g), 128, 1)) != MSIG_OK) return rc;
    cudaError_t ce = launch_fprop_ring64(p, sm_count(), static_cast<cudaStream_t>(stream));
    if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(rowpatch ring) launch: %s", cudaGetErrorString(ce));
    return MSIG_OK;
  }
  cudaError_t ce = launch_fprop(p, block_n, sm_count(), static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "fprop(rowpatch) launch: %s", cudaGetErrorString(ce));
  return MSIG_OK;
}

// Row-patch weight gradients: groups of four filter rows are the four B boxes of one CTA. With exactly two
// groups (R = 5..8, the 7x7 convs of the generator) the groups are STACKED in M: the second 64-row A box is the
// `other` tile shifted up by four rows, so rows 64..127 of the same MMA accumulate filter rows 4..7 against the
// same B boxes (sum_y o[y-4] * img[y + r] == sum_y' o[y'] * img[y' + 4 + r]); the K range grows by four rows.
struct RowpatchWgradPlan {
  int groups, stacked, taps, M, PW, PH, blocks_w, blocks_h;
  int64_t kb_total;
  WgradPlan pl;
};
static RowpatchWgradPlan plan_rowpatch_wgrad(const msig_conv_geom* g) {
  RowpatchWgradPlan r;
  pick_kblock(g->ow, r.PW, r.PH);
  r.groups = static_cast<int>(ceil_div(g->r, 4));
  r.stacked = (r.groups == 2 && g_rowpatch_stack) ? 1 : 0;
  r.taps = r.stacked ? 1 : r.groups;
  r.M = r.stacked ? 128 : 64;
  r.blocks_w = static_cast<int>(ceil_div(g->ow, r.PW));
  r.blocks_h = static_cast<int>(ceil_div(g->oh + (r.stacked ? 4 : 0), r.PH));
  r.kb_total = int64_t(g->n) * r.blocks_h * r.blocks_w;
  r.pl = plan_wgrad(r.M, 256, r.taps, r.kb_total);
  return r;
}

size_t msig_conv_rowpatch_wgrad_workspace(const msig_conv_geom* g) {
  if (!g) return 0;
  const RowpatchWgradPlan r = plan_rowpatch_wgrad(g);
  return size_t(r.pl.splits) * 64 * r.groups * 256 * sizeof(float);
}

// flip = 0: dw[k][c][r][s] (+)= sum_pix dy[pix, k] * patch_r[pix, (s, c)]         (k = 64 output channels of
//           a conv over the pad8 image; `other` = dy, its bf16 NHWC output gradient [n, oh, ow, 64])
// flip = 1: the image is a few-channel GRADIENT (pad8 of dz, pad = R-1) and `other` is the 64-channel
//           input xp [n, oh, ow, 64] of a small-O conv: dw[o][c][r][s] (+)= sum_q xp[q, c] * dzpatch_r'[q, (s', o)],
//           r' = R-1-r, s' = S-1-s   (the weight gradient of the generator's final conv, model.py:141)
// GEMM: M = the 64 channels of `other` (one A box; the second half of the 128-row tile is never
// stored), N = 256 = FOUR filter rows' (s, c) windows (one B box per filter row), K = pixels.
int msig_conv_rowpatch_wgrad(const msig_conv_geom* g, const void* x_pad8, const void* other, int flip, float* dw,
                             int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(g && x_pad8 && other && dw && workspace, "msig_conv_rowpatch_wgrad: null argument");
  int rc;
  if ((rc = rowpatch_check(g, "msig_conv_rowpatch_wgrad")) != MSIG_OK) return rc;
  MSIG_REQUIRE(g->k == 64, "msig_conv_rowpatch_wgrad: the wide side must have 64 channels (k=%d)", g->k);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  const RowpatchWgradPlan rp = plan_rowpatch_wgrad(g);
  const WgradPlan& pl = rp.pl;
  p.PW = rp.PW; p.PH = rp.PH;
  p.blocks_w = rp.blocks_w; p.blocks_h = rp.blocks_h;
  p.n_img = g->n;
  p.taps = rp.taps;                                     // CTA taps: groups of four filter rows (stacked: both groups)
  MSIG_REQUIRE(rp.groups * 4 <= kMaxTaps, "rowpatch wgrad: too many filter rows");
  const size_t need = size_t(pl.splits) * 64 * rp.groups * 256 * sizeof(float);
  MSIG_REQUIRE(workspace_bytes >= need, "rowpatch wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.kb_total = pl.kb_total;
  p.out = reinterpret_cast<float*>(workspace);
  // partial layout [group][64 rows][256]: stacked rows 64..127 of the single CTA tap ARE group 1
  p.o_row = 256; p.o_tap = int64_t(rp.M) * 256; p.o_split = int64_t(rp.groups) * 64 * 256;
  p.alpha = 1.f; p.m_valid = rp.M; p.n_valid = 256;
  p.a_boxes = 1; p.b_box_tap = 1; p.a_box_tap = rp.stacked;
  ActView vo{other, 64, g->ow, g->oh, g->n, 64, int64_t(g->ow) * 64, int64_t(g->oh) * g->ow * 64};
  if ((rc = make_act_map(&p.tmA[0], vo, p.PW, p.PH)) != MSIG_OK) return rc;
  if ((rc = make_act_map(&p.tmB[0], pad8_view(x_pad8, g), p.PW, p.PH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) { p.tmA[i] = p.tmA[0]; p.tmB[i] = p.tmB[0]; }
  if (rp.stacked) {
    p.tapA[0] = Tap{0, 0, 0, 0};
    p.tapA[1] = Tap{-4, 0, 0, 0};                       // rows above / below the plane are zero-filled by TMA
  } else {
    for (int y = 0; y < p.taps; ++y) p.tapA[y] = Tap{0, 0, 0, 0};
  }
  // filter rows past R read whatever lies below (or TMA zero fill); their columns are never unpacked
  for (int r = 0; r < p.taps * 4; ++r) p.tapB[r] = Tap{int8_t(r), 0, 0, 0};
  cudaError_t ce = launch_wgrad(p, pl.block_n, st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "rowpatch wgrad launch: %s", cudaGetErrorString(ce));
  msig_wpack_desc d{flip ? MSIG_WPACK_ROWPATCH_FLIP : MSIG_WPACK_ROWPATCH, flip ? g->c : 64, flip ? 64 : g->c, g->r, g->s};
  PackGeom pg = make_pack_geom(&d, 0, 0);
  return launch_wgrad_reduce(pg, p.out, pl.splits, p.o_split, dw, accumulate, st);
}

// Narrow-output stride-1 conv through the row-fold kernel (see RowfoldParams).
int msig_conv_narrow_fwd(const msig_conv_geom* g, const void* x, const void* w_rowfold, const msig_epilogue* e,
                         void* y, void* stream) {
  MSIG_REQUIRE(g && x && w_rowfold && y, "msig_conv_narrow_fwd: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(g->c == 64 && g->k >= 1 && g->k <= 4 && g->stride == 1 && g->r >= 1 && g->r <= 7 && g->s >= 1 &&
                   g->s <= 8,
               "msig_conv_narrow_fwd: needs c = 64, k <= 4, stride 1, r <= 7, s <= 8 (got c=%d k=%d stride=%d %dx%d)",
               g->c, g->k, g->stride, g->r, g->s);
  const int layout = e ? e->out_layout : MSIG_OUT_F32_NCHW;
  MSIG_REQUIRE(layout == MSIG_OUT_F32_NCHW || layout == MSIG_OUT_F32_NHWC, "msig_conv_narrow_fwd: fp32 outputs only");
  MSIG_REQUIRE(!e || e->aux == nullptr, "msig_conv_narrow_fwd: no aux operand");
  RowfoldParams p;
  memset(&p, 0, sizeof(p));
  p.R = g->r; p.S = g->s;
  p.org_h = -g->pad_t; p.org_w = -g->pad_l;
  p.OH = g->oh; p.OW = g->ow; p.n_img = g->n;
  const int tile_out = 128 - (g->s - 1);
  p.tiles_w = static_cast<int>(ceil_div(g->ow, tile_out));
  // chunk of output rows per work item: long enough to amortise the R-1 extra strips, short enough
  // for several waves of items over the SMs
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int rows = 64;
  while (rows > 8 && int64_t(g->n) * p.tiles_w * ceil_div(g->oh, rows) < int64_t(6) * sms) rows /= 2;
  p.rows_per_item = rows;
  p.chunks_h = static_cast<int>(ceil_div(g->oh, rows));
  int rc;
  ActView v{x, g->c, g->w, g->h, g->n, g->c, int64_t(g->w) * g->c, int64_t(g->h) * g->w * g->c};
  if ((rc = make_act_map(&p.tmA, v, 128, 1)) != MSIG_OK) return rc;
  if ((rc = make_w_map(&p.tmB, w_rowfold, int64_t(g->r) * 32, 64, 32)) != MSIG_OK) return rc;
  const OutView ov = make_out_view(y, layout, g->oh, g->ow, g->k);
  p.out = reinterpret_cast<float*>(y);
  p.o_sn = ov.sN; p.o_sh = ov.sH; p.o_sw = ov.sW; p.o_sc = ov.sC;
  p.n_valid = g->k;
  p.bias = e ? e->bias : nullptr;
  p.alpha = e ? e->alpha : 1.f;
  p.alpha_ptr = e ? e->alpha_ptr : nullptr;
  p.act = e ? e->act : ACT_NONE;
  p.ch_scale = e ? e->ch_scale : nullptr;
  cudaError_t ce = launch_rowfold(p, sms, static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "rowfold launch: %s", cudaGetErrorString(ce));
  return MSIG_OK;
}

int msig_conv2d_dgrad(const msig_conv_geom* g, const void* dy, const void* w_dgrad,
                      const msig_epilogue* e, void* dx, void* stream) {
  MSIG_REQUIRE(g && dy && w_dgrad && dx, "msig_conv2d_dgrad: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g->stride == 1) {
    // dx[ih] = sum_r dy[ih + pad - r] w[r]  ==  conv of dy with the flipped filter, pad' = R-1-pad
    return run_conv(dy, g->n, g->oh, g->ow, g->k, g->c, g->r, g->s, 1, g->r - 1 - g->pad_t,
                    g->s - 1 - g->pad_l, g->h, g->w, w_dgrad, e, dx, st);
  }
  MSIG_REQUIRE(g->stride == 2 && g->r == 4 && g->s == 4 && g->pad_t == 1 && g->pad_l == 1 &&
                   g->h == 2 * g->oh && g->w == 2 * g->ow,
               "msig_conv2d_dgrad: stride-2 dgrad supports k=4,s=2,p=1 only");
  return run_phased(dy, g->n, g->oh, g->ow, g->k, g->c, w_dgrad, e, dx, st);
}

int msig_convT2d_fwd(const msig_conv_geom* g, const void* x, const void* w, const msig_epilogue* e,
                     void* y, void* stream) {
  MSIG_REQUIRE(g && x && w && y, "msig_convT2d_fwd: null argument");
  MSIG_REQUIRE(g->r == 4 && g->s == 4 && g->stride == 2 && g->oh == 2 * g->h && g->ow == 2 * g->w,
               "msig_convT2d_fwd: supports k=4,s=2,p=1 only");
  return run_phased(x, g->n, g->h, g->w, g->c, g->k, w, e, y, static_cast<cudaStream_t>(stream));
}

int msig_convT2d_dgrad(const msig_conv_geom* g, const void* dy, const void* w, const msig_epilogue* e,
                       void* dx, void* stream) {
  MSIG_REQUIRE(g && dy && w && dx, "msig_convT2d_dgrad: null argument");
  // dx[ih, ci] = sum_{r,co} dy[2 ih - 1 + r, co] wT[ci][co][r]  ==  4x4 stride-2 pad-1 conv of dy
  return run_conv(dy, g->n, g->oh, g->ow, g->k, g->c, 4, 4, 2, 1, 1, g->h, g->w, w, e, dx,
                  static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------- weight gradients
static size_t wgrad_ws_bytes(int M, int N, int taps, int64_t kb_total) {
  WgradPlan pl = plan_wgrad(M, N, taps, kb_total);
  return size_t(pl.splits) * size_t(M) * taps * N * sizeof(float);
}

// 64-input-channel convs (model.py:132 and the second layer of the SE / D trunks): dy is the 128-wide M
// operand and FOUR filter taps share one CTA as four 64-channel B boxes (N = 256), so the dy tile is
// fetched once per four taps instead of once per tap (these layers sit on the L2 -> shared-memory roof).
static bool wgrad_groups_taps(const msig_conv_geom* g) { return g->c == 64 && (g->r * g->s) % 4 == 0; }

size_t msig_conv2d_wgrad_workspace(const msig_conv_geom* g) {
  if (!g) return 0;
  int PW, PH;
  pick_kblock(g->ow, PW, PH);
  const int64_t kb = int64_t(g->n) * ceil_div(g->oh, PH) * ceil_div(g->ow, PW);
  if (wgrad_groups_taps(g)) return wgrad_ws_bytes(g->k, 256, g->r * g->s / 4, kb);
  const bool swap = g->c >= 128;
  return wgrad_ws_bytes(swap ? g->c : g->k, swap ? g->k : g->c, g->r * g->s, kb);
}

int msig_conv2d_wgrad(const msig_conv_geom* g, const void* x, const void* dy, float* dw, int accumulate,
                      void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(g && x && dy && dw && workspace, "msig_conv2d_wgrad: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(g->c % 64 == 0 && g->k % 64 == 0, "wgrad: c (%d) and k (%d) must be multiples of 64", g->c, g->k);
  MSIG_REQUIRE(g->r * g->s <= kMaxTaps, "wgrad: too many taps");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  pick_kblock(g->ow, p.PW, p.PH);
  p.blocks_w = static_cast<int>(ceil_div(g->ow, p.PW));
  p.blocks_h = static_cast<int>(ceil_div(g->oh, p.PH));
  p.n_img = g->n;
  const int taps = g->r * g->s;
  const bool group = wgrad_groups_taps(g);
  p.taps = group ? taps / 4 : taps;
  const int64_t kb_total = int64_t(g->n) * p.blocks_h * p.blocks_w;
  // Operand roles. All CTAs of one pixel range read the SAME dy tile (L2 serves it once) but a
  // different shifted x window per tap, so the per-CTA-unique operand should be the 128-wide M side:
  // with >= 128 input channels x is the M operand (partials come out [ci][tap-major][co]).
  const bool swap = !group && g->c >= 128;
  const int M = swap ? g->c : g->k, N = group ? 256 : (swap ? g->k : g->c);
  WgradPlan pl = plan_wgrad(M, N, p.taps, kb_total);
  const size_t need = size_t(pl.splits) * size_t(g->k) * taps * g->c * sizeof(float);
  MSIG_REQUIRE(workspace_bytes >= need, "wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.kb_total = pl.kb_total;
  p.out = reinterpret_cast<float*>(workspace);
  // partial layout [m][tap][n]: with grouped taps a CTA's 256 columns are 4 consecutive taps x 64 channels
  p.o_row = int64_t(taps) * (group ? 64 : N); p.o_tap = N;
  p.o_split = int64_t(g->k) * taps * g->c;
  p.alpha = 1.f; p.m_valid = M; p.n_valid = N;
  p.b_box_tap = group ? 1 : 0;
  int rc;
  CUtensorMap* tm_dy = swap ? p.tmB : p.tmA;
  CUtensorMap* tm_x = swap ? p.tmA : p.tmB;
  Tap* tap_dy = swap ? p.tapB : p.tapA;
  Tap* tap_x = swap ? p.tapA : p.tapB;
  ActView va{dy, g->k, g->ow, g->oh, g->n, g->k, int64_t(g->ow) * g->k, int64_t(g->oh) * g->ow * g->k};
  if ((rc = make_act_map(&tm_dy[0], va, p.PW, p.PH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) tm_dy[i] = tm_dy[0];
  for (int t = 0; t < taps; ++t) tap_dy[t] = Tap{0, 0, 0, 0};
  if (g->stride == 1) {
    ActView vb{x, g->c, g->w, g->h, g->n, g->c, int64_t(g->w) * g->c, int64_t(g->h) * g->w * g->c};
    if ((rc = make_act_map(&tm_x[0], vb, p.PW, p.PH)) != MSIG_OK) return rc;
    for (int i = 1; i < 4; ++i) tm_x[i] = tm_x[0];
    for (int r = 0; r < g->r; ++r)
      for (int s = 0; s < g->s; ++s)
        tap_x[r * g->s + s] = Tap{int8_t(r - g->pad_t), int8_t(s - g->pad_l), 0, 0};
  } else {
    MSIG_REQUIRE(g->stride == 2 && g->h % 2 == 0 && g->w % 2 == 0, "wgrad: stride-2 needs even dims");
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const __nv_bfloat16* base =
            reinterpret_cast<const __nv_bfloat16*>(x) + (int64_t(ph) * g->w + pw) * g->c;
        ActView vb{base, g->c, g->w / 2, g->h / 2, g->n, int64_t(2) * g->c, int64_t(2) * g->w * g->c,
                   int64_t(g->h) * g->w * g->c};
        if ((rc = make_act_map(&tm_x[ph * 2 + pw], vb, p.PW, p.PH)) != MSIG_OK) return rc;
      }
    for (int r = 0; r < g->r; ++r)
      for (int s = 0; s < g->s; ++s) {
        const int rr = r - g->pad_t, ss = s - g->pad_l;
        const int ph = ((rr % 2) + 2) % 2, pw = ((ss % 2) + 2) % 2;
        tap_x[r * g->s + s] = Tap{int8_t((rr - ph) / 2), int8_t((ss - pw) / 2), int8_t(ph * 2 + pw), 0};
      }
  }
  cudaError_t ce = launch_wgrad(p, pl.block_n, st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "wgrad launch: %s", cudaGetErrorString(ce));
  msig_wpack_desc d{MSIG_WPACK_FWD, g->k, g->c, g->r, g->s};
  PackGeom pg = make_pack_geom(&d, 0, 0);
  pg.partT = swap ? 1 : 0;
  return launch_wgrad_reduce(pg, p.out, pl.splits, p.o_split, dw, accumulate, st);
}

// 64 output channels and >= 128 input channels (model.py:138, the generator's second up-sampling layer): x is
// the 128-wide M operand and the FOUR taps of one output phase share a CTA as four 64-channel dy boxes (N = 256),
// sum_p x[p + d] dy_ph[p] = sum_q x[q] dy_ph[q - d]: the shift moves to the dy side (TMA zero fill either way).
static bool convT_wgrad_groups(const msig_conv_geom* g) { return g_convt_group && g->k == 64 && g->c % 128 == 0; }

size_t msig_convT2d_wgrad_workspace(const msig_conv_geom* g) {
  if (!g) return 0;
  int PW, PH;
  pick_kblock(g->w, PW, PH);
  const int64_t kb = int64_t(g->n) * ceil_div(g->h, PH) * ceil_div(g->w, PW);
  if (convT_wgrad_groups(g)) return size_t(plan_wgrad(g->c, 256, 4, kb).splits) * g->k * 16 * g->c * sizeof(float);
  return wgrad_ws_bytes(g->k, g->c, 16, kb);
}

// Grouped-tap convT partials [ci][T = phase*4 + tap][co] -> dw[ci][co][r][s] (+)=: one block per input channel,
// (splits x) reads coalesced along co, one coalesced write of the 16*O contiguous master elements.
__global__ void __launch_bounds__(256) wgrad_reduce_convT_t_kernel(const float* __restrict__ partial, int splits,
                                                                  int64_t split_stride, float* __restrict__ dw,
                                                                  int accumulate, int O) {
  pdl_entry();
  extern __shared__ float sm_t[];                       // [O][17]
  const int i = blockIdx.x;
  const float* pp = partial + int64_t(i) * 16 * O;
  for (int q = threadIdx.x; q < 16 * O; q += 256) {
    const int T = q / O, o = q - T * O;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += pp[s * split_stride + q];
    const int ph = T >> 2, tp = T & 3;
    const int t = ph_r(ph >> 1, tp >> 1) * 4 + ph_r(ph & 1, tp & 1);
    sm_t[o * 17 + t] = acc;
  }
  __syncthreads();
  float* d = dw + int64_t(i) * 16 * O;
  for (int e = threadIdx.x; e < 16 * O; e += 256) {
    const float v = sm_t[(e >> 4) * 17 + (e & 15)];
    d[e] = accumulate ? d[e] + v : v;
  }
}

// dW[ci][co][r][s] = sum x[n, i+dh, j+dw, ci] * dy[n, 2i+py, 2j+px, co] over the 4 phases x 4 taps.
int msig_convT2d_wgrad(const msig_conv_geom* g, const void* x, const void* dy, float* dw, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(g && x && dy && dw && workspace, "msig_convT2d_wgrad: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(g->c % 64 == 0 && g->k % 64 == 0, "convT wgrad: c and k must be multiples of 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  pick_kblock(g->w, p.PW, p.PH);
  p.blocks_w = static_cast<int>(ceil_div(g->w, p.PW));
  p.blocks_h = static_cast<int>(ceil_div(g->h, p.PH));
  p.n_img = g->n;
  const bool group = convT_wgrad_groups(g);
  p.taps = group ? 4 : 16;
  const int64_t kb_total = int64_t(g->n) * p.blocks_h * p.blocks_w;
  WgradPlan pl = group ? plan_wgrad(g->c, 256, 4, kb_total) : plan_wgrad(g->k, g->c, 16, kb_total);
  const size_t need = size_t(pl.splits) * size_t(g->k) * 16 * g->c * sizeof(float);
  MSIG_REQUIRE(workspace_bytes >= need, "convT wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.kb_total = pl.kb_total;
  p.out = reinterpret_cast<float*>(workspace);
  p.o_split = int64_t(g->k) * 16 * g->c;
  p.alpha = 1.f;
  if (group) {   // partials [ci][T][co]: the CTA of phase ph owns columns [ph*256, ph*256 + 256) of each row
    p.o_row = int64_t(16) * g->k; p.o_tap = 256;
    p.m_valid = g->c; p.n_valid = 256;
    p.b_box_tap = 1;
  } else {       // partials [co][T][ci]
    p.o_row = int64_t(16) * g->c; p.o_tap = g->c;
    p.m_valid = g->k; p.n_valid = g->c;
  }
  int rc;
  CUtensorMap* tm_dy = group ? p.tmB : p.tmA;
  CUtensorMap* tm_x = group ? p.tmA : p.tmB;
  Tap* tap_dy = group ? p.tapB : p.tapA;
  Tap* tap_x = group ? p.tapA : p.tapB;
  const int64_t OH = 2 * int64_t(g->h), OW = 2 * int64_t(g->w);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(dy) + (py * OW + px) * g->k;
      ActView va{base, g->k, g->w, g->h, g->n, int64_t(2) * g->k, 2 * OW * g->k, OH * OW * g->k};
      if ((rc = make_act_map(&tm_dy[py * 2 + px], va, p.PW, p.PH)) != MSIG_OK) return rc;
      for (int ty = 0; ty < 2; ++ty)
        for (int tx = 0; tx < 2; ++tx) {
          const int T = (py * 2 + px) * 4 + ty * 2 + tx;
          const int dh = ph_d(py, ty), dw_ = ph_d(px, tx);
          if (group) {
            tap_dy[T] = Tap{int8_t(-dh), int8_t(-dw_), int8_t(py * 2 + px), 0};
          } else {
            tap_dy[T] = Tap{0, 0, int8_t(py * 2 + px), 0};
            tap_x[T] = Tap{int8_t(dh), int8_t(dw_), 0, 0};
          }
        }
    }
  if (group)
    for (int ph = 0; ph < 4; ++ph) tap_x[ph] = Tap{0, 0, 0, 0};
  ActView vb{x, g->c, g->w, g->h, g->n, g->c, int64_t(g->w) * g->c, int64_t(g->h) * g->w * g->c};
  if ((rc = make_act_map(&tm_x[0], vb, p.PW, p.PH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) tm_x[i] = tm_x[0];
  cudaError_t ce = launch_wgrad(p, pl.block_n, st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "convT wgrad launch: %s", cudaGetErrorString(ce));
  if (group) {
    MSIG_LAUNCH((wgrad_reduce_convT_t_kernel), g->c, 256, size_t(g->k) * 17 * sizeof(float), st, p.out, pl.splits, p.o_split, dw,
                                                                                      accumulate, g->k);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
    return MSIG_OK;
  }
  msig_wpack_desc d{MSIG_WPACK_CONVT_FWD, g->k, g->c, 4, 4};   // master weight [I=c][O=k][4][4]
  PackGeom pg = make_pack_geom(&d, 0, 0);
  return launch_wgrad_reduce(pg, p.out, pl.splits, p.o_split, dw, accumulate, st);
}

// Flat wgrad for gathered-patch GEMMs: a [rows][m] and b [rows][ncols] bf16 row-major,
// result[m][ncols] = a^T b, un-packed into the master weight layout described by `d`.
size_t msig_patch_wgrad_workspace(int64_t rows, int32_t m, int32_t ncols) {
  return wgrad_ws_bytes(m, ncols, 1, ceil_div(rows, 64));
}

int msig_gemm_tn_partial(int64_t rows, const void* a_rows_m, int32_t m, const void* b_rows_n, int32_t ncols,
                         void* workspace, size_t workspace_bytes, int32_t* splits_out, void* stream) {
  MSIG_REQUIRE(a_rows_m && b_rows_n && workspace && splits_out, "msig_gemm_tn_partial: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(m % 64 == 0 && ncols % 64 == 0, "gemm_tn: m (%d) and ncols (%d) must be multiples of 64", m, ncols);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.PW = 64; p.PH = 1;
  p.blocks_w = static_cast<int>(ceil_div(rows, 64));
  p.blocks_h = 1;
  p.n_img = 1;
  p.taps = 1;
  const int64_t kb_total = p.blocks_w;
  WgradPlan pl = plan_wgrad(m, ncols, 1, kb_total);
  const size_t need = size_t(pl.splits) * size_t(m) * ncols * sizeof(float);
  MSIG_REQUIRE(workspace_bytes >= need, "gemm_tn: workspace too small (%zu < %zu)", workspace_bytes, need);
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.kb_total = pl.kb_total;
  p.out = reinterpret_cast<float*>(workspace);
  p.o_row = ncols; p.o_tap = 0; p.o_split = int64_t(m) * ncols;
  p.alpha = 1.f; p.m_valid = m; p.n_valid = ncols;
  int rc;
  ActView va{a_rows_m, m, rows, 1, 1, m, int64_t(rows) * m, int64_t(rows) * m};
  if ((rc = make_act_map(&p.tmA[0], va, 64, 1)) != MSIG_OK) return rc;
  ActView vb{b_rows_n, ncols, rows, 1, 1, ncols, int64_t(rows) * ncols, int64_t(rows) * ncols};
  if ((rc = make_act_map(&p.tmB[0], vb, 64, 1)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) { p.tmA[i] = p.tmA[0]; p.tmB[i] = p.tmB[0]; }
  cudaError_t ce = launch_wgrad(p, pl.block_n, st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "gemm_tn launch: %s", cudaGetErrorString(ce));
  *splits_out = pl.splits;
  return MSIG_OK;
}

// partial layouts: IM2COL [O][Kpad] (a = dy, b = patches); IM2COL_FLIP [Kpad][I] (a = patches of dy,
// b = x); FWD with r=s=1 [O][I] (Linear; a = dy, b = x).
int msig_wgrad_unpack(const msig_wpack_desc* d, int32_t oc, int32_t o_off, const float* partial,
                      int32_t splits, int64_t split_stride, float* dw, int accumulate, void* stream) {
  MSIG_REQUIRE(d && partial && dw && splits >= 1, "msig_wgrad_unpack: bad argument");
  PackGeom pg = make_pack_geom(d, oc, o_off);
  MSIG_REQUIRE(pg.kind == MSIG_WPACK_FWD || pg.kind == MSIG_WPACK_IM2COL || pg.kind == MSIG_WPACK_IM2COL_FLIP ||
                   pg.kind == MSIG_WPACK_CONVT_FWD || pg.kind == MSIG_WPACK_ROWPATCH ||
                   pg.kind == MSIG_WPACK_ROWPATCH_FLIP,
               "msig_wgrad_unpack: kind %d has no weight-gradient layout", pg.kind);
  return launch_wgrad_reduce(pg, partial, splits, split_stride, dw, accumulate, static_cast<cudaStream_t>(stream));
}

// Weight AND bias gradients of `layers` Linear layers that were evaluated as one batched GEMM (the 16 AdaIN
// style Linears of a generator, model.py:18,28; the per-domain 1x1 heads of the style encoder, model.py:84),
// in ONE launch: dW_l[o][i] += sum_splits partial[split][(l*O + o)][i] and db_l[o] += sum_rows dy[row][l*O + o],
// scattered to the layers' own gradient tensors through device-resident pointer tables. Replaces one
// msig_wgrad_unpack + one msig_colsum_f32 per layer (2 x 16 launches per generator backward).
__global__ void __launch_bounds__(256) multi_linear_grads_kernel(const float* __restrict__ partial, int splits,
                                                                 int64_t split_stride, const float* __restrict__ dy,
                                                                 int64_t rows, int64_t ld, int layers, int O, int I,
                                                                 float* const* __restrict__ wgrads,
                                                                 float* const* __restrict__ bgrads) {
  pdl_entry();
  const int64_t per_layer = int64_t(O) * I;
  const int64_t n_w = per_layer * layers, n_b = int64_t(O) * layers;
  for (int64_t q = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; q < n_w + n_b;
       q += int64_t(gridDim.x) * blockDim.x) {
    if (q < n_w) {
      const int l = static_cast<int>(q / per_layer);
      float acc = 0.f;
      for (int s = 0; s < splits; ++s) acc += partial[s * split_stride + q];
      float* d = wgrads[l] + (q - l * per_layer);
      *d += acc;
    } else {
      const int64_t c = q - n_w;                       // column l*O + o of dy
      const int l = static_cast<int>(c / O);
      float acc = 0.f;
      for (int64_t r = 0; r < rows; ++r) acc += dy[r * ld + c];
      float* d = bgrads[l] + (c - int64_t(l) * O);
      *d += acc;
    }
  }
}

int msig_multi_linear_grads(const float* partial, int32_t splits, int64_t split_stride, const float* dy,
                            int64_t rows, int64_t ld, int32_t layers, int32_t out_features, int32_t in_features,
                            const void* wgrad_ptrs, const void* bgrad_ptrs, void* stream) {
  MSIG_REQUIRE(partial && dy && wgrad_ptrs && bgrad_ptrs && splits >= 1 && layers >= 1 && out_features >= 1 &&
                   in_features >= 1 && rows >= 1,
               "msig_multi_linear_grads: bad argument");
  const int64_t total = int64_t(layers) * out_features * (int64_t(in_features) + 1);
  const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(total, 256), 148 * 16));
  MSIG_LAUNCH((multi_linear_grads_kernel), blocks, 256, 0, static_cast<cudaStream_t>(stream), 
      partial, splits, split_stride, dy, rows, ld, layers, out_features, in_features,
      reinterpret_cast<float* const*>(wgrad_ptrs), reinterpret_cast<float* const*>(bgrad_ptrs));
  count_launch(1);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

int msig_patch_wgrad_part(const msig_wpack_desc* d, int32_t oc, int32_t o_off, int64_t rows,
                          const void* a_rows_m, int32_t m, const void* b_rows_n, int32_t ncols, float* dw,
                          int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(d && dw, "msig_patch_wgrad: null argument");
  PackGeom pg = make_pack_geom(d, oc, o_off);
  if (pg.kind == MSIG_WPACK_IM2COL) MSIG_REQUIRE(ncols == pg.Kpad && m >= pg.O, "patch wgrad: IM2COL shape mismatch");
  if (pg.kind == MSIG_WPACK_IM2COL_FLIP) MSIG_REQUIRE(m == pg.Kpad && ncols == pg.I, "patch wgrad: FLIP shape mismatch");
  if (pg.kind == MSIG_WPACK_FWD) MSIG_REQUIRE(pg.RS == 1 && ncols == pg.I, "patch wgrad: FWD needs r=s=1");
  int32_t splits = 0;
  int rc = msig_gemm_tn_partial(rows, a_rows_m, m, b_rows_n, ncols, workspace, workspace_bytes, &splits, stream);
  if (rc != MSIG_OK) return rc;
  return msig_wgrad_unpack(d, oc, o_off, reinterpret_cast<const float*>(workspace), splits, int64_t(m) * ncols,
                           dw, accumulate, stream);
}

int msig_patch_wgrad(const msig_wpack_desc* d, int64_t rows, const void* a_rows_m, const void* b_rows_n,
                     float* dw, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(d, "msig_patch_wgrad: null desc");
  PackGeom pg = make_pack_geom(d, 0, 0);
  int m, ncols;
  if (pg.kind == MSIG_WPACK_IM2COL) { m = static_cast<int>(round_up(pg.O, 64)); ncols = pg.Kpad; }
  else if (pg.kind == MSIG_WPACK_IM2COL_FLIP) { m = pg.Kpad; ncols = pg.I; }
  else { m = static_cast<int>(round_up(pg.O, 64)); ncols = pg.I; }
  return msig_patch_wgrad_part(d, 0, 0, rows, a_rows_m, m, b_rows_n, ncols, dw, accumulate, workspace,
                               workspace_bytes, stream);
}


// ---------------------------------------------------------------- Gram matrices (losses.py:70-78)
static WgradPlan plan_gram(int n, int h, int w, int c, int& PW, int& PH) {
  pick_kblock(w, PW, PH);
  const int64_t kb = ceil_div(h, PH) * ceil_div(w, PW);
  return plan_wgrad(n * c, n * c, 1, kb, true);
}

size_t msig_gram_workspace(int32_t n, int32_t h, int32_t w, int32_t c) {
  int PW, PH;
  WgradPlan pl = plan_gram(n, h, w, c, PW, PH);
  return pl.splits > 1 ? size_t(pl.splits) * size_t(n) * c * size_t(n) * c * sizeof(float) : 16;
}

int msig_gram_fwd(const void* f, int32_t n, int32_t h, int32_t w, int32_t c, float* gram, void* workspace,
                  size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(f && gram, "msig_gram_fwd: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(c % 64 == 0, "msig_gram_fwd: channels (%d) must be a multiple of 64", c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  WgradPlan pl = plan_gram(n, h, w, c, p.PW, p.PH);
  const int dim = n * c;
  p.blocks_w = static_cast<int>(ceil_div(w, p.PW));
  p.blocks_h = static_cast<int>(ceil_div(h, p.PH));
  p.n_img = n; p.taps = 1; p.fold_img = 1; p.CA = c; p.CB = c;
  p.m_blocks = pl.m_blocks; p.n_blocks = pl.n_blocks; p.splits = pl.splits;
  p.kb_per_split = pl.kb_per_split; p.kb_total = pl.kb_total;
  p.o_row = dim; p.o_tap = 0; p.o_split = int64_t(dim) * dim;
  p.alpha = 1.f / (float(n) * float(c) * float(h) * float(w));
  p.m_valid = dim; p.n_valid = dim;
  p.upper_only = 1;   // G is symmetric: tiles strictly below the diagonal are neither computed nor read
  if (pl.splits > 1) {
    MSIG_REQUIRE(workspace && workspace_bytes >= size_t(pl.splits) * dim * size_t(dim) * sizeof(float),
                 "msig_gram_fwd: workspace too small");
    p.out = reinterpret_cast<float*>(workspace);
  } else {
    p.out = gram;
  }
  int rc;
  ActView v{f, c, w, h, n, c, int64_t(w) * c, int64_t(h) * w * c};
  if ((rc = make_act_map(&p.tmA[0], v, p.PW, p.PH)) != MSIG_OK) return rc;
  for (int i = 0; i < 4; ++i) { p.tmA[i] = p.tmA[0]; p.tmB[i] = p.tmA[0]; }
  cudaError_t ce = launch_wgrad(p, pl.block_n, st);
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "gram launch: %s", cudaGetErrorString(ce));
  if (pl.splits > 1) {
    const int64_t numel = int64_t(dim) * dim;
    MSIG_LAUNCH((sum_splits_kernel), static_cast<int>(std::min<int64_t>(ceil_div(numel, 256), 4096)), 256, 0, st, 
        p.out, pl.splits, p.o_split, gram, numel);
    count_launch(1);
    MSIG_CHECK_LAUNCH();
  }
  return MSIG_OK;
}

// df[b, p, c] = alpha * sum_{b', c'} ssym[(b,c)][(b',c')] * f[b', p, c']   (+ aux)
int msig_gram_bwd(const void* f, const void* ssym, int32_t n, int32_t h, int32_t w, int32_t c, float alpha,
                  const float* gscale, const void* aux, int relu_mask, void* df, void* stream) {
  MSIG_REQUIRE(f && ssym && df, "msig_gram_bwd: null argument");
  MSIG_REQUIRE(context_ready(), "msig_init() has not been called");
  MSIG_REQUIRE(c % 64 == 0, "msig_gram_bwd: channels (%d) must be a multiple of 64", c);
  FpropParams p;
  init_fprop(p);
  // GEMM view: rows = the h*w pixel positions, K = N = (image, channel) pairs. An output tile is 128
  // pixel positions x BLOCK_N columns that may span several images (fold_c), so narrow feature maps
  // (c = 64) still run 256-wide MMAs; the A tile of K block (image b', chunk) is shared by all of them.
  const int block_n = pick_block_n(n * c);
  pick_tile(w, p.TW, p.TH);
  p.OH = h; p.OW = w;
  p.tiles_h = static_cast<int>(ceil_div(h, p.TH));
  p.tiles_w = static_cast<int>(ceil_div(w, p.TW));
  p.n_img = 1;
  p.n_blocks = (n * c) / block_n;
  p.taps = n;                 // one "tap" per source image
  p.cblocks = c / 64;
  p.tap_is_image = 1;
  p.b_row_per_image = 0;
  p.fold_c = c;
  int rc;
  ActView v{f, c, w, h, n, c, int64_t(w) * c, int64_t(h) * w * c};
  if ((rc = make_act_map(&p.tmA[0], v, p.TW, p.TH)) != MSIG_OK) return rc;
  for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  const int64_t dim = int64_t(n) * c;
  if ((rc = make_w_map(&p.tmB, ssym, dim, dim, fprop_uses_pairs(p, block_n) ? 128 : block_n)) != MSIG_OK) return rc;
  msig_epilogue e;
  memset(&e, 0, sizeof(e));
  e.alpha = alpha; e.alpha_ptr = gscale; e.aux = aux; e.aux_mode = aux ? MSIG_AUX_ADD : MSIG_AUX_NONE;
  e.out_layout = MSIG_OUT_BF16_NHWC;
  const OutView ov = make_out_view(df, MSIG_OUT_BF16_NHWC, h, w, c);
  if ((rc = fill_epilogue(p, &e, ov, c)) != MSIG_OK) return rc;
  if (relu_mask) {            // df *= (f > 0): the ReLU backward of the tapped feature map, fused
    p.stat_z = reinterpret_cast<const __nv_bfloat16*>(f);
    p.z_mask = 1;
  }
  cudaError_t ce = launch_fprop(p, block_n, sm_count(), static_cast<cudaStream_t>(stream));
  if (ce != cudaSuccess) return set_error(MSIG_ERR_CUDA, "gram bwd launch: %s", cudaGetErrorString(ce));
  return MSIG_OK;
}

}  // extern "C"
