// tcgen05 implicit-GEMM kernels for sm_100a (TMA-staged operands, TMEM accumulators, mbarrier pipelines).
//
//   fprop_kernel<BLOCK_N>   persistent, warp-specialised (warp 0 TMA producer, warp 1 MMA issuer, warp 2
//                           TMEM allocator, warps 4..11 epilogue); the 128 x BLOCK_N fp32 accumulator is
//                           double buffered in TMEM so the epilogue of tile i overlaps the main loop of
//                           tile i+1. Epilogue: alpha / bias / residual / activation mask / activation,
//                           fused per-(image, channel) reductions, 256-bit loads and stores.
//   fprop2_kernel           the same GEMM for 256-wide tiles on a CTA pair (cluster of 2, cta_group::2):
//                           each CTA stages half of the weight tile (the 1-CTA kernel sits on the L2 roof).
//   fprop_m2_kernel         the 128-wide tiles with TWO m-tiles per CTA sharing one weight tile (same roof).
//   fprop_ring64_kernel     64-channel stride-1 layers: resident filter, strip ring shared by output rows,
//                           horizontal taps as row-shifted descriptors (also the row-patch 7x7 convs and
//                           the phases of the 128 -> 64 transposed conv / stride-2 dgrad).
//   fprop_rowfold_kernel    <= 4 output channels: horizontal taps folded into N, shift-add epilogue.
//   wgrad_kernel<BLOCK_N>   one 128 x BLOCK_N tile per CTA, K = pixels, both operands pixel-major NHWC tiles
//                           consumed as MN-major UMMA operands; split-K into fp32 partials.
//   wgrad2_kernel           the 256 x 256 tile on a CTA pair (cta_group::2).
//
// Reference ops these replace: every nn.Conv2d / nn.ConvTranspose2d / nn.Linear in
// /root/reference/model.py:18,45,48,72-75,84,131-141,165,183 and the VGG convs + torch.mm Gram of
// /root/reference/losses.py:15,76 (forward, dgrad, wgrad).
#include "common.h"
#include "igemm.cuh"
#include "ptx.cuh"

#include <atomic>

namespace msig {

static std::atomic<int> g_launches{0};
int igemm_kernel_launches() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

constexpr int kTileM = 128;
constexpr int kBlockK = 64;                     // bf16 elements per K block = one 128 B swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;   // 16 KiB

template <int BLOCK_N>
struct FpropCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  // Accumulator stages in TMEM: double buffered. (Eight stages + alternate-tile epilogue warpgroups for the 64-wide
  // tiles were tried and were SLOWER: those layers are bound by the epilogue's instruction stream, not by the
  // accumulator hand-off latency -- profiles/README_r2.md section 4.)
  static constexpr int kAccStages = 2;
  static constexpr bool kAltTiles = false;
  static constexpr int kTmemCols = (kAccStages * BLOCK_N) < 32 ? 32 : kAccStages * BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Column sums over the 32 lanes of a warp for 32 per-lane values (a 32x32 transpose-reduce:
// 31 shuffles instead of 160). On return lane l holds the total of column l in v[0].
__device__ __forceinline__ float warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Loads the epilogue's global operands for NC columns of one output pixel: `av` = aux (residual /
// activation mask source), `zv` = stat_z. Issued one chunk ahead of their use so their latency hides
// behind the previous chunk (and, for the first chunk, behind the tile's main loop).
// 32-byte read-only load: one full DRAM/L2 sector per thread per instruction. (With 16-byte loads a
// warp touches 32 half-used sectors per instruction and relies on L1 to merge the two halves; with the
// aux AND z streams in flight that merging breaks down and the L2 -> SM traffic of the epilogue doubles.)
__device__ __forceinline__ void ldg256(const void* ptr, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(ptr));
}

__device__ __forceinline__ void stg256(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z),
               "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

template <int NC>
__device__ __forceinline__ void epilogue_load(const FpropParams& p, int col0, bool row_valid, int64_t out_off,
                                              int64_t aux_off, uint4 (&av)[NC / 8], uint4 (&zv)[NC / 8]) {
#pragma unroll
  for (int g = 0; g < NC / 8; g += 2) {
    const int c = col0 + 8 * g;
    const bool ld_aux = p.aux_mode != AUX_NONE && p.mask_scale == nullptr;
    if (row_valid && c + 16 <= p.n_valid) {          // column pairs are 32-byte aligned (channels % 16 == 0)
      if (ld_aux) ldg256(p.aux + aux_off + c, av[g], av[g + 1]);
      if (p.stat_z != nullptr) ldg256(p.stat_z + out_off + c, zv[g], zv[g + 1]);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool live = row_valid && (c + 8 * h + 8 <= p.n_valid);
        if (live && ld_aux)
          av[g + h] = __ldg(reinterpret_cast<const uint4*>(p.aux + aux_off + c + 8 * h));
        if (live && p.stat_z != nullptr)
          zv[g + h] = __ldg(reinterpret_cast<const uint4*>(p.stat_z + out_off + c + 8 * h));
      }
    }
  }
}

// Epilogue for NC accumulator columns of one output pixel (thread-per-row). With STATS the warp
// also reduces the stored values over its 32 pixels (see FpropParams::stat_out); `stat_row` is the
// partial-sum row of this warp and `stat_col` the GEMM column of r[0].
// Epilogue for NC accumulator columns of one output pixel (thread-per-row). With STATS the warp
// also reduces the stored values over its 32 pixels (see FpropParams::stat_out); `stat_row` is the
// partial-sum row of this warp and `stat_col` the GEMM column of r[0].
template <int NC, bool STATS>
__device__ __forceinline__ void fprop_epilogue_chunk(const FpropParams& p, const uint32_t (&r)[NC],
                                                     const uint4 (&av)[NC / 8], const uint4 (&zv)[NC / 8],
                                                     int col0, bool row_valid, int64_t out_off, float alpha,
                                                     int64_t stat_row = 0, int stat_col = 0, int lane = 0,
                                                     const float* msc = nullptr, const float* msh = nullptr) {
  if (!STATS && !row_valid) return;
  if (!p.out_f32 && p.o_sc == 1) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    float sv[STATS ? NC : 1], sq[STATS ? NC : 1];
    uint4 o_even = make_uint4(0, 0, 0, 0);      // even group's packed output, stored with the odd one (32 bytes)
#pragma unroll
    for (int g = 0; g < NC / 8; ++g) {
      const int c = col0 + 8 * g;
      const bool live = row_valid && (c + 8 <= p.n_valid);
      if (STATS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sv[8 * g + j] = sq[8 * g + j] = 0.f;
      }
      if (!live) continue;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[8 * g + j]) * alpha;
      if (p.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (p.aux_mode != AUX_NONE && msc != nullptr) {
        // act' of the normalised value, recomputed exactly as the forward pass formed it (ops.cu, norm_act_fwd)
        const __nv_bfloat162* zh2 = reinterpret_cast<const __nv_bfloat162*>(&zv[g]);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(msc + c)), s1 = __ldg(reinterpret_cast<const float4*>(msc + c + 4));
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(msh + c)), h1 = __ldg(reinterpret_cast<const float4*>(msh + c + 4));
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float off = p.aux_mode == AUX_RELU_MASK ? 0.f : p.slope;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 z = __bfloat1622float2(zh2[j]);
          v[2 * j] = fmaf(z.x, sc[2 * j], sh[2 * j]) > 0.f ? v[2 * j] : v[2 * j] * off;
          v[2 * j + 1] = fmaf(z.y, sc[2 * j + 1], sh[2 * j + 1]) > 0.f ? v[2 * j + 1] : v[2 * j + 1] * off;
        }
      } else if (p.aux_mode != AUX_NONE) {
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&av[g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(ah[j]);
          if (p.aux_mode == AUX_ADD) {
            v[2 * j] += f.x;
            v[2 * j + 1] += f.y;
          } else if (p.aux_mode == AUX_RELU_MASK) {
            v[2 * j] = f.x > 0.f ? v[2 * j] : 0.f;
            v[2 * j + 1] = f.y > 0.f ? v[2 * j + 1] : 0.f;
          } else {
            v[2 * j] = f.x > 0.f ? v[2 * j] : v[2 * j] * p.slope;
            v[2 * j + 1] = f.y > 0.f ? v[2 * j + 1] : v[2 * j + 1] * p.slope;
          }
        }
      }
      if (p.z_mask != 0) {
        const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&zv[g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 m = __bfloat1622float2(mh[j]);
          v[2 * j] = m.x > 0.f ? v[2 * j] : 0.f;
          v[2 * j + 1] = m.y > 0.f ? v[2 * j + 1] : 0.f;
        }
      }
      if (p.act != ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act, p.slope);
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]);
      o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]);
      o.w = pack_bf16x2(v[6], v[7]);
      // one full 32-byte sector per store where the group pair is complete
      if ((g & 1) == 0) {
        if (c + 16 <= p.n_valid) o_even = o;
        else *reinterpret_cast<uint4*>(out + out_off + c) = o;
      } else {
        stg256(out + out_off + c - 8, o_even, o);
      }
      if (STATS) {
        // statistics of the values as stored (bf16-rounded): what the consumer will normalise
        const __nv_bfloat162* oh2 = reinterpret_cast<const __nv_bfloat162*>(&o);
        const __nv_bfloat162* zh = (p.stat_z != nullptr) ? reinterpret_cast<const __nv_bfloat162*>(&zv[g]) : oh2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(oh2[j]);
          const float2 z = __bfloat1622float2(zh[j]);
          sv[8 * g + 2 * j] = f.x;
          sv[8 * g + 2 * j + 1] = f.y;
          sq[8 * g + 2 * j] = f.x * z.x;
          sq[8 * g + 2 * j + 1] = f.y * z.y;
        }
      }
    }
    if constexpr (STATS && NC == 32) {
      const float s1 = warp_col_reduce32(sv, lane);
      const float s2 = warp_col_reduce32(sq, lane);
      float* so = p.stat_out + stat_row * 2 * p.stat_ld + stat_col + lane;
      so[0] = s1;
      so[p.stat_ld] = s2;
    }
  } else {
    if (!row_valid) return;
    // generic strided path (fp32 outputs, narrow channel counts)
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const int c = col0 + j;
      if (c < p.n_valid) {
        float v = __uint_as_float(r[j]) * alpha;
        if (p.bias != nullptr) v += __ldg(p.bias + c);
        v = apply_act(v, p.act, p.slope);
        if (p.out_f32)
          reinterpret_cast<float*>(p.out)[out_off + c * p.o_sc] = v;
        else
          reinterpret_cast<__nv_bfloat16*>(p.out)[out_off + c * p.o_sc] = __float2bfloat16(v);
      }
    }
  }
}

// Epilogue of ONE output tile for one epilogue warp: waits for the accumulator stage, then
// alpha / bias / aux / activation / store (+ fused statistics) for this warp's columns
// [c_begin, c_end) of accumulator rows q*32 .. q*32+31. `mt_full` = m-tile index incl. phase.
// fprop_epilogue_at: the tile's (phase, image) and this thread's output pixel (oh, ow) are already known.
template <int BLOCK_N>
__device__ __forceinline__ void fprop_epilogue_at(const FpropParams& p, int mt_full, int ph, int img, int oh, int ow,
                                                  int n_blk, uint32_t tmem_base, int as, uint32_t aphase,
                                                  uint64_t* tfull_bar, int q, int lane, int c_begin, int c_end,
                                                  float alpha);

template <int BLOCK_N>
__device__ __forceinline__ void fprop_epilogue_tile(const FpropParams& p, int mt_full, int n_blk, uint32_t tmem_base,
                                                    int as, uint32_t aphase, uint64_t* tfull_bar, int q, int lane,
                                                    int c_begin, int c_end, float alpha) {
  const int row = q * 32 + lane;
  int mt = mt_full;
  const int ph = mt % p.phases;
  mt /= p.phases;
  const int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int th = mt % p.tiles_h;
  const int img = mt / p.tiles_h;
  const int oh = th * p.TH + row / p.TW;
  const int ow = tw * p.TW + row % p.TW;
  fprop_epilogue_at<BLOCK_N>(p, mt_full, ph, img, oh, ow, n_blk, tmem_base, as, aphase, tfull_bar, q, lane, c_begin,
                             c_end, alpha);
}

template <int BLOCK_N>
__device__ __forceinline__ void fprop_epilogue_at(const FpropParams& p, int mt_full, int ph, int img, int oh, int ow,
                                                  int n_blk, uint32_t tmem_base, int as, uint32_t aphase,
                                                  uint64_t* tfull_bar, int q, int lane, int c_begin, int c_end,
                                                  float alpha) {
  const bool row_valid = (oh < p.OH) && (ow < p.OW);
  const int64_t out_off = p.o_ph[ph] + img * p.o_sn + oh * p.o_sh + ow * p.o_sw;
  const int64_t aux_off = p.a_ph[ph] + img * p.a_sn + oh * p.a_sh + ow * p.a_sw;
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
  if constexpr (BLOCK_N == 16) {
    mbar_wait(&tfull_bar[as], aphase);
    tc_fence_after();
    uint32_t r[16];
    uint4 av[2], zv[2];
    epilogue_load<16>(p, n_blk * BLOCK_N, row_valid, out_off, aux_off, av, zv);
    tmem_ld_32x16(taddr, r);
    tmem_ld_wait();
    fprop_epilogue_chunk<16, false>(p, r, av, zv, n_blk * BLOCK_N, row_valid, out_off, alpha);
  } else {
    const int64_t stat_row = int64_t(mt_full) * 4 + q;
    // (column, output offset, aux offset) of accumulator chunk c
    auto locate = [&](int c, int& col, int64_t& oo, int64_t& ao) {
      col = n_blk * BLOCK_N + c;
      oo = out_off;
      ao = aux_off;
      if (p.fold_c > 0) {                 // column -> (image, channel)
        const int img_o = col / p.fold_c;
        col -= img_o * p.fold_c;
        oo += img_o * p.o_sn;
        ao += img_o * p.a_sn;
      }
    };
    // While the main loop of this tile is still running: pull the tile's aux / stat_z rows into L2
    // and the first chunk's operands into registers.
    const bool has_aux = p.aux_mode != AUX_NONE && p.mask_scale == nullptr, has_z = p.stat_z != nullptr;
    const float* msc = p.mask_scale ? p.mask_scale + int64_t(img) * p.mask_ld : nullptr;
    const float* msh = p.mask_shift ? p.mask_shift + int64_t(img) * p.mask_ld : nullptr;
    uint4 pa[4], pz[4];
    if (has_aux || has_z) {
      if (row_valid) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 64) {
          int col;
          int64_t oo, ao;
          locate(c, col, oo, ao);
          if (has_aux) prefetch_l2(p.aux + ao + col);
          if (has_z) prefetch_l2(p.stat_z + oo + col);
        }
      }
      int col;
      int64_t oo, ao;
      locate(c_begin, col, oo, ao);
      epilogue_load<32>(p, col, row_valid, oo, ao, pa, pz);
    }
    mbar_wait(&tfull_bar[as], aphase);
    tc_fence_after();
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c, r);
      uint4 ca[4], cz[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        ca[g] = pa[g];
        cz[g] = pz[g];
      }
      if ((has_aux || has_z) && c + 32 < c_end) {
        int ncol;
        int64_t noo, nao;
        locate(c + 32, ncol, noo, nao);
        epilogue_load<32>(p, ncol, row_valid, noo, nao, pa, pz);
      }
      int col;
      int64_t oo, ao;
      locate(c, col, oo, ao);
      tmem_ld_wait();
      if (p.stat_out != nullptr)
        fprop_epilogue_chunk<32, true>(p, r, ca, cz, col, row_valid, oo, alpha, stat_row, n_blk * BLOCK_N + c, lane,
                                       msc, msh);
      else
        fprop_epilogue_chunk<32, false>(p, r, ca, cz, col, row_valid, oo, alpha, 0, 0, 0, msc, msh);
    }
  }
}

// Lean epilogue of the 64-wide tiles for the common case (see epilogue_is_lean): NC accumulator columns of one
// output pixel -> alpha * acc + bias -> ReLU / LeakyReLU / none -> bf16, 32-byte stores. The accumulator stage is
// handed back as soon as it sits in registers. `bias_r` stays in registers for the whole kernel.
__device__ __forceinline__ bool epilogue_is_lean(const FpropParams& p, int block_n, bool with_stats = false) {
  return p.aux_mode == AUX_NONE && (with_stats || p.stat_out == nullptr) && p.stat_z == nullptr && p.z_mask == 0 &&
         !p.out_f32 && p.o_sc == 1 && p.fold_c == 0 && p.n_valid == block_n && p.n_blocks == 1 && p.act != ACT_TANH;
}
template <int NC>
__device__ __forceinline__ void epilogue_lean(const FpropParams& p, const float (&bias_r)[NC], float alpha, uint32_t taddr,
                                              uint64_t* tempty, bool store, __nv_bfloat16* o) {
  static_assert(NC == 32, "one tcgen05.ld.32x32b.x32 per call");
  tc_fence_after();
  uint32_t r[NC];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(tempty);                            // the accumulator is in registers
  if (!store) return;
  float v[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) v[j] = fmaf(__uint_as_float(r[j]), alpha, bias_r[j]);
  if (p.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
  } else if (p.act == ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
  }
#pragma unroll
  for (int g = 0; g < NC / 8; g += 2) {
    uint4 a, b;
    a.x = pack_bf16x2(v[8 * g], v[8 * g + 1]);
    a.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
    a.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
    a.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
    b.x = pack_bf16x2(v[8 * g + 8], v[8 * g + 9]);
    b.y = pack_bf16x2(v[8 * g + 10], v[8 * g + 11]);
    b.z = pack_bf16x2(v[8 * g + 12], v[8 * g + 13]);
    b.w = pack_bf16x2(v[8 * g + 14], v[8 * g + 15]);
    stg256(o + 8 * g, a, b);
  }
}

// Lean epilogue that also accumulates the InstanceNorm statistics of the values AS STORED (bf16-rounded) into
// per-thread registers: s1 += v, s2 += v*v for this thread's pixel, NC channels. The ring kernel's work item is a
// column of consecutive output rows, so one thread sees the same 128-pixel column position row after row and needs
// ONE warp transpose-reduce per item instead of one per tile (ring_item_stats). Bias comes from shared memory
// (the registers hold the 2 x NC sums instead).
template <int NC>
__device__ __forceinline__ void epilogue_lean_stats(const FpropParams& p, const float* bias_s, float alpha, uint32_t taddr,
                                                    uint64_t* tempty, bool store, __nv_bfloat16* o, float (&s1)[NC],
                                                    float (&s2)[NC]) {
  static_assert(NC == 32, "one tcgen05.ld.32x32b.x32 per call");
  tc_fence_after();
  uint32_t r[NC];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(tempty);                            // the accumulator is in registers
  if (!store) return;
#pragma unroll
  for (int g = 0; g < NC / 8; g += 2) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = fmaf(__uint_as_float(r[8 * g + 2 * j]), alpha, bias_s[8 * g + 2 * j]);
      float b = fmaf(__uint_as_float(r[8 * g + 2 * j + 1]), alpha, bias_s[8 * g + 2 * j + 1]);
      if (p.act == ACT_RELU) {
        a = a > 0.f ? a : 0.f;
        b = b > 0.f ? b : 0.f;
      } else if (p.act == ACT_LRELU) {
        a = a > 0.f ? a : a * p.slope;
        b = b > 0.f ? b : b * p.slope;
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      const float2 f = __bfloat1622float2(h);
      s1[8 * g + 2 * j] += f.x;
      s2[8 * g + 2 * j] = fmaf(f.x, f.x, s2[8 * g + 2 * j]);
      s1[8 * g + 2 * j + 1] += f.y;
      s2[8 * g + 2 * j + 1] = fmaf(f.y, f.y, s2[8 * g + 2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    stg256(o + 8 * g, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
  }
}

// Lean epilogue of a dgrad whose output feeds an InstanceNorm backward (ring_item_stats with stat_z): the activation
// mask is recomputed from the norm input z and its per-(image, channel) scale / shift (msk = [scale[64] | shift[64]] in
// shared memory), g = act'(z*scale + shift) * alpha * acc is stored as bf16, and the two reductions of the norm
// backward -- sum g, sum g*z of the values as stored -- stay in this thread's registers over the rows of the work item.
template <int NC>
__device__ __forceinline__ void epilogue_lean_mask_stats(const FpropParams& p, const float* msk, float alpha, float off,
                                                         uint32_t taddr, uint64_t* tempty, bool store,
                                                         const uint4 (&zv)[NC / 8], __nv_bfloat16* o, float (&s1)[NC],
                                                         float (&s2)[NC]) {
  static_assert(NC == 32, "one tcgen05.ld.32x32b.x32 per call");
  tc_fence_after();
  uint32_t r[NC];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
  tc_fence_before();
  mbar_arrive(tempty);                            // the accumulator is in registers
  if (!store) return;
#pragma unroll
  for (int g = 0; g < NC / 8; g += 2) {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * g + 2 * j;
      const __nv_bfloat162* zh = reinterpret_cast<const __nv_bfloat162*>(&zv[c / 8]);
      const float2 z = __bfloat1622float2(zh[(c % 8) / 2]);
      float a = __uint_as_float(r[c]) * alpha, b = __uint_as_float(r[c + 1]) * alpha;
      a = fmaf(z.x, msk[c], msk[64 + c]) > 0.f ? a : a * off;
      b = fmaf(z.y, msk[c + 1], msk[64 + c + 1]) > 0.f ? b : b * off;
      const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      const float2 f = __bfloat1622float2(h);
      s1[c] += f.x;
      s2[c] = fmaf(f.x, z.x, s2[c]);
      s1[c + 1] += f.y;
      s2[c + 1] = fmaf(f.y, z.y, s2[c + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    stg256(o + 8 * g, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
  }
}

// 12 warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 idle, 4..11 = epilogue. The two
// epilogue warps of a TMEM lane quadrant (warp % 4) split the accumulator columns in halves.
constexpr int kFpropThreads = 384;

template <int BLOCK_N>
__global__ void __launch_bounds__(kFpropThreads, 1) fprop_kernel(const __grid_constant__ FpropParams p) {
  pdl_entry();
  using Cfg = FpropCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  constexpr int NACC = Cfg::kAccStages;
  uint64_t* tempty_bar = tfull_bar + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + NACC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], (BLOCK_N >= 64 && !Cfg::kAltTiles) ? 256 : 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w * p.phases;
  const int total = m_tiles * p.n_blocks;
  const int kblocks = p.taps * p.cblocks;

  // Both single-thread roles run warp-converged: every lane walks the loops and waits on the
  // barriers, one elected lane issues TMA / MMA / commit. (Issuing from inside `if (lane == 0)`
  // makes ptxas wrap every uniform-datapath instruction in an ELECT / BRA.U.ANY loop, which costs
  // more than the MMA itself for narrow tiles.)
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int n_blk = tile % p.n_blocks;
      int mt = tile / p.n_blocks;
      const int ph = mt % p.phases;
      mt /= p.phases;
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int oh0 = th * p.TH, ow0 = tw * p.TW;
      const int n_off = n_blk * BLOCK_N + img * p.b_row_per_image + ph * p.b_row_per_phase;
      int kb = 0;
      for (int t = 0; t < p.taps; ++t) {
        int map = 0, cw = ow0, chh = oh0, cn = t;
        if (!p.tap_is_image) {
          const Tap tp = p.tap[ph * p.taps + t];
          map = tp.map;
          cw = ow0 + tp.dw;
          chh = oh0 + tp.dh;
          cn = img;
        }
        for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_4d(sa, &p.tmA[map], &full_bar[stage], cb * kBlockK, cw, chh, cn);
            tma_load_2d(sa + kABytes, &p.tmB, &full_bar[stage], kb * kBlockK, n_off);
          }
          __syncwarp();
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int as = it % NACC;
      const uint32_t aphase = (it / NACC) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = make_smem_desc(sa, 0, 1024);
          const uint64_t db = make_smem_desc(sa + kABytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)   // +32 B per K step = +2 in the descriptor's address field
            umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == kblocks - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4 && (BLOCK_N >= 64 || warp < 8)) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;                       // which half of the columns (kAltTiles: which tiles)
    constexpr int kHalfCols = (BLOCK_N >= 64 && !Cfg::kAltTiles) ? BLOCK_N / 2 : BLOCK_N;
    const int c_begin = Cfg::kAltTiles ? 0 : half * kHalfCols, c_end = c_begin + kHalfCols;
    const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    int it = 0;
    if constexpr (BLOCK_N == 64) {
      // 64-wide tiles with a short K (the first layers of SE / D: K = 64) are bound by the epilogue: lean path
      if (epilogue_is_lean(p, BLOCK_N) && p.phases == 1) {
        float bias_r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bias_r[j] = p.bias != nullptr ? __ldg(p.bias + c_begin + j) : 0.f;
        const int row = q * 32 + lane;
        const int dh = row / p.TW, dw = row % p.TW;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
          const int as = it % NACC;
          int mt = tile;
          const int tw = mt % p.tiles_w;
          mt /= p.tiles_w;
          const int th = mt % p.tiles_h;
          const int img = mt / p.tiles_h;
          const int oh = th * p.TH + dh, ow = tw * p.TW + dw;
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + img * p.o_sn + oh * p.o_sh + ow * p.o_sw + c_begin;
          mbar_wait(&tfull_bar[as], (it / NACC) & 1);
          epilogue_lean<32>(p, bias_r, alpha, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c_begin,
                            &tempty_bar[as], oh < p.OH && ow < p.OW, o);
        }
        it = -1;
      }
    }
    for (int tile = blockIdx.x; it >= 0 && tile < total; tile += gridDim.x, ++it) {
      if (Cfg::kAltTiles && (it & 1) != half) continue;
      const int as = it % NACC;
      const uint32_t aphase = (it / NACC) & 1;
      fprop_epilogue_tile<BLOCK_N>(p, tile / p.n_blocks, tile % p.n_blocks, tmem_base, as, aphase, tfull_bar, q, lane,
                                   c_begin, c_end, alpha);
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ M = 256 per CTA, N = 128
// fprop_m2_kernel: the BLOCK_N = 128 implicit GEMM with TWO 128-pixel m-tiles per CTA (vertically adjacent
// logical tiles 2*th and 2*th+1 of one image / phase), both multiplied by the SAME weight tile: per 64-deep
// K block the CTA stages 32 KiB of activations (one TMA box {64, TW, 2*TH}) + 16 KiB of weights for 4.2 MF
// instead of 16 + 16 KiB for 2.1 MF. The 1-CTA N = 128 kernel sits on the L2 -> shared-memory roof (the
// phased 256 -> 128 transposed conv and the stride-2 dgrad of 128 -> 256 measured 610 TF/s); this halves the
// weight traffic per FLOP, the ratio the 256-wide tiles have. Two accumulators (128 columns each) per TMEM
// stage, double buffered: 512 columns. Logical m-tile indices (and so the statistics rows) are unchanged.
constexpr int kM2Stages = 4;
constexpr int kM2ABytes = 2 * kABytes;                       // 32 KiB
constexpr int kM2BBytes = 128 * kBlockK * 2;                 // 16 KiB
constexpr int kM2StageBytes = kM2ABytes + kM2BBytes;         // 48 KiB
constexpr int kM2SmemBytes = kM2Stages * kM2StageBytes + 1024 + 256;

__global__ void __launch_bounds__(kFpropThreads, 1) fprop_m2_kernel(const __grid_constant__ FpropParams p) {
  pdl_entry();
  constexpr int BLOCK_N = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kM2Stages * kM2StageBytes);
  uint64_t* empty_bar = full_bar + kM2Stages;
  uint64_t* tfull_bar = empty_bar + kM2Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kM2Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item = (image, pair of tile rows, tile column, phase, n block); phase and n block vary fastest
  const int th_pairs = p.tiles_h / 2;
  const int total = p.n_img * th_pairs * p.tiles_w * p.phases * p.n_blocks;
  const int kblocks = p.taps * p.cblocks;
  auto decode = [&](int item, int& n_blk, int& ph, int& tw, int& th2, int& img) {
    n_blk = item % p.n_blocks;
    int r = item / p.n_blocks;
    ph = r % p.phases;
    r /= p.phases;
    tw = r % p.tiles_w;
    r /= p.tiles_w;
    th2 = r % th_pairs;
    img = r / th_pairs;
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int n_blk, ph, tw, th2, img;
      decode(item, n_blk, ph, tw, th2, img);
      const int oh0 = 2 * th2 * p.TH, ow0 = tw * p.TW;
      const int n_off = n_blk * BLOCK_N + ph * p.b_row_per_phase;
      int kb = 0;
      for (int t = 0; t < p.taps; ++t) {
        const Tap tp = p.tap[ph * p.taps + t];
        for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = smem + stage * kM2StageBytes;
            mbar_expect_tx(&full_bar[stage], kM2StageBytes);
            tma_load_4d(sa, &p.tmA[tp.map], &full_bar[stage], cb * kBlockK, ow0 + tp.dw, oh0 + tp.dh, img);
            tma_load_2d(sa + kM2ABytes, &p.tmB, &full_bar[stage], kb * kBlockK, n_off);
          }
          __syncwarp();
          if (++stage == kM2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 2 * BLOCK_N;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kM2StageBytes);
          const uint64_t da0 = make_smem_desc(sa, 0, 1024);
          const uint64_t da1 = make_smem_desc(sa + kABytes, 0, 1024);
          const uint64_t db = make_smem_desc(sa + kM2ABytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            umma_bf16(d_tmem, da0 + 2 * k, db + 2 * k, idesc, acc);
            umma_bf16(d_tmem + BLOCK_N, da1 + 2 * k, db + 2 * k, idesc, acc);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kblocks - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++stage == kM2Stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int c_begin = half * (BLOCK_N / 2), c_end = c_begin + BLOCK_N / 2;
    const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    int it = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      int n_blk, ph, tw, th2, img;
      decode(item, n_blk, ph, tw, th2, img);
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        // logical m-tile index in the 1-CTA kernel's order (phase fastest, then tile column, tile row, image)
        const int mt_full = ((img * p.tiles_h + 2 * th2 + sub) * p.tiles_w + tw) * p.phases + ph;
        // fprop_epilogue_tile addresses TMEM at base + as * BLOCK_N: fold this kernel's stage / sub-tile layout in
        const uint32_t base = tmem_base + as * BLOCK_N + sub * BLOCK_N;
        fprop_epilogue_tile<BLOCK_N>(p, mt_full, n_blk, base, as, aphase, tfull_bar, q, lane, c_begin, c_end, alpha);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ CTA-pair fprop
// fprop2_kernel: the BLOCK_N = 256 implicit GEMM on a CTA PAIR (cluster of 2, tcgen05 cta_group::2).
// One MMA covers 256 output pixels (two 128-pixel m-tiles, one per CTA) x 256 channels; each CTA
// stages its own A tile and only HALF of the weight tile (128 of the 256 rows), so the L2 -> shared
// memory traffic per MMA drops from 48 KiB to 32 KiB per SM (the 1-CTA kernel runs at the L2
// bandwidth roof) and the stage ring gets 6 slots instead of 4. Rank 0 issues the MMAs and
// owns the full / accumulator-empty barriers; commits are multicast to both CTAs.
// Requirements (host): phases == 1, no per-image B offsets, an even number of m-tiles.
constexpr int k2Stages = 6;
constexpr int k2BHalfBytes = 128 * kBlockK * 2;           // 16 KiB
constexpr int k2StageBytes = kABytes + k2BHalfBytes;      // 32 KiB
constexpr int k2SmemBytes = k2Stages * k2StageBytes + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFpropThreads, 1)
    fprop2_kernel(const __grid_constant__ FpropParams p) {
  pdl_entry();
  constexpr int BLOCK_N = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes);
  uint64_t* empty_bar = full_bar + k2Stages;
  uint64_t* tfull_bar = empty_bar + k2Stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmA[i]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 512);          // the 8 epilogue warps of BOTH CTAs (rank 0's copy is used)
    }
    fence_barrier_init();
  }
  cluster_sync_all();                          // barriers of both CTAs are initialised
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                          // both halves of the accumulator are allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w;          // phases == 1
  const int pairs = (m_tiles / 2) * p.n_blocks;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int kblocks = p.taps * p.cblocks;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pr = cluster_id; pr < pairs; pr += n_clusters) {
      const int n_blk = pr % p.n_blocks;
      int mt = (pr / p.n_blocks) * 2 + static_cast<int>(rank);
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int img = mt / p.tiles_h;
      const int oh0 = th * p.TH, ow0 = tw * p.TW;
      const int n_off = n_blk * BLOCK_N + static_cast<int>(rank) * 128;      // this CTA's half of the weight rows
      int kb = 0;
      for (int t = 0; t < p.taps; ++t) {
        int map = 0, cw = ow0, chh = oh0, cn = t;          // Gram backward: tap t reads image t
        if (!p.tap_is_image) {
          const Tap tp = p.tap[t];
          map = tp.map;
          cw = ow0 + tp.dw;
          chh = oh0 + tp.dh;
          cn = img;
        }
        for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            uint8_t* sa = smem + stage * k2StageBytes;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * k2StageBytes);   // both CTAs' bytes
            tma_load_4d_2sm(sa, &p.tmA[map], &full_bar[stage], cb * kBlockK, cw, chh, cn);
            tma_load_2d_2sm(sa + kABytes, &p.tmB, &full_bar[stage], kb * kBlockK, n_off);
          }
          __syncwarp();
          if (++stage == k2Stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int pr = cluster_id; pr < pairs; pr += n_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * k2StageBytes);
          const uint64_t da = make_smem_desc(sa, 0, 1024);
          const uint64_t db = make_smem_desc(sa + kABytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 3);
          if (kb == kblocks - 1) umma_commit_2sm(&tfull_bar[as], 3);
        }
        __syncwarp();
        if (++stage == k2Stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int c_begin = half * (BLOCK_N / 2), c_end = c_begin + BLOCK_N / 2;
    const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    int it = 0;
    for (int pr = cluster_id; pr < pairs; pr += n_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_blk = pr % p.n_blocks;
      const int mt = (pr / p.n_blocks) * 2 + static_cast<int>(rank);
      fprop_epilogue_tile<BLOCK_N>(p, mt, n_blk, tmem_base, as, aphase, tfull_bar, q, lane, c_begin, c_end, alpha);
      tc_fence_before();
      mbar_arrive_rank0(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  cluster_sync_all();                          // no CTA exits (or frees TMEM) while its pair still works
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ row-fold
// Every input strip is used ONCE: its MMAs add the strip's contribution to all R output rows it belongs to
// (filter row j of input row y feeds output row y - j). Output row o accumulates in TMEM column block o % 16
// (32 columns each, all 512 columns allocated), so the R blocks one strip touches are CONSECUTIVE (mod 16) and
// the B operand is the fixed matrix [W_{R-1}; ...; W_0]: per 16-channel K step one N = 32*(R-1) accumulate MMA
// (two when the block window wraps) plus the N = 32 MMA that initialises the newest row's block -- a quarter of
// the MMA time of accumulating each output row from its R strips (R * 4 N = 32 MMAs, shared-memory bound).
constexpr int kRfRing = 8;                       // strips in flight
constexpr int kRfStripBytes = kTileM * 128;      // 128 pixels x 64 ch bf16
constexpr int kRfMaxR = 7;
constexpr int kRfWBytes = 32 * 128;              // one filter row: 32 (s, co) rows x 64 ch
constexpr int kRfFoldLd = 33;
constexpr int kRfThreads = 384;                  // warps 0..3 as in the fprop kernels, 4..7 / 8..11 = two epilogue groups
constexpr int kRfSmemBytes = kRfMaxR * kRfWBytes + kRfRing * kRfStripBytes + 4 * kTileM * kRfFoldLd * 4 + 1024 + 512;

__global__ void __launch_bounds__(kRfThreads, 1) fprop_rowfold_kernel(const __grid_constant__ RowfoldParams p) {
  pdl_entry();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* wsm = smem;                                              // R x [32][64] bf16
  uint8_t* ring = smem + kRfMaxR * kRfWBytes;                        // 28 KiB -> 1024-aligned
  float* fold = reinterpret_cast<float*>(ring + kRfRing * kRfStripBytes);   // 2 groups x 2 x [128][33] fp32
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(fold) + 4 * kTileM * kRfFoldLd * 4);
  uint64_t* empty_bar = full_bar + kRfRing;
  uint64_t* tfull_bar = empty_bar + kRfRing;
  uint64_t* tempty_bar = tfull_bar + 16;
  uint64_t* wfull_bar = tempty_bar + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.R, S = p.S;
  const int tile_out = kTileM - (S - 1);             // output pixels per tile

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kRfRing; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 16; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int items = p.n_img * p.tiles_w * p.chunks_h;

  // work item -> (image, column tile, first output row, number of output rows)
  auto decode = [&](int item, int& img, int& tw, int& h0, int& nrows) {
    const int ch = item % p.chunks_h;
    const int rest = item / p.chunks_h;
    tw = rest % p.tiles_w;
    img = rest / p.tiles_w;
    h0 = ch * p.rows_per_item;
    nrows = min(p.rows_per_item, p.OH - h0);
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, static_cast<uint32_t>(R) * kRfWBytes);
      // resident filter in DESCENDING filter-row order: block j' holds W_{R-1-j'}
      for (int r = 0; r < R; ++r) tma_load_2d(wsm + (R - 1 - r) * kRfWBytes, &p.tmB, wfull_bar, 0, r * 32);
    }
    __syncwarp();
    uint32_t g = 0;                                  // strips issued so far (ring position)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      const int x0 = p.org_w + tw * tile_out;
      const int y0 = p.org_h + h0;
      const int nstrips = nrows + R - 1;
      for (int j = 0; j < nstrips; ++j, ++g) {
        const uint32_t slot = g % kRfRing;
        mbar_wait(&empty_bar[slot], ((g / kRfRing) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[slot], kRfStripBytes);
          tma_load_4d(ring + slot * kRfStripBytes, &p.tmA, &full_bar[slot], 0, x0, y0 + j, img);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    mbar_wait(wfull_bar, 0);
    const uint32_t w_addr = smem_u32(wsm);
    const uint32_t ring_addr = smem_u32(ring);
    const int nacc = R - 1;                          // blocks that accumulate (filter rows R-1 .. 1)
    const uint32_t idesc_new = make_idesc_bf16(kTileM, 32, 0, 0);
    const uint32_t idesc_all = make_idesc_bf16(kTileM, 32 * R, 0, 0);
    const uint64_t db_new = make_smem_desc(w_addr + nacc * kRfWBytes, 0, 1024);     // W_0
    uint32_t g = 0;                                  // strips consumed so far == counter of the newest output row
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      const int nstrips = nrows + R - 1;
      for (int t = 0; t < nstrips; ++t, ++g) {
        const uint32_t blk = g & 15u;
        mbar_wait(&tempty_bar[blk], ((g >> 4) & 1u) ^ 1u);          // the row that used this block 16 rows ago is read
        mbar_wait(&full_bar[g % kRfRing], (g / kRfRing) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = make_smem_desc(ring_addr + (g % kRfRing) * kRfStripBytes, 0, 1024);
          const uint32_t b0 = (g - static_cast<uint32_t>(nacc)) & 15u;   // block of filter row R-1
          const int n1 = min(nacc, static_cast<int>(16u - b0));           // accumulate blocks before the wrap
          const int n2 = nacc - n1;
          const uint64_t db0 = make_smem_desc(w_addr, 0, 1024);
          const uint64_t db1 = make_smem_desc(w_addr + n1 * kRfWBytes, 0, 1024);
          const uint32_t idesc1 = make_idesc_bf16(kTileM, n1 > 0 ? 32 * n1 : 32, 0, 0);
          const uint32_t idesc2 = make_idesc_bf16(kTileM, n2 > 0 ? 32 * n2 : 32, 0, 0);
          const bool contiguous = (n2 == 0) && (nacc == 0 || blk == b0 + static_cast<uint32_t>(nacc));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (k > 0 && contiguous) {               // one MMA over all R blocks
              umma_bf16(tmem_base + b0 * 32, da + 2 * k, db0 + 2 * k, idesc_all, 1u);
              continue;
            }
            if (n1 > 0) umma_bf16(tmem_base + b0 * 32, da + 2 * k, db0 + 2 * k, idesc1, 1u);
            if (n2 > 0) umma_bf16(tmem_base, da + 2 * k, db1 + 2 * k, idesc2, 1u);
            umma_bf16(tmem_base + blk * 32, da + 2 * k, db_new + 2 * k, idesc_new, k > 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[g % kRfRing]);      // the strip is dead
          // output row counter g - (R-1) is complete (for t < R-1 that is one of the previous item's phantom
          // counters: committed all the same, the epilogue walks every counter in order)
          if (g >= static_cast<uint32_t>(nacc)) umma_commit(&tfull_bar[(g - static_cast<uint32_t>(nacc)) & 15u]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int t = q * 32 + lane;                     // TMEM lane = strip pixel = local output pixel
    const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    float chs[4] = {alpha, alpha, alpha, alpha};
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      if (co < p.n_valid && p.bias != nullptr) bias[co] = __ldg(p.bias + co);
      if (co < p.n_valid && p.ch_scale != nullptr) chs[co] = alpha * __ldg(p.ch_scale + co);
    }
    // The two epilogue groups take ALTERNATE output rows (one row's chain -- tfull, tcgen05.ld, transpose through
    // shared memory, fold, store -- is longer than a strip's MMAs); both walk every counter to stay in step.
    const int grp = (warp - 4) >> 2;
    float* gfold = fold + grp * 2 * kTileM * kRfFoldLd;
    int it = 0;                                      // real rows seen
    int mine = 0;                                    // real rows handled by this group (fold buffer parity)
    uint32_t c = 0;                                  // output-row counter, phantom ones (R-1 per item) included
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      const int ow = tw * tile_out + t;
      const bool valid = (t < tile_out) && (ow < p.OW);
      const bool last_item = item + static_cast<int>(gridDim.x) >= items;
      const int ncount = last_item ? nrows : nrows + R - 1;         // the last phantom counters never complete
      for (int i = 0; i < ncount; ++i, ++c) {
        const uint32_t blk = c & 15u;
        if (i >= nrows) {                            // phantom: nothing to read, group 0 hands the block back
          if (grp == 0) {
            mbar_wait(&tfull_bar[blk], (c >> 4) & 1u);
            mbar_arrive(&tempty_bar[blk]);
          }
          continue;
        }
        if (((it++) & 1) != grp) continue;
        mbar_wait(&tfull_bar[blk], (c >> 4) & 1u);
        tc_fence_after();
        uint32_t r32[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * 32, r32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty_bar[blk]);               // accumulator is in registers: free the block
        float* fb = gfold + (mine & 1) * kTileM * kRfFoldLd;
        ++mine;
#pragma unroll
        for (int j = 0; j < 32; ++j) fb[t * kRfFoldLd + j] = __uint_as_float(r32[j]);
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // the four warps of this group only
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (valid) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          for (int s = 0; s < S; ++s) {
            const float* row = fb + (t + s) * kRfFoldLd + s * 4;
#pragma unroll
            for (int co = 0; co < 4; ++co) acc[co] += row[co];
          }
          const int oh = h0 + i;
          float* o = p.out + img * p.o_sn + oh * p.o_sh + ow * p.o_sw;
#pragma unroll
          for (int co = 0; co < 4; ++co)
            if (co < p.n_valid) o[co * p.o_sc] = apply_act(acc[co] * chs[co] + bias[co], p.act, 0.f);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------ ring (N = 64)
// Stride-1 R x S convolution with 64 input and <= 64 output channels (VGG conv 1_2 fwd / dgrad,
// losses.py:15) and the row-patch 7x7 convs (R = 7 overlapped-window rows, S = 1). The generic kernel
// re-fetches the A tile once per filter tap and the weight tile once per K block, and at N = 64 that
// traffic (24 KiB per 2 MF) pins it to the L2 -> shared-memory roof at ~25 % tensor-pipe activity. Here
//   * the whole packed filter (R*S x [64][64] bf16) stays resident in shared memory,
//   * one TMA strip of 128 + S - 1 pixels per INPUT row lives in an 8-slot ring: the S horizontal taps
//     are MMAs whose A descriptors start s rows into the strip, and the R - 1 strips shared with the
//     next output row are re-used instead of re-loaded (work item = a column of consecutive output rows).
// L2 traffic per 128-pixel tile drops from R*S*24 KiB to ~17 KiB. Epilogue = the generic one (BLOCK_N = 64).
//   * with ring_phases = 4 (the four output phases of a stride-2 transposed conv / dgrad, each a 2x2 stride-1 conv
//     of the SAME input) CTA b works on phase b % 4: the four phases walk the images side by side in ONE launch,
//     so the input comes from HBM once and the other three phases hit L2.
// The ring takes whatever shared memory the resident filter leaves (ring_slots, up to 16) as look-ahead.
//   * ring_stack (default): the MMA warp walks INPUT rows. A strip feeds the R output rows y .. y-R+1 through filter
//     rows 0 .. R-1; their accumulators sit side by side in 8 TMEM stages (output row `it` in stage it % 8) and the
//     R filter-row tiles of one (s, channel block) back to back in shared memory (r descending), so one
//     N = 64 R MMA per (strip, s, k step) replaces R N = 64 ones and a strip is consumed exactly once. Legacy order
//     (msig_debug_set_ring_mode bit 2): one N = 64 MMA chain per OUTPUT row over its R resident strips, 4 stages.
//   * epilogues: lean (bias + ReLU / LeakyReLU, bf16 NHWC: the accumulator stage is handed back as soon as it is in
//     registers); lean + per-item statistics (ring_item_stats: a thread keeps the sums of its pixel column over the
//     rows of the work item, one warp transpose-reduce per item -- plain sum v, sum v^2, or with stat_z and
//     mask_scale / mask_shift the masked dgrad's sum g, sum g*z); anything else goes through the generic epilogue.
#ifdef MSIG_RING_PROFILE
// Probe build only (make PROF=1 -> libmsig_prof.so): cycles per role, summed over CTAs.
//  [0] MMA warp: wait tempty  [1] wait strips  [2] issue + commit  [3] tiles
//  [4] epilogue warp 4: wait tfull  [5] tcgen05.ld  [6] math + stores  [7] tiles
//  [8] producer: wait empty slot  [9] strips   [10] kernel cycles (CTA 0)
//  [11] stacked MMA loop: piece setup  [12] MMA issue  [13] commits (issuing thread)
__device__ unsigned long long g_ring_prof[16];
#define RING_PROF_T(var) const long long var = clock64()
#define RING_PROF_ADD(acc, a, b) acc += (b) - (a)
#else
#define RING_PROF_T(var)
#define RING_PROF_ADD(acc, a, b)
#endif
constexpr int kRingMaxSlots = 16;
constexpr int kRingMaxTaps = 9;                    // resident [64][64] filter tiles (R * S * channel blocks)
constexpr int kRingWBytes = 64 * 128;              // one tap: 64 output channels x 64 K
constexpr int kRingBarBytes = 2048;             // barriers + TMEM slot | +512: bias[64] fp32 | +1024: 2 x (mask scale[64], shift[64])
constexpr int kRingSmemMax = 232448;               // 227 KiB
static inline int ring_slot_bytes(int S) { return ((kTileM + S - 1) * 128 + 1023) & ~1023; }

__global__ void __launch_bounds__(kFpropThreads, 1) fprop_ring64_kernel(const __grid_constant__ FpropParams p) {
  pdl_entry();
  constexpr int BLOCK_N = 64;
  constexpr int kMaxAcc = 8;
  const bool STACK = p.ring_stack != 0;
  const int NACC = STACK ? 8 : 4;          // accumulator stages in TMEM (64 columns each)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int R = p.strip_r, S = p.strip_s;
  const int CB = p.ring_cb > 1 ? p.ring_cb : 1;                    // 64-channel blocks per input pixel
  const int SLOTS = p.ring_slots;
  const uint32_t slot_bytes = static_cast<uint32_t>(((kTileM + S - 1) * 128 + 1023) & ~1023);
  uint8_t* wsm = smem;                                             // R*S*CB x 8 KiB
  uint8_t* ring = smem + R * S * CB * kRingWBytes;                 // 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + SLOTS * slot_bytes);
  uint64_t* empty_bar = full_bar + kRingMaxSlots;
  uint64_t* tfull_bar = empty_bar + kRingMaxSlots;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint64_t* wfull_bar = tempty_bar + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  const uint32_t strip_tx = static_cast<uint32_t>(kTileM + S - 1) * 128u;
  // phase interleave: CTA b -> phase b % NPH, item sequence b / NPH, b / NPH + gridDim / NPH, ...
  const int NPH = p.ring_phases > 1 ? p.ring_phases : 1;
  const int ph = NPH > 1 ? static_cast<int>(blockIdx.x) % NPH : 0;
  const int cta0 = static_cast<int>(blockIdx.x) / NPH;
  const int cta_stride = static_cast<int>(gridDim.x) / NPH;
  const int org_h = NPH > 1 ? p.ring_org_h[ph] : p.org_h;
  const int org_w = NPH > 1 ? p.ring_org_w[ph] : p.org_w;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA[1]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (STACK) tmem_alloc(tmem_slot, 512); else tmem_alloc(tmem_slot, 256);
  }
  float* const bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);   // item-stats epilogue
  if (warp == 3) {
    bias_s[lane] = (p.bias != nullptr && lane < p.n_valid) ? __ldg(p.bias + lane) : 0.f;
    bias_s[lane + 32] = (p.bias != nullptr && lane + 32 < p.n_valid) ? __ldg(p.bias + lane + 32) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int items = p.n_img * p.tiles_w * p.ring_chunks;

  auto decode = [&](int item, int& img, int& tw, int& h0, int& nrows) {
    const int ch = item % p.ring_chunks;
    const int rest = item / p.ring_chunks;
    tw = rest % p.tiles_w;
    img = rest / p.tiles_w;
    h0 = ch * p.ring_rows;
    nrows = min(p.ring_rows, p.OH - h0);
  };

  if (warp == 0) {
    if (elect_one()) {
      // resident filter, in ring order: slot (r*S + s)*CB + cb <- K block (tap(r,s)*CB + cb) of the packed matrix
      mbar_expect_tx(wfull_bar, static_cast<uint32_t>(R * S * CB) * kRingWBytes);
      // (row-stacked mode: the R filter rows of one (s, cb) sit back to back, r descending, as ONE N = 64 R operand)
      for (int t = 0; t < R * S; ++t)
        for (int cb = 0; cb < CB; ++cb) {
          const int dst = STACK ? ((t % S) * CB + cb) * R + (R - 1 - t / S) : t * CB + cb;
          tma_load_2d(wsm + dst * kRingWBytes, &p.tmB, wfull_bar, (p.ring_tap[t] * CB + cb) * kBlockK,
                      ph * p.b_row_per_phase);
        }
    }
    __syncwarp();
    uint32_t slot = 0, sphase = 0;
#ifdef MSIG_RING_PROFILE
    long long pw = 0, pn = 0;
    const long long k0 = clock64();
#endif
    for (int item = cta0; item < items; item += cta_stride) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      const int x0 = org_w + tw * kTileM;
      const int y0 = org_h + h0;
      const int nstrips = (nrows + R - 1) * CB;                    // ring entry = (input row, channel block)
      for (int j = 0; j < nstrips; ++j) {
        RING_PROF_T(t0);
        mbar_wait(&empty_bar[slot], sphase ^ 1u);
        RING_PROF_T(t1);
        RING_PROF_ADD(pw, t0, t1);
#ifdef MSIG_RING_PROFILE
        ++pn;
#endif
        if (elect_one()) {
          mbar_expect_tx(&full_bar[slot], strip_tx);
          tma_load_4d(ring + slot * slot_bytes, &p.tmA[1], &full_bar[slot], (j % CB) * kBlockK, x0, y0 + j / CB, img);
        }
        __syncwarp();
        if (++slot == static_cast<uint32_t>(SLOTS)) { slot = 0; sphase ^= 1u; }
      }
    }
#ifdef MSIG_RING_PROFILE
    if (lane == 0) {
      atomicAdd(&g_ring_prof[8], static_cast<unsigned long long>(pw));
      atomicAdd(&g_ring_prof[9], static_cast<unsigned long long>(pn));
      if (blockIdx.x == 0) atomicAdd(&g_ring_prof[10], static_cast<unsigned long long>(clock64() - k0));
    }
#endif
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
    mbar_wait(wfull_bar, 0);
    const uint32_t w_addr = smem_u32(wsm);
    const uint32_t ring_addr = smem_u32(ring);
    // ring entries are numbered g = 0, 1, ... in load order; entry g sits in slot g % SLOTS, pass g / SLOTS.
    // g0 = first entry of the current item, kept as (slot, pass parity) to stay clear of divisions by SLOTS.
    uint32_t g0s = 0, g0p = 0;
    int it = 0;
#ifdef MSIG_RING_PROFILE
    long long m_te = 0, m_fu = 0, m_is = 0, m_setup = 0, m_mma = 0, m_commit = 0;
#endif
    auto entry = [&](int off, uint32_t& sl, uint32_t& par) {      // entry g0 + off (off < 2 * SLOTS)
      uint32_t x = g0s + static_cast<uint32_t>(off);
      par = g0p;
      while (x >= static_cast<uint32_t>(SLOTS)) { x -= SLOTS; par ^= 1u; }
      sl = x;
    };
    // Row-stacked issue order (ring_stack): an input-row strip feeds the R output rows y, y-1, .. y-R+1 through the
    // filter rows r = 0 .. R-1. Their accumulators sit side by side in TMEM (output row `it` in stage it % 8), and
    // the R filter-row tiles of one (s, cb) sit back to back in shared memory (r descending), so ONE N = 64 R MMA
    // per (strip, s, k) replaces R N = 64 MMAs: the N = 64 instruction is bound by its 6 KiB of shared-memory
    // operand reads (~60 cycles for 32 cycles of math); at N = 192 the A strip is read once for three rows' worth.
    // Pieces: an MMA may not wrap around the 8 stages, N <= 256, and the first MMA into a NEW output row
    // overwrites (accumulate = 0) that row's 64 columns only.
    // What bounds it now (profiles/probe/mma_rate.cu, ring_profile_r2.txt): the pipe runs these MMAs at their math
    // floor (N >= 128: N / 2 cycles; N = 64: 48, its 6 KiB of operand reads) but queues only ~2 instructions beyond
    // the one in flight, so whatever the issuing thread does between rows beyond ~190 cycles -- barrier waits
    // ~350, piece setup ~290, commits ~170 per row -- drains it: ~1.3 k cycles of MMAs in a ~2.1 k cycle tile.
    // Preparing the next row between the current row's MMAs (one thread running the whole loop, phases after every
    // k step) measured 2-4x SLOWER, inlined (the issue loop outgrows the instruction cache: ~300 cycles per MMA)
    // and as a non-inlined call (state in local memory): not adopted.
    if (STACK) {
      uint32_t gs = 0, gp = 0;               // ring cursor: strips are consumed strictly in load order
      int it0 = 0;                           // running output-row counter at the start of the item
      for (int item = cta0; item < items; item += cta_stride) {
        int img, tw, h0, nrows;
        decode(item, img, tw, h0, nrows);
        const int nin = nrows + R - 1;
        for (int j = 0; j < nin; ++j) {
          const int i_lo = max(0, j - R + 1), i_hi = min(nrows - 1, j);
          const int nb = i_hi - i_lo + 1;
          const bool has_new = j <= nrows - 1;
          RING_PROF_T(t0);
          // The row's barriers (accumulator stage of its new output row, its CB strips) are polled back to back
          // before any result is looked at: a poll is a ~100-cycle round trip, and this thread's idle time between
          // rows is exposed (the pipe queues only ~2 MMAs).
          const int itn = it0 + j;
          uint64_t* const te = &tempty_bar[itn & 7];
          const uint32_t tpar = ((itn >> 3) & 1) ^ 1u;
          uint32_t sl1 = gs + 1, par1 = gp;
          if (sl1 == static_cast<uint32_t>(SLOTS)) { sl1 = 0; par1 ^= 1u; }
          const bool ok_e = has_new ? mbar_try_wait(te, tpar) : true;
          const bool ok_0 = mbar_try_wait(&full_bar[gs], gp);
          const bool ok_1 = CB > 1 ? mbar_try_wait(&full_bar[sl1], par1) : true;
          if (!ok_e) mbar_wait(te, tpar);
          RING_PROF_T(t1);
          if (!ok_0) mbar_wait(&full_bar[gs], gp);
          if (!ok_1) mbar_wait(&full_bar[sl1], par1);
          RING_PROF_T(t2);
          RING_PROF_ADD(m_te, t0, t1);
          RING_PROF_ADD(m_fu, t1, t2);
          tc_fence_after();
          if (elect_one()) {
            const int b0 = R - 1 - (j - i_lo);           // first block of the stacked filter that takes part
            const int s0 = (it0 + i_lo) & 7;             // TMEM stage of output row i_lo
            // MMA pieces of this input row, fixed for all its (cb, s, k): (TMEM address, descriptor offset of the
            // first filter block, instruction descriptor). The issue loop below is a handful of integer adds per MMA
            // -- with per-MMA descriptor construction the single issuing thread, not the tensor pipe, set the pace.
            uint32_t d0 = 0, d1 = 0, d2 = 0, i0 = 0, i1 = 0, i2 = 0, o1 = 0, o2 = 0;   // piece 0 starts at block 0
            int np = 0;
            for (int b = 0; b < nb;) {
              const int stage = (s0 + b) & 7;
              const int len = min(min(nb - b, 4), 8 - stage);
              const uint32_t dd = tmem_base + stage * BLOCK_N, ii = make_idesc_bf16(kTileM, BLOCK_N * len, 0, 0);
              const uint32_t oo = static_cast<uint32_t>(b * (kRingWBytes >> 4));
              if (np == 0) { d0 = dd; i0 = ii; }
              else if (np == 1) { d1 = dd; i1 = ii; o1 = oo; }
              else { d2 = dd; i2 = ii; o2 = oo; }
              ++np;
              b += len;
            }
            const uint64_t wd0 = make_smem_desc(w_addr, 0, 1024) + static_cast<uint64_t>(b0 * (kRingWBytes >> 4));
            uint32_t sl = gs;
            RING_PROF_T(ts1);
            RING_PROF_ADD(m_setup, t2, ts1);
            for (int cb = 0; cb < CB; ++cb) {
              const uint32_t sa = ring_addr + sl * slot_bytes;
              for (int s = 0; s < S; ++s) {
                const uint64_t da = make_smem_desc(sa + s * 128, 0, 1024);      // s pixels into the strip
                const uint64_t db = wd0 + static_cast<uint64_t>(((s * CB + cb) * R) * (kRingWBytes >> 4));
                int k = 0;
                if (has_new && cb == 0 && s == 0) {
                  // first MMA into the NEW output row (last block): overwrite its 64 columns, accumulate the others
                  for (int b = 0; b < nb - 1;) {
                    const int stage = (s0 + b) & 7;
                    const int len = min(min(nb - 1 - b, 4), 8 - stage);
                    umma_bf16(tmem_base + stage * BLOCK_N, da, db + static_cast<uint64_t>(b * (kRingWBytes >> 4)),
                              make_idesc_bf16(kTileM, BLOCK_N * len, 0, 0), 1u);
                    b += len;
                  }
                  umma_bf16(tmem_base + ((s0 + nb - 1) & 7) * BLOCK_N, da,
                            db + static_cast<uint64_t>((nb - 1) * (kRingWBytes >> 4)),
                            make_idesc_bf16(kTileM, BLOCK_N, 0, 0), 0u);
                  k = 1;
                }
                for (; k < kBlockK / 16; ++k) {
                  const uint64_t ak = da + 2 * k, bk = db + 2 * k;
                  umma_bf16(d0, ak, bk, i0, 1u);
                  if (np > 1) umma_bf16(d1, ak, bk + o1, i1, 1u);
                  if (np > 2) umma_bf16(d2, ak, bk + o2, i2, 1u);
                }
              }
              if (++sl == static_cast<uint32_t>(SLOTS)) sl = 0;
            }
            RING_PROF_T(ts2);
            RING_PROF_ADD(m_mma, ts1, ts2);
            sl = gs;
            for (int cb = 0; cb < CB; ++cb) {                                    // this input row is dead now
              umma_commit(&empty_bar[sl]);
              if (++sl == static_cast<uint32_t>(SLOTS)) sl = 0;
            }
            if (j >= R - 1) umma_commit(&tfull_bar[(it0 + j - R + 1) & 7]);      // output row j - R + 1 is complete
            RING_PROF_T(ts3);
            RING_PROF_ADD(m_commit, ts2, ts3);
          }
          __syncwarp();
          RING_PROF_T(t3);
          RING_PROF_ADD(m_is, t2, t3);
          gs += CB;
          if (gs >= static_cast<uint32_t>(SLOTS)) { gs -= SLOTS; gp ^= 1u; }
        }
        it0 += nrows;
      }
      it = it0;
    } else
    for (int item = cta0; item < items; item += cta_stride) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      for (int i = 0; i < nrows; ++i, ++it) {
        const int as = it % NACC;
        RING_PROF_T(t0);
        mbar_wait(&tempty_bar[as], (((it / NACC) & 1) ^ 1u));
        RING_PROF_T(t1);
        for (int e = (i == 0 ? 0 : (R - 1) * CB); e < R * CB; ++e) {
          uint32_t sl, par;
          entry(e, sl, par);
          mbar_wait(&full_bar[sl], par);
        }
        RING_PROF_T(t2);
        RING_PROF_ADD(m_te, t0, t1);
        RING_PROF_ADD(m_fu, t1, t2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + as * BLOCK_N;
          for (int r = 0; r < R; ++r)
            for (int cb = 0; cb < CB; ++cb) {
              uint32_t sl, par;
              entry(r * CB + cb, sl, par);
              const uint32_t sa = ring_addr + sl * slot_bytes;
              for (int s = 0; s < S; ++s) {
                const uint64_t da = make_smem_desc(sa + s * 128, 0, 1024);      // s pixels into the strip
                const uint64_t db = make_smem_desc(w_addr + ((r * S + s) * CB + cb) * kRingWBytes, 0, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (r > 0 || cb > 0 || s > 0 || k > 0) ? 1u : 0u);
              }
            }
          // the first input row of this output row is dead now; after the item's last row so are the others
          for (int e = 0; e < (i == nrows - 1 ? R * CB : CB); ++e) {
            uint32_t sl, par;
            entry(e, sl, par);
            umma_commit(&empty_bar[sl]);
          }
          umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        RING_PROF_T(t3);
        RING_PROF_ADD(m_is, t2, t3);
        // the next output row starts one input row (CB entries) further
        g0s += CB;
        if (g0s >= static_cast<uint32_t>(SLOTS)) { g0s -= SLOTS; g0p ^= 1u; }
      }
      // ... and the next item after the R - 1 rows the last output row still used
      g0s += (R - 1) * CB;
      while (g0s >= static_cast<uint32_t>(SLOTS)) { g0s -= SLOTS; g0p ^= 1u; }
    }
#ifdef MSIG_RING_PROFILE
    if (lane == 0) {
      atomicAdd(&g_ring_prof[0], static_cast<unsigned long long>(m_te));
      atomicAdd(&g_ring_prof[1], static_cast<unsigned long long>(m_fu));
      atomicAdd(&g_ring_prof[2], static_cast<unsigned long long>(m_is));
      atomicAdd(&g_ring_prof[3], static_cast<unsigned long long>(it));
    }
    if (elect_one()) {       // (the stacked loop's finer counters live in the issuing thread)
      atomicAdd(&g_ring_prof[11], static_cast<unsigned long long>(m_setup));
      atomicAdd(&g_ring_prof[12], static_cast<unsigned long long>(m_mma));
      atomicAdd(&g_ring_prof[13], static_cast<unsigned long long>(m_commit));
    }
#endif
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;                 // which half of the 64 columns
    const int c_begin = half * (BLOCK_N / 2), c_end = c_begin + BLOCK_N / 2;
    const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
    // Cycle counters (make PROF=1, profiles/probe/ring_profile.py) put the bound of these layers on the MMA warp:
    // ~60 cycles per 128 x 64 x 16 MMA (the N = 64 instruction moves 6 KiB of shared-memory operands for 32 cycles
    // of math), 2.0-2.2 k cycles per tile against ~1 k for the lean epilogue below and ~3.4 k for the generic one.
    // Lean path = the common case (full 64-channel bf16 NHWC output, bias + ReLU / LeakyReLU / none, no second
    // operand, no fused statistics): bias in registers for the whole kernel, pixel coordinates straight from the
    // work item, the accumulator stage handed back as soon as it is in registers.
    const bool lean = epilogue_is_lean(p, BLOCK_N);
    float bias_r[BLOCK_N / 2];
#pragma unroll
    for (int j = 0; j < BLOCK_N / 2; ++j) bias_r[j] = (lean && p.bias != nullptr) ? __ldg(p.bias + c_begin + j) : 0.f;
    __nv_bfloat16* const out_ph = reinterpret_cast<__nv_bfloat16*>(p.out) + p.o_ph[ph] + c_begin;
    int it = 0;
#ifdef MSIG_RING_PROFILE
    long long e_w = 0, e_ld = 0, e_m = 0;
#endif
    if (p.ring_item_stats != 0 && p.stat_out != nullptr && p.stat_z != nullptr) {
      // dgrad + activation mask from z + the norm-backward reductions (sum g, sum g*z), one partial row per item
      const int rows_per_img = p.tiles_w * p.ring_chunks * NPH * 4;
      float* const msk_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 1024);   // 2 x [scale | shift]
      const float off = p.aux_mode == AUX_RELU_MASK ? 0.f : p.slope;
      const __nv_bfloat16* const z_ph = p.stat_z + p.o_ph[ph] + c_begin;
      int nitem = 0;
      for (int item = cta0; item < items; item += cta_stride, ++nitem) {
        int img, tw, h0, nrows;
        decode(item, img, tw, h0, nrows);
        const int ow = tw * kTileM + q * 32 + lane;
        const bool valid = ow < p.OW;
        float* const msk = msk_s + (nitem & 1) * 128;
        if (warp == 4) {                           // this image's mask coefficients (two buffers: see the barrier below)
          msk[lane] = __ldg(p.mask_scale + int64_t(img) * p.mask_ld + lane);
          msk[lane + 32] = __ldg(p.mask_scale + int64_t(img) * p.mask_ld + lane + 32);
          msk[64 + lane] = __ldg(p.mask_shift + int64_t(img) * p.mask_ld + lane);
          msk[64 + lane + 32] = __ldg(p.mask_shift + int64_t(img) * p.mask_ld + lane + 32);
        }
        // all eight epilogue warps: when any of them is past this barrier, all of them have finished the previous
        // item, so the buffer written for the item after this one is free again
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float s1[BLOCK_N / 2], s2[BLOCK_N / 2];
#pragma unroll
        for (int j = 0; j < BLOCK_N / 2; ++j) s1[j] = s2[j] = 0.f;
        for (int i = 0; i < nrows; ++i, ++it) {
          const int as = it % NACC;
          const int64_t off_o = img * p.o_sn + int64_t(h0 + i) * p.o_sh + int64_t(ow) * p.o_sw;
          uint4 zv[4];
          if (valid) {                             // issued ahead of the accumulator wait
            ldg256(z_ph + off_o, zv[0], zv[1]);
            ldg256(z_ph + off_o + 16, zv[2], zv[3]);
          }
          mbar_wait(&tfull_bar[as], (it / NACC) & 1);
          epilogue_lean_mask_stats<BLOCK_N / 2>(p, msk + c_begin, alpha, off,
                                                tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c_begin,
                                                &tempty_bar[as], valid, zv, out_ph + off_o, s1, s2);
        }
        const float t1 = warp_col_reduce32(s1, lane);
        const float t2 = warp_col_reduce32(s2, lane);
        const int li = tw * p.ring_chunks + item % p.ring_chunks;
        float* so = p.stat_out + ((int64_t(img) * rows_per_img + (li * NPH + ph) * 4 + q) * 2) * p.stat_ld + c_begin + lane;
        so[0] = t1;
        so[p.stat_ld] = t2;
      }
    } else if (p.ring_item_stats != 0 && p.stat_out != nullptr && epilogue_is_lean(p, BLOCK_N, true)) {
      // InstanceNorm statistics of the stored output, one partial row per (item, phase, quadrant): this thread's
      // pixel column over the item's rows in registers, one warp transpose-reduce per item.
      const int rows_per_img = p.tiles_w * p.ring_chunks * NPH * 4;
      for (int item = cta0; item < items; item += cta_stride) {
        int img, tw, h0, nrows;
        decode(item, img, tw, h0, nrows);
        const int ow = tw * kTileM + q * 32 + lane;
        float s1[BLOCK_N / 2], s2[BLOCK_N / 2];
#pragma unroll
        for (int j = 0; j < BLOCK_N / 2; ++j) s1[j] = s2[j] = 0.f;
        for (int i = 0; i < nrows; ++i, ++it) {
          const int as = it % NACC;
          mbar_wait(&tfull_bar[as], (it / NACC) & 1);
          __nv_bfloat16* o = out_ph + img * p.o_sn + int64_t(h0 + i) * p.o_sh + int64_t(ow) * p.o_sw;
          epilogue_lean_stats<BLOCK_N / 2>(p, bias_s + c_begin, alpha,
                                           tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c_begin,
                                           &tempty_bar[as], ow < p.OW, o, s1, s2);
        }
        const float t1 = warp_col_reduce32(s1, lane);
        const float t2 = warp_col_reduce32(s2, lane);
        const int li = tw * p.ring_chunks + item % p.ring_chunks;
        float* so = p.stat_out + ((int64_t(img) * rows_per_img + (li * NPH + ph) * 4 + q) * 2) * p.stat_ld + c_begin + lane;
        so[0] = t1;
        so[p.stat_ld] = t2;
      }
    } else
    for (int item = cta0; item < items; item += cta_stride) {
      int img, tw, h0, nrows;
      decode(item, img, tw, h0, nrows);
      const int ow = tw * kTileM + q * 32 + lane;
      for (int i = 0; i < nrows; ++i, ++it) {
        const int as = it % NACC;
        const uint32_t aphase = (it / NACC) & 1;
        // m-tile index in the generic numbering (TH = 1, TW = 128): (img, oh, tw), phase innermost
        const int mt = ((img * p.tiles_h + (h0 + i)) * p.tiles_w + tw) * NPH + ph;
        if (!lean) {
          fprop_epilogue_at<BLOCK_N>(p, mt, ph, img, h0 + i, ow, 0, tmem_base, as, aphase, tfull_bar, q, lane, c_begin,
                                     c_end, alpha);
          tc_fence_before();
          mbar_arrive(&tempty_bar[as]);
          continue;
        }
        RING_PROF_T(t0);
        mbar_wait(&tfull_bar[as], aphase);
        RING_PROF_T(t1);
        __nv_bfloat16* o = out_ph + img * p.o_sn + int64_t(h0 + i) * p.o_sh + int64_t(ow) * p.o_sw;
        epilogue_lean<BLOCK_N / 2>(p, bias_r, alpha, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + c_begin,
                                   &tempty_bar[as], ow < p.OW, o);
        RING_PROF_T(t2);
        RING_PROF_ADD(e_w, t0, t1);
        RING_PROF_ADD(e_ld, t1, t2);
        RING_PROF_T(t3);
        RING_PROF_ADD(e_m, t2, t3);
      }
    }
#ifdef MSIG_RING_PROFILE
    if (warp == 4 && lane == 0) {
      atomicAdd(&g_ring_prof[4], static_cast<unsigned long long>(e_w));
      atomicAdd(&g_ring_prof[5], static_cast<unsigned long long>(e_ld));
      atomicAdd(&g_ring_prof[6], static_cast<unsigned long long>(e_m));
      atomicAdd(&g_ring_prof[7], static_cast<unsigned long long>(it));
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (STACK) tmem_dealloc(tmem_base, 512); else tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------ wgrad
template <int BLOCK_N>
struct WgradCfg {
  static constexpr int kABytes = 2 * 8192;                 // 128 channels x 64 pixels
  static constexpr int kBBytes = (BLOCK_N / 64) * 8192;    // BLOCK_N channels x 64 pixels
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  pdl_entry();
  using Cfg = WgradCfg<BLOCK_N>;
  const int m_blk = blockIdx.x / p.n_blocks;
  const int n_blk = blockIdx.x % p.n_blocks;
  const int tap = blockIdx.y;
  const int split = blockIdx.z;
  if (p.upper_only && (n_blk + 1) * BLOCK_N <= m_blk * kTileM) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      prefetch_tmap(&p.tmA[i]);
      prefetch_tmap(&p.tmB[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.kb_total) kb1 = p.kb_total;
  const int nk = kb1 > kb0 ? kb1 - kb0 : 0;

  if (warp == 0) {
    const Tap ta = p.tapA[tap];
    const Tap tb = p.tapB[tap];
    const int per_img = p.blocks_h * p.blocks_w;
    int stage = 0;
    uint32_t phase = 0;
    // (img, hb, wb) of the first K block, then incremented (no division in the loop)
    int img = 0, rem = kb0;
    if (!p.fold_img) {
      img = kb0 / per_img;
      rem = kb0 - img * per_img;
    }
    int hb = rem / p.blocks_w;
    int wb = rem - hb * p.blocks_w;
    int vca[2], na[2], vcb[BLOCK_N / 64], nb[BLOCK_N / 64];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      vca[i] = m_blk * kTileM + 64 * i;
      na[i] = 0;
      if (p.fold_img) {
        na[i] = vca[i] / p.CA;
        vca[i] -= na[i] * p.CA;
      }
    }
#pragma unroll
    for (int i = 0; i < BLOCK_N / 64; ++i) {
      vcb[i] = n_blk * BLOCK_N + 64 * i;
      nb[i] = 0;
      if (p.fold_img) {
        nb[i] = vcb[i] / p.CB;
        vcb[i] -= nb[i] * p.CB;
      }
    }
    const int a_boxes = (p.a_boxes == 1 && !p.a_box_tap) ? 1 : 2;
    const uint32_t stage_tx = static_cast<uint32_t>(a_boxes * 8192 + Cfg::kBBytes);
    for (int kb = 0; kb < nk; ++kb) {
      const int oh0 = hb * p.PH, ow0 = wb * p.PW;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        mbar_expect_tx(&full_bar[stage], stage_tx);
        if (p.a_box_tap) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const Tap tai = p.tapA[tap * 2 + i];
            tma_load_4d(sa + i * 8192, &p.tmA[tai.map], &full_bar[stage], 0, ow0 + tai.dw, oh0 + tai.dh, img);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (i < a_boxes)
              tma_load_4d(sa + i * 8192, &p.tmA[ta.map], &full_bar[stage], vca[i], ow0 + ta.dw, oh0 + ta.dh,
                          p.fold_img ? na[i] : img);
        }
        if (p.b_box_tap) {
#pragma unroll
          for (int i = 0; i < BLOCK_N / 64; ++i) {
            const Tap tbi = p.tapB[tap * (BLOCK_N / 64) + i];
            tma_load_4d(sb + i * 8192, &p.tmB[tbi.map], &full_bar[stage], 0, ow0 + tbi.dw, oh0 + tbi.dh, img);
          }
        } else {
#pragma unroll
          for (int i = 0; i < BLOCK_N / 64; ++i)
            tma_load_4d(sb + i * 8192, &p.tmB[tb.map], &full_bar[stage], vcb[i], ow0 + tb.dw, oh0 + tb.dh,
                        p.fold_img ? nb[i] : img);
        }
      }
      __syncwarp();
      if (++wb == p.blocks_w) {
        wb = 0;
        if (++hb == p.blocks_h) {
          hb = 0;
          ++img;
        }
      }
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t da = make_smem_desc(sa, 8192, 1024);
        const uint64_t db = make_smem_desc(sa + Cfg::kABytes, 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 16 pixel rows (2 swizzle atoms of 8 rows = 2048 B) per MMA
          umma_bf16(tmem_base, da + 128 * k, db + 128 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (kb == nk - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int m = m_blk * kTileM + q * 32 + lane;
    const bool row_valid = m < p.m_valid;
    float* orow = p.out + split * p.o_split + tap * p.o_tap + static_cast<int64_t>(m) * p.o_row;
    if (nk > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t r[32];
      if (nk > 0) {
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row_valid) {
        const int col0 = n_blk * BLOCK_N + c;
        if (col0 + 32 <= p.n_valid && (p.o_row & 7) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {        // one full 32-byte sector per store
            uint4 a, b;
            a.x = __float_as_uint(__uint_as_float(r[j]) * p.alpha);
            a.y = __float_as_uint(__uint_as_float(r[j + 1]) * p.alpha);
            a.z = __float_as_uint(__uint_as_float(r[j + 2]) * p.alpha);
            a.w = __float_as_uint(__uint_as_float(r[j + 3]) * p.alpha);
            b.x = __float_as_uint(__uint_as_float(r[j + 4]) * p.alpha);
            b.y = __float_as_uint(__uint_as_float(r[j + 5]) * p.alpha);
            b.z = __float_as_uint(__uint_as_float(r[j + 6]) * p.alpha);
            b.w = __float_as_uint(__uint_as_float(r[j + 7]) * p.alpha);
            stg256(orow + col0 + j, a, b);
          }
        } else if (col0 + 32 <= p.n_valid && (p.o_row & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(r[j]) * p.alpha;
            o.y = __uint_as_float(r[j + 1]) * p.alpha;
            o.z = __uint_as_float(r[j + 2]) * p.alpha;
            o.w = __uint_as_float(r[j + 3]) * p.alpha;
            *reinterpret_cast<float4*>(orow + col0 + j) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.n_valid) orow[col0 + j] = __uint_as_float(r[j]) * p.alpha;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// ------------------------------------------------------------------------------------ CTA-pair wgrad
// wgrad2_kernel: the 256 x 256 weight-gradient tile (M = 256 channels of the A side, N = 256 of the B
// side) of one (tap, K split) on a CTA PAIR: one tcgen05.mma.cta_group::2 per 16 pixels covers both CTAs'
// 128 M rows; each CTA stages its own 128-channel A tile and HALF of the B tile (32 KiB instead of
// 48 KiB per 64-pixel K block per SM), 6-stage ring. Rank 0 issues the MMAs. Used for the residual
// blocks' 3x3 256->256 layers (m_blocks == 2, n_blocks == 1).
constexpr int kW2Stages = 6;
constexpr int kW2StageBytes = 4 * 8192;                 // A: 2 x 64 ch x 64 px, B half: 2 x 64 ch x 64 px
constexpr int kW2SmemBytes = kW2Stages * kW2StageBytes + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) wgrad2_kernel(const __grid_constant__ WgradParams p) {
  pdl_entry();
  constexpr int BLOCK_N = 256;
  const uint32_t rank = cluster_ctarank();               // == m_blk
  const int m_blk = static_cast<int>(rank);
  const int tap = blockIdx.y;
  const int split = blockIdx.z;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kW2Stages * kW2StageBytes);
  uint64_t* empty_bar = full_bar + kW2Stages;
  uint64_t* tfull_bar = empty_bar + kW2Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      prefetch_tmap(&p.tmA[i]);
      prefetch_tmap(&p.tmB[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kW2Stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_slot, BLOCK_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.kb_total) kb1 = p.kb_total;
  const int nk = kb1 > kb0 ? kb1 - kb0 : 0;

  if (warp == 0) {
    const Tap ta = p.tapA[tap];
    const Tap tb = p.tapB[tap];
    const int per_img = p.blocks_h * p.blocks_w;
    int stage = 0;
    uint32_t phase = 0;
    int img = kb0 / per_img;
    const int rem = kb0 - img * per_img;
    int hb = rem / p.blocks_w;
    int wb = rem - hb * p.blocks_w;
    for (int kb = 0; kb < nk; ++kb) {
      const int oh0 = hb * p.PH, ow0 = wb * p.PW;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      if (elect_one()) {
        uint8_t* sa = smem + stage * kW2StageBytes;
        uint8_t* sb = sa + 2 * 8192;
        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kW2StageBytes);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          tma_load_4d_2sm(sa + i * 8192, &p.tmA[ta.map], &full_bar[stage], m_blk * kTileM + 64 * i, ow0 + ta.dw,
                          oh0 + ta.dh, img);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          tma_load_4d_2sm(sb + i * 8192, &p.tmB[tb.map], &full_bar[stage], static_cast<int>(rank) * 128 + 64 * i,
                          ow0 + tb.dw, oh0 + tb.dh, img);
      }
      __syncwarp();
      if (++wb == p.blocks_w) {
        wb = 0;
        if (++hb == p.blocks_h) {
          hb = 0;
          ++img;
        }
      }
      if (++stage == kW2Stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nk; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * kW2StageBytes);
        const uint64_t da = make_smem_desc(sa, 8192, 1024);
        const uint64_t db = make_smem_desc(sa + 2 * 8192, 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_2sm(tmem_base, da + 128 * k, db + 128 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage], 3);
        if (kb == nk - 1) umma_commit_2sm(tfull_bar, 3);
      }
      __syncwarp();
      if (++stage == kW2Stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int m = m_blk * kTileM + q * 32 + lane;
    const bool row_valid = m < p.m_valid;
    float* orow = p.out + split * p.o_split + tap * p.o_tap + static_cast<int64_t>(m) * p.o_row;
    if (nk > 0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t r[32];
      if (nk > 0) {
        tmem_ld_32x32(taddr + c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row_valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 a, b;
          a.x = __float_as_uint(__uint_as_float(r[j]) * p.alpha);
          a.y = __float_as_uint(__uint_as_float(r[j + 1]) * p.alpha);
          a.z = __float_as_uint(__uint_as_float(r[j + 2]) * p.alpha);
          a.w = __float_as_uint(__uint_as_float(r[j + 3]) * p.alpha);
          b.x = __float_as_uint(__uint_as_float(r[j + 4]) * p.alpha);
          b.y = __float_as_uint(__uint_as_float(r[j + 5]) * p.alpha);
          b.z = __float_as_uint(__uint_as_float(r[j + 6]) * p.alpha);
          b.w = __float_as_uint(__uint_as_float(r[j + 7]) * p.alpha);
          stg256(orow + c + j, a, b);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, BLOCK_N);
  }
}

// ------------------------------------------------------------------------------------ launch
// cudaFuncSetAttribute is per device: one bit per device ordinal records where it has been applied.
static int current_dev_bit() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
static bool attr_needed(uint64_t mask) { return ((mask >> current_dev_bit()) & 1ull) == 0; }
static void attr_done(uint64_t& mask) { mask |= 1ull << current_dev_bit(); }

template <int BLOCK_N>
static cudaError_t launch_fprop_t(const FpropParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = FpropCfg<BLOCK_N>;
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(fprop_kernel<BLOCK_N>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  const int total = p.n_img * p.tiles_h * p.tiles_w * p.phases * p.n_blocks;
  const int grid = total < num_sms ? total : num_sms;
  if (grid <= 0) return cudaSuccess;
  MSIG_LAUNCH((fprop_kernel<BLOCK_N>), grid, kFpropThreads, Cfg::kSmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

static bool g_pair_mode = true;
static int g_ring_slots_cap = 0;      // test hook: > 0 caps the ring depth (8 = the fixed depth of earlier builds)
void set_ring_slots_cap(int n) { g_ring_slots_cap = n; }
static bool g_ring_legacy = false;    // test hook: per-output-row N = 64 MMAs
void set_ring_legacy(bool on) { g_ring_legacy = on; }
void set_pair_mode(bool on) { g_pair_mode = on; }

static cudaError_t launch_fprop2(const FpropParams& p, int num_sms, cudaStream_t stream) {
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(fprop2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  const int pairs = (p.n_img * p.tiles_h * p.tiles_w / 2) * p.n_blocks;
  int clusters = num_sms / 2;
  if (pairs < clusters) clusters = pairs;
  if (clusters <= 0) return cudaSuccess;
  MSIG_LAUNCH((fprop2_kernel), 2 * clusters, kFpropThreads, k2SmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

// CTA pairs for the wide tiles whenever both CTAs of a pair can share the weight tile. (The weight
// tensor map of a paired launch must have a 128-row box: each CTA loads half of the 256 rows.)
static bool g_m2_mode = true;
void set_m2_mode(bool on) { g_m2_mode = on; }

// Two m-tiles per CTA for the N = 128 layers whenever tile rows pair up inside an image and there is at
// least one full wave of (pair, n block) work items.
bool fprop_uses_m2(const FpropParams& p, int block_n) {
  if (!g_m2_mode || block_n != 128 || p.tap_is_image || p.fold_c != 0 || p.b_row_per_image != 0) return false;
  if ((p.tiles_h % 2) != 0 || 2 * p.TH > 256 || p.taps * p.cblocks < 4) return false;
  const int64_t items = int64_t(p.n_img) * (p.tiles_h / 2) * p.tiles_w * p.phases * p.n_blocks;
  return items >= 148;
}

static cudaError_t launch_fprop_m2(const FpropParams& p, int num_sms, cudaStream_t stream) {
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(fprop_m2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kM2SmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  const int total = p.n_img * (p.tiles_h / 2) * p.tiles_w * p.phases * p.n_blocks;
  const int grid = total < num_sms ? total : num_sms;
  if (grid <= 0) return cudaSuccess;
  MSIG_LAUNCH((fprop_m2_kernel), grid, kFpropThreads, kM2SmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

bool fprop_uses_pairs(const FpropParams& p, int block_n) {
  return g_pair_mode && block_n == 256 && p.phases == 1 && p.b_row_per_image == 0 &&
         ((p.n_img * p.tiles_h * p.tiles_w) % 2) == 0 && p.taps * p.cblocks >= 4;
}

cudaError_t launch_fprop(const FpropParams& p, int block_n, int num_sms, cudaStream_t stream) {
  if (p.ring_item_stats != 0) return cudaErrorInvalidValue;   // per-item statistics rows exist in the ring kernel only
  if (fprop_uses_pairs(p, block_n)) return launch_fprop2(p, num_sms, stream);
  if (p.m2 != 0) return block_n == 128 ? launch_fprop_m2(p, num_sms, stream) : cudaErrorInvalidValue;
  switch (block_n) {
    case 16: return launch_fprop_t<16>(p, num_sms, stream);
    case 64: return launch_fprop_t<64>(p, num_sms, stream);
    case 128: return launch_fprop_t<128>(p, num_sms, stream);
    case 256: return launch_fprop_t<256>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_rowfold(const RowfoldParams& p, int num_sms, cudaStream_t stream) {
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(fprop_rowfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kRfSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  if (p.R < 1 || p.R > kRfMaxR || p.S < 1 || p.S * 4 > 32 || p.n_valid > 4 || p.R + 1 > kRfRing)
    return cudaErrorInvalidValue;
  const int items = p.n_img * p.tiles_w * p.chunks_h;
  const int grid = items < num_sms ? items : num_sms;
  if (grid <= 0) return cudaSuccess;
  MSIG_LAUNCH((fprop_rowfold_kernel), grid, kRfThreads, kRfSmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

#ifdef MSIG_RING_PROFILE
extern "C" int msig_debug_ring_profile(unsigned long long* out16, int reset) {
  unsigned long long z[16] = {0};
  if (out16 && cudaMemcpyFromSymbol(out16, g_ring_prof, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(g_ring_prof, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif

// Ring slots that fit next to the resident filter (0: the shape does not fit the ring kernel at all).
int ring_slots_for(int R, int S, int cbs) {
  if (R < 1 || S < 1 || R * S * cbs > kRingMaxTaps || R * S > 16) return 0;
  const int room = kRingSmemMax - 1024 - kRingBarBytes - R * S * cbs * kRingWBytes;
  int slots = room / ring_slot_bytes(S);
  if (slots > kRingMaxSlots) slots = kRingMaxSlots;
  return slots >= (R + 1) * cbs ? slots : 0;
}

cudaError_t launch_fprop_ring64(const FpropParams& p0, int num_sms, cudaStream_t stream) {
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(fprop_ring64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kRingSmemMax);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  FpropParams p = p0;
  const int cbs = p.ring_cb > 1 ? p.ring_cb : 1;
  const int nph = p.ring_phases > 1 ? p.ring_phases : 1;
  const int slots = ring_slots_for(p.strip_r, p.strip_s, cbs);
  if (slots == 0 || p.TW != kTileM || p.TH != 1 || p.phases != nph || (nph != 1 && nph != 4) || p.n_blocks != 1)
    return cudaErrorInvalidValue;
  p.ring_stack = g_ring_legacy ? 0 : 1;
  if (p.ring_item_stats != 0) {         // requested rows per image (msig_epilogue.stats_rows) must be what this launch writes
    const bool plain = p.stat_z == nullptr && p.aux_mode == AUX_NONE;                 // statistics of the output
    const bool masked = p.stat_z != nullptr && p.mask_scale != nullptr && p.mask_shift != nullptr && p.mask_ld >= 64 &&
                        (p.aux_mode == AUX_RELU_MASK || p.aux_mode == AUX_LRELU_MASK) && p.bias == nullptr &&
                        p.act == ACT_NONE && p.z_mask == 0;                           // norm-backward reductions
    if (p.stat_out == nullptr || !(plain || masked) || p.ring_item_stats != p.tiles_w * p.ring_chunks * nph * 4 ||
        p.n_valid != 64 || p.out_f32 || p.o_sc != 1 || p.act == ACT_TANH || p.fold_c != 0)
      return cudaErrorInvalidValue;
    p.ring_item_stats = 1;
  }
  p.ring_slots = g_ring_slots_cap > 0 && g_ring_slots_cap < slots ? g_ring_slots_cap : slots;
  if (p.ring_slots < (p.strip_r + 1) * cbs) p.ring_slots = (p.strip_r + 1) * cbs;
  const int items = p.n_img * p.tiles_w * p.ring_chunks;
  int grid = items * nph < num_sms ? items * nph : num_sms / nph * nph;
  if (grid <= 0) return cudaSuccess;
  const int smem_bytes = p.strip_r * p.strip_s * cbs * kRingWBytes + p.ring_slots * ring_slot_bytes(p.strip_s) + 1024 +
                         kRingBarBytes;
  MSIG_LAUNCH((fprop_ring64_kernel), grid, kFpropThreads, smem_bytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

template <int BLOCK_N>
static cudaError_t launch_wgrad_t(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<BLOCK_N>;
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BLOCK_N>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  dim3 grid(p.m_blocks * p.n_blocks, p.taps, p.splits);
  MSIG_LAUNCH((wgrad_kernel<BLOCK_N>), grid, 256, Cfg::kSmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

static cudaError_t launch_wgrad2(const WgradParams& p, cudaStream_t stream) {
  static uint64_t attr_devs = 0;
  if (attr_needed(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2SmemBytes);
    if (e != cudaSuccess) return e;
    attr_done(attr_devs);
  }
  dim3 grid(2, p.taps, p.splits);
  MSIG_LAUNCH((wgrad2_kernel), grid, 256, kW2SmemBytes, stream, p);
  count_launch(1);
  return cudaGetLastError();
}

cudaError_t launch_wgrad(const WgradParams& p, int block_n, cudaStream_t stream) {
  // CTA pairs for the full 256 x 256 tiles (both m-blocks of one (tap, split) share the B tile)
  if (g_pair_mode && block_n == 256 && p.m_blocks == 2 && p.n_blocks == 1 && !p.fold_img && !p.upper_only &&
      !p.b_box_tap && !p.a_box_tap && p.a_boxes != 1 && p.m_valid == 256 && p.n_valid == 256 && (p.o_row & 7) == 0)
    return launch_wgrad2(p, stream);
  switch (block_n) {
    case 64: return launch_wgrad_t<64>(p, stream);
    case 128: return launch_wgrad_t<128>(p, stream);
    case 256: return launch_wgrad_t<256>(p, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace msig
