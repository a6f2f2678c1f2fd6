// GPU-side training augmentation: uint8 HWC image batch -> RandomResizedCrop (given crop boxes) ->
// rotation by a multiple of 90 degrees -> ToTensor -> Normalize(0.5, 0.5) -> fp32 NCHW in [-1, 1].
//
// Replaces the per-image CPU transform pipeline of /root/reference/dataset.py:16-22 (PIL + torchvision in
// four DataLoader worker processes, /root/reference/trainer.py:287-290) for the deterministic part; the
// random draws (crop box, quarter turns) stay on the host (msig_b200/augment.py). Results are BIT-EXACT with
// Pillow's 8-bit bilinear resample (ImagingResample: double-precision triangle-filter coefficients,
// normalised, quantised to 22-bit fixed point; horizontal pass rounded to uint8, then vertical pass rounded
// to uint8) followed by torchvision's to_tensor / normalize: the coefficient arithmetic below uses explicit
// round-to-nearest double intrinsics in Pillow's operation order (no FMA contraction), the accumulation is
// integer. Byte work, HBM / latency bound; a batch of 32 256x256 images is 6 MB in and 25 MB out.
#include "common.h"

namespace msig {

constexpr int kAugPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS
constexpr int kAugMaxTaps = 17;                 // ceil(support) * 2 + 1 with support <= 8 (shrink factor <= 8)

struct AugCoef {
  int xmin, count;
  int kk[kAugMaxTaps];
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output index xx of an axis resampled from in_size to
// out_size samples (bilinear filter, box = the whole cropped axis).
__device__ __forceinline__ void aug_coeffs(int in_size, int out_size, int xx, AugCoef& c) {
  const double scale = __ddiv_rn(static_cast<double>(static_cast<float>(in_size)), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // filter support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dadd_rn(0.0, __dmul_rn(static_cast<double>(xx) + 0.5, scale));
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kAugMaxTaps) xmax = kAugMaxTaps;               // (excluded by the host-side shape check)
  double w[kAugMaxTaps];
  double ww = 0.0;
#pragma unroll 1
  for (int x = 0; x < xmax; ++x) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
    if (a < 0.0) a = -a;
    w[x] = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
    ww = __dadd_rn(ww, w[x]);
  }
#pragma unroll 1
  for (int x = 0; x < xmax; ++x) {
    double v = w[x];
    if (ww != 0.0) v = __ddiv_rn(v, ww);
    const double s = __dmul_rn(v, static_cast<double>(1 << kAugPrecisionBits));
    c.kk[x] = v < 0.0 ? __double2int_rz(__dadd_rn(-0.5, s)) : __double2int_rz(__dadd_rn(0.5, s));
  }
  c.xmin = xmin;
  c.count = xmax;
}

__device__ __forceinline__ uint8_t aug_clip8(int v) {
  v >>= kAugPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct AugBox {
  int top, left, h, w;
};
__device__ __forceinline__ AugBox aug_box(const int32_t* boxes, int b, int H, int W) {
  AugBox r;
  r.top = min(max(boxes[4 * b + 0], 0), H - 1);
  r.left = min(max(boxes[4 * b + 1], 0), W - 1);
  r.h = min(max(boxes[4 * b + 2], 1), H - r.top);
  r.w = min(max(boxes[4 * b + 3], 1), W - r.left);
  return r;
}

// Horizontal pass: tmp[b][y][xx][c] = clip8(sum_x src[b][top + y][left + xmin + x][c] * kk[x]) for the crop's rows.
// grid (row tiles, n); a thread owns output column(s) xx and walks the rows of its tile.
__global__ void __launch_bounds__(256) augment_h_kernel(const uint8_t* __restrict__ src, int H, int W,
                                                        const int32_t* __restrict__ boxes, int size,
                                                        uint8_t* __restrict__ tmp, int rows_per_block) {
  pdl_entry();
  const int b = blockIdx.y;
  const AugBox bx = aug_box(boxes, b, H, W);
  const int y0 = blockIdx.x * rows_per_block;
  if (y0 >= bx.h) return;
  const int y1 = min(y0 + rows_per_block, bx.h);
  const uint8_t* sb = src + (int64_t(b) * H + bx.top) * W * 3 + int64_t(bx.left) * 3;
  uint8_t* tb = tmp + int64_t(b) * H * size * 3;
  for (int xx = threadIdx.x; xx < size; xx += blockDim.x) {
    AugCoef c;
    aug_coeffs(bx.w, size, xx, c);
    for (int y = y0; y < y1; ++y) {
      const uint8_t* row = sb + int64_t(y) * W * 3 + c.xmin * 3;
      int s0 = 1 << (kAugPrecisionBits - 1), s1 = s0, s2 = s0;
      for (int x = 0; x < c.count; ++x) {
        s0 += int(row[3 * x + 0]) * c.kk[x];
        s1 += int(row[3 * x + 1]) * c.kk[x];
        s2 += int(row[3 * x + 2]) * c.kk[x];
      }
      uint8_t* o = tb + (int64_t(y) * size + xx) * 3;
      o[0] = aug_clip8(s0);
      o[1] = aug_clip8(s1);
      o[2] = aug_clip8(s2);
    }
  }
}

// Vertical pass + counter-clockwise rotation by q quarter turns + ToTensor + Normalize(0.5, 0.5):
// out[b][c][i][j] fp32. grid (row tiles, n); a thread owns column(s) xx of the resized image.
__global__ void __launch_bounds__(256) augment_v_kernel(const uint8_t* __restrict__ tmp, int H, int W,
                                                        const int32_t* __restrict__ boxes,
                                                        const int32_t* __restrict__ quarter_turns, int size,
                                                        float* __restrict__ out, int rows_per_block) {
  pdl_entry();
  const int b = blockIdx.y;
  const AugBox bx = aug_box(boxes, b, H, W);
  const int q = quarter_turns ? (quarter_turns[b] & 3) : 0;
  const int yy0 = blockIdx.x * rows_per_block;
  const int yy1 = min(yy0 + rows_per_block, size);
  const uint8_t* tb = tmp + int64_t(b) * H * size * 3;
  float* ob = out + int64_t(b) * 3 * size * size;
  const int64_t plane = int64_t(size) * size;
  for (int yy = yy0; yy < yy1; ++yy) {
    AugCoef c;
    aug_coeffs(bx.h, size, yy, c);
    for (int xx = threadIdx.x; xx < size; xx += blockDim.x) {
      int s0 = 1 << (kAugPrecisionBits - 1), s1 = s0, s2 = s0;
      for (int y = 0; y < c.count; ++y) {
        const uint8_t* p = tb + (int64_t(c.xmin + y) * size + xx) * 3;
        s0 += int(p[0]) * c.kk[y];
        s1 += int(p[1]) * c.kk[y];
        s2 += int(p[2]) * c.kk[y];
      }
      // PIL Transpose.ROTATE_90 / 180 / 270 (counter-clockwise): destination of source pixel (yy, xx)
      int i = yy, j = xx;
      if (q == 1) { i = size - 1 - xx; j = yy; }
      else if (q == 2) { i = size - 1 - yy; j = size - 1 - xx; }
      else if (q == 3) { i = xx; j = size - 1 - yy; }
      const int64_t o = int64_t(i) * size + j;
      const int v[3] = {aug_clip8(s0), aug_clip8(s1), aug_clip8(s2)};
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float t = __fdiv_rn(static_cast<float>(v[ch]), 255.0f);         // to_tensor
        ob[ch * plane + o] = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);             // normalize(0.5, 0.5)
      }
    }
  }
}

}  // namespace msig

using namespace msig;

extern "C" {

size_t msig_augment_workspace(int32_t n, int32_t h, int32_t w, int32_t size) {
  (void)w;
  return size_t(n) * h * size * 3;       // the horizontally resampled crop rows, uint8
}

int msig_augment_u8(const void* src, int32_t n, int32_t h, int32_t w, const int32_t* boxes,
                    const int32_t* quarter_turns, int32_t size, float* out, void* workspace,
                    size_t workspace_bytes, void* stream) {
  MSIG_REQUIRE(src && boxes && out && workspace && n >= 1 && h >= 1 && w >= 1 && size >= 1,
               "msig_augment_u8: bad argument");
  MSIG_REQUIRE(h <= 8 * size && w <= 8 * size,
               "msig_augment_u8: shrink factor above 8 (%dx%d -> %d) is not supported", h, w, size);
  MSIG_REQUIRE(workspace_bytes >= msig_augment_workspace(n, h, w, size), "msig_augment_u8: workspace too small");
  const int rpb = 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MSIG_LAUNCH((augment_h_kernel), dim3(static_cast<unsigned>(ceil_div(h, rpb)), n), 256, 0, st, 
      static_cast<const uint8_t*>(src), h, w, boxes, size, static_cast<uint8_t*>(workspace), rpb);
  MSIG_CHECK_LAUNCH();
  MSIG_LAUNCH((augment_v_kernel), dim3(static_cast<unsigned>(ceil_div(size, rpb)), n), 256, 0, st, 
      static_cast<const uint8_t*>(workspace), h, w, boxes, quarter_turns, size, out, rpb);
  count_launch(2);
  MSIG_CHECK_LAUNCH();
  return MSIG_OK;
}

}  // extern "C"
