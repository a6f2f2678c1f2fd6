// Parameter blocks of the two tcgen05 implicit-GEMM kernels (see igemm.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace msig {

constexpr int kMaxTaps = 64;

// One filter tap of an implicit GEMM: where the operand tile of this tap sits relative to the
// output tile (in the coordinates of tensor map `map`).
struct Tap {
  int8_t dh, dw, map, pad_;
};

enum : int32_t { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };
enum : int32_t { AUX_NONE = 0, AUX_ADD = 1, AUX_RELU_MASK = 2, AUX_LRELU_MASK = 3 };

// ---- "fprop" kernel: D[pixels, N] = sum_{tap, c} A[pixel + tap, c] * B[n, (tap, c)]
// A: bf16 NHWC activations read through 4-D TMA maps (dims C,W,H,N), one 128-pixel x 64-channel
//    box per K block; out-of-bounds rows/cols are zero-filled by TMA (= zero padding).
// B: bf16 [N_total][K_total] K-major packed weights (2-D TMA map).
// Used for: conv forward, conv dgrad, transposed conv (4 output phases), 1x1 / Linear GEMMs,
// Gram backward (taps = images).
struct FpropParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  Tap tap[kMaxTaps];
  int32_t taps, cblocks;          // K blocks = taps * cblocks, 64 channels each (taps per phase)
  int32_t phases;                 // 1, or 4 output phases of a stride-2 transposed conv / dgrad:
                                  //   phase ph uses taps [ph*taps, (ph+1)*taps), B rows offset by
                                  //   ph*b_row_per_phase and the output / aux view offset o_ph/a_ph
  int32_t b_row_per_phase;
  int64_t o_ph[4], a_ph[4];
  int32_t tap_is_image;           // Gram backward: tap t reads image t (not the tile's image)
  int32_t b_row_per_image;        // B row offset added per output image (Gram backward)
  int32_t fold_c;                 // >0: GEMM columns are (image, channel) pairs, fold_c channels per image;
                                  //     column j is stored to image j / fold_c (Gram backward with N spanning images)
  int32_t strip_r, strip_s;       // strip / ring kernels: filter extent; tmA[1] has a (128 + strip_s - 1)-pixel box
  int32_t ring_rows, ring_chunks; // ring kernel: output rows per work item, items per image column
  int32_t org_h, org_w;           // ring kernel: input coordinate read by output (0,0) through tap (0,0)
  int32_t ring_cb;                // ring kernel: 64-channel blocks of the input (1 or 2; 0 means 1)
  int8_t ring_tap[16];            // ring kernel: filter position (r*S + s) -> tap index in the packed weights
  int32_t ring_slots;             // ring kernel: strips resident in shared memory (set by the launcher)
  int32_t ring_item_stats;        // ring kernel: 1 = stat_out holds ONE partial row per (work item, phase, accumulator
                                  //   quadrant): the lean epilogue sums its pixels over the item's rows in registers
  int32_t ring_stack;             // ring kernel: 1 = one MMA per INPUT row strip over all the output rows it feeds
                                  //   (N = 64 x rows, accumulators of consecutive rows side by side in TMEM)
  int32_t ring_phases;            // ring kernel: 0/1, or 4 = CTA b runs output phase b % 4 (`phases` must be 4 too):
  int8_t ring_org_h[4], ring_org_w[4];   //   per-phase org_h / org_w; weights of phase ph start at B row ph*b_row_per_phase
  int32_t OH, OW, TH, TW;         // output plane and the 128-pixel tile (TH*TW == 128)
  int32_t tiles_h, tiles_w, n_img, n_blocks;
  // epilogue
  void* out;
  int64_t o_sn, o_sh, o_sw, o_sc;  // element strides of the output view
  int32_t out_f32, n_valid;
  const float* bias;
  float alpha;
  const float* alpha_ptr;          // optional device scalar folded into alpha
  int32_t act;
  float slope;
  const __nv_bfloat16* aux;        // channel-contiguous tensor congruent with the output tile
  int64_t a_sn, a_sh, a_sw;
  int32_t aux_mode;
  // per-(image, channel) partial sums of the stored values v, produced by the epilogue (bf16 NHWC
  // outputs only): stat_out[(m_tile*4 + warp)*2 + q][stat_ld], q = 0: sum v, q = 1: sum v*v
  // (stat_z == nullptr: InstanceNorm statistics) or sum v*z (stat_z congruent with the output:
  // the two reductions of the InstanceNorm / AdaIN backward). An m-tile never spans two images.
  float* stat_out;
  const __nv_bfloat16* stat_z;
  int32_t stat_ld;
  // z_mask = 1 (stat_out == nullptr): the result (after alpha / bias / aux) is multiplied by (stat_z > 0): a
  // ReLU mask taken from a SECOND operand next to an additive aux (Gram backward + the ReLU backward of the
  // tapped feature map, losses.py:70-89 with the ReLUs of the VGG stack).
  int32_t z_mask;
  // mask_scale / mask_shift (fp32 [image][mask_ld], both or neither): with aux_mode = AUX_RELU_MASK / AUX_LRELU_MASK
  // the mask is act'(stat_z * scale + shift) -- recomputed from the norm INPUT z the reductions read anyway and
  // the per-(image, channel) scale / shift of its InstanceNorm / AdaIN -- and `aux` is not read at all.
  const float* mask_scale;
  const float* mask_shift;
  int32_t mask_ld;
  // m2 = 1: run on fprop_m2_kernel -- one CTA computes TWO vertically adjacent 128-pixel m-tiles (2*th, 2*th+1)
  // against the same weight tile; tmA maps then carry a {64, TW, 2*TH} box (see fprop_uses_m2).
  int32_t m2;
};

// ---- "wgrad" kernel: D[m, n] = sum_{pixels} A[pixel + tapA, m] * B[pixel + tapB, n]
// Both operands are pixel-major NHWC tiles (MN-major UMMA operands). Used for weight gradients
// (A = dy, B = x), and for Gram matrices (A = B = features, channels folded with images).
struct WgradParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  Tap tapA[kMaxTaps];
  Tap tapB[kMaxTaps];
  int32_t taps;
  int32_t PW, PH;                  // 64-pixel K block (PW*PH == 64)
  int32_t blocks_w, blocks_h, n_img;
  int32_t fold_img;                // 1: "channel" index = img*C + c, K loop stays inside an image
  int32_t CA, CB;                  // channels per image of A / B (fold_img only)
  int32_t m_blocks, n_blocks, splits, kb_per_split, kb_total;
  float* out;                      // fp32 [split][tap][m][n]
  int64_t o_split, o_tap, o_row;
  float alpha;
  int32_t m_valid, n_valid;
  int32_t upper_only;              // Gram: skip tiles strictly below the diagonal
  int32_t a_boxes;                 // 0/2: both 64-channel A boxes are loaded; 1: only the first (M <= 64 valid rows;
                                   //      rows 64..127 of the tile are never stored)
  int32_t a_box_tap;               // 1: the two 64-row A boxes are the SAME 64 channels read through different taps:
                                   //    box i of CTA-tap y uses tapA[2*y + i] (row-patch weight gradients: the
                                   //    second box is the first shifted up by four rows, so rows 64..127 of the
                                   //    tile accumulate the NEXT four filter rows against the same B boxes)
  int32_t b_box_tap;               // 1: the BLOCK_N/64 boxes of the B tile are the SAME 64 "channels" read through
                                   //    different taps: box i of CTA-tap y uses tapB[y*(BLOCK_N/64) + i]
                                   //    (row-patch weight gradients: 4 filter rows per CTA)
};

// ---- "row-fold" kernel: stride-1 convolution with <= 4 output channels (the generator's final 7x7
// 64->3 conv, model.py:141, and the input gradient of its first 7x7 3->64 conv, model.py:131).
// The S horizontal taps are folded into the GEMM N dimension: for one output row,
//   Q[q, (s, co)] = sum_{r, c} x[oh + r, q, c] * w[r][(s, co)][c]          (N = 4*S <= 32, K = R*64)
// is accumulated in TMEM over the R input rows (one 128-pixel TMA strip each, kept in a shared-memory
// ring so that consecutive output rows re-use R-1 of them), and the epilogue finishes
//   y[oh, ow, co] = act(bias[co] + sum_s Q[ow + s, (s, co)])
// through a shared-memory transpose. One tile = 128 - (S-1) output pixels of one row.
struct RowfoldParams {
  CUtensorMap tmA;                 // input [64 ch, W, H, N], box {64, 128, 1, 1}
  CUtensorMap tmB;                 // packed weights [R*32 rows][64], box {64, 32}
  int32_t R, S;
  int32_t org_h, org_w;            // input coordinate read by output (0,0) through tap (0,0) (= -pad)
  int32_t OH, OW, n_img;
  int32_t tiles_w, rows_per_item, chunks_h;   // work item = (image, column tile, chunk of output rows)
  float* out;                      // fp32
  int64_t o_sn, o_sh, o_sw, o_sc;
  int32_t n_valid;                 // output channels (<= 4)
  const float* bias;
  float alpha;
  const float* alpha_ptr;
  int32_t act;
  const float* ch_scale;           // optional per-output-channel multiplier applied with alpha (before the bias)
};

cudaError_t launch_fprop(const FpropParams& p, int block_n, int num_sms, cudaStream_t stream);
cudaError_t launch_rowfold(const RowfoldParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_fprop_ring64(const FpropParams& p, int num_sms, cudaStream_t stream);
int ring_slots_for(int R, int S, int cbs);   // ring depth the shape gets (0: does not fit the ring kernel)
void set_ring_slots_cap(int n);              // test hook: > 0 caps the ring depth
void set_ring_legacy(bool on);               // test hook: per-output-row N = 64 MMAs (the round-1 issue order)
cudaError_t launch_wgrad(const WgradParams& p, int block_n, cudaStream_t stream);
bool fprop_uses_pairs(const FpropParams& p, int block_n);
bool fprop_uses_m2(const FpropParams& p, int block_n);   // call with tiles / phases / taps / n_blocks already set
void set_m2_mode(bool on);
void set_pair_mode(bool on);   // test hook: CTA-pair (cta_group::2) kernel for 256-wide tiles on/off
int igemm_kernel_launches();  // launches issued since process start (bench: gpu_launches)

}  // namespace msig
