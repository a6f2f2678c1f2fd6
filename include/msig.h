/* msig.h — C ABI of the B200-native operator library behind the drop-in
 * Generator / StyleEncoder / Discriminator / VGG-loss modules.
 *
 * The reference (/root/reference) has no FFI: its arithmetic is issued through torch.nn
 * (cuDNN / cuBLAS / ATen). Each entry point below names the reference call sites it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch caching allocator); the
 *     library never allocates, frees or retains buffers;
 *   - activations are bf16 NHWC ("pixel-major"), statistics / losses / gradients of parameters
 *     are fp32; images at the module surface are fp32 NCHW like the reference's;
 *   - every call is asynchronous on the cudaStream_t passed as `stream` (void*), never
 *     synchronises the device, and is safe under CUDA-graph capture;
 *   - return value: 0 = ok, negative = error (message: msig_last_error()). Unsupported shapes
 *     fail loudly; there is no CPU or library fallback.
 */
#ifndef MSIG_H_
#define MSIG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSIG_VERSION 100

enum { MSIG_OK = 0, MSIG_ERR_ARG = -1, MSIG_ERR_CUDA = -2, MSIG_ERR_UNSUPPORTED = -3 };

enum { MSIG_ACT_NONE = 0, MSIG_ACT_RELU = 1, MSIG_ACT_LRELU = 2, MSIG_ACT_TANH = 3 };
enum { MSIG_AUX_NONE = 0, MSIG_AUX_ADD = 1, MSIG_AUX_RELU_MASK = 2, MSIG_AUX_LRELU_MASK = 3 };
enum { MSIG_OUT_BF16_NHWC = 0, MSIG_OUT_F32_NCHW = 1, MSIG_OUT_F32_NHWC = 2 };

/* ---- context ------------------------------------------------------------------------ */
int msig_init(int device);               /* binds the device, resolves cuTensorMapEncodeTiled */
int msig_version(void);
const char* msig_last_error(void);       /* thread-local */
int msig_sm_count(void);
int msig_debug_set_ring_mode(int mode);   /* test hook: strip-ring kernel for 64-channel layers: bit 0 on, bit 1 four convT phases in one launch, bit 2 legacy per-output-row MMA order, bits 8..15 ring-depth cap; default 3 */
int msig_debug_set_wgrad_mode(int mask);  /* test hook: bit 0 = M-stacked row-patch weight gradients, bit 1 = tap-grouped convT ones; default 3 */
int msig_debug_set_m2_mode(int on);       /* test hook: two m-tiles per CTA for the 128-wide conv tiles, default on */
int msig_debug_set_pdl(int on);           /* test hook: programmatic dependent launch of every kernel (csrc/common.h), default OFF (measured: no gain);
                                           * the environment variable MSIG_PDL=1 turns it on */
int msig_debug_set_pair_mode(int on);     /* test hook: CTA-pair (tcgen05 cta_group::2) kernel for 256-wide conv tiles, default on */
long long msig_kernel_launches(void);    /* kernels launched by this library since load */

/* ---- convolution geometry -------------------------------------------------------------
 * x: [n,h,w,c] bf16 NHWC -> y: [n,oh,ow,k]. stride 1 or 2; pad_t/pad_l are the top/left zero
 * padding (bottom/right follow from oh/ow; asymmetric padding as in ZeroPad2d((1,0,1,0)) +
 * padding=1, model.py:182-183, is pad_t=pad_l=2). c must be a multiple of 64. */
typedef struct msig_conv_geom {
  int32_t n, h, w, c;
  int32_t k;
  int32_t r, s, stride, pad_t, pad_l;
  int32_t oh, ow;
} msig_conv_geom;

/* Fused epilogue of every GEMM-shaped op:  y = act( aux_op( alpha*acc + bias ) ). */
typedef struct msig_epilogue {
  const float* bias;      /* [k] fp32 or NULL */
  const void* aux;        /* bf16 tensor shaped like the bf16 NHWC output, or NULL */
  int32_t aux_mode;       /* MSIG_AUX_* : residual add, or multiply by act'(aux) (ReLU / LeakyReLU) */
  int32_t act;            /* MSIG_ACT_* */
  float alpha;            /* 1.0f for plain convs */
  const float* alpha_ptr; /* optional DEVICE scalar multiplied into alpha (upstream loss gradient) */
  float slope;            /* LeakyReLU slope */
  int32_t out_layout;     /* MSIG_OUT_* */
  /* Optional fused per-(image, channel) reductions of the stored output v (bf16 NHWC outputs,
   * k >= 64): stats_partial receives [rows][2][ld] fp32 partial sums, rows = 4 per 128-pixel output
   * tile (msig_epilogue_stats_rows), ld = k rounded up to 64; q=0: sum v, q=1: sum v*v, or sum v*z
   * when stats_z (bf16, shaped like the output) is given. They replace the separate statistics pass
   * of InstanceNorm / AdaIN (model.py:16) and the two reductions of its backward; finish them with
   * msig_in_stats_from_partials / msig_norm_bwd_from_partials. */
  float* stats_partial;
  const void* stats_z;
  const float* ch_scale;  /* optional [k] fp32 per-output-channel multiplier, applied with alpha before the bias
                           * (msig_conv_narrow_fwd only: the VGG input renormalisation's 0.5/std on d(image)) */
  /* Optional (both or neither; fp32 [n][k], 16-byte aligned; needs stats_z, k % 16 == 0 and aux_mode RELU_MASK /
   * LRELU_MASK): the activation mask is act'(stats_z * mask_scale + mask_shift), i.e. recomputed from the norm
   * INPUT z and the scale / shift its InstanceNorm / AdaIN applied (model.py:16, 28-36), instead of being read
   * from the saved activation: `aux` is ignored and one 2-byte-per-element stream leaves the epilogue. */
  const float* mask_scale;
  const float* mask_shift;
  /* 0: stats_partial has the per-tile rows of msig_epilogue_stats_rows. > 0 (= msig_ring_stats_rows of this layer):
   * the strip-ring kernel writes ONE partial row per (work item, phase, accumulator quadrant) -- its lean epilogue
   * keeps a pixel column's sums in registers over the item's rows (plain statistics only: no stats_z, no aux);
   * any other kernel refuses the call. */
  int32_t stats_rows;
} msig_epilogue;
/* rows of stats_partial PER IMAGE in the ring kernel's per-item layout (msig_epilogue.stats_rows); 0 = this layer does
 * not run on the ring kernel. kind 0: msig_conv_rowpatch_fwd, 1: msig_convT2d_fwd, 2: msig_conv2d_fwd (64 -> 64, s1) */
int32_t msig_ring_stats_rows(int32_t kind, const msig_conv_geom* g);
/* rows of stats_partial PER IMAGE for an output plane oh x ow produced in `phases` (1, or 4 for the
 * k4 s2 transposed conv / stride-2 dgrad, where oh x ow is the per-phase plane = the INPUT plane). */
int32_t msig_epilogue_stats_rows(int32_t oh, int32_t ow, int32_t phases);

/* ---- weight packing --------------------------------------------------------------------
 * Master weights stay fp32 in the reference's state_dict layout (OIHW; ConvTranspose2d IOHW;
 * Linear [out,in]). Packed bf16 copies are derived caches, refreshed after each optimizer step. */
enum {
  MSIG_WPACK_FWD = 0,         /* OIHW -> [Opad][r*s][I]            conv fwd / Linear fwd (r=s=1)      */
  MSIG_WPACK_DGRAD_S1 = 1,    /* OIHW -> [Ipad][flip(r*s)][O]      dgrad of stride-1 conv / Linear    */
  MSIG_WPACK_DGRAD_S2 = 2,    /* OIHW(4x4,s2,p1) -> 4 x [Ipad][2*2][O]  dgrad as 4 output phases      */
  MSIG_WPACK_CONVT_FWD = 3,   /* IOHW(4x4,s2,p1) -> 4 x [Opad][2*2][I]  transposed conv, 4 phases     */
  MSIG_WPACK_CONVT_DGRAD = 4, /* IOHW -> [Ipad][4*4][O]            dgrad of transposed conv (s2 conv) */
  MSIG_WPACK_IM2COL = 5,      /* OIHW (small I) -> [Opad][Kpad], k=(r*s)*I+i   gathered-patch GEMM    */
  MSIG_WPACK_IM2COL_DGRAD = 6,/* OIHW (small I) -> [Kpad][O]                   its dgrad              */
  MSIG_WPACK_IM2COL_FLIP = 7, /* OIHW (small O) -> [Ipad][Kpad], k=flip(r*s)*O+o  dgrad of a small-O conv via gathered dy */
  MSIG_WPACK_ROWPATCH = 10,   /* OIHW (I <= 8) -> [Opad][r][s*8+i]                msig_conv_rowpatch_fwd          */
  MSIG_WPACK_ROWPATCH_FLIP = 11,/* OIHW (O <= 8) -> [Ipad][R-1-r][(S-1-s)*8+o]     dgrad of a small-O conv as a
                                 msig_conv_rowpatch_fwd over the pad8 output gradient (pad = R-1)              */
  MSIG_WPACK_ROWFOLD = 8,     /* OIHW (O <= 4, I = 64) -> [r][s*4+o][I]            msig_conv_narrow_fwd            */
  MSIG_WPACK_ROWFOLD_DGRAD = 9,/* OIHW (I <= 4, O = 64) -> [R-1-r][(S-1-s)*4+i][O]  dgrad (w.r.t. the image) of a
                                 small-I conv, run as msig_conv_narrow_fwd over dy with pad = R-1-pad          */
};
typedef struct msig_wpack_desc {
  int32_t kind;
  int32_t o, i, r, s;     /* dims of the fp32 master weight as the reference stores it */
} msig_wpack_desc;
size_t msig_wpack_elems(const msig_wpack_desc* d);   /* bf16 elements of the packed buffer */
int msig_wpack(const msig_wpack_desc* d, const float* w, void* packed, void* stream);

/* All packs of one network in one launch (they are refreshed after every optimizer step): describe them
 * once as jobs, build the table (host memory, msig_wpack_table_bytes), copy it to the device, then call
 * msig_wpack_multi each step. d.kind < 0 marks a plain fp32 copy of copy_numel elements (bias tables). */
typedef struct msig_wpack_job {
  msig_wpack_desc d;
  int32_t oc, o_off;      /* composite packs, as in msig_wpack_part */
  const float* src;       /* fp32 master weight (device) */
  void* dst;              /* packed bf16 buffer (device); fp32 for copies */
  int64_t copy_numel;
} msig_wpack_job;
size_t msig_wpack_table_bytes(int32_t n_jobs);
/* total_out: elements handled by the generic (scatter) kernel; total_tiles_out: shared-memory tiles of the
 * big FWD / DGRAD_S1 packs (coalesced reads AND writes). Pass both to msig_wpack_multi. */
int msig_wpack_table_build(const msig_wpack_job* jobs, int32_t n_jobs, void* table_host, int64_t* total_out,
                           int64_t* total_tiles_out);
int msig_wpack_multi(const void* table_dev, int32_t n_jobs, int64_t total, int64_t total_tiles, void* stream);

/* ---- convolutions (tcgen05 implicit GEMM) ------------------------------------------------
 * Replace nn.Conv2d (model.py:45,48,72-75,132-133,165,183; losses.py:15), nn.Linear
 * (model.py:18; r=s=1 on a [1,1,M,K] view) and the 1x1 heads (model.py:84). */
int msig_conv2d_fwd(const msig_conv_geom* g, const void* x, const void* w_fwd,
                    const msig_epilogue* e, void* y, void* stream);
/* ---- row-patch convolutions of few-channel images (model.py:131 forward + weight gradient; model.py:141
 * input gradient + weight gradient): no gathered patch matrix. msig_img_pad8 stores the fp32 NCHW image
 * (c <= 8) reflect- or zero-padded as bf16 [n][h+2p][w+2p+2][8]; a TMA map with a one-pixel (16-byte) W
 * stride then reads the (s, c) window of filter row r for every output pixel as one K-major 128-byte
 * row. Geometry: n,h,w,c = the UNPADDED image, pad_t = pad_l = p, stride 1, s <= 8.
 * dst_pad8 must have room for MSIG_PAD8_SLACK_PIXELS (8) extra pixels (128 bytes) behind the last row;
 * msig_img_pad8 zeroes them: the 8-pixel window of the last output columns of the LAST padded row reaches up to
 * 7 - s pixels past the row (in every other row that is the start of the next row, met by zero weights). */
#define MSIG_PAD8_SLACK_PIXELS 8
int msig_img_pad8(const float* src_nchw, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                  int32_t reflect, const float* scale /* [c] or NULL */, const float* shift /* [c] or NULL */,
                  void* dst_pad8, void* stream);   /* stores x*scale + shift (VGG renorm, losses.py:49-56); padding = 0 */
int msig_conv_rowpatch_fwd(const msig_conv_geom* g, const void* x_pad8, const void* w_rowpatch,
                           const msig_epilogue* e, void* y, void* stream);
size_t msig_conv_rowpatch_wgrad_workspace(const msig_conv_geom* g);
/* flip=0: dw[k][c][r][s] (+)= wgrad(pad8 image, dy [n,oh,ow,64]); flip=1: pad8 holds the few-channel output
 * gradient (pad = R-1) and `other` the 64-channel input [n,oh,ow,64] of a small-O conv (dw is [c][64][r][s]). */
int msig_conv_rowpatch_wgrad(const msig_conv_geom* g, const void* x_pad8, const void* other, int flip,
                             float* dw, int accumulate, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Stride-1 convolution with k <= 4 output channels of a 64-channel input (model.py:141, the final 7x7
 * 64->3 conv; and, with MSIG_WPACK_ROWFOLD_DGRAD weights, the image gradient of the first 7x7 3->64 conv,
 * model.py:131): the horizontal taps are folded into the GEMM N dimension and consecutive output rows
 * share their input strips in shared memory. y is fp32 (NCHW or NHWC); epilogue: alpha, bias, act. */
int msig_conv_narrow_fwd(const msig_conv_geom* g, const void* x, const void* w_rowfold,
                         const msig_epilogue* e, void* y, void* stream);
/* dx = dgrad(dy): w_dgrad is MSIG_WPACK_DGRAD_S1 (stride 1) or MSIG_WPACK_DGRAD_S2 (stride 2). */
int msig_conv2d_dgrad(const msig_conv_geom* g, const void* dy, const void* w_dgrad,
                      const msig_epilogue* e, void* dx, void* stream);
/* dw (fp32, reference layout OIHW) (+)= wgrad(x, dy). workspace holds split-K partials. */
size_t msig_conv2d_wgrad_workspace(const msig_conv_geom* g);
int msig_conv2d_wgrad(const msig_conv_geom* g, const void* x, const void* dy, float* dw,
                      int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ConvTranspose2d(k=4, s=2, p=1) (model.py:139-140): x [n,h,w,c] -> y [n,2h,2w,k];
 * geometry fields: n,h,w,c = input, k = output channels, oh=2h, ow=2w, r=s=4, stride=2. */
int msig_convT2d_fwd(const msig_conv_geom* g, const void* x, const void* w_convt_fwd,
                     const msig_epilogue* e, void* y, void* stream);
int msig_convT2d_dgrad(const msig_conv_geom* g, const void* dy, const void* w_convt_dgrad,
                       const msig_epilogue* e, void* dx, void* stream);
size_t msig_convT2d_wgrad_workspace(const msig_conv_geom* g);
int msig_convT2d_wgrad(const msig_conv_geom* g, const void* x, const void* dy, float* dw,
                       int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- gathered-patch ("im2col") path for the narrow convs (3-channel images, few-channel
 * gradients): model.py:131 (7x7 reflect, 3->64), :72/:165 (4x4 s2, 3->64), :141 (64->3 backward),
 * :183 (head backward), losses.py conv_1_1. Source is fp32 NCHW [n,c,h,w]; patches go to a bf16
 * [n*oh*ow][kpad] matrix with k = (r*S+s)*c + ch, kpad = roundup(r*s*c, 64); then the GEMM runs
 * through msig_conv2d_fwd / msig_conv2d_wgrad with r=s=1. scale/shift (per channel, may be NULL)
 * fold the VGG input renormalisation (losses.py:49-56) into the gather. */
typedef struct msig_patch_geom {
  int32_t n, c, h, w;
  int32_t r, s, stride, pad_t, pad_l;
  int32_t oh, ow;
  int32_t reflect;        /* 1: reflect padding (padding_mode='reflect'), 0: zero padding */
  int32_t kpad;
} msig_patch_geom;
int msig_patch_gather(const msig_patch_geom* g, const float* src_nchw, const float* scale,
                      const float* shift, void* patches, void* stream);
/* Adjoint: dsrc[n,c,h,w] (fp32, (+)=) = scale[c] * sum of the patch-gradient entries that read it. */
int msig_patch_scatter(const msig_patch_geom* g, const void* dpatches, const float* scale,
                       float* dsrc_nchw, int accumulate, void* stream);
/* wgrad for a gathered-patch GEMM: dw (OIHW fp32, o x c x r x s) (+)= dy^T * patches.
 * `flip`=1 is the MSIG_WPACK_IM2COL_FLIP arrangement (patches gathered from dy, `other` = input). */
size_t msig_patch_wgrad_workspace(int64_t rows, int32_t m, int32_t ncols);
/* Two-step form for weights stored as several master tensors (per-domain heads model.py:84,183;
 * the 16 AdaIN Linears model.py:18): one GEMM  partial[split][m][ncols] = a^T b  into the
 * workspace, then one msig_wgrad_unpack per master tensor. */
int msig_gemm_tn_partial(int64_t rows, const void* a_rows_m, int32_t m, const void* b_rows_n, int32_t ncols,
                         void* workspace, size_t workspace_bytes, int32_t* splits_out, void* stream);
/* ... or ONE launch for `layers` equally shaped Linear layers [out_features][in_features] (+ bias) whose
 * outputs are consecutive column blocks of the batched GEMM: dW_l += sum of the split partials, db_l += column
 * sums of the fp32 output gradient dy [rows][ld]; wgrad_ptrs / bgrad_ptrs are DEVICE arrays of `layers`
 * fp32 pointers (the layers' own gradient tensors). */
int msig_multi_linear_grads(const float* partial, int32_t splits, int64_t split_stride, const float* dy,
                            int64_t rows, int64_t ld, int32_t layers, int32_t out_features, int32_t in_features,
                            const void* wgrad_ptrs, const void* bgrad_ptrs, void* stream);
int msig_wgrad_unpack(const msig_wpack_desc* d, int32_t oc, int32_t o_off, const float* partial,
                      int32_t splits, int64_t split_stride, float* dw, int accumulate, void* stream);
size_t msig_wpack_part_elems(const msig_wpack_desc* d, int32_t oc);
int msig_wpack_part(const msig_wpack_desc* d, int32_t oc, int32_t o_off, const float* w, void* packed,
                    void* stream);
int msig_patch_wgrad_part(const msig_wpack_desc* d, int32_t oc, int32_t o_off, int64_t rows,
                          const void* a_rows_m, int32_t m, const void* b_rows_n, int32_t ncols, float* dw,
                          int accumulate, void* workspace, size_t workspace_bytes, void* stream);
int msig_patch_wgrad(const msig_wpack_desc* d, int64_t rows, const void* a_rows_m, const void* b_rows_n,
                     float* dw, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* ---- reflect padding of NHWC bf16 (7x7 reflect conv on 64 channels, model.py:141) -------- */
/* Adjoint of a reflect padding for an fp32 NCHW gradient (image side of the reflect-padded first conv, model.py:131):
 * dx[n,c,h,w] = sum of dy_padded[n,c,.,.] over the padded positions that mirror onto (h,w). */
int msig_reflect_fold_nchw(const float* dy_padded, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                           float* dx, void* stream);

/* ---- InstanceNorm / AdaIN (model.py:16,20-36,131-133,139-140,167) -----------------------
 * x: [n,hw,c] bf16. Statistics are biased, eps inside the sqrt, fp32.
 * gamma/beta: fp32 with row stride `gb_stride` (NULL = plain InstanceNorm, gamma=1, beta=0);
 * a row stride of 0 broadcasts one style code over the batch. */
size_t msig_in_stats_workspace(int32_t n, int32_t hw, int32_t c);
int msig_in_stats(const void* x, int32_t n, int32_t hw, int32_t c, float eps,
                  const float* gamma, const float* beta, int64_t gb_stride,
                  float* mean, float* rstd, float* scale, float* shift,
                  void* workspace, size_t workspace_bytes, void* stream);
/* y = act(x*scale + shift) (+ residual) */
int msig_norm_act_fwd(const void* x, const float* scale, const float* shift, const void* residual,
                      int32_t act, float slope, int32_t n, int32_t hw, int32_t c, void* y,
                      void* stream);
/* dx from dy (grad of y); also dgamma/dbeta [n,c] with row stride gb_stride (may be NULL). */
int msig_norm_act_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                      const float* scale, const float* shift, const float* gamma, int64_t gb_stride,
                      int32_t act, float slope, int32_t n, int32_t hw, int32_t c, void* dx,
                      float* dgamma, float* dbeta, int64_t dgb_stride, int accumulate_dgb,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The same two kernels with the reflect padding of the generator's last conv (model.py:141) fused in: the
 * forward writes act(x*scale+shift) straight into the reflect-padded buffer [n][h+2p][w+2p][c] (interior
 * plus mirrored border copies); the backward reads dy through the fold of the padded gradient
 * [n][h+2p][w+2p][c]: no separate pad / fold pass exists. */
int msig_norm_act_fwd_pad(const void* x, const float* scale, const float* shift, int32_t act, float slope,
                          int32_t n, int32_t h, int32_t w, int32_t c, int32_t pad, void* y_padded,
                          void* stream);
size_t msig_norm_act_bwd_pad_workspace(int32_t n, int32_t h, int32_t w, int32_t c);
int msig_norm_act_bwd_pad(const void* dy_padded, const void* x, const float* mean, const float* rstd,
                          const float* scale, const float* shift, int32_t act, float slope, int32_t n,
                          int32_t h, int32_t w, int32_t c, int32_t pad, void* dx, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Finish epilogue-fused reductions (msig_epilogue.stats_partial): same results as msig_in_stats /
 * msig_norm_act_bwd without re-reading the activation for the reduction. For the backward form the
 * epilogue must already have applied act' to dy (aux mask), i.e. `g` is dy*act'(u), and stats_z = x. */
int msig_in_stats_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld,
                                int32_t hw, int32_t c, float eps, const float* gamma, const float* beta,
                                int64_t gb_stride, float* mean, float* rstd, float* scale, float* shift,
                                void* stream);
int msig_norm_bwd_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld,
                                const void* g, const void* x, const float* mean, const float* rstd,
                                const float* scale, const float* shift, int32_t hw, int32_t c, void* dx,
                                float* dgamma, float* dbeta, int64_t dgb_stride, int accumulate_dgb,
                                float* coef_workspace /* [n][2][c] fp32 */, void* stream);

/* Same results in ONE launch each: every block of the apply kernel folds the partial rows of its image in its
 * prologue (meant for few rows per image, e.g. the 32 rows of a 64x64 plane; the caller chooses), block 0 of
 * the image writes mean / rstd / scale / shift (forward) or coef / dgamma / dbeta (backward). Replaces
 * msig_in_stats_from_partials + msig_norm_act_fwd, and msig_norm_bwd_from_partials (model.py:16,28-36,51-55). */
int msig_norm_act_fwd_from_partials(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld,
                                    int32_t hw, int32_t c, float eps, const float* gamma, const float* beta,
                                    int64_t gb_stride, float* mean, float* rstd, float* scale, float* shift,
                                    const void* x, const void* residual, int32_t act, float slope, void* y,
                                    void* stream);
int msig_norm_bwd_from_partials_fused(const float* partial, int32_t n, int32_t rows_per_img, int32_t ld,
                                      const void* g, const void* x, const float* mean, const float* rstd,
                                      const float* scale, const float* shift, int32_t hw, int32_t c, void* dx,
                                      float* dgamma, float* dbeta, int64_t dgb_stride, int accumulate_dgb,
                                      float* coef_workspace /* [n][2][c] fp32 */, void* stream);

/* ---- small bandwidth ops ------------------------------------------------------------------ */
/* dz = dy * act'(y) (ReLU / LeakyReLU), bf16 */
int msig_act_bwd(const void* dy, const void* y, int32_t act, float slope, int64_t numel, void* dz,
                 void* stream);
/* bias gradient: db[c] (+)= sum over rows of dy[rows][c] (bf16 rows, fp32 out). Deterministic two-stage
 * sum (per-block partials in the workspace, folded in block order by the last block; no fp32 atomics). */
size_t msig_colsum_workspace(int64_t rows, int32_t c);
int msig_colsum(const void* dy, int64_t rows, int32_t c, float* db, int accumulate, void* workspace,
                size_t workspace_bytes, void* stream);
/* MaxPool2d(2) on NHWC bf16 (losses.py pool_2 / pool_4) */
int msig_maxpool2_fwd(const void* x, int32_t n, int32_t h, int32_t w, int32_t c, void* y, void* stream);
int msig_maxpool2_bwd(const void* dy, const void* x, const void* y, int32_t n, int32_t h, int32_t w,
                      int32_t c, void* dx, void* stream);
/* AdaptiveAvgPool2d(1) (model.py:76): [n,hw,c] bf16 -> [n,c] bf16, and its backward */
int msig_avgpool_fwd(const void* x, int32_t n, int32_t hw, int32_t c, void* y, void* stream);
int msig_avgpool_bwd(const void* dy, int32_t n, int32_t hw, int32_t c, void* dx, void* stream);
/* per-sample head selection (model.py:112-116, 208-212): out[b, :] = all[b, idx[b], :] over a
 * [n, heads*per_head] fp32 row (SE) or [n, pix, heads_pad] (D); the backward writes zeros to the
 * unselected heads. idx is int64 like the reference's domain_idx; `heads` = the number of real heads
 * (<= heads_ld, the padded stride): negative indices wrap like torch's, an index outside [-heads, heads)
 * never reads out of bounds -- the selected output is NaN and the backward routes nothing. */
int msig_head_gather(const float* all, const int64_t* idx, int32_t n, int32_t pix, int32_t heads_ld,
                     int32_t heads, int32_t per_head, int32_t head_major, float* out, void* stream);
int msig_head_scatter(const float* dout, const int64_t* idx, int32_t n, int32_t pix, int32_t heads_ld,
                      int32_t heads, int32_t per_head, int32_t head_major, float* dall, void* stream);
/* dtype / layout converters at the module surface */
int msig_f32_to_bf16(const float* x, int64_t numel, void* y, void* stream);
int msig_bf16_to_f32(const void* x, int64_t numel, float* y, void* stream);
int msig_tanh_bwd(const float* dy, const float* y, int64_t numel, float* dz, void* stream);
/* sum over (n, h, w) of an fp32 NCHW tensor per channel (bias grad of the final conv) */
size_t msig_nchw_chansum_workspace(int32_t n, int32_t c, int64_t hw);
int msig_nchw_chansum(const float* x, int32_t n, int32_t c, int64_t hw, int64_t img_stride, float* out,
                      int accumulate, void* workspace, size_t workspace_bytes,
                      void* stream);   /* img_stride: elements between images (c*hw if dense); deterministic */

/* ---- losses (trainer.py:50-52, losses.py:70-98) ---------------------------------------------
 * Forward kernels write the mean-reduced loss (fp32 device scalar). Backward kernels write
 * d(loss)/da * (*gscale), where gscale is the upstream gradient as a DEVICE scalar, so no host
 * synchronisation is needed anywhere in a training step. The reductions are deterministic (per-block
 * partial sums in `workspace`, folded in a fixed order by the last block; no floating-point atomics):
 * msig_reduce_workspace() bytes cover every 1-D loss kernel and msig_sumsq. */
size_t msig_reduce_workspace(void);
int msig_l1_loss_f32_fwd(const float* a, const float* b, int64_t numel, float* loss, void* workspace,
                         size_t workspace_bytes, void* stream);
int msig_l1_loss_f32_bwd(const float* a, const float* b, int64_t numel, const float* gscale,
                         float* grad_a, void* stream);                   /* nn.L1Loss on images */
int msig_l1_loss_bf16_fwd(const void* a, const void* b, int64_t numel, float* loss, void* workspace,
                          size_t workspace_bytes, void* stream);
int msig_l1_loss_bf16_bwd(const void* a, const void* b, int64_t numel, const float* gscale,
                          const void* aux, void* grad_a, void* stream);  /* F.l1_loss on VGG features; (+ aux) */
int msig_mse_const_fwd(const float* a, float target, int64_t numel, float* loss, void* workspace,
                       size_t workspace_bytes, void* stream);
int msig_mse_const_bwd(const float* a, float target, int64_t numel, const float* gscale,
                       float* grad_a, void* stream);                     /* nn.MSELoss vs ones / zeros */
/* nn.MSELoss against a target TENSOR (trainer.py:84-86 builds `valid` / `fake` tensors): no host read of
 * the target is needed when a caller keeps the reference's calling convention. */
int msig_mse_loss_fwd(const float* a, const float* target, int64_t numel, float* loss, void* workspace,
                      size_t workspace_bytes, void* stream);
int msig_mse_loss_bwd(const float* a, const float* target, int64_t numel, const float* gscale,
                      float* grad_a, void* stream);
/* Gram matrix (losses.py:70-78): f [n,h,w,c] bf16 -> G [n*c][n*c] fp32 = F F^T / (n*c*h*w), batch
 * folded into rows exactly like the reference's view(a*b, c*d). */
size_t msig_gram_workspace(int32_t n, int32_t h, int32_t w, int32_t c);
int msig_gram_fwd(const void* f, int32_t n, int32_t h, int32_t w, int32_t c, float* gram,
                  void* workspace, size_t workspace_bytes, void* stream);
/* loss = mean |G_a - G_b|; ssym (bf16 [dim][dim]) = sign(D) + sign(D)^T, D = G_a - G_b. */
size_t msig_gram_l1_workspace(int32_t dim);
int msig_gram_l1(const float* ga, const float* gb, int32_t dim, float* loss, int accumulate, void* ssym,
                 void* workspace, size_t workspace_bytes,
                 void* stream);   /* accumulate=1: loss += (sum over the five taps, losses.py:84-89) */
/* df = alpha * (*gscale) * ssym * F (+ aux): gradient of the style term w.r.t. the generated
 * features; alpha = 1 / (dim^2 * n*c*h*w) supplied by the caller. relu_mask = 1: df is additionally
 * multiplied by (F > 0) -- F is the output of a ReLU (losses.py:23-35), whose backward is thereby fused. */
int msig_gram_bwd(const void* f, const void* ssym, int32_t n, int32_t h, int32_t w, int32_t c,
                  float alpha, const float* gscale, const void* aux, int relu_mask, void* df, void* stream);
/* column sums of an fp32 [rows][c] matrix (bias gradients of the fp32 heads / style Linear) */
int msig_colsum_f32(const float* x, int64_t rows, int32_t c, int64_t ld, float* out, int accumulate,
                    void* stream);   /* ld: elements between rows */

/* ---- optimizer-side multi-tensor ops on flat fp32 buffers (trainer.py:127-134,152-153;
 *      utils.py:80-91): global grad-norm clip + Adam + EMA + in one pass ------------------- */
int msig_sumsq(const float* x, int64_t numel, float* out /* fp32 scalar, (+)= */, int accumulate,
               void* workspace /* msig_reduce_workspace() bytes */, size_t workspace_bytes, void* stream);
int msig_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   float* ema /* may be NULL */, int64_t numel, const float* grad_sumsq, float max_norm,
                   float grad_scale, float lr, float beta1, float beta2, float eps, int32_t step,
                   float ema_beta, void* stream);
/* Same, with the step number kept in DEVICE memory: increments *step_counter, then uses it for the
 * bias corrections, so the call has no per-step host argument and can be replayed from a CUDA graph. */
int msig_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       float* ema /* may be NULL */, int64_t numel, const float* grad_sumsq, float max_norm,
                       float grad_scale, float lr, float beta1, float beta2, float eps,
                       int32_t* step_counter, float ema_beta, void* stream);

/* ---- GPU-side training augmentation (dataset.py:16-22: RandomResizedCrop + 0/90/180/270 rotation +
 *      ToTensor + Normalize(0.5)) for GIVEN draws ---------------------------------------------------
 * src: uint8 [n][h][w][3] decoded images; boxes: int32 [n][4] = (top, left, height, width) of each crop;
 * quarter_turns: int32 [n], counter-clockwise multiples of 90 degrees (NULL = none); out: fp32
 * [n][3][size][size] in [-1, 1]. Bit-exact with Pillow's 8-bit bilinear resample (22-bit fixed-point
 * coefficients, horizontal then vertical pass, each rounded to uint8) + torchvision to_tensor / normalize.
 * workspace: msig_augment_workspace bytes (the horizontally resampled rows). Shrink factors up to 8. */
size_t msig_augment_workspace(int32_t n, int32_t h, int32_t w, int32_t size);
int msig_augment_u8(const void* src, int32_t n, int32_t h, int32_t w, const int32_t* boxes,
                    const int32_t* quarter_turns, int32_t size, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSIG_H_ */
