/* nefnet_b200 -- C ABI of the B200-native Nef-Net hot path (libnefnet_b200.so).
 *
 * The reference (WhatAShot/Electrocardio-Panorama) is pure PyTorch and has no FFI of its own; the hot
 * path it runs is codes/network/model_nefnet.py:109-218 (Model_nefnet.forward / gen_ecg),
 * codes/network/encoder/{encoder.py:28-40,resnet_1d.py:39-53,102-105}, codes/network/utils/
 * {theta_encoder.py:13-29,roi_pooling_1d.py:38-99} and codes/network/loss/losses.py:21-50, driven by
 * codes/solver/solver.py:171-235.  Each entry point below names the reference lines it replaces.
 *
 * Conventions: every function returns 0 on success, non-zero on error (nef_last_error() gives the
 * thread-local message).  All pointers are DEVICE pointers unless marked host.  The caller owns every
 * buffer, including the workspace; the library never allocates device memory and never synchronises.
 * Every launch goes to the stream passed in.  There is no CPU path: on a machine without an sm_100
 * device nef_init() fails.
 *
 * Internal activation layout ("CBL4"): a tensor of C channels (C % 4 == 0), B segments, L samples is
 *   float4 T[C/4][B * (L + 2*NEF_HALO)],  row = b * (L + 2*NEF_HALO) + NEF_HALO + l,  lane = c % 4
 * with the halo rows kept zero.  nef_ncl_to_cbl4 / nef_cbl4_to_ncl convert from/to (B, C, L).
 */
#ifndef NEFNET_B200_H
#define NEFNET_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEF_ABI_VERSION 1
#define NEF_HALO_ROWS 3
#define NEF_GUARD_ROWS_ABI 528

typedef void* nef_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int nef_version(void);
const char* nef_last_error(void);
/* Selects `device`, checks it is sm_100, opts the kernels into their shared-memory sizes. */
int nef_init(int device);
/* 0 = CUDA-core fp32 implicit GEMM, 1 = tcgen05 TF32 (default when built in).  Test hook. */
int nef_set_conv_impl(int impl);
int nef_get_conv_impl(void);
/* Split-precision terms of the decoder's first convolution (model_nefnet.py:101-103, DoubleConv conv 256 -> 128):
 * 3 = x_hi w_hi + x_lo w_hi + x_hi w_lo (default), 2 = without x_hi w_lo, 1 = plain TF32.  Measurement hook. */
int nef_set_dec1_terms(int n);
/* 1 (default) = the encoder's forward convolutions read fp16 operand copies (NefConvTerm.x_f16); 0 = TF32 operands
 * everywhere.  Measurement hook. */
int nef_set_fwd_f16(int on);
/* 1 (default) = fp16 decoder dataflow in training (post-BatchNorm activations kept as fp16 operand copies only, decoder
 * convolutions 2-4 and all decoder data / weight gradients in kind::f16 on loss-scaled fp16 gradient copies); 0 = the
 * TF32 decoder.  Measurement hook. */
int nef_set_dec_f16(int on);
/* Test hook: 1 = round nothing to TF32 (with conv impl 0 the whole path is then plain fp32 and can be
 * compared tightly with the fp32 oracle); 0 = production behaviour.  Synchronous, call between steps. */
int nef_set_exact_fp32(int on);
/* Cumulative number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t nef_launch_count(void);
/* Test hook: which kernel forms have been launched since the last reset.  out[0..7] = launches of: 0 the persistent
 * convolution kernel, 1 the one-CTA-per-tile forms, 2 the single-tile generic form, 3 the tcgen05 weight gradient,
 * 4 its CUDA-core ragged tail, 5 the fp16-operand weight gradient; out[8..] = the distinct specialised-epilogue codes the
 * persistent kernel ran with (64 = the generic epilogue), -1 padded.  Returns the number of distinct codes.  A parity test
 * uses it to prove that it exercised the dispatch the benchmark times. */
int nef_tc_dispatch_stats(int64_t* out_host, int n);
void nef_tc_dispatch_reset(void);
/* Test hook: number of (256-row tile, group) work items from which nef_gconv_fwd takes the persistent kernel
 * (0 = default, twice the SM count; 1 = always), so that small shapes can run the production kernel form. */
int nef_tc_set_persist_min(int tiles);
/* sizeof() of the ABI structs as this library was compiled (0 NefConvTerm, 1 NefConvDesc, 2 NefWgradDesc,
 * 3 NefForwardArgs, 4 NefBackwardArgs): lets a foreign-language binding verify its mirror. */
size_t nef_struct_size(int which);

/* ---- parameters (state_dict contract; model_nefnet.py:67-107) -------------------- */
/* Number of state_dict entries for lead_num = G, their names (host strings), element counts.   */
int nef_param_count(int G);
const char* nef_param_name(int G, int index);
int64_t nef_param_numel(int G, int index);

/* ---- layout -------------------------------------------------------------------------------- */
/* rows of one channel chunk of a CBL4 tensor, and float count of a whole tensor incl. tail guard */
int64_t nef_cbl4_rows(int B, int L);
int64_t nef_cbl4_floats(int C, int B, int L);
int nef_ncl_to_cbl4(const float* src, float* dst, int B, int C, int L, int round_tf32, nef_stream_t s);
int nef_cbl4_to_ncl(const float* src, float* dst, int B, int C, int L, nef_stream_t s);
/* fp16 operand copies: `half8 T16[C/8][B * (L + 2*NEF_HALO)]`, 8 channels per 16-byte row, same row indexing and zero halo
 * as CBL4 (C %% 8 == 0; NEF_GUARD_ROWS_ABI readable 16-byte rows on both sides).  nef_ncl_to_h8 multiplies by `scale` and
 * saturates to the largest finite fp16. */
int nef_ncl_to_h8(const float* src, void* dst, int B, int C, int L, float scale, nef_stream_t s);
int nef_h8_to_ncl(const void* src, float* dst, int B, int C, int L, nef_stream_t s);

/* ---- grouped 1-D convolution as implicit GEMM ---------------------------------------------- */
/* replaces every nn.Conv1d / nn.ConvTranspose1d call of resnet_1d.py:21-24,39-53 and
 * model_nefnet.py:10-60,84-107 (forward and data-gradient; the latter is the same contraction with
 * flipped / transposed packed weights).                                                         */
typedef struct NefConvTerm {
  const float* x;       /* CBL4 input */
  int64_t x_cstride;    /* rows between channel chunks (= B * Lp) */
  int32_t x_c4_off;     /* first chunk read by group 0 */
  int32_t x_c4_gstride; /* chunk step between groups */
  int32_t cin_g;        /* input channels per group, multiple of 32 */
  int32_t taps;         /* 1, 3 or 7 */
  int32_t tap_off;      /* row offset of tap 0 (= -(taps/2) for a "same" convolution) */
  int32_t x_f16;        /* 1: x and w hold fp16 pairs -- 8 channels per 16-byte row (chunk = 8 channels), cin_g counts PAIRS
                         * of channels (a 128-channel group has cin_g = 64, 16 chunks), weights packed with flag bit 2.
                         * Same 11-bit significand as TF32 at twice the MMA rate and half the operand bytes; tensor-core
                         * implementation only.                                                                   */
  const float* w;       /* packed weights [group][tap][cin_g/32][8][N][4], see nef_pack_weights */
} NefConvTerm;

typedef struct NefConvDesc {
  int32_t n_terms;      /* 1..3: the terms accumulate into the same output */
  int32_t groups;
  int32_t N;            /* output channels per group: 64 or 128 */
  int32_t round_tf32;   /* round the stored result to TF32 (it feeds another tensor-core conv) */
  NefConvTerm term[3];
  int64_t rows;         /* B * Lp rows of the input row space are visited */
  int32_t Lp, L;        /* segment pitch and length of the input row space */
  /* output mapping: input row (b, l) -> output row b * y_Lp + HALO + l * y_lmul + y_ladd */
  float* y;
  int64_t y_cstride;
  int32_t y_c4_off, y_c4_gstride;
  int32_t y_Lp, y_lmul, y_ladd;
  /* epilogue, in this order: v = acc + bias + res ; stats(v) ; relu ; dropout ; bscale ;
   * bscale_grad ; mask ; round ; store                                                        */
  int32_t relu;
  const float* bias;    /* [groups * N] or NULL */
  const float* res;     /* CBL4, output row space, or NULL */
  int64_t res_cstride;
  int32_t res_c4_off, res_c4_gstride;
  float drop_p;         /* 0 = no dropout; survivors are scaled by 1/(1-p) */
  int32_t mask_mode;    /* 0 none, 1: v *= (mask > 0) * mask_scale, 2: v *= (mask != 0) * mask_scale */
  uint64_t drop_seed;
  const float* bscale;  /* [B][groups * N] per-segment channel scale (angular encoding), or NULL */
  float* bscale_grad;   /* [B][groups * N] += sum_l v * mask / bscale, then v *= bscale ; needs mask_mode 2 */
  const float* mask;    /* CBL4, output row space */
  int64_t mask_cstride;
  int32_t mask_c4_off, mask_c4_gstride;
  float mask_scale;
  int32_t reserved2;
  /* BatchNorm batch statistics, or NULL: record t = the sums over the valid rows of the 128-row tile t of
   * the input row space, [ceil(rows / 128)][groups * N] floats each, every record overwritten.  They are
   * reduced in a fixed order (nef_plan's finalize kernel), so equal inputs give bit-equal statistics. */
  float* stat_sum;      /* sum of v */
  float* stat_sq;       /* sum of v * v */
  /* One-bit activation masks (tensor-core implementation; the CUDA-core cross-check keeps using `mask`):
   * word [(chunk / 8)][row], bit 4 i + j <-> channel 32 (chunk / 8) + 4 i + j; planes are y_cstride (mask_cstride) words
   * apart and rows index like the float tensors.  out_bits: the epilogue also records (stored value != 0) of every
   * output element; mask_bits: mask modes 1 / 2 read these bits (plane (mask_c4_off + g * mask_c4_gstride) / 8 + ...)
   * instead of the float `mask` tensor -- 4 bytes instead of 128 per row and 32 channels.  Chunk offsets must be
   * multiples of 8.                                                                                          */
  uint32_t* out_bits;
  const uint32_t* mask_bits;
  /* fp16 copy of the output, or NULL: 8 channels per 16-byte row, chunk (y chunk / 2), planes y_cstride rows apart,
   * values saturate to the largest finite fp16.  It is the x_f16 operand of the next convolution.          */
  void* y16;
  /* Loss-scaled fp16 gradient copies (backward pass; device scalars, NULL = 1): the accumulators are multiplied by
   * acc_scale[0] before anything else (fp16 gradient operands carry the scale S, acc_scale = 1 / S), and the fp16 copy
   * stores fp16(value * y16_scale[0]) (= S).  y may be NULL when y16 is set: only the fp16 copy is kept.        */
  const float* acc_scale;
  const float* y16_scale;
  /* Residual operand from an fp16 copy (same layout as y16, geometry = the res_* fields, chunk offsets even) instead of
   * the fp32 tensor `res`: v += res16 * res16_scale[0] (device scalar, NULL = 1).  TF32-rounded activations and the
   * loss-scaled gradient copies hold the same 11-bit significands as their fp32 originals, at half the bytes.  */
  const void* res16;
  const float* res16_scale;
} NefConvDesc;

/* Packs reference-layout weights into the layout NefConvTerm.w expects, rounding to TF32 (RN):
 *   dst[g][t][kb][c][n][j] = src[g*sg + n*sn + (kb*32 + c*4 + j)*sk + ((flags & 1) ? taps-1-t : t)*st]
 *   flags bit 0: flip the taps (data gradient); bit 1: store the TF32 residual w - tf32(w) instead
 *   of tf32(w) (the low part of a split-precision contraction, used for the decoder's first conv);
 *   bit 2: fp16 operand packing for NefConvTerm.x_f16 -- dst[g][t][K/64][8][n][8 halves], a 16-byte slot = 8 consecutive
 *   input channels of one output channel, RN-even from the fp32 weight, saturating (K %% 64 == 0)      */
int nef_pack_weights(const float* src, float* dst, int groups, int N, int K, int taps, int64_t sg, int64_t sn,
                     int64_t sk, int64_t st, int flags, nef_stream_t s);
int nef_gconv_fwd(const NefConvDesc* d, nef_stream_t s);

/* weight gradient (accumulated atomically): dw[g*sg + m*sm + n*sn + t*st] += sum_rows dy[row][g, m] * x[row + t + tap_off][g, n];
 * db[g*cout_g + m] += sum_rows dy[row][g, m]                                                     */
typedef struct NefWgradDesc {
  const float* dy;
  int64_t dy_cstride;
  int32_t dy_c4_off, dy_c4_gstride;
  const float* x;
  int64_t x_cstride;
  int32_t x_c4_off, x_c4_gstride;
  int32_t cout_g, cin_g; /* multiples of 64 */
  int32_t groups, taps, tap_off;
  int32_t wg_mod; /* 0: one weight slot per group; m > 0: group g accumulates into slot g mod m of dw / db (weights shared by
                   * groups: Model_nefnet2's single-lead trunk applied to every lead) */
  int64_t rows;
  float* dw;
  int64_t sg, sm, sn, st;
  float* db; /* or NULL */
} NefWgradDesc;
int nef_gconv_wgrad(const NefWgradDesc* d, nef_stream_t s);
/* The same weight gradient from fp16 operand copies (`half8 [C/8][rows]`: 8 channels per 16-byte row, the layout of
 * NefConvDesc.y16, addressed from the same row origin as the fp32 tensors), read by the tensor core as the bulk copy lands
 * them (MN-major, no shared-memory re-tile pass).  d gives the geometry, dw and its strides (d->dy, d->x, d->db are not
 * read; chunk offsets / group strides must be even); cout_g must be 64 or 128, cin_g a multiple of 64.  The accumulated block is
 * multiplied by out_scale[0] (device scalar, NULL = 1: the inverse of the loss scale the dy16 copy carries) before the
 * fp32 accumulation into dw.  tcgen05 kind::f16: 11-bit significands like TF32, twice the rate, half the operand bytes. */
int nef_gconv_wgrad_f16(const NefWgradDesc* d, const void* dy16, const void* x16, const float* out_scale, nef_stream_t s);

/* ---- whole path ---------------------------------------------------------------------------- */
typedef struct NefPlan NefPlan;
/* V = number of extra views decoded in phase 'test' (0 for training). host call. */
int nef_plan_create(int B, int G, int L, int V, NefPlan** plan);
/* The same for a model variant: 1 = Model_nefnet (network/model_nefnet.py), 2 = Model_nefnet2 (network/model_nefnet2.py:63-203):
 * ONE single-lead trunk whose weights every lead shares (the grouped kernels read weight slice g mod m and accumulate
 * their weight gradients into it) plus single_conv_z1 / single_conv_z2, which -- being linear -- are applied to the lead
 * means and the picked leads instead of to every lead.  `params` / `grads` of nef_forward / nef_backward follow
 * nef_param_name_v(G, 2, i): the G = 1 table of variant 1 followed by single_conv_z1.0.{weight,bias},
 * single_conv_z2.0.{weight,bias}.  Phase 'gen' and nef_gen_ecg are variant-1 only. */
int nef_plan_create_v(int B, int G, int L, int V, int variant, NefPlan** plan);
int nef_param_count_v(int G, int variant);
const char* nef_param_name_v(int G, int variant, int i);
int64_t nef_param_numel_v(int G, int variant, int i);
void nef_plan_destroy(NefPlan* plan);
size_t nef_plan_workspace_bytes(const NefPlan* plan);
/* Carves the caller's workspace (must be nef_plan_workspace_bytes) and zero-fills it on `s`. */
int nef_plan_bind(NefPlan* plan, void* workspace, size_t bytes, nef_stream_t s);

enum { NEF_PHASE_TRAIN = 0, NEF_PHASE_TEST = 1, NEF_PHASE_GEN = 2 };

typedef struct NefForwardArgs {
  const float* const* params; /* host array of nef_param_count(G) device pointers, state_dict order
                                 (BatchNorm running stats included; num_batches_tracked entries are int64) */
  const float* x;             /* (B, G, L) */
  const float* input_thetas;  /* (B, G, 2) */
  const float* query_theta;   /* (B, 2) */
  const int64_t* rois;        /* (B, 7, 2) */
  const float* rest_theta;    /* (B, V, 2) or NULL */
  int32_t phase;              /* NEF_PHASE_* */
  int32_t bn_training;        /* module.training: batch statistics + running-stat update */
  int32_t lead_choice_z1, lead_choice_z2; /* the two random.randint draws, model_nefnet.py:154,156 */
  float drop_p;               /* 0.2 in training, 0 in eval */
  int32_t save_for_backward;
  uint64_t drop_seed;
  float* out;                 /* (B, 1, L) x3 ; phase GEN: out = z1 (B,128G,L/4), out_p = z2 (B,128G,7,32) */
  float* out_p;
  float* out_l;
  float* rest_out;            /* (B, V, L) */
} NefForwardArgs;
/* Model_nefnet.forward, model_nefnet.py:109-194 */
int nef_forward(NefPlan* plan, const NefForwardArgs* a, nef_stream_t s);

typedef struct NefBackwardArgs {
  const float* const* params;
  float* const* grads;        /* host array, same order; NULL entries are skipped; gradients ACCUMULATE (+=) */
  const float* dout;          /* (B, 1, L) x3, NULL = zero */
  const float* dout_p;
  const float* dout_l;
  /* Optional cudaEvent_t, recorded on the stream as soon as the gradients of every parameter from "z1_conv.0.conv1.weight"
   * to the end of the state_dict order (z1_conv, z2_conv1, z2_conv2, decoder: 2/3 of the bytes) are final -- before the
   * w_conv and encoder blocks run their backward.  A data-parallel caller all-reduces that bucket on a side stream behind
   * this event, overlapped with the rest of the backward pass.                                         */
  void* ev_late_params_done;
} NefBackwardArgs;
/* autograd of the above (solver.py:233) for the last nef_forward(save_for_backward = 1) on this plan */
int nef_backward(NefPlan* plan, const NefBackwardArgs* a, nef_stream_t s);

/* Test hook: internal activations of the last nef_forward on this plan, by name, as (B, C, L) fp32 (vectors: C floats,
 * L = 0).  Names: "stem"; "<block>.h" / "<block>.y" for the residual blocks W_encoder.layer1.{0,1,2}, w_conv.0, z1_conv.0,
 * z2_conv1.0 (centre window only), z2_conv2.0, z2_conv2.2; "roi_align"; "z2_conv2.1"; per decoder call k = 0..2:
 * "dec<k>.u0", "dec<k>.decoder.{1,3}.{0,3}" (convolution outputs before BatchNorm), "dec<k>.a1|u1|a3",
 * "dec<k>.bn<i>.scale|shift"; "stem.argmax" (the max-pool selections of the stem as floats: 0..2 = the conv position
 * 2j-1 / 2j / 2j+1 that won, 3 = clipped by the ReLU); "<block>.h.mask" / "<block>.y.mask" (encoder blocks, w_conv.0, z1_conv.0.h): the one-bit
 * (value != 0) planes the masked data gradients read, as 0 / 1 floats.  The parity tests compare them layer by layer with the oracle and evaluate the oracle's
 * backward pass on the device's own ReLU / dropout patterns. */
int nef_plan_tensor_info(const NefPlan* plan, const char* name, int* C, int* L);
int nef_plan_export(NefPlan* plan, const char* name, float* dst, nef_stream_t s);
/* Test hook: fills the fp32 storage of a named activation with NaN (a test then proves that nothing reads a tensor the
 * dataflow dropped).  scratch: B * C * L floats of device memory (shape from nef_plan_tensor_info). */
int nef_plan_poison(NefPlan* plan, const char* name, float* scratch, nef_stream_t s);

/* Model_nefnet.gen_ecg, model_nefnet.py:196-218: decode V views from supplied latents (eval-mode BN) */
int nef_gen_ecg(NefPlan* plan, const float* const* params, const float* z1, const float* z2,
                const float* query_theta /* (B, V, 2) */, const int64_t* rois, int V, float* out /* (B, V, L) */,
                nef_stream_t s);

/* ---- Standin-Learning loss, losses.py:21-50 ------------------------------------------------- */
/* ROI tiling check -- replaces the shape errors of roi_pooling_reverse (network/utils/roi_pooling_1d.py:83-98: per segment
 * the 7 reversed ROIs of lengths long(r1 * .25) - long(r0 * .25) are concatenated and the segments stacked, so a segment whose
 * lengths do not sum to L / 4, or with a negative length, makes torch.stack / torch.cat / F.interpolate raise).  The kernels
 * themselves clamp; this check lets the host raise the reference's RuntimeError.  rois int64 (B, 7, 2); flag: 3 device ints
 * written by one small launch -- [0] number of offending segments, [1] the first one (-1: none), [2] its length sum.  The
 * caller copies them back when it chooses to synchronise (Model_nefnet: at its next entry point, or at once with
 * roi_check = "sync").                                                                                               */
int nef_roi_check(const int64_t* rois, int B, int L, int32_t* flag, nef_stream_t s);

/* sums[0..2] = sum|out - out_p|, sum|out - out_l|, sum|out - target| (or squared for mse) ; doubles,
 * zeroed by the call.  losses[0..3] = total, l1*f0, l2*f1, l3*f2 as floats.                      */
int nef_loss_fwd(const float* out, const float* out_p, const float* out_l, const float* target, int64_t n,
                 int use_mse, const float* factors3_host, int using_mask, double* sums, float* losses,
                 nef_stream_t s);
/* gradients of the total w.r.t. the three predictions (out is detached in the first two terms) */
int nef_loss_bwd(const float* out, const float* out_p, const float* out_l, const float* target, int64_t n,
                 int use_mse, const float* factors3_host, int using_mask, const float* dloss /* device float[4] or NULL */,
                 float* dout, float* dout_p, float* dout_l, nef_stream_t s);
/* mean |a - b| or mean (a-b)^2 into result[0] (the unsupervised validation term, losses.py:47-49) */
int nef_pair_loss(const float* a, const float* b, int64_t n, int use_mse, double* sum, float* result,
                  nef_stream_t s);

/* ---- optimiser (solver/optim_scheduler.py:10, solver.py:234-235) ---------------------------- */
/* g *= gscale ; m = momentum * m + g ; p -= lr * m    over flat buffers (lr read from host value) */
int nef_sgd_step(float* p, const float* g, float* m, int64_t n, float lr, float momentum, float gscale,
                 nef_stream_t s);

/* ---- callers either side of the path ---------------------------------- */
/* utils/mertic.py:7-21 PSNR(pred, gt, rois): per (segment, view) row 20 log10(1 / rmse) over [0, rois[b, 6, 0]) (whole
 * row when rois is NULL), 100 when rmse == 0.  rows: scratch double[B * V] (the per-row values, kept for inspection);
 * acc: double[2] = {sum of row values, rows seen}, accumulated across calls (zero it at the start of an epoch);
 * result (optional, device float): running mean acc[0] / acc[1].  No host synchronisation.                     */
int nef_psnr(const float* pred, const float* gt, const int64_t* rois, int B, int V, int L, double* rows, double* acc,
             float* result, nef_stream_t s);
/* dataset/tianchi.py:84-111, 212-225 for a batch of whole records already in device memory.
 * raw: packed records, record b = 8 leads x rec_len[b] doubles (row-major) starting at element rec_off[b];
 * marks (B, 7) int64 = p_on, p_off, r_on, r_off, t_on, t_off, end_point of the chosen heartbeat (:96-102);
 * derives III, aVR, aVL, aVF (:88-93), crops [p_on, end_point) (:107), min-max normalises over the 12 x crop block
 * (:110-111), zero-pads / truncates to L (:212-219).  Outputs (each optional): ori (B, 12, L) = 'ori_data';
 * data (B, G, L) = leads select[b, :] ('data'); target (B, L) = lead target_index[b] ('target_view');
 * rois (B, 7, 2) int64 (:103-106, the literal 512 is L).  scratch: nef_prepare_scratch_bytes(B) bytes of device
 * memory (partial minima / maxima).                                                                              */
size_t nef_prepare_scratch_bytes(int B);
int nef_prepare_segments(const double* raw, const int64_t* rec_off, const int32_t* rec_len, const int64_t* marks,
                         int B, int L, const int32_t* select, int G, const int32_t* target_index, double* scratch,
                         float* ori, float* data, float* target, int64_t* rois, nef_stream_t s);

/* ---- single ops, exported for unit tests ---------------------------------------------------- */
/* encoder stem, resnet_1d.py:102-105 + encoder.py:35-38: x (B,G,L) -> CBL4 (128G, L/4) */
/* argmax: one byte per output element (a uint32 per float4 row of y, same indexing): which pooled conv position won
 * (0..2, MaxPool1d's first maximum) or 3 where the ReLU clipped; written by fwd (may be NULL), required by bwd. */
int nef_stem_fwd(const float* x, const float* w, float* y, uint32_t* argmax, int B, int G, int L, nef_stream_t s);
int nef_stem_bwd(const float* x, const uint32_t* argmax, const float* dy, float* dw, int B, int G, int L, nef_stream_t s);
/* The stem forward on the tensor cores (split-precision fp16 MMAs, fp32-accurate): writes the fp16 copy y16 (half8 rows,
 * layout of nef_ncl_to_h8 for (B, 128 G, L / 4)) and the argmax codes (may be NULL) -- the production form of the stem. */
int nef_stem_tc_fwd(const float* x, const float* w, void* y16, uint32_t* argmax, int B, int G, int L, nef_stream_t s);
/* The stem weight gradient on the tensor cores: dy16 = fp16 copy (layout of y16 above) of the gradient of the stem output,
 * multiplied by a loss scale S; inv_scale = device pointer to 1 / S (NULL = 1); dw (128 G, 1, 15) accumulates (+=). */
int nef_stem_tc_bwd(const float* x, const uint32_t* argmax, const void* dy16, float* dw, const float* inv_scale, int B, int G,
                    int L, nef_stream_t s);
/* Angular encoding + Linear, theta_encoder.py:13-29 + model_nefnet.py:76-77: (n,2) -> (n,D) */
int nef_angular_fwd(const float* theta, const float* w, const float* b, float* out, int n, int D, nef_stream_t s);
int nef_angular_bwd(const float* theta, const float* dout, float* dw, float* db, int n, int D, nef_stream_t s);

#ifdef __cplusplus
}
#endif
#endif
