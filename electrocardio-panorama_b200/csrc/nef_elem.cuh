// Host-callable launchers of the non-GEMM kernels (nef_elem.cu), used by the plan (nef_plan.cu).
#pragma once
#include "nef_common.cuh"
#include "../../include/nefnet_b200.h"

namespace nef {

// A CBL4 tensor view: p = row 0 of chunk 0, cs = rows per chunk (= B * Lp), Lp = L + 2 * HALO.
struct T4 {
  float4* p;
  long cs;
  int C, B, L, Lp;
  __host__ __device__ long row(int b, int l) const { return (long)b * Lp + NEF_HALO + l; }
  __host__ __device__ float4* at(int c4, int b, int l) const { return p + (long)c4 * cs + row(b, l); }
};

struct BnLayer {          // one BatchNorm1d of the decoder for one decoder call
  float* sum;             // [n_rec][C] per-128-row-tile sums of the conv output (NefConvDesc.stat_sum)
  float* sq;              // [n_rec][C]
  int n_rec;              // ceil(rows / 128)
  float* scale;           // [C] gamma * invstd
  float* shift;           // [C] beta - mean * scale
  float* mean;            // [C] saved for backward
  float* invstd;          // [C]
  double* s1;             // [C] backward: sum g
  double* s2;             // [C] backward: sum g * xhat
};

int stem_fwd(const float* x, const float* w, T4 y, uint32_t* amax, void* y16, int G, cudaStream_t s, int store32 = 1, int w_shared = 0);  // y16: optional fp16 copy (store32 = 0: only that copy is written)
int stem_bwd(const float* x, const uint32_t* amax, T4 dy, float* dw, int G, cudaStream_t s, int w_shared = 0);
// the same on the tensor cores (nef_stem_tc.cu): split-precision fp16 MMAs, fp32-accurate; outputs = the fp16 copy y16 and
// the codes only (geometry of y)
// weight gradient on the tensor cores from the loss-scaled fp16 copy dy16 of the stem output's gradient (geometry of dy)
int stem_tc_bwd(const float* x, const uint32_t* amax, const void* dy16, T4 dy, float* dw, const float* inv_scale, int G,
                cudaStream_t s, int w_shared = 0);
int stem_tc_fwd(const float* x, const float* w, T4 y, uint32_t* amax, void* y16, int G, cudaStream_t s, int w_shared = 0);
int angular_fwd(const float* theta, const float* w, const float* b, float* out, int n, int D, cudaStream_t s);
int angular_bwd(const float* theta, const float* dout, float* dw, float* db, int n, int D, cudaStream_t s);

// z2_conv1 only matters at the centre columns (roi_algin samples nothing else)
struct Window { int w0, Lw, y0; float wy1; };
Window centre_window(int L4);
int window_extract(T4 w, T4 xw, int G, Window win, const void* w16, cudaStream_t s);   // w16: read the fp16 copy of w instead                  // w z2-half -> (64G, Lw)
// -> z2 half of g_w, zeros elsewhere; gw16 (optional): fp16 copy of the same rows scaled by s16[0] (device scalar)
//   store32 = 0: only the fp16 copy is written
int window_scatter(T4 gxw, T4 gw, int G, Window win, void* gw16, const float* s16, int store32, cudaStream_t s);
// scale[0] = S, scale[1] = 1 / S, S = 2^k with S * max|d0, d1, d2| in [32, 64) (1 if all zero); scale[2] is scratch
int grad_loss_scale(const float* d0, const float* d1, const float* d2, long n, float* scale, cudaStream_t s);
int roi_check(const int64_t* rois, int B, int L4, int* flag, cudaStream_t s);   // flag[3]: bad segments, first one, its length sum
int roi_align_fwd(T4 z2c, const int64_t* rois, T4 ra, Window win, int L4, cudaStream_t s, void* ra16 = nullptr);   // ra16: write only the fp16 copy
int roi_align_bwd(T4 dra, const int64_t* rois, T4 z2c, T4 gz2c, Window win, int L4, cudaStream_t s);
int deinterleave2(T4 src, T4 even, T4 odd, cudaStream_t s);
int deinterleave2_h(const void* src16, T4 src, void* even16, void* odd16, T4 even, cudaStream_t s);   // fp16 copies (half8 rows)
// gradient of the angular scale s (model_nefnet.py:120-123): ys = relu(u) * s[b, c]; gx = d ys * s * (ys != 0)  ->
//   ds[b, c] = sum_l d ys * relu(u) = sum_l gx * ys / s^2      (overwrites ds)
int bscale_grad(T4 gx, T4 ys, const float* scale, float* ds, cudaStream_t s);
// the same from fp16 copies of gx (times the loss scale; inv[0] = 1 / S) and ys, both with gx's row geometry
int bscale_grad_h(T4 gx, const void* gx16, const void* ys16, const float* inv, const float* scale, float* ds, cudaStream_t s);                          // (C, 2n) -> 2 x (C, n)

struct LatentArgs {
  T4 z1, z2o;             // (128G, L4), (896G, 32)
  const int64_t* rois;
  const float* q;         // (B, 256) query scale (mlp2 output), or (B, V, 256) with view index
  int q_stride;           // floats between segments in q
  int G, c1, c2;
  T4 lat[3];              // unscaled latents (256, L4): all, patient-shuffled, lead-shuffled
  T4 u0[3];               // upsample2(q * lat) (256, L/2), TF32-rounded
  T4 u0lo[3];             // TF32 residual of the same (split-precision input of the decoder's first conv)
  void* u0h[3];           // optional fp16 operand copies (8 channels per 16-byte row) of u0 and of its residual; when set,
  void* u0loh[3];         //   the fp32 residual u0lo is not written (only the fp16 forward convolution reads a residual)
  int n_lat;              // 3 (train) or 1 (extra views: only lat[0] / u0[0])
  int write_lat;          // 0: lat already built, only (re)build u0 from it with another q
  int skip_u0;            // 1: only build the latents (Model_nefnet2 convolves them before the query scaling)
  int round_lat;          // 1: the stored latents are TF32-rounded (they feed a tensor-core convolution)
  int skip_u032;          // 1: the fp32 u0 is not stored (needs u0h: the decoder reads the fp16 copies only)
  int store_mask;         // with write_lat: bit (2 k + half) = store half (0: z1 channels, 1: z2 channels) of lat[k].
                          //   training needs only the z2 halves of lat[0] and lat[2] (latent_bwd rebuilds the rest from z1);
                          //   the extra views of the test phase / gen_ecg re-read both halves of lat[0]
};
int latent_fwd(const LatentArgs& a, cudaStream_t s);
int elem_init();  // shared-memory opt-ins of the kernels in nef_elem.cu (once per device, from nef_init)
struct LatentBwdArgs {
  T4 du0[3]; T4 lat[3]; T4 z1, z2o; const int64_t* rois; const float* q; int q_stride; int G, c1, c2;
  T4 gz1;                 // out: grad wrt pre-ReLU z1 (128G, L4)
  void* gz1_h;            // optional: fp16 copy of gz1 (8 channels per 16-byte row) multiplied by s16[0] (device scalar)
  const float* s16;
  int skip_gz1_32;        // 1: only the fp16 copy gz1_h is written
  void* gz2o_h;           // optional: gz2o is written ONLY as this fp16 copy (times s16[0])
  T4 gz2o;                // out: grad wrt pre-ReLU z2o (896G, 32)
  float* dq;              // out (B, 256), overwritten
  int direct;             // 1: dlat[k] ARE the latent gradients (no upsample / query adjoint, dq untouched); 0: from du0
  T4 dlat[3];
  const void* du0h[3];    // optional: read d u0_k from these loss-scaled fp16 copies (geometry of du0[k], times s16[0]; s16[1] = 1 / S)
  int only_half;          // set by latent_bwd(): -1 = the general kernel does both halves, 1 = the z2 half only (the z1 half ran in latent_bwd_z1_kernel)
};
int latent_bwd(const LatentBwdArgs& a, cudaStream_t s);
struct UpqAdjArgs { T4 du0[3]; T4 lat2[3]; T4 dlat2[3]; const float* q; int q_stride; float* dq; };
int upq_adjoint(const UpqAdjArgs& a, cudaStream_t s);   // Model_nefnet2: d(q * lat') and dq from the decoder input gradients
int replicate_f32(const float* src, float* dst, int n, int total, cudaStream_t s);   // dst[i] = src[i mod n]

int bn_finalize(const BnLayer& bn, int C, double count, const float* gamma, const float* beta, float* rmean,
                float* rvar, int64_t* nbt, int training, cudaStream_t s);
int fill_f32(float* p, float v, int n, cudaStream_t s);
int bn_fold_eval(const float* gamma, const float* beta, const float* rmean, const float* rvar, const float* bias, float* wscale,
                 float* fbias, int C, cudaStream_t s);
int bn_relu(T4 c, const float* scale, const float* shift, T4 out, int upsample, cudaStream_t s);
int up_adjoint(T4 du, T4 da, cudaStream_t s);
int bnbwd_stats(T4 da, T4 c, const BnLayer& bn, cudaStream_t s);
// training == 0: the forward used running statistics, so the batch-mean terms of the BatchNorm gradient vanish
int bnbwd_apply(T4 da, T4 c, const BnLayer& bn, const float* gamma, double count, T4 dc, float* dgamma, float* dbeta,
                int training, cudaStream_t s);
// fp16 decoder dataflow: the same passes on / into fp16 copies (half8 rows, geometry of the given T4; gradients times the
// loss scale lscale[0] = S, lscale[1] = 1 / S)
int bn_relu_h(T4 c, const float* scale, const float* shift, void* out16, T4 og, int upsample, cudaStream_t s);
int up_adjoint_h(const void* du16, T4 du, void* da16, T4 da, cudaStream_t s);
int bnbwd_stats_h(const void* da16, T4 c, const BnLayer& bn, cudaStream_t s);
// up_adjoint_h and bnbwd_stats_h in one pass: da16 (geometry of c) = adjoint of the upsampling of du16, statistics of da16
int up_adjoint_stats_h(const void* du16, T4 du, void* da16, T4 c, const BnLayer& bn, cudaStream_t s);
// da: fp32 incoming gradient (unscaled) or nullptr to read da16 (scaled); dc16 may alias da16
int bnbwd_apply_h(const T4* da, const void* da16, T4 c, const BnLayer& bn, const float* gamma, double count, void* dc16,
                  float* dgamma, float* dbeta, int training, const float* lscale, cudaStream_t s);
int dec_out_fwd(T4 c4, const float* scale, const float* shift, const float* w, const float* b, float* out,
                int out_bstride, cudaStream_t s);
int dec_out_bwd(T4 c4, const BnLayer& bn, const float* w, const float* out, const float* dout, T4 g4, float* dw,
                float* db, cudaStream_t s);
// fp16 decoder dataflow: g4 leaves as the loss-scaled fp16 copy g4h only (geometry of g4), s1 / s2 in scaled units
int dec_out_bwd_h(T4 c4, const BnLayer& bn, const float* w, const float* out, const float* dout, T4 g4, void* g4h,
                  const float* lscale, float* dw, float* db, cudaStream_t s);

}  // namespace nef
