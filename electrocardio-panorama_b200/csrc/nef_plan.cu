// Host-side orchestration of the Nef-Net hot path and the C ABI around it (include/nefnet_b200.h).
// Follows Model_nefnet.forward (network/model_nefnet.py:109-194) and its autograd backward kernel by
// kernel; every launch goes to the caller's stream, all memory comes from the caller's workspace.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nef_elem.cuh"

using namespace nef;

// ---------------------------------------------------------------------------------------------
// errors / library state
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
extern "C" void nef_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* nef_last_error(void) { return g_err; }
static long long g_launches = 0;
extern "C" void nef_count_launch(void) { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
extern "C" int64_t nef_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int nef_version(void) { return NEF_ABI_VERSION; }
extern "C" size_t nef_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(NefConvTerm);
    case 1: return sizeof(NefConvDesc);
    case 2: return sizeof(NefWgradDesc);
    case 3: return sizeof(NefForwardArgs);
    case 4: return sizeof(NefBackwardArgs);
    default: return 0;
  }
}

extern "C" int nef_gconv_fwd_simt(const NefConvDesc* d, nef_stream_t s);
extern "C" int nef_gconv_wgrad_simt(const NefWgradDesc* d, nef_stream_t s);
extern "C" int nef_gconv_fwd_tc(const NefConvDesc* d, nef_stream_t s);
extern "C" int nef_gconv_wgrad_tc(const NefWgradDesc* d, nef_stream_t s);
extern "C" int nef_tc_init(void);
extern "C" int nef_h8_to_ncl(const void* src, float* dst, int B, int C, int L, nef_stream_t s);
extern "C" int nef_bits_to_ncl(const uint32_t* src, float* dst, int B, int C, int L, nef_stream_t s);
extern "C" int nef_codes_to_ncl(const uint32_t* src, float* dst, int B, int C, int L, nef_stream_t s);

static int g_conv_impl = 1;
extern "C" int nef_set_conv_impl(int impl) {
  g_conv_impl = impl;
  return 0;
}
extern "C" int nef_get_conv_impl(void) { return g_conv_impl; }
// Decoder first conv (256 -> 128 on the query-scaled latent): number of split-precision terms.
//   3: x_hi w_hi + x_lo w_hi + x_hi w_lo   2: x_hi w_hi + x_lo w_hi   1: x_hi w_hi only
// Default 3.  Measured worst output error (tests/probe_dec1_terms.py, profiles/r01_dec1_split_terms_probe.txt): 9.4e-4 /
// 6.4e-4 / 5.9e-4 relative for 1 / 2 / 3 terms against the 1e-3 bar; 2 terms save ~0.2 ms per decoder call at batch 256 but
// leave less margin (and move the noisiest weight gradient, W_encoder.layer1.2.conv2, from cosine 0.86 to 0.75 at B = 1).
static int g_dec1_terms = 3;
// Encoder forward convolutions on fp16 operand copies (kind::f16: the significand of TF32, twice the MMA rate, half the
// operand bytes).  Only the forward reads the copies; the backward pass keeps using the fp32 (TF32-rounded) tensors.
static int g_fwd_f16 = 1;
extern "C" int nef_set_fwd_f16(int on) { g_fwd_f16 = on; return 0; }
// Backward pass on fp16 operand copies (needs the fp16 forward): the weight gradients of the big 128-channel layers read
// loss-scaled fp16 copies of dY and fp16 copies of the saved activations MN-major, straight as the bulk copy lands them
// (nef_gconv_wgrad_f16; the TF32 kernel has to re-tile every staged tile in shared memory).  0 = TF32 backward everywhere.
static int g_bwd_f16 = 1;
extern "C" int nef_set_bwd_f16(int on) { g_bwd_f16 = on; return 0; }
// A/B switches (environment, read at nef_init): NEF_KEEP_H32=1 also stores the fp32 hidden activations of the big blocks;
// NEF_K3_TF32=1 runs the forward of w_conv / z1_conv on TF32 operands
static int g_keep_h32 = 0, g_k3_tf32 = 0;
// 1 (default) = fp16 decoder dataflow in training: a1 / u1 / a3 kept as fp16 operand copies only, convolutions 2-4 and every
// decoder data / weight gradient in kind::f16 on loss-scaled fp16 gradient copies (NEF_DEC_F16=0: the TF32 decoder)
static int g_dec_f16 = 1;
// 1 (default) = the stem runs on the tensor cores (nef_stem_tc.cu) whenever only its fp16 copy is kept (NEF_STEM_TC=0: CUDA cores)
static int g_stem_tc = 1;
// 1 (default) = the z2 deflection branch (z2_conv2: residual block, ConvTranspose, residual block on 7 G groups) runs on fp16
// operand copies forward and backward like the big blocks (NEF_Z2_F16=0: TF32 operands, fp32 tensors)
static int g_z2_f16 = 1;
extern "C" int nef_set_dec_f16(int on) { g_dec_f16 = on; return 0; }
extern "C" int nef_gconv_wgrad_f16(const NefWgradDesc* d, const void* dy16, const void* x16, const float* out_scale, nef_stream_t s);
extern "C" int nef_set_dec1_terms(int n) {
  NEF_REQUIRE(n >= 1 && n <= 3, "nef_set_dec1_terms: 1, 2 or 3");
  g_dec1_terms = n;
  return 0;
}
extern "C" int nef_set_exact_simt(int on);
extern "C" int nef_set_exact_elem(int on);
extern "C" int nef_set_exact_tc(int on);
extern "C" int nef_set_exact_fp32(int on) {
  int rc = nef_set_exact_simt(on) | nef_set_exact_elem(on) | nef_set_exact_tc(on);
  NEF_REQUIRE(rc == 0, "nef_set_exact_fp32: cudaMemcpyToSymbol failed");
  return 0;
}

extern "C" int nef_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  NEF_REQUIRE(e == cudaSuccess && n > 0, "nef_init: no CUDA device (%s); this library has no CPU path",
              cudaGetErrorString(e));
  NEF_REQUIRE(device >= 0 && device < n, "nef_init: device %d out of range (%d devices)", device, n);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, device);
  NEF_REQUIRE(p.major == 10, "nef_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
              p.major, p.minor);
  cudaSetDevice(device);
  if (getenv("NEF_KEEP_H32")) g_keep_h32 = atoi(getenv("NEF_KEEP_H32"));
  if (getenv("NEF_K3_TF32")) g_k3_tf32 = atoi(getenv("NEF_K3_TF32"));
  if (getenv("NEF_BWD_F16")) g_bwd_f16 = atoi(getenv("NEF_BWD_F16"));
  if (getenv("NEF_DEC_F16")) g_dec_f16 = atoi(getenv("NEF_DEC_F16"));
  if (getenv("NEF_STEM_TC")) g_stem_tc = atoi(getenv("NEF_STEM_TC"));
  if (getenv("NEF_Z2_F16")) g_z2_f16 = atoi(getenv("NEF_Z2_F16"));
  int rc = elem_init();
  if (rc) return rc;
  return nef_tc_init();
}

extern "C" int nef_gconv_fwd(const NefConvDesc* d, nef_stream_t s) {
  NEF_REQUIRE(d->N == 64 || d->N == 128, "nef_gconv_fwd: N must be 64 or 128 (got %d)", d->N);
  for (int i = 0; i < d->n_terms; ++i)
    NEF_REQUIRE(d->term[i].cin_g % 32 == 0 && d->term[i].taps >= 1 && d->term[i].taps <= 7,
                "nef_gconv_fwd: term %d: cin_g %% 32 and taps in [1,7] required", i);
  return g_conv_impl == 1 ? nef_gconv_fwd_tc(d, s) : nef_gconv_fwd_simt(d, s);
}
extern "C" int nef_gconv_wgrad(const NefWgradDesc* d, nef_stream_t s) {
  NEF_REQUIRE(d->cout_g % 64 == 0 && d->cin_g % 64 == 0, "nef_gconv_wgrad: cout_g, cin_g must be multiples of 64");
  return g_conv_impl == 1 ? nef_gconv_wgrad_tc(d, s) : nef_gconv_wgrad_simt(d, s);
}

// ---------------------------------------------------------------------------------------------
// parameter table (state_dict order of the reference)
// ---------------------------------------------------------------------------------------------
enum ParamIdx {
  P_STEM = 0,
  P_ENC = 1,  // + 2*i + {0: conv1, 1: conv2}
  P_MLP1_W = 7, P_MLP1_B, P_MLP2_W, P_MLP2_B, P_WFE_W, P_WFE_B,
  P_WCONV = 13,   // + {0 conv1, 1 conv2, 2 res.w, 3 res.b}
  P_Z1 = 17, P_Z2C1 = 21, P_Z2A = 25,
  P_CT_W = 29, P_CT_B = 30,
  P_Z2B = 31,
  P_DEC1 = 35,    // + {0 c0.w,1 c0.b,2 bn1.w,3 bn1.b,4 rm,5 rv,6 nbt,7 c3.w,8 c3.b,9 bn4.w,10 bn4.b,11 rm,12 rv,13 nbt}
  P_DEC3 = 49,
  P_OUT_W = 63, P_OUT_B = 64,
  P_COUNT = 65,
  // Model_nefnet2 (variant 2): the G = 1 table above plus its two plain k3 convolutions (model_nefnet2.py:102-107)
  P_S1_W = 65, P_S1_B = 66, P_S2_W = 67, P_S2_B = 68,
  P_COUNT2 = 69
};

struct ParamInfo { std::string name; int64_t numel; };
static std::vector<ParamInfo> param_table(int G, int variant = 1) {
  if (variant == 2) G = 1;   // one single-lead trunk shared by all leads
  std::vector<ParamInfo> t;
  auto add = [&](const std::string& n, int64_t e) { t.push_back({n, e}); };
  add("W_encoder.conv1.weight", 128LL * G * 15);
  for (int i = 0; i < 3; ++i) {
    add("W_encoder.layer1." + std::to_string(i) + ".conv1.weight", 128LL * G * 128 * 7);
    add("W_encoder.layer1." + std::to_string(i) + ".conv2.weight", 128LL * G * 128 * 7);
  }
  add("mlp1.weight", 128 * 12); add("mlp1.bias", 128);
  add("mlp2.weight", 256 * 12); add("mlp2.bias", 256);
  add("w_feature_extractor.0.weight", 128 * 128 * 3); add("w_feature_extractor.0.bias", 128);
  auto block = [&](const std::string& p, int cin_g, int groups) {
    add(p + ".conv1.weight", 128LL * groups * cin_g * 3);
    add(p + ".conv2.weight", 128LL * groups * 128 * 3);
    add(p + ".residual_conv.weight", 128LL * groups * cin_g);
    add(p + ".residual_conv.bias", 128LL * groups);
  };
  block("w_conv.0", 128, G);
  block("z1_conv.0", 64, G);
  block("z2_conv1.0", 64, G);
  block("z2_conv2.0", 128, 7 * G);
  add("z2_conv2.1.weight", 896LL * G * 64 * 2);
  add("z2_conv2.1.bias", 448LL * G);
  block("z2_conv2.2", 64, 7 * G);
  const int cin[2] = {256, 128}, cout[2] = {128, 64};
  const char* st[2] = {"decoder.1", "decoder.3"};
  for (int s = 0; s < 2; ++s) {
    const std::string p = std::string(st[s]) + ".double_conv.";
    add(p + "0.weight", (int64_t)cout[s] * cin[s] * 3); add(p + "0.bias", cout[s]);
    add(p + "1.weight", cout[s]); add(p + "1.bias", cout[s]);
    add(p + "1.running_mean", cout[s]); add(p + "1.running_var", cout[s]); add(p + "1.num_batches_tracked", 1);
    add(p + "3.weight", (int64_t)cout[s] * cout[s] * 3); add(p + "3.bias", cout[s]);
    add(p + "4.weight", cout[s]); add(p + "4.bias", cout[s]);
    add(p + "4.running_mean", cout[s]); add(p + "4.running_var", cout[s]); add(p + "4.num_batches_tracked", 1);
  }
  add("decoder.4.weight", 64 * 3); add("decoder.4.bias", 1);
  if (variant == 2) {
    add("single_conv_z1.0.weight", 128 * 128 * 3); add("single_conv_z1.0.bias", 128);
    add("single_conv_z2.0.weight", 128 * 128 * 3); add("single_conv_z2.0.bias", 128);
  }
  return t;
}
static thread_local std::string g_name_tmp;
extern "C" int nef_param_count(int G) { (void)G; return P_COUNT; }
extern "C" const char* nef_param_name(int G, int i) {
  auto t = param_table(G);
  if (i < 0 || i >= (int)t.size()) return "";
  g_name_tmp = t[i].name;
  return g_name_tmp.c_str();
}
extern "C" int64_t nef_param_numel(int G, int i) {
  auto t = param_table(G);
  if (i < 0 || i >= (int)t.size()) return -1;
  return t[i].numel;
}
// the same for a model variant (1 = Model_nefnet, 2 = Model_nefnet2: the order of the `params` / `grads` arrays)
extern "C" int nef_param_count_v(int G, int variant) { return (int)param_table(G, variant).size(); }
extern "C" const char* nef_param_name_v(int G, int variant, int i) {
  auto t = param_table(G, variant);
  if (i < 0 || i >= (int)t.size()) return "";
  g_name_tmp = t[i].name;
  return g_name_tmp.c_str();
}
extern "C" int64_t nef_param_numel_v(int G, int variant, int i) {
  auto t = param_table(G, variant);
  if (i < 0 || i >= (int)t.size()) return -1;
  return t[i].numel;
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct ConvW {            // one convolution's weights: reference tensor + packed copies
  int pidx;               // index in the parameter table
  int groups, cout_g, cin_g, taps;
  float* pk_f;            // forward packing  [g][t][cin_g/32][8][cout_g][4]
  float* pk_d;            // data-gradient packing (flipped taps, transposed); N = min(cin_g, 128)
  void* pk_h;             // forward packing in fp16 (encoder convolutions only), or nullptr
  void* pk_dh;            // data-gradient packing in fp16 (layers of the fp16 backward), or nullptr
  int src_gmod;           // 0: one weight tensor slice per group; m > 0: group g uses (and accumulates into) slice g mod m
};

struct DecBufs {          // one decoder call
  T4 c1, a1, c2, u1, c3, a3, c4;
  void *a1_h, *u1_h, *a3_h;   // fp16 copies of a1 / u1 / a3 (dec_f16: the only copies written)
  BnLayer bn[4];
  float* out;             // (B, L) saved sigmoid output
};

struct NefPlan {
  int B, G, L, V, L2, L4, C1;
  int variant;            // 1 = Model_nefnet, 2 = Model_nefnet2 (the trunk weights are shared by the G leads; two more convolutions)
  Window win;
  size_t ws_bytes;
  char* base;
  bool bound;
  // bump allocator state (offsets in bytes, relative to base)
  size_t cursor;
  // activations
  T4 s0, eh[3], ey[3], hw, w, h1, z1, xw, hz, z2c, ra, h20, y20, t21, h22, z2o;
  T4 lat[3], u0[3], u0lo[3];
  DecBufs dec[3];
  float *s_in, *q, *rq;
  uint32_t* s0_amax;      // stem max-pool argmax / ReLU codes, indexed like s0
  // one-bit (value != 0) masks of the big post-ReLU activations, written by the forward epilogues and read by the masked
  // data-gradient epilogues instead of the fp32 tensors (NefConvDesc.out_bits / mask_bits)
  uint32_t *b_eh[3], *b_ey[3], *b_hw, *b_w, *b_h1;
  void *s0_h, *eh_h[3], *ey_h[3];   // fp16 operand copies (8 channels per 16-byte row) of s0, eh[i], ey[i]
  void *hw_h, *w_h, *h1_h;          //   ... of w_conv's h and y and of z1_conv's h (operands of the fp16 backward)
  void* GA_h[3];                    // loss-scaled fp16 copies of the gradient buffers GA[i] (backward, bwd_f16)
  float* lscale;                    // device scalars {S, 1 / S}: loss scale of the fp16 gradient copies; [2] = amax scratch
  bool bwd_f16;                     // this forward / backward pair runs the fp16 backward of the big blocks
  bool h_f16_only;                  // the hidden activations of the big blocks (eh, hw, h1) were stored as fp16 copies only
  void *u0_h[3], *u0lo_h[3];        //   ... of the decoder inputs u0 and of their rounding residuals
  void* dec1_lo_h;                  // fp16 residual of the decoder first conv weights (decw[0].pk_h holds the fp16 weights)
  bool fwd_f16;                     // this forward runs the encoder convolutions on them
  // gradients
  T4 GA[3];
  T4 gz2o, gh22, dt21, dte, dto, gy20, gh20, dra, gz2c, ghz, gxw;
  T4 dg4, dg3, du1, dg2, dg1, du0[3];
  void *dg4_h, *dg3_h, *du1_h, *dg2_h, *dg1_h;   // loss-scaled fp16 gradient copies of the decoder backward (dec_f16)
  bool dec_f16;                     // this forward / backward pair runs the fp16 decoder dataflow
  // z2 deflection branch on fp16 copies (z2_f16): activations ra, h20, y20, t21, h22 and the gradients between the layers exist
  // as fp16 copies (+ one-bit planes of h20, y20, h22) only
  void *ra_h, *h20_h, *y20_h, *t21_h, *h22_h;
  uint32_t *b_h20, *b_y20, *b_h22;
  void *gz2o_h, *gh22_h, *dt21_h, *dte_h, *dto_h, *gy20_h, *gh20_h;
  void *ct_h[2], *ct_dh[2];         // fp16 packings of the ConvTranspose weights (forward / data gradient, per tap)
  bool z2_f16;
  bool u0_f16_only;                 // the decoder inputs u0 of this forward exist as fp16 copies (u0_h, u0lo_h) only
  void* du0_h[3];                   // loss-scaled fp16 gradients of the decoder inputs (dec_f16, variant 1)
  float *ds_in, *dq;
  double* bn_stats;       // BatchNorm backward accumulators (s1, s2) of all layers, contiguous
  size_t bn_stats_count;
  // weights
  ConvW enc[6], wc[2], z1c[3], z2c1[3], z2a[2], z2b[3], decw[4];
  // variant 2: single_conv_z1 / single_conv_z2 (applied to the lead means and the picked leads: they are linear, so
  // conv(mean_i z_i) = mean_i conv(z_i)), convolved latents and their gradients, per-lead copies of the shared biases
  ConvW s12[2];
  T4 lat2[3], dlat2[3], dlat[3];
  float* rb[4];
  float *ct_f[2], *ct_d[2];
  float* dec1_lo;         // TF32 residual of the decoder first conv weights, forward packing
  float *fold_scale[4], *fold_bias[4];  // inference: BatchNorm folded into the decoder convolutions (per output channel)
  float *ident_scale, *ident_shift;     // 128 ones / zeros: identity "BatchNorm" for the kernels that still apply one
  bool folded;            // the packed decoder weights of this forward carry the folded BatchNorm
  // saved forward state
  const float* x_in; const float* thetas_in; const float* query_in; const int64_t* rois_in;
  int c1, c2; float drop_p; int bn_training; bool have_fwd;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
  NefPlan* p;
  bool dry;
  size_t cur;
  void* take(size_t bytes) {
    cur = align_up(cur, 256);
    void* r = dry ? nullptr : (void*)(p->base + cur);
    cur += bytes;
    return r;
  }
  T4 t4(int C, int L) {
    T4 t;
    t.C = C; t.B = p->B; t.L = L; t.Lp = L + 2 * NEF_HALO;
    t.cs = (long)p->B * t.Lp;
    const size_t bytes = ((size_t)(C / 4) * t.cs + NEF_GUARD_ROWS) * sizeof(float4);
    t.p = reinterpret_cast<float4*>(take(bytes));
    return t;
  }
  float* f32(size_t n) { return reinterpret_cast<float*>(take(n * sizeof(float))); }
};

static void carve_convw(Carver& c, ConvW& w, int pidx, int groups, int cout_g, int cin_g, int taps, int src_gmod = 0) {
  w.pidx = pidx; w.groups = groups; w.cout_g = cout_g; w.cin_g = cin_g; w.taps = taps; w.src_gmod = src_gmod;
  const size_t n = (size_t)groups * cout_g * cin_g * taps;
  w.pk_f = c.f32(n);
  w.pk_d = c.f32(n);
  w.pk_h = nullptr;
  w.pk_dh = nullptr;
}

static void carve(NefPlan* p, bool dry) {
  Carver c{p, dry, 0};
  c.take(NEF_GUARD_ROWS * sizeof(float4));  // front guard
  const int G = p->G, C1 = p->C1, L4 = p->L4, L2 = p->L2, L = p->L, B = p->B;
  const int m1 = p->variant == 2 ? 1 : 0, m7 = p->variant == 2 ? 7 : 0;   // shared weights: source slice = group mod m
  p->s0 = c.t4(C1, L4);
  p->s0_amax = reinterpret_cast<uint32_t*>(c.take(((size_t)(C1 / 4) * p->s0.cs + NEF_GUARD_ROWS) * sizeof(uint32_t)));
  for (int i = 0; i < 3; ++i) { p->eh[i] = c.t4(C1, L4); p->ey[i] = c.t4(C1, L4); }
  p->hw = c.t4(C1, L4); p->w = c.t4(C1, L4); p->h1 = c.t4(C1, L4); p->z1 = c.t4(C1, L4);
  {
    const size_t bytes = ((size_t)(C1 / 8) * p->s0.cs + NEF_GUARD_ROWS) * 16;
    void** hp[13] = {&p->s0_h, &p->eh_h[0], &p->eh_h[1], &p->eh_h[2], &p->ey_h[0], &p->ey_h[1], &p->ey_h[2], &p->hw_h, &p->w_h,
                     &p->h1_h, &p->GA_h[0], &p->GA_h[1], &p->GA_h[2]};
    for (auto q : hp) *q = c.take(bytes);
    p->lscale = c.f32(4);
  }
  {
    const size_t words = (size_t)(C1 / 32) * p->s0.cs + NEF_GUARD_ROWS;
    uint32_t** bp[9] = {&p->b_eh[0], &p->b_eh[1], &p->b_eh[2], &p->b_ey[0], &p->b_ey[1], &p->b_ey[2], &p->b_hw, &p->b_w, &p->b_h1};
    for (auto q : bp) *q = reinterpret_cast<uint32_t*>(c.take(words * sizeof(uint32_t)));
  }
  p->xw = c.t4(64 * G, p->win.Lw); p->hz = c.t4(C1, p->win.Lw); p->z2c = c.t4(C1, p->win.Lw);
  p->ra = c.t4(896 * G, 16); p->h20 = c.t4(896 * G, 16); p->y20 = c.t4(896 * G, 16);
  p->t21 = c.t4(448 * G, 32); p->h22 = c.t4(896 * G, 32); p->z2o = c.t4(896 * G, 32);
  {
    auto h8 = [&](const T4& t) { return c.take(((size_t)(t.C / 8) * t.cs + NEF_GUARD_ROWS) * 16); };
    auto bits = [&](const T4& t) { return reinterpret_cast<uint32_t*>(c.take(((size_t)(t.C / 32) * t.cs + NEF_GUARD_ROWS) * sizeof(uint32_t))); };
    p->ra_h = h8(p->ra); p->h20_h = h8(p->h20); p->y20_h = h8(p->y20); p->t21_h = h8(p->t21); p->h22_h = h8(p->h22);
    p->b_h20 = bits(p->h20); p->b_y20 = bits(p->y20); p->b_h22 = bits(p->h22);
  }
  for (int k = 0; k < 3; ++k) { p->lat[k] = c.t4(256, L4); p->u0[k] = c.t4(256, L2); p->u0lo[k] = c.t4(256, L2); }
  for (int k = 0; k < 3; ++k) {
    const size_t bytes = ((size_t)(256 / 8) * p->u0[k].cs + NEF_GUARD_ROWS) * 16;
    p->u0_h[k] = c.take(bytes);
    p->u0lo_h[k] = c.take(bytes);
  }
  for (int k = 0; k < 3; ++k) {
    DecBufs& d = p->dec[k];
    d.c1 = c.t4(128, L2); d.a1 = c.t4(128, L2); d.c2 = c.t4(128, L2); d.u1 = c.t4(128, L);
    d.c3 = c.t4(64, L); d.a3 = c.t4(64, L); d.c4 = c.t4(64, L);
    d.a1_h = c.take(((size_t)(128 / 8) * d.a1.cs + NEF_GUARD_ROWS) * 16);
    d.u1_h = c.take(((size_t)(128 / 8) * d.u1.cs + NEF_GUARD_ROWS) * 16);
    d.a3_h = c.take(((size_t)(64 / 8) * d.a3.cs + NEF_GUARD_ROWS) * 16);
    d.out = c.f32((size_t)B * L);
    const int ch[4] = {128, 128, 64, 64};
    for (int i = 0; i < 4; ++i) {
      d.bn[i].scale = c.f32(ch[i]); d.bn[i].shift = c.f32(ch[i]);
      d.bn[i].mean = c.f32(ch[i]); d.bn[i].invstd = c.f32(ch[i]);
    }
  }
  // BatchNorm forward partial records (per 128-row tile) and backward double accumulators
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 4; ++i) {
      const int ch = i < 2 ? 128 : 64;
      const long rows = (long)B * ((i < 2 ? L2 : L) + 2 * NEF_HALO);
      const int n_rec = (int)((rows + 127) / 128);
      p->dec[k].bn[i].n_rec = n_rec;
      p->dec[k].bn[i].sum = c.f32((size_t)n_rec * ch);
      p->dec[k].bn[i].sq = c.f32((size_t)n_rec * ch);
    }
  p->bn_stats_count = 3 * 4 * 2 * 128;
  p->bn_stats = reinterpret_cast<double*>(c.take(p->bn_stats_count * sizeof(double)));
  if (!dry) {
    double* q = p->bn_stats;
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 4; ++i) {
        p->dec[k].bn[i].s1 = q; q += 128;
        p->dec[k].bn[i].s2 = q; q += 128;
      }
  }
  p->s_in = c.f32((size_t)B * C1); p->ds_in = c.f32((size_t)B * C1);
  p->q = c.f32((size_t)B * 256); p->dq = c.f32((size_t)B * 256);
  p->rq = c.f32((size_t)B * (p->V > 0 ? p->V : 1) * 256);
  for (int i = 0; i < 3; ++i) p->GA[i] = c.t4(C1, L4);
  p->gz2o = c.t4(896 * G, 32); p->gh22 = c.t4(896 * G, 32); p->dt21 = c.t4(448 * G, 32);
  p->dte = c.t4(448 * G, 16); p->dto = c.t4(448 * G, 16);
  p->gy20 = c.t4(896 * G, 16); p->gh20 = c.t4(896 * G, 16); p->dra = c.t4(896 * G, 16);
  p->gz2c = c.t4(C1, p->win.Lw); p->ghz = c.t4(C1, p->win.Lw); p->gxw = c.t4(64 * G, p->win.Lw);
  {
    auto h8 = [&](const T4& t) { return c.take(((size_t)(t.C / 8) * t.cs + NEF_GUARD_ROWS) * 16); };
    p->gz2o_h = h8(p->gz2o); p->gh22_h = h8(p->gh22); p->dt21_h = h8(p->dt21); p->dte_h = h8(p->dte); p->dto_h = h8(p->dto);
    p->gy20_h = h8(p->gy20); p->gh20_h = h8(p->gh20);
  }
  p->dg4 = c.t4(64, L); p->dg3 = c.t4(64, L); p->du1 = c.t4(128, L); p->dg2 = c.t4(128, L2); p->dg1 = c.t4(128, L2);
  for (int k = 0; k < 3; ++k) p->du0[k] = c.t4(256, L2);
  for (int k = 0; k < 3; ++k) p->du0_h[k] = c.take(((size_t)(256 / 8) * p->du0[k].cs + NEF_GUARD_ROWS) * 16);
  p->dg4_h = c.take(((size_t)(64 / 8) * p->dg4.cs + NEF_GUARD_ROWS) * 16);
  p->dg3_h = c.take(((size_t)(64 / 8) * p->dg3.cs + NEF_GUARD_ROWS) * 16);
  p->du1_h = c.take(((size_t)(128 / 8) * p->du1.cs + NEF_GUARD_ROWS) * 16);
  p->dg2_h = c.take(((size_t)(128 / 8) * p->dg2.cs + NEF_GUARD_ROWS) * 16);
  p->dg1_h = c.take(((size_t)(128 / 8) * p->dg1.cs + NEF_GUARD_ROWS) * 16);
  // weights
  for (int i = 0; i < 6; ++i) {
    carve_convw(c, p->enc[i], P_ENC + i, G, 128, 128, 7, m1);
    p->enc[i].pk_h = c.take((size_t)G * 128 * 128 * 7 * 2);
    p->enc[i].pk_dh = c.take((size_t)G * 128 * 128 * 7 * 2);
  }
  carve_convw(c, p->wc[0], P_WCONV + 0, G, 128, 128, 3, m1);
  carve_convw(c, p->wc[1], P_WCONV + 1, G, 128, 128, 3, m1);
  carve_convw(c, p->z1c[0], P_Z1 + 0, G, 128, 64, 3, m1);
  carve_convw(c, p->z1c[1], P_Z1 + 1, G, 128, 128, 3, m1);
  carve_convw(c, p->z1c[2], P_Z1 + 2, G, 128, 64, 1, m1);
  for (ConvW* w : {&p->wc[0], &p->wc[1], &p->z1c[0], &p->z1c[1], &p->z1c[2]}) {
    w->pk_h = c.take((size_t)w->groups * w->cout_g * w->cin_g * w->taps * 2);
    w->pk_dh = c.take((size_t)w->groups * w->cout_g * w->cin_g * w->taps * 2);
  }
  carve_convw(c, p->z2c1[0], P_Z2C1 + 0, G, 128, 64, 3, m1);
  carve_convw(c, p->z2c1[1], P_Z2C1 + 1, G, 128, 128, 3, m1);
  carve_convw(c, p->z2c1[2], P_Z2C1 + 2, G, 128, 64, 1, m1);
  carve_convw(c, p->z2a[0], P_Z2A + 0, 7 * G, 128, 128, 3, m7);
  carve_convw(c, p->z2a[1], P_Z2A + 1, 7 * G, 128, 128, 3, m7);
  carve_convw(c, p->z2b[0], P_Z2B + 0, 7 * G, 128, 64, 3, m7);
  carve_convw(c, p->z2b[1], P_Z2B + 1, 7 * G, 128, 128, 3, m7);
  carve_convw(c, p->z2b[2], P_Z2B + 2, 7 * G, 128, 64, 1, m7);
  for (ConvW* w : {&p->z2a[0], &p->z2a[1], &p->z2b[0], &p->z2b[1], &p->z2b[2]}) {
    w->pk_h = c.take((size_t)w->groups * w->cout_g * w->cin_g * w->taps * 2);
    w->pk_dh = c.take((size_t)w->groups * w->cout_g * w->cin_g * w->taps * 2);
  }
  carve_convw(c, p->decw[0], P_DEC1 + 0, 1, 128, 256, 3);
  p->decw[0].pk_h = c.take((size_t)128 * 256 * 3 * 2);
  p->dec1_lo_h = c.take((size_t)128 * 256 * 3 * 2);
  p->dec1_lo = c.f32((size_t)128 * 256 * 3);
  carve_convw(c, p->decw[1], P_DEC1 + 7, 1, 128, 128, 3);
  carve_convw(c, p->decw[2], P_DEC3 + 0, 1, 64, 128, 3);
  carve_convw(c, p->decw[3], P_DEC3 + 7, 1, 64, 64, 3);
  for (int i = 0; i < 4; ++i) {   // fp16 packings of the decoder (dec_f16): forward of convolutions 2-4, every data gradient
    ConvW& w = p->decw[i];
    if (i > 0) w.pk_h = c.take((size_t)w.cout_g * w.cin_g * w.taps * 2);
    w.pk_dh = c.take((size_t)w.cout_g * w.cin_g * w.taps * 2);
  }
  for (int i = 0; i < 4; ++i) { p->fold_scale[i] = c.f32(128); p->fold_bias[i] = c.f32(128); }
  p->ident_scale = c.f32(128); p->ident_shift = c.f32(128);
  for (int t = 0; t < 2; ++t) {
    p->ct_f[t] = c.f32((size_t)7 * G * 128 * 64);
    p->ct_d[t] = c.f32((size_t)7 * G * 128 * 64);
    p->ct_h[t] = c.take((size_t)7 * G * 128 * 64 * 2);
    p->ct_dh[t] = c.take((size_t)7 * G * 128 * 64 * 2);
  }
  if (p->variant == 2) {
    carve_convw(c, p->s12[0], P_S1_W, 1, 128, 128, 3);
    carve_convw(c, p->s12[1], P_S2_W, 1, 128, 128, 3);
    for (int k = 0; k < 3; ++k) { p->lat2[k] = c.t4(256, L4); p->dlat2[k] = c.t4(256, L4); p->dlat[k] = c.t4(256, L4); }
    const int nb[4] = {128 * G, 128 * G, 896 * G, 448 * G};
    for (int i = 0; i < 4; ++i) p->rb[i] = c.f32(nb[i]);
  }
  c.take(NEF_GUARD_ROWS * sizeof(float4));
  p->ws_bytes = align_up(c.cur, 256);
}

extern "C" int nef_plan_create_v(int B, int G, int L, int V, int variant, NefPlan** out) {
  NEF_REQUIRE(B >= 1 && G >= 1 && L >= 16 && L % 4 == 0, "nef_plan_create: need B>=1, G>=1, L>=16, L %% 4 == 0 (B=%d G=%d L=%d)",
              B, G, L);
  NEF_REQUIRE(variant == 1 || variant == 2, "nef_plan_create_v: variant must be 1 (Model_nefnet) or 2 (Model_nefnet2)");
  NEF_REQUIRE((long)B * (L + 2 * NEF_HALO) < (1L << 31), "nef_plan_create: B * (L + halo) must stay below 2^31 rows per chunk plane (the stem kernels multiply chunk strides in 32 bits; B=%d L=%d)", B, L);
  NefPlan* p = new NefPlan();
  memset(p, 0, sizeof(NefPlan));
  p->variant = variant;
  p->B = B; p->G = G; p->L = L; p->V = V; p->L2 = L / 2; p->L4 = L / 4; p->C1 = 128 * G;
  p->win = centre_window(p->L4);
  carve(p, true);
  *out = p;
  return 0;
}
extern "C" int nef_plan_create(int B, int G, int L, int V, NefPlan** out) { return nef_plan_create_v(B, G, L, V, 1, out); }
extern "C" void nef_plan_destroy(NefPlan* p) { delete p; }
extern "C" size_t nef_plan_workspace_bytes(const NefPlan* p) { return p->ws_bytes; }
extern "C" int nef_plan_bind(NefPlan* p, void* ws, size_t bytes, nef_stream_t s) {
  NEF_REQUIRE(bytes >= p->ws_bytes, "nef_plan_bind: workspace too small (%zu < %zu)", bytes, p->ws_bytes);
  NEF_REQUIRE(((uintptr_t)ws & 255) == 0, "nef_plan_bind: workspace must be 256-byte aligned");
  p->base = (char*)ws;
  carve(p, false);
  cudaError_t e = cudaMemsetAsync(ws, 0, p->ws_bytes, (cudaStream_t)s);
  NEF_REQUIRE(e == cudaSuccess, "nef_plan_bind: memset failed: %s", cudaGetErrorString(e));
  {
    int rc = fill_f32(p->ident_scale, 1.0f, 128, (cudaStream_t)s);
    if (rc) return rc;
  }
  p->bound = true;
  p->have_fwd = false;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// descriptor builders
// ---------------------------------------------------------------------------------------------
struct CD {
  NefConvDesc d;
  CD(int groups, int N, const T4& space) {
    memset(&d, 0, sizeof(d));
    d.groups = groups; d.N = N; d.rows = space.cs; d.Lp = space.Lp; d.L = space.L;
    d.mask_scale = 1.f;
  }
  CD& term(const T4& x, int off, int gs, int cin_g, int taps, const float* w) {
    NefConvTerm& t = d.term[d.n_terms++];
    t.x = reinterpret_cast<const float*>(x.p); t.x_cstride = x.cs; t.x_c4_off = off; t.x_c4_gstride = gs;
    t.cin_g = cin_g; t.taps = taps; t.tap_off = -(taps / 2); t.w = w;
    return *this;
  }
  CD& out(const T4& y, int off, int gs, int lmul = 1, int ladd = 0) {
    d.y = reinterpret_cast<float*>(y.p); d.y_cstride = y.cs; d.y_c4_off = off; d.y_c4_gstride = gs;
    d.y_Lp = y.Lp; d.y_lmul = lmul; d.y_ladd = ladd;
    return *this;
  }
  CD& bias(const float* b) { d.bias = b; return *this; }
  CD& res(const T4& r, int off, int gs) {
    d.res = reinterpret_cast<const float*>(r.p); d.res_cstride = r.cs; d.res_c4_off = off; d.res_c4_gstride = gs;
    return *this;
  }
  // residual operand from the fp16 copy of a tensor with geometry (cs, off, gs in 4-channel chunk units), times scale[0]
  CD& res16(const void* r16, long cs, int off, int gs, const float* scale = nullptr) {
    d.res16 = r16; d.res_cstride = cs; d.res_c4_off = off; d.res_c4_gstride = gs; d.res16_scale = scale;
    return *this;
  }
  CD& relu() { d.relu = 1; return *this; }
  CD& drop(float p, uint64_t seed) { d.drop_p = p; d.drop_seed = seed; return *this; }
  CD& bscale(const float* s) { d.bscale = s; return *this; }
  CD& bsgrad(float* g) { d.bscale_grad = g; return *this; }
  CD& mask(const T4& m, int off, int gs, int mode, float scale) {
    d.mask = reinterpret_cast<const float*>(m.p); d.mask_cstride = m.cs; d.mask_c4_off = off; d.mask_c4_gstride = gs;
    d.mask_mode = mode; d.mask_scale = scale;
    return *this;
  }
  CD& stats(float* s1, float* s2) { d.stat_sum = s1; d.stat_sq = s2; return *this; }
  // fp16 operand term: x16 = fp16 copy (8 channels per row) of a tensor with the row space of `space`, cin real channels
  CD& term16(const void* x16, long cs, int off8, int gs8, int cin, int taps, const void* w16) {
    NefConvTerm& t = d.term[d.n_terms++];
    t.x = reinterpret_cast<const float*>(x16); t.x_cstride = cs; t.x_c4_off = off8; t.x_c4_gstride = gs8;
    t.cin_g = cin / 2; t.taps = taps; t.tap_off = -(taps / 2); t.w = reinterpret_cast<const float*>(w16); t.x_f16 = 1;
    return *this;
  }
  CD& y16(void* h) { d.y16 = h; return *this; }                     // also store the fp16 copy of the output
  CD& y16s(void* h, const float* scale) { d.y16 = h; d.y16_scale = scale; return *this; }   // ... multiplied by scale[0] (device)
  CD& obits(uint32_t* b) { d.out_bits = b; return *this; }          // record (output != 0) bits next to the output
  CD& mbits(const uint32_t* b) { d.mask_bits = b; return *this; }   // read the mask from bits (same chunk offsets as .mask)
  CD& round() { d.round_tf32 = 1; return *this; }
  int run(cudaStream_t s) { return nef_gconv_fwd(&d, (nef_stream_t)s); }
};

static int wgrad(const T4& dy, int dy_off, int dy_gs, int cout_g, const T4& x, int x_off, int x_gs, int cin_g,
                 int groups, int taps, float* dw, int64_t sg, int64_t sm, int64_t sn, int64_t st, float* db,
                 cudaStream_t s, int wg_mod = 0) {
  if (!dw) return 0;
  NefWgradDesc d;
  memset(&d, 0, sizeof(d));
  d.dy = reinterpret_cast<const float*>(dy.p); d.dy_cstride = dy.cs; d.dy_c4_off = dy_off; d.dy_c4_gstride = dy_gs;
  d.x = reinterpret_cast<const float*>(x.p); d.x_cstride = x.cs; d.x_c4_off = x_off; d.x_c4_gstride = x_gs;
  d.cout_g = cout_g; d.cin_g = cin_g; d.groups = groups; d.taps = taps; d.tap_off = -(taps / 2);
  d.rows = dy.cs; d.dw = dw; d.sg = sg; d.sm = sm; d.sn = sn; d.st = st; d.db = db; d.wg_mod = wg_mod;
  return nef_gconv_wgrad(&d, (nef_stream_t)s);
}
// standard Conv1d weight (groups*cout_g, cin_g, taps)
static int wgrad_std(const T4& dy, int dy_off, int dy_gs, const T4& x, int x_off, int x_gs, const ConvW& w, float* dw,
                     float* db, cudaStream_t s) {
  return wgrad(dy, dy_off, dy_gs, w.cout_g, x, x_off, x_gs, w.cin_g, w.groups, w.taps, dw,
               (int64_t)w.cout_g * w.cin_g * w.taps, (int64_t)w.cin_g * w.taps, w.taps, 1, db, s, w.src_gmod);
}

#define RUN(x)            \
  do {                    \
    int rc__ = (x);       \
    if (rc__) return rc__; \
  } while (0)

// packing jobs are queued in a table and launched together (nef_pack_weights_batch)
static int queue_pack(NefPackTable& t, const float* src, float* dst, int groups, int N, int K, int taps, int64_t sg, int64_t sn,
                      int64_t sk, int64_t st, int flags, cudaStream_t s, const float* nscale = nullptr, int gmod = 0) {
  if (t.n == NEF_PACK_MAX) RUN(nef_pack_weights_batch(&t, s));
  NefPackJob& q = t.job[t.n++];
  q.gmod = gmod;
  q.src = src; q.dst = dst; q.groups = groups; q.N = N; q.K = K; q.taps = taps;
  q.sg = sg; q.sn = sn; q.sk = sk; q.st = st; q.flags = flags; q.first_block = 0; q.nscale = nscale;
  return 0;
}
static int pack_fwd(NefPackTable& t, const ConvW& w, const float* const* P, cudaStream_t s, const float* nscale = nullptr) {
  return queue_pack(t, P[w.pidx], w.pk_f, w.groups, w.cout_g, w.cin_g, w.taps, (int64_t)w.cout_g * w.cin_g * w.taps,
                    (int64_t)w.cin_g * w.taps, w.taps, 1, 0, s, nscale, w.src_gmod);
}
// dgrad: N' = cin_g (split into sub-groups of 128 when larger; only for groups == 1), K' = cout_g, flipped taps
static int pack_dgrad(NefPackTable& t, const ConvW& w, const float* const* P, cudaStream_t s) {
  if (w.cin_g > 128) {
    const int sub = w.cin_g / 128;
    return queue_pack(t, P[w.pidx], w.pk_d, sub, 128, w.cout_g, w.taps, (int64_t)128 * w.taps, w.taps,
                      (int64_t)w.cin_g * w.taps, 1, 1, s);   // (decoder only: one group)
  }
  return queue_pack(t, P[w.pidx], w.pk_d, w.groups, w.cin_g, w.cout_g, w.taps, (int64_t)w.cout_g * w.cin_g * w.taps,
                    w.taps, (int64_t)w.cin_g * w.taps, 1, 1, s, nullptr, w.src_gmod);
}

// the same in fp16 (layers of the fp16 backward)
static int pack_dgrad_h(NefPackTable& t, const ConvW& w, const float* const* P, cudaStream_t s) {
  if (w.cin_g > 128)   // decoder first convolution: 256 input channels as two sub-groups of 128 (as pack_dgrad)
    return queue_pack(t, P[w.pidx], reinterpret_cast<float*>(w.pk_dh), w.cin_g / 128, 128, w.cout_g, w.taps, (int64_t)128 * w.taps,
                      w.taps, (int64_t)w.cin_g * w.taps, 1, 1 | 4, s);
  return queue_pack(t, P[w.pidx], reinterpret_cast<float*>(w.pk_dh), w.groups, w.cin_g, w.cout_g, w.taps,
                    (int64_t)w.cout_g * w.cin_g * w.taps, w.taps, (int64_t)w.cin_g * w.taps, 1, 1 | 4, s, nullptr, w.src_gmod);
}

static int pack_dec1_lo(NefPackTable& t, NefPlan* p, const float* const* P, cudaStream_t s, const float* nscale = nullptr) {
  const ConvW& w = p->decw[0];
  return queue_pack(t, P[w.pidx], p->dec1_lo, 1, 128, 256, 3, 0, 256 * 3, 3, 1, 2, s, nscale);
}

// decoder layer i: parameter indices of its conv bias and of its BatchNorm (weight, bias, running_mean, running_var, nbt)
static const int kDecBias[4] = {P_DEC1 + 1, P_DEC1 + 8, P_DEC3 + 1, P_DEC3 + 8};
static const int kDecBn[4] = {P_DEC1 + 2, P_DEC1 + 9, P_DEC3 + 2, P_DEC3 + 9};

// Inference (running statistics, nothing saved for backward): fold every decoder BatchNorm into its convolution and queue
// the decoder weight packing with the folded scales.  The decoder then runs conv + bias + ReLU epilogues only.
static int queue_decoder_packs(NefPlan* p, NefPackTable& t, const float* const* P, bool fold, cudaStream_t s) {
  p->folded = fold;
  for (int i = 0; i < 4; ++i) {
    const float* ns = nullptr;
    if (fold) {
      RUN(bn_fold_eval(P[kDecBn[i]], P[kDecBn[i] + 1], P[kDecBn[i] + 2], P[kDecBn[i] + 3], P[kDecBias[i]], p->fold_scale[i],
                       p->fold_bias[i], p->decw[i].cout_g, s));
      ns = p->fold_scale[i];
    }
    if (i == 0 && p->fwd_f16) {  // first conv in kind::f16: weights and their residuals in fp16
      const ConvW& w = p->decw[0];
      RUN(queue_pack(t, P[w.pidx], reinterpret_cast<float*>(w.pk_h), 1, 128, 256, 3, 0, 256 * 3, 3, 1, 4, s, ns));
      RUN(queue_pack(t, P[w.pidx], reinterpret_cast<float*>(p->dec1_lo_h), 1, 128, 256, 3, 0, 256 * 3, 3, 1, 4 | 2, s, ns));
      continue;
    }
    if (i > 0 && (p->dec_f16 || (fold && p->fwd_f16 && g_dec_f16))) {   // fp16 operand packing (with the folded scales in inference)
      const ConvW& w = p->decw[i];
      RUN(queue_pack(t, P[w.pidx], reinterpret_cast<float*>(w.pk_h), 1, w.cout_g, w.cin_g, 3, 0, (int64_t)w.cin_g * 3, 3, 1, 4, s, ns));
      continue;
    }
    RUN(pack_fwd(t, p->decw[i], P, s, ns));
    if (i == 0) RUN(pack_dec1_lo(t, p, P, s, ns));
  }
  return 0;
}

template <class F>
static int for_all_convw(NefPlan* p, F f) {
  for (int i = 0; i < 6; ++i) RUN(f(p->enc[i]));
  for (int i = 0; i < 2; ++i) RUN(f(p->wc[i]));
  for (int i = 0; i < 3; ++i) RUN(f(p->z1c[i]));
  for (int i = 0; i < 3; ++i) RUN(f(p->z2c1[i]));
  for (int i = 0; i < 2; ++i) RUN(f(p->z2a[i]));
  for (int i = 0; i < 3; ++i) RUN(f(p->z2b[i]));
  for (int i = 0; i < 4; ++i) RUN(f(p->decw[i]));
  if (p->variant == 2)
    for (int i = 0; i < 2; ++i) RUN(f(p->s12[i]));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// residual block: h = drop(relu(conv1(x))) ; y = relu(conv2(h) + r(x))   (resnet_1d.py:39-53, model_nefnet.py:48-60)
struct BlockIO {
  T4 x; int x_off, x_gs;   // input view
  T4 h, y;
  const ConvW* c1; const ConvW* c2; const ConvW* cr;  // cr == nullptr: identity residual
  const float* res_bias;
  int groups;
  uint32_t* hbits = nullptr;   // one-bit masks of h and y (big layers only)
  uint32_t* ybits = nullptr;
  const void* x16 = nullptr;   // fp16 operand copies: when x16 is set the two convolutions run in kind::f16 (h16 required)
  void* h16 = nullptr;         // fp16 copies of h / y written by the epilogues (operands of the next convolution and / or of
  void* y16 = nullptr;         //   the fp16 weight gradients); optional, also without x16
  bool h_f16_only = false;     // do not store the fp32 h at all (needs x16 and h16: nothing reads it then)
  bool y_f16_only = false;     // the same for y (needs y16; its readers take the fp16 copy and the bit plane)
};

static int block_fwd(const BlockIO& io, float drop_p, uint64_t seed, const float* bscale, cudaStream_t s) {
  CD a(io.groups, 128, io.x);
  if (io.x16) a.term16(io.x16, io.x.cs, io.x_off / 2, io.x_gs / 2, io.c1->cin_g, io.c1->taps, io.c1->pk_h);
  else a.term(io.x, io.x_off, io.x_gs, io.c1->cin_g, io.c1->taps, io.c1->pk_f);
  if (io.h16) a.y16(io.h16);
  a.out(io.h, 0, 32).relu().round();
  if (io.hbits) a.obits(io.hbits);
  if (drop_p > 0.f) a.drop(drop_p, seed);
  if (io.h_f16_only && io.x16 && io.h16) a.d.y = nullptr;
  RUN(a.run(s));
  CD b(io.groups, 128, io.x);
  if (io.x16) b.term16(io.h16, io.h.cs, 0, 16, 128, io.c2->taps, io.c2->pk_h);
  else b.term(io.h, 0, 32, 128, io.c2->taps, io.c2->pk_f);
  b.out(io.y, 0, 32).relu().round();
  if (io.y16) b.y16(io.y16);
  if (io.cr && io.x16) b.term16(io.x16, io.x.cs, io.x_off / 2, io.x_gs / 2, io.cr->cin_g, 1, io.cr->pk_h).bias(io.res_bias);
  else if (io.cr) b.term(io.x, io.x_off, io.x_gs, io.cr->cin_g, 1, io.cr->pk_f).bias(io.res_bias);
  else if (io.x16) b.res16(io.x16, io.x.cs, io.x_off, io.x_gs);   // identity residual from the fp16 copy (same significand)
  else b.res(io.x, io.x_off, io.x_gs);
  if (bscale) b.bscale(bscale);
  if (io.ybits) b.obits(io.ybits);
  if (io.y_f16_only && io.y16) b.d.y = nullptr;
  RUN(b.run(s));
  return 0;
}

static int decoder_fwd(NefPlan* p, const float* const* P, int slot, int lslot, const T4& u0, const T4& u0lo, int training,
                       float* out_user, int out_bstride, cudaStream_t s) {
  DecBufs& d = p->dec[slot];
  const void* u0h = p->u0_h[lslot];       // fp16 copies of u0 / its residual (latent slot lslot), used when p->fwd_f16
  const void* u0loh = p->u0lo_h[lslot];
  const int B = p->B;
  if (p->folded) {  // inference: conv (BatchNorm folded into weights and bias) + ReLU epilogues, one upsample pass
    const T4 ins[4] = {u0, d.a1, d.u1, d.a3};
    const T4 outs[4] = {d.a1, d.c2, d.a3, d.c4};
    const bool h = p->fwd_f16 && g_dec_f16;   // a1 / u1 / a3 as fp16 operand copies only (c2 and c4 stay fp32)
    const void* in16[4] = {nullptr, d.a1_h, d.u1_h, d.a3_h};
    void* out16[4] = {d.a1_h, nullptr, d.a3_h, nullptr};
    for (int i = 0; i < 4; ++i) {
      const ConvW& w = p->decw[i];
      CD c(1, w.cout_g, ins[i]);
      if (i == 0 && p->fwd_f16) {
        c.term16(u0h, u0.cs, 0, 0, 256, 3, w.pk_h);
        if (g_dec1_terms >= 2) c.term16(u0loh, u0.cs, 0, 0, 256, 3, w.pk_h);
        if (g_dec1_terms >= 3) c.term16(u0h, u0.cs, 0, 0, 256, 3, p->dec1_lo_h);
      } else if (h) {
        c.term16(in16[i], ins[i].cs, 0, 0, w.cin_g, 3, w.pk_h);
      } else {
        c.term(ins[i], 0, 0, w.cin_g, 3, w.pk_f);
        if (i == 0) {
          if (g_dec1_terms >= 2) c.term(u0lo, 0, 0, 256, 3, w.pk_f);
          if (g_dec1_terms >= 3) c.term(ins[i], 0, 0, 256, 3, p->dec1_lo);
        }
      }
      c.out(outs[i], 0, 0).bias(p->fold_bias[i]).relu();
      if (h && out16[i]) { c.y16(out16[i]); c.d.y = nullptr; }
      else if (i == 0 || i == 2) c.round();  // a1, a3 feed the next tensor-core convolution directly
      RUN(c.run(s));
      if (i == 1 && h) RUN(bn_relu_h(d.c2, p->ident_scale, p->ident_shift, d.u1_h, d.u1, 1, s));
      else if (i == 1) RUN(bn_relu(d.c2, p->ident_scale, p->ident_shift, d.u1, 1, s));
    }
    RUN(dec_out_fwd(d.c4, p->ident_scale, p->ident_shift, P[P_OUT_W], P[P_OUT_B], d.out, p->L, s));
  } else {
  struct Lay { const ConvW* w; int pb; T4 in; T4 c; int bnp; double count; };
  const Lay lay[4] = {{&p->decw[0], P_DEC1 + 1, u0, d.c1, P_DEC1 + 2, (double)B * p->L2},
                      {&p->decw[1], P_DEC1 + 8, d.a1, d.c2, P_DEC1 + 9, (double)B * p->L2},
                      {&p->decw[2], P_DEC3 + 1, d.u1, d.c3, P_DEC3 + 2, (double)B * p->L},
                      {&p->decw[3], P_DEC3 + 8, d.a3, d.c4, P_DEC3 + 9, (double)B * p->L}};
  for (int i = 0; i < 4; ++i) {
    const Lay& l = lay[i];
    CD c(1, l.w->cout_g, l.in);
    if (i == 0 && p->fwd_f16) {  // split precision in kind::f16 (same significand as TF32): x_hi w_hi + x_lo w_hi + x_hi w_lo
      c.term16(u0h, u0.cs, 0, 0, 256, 3, l.w->pk_h);
      if (g_dec1_terms >= 2) c.term16(u0loh, u0.cs, 0, 0, 256, 3, l.w->pk_h);
      if (g_dec1_terms >= 3) c.term16(u0h, u0.cs, 0, 0, 256, 3, p->dec1_lo_h);
    } else if (p->dec_f16) {     // the post-BatchNorm activations exist as fp16 copies only
      const void* in16[4] = {nullptr, d.a1_h, d.u1_h, d.a3_h};
      c.term16(in16[i], l.in.cs, 0, 0, l.w->cin_g, 3, l.w->pk_h);
    } else {
      c.term(l.in, 0, 0, l.w->cin_g, 3, l.w->pk_f);
      if (i == 0) {  // split precision: x_hi w_hi + x_lo w_hi + x_hi w_lo  (this layer dominates the TF32 error budget)
        if (g_dec1_terms >= 2) c.term(u0lo, 0, 0, 256, 3, l.w->pk_f);
        if (g_dec1_terms >= 3) c.term(l.in, 0, 0, 256, 3, p->dec1_lo);
      }
    }
    c.out(l.c, 0, 0).bias(P[l.pb]);
    if (training) c.stats(d.bn[i].sum, d.bn[i].sq);
    RUN(c.run(s));
    RUN(bn_finalize(d.bn[i], l.w->cout_g, l.count, P[l.bnp], P[l.bnp + 1], const_cast<float*>(P[l.bnp + 2]),
                    const_cast<float*>(P[l.bnp + 3]),
                    reinterpret_cast<int64_t*>(const_cast<float*>(P[l.bnp + 4])), training, s));
    if (p->dec_f16) {
      if (i == 0) RUN(bn_relu_h(d.c1, d.bn[0].scale, d.bn[0].shift, d.a1_h, d.a1, 0, s));
      if (i == 1) RUN(bn_relu_h(d.c2, d.bn[1].scale, d.bn[1].shift, d.u1_h, d.u1, 1, s));
      if (i == 2) RUN(bn_relu_h(d.c3, d.bn[2].scale, d.bn[2].shift, d.a3_h, d.a3, 0, s));
      continue;
    }
    if (i == 0) RUN(bn_relu(d.c1, d.bn[0].scale, d.bn[0].shift, d.a1, 0, s));
    if (i == 1) RUN(bn_relu(d.c2, d.bn[1].scale, d.bn[1].shift, d.u1, 1, s));
    if (i == 2) RUN(bn_relu(d.c3, d.bn[2].scale, d.bn[2].shift, d.a3, 0, s));
  }
  RUN(dec_out_fwd(d.c4, d.bn[3].scale, d.bn[3].shift, P[P_OUT_W], P[P_OUT_B], d.out, p->L, s));
  }
  if (out_user) {
    cudaError_t e = cudaMemcpy2DAsync(out_user, (size_t)out_bstride * sizeof(float), d.out, (size_t)p->L * sizeof(float),
                                      (size_t)p->L * sizeof(float), B, cudaMemcpyDeviceToDevice, s);
    NEF_REQUIRE(e == cudaSuccess, "decoder_fwd: output copy failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

static int latents_to_decoders(NefPlan* p, const float* const* P, const float* query_theta, const float* rest_theta,
                               const int64_t* rois, int phase, int training, int V, float* out, float* out_p,
                               float* out_l, float* rest_out, bool only_views, cudaStream_t s) {
  const int B = p->B;
  const bool v2 = p->variant == 2;
  LatentArgs la = {};
  la.z1 = p->z1; la.z2o = p->z2o; la.rois = rois; la.G = p->G; la.c1 = p->c1; la.c2 = p->c2;
  for (int k = 0; k < 3; ++k) {
    la.lat[k] = p->lat[k]; la.u0[k] = p->u0[k]; la.u0lo[k] = p->u0lo[k];
    la.u0h[k] = p->fwd_f16 ? p->u0_h[k] : nullptr;
    la.u0loh[k] = p->fwd_f16 ? p->u0lo_h[k] : nullptr;
  }
  la.skip_u032 = p->u0_f16_only ? 1 : 0;
  if (!only_views) {
    RUN(angular_fwd(query_theta, P[P_MLP2_W], P[P_MLP2_B], p->q, B, 256, s));
    la.q = p->q; la.q_stride = 256; la.n_lat = 3; la.write_lat = 1;
    la.store_mask = phase == NEF_PHASE_TEST ? 3 : (2 | 32);
    if (v2) { la.store_mask = 63; la.skip_u0 = 1; la.round_lat = 1; }   // every half of every latent feeds a convolution
    RUN(latent_fwd(la, s));
    if (v2) {
      // single_conv_z1 / single_conv_z2 (model_nefnet2.py:140,148) on the z1 / z2 halves of the three latents -- applied
      // after the lead mean / pick instead of before (the convolutions are linear) -- then the query scaling + upsampling
      for (int k = 0; k < 3; ++k)
        for (int h = 0; h < 2; ++h) {
          CD c(1, 128, p->lat[k]);
          c.term(p->lat[k], 32 * h, 0, 128, 3, p->s12[h].pk_f).out(p->lat2[k], 32 * h, 0).bias(P[h ? P_S2_B : P_S1_B]);
          RUN(c.run(s));
        }
      LatentArgs lu = la;
      lu.write_lat = 0; lu.skip_u0 = 0; lu.n_lat = 1;
      for (int k = 0; k < 3; ++k) {
        lu.lat[0] = p->lat2[k]; lu.u0[0] = p->u0[k]; lu.u0lo[0] = p->u0lo[k]; lu.u0h[0] = la.u0h[k]; lu.u0loh[0] = la.u0loh[k];
        RUN(latent_fwd(lu, s));
      }
      la.lat[0] = p->lat2[0];   // the extra views below re-read the convolved mean latent
      la.skip_u0 = 0;
    }
    float* outs[3] = {out, out_p, out_l};
    for (int k = 0; k < 3; ++k) RUN(decoder_fwd(p, P, k, k, p->u0[k], p->u0lo[k], training, outs[k], p->L, s));
  } else {
    // gen_ecg: build lat[0] only (mean latents); q unused for that -> use rq view 0 below
    la.q = p->rq; la.q_stride = V * 256; la.n_lat = 1; la.write_lat = 1; la.store_mask = 3;
  }
  if (V > 0 && (phase == NEF_PHASE_TEST || only_views)) {
    NEF_REQUIRE(V <= p->V, "nef_forward: V=%d exceeds the plan's V=%d", V, p->V);
    RUN(angular_fwd(rest_theta, P[P_MLP2_W], P[P_MLP2_B], p->rq, B * V, 256, s));
    for (int v = 0; v < V; ++v) {
      la.q = p->rq + (size_t)v * 256; la.q_stride = V * 256; la.n_lat = 1;
      la.write_lat = (only_views && v == 0) ? 1 : 0;
      la.store_mask = 3;
      RUN(latent_fwd(la, s));
      RUN(decoder_fwd(p, P, 0, 0, p->u0[0], p->u0lo[0], training, rest_out + (size_t)v * p->L, V * p->L, s));
    }
  }
  return 0;
}

extern "C" int nef_forward(NefPlan* p, const NefForwardArgs* a, nef_stream_t sv) {
  cudaStream_t s = (cudaStream_t)sv;
  NEF_REQUIRE(p && p->bound, "nef_forward: plan not bound to a workspace");
  const float* const* P = a->params;
  const int G = p->G, B = p->B;
  const bool v2 = p->variant == 2;
  NEF_REQUIRE(a->lead_choice_z1 >= 0 && a->lead_choice_z1 < G && a->lead_choice_z2 >= 0 && a->lead_choice_z2 < G,
              "nef_forward: lead choice out of range");
  p->have_fwd = false;
  p->c1 = a->lead_choice_z1; p->c2 = a->lead_choice_z2; p->drop_p = a->drop_p; p->bn_training = a->bn_training;
  p->x_in = a->x; p->thetas_in = a->input_thetas; p->query_in = a->query_theta; p->rois_in = a->rois;

  NefPackTable packs;
  packs.n = 0;
  p->fwd_f16 = g_fwd_f16 && g_conv_impl == 1;
  p->bwd_f16 = p->fwd_f16 && g_bwd_f16 && a->save_for_backward;
  // fp16 decoder: training-mode BatchNorm (with running statistics and a backward the conv biases need gradients from the
  // fp32 tensors; inference without a backward folds the BatchNorm and takes its own path)
  p->dec_f16 = p->fwd_f16 && g_dec_f16 && a->bn_training && (p->bwd_f16 || !a->save_for_backward);
  // the fp32 u0 has readers only without the fp16 forward (first decoder convolution) or in the TF32 decoder backward
  p->u0_f16_only = p->fwd_f16 && (p->dec_f16 || !a->save_for_backward);
  // variant 2: the biases its grouped epilogues index per group are shared by the leads -> one copy per lead
  const float* b_z1 = P[P_Z1 + 3];
  const float* b_z2c1 = P[P_Z2C1 + 3];
  const float* b_z2b = P[P_Z2B + 3];
  const float* b_ct = P[P_CT_B];
  if (v2) {
    RUN(replicate_f32(P[P_Z1 + 3], p->rb[0], 128, 128 * G, s));
    RUN(replicate_f32(P[P_Z2C1 + 3], p->rb[1], 128, 128 * G, s));
    RUN(replicate_f32(P[P_Z2B + 3], p->rb[2], 896, 896 * G, s));
    RUN(replicate_f32(P[P_CT_B], p->rb[3], 448, 448 * G, s));
    b_z1 = p->rb[0]; b_z2c1 = p->rb[1]; b_z2b = p->rb[2]; b_ct = p->rb[3];
  }
  // the fp32 hidden activations of the big blocks have a reader only in the TF32 backward (its weight gradients)
  p->h_f16_only = p->fwd_f16 && (p->bwd_f16 || !a->save_for_backward) && !g_keep_h32;
  p->z2_f16 = p->h_f16_only && g_z2_f16;
  RUN(for_all_convw(p, [&](const ConvW& w) {
    if (&w >= p->decw && &w < p->decw + 4) return 0;
    const bool k3 = (&w >= p->wc && &w < p->wc + 2) || (&w >= p->z1c && &w < p->z1c + 3);
    const bool z2 = (&w >= p->z2a && &w < p->z2a + 2) || (&w >= p->z2b && &w < p->z2b + 3);
    if (p->fwd_f16 && w.pk_h && !(k3 && g_k3_tf32) && !(z2 && !p->z2_f16))
      return queue_pack(packs, P[w.pidx], reinterpret_cast<float*>(w.pk_h), w.groups, w.cout_g, w.cin_g, w.taps,
                        (int64_t)w.cout_g * w.cin_g * w.taps, (int64_t)w.cin_g * w.taps, w.taps, 1, 4, s, nullptr, w.src_gmod);
    return pack_fwd(packs, w, P, s);
  }));
  RUN(queue_decoder_packs(p, packs, P, !a->bn_training && !a->save_for_backward, s));
  for (int t = 0; t < 2; ++t)  // ConvTranspose1d weight (Cin_total, Cout/groups, 2): one 1x1 conv per tap
    RUN(queue_pack(packs, P[P_CT_W] + t, p->z2_f16 ? reinterpret_cast<float*>(p->ct_h[t]) : p->ct_f[t], 7 * G, 64, 128, 1, 128LL * 64 * 2, 2,
                   64 * 2, 0, p->z2_f16 ? 4 : 0, s, nullptr, v2 ? 7 : 0));
  RUN(nef_pack_weights_batch(&packs, s));

  if (p->fwd_f16 && p->h_f16_only && g_stem_tc)   // fp16 dataflow: the first block reads only the fp16 copy
    RUN(stem_tc_fwd(a->x, P[P_STEM], p->s0, a->save_for_backward ? p->s0_amax : nullptr, p->s0_h, G, s, v2 ? 1 : 0));
  else
    RUN(stem_fwd(a->x, P[P_STEM], p->s0, a->save_for_backward ? p->s0_amax : nullptr, p->fwd_f16 ? p->s0_h : nullptr, G, s,
                 p->h_f16_only ? 0 : 1, v2 ? 1 : 0));
  RUN(angular_fwd(a->input_thetas, P[P_MLP1_W], P[P_MLP1_B], p->s_in, B * G, 128, s));
  const float dp = a->drop_p;
  const uint64_t seed = a->drop_seed * 16;
  for (int i = 0; i < 3; ++i) {
    BlockIO io{i == 0 ? p->s0 : p->ey[i - 1], 0, 32, p->eh[i], p->ey[i], &p->enc[2 * i], &p->enc[2 * i + 1], nullptr,
               nullptr, G};
    if (a->save_for_backward) { io.hbits = p->b_eh[i]; io.ybits = p->b_ey[i]; }
    if (p->fwd_f16) {
      io.x16 = i == 0 ? p->s0_h : p->ey_h[i - 1];
      io.h16 = p->eh_h[i];
      io.y16 = p->ey_h[i];
      io.h_f16_only = p->h_f16_only;
      io.y_f16_only = p->h_f16_only && (i < 2 || !g_k3_tf32);   // every reader takes the fp16 copy / the bit plane
    }
    RUN(block_fwd(io, dp, seed + i, i == 2 ? p->s_in : nullptr, s));
  }
  {
    BlockIO io{p->ey[2], 0, 32, p->hw, p->w, &p->wc[0], &p->wc[1], nullptr, nullptr, G};
    if (a->save_for_backward) { io.hbits = p->b_hw; io.ybits = p->b_w; }
    if (p->fwd_f16 && !g_k3_tf32) {
      io.x16 = p->ey_h[2]; io.h16 = p->hw_h; io.y16 = p->w_h; io.h_f16_only = p->h_f16_only; io.y_f16_only = p->h_f16_only;
    } else if (p->bwd_f16) { io.h16 = p->hw_h; io.y16 = p->w_h; }
    RUN(block_fwd(io, dp, seed + 3, nullptr, s));
  }
  {
    BlockIO io{p->w, 0, 32, p->h1, p->z1, &p->z1c[0], &p->z1c[1], &p->z1c[2], b_z1, G};
    if (a->save_for_backward) io.hbits = p->b_h1;
    if (p->fwd_f16 && !g_k3_tf32) { io.x16 = p->w_h; io.h16 = p->h1_h; io.h_f16_only = p->h_f16_only; }
    else if (p->bwd_f16) io.h16 = p->h1_h;
    RUN(block_fwd(io, dp, seed + 4, nullptr, s));
  }
  RUN(window_extract(p->w, p->xw, G, p->win, (p->fwd_f16 && !g_k3_tf32 && p->h_f16_only) ? p->w_h : nullptr, s));
  {
    BlockIO io{p->xw, 0, 16, p->hz, p->z2c, &p->z2c1[0], &p->z2c1[1], &p->z2c1[2], b_z2c1, G};
    RUN(block_fwd(io, dp, seed + 5, nullptr, s));
  }
  RUN(roi_align_fwd(p->z2c, a->rois, p->ra, p->win, p->L4, s, p->z2_f16 ? p->ra_h : nullptr));
  {
    BlockIO io{p->ra, 0, 32, p->h20, p->y20, &p->z2a[0], &p->z2a[1], nullptr, nullptr, 7 * G};
    if (p->z2_f16) {   // fp16 copies (+ bit planes for the backward) only
      io.x16 = p->ra_h; io.h16 = p->h20_h; io.y16 = p->y20_h; io.h_f16_only = true; io.y_f16_only = true;
      if (a->save_for_backward) { io.hbits = p->b_h20; io.ybits = p->b_y20; }
    }
    RUN(block_fwd(io, dp, seed + 6, nullptr, s));
  }
  for (int t = 0; t < 2; ++t) {  // ConvTranspose1d(k2, s2): out[2l + t] = W_t x[l] + b
    CD c(7 * G, 64, p->y20);
    if (p->z2_f16) {
      c.term16(p->y20_h, p->y20.cs, 0, 16, 128, 1, p->ct_h[t]).out(p->t21, 0, 16, 2, t).bias(b_ct).y16(p->t21_h);
      c.d.y = nullptr;
    } else {
      c.term(p->y20, 0, 32, 128, 1, p->ct_f[t]).out(p->t21, 0, 16, 2, t).bias(b_ct).round();
    }
    c.d.term[0].tap_off = 0;
    RUN(c.run(s));
  }
  {
    BlockIO io{p->t21, 0, 16, p->h22, p->z2o, &p->z2b[0], &p->z2b[1], &p->z2b[2], b_z2b, 7 * G};
    if (p->z2_f16) {   // the block output z2o stays fp32 (latent mixing reads it)
      io.x16 = p->t21_h; io.h16 = p->h22_h; io.h_f16_only = true;
      if (a->save_for_backward) io.hbits = p->b_h22;
    }
    RUN(block_fwd(io, dp, seed + 7, nullptr, s));
  }
  NEF_REQUIRE(!(v2 && a->phase == NEF_PHASE_GEN), "nef_forward: phase 'gen' is not built for Model_nefnet2 (its gen_ecg cannot consume it)");
  if (a->phase == NEF_PHASE_GEN) {
    RUN(nef_cbl4_to_ncl(reinterpret_cast<const float*>(p->z1.p), a->out, B, p->C1, p->L4, sv));
    RUN(nef_cbl4_to_ncl(reinterpret_cast<const float*>(p->z2o.p), a->out_p, B, 896 * G, 32, sv));
    return 0;
  }
  RUN(latents_to_decoders(p, P, a->query_theta, a->rest_theta, a->rois, a->phase, a->bn_training,
                          a->phase == NEF_PHASE_TEST ? p->V : 0, a->out, a->out_p, a->out_l, a->rest_out, false, s));
  p->have_fwd = a->save_for_backward != 0 && a->phase == NEF_PHASE_TRAIN;
  return 0;
}

extern "C" int nef_gen_ecg(NefPlan* p, const float* const* P, const float* z1, const float* z2, const float* query_theta,
                           const int64_t* rois, int V, float* out, nef_stream_t sv) {
  cudaStream_t s = (cudaStream_t)sv;
  NEF_REQUIRE(p && p->bound, "nef_gen_ecg: plan not bound to a workspace");
  NEF_REQUIRE(V >= 1 && V <= p->V, "nef_gen_ecg: V=%d not in [1, plan V=%d]", V, p->V);
  NEF_REQUIRE(p->variant == 1, "nef_gen_ecg: not built for Model_nefnet2 (the reference's gen_ecg cannot consume its own phase 'gen' output)");
  p->have_fwd = false;
  p->c1 = 0; p->c2 = 0;
  NefPackTable packs;
  packs.n = 0;
  p->fwd_f16 = g_fwd_f16 && g_conv_impl == 1;
  p->dec_f16 = false;
  p->z2_f16 = false;
  p->u0_f16_only = p->fwd_f16;
  RUN(queue_decoder_packs(p, packs, P, true, s));   // gen_ecg runs the module in eval mode (model_nefnet.py:197)
  RUN(nef_pack_weights_batch(&packs, s));
  RUN(nef_ncl_to_cbl4(z1, reinterpret_cast<float*>(p->z1.p), p->B, p->C1, p->L4, 0, sv));
  RUN(nef_ncl_to_cbl4(z2, reinterpret_cast<float*>(p->z2o.p), p->B, 896 * p->G, 32, 0, sv));
  return latents_to_decoders(p, P, nullptr, query_theta, rois, NEF_PHASE_TEST, 0, V, nullptr, nullptr, nullptr, out,
                             true, s);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// residual block backward.  gy = grad w.r.t. the block's pre-ReLU output (already masked).
// Produces gx through `fin` (a descriptor prepared by the caller with its output / mask / extras).
struct BlockBwd {
  BlockIO io;
  T4 gy, gh;
  float *dw1, *dw2, *dwr, *dbr;
  // fp16 backward (nullptr = the TF32 kernels): loss-scaled fp16 copies of gy (given) and gh (written here), the scale
  // scalars {S, 1 / S}; the saved activations' fp16 copies are io.x16 / io.h16
  const void* gy16 = nullptr;
  void* gh16 = nullptr;
  const float* lscale = nullptr;
};
// weight gradient from fp16 copies (same geometry arguments as wgrad_std; chunk offsets in 4-channel units)
static int wgrad_h(const void* dy16, const T4& dy, int dy_off, int dy_gs, const void* x16, const T4& x, int x_off, int x_gs,
                   const ConvW& w, float* dw, const float* inv_scale, cudaStream_t s) {
  if (!dw) return 0;
  NefWgradDesc d;
  memset(&d, 0, sizeof(d));
  d.dy_cstride = dy.cs; d.dy_c4_off = dy_off; d.dy_c4_gstride = dy_gs;
  d.x_cstride = x.cs; d.x_c4_off = x_off; d.x_c4_gstride = x_gs;
  d.cout_g = w.cout_g; d.cin_g = w.cin_g; d.groups = w.groups; d.taps = w.taps; d.tap_off = -(w.taps / 2);
  d.rows = dy.cs; d.dw = dw;
  d.sg = (int64_t)w.cout_g * w.cin_g * w.taps; d.sm = (int64_t)w.cin_g * w.taps; d.sn = w.taps; d.st = 1;
  d.wg_mod = w.src_gmod;
  return nef_gconv_wgrad_f16(&d, dy16, x16, inv_scale, (nef_stream_t)s);
}
extern "C" int nef_bias_grad_tc(const NefWgradDesc* d, nef_stream_t s);
extern "C" int nef_bias_grad_h(const NefWgradDesc* d, const void* dy16, const float* scale, nef_stream_t s);
static int block_bwd(const BlockBwd& b, float drop_p, CD& fin, cudaStream_t s) {
  const BlockIO& io = b.io;
  // Full fp16 backward of the block: every operand of its convolutions' data and weight gradients is an fp16 copy (the
  // gradient copies carry the loss scale S: accumulators are multiplied by 1 / S, lscale[1]).  gh is then kept in fp16 only.
  const bool h = b.lscale && b.gy16 && b.gh16 && io.x16 && io.h16 && io.c1->pk_dh && io.c2->pk_dh && (!io.cr || io.cr->pk_dh);
  if (h) {
    const float* inv = b.lscale + 1;
    RUN(wgrad_h(b.gy16, b.gy, 0, 32, io.h16, io.h, 0, 32, *io.c2, b.dw2, inv, s));
    if (io.cr) {
      RUN(wgrad_h(b.gy16, b.gy, 0, 32, io.x16, io.x, io.x_off, io.x_gs, *io.cr, b.dwr, inv, s));
      if (b.dbr) {
        NefWgradDesc d;
        memset(&d, 0, sizeof(d));
        d.dy_cstride = b.gy.cs; d.dy_c4_off = 0; d.dy_c4_gstride = 32;
        d.cout_g = io.cr->cout_g; d.groups = io.cr->groups; d.rows = b.gy.cs; d.db = b.dbr; d.wg_mod = io.cr->src_gmod;
        RUN(nef_bias_grad_h(&d, b.gy16, inv, (nef_stream_t)s));   // the fp32 gy of such a block is not stored
      }
    }
    CD a(io.groups, 128, io.x);
    a.term16(b.gy16, b.gy.cs, 0, 16, 128, io.c2->taps, io.c2->pk_dh).out(b.gh, 0, 32)
        .mask(io.h, 0, 32, 1, 1.f / (1.f - drop_p)).round().y16s(b.gh16, b.lscale);
    a.d.acc_scale = inv;
    a.d.y = nullptr;   // gh: fp16 copy only
    if (io.hbits) a.mbits(io.hbits);
    RUN(a.run(s));
    RUN(wgrad_h(b.gh16, b.gh, 0, 32, io.x16, io.x, io.x_off, io.x_gs, *io.c1, b.dw1, inv, s));
    fin.term16(b.gh16, b.gh.cs, 0, 16, 128, io.c1->taps, io.c1->pk_dh);
    fin.d.acc_scale = inv;
    if (io.cr) fin.term16(b.gy16, b.gy.cs, 0, 16, 128, 1, io.cr->pk_dh);
    else fin.res16(b.gy16, b.gy.cs, 0, 32, inv);
    RUN(fin.run(s));
    return 0;
  }
  if (b.gy16 && io.h16) RUN(wgrad_h(b.gy16, b.gy, 0, 32, io.h16, io.h, 0, 32, *io.c2, b.dw2, b.lscale + 1, s));
  else RUN(wgrad_std(b.gy, 0, 32, io.h, 0, 32, *io.c2, b.dw2, nullptr, s));
  if (io.cr) RUN(wgrad_std(b.gy, 0, 32, io.x, io.x_off, io.x_gs, *io.cr, b.dwr, b.dbr, s));
  CD a(io.groups, 128, io.x);
  a.term(b.gy, 0, 32, 128, io.c2->taps, io.c2->pk_d).out(b.gh, 0, 32).mask(io.h, 0, 32, 1, 1.f / (1.f - drop_p)).round();
  if (io.hbits) a.mbits(io.hbits);
  if (b.gh16) a.y16s(b.gh16, b.lscale);
  RUN(a.run(s));
  if (b.gh16 && io.x16) RUN(wgrad_h(b.gh16, b.gh, 0, 32, io.x16, io.x, io.x_off, io.x_gs, *io.c1, b.dw1, b.lscale + 1, s));
  else RUN(wgrad_std(b.gh, 0, 32, io.x, io.x_off, io.x_gs, *io.c1, b.dw1, nullptr, s));
  // gx = conv1^T(gh) + (identity: gy | 1x1: res^T(gy))
  fin.term(b.gh, 0, 32, 128, io.c1->taps, io.c1->pk_d);
  if (io.cr) fin.term(b.gy, 0, 32, 128, 1, io.cr->pk_d);
  else fin.res(b.gy, 0, 32);
  RUN(fin.run(s));
  return 0;
}

static int decoder_bwd(NefPlan* p, const float* const* P, float* const* Gd, int slot, const float* dout, cudaStream_t s) {
  DecBufs& d = p->dec[slot];
  const int B = p->B;
  const double n2 = (double)B * p->L2, n1 = (double)B * p->L;
  auto g = [&](int i) { return Gd[i]; };
  // A convolution bias followed by a training-mode BatchNorm has no gradient (the batch mean absorbs it; the reference's
  // autograd value is rounding noise, oracle ZERO_GRAD_PARAMS): its slot is left at zero instead of spending a reduction.
  auto gb = [&](int i) { return p->bn_training ? (float*)nullptr : Gd[i]; };
  // output layer + bn4 statistics
  RUN(dec_out_bwd(d.c4, d.bn[3], P[P_OUT_W], d.out, dout, p->dg4, g(P_OUT_W), g(P_OUT_B), s));
  RUN(bnbwd_apply(p->dg4, d.c4, d.bn[3], P[P_DEC3 + 9], n1, p->dg4, g(P_DEC3 + 9), g(P_DEC3 + 10), p->bn_training, s));
  RUN(wgrad_std(p->dg4, 0, 0, d.a3, 0, 0, p->decw[3], g(P_DEC3 + 7), gb(P_DEC3 + 8), s));
  {
    CD c(1, 64, p->dg4);
    c.term(p->dg4, 0, 0, 64, 3, p->decw[3].pk_d).out(p->dg3, 0, 0);
    RUN(c.run(s));
  }
  RUN(bnbwd_stats(p->dg3, d.c3, d.bn[2], s));
  RUN(bnbwd_apply(p->dg3, d.c3, d.bn[2], P[P_DEC3 + 2], n1, p->dg3, g(P_DEC3 + 2), g(P_DEC3 + 3), p->bn_training, s));
  RUN(wgrad_std(p->dg3, 0, 0, d.u1, 0, 0, p->decw[2], g(P_DEC3 + 0), gb(P_DEC3 + 1), s));
  {
    CD c(1, 128, p->dg3);
    c.term(p->dg3, 0, 0, 64, 3, p->decw[2].pk_d).out(p->du1, 0, 0);
    RUN(c.run(s));
  }
  RUN(up_adjoint(p->du1, p->dg2, s));
  RUN(bnbwd_stats(p->dg2, d.c2, d.bn[1], s));
  RUN(bnbwd_apply(p->dg2, d.c2, d.bn[1], P[P_DEC1 + 9], n2, p->dg2, g(P_DEC1 + 9), g(P_DEC1 + 10), p->bn_training, s));
  RUN(wgrad_std(p->dg2, 0, 0, d.a1, 0, 0, p->decw[1], g(P_DEC1 + 7), gb(P_DEC1 + 8), s));
  {
    CD c(1, 128, p->dg2);
    c.term(p->dg2, 0, 0, 128, 3, p->decw[1].pk_d).out(p->dg1, 0, 0);
    RUN(c.run(s));
  }
  RUN(bnbwd_stats(p->dg1, d.c1, d.bn[0], s));
  RUN(bnbwd_apply(p->dg1, d.c1, d.bn[0], P[P_DEC1 + 2], n2, p->dg1, g(P_DEC1 + 2), g(P_DEC1 + 3), p->bn_training, s));
  RUN(wgrad_std(p->dg1, 0, 0, p->u0[slot], 0, 0, p->decw[0], g(P_DEC1 + 0), gb(P_DEC1 + 1), s));
  {
    CD c(2, 128, p->dg1);  // 256 output channels as two sub-groups reading the same input
    c.term(p->dg1, 0, 0, 128, 3, p->decw[0].pk_d).out(p->du0[slot], 0, 32).round();
    RUN(c.run(s));
  }
  return 0;
}

// The same on the fp16 decoder dataflow (dec_f16): the gradients between the layers are loss-scaled fp16 copies only, every
// data / weight gradient runs in kind::f16 on them and on the fp16 activation copies a1_h / u1_h / a3_h / u0_h.
static int decoder_bwd_h(NefPlan* p, const float* const* P, float* const* Gd, int slot, const float* dout, cudaStream_t s) {
  DecBufs& d = p->dec[slot];
  const int B = p->B;
  const double n2 = (double)B * p->L2, n1 = (double)B * p->L;
  const float* ls = p->lscale;
  const float* inv = ls + 1;
  auto g = [&](int i) { return Gd[i]; };
  // data gradient of decoder convolution i from the fp16 gradient copy gin16 (row space `space`) into the fp16 copy gout16
  auto dgrad = [&](int i, const T4& space, const void* gin16, void* gout16, const T4& outgeom) {
    const ConvW& w = p->decw[i];
    CD c(1, w.cin_g, space);
    c.term16(gin16, space.cs, 0, 0, w.cout_g, 3, w.pk_dh).out(outgeom, 0, 0).y16s(gout16, ls);
    c.d.acc_scale = inv;
    c.d.y = nullptr;
    return c.run(s);
  };
  // output layer + bn4 statistics (fp32 g4, unscaled), then bn4's backward into the scaled fp16 copy
  static const bool out_h = !getenv("NEF_DEC_OUT_H") || atoi(getenv("NEF_DEC_OUT_H")) != 0;   // A/B switch: 0 = fp32 g4
  if (out_h) {   // g4 as the loss-scaled fp16 copy only, bn4's backward in place on it
    RUN(dec_out_bwd_h(d.c4, d.bn[3], P[P_OUT_W], d.out, dout, p->dg4, p->dg4_h, ls, g(P_OUT_W), g(P_OUT_B), s));
    RUN(bnbwd_apply_h(nullptr, p->dg4_h, d.c4, d.bn[3], P[P_DEC3 + 9], n1, p->dg4_h, g(P_DEC3 + 9), g(P_DEC3 + 10), 1, ls, s));
  } else {
    RUN(dec_out_bwd(d.c4, d.bn[3], P[P_OUT_W], d.out, dout, p->dg4, g(P_OUT_W), g(P_OUT_B), s));
    RUN(bnbwd_apply_h(&p->dg4, nullptr, d.c4, d.bn[3], P[P_DEC3 + 9], n1, p->dg4_h, g(P_DEC3 + 9), g(P_DEC3 + 10), 1, ls, s));
  }
  RUN(wgrad_h(p->dg4_h, p->dg4, 0, 0, d.a3_h, d.a3, 0, 0, p->decw[3], g(P_DEC3 + 7), inv, s));
  RUN(dgrad(3, p->dg4, p->dg4_h, p->dg3_h, p->dg3));
  RUN(bnbwd_stats_h(p->dg3_h, d.c3, d.bn[2], s));
  RUN(bnbwd_apply_h(nullptr, p->dg3_h, d.c3, d.bn[2], P[P_DEC3 + 2], n1, p->dg3_h, g(P_DEC3 + 2), g(P_DEC3 + 3), 1, ls, s));
  RUN(wgrad_h(p->dg3_h, p->dg3, 0, 0, d.u1_h, d.u1, 0, 0, p->decw[2], g(P_DEC3 + 0), inv, s));
  RUN(dgrad(2, p->dg3, p->dg3_h, p->du1_h, p->du1));
  RUN(up_adjoint_stats_h(p->du1_h, p->du1, p->dg2_h, d.c2, d.bn[1], s));
  RUN(bnbwd_apply_h(nullptr, p->dg2_h, d.c2, d.bn[1], P[P_DEC1 + 9], n2, p->dg2_h, g(P_DEC1 + 9), g(P_DEC1 + 10), 1, ls, s));
  RUN(wgrad_h(p->dg2_h, p->dg2, 0, 0, d.a1_h, d.a1, 0, 0, p->decw[1], g(P_DEC1 + 7), inv, s));
  RUN(dgrad(1, p->dg2, p->dg2_h, p->dg1_h, p->dg1));
  RUN(bnbwd_stats_h(p->dg1_h, d.c1, d.bn[0], s));
  RUN(bnbwd_apply_h(nullptr, p->dg1_h, d.c1, d.bn[0], P[P_DEC1 + 2], n2, p->dg1_h, g(P_DEC1 + 2), g(P_DEC1 + 3), 1, ls, s));
  RUN(wgrad_h(p->dg1_h, p->dg1, 0, 0, p->u0_h[slot], p->u0[slot], 0, 0, p->decw[0], g(P_DEC1 + 0), inv, s));
  {
    CD c(2, 128, p->dg1);  // 256 output channels as two sub-groups reading the same input
    c.term16(p->dg1_h, p->dg1.cs, 0, 0, 128, 3, p->decw[0].pk_dh).out(p->du0[slot], 0, 32);
    c.d.acc_scale = inv;
    if (p->variant == 1) { c.y16s(p->du0_h[slot], ls); c.d.y = nullptr; }   // latent_bwd reads the loss-scaled fp16 copy
    else c.round();                                                        // Model_nefnet2: fp32 result for upq_adjoint
    RUN(c.run(s));
  }
  return 0;
}

static int zero_t4(const T4& t, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(t.p, 0, (size_t)(t.C / 4) * t.cs * sizeof(float4), s);
  NEF_REQUIRE(e == cudaSuccess, "memset failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int nef_backward(NefPlan* p, const NefBackwardArgs* a, nef_stream_t sv) {
  cudaStream_t s = (cudaStream_t)sv;
  NEF_REQUIRE(p && p->bound && p->have_fwd, "nef_backward: no saved training forward on this plan");
  const float* const* P = a->params;
  float* const* Gd = a->grads;
  const int G = p->G, B = p->B;
  const bool v2 = p->variant == 2;
  const float dp = p->drop_p;
  p->have_fwd = false;

  NefPackTable packs;
  packs.n = 0;
  const bool f16 = p->bwd_f16 && g_conv_impl == 1;
  RUN(for_all_convw(p, [&](const ConvW& w) {
    const bool dec = &w >= p->decw && &w < p->decw + 4;
    const bool z2 = (&w >= p->z2a && &w < p->z2a + 2) || (&w >= p->z2b && &w < p->z2b + 3);
    if (f16 && w.pk_dh && (!dec || p->dec_f16) && (!z2 || p->z2_f16)) return pack_dgrad_h(packs, w, P, s);   // the TF32 data-gradient packing of these layers is not read
    return pack_dgrad(packs, w, P, s);
  }));
  for (int t = 0; t < 2; ++t)  // ConvTranspose dgrad: dx[l] = sum_t W_t^T dy[2l + t] ; N' = ci (128), K' = co (64)
    RUN(queue_pack(packs, P[P_CT_W] + t, (f16 && p->z2_f16) ? reinterpret_cast<float*>(p->ct_dh[t]) : p->ct_d[t], 7 * G, 128, 64, 1, 128LL * 64 * 2,
                   64 * 2, 2, 0, (f16 && p->z2_f16) ? 4 : 0, s, nullptr, v2 ? 7 : 0));
  RUN(nef_pack_weights_batch(&packs, s));
  // zero the BatchNorm backward accumulators (s1, s2 of every layer)
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 4; ++i) cudaMemsetAsync(p->dec[k].bn[i].s1, 0, 2 * 128 * sizeof(double), s);

  const float* douts[3] = {a->dout, a->dout_p, a->dout_l};
  const float* ls = p->lscale;
  if (f16) RUN(grad_loss_scale(a->dout, a->dout_p, a->dout_l, (long)B * p->L, p->lscale, s));
  for (int k = 0; k < 3; ++k) {
    if (douts[k]) RUN((f16 && p->dec_f16) ? decoder_bwd_h(p, P, Gd, k, douts[k], s) : decoder_bwd(p, P, Gd, k, douts[k], s));
    else RUN(zero_t4(p->du0[k], s));
  }
  // latents
  LatentBwdArgs lb = {};
  if (v2) {
    // Model_nefnet2: (query scaling, upsampling) adjoint first, then single_conv_z1 / z2 backward on the latent halves
    // (weight gradients summed over the three latents, data gradients per latent), then the lead distribution below
    UpqAdjArgs ua = {};
    for (int k = 0; k < 3; ++k) { ua.du0[k] = p->du0[k]; ua.lat2[k] = p->lat2[k]; ua.dlat2[k] = p->dlat2[k]; }
    ua.q = p->q; ua.q_stride = 256; ua.dq = p->dq;
    RUN(upq_adjoint(ua, s));
    for (int h = 0; h < 2; ++h)
      for (int k = 0; k < 3; ++k) {
        RUN(wgrad_std(p->dlat2[k], 32 * h, 0, p->lat[k], 32 * h, 0, p->s12[h], Gd[h ? P_S2_W : P_S1_W], Gd[h ? P_S2_B : P_S1_B], s));
        CD c(1, 128, p->dlat2[k]);
        c.term(p->dlat2[k], 32 * h, 0, 128, 3, p->s12[h].pk_d).out(p->dlat[k], 32 * h, 0);
        RUN(c.run(s));
      }
    lb.direct = 1;
    for (int k = 0; k < 3; ++k) lb.dlat[k] = p->dlat[k];
  }
  for (int k = 0; k < 3; ++k) {
    lb.du0[k] = p->du0[k]; lb.lat[k] = p->lat[k];
    lb.du0h[k] = (f16 && p->dec_f16 && !v2 && douts[k]) ? p->du0_h[k] : nullptr;
  }
  lb.z1 = p->z1; lb.z2o = p->z2o; lb.rois = p->rois_in; lb.q = p->q; lb.q_stride = 256; lb.G = G; lb.c1 = p->c1; lb.c2 = p->c2;
  lb.gz1 = p->GA[0]; lb.gz2o = p->gz2o; lb.dq = p->dq;
  lb.gz1_h = f16 ? p->GA_h[0] : nullptr; lb.s16 = ls;
  lb.skip_gz1_32 = f16 ? 1 : 0;   // the fp16 backward of z1_conv reads the fp16 copy only
  const bool z2h = f16 && p->z2_f16;
  lb.gz2o_h = z2h ? p->gz2o_h : nullptr;
  RUN(latent_bwd(lb, s));
  if (Gd[P_MLP2_W]) RUN(angular_bwd(p->query_in, p->dq, Gd[P_MLP2_W], Gd[P_MLP2_B], B, 256, s));

  // ---- z1 branch: gz1 = GA0, gh = GA1, result into the z1 half of g_w = GA2
  {
    BlockBwd bb{{p->w, 0, 32, p->h1, p->z1, &p->z1c[0], &p->z1c[1], &p->z1c[2], nullptr, G}, p->GA[0], p->GA[1],
                Gd[P_Z1 + 0], Gd[P_Z1 + 1], Gd[P_Z1 + 2], Gd[P_Z1 + 3]};
    bb.io.hbits = p->b_h1;
    CD fin(G, 64, p->w);
    fin.out(p->GA[2], 0, 32).mask(p->w, 0, 32, 1, 1.f).mbits(p->b_w).round();
    if (f16) {
      bb.io.x16 = p->w_h; bb.io.h16 = p->h1_h; bb.gy16 = p->GA_h[0]; bb.gh16 = p->GA_h[1]; bb.lscale = ls;
      fin.y16s(p->GA_h[2], ls);
      fin.d.y = nullptr;   // g_w is read as fp16 only (w_conv's data / weight gradients, its residual operand)
    }
    RUN(block_bwd(bb, dp, fin, s));
  }
  // ---- z2 branch
  {
    BlockBwd bb{{p->t21, 0, 16, p->h22, p->z2o, &p->z2b[0], &p->z2b[1], &p->z2b[2], nullptr, 7 * G}, p->gz2o, p->gh22,
                Gd[P_Z2B + 0], Gd[P_Z2B + 1], Gd[P_Z2B + 2], Gd[P_Z2B + 3]};
    CD fin(7 * G, 64, p->t21);
    fin.out(p->dt21, 0, 16).round();
    if (z2h) {
      bb.io.x16 = p->t21_h; bb.io.h16 = p->h22_h; bb.io.hbits = p->b_h22;
      bb.gy16 = p->gz2o_h; bb.gh16 = p->gh22_h; bb.lscale = ls;
      fin.d.round_tf32 = 0;
      fin.y16s(p->dt21_h, ls);
      fin.d.y = nullptr;
    }
    RUN(block_bwd(bb, dp, fin, s));
  }
  if (z2h) RUN(deinterleave2_h(p->dt21_h, p->dt21, p->dte_h, p->dto_h, p->dte, s));
  else RUN(deinterleave2(p->dt21, p->dte, p->dto, s));
  {
    const T4* dts[2] = {&p->dte, &p->dto};
    const void* dts16[2] = {p->dte_h, p->dto_h};
    for (int t = 0; t < 2; ++t) {
      float* dw = Gd[P_CT_W] ? Gd[P_CT_W] + t : nullptr;
      if (dw) {
        NefWgradDesc d;
        memset(&d, 0, sizeof(d));
        d.dy = reinterpret_cast<const float*>(dts[t]->p); d.dy_cstride = dts[t]->cs; d.dy_c4_off = 0; d.dy_c4_gstride = 16;
        d.x = reinterpret_cast<const float*>(p->y20.p); d.x_cstride = p->y20.cs; d.x_c4_off = 0; d.x_c4_gstride = 32;
        d.cout_g = 64; d.cin_g = 128; d.groups = 7 * G; d.taps = 1; d.tap_off = 0; d.rows = p->y20.cs;
        d.dw = dw; d.sg = 128LL * 64 * 2; d.sm = 2; d.sn = 64 * 2; d.st = 0; d.db = Gd[P_CT_B]; d.wg_mod = v2 ? 7 : 0;
        if (z2h) {
          RUN(nef_gconv_wgrad_f16(&d, dts16[t], p->y20_h, ls + 1, sv));
          if (d.db) RUN(nef_bias_grad_h(&d, dts16[t], ls + 1, sv));
        } else {
          RUN(nef_gconv_wgrad(&d, sv));
        }
      }
    }
    CD c(7 * G, 128, p->y20);
    if (z2h) {
      c.term16(p->dte_h, p->dte.cs, 0, 8, 64, 1, p->ct_dh[0]).term16(p->dto_h, p->dto.cs, 0, 8, 64, 1, p->ct_dh[1]);
      c.d.acc_scale = ls + 1;
    } else {
      c.term(p->dte, 0, 16, 64, 1, p->ct_d[0]).term(p->dto, 0, 16, 64, 1, p->ct_d[1]);
    }
    c.d.term[0].tap_off = 0; c.d.term[1].tap_off = 0;
    c.out(p->gy20, 0, 32).mask(p->y20, 0, 32, 1, 1.f).round();
    if (z2h) {
      c.mbits(p->b_y20).y16s(p->gy20_h, ls);
      c.d.y = nullptr;
    }
    RUN(c.run(s));
  }
  {
    BlockBwd bb{{p->ra, 0, 32, p->h20, p->y20, &p->z2a[0], &p->z2a[1], nullptr, nullptr, 7 * G}, p->gy20, p->gh20,
                Gd[P_Z2A + 0], Gd[P_Z2A + 1], nullptr, nullptr};
    CD fin(7 * G, 128, p->ra);
    fin.out(p->dra, 0, 32);
    if (z2h) {
      bb.io.x16 = p->ra_h; bb.io.h16 = p->h20_h; bb.io.hbits = p->b_h20;
      bb.gy16 = p->gy20_h; bb.gh16 = p->gh20_h; bb.lscale = ls;
    }
    RUN(block_bwd(bb, dp, fin, s));
  }
  RUN(roi_align_bwd(p->dra, p->rois_in, p->z2c, p->gz2c, p->win, p->L4, s));
  {
    BlockBwd bb{{p->xw, 0, 16, p->hz, p->z2c, &p->z2c1[0], &p->z2c1[1], &p->z2c1[2], nullptr, G}, p->gz2c, p->ghz,
                Gd[P_Z2C1 + 0], Gd[P_Z2C1 + 1], Gd[P_Z2C1 + 2], Gd[P_Z2C1 + 3]};
    CD fin(G, 64, p->xw);
    fin.out(p->gxw, 0, 16).mask(p->xw, 0, 16, 1, 1.f).round();
    RUN(block_bwd(bb, dp, fin, s));
  }
  RUN(window_scatter(p->gxw, p->GA[2], G, p->win, f16 ? p->GA_h[2] : nullptr, ls, f16 ? 0 : 1, s));
  if (a->ev_late_params_done) {   // every parameter gradient from z1_conv.0 on is final: the caller may start reducing them
    cudaError_t e = cudaEventRecord((cudaEvent_t)a->ev_late_params_done, s);
    NEF_REQUIRE(e == cudaSuccess, "nef_backward: cudaEventRecord failed: %s", cudaGetErrorString(e));
  }
  // ---- w_conv: g_w = GA2, gh = GA0, result (grad of the unscaled last encoder output, pre-ReLU) = GA1
  {
    BlockBwd bb{{p->ey[2], 0, 32, p->hw, p->w, &p->wc[0], &p->wc[1], nullptr, nullptr, G}, p->GA[2], p->GA[0],
                Gd[P_WCONV + 0], Gd[P_WCONV + 1], nullptr, nullptr};
    bb.io.hbits = p->b_hw;
    CD fin(G, 128, p->ey[2]);
    fin.out(p->GA[1], 0, 32).bscale(p->s_in).mask(p->ey[2], 0, 32, 2, 1.f).mbits(p->b_ey[2]).round();
    if (f16) {
      bb.io.x16 = p->ey_h[2]; bb.io.h16 = p->hw_h; bb.gy16 = p->GA_h[2]; bb.gh16 = p->GA_h[0]; bb.lscale = ls;
      fin.y16s(p->GA_h[1], ls);
      fin.d.y = nullptr;   // read as fp16 only: by bscale_grad_h below and by the last encoder block's backward
    }
    RUN(block_bwd(bb, dp, fin, s));
    if (f16) RUN(bscale_grad_h(p->GA[1], p->GA_h[1], p->ey_h[2], ls + 1, p->s_in, p->ds_in, s));
    else RUN(bscale_grad(p->GA[1], p->ey[2], p->s_in, p->ds_in, s));
  }
  if (Gd[P_MLP1_W]) RUN(angular_bwd(p->thetas_in, p->ds_in, Gd[P_MLP1_W], Gd[P_MLP1_B], B * G, 128, s));
  // ---- encoder blocks 2, 1, 0
  {
    int gy_i = 1;  // GA index holding gy
    for (int i = 2; i >= 0; --i) {
      const int gx_i = gy_i == 1 ? 2 : 1;
      BlockBwd bb{{i == 0 ? p->s0 : p->ey[i - 1], 0, 32, p->eh[i], p->ey[i], &p->enc[2 * i], &p->enc[2 * i + 1], nullptr,
                   nullptr, G},
                  p->GA[gy_i], p->GA[0], Gd[P_ENC + 2 * i], Gd[P_ENC + 2 * i + 1], nullptr, nullptr};
      bb.io.hbits = p->b_eh[i];
      CD fin(G, 128, p->s0);
      fin.out(p->GA[gx_i], 0, 32);
      if (i > 0) fin.mask(p->ey[i - 1], 0, 32, 1, 1.f).mbits(p->b_ey[i - 1]).round();
      if (f16) {
        bb.io.x16 = i == 0 ? p->s0_h : p->ey_h[i - 1]; bb.io.h16 = p->eh_h[i];
        bb.gy16 = p->GA_h[gy_i]; bb.gh16 = p->GA_h[0]; bb.lscale = ls;
        if (i > 0 || g_stem_tc) {
          fin.y16s(p->GA_h[gx_i], ls);
          fin.d.y = nullptr;   // the next block (the tensor-core stem gradient) reads this gradient as fp16 only
        }
      }
      RUN(block_bwd(bb, dp, fin, s));
      gy_i = gx_i;
    }
    if (Gd[P_STEM] && f16 && g_stem_tc)
      RUN(stem_tc_bwd(p->x_in, p->s0_amax, p->GA_h[gy_i], p->GA[gy_i], Gd[P_STEM], ls + 1, G, s, v2 ? 1 : 0));
    else if (Gd[P_STEM]) RUN(stem_bwd(p->x_in, p->s0_amax, p->GA[gy_i], Gd[P_STEM], G, s, v2 ? 1 : 0));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// workspace inspection (test hook): named internal tensors of the last forward
// ---------------------------------------------------------------------------------------------
struct NamedTensor { const T4* t; const float* v; int n; const void* h16; const uint32_t* bits; const uint32_t* codes; };
static bool find_tensor(const NefPlan* p, const char* name, NamedTensor* out) {
  out->t = nullptr; out->v = nullptr; out->n = 0; out->h16 = nullptr; out->bits = nullptr; out->codes = nullptr;
  std::string n(name);
  // the stem's max-pool selections: 0..2 = which conv position of the window (2j-1, 2j, 2j+1) won, 3 = clipped by the ReLU
  if (n == "stem.argmax") { out->t = &p->s0; out->codes = p->s0_amax; return true; }
  // "<block>.h.mask" / "<block>.y.mask": the one-bit (value != 0) plane the masked data-gradient epilogues read, as 0 / 1 floats
  if (n.size() > 5 && n.substr(n.size() - 5) == ".mask") {
    n = n.substr(0, n.size() - 5);
    struct { const char* name; const T4* t; const uint32_t* b; } planes[] = {
        {"W_encoder.layer1.0.h", &p->eh[0], p->b_eh[0]}, {"W_encoder.layer1.1.h", &p->eh[1], p->b_eh[1]},
        {"W_encoder.layer1.2.h", &p->eh[2], p->b_eh[2]}, {"W_encoder.layer1.0.y", &p->ey[0], p->b_ey[0]},
        {"W_encoder.layer1.1.y", &p->ey[1], p->b_ey[1]}, {"W_encoder.layer1.2.y", &p->ey[2], p->b_ey[2]},
        {"w_conv.0.h", &p->hw, p->b_hw}, {"w_conv.0.y", &p->w, p->b_w}, {"z1_conv.0.h", &p->h1, p->b_h1},
        {"z2_conv2.0.h", &p->h20, p->b_h20}, {"z2_conv2.0.y", &p->y20, p->b_y20}, {"z2_conv2.2.h", &p->h22, p->b_h22}};
    for (auto& q : planes)
      if (n == q.name) { out->t = q.t; out->bits = q.b; return true; }
    return false;
  }
  auto is = [&](const char* s) { return n == s; };
  auto blk = [&](const char* prefix, const T4& h, const T4& y) {
    const std::string pre(prefix);
    if (n == pre + ".h") { out->t = &h; return true; }
    if (n == pre + ".y") { out->t = &y; return true; }
    return false;
  };
  if (is("stem")) { out->t = &p->s0; if (p->h_f16_only) out->h16 = p->s0_h; return true; }
  // hidden activations that the last forward kept as fp16 copies only
  auto h_only = [&](const void* h16) { if (p->h_f16_only && out->t && n.size() > 2 && n.substr(n.size() - 2) == ".h") out->h16 = h16; return true; };
  for (int i = 0; i < 3; ++i)
    if (blk(("W_encoder.layer1." + std::to_string(i)).c_str(), p->eh[i], p->ey[i])) {
      if (p->h_f16_only && (i < 2 || !g_k3_tf32) && out->t == &p->ey[i]) { out->h16 = p->ey_h[i]; return true; }
      return h_only(p->eh_h[i]);
    }
  if (blk("w_conv.0", p->hw, p->w)) {
    if (p->h_f16_only && p->fwd_f16 && !g_k3_tf32 && out->t == &p->w) { out->h16 = p->w_h; return true; }
    return h_only(p->hw_h);
  }
  if (blk("z1_conv.0", p->h1, p->z1)) return h_only(p->h1_h);
  if (blk("z2_conv1.0", p->hz, p->z2c)) return true;
  if (blk("z2_conv2.0", p->h20, p->y20)) { if (p->z2_f16) out->h16 = out->t == &p->h20 ? p->h20_h : p->y20_h; return true; }
  if (blk("z2_conv2.2", p->h22, p->z2o)) { if (p->z2_f16 && out->t == &p->h22) out->h16 = p->h22_h; return true; }
  if (is("roi_align")) { out->t = &p->ra; if (p->z2_f16) out->h16 = p->ra_h; return true; }
  if (is("z2_conv2.1")) { out->t = &p->t21; if (p->z2_f16) out->h16 = p->t21_h; return true; }
  for (int k = 0; k < 3; ++k) {
    const DecBufs& d = p->dec[k];
    const std::string pre = "dec" + std::to_string(k) + ".";
    const T4* ts[7] = {&d.c1, &d.a1, &d.c2, &d.u1, &d.c3, &d.a3, &d.c4};
    const char* nm[7] = {"decoder.1.0", "a1", "decoder.1.3", "u1", "decoder.3.0", "a3", "decoder.3.3"};
    const void* hs[7] = {nullptr, d.a1_h, nullptr, d.u1_h, nullptr, d.a3_h, nullptr};
    for (int i = 0; i < 7; ++i)
      if (n == pre + nm[i]) { out->t = ts[i]; if (p->dec_f16 || (p->folded && p->fwd_f16 && g_dec_f16)) out->h16 = hs[i]; return true; }
    if (n == pre + "u0") { out->t = &p->u0[k]; if (p->u0_f16_only) out->h16 = p->u0_h[k]; return true; }
    for (int i = 0; i < 4; ++i) {
      const std::string b = pre + "bn" + std::to_string(i) + ".";
      const int ch = i < 2 ? 128 : 64;
      if (n == b + "scale") { out->v = d.bn[i].scale; out->n = ch; return true; }
      if (n == b + "shift") { out->v = d.bn[i].shift; out->n = ch; return true; }
    }
  }
  return false;
}
extern "C" int nef_plan_tensor_info(const NefPlan* p, const char* name, int* C, int* L) {
  NamedTensor t;
  NEF_REQUIRE(p && name && find_tensor(p, name, &t), "nef_plan_tensor_info: unknown tensor '%s'", name ? name : "");
  if (t.t) { *C = t.t->C; *L = t.t->L; }
  else { *C = t.n; *L = 0; }
  return 0;
}
// Test hook: overwrite the fp32 storage of a named activation (valid rows only; the zero halo is part of the layout) with
// NaN, so that a test can prove that nothing reads a tensor the dataflow claims to have dropped.
extern "C" int nef_plan_poison(NefPlan* p, const char* name, float* scratch, nef_stream_t s) {
  NamedTensor t;
  NEF_REQUIRE(p && p->bound && name && scratch && find_tensor(p, name, &t) && t.t && !t.bits && !t.codes,
              "nef_plan_poison: unknown tensor '%s'", name ? name : "");
  cudaMemsetAsync(scratch, 0xFF, (size_t)p->B * t.t->C * t.t->L * sizeof(float), (cudaStream_t)s);   // all-ones bits = NaN
  return nef_ncl_to_cbl4(scratch, reinterpret_cast<float*>(t.t->p), p->B, t.t->C, t.t->L, 0, s);
}
extern "C" int nef_plan_export(NefPlan* p, const char* name, float* dst, nef_stream_t s) {
  NamedTensor t;
  NEF_REQUIRE(p && p->bound && name && find_tensor(p, name, &t), "nef_plan_export: unknown tensor '%s'", name ? name : "");
  if (t.t && t.codes) return nef_codes_to_ncl(t.codes, dst, p->B, t.t->C, t.t->L, s);
  if (t.t && t.bits) return nef_bits_to_ncl(t.bits, dst, p->B, t.t->C, t.t->L, s);
  if (t.t && t.h16) return nef_h8_to_ncl(t.h16, dst, p->B, t.t->C, t.t->L, s);
  if (t.t) return nef_cbl4_to_ncl(reinterpret_cast<const float*>(t.t->p), dst, p->B, t.t->C, t.t->L, s);
  cudaError_t e = cudaMemcpyAsync(dst, t.v, (size_t)t.n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)s);
  NEF_REQUIRE(e == cudaSuccess, "nef_plan_export: copy failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// single-op exports for unit tests
// ---------------------------------------------------------------------------------------------
static T4 view_t4(const float* ptr, int C, int B, int L) {
  T4 t;
  t.p = reinterpret_cast<float4*>(const_cast<float*>(ptr));
  t.C = C; t.B = B; t.L = L; t.Lp = L + 2 * NEF_HALO; t.cs = (long)B * t.Lp;
  return t;
}
extern "C" int nef_roi_check(const int64_t* rois, int B, int L, int32_t* flag, nef_stream_t s) {
  NEF_REQUIRE(rois && flag && B >= 1 && L % 4 == 0, "nef_roi_check: bad arguments");
  return roi_check(rois, B, L / 4, flag, (cudaStream_t)s);
}
extern "C" int nef_stem_fwd(const float* x, const float* w, float* y, uint32_t* argmax, int B, int G, int L,
                            nef_stream_t s) {
  return stem_fwd(x, w, view_t4(y, 128 * G, B, L / 4), argmax, nullptr, G, (cudaStream_t)s);
}
extern "C" int nef_stem_tc_fwd(const float* x, const float* w, void* y16, uint32_t* argmax, int B, int G, int L, nef_stream_t s) {
  NEF_REQUIRE(x && w && y16 && L % 4 == 0, "nef_stem_tc_fwd: bad arguments");
  return stem_tc_fwd(x, w, view_t4(nullptr, 128 * G, B, L / 4), argmax, y16, G, (cudaStream_t)s);
}
extern "C" int nef_stem_tc_bwd(const float* x, const uint32_t* argmax, const void* dy16, float* dw, const float* inv_scale, int B, int G,
                               int L, nef_stream_t s) {
  NEF_REQUIRE(x && argmax && dy16 && dw && L % 4 == 0, "nef_stem_tc_bwd: bad arguments");
  return stem_tc_bwd(x, argmax, dy16, view_t4(nullptr, 128 * G, B, L / 4), dw, inv_scale, G, (cudaStream_t)s);
}
extern "C" int nef_stem_bwd(const float* x, const uint32_t* argmax, const float* dy, float* dw, int B, int G, int L,
                            nef_stream_t s) {
  NEF_REQUIRE(argmax, "nef_stem_bwd: the argmax codes written by nef_stem_fwd are required");
  return stem_bwd(x, argmax, view_t4(dy, 128 * G, B, L / 4), dw, G, (cudaStream_t)s);
}
extern "C" int nef_angular_fwd(const float* theta, const float* w, const float* b, float* out, int n, int D,
                               nef_stream_t s) {
  return angular_fwd(theta, w, b, out, n, D, (cudaStream_t)s);
}
extern "C" int nef_angular_bwd(const float* theta, const float* dout, float* dw, float* db, int n, int D,
                               nef_stream_t s) {
  return angular_bwd(theta, dout, dw, db, n, D, (cudaStream_t)s);
}
