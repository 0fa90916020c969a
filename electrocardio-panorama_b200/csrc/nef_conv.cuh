// Epilogue shared by the CUDA-core and the tcgen05 implicit-GEMM convolution kernels.
#pragma once
#include "nef_common.cuh"
#include "../../include/nefnet_b200.h"

namespace nef {

struct EpiRow {
  bool valid;
  int b;
  long out_row;
};

__device__ __forceinline__ EpiRow epi_row(const NefConvDesc& d, long row) {
  EpiRow r;
  int b = int(row / d.Lp);
  int l = int(row - (long)b * d.Lp) - NEF_HALO;
  r.valid = (row < d.rows) && (l >= 0) && (l < d.L);
  r.b = b;
  r.out_row = (long)b * d.y_Lp + NEF_HALO + (long)l * d.y_lmul + d.y_ladd;
  return r;
}

// One float4 = 4 consecutive output channels (chunk n4 of group g) of one row.
//   *pre  <- v after bias/residual (what BatchNorm statistics are taken over)
//   *bsg  <- contribution to bscale_grad (0 unless d.bscale_grad)
// Stores the finished value.  Must only be called for rows with r.valid.
__device__ __forceinline__ void epi_apply_store(const NefConvDesc& d, const EpiRow& r, int g, int n4, float4 v,
                                                float4* pre, float4* bsg) {
  const int ch4 = g * (d.N >> 2) + n4;  // chunk index in the [groups*N] channel space
  if (d.bias) v = v + __ldg(reinterpret_cast<const float4*>(d.bias) + ch4);
  if (d.res) {
    const float4* rp = reinterpret_cast<const float4*>(d.res) +
                       (long)(d.res_c4_off + g * d.res_c4_gstride + n4) * d.res_cstride + r.out_row;
    v = v + __ldg(rp);
  }
  *pre = v;
  if (d.relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
  if (d.drop_p > 0.f) {
    uint64_t bits = drop_bits(d.drop_seed, r.out_row, d.y_c4_off + g * d.y_c4_gstride + n4);
    const uint32_t thr = (uint32_t)(d.drop_p * 65536.f);
    const float sc = 1.f / (1.f - d.drop_p);
    v.x = ((bits & 0xffff) >= thr) ? v.x * sc : 0.f;
    v.y = (((bits >> 16) & 0xffff) >= thr) ? v.y * sc : 0.f;
    v.z = (((bits >> 32) & 0xffff) >= thr) ? v.z * sc : 0.f;
    v.w = (((bits >> 48) & 0xffff) >= thr) ? v.w * sc : 0.f;
  }
  const float4 v0 = v;
  float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (d.bscale) {
    sc4 = __ldg(reinterpret_cast<const float4*>(d.bscale) + (long)r.b * (d.groups * (d.N >> 2)) + ch4);
    v = v * sc4;
  }
  float4 m = f4zero();
  if (d.mask_mode) {
    const float4* mp = reinterpret_cast<const float4*>(d.mask) +
                       (long)(d.mask_c4_off + g * d.mask_c4_gstride + n4) * d.mask_cstride + r.out_row;
    m = __ldg(mp);
  }
  *bsg = f4zero();
  if (d.bscale_grad) {
    // forward was ys = relu(.) * s and the mask tensor is ys; here v0 = d ys:
    //   d s += v0 * relu(.) = v0 * ys / s     (d relu(.) = v0 * s * (ys != 0) is the generic path below)
    bsg->x = sc4.x != 0.f ? v0.x * m.x / sc4.x : 0.f;
    bsg->y = sc4.y != 0.f ? v0.y * m.y / sc4.y : 0.f;
    bsg->z = sc4.z != 0.f ? v0.z * m.z / sc4.z : 0.f;
    bsg->w = sc4.w != 0.f ? v0.w * m.w / sc4.w : 0.f;
  }
  if (d.mask_mode == 1) {
    v.x = m.x > 0.f ? v.x * d.mask_scale : 0.f;
    v.y = m.y > 0.f ? v.y * d.mask_scale : 0.f;
    v.z = m.z > 0.f ? v.z * d.mask_scale : 0.f;
    v.w = m.w > 0.f ? v.w * d.mask_scale : 0.f;
  } else if (d.mask_mode == 2) {
    v.x = m.x != 0.f ? v.x * d.mask_scale : 0.f;
    v.y = m.y != 0.f ? v.y * d.mask_scale : 0.f;
    v.z = m.z != 0.f ? v.z * d.mask_scale : 0.f;
    v.w = m.w != 0.f ? v.w * d.mask_scale : 0.f;
  }
  if (d.round_tf32) v = tf32_rn4(v);
  float4* yp = reinterpret_cast<float4*>(d.y) + (long)(d.y_c4_off + g * d.y_c4_gstride + n4) * d.y_cstride + r.out_row;
  *yp = v;
}

}  // namespace nef
