// CUDA-core fp32 implementation of the grouped implicit-GEMM convolution and its weight gradient.
// This is the bring-up / cross-check implementation of the contraction (nef_set_conv_impl(0));
// the production path is the tcgen05 TF32 kernel in nef_conv_tc.cu.  Same descriptor, same epilogue.
#include <cuda_fp16.h>
#include "nef_conv.cuh"

namespace nef {

constexpr int ST_ROWS = 128;  // rows per block tile
constexpr int ST_N = 64;      // output channels per block tile
constexpr int ST_XROWS = ST_ROWS + 6;

// block 256 threads: warp w -> rows (w & 3) * 32 + lane, channel half (w >> 2) * 32
__global__ void __launch_bounds__(256, 2) conv_simt_kernel(const NefConvDesc d) {
  __shared__ float4 Xs[8][ST_XROWS];
  __shared__ float4 Ws[8][ST_N];
  __shared__ float s_stat[2][4][ST_N];  // [sum | sq][row quarter][channel]: one writer each (deterministic)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = d.N / ST_N;
  const int g = blockIdx.y / ntiles, nt = blockIdx.y % ntiles;
  const long r0 = (long)blockIdx.x * ST_ROWS;
  const int rw = (warp & 3) * 32 + lane;
  const int half = warp >> 2;

  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = f4zero();

  for (int ti = 0; ti < d.n_terms; ++ti) {
    const NefConvTerm& t = d.term[ti];
    const int nkb = t.cin_g >> 5;
    const float4* xg = reinterpret_cast<const float4*>(t.x);
    const float4* wg = reinterpret_cast<const float4*>(t.w);
    const int xrows = ST_ROWS + t.taps - 1;
    for (int kb = 0; kb < nkb; ++kb) {
      __syncthreads();
      for (int i = tid; i < 8 * xrows; i += 256) {
        int c = i / xrows, rr = i - c * xrows;
        long chunk = t.x_c4_off + (long)g * t.x_c4_gstride + kb * 8 + c;
        Xs[c][rr] = __ldg(xg + chunk * t.x_cstride + (r0 + t.tap_off + rr));
      }
      for (int tp = 0; tp < t.taps; ++tp) {
        if (tp > 0) __syncthreads();
        const float4* wt = wg + ((((long)g * t.taps + tp) * nkb + kb) * 8) * d.N + nt * ST_N;
        for (int i = tid; i < 8 * ST_N; i += 256) {
          int c = i / ST_N, n = i - c * ST_N;
          Ws[c][n] = __ldg(wt + (long)c * d.N + n);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 xv = Xs[c][rw + tp];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 w0 = Ws[c][half * 32 + i * 4 + 0];
            const float4 w1 = Ws[c][half * 32 + i * 4 + 1];
            const float4 w2 = Ws[c][half * 32 + i * 4 + 2];
            const float4 w3 = Ws[c][half * 32 + i * 4 + 3];
            acc[i].x += xv.x * w0.x + xv.y * w0.y + xv.z * w0.z + xv.w * w0.w;
            acc[i].y += xv.x * w1.x + xv.y * w1.y + xv.z * w1.z + xv.w * w1.w;
            acc[i].z += xv.x * w2.x + xv.y * w2.y + xv.z * w2.z + xv.w * w2.w;
            acc[i].w += xv.x * w3.x + xv.y * w3.y + xv.z * w3.z + xv.w * w3.w;
          }
        }
      }
    }
  }

  // ---- epilogue
  const bool want_stats = d.stat_sum != nullptr;
  const EpiRow er = epi_row(d, r0 + rw);
  // is the whole warp inside one segment?  (for the bscale_grad reduction)
  const int b0 = __shfl_sync(0xffffffffu, er.b, 0);
  const bool uniform_b = __all_sync(0xffffffffu, er.b == b0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n4 = nt * (ST_N / 4) + half * 8 + i;
    float4 pre = f4zero(), bsg = f4zero();
    if (er.valid) epi_apply_store(d, er, g, n4, acc[i], &pre, &bsg);
    if (want_stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = er.valid ? f4get(pre, j) : 0.f;
        float s1 = warp_sum(v), s2 = warp_sum(v * v);
        if (lane == 0) {
          s_stat[0][warp & 3][half * 32 + i * 4 + j] = s1;
          s_stat[1][warp & 3][half * 32 + i * 4 + j] = s2;
        }
      }
    }
    if (d.bscale_grad) {
      const long cbase = (long)(g * d.N + n4 * 4);
      const long ctot = (long)d.groups * d.N;
      if (uniform_b) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s1 = warp_sum(er.valid ? f4get(bsg, j) : 0.f);
          if (lane == 0 && s1 != 0.f) atomicAdd(d.bscale_grad + (long)b0 * ctot + cbase + j, s1);
        }
      } else if (er.valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(d.bscale_grad + (long)er.b * ctot + cbase + j, f4get(bsg, j));
      }
    }
  }
  if (want_stats) {
    __syncthreads();
    for (int i = tid; i < ST_N; i += 256) {
      const long o = (long)blockIdx.x * ((long)d.groups * d.N) + (long)g * d.N + nt * ST_N + i;
      d.stat_sum[o] = (s_stat[0][0][i] + s_stat[0][1][i]) + (s_stat[0][2][i] + s_stat[0][3][i]);
      d.stat_sq[o] = (s_stat[1][0][i] + s_stat[1][1][i]) + (s_stat[1][2][i] + s_stat[1][3][i]);
    }
  }
}

// ---- weight gradient ------------------------------------------------------------------------
constexpr int WG_TR = 64;        // rows per smem stage
constexpr int WG_XS = WG_TR + 3; // odd chunk pitch -> conflict-free

template <int TP>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const NefWgradDesc d, long rows_per_split, long row0) {
  __shared__ float4 Ys[16][WG_TR];
  __shared__ float4 Xs[16][WG_XS];
  const int tid = threadIdx.x;
  const int co4 = tid >> 4, ci4 = tid & 15;
  const int mt = d.cout_g / 64, nt = d.cin_g / 64;
  int idx = blockIdx.y;
  const int cit = idx % nt; idx /= nt;
  const int cot = idx % mt; idx /= mt;
  const int g = idx;
  const int tap_base = blockIdx.z * 4;
  const int ntap = min(TP, d.taps - tap_base);
  const long rbeg = row0 + (long)blockIdx.x * rows_per_split;
  const long rend = min(d.rows, rbeg + rows_per_split);

  float acc[TP][4][4];
#pragma unroll
  for (int a = 0; a < TP; ++a)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.f;
  float4 bacc = f4zero();

  const float4* yg = reinterpret_cast<const float4*>(d.dy) + (long)(d.dy_c4_off + g * d.dy_c4_gstride + cot * 16) * d.dy_cstride;
  const float4* xg = reinterpret_cast<const float4*>(d.x) + (long)(d.x_c4_off + g * d.x_c4_gstride + cit * 16) * d.x_cstride;

  for (long r = rbeg; r < rend; r += WG_TR) {
    __syncthreads();
    for (int i = tid; i < 16 * WG_TR; i += 256) {
      int c = i / WG_TR, rr = i - c * WG_TR;
      Ys[c][rr] = (r + rr < rend) ? __ldg(yg + (long)c * d.dy_cstride + r + rr) : f4zero();
    }
    for (int i = tid; i < 16 * (WG_TR + TP - 1); i += 256) {
      int c = i / (WG_TR + TP - 1), rr = i - c * (WG_TR + TP - 1);
      Xs[c][rr] = __ldg(xg + (long)c * d.x_cstride + r + rr + d.tap_off + tap_base);
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < WG_TR; ++rr) {
      const float4 a = Ys[co4][rr];
      bacc = bacc + a;
#pragma unroll
      for (int tt = 0; tt < TP; ++tt) {
        const float4 b = Xs[ci4][rr + tt];
        acc[tt][0][0] += a.x * b.x; acc[tt][0][1] += a.x * b.y; acc[tt][0][2] += a.x * b.z; acc[tt][0][3] += a.x * b.w;
        acc[tt][1][0] += a.y * b.x; acc[tt][1][1] += a.y * b.y; acc[tt][1][2] += a.y * b.z; acc[tt][1][3] += a.y * b.w;
        acc[tt][2][0] += a.z * b.x; acc[tt][2][1] += a.z * b.y; acc[tt][2][2] += a.z * b.z; acc[tt][2][3] += a.z * b.w;
        acc[tt][3][0] += a.w * b.x; acc[tt][3][1] += a.w * b.y; acc[tt][3][2] += a.w * b.z; acc[tt][3][3] += a.w * b.w;
      }
    }
  }
#pragma unroll
  for (int tt = 0; tt < TP; ++tt) {
    if (tt < ntap) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long m = cot * 64 + co4 * 4 + i, n = cit * 64 + ci4 * 4 + j;
          atomicAdd(d.dw + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.sg + m * d.sm + n * d.sn + (tap_base + tt) * d.st, acc[tt][i][j]);
        }
    }
  }
  if (d.db && cit == 0 && blockIdx.z == 0 && ci4 == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(d.db + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.cout_g + cot * 64 + co4 * 4 + i, f4get(bacc, i));
  }
}

// Several packing jobs in one launch (the plan packs ~28 weight tensors per pass): the job table travels as a kernel
// parameter, a block finds its job from the cumulative block counts and packs NEF_PACK_CHUNK elements of it.
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const __grid_constant__ NefPackTable tab) {
  int jb = 0;
  while (jb + 1 < tab.n && (int)blockIdx.x >= tab.job[jb + 1].first_block) ++jb;
  const NefPackJob& q = tab.job[jb];
  const bool f16 = (q.flags & 4) != 0;  // fp16 operand packing: a 16-byte slot holds 8 consecutive input channels
  const long total = (long)q.groups * q.taps * (f16 ? q.K >> 1 : q.K) * q.N;   // in 4-byte units of the destination
  const long beg = (long)((int)blockIdx.x - q.first_block) * NEF_PACK_CHUNK;
  const long end = beg + NEF_PACK_CHUNK < total ? beg + NEF_PACK_CHUNK : total;
  if (f16) {
    const int nkb16 = q.K >> 6;
    for (long i4 = (beg >> 2) + threadIdx.x; i4 < (end >> 2); i4 += 256) {
      long r = i4;
      const int n = r % q.N; r /= q.N;
      const int c = r & 7; r >>= 3;
      const int kb = r % nkb16; r /= nkb16;
      const int t = r % q.taps; r /= q.taps;
      const int g = (int)r;
      const int k = kb * 64 + c * 8;
      const int ts = (q.flags & 1) ? q.taps - 1 - t : t;
      const float* sp = q.src + (q.gmod > 0 ? g % q.gmod : g) * q.sg + n * q.sn + k * q.sk + ts * q.st;
      const float sc = q.nscale ? q.nscale[(q.gmod > 0 ? g % q.gmod : g) * q.N + n] : 1.0f;
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo = sp[(2 * j) * q.sk] * sc, hi = sp[(2 * j + 1) * q.sk] * sc;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(hi), "f"(lo));
        if (q.flags & 2) {  // residual w - fp16(w), itself in fp16
          const __half2 h2 = *reinterpret_cast<const __half2*>(&o[j]);
          lo -= __low2float(h2);
          hi -= __high2float(h2);
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(hi), "f"(lo));
        }
      }
      reinterpret_cast<uint4*>(q.dst)[i4] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  const int nkb = q.K >> 5;
  // one thread per float4 of the packed layout (4 consecutive input channels of one output channel): the index
  // arithmetic is paid once per four elements
  for (long i4 = (beg >> 2) + threadIdx.x; i4 < (end >> 2); i4 += 256) {
    long r = i4;
    const int n = r % q.N; r /= q.N;
    const int c = r & 7; r >>= 3;
    const int kb = r % nkb; r /= nkb;
    const int t = r % q.taps; r /= q.taps;
    const int g = (int)r;
    const int k = kb * 32 + c * 4;
    const int ts = (q.flags & 1) ? q.taps - 1 - t : t;
    const float* sp = q.src + (q.gmod > 0 ? g % q.gmod : g) * q.sg + n * q.sn + k * q.sk + ts * q.st;
    const float sc = q.nscale ? q.nscale[(q.gmod > 0 ? g % q.gmod : g) * q.N + n] : 1.0f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float wv = sp[j * q.sk] * sc;
      const float hi = tf32_rn(wv);
      o[j] = (q.flags & 2) ? tf32_rn(wv - hi) : hi;
    }
    reinterpret_cast<float4*>(q.dst)[i4] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void ncl_to_cbl4_kernel(const float* __restrict__ src, float4* __restrict__ dst, int B, int C, int L,
                                   int round_tf32) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 4) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int c4 = r / B;
    const float* s = src + ((long)b * C + c4 * 4) * L + l;
    float4 v = make_float4(s[0], s[L], s[2L * L], s[3L * L]);
    if (round_tf32) v = tf32_rn4(v);
    dst[(long)c4 * B * Lp + (long)b * Lp + NEF_HALO + l] = v;
  }
}

__global__ void cbl4_to_ncl_kernel(const float4* __restrict__ src, float* __restrict__ dst, int B, int C, int L) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 4) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int c4 = r / B;
    const float4 v = src[(long)c4 * B * Lp + (long)b * Lp + NEF_HALO + l];
    float* o = dst + ((long)b * C + c4 * 4) * L + l;
    o[0] = v.x; o[L] = v.y; o[2L * L] = v.z; o[3L * L] = v.w;
  }
}

// fp16 copy `half8 [C/8][B * Lp]` (NefConvDesc.y16) -> (B, C, L) fp32
__global__ void h8_to_ncl_kernel(const uint4* __restrict__ src, float* __restrict__ dst, int B, int C, int L) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 8) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int c8 = r / B;
    const uint4 v = src[(long)c8 * B * Lp + (long)b * Lp + NEF_HALO + l];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float* o = dst + ((long)b * C + c8 * 8) * L + l;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
      o[(long)(2 * j) * L] = __low2float(h);
      o[(long)(2 * j + 1) * L] = __high2float(h);
    }
  }
}

// one-bit planes `uint32 [C/32][B * Lp]` (NefConvDesc.out_bits: bit 4 i + j <-> channel 32 plane + 4 i + j) -> (B, C, L) 0 / 1
__global__ void bits_to_ncl_kernel(const uint32_t* __restrict__ src, float* __restrict__ dst, int B, int C, int L) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 32) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int pl = r / B;
    const uint32_t w = src[(long)pl * B * Lp + (long)b * Lp + NEF_HALO + l];
    float* o = dst + ((long)b * C + pl * 32) * L + l;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[(long)j * L] = (w >> j) & 1u ? 1.f : 0.f;
  }
}

// one code byte per channel (a uint32 per 4-channel chunk and row, indexed like a CBL4 tensor) -> floats (B, C, L)
__global__ void codes_to_ncl_kernel(const uint32_t* __restrict__ src, float* __restrict__ dst, int B, int C, int L) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 4) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int c4 = r / B;
    const uint32_t w = src[(long)c4 * B * Lp + (long)b * Lp + NEF_HALO + l];
    float* o = dst + ((long)b * C + c4 * 4) * L + l;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[(long)j * L] = (float)((w >> (8 * j)) & 0xffu);
  }
}

}  // namespace nef

using namespace nef;

extern "C" int nef_codes_to_ncl(const uint32_t* src, float* dst, int B, int C, int L, nef_stream_t s) {
  const long total = (long)B * (C / 4) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  codes_to_ncl_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(src, dst, B, C, L);
  NEF_CHECK_LAUNCH("codes_to_ncl_kernel");
  return 0;
}

extern "C" int nef_bits_to_ncl(const uint32_t* src, float* dst, int B, int C, int L, nef_stream_t s) {
  const long total = (long)B * (C / 32) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bits_to_ncl_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(src, dst, B, C, L);
  NEF_CHECK_LAUNCH("bits_to_ncl_kernel");
  return 0;
}

extern "C" int nef_h8_to_ncl(const void* src, float* dst, int B, int C, int L, nef_stream_t s) {
  const long total = (long)B * (C / 8) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  h8_to_ncl_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(reinterpret_cast<const uint4*>(src), dst, B, C, L);
  NEF_CHECK_LAUNCH("h8_to_ncl_kernel");
  return 0;
}

extern "C" int nef_gconv_fwd_simt(const NefConvDesc* d, nef_stream_t s) {
  dim3 grid((unsigned)((d->rows + ST_ROWS - 1) / ST_ROWS), (unsigned)(d->groups * (d->N / ST_N)));
  conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(*d);
  NEF_CHECK_LAUNCH("conv_simt_kernel");
  return 0;
}

// weight gradient over rows [row0, d->rows)
extern "C" int nef_gconv_wgrad_simt_range(const NefWgradDesc* d, long row0, nef_stream_t s) {
  const long nrows = d->rows - row0;
  if (nrows <= 0) return 0;
  const int passes = (d->taps + 3) / 4;
  const int tiles = d->groups * (d->cout_g / 64) * (d->cin_g / 64);
  long splits = (148L * 4 + (long)tiles * passes - 1) / ((long)tiles * passes);
  const long max_splits = (nrows + WG_TR - 1) / WG_TR;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long rps = (nrows + splits - 1) / splits;
  rps = (rps + WG_TR - 1) / WG_TR * WG_TR;
  splits = (nrows + rps - 1) / rps;
  dim3 grid((unsigned)splits, (unsigned)tiles, (unsigned)passes);
  const int tp = d->taps >= 4 ? 4 : d->taps;
  switch (tp) {
    case 1: wgrad_simt_kernel<1><<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps, row0); break;
    case 2: wgrad_simt_kernel<2><<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps, row0); break;
    case 3: wgrad_simt_kernel<3><<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps, row0); break;
    default: wgrad_simt_kernel<4><<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps, row0); break;
  }
  NEF_CHECK_LAUNCH("wgrad_simt_kernel");
  return 0;
}
extern "C" int nef_gconv_wgrad_simt(const NefWgradDesc* d, nef_stream_t s) { return nef_gconv_wgrad_simt_range(d, 0, s); }

extern "C" int nef_pack_weights(const float* src, float* dst, int groups, int N, int K, int taps, int64_t sg,
                                int64_t sn, int64_t sk, int64_t st, int flags, nef_stream_t s) {
  // one job of the batched packer (the kernel the plan uses), so that every flag -- including bit 2, the fp16 operand
  // packing of NefConvTerm.x_f16 -- is available through the C ABI
  NefPackTable t;
  t.n = 1;
  NefPackJob& q = t.job[0];
  q.src = src; q.dst = dst; q.groups = groups; q.N = N; q.K = K; q.taps = taps;
  q.sg = sg; q.sn = sn; q.sk = sk; q.st = st; q.flags = flags; q.first_block = 0; q.nscale = nullptr; q.gmod = 0;
  return nef_pack_weights_batch(&t, (cudaStream_t)s);
}

// (B, C, L) fp32 -> fp16 copy `half8 [C/8][B * Lp]` (the layout of NefConvDesc.y16), values multiplied by scale, saturating
__global__ void ncl_to_h8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int B, int C, int L, float scale) {
  const int Lp = L + 2 * NEF_HALO;
  const long total = (long)B * (C / 8) * L;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int l = i % L;
    long r = i / L;
    const int b = r % B;
    const int c8 = r / B;
    const float* p = src + ((long)b * C + c8 * 8) * L + l;
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = p[(long)(2 * j) * L] * scale, hi = p[(long)(2 * j + 1) * L] * scale;
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o[j]) : "f"(hi), "f"(lo));
    }
    dst[(long)c8 * B * Lp + (long)b * Lp + NEF_HALO + l] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
extern "C" int nef_ncl_to_h8(const float* src, void* dst, int B, int C, int L, float scale, nef_stream_t s) {
  NEF_REQUIRE(C % 8 == 0, "nef_ncl_to_h8: C %% 8 required");
  const long total = (long)B * (C / 8) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ncl_to_h8_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(src, reinterpret_cast<uint4*>(dst), B, C, L, scale);
  NEF_CHECK_LAUNCH("ncl_to_h8_kernel");
  return 0;
}

int nef_pack_weights_batch(NefPackTable* tab, cudaStream_t s) {
  if (tab->n == 0) return 0;
  int blocks = 0;
  for (int i = 0; i < tab->n; ++i) {
    NefPackJob& q = tab->job[i];
    NEF_REQUIRE(q.K % 32 == 0 && q.N % 4 == 0, "nef_pack_weights_batch: K %% 32 and N %% 4 required (K=%d N=%d)", q.K, q.N);
    NEF_REQUIRE(!(q.flags & 4) || q.K % 64 == 0, "nef_pack_weights_batch: fp16 packing needs K %% 64 (K=%d)", q.K);
    q.first_block = blocks;
    const long total = (long)q.groups * q.taps * ((q.flags & 4) ? q.K >> 1 : q.K) * q.N;
    blocks += (int)((total + NEF_PACK_CHUNK - 1) / NEF_PACK_CHUNK);
  }
  pack_weights_batch_kernel<<<blocks, 256, 0, s>>>(*tab);
  NEF_CHECK_LAUNCH("pack_weights_batch_kernel");
  tab->n = 0;
  return 0;
}

extern "C" int64_t nef_cbl4_rows(int B, int L) { return (int64_t)B * (L + 2 * NEF_HALO); }
extern "C" int64_t nef_cbl4_floats(int C, int B, int L) {
  return ((int64_t)(C / 4) * nef_cbl4_rows(B, L) + NEF_GUARD_ROWS) * 4;
}

extern "C" int nef_ncl_to_cbl4(const float* src, float* dst, int B, int C, int L, int round_tf32, nef_stream_t s) {
  NEF_REQUIRE(C % 4 == 0, "nef_ncl_to_cbl4: C %% 4 required");
  const long total = (long)B * (C / 4) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ncl_to_cbl4_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(src, reinterpret_cast<float4*>(dst), B, C, L, round_tf32);
  NEF_CHECK_LAUNCH("ncl_to_cbl4_kernel");
  return 0;
}

extern "C" int nef_cbl4_to_ncl(const float* src, float* dst, int B, int C, int L, nef_stream_t s) {
  NEF_REQUIRE(C % 4 == 0, "nef_cbl4_to_ncl: C %% 4 required");
  const long total = (long)B * (C / 4) * L;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cbl4_to_ncl_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(reinterpret_cast<const float4*>(src), dst, B, C, L);
  NEF_CHECK_LAUNCH("cbl4_to_ncl_kernel");
  return 0;
}

NEF_DEFINE_EXACT_SETTER(nef_set_exact_simt)
