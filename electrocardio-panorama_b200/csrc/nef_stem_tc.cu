// Encoder stem on the tensor cores (sm_100a): Conv1d(G -> 128G, k15, s2, p7, groups=G, no bias) -> ReLU -> MaxPool1d(3,2,1),
// network/encoder/resnet_1d.py:102-105 + network/encoder/encoder.py:35-38, and its weight gradient.  The CUDA-core kernels
// in nef_elem.cu (stem_fwd_kernel / stem_bwd_kernel) compute the same thing at 30 fp32 FMAs per output element and are bound
// by the FP32 pipe at a fifth of the HBM roof; they stay as the exact-fp32 tier and as the fallback for the fp32 store.
//
// Forward.  Per lead the convolution is a GEMM over the 15 taps (K = 16 with a zero tap): conv[p][c] = sum_t x[2p + t - 7] w[c][t].
// Three accumulator sets indexed by the POOLED position j hold the three conv positions of its window,
//     E[j] = conv[2j],   O[j] = conv[2j + 1],   Om[j] = conv[2j - 1] = O[j - 1],
// so the max-pool is lane-local in the epilogue (TMEM lane = j).  The A operands are im2col tiles built in shared memory from
// the input window, K-major no-swizzle `[k chunk][row][8 halves]` (rows 16 bytes apart, as the convolution kernels stage
// their tiles): A_E[j][t] = x[4j + t - 7], A_O[j][t] = x[4j + t - 5]; Om reads the O tile through a start address one row
// (16 bytes) lower.  fp32 accuracy comes from split precision, x = x_hi + x_lo, w = w_hi + w_lo in fp16:
//     x_hi w_hi + x_lo w_hi + x_hi w_lo      (three accumulating kind::f16 MMAs per set; the dropped x_lo w_lo term is 2^-22)
// A CTA owns 32 output channels of one lead (M128 x N32 x K16 MMAs, 96 TMEM columns) and walks (segment, 128-output tile)
// units; four such CTAs share an SM, so one CTA's epilogue overlaps the others' loads and MMAs without any in-kernel pipeline.
#include <cstdlib>
#include <cuda_fp16.h>

#include "nef_elem.cuh"

namespace nef {
namespace stc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// descriptor words (SWIZZLE_NONE, version 1): lo = start address | LBO, hi = SBO
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFF) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_of(uint32_t hi, uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// fp32 accumulate (bit 4), F16 x F16, a_mn / b_mn: operand is MN-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// phase clocks of CTA (0, 0, 0), tiles 0..7 (profiling aid, nef_stem_tc_debug): start, window stored, tiles built, MMAs
// committed, accumulators ready, epilogue done
__device__ long long g_stem_dbg[8][8];
#define STEM_STAMP(slot)                                                                     \
  do {                                                                                       \
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 8) g_stem_dbg[it][slot] = clock64(); \
  } while (0)
constexpr int SF_THREADS = 128;
constexpr int SF_TJ = 128;                 // pooled outputs per tile = MMA M
constexpr int SF_XW = 4 * SF_TJ + 24;      // input window of a tile: x[4 j0 - 9 .. 4 j0 + 4 * 128 + 14]
constexpr int SF_AROWS = SF_TJ + 8;        // rows per k chunk plane of an A tile (the O tile holds rows j = -1 .. 127)
constexpr int SF_APITCH = SF_AROWS * 16;   // bytes between the two k chunks
constexpr int SF_ABYTES = 2 * SF_APITCH;   // one A tile (K = 16 halves = 2 chunks)

// SF_NC = output channels per CTA = MMA N (32: 96 TMEM columns, four CTAs per SM; 16: 48 columns, eight CTAs per SM)
template <int SF_NC>
__global__ void __launch_bounds__(SF_THREADS, 512 / (SF_NC == 32 ? 128 : 64)) stem_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, T4 y,
                                                                    uint32_t* __restrict__ amax, uint4* __restrict__ y16, int G, int L,
                                                                    int w_shared) {
  constexpr int SF_BBYTES = 2 * SF_NC * 16;   // one B tile: [k chunk][channel][8 halves]
  constexpr uint32_t TM_COLS = SF_NC == 32 ? 128 : 64;
  __shared__ __align__(128) uint8_t s_a[4 * SF_ABYTES];   // E_hi, E_lo, O_hi, O_lo
  __shared__ __align__(128) uint8_t s_b[2 * SF_BBYTES];   // w_hi, w_lo
  __shared__ __align__(16) __half s_xh[SF_XW + 8], s_xl[SF_XW + 8];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y, cq = blockIdx.z;   // lead, 32-channel quarter
  const int L4 = L / 4;
  const int ntile = (L4 + SF_TJ - 1) / SF_TJ;
  const uint32_t bar = smem_u32(&s_bar);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), TM_COLS);
  // weights of this CTA's 32 channels, split into fp16 hi / lo, K-major: slot [k chunk][n] = taps 8 kc .. 8 kc + 7 (tap 15 = 0)
  for (int i = tid; i < 2 * SF_NC; i += SF_THREADS) {
    const int kc = i / SF_NC, n = i % SF_NC;
    const float* wp = w + ((long)(w_shared ? 0 : g) * 128 + cq * SF_NC + n) * 15;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[2], h[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = kc * 8 + 2 * q + e;
        v[e] = t < 15 ? wp[t] : 0.f;
        h[e] = __half2float(__float2half_rn(v[e]));
      }
      hi[q] = f16x2_sat(v[0], v[1]);
      lo[q] = f16x2_sat(v[0] - h[0], v[1] - h[1]);
    }
    *reinterpret_cast<uint4*>(s_b + (kc * SF_NC + n) * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(s_b + SF_BBYTES + (kc * SF_NC + n) * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc_f16(128, SF_NC, 0, 0);
  const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);
  const uint32_t a0 = smem_u32(s_a), b0 = smem_u32(s_b);
  uint32_t ph = 0;

  // the input window of a tile (x[4 j0 - 9 + i], i < SF_XW + 8) is fetched into registers one tile ahead, behind the epilogue
  constexpr int XPT = (SF_XW + 8 + SF_THREADS - 1) / SF_THREADS;
  float xv[XPT];
  auto fetch_window = [&](int b, int j0) {
    const float* xb = x + ((long)b * G + g) * L;
#pragma unroll
    for (int k = 0; k < XPT; ++k) {
      const int i = tid + k * SF_THREADS;
      const int p = 4 * j0 - 9 + i;
      xv[k] = (i < SF_XW + 8 && p >= 0 && p < L) ? __ldg(xb + p) : 0.f;
    }
  };
  // (segment, tile) of unit u advance incrementally: one division pair per kernel instead of two per tile
  const int step_b = (int)gridDim.x / ntile, step_t = (int)gridDim.x % ntile;
  int b = (int)blockIdx.x / ntile, tl_i = (int)blockIdx.x % ntile;
  // per-CTA output bases: 4-channel chunk g * 32 + cq * (SF_NC / 4) of the codes, 8-channel chunk of the fp16 copy; the chunk
  // stride fits 32 bits (rows per chunk plane), so a chunk offset is one 32 x 32 -> 64-bit multiply
  const uint32_t cs32 = (uint32_t)y.cs;
  uint32_t* const amax_c = amax ? amax + (size_t)(g * 32 + cq * (SF_NC / 4)) * cs32 : nullptr;
  uint4* const y16_c = y16 + (size_t)((g * 32 + cq * (SF_NC / 4)) >> 1) * cs32;
  fetch_window(b, tl_i * SF_TJ);
  int it = -1;
  for (int u = blockIdx.x; u < y.B * ntile; u += gridDim.x) {
    ++it;
    STEM_STAMP(0);
    const int j0 = tl_i * SF_TJ;
    // ---- input window as fp16 hi / lo
#pragma unroll
    for (int k = 0; k < XPT; ++k) {
      const int i = tid + k * SF_THREADS;
      if (i < SF_XW + 8) {
        const __half h = __float2half_rn(xv[k]);
        s_xh[i] = h;
        s_xl[i] = __float2half_rn(xv[k] - __half2float(h));
      }
    }
    __syncthreads();
    STEM_STAMP(1);
    // ---- im2col tiles.  E row j (tile-local), chunk kc: x[4j - 7 + 8kc ..] = s_x[4j + 2 + 8kc ..] ; O row r = j + 1:
    //      x[4j - 5 + 8kc ..] = s_x[4r + 8kc ..]  (rows 0 .. 128).  Thread t builds row t of both tiles (both k chunks are 16
    //      consecutive halves of the window); threads 0, 1 add the two chunks of row 128 of the O tile.
    {
      const uint32_t* eh = reinterpret_cast<const uint32_t*>(s_xh + 4 * tid + 2);   // 4-byte aligned
      const uint32_t* el = reinterpret_cast<const uint32_t*>(s_xl + 4 * tid + 2);
      const uint2* oh = reinterpret_cast<const uint2*>(s_xh + 4 * tid);             // 8-byte aligned
      const uint2* ol = reinterpret_cast<const uint2*>(s_xl + 4 * tid);
      uint32_t a[8], c[8];
      uint2 d[4], e[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) { a[q] = eh[q]; c[q] = el[q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q) { d[q] = oh[q]; e[q] = ol[q]; }
      uint8_t* pe = s_a + tid * 16;
      *reinterpret_cast<uint4*>(pe) = make_uint4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<uint4*>(pe + SF_APITCH) = make_uint4(a[4], a[5], a[6], a[7]);
      *reinterpret_cast<uint4*>(pe + SF_ABYTES) = make_uint4(c[0], c[1], c[2], c[3]);
      *reinterpret_cast<uint4*>(pe + SF_ABYTES + SF_APITCH) = make_uint4(c[4], c[5], c[6], c[7]);
      uint8_t* po = s_a + 2 * SF_ABYTES + tid * 16;
      *reinterpret_cast<uint4*>(po) = make_uint4(d[0].x, d[0].y, d[1].x, d[1].y);
      *reinterpret_cast<uint4*>(po + SF_APITCH) = make_uint4(d[2].x, d[2].y, d[3].x, d[3].y);
      *reinterpret_cast<uint4*>(po + SF_ABYTES) = make_uint4(e[0].x, e[0].y, e[1].x, e[1].y);
      *reinterpret_cast<uint4*>(po + SF_ABYTES + SF_APITCH) = make_uint4(e[2].x, e[2].y, e[3].x, e[3].y);
      if (tid < 2) {   // row 128 of the O tile, chunk kc = tid
        const uint2* qh = reinterpret_cast<const uint2*>(s_xh + 4 * SF_TJ + 8 * tid);
        const uint2* ql = reinterpret_cast<const uint2*>(s_xl + 4 * SF_TJ + 8 * tid);
        uint8_t* pr = s_a + 2 * SF_ABYTES + tid * SF_APITCH + SF_TJ * 16;
        *reinterpret_cast<uint4*>(pr) = make_uint4(qh[0].x, qh[0].y, qh[1].x, qh[1].y);
        *reinterpret_cast<uint4*>(pr + SF_ABYTES) = make_uint4(ql[0].x, ql[0].y, ql[1].x, ql[1].y);
      }
    }
    fence_proxy_async();
    __syncthreads();
    STEM_STAMP(2);
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bh = desc_of(b_hi, desc_lo(b0, SF_NC * 16)), bl = desc_of(b_hi, desc_lo(b0 + SF_BBYTES, SF_NC * 16));
        // sets: 0 = E, 1 = O (row 1 of the O tile), 2 = Om (row 0 of the O tile)
        const uint32_t sa[3] = {a0, a0 + 2 * SF_ABYTES + 16, a0 + 2 * SF_ABYTES};
#pragma unroll
        for (int st = 0; st < 3; ++st) {
          const uint64_t ah = desc_of(a_hi, desc_lo(sa[st], SF_APITCH)), al = desc_of(a_hi, desc_lo(sa[st] + SF_ABYTES, SF_APITCH));
          mma_f16(tmem + st * SF_NC, ah, bh, idesc, 0u);
          mma_f16(tmem + st * SF_NC, al, bh, idesc, 1u);
          mma_f16(tmem + st * SF_NC, ah, bl, idesc, 1u);
        }
        tc_commit(bar);
      }
      __syncwarp();
    }
    STEM_STAMP(3);
    int nb = b + step_b, nt = tl_i + step_t;   // the next unit of this CTA
    if (nt >= ntile) { nt -= ntile; ++nb; }
    if (u + (int)gridDim.x < y.B * ntile) fetch_window(nb, nt * SF_TJ);   // in flight behind the MMA wait and the epilogue
    mbar_wait(bar, ph);
    ph ^= 1;
    tc_fence_after();
    STEM_STAMP(4);
    // ---- epilogue: thread = pooled position j0 + 32 warp + lane; max over (2j-1, 2j, 2j+1), ReLU, code, fp16 row stores
    {
      const int j = j0 + warp * 32 + lane;
      const bool ok = j < L4;
      const bool va = j > 0;   // conv position 2j - 1 exists
      const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
      const long row = y.row(b, ok ? j : 0);
      // all accumulator columns of this thread's row in one round trip (3 x SF_NC registers), then one pass over them
      uint32_t ve[SF_NC], vo[SF_NC], vm[SF_NC];
      if constexpr (SF_NC == 32) {
        tmem_ld32(tl + 0 * SF_NC, ve);
        tmem_ld32(tl + 1 * SF_NC, vo);
        tmem_ld32(tl + 2 * SF_NC, vm);
      } else {
        tmem_ld16(tl + 0 * SF_NC, ve);
        tmem_ld16(tl + 1 * SF_NC, vo);
        tmem_ld16(tl + 2 * SF_NC, vm);
      }
      tmem_ld_wait();
      uint32_t hp[4];
#pragma unroll
      for (int c4 = 0; c4 < SF_NC / 4; ++c4) {
        uint32_t code = 0;
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float c0 = __uint_as_float(vm[4 * c4 + k]), c1 = __uint_as_float(ve[4 * c4 + k]), c2 = __uint_as_float(vo[4 * c4 + k]);
          uint32_t best = 1;
          float bv = c1;
          if (va && c0 >= c1) { best = 0; bv = c0; }   // first maximum wins (window order), as MaxPool1d
          if (c2 > bv) { best = 2; bv = c2; }
          if (!(bv > 0.f)) { best = 3; bv = 0.f; }
          m[k] = bv;   // the fp16 conversion below is the only rounding
          code |= best << (8 * k);
        }
        // 4-channel chunk g * 32 + cq * (SF_NC / 4) + c4 of the 128 G channel space
        if (amax_c && ok) amax_c[(size_t)c4 * cs32 + row] = code;
        hp[(c4 & 1) * 2 + 0] = f16x2_sat(m[0], m[1]);
        hp[(c4 & 1) * 2 + 1] = f16x2_sat(m[2], m[3]);
        if ((c4 & 1) && ok) y16_c[(size_t)(c4 >> 1) * cs32 + row] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
      }
    }
    tc_fence_before();
    __syncthreads();
    STEM_STAMP(5);
    b = nb;
    tl_i = nt;
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------
// dW[c][t] = sum_j  gE[j][c] x[4j + t - 7] + gO[j][c] x[4j + t - 5], with the pooled gradient routed to the conv positions by
// the argmax codes of the forward:  gE[j] = dy[j] [code(j) == 1]   (position 2j)
//                                    gO[j] = dy[j] [code(j) == 2] + dy[j + 1] [code(j + 1) == 0]   (position 2j + 1).
// As a GEMM over the pooled positions (K = rows): D[128 channels x 16 taps] += gE^T . X_E + gO^T . X_O.  Both operands are
// MN-major exactly as they sit in shared memory: the routed gradients as `[8-channel chunk][row][8 halves]` (the layout of
// the fp16 gradient copy itself) and the im2col tiles of the forward kernel, `[8-tap chunk][row][8 halves]` -- the same bytes
// the forward reads K-major.  x is split hi / lo (fp32-accurate), the gradient is the loss-scaled fp16 copy the last
// data-gradient epilogue wrote.  A CTA accumulates all its (segment, 128-row tile) units of one lead in 16 TMEM columns and
// drains them once, times 1 / S, with fp32 REDs.
constexpr int SB_THREADS = 256;
constexpr int SB_TJ = 128;
constexpr int SB_GP = SB_TJ * 16;          // bytes between 8-channel chunks of a routed-gradient tile
constexpr int SB_GBYTES = 16 * SB_GP;      // one routed-gradient tile: 128 channels x 128 rows of fp16
constexpr int SB_XOFF = 2 * SB_GBYTES;     // the four im2col tiles behind the two gradient tiles
constexpr int SB_TOTAL = SB_XOFF + 4 * SF_ABYTES + 64;

// per 8-channel row: keep the halves whose code byte equals `want` (codes: one byte per channel, two words per row)
__device__ __forceinline__ uint4 keep_code(uint4 g, uint32_t c0, uint32_t c1, uint32_t want) {
  auto m2 = [&](uint32_t c, int k) {   // 16-bit masks of channels k, k + 1 of the word c
    const uint32_t lo = ((c >> (8 * k)) & 0xffu) == want ? 0x0000ffffu : 0u;
    const uint32_t hi = ((c >> (8 * k + 8)) & 0xffu) == want ? 0xffff0000u : 0u;
    return lo | hi;
  };
  return make_uint4(g.x & m2(c0, 0), g.y & m2(c0, 2), g.z & m2(c1, 0), g.w & m2(c1, 2));
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__global__ void __launch_bounds__(SB_THREADS, 2) stem_tc_bwd_kernel(const float* __restrict__ x, const uint32_t* __restrict__ amax,
                                                                    const uint4* __restrict__ dy16, T4 dy, float* __restrict__ dw,
                                                                    const float* __restrict__ inv_scale, int G, int L, int w_shared) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) __half s_xh[SF_XW + 8], s_xl[SF_XW + 8];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.y;
  const int L4 = L / 4;
  const int ntile = (L4 + SB_TJ - 1) / SB_TJ;
  const uint32_t bar = smem_u32(&s_bar);
  uint8_t* s_ge = smem;
  uint8_t* s_go = smem + SB_GBYTES;
  uint8_t* s_x = smem + SB_XOFF;   // E_hi, E_lo, O_hi, O_lo

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc_f16(128, 16, 1, 1);
  const uint32_t hi_g = desc_hi(SB_GP), hi_x = desc_hi(SF_APITCH);
  const uint32_t ge0 = smem_u32(s_ge), go0 = smem_u32(s_go), x0 = smem_u32(s_x);
  uint32_t ph = 0;
  int issued = 0;
  const int row_l = tid & (SB_TJ - 1), chalf = tid >> 7;   // this thread's tile row, and which 8 of the 16 channel chunks

  // (segment, tile) of unit u advance incrementally; per-thread bases of the gradient rows and codes (the chunk stride fits 32
  // bits: a chunk offset is one 32 x 32 -> 64-bit multiply instead of a 64 x 64 one per load)
  const int step_b = (int)gridDim.x / ntile, step_t = (int)gridDim.x % ntile;
  int b = (int)blockIdx.x / ntile, tl_i = (int)blockIdx.x % ntile;
  const uint32_t cs32 = (uint32_t)dy.cs;
  const uint4* const gp_c = dy16 + (size_t)(g * 16 + chalf * 8) * cs32;
  const uint32_t* const cp_c = amax + (size_t)(2 * (g * 16 + chalf * 8)) * cs32;
  for (int u = blockIdx.x; u < dy.B * ntile; u += gridDim.x) {
    const int j0 = tl_i * SB_TJ;
    const float* xb = x + ((long)b * G + g) * L;
    // ---- global loads of this tile (all in flight before the previous tile's MMAs are waited for)
    float xv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * SB_THREADS;
      const int p = 4 * j0 - 9 + i;
      xv[k] = (i < SF_XW + 8 && p >= 0 && p < L) ? __ldg(xb + p) : 0.f;
    }
    const int j = j0 + row_l;
    const bool v0 = j < L4, v1 = j + 1 < L4;
    const long row = dy.row(b, v0 ? j : 0);
    uint4 ga[8], gb[8];
    uint32_t ca[8][2], cb[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4* gp = gp_c + (size_t)q * cs32 + row;              // 8-channel chunk g * 16 + chalf * 8 + q
      const uint32_t* cp = cp_c + (size_t)(2 * q) * cs32 + row;
      ga[q] = v0 ? __ldg(gp) : make_uint4(0u, 0u, 0u, 0u);
      gb[q] = v1 ? __ldg(gp + 1) : make_uint4(0u, 0u, 0u, 0u);
      ca[q][0] = __ldg(cp); ca[q][1] = __ldg(cp + cs32);
      cb[q][0] = __ldg(cp + 1); cb[q][1] = __ldg(cp + cs32 + 1);
    }
    // ---- the previous tile's MMAs must have read the shared-memory tiles before they are overwritten
    if (issued) {
      mbar_wait(bar, ph);
      ph ^= 1;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = tid + k * SB_THREADS;
      if (i < SF_XW + 8) {
        const __half h = __float2half_rn(xv[k]);
        s_xh[i] = h;
        s_xl[i] = __float2half_rn(xv[k] - __half2float(h));
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t off = (uint32_t)((chalf * 8 + q) * SB_GP + row_l * 16);
      const uint4 e = keep_code(ga[q], ca[q][0], ca[q][1], 1u);
      const uint4 o2 = keep_code(ga[q], ca[q][0], ca[q][1], 2u);
      const uint4 o0 = keep_code(gb[q], cb[q][0], cb[q][1], 0u);   // disjoint supports are not guaranteed: add
      *reinterpret_cast<uint4*>(s_ge + off) = e;
      *reinterpret_cast<uint4*>(s_go + off) =
          make_uint4(hadd2_u32(o2.x, o0.x), hadd2_u32(o2.y, o0.y), hadd2_u32(o2.z, o0.z), hadd2_u32(o2.w, o0.w));
    }
    __syncthreads();
    // ---- im2col tiles (as in the forward): E row r chunk tc = s_x[4r + 2 + 8tc ..], O row r chunk tc = s_x[4r + 4 + 8tc ..]
    if (tid < SB_TJ) {
      const uint32_t* eh = reinterpret_cast<const uint32_t*>(s_xh + 4 * tid + 2);
      const uint32_t* el = reinterpret_cast<const uint32_t*>(s_xl + 4 * tid + 2);
      const uint2* oh = reinterpret_cast<const uint2*>(s_xh + 4 * tid + 4);
      const uint2* ol = reinterpret_cast<const uint2*>(s_xl + 4 * tid + 4);
      uint8_t* pe = s_x + tid * 16;
      *reinterpret_cast<uint4*>(pe) = make_uint4(eh[0], eh[1], eh[2], eh[3]);
      *reinterpret_cast<uint4*>(pe + SF_APITCH) = make_uint4(eh[4], eh[5], eh[6], eh[7]);
      *reinterpret_cast<uint4*>(pe + SF_ABYTES) = make_uint4(el[0], el[1], el[2], el[3]);
      *reinterpret_cast<uint4*>(pe + SF_ABYTES + SF_APITCH) = make_uint4(el[4], el[5], el[6], el[7]);
      uint8_t* po = s_x + 2 * SF_ABYTES + tid * 16;
      *reinterpret_cast<uint4*>(po) = make_uint4(oh[0].x, oh[0].y, oh[1].x, oh[1].y);
      *reinterpret_cast<uint4*>(po + SF_APITCH) = make_uint4(oh[2].x, oh[2].y, oh[3].x, oh[3].y);
      *reinterpret_cast<uint4*>(po + SF_ABYTES) = make_uint4(ol[0].x, ol[0].y, ol[1].x, ol[1].y);
      *reinterpret_cast<uint4*>(po + SF_ABYTES + SF_APITCH) = make_uint4(ol[2].x, ol[2].y, ol[3].x, ol[3].y);
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t gel = desc_lo(ge0, 128), gol = desc_lo(go0, 128);
        const uint32_t xeh = desc_lo(x0, 128), xel = desc_lo(x0 + SF_ABYTES, 128);
        const uint32_t xoh = desc_lo(x0 + 2 * SF_ABYTES, 128), xol = desc_lo(x0 + 3 * SF_ABYTES, 128);
#pragma unroll
        for (int ks = 0; ks < SB_TJ / 16; ++ks) {   // 16 rows = 256 bytes = 16 descriptor units per K step
          const uint64_t ae = desc_of(hi_g, gel + ks * 16), ao = desc_of(hi_g, gol + ks * 16);
          mma_f16(tmem, ae, desc_of(hi_x, xeh + ks * 16), idesc, (uint32_t)(issued | ks));
          mma_f16(tmem, ae, desc_of(hi_x, xel + ks * 16), idesc, 1u);
          mma_f16(tmem, ao, desc_of(hi_x, xoh + ks * 16), idesc, 1u);
          mma_f16(tmem, ao, desc_of(hi_x, xol + ks * 16), idesc, 1u);
        }
        tc_commit(bar);
      }
      __syncwarp();
    }
    issued = 1;
    b += step_b;
    tl_i += step_t;
    if (tl_i >= ntile) { tl_i -= ntile; ++b; }
  }
  if (issued) {
    mbar_wait(bar, ph);
    tc_fence_after();
    if (warp < 4) {   // TMEM lane = channel, column = tap
      uint32_t v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld_wait();
      const float sc = inv_scale ? __ldg(inv_scale) : 1.f;
      float* dst = dw + ((long)(w_shared ? 0 : g) * 128 + tid) * 15;
#pragma unroll
      for (int t = 0; t < 15; ++t) atomicAdd(dst + t, __uint_as_float(v[t]) * sc);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace stc

// y16: fp16 copy (half8 rows) -- the only output of this form besides the argmax codes (amax may be nullptr)
int stem_tc_fwd(const float* x, const float* w, T4 y, uint32_t* amax, void* y16, int G, cudaStream_t s, int w_shared) {
  const int L = y.L * 4;
  const int ntile = (y.L + stc::SF_TJ - 1) / stc::SF_TJ;
  const long units = (long)y.B * ntile;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static int nc = getenv("NEF_STEM_NC") ? atoi(getenv("NEF_STEM_NC")) : 32;   // A/B: 16 = eight CTAs of 16 channels per SM
  const int per_sm = nc == 16 ? 8 : 4, nz = 128 / (nc == 16 ? 16 : 32);
  long gx = ((long)sms * per_sm) / ((long)G * nz);   // per_sm CTAs per SM over (lead, channel slice) pairs
  if (gx < 1) gx = 1;
  if (gx > units) gx = units;
  dim3 grid((unsigned)gx, (unsigned)G, (unsigned)nz);
  // unused dynamic shared memory keeps one CTA too many off the SM: it would only wait for TMEM columns (all 512 are taken)
  if (nc == 16)
    stc::stem_tc_fwd_kernel<16><<<grid, stc::SF_THREADS, 6 * 1024, s>>>(x, w, y, amax, reinterpret_cast<uint4*>(y16), G, L, w_shared);
  else
    stc::stem_tc_fwd_kernel<32><<<grid, stc::SF_THREADS, 25 * 1024, s>>>(x, w, y, amax, reinterpret_cast<uint4*>(y16), G, L, w_shared);
  NEF_CHECK_LAUNCH("stem_tc_fwd_kernel");
  return 0;
}

// dy16: loss-scaled fp16 copy of the gradient of the stem output (half8 rows, geometry of dy); inv_scale[0] = 1 / S (device)
int stem_tc_bwd(const float* x, const uint32_t* amax, const void* dy16, T4 dy, float* dw, const float* inv_scale, int G,
                cudaStream_t s, int w_shared) {
  const int L = dy.L * 4;
  const int ntile = (dy.L + stc::SB_TJ - 1) / stc::SB_TJ;
  const long units = (long)dy.B * ntile;
  cudaError_t e = cudaFuncSetAttribute(stc::stem_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stc::SB_TOTAL);
  NEF_REQUIRE(e == cudaSuccess, "stem_tc_bwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long gx = ((long)sms * 2) / G;   // two CTAs per SM over the leads
  if (gx < 1) gx = 1;
  if (gx > units) gx = units;
  dim3 grid((unsigned)gx, (unsigned)G);
  stc::stem_tc_bwd_kernel<<<grid, stc::SB_THREADS, stc::SB_TOTAL, s>>>(x, amax, reinterpret_cast<const uint4*>(dy16), dy, dw, inv_scale, G, L,
                                                                        w_shared);
  NEF_CHECK_LAUNCH("stem_tc_bwd_kernel");
  return 0;
}

}  // namespace nef

extern "C" int nef_stem_tc_debug(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, nef::stc::g_stem_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
