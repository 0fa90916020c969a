// tcgen05 (5th-generation tensor core) implementation of the grouped implicit-GEMM convolution, its data
// gradient (same kernels, flipped / transposed packed weights) and its weight gradient, for sm_100a.
//
//   * TF32 operands (fp32 storage, pre-rounded to TF32 by the producing kernel), fp32 accumulation in TMEM.
//   * Operands are staged global -> shared with 1-D bulk async copies (cp.async.bulk, the TMA engine's
//     linear mode) completing on mbarriers: a CBL4 tile is one contiguous run of 16-byte rows per 4-channel
//     chunk, so no tensor map is needed, and that run IS a tcgen05 no-swizzle core-matrix column
//     (8 rows x 16 B = 128 contiguous bytes).  A k-tap convolution issues the same staged tile k times with
//     the descriptor start address advanced by 16 B per tap -- no im2col, no per-tap reload.
//   * Warp-specialised: copy producer warp(s), one MMA warp (converged, one elected lane issues; descriptor
//     words precomputed so that consecutive UTCHMMAs are back to back), 8 epilogue warps (tcgen05.ld ->
//     fused bias / residual / ReLU / dropout / angular scale / mask / BatchNorm partial statistics ->
//     coalesced float4 stores), epilogue specialised at compile time (EPI).
//
// forward / dgrad :  D[128 rows x N] += X[128 rows x 32 ch](shifted by tap) . W_tap[N x 32 ch]^T   (both K-major)
//     conv_tc_persist_kernel: one CTA per SM walks (group, 256-row tile) items, two TMEM accumulator sets, so the
//     epilogue and the pipeline fill of the next tile overlap the main loop (production path);
//     conv_tc_kernel<MT>: one CTA per MT x 128 rows (small row spaces, A/B measurements).
// wgrad           :  D_tap[cout x cin] += dY[rows x cout]^T . X[rows (+tap) x cin]: the contraction runs over rows and
//     kind::tf32 cannot read MN-major operands, so the staged tiles are re-tiled in shared memory (see wgrad_tc_kernel).
//
// Measured facts behind the structure (tools/probe_umma.cu, tools/probe_rate.cu, profiles/): kind::tf32 reads zeros for
// MN-major operands; M128 x N128 x K8 issues every 64 cycles, N = 64 every 48 (A-operand bound) unless A or B is held in
// the collector; SWIZZLE_NONE costs nothing over SWIZZLE_128B for these tiles; a bulk copy is issued from the uniform
// datapath, one at a time per warp.
#include <cstdlib>
#include <cuda_fp16.h>
#include "nef_conv.cuh"

extern "C" int nef_gconv_fwd_simt(const NefConvDesc* d, nef_stream_t s);
extern "C" int nef_gconv_wgrad_simt_range(const NefWgradDesc* d, long row0, nef_stream_t s);

namespace nef {
namespace tc {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on `bar` when all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same contraction with fp16 operands (K = 16 halves = the same 32 bytes per K step, so the shared-memory descriptors are
// those of the TF32 form): twice the MMA rate, the same 11-bit significand.
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Weight-stationary form: the B operand is latched in collector buffer b0 by ::fill and re-used by ::use / ::lastuse
// without being read from shared memory again (B = one weight stage, shared by the MT row tiles of a CTA).
//   mode 0 = fill, 1 = use, 2 = lastuse, 3 = fill and discard (no re-use)
template <int MODE>
__device__ __forceinline__ void mma_tf32_ws(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (MODE == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::tf32.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (MODE == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::tf32.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (MODE == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::tf32.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::tf32.collector::b0::discard [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// the same for fp16 operands: at kind::f16 rate the MMAs are bound by shared-memory operand reads (4 KB of A + 4 KB of B per
// M128 x N128 x K16 step), so re-using the weight stage B for the second row tile of the CTA removes a quarter of them
template <int MODE>
__device__ __forceinline__ void mma_f16_ws(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (MODE == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// A-operand re-use: the A tile is latched in the collector by ::fill and re-used by ::use / ::lastuse (same encoding of
// MODE as above) -- for consecutive MMAs that share A (the taps of a weight-gradient K step share dY^T).
template <int MODE>
__device__ __forceinline__ void mma_tf32_areuse(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (MODE == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (MODE == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, 16-byte units:
//   K-major  (K = channels):  core matrix = 8 rows x 16 B;  SBO = next 8 rows,        LBO = next 4-channel chunk
//   MN-major (K = rows)    :  core matrix = 8 rows x 16 B;  SBO = next 4-channel chunk, LBO = next 8 rows
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// The issue loop must stay short: a descriptor is (hi = SBO | version, lo = start address | LBO) and only the start
// address changes between the MMAs of a stage, so lo words are formed once per stage and advanced by plain adds.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFF) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint64_t desc_of(uint32_t hi, uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
constexpr uint32_t DESC_HI_SBO128 = (128u >> 4) | (1u << 14);  // SBO = 128 bytes, descriptor version 1, SWIZZLE_NONE
// one lane of a converged warp (the lowest); the MMAs and their commits are issued under it so that the compiler keeps
// the operands in uniform registers instead of serialising every MMA through a lane-election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Instruction descriptor: fp32 accumulate, TF32 x TF32, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor: fp32 accumulate, F16 x F16 (kind::f16, operand format 0), M x N, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// phase timestamps of the first CTAs of the last conv launch (profiling aid, read by nef_tc_debug_dump)
__device__ unsigned long long g_tc_dbg[1024][8];
__device__ __forceinline__ void dbg_stamp(int slot) {
  const int cta = blockIdx.x + gridDim.x * blockIdx.y;
  if (cta < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tc_dbg[cta][slot] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// forward / data-gradient kernel
// ---------------------------------------------------------------------------------------------
constexpr int FW_THREADS = 320;  // warp 0 copy producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int FW_XST = 2;  // activation stages (one per 32-channel block)
constexpr int FW_WBYTES = 8 * 128 * 16;  // one weight stage: 32 input channels x up to 128 outputs x one tap

template <int MT>
struct FwSmem {
  static constexpr int XROWS = MT * 128 + 8;          // rows per chunk in a stage (taps - 1 <= 6 extra)
  static constexpr int XPITCH = XROWS * 16;           // bytes between channel chunks
  static constexpr int XBYTES = 8 * XPITCH;           // one activation stage
  // MT == 4: one CTA per SM (all 512 TMEM columns).  MT == 2: two CTAs per SM (256 columns each), so that one CTA's
  // epilogue overlaps the other's main loop; each gets half of the shared memory.
  static constexpr int BUDGET = (MT == 2 ? 113 : 227) * 1024;
  static constexpr int WST_FIT = (BUDGET - 6144 - FW_XST * XBYTES) / FW_WBYTES;
  static constexpr int WST = WST_FIT > 6 ? 6 : WST_FIT;  // weight stages
  static constexpr int BAR_OFF = FW_XST * XBYTES + WST * FW_WBYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = STAT_OFF + 4 * 2 * 128 * 4 + 128;
};

// EPI: compile-time epilogue selection (bit 0 bias, 1 residual tensor, 2 ReLU, 3 dropout, 4 mask mode 1, 5 TF32 rounding of
// the stored value; bit 6 = generic: every option read from the descriptor at run time, incl. BatchNorm statistics, the
// angular scale and its gradient, mask mode 2).  The specialised bodies are ~4x shorter, which matters: the epilogue runs
// on 8 warps only and the generic body does not fit the instruction cache.
constexpr int EPI_BIAS = 1, EPI_RES = 2, EPI_RELU = 4, EPI_DROP = 8, EPI_MASK1 = 16, EPI_ROUND = 32, EPI_GENERIC = 64;
constexpr int EPI_MASK2 = 128, EPI_BSCALE = 256, EPI_STATS = 512;  // mask mode 2, angular scale, BatchNorm partial statistics
constexpr int EPI_OBITS = 1024, EPI_MBITS = 2048;                  // write / read one-bit activation masks
constexpr int EPI_Y16 = 4096;                                      // also store an fp16 copy of the output (next conv's operand)
constexpr int EPI_RES16 = 16384;   // the residual operand is read from an fp16 copy (times res16_scale[0]) instead of the fp32 tensor
constexpr int EPI_NOY = 32768;     // no fp32 store at all (y == NULL): only the fp16 copy / the bit plane of the result is kept
constexpr int EPI_GSCALE = 8192;   // backward pass with loss-scaled fp16 copies: acc *= acc_scale[0]; y16 = fp16(v * y16_scale[0])

// Sum over the 32 lanes of 32 per-lane values at once: after the 5 exchange levels lane j holds the warp total of a[j].
// 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_sum32(float (&a)[32], int lane) {
#pragma unroll
  for (int bit = 16; bit >= 1; bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < bit; ++i) {
      const float send = upper ? a[i] : a[i + bit];
      const float keep = upper ? a[i + bit] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return a[0];
}

// two floats -> packed fp16 pair (lo in the low half = lower address), saturating to the largest finite value
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float4 rn4_tf32(float4 v) {
  uint32_t a, b, c, e;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(v.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(e) : "f"(v.w));
  return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(e));
}

// Epilogue of one CTA tile (MT row tiles of 128 x N output channels of group g, rows from r0), run by the 8 epilogue
// warps: warp w owns TMEM lanes 32 * (w % 4) .. + 31 and the 32-column groups cg = (w - 2) / 4 (mod 2).
// tmem_acc = TMEM address of the tile's first accumulator column.
template <int MT, int EPI, int EW = 8>
__device__ __forceinline__ void epilogue_tile(const NefConvDesc& d, int g, long r0, uint32_t tmem, int warp, int lane, int tid,
                                              float* s_stat) {
  const int N = d.N;
  const int q = warp & 3, chalf = (warp - 2) >> 2;   // EW epilogue warps: EW / 4 share a lane quarter, splitting its column groups
  constexpr bool GEN = (EPI & EPI_GENERIC) != 0;
  const bool want_stats = GEN ? d.stat_sum != nullptr : (EPI & EPI_STATS) != 0;
  const bool f_bias = GEN ? d.bias != nullptr : (EPI & EPI_BIAS) != 0;
  const bool f_res = GEN ? d.res != nullptr : (EPI & EPI_RES) != 0;
  const bool f_relu = GEN ? d.relu != 0 : (EPI & EPI_RELU) != 0;
  const bool f_drop = GEN ? d.drop_p > 0.f : (EPI & EPI_DROP) != 0;
  const bool f_round = GEN ? d.round_tf32 != 0 : (EPI & EPI_ROUND) != 0;
  const bool f_bscale = GEN ? d.bscale != nullptr : (EPI & EPI_BSCALE) != 0;
  const bool f_bsgrad = GEN && d.bscale_grad != nullptr;
  const bool f_y16 = GEN ? d.y16 != nullptr : (EPI & EPI_Y16) != 0;
  const bool f_obits = GEN ? d.out_bits != nullptr : (EPI & EPI_OBITS) != 0;
  const bool f_mbits = GEN ? (d.mask_bits != nullptr && d.mask_mode != 0) : (EPI & EPI_MBITS) != 0;
  const bool f_gs = GEN ? (d.acc_scale != nullptr || d.y16_scale != nullptr) : (EPI & EPI_GSCALE) != 0;
  const float acc_sc = (f_gs && d.acc_scale) ? __ldg(d.acc_scale) : 1.f;
  const float y16_sc = (f_gs && d.y16_scale) ? __ldg(d.y16_scale) : 1.f;
  const bool store_y = GEN ? d.y != nullptr : (EPI & EPI_NOY) == 0;   // NULL: only the fp16 copy (and the bit plane) of the result is kept
  const bool f_res16 = GEN ? d.res16 != nullptr : (EPI & EPI_RES16) != 0;
  const float res_sc = (f_res16 && d.res16_scale) ? __ldg(d.res16_scale) : 1.f;
  // mask operand: 0 none, 1 / 2 float tensor (> 0 / != 0), 3 one-bit masks
  const int mask_mode = f_mbits ? 3 : (GEN ? d.mask_mode : ((EPI & EPI_MASK1) ? 1 : ((EPI & EPI_MASK2) ? 2 : 0)));
  const float mask_scale = d.mask_scale;
  const uint32_t drop_thr = (uint32_t)(d.drop_p * 65536.f);
  const float drop_sc = 1.f / (1.f - d.drop_p);
  const long ctot = (long)d.groups * N;
  const long n_rec = (d.rows + 127) / 128;
  const float4* bias4 = reinterpret_cast<const float4*>(d.bias) + g * (N >> 2);
  for (int mt = 0; mt < MT; ++mt) {
    const long row = r0 + mt * 128 + q * 32 + lane;
    const EpiRow er = epi_row(d, row);
    const int b0 = __shfl_sync(0xffffffffu, er.b, 0);
    const bool uniform_b = __all_sync(0xffffffffu, er.b == b0);
    // per-row base pointers; chunk n4 of the group is n4 * (chunk stride) further.  Rows outside the tensor read row 0
    // (always mapped) so that the operand loads are unconditional and can all be in flight together.
    const long orow = er.valid ? er.out_row : 0;
    float4* yp = reinterpret_cast<float4*>(d.y) + (long)(d.y_c4_off + g * d.y_c4_gstride) * d.y_cstride + orow;
    const float4* rp = reinterpret_cast<const float4*>(d.res) + (long)(d.res_c4_off + g * d.res_c4_gstride) * d.res_cstride + orow;
    const float4* mp = reinterpret_cast<const float4*>(d.mask) + (long)(d.mask_c4_off + g * d.mask_c4_gstride) * d.mask_cstride + orow;
    const float4* bs4 = reinterpret_cast<const float4*>(d.bscale) + (long)(er.valid ? er.b : 0) * (ctot >> 2) + g * (N >> 2);
    const uint32_t* mbp = d.mask_bits + (long)((d.mask_c4_off + g * d.mask_c4_gstride) >> 3) * d.mask_cstride + orow;
    uint32_t* obp = d.out_bits + (long)((d.y_c4_off + g * d.y_c4_gstride) >> 3) * d.y_cstride + orow;
    uint4* y16p = reinterpret_cast<uint4*>(d.y16) + (long)((d.y_c4_off + g * d.y_c4_gstride) >> 1) * d.y_cstride + orow;
    const uint4* r16p = reinterpret_cast<const uint4*>(d.res16) + (long)((d.res_c4_off + g * d.res_c4_gstride) >> 1) * d.res_cstride + orow;
    for (int cg = chalf; cg < N / 32; cg += EW / 4) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * N + cg * 32), v);
      const uint32_t mword = mask_mode == 3 ? __ldg(mbp + (long)cg * d.mask_cstride) : 0u;
      uint32_t oword = 0u;
      // the residual / mask operands of these 8 chunks are fetched while the TMEM load is in flight
      float4 rr[8], mm[8];
      {
        const float4* rpc = rp + (long)(cg * 8) * d.res_cstride;
        const float4* mpc = mp + (long)(cg * 8) * d.mask_cstride;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rr[i] = f_res ? __ldg(rpc) : f4zero();
          mm[i] = (mask_mode == 1 || mask_mode == 2) ? __ldg(mpc) : f4zero();
          rpc += d.res_cstride;
          mpc += d.mask_cstride;
        }
        if (f_res16) {   // 8 channels per 16-byte row: chunk pair (2 m, 2 m + 1) of this column group
          const uint4* rq = r16p + (long)(cg * 4) * d.res_cstride;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const uint4 h = __ldg(rq);
            rq += d.res_cstride;
            const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&h.z)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&h.w));
            rr[2 * m] = make_float4(a0.x * res_sc, a0.y * res_sc, a1.x * res_sc, a1.y * res_sc);
            rr[2 * m + 1] = make_float4(b0.x * res_sc, b0.y * res_sc, b1.x * res_sc, b1.y * res_sc);
          }
        }
      }
      tmem_ld_wait();
      float st1[32], st2[32];
      uint32_t h16[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n4 = cg * 8 + i;
        float4 x = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
        if (f_gs) x = x * acc_sc;
        if (f_bias) x = x + __ldg(bias4 + n4);
        if (f_res || f_res16) x = x + rr[i];
        const float4 pre = x;
        if (f_relu) x = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
        if (f_drop) {
          const uint64_t bits = drop_bits(d.drop_seed, er.out_row, d.y_c4_off + g * d.y_c4_gstride + n4);
          const uint32_t blo = (uint32_t)bits, bhi = (uint32_t)(bits >> 32);   // 32-bit compares (a 64-bit one is two instructions)
          x.x = ((blo & 0xffffu) >= drop_thr) ? x.x * drop_sc : 0.f;
          x.y = ((blo >> 16) >= drop_thr) ? x.y * drop_sc : 0.f;
          x.z = ((bhi & 0xffffu) >= drop_thr) ? x.z * drop_sc : 0.f;
          x.w = ((bhi >> 16) >= drop_thr) ? x.w * drop_sc : 0.f;
        }
        if (f_bscale) {
          const float4 sc4 = __ldg(bs4 + n4);
          if (f_bsgrad) {  // forward was ys = relu(.) * s with the mask tensor = ys ; d s += x * ys / s
            const float4 m = mm[i];
            float4 bsg;
            bsg.x = sc4.x != 0.f ? x.x * m.x / sc4.x : 0.f;
            bsg.y = sc4.y != 0.f ? x.y * m.y / sc4.y : 0.f;
            bsg.z = sc4.z != 0.f ? x.z * m.z / sc4.z : 0.f;
            bsg.w = sc4.w != 0.f ? x.w * m.w / sc4.w : 0.f;
            const long cbase = (long)g * N + n4 * 4;
            if (uniform_b) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float s1 = warp_sum(er.valid ? f4get(bsg, j) : 0.f);
                if (lane == 0 && s1 != 0.f) atomicAdd(d.bscale_grad + (long)b0 * ctot + cbase + j, s1);
              }
            } else if (er.valid) {
#pragma unroll
              for (int j = 0; j < 4; ++j) atomicAdd(d.bscale_grad + (long)er.b * ctot + cbase + j, f4get(bsg, j));
            }
          }
          x = x * sc4;
        }
        if (mask_mode == 1) {
          const float4 m = mm[i];
          x = make_float4(m.x > 0.f ? x.x * mask_scale : 0.f, m.y > 0.f ? x.y * mask_scale : 0.f,
                          m.z > 0.f ? x.z * mask_scale : 0.f, m.w > 0.f ? x.w * mask_scale : 0.f);
        } else if (mask_mode == 2) {
          const float4 m = mm[i];
          x = make_float4(m.x != 0.f ? x.x * mask_scale : 0.f, m.y != 0.f ? x.y * mask_scale : 0.f,
                          m.z != 0.f ? x.z * mask_scale : 0.f, m.w != 0.f ? x.w * mask_scale : 0.f);
        } else if (mask_mode == 3) {
          const uint32_t mb = mword >> (4 * i);
          x = make_float4((mb & 1u) ? x.x * mask_scale : 0.f, (mb & 2u) ? x.y * mask_scale : 0.f,
                          (mb & 4u) ? x.z * mask_scale : 0.f, (mb & 8u) ? x.w * mask_scale : 0.f);
        }
        // The stored fp32 value is pre-rounded to TF32 (RNA) so that it and its fp16 copy hold the same significand; with
        // no fp32 store the fp16 conversion below (RN) is the only rounding (cvt.rna.tf32 is three instructions per value).
        if (f_round && store_y) x = rn4_tf32(x);
        if (er.valid && store_y) yp[(long)n4 * d.y_cstride] = x;
        if (f_y16) {  // chunks n4 = 2m, 2m + 1 form the 8-channel fp16 chunk m
          const float4 xs = f_gs ? x * y16_sc : x;
          h16[(i & 1) * 2 + 0] = pack_f16x2(xs.x, xs.y);
          h16[(i & 1) * 2 + 1] = pack_f16x2(xs.z, xs.w);
          if ((i & 1) && er.valid) y16p[(long)(n4 >> 1) * d.y_cstride] = make_uint4(h16[0], h16[1], h16[2], h16[3]);
        }
        if (f_obits)
          oword |= ((x.x != 0.f ? 1u : 0u) | (x.y != 0.f ? 2u : 0u) | (x.z != 0.f ? 4u : 0u) | (x.w != 0.f ? 8u : 0u)) << (4 * i);
        if (want_stats) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pv = er.valid ? f4get(pre, j) : 0.f;
            st1[4 * i + j] = pv;
            st2[4 * i + j] = pv * pv;
          }
        }
      }
      if (f_obits && er.valid) obp[(long)cg * d.y_cstride] = oword;
      if (want_stats) {  // lane j ends up with the 32-row totals of channel cg * 32 + j
        const float t1 = warp_sum32(st1, lane), t2 = warp_sum32(st2, lane);
        s_stat[(q * 2 + 0) * 128 + cg * 32 + lane] = t1;
        s_stat[(q * 2 + 1) * 128 + cg * 32 + lane] = t2;
      }
    }
    if (want_stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      const int et = tid - 64;  // 0 .. EW * 32 - 1
      const long rec = r0 / 128 + mt;
      if (et < N && rec < n_rec) {
        const long o = rec * ctot + (long)g * N + et;
        d.stat_sum[o] = (s_stat[0 * 128 + et] + s_stat[2 * 128 + et]) + (s_stat[4 * 128 + et] + s_stat[6 * 128 + et]);
        d.stat_sq[o] = (s_stat[1 * 128 + et] + s_stat[3 * 128 + et]) + (s_stat[5 * 128 + et] + s_stat[7 * 128 + et]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
  }
}

// L2 prefetch of the fp16 residual rows (and the mask words that go with them) of a tile, issued by the epilogue warps of the
// persistent kernel BEFORE they wait for the tile's accumulators: the loads in epilogue_tile then hit L2 instead of paying the
// DRAM latency once per (row tile, column group) round while the accumulator set is held.  Measured at 256 x 12 x 5000: k7 data
// gradient with the skip gradient added 0.82 -> 0.73 ms, the k3 one 0.73 -> 0.59 ms.  (Loading the next round's rows into
// registers one round ahead on top of this changed nothing.)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int MT, int EPI, int EW>
__device__ __forceinline__ void epilogue_prefetch(const NefConvDesc& d, int g, long r0, int warp, int lane) {
  if constexpr ((EPI & EPI_GENERIC) == 0 && (EPI & EPI_RES16) != 0) {
    const int N = d.N;
    const int q = warp & 3, chalf = (warp - 2) >> 2;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const EpiRow er = epi_row(d, r0 + mt * 128 + q * 32 + lane);
      const long orow = er.valid ? er.out_row : 0;
      for (int cg = chalf; cg < N / 32; cg += EW / 4) {
        if ((lane & 3) == 0) {   // 16 bytes per row: four rows (two with an output stride of 2) share a 64-byte half line
          const uint4* rq = reinterpret_cast<const uint4*>(d.res16) + (long)(((d.res_c4_off + g * d.res_c4_gstride) >> 1) + cg * 4) * d.res_cstride + orow;
#pragma unroll
          for (int m = 0; m < 4; ++m) prefetch_l2(rq + (long)m * d.res_cstride);
        }
        if constexpr ((EPI & EPI_MBITS) != 0) {   // (the mask words alone, without a residual, cost the memory-bound k3 data gradients 5 %)
          if ((lane & 15) == 0) prefetch_l2(d.mask_bits + (long)(((d.mask_c4_off + g * d.mask_c4_gstride) >> 3) + cg) * d.mask_cstride + orow);
        }
      }
    }
  }
}

template <int MT, int EPI>
__global__ void __launch_bounds__(FW_THREADS, MT == 2 ? 2 : 1) conv_tc_kernel(const __grid_constant__ NefConvDesc d,
                                                                              int first_wave, int stagger_cycles, int use_ws) {
  using S = FwSmem<MT>;
  // Two CTAs share an SM (MT == 2) so that one's epilogue and pipeline fill overlap the other's main loop -- but CTAs
  // launched together run in lock step (same duration), both in the main loop, then both in the epilogue.  The CTAs
  // that take the second slot of each SM in the first wave therefore start half a CTA lifetime late; equal durations
  // keep the two slots out of phase for the rest of the launch.
  if (stagger_cycles > 0) {
    const int cta = blockIdx.x + gridDim.x * blockIdx.y;
    if (cta >= first_wave / 2 && cta < first_wave) {
      const long long t0 = clock64();
      while (clock64() - t0 < stagger_cycles) {}
    }
  }
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t xs0 = sbase, ws0 = sbase + FW_XST * S::XBYTES, bar0 = sbase + S::BAR_OFF;
  // barriers (8 bytes each): full_x[2], empty_x[2], full_w[WST], empty_w[WST], acc_full ; then the TMEM base
  auto full_x = [&](int i) { return bar0 + 8 * i; };
  auto empty_x = [&](int i) { return bar0 + 8 * (FW_XST + i); };
  auto full_w = [&](int i) { return bar0 + 8 * (2 * FW_XST + i); };
  auto empty_w = [&](int i) { return bar0 + 8 * (2 * FW_XST + S::WST + i); };
  const uint32_t acc_full = bar0 + 8 * (2 * FW_XST + 2 * S::WST);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::BAR_OFF + 8 * (2 * FW_XST + 2 * S::WST + 1));
  float* s_stat = reinterpret_cast<float*>(smem + S::STAT_OFF);  // [4 quarters][2][128]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y;
  const long r0 = (long)blockIdx.x * (MT * 128);
  const int N = d.N;
  constexpr uint32_t TM_COLS = MT * 128 <= 32 ? 32 : (MT * 128 <= 64 ? 64 : (MT * 128 <= 128 ? 128 : (MT * 128 <= 256 ? 256 : 512)));

  if (tid == 0) dbg_stamp(0);
  if (tid == 0) {
    for (int i = 0; i < FW_XST; ++i) { mbar_init(full_x(i), 1); mbar_init(empty_x(i), 1); }
    for (int i = 0; i < S::WST; ++i) { mbar_init(full_w(i), 1); mbar_init(empty_w(i), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) dbg_stamp(1);

  if (warp == 0) {
    // ===== copy producer (lane c copies channel chunk c of every activation stage, lane 0 the weight stages) =====
    int xs = 0, xph = 0, wst = 0, wph = 0;
    for (int ti = 0; ti < d.n_terms; ++ti) {
      const NefConvTerm& t = d.term[ti];
      const int nkb = t.cin_g >> 5;
      const uint32_t xbytes = (uint32_t)(MT * 128 + t.taps - 1) * 16;
      const uint32_t wbytes = (uint32_t)(8 * N * 16);
      const float4* xg = reinterpret_cast<const float4*>(t.x) + (r0 + t.tap_off);
      const float4* wg = reinterpret_cast<const float4*>(t.w);
      for (int kb = 0; kb < nkb; ++kb) {
        if (lane == 0) {
          mbar_wait(empty_x(xs), xph ^ 1);
          mbar_expect_tx(full_x(xs), 8 * xbytes);
        }
        __syncwarp();
        if (lane < 8) {
          const long chunk = t.x_c4_off + (long)g * t.x_c4_gstride + kb * 8 + lane;
          bulk_g2s(xs0 + xs * S::XBYTES + lane * S::XPITCH, xg + chunk * t.x_cstride, xbytes, full_x(xs));
        }
        if (++xs == FW_XST) { xs = 0; xph ^= 1; }
        if (lane == 0) {
          for (int tp = 0; tp < t.taps; ++tp) {
            mbar_wait(empty_w(wst), wph ^ 1);
            mbar_expect_tx(full_w(wst), wbytes);
            bulk_g2s(ws0 + wst * FW_WBYTES, wg + ((((long)g * t.taps + tp) * nkb + kb) * 8) * N, wbytes, full_w(wst));
            if (++wst == S::WST) { wst = 0; wph ^= 1; }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp runs the (warp-uniform) control flow converged, one elected lane issues =====
    const uint32_t idesc = make_idesc(128, N, 0, 0), idesc16 = make_idesc_f16(128, N);
    const uint32_t nb = 2u * (uint32_t)N;                    // descriptor units between two K = 8 steps of a weight stage
    int xs = 0, xph = 0, wst = 0, wph = 0;
    uint32_t accum = 0;
    for (int ti = 0; ti < d.n_terms; ++ti) {
      const NefConvTerm& t = d.term[ti];
      const int nkb = t.cin_g >> 5;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_x(xs), xph);
        for (int tp = 0; tp < t.taps; ++tp) {
          mbar_wait(full_w(wst), wph);
          if (accum == 0 && lane == 0) dbg_stamp(2);
          tc_fence_after();
          const uint32_t xa = desc_lo(xs0 + xs * S::XBYTES + tp * 16, S::XPITCH);
          const uint32_t wa = desc_lo(ws0 + wst * FW_WBYTES, (uint32_t)N * 16);
          if (elect_one()) {
            if (use_ws && MT > 1 && !t.x_f16) {
#pragma unroll
              for (int k8 = 0; k8 < 4; ++k8) {
                const uint64_t bd = desc_of(DESC_HI_SBO128, wa + k8 * nb);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint64_t ad = desc_of(DESC_HI_SBO128, xa + k8 * 2 * (S::XPITCH >> 4) + mt * 128);
                  if (mt == 0) mma_tf32_ws<0>(tmem + mt * N, ad, bd, idesc, accum | (uint32_t)k8);
                  else if (mt == MT - 1) mma_tf32_ws<2>(tmem + mt * N, ad, bd, idesc, accum | (uint32_t)k8);
                  else mma_tf32_ws<1>(tmem + mt * N, ad, bd, idesc, accum | (uint32_t)k8);
                }
              }
            } else if (t.x_f16) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  mma_f16(tmem + mt * N, desc_of(DESC_HI_SBO128, xa + k8 * 2 * (S::XPITCH >> 4) + mt * 128),
                          desc_of(DESC_HI_SBO128, wa + k8 * nb), idesc16, accum | (uint32_t)k8);
              }
            } else {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k8 = 0; k8 < 4; ++k8)
                  mma_tf32(tmem + mt * N, desc_of(DESC_HI_SBO128, xa + k8 * 2 * (S::XPITCH >> 4) + mt * 128),
                           desc_of(DESC_HI_SBO128, wa + k8 * nb), idesc, accum | (uint32_t)k8);
              }
            }
            tc_commit(empty_w(wst));
          }
          __syncwarp();
          accum = 1;
          if (++wst == S::WST) { wst = 0; wph ^= 1; }
        }
        if (elect_one()) tc_commit(empty_x(xs));
        __syncwarp();
        if (++xs == FW_XST) { xs = 0; xph ^= 1; }
      }
    }
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
    if (lane == 0) dbg_stamp(3);
  } else {
    // ===== epilogue: 8 warps; warp w owns TMEM lanes 32 * (w % 4) .. + 31 and the 32-column groups cg = (w - 2) / 4 (mod 2)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tid == 64) dbg_stamp(4);
    epilogue_tile<MT, EPI>(d, g, r0, tmem, warp, lane, tid, s_stat);
    if (tid == 64) dbg_stamp(5);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
  if (tid == 0) dbg_stamp(6);
}

// ---------------------------------------------------------------------------------------------
// persistent forward / data-gradient kernel: one CTA per SM walks (group, 256-row tile) work items; two accumulator
// sets in TMEM (2 x (2 row tiles x N columns)), so the epilogue of tile i and the operand pipeline fill of tile i + 1
// overlap the main loop.  Same operand staging, MMA issue and epilogue as conv_tc_kernel.
// ---------------------------------------------------------------------------------------------
struct PsSmem {
  static constexpr int MT = 2;
  static constexpr int XROWS = MT * 128 + 8;
  static constexpr int XPITCH = XROWS * 16;
  static constexpr int XBYTES = 8 * XPITCH;
#ifndef NEF_PS_XST
#define NEF_PS_XST 3
#endif
#ifndef NEF_PS_WST
#define NEF_PS_WST 6
#endif
  static constexpr int XST = NEF_PS_XST;   // activation stages (33 KB each) and weight stages (16 KB each) of the operand pipeline
  static constexpr int WST = NEF_PS_WST;   // (-DNEF_PS_XST / -DNEF_PS_WST: A/B builds, loaded through NEFNET_B200_LIB)
  static constexpr int BAR_OFF = XST * XBYTES + WST * FW_WBYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int TOTAL = STAT_OFF + 4 * 2 * 128 * 4 + 128;
};
static_assert(PsSmem::TOTAL <= 227 * 1024, "persistent conv shared memory");

// Epilogue warps of the persistent kernel: 16 (four per scheduler) hide the TMEM-load / global-load latency of the
// memory-bound launches (k <= 3: the masked k3 data gradient is 16 % faster than with 8); the BatchNorm-statistics and generic
// bodies need more than the 112 registers that leaves per thread and stay at 8.  The tensor-bound k7 launches are better off
// with 8 as well (3-5 %: the MMA-issuing warp shares its scheduler with two busy epilogue warps instead of four), so the
// epilogues those layers use (NEF_TC_EPI_HEAVY) are also instantiated with 8 and picked per launch by MMAs per tile.
template <int EPI>
#ifndef NEF_PS_EW
#define NEF_PS_EW 16
#endif
constexpr int ps_epi_warps() { return (EPI & (EPI_STATS | EPI_GENERIC)) ? 8 : NEF_PS_EW; }

template <int EPI, int EW = ps_epi_warps<EPI>()>
__global__ void __launch_bounds__(64 + 32 * EW, 1) conv_tc_persist_kernel(const __grid_constant__ NefConvDesc d, int tiles_per_group,
                                                                        int n_tiles, int use_ws) {
  using S = PsSmem;
  constexpr int MT = S::MT;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t xs0 = sbase, ws0 = sbase + S::XST * S::XBYTES, bar0 = sbase + S::BAR_OFF;
  auto full_x = [&](int i) { return bar0 + 8 * i; };
  auto empty_x = [&](int i) { return bar0 + 8 * (S::XST + i); };
  auto full_w = [&](int i) { return bar0 + 8 * (2 * S::XST + i); };
  auto empty_w = [&](int i) { return bar0 + 8 * (2 * S::XST + S::WST + i); };
  auto acc_full = [&](int i) { return bar0 + 8 * (2 * S::XST + 2 * S::WST + i); };
  auto acc_empty = [&](int i) { return bar0 + 8 * (2 * S::XST + 2 * S::WST + 2 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + S::BAR_OFF + 8 * (2 * S::XST + 2 * S::WST + 4));
  float* s_stat = reinterpret_cast<float*>(smem + S::STAT_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = d.N;
  constexpr uint32_t TM_COLS = 512;

  if (tid == 0) {
    for (int i = 0; i < S::XST; ++i) { mbar_init(full_x(i), 1); mbar_init(empty_x(i), 1); }
    for (int i = 0; i < S::WST; ++i) { mbar_init(full_w(i), 1); mbar_init(empty_w(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full(i), 1); mbar_init(acc_empty(i), 32 * EW); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== copy producer =====
    int xs = 0, xph = 0, wst = 0, wph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int g = tile / tiles_per_group;
      const long r0 = (long)(tile - g * tiles_per_group) * (MT * 128);
      for (int ti = 0; ti < d.n_terms; ++ti) {
        const NefConvTerm& t = d.term[ti];
        const int nkb = t.cin_g >> 5;
        const uint32_t xbytes = (uint32_t)(MT * 128 + t.taps - 1) * 16;
        const uint32_t wbytes = (uint32_t)(8 * N * 16);
        const float4* xg = reinterpret_cast<const float4*>(t.x) + (r0 + t.tap_off);
        const float4* wg = reinterpret_cast<const float4*>(t.w);
        for (int kb = 0; kb < nkb; ++kb) {
          if (lane == 0) {
            mbar_wait(empty_x(xs), xph ^ 1);
            if (use_ws & 4) mbar_arrive(full_x(xs));   // timing experiment (NEF_TC_WS bit 2): no operand copies at all
            else mbar_expect_tx(full_x(xs), 8 * xbytes);
          }
          __syncwarp();
          if (lane < 8 && !(use_ws & 4)) {
            const long chunk = t.x_c4_off + (long)g * t.x_c4_gstride + kb * 8 + lane;
            bulk_g2s(xs0 + xs * S::XBYTES + lane * S::XPITCH, xg + chunk * t.x_cstride, xbytes, full_x(xs));
          }
          if (++xs == S::XST) { xs = 0; xph ^= 1; }
          if (lane == 0) {
            for (int tp = 0; tp < t.taps; ++tp) {
              mbar_wait(empty_w(wst), wph ^ 1);
              if (use_ws & 4) {
                mbar_arrive(full_w(wst));
              } else {
                mbar_expect_tx(full_w(wst), wbytes);
                bulk_g2s(ws0 + wst * FW_WBYTES, wg + ((((long)g * t.taps + tp) * nkb + kb) * 8) * N, wbytes, full_w(wst));
              }
              if (++wst == S::WST) { wst = 0; wph ^= 1; }
            }
          } else {
            wst = (wst + t.taps) % S::WST;  // (only lane 0 uses the weight stage counters)
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The tensor pipe does not run far ahead of this warp (measured: every cycle between the last MMA of a tap and the
    // first of the next shows up in the launch time), so the per-tap path is kept as short as possible: operand kind and
    // tap count are hoisted out of the tap loop, descriptor words advance by adds, the wait accounting (clock reads) only
    // runs in the timing mode of tools/issuer_waits.py (NEF_TC_WS bit 3).
    const uint32_t idesc = make_idesc(128, N, 0, 0), idesc16 = make_idesc_f16(128, N);
    const uint32_t nb = 2u * (uint32_t)N;
    const bool timing = (use_ws & 8) != 0;
    int xs = 0, xph = 0, wst = 0, wph = 0, as = 0, aph = 0;
    long long wt_a = 0, wt_x = 0, wt_w = 0, tq = 0;
    const long long t_begin = clock64();
    const uint32_t xlo0 = desc_lo(xs0, S::XPITCH), wlo0 = desc_lo(ws0, (uint32_t)N * 16);
    constexpr uint32_t XK = 2 * (S::XPITCH >> 4);   // descriptor units between two K = 16-half (8-float) steps of an activation stage
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (timing) tq = clock64();
      mbar_wait(acc_empty(as), aph ^ 1);
      if (timing) wt_a += clock64() - tq;
      tc_fence_after();
      const uint32_t acc0 = tmem + (uint32_t)(as * 256);
      uint32_t accum = 0;
      for (int ti = 0; ti < d.n_terms; ++ti) {
        const int nkb = d.term[ti].cin_g >> 5, taps = d.term[ti].taps;
        const bool f16 = d.term[ti].x_f16 != 0;
        for (int kb = 0; kb < nkb; ++kb) {
          if (timing) tq = clock64();
          mbar_wait(full_x(xs), xph);
          if (timing) wt_x += clock64() - tq;
          uint32_t xa = xlo0 + (uint32_t)xs * (S::XBYTES >> 4);
          // All taps of this channel block are issued by the elected lane alone.  Per tap: the 2 x 4 MMAs of the two row
          // tiles; the wait for the NEXT tap's weight stage sits between the sixth and the seventh, where the issuing
          // thread is held back by the tensor pipe anyway, so the first MMA of the next tap follows the last of this one.
          // (No tcgen05 fence here: the operands arrive through the async proxy and their mbarrier; the fence after the
          // accumulator hand-over above is the one that orders against the epilogue's tcgen05.ld.)
          if (elect_one()) {
            int w = wst, ph = wph;
            if (timing) tq = clock64();
            mbar_wait(full_w(w), ph);
            if (timing) wt_w += clock64() - tq;
#define NEF_PS_TAPS(MMA, IDESC)                                                                                      \
            for (int tp = 0; tp < taps; ++tp, ++xa) {                                                                \
              const uint32_t wa = wlo0 + (uint32_t)w * (FW_WBYTES >> 4);                                             \
              const uint32_t ebar = empty_w(w);                                                                      \
              if (++w == S::WST) { w = 0; ph ^= 1; }                                                                 \
              _Pragma("unroll") for (int k8 = 0; k8 < 4; ++k8)                                                       \
                MMA(acc0, desc_of(DESC_HI_SBO128, xa + k8 * XK), desc_of(DESC_HI_SBO128, wa + k8 * nb), IDESC, accum | (uint32_t)k8); \
              _Pragma("unroll") for (int k8 = 0; k8 < 2; ++k8)                                                       \
                MMA(acc0 + N, desc_of(DESC_HI_SBO128, xa + k8 * XK + 128), desc_of(DESC_HI_SBO128, wa + k8 * nb), IDESC, accum | (uint32_t)k8); \
              if (tp + 1 < taps) {                                                                                   \
                if (timing) tq = clock64();                                                                          \
                mbar_wait(full_w(w), ph);                                                                            \
                if (timing) wt_w += clock64() - tq;                                                                  \
              }                                                                                                      \
              _Pragma("unroll") for (int k8 = 2; k8 < 4; ++k8)                                                       \
                MMA(acc0 + N, desc_of(DESC_HI_SBO128, xa + k8 * XK + 128), desc_of(DESC_HI_SBO128, wa + k8 * nb), IDESC, 1u); \
              tc_commit(ebar);                                                                                       \
              accum = 1;                                                                                             \
            }
            if (f16) { NEF_PS_TAPS(mma_f16, idesc16) } else { NEF_PS_TAPS(mma_tf32, idesc) }
#undef NEF_PS_TAPS
            tc_commit(empty_x(xs));
          }
          __syncwarp();
          {  // every lane keeps the stage counters (the elected lane advanced its private copies)
            const int total = wst + taps;
            wph ^= (total / S::WST) & 1;
            wst = total % S::WST;
            accum = 1;
          }
          if (++xs == S::XST) { xs = 0; xph ^= 1; }
        }
      }
      if (elect_one()) tc_commit(acc_full(as));
      __syncwarp();
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (lane == 0 && blockIdx.x < 1024) {  // profiling aid: cycles the issuer spent waiting (nef_tc_debug_dump)
      g_tc_dbg[blockIdx.x][0] = (unsigned long long)(clock64() - t_begin);
      g_tc_dbg[blockIdx.x][1] = (unsigned long long)wt_a;
      g_tc_dbg[blockIdx.x][2] = (unsigned long long)wt_x;
      g_tc_dbg[blockIdx.x][3] = (unsigned long long)wt_w;
    }
  } else {
    // ===== epilogue warps =====
    int as = 0, aph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int g = tile / tiles_per_group;
      const long r0 = (long)(tile - g * tiles_per_group) * (MT * 128);
      if (!(use_ws & 16)) epilogue_prefetch<MT, EPI, EW>(d, g, r0, warp, lane);   // (NEF_TC_WS bit 4: A/B switch)
      mbar_wait(acc_full(as), aph);
      tc_fence_after();
      if (!(use_ws & 2))   // (NEF_TC_WS bit 1: timing experiment without the epilogue)
        epilogue_tile<MT, EPI, EW>(d, g, r0, tmem + (uint32_t)(as * 256), warp, lane, tid, s_stat);
      tc_fence_before();
      mbar_arrive(acc_empty(as));
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// weight-gradient kernel
//
// The contraction runs over rows, so both operands would be "MN-major" in the CBL4 layout -- and
// tcgen05.mma kind::tf32 returns zeros for MN-major operands on this part (tools/probe_umma.cu,
// profiles/r01_umma_tf32_mn_major_probe.txt; 16-bit kinds transpose fine, 32-bit ones do not).  The kernel
// therefore re-tiles each staged CBL4 tile in shared memory into K-major core matrices with the four
// epilogue warps (idle during the main loop) -- a register 4x4 transpose per (chunk, unit):
//     unit a of a 32-row stage = rows (a, a+8, a+16, a+24) of one channel = one 16-byte K group
// Because a K group strides over the stage instead of packing 4 adjacent rows, a tap shift of t rows is a
// shift of t UNITS (16 bytes * plane pitch), so one transposed copy of x serves every tap: the descriptor
// start address advances by one unit plane per tap, exactly as the forward kernel advances by one row.
//     D_tap[cout x cin] += dY^T[cout x (a,b)] . X^T[cin x (a + tap, b)]
// ---------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 416;  // warps 0, 10, 11, 12: copy producers; warp 1: MMA issuer; warps 2..9: re-tiling + epilogue
constexpr int WG_RT = 256;       // re-tiling threads
constexpr int WG_KR = 32;                        // rows (= contraction length) per re-tiled sub-stage
constexpr int WG_S = WG_KR / 4;                  // units per sub-stage; unit a = rows a + WG_S * b
constexpr int WG_RROWS = 64;                     // rows per raw (copy) stage = 2 sub-stages: 1 KB contiguous per chunk from HBM
constexpr int WG_YPITCH = WG_RROWS * 16;         // raw tiles: bytes between channel chunks
constexpr int WG_XPITCH = (WG_RROWS + 8) * 16;
constexpr int WG_RAW = 32 * WG_YPITCH + 32 * WG_XPITCH;   // one raw stage (128 + 128 channels)
constexpr int WG_NRAW = 2;
constexpr int WG_LBO = 128 * 16 + 16;            // transposed tiles: bytes between unit planes (+16: conflict-free stores)
constexpr int WG_TY = WG_S * WG_LBO;             // dY^T: units [0, 8)
constexpr int WG_XUNITS = WG_S + 6;              // X^T: units [0, 8 + taps - 1)
constexpr int WG_TR = WG_TY + WG_XUNITS * WG_LBO;
constexpr int WG_NTR = 2;
constexpr int WG_BAR_OFF = WG_NRAW * WG_RAW + WG_NTR * WG_TR;
constexpr int WG_TOTAL = WG_BAR_OFF + 128;
static_assert(WG_TOTAL <= 227 * 1024, "wgrad shared memory");

// Every CTA accumulates ALL taps (<= 7) of a 128 x 64 block of the weight gradient (7 x 64 = 448 TMEM columns), so
// each row of X and dY is staged once per CTA.  Two orientations:
//  swap == 0 (cout_g == 128, or 64 x 64):  D_tap[cout x cin_half] = dY^T . X(+tap)   M = cout (64 is zero-padded to 128),
//     N = 64 input channels; the taps of a K step share the A operand (dY^T): collector::a fill / use / lastuse, so the
//     MMAs are not bound by re-reading A from shared memory (N = 64 alone would be: tools/probe_rate.cu, 48 vs 32 cycles)
//  swap == 1 (cout_g == 64, cin_g >= 128): D_tap[cin_tile x cout] = X^T(+tap) . dY   M = 128 input channels, N = 64; the
//     shared operand is B: weight-stationary form (collector::b0)
//  wide (mode 2; cout_g == 128, cin_g % 128 == 0, <= 3 taps): as swap == 0 with N = 128 input channels per CTA
//     (3 x 128 = 384 TMEM columns), so dY is staged and re-tiled once per 128 input channels instead of once per 64
// CTAs that take different input-channel tiles do identical work on the same rows of the shared operand (L2 serves it).
// grid: x = input-channel tile (swap == 0), y = row split, z = group (x cin tile of 128 when swap == 1)
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ NefWgradDesc d, int NT, long rows_per_split,
                                                                 long rows_main, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tr0 = sbase + WG_NRAW * WG_RAW;
  const uint32_t bar0 = sbase + WG_BAR_OFF;
  auto raw_full = [&](int i) { return bar0 + 8 * i; };
  auto raw_empty = [&](int i) { return bar0 + 8 * (WG_NRAW + i); };
  auto tr_full = [&](int i) { return bar0 + 8 * (2 * WG_NRAW + i); };
  auto tr_empty = [&](int i) { return bar0 + 8 * (2 * WG_NRAW + WG_NTR + i); };
  const uint32_t acc_full = bar0 + 8 * (2 * WG_NRAW + 2 * WG_NTR);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + WG_BAR_OFF + 8 * (2 * WG_NRAW + 2 * WG_NTR + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int swap = mode == 1, wide = mode == 2;
  const int ntile = swap ? d.cin_g / 128 : 1;
  const int g = blockIdx.z / ntile, nt = blockIdx.z % ntile;
  const int ntap = d.taps;
  const int half = blockIdx.x;                       // swap == 0: input channels [64 half, 64 half + 64)
  const int ych = d.cout_g >> 2;                     // dY chunks staged (all output channels of the group)
  const int xch = (swap || wide) ? 32 : 16;          // X chunks staged
  const int xch0 = swap ? nt * 32 : half * xch;      // first X chunk staged
  const int dcols = wide ? 128 : 64;                 // accumulator columns per tap
  (void)NT;
  const long rbeg = (long)blockIdx.y * rows_per_split;
  const long rend = min(rows_main, rbeg + rows_per_split);
  const int nstage = (int)((rend - rbeg) / WG_RROWS);
  const int xunits = WG_S + ntap - 1;
  const uint32_t TM_COLS = 512;

  if (tid == 0) {
    for (int i = 0; i < WG_NRAW; ++i) { mbar_init(raw_full(i), 4); mbar_init(raw_empty(i), WG_RT); }
    for (int i = 0; i < WG_NTR; ++i) { mbar_init(tr_full(i), WG_RT); mbar_init(tr_empty(i), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (!swap && ych < 32) {  // cout_g == 64: channels 64..127 of the M = 128 operand are zero (never written afterwards)
    for (int u = 0; u < WG_NTR; ++u)
      for (int i = tid; i < WG_S * 64; i += WG_THREADS) {
        const int a = i >> 6, c = 64 + (i & 63);
        *reinterpret_cast<float4*>(smem + WG_NRAW * WG_RAW + u * WG_TR + a * WG_LBO + c * 16) = f4zero();
      }
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 || warp >= 10) {
    // ===== four copy-producer warps (a bulk copy is issued from the uniform datapath, one at a time per warp):
    //       producer pw stages chunks [8 pw, 8 pw + 8) of dY (lanes 0..7) and of X (lanes 8..15) =====
    const int pw = warp == 0 ? 0 : warp - 9;
    const float4* yg = reinterpret_cast<const float4*>(d.dy) + (long)(d.dy_c4_off + g * d.dy_c4_gstride) * d.dy_cstride;
    const float4* xg = reinterpret_cast<const float4*>(d.x) + (long)(d.x_c4_off + g * d.x_c4_gstride + xch0) * d.x_cstride + d.tap_off;
    const uint32_t xbytes = (uint32_t)(WG_RROWS + ntap - 1) * 16;
    const int c = pw * 8 + (lane & 7);
    const int ny = min(max(ych - pw * 8, 0), 8), nx = min(max(xch - pw * 8, 0), 8);
    int st = 0, ph = 0;
    for (int it = 0; it < nstage; ++it) {
      const long r = rbeg + (long)it * WG_RROWS;
      if (lane == 0) {
        mbar_wait(raw_empty(st), ph ^ 1);
        mbar_expect_tx(raw_full(st), (uint32_t)ny * WG_YPITCH + (uint32_t)nx * xbytes);
      }
      __syncwarp();
      const uint32_t ys = sbase + st * WG_RAW, xs = ys + 32 * WG_YPITCH;
      if (lane < 8) {
        if (c < ych) bulk_g2s(ys + c * WG_YPITCH, yg + (long)c * d.dy_cstride + r, WG_YPITCH, raw_full(st));
      } else if (lane < 16) {
        if (c < xch) bulk_g2s(xs + c * WG_XPITCH, xg + (long)c * d.x_cstride + r, xbytes, raw_full(st));
      }
      if (++st == WG_NRAW) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues (see the forward kernel) =====
    const uint32_t idesc = make_idesc(128, dcols, 0, 0);
    constexpr uint32_t UNIT = WG_LBO >> 4;  // descriptor units per K group plane
    int u = 0, ph = 0;
    for (int it = 0; it < 2 * nstage; ++it) {
      mbar_wait(tr_full(u), ph);
      tc_fence_after();
      const uint32_t ty = desc_lo(tr0 + u * WG_TR, WG_LBO), tx = desc_lo(tr0 + u * WG_TR + WG_TY, WG_LBO);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < WG_S / 2; ++ks) {
          const uint64_t yd = desc_of(DESC_HI_SBO128, ty + 2 * ks * UNIT);
          const uint32_t xk = tx + 2 * ks * UNIT;
          const uint32_t acc = (uint32_t)(it | ks);
          if (ntap == 1) {
            if (!swap) mma_tf32(tmem, yd, desc_of(DESC_HI_SBO128, xk), idesc, acc);
            else mma_tf32(tmem, desc_of(DESC_HI_SBO128, xk), yd, idesc, acc);
          } else if (!swap) {
            mma_tf32_areuse<0>(tmem, yd, desc_of(DESC_HI_SBO128, xk), idesc, acc);
#pragma unroll
            for (int tp = 1; tp < 6; ++tp)
              if (tp < ntap - 1) mma_tf32_areuse<1>(tmem + tp * dcols, yd, desc_of(DESC_HI_SBO128, xk + tp * UNIT), idesc, acc);
            mma_tf32_areuse<2>(tmem + (ntap - 1) * dcols, yd, desc_of(DESC_HI_SBO128, xk + (ntap - 1) * UNIT), idesc, acc);
          } else {
            mma_tf32_ws<0>(tmem, desc_of(DESC_HI_SBO128, xk), yd, idesc, acc);
#pragma unroll
            for (int tp = 1; tp < 6; ++tp)
              if (tp < ntap - 1) mma_tf32_ws<1>(tmem + tp * dcols, desc_of(DESC_HI_SBO128, xk + tp * UNIT), yd, idesc, acc);
            mma_tf32_ws<2>(tmem + (ntap - 1) * dcols, desc_of(DESC_HI_SBO128, xk + (ntap - 1) * UNIT), yd, idesc, acc);
          }
        }
        tc_commit(tr_empty(u));
      }
      __syncwarp();
      if (++u == WG_NTR) { u = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
  } else if (warp < 10) {
    // ===== warps 2..9: re-tile every sub-stage into K-major core matrices, then drain the accumulators =====
    const int e = tid - 64;  // 0..255
    {
      // the (chunk, unit) blocks of a thread are the same in every sub-stage: byte offsets precomputed, -1 = none
      int ysrc = -1, ydst = 0, xsrc[2] = {-1, -1}, xdst[2] = {0, 0};
      if (e < ych * WG_S) {
        const int c = e >> 3, a = e & 7;
        ysrc = c * WG_YPITCH + a * 16;
        ydst = a * WG_LBO + (4 * c) * 16;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int id = e + k * WG_RT;
        if (id < xch * xunits) {
          const int c = id / xunits, a = id - c * xunits;
          xsrc[k] = 32 * WG_YPITCH + c * WG_XPITCH + a * 16;
          xdst[k] = WG_TY + a * WG_LBO + (4 * c) * 16;
        }
      }
      int st = 0, rph = 0, u = 0, uph = 0;
      for (int it = 0; it < nstage; ++it) {
        mbar_wait(raw_full(st), rph);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const uint8_t* raw = smem + st * WG_RAW + sub * (WG_KR * 16);
          // all loads of the sub-stage first (up to 12 x 16 bytes in flight per thread), then the transposed stores
          float4 r[3][4];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int so = k == 0 ? ysrc : xsrc[k - 1];
            if (so >= 0) {
              const float4* src = reinterpret_cast<const float4*>(raw + so);
              r[k][0] = src[0]; r[k][1] = src[WG_S]; r[k][2] = src[2 * WG_S]; r[k][3] = src[3 * WG_S];
            }
          }
          mbar_wait(tr_empty(u), uph ^ 1);
          tc_fence_after();
          uint8_t* tr = smem + WG_NRAW * WG_RAW + u * WG_TR;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int so = k == 0 ? ysrc : xsrc[k - 1];
            if (so >= 0) {
              float4* dst = reinterpret_cast<float4*>(tr + (k == 0 ? ydst : xdst[k - 1]));
              dst[0] = make_float4(r[k][0].x, r[k][1].x, r[k][2].x, r[k][3].x);
              dst[1] = make_float4(r[k][0].y, r[k][1].y, r[k][2].y, r[k][3].y);
              dst[2] = make_float4(r[k][0].z, r[k][1].z, r[k][2].z, r[k][3].z);
              dst[3] = make_float4(r[k][0].w, r[k][1].w, r[k][2].w, r[k][3].w);
            }
          }
          fence_proxy_async();
          mbar_arrive(tr_full(u));
          if (sub == 1) mbar_arrive(raw_empty(st));   // both halves of the raw stage have been re-tiled
          if (++u == WG_NTR) { u = 0; uph ^= 1; }
        }
        if (++st == WG_NRAW) { st = 0; rph ^= 1; }
      }
    }
    if (nstage > 0) {
      const int q = warp & 3;
      const int lr = q * 32 + lane;  // accumulator row: output channel (cout) -- or input channel of this tile when swapped
      const int chalf = (warp - 2) >> 2;  // the two warps of a TMEM lane quarter take alternate 32-column groups
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int gw = d.wg_mod > 0 ? g % d.wg_mod : g;   // shared weights: group g accumulates into slot g mod wg_mod
      for (int tp = 0; tp < ntap; ++tp) {
        for (int cg = chalf; cg < dcols / 32; cg += 2) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * dcols + cg * 32), v);
          tmem_ld_wait();
          if (!swap) {
            if (lr < d.cout_g) {
              float* dst = d.dw + (long)gw * d.sg + (long)lr * d.sm + (long)(half * dcols + cg * 32) * d.sn + (long)tp * d.st;
#pragma unroll
              for (int i = 0; i < 32; ++i) atomicAdd(dst + (long)i * d.sn, __uint_as_float(v[i]));
            }
          } else {
            float* dst = d.dw + (long)gw * d.sg + (long)(cg * 32) * d.sm + (long)(nt * 128 + lr) * d.sn + (long)tp * d.st;
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(dst + (long)i * d.sm, __uint_as_float(v[i]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

// db[g * cout_g + m] += sum over rows of dy[row][g, m]   (grid: x = row splits, y = 4-channel chunks of all groups)
__global__ void __launch_bounds__(256) bias_grad_kernel(const NefWgradDesc d, long rows_per_split) {
  __shared__ float4 red[8];
  const int ch4 = blockIdx.y;  // chunk among groups * cout_g / 4
  const int g = ch4 / (d.cout_g >> 2), c = ch4 % (d.cout_g >> 2);
  const float4* yg = reinterpret_cast<const float4*>(d.dy) + (long)(d.dy_c4_off + g * d.dy_c4_gstride + c) * d.dy_cstride;
  const long rbeg = (long)blockIdx.x * rows_per_split, rend = min(d.rows, rbeg + rows_per_split);
  float4 acc = f4zero();
  for (long r = rbeg + threadIdx.x; r < rend; r += 256) acc = acc + __ldg(yg + r);
  acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y); acc.z = warp_sum(acc.z); acc.w = warp_sum(acc.w);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = f4zero();
    for (int w = 0; w < 8; ++w) t = t + red[w];
    float* o = d.db + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.cout_g + c * 4;
    atomicAdd(o + 0, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
  }
}

// the same from an fp16 gradient copy (half8 rows: 8 channels per 16-byte row, times a loss scale): db += scale[0] * sum
__global__ void __launch_bounds__(256) bias_grad_h_kernel(const NefWgradDesc d, const uint4* __restrict__ dy16, const float* __restrict__ scale,
                                                          long rows_per_split) {
  __shared__ float red[8][8];
  const int ch8 = blockIdx.y;  // chunk among groups * cout_g / 8
  const int g = ch8 / (d.cout_g >> 3), c = ch8 % (d.cout_g >> 3);
  const uint4* yg = dy16 + (long)((d.dy_c4_off >> 1) + g * (d.dy_c4_gstride >> 1) + c) * d.dy_cstride;
  const long rbeg = (long)blockIdx.x * rows_per_split, rend = min(d.rows, rbeg + rows_per_split);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long r = rbeg + threadIdx.x; r < rend; r += 256) {
    const uint4 h = __ldg(yg + r);
    const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&h.z)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&h.w));
    acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
    acc[4] += b0.x; acc[5] += b0.y; acc[6] += b1.x; acc[7] += b1.y;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(d.db + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.cout_g + c * 8 + threadIdx.x, t * (scale ? __ldg(scale) : 1.f));
  }
}

}  // namespace tc
}  // namespace nef

using namespace nef;

static int g_sm_count = 148;
static int g_tc_stagger = -1;  // first-wave start stagger in cycles; -1 = one estimated CTA lifetime (NEF_TC_STAGGER)
static int g_tc_ws = 0;   // 1 = weight-stationary MMA form (NEF_TC_WS)
static int g_tc_persist = 1;  // persistent forward kernel (NEF_TC_PERSIST=0: one CTA per tile)
static int g_tc_persist_min = 0;  // row-tile x group count from which the persistent kernel is used; 0 = 2 x SM count
extern "C" int nef_tc_set_persist_min(int tiles) { g_tc_persist_min = tiles; return 0; }  // test hook: 1 = always persistent
static int g_tc_heavy_mmas = 96;  // MMAs per 256-row tile from which a launch counts as tensor-bound (8 epilogue warps); k7 x 128 ch = 112
static int g_tc_mt = 0;  // 0 = automatic; 4 forces four row tiles per CTA, one CTA per SM (NEF_TC_MT, for A/B measurements)

// the specialised epilogues instantiated for the 4-row-tile kernel (everything else takes the generic one)
#define NEF_TC_EPI_LIST(X) X(0) X(2) X(5) X(32) X(33) X(36) X(37) X(38) X(44) X(48) X(50) X(64) X(294) X(418) X(513) \
  X(1062) X(1068) X(1318) X(2080) X(2082) X(2338) X(4132) X(4134) X(4390) X(5158) X(5164) X(5414) X(8194) X(14368) X(14370) X(14626) X(21540) X(21796) X(24576) X(30752) X(31008) \
  X(37932) X(54308) X(54564) X(47136) X(63520) X(63776) X(45056) X(8224) X(61440) X(36865) X(36869)   /* the production fp16-only stores: 5164, 21540, 21796, 14368, 30752, 31008 | EPI_NOY */

// the epilogues of the k7 / 128-channel fp16 layers (forward with dropout, second convolutions with the fp16 residual, masked
// and unmasked loss-scaled data gradients): instantiated with 8 epilogue warps too
#define NEF_TC_EPI_HEAVY(X) X(37932) X(54308) X(54564) X(47136) X(63520) X(24576) X(61440)

template <int MT, int EPI>
static int tc_optin() {
  cudaError_t e = cudaFuncSetAttribute(tc::conv_tc_kernel<MT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::FwSmem<MT>::TOTAL);
  NEF_REQUIRE(e == cudaSuccess, "nef_tc_init: conv_tc_kernel<%d,%d> shared-memory opt-in failed: %s", MT, EPI, cudaGetErrorString(e));
  return 0;
}

extern "C" int nef_tc_init(void) {
#define X(E) { int rc = tc_optin<4, E>(); if (rc) return rc; rc = tc_optin<2, E>(); if (rc) return rc; \
               cudaError_t pe = cudaFuncSetAttribute(tc::conv_tc_persist_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::PsSmem::TOTAL); \
               NEF_REQUIRE(pe == cudaSuccess, "nef_tc_init: conv_tc_persist_kernel<%d> shared-memory opt-in failed: %s", E, cudaGetErrorString(pe)); }
  NEF_TC_EPI_LIST(X)
#undef X
#define X(E) { cudaError_t pe = cudaFuncSetAttribute(tc::conv_tc_persist_kernel<E, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::PsSmem::TOTAL); \
               NEF_REQUIRE(pe == cudaSuccess, "nef_tc_init: conv_tc_persist_kernel<%d, 8> shared-memory opt-in failed: %s", E, cudaGetErrorString(pe)); }
  NEF_TC_EPI_HEAVY(X)
#undef X
  if (getenv("NEF_TC_HEAVY_MMAS")) g_tc_heavy_mmas = atoi(getenv("NEF_TC_HEAVY_MMAS"));
  if (getenv("NEF_TC_MT")) g_tc_mt = atoi(getenv("NEF_TC_MT"));
  if (getenv("NEF_TC_PERSIST")) g_tc_persist = atoi(getenv("NEF_TC_PERSIST"));
  if (getenv("NEF_TC_PERSIST_MIN")) g_tc_persist_min = atoi(getenv("NEF_TC_PERSIST_MIN"));   // 1: persistent kernel at every size (sanitizer runs)
  if (getenv("NEF_TC_STAGGER")) g_tc_stagger = atoi(getenv("NEF_TC_STAGGER"));
  if (getenv("NEF_TC_WS")) g_tc_ws = atoi(getenv("NEF_TC_WS"));
  { int rc = tc_optin<1, tc::EPI_GENERIC>(); if (rc) return rc; }
  cudaError_t e = cudaFuncSetAttribute(tc::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::WG_TOTAL);
  NEF_REQUIRE(e == cudaSuccess, "nef_tc_init: wgrad_tc_kernel smem opt-in failed: %s", cudaGetErrorString(e));
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  return 0;
}

// Dispatch record (test hook, nef_tc_dispatch_stats): which kernel forms and which specialised epilogues have been launched, so
// that a parity test can assert that it exercised the dispatch the benchmark times.
static long long g_disp[8];           // 0 persistent, 1 one-CTA-per-tile (MT 2 / 4), 2 single-tile generic, 3 wgrad_tc, 4 wgrad simt tail, 5 wgrad_f16
static int g_disp_epi[64], g_disp_nepi = 0;   // distinct EPI codes launched through the persistent kernel
static void note_persist_epi(int e) {
  for (int i = 0; i < g_disp_nepi; ++i)
    if (g_disp_epi[i] == e) return;
  if (g_disp_nepi < 64) g_disp_epi[g_disp_nepi++] = e;
}
extern "C" void nef_tc_note_dispatch(int which) { __atomic_add_fetch(&g_disp[which & 7], 1, __ATOMIC_RELAXED); }
extern "C" int nef_tc_dispatch_stats(int64_t* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = -1;
  for (int i = 0; i < 8 && i < n; ++i) out[i] = g_disp[i];
  for (int i = 0; i < g_disp_nepi && 8 + i < n; ++i) out[8 + i] = g_disp_epi[i];
  return g_disp_nepi;
}
extern "C" void nef_tc_dispatch_reset(void) {
  for (int i = 0; i < 8; ++i) g_disp[i] = 0;
  g_disp_nepi = 0;
}

static int epi_code(const NefConvDesc* d) {
  if (d->bscale_grad) return tc::EPI_GENERIC;
  int e = 0;
  if (d->stat_sum) e |= tc::EPI_STATS;
  if (d->bscale) e |= tc::EPI_BSCALE;
  const bool mbits = d->mask_bits != nullptr && d->mask_mode != 0;
  if (d->mask_mode == 2 && !mbits) e |= tc::EPI_MASK2;
  if (mbits) e |= tc::EPI_MBITS;
  if (d->out_bits) e |= tc::EPI_OBITS;
  if (d->y16) e |= tc::EPI_Y16;
  if (d->acc_scale || d->y16_scale) e |= tc::EPI_GSCALE;
  if (d->bias) e |= tc::EPI_BIAS;
  if (d->res) e |= tc::EPI_RES;
  if (d->res16) e |= tc::EPI_RES16;
  if (d->relu) e |= tc::EPI_RELU;
  if (d->drop_p > 0.f) e |= tc::EPI_DROP;
  if (d->mask_mode == 1 && !mbits) e |= tc::EPI_MASK1;
  if (d->round_tf32) e |= tc::EPI_ROUND;
  if (!d->y) e |= tc::EPI_NOY;
  return e;
}


template <int MT>
static int launch_conv_tc(const NefConvDesc* d, cudaStream_t s) {
  dim3 grid((unsigned)((d->rows + MT * 128 - 1) / (MT * 128)), (unsigned)d->groups);
  const int fw = g_sm_count * (MT == 2 ? 2 : 1);
  // one CTA lifetime ~ (cycles of MMA work per row tile) * MT + epilogue; the stagger spans about one lifetime
  long kw = 0;
  for (int i = 0; i < d->n_terms; ++i) kw += (long)(d->term[i].cin_g / 8) * d->term[i].taps;
  const int stg = MT != 2 ? 0 : (g_tc_stagger < 0 ? (int)((kw * (d->N / 2) * 1.3 + 14000) * MT / 2) : g_tc_stagger);
  switch (epi_code(d)) {
#define X(E) case E: tc::conv_tc_kernel<MT, E><<<grid, tc::FW_THREADS, tc::FwSmem<MT>::TOTAL, s>>>(*d, fw, stg, g_tc_ws); break;
    NEF_TC_EPI_LIST(X)
#undef X
    default: tc::conv_tc_kernel<MT, tc::EPI_GENERIC><<<grid, tc::FW_THREADS, tc::FwSmem<MT>::TOTAL, s>>>(*d, fw, stg, g_tc_ws); break;
  }
  return 0;
}

static int launch_conv_persist(const NefConvDesc* d, cudaStream_t s) {
  const int tpg = (int)((d->rows + 255) / 256);
  const int n_tiles = tpg * d->groups;
  const int grid = n_tiles < g_sm_count ? n_tiles : g_sm_count;
  bool specialised = false;
  switch (epi_code(d)) {
#define X(E) case E: specialised = true; break;
    NEF_TC_EPI_LIST(X)
#undef X
    default: break;
  }
  nef_tc_note_dispatch(0);
  note_persist_epi(specialised ? epi_code(d) : tc::EPI_GENERIC);
  int mmas = 0;   // MMAs per 256-row tile
  for (int i = 0; i < d->n_terms; ++i) mmas += (d->term[i].cin_g >> 5) * d->term[i].taps * 8;
  if (mmas >= g_tc_heavy_mmas) {
    switch (epi_code(d)) {
#define X(E) case E: tc::conv_tc_persist_kernel<E, 8><<<grid, 64 + 32 * 8, tc::PsSmem::TOTAL, s>>>(*d, tpg, n_tiles, g_tc_ws); return 0;
      NEF_TC_EPI_HEAVY(X)
#undef X
      default: break;
    }
  }
  switch (epi_code(d)) {
#define X(E) case E: tc::conv_tc_persist_kernel<E><<<grid, 64 + 32 * tc::ps_epi_warps<E>(), tc::PsSmem::TOTAL, s>>>(*d, tpg, n_tiles, g_tc_ws); break;
    NEF_TC_EPI_LIST(X)
#undef X
    default: tc::conv_tc_persist_kernel<tc::EPI_GENERIC><<<grid, tc::FW_THREADS, tc::PsSmem::TOTAL, s>>>(*d, tpg, n_tiles, g_tc_ws); break;
  }
  return 0;
}

extern "C" int nef_gconv_fwd_tc(const NefConvDesc* d, nef_stream_t s) {
  const long tiles4 = (d->rows + 511) / 512, tiles2 = (d->rows + 255) / 256;
  const long persist_min = g_tc_persist_min > 0 ? g_tc_persist_min : 2L * g_sm_count;
  if (g_tc_persist && g_tc_mt == 0 && tiles2 * d->groups >= persist_min) {
    launch_conv_persist(d, (cudaStream_t)s);
    NEF_CHECK_LAUNCH("conv_tc_persist_kernel");
    return 0;
  }
  int mt = 1;  // small row spaces (the z2 deflection branch at small batch): one 128-row tile per CTA keeps the grid wide
  if (tiles2 * d->groups >= 2L * g_sm_count) mt = 2;
  if (g_tc_mt == 4 && tiles4 * d->groups >= g_sm_count) mt = 4;
  nef_tc_note_dispatch(mt == 1 ? 2 : 1);
  if (mt == 2) launch_conv_tc<2>(d, (cudaStream_t)s);
  else if (mt == 4) launch_conv_tc<4>(d, (cudaStream_t)s);
  else {
    dim3 grid((unsigned)((d->rows + 127) / 128), (unsigned)d->groups);
    tc::conv_tc_kernel<1, tc::EPI_GENERIC><<<grid, tc::FW_THREADS, tc::FwSmem<1>::TOTAL, (cudaStream_t)s>>>(*d, 0, 0, 0);
  }
  NEF_CHECK_LAUNCH("conv_tc_kernel");
  return 0;
}

// profiling aid: copies the phase timestamps (ns, %globaltimer) of the first n CTAs of the last conv launch
extern "C" int nef_tc_debug_dump(unsigned long long* host_out, int n) {
  if (n > 1024) n = 1024;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, tc::g_tc_dbg, sizeof(unsigned long long) * 8 * n) == cudaSuccess ? 0 : 1;
}

static int g_wg_wide = 1;
extern "C" int nef_tc_set_wgrad_wide(int on) { g_wg_wide = on; return 0; }  // A/B measurement hook

extern "C" int nef_gconv_wgrad_tc(const NefWgradDesc* d, nef_stream_t s) {
  NEF_REQUIRE(d->cout_g == 64 || d->cout_g == 128, "nef_gconv_wgrad_tc: cout_g must be 64 or 128 (got %d)", d->cout_g);
  NEF_REQUIRE(d->cin_g % 64 == 0 && d->taps <= 7, "nef_gconv_wgrad_tc: cin_g %% 64 == 0 and at most 7 taps required (cin_g=%d taps=%d)",
              d->cin_g, d->taps);
  const int swap = (d->cout_g == 64 && d->cin_g % 128 == 0) ? 1 : 0;
  const int wide = (!swap && g_wg_wide && d->cout_g == 128 && d->cin_g % 128 == 0 && d->taps <= 3) ? 1 : 0;
  const int NT = (swap || wide) ? 128 : 64;
  const long nst_total = d->rows / tc::WG_RROWS;
  const long rows_main = nst_total * tc::WG_RROWS;
  if (nst_total > 0) {
    const int halves = swap ? 1 : d->cin_g / (wide ? 128 : 64);
    const int ztiles = swap ? d->groups * (d->cin_g / 128) : d->groups;
    const long tiles = (long)ztiles * halves;
    // row splits: the smallest count whose last wave is >= 90 % full (else the best seen)
    long best = 1;
    double best_eff = 0.0;
    const long max_splits = nst_total < 2L * g_sm_count ? nst_total : 2L * g_sm_count;
    for (long sp = 1; sp <= max_splits; ++sp) {
      const long ctas = tiles * sp;
      const double eff = (double)ctas / (double)(((ctas + g_sm_count - 1) / g_sm_count) * g_sm_count);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
      if (eff >= 0.9) { best = sp; break; }
    }
    long st_per_split = (nst_total + best - 1) / best;
    const long splits = (nst_total + st_per_split - 1) / st_per_split;
    dim3 grid((unsigned)halves, (unsigned)splits, (unsigned)ztiles);
    tc::wgrad_tc_kernel<<<grid, tc::WG_THREADS, tc::WG_TOTAL, (cudaStream_t)s>>>(*d, NT, st_per_split * tc::WG_RROWS, rows_main, swap ? 1 : (wide ? 2 : 0));
    NEF_CHECK_LAUNCH("wgrad_tc_kernel");
    nef_tc_note_dispatch(3);
  }
  if (rows_main < d->rows) {  // ragged tail (< 64 rows): CUDA-core kernel over [rows_main, rows), no bias term
    NefWgradDesc t = *d;
    t.db = nullptr;
    int rc = nef_gconv_wgrad_simt_range(&t, rows_main, s);
    if (rc) return rc;
    nef_tc_note_dispatch(4);
  }
  if (d->db) {
    long splits = (d->rows + 4095) / 4096;
    if (splits > 64) splits = 64;
    const long rps = (d->rows + splits - 1) / splits;
    dim3 grid((unsigned)((d->rows + rps - 1) / rps), (unsigned)(d->groups * (d->cout_g / 4)));
    tc::bias_grad_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps);
    NEF_CHECK_LAUNCH("bias_grad_kernel");
  }
  return 0;
}

// db[g * cout_g + m] += sum over rows of dy[row][g, m] alone (the weight gradient of the layer ran on fp16 copies)
extern "C" int nef_bias_grad_tc(const NefWgradDesc* d, nef_stream_t s) {
  long splits = (d->rows + 4095) / 4096;
  if (splits > 64) splits = 64;
  const long rps = (d->rows + splits - 1) / splits;
  dim3 grid((unsigned)((d->rows + rps - 1) / rps), (unsigned)(d->groups * (d->cout_g / 4)));
  tc::bias_grad_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(*d, rps);
  NEF_CHECK_LAUNCH("bias_grad_kernel");
  return 0;
}

// the same from the loss-scaled fp16 copy dy16 of the gradient (geometry of d; d->dy is not read); scale[0] = 1 / S (device)
extern "C" int nef_bias_grad_h(const NefWgradDesc* d, const void* dy16, const float* scale, nef_stream_t s) {
  long splits = (d->rows + 4095) / 4096;
  if (splits > 64) splits = 64;
  const long rps = (d->rows + splits - 1) / splits;
  dim3 grid((unsigned)((d->rows + rps - 1) / rps), (unsigned)(d->groups * (d->cout_g / 8)));
  tc::bias_grad_h_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(*d, reinterpret_cast<const uint4*>(dy16), scale, rps);
  NEF_CHECK_LAUNCH("bias_grad_h_kernel");
  return 0;
}

NEF_DEFINE_EXACT_SETTER(nef_set_exact_tc)
