// nef_gconv_wgrad_f16 (include/nefnet_b200.h): op-level parity in tests/test_gpu_conv_ops.py, used by nef_backward for the
// big 128-output-channel layers when the fp16 backward is on (nef_set_bwd_f16).
//
// Weight gradient of a grouped k-tap convolution from fp16 operand copies, WITHOUT a re-tile pass:
//     D_tap[cout x cin] += dY16[rows x cout]^T . X16[rows (+tap) x cin]          (fp32 accumulation in TMEM)
// The production kernel (wgrad_tc_kernel) transposes every staged tile in shared memory because tcgen05.mma kind::tf32
// reads zeros for MN-major operands; the 16-bit kinds accept them (tools/probe_umma16.cu).  In the fp16 layout
// `half8 T16[C/8][rows]` eight consecutive rows of one 8-channel chunk are 8 x 16 contiguous bytes = one MN-major no-swizzle
// core matrix (8 contraction steps x 8 channels), canonical form ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) with T = 8 halves:
//     LBO = 128 bytes (the next 8 rows), SBO = the staged chunk pitch,
// so both operands are consumed exactly as the bulk copy lands them; a tap is the B start address advanced by 16 bytes.
// All taps of a K step share the A operand (collector::a fill / use / lastuse), as in the production kernel.
//
// Geometry is the fp32 descriptor's (NefWgradDesc): same row space and chunk-plane pitch; chunk offsets / group strides are
// the 4-channel ones halved (they must be even).  Rows beyond the last full 128-row stage go through a small CUDA-core tail
// kernel on the same fp16 operands.  `out_scale` multiplies the accumulators before the fp32 RED (1 / loss scale).
#include <cuda_fp16.h>

#include "nef_conv.cuh"

namespace nef {
namespace wf16 {

// ---- PTX wrappers (same forms as in nef_conv_tc.cu) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A . B, fp16 operands, fp32 accumulate.  MODE: 0 = latch A in the collector (fill), 1 = re-use it, 2 = last use,
// 3 = no re-use (single-tap layers)
template <int MODE>
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (MODE == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (MODE == 1) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else if constexpr (MODE == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
                 "r"(accumulate) : "memory");
  }
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// descriptor words: lo = start address | LBO, hi = SBO | version 1 (SWIZZLE_NONE); only the start address moves between MMAs
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFF) | ((lbo_bytes >> 4) << 16); }
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
// Instruction descriptor: fp32 accumulate (bit 4), F16 x F16 (formats 0), BOTH operands MN-major (bits 15, 16), M x N
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int M, int N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int THREADS = 256;            // warp 0: MMA issuer; warps 1..3: copy producers; warps 4..7: accumulator drain
constexpr int NPROD = 3;
constexpr int ROWS = 128;               // rows (contraction length) per stage = 8 MMAs of K = 16 per tap
constexpr int YCH = 16;                 // 8-channel chunks of dY staged: all 128 output channels of the group
constexpr int YP = ROWS * 16;           // bytes between chunks of the staged dY tile
constexpr int XP = (ROWS + 8) * 16;     // ... of the staged X tile (ROWS + taps - 1 rows are used)
constexpr int YBYTES = YCH * YP;
constexpr int XBYTES = 16 * XP;         // up to 128 input channels per CTA
constexpr int STAGE = YBYTES + XBYTES;
constexpr int NST = 3;
constexpr int BAR_OFF = NST * STAGE;
constexpr int TOTAL = BAR_OFF + 128;
static_assert(TOTAL <= 227 * 1024, "wgrad_f16 shared memory");
static_assert(STAGE % 128 == 0 && YBYTES % 128 == 0, "stage alignment");

// grid: x = input-channel tile (dcols channels), y = row split, z = group
__global__ void __launch_bounds__(THREADS, 1) wgrad_f16_kernel(const __grid_constant__ NefWgradDesc d, const uint4* __restrict__ dy16,
                                                               const uint4* __restrict__ x16, const float* __restrict__ out_scale_p, int dcols,
                                                               long rows_per_split, long rows_main, int coalesced_drain) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + BAR_OFF;
  auto full = [&](int i) { return bar0 + 8 * i; };
  auto empty = [&](int i) { return bar0 + 8 * (NST + i); };
  const uint32_t acc_full = bar0 + 8 * (2 * NST);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * (2 * NST + 1));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.z, tile = blockIdx.x;
  const int ntap = d.taps;
  const int xch = dcols >> 3;   // 8-channel chunks of X staged
  // 8-channel chunks of dY staged.  cout_g = 64: the MMA still spans M = 128 (an M = 64 MMA runs at the same rate), its upper
  // 64 rows read the unused half of the staged dY region and their accumulator lanes are never drained
  const int ych = d.cout_g >> 3;
  const long rbeg = (long)blockIdx.y * rows_per_split;
  const long rend = min(rows_main, rbeg + rows_per_split);
  const int nstage = rend > rbeg ? (int)((rend - rbeg) / ROWS) : 0;
  const uint32_t TM_COLS = 512;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(full(i), NPROD); mbar_init(empty(i), 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32((const void*)tmem_slot), TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 1 && warp <= NPROD) {
    // ===== copy producers: copy id = pw + NPROD * lane; ids [0, ych) are dY chunks, [ych, ych + xch) are X chunks =====
    const int pw = warp - 1;
    const int id = pw + NPROD * lane;
    const int ncopy = ych + xch;
    const uint32_t xbytes = (uint32_t)(ROWS + ntap - 1) * 16;
    uint32_t my_bytes = 0;   // bytes this warp lands per stage (uniform)
    for (int l = 0; l < 32; ++l) {
      const int i = pw + NPROD * l;
      if (i < ych) my_bytes += YP;
      else if (i < ncopy) my_bytes += xbytes;
    }
    const uint4* src = nullptr;
    uint32_t dst_off = 0, bytes = 0;
    if (id < ych) {
      src = dy16 + (long)((d.dy_c4_off >> 1) + g * (d.dy_c4_gstride >> 1) + id) * d.dy_cstride;
      dst_off = (uint32_t)id * YP;
      bytes = YP;
    } else if (id < ncopy) {
      src = x16 + (long)((d.x_c4_off >> 1) + g * (d.x_c4_gstride >> 1) + tile * xch + (id - ych)) * d.x_cstride + d.tap_off;
      dst_off = YBYTES + (uint32_t)(id - ych) * XP;
      bytes = xbytes;
    }
    int st = 0, ph = 0;
    for (int it = 0; it < nstage; ++it) {
      const long r = rbeg + (long)it * ROWS;
      if (lane == 0) {
        mbar_wait(empty(st), ph ^ 1);
        mbar_expect_tx(full(st), my_bytes);
      }
      __syncwarp();
      if (bytes) bulk_g2s(sbase + st * STAGE + dst_off, src + r, bytes, full(st));
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  } else if (warp == 0) {
    // ===== MMA issuer: converged warp, one elected lane issues =====
    const uint32_t idesc = make_idesc_f16_mn(128, dcols);
    const uint32_t hi_y = ((uint32_t)YP >> 4) | (1u << 14), hi_x = ((uint32_t)XP >> 4) | (1u << 14);
    int st = 0, ph = 0;
    for (int it = 0; it < nstage; ++it) {
      mbar_wait(full(st), ph);
      tc_fence_after();
      const uint32_t ylo = desc_lo(sbase + st * STAGE, 128), xlo = desc_lo(sbase + st * STAGE + YBYTES, 128);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ROWS / 16; ++ks) {   // 16 rows = 256 bytes = 16 descriptor units per K step
          const uint64_t yd = desc_of(hi_y, ylo + ks * 16);
          const uint32_t xk = xlo + ks * 16;
          const uint32_t acc = (uint32_t)(it | ks);
          if (ntap == 1) {
            mma_f16<3>(tmem, yd, desc_of(hi_x, xk), idesc, acc);
          } else {
            mma_f16<0>(tmem, yd, desc_of(hi_x, xk), idesc, acc);
#pragma unroll
            for (int tp = 1; tp < 6; ++tp)
              if (tp < ntap - 1) mma_f16<1>(tmem + tp * dcols, yd, desc_of(hi_x, xk + tp), idesc, acc);
            mma_f16<2>(tmem + (ntap - 1) * dcols, yd, desc_of(hi_x, xk + (ntap - 1)), idesc, acc);
          }
        }
        tc_commit(empty(st));
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
    if (nstage > 0) {
      if (elect_one()) tc_commit(acc_full);
      __syncwarp();
    }
  } else if (nstage > 0 && (warp & 3) * 32 < d.cout_g) {
    // ===== warps 4..7: drain the accumulators (TMEM lane quarter = warp % 4), scaled fp32 RED into the gradient =====
    const int q = warp & 3;
    const int lr = q * 32 + lane;   // output channel
    const float out_scale = out_scale_p ? __ldg(out_scale_p) : 1.f;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* const dwg = d.dw + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.sg;
    if (coalesced_drain) {
      // Conv1d weight layout [cout][cin][k]: for one output channel the (cin, tap) plane is one contiguous run.  A TMEM lane
      // is an output channel, so a thread-per-lane RED touches 32 different lines per instruction (the drain of a k7 block
      // was 114 k such sector requests per CTA).  Each drain warp transposes its 32 channels x (32 cin x taps) slab through
      // the pipeline's shared memory (free once acc_full has fired; odd pitch = conflict-free both ways) and issues REDs whose
      // 32 lanes cover 128 contiguous bytes.
      const int run = 32 * ntap, pitch = run + 1;
      float* slab = reinterpret_cast<float*>(smem) + q * (32 * 225);
      for (int cg = 0; cg < dcols / 32; ++cg) {
        for (int tp = 0; tp < ntap; ++tp) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * dcols + cg * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) slab[lane * pitch + i * ntap + tp] = __uint_as_float(v[i]) * out_scale;
        }
        __syncwarp();
        float* dst = dwg + (long)(q * 32) * d.sm + (long)(tile * dcols + cg * 32) * ntap;
        for (int r = 0; r < 32; ++r)
          for (int c = lane; c < run; c += 32) atomicAdd(dst + (long)r * d.sm + c, slab[r * pitch + c]);
        __syncwarp();
      }
    } else
    for (int tp = 0; tp < ntap; ++tp) {
      for (int cg = 0; cg < dcols / 32; ++cg) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(tp * dcols + cg * 32), v);
        tmem_ld_wait();
        float* dst = d.dw + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.sg + (long)lr * d.sm + (long)(tile * dcols + cg * 32) * d.sn + (long)tp * d.st;
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dst + (long)i * d.sn, __uint_as_float(v[i]) * out_scale);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, TM_COLS);
  }
}

// rows [row0, d.rows): one thread per (group, cout, cin, tap)
__global__ void __launch_bounds__(256) wgrad_f16_tail_kernel(const NefWgradDesc d, const __half* __restrict__ dy16,
                                                             const __half* __restrict__ x16, const float* __restrict__ out_scale_p, long row0) {
  const long total = (long)d.groups * d.cout_g * d.cin_g * d.taps;
  const long idx = (long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int t = (int)(idx % d.taps);
  const int n = (int)((idx / d.taps) % d.cin_g);
  const int m = (int)((idx / ((long)d.taps * d.cin_g)) % d.cout_g);
  const int g = (int)(idx / ((long)d.taps * d.cin_g * d.cout_g));
  const long yc = (long)((d.dy_c4_off >> 1) + g * (d.dy_c4_gstride >> 1) + (m >> 3)) * d.dy_cstride;
  const long xc = (long)((d.x_c4_off >> 1) + g * (d.x_c4_gstride >> 1) + (n >> 3)) * d.x_cstride + d.tap_off + t;
  float acc = 0.f;
  for (long r = row0; r < d.rows; ++r)
    acc += __half2float(dy16[(yc + r) * 8 + (m & 7)]) * __half2float(x16[(xc + r) * 8 + (n & 7)]);
  atomicAdd(d.dw + (long)(d.wg_mod > 0 ? g % d.wg_mod : g) * d.sg + (long)m * d.sm + (long)n * d.sn + (long)t * d.st, acc * (out_scale_p ? __ldg(out_scale_p) : 1.f));
}

}  // namespace wf16
}  // namespace nef

using namespace nef;
extern "C" void nef_tc_note_dispatch(int which);

// d: the fp32 descriptor of the same gradient (geometry, dw and its strides; d->dy, d->x and d->db are not read here);
// dy16 / x16: the fp16 copies, `half8 [C/8][cstride rows]`, addressed from the same row origin as the fp32 tensors.
extern "C" int nef_gconv_wgrad_f16(const NefWgradDesc* d, const void* dy16, const void* x16, const float* out_scale, nef_stream_t s) {
  NEF_REQUIRE(d && dy16 && x16, "nef_gconv_wgrad_f16: null argument");
  NEF_REQUIRE(d->cout_g == 128 || d->cout_g == 64, "nef_gconv_wgrad_f16: cout_g must be 64 or 128 (got %d)", d->cout_g);
  NEF_REQUIRE(d->cin_g % 64 == 0 && d->taps >= 1 && d->taps <= 7, "nef_gconv_wgrad_f16: cin_g %% 64 == 0 and 1..7 taps required");
  NEF_REQUIRE(((d->dy_c4_off | d->dy_c4_gstride | d->x_c4_off | d->x_c4_gstride) & 1) == 0,
              "nef_gconv_wgrad_f16: chunk offsets and group strides must be even (8-channel fp16 chunks)");
  const int dcols = (d->taps <= 3 && d->cin_g % 128 == 0) ? 128 : 64;   // taps * dcols <= 448 TMEM columns
  cudaError_t e = cudaFuncSetAttribute(wf16::wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wf16::TOTAL);
  NEF_REQUIRE(e == cudaSuccess, "nef_gconv_wgrad_f16: shared-memory opt-in failed: %s", cudaGetErrorString(e));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long nst_total = d->rows / wf16::ROWS;
  const long rows_main = nst_total * wf16::ROWS;
  if (nst_total > 0) {
    const long tiles = (long)(d->cin_g / dcols) * d->groups;
    // row splits: the smallest count whose last wave is >= 90 % full (else the best seen) -- every extra split costs
    // another cout x cin x taps block of fp32 REDs
    long best = 1;
    double best_eff = 0.0;
    const long max_splits = nst_total < 2L * sms ? nst_total : 2L * sms;
    for (long sp = 1; sp <= max_splits; ++sp) {
      const long ctas = tiles * sp;
      const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
      if (eff >= 0.9) { best = sp; break; }
    }
    const long st_per_split = (nst_total + best - 1) / best;
    const long splits = (nst_total + st_per_split - 1) / st_per_split;
    dim3 grid((unsigned)(d->cin_g / dcols), (unsigned)splits, (unsigned)d->groups);
    // NEF_WGRAD_DRAIN=0: the thread-per-channel RED drain for every layout (A/B switch)
    static const int drain_mode = getenv("NEF_WGRAD_DRAIN") ? atoi(getenv("NEF_WGRAD_DRAIN")) : 1;
    const int coalesced = drain_mode && d->st == 1 && d->sn == d->taps;
    wf16::wgrad_f16_kernel<<<grid, wf16::THREADS, wf16::TOTAL, (cudaStream_t)s>>>(
        *d, reinterpret_cast<const uint4*>(dy16), reinterpret_cast<const uint4*>(x16), out_scale, dcols,
        st_per_split * wf16::ROWS, rows_main, coalesced);
    NEF_CHECK_LAUNCH("wgrad_f16_kernel");
    nef_tc_note_dispatch(5);
  }
  if (rows_main < d->rows) {
    const long total = (long)d->groups * d->cout_g * d->cin_g * d->taps;
    wf16::wgrad_f16_tail_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(
        *d, reinterpret_cast<const __half*>(dy16), reinterpret_cast<const __half*>(x16), out_scale, rows_main);
    NEF_CHECK_LAUNCH("wgrad_f16_tail_kernel");
  }
  return 0;
}
