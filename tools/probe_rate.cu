// Probe: sustained tcgen05.mma issue/execute rate (cycles per instruction) for kind::tf32 / kind::f16, M = 128,
// N in {64, 128, 256}, SWIZZLE_NONE vs SWIZZLE_128B K-major operands, regular vs weight-stationary form,
// one CTA per SM on all SMs.  Operands are zeros (timing only).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t lt) {
  return ((uint64_t)lt << 61) | (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void rate(int kind_f16, int N, int lt, int ws, int nacc, int iters, long long* out, int a_shift = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint32_t fmt = kind_f16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    // A at smem + 0 (64 KB window), B at smem + 64 KB (128 KB window)
    const uint32_t a0 = smem_u32(smem) + a_shift, b0 = smem_u32(smem) + 65536;   // a_shift: the tap shift of the conv kernels (16 B = one row)
    const uint64_t ad = lt ? (make_desc(a0, 16, 1024, 2) | ((uint64_t)((a0 >> 7) & 7) << 49)) : make_desc(a0, 8320, 128, 0);
    const uint64_t bd = lt ? make_desc(b0, 16, 1024, 2) : make_desc(b0, (uint32_t)N * 16, 128, 0);
    long long t0 = clock64();
#define MMA_F16(D) asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(D), "l"(ad), "l"(bd), "r"(idesc) : "memory")
#define MMA_TF32(D) asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(D), "l"(ad), "l"(bd), "r"(idesc) : "memory")
#define MMA_WS(D, MODE) asm volatile("tcgen05.mma.ws.cta_group::1.kind::tf32.collector::b0::" MODE " [%0], %1, %2, %3, 1;" ::"r"(D), "l"(ad), "l"(bd), "r"(idesc) : "memory")
    const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? N : 0), d2 = tmem + (nacc > 2 ? 2 * N : 0), d3 = tmem + (nacc > 3 ? 3 * N : 0);
#define MMA_F16_WS(D, MODE) asm volatile("tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::" MODE " [%0], %1, %2, %3, 1;" ::"r"(D), "l"(ad), "l"(bd), "r"(idesc) : "memory")
#define MMA_F16_A(D, MODE) asm volatile("tcgen05.mma.cta_group::1.kind::f16.collector::a::" MODE " [%0], %1, %2, %3, 1;" ::"r"(D), "l"(ad), "l"(bd), "r"(idesc) : "memory")
    if (kind_f16 && ws == 1) {        // weight-stationary pairs, as the persistent conv kernel issues them (2 row tiles)
      for (int i = 0; i < iters; i += 4) { MMA_F16_WS(d0, "fill"); MMA_F16_WS(d1, "lastuse"); MMA_F16_WS(d2, "fill"); MMA_F16_WS(d3, "lastuse"); }
    } else if (kind_f16 && ws == 2) { // A held in the collector over 4 MMAs (the weight-gradient kernels' tap loop)
      for (int i = 0; i < iters; i += 4) { MMA_F16_A(d0, "fill"); MMA_F16_A(d1, "use"); MMA_F16_A(d2, "use"); MMA_F16_A(d3, "lastuse"); }
    } else if (kind_f16) {
      for (int i = 0; i < iters; i += 4) { MMA_F16(d0); MMA_F16(d1); MMA_F16(d2); MMA_F16(d3); }
    } else if (!ws) {
      for (int i = 0; i < iters; i += 4) { MMA_TF32(d0); MMA_TF32(d1); MMA_TF32(d2); MMA_TF32(d3); }
    } else {
      for (int i = 0; i < iters; i += 4) { MMA_WS(d0, "fill"); MMA_WS(d1, "use"); MMA_WS(d2, "use"); MMA_WS(d3, "lastuse"); }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); } while (!done);
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 148 * 8);
  static long long h[148];
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct V { const char* name; int f16, N, lt, ws, nacc, shift; } vs[] = {
    {"tf32 N=128 NONE  4 acc", 0, 128, 0, 0, 4}, {"tf32 N=128 SW128 4 acc", 0, 128, 1, 0, 4}, {"tf32 N=128 NONE  1 acc", 0, 128, 0, 0, 1},
    {"tf32 N=64  NONE  4 acc", 0, 64, 0, 0, 4},  {"tf32 N=64  SW128 4 acc", 0, 64, 1, 0, 4},  {"tf32 N=256 NONE  2 acc", 0, 256, 0, 0, 2},
    {"tf32 N=256 SW128 2 acc", 0, 256, 1, 0, 2}, {"tf32 N=128 NONE  4 acc WS", 0, 128, 0, 1, 4}, {"tf32 N=64 NONE 4 acc WS", 0, 64, 0, 1, 4},
    {"bf16 N=128 NONE  4 acc", 1, 128, 0, 0, 4}, {"bf16 N=128 SW128 4 acc", 1, 128, 1, 0, 4}, {"bf16 N=256 SW128 2 acc", 1, 256, 1, 0, 2},
    {"f16 N=256 NONE 2 acc", 1, 256, 0, 0, 2}, {"f16 N=128 NONE 4 acc WS pairs", 1, 128, 0, 1, 4}, {"f16 N=128 NONE 4 acc A-reuse", 1, 128, 0, 2, 4},
    {"f16 N=64 NONE 4 acc", 1, 64, 0, 0, 4}, {"f16 N=64 NONE 4 acc A-reuse", 1, 64, 0, 2, 4},
    {"f16 N=128 NONE A + 16 B", 1, 128, 0, 0, 4, 16}, {"f16 N=128 NONE A + 48 B", 1, 128, 0, 0, 4, 48}, {"f16 N=128 NONE A + 64 B", 1, 128, 0, 0, 4, 64},
    {"tf32 N=128 NONE A + 16 B", 0, 128, 0, 0, 4, 16}, {"f16 N=128 SW128 A + 128 B", 1, 128, 1, 0, 4, 128}, {"f16 N=128 SW128 A + 384 B", 1, 128, 1, 0, 4, 384},
  };
  const int iters = 4000;
  for (int grid : {148}) for (auto& v : vs) {
    rate<<<grid, 128, 200 * 1024>>>(v.f16, v.N, v.lt, v.ws, v.nacc, iters, out, v.shift);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d  %-28s %s  %7.1f cycles/MMA  (ideal %d)\n", grid, v.name, cudaGetErrorString(e), (double)mx / iters, 128 * v.N / 256);
  }
  return 0;
}
