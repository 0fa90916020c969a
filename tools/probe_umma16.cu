// Probe: MN-major A with kind::f16 (bf16) vs kind::tf32 -- is the zero result a descriptor mistake or a tf32 limitation?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t lt = 0) {
  return ((uint64_t)lt << 61) | (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void probe(int a_major, uint32_t a_lbo, uint32_t a_sbo, int lt, float* out, uint32_t a_shift = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // 64 KB = 32768 halves
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + 65536);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // value encodes (16-byte unit index) : unit u -> value (u % 251) + element/16 fraction is not representable; use unit id only
  for (int i = tid; i < 32768; i += 128) A[i] = __float2bfloat16((float)((i / 8) % 256));
  for (int i = tid; i < 32768; i += 128) Bm[i] = __float2bfloat16(0.f);
  __syncthreads();
  // K-major B (N=128, K=16): chunk c = k/8 at +c*128*8 halves ; row n at +n*8 halves
  if (tid < 16) { int k = tid, n = tid; Bm[(k / 8) * 1024 + n * 8 + (k % 8)] = __float2bfloat16(1.f); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (tid == 0) {
    // c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10)
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_major << 15) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t ad = make_desc(smem_u32(A) + a_shift, a_lbo, a_sbo, lt), bd = make_desc(smem_u32(Bm), 128 * 16, 128);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;");
  {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(tmem + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
int main() {
  float* out; cudaMalloc(&out, 128 * 32 * 4);
  static float h[128 * 32];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 256);
  struct V { const char* name; int a_major; uint32_t lbo, sbo; int lt; uint32_t shift; } vs[] = {
    {"bf16 A K-major  LBO=2048 SBO=128 (sanity): m -> unit m (k<8), unit 128+m (k>=8)", 0, 2048, 128, 0},
    {"bf16 A MN-major NONE LBO=128 SBO=1024", 1, 128, 1024, 0},
    {"bf16 A MN-major NONE LBO=1024 SBO=128", 1, 1024, 128, 0},
    {"bf16 A MN-major SW128 LBO=2048 SBO=1024", 1, 2048, 1024, 2},
    {"bf16 A MN-major SW128 LBO=1024 SBO=2048", 1, 1024, 2048, 2},
    // the weight-gradient kernel without a re-tile pass (csrc/nef_wgrad_f16.cu) shifts the start address by one 16-byte
    // row per tap: expected = every unit index of the LBO=128 / SBO=1024 variant above plus 1 (plus 3)
    {"bf16 A MN-major NONE LBO=128 SBO=1024, start + 16 B", 1, 128, 1024, 0, 16},
    {"bf16 A MN-major NONE LBO=128 SBO=1024, start + 48 B", 1, 128, 1024, 0, 48},
  };
  for (auto& v : vs) {
    cudaMemset(out, 0xff, 128 * 32 * 4);
    probe<<<1, 128, 131072 + 256>>>(v.a_major, v.lbo, v.sbo, v.lt, out, v.shift);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== %s : %s\n", v.name, cudaGetErrorString(e));
    const int ms[] = {0, 1, 7, 8, 9, 15, 16, 17, 32, 64, 127};
    for (int m : ms) {
      printf("  m=%3d: unit read for k=0..15 :", m);
      for (int k = 0; k < 16; ++k) printf(" %4.0f", h[m * 32 + k]);
      printf("\n");
    }
  }
  return 0;
}
