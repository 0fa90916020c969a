// Probe: how does tcgen05.mma (kind::tf32, SWIZZLE_NONE) address an MN-major A operand?
// B is a K-major "selector" (B[n][k] = (n == k)), so D[m][n<8] = A[m][k = n]; A's shared memory is filled with
// the float value of each word's index, so D reveals which word the hardware read for (m, k).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ int g_layout;
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint64_t lt = 0) {
  return ((uint64_t)lt << 61) | (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void probe(int a_major, int b_major, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_off, int mode, float* out, int lt) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* A = reinterpret_cast<float*>(smem);                 // 64 KB = 16384 words
  float* Bm = reinterpret_cast<float*>(smem + 65536);        // 64 KB
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (mode == 0) {   // probe A: A = word index, B = K-major selector
    for (int i = tid; i < 16384; i += 128) A[i] = (float)i;
    for (int i = tid; i < 16384; i += 128) Bm[i] = 0.f;
    __syncthreads();
    // K-major B: chunk c (k/4) at Bm + c*128*4 floats; row n at +n*4 floats
    if (tid < 8) { int k = tid, n = tid; Bm[(k / 4) * 512 + n * 4 + (k % 4)] = 1.f; }
  } else {           // probe B: B = word index (descriptor under test applies to B), A = K-major selector
    for (int i = tid; i < 16384; i += 128) Bm[i] = (float)i;
    for (int i = tid; i < 16384; i += 128) A[i] = 0.f;
    __syncthreads();
    if (tid < 8) { int k = tid, m = tid; A[(k / 4) * 512 + m * 4 + (k % 4)] = 1.f; }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = slot;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t ad, bd;
    if (mode == 0) { ad = make_desc(smem_u32(A) + a_off, a_lbo, a_sbo, lt); bd = make_desc(smem_u32(Bm), 128 * 16, 128); }
    else { ad = make_desc(smem_u32(A), 128 * 16, 128); bd = make_desc(smem_u32(Bm) + a_off, a_lbo, a_sbo, lt); }
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int cg = 0; cg < 4; ++cg) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cg * 32) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + cg * 32 + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
int main() {
  float* out; cudaMalloc(&out, 128 * 128 * 4);
  static float h[128 * 128];
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 256);
  struct V { const char* name; int mode, a_major, b_major; uint32_t lbo, sbo, off; int lt; } vs[] = {
    {"A K-major  LBO=2048 SBO=128 (sanity)", 0, 0, 0, 2048, 128, 0, 0},
    {"A MN-major SW32  LBO=4096 SBO=256", 0, 1, 0, 4096, 256, 0, 6},
    {"A MN-major SW32  LBO=256 SBO=4096", 0, 1, 0, 256, 4096, 0, 6},
    {"A MN-major SW64  LBO=8192 SBO=512", 0, 1, 0, 8192, 512, 0, 4},
    {"A MN-major SW128 LBO=1024 SBO=4096", 0, 1, 0, 1024, 4096, 0, 2},
    {"A MN-major SW128 LBO=4096 SBO=1024", 0, 1, 0, 4096, 1024, 0, 2},
    {"A MN-major SW128 LBO=4096 SBO=1024 off=128 (one k row)", 0, 1, 0, 4096, 1024, 128, 2},
    {"A MN-major SW128 LBO=4096 SBO=1024 off=384 (3 k rows)", 0, 1, 0, 4096, 1024, 384, 2},
    {"A K-major  SW128 SBO=1024", 0, 0, 0, 16, 1024, 0, 2},
    {"A K-major  SW128 SBO=1024 off=128 (one m row)", 0, 0, 0, 16, 1024, 128, 2},
    {"A K-major  SW128 SBO=1024 off=384 (3 m rows)", 0, 0, 0, 16, 1024, 384, 2},
    {"A K-major  SW32 SBO=256", 0, 0, 0, 16, 256, 0, 6},
    {"A K-major  SW32 SBO=256 off=32 (one m row)", 0, 0, 0, 16, 256, 32, 6},
    {"A K-major  SW32 SBO=256 off=160 (5 m rows)", 0, 0, 0, 16, 256, 160, 6},
  };
  for (auto& v : vs) {
    cudaMemset(out, 0xff, 128 * 128 * 4);
    probe<<<1, 128, 131072 + 256>>>(v.a_major, v.b_major, v.lbo, v.sbo, v.off, v.mode, out, v.lt);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== %s : %s\n", v.name, cudaGetErrorString(e));
    const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 64, 127};
    for (int m : ms) {
      printf("  %s=%3d: word read for k=0..7 :", v.mode == 0 ? "m" : "n", m);
      for (int k = 0; k < 8; ++k) printf(" %6.0f", v.mode == 0 ? h[m * 128 + k] : h[k * 128 + m]);
      printf("\n");
    }
  }
  return 0;
}
