// Shared definitions for the Nef-Net B200 hot path (sm_100a only).
//
// Internal activation layout "CBL4" (channel-chunk major, zero halo):
//   a tensor with C channels (C % 4 == 0), B segments and L samples is an array of float4 rows
//       T[c4][b * Lp + P + l]   (c4 = channel / 4, lane = channel % 4),  Lp = L + 2 * P, P = NEF_HALO
//   the P halo rows on each side of every segment are kept zero for the life of the tensor: kernels
//   never store to them, so a k-tap "same" convolution over the flattened row axis needs no boundary
//   logic and a tile of rows is one contiguous run of 16-byte rows per channel chunk (bulk-copyable,
//   and directly a tcgen05 no-swizzle operand: K-major when K = channels, MN-major when K = rows).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NEF_HALO 3
#define NEF_GUARD_ROWS 528   // readable rows after (and before) every CBL4 tensor: tiles may overrun
#define NEF_NROI 7
#define NEF_ROI_SIZE 16

extern "C" void nef_set_error(const char* fmt, ...);

extern "C" void nef_count_launch(void);  // every kernel launch of the library is counted (nef_launch_count)

#define NEF_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    nef_count_launch();                                                               \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      nef_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));          \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define NEF_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      nef_set_error(__VA_ARGS__);       \
      return 2;                         \
    }                                   \
  } while (0)

// One launch packs up to NEF_PACK_MAX weight tensors (nef_pack_weights semantics per job); the table is a kernel parameter.
#define NEF_PACK_MAX 40
#define NEF_PACK_CHUNK 4096   // elements of a job per block
struct NefPackJob {
  const float* src;
  float* dst;
  int groups, N, K, taps;
  long sg, sn, sk, st;
  int flags;        // bit 0: flip taps, bit 1: TF32 residual, bit 2: fp16 operand packing (8 channels per 16-byte slot)
  int first_block;  // filled by nef_pack_weights_batch
  int gmod;         // 0: every group has its own source weights; m > 0: group g reads source group g mod m (shared weights)
  const float* nscale;  // optional [groups * N]: the weights of output channel (g, n) are multiplied by nscale[g * N + n]
                        //   before rounding (inference-time BatchNorm folding)
};
struct NefPackTable {
  NefPackJob job[NEF_PACK_MAX];
  int n;
};
static_assert(sizeof(NefPackTable) <= 4000, "the packing table travels as a kernel parameter");
int nef_pack_weights_batch(NefPackTable* tab, cudaStream_t s);  // launches the jobs queued in tab and empties it

namespace nef {

// Test hook (nef_set_exact_fp32): when set, nothing is rounded to TF32, so that the CUDA-core
// implementation is a plain fp32 computation that can be compared tightly with the fp32 oracle.
// One copy per translation unit; NEF_DEFINE_EXACT_SETTER(name) defines the unit's setter.
static __constant__ int c_nef_exact;
#define NEF_DEFINE_EXACT_SETTER(name) \
  extern "C" int name(int on) { return cudaMemcpyToSymbol(nef::c_nef_exact, &on, sizeof(int)) == cudaSuccess ? 0 : 1; }

__device__ __forceinline__ float tf32_rn(float x) {
  if (c_nef_exact) return x;
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_rn4(float4 v) {
  return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}

// Counter-based dropout bits.  The epilogues that apply dropout are instruction-issue-bound, so the generator is built from
// 32-bit multiplies only (a 64-bit splitmix per float4 cost ~25 instructions per element, this costs ~6): hash32 is the
// two-multiply "lowbias32" finaliser (full avalanche); the row index goes through it once, keyed by the seed (a bijection
// per seed, hoisted out of the chunk loop by the compiler), and every float4 takes two more hashes of row key + chunk * odd.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  return x ^ (x >> 16);
}
// four 16-bit uniforms for the float4 at (row, chunk) of a tensor; keep lane i iff u_i >= p * 65536
// (host mirror: tests/_dropmask.py)
__device__ __forceinline__ uint64_t drop_bits(uint64_t seed, long row, int c4) {
  const uint32_t sk = hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + 0x9E3779B9u));
  const uint32_t rk = hash32((uint32_t)row ^ sk);
  const uint32_t k = rk + (uint32_t)c4 * 0x9E3779B1u;
  const uint32_t lo = hash32(k), hi = hash32(k ^ 0x85EBCA6Bu);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float& f4at(float4& v, int i) { return reinterpret_cast<float*>(&v)[i]; }
__device__ __forceinline__ float f4get(const float4& v, int i) { return reinterpret_cast<const float*>(&v)[i]; }

}  // namespace nef
