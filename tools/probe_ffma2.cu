// FP32 FMA issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/probe_ffma2 tools/probe_ffma2.cu && tools/_bin/probe_ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  unsigned long long p[8], ss;
  float2 s2 = make_float2(s, s);
  ss = *reinterpret_cast<unsigned long long*>(&s2);
#pragma unroll
  for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(0.5f));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
    }
  }
  float r = 0.f;
  if (MODE == 0) { for (int i = 0; i < 16; ++i) r += a[i]; }
  else { for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&p[i]); r += t.x + t.y; } }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)148 * 8 * 256 * 16.0 * iters;
      if (rep) printf("%s: %.3f ms  %.1f TFLOP/s fp32\n", mode ? "FFMA2 (f32x2)" : "FFMA scalar ", ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
