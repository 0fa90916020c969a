// Callers either side of the hot path, kept on the device so the step loop has no host
// round trips:
//   * nef_psnr            -- utils/mertic.py:7-21 (PSNR over the valid part of every synthesized view), accumulated in
//                            device memory across validation batches, one read-back per epoch;
//   * nef_prepare_segments -- dataset/tianchi.py:84-111, 212-225: lead derivation, heartbeat crop, min-max
//                            normalisation, zero padding / truncation to L, ROI table, lead selection.
// Both are HBM-trivial byte/float shuffles; arithmetic is done in double exactly as numpy does it in the reference.
#include "nef_common.cuh"
#include "../../include/nefnet_b200.h"

namespace nef {

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];  // fixed order
  return t;
}

// one block per (segment, view) row
__global__ void __launch_bounds__(256) psnr_rows_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                        const int64_t* __restrict__ rois, int V, int L,
                                                        double* __restrict__ rows) {
  __shared__ double red[8];
  const long r = blockIdx.x;
  const int b = (int)(r / V);
  long end = L;
  if (rois) {  // rois[i, -1, 0] (mertic.py:11); python slice semantics for out-of-range / negative values
    end = rois[((long)b * NEF_NROI + (NEF_NROI - 1)) * 2];
    if (end < 0) end += L;
    if (end < 0) end = 0;
    if (end > L) end = L;
  }
  const float* p = pred + r * L;
  const float* g = gt + r * L;
  double acc = 0.0;
  for (long i = threadIdx.x; i < end; i += 256) {
    const float d = p[i] - g[i];  // float32 difference, as numpy computes it (:14)
    acc += (double)d * (double)d;
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const double rmse = sqrt(tot / (double)end);  // end == 0: 0/0 = nan, as np.mean of an empty slice
    rows[r] = rmse == 0.0 ? 100.0 : 20.0 * log10(1.0 / rmse);
  }
}

// acc[0] += sum of the row values (fixed order), acc[1] += n; result[0] = running mean
__global__ void __launch_bounds__(256) psnr_accumulate_kernel(const double* __restrict__ rows, long n, double* acc,
                                                              float* result) {
  __shared__ double red[8];
  double s = 0.0;
  for (long i = threadIdx.x; i < n; i += 256) s += rows[i];
  const double tot = block_sum(s, red);
  if (threadIdx.x == 0) {
    acc[0] += tot;
    acc[1] += (double)n;
    if (result) result[0] = (float)(acc[0] / acc[1]);
  }
}

// lead k (0..11) of a record at sample t: the 8 recorded leads, then III, aVR, aVL, aVF (tianchi.py:88-93)
__device__ __forceinline__ double lead_value(const double* rec, long T, int k, long t) {
  if (k < 8) return rec[(long)k * T + t];
  const double I = rec[t], II = rec[T + t];
  switch (k) {
    case 8: return II - I;
    case 9: return -0.5 * (I + II);
    case 10: return I - 0.5 * II;
    default: return II - 0.5 * I;
  }
}

// numpy slice [p_on:end_point] on an axis of length T -> [lo, lo + n)
__device__ __forceinline__ void crop_range(const int64_t* mk, long T, long& lo, long& n) {
  const long p_on = mk[0], end = mk[6];
  lo = p_on < 0 ? p_on + T : p_on;
  long hi = end < 0 ? end + T : end;
  lo = lo < 0 ? 0 : (lo > T ? T : lo);
  hi = hi < 0 ? 0 : (hi > T ? T : hi);
  n = hi > lo ? hi - lo : 0;
}

constexpr int PREP_SPLITS = 16;  // blocks per segment (both passes)

// pass 1: partial min / max of the 12 x crop block (tianchi.py:110), one (segment, split) per block
__global__ void __launch_bounds__(256) prepare_minmax_kernel(const double* __restrict__ raw, const int64_t* __restrict__ rec_off,
                                                             const int32_t* __restrict__ rec_len,
                                                             const int64_t* __restrict__ marks, double* __restrict__ part) {
  __shared__ double smin[8], smax[8];
  const int b = blockIdx.y, sp = blockIdx.x, tid = threadIdx.x;
  const long T = rec_len[b];
  const double* rec = raw + rec_off[b];
  long lo, n;
  crop_range(marks + (long)b * 7, T, lo, n);
  double mn = INFINITY, mx = -INFINITY;
  for (long l = (long)sp * 256 + tid; l < n; l += (long)PREP_SPLITS * 256) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const double v = lead_value(rec, T, k, lo + l);
      mn = fmin(mn, v);
      mx = fmax(mx, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((tid & 31) == 0) { smin[tid >> 5] = mn; smax[tid >> 5] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 8; ++i) { mn = fmin(mn, smin[i]); mx = fmax(mx, smax[i]); }
    part[((long)b * PREP_SPLITS + sp) * 2 + 0] = mn;
    part[((long)b * PREP_SPLITS + sp) * 2 + 1] = mx;
  }
}

// pass 2: normalise, pad / truncate, select; one (segment, split) per block
__global__ void __launch_bounds__(256) prepare_segments_kernel(const double* __restrict__ raw, const int64_t* __restrict__ rec_off,
                                                               const int32_t* __restrict__ rec_len,
                                                               const int64_t* __restrict__ marks, int L,
                                                               const int32_t* __restrict__ select, int G,
                                                               const int32_t* __restrict__ target_index,
                                                               const double* __restrict__ part,
                                                               float* __restrict__ ori, float* __restrict__ data,
                                                               float* __restrict__ target, int64_t* __restrict__ rois) {
  const int b = blockIdx.y, sp = blockIdx.x, tid = threadIdx.x;
  const int64_t* mk = marks + (long)b * 7;
  const long T = rec_len[b];
  const double* rec = raw + rec_off[b];
  long lo, n;
  crop_range(mk, T, lo, n);
  if (rois && sp == 0 && tid < 7) {  // tianchi.py:103-106
    const int64_t e = tid < 6 ? mk[tid + 1] : (int64_t)L + mk[0];
    rois[((long)b * 7 + tid) * 2 + 0] = mk[tid] - mk[0];
    rois[((long)b * 7 + tid) * 2 + 1] = e - mk[0];
  }
  double mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < PREP_SPLITS; ++i) {
    mn = fmin(mn, part[((long)b * PREP_SPLITS + i) * 2 + 0]);
    mx = fmax(mx, part[((long)b * PREP_SPLITS + i) * 2 + 1]);
  }
  const double vmin = mn, range = mx - mn;  // (x - min) / (max - min)  (:110-111)
  const int tgt = target_index ? target_index[b] : -1;
  for (long l = (long)sp * 256 + tid; l < L; l += (long)PREP_SPLITS * 256) {
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
      v[k] = l < n ? (float)((lead_value(rec, T, k, lo + l) - vmin) / range) : 0.f;  // pad / truncate (:212-219)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      if (ori) ori[((long)b * 12 + k) * L + l] = v[k];
      if (k == tgt) target[(long)b * L + l] = v[k];
    }
    if (data) {
      for (int gi = 0; gi < G; ++gi) {
        const int k = select[(long)b * G + gi];
        float sel = 0.f;
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) sel = kk == k ? v[kk] : sel;
        data[((long)b * G + gi) * L + l] = sel;
      }
    }
  }
}

}  // namespace nef

using namespace nef;

extern "C" int nef_psnr(const float* pred, const float* gt, const int64_t* rois, int B, int V, int L, double* rows,
                        double* acc, float* result, nef_stream_t s) {
  NEF_REQUIRE(pred && gt && rows && acc && B >= 1 && V >= 1 && L >= 1, "nef_psnr: bad arguments");
  psnr_rows_kernel<<<(unsigned)((long)B * V), 256, 0, (cudaStream_t)s>>>(pred, gt, rois, V, L, rows);
  NEF_CHECK_LAUNCH("psnr_rows_kernel");
  psnr_accumulate_kernel<<<1, 256, 0, (cudaStream_t)s>>>(rows, (long)B * V, acc, result);
  NEF_CHECK_LAUNCH("psnr_accumulate_kernel");
  return 0;
}

extern "C" size_t nef_prepare_scratch_bytes(int B) { return (size_t)B * nef::PREP_SPLITS * 2 * sizeof(double); }

extern "C" int nef_prepare_segments(const double* raw, const int64_t* rec_off, const int32_t* rec_len, const int64_t* marks,
                                    int B, int L, const int32_t* select, int G, const int32_t* target_index, double* scratch,
                                    float* ori, float* data, float* target, int64_t* rois, nef_stream_t s) {
  NEF_REQUIRE(raw && rec_off && rec_len && marks && scratch && B >= 1 && L >= 1, "nef_prepare_segments: bad arguments");
  NEF_REQUIRE(!data || (select && G >= 1), "nef_prepare_segments: data needs select (B, G)");
  NEF_REQUIRE(!target_index || target, "nef_prepare_segments: target_index needs target (B, L)");
  NEF_REQUIRE(B <= 65535, "nef_prepare_segments: at most 65535 segments per call");
  dim3 grid(PREP_SPLITS, (unsigned)B);
  prepare_minmax_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(raw, rec_off, rec_len, marks, scratch);
  NEF_CHECK_LAUNCH("prepare_minmax_kernel");
  prepare_segments_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(raw, rec_off, rec_len, marks, L, select, G, target_index, scratch,
                                                            ori, data, target, rois);
  NEF_CHECK_LAUNCH("prepare_segments_kernel");
  return 0;
}
