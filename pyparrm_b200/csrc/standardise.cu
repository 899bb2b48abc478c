// Standardisation prologue of the period search: PARRM._standardise_data (parrm.py:272-280).
//   d[c,t] = x[c,t+1] - x[c,t];  s_c = mean_t |d[c,t]|;  z[c,t] = clip(d[c,t] / s_c, -ob, +ob)
// The reference materialises z for the whole recording; the search only ever reads z at the
// <= 25 001 fitted indices (parrm.py:589-591), so the device path is one streaming pass for
// s_c (HBM-bound, 8 B read per channel-sample) plus a gather of the fitted columns.
#include "common.cuh"

namespace parrm {

constexpr int kScaleThreads = 256;
constexpr int kScaleChunk = 16384;  // diffs per CTA

// 16-byte loads: a thread takes VEC consecutive samples per pass (2 doubles / 4 floats) and the
// first sample of its right-hand neighbour by shuffle; four passes are in flight per thread.
// Rows whose base is not 16-byte aligned (odd row stride) take the scalar loop.
template <typename T>
struct Vec16;
template <>
struct Vec16<double> {
  typedef double2 type;
  static constexpr int N = 2;
};
template <>
struct Vec16<float> {
  typedef float4 type;
  static constexpr int N = 4;
};

template <typename T>
__global__ void __launch_bounds__(kScaleThreads)
abs_diff_partial_kernel(const T* __restrict__ x, int64_t n_diffs, int64_t ld, int n_chunks,
                        double* __restrict__ partial) {
  __shared__ double warp_part[kScaleThreads / 32];
  typedef typename Vec16<T>::type V;
  constexpr int N = Vec16<T>::N;
  const int64_t chan = blockIdx.y;
  const T* row = x + chan * ld;
  const int64_t begin = int64_t(blockIdx.x) * kScaleChunk;
  const int64_t end = min(begin + kScaleChunk, n_diffs);
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  int64_t t = begin;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    // whole warps of vector passes while every lane's N samples and its neighbour's first
    // sample exist: lane l covers samples [t0 + l N, t0 + (l + 1) N], diffs t0 + l N .. + N - 1
    constexpr int kPass = kScaleThreads * N;
    for (; t + kPass <= end; t += kPass) {
      const int64_t t0 = t + int64_t(threadIdx.x) * N;
      const V v = *reinterpret_cast<const V*>(row + t0);
      T e[N + 1];
      if (N == 2) {
        e[0] = reinterpret_cast<const T*>(&v)[0];
        e[1] = reinterpret_cast<const T*>(&v)[1];
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) e[i] = reinterpret_cast<const T*>(&v)[i];
      }
      T next = __shfl_down_sync(0xffffffffu, e[0], 1);
      if (lane == 31) next = row[t0 + N];  // t0 + N <= end <= n_diffs: a valid sample
      e[N] = next;
#pragma unroll
      for (int i = 0; i < N; ++i) acc += fabs(double(e[i + 1] - e[i]));  // data's own precision
    }
  }
  for (int64_t u = t + threadIdx.x; u < end; u += kScaleThreads)
    acc += fabs(double(row[u + 1] - row[u]));
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < kScaleThreads / 32; ++i) s += warp_part[i];
    partial[chan * n_chunks + blockIdx.x] = s;
  }
}

// One warp per channel; partials are added in a fixed order (deterministic result).
__global__ void scale_finalise_kernel(const double* __restrict__ partial, int n_chunks,
                                      int64_t n_diffs, int64_t n_chans,
                                      double* __restrict__ scale) {
  const int64_t chan = int64_t(blockIdx.x) * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (chan >= n_chans) return;
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int i = lane; i < n_chunks; i += 32) acc += partial[chan * n_chunks + i];
  acc = warp_sum(acc);
  if (lane == 0) scale[chan] = acc / double(n_diffs);
}

__device__ __forceinline__ double clip_keep_nan(double v, double bound) {
  return v < -bound ? -bound : (v > bound ? bound : v);  // NaN passes through, like np.clip
}

template <typename T>
__global__ void __launch_bounds__(256)
standardise_gather_kernel(const T* __restrict__ x, int64_t ld, const int64_t* __restrict__ indices,
                          int64_t n_indices, const double* __restrict__ scale, double bound,
                          double* __restrict__ y, int64_t ld_y) {
  const int64_t chan = blockIdx.y;
  const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_indices) return;
  const T* row = x + chan * ld;
  const int64_t t = indices[j];
  const T d = row[t + 1] - row[t];
  y[j * ld_y + chan] = clip_keep_nan(double(T(d / T(scale[chan]))), bound);  // sample-major
}

template <typename T>
__global__ void __launch_bounds__(256)
channel_sumsq_kernel(const T* __restrict__ y, int64_t ld_y, int64_t n,
                     double* __restrict__ sumsq) {
  __shared__ double warp_part[8];
  const T* col = y + blockIdx.x;  // channel blockIdx.x of the sample-major tile
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += 256) {
    const double v = double(col[j * ld_y]);
    acc = fma(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += warp_part[i];
    sumsq[blockIdx.x] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
standardise_full_kernel(const T* __restrict__ x, int64_t n_diffs, int64_t ld,
                        const double* __restrict__ scale, double bound, T* __restrict__ z,
                        int64_t ld_z) {
  const int64_t chan = blockIdx.y;
  const T* row = x + chan * ld;
  const T s = T(scale[chan]);
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n_diffs;
       t += int64_t(gridDim.x) * blockDim.x) {
    const T d = row[t + 1] - row[t];
    z[chan * ld_z + t] = T(clip_keep_nan(double(T(d / s)), bound));
  }
}

inline int scale_chunks(int64_t n_samples) {
  const int64_t n_diffs = n_samples > 1 ? n_samples - 1 : 0;
  return int(max64(1, ceil_div(n_diffs, kScaleChunk)));
}

}  // namespace parrm

extern "C" {

size_t parrm_channel_scales_workspace_bytes(int64_t n_chans, int64_t n_samples) {
  if (n_chans <= 0) return 0;
  return size_t(n_chans) * size_t(parrm::scale_chunks(n_samples)) * sizeof(double);
}

int parrm_channel_scales(const void* d_x, int64_t n_chans, int64_t n_samples, int64_t ld,
                         double* d_scale, void* d_workspace, size_t workspace_bytes, int dtype,
                         void* stream) {
  PARRM_NVTX("parrm_channel_scales");
  using namespace parrm;
  PARRM_REQUIRE(n_chans > 0 && n_chans <= 65535 && n_samples >= 1 && ld >= n_samples,
                "parrm_channel_scales: bad shape (%lld x %lld, ld %lld)", (long long)n_chans,
                (long long)n_samples, (long long)ld);
  PARRM_REQUIRE(d_x && d_scale && d_workspace, "parrm_channel_scales: null pointer");
  if (workspace_bytes < parrm_channel_scales_workspace_bytes(n_chans, n_samples)) {
    set_error("parrm_channel_scales: workspace too small");
    return PARRM_ERR_WORKSPACE;
  }
  const int64_t n_diffs = n_samples - 1;
  const int n_chunks = scale_chunks(n_samples);
  double* partial = static_cast<double*>(d_workspace);
  cudaStream_t s = as_stream(stream);
  dim3 grid(n_chunks, (unsigned)n_chans);
  if (dtype == PARRM_F64) {
    abs_diff_partial_kernel<double><<<grid, kScaleThreads, 0, s>>>(
        static_cast<const double*>(d_x), n_diffs, ld, n_chunks, partial);
  } else if (dtype == PARRM_F32) {
    abs_diff_partial_kernel<float><<<grid, kScaleThreads, 0, s>>>(
        static_cast<const float*>(d_x), n_diffs, ld, n_chunks, partial);
  } else {
    set_error("parrm_channel_scales: bad dtype %d", dtype);
    return PARRM_ERR_INVALID_ARGUMENT;
  }
  PARRM_LAUNCH_OK("abs_diff_partial_kernel");
  scale_finalise_kernel<<<(unsigned)ceil_div(n_chans, 4), 128, 0, s>>>(partial, n_chunks, n_diffs,
                                                                      n_chans, d_scale);
  PARRM_LAUNCH_OK("scale_finalise_kernel");
  return PARRM_OK;
}

int parrm_standardise_gather(const void* d_x, int64_t n_chans, int64_t n_samples, int64_t ld,
                             const int64_t* d_indices, int64_t n_indices, const double* d_scale,
                             double outlier_boundary, double* d_y, int64_t ld_y, double* d_sumsq,
                             int dtype, void* stream) {
  PARRM_NVTX("parrm_standardise_gather");
  using namespace parrm;
  PARRM_REQUIRE(n_chans > 0 && n_chans <= 65535 && n_samples >= 2 && ld >= n_samples,
                "parrm_standardise_gather: bad shape");
  PARRM_REQUIRE(n_indices > 0 && ld_y >= n_chans, "parrm_standardise_gather: bad output shape");
  PARRM_REQUIRE(d_x && d_indices && d_scale && d_y, "parrm_standardise_gather: null pointer");
  cudaStream_t s = as_stream(stream);
  dim3 grid((unsigned)ceil_div(n_indices, 256), (unsigned)n_chans);
  if (dtype == PARRM_F64) {
    standardise_gather_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(d_x), ld,
                                                           d_indices, n_indices, d_scale,
                                                           outlier_boundary, d_y, ld_y);
  } else if (dtype == PARRM_F32) {
    standardise_gather_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_x), ld,
                                                          d_indices, n_indices, d_scale,
                                                          outlier_boundary, d_y, ld_y);
  } else {
    set_error("parrm_standardise_gather: bad dtype %d", dtype);
    return PARRM_ERR_INVALID_ARGUMENT;
  }
  PARRM_LAUNCH_OK("standardise_gather_kernel");
  if (d_sumsq) {
    channel_sumsq_kernel<double><<<(unsigned)n_chans, 256, 0, s>>>(d_y, ld_y, n_indices, d_sumsq);
    PARRM_LAUNCH_OK("channel_sumsq_kernel");
  }
  return PARRM_OK;
}

int parrm_standardise_full(const void* d_x, int64_t n_chans, int64_t n_samples, int64_t ld,
                           const double* d_scale, double outlier_boundary, void* d_z,
                           int64_t ld_z, int dtype, void* stream) {
  PARRM_NVTX("parrm_standardise_full");
  using namespace parrm;
  PARRM_REQUIRE(n_chans > 0 && n_chans <= 65535 && n_samples >= 2 && ld >= n_samples &&
                    ld_z >= n_samples - 1,
                "parrm_standardise_full: bad shape");
  PARRM_REQUIRE(d_x && d_scale && d_z, "parrm_standardise_full: null pointer");
  const int64_t n_diffs = n_samples - 1;
  cudaStream_t s = as_stream(stream);
  dim3 grid((unsigned)min64(ceil_div(n_diffs, 256), 148 * 16), (unsigned)n_chans);
  if (dtype == PARRM_F64) {
    standardise_full_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(d_x), n_diffs,
                                                         ld, d_scale, outlier_boundary,
                                                         static_cast<double*>(d_z), ld_z);
  } else if (dtype == PARRM_F32) {
    standardise_full_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(d_x), n_diffs,
                                                        ld, d_scale, outlier_boundary,
                                                        static_cast<float*>(d_z), ld_z);
  } else {
    set_error("parrm_standardise_full: bad dtype %d", dtype);
    return PARRM_ERR_INVALID_ARGUMENT;
  }
  PARRM_LAUNCH_OK("standardise_full_kernel");
  return PARRM_OK;
}

}  // extern "C"

// sum_j y[j, c]^2 of a sample-major tile (the y'y term of the evaluator's quadratic form) for
// callers that bring an already standardised tile (PARRM._optimise_local's seam) and for the
// float32 storage mode, where the sums must be those of the rounded values the fit sees.
extern "C" int parrm_channel_sumsq(const void* d_y, int y_dtype, int64_t ld_y, int64_t n_chans,
                                   int64_t n_indices, double* d_sumsq, void* stream) {
  PARRM_REQUIRE(n_chans >= 0 && n_chans <= 65535 && n_indices >= 0 && ld_y >= n_chans,
                "parrm_channel_sumsq: bad shape");
  PARRM_REQUIRE(y_dtype == PARRM_F64 || y_dtype == PARRM_F32, "parrm_channel_sumsq: bad dtype");
  if (n_chans == 0) return PARRM_OK;
  PARRM_REQUIRE(d_y != nullptr && d_sumsq != nullptr, "parrm_channel_sumsq: null pointer");
  if (y_dtype == PARRM_F64)
    parrm::channel_sumsq_kernel<double><<<unsigned(n_chans), 256, 0, parrm::as_stream(stream)>>>(
        static_cast<const double*>(d_y), ld_y, n_indices, d_sumsq);
  else
    parrm::channel_sumsq_kernel<float><<<unsigned(n_chans), 256, 0, parrm::as_stream(stream)>>>(
        static_cast<const float*>(d_y), ld_y, n_indices, d_sumsq);
  PARRM_LAUNCH_OK("channel_sumsq_kernel");
  return PARRM_OK;
}
