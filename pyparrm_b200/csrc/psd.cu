// Periodogram behind compute_psd (reference src/pyparrm/_utils/_power.py:10-68), the power
// spectrum the parameter explorer redraws after every re-filter (_plotting.py:568-584,
// 637-642).  The reference calls scipy.fft.fft(data.astype(float32), n_points): the FIRST
// n_points samples of every channel (zero-padded when the recording is shorter), single
// precision, and keeps bins 1 .. n_points/2.  n_points is small (sampling_freq / freq_res,
// tens to a few thousand), so this is a direct DFT: one thread per (channel, bin), the
// samples of the channel and a table of the n_points roots of unity staged in shared memory,
// twiddle index (k * t) mod n_points kept as an exact integer.  Sums run in float64 and the
// coefficient is rounded to float32 once, where the reference's float32 FFT rounds at every
// butterfly; |X|^2 / (fs * n) is then formed in float32 as the reference does.
#include "common.cuh"

namespace parrm {

constexpr int kPsdThreads = 256;
constexpr int kPsdTile = 2048;  // samples staged per pass

template <typename T>
__global__ void __launch_bounds__(kPsdThreads)
periodogram_kernel(const T* __restrict__ x, int64_t ld, int64_t n_samples, int n_points,
                   int n_bins, float scale, float* __restrict__ psd, int64_t ld_psd) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* const root = reinterpret_cast<float2*>(smem_raw);                 // [n_points]
  float* const tile = reinterpret_cast<float*>(smem_raw + size_t(n_points) * sizeof(float2));
  const int64_t chan = blockIdx.y;
  const T* const row = x + chan * ld;
  for (int j = threadIdx.x; j < n_points; j += kPsdThreads) {
    double s, c;
    sincospi(-2.0 * double(j) / double(n_points), &s, &c);   // exp(-2 pi i j / n)
    root[j] = make_float2(float(c), float(s));
  }
  const int n_used = int(min64(n_samples, n_points));         // fft(x, n): crop or zero-pad
  const int k = 1 + blockIdx.x * kPsdThreads + threadIdx.x;   // bin (1-based: DC is dropped)
  double re = 0.0, im = 0.0;
  for (int t0 = 0; t0 < n_used; t0 += kPsdTile) {
    const int n_t = min(kPsdTile, n_used - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < n_t; j += kPsdThreads) tile[j] = float(row[t0 + j]);
    __syncthreads();
    if (k <= n_bins) {
      int idx = int((int64_t(k) * t0) % n_points);
      for (int j = 0; j < n_t; ++j) {
        const float2 w = root[idx];
        const double v = double(tile[j]);
        re = fma(v, double(w.x), re);
        im = fma(v, double(w.y), im);
        idx += k;
        if (idx >= n_points) idx -= n_points;
      }
    }
  }
  if (k <= n_bins) {
    const float fr = float(re), fi = float(im);
    const float mag = hypotf(fr, fi);                         // np.abs(complex64) -> float32
    psd[chan * ld_psd + (k - 1)] = scale * (mag * mag);
  }
}

}  // namespace parrm

extern "C" int parrm_periodogram(const void* d_x, int64_t n_chans, int64_t n_samples, int64_t ld,
                                 int dtype, int64_t n_points, double sampling_freq, float* d_psd,
                                 int64_t ld_psd, void* stream) {
  PARRM_NVTX("parrm_periodogram");
  using namespace parrm;
  PARRM_REQUIRE(n_points >= 2 && n_points <= 16384,
                "parrm_periodogram: n_points must lie in [2, 16384] (got %lld)", (long long)n_points);
  PARRM_REQUIRE(n_chans >= 0 && n_chans <= 65535 && n_samples >= 0 && sampling_freq > 0,
                "parrm_periodogram: bad shape");
  PARRM_REQUIRE(dtype == PARRM_F64 || dtype == PARRM_F32, "parrm_periodogram: bad dtype %d", dtype);
  const int n_bins = int(n_points / 2);
  if (n_chans == 0) return PARRM_OK;
  PARRM_REQUIRE(d_x != nullptr && d_psd != nullptr && ld_psd >= n_bins,
                "parrm_periodogram: null pointer or short output rows");
  const float scale = float(1.0 / (sampling_freq * double(n_points)));
  const size_t smem = size_t(n_points) * sizeof(float2) + kPsdTile * sizeof(float);
  dim3 grid(unsigned(ceil_div(n_bins, kPsdThreads)), unsigned(n_chans));
  if (dtype == PARRM_F64) {
    auto kernel = periodogram_kernel<double>;
    PARRM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kernel<<<grid, kPsdThreads, smem, as_stream(stream)>>>(
        static_cast<const double*>(d_x), ld, n_samples, int(n_points), n_bins, scale, d_psd, ld_psd);
  } else {
    auto kernel = periodogram_kernel<float>;
    PARRM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kernel<<<grid, kPsdThreads, smem, as_stream(stream)>>>(
        static_cast<const float*>(d_x), ld, n_samples, int(n_points), n_bins, scale, d_psd, ld_psd);
  }
  PARRM_LAUNCH_OK("periodogram_kernel");
  return PARRM_OK;
}
