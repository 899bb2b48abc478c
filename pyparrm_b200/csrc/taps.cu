// Tap builder: device restatement of PARRM._generate_filter's mask (parrm.py:803-820).
// Integer-exact: the same IEEE operations NumPy performs (fmod, one add, one subtract,
// comparisons), then an order-preserving compaction of the surviving window offsets.
#include "common.cuh"

namespace parrm {

constexpr int kTapThreads = 1024;

// NumPy's float remainder (npy_divmod): fmod, then shift into [0, b) for b > 0.
__device__ __forceinline__ double numpy_mod_pos(double a, double b) {
  double r = fmod(a, b);  // exact in CUDA (0 ulp)
  if (r != 0.0) {
    if (r < 0.0) r = __dadd_rn(r, b);
  } else {
    r = 0.0;  // copysign(0, b), b > 0
  }
  return r;
}

__device__ __forceinline__ void build_taps_cta(double period, double phw, int64_t hw, int64_t omit,
                                               int direction, int32_t* __restrict__ taps,
                                               int32_t* __restrict__ n_taps) {
  __shared__ int warp_count[kTapThreads / 32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  const double upper = __dsub_rn(period, phw);  // parrm.py:812
  for (int64_t start = -hw; start <= hw; start += kTapThreads) {
    const int64_t w = start + tid;
    bool keep = false;
    if (w <= hw) {
      const double r = numpy_mod_pos(static_cast<double>(w), period);
      const int64_t aw = w < 0 ? -w : w;
      keep = (r <= phw || r >= upper) && (aw > omit);
      if (direction == PARRM_DIR_PAST && w > 0) keep = false;     // parrm.py:817-818
      if (direction == PARRM_DIR_FUTURE && w <= 0) keep = false;  // parrm.py:819-820
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    int offset = base;
    for (int i = 0; i < warp; ++i) offset += warp_count[i];
    if (keep) taps[offset + __popc(ballot & ((1u << lane) - 1u))] = static_cast<int32_t>(w);
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int i = 0; i < kTapThreads / 32; ++i) total += warp_count[i];
      base += total;
    }
    __syncthreads();
  }
  if (tid == 0) *n_taps = base;
}

__global__ void __launch_bounds__(kTapThreads)
build_taps_kernel(double period, double phw, int64_t hw, int64_t omit, int direction,
                  int32_t* __restrict__ taps, int32_t* __restrict__ n_taps) {
  build_taps_cta(period, phw, hw, omit, direction, taps, n_taps);
}

// One CTA per parameter set: the whole sweep of an explorer session in one launch.
__global__ void __launch_bounds__(kTapThreads)
build_taps_batch_kernel(const double* __restrict__ period, const double* __restrict__ phw,
                        const int64_t* __restrict__ hw, const int64_t* __restrict__ omit,
                        const int32_t* __restrict__ direction, int32_t* __restrict__ taps,
                        int64_t stride, int32_t* __restrict__ n_taps) {
  const int64_t s = blockIdx.x;
  build_taps_cta(period[s], phw[s], hw[s], omit[s], direction[s], taps + s * stride, n_taps + s);
}

// Default filter half-width (parrm.py:788-801): walk hw = omit + 1, omit + 2, ... counting the
// offsets whose phase  mod(hw, period)  is <= phw  or  >= period + phw  (the second clause as
// the reference writes it: never true) until 50 are found or hw reaches the limit.  One CTA
// per parameter set scans blocks of 1024 offsets; ballots locate the 50th match.
__global__ void __launch_bounds__(kTapThreads)
default_half_width_kernel(const double* __restrict__ period, const double* __restrict__ phw,
                          const int64_t* __restrict__ omit, const int64_t* __restrict__ limit,
                          int needed, int64_t* __restrict__ out) {
  __shared__ int warp_count[kTapThreads / 32];
  __shared__ int found_total;
  __shared__ long long answer;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t s = blockIdx.x;
  const double per = period[s], half = phw[s];
  const double upper = __dadd_rn(per, half);
  const int64_t lim = limit[s];
  if (tid == 0) {
    found_total = 0;
    answer = -1;
  }
  __syncthreads();
  int64_t start = omit[s];
  while (start < lim) {
    const int64_t hw = start + 1 + tid;
    bool hit = false;
    if (hw <= lim) {
      const double m = numpy_mod_pos(static_cast<double>(hw), per);
      hit = (m <= half) || (m >= upper);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_count[warp] = __popc(ballot);
    __syncthreads();
    int before = found_total;
    for (int i = 0; i < warp; ++i) before += warp_count[i];
    const int rank = before + __popc(ballot & ((1u << lane) - 1u)) + 1;  // 1-based, if hit
    if (hit && rank == needed) answer = hw;
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int i = 0; i < kTapThreads / 32; ++i) total += warp_count[i];
      found_total += total;
    }
    __syncthreads();
    if (answer >= 0) break;
    start += kTapThreads;
  }
  if (tid == 0) out[s] = answer >= 0 ? int64_t(answer) : max64(lim, omit[s]);
}

}  // namespace parrm

extern "C" int parrm_build_taps(double period, double period_half_width,
                                int64_t filter_half_width, int64_t omit_n_samples, int direction,
                                int32_t* d_taps, int32_t* d_n_taps, void* stream) {
  PARRM_NVTX("parrm_build_taps");
  PARRM_REQUIRE(period > 0.0, "parrm_build_taps: period must be > 0");
  PARRM_REQUIRE(filter_half_width >= 0 && filter_half_width < (int64_t(1) << 30),
                "parrm_build_taps: filter_half_width out of range");
  PARRM_REQUIRE(direction >= PARRM_DIR_BOTH && direction <= PARRM_DIR_FUTURE,
                "parrm_build_taps: unknown direction %d", direction);
  PARRM_REQUIRE(d_taps != nullptr && d_n_taps != nullptr, "parrm_build_taps: null output");
  parrm::build_taps_kernel<<<1, parrm::kTapThreads, 0, parrm::as_stream(stream)>>>(
      period, period_half_width, filter_half_width, omit_n_samples, direction, d_taps, d_n_taps);
  PARRM_LAUNCH_OK("build_taps_kernel");
  return PARRM_OK;
}

extern "C" int parrm_build_taps_batch(const double* d_period, const double* d_period_half_width,
                                      const int64_t* d_filter_half_width,
                                      const int64_t* d_omit_n_samples, const int32_t* d_direction,
                                      int64_t n_sets, int32_t* d_taps, int64_t stride,
                                      int32_t* d_n_taps, void* stream) {
  PARRM_NVTX("parrm_build_taps_batch");
  PARRM_REQUIRE(n_sets >= 0 && n_sets <= 65535, "parrm_build_taps_batch: 0..65535 parameter sets");
  if (n_sets == 0) return PARRM_OK;
  PARRM_REQUIRE(d_period && d_period_half_width && d_filter_half_width && d_omit_n_samples &&
                    d_direction && d_taps && d_n_taps && stride > 0,
                "parrm_build_taps_batch: null pointer or empty rows");
  parrm::build_taps_batch_kernel<<<unsigned(n_sets), parrm::kTapThreads, 0,
                                   parrm::as_stream(stream)>>>(
      d_period, d_period_half_width, d_filter_half_width, d_omit_n_samples, d_direction, d_taps,
      stride, d_n_taps);
  PARRM_LAUNCH_OK("build_taps_batch_kernel");
  return PARRM_OK;
}

extern "C" int parrm_default_half_width(const double* d_period, const double* d_period_half_width,
                                        const int64_t* d_omit_n_samples, const int64_t* d_limit,
                                        int64_t n_sets, int64_t* d_half_width, void* stream) {
  PARRM_NVTX("parrm_default_half_width");
  PARRM_REQUIRE(n_sets >= 0 && n_sets <= 65535, "parrm_default_half_width: 0..65535 parameter sets");
  if (n_sets == 0) return PARRM_OK;
  PARRM_REQUIRE(d_period && d_period_half_width && d_omit_n_samples && d_limit && d_half_width,
                "parrm_default_half_width: null pointer");
  parrm::default_half_width_kernel<<<unsigned(n_sets), parrm::kTapThreads, 0,
                                     parrm::as_stream(stream)>>>(
      d_period, d_period_half_width, d_omit_n_samples, d_limit, 50, d_half_width);
  PARRM_LAUNCH_OK("default_half_width_kernel");
  return PARRM_OK;
}
