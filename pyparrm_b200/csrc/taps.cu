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

__global__ void __launch_bounds__(kTapThreads)
build_taps_kernel(double period, double phw, int64_t hw, int64_t omit, int direction,
                  int32_t* __restrict__ taps, int32_t* __restrict__ n_taps) {
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

}  // namespace parrm

extern "C" int parrm_build_taps(double period, double period_half_width,
                                int64_t filter_half_width, int64_t omit_n_samples, int direction,
                                int32_t* d_taps, int32_t* d_n_taps, void* stream) {
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
