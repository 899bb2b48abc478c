// Run-time specialised comb kernel (filter_comb_e.cuh through NVRTC): host interface used by
// filter.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "filter_plan.h"

namespace parrm {

// Optional launch tuning (parrm_filter_options_t in parrm_b200.h); zero = library default.
struct FilterTuning {
  int kernel;           // PARRM_FILTER_KERNEL_*
  int steps_per_chunk;  // U
  int prefetch_chunks;  // chunks in flight beyond the window
  int ctas_per_sm;
  int variant;          // source variant for A/B measurements (PE_VARIANT)
  unsigned long long timeline;  // device pointer of the per-CTA timeline (variant bit 1), or 0
};

struct CombEShape {
  int es;                 // element bytes
  int d, nk, m[2], nb[2];
  const int32_t* off[2];  // box offsets per kind (point into the plan's term list)
  int n_plus, n_minus;
  const int32_t *plus, *minus;
  int centre, n_taps, w_lo, w_hi, back, fwd;
  int u, pf, ctas, smem_bytes, variant;
};

bool comb_e_shape(const FilterPlanHeader* hdr, const int32_t* terms, int dtype,
                  const FilterTuning* tune, CombEShape* out);
std::string comb_e_key(const CombEShape& s, int dev);
bool comb_e_cached(const CombEShape& s);
int comb_e_compile_only(const CombEShape& s, size_t* cubin_bytes);
int launch_comb_e(const CombEShape& s, const void* d_x, void* d_out, const int32_t* d_taps,
                  const int32_t* d_count, const double* d_recip,
                  int64_t ld_x, int64_t x_t0, int64_t n_x, int64_t ld_out, int64_t t0,
                  int64_t n_out, int64_t n_total, int64_t n_chans, cudaStream_t stream,
                  int* regs_out, unsigned long long timeline = 0);

}  // namespace parrm
