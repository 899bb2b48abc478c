// Relocatable filter plan blob shared by host and device (offsets, no pointers).
// Built on the host by parrm_filter_plan(), uploaded verbatim by the caller.
//
// A plan is an exact integer re-association of the tap set T of PARRM._generate_filter
// (parrm.py:803-833).  Every tap has the same weight -1/n_taps (parrm.py:829), so
//
//     sum_{w in T} x[t - w]  =  sum_b D_{k(b)}[t - a_b]  +  sum_{w in plus} x[t - w]
//                               -  sum_{w in minus} x[t - w]  +  centre * x[t]
//
// with "comb boxes"  D_k[i] = sum_{q=0}^{m_k - 1} x[i - q*d]  of at most two window lengths
// m_0, m_1 over one common stride d.  The taps of a phase-matched comb sit near the multiples
// of the period, so for d ~ an integer multiple of the period they fall into a few long
// arithmetic progressions.  The kernel sums over the boxes first (a short "pattern"
// E_k[i] = sum_b x[i - a_kb], one shared-memory load per box) and runs the comb as a sliding
// sum of E_k along the chain t, t + d, t + 2d, ... in registers.  The identity is over
// integers (which offsets are summed), so the result differs from the plain gather only by
// floating-point association.
#pragma once
#include <stdint.h>

namespace parrm {

constexpr uint32_t kPlanMagic = 0x4D525250u;  // "PRRM"
constexpr uint32_t kPlanVersion = 4;
constexpr int kMaxTerms = 120;                // structured terms passed as kernel parameters
constexpr int kMaxBoxKinds = 2;

enum PlanKind : int32_t {
  kPlanGather = 0,  // one shared-memory load per tap
  kPlanComb = 1     // comb boxes + single taps (see above); taps are kept for the gather
};

struct FilterPlanHeader {
  uint32_t magic;
  uint32_t version;
  int32_t n_taps;
  int32_t w_min, w_max;   // smallest / largest signed tap offset
  int32_t kind;           // PlanKind
  int32_t taps_offset;    // byte offset of int32 taps[n_taps]
  int32_t dtype;          // parrm_dtype_t the plan was built for
  // ---- kPlanComb ----
  int32_t stride;                   // d
  int32_t n_kinds;                  // 1 or 2 box lengths in use
  int32_t window[kMaxBoxKinds];     // m_k
  int32_t n_box[kMaxBoxKinds];      // boxes of each length
  int32_t a_min[kMaxBoxKinds];      // smallest / largest box offset of each length
  int32_t a_max[kMaxBoxKinds];
  int32_t n_plus, n_minus;          // single taps with coefficient +1 / -1
  int32_t centre;                   // coefficient of x[t] inside the tap sum (0 or negative)
  int32_t terms_offset;             // byte offset of int32 terms[]: boxes kind 0, boxes kind 1,
                                    // plus singles, minus singles
  int32_t cost_milli;               // modelled shared-memory loads per output x 1000
  // ---- recording edges (every kind) ----
  int32_t count_offset;             // byte offset of int32 count[w_hi - w_lo + 2]:
                                    //   count[v - (w_lo - 1)] = #{taps <= v},  w_lo - 1 <= v <= w_hi
                                    //   (w_lo = min(w_min, 0), w_hi = max(w_max, 0))
  int32_t recip_offset;             // byte offset of float64 recip[n_taps + 1]: 0, 1/1, 1/2, ...
  int32_t total_bytes;              // size of the whole blob
  int32_t reserved[6];
};
static_assert(sizeof(FilterPlanHeader) == 128, "plan header is 128 bytes");

// The decomposition is chosen for the run-time specialised kernel (filter_comb_e.cuh): one
// thread per residue of the stride, so 64 <= d <= 992; every term is one shared-memory load
// per output, so cost = number of terms; the box lengths are bounded by the register rings.
constexpr int kPatternFirstMinStride = 64;
constexpr int kPatternFirstMaxStride = 992;
constexpr int kPatternFirstMaxRing = 60;   // M0 + M1
constexpr int kPatternFirstStepRegs = 52;  // registers a step needs besides the rings

// Threads of the specialised kernel's CTA (one per chain, whole warps, plus the producer warp)
// and the registers each may use with `ctas` CTAs resident per SM.
inline int pattern_first_threads(int d) { return ((d + 31) / 32) * 32 + 32; }
inline int pattern_first_reg_budget(int d, int ctas) {
  const int r = (65536 / (ctas * pattern_first_threads(d))) / 8 * 8;
  return r < 255 ? r : 255;
}
// Longest rings (M0 + M1 pattern values of `es` bytes per chain) the kernel can hold in
// registers at stride d: a wide stride is a wide CTA, which leaves few registers per thread.
// The planner only proposes plans inside this limit, so that what it returns is runnable.
inline int pattern_first_max_ring(int d, int es) {
  const int ring = (pattern_first_reg_budget(d, 1) - kPatternFirstStepRegs) / (es / 4);
  return ring < kPatternFirstMaxRing ? ring : kPatternFirstMaxRing;
}

}  // namespace parrm
