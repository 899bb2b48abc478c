// Relocatable filter plan blob shared by host and device (offsets, no pointers).
// Built on the host by parrm_filter_plan(), uploaded verbatim by the caller.
#pragma once
#include <stdint.h>

namespace parrm {

constexpr uint32_t kPlanMagic = 0x4D525250u;  // "PRRM"
constexpr uint32_t kPlanVersion = 2;
constexpr int kMaxTerms = 96;                 // structured terms passed as kernel parameters

enum PlanKind : int32_t {
  kPlanGather = 0,      // y = x[t] - mean of the in-range taps, one shared-memory load per tap
  kPlanStridePrefix = 1 // taps grouped into arithmetic progressions of one common stride d;
                        // each progression costs two loads from a stride-d prefix sum
};

struct FilterPlanHeader {
  uint32_t magic;
  uint32_t version;
  int32_t n_taps;
  int32_t w_min, w_max;  // smallest / largest signed tap offset
  int32_t kind;          // PlanKind
  int32_t taps_offset;   // byte offset of int32 taps[n_taps]
  int32_t dtype;         // parrm_dtype_t the plan was built for
  int32_t stride;        // common difference d of the progressions (kind 1)
  int32_t n_terms;       // number of (offset, coefficient) terms (kind 1)
  int32_t off_offset;    // byte offset of int32 term_off[n_terms]
  int32_t coef_offset;   // byte offset of double term_coef[n_terms]
  int32_t n_progressions;
  int32_t reserved[3];
};
static_assert(sizeof(FilterPlanHeader) == 64, "plan header is 64 bytes");

}  // namespace parrm
