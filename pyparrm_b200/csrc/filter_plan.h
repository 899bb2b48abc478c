// Relocatable filter plan blob shared by host and device (offsets, no pointers).
// Built on the host by parrm_filter_plan(), uploaded verbatim by the caller.
#pragma once
#include <stdint.h>

namespace parrm {

constexpr uint32_t kPlanMagic = 0x4D525250u;  // "PRRM"

struct FilterPlanHeader {
  uint32_t magic;
  uint32_t version;
  int32_t n_taps;
  int32_t w_min, w_max;  // smallest / largest signed tap offset
  int32_t kind;          // 0 = plain tap gather
  int32_t taps_offset;   // byte offset of int32 taps[n_taps]
  int32_t dtype;         // parrm_dtype_t the plan was built for
};
static_assert(sizeof(FilterPlanHeader) == 32, "plan header is 32 bytes");

}  // namespace parrm
