// Host-side filter planner: turns the ascending tap offsets of PARRM._generate_filter
// (parrm.py:803-833) into the relocatable plan blob of filter_plan.h.  Pure integer work on
// the host; no CUDA calls (the file is compiled by nvcc only to share the build line).
#include <algorithm>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "filter_plan.h"

namespace parrm {
namespace {

struct CombPlan {
  int d = 0, nk = 0;
  int m[kMaxBoxKinds] = {0, 0};
  std::vector<int> box[kMaxBoxKinds];
  std::vector<int> plus, minus;
  int centre = 0;
  double cost = 1e30;
  int n_terms() const { return int(box[0].size() + box[1].size() + plus.size() + minus.size()); }
};

// bit[w - lo] = 1 for taps, lo <= 0 <= hi.  Exact cover of every residue class of stride d by
// boxes of the given lengths plus +/- single taps, minimising loads (dynamic programme).
bool cover(const std::vector<uint8_t>& bit, int lo, int hi, int d, int nk, const int* m,
           CombPlan* out) {
  const int span = hi - lo + 1;
  CombPlan p;
  p.d = d;
  p.nk = nk;
  for (int k = 0; k < nk; ++k) p.m[k] = m[k];
  std::vector<int> dp, zeros, choice;
  for (int c = 0; c < d && c < span; ++c) {
    const int len = (span - 1 - c) / d + 1;
    dp.assign(len + 1, 0);
    zeros.assign(len + 1, 0);
    choice.assign(len + 1, -1);
    auto w_of = [&](int pos) { return lo + c + pos * d; };
    for (int q = 0; q < len; ++q) {
      const int w = w_of(q);
      const bool free_zero = (w == 0);  // the centre sample is loaded anyway
      zeros[q + 1] = zeros[q] + ((bit[w - lo] == 0 && !free_zero) ? 1 : 0);
    }
    // forward DP over "prefix [0, q) represented"
    const int kInf = 1 << 28;
    for (int q = 1; q <= len; ++q) dp[q] = kInf;
    for (int q = 0; q < len; ++q) {
      if (dp[q] >= kInf) continue;
      const int single = dp[q] + (bit[w_of(q) - lo] ? 1 : 0);
      if (single < dp[q + 1]) {
        dp[q + 1] = single;
        choice[q + 1] = -1;
      }
      for (int k = 0; k < nk; ++k) {
        if (q + m[k] > len) continue;
        const int v = dp[q] + 1 + (zeros[q + m[k]] - zeros[q]);
        if (v < dp[q + m[k]]) {
          dp[q + m[k]] = v;
          choice[q + m[k]] = k;
        }
      }
    }
    // backtrack
    for (int q = len; q > 0;) {
      const int k = choice[q];
      if (k < 0) {
        if (bit[w_of(q - 1) - lo]) p.plus.push_back(w_of(q - 1));
        --q;
      } else {
        const int q0 = q - m[k];
        p.box[k].push_back(w_of(q0));
        for (int r = q0; r < q; ++r) {
          const int w = w_of(r);
          if (bit[w - lo]) continue;
          if (w == 0) --p.centre; else p.minus.push_back(w);
        }
        q = q0;
      }
    }
  }
  for (int k = 0; k < kMaxBoxKinds; ++k) std::sort(p.box[k].begin(), p.box[k].end());
  std::sort(p.plus.begin(), p.plus.end());
  std::sort(p.minus.begin(), p.minus.end());
  // drop an unused second length; keep kind 0 the one in use
  if (p.nk == 2 && p.box[1].empty()) p.nk = 1;
  if (p.nk == 2 && p.box[0].empty()) {
    p.box[0].swap(p.box[1]);
    p.m[0] = p.m[1];
    p.nk = 1;
  }
  if (p.box[0].empty()) return false;
  p.cost = double(p.n_terms());
  *out = p;
  return true;
}

// Verifies box/plus/minus/centre expand to exactly the tap indicator.
bool exact(const CombPlan& p, const std::vector<uint8_t>& bit, int lo, int hi) {
  std::vector<int> acc(hi - lo + 1, 0);
  for (int k = 0; k < p.nk; ++k)
    for (int a : p.box[k])
      for (int q = 0; q < p.m[k]; ++q) {
        const int w = a + q * p.d;
        if (w < lo || w > hi) return false;
        ++acc[w - lo];
      }
  for (int w : p.plus) ++acc[w - lo];
  for (int w : p.minus) --acc[w - lo];
  acc[0 - lo] += p.centre;
  for (int w = lo; w <= hi; ++w)
    if (acc[w - lo] != int(bit[w - lo])) return false;
  return true;
}

// Searches comb decompositions over strides d_lo..d_hi under the cost model of the run-time
// specialised kernel (filter_comb_e.cuh): loads per output = number of terms, whatever the
// number of box lengths, and the two box lengths together must fit the register rings
// (pattern_first_max_ring values of `es` bytes per chain at that stride).
bool best_comb(const int32_t* taps, int n, int d_lo, int d_hi, int es, CombPlan* best) {
  const int lo = std::min(taps[0], 0), hi = std::max(taps[n - 1], 0);
  const int64_t span = int64_t(hi) - lo + 1;
  if (n < 8 || span > (1 << 20)) return false;
  std::vector<uint8_t> bit(span, 0);
  for (int i = 0; i < n; ++i) bit[taps[i] - lo] = 1;
  // candidate strides: fewest maximal progressions
  int64_t d_max = std::min<int64_t>(span - 1, 8192);
  d_max = std::min<int64_t>(d_max, std::max<int64_t>(64, 40000000 / n));
  d_max = std::min<int64_t>(d_max, d_hi);
  std::vector<std::pair<int, int>> ranked;  // (progressions, d)
  for (int d = d_lo; d <= d_max; ++d) {
    int chains = 0;
    for (int i = 0; i < n; ++i) {
      const int64_t prev = int64_t(taps[i]) - d;
      if (prev < lo || !bit[prev - lo]) ++chains;
    }
    ranked.emplace_back(chains, d);
  }
  std::sort(ranked.begin(), ranked.end());
  bool found = false;
  const int n_strides = std::min<int>(12, int(ranked.size()));
  for (int r = 0; r < n_strides; ++r) {
    const int d = ranked[r].second;
    // progression lengths, with and without bridging the (absent) centre tap
    std::vector<std::pair<int64_t, int>> lengths;  // (taps covered, length)
    for (int bridge = 0; bridge < 2; ++bridge) {
      auto has = [&](int64_t w) {
        if (w < lo || w > hi) return false;
        return bit[w - lo] != 0 || (bridge && w == 0);
      };
      for (int64_t w = lo; w <= hi; ++w) {
        if (!has(w) || has(w - d)) continue;
        int len = 1;
        while (has(w + int64_t(len) * d)) ++len;
        if (len < 2) continue;
        bool seen = false;
        for (auto& e : lengths)
          if (e.second == len) {
            e.first += len;
            seen = true;
          }
        if (!seen) lengths.emplace_back(len, len);
      }
    }
    std::sort(lengths.rbegin(), lengths.rend());
    // what the kernel can keep in registers at this stride (a wide stride is a wide CTA)
    const int max_ring = pattern_first_max_ring(d, es);
    if (max_ring < 2) continue;
    // candidate box lengths: the progressions' own lengths and, where those do not fit the
    // rings, halves and thirds of them (two or three boxes per progression)
    std::vector<int> cand;
    auto add_len = [&](int len) {
      if (len >= 2 && len <= max_ring && std::find(cand.begin(), cand.end(), len) == cand.end())
        cand.push_back(len);
    };
    for (size_t i = 0; i < lengths.size() && i < 6; ++i) add_len(lengths[i].second);
    for (size_t i = 0; i < lengths.size() && i < 3 && cand.size() < 8; ++i) {
      const int len = lengths[i].second;
      if (len > max_ring || 2 * len > max_ring) {
        add_len((len + 1) / 2);
        add_len((len + 2) / 3);
      }
    }
    const int n_len = int(cand.size());
    for (int i = 0; i < n_len; ++i) {
      for (int j = i; j < n_len; ++j) {
        int m[2] = {cand[i], cand[j]};
        if (i != j && m[0] + m[1] > max_ring) continue;
        CombPlan p;
        if (!cover(bit, lo, hi, d, i == j ? 1 : 2, m, &p)) continue;
        if (p.n_terms() > kMaxTerms) continue;
        const int ring = p.m[0] + (p.nk > 1 ? p.m[1] : 0);
        if (ring > max_ring) continue;
        p.cost = double(p.n_terms()) + 1e-3 * ring;  // fewer loads, then fewer registers
        if (p.cost < best->cost) {
          *best = p;
          found = true;
        }
      }
    }
  }
  if (!found) return false;
  return exact(*best, bit, lo, hi);
}

}  // namespace
}  // namespace parrm

extern "C" {

// blob layout: header | taps[n] | terms[kMaxTerms] | (8-byte aligned) recip[n + 1] | count[span + 2]
static size_t plan_layout(const int32_t* h_taps, int32_t n_taps, size_t* recip_off, size_t* count_off) {
  const size_t n = size_t(n_taps > 0 ? n_taps : 0);
  size_t off = sizeof(parrm::FilterPlanHeader) + (n + parrm::kMaxTerms) * sizeof(int32_t);
  off = (off + 7) & ~size_t(7);
  if (recip_off) *recip_off = off;
  off += (n + 1) * sizeof(double);
  if (count_off) *count_off = off;
  int64_t span = 0;
  if (n > 0 && h_taps != nullptr) {
    const int64_t lo = h_taps[0] < 0 ? h_taps[0] : 0;
    const int64_t hi = h_taps[n - 1] > 0 ? h_taps[n - 1] : 0;
    span = hi - lo;
  }
  off += size_t(span + 2) * sizeof(int32_t);
  return (off + 15) & ~size_t(15);
}

size_t parrm_filter_plan_bytes(const int32_t* h_taps, int32_t n_taps) {
  return plan_layout(h_taps, n_taps, nullptr, nullptr);
}

int parrm_filter_plan(const int32_t* h_taps, int32_t n_taps, int dtype, int strategy,
                      void* h_plan, size_t plan_bytes) {
  using namespace parrm;
  PARRM_REQUIRE(n_taps > 0 && h_taps != nullptr, "parrm_filter_plan: empty tap list");
  PARRM_REQUIRE(h_plan != nullptr, "parrm_filter_plan: null plan buffer");
  PARRM_REQUIRE(dtype == PARRM_F64 || dtype == PARRM_F32, "parrm_filter_plan: bad dtype");
  PARRM_REQUIRE(strategy >= PARRM_PLAN_AUTO && strategy <= PARRM_PLAN_COMB,
                "parrm_filter_plan: unknown strategy %d", strategy);
  for (int32_t i = 1; i < n_taps; ++i)
    PARRM_REQUIRE(h_taps[i] > h_taps[i - 1], "parrm_filter_plan: taps must be strictly ascending");
  PARRM_REQUIRE(h_taps[0] > -(1 << 30) && h_taps[n_taps - 1] < (1 << 30),
                "parrm_filter_plan: tap offset out of range");
  size_t recip_off = 0, count_off = 0;
  const size_t total = plan_layout(h_taps, n_taps, &recip_off, &count_off);
  PARRM_REQUIRE(plan_bytes >= total, "parrm_filter_plan: plan buffer too small (%zu < %zu)",
                plan_bytes, total);
  PARRM_REQUIRE(total < (size_t(1) << 31), "parrm_filter_plan: tap window too wide");
  memset(h_plan, 0, total);
  FilterPlanHeader* hdr = static_cast<FilterPlanHeader*>(h_plan);
  hdr->magic = kPlanMagic;
  hdr->version = kPlanVersion;
  hdr->n_taps = n_taps;
  hdr->w_min = h_taps[0];
  hdr->w_max = h_taps[n_taps - 1];
  hdr->kind = kPlanGather;
  hdr->taps_offset = int32_t(sizeof(FilterPlanHeader));
  hdr->dtype = dtype;
  hdr->terms_offset = hdr->taps_offset + n_taps * int32_t(sizeof(int32_t));
  hdr->cost_milli = n_taps * 1000;
  unsigned char* base = static_cast<unsigned char*>(h_plan);
  memcpy(base + hdr->taps_offset, h_taps, size_t(n_taps) * sizeof(int32_t));
  {  // edge tables: in-range tap counts by lookup, reciprocals of the possible counts
    hdr->recip_offset = int32_t(recip_off);
    hdr->count_offset = int32_t(count_off);
    hdr->total_bytes = int32_t(total);
    double* recip = reinterpret_cast<double*>(base + recip_off);
    recip[0] = 0.0;
    for (int32_t i = 1; i <= n_taps; ++i) recip[i] = 1.0 / double(i);
    int32_t* count = reinterpret_cast<int32_t*>(base + count_off);
    const int64_t w_lo = h_taps[0] < 0 ? h_taps[0] : 0;
    const int64_t w_hi = h_taps[n_taps - 1] > 0 ? h_taps[n_taps - 1] : 0;
    int32_t i = 0;
    for (int64_t v = w_lo - 1; v <= w_hi; ++v) {
      while (i < n_taps && h_taps[i] <= v) ++i;
      count[v - (w_lo - 1)] = i;
    }
  }
  if (strategy == PARRM_PLAN_GATHER) return PARRM_OK;

  CombPlan best;
  const bool ok = best_comb(h_taps, n_taps, kPatternFirstMinStride, kPatternFirstMaxStride,
                            dtype == PARRM_F64 ? 8 : 4, &best);
  const bool worth = ok && (strategy == PARRM_PLAN_COMB || best.cost < 0.6 * double(n_taps));
  if (!worth) {
    if (strategy == PARRM_PLAN_COMB) {
      set_error("parrm_filter_plan: this tap set has no comb structure");
      return PARRM_ERR_UNSUPPORTED;
    }
    return PARRM_OK;
  }
  hdr->kind = kPlanComb;
  hdr->stride = best.d;
  hdr->n_kinds = best.nk;
  int32_t* terms = reinterpret_cast<int32_t*>(base + hdr->terms_offset);
  int n = 0;
  for (int k = 0; k < best.nk; ++k) {
    hdr->window[k] = best.m[k];
    hdr->n_box[k] = int32_t(best.box[k].size());
    hdr->a_min[k] = best.box[k].front();
    hdr->a_max[k] = best.box[k].back();
    for (int a : best.box[k]) terms[n++] = a;
  }
  hdr->n_plus = int32_t(best.plus.size());
  hdr->n_minus = int32_t(best.minus.size());
  for (int w : best.plus) terms[n++] = w;
  for (int w : best.minus) terms[n++] = w;
  hdr->centre = best.centre;
  hdr->cost_milli = int32_t(best.cost * 1000.0);
  return PARRM_OK;
}

}  // extern "C"

extern "C" int parrm_filter_plan_info(const void* h_plan, int32_t* info, int32_t* terms,
                                      int32_t capacity) {
  using namespace parrm;
  PARRM_REQUIRE(h_plan != nullptr && info != nullptr, "parrm_filter_plan_info: null pointer");
  const FilterPlanHeader* hdr = static_cast<const FilterPlanHeader*>(h_plan);
  PARRM_REQUIRE(hdr->magic == kPlanMagic && hdr->version == kPlanVersion,
                "parrm_filter_plan_info: not a filter plan");
  const int32_t n_terms = hdr->kind == kPlanComb
                              ? hdr->n_box[0] + hdr->n_box[1] + hdr->n_plus + hdr->n_minus
                              : 0;
  const int32_t v[16] = {hdr->kind, hdr->stride, hdr->n_kinds, hdr->window[0], hdr->window[1],
                         hdr->n_box[0], hdr->n_box[1], hdr->n_plus, hdr->n_minus, hdr->centre,
                         hdr->cost_milli, hdr->n_taps, hdr->w_min, hdr->w_max, n_terms, 0};
  for (int i = 0; i < 16; ++i) info[i] = v[i];
  if (terms != nullptr) {
    PARRM_REQUIRE(capacity >= n_terms, "parrm_filter_plan_info: terms buffer too small");
    const int32_t* src = reinterpret_cast<const int32_t*>(
        static_cast<const unsigned char*>(h_plan) + hdr->terms_offset);
    for (int i = 0; i < n_terms; ++i) terms[i] = src[i];
  }
  return PARRM_OK;
}
