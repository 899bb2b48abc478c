// Period filter: body of PARRM.filter_data (parrm.py:861-869) as a direct gather.
//
//   y[c,t] = x[c,t] - (1/n_in(t)) * sum_{w in taps, 0 <= t-w < T} x[c,t-w];   0 where n_in(t) = 0
//
// The reference evaluates this with two FFT convolutions (the second one, of an all-ones
// array, only counts the in-range taps).  Here every CTA stages one time tile plus its halo
// [t0 - w_max, t0 + tile - w_min) of one channel into shared memory with a single TMA bulk
// copy (cp.async.bulk, SASS UBLKCP) and gathers the taps from there.  HBM traffic is the
// algorithmic 2 * sizeof(T) bytes per channel-sample; the halo re-reads of neighbouring
// tiles are served by L2.
#include "filter_plan.h"

#include "common.cuh"

namespace parrm {

template <typename T>
struct FilterArgs {
  const T* x;
  T* out;
  const int32_t* taps;  // device, ascending
  int64_t ld_x, x_t0, n_x;
  int64_t ld_out, t0, n_out;
  int64_t n_total;
  int32_t n_taps, w_lo, w_hi;  // w_lo = min(w_min, 0), w_hi = max(w_max, 0)
  int32_t tile;
};

constexpr int kFilterThreads = 256;
constexpr int kOutPerThread = 4;

__host__ __device__ inline int round16(int bytes) { return (bytes + 15) & ~15; }

// ---- shared-memory gather --------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFilterThreads)
filter_gather_smem_kernel(const FilterArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  int32_t* s_taps = reinterpret_cast<int32_t*>(smem_raw + 16);
  T* s_win = reinterpret_cast<T*>(smem_raw + 16 + round16(a.n_taps * 4));
  constexpr int VEC = 16 / sizeof(T);

  const int tid = threadIdx.x;
  const int64_t chan = blockIdx.y;
  const int64_t tile_t0 = a.t0 + int64_t(blockIdx.x) * a.tile;
  const int n_tile = int(min(int64_t(a.tile), a.t0 + a.n_out - tile_t0));
  // window of global sample times held in shared memory
  const int64_t g_lo = tile_t0 - a.w_hi;
  const int64_t g_hi = tile_t0 + n_tile - a.w_lo;
  const int64_t v_lo = max(g_lo, max(int64_t(0), a.x_t0));
  const int64_t v_hi = min(g_hi, min(a.n_total, a.x_t0 + a.n_x));
  const T* xrow = a.x + chan * a.ld_x - a.x_t0;  // xrow[g] = sample at global time g

  // Element shift of the window so that 16-byte aligned global addresses land on 16-byte
  // aligned shared addresses (bulk-copy requirement).
  const int g_mis = int((reinterpret_cast<uintptr_t>(xrow + v_lo) / sizeof(T)) % VEC);
  const int shift = (g_mis - int((v_lo - g_lo) % VEC) + VEC) % VEC;
  T* s_x = s_win + shift;  // s_x[g - g_lo]
  const int n_valid = int(v_hi - v_lo);
  const int head = min((VEC - g_mis) % VEC, n_valid);
  const int n_bulk = ((n_valid - head) / VEC) * VEC;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0 && n_bulk > 0) {
    mbar_expect_tx(bar, uint32_t(n_bulk) * sizeof(T));
    bulk_g2s(s_x + (v_lo - g_lo) + head, xrow + v_lo + head, uint32_t(n_bulk) * sizeof(T), bar);
  }
  for (int i = tid; i < a.n_taps; i += kFilterThreads) s_taps[i] = a.taps[i];
  // scalar head / tail around the bulk copy, zero fill outside the recording
  if (tid < head) s_x[(v_lo - g_lo) + tid] = xrow[v_lo + tid];
  for (int i = head + n_bulk + tid; i < n_valid; i += kFilterThreads)
    s_x[(v_lo - g_lo) + i] = xrow[v_lo + i];
  for (int i = tid; i < int(v_lo - g_lo); i += kFilterThreads) s_x[i] = T(0);
  for (int i = int(v_hi - g_lo) + tid; i < int(g_hi - g_lo); i += kFilterThreads) s_x[i] = T(0);
  __syncthreads();
  if (n_bulk > 0) mbar_wait(bar, 0);

  const T* s_c = s_x + a.w_hi;  // s_c[i] = sample at tile_t0 + i
  T* orow = a.out + chan * a.ld_out + (tile_t0 - a.t0);
  const bool interior = (g_lo >= 0) && (g_hi <= a.n_total);
  const int n_taps = a.n_taps;

  if (interior) {
    const T inv_scale = T(n_taps);
    for (int i0 = tid; i0 < n_tile; i0 += kFilterThreads * kOutPerThread) {
      T acc[kOutPerThread];
#pragma unroll
      for (int r = 0; r < kOutPerThread; ++r) acc[r] = T(0);
      // clamp the per-thread outputs of a ragged last pass onto a valid one
      int idx[kOutPerThread];
#pragma unroll
      for (int r = 0; r < kOutPerThread; ++r)
        idx[r] = min(i0 + r * kFilterThreads, n_tile - 1);
#pragma unroll 4
      for (int k = 0; k < n_taps; ++k) {
        const int w = s_taps[k];
#pragma unroll
        for (int r = 0; r < kOutPerThread; ++r) acc[r] += s_c[idx[r] - w];
      }
#pragma unroll
      for (int r = 0; r < kOutPerThread; ++r) {
        const int i = i0 + r * kFilterThreads;
        if (i < n_tile) orow[i] = s_c[i] - acc[r] / inv_scale;
      }
    }
  } else {
    for (int i = tid; i < n_tile; i += kFilterThreads) {
      const int64_t t = tile_t0 + i;
      T acc = T(0);
      int n_in = 0;
      for (int k = 0; k < n_taps; ++k) {
        const int w = s_taps[k];
        const int64_t src = t - w;
        if (src >= 0 && src < a.n_total) {
          acc += s_c[i - w];
          ++n_in;
        }
      }
      orow[i] = n_in > 0 ? s_c[i] - acc / T(n_in) : T(0);
    }
  }
}

// ---- global-memory gather (spans or tap lists too large for shared memory) --------
template <typename T>
__global__ void __launch_bounds__(kFilterThreads)
filter_gather_global_kernel(const FilterArgs<T> a) {
  const int64_t chan = blockIdx.y;
  const T* xrow = a.x + chan * a.ld_x - a.x_t0;
  T* orow = a.out + chan * a.ld_out - a.t0;
  for (int64_t t = a.t0 + int64_t(blockIdx.x) * kFilterThreads + threadIdx.x; t < a.t0 + a.n_out;
       t += int64_t(gridDim.x) * kFilterThreads) {
    T acc = T(0);
    int n_in = 0;
    for (int k = 0; k < a.n_taps; ++k) {
      const int64_t src = t - a.taps[k];
      if (src >= 0 && src < a.n_total) {
        acc += xrow[src];
        ++n_in;
      }
    }
    orow[t] = n_in > 0 ? xrow[t] - acc / T(n_in) : T(0);
  }
}

constexpr int kSmemBudget = 200 * 1024;

template <typename T>
int launch_filter(const FilterArgs<T>& args_in, int64_t n_chans, cudaStream_t stream) {
  FilterArgs<T> a = args_in;
  const int64_t span = int64_t(a.w_hi) - a.w_lo;
  const int64_t fixed = 16 + round16(a.n_taps * 4) + 32;
  const int64_t min_window = (1024 + span) * int64_t(sizeof(T));
  if (fixed + min_window > kSmemBudget) {
    const int64_t blocks = min64(ceil_div(a.n_out, kFilterThreads), 148 * 32);
    dim3 grid((unsigned)blocks, (unsigned)n_chans);
    filter_gather_global_kernel<T><<<grid, kFilterThreads, 0, stream>>>(a);
    PARRM_LAUNCH_OK("filter_gather_global_kernel");
    return PARRM_OK;
  }
  // tile: at least the halo span (<= 2x read amplification from L2), in 1024-output passes
  int64_t tile = ((span + 1023) / 1024) * 1024;
  tile = max64(4096, min64(tile, 8192));
  while (fixed + (tile + span) * int64_t(sizeof(T)) > kSmemBudget) tile -= 1024;
  tile = min64(tile, ((a.n_out + 1023) / 1024) * 1024);
  a.tile = int32_t(tile);
  const size_t smem = size_t(fixed + (tile + span + 16 / sizeof(T)) * sizeof(T));
  PARRM_CUDA_OK(cudaFuncSetAttribute(filter_gather_smem_kernel<T>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dim3 grid((unsigned)ceil_div(a.n_out, tile), (unsigned)n_chans);
  filter_gather_smem_kernel<T><<<grid, kFilterThreads, smem, stream>>>(a);
  PARRM_LAUNCH_OK("filter_gather_smem_kernel");
  return PARRM_OK;
}

}  // namespace parrm

extern "C" {

size_t parrm_filter_plan_bytes(int32_t n_taps) {
  return sizeof(parrm::FilterPlanHeader) + size_t(n_taps > 0 ? n_taps : 0) * sizeof(int32_t);
}

int parrm_filter_plan(const int32_t* h_taps, int32_t n_taps, int dtype, void* h_plan,
                      size_t plan_bytes) {
  using parrm::FilterPlanHeader;
  PARRM_REQUIRE(n_taps > 0 && h_taps != nullptr, "parrm_filter_plan: empty tap list");
  PARRM_REQUIRE(h_plan != nullptr && plan_bytes >= parrm_filter_plan_bytes(n_taps),
                "parrm_filter_plan: plan buffer too small");
  PARRM_REQUIRE(dtype == PARRM_F64 || dtype == PARRM_F32, "parrm_filter_plan: bad dtype");
  for (int32_t i = 1; i < n_taps; ++i)
    PARRM_REQUIRE(h_taps[i] > h_taps[i - 1], "parrm_filter_plan: taps must be strictly ascending");
  FilterPlanHeader* hdr = static_cast<FilterPlanHeader*>(h_plan);
  hdr->magic = parrm::kPlanMagic;
  hdr->version = 1;
  hdr->n_taps = n_taps;
  hdr->w_min = h_taps[0];
  hdr->w_max = h_taps[n_taps - 1];
  hdr->kind = 0;
  hdr->taps_offset = int32_t(sizeof(FilterPlanHeader));
  hdr->dtype = dtype;
  int32_t* taps = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(h_plan) + hdr->taps_offset);
  for (int32_t i = 0; i < n_taps; ++i) taps[i] = h_taps[i];
  return PARRM_OK;
}

int parrm_filter_apply(const void* d_x, int64_t ld_x, int64_t x_t0, int64_t n_x, void* d_out,
                       int64_t ld_out, int64_t t0, int64_t n_out, int64_t n_samples_total,
                       int64_t n_chans, const void* d_plan, const void* h_plan, int dtype,
                       void* stream) {
  using parrm::FilterPlanHeader;
  PARRM_REQUIRE(d_plan != nullptr && h_plan != nullptr, "parrm_filter_apply: null plan");
  const FilterPlanHeader* hdr = static_cast<const FilterPlanHeader*>(h_plan);
  PARRM_REQUIRE(hdr->magic == parrm::kPlanMagic && hdr->version == 1,
                "parrm_filter_apply: not a filter plan");
  PARRM_REQUIRE(hdr->dtype == dtype, "parrm_filter_apply: plan built for another dtype");
  PARRM_REQUIRE(n_chans >= 0 && n_out >= 0 && n_x >= 0 && n_samples_total >= 0,
                "parrm_filter_apply: negative size");
  PARRM_REQUIRE(n_chans <= 65535, "parrm_filter_apply: more than 65535 channels per call");
  if (n_chans == 0 || n_out == 0) return PARRM_OK;
  PARRM_REQUIRE(d_x != nullptr && d_out != nullptr, "parrm_filter_apply: null data pointer");
  PARRM_REQUIRE(t0 >= 0 && t0 + n_out <= n_samples_total,
                "parrm_filter_apply: output range outside the recording");
  const int32_t w_lo = hdr->w_min < 0 ? hdr->w_min : 0;
  const int32_t w_hi = hdr->w_max > 0 ? hdr->w_max : 0;
  {
    const int64_t need_lo = t0 - w_hi > 0 ? t0 - w_hi : 0;
    const int64_t need_hi =
        t0 + n_out - w_lo < n_samples_total ? t0 + n_out - w_lo : n_samples_total;
    PARRM_REQUIRE(x_t0 <= need_lo && x_t0 + n_x >= need_hi,
                  "parrm_filter_apply: input chunk [%lld, %lld) does not cover the halo [%lld, %lld)",
                  (long long)x_t0, (long long)(x_t0 + n_x), (long long)need_lo, (long long)need_hi);
  }
  const int32_t* d_taps = reinterpret_cast<const int32_t*>(
      static_cast<const unsigned char*>(d_plan) + hdr->taps_offset);
  cudaStream_t s = parrm::as_stream(stream);
  if (dtype == PARRM_F64) {
    parrm::FilterArgs<double> a{static_cast<const double*>(d_x), static_cast<double*>(d_out),
                                d_taps, ld_x, x_t0, n_x, ld_out, t0, n_out, n_samples_total,
                                hdr->n_taps, w_lo, w_hi, 0};
    return parrm::launch_filter<double>(a, n_chans, s);
  }
  if (dtype == PARRM_F32) {
    parrm::FilterArgs<float> a{static_cast<const float*>(d_x), static_cast<float*>(d_out),
                               d_taps, ld_x, x_t0, n_x, ld_out, t0, n_out, n_samples_total,
                               hdr->n_taps, w_lo, w_hi, 0};
    return parrm::launch_filter<float>(a, n_chans, s);
  }
  parrm::set_error("parrm_filter_apply: bad dtype %d", dtype);
  return PARRM_ERR_INVALID_ARGUMENT;
}

}  // extern "C"
