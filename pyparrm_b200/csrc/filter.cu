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
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "filter_jit.h"
#include "filter_plan.h"

namespace parrm {

// name of the kernel the last parrm_filter_apply* call of this thread enqueued
static thread_local const char* g_last_kernel = "";

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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
    g_last_kernel = "filter_gather_global_kernel";
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
  g_last_kernel = "filter_gather_smem_kernel";
  return PARRM_OK;
}


// ====================================================================================
// Comb-box strip kernel (plan kind kPlanComb, filter_plan.h)
//
// A persistent CTA walks a strip of consecutive output chunks of one channel.  Shared memory
// holds three rings, all in units of `tile`-sample chunks:
//     X    the signal window  [cur - w_hi - d, cur + tile - w_lo)  plus `prefetch` chunks that
//          the TMA engine is filling for the coming steps (cp.async.bulk -> mbarrier);
//     D_k  comb boxes  D_k[i] = sum_{q < m_k} x[i - q d]  for the offsets the gather reads.
// Each ring carries one extra "mirror" chunk after its end that duplicates ring chunk 0, so a
// run of `tile` consecutive ring elements starting anywhere is contiguous: the gather
// addresses are  (uniform per-term base) + (output index), with no per-lane wrap.
// Per step:  issue the TMA for a future chunk -> slide every D_k by one chunk
// (D[i] = D[i - d] + x[i] - x[i - m d]; every x sample is read from HBM exactly once per
// strip) -> gather  y = x - (sum of terms)/n_in  -> coalesced store.
// ====================================================================================
constexpr int kMaxPrefetch = 8;
constexpr int kMaxStages = 4;

template <typename T>
struct StripArgs {
  const T* x;
  T* out;
  const int32_t* taps;  // device, ascending (edge counts only)
  int64_t ld_x, x_t0, n_x;
  int64_t ld_out, t0, n_out;
  int64_t n_total;
  int64_t total_steps;       // n_chans * steps_per_chan
  int32_t steps_per_chan;
  int32_t n_taps, w_lo, w_hi;
  int32_t tile;              // outputs per step, multiple of blockDim * RU and of 16 / sizeof(T)
  int32_t d, nk;
  int32_t m[kMaxBoxKinds], n_box[kMaxBoxKinds], a_lo[kMaxBoxKinds];
  int32_t n_plus, n_minus, centre;
  int32_t h_back, h_fwd, prefetch;
  int32_t nq_x;                        // X ring chunks (mirror excluded)
  int32_t nq_d[kMaxBoxKinds];          // D ring chunks (mirror excluded)
  int32_t cx1[kMaxBoxKinds];           // (-a_lo_k)           mod ring size of X
  int32_t cx2[kMaxBoxKinds];           // (-(a_lo_k + m_k d)) mod ring size of X
  int32_t cprev[kMaxBoxKinds];         // (-d)                mod ring size of D_k
  int32_t seg_len[kMaxBoxKinds];       // D-pass: chain elements per work item
  int32_t n_seg[kMaxBoxKinds];
  int32_t reinit_every;                // steps between direct re-evaluations of the D rings
  int32_t chains, chain_mode, chain_lanes;  // min(d, tile); 1 = one thread per chain
  int32_t q_full, q_rem;               // tile = q_full * d + q_rem
  int32_t tab_bytes;                   // gather address table (after the barriers)
  int32_t tab_d0, tab_d1, tab_x;       // int offsets of the three sub-tables
  int32_t tab_stride_d0, tab_stride_d1, tab_stride_x;
  int32_t small_plan;                  // every table row fits the preloaded registers
  int32_t piece_steps;                 // pipelined kernel: chunks per piece (direct re-evaluation)
  int32_t stages;                      // pipelined kernel: hand-over stages (slide runs stages-1 ahead)
  int32_t off[kMaxTerms];              // per-term ring offsets: (-(a - a_lo_k)) mod |D_k| for
                                       // boxes, (-w) mod |X| for single taps
};

__device__ __forceinline__ int wrap_up(int v, int ring) { return v >= ring ? v - ring : v; }
__device__ __forceinline__ int wrap_both(int v, int ring) {
  return v < 0 ? v + ring : (v >= ring ? v - ring : v);
}
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b) {
  int64_t q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// number of taps w with 0 <= t - w < n_total
__device__ __forceinline__ int taps_in_range(const int32_t* __restrict__ taps, int n_taps, int64_t t,
                                             int64_t n_total) {
  auto upper = [&](int64_t v) {  // #taps <= v
    int lo = 0, hi = n_taps;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (int64_t(taps[mid]) <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  return upper(t) - upper(t - n_total);
}

#ifdef PARRM_STRIP_TIMING
// Debug build only: cycles CTA 0 / thread 0 spends in each phase of a step.
__device__ unsigned long long g_strip_timing[8];
#define STRIP_TICK(slot)                                             \
  do {                                                               \
    if (blockIdx.x == 0 && tid == 0) {                               \
      const long long now__ = clock64();                             \
      atomicAdd(&g_strip_timing[slot], (unsigned long long)(now__ - tick__)); \
      tick__ = now__;                                                \
    }                                                                \
  } while (0)
#define PIPE_TICK(who, slot)                                         \
  do {                                                               \
    if (blockIdx.x == 0 && tid == (who)) {                           \
      const long long now__ = clock64();                             \
      atomicAdd(&g_strip_timing[slot], (unsigned long long)(now__ - ptick__)); \
      ptick__ = now__;                                               \
    }                                                                \
  } while (0)
#else
#define STRIP_TICK(slot) do {} while (0)
#define PIPE_TICK(who, slot) do {} while (0)
#endif

// One chain of a comb box: D[e] = D[e - d] + x[e] - x[e - m d] for the chain's n elements, the
// loads batched four deep in front of the dependent adds.  MIRROR stores ring chunk 0 twice.
template <typename T, bool MIRROR>
__device__ __forceinline__ void slide_chain(const T* __restrict__ p1, const T* __restrict__ p2,
                                            T* __restrict__ pd, int mirror, int d, int n, T sum) {
  int q = 0;
#pragma unroll 1
  for (; q + 4 <= n; q += 4) {
    const T d0 = p1[0] - p2[0];
    const T d1 = p1[d] - p2[d];
    const T d2 = p1[2 * d] - p2[2 * d];
    const T d3 = p1[3 * d] - p2[3 * d];
    sum += d0; pd[0] = sum;     if (MIRROR) pd[mirror] = sum;
    sum += d1; pd[d] = sum;     if (MIRROR) pd[mirror + d] = sum;
    sum += d2; pd[2 * d] = sum; if (MIRROR) pd[mirror + 2 * d] = sum;
    sum += d3; pd[3 * d] = sum; if (MIRROR) pd[mirror + 3 * d] = sum;
    p1 += 4 * d; p2 += 4 * d; pd += 4 * d;
  }
#pragma unroll 1
  for (; q < n; ++q) {
    sum += p1[0] - p2[0];
    pd[0] = sum;
    if (MIRROR) pd[mirror] = sum;
    p1 += d; p2 += d; pd += d;
  }
}

// The same chain for an X ring without a mirror chunk: positions i1 / i2 (ring elements) wrap.
template <typename T, bool MIRROR>
__device__ __forceinline__ void slide_chain_ring(const T* __restrict__ sX, int RX, int i1, int i2,
                                                 T* __restrict__ pd, int mirror, int d, int n,
                                                 T sum) {
  auto at = [&](int pos) { return sX[pos >= RX ? pos - RX : pos]; };
  int q = 0;
#pragma unroll 1
  for (; q + 4 <= n; q += 4) {
    const T d0 = at(i1) - at(i2);
    const T d1 = at(i1 + d) - at(i2 + d);
    const T d2 = at(i1 + 2 * d) - at(i2 + 2 * d);
    const T d3 = at(i1 + 3 * d) - at(i2 + 3 * d);
    sum += d0; pd[0] = sum;     if (MIRROR) pd[mirror] = sum;
    sum += d1; pd[d] = sum;     if (MIRROR) pd[mirror + d] = sum;
    sum += d2; pd[2 * d] = sum; if (MIRROR) pd[mirror + 2 * d] = sum;
    sum += d3; pd[3 * d] = sum; if (MIRROR) pd[mirror + 3 * d] = sum;
    i1 += 4 * d; i2 += 4 * d; pd += 4 * d;
  }
#pragma unroll 1
  for (; q < n; ++q) {
    sum += at(i1) - at(i2);
    pd[0] = sum;
    if (MIRROR) pd[mirror] = sum;
    i1 += d; i2 += d; pd += d;
  }
}

template <typename T, int NT, int RU>
__global__ void __launch_bounds__(NT) filter_comb_strip_kernel(const StripArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [kMaxPrefetch]
  int32_t* const tab = reinterpret_cast<int32_t*>(smem_raw + 128);
  constexpr int VEC = 16 / sizeof(T);
  constexpr int ES = int(sizeof(T));
  const int tile = a.tile;
  const int RX = a.nq_x * tile;
  T* const sX = reinterpret_cast<T*>(smem_raw + 128 + a.tab_bytes);
  T* const sD0 = sX + RX + tile;
  const int RD0 = a.nq_d[0] * tile;
  T* const sD1 = sD0 + RD0 + tile;
  const int RD1 = a.nk > 1 ? a.nq_d[1] * tile : 0;
  const int tid = threadIdx.x;
  const int P = a.prefetch;
  const int Sc = a.steps_per_chan;
  const int d = a.d;
  const int64_t lo_valid = max64(0, a.x_t0);
  const int64_t hi_valid = min64(a.n_total, a.x_t0 + a.n_x);
  const T inv_n = T(1) / T(a.n_taps);
  const T centre = T(a.centre);
  const int n_box0 = a.n_box[0], n_box1 = a.nk > 1 ? a.n_box[1] : 0;
  const int n_plus = a.n_plus, n_minus = a.n_minus;

  if (tid == 0) {
    for (int b = 0; b < kMaxPrefetch; ++b) mbar_init(&bars[b], 1);
    fence_mbar_init();
  }
  // Gather address table: for every ring chunk slot and term, the byte offset (from smem_raw)
  // of the first element the term reads when the current chunk sits in that slot.
  {
    const int x_base = 128 + a.tab_bytes;
    const int d0_base = x_base + (RX + tile) * ES;
    const int d1_base = d0_base + (RD0 + tile) * ES;
    const int n_x_terms = 1 + a.n_plus + a.n_minus;
    for (int i = tid; i < a.nq_d[0] * a.n_box[0]; i += NT) {
      const int slot = i / a.n_box[0], t = i - slot * a.n_box[0];
      tab[a.tab_d0 + slot * a.tab_stride_d0 + t] =
          d0_base + wrap_up(slot * tile + a.off[t], RD0) * ES;
    }
    if (a.nk > 1)
      for (int i = tid; i < a.nq_d[1] * a.n_box[1]; i += NT) {
        const int slot = i / a.n_box[1], t = i - slot * a.n_box[1];
        tab[a.tab_d1 + slot * a.tab_stride_d1 + t] =
            d1_base + wrap_up(slot * tile + a.off[a.n_box[0] + t], RD1) * ES;
      }
    for (int i = tid; i < a.nq_x * n_x_terms; i += NT) {
      const int slot = i / n_x_terms, t = i - slot * n_x_terms;
      const int off = t == 0 ? 0 : a.off[a.n_box[0] + a.n_box[1] + t - 1];  // entry 0: centre
      tab[a.tab_x + slot * a.tab_stride_x + t] = x_base + wrap_up(slot * tile + off, RX) * ES;
    }
  }
  uint32_t phase_bits = 0;  // parity to wait for on each barrier (identical in every thread)
  uint32_t tma_bits = 0;    // whether the chunk in flight on a barrier went through TMA

  const int64_t F_begin = a.total_steps * int64_t(blockIdx.x) / int64_t(gridDim.x);
  const int64_t F_end = a.total_steps * int64_t(blockIdx.x + 1) / int64_t(gridDim.x);

  for (int64_t F = F_begin; F < F_end;) {
    const int64_t chan = F / Sc;
    const int s0 = int(F - chan * Sc);
    const int s1 = int(min64(Sc, s0 + (F_end - F)));
    F += s1 - s0;

    const T* xrow = a.x + chan * a.ld_x - a.x_t0;  // xrow[g] = sample at global time g
    T* orow = a.out + chan * a.ld_out - a.t0;      // orow[g]
    const int gamma =
        int((VEC - int((reinterpret_cast<uintptr_t>(xrow) / sizeof(T)) % VEC)) % VEC);
    const int64_t j_first = floor_div(a.t0 - gamma, tile);
    const int64_t j_last = floor_div(a.t0 + a.n_out - 1 - gamma, tile);  // last chunk with outputs
    const int64_t js0 = j_first + s0;
    const int64_t js_end = min64(j_first + s1, j_last + 1);
    if (js0 >= js_end) continue;
    const int n_steps = int(js_end - js0);
    // Chunks are addressed by r = j - js0 + h_back >= 0 (ring chunk slot r mod nq_x, barrier
    // r mod P); the last one any step of this piece reads is r_need_max.
    const int r_need_max = n_steps - 1 + a.h_back + a.h_fwd;
    const int64_t g_ring0 = gamma + (js0 - a.h_back) * tile;  // global time of chunk r = 0

    auto load_chunk_sync = [&](int r, int slot) {
      const int64_t g0 = g_ring0 + int64_t(r) * tile;
      T* dst = sX + slot * tile;
      for (int e = tid; e < tile; e += NT) {
        const int64_t g = g0 + e;
        const T v = (g >= lo_valid && g < hi_valid) ? xrow[g] : T(0);
        dst[e] = v;
        if (slot == 0) sX[RX + e] = v;
      }
    };
    auto issue_chunk = [&](int r, int slot, int b) {  // asynchronous inside the recording
      const int64_t g0 = g_ring0 + int64_t(r) * tile;
      if (g0 >= lo_valid && g0 + tile <= hi_valid) {
        tma_bits |= 1u << b;
        if (tid == NT - 32) {  // the last warp has no chain work: keep the issue off that path
          const uint32_t bytes = uint32_t(tile) * sizeof(T);
          fence_proxy_async();
          mbar_expect_tx(&bars[b], slot == 0 ? 2 * bytes : bytes);
          bulk_g2s(sX + slot * tile, xrow + g0, bytes, &bars[b]);
          if (slot == 0) bulk_g2s(sX + RX, xrow + g0, bytes, &bars[b]);
        }
      } else {
        tma_bits &= ~(1u << b);
        load_chunk_sync(r, slot);
      }
    };
    // Direct evaluation of the D_k chunks behind the one the coming step slides into;
    // slot_x / slot_d are the ring chunk slots of the current output chunk / new D chunk.
    auto init_boxes = [&](int slot_x, int slot_d0, int slot_d1) {
      const int sxn = slot_x * tile;
#pragma unroll
      for (int k = 0; k < kMaxBoxKinds; ++k) {
        if (k >= a.nk) break;
        T* const sD = k == 0 ? sD0 : sD1;
        const int RD = k == 0 ? RD0 : RD1;
        const int nq = a.nq_d[k], back = nq - 1, n_back = back * tile, m = a.m[k];
        // chunk slot of the oldest D chunk kept: new slot - back (mod nq) = new slot + 1
        const int slot_first = wrap_up((k == 0 ? slot_d0 : slot_d1) + 1, nq);
        for (int e = tid; e < n_back; e += NT) {
          const int rel_i = -a.a_lo[k] - n_back + e;  // relative to cur
          T sum = T(0);
          for (int q = 0; q < m; ++q) {
            const int rel = rel_i - q * d;
            if (rel < -a.h_back * tile) break;
            sum += sX[wrap_both(sxn + rel, RX)];
          }
          const int chunk = e / tile, within = e - chunk * tile;
          const int slot = wrap_up(slot_first + chunk, nq);
          sD[slot * tile + within] = sum;
          if (slot == 0) sD[RD + within] = sum;
        }
      }
    };

    __syncthreads();  // the previous piece is done with the rings; barriers and table ready
    {
      int slot = 0;
      for (int r = 0; r <= a.h_back + a.h_fwd; ++r) {
        load_chunk_sync(r, slot);
        slot = wrap_up(slot + 1, a.nq_x);
      }
    }
    // incremental ring state
    int slot_x = a.h_back % a.nq_x;                          // chunk js
    int r_issue = a.h_back + a.h_fwd + 1;                    // next chunk to issue
    int slot_issue = r_issue % a.nq_x, bar_issue = r_issue % P;
    int bar_wait = bar_issue;                                // chunk js + h_fwd + 1
    for (int i = 0; i < P - 1; ++i) {
      if (r_issue <= r_need_max) issue_chunk(r_issue, slot_issue, bar_issue);
      ++r_issue;
      slot_issue = wrap_up(slot_issue + 1, a.nq_x);
      bar_issue = wrap_up(bar_issue + 1, P);
    }
    int slot_d0 = a.nq_d[0] - 1, slot_d1 = a.nk > 1 ? a.nq_d[1] - 1 : 0;
    int reinit_in = a.reinit_every;
    __syncthreads();
    init_boxes(slot_x, slot_d0, slot_d1);
    __syncthreads();

    int64_t cur = gamma + js0 * tile;
#ifdef PARRM_STRIP_TIMING
    long long tick__ = clock64();
#endif
    for (int n = 0; n < n_steps; ++n, cur += tile) {
      if (r_issue <= r_need_max) issue_chunk(r_issue, slot_issue, bar_issue);
      STRIP_TICK(0);
      ++r_issue;
      slot_issue = wrap_up(slot_issue + 1, a.nq_x);
      bar_issue = wrap_up(bar_issue + 1, P);
      if (a.reinit_every > 0 && --reinit_in == 0) {
        reinit_in = a.reinit_every;
        init_boxes(slot_x, slot_d0, slot_d1);
        __syncthreads();
      }
      const int sxn = slot_x * tile;

      // ---- slide the comb boxes into chunk js ----
      if (a.chain_mode) {
        // one thread per chain; the two box lengths run side by side on different warps
        const int lanes = a.chain_lanes;  // chains rounded up to a whole number of warps
        for (int u = tid; u < lanes * a.nk; u += NT) {
          const int k = u >= lanes ? 1 : 0;
          const int c = u - (k ? lanes : 0);
          if (c >= a.chains) continue;
          T* const sD = k ? sD1 : sD0;
          const int RD = k ? RD1 : RD0;
          const int slot = k ? slot_d1 : slot_d0;
          const T* p1 = sX + wrap_up(sxn + a.cx1[k], RX) + c;
          const T* p2 = sX + wrap_up(sxn + a.cx2[k], RX) + c;
          T* pd = sD + slot * tile + c;
          const T sum = sD[wrap_up(slot * tile + a.cprev[k], RD) + c];
          const int n_el = a.q_full + (c < a.q_rem ? 1 : 0);
          if (slot == 0) slide_chain<T, true>(p1, p2, pd, RD, d, n_el, sum);
          else slide_chain<T, false>(p1, p2, pd, 0, d, n_el, sum);
        }
      } else {
        // few chains (small stride): chains are cut into segments; later segments start from
        // a directly summed D[i - d]
        const int chains = a.chains;
#pragma unroll
        for (int k = 0; k < kMaxBoxKinds; ++k) {
          if (k >= a.nk) break;
          T* const sD = k == 0 ? sD0 : sD1;
          const int RD = k == 0 ? RD0 : RD1;
          const int slot = k == 0 ? slot_d0 : slot_d1;
          const int L = a.seg_len[k];
          const int base1 = wrap_up(sxn + a.cx1[k], RX);
          const T* const x1 = sX + base1;
          const T* const x2 = sX + wrap_up(sxn + a.cx2[k], RX);
          T* const dnew = sD + slot * tile;
          for (int u = tid; u < chains * a.n_seg[k]; u += NT) {
            const int s = u / chains, c = u - s * chains;
            int e = c + s * L * d;
            if (e >= tile) continue;
            T sum;
            if (s == 0) {
              sum = sD[wrap_up(slot * tile + a.cprev[k], RD) + c];
            } else {
              sum = T(0);
              for (int q = 1; q <= a.m[k]; ++q) sum += sX[wrap_both(base1 + e - q * d, RX)];
            }
            const int n_el = (min(tile, e + L * d) - e + d - 1) / d;
            if (slot == 0) slide_chain<T, true>(x1 + e, x2 + e, dnew + e, RD, d, n_el, sum);
            else slide_chain<T, false>(x1 + e, x2 + e, dnew + e, 0, d, n_el, sum);
          }
        }
      }
      STRIP_TICK(1);
      __syncthreads();
      STRIP_TICK(2);

      // ---- gather ----
      const bool interior = (cur - a.w_hi >= 0) && (cur + tile - a.w_lo <= a.n_total);
      const bool all_out = (cur >= a.t0) && (cur + tile <= a.t0 + a.n_out);
      const int32_t* const row0 = tab + a.tab_d0 + slot_d0 * a.tab_stride_d0;
      const int32_t* const row1 = tab + a.tab_d1 + slot_d1 * a.tab_stride_d1;
      const int32_t* const rowx = tab + a.tab_x + slot_x * a.tab_stride_x;
      for (int i0 = tid; i0 < tile; i0 += NT * RU) {
        const unsigned char* const lane = smem_raw + i0 * ES;
        T acc[RU];
#pragma unroll
        for (int r = 0; r < RU; ++r) acc[r] = T(0);
        auto add_term = [&](int off) {
          const T* p = reinterpret_cast<const T*>(lane + off);
#pragma unroll
          for (int r = 0; r < RU; ++r) acc[r] += p[r * NT];
        };
        int x_centre;
        if (a.small_plan) {
          // <= 8 + 4 boxes and <= 3 single taps: fetch every table entry first so that all the
          // data loads of the pass are issued back to back
          const int4 oa = *reinterpret_cast<const int4*>(row0);
          const int4 ob = *reinterpret_cast<const int4*>(row0 + 4);
          const int4 oc = *reinterpret_cast<const int4*>(row1);
          const int4 ox = *reinterpret_cast<const int4*>(rowx);
          x_centre = ox.x;
          if (n_box0 > 0) add_term(oa.x);
          if (n_box0 > 1) add_term(oa.y);
          if (n_box0 > 2) add_term(oa.z);
          if (n_box0 > 3) add_term(oa.w);
          if (n_box0 > 4) add_term(ob.x);
          if (n_box0 > 5) add_term(ob.y);
          if (n_box0 > 6) add_term(ob.z);
          if (n_box0 > 7) add_term(ob.w);
          if (n_box1 > 0) add_term(oc.x);
          if (n_box1 > 1) add_term(oc.y);
          if (n_box1 > 2) add_term(oc.z);
          if (n_box1 > 3) add_term(oc.w);
          const int xo[3] = {ox.y, ox.z, ox.w};
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            if (t < n_plus) {
              add_term(xo[t]);
            } else if (t < n_plus + n_minus) {
              const T* p = reinterpret_cast<const T*>(lane + xo[t]);
#pragma unroll
              for (int r = 0; r < RU; ++r) acc[r] -= p[r * NT];
            }
          }
        } else {
          // (runtime trip counts: keep ptxas from unrolling these further, the remainder
          // scaffolding would cost more than the loops)
          x_centre = rowx[0];
          int t = 0;
#pragma unroll 1
          for (; t + 4 <= n_box0; t += 4) {
            const int4 o = *reinterpret_cast<const int4*>(row0 + t);
            add_term(o.x); add_term(o.y); add_term(o.z); add_term(o.w);
          }
#pragma unroll 1
          for (; t < n_box0; ++t) add_term(row0[t]);
          t = 0;
#pragma unroll 1
          for (; t + 4 <= n_box1; t += 4) {
            const int4 o = *reinterpret_cast<const int4*>(row1 + t);
            add_term(o.x); add_term(o.y); add_term(o.z); add_term(o.w);
          }
#pragma unroll 1
          for (; t < n_box1; ++t) add_term(row1[t]);
#pragma unroll 1
          for (t = 1; t <= n_plus; ++t) add_term(rowx[t]);
#pragma unroll 1
          for (t = 1 + n_plus; t <= n_plus + n_minus; ++t) {
            const T* p = reinterpret_cast<const T*>(lane + rowx[t]);
#pragma unroll
            for (int r = 0; r < RU; ++r) acc[r] -= p[r * NT];
          }
        }
        const T* const xc = reinterpret_cast<const T*>(lane + x_centre);
        T* const og = orow + cur + i0;
        if (interior && all_out) {
#pragma unroll
          for (int r = 0; r < RU; ++r) {
            const T x0 = xc[r * NT];
            og[r * NT] = x0 - (acc[r] + centre * x0) * inv_n;
          }
        } else {
#pragma unroll
          for (int r = 0; r < RU; ++r) {
            const int64_t g = cur + i0 + r * NT;
            const T x0 = xc[r * NT];
            const T sum = acc[r] + centre * x0;
            const int n_in = interior ? a.n_taps : taps_in_range(a.taps, a.n_taps, g, a.n_total);
            const T y = n_in > 0 ? x0 - sum / T(n_in) : T(0);
            if (g >= a.t0 && g < a.t0 + a.n_out) og[r * NT] = y;
          }
        }
      }
      STRIP_TICK(3);
      // chunk js + h_fwd + 1 must have landed before the next step
      if (tma_bits & (1u << bar_wait)) {
        mbar_wait(&bars[bar_wait], (phase_bits >> bar_wait) & 1u);
        phase_bits ^= 1u << bar_wait;
        tma_bits &= ~(1u << bar_wait);
      }
      bar_wait = wrap_up(bar_wait + 1, P);
      slot_x = wrap_up(slot_x + 1, a.nq_x);
      slot_d0 = wrap_up(slot_d0 + 1, a.nq_d[0]);
      if (a.nk > 1) slot_d1 = wrap_up(slot_d1 + 1, a.nq_d[1]);
      STRIP_TICK(4);
      __syncthreads();
      STRIP_TICK(5);
    }
  }
}

// ====================================================================================
// Pipelined strip kernel: the same rings and arithmetic as filter_comb_strip_kernel, but the
// box slide and the gather run on different warps and overlap.  ND "slide" threads run one
// chunk ahead: they issue the TMA copies, wait for them, and slide the boxes into chunk n + 1
// while the NG "gather" threads produce the outputs of chunk n.  The D rings hold one chunk
// more than the gather reads (X two more), and two mbarrier pairs hand chunks over:
//     full[n & 1]   slide -> gather   chunk n of every D ring is written
//     empty[n & 1]  gather -> slide   the gather of chunk n has read everything it needs
// The latency-bound dependent adds of the slide thus hide behind the bandwidth-bound gather,
// and a step has no CTA-wide barrier at all.  A piece covers at most `piece_steps` chunks, so
// the boxes are re-evaluated directly often enough to bound rounding drift.
// ====================================================================================
template <typename T, int NG, int ND, int RU>
__global__ void __launch_bounds__(NG + ND) filter_comb_pipe_kernel(const StripArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem_raw);  // [kMaxPrefetch] TMA chunks
  uint64_t* const full = bars + kMaxPrefetch;                    // [kMaxStages]
  uint64_t* const empty = full + kMaxStages;                     // [kMaxStages]
  int32_t* const tab = reinterpret_cast<int32_t*>(smem_raw + 128);
  constexpr int VEC = 16 / sizeof(T);
  constexpr int ES = int(sizeof(T));
  constexpr int NT = NG + ND;
  const int tile = a.tile;
  const int RX = a.nq_x * tile;
  T* const sX = reinterpret_cast<T*>(smem_raw + 128 + a.tab_bytes);
  T* const sD0 = sX + RX;  // the X ring has no mirror chunk here: X reads wrap per lane
  const int RD0 = a.nq_d[0] * tile;
  T* const sD1 = sD0 + RD0 + tile;
  const int RD1 = a.nk > 1 ? a.nq_d[1] * tile : 0;
  const int tid = threadIdx.x;
  const bool is_gather = tid < NG;
  const int gt = tid;        // gather thread index (valid when is_gather)
  const int dt = tid - NG;   // slide thread index (valid otherwise)
  const int P = a.prefetch;
  const int Sc = a.steps_per_chan;
  const int d = a.d;
  const int64_t lo_valid = max64(0, a.x_t0);
  const int64_t hi_valid = min64(a.n_total, a.x_t0 + a.n_x);
  const T inv_n = T(1) / T(a.n_taps);
  const T centre = T(a.centre);
  const int n_box0 = a.n_box[0], n_box1 = a.nk > 1 ? a.n_box[1] : 0;
  const int n_plus = a.n_plus, n_minus = a.n_minus;
  const int K = a.stages;  // the slide may run K - 1 chunks ahead of the gather
  const int back0 = a.nq_d[0] - K, back1 = a.nk > 1 ? a.nq_d[1] - K : 0;

  if (tid == 0) {
    for (int b = 0; b < kMaxPrefetch; ++b) mbar_init(&bars[b], 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full[s], 1);         // one elected slide thread arrives
      mbar_init(&empty[s], NG / 32);  // one lane of every gather warp arrives
    }
    fence_mbar_init();
  }
  {  // gather address table (see filter_comb_strip_kernel); padding entries stay valid offsets
    const int x_base = 128 + a.tab_bytes;
    for (int i = tid; i < a.tab_bytes / 4; i += NT) tab[i] = x_base;
    __syncthreads();
    const int d0_base = x_base + RX * ES;
    const int d1_base = d0_base + (RD0 + tile) * ES;
    const int n_x_terms = 1 + n_plus + n_minus;
    for (int i = tid; i < a.nq_d[0] * n_box0; i += NT) {
      const int slot = i / n_box0, t = i - slot * n_box0;
      tab[a.tab_d0 + slot * a.tab_stride_d0 + t] =
          d0_base + wrap_up(slot * tile + a.off[t], RD0) * ES;
    }
    if (a.nk > 1)
      for (int i = tid; i < a.nq_d[1] * n_box1; i += NT) {
        const int slot = i / n_box1, t = i - slot * n_box1;
        tab[a.tab_d1 + slot * a.tab_stride_d1 + t] =
            d1_base + wrap_up(slot * tile + a.off[n_box0 + t], RD1) * ES;
      }
    for (int i = tid; i < a.nq_x * n_x_terms; i += NT) {
      const int slot = i / n_x_terms, t = i - slot * n_x_terms;
      const int off = t == 0 ? 0 : a.off[n_box0 + n_box1 + t - 1];
      tab[a.tab_x + slot * a.tab_stride_x + t] = x_base + wrap_up(slot * tile + off, RX) * ES;
    }
  }
  // phase parities; every thread keeps the ones its role waits on
  uint32_t tma_phase = 0, tma_bits = 0;  // slide threads
  // Chunk hand-overs are numbered by a counter g that runs on through the pieces of this CTA:
  // hand-over g uses barrier pair g % K and completes its (g / K)-th phase, so the parity a
  // waiter needs follows from g alone.  Kept incrementally: stage = g % K, phase = (g / K) & 1.
  int hand_stage = 0, hand_count = 0;
  uint32_t hand_phase = 0;

  const int64_t F_begin = a.total_steps * int64_t(blockIdx.x) / int64_t(gridDim.x);
  const int64_t F_end = a.total_steps * int64_t(blockIdx.x + 1) / int64_t(gridDim.x);

  for (int64_t F = F_begin; F < F_end;) {
    const int64_t chan = F / Sc;
    const int s0 = int(F - chan * Sc);
    const int s1 = int(min64(min64(Sc, s0 + (F_end - F)), int64_t(s0) + a.piece_steps));
    F += s1 - s0;

    const T* xrow = a.x + chan * a.ld_x - a.x_t0;
    T* orow = a.out + chan * a.ld_out - a.t0;
    const int gamma =
        int((VEC - int((reinterpret_cast<uintptr_t>(xrow) / sizeof(T)) % VEC)) % VEC);
    const int64_t j_first = floor_div(a.t0 - gamma, tile);
    const int64_t j_last = floor_div(a.t0 + a.n_out - 1 - gamma, tile);
    const int64_t js0 = j_first + s0;
    const int64_t js_end = min64(j_first + s1, j_last + 1);
    if (js0 >= js_end) continue;
    const int n_steps = int(js_end - js0);
    const int r_need_max = n_steps - 1 + a.h_back + a.h_fwd;
    const int64_t g_ring0 = gamma + (js0 - a.h_back) * tile;

    // chunk r -> ring chunk slot, cooperative copy by `nthr` threads (zero fill outside)
    auto load_chunk_sync = [&](int r, int slot, int t0, int nthr) {
      const int64_t g0 = g_ring0 + int64_t(r) * tile;
      T* dst = sX + slot * tile;
      for (int e = t0; e < tile; e += nthr) {
        const int64_t g = g0 + e;
        const T v = (g >= lo_valid && g < hi_valid) ? xrow[g] : T(0);
        dst[e] = v;
      }
    };
    // issued by the slide threads (or, in the prologue, by everybody with t0/nthr of the CTA)
    auto issue_chunk = [&](int r, int slot, int b, int t0, int nthr) {
      const int64_t g0 = g_ring0 + int64_t(r) * tile;
      if (g0 >= lo_valid && g0 + tile <= hi_valid) {
        tma_bits |= 1u << b;
        if (t0 == 0) {
          const uint32_t bytes = uint32_t(tile) * sizeof(T);
          fence_proxy_async();
          mbar_expect_tx(&bars[b], bytes);
          bulk_g2s(sX + slot * tile, xrow + g0, bytes, &bars[b]);
        }
      } else {
        tma_bits &= ~(1u << b);
        load_chunk_sync(r, slot, t0, nthr);
      }
    };

    __syncthreads();  // previous piece fully drained; barriers and table ready
    {
      int slot = 0;
      for (int r = 0; r <= a.h_back + a.h_fwd; ++r) {
        load_chunk_sync(r, slot, tid, NT);
        slot = wrap_up(slot + 1, a.nq_x);
      }
    }
    int r_issue = a.h_back + a.h_fwd + 1;
    int slot_issue = r_issue % a.nq_x, bar_issue = r_issue % P;
    int bar_wait = bar_issue;  // barrier of chunk (n + 1) + h_fwd while step n runs
    // the slide threads own the TMA bookkeeping; thread `NG` (dt == 0) issues
    for (int i = 0; i < P - 1; ++i) {
      if (r_issue <= r_need_max) {
        if (!is_gather) issue_chunk(r_issue, slot_issue, bar_issue, dt, ND);
      }
      ++r_issue;
      slot_issue = wrap_up(slot_issue + 1, a.nq_x);
      bar_issue = wrap_up(bar_issue + 1, P);
    }
    __syncthreads();
    {  // direct evaluation of the D chunks behind chunk 0 (ring chunk slots 0 .. back-1)
      const int sxn = a.h_back * tile;
#pragma unroll
      for (int k = 0; k < kMaxBoxKinds; ++k) {
        if (k >= a.nk) break;
        T* const sD = k == 0 ? sD0 : sD1;
        const int RD = k == 0 ? RD0 : RD1;
        const int back = k == 0 ? back0 : back1, n_back = back * tile, m = a.m[k];
        for (int e = tid; e < n_back; e += NT) {
          const int rel_i = -a.a_lo[k] - n_back + e;
          T sum = T(0);
          for (int q = 0; q < m; ++q) {
            const int rel = rel_i - q * d;
            if (rel < -a.h_back * tile) break;
            sum += sX[wrap_both(sxn + rel, RX)];
          }
          sD[e] = sum;
          if (e < tile) sD[RD + e] = sum;
        }
      }
    }
    __syncthreads();

    if (!is_gather) {
      // ------------------------------ slide warps ------------------------------
      int slot_x = a.h_back;  // ring chunk slot of chunk n
      int slot_d0 = back0, slot_d1 = back1;
#ifdef PARRM_STRIP_TIMING
      long long ptick__ = clock64();
#endif
      for (int n = 0; n < n_steps; ++n) {
        if (hand_count >= K)  // the gather K hand-overs back has released the slots written below
          mbar_wait(&empty[hand_stage], hand_phase ^ 1u);
        PIPE_TICK(NG, 4);
        if (r_issue <= r_need_max) issue_chunk(r_issue, slot_issue, bar_issue, dt, ND);
        ++r_issue;
        slot_issue = wrap_up(slot_issue + 1, a.nq_x);
        bar_issue = wrap_up(bar_issue + 1, P);
        const int sxn = slot_x * tile;
#ifndef PARRM_DEBUG_SKIP_SLIDE
        if (a.chain_mode) {
          const int lanes = a.chain_lanes;
          for (int u = dt; u < lanes * a.nk; u += ND) {
            const int k = u >= lanes ? 1 : 0;
            const int c = u - (k ? lanes : 0);
            if (c >= a.chains) continue;
            T* const sD = k ? sD1 : sD0;
            const int RD = k ? RD1 : RD0;
            const int slot = k ? slot_d1 : slot_d0;
            const int i1 = wrap_up(sxn + a.cx1[k], RX) + c;
            const int i2 = wrap_up(sxn + a.cx2[k], RX) + c;
            T* pd = sD + slot * tile + c;
            const T sum = sD[wrap_up(slot * tile + a.cprev[k], RD) + c];
            const int n_el = a.q_full + (c < a.q_rem ? 1 : 0);
            if (slot == 0) slide_chain_ring<T, true>(sX, RX, i1, i2, pd, RD, d, n_el, sum);
            else slide_chain_ring<T, false>(sX, RX, i1, i2, pd, 0, d, n_el, sum);
          }
        } else {
          const int chains = a.chains;
#pragma unroll
          for (int k = 0; k < kMaxBoxKinds; ++k) {
            if (k >= a.nk) break;
            T* const sD = k == 0 ? sD0 : sD1;
            const int RD = k == 0 ? RD0 : RD1;
            const int slot = k == 0 ? slot_d0 : slot_d1;
            const int L = a.seg_len[k];
            const int base1 = wrap_up(sxn + a.cx1[k], RX);
            const int base2 = wrap_up(sxn + a.cx2[k], RX);
            T* const dnew = sD + slot * tile;
            for (int u = dt; u < chains * a.n_seg[k]; u += ND) {
              const int s = u / chains, c = u - s * chains;
              const int e = c + s * L * d;
              if (e >= tile) continue;
              T sum;
              if (s == 0) {
                sum = sD[wrap_up(slot * tile + a.cprev[k], RD) + c];
              } else {
                sum = T(0);
                for (int q = 1; q <= a.m[k]; ++q) sum += sX[wrap_both(base1 + e - q * d, RX)];
              }
              const int n_el = (min(tile, e + L * d) - e + d - 1) / d;
              if (slot == 0)
                slide_chain_ring<T, true>(sX, RX, base1 + e, base2 + e, dnew + e, RD, d, n_el, sum);
              else
                slide_chain_ring<T, false>(sX, RX, base1 + e, base2 + e, dnew + e, 0, d, n_el, sum);
            }
          }
        }
#endif
        PIPE_TICK(NG, 5);
        // chunk (n + 1) + h_fwd must have landed before the next slide (and the next gather)
        if (tma_bits & (1u << bar_wait)) {
          mbar_wait(&bars[bar_wait], (tma_phase >> bar_wait) & 1u);
          tma_phase ^= 1u << bar_wait;
          tma_bits &= ~(1u << bar_wait);
        }
        bar_wait = wrap_up(bar_wait + 1, P);
        PIPE_TICK(NG, 6);
        named_bar_sync(1, ND);  // every slide thread is done with chunk n (and any sync loads)
        if (dt == 0) mbar_arrive(&full[hand_stage]);
        ++hand_count;
        if (++hand_stage == K) {
          hand_stage = 0;
          hand_phase ^= 1u;
        }
        PIPE_TICK(NG, 7);
        slot_x = wrap_up(slot_x + 1, a.nq_x);
        slot_d0 = wrap_up(slot_d0 + 1, a.nq_d[0]);
        if (a.nk > 1) slot_d1 = wrap_up(slot_d1 + 1, a.nq_d[1]);
      }
    } else {
      // ------------------------------ gather warps ------------------------------
      int slot_x = a.h_back, slot_d0 = back0, slot_d1 = back1;
      int64_t cur = gamma + js0 * tile;
#ifdef PARRM_STRIP_TIMING
      long long ptick__ = clock64();
#endif
      for (int n = 0; n < n_steps; ++n, cur += tile) {
        mbar_wait(&full[hand_stage], hand_phase);
        PIPE_TICK(0, 0);
        const bool interior = (cur - a.w_hi >= 0) && (cur + tile - a.w_lo <= a.n_total);
        const bool all_out = (cur >= a.t0) && (cur + tile <= a.t0 + a.n_out);
        const int32_t* const row0 = tab + a.tab_d0 + slot_d0 * a.tab_stride_d0;
        const int32_t* const row1 = tab + a.tab_d1 + slot_d1 * a.tab_stride_d1;
        const int32_t* const rowx = tab + a.tab_x + slot_x * a.tab_stride_x;
#ifndef PARRM_DEBUG_SKIP_GATHER
        for (int i0 = gt; i0 < tile; i0 += NG * RU) {
          const unsigned char* const lane = smem_raw + i0 * ES;
          T acc[RU];
#pragma unroll
          for (int r = 0; r < RU; ++r) acc[r] = T(0);
          auto add_term = [&](int off) {
            const T* p = reinterpret_cast<const T*>(lane + off);
#pragma unroll
            for (int r = 0; r < RU; ++r) acc[r] += p[r * NG];
          };
          // X-ring terms (single taps, the centre sample): no mirror chunk, wrap per lane
          const int x_end = 128 + a.tab_bytes + RX * ES, x_ring = RX * ES;
          auto load_x = [&](int off, T (&v)[RU]) {
#pragma unroll
            for (int r = 0; r < RU; ++r) {
              int pos = off + (i0 + r * NG) * ES;
              if (pos >= x_end) pos -= x_ring;
              v[r] = *reinterpret_cast<const T*>(smem_raw + pos);
            }
          };
          auto add_x = [&](int off, T sign) {
            T v[RU];
            load_x(off, v);
#pragma unroll
            for (int r = 0; r < RU; ++r) acc[r] = fma(sign, v[r], acc[r]);
          };
          int x_centre;
          if (a.small_plan) {
            const int4 oa = *reinterpret_cast<const int4*>(row0);
            const int4 ob = *reinterpret_cast<const int4*>(row0 + 4);
            const int4 oc = *reinterpret_cast<const int4*>(row1);
            const int4 ox = *reinterpret_cast<const int4*>(rowx);
            x_centre = ox.x;
            switch (n_box0) {  // one indirect branch instead of a compare per term
              case 8: add_term(ob.w);
              case 7: add_term(ob.z);
              case 6: add_term(ob.y);
              case 5: add_term(ob.x);
              case 4: add_term(oa.w);
              case 3: add_term(oa.z);
              case 2: add_term(oa.y);
              case 1: add_term(oa.x);
              default: break;
            }
            switch (n_box1) {
              case 4: add_term(oc.w);
              case 3: add_term(oc.z);
              case 2: add_term(oc.y);
              case 1: add_term(oc.x);
              default: break;
            }
            const int xo[3] = {ox.y, ox.z, ox.w};
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              if (t < n_plus) add_x(xo[t], T(1));
              else if (t < n_plus + n_minus) add_x(xo[t], T(-1));
            }
          } else {
            x_centre = rowx[0];
            int t = 0;
#pragma unroll 1
            for (; t + 4 <= n_box0; t += 4) {
              const int4 o = *reinterpret_cast<const int4*>(row0 + t);
              add_term(o.x); add_term(o.y); add_term(o.z); add_term(o.w);
            }
#pragma unroll 1
            for (; t < n_box0; ++t) add_term(row0[t]);
            t = 0;
#pragma unroll 1
            for (; t + 4 <= n_box1; t += 4) {
              const int4 o = *reinterpret_cast<const int4*>(row1 + t);
              add_term(o.x); add_term(o.y); add_term(o.z); add_term(o.w);
            }
#pragma unroll 1
            for (; t < n_box1; ++t) add_term(row1[t]);
#pragma unroll 1
            for (t = 1; t <= n_plus; ++t) add_x(rowx[t], T(1));
#pragma unroll 1
            for (t = 1 + n_plus; t <= n_plus + n_minus; ++t) add_x(rowx[t], T(-1));
          }
          T xc[RU];
          load_x(x_centre, xc);
          T* const og = orow + cur + i0;
          if (interior && all_out) {
#pragma unroll
            for (int r = 0; r < RU; ++r) {
              const T x0 = xc[r];
              og[r * NG] = x0 - (acc[r] + centre * x0) * inv_n;
            }
          } else {
#pragma unroll
            for (int r = 0; r < RU; ++r) {
              const int64_t g = cur + i0 + r * NG;
              const T x0 = xc[r];
              const T sum = acc[r] + centre * x0;
              const int n_in =
                  interior ? a.n_taps : taps_in_range(a.taps, a.n_taps, g, a.n_total);
              const T y = n_in > 0 ? x0 - sum / T(n_in) : T(0);
              if (g >= a.t0 && g < a.t0 + a.n_out) og[r * NG] = y;
            }
          }
        }
#endif
        PIPE_TICK(0, 1);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[hand_stage]);
        if (++hand_stage == K) {
          hand_stage = 0;
          hand_phase ^= 1u;
        }
        PIPE_TICK(0, 2);
        slot_x = wrap_up(slot_x + 1, a.nq_x);
        slot_d0 = wrap_up(slot_d0 + 1, a.nq_d[0]);
        if (a.nk > 1) slot_d1 = wrap_up(slot_d1 + 1, a.nq_d[1]);
      }
    }
  }
}

struct StripTuning {
  int pipe;         // 0 two-phase kernel, 1 producer/consumer (slide warps a chunk ahead)
  int threads;      // gather threads (all threads when pipe = 0)
  int slide;        // slide threads (pipe = 1)
  int ru, tile, prefetch, ctas_per_sm;
  int stages;       // pipe = 1: hand-over stages (2 or 3); the slide runs stages - 1 chunks ahead
};

template <typename T>
int launch_strip(const FilterPlanHeader* hdr, const int32_t* h_terms, const int32_t* d_taps,
                 const FilterArgs<T>& f, int64_t n_chans, cudaStream_t stream, bool* launched) {
  *launched = false;
  StripArgs<T> a;
  memset(&a, 0, sizeof(a));
  a.x = f.x; a.out = f.out; a.taps = d_taps;
  a.ld_x = f.ld_x; a.x_t0 = f.x_t0; a.n_x = f.n_x;
  a.ld_out = f.ld_out; a.t0 = f.t0; a.n_out = f.n_out; a.n_total = f.n_total;
  a.n_taps = f.n_taps; a.w_lo = f.w_lo; a.w_hi = f.w_hi;
  a.d = hdr->stride; a.nk = hdr->n_kinds;
  a.n_plus = hdr->n_plus; a.n_minus = hdr->n_minus; a.centre = hdr->centre;
  const int n_terms = hdr->n_box[0] + hdr->n_box[1] + hdr->n_plus + hdr->n_minus;
  if (n_terms > kMaxTerms || a.nk < 1 || a.nk > kMaxBoxKinds) return PARRM_OK;

  auto round4 = [](int64_t v) { return (v + 3) & ~int64_t(3); };
  auto env_int = [](const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
  };
  const int sm_budget = 227 * 1024;
  // Candidate shapes, best first as measured on cfg2 (scripts/sweep_filter.py).  The
  // pipelined kernel needs one more chunk per ring; where that does not fit, the two-phase
  // kernel with a big tile (which amortises its two barriers per step) comes next, then
  // smaller tiles for wide tap windows.
  const StripTuning shapes[] = {
      {1, 512, 512, 4, 2048, 1, 1, 2}, {1, 256, 256, 6, 1536, 2, 1, 2}, {1, 512, 256, 2, 1024, 3, 1, 2},
      {0, 512, 0, 4, 2048, 3, 1, 1},   {1, 256, 256, 2, 512, 3, 1, 2},  {0, 256, 0, 3, 768, 2, 2, 1},
      {0, 512, 0, 2, 1024, 4, 1, 1},   {0, 256, 0, 2, 512, 4, 2, 1},    {0, 256, 0, 2, 512, 3, 1, 1},
      {0, 256, 0, 1, 256, 4, 1, 1}};
  StripTuning pick{0, 0, 0, 0, 0, 0, 0, 0};
  size_t pick_smem = 0;
  const int forced_tile = env_int("PARRM_FILTER_TILE", 0);
  const int allow_pipe = env_int("PARRM_FILTER_PIPE", 1);
  for (const StripTuning& s0 : shapes) {
    StripTuning s = s0;
    if (forced_tile) {
      s.tile = forced_tile;
      s.pipe = env_int("PARRM_FILTER_PIPE", s.pipe);
      s.threads = env_int("PARRM_FILTER_THREADS", s.threads);
      s.slide = env_int("PARRM_FILTER_SLIDE", s.pipe == 1 ? 256 : 0);
      s.ru = env_int("PARRM_FILTER_RU", s.ru);
      s.prefetch = env_int("PARRM_FILTER_PREFETCH", s.prefetch);
      s.ctas_per_sm = env_int("PARRM_FILTER_CTAS", s.ctas_per_sm);
      s.stages = env_int("PARRM_FILTER_STAGES", s.stages);
    }
    if (s.pipe && (s.stages < 2 || s.stages > kMaxStages)) continue;
    if (s.pipe && !allow_pipe) continue;
    if (s.tile % (s.threads * s.ru) != 0 || s.prefetch < 1 || s.prefetch > kMaxPrefetch) continue;
    const int64_t tile = s.tile;
    const int extra = s.pipe ? s.stages : 1;  // ring chunks beyond what one step reads
    const int64_t h_back = ceil_div(int64_t(f.w_hi) + a.d, tile);
    const int64_t h_fwd = ceil_div(-int64_t(f.w_lo), tile);
    const int64_t nq_x = h_back + h_fwd + extra + s.prefetch;
    int64_t elems = (nq_x + (s.pipe ? 0 : 1)) * tile;  // the pipelined kernel's X ring has no mirror
    int64_t tab_ints = nq_x * round4(1 + hdr->n_plus + hdr->n_minus) + 8;
    for (int k = 0; k < kMaxBoxKinds; ++k) {
      const int64_t reach = max64(int64_t(hdr->a_max[k]) - hdr->a_min[k], a.d);
      const int64_t nq_d = ceil_div(reach, tile) + extra;
      if (k < a.nk) elems += (nq_d + 1) * tile;
      tab_ints += nq_d * max64(8, round4(hdr->n_box[k]));
    }
    const int64_t tab_bytes = ((tab_ints * 4 + 127) / 128) * 128;
    const size_t smem = 128 + size_t(tab_bytes) + size_t(elems) * sizeof(T);
    if (smem + 1024 > size_t(sm_budget) / s.ctas_per_sm) continue;
    pick = s;
    pick_smem = smem;
    break;
  }
  if (pick.tile == 0) return PARRM_OK;  // rings do not fit: the caller falls back to the gather

  const int64_t tile = pick.tile;
  const int extra = pick.pipe ? pick.stages : 1;
  a.stages = pick.stages;
  a.tile = pick.tile;
  a.prefetch = pick.prefetch;
  a.h_back = int32_t(ceil_div(int64_t(f.w_hi) + a.d, tile));
  a.h_fwd = int32_t(ceil_div(-int64_t(f.w_lo), tile));
  a.nq_x = a.h_back + a.h_fwd + extra + a.prefetch;
  const int64_t RX = int64_t(a.nq_x) * tile;
  auto mod = [](int64_t v, int64_t ring) { return int32_t(((v % ring) + ring) % ring); };
  const int slide_threads = pick.pipe == 1 ? pick.slide : pick.threads;
  int t = 0;
  for (int k = 0; k < a.nk; ++k) {
    a.m[k] = hdr->window[k];
    a.n_box[k] = hdr->n_box[k];
    a.a_lo[k] = hdr->a_min[k];
    const int64_t reach = max64(int64_t(hdr->a_max[k]) - hdr->a_min[k], a.d);
    a.nq_d[k] = int32_t(ceil_div(reach, tile) + extra);
    const int64_t RD = int64_t(a.nq_d[k]) * tile;
    a.cx1[k] = mod(-int64_t(a.a_lo[k]), RX);
    a.cx2[k] = mod(-(int64_t(a.a_lo[k]) + int64_t(a.m[k]) * a.d), RX);
    a.cprev[k] = mod(-int64_t(a.d), RD);
    // slide work split: one item per chain unless there are few chains
    const int64_t chains = min64(a.d, tile);
    const int64_t per_chain = ceil_div(tile, a.d);
    int64_t seg = per_chain;
    if (chains * a.nk < slide_threads / 2) {
      seg = max64(9, a.m[k]) | 1;  // odd: conflict-free shared-memory strides when d is small
      seg = min64(seg, per_chain);
    }
    a.seg_len[k] = int32_t(seg);
    a.n_seg[k] = int32_t(ceil_div(per_chain, seg));
    for (int b = 0; b < a.n_box[k]; ++b, ++t)
      a.off[t] = mod(-(int64_t(h_terms[t]) - a.a_lo[k]), RD);
  }
  for (int b = 0; b < a.n_plus + a.n_minus; ++b, ++t) a.off[t] = mod(-int64_t(h_terms[t]), RX);
  a.reinit_every = env_int("PARRM_FILTER_REINIT", 256);
  a.piece_steps = a.reinit_every > 0 ? a.reinit_every : (1 << 30);
  a.chains = int32_t(min64(a.d, tile));
  a.chain_mode = 1;
  for (int k = 0; k < a.nk; ++k)
    if (a.n_seg[k] != 1) a.chain_mode = 0;
  a.chain_lanes = (a.chains + 31) & ~31;
  a.q_full = int32_t(tile / a.d);
  a.q_rem = int32_t(tile - int64_t(a.q_full) * a.d);
  a.tab_stride_d0 = int32_t(max64(8, round4(a.n_box[0])));
  a.tab_stride_d1 = int32_t(max64(4, round4(a.n_box[1])));
  a.tab_stride_x = int32_t(round4(1 + a.n_plus + a.n_minus));
  a.tab_d0 = 0;
  a.tab_d1 = a.tab_d0 + a.nq_d[0] * a.tab_stride_d0;
  a.tab_x = a.tab_d1 + max64(1, a.nq_d[1]) * a.tab_stride_d1;
  a.small_plan = (a.n_box[0] <= 8 && a.n_box[1] <= 4 && a.n_plus + a.n_minus <= 3) ? 1 : 0;
  a.tab_bytes = int32_t(((int64_t(a.tab_x + a.nq_x * a.tab_stride_x) * 4 + 127) / 128) * 128);
  if (128 + size_t(a.tab_bytes) > pick_smem) return PARRM_OK;

  a.steps_per_chan = int32_t(ceil_div(f.n_out + tile - 1, tile));
  a.total_steps = n_chans * int64_t(a.steps_per_chan);
  // one strip per resident CTA; a strip is at least 8 steps so its prologue (halo load and
  // direct box evaluation) stays a small fraction of the work
  const int64_t resident = int64_t(kNumSMs) * pick.ctas_per_sm;
  const int64_t grid = max64(1, min64(resident, a.total_steps / 8));

  void (*kernel)(const StripArgs<T>) = nullptr;
  int block = pick.threads;
  if (pick.pipe == 1) {
    block = pick.threads + pick.slide;
#define PARRM_PIPE_SHAPE(NG_, ND_, RU_) \
  if (pick.threads == NG_ && pick.slide == ND_ && pick.ru == RU_) \
    kernel = filter_comb_pipe_kernel<T, NG_, ND_, RU_>;
    PARRM_PIPE_SHAPE(512, 256, 2) PARRM_PIPE_SHAPE(512, 256, 3) PARRM_PIPE_SHAPE(512, 256, 4)
    PARRM_PIPE_SHAPE(256, 256, 2) PARRM_PIPE_SHAPE(256, 256, 4) PARRM_PIPE_SHAPE(512, 448, 2)
    PARRM_PIPE_SHAPE(256, 128, 2) PARRM_PIPE_SHAPE(256, 128, 4) PARRM_PIPE_SHAPE(512, 128, 2)
    PARRM_PIPE_SHAPE(768, 256, 2) PARRM_PIPE_SHAPE(384, 256, 4) PARRM_PIPE_SHAPE(768, 128, 2)
    PARRM_PIPE_SHAPE(256, 256, 6) PARRM_PIPE_SHAPE(384, 128, 4) PARRM_PIPE_SHAPE(384, 384, 4)
    PARRM_PIPE_SHAPE(256, 256, 8) PARRM_PIPE_SHAPE(512, 512, 4) PARRM_PIPE_SHAPE(640, 256, 2)
    PARRM_PIPE_SHAPE(320, 320, 4) PARRM_PIPE_SHAPE(512, 512, 2) PARRM_PIPE_SHAPE(256, 256, 5)
    PARRM_PIPE_SHAPE(320, 256, 4) PARRM_PIPE_SHAPE(512, 512, 3)
#undef PARRM_PIPE_SHAPE
  } else {
#define PARRM_STRIP_SHAPE(NT_, RU_) \
  if (pick.threads == NT_ && pick.ru == RU_) kernel = filter_comb_strip_kernel<T, NT_, RU_>;
    PARRM_STRIP_SHAPE(256, 1) PARRM_STRIP_SHAPE(256, 2) PARRM_STRIP_SHAPE(256, 3)
    PARRM_STRIP_SHAPE(256, 4) PARRM_STRIP_SHAPE(256, 8) PARRM_STRIP_SHAPE(512, 2)
    PARRM_STRIP_SHAPE(512, 4) PARRM_STRIP_SHAPE(1024, 2)
#undef PARRM_STRIP_SHAPE
  }
  if (kernel == nullptr) return PARRM_OK;  // unknown shape: plain gather
  PARRM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(pick_smem)));
  PARRM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
  kernel<<<unsigned(grid), block, pick_smem, stream>>>(a);
  PARRM_LAUNCH_OK("filter_comb_strip_kernel");
  g_last_kernel = pick.pipe == 1 ? "filter_comb_pipe_kernel" : "filter_comb_strip_kernel";
  *launched = true;
  return PARRM_OK;
}

}  // namespace parrm

extern "C" {

#ifdef PARRM_STRIP_TIMING
int parrm_debug_strip_timing(unsigned long long* h_out, int reset) {
  if (h_out) cudaMemcpyFromSymbol(h_out, parrm::g_strip_timing, sizeof(parrm::g_strip_timing));
  if (reset) {
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(parrm::g_strip_timing, zero, sizeof(zero));
  }
  return 0;
}
#endif

const char* parrm_filter_last_kernel(void) { return parrm::g_last_kernel; }

int parrm_filter_specialise_check(const void* h_plan, int dtype,
                                  const parrm_filter_options_t* options, int32_t* shape,
                                  size_t* cubin_bytes) {
  using namespace parrm;
  PARRM_REQUIRE(h_plan != nullptr, "parrm_filter_specialise_check: null plan");
  const FilterPlanHeader* hdr = static_cast<const FilterPlanHeader*>(h_plan);
  PARRM_REQUIRE(hdr->magic == kPlanMagic && hdr->version == kPlanVersion,
                "parrm_filter_specialise_check: not a filter plan");
  const int32_t* h_terms = reinterpret_cast<const int32_t*>(
      static_cast<const unsigned char*>(h_plan) + hdr->terms_offset);
  FilterTuning tune{0, options ? options->steps_per_chunk : 0,
                    options ? options->prefetch_chunks : 0, options ? options->ctas_per_sm : 0};
  CombEShape s;
  if (!comb_e_shape(hdr, h_terms, dtype, &tune, &s)) {
    set_error("parrm_filter_specialise_check: this plan is outside the specialised kernel's range");
    return PARRM_ERR_UNSUPPORTED;
  }
  if (shape) {
    const int32_t v[12] = {s.d, s.nk, s.m[0], s.m[1], s.nb[0], s.nb[1], s.n_plus + s.n_minus,
                           s.u, s.pf, s.ctas, s.smem_bytes, ((s.d + 31) / 32) * 32 + 32};
    for (int i = 0; i < 12; ++i) shape[i] = v[i];
  }
  return comb_e_compile_only(s, cubin_bytes);
}

int parrm_filter_apply(const void* d_x, int64_t ld_x, int64_t x_t0, int64_t n_x, void* d_out,
                       int64_t ld_out, int64_t t0, int64_t n_out, int64_t n_samples_total,
                       int64_t n_chans, const void* d_plan, const void* h_plan, int dtype,
                       void* stream) {
  return parrm_filter_apply_ex(d_x, ld_x, x_t0, n_x, d_out, ld_out, t0, n_out, n_samples_total,
                               n_chans, d_plan, h_plan, dtype, nullptr, stream);
}

int parrm_filter_apply_ex(const void* d_x, int64_t ld_x, int64_t x_t0, int64_t n_x, void* d_out,
                          int64_t ld_out, int64_t t0, int64_t n_out, int64_t n_samples_total,
                          int64_t n_chans, const void* d_plan, const void* h_plan, int dtype,
                          const parrm_filter_options_t* options, void* stream) {
  using namespace parrm;
  PARRM_REQUIRE(d_plan != nullptr && h_plan != nullptr, "parrm_filter_apply: null plan");
  const FilterPlanHeader* hdr = static_cast<const FilterPlanHeader*>(h_plan);
  PARRM_REQUIRE(hdr->magic == kPlanMagic && hdr->version == kPlanVersion,
                "parrm_filter_apply: not a filter plan");
  PARRM_REQUIRE(hdr->dtype == dtype, "parrm_filter_apply: plan built for another dtype");
  PARRM_REQUIRE(dtype == PARRM_F64 || dtype == PARRM_F32, "parrm_filter_apply: bad dtype %d", dtype);
  PARRM_REQUIRE(n_chans >= 0 && n_out >= 0 && n_x >= 0 && n_samples_total >= 0,
                "parrm_filter_apply: negative size");
  PARRM_REQUIRE(n_chans <= 65535, "parrm_filter_apply: more than 65535 channels per call");
  const int want = options ? options->kernel : PARRM_FILTER_KERNEL_AUTO;
  PARRM_REQUIRE(want >= PARRM_FILTER_KERNEL_AUTO && want <= PARRM_FILTER_KERNEL_SPECIALISED,
                "parrm_filter_apply: unknown kernel choice %d", want);
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
  const unsigned char* h_base = static_cast<const unsigned char*>(h_plan);
  const int32_t* d_taps = reinterpret_cast<const int32_t*>(
      static_cast<const unsigned char*>(d_plan) + hdr->taps_offset);
  const int32_t* h_terms = reinterpret_cast<const int32_t*>(h_base + hdr->terms_offset);
  cudaStream_t s = as_stream(stream);

  // 1. kernel specialised for this plan at run time (pattern-first comb, filter_comb_e.cuh).
  //    Building it costs about a second once per plan, so short one-off calls keep the
  //    pre-built kernels unless the specialisation already exists.
  if (hdr->kind == kPlanComb &&
      (want == PARRM_FILTER_KERNEL_AUTO || want == PARRM_FILTER_KERNEL_SPECIALISED)) {
    FilterTuning tune{want, options ? options->steps_per_chunk : 0,
                      options ? options->prefetch_chunks : 0, options ? options->ctas_per_sm : 0};
    CombEShape shape;
    const bool fits = comb_e_shape(hdr, h_terms, dtype, &tune, &shape);
    const bool worth = want == PARRM_FILTER_KERNEL_SPECIALISED ||
                       n_chans * n_out >= (int64_t(1) << 24) || (fits && comb_e_cached(shape));
    if (fits && worth) {
      const unsigned char* d_base = static_cast<const unsigned char*>(d_plan);
      const int rc = launch_comb_e(
          shape, d_x, d_out, reinterpret_cast<const int32_t*>(d_base + hdr->count_offset),
          reinterpret_cast<const double*>(d_base + hdr->recip_offset), ld_x, x_t0, n_x, ld_out,
          t0, n_out, n_samples_total, n_chans, s, nullptr);
      if (rc == PARRM_OK) {
        g_last_kernel = "parrm_filter_comb_e";
        return rc;
      }
      if (rc != PARRM_ERR_UNSUPPORTED || want == PARRM_FILTER_KERNEL_SPECIALISED) return rc;
    } else if (want == PARRM_FILTER_KERNEL_SPECIALISED) {
      set_error("parrm_filter_apply: this plan is outside the specialised kernel's range");
      return PARRM_ERR_UNSUPPORTED;
    }
  } else if (want == PARRM_FILTER_KERNEL_SPECIALISED) {
    set_error("parrm_filter_apply: the specialised kernel needs a comb plan");
    return PARRM_ERR_UNSUPPORTED;
  }

  const bool try_strip = hdr->kind == kPlanComb && want != PARRM_FILTER_KERNEL_GATHER;
  if (dtype == PARRM_F64) {
    FilterArgs<double> a{static_cast<const double*>(d_x), static_cast<double*>(d_out),
                         d_taps, ld_x, x_t0, n_x, ld_out, t0, n_out, n_samples_total,
                         hdr->n_taps, w_lo, w_hi, 0};
    if (try_strip) {
      bool launched = false;
      const int rc = launch_strip<double>(hdr, h_terms, d_taps, a, n_chans, s, &launched);
      if (rc != PARRM_OK || launched) return rc;
    }
    return launch_filter<double>(a, n_chans, s);
  }
  FilterArgs<float> a{static_cast<const float*>(d_x), static_cast<float*>(d_out),
                      d_taps, ld_x, x_t0, n_x, ld_out, t0, n_out, n_samples_total,
                      hdr->n_taps, w_lo, w_hi, 0};
  if (try_strip) {
    bool launched = false;
    const int rc = launch_strip<float>(hdr, h_terms, d_taps, a, n_chans, s, &launched);
    if (rc != PARRM_OK || launched) return rc;
  }
  return launch_filter<float>(a, n_chans, s);
}

}  // extern "C"
