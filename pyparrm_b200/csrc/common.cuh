// Shared helpers for libparrm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "parrm_b200.h"

namespace parrm {

// Thread-local last-error string behind parrm_last_error().
void set_error(const char* fmt, ...);

inline int cuda_fail(cudaError_t err, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(err));
  return PARRM_ERR_CUDA;
}

#define PARRM_CUDA_OK(expr)                                  \
  do {                                                       \
    cudaError_t err__ = (expr);                              \
    if (err__ != cudaSuccess) return ::parrm::cuda_fail(err__, #expr); \
  } while (0)

#define PARRM_REQUIRE(cond, ...)              \
  do {                                        \
    if (!(cond)) {                            \
      ::parrm::set_error(__VA_ARGS__);        \
      return PARRM_ERR_INVALID_ARGUMENT;      \
    }                                         \
  } while (0)

// Launch check: catches configuration errors at enqueue time without synchronising.
#define PARRM_LAUNCH_OK(name)                                        \
  do {                                                               \
    cudaError_t err__ = cudaGetLastError();                          \
    if (err__ != cudaSuccess) return ::parrm::cuda_fail(err__, name); \
  } while (0)

// NVTX range around a C-ABI call (header-only NVTX 3: a no-op unless a profiler is attached).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define PARRM_NVTX(name) ::parrm::NvtxRange parrm_nvtx_range__(name)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// plain arrival (release.cta) for producer/consumer hand-offs between warps of one CTA
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over a subset of the CTA's warps (count = participating threads, multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace parrm
