// Run-time specialisation of the pattern-first comb kernel (filter_comb_e.cuh) for one
// filter plan: the plan's stride, box lengths and tap offsets become compile-time constants
// of a kernel built with NVRTC for sm_100a and cached per (device, plan shape, tuning).
//
// NVRTC is opened with dlopen (it ships with the CUDA toolkit of this image and with torch);
// the driver entry points come from cudaGetDriverEntryPoint, so the library links against
// neither libnvrtc nor libcuda.  When NVRTC cannot be opened the caller (filter.cu) keeps the
// pre-built strip kernels -- still CUDA, just not specialised -- and parrm_filter_last_kernel()
// says which one ran.
#include <cuda.h>
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "filter_jit.h"
#include "filter_plan.h"

namespace parrm {
namespace {

const char kCombESource[] =
#include "filter_comb_e_src.inc"
    ;

// ---- NVRTC through dlopen ----------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* handle = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                       const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nvrtc* nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* name : names) {
      n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
#define PARRM_NVRTC_SYM(field, sym)                                             \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, sym));          \
  if (!n.field) return;
    PARRM_NVRTC_SYM(CreateProgram, "nvrtcCreateProgram")
    PARRM_NVRTC_SYM(CompileProgram, "nvrtcCompileProgram")
    PARRM_NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    PARRM_NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
    PARRM_NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    PARRM_NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
    PARRM_NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
    PARRM_NVRTC_SYM(GetErrorString, "nvrtcGetErrorString")
#undef PARRM_NVRTC_SYM
    n.ok = true;
  });
  return &n;
}

// ---- driver entry points through the runtime ------------------------------------------
struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

Driver* driver() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult st;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess &&
             st == cudaDriverEntryPointSuccess && *fn != nullptr;
    };
    d.ok = get("cuModuleLoadData", reinterpret_cast<void**>(&d.ModuleLoadData)) &&
           get("cuModuleGetFunction", reinterpret_cast<void**>(&d.ModuleGetFunction)) &&
           get("cuFuncSetAttribute", reinterpret_cast<void**>(&d.FuncSetAttribute)) &&
           get("cuFuncGetAttribute", reinterpret_cast<void**>(&d.FuncGetAttribute)) &&
           get("cuLaunchKernel", reinterpret_cast<void**>(&d.LaunchKernel)) &&
           get("cuGetErrorString", reinterpret_cast<void**>(&d.GetErrorString));
  });
  return &d;
}

int driver_fail(CUresult rc, const char* what) {
  const char* msg = nullptr;
  if (driver()->GetErrorString) driver()->GetErrorString(rc, &msg);
  set_error("%s: %s", what, msg ? msg : "driver error");
  return PARRM_ERR_CUDA;
}

struct Kernel {
  CUfunction fn = nullptr;
  int smem_bytes = 0, threads = 0, regs = 0, ctas_per_sm = 1;
  int chunk = 0, priming_groups = 0;
};

std::mutex g_mutex;
std::map<std::string, Kernel> g_cache;  // key: device + option string

std::string join(const int32_t* v, int n) {
  std::string s;
  for (int i = 0; i < n; ++i) {
    if (i) s += ",";
    s += std::to_string(v[i]);
  }
  if (n == 0) s = "0";
  return s;
}

}  // namespace

// Shape of the specialisation the plan would get; false when the plan is outside what the
// pattern-first kernel handles (the caller then keeps the strip kernels).
bool comb_e_shape(const FilterPlanHeader* hdr, const int32_t* terms, int dtype,
                  const FilterTuning* tune, CombEShape* out) {
  const FilterPlanHeader* sec = hdr;
  if (hdr->kind != kPlanComb || sec->n_kinds < 1 || sec->n_kinds > 2) return false;
  CombEShape s;
  memset(&s, 0, sizeof(s));
  s.es = dtype == PARRM_F64 ? 8 : 4;
  s.d = sec->stride;
  s.nk = sec->n_kinds;
  const int nb[2] = {sec->n_box[0], sec->n_kinds > 1 ? sec->n_box[1] : 0};
  int first = 0;  // kind with the longer box goes first
  if (s.nk == 2 && sec->window[1] > sec->window[0]) first = 1;
  const int32_t* box[2] = {terms, terms + nb[0]};
  s.m[0] = sec->window[first];
  s.nb[0] = nb[first];
  s.off[0] = box[first];
  if (s.nk == 2) {
    s.m[1] = sec->window[1 - first];
    s.nb[1] = nb[1 - first];
    s.off[1] = box[1 - first];
  } else {
    s.m[1] = 1;
    s.nb[1] = 0;
    s.off[1] = nullptr;
  }
  s.n_plus = sec->n_plus;
  s.n_minus = sec->n_minus;
  s.plus = terms + nb[0] + nb[1];
  s.minus = s.plus + s.n_plus;
  s.centre = sec->centre;
  s.n_taps = hdr->n_taps;
  s.w_lo = hdr->w_min < 0 ? hdr->w_min : 0;
  s.w_hi = hdr->w_max > 0 ? hdr->w_max : 0;
  int lo = 0, hi = 0;
  const int n_all = nb[0] + nb[1] + s.n_plus + s.n_minus;
  for (int i = 0; i < n_all; ++i) {
    lo = terms[i] < lo ? terms[i] : lo;
    hi = terms[i] > hi ? terms[i] : hi;
  }
  s.back = hi;
  s.fwd = -lo;
  if (s.d < kPatternFirstMinStride || s.d > kPatternFirstMaxStride) return false;
  if (n_all > 96 || s.nb[0] < 1) return false;
  // Shape choice, from interleaved A/B runs on one box (scripts/ab_filter_shapes.py, fraction of
  // the HBM roofline at full size): (1) two CTAs per SM whenever the register rings (live:
  // M0 + M1 values) leave ~50 registers for the rest of the step inside the per-thread budget
  // at that occupancy, even at the price of a few spilled bytes -- cfg3 (M0 + M1 = 37): 0.69-0.75
  // with two CTAs, 0.60 with one; cfg2: 0.70 / 0.68.  (2) a chunk of U steps with U a divisor
  // of M0, so that the unrolled block B = U * ceil(M0 / U) is exactly M0 steps -- cfg3 U = 5
  // (B = 25) 0.69-0.75, U = 7 (B = 28) 0.65-0.69; cfg4 U = 10 0.63-0.64, U = 8 (B = 16) 0.58-0.64;
  // cfg1 U = 5 0.61, U = 3 0.60 -- the largest such chunk up to 32 KB that fits shared memory;
  // without a usable divisor, the largest chunk up to 20 KB.  A chunk is a whole number of
  // 16-byte units.
  const int ring_regs = (s.m[0] + (s.nk > 1 ? s.m[1] : 0)) * (s.es / 4);
  const int threads = pattern_first_threads(s.d);
  auto reg_budget = [&](int ctas) { return pattern_first_reg_budget(s.d, ctas); };
  // (the planner proposes only plans inside this limit: pattern_first_max_ring)
  if (ring_regs + kPatternFirstStepRegs > reg_budget(1)) return false;
  int ctas_first = ring_regs + kPatternFirstStepRegs <= reg_budget(2) ? 2 : 1;
  // float32: half the wavefronts per output, so the step loop is latency-bound and a third CTA
  // per SM pays where the rings are short (cfg3 shape, float32: 0.62 of its 8 B roofline with
  // three CTAs and 4 KB chunks, 0.52 with two CTAs and 16 KB; cfg2 / cfg4: no difference)
  if (s.es == 4 && 3 * threads <= 2048 && ring_regs + 40 <= reg_budget(3)) ctas_first = 3;
  if (tune && tune->ctas_per_sm > 0) ctas_first = tune->ctas_per_sm;
  // chunks in flight beyond the window: 2 measured best wherever more would fit (cfg3 shape,
  // one CTA per SM: 0.637 of the HBM roofline with 2, 0.606 with 4; cfg1 taps: 0.591 / 0.582)
  const int pf_max = (tune && tune->prefetch_chunks > 0) ? tune->prefetch_chunks : 2;
  const int pf_min = std::min(pf_max, 2);
  auto smem_for = [&](int u, int pf) {
    const int64_t ch = int64_t(u) * s.d;
    const int64_t hb = (s.back + ch - 1) / ch, hf = (s.fwd + ch - 1) / ch;
    const int64_t q = hb + hf + 1 + pf;
    return ((2 * q * 8 + 127) / 128) * 128 + (q + hb + hf) * ch * s.es;
  };
  auto usable = [&](int u) {
    return u >= 1 && (int64_t(u) * s.d * s.es) % 16 == 0 && (s.m[0] + u - 1) / u * u <= 64;
  };
  for (int ctas = ctas_first; ctas >= 1; --ctas) {
    const int64_t budget = int64_t(227) * 1024 / ctas - 1024;
    int u = 0;
    if (tune && tune->steps_per_chunk > 0) {
      u = tune->steps_per_chunk;
      if (!usable(u) || smem_for(u, pf_min) > budget) u = 0;
    } else {
      for (int cand = s.m[0]; cand >= 1 && u == 0; --cand)  // largest divisor of M0, <= 32 KB
        if (s.m[0] % cand == 0 && int64_t(cand) * s.d * s.es <= 32768 &&
            int64_t(cand) * s.d >= 512 && usable(cand) && smem_for(cand, pf_min) <= budget)
          u = cand;
      const int u_hi = int(std::max<int64_t>(1, 20480 / (int64_t(s.d) * s.es)));
      for (int cand = u_hi; cand >= 1 && u == 0; --cand)  // largest chunk <= 20 KB that fits
        if (usable(cand) && smem_for(cand, pf_min) <= budget) u = cand;
      for (int cand = 1; cand <= 64 && u == 0; ++cand)    // stride alone exceeds 20 KB
        if (usable(cand) && smem_for(cand, pf_min) <= budget) u = cand;
    }
    if (u == 0) continue;
    int pf = pf_min;
    while (pf < pf_max && smem_for(u, pf + 1) <= budget) ++pf;
    s.smem_bytes = int(smem_for(u, pf));
    s.ctas = ctas;
    s.u = u;
    s.pf = pf;
    s.variant = tune ? tune->variant : 0;
    *out = s;
    return true;
  }
  return false;
}

static int compile_cubin(const CombEShape& s, std::vector<char>* cubin) {
  Nvrtc* n = nvrtc();
  if (!n->ok) {
    const char* why = dlerror();
    set_error("parrm filter: NVRTC (libnvrtc.so.12) could not be opened: %s", why ? why : "?");
    return PARRM_ERR_UNSUPPORTED;
  }
  std::vector<std::string> opts = {
      "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
      std::string("-DPE_T=") + (s.es == 8 ? "double" : "float"),
      "-DPE_D=" + std::to_string(s.d), "-DPE_NK=" + std::to_string(s.nk),
      "-DPE_M0=" + std::to_string(s.m[0]), "-DPE_M1=" + std::to_string(s.m[1]),
      "-DPE_NB0=" + std::to_string(s.nb[0]), "-DPE_NB1=" + std::to_string(s.nb[1]),
      "-DPE_OFF0=" + join(s.off[0], s.nb[0]), "-DPE_OFF1=" + join(s.off[1], s.nb[1]),
      "-DPE_NPLUS=" + std::to_string(s.n_plus), "-DPE_PLUS=" + join(s.plus, s.n_plus),
      "-DPE_NMINUS=" + std::to_string(s.n_minus), "-DPE_MINUS=" + join(s.minus, s.n_minus),
      "-DPE_CENTRE=" + std::to_string(s.centre), "-DPE_U=" + std::to_string(s.u),
      "-DPE_PF=" + std::to_string(s.pf), "-DPE_NTAPS=" + std::to_string(s.n_taps),
      "-DPE_WLO=" + std::to_string(s.w_lo), "-DPE_WHI=" + std::to_string(s.w_hi),
      "-DPE_BACK=" + std::to_string(s.back), "-DPE_FWD=" + std::to_string(s.fwd),
      "-DPE_CTAS=" + std::to_string(s.ctas), "-DPE_VARIANT=" + std::to_string(s.variant)};
  std::vector<const char*> copts;
  for (const std::string& o : opts) copts.push_back(o.c_str());
  nvrtcProgram prog = nullptr;
  int rc = n->CreateProgram(&prog, kCombESource, "filter_comb_e.cu", 0, nullptr, nullptr);
  if (rc != 0) {
    set_error("nvrtcCreateProgram: %s", n->GetErrorString(rc));
    return PARRM_ERR_CUDA;
  }
  rc = n->CompileProgram(prog, int(copts.size()), copts.data());
  if (rc != 0) {
    size_t log_size = 0;
    n->GetProgramLogSize(prog, &log_size);
    std::string log(log_size + 1, '\0');
    if (log_size) n->GetProgramLog(prog, &log[0]);
    set_error("nvrtcCompileProgram: %s\n%.380s", n->GetErrorString(rc), log.c_str());
    n->DestroyProgram(&prog);
    return PARRM_ERR_CUDA;
  }
  size_t cubin_size = 0;
  n->GetCUBINSize(prog, &cubin_size);
  cubin->resize(cubin_size);
  rc = n->GetCUBIN(prog, cubin->data());
  n->DestroyProgram(&prog);
  if (rc != 0 || cubin_size == 0) {
    set_error("nvrtcGetCUBIN: %s", n->GetErrorString(rc));
    return PARRM_ERR_CUDA;
  }
  return PARRM_OK;
}

int comb_e_compile_only(const CombEShape& s, size_t* cubin_bytes) {
  std::vector<char> cubin;
  const int rc = compile_cubin(s, &cubin);
  if (rc == PARRM_OK && cubin_bytes) *cubin_bytes = cubin.size();
  return rc;
}

static int build_kernel(const CombEShape& s, Kernel* out) {
  Driver* d = driver();
  if (!d->ok) {
    set_error("parrm filter: CUDA driver entry points unavailable");
    return PARRM_ERR_UNSUPPORTED;
  }
  std::vector<char> cubin;
  const int rc = compile_cubin(s, &cubin);
  if (rc != PARRM_OK) return rc;
  PARRM_CUDA_OK(cudaFree(nullptr));  // the primary context exists and is current
  CUmodule mod = nullptr;
  CUresult cr = d->ModuleLoadData(&mod, cubin.data());
  if (cr != CUDA_SUCCESS) return driver_fail(cr, "cuModuleLoadData");
  Kernel k;
  cr = d->ModuleGetFunction(&k.fn, mod, "parrm_filter_comb_e");
  if (cr != CUDA_SUCCESS) return driver_fail(cr, "cuModuleGetFunction");
  k.smem_bytes = s.smem_bytes;
  k.threads = ((s.d + 31) / 32) * 32 + 32;
  k.ctas_per_sm = s.ctas;
  k.chunk = s.u * s.d;
  k.priming_groups = (s.m[0] + s.u - 1) / s.u;
  cr = d->FuncSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, k.smem_bytes);
  if (cr != CUDA_SUCCESS) return driver_fail(cr, "cuFuncSetAttribute(max dynamic shared memory)");
  d->FuncSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
  d->FuncGetAttribute(&k.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, k.fn);
  *out = k;
  return PARRM_OK;
}

bool comb_e_cached(const CombEShape& s) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  const std::string key = comb_e_key(s, dev);
  std::lock_guard<std::mutex> lock(g_mutex);
  return g_cache.count(key) != 0;
}

std::string comb_e_key(const CombEShape& s, int dev) {
  std::string key = std::to_string(dev) + "|" + std::to_string(s.es) + "|" + std::to_string(s.d) +
                    "|" + std::to_string(s.nk) + "|" + std::to_string(s.m[0]) + "|" +
                    std::to_string(s.m[1]) + "|" + join(s.off[0], s.nb[0]) + "|" +
                    join(s.off[1], s.nb[1]) + "|" + join(s.plus, s.n_plus) + "|" +
                    join(s.minus, s.n_minus) + "|" + std::to_string(s.centre) + "|" +
                    std::to_string(s.u) + "|" + std::to_string(s.pf) + "|" +
                    std::to_string(s.n_taps) + "|" + std::to_string(s.w_lo) + "|" +
                    std::to_string(s.w_hi) + "|" + std::to_string(s.ctas) + "|" +
                    std::to_string(s.variant);
  return key;
}

// Launches the specialised kernel (building it on first use).  PARRM_ERR_UNSUPPORTED means
// "could not specialise here"; the caller falls back to the pre-built kernels.
int launch_comb_e(const CombEShape& s, const void* d_x, void* d_out, const int32_t* d_taps,
                  const int32_t* d_count, const double* d_recip,
                  int64_t ld_x, int64_t x_t0, int64_t n_x, int64_t ld_out, int64_t t0,
                  int64_t n_out, int64_t n_total, int64_t n_chans, cudaStream_t stream,
                  int* regs_out, unsigned long long timeline) {
  int dev = 0;
  PARRM_CUDA_OK(cudaGetDevice(&dev));
  const std::string key = comb_e_key(s, dev);
  Kernel k;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    auto it = g_cache.find(key);
    if (it == g_cache.end()) {
      const int rc = build_kernel(s, &k);
      if (rc != PARRM_OK) return rc;
      g_cache[key] = k;
    } else {
      k = it->second;
    }
  }
  if (regs_out) *regs_out = k.regs;
  struct Args {
    const void* x;
    void* out;
    const int32_t* taps;
    const int32_t* count;
    const double* recip;
    int64_t ld_x, x_t0, n_x, ld_out, t0, n_out, n_total, total_groups;
    int32_t groups_per_chan, edge_start, edge_total, pad;
    unsigned long long timeline;
    unsigned char consts[16];  // T neg_inv_n, t_max in the kernel's element type
  } a;
  a.x = d_x; a.out = d_out; a.taps = d_taps; a.count = d_count; a.recip = d_recip;
  a.ld_x = ld_x; a.x_t0 = x_t0; a.n_x = n_x;
  a.ld_out = ld_out; a.t0 = t0; a.n_out = n_out; a.n_total = n_total;
  const int64_t ch = k.chunk;
  a.groups_per_chan = int32_t(ceil_div(n_out + ch - 1, ch));
  // cost of a channel's ends in groups (see the kernel): priming block + window fill in front,
  // and the generic blocks (about 2.5x the time of a fast one) at either recording edge;
  // measured with scripts/filter_timeline.py: 11.6 groups at the cfg2 shape (2 groups per
  // block), 33 at cfg3 (5), 7 at cfg4 (1)
  const int gpb = k.priming_groups;
  const int edge_end = 2 * gpb;
  a.edge_start = s.variant & 4 ? 0 : 2 * gpb + 2 + edge_end;
  a.edge_total = s.variant & 4 ? 0 : a.edge_start + edge_end;
  a.total_groups = n_chans * int64_t(a.groups_per_chan + a.edge_total);
  a.pad = 0;
  a.timeline = timeline;
  memset(a.consts, 0, sizeof(a.consts));
  if (s.es == 8) {
    const double v[2] = {-1.0 / double(s.n_taps), 1.7976931348623157e308};
    memcpy(a.consts, v, sizeof(v));
  } else {
    const float v[2] = {-1.0f / float(s.n_taps), 3.402823466e38f};
    memcpy(a.consts, v, sizeof(v));
  }
  // one strip per resident CTA; a strip is at least 4x its priming so the warm-up of the
  // register rings stays a small fraction of the work
  const int64_t resident = int64_t(kNumSMs) * k.ctas_per_sm;
  const int64_t min_groups = int64_t(4) * k.priming_groups;
  const int64_t grid =
      max64(1, min64(resident, n_chans * int64_t(a.groups_per_chan) / min_groups));
  void* params[] = {&a};
  const CUresult cr = driver()->LaunchKernel(k.fn, unsigned(grid), 1, 1, unsigned(k.threads), 1, 1,
                                             unsigned(k.smem_bytes), stream, params, nullptr);
  if (cr != CUDA_SUCCESS) return driver_fail(cr, "cuLaunchKernel(parrm_filter_comb_e)");
  return PARRM_OK;
}

}  // namespace parrm
