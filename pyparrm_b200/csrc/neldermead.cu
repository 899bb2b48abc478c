// Device-resident 1-D Nelder-Mead: the refinement stages of find_period (parrm.py:499-517,
// 545-550, scipy.optimize.fmin at its defaults) without a host round trip per iteration.
//
// The host-side restatement (pyparrm_b200/_neldermead.py) runs every chain as a small state
// machine in lock step: one round asks the evaluator for the five points any chain could need
// in its current iteration (reflection, expansion, outside / inside contraction, shrink point)
// and then consumes exactly the values SciPy's control flow would have looked at, with the
// same comparisons and the same call counting.  Here that state machine is one tiny kernel, so
// a round is  parrm_eval_periods (d_points -> d_values)  +  parrm_nm_step  on the stream, and
// several rounds are captured into one CUDA graph; the host only reads the count of chains
// still running after each replay.
//
// Bit-exactness: the simplex arithmetic is written with round-to-nearest intrinsics in the
// order NumPy evaluates it ((1 + rho) * xbar - rho * worst, ...), never contracted to FMAs, so
// the trajectory equals SciPy's whenever the objective values are equal.
#include "common.cuh"

namespace parrm {

struct NmChain {
  double sim[2];
  double fsim[2];
  double points[5];
  int32_t fcalls, iterations, done, started;
};
static_assert(sizeof(NmChain) == PARRM_NM_STATE_BYTES, "NmChain layout is part of the ABI");

__device__ __forceinline__ void nm_sort(NmChain& c) {
  // numpy.argsort on two values: ascending, NaN last, stable
  const bool swap = (c.fsim[1] < c.fsim[0]) || (isnan(c.fsim[0]) && !isnan(c.fsim[1]));
  if (swap) {
    const double x = c.sim[0], f = c.fsim[0];
    c.sim[0] = c.sim[1];
    c.fsim[0] = c.fsim[1];
    c.sim[1] = x;
    c.fsim[1] = f;
  }
}

__device__ __forceinline__ void nm_check_done(NmChain& c, double xtol, double ftol, int maxiter,
                                              int maxfun) {
  if (!(c.fcalls < maxfun && c.iterations < maxiter)) {
    c.done = 1;
    return;
  }
  if (fabs(__dsub_rn(c.sim[1], c.sim[0])) <= xtol && fabs(__dsub_rn(c.fsim[0], c.fsim[1])) <= ftol)
    c.done = 1;
}

// SciPy's function wrapper refuses the call once maxfun evaluations have been spent.
#define NM_SPEND()            \
  do {                        \
    if (c.fcalls >= maxfun) goto out_of_calls; \
    ++c.fcalls;               \
  } while (0)

__device__ void nm_advance(NmChain& c, const double* v, int maxfun) {
  const double xr = c.points[0], xe = c.points[1], xc = c.points[2], xcc = c.points[3],
               xs = c.points[4];
  {
    NM_SPEND();
    const double fxr = v[0];
    bool shrink = false;
    if (fxr < c.fsim[0]) {
      NM_SPEND();
      const double fxe = v[1];
      if (fxe < fxr) {
        c.sim[1] = xe;
        c.fsim[1] = fxe;
      } else {
        c.sim[1] = xr;
        c.fsim[1] = fxr;
      }
    } else if (fxr < c.fsim[0]) {  // "second worst" is the best vertex when N = 1: never true
      c.sim[1] = xr;
      c.fsim[1] = fxr;
    } else {
      if (fxr < c.fsim[1]) {
        NM_SPEND();
        const double fxc = v[2];
        if (fxc <= fxr) {
          c.sim[1] = xc;
          c.fsim[1] = fxc;
        } else {
          shrink = true;
        }
      } else {
        NM_SPEND();
        const double fxcc = v[3];
        if (fxcc < c.fsim[1]) {
          c.sim[1] = xcc;
          c.fsim[1] = fxcc;
        } else {
          shrink = true;
        }
      }
      if (shrink) {
        c.sim[1] = xs;
        NM_SPEND();
        c.fsim[1] = v[4];
      }
    }
    ++c.iterations;
  }
out_of_calls:
  return;
}

__device__ __forceinline__ void nm_propose(NmChain& c) {
  const double best = c.sim[0], worst = c.sim[1];
  const double xbar = best;  // centroid of all vertices but the worst
  c.points[0] = __dsub_rn(__dmul_rn(2.0, xbar), worst);                        // (1+rho) xbar - rho w
  c.points[1] = __dsub_rn(__dmul_rn(3.0, xbar), __dmul_rn(2.0, worst));        // (1+rho chi) ..
  c.points[2] = __dsub_rn(__dmul_rn(1.5, xbar), __dmul_rn(0.5, worst));        // (1+psi rho) ..
  c.points[3] = __dadd_rn(__dmul_rn(0.5, xbar), __dmul_rn(0.5, worst));        // (1-psi) xbar + psi w
  c.points[4] = __dadd_rn(best, __dmul_rn(0.5, __dsub_rn(worst, best)));       // shrink
}

__global__ void nm_init_kernel(const double* __restrict__ starts, int n_chains,
                               NmChain* __restrict__ state, double* __restrict__ points) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_chains) return;
  NmChain c;
  const double x0 = starts[k];
  c.sim[0] = x0;
  c.sim[1] = x0 != 0.0 ? __dmul_rn(1.05, x0) : 0.00025;
  c.fsim[0] = c.fsim[1] = __longlong_as_double(0x7ff0000000000000LL);
  c.fcalls = 0;
  c.iterations = 1;
  c.done = 0;
  c.started = 0;
  for (int i = 0; i < 5; ++i) c.points[i] = c.sim[0];
  c.points[1] = c.sim[1];
  state[k] = c;
  for (int i = 0; i < 5; ++i) points[5 * k + i] = c.points[i];
}

__global__ void nm_step_kernel(NmChain* __restrict__ state, int n_chains,
                               const double* __restrict__ values, double* __restrict__ points,
                               double xtol, double ftol, int maxiter, int maxfun,
                               int32_t* __restrict__ n_active) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) *n_active = 0;
  __syncthreads();
  if (k >= n_chains) return;
  NmChain c = state[k];
  if (!c.done) {
    if (!c.started) {  // the two vertices of the initial simplex
      for (int i = 0; i < 2; ++i) {
        if (c.fcalls >= maxfun) break;
        ++c.fcalls;
        c.fsim[i] = values[5 * k + i];
      }
      c.started = 1;
    } else {
      nm_advance(c, values + 5 * k, maxfun);
    }
    nm_sort(c);
    nm_check_done(c, xtol, ftol, maxiter, maxfun);
    if (!c.done) {
      nm_propose(c);
    } else {
      for (int i = 0; i < 5; ++i) c.points[i] = c.sim[0];  // keeps the evaluator fed, ignored
    }
    state[k] = c;
    for (int i = 0; i < 5; ++i) points[5 * k + i] = c.points[i];
    if (!c.done) atomicAdd(n_active, 1);
  }
}

}  // namespace parrm

extern "C" {

int parrm_nm_init(const double* d_starts, int32_t n_chains, void* d_state, double* d_points,
                  void* stream) {
  PARRM_REQUIRE(n_chains >= 1 && n_chains <= 1024, "parrm_nm_init: 1..1024 chains");
  PARRM_REQUIRE(d_starts && d_state && d_points, "parrm_nm_init: null pointer");
  parrm::nm_init_kernel<<<1, 1024, 0, parrm::as_stream(stream)>>>(
      d_starts, n_chains, static_cast<parrm::NmChain*>(d_state), d_points);
  PARRM_LAUNCH_OK("nm_init_kernel");
  return PARRM_OK;
}

int parrm_nm_step(void* d_state, int32_t n_chains, const double* d_values, double* d_points,
                  double xtol, double ftol, int32_t maxiter, int32_t maxfun, int32_t* d_n_active,
                  void* stream) {
  PARRM_NVTX("parrm_nm_step");
  PARRM_REQUIRE(n_chains >= 1 && n_chains <= 1024, "parrm_nm_step: 1..1024 chains");
  PARRM_REQUIRE(d_state && d_values && d_points && d_n_active, "parrm_nm_step: null pointer");
  parrm::nm_step_kernel<<<1, 1024, 0, parrm::as_stream(stream)>>>(
      static_cast<parrm::NmChain*>(d_state), n_chains, d_values, d_points, xtol, ftol, maxiter,
      maxfun, d_n_active);
  PARRM_LAUNCH_OK("nm_step_kernel");
  return PARRM_OK;
}

}  // extern "C"
