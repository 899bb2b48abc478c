// Period-candidate evaluator: PARRM._optimise_local / _fit_waves_to_data (parrm.py:552-632),
// batched over candidate periods.
//
// Reference, per candidate p and channel c (N fitted samples, M = 2*bw + 1):
//     a_i = (idx_i + 1) * (2 pi / p);   W = [1, sin(k a_i), cos(k a_i)]_{k=1..bw}      (N x M)
//     beta = solve(W'W, W'y_c);  e_c = mean_i (y_c - W beta)_i^2 + sum_j lambda*j/sum(1..M) * beta_j^2
//     fit_error(p) = sum_c e_c / n_chans
//
// Device formulation (one pass over the samples, no N x M matrix ever stored):
//   * the Gram matrix W'W does not depend on the channel (the reference rebuilds it per
//     channel) and, by the product-to-sum identities, is a linear function of the 4*bw+1
//     harmonic sums  C_m = sum_i cos(m a_i),  S_m = sum_i sin(m a_i),  m = 0..2*bw;
//   * the right-hand sides are one small GEMM  B = W' Y  (M x C, K = N);
//   * the residual sum of squares is the quadratic form  y'y - 2 beta'(W'y) + beta'(W'W) beta,
//     exact for any beta (so, like the reference's explicit residual, insensitive to first
//     order to rounding in the solve).
// Kernel 1 (accumulate) walks the samples: one accurate sincos(a_i) per sample, harmonics by
// the angle-addition recurrence in registers, harmonic sums in registers, and B from
// shared-memory tiles -- on the FP64 tensor cores (eval_accumulate_tensor_kernel, the default),
// or by scalar FMAs for one or two channels (eval_accumulate_narrow_kernel).  Samples can be
// split over several CTAs per candidate (few candidates, e.g. Nelder-Mead steps); partials are
// combined in a fixed order.
// Kernel 2 (solve) assembles the Gram matrix, factorises it with partial pivoting (zero pivot
// -> +inf, the reference's LinAlgError path), solves all channels and reduces the objective.

#include <type_traits>

#include "common.cuh"

namespace parrm {

constexpr int kSplitQuantum = 512;  // sample splits are multiples of every kernel's batch
constexpr int kKT = 64;           // narrow kernel: samples per sub-batch
constexpr int kMaxRows = 2 * PARRM_MAX_BANDWIDTH + 1;                     // 47
constexpr int kRowsPad = 48;      // padded design-matrix rows: six 8-row blocks
static_assert(kMaxRows <= kRowsPad, "row groups must cover the widest design matrix");
constexpr int kChanTile = 64;     // channels per CTA (16 column groups x 4)

struct EvalShape {
  int64_t n_chans, n_indices, n_periods;
  int bandwidth, n_rows;          // n_rows = 2*bw + 1
  int n_splits, n_chan_tiles;
  int64_t split_len;              // samples per split (multiple of kSplitQuantum)
  int64_t ld_y;
  // workspace layout (doubles)
  int64_t b_stride_split;         // n_rows * n_chans
  int64_t b_stride_period;        // 2 * n_splits * b_stride_split   (two K-halves per split)
  int64_t t_stride_split;         // 4 * bw   (C_1..C_2bw, S_1..S_2bw)
  int64_t t_stride_period;        // n_splits * t_stride_split
  int64_t t_offset;               // start of the harmonic-sum area
  int64_t c_offset;               // start of the per-channel sums of y (row 0 of W'Y)
  int row0_from_colsum;           // 1: the accumulate kernel leaves row 0 to eval_colsum_kernel
  int b_reduced;                  // 1: eval_reduce_partials_kernel has summed the partials of W'Y
  int64_t y_offset;               // start of the re-tiled copy of Y (tensor kernel, when needed)
  // what the tensor kernel reads: per channel tile a dense [n_indices][y_row_chans] block
  int y_row_chans;                // doubles per row (even; the whole row is one channel tile)
  int64_t y_tile_stride;          // doubles between channel tiles
};

__device__ __forceinline__ void cmul(double& c, double& s, double c2, double s2) {
  const double nc = fma(c, c2, -(s * s2));
  const double ns = fma(s, c2, c * s2);
  c = nc;
  s = ns;
}

// sin and cos of a phase angle a = (index + 1) * 2 pi / period (parrm.py:619).  The angles reach
// ~5e5 rad on a ten-minute recording, far beyond the threshold (|a| > 105615) above which
// CUDA's sincos() takes its Payne-Hanek slow path through local memory -- measured as the
// bottleneck of the whole evaluator.  For |a| < 1e9 a three-term Cody-Waite reduction with
// FMAs is exact to ~1e-16 rad: q = rint(a * 2/pi) < 2^30, each fma(-q, C_i, r) rounds once, and
// C1 + C2 + C3 = pi/2 to 5.6e-50.  The reduced argument |r| <= pi/4 goes through the polynomial
// kernels below and the quadrant is applied by hand.
__device__ __forceinline__ void sincos_phase(double a, double* sn, double* cs) {
  if (!(fabs(a) < 1.0e9)) {
    sincos(a, sn, cs);
    return;
  }
  const double q = rint(a * 0.6366197723675814);
  double r = fma(-q, 1.5707963267948966, a);
  r = fma(-q, 6.123233995736766e-17, r);
  r = fma(-q, -1.4973849048591698e-33, r);
  // |r| <= pi/4: minimax kernels in z = r^2 (the classic fdlibm k_sin / k_cos coefficients;
  // plain Horner evaluation is within 1.2e-16 absolute of sin / cos on this interval), two
  // independent chains instead of sincos()'s own second reduction
  const double z = r * r;
  double ps = 1.58969099521155010221e-10, pc = -1.13596475577881948265e-11;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  pc = fma(pc, z, 2.08757232129817482790e-09);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double s = fma(r * z, ps, r);
  const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int quadrant = int(static_cast<long long>(q) & 3);
  const double s1 = (quadrant & 1) ? c : s;
  const double c1 = (quadrant & 1) ? s : c;
  *sn = (quadrant & 2) ? -s1 : s1;
  *cs = ((quadrant + 1) & 2) ? -c1 : c1;
}

// (c, s) = (c1 + i s1)^n by binary powering
__device__ __forceinline__ void cpow(double c1, double s1, int n, double& c, double& s) {
  c = 1.0;
  s = 0.0;
  double bc = c1, bs = s1;
  while (n > 0) {
    if (n & 1) cmul(c, s, bc, bs);
    n >>= 1;
    if (n) cmul(bc, bs, bc, bs);
  }
}

// Shared-memory slot of channel c (0..63) of sample k in a Y tile.  A GEMM thread reads its
// four channels as two 16-byte loads (c = 4 cg .. 4 cg + 3); with the plain [k][c] layout the
// 16 column groups of a quarter-warp would span 256 bytes per load (2-way bank conflict), so
// the two halves of every group live in separate 128-byte planes.
__device__ __forceinline__ int y_slot(int k, int c) {
  return k * kChanTile + ((c >> 1) & 1) * (kChanTile / 2) + (c >> 2) * 2 + (c & 1);
}
// 8-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 writes zeros.
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------
// Tensor-core form of the accumulate kernel (default for more than two channels): B = W'Y on
// the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA; tcgen05 has no FP64 kind).
//
// One 512-thread CTA per SM alternates between two phases over 128-sample tiles:
//   generate  all 16 warps produce the tile of W' (residue class mod 8 of the harmonics of two
//             samples per thread: <= 5 dependent complex products from a table of z, z^2, z^4,
//             z^8 built once per 512-sample batch), tensor pipe idle;
//   multiply  all 16 warps issue DMMA only, operands from shared memory.
// Measured on the two overlapped forms this replaces (generator warps feeding tensor warps
// through a 3-stage ring, 159.6k candidates/s; two out-of-phase CTAs per SM, slower): a DMMA
// holds the FP64 pipe of its scheduler partition for 16 cycles and FP64 is one shared pipe, so
// the generator's dependent scalar FP64 products wait ~10x longer behind tensor work than on an
// idle pipe -- overlapping the two costs more than it hides.  Y tiles stream in by bulk copy a
// tile ahead.  The constant row of W'Y (column sums of Y) comes from eval_colsum_kernel, which
// leaves 2 bw rows = exactly five 8-row blocks at the reference's bandwidth 20.
constexpr int kGenGroups = 4;   // narrow kernel: harmonic groups (residue classes mod 4) per sample
constexpr int kGenH = (2 * PARRM_MAX_BANDWIDTH + kGenGroups - 1) / kGenGroups;  // harmonics per group
constexpr int kGenH8 = (2 * PARRM_MAX_BANDWIDTH + 7) / 8;  // harmonics per residue class mod 8
// Shared-memory tiles.  W is stored transposed, [row][tile position], Y as [sample][channel];
// row strides = 4 (mod 16) doubles make the m8n8k4 fragment loads (lane -> (l % 4, l / 4)) and
// the generator's stores (lane -> position) conflict-free.
constexpr int kYStride = 68;    // >= kChanTile
constexpr int kTensorThreads = 512;
constexpr int kTensorBatch = kTensorThreads;  // samples per sincos batch (one per thread)
constexpr int kTensorKT = 128;              // samples per tile: two per thread while generating
constexpr int kTensorWtStride = 132;        // >= kTensorKT, = 4 (mod 16)
constexpr int kTensorWtTile = kRowsPad * kTensorWtStride;
constexpr int kTensorYTile = kTensorKT * kYStride;

#ifdef PARRM_TENSOR_TIMING
// Debug build only: cycles lane 0 of every warp of CTA (0,0,0) spends in each phase.
__device__ unsigned long long g_tensor_timing[128];  // [warp][phase]
#define TENSOR_TICK(slot)                          \
  do {                                           \
    const long long now__ = clock64();           \
    tacc__[slot] += now__ - tick__;              \
    tick__ = now__;                              \
  } while (0)
#else
#define TENSOR_TICK(slot)
#endif

template <int N>
struct IntTag {
  static constexpr int value = N;
};

template <int N_MB>  // 8-row blocks of W' in use: ceil(2 bw / 8)
__global__ void __launch_bounds__(kTensorThreads, 1)
eval_accumulate_tensor_kernel(const double* __restrict__ y, const int64_t* __restrict__ indices,
                            const double* __restrict__ periods, double* __restrict__ ws,
                            const EvalShape sh) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* s_cs = reinterpret_cast<double2*>(smem_raw);  // [4][kTensorBatch] (cos, sin) of z, z^2, z^4, z^8
  double* s_w = reinterpret_cast<double*>(smem_raw + 4 * kTensorBatch * 16);  // [kRowsPad][kTensorWtStride]
  double* s_y = s_w + kTensorWtTile;                                        // 2 x [kTensorKT][kYStride]
  double* s_red = s_y + 2 * kTensorYTile;                                   // [16 warps][2 * kGenH8]

  const int tid = threadIdx.x;
  const int64_t cand = blockIdx.x;
  const int split = blockIdx.y;
  const int ctile = blockIdx.z;
  const int bw = sh.bandwidth, two_bw = 2 * bw;
  const int n_rows = sh.n_rows;
  const int64_t n_begin = int64_t(split) * sh.split_len;
  const int64_t n_end = min(n_begin + sh.split_len, sh.n_indices);
  const int chan0 = ctile * kChanTile;
  const int n_chan_here = int(min64(kChanTile, sh.n_chans - chan0));
  const double delta = 6.283185307179586 / periods[cand];  // 2*pi/period (parrm.py:619)
  // this channel tile of Y: dense rows of row_chans doubles, 16-byte aligned (the caller's array
  // when it already has that form, else the copy made by eval_retile_kernel)
  const double* const yt = y + ctile * sh.y_tile_stride;
  const int row_chans = sh.y_row_chans;

  // generator role: tile positions gi and gi + 64 (two independent chains per thread: the
  // complex products are latency-bound), residue class r8 -> harmonics r8+1, r8+9, r8+17, ...
  // Classes 2, 4, 5, 6 pay one more product for their seed; a class is two warps = two
  // scheduler partitions (warp % 4), so this order gives every partition two of them
  const int gi = tid & 63, r8 = (0x71305462 >> (4 * (tid >> 6))) & 7;  // slots 0..7 -> 2 6 4 5 0 3 1 7
  const int h = (two_bw - r8 + 7) / 8;
  double sum_c[kGenH8], sum_s[kGenH8];
#pragma unroll
  for (int j = 0; j < kGenH8; ++j) sum_c[j] = sum_s[j] = 0.0;
  // tensor role: warp (kh, nq, mh): samples kh*64..+64, channels 16 nq..+16, row blocks
  // 0..2 (mh = 0) or 3..5 (mh = 1)
  const int warp = tid >> 5, l = tid & 31;
  // (warp & 3 picks the scheduler partition: each partition gets both row halves and K halves)
  const int kh = warp >> 3, mh = (warp >> 2) & 1, nq = warp & 3;
  // tile row r = W column r + 1 (row 0: eval_colsum_kernel); a warp with mh = 0 takes the first
  // kCnt0 row blocks, mh = 1 the remaining kCnt1 -- compile-time counts, so that the DMMA
  // sequence below is straight-line code (mma.sync inside runtime conditionals costs a
  // WARPSYNC and NOPs per instruction)
  constexpr int kCnt0 = (N_MB + 1) / 2, kCnt1 = N_MB / 2;
  double acc[3][2][2];
#pragma unroll
  for (int mb = 0; mb < 3; ++mb)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

  // Tile position p (the k index of the tensor products) holds sample (p % 4) * 32 + p / 4 of
  // the tile: the four samples of one k-step then come from four 32-sample quarters.  Rows of
  // Y are dense, so a quarter is one contiguous block (16 KB at 64 channels) and a tile arrives
  // in four bulk copies (TMA), issued a tile ahead, each landing 32 bytes further round the
  // banks than the one before (the fragment-load skew a padded row stride would give).  Bulk
  // copies must be few and large: one 512-byte copy per row costs the TMA unit ~57 cycles each
  // (7300 cycles per tile, measured), and 512 threads issuing 16-byte cp.async stall ~1500
  // cycles per tile on the L2 -> SM path.
  __shared__ uint64_t s_full[2];
  const int quarter = kTensorKT / 4;  // samples per bulk copy
  const int y_k_stride = quarter * row_chans + 4;
  {
    // rows past the end of the range (and channels past row_chans) are never copied: zero
    // everything once so that W = 0 (samples past the end) never meets a non-finite leftover
    for (int e = tid; e < 2 * kTensorYTile; e += kTensorThreads) s_y[e] = 0.0;
    if (tid == 0) {
      mbar_init(&s_full[0], 1);
      mbar_init(&s_full[1], 1);
      fence_mbar_init();
    }
    fence_proxy_async();
    __syncthreads();
  }
  auto stage_y = [&](int buf, int64_t n_tile) {
    // one copy per warp 4..7: issuing a bulk copy holds the thread ~150 cycles, and these warps
    // own the smaller share of the row blocks, so they leave the multiply phase first.  The
    // barrier phase cannot complete before the arrive with the byte count, whichever order the
    // completions arrive in.
    if ((tid & 31) == 0 && (tid >> 7) == 1) {
      double* dst = s_y + buf * kTensorYTile;
      const int q = (tid >> 5) - 4;
      const int rows = int(min64(kTensorKT, n_end - n_tile));
      const uint32_t row_bytes = uint32_t(row_chans) * 8u;
      if (q == 0) mbar_expect_tx(&s_full[buf], uint32_t(rows) * row_bytes);
      if (q * quarter < rows)
        bulk_g2s(dst + q * y_k_stride, yt + (n_tile + q * quarter) * row_chans,
                 uint32_t(min(quarter, rows - q * quarter)) * row_bytes, &s_full[buf]);
    }
  };
  // rows of W' between the last harmonic and the end of its 8-row block are read by the tensor
  // products but never generated: zero them once (the first barrier of the loop orders this)
  for (int e = tid; e < (N_MB * 8 - (n_rows - 1)) * kTensorKT; e += kTensorThreads)
    s_w[(n_rows - 1 + e / kTensorKT) * kTensorWtStride + e % kTensorKT] = 0.0;
  int y_buf = 0;
  uint32_t y_phase = 0;  // bit b: parity of the next completion of s_full[b]
  if (n_begin < n_end) stage_y(0, n_begin);
#ifdef PARRM_TENSOR_TIMING
  const bool timed__ = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid & 31) == 0;
  const int tbase__ = (tid >> 5) * 8;
  long long tacc__[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tick__ = clock64();
#endif

  // thread = tile position: sample (p % 4) * 32 + p / 4 of tile tid / 128 of the batch
  const int n_mine = (tid & ~(kTensorKT - 1)) + (tid & 3) * quarter + ((tid & (kTensorKT - 1)) >> 2);
  int64_t index_mine = n_begin + n_mine < n_end ? indices[n_begin + n_mine] : 0;

  for (int64_t n_super = n_begin; n_super < n_end; n_super += kTensorBatch) {
    {  // one accurate sincos per sample of this batch; samples past the end hold z = 0
      const int64_t n = n_super + n_mine;
      double2 cs = make_double2(0.0, 0.0);
      if (n < n_end) {
        const double angle = double(index_mine + 1) * delta;
        sincos_phase(angle, &cs.y, &cs.x);
      }
      // next batch's index now: its L2 latency hides behind four tiles of work
      if (n + kTensorBatch < n_end) index_mine = indices[n + kTensorBatch];
      double2 z2 = cs, z4, z8;
      cmul(z2.x, z2.y, cs.x, cs.y);
      z4 = z2;
      cmul(z4.x, z4.y, z2.x, z2.y);
      z8 = z4;
      cmul(z8.x, z8.y, z4.x, z4.y);
      s_cs[tid] = cs;
      s_cs[kTensorBatch + tid] = z2;
      s_cs[2 * kTensorBatch + tid] = z4;
      s_cs[3 * kTensorBatch + tid] = z8;
    }
    TENSOR_TICK(0);
    __syncthreads();
    TENSOR_TICK(1);
    for (int sub = 0; sub < kTensorBatch / kTensorKT; ++sub) {
      const int64_t n_tile = n_super + sub * kTensorKT;
      if (n_tile >= n_end) break;  // uniform
      TENSOR_TICK(3);
#ifndef PARRM_DEBUG_TENSOR_NO_GEN
      {  // ---- generate: class r8 of positions gi, gi + 64 (W transposed, tile row = column - 1) ----
        // z, z^2, z^4, z^8 of both samples come from the batch table; seed z^(r8+1) costs at
        // most one more product
        const double2* pa = s_cs + sub * kTensorKT + gi;
        const double2* pb = pa + 64;
        const double2 a8 = pa[3 * kTensorBatch], b8 = pb[3 * kTensorBatch];
        const double a8c = a8.x, a8s = a8.y, b8c = b8.x, b8s = b8.y;
        double ca, sa, cb, sb;
        {
          // base power and multiplier per class: z^(r8+1) = base * mult
          //   r8: 0 z | 1 z^2 | 2 z^2 z | 3 z^4 | 4 z^4 z | 5 z^4 z^2 | 6 z^8 conj(z) | 7 z^8
          const int base = r8 == 0 ? 0 : r8 < 3 ? 1 : r8 < 6 ? 2 : 3;
          const double2 ba = pa[base * kTensorBatch], bb = pb[base * kTensorBatch];
          ca = ba.x; sa = ba.y; cb = bb.x; sb = bb.y;
          if (r8 == 2 || r8 == 4 || r8 == 5 || r8 == 6) {
            const int mult = r8 == 5 ? 1 : 0;
            double2 ma = pa[mult * kTensorBatch], mb = pb[mult * kTensorBatch];
            if (r8 == 6) {  // |z| = 1: z^8 conj(z) = z^7
              ma.y = -ma.y;
              mb.y = -mb.y;
            }
            cmul(ca, sa, ma.x, ma.y);
            cmul(cb, sb, mb.x, mb.y);
          }
        }
        double* wcol = s_w + gi;
#pragma unroll
        for (int j = 0; j < kGenH8; ++j) {
          if (j >= h) break;  // uniform within a warp
          if (j > 0) {
            cmul(ca, sa, a8c, a8s);
            cmul(cb, sb, b8c, b8s);
          }
          const int m = r8 + 1 + 8 * j;
          sum_c[j] += ca + cb;
          sum_s[j] += sa + sb;
          if (m <= bw) {  // columns 2m-1 = sin, 2m = cos (parrm.py:622-623)
            wcol[(2 * m - 2) * kTensorWtStride] = sa;
            wcol[(2 * m - 1) * kTensorWtStride] = ca;
            wcol[(2 * m - 2) * kTensorWtStride + 64] = sb;
            wcol[(2 * m - 1) * kTensorWtStride + 64] = cb;
          }
        }
      }
#endif
      TENSOR_TICK(2);
      mbar_wait(&s_full[y_buf], (y_phase >> y_buf) & 1u);
      y_phase ^= 1u << y_buf;
      TENSOR_TICK(4);
      __syncthreads();
      TENSOR_TICK(5);
#ifndef PARRM_DEBUG_TENSOR_NO_MMA
      {  // ---- multiply: B += W' Y on the FP64 tensor cores, every warp ----
        const double* yb = s_y + y_buf * kTensorYTile + (l & 3) * y_k_stride +
                           kh * (kTensorKT / 8) * row_chans + nq * 16 + (l >> 2);
        auto multiply = [&](auto count, int first_block) {
          constexpr int CNT = decltype(count)::value;
          const double* wa =
              s_w + (first_block * 8 + (l >> 2)) * kTensorWtStride + kh * (kTensorKT / 2) + (l & 3);
#pragma unroll
          for (int step = 0; step < kTensorKT / 8; ++step) {
            double a[CNT > 0 ? CNT : 1], b[2];
#pragma unroll
            for (int mb = 0; mb < CNT; ++mb) a[mb] = wa[mb * 8 * kTensorWtStride + step * 4];
            b[0] = yb[step * row_chans];
            b[1] = yb[step * row_chans + 8];
#pragma unroll
            for (int mb = 0; mb < CNT; ++mb)
#pragma unroll
              for (int nb = 0; nb < 2; ++nb)
                asm volatile(
                    "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(acc[mb][nb][0]), "+d"(acc[mb][nb][1])
                    : "d"(a[mb]), "d"(b[nb]));
          }
        };
        if (mh == 0) multiply(IntTag<kCnt0>{}, 0);
        else multiply(IntTag<kCnt1>{}, kCnt0);
      }
#endif
      // next tile's Y: the other buffer was last read before the barrier that ended the previous
      // tile
      if (n_tile + kTensorKT < n_end) stage_y(y_buf ^ 1, n_tile + kTensorKT);
      TENSOR_TICK(6);
      __syncthreads();
      TENSOR_TICK(7);
      y_buf ^= 1;
    }
  }

#ifdef PARRM_TENSOR_TIMING
  if (timed__)
    for (int i = 0; i < 8; ++i) g_tensor_timing[tbase__ + i] += (unsigned long long)tacc__[i];
#endif
  {  // ---- write the B partial of this (candidate, split, K-half) ----
    double* bp = ws + cand * sh.b_stride_period + (int64_t(split) * 2 + kh) * sh.b_stride_split;
#pragma unroll
    for (int mb = 0; mb < 3; ++mb) {
      const int row = ((mh ? kCnt0 : 0) + mb) * 8 + (l >> 2) + 1;
      if (mb < (mh ? kCnt1 : kCnt0) && row < n_rows) {
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int ch = nq * 16 + nb * 8 + 2 * (l & 3) + j;
            if (ch < n_chan_here) bp[int64_t(row) * sh.n_chans + chan0 + ch] = acc[mb][nb][j];
          }
      }
    }
  }
  if (ctile == 0) {  // ---- harmonic sums: reduce the 64 position lanes (2 warps) of each class ----
#pragma unroll
    for (int j = 0; j < kGenH8; ++j) {
      const double c = warp_sum(sum_c[j]);
      const double sn = warp_sum(sum_s[j]);
      if (l == 0) {
        s_red[(2 * r8 + (warp & 1)) * 2 * kGenH8 + j] = c;
        s_red[(2 * r8 + (warp & 1)) * 2 * kGenH8 + kGenH8 + j] = sn;
      }
    }
    __syncthreads();
    double* tp = ws + sh.t_offset + cand * sh.t_stride_period + int64_t(split) * sh.t_stride_split;
    for (int e = tid; e < 8 * kGenH8; e += kTensorThreads) {
      const int r = e / kGenH8, j = e % kGenH8;
      const int m = r + 1 + 8 * j;
      if (m <= two_bw) {
        const double* wa = s_red + (2 * r) * 2 * kGenH8;
        const double* wb = s_red + (2 * r + 1) * 2 * kGenH8;
        tp[m - 1] = wa[j] + wb[j];
        tp[two_bw + m - 1] = wa[kGenH8 + j] + wb[kGenH8 + j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Narrow form of the accumulate kernel for one or two channels (the bundled single-channel
// example, cfg1, and the single-channel candidate sweeps, cfg5).  With so few right-hand sides
// the W'Y products are 2 C FMAs per harmonic -- not worth a tensor-core tile that is 64 channels
// wide -- so every thread is a generator: it walks the samples of its lane, carries its
// harmonics' sums AND their products with y in registers, and the CTA reduces once at the end.
// Same arithmetic and workspace layout as above (the K-half 1 partial is written as zeros).
constexpr int kNarrowThreads = 256;
constexpr int kNarrowBatch = 256;  // samples per sincos batch (one per thread)
constexpr int kNarrowRowsPerGroup = (PARRM_MAX_BANDWIDTH + kGenGroups - 1) / kGenGroups;  // 6

template <int C>
__global__ void __launch_bounds__(kNarrowThreads, 2)
eval_accumulate_narrow_kernel(const double* __restrict__ y, const int64_t* __restrict__ indices,
                              const double* __restrict__ periods, double* __restrict__ ws,
                              const EvalShape sh) {
  __shared__ double2 s_cs[kNarrowBatch];
  __shared__ double s_yv[kNarrowBatch * C];
  const int tid = threadIdx.x;
  const int64_t cand = blockIdx.x;
  const int split = blockIdx.y;
  const int bw = sh.bandwidth, two_bw = 2 * bw;
  const int64_t n_begin = int64_t(split) * sh.split_len;
  const int64_t n_end = min(n_begin + sh.split_len, sh.n_indices);
  const double delta = 6.283185307179586 / periods[cand];  // 2*pi/period (parrm.py:619)
  const int gi = tid & (kKT - 1), gg = tid / kKT;           // sample lane, harmonic residue group
  const int h = (two_bw - gg + kGenGroups - 1) / kGenGroups;  // harmonics gg+1, gg+5, ... <= 2 bw
  const int hb = (bw - gg + kGenGroups - 1) / kGenGroups;     // ... of which <= bw (rows of W)
  double sum_c[kGenH], sum_s[kGenH];
  double dot_c[kNarrowRowsPerGroup][C], dot_s[kNarrowRowsPerGroup][C], dot_1[C];
#pragma unroll
  for (int j = 0; j < kGenH; ++j) sum_c[j] = sum_s[j] = 0.0;
#pragma unroll
  for (int j = 0; j < kNarrowRowsPerGroup; ++j)
#pragma unroll
    for (int c = 0; c < C; ++c) dot_c[j][c] = dot_s[j][c] = 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) dot_1[c] = 0.0;

  for (int64_t n_batch = n_begin; n_batch < n_end; n_batch += kNarrowBatch) {
    __syncthreads();  // the previous batch has been consumed
    {
      const int64_t n = n_batch + tid;
      double2 cs = make_double2(0.0, 0.0);
      double yv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) yv[c] = 0.0;
      if (n < n_end) {
        const double angle = double(indices[n] + 1) * delta;
        sincos_phase(angle, &cs.y, &cs.x);
#pragma unroll
        for (int c = 0; c < C; ++c) yv[c] = y[n * sh.ld_y + c];
      }
      s_cs[tid] = cs;  // invalid samples: z = 0, y = 0 -> they add nothing below
#pragma unroll
      for (int c = 0; c < C; ++c) s_yv[tid * C + c] = yv[c];
    }
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < kNarrowBatch / kKT; ++sub) {
      if (n_batch + sub * kKT >= n_end) break;  // uniform
      const double2 z = s_cs[sub * kKT + gi];
      double yv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) yv[c] = s_yv[(sub * kKT + gi) * C + c];
      if (gg == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) dot_1[c] += yv[c];  // row 0 of W is the constant 1
      }
      double z2c = z.x, z2s = z.y;
      cmul(z2c, z2s, z.x, z.y);
      double z4c = z2c, z4s = z2s;
      cmul(z4c, z4s, z2c, z2s);
      double c = z.x, sn = z.y;  // seed z^(gg+1)
      if (gg == 1) {
        c = z2c; sn = z2s;
      } else if (gg == 2) {
        c = z2c; sn = z2s;
        cmul(c, sn, z.x, z.y);
      } else if (gg == 3) {
        c = z4c; sn = z4s;
      }
#pragma unroll
      for (int j = 0; j < kGenH; ++j) {
        if (j >= h) break;  // uniform within a warp
        if (j > 0) cmul(c, sn, z4c, z4s);
        sum_c[j] += c;
        sum_s[j] += sn;
        if (j < kNarrowRowsPerGroup && j < hb) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            dot_c[j][ch] = fma(c, yv[ch], dot_c[j][ch]);
            dot_s[j][ch] = fma(sn, yv[ch], dot_s[j][ch]);
          }
        }
      }
    }
  }

  // ---- reduce over the 64 sample lanes of every group (2 warps) and write ----
  __shared__ double s_red[kNarrowThreads / 32][2 * kGenH + (2 * kNarrowRowsPerGroup + 1) * C];
  const int warp = tid >> 5, lane = tid & 31;
  {
    int v = 0;
#pragma unroll
    for (int j = 0; j < kGenH; ++j) {
      const double a = warp_sum(sum_c[j]), b = warp_sum(sum_s[j]);
      if (lane == 0) {
        s_red[warp][v] = a;
        s_red[warp][v + 1] = b;
      }
      v += 2;
    }
#pragma unroll
    for (int j = 0; j < kNarrowRowsPerGroup; ++j)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const double a = warp_sum(dot_c[j][ch]), b = warp_sum(dot_s[j][ch]);
        if (lane == 0) {
          s_red[warp][v] = a;
          s_red[warp][v + 1] = b;
        }
        v += 2;
      }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      const double a = warp_sum(dot_1[ch]);
      if (lane == 0) s_red[warp][v] = a;
      ++v;
    }
  }
  __syncthreads();
  double* tp = ws + sh.t_offset + cand * sh.t_stride_period + int64_t(split) * sh.t_stride_split;
  double* bp0 = ws + cand * sh.b_stride_period + (int64_t(split) * 2) * sh.b_stride_split;
  double* bp1 = bp0 + sh.b_stride_split;
  // harmonic sums
  for (int e = tid; e < kGenGroups * kGenH; e += kNarrowThreads) {
    const int g = e / kGenH, j = e % kGenH;
    const int m = g + 1 + kGenGroups * j;
    if (m <= two_bw) {
      tp[m - 1] = s_red[2 * g][2 * j] + s_red[2 * g + 1][2 * j];
      tp[two_bw + m - 1] = s_red[2 * g][2 * j + 1] + s_red[2 * g + 1][2 * j + 1];
    }
  }
  // right-hand sides: rows 2m-1 (sin), 2m (cos), m = g + 1 + 4 j <= bw; row 0 from group 0
  for (int e = tid; e < kGenGroups * kNarrowRowsPerGroup * C; e += kNarrowThreads) {
    const int ch = e % C, j = (e / C) % kNarrowRowsPerGroup, g = e / (C * kNarrowRowsPerGroup);
    const int m = g + 1 + kGenGroups * j;
    if (m <= bw) {
      const int v = 2 * kGenH + (j * C + ch) * 2;
      bp0[int64_t(2 * m) * sh.n_chans + ch] = s_red[2 * g][v] + s_red[2 * g + 1][v];
      bp0[int64_t(2 * m - 1) * sh.n_chans + ch] = s_red[2 * g][v + 1] + s_red[2 * g + 1][v + 1];
      bp1[int64_t(2 * m) * sh.n_chans + ch] = 0.0;
      bp1[int64_t(2 * m - 1) * sh.n_chans + ch] = 0.0;
    }
  }
  if (tid < C) {
    const int v = 2 * kGenH + 2 * kNarrowRowsPerGroup * C + tid;
    bp0[tid] = s_red[0][v] + s_red[1][v];
    bp1[tid] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// Row 0 of W'Y: W's first column is the constant 1, so that row is sum_i y_i per channel -- the
// same for every candidate.  The tensor-core kernels therefore multiply only the 2 bw sine and
// cosine rows (40 = five 8-row blocks at bandwidth 20, instead of six for 41) and this small
// kernel computes the sums once per call, in a fixed order.
// Two stages so that the whole GPU reads Y (a 2-CTA version took 135 us per call, more than the
// rest of a five-candidate Nelder-Mead round): kColsumBlocks row blocks per 32-channel tile
// write partial sums, a second small kernel adds them in block order.
constexpr int kColsumBlocks = 64;

__global__ void __launch_bounds__(1024)
eval_colsum_kernel(const double* __restrict__ y, int64_t ld_y, int64_t n_indices, int64_t n_chans,
                   double* __restrict__ partial /* [kColsumBlocks][n_chans] */) {
  __shared__ double s_sum[32][33];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int64_t ch = int64_t(blockIdx.x) * 32 + cx;
  const int64_t rows = ceil_div(n_indices, kColsumBlocks);
  const int64_t n0 = int64_t(blockIdx.y) * rows, n1 = min64(n0 + rows, n_indices);
  double v = 0.0;
  if (ch < n_chans)
    for (int64_t n = n0 + ry; n < n1; n += 32) v += y[n * ld_y + ch];
  s_sum[ry][cx] = v;
  __syncthreads();
  if (ry == 0 && ch < n_chans) {
    double total = 0.0;
    for (int r = 0; r < 32; ++r) total += s_sum[r][cx];
    partial[int64_t(blockIdx.y) * n_chans + ch] = total;
  }
}

__global__ void __launch_bounds__(256)
eval_colsum_finish_kernel(const double* __restrict__ partial, int64_t n_chans,
                          double* __restrict__ colsum) {
  const int64_t ch = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (ch >= n_chans) return;
  double t[kColsumBlocks];
#pragma unroll
  for (int b = 0; b < kColsumBlocks; ++b) t[b] = partial[int64_t(b) * n_chans + ch];
  double total = 0.0;
#pragma unroll
  for (int b = 0; b < kColsumBlocks; ++b) total += t[b];
  colsum[ch] = total;
}

// ------------------------------------------------------------------------------------------
// Sample splits (few candidates, e.g. a Nelder-Mead round) leave 2 * n_splits partials of W'Y
// per candidate.  Added up inside the solve kernel they are ~70 dependent L2 round trips on
// one CTA per candidate (60k of its 235k cycles, measured); here every (candidate, row) gets
// its own CTA, so the sums cost one short wave.  Fixed order; the total lands in partial 0.
__global__ void __launch_bounds__(64)
eval_reduce_partials_kernel(double* __restrict__ ws, const EvalShape sh) {
  const int64_t cand = blockIdx.x / sh.n_rows;
  const int m = int(blockIdx.x % sh.n_rows);
  if (m == 0 && sh.row0_from_colsum) return;  // that row comes from the column sums
  double* pm = ws + cand * sh.b_stride_period + int64_t(m) * sh.n_chans;
  const int n_part = 2 * sh.n_splits;
  for (int64_t ch = threadIdx.x; ch < sh.n_chans; ch += 64) {
    double v = 0.0;
    int sp = 0;
    for (; sp + 32 <= n_part; sp += 32) {  // every load is an L2 round trip: many in flight,
      double t[32];                        // the adds in the fixed order all the same
#pragma unroll
      for (int u = 0; u < 32; ++u) t[u] = pm[(sp + u) * sh.b_stride_split + ch];
#pragma unroll
      for (int u = 0; u < 32; ++u) v += t[u];
    }
    for (; sp + 8 <= n_part; sp += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = pm[(sp + u) * sh.b_stride_split + ch];
#pragma unroll
      for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; sp < n_part; ++sp) v += pm[sp * sh.b_stride_split + ch];
    pm[ch] = v;
  }
}

// ------------------------------------------------------------------------------------------
// 1/x for the pivots of the solve: hardware seed (MUFU.RCP64H, ~20 bits) and two Newton steps,
// straight-line code that the scheduler can interleave with the trailing update -- the IEEE
// division is a ~400-cycle dependent sequence with a call for the special cases.  Within an
// ulp of the rounded reciprocal for normal x; 0 -> NaN/inf and non-finite x -> NaN, which
// only happen where the factorisation is reported singular or is NaN anyway.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

#ifdef PARRM_SOLVE_TIMING
// Debug build only: cycles thread 0 of CTA 0 spends in each phase of the solve kernel.
__device__ unsigned long long g_solve_timing[8];
#define SOLVE_TICK(slot)                                                    \
  do {                                                                      \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                              \
      const long long now__ = clock64();                                    \
      g_solve_timing[slot] += now__ - stick__;                              \
      stick__ = now__;                                                      \
    }                                                                       \
  } while (0)
#else
#define SOLVE_TICK(slot)
#endif

constexpr int kSolveChans = 64;     // channels per pass: one per thread in the substitutions
constexpr int kSolveThreads = 256;  // eight warps: rows over warps in the factorisation, four
                                    // row roles per channel in the substitutions.  80 registers
                                    // and 74 KB of shared memory: three CTAs per SM
constexpr int kGStride = kMaxRows + 2;  // row stride of the Gram matrix (odd: walking down a
                                        // column touches every bank once)

__global__ void __launch_bounds__(kSolveThreads, 3)
eval_solve_kernel(const double* __restrict__ ws, const double* __restrict__ sumsq, double lambda,
                  int64_t n_chans_divisor, double* __restrict__ fit_error, const EvalShape sh) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = sh.n_rows, bw = sh.bandwidth, two_bw = 2 * bw;
  double* s_g = reinterpret_cast<double*>(smem_raw);   // [M][kGStride] LU factors
  double* s_g0 = s_g + M * kGStride;          // [M][kGStride] Gram matrix, kept unfactored
  double* s_hc = s_g0 + M * kGStride;         // C_0..C_2bw
  double* s_hs = s_hc + kMaxRows;             // S_0..S_2bw
  double* s_b = s_hs + kMaxRows;              // [M][kSolveChans] right-hand sides
  double* s_x = s_b + M * kSolveChans;        // [M][kSolveChans] work / solution
  __shared__ int s_perm[kMaxRows];
  __shared__ double s_inv[kMaxRows];  // reciprocals of the pivots
  __shared__ unsigned long long s_cand[2 * (kSolveThreads / 32)];  // pivot candidates per warp,
                                                                   // double-buffered
  __shared__ int s_singular;
  __shared__ double s_part[kSolveThreads / 32];

  const int tid = threadIdx.x;
  const int64_t cand = blockIdx.x;
#ifdef PARRM_SOLVE_TIMING
  long long stick__ = clock64();
#endif

  // harmonic sums, splits added in a fixed order
  for (int e = tid; e <= two_bw; e += kSolveThreads) {
    double c = 0.0, s = 0.0;
    if (e == 0) {
      c = double(sh.n_indices);
    } else {
      const double* tp = ws + sh.t_offset + cand * sh.t_stride_period;
      int sp = 0;
      for (; sp + 16 <= sh.n_splits; sp += 16) {  // fixed order, loads ahead of the adds
        double tc[16], ts[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          tc[u] = tp[(sp + u) * sh.t_stride_split + e - 1];
          ts[u] = tp[(sp + u) * sh.t_stride_split + two_bw + e - 1];
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          c += tc[u];
          s += ts[u];
        }
      }
      for (; sp + 4 <= sh.n_splits; sp += 4) {
        double tc[4], ts[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          tc[u] = tp[(sp + u) * sh.t_stride_split + e - 1];
          ts[u] = tp[(sp + u) * sh.t_stride_split + two_bw + e - 1];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          c += tc[u];
          s += ts[u];
        }
      }
      for (; sp < sh.n_splits; ++sp) {
        c += tp[sp * sh.t_stride_split + e - 1];
        s += tp[sp * sh.t_stride_split + two_bw + e - 1];
      }
    }
    s_hc[e] = c;
    s_hs[e] = s;
  }
  if (tid == 0) s_singular = 0;
  __syncthreads();
  // Gram matrix by product-to-sum: column 0 = 1, 2k-1 = sin(k a), 2k = cos(k a)
  for (int e = tid; e < M * M; e += kSolveThreads) {
    const int i = e / M, j = e % M;
    const int ki = (i + 1) / 2, kj = (j + 1) / 2;      // harmonic numbers (0 for the constant)
    const bool si = (i & 1), sj = (j & 1);             // odd column = sine
    const int kd = ki > kj ? ki - kj : kj - ki, ksum = ki + kj;
    double v;
    if (i == 0 && j == 0) {
      v = s_hc[0];
    } else if (i == 0 || j == 0) {
      const int k = ki + kj;
      v = (si || sj) ? s_hs[k] : s_hc[k];
    } else if (si && sj) {
      v = 0.5 * (s_hc[kd] - s_hc[ksum]);
    } else if (!si && !sj) {
      v = 0.5 * (s_hc[kd] + s_hc[ksum]);
    } else {
      // sin(ks a) cos(kc a) = 0.5 * (sin((ks+kc) a) + sin((ks-kc) a))
      const int ks = si ? ki : kj, kc = si ? kj : ki;
      const double sd = ks >= kc ? s_hs[ks - kc] : -s_hs[kc - ks];
      v = 0.5 * (s_hs[ksum] + sd);
    }
    s_g[i * kGStride + j] = v;
    s_g0[i * kGStride + j] = v;
  }
  __syncthreads();

  SOLVE_TICK(0);
  // LU with partial pivoting, in place, rows left where they are: dgetrf's interchanges are
  // recorded in s_perm and never carried out.  The arithmetic is that of the row-swapping form,
  // operation for operation, but a step costs ONE block barrier -- there is no swap, and the
  // search for the next pivot (with the reciprocal it will need) rides on the update that
  // produces the column it looks at.  Warp w owns rows w, w+8, ... and keeps them in
  // registers (lane l: columns l and l+32), storing every update through to shared memory,
  // where the other warps read the pivot row and the substitutions read the factors.  The
  // lane that holds column k+1 keeps the largest |entry| of the warp's open rows, compared
  // through its bit pattern (for non-negative doubles the unsigned order is the numeric order,
  // a NaN sorts above infinity so it wins like a propagating max, and integer compares do not
  // wait on the FP64 pipe).  After the barrier every thread reduces the eight candidates for
  // itself.  Equal magnitudes go to the lower row index.  The candidate cells are
  // double-buffered by the parity of k: a fast warp writes those of step k+1 while a slow one
  // may still be reading those of step k.
  constexpr int kWarps = kSolveThreads / 32;
  constexpr int kSlots = (kMaxRows + kWarps - 1) / kWarps;  // rows per warp
  static_assert(kMaxRows <= 64, "two column registers per lane");
  const int warp = tid >> 5, ln = tid & 31;
  unsigned long long used = 0ull;  // rows already taken as pivots (the same in every thread)
  double g0[kSlots], g1[kSlots];
#pragma unroll
  for (int u = 0; u < kSlots; ++u) {
    const int i = warp + kWarps * u;
    g0[u] = (i < M && ln < M) ? s_g[i * kGStride + ln] : 0.0;
    g1[u] = (i < M && ln + 32 < M) ? s_g[i * kGStride + ln + 32] : 0.0;
  }
  // A candidate is one 64-bit key: the bit pattern of |entry| with its six lowest mantissa
  // bits replaced by 63 - row, so the arg-max is a plain unsigned max (equal magnitudes -- to
  // within 64 ulp -- go to the lower row; a key below 64 is an exact zero pivot; 0 = no row).
  auto make_key = [](double v, int row) {
    return (static_cast<unsigned long long>(__double_as_longlong(v)) & 0x7fffffffffffffc0ull) |
           static_cast<unsigned long long>(63 - row);
  };
  auto umax = [](unsigned long long a, unsigned long long b) { return a > b ? a : b; };
  {
    // candidates of column 0: lane 0 holds it for every row of the warp
    unsigned long long key = 0ull;
#pragma unroll
    for (int u = 0; u < kSlots; ++u) {
      const int i = warp + kWarps * u;
      key = umax(key, i < M ? make_key(g0[u], i) : 0ull);
    }
    if (ln == 0) s_cand[warp] = key;
  }
  __syncthreads();
  static_assert(kWarps == 8, "the candidate reduction below shuffles over eight lanes");
  // One elimination step; HI = column k lives in the second register of its lane (k >= 32).
  // Straight-line code (selects and predicated stores), so the six rows of a warp overlap.
  // Registers of columns <= k are dead once their multiplier is stored: they keep computing
  // (garbage) and are never stored or looked at again.
  auto lu_step = [&](int k, auto hi_tag) {
    constexpr bool HI = decltype(hi_tag)::value;
    const int buf = (k & 1) * kWarps;
    unsigned long long key = ln < kWarps ? s_cand[buf + ln] : 0ull;
#pragma unroll
    for (int o = kWarps / 2; o > 0; o >>= 1) key = umax(key, __shfl_xor_sync(0xffffffffu, key, o));
    key = __shfl_sync(0xffffffffu, key, 0);
    const int piv = 63 - int(key & 63ull);
    used |= 1ull << piv;
    const double* prow = s_g + piv * kGStride;
    const double inv_p = fast_rcp(prow[k]);
    const double p0 = (!HI && ln < M) ? prow[ln] : 0.0, p1 = ln + 32 < M ? prow[ln + 32] : 0.0;
    if (tid == 0) {
      s_perm[k] = piv;
      s_inv[k] = inv_p;  // reused by the back substitution
      if ((key >> 6) == 0ull) s_singular = 1;
    }
    const int kl = k & 31;
    const bool nhi = k + 1 >= 32;  // which register holds column k+1
    double rk[kSlots];
#pragma unroll
    for (int u = 0; u < kSlots; ++u) rk[u] = __shfl_sync(0xffffffffu, HI ? g1[u] : g0[u], kl);
    unsigned long long nkey = 0ull;
#pragma unroll
    for (int u = 0; u < kSlots; ++u) {
      const int i = warp + kWarps * u;
      const bool act = i < M && !((used >> i) & 1ull);
      const double lik = rk[u] * inv_p;
      if (!HI) {
        g0[u] = fma(-lik, p0, g0[u]);
        g1[u] = fma(-lik, p1, g1[u]);
        if (act && ln >= kl && ln < M) s_g[i * kGStride + ln] = ln == kl ? lik : g0[u];
        if (act && ln + 32 < M) s_g[i * kGStride + ln + 32] = g1[u];
      } else {
        g1[u] = fma(-lik, p1, g1[u]);
        if (act && ln >= kl && ln + 32 < M) s_g[i * kGStride + ln + 32] = ln == kl ? lik : g1[u];
      }
      nkey = umax(nkey, act ? make_key(nhi ? g1[u] : g0[u], i) : 0ull);
    }
    // this lane's column is the one the next step searches
    if (ln == ((k + 1) & 31)) s_cand[(kWarps - buf) + warp] = nkey;
    __syncthreads();
  };
  {
    const int m_lo = M < 32 ? M : 32;
    for (int k = 0; k < m_lo; ++k) lu_step(k, std::false_type{});
    for (int k = 32; k < M; ++k) lu_step(k, std::true_type{});
  }
  SOLVE_TICK(1);

  // per-channel solve; channels in passes of kSolveChans.  Thread (lane, role): lane = channel
  // of the pass, the four roles split the rows and keep theirs in registers.  Both
  // substitutions are column-oriented: once an unknown is final it is eliminated from every
  // row still open, all rows and channels at once -- one independent FMA per (row, channel)
  // and one barrier per unknown, instead of a chain of M dependent dot products per channel.
  // Updates are stored through to s_x, where the next step finds its unknown.
  const double tri = 0.5 * double(M) * double(M + 1);
  const int lane = tid % kSolveChans, role = tid / kSolveChans;
  constexpr int kRoles = kSolveThreads / kSolveChans;
  constexpr int kRowSlots = (kMaxRows + kRoles - 1) / kRoles;
  double local = 0.0;
  for (int64_t c0 = 0; c0 < sh.n_chans; c0 += kSolveChans) {
    const int64_t ch = c0 + lane;
    const bool live = ch < sh.n_chans;
    double xr[kRowSlots];
    {
      // right-hand sides: partials in a fixed order, the loads of all rows of a thread in
      // flight together (each is an L2 round trip)
      const double* bp = ws + cand * sh.b_stride_period + ch;
      const int n_part = sh.b_reduced ? 1 : 2 * sh.n_splits;
      const bool colsum0 = role == 0 && sh.row0_from_colsum;  // slot 0 of role 0 is row 0
#pragma unroll
      for (int u = 0; u < kRowSlots; ++u) xr[u] = 0.0;
      for (int sp = 0; sp < n_part; ++sp) {
        double t[kRowSlots];
#pragma unroll
        for (int u = 0; u < kRowSlots; ++u) {
          const int m = role + kRoles * u;
          const bool get = live && m < M && !(u == 0 && colsum0);
          t[u] = get ? bp[int64_t(m) * sh.n_chans + int64_t(sp) * sh.b_stride_split] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kRowSlots; ++u) xr[u] += t[u];
      }
      // the constant column of W: the same sum of y for every candidate
      if (colsum0 && live) xr[0] = ws[sh.c_offset + ch];
#pragma unroll
      for (int u = 0; u < kRowSlots; ++u) {
        const int m = role + kRoles * u;
        if (m < M) {
          s_b[m * kSolveChans + lane] = xr[u];
          s_x[m * kSolveChans + lane] = xr[u];
        }
      }
    }
    __syncthreads();
    SOLVE_TICK(2);
    double* xc = s_x + lane;
    // Both substitutions run in blocks of kBlk unknowns: every thread first resolves the
    // block's own small triangle for its channel (redundantly over the four roles; the same
    // FMA chains, in the same order, as the one-unknown-at-a-time form), then each row still
    // open loses all kBlk terms at once -- two barriers per block instead of one per unknown,
    // and a quarter of the loads, stores and predicate tests.
    constexpr int kBlk = 4;
    auto own_bit = [&](int row) { return (row & (kRoles - 1)) == role ? (1u << (row / kRoles)) : 0u; };
    unsigned all_rows = 0u;
#pragma unroll
    for (int u = 0; u < kRowSlots; ++u) all_rows |= (role + kRoles * u < M) ? (1u << u) : 0u;
    // L (unit lower): after step k the rows not yet taken as pivots lose l_ik * x[pivot k]
    unsigned cur = all_rows;  // this thread's rows not yet taken as pivots
    for (int k0 = 0; k0 + 1 < M; k0 += kBlk) {
      const int nb = min(kBlk, M - 1 - k0);
      int pv[kBlk];
      double z[kBlk];
      unsigned msk[kBlk];
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        pv[j] = j < nb ? s_perm[k0 + j] : 0;
        cur &= ~(j < nb ? own_bit(pv[j]) : 0u);
        msk[j] = j < nb ? cur : 0u;
      }
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        double v = 0.0;
        if (j < nb) {
          v = xc[pv[j] * kSolveChans];
#pragma unroll
          for (int t = 0; t < j; ++t) v = fma(-s_g[pv[j] * kGStride + k0 + t], z[t], v);
        }
        z[j] = v;
      }
      __syncthreads();  // every read of the block's pivot rows precedes their update below
#pragma unroll
      for (int u = 0; u < kRowSlots; ++u) {
        const int i = role + kRoles * u;
        if ((msk[0] >> u) & 1u) {
#pragma unroll
          for (int j = 0; j < kBlk; ++j)
            if ((msk[j] >> u) & 1u) xr[u] = fma(-s_g[i * kGStride + k0 + j], z[j], xr[u]);
          xc[i * kSolveChans] = xr[u];
        }
      }
      __syncthreads();
    }
    SOLVE_TICK(3);
    // U: unknown k is final in its pivot row (scaled by 1/u_kk when it is read); the pivot
    // rows of the earlier steps lose u_jk * beta_k
    cur = all_rows;  // this thread's rows whose unknown is not final yet
    for (int k0 = M - 1; k0 > 0; k0 -= kBlk) {
      const int nb = min(kBlk, k0);
      int pv[kBlk];
      double bt[kBlk];
      unsigned msk[kBlk];
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        pv[j] = j < nb ? s_perm[k0 - j] : 0;
        cur &= ~(j < nb ? own_bit(pv[j]) : 0u);
        msk[j] = j < nb ? cur : 0u;
      }
#pragma unroll
      for (int j = 0; j < kBlk; ++j) {
        double v = 0.0;
        if (j < nb) {
          v = xc[pv[j] * kSolveChans];
#pragma unroll
          for (int t = 0; t < j; ++t) v = fma(-s_g[pv[j] * kGStride + k0 - t], bt[t], v);
          v *= s_inv[k0 - j];
        }
        bt[j] = v;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kRowSlots; ++u) {
        const int i = role + kRoles * u;
        if ((msk[0] >> u) & 1u) {
#pragma unroll
          for (int j = 0; j < kBlk; ++j)
            if ((msk[j] >> u) & 1u) xr[u] = fma(-s_g[i * kGStride + k0 - j], bt[j], xr[u]);
          xc[i * kSolveChans] = xr[u];
        }
      }
      __syncthreads();
    }
    SOLVE_TICK(4);
    // beta into the natural order of the unknowns (through registers: s_x is both ends)
    {
      double beta_r[(kMaxRows + kRoles - 1) / kRoles];
#pragma unroll
      for (int u = 0; u < (kMaxRows + kRoles - 1) / kRoles; ++u) {
        const int m = role + u * kRoles;
        beta_r[u] = m < M ? xc[s_perm[m] * kSolveChans] * s_inv[m] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < (kMaxRows + kRoles - 1) / kRoles; ++u) {
        const int m = role + u * kRoles;
        if (m < M) xc[m * kSolveChans] = beta_r[u];
      }
      __syncthreads();
    }
    {
      // sum_i (y - W beta)^2 = y'y - 2 beta'b + beta'G beta holds for ANY beta, so like the
      // reference's explicit residual (parrm.py:630) it is only second-order sensitive to the
      // rounding of the solve; y'y - beta'b would be first-order sensitive.  Two rows of G at
      // a time with four partial sums each: a single FMA chain is bound by the FP64 latency.
      // G is symmetric: beta'G beta = sum_m beta_m (G_mm beta_m + 2 sum_{j<m} G_mj beta_j).
      double cross = 0.0, quad = 0.0, penalty = 0.0;
      for (int m = role; m < M; m += 2 * kRoles) {
        const bool two = m + kRoles < M;
        const int m2 = two ? m + kRoles : m;
        const double* ga = s_g0 + m * kGStride;
        const double* gb = s_g0 + m2 * kGStride;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int j = 0;
        for (; j + 4 <= m; j += 4) {  // columns below both diagonals
          const double x0 = xc[j * kSolveChans], x1 = xc[(j + 1) * kSolveChans];
          const double x2 = xc[(j + 2) * kSolveChans], x3 = xc[(j + 3) * kSolveChans];
          a0 = fma(ga[j], x0, a0);
          b0 = fma(gb[j], x0, b0);
          a1 = fma(ga[j + 1], x1, a1);
          b1 = fma(gb[j + 1], x1, b1);
          a2 = fma(ga[j + 2], x2, a2);
          b2 = fma(gb[j + 2], x2, b2);
          a3 = fma(ga[j + 3], x3, a3);
          b3 = fma(gb[j + 3], x3, b3);
        }
        for (; j < m; ++j) {
          const double x0 = xc[j * kSolveChans];
          a0 = fma(ga[j], x0, a0);
          b0 = fma(gb[j], x0, b0);
        }
        for (; j < m2; ++j) b1 = fma(gb[j], xc[j * kSolveChans], b1);  // at most kRoles more
        const double beta = xc[m * kSolveChans];
        cross = fma(beta, s_b[m * kSolveChans + lane], cross);
        quad = fma(beta, fma(ga[m], beta, 2.0 * ((a0 + a1) + (a2 + a3))), quad);
        penalty = fma((lambda * double(m + 1)) / tri, beta * beta, penalty);
        if (two) {
          const double beta2 = xc[m2 * kSolveChans];
          cross = fma(beta2, s_b[m2 * kSolveChans + lane], cross);
          quad = fma(beta2, fma(gb[m2], beta2, 2.0 * ((b0 + b1) + (b2 + b3))), quad);
          penalty = fma((lambda * double(m2 + 1)) / tri, beta2 * beta2, penalty);
        }
      }
      if (live)
        local += ((role == 0 ? sumsq[ch] : 0.0) - 2.0 * cross + quad) / double(sh.n_indices) + penalty;
    }
    __syncthreads();  // the next pass overwrites s_b / s_x
    SOLVE_TICK(5);
  }
  local = warp_sum(local);
  if ((tid & 31) == 0) s_part[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    double total = 0.0;
    for (int i = 0; i < kSolveThreads / 32; ++i) total += s_part[i];
    total /= double(n_chans_divisor);
    fit_error[cand] = s_singular ? __longlong_as_double(0x7ff0000000000000LL) : total;
  }
}

// Dense, zero-padded copy of Y per 64-channel tile, [tile][n][64], for the tensor kernel's bulk
// copies when the caller's array is not already one dense, aligned, even-width tile (more than
// 64 channels, odd widths, padded rows).  Once per call: 2 x the bytes of Y.
__global__ void __launch_bounds__(256)
eval_retile_kernel(const double* __restrict__ y, int64_t ld_y, int64_t n_indices, int64_t n_chans,
                   int n_chan_tiles, double* __restrict__ out) {
  const int64_t total = int64_t(n_chan_tiles) * n_indices * kChanTile;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int c = int(e % kChanTile);
    const int64_t row = (e / kChanTile) % n_indices;
    const int64_t ch = (e / kChanTile / n_indices) * kChanTile + c;
    out[e] = ch < n_chans ? y[row * ld_y + ch] : 0.0;
  }
}

// float32 storage of the search tile: widened once per call into a dense float64 [N, C] array
// in the workspace, which the float64 kernels then read (the arithmetic stays float64).
__global__ void __launch_bounds__(256)
eval_widen_kernel(const float* __restrict__ y, int64_t ld_y, int64_t n_indices, int64_t n_chans,
                  double* __restrict__ out) {
  const int64_t total = n_indices * n_chans;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256)
    out[e] = double(y[(e / n_chans) * ld_y + (e % n_chans)]);
}

// first index of the smallest non-NaN value
__global__ void __launch_bounds__(1024)
argmin_kernel(const double* __restrict__ v, int64_t n, double* min_value, int64_t* min_index) {
  __shared__ double s_v[32];
  __shared__ int64_t s_i[32];
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double best = inf;
  int64_t best_i = INT64_MAX;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const double x = v[i];
    if (x < best || (x == best && i < best_i)) {
      best = x;
      best_i = i;
    }
  }
  auto combine = [&](double ov, int64_t oi) {
    if (ov < best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    combine(ov, oi);
  }
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = best_i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    best = s_v[0];
    best_i = s_i[0];
    for (int w = 1; w < 32; ++w) combine(s_v[w], s_i[w]);
    *min_value = best;
    *min_index = best_i == INT64_MAX ? 0 : best_i;
  }
}

static int make_shape(int64_t n_chans, int64_t n_indices, int64_t n_periods, int bandwidth,
                      int64_t ld_y, EvalShape* sh) {
  sh->n_chans = n_chans;
  sh->n_indices = n_indices;
  sh->n_periods = n_periods;
  sh->bandwidth = bandwidth;
  sh->n_rows = 2 * bandwidth + 1;
  sh->n_chan_tiles = int(ceil_div(n_chans, kChanTile));
  // Few candidates (Nelder-Mead rounds): split the samples so that every SM has work -- one
  // CTA per SM for the tensor kernel, ~4 for the 256-thread kernels.  Every split costs the
  // solve kernel two more partials to add per row, so no more than that; a split is a whole
  // number of sincos batches.
  const int64_t per_sm = n_chans > 2 ? 1 : 4;
  const int64_t want = ceil_div(per_sm * kNumSMs, n_periods * sh->n_chan_tiles);
  const int64_t max_splits = ceil_div(n_indices, kSplitQuantum);
  int64_t n_splits = max64(1, min(want, max_splits));
  sh->split_len = ceil_div(ceil_div(n_indices, n_splits), kSplitQuantum) * kSplitQuantum;
  sh->n_splits = int(ceil_div(n_indices, sh->split_len));
  sh->ld_y = ld_y;
  sh->b_stride_split = int64_t(sh->n_rows) * n_chans;
  sh->b_stride_period = 2 * sh->n_splits * sh->b_stride_split;
  sh->t_stride_split = 4 * int64_t(bandwidth);
  sh->t_stride_period = sh->n_splits * sh->t_stride_split;
  sh->t_offset = n_periods * sh->b_stride_period;
  sh->c_offset = sh->t_offset + n_periods * sh->t_stride_period;
  sh->row0_from_colsum = 0;
  sh->b_reduced = 0;
  // c_offset: n_chans column sums, then kColsumBlocks x n_chans partials
  sh->y_offset = (sh->c_offset + (1 + kColsumBlocks) * n_chans + 1) / 2 * 2;  // 16-byte aligned
  sh->y_row_chans = 0;
  sh->y_tile_stride = 0;
  return PARRM_OK;
}

}  // namespace parrm

extern "C" {
#ifdef PARRM_SOLVE_TIMING
int parrm_debug_solve_timing(unsigned long long* h_out, int reset) {
  if (h_out) cudaMemcpyFromSymbol(h_out, parrm::g_solve_timing, sizeof(parrm::g_solve_timing));
  if (reset) {
    unsigned long long zero[8] = {0};
    cudaMemcpyToSymbol(parrm::g_solve_timing, zero, sizeof(zero));
  }
  return 0;
}
#endif
#ifdef PARRM_TENSOR_TIMING
int parrm_debug_tensor_timing(unsigned long long* h_out, int reset) {
  if (h_out) cudaMemcpyFromSymbol(h_out, parrm::g_tensor_timing, sizeof(parrm::g_tensor_timing));
  if (reset) {
    unsigned long long zero[128] = {0};
    cudaMemcpyToSymbol(parrm::g_tensor_timing, zero, sizeof(zero));
  }
  return 0;
}
#endif


size_t parrm_eval_workspace_bytes(int64_t n_chans, int64_t n_indices, int64_t n_periods,
                                  int bandwidth) {
  if (n_chans <= 0 || n_indices <= 0 || n_periods <= 0 || bandwidth < 0) return 0;
  parrm::EvalShape sh;
  parrm::make_shape(n_chans, n_indices, n_periods, bandwidth, n_chans, &sh);
  // the re-tiled copy of Y is only needed by the tensor kernel (more than two channels)
  const int64_t retile = n_chans > 2 ? int64_t(sh.n_chan_tiles) * n_indices * parrm::kChanTile : 0;
  return size_t(sh.y_offset + retile + 2) * sizeof(double);
}

static bool y_is_one_dense_tile(const double* d_y, int64_t ld_y, int64_t n_chans) {
  return n_chans <= parrm::kChanTile && ld_y == n_chans && n_chans % 2 == 0 &&
         (reinterpret_cast<uintptr_t>(d_y) & 15) == 0;
}

int parrm_eval_launch_count(const double* d_y, int64_t ld_y, int64_t n_chans, int64_t n_indices,
                            int64_t n_periods, int bandwidth) {
  if (n_chans <= 0 || n_indices <= 0 || n_periods <= 0 || bandwidth < 0 ||
      bandwidth > PARRM_MAX_BANDWIDTH)
    return 0;
  parrm::EvalShape sh;
  parrm::make_shape(n_chans, n_indices, n_periods, bandwidth, ld_y, &sh);
  int n = 2;  // accumulate + solve
  if (n_chans > 2) n += 2 + (y_is_one_dense_tile(d_y, ld_y, n_chans) ? 0 : 1);
  if (sh.n_splits > 1) n += 1;
  return n;
}

int parrm_eval_periods(const double* d_y, int64_t ld_y, const double* d_sumsq,
                       const int64_t* d_indices, int64_t n_chans, int64_t n_indices,
                       const double* d_periods, int64_t n_periods, int bandwidth, double lambda,
                       int64_t n_chans_divisor, double* d_fit_error, void* d_workspace,
                       size_t workspace_bytes, void* stream) {
  PARRM_NVTX("parrm_eval_periods");
  using namespace parrm;
  PARRM_REQUIRE(n_chans > 0 && n_indices > 0 && n_periods >= 0 && ld_y >= n_chans,
                "parrm_eval_periods: bad shape");
  PARRM_REQUIRE(n_chans_divisor > 0, "parrm_eval_periods: n_chans_divisor must be > 0");
  if (bandwidth < 0 || bandwidth > PARRM_MAX_BANDWIDTH) {
    set_error("parrm_eval_periods: bandwidth %d outside [0, %d]", bandwidth, PARRM_MAX_BANDWIDTH);
    return PARRM_ERR_UNSUPPORTED;
  }
  if (n_periods == 0) return PARRM_OK;
  PARRM_REQUIRE(d_y && d_sumsq && d_indices && d_periods && d_fit_error && d_workspace,
                "parrm_eval_periods: null pointer");
  PARRM_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0,
                "parrm_eval_periods: workspace must be 16-byte aligned");
  if (workspace_bytes < parrm_eval_workspace_bytes(n_chans, n_indices, n_periods, bandwidth)) {
    set_error("parrm_eval_periods: workspace too small");
    return PARRM_ERR_WORKSPACE;
  }
  EvalShape sh;
  make_shape(n_chans, n_indices, n_periods, bandwidth, ld_y, &sh);
  PARRM_REQUIRE(sh.n_splits <= 65535 && sh.n_chan_tiles <= 65535,
                "parrm_eval_periods: grid too large");
  cudaStream_t s = as_stream(stream);
  double* ws = static_cast<double*>(d_workspace);
  dim3 grid((unsigned)n_periods, (unsigned)sh.n_splits, (unsigned)sh.n_chan_tiles);
  if (n_chans <= 2) {
    dim3 narrow_grid((unsigned)n_periods, (unsigned)sh.n_splits, 1);
    if (n_chans == 1)
      eval_accumulate_narrow_kernel<1><<<narrow_grid, kNarrowThreads, 0, s>>>(d_y, d_indices,
                                                                              d_periods, ws, sh);
    else
      eval_accumulate_narrow_kernel<2><<<narrow_grid, kNarrowThreads, 0, s>>>(d_y, d_indices,
                                                                              d_periods, ws, sh);
  } else {
    sh.row0_from_colsum = 1;
    eval_colsum_kernel<<<dim3(unsigned(ceil_div(n_chans, 32)), kColsumBlocks), dim3(32, 32), 0, s>>>(
        d_y, ld_y, n_indices, n_chans, ws + sh.c_offset + n_chans);
    eval_colsum_finish_kernel<<<unsigned(ceil_div(n_chans, 256)), 256, 0, s>>>(
        ws + sh.c_offset + n_chans, n_chans, ws + sh.c_offset);
    const size_t smem =
        size_t(4 * kTensorBatch * 16 + (kTensorWtTile + 2 * kTensorYTile + 16 * 2 * kGenH8) * sizeof(double));
    const double* y_src = d_y;
    if (y_is_one_dense_tile(d_y, ld_y, n_chans)) {
      sh.y_row_chans = int(n_chans);  // the caller's array is one dense tile already
      sh.y_tile_stride = 0;
    } else {
      double* tiles = ws + sh.y_offset;
      const int64_t total = int64_t(sh.n_chan_tiles) * n_indices * kChanTile;
      eval_retile_kernel<<<unsigned(min64(ceil_div(total, 256), 8 * kNumSMs)), 256, 0, s>>>(
          d_y, ld_y, n_indices, n_chans, sh.n_chan_tiles, tiles);
      y_src = tiles;
      sh.y_row_chans = kChanTile;
      sh.y_tile_stride = n_indices * kChanTile;
    }
    void (*tensor)(const double*, const int64_t*, const double*, double*, const EvalShape) = nullptr;
    switch ((sh.n_rows - 1 + 7) / 8) {  // 8-row blocks of W' without its constant row
      case 0: case 1: tensor = eval_accumulate_tensor_kernel<1>; break;
      case 2: tensor = eval_accumulate_tensor_kernel<2>; break;
      case 3: tensor = eval_accumulate_tensor_kernel<3>; break;
      case 4: tensor = eval_accumulate_tensor_kernel<4>; break;
      case 5: tensor = eval_accumulate_tensor_kernel<5>; break;
      default: tensor = eval_accumulate_tensor_kernel<6>; break;
    }
    PARRM_CUDA_OK(cudaFuncSetAttribute(tensor, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    tensor<<<grid, kTensorThreads, smem, s>>>(y_src, d_indices, d_periods, ws, sh);
  }
  PARRM_LAUNCH_OK("eval_accumulate_kernel");
  if (sh.n_splits > 1) {
    eval_reduce_partials_kernel<<<unsigned(n_periods * sh.n_rows), 64, 0, s>>>(ws, sh);
    sh.b_reduced = 1;
  }
  const size_t solve_smem =
      size_t(sh.n_rows * (2 * kGStride + 2 * kSolveChans) + 2 * kMaxRows) * sizeof(double);
  PARRM_CUDA_OK(cudaFuncSetAttribute(eval_solve_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(solve_smem)));
  eval_solve_kernel<<<(unsigned)n_periods, kSolveThreads, solve_smem, s>>>(
      ws, d_sumsq, lambda, n_chans_divisor, d_fit_error, sh);
  PARRM_LAUNCH_OK("eval_solve_kernel");
  return PARRM_OK;
}

// ---- float32 storage of the tile (the "fp32" mode of the search) ---------------------------
static size_t eval_base_bytes(int64_t n_chans, int64_t n_indices, int64_t n_periods, int bandwidth) {
  return (parrm_eval_workspace_bytes(n_chans, n_indices, n_periods, bandwidth) + 15) & ~size_t(15);
}

size_t parrm_eval_workspace_bytes_typed(int64_t n_chans, int64_t n_indices, int64_t n_periods,
                                        int bandwidth, int y_dtype) {
  const size_t base = parrm_eval_workspace_bytes(n_chans, n_indices, n_periods, bandwidth);
  if (y_dtype != PARRM_F32 || base == 0) return base;
  return eval_base_bytes(n_chans, n_indices, n_periods, bandwidth) +
         size_t(n_chans) * size_t(n_indices) * sizeof(double);
}

int parrm_eval_periods_typed(const void* d_y, int y_dtype, int64_t ld_y, const double* d_sumsq,
                             const int64_t* d_indices, int64_t n_chans, int64_t n_indices,
                             const double* d_periods, int64_t n_periods, int bandwidth,
                             double lambda, int64_t n_chans_divisor, double* d_fit_error,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
  PARRM_NVTX("parrm_eval_periods_typed");
  using namespace parrm;
  if (y_dtype == PARRM_F64)
    return parrm_eval_periods(static_cast<const double*>(d_y), ld_y, d_sumsq, d_indices, n_chans,
                              n_indices, d_periods, n_periods, bandwidth, lambda, n_chans_divisor,
                              d_fit_error, d_workspace, workspace_bytes, stream);
  PARRM_REQUIRE(y_dtype == PARRM_F32, "parrm_eval_periods_typed: y must be float64 or float32");
  PARRM_REQUIRE(n_chans > 0 && n_indices > 0 && n_periods >= 0 && ld_y >= n_chans,
                "parrm_eval_periods_typed: bad shape");
  if (n_periods == 0) return PARRM_OK;
  if (bandwidth < 0 || bandwidth > PARRM_MAX_BANDWIDTH) {
    set_error("parrm_eval_periods: bandwidth %d outside [0, %d]", bandwidth, PARRM_MAX_BANDWIDTH);
    return PARRM_ERR_UNSUPPORTED;
  }
  PARRM_REQUIRE(d_y && d_workspace, "parrm_eval_periods_typed: null pointer");
  if (workspace_bytes <
      parrm_eval_workspace_bytes_typed(n_chans, n_indices, n_periods, bandwidth, y_dtype)) {
    set_error("parrm_eval_periods_typed: workspace too small");
    return PARRM_ERR_WORKSPACE;
  }
  const size_t base = eval_base_bytes(n_chans, n_indices, n_periods, bandwidth);
  double* wide = reinterpret_cast<double*>(static_cast<unsigned char*>(d_workspace) + base);
  const int64_t total = n_chans * n_indices;
  eval_widen_kernel<<<unsigned(min64(ceil_div(total, 256), 8 * kNumSMs)), 256, 0,
                      as_stream(stream)>>>(static_cast<const float*>(d_y), ld_y, n_indices,
                                           n_chans, wide);
  PARRM_LAUNCH_OK("eval_widen_kernel");
  return parrm_eval_periods(wide, n_chans, d_sumsq, d_indices, n_chans, n_indices, d_periods,
                            n_periods, bandwidth, lambda, n_chans_divisor, d_fit_error,
                            d_workspace, base, stream);
}

int parrm_argmin(const double* d_values, int64_t n, double* d_min_value, int64_t* d_min_index,
                 void* stream) {
  PARRM_REQUIRE(n > 0 && d_values && d_min_value && d_min_index, "parrm_argmin: bad arguments");
  parrm::argmin_kernel<<<1, 1024, 0, parrm::as_stream(stream)>>>(d_values, n, d_min_value,
                                                                 d_min_index);
  PARRM_LAUNCH_OK("argmin_kernel");
  return PARRM_OK;
}

}  // extern "C"
