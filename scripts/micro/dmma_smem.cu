// Micro-benchmark: the evaluator's tensor-core inner loop in isolation -- 6 A fragments and 2 B
// fragments per k4 step read from shared memory (conflict-free strides), 12 DMMAs -- to see
// what fraction of the 37 TFLOP/s DMMA peak this operand pattern can reach.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kStrideW = 68, kStrideY = 68;

template <bool PREFETCH>
__global__ void loop_kernel(int tiles, double* sink) {
  extern __shared__ double sm[];
  double* s_w = sm;                    // [48][68]
  double* s_y = sm + 48 * kStrideW;    // [64][68]
  for (int i = threadIdx.x; i < 48 * kStrideW + 64 * kStrideY; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int kh = (warp >> 2) & 1, nq = warp & 3;
  double acc[6][2][2];
  for (int mb = 0; mb < 6; ++mb) for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  const double* wa = s_w + (l >> 2) * kStrideW + kh * 32 + (l & 3);
  const double* yb = s_y + (kh * 32 + (l & 3)) * kStrideY + nq * 16 + (l >> 2);
  for (int t = 0; t < tiles; ++t) {
    double a_cur[6], b_cur[2];
    for (int mb = 0; mb < 6; ++mb) a_cur[mb] = wa[mb * 8 * kStrideW];
    b_cur[0] = yb[0]; b_cur[1] = yb[8];
#pragma unroll 2
    for (int step = 0; step < 8; ++step) {
      double a_nxt[6], b_nxt[2];
      if (PREFETCH) {
        const int nx = (step + 1) & 7;
        for (int mb = 0; mb < 6; ++mb) a_nxt[mb] = wa[mb * 8 * kStrideW + nx * 4];
        b_nxt[0] = yb[nx * 4 * kStrideY]; b_nxt[1] = yb[nx * 4 * kStrideY + 8];
      } else {
        for (int mb = 0; mb < 6; ++mb) a_cur[mb] = wa[mb * 8 * kStrideW + step * 4];
        b_cur[0] = yb[step * 4 * kStrideY]; b_cur[1] = yb[step * 4 * kStrideY + 8];
      }
#pragma unroll
      for (int mb = 0; mb < 6; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[mb][nb][0]), "+d"(acc[mb][nb][1]) : "d"(a_cur[mb]), "d"(b_cur[nb]));
      if (PREFETCH) {
        for (int mb = 0; mb < 6; ++mb) a_cur[mb] = a_nxt[mb];
        b_cur[0] = b_nxt[0]; b_cur[1] = b_nxt[1];
      }
    }
  }
  double r = 0;
  for (int mb = 0; mb < 6; ++mb) for (int nb = 0; nb < 2; ++nb) r += acc[mb][nb][0] + acc[mb][nb][1];
  if (r == 123.456) sink[0] = r;
}

template <bool PREFETCH>
void run(int threads, int ctas) {
  double* sink; cudaMalloc(&sink, 8);
  const int tiles = 4000;
  const size_t smem = (48 * kStrideW + 64 * kStrideY) * 8;
  cudaFuncSetAttribute(loop_kernel<PREFETCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  loop_kernel<PREFETCH><<<148 * ctas, threads, smem>>>(tiles, sink);
  cudaEventRecord(e0);
  loop_kernel<PREFETCH><<<148 * ctas, threads, smem>>>(tiles, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = double(tiles) * 8 * 12 * 256.0 * (threads / 32) * 148 * ctas;
  printf("prefetch %d  threads %4d x %d CTA/SM: %.3f ms  %.2f TFLOP/s\n", PREFETCH, threads, ctas, ms,
         2 * fma / ms / 1e9);
  cudaFree(sink);
}

int main() {
  run<true>(256, 1); run<false>(256, 1); run<true>(256, 2); run<false>(256, 2); run<true>(512, 1);
  return 0;
}
