// Micro-benchmark: FP64 tensor-core (mma.sync m8n8k4 f64) throughput on this GPU, next to
// the vector DFMA rate, to decide whether the evaluator's W'Y accumulation should use it.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ACC>
__global__ void dmma_kernel(int iters, double* sink) {
  double c[ACC][2];
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double r = 0;
  for (int i = 0; i < ACC; ++i) r += c[i][0] + c[i][1];
  if (r == 123.456) sink[0] = r;
}

template <int ACC>
void run(int threads, int blocks_per_sm) {
  double* sink;
  cudaMalloc(&sink, 8);
  const int iters = 1 << 14;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dmma_kernel<ACC><<<148 * blocks_per_sm, threads>>>(iters, sink);
  cudaEventRecord(e0);
  dmma_kernel<ACC><<<148 * blocks_per_sm, threads>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = double(iters) * ACC * 256.0 * (threads / 32) * 148 * blocks_per_sm;
  printf("acc %2d  threads %4d x %d/SM: %.3f ms  %.2f TFLOP/s (FMA = 2 flops)\n", ACC, threads,
         blocks_per_sm, ms, 2 * fma / ms / 1e9);
  cudaFree(sink);
}

int main() {
  run<4>(256, 1); run<8>(256, 1); run<12>(256, 1); run<8>(512, 1); run<12>(256, 2); run<12>(1024, 1);
  return 0;
}
