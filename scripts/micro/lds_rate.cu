// Micro-benchmark: shared-memory LDS.64 throughput per SM versus warps per SM and loads in
// flight per warp, with each load consumed by a dependent DADD (the filter gather's pattern).
#include <cstdio>
#include <cuda_runtime.h>

template <int G>
__global__ void lds_kernel(int iters, int stride, double* sink, long long* cycles) {
  extern __shared__ double s[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = i * 1e-3;
  __syncthreads();
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const double* p = s + (threadIdx.x & 1023);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double v[G];
    const double* q = p + ((it * stride) & 1023);
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = q[g * 256];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g % 6] += v[g];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  double r = 0;
  for (int g = 0; g < 6; ++g) r += acc[g];
  if (r == 123.456) sink[0] = r;
}

template <int G>
void run(int threads) {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  cudaFuncSetAttribute(lds_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192 * 8);
  lds_kernel<G><<<148, threads, 8192 * 8 + 1024 * 8 * 0, 0>>>(iters, 7, sink, cyc);
  lds_kernel<G><<<148, threads, 8192 * 8, 0>>>(iters, 7, sink, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = h[0];
  double loads = double(iters) * G * (threads / 32);
  printf("threads %4d  G %2d  cycles %9.0f  LDS.64/clk %.3f  wavefronts/clk %.3f\n", threads, G, c,
         loads / c, 2 * loads / c);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 384, 512, 1024}) {
    run<4>(threads); run<6>(threads); run<12>(threads); run<24>(threads);
  }
  return 0;
}
