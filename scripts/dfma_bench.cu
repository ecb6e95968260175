// FP64 FMA issue rate per warp on B200: independent DFMA chains, 1 CTA per SM, W warps per CTA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_bench scripts/dfma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, int iters, long long* clk) {
  double a[CHAINS];
  const double x = 1.0000001, y = 1e-9 * threadIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) a[c] = c + y;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], x, y);
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += a[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&clk, 8);
  const int iters = 4096;
  printf("warps/SM  chains  cycles/DFMA-per-warp  DFMA lanes/clk/SM\n");
  for (int warps : {1, 2, 4, 8, 12, 16, 24, 32}) {
    for (int chains : {1, 4, 16}) {
      if (chains == 1) k<1><<<148, warps * 32>>>(out, iters, clk);
      else if (chains == 4) k<4><<<148, warps * 32>>>(out, iters, clk);
      else k<16><<<148, warps * 32>>>(out, iters, clk);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / ((double)iters * chains);
      printf("%5d %7d %16.2f %20.1f\n", warps, chains, per, warps * 32 / per);
    }
  }
  return 0;
}
