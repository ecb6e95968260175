// Experiment (not part of the product): latency of grid-wide barriers / reductions and
// FP64 atomic throughput on B200, to size the persistent PCG kernel and the Schur scatter.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_cg_sync(int iters, double* out) {
  cg::grid_group grid = cg::this_grid();
  for (int i = 0; i < iters; ++i) grid.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.0;
}

// hand-rolled barrier: one arrive per CTA on a monotonically increasing counter
__device__ __forceinline__ void my_barrier(unsigned int* ctr, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (*(volatile unsigned int*)ctr < target) {}
    __threadfence();
  }
  __syncthreads();
}
__global__ void k_my_sync(int iters, unsigned int* ctr, double* out) {
  unsigned int target = 0;
  for (int i = 0; i < iters; ++i) my_barrier(ctr, target);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.0;
}
// barrier + all-reduce of 2 doubles through per-CTA slots (what PCG needs)
__global__ void k_my_allreduce(int iters, unsigned int* ctr, double* slots, double* out) {
  unsigned int target = 0;
  __shared__ double sm[64];
  double v0 = threadIdx.x * 1e-3, v1 = 1.0;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    double a = v0, b = v1;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(~0u, a, o); b += __shfl_xor_sync(~0u, b, o); }
    if ((threadIdx.x & 31) == 0) { sm[(threadIdx.x >> 5) * 2] = a; sm[(threadIdx.x >> 5) * 2 + 1] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
      double x = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x * 2] : 0.0;
      double y = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x * 2 + 1] : 0.0;
      for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(~0u, x, o); y += __shfl_xor_sync(~0u, y, o); }
      if (threadIdx.x == 0) {
        double* s = slots + ((it & 1) * 2048 + blockIdx.x) * 2;
        s[0] = x; s[1] = y;
      }
    }
    my_barrier(ctr, target);
    // every warp 0 lane reads slots
    if (threadIdx.x < 32) {
      double x = 0.0, y = 0.0;
      for (int b2 = threadIdx.x; b2 < gridDim.x; b2 += 32) {
        const volatile double* s = slots + ((it & 1) * 2048 + b2) * 2;
        x += s[0]; y += s[1];
      }
      for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(~0u, x, o); y += __shfl_xor_sync(~0u, y, o); }
      if (threadIdx.x == 0) { sm[32] = x; sm[33] = y; }
    }
    __syncthreads();
    acc += sm[32] + sm[33];
    v0 = acc * 1e-9;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = acc;
}

__global__ void k_atomics(double* buf, size_t n_elems, int per_thread, unsigned seed) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
  for (int i = 0; i < per_thread; ++i) {
    x = x * 1664525u + 1013904223u;
    size_t blk = (size_t)(x >> 4) % (n_elems / 36);
    // 36 consecutive doubles like one 6x6 block
    for (int k = 0; k < 36; ++k) atomicAdd(buf + blk * 36 + k, 1.0);
  }
}
__global__ void k_atomics_lane(double* buf, size_t n_elems, int per_warp, unsigned seed) {
  // one 6x6 block per warp-step: lanes 0..31 (+4 tail) hit consecutive doubles
  const int lane = threadIdx.x & 31;
  unsigned x = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2654435761u + seed;
  for (int i = 0; i < per_warp; ++i) {
    x = x * 1664525u + 1013904223u;
    size_t blk = (size_t)(x >> 4) % (n_elems / 36);
    atomicAdd(buf + blk * 36 + lane, 1.0);
    if (lane < 4) atomicAdd(buf + blk * 36 + 32 + lane, 1.0);
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount;
  printf("SMs %d\n", nsm);
  double* out; CK(cudaMalloc(&out, 64));
  unsigned int* ctr; CK(cudaMalloc(&ctr, 64));
  double* slots; CK(cudaMalloc(&slots, 2 * 2048 * 2 * 8));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 2000;
  for (int threads : {256, 1024}) {
    for (int mult : {1, 2}) {
      if (threads == 1024 && mult == 2) continue;
      int grid = nsm * mult;
      float ms;
      void* args1[] = {(void*)&iters, (void*)&out};
      CK(cudaLaunchCooperativeKernel((void*)k_cg_sync, dim3(grid), dim3(threads), args1, 0, 0));
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      CK(cudaLaunchCooperativeKernel((void*)k_cg_sync, dim3(grid), dim3(threads), args1, 0, 0));
      cudaEventRecord(b); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, a, b);
      printf("cg grid.sync      grid %4d x %4d : %.3f us/sync\n", grid, threads, 1e3 * ms / iters);
      CK(cudaMemset(ctr, 0, 64));
      void* args2[] = {(void*)&iters, (void*)&ctr, (void*)&out};
      cudaEventRecord(a);
      CK(cudaLaunchCooperativeKernel((void*)k_my_sync, dim3(grid), dim3(threads), args2, 0, 0));
      cudaEventRecord(b); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, a, b);
      printf("atomic barrier    grid %4d x %4d : %.3f us/sync\n", grid, threads, 1e3 * ms / iters);
      CK(cudaMemset(ctr, 0, 64));
      void* args3[] = {(void*)&iters, (void*)&ctr, (void*)&slots, (void*)&out};
      cudaEventRecord(a);
      CK(cudaLaunchCooperativeKernel((void*)k_my_allreduce, dim3(grid), dim3(threads), args3, 0, 0));
      cudaEventRecord(b); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, a, b);
      printf("barrier+allreduce grid %4d x %4d : %.3f us/op\n", grid, threads, 1e3 * ms / iters);
    }
  }
  // atomics: 30 MB region (L2 resident) and 7 GB region
  for (size_t mb : {30ull, 4000ull}) {
    size_t n = mb * 1000000ull / 8 / 36 * 36;
    double* buf; CK(cudaMalloc(&buf, n * 8)); CK(cudaMemset(buf, 0, n * 8));
    float ms;
    k_atomics<<<nsm * 8, 256>>>(buf, n, 4, 1); CK(cudaDeviceSynchronize());
    cudaEventRecord(a); k_atomics<<<nsm * 8, 256>>>(buf, n, 16, 2); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b);
    double cnt = (double)nsm * 8 * 256 * 16 * 36;
    printf("atomicAdd f64 thread-per-block  region %5zu MB: %.1f G atomics/s\n", mb, cnt / ms * 1e-6);
    cudaEventRecord(a); k_atomics_lane<<<nsm * 8, 256>>>(buf, n, 256, 3); cudaEventRecord(b); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, a, b);
    cnt = (double)nsm * 8 * 8 * 256 * 36;
    printf("atomicAdd f64 lane-per-element  region %5zu MB: %.1f G atomics/s\n", mb, cnt / ms * 1e-6);
    cudaFree(buf);
  }
  return 0;
}
