// Phase timing (clock64) of the 64x64 diagonal-block kernel of the dense Cholesky, in isolation.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o potrf_bench scripts/potrf_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
constexpr int CB = 64;
// variant 0: left-looking, thread per row, dot products with NCH chains; variant 1: right-looking
template <int VARIANT, int NCH>
__global__ void __launch_bounds__(CB) potrf(double* A, long long ld, double* Linv, long long* clk) {
  __shared__ double T[CB][CB + 1];
  __shared__ double dinv[CB], lcol[CB];
  __shared__ double piv;
  const int i = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int r = 0; r < CB; ++r) T[r][i] = (i <= r) ? A[(size_t)r * ld + i] : 0.0;
  __syncthreads();
  long long t1 = clock64();
  if (VARIANT == 0) {
    for (int j = 0; j < CB; ++j) {
      double s[4] = {0, 0, 0, 0};
      if (i >= j) {
        int k = 0;
        for (; k + NCH - 1 < j; k += NCH)
#pragma unroll
          for (int q = 0; q < NCH; ++q) s[q] += T[i][k + q] * T[j][k + q];
        for (; k < j; ++k) s[0] += T[i][k] * T[j][k];
      }
      const double sres = (i >= j) ? T[i][j] - ((s[0] + s[1]) + (s[2] + s[3])) : 0.0;
      if (i == j) piv = rsqrt(sres);
      __syncthreads();
      if (i >= j) T[i][j] = sres * piv;
      if (i == j) dinv[j] = piv;
      __syncthreads();
    }
  } else {
    for (int j = 0; j < CB; ++j) {
      if (i == j) piv = rsqrt(T[j][j]);
      __syncthreads();
      if (i >= j) { const double v = T[i][j] * piv; T[i][j] = v; lcol[i] = v; if (i == j) dinv[j] = piv; }
      __syncthreads();
      if (i > j) { const double li = lcol[i];
#pragma unroll 4
        for (int c = j + 1; c <= i; ++c) T[i][c] -= li * lcol[c]; }
    }
    __syncthreads();
  }
  long long t2 = clock64();
  if (VARIANT == 0) {
    double xi = dinv[i];
    for (int r = i + 1; r < CB; ++r) {
      double s[4] = {-T[r][i] * xi, 0, 0, 0};
      int k = i + 1;
      for (; k + NCH - 1 < r; k += NCH)
#pragma unroll
        for (int q = 0; q < NCH; ++q) s[q] -= T[r][k + q] * T[i][k + q];
      for (; k < r; ++k) s[0] -= T[r][k] * T[i][k];
      T[i][r] = ((s[0] + s[1]) + (s[2] + s[3])) * dinv[r];
    }
  } else {
    for (int k = 0; k < CB - 1; ++k) {
      if (k >= i) {
        const double xk = k == i ? dinv[i] : T[i][k] * dinv[k];
        if (k > i) T[i][k] = xk;
#pragma unroll 4
        for (int r = k + 1; r < CB; ++r) T[i][r] -= T[r][k] * xk;
      }
    }
    if (i < CB - 1) T[i][CB - 1] *= dinv[CB - 1];
  }
  __syncthreads();
  long long t3 = clock64();
#pragma unroll 8
  for (int r = 0; r < CB; ++r) {
    if (i <= r) A[(size_t)r * ld + i] = T[r][i];
    Linv[r * CB + i] = i < r ? T[i][r] : (i == r ? dinv[r] : 0.0);
  }
  long long t4 = clock64();
  if (i == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}

// variant 2: left-looking, 128-bit shared loads (row stride 66 doubles), two chains per half
__global__ void __launch_bounds__(CB) potrf_v2(double* A, long long ld, double* Linv, long long* clk) {
  constexpr int LD = CB + 2;
  __shared__ __align__(16) double T[CB * LD];
  __shared__ double dinv[CB];
  __shared__ double piv;
  const int i = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int r = 0; r < CB; ++r) T[r * LD + i] = (i <= r) ? A[(size_t)r * ld + i] : 0.0;
  __syncthreads();
  long long t1 = clock64();
  for (int j = 0; j < CB; ++j) {
    double s0 = 0.0, s1 = 0.0;
    if (i >= j) {
      const double2* a = reinterpret_cast<const double2*>(T + i * LD);
      const double2* b = reinterpret_cast<const double2*>(T + j * LD);
      int k = 0;
#pragma unroll 4
      for (; k + 1 < j; k += 2) {
        const double2 x = a[k >> 1], y = b[k >> 1];
        s0 += x.x * y.x;
        s1 += x.y * y.y;
      }
      if (k < j) s0 += T[i * LD + k] * T[j * LD + k];
    }
    const double sres = (i >= j) ? T[i * LD + j] - (s0 + s1) : 0.0;
    if (i == j) piv = rsqrt(sres);
    __syncthreads();
    if (i >= j) T[i * LD + j] = sres * piv;
    if (i == j) dinv[j] = piv;
    __syncthreads();
  }
  long long t2 = clock64();
  {
    // column i of L^-1 kept at T[i][r], r > i (upper triangle)
    double xi = dinv[i];
    for (int r = i + 1; r < CB; ++r) {
      double s0 = -T[r * LD + i] * xi, s1 = 0.0;
      int k = i + 1;
      if (k & 1) { if (k < r) { s0 -= T[r * LD + k] * T[i * LD + k]; } ++k; }
      const double2* a = reinterpret_cast<const double2*>(T + r * LD);
      const double2* b = reinterpret_cast<const double2*>(T + i * LD);
#pragma unroll 4
      for (; k + 1 < r; k += 2) {
        const double2 x = a[k >> 1], y = b[k >> 1];
        s0 -= x.x * y.x;
        s1 -= x.y * y.y;
      }
      if (k < r) s0 -= T[r * LD + k] * T[i * LD + k];
      T[i * LD + r] = (s0 + s1) * dinv[r];
    }
  }
  __syncthreads();
  long long t3 = clock64();
#pragma unroll 8
  for (int r = 0; r < CB; ++r) {
    if (i <= r) A[(size_t)r * ld + i] = T[r * LD + i];
    Linv[r * CB + i] = i < r ? T[i * LD + r] : (i == r ? dinv[r] : 0.0);
  }
  long long t4 = clock64();
  if (i == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}
template <int V, int N>
void run(const char* name, double* dA0, double* dA, double* dL, long long* dclk, int ld) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 20; ++rep) {
    cudaMemcpy(dA, dA0, sizeof(double) * 64 * ld, cudaMemcpyDeviceToDevice);
    cudaEventRecord(a); if (V == 2) potrf_v2<<<1, CB>>>(dA, ld, dL, dclk); else potrf<V == 2 ? 0 : V, N><<<1, CB>>>(dA, ld, dL, dclk); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); best = fminf(best, ms);
  }
  long long h[4]; cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost);
  std::vector<double> L(64 * 64), A(64 * ld);
  cudaMemcpy(L.data(), dL, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
  cudaMemcpy(A.data(), dA, sizeof(double) * 64 * ld, cudaMemcpyDeviceToHost);
  double err = 0;  // ||Linv * L - I||
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) { double s = 0; for (int k = c; k <= r; ++k) s += L[r * 64 + k] * A[(size_t)k * ld + c]; err = fmax(err, fabs(s - (r == c))); }
  printf("%-28s %7.1f us   cycles: load %lld  cholesky %lld  inverse %lld  store %lld   |Linv L - I| %.1e\n", name, best * 1e3, h[0], h[1], h[2], h[3], err);
}
int main() {
  const int ld = 12032;
  std::vector<double> M(64 * 64), A((size_t)64 * ld, 0.0);
  srand(1); for (auto& v : M) v = rand() / (double)RAND_MAX - 0.5;
  for (int r = 0; r < 64; ++r) for (int c = 0; c <= r; ++c) { double s = r == c ? 64.0 : 0.0; for (int k = 0; k < 64; ++k) s += M[r * 64 + k] * M[c * 64 + k]; A[(size_t)r * ld + c] = s; }
  double *dA0, *dA, *dL; long long* dclk;
  cudaMalloc(&dA0, sizeof(double) * 64 * ld); cudaMalloc(&dA, sizeof(double) * 64 * ld); cudaMalloc(&dL, sizeof(double) * 4096); cudaMalloc(&dclk, 64);
  cudaMemcpy(dA0, A.data(), sizeof(double) * 64 * ld, cudaMemcpyHostToDevice);
  run<0, 1>("left-looking, 1 chain", dA0, dA, dL, dclk, ld);
  run<0, 2>("left-looking, 2 chains", dA0, dA, dL, dclk, ld);
  run<0, 4>("left-looking, 4 chains", dA0, dA, dL, dclk, ld);
  run<1, 1>("right-looking", dA0, dA, dL, dclk, ld);
  run<2, 2>("left-looking, 128-bit loads", dA0, dA, dL, dclk, ld);
  return 0;
}
