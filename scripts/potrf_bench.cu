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

// variant 3: blocked.  The 64 x 64 block is a 4 x 4 grid of 16 x 16 sub-blocks.  The diagonal sub-blocks are
// factored and inverted by ONE warp with the rows in registers and shuffles instead of shared-memory round
// trips and CTA barriers; panel solves, trailing updates and the block inverse are 16 x 16 x 16 products by
// all 256 threads (one output element each).
constexpr int SB = 16;
constexpr int TLD = CB + 1;
__device__ __forceinline__ void potrf16_warp(double* T /* block (kb,kb), row stride TLD */, double* X /* same position in the inverse */,
                                             double* fail) {
  const int lane = threadIdx.x & 31;
  const int i = lane & 15;  // lanes 16..31 mirror lanes 0..15 (keeps the shuffles full-warp)
  double a[SB], p[SB], x[SB];
#pragma unroll
  for (int c = 0; c < SB; ++c) a[c] = (c <= i) ? T[i * TLD + c] : 0.0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    if (!(d > 0.0) && lane == 0) *fail = 1.0;
    p[j] = rsqrt(d);
    a[j] *= p[j];  // lane j: d / sqrt(d) = sqrt(d); lanes i > j: l_ij
#pragma unroll
    for (int c = j + 1; c < SB; ++c) {
      const double lc = __shfl_sync(0xffffffffu, a[j], c);
      a[c] -= a[j] * lc;  // only c <= i is used later
    }
  }
  // inverse, lane c owns column c: x_r = p_r (delta_rc - sum_{k<r} L[r][k] x_k)
#pragma unroll
  for (int r = 0; r < SB; ++r) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < r; ++k) sum += __shfl_sync(0xffffffffu, a[k], r) * x[k];
    x[r] = r < i ? 0.0 : (r == i ? p[r] : -p[r] * sum);
  }
  if (lane < SB) {
#pragma unroll
    for (int c = 0; c < SB; ++c) {
      if (c <= i) T[i * TLD + c] = a[c];
      X[c * TLD + i] = c >= i ? x[c] : 0.0;  // X[r][col i] = x[r]
    }
  }
}
__global__ void __launch_bounds__(256) potrf_v3(double* A, long long ld, double* Linv, long long* clk, double* fail) {
  extern __shared__ double sm3[];
  double* T = sm3;                 // [64][65] L
  double* X = sm3 + CB * TLD;      // [64][65] L^-1 (lower)
  double* S = X + CB * TLD;        // [3][16][17] temporaries
  const int tid = threadIdx.x;
  const int er = tid >> 4, ec = tid & 15;  // the output element this thread owns in a 16 x 16 product
  long long t0 = clock64();
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e >> 6, c = e & 63;
    T[r * TLD + c] = (c <= r) ? A[(size_t)r * ld + c] : 0.0;
    X[r * TLD + c] = 0.0;
  }
  __syncthreads();
  long long t1 = clock64();
  for (int kb = 0; kb < 4; ++kb) {
    if (tid < 32) potrf16_warp(T + (kb * SB) * TLD + kb * SB, X + (kb * SB) * TLD + kb * SB, fail);
    __syncthreads();
    // panel: L(ib,kb) = A(ib,kb) Xkk^T, element (r,c) = sum_{m<=c} A(r,m) Xkk(c,m)
    double v[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int ib = kb + 1 + q;
      v[q] = 0.0;
      if (ib < 4) {
        const double* a = T + (ib * SB + er) * TLD + kb * SB;
        const double* xk = X + (kb * SB + ec) * TLD + kb * SB;
#pragma unroll
        for (int m = 0; m < SB; ++m) v[q] += a[m] * xk[m];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int ib = kb + 1 + q;
      if (ib < 4) T[(ib * SB + er) * TLD + kb * SB + ec] = v[q];
    }
    __syncthreads();
    // trailing update: T(ib,jb) -= L(ib,kb) L(jb,kb)^T, ib >= jb > kb
    for (int ib = kb + 1; ib < 4; ++ib)
      for (int jb = kb + 1; jb <= ib; ++jb) {
        const double* a = T + (ib * SB + er) * TLD + kb * SB;
        const double* b = T + (jb * SB + ec) * TLD + kb * SB;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < SB; ++m) s += a[m] * b[m];
        T[(ib * SB + er) * TLD + jb * SB + ec] -= s;
      }
    __syncthreads();
  }
  long long t2 = clock64();
  // block inverse by anti-diagonals: X(i,j) = -Xii * sum_{k=j}^{i-1} L(i,k) X(k,j)
  for (int d = 1; d < 4; ++d) {
    const int nb = 4 - d;  // blocks (i, j) = (j + d, j), j = 0 .. nb - 1
    for (int j = 0; j < nb; ++j) {
      const int i = j + d;
      double s = 0.0;
      for (int k = j; k < i; ++k) {
        const double* a = T + (i * SB + er) * TLD + k * SB;   // L(i,k) row er
        const double* b = X + (k * SB) * TLD + j * SB + ec;   // X(k,j) column ec
#pragma unroll
        for (int m = 0; m < SB; ++m) s += a[m] * b[m * TLD];
      }
      S[(j * SB + er) * (SB + 1) + ec] = s;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      const int i = j + d;
      const double* a = X + (i * SB + er) * TLD + i * SB;     // Xii row er (lower triangular)
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < SB; ++m) s += a[m] * S[(j * SB + m) * (SB + 1) + ec];
      X[(i * SB + er) * TLD + j * SB + ec] = -s;
    }
    __syncthreads();
  }
  long long t3 = clock64();
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (c <= r) A[(size_t)r * ld + c] = T[r * TLD + c];
    Linv[e] = c <= r ? X[r * TLD + c] : 0.0;
  }
  long long t4 = clock64();
  if (tid == 0 && clk) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}
template <int V, int N>
void run(const char* name, double* dA0, double* dA, double* dL, long long* dclk, int ld) {
  static double* dfail = nullptr;
  if (!dfail) { cudaMalloc(&dfail, 8); cudaMemset(dfail, 0, 8); cudaFuncSetAttribute(potrf_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * CB * TLD + 3 * SB * (SB + 1)) * sizeof(double))); }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 20; ++rep) {
    cudaMemcpy(dA, dA0, sizeof(double) * 64 * ld, cudaMemcpyDeviceToDevice);
    cudaEventRecord(a); if (V == 3) potrf_v3<<<1, 256, (2 * CB * TLD + 3 * SB * (SB + 1)) * sizeof(double)>>>(dA, ld, dL, dclk, dfail); else if (V == 2) potrf_v2<<<1, CB>>>(dA, ld, dL, dclk); else potrf<(V >= 2) ? 0 : V, N><<<1, CB>>>(dA, ld, dL, dclk); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); best = fminf(best, ms);
  }
  long long h[4]; cudaMemcpy(h, dclk, sizeof(h), cudaMemcpyDeviceToHost);
  std::vector<double> L(64 * 64), A(64 * ld);
  cudaMemcpy(L.data(), dL, sizeof(double) * 4096, cudaMemcpyDeviceToHost);
  cudaMemcpy(A.data(), dA, sizeof(double) * 64 * ld, cudaMemcpyDeviceToHost);
  double err = 0;  // ||Linv * L - I||
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) { double s = 0; for (int k = c; k <= r; ++k) s += L[r * 64 + k] * A[(size_t)k * ld + c]; err = fmax(err, fabs(s - (r == c))); }
  { std::vector<double> A0((size_t)64 * ld); cudaMemcpy(A0.data(), dA0, sizeof(double) * 64 * ld, cudaMemcpyDeviceToHost); double e2 = 0; for (int r = 0; r < 64; ++r) for (int c = 0; c <= r; ++c) { double s2 = 0; for (int k = 0; k <= c; ++k) s2 += A[(size_t)r * ld + k] * A[(size_t)c * ld + k]; e2 = fmax(e2, fabs(s2 - A0[(size_t)r * ld + c])); } printf("|L L^T - A| %.1e  ", e2); }
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
  run<3, 1>("blocked 16x16, 256 threads", dA0, dA, dL, dclk, ld);
  return 0;
}
