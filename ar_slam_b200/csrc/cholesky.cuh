// cholesky.cuh -- kernel (4): blocked dense Cholesky of the reduced system in
// FP64 with the trailing update on the FP64 tensor pipe (DMMA,
// mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind).
//
// Replaces Eigen::LLT inside Ceres' DenseSchurComplementSolver (reached from
// reference ar_slam/src/ar_slam_util.cpp:1011 DENSE_SCHUR).  Lower triangle,
// row-major, leading dimension ld, dimension n_pad (multiple of 64).  The
// right-hand side rides along as one extra row, so the sweep also does the
// forward substitution; dense_backsolve then solves L^T y = w.
// Failure (non-positive pivot) raises *fail like Eigen's info() != Success.
#pragma once
#include <cuda_runtime.h>

namespace ars {

constexpr int CB = 64;        // panel width / tile size
constexpr int CB_LD = CB + 4; // shared-memory row stride (bank-conflict free for the DMMA fragments)

// --- factor the CB x CB diagonal block in shared memory (one CTA) ---------
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, long long ld, int k0,
                                                         double* __restrict__ fail) {
  __shared__ double T[CB][CB + 1];
  const int tid = threadIdx.x;
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e / CB, c = e % CB;
    T[r][c] = (c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < CB; ++j) {
    if (tid == 0) {
      const double d = T[j][j];
      if (!(d > 0.0)) *fail = 1.0;
      T[j][j] = sqrt(d);
    }
    __syncthreads();
    if (tid > j && tid < CB) T[tid][j] /= T[j][j];
    __syncthreads();
    // trailing update of the lower triangle: T[i][c] -= T[i][j] T[c][j], j < c <= i
    const int m = CB - 1 - j;  // remaining rows/cols
    for (int e = tid; e < m * m; e += 256) {
      const int i = j + 1 + e / m, c = j + 1 + e % m;
      if (c <= i) T[i][c] -= T[i][j] * T[c][j];
    }
    __syncthreads();
  }
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e / CB, c = e % CB;
    if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = T[r][c];
  }
}

// --- panel solve: X L_kk^T = A_tile, one thread per row, 64 rows per CTA ---
__global__ void __launch_bounds__(CB) trsm_panel_kernel(double* __restrict__ A, long long ld, int k0) {
  extern __shared__ double sm[];
  double(*L)[CB + 1] = reinterpret_cast<double(*)[CB + 1]>(sm);
  double(*X)[CB + 1] = reinterpret_cast<double(*)[CB + 1]>(sm + CB * (CB + 1));
  const int tid = threadIdx.x;
  const int r0 = k0 + CB + blockIdx.x * CB;
  for (int e = tid; e < CB * CB; e += CB) {
    const int r = e / CB, c = e % CB;
    L[r][c] = A[(size_t)(k0 + r) * ld + k0 + c];
    X[r][c] = A[(size_t)(r0 + r) * ld + k0 + c];
  }
  __syncthreads();
  double x[CB];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    double s = X[tid][c];
#pragma unroll
    for (int m = 0; m < c; ++m) s -= x[m] * L[c][m];
    x[c] = s / L[c][c];
  }
#pragma unroll
  for (int c = 0; c < CB; ++c) X[tid][c] = x[c];
  __syncthreads();
  for (int e = tid; e < CB * CB; e += CB) {
    const int r = e / CB, c = e % CB;
    A[(size_t)(r0 + r) * ld + k0 + c] = X[r][c];
  }
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// --- trailing update C(ti,tj) -= P(ti) P(tj)^T on 64x64 tiles, tj <= ti ----
// 4 warps per CTA, each a 32x32 quadrant = 4x4 DMMA tiles, K = 64.
__global__ void __launch_bounds__(128) syrk_dmma_kernel(double* __restrict__ A, long long ld, int k0) {
  extern __shared__ double sm[];
  double(*Pi)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm);
  double(*Pj)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm + CB * CB_LD);
  // decode linear tile id -> (ti >= tj)
  const int p = blockIdx.x;
  int tj, ti;
  ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
  while (ti * (ti + 1) / 2 > p) --ti;
  tj = p - ti * (ti + 1) / 2;
  const int r0 = k0 + CB + ti * CB, c0 = k0 + CB + tj * CB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int e = tid; e < CB * CB; e += 128) {
    const int r = e / CB, c = e % CB;
    Pi[r][c] = A[(size_t)(r0 + r) * ld + k0 + c];
    Pj[r][c] = A[(size_t)(c0 + r) * ld + k0 + c];
  }
  __syncthreads();
  const int rb = (w >> 1) * 32, cb = (w & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < CB; kk += 4) {
    double af[4], bf[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) af[mi] = Pi[rb + mi * 8 + fr][kk + fk];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) bf[ni] = Pj[cb + ni * 8 + fr][kk + fk];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2* c = reinterpret_cast<double2*>(A + (size_t)(r0 + rb + mi * 8 + fr) * ld + c0 + cb + ni * 8 + 2 * fk);
      double2 v = *c;
      v.x -= acc[mi][ni][0];
      v.y -= acc[mi][ni][1];
      *c = v;
    }
}

// --- one 64-block step of L^T y = w (w = row rhs_row), from the bottom up ---
// Every CTA solves the diagonal block redundantly in shared memory, CTA 0
// publishes y_k, and CTA b folds y_k into w[256 b .. 256 b + 255] (< k0).
__global__ void __launch_bounds__(256) backsolve_step_kernel(double* __restrict__ A, long long ld, int k0,
                                                             int n, int rhs_row, double* __restrict__ y) {
  __shared__ double L[CB][CB + 1];
  __shared__ double wv[CB];
  const int tid = threadIdx.x;
  const int nv = min(CB, n - k0);
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e / CB, c = e % CB;
    L[r][c] = (c <= r && r < nv) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
  }
  if (tid < CB) wv[tid] = tid < nv ? A[(size_t)rhs_row * ld + k0 + tid] : 0.0;
  __syncthreads();
  for (int c = nv - 1; c >= 0; --c) {
    if (tid == 0) wv[c] = wv[c] / L[c][c];
    __syncthreads();
    if (tid < c) wv[tid] -= L[c][tid] * wv[c];
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid < nv) y[k0 + tid] = wv[tid];
  const int j = blockIdx.x * 256 + tid;
  if (j < k0) {
    double s = 0.0;
#pragma unroll 8
    for (int r = 0; r < nv; ++r) s += A[(size_t)(k0 + r) * ld + j] * wv[r];
    A[(size_t)rhs_row * ld + j] -= s;
  }
}

struct DenseCholesky {
  static constexpr size_t kTrsmSmem = 2 * CB * (CB + 1) * sizeof(double);
  static constexpr size_t kSyrkSmem = 2 * CB * CB_LD * sizeof(double);
  static cudaError_t init() {
    cudaError_t e = cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kTrsmSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem);
  }
  // factor rows/cols [0, n_pad); returns number of launches
  static int factor(double* A, long long ld, int n_pad, double* fail, cudaStream_t st) {
    int launches = 0;
    for (int k0 = 0; k0 < n_pad; k0 += CB) {
      potrf_diag_kernel<<<1, 256, 0, st>>>(A, ld, k0, fail);
      ++launches;
      const int rem = (n_pad - k0 - CB) / CB;
      if (rem > 0) {
        trsm_panel_kernel<<<rem, CB, kTrsmSmem, st>>>(A, ld, k0);
        syrk_dmma_kernel<<<rem * (rem + 1) / 2, 128, kSyrkSmem, st>>>(A, ld, k0);
        launches += 2;
      }
    }
    return launches;
  }
  // solves L^T y = w for the leading n unknowns; w is row rhs_row (== n) and is destroyed
  static int backsolve(double* A, long long ld, int n, int rhs_row, double* y, cudaStream_t st) {
    int launches = 0;
    for (int k0 = ((n - 1) / CB) * CB; k0 >= 0; k0 -= CB) {
      const int grid = k0 > 0 ? (k0 + 255) / 256 : 1;
      backsolve_step_kernel<<<grid, 256, 0, st>>>(A, ld, k0, n, rhs_row, y);
      ++launches;
    }
    return launches;
  }
};

}  // namespace ars
