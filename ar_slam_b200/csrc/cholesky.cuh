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

// --- factor the CB x CB diagonal block and invert its factor (one CTA, one thread per row) --
// Left-looking column Cholesky in shared memory, then every thread forward-substitutes one
// column of the identity: Linv = L^-1 (lower triangular).  With Linv the panel solve below is
// a GEMM and runs on the FP64 tensor pipe as well.
__global__ void __launch_bounds__(CB) potrf_inv_kernel(double* __restrict__ A, long long ld, int k0,
                                                       double* __restrict__ Linv, double* __restrict__ fail) {
  // one array, two triangles: the lower one (with diagonal) holds L, the strictly upper one
  // receives Linv^T (Linv is lower triangular, its diagonal is 1 / L[r][r])
  __shared__ double T[CB][CB + 1];
  __shared__ double dinv[CB];
  __shared__ double piv;
  const int i = threadIdx.x;
  for (int c = 0; c < CB; ++c) T[i][c] = (c <= i) ? A[(size_t)(k0 + i) * ld + k0 + c] : 0.0;
  __syncthreads();
  for (int j = 0; j < CB; ++j) {
    double s0 = 0.0, s1 = 0.0;
    if (i >= j) {
      int k = 0;
      for (; k + 1 < j; k += 2) {
        s0 += T[i][k] * T[j][k];
        s1 += T[i][k + 1] * T[j][k + 1];
      }
      if (k < j) s0 += T[i][k] * T[j][k];
    }
    const double sres = (i >= j) ? T[i][j] - (s0 + s1) : 0.0;
    if (i == j) {
      if (!(sres > 0.0)) *fail = 1.0;
      piv = 1.0 / sqrt(sres);
    }
    __syncthreads();
    if (i >= j) T[i][j] = sres * piv;  // i == j: d / sqrt(d) = sqrt(d)
    if (i == j) dinv[j] = piv;
    __syncthreads();
  }
  // column i of L^-1: x_r = (delta_ri - sum_{k=i}^{r-1} L[r][k] x_k) / L[r][r], x_r kept at T[i][r]
  double xi = dinv[i];  // x_i
  for (int r = i + 1; r < CB; ++r) {
    double s0 = -T[r][i] * xi, s1 = 0.0;
    int k = i + 1;
    for (; k + 1 < r; k += 2) {
      s0 -= T[r][k] * T[i][k];
      s1 -= T[r][k + 1] * T[i][k + 1];
    }
    if (k < r) s0 -= T[r][k] * T[i][k];
    T[i][r] = (s0 + s1) * dinv[r];
  }
  __syncthreads();
  for (int c = 0; c < CB; ++c) {
    if (c <= i) A[(size_t)(k0 + i) * ld + k0 + c] = T[i][c];
    // Linv[i][c] = x_i of column c = T[c][i] for c < i, dinv on the diagonal, 0 above
    Linv[i * CB + c] = c < i ? T[c][i] : (c == i ? dinv[i] : 0.0);
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

// --- panel solve as a GEMM: X = A_tile Linv^T on 64x64 tiles (in place), DMMA ----------------
__global__ void __launch_bounds__(128) trsm_dmma_kernel(double* __restrict__ A, long long ld, int k0,
                                                        const double* __restrict__ Linv) {
  extern __shared__ double sm[];
  double(*Pi)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm);
  double(*Pj)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm + CB * CB_LD);
  const int r0 = k0 + CB + blockIdx.x * CB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int e = tid; e < CB * CB; e += 128) {
    const int r = e / CB, c = e % CB;
    Pi[r][c] = A[(size_t)(r0 + r) * ld + k0 + c];
    Pj[r][c] = Linv[e];
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
      double2* c = reinterpret_cast<double2*>(A + (size_t)(r0 + rb + mi * 8 + fr) * ld + k0 + cb + ni * 8 + 2 * fk);
      *c = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
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
  static constexpr size_t kTileSmem = 2 * CB * CB_LD * sizeof(double);
  static cudaError_t init() {
    cudaError_t e = cudaFuncSetAttribute(trsm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
  }
  // factor rows/cols [0, n_pad); linv = 64 x 64 scratch; returns number of launches
  static int factor(double* A, long long ld, int n_pad, double* linv, double* fail, cudaStream_t st) {
    int launches = 0;
    for (int k0 = 0; k0 < n_pad; k0 += CB) {
      potrf_inv_kernel<<<1, CB, 0, st>>>(A, ld, k0, linv, fail);
      ++launches;
      const int rem = (n_pad - k0 - CB) / CB;
      if (rem > 0) {
        trsm_dmma_kernel<<<rem, 128, kTileSmem, st>>>(A, ld, k0, linv);
        syrk_dmma_kernel<<<rem * (rem + 1) / 2, 128, kTileSmem, st>>>(A, ld, k0);
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
