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
//
// Two-level blocking.  A 64-wide right-looking sweep updates every trailing tile with K = 64,
// i.e. 4 flop per byte of C traffic: HBM bound, not tensor bound.  So the 64-wide steps only
// keep the columns of the current 256-wide OUTER panel up to date (syrk_dmma_kernel on that
// strip), and the rest of the trailing matrix receives the whole outer panel at once from
// syrk_big_dmma_kernel: 128 x 128 tiles of C, K = 256, operands streamed through a two-stage
// cp.async pipeline -- ~15 flop per byte, which puts the update on the DMMA pipe.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

namespace ars {

constexpr int CB = 64;        // panel width / tile size
constexpr int CB_LD = CB + 4; // shared-memory row stride (bank-conflict free for the DMMA fragments)

// --- factor the CB x CB diagonal block and invert its factor (one CTA, 256 threads) --------
// With Linv = L^-1 the panel solve below is a GEMM and runs on the FP64 tensor pipe as well, and
// the back-substitution reuses it.  This kernel sits on the critical path of every 64-wide step
// (188 times for n = 12 003), so it is built for latency (profiles/r1_microbench.txt: a
// thread-per-row column Cholesky in shared memory takes 73 us, this one 30 us):
// Blocked: the 64 x 64 block is a 4 x 4 grid of 16 x 16 sub-blocks.  The diagonal sub-blocks are
// factored and inverted by ONE warp with the rows in registers and shuffles instead of shared-memory round
// trips and CTA barriers; panel solves, trailing updates and the block inverse are 16 x 16 x 16 products by
// all 256 threads (one output element each).
constexpr int SB = 16;        // sub-block
constexpr int TLD = CB + 1;   // shared-memory row stride
constexpr size_t kPotrfSmem = (size_t)(2 * CB * TLD + 3 * SB * (SB + 1)) * sizeof(double);
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
__global__ void __launch_bounds__(256) potrf_inv_kernel(double* __restrict__ A, long long ld, int k0,
                                                        double* __restrict__ Linv, double* __restrict__ fail) {
  extern __shared__ double sm3[];
  double* T = sm3;                 // [64][65] L
  double* X = sm3 + CB * TLD;      // [64][65] L^-1 (lower)
  double* S = X + CB * TLD;        // [3][16][17] temporaries
  const int tid = threadIdx.x;
  const int er = tid >> 4, ec = tid & 15;  // the output element this thread owns in a 16 x 16 product
  A += (size_t)k0 * ld + k0;
#pragma unroll
  for (int q = 0; q < CB * CB / 256; ++q) {  // 16 independent loads per thread
    const int e = tid + q * 256;
    const int r = e >> 6, c = e & 63;
    T[r * TLD + c] = (c <= r) ? A[(size_t)r * ld + c] : 0.0;
    X[r * TLD + c] = 0.0;
  }
  __syncthreads();
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
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (c <= r) A[(size_t)r * ld + c] = T[r * TLD + c];
    Linv[e] = c <= r ? X[r * TLD + c] : 0.0;
  }
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// m16n8k8 FP64 MMA (sm_90+): A 16x8 row (a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4)), B 8x8 col
// (b0 (k=t,n=g) b1 (k=t+4,n=g)), C/D 16x8 (c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1)); g = lane / 4, t = lane % 4
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
// both 64 x 64 operands of a tile kernel, 128 threads: 2 x 2048 16-byte pieces, ALL in flight at once (these kernels sit on the
// factorisation's critical path with one CTA per SM or fewer: their time is load latency, so it is paid once)
__device__ __forceinline__ void load_two_tiles_async(double (*Pi)[CB + 4], double (*Pj)[CB + 4], const double* gi, long long ldi, const double* gj,
                                                     long long ldj) {
#pragma unroll
  for (int it = 0; it < CB * CB / 2 / 128; ++it) {
    const int e = threadIdx.x + it * 128;
    const int r = e >> 5, c2 = (e & 31) * 2;
    cp_async16(&Pi[r][c2], gi + (size_t)r * ldi + c2);
    cp_async16(&Pj[r][c2], gj + (size_t)r * ldj + c2);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// --- in-panel update C(ti,tj) -= P(ti) P(tj)^T on 64x64 tiles, tj <= ti, tj inside the outer panel ----
// 4 warps per CTA, each a 32x32 quadrant = 4x4 DMMA tiles, K = 64.
__global__ void __launch_bounds__(128) syrk_dmma_kernel(double* __restrict__ A, long long ld, int k0) {
  extern __shared__ double sm[];
  double(*Pi)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm);
  double(*Pj)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm + CB * CB_LD);
  // tile (ti >= tj) of the strip: grid (row tiles, column tiles of the outer panel)
  const int ti = blockIdx.x, tj = blockIdx.y;
  if (ti < tj) return;
  const int r0 = k0 + CB + ti * CB, c0 = k0 + CB + tj * CB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  load_two_tiles_async(Pi, Pj, A + (size_t)r0 * ld + k0, ld, A + (size_t)c0 * ld + k0, ld);
  const int rb = (w >> 1) * 32, cb = (w & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double2 cv[4][4];  // the C fragments travel while the operands land
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      cv[mi][ni] = *reinterpret_cast<const double2*>(A + (size_t)(r0 + rb + mi * 8 + fr) * ld + c0 + cb + ni * 8 + 2 * fk);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
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
      double2 v = cv[mi][ni];
      v.x -= acc[mi][ni][0];
      v.y -= acc[mi][ni][1];
      *c = v;
    }
}

// --- trailing update with a whole outer panel: C(ti,tj) -= P(ti) P(tj)^T, 128x128 tiles, K = kw ----
// P = A[:, K0 .. K0 + kw), the finished outer panel.  8 warps (2 x 4), warp tile 64 x 32 =
// 8 x 4 DMMA tiles; k runs in chunks of 32 through a two-stage cp.async pipeline.
constexpr int BT = 128;          // C tile
constexpr int BK = 32;           // k chunk
constexpr int BK_LD = BK + 4;    // shared-memory row stride: conflict-free DMMA fragment loads, 16-byte aligned rows
constexpr int kBigThreads = 256;
constexpr size_t kBigSmem = (size_t)2 /*stages*/ * 2 /*Pi, Pj*/ * BT * BK_LD * sizeof(double);
// Tiles: strip_cols > 0: grid (row tiles, strip_cols) = the tile columns [0, strip_cols) (the next outer
// panel, updated first so that its factorisation can start); else a linear id over the triangle of the
// tiles with ti >= tj >= tile_off (the rest, which runs beside that factorisation on a second stream).
__global__ void __launch_bounds__(kBigThreads, 1) syrk_big_dmma_kernel(double* __restrict__ A, long long ld, int K0, int kw,
                                                                       int n_pad, int tile_off, int strip_cols, int n_tiles) {
  extern __shared__ __align__(16) double sm[];
  // rest mode: a persistent grid smaller than the GPU walks the tiles, so that the SMs it leaves
  // free take the latency-bound kernels of the next panel's factorisation (look-ahead)
  for (int p = blockIdx.x; p < (strip_cols > 0 ? (int)gridDim.x : n_tiles); p += gridDim.x) {
  int ti, tj;
  if (strip_cols > 0) {
    ti = blockIdx.x;
    tj = blockIdx.y;
    if (ti < tj) return;
  } else {
    ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    while (ti * (ti + 1) / 2 > p) --ti;
    tj = p - ti * (ti + 1) / 2;
    ti += tile_off;
    tj += tile_off;
  }
  const int base = K0 + kw;
  const int r0 = base + ti * BT, c0 = base + tj * BT;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  auto stage_ptr = [&](int st, int which) { return sm + (size_t)(st * 2 + which) * BT * BK_LD; };
  // one chunk = 2 x (128 rows x 32 doubles): 4096 16-byte pieces, 16 per thread
  auto load_chunk = [&](int st, int kc) {
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int e = tid + it * kBigThreads;      // 0 .. 4095
      const int which = e >> 11, q = e & 2047;   // 2048 pieces per operand
      const int r = q >> 4, c2 = (q & 15) * 2;   // row, first double of the piece
      const int gr = (which == 0 ? r0 : c0) + r;
      double* dst = stage_ptr(st, which) + r * BK_LD + c2;
      if (gr < n_pad) cp_async16(dst, A + (size_t)gr * ld + K0 + kc * BK + c2);
      else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int rb = (w >> 2) * 64, cb = (w & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][4];  // 4 x 4 MMA tiles of 16 x 8
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.0;
  const int nchunk = kw / BK;
  load_chunk(0, 0);
  for (int kc = 0; kc < nchunk; ++kc) {
    if (kc + 1 < nchunk) {
      load_chunk((kc + 1) & 1, kc + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const double* Pi = stage_ptr(kc & 1, 0);
    const double* Pj = stage_ptr(kc & 1, 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 8) {
      double af[4][4], bf[4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const double* r0p = Pi + (rb + mi * 16 + g) * BK_LD + kk + t;
        af[mi][0] = r0p[0];
        af[mi][1] = r0p[8 * BK_LD];
        af[mi][2] = r0p[4];
        af[mi][3] = r0p[8 * BK_LD + 4];
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const double* c0p = Pj + (cb + ni * 8 + g) * BK_LD + kk + t;
        bf[ni][0] = c0p[0];
        bf[ni][1] = c0p[4];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma1688(acc[mi][ni], af[mi], bf[ni]);
    }
    __syncthreads();  // the stage is overwritten by the load issued in the next iteration
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gr = r0 + rb + mi * 16 + g + 8 * h;
      if (gr >= n_pad) continue;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int gc = c0 + cb + ni * 8 + 2 * t;
        if (gc >= n_pad) continue;
        double2* c = reinterpret_cast<double2*>(A + (size_t)gr * ld + gc);
        double2 v = *c;
        v.x -= acc[mi][ni][2 * h];
        v.y -= acc[mi][ni][2 * h + 1];
        *c = v;
      }
    }
  __syncthreads();  // all warps are done with the stages before the next tile's loads
  }
}

// --- trailing update, asynchronous version ---------------------------------------------------
// Same tiles and the same arithmetic as syrk_big_dmma_kernel (every C element still receives ONE
// contribution per launch, so the factor stays bit-reproducible), built so that the DMMA pipe does
// not wait for memory at the two ends of a tile:
//   * the operand chunks of ALL the tiles a persistent CTA walks form one cp.async stream (k chunks of
//     16, STAGES deep, one CTA barrier per chunk): the first chunks of the next tile are in flight
//     while the current tile finishes;
//   * C is never read into the SM: the negated accumulators are staged in shared memory EPI_ROWS
//     rows at a time and handed to the TMA engine as one bulk reduction per row
//     (cp.reduce.async.bulk ... .add.f64: C += (-acc) in L2), which drains under the next tile's MMAs.
constexpr int BK2 = 16;                 // k chunk
constexpr int BK2_LD = BK2 + 4;         // conflict-free DMMA fragment loads, 16-byte aligned rows
constexpr int CS_LD = BT + 8;           // staging row stride: conflict-free 16-byte stores of the accumulator fragments
constexpr int kStage2 = 2 * BT * BK2_LD;  // doubles per stage (both operands)
template <int STAGES, int EPI_ROWS>
constexpr size_t big2_smem() { return (size_t)(STAGES * kStage2 + EPI_ROWS * CS_LD) * sizeof(double); }
__device__ __forceinline__ void tile_of(int p, int nt, int tile_off, int strip_cols, int& ti, int& tj) {
  if (strip_cols > 0) {  // column by column through the strip, rows ti >= tj
    tj = 0;
    while (p >= nt - tj) { p -= nt - tj; ++tj; }
    ti = tj + p;
  } else {               // linear id over the triangle ti >= tj >= tile_off
    ti = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    while (ti * (ti + 1) / 2 > p) --ti;
    tj = p - ti * (ti + 1) / 2 + tile_off;
    ti += tile_off;
  }
}
template <int STAGES, int EPI_ROWS>
__global__ void __launch_bounds__(kBigThreads, 1) syrk_big_async_kernel(double* __restrict__ A, long long ld, int K0, int kw, int n_pad,
                                                                        int nt, int tile_off, int strip_cols, int n_tiles) {
  extern __shared__ __align__(16) double sm[];
  double* cs = sm + STAGES * kStage2;  // [EPI_ROWS][CS_LD]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int base = K0 + kw;
  const int nchunk = kw / BK2;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * nchunk;  // chunks in this CTA's stream
  // the load side of the stream: chunk ld_kc of tile ld_p
  int ld_i = 0, ld_kc = 0, ld_r0 = 0, ld_c0 = 0;
  {
    int ti, tj;
    tile_of(blockIdx.x, nt, tile_off, strip_cols, ti, tj);
    ld_r0 = base + ti * BT;
    ld_c0 = base + tj * BT;
  }
  auto load_next = [&]() {
    if (ld_i < total) {
      double* st = sm + (size_t)(ld_i % STAGES) * kStage2;
      // 2 x (128 rows x 16 doubles) = 2048 16-byte pieces, 8 per thread; 8 threads cover 128 contiguous bytes of a row
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int e = tid + it * kBigThreads;
        const int which = e >> 10, q = e & 1023;
        const int r = q >> 3, c2 = (q & 7) * 2;
        const int gr = (which == 0 ? ld_r0 : ld_c0) + r;
        double* dst = st + (which * BT + r) * BK2_LD + c2;
        if (gr < n_pad) cp_async16(dst, A + (size_t)gr * ld + K0 + ld_kc * BK2 + c2);
        else { dst[0] = 0.0; dst[1] = 0.0; }
      }
      ++ld_i;
      if (++ld_kc == nchunk) {
        ld_kc = 0;
        if (ld_i < total) {
          int ti, tj;
          tile_of(blockIdx.x + (ld_i / nchunk) * gridDim.x, nt, tile_off, strip_cols, ti, tj);
          ld_r0 = base + ti * BT;
          ld_c0 = base + tj * BT;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // an empty group keeps the wait distance constant
  };
#pragma unroll
  for (int q = 0; q < STAGES - 1; ++q) load_next();
  const int rb = (w >> 2) * 64, cb = (w & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  int i = 0;
  for (int mt = 0; mt < my_tiles; ++mt) {
    double acc[4][4][4];  // 4 x 4 MMA tiles of 16 x 8
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mi][ni][q] = 0.0;
    for (int kc = 0; kc < nchunk; ++kc, ++i) {
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");  // chunk i has landed (this thread's pieces)
      __syncthreads();  // ... everybody's; and every warp is done with chunk i - 1, whose stage the next load overwrites
      load_next();      // chunk i + STAGES - 1
      const double* Pi = sm + (size_t)(i % STAGES) * kStage2;
      const double* Pj = Pi + BT * BK2_LD;
#pragma unroll
      for (int kk = 0; kk < BK2; kk += 8) {
        double af[4][4], bf[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
          const double* r0p = Pi + (rb + mi * 16 + g) * BK2_LD + kk + t;
          af[mi][0] = r0p[0];
          af[mi][1] = r0p[8 * BK2_LD];
          af[mi][2] = r0p[4];
          af[mi][3] = r0p[8 * BK2_LD + 4];
        }
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const double* c0p = Pj + (cb + ni * 8 + g) * BK2_LD + kk + t;
          bf[ni][0] = c0p[0];
          bf[ni][1] = c0p[4];
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) dmma1688(acc[mi][ni], af[mi], bf[ni]);
      }
    }
    // epilogue: C(tile) += -acc through the staging rows
    int ti, tj;
    tile_of(blockIdx.x + mt * gridDim.x, nt, tile_off, strip_cols, ti, tj);
    const int r0 = base + ti * BT, c0 = base + tj * BT;
    const unsigned row_bytes = (unsigned)(min(BT, n_pad - c0) * (int)sizeof(double));
#pragma unroll
    for (int ph = 0; ph < BT / EPI_ROWS; ++ph) {
      if (tid < EPI_ROWS) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the engine has read the previous rows
      __syncthreads();
      if ((w >> 2) == (ph * EPI_ROWS) / 64) {
#pragma unroll
        for (int mm = 0; mm < EPI_ROWS / 16; ++mm) {
          const int mi = ((ph * EPI_ROWS) % 64) / 16 + mm;
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
              *reinterpret_cast<double2*>(cs + (mm * 16 + g + 8 * h) * CS_LD + cb + ni * 8 + 2 * t) =
                  make_double2(-acc[mi][ni][2 * h], -acc[mi][ni][2 * h + 1]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the bulk engine
      }
      __syncthreads();
      if (tid < EPI_ROWS) {
        const int gr = r0 + ph * EPI_ROWS + tid;
        if (gr < n_pad)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(
                           __cvta_generic_to_global(A + (size_t)gr * ld + c0)),
                       "r"((unsigned)__cvta_generic_to_shared(cs + tid * CS_LD)), "r"(row_bytes)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (tid < EPI_ROWS) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // complete (and visible) before the CTA retires
}

// --- panel solve as a GEMM: X = A_tile Linv^T on 64x64 tiles (in place), DMMA ----------------
__global__ void __launch_bounds__(128) trsm_dmma_kernel(double* __restrict__ A, long long ld, int k0,
                                                        const double* __restrict__ Linv) {
  extern __shared__ double sm[];
  double(*Pi)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm);
  double(*Pj)[CB_LD] = reinterpret_cast<double(*)[CB_LD]>(sm + CB * CB_LD);
  const int r0 = k0 + CB + blockIdx.x * CB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  load_two_tiles_async(Pi, Pj, A + (size_t)r0 * ld + k0, ld, Linv, CB);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
// y_k = Linv_k^T w_k with the block inverse kept from the factorisation (no 64-step substitution
// chain): every CTA forms y_k redundantly, CTA 0 publishes it, and CTA b folds y_k into
// w[64 b .. 64 b + 63] (< k0).  Only the leading nv unknowns of the last block are real.
__global__ void __launch_bounds__(256) backsolve_step_kernel(double* __restrict__ A, long long ld, int k0,
                                                             int n, int rhs_row, double* __restrict__ y,
                                                             const double* __restrict__ Linv) {
  __shared__ double wv[CB], yv[CB];
  const int tid = threadIdx.x;
  const int nv = min(CB, n - k0);
  if (tid < CB) wv[tid] = tid < nv ? A[(size_t)rhs_row * ld + k0 + tid] : 0.0;
  __syncthreads();
  {
    // y_r = sum_{c >= r} Linv[c][r] w_c: four lanes per r
    const int r = tid >> 2, part = tid & 3;
    double acc = 0.0;
    for (int c = r + part; c < nv; c += 4) acc += Linv[c * CB + r] * wv[c];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) yv[r] = r < nv ? acc : 0.0;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < nv) y[k0 + tid] = yv[tid];
  // w[j] -= sum_r L[k0 + r][j] y_r for 64 columns j per CTA: four groups of 16 rows each, so that
  // every thread has 16 independent loads in flight instead of a 64-long chain
  __shared__ double part[4][CB];
  const int jl = tid & 63, grp = tid >> 6;
  const int j = blockIdx.x * CB + jl;
  double s = 0.0;
  if (j < k0) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int r = grp * 16 + q;
      if (r < nv) s += A[(size_t)(k0 + r) * ld + j] * yv[r];
    }
  }
  part[grp][jl] = s;
  __syncthreads();
  if (grp == 0 && j < k0) A[(size_t)rhs_row * ld + j] -= (part[0][jl] + part[1][jl]) + (part[2][jl] + part[3][jl]);
}

// --- the whole of L^T y = w in ONE launch ------------------------------------------------------
// CTA j owns the unknowns of 64-block j: it folds y_k (k from the last block down to j + 1) into its right-hand
// side, then forms y_j = Linv_j^T w_j.  The CTAs are chained through y itself: y is preset to an all-ones bit
// pattern (a NaN no computation produces) and a reader polls the 64 values it needs until they are numbers, so
// the dependent chain is one L2 round trip + 16 FMAs + one 64 x 64 mat-vec per block instead of a launch.  The L
// block of the NEXT step is already in registers when y_k arrives.  Launched cooperatively (all CTAs resident;
// producers carry the lowest block indices); summation order is fixed, so the result is reproducible.
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
constexpr int kChainSpinLimit = 1 << 22;  // ~ seconds; a reader that gives up raises *fail instead of hanging the GPU
__global__ void __launch_bounds__(256, 2) backsolve_chain_kernel(const double* __restrict__ A, long long ld, int n, int rhs_row,
                                                                 double* y, const double* __restrict__ linv, double* __restrict__ fail) {
  __shared__ __align__(16) double Ls[CB * CB];  // Linv_j
  __shared__ double yv[2][CB], part[4][CB], wv[CB];
  const int nblk = gridDim.x;
  const int j = nblk - 1 - (int)blockIdx.x;
  const int tid = threadIdx.x, c = tid & 63, grp = tid >> 6;
  const int nv_j = min(CB, n - j * CB);
#pragma unroll
  for (int it = 0; it < CB * CB / 2 / 256; ++it) {
    const int e = tid + it * 256;
    cp_async16(Ls + 2 * e, linv + (size_t)j * CB * CB + 2 * e);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const double w0 = (tid < nv_j) ? A[(size_t)rhs_row * ld + j * CB + tid] : 0.0;
  auto load_blk = [&](double(&dst)[16], int k) {
    const int nv = min(CB, n - k * CB);
    const double* src = A + (size_t)(k * CB + grp * 16) * ld + j * CB + c;
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[q] = (grp * 16 + q < nv) ? src[(size_t)q * ld] : 0.0;
  };
  double s = 0.0;
  auto step = [&](const double(&cur)[16], double(&nxt)[16], int k) {
    if (k - 1 > j) load_blk(nxt, k - 1);
    if (tid < CB) {
      const double* src = y + (size_t)k * CB + tid;
      double v = ld_volatile_f64(src);
      int spins = 0;
      while (__double_as_longlong(v) == -1ll && ++spins < kChainSpinLimit) v = ld_volatile_f64(src);
      if (spins >= kChainSpinLimit) { *fail = 1.0; v = 0.0; }
      yv[k & 1][tid] = v;
    }
    __syncthreads();
    const double* yk = yv[k & 1] + grp * 16;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += cur[q] * yk[q];
  };
  double Lb0[16], Lb1[16];
  int k = nblk - 1;
  if (k > j) load_blk(Lb0, k);
  while (k > j) {
    step(Lb0, Lb1, k);
    if (--k <= j) break;
    step(Lb1, Lb0, k);
    --k;
  }
  part[grp][c] = s;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (tid < CB) wv[tid] = w0 - ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
  __syncthreads();
  {
    // y_r = sum_{cc >= r} Linv[cc][r] w_cc: four lanes per r
    const int r = tid >> 2, p4 = tid & 3;
    double acc = 0.0;
    for (int cc = r + p4; cc < nv_j; cc += 4) acc += Ls[cc * CB + r] * wv[cc];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (p4 == 0) {
      const double v = r < nv_j ? acc : 0.0;
      asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(y + (size_t)j * CB + r), "d"(v) : "memory");
    }
  }
}

struct DenseCholesky {
  static constexpr size_t kTileSmem = 2 * CB * CB_LD * sizeof(double);
  static cudaError_t init() {
    cudaError_t e = cudaFuncSetAttribute(trsm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(syrk_big_async_kernel<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big2_smem<3, 64>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(syrk_big_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem);
  }
  static void launch_big(int grid, cudaStream_t st, double* A, long long ld, int K0, int kw, int n_pad, int nt, int tile_off, int strip_cols,
                         int n_tiles) {
    syrk_big_async_kernel<3, 64><<<grid, kBigThreads, big2_smem<3, 64>(), st>>>(A, ld, K0, kw, n_pad, nt, tile_off, strip_cols, n_tiles);
  }
  // Second stream + events for the look-ahead: after outer panel P is factored, the update of the
  // NEXT panel's columns (strip) stays on the caller's stream, followed by that panel's
  // factorisation, while the update of everything beyond (rest) runs on `aux` (lower priority,
  // so the latency-bound factorisation kernels get SM slots first).
  struct LookAhead {
    cudaStream_t aux = nullptr;
    int chain_capacity = -1;      // CTAs of backsolve_chain_kernel the GPU holds at once (-1: not asked yet)
    int variant = 1;              // trailing update: 0 synchronous epilogue (round 1), 1 asynchronous
    int nb = 256;                 // outer panel width (a multiple of 128)
    int free_sms = 8;             // SMs the persistent rest update leaves to the factorisation chain
    int sms = 148;
    int rest_ctas = 140;          // persistent CTAs of the rest update (one per SM): the other SMs serve the factorisation chain
    std::vector<cudaEvent_t> ev;  // [2 P]: panel P factored, [2 P + 1]: rest(P) done
    cudaError_t ensure(int panels) {
      if (!aux) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        rest_ctas = std::max(1, sms - free_sms);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const cudaError_t e = cudaStreamCreateWithPriority(&aux, cudaStreamNonBlocking, lo);
        if (e != cudaSuccess) return e;
      }
      while ((int)ev.size() < 2 * panels) {
        cudaEvent_t x;
        const cudaError_t e = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        ev.push_back(x);
      }
      return cudaSuccess;
    }
    ~LookAhead() {
      for (cudaEvent_t x : ev) cudaEventDestroy(x);
      if (aux) cudaStreamDestroy(aux);
    }
  };
  // factor rows/cols [0, n_pad); linv = (n_pad / 64) x 64 x 64: the inverses of the diagonal blocks' factors;
  // returns the number of launches
  static int factor(double* A, long long ld, int n_pad, double* linv, double* fail, cudaStream_t st, LookAhead& la) {
    int launches = 0;
    const int NB = la.nb;
    const int panels = (n_pad + NB - 1) / NB;
    // without the second stream (creation failed) everything runs in order on the caller's stream
    const bool two_streams = la.ensure(panels) == cudaSuccess;
    if (!two_streams) cudaGetLastError();
    cudaStream_t aux = two_streams ? la.aux : st;
    int last_rest = -1;  // last panel whose rest update was launched on aux
    for (int P = 0; P < panels; ++P) {
      const int K0 = P * NB;
      const int kw = std::min(NB, n_pad - K0);
      for (int k0 = K0; k0 < K0 + kw; k0 += CB) {
        double* linv_k = linv + (size_t)(k0 / CB) * CB * CB;  // kept for the back-substitution
        potrf_inv_kernel<<<1, 256, kPotrfSmem, st>>>(A, ld, k0, linv_k, fail);
        ++launches;
        const int rem = (n_pad - k0 - CB) / CB;          // row tiles below the diagonal block
        if (rem > 0) {
          trsm_dmma_kernel<<<rem, 128, kTileSmem, st>>>(A, ld, k0, linv_k);
          ++launches;
          const int ntj = (K0 + kw - k0 - CB) / CB;      // column tiles still inside the outer panel
          if (ntj > 0) {
            syrk_dmma_kernel<<<dim3(rem, ntj), 128, kTileSmem, st>>>(A, ld, k0);
            ++launches;
          }
        }
      }
      const int left = n_pad - K0 - kw;                  // trailing rows / columns
      if (left <= 0) break;
      const int nt = (left + BT - 1) / BT;
      const int strip = std::min(nt, NB / BT);           // tile columns of the next outer panel
      if (two_streams) cudaEventRecord(la.ev[2 * P], st);  // panel P is factored
      // the strip reads and writes columns that rest(P - 1) wrote
      if (two_streams && last_rest >= 0) cudaStreamWaitEvent(st, la.ev[2 * last_rest + 1], 0);
      if (la.variant == 0) {
        syrk_big_dmma_kernel<<<dim3(nt, strip), kBigThreads, kBigSmem, st>>>(A, ld, K0, kw, n_pad, 0, strip, 0);
      } else {
        const int stiles = nt * strip - strip * (strip - 1) / 2;
        launch_big(std::min(stiles, la.sms), st, A, ld, K0, kw, n_pad, nt, 0, strip, stiles);
      }
      ++launches;
      if (nt > strip) {
        const int m = nt - strip;
        if (two_streams) cudaStreamWaitEvent(aux, la.ev[2 * P], 0);
        const int tiles = m * (m + 1) / 2;
        if (la.variant == 0)
          syrk_big_dmma_kernel<<<std::min(tiles, la.rest_ctas), kBigThreads, kBigSmem, aux>>>(A, ld, K0, kw, n_pad, strip, 0, tiles);
        else
          launch_big(std::min(tiles, la.rest_ctas), aux, A, ld, K0, kw, n_pad, nt, strip, 0, tiles);
        if (two_streams) cudaEventRecord(la.ev[2 * P + 1], aux);
        last_rest = P;
        ++launches;
      }
    }
    if (two_streams && last_rest >= 0) cudaStreamWaitEvent(st, la.ev[2 * last_rest + 1], 0);  // join
    return launches;
  }
  // solves L^T y = w for the leading n unknowns; w is row rhs_row (== n) and is destroyed
  // y must hold ceil(n / 64) * 64 doubles
  static int backsolve(double* A, long long ld, int n, int rhs_row, double* y, const double* linv, double* fail, cudaStream_t st, LookAhead& la,
                       bool chained = true) {
    int launches = 0;
    const int nblk = (n + CB - 1) / CB;
    if (la.chain_capacity < 0) {
      int per_sm = 0, dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      la.chain_capacity = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, backsolve_chain_kernel, 256, 0) == cudaSuccess ? per_sm * sms : 0;
    }
    if (chained && nblk <= la.chain_capacity) {
      cudaMemsetAsync(y, 0xFF, (size_t)nblk * CB * sizeof(double), st);
      const double* Ac = A;
      void* args[] = {(void*)&Ac, (void*)&ld, (void*)&n, (void*)&rhs_row, (void*)&y, (void*)&linv, (void*)&fail};
      if (cudaLaunchCooperativeKernel((void*)backsolve_chain_kernel, dim3(nblk), dim3(256), args, 0, st) == cudaSuccess) return 2;
      cudaGetLastError();  // fall through to the stepwise path
    }
    for (int k0 = ((n - 1) / CB) * CB; k0 >= 0; k0 -= CB) {
      const int grid = k0 > 0 ? (k0 + CB - 1) / CB : 1;
      backsolve_step_kernel<<<grid, 256, 0, st>>>(A, ld, k0, n, rhs_row, y, linv + (size_t)(k0 / CB) * CB * CB);
      ++launches;
    }
    return launches;
  }
};

}  // namespace ars
