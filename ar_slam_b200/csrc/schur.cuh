// schur.cuh -- kernels (3)/(4): Schur complement of the E poses, dense
// reduced system assembly, back-substitution and the LM step vectors.
//
// Replaces Ceres' SchurEliminator::Eliminate / BackSubstitute and
// LevenbergMarquardtStrategy's damping (reached from ceres::Solve,
// reference ar_slam/src/ar_slam_util.cpp:1011-1015).
//
// Scaled, damped normal equations (Jacobi scaling sigma, LM diagonal D):
//   Ht = Sig H Sig + D^2,  gt = Sig g,  solve Ht y = gt,  delta = -Sig y.
// Reduced system ordering: F pose f -> rows 6f..6f+5, focal length -> row
// cam_row = 6 n_f, right-hand side -> extra row rhs_row = cam_row + 1 of the
// same lower-triangular array, so that the Cholesky sweep also performs the
// forward substitution.
#pragma once
#include "kernels.cuh"

namespace ars {

struct SchurArgs {
  int n_e, plane;
  const int32_t* e_off;   // [n_e] first block of each E pose's segment (E-sorted order)
  const int32_t* e_end;   // [n_e] one past its last block (== e_off + 1 when the segments are in index order)
  const int32_t* f_idx;   // [n_blk]   F pose per E-sorted block
  const int32_t* pair_off; // [n_blk]  number of (partner, block) pairs before each block (sparse target)
  const double* HE;       // [n_e][NV]
  const double* HEx;      // [n_e][NVX] l1, l2 borders (radial model), else null
  const double* W;        // 36 planes
  const double* sig_e;    // [6 n_e]
  double radius, inv_radius, min_diag, max_diag;  // inv_radius = 1 / radius (host)
  double* Z;              // [n_e][8]: z = Ht_ee^-1 sig_e g_e (6), ok flag, pad
  double* YB;             // [n_e][6 NK]: Ht_ee^-1 sig_e H_e,intrinsic_q
  double* seg_cam;        // [grid][12] CTA partials of M = hk^T yb (ff, f l1, f l2, l1l1, l1l2, l2l2) | hk^T z (3) | failed | max |g_e| | 0
  const unsigned char* e_const;  // [n_e] != 0: E pose held constant (left out of the gradient norm), or null
  // second launch (products whose partner sits in another CTA): the CTAs that hold a part of a segment that leaves
  // them, found once per problem; with equally long aligned segments (every capture sees 8 tags) there are none
  const int32_t* straddle_ctas;  // [n_straddle]
  int n_straddle;
};

// CTAs (of kSchurCtaBlocks consecutive E-sorted blocks) that hold a part of a segment leaving them
constexpr int kSchurCtaBlocks = 128;
__global__ void straddle_ctas_kernel(int n_blk, const int32_t* __restrict__ e_idx, const int32_t* __restrict__ e_off,
                                     const int32_t* __restrict__ e_end, int32_t* __restrict__ list, int* __restrict__ count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int first = c * kSchurCtaBlocks;
  if (first >= n_blk) return;
  const int last = min(first + kSchurCtaBlocks, n_blk) - 1;
  const bool open_head = e_off[e_idx[first]] < first;
  const bool open_tail = e_end[e_idx[last]] > first + kSchurCtaBlocks;
  if (open_head || open_tail) list[atomicAdd(count, 1)] = c;
}

// FP64 add to global memory without a return value.  atomicAdd through a pointer whose address
// space the compiler cannot prove (e.g. one that went through a shuffle) becomes a generic ATOM
// that returns the old value; consecutive ones then serialise on the destination register.
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

// decode p -> (i <= j) with p = j (j + 1) / 2 + i
__device__ __forceinline__ void tri_decode(int p, int& i, int& j) {
  j = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
  while ((j + 1) * (j + 2) / 2 <= p) ++j;
  while (j * (j + 1) / 2 > p) --j;
  i = p - j * (j + 1) / 2;
}

// scaled borders of the l1, l2 columns of an E pose (radial model)
__device__ __forceinline__ void load_scaled_Ex(const SchurArgs& a, int e, const double s[6], double hk1[6], double hk2[6]) {
  const double* rx = a.HEx + (size_t)e * NVX;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    hk1[i] = rx[i] * s[i];
    hk2[i] = rx[6 + i] * s[i];
  }
}

__device__ __forceinline__ void load_scaled_E(const SchurArgs& a, int e, double L[36], double g[6],
                                              double hk[6], double s[6]) {
  const double* rec = a.HE + (size_t)e * NV;
  const double inv_radius = a.inv_radius;
#pragma unroll
  for (int i = 0; i < 6; ++i) s[i] = a.sig_e[6 * (size_t)e + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) {
      const double h = rec[tri6(i, j)] * s[i] * s[j];
      L[i * 6 + j] = h;
      L[j * 6 + i] = h;
    }
    const double d = fmin(fmax(L[i * 6 + i], a.min_diag), a.max_diag);
    L[i * 6 + i] += d * inv_radius;
    g[i] = rec[21 + i] * s[i];
    hk[i] = rec[27 + i] * s[i];
  }
}

// Where the Schur terms sum W~_i^T Y_j land: the dense lower-triangular array
// (DenseTarget) or the block-sparse value array (SparseTarget, pcg.cuh).
struct DenseTarget {
  double* S;
  long long ld;
  int cam_row, rhs_row;
  __device__ __forceinline__ void add_border(int f, int c, double b0, double b1) const {
    red_add_f64(S + (size_t)cam_row * ld + 6 * f + c, b0);
    red_add_f64(S + (size_t)rhs_row * ld + 6 * f + c, b1);
  }
  __device__ __forceinline__ void add_border_x(int f, int c, double bl1, double bl2) const {
    red_add_f64(S + (size_t)(cam_row + 1) * ld + 6 * f + c, bl1);
    red_add_f64(S + (size_t)(cam_row + 2) * ld + 6 * f + c, bl2);
  }
  // M[r][c] = element (6 fi + r, 6 fj + c), fi <= fj, of sum W~^T Y; stored in the lower triangle
  __device__ __forceinline__ double* block(int fi, int fj, long long /*pair*/) const { return S + (size_t)(6 * fj) * ld + 6 * fi; }
  // address of element e = 6 * row + col of a 6x6 block
  __device__ __forceinline__ double* elem(double* blk, int e) const { return blk + (size_t)(e / 6) * ld + (e % 6); }
};

// number of (block, partner) products thread j of a k-block segment owns: partners are
// (j + d) mod k for d = 0 .. k/2, the d = k/2 ring of an even k being split between the halves
__host__ __device__ inline int schur_pairs_of(int j, int k) {
  return (k & 1) ? (k + 1) / 2 : k / 2 + (j < k / 2 ? 1 : 0);
}

// forward substitution with the reciprocal-pivot factor: v = L^-1 v
__device__ __forceinline__ void chol6_forward(const double L[36], double b[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < i) s -= L[i * 6 + k] * b[k];
    b[i] = s * L[i * 6 + i];
  }
}
__device__ __forceinline__ void chol6_backward(const double L[36], double b[6]) {
#pragma unroll
  for (int ii = 0; ii < 6; ++ii) {
    const int i = 5 - ii;
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > i) s -= L[k * 6 + i] * b[k];
    b[i] = s * L[i * 6 + i];
  }
}

constexpr int kSchurThreads = 128;
constexpr int kSchurStageLd = 38;  // staged 6x6 product per thread: 36 doubles, 16-byte aligned rows (bulk reduction source)
constexpr size_t kSchurSmem = (size_t)(kSchurThreads * 37 + 4 * 32 * kSchurStageLd) * sizeof(double);

// One TMA bulk reduction adds a staged, contiguous block of doubles to global memory
// (cp.reduce.async.bulk ... .add.f64): the 36 elements of a Schur product travel as one request
// instead of 36 per-lane FP64 reductions.  src: shared memory, 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_reduce_add_f64(double* dst_global, const double* src_shared, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy stores that staged the block
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(dst_global)),
               "r"((unsigned)__cvta_generic_to_shared(src_shared)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One thread per residual block (E-sorted order).  With Ht_ee = L L^T the Schur term of a
// segment is sum_ij V_i^T V_j, V_j = L^-1 (sig_e W_j).  Thread j
//   * rebuilds the segment's damped 6x6 block and factors it in registers (batched 6x6
//     Cholesky; the redundancy inside a segment is cheaper than broadcasting 21 + 12 doubles),
//   * computes V_j (its 36 W loads are issued before the factorisation) and publishes it in
//     shared memory,
//   * forms its share of the k (k + 1) / 2 products V_a^T V_b -- partners (j + d) mod k, so
//     every lane of the segment carries the same load -- reading the partner's V from shared
//     memory, stages each 6x6 result in shared memory, and the warp adds it to the reduced
//     system with one FP64 reduction per lane on consecutive addresses (coalesced at L2).
// A partner that sits in a neighbouring CTA (a segment cut by a CTA boundary, or longer than a
// CTA) is not in shared memory.  Those products are left to a second launch of the same kernel
// with STRADDLE = true, which rebuilds the partner's V column by column from W (same L); the
// main launch then does not need L after V and fits three CTAs per SM.
// BULK (sparse target only: its blocks are 36 contiguous doubles): every thread hands its staged product to
// the TMA engine as ONE bulk reduction instead of the warp adding it element by element.
template <typename Target, int NK, bool STRADDLE, bool BULK = false>
__global__ void __launch_bounds__(kSchurThreads, STRADDLE ? 1 : 3)
schur_eliminate_kernel(const SchurArgs a, const Target t, int n_blk, const int32_t* __restrict__ e_idx) {
  extern __shared__ __align__(16) double schur_sm[];
  double(*Vs)[37] = reinterpret_cast<double(*)[37]>(schur_sm);                       // [128][37]
  double(*stage)[32][kSchurStageLd] = reinterpret_cast<double(*)[32][kSchurStageLd]>(schur_sm + kSchurThreads * 37);  // [4][32][38]
  const int cta = STRADDLE ? a.straddle_ctas[blockIdx.x] : (int)blockIdx.x;
  const int pos = cta * blockDim.x + threadIdx.x;
  const int cta0 = cta * blockDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool valid = pos < n_blk;
  const int e = valid ? e_idx[pos] : 0;
  const int beg = valid ? a.e_off[e] : 0;
  const int k = valid ? a.e_end[e] - beg : 0;
  const int j = valid ? pos - beg : 0;
  // the second launch only concerns segments that leave their CTA
  if (STRADDLE) valid = valid && (beg < cta0 || beg + k > cta0 + kSchurThreads);
  double L[36], V[36], s[6];
  int fj = 0;
  const size_t ps = a.plane;
  // camera terms of the segments (written by their first thread), parked in the staging area
  // until the CTA reduces them below
  double(*cs)[12] = reinterpret_cast<double(*)[12]>(schur_sm + kSchurThreads * 37);
  if (!STRADDLE && !(valid && j == 0)) {
#pragma unroll
    for (int i = 0; i < 12; ++i) cs[threadIdx.x][i] = 0.0;
  }
  if (valid) {
#pragma unroll
    for (int q = 0; q < 36; ++q) V[q] = a.W[(size_t)q * ps + pos];  // in flight during the factorisation
    double zl[6], ybl[6], hk[6];
    double hk1[6], hk2[6], ybl1[6], ybl2[6];  // radial model: l1, l2 borders
    load_scaled_E(a, e, L, zl, hk, s);
    const bool ok = chol6(L);
    fj = a.f_idx[pos];
    if (!STRADDLE) {
#pragma unroll
      for (int i = 0; i < 6; ++i) ybl[i] = hk[i];
      chol6_forward(L, zl);
      chol6_forward(L, ybl);
      if (NK == 3) {
        load_scaled_Ex(a, e, s, hk1, hk2);
#pragma unroll
        for (int i = 0; i < 6; ++i) { ybl1[i] = hk1[i]; ybl2[i] = hk2[i]; }
        chol6_forward(L, ybl1);
        chol6_forward(L, ybl2);
      }
      if (j == 0) {
        double z[6], yb[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { z[i] = zl[i]; yb[i] = ybl[i]; }
        chol6_backward(L, z);
        chol6_backward(L, yb);
        double* zo = a.Z + 8 * (size_t)e;
        double* sg = cs[threadIdx.x];
        double m00 = 0.0, v0 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          zo[i] = z[i];
          a.YB[6 * NK * (size_t)e + i] = yb[i];
          m00 += hk[i] * yb[i];
          v0 += hk[i] * z[i];
        }
        zo[6] = ok ? 0.0 : 1.0;
        zo[7] = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) sg[i] = 0.0;
        sg[0] = m00; sg[6] = v0; sg[9] = ok ? 0.0 : 1.0;
        if (!(a.e_const && a.e_const[e])) {  // max |g| of this E pose (unscaled), for the gradient convergence test
          const double* rec = a.HE + (size_t)e * NV;
          double gm = 0.0;
#pragma unroll
          for (int i = 0; i < 6; ++i) gm = fmax(gm, fabs(rec[21 + i]));
          sg[10] = gm;
        }
        if (NK == 3) {
          double yb1[6], yb2[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) { yb1[i] = ybl1[i]; yb2[i] = ybl2[i]; }
          chol6_backward(L, yb1);
          chol6_backward(L, yb2);
          double m01 = 0, m02 = 0, m11 = 0, m12 = 0, m22 = 0, v1 = 0, v2 = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            a.YB[18 * (size_t)e + 6 + i] = yb1[i];
            a.YB[18 * (size_t)e + 12 + i] = yb2[i];
            m01 += hk[i] * yb1[i]; m02 += hk[i] * yb2[i];
            m11 += hk1[i] * yb1[i]; m12 += hk1[i] * yb2[i]; m22 += hk2[i] * yb2[i];
            v1 += hk1[i] * z[i]; v2 += hk2[i] * z[i];
          }
          sg[1] = m01; sg[2] = m02; sg[3] = m11; sg[4] = m12; sg[5] = m22; sg[7] = v1; sg[8] = v2;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double col[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) col[i] = V[i * 6 + c] * s[i];
      chol6_forward(L, col);
#pragma unroll
      for (int i = 0; i < 6; ++i) V[i * 6 + c] = col[i];
      if (!STRADDLE) {
        double b0 = 0.0, b1 = 0.0, bl1 = 0.0, bl2 = 0.0;  // (sig_e W)^T Ht^-1 h = V^T (L^-1 h)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          Vs[threadIdx.x][i * 6 + c] = col[i];
          b0 += col[i] * ybl[i];
          b1 += col[i] * zl[i];
          if (NK == 3) { bl1 += col[i] * ybl1[i]; bl2 += col[i] * ybl2[i]; }
        }
        t.add_border(fj, c, b0, b1);
        if (NK == 3) t.add_border_x(fj, c, bl1, bl2);
      }
    }
  }
  if (!STRADDLE) {
    __syncthreads();
    double cm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) cm[i] = (NK == 3 || i == 0 || i == 6 || i == 9) ? warp_sum(cs[threadIdx.x][i]) : 0.0;
    cm[10] = warp_max(cs[threadIdx.x][10]);
    cta_partial<12, false, kSchurThreads / 32, 10>(cm, a.seg_cam, schur_sm + kSchurThreads * 37 + kSchurThreads * 12);
    __syncthreads();  // the staging area is reused by the products
  }
  const int my_pairs = valid ? schur_pairs_of(j, k) : 0;
  int dmax = my_pairs;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  for (int d = 0; d < dmax; ++d) {
    bool active = d < my_pairs;
    double* blk = nullptr;
    if (active) {
      int i2 = j + d;
      if (i2 >= k) i2 -= k;
      const int ppos = beg + i2;
      const bool in_cta = ppos >= cta0 && ppos < cta0 + kSchurThreads;
      active = STRADDLE ? !in_cta : in_cta;
      if (active) {
        const int fp = a.f_idx[ppos];
        // order the pair so that the block lands in the lower triangle: row = larger F pose
        const bool own_first = fj <= fp;
        const int fa = own_first ? fj : fp, fb = own_first ? fp : fj;
        const bool diag = fa == fb, twice = diag && d != 0;
        blk = t.block(fa, fb, (a.pair_off ? (long long)a.pair_off[pos] : 0) + d);
        double* st = stage[wid][lane];
        if (BULK) bulk_wait_read();  // the engine has read the product staged in the previous round
        // m1 = V^T P, P the partner's V.  The product wanted is M = V_a^T V_b (element
        // (6 fa + r, 6 fb + c)), stored transposed in the lower block (fb, fa): with the own block
        // first that is m1^T, with the partner first it is m1 itself; on the diagonal (same F
        // pose) m1 is stored as is.
        const bool transpose = own_first && !diag;
        const double* Pp = Vs[STRADDLE ? 0 : ppos - cta0];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double pc[6];  // column c of P
          if (STRADDLE) {
#pragma unroll
            for (int i = 0; i < 6; ++i) pc[i] = a.W[(size_t)(i * 6 + c) * ps + ppos] * s[i];
            chol6_forward(L, pc);
          } else {
#pragma unroll
            for (int m = 0; m < 6; ++m) pc[m] = Pp[m * 6 + c];
          }
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            double m1 = 0.0;
#pragma unroll
            for (int m = 0; m < 6; ++m) m1 += V[m * 6 + r] * pc[m];
            st[transpose ? c * 6 + r : r * 6 + c] = m1;
          }
        }
        if (twice) {  // a tag seen twice by one capture: the two cross products M + M^T share a diagonal block
#pragma unroll
          for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = r; c < 6; ++c) {
              const double v = r == c ? 2.0 * st[r * 6 + c] : st[r * 6 + c] + st[c * 6 + r];
              st[r * 6 + c] = v;
              st[c * 6 + r] = v;
            }
        }
        if (BULK) bulk_reduce_add_f64(blk, st, 36 * sizeof(double));
      }
    }
    if (BULK) continue;
    __syncwarp();
    const unsigned m = __ballot_sync(0xffffffffu, active);
    // lanes 0..31 of the warp add elements 0..31 of every staged block; the 4-element tails
    // (elements 32..35) of eight blocks at a time share one more warp-wide reduction
    for (int base = 0; base < 32; base += 8) {
      if (!((m >> base) & 0xffu)) continue;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int src = base + q;
        double* dst = reinterpret_cast<double*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(blk), src));
        if ((m >> src) & 1u) red_add_f64(t.elem(dst, lane), stage[wid][src][lane]);
      }
      const int src = base + (lane >> 2);
      double* dst = reinterpret_cast<double*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(blk), src));
      if ((m >> src) & 1u) red_add_f64(t.elem(dst, 32 + (lane & 3)), stage[wid][src][32 + (lane & 3)]);
    }
    __syncwarp();
  }
  if (BULK) bulk_wait_all();  // the reductions are complete (and visible) before the thread exits
}

// One CTA closes the elimination: column sums of the CTA partials (cam_minus[0..9], fixed order), the
// maximum of the E poses' gradient entries, and the failure flag of the linear solve that starts here.
__global__ void __launch_bounds__(1024) schur_finish_kernel(const double* __restrict__ seg_cam, int n, double* __restrict__ cam_minus,
                                                            double* __restrict__ sc) {
  __shared__ double scratch[(1024 / 12) * 12];
  reduce_partials<12, false, 1024, 10>(seg_cam, n, cam_minus, scratch);
  __syncthreads();
  if (threadIdx.x == 0) {
    sc[16] = cam_minus[10];  // LmScalars::gmax_e
    sc[12] = 0.0;            // LmScalars::chol_fail (the solvers raise it; the E-pose failures travel in cam_minus[9])
  }
}

// The dense reduced system is only ever read on and below the diagonal (and inside the 128 x 128 diagonal tiles
// of the factorisation): the two full-matrix passes touch 64 x 32 tiles of that region only.
constexpr int kDenseTileCols = 64, kDenseTileRows = 32;
// zero the rows [0, rows) up to the end of the 128-wide tile that holds the diagonal
__global__ void __launch_bounds__(256) dense_zero_lower_kernel(double* __restrict__ S, long long ld, int rows) {
  const int j0 = blockIdx.x * kDenseTileCols, i0 = blockIdx.y * kDenseTileRows;
  if (j0 >= ((i0 + kDenseTileRows - 1) / 128 + 1) * 128) return;
  const int j = j0 + (threadIdx.x & 31) * 2;
  const int ib = i0 + (threadIdx.x >> 5) * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ib + q;
    if (i < rows && j < ld) *reinterpret_cast<double2*>(S + (size_t)i * ld + j) = make_double2(0.0, 0.0);
  }
}
// The same region as a contiguous buffer, for the multi-GPU sum: 128-row bands, band t holding 128 (t + 1) columns.
__host__ __device__ inline size_t dense_packed_offset(int band) { return (size_t)(128 * 128) * band * (band + 1) / 2; }
inline size_t dense_packed_count(int n_pad) { return dense_packed_offset((n_pad + 127) / 128); }
template <bool UNPACK>
__global__ void __launch_bounds__(256) dense_pack_kernel(double* __restrict__ S, long long ld, int rows, double* __restrict__ packed) {
  const int j0 = blockIdx.x * kDenseTileCols, i0 = blockIdx.y * kDenseTileRows;
  const int band = i0 / 128, width = (band + 1) * 128;
  if (j0 >= width) return;
  const int j = j0 + (threadIdx.x & 31) * 2;
  const int ib = i0 + (threadIdx.x >> 5) * 4;
  if (j >= ld) return;
  double* pk = packed + dense_packed_offset(band) + (size_t)(ib - band * 128) * width + j;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (ib + q >= rows) break;
    double2* a = reinterpret_cast<double2*>(S + (size_t)(ib + q) * ld + j);
    double2* b = reinterpret_cast<double2*>(pk + (size_t)q * width);
    if (UNPACK) *a = *b; else *b = *a;
  }
}
// S <- -sigF_i sigF_j S on the lower triangle (rhs row: -sigF_j S).
__global__ void __launch_bounds__(256) dense_scale_kernel(double* __restrict__ S, long long ld, int n /* = rhs_row */,
                                                          const double* __restrict__ sigF) {
  const int j0 = blockIdx.x * kDenseTileCols, i0 = blockIdx.y * kDenseTileRows;
  if (j0 > i0 + kDenseTileRows - 1) return;
  const int j = j0 + (threadIdx.x & 31) * 2;
  const int ib = i0 + (threadIdx.x >> 5) * 4;
  if (j >= n) return;
  const double sj0 = sigF[j], sj1 = j + 1 < n ? sigF[j + 1] : 0.0;
  double2 v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ib + q;
    v[q] = (i <= n && j <= i) ? *reinterpret_cast<const double2*>(S + (size_t)i * ld + j) : make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ib + q;
    if (i > n || j > i) continue;
    const double si = (i == n) ? 1.0 : sigF[i];
    double* dst = S + (size_t)i * ld + j;
    dst[0] = -si * sj0 * v[q].x;
    if (j + 1 <= i && j + 1 < n) dst[1] = -si * sj1 * v[q].y;
  }
}

// Adds the F-pose diagonal blocks sig H_ff sig + D^2, the camera border and
// the gradient to the dense system.  One thread per F pose.
__global__ void dense_add_pose_kernel(int n_f, const double* __restrict__ HF, const double* __restrict__ HFx,
                                      const double* __restrict__ sigF, double radius, double min_diag, double max_diag,
                                      double* __restrict__ S, long long ld, int cam_row, int rhs_row) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_f) return;
  const double* rec = HF + (size_t)f * NV;
  const double sc_cam = sigF[cam_row];
  double s[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) s[i] = sigF[6 * f + i];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double h = rec[tri6(j, i)] * s[i] * s[j];
      if (i == j) h += fmin(fmax(h, min_diag), max_diag) / radius;
      S[(size_t)(6 * f + i) * ld + 6 * f + j] += h;
    }
    S[(size_t)cam_row * ld + 6 * f + i] += rec[27 + i] * s[i] * sc_cam;
    S[(size_t)rhs_row * ld + 6 * f + i] += rec[21 + i] * s[i];
    if (HFx) {  // radial model: rows of l1, l2
      S[(size_t)(cam_row + 1) * ld + 6 * f + i] += HFx[(size_t)f * NVX + i] * s[i] * sigF[cam_row + 1];
      S[(size_t)(cam_row + 2) * ld + 6 * f + i] += HFx[(size_t)f * NVX + 6 + i] * s[i] * sigF[cam_row + 2];
    }
  }
}

// Intrinsics block / rhs tail / identity padding.  cam_minus = column sums of seg_cam (12 values).
__global__ void dense_add_camera_kernel(LmScalars* __restrict__ sc, const double* __restrict__ cam_minus, int nk,
                                        double radius, double min_diag, double max_diag, double* __restrict__ S,
                                        long long ld, int cam_row, int rhs_row, int n_pad) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double H[6] = {sc->cam_H, sc->H_f_l1, sc->H_f_l2, sc->H_l1_l1, sc->H_l1_l2, sc->H_l2_l2};  // 00 01 02 11 12 22
    const double g[3] = {sc->cam_g, sc->g_l1, sc->g_l2};
    const double sg[3] = {sc->sigma_f, sc->sigma_l1, sc->sigma_l2};
    const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int q = 0; q < nk; ++q) {
      for (int p = 0; p <= q; ++p) {
        double v = sg[q] * sg[p] * (H[idx[q][p]] - cam_minus[idx[q][p]]);
        if (p == q) v += fmin(fmax(sg[q] * sg[q] * H[idx[q][q]], min_diag), max_diag) / radius;
        S[(size_t)(cam_row + q) * ld + cam_row + p] = v;
      }
      S[(size_t)rhs_row * ld + cam_row + q] = sg[q] * (g[q] - cam_minus[6 + q]);
    }
    S[(size_t)rhs_row * ld + rhs_row] = 1e300;
    if (cam_minus[9] != 0.0) sc->chol_fail = 1.0;  // some E pose's damped 6x6 block was not positive definite
  }
  const int i = rhs_row + 1 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) S[(size_t)i * ld + i] = 1.0;
}

// uF = sigF * yF (the unscaled, un-negated F step); yF is row rhs_row of the
// factored array after back substitution.
__global__ void scale_uF_kernel(int n, const double* __restrict__ y, const double* __restrict__ sigF,
                                double* __restrict__ uF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) uF[i] = y[i] * sigF[i];
}

// Back substitution of the E poses and the cross term of the model cost change.
//   y_e = z - Ht_ee^-1 (sum_j sig_e W_j uF[f_j]) - yb uF[cam],   delta_e = -sig_e y_e
//   cross_e = sum_j delta_e^T W_j delta_f,   delta_f = -uF[f_j]
// One warp per E pose, lanes over its blocks; the 6x6 factor is rebuilt in
// registers from the pose record (cheaper than keeping Ht_ee^-1 W in HBM).
// Bytes per corner: 72 (W) + 1 (index).
struct BacksubArgs {
  SchurArgs sa;       // HE, W, sig_e, radius, e_off, f_idx, Z, YB
  const double* uF;   // [6 n_f + nk]
  const int32_t* e_order;  // [n_e] E pose handled by each thread group (storage order of the segments), or null: index order
  int cam_row, nk;
  double* d_e;        // [6 n_e] step of the E poses
  double* seg_part;   // [grid][5] CTA partials of (cross term -sum_j d_e^T W_j u_j, 0, |x - x_cand|^2, |x|^2, model terms)
  unsigned* ticket;
  double* out5;       // -> LmScalars::cross (5 consecutive scalars: cross, cand_r2 (rewritten later), step2_e, xnorm2_e, mq_e)
  // the E side of the step (what apply_step_kernel does for the F side), folded in: candidate point,
  // norms, model-cost terms and the candidate's prep record
  const double* x_e;      // [6 n_e] current E poses
  double* x_cand;         // [6 n_e] out: x + d
  int e_is_capture;       // which prep record the E poses get
  double tag_size;
  double* cap_pre_c;      // candidate prep records (E = captures): [n_e][kCapPre], cap_rt_c [n_e][12]
  double* cap_rt_c;
  double* tag_pre_c;      // (E = tags): [n_e][kTagPre], tag_cor_c [n_e][12]
  double* tag_cor_c;
};
constexpr int kBsGroup = 8;  // lanes per E pose: four poses per warp (8 blocks per capture is the common case)
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = kBsGroup / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void store_cap_prep(const double pose[6], double* __restrict__ cap_pre, double* __restrict__ cap_rt, size_t i) {
  double rec[kCapPre];
  prep_capture(pose, rec);
  double2* o = reinterpret_cast<double2*>(cap_pre + (size_t)kCapPre * i);
#pragma unroll
  for (int k = 0; k < kCapPre / 2; ++k) o[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
  if (cap_rt) {
    double2* c = reinterpret_cast<double2*>(cap_rt + (size_t)12 * i);
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
    c[4] = make_double2(rec[8], rec[18]);
    c[5] = make_double2(rec[19], rec[20]);
  }
}
__device__ __forceinline__ void store_tag_prep(const double pose[6], double tag_size, double* __restrict__ tag_pre,
                                               double* __restrict__ tag_cor, size_t i) {
  double rec[kTagPre];
  prep_tag(pose, tag_size, rec);
  double2* o = reinterpret_cast<double2*>(tag_pre + (size_t)kTagPre * i);
#pragma unroll
  for (int k = 0; k < kTagPre / 2; ++k) o[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
  if (tag_cor) {
    double2* c = reinterpret_cast<double2*>(tag_cor + (size_t)12 * i);
    const double w[12] = {rec[0], rec[1], rec[2], rec[12], rec[13], rec[14], rec[24], rec[25], rec[26], rec[36], rec[37], rec[38]};
#pragma unroll
    for (int k = 0; k < 6; ++k) c[k] = make_double2(w[2 * k], w[2 * k + 1]);
  }
}
// Six CTAs per SM (80 registers, a few spilled doubles): the kernel is a chain of dependent loads per pose, so
// resident warps count for more than registers (4 CTAs at 128 registers: 125 us, 6 at 80: 105 us, 8 at 64: 125 us).
__global__ void __launch_bounds__(128, 6) backsub_kernel(const BacksubArgs a) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int slot = gid / kBsGroup, gl = gid % kBsGroup;
  const bool valid = slot < a.sa.n_e;
  // groups walk the E poses in the order their segments are stored (e_order: locality rank -> pose), so that W streams
  const int e = valid ? (a.e_order ? a.e_order[slot] : slot) : 0;
  const int beg = valid ? a.sa.e_off[e] : 0, k = valid ? a.sa.e_end[e] - beg : 0;
  double L[36], zt[6], hk[6], s[6];
  const size_t ps = a.sa.plane;
  double t[6] = {0, 0, 0, 0, 0, 0};
  // the 36 W loads of a block are in flight before the pose record is fetched and factored
  for (int j = gl; j < k; j += kBsGroup) {
    const int blk = beg + j;
    const int f = a.sa.f_idx[blk];
    double u[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) u[c] = a.uF[6 * (size_t)f + c];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int c = 0; c < 6; ++c) t[i] += a.sa.W[(size_t)(i * 6 + c) * ps + blk] * u[c];
  }
  if (k > 0) {
    load_scaled_E(a.sa, e, L, zt, hk, s);
    chol6(L);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) t[i] = group_sum(t[i]);
  double de[6] = {0, 0, 0, 0, 0, 0};
  double cross = 0.0;
  // a constant E pose has sig_e = 0: everything below then gives d = 0 exactly
  const bool active = k > 0 && !(a.sa.e_const && a.sa.e_const[e]);
  if (k > 0) {
    double tr[6];  // sum_j W_j u_j, unscaled
#pragma unroll
    for (int i = 0; i < 6; ++i) { tr[i] = t[i]; t[i] *= s[i]; }
    chol6_solve(L, t);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double ybu = a.sa.YB[6 * a.nk * (size_t)e + i] * a.uF[a.cam_row];
      if (a.nk == 3)
        ybu += a.sa.YB[18 * (size_t)e + 6 + i] * a.uF[a.cam_row + 1] + a.sa.YB[18 * (size_t)e + 12 + i] * a.uF[a.cam_row + 2];
      de[i] = active ? -s[i] * (a.sa.Z[8 * (size_t)e + i] - t[i] - ybu) : 0.0;
      // cross term of the model cost change, -sum_j d_e^T W_j u_j = -d_e . (sum_j W_j u_j)
      cross -= de[i] * tr[i];
    }
  }
  // ---- the step of this E pose: candidate point, norms, model-cost terms (one row of H per lane)
  double d2 = 0.0, x2 = 0.0, mq = 0.0;
  if (valid) {
    const double* rec = a.sa.HE + (size_t)e * NV;
    if (active && gl < 6) {
      double hd = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) hd += rec[gl <= c ? tri6(gl, c) : tri6(c, gl)] * de[c];
      double border = -a.uF[a.cam_row] * rec[27 + gl];
      if (a.sa.HEx) border += -a.uF[a.cam_row + 1] * a.sa.HEx[(size_t)e * NVX + gl] - a.uF[a.cam_row + 2] * a.sa.HEx[(size_t)e * NVX + 6 + gl];
      double dg = 0.0;  // de[gl] without dynamic register indexing
#pragma unroll
      for (int c = 0; c < 6; ++c) dg = c == gl ? de[c] : dg;
      mq = dg * (rec[21 + gl] + 0.5 * hd + border);
    }
    if (gl == 0) {
      double xc[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double x = a.x_e[6 * (size_t)e + i];
        xc[i] = x + de[i];
        a.d_e[6 * (size_t)e + i] = de[i];
        a.x_cand[6 * (size_t)e + i] = xc[i];
        if (active) {
          const double dd = x - xc[i];  // Ceres: (x - candidate_x).norm()
          d2 += dd * dd;
          x2 += x * x;
        }
      }
      if (a.e_is_capture) { if (a.cap_pre_c) store_cap_prep(xc, a.cap_pre_c, a.cap_rt_c, (size_t)e); }
      else if (a.tag_pre_c) store_tag_prep(xc, a.tag_size, a.tag_pre_c, a.tag_cor_c, (size_t)e);
    }
  }
  const double v[5] = {warp_sum(valid && gl == 0 ? cross : 0.0), 0.0, warp_sum(d2), warp_sum(x2), warp_sum(mq)};
  grid_reduce_last_cta<5, false, 4>(v, a.seg_part, a.ticket, a.out5);
}

// delta_F = -uF ; candidate = x + delta on both pose sides and the camera;
// per-warp partials of ||delta||^2 and ||x||^2 over poses that own blocks.
struct ApplyArgs {
  int n_pose;
  const int32_t* seg_off;   // [n_pose] segment start / seg_end [n_pose] segment end (sorted by this side)
  const int32_t* seg_end;
  const double* blocks_all; // multi-GPU, F side: blocks of the pose summed over all ranks (a rank may hold none), else null
  const unsigned char* constant; // [n_pose] != 0: the pose is held constant (arslam_set_constant), or null
  const double* x;          // [6 n_pose]
  const double* step;       // [6 n_pose]: d_e (already the step) or uF (to be negated)
  int negate;
  const double* rec;        // [n_pose][NV] normal-equation records of this side (unscaled)
  const double* recx;       // [n_pose][NVX] l1, l2 borders (radial model) or null
  const double* uF_cam;     // -> uF[cam_row .. cam_row + nk); the intrinsics step is its negative
  double* delta;            // [6 n_pose] out (unscaled step)
  double* x_cand;           // [6 n_pose] out
  double* warp_out;         // [grid][3] CTA partials of sum delta^2, sum x^2, sum (g.d + d^T H d / 2 + d_cam H_pose,f.d)
  unsigned* ticket;
  double* out;              // [3]
  int count_norms;          // 0: this rank does not own the sums of this side (multi-GPU)
  // camera step (done by one thread of the launch that has cam != null)
  const double* cam;        // [3] current intrinsics, or null
  double* cam_c;            // [3] candidate intrinsics
  double* d_cam;            // [3] step
  double* sc;               // LmScalars
  int nk;
  // prep record of the candidate point (saves the separate prep launch): pose_is_capture selects which
  int pose_is_capture;
  double tag_size;
  double* cap_pre_c;        // [n_pose][kCapPre] / cap_rt_c [n_pose][12], or
  double* cap_rt_c;
  double* tag_pre_c;        // [n_pose][kTagPre] / tag_cor_c [n_pose][12]; all null: no prep
  double* tag_cor_c;
};
__global__ void apply_step_kernel(const ApplyArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (a.cam && i == 0) {
    const int dslot[3] = {13, 32, 33}, xslot[3] = {15, 36, 37};
    for (int q = 0; q < 3; ++q) {
      const double d = q < a.nk ? -a.uF_cam[q] : 0.0;
      a.d_cam[q] = d;
      a.cam_c[q] = a.cam[q] + d;
      a.sc[dslot[q]] = a.cam[q] - (a.cam[q] + d);
      a.sc[xslot[q]] = a.cam[q];
    }
  }
  double d2 = 0.0, x2 = 0.0, mq = 0.0;
  if (i < a.n_pose) {
    const bool active = (a.blocks_all ? a.blocks_all[i] > 0.0 : a.seg_end[i] > a.seg_off[i]) && !(a.constant && a.constant[i]);
    double d[6], xcand[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double x = a.x[6 * (size_t)i + k];
      double dk = a.step[6 * (size_t)i + k];
      dk = active ? (a.negate ? -dk : dk) : 0.0;
      d[k] = dk;
      const double xc = x + dk;
      xcand[k] = xc;
      a.delta[6 * (size_t)i + k] = dk;
      a.x_cand[6 * (size_t)i + k] = xc;
      if (active && a.count_norms) {
        const double dd = x - xc;  // Ceres: (x - candidate_x).norm()
        d2 += dd * dd;
        x2 += x * x;
      }
    }
    if (active && a.count_norms) {
      const double* rec = a.rec + (size_t)i * NV;
      const double dcam = -a.uF_cam[0];
      const double dl1 = a.recx ? -a.uF_cam[1] : 0.0, dl2 = a.recx ? -a.uF_cam[2] : 0.0;
      double q = 0.0;
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        double hd = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) hd += rec[p <= c ? tri6(p, c) : tri6(c, p)] * d[c];
        double border = dcam * rec[27 + p];
        if (a.recx) border += dl1 * a.recx[(size_t)i * NVX + p] + dl2 * a.recx[(size_t)i * NVX + 6 + p];
        q += d[p] * (rec[21 + p] + 0.5 * hd + border);
      }
      mq = q;
    }
    if (a.pose_is_capture) { if (a.cap_pre_c) store_cap_prep(xcand, a.cap_pre_c, a.cap_rt_c, (size_t)i); }
    else if (a.tag_pre_c) store_tag_prep(xcand, a.tag_size, a.tag_pre_c, a.tag_cor_c, (size_t)i);
  }
  const double v[3] = {warp_sum(d2), warp_sum(x2), warp_sum(mq)};
  grid_reduce_last_cta<3, false, 4>(v, a.warp_out, a.ticket, a.out);
}

// max |g| over poses that own blocks
__global__ void __launch_bounds__(128) gradmax_kernel(int n_pose, const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_end,
                                                      const double* __restrict__ rec,
                                                      double* __restrict__ part, unsigned* __restrict__ ticket, double* __restrict__ out,
                                                      const double* __restrict__ head, double* __restrict__ sc, int nk,
                                                      const double* __restrict__ blocks_all, const unsigned char* __restrict__ constant) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // the freshly summed camera scalars move next to the other LM scalars (head may be null)
  if (head && i < 3) sc[i] = head[i];
  if (head && nk == 3 && i >= 4 && i < 12) sc[24 + (i - 4)] = head[i];
  double m = 0.0;
  if (i < n_pose && (blocks_all ? blocks_all[i] > 0.0 : seg_end[i] > seg_off[i]) && !(constant && constant[i])) {
#pragma unroll
    for (int k = 0; k < 6; ++k) m = fmax(m, fabs(rec[(size_t)i * NV + 21 + k]));
  }
  const double v[1] = {warp_max(m)};
  grid_reduce_last_cta<1, true, 4>(v, part, ticket, out);
}

}  // namespace ars
