// pcg.cuh -- kernel (4), sparse branch: the reduced system kept as a 6x6
// block-sparse (BSR) matrix over the F poses plus a dense focal-length border,
// solved by block-Jacobi preconditioned conjugate gradients in ONE persistent
// cooperative kernel (two grid syncs per iteration, deterministic reductions).
//
// North-star item (4): "a PCG solve otherwise" -- used when the reduced matrix
// is large and sparse (BASELINE config 3: 30 001 unknowns, ~21 blocks per
// block row, ~30 MB, L2 resident), where the reference's DENSE_SCHUR
// (ar_slam/src/ar_slam_util.cpp:1011) would need a 7.2 GB dense factorisation.
#pragma once
#include <algorithm>
#include <cooperative_groups.h>
#include <cstdint>
#include <string>
#include <vector>

#include "schur.cuh"

namespace ars {
namespace cg = cooperative_groups;

constexpr int kPcgThreads = 1024;  // one CTA per SM, ~one block row per warp
constexpr int kPcgWarps = kPcgThreads / 32;

struct PcgWorkspace {
  int n_f = 0, nnzb = 0, grid = 0;
  bool valid = false;
  long long total_iterations = 0;
  int32_t* row_ptr = nullptr;   // [n_f + 1]
  int32_t* col_idx = nullptr;   // [nnzb]
  int32_t* src_slot = nullptr;  // [nnzb]: slot holding the (lower) source block, transposed if col > row
  double* Sfin = nullptr;       // [nnzb][36] final scaled + damped blocks
  double* Minv = nullptr;       // [n_f][36]  inverse diagonal blocks (block-Jacobi)
  double* vec = nullptr;        // x | r | z | p0 | p1 | q | border | rhs : 8 x (6 n_f + 2)
  double* partial = nullptr;    // [grid][8]
  double* scal = nullptr;       // [16]: S_kk, 1/S_kk, iterations, fail, ...
  size_t value_count() const { return (size_t)36 * nnzb + (size_t)12 * n_f; }
  void release() {
    cudaFree(row_ptr); cudaFree(col_idx); cudaFree(src_slot); cudaFree(Sfin); cudaFree(Minv);
    cudaFree(vec); cudaFree(partial); cudaFree(scal);
    row_ptr = col_idx = src_slot = nullptr; Sfin = Minv = vec = partial = scal = nullptr;
    valid = false;
  }
  ~PcgWorkspace() { release(); }
};

// ---- symbolic phase (host): block pattern of sum_e W_e^T W_e over the E segments
inline int pcg_symbolic(PcgWorkspace& ws, int n_e, int n_f, const int32_t* e_off, const int32_t* f_of_blk,
                        cudaStream_t st, std::string& err) {
  ws.release();
  std::vector<uint64_t> keys;
  keys.reserve((size_t)e_off[n_e] * 8);
  for (int e = 0; e < n_e; ++e)
    for (int i = e_off[e]; i < e_off[e + 1]; ++i)
      for (int j = e_off[e]; j < e_off[e + 1]; ++j)
        keys.push_back((uint64_t)(uint32_t)f_of_blk[i] << 32 | (uint32_t)f_of_blk[j]);
  for (int f = 0; f < n_f; ++f) keys.push_back((uint64_t)(uint32_t)f << 32 | (uint32_t)f);  // every diagonal block exists
  std::sort(keys.begin(), keys.end());
  keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
  const int nnzb = (int)keys.size();
  std::vector<int32_t> row_ptr(n_f + 1, 0), col(nnzb), src(nnzb);
  for (int s = 0; s < nnzb; ++s) {
    row_ptr[(keys[s] >> 32) + 1]++;
    col[s] = (int32_t)(keys[s] & 0xffffffffu);
  }
  for (int f = 0; f < n_f; ++f) row_ptr[f + 1] += row_ptr[f];
  for (int r = 0; r < n_f; ++r)
    for (int s = row_ptr[r]; s < row_ptr[r + 1]; ++s) {
      const int c = col[s];
      if (c <= r) { src[s] = s; continue; }
      // transpose partner: slot of (c, r)
      const int32_t* b = col.data() + row_ptr[c];
      const int32_t* e = col.data() + row_ptr[c + 1];
      const int32_t* it = std::lower_bound(b, e, r);
      src[s] = (int32_t)(it - col.data());
    }
  ws.n_f = n_f;
  ws.nnzb = nnzb;
  const size_t nvec = (size_t)6 * n_f + 2;
  cudaError_t ce = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (ce == cudaSuccess) ce = cudaMalloc(p, bytes); };
  A((void**)&ws.row_ptr, sizeof(int32_t) * (n_f + 1));
  A((void**)&ws.col_idx, sizeof(int32_t) * nnzb);
  A((void**)&ws.src_slot, sizeof(int32_t) * nnzb);
  A((void**)&ws.Sfin, sizeof(double) * 36 * nnzb);
  A((void**)&ws.Minv, sizeof(double) * 36 * n_f);
  A((void**)&ws.vec, sizeof(double) * 8 * nvec);
  A((void**)&ws.partial, sizeof(double) * 8 * 4096);
  A((void**)&ws.scal, sizeof(double) * 16);
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  cudaMemcpyAsync(ws.row_ptr, row_ptr.data(), sizeof(int32_t) * (n_f + 1), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(ws.col_idx, col.data(), sizeof(int32_t) * nnzb, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(ws.src_slot, src.data(), sizeof(int32_t) * nnzb, cudaMemcpyHostToDevice, st);
  ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { err = std::string("pcg symbolic upload: ") + cudaGetErrorString(ce); return -2; }
  ws.valid = true;
  return 0;
}

// ---- numeric formation: same elimination as schur_eliminate_kernel, but the
// pair products land in the BSR value array (lower blocks only; the upper ones
// are mirrored by pcg_finalize_kernel).
struct SparseTarget {
  const int32_t* row_ptr;
  const int32_t* col_idx;
  double* Sraw;     // [nnzb][36]
  double* borderm;  // [6 n_f]  sum W~^T yb
  double* rhsm;     // [6 n_f]  sum W~^T z
  __device__ __forceinline__ void add_border(int f, int c, double b0, double b1) const {
    atomicAdd(borderm + 6 * (size_t)f + c, b0);
    atomicAdd(rhsm + 6 * (size_t)f + c, b1);
  }
  // lower block (row fj, col fi), found by bisection in the row's sorted column list
  __device__ __forceinline__ double* block(int fi, int fj) const {
    int lo = row_ptr[fj], hi = row_ptr[fj + 1] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col_idx[mid] < fi) lo = mid + 1; else hi = mid;
    }
    return Sraw + 36 * (size_t)lo;
  }
  __device__ __forceinline__ void add(double* blk, int r, int c, double v, bool diag, bool twice) const {
    if (!diag) { atomicAdd(blk + c * 6 + r, v); return; }
    atomicAdd(blk + r * 6 + c, v);
    if (twice) atomicAdd(blk + c * 6 + r, v);
  }
};

// ---- finalize: scale by sigma_F, add the F-pose diagonal blocks and damping,
// mirror the upper blocks, invert the diagonal blocks for the preconditioner.
// One thread per block slot.
struct PcgFinalizeArgs {
  int n_f, nnzb;
  const int32_t* row_ptr;
  const int32_t* col_idx;
  const int32_t* src_slot;
  const double* Sraw;
  const double* borderm;
  const double* rhsm;
  const double* HF;       // [n_f][NV]
  const double* sigF;     // [6 n_f + 1]
  const LmScalars* sc;
  const double* cam_minus;
  double radius, min_diag, max_diag;
  double* Sfin;
  double* Minv;
  double* border;         // [6 n_f] final S_f,cam
  double* rhs;            // [6 n_f + 1] final right-hand side
  double* scal;           // [0] S_kk  [1] 1/S_kk  [3] fail
};

__global__ void pcg_finalize_kernel(const PcgFinalizeArgs a) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per block row
  if (row == 0) {
    const double sf = a.sc->sigma_f;
    const double h = a.sc->cam_H * sf * sf;
    const double d = fmin(fmax(h, a.min_diag), a.max_diag) / a.radius;
    const double skk = sf * sf * (a.sc->cam_H - a.cam_minus[0]) + d;
    a.scal[0] = skk;
    a.scal[1] = 1.0 / skk;
    a.rhs[6 * (size_t)a.n_f] = sf * (a.sc->cam_g - a.cam_minus[1]);
    if (!(skk > 0.0)) a.scal[3] = 1.0;
  }
  if (row >= a.n_f) return;
  double sr[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) sr[i] = a.sigF[6 * (size_t)row + i];
  const double scam = a.sigF[6 * (size_t)a.n_f];
  const double* rec = a.HF + (size_t)row * NV;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    a.border[6 * (size_t)row + i] = sr[i] * scam * (rec[27 + i] - a.borderm[6 * (size_t)row + i]);
    a.rhs[6 * (size_t)row + i] = sr[i] * (rec[21 + i] - a.rhsm[6 * (size_t)row + i]);
  }
  for (int s = a.row_ptr[row]; s < a.row_ptr[row + 1]; ++s) {
    const int col = a.col_idx[s];
    const double* src = a.Sraw + 36 * (size_t)a.src_slot[s];
    double* dst = a.Sfin + 36 * (size_t)s;
    if (col != row) {
      double scl[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) scl[i] = a.sigF[6 * (size_t)col + i];
      const bool tr = col > row;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) dst[i * 6 + j] = -sr[i] * scl[j] * (tr ? src[j * 6 + i] : src[i * 6 + j]);
    } else {
      double D[36];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const double h = rec[i <= j ? tri6(i, j) : tri6(j, i)];
          // the raw diagonal block holds its lower triangle (+ upper from the i != j duplicates): symmetrise from the lower part
          const double m = i >= j ? src[i * 6 + j] : src[j * 6 + i];
          double v = sr[i] * sr[j] * (h - m);
          if (i == j) v += fmin(fmax(sr[i] * sr[i] * h, a.min_diag), a.max_diag) / a.radius;
          D[i * 6 + j] = v;
          dst[i * 6 + j] = v;
        }
      // inverse through the Cholesky factor (columns of the identity)
      const bool ok = chol6(D);
      if (!ok) a.scal[3] = 1.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double e[6] = {0, 0, 0, 0, 0, 0};
        e[c] = 1.0;
        chol6_solve(D, e);
#pragma unroll
        for (int i = 0; i < 6; ++i) a.Minv[36 * (size_t)row + i * 6 + c] = e[i];
      }
    }
  }
}

// ---- the solver: persistent cooperative kernel ------------------------------
struct PcgArgs {
  int n_f, max_iter;
  double tol;
  const int32_t* row_ptr;
  const int32_t* col_idx;
  const double* S;       // [nnzb][36]
  const double* Minv;    // [n_f][36]
  const double* border;  // [6 n_f]
  const double* rhs;     // [6 n_f + 1]
  double* x;             // [6 n_f + 1] out
  double* r;
  double* z;
  double* p0;
  double* p1;
  double* q;
  double* partial;       // [grid][8]
  double* scal;          // [0] S_kk [1] 1/S_kk [2] iterations (out) [3] fail (in/out)
};

// deterministic grid-wide sums: every CTA adds the same numbers in the same order
template <int NS>
__device__ __forceinline__ void grid_sums(cg::grid_group& grid, double (&v)[NS], double* partial_base, double* sm, int& parity) {
  // two alternating partial buffers: a CTA may start the next reduction while a slower one still reads this one
  double* partial = partial_base + (parity ? 8 * 2048 : 0);
  parity ^= 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NS; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NS; ++i) sm[wid * NS + i] = v[i];
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      double acc = lane < kPcgWarps ? sm[lane * NS + i] : 0.0;
      acc = warp_sum(acc);
      if (lane == 0) partial[(size_t)blockIdx.x * 8 + i] = acc;
    }
  }
  grid.sync();
  // every CTA reads all CTA partials (one load per thread) and reduces them in the same fixed tree
  double acc[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) acc[i] = threadIdx.x < gridDim.x ? partial[(size_t)threadIdx.x * 8 + i] : 0.0;
  const int nwarp_used = ((int)gridDim.x + 31) >> 5;
  if (wid < nwarp_used) {
#pragma unroll
    for (int i = 0; i < NS; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < NS; ++i) sm[64 + wid * NS + i] = acc[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double t = 0.0;
    for (int w = 0; w < nwarp_used; ++w) t += sm[64 + w * NS + i];
    v[i] = t;
  }
  __syncthreads();
}

__global__ void pcg_publish_kernel(const double* scal, double* sc) {
  sc[3] = scal[2];                       // PCG iterations of this solve
  if (scal[3] != 0.0) sc[12] = 1.0;      // failure -> invalid LM step
}

__global__ void __launch_bounds__(kPcgThreads, 1) pcg_kernel(const PcgArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[64 + 32 * 2 + 8];
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * kPcgThreads + threadIdx.x) >> 5, nw = (gridDim.x * kPcgThreads) >> 5;
  const int g = lane >> 3, rr_ = lane & 7;  // 4 block groups x 8 lanes (6 active rows)
  const bool act = rr_ < 6;
  const int n_f = a.n_f, camrow = 6 * n_f;
  const double skk = a.scal[0], iskk = a.scal[1];
  const bool is_cam_owner = (blockIdx.x == 0 && threadIdx.x == 0);
  int parity = 0;

  // x = 0, r = b, z = M^-1 r, p = z ; sums: rz, bb
  double s2[2] = {0.0, 0.0};
  for (int f = gw; f < n_f; f += nw) {
    double rv = 0.0;
    if (lane < 6) rv = a.rhs[6 * (size_t)f + lane];
    double zv = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double rc = __shfl_sync(0xffffffffu, rv, c);
      if (lane < 6) zv += a.Minv[36 * (size_t)f + lane * 6 + c] * rc;
    }
    if (lane < 6) {
      a.x[6 * (size_t)f + lane] = 0.0;
      a.r[6 * (size_t)f + lane] = rv;
      a.z[6 * (size_t)f + lane] = zv;
      a.p0[6 * (size_t)f + lane] = 0.0;
      s2[0] += rv * zv;
      s2[1] += rv * rv;
    }
  }
  if (is_cam_owner) {
    const double rv = a.rhs[camrow], zv = rv * iskk;
    a.x[camrow] = 0.0; a.r[camrow] = rv; a.z[camrow] = zv; a.p0[camrow] = 0.0;
    s2[0] += rv * zv;
    s2[1] += rv * rv;
  }
  grid_sums<2>(grid, s2, a.partial, sm, parity);
  double rz = s2[0];
  const double bb = s2[1];
  const double thresh = a.tol * a.tol * bb;
  double beta = 0.0;
  double* p_old = a.p0;
  double* p_new = a.p1;
  int it = 0;
  bool fail = !(bb >= 0.0) || !isfinite(bb);
  if (bb == 0.0 || fail) {
    if (is_cam_owner) { a.scal[2] = 0.0; if (fail) a.scal[3] = 1.0; }
    return;
  }
  while (it < a.max_iter) {
    ++it;
    // ---- phase A: p_new = z + beta p_old ; q = S p_new ; sums: p.q and the camera row of q
    double sa[2] = {0.0, 0.0};
    const double pk = a.z[camrow] + beta * p_old[camrow];
    for (int f = gw; f < n_f; f += nw) {
      double acc = 0.0, acc2 = 0.0;
      const int s0 = a.row_ptr[f], s1 = a.row_ptr[f + 1];
      for (int s = s0 + g; s < s1; s += 8) {   // two independent blocks in flight per lane group
        const int sB = s + 4;
        const bool hasB = sB < s1;
        const int c = a.col_idx[s];
        const int cB = hasB ? a.col_idx[sB] : c;
        if (act) {
          const double* B = a.S + 36 * (size_t)s + rr_ * 6;
          const double* zc = a.z + 6 * (size_t)c;
          const double* pc = p_old + 6 * (size_t)c;
          const double* B2 = a.S + 36 * (size_t)(hasB ? sB : s) + rr_ * 6;
          const double* zc2 = a.z + 6 * (size_t)cB;
          const double* pc2 = p_old + 6 * (size_t)cB;
          double t1 = 0.0, t2 = 0.0;
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            t1 += B[j] * (zc[j] + beta * pc[j]);
            t2 += B2[j] * (zc2[j] + beta * pc2[j]);
          }
          acc += t1;
          if (hasB) acc2 += t2;
        }
      }
      acc += acc2;
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      if (lane < 6) {
        const double pf = a.z[6 * (size_t)f + lane] + beta * p_old[6 * (size_t)f + lane];
        const double bd = a.border[6 * (size_t)f + lane];
        const double qf = acc + bd * pk;
        p_new[6 * (size_t)f + lane] = pf;
        a.q[6 * (size_t)f + lane] = qf;
        sa[0] += pf * qf;
        sa[1] += bd * pf;
      }
    }
    grid_sums<2>(grid, sa, a.partial, sm, parity);
    const double qk = sa[1] + skk * pk;
    const double pq = sa[0] + pk * qk;
    if (!(pq > 0.0) || !isfinite(pq)) { fail = true; break; }
    const double alpha = rz / pq;
    // ---- phase B: x += alpha p ; r -= alpha q ; z = M^-1 r ; sums: r.z, r.r
    double sb[2] = {0.0, 0.0};
    for (int f = gw; f < n_f; f += nw) {
      double rv = 0.0;
      if (lane < 6) {
        const size_t i = 6 * (size_t)f + lane;
        a.x[i] += alpha * p_new[i];
        rv = a.r[i] - alpha * a.q[i];
        a.r[i] = rv;
      }
      double zv = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double rc = __shfl_sync(0xffffffffu, rv, c);
        if (lane < 6) zv += a.Minv[36 * (size_t)f + lane * 6 + c] * rc;
      }
      if (lane < 6) {
        a.z[6 * (size_t)f + lane] = zv;
        sb[0] += rv * zv;
        sb[1] += rv * rv;
      }
    }
    if (is_cam_owner) {
      p_new[camrow] = pk;
      a.x[camrow] += alpha * pk;
      const double rv = a.r[camrow] - alpha * qk;
      a.r[camrow] = rv;
      const double zv = rv * iskk;
      a.z[camrow] = zv;
      sb[0] += rv * zv;
      sb[1] += rv * rv;
    }
    grid_sums<2>(grid, sb, a.partial, sm, parity);
    beta = sb[0] / rz;
    rz = sb[0];
    double* t = p_old; p_old = p_new; p_new = t;
    if (sb[1] <= thresh) break;
    if (!isfinite(sb[1])) { fail = true; break; }
  }
  if (is_cam_owner) {
    a.scal[2] = (double)it;
    if (fail) a.scal[3] = 1.0;
  }
}

inline cudaError_t pcg_init() { return cudaSuccess; }

}  // namespace ars
