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
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cstdint>
#include <string>
#include <vector>

#include "schur.cuh"

namespace ars {
namespace cg = cooperative_groups;

constexpr int kPcgThreads = 1024;  // one CTA per SM, ~one block row per warp
constexpr int kPcgWarps = kPcgThreads / 32;

template <class T>
struct PcgBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    // a buffer that grows AGAIN grows by half as much again (the incremental schedules append one capture per solve)
    const size_t want = p ? n + n / 2 + 64 : n + n / 8 + 64;
    cudaFree(p);
    p = nullptr; cap = 0;
    const cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  PcgBuf() = default;
  PcgBuf(const PcgBuf&) = delete;
  PcgBuf& operator=(const PcgBuf&) = delete;
  ~PcgBuf() { cudaFree(p); }
};

// length of one PCG vector: the F poses' components + up to three intrinsics (even, so every vector is 16-byte aligned)
inline size_t pcg_nvec(int n_f) { return (size_t)6 * n_f + 4; }

struct PcgWorkspace {
  int n_f = 0, nnzb = 0, nnz_lower = 0, grid = 0;  // nnz_lower: blocks with col <= row, the ones that are formed
  bool valid = false;
  long long total_iterations = 0;
  int32_t* row_ptr = nullptr;   // [n_f + 1]
  int32_t* col_idx = nullptr;   // [nnzb]
  int32_t* src_slot = nullptr;  // [nnzb]: index, in the compact array of lower blocks, of the source block (transposed if col > row)
  double* Sfin = nullptr;       // [nnzb][36] final scaled + damped blocks
  double* Minv = nullptr;       // [n_f][36]  inverse diagonal blocks (block-Jacobi)
  double* vec = nullptr;        // x | r | z | p0 | p1 | q | border | rhs | border_l1 | border_l2 : 10 x pcg_nvec(n_f)
  double* partial = nullptr;    // [grid][8]
  double* scal = nullptr;       // [16]: S_kk, 1/S_kk, iterations, fail, ...
  // shared-memory resident variant: block rows are split into one contiguous range per CTA
  bool smem_ok = false;
  int smem_grid = 0, cap_slots = 0, max_halo = 0, max_slots = 0, max_rows = 0;
  size_t smem_bytes = 0;
  int32_t* cta_row = nullptr;   // [smem_grid + 1]
  int32_t* halo_ptr = nullptr;  // [smem_grid + 1]
  int32_t* halo_col = nullptr;  // distinct block columns each CTA reads
  uint16_t* lcol = nullptr;     // [nnzb] column of a slot as an index into its CTA's halo list
  int32_t* diag_slot = nullptr; // [n_f]
  int32_t* slot_row = nullptr;  // [nnzb]
  int32_t* pair_off = nullptr;  // [n_blk] (E-sorted) pairs before each block
  int32_t* pair_slot = nullptr; // [n_pairs]
  long long n_pairs = 0;
  // lower blocks | border (focal) | rhs | borders of l1, l2 (radial model)
  size_t value_count(int nk = 1) const { return (size_t)36 * nnz_lower + (size_t)12 * n_f + (nk == 3 ? (size_t)12 * n_f : 0); }
  // storage (grow-only, so a new problem of similar size allocates nothing); the pointers above alias it
  PcgBuf<int32_t> b_lowflag, b_lrank, b_row_ptr, b_col_idx, b_src_slot, b_cta_row, b_halo_ptr, b_halo_col, b_diag_slot, b_slot_row, b_pair_off, b_pair_slot;
  PcgBuf<uint16_t> b_lcol;
  PcgBuf<double> b_Sfin, b_Minv, b_vec, b_partial, b_scal;
  // symbolic phase scratch
  PcgBuf<unsigned long long> keys[2], hkeys[2];
  PcgBuf<long long> cnt_keys, cnt_pairs, off_keys, off_pairs;
  PcgBuf<int> counts;
  PcgBuf<unsigned char> tmp;
  int* h_counts = nullptr;  // pinned
  ~PcgWorkspace() { if (h_counts) cudaFreeHost(h_counts); }
};

// ---- symbolic phase (device): block pattern of sum_e W_e^T W_e over the E segments.
// Every E-sorted block (e, f) emits the keys (f, f') of all blocks (e, f') of its segment; the
// keys row * n_f + col are radix-sorted and de-duplicated, which is the BSR pattern in row-major
// order.  Everything below (row pointers, transpose partners, the per-SM row ranges of the
// shared-memory solver and their halo lists) is derived from the sorted keys by bisection, so
// the host only ever sees a handful of counts.
__global__ void sym_counts_kernel(int n_blk, const int32_t* __restrict__ e_idx, const int32_t* __restrict__ e_off,
                                  const int32_t* __restrict__ e_end, long long* __restrict__ n_keys, long long* __restrict__ n_pairs) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos > n_blk) return;
  if (pos == n_blk) { n_keys[pos] = 0; n_pairs[pos] = 0; return; }  // the scans' totals land here
  const int e = e_idx[pos];
  const int beg = e_off[e], k = e_end[e] - beg;
  n_keys[pos] = k;
  n_pairs[pos] = schur_pairs_of(pos - beg, k);
}
__global__ void sym_emit_keys_kernel(int n_blk, int n_f, const int32_t* __restrict__ e_idx, const int32_t* __restrict__ e_off,
                                     const int32_t* __restrict__ e_end, const int32_t* __restrict__ f_idx, const long long* __restrict__ key_off,
                                     unsigned long long* __restrict__ keys) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos < n_f) keys[key_off[n_blk] + pos] = (unsigned long long)pos * n_f + pos;  // every diagonal block exists
  if (pos >= n_blk) return;
  const int e = e_idx[pos];
  const int beg = e_off[e], k = e_end[e] - beg;
  const unsigned long long row = (unsigned long long)f_idx[pos] * n_f;
  unsigned long long* out = keys + key_off[pos];
  for (int q = 0; q < k; ++q) out[q] = row + f_idx[beg + q];
}
__global__ void sym_narrow_kernel(int n, const long long* __restrict__ in, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}
__device__ __forceinline__ int sym_lower_bound(const unsigned long long* __restrict__ a, int n, unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__global__ void sym_rows_kernel(int n_f, int nnzb, const unsigned long long* __restrict__ keys, int32_t* __restrict__ row_ptr) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n_f) row_ptr[r] = sym_lower_bound(keys, nnzb, (unsigned long long)r * n_f);
}
__global__ void sym_slots_kernel(int n_f, int nnzb, const unsigned long long* __restrict__ keys, int32_t* __restrict__ col,
                                 int32_t* __restrict__ slot_row, int32_t* __restrict__ src, int32_t* __restrict__ diag_slot) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl >= nnzb) return;
  const unsigned long long key = keys[sl];
  const int r = (int)(key / n_f), c = (int)(key % n_f);
  col[sl] = c;
  slot_row[sl] = r;
  if (c == r) diag_slot[r] = sl;
  // an upper block takes its values from the transpose partner, the slot of (c, r)
  src[sl] = c <= r ? sl : sym_lower_bound(keys, nnzb, (unsigned long long)c * n_f + r);
}
// The elimination only forms the lower blocks (col <= row); they are stored compactly, so the
// array that is zeroed, reduced across ranks and read by the finalisation is half the pattern.
__global__ void sym_lower_flag_kernel(int nnzb, const int32_t* __restrict__ col, const int32_t* __restrict__ slot_row, int32_t* __restrict__ flag) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl <= nnzb) flag[sl] = (sl < nnzb && col[sl] <= slot_row[sl]) ? 1 : 0;
}
__global__ void sym_compact_src_kernel(int nnzb, const int32_t* __restrict__ lrank, int32_t* __restrict__ src) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl < nnzb) src[sl] = lrank[src[sl]];
}
// one contiguous block-row range per CTA, balanced by block count
__global__ void sym_cta_rows_kernel(int G, int n_f, int nnzb, const int32_t* __restrict__ row_ptr, int32_t* __restrict__ cta_row) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > G) return;
  if (c == 0) { cta_row[0] = 0; return; }
  if (c == G) { cta_row[G] = n_f; return; }
  const long long target = (long long)nnzb * c / G;
  int lo = 0, hi = n_f;  // number of rows r with row_ptr[r + 1] <= target
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (row_ptr[mid + 1] <= target) lo = mid + 1; else hi = mid;
  }
  cta_row[c] = lo;
}
__device__ __forceinline__ int sym_cta_of_row(const int32_t* __restrict__ cta_row, int G, int r) {
  int lo = 0, hi = G;  // last c with cta_row[c] <= r
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (cta_row[mid] <= r) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__global__ void sym_halo_keys_kernel(int nnzb, int n_f, int G, const int32_t* __restrict__ cta_row, const int32_t* __restrict__ slot_row,
                                     const int32_t* __restrict__ col, unsigned long long* __restrict__ hkeys) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl >= nnzb) return;
  hkeys[sl] = (unsigned long long)sym_cta_of_row(cta_row, G, slot_row[sl]) * n_f + col[sl];
}
__global__ void sym_halo_kernel(int n_halo, int n_f, int G, const unsigned long long* __restrict__ hkeys, int32_t* __restrict__ halo_col,
                                int32_t* __restrict__ halo_ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_halo) halo_col[i] = (int32_t)(hkeys[i] % n_f);
  if (i <= G) halo_ptr[i] = sym_lower_bound(hkeys, n_halo, (unsigned long long)i * n_f);
}
__global__ void sym_lcol_kernel(int nnzb, int n_halo, int n_f, int G, const int32_t* __restrict__ cta_row, const int32_t* __restrict__ slot_row,
                                const int32_t* __restrict__ col, const unsigned long long* __restrict__ hkeys,
                                const int32_t* __restrict__ halo_ptr, uint16_t* __restrict__ lcol) {
  const int sl = blockIdx.x * blockDim.x + threadIdx.x;
  if (sl >= nnzb) return;
  const int c = sym_cta_of_row(cta_row, G, slot_row[sl]);
  lcol[sl] = (uint16_t)(sym_lower_bound(hkeys, n_halo, (unsigned long long)c * n_f + col[sl]) - halo_ptr[c]);
}
// out: max halo size, max slots, max rows over the CTAs (single CTA)
__global__ void sym_maxima_kernel(int G, const int32_t* __restrict__ cta_row, const int32_t* __restrict__ halo_ptr,
                                  const int32_t* __restrict__ row_ptr, int32_t* __restrict__ out) {
  __shared__ int m[3];
  if (threadIdx.x < 3) m[threadIdx.x] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < G; c += blockDim.x) {
    atomicMax(&m[0], halo_ptr[c + 1] - halo_ptr[c]);
    atomicMax(&m[1], row_ptr[cta_row[c + 1]] - row_ptr[cta_row[c]]);
    atomicMax(&m[2], cta_row[c + 1] - cta_row[c]);
  }
  __syncthreads();
  if (threadIdx.x < 3) out[threadIdx.x] = m[threadIdx.x];
}

inline int sym_key_bits(int n_f) {
  unsigned long long top = (unsigned long long)n_f * (unsigned long long)n_f;  // also the padding sentinel
  int bits = 1;
  while (bits < 64 && (top >> bits)) ++bits;
  return bits;
}
// sorts keys[0..n) (ws.keys[0] -> de-duplicated into ws.keys[0] again); returns the count in *n_out
inline cudaError_t pcg_sort_unique(PcgWorkspace& ws, int n, int bits, cudaStream_t st, int* n_out) {
  cudaError_t ce;
  size_t t1 = 0, t2 = 0;
  if ((ce = cub::DeviceRadixSort::SortKeys(nullptr, t1, ws.keys[0].p, ws.keys[1].p, n, 0, bits, st)) != cudaSuccess) return ce;
  if ((ce = cub::DeviceSelect::Unique(nullptr, t2, ws.keys[1].p, ws.keys[0].p, ws.counts.p, n, st)) != cudaSuccess) return ce;
  if ((ce = ws.tmp.ensure(std::max(t1, t2))) != cudaSuccess) return ce;
  t1 = t2 = ws.tmp.cap;
  if ((ce = cub::DeviceRadixSort::SortKeys(ws.tmp.p, t1, ws.keys[0].p, ws.keys[1].p, n, 0, bits, st)) != cudaSuccess) return ce;
  if ((ce = cub::DeviceSelect::Unique(ws.tmp.p, t2, ws.keys[1].p, ws.keys[0].p, ws.counts.p, n, st)) != cudaSuccess) return ce;
  if ((ce = cudaMemcpyAsync(ws.h_counts, ws.counts.p, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return ce;
  if ((ce = cudaStreamSynchronize(st)) != cudaSuccess) return ce;
  *n_out = ws.h_counts[0];
  return cudaSuccess;
}

// phase 1: the sorted unique keys of this rank's blocks -> ws.keys[0][0 .. ws.nnzb), and pair_off
inline int pcg_symbolic_keys(PcgWorkspace& ws, int n_e, int n_f, int n_blk, const int32_t* e_idx, const int32_t* e_off, const int32_t* e_end,
                             const int32_t* f_idx, cudaStream_t st, std::string& err) {
  (void)n_e;
  ws.valid = false;
  cudaError_t ce = cudaSuccess;
#define PCG_TRY(x) do { if (ce == cudaSuccess) ce = (x); } while (0)
  if (!ws.h_counts) PCG_TRY(cudaMallocHost((void**)&ws.h_counts, 8 * sizeof(long long)));
  PCG_TRY(ws.cnt_keys.ensure((size_t)n_blk + 1)); PCG_TRY(ws.cnt_pairs.ensure((size_t)n_blk + 1));
  PCG_TRY(ws.off_keys.ensure((size_t)n_blk + 1)); PCG_TRY(ws.off_pairs.ensure((size_t)n_blk + 1));
  PCG_TRY(ws.counts.ensure(8));
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  sym_counts_kernel<<<(n_blk + 256) / 256, 256, 0, st>>>(n_blk, e_idx, e_off, e_end, ws.cnt_keys.p, ws.cnt_pairs.p);
  size_t t1 = 0;
  PCG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t1, ws.cnt_keys.p, ws.off_keys.p, n_blk + 1, st));
  PCG_TRY(ws.tmp.ensure(t1));
  t1 = ws.tmp.cap;
  PCG_TRY(cub::DeviceScan::ExclusiveSum(ws.tmp.p, t1, ws.cnt_keys.p, ws.off_keys.p, n_blk + 1, st));
  t1 = ws.tmp.cap;
  PCG_TRY(cub::DeviceScan::ExclusiveSum(ws.tmp.p, t1, ws.cnt_pairs.p, ws.off_pairs.p, n_blk + 1, st));
  long long* hc = reinterpret_cast<long long*>(ws.h_counts);
  PCG_TRY(cudaMemcpyAsync(hc, ws.off_keys.p + n_blk, sizeof(long long), cudaMemcpyDeviceToHost, st));
  PCG_TRY(cudaMemcpyAsync(hc + 1, ws.off_pairs.p + n_blk, sizeof(long long), cudaMemcpyDeviceToHost, st));
  PCG_TRY(cudaStreamSynchronize(st));
  if (ce != cudaSuccess) { err = std::string("pcg symbolic (counts): ") + cudaGetErrorString(ce); return -2; }
  const long long n_keys = hc[0] + n_f;
  ws.n_pairs = hc[1];
  if (n_keys >= (1LL << 31)) { err = "pcg symbolic: more than 2^31 block pairs"; return -2; }
  PCG_TRY(ws.keys[0].ensure((size_t)n_keys)); PCG_TRY(ws.keys[1].ensure((size_t)n_keys));
  if (ws.n_pairs < (1LL << 31)) {
    PCG_TRY(ws.b_pair_off.ensure((size_t)n_blk + 1)); PCG_TRY(ws.b_pair_slot.ensure((size_t)std::max<long long>(ws.n_pairs, 1)));
    if (ce == cudaSuccess) sym_narrow_kernel<<<(n_blk + 255) / 256, 256, 0, st>>>(n_blk, ws.off_pairs.p, ws.b_pair_off.p);
    ws.pair_off = ws.b_pair_off.p; ws.pair_slot = ws.b_pair_slot.p;
  } else {
    ws.pair_off = nullptr; ws.pair_slot = nullptr;
  }
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  sym_emit_keys_kernel<<<(std::max(n_blk, n_f) + 255) / 256, 256, 0, st>>>(n_blk, n_f, e_idx, e_off, e_end, f_idx, ws.off_keys.p, ws.keys[0].p);
  int nn = 0;
  PCG_TRY(pcg_sort_unique(ws, (int)n_keys, sym_key_bits(n_f), st, &nn));
  if (ce != cudaSuccess) { err = std::string("pcg symbolic (sort): ") + cudaGetErrorString(ce); return -2; }
  ws.n_f = n_f;
  ws.nnzb = nn;
  return 0;
}

// phase 2: everything derived from the sorted keys ws.keys[0][0 .. ws.nnzb)
inline int pcg_symbolic_build(PcgWorkspace& ws, cudaStream_t st, std::string& err, int n_sm, size_t smem_limit) {
  const int n_f = ws.n_f, nnzb = ws.nnzb, G = n_sm;
  const size_t nvec = pcg_nvec(n_f);
  cudaError_t ce = cudaSuccess;
  PCG_TRY(ws.b_row_ptr.ensure((size_t)n_f + 1)); PCG_TRY(ws.b_col_idx.ensure(nnzb)); PCG_TRY(ws.b_src_slot.ensure(nnzb));
  PCG_TRY(ws.b_Sfin.ensure((size_t)36 * nnzb)); PCG_TRY(ws.b_Minv.ensure((size_t)36 * n_f)); PCG_TRY(ws.b_vec.ensure(10 * nvec));
  PCG_TRY(ws.b_partial.ensure(8 * 4096)); PCG_TRY(ws.b_scal.ensure(16));
  PCG_TRY(ws.b_diag_slot.ensure(std::max(n_f, 1))); PCG_TRY(ws.b_slot_row.ensure(std::max(nnzb, 1)));
  PCG_TRY(ws.b_cta_row.ensure((size_t)G + 1)); PCG_TRY(ws.b_halo_ptr.ensure((size_t)G + 1)); PCG_TRY(ws.b_lcol.ensure(std::max(nnzb, 1)));
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  ws.row_ptr = ws.b_row_ptr.p; ws.col_idx = ws.b_col_idx.p; ws.src_slot = ws.b_src_slot.p; ws.Sfin = ws.b_Sfin.p;
  ws.Minv = ws.b_Minv.p; ws.vec = ws.b_vec.p; ws.partial = ws.b_partial.p; ws.scal = ws.b_scal.p;
  ws.diag_slot = ws.b_diag_slot.p; ws.slot_row = ws.b_slot_row.p; ws.cta_row = ws.b_cta_row.p; ws.halo_ptr = ws.b_halo_ptr.p;
  ws.lcol = ws.b_lcol.p;
  const unsigned long long* keys = ws.keys[0].p;
  sym_rows_kernel<<<(n_f + 256) / 256, 256, 0, st>>>(n_f, nnzb, keys, ws.row_ptr);
  sym_slots_kernel<<<(nnzb + 255) / 256, 256, 0, st>>>(n_f, nnzb, keys, ws.col_idx, ws.slot_row, ws.src_slot, ws.diag_slot);
  {
    PCG_TRY(ws.b_lowflag.ensure((size_t)nnzb + 1)); PCG_TRY(ws.b_lrank.ensure((size_t)nnzb + 1));
    if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
    sym_lower_flag_kernel<<<(nnzb + 256) / 256, 256, 0, st>>>(nnzb, ws.col_idx, ws.slot_row, ws.b_lowflag.p);
    size_t t1 = 0;
    PCG_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t1, ws.b_lowflag.p, ws.b_lrank.p, nnzb + 1, st));
    PCG_TRY(ws.tmp.ensure(t1));
    t1 = ws.tmp.cap;
    PCG_TRY(cub::DeviceScan::ExclusiveSum(ws.tmp.p, t1, ws.b_lowflag.p, ws.b_lrank.p, nnzb + 1, st));
    sym_compact_src_kernel<<<(nnzb + 255) / 256, 256, 0, st>>>(nnzb, ws.b_lrank.p, ws.src_slot);
    PCG_TRY(cudaMemcpyAsync(ws.h_counts + 4, ws.b_lrank.p + nnzb, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  sym_cta_rows_kernel<<<(G + 256) / 256, 256, 0, st>>>(G, n_f, nnzb, ws.row_ptr, ws.cta_row);
  // halo lists: distinct (CTA, column) pairs.  keys[0] still holds the pattern, so sort from a copy in keys[1]
  // through the spare halves: hkeys live in ws.hkeys[0/1]
  PCG_TRY(ws.hkeys[0].ensure(std::max(nnzb, 1))); PCG_TRY(ws.hkeys[1].ensure(std::max(nnzb, 1)));
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  sym_halo_keys_kernel<<<(nnzb + 255) / 256, 256, 0, st>>>(nnzb, n_f, G, ws.cta_row, ws.slot_row, ws.col_idx, ws.hkeys[0].p);
  int n_halo = 0;
  {
    unsigned long long top = (unsigned long long)(G + 1) * (unsigned long long)n_f;
    int bits = 1;
    while (bits < 64 && (top >> bits)) ++bits;
    size_t t1 = 0, t2 = 0;
    PCG_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t1, ws.hkeys[0].p, ws.hkeys[1].p, nnzb, 0, bits, st));
    PCG_TRY(cub::DeviceSelect::Unique(nullptr, t2, ws.hkeys[1].p, ws.hkeys[0].p, ws.counts.p, nnzb, st));
    PCG_TRY(ws.tmp.ensure(std::max(t1, t2)));
    t1 = t2 = ws.tmp.cap;
    PCG_TRY(cub::DeviceRadixSort::SortKeys(ws.tmp.p, t1, ws.hkeys[0].p, ws.hkeys[1].p, nnzb, 0, bits, st));
    PCG_TRY(cub::DeviceSelect::Unique(ws.tmp.p, t2, ws.hkeys[1].p, ws.hkeys[0].p, ws.counts.p, nnzb, st));
    PCG_TRY(cudaMemcpyAsync(ws.h_counts, ws.counts.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PCG_TRY(cudaStreamSynchronize(st));
    if (ce != cudaSuccess) { err = std::string("pcg symbolic (halo): ") + cudaGetErrorString(ce); return -2; }
    n_halo = ws.h_counts[0];
  }
  PCG_TRY(ws.b_halo_col.ensure(std::max(n_halo, 1)));
  if (ce != cudaSuccess) { err = std::string("pcg workspace: ") + cudaGetErrorString(ce); return -2; }
  ws.halo_col = ws.b_halo_col.p;
  sym_halo_kernel<<<(std::max(n_halo, G + 1) + 255) / 256, 256, 0, st>>>(n_halo, n_f, G, ws.hkeys[0].p, ws.halo_col, ws.halo_ptr);
  sym_lcol_kernel<<<(nnzb + 255) / 256, 256, 0, st>>>(nnzb, n_halo, n_f, G, ws.cta_row, ws.slot_row, ws.col_idx, ws.hkeys[0].p,
                                                      ws.halo_ptr, ws.lcol);
  sym_maxima_kernel<<<1, 256, 0, st>>>(G, ws.cta_row, ws.halo_ptr, ws.row_ptr, ws.counts.p + 1);
  PCG_TRY(cudaMemcpyAsync(ws.h_counts + 1, ws.counts.p + 1, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
  PCG_TRY(cudaStreamSynchronize(st));
  if (ce != cudaSuccess) { err = std::string("pcg symbolic (layout): ") + cudaGetErrorString(ce); return -2; }
#undef PCG_TRY
  const int max_halo = ws.h_counts[1], max_slots = ws.h_counts[2], max_rows = ws.h_counts[3];
  ws.nnz_lower = ws.h_counts[4];
  ws.smem_grid = G;
  ws.max_halo = max_halo;
  ws.max_slots = max_slots;
  ws.max_rows = max_rows;
  // dynamic shared memory: matrix slice | gathered halo vector | local column ids
  const size_t fixed = (size_t)max_halo * 48 + (size_t)max_rows * (36 + 6) * 8 + ((size_t)max_halo + max_rows + 4) * 4 +
                       (((size_t)max_slots * 2 + 15) / 16) * 16 + 2048;
  ws.smem_ok = max_halo < 65535 && max_rows <= 64 && smem_limit > fixed + 288 * 16;
  if (ws.smem_ok) {
    ws.cap_slots = (int)std::min<size_t>((smem_limit - fixed) / 288, (size_t)max_slots);
    ws.smem_bytes = (size_t)ws.cap_slots * 288 + fixed - 2048;
  }
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
  double* borderx;  // [2][6 n_f] the same for the columns of l1, l2 (radial model), else null
  int n_f;
  __device__ __forceinline__ void add_border(int f, int c, double b0, double b1) const {
    red_add_f64(borderm + 6 * (size_t)f + c, b0);
    red_add_f64(rhsm + 6 * (size_t)f + c, b1);
  }
  __device__ __forceinline__ void add_border_x(int f, int c, double bl1, double bl2) const {
    red_add_f64(borderx + 6 * (size_t)f + c, bl1);
    red_add_f64(borderx + 6 * (size_t)(n_f + f) + c, bl2);
  }
  const int32_t* pair_slot;  // [n_pairs] precomputed slot of every (partner, block) pair, or null
  // lower block (row fj, col fi): precomputed slot, else bisection in the row's sorted column list
  __device__ __forceinline__ int find(int fi, int fj) const {
    int lo = row_ptr[fj], hi = row_ptr[fj + 1] - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col_idx[mid] < fi) lo = mid + 1; else hi = mid;
    }
    return lo;
  }
  const int32_t* lower_of;  // [nnzb] compact index of a slot's lower source block (= PcgWorkspace::src_slot)
  __device__ __forceinline__ double* block(int fi, int fj, long long pair) const {
    return Sraw + 36 * (size_t)(pair_slot ? pair_slot[pair] : lower_of[find(fi, fj)]);
  }
  __device__ __forceinline__ double* elem(double* blk, int e) const { return blk + e; }
};

// fills pair_slot once per problem: thread per E-sorted block, same partner enumeration as
// schur_eliminate_kernel
__global__ void pair_slot_kernel(int n_blk, const int32_t* __restrict__ e_idx, const int32_t* __restrict__ e_off,
                                 const int32_t* __restrict__ e_end, const int32_t* __restrict__ f_idx, const int32_t* __restrict__ pair_off,
                                 SparseTarget t, int32_t* __restrict__ out) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_blk) return;
  const int e = e_idx[pos];
  const int beg = e_off[e], k = e_end[e] - beg, j = pos - beg;
  const int fj = f_idx[pos];
  const int np = schur_pairs_of(j, k);
  for (int d = 0; d < np; ++d) {
    int i2 = j + d;
    if (i2 >= k) i2 -= k;
    const int fp = f_idx[beg + i2];
    out[(size_t)pair_off[pos] + d] = t.lower_of[t.find(min(fj, fp), max(fj, fp))];
  }
}

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
  const double* HFx;      // [n_f][NVX] pose x (l1, l2) terms of the radial model, or null
  const double* borderx;  // [2][6 n_f] (SparseTarget::borderx), or null
  int nk;                 // live intrinsics: 1 (focal) or 3 (radial model)
  double* border1;        // [6 n_f] final S_f,l1
  double* border2;        // [6 n_f] final S_f,l2
  const double* sigF;     // [6 n_f + nk]
  const LmScalars* sc;
  const double* cam_minus;
  double radius, min_diag, max_diag;
  double* Sfin;
  double* Minv;
  double* border;         // [6 n_f] final S_f,cam
  double* rhs;            // [6 n_f + 1] final right-hand side
  double* scal;           // [0] S_kk  [1] 1/S_kk  [3] fail; radial model: [4..9] the 3 x 3 intrinsics block (00 01 02 11 12 22), [10..15] its inverse
  double* partial;        // PcgWorkspace::partial (cleared here)
};

// off-diagonal blocks: one thread per (slot, block row i) -- 6 consecutive doubles each
__global__ void pcg_finalize_offdiag_kernel(const PcgFinalizeArgs a, const int32_t* __restrict__ slot_row) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 16) a.scal[t] = 0.0;          // (pcg_finalize_kernel, the next launch, fills [0], [1], [3])
  if (t < 24) a.partial[t] = 0.0;       // the three rotating accumulators of grid_sums_atomic
  if (t >= a.nnzb * 6) return;
  const int s = t / 6, i = t - s * 6;
  const int row = slot_row[s], col = a.col_idx[s];
  if (col == row) return;
  const double* src = a.Sraw + 36 * (size_t)a.src_slot[s];
  double* dst = a.Sfin + 36 * (size_t)s + i * 6;
  const double sri = a.sigF[6 * (size_t)row + i];
  const bool tr = col > row;
#pragma unroll
  for (int j = 0; j < 6; ++j)
    dst[j] = -sri * a.sigF[6 * (size_t)col + j] * (tr ? src[j * 6 + i] : src[i * 6 + j]);
}

// diagonal blocks, border, right-hand side and the block-Jacobi inverses: one thread per block row
// Eight lanes per block row (six at work): lane c owns component c of the border / right-hand side, row c of the final
// diagonal block and column c of its inverse; the 6 x 6 factorisation is repeated by the six lanes, which is cheaper than
// the six dependent solves one thread per row used to walk.
constexpr int kFinLanes = 8;
__global__ void pcg_finalize_kernel(const PcgFinalizeArgs a, const int32_t* __restrict__ diag_slot) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gid / kFinLanes, c = gid % kFinLanes;
  if (gid == 0) {
    const double sf = a.sc->sigma_f;
    const double h = a.sc->cam_H * sf * sf;
    const double d = fmin(fmax(h, a.min_diag), a.max_diag) / a.radius;
    const double skk = sf * sf * (a.sc->cam_H - a.cam_minus[0]) + d;
    a.scal[0] = skk;
    a.scal[1] = 1.0 / skk;
    a.rhs[6 * (size_t)a.n_f] = sf * (a.sc->cam_g - a.cam_minus[6]);
    if (!(skk > 0.0) || a.cam_minus[9] != 0.0) a.scal[3] = 1.0;
    if (a.nk == 3) {
      // the damped 3 x 3 block of (f, l1, l2) as in dense_add_camera_kernel, and its inverse (adjugate) for the preconditioner
      const double H[6] = {a.sc->cam_H, a.sc->H_f_l1, a.sc->H_f_l2, a.sc->H_l1_l1, a.sc->H_l1_l2, a.sc->H_l2_l2};
      const double g[3] = {a.sc->cam_g, a.sc->g_l1, a.sc->g_l2};
      const double sg[3] = {a.sc->sigma_f, a.sc->sigma_l1, a.sc->sigma_l2};
      const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
      double K[6];
      for (int q = 0; q < 3; ++q) {
        for (int p2 = q; p2 < 3; ++p2) {
          double v = sg[q] * sg[p2] * (H[idx[q][p2]] - a.cam_minus[idx[q][p2]]);
          if (p2 == q) v += fmin(fmax(sg[q] * sg[q] * H[idx[q][q]], a.min_diag), a.max_diag) / a.radius;
          K[idx[q][p2]] = v;
        }
        a.rhs[6 * (size_t)a.n_f + q] = sg[q] * (g[q] - a.cam_minus[6 + q]);
      }
      const double c00 = K[3] * K[5] - K[4] * K[4], c01 = K[2] * K[4] - K[1] * K[5], c02 = K[1] * K[4] - K[2] * K[3];
      const double det = K[0] * c00 + K[1] * c01 + K[2] * c02;
      const double m2 = K[0] * K[3] - K[1] * K[1];
      if (!(K[0] > 0.0) || !(m2 > 0.0) || !(det > 0.0)) a.scal[3] = 1.0;
      const double id = 1.0 / det;
      for (int q = 0; q < 6; ++q) a.scal[4 + q] = K[q];
      a.scal[10] = c00 * id; a.scal[11] = c01 * id; a.scal[12] = c02 * id;
      a.scal[13] = (K[0] * K[5] - K[2] * K[2]) * id; a.scal[14] = (K[1] * K[2] - K[0] * K[4]) * id; a.scal[15] = m2 * id;
    }
  }
  if (row >= a.n_f || c >= 6) return;
  double sr[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) sr[i] = a.sigF[6 * (size_t)row + i];
  const double scam = a.sigF[6 * (size_t)a.n_f];
  const double* rec = a.HF + (size_t)row * NV;
  {
    const double sc_ = a.sigF[6 * (size_t)row + c];
    a.border[6 * (size_t)row + c] = sc_ * scam * (rec[27 + c] - a.borderm[6 * (size_t)row + c]);
    a.rhs[6 * (size_t)row + c] = sc_ * (rec[21 + c] - a.rhsm[6 * (size_t)row + c]);
    if (a.nk == 3) {
      const double* rx = a.HFx + (size_t)row * NVX;
      a.border1[6 * (size_t)row + c] = sc_ * a.sigF[6 * (size_t)a.n_f + 1] * (rx[c] - a.borderx[6 * (size_t)row + c]);
      a.border2[6 * (size_t)row + c] = sc_ * a.sigF[6 * (size_t)a.n_f + 2] * (rx[6 + c] - a.borderx[6 * (size_t)(a.n_f + row) + c]);
    }
  }
  const int s = diag_slot[row];
  const double* src = a.Sraw + 36 * (size_t)a.src_slot[s];
  double* dst = a.Sfin + 36 * (size_t)s;
  double D[36];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double h = rec[i <= j ? tri6(i, j) : tri6(j, i)];
      const double m = i >= j ? src[i * 6 + j] : src[j * 6 + i];  // raw diagonal block is symmetric
      double v = sr[i] * sr[j] * (h - m);
      if (i == j) v += fmin(fmax(sr[i] * sr[i] * h, a.min_diag), a.max_diag) / a.radius;
      D[i * 6 + j] = v;
      if (i == c) dst[i * 6 + j] = v;
    }
  const bool ok = chol6(D);
  if (!ok && c == 0) a.scal[3] = 1.0;
  double e[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) e[i] = i == c ? 1.0 : 0.0;
  chol6_solve(D, e);
#pragma unroll
  for (int i = 0; i < 6; ++i) a.Minv[36 * (size_t)row + i * 6 + c] = e[i];
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
  const double* border1; // [6 n_f] l1, l2 columns (radial model: pcg_kernel<3>)
  const double* border2;
  const double* rhs;     // [6 n_f + nk]
  double* x;             // [6 n_f + 1] out
  double* r;
  double* z;
  double* p0;
  double* p1;
  double* q;
  double* partial;       // [grid][8]
  double* scal;          // [0] S_kk [1] 1/S_kk [2] iterations (out) [3] fail (in/out)
  double q_tol;               // quadratic-model termination (pipelined kernel): 0 = off
  unsigned long long* trace;  // debug: [iteration][cta][5] globaltimer stamps, or null
  // results go straight to the LM loop: unscaled step uF = sigF . x, iteration count -> sc[18],
  // failure -> sc[12]
  const double* sigF;
  double* uF;
  double* sc;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define PCG_TRACE(slot) do { if (a.trace && threadIdx.x == 0 && it <= 64) a.trace[((size_t)(it - 1) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime(); } while (0)

// deterministic grid-wide sums: every CTA adds the same numbers in the same order
constexpr int kGsOff = 32 * 4;         // NS <= 4: warp partials in sm[0 .. 128), the second stage behind them
constexpr int kGsSmem = 2 * kGsOff;
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
      for (int i = 0; i < NS; ++i) sm[kGsOff + wid * NS + i] = acc[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    double t = 0.0;
    for (int w = 0; w < nwarp_used; ++w) t += sm[kGsOff + w * NS + i];
    v[i] = t;
  }
  __syncthreads();
}


// NK = live intrinsics: 1 (focal: the scalar S_kk) or 3 (radial model: the 3 x 3 block of f, l1, l2 and three border columns)
template <int NK>
__global__ void __launch_bounds__(kPcgThreads, 1) pcg_kernel(const PcgArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[kGsSmem];
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * kPcgThreads + threadIdx.x) >> 5, nw = (gridDim.x * kPcgThreads) >> 5;
  const int g = lane >> 3, rr_ = lane & 7;  // 4 block groups x 8 lanes (6 active rows)
  const bool act = rr_ < 6;
  const int n_f = a.n_f, camrow = 6 * n_f;
  // intrinsics block K (symmetric) and its inverse Ki
  double K[NK][NK], Ki[NK][NK];
  if (NK == 1) {
    K[0][0] = a.scal[0];
    Ki[0][0] = a.scal[1];
  } else {
    const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
    for (int q = 0; q < NK; ++q)
#pragma unroll
      for (int p2 = 0; p2 < NK; ++p2) { K[q][p2] = a.scal[4 + idx[q][p2]]; Ki[q][p2] = a.scal[10 + idx[q][p2]]; }
  }
  const double* bcol[3] = {a.border, a.border1, a.border2};
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
    double rv[NK];
#pragma unroll
    for (int q = 0; q < NK; ++q) rv[q] = a.rhs[camrow + q];
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      double zv = 0.0;
#pragma unroll
      for (int p2 = 0; p2 < NK; ++p2) zv += Ki[q][p2] * rv[p2];
      a.x[camrow + q] = 0.0; a.r[camrow + q] = rv[q]; a.z[camrow + q] = zv; a.p0[camrow + q] = 0.0;
      s2[0] += rv[q] * zv;
      s2[1] += rv[q] * rv[q];
    }
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
    // ---- phase A: p_new = z + beta p_old ; q = S p_new ; sums: p.q and the intrinsics rows of q
    double sa[1 + NK];
#pragma unroll
    for (int q = 0; q <= NK; ++q) sa[q] = 0.0;
    double pk[NK];
#pragma unroll
    for (int q = 0; q < NK; ++q) pk[q] = a.z[camrow + q] + beta * p_old[camrow + q];
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
        double qf = acc;
#pragma unroll
        for (int q = 0; q < NK; ++q) {
          const double bd = bcol[q][6 * (size_t)f + lane];
          qf += bd * pk[q];
          sa[1 + q] += bd * pf;
        }
        p_new[6 * (size_t)f + lane] = pf;
        a.q[6 * (size_t)f + lane] = qf;
        sa[0] += pf * qf;
      }
    }
    grid_sums<1 + NK>(grid, sa, a.partial, sm, parity);
    double qk[NK];
    double pq = sa[0];
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      qk[q] = sa[1 + q];
#pragma unroll
      for (int p2 = 0; p2 < NK; ++p2) qk[q] += K[q][p2] * pk[p2];
      pq += pk[q] * qk[q];
    }
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
      double rv[NK];
#pragma unroll
      for (int q = 0; q < NK; ++q) {
        p_new[camrow + q] = pk[q];
        a.x[camrow + q] += alpha * pk[q];
        rv[q] = a.r[camrow + q] - alpha * qk[q];
        a.r[camrow + q] = rv[q];
      }
#pragma unroll
      for (int q = 0; q < NK; ++q) {
        double zv = 0.0;
#pragma unroll
        for (int p2 = 0; p2 < NK; ++p2) zv += Ki[q][p2] * rv[p2];
        a.z[camrow + q] = zv;
        sb[0] += rv[q] * zv;
        sb[1] += rv[q] * rv[q];
      }
    }
    grid_sums<2>(grid, sb, a.partial, sm, parity);
    beta = sb[0] / rz;
    rz = sb[0];
    double* t = p_old; p_old = p_new; p_new = t;
    if (sb[1] <= thresh) break;
    if (!isfinite(sb[1])) { fail = true; break; }
  }
  // unscaled step for the LM loop (x is complete: the loop leaves through a grid-wide sum)
  for (int i = blockIdx.x * kPcgThreads + threadIdx.x; i < camrow + NK; i += gridDim.x * kPcgThreads) a.uF[i] = a.sigF[i] * a.x[i];
  if (is_cam_owner) {
    a.scal[2] = (double)it;
    if (fail) a.scal[3] = 1.0;
    a.sc[18] = (double)it;
    if (fail || a.scal[3] != 0.0) a.sc[12] = 1.0;
  }
}


// ---- shared-memory resident variant -------------------------------------------
// Each CTA owns a contiguous range of block rows and keeps that slice of the
// matrix in shared memory across all iterations (config 3: 30.6 MB over 148 SMs
// = 207 KB per SM); per iteration it gathers the ~250 neighbouring 6-vectors it
// needs from L2 once, multiplies out of shared memory, and keeps x, r, q, p of
// its own rows in registers.  HBM / L2 traffic per iteration is the halo only.
struct PcgSmemArgs {
  PcgArgs a;
  const int32_t* cta_row;
  const int32_t* halo_ptr;
  const int32_t* halo_col;
  const uint16_t* lcol;
  int cap_slots, max_halo, max_slots, max_rows;
};

// grid-wide sum of NS doubles: CTA totals are added to a global accumulator with FP64 atomics
// before the barrier and read back after it (three rotating accumulators, the next one is
// cleared by CTA 0).  One L2 round trip after the barrier instead of a 148-slot read + reduce.
template <int NS>
__device__ __forceinline__ void grid_sums_atomic(cg::grid_group& grid, double (&v)[NS], double* accum /* [3][8] */,
                                                 double* sm /* shared, (kPcgWarps + 1) * NS doubles */, int& round) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double* cur = accum + 8 * (round % 3);
  double* nxt = accum + 8 * ((round + 1) % 3);
  ++round;
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
      if (lane == 0) atomicAdd(cur + i, acc);
    }
    if (blockIdx.x == 0 && lane < 8) nxt[lane] = 0.0;
  }
  grid.sync();
  // ONE thread per CTA fetches the totals (every thread of every CTA loading the same line
  // serialises at its L2 slice: ~10 us for 148 x 1024 threads), the rest get them from smem
  if (threadIdx.x < NS) sm[kPcgWarps * NS + threadIdx.x] = __ldcg(cur + threadIdx.x);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NS; ++i) v[i] = sm[kPcgWarps * NS + i];
}

__global__ void __launch_bounds__(kPcgThreads, 1) pcg_smem_kernel(const PcgSmemArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ double sm[96];
  const PcgArgs& a = A.a;
  // dynamic shared memory: matrix slice | halo vector | Minv | border | halo columns | row offsets | local column ids
  double* Ss = reinterpret_cast<double*>(dyn);                       // [cap_slots][36]
  double* xs = Ss + (size_t)A.cap_slots * 36;                        // [max_halo][6]
  double* Ms = xs + (size_t)A.max_halo * 6;                          // [max_rows][36]
  double* bs = Ms + (size_t)A.max_rows * 36;                         // [max_rows][6]
  int32_t* hc = reinterpret_cast<int32_t*>(bs + (size_t)A.max_rows * 6);  // [max_halo]
  int32_t* rp = hc + A.max_halo;                                     // [max_rows + 1] local slot offsets
  uint16_t* lc = reinterpret_cast<uint16_t*>(rp + A.max_rows + 2);   // [max_slots]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_f = a.n_f, camrow = 6 * n_f;
  const int r0 = A.cta_row[blockIdx.x], r1 = A.cta_row[blockIdx.x + 1];
  const int nrow = r1 - r0;
  const int s_beg = a.row_ptr[r0], s_end = a.row_ptr[r1];
  const int nslot = s_end - s_beg, ncache = min(nslot, A.cap_slots);
  const int h0 = A.halo_ptr[blockIdx.x], nhalo = A.halo_ptr[blockIdx.x + 1] - h0;
  if (tid == 0) { sm[81] = a.scal[0]; sm[82] = a.scal[1]; }
  const bool is_cam_owner = (blockIdx.x == 0 && tid == 0);
  int round = 0;
  // everything that is constant over the iterations -> shared memory (once)
  {
    // blocks are stored TRANSPOSED in shared memory ([j][i]): lane (block, j) then reads one
    // 48-byte column with three conflict-free 128-bit loads and feeds all six row sums
    const double* src = a.S + 36 * (size_t)s_beg;
    for (int i = tid; i < ncache * 36; i += kPcgThreads) {
      const int sl = i / 36, k = i - sl * 36, r = k / 6, c = k - r * 6;
      Ss[sl * 36 + c * 6 + r] = src[i];
    }
    for (int i = tid; i < nslot; i += kPcgThreads) lc[i] = A.lcol[s_beg + i];
    for (int i = tid; i < nhalo; i += kPcgThreads) hc[i] = A.halo_col[h0 + i];
    for (int i = tid; i <= nrow; i += kPcgThreads) rp[i] = a.row_ptr[r0 + i] - s_beg;
    for (int i = tid; i < nrow * 36; i += kPcgThreads) Ms[i] = a.Minv[36 * (size_t)r0 + i];
    for (int i = tid; i < nrow * 6; i += kPcgThreads) bs[i] = a.border[6 * (size_t)r0 + i];
  }
  __syncthreads();
  const double skk = sm[81], iskk = sm[82];
  // this warp's rows (at most two); x, r, z, p, q of a row live in registers of lanes 0..5
  const int lrA = wid, lrB = wid + 32;
  const bool hasA = lrA < nrow, hasB = lrB < nrow;
  const int rowA = r0 + lrA, rowB = r0 + lrB;
  double xA = 0.0, xB = 0.0, rA = 0.0, rB = 0.0, qA = 0.0, qB = 0.0, pA = 0.0, pB = 0.0, zA = 0.0, zB = 0.0;
  double s2[2] = {0.0, 0.0};
  auto precond = [&](int lrow, double rv) {
    double zv = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double rc = __shfl_sync(0xffffffffu, rv, c);
      if (lane < 6) zv += Ms[36 * lrow + lane * 6 + c] * rc;
    }
    return zv;
  };
  if (hasA) {
    if (lane < 6) rA = a.rhs[6 * (size_t)rowA + lane];
    zA = precond(lrA, rA);
    if (lane < 6) { a.z[6 * (size_t)rowA + lane] = zA; a.p0[6 * (size_t)rowA + lane] = 0.0; s2[0] += rA * zA; s2[1] += rA * rA; }
  }
  if (hasB) {
    if (lane < 6) rB = a.rhs[6 * (size_t)rowB + lane];
    zB = precond(lrB, rB);
    if (lane < 6) { a.z[6 * (size_t)rowB + lane] = zB; a.p0[6 * (size_t)rowB + lane] = 0.0; s2[0] += rB * zB; s2[1] += rB * rB; }
  }
  double xk = 0.0, rk = 0.0;
  if (is_cam_owner) {
    rk = a.rhs[camrow];
    const double zv = rk * iskk;
    a.z[camrow] = zv; a.p0[camrow] = 0.0;
    s2[0] += rk * zv;
    s2[1] += rk * rk;
  }
  grid_sums_atomic<2>(grid, s2, a.partial, sm, round);
  double rz = s2[0];
  const double bb = s2[1];
  const double thresh = a.tol * a.tol * bb;
  double beta = 0.0;
  double* p_old = a.p0;
  double* p_new = a.p1;
  int it = 0;
  bool fail = !(bb >= 0.0) || !isfinite(bb);
  if (!(bb == 0.0 || fail)) {
    while (it < a.max_iter) {
      ++it;
      PCG_TRACE(0);
      // ---- phase A: gather p = z + beta p_old for the halo columns, q = S p from shared memory
      if (tid == 0) sm[80] = __ldcg(a.z + camrow) + beta * __ldcg(p_old + camrow);
      for (int i = tid; i < nhalo * 6; i += kPcgThreads) {
        const int h = i / 6, k = i - h * 6;
        const size_t gi = 6 * (size_t)hc[h] + k;
        xs[i] = __ldcg(a.z + gi) + beta * __ldcg(p_old + gi);
      }
      __syncthreads();
      PCG_TRACE(5);
      const double pk = sm[80];
      double sa[2] = {0.0, 0.0};
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        if (!(which == 0 ? hasA : hasB)) continue;
        const int lrow = which == 0 ? lrA : lrB;
        const int row = r0 + lrow;
        const int s0 = rp[lrow], s1 = rp[lrow + 1];
        // the block row as a dense 6 x 6(s1 - s0) matrix: lanes over its columns
        double ac[6] = {0, 0, 0, 0, 0, 0};
        for (int c = lane; c < 6 * (s1 - s0); c += 32) {
          const int sl = s0 + c / 6, j = c - (c / 6) * 6;
          const double xv = xs[6 * (int)lc[sl] + j];
          if (sl < ncache) {
            const double2* col = reinterpret_cast<const double2*>(Ss + 36 * (size_t)sl + 6 * j);
            const double2 c0 = col[0], c1 = col[1], c2 = col[2];
            ac[0] += c0.x * xv; ac[1] += c0.y * xv; ac[2] += c1.x * xv;
            ac[3] += c1.y * xv; ac[4] += c2.x * xv; ac[5] += c2.y * xv;
          } else {  // the few slots that did not fit in shared memory: row-major in global memory
            const double* Bg = a.S + 36 * (size_t)(s_beg + sl) + j;
#pragma unroll
            for (int i = 0; i < 6; ++i) ac[i] += __ldg(Bg + 6 * i) * xv;
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) ac[i] = warp_sum(ac[i]);
        const double acc = lane == 0 ? ac[0] : lane == 1 ? ac[1] : lane == 2 ? ac[2] : lane == 3 ? ac[3] : lane == 4 ? ac[4] : ac[5];
        if (lane < 6) {
          const double pf = (which == 0 ? zA : zB) + beta * (which == 0 ? pA : pB);
          const double bd = bs[6 * lrow + lane];
          const double qf = acc + bd * pk;
          p_new[6 * (size_t)row + lane] = pf;   // published for the neighbours' next gather
          sa[0] += pf * qf;
          sa[1] += bd * pf;
          if (which == 0) { pA = pf; qA = qf; } else { pB = pf; qB = qf; }
        }
        if (which == 0) PCG_TRACE(6);
      }
      PCG_TRACE(1);
      grid_sums_atomic<2>(grid, sa, a.partial, sm, round);
      PCG_TRACE(2);
      const double qk = sa[1] + skk * pk;
      const double pq = sa[0] + pk * qk;
      if (!(pq > 0.0) || !isfinite(pq)) { fail = true; break; }
      const double alpha = rz / pq;
      // ---- phase B: own rows only, state in registers
      double sb[2] = {0.0, 0.0};
      if (hasA) {
        xA += alpha * pA;
        rA -= alpha * qA;
        zA = precond(lrA, rA);
        if (lane < 6) { a.z[6 * (size_t)rowA + lane] = zA; sb[0] += rA * zA; sb[1] += rA * rA; }
      }
      if (hasB) {
        xB += alpha * pB;
        rB -= alpha * qB;
        zB = precond(lrB, rB);
        if (lane < 6) { a.z[6 * (size_t)rowB + lane] = zB; sb[0] += rB * zB; sb[1] += rB * rB; }
      }
      if (is_cam_owner) {
        p_new[camrow] = pk;
        xk += alpha * pk;
        rk -= alpha * qk;
        const double zv = rk * iskk;
        a.z[camrow] = zv;
        sb[0] += rk * zv;
        sb[1] += rk * rk;
      }
      PCG_TRACE(3);
      grid_sums_atomic<2>(grid, sb, a.partial, sm, round);
      PCG_TRACE(4);
      beta = sb[0] / rz;
      rz = sb[0];
      double* t = p_old; p_old = p_new; p_new = t;
      if (sb[1] <= thresh) break;
      if (!isfinite(sb[1])) { fail = true; break; }
    }
  }
  if (hasA && lane < 6) { a.x[6 * (size_t)rowA + lane] = xA; a.uF[6 * (size_t)rowA + lane] = a.sigF[6 * (size_t)rowA + lane] * xA; }
  if (hasB && lane < 6) { a.x[6 * (size_t)rowB + lane] = xB; a.uF[6 * (size_t)rowB + lane] = a.sigF[6 * (size_t)rowB + lane] * xB; }
  if (is_cam_owner) {
    a.x[camrow] = xk;
    a.uF[camrow] = a.sigF[camrow] * xk;
    a.scal[2] = (double)it;
    if (fail) a.scal[3] = 1.0;
    a.sc[18] = (double)it;
    if (fail || a.scal[3] != 0.0) a.sc[12] = 1.0;
  }
}

// Six per-lane partial sums -> component c in lanes 4c .. 4c + 3 (recursive halving over the
// lane bits 4, 3, 2, then a plain butterfly over bits 1, 0): 9 double shuffles instead of 30.
__device__ __forceinline__ double reduce6_halving(const double (&ac)[6], int lane) {
  const double a8[8] = {ac[0], ac[1], ac[2], ac[3], ac[4], ac[5], 0.0, 0.0};
  double a4[4], a2[2];
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double recv = __shfl_xor_sync(0xffffffffu, hi ? a8[t] : a8[t + 4], 16);
    a4[t] = (hi ? a8[t + 4] : a8[t]) + recv;
  }
  hi = (lane & 8) != 0;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const double recv = __shfl_xor_sync(0xffffffffu, hi ? a4[t] : a4[t + 2], 8);
    a2[t] = (hi ? a4[t + 2] : a4[t]) + recv;
  }
  hi = (lane & 4) != 0;
  double a1 = (hi ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, hi ? a2[0] : a2[1], 4);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

// ---- pipelined PCG (Ghysels & Vanroose): ONE grid barrier per iteration --------------------
// The classic recurrence needs two grid-wide sums per iteration (p.q, then r.z) with the
// matrix product between them.  The pipelined form carries w = A u, s = A p, z = A q along by
// recurrences, so the sums of iteration i (r.u, w.u, r.r and the border dot b.m) are formed
// BEFORE the product n = A m, m = M^-1 w, and the barrier that publishes m to the neighbours is
// the same barrier that completes the sums.  Same shared-memory resident matrix slice as
// pcg_smem_kernel.  A row's state (x r u w p s q z) lives in one register set per lane: lanes
// 0..5 hold the warp's first row, lanes 8..13 its second, lane 16 of CTA 0 / warp 0 the camera
// scalar.  Rounding differs from the classic recurrence and the attainable accuracy is lower,
// so the host picks this kernel for inexact-Newton tolerances only (pcg_tolerance >= 1e-6).
// NK = 1: the focal length; NK = 3: f, l1, l2 of the radial model in lanes 16..18, their 3 x 3 block and its inverse in
// shared memory, three border columns (one register each per row lane), three border dots among the fused sums.
template <int NK>
__global__ void __launch_bounds__(kPcgThreads, 1) pcg_pipe_kernel(const PcgSmemArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char dyn[];
  constexpr int NSUM = 5 + NK;  // r.u  w.u  r.r  x.b  x.r | b_q.m
  constexpr int kVk = (kPcgWarps + 1) * NSUM + 4;  // intrinsics components of the gathered vector (NK); then K (9) and K^-1 (9)
  constexpr int kKm = kVk + 4, kKi = kKm + 9;
  __shared__ double sm[kKi + 9];
  const PcgArgs& a = A.a;
  double* Ss = reinterpret_cast<double*>(dyn);                       // [cap_slots][36]
  double* xs = Ss + (size_t)A.cap_slots * 36;                        // [max_halo][6]
  double* Ms = xs + (size_t)A.max_halo * 6;                          // [max_rows][36]
  double* bs = Ms + (size_t)A.max_rows * 36;                         // [max_rows][6]
  int32_t* hc = reinterpret_cast<int32_t*>(bs + (size_t)A.max_rows * 6);  // [max_halo]
  int32_t* rp = hc + A.max_halo;                                     // [max_rows + 1] local slot offsets
  uint16_t* lc = reinterpret_cast<uint16_t*>(rp + A.max_rows + 2);   // [max_slots]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_f = a.n_f, camrow = 6 * n_f;
  const int r0 = A.cta_row[blockIdx.x], r1 = A.cta_row[blockIdx.x + 1];
  const int nrow = r1 - r0;
  const int s_beg = a.row_ptr[r0], s_end = a.row_ptr[r1];
  const int nslot = s_end - s_beg, ncache = min(nslot, A.cap_slots);
  const int h0 = A.halo_ptr[blockIdx.x], nhalo = A.halo_ptr[blockIdx.x + 1] - h0;
  if (tid < 9) {
    // the intrinsics block and its inverse, row-major 3 x 3 (NK = 1: element 0 only)
    const int q = tid / 3, p2 = tid - 3 * q;
    const int sym = q <= p2 ? (q == 0 ? p2 : q + p2 + 1) : (p2 == 0 ? q : q + p2 + 1);  // 00 01 02 11 12 22
    sm[kKm + tid] = NK == 1 ? (tid == 0 ? a.scal[0] : 0.0) : a.scal[4 + sym];
    sm[kKi + tid] = NK == 1 ? (tid == 0 ? a.scal[1] : 0.0) : a.scal[10 + sym];
  }
  int round = 0;
  {
    const double* src = a.S + 36 * (size_t)s_beg;
    for (int i = tid; i < ncache * 36; i += kPcgThreads) {
      const int sl = i / 36, k = i - sl * 36, r = k / 6, c = k - r * 6;
      Ss[sl * 36 + c * 6 + r] = src[i];  // transposed blocks, see pcg_smem_kernel
    }
    for (int i = tid; i < nslot; i += kPcgThreads) lc[i] = A.lcol[s_beg + i];
    for (int i = tid; i < nhalo; i += kPcgThreads) hc[i] = A.halo_col[h0 + i];
    for (int i = tid; i <= nrow; i += kPcgThreads) rp[i] = a.row_ptr[r0 + i] - s_beg;
    for (int i = tid; i < nrow * 36; i += kPcgThreads) Ms[i] = a.Minv[36 * (size_t)r0 + i];
    for (int i = tid; i < nrow * 6; i += kPcgThreads) bs[i] = a.border[6 * (size_t)r0 + i];
  }
  __syncthreads();
  const int lrA = wid, lrB = wid + 32;
  const bool hasA = lrA < nrow, hasB = lrB < nrow;
  const bool isA = hasA && lane < 6, isB = hasB && lane >= 8 && lane < 14;
  const bool isK = blockIdx.x == 0 && wid == 0 && lane >= 16 && lane < 16 + NK;
  const int kq = isK ? lane - 16 : 0;
  const bool isRow = isA || isB, owner = isRow || isK;
  const int comp = isB ? lane - 8 : (lane < 6 ? lane : 0);
  const int lrow = isB ? lrB : lrA;
  const size_t gi = isK ? (size_t)(camrow + kq) : 6 * (size_t)(r0 + lrow) + comp;
  double bd[NK];
  bd[0] = isRow ? bs[6 * lrow + comp] : 0.0;
  if (NK == 3) {
    bd[1] = isRow ? a.border1[gi] : 0.0;
    bd[2] = isRow ? a.border2[gi] : 0.0;
  }

  auto precond = [&](double v) {  // M^-1 v: the row's inverse diagonal block, 1 / S_kk for the camera
    double o = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double va = __shfl_sync(0xffffffffu, v, c), vb = __shfl_sync(0xffffffffu, v, 8 + c);
      if (isA) o += Ms[36 * lrA + lane * 6 + c] * va;
      else if (isB) o += Ms[36 * lrB + (lane - 8) * 6 + c] * vb;
    }
    if (NK == 1) {
      if (isK) o = v * sm[kKi];
    } else {
      const double v0 = __shfl_sync(0xffffffffu, v, 16), v1 = __shfl_sync(0xffffffffu, v, 17), v2 = __shfl_sync(0xffffffffu, v, 18);
      if (isK) o = sm[kKi + 3 * kq] * v0 + sm[kKi + 3 * kq + 1] * v1 + sm[kKi + 3 * kq + 2] * v2;
    }
    return o;
  };
  // halo of `buf` -> xs, camera component -> sm[kVk]
  auto gather = [&](const double* buf) {
    if (tid < NK) sm[kVk + tid] = __ldcg(buf + camrow + tid);
    for (int i = tid; i < nhalo * 6; i += kPcgThreads) {
      const int h = i / 6, k = i - h * 6;
      xs[i] = __ldcg(buf + 6 * (size_t)hc[h] + k);
    }
    __syncthreads();
  };
  // this lane's component of A v, v gathered in xs / sm[kVk]; bv = b . v over all rows
  auto product = [&](const double* bv) {  // bv[q] = b_q . v
    double vk[NK];
#pragma unroll
    for (int q = 0; q < NK; ++q) vk[q] = sm[kVk + q];
    double res = 0.0;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      if (!(which == 0 ? hasA : hasB)) continue;
      const int lr = which == 0 ? lrA : lrB;
      const int s0 = rp[lr], s1 = rp[lr + 1];
      double ac[6] = {0, 0, 0, 0, 0, 0};
      for (int c = lane; c < 6 * (s1 - s0); c += 32) {
        const int sl = s0 + c / 6, j = c - (c / 6) * 6;
        const double xv = xs[6 * (int)lc[sl] + j];
        if (sl < ncache) {
          const double2* col = reinterpret_cast<const double2*>(Ss + 36 * (size_t)sl + 6 * j);
          const double2 c0 = col[0], c1 = col[1], c2 = col[2];
          ac[0] += c0.x * xv; ac[1] += c0.y * xv; ac[2] += c1.x * xv;
          ac[3] += c1.y * xv; ac[4] += c2.x * xv; ac[5] += c2.y * xv;
        } else {
          const double* Bg = a.S + 36 * (size_t)(s_beg + sl) + j;
#pragma unroll
          for (int i = 0; i < 6; ++i) ac[i] += __ldg(Bg + 6 * i) * xv;
        }
      }
      const double tot = reduce6_halving(ac, lane);
      const double mine = __shfl_sync(0xffffffffu, tot, (4 * (which == 0 ? lane : lane - 8)) & 31);
      if (which == 0 ? isA : isB) {
        res = mine;
#pragma unroll
        for (int q = 0; q < NK; ++q) res += bd[q] * vk[q];
      }
    }
    if (isK) {
      res = bv[kq];
#pragma unroll
      for (int q = 0; q < NK; ++q) res += sm[kKm + 3 * kq + q] * vk[q];
    }
    return res;
  };

  double x = 0.0, r = 0.0, u = 0.0, w = 0.0, p = 0.0, s_ = 0.0, q = 0.0, z = 0.0;
  double* buf[2] = {a.p0, a.p1};
  // r0 = rhs (x0 = 0), u0 = M^-1 r0, w0 = A u0
  if (owner) r = a.rhs[gi];
  const double b0 = r;      // this lane's component of the right-hand side (quadratic-model termination)
  double Q_prev = 0.0;      // Q(x) = x'Ax - 2 b'x = -x.(b + r); Q(0) = 0
  u = precond(r);
  if (owner) buf[0][gi] = u;
  double s2[1 + NK];
  s2[0] = owner ? r * r : 0.0;
#pragma unroll
  for (int q = 0; q < NK; ++q) s2[1 + q] = isRow ? bd[q] * u : 0.0;
  grid_sums_atomic<1 + NK>(grid, s2, a.partial, sm, round);
  const double bb = s2[0];
  const double thresh = a.tol * a.tol * bb;
  bool fail = !(bb >= 0.0) || !isfinite(bb);
  int it = 0;
  if (!(bb == 0.0 || fail)) {
    gather(buf[0]);
    w = product(s2 + 1);
    double gamma_prev = 1.0, alpha_prev = 1.0;
    while (it < a.max_iter) {
      const double m = precond(w);
      double* mb = buf[(it + 1) & 1];
      if (owner) mb[gi] = m;
      double sv[NSUM];
      sv[0] = owner ? r * u : 0.0; sv[1] = owner ? w * u : 0.0; sv[2] = owner ? r * r : 0.0;
      sv[3] = owner ? x * b0 : 0.0; sv[4] = owner ? x * r : 0.0;
#pragma unroll
      for (int kq2 = 0; kq2 < NK; ++kq2) sv[5 + kq2] = isRow ? bd[kq2] * m : 0.0;
      grid_sums_atomic<NSUM>(grid, sv, a.partial, sm, round);
      const double gamma = sv[0], delta = sv[1];
      if (!isfinite(sv[2])) { fail = true; break; }
      if (sv[2] <= thresh) break;
      if (a.q_tol > 0.0 && it >= 1) {
        // Ceres' ConjugateGradientsSolver rule, the one its trust-region strategies use for inexact steps
        // (q_tolerance = eta, r_tolerance off): stop when zeta = i (Q_i - Q_{i-1}) / Q_i < q_tolerance
        const double Q = -(sv[3] + sv[4]);
        const double zeta = (double)it * (Q - Q_prev) / Q;
        Q_prev = Q;
        if (zeta < a.q_tol) break;
      }
      double beta = 0.0, alpha;
      if (it == 0) {
        if (!(delta > 0.0) || !isfinite(delta)) { fail = true; break; }
        alpha = gamma / delta;
      } else {
        beta = gamma / gamma_prev;
        const double den = delta - beta * gamma / alpha_prev;
        if (!(den > 0.0) || !isfinite(den)) { fail = true; break; }
        alpha = gamma / den;
      }
      gather(mb);
      const double n = product(sv + 5);
      z = n + beta * z;
      q = m + beta * q;
      s_ = w + beta * s_;
      p = u + beta * p;
      x += alpha * p;
      r -= alpha * s_;
      u -= alpha * q;
      w -= alpha * z;
      gamma_prev = gamma;
      alpha_prev = alpha;
      ++it;
    }
  }
  if (owner) { a.x[gi] = x; a.uF[gi] = a.sigF[gi] * x; }
  if (isK) {
    a.scal[2] = (double)it;
    if (fail) a.scal[3] = 1.0;
    a.sc[18] = (double)it;
    if (fail || a.scal[3] != 0.0) a.sc[12] = 1.0;
  }
}

inline cudaError_t pcg_init() {
  cudaError_t e = cudaFuncSetAttribute(pcg_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pcg_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(pcg_pipe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096);
  return e;
}

}  // namespace ars
