// schur_local.cuh -- kernel (3), block-sparse target: Schur elimination with the products
// pre-reduced inside the CTA.
//
// schur_eliminate_kernel (schur.cuh) walks the E poses in index order; every pair of blocks of an
// E pose sends one 6x6 product to the reduced system, 3.6 M reductions of 36 doubles at config 3,
// and ncu shows L2 (62 % of peak, round 2) busy with exactly those.  Two E poses add to the same
// block of S whenever they see the same pair of F poses -- which neighbouring captures do all the
// time -- but in index order such captures never meet in one CTA (distinct destinations per CTA:
// 99.4 % of its products).
//
// The E-sorted copy of the blocks is therefore stored in LOCALITY order (arslam.cu, rebuild_views: E
// poses sorted by their smallest F pose -- all E poses that share it see overlapping sets of F
// poses), and this kernel packs whole segments, up to 128 blocks, into a CTA.  A plan, built once
// per problem on the device, holds the CTA's products sorted by destination block as (first,
// second) pairs of local block ids whose V_first^T V_second land there.  The kernel computes
// V_j = L^-1 (sig_e W_j) as before, publishes it in shared memory, and then every thread takes an
// equal share of the CTA's pair list, adds the products of a run of equal destinations in
// registers and sends one reduction per run.  Config 3: 3.6 M products, 0.59 M distinct
// (CTA, destination) pairs; no segment straddles a CTA, so the second launch disappears too.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include "pcg.cuh"
#include "schur.cuh"

namespace ars {

constexpr int kSlThreads = 128;
constexpr int kSlMaxSeg = 33;  // longest E segment the plan accepts (the CTA keeps kSlThreads - (kmax - 1) slots as packing target)

struct SchurLocalView {
  int T;                     // packing target: CTA c owns the segments that START in positions [c T, (c + 1) T)
  int n_cta, stride;         // launch index b works on range (b * stride) mod n_cta: CTAs that run at the same time are far
                             // apart in the locality order, so that their reductions do not meet on the same blocks of S
  const int32_t* cta_pair;   // [n_cta + 1]   the CTA's range in the pair list (sorted by destination inside the CTA)
  const int32_t* pair_dst;   // [n_pairs]     compact lower-block index of the pair's destination
  const uint16_t* pairs;     // [n_pairs]     first | second << 7 | symmetrise << 14 (local block ids)
};

struct SchurLocalPlan {
  bool valid = false;
  int n_cta = 0;
  long long n_pairs = 0;
  PcgBuf<int32_t> cta_pair, pair_dst;
  int T = 0, stride = 1;
  PcgBuf<uint16_t> pairs[2];
  // scratch of the build
  PcgBuf<unsigned long long> keys[2];
  PcgBuf<int32_t> cnt, dev_ints;
  PcgBuf<unsigned char> tmp;
  int* h_ints = nullptr;  // pinned: kmax, n_items
  ~SchurLocalPlan() { if (h_ints) cudaFreeHost(h_ints); }
  SchurLocalView view() const {
    SchurLocalView v;
    v.T = T; v.n_cta = n_cta; v.stride = stride; v.cta_pair = cta_pair.p; v.pair_dst = pair_dst.p; v.pairs = pairs[0].p;
    return v;
  }
};

// ---- plan construction (device) ------------------------------------------------------------------
__global__ void sl_counts_kernel(int n_e, const int32_t* __restrict__ e_off, const int32_t* __restrict__ e_end, int32_t* __restrict__ cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_e) cnt[e] = e_end[e] - e_off[e];
}
// (CTA, local slot) of the block at E-sorted position pos: a segment belongs to the CTA its FIRST block falls into and
// ends before slot T + kmax - 1 <= 128
__device__ __forceinline__ int sl_loc(int pos, int seg_start, int T) {
  const int c = seg_start / T;
  return c * kSlThreads + (pos - c * T);
}
// one (key, pair) per product, same enumeration as schur_eliminate_kernel / pair_slot_kernel
__global__ void sl_emit_pairs_kernel(int n_blk, int T, const int32_t* __restrict__ e_idx, const int32_t* __restrict__ e_off,
                                     const int32_t* __restrict__ e_end, const int32_t* __restrict__ f_idx,
                                     const int32_t* __restrict__ pair_off, const int32_t* __restrict__ pair_slot,
                                     unsigned long long nnz_lower, unsigned long long* __restrict__ keys, uint16_t* __restrict__ vals) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_blk) return;
  const int e = e_idx[pos];
  const int beg = e_off[e], k = e_end[e] - beg, j = pos - beg;
  const int fj = f_idx[pos];
  const int loc = sl_loc(pos, beg, T);
  const unsigned long long cta = (unsigned long long)(loc >> 7);
  const unsigned la = loc & 127;
  const int np = schur_pairs_of(j, k);
  for (int d = 0; d < np; ++d) {
    int i2 = j + d;
    if (i2 >= k) i2 -= k;
    const int ppos = beg + i2;
    const int fp = f_idx[ppos];
    const unsigned lb = sl_loc(ppos, beg, T) & 127;
    const bool own_first = fj <= fp, diag = fj == fp;
    // the lower block (row = larger F pose, col = smaller) holds V_first^T V_second with first = the block of the
    // LARGER F pose, exactly as schur_eliminate_kernel stores it (own first and not diagonal -> its m1 transposed,
    // i.e. partner^T own)
    const bool swap = own_first && !diag;
    const unsigned first = swap ? lb : la, second = swap ? la : lb;
    const unsigned sym = (diag && d != 0) ? 1u : 0u;  // one capture sees a tag twice: M + M^T share the diagonal block
    const size_t o = (size_t)pair_off[pos] + d;
    keys[o] = cta * nnz_lower + (unsigned long long)pair_slot[o];
    vals[o] = (uint16_t)(first | (second << 7) | (sym << 14));
  }
}
__global__ void sl_pair_dst_kernel(long long n, const unsigned long long* __restrict__ keys, unsigned long long nnz_lower,
                                   int32_t* __restrict__ pair_dst) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) pair_dst[i] = (int32_t)(keys[i] % nnz_lower);
}
__global__ void sl_cta_pair_kernel(int n_cta, long long n, const unsigned long long* __restrict__ keys, unsigned long long nnz_lower,
                                   int32_t* __restrict__ cta_pair) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cta) return;
  const unsigned long long want = (unsigned long long)c * nnz_lower;
  long long lo = 0, hi = n;
  while (lo < hi) {  // first pair whose key is >= c * nnz_lower
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  cta_pair[c] = (int32_t)lo;
}

#define SL_TRY(call)                                                   \
  do {                                                                 \
    const cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) { err = std::string("schur plan: ") + cudaGetErrorString(e__); return 1; } \
  } while (0)

// Builds the plan for the E-sorted blocks (e_idx, e_off, f_idx) whose pair slots the PCG symbolic phase
// has already computed (ws.pair_off / ws.pair_slot).  Returns 0 and plan.valid == false when the problem does
// not qualify (a segment longer than kSlMaxSeg, or no pair-slot table): the caller then uses schur_eliminate_kernel.
inline int schur_local_build(SchurLocalPlan& plan, const PcgWorkspace& ws, int n_e, int n_blk, const int32_t* e_idx,
                             const int32_t* e_off, const int32_t* e_end, const int32_t* f_idx, int n_sm, cudaStream_t st, std::string& err) {
  plan.valid = false;
  if (!ws.pair_slot || ws.n_pairs <= 0 || ws.n_pairs >= (1LL << 31) || n_blk <= 0) return 0;
  const long long np = ws.n_pairs;
  if (!plan.h_ints) SL_TRY(cudaMallocHost((void**)&plan.h_ints, 4 * sizeof(int)));
  SL_TRY(plan.keys[0].ensure((size_t)np)); SL_TRY(plan.keys[1].ensure((size_t)np));
  SL_TRY(plan.cnt.ensure(n_e)); SL_TRY(plan.dev_ints.ensure(4));
  // 1. the longest segment decides how many slots of a CTA are the packing target
  size_t t3 = 0, t4 = 0;
  sl_counts_kernel<<<(n_e + 255) / 256, 256, 0, st>>>(n_e, e_off, e_end, plan.cnt.p);
  SL_TRY(cub::DeviceReduce::Max(nullptr, t3, plan.cnt.p, plan.dev_ints.p, n_e, st));
  SL_TRY(plan.tmp.ensure(t3));
  SL_TRY(cub::DeviceReduce::Max(plan.tmp.p, t3, plan.cnt.p, plan.dev_ints.p, n_e, st));
  SL_TRY(cudaMemcpyAsync(plan.h_ints, plan.dev_ints.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SL_TRY(cudaStreamSynchronize(st));
  const int kmax = plan.h_ints[0];
  if (kmax > kSlMaxSeg) return 0;
  const int T = kSlThreads - (std::max(kmax, 1) - 1);
  plan.T = T;
  plan.n_cta = (n_blk - 1) / T + 1;
  {
    // co-resident CTAs (3 per SM) are spread evenly over the locality order: stride ~ n_cta / resident, coprime with n_cta
    auto gcd = [](long long a, long long b) { while (b) { const long long t = a % b; a = b; b = t; } return a; };
    long long st2 = std::max(1, plan.n_cta / std::max(1, 3 * n_sm));
    while (gcd(st2, plan.n_cta) != 1) ++st2;
    plan.stride = (int)st2;
  }
  // 3. products keyed by (CTA, destination), sorted
  SL_TRY(plan.pairs[0].ensure((size_t)np)); SL_TRY(plan.pairs[1].ensure((size_t)np));
  SL_TRY(plan.pair_dst.ensure((size_t)np)); SL_TRY(plan.cta_pair.ensure((size_t)plan.n_cta + 1));
  const unsigned long long nnz_lower = (unsigned long long)std::max(ws.nnz_lower, 1);
  sl_emit_pairs_kernel<<<(n_blk + 127) / 128, 128, 0, st>>>(n_blk, T, e_idx, e_off, e_end, f_idx, ws.pair_off, ws.pair_slot, nnz_lower,
                                                          plan.keys[0].p, plan.pairs[1].p);
  int bits = 1;
  while (bits < 64 && ((unsigned long long)plan.n_cta * nnz_lower) >> bits) ++bits;
  SL_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t4, plan.keys[0].p, plan.keys[1].p, plan.pairs[1].p, plan.pairs[0].p, (int)np, 0, bits, st));
  SL_TRY(plan.tmp.ensure(t4));
  SL_TRY(cub::DeviceRadixSort::SortPairs(plan.tmp.p, t4, plan.keys[0].p, plan.keys[1].p, plan.pairs[1].p, plan.pairs[0].p, (int)np, 0, bits, st));
  // 4. per pair its destination, per CTA its range of the list
  sl_pair_dst_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(np, plan.keys[1].p, nnz_lower, plan.pair_dst.p);
  sl_cta_pair_kernel<<<(plan.n_cta + 1 + 255) / 256, 256, 0, st>>>(plan.n_cta, np, plan.keys[1].p, nnz_lower, plan.cta_pair.p);
  SL_TRY(cudaGetLastError());
  plan.n_pairs = np;
  plan.valid = true;
  return 0;
}
#undef SL_TRY

// ---- the kernel ---------------------------------------------------------------------------------
// V rows in shared memory: row r (local block id) starts at r * 37 + (r >> 3) doubles.  Rows of the same destination
// come from the same slot of different E poses, i.e. 8 rows apart when every capture sees 8 tags; with a plain odd
// stride those rows fall on two banks (16-way conflicts: 437 us); the extra (r >> 3) spreads them over all sixteen.
constexpr int kSlVsDoubles = kSlThreads * 37 + kSlThreads / 8;
__device__ __forceinline__ int sl_row(int r) { return r * 37 + (r >> 3); }
constexpr size_t kSlSmem = (size_t)(kSlVsDoubles + kSlThreads * 12 + 64) * sizeof(double);

// Phase 1 is schur_eliminate_kernel's: the thread of block j rebuilds the damped 6x6 block of its E pose, factors
// it in registers, computes V_j = L^-1 (sig_e W_j) and publishes it; the segment's first thread also writes z, yb
// and the camera terms.  Phase 2: every thread takes an equal share of the CTA's pair list (sorted by destination).
template <bool BULK /* unused: the few reductions that are left go out as per-lane FP64 reductions */>
__global__ void __launch_bounds__(kSlThreads, 3)
schur_local_kernel(const SchurArgs a, const SparseTarget t, int n_blk, const int32_t* __restrict__ e_idx, const SchurLocalView v) {
  extern __shared__ __align__(16) double sl_sm[];
  double* Vs = sl_sm;                                                            // 128 swizzled rows of 36
  double(*cs)[12] = reinterpret_cast<double(*)[12]>(sl_sm + kSlVsDoubles);       // camera terms of the segments
  // local slot l holds E-sorted position c T + l when that block's segment starts inside this CTA's range
  // (slots beyond T finish the last segment; the first slots may belong to the previous CTA's last segment)
  const int cta = (int)(((long long)blockIdx.x * v.stride) % v.n_cta);
  const int pos = cta * v.T + threadIdx.x;
  const int e = pos < n_blk ? e_idx[pos] : 0;
  const int beg = pos < n_blk ? a.e_off[e] : 0;
  const bool valid = pos < n_blk && beg / v.T == cta;
  const int j = valid ? pos - beg : 0;
  const size_t ps = a.plane;
  // the pair list of this thread (phase 2) starts travelling now
  const int pb = v.cta_pair[cta], pe_cta = v.cta_pair[cta + 1];
  const int chunk = (pe_cta - pb + kSlThreads - 1) / kSlThreads;
  int p = pb + threadIdx.x * chunk;
  const int pe = min(p + chunk, pe_cta);
  int nxt_dst = p < pe ? v.pair_dst[p] : -1;
  unsigned nxt_pr = p < pe ? v.pairs[p] : 0u;
  if (!(valid && j == 0)) {
#pragma unroll
    for (int i = 0; i < 12; ++i) cs[threadIdx.x][i] = 0.0;
  }
  if (valid) {
    double L[36], V[36], s[6];
#pragma unroll
    for (int q = 0; q < 36; ++q) V[q] = a.W[(size_t)q * ps + pos];  // in flight during the factorisation
    double zl[6], ybl[6], hk[6];
    load_scaled_E(a, e, L, zl, hk, s);
    const bool ok = chol6(L);
    const int fj = a.f_idx[pos];
#pragma unroll
    for (int i = 0; i < 6; ++i) ybl[i] = hk[i];
    chol6_forward(L, zl);
    chol6_forward(L, ybl);
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
        a.YB[6 * (size_t)e + i] = yb[i];
        m00 += hk[i] * yb[i];
        v0 += hk[i] * z[i];
      }
      zo[6] = ok ? 0.0 : 1.0;
      zo[7] = 0.0;
#pragma unroll
      for (int i = 0; i < 12; ++i) sg[i] = 0.0;
      sg[0] = m00; sg[6] = v0; sg[9] = ok ? 0.0 : 1.0;
      if (!(a.e_const && a.e_const[e])) {
        const double* rec = a.HE + (size_t)e * NV;
        double gm = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) gm = fmax(gm, fabs(rec[21 + i]));
        sg[10] = gm;
      }
    }
    double* myV = Vs + sl_row(threadIdx.x);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double col[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) col[i] = V[i * 6 + c] * s[i];
      chol6_forward(L, col);
      double b0 = 0.0, b1 = 0.0;  // (sig_e W)^T Ht^-1 h = V^T (L^-1 h)
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        myV[i * 6 + c] = col[i];
        b0 += col[i] * ybl[i];
        b1 += col[i] * zl[i];
      }
      t.add_border(fj, c, b0, b1);
    }
  }
  __syncthreads();
  {
    double cm[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) cm[i] = (i == 0 || i == 6 || i == 9) ? warp_sum(cs[threadIdx.x][i]) : 0.0;
    cm[10] = warp_max(cs[threadIdx.x][10]);
    cta_partial<12, false, kSlThreads / 32, 10>(cm, a.seg_cam, sl_sm + kSlVsDoubles + kSlThreads * 12);
  }
  // ---- phase 2: the CTA's pair list (sorted by destination) is cut into 128 equal chunks, so every thread multiplies
  // the same number of pairs (one thread per DESTINATION left most of the CTA idle behind its largest destination:
  // 539 us at config 3).  A thread adds the products of a run of equal destinations in registers and sends one
  // reduction per run; runs cut by a chunk boundary are completed by the L2 reduction.
  double acc[36];
  int cur = -1;
  auto flush = [&]() {
    double* dst = t.Sraw + 36 * (size_t)cur;
#pragma unroll
    for (int q = 0; q < 36; ++q) red_add_f64(dst + q, acc[q]);
  };
  for (; p < pe; ++p) {
    const int dst_slot = nxt_dst;
    const unsigned pr = nxt_pr;
    if (p + 1 < pe) { nxt_dst = v.pair_dst[p + 1]; nxt_pr = v.pairs[p + 1]; }  // next pair's words fly during this product
    if (dst_slot != cur) {
      if (cur >= 0) flush();
      cur = dst_slot;
#pragma unroll
      for (int q = 0; q < 36; ++q) acc[q] = 0.0;
    }
    const double* A = Vs + sl_row(pr & 127u);
    const double* B = Vs + sl_row((pr >> 7) & 127u);
    const bool sym = (pr >> 14) != 0;
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      double am[6], bm[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) { am[q] = A[m * 6 + q]; bm[q] = B[m * 6 + q]; }
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[r * 6 + c] += am[r] * bm[c];
      if (sym) {
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[r * 6 + c] += bm[r] * am[c];
      }
    }
  }
  if (cur >= 0) flush();
}

}  // namespace ars
