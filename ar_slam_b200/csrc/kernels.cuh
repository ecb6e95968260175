// kernels.cuh -- bundle-adjustment kernels (sm_100a).
//
// Terminology: the two pose sets (captures, tags) play two roles per solve:
// the E side is eliminated by the Schur complement, the F side (plus the
// camera intrinsics) is retained.  Residual blocks are kept twice in HBM,
// once sorted by E pose and once sorted by F pose, as structure-of-arrays
// planes, so that each pass sees every pose's blocks as one contiguous
// segment and can reduce J^T J / J^T r inside the CTA (warp-shuffle + shared
// memory staging) instead of with global atomics.
#pragma once
#include <cstdint>
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include "model.cuh"

namespace ars {

constexpr int NV = 33;         // per-pose normal-equation record: H upper (21) | g (6) | H_pose,f (6)
constexpr int kAccumThreads = 128;
constexpr int kAccumWarps = kAccumThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Grid-wide, bit-reproducible reductions folded into the producing kernel.  v[] holds
// warp-reduced values (valid in lane 0 of every warp); every thread of the CTA must call.
//   cta_partial:           the CTA's totals -> part[blockIdx.x][M]
//   reduce_partials:       one CTA adds n partial rows in a fixed order -> out[0..M)
//   grid_reduce_last_cta:  both; the CTA that arrives last (atomic ticket, which wraps back to
//                          zero for the next launch) runs reduce_partials.  The ticket costs every
//                          CTA a fence + an atomic round trip, so it is used by the light kernels
//                          only; the heavy ones write partials and leave the sum to
//                          reduce_partials_kernel.
__host__ __device__ constexpr int grid_reduce_scratch_doubles(int M, int WARPS) { return (WARPS + (WARPS * 32) / M) * M + 2; }
// MAXCOL >= 0: that one column is combined with max instead of + (a sum-type reduction that carries one maximum)
template <int M, bool MAX, int WARPS, int MAXCOL = -1>
__device__ __forceinline__ bool cta_partial(const double (&v)[M], double* __restrict__ part, double* scratch) {
  double(*gr_sm)[M] = reinterpret_cast<double(*)[M]>(scratch);  // [WARPS][M]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < M; ++c) gr_sm[wid][c] = v[c];
  __syncthreads();
  if (threadIdx.x < M) {
    const bool mx = MAX || (int)threadIdx.x == MAXCOL;
    double t = gr_sm[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) t = mx ? fmax(t, gr_sm[w][threadIdx.x]) : t + gr_sm[w][threadIdx.x];
    __stcg(part + (size_t)blockIdx.x * M + threadIdx.x, t);
    return true;
  }
  return false;
}
template <int M, bool MAX, int THREADS, int MAXCOL = -1>
__device__ __forceinline__ void reduce_partials(const double* __restrict__ part, int n, double* __restrict__ out,
                                                double* scratch /* [THREADS / M][M] */) {
  // thread (g, c): partial rows g, g + G, ... of column c; then the G groups in order.  The rows are
  // loaded eight at a time before they are added (same order of additions): this kernel is one CTA
  // deep on the critical path of every LM iteration, and a load -> add chain per row made it
  // latency bound (13 us for 6250 rows, round 1).
  constexpr int G = THREADS / M;
  double(*gr_fin)[M] = reinterpret_cast<double(*)[M]>(scratch);
  const int c = threadIdx.x % M, g = threadIdx.x / M;
  const bool mx = MAX || c == MAXCOL;
  if (g < G) {
    double t = 0.0;
    for (int i = g; i < n; i += 8 * G) {
      double x[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = (i + k * G < n) ? __ldcg(part + (size_t)(i + k * G) * M + c) : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (i + k * G < n) t = mx ? fmax(t, x[k]) : t + x[k];
    }
    gr_fin[g][c] = t;
  }
  __syncthreads();
  if (threadIdx.x < M) {
    const bool mxo = MAX || (int)threadIdx.x == MAXCOL;
    double r = gr_fin[0][threadIdx.x];
    for (int q = 1; q < G; ++q) r = mxo ? fmax(r, gr_fin[q][threadIdx.x]) : r + gr_fin[q][threadIdx.x];
    out[threadIdx.x] = r;
  }
}
template <int M, bool MAX>
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double scratch[(1024 / M) * M];
  reduce_partials<M, MAX, 1024>(part, n, out, scratch);
}
template <int M, bool MAX, int WARPS>
__device__ __forceinline__ void grid_reduce_last_cta(const double (&v)[M], double* __restrict__ part,
                                                     unsigned* __restrict__ ticket, double* __restrict__ out,
                                                     double* scratch /* shared, grid_reduce_scratch_doubles(M, WARPS) */) {
  constexpr int G = (WARPS * 32) / M;
  int* gr_last = reinterpret_cast<int*>(scratch + (WARPS + G) * M);
  if (cta_partial<M, MAX, WARPS>(v, part, scratch)) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *gr_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (!*gr_last) return;
  __threadfence();
  reduce_partials<M, MAX, WARPS * 32>(part, (int)gridDim.x, out, scratch + WARPS * M);
}
template <int M, bool MAX, int WARPS>
__device__ __forceinline__ void grid_reduce_last_cta(const double (&v)[M], double* __restrict__ part,
                                                     unsigned* __restrict__ ticket, double* __restrict__ out) {
  __shared__ double gr_scratch[grid_reduce_scratch_doubles(M, WARPS)];
  grid_reduce_last_cta<M, MAX, WARPS>(v, part, ticket, out, gr_scratch);
}

// ---------------------------------------------------------------- prep -----
// both pose sides in one launch: CTAs [0, cap_ctas) prepare captures, the rest tags
// tag_cor (optional): the four world corners alone, 12 contiguous doubles per tag, for the residual-only
// kernels (six 128-bit loads instead of twelve scalar ones spread over the 384-byte record)
// cap_rt (optional): R | t alone, 12 contiguous doubles per capture, for the passes that never form the
// capture-rotation columns (accum_f_pipe_kernel)
// A thread's records are 96 .. 384 contiguous bytes: written directly, every warp-wide store would touch 32 half-filled
// sectors (ncu: lg_throttle 12 stall cycles per issue).  They go through a per-warp shared-memory stage instead, 32 consecutive
// 16-byte pieces per store instruction.
template <int N2, int STRIDE2>
__device__ __forceinline__ void warp_store_records(double2* __restrict__ dst_warp, const double2 (&v)[N2], double2* stage, int lane, int n_valid) {
#pragma unroll
  for (int k = 0; k < N2; ++k) stage[lane * N2 + k] = v[k];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    const int i = k * 32 + lane;
    const int t = i / N2, j = i - t * N2;
    if (t < n_valid) dst_warp[(size_t)t * STRIDE2 + j] = stage[i];
  }
  __syncwarp();
}
__global__ void __launch_bounds__(128) prep_poses_kernel(int n_cap, const double* __restrict__ cap_pose, double* __restrict__ cap_out,
                                                         int n_tag, const double* __restrict__ tag_pose, double tag_size,
                                                         double* __restrict__ tag_out, int cap_ctas, double* __restrict__ tag_cor,
                                                         double* __restrict__ cap_rt) {
  __shared__ double2 stage_all[4][32 * 12];
  const int lane = threadIdx.x & 31;
  double2* stage = stage_all[threadIdx.x >> 5];
  if ((int)blockIdx.x < cap_ctas) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int w0 = i - lane;                       // first capture of this warp
    const int n_valid = min(32, n_cap - w0);
    if (n_valid <= 0) return;
    double p[6], rec[kCapPre];
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = i < n_cap ? cap_pose[6 * (size_t)i + k] : 0.0;
    prep_capture(p, rec);
    double2 v[kCapPre / 2];
#pragma unroll
    for (int k = 0; k < kCapPre / 2; ++k) v[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
    warp_store_records<kCapPre / 2, kCapPre / 2>(reinterpret_cast<double2*>(cap_out + (size_t)kCapPre * w0), v, stage, lane, n_valid);
    if (cap_rt) {
      const double2 c[6] = {v[0], v[1], v[2], v[3], make_double2(rec[8], rec[18]), make_double2(rec[19], rec[20])};
      warp_store_records<6, 6>(reinterpret_cast<double2*>(cap_rt + (size_t)12 * w0), c, stage, lane, n_valid);
    }
  } else {
    const int i = (blockIdx.x - cap_ctas) * blockDim.x + threadIdx.x;
    const int w0 = i - lane;
    const int n_valid = min(32, n_tag - w0);
    if (n_valid <= 0) return;
    double p[6], rec[kTagPre];
#pragma unroll
    for (int k = 0; k < 6; ++k) p[k] = i < n_tag ? tag_pose[6 * (size_t)i + k] : 0.0;
    prep_tag(p, tag_size, rec);
    double2* o = reinterpret_cast<double2*>(tag_out + (size_t)kTagPre * w0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double2 v[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) v[k] = make_double2(rec[24 * half + 2 * k], rec[24 * half + 2 * k + 1]);
      warp_store_records<12, kTagPre / 2>(o + 12 * half, v, stage, lane, n_valid);
    }
    if (tag_cor) {
      const double2 c[6] = {make_double2(rec[0], rec[1]),   make_double2(rec[2], rec[12]),  make_double2(rec[13], rec[14]),
                            make_double2(rec[24], rec[25]), make_double2(rec[26], rec[36]), make_double2(rec[37], rec[38])};
      warp_store_records<6, 6>(reinterpret_cast<double2*>(tag_cor + (size_t)12 * w0), c, stage, lane, n_valid);
    }
  }
}

// ------------------------------------------------ kernel (1): evaluate -----
// One thread per observation corner, original block order.  Writes what
// ceres::Problem::Evaluate would: residuals and the three Jacobian blocks in
// Ceres' row-major layout.  Thread (b, i) owns rows 2i, 2i+1 of block b, i.e.
// 16 B of residuals, 48 B of jac_cam, 96 B of jac_cap and of jac_tag, all
// contiguous across consecutive threads (coalesced 128-bit stores).
// Algorithmic bytes per corner: 16 obs + 2 idx + 16 r + 240 J = 274.
// The outputs are written through a per-warp shared-memory transpose: a thread's rows are 16 / 48 / 96 / 96
// contiguous bytes, so direct 128-bit stores would leave every warp-wide store touching 32 half-filled
// sectors (ncu, round 2: 243 us, lg_throttle + mio_throttle 20 stall cycles per issue, L2 at 61 %).  Staged,
// every store instruction of a warp covers 512 contiguous bytes.
constexpr int kEvalThreads = 256;
template <int N2 /* double2 per thread */>
__device__ __forceinline__ void warp_store_rows(double2* __restrict__ dst_warp, const double2 (&v)[N2], double2* stage, int lane, int n_valid_threads) {
  // stage[thread][N2] -> dst_warp[0 .. 32 N2): the same linear order, written 32 consecutive double2 at a time
#pragma unroll
  for (int k = 0; k < N2; ++k) stage[lane * N2 + k] = v[k];
  __syncwarp();
  const int n = n_valid_threads * N2;
#pragma unroll
  for (int k = 0; k < N2; ++k) {
    const int i = k * 32 + lane;
    if (i < n) dst_warp[i] = stage[i];
  }
  __syncwarp();
}
template <int MODEL>
__global__ void __launch_bounds__(kEvalThreads)
eval_jacobian_kernel(int n_corner, const int32_t* __restrict__ cap_idx, const int32_t* __restrict__ tag_idx,
                     const double2* __restrict__ obs /* [n_corner] (x,y) */,
                     const double* __restrict__ cap_pre, const double* __restrict__ tag_pre,
                     const double* __restrict__ cam, double2* __restrict__ res /* may be null */,
                     double2* __restrict__ jac_cam, double2* __restrict__ jac_cap,
                     double2* __restrict__ jac_tag, double* __restrict__ warp_cost) {
  __shared__ double2 stage_all[kEvalThreads / 32][32 * 6];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  double2* stage = stage_all[threadIdx.x >> 5];
  const int warp_first = t - lane;                              // first corner of this warp
  const int n_valid = min(32, max(0, n_corner - warp_first));   // corners of this warp that exist
  double rr = 0.0;
  CornerJ j;
  double Kl[2][2] = {{0, 0}, {0, 0}};
  if (t < n_corner) {
    const int b = t >> 2, i = t & 3;
    const double cm[3] = {cam[0], cam[1], cam[2]};
    const double2 o = obs[t];
    const double* cp = cap_pre + (size_t)kCapPre * cap_idx[b];
    const double* tp = tag_pre + (size_t)kTagPre * tag_idx[b] + 12 * i;
    corner_jacobian_m<MODEL>(cp, tp, cm, o.x, o.y, j, Kl);
    rr = j.r[0] * j.r[0] + j.r[1] * j.r[1];
    if (res) res[t] = make_double2(j.r[0], j.r[1]);   // 16 B per thread: already contiguous across the warp
  }
  if (n_valid > 0) {
    if (jac_cam) {
      const double2 v[3] = {make_double2(j.K[0], Kl[0][0]), make_double2(Kl[0][1], j.K[1]), make_double2(Kl[1][0], Kl[1][1])};
      warp_store_rows<3>(jac_cam + 3 * (size_t)warp_first, v, stage, lane, n_valid);
    }
    if (jac_cap) {
      const double2 v[6] = {make_double2(j.A[0][0], j.A[0][1]), make_double2(j.A[0][2], j.B[0][0]), make_double2(j.B[0][1], j.B[0][2]),
                            make_double2(j.A[1][0], j.A[1][1]), make_double2(j.A[1][2], j.B[1][0]), make_double2(j.B[1][1], j.B[1][2])};
      warp_store_rows<6>(jac_cap + 6 * (size_t)warp_first, v, stage, lane, n_valid);
    }
    if (jac_tag) {
      const double2 v[6] = {make_double2(j.A[0][0], j.A[0][1]), make_double2(j.A[0][2], j.C[0][0]), make_double2(j.C[0][1], j.C[0][2]),
                            make_double2(j.A[1][0], j.A[1][1]), make_double2(j.A[1][2], j.C[1][0]), make_double2(j.C[1][1], j.C[1][2])};
      warp_store_rows<6>(jac_tag + 6 * (size_t)warp_first, v, stage, lane, n_valid);
    }
  }
  rr = warp_sum(rr);
  if (lane == 0) warp_cost[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = rr;
}

// Reduces the staged per-lane values ([value][lane], padded) over the runs of lanes that share
// a pose and writes one record per run: runs inside the warp go straight to out_seg, pieces
// that continue in a neighbouring warp go to `partial` (slot 0: the run started before this
// warp or fills it, slot 1: it starts inside and continues) for seg_fixup_kernel.
// sg[0] = pose of the lane before the warp, sg[1..32] = the lanes' poses, sg[33] = the next one.
template <int NVT>
__device__ __forceinline__ void segment_flush(const double (*st)[33], const int* sg, int own, int lane, int gwarp,
                                              double* __restrict__ out_seg, double* __restrict__ partial) {
  // run structure of this warp, identical for every lane: bit j of `ends` = lane j closes a run
  const unsigned ends = __ballot_sync(0xffffffffu, sg[lane + 2] != own);
  const bool head_open = sg[0] == sg[1];           // the first run started in the previous warp
  const bool tail_open = !((ends >> 31) & 1u);     // the last run continues in the next warp
  if (lane < NVT) {  // values 0..31: lane v walks the 32 columns of its value and flushes at every run end
    const int v = lane;
    const double* col = st[v];
    double acc = 0.0;
    unsigned m = ends;
    int j0 = 0;
    while (m) {
      const int j1 = __ffs(m) - 1;  // run [j0, j1]
      m &= m - 1;
      for (int j = j0; j <= j1; ++j) acc += col[j];
      const int sj = sg[j1 + 1];
      if (sj >= 0) {
        if (j0 == 0 && head_open) partial[((size_t)gwarp * 2 + 0) * NVT + v] = acc;
        else out_seg[(size_t)sj * NVT + v] = acc;
      }
      acc = 0.0;
      j0 = j1 + 1;
    }
    if (tail_open && sg[32] >= 0) {
      for (int j = j0; j < 32; ++j) acc += col[j];
      partial[((size_t)gwarp * 2 + (j0 == 0 ? 0 : 1)) * NVT + v] = acc;
    }
  }
  static_assert(NVT <= 33, "at most one value beyond the 32 lanes");
  if (NVT == 33) {  // value 32: a second walk would keep 31 lanes idle -> segmented inclusive scan over the lanes
    double x = st[NVT - 1][lane];
    const int start = lane == 0 ? 0 : 32 - __clz(ends & ((1u << lane) - 1u));  // first lane of this lane's run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane - o >= start) x += y;
    }
    const bool run_end = (ends >> lane) & 1u;
    if (run_end && own >= 0) {
      if (start == 0 && head_open) partial[((size_t)gwarp * 2 + 0) * NVT + 32] = x;
      else out_seg[(size_t)own * NVT + 32] = x;
    } else if (lane == 31 && tail_open && own >= 0) {
      partial[((size_t)gwarp * 2 + (start == 0 ? 0 : 1)) * NVT + 32] = x;
    }
  }
}

// -------------------------------------- kernels (1)+(2) fused: accumulate --
// One thread per residual block (4 corners), blocks sorted by the "own" pose
// side of this pass (SIDE 0: capture-sorted, SIDE 1: tag-sorted).  Per thread:
// evaluate the Jacobian pieces of its 4 corners in registers and accumulate
// the Gram entries it needs; then
//   * WITH_W: write the 6x6 cross block W = J_own^T J_other (36 coalesced
//     plane stores) and the warp's camera/cost partial;
//   * reduce the per-pose record (H_own,own upper | g_own | H_own,f; NV = 33
//     doubles) over the segment of threads that share the pose: values are
//     staged transposed in shared memory ([value][lane], padded), then lane v
//     walks the 32 columns and flushes a sum at every segment boundary.
//     Segments inside a warp are written straight to out_seg; pieces that
//     straddle warps go to `partial` and are summed in a fixed order by
//     seg_fixup_kernel, so the result is bit-reproducible.  No global atomics.
// J never touches HBM.  Algorithmic bytes per corner (E pass): 16 obs + 2 idx
// in, 72 W out; per pose 264 out.  (F pass: 18 in, 264 per pose out.)
struct AccumArgs {
  int n_blk;                 // blocks in this pass (sorted order)
  int plane;                 // plane stride (n_blk rounded up to 32)
  const int32_t* own_idx;    // [n_blk] pose index on the sorted side
  const int32_t* oth_idx;    // [n_blk] pose index on the other side
  const double* obs;         // 8 planes [k][plane]: x0,y0,x1,y1,x2,y2,x3,y3
  const double* cap_pre;
  const double* tag_pre;
  const double* cam;
  double* out_seg;           // [n_pose][NV]
  double* partial;           // [n_warp][2][NV]
  double* W;                 // 36 planes [e][plane] (WITH_W)
  double* warp_cam;          // [grid][4] CTA partials of sum K^2, sum K r, sum r^2, 0 (WITH_W)
};

template <int SIDE, bool WITH_W, int MODEL>
__global__ void __launch_bounds__(kAccumThreads, WITH_W ? 2 : 3) accum_kernel(const AccumArgs a) {
  // one shared buffer, two lives: during the corner loop it holds the per-thread cp.async
  // landing slots of the tag records (2 stages x 128 threads x 112 B, 16 B padding keeps the
  // 128-bit reads conflict free); afterwards it is the [value][lane] staging of the reduction
  __shared__ __align__(16) double raw[kAccumWarps * NV * 33];
  double(*stage)[NV][33] = reinterpret_cast<double(*)[NV][33]>(raw);
  __shared__ int sseg[kAccumWarps][36];
  constexpr int kSlot = 112;
  unsigned char* tslot = reinterpret_cast<unsigned char*>(raw) + threadIdx.x * kSlot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pos = blockIdx.x * kAccumThreads + threadIdx.x;
  const int gwarp = pos >> 5;
  const bool valid = pos < a.n_blk;
  const int own = valid ? a.own_idx[pos] : -1;
  const int oth = valid ? a.oth_idx[pos] : 0;
  sseg[wid][lane + 1] = own;
  if (lane == 0) {
    const int w0 = gwarp << 5;
    sseg[wid][0] = (w0 > 0 && w0 - 1 < a.n_blk) ? a.own_idx[w0 - 1] : -2;
    sseg[wid][33] = (w0 + 32 < a.n_blk) ? a.own_idx[w0 + 32] : -1;
  }

  // Gram accumulators.  own = [A | O2], other = [A | X2]; SIDE 0: O2 = B
  // (capture rotation), X2 = C (tag rotation); SIDE 1 the other way round.
  double AA[6] = {0, 0, 0, 0, 0, 0}, AO[9], OO[6] = {0, 0, 0, 0, 0, 0};
  double AX[9], OX[9];
  double Ar[3] = {0, 0, 0}, Or[3] = {0, 0, 0}, AK[3] = {0, 0, 0}, OK[3] = {0, 0, 0};
  double KK = 0.0, Kr = 0.0, rr = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) { AO[i] = 0.0; AX[i] = 0.0; OX[i] = 0.0; }

  if (valid) {
    const int cap = SIDE == 0 ? own : oth;
    const int tag = SIDE == 0 ? oth : own;
    const double cm[3] = {a.cam[0], MODEL ? a.cam[1] : 0.0, MODEL ? a.cam[2] : 0.0};
    double cp[kCapPre];
    {
      const double2* src = reinterpret_cast<const double2*>(a.cap_pre + (size_t)kCapPre * cap);
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const double2 v = __ldg(src + k);
        cp[2 * k] = v.x;
        cp[2 * k + 1] = v.y;
      }
    }
    const double2* tsrc = reinterpret_cast<const double2*>(a.tag_pre + (size_t)kTagPre * tag);
    // SIDE 0: the tag record is a random gather from L2 -> software pipeline with cp.async:
    // corner i + 1's 96 bytes fly into shared memory while corner i is being computed.
    if (SIDE == 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(tslot + 16 * k, tsrc + k, 16);
      __pipeline_memcpy_async(tslot + 96, a.obs + pos, 8);
      __pipeline_memcpy_async(tslot + 104, a.obs + (size_t)a.plane + pos, 8);
      __pipeline_commit();
    }
    double tn[12];  // SIDE 1: next corner's (own, warp-broadcast) tag record, prefetched in registers
    if (SIDE == 1) {
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double2 v = __ldg(tsrc + k);
        tn[2 * k] = v.x;
        tn[2 * k + 1] = v.y;
      }
    }
    double oxn = 0.0, oyn = 0.0;
    if (SIDE == 1) { oxn = a.obs[pos]; oyn = a.obs[(size_t)a.plane + pos]; }
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      double tp[12];
      double ox, oy;
      if (SIDE == 0) {
        if (i < 3) {
          unsigned char* nxt = tslot + ((i + 1) & 1) * (kAccumThreads * kSlot);
#pragma unroll
          for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(nxt + 16 * k, tsrc + 6 * (i + 1) + k, 16);
          __pipeline_memcpy_async(nxt + 96, a.obs + (size_t)(2 * i + 2) * a.plane + pos, 8);
          __pipeline_memcpy_async(nxt + 104, a.obs + (size_t)(2 * i + 3) * a.plane + pos, 8);
        }
        __pipeline_commit();
        __pipeline_wait_prior(1);
        const double2* cur = reinterpret_cast<const double2*>(tslot + (i & 1) * (kAccumThreads * kSlot));
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const double2 v = cur[k];
          tp[2 * k] = v.x;
          tp[2 * k + 1] = v.y;
        }
        const double2 o2 = cur[6];
        ox = o2.x;
        oy = o2.y;
      } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) tp[k] = tn[k];
        ox = oxn;
        oy = oyn;
        if (i < 3) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const double2 v = __ldg(tsrc + 6 * (i + 1) + k);
            tn[2 * k] = v.x;
            tn[2 * k + 1] = v.y;
          }
          oxn = a.obs[(size_t)(2 * i + 2) * a.plane + pos];
          oyn = a.obs[(size_t)(2 * i + 3) * a.plane + pos];
        }
      }
      CornerJ j;
      if (MODEL == 0) {
        corner_jacobian(cp, tp, cm[0], ox, oy, j);
      } else {
        double Kl[2][2];
        corner_jacobian_m<1>(cp, tp, cm, ox, oy, j, Kl);  // the l1, l2 columns belong to accum_cam_kernel
      }
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double* A = j.A[row];
        const double* O = SIDE == 0 ? j.B[row] : j.C[row];
        const double* X = SIDE == 0 ? j.C[row] : j.B[row];
        const double r = j.r[row], K = j.K[row];
        AA[0] += A[0] * A[0]; AA[1] += A[0] * A[1]; AA[2] += A[0] * A[2];
        AA[3] += A[1] * A[1]; AA[4] += A[1] * A[2]; AA[5] += A[2] * A[2];
        OO[0] += O[0] * O[0]; OO[1] += O[0] * O[1]; OO[2] += O[0] * O[2];
        OO[3] += O[1] * O[1]; OO[4] += O[1] * O[2]; OO[5] += O[2] * O[2];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            AO[p * 3 + q] += A[p] * O[q];
            if (WITH_W) {
              AX[p * 3 + q] += A[p] * X[q];
              OX[p * 3 + q] += O[p] * X[q];
            }
          }
          Ar[p] += A[p] * r;
          Or[p] += O[p] * r;
          AK[p] += A[p] * K;
          OK[p] += O[p] * K;
        }
        if (WITH_W) {
          KK += K * K;
          Kr += K * r;
          rr += r * r;
        }
      }
    }
    if (WITH_W) {
      // W = own^T other = [A^T A, A^T X2 ; O2^T A, O2^T X2], row-major 6x6 planes
      double* w = a.W + pos;
      const size_t ps = a.plane;
      const double aa[9] = {AA[0], AA[1], AA[2], AA[1], AA[3], AA[4], AA[2], AA[4], AA[5]};
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          w[(size_t)(p * 6 + q) * ps] = aa[p * 3 + q];
          w[(size_t)(p * 6 + 3 + q) * ps] = AX[p * 3 + q];
          w[(size_t)((p + 3) * 6 + q) * ps] = AO[q * 3 + p];
          w[(size_t)((p + 3) * 6 + 3 + q) * ps] = OX[p * 3 + q];
        }
    }
  }
  if (WITH_W) {
    const double v[4] = {warp_sum(KK), warp_sum(Kr), warp_sum(rr), 0.0};
    __shared__ double cam_scratch[kAccumWarps * 4];
    cta_partial<4, false, kAccumWarps>(v, a.warp_cam, cam_scratch);
  }
  if (SIDE == 0) __syncthreads();  // every thread is done with its cp.async slots before the buffer is reused
  // stage the per-pose record, transposed: stage[v][lane]
  {
    double(*st)[33] = stage[wid];
    // H upper packed, own = [A(0..2) | O2(3..5)]
    st[tri6(0, 0)][lane] = AA[0]; st[tri6(0, 1)][lane] = AA[1]; st[tri6(0, 2)][lane] = AA[2];
    st[tri6(1, 1)][lane] = AA[3]; st[tri6(1, 2)][lane] = AA[4]; st[tri6(2, 2)][lane] = AA[5];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q) st[tri6(p, 3 + q)][lane] = AO[p * 3 + q];
    st[tri6(3, 3)][lane] = OO[0]; st[tri6(3, 4)][lane] = OO[1]; st[tri6(3, 5)][lane] = OO[2];
    st[tri6(4, 4)][lane] = OO[3]; st[tri6(4, 5)][lane] = OO[4]; st[tri6(5, 5)][lane] = OO[5];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      st[21 + p][lane] = Ar[p];
      st[24 + p][lane] = Or[p];
      st[27 + p][lane] = AK[p];
      st[30 + p][lane] = OK[p];
    }
  }
  __syncwarp();
  segment_flush<NV>(stage[wid], sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
}

// ---- radial model only: the l1, l2 columns of the normal equations ---------------------------
// accum_kernel keeps its register budget for the pose blocks and the focal column; with the
// radial model this second pass (same thread -> block mapping, Jacobians recomputed) adds
//   * per pose: H_pose,l1 | H_pose,l2 (12 doubles, reduced with the same staged flush),
//   * WITH_CAM (the E pass): per warp the intrinsics block entries that involve l1, l2:
//     [f.l1, f.l2, l1.l1, l1.l2, l2.l2, l1.r, l2.r, 0].
constexpr int NVX = 12;
struct AccumCamArgs {
  int n_blk, plane;
  const int32_t* own_idx;
  const int32_t* oth_idx;
  const double* obs;
  const double* cap_pre;
  const double* tag_pre;
  const double* cam;
  double* out_seg;   // [n_pose][NVX]
  double* partial;   // [n_warp][2][NVX]
  double* warp_cam;  // [n_warp][8] (WITH_CAM)
};
template <int SIDE, bool WITH_CAM>
__global__ void __launch_bounds__(kAccumThreads) accum_cam_kernel(const AccumCamArgs a) {
  __shared__ double stage[kAccumWarps][NVX][33];
  __shared__ int sseg[kAccumWarps][36];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pos = blockIdx.x * kAccumThreads + threadIdx.x;
  const int gwarp = pos >> 5;
  const bool valid = pos < a.n_blk;
  const int own = valid ? a.own_idx[pos] : -1;
  const int oth = valid ? a.oth_idx[pos] : 0;
  sseg[wid][lane + 1] = own;
  if (lane == 0) {
    const int w0 = gwarp << 5;
    sseg[wid][0] = (w0 > 0 && w0 - 1 < a.n_blk) ? a.own_idx[w0 - 1] : -2;
    sseg[wid][33] = (w0 + 32 < a.n_blk) ? a.own_idx[w0 + 32] : -1;
  }
  double AL[3][2] = {{0, 0}, {0, 0}, {0, 0}}, OL[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  double cw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (valid) {
    const int cap = SIDE == 0 ? own : oth, tag = SIDE == 0 ? oth : own;
    const double cm[3] = {a.cam[0], a.cam[1], a.cam[2]};
    const double* cp = a.cap_pre + (size_t)kCapPre * cap;
    const double* tpb = a.tag_pre + (size_t)kTagPre * tag;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const double ox = a.obs[(size_t)(2 * i) * a.plane + pos];
      const double oy = a.obs[(size_t)(2 * i + 1) * a.plane + pos];
      CornerJ j;
      double Kl[2][2];
      corner_jacobian_m<1>(cp, tpb + 12 * i, cm, ox, oy, j, Kl);
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double* O = SIDE == 0 ? j.B[row] : j.C[row];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            AL[p][q] += j.A[row][p] * Kl[row][q];
            OL[p][q] += O[p] * Kl[row][q];
          }
        if (WITH_CAM) {
          cw[0] += j.K[row] * Kl[row][0];
          cw[1] += j.K[row] * Kl[row][1];
          cw[2] += Kl[row][0] * Kl[row][0];
          cw[3] += Kl[row][0] * Kl[row][1];
          cw[4] += Kl[row][1] * Kl[row][1];
          cw[5] += Kl[row][0] * j.r[row];
          cw[6] += Kl[row][1] * j.r[row];
        }
      }
    }
  }
  if (WITH_CAM) {
#pragma unroll
    for (int i = 0; i < 7; ++i) cw[i] = warp_sum(cw[i]);
    if (lane == 0) {
      double* wc = a.warp_cam + 8 * (size_t)gwarp;
#pragma unroll
      for (int i = 0; i < 8; ++i) wc[i] = cw[i];
    }
  }
  // record: [pose dof p (A: 0..2, O: 3..5)][l1] at p, [..][l2] at 6 + p
  double(*st)[33] = stage[wid];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    st[p][lane] = AL[p][0];
    st[3 + p][lane] = OL[p][0];
    st[6 + p][lane] = AL[p][1];
    st[9 + p][lane] = OL[p][1];
  }
  __syncwarp();
  segment_flush<NVX>(stage[wid], sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
}

// Sums the pieces of poses whose blocks straddle warps (fixed order), and zero-fills poses without
// blocks.  A segment can only be cut where a warp ends, so the work is indexed by warp boundary:
// thread (w, v) looks at the boundary between sorted positions 32 w - 1 and 32 w and, when a
// segment crosses it for the FIRST time there, adds that segment's pieces of value v in warp order.
// (Round 1 ran one thread per (pose, value) -- 3.3 M threads per side at config 3, 13.6 us each.)
// Up to four records are fixed in one launch: both pose sides and, for the radial model, their
// l1 / l2 border records.
struct FixupJob {
  int n_pose, n_blk, nv;
  const int32_t* own_idx;   // [n_blk] sorted pose index
  const int32_t* seg_off;   // [n_pose] first sorted position of the pose's segment
  const int32_t* seg_end;   // [n_pose] one past its last position (== seg_off + 1 when the segments are in index order)
  const double* partial;    // [n_warp][2][nv]
  double* out_seg;          // [n_pose][nv]
  int cta_end;              // this job owns CTAs [previous cta_end, cta_end)
  int cta_zero;             // of those, CTAs >= cta_zero zero-fill the poses without blocks
};
struct FixupJobs {
  FixupJob j[4];
  int n;
};
__global__ void __launch_bounds__(256) seg_fixup_kernel(const FixupJobs jobs) {
  int q = 0;
  while (q + 1 < jobs.n && (int)blockIdx.x >= jobs.j[q].cta_end) ++q;
  const FixupJob& f = jobs.j[q];
  const int cta0 = q ? jobs.j[q - 1].cta_end : 0;
  if ((int)blockIdx.x >= f.cta_zero) {  // poses without blocks: their records are never written by the accumulation
    const int s = ((int)blockIdx.x - f.cta_zero) * 256 + threadIdx.x;
    if (s < f.n_pose && f.seg_end[s] == f.seg_off[s])
      for (int v = 0; v < f.nv; ++v) f.out_seg[(size_t)s * f.nv + v] = 0.0;
    return;
  }
  const int t = ((int)blockIdx.x - cta0) * 256 + threadIdx.x;
  const int w = 1 + t / f.nv, v = t % f.nv;
  const int blk = w << 5;
  if (blk >= f.n_blk) return;
  const int s = f.own_idx[blk - 1];
  if (f.own_idx[blk] != s) return;             // no segment crosses this boundary
  const int b0 = f.seg_off[s], b1 = f.seg_end[s];
  const int w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
  if (w0 != w - 1) return;                     // the segment was already open at the previous boundary
  double acc = 0.0;
  for (int ww = w0; ww <= w1; ++ww) {
    const int slot = (ww == w0 && (b0 & 31) != 0) ? 1 : 0;
    acc += f.partial[((size_t)ww * 2 + slot) * f.nv + v];
  }
  f.out_seg[(size_t)s * f.nv + v] = acc;
}

// Deterministic column sums of a [n][m] array (m <= 12) into out[m]: every CTA
// reduces one contiguous chunk in a fixed order (stage 1), one CTA adds the
// chunk sums (stage 2).  Same result on every run, no atomics.
constexpr int kColsumChunks = 128;
__global__ void __launch_bounds__(256) colsum_stage1_kernel(int n, int m, const double* __restrict__ in,
                                                            double* __restrict__ part /* [chunks][12] */) {
  __shared__ double sm[8][12];
  const int chunk = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  double acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lo + threadIdx.x; i < hi; i += 256)
#pragma unroll
    for (int c = 0; c < 12; ++c) if (c < m) acc[c] += in[(size_t)i * m + c];
#pragma unroll
  for (int c = 0; c < 12; ++c) acc[c] = warp_sum(acc[c]);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int c = 0; c < 12; ++c) sm[threadIdx.x >> 5][c] = acc[c];
  __syncthreads();
  if (threadIdx.x < 12) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
    part[blockIdx.x * 12 + threadIdx.x] = t;
  }
}
__global__ void colsum_stage2_kernel(int chunks, int m, const double* __restrict__ part, double* __restrict__ out) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;  // one warp per column
  if (c >= m) return;
  double t = 0.0;
  for (int i = lane; i < chunks; i += 32) t += part[i * 12 + c];
  t = warp_sum(t);
  if (lane == 0) out[c] = t;
}
// small inputs: one CTA
__global__ void __launch_bounds__(1024) colsum_kernel(int n, int m, const double* __restrict__ in,
                                                      double* __restrict__ out) {
  __shared__ double sm[1024];
  for (int c = 0; c < m; ++c) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += in[(size_t)i * m + c];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = sm[0];
    __syncthreads();
  }
}

// ------------------------------------------------------------ candidate -----
// Cost at the candidate point x + delta: one thread per block (E-sorted order),
// residuals only.  warp partials of sum r^2 -> colsum_kernel.  Bytes per corner: 18 in.
struct CandArgs {
  int n_blk, plane;
  const int32_t* own_idx;
  const int32_t* oth_idx;
  const double* obs;
  const double* cap_pre_c; // prep records at x + delta
  const double* tag_cor_c; // [n_tag][12] world corners at x + delta
  const double* cam_c;
  double* warp_out;        // [grid] CTA partials of the candidate's sum r^2
  unsigned* ticket;
  double* out;             // [1]
};
template <int SIDE, int MODEL>
__global__ void __launch_bounds__(256, 4) candidate_kernel(const CandArgs a) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  double c2 = 0.0;
  if (pos < a.n_blk) {
    const int own = a.own_idx[pos], oth = a.oth_idx[pos];
    const int cap = SIDE == 0 ? own : oth, tag = SIDE == 0 ? oth : own;
    const double cm[3] = {a.cam_c[0], MODEL ? a.cam_c[1] : 0.0, MODEL ? a.cam_c[2] : 0.0};
    double cp[kCapPre];
    {
      const double2* src = reinterpret_cast<const double2*>(a.cap_pre_c + (size_t)kCapPre * cap);
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const double2 v = __ldg(src + k);
        cp[2 * k] = v.x;
        cp[2 * k + 1] = v.y;
      }
    }
    double wc[12];
    {
      const double2* src = reinterpret_cast<const double2*>(a.tag_cor_c + (size_t)12 * tag);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double2 v = __ldg(src + k);
        wc[2 * k] = v.x;
        wc[2 * k + 1] = v.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double ox = a.obs[(size_t)(2 * i) * a.plane + pos];
      const double oy = a.obs[(size_t)(2 * i + 1) * a.plane + pos];
      double tp3[3] = {wc[3 * i], wc[3 * i + 1], wc[3 * i + 2]};
      double rc[2];
      corner_residual_m<MODEL>(cp, tp3, cm, ox, oy, rc);
      c2 += rc[0] * rc[0] + rc[1] * rc[1];
    }
  }
  const double v[1] = {warp_sum(c2)};
  grid_reduce_last_cta<1, false, 8>(v, a.warp_out, a.ticket, a.out);
}

// ------------------------------------------------------------ LM vectors ---
// sigma = 1 / (1 + sqrt(H_jj)) once at iteration 0 (Ceres jacobi_scaling).
// One thread per pose; rec = out_seg layout.
// A constant pose (arslam_set_constant; ceres::Problem::SetParameterBlockConstant) gets sigma = 0:
// its scaled block is then D^2 alone, its gradient and every cross term vanish, so its row of the
// (reduced) system decouples and its step is exactly zero -- the same system Ceres solves after
// removing the block from the program.
__global__ void sigma_pose_kernel(int n_pose, const double* __restrict__ rec, int enabled,
                                  double* __restrict__ sigma, const unsigned char* __restrict__ constant) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pose) return;
  const bool fixed = constant && constant[i];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double h = rec[(size_t)i * NV + tri6(k, k)];
    sigma[6 * (size_t)i + k] = fixed ? 0.0 : (enabled ? 1.0 / (1.0 + sqrt(h)) : 1.0);
  }
}

// Small device-resident scalar block shared by the LM kernels (kNumScalars doubles; the
// in-kernel grid reductions write straight into its fields).
struct LmScalars {
  double cam_H;      // [0] sum K^2   (J^T J of the focal length)
  double cam_g;      // [1] sum K r
  double sum_r2;     // [2] sum r^2 at x
  double unused0;    // [3]
  double cross;      // [4] sum over blocks delta_e^T W delta_f
  double cand_r2;    // [5] sum r^2 at x + delta
  double step2_e;    // [6] ||delta||^2, E poses
  double xnorm2_e;   // [7] ||x||^2, E poses that own blocks
  double mq_e;       // [8] sum g.d + d^T H d / 2 + d_f H_e,f.d over E poses
  double step2_f;    // [9]
  double xnorm2_f;   // [10]
  double mq_f;       // [11]
  double chol_fail;  // [12] != 0: a factorisation failed (non-positive pivot)
  double d_cam;      // [13] step of the focal length
  double sigma_f;    // [14] Jacobi scale of the focal length
  double focal;      // [15] current focal length
  double gmax_e;     // [16] max |g_i|, E poses
  double gmax_f;     // [17]
  double pcg_iters;  // [18]
  double pad[5];     // [19..23]
  // radial model (num_intrinsics = 3) only
  double H_f_l1, H_f_l2, H_l1_l1, H_l1_l2, H_l2_l2;  // [24..28] intrinsics block entries with l1, l2
  double g_l1, g_l2;                                  // [29], [30]
  double unused2;                                     // [31]
  double d_l1, d_l2;                                  // [32], [33] steps
  double sigma_l1, sigma_l2;                          // [34], [35] Jacobi scales
  double l1, l2;                                      // [36], [37] current values
  double pad2[2];
};
constexpr int kNumScalars = 40;

}  // namespace ars
