// accum_pipe.cuh -- kernels (1)+(2), software-pipelined across residual blocks.
//
// Same arithmetic, thread -> block mapping, summation order and outputs as accum_kernel
// (kernels.cuh): the records and W written here are bit-identical to it.  What changes is how
// the operands arrive.  accum_kernel starts every block cold: index load -> dependent gather of
// the pose records -> first corner waits a full L2 round trip, with two warps per scheduler and
// nothing to hide it (ncu round 1: long_scoreboard 23 % / 50 % of the stalls of the E / F pass).
// Here every CTA walks a list of 128-block chunks and each thread keeps a ring of cp.async
// landing slots in shared memory that runs AHEAD ACROSS BLOCKS: while corner c is computed, the
// records of corner c + 2 (possibly already the next chunk's block) are in flight, and the next
// chunk's indices were loaded one whole block earlier.  No register ever waits on global memory
// inside the corner loop.
//
//   accum_e_pipe_kernel<MODEL>  capture-sorted pass that also writes W (SIDE 0, WITH_W)
//   accum_f_pipe_kernel<MODEL>  tag-sorted pass (SIDE 1, no W)
// The two other side/role combinations (tags eliminated) stay on accum_kernel.
#pragma once
#include "kernels.cuh"

namespace ars {

// segment_flush of kernels.cuh for a slice [V0, V0 + NVT) of the NV-wide record, so that the
// transposed staging buffer only has to hold NVT <= 17 rows at a time -- and without its serial
// load -> add chain: ncu (round 2, profiles/r2_accum_stalls.txt) put a third of accum_kernel's time
// into that walk (a shared-memory round trip per column, loop bounds that depend on the run
// structure).  Here lane v first pulls its whole row (32 columns, 16 x 128-bit loads in flight
// together; row stride kStageLd = 34 keeps them aligned and conflict free), then adds the columns in
// the same left-to-right order and flushes at every run end; the branches are warp-uniform.
constexpr int kStageLd = 34;
template <int NVT, int V0>
__device__ __forceinline__ void segment_flush_slice(const double (*st)[kStageLd], const int* sg, int own, int lane, int gwarp,
                                                    double* __restrict__ out_seg, double* __restrict__ partial) {
  static_assert(NVT <= 32, "one lane per value");
  const unsigned ends = __ballot_sync(0xffffffffu, sg[lane + 2] != own);  // bit j: lane j closes a run
  if (lane >= NVT) return;
  const bool head_open = sg[0] == sg[1];           // the first run started in the previous warp
  const bool tail_open = !((ends >> 31) & 1u);     // the last run continues in the next warp
  const int v = lane;
  double x[32];
  {
    const double2* row = reinterpret_cast<const double2*>(st[v]);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double2 t = row[j];
      x[2 * j] = t.x;
      x[2 * j + 1] = t.y;
    }
  }
  double acc = 0.0;
  bool first = true;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    acc += x[j];
    if ((ends >> j) & 1u) {
      const int sj = sg[j + 1];
      if (sj >= 0) {
        if (first && head_open) partial[((size_t)gwarp * 2 + 0) * NV + V0 + v] = acc;
        else out_seg[(size_t)sj * NV + V0 + v] = acc;
      }
      acc = 0.0;
      first = false;
    }
  }
  if (tail_open && sg[32] >= 0) partial[((size_t)gwarp * 2 + (first ? 0 : 1)) * NV + V0 + v] = acc;
}

// the serial walk of kernels.cuh's segment_flush (run by run, loads and adds chained), same order of additions
template <int NVT, int V0>
__device__ __forceinline__ void segment_flush_slice_serial(const double (*st)[kStageLd], const int* sg, int own, int lane, int gwarp,
                                                           double* __restrict__ out_seg, double* __restrict__ partial) {
  const unsigned ends = __ballot_sync(0xffffffffu, sg[lane + 2] != own);
  if (lane >= NVT) return;
  const bool head_open = sg[0] == sg[1];
  const bool tail_open = !((ends >> 31) & 1u);
  const int v = lane;
  const double* col = st[v];
  double acc = 0.0;
  unsigned m = ends;
  int j0 = 0;
  while (m) {
    const int j1 = __ffs(m) - 1;
    m &= m - 1;
    for (int j = j0; j <= j1; ++j) acc += col[j];
    const int sj = sg[j1 + 1];
    if (sj >= 0) {
      if (j0 == 0 && head_open) partial[((size_t)gwarp * 2 + 0) * NV + V0 + v] = acc;
      else out_seg[(size_t)sj * NV + V0 + v] = acc;
    }
    acc = 0.0;
    j0 = j1 + 1;
  }
  if (tail_open && sg[32] >= 0) {
    for (int j = j0; j < 32; ++j) acc += col[j];
    partial[((size_t)gwarp * 2 + (j0 == 0 ? 0 : 1)) * NV + V0 + v] = acc;
  }
}
template <bool UNROLLED, int NVT, int V0>
__device__ __forceinline__ void flush_slice(const double (*st)[kStageLd], const int* sg, int own, int lane, int gwarp,
                                            double* __restrict__ out_seg, double* __restrict__ partial) {
  if (UNROLLED) segment_flush_slice<NVT, V0>(st, sg, own, lane, gwarp, out_seg, partial);
  else segment_flush_slice_serial<NVT, V0>(st, sg, own, lane, gwarp, out_seg, partial);
}

constexpr int kPipeThreads = 128;
constexpr int kPipeWarps = kPipeThreads / 32;
constexpr int kPipeStages = 3;        // corner c + 2 is in flight while corner c is computed
constexpr int kPipeHalf = 17;         // staging rows per flush (NV = 33 = 17 + 16)

// ---- E pass -----------------------------------------------------------------------------
// per-thread shared memory: ring of kPipeStages corner slots (96 B tag-corner record + 16 B
// observation, 112 B: the 16 B of padding keep the 128-bit reads conflict free) and one slot for
// the capture record (176 B).
constexpr int kESlot = 112, kECap = 176;
constexpr size_t kEPipeSmem = (size_t)kPipeThreads * (kPipeStages * kESlot + kECap) +
                              (size_t)kPipeWarps * kPipeHalf * kStageLd * sizeof(double) + kPipeWarps * 36 * sizeof(int) +
                              kPipeWarps * 4 * sizeof(double);

template <int MODEL, bool UNROLLED>
__global__ void __launch_bounds__(kPipeThreads, 2) accum_e_pipe_kernel(const AccumArgs a, int n_chunks) {
  extern __shared__ __align__(16) unsigned char pipe_sm[];
  unsigned char* ring = pipe_sm;                                               // [stage][thread][112]
  unsigned char* capsl = ring + (size_t)kPipeThreads * kPipeStages * kESlot;   // [thread][176]
  double(*stage)[kPipeHalf][kStageLd] = reinterpret_cast<double(*)[kPipeHalf][kStageLd]>(capsl + (size_t)kPipeThreads * kECap);
  int(*sseg)[36] = reinterpret_cast<int(*)[36]>(reinterpret_cast<unsigned char*>(stage) + sizeof(double) * kPipeWarps * kPipeHalf * kStageLd);
  double* cam_scratch = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sseg) + sizeof(int) * kPipeWarps * 36);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* myring = ring + (size_t)threadIdx.x * kESlot;
  unsigned char* mycap = capsl + (size_t)threadIdx.x * kECap;
  const size_t ring_stride = (size_t)kPipeThreads * kESlot;
  const double cm[3] = {a.cam[0], MODEL ? a.cam[1] : 0.0, MODEL ? a.cam[2] : 0.0};
  const size_t ps = a.plane;

  // issues the copies of corner i of the block (pos, tag) into ring stage `st`
  auto fetch_corner = [&](int st, int pos, int tag, int i) {
    unsigned char* dst = myring + st * ring_stride;
    const double2* tsrc = reinterpret_cast<const double2*>(a.tag_pre + (size_t)kTagPre * tag) + 6 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(dst + 16 * k, tsrc + k, 16);
    __pipeline_memcpy_async(dst + 96, a.obs + (size_t)(2 * i) * ps + pos, 8);
    __pipeline_memcpy_async(dst + 104, a.obs + (size_t)(2 * i + 1) * ps + pos, 8);
  };
  auto fetch_cap = [&](int cap) {
    const double2* src = reinterpret_cast<const double2*>(a.cap_pre + (size_t)kCapPre * cap);
#pragma unroll
    for (int k = 0; k < 11; ++k) __pipeline_memcpy_async(mycap + 16 * k, src + k, 16);
  };

  const int gstride = gridDim.x * kPipeThreads;
  int chunk = blockIdx.x;
  int pos = chunk * kPipeThreads + threadIdx.x;
  bool valid = chunk < n_chunks && pos < a.n_blk;
  int own = valid ? a.own_idx[pos] : -1;
  int oth = valid ? a.oth_idx[pos] : 0;
  // the next chunk's indices are always one whole block old when they are first used
  int npos = pos + gstride;
  bool nvalid = chunk + (int)gridDim.x < n_chunks && npos < a.n_blk;
  int nown = nvalid ? a.own_idx[npos] : -1;
  int noth = nvalid ? a.oth_idx[npos] : 0;
  // prologue: corners 0 and 1 of the first block (+ its capture record)
  if (valid) { fetch_cap(own); fetch_corner(0, pos, oth, 0); }
  __pipeline_commit();
  if (valid) fetch_corner(1, pos, oth, 1);
  __pipeline_commit();
  int slot = 0;  // ring stage of the corner computed next
  double KK = 0.0, Kr = 0.0, rr = 0.0;

  for (; chunk < n_chunks; chunk += gridDim.x) {
    const int n2pos = npos + gstride;
    const bool n2valid = chunk + 2 * (int)gridDim.x < n_chunks && n2pos < a.n_blk;
    const int n2own = n2valid ? a.own_idx[n2pos] : -1;
    const int n2oth = n2valid ? a.oth_idx[n2pos] : 0;
    const int gwarp = pos >> 5;
    // the poses of the lanes next to this warp (lane 0 only); parked in registers until the flush so
    // that no instruction of the corner loop waits for them
    int own_before = -2, own_after = -1;
    if (lane == 0) {
      const int w0 = gwarp << 5;
      if (w0 > 0 && w0 - 1 < a.n_blk) own_before = a.own_idx[w0 - 1];
      if (w0 + 32 < a.n_blk) own_after = a.own_idx[w0 + 32];
    }
    double AA[6] = {0, 0, 0, 0, 0, 0}, AO[9], OO[6] = {0, 0, 0, 0, 0, 0};
    double AX[9], OX[9];
    double Ar[3] = {0, 0, 0}, Or[3] = {0, 0, 0}, AK[3] = {0, 0, 0}, OK[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) { AO[i] = 0.0; AX[i] = 0.0; OX[i] = 0.0; }
    double cp[kCapPre];
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      // corner c has landed (the two groups behind it may still be in flight); its slot is read into
      // registers FIRST, then the copies of corner c + 2 are issued -- into the slot of corner c - 1 --
      // so that their issue slots hide the shared-memory latency of these reads
      __pipeline_wait_prior(1);
      double tp[12];
      double2 o2 = make_double2(0.0, 0.0);
      {
        if (i == 0) {
          const double2* c2 = reinterpret_cast<const double2*>(mycap);
#pragma unroll
          for (int k = 0; k < 11; ++k) {
            const double2 v = c2[k];
            cp[2 * k] = v.x;
            cp[2 * k + 1] = v.y;
          }
        }
        const double2* cur = reinterpret_cast<const double2*>(myring + slot * ring_stride);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const double2 v = cur[k];
          tp[2 * k] = v.x;
          tp[2 * k + 1] = v.y;
        }
        o2 = cur[6];
      }
      // corner c + 2: still this block for i < 2, else the next chunk's block (with its capture record:
      // the capture slot was emptied into registers at i == 0)
      int st2 = slot + 2;
      if (st2 >= kPipeStages) st2 -= kPipeStages;
      if (i < 2) {
        if (valid) fetch_corner(st2, pos, oth, i + 2);
      } else if (nvalid) {
        if (i == 2) fetch_cap(nown);
        fetch_corner(st2, npos, noth, i - 2);
      }
      __pipeline_commit();
      if (valid) {
        CornerJ j;
        if (MODEL == 0) {
          corner_jacobian(cp, tp, cm[0], o2.x, o2.y, j);
        } else {
          double Kl[2][2];
          corner_jacobian_m<1>(cp, tp, cm, o2.x, o2.y, j, Kl);
        }
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const double* A = j.A[row];
          const double* O = j.B[row];
          const double* X = j.C[row];
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
              AX[p * 3 + q] += A[p] * X[q];
              OX[p * 3 + q] += O[p] * X[q];
            }
            Ar[p] += A[p] * r;
            Or[p] += O[p] * r;
            AK[p] += A[p] * K;
            OK[p] += O[p] * K;
          }
          KK += K * K;
          Kr += K * r;
          rr += r * r;
        }
      }
      if (++slot == kPipeStages) slot = 0;
    }
    if (valid) {
      double* w = a.W + pos;
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
    // per-pose record, two staged flushes: values [0, 17) then [17, 33)
    {
      double(*st)[kStageLd] = stage[wid];
      sseg[wid][lane + 1] = own;
      if (lane == 0) { sseg[wid][0] = own_before; sseg[wid][33] = own_after; }
      st[tri6(0, 0)][lane] = AA[0]; st[tri6(0, 1)][lane] = AA[1]; st[tri6(0, 2)][lane] = AA[2];
      st[tri6(0, 3)][lane] = AO[0]; st[tri6(0, 4)][lane] = AO[1]; st[tri6(0, 5)][lane] = AO[2];
      st[tri6(1, 1)][lane] = AA[3]; st[tri6(1, 2)][lane] = AA[4];
      st[tri6(1, 3)][lane] = AO[3]; st[tri6(1, 4)][lane] = AO[4]; st[tri6(1, 5)][lane] = AO[5];
      st[tri6(2, 2)][lane] = AA[5];
      st[tri6(2, 3)][lane] = AO[6]; st[tri6(2, 4)][lane] = AO[7]; st[tri6(2, 5)][lane] = AO[8];
      st[tri6(3, 3)][lane] = OO[0]; st[tri6(3, 4)][lane] = OO[1];  // tri6(3, 4) == 16: the last row of the first slice
      __syncwarp();
      flush_slice<UNROLLED, kPipeHalf, 0>(st, sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
      __syncwarp();
      st[tri6(3, 5) - kPipeHalf][lane] = OO[2];
      st[tri6(4, 4) - kPipeHalf][lane] = OO[3]; st[tri6(4, 5) - kPipeHalf][lane] = OO[4];
      st[tri6(5, 5) - kPipeHalf][lane] = OO[5];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        st[21 + p - kPipeHalf][lane] = Ar[p];
        st[24 + p - kPipeHalf][lane] = Or[p];
        st[27 + p - kPipeHalf][lane] = AK[p];
        st[30 + p - kPipeHalf][lane] = OK[p];
      }
      __syncwarp();
      flush_slice<UNROLLED, NV - kPipeHalf, kPipeHalf>(st, sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
      __syncwarp();
    }
    pos = npos; valid = nvalid; own = nown; oth = noth;
    npos = n2pos; nvalid = n2valid; nown = n2own; noth = n2oth;
  }
  __pipeline_wait_prior(0);
  const double v[4] = {warp_sum(KK), warp_sum(Kr), warp_sum(rr), 0.0};
  cta_partial<4, false, kPipeWarps>(v, a.warp_cam, cam_scratch);
}

// ---- F pass -----------------------------------------------------------------------------
// The own pose (tag) is the same for ~every lane of a warp: its corner records come through L1,
// one corner ahead in registers (also across blocks).  The gathered operand is the capture's
// R | t (12 doubles, the compact copy written by prep_poses_kernel) plus the block's 8
// observations: one 160-byte slot per block, double-buffered across blocks; the indices it is
// addressed with are always one whole block old when they are first used.
constexpr int kFSlot = 160 + 16;  // +16 B padding: conflict-free 128-bit reads
constexpr size_t kFPipeSmem = (size_t)kPipeThreads * 2 * kFSlot + (size_t)kPipeWarps * kPipeHalf * kStageLd * sizeof(double) +
                              kPipeWarps * 36 * sizeof(int);

struct AccumFArgs {
  AccumArgs a;
  const double* cap_rt;  // [n_cap][12]  R (9) | t (3)
};

template <int MODEL, bool UNROLLED>
__global__ void __launch_bounds__(kPipeThreads, 3) accum_f_pipe_kernel(const AccumFArgs fa, int n_chunks) {
  const AccumArgs& a = fa.a;
  extern __shared__ __align__(16) unsigned char pipe_sm[];
  unsigned char* slots = pipe_sm;  // [2][thread][kFSlot]
  double(*stage)[kPipeHalf][kStageLd] = reinterpret_cast<double(*)[kPipeHalf][kStageLd]>(slots + (size_t)kPipeThreads * 2 * kFSlot);
  int(*sseg)[36] = reinterpret_cast<int(*)[36]>(reinterpret_cast<unsigned char*>(stage) + sizeof(double) * kPipeWarps * kPipeHalf * kStageLd);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* myslot = slots + (size_t)threadIdx.x * kFSlot;
  const size_t slot_stride = (size_t)kPipeThreads * kFSlot;
  const double cm[3] = {a.cam[0], MODEL ? a.cam[1] : 0.0, MODEL ? a.cam[2] : 0.0};
  const size_t ps = a.plane;

  auto fetch_block = [&](int st, int pos, int cap) {
    unsigned char* dst = myslot + st * slot_stride;
    const double2* src = reinterpret_cast<const double2*>(fa.cap_rt + (size_t)12 * cap);
#pragma unroll
    for (int k = 0; k < 6; ++k) __pipeline_memcpy_async(dst + 16 * k, src + k, 16);
#pragma unroll
    for (int k = 0; k < 8; ++k) __pipeline_memcpy_async(dst + 96 + 8 * k, a.obs + (size_t)k * ps + pos, 8);
  };

  const int gstride = gridDim.x * kPipeThreads;
  int chunk = blockIdx.x;
  int pos = chunk * kPipeThreads + threadIdx.x;
  bool valid = chunk < n_chunks && pos < a.n_blk;
  int own = valid ? a.own_idx[pos] : -1;
  int oth = valid ? a.oth_idx[pos] : 0;
  // the next chunk's indices are always one whole block old when they are first used
  int npos = pos + gstride;
  bool nvalid = chunk + (int)gridDim.x < n_chunks && npos < a.n_blk;
  int nown = nvalid ? a.own_idx[npos] : -1;
  int noth = nvalid ? a.oth_idx[npos] : 0;
  if (valid) fetch_block(0, pos, oth);
  __pipeline_commit();
  int slot = 0;
  // the own tag's corner records come through L1, one corner ahead in registers (also across blocks)
  double tn[12];
  {
    const double2* t0 = reinterpret_cast<const double2*>(a.tag_pre + (size_t)kTagPre * (valid ? own : 0));
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double2 v = __ldg(t0 + k);
      tn[2 * k] = v.x;
      tn[2 * k + 1] = v.y;
    }
  }
  for (; chunk < n_chunks; chunk += gridDim.x) {
    const int n2pos = npos + gstride;
    const bool n2valid = chunk + 2 * (int)gridDim.x < n_chunks && n2pos < a.n_blk;
    const int n2own = n2valid ? a.own_idx[n2pos] : -1;
    const int n2oth = n2valid ? a.oth_idx[n2pos] : 0;
    const int gwarp = pos >> 5;
    int own_before = -2, own_after = -1;
    if (lane == 0) {
      const int w0 = gwarp << 5;
      if (w0 > 0 && w0 - 1 < a.n_blk) own_before = a.own_idx[w0 - 1];
      if (w0 + 32 < a.n_blk) own_after = a.own_idx[w0 + 32];
    }
    double AA[6] = {0, 0, 0, 0, 0, 0}, AO[9], OO[6] = {0, 0, 0, 0, 0, 0};
    double Ar[3] = {0, 0, 0}, Or[3] = {0, 0, 0}, AK[3] = {0, 0, 0}, OK[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) AO[i] = 0.0;
    if (nvalid) fetch_block(slot ^ 1, npos, noth);
    __pipeline_commit();
    __pipeline_wait_prior(1);
    const double2* tsrc = reinterpret_cast<const double2*>(a.tag_pre + (size_t)kTagPre * (valid ? own : 0));
    const double2* tnext = reinterpret_cast<const double2*>(a.tag_pre + (size_t)kTagPre * (nvalid ? nown : 0));
    {
      const double2* cur = reinterpret_cast<const double2*>(myslot + slot * slot_stride);
      double cp[kCapPre];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double2 v = cur[k];
        cp[2 * k] = v.x;
        cp[2 * k + 1] = v.y;
      }
      {
        const double2 v4 = cur[4], v5 = cur[5];
        cp[8] = v4.x; cp[18] = v4.y; cp[19] = v5.x; cp[20] = v5.y;
      }
#pragma unroll
      for (int k = 9; k < 18; ++k) cp[k] = 0.0;  // M: the capture rotation columns are not formed in this pass
      cp[21] = 0.0; cp[22] = 0.0; cp[23] = 0.0;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        double tp[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) tp[k] = tn[k];
        {
          const double2* nx = i < 3 ? tsrc + 6 * (i + 1) : tnext;  // i == 3: corner 0 of the next block's tag
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const double2 v = __ldg(nx + k);
            tn[2 * k] = v.x;
            tn[2 * k + 1] = v.y;
          }
        }
        if (valid) {
          const double2 o2 = cur[6 + i];
          CornerJ j;
          if (MODEL == 0) {
            corner_jacobian(cp, tp, cm[0], o2.x, o2.y, j);
          } else {
            double Kl[2][2];
            corner_jacobian_m<1>(cp, tp, cm, o2.x, o2.y, j, Kl);
          }
#pragma unroll
          for (int row = 0; row < 2; ++row) {
            const double* A = j.A[row];
            const double* O = j.C[row];
            const double r = j.r[row], K = j.K[row];
            AA[0] += A[0] * A[0]; AA[1] += A[0] * A[1]; AA[2] += A[0] * A[2];
            AA[3] += A[1] * A[1]; AA[4] += A[1] * A[2]; AA[5] += A[2] * A[2];
            OO[0] += O[0] * O[0]; OO[1] += O[0] * O[1]; OO[2] += O[0] * O[2];
            OO[3] += O[1] * O[1]; OO[4] += O[1] * O[2]; OO[5] += O[2] * O[2];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
#pragma unroll
              for (int q = 0; q < 3; ++q) AO[p * 3 + q] += A[p] * O[q];
              Ar[p] += A[p] * r;
              Or[p] += O[p] * r;
              AK[p] += A[p] * K;
              OK[p] += O[p] * K;
            }
          }
        }
      }
    }
    {
      double(*st)[kStageLd] = stage[wid];
      sseg[wid][lane + 1] = own;
      if (lane == 0) { sseg[wid][0] = own_before; sseg[wid][33] = own_after; }
      st[tri6(0, 0)][lane] = AA[0]; st[tri6(0, 1)][lane] = AA[1]; st[tri6(0, 2)][lane] = AA[2];
      st[tri6(0, 3)][lane] = AO[0]; st[tri6(0, 4)][lane] = AO[1]; st[tri6(0, 5)][lane] = AO[2];
      st[tri6(1, 1)][lane] = AA[3]; st[tri6(1, 2)][lane] = AA[4];
      st[tri6(1, 3)][lane] = AO[3]; st[tri6(1, 4)][lane] = AO[4]; st[tri6(1, 5)][lane] = AO[5];
      st[tri6(2, 2)][lane] = AA[5];
      st[tri6(2, 3)][lane] = AO[6]; st[tri6(2, 4)][lane] = AO[7]; st[tri6(2, 5)][lane] = AO[8];
      st[tri6(3, 3)][lane] = OO[0]; st[tri6(3, 4)][lane] = OO[1];  // tri6(3, 4) == 16: the last row of the first slice
      __syncwarp();
      flush_slice<UNROLLED, kPipeHalf, 0>(st, sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
      __syncwarp();
      st[tri6(3, 5) - kPipeHalf][lane] = OO[2];
      st[tri6(4, 4) - kPipeHalf][lane] = OO[3]; st[tri6(4, 5) - kPipeHalf][lane] = OO[4];
      st[tri6(5, 5) - kPipeHalf][lane] = OO[5];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        st[21 + p - kPipeHalf][lane] = Ar[p];
        st[24 + p - kPipeHalf][lane] = Or[p];
        st[27 + p - kPipeHalf][lane] = AK[p];
        st[30 + p - kPipeHalf][lane] = OK[p];
      }
      __syncwarp();
      flush_slice<UNROLLED, NV - kPipeHalf, kPipeHalf>(st, sseg[wid], own, lane, gwarp, a.out_seg, a.partial);
      __syncwarp();
    }
    slot ^= 1;
    pos = npos; valid = nvalid; own = nown; oth = noth;
    npos = n2pos; nvalid = n2valid; nown = n2own; noth = n2oth;
  }
  __pipeline_wait_prior(0);
}

}  // namespace ars
