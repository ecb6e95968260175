// arslam.cu -- host side of libar_slam_b200.so: the C-ABI of
// include/ar_slam_b200.h, device memory, the LM control loop and NCCL.
//
// Replaces ArSlamSolver::optimize / resetProblem / the AddResidualBlock sites
// and localizeOne's per-capture solve (reference
// ar_slam/src/ar_slam_util.cpp:720-727, 829-836, 903-979, 1001-1025).
// There is no CPU fallback: without a CUDA device every entry point fails.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/ar_slam_b200.h"
#include "cholesky.cuh"
#include "kernels.cuh"
#include "accum_pipe.cuh"
#include "localize.cuh"
#include "pcg.cuh"
#include "schur.cuh"
#include "schur_local.cuh"

using namespace ars;

namespace {

thread_local std::string g_create_error;

double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- NCCL through dlopen: no link-time dependency, and inside a torch
// process the already loaded libnccl.so.2 is reused. -------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, const void* /* by value, see call */, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (lib) break;
    }
    for (const char* n : names) {
      if (lib) break;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(void**, int, const void*, int))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather) { err = "libnccl misses symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;
struct NcclId { char bytes[128]; };
constexpr int kNcclDouble = 8, kNcclInt64 = 4, kNcclSum = 0, kNcclMax = 2;  // ncclFloat64, ncclInt64, ncclSum, ncclMax

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  // contents are NOT preserved.  A buffer that has to grow a second time grows by half again: the
  // incremental schedules enlarge the problem by one capture per solve, and a cudaFree + cudaMalloc per buffer
  // per solve cost more than the solve itself (3.5 ms per optimize() on a 1600-block map).  The first
  // allocation is exact, so one-shot problems (config 3: 230 MB of W) do not over-allocate.
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    const size_t cap = p ? count + count / 2 + 16 : std::max<size_t>(count, 1);
    release();
    cudaError_t e = cudaMalloc(&p, cap * sizeof(T));
    if (e == cudaSuccess) n = cap;
    return e;
  }
  // exactly like grow_keep, and the elements [keep, count) are zero afterwards (also when no reallocation happened)
  cudaError_t grow_keep_zero(size_t count, size_t keep, cudaStream_t st) {
    cudaError_t e = grow_keep(count, keep, st);
    if (e == cudaSuccess && count > keep) e = cudaMemsetAsync(p + keep, 0, (count - keep) * sizeof(T), st);
    return e;
  }
  // capacity >= count with the first `keep` elements preserved (amortised growth)
  cudaError_t grow_keep(size_t count, size_t keep, cudaStream_t st) {
    if (count <= n && p) return cudaSuccess;
    T* q = nullptr;
    const size_t cap = count + count / 2 + 16;
    cudaError_t e = cudaMalloc(&q, cap * sizeof(T));
    if (e != cudaSuccess) return e;
    if (p && keep) e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (p) cudaFree(p);
    p = q;
    n = cap;
    return e;
  }
};

struct Profiler {
  bool on = false;
  struct Rec { int id; cudaEvent_t a, b; };
  std::vector<std::string> names;
  std::vector<double> bytes, total_ms;
  std::vector<long long> count;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  int id_of(const char* n, double b) {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) { bytes[i] = b; return (int)i; }
    names.push_back(n); bytes.push_back(b); total_ms.push_back(0); count.push_back(0);
    return (int)names.size() - 1;
  }
  cudaEvent_t ev() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void clear() {
    for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); }
    recs.clear(); names.clear(); bytes.clear(); total_ms.clear(); count.clear();
  }
  void resolve() {
    for (auto& r : recs) {
      float ms = 0; cudaEventElapsedTime(&ms, r.a, r.b);
      total_ms[r.id] += ms; count[r.id]++;
      pool.push_back(r.a); pool.push_back(r.b);
    }
    recs.clear();
  }
  ~Profiler() { for (auto e : pool) cudaEventDestroy(e); for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

}  // namespace

// ---- device-side construction of the two sorted SoA copies (set_problem) ------------
__global__ void make_sort_keys_kernel(int n, const int32_t* __restrict__ own, const int32_t* __restrict__ oth,
                                      unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  keys[b] = ((unsigned long long)(unsigned)own[b] << 32) | (unsigned)oth[b];
  vals[b] = b;
}
__global__ void shift_index_kernel(int n, int32_t* idx, int32_t by) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n) idx[b] += by;
}
__global__ void segment_sizes_kernel(int n, const int32_t* __restrict__ off, const int32_t* __restrict__ end, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)(end[i] - off[i]);
}
// ---- locality order of the capture-sorted copy: captures sorted by their smallest tag ---------------
// (captures that share it see overlapping tag sets; the elimination kernel pre-reduces the products of such
// neighbours inside a CTA, schur_local.cuh, and the tag records their blocks gather share cache lines)
__global__ void fill_i32_kernel(int n, int32_t* out, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
__global__ void min_other_kernel(int n_blk, const int32_t* __restrict__ own, const int32_t* __restrict__ oth, int32_t* __restrict__ min_oth) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_blk) atomicMin(min_oth + own[b], oth[b]);
}
__global__ void locality_keys_kernel(int n_own, const int32_t* __restrict__ min_oth, unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_own) keys[i] = ((unsigned long long)(unsigned)min_oth[i] << 32) | (unsigned)i;  // poses without blocks (INT_MAX) last
}
__global__ void locality_rank_kernel(int n_own, const unsigned long long* __restrict__ sorted, int32_t* __restrict__ by_rank, int32_t* __restrict__ rank) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_own) return;
  const int i = (int)(sorted[r] & 0xffffffffu);
  by_rank[r] = i;
  rank[i] = r;
}
__global__ void make_sort_keys_ranked_kernel(int n, const int32_t* __restrict__ own, const int32_t* __restrict__ oth, const int32_t* __restrict__ rank,
                                             unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  keys[b] = ((unsigned long long)(unsigned)rank[own[b]] << 32) | (unsigned)oth[b];
  vals[b] = b;
}
// segment bounds when the sorted order is by rank: start_by_rank[r] = first position whose rank is >= r (r = 0 .. n_own)
__global__ void rank_offsets_kernel(int n, int n_own, const unsigned long long* __restrict__ sorted_keys, int32_t* __restrict__ start_by_rank) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_own) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int)(sorted_keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
  }
  start_by_rank[r] = lo;
}
__global__ void rank_bounds_kernel(int n_own, const int32_t* __restrict__ start_by_rank, const int32_t* __restrict__ by_rank,
                                   int32_t* __restrict__ off, int32_t* __restrict__ end) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_own) return;
  const int i = by_rank[r];
  off[i] = start_by_rank[r];
  end[i] = start_by_rank[r + 1];
}
// sorted position p takes block perm[p]: indices and the 8 observation planes
__global__ void gather_sorted_kernel(int n, int plane, const int32_t* __restrict__ perm,
                                     const unsigned long long* __restrict__ keys, const double* __restrict__ rect8,
                                     int32_t* __restrict__ s_own, int32_t* __restrict__ s_oth, double* __restrict__ s_obs,
                                     const int32_t* __restrict__ by_rank /* key's high word is a rank: pose = by_rank[rank]; or null */) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= plane) return;
  if (p < n) {
    const int b = perm[p];
    const unsigned long long k = keys[p];
    s_own[p] = by_rank ? by_rank[(int32_t)(k >> 32)] : (int32_t)(k >> 32);
    s_oth[p] = (int32_t)(k & 0xffffffffu);
    const double2* src = reinterpret_cast<const double2*>(rect8 + 8 * (size_t)b);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 v = src[q];
      s_obs[(size_t)(2 * q) * plane + p] = v.x;
      s_obs[(size_t)(2 * q + 1) * plane + p] = v.y;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) s_obs[(size_t)q * plane + p] = 0.0;
  }
}
// seg_off[i] = first sorted position whose own index is >= i  (i = 0 .. n_own)
__global__ void segment_offsets_kernel(int n, int n_own, const int32_t* __restrict__ s_own, int32_t* __restrict__ off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_own) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_own[mid] < i) lo = mid + 1; else hi = mid;
  }
  off[i] = lo;
}

struct arslam_solver {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  arslam_options opt;
  std::string err;
  // problem
  int n_cap = 0, n_tag = 0, n_blk = 0, plane = 0, n_warp = 0;
  bool have_problem = false, have_params = false;
  bool keep_params = false;           // rebuild_views: keep the current parameter set (append_blocks with parameters on the device)
  int param_caps = 0, param_tags = 0; // poses that hold parameters
  std::vector<double> h_cam;  // host mirror of the camera (poses stay on the device between calls)
  // multi-GPU: captures are re-indexed to the rank's own range [cap_lo, cap_lo + n_cap) of the n_cap_global captures
  int cap_lo = 0, n_cap_global = 0;
  // original order (evaluate API)
  DevBuf<int32_t> o_cap, o_tag;
  DevBuf<double> o_obs;
  // sorted copies: side 0 capture-sorted, side 1 tag-sorted
  DevBuf<int32_t> s_own[2], s_oth[2], s_off[2];
  // side 0 (captures) is stored in locality order: its segments are not in index order and carry explicit ends
  DevBuf<int32_t> s_end0, rank0, by_rank0, min_oth0, start_by_rank0;
  bool locality0 = false;
  const int32_t* seg_end(int side) const { return (side == 0 && locality0) ? s_end0.p : s_off[side].p + 1; }
  const int32_t* order(int side) const { return (side == 0 && locality0) ? by_rank0.p : nullptr; }
  DevBuf<double> s_obs[2];
  long long problem_version = 0, pcg_version = -1;
  int pcg_side = -1, n_sm = 0;
  // parameters: two sets (current / candidate)
  DevBuf<double> cam[2], cap[2], tag[2], cap_pre[2], tag_pre[2], tag_cor[2], cap_rt[2];
  int cur = 0;
  // normal equations
  DevBuf<double> H[2], partial[2], W, Z, YB, seg_cam, seg_cross, warp_cam, warp_cand, warp_norm[2], warp_gmax[2];
  DevBuf<int> check3;  // store_blocks: first bad block, capture range
  DevBuf<int32_t> straddle_list[2];  // per sorted copy: elimination CTAs with a segment that leaves them (schur.cuh)
  DevBuf<int> straddle_count;
  int n_straddle[2] = {0, 0};
  DevBuf<double> sigE, sigF, d_cam, d_pose[2], uF, yF, sc, cam_minus, red, tri;  // tri: packed lower region of the dense system (multi-GPU sum)  // red: S | cam_minus | HF | sc head
  DevBuf<double> eval_out, small, colsum_part, linv;
  DenseCholesky::LookAhead lookahead;  // second stream + events of the dense factorisation
  DevBuf<long long> agree;   // multi-GPU: status word of comm_agree / capture ranges of the ranks
  DevBuf<double> f_blocks;   // multi-GPU: per F pose, number of blocks over all ranks
  DevBuf<unsigned> tickets;  // one ticket per in-kernel grid reduction (kernels.cuh), zero between launches
  // constant parameter blocks (arslam_set_constant); cleared by set_problem / append_blocks
  DevBuf<unsigned char> const_cap, const_tag;
  bool have_const = false;
  int cam_const = 0;
  DevBuf<double> Hx[2], partialx[2], warp_cam8;  // radial model: l1, l2 borders per pose side
  DevBuf<unsigned long long> sort_keys[2];
  DevBuf<int32_t> sort_vals[2];
  DevBuf<unsigned char> sort_tmp;
  // localisation batch buffers (kept between calls)
  DevBuf<int32_t> l_off, l_tag, l_seed, l_it, l_term;
  DevBuf<double> l_obs, l_tagpose, l_tagpre, l_pose, l_cost;
  DevBuf<int> l_invalid;
  DevBuf<int32_t> seed_idx;   // arslam_seed_*: staged indices and observations
  DevBuf<double> seed_rect;
  cudaStream_t loc_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t loc_done[3] = {nullptr, nullptr, nullptr}, loc_ready = nullptr;
  double* h_sc = nullptr;  // pinned
  long long ld = 0;
  int n_pad = 0;
  // pcg
  PcgWorkspace pcg;
  SchurLocalPlan schur_plan;   // locality-ordered elimination with in-CTA pre-reduction (schur_local.cuh)
  // comm
  void* comm = nullptr;
  int rank = 0, world = 1;
  // stats
  long long launches = 0;
  Profiler prof;
  // tuning switches (arslam_set_tuning), per handle
  int tune_accum_pipe = 2, tune_accum_flush = 0, tune_pcg_smem = 1, tune_pcg_pipelined = 1, tune_schur_bulk = 1, tune_schur_local = 0, tune_locality = 0, tune_chol_chain = 1;  // the locality-ordered elimination is parked: measured slower
  long long tune_loc_chunk = 0;  // captures per localisation chunk (0: default)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    err = buf;
    return code;
  }
};

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return s->fail(ARSLAM_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

// launch wrapper: counts, and brackets with events when profiling
#define LAUNCH(name, bytes, ...)                                                              \
  do {                                                                                        \
    if (s->prof.on) {                                                                         \
      Profiler::Rec r__{s->prof.id_of(name, (double)(bytes)), s->prof.ev(), s->prof.ev()};    \
      cudaEventRecord(r__.a, s->stream);                                                      \
      __VA_ARGS__;                                                                            \
      cudaEventRecord(r__.b, s->stream);                                                      \
      s->prof.recs.push_back(r__);                                                            \
    } else {                                                                                  \
      __VA_ARGS__;                                                                            \
    }                                                                                         \
    ++s->launches;                                                                            \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// main launch + the launch for the products whose partner sits in another CTA (schur.cuh)
template <typename Target, int NK>
void launch_schur(arslam_solver* s, const SchurArgs& a, const Target& t, const int32_t* e_idx, double bytes, bool bulk = false) {
  static_assert(kSchurThreads == kSchurCtaBlocks, "the straddle list is built for the elimination kernel's CTA size");
  const int grid = cdiv(s->n_blk, kSchurThreads);
  if (bulk) {
    LAUNCH("schur_eliminate", bytes, schur_eliminate_kernel<Target, NK, false, true><<<grid, kSchurThreads, kSchurSmem, s->stream>>>(a, t, s->n_blk, e_idx));
    if (a.n_straddle > 0)
      LAUNCH("schur_straddle", 0.0, schur_eliminate_kernel<Target, NK, true, true><<<a.n_straddle, kSchurThreads, kSchurSmem, s->stream>>>(a, t, s->n_blk, e_idx));
  } else {
    LAUNCH("schur_eliminate", bytes, schur_eliminate_kernel<Target, NK, false><<<grid, kSchurThreads, kSchurSmem, s->stream>>>(a, t, s->n_blk, e_idx));
    if (a.n_straddle > 0)
      LAUNCH("schur_straddle", 0.0, schur_eliminate_kernel<Target, NK, true><<<a.n_straddle, kSchurThreads, kSchurSmem, s->stream>>>(a, t, s->n_blk, e_idx));
  }
}
template <typename Target, int NK>
cudaError_t schur_kernel_attributes() {
  cudaError_t e = cudaFuncSetAttribute(schur_eliminate_kernel<Target, NK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSchurSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_eliminate_kernel<Target, NK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSchurSmem);
  // three CTAs of 75 KB per SM need the full shared-memory carve-out
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_eliminate_kernel<Target, NK, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return e;
}
static cudaError_t schur_local_attributes() {
  cudaError_t e = cudaFuncSetAttribute(schur_local_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_local_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSlSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_local_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_local_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return e;
}
template <int NK>
static cudaError_t schur_bulk_attributes() {
  cudaError_t e = cudaFuncSetAttribute(schur_eliminate_kernel<SparseTarget, NK, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSchurSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_eliminate_kernel<SparseTarget, NK, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSchurSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(schur_eliminate_kernel<SparseTarget, NK, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return e;
}



static cudaError_t pipe_kernel_attributes() {
  cudaError_t e = cudaSuccess;
#define ARS_ATTR(K_, BYTES_)                                                                                  \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES_)); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(K_, cudaFuncAttributePreferredSharedMemoryCarveout, 100)
  ARS_ATTR((accum_e_pipe_kernel<0, false>), kEPipeSmem); ARS_ATTR((accum_e_pipe_kernel<0, true>), kEPipeSmem);
  ARS_ATTR((accum_e_pipe_kernel<1, false>), kEPipeSmem); ARS_ATTR((accum_e_pipe_kernel<1, true>), kEPipeSmem);
  ARS_ATTR((accum_f_pipe_kernel<0, false>), kFPipeSmem); ARS_ATTR((accum_f_pipe_kernel<0, true>), kFPipeSmem);
  ARS_ATTR((accum_f_pipe_kernel<1, false>), kFPipeSmem); ARS_ATTR((accum_f_pipe_kernel<1, true>), kFPipeSmem);
#undef ARS_ATTR
  return e;
}

extern "C" {

namespace { int launch_colsum(arslam_solver* s, int n, int m, const double* in, double* out); }

int arslam_abi_version(void) { return ARSLAM_ABI_VERSION; }

void arslam_default_options(arslam_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_num_iterations = 50;  // ar_slam_util.cpp:1004
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->elimination = ARSLAM_ELIM_AUTO;
  o->linear_solver = ARSLAM_LINSOLVE_AUTO;  // ar_slam_util.cpp:1011 DENSE_SCHUR where it is dense
  o->pcg_max_iterations = 500;
  o->num_intrinsics = 1;
  o->verbose = 0;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->pcg_tolerance = 0.1;  // inexact-Newton forcing term, the value of Ceres' Solver::Options::eta
  o->pcg_q_tolerance = 0.0;
  o->tag_size = 0.0635;  // ar_slam_util.hpp:319
  o->dense_max_dim = 16384;
}

const char* arslam_last_error(const arslam_solver* s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int arslam_create(int device, const arslam_options* opt, arslam_solver** out) {
  if (!out) return ARSLAM_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    g_create_error = std::string("no CUDA device (") + cudaGetErrorString(e) +
                     "); this solver has no CPU fallback";
    cudaGetLastError();
    return ARSLAM_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) { g_create_error = "device ordinal out of range"; return ARSLAM_ERR_INVALID; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int n_sm = prop.multiProcessorCount;
  if (prop.major < 10) {
    g_create_error = "device is not sm_100 class; kernels are built for sm_100a only";
    return ARSLAM_ERR_NO_DEVICE;
  }
  arslam_solver* s = new arslam_solver();
  s->device = device;
  s->n_sm = n_sm;
  if (opt) s->opt = *opt; else arslam_default_options(&s->opt);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(&s->h_sc, 256 * sizeof(double)) != cudaSuccess || DenseCholesky::init() != cudaSuccess ||
      schur_kernel_attributes<SparseTarget, 1>() != cudaSuccess || schur_kernel_attributes<DenseTarget, 1>() != cudaSuccess ||
      schur_kernel_attributes<DenseTarget, 3>() != cudaSuccess ||
      pcg_init() != cudaSuccess || schur_bulk_attributes<1>() != cudaSuccess || schur_bulk_attributes<3>() != cudaSuccess ||
      schur_kernel_attributes<SparseTarget, 3>() != cudaSuccess || schur_local_attributes() != cudaSuccess ||
      pipe_kernel_attributes() != cudaSuccess) {
    g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
    delete s;
    return ARSLAM_ERR_CUDA;
  }
  s->stream = s->own_stream;
  for (auto& ev : s->ev) cudaEventCreate(&ev);
  *out = s;
  return ARSLAM_OK;
}

void arslam_destroy(arslam_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  for (auto& ev : s->ev) if (ev) cudaEventDestroy(ev);
  for (auto& st : s->loc_stream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (auto& ev : s->loc_done) if (ev) cudaEventDestroy(ev);
  if (s->loc_ready) cudaEventDestroy(s->loc_ready);
  if (s->h_sc) cudaFreeHost(s->h_sc);
  if (s->own_stream) cudaStreamDestroy(s->own_stream);
  delete s;
}

int arslam_set_options(arslam_solver* s, const arslam_options* opt) {
  if (!s || !opt) return ARSLAM_ERR_INVALID;
  if (opt->num_intrinsics != 1 && opt->num_intrinsics != 3)
    return s->fail(ARSLAM_ERR_INVALID, "num_intrinsics=%d: 1 (focal only, the reference's live model) or 3 (focal + radial l1, l2)", opt->num_intrinsics);
  s->opt = *opt;
  return ARSLAM_OK;
}

int arslam_set_stream(arslam_solver* s, void* cuda_stream) {
  if (!s) return ARSLAM_ERR_INVALID;
  cudaSetDevice(s->device);
  cudaStreamSynchronize(s->stream);
  s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->own_stream;
  return ARSLAM_OK;
}

int arslam_set_profiling(arslam_solver* s, int on) {
  if (!s) return ARSLAM_ERR_INVALID;
  s->prof.on = on != 0;
  return ARSLAM_OK;
}

int arslam_set_tuning(arslam_solver* s, const char* key, int64_t value) {
  if (!s || !key) return ARSLAM_ERR_INVALID;
  const std::string k(key);
  if (k == "accum_pipe") s->tune_accum_pipe = (int)value & 3;  // bit 0: E pass, bit 1: F pass
  else if (k == "accum_flush") s->tune_accum_flush = value != 0;
  else if (k == "pcg_smem") s->tune_pcg_smem = value != 0;
  else if (k == "schur_bulk") s->tune_schur_bulk = value != 0;
  else if (k == "locality") s->tune_locality = value != 0;  // takes effect at the next arslam_set_problem / append_blocks
  else if (k == "schur_local") { s->tune_schur_local = value != 0; s->pcg.valid = false; }  // the plan is (re)built with the symbolic phase
  else if (k == "loc_chunk") s->tune_loc_chunk = value;
  else if (k == "pcg_pipelined") s->tune_pcg_pipelined = value != 0;
  else if (k == "chol_chain") s->tune_chol_chain = value != 0;
  else if (k == "chol_big") s->lookahead.variant = value != 0;
  else if (k == "chol_nb") s->lookahead.nb = (int)std::max<int64_t>(128, std::min<int64_t>(1024, value / 128 * 128));
  else if (k == "chol_free_sms") { s->lookahead.free_sms = (int)std::max<int64_t>(0, std::min<int64_t>(64, value)); s->lookahead.rest_ctas = std::max(1, s->lookahead.sms - s->lookahead.free_sms); }
  else return s->fail(ARSLAM_ERR_INVALID, "set_tuning: unknown key '%s'", key);
  return ARSLAM_OK;
}

int arslam_kernel_times(arslam_solver* s, arslam_kernel_time* out, int32_t cap) {
  if (!s) return ARSLAM_ERR_INVALID;
  int n = 0;
  for (size_t i = 0; i < s->prof.names.size() && n < cap; ++i, ++n) {
    std::memset(&out[n], 0, sizeof(out[n]));
    std::snprintf(out[n].name, sizeof(out[n].name), "%s", s->prof.names[i].c_str());
    out[n].total_ms = s->prof.total_ms[i];
    out[n].launches = s->prof.count[i];
    out[n].algorithmic_bytes = s->prof.bytes[i];
  }
  return n;
}

namespace {

// validates blocks [first, first + n_new) and moves them into the original-order arrays
// range check of freshly uploaded indices + the capture range they span: out[0] = first bad block (or INT_MAX), out[1] / out[2] = min / max capture
__global__ void validate_blocks_kernel(int n, const int32_t* __restrict__ cap, const int32_t* __restrict__ tag, int n_cap, int n_tag, int* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0x7fffffff, lo = 0x7fffffff, hi = -1;
  if (b < n) {
    const int c = cap[b], t = tag[b];
    if (c < 0 || c >= n_cap || t < 0 || t >= n_tag) bad = b;
    lo = hi = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bad = min(bad, __shfl_xor_sync(0xffffffffu, bad, o));
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (bad != 0x7fffffff) atomicMin(out, bad);
    atomicMin(out + 1, lo);
    atomicMax(out + 2, hi);
  }
}

// Uploads blocks [first, first + n_new) and validates them ON THE DEVICE (a serial host loop over 800 k blocks cost
// 0.6 ms per call at config 3).  The stream is drained before an error is returned, so no copy is left reading the
// caller's (possibly pinned, possibly about to be freed) arrays; the blocks before `first` are untouched.
int store_blocks(arslam_solver* s, int64_t n_cap, int64_t n_tag, int64_t first, int64_t n_new, const int32_t* cap_idx,
                 const int32_t* tag_idx, const double* rect8, int32_t* lo_out, int32_t* hi_out) {
  const bool on_host = n_new <= 4096;  // the incremental schedules append a handful of blocks: cheaper than a kernel + read-back
  if (on_host) {
    int32_t lo = cap_idx[0], hi = cap_idx[0];
    for (int64_t b = 0; b < n_new; ++b) {
      if (cap_idx[b] < 0 || cap_idx[b] >= n_cap || tag_idx[b] < 0 || tag_idx[b] >= n_tag)
        return s->fail(ARSLAM_ERR_INVALID, "block %lld has an index out of range", (long long)(first + b));
      lo = std::min(lo, cap_idx[b]);
      hi = std::max(hi, cap_idx[b]);
    }
    *lo_out = lo;
    *hi_out = hi;
  }
  CU(s->o_obs.grow_keep((size_t)(first + n_new) * 8, (size_t)first * 8, s->stream));
  CU(cudaMemcpyAsync(s->o_obs.p + 8 * first, rect8, sizeof(double) * 8 * n_new, cudaMemcpyHostToDevice, s->stream));
  CU(s->o_cap.grow_keep((size_t)(first + n_new), (size_t)first, s->stream));
  CU(s->o_tag.grow_keep((size_t)(first + n_new), (size_t)first, s->stream));
  CU(cudaMemcpyAsync(s->o_cap.p + first, cap_idx, sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->o_tag.p + first, tag_idx, sizeof(int32_t) * n_new, cudaMemcpyHostToDevice, s->stream));
  if (on_host) return ARSLAM_OK;
  CU(s->check3.ensure(4));
  CU(cudaMemsetAsync(s->check3.p, 0x7f, 2 * sizeof(int), s->stream));
  CU(cudaMemsetAsync(s->check3.p + 2, 0xff, sizeof(int), s->stream));
  validate_blocks_kernel<<<cdiv(n_new, 256), 256, 0, s->stream>>>((int)n_new, s->o_cap.p + first, s->o_tag.p + first, (int)n_cap, (int)n_tag, s->check3.p);
  int* h = reinterpret_cast<int*>(s->h_sc);  // pinned
  CU(cudaMemcpyAsync(h, s->check3.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (h[0] < (int)n_new) return s->fail(ARSLAM_ERR_INVALID, "block %lld has an index out of range", (long long)(first + h[0]));
  *lo_out = h[1];
  *hi_out = h[2];
  return ARSLAM_OK;
}

// both sorted copies are built on the GPU from the original-order arrays: stable radix sort of
// (own << 32 | other) keys, then one gather kernel writes the index arrays and the 8 planes
int rebuild_views(arslam_solver* s) {
  s->plane = (s->n_blk + 31) / 32 * 32;
  s->n_warp = s->plane / 32;
  const int nb = s->n_blk, plane = s->plane;
  CU(s->sort_keys[0].ensure(nb)); CU(s->sort_keys[1].ensure(nb)); CU(s->sort_vals[0].ensure(nb)); CU(s->sort_vals[1].ensure(nb));
  for (int side = 0; side < 2; ++side) {
    const int32_t* d_own = side == 0 ? s->o_cap.p : s->o_tag.p;
    const int32_t* d_oth = side == 0 ? s->o_tag.p : s->o_cap.p;
    const int n_own = side == 0 ? s->n_cap : s->n_tag;
    CU(s->s_own[side].ensure(plane)); CU(s->s_oth[side].ensure(plane)); CU(s->s_off[side].ensure(n_own + 1));
    CU(s->s_obs[side].ensure((size_t)8 * plane));
    const bool ranked = side == 0 && s->tune_locality;
    if (side == 0) s->locality0 = ranked;
    size_t tmp_bytes = 0;
    if (ranked) {
      // captures in locality order: sort them by (smallest tag, index), then sort the blocks by (rank, tag)
      CU(s->min_oth0.ensure(n_own)); CU(s->rank0.ensure(n_own)); CU(s->by_rank0.ensure(n_own)); CU(s->s_end0.ensure(n_own));
      CU(s->start_by_rank0.ensure(n_own + 1));
      CU(s->sort_keys[0].ensure(std::max(nb, n_own))); CU(s->sort_keys[1].ensure(std::max(nb, n_own)));
      fill_i32_kernel<<<cdiv(n_own, 256), 256, 0, s->stream>>>(n_own, s->min_oth0.p, 0x7fffffff);
      min_other_kernel<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, d_own, d_oth, s->min_oth0.p);
      locality_keys_kernel<<<cdiv(n_own, 256), 256, 0, s->stream>>>(n_own, s->min_oth0.p, s->sort_keys[0].p);
      CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, s->sort_keys[0].p, s->sort_keys[1].p, n_own, 0, 64, s->stream));
      CU(s->sort_tmp.ensure(tmp_bytes));
      CU(cub::DeviceRadixSort::SortKeys(s->sort_tmp.p, tmp_bytes, s->sort_keys[0].p, s->sort_keys[1].p, n_own, 0, 64, s->stream));
      locality_rank_kernel<<<cdiv(n_own, 256), 256, 0, s->stream>>>(n_own, s->sort_keys[1].p, s->by_rank0.p, s->rank0.p);
      make_sort_keys_ranked_kernel<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, d_own, d_oth, s->rank0.p, s->sort_keys[0].p, s->sort_vals[0].p);
    } else {
      make_sort_keys_kernel<<<cdiv(nb, 256), 256, 0, s->stream>>>(nb, d_own, d_oth, s->sort_keys[0].p, s->sort_vals[0].p);
    }
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s->sort_keys[0].p, s->sort_keys[1].p, s->sort_vals[0].p,
                                       s->sort_vals[1].p, nb, 0, 64, s->stream));
    CU(s->sort_tmp.ensure(tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(s->sort_tmp.p, tmp_bytes, s->sort_keys[0].p, s->sort_keys[1].p, s->sort_vals[0].p,
                                       s->sort_vals[1].p, nb, 0, 64, s->stream));
    gather_sorted_kernel<<<cdiv(plane, 256), 256, 0, s->stream>>>(nb, plane, s->sort_vals[1].p, s->sort_keys[1].p, s->o_obs.p,
                                                                 s->s_own[side].p, s->s_oth[side].p, s->s_obs[side].p,
                                                                 ranked ? s->by_rank0.p : nullptr);
    if (ranked) {
      rank_offsets_kernel<<<cdiv(n_own + 1, 256), 256, 0, s->stream>>>(nb, n_own, s->sort_keys[1].p, s->start_by_rank0.p);
      rank_bounds_kernel<<<cdiv(n_own, 256), 256, 0, s->stream>>>(n_own, s->start_by_rank0.p, s->by_rank0.p, s->s_off[side].p, s->s_end0.p);
    } else {
      segment_offsets_kernel<<<cdiv(n_own + 1, 256), 256, 0, s->stream>>>(nb, n_own, s->s_own[side].p, s->s_off[side].p);
    }
    CU(s->H[side].ensure((size_t)n_own * NV));
    CU(s->partial[side].ensure((size_t)s->n_warp * 2 * NV));
    CU(s->d_pose[side].ensure((size_t)6 * n_own));
    CU(s->warp_norm[side].ensure((size_t)12 * cdiv(n_own, 128) + 12));
    CU(s->warp_gmax[side].ensure((size_t)4 * cdiv(n_own, 128) + 4));
  }
  {
    const int nc = cdiv(nb, kSchurCtaBlocks);
    CU(s->straddle_count.ensure(2));
    CU(cudaMemsetAsync(s->straddle_count.p, 0, 2 * sizeof(int), s->stream));
    for (int side = 0; side < 2; ++side) {
      CU(s->straddle_list[side].ensure(nc));
      straddle_ctas_kernel<<<cdiv(nc, 256), 256, 0, s->stream>>>(nb, s->s_own[side].p, s->s_off[side].p, s->seg_end(side), s->straddle_list[side].p,
                                                                 s->straddle_count.p + side);
    }
    CU(cudaMemcpyAsync(s->h_sc, s->straddle_count.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  }
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  s->n_straddle[0] = reinterpret_cast<const int*>(s->h_sc)[0];
  s->n_straddle[1] = reinterpret_cast<const int*>(s->h_sc)[1];
  for (int k = 0; k < 2; ++k) {
    CU(s->cam[k].ensure(4));
    // the current parameter set survives a growing problem (arslam_append_blocks): poses that did not exist
    // before are zero until arslam_set_poses / arslam_seed_* / arslam_set_params fills them
    const bool keep = s->keep_params && k == s->cur;
    CU(s->cap[k].grow_keep_zero((size_t)6 * s->n_cap, keep ? (size_t)6 * s->param_caps : 0, s->stream));
    CU(s->tag[k].grow_keep_zero((size_t)6 * s->n_tag, keep ? (size_t)6 * s->param_tags : 0, s->stream));
    CU(s->cap_pre[k].ensure((size_t)kCapPre * s->n_cap)); CU(s->tag_pre[k].ensure((size_t)kTagPre * s->n_tag));
    CU(s->tag_cor[k].ensure((size_t)12 * s->n_tag)); CU(s->cap_rt[k].ensure((size_t)12 * s->n_cap));
  }
  CU(s->W.ensure((size_t)36 * plane));
  CU(s->warp_cam.ensure((size_t)4 * s->n_warp)); CU(s->warp_cand.ensure((size_t)2 * s->n_warp + 8));
  if (!s->tickets.p) { CU(s->tickets.ensure(16)); CU(cudaMemsetAsync(s->tickets.p, 0, 16 * sizeof(unsigned), s->stream)); }
  CU(s->d_cam.ensure(4)); CU(s->sc.ensure(kNumScalars)); CU(s->cam_minus.ensure(12)); CU(s->colsum_part.ensure(12 * kColsumChunks)); CU(s->linv.ensure(CB * CB));
  s->have_problem = true;
  s->have_const = false;
  s->cam_const = 0;
  ++s->problem_version;
  return ARSLAM_OK;
}

}  // namespace

namespace { int comm_agree(arslam_solver* s, int local, const char* what); }

int arslam_set_problem(arslam_solver* s, int64_t n_cap, int64_t n_tag, int64_t n_blk, const int32_t* cap_idx,
                       const int32_t* tag_idx, const double* rect8) {
  if (!s) return ARSLAM_ERR_INVALID;
  int rc = ARSLAM_OK;
  if (n_cap <= 0 || n_tag <= 0 || n_blk <= 0 || !cap_idx || !tag_idx || !rect8)
    rc = s->fail(ARSLAM_ERR_INVALID, "set_problem: empty problem or null pointer%s", s->world > 1 ? " (every rank needs at least one block)" : "");
  else if (n_blk > (1LL << 28) || n_cap > (1LL << 28) || n_tag > (1LL << 28))
    rc = s->fail(ARSLAM_ERR_INVALID, "set_problem: problem too large for 32-bit block indices");
  CU(cudaSetDevice(s->device));
  s->have_problem = false;
  s->have_params = false;
  int32_t lo = 0, hi = 0;
  if (!rc) rc = store_blocks(s, n_cap, n_tag, 0, n_blk, cap_idx, tag_idx, rect8, &lo, &hi);
  // under arslam_comm_init every rank learns here whether ALL ranks declared a valid shard
  rc = comm_agree(s, rc, "set_problem");
  if (rc) return rc;
  s->cap_lo = 0;
  s->n_cap_global = (int)n_cap;
  if (s->world > 1) {
    // a rank only ever touches the captures of its own blocks: work on that index range alone.  The
    // ranges of the ranks must not overlap: a capture whose blocks sit on two ranks would be eliminated
    // twice from partial blocks (silently wrong Schur complement).
    CU(s->agree.ensure((size_t)2 * s->world));
    long long mine[2] = {lo, hi};
    std::vector<long long> all((size_t)2 * s->world);
    CU(cudaMemcpyAsync(s->agree.p + 2 * s->rank, mine, sizeof(mine), cudaMemcpyHostToDevice, s->stream));
    const int nrc = g_nccl.AllGather(s->agree.p + 2 * s->rank, s->agree.p, 2, kNcclInt64, s->comm, s->stream);
    if (nrc) return s->fail(ARSLAM_ERR_NCCL, "ncclAllGather(ranges): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?");
    CU(cudaMemcpyAsync(all.data(), s->agree.p, sizeof(long long) * 2 * s->world, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int a = 0; a < s->world; ++a)
      for (int b = a + 1; b < s->world; ++b)
        if (all[2 * a] <= all[2 * b + 1] && all[2 * b] <= all[2 * a + 1])  // identical on every rank: all fail together
          return s->fail(ARSLAM_ERR_INVALID, "set_problem: ranks %d and %d declare overlapping capture ranges [%lld, %lld] and [%lld, %lld]; "
                         "all blocks of a capture must be on one rank", a, b, all[2 * a], all[2 * a + 1], all[2 * b], all[2 * b + 1]);
    s->cap_lo = lo;
    n_cap = (int64_t)hi - lo + 1;
    if (s->cap_lo) shift_index_kernel<<<cdiv(n_blk, 256), 256, 0, s->stream>>>((int)n_blk, s->o_cap.p, -s->cap_lo);
  }
  s->n_cap = (int)n_cap; s->n_tag = (int)n_tag; s->n_blk = (int)n_blk;
  s->keep_params = false;
  return comm_agree(s, rebuild_views(s), "set_problem");
}

int arslam_append_blocks(arslam_solver* s, int64_t n_cap, int64_t n_tag, int64_t n_new, const int32_t* cap_idx,
                         const int32_t* tag_idx, const double* rect8) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (!s->have_problem) return arslam_set_problem(s, n_cap, n_tag, n_new, cap_idx, tag_idx, rect8);
  if (s->world > 1) return s->fail(ARSLAM_ERR_UNSUPPORTED, "append_blocks under arslam_comm_init: declare the rank's blocks with set_problem");
  if (n_new <= 0 || !cap_idx || !tag_idx || !rect8) return s->fail(ARSLAM_ERR_INVALID, "append_blocks: no blocks or null pointer");
  if (n_cap < s->n_cap || n_tag < s->n_tag) return s->fail(ARSLAM_ERR_INVALID, "append_blocks: pose counts cannot shrink");
  if (s->n_blk + n_new > (1LL << 28) || n_cap > (1LL << 28) || n_tag > (1LL << 28))
    return s->fail(ARSLAM_ERR_INVALID, "append_blocks: problem too large for 32-bit block indices");
  CU(cudaSetDevice(s->device));
  const bool had_params = s->have_params;
  s->have_problem = false;
  s->have_params = false;
  int32_t lo = 0, hi = 0;
  const int rc = store_blocks(s, n_cap, n_tag, s->n_blk, n_new, cap_idx, tag_idx, rect8, &lo, &hi);
  if (rc) {  // the rejected blocks sit behind the n_blk the handle knows: the earlier blocks and parameters are intact
    s->have_problem = true;
    s->have_params = had_params;
    return rc;
  }
  s->n_cap = (int)n_cap; s->n_tag = (int)n_tag; s->n_blk += (int)n_new;
  s->n_cap_global = (int)n_cap;
  // the parameters on the device stay (the incremental schedules continue from the previous solve's result);
  // poses that are new are zero until they are set or seeded
  s->keep_params = had_params;
  const int rrc = rebuild_views(s);
  s->keep_params = false;
  if (rrc) return rrc;
  s->have_params = had_params;
  s->param_caps = s->n_cap; s->param_tags = s->n_tag;
  return ARSLAM_OK;
}

int arslam_set_constant(arslam_solver* s, int camera_constant, const uint8_t* cap_constant, const uint8_t* tag_constant) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (!s->have_problem) return s->fail(ARSLAM_ERR_INVALID, "set_constant before set_problem");
  CU(cudaSetDevice(s->device));
  CU(s->const_cap.ensure((size_t)s->n_cap)); CU(s->const_tag.ensure((size_t)s->n_tag));
  if (cap_constant) CU(cudaMemcpyAsync(s->const_cap.p, cap_constant + s->cap_lo, (size_t)s->n_cap, cudaMemcpyHostToDevice, s->stream));
  else CU(cudaMemsetAsync(s->const_cap.p, 0, (size_t)s->n_cap, s->stream));
  if (tag_constant) CU(cudaMemcpyAsync(s->const_tag.p, tag_constant, (size_t)s->n_tag, cudaMemcpyHostToDevice, s->stream));
  else CU(cudaMemsetAsync(s->const_tag.p, 0, (size_t)s->n_tag, s->stream));
  CU(cudaStreamSynchronize(s->stream));  // the caller may reuse its arrays
  s->cam_const = camera_constant != 0;
  s->have_const = cap_constant || tag_constant;
  return ARSLAM_OK;
}

int arslam_set_params(arslam_solver* s, const double* camera3, const double* cap_pose6, const double* tag_pose6) {
  if (!s || !camera3 || !cap_pose6 || !tag_pose6) return ARSLAM_ERR_INVALID;
  if (!s->have_problem) return s->fail(ARSLAM_ERR_INVALID, "set_params before set_problem");
  CU(cudaSetDevice(s->device));
  s->h_cam.assign(camera3, camera3 + 3);
  // straight from the caller's arrays into parameter set 0 (DMA when they are pinned); in a
  // multi-GPU solve only the rank's own capture range is moved
  double* cam4 = s->h_sc + 200;  // pinned staging
  cam4[0] = camera3[0]; cam4[1] = camera3[1]; cam4[2] = camera3[2]; cam4[3] = 0.0;
  CU(cudaMemcpyAsync(s->cam[0].p, cam4, 4 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->cap[0].p, cap_pose6 + (size_t)6 * s->cap_lo, sizeof(double) * 6 * s->n_cap, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->tag[0].p, tag_pose6, sizeof(double) * 6 * s->n_tag, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));  // the caller may reuse its arrays
  s->cur = 0;
  s->have_params = true;
  s->param_caps = s->n_cap; s->param_tags = s->n_tag;
  return ARSLAM_OK;
}

// ---- parameters that stay on the device between solves (the incremental schedules) -----------------
namespace {
__global__ void seed_captures_kernel(int n, const int32_t* __restrict__ cap_idx, const int32_t* __restrict__ tag_idx,
                                     const double* __restrict__ rect8, const double* __restrict__ cam, const double* __restrict__ tag_pose,
                                     double tag_size, double* __restrict__ cap_pose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rect[8], tp[6], out[6];
  for (int k = 0; k < 8; ++k) rect[k] = rect8[8 * (size_t)i + k];
  for (int k = 0; k < 6; ++k) tp[k] = tag_pose[6 * (size_t)tag_idx[i] + k];
  seed_capture_pose(rect, cam[0], tp, tag_size, out);  // initCapturePose, ar_slam_util.cpp:91-108
  for (int k = 0; k < 6; ++k) cap_pose[6 * (size_t)cap_idx[i] + k] = out[k];
}
__global__ void seed_tags_kernel(int n, const int32_t* __restrict__ tag_idx, const int32_t* __restrict__ cap_idx,
                                 const double* __restrict__ rect8, const double* __restrict__ cam, const double* __restrict__ cap_pose,
                                 double tag_size, double* __restrict__ tag_pose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rect[8], cp[6], out[6];
  for (int k = 0; k < 8; ++k) rect[k] = rect8[8 * (size_t)i + k];
  for (int k = 0; k < 6; ++k) cp[k] = cap_pose[6 * (size_t)cap_idx[i] + k];
  seed_tag_pose(rect, cam[0], cp, tag_size, out);      // initArPose, ar_slam_util.cpp:111-128
  for (int k = 0; k < 6; ++k) tag_pose[6 * (size_t)tag_idx[i] + k] = out[k];
}
}  // namespace

static int pose_range_ok(arslam_solver* s, int which, int64_t first, int64_t count, const char* what) {
  if (!s->have_problem || !s->have_params) return s->fail(ARSLAM_ERR_INVALID, "%s needs set_problem and parameters on the device", what);
  if (s->world > 1) return s->fail(ARSLAM_ERR_UNSUPPORTED, "%s: single-GPU handles only", what);
  const int64_t n = which == 0 ? s->n_cap : s->n_tag;
  if ((which != 0 && which != 1) || first < 0 || count < 0 || first + count > n)
    return s->fail(ARSLAM_ERR_INVALID, "%s: pose range [%lld, %lld) outside the problem's %lld poses", what, (long long)first,
                   (long long)(first + count), (long long)n);
  return ARSLAM_OK;
}

int arslam_set_poses(arslam_solver* s, int which, int64_t first, int64_t count, const double* pose6) {
  if (!s || !pose6) return ARSLAM_ERR_INVALID;
  const int rc = pose_range_ok(s, which, first, count, "set_poses");
  if (rc || count == 0) return rc;
  CU(cudaSetDevice(s->device));
  double* dst = (which == 0 ? s->cap[s->cur].p : s->tag[s->cur].p) + 6 * first;
  CU(cudaMemcpyAsync(dst, pose6, sizeof(double) * 6 * count, cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return ARSLAM_OK;
}

int arslam_get_poses(arslam_solver* s, int which, int64_t first, int64_t count, double* pose6) {
  if (!s || !pose6) return ARSLAM_ERR_INVALID;
  const int rc = pose_range_ok(s, which, first, count, "get_poses");
  if (rc || count == 0) return rc;
  CU(cudaSetDevice(s->device));
  const double* src = (which == 0 ? s->cap[s->cur].p : s->tag[s->cur].p) + 6 * first;
  CU(cudaMemcpyAsync(pose6, src, sizeof(double) * 6 * count, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return ARSLAM_OK;
}

int arslam_set_camera(arslam_solver* s, const double* camera3) {
  if (!s || !camera3) return ARSLAM_ERR_INVALID;
  if (!s->have_problem) return s->fail(ARSLAM_ERR_INVALID, "set_camera before set_problem");
  CU(cudaSetDevice(s->device));
  s->h_cam.assign(camera3, camera3 + 3);
  double* cam4 = s->h_sc + 200;
  cam4[0] = camera3[0]; cam4[1] = camera3[1]; cam4[2] = camera3[2]; cam4[3] = 0.0;
  CU(cudaMemcpyAsync(s->cam[s->cur].p, cam4, 4 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (!s->have_params) {  // a fresh problem whose poses will all be set / seeded on the device: they start at zero
    CU(cudaMemsetAsync(s->cap[s->cur].p, 0, sizeof(double) * 6 * s->n_cap, s->stream));
    CU(cudaMemsetAsync(s->tag[s->cur].p, 0, sizeof(double) * 6 * s->n_tag, s->stream));
    s->have_params = true;
    s->param_caps = s->n_cap; s->param_tags = s->n_tag;
  }
  return ARSLAM_OK;
}

static int seed_common(arslam_solver* s, int64_t n, const int32_t* a_idx, int64_t n_a, const int32_t* b_idx, int64_t n_b,
                       const double* rect8, const char* what, int32_t** d_a, int32_t** d_b, double** d_rect) {
  if (!s->have_problem || !s->have_params) return s->fail(ARSLAM_ERR_INVALID, "%s needs set_problem and parameters on the device", what);
  if (s->world > 1) return s->fail(ARSLAM_ERR_UNSUPPORTED, "%s: single-GPU handles only", what);
  for (int64_t i = 0; i < n; ++i)
    if (a_idx[i] < 0 || a_idx[i] >= n_a || b_idx[i] < 0 || b_idx[i] >= n_b) return s->fail(ARSLAM_ERR_INVALID, "%s: index out of range", what);
  CU(s->seed_idx.ensure((size_t)2 * n)); CU(s->seed_rect.ensure((size_t)8 * n));
  CU(cudaMemcpyAsync(s->seed_idx.p, a_idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->seed_idx.p + n, b_idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->seed_rect.p, rect8, sizeof(double) * 8 * n, cudaMemcpyHostToDevice, s->stream));
  *d_a = s->seed_idx.p; *d_b = s->seed_idx.p + n; *d_rect = s->seed_rect.p;
  return ARSLAM_OK;
}

int arslam_seed_captures(arslam_solver* s, int64_t n, const int32_t* cap_idx, const int32_t* tag_idx, const double* rect8) {
  if (!s || n < 0 || (n > 0 && (!cap_idx || !tag_idx || !rect8))) return ARSLAM_ERR_INVALID;
  if (n == 0) return ARSLAM_OK;
  CU(cudaSetDevice(s->device));
  int32_t *d_c, *d_t; double* d_r;
  const int rc = seed_common(s, n, cap_idx, s->n_cap, tag_idx, s->n_tag, rect8, "seed_captures", &d_c, &d_t, &d_r);
  if (rc) return rc;
  seed_captures_kernel<<<cdiv(n, 128), 128, 0, s->stream>>>((int)n, d_c, d_t, d_r, s->cam[s->cur].p, s->tag[s->cur].p, s->opt.tag_size, s->cap[s->cur].p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));  // the caller may reuse its arrays
  return ARSLAM_OK;
}

int arslam_seed_tags(arslam_solver* s, int64_t n, const int32_t* tag_idx, const int32_t* cap_idx, const double* rect8) {
  if (!s || n < 0 || (n > 0 && (!cap_idx || !tag_idx || !rect8))) return ARSLAM_ERR_INVALID;
  if (n == 0) return ARSLAM_OK;
  CU(cudaSetDevice(s->device));
  int32_t *d_t, *d_c; double* d_r;
  const int rc = seed_common(s, n, tag_idx, s->n_tag, cap_idx, s->n_cap, rect8, "seed_tags", &d_t, &d_c, &d_r);
  if (rc) return rc;
  seed_tags_kernel<<<cdiv(n, 128), 128, 0, s->stream>>>((int)n, d_t, d_c, d_r, s->cam[s->cur].p, s->cap[s->cur].p, s->opt.tag_size, s->tag[s->cur].p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(s->stream));
  return ARSLAM_OK;
}

int arslam_get_params(arslam_solver* s, double* camera3, double* cap_pose6, double* tag_pose6) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (!s->have_params) return s->fail(ARSLAM_ERR_INVALID, "get_params before set_params");
  CU(cudaSetDevice(s->device));
  const int k = s->cur;
  if (camera3) std::memcpy(camera3, s->h_cam.data(), 3 * sizeof(double));
  if (cap_pose6)
    CU(cudaMemcpyAsync(cap_pose6 + (size_t)6 * s->cap_lo, s->cap[k].p, sizeof(double) * 6 * s->n_cap, cudaMemcpyDeviceToHost, s->stream));
  if (tag_pose6) CU(cudaMemcpyAsync(tag_pose6, s->tag[k].p, sizeof(double) * 6 * s->n_tag, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return ARSLAM_OK;
}

static int launch_prep(arslam_solver* s, int k) {
  const int cap_ctas = cdiv(s->n_cap, 128);
  LAUNCH("prep_poses", (48.0 + 8.0 * kCapPre) * s->n_cap + (48.0 + 8.0 * kTagPre) * s->n_tag,
         prep_poses_kernel<<<cap_ctas + cdiv(s->n_tag, 128), 128, 0, s->stream>>>(s->n_cap, s->cap[k].p, s->cap_pre[k].p, s->n_tag, s->tag[k].p,
                                                                               s->opt.tag_size, s->tag_pre[k].p, cap_ctas, s->tag_cor[k].p, s->cap_rt[k].p));
  return ARSLAM_OK;
}

int arslam_evaluate(arslam_solver* s, double* cost, double* residuals, double* jac_cam, double* jac_cap, double* jac_tag) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (!s->have_problem || !s->have_params) return s->fail(ARSLAM_ERR_INVALID, "evaluate needs set_problem and set_params");
  CU(cudaSetDevice(s->device));
  s->prof.clear();
  const int nb = s->n_blk, nc = 4 * nb;
  const int nwarp = cdiv(nc, 256) * 8;
  const int kp = s->cur;
  launch_prep(s, kp);
  // outputs live in one scratch allocation: res 8 | jc 24 | jp 48 | ja 48 per block, then warp costs
  const size_t per_blk = 8 + 24 + 48 + 48;
  CU(s->eval_out.ensure(per_blk * nb + nwarp + 8));
  double* d_res = s->eval_out.p;
  double* d_jc = d_res + (size_t)8 * nb;
  double* d_jp = d_jc + (size_t)24 * nb;
  double* d_ja = d_jp + (size_t)48 * nb;
  double* d_wc = d_ja + (size_t)48 * nb;
  double* d_cost = d_wc + nwarp;
  const bool want_j = jac_cam || jac_cap || jac_tag;
  const bool dist = s->opt.num_intrinsics == 3;
  if (dist)
    LAUNCH("eval_jacobian", (want_j ? 274.0 : 34.0) * nc,
           eval_jacobian_kernel<1><<<cdiv(nc, 256), 256, 0, s->stream>>>(
               nc, s->o_cap.p, s->o_tag.p, reinterpret_cast<const double2*>(s->o_obs.p), s->cap_pre[kp].p, s->tag_pre[kp].p,
               s->cam[kp].p, reinterpret_cast<double2*>(d_res), want_j ? reinterpret_cast<double2*>(d_jc) : nullptr,
               want_j ? reinterpret_cast<double2*>(d_jp) : nullptr, want_j ? reinterpret_cast<double2*>(d_ja) : nullptr, d_wc));
  else
    LAUNCH("eval_jacobian", (want_j ? 274.0 : 34.0) * nc,
           eval_jacobian_kernel<0><<<cdiv(nc, 256), 256, 0, s->stream>>>(
               nc, s->o_cap.p, s->o_tag.p, reinterpret_cast<const double2*>(s->o_obs.p), s->cap_pre[kp].p, s->tag_pre[kp].p,
               s->cam[kp].p, reinterpret_cast<double2*>(d_res), want_j ? reinterpret_cast<double2*>(d_jc) : nullptr,
               want_j ? reinterpret_cast<double2*>(d_jp) : nullptr, want_j ? reinterpret_cast<double2*>(d_ja) : nullptr, d_wc));
  launch_colsum(s, nwarp, 1, d_wc, d_cost);
  CU(cudaGetLastError());
  double h_cost = 0.0;
  CU(cudaMemcpyAsync(&h_cost, d_cost, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (residuals) CU(cudaMemcpyAsync(residuals, d_res, sizeof(double) * 8 * nb, cudaMemcpyDeviceToHost, s->stream));
  if (jac_cam) CU(cudaMemcpyAsync(jac_cam, d_jc, sizeof(double) * 24 * nb, cudaMemcpyDeviceToHost, s->stream));
  if (jac_cap) CU(cudaMemcpyAsync(jac_cap, d_jp, sizeof(double) * 48 * nb, cudaMemcpyDeviceToHost, s->stream));
  if (jac_tag) CU(cudaMemcpyAsync(jac_tag, d_ja, sizeof(double) * 48 * nb, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->prof.resolve();
  if (cost) *cost = 0.5 * h_cost;
  return ARSLAM_OK;
}

// ------------------------------------------------------------------ solve ---
namespace {

// deterministic column sums, two stages when the input is large
int launch_colsum(arslam_solver* s, int n, int m, const double* in, double* out) {
  if (n <= 8192) {
    LAUNCH("colsum", 8.0 * n * m, colsum_kernel<<<1, 1024, 0, s->stream>>>(n, m, in, out));
  } else {
    LAUNCH("colsum", 8.0 * n * m, colsum_stage1_kernel<<<kColsumChunks, 256, 0, s->stream>>>(n, m, in, s->colsum_part.p));
    LAUNCH("colsum2", 4096.0, colsum_stage2_kernel<<<1, 384, 0, s->stream>>>(kColsumChunks, m, s->colsum_part.p, out));
  }
  return ARSLAM_OK;
}

// ---- multi-GPU helpers (one process per GPU; captures are sharded) ----------
__global__ void small_pack_kernel(const double* sc, double* buf, int rank, int world) {
  const int i = threadIdx.x;
  if (i < 8) buf[i] = sc[4 + i];   // cross, cand_r2, step2_e, xnorm2_e, mq_e, step2_f, xnorm2_f, mq_f
  if (i >= 8 && i < 8 + world) buf[i] = (i - 8 == rank) ? sc[16] : 0.0;
}
__global__ void small_unpack_kernel(double* sc, const double* buf, int world) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) sc[4 + i] = buf[i];
    double m = 0.0;
    for (int r = 0; r < world; ++r) m = fmax(m, buf[8 + r]);
    sc[16] = m;
  }
}

int nccl_sum(arslam_solver* s, double* buf, size_t count) {
  Profiler::Rec r{0, nullptr, nullptr};
  if (s->prof.on) {
    r = Profiler::Rec{s->prof.id_of(count > 4096 ? "nccl_allreduce_large" : "nccl_allreduce_small", 8.0 * count), s->prof.ev(), s->prof.ev()};
    cudaEventRecord(r.a, s->stream);
  }
  const int rc = g_nccl.AllReduce(buf, buf, count, kNcclDouble, kNcclSum, s->comm, s->stream);
  if (rc) return s->fail(ARSLAM_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  if (s->prof.on) {
    cudaEventRecord(r.b, s->stream);
    s->prof.recs.push_back(r);
  }
  return ARSLAM_OK;
}

// Multi-GPU: a rank-local failure (bad index, allocation) must fail EVERY rank, or the others wait in
// the next collective forever.  Max-reduces |code| over the ranks; returns `local` when this rank
// failed, ARSLAM_ERR_INVALID-style "a peer failed" otherwise, ARSLAM_OK when all are fine.
int comm_agree(arslam_solver* s, int local, const char* what) {
  if (s->world <= 1) return local;
  if (s->agree.ensure(1) != cudaSuccess) return s->fail(ARSLAM_ERR_CUDA, "comm_agree: allocation failed");
  long long v = local < 0 ? -local : local, worst = 0;
  CU(cudaMemcpyAsync(s->agree.p, &v, sizeof(v), cudaMemcpyHostToDevice, s->stream));
  const int rc = g_nccl.AllReduce(s->agree.p, s->agree.p, 1, kNcclInt64, kNcclMax, s->comm, s->stream);
  if (rc) return s->fail(ARSLAM_ERR_NCCL, "ncclAllReduce(status): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  CU(cudaMemcpyAsync(&worst, s->agree.p, sizeof(worst), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  if (local) return local;
  if (worst) return s->fail(-(int)worst, "%s: another rank failed (code %d); this rank stops with it", what, -(int)worst);
  return ARSLAM_OK;
}

// sums the eight block-local LM scalars; max of gmax_e via per-rank slots
int small_allreduce(arslam_solver* s, double* sc) {
  CU(s->small.ensure(8 + s->world));
  LAUNCH("small_pack", 128.0, small_pack_kernel<<<1, 64 + s->world, 0, s->stream>>>(sc, s->small.p, s->rank, s->world));
  int rc = nccl_sum(s, s->small.p, 8 + s->world);
  if (rc) return rc;
  LAUNCH("small_unpack", 128.0, small_unpack_kernel<<<1, 32, 0, s->stream>>>(sc, s->small.p, s->world));
  return ARSLAM_OK;
}

// ---- PCG hooks ---------------------------------------------------------------
// Multi-GPU: every rank must use the same block pattern of the reduced system (the partial
// values are summed element-wise), so the ranks' pair-key lists are all-gathered and united.
__global__ void pad_keys_kernel(unsigned long long* keys, long long from, long long to, unsigned long long sentinel) {
  const long long i = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < to) keys[i] = sentinel;
}
int union_keys_across_ranks(arslam_solver* s) {
  if (s->world <= 1) return ARSLAM_OK;
  PcgWorkspace& w = s->pcg;
  const unsigned long long sentinel = (unsigned long long)w.n_f * (unsigned long long)w.n_f;  // above every key
  DevBuf<long long> cnt;
  CU(cnt.ensure(1));
  long long n_local = w.nnzb, n_max = 0;
  CU(cudaMemcpyAsync(cnt.p, &n_local, sizeof(long long), cudaMemcpyHostToDevice, s->stream));
  int rc = g_nccl.AllReduce(cnt.p, cnt.p, 1, kNcclInt64, kNcclMax, s->comm, s->stream);
  if (rc) return s->fail(ARSLAM_ERR_NCCL, "ncclAllReduce(max): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  CU(cudaMemcpyAsync(&n_max, cnt.p, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  const long long n_all = n_max * s->world;
  if (n_all >= (1LL << 31)) return s->fail(ARSLAM_ERR_UNSUPPORTED, "pcg symbolic: more than 2^31 block pairs across ranks");
  // padded local list -> keys[1][0 .. n_max), gathered into keys[0][0 .. n_max * world)
  CU(w.keys[1].ensure((size_t)n_all));
  CU(cudaMemcpyAsync(w.keys[1].p, w.keys[0].p, sizeof(unsigned long long) * n_local, cudaMemcpyDeviceToDevice, s->stream));
  if (n_max > n_local)
    pad_keys_kernel<<<cdiv(n_max - n_local, 256), 256, 0, s->stream>>>(w.keys[1].p, n_local, n_max, sentinel);
  CU(w.keys[0].ensure((size_t)n_all));
  rc = g_nccl.AllGather(w.keys[1].p, w.keys[0].p, (size_t)n_max, kNcclInt64, s->comm, s->stream);
  if (rc) return s->fail(ARSLAM_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  int nn = 0;
  CU(pcg_sort_unique(w, (int)n_all, sym_key_bits(w.n_f), s->stream, &nn));
  // the sentinel, if any rank padded, sorts last
  if (nn > 0) {
    unsigned long long last = 0;
    CU(cudaMemcpyAsync(&last, w.keys[0].p + (nn - 1), sizeof(last), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (last == sentinel) --nn;
  }
  w.nnzb = nn;
  return ARSLAM_OK;
}

int pcg_prepare(arslam_solver* s, int side_e, int n_e, int n_f) {
  if (s->pcg.valid && s->pcg_version == s->problem_version && s->pcg_side == side_e) return ARSLAM_OK;
  std::string err;
  int rc = pcg_symbolic_keys(s->pcg, n_e, n_f, s->n_blk, s->s_own[side_e].p, s->s_off[side_e].p, s->seg_end(side_e), s->s_oth[side_e].p, s->stream, err);
  rc = comm_agree(s, rc ? s->fail(ARSLAM_ERR_CUDA, "%s", err.c_str()) : ARSLAM_OK, "pcg symbolic phase");
  if (rc) return rc;
  const int urc = union_keys_across_ranks(s);
  if (urc) return urc;
  rc = pcg_symbolic_build(s->pcg, s->stream, err, s->n_sm, (size_t)227 * 1024 - 2048);
  rc = comm_agree(s, rc ? s->fail(ARSLAM_ERR_CUDA, "%s", err.c_str()) : ARSLAM_OK, "pcg symbolic phase");
  if (rc) return rc;
  int occ = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_kernel<3>, kPcgThreads, 0));
  if (occ < 1) return s->fail(ARSLAM_ERR_CUDA, "pcg_kernel cannot be made resident");
  s->pcg.grid = std::min(s->n_sm * std::min(occ, 1), 1024);
  if (s->pcg.pair_slot) {
    SparseTarget t;
    t.row_ptr = s->pcg.row_ptr; t.col_idx = s->pcg.col_idx; t.Sraw = nullptr; t.borderm = nullptr; t.rhsm = nullptr; t.borderx = nullptr; t.n_f = s->pcg.n_f;
    t.pair_slot = nullptr; t.lower_of = s->pcg.src_slot;
    LAUNCH("pair_slot", 4.0 * s->pcg.n_pairs,
           pair_slot_kernel<<<cdiv(s->n_blk, 128), 128, 0, s->stream>>>(s->n_blk, s->s_own[side_e].p, s->s_off[side_e].p, s->seg_end(side_e),
                                                                      s->s_oth[side_e].p, s->pcg.pair_off, t, s->pcg.pair_slot));
  }
  // the locality plan of the elimination kernel hangs off the same pair-slot table
  s->schur_plan.valid = false;
  if (s->pcg.pair_slot && s->tune_schur_local) {
    const int prc = schur_local_build(s->schur_plan, s->pcg, n_e, s->n_blk, s->s_own[side_e].p, s->s_off[side_e].p, s->seg_end(side_e), s->s_oth[side_e].p, s->n_sm, s->stream, err);
    if (prc) return s->fail(ARSLAM_ERR_CUDA, "%s", err.c_str());
  }
  s->pcg_version = s->problem_version;
  s->pcg_side = side_e;
  return ARSLAM_OK;
}

int pcg_launch_eliminate(arslam_solver* s, const SchurArgs& a, double* Sraw, const int32_t* e_idx, int nk) {
  SparseTarget t;
  t.row_ptr = s->pcg.row_ptr; t.col_idx = s->pcg.col_idx;
  t.Sraw = Sraw;
  t.borderm = Sraw + (size_t)36 * s->pcg.nnz_lower;
  t.rhsm = t.borderm + (size_t)6 * s->pcg.n_f;
  t.borderx = nk == 3 ? t.rhsm + (size_t)6 * s->pcg.n_f : nullptr;
  t.n_f = s->pcg.n_f;
  t.pair_slot = s->pcg.pair_slot; t.lower_of = s->pcg.src_slot;
  SchurArgs a2 = a;
  a2.pair_off = s->pcg.pair_slot ? s->pcg.pair_off : nullptr;
  if (nk == 3) {
    launch_schur<SparseTarget, 3>(s, a2, t, e_idx, (288.0 + 8) * s->n_blk + (264.0 + 128 + 96) * a.n_e + 288.0 * s->pcg.nnz_lower,
                                  s->tune_schur_bulk != 0);
    return ARSLAM_OK;
  }
  if (s->schur_plan.valid && s->tune_schur_local) {
    // locality-ordered CTAs, products pre-reduced per destination inside the CTA; no straddle launch
    const double bytes = (288.0 + 8) * s->n_blk + (264.0 + 128) * a.n_e + 288.0 * s->pcg.nnz_lower + 2.0 * s->schur_plan.n_pairs;
    const SchurLocalView v = s->schur_plan.view();
    if (s->tune_schur_bulk)
      LAUNCH("schur_eliminate", bytes, schur_local_kernel<true><<<s->schur_plan.n_cta, kSlThreads, kSlSmem, s->stream>>>(a2, t, s->n_blk, e_idx, v));
    else
      LAUNCH("schur_eliminate", bytes, schur_local_kernel<false><<<s->schur_plan.n_cta, kSlThreads, kSlSmem, s->stream>>>(a2, t, s->n_blk, e_idx, v));
    return ARSLAM_OK;
  }
  // compulsory bytes: W read once (288 B / block) + indices (8 B / block) + E records and Z / YB (264 + 128 B / pose)
  // + the lower blocks of the reduced system written once (288 B each)
  launch_schur<SparseTarget, 1>(s, a2, t, e_idx, (288.0 + 8) * s->n_blk + (264.0 + 128) * a.n_e + 288.0 * s->pcg.nnz_lower,
                                s->tune_schur_bulk != 0);
  return ARSLAM_OK;
}

int pcg_launch_solve(arslam_solver* s, int n_f, double* Sraw, const double* HF, const double* HFx, int nk, const double* sc,
                     const double* cam_minus, double radius, double* x_out) {
  PcgWorkspace& w = s->pcg;
  const size_t nvec = pcg_nvec(n_f);
  double* v = w.vec;
  PcgFinalizeArgs f;
  f.n_f = n_f; f.nnzb = w.nnzb; f.row_ptr = w.row_ptr; f.col_idx = w.col_idx; f.src_slot = w.src_slot;
  f.Sraw = Sraw; f.borderm = Sraw + (size_t)36 * w.nnz_lower; f.rhsm = f.borderm + (size_t)6 * n_f;
  f.HF = HF; f.sigF = s->sigF.p; f.sc = reinterpret_cast<const LmScalars*>(sc); f.cam_minus = cam_minus;
  f.radius = radius; f.min_diag = s->opt.min_lm_diagonal; f.max_diag = s->opt.max_lm_diagonal;
  f.Sfin = w.Sfin; f.Minv = w.Minv; f.border = v + 6 * nvec; f.rhs = v + 7 * nvec; f.scal = w.scal; f.partial = w.partial;
  f.nk = nk; f.HFx = HFx; f.borderx = nk == 3 ? f.rhsm + (size_t)6 * n_f : nullptr; f.border1 = v + 8 * nvec; f.border2 = v + 9 * nvec;
  LAUNCH("pcg_finalize_offdiag", 2.0 * 288.0 * w.nnzb,
         pcg_finalize_offdiag_kernel<<<cdiv((long long)w.nnzb * 6, 256), 256, 0, s->stream>>>(f, w.slot_row));
  LAUNCH("pcg_finalize", 8.0 * (NV + 36 + 36 + 36 + 24) * n_f,
         pcg_finalize_kernel<<<cdiv((long long)std::max(n_f, 1) * kFinLanes, 128), 128, 0, s->stream>>>(f, w.diag_slot));
  PcgSmemArgs sa;
  PcgArgs& a = sa.a;
  a.n_f = n_f; a.max_iter = s->opt.pcg_max_iterations; a.tol = s->opt.pcg_tolerance; a.q_tol = s->opt.pcg_q_tolerance;
  a.row_ptr = w.row_ptr; a.col_idx = w.col_idx; a.S = w.Sfin; a.Minv = w.Minv;
  a.border = f.border; a.border1 = f.border1; a.border2 = f.border2; a.rhs = f.rhs;
  a.x = x_out; a.r = v + 1 * nvec; a.z = v + 2 * nvec; a.p0 = v + 3 * nvec; a.p1 = v + 4 * nvec; a.q = v + 5 * nvec;
  a.partial = w.partial; a.scal = w.scal; a.trace = nullptr;
  a.sigF = s->sigF.p; a.uF = s->uF.p; a.sc = const_cast<double*>(sc);
  sa.cta_row = w.cta_row; sa.halo_ptr = w.halo_ptr; sa.halo_col = w.halo_col; sa.lcol = w.lcol;
  sa.cap_slots = w.cap_slots; sa.max_halo = w.max_halo; sa.max_slots = w.max_slots; sa.max_rows = w.max_rows;
  const bool smem_fits = w.smem_ok && s->tune_pcg_smem;
  const bool use_smem = smem_fits && nk == 1;  // the classic shared-memory kernel knows the focal-only border
  void* args_g[] = {(void*)&a};
  void* args_s[] = {(void*)&sa};
  Profiler::Rec r{0, nullptr, nullptr};
  if (s->prof.on) {
    r = Profiler::Rec{s->prof.id_of("pcg_solve", 288.0 * w.nnzb), s->prof.ev(), s->prof.ev()};
    cudaEventRecord(r.a, s->stream);
  }
  // one barrier per iteration for the inexact-Newton tolerances; the classic recurrence when the
  // system is to be solved tightly (its attainable accuracy is higher)
  // (radial model: pcg_pipe_kernel<3>, or pcg_kernel<3> for the tight solves)
  const bool pipelined = smem_fits && (s->opt.pcg_tolerance >= 1e-6 || s->opt.pcg_q_tolerance > 0.0) && s->tune_pcg_pipelined;
  if (pipelined)
    CU(cudaLaunchCooperativeKernel(nk == 3 ? (void*)pcg_pipe_kernel<3> : (void*)pcg_pipe_kernel<1>, dim3(w.smem_grid), dim3(kPcgThreads), args_s,
                                   w.smem_bytes, s->stream));
  else if (use_smem)
    CU(cudaLaunchCooperativeKernel((void*)pcg_smem_kernel, dim3(w.smem_grid), dim3(kPcgThreads), args_s, w.smem_bytes, s->stream));
  else
    CU(cudaLaunchCooperativeKernel(nk == 3 ? (void*)pcg_kernel<3> : (void*)pcg_kernel<1>, dim3(w.grid), dim3(kPcgThreads), args_g, 0, s->stream));
  if (s->prof.on) {
    cudaEventRecord(r.b, s->stream);
    s->prof.recs.push_back(r);
  }
  ++s->launches;
  return ARSLAM_OK;
}

struct Sides {
  int e, f;        // side index (0 capture, 1 tag) of the eliminated / retained poses
  int n_e, n_f;
};

// HF / HFx: F-side records (inside the cross-rank reduction buffer); head: 12 doubles
// [H_ff, g_f, sum r^2, 0 | f.l1, f.l2, l1.l1, l1.l2, l2.l2, g_l1, g_l2, 0]
int launch_accumulate(arslam_solver* s, const Sides& sd, int k, double* HF, double* HFx, double* head) {
  const int nb = s->n_blk, plane = s->plane, grid = cdiv(plane, kAccumThreads);
  const bool dist = s->opt.num_intrinsics == 3;
  int pipe_grid = 0;  // CTAs of the pipelined E pass (its camera partials), 0: accum_kernel ran
  FixupJobs fix;
  fix.n = 0;
  int fix_ctas = 0;
  auto add_fixup = [&](int side, int n_own, int nv, const double* partial, double* out_seg) {
    FixupJob& f = fix.j[fix.n++];
    f.n_pose = n_own; f.n_blk = nb; f.nv = nv;
    f.own_idx = s->s_own[side].p; f.seg_off = s->s_off[side].p; f.seg_end = s->seg_end(side); f.partial = partial; f.out_seg = out_seg;
    fix_ctas += cdiv((long long)std::max(s->n_warp - 1, 0) * nv, 256);
    f.cta_zero = fix_ctas;
    fix_ctas += cdiv(n_own, 256);
    f.cta_end = fix_ctas;
  };
  for (int pass = 0; pass < 2; ++pass) {
    const int side = pass == 0 ? sd.e : sd.f;
    const int n_own = side == 0 ? s->n_cap : s->n_tag;
    AccumArgs a;
    a.n_blk = nb; a.plane = plane;
    a.own_idx = s->s_own[side].p; a.oth_idx = s->s_oth[side].p; a.obs = s->s_obs[side].p;
    a.cap_pre = s->cap_pre[k].p; a.tag_pre = s->tag_pre[k].p; a.cam = s->cam[k].p;
    a.out_seg = pass == 0 ? s->H[side].p : HF;
    a.partial = s->partial[side].p;
    a.W = s->W.p; a.warp_cam = s->warp_cam.p;
    const double bytes_e = (4.0 * 18.0 + 288.0) * nb + 264.0 * n_own;
    const double bytes_f = (4.0 * 18.0) * nb + 264.0 * n_own;
#define ARS_ACC(SIDE_, W_, NAME_, BYTES_)                                                                      \
    do {                                                                                                        \
      if (dist) LAUNCH(NAME_, BYTES_, accum_kernel<SIDE_, W_, 1><<<grid, kAccumThreads, 0, s->stream>>>(a));    \
      else LAUNCH(NAME_, BYTES_, accum_kernel<SIDE_, W_, 0><<<grid, kAccumThreads, 0, s->stream>>>(a));         \
    } while (0)
    // the hot role/side combinations (captures eliminated) run the cross-block pipelined kernels
    const int n_chunks = cdiv(nb, kPipeThreads);
    if (pass == 0 && side == 0 && (s->tune_accum_pipe & 1)) {
      pipe_grid = std::min(n_chunks, 2 * s->n_sm);
#define ARS_EP(M_, U_) LAUNCH("accum_E", bytes_e, accum_e_pipe_kernel<M_, U_><<<pipe_grid, kPipeThreads, kEPipeSmem, s->stream>>>(a, n_chunks))
      if (dist) { if (s->tune_accum_flush) ARS_EP(1, true); else ARS_EP(1, false); }
      else { if (s->tune_accum_flush) ARS_EP(0, true); else ARS_EP(0, false); }
#undef ARS_EP
    } else if (pass == 1 && side == 1 && (s->tune_accum_pipe & 2)) {
      AccumFArgs fa;
      fa.a = a; fa.cap_rt = s->cap_rt[k].p;
      const int g = std::min(n_chunks, 3 * s->n_sm);
#define ARS_FP(M_, U_) LAUNCH("accum_F", bytes_f, accum_f_pipe_kernel<M_, U_><<<g, kPipeThreads, kFPipeSmem, s->stream>>>(fa, n_chunks))
      if (dist) { if (s->tune_accum_flush) ARS_FP(1, true); else ARS_FP(1, false); }
      else { if (s->tune_accum_flush) ARS_FP(0, true); else ARS_FP(0, false); }
#undef ARS_FP
    } else if (pass == 0) {
      if (side == 0) ARS_ACC(0, true, "accum_E", bytes_e); else ARS_ACC(1, true, "accum_E", bytes_e);
    } else {
      if (side == 0) ARS_ACC(0, false, "accum_F", bytes_f); else ARS_ACC(1, false, "accum_F", bytes_f);
    }
#undef ARS_ACC
    add_fixup(side, n_own, NV, s->partial[side].p, a.out_seg);
    if (dist) {  // the l1, l2 columns
      AccumCamArgs c;
      c.n_blk = nb; c.plane = plane;
      c.own_idx = a.own_idx; c.oth_idx = a.oth_idx; c.obs = a.obs;
      c.cap_pre = a.cap_pre; c.tag_pre = a.tag_pre; c.cam = a.cam;
      c.out_seg = pass == 0 ? s->Hx[side].p : HFx;
      c.partial = s->partialx[side].p;
      c.warp_cam = s->warp_cam8.p;
      const double bytes_c = 72.0 * nb + 96.0 * n_own;
      if (pass == 0) {
        if (side == 0) LAUNCH("accum_cam_E", bytes_c, accum_cam_kernel<0, true><<<grid, kAccumThreads, 0, s->stream>>>(c));
        else LAUNCH("accum_cam_E", bytes_c, accum_cam_kernel<1, true><<<grid, kAccumThreads, 0, s->stream>>>(c));
      } else {
        if (side == 0) LAUNCH("accum_cam_F", bytes_c, accum_cam_kernel<0, false><<<grid, kAccumThreads, 0, s->stream>>>(c));
        else LAUNCH("accum_cam_F", bytes_c, accum_cam_kernel<1, false><<<grid, kAccumThreads, 0, s->stream>>>(c));
      }
      add_fixup(side, n_own, NVX, s->partialx[side].p, c.out_seg);
    }
  }
  // the pieces of segments that straddle warps, all records of both passes in one launch
  LAUNCH("seg_fixup", 16.0 * NV * s->n_warp, seg_fixup_kernel<<<fix_ctas, 256, 0, s->stream>>>(fix));
  LAUNCH("reduce_partials", 32.0 * grid, reduce_partials_kernel<4, false><<<1, 1024, 0, s->stream>>>(s->warp_cam.p, pipe_grid ? pipe_grid : grid, head));
  if (dist) launch_colsum(s, s->n_warp, 8, s->warp_cam8.p, head + 4);
  return ARSLAM_OK;
}

__global__ void cam_sigma_kernel(double* sc, double* sigF_cam, int enabled, int nk, int cam_const) {
  const double h[3] = {sc[0], sc[26], sc[28]};  // H_ff, H_l1l1, H_l2l2
  const int slot[3] = {14, 34, 35};
  for (int q = 0; q < nk; ++q) {
    const double sg = cam_const ? 0.0 : (enabled ? 1.0 / (1.0 + sqrt(h[q])) : 1.0);
    sc[slot[q]] = sg;
    sigF_cam[q] = sg;
  }
}

}  // namespace

int arslam_solve(arslam_solver* s, arslam_summary* summary, double* iter_log, int32_t log_rows) {
  if (!s || !summary) return ARSLAM_ERR_INVALID;
  CU(cudaSetDevice(s->device));
  const double t_start = wall_ms();
  const arslam_options& o = s->opt;
  std::memset(summary, 0, sizeof(*summary));
  s->launches = 0;
  s->prof.clear();

  // ---- which pose side is eliminated
  Sides sd;
  int elim = o.elimination;
  // AUTO eliminates the larger pose set; under arslam_comm_init the captures are sharded (n_cap is this
  // rank's range) and are the only side that can be eliminated locally
  if (elim == ARSLAM_ELIM_AUTO)
    elim = (s->world > 1 || s->n_cap >= s->n_tag) ? ARSLAM_ELIM_CAPTURES : ARSLAM_ELIM_TAGS;
  {
    // every rank-local reason to refuse is settled among the ranks before the first collective
    int pre = ARSLAM_OK;
    if (!s->have_problem || !s->have_params) pre = s->fail(ARSLAM_ERR_INVALID, "solve needs set_problem and set_params");
    else if (s->world > 1 && elim != ARSLAM_ELIM_CAPTURES)
      pre = s->fail(ARSLAM_ERR_UNSUPPORTED, "multi-GPU solve shards captures and must eliminate them");
    pre = comm_agree(s, pre, "solve");
    if (pre) return pre;
  }
  sd.e = elim == ARSLAM_ELIM_CAPTURES ? 0 : 1;
  sd.f = 1 - sd.e;
  sd.n_e = sd.e == 0 ? s->n_cap : s->n_tag;
  sd.n_f = sd.f == 0 ? s->n_cap : s->n_tag;
  const unsigned char* const_e = s->have_const ? (sd.e == 0 ? s->const_cap.p : s->const_tag.p) : nullptr;
  const unsigned char* const_f = s->have_const ? (sd.f == 0 ? s->const_cap.p : s->const_tag.p) : nullptr;
  const int nk = o.num_intrinsics == 3 ? 3 : 1;  // live intrinsics: focal (+ l1, l2 of the radial model)
  const bool dist = nk == 3;
  const int n = 6 * sd.n_f + nk;  // reduced dimension (F poses + intrinsics)
  const int cam_row = 6 * sd.n_f, rhs_row = n;
  // AUTO: the dense DMMA Cholesky where the reduced matrix is actually dense (>= 25 % of its
  // 6x6 blocks are structurally non-zero, or it is tiny), block-sparse PCG otherwise
  int lin = o.linear_solver;
  if (lin == ARSLAM_LINSOLVE_AUTO) {
    lin = ARSLAM_LINSOLVE_PCG;
    if (n <= 128) {
      lin = ARSLAM_LINSOLVE_DENSE;
    } else if (n <= o.dense_max_dim) {
      int prc = pcg_prepare(s, sd.e, sd.n_e, sd.n_f);
      if (prc) return prc;
      const double fill = (double)s->pcg.nnzb / ((double)sd.n_f * (double)sd.n_f);
      if (fill >= 0.25) lin = ARSLAM_LINSOLVE_DENSE;
    }
  }
  summary->eliminated_side = elim;
  summary->linear_solver = lin;
  summary->reduced_dim = n;

  // ---- buffers that depend on the roles
  CU(s->Z.ensure((size_t)8 * sd.n_e)); CU(s->YB.ensure((size_t)6 * nk * sd.n_e)); CU(s->seg_cam.ensure((size_t)12 * (cdiv(s->n_blk, 90) + 2)));  // CTA partials of either elimination kernel (the local one packs >= 96 blocks per CTA)
  // E poses without blocks never reach the elimination kernel: their Z / YB stay zero for the whole solve
  CU(cudaMemsetAsync(s->Z.p, 0, sizeof(double) * 8 * sd.n_e, s->stream));
  CU(cudaMemsetAsync(s->YB.p, 0, sizeof(double) * 6 * nk * sd.n_e, s->stream));
  if (dist) {
    CU(s->Hx[sd.e].ensure((size_t)NVX * sd.n_e));
    for (int side = 0; side < 2; ++side) CU(s->partialx[side].ensure((size_t)s->n_warp * 2 * NVX));
    CU(s->warp_cam8.ensure((size_t)8 * s->n_warp));
  }
  CU(s->seg_cross.ensure((size_t)5 * (cdiv((long long)sd.n_e * kBsGroup, 128) + 1)));
  CU(s->sigE.ensure((size_t)6 * sd.n_e)); CU(s->sigF.ensure((size_t)n + 1)); CU(s->uF.ensure((size_t)n + 1));
  size_t s_elems = 0;
  if (lin == ARSLAM_LINSOLVE_DENSE) {
    s->n_pad = (n + 1 + CB - 1) / CB * CB;
    s->ld = s->n_pad;
    s_elems = (size_t)s->n_pad * s->ld;
    CU(s->linv.ensure((size_t)(s->n_pad / CB) * CB * CB));
    CU(s->yF.ensure((size_t)s->n_pad));
    if (s->world > 1) {
      const size_t had = s->tri.n;
      CU(s->tri.ensure(dense_packed_count(s->n_pad)));
      if (s->tri.n != had) CU(cudaMemsetAsync(s->tri.p, 0, s->tri.n * sizeof(double), s->stream));  // the slots past ld are summed too, never read
    }
  } else {
    int rc = pcg_prepare(s, sd.e, sd.n_e, sd.n_f);
    if (rc) return rc;
    s_elems = s->pcg.value_count(nk);  // block values + border + rhs (+ the l1, l2 borders)
    CU(s->yF.ensure((size_t)n + 1));
  }
  // reduction buffer: [S or sparse values | cam_minus (12) | HF (n_f NV) | HFx (n_f NVX, radial model) | head (12)]
  const size_t red_tail = 12 + (size_t)sd.n_f * NV + (dist ? (size_t)sd.n_f * NVX : 0) + 12;
  CU(s->red.ensure(s_elems + red_tail));
  double* S = s->red.p;
  double* cam_minus = S + s_elems;
  double* HF = cam_minus + 12;
  double* HFx = dist ? HF + (size_t)sd.n_f * NV : nullptr;
  double* sc_head = HF + (size_t)sd.n_f * NV + (dist ? (size_t)sd.n_f * NVX : 0);  // summed across ranks
  CU(cudaMemsetAsync(sc_head, 0, 12 * sizeof(double), s->stream));
  double* sc = s->sc.p;
  CU(cudaMemsetAsync(sc, 0, kNumScalars * sizeof(double), s->stream));

  int rc = ARSLAM_OK;
  // multi-GPU: an F pose (tag) is live when ANY rank holds a block of it; a rank that holds none must
  // still move it, or the replicated tag poses drift apart
  const double* f_blocks_all = nullptr;
  if (s->world > 1) {
    CU(s->f_blocks.ensure((size_t)sd.n_f));
    segment_sizes_kernel<<<cdiv(sd.n_f, 256), 256, 0, s->stream>>>(sd.n_f, s->s_off[sd.f].p, s->seg_end(sd.f), s->f_blocks.p);
    int rcb = nccl_sum(s, s->f_blocks.p, (size_t)sd.n_f);
    if (rcb) return rcb;
    f_blocks_all = s->f_blocks.p;
  }
  launch_prep(s, s->cur);
  launch_accumulate(s, sd, s->cur, HF, HFx, sc_head);

  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool have_sigma = false, last_successful = true, fresh_linearisation = true, pending_acc = false;
  int iteration = 0, invalid = 0;
  int termination = ARSLAM_NO_CONVERGENCE, reason = ARSLAM_REASON_MAX_ITERATIONS;
  double x_cost = 0.0, grad_max = 0.0, x_norm = 0.0;
  double eval_ms = 0.0, lin_ms = 0.0;
  summary->num_jacobian_evals = 1;
  summary->num_successful_steps = 1;

  auto log_iter = [&](int it, double cost, double cost_change, double step_norm, double rho, int valid, int ok) {
    if (iter_log && it < log_rows) {
      double* L = iter_log + 8 * (size_t)it;
      L[0] = cost; L[1] = cost_change; L[2] = grad_max; L[3] = step_norm;
      L[4] = rho; L[5] = radius; L[6] = valid; L[7] = ok;
    }
    if (o.verbose)
      std::printf("%4d  cost %.6e  change %.3e  |grad| %.3e  |step| %.3e  rho %.3e  radius %.3e%s\n", it, cost,
                  cost_change, grad_max, step_norm, rho, radius, valid ? (ok ? "" : "  (rejected)") : "  (invalid)");
  };

  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue, the tests that need no new data
    if (iteration >= o.max_num_iterations && iteration > 0) {
      termination = ARSLAM_NO_CONVERGENCE; reason = ARSLAM_REASON_MAX_ITERATIONS; break; }
    if (radius <= o.min_trust_region_radius) {
      termination = ARSLAM_CONVERGENCE; reason = ARSLAM_REASON_MIN_RADIUS; break; }
    const bool stop_after_readback = iteration >= o.max_num_iterations;  // max_num_iterations == 0

    // ---------------- linear solve for this radius (speculative; see below)
    const int k = s->cur, kc = 1 - s->cur;
    cudaEventRecord(s->ev[0], s->stream);
    if (!have_sigma) {
      LAUNCH("sigma", 8.0 * 12 * sd.n_e,
             sigma_pose_kernel<<<cdiv(sd.n_e, 128), 128, 0, s->stream>>>(sd.n_e, s->H[sd.e].p, o.jacobi_scaling, s->sigE.p, const_e));
    }
    SchurArgs schur_args;
    {
      SchurArgs a;
      a.n_e = sd.n_e; a.plane = s->plane;
      a.e_off = s->s_off[sd.e].p; a.e_end = s->seg_end(sd.e); a.f_idx = s->s_oth[sd.e].p;
      a.HE = s->H[sd.e].p; a.HEx = dist ? s->Hx[sd.e].p : nullptr; a.W = s->W.p; a.sig_e = s->sigE.p;
      a.radius = radius; a.inv_radius = 1.0 / radius; a.min_diag = o.min_lm_diagonal; a.max_diag = o.max_lm_diagonal;
      a.Z = s->Z.p; a.YB = s->YB.p; a.seg_cam = s->seg_cam.p; a.pair_off = nullptr; a.e_const = const_e;
      a.straddle_ctas = s->straddle_list[sd.e].p; a.n_straddle = s->n_straddle[sd.e];
      schur_args = a;
      if (lin == ARSLAM_LINSOLVE_DENSE) {
        LAUNCH("dense_zero", 4.0 * s->n_pad * (double)s->n_pad,
               dense_zero_lower_kernel<<<dim3(cdiv(s->n_pad, kDenseTileCols), cdiv(s->n_pad, kDenseTileRows)), 256, 0, s->stream>>>(S, s->ld, s->n_pad));
      } else {
        CU(cudaMemsetAsync(S, 0, s_elems * sizeof(double), s->stream));
      }
      if (lin == ARSLAM_LINSOLVE_DENSE) {
        DenseTarget t;
        t.S = S; t.ld = s->ld; t.cam_row = cam_row; t.rhs_row = rhs_row;
        if (dist) launch_schur<DenseTarget, 3>(s, a, t, s->s_own[sd.e].p, (288.0 + 8) * s->n_blk + (264.0 + 128) * sd.n_e + 8.0 * n * (double)n / 2);
        else launch_schur<DenseTarget, 1>(s, a, t, s->s_own[sd.e].p, (288.0 + 8) * s->n_blk + (264.0 + 128) * sd.n_e + 8.0 * n * (double)n / 2);
      } else {
        rc = pcg_launch_eliminate(s, a, S, s->s_own[sd.e].p, nk);
        if (rc) return rc;
      }
      // closes the elimination: cam_minus, max |g_e| -> sc[16], and the failure flag of the solve that starts here
      const int schur_ctas = (lin != ARSLAM_LINSOLVE_DENSE && s->schur_plan.valid && s->tune_schur_local) ? s->schur_plan.n_cta : cdiv(s->n_blk, kSchurThreads);
      LAUNCH("schur_finish", 96.0 * schur_ctas, schur_finish_kernel<<<1, 1024, 0, s->stream>>>(s->seg_cam.p, schur_ctas, cam_minus, sc));
    }
    if (s->world > 1) {
      // one allreduce per linearisation: partial Schur terms (+ on a fresh
      // linearisation the partial tag blocks, focal terms and cost)
      const size_t tail = fresh_linearisation ? red_tail : 12;
      if (lin == ARSLAM_LINSOLVE_DENSE) {
        // only the populated region of the square travels (half the bytes), packed into a contiguous buffer
        const dim3 grd(cdiv(s->n_pad, kDenseTileCols), cdiv(s->n_pad, kDenseTileRows));
        LAUNCH("dense_pack", 8.0 * s->n_pad * (double)s->n_pad, dense_pack_kernel<false><<<grd, 256, 0, s->stream>>>(S, s->ld, s->n_pad, s->tri.p));
        rc = nccl_sum(s, s->tri.p, dense_packed_count(s->n_pad));
        if (rc) return rc;
        LAUNCH("dense_unpack", 8.0 * s->n_pad * (double)s->n_pad, dense_pack_kernel<true><<<grd, 256, 0, s->stream>>>(S, s->ld, s->n_pad, s->tri.p));
        rc = nccl_sum(s, S + s_elems, tail);
      } else {
        rc = nccl_sum(s, S, s_elems + tail);
      }
      if (rc) return rc;
    }
    if (fresh_linearisation) {
      // (the E side's maximum comes out of the elimination kernel; this launch also moves the freshly summed
      // camera scalars next to the other LM scalars)
      LAUNCH("gradmax", 8.0 * 6 * sd.n_f, gradmax_kernel<<<cdiv(sd.n_f, 128), 128, 0, s->stream>>>(sd.n_f, s->s_off[sd.f].p, s->seg_end(sd.f), HF, s->warp_gmax[sd.f].p, s->tickets.p + 3, sc + 17, sc_head, sc, nk, f_blocks_all, const_f));
    }
    if (!have_sigma) {
      LAUNCH("sigma", 8.0 * 12 * sd.n_f,
             sigma_pose_kernel<<<cdiv(sd.n_f, 128), 128, 0, s->stream>>>(sd.n_f, HF, o.jacobi_scaling, s->sigF.p, const_f));
      LAUNCH("cam_sigma", 16.0, cam_sigma_kernel<<<1, 1, 0, s->stream>>>(sc, s->sigF.p + cam_row, o.jacobi_scaling, nk, s->cam_const));
      have_sigma = true;
    }
    if (lin == ARSLAM_LINSOLVE_DENSE) {
      LAUNCH("dense_scale", 8.0 * n * n,
             dense_scale_kernel<<<dim3(cdiv(n, kDenseTileCols), cdiv(n + 1, kDenseTileRows)), 256, 0, s->stream>>>(S, s->ld, rhs_row, s->sigF.p));
      LAUNCH("dense_add_pose", 8.0 * NV * sd.n_f,
             dense_add_pose_kernel<<<cdiv(sd.n_f, 128), 128, 0, s->stream>>>(sd.n_f, HF, HFx, s->sigF.p, radius, o.min_lm_diagonal, o.max_lm_diagonal, S, s->ld, cam_row, rhs_row));
      LAUNCH("dense_add_camera", 64.0,
             dense_add_camera_kernel<<<cdiv(std::max(1, s->n_pad - rhs_row - 1), 128), 128, 0, s->stream>>>(
                 reinterpret_cast<LmScalars*>(sc), cam_minus, nk, radius, o.min_lm_diagonal, o.max_lm_diagonal, S, s->ld, cam_row, rhs_row, s->n_pad));
      if (s->prof.on) {
        Profiler::Rec r{s->prof.id_of("dense_cholesky", 0.0), s->prof.ev(), s->prof.ev()};
        cudaEventRecord(r.a, s->stream);
        s->launches += DenseCholesky::factor(S, s->ld, s->n_pad, s->linv.p, sc + 12, s->stream, s->lookahead);
        s->launches += DenseCholesky::backsolve(S, s->ld, n, rhs_row, s->yF.p, s->linv.p, sc + 12, s->stream, s->lookahead, s->tune_chol_chain != 0);
        cudaEventRecord(r.b, s->stream);
        s->prof.recs.push_back(r);
      } else {
        s->launches += DenseCholesky::factor(S, s->ld, s->n_pad, s->linv.p, sc + 12, s->stream, s->lookahead);
        s->launches += DenseCholesky::backsolve(S, s->ld, n, rhs_row, s->yF.p, s->linv.p, sc + 12, s->stream, s->lookahead, s->tune_chol_chain != 0);
      }
    } else {
      rc = pcg_launch_solve(s, sd.n_f, S, HF, HFx, nk, sc, cam_minus, radius, s->yF.p);
      if (rc) return rc;
    }
    if (lin == ARSLAM_LINSOLVE_DENSE)  // (the PCG kernels write uF themselves)
      LAUNCH("scale_uF", 24.0 * n, scale_uF_kernel<<<cdiv(n, 256), 256, 0, s->stream>>>(n, s->yF.p, s->sigF.p, s->uF.p));
    double* x_e = sd.e == 0 ? s->cap[k].p : s->tag[k].p;
    double* x_f = sd.f == 0 ? s->cap[k].p : s->tag[k].p;
    double* xc_e = sd.e == 0 ? s->cap[kc].p : s->tag[kc].p;
    double* xc_f = sd.f == 0 ? s->cap[kc].p : s->tag[kc].p;
    {
      // back-substitution + the E side of the step: candidate poses, norms, model-cost terms and the
      // candidate's prep records all come out of this one launch
      BacksubArgs b;
      b.sa = schur_args; b.uF = s->uF.p; b.cam_row = cam_row; b.nk = nk; b.e_order = s->order(sd.e);
      b.d_e = s->d_pose[sd.e].p; b.seg_part = s->seg_cross.p; b.ticket = s->tickets.p + 4; b.out5 = sc + 4;
      b.x_e = x_e; b.x_cand = xc_e; b.e_is_capture = sd.e == 0; b.tag_size = o.tag_size;
      // (the candidate's prep records stay with prep_poses_kernel: one lane in eight would run the sincos /
      // Rodrigues chain here, measured +40 us on this kernel against 19 us for the separate launch)
      b.cap_pre_c = nullptr; b.cap_rt_c = nullptr; b.tag_pre_c = nullptr; b.tag_cor_c = nullptr;
      LAUNCH("backsub", 292.0 * s->n_blk + (400.0 + 96.0) * sd.n_e,
             backsub_kernel<<<cdiv((long long)sd.n_e * kBsGroup, 128), 128, 0, s->stream>>>(b));
    }
    {
      ApplyArgs ap;
      ap.uF_cam = s->uF.p + cam_row;
      ap.sc = sc; ap.nk = nk;
      ap.blocks_all = f_blocks_all; ap.constant = const_f;
      ap.n_pose = sd.n_f; ap.seg_off = s->s_off[sd.f].p; ap.seg_end = s->seg_end(sd.f); ap.x = x_f; ap.step = s->uF.p; ap.negate = 1;
      ap.rec = HF; ap.recx = HFx;
      ap.delta = s->d_pose[sd.f].p; ap.x_cand = xc_f; ap.warp_out = s->warp_norm[sd.f].p;
      ap.ticket = s->tickets.p + 6; ap.out = sc + 9;
      ap.cam = s->cam[k].p; ap.cam_c = s->cam[kc].p; ap.d_cam = s->d_cam.p;  // this launch also steps the intrinsics
      ap.count_norms = (s->rank == 0) ? 1 : 0;
      ap.pose_is_capture = sd.f == 0; ap.tag_size = o.tag_size;
      ap.cap_pre_c = nullptr; ap.cap_rt_c = nullptr; ap.tag_pre_c = nullptr; ap.tag_cor_c = nullptr;
      LAUNCH("apply_step", (144.0 + 8.0 * NV) * sd.n_f,
             apply_step_kernel<<<cdiv(sd.n_f, 128), 128, 0, s->stream>>>(ap));
    }
    cudaEventRecord(s->ev[1], s->stream);
    // ---------------- candidate point: cost at x + delta (residuals only)
    launch_prep(s, kc);
    {
      CandArgs c;
      c.n_blk = s->n_blk; c.plane = s->plane;
      c.own_idx = s->s_own[sd.e].p; c.oth_idx = s->s_oth[sd.e].p; c.obs = s->s_obs[sd.e].p;
      c.cap_pre_c = s->cap_pre[kc].p; c.tag_cor_c = s->tag_cor[kc].p; c.cam_c = s->cam[kc].p;
      c.warp_out = s->warp_cand.p; c.ticket = s->tickets.p + 7; c.out = sc + 5;
      const int grid = cdiv(s->plane, 256);
      if (dist) {
        if (sd.e == 0) LAUNCH("candidate", 72.0 * s->n_blk, candidate_kernel<0, 1><<<grid, 256, 0, s->stream>>>(c));
        else LAUNCH("candidate", 72.0 * s->n_blk, candidate_kernel<1, 1><<<grid, 256, 0, s->stream>>>(c));
      } else {
        if (sd.e == 0) LAUNCH("candidate", 72.0 * s->n_blk, candidate_kernel<0, 0><<<grid, 256, 0, s->stream>>>(c));
        else LAUNCH("candidate", 72.0 * s->n_blk, candidate_kernel<1, 0><<<grid, 256, 0, s->stream>>>(c));
      }
    }
    if (s->world > 1) {
      // small allreduce: sums [model, cand_r2, step2_e, xnorm2_e, step2_f, xnorm2_f] and
      // max of gmax_e through per-rank slots
      rc = small_allreduce(s, sc);
      if (rc) return rc;
    }
    cudaEventRecord(s->ev[2], s->stream);
    CU(cudaMemcpyAsync(s->h_sc, sc, kNumScalars * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    {
      float a = 0, b = 0;
      cudaEventElapsedTime(&a, s->ev[0], s->ev[1]);
      cudaEventElapsedTime(&b, s->ev[1], s->ev[2]);
      lin_ms += a; eval_ms += b;
    }
    const double* h = s->h_sc;
    if (lin == ARSLAM_LINSOLVE_PCG) summary->linear_solver_iterations += (long long)h[18];
    if (fresh_linearisation) {
      x_cost = 0.5 * h[2];
      grad_max = std::max(h[16], h[17]);
      if (!s->cam_const) {
        grad_max = std::max(grad_max, std::fabs(h[1]));
        if (dist) grad_max = std::max(grad_max, std::max(std::fabs(h[29]), std::fabs(h[30])));
      }
      if (iteration == 0) { summary->initial_cost = x_cost; log_iter(0, x_cost, 0.0, 0.0, 0.0, 1, 1); }
    }
    fresh_linearisation = false;
    // the data-dependent test of FinalizeIteration...: the speculative work above is discarded
    if (last_successful && grad_max <= o.gradient_tolerance) {
      termination = ARSLAM_CONVERGENCE; reason = ARSLAM_REASON_GRADIENT; break; }
    if (stop_after_readback) { termination = ARSLAM_NO_CONVERGENCE; reason = ARSLAM_REASON_MAX_ITERATIONS; break; }
    ++iteration;
    last_successful = false;
    const double f_cur = h[15];
    const double l1_cur = dist ? h[36] : s->h_cam[1], l2_cur = dist ? h[37] : s->h_cam[2];
    // Ceres' x is the reduced parameter vector: constant blocks are not part of it
    x_norm = std::sqrt(h[7] + h[10] + (s->cam_const ? 0.0 : f_cur * f_cur + l1_cur * l1_cur + l2_cur * l2_cur));
    const double step_norm = std::sqrt(h[6] + h[9] + h[13] * h[13] + (dist ? h[32] * h[32] + h[33] * h[33] : 0.0));
    // model_cost_change = -(J d).(r + J d / 2) = -(g.d + d^T H d / 2), assembled from the block pieces
    const double d_focal = -h[13];
    double model_sum = h[4] + h[8] + h[11] + h[1] * d_focal + 0.5 * h[0] * d_focal * d_focal;
    if (dist) {  // intrinsics block of the radial model
      const double d1 = -h[32], d2 = -h[33];
      model_sum += h[29] * d1 + h[30] * d2 + h[24] * d_focal * d1 + h[25] * d_focal * d2 +
                   0.5 * h[26] * d1 * d1 + h[27] * d1 * d2 + 0.5 * h[28] * d2 * d2;
    }
    const double model_cost_change = -model_sum;
    const double cand_cost_raw = 0.5 * h[5];
    const bool lin_ok = (h[12] == 0.0) && std::isfinite(step_norm) && std::isfinite(model_sum);
    if (!lin_ok || !(model_cost_change > 0.0)) {
      // HandleInvalidStep
      summary->num_unsuccessful_steps++;
      if (++invalid >= o.max_num_consecutive_invalid_steps) {
        termination = ARSLAM_FAILURE; reason = ARSLAM_REASON_INVALID_STEPS; break; }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      log_iter(iteration, x_cost, 0.0, 0.0, 0.0, 0, 0);
      continue;
    }
    invalid = 0;
    summary->num_cost_evals++;
    const double cand_cost = std::isfinite(cand_cost_raw) ? cand_cost_raw : std::numeric_limits<double>::max();
    const double cost_change = x_cost - cand_cost;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      termination = ARSLAM_CONVERGENCE; reason = ARSLAM_REASON_PARAMETER;
      log_iter(iteration, x_cost, cost_change, step_norm, 0.0, 1, 0);
      break;
    }
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) {
      termination = ARSLAM_CONVERGENCE; reason = ARSLAM_REASON_FUNCTION;
      log_iter(iteration, x_cost, cost_change, step_norm, 0.0, 1, 0);
      break;
    }
    const double rho = cost_change / model_cost_change;
    if (rho > o.min_relative_decrease) {
      // HandleSuccessfulStep: the candidate becomes the current point
      s->cur = kc;
      if (pending_acc) {  // timing of the previous accumulate is resolved now (the stream is idle)
        float a_ms = 0; cudaEventElapsedTime(&a_ms, s->ev[3], s->ev[4]); eval_ms += a_ms;
      }
      cudaEventRecord(s->ev[3], s->stream);
      launch_accumulate(s, sd, kc, HF, HFx, sc_head);
      cudaEventRecord(s->ev[4], s->stream);
      pending_acc = true;
      summary->num_jacobian_evals++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      last_successful = true;
      fresh_linearisation = true;
      summary->num_successful_steps++;
      x_cost = cand_cost;  // re-evaluated value arrives with the next readback
      log_iter(iteration, cand_cost, cost_change, step_norm, rho, 1, 1);
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      summary->num_unsuccessful_steps++;
      log_iter(iteration, cand_cost, cost_change, step_norm, rho, 1, 0);
    }
  }
  // ---- results: the poses stay on the device (arslam_get_params); camera and final cost come back
  {
    const int k = s->cur;
    CU(cudaMemcpyAsync(s->h_sc + 32, s->cam[k].p, 3 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(s->h_sc + 48, sc_head, 4 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    s->h_cam[0] = s->h_sc[32];
    if (dist) { s->h_cam[1] = s->h_sc[33]; s->h_cam[2] = s->h_sc[34]; }
    if (pending_acc) { float a_ms = 0; cudaEventElapsedTime(&a_ms, s->ev[3], s->ev[4]); eval_ms += a_ms; }
    // a successful last step was re-linearised but not read back yet: exact cost of the final point
    if (fresh_linearisation && iteration > 0 && s->world == 1) x_cost = 0.5 * s->h_sc[50];
  }
  s->prof.resolve();
  summary->iterations = iteration;
  summary->termination = termination;
  summary->reason = reason;
  summary->final_cost = x_cost;
  summary->final_radius = radius;
  summary->gradient_max_norm = grad_max;
  summary->gpu_launches = s->launches;
  summary->eval_ms = eval_ms;
  summary->linsolve_ms = lin_ms;
  summary->total_ms = wall_ms() - t_start;
  return ARSLAM_OK;
}

// ------------------------------------------------- normal equations (parity) ---
int arslam_get_normal_equations(arslam_solver* s, int32_t* eliminated_side, int32_t* blk_cap, int32_t* blk_tag,
                                double* W36, double* H_cap, double* H_tag, double* camera4) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (!s->have_problem || !s->have_params) return s->fail(ARSLAM_ERR_INVALID, "get_normal_equations needs set_problem and set_params");
  if (s->opt.num_intrinsics != 1) return s->fail(ARSLAM_ERR_UNSUPPORTED, "get_normal_equations: focal-only model");
  if (s->world > 1) return s->fail(ARSLAM_ERR_UNSUPPORTED, "get_normal_equations: single-GPU handles only");
  CU(cudaSetDevice(s->device));
  s->prof.clear();
  Sides sd;
  int elim = s->opt.elimination;
  if (elim == ARSLAM_ELIM_AUTO) elim = s->n_cap >= s->n_tag ? ARSLAM_ELIM_CAPTURES : ARSLAM_ELIM_TAGS;
  sd.e = elim == ARSLAM_ELIM_CAPTURES ? 0 : 1;
  sd.f = 1 - sd.e;
  sd.n_e = sd.e == 0 ? s->n_cap : s->n_tag;
  sd.n_f = sd.f == 0 ? s->n_cap : s->n_tag;
  if (eliminated_side) *eliminated_side = elim;
  CU(s->eval_out.ensure((size_t)sd.n_f * NV + 16));
  double* HF = s->eval_out.p;
  double* head = HF + (size_t)sd.n_f * NV;
  launch_prep(s, s->cur);
  launch_accumulate(s, sd, s->cur, HF, nullptr, head);
  CU(cudaGetLastError());
  const int nb = s->n_blk;
  const size_t ps = s->plane;
  // W leaves the device as it is stored (36 planes in E-sorted order) and is turned into one 6 x 6
  // row-major block per residual block here
  std::vector<double> planes;
  if (W36) {
    planes.resize((size_t)36 * ps);
    CU(cudaMemcpyAsync(planes.data(), s->W.p, sizeof(double) * 36 * ps, cudaMemcpyDeviceToHost, s->stream));
  }
  int32_t* out_own = sd.e == 0 ? blk_cap : blk_tag;
  int32_t* out_oth = sd.e == 0 ? blk_tag : blk_cap;
  if (out_own) CU(cudaMemcpyAsync(out_own, s->s_own[sd.e].p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s->stream));
  if (out_oth) CU(cudaMemcpyAsync(out_oth, s->s_oth[sd.e].p, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s->stream));
  double* out_e = sd.e == 0 ? H_cap : H_tag;
  double* out_f = sd.e == 0 ? H_tag : H_cap;
  if (out_e) CU(cudaMemcpyAsync(out_e, s->H[sd.e].p, sizeof(double) * NV * sd.n_e, cudaMemcpyDeviceToHost, s->stream));
  if (out_f) CU(cudaMemcpyAsync(out_f, HF, sizeof(double) * NV * sd.n_f, cudaMemcpyDeviceToHost, s->stream));
  if (camera4) CU(cudaMemcpyAsync(camera4, head, sizeof(double) * 4, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->prof.resolve();
  if (W36)
    for (int b = 0; b < nb; ++b)
      for (int q = 0; q < 36; ++q) W36[(size_t)36 * b + q] = planes[(size_t)q * ps + b];
  if (out_own && s->cap_lo && sd.e == 0)
    for (int b = 0; b < nb; ++b) out_own[b] += s->cap_lo;
  return ARSLAM_OK;
}

// -------------------------------------------------------------- localise ---
int arslam_localize_batch(arslam_solver* s, int64_t n_loc, const int32_t* blk_offsets, const int32_t* tag_idx,
                          const double* rect8, const int32_t* seed_block, int64_t n_tag, const double* camera3,
                          const double* tag_pose6, double* cap_pose6, int32_t* iterations, double* final_cost,
                          int32_t* termination) {
  if (!s) return ARSLAM_ERR_INVALID;
  if (n_loc <= 0 || !blk_offsets || !tag_idx || !rect8 || !seed_block || n_tag <= 0 || !camera3 || !tag_pose6 || !cap_pose6)
    return s->fail(ARSLAM_ERR_INVALID, "localize_batch: null pointer or empty batch");
  const int64_t nb = blk_offsets[n_loc];
  if (blk_offsets[0] != 0 || nb < 0 || nb > (1LL << 28)) return s->fail(ARSLAM_ERR_INVALID, "localize_batch: bad block offsets");
  CU(cudaSetDevice(s->device));
  s->prof.clear();
  s->launches = 0;
  DevBuf<int32_t>&d_off = s->l_off, &d_tag = s->l_tag, &d_seed = s->l_seed, &d_it = s->l_it, &d_term = s->l_term;
  DevBuf<double>&d_obs = s->l_obs, &d_tagpose = s->l_tagpose, &d_tagpre = s->l_tagpre, &d_pose = s->l_pose, &d_cost = s->l_cost;
  CU(d_off.ensure(n_loc + 1)); CU(d_tag.ensure(nb)); CU(d_seed.ensure(n_loc)); CU(d_it.ensure(n_loc)); CU(d_term.ensure(n_loc));
  CU(d_obs.ensure((size_t)8 * nb)); CU(d_tagpose.ensure((size_t)6 * n_tag)); CU(d_tagpre.ensure((size_t)kTagPre * n_tag));
  CU(d_pose.ensure((size_t)6 * n_loc)); CU(d_cost.ensure(n_loc)); CU(s->l_invalid.ensure(1));
  // chunk streams: the upload of chunk i + 1 runs under the kernel of chunk i, the download of chunk i - 1 under both
  constexpr int kLocStreams = 3;
  if (!s->loc_stream[0]) {
    for (int i = 0; i < kLocStreams; ++i) {
      CU(cudaStreamCreateWithFlags(&s->loc_stream[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&s->loc_done[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&s->loc_ready, cudaEventDisableTiming));
  }
  CU(cudaMemsetAsync(s->l_invalid.p, 0, sizeof(int), s->stream));
  CU(cudaMemcpyAsync(d_tagpose.p, tag_pose6, sizeof(double) * 6 * n_tag, cudaMemcpyHostToDevice, s->stream));
  LAUNCH("prep_poses", 8.0 * (6 + kTagPre) * n_tag,
         prep_poses_kernel<<<cdiv(n_tag, 128), 128, 0, s->stream>>>(0, nullptr, nullptr, (int)n_tag, d_tagpose.p, s->opt.tag_size, d_tagpre.p, 0, nullptr, nullptr));
  CU(cudaEventRecord(s->loc_ready, s->stream));
  LocArgs a;
  a.tag_idx = d_tag.p; a.obs = reinterpret_cast<const double2*>(d_obs.p);
  a.tag_pose = d_tagpose.p; a.tag_pre = d_tagpre.p;
  a.cam[0] = camera3[0]; a.cam[1] = camera3[1]; a.cam[2] = camera3[2];
  const arslam_options& o = s->opt;
  a.o.max_num_iterations = o.max_num_iterations; a.o.max_invalid = o.max_num_consecutive_invalid_steps;
  a.o.jacobi_scaling = o.jacobi_scaling; a.o.initial_radius = o.initial_trust_region_radius;
  a.o.max_radius = o.max_trust_region_radius; a.o.min_radius = o.min_trust_region_radius;
  a.o.min_relative_decrease = o.min_relative_decrease; a.o.min_diag = o.min_lm_diagonal; a.o.max_diag = o.max_lm_diagonal;
  a.o.function_tolerance = o.function_tolerance; a.o.gradient_tolerance = o.gradient_tolerance;
  a.o.parameter_tolerance = o.parameter_tolerance; a.o.tag_size = o.tag_size;
  a.invalid = s->l_invalid.p;
  // ~128 k captures per chunk (64 MB of observations): large enough for full-rate DMA and a full grid
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_loc, s->tune_loc_chunk > 0 ? s->tune_loc_chunk : (1 << 17)));
  const int n_chunks = (int)((n_loc + chunk - 1) / chunk);
  for (int c = 0; c < n_chunks; ++c) {
    cudaStream_t st = s->loc_stream[c % kLocStreams];
    const int64_t c0 = c * chunk, c1 = std::min<int64_t>(n_loc, c0 + chunk), nc = c1 - c0;
    const int64_t b0 = blk_offsets[c0], b1 = blk_offsets[c1];
    if (b0 < 0 || b1 < b0 || b1 > nb) return s->fail(ARSLAM_ERR_INVALID, "localize_batch: block offsets are not monotone");
    if (c < kLocStreams) CU(cudaStreamWaitEvent(st, s->loc_ready, 0));
    CU(cudaMemcpyAsync(d_off.p + c0, blk_offsets + c0, sizeof(int32_t) * (nc + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_seed.p + c0, seed_block + c0, sizeof(int32_t) * nc, cudaMemcpyHostToDevice, st));
    if (b1 > b0) {
      CU(cudaMemcpyAsync(d_tag.p + b0, tag_idx + b0, sizeof(int32_t) * (b1 - b0), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(d_obs.p + 8 * b0, rect8 + 8 * b0, sizeof(double) * 8 * (b1 - b0), cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(d_pose.p + 6 * c0, cap_pose6 + 6 * c0, sizeof(double) * 6 * nc, cudaMemcpyHostToDevice, st));
    loc_validate_kernel<<<cdiv(nc, 256), 256, 0, st>>>((int)nc, d_off.p + c0, d_tag.p, d_seed.p + c0, (int)n_tag, (int)b0, (int)b1, s->l_invalid.p);
    a.n_loc = (int)nc; a.blk_off = d_off.p + c0; a.seed_block = d_seed.p + c0;
    a.pose = d_pose.p + 6 * c0; a.iterations = d_it.p + c0; a.final_cost = d_cost.p + c0; a.termination = d_term.p + c0;
    {
      Profiler::Rec r{0, nullptr, nullptr};
      if (s->prof.on) {
        r = Profiler::Rec{s->prof.id_of("localize", 17.0 * 4 * nb + 116.0 * n_loc), s->prof.ev(), s->prof.ev()};
        cudaEventRecord(r.a, st);
      }
      if (s->opt.num_intrinsics == 3) localize_kernel<1><<<cdiv(nc * kLocGroup, 128), 128, 0, st>>>(a);
      else localize_kernel<0><<<cdiv(nc * kLocGroup, 128), 128, 0, st>>>(a);
      if (s->prof.on) {
        cudaEventRecord(r.b, st);
        s->prof.recs.push_back(r);
      }
      s->launches += 2;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(cap_pose6 + 6 * c0, d_pose.p + 6 * c0, sizeof(double) * 6 * nc, cudaMemcpyDeviceToHost, st));
    if (iterations) CU(cudaMemcpyAsync(iterations + c0, d_it.p + c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, st));
    if (final_cost) CU(cudaMemcpyAsync(final_cost + c0, d_cost.p + c0, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
    if (termination) CU(cudaMemcpyAsync(termination + c0, d_term.p + c0, sizeof(int32_t) * nc, cudaMemcpyDeviceToHost, st));
  }
  // join the chunk streams back into the handle's stream, then wait for it
  for (int i = 0; i < std::min(kLocStreams, n_chunks); ++i) {
    CU(cudaEventRecord(s->loc_done[i], s->loc_stream[i]));
    CU(cudaStreamWaitEvent(s->stream, s->loc_done[i], 0));
  }
  int h_invalid = 0;
  CU(cudaMemcpyAsync(&h_invalid, s->l_invalid.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  s->prof.resolve();
  if (h_invalid) return s->fail(ARSLAM_ERR_INVALID, "localize_batch: bad block offsets, seed block or tag index (checked on the device; no result was written)");
  return ARSLAM_OK;
}

// ------------------------------------------------------------------ comm ----
int arslam_comm_unique_id(void* id128) {
  if (!id128) return ARSLAM_ERR_INVALID;
  std::string err;
  if (!g_nccl.load(err)) { g_create_error = err; return ARSLAM_ERR_NCCL; }
  NcclId id;
  std::memset(&id, 0, sizeof(id));
  if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return ARSLAM_ERR_NCCL; }
  std::memcpy(id128, &id, 128);
  return ARSLAM_OK;
}

int arslam_comm_init(arslam_solver* s, int rank, int world_size, const void* id128) {
  if (!s || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return ARSLAM_ERR_INVALID;
  std::string err;
  if (!g_nccl.load(err)) return s->fail(ARSLAM_ERR_NCCL, "%s", err.c_str());
  CU(cudaSetDevice(s->device));
  NcclId id;
  std::memcpy(&id, id128, 128);
  // ncclCommInitRank takes the id BY VALUE (128-byte struct, passed in memory by the SysV ABI)
  typedef int (*init_fn)(void**, int, NcclId, int);
  init_fn init = reinterpret_cast<init_fn>(reinterpret_cast<void*>(g_nccl.CommInitRank));
  int rc = init(&s->comm, world_size, id, rank);
  if (rc != 0) return s->fail(ARSLAM_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  s->rank = rank;
  s->have_problem = false;  // blocks must be (re)declared after the communicator exists: ranks re-index their captures
  s->have_params = false;
  s->world = world_size;
  return ARSLAM_OK;
}

}  // extern "C"
