// localize.cuh -- kernel (5): batched independent per-capture localisation
// against a fixed map.  Eight lanes per capture (four captures per warp) run the whole pipeline of
// ArSlamSolver::localizeOne (reference ar_slam/src/ar_slam_util.cpp:903-979)
// on the device: seed from one tag (initCapturePose, :91-108), then the
// trust-region LM that ceres::Solve would run on a problem whose tags and
// camera are constant (:965, :972) -- a single free 6-vector, so the "Schur
// complement" is one damped 6x6 Cholesky solve held in registers.
// Lanes own residual blocks (64 B of the capture's contiguous rect array each);
// J^T J / J^T r are reduced with shuffles inside the 8-lane group; every lane of
// the group then runs the same scalar LM control code.  No communication between
// captures, hence none between GPUs.
#pragma once
#include "kernels.cuh"

namespace ars {

struct LocOptions {
  int max_num_iterations, max_invalid, jacobi_scaling;
  double initial_radius, max_radius, min_radius, min_relative_decrease;
  double min_diag, max_diag, function_tolerance, gradient_tolerance, parameter_tolerance;
  double tag_size;
};

struct LocArgs {
  int n_loc;
  const int32_t* blk_off;     // [n_loc + 1]
  const int32_t* tag_idx;     // [n_blk]
  const double2* obs;         // [4 n_blk] corner (x, y), ArucoRect order
  const int32_t* seed_block;  // [n_loc] index within the capture, < 0: skip
  const double* tag_pose;     // [n_tag][6]
  const double* tag_pre;      // [n_tag][kTagPre] (world corners used)
  double cam[3];              // f, l1, l2
  LocOptions o;
  double* pose;               // [n_loc][6] out
  int32_t* iterations;        // optional
  double* final_cost;         // optional
  int32_t* termination;       // optional
  const int* invalid;         // device flag written by loc_validate_kernel: != 0 -> do nothing (indices cannot be trusted)
};

// Index validation on the device (the host used to walk 1 M offsets + 8 M tag indices serially inside
// the timed path): per capture 0 <= blocks, seed < blocks, block range inside [0, n_blk_total]; per block
// 0 <= tag < n_tag.  One thread per capture, its blocks' tags checked by the same thread.
__global__ void loc_validate_kernel(int n_loc, const int32_t* __restrict__ blk_off, const int32_t* __restrict__ tag_idx,
                                    const int32_t* __restrict__ seed_block, int n_tag, int b_lo, int b_hi,
                                    int* __restrict__ invalid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loc) return;
  const int b0 = blk_off[i], b1 = blk_off[i + 1];
  bool bad = b0 < b_lo || b1 > b_hi || b1 < b0 || seed_block[i] >= b1 - b0;
  if (!bad)
    for (int b = b0; b < b1; ++b) bad = bad || tag_idx[b] < 0 || tag_idx[b] >= n_tag;
  if (bad) atomicExch(invalid, 1);
}

// Eight lanes per capture (one lane per residual block, its four corners in sequence), four
// captures per warp: the scalar LM control code, the pose prep (sincos) and the 6x6 Cholesky
// are shared by 8 lanes instead of 32, and the Gram reductions are 3 shuffle levels inside the
// group (shuffles name only the group's lanes, so groups may diverge freely).
constexpr int kLocGroup = 8;
__device__ __forceinline__ double loc_group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = kLocGroup / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// sum r^2 at pose x over the capture's corners (all lanes of the group get the result)
template <int MODEL>
__device__ __forceinline__ double loc_cost(const LocArgs& a, int b0, int nb, const double x[6], int gl, unsigned mask) {
  double cp[kCapPre];
  prep_capture(x, cp);
  double s = 0.0;
  for (int j = gl; j < nb; j += kLocGroup) {
    const int blk = b0 + j;
    const double2* ob = a.obs + 4 * (size_t)blk;
    const double* tp = a.tag_pre + (size_t)kTagPre * a.tag_idx[blk];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 o = ob[i];
      double r[2];
      corner_residual_m<MODEL>(cp, tp + 12 * i, a.cam, o.x, o.y, r);
      s += r[0] * r[0] + r[1] * r[1];
    }
  }
  return loc_group_sum(s, mask);
}

// J^T J (upper packed 21), J^T r (6) and sum r^2 at pose x
template <int MODEL>
__device__ __forceinline__ void loc_normal_eq(const LocArgs& a, int b0, int nb, const double x[6], int gl,
                                              unsigned mask, double H[21], double g[6], double& rr) {
  double cp[kCapPre];
  prep_capture(x, cp);
#pragma unroll
  for (int i = 0; i < 21; ++i) H[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = 0.0;
  rr = 0.0;
  for (int j = gl; j < nb; j += kLocGroup) {
    const int blk = b0 + j;
    const double2* ob = a.obs + 4 * (size_t)blk;
    const double* tp = a.tag_pre + (size_t)kTagPre * a.tag_idx[blk];
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const double2 o = ob[i];
      CornerJ cj;
      double Kl[2][2];
      corner_jacobian_m<MODEL>(cp, tp + 12 * i, a.cam, o.x, o.y, cj, Kl);
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double J[6] = {cj.A[row][0], cj.A[row][1], cj.A[row][2], cj.B[row][0], cj.B[row][1], cj.B[row][2]};
#pragma unroll
        for (int p = 0; p < 6; ++p) {
#pragma unroll
          for (int q = p; q < 6; ++q) H[tri6(p, q)] += J[p] * J[q];
          g[p] += J[p] * cj.r[row];
        }
        rr += cj.r[row] * cj.r[row];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 21; ++i) H[i] = loc_group_sum(H[i], mask);
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = loc_group_sum(g[i], mask);
  rr = loc_group_sum(rr, mask);
}

template <int MODEL>
__global__ void __launch_bounds__(128, 3) localize_kernel(const LocArgs a) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int cap = gid / kLocGroup;
  const int lane = threadIdx.x & 31, gl = lane & (kLocGroup - 1);
  const unsigned mask = ((1u << kLocGroup) - 1u) << (lane & ~(kLocGroup - 1));
  if (cap >= a.n_loc) return;
  if (a.invalid && *a.invalid) return;
  const int b0 = a.blk_off[cap], nb = a.blk_off[cap + 1] - b0;
  const int seed = a.seed_block[cap];
  if (seed < 0 || nb <= 0) {
    if (gl == 0) {
      if (a.iterations) a.iterations[cap] = -1;
      if (a.final_cost) a.final_cost[cap] = 0.0;
      if (a.termination) a.termination[cap] = -1;
    }
    return;
  }
  const LocOptions& o = a.o;
  double x[6];
  {
    double rect[8];
    const double2* so = a.obs + 4 * (size_t)(b0 + seed);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 v = so[i];
      rect[2 * i] = v.x;
      rect[2 * i + 1] = v.y;
    }
    double tpose[6];
    const double* tp = a.tag_pose + 6 * (size_t)a.tag_idx[b0 + seed];
#pragma unroll
    for (int i = 0; i < 6; ++i) tpose[i] = tp[i];
    seed_capture_pose(rect, a.cam[0], tpose, o.tag_size, x);
  }

  double H[21], g[6], rr;
  double scale[6], diag[6];
  loc_normal_eq<MODEL>(a, b0, nb, x, gl, mask, H, g, rr);
  double x_cost = 0.5 * rr;
#pragma unroll
  for (int i = 0; i < 6; ++i) scale[i] = o.jacobi_scaling ? 1.0 / (1.0 + sqrt(H[tri6(i, i)])) : 1.0;
  double grad_max = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) grad_max = fmax(grad_max, fabs(x[i] - (x[i] - g[i])));
  double x_norm = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) x_norm += x[i] * x[i];
  x_norm = sqrt(x_norm);
  double radius = o.initial_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false, last_successful = true;
  int invalid = 0, iteration = 0;
  int termination = 1 /* NO_CONVERGENCE */;

  while (true) {
    if (iteration >= o.max_num_iterations) { termination = 1; break; }
    if (last_successful && grad_max <= o.gradient_tolerance) { termination = 0; break; }
    if (radius <= o.min_radius) { termination = 0; break; }
    ++iteration;
    last_successful = false;
    if (!reuse_diagonal) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
        diag[i] = fmin(fmax(H[tri6(i, i)] * scale[i] * scale[i], o.min_diag), o.max_diag);
    }
    reuse_diagonal = true;
    double L[36], y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = i; j < 6; ++j) {
        const double h = H[tri6(i, j)] * scale[i] * scale[j];
        L[i * 6 + j] = h;
        L[j * 6 + i] = h;
      }
      L[i * 6 + i] += diag[i] / radius;
      y[i] = g[i] * scale[i];
    }
    bool ok = chol6(L);
    chol6_solve(L, y);
    double delta[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      ok = ok && isfinite(y[i]);
      delta[i] = -y[i] * scale[i];
    }
    // model_cost_change = -(J d).(r + J d / 2) = -g.d - d^T H d / 2
    double gd = 0.0, dHd = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      gd += g[i] * delta[i];
#pragma unroll
      for (int j = 0; j < 6; ++j) dHd += delta[i] * delta[j] * H[i <= j ? tri6(i, j) : tri6(j, i)];
    }
    const double model_cost_change = -gd - 0.5 * dHd;
    if (!ok || !(model_cost_change > 0.0)) {
      if (++invalid >= o.max_invalid) { termination = 2; break; }
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    double xc[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xc[i] = x[i] + delta[i];
    double cand_cost = 0.5 * loc_cost<MODEL>(a, b0, nb, xc, gl, mask);
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
    double sn = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double d = x[i] - xc[i]; sn += d * d; }
    const double step_norm = sqrt(sn);
    const double cost_change = x_cost - cand_cost;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { termination = 0; break; }
    if (fabs(cost_change) <= o.function_tolerance * x_cost) { termination = 0; break; }
    const double rho = cost_change / model_cost_change;
    if (rho > o.min_relative_decrease) {
#pragma unroll
      for (int i = 0; i < 6; ++i) x[i] = xc[i];
      x_norm = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) x_norm += x[i] * x[i];
      x_norm = sqrt(x_norm);
      loc_normal_eq<MODEL>(a, b0, nb, x, gl, mask, H, g, rr);
      x_cost = 0.5 * rr;
      grad_max = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) grad_max = fmax(grad_max, fabs(x[i] - (x[i] - g[i])));
      const double t = 2.0 * rho - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(o.max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
    }
  }
  if (gl == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) a.pose[6 * (size_t)cap + i] = x[i];
    if (a.iterations) a.iterations[cap] = iteration;
    if (a.final_cost) a.final_cost[cap] = x_cost;
    if (a.termination) a.termination[cap] = termination;
  }
}

}  // namespace ars
