// ar_oracle.hpp -- CPU ORACLE (test infrastructure, NOT the product).
//
// A dependency-free restatement of the reference's optimisation hot path so
// that the CUDA solver has something to be checked against.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may link or call this.  The product (ar_slam_b200/csrc) never does.
//
// What is restated and from where (paths relative to /root/reference):
//   * the cost model             ar_slam/src/ar_slam_util.cpp:131-216
//   * the seed heuristics        ar_slam/src/ar_slam_util.cpp:41-128
//   * tag size / corner order    ar_slam/include/ar_slam/ar_slam_util.hpp:319-351
//   * solver options             ar_slam/src/ar_slam_util.cpp:1001-1018
// The arithmetic underneath those call sites lives in Ceres Solver 2.0.0
// (libceres-dev of Ubuntu 22.04; not vendored, not installed here).  Its
// published algorithms are restated here: forward-mode Jets (jet.h), the
// angle-axis / quaternion helpers (rotation.h), trust-region LM
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc), the greedy
// independent-set Schur ordering (reorder_program.cc) and the dense Schur
// complement solver (schur_eliminator_impl.h, schur_complement_solver.cc).
//
// PARITY STATUS: the reference has no numeric test of this path and real
// Ceres cannot be built here, so this oracle is pinned by (i) a 50-digit
// mpmath known-answer vector (tests/golden/kat_projection.json), (ii)
// complex-step / finite-difference Jacobian checks, (iii) scipy
// least_squares on the frozen demo detections.  Against real Ceres the
// parity is UNPINNED.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------- Jet ------
// Forward-mode dual number, same operator definitions as Ceres' jet.h.
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // jet.h: one reciprocal, then (f.v - f.a/g.a * g.v) * (1/g.a)
  const double g_a_inverse = 1.0 / g.a;
  const double f_a_by_g_a = f.a * g_a_inverse;
  Jet<N> h; h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) {
  Jet<N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  const double minus_s_g_a_inverse2 = -s / (g.a * g.a);
  Jet<N> h; h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2; return h; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline Jet<N> jsqrt(const Jet<N>& f) {
  const double tmp = std::sqrt(f.a); const double two_a_inverse = 1.0 / (2.0 * tmp);
  Jet<N> h; h.a = tmp; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) {
  const double c = std::cos(f.a);
  Jet<N> h; h.a = std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> jcos(const Jet<N>& f) {
  const double ms = -std::sin(f.a);
  Jet<N> h; h.a = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = ms * f.v[i]; return h; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jcos(double x) { return std::cos(x); }
template <typename T> inline T make_const(double s) { return T(s); }

// --------------------------------------------------- rotation helpers ------
// Ceres rotation.h AngleAxisRotatePoint: Rodrigues away from zero, the first
// order form p + w x p when theta^2 <= DBL_EPSILON.
template <typename T>
inline void angle_axis_rotate_point(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > T(DBL_EPSILON)) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1],
                      w[2] * pt[0] - w[0] * pt[2],
                      w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    out[0] = pt[0] * costheta + wxp[0] * sintheta + w[0] * tmp;
    out[1] = pt[1] * costheta + wxp[1] * sintheta + w[1] * tmp;
    out[2] = pt[2] * costheta + wxp[2] * sintheta + w[2] * tmp;
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1],
                      aa[2] * pt[0] - aa[0] * pt[2],
                      aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + wxp[0];
    out[1] = pt[1] + wxp[1];
    out[2] = pt[2] + wxp[2];
  }
}

// rotation.h AngleAxisToQuaternion / QuaternionProduct / QuaternionToAngleAxis
// (doubles only; the reference uses them for seeds, ar_slam_util.cpp:41-50).
inline void angle_axis_to_quaternion(const double aa[3], double q[4]) {
  const double t2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (t2 > 0.0) {
    const double theta = std::sqrt(t2);
    const double half = theta * 0.5;
    const double k = std::sin(half) / theta;
    q[0] = std::cos(half); q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
  } else {
    q[0] = 1.0; q[1] = aa[0] * 0.5; q[2] = aa[1] * 0.5; q[3] = aa[2] * 0.5;
  }
}
inline void quaternion_product(const double z[4], const double w[4], double zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}
inline void quaternion_to_angle_axis(const double q[4], double aa[3]) {
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = std::sqrt(s2);
    const double c = q[0];
    const double two_theta = 2.0 * ((c < 0.0) ? std::atan2(-s, -c) : std::atan2(s, c));
    const double k = two_theta / s;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
  } else {
    aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
  }
}
// ar_slam_util.cpp:41-50
inline void compose_axis_angle(const double r1[3], const double r2[3], double out[3]) {
  double q1[4], q2[4], q3[4];
  angle_axis_to_quaternion(r1, q1);
  angle_axis_to_quaternion(r2, q2);
  quaternion_product(q1, q2, q3);
  quaternion_to_angle_axis(q3, out);
}

// ------------------------------------------------------------ model --------
// Corner order TL,TR,BR,BL with +y down (ar_slam_util.hpp:335-345).
static const double kCornerDir[4][2] = {{-1, -1}, {+1, -1}, {+1, +1}, {-1, +1}};

// model 0: the reference's live model, focal length only
//          (ar_slam_util.cpp:157-162).
// model 1: the radial model sketched in the TODO comment at
//          ar_slam_util.cpp:164-171 with l1 = camera[1], l2 = camera[2]
//          (BASELINE config 5 extension; no Ceres parity target).
template <typename T>
inline void project_corner(const T* camera, const T* inv_cap_pose, const T* ar_pose,
                           unsigned idx, double tag_size, int model, T* projected) {
  T corner[3] = {T(0.5 * tag_size * kCornerDir[idx][0]),
                 T(0.5 * tag_size * kCornerDir[idx][1]), T(0.0)};
  T world[3];
  angle_axis_rotate_point(ar_pose + 3, corner, world);
  world[0] = world[0] + ar_pose[0];
  world[1] = world[1] + ar_pose[1];
  world[2] = world[2] + ar_pose[2];
  // capture pose is an inverse pose, translation applied BEFORE rotation
  world[0] = world[0] + inv_cap_pose[0];
  world[1] = world[1] + inv_cap_pose[1];
  world[2] = world[2] + inv_cap_pose[2];
  T cam[3];
  angle_axis_rotate_point(inv_cap_pose + 3, world, cam);
  const T xp = cam[0] / cam[2];
  const T yp = cam[1] / cam[2];
  const T& focal = camera[0];
  if (model == 0) {
    projected[0] = focal * xp;
    projected[1] = focal * yp;
  } else {
    const T r2 = xp * xp + yp * yp;
    const T distortion = r2 * (camera[1] + camera[2] * r2) + 1.0;
    projected[0] = focal * distortion * xp;
    projected[1] = focal * distortion * yp;
  }
}

// 8 residuals of one (capture, tag) block, ordered x0,y0,x1,y1,...
// (ar_slam_util.cpp:198-211).  rect = x0,y0,...,x3,y3 centred pixels.
template <typename T>
inline void block_residuals(const double rect[8], const T* camera, const T* cap, const T* tag,
                            double tag_size, int model, T* residuals) {
  for (unsigned idx = 0; idx < 4; ++idx) {
    T p[2];
    project_corner(camera, cap, tag, idx, tag_size, model, p);
    residuals[2 * idx + 0] = p[0] - rect[2 * idx + 0];
    residuals[2 * idx + 1] = p[1] - rect[2 * idx + 1];
  }
}

// hpp:348-351
inline double normalize_angle(double a) {
  return std::fmod(std::fmod(a, 2 * M_PI) + 3 * M_PI, 2 * M_PI) - M_PI;
}

// ar_slam_util.cpp:52-88: depth from the longest edge, centroid, in-plane yaw.
inline void calc_init_values(const double rect[8], double focal, double tag_size,
                             double* lx, double* ly, double* lz, double* yaw) {
  double max_d2 = 0.0, ax = 0.0, ay = 0.0;
  for (unsigned i = 0; i < 4; ++i) {
    const unsigned j = (i + 1) & 3;
    const double dx = rect[2 * i] - rect[2 * j], dy = rect[2 * i + 1] - rect[2 * j + 1];
    const double d2 = std::pow(dx, 2) + std::pow(dy, 2);
    max_d2 = std::max(d2, max_d2);
    ax += rect[2 * i];
    ay += rect[2 * i + 1];
  }
  ax *= 0.25;
  ay *= 0.25;
  double avg = 0.0;
  for (unsigned i = 0; i < 4; ++i) {
    const double expected = std::atan2(kCornerDir[i][1], kCornerDir[i][0]);
    const double actual = std::atan2(rect[2 * i + 1] - ay, rect[2 * i] - ax);
    const double delta = normalize_angle(actual - expected);
    avg += normalize_angle(delta - avg) / (i + 1);
  }
  *lz = focal * tag_size / std::sqrt(max_d2);
  *lx = ax * (*lz) / focal;
  *ly = ay * (*lz) / focal;
  *yaw = avg;
}

// ar_slam_util.cpp:91-108
inline void init_capture_pose(const double rect[8], const double* camera, const double* ar_pose,
                              double tag_size, double* inv_cap_pose) {
  double lx, ly, lz, yaw;
  calc_init_values(rect, camera[0], tag_size, &lx, &ly, &lz, &yaw);
  const double local_position[3] = {lx, ly, lz};
  const double local_rot[3] = {0.0, 0.0, yaw};
  const double inv_ar_rot[3] = {-ar_pose[3], -ar_pose[4], -ar_pose[5]};
  compose_axis_angle(local_rot, inv_ar_rot, inv_cap_pose + 3);
  const double cap_rotation[3] = {-inv_cap_pose[3], -inv_cap_pose[4], -inv_cap_pose[5]};
  double t[3];
  angle_axis_rotate_point(cap_rotation, local_position, t);
  inv_cap_pose[0] = t[0] - ar_pose[0];
  inv_cap_pose[1] = t[1] - ar_pose[1];
  inv_cap_pose[2] = t[2] - ar_pose[2];
}

// ar_slam_util.cpp:111-128
inline void init_ar_pose(const double rect[8], const double* camera, const double* inv_cap_pose,
                         double tag_size, double* ar_pose) {
  double lx, ly, lz, yaw;
  calc_init_values(rect, camera[0], tag_size, &lx, &ly, &lz, &yaw);
  const double local_position[3] = {lx, ly, lz};
  const double cap_rotation[3] = {-inv_cap_pose[3], -inv_cap_pose[4], -inv_cap_pose[5]};
  double t[3];
  angle_axis_rotate_point(cap_rotation, local_position, t);
  ar_pose[0] = t[0] - inv_cap_pose[0];
  ar_pose[1] = t[1] - inv_cap_pose[1];
  ar_pose[2] = t[2] - inv_cap_pose[2];
  const double local_rot[3] = {0.0, 0.0, yaw};
  compose_axis_angle(cap_rotation, local_rot, ar_pose + 3);
}

}  // namespace oracle
