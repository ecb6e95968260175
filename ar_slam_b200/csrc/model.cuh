// model.cuh -- closed-form reprojection model and analytic Jacobians.
//
// Replaces projectCorner<T> + ArucoReprojectionError evaluated with
// ceres::Jet<double,15> (reference ar_slam/src/ar_slam_util.cpp:131-216,
// AutoDiffCostFunction at :722/:831/:958).  Nothing here is autodiff: the
// work that only depends on a pose is hoisted into a per-pose "prep" record
// (one thread per pose), so the per-observation code is a handful of 3x3
// products.  Functions are __host__ __device__ only so that the formulas can
// be unit-tested on the CPU box against the oracle (tests/test_model_host.py);
// the product never runs them on the host.
//
// Notation, block = (capture c, tag a), corner i:
//   m_i = tag corner in the tag frame, pw_i = R(w_a) m_i + t_a
//   q   = pw_i + t_c,  p = R(w_c) q,  (u,v) = f (p.x/p.z, p.y/p.z)
//   dp  = d(u|v)/dp (row 3-vector),  a = dp^T R_c  (== d/dt_c == d/dt_a)
//   d/dw_c[k] = dp^T dR_c/dw_k q = M_k . (q x a)   with [M_k]x = R_c^T dR_c/dw_k
//   d/dw_a[k] = a^T (dR_a/dw_k m_i) = a^T Ga_i[:,k]
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define ARS_HD __host__ __device__ __forceinline__
#else
#define ARS_HD inline
#endif

namespace ars {

// corner directions TL,TR,BR,BL, +y down (ar_slam_util.hpp:340-345)
ARS_HD double corner_dx(int i) { return (i == 1 || i == 2) ? 1.0 : -1.0; }
ARS_HD double corner_dy(int i) { return (i >= 2) ? 1.0 : -1.0; }

// R(w) and dR/dw_k, differentiating the Rodrigues expression term by term in
// the same form Ceres' AngleAxisRotatePoint uses, including its
// theta^2 <= DBL_EPSILON first-order branch (R = I + [w]x, dR_k = [e_k]x).
// Returns true when the small-angle branch was taken.
ARS_HD bool rot_and_derivs(const double w[3], double R[9], double dR[3][9]) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(theta, &s, &c);
#else
    s = sin(theta);
    c = cos(theta);
#endif
    const double ti = 1.0 / theta;
    const double n[3] = {w[0] * ti, w[1] * ti, w[2] * ti};
    const double omc = 1.0 - c;
    // R = c I + s [n]x + omc n n^T
    R[0] = c + omc * n[0] * n[0];
    R[1] = -s * n[2] + omc * n[0] * n[1];
    R[2] = s * n[1] + omc * n[0] * n[2];
    R[3] = s * n[2] + omc * n[1] * n[0];
    R[4] = c + omc * n[1] * n[1];
    R[5] = -s * n[0] + omc * n[1] * n[2];
    R[6] = -s * n[1] + omc * n[2] * n[0];
    R[7] = s * n[0] + omc * n[2] * n[1];
    R[8] = c + omc * n[2] * n[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dc = -s * n[k];   // d cos(theta)
      const double ds = c * n[k];    // d sin(theta)
      const double domc = s * n[k];  // d (1 - cos)
      double dn[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) dn[i] = ((i == k ? 1.0 : 0.0) - n[i] * n[k]) * ti;
      double* D = dR[k];
      // d(c I) + d(s [n]x) + d(omc n n^T)
      D[0] = dc + domc * n[0] * n[0] + omc * (2.0 * dn[0] * n[0]);
      D[4] = dc + domc * n[1] * n[1] + omc * (2.0 * dn[1] * n[1]);
      D[8] = dc + domc * n[2] * n[2] + omc * (2.0 * dn[2] * n[2]);
      const double s01 = domc * n[0] * n[1] + omc * (dn[0] * n[1] + n[0] * dn[1]);
      const double s02 = domc * n[0] * n[2] + omc * (dn[0] * n[2] + n[0] * dn[2]);
      const double s12 = domc * n[1] * n[2] + omc * (dn[1] * n[2] + n[1] * dn[2]);
      const double a0 = ds * n[0] + s * dn[0];
      const double a1 = ds * n[1] + s * dn[1];
      const double a2 = ds * n[2] + s * dn[2];
      D[1] = -a2 + s01;
      D[2] = a1 + s02;
      D[3] = a2 + s01;
      D[5] = -a0 + s12;
      D[6] = -a1 + s02;
      D[7] = a0 + s12;
    }
    return false;
  }
  R[0] = 1.0;   R[1] = -w[2]; R[2] = w[1];
  R[3] = w[2];  R[4] = 1.0;   R[5] = -w[0];
  R[6] = -w[1]; R[7] = w[0];  R[8] = 1.0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 9; ++i) dR[k][i] = 0.0;
  dR[0][5] = -1.0; dR[0][7] = 1.0;
  dR[1][2] = 1.0;  dR[1][6] = -1.0;
  dR[2][1] = -1.0; dR[2][3] = 1.0;
  return true;
}

// Per-capture record: R (9) | M (9, M[i*3+k] = i-th component of M_k) | t (3)
// | small-angle flag (1) ; padded to 24 doubles (192 B, 16 B aligned).
constexpr int kCapPre = 24;
// Per-tag record: for corner i: pw_i (3) | Ga_i (9, row-major [j][k]) ; 48 doubles.
constexpr int kTagPre = 48;

ARS_HD void prep_capture(const double pose[6], double* out) {
  double R[9], dR[3][9];
  const bool small = rot_and_derivs(pose + 3, R, dR);
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = R[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // X = R^T dR_k, antisymmetric: M_k = vee(X)
    const double* D = dR[k];
    const double x21 = R[2] * D[1] + R[5] * D[4] + R[8] * D[7];
    const double x12 = R[1] * D[2] + R[4] * D[5] + R[7] * D[8];
    const double x02 = R[0] * D[2] + R[3] * D[5] + R[6] * D[8];
    const double x20 = R[2] * D[0] + R[5] * D[3] + R[8] * D[6];
    const double x10 = R[1] * D[0] + R[4] * D[3] + R[7] * D[6];
    const double x01 = R[0] * D[1] + R[3] * D[4] + R[6] * D[7];
    out[9 + 0 * 3 + k] = 0.5 * (x21 - x12);
    out[9 + 1 * 3 + k] = 0.5 * (x02 - x20);
    out[9 + 2 * 3 + k] = 0.5 * (x10 - x01);
  }
  out[18] = pose[0];
  out[19] = pose[1];
  out[20] = pose[2];
  out[21] = small ? 1.0 : 0.0;
  out[22] = 0.0;
  out[23] = 0.0;
}

ARS_HD void prep_tag(const double pose[6], double tag_size, double* out) {
  double R[9], dR[3][9];
  rot_and_derivs(pose + 3, R, dR);
  const double h = 0.5 * tag_size;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double mx = h * corner_dx(i), my = h * corner_dy(i);  // m_i = (mx, my, 0)
    double* o = out + 12 * i;
    o[0] = (R[0] * mx + R[1] * my) + pose[0];
    o[1] = (R[3] * mx + R[4] * my) + pose[1];
    o[2] = (R[6] * mx + R[7] * my) + pose[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double* D = dR[k];
      o[3 + 0 * 3 + k] = D[0] * mx + D[1] * my;
      o[3 + 1 * 3 + k] = D[3] * mx + D[4] * my;
      o[3 + 2 * 3 + k] = D[6] * mx + D[7] * my;
    }
  }
}

// Everything one observation corner contributes (focal-only model,
// ar_slam_util.cpp:157-162): residual r[2], d/df K[2], and the three 2x3
// Jacobian pieces A (d/dt_c == d/dt_a), B (d/dw_c), C (d/dw_a).
struct CornerJ {
  double r[2], K[2];
  double A[2][3], B[2][3], C[2][3];
};

// residual only
ARS_HD void corner_residual(const double* __restrict__ cp, const double* __restrict__ tp, double f,
                            double ox, double oy, double r[2]) {
  const double q0 = tp[0] + cp[18], q1 = tp[1] + cp[19], q2 = tp[2] + cp[20];
  const double px = cp[0] * q0 + cp[1] * q1 + cp[2] * q2;
  const double py = cp[3] * q0 + cp[4] * q1 + cp[5] * q2;
  const double pz = cp[6] * q0 + cp[7] * q1 + cp[8] * q2;
  const double iz = 1.0 / pz;
  r[0] = f * (px * iz) - ox;
  r[1] = f * (py * iz) - oy;
}

ARS_HD void corner_jacobian(const double* __restrict__ cp, const double* __restrict__ tp, double f,
                            double ox, double oy, CornerJ& o) {
  const double q[3] = {tp[0] + cp[18], tp[1] + cp[19], tp[2] + cp[20]};
  const double px = cp[0] * q[0] + cp[1] * q[1] + cp[2] * q[2];
  const double py = cp[3] * q[0] + cp[4] * q[1] + cp[5] * q[2];
  const double pz = cp[6] * q[0] + cp[7] * q[1] + cp[8] * q[2];
  const double iz = 1.0 / pz;
  const double xp = px * iz, yp = py * iz;
  o.r[0] = f * xp - ox;
  o.r[1] = f * yp - oy;
  o.K[0] = xp;
  o.K[1] = yp;
  const double fz = f * iz;
  const bool small = cp[21] != 0.0;
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    const double e = row == 0 ? xp : yp;
    // a = dp^T R with dp = fz * (row 0: (1,0,-xp), row 1: (0,1,-yp))
    double a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fz * (cp[3 * row + j] - e * cp[6 + j]);
    o.A[row][0] = a[0];
    o.A[row][1] = a[1];
    o.A[row][2] = a[2];
    if (!small) {
      // B[k] = M_k . (q x a)
      const double c0 = q[1] * a[2] - q[2] * a[1];
      const double c1 = q[2] * a[0] - q[0] * a[2];
      const double c2 = q[0] * a[1] - q[1] * a[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) o.B[row][k] = c0 * cp[9 + k] + c1 * cp[12 + k] + c2 * cp[15 + k];
    } else {
      // first-order branch: dp^T [e_k]x q = (q x dp)_k
      const double d0 = row == 0 ? fz : 0.0, d1 = row == 0 ? 0.0 : fz, d2 = -fz * e;
      o.B[row][0] = q[1] * d2 - q[2] * d1;
      o.B[row][1] = q[2] * d0 - q[0] * d2;
      o.B[row][2] = q[0] * d1 - q[1] * d0;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) o.C[row][k] = a[0] * tp[3 + k] + a[1] * tp[6 + k] + a[2] * tp[9 + k];
  }
}

// ---- radial model of the TODO at ar_slam_util.cpp:164-171 (BASELINE config 5) -------------
//   r2 = xp^2 + yp^2,  d = 1 + r2 (l1 + l2 r2),  (u, v) = f d (xp, yp),  camera = [f, l1, l2]
// All three intrinsics are live.  MODEL 0 is the reference's live focal-only model above.
template <int MODEL>
ARS_HD void corner_residual_m(const double* __restrict__ cp, const double* __restrict__ tp,
                              const double cam[3], double ox, double oy, double r[2]) {
  if (MODEL == 0) { corner_residual(cp, tp, cam[0], ox, oy, r); return; }
  const double q0 = tp[0] + cp[18], q1 = tp[1] + cp[19], q2 = tp[2] + cp[20];
  const double px = cp[0] * q0 + cp[1] * q1 + cp[2] * q2;
  const double py = cp[3] * q0 + cp[4] * q1 + cp[5] * q2;
  const double pz = cp[6] * q0 + cp[7] * q1 + cp[8] * q2;
  const double iz = 1.0 / pz;
  const double xp = px * iz, yp = py * iz;
  const double r2 = xp * xp + yp * yp;
  const double d = r2 * (cam[1] + cam[2] * r2) + 1.0;
  r[0] = cam[0] * d * xp - ox;
  r[1] = cam[0] * d * yp - oy;
}

// Kl[row][0..1] = d r / d l1, d r / d l2 (MODEL 1 only); o.K is d r / d f
template <int MODEL>
ARS_HD void corner_jacobian_m(const double* __restrict__ cp, const double* __restrict__ tp,
                              const double cam[3], double ox, double oy, CornerJ& o, double Kl[2][2]) {
  if (MODEL == 0) {
    corner_jacobian(cp, tp, cam[0], ox, oy, o);
    Kl[0][0] = Kl[0][1] = Kl[1][0] = Kl[1][1] = 0.0;
    return;
  }
  const double f = cam[0], l1 = cam[1], l2 = cam[2];
  const double q[3] = {tp[0] + cp[18], tp[1] + cp[19], tp[2] + cp[20]};
  const double px = cp[0] * q[0] + cp[1] * q[1] + cp[2] * q[2];
  const double py = cp[3] * q[0] + cp[4] * q[1] + cp[5] * q[2];
  const double pz = cp[6] * q[0] + cp[7] * q[1] + cp[8] * q[2];
  const double iz = 1.0 / pz;
  const double xp = px * iz, yp = py * iz;
  const double r2 = xp * xp + yp * yp;
  const double d = r2 * (l1 + l2 * r2) + 1.0;
  const double g2 = 2.0 * (l1 + 2.0 * l2 * r2);  // d(d)/d(xp) = g2 xp
  o.r[0] = f * d * xp - ox;
  o.r[1] = f * d * yp - oy;
  o.K[0] = d * xp;
  o.K[1] = d * yp;
  Kl[0][0] = f * r2 * xp; Kl[0][1] = f * r2 * r2 * xp;
  Kl[1][0] = f * r2 * yp; Kl[1][1] = f * r2 * r2 * yp;
  // P = d(u, v)/d(xp, yp)
  const double Pxy = f * g2 * xp * yp;
  const double P[2][2] = {{f * (d + g2 * xp * xp), Pxy}, {Pxy, f * (d + g2 * yp * yp)}};
  const bool small = cp[21] != 0.0;
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    // dp = d(u|v)/dp = P[row] . [[iz, 0, -xp iz], [0, iz, -yp iz]]
    const double dp[3] = {P[row][0] * iz, P[row][1] * iz, -(P[row][0] * xp + P[row][1] * yp) * iz};
    double a[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = dp[0] * cp[j] + dp[1] * cp[3 + j] + dp[2] * cp[6 + j];
    o.A[row][0] = a[0];
    o.A[row][1] = a[1];
    o.A[row][2] = a[2];
    if (!small) {
      const double c0 = q[1] * a[2] - q[2] * a[1];
      const double c1 = q[2] * a[0] - q[0] * a[2];
      const double c2 = q[0] * a[1] - q[1] * a[0];
#pragma unroll
      for (int k = 0; k < 3; ++k) o.B[row][k] = c0 * cp[9 + k] + c1 * cp[12 + k] + c2 * cp[15 + k];
    } else {
      o.B[row][0] = q[1] * dp[2] - q[2] * dp[1];
      o.B[row][1] = q[2] * dp[0] - q[0] * dp[2];
      o.B[row][2] = q[0] * dp[1] - q[1] * dp[0];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) o.C[row][k] = a[0] * tp[3 + k] + a[1] * tp[6 + k] + a[2] * tp[9 + k];
  }
}

// ---- seeds (calcInitValues / initCapturePose, ar_slam_util.cpp:52-108),
// needed on the device by the batched localisation kernel.
ARS_HD double normalize_angle(double a) {
  const double two_pi = 2.0 * M_PI;
  return fmod(fmod(a, two_pi) + 3.0 * M_PI, two_pi) - M_PI;
}

ARS_HD void aa_to_quat(const double a[3], double q[4]) {
  const double t2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  if (t2 > 0.0) {
    const double th = sqrt(t2), h = th * 0.5;
    const double k = sin(h) / th;
    q[0] = cos(h); q[1] = a[0] * k; q[2] = a[1] * k; q[3] = a[2] * k;
  } else {
    q[0] = 1.0; q[1] = a[0] * 0.5; q[2] = a[1] * 0.5; q[3] = a[2] * 0.5;
  }
}
ARS_HD void quat_to_aa(const double q[4], double a[3]) {
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = sqrt(s2);
    const double two_theta = 2.0 * ((q[0] < 0.0) ? atan2(-s, -q[0]) : atan2(s, q[0]));
    const double k = two_theta / s;
    a[0] = q[1] * k; a[1] = q[2] * k; a[2] = q[3] * k;
  } else {
    a[0] = q[1] * 2.0; a[1] = q[2] * 2.0; a[2] = q[3] * 2.0;
  }
}
ARS_HD void compose_aa(const double r1[3], const double r2[3], double out[3]) {
  double z[4], w[4], zw[4];
  aa_to_quat(r1, z);
  aa_to_quat(r2, w);
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
  quat_to_aa(zw, out);
}
// value-only Rodrigues rotation with Ceres' branch
ARS_HD void rotate_point(const double w[3], const double p[3], double out[3]) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    const double c = cos(theta), s = sin(theta), ti = 1.0 / theta;
    const double n[3] = {w[0] * ti, w[1] * ti, w[2] * ti};
    const double x[3] = {n[1] * p[2] - n[2] * p[1], n[2] * p[0] - n[0] * p[2], n[0] * p[1] - n[1] * p[0]};
    const double tmp = (n[0] * p[0] + n[1] * p[1] + n[2] * p[2]) * (1.0 - c);
    out[0] = p[0] * c + x[0] * s + n[0] * tmp;
    out[1] = p[1] * c + x[1] * s + n[1] * tmp;
    out[2] = p[2] * c + x[2] * s + n[2] * tmp;
  } else {
    out[0] = p[0] + (w[1] * p[2] - w[2] * p[1]);
    out[1] = p[1] + (w[2] * p[0] - w[0] * p[2]);
    out[2] = p[2] + (w[0] * p[1] - w[1] * p[0]);
  }
}
// calcInitValues (ar_slam_util.cpp:52-88): depth from the longest edge, centroid, in-plane yaw
ARS_HD void seed_local_values(const double rect[8], double focal, double tag_size, double lp[3], double* yaw) {
  double max_d2 = 0.0, ax = 0.0, ay = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const double dx = rect[2 * i] - rect[2 * j], dy = rect[2 * i + 1] - rect[2 * j + 1];
    const double d2 = dx * dx + dy * dy;
    max_d2 = fmax(d2, max_d2);
    ax += rect[2 * i];
    ay += rect[2 * i + 1];
  }
  ax *= 0.25;
  ay *= 0.25;
  double avg = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double expected = atan2(corner_dy(i), corner_dx(i));
    const double actual = atan2(rect[2 * i + 1] - ay, rect[2 * i] - ax);
    const double delta = normalize_angle(actual - expected);
    avg += normalize_angle(delta - avg) / (double)(i + 1);
  }
  const double lz = focal * tag_size / sqrt(max_d2);
  lp[0] = ax * lz / focal;
  lp[1] = ay * lz / focal;
  lp[2] = lz;
  *yaw = avg;
}
// initArPose (ar_slam_util.cpp:111-128): tag pose from one quad and the capture's pose
ARS_HD void seed_tag_pose(const double rect[8], double focal, const double cap_pose[6], double tag_size,
                          double out[6]) {
  double lp[3], yaw;
  seed_local_values(rect, focal, tag_size, lp, &yaw);
  const double cap_rot[3] = {-cap_pose[3], -cap_pose[4], -cap_pose[5]};
  double t[3];
  rotate_point(cap_rot, lp, t);
  out[0] = t[0] - cap_pose[0];
  out[1] = t[1] - cap_pose[1];
  out[2] = t[2] - cap_pose[2];
  const double local_rot[3] = {0.0, 0.0, yaw};
  compose_aa(cap_rot, local_rot, out + 3);
}
// rect: x0,y0,...,y3
ARS_HD void seed_capture_pose(const double rect[8], double focal, const double tag_pose[6],
                              double tag_size, double out[6]) {
  double max_d2 = 0.0, ax = 0.0, ay = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const double dx = rect[2 * i] - rect[2 * j], dy = rect[2 * i + 1] - rect[2 * j + 1];
    const double d2 = dx * dx + dy * dy;
    max_d2 = fmax(d2, max_d2);
    ax += rect[2 * i];
    ay += rect[2 * i + 1];
  }
  ax *= 0.25;
  ay *= 0.25;
  double avg = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double expected = atan2(corner_dy(i), corner_dx(i));
    const double actual = atan2(rect[2 * i + 1] - ay, rect[2 * i] - ax);
    const double delta = normalize_angle(actual - expected);
    avg += normalize_angle(delta - avg) / (double)(i + 1);
  }
  const double lz = focal * tag_size / sqrt(max_d2);
  const double lp[3] = {ax * lz / focal, ay * lz / focal, lz};
  const double local_rot[3] = {0.0, 0.0, avg};
  const double inv_ar_rot[3] = {-tag_pose[3], -tag_pose[4], -tag_pose[5]};
  compose_aa(local_rot, inv_ar_rot, out + 3);
  const double cap_rot[3] = {-out[3], -out[4], -out[5]};
  double t[3];
  rotate_point(cap_rot, lp, t);
  out[0] = t[0] - tag_pose[0];
  out[1] = t[1] - tag_pose[1];
  out[2] = t[2] - tag_pose[2];
}

// Cholesky factor + solve of a 6x6 SPD system held in registers
// ("batched 6x6 Cholesky in registers").  H is full row-major; the factor
// overwrites the lower triangle of the same array, with the RECIPROCALS of the
// pivots on the diagonal (the solves then need no division).  Returns false on
// a non-positive pivot (Eigen::LLT's failure condition).
// (All loops below have constant trip counts and compile-time predicates: loops whose bounds
// depend on an outer index are not reliably unrolled, and then the 6x6 array lands in local memory.)
ARS_HD bool chol6(double H[36]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = H[j * 6 + j];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < j) d -= H[j * 6 + k] * H[j * 6 + k];
    ok = ok && (d > 0.0);
#if defined(__CUDA_ARCH__)
    const double il = rsqrt(d);
#else
    const double il = 1.0 / sqrt(d);
#endif
    H[j * 6 + j] = il;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i > j) {
        double s = H[i * 6 + j];
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k < j) s -= H[i * 6 + k] * H[j * 6 + k];
        H[i * 6 + j] = s * il;
      }
    }
  }
  return ok;
}
ARS_HD void chol6_solve(const double L[36], double b[6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < i) s -= L[i * 6 + k] * b[k];
    b[i] = s * L[i * 6 + i];
  }
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

// index of (i,j), i<=j, in a packed upper triangle of a 6x6 symmetric matrix
ARS_HD constexpr int tri6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

}  // namespace ars
