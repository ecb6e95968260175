// Test-only shim: compiles ar_slam_b200/csrc/model.cuh as HOST code so that the
// closed-form Jacobian formulas can be checked against the oracle on the CPU
// box (tests/test_model_host.py).  Not part of the product.
#include <cstdint>
#include "../../ar_slam_b200/csrc/model.cuh"

extern "C" {
// residuals [8], jac_cam [8][3], jac_cap [8][6], jac_tag [8][6] (Ceres layout)
void shim_eval_block(const double* cam, const double* cap, const double* tag, double tag_size,
                     const double* rect8, double* res, double* jc, double* jp, double* ja) {
  double cp[ars::kCapPre], tp[ars::kTagPre];
  ars::prep_capture(cap, cp);
  ars::prep_tag(tag, tag_size, tp);
  for (int i = 0; i < 4; ++i) {
    ars::CornerJ o;
    ars::corner_jacobian(cp, tp + 12 * i, cam[0], rect8[2 * i], rect8[2 * i + 1], o);
    double r2[2];
    ars::corner_residual(cp, tp + 12 * i, cam[0], rect8[2 * i], rect8[2 * i + 1], r2);
    for (int row = 0; row < 2; ++row) {
      const int k = 2 * i + row;
      res[k] = o.r[row];
      if (r2[row] != o.r[row]) res[k] = 1e300;  // the two code paths must agree exactly
      jc[k * 3 + 0] = o.K[row]; jc[k * 3 + 1] = 0.0; jc[k * 3 + 2] = 0.0;
      for (int j = 0; j < 3; ++j) {
        jp[k * 6 + j] = o.A[row][j];
        jp[k * 6 + 3 + j] = o.B[row][j];
        ja[k * 6 + j] = o.A[row][j];
        ja[k * 6 + 3 + j] = o.C[row][j];
      }
    }
  }
}
// same with the radial model (camera = f, l1, l2)
void shim_eval_block_dist(const double* cam, const double* cap, const double* tag, double tag_size,
                          const double* rect8, double* res, double* jc, double* jp, double* ja) {
  double cp[ars::kCapPre], tp[ars::kTagPre];
  ars::prep_capture(cap, cp);
  ars::prep_tag(tag, tag_size, tp);
  for (int i = 0; i < 4; ++i) {
    ars::CornerJ o;
    double Kl[2][2], r2[2];
    ars::corner_jacobian_m<1>(cp, tp + 12 * i, cam, rect8[2 * i], rect8[2 * i + 1], o, Kl);
    ars::corner_residual_m<1>(cp, tp + 12 * i, cam, rect8[2 * i], rect8[2 * i + 1], r2);
    for (int row = 0; row < 2; ++row) {
      const int k = 2 * i + row;
      res[k] = o.r[row];
      if (r2[row] != o.r[row]) res[k] = 1e300;
      jc[k * 3 + 0] = o.K[row]; jc[k * 3 + 1] = Kl[row][0]; jc[k * 3 + 2] = Kl[row][1];
      for (int j = 0; j < 3; ++j) {
        jp[k * 6 + j] = o.A[row][j];
        jp[k * 6 + 3 + j] = o.B[row][j];
        ja[k * 6 + j] = o.A[row][j];
        ja[k * 6 + 3 + j] = o.C[row][j];
      }
    }
  }
}
void shim_seed_capture_pose(const double* rect8, double focal, const double* tag_pose, double tag_size,
                            double* out6) {
  ars::seed_capture_pose(rect8, focal, tag_pose, tag_size, out6);
}
void shim_seed_tag_pose(const double* rect8, double focal, const double* cap_pose, double tag_size,
                        double* out6) {
  ars::seed_tag_pose(rect8, focal, cap_pose, tag_size, out6);
}
int shim_chol6_solve(const double* H36, const double* b6, double* x6) {
  double L[36];
  for (int i = 0; i < 36; ++i) L[i] = H36[i];
  const bool ok = ars::chol6(L);
  for (int i = 0; i < 6; ++i) x6[i] = b6[i];
  ars::chol6_solve(L, x6);
  return ok ? 1 : 0;
}
}
