/* ar_oracle_capi.h -- C entry points of the CPU ORACLE (test infrastructure,
 * NOT the product; see ar_oracle.hpp).  Loaded through ctypes by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference. */
#ifndef AR_ORACLE_CAPI_H_
#define AR_ORACLE_CAPI_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_CONVERGENCE = 0, ORACLE_NO_CONVERGENCE = 1, ORACLE_FAILURE = 2 };
enum {
  ORACLE_REASON_GRADIENT = 1, ORACLE_REASON_PARAMETER = 2, ORACLE_REASON_FUNCTION = 3,
  ORACLE_REASON_MIN_RADIUS = 4, ORACLE_REASON_MAX_ITERATIONS = 5, ORACLE_REASON_INVALID_STEPS = 6
};

typedef struct {
  int32_t max_num_iterations;
  int32_t max_num_consecutive_invalid_steps;
  int32_t jacobi_scaling;
  int32_t num_threads;
  int32_t elimination; /* 0 Ceres' independent set, 1 tags, 2 captures, 3 none */
  int32_t _pad;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
} oracle_options;

typedef struct {
  int32_t iterations, num_successful_steps, num_unsuccessful_steps;
  int32_t termination, reason, n_e_blocks, reduced_dim, num_parameters;
  double initial_cost, final_cost, final_radius, gradient_max_norm;
  double total_seconds, jacobian_seconds, linear_solver_seconds;
} oracle_summary;

void oracle_default_options(oracle_options* o);
void oracle_project_block(const double* cam, const double* cap, const double* tag, double tag_size,
                          int model, double* uv8);
int oracle_evaluate(int n_blk, const int32_t* cap_idx, const int32_t* tag_idx, const double* obs,
                    const double* cam, const double* cap, const double* tag, double tag_size, int model,
                    int num_threads, double* cost, double* residuals, double* jac_cam, double* jac_cap,
                    double* jac_tag);
void oracle_init_capture_pose(const double* rect8, const double* cam, const double* tag_pose,
                              double tag_size, double* cap_pose_out);
void oracle_init_tag_pose(const double* rect8, const double* cam, const double* cap_pose,
                          double tag_size, double* tag_pose_out);
void oracle_compose_axis_angle(const double* r1, const double* r2, double* out);
void oracle_rotate_point(const double* aa, const double* pt, double* out);
/* iter_log: log_cap rows of 8 doubles: cost, cost_change, gradient_max_norm,
 * step_norm, relative_decrease, radius, step_is_valid, step_is_successful */
int oracle_solve(int n_cap, int n_tag, int n_blk, const int32_t* cap_idx, const int32_t* tag_idx,
                 const double* obs, double tag_size, int model, int cam_const, const uint8_t* cap_const,
                 const uint8_t* tag_const, const oracle_options* opt, double* cam, double* cap, double* tag,
                 oracle_summary* summary, double* iter_log, int log_cap);
int oracle_localize_batch(int n_loc, const int32_t* blk_offsets, const int32_t* tag_idx, const double* obs,
                          const int32_t* seed_block, int n_tag, const double* cam, const double* tag,
                          double tag_size, int model, const oracle_options* opt, int num_threads,
                          double* cap_pose, int32_t* iterations, double* final_cost, int32_t* termination);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
