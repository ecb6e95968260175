/* ar_slam_b200.h -- C-ABI of the B200-native solver for ar_slam's optimisation
 * hot path.
 *
 * The reference has no FFI: its seam is the C++ class ArSlamSolver
 * (ar_slam/include/ar_slam/ar_slam_util.hpp:367-497).  This header is what a
 * maintainer binds in place of the `ceres::Problem problem_` member
 * (ar_slam_util.hpp:473) and of the Ceres calls in ar_slam_util.cpp; each entry
 * point names the reference code it replaces.  Plain pointers and sizes only,
 * no C++ or torch types, no exceptions across the boundary: every call returns
 * ARSLAM_OK (0) or a negative error code and leaves a message in
 * arslam_last_error().
 *
 * Threading (ar_slam/src/ar_slam.cpp:87,93-98): one caller at a time per
 * handle, from any thread; every entry binds the handle's CUDA device first.
 * Handles share no mutable state.
 *
 * Conventions:
 *   camera   [f, l1, l2]                        (ar_slam_util.hpp:64-76)
 *   capture  [tx,ty,tz, wx,wy,wz]  inverse pose, p_cam = R(w)(p_world + t)
 *   tag      [tx,ty,tz, wx,wy,wz]  forward pose, p_world = R(w) c + t
 *   rect     x0,y0,x1,y1,x2,y2,x3,y3 centred pixels, TL,TR,BR,BL
 *                                                (ar_slam_util.hpp:266-293,335-345)
 */
#ifndef AR_SLAM_B200_H_
#define AR_SLAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARSLAM_ABI_VERSION 4 /* 4: marker detection (arslam_detector_*, arslam_detect_markers); 2: arslam_append_blocks, device-resident parameters; 3: arslam_set_constant,
                                arslam_get_normal_equations, arslam_set_tuning, device-resident schedule entries
                                (arslam_set_camera / set_poses / get_poses / seed_captures / seed_tags) */

enum {
  ARSLAM_OK = 0,
  ARSLAM_ERR_INVALID = -1,    /* bad argument / bad index / call order     */
  ARSLAM_ERR_CUDA = -2,       /* CUDA runtime error (message has details)  */
  ARSLAM_ERR_NO_DEVICE = -3,  /* no usable sm_100 device: there is NO CPU fallback */
  ARSLAM_ERR_NCCL = -4,
  ARSLAM_ERR_UNSUPPORTED = -5
};

/* Mirrors ceres::TerminationType as far as the reference can observe it. */
enum { ARSLAM_CONVERGENCE = 0, ARSLAM_NO_CONVERGENCE = 1, ARSLAM_FAILURE = 2 };
enum {
  ARSLAM_REASON_GRADIENT = 1, ARSLAM_REASON_PARAMETER = 2, ARSLAM_REASON_FUNCTION = 3,
  ARSLAM_REASON_MIN_RADIUS = 4, ARSLAM_REASON_MAX_ITERATIONS = 5, ARSLAM_REASON_INVALID_STEPS = 6
};

enum { ARSLAM_ELIM_AUTO = 0, ARSLAM_ELIM_TAGS = 1, ARSLAM_ELIM_CAPTURES = 2 };
enum { ARSLAM_LINSOLVE_AUTO = 0, ARSLAM_LINSOLVE_DENSE = 1, ARSLAM_LINSOLVE_PCG = 2 };

/* Solver options.  Defaults (arslam_default_options) are exactly what
 * ArSlamSolver::optimize sets (ar_slam_util.cpp:1003-1012: 50 iterations,
 * DENSE_SCHUR) plus the Ceres 2.0.0 Solver::Options defaults it leaves alone. */
typedef struct arslam_options {
  int32_t max_num_iterations;                /* 50  (ar_slam_util.cpp:1004) */
  int32_t max_num_consecutive_invalid_steps; /* 5   */
  int32_t jacobi_scaling;                    /* 1   */
  int32_t elimination;                       /* ARSLAM_ELIM_*; AUTO eliminates the larger pose set */
  int32_t linear_solver;                     /* ARSLAM_LINSOLVE_*; AUTO: dense Cholesky when the
                                                reduced system is tiny or structurally dense
                                                (>= 25 % block fill), block-sparse PCG otherwise */
  int32_t pcg_max_iterations;                /* 500 */
  int32_t num_intrinsics;                    /* 1: focal only (the reference's live model,
                                                ar_slam_util.cpp:160-162, l1 / l2 inert);
                                                3: focal + l1 + l2 of the radial TODO model
                                                (:164-171): evaluation, localisation and the LM
                                                solve with either linear solver */
  int32_t verbose;                           /* 0; 1 prints the per-iteration table like
                                                minimizer_progress_to_stdout (:1012)          */
  double initial_trust_region_radius;        /* 1e4   */
  double max_trust_region_radius;            /* 1e16  */
  double min_trust_region_radius;            /* 1e-32 */
  double min_relative_decrease;              /* 1e-3  */
  double min_lm_diagonal;                    /* 1e-6  */
  double max_lm_diagonal;                    /* 1e32  */
  double function_tolerance;                 /* 1e-6  */
  double gradient_tolerance;                 /* 1e-10 */
  double parameter_tolerance;                /* 1e-8  */
  double pcg_tolerance;                      /* 0.1: relative residual ||S y - b|| / ||b|| at which
                                                PCG stops (inexact Newton; Ceres' eta).  >= 1e-6
                                                runs the one-barrier pipelined recurrence, tighter
                                                values the classic one                          */
  double pcg_q_tolerance;                    /* 0: off.  > 0: PCG also stops when i (Q_i - Q_{i-1}) / Q_i < this, Q(x) = x'Sx - 2 b'x
                                                -- the rule Ceres' own ConjugateGradientsSolver applies for the
                                                inexact steps of its trust-region strategies (q_tolerance = eta = 0.1,
                                                r_tolerance off); pipelined PCG kernel only          */
  double tag_size;                           /* 0.0635 m (ar_slam_util.hpp:319)                */
  int64_t dense_max_dim;                     /* AUTO never picks the dense Cholesky above this
                                                reduced dimension (default 16384)              */
} arslam_options;

/* What ceres::Solver::Summary would have told the reference had it looked
 * (ar_slam_util.cpp:1013-1017 discards it). */
typedef struct arslam_summary {
  int32_t iterations;             /* LM iterations run (successful + unsuccessful + invalid) */
  int32_t num_successful_steps;   /* includes iteration 0, like Ceres */
  int32_t num_unsuccessful_steps;
  int32_t termination;            /* ARSLAM_CONVERGENCE / NO_CONVERGENCE / FAILURE */
  int32_t reason;                 /* ARSLAM_REASON_* */
  int32_t eliminated_side;        /* ARSLAM_ELIM_TAGS or ARSLAM_ELIM_CAPTURES actually used */
  int32_t linear_solver;          /* ARSLAM_LINSOLVE_DENSE or _PCG actually used */
  int32_t reduced_dim;            /* dimension of the reduced (Schur) system */
  int64_t num_jacobian_evals;     /* residual+Jacobian+accumulation evaluations */
  int64_t num_cost_evals;         /* candidate (residual-only) evaluations */
  int64_t linear_solver_iterations; /* total PCG iterations (0 for dense) */
  int64_t gpu_launches;           /* kernels launched by this solve */
  double initial_cost, final_cost;
  double final_radius;
  double gradient_max_norm;
  double total_ms;                /* host wall time of the call */
  double eval_ms, linsolve_ms;    /* device time (CUDA events) of the two phases */
} arslam_summary;

typedef struct arslam_solver arslam_solver; /* opaque; owns all device memory */

void arslam_default_options(arslam_options* opt);
int arslam_abi_version(void);

/* Lifetime == lifetime of the ArSlamSolver that owns `problem_`
 * (ar_slam_util.hpp:473).  `device` is a CUDA ordinal.  Fails with
 * ARSLAM_ERR_NO_DEVICE when there is no GPU: no CPU path exists. */
int arslam_create(int device, const arslam_options* opt, arslam_solver** out);
void arslam_destroy(arslam_solver* s);
const char* arslam_last_error(const arslam_solver* s); /* s may be NULL: creation errors */
int arslam_set_options(arslam_solver* s, const arslam_options* opt);
/* Run all work of this handle on the caller's CUDA stream (a cudaStream_t cast
 * to void*; NULL restores the handle's own stream), e.g. to bracket calls with
 * the caller's CUDA events.  (The dense factorisation forks part of its
 * trailing update to an internal lower-priority stream and joins it back
 * before it returns, so stream order is preserved for the caller.) */
int arslam_set_stream(arslam_solver* s, void* cuda_stream);

/* Replaces resetProblem (ar_slam_util.cpp:1021-1025) + the AddResidualBlock
 * loops (:720-727, :829-836): defines the whole set of residual blocks.
 * Block b couples camera, capture cap_idx[b] and tag tag_idx[b]; rect8 holds
 * 8 doubles per block (ArucoRect order).  Copies everything to HBM in
 * structure-of-arrays form, sorted once per pose side.  Under
 * arslam_comm_init every rank passes only ITS blocks (all blocks of a capture
 * on one rank) with GLOBAL n_cap / n_tag and global indices. */
int arslam_set_problem(arslam_solver* s, int64_t n_cap, int64_t n_tag, int64_t n_blk,
                       const int32_t* cap_idx, const int32_t* tag_idx, const double* rect8);

/* The incremental schedules (solveIncremental :629-678, solve :744-866) add the
 * blocks of one capture to the live ceres::Problem and solve again (:723, :832).
 * arslam_append_blocks is that AddResidualBlock loop for a problem already on
 * the device: only the n_new blocks cross the bus, the earlier observations
 * stay in HBM and the sorted copies are rebuilt there.  n_cap / n_tag are the
 * new totals (they may grow, never shrink).  Afterwards the result is the same,
 * bit for bit, as arslam_set_problem with all blocks; parameters must be set
 * again (arslam_set_params).  Without a current problem it is arslam_set_problem.
 * Not available under arslam_comm_init (ARSLAM_ERR_UNSUPPORTED). */
int arslam_append_blocks(arslam_solver* s, int64_t n_cap, int64_t n_tag, int64_t n_new,
                         const int32_t* cap_idx, const int32_t* tag_idx, const double* rect8);

/* Parameter blocks live in the caller's Capture::inv_pose / Aruco::pose /
 * camera_.params (ar_slam_util.hpp:72,208,237); Ceres updated them in place
 * through raw pointers, here they are copied in before and out after, straight
 * between the caller's arrays and HBM (DMA when the arrays are pinned; both
 * calls return after the copy has completed).  All arrays are indexed by the
 * global capture / tag index.  Under arslam_comm_init a rank reads and writes
 * only the captures of its own blocks -- the index range [min, max] of its
 * cap_idx; the rest of cap_pose6 is neither read nor written -- so ranks that
 * own disjoint capture ranges can share one output array or gather the
 * ranges themselves.  Camera and tag poses are replicated on every rank. */
int arslam_set_params(arslam_solver* s, const double* camera3, const double* cap_pose6,
                      const double* tag_pose6);
int arslam_get_params(arslam_solver* s, double* camera3, double* cap_pose6, double* tag_pose6);

/* Device-resident parameters for the incremental schedules (solveIncremental :629-678, solveCapture
 * :680-742, solve :744-866: one optimize() per added capture).  Between two solves only the new
 * capture's pose and the poses of tags seen for the first time change on the host side of the
 * reference; with these entries nothing else crosses the bus:
 *   arslam_set_camera      intrinsics alone (on a fresh problem it also zeroes all poses, so that a
 *                          schedule can start without arslam_set_params)
 *   arslam_set_poses / arslam_get_poses   `count` poses from index `first`; which: 0 captures, 1 tags
 *   arslam_seed_captures   initCapturePose (:91-108) on the device: capture cap_idx[i] is seeded from the
 *                          quad rect8[8 i ..] of tag tag_idx[i], whose pose is read from the device
 *   arslam_seed_tags       initArPose (:111-128): tag tag_idx[i] from the quad seen by capture cap_idx[i]
 * All use the current focal length on the device (camera_.params[0] in the reference).  Single GPU. */
int arslam_set_camera(arslam_solver* s, const double* camera3);
int arslam_set_poses(arslam_solver* s, int which, int64_t first, int64_t count, const double* pose6);
int arslam_get_poses(arslam_solver* s, int which, int64_t first, int64_t count, double* pose6);
int arslam_seed_captures(arslam_solver* s, int64_t n, const int32_t* cap_idx, const int32_t* tag_idx, const double* rect8);
int arslam_seed_tags(arslam_solver* s, int64_t n, const int32_t* tag_idx, const int32_t* cap_idx, const double* rect8);

/* Replaces ceres::Problem::SetParameterBlockConstant (ar_slam_util.cpp:965 camera, :972 tags in
 * localizeOne; the disabled gauge fix of the first capture at :697-700, :776-779) for the blocks of
 * arslam_set_problem: camera_constant != 0 holds all intrinsics, cap_constant[c] / tag_constant[t]
 * != 0 hold that pose (arrays over the global indices; NULL = none).  A constant block keeps its
 * value, is left out of the step, of the gradient and parameter norms of the convergence tests and
 * of the linear system -- what Ceres does by removing it from the program.  The masks are cleared
 * by arslam_set_problem / arslam_append_blocks. */
int arslam_set_constant(arslam_solver* s, int camera_constant, const uint8_t* cap_constant,
                        const uint8_t* tag_constant);

/* What ceres::Problem::Evaluate would return for the current parameters:
 * cost = 1/2 sum r^2, residuals [8 n_blk] (x0,y0,..,y3 per block,
 * ar_slam_util.cpp:207-208) and the AutoDiffCostFunction<..,8,3,6,6>
 * Jacobians (:722), row-major per parameter block: jac_cam [n_blk][8][3],
 * jac_cap [n_blk][8][6], jac_tag [n_blk][8][6].  Any output may be NULL.
 * Host pointers. */
int arslam_evaluate(arslam_solver* s, double* cost, double* residuals, double* jac_cam,
                    double* jac_cap, double* jac_tag);

/* Parity hook for the fused evaluation + accumulation kernels (kernels (1)+(2)): the block pieces
 * of J^T J and J^T r at the current parameters, exactly as the LM loop consumes them.  Focal-only
 * model, single GPU.  Any output may be NULL; all are host pointers.
 *   eliminated_side  ARSLAM_ELIM_CAPTURES / _TAGS actually used (the E side owns the rows of W)
 *   blk_cap, blk_tag [n_blk]      the blocks in the order W36 is returned in (sorted by E pose)
 *   W36     [n_blk][6][6]         J_E^T J_F of each block, rows = the E pose's (t, w), cols = the F pose's
 *   H_cap   [n_cap][33]           per capture: upper triangle of J_c^T J_c (21, row-major) | J_c^T r (6) | J_c^T J_f (6)
 *   H_tag   [n_tag][33]           the same per tag
 *   camera4                       sum (dr/df)^2, sum (dr/df) r, sum r^2, 0 */
int arslam_get_normal_equations(arslam_solver* s, int32_t* eliminated_side, int32_t* blk_cap, int32_t* blk_tag,
                                double* W36, double* H_cap, double* H_tag, double* camera4);

/* Replaces ArSlamSolver::optimize == ceres::Solve (ar_slam_util.cpp:1001-1018)
 * on the blocks of arslam_set_problem, starting from arslam_set_params.
 * iter_log (optional, host): log_rows x 8 doubles per iteration: cost,
 * cost_change, gradient_max_norm, step_norm, relative_decrease, radius,
 * step_is_valid, step_is_successful. */
int arslam_solve(arslam_solver* s, arslam_summary* summary, double* iter_log, int32_t log_rows);

/* Replaces localizeMany / localizeOne (ar_slam_util.cpp:888-979) for a batch
 * of independent captures against the fixed map given by camera3 / tag_pose6
 * (n_tag tags).  Capture i owns blocks blk_offsets[i] .. blk_offsets[i+1];
 * seed_block[i] is the index (within the capture) of the block whose tag is
 * shared with the map (:911-927) and seeds the pose through initCapturePose
 * (:91-108); seed_block[i] < 0 leaves capture i untouched (:929-933).
 * cap_pose6 [6 n_loc] out; iterations / final_cost / termination are optional
 * per-capture outputs (host pointers).  The batch is processed in chunks on three
 * internal streams (upload of chunk i + 1 under the kernel of chunk i, download of
 * chunk i - 1 under both; pinned host arrays make the copies asynchronous) which are
 * joined into the handle's stream before the call returns.  Offsets, seed blocks and
 * tag indices are validated on the device; on ARSLAM_ERR_INVALID no result is valid. */
int arslam_localize_batch(arslam_solver* s, int64_t n_loc, const int32_t* blk_offsets,
                          const int32_t* tag_idx, const double* rect8, const int32_t* seed_block,
                          int64_t n_tag, const double* camera3, const double* tag_pose6,
                          double* cap_pose6, int32_t* iterations, double* final_cost,
                          int32_t* termination);

/* Multi-GPU (new; the reference is single-threaded): one process per GPU.
 * Rank 0 calls arslam_comm_unique_id, ships the 128 bytes to the other ranks
 * by any means (torch.distributed, MPI, a file), then every rank calls
 * arslam_comm_init (before arslam_set_problem: it invalidates a problem
 * declared earlier).  Afterwards arslam_solve sums the per-rank partial normal
 * equations with one ncclAllReduce per linearisation over NVLink. */
int arslam_comm_unique_id(void* id128);
int arslam_comm_init(arslam_solver* s, int rank, int world_size, const void* id128);

/* Device-side timing of the last arslam_solve / arslam_localize_batch /
 * arslam_evaluate, for bench.py: fills up to `cap` (name,ms,launches) rows. */
typedef struct arslam_kernel_time {
  char name[48];
  double total_ms;
  int64_t launches;
  double algorithmic_bytes; /* per launch, by the formulas in DESIGN.md */
} arslam_kernel_time;
int arslam_set_profiling(arslam_solver* s, int on); /* brackets every launch with CUDA events */
/* Kernel-variant switches for A/B measurements (bench.py, tests): per handle, never read from the
 * environment.  Keys: "accum_pipe" (bit 0: E pass, bit 1: F pass on the cross-block pipelined
 * accumulation kernels instead of the thread-per-block ones that start every block cold; default 2:
 * measured faster for the F pass only); "accum_flush" (1: unrolled segment flush in those kernels); "schur_bulk" (1: the Schur products reach the
 * block-sparse reduced system as TMA bulk reductions, default; 0: per-lane FP64 reductions);
 * "locality" (1: the capture-sorted copy of the blocks is stored with the captures ordered by their
 * smallest tag; 0, default: by capture index; read by the next arslam_set_problem / append_blocks);
 * "schur_local" (1: elimination CTAs hold whole segments and pre-reduce the products per destination
 * inside the CTA -- an experiment kept for reference, measured slower than the default; 0, default: one
 * reduction per product); "loc_chunk" (captures per chunk of arslam_localize_batch's upload / kernel / download pipeline, 0:
 * default 131072); "pcg_smem" (1: the PCG kernels that keep the
 * reduced matrix in shared memory, default); "pcg_pipelined" (1: one-barrier pipelined recurrence for
 * pcg_tolerance >= 1e-6, default; 0: always the classic two-barrier recurrence);
 * dense Cholesky: "chol_big" (1: asynchronous trailing update -- cp.async operand stream across tiles, C through TMA
 * bulk reductions, default; 0: the synchronous round-1 kernel), "chol_chain" (1: back-substitution as one chained
 * cooperative launch, default; 0: one launch per 64-block), "chol_nb" (outer panel width, default 256),
 * "chol_free_sms" (SMs the persistent trailing update leaves to the factorisation chain, default 8).
 * Unknown key: ARSLAM_ERR_INVALID. */
int arslam_set_tuning(arslam_solver* s, const char* key, int64_t value);
int arslam_kernel_times(arslam_solver* s, arslam_kernel_time* out, int32_t cap); /* returns rows */

/* ---- Marker detection: the step before the path --------------------------------
 * Replaces cv::aruco::detectMarkers at ar_slam/src/aruco_detector.cpp:106 (ArucoDetector::image_callback) and
 * ar_slam/src/ar_slam_util.cpp:268 (ArSlamSolver::loadImages): 8-bit BGR or grey frames in, marker ids and the
 * four corners (pixel coordinates, marker's top-left first, clockwise, exactly cv::aruco's convention; the caller
 * centres them like from_cv_img, ar_slam_util.hpp:257-263) out.  Everything that touches pixels runs on the GPU
 * (grey conversion + three adaptive thresholds, border following, polygon approximation, perspective removal,
 * Otsu, bit extraction, dictionary match); the grouping of near-duplicate candidates (tens of quads per frame) is
 * host code, like the reference's schedules.  No corner refinement (cv's default CORNER_REFINE_NONE, which is
 * what the reference uses), no inverted markers, no ArUco3 pyramid.
 * Fields and defaults of cv::aruco::DetectorParameters that the path reads. */
typedef struct arslam_detect_params {
  int32_t adaptive_thresh_win_size_min;  /* 3  */
  int32_t adaptive_thresh_win_size_max;  /* 23 */
  int32_t adaptive_thresh_win_size_step; /* 10: windows 3, 13, 23 (at most 8 windows, each odd and <= 31) */
  int32_t min_distance_to_border;        /* 3  */
  int32_t marker_border_bits;            /* 1  */
  int32_t perspective_remove_pixel_per_cell; /* 4 */
  double adaptive_thresh_constant;       /* 7  */
  double min_marker_perimeter_rate;      /* 0.03 */
  double max_marker_perimeter_rate;      /* 4.0  */
  double polygonal_approx_accuracy_rate; /* 0.03 */
  double min_corner_distance_rate;       /* 0.05; ArSlamSolver::loadImages sets 0.1 (ar_slam_util.cpp:250) */
  double min_marker_distance_rate;       /* 0.125 */
  double min_group_distance;             /* 0.21 */
  double perspective_remove_ignored_margin_per_cell; /* 0.13 */
  double max_erroneous_bits_in_border_rate;          /* 0.35 */
  double min_otsu_std_dev;               /* 5.0 */
  double error_correction_rate;          /* 0.6 */
} arslam_detect_params;
void arslam_detect_default_params(arslam_detect_params* p);

typedef struct arslam_detector arslam_detector; /* opaque; owns its device workspace and stream */

/* Workspace for batches of up to max_images frames of up to max_width x max_height pixels.  The dictionary is
 * DICT_4X4_50 (the reference's default, aruco_detector.cpp:71, ar_slam_util.cpp:251-252) until
 * arslam_detector_set_dictionary replaces it.  ARSLAM_ERR_NO_DEVICE without an sm_100 GPU: no CPU path. */
int arslam_detector_create(int device, int32_t max_images, int32_t max_width, int32_t max_height,
                           arslam_detector** out);
void arslam_detector_destroy(arslam_detector* d);
const char* arslam_detector_last_error(const arslam_detector* d); /* d may be NULL: creation errors */
/* Any square dictionary (aruco_detector.cpp:148-152 offers 4X4_50, 5X5_100, 6X6_250): n_markers codes of
 * marker_size x marker_size bits, one byte per bit, row-major, rotation 0 (cv::aruco::Dictionary::getBitsFromByteList);
 * max_correction_bits as in cv::aruco::Dictionary.  marker_size <= 6. */
int arslam_detector_set_dictionary(arslam_detector* d, int32_t n_markers, int32_t marker_size,
                                   int32_t max_correction_bits, const uint8_t* bits);
/* The embedded tables by the names of aruco_detector.cpp:148-152: "4X4_50", "5X5_100", "6X6_250". */
int arslam_detector_set_predefined_dictionary(arslam_detector* d, const char* name);
/* n_images frames of width x height pixels, channels = 1 (grey) or 3 (BGR, cv::imread / cv_bridge order),
 * tightly packed one after the other; on_device != 0: `images` is a device pointer (frames already in HBM, complete
 * before the call: the detector works on its own stream).
 * Per image i: n_found[i] markers, written to ids[i * max_markers ...] and corners[(i * max_markers + k) * 8 ...]
 * as x0,y0,...,x3,y3 in cv::aruco's order of detection.  More than max_markers detections in a frame:
 * ARSLAM_ERR_INVALID.  The contour workspace grows on demand (a batch of pure noise costs one re-run). */
int arslam_detect_markers(arslam_detector* d, const uint8_t* images, int32_t n_images, int32_t width,
                          int32_t height, int32_t channels, int32_t on_device, const arslam_detect_params* params,
                          int32_t max_markers, int32_t* n_found, int32_t* ids, float* corners);
/* Candidate quads of the last call before grouping (parity tests and debugging): per candidate image index,
 * threshold window index, 8 corner coordinates (clockwise), too-near-border flag, identified id (-1: none) and
 * rotation; in cv::aruco's candidate order.  Returns the number of candidates (<= cap rows are written). */
int arslam_detector_candidates(arslam_detector* d, int32_t cap, int32_t* image, int32_t* window, float* corners,
                               int32_t* near_border, int32_t* id, int32_t* rotation);
/* Intermediate results of the last call, for the stage-by-stage parity tests: what = 0 grey frames (n x h x w bytes),
 * 1 threshold bits (n x h x w bytes, bit k = window k), 2 the traced borders (5 int32 each: image, window, raster
 * index of discovery in the zero-padded frame -- pitch ((w + 32 + 15) / 16) * 16, pixel (x, y) at (y + 1) * pitch + 16 + x --,
 * length, offset into the points), 3 border
 * points (int32 x | y << 16).  Returns the bytes written; out == NULL: the bytes needed. */
int64_t arslam_detector_read_stage(arslam_detector* d, int32_t what, void* out, int64_t cap_bytes);
/* Device time of the stages of the last arslam_detect_markers in ms: [0] grey + thresholds, [1] border starts,
 * [2] border following (incl. the host's read of the start count before it), [3] polygon approximation,
 * [4] identification + candidate table download, [5] whole call on the device incl. copies; kernel launches. */
int arslam_detector_times(arslam_detector* d, double* ms6, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* AR_SLAM_B200_H_ */
