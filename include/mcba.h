/*
 * mcba.h — C ABI of libmcba, the B200-native replacement for the bundle-adjustment hot path that
 * MC-Calib hands to Ceres.
 *
 * One mcba_problem corresponds to one `ceres::Problem` the reference builds inside a refine* stage:
 *   stage 1  Camera::refineIntrinsicCalibration            McCalib/src/Camera.cpp:268-305
 *   stage 2  Object3D::refineObject                        McCalib/src/Object3D.cpp:160-243
 *   stage 3  CameraGroup::refineCameraGroup                McCalib/src/CameraGroup.cpp:191-279
 *   stage 4  CameraGroup::refineCameraGroupAndObjects      McCalib/src/CameraGroup.cpp:366-474
 *   stage 5  CameraGroup::refineCameraGroupAndObjectsAndIntrinsics  McCalib/src/CameraGroup.cpp:483-583
 * mcba_solve replaces the `ceres::Solve(options, &problem, &summary)` call at the end of each
 * (Camera.cpp:301, Object3D.cpp:223, CameraGroup.cpp:273,468,577); mcba_eval replaces one
 * `Problem::Evaluate` (residuals, robustified Jacobian blocks, cost) of the cost functors in
 * McCalib/include/OptimizationCeres.h:11-665.
 *
 * All pointers are caller-owned HOST memory; nothing here is a torch or C++ type. Functions return
 * 0 on success and a negative mcba_status otherwise (parameters are then left untouched);
 * mcba_last_error() gives the text. A handle is thread-compatible, not thread-safe.
 */
#ifndef MCBA_H
#define MCBA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum mcba_status {
  MCBA_OK = 0,
  MCBA_ERR_INVALID = -1,   /* bad argument / inconsistent problem description */
  MCBA_ERR_CUDA = -2,      /* CUDA runtime error, or no sm_100 device */
  MCBA_ERR_NCCL = -3,      /* NCCL error / libnccl missing when n_ranks > 1 */
  MCBA_ERR_NUMERIC = -4    /* non-finite initial residual or Jacobian (Ceres: "Initial residual and Jacobian evaluation failed") */
};

enum mcba_termination { MCBA_CONVERGENCE = 0, MCBA_NO_CONVERGENCE = 1, MCBA_FAILURE = 2 };

/*
 * Observations are grouped in "board observations" (bobs): runs of corners that share the same
 * (camera, e-block pose, board). Which parameter blocks exist is fixed by `stage`:
 *   stage 1: e-block = BoardObs::pose_ (BoardObs.hpp:33);  f-blocks = Camera::intrinsics_ (Camera.hpp:46)
 *   stage 2: e-block = Object3DObs::pose_ (Object3DObs.hpp:40); f-blocks = Object3D::relative_board_pose_ (Object3D.hpp:48);
 *            camera intrinsics are constants
 *   stage 3: e-block = CameraGroupObs::object_pose_ (CameraGroupObs.hpp:29-30); f-blocks = CameraGroup::relative_camera_pose_
 *            (CameraGroup.hpp:40-41); points are object-frame points (Object3D::pts_3d_), no board block
 *   stage 4: stage 3 + Object3D::relative_board_pose_; points are board-local (Board::pts_3d_)
 *   stage 5: stage 4 + Camera::intrinsics_
 * The e-blocks are the ones Ceres' automatic SPARSE_SCHUR ordering eliminates.
 */
typedef struct mcba_problem {
  int32_t stage;               /* 1..5 */
  int32_t n_cam, n_pose, n_board, n_pts;
  int64_t n_obs, n_bobs;
  const float   *u, *v;        /* [n_obs] observed pixel, fp32 as cv::Point2f */
  const int32_t *pt_idx;       /* [n_obs] index into pts_xyz */
  const int64_t *bobs_ptr;     /* [n_bobs+1] CSR offsets into the observation arrays */
  const int32_t *bobs_cam;     /* [n_bobs] camera slot (always present: intrinsics come from it) */
  const int32_t *bobs_pose;    /* [n_bobs] e-block slot; must be non-decreasing (bobs sorted by e-block) */
  const int32_t *bobs_board;   /* [n_bobs] board slot, ignored in stages 1 and 3 */
  const float   *pts_xyz;      /* [n_pts*3] fp32 as cv::Point3f */
  const int32_t *cam_model;    /* [n_cam] 0 = Brown (perspective), 1 = Kannala (fisheye) */
  const uint8_t *cam_is_ref;   /* [n_cam] 1: pose block has a structurally zero Jacobian (refine_camera=false,
                                  OptimizationCeres.h:420-427); may be NULL (= none) */
  const uint8_t *board_is_ref; /* [n_board] 1: refine_board=false (OptimizationCeres.h:401-409); may be NULL */
  const int32_t *cam_problem;  /* batched stage 1 only: [n_cam] independent-problem id of each camera, or NULL
                                  (= one problem). Problems are solved simultaneously and independently. */
  int32_t n_problems;          /* number of independent problems (1 unless cam_problem != NULL) */
} mcba_problem;

/* Parameter blocks, in/out. Layouts as the reference: pose = angle-axis(3) ‖ translation(3);
 * intrinsics = fx fy u0 v0 | Brown k1 k2 p1 p2 k3 | Kannala k1 k2 k3 k4 (unused) (Camera.hpp:43-46). */
typedef struct mcba_params {
  double *cam_pose;    /* [n_cam*6]   CameraGroup::relative_camera_pose_ (stages 3-5; may be NULL otherwise) */
  double *pose;        /* [n_pose*6]  the stage's e-blocks */
  double *board_pose;  /* [n_board*6] Object3D::relative_board_pose_ (stages 2,4,5; may be NULL otherwise) */
  double *intrinsics;  /* [n_cam*9]   always read; written only in stages 1 and 5 */
} mcba_params;

/* Defaults (mcba_default_options) are the Ceres 2.2.0 Solver::Options defaults with the three
 * overrides the reference makes (SPARSE_SCHUR, max_num_iterations, progress to stdout). */
typedef struct mcba_options {
  int32_t max_num_iterations;            /* reference: YAML number_iterations */
  double function_tolerance;             /* 1e-6  */
  double gradient_tolerance;             /* 1e-10 */
  double parameter_tolerance;            /* 1e-8  */
  double initial_trust_region_radius;    /* 1e4   */
  double max_trust_region_radius;        /* 1e16  */
  double min_trust_region_radius;        /* 1e-32 */
  double min_relative_decrease;          /* 1e-3  */
  double min_lm_diagonal;                /* 1e-6  */
  double max_lm_diagonal;                /* 1e32  */
  double huber_a;                        /* 1.0; <= 0 disables the loss (trivial loss) */
  int32_t jacobi_scaling;                /* 1 */
  int32_t max_num_consecutive_invalid_steps; /* 5 */
  int32_t verbose;                       /* 1: print the Ceres-style iteration table to stdout */
  int32_t materialize_jacobian;          /* 0: fused resid+Jac+JtJ kernel; 1: K1 writes J to HBM, K2 reads it back */
} mcba_options;

typedef struct mcba_summary {
  int32_t termination;        /* mcba_termination */
  int32_t num_successful_steps, num_unsuccessful_steps;  /* as ceres::Solver::Summary (iteration 0 counts as successful) */
  int32_t num_iterations;     /* recorded iterations, including iteration 0 */
  double initial_cost, final_cost;
  double rms_px;              /* sqrt(sum(du^2+dv^2)/n_obs) of the unrobustified residuals at the final parameters */
  double t_total_ms, t_jacobian_ms, t_schur_ms, t_chol_ms, t_cost_ms, t_comm_ms;  /* device time, CUDA events */
} mcba_summary;

/* One row per recorded iteration, same columns as Ceres' minimizer_progress_to_stdout table. */
typedef struct mcba_trace_row {
  int32_t iteration, step_is_valid, step_is_successful, _pad;
  double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, trust_region_radius;
} mcba_trace_row;

void mcba_default_options(mcba_options *opt);

/* Pack the problem to the device SoA layout and copy it to HBM once; the handle can be re-solved
 * (the reference runs stage 4 then stage 5 on the same observations, apps/calibrate/src/calibrate.cpp:60-64:
 * use mcba_set_stage). device = CUDA ordinal. */
int mcba_create(void **handle, const mcba_problem *prob, int device);
int mcba_set_stage(void *handle, int stage);

/* Device-resident observations shared between calls. The pose initialisation (mcba_pnp_ransac, mcba_pnp.h) and the
 * bundle adjustment that follows it read the same corner arrays (the reference keeps them once, in BoardObs /
 * Object3DObs: McCalib/include/BoardObs.hpp:27-33). After mcba_obs_register every mcba_create / mcba_pnp_ransac whose
 * u, v, pt_idx pointers, n_obs and device match uses the resident copy instead of uploading again (12 B per corner).
 * mcba_obs_release frees it; it fails with MCBA_ERR_INVALID while a handle still uses the copy. */
int mcba_obs_register(const float *u, const float *v, const int32_t *pt_idx, int64_t n_obs, int device);
int mcba_obs_release(const float *u, int device);

/* Multi-GPU: this process owns one GPU and one shard (contiguous e-block range, all f-blocks replicated).
 * nccl_unique_id is the 128-byte ncclUniqueId produced by mcba_nccl_unique_id on rank 0 and broadcast
 * by the caller (torch.distributed / MPI / file). Must be called before mcba_solve on every rank. */
int mcba_nccl_unique_id(void *id_out_128B);
int mcba_comm_init(void *handle, int rank, int n_ranks, const void *nccl_unique_id_128B);
/* Which data plane sums the reduced system across ranks: 0 = single rank, 1 = NCCL allreduce, 2 = fused into the
 * Schur kernel's epilogue over NVLink peer memory (the default when CUDA IPC is available on every rank). */
int mcba_schur_reduce_mode(void *handle);

/* ceres::Solve replacement. trace may be NULL. For batched problems summary/trace describe problem 0
 * unless summaries (an array of n_problems) is given through mcba_solve_batched. */
int mcba_solve(void *handle, mcba_params *params, const mcba_options *opt, mcba_summary *summary,
               mcba_trace_row *trace, int trace_cap);
int mcba_solve_batched(void *handle, mcba_params *params, const mcba_options *opt,
                       mcba_summary *summaries /* [n_problems] */);

/* Stepping interface under mcba_solve (used by the benchmark to time exactly K LM iterations):
 * begin = upload parameters + iteration 0; step = one trust-region iteration, returns 1 while the
 * minimizer can continue, 0 when it terminated; end = download parameters + summary. */
int mcba_solve_begin(void *handle, const mcba_params *params, const mcba_options *opt);
int mcba_solve_step(void *handle);
int mcba_solve_end(void *handle, mcba_params *params, mcba_summary *summary);

/* Problem::Evaluate replacement at the given parameters: cost = sum 1/2 rho(|r|^2); residuals [n_obs*2]
 * and jacobian [n_obs][2][K] (row-major, K = sum of the stage's block sizes in the reference's
 * parameter-block order) are the loss-corrected ones Ceres hands to its linear solver. Any output may be NULL. */
int mcba_eval(void *handle, const mcba_params *params, const mcba_options *opt, double *cost,
              double *residuals, double *jacobian);
int mcba_jacobian_width(int stage);

/* Per-observation reprojection error sqrt(du^2+dv^2) at the given parameters (unrobustified), the quantity
 * Calibration::computeAvgReprojectionError averages (McCalib/src/McCalib.cpp:2168-2245). */
int mcba_reprojection_errors(void *handle, const mcba_params *params, double *err /* [n_obs] */);
/* The same errors as the reference's OUTPUTS hold them, i.e. with its intermediate roundings: object-frame points
 * after the object pose stored as cv::Point3f (transform3DPts, McCalib/src/geometrytools.cpp:328-352), projections
 * returned as cv::Point2f by cv::projectPoints, float difference, double sqrt stored in a float
 * (computeDistanceBetweenPoints, McCalib/src/McCalib.cpp:2150-2160). Use on a stage-3 packing (Object3D::pts_3d_). */
int mcba_reprojection_errors_f32(void *handle, const mcba_params *params, float *err /* [n_obs] */);

void mcba_destroy(void *handle);
const char *mcba_last_error(void *handle);

/* Timing hooks for benchmarks: device time (ms, CUDA events on the library's stream) of the last
 * launch of each kernel family; names in mcba_kernel_name. */
int mcba_num_kernels(void);
const char *mcba_kernel_name(int k);
int mcba_kernel_stats(void *handle, int k, int64_t *launches, double *total_ms);
int mcba_reset_kernel_stats(void *handle);
int64_t mcba_kernel_launches(void *handle);          /* kernels launched by this handle so far */
int mcba_timer_start(void *handle);                  /* CUDA event on the library's stream (after a stream sync) */
int mcba_timer_stop(void *handle, double *ms);       /* second event + elapsed device time between the two */

/* Binary dump / load of a packed problem with (optionally) its parameter blocks: the flat arrays above, as they are
 * ("MCBAPRB1", layout in csrc/mcba_io.cpp). Replay of a bundle adjustment, resume from the parameters of an earlier
 * solve, parity fixtures. The reference's own checkpoint is one step earlier and textual (detected_keypoints_data.yml +
 * the camera-parameter file, McCalib/src/McCalib.cpp:493-548, :172-226, :662-688; host/McCalibInit.cpp reads those).
 * mcba_problem_load allocates the arrays; *storage owns them until mcba_problem_free. params may be NULL on save (no
 * parameter section) and on load (sections skipped); absent sections load as NULL pointers. Host only. */
int mcba_problem_save(const char *path, const mcba_problem *prob, const mcba_params *params);
int mcba_problem_load(const char *path, mcba_problem *prob, mcba_params *params, void **storage);
void mcba_problem_free(void *storage);
const char *mcba_io_last_error(void);

/* Test hook: copy an internal device buffer (normal-equation blocks, reduced system, scaling) to the host.
 * Names: "FF" "gf" "S" "rhs" "zf" "ZtZ" "Hoo" "Wt" "Z" "brec" "pair_rec" "scale_e" "scale_f" "diag_e" "diag_f". */
int mcba_debug_fetch(void *handle, const char *name, double *out, int64_t cap, int64_t *n_out);
/* Test hook: fused observation kernel of the handles created afterwards. 3 = k_observations_pair (two board
 * observations per warp; the default at every stage), 0 = k_observations_fused (one per warp), 1 = first version
 * (k_observations<S, MODE_FUSED>), 2 = streaming experiment, 4 = 24-corner passes with 16 warps per SM (experiment);
 * -1 = back to the MCBA_K1 environment variable. Variants 0-3 write bit-identical per-board-observation records; 4
 * differs at rounding level (other multiply-add contractions). + 32: variants 0 and 3 keep the f x f / f x r part of
 * J~^T J~ in registers over runs of one (camera, board) pair and write partial pair records ("pair_rec" after the
 * assembly) instead of that part of every record (MCBA_K1_ACC=1 does the same; measured slower, off by default);
 * + 16 forces it off. */
int mcba_debug_set_k1(int variant);
/* Test / benchmark hook: factor and solve one dense symmetric positive definite system with the reduced-system
 * kernels alone. S_host is (n+1) x n row-major (matrix rows, then the right-hand side as row n); z_out [n] = solution,
 * L_out [(n+1) x n] = the factor in the lower triangle (row n = L^-1 rhs). variant 0: grid-synchronous kernel,
 * 1: dataflow kernel (one CTA per 64x64 tile), -1: the solver's own choice. ms_out = mean kernel time over iters. */
int mcba_debug_cholesky(void *handle, int n, const double *S_host, double *z_out, double *L_out, int variant,
                        int iters, double *ms_out, int *fail_out);

#ifdef __cplusplus
}
#endif
#endif /* MCBA_H */
