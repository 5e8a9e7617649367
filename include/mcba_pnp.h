/*
 * mcba_pnp.h — C ABI of the pose initialisation that feeds the bundle adjustment: every board
 * (or object) observation gets its pose w.r.t. its camera from a P3P RANSAC followed by a pose-only non-linear
 * refinement on the inliers — all observations of a calibration in ONE call (four kernels; a group of 4-32 lanes per
 * observation, several observations per warp).
 *
 * Replaces, in the reference (paths under /root/reference/McCalib/):
 *   ransacP3PDistortion / ransacP3P          src/geometrytools.cpp:709-751, :233-326
 *   BoardObs::estimatePose                   src/BoardObs.cpp:121-163     (called per board observation by
 *   Object3DObs::estimatePose                src/Object3DObs.cpp:237       Calibration::estimatePoseAllBoards / AllObjects,
 *                                                                         src/McCalib.cpp:703-706, :990-993)
 * which loop on the host over cv::solvePnP(SOLVEPNP_P3P) + cv::projectPoints per hypothesis and finish with
 * cv::solvePnP(SOLVEPNP_ITERATIVE) (OpenCV 4.x, a third-party dependency of the reference).
 *
 * Differences that are deliberate and documented:
 *   - the reference draws its samples from std::mt19937 re-seeded from std::random_device at every draw, i.e. it is
 *     not reproducible; here draw h of observation b is a pure function of (seed, b, h) (splitmix64), so a run is
 *     reproducible and the CPU checker can replay exactly the same hypotheses;
 *   - a hypothesis whose three points are (nearly) collinear, or for which P3P has no solution in front of the
 *     camera, scores zero inliers (the reference scores whatever stale pose cv::solvePnP left in its output);
 *   - arithmetic is fp64 throughout (OpenCV projects into cv::Point2f).
 *
 * All pointers are caller-owned HOST memory. Returns 0 or a negative mcba_status (mcba.h).
 */
#ifndef MCBA_PNP_H
#define MCBA_PNP_H

#include <stdint.h>
#include "mcba.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcba_pnp_problem {
  int64_t n_obs;
  const float   *u, *v;        /* [n_obs] detected corners (BoardObs::pts_2d_ / Object3DObs::pts_2d_), fp32 as cv::Point2f */
  const int32_t *pt_idx;       /* [n_obs] index into pts_xyz (BoardObs::charuco_id_ / Object3DObs::pts_id_ + table offset) */
  int64_t n_bobs;
  const int64_t *bobs_ptr;     /* [n_bobs+1] CSR offsets: the corners of one board (object) observation */
  const int32_t *bobs_cam;     /* [n_bobs] camera slot */
  int32_t n_pts;
  const float   *pts_xyz;      /* [n_pts*3] Board::pts_3d_ (Object3D::pts_3d_), fp32 as cv::Point3f */
  int32_t n_cam;
  const int32_t *cam_model;    /* [n_cam] 0 Brown, 1 Kannala */
  const double  *intrinsics;   /* [n_cam*9] fx fy u0 v0 | k1 k2 p1 p2 k3 (Brown) | k1 k2 k3 k4 - (Kannala) */
} mcba_pnp_problem;

typedef struct mcba_pnp_options {
  double  threshold_px;        /* YAML ransac_threshold (McCalib.cpp:58); inlier: reprojection distance < threshold */
  int32_t max_iterations;      /* `it`: the reference passes YAML number_iterations (McCalib.cpp:705) */
  double  confidence;          /* p = 0.99 (geometrytools.hpp:38) */
  int32_t refine;              /* 1: pose-only LM on the inliers when there are >= 4 (geometrytools.cpp:314-324) */
  int32_t refine_max_iterations; /* 20 in cv::solvePnP; the LM here stops earlier on convergence */
  uint64_t seed;
} mcba_pnp_options;

typedef struct mcba_pnp_result {
  double  *pose;               /* [n_bobs*6] angle-axis || translation; zeros when no hypothesis had an inlier */
  int32_t *n_inliers;          /* [n_bobs]   BoardObs::valid_ = (n_inliers >= 4), BoardObs.cpp:149 */
  uint8_t *inlier_mask;        /* [n_obs]    1 = inlier of the selected hypothesis (the corners BoardObs.cpp:153-161 keeps) */
  int32_t *best_hypothesis;    /* [n_bobs]   draw index of the selected hypothesis, -1 = none (may be NULL) */
  int32_t *n_hypotheses;       /* [n_bobs]   draws consumed by the stopping rule (may be NULL) */
  double  *unrefined_pose;     /* [n_bobs*6] pose of the selected hypothesis before refinement (may be NULL) */
  double   kernel_ms;          /* out: device time of the kernels (CUDA events) */
  double   total_ms;           /* out: wall time of the call including H2D / D2H */
  int64_t  kernel_launches;    /* out */
} mcba_pnp_result;

void mcba_pnp_default_options(mcba_pnp_options *opt);
int  mcba_pnp_ransac(const mcba_pnp_problem *prob, const mcba_pnp_options *opt, int device, mcba_pnp_result *res);
const char *mcba_pnp_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MCBA_PNP_H */
