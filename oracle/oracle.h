/*
 * oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A from-scratch CPU restatement of the path MC-Calib hands to ceres-solver 2.2.0
 * (pinned at /root/reference/Dockerfile:48, un-vendored, absent from this image):
 * forward-mode dual-number evaluation of the five cost functors of
 * McCalib/include/OptimizationCeres.h, ceres::HuberLoss + Corrector, Jacobi scaling,
 * Levenberg-Marquardt trust region with Ceres' step-acceptance rules, Schur elimination of
 * the e-blocks and a dense Cholesky of the reduced system.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference has no golden vectors, known-answer tests or
 * fixtures for this path and cannot be compiled here (needs Ceres, Eigen, OpenCV,
 * Boost). The oracle is instead pinned by tests/test_oracle_*.py against independent evaluators
 * (mpmath 50-digit functor KATs, cv2.projectPoints / cv2.fisheye.projectPoints values and
 * Jacobians, scipy.optimize.least_squares(loss='huber') minima) and by committed trace fixtures
 * under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this. It shares the problem/parameter/option structs of include/mcba.h so both
 * sides are driven by identical inputs.
 */
#ifndef MCBA_ORACLE_H
#define MCBA_ORACLE_H
#include "../include/mcba.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Problem::Evaluate. jacobian is [n_obs][2][K] row-major in the reference's parameter-block order. */
int oracle_eval(const mcba_problem *prob, const mcba_params *params, const mcba_options *opt,
                double *cost, double *residuals, double *jacobian, double *gradient_max_norm);

/* ceres::Solve. n_threads = 1 is the reference as shipped (Solver::Options::num_threads default). */
int oracle_solve(const mcba_problem *prob, mcba_params *params, const mcba_options *opt,
                 mcba_summary *summary, mcba_trace_row *trace, int trace_cap, int n_threads);

/* Timed pieces for the CPU baseline: n_iters repetitions of (Jacobian evaluation + normal equations
 * + Schur + Cholesky + back-substitution + candidate cost) at fixed parameters; returns seconds per
 * repetition and the split. */
int oracle_time_iteration(const mcba_problem *prob, const mcba_params *params, const mcba_options *opt,
                          int n_iters, int n_threads, double *sec_per_iter, double *sec_eval,
                          double *sec_linear);

/* Single functor evaluation for known-answer tests. blocks: cam_pose[6], pose[6], board_pose[6],
 * intr[9]; flags bit0 = cam_is_ref, bit1 = board_is_ref. Outputs the raw (un-robustified) residual
 * and Jacobian [2][K]. */
int oracle_functor(int stage, int model, int flags, const double *cam_pose, const double *pose,
                   const double *board_pose, const double *intr, const double *xyz, const double *uv,
                   double *residual, double *jacobian);

/* ceres::AngleAxisRotatePoint and its Jacobians d(result)/d(angle_axis) [3][3], d(result)/d(pt) [3][3]. */
void oracle_angle_axis_rotate(const double *angle_axis, const double *pt, double *result,
                              double *d_aa, double *d_pt);

/* HuberLoss::Evaluate: rho[3]. */
void oracle_huber(double a, double s, double *rho);

int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
