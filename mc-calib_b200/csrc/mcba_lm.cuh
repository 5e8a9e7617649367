// mcba_lm.cuh — the trust-region bookkeeping of ceres-solver 2.2.0's TrustRegionMinimizer +
// LevenbergMarquardtStrategy (defaults), as plain __host__ __device__ functions on a POD
// state so that the single-problem driver (host, mcba_api.cu) and the batched driver (one device thread per
// independent problem, mcba_batched.cuh) take decisions from exactly the same code.
#pragma once
#include <cfloat>
#include <cmath>
#include "../../include/mcba.h"
#include "mcba_math.cuh"

namespace mcba {

struct LmState {
  double radius, decrease_factor, x_cost, x_norm, grad_max, minimum_cost, initial_cost, prev_grad_max;
  int reuse_diagonal, at_least_one_success, num_consecutive_invalid, termination, done;
  int num_successful, num_unsuccessful, num_recorded;
  mcba_trace_row it;   // summary of the iteration in flight
};
struct LmStep { double mcc, step_norm, cand_norm, cand_cost; int solver_ok; };
enum { LM_INVALID = 0, LM_TERMINATED = 1, LM_ACCEPT = 2, LM_REJECT = 3 };

MCBA_HD void lm_clear_row(mcba_trace_row& r) {
  r.iteration = 0; r.step_is_valid = 0; r.step_is_successful = 0; r._pad = 0;
  r.cost = 0; r.cost_change = 0; r.gradient_max_norm = 0; r.step_norm = 0; r.relative_decrease = 0; r.trust_region_radius = 0;
}

MCBA_HD void lm_init(LmState& s, const mcba_options& o) {
  s.radius = o.initial_trust_region_radius; s.decrease_factor = 2.0;
  s.x_cost = 0; s.x_norm = 0; s.grad_max = 0; s.minimum_cost = DBL_MAX; s.initial_cost = 0; s.prev_grad_max = 0;
  s.reuse_diagonal = 0; s.at_least_one_success = 0; s.num_consecutive_invalid = 0;
  s.termination = MCBA_NO_CONVERGENCE; s.done = 0;
  s.num_successful = 0; s.num_unsuccessful = 0; s.num_recorded = 0;
  lm_clear_row(s.it);
}

// IterationZero, after the first evaluation of cost / gradient / Jacobian.
MCBA_HD void lm_iteration_zero(LmState& s, double cost, double grad_max, double x_norm) {
  s.x_cost = cost; s.grad_max = grad_max; s.x_norm = x_norm; s.initial_cost = cost;
  lm_clear_row(s.it);
  s.it.iteration = 0; s.it.step_is_valid = 1; s.it.step_is_successful = 1; s.it.cost = cost; s.it.gradient_max_norm = grad_max;
}

// FinalizeIterationAndCheckIfMinimizerCanContinue: records the iteration in flight (the caller may copy s.it out as
// a trace row right after), then MaxSolverIterationsReached / GradientToleranceReached / MinTrustRegionRadiusReached.
// Returns 1 and opens the next iteration when the minimizer continues, 0 (done, termination set) otherwise.
MCBA_HD int lm_finalize(LmState& s, const mcba_options& o, mcba_trace_row* recorded) {
  if (s.it.step_is_successful) {
    ++s.num_successful;
    if (s.x_cost < s.minimum_cost) s.minimum_cost = s.x_cost;
  } else {
    ++s.num_unsuccessful;
  }
  s.it.trust_region_radius = s.radius;
  ++s.num_recorded;
  if (recorded) *recorded = s.it;
  if (s.it.iteration >= o.max_num_iterations) { s.termination = MCBA_NO_CONVERGENCE; s.done = 1; return 0; }
  if (s.it.step_is_successful && s.it.gradient_max_norm <= o.gradient_tolerance) { s.termination = MCBA_CONVERGENCE; s.done = 1; return 0; }
  if (s.it.trust_region_radius <= o.min_trust_region_radius) { s.termination = MCBA_CONVERGENCE; s.done = 1; return 0; }
  s.prev_grad_max = s.it.gradient_max_norm;
  const int iter = s.it.iteration + 1;
  lm_clear_row(s.it);
  s.it.iteration = iter;
  return 1;
}

// Everything between ComputeTrustRegionStep and the accept/reject branch. On LM_ACCEPT the caller makes the
// candidate the current point, re-evaluates there and calls lm_accepted_eval.
MCBA_HD int lm_decide(LmState& s, const mcba_options& o, const LmStep& st) {
  s.it.step_is_valid = st.solver_ok && st.mcc > 0.0;
  if (!s.it.step_is_valid) {   // HandleInvalidStep
    if (++s.num_consecutive_invalid >= o.max_num_consecutive_invalid_steps) { s.termination = MCBA_FAILURE; s.done = 1; return LM_TERMINATED; }
    s.radius /= s.decrease_factor; s.decrease_factor *= 2.0; s.reuse_diagonal = 1;   // StepIsInvalid -> StepRejected(0)
    s.it.cost = s.x_cost; s.it.gradient_max_norm = s.prev_grad_max;
    return LM_INVALID;
  }
  s.num_consecutive_invalid = 0;
  if (s.at_least_one_success) {
    s.it.step_norm = st.step_norm;   // ParameterToleranceReached
    if (st.step_norm <= o.parameter_tolerance * (s.x_norm + o.parameter_tolerance)) { s.termination = MCBA_CONVERGENCE; s.done = 1; return LM_TERMINATED; }
    s.it.cost_change = s.x_cost - st.cand_cost;   // FunctionToleranceReached
    if (fabs(s.it.cost_change) <= o.function_tolerance * s.x_cost) { s.termination = MCBA_CONVERGENCE; s.done = 1; return LM_TERMINATED; }
  }
  double rel;   // TrustRegionStepEvaluator::StepQuality, monotonic: reference cost == current cost
  if (st.cand_cost >= DBL_MAX) rel = -DBL_MAX;
  else rel = (s.x_cost - st.cand_cost) / st.mcc;
  s.it.relative_decrease = rel;
  if (rel > o.min_relative_decrease) {
    s.at_least_one_success = 1;
    s.x_norm = st.cand_norm;
    return LM_ACCEPT;
  }
  s.it.step_is_successful = 0; s.it.cost = st.cand_cost; s.it.gradient_max_norm = s.prev_grad_max;
  s.radius /= s.decrease_factor; s.decrease_factor *= 2.0; s.reuse_diagonal = 1;   // StepRejected
  return LM_REJECT;
}

// HandleSuccessfulStep after EvaluateGradientAndJacobian at the accepted point; LevenbergMarquardtStrategy::StepAccepted.
MCBA_HD void lm_accepted_eval(LmState& s, const mcba_options& o, double cost, double grad_max) {
  s.x_cost = cost; s.grad_max = grad_max;
  s.it.step_is_successful = 1; s.it.cost = cost; s.it.gradient_max_norm = grad_max;
  const double rel = s.it.relative_decrease;
  const double t = 2.0 * rel - 1.0;
  s.radius = s.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
  s.radius = fmin(o.max_trust_region_radius, s.radius);
  s.decrease_factor = 2.0;
  s.reuse_diagonal = 0;
}

}  // namespace mcba
