// stub_mcba.cpp — TEST INFRASTRUCTURE: a CPU stand-in for libmcba's C ABI (include/mcba.h, include/mcba_pnp.h) so that
// the host-side facade (mc-calib_b200/host: packers, refineAll* flows, pose-initialisation flows, file formats) can be
// exercised by the CPU-only test suite. It forwards to the CPU oracle (oracle/liboracle.so) and to the sequential host
// restatement of the PnP kernels (tests/host_math). It is linked ONLY into tests/stub_abi/*_cpu binaries, never into
// the product: the product library has no CPU path.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcba_pnp.h"
#include "../../oracle/oracle.h"
#include "../../mc-calib_b200/csrc/mcba_math.cuh"

extern "C" int pnptest_ransac(int n, const float* u, const float* v, const float* xyz, int model, const double* intr,
                              double thresh, int it, double p, int refine, int refine_it, unsigned long long seed,
                              unsigned long long b, double* pose, double* pose0, unsigned char* mask, int* best_h, int* n_hyp);

namespace {
struct Stub {
  mcba_problem P;
  std::vector<float> u, v, pts; std::vector<int32_t> pt, cam, pose, board, model, cam_problem; std::vector<int64_t> ptr;
  std::vector<uint8_t> cam_ref, board_ref;
  std::string err;
};
template <class T> void copy(std::vector<T>& dst, const T* src, size_t n) { dst.assign(src ? src : nullptr, src ? src + n : nullptr); }
}  // namespace

extern "C" {

void mcba_default_options(mcba_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_num_iterations = 50; o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->huber_a = 1.0;
  o->jacobi_scaling = 1; o->max_num_consecutive_invalid_steps = 5; o->verbose = 0; o->materialize_jacobian = 0;
}

int mcba_create(void** handle, const mcba_problem* P, int) {
  Stub* S = new Stub();
  *handle = S;
  S->P = *P;
  copy(S->u, P->u, P->n_obs); copy(S->v, P->v, P->n_obs); copy(S->pt, P->pt_idx, P->n_obs);
  copy(S->ptr, P->bobs_ptr, P->n_bobs + 1); copy(S->cam, P->bobs_cam, P->n_bobs); copy(S->pose, P->bobs_pose, P->n_bobs);
  copy(S->board, P->bobs_board, P->n_bobs); copy(S->pts, P->pts_xyz, (size_t)P->n_pts * 3); copy(S->model, P->cam_model, P->n_cam);
  if (P->cam_is_ref) copy(S->cam_ref, P->cam_is_ref, P->n_cam);
  if (P->board_is_ref) copy(S->board_ref, P->board_is_ref, P->n_board);
  if (P->cam_problem) copy(S->cam_problem, P->cam_problem, P->n_cam);
  S->P.u = S->u.data(); S->P.v = S->v.data(); S->P.pt_idx = S->pt.data(); S->P.bobs_ptr = S->ptr.data();
  S->P.bobs_cam = S->cam.data(); S->P.bobs_pose = S->pose.data(); S->P.bobs_board = S->board.data();
  S->P.pts_xyz = S->pts.data(); S->P.cam_model = S->model.data();
  S->P.cam_is_ref = S->cam_ref.empty() ? nullptr : S->cam_ref.data();
  S->P.board_is_ref = S->board_ref.empty() ? nullptr : S->board_ref.data();
  S->P.cam_problem = nullptr; S->P.n_problems = 1;
  for (int64_t b = 1; b < P->n_bobs; ++b) if (P->bobs_pose[b] < P->bobs_pose[b - 1]) { S->err = "bobs not sorted by e-block"; return MCBA_ERR_INVALID; }
  return MCBA_OK;
}
int mcba_set_stage(void* h, int stage) { ((Stub*)h)->P.stage = stage; return MCBA_OK; }
void mcba_destroy(void* h) { delete (Stub*)h; }
const char* mcba_last_error(void* h) { return h ? ((Stub*)h)->err.c_str() : "null handle"; }

int mcba_solve(void* h, mcba_params* params, const mcba_options* opt, mcba_summary* summary, mcba_trace_row* trace, int cap) {
  Stub* S = (Stub*)h;
  mcba_options o = *opt; o.verbose = 0;
  return oracle_solve(&S->P, params, &o, summary, trace, cap, 1);
}

// one independent stage-1 problem per camera (the bobs are camera-major, one e-block per board observation)
int mcba_solve_batched(void* h, mcba_params* params, const mcba_options* opt, mcba_summary* summaries) {
  Stub* S = (Stub*)h;
  if (S->cam_problem.empty()) { S->err = "not a batched handle"; return MCBA_ERR_INVALID; }
  mcba_options o = *opt; o.verbose = 0;
  const mcba_problem& P = S->P;
  int64_t b0 = 0;
  for (int c = 0; c < P.n_cam; ++c) {
    int64_t b1 = b0;
    while (b1 < P.n_bobs && P.bobs_cam[b1] == c) ++b1;
    const int64_t nb = b1 - b0, o0 = P.bobs_ptr[b0], o1 = P.bobs_ptr[b1];
    std::vector<int64_t> ptr(nb + 1); std::vector<int32_t> cam(nb, 0), pose(nb), board(nb, 0);
    for (int64_t b = 0; b <= nb; ++b) ptr[b] = P.bobs_ptr[b0 + b] - o0;
    for (int64_t b = 0; b < nb; ++b) pose[b] = (int32_t)b;
    mcba_problem Q = P;
    Q.stage = 1; Q.n_cam = 1; Q.n_pose = (int32_t)nb; Q.n_board = 0; Q.n_obs = o1 - o0; Q.n_bobs = nb;
    Q.u = P.u + o0; Q.v = P.v + o0; Q.pt_idx = P.pt_idx + o0; Q.bobs_ptr = ptr.data(); Q.bobs_cam = cam.data();
    Q.bobs_pose = pose.data(); Q.bobs_board = board.data(); Q.cam_model = P.cam_model + c; Q.cam_is_ref = nullptr; Q.board_is_ref = nullptr;
    double zero6[6] = {0, 0, 0, 0, 0, 0};
    mcba_params X{zero6, params->pose + 6 * P.bobs_pose[b0], nullptr, params->intrinsics + 9 * c};
    if (nb > 0) { const int rc = oracle_solve(&Q, &X, &o, &summaries[c], nullptr, 0, 1); if (rc) return rc; }
    b0 = b1;
  }
  return MCBA_OK;
}

int mcba_reprojection_errors(void* h, const mcba_params* params, double* err) {
  Stub* S = (Stub*)h;
  mcba_options o; mcba_default_options(&o); o.huber_a = 0.0;   // trivial loss: raw residuals
  std::vector<double> r((size_t)S->P.n_obs * 2);
  double cost = 0, g = 0;
  const int rc = oracle_eval(&S->P, params, &o, &cost, r.data(), nullptr, &g);
  if (rc) return rc;
  for (int64_t i = 0; i < S->P.n_obs; ++i) err[i] = std::sqrt(r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1]);
  return MCBA_OK;
}

// the reference-rounded error path on the host: the same MCBA_HD functions the kernel runs (mcba_math.cuh), per corner
int mcba_reprojection_errors_f32(void* h, const mcba_params* params, float* err) {
  Stub* S = (Stub*)h; const mcba_problem& P = S->P;
  const bool has_cam = P.stage >= 3, has_board = P.stage == 2 || P.stage >= 4;
  for (int64_t b = 0; b < P.n_bobs; ++b) {
    const int cam = P.bobs_cam[b], o = P.bobs_pose[b], bd = has_board ? P.bobs_board[b] : 0;
    const bool act_c = has_cam && !(P.cam_is_ref && P.cam_is_ref[cam]);
    const bool act_b = has_board && !(P.board_is_ref && P.board_is_ref[bd]);
    double pc[mcba::PREP], po[mcba::PREP], pb[mcba::PREP], ctx[mcba::CTX_SIZE];
    double zero6[6] = {0, 0, 0, 0, 0, 0};
    mcba::prep_block(has_cam && params->cam_pose ? params->cam_pose + 6 * cam : zero6, pc);
    mcba::prep_block(params->pose + 6 * o, po);
    mcba::prep_block(has_board && params->board_pose ? params->board_pose + 6 * bd : zero6, pb);
    for (int e = 0; e < mcba::CTX_SIZE; ++e)
      if (e < mcba::CTX_NO || e >= mcba::CTX_IN) ctx[e] = mcba::ctx_entry_a(e, pc, act_c, po, pb, act_b, params->intrinsics + 9 * cam);
    for (int64_t i = P.bobs_ptr[b]; i < P.bobs_ptr[b + 1]; ++i) {
      const float* p3 = P.pts_xyz + 3 * (size_t)P.pt_idx[i];
      const double X0[3] = {p3[0], p3[1], p3[2]};
      err[i] = mcba::obs_error_ref(ctx, P.cam_model[cam], X0, P.u[i], P.v[i]);
    }
  }
  return MCBA_OK;
}

void mcba_pnp_default_options(mcba_pnp_options* o) {
  o->threshold_px = 10.0; o->max_iterations = 1000; o->confidence = 0.99; o->refine = 1; o->refine_max_iterations = 20; o->seed = 0;
}
const char* mcba_pnp_last_error(void) { return ""; }
int mcba_pnp_ransac(const mcba_pnp_problem* P, const mcba_pnp_options* opt, int, mcba_pnp_result* res) {
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    const int64_t o0 = P->bobs_ptr[b], n = P->bobs_ptr[b + 1] - o0;
    std::vector<float> xyz(3 * n);
    for (int64_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] = P->pts_xyz[3 * (size_t)P->pt_idx[o0 + i] + k];
    std::vector<unsigned char> mask(n > 0 ? n : 1);
    double pose0[6]; int bh = -1, nh = 0;
    const int cam = P->bobs_cam[b];
    res->n_inliers[b] = pnptest_ransac((int)n, P->u + o0, P->v + o0, xyz.data(), P->cam_model[cam], P->intrinsics + 9 * cam,
                                       opt->threshold_px, opt->max_iterations, opt->confidence, opt->refine, opt->refine_max_iterations,
                                       opt->seed, (unsigned long long)b, res->pose + 6 * b, pose0, mask.data(), &bh, &nh);
    for (int64_t i = 0; i < n; ++i) res->inlier_mask[o0 + i] = mask[i];
    if (res->best_hypothesis) res->best_hypothesis[b] = bh;
    if (res->n_hypotheses) res->n_hypotheses[b] = nh;
    if (res->unrefined_pose) std::memcpy(res->unrefined_pose + 6 * b, pose0, sizeof(pose0));
  }
  res->kernel_ms = 0; res->total_ms = 0; res->kernel_launches = 0;
  return MCBA_OK;
}

}  // extern "C"
