// Host-side harness around csrc/mcba_math.cuh (the arithmetic the CUDA kernels run), so that it can be
// checked against the oracle on a box without a GPU. TEST CODE: never linked into libmcba.
#include <vector>
#include "../../mc-calib_b200/csrc/mcba_math.cuh"
using namespace mcba;

template <int STAGE> struct ArraySink {
  double* J; double* r;
  void put(int col, double ju, double jv) {
    typedef Cols<STAGE> C;
    if (col == C::R0) { r[0] = ju; r[1] = jv; return; }
    const int k = C::to_ref(col);
    J[k] = ju; J[C::K + k] = jv;
  }
};

template <int STAGE>
static double run(int model, int flags, const double* cam, const double* pose, const double* board,
                  const double* intr, const double* xyz, const double* uv, double huber_a, double* r, double* J,
                  double* cost_only) {
  typedef StageTraits<STAGE> T;
  double pc[PREP], po[PREP], pb[PREP], ctx[CTX_SIZE];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  prep_block(cam ? cam : zero6, pc);
  prep_block(pose, po);
  prep_block(board ? board : zero6, pb);
  const bool act_c = T::CAM && !(flags & 1), act_b = T::BOARD && !(flags & 2);
  for (int e = 0; e < CTX_SIZE; ++e) if (e < CTX_NO || e >= CTX_IN) ctx[e] = ctx_entry_a(e, pc, act_c, po, pb, act_b, intr);
  for (int e = CTX_NO; e < CTX_IN; ++e) ctx[e] = ctx_entry_b(e, ctx, po, pb, act_b);
  ArraySink<STAGE> sink{J, r};
  const double hr = obs_jacobian<STAGE>(ctx, act_c, act_b, model, xyz, uv[0], uv[1], huber_a, sink);
  double ru, rv;
  *cost_only = obs_cost(ctx, model, xyz, uv[0], uv[1], huber_a, ru, rv);
  return hr;
}

extern "C" double mathtest_obs(int stage, int model, int flags, const double* cam, const double* pose,
                               const double* board, const double* intr, const double* xyz, const double* uv,
                               double huber_a, double* r, double* J, double* cost_only) {
  switch (stage) {
    case 1: return run<1>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 2: return run<2>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 3: return run<3>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 4: return run<4>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
    case 5: return run<5>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J, cost_only);
  }
  return -1;
}
extern "C" int mathtest_chol6(double* A, double* b) {
  if (!chol6(A)) return 0;
  chol6_fwd(A, b); chol6_bwd(A, b); return 1;
}

// ---- pose initialisation arithmetic (csrc/mcba_pnp_math.cuh) ----
#include "../../mc-calib_b200/csrc/mcba_pnp_math.cuh"
extern "C" int pnptest_undistort(int model, const double* intr, double u, double v, double* ab) {
  return undistort_pixel(model, intr, u, v, ab[0], ab[1]) ? 1 : 0;
}
// f: 3 unit bearings (row-major 3x3), P: 3 points; out R [4][9], t [4][3]; returns the number of solutions
extern "C" int pnptest_p3p(const double* f, const double* P, double* R, double* t) {
  double ff[3][3], PP[3][3];
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) { ff[i][k] = f[i * 3 + k]; PP[i][k] = P[i * 3 + k]; }
  P3PSolutions s;
  p3p_solve(ff, PP, s);
  for (int i = 0; i < s.n; ++i) { for (int k = 0; k < 9; ++k) R[i * 9 + k] = s.R[i][k]; for (int k = 0; k < 3; ++k) t[i * 3 + k] = s.t[i][k]; }
  return s.n;
}
extern "C" int pnptest_hypothesis(const double* ab, const double* X, const double* intr, const double* w4, double* rvec, double* t) {
  double a[4][2], x[4][3], R[9];
  for (int i = 0; i < 4; ++i) { a[i][0] = ab[2 * i]; a[i][1] = ab[2 * i + 1]; for (int k = 0; k < 3; ++k) x[i][k] = X[3 * i + k]; }
  if (!p3p_hypothesis(a, x, intr, w4, R, t)) return 0;
  rot_to_rvec(R, rvec);
  return 1;
}
extern "C" void pnptest_rvec(const double* R, double* rv) { rot_to_rvec(R, rv); }
extern "C" void pnptest_sample4(unsigned long long seed, unsigned long long b, unsigned long long h, int n, int* idx) {
  ransac_sample4(seed, b, h, n, idx);
}
// replays the stopping rule over a list of inlier counts; returns the index of the chosen hypothesis (-1: none),
// *n_used = hypotheses consumed
extern "C" int pnptest_replay(const int* counts, int n_counts, int total, int it, double p, int* n_used) {
  RansacState s; ransac_init(s, it);
  int h = 0;
  while (ransac_continue(s, it) && h < n_counts) { ransac_update(s, h, counts[h], total, p); ++h; }
  *n_used = h;
  return s.best_h;
}
extern "C" int pnptest_residual_jac(const double* pose, const double* intr, const double* X0, double u, double v,
                                    double* r, double* J) {
  double prep[PREP];
  prep_block(pose, prep);
  return pnp_residual_jac(prep, intr, X0, u, v, r, J, J + 6) ? 1 : 0;
}

// Sequential restatement of k_pnp_prepare + k_pnp_ransac (csrc/mcba_pnp.cu) for ONE observation run, on the same
// header arithmetic: lets the CPU-only suite compare the algorithm with the OpenCV-backed checker before a GPU is
// involved. Sums of the refinement run in corner order instead of the kernel's shuffle tree (rounding-level difference).
extern "C" int pnptest_ransac(int n, const float* u, const float* v, const float* xyz /*[n][3]*/, int model,
                              const double* intr, double thresh, int it, double p, int refine, int refine_it,
                              unsigned long long seed, unsigned long long b, double* pose, double* pose0,
                              unsigned char* mask, int* best_h, int* n_hyp) {
  std::vector<double> ab(2 * n), w(2 * n); std::vector<unsigned char> okv(n);
  double in[9];
  for (int k = 0; k < 9; ++k) in[k] = (k < 4 || model == 0) ? intr[k] : 0.0;
  for (int i = 0; i < n; ++i) {
    double x = 0, y = 0;
    okv[i] = undistort_pixel(model, intr, (double)u[i], (double)v[i], x, y) ? 1 : 0;
    ab[2 * i] = x; ab[2 * i + 1] = y;
    w[2 * i] = model == 0 ? (double)u[i] : x * intr[0] + intr[2];
    w[2 * i + 1] = model == 0 ? (double)v[i] : y * intr[1] + intr[3];
  }
  const double thr2 = thresh * thresh;
  auto inlier = [&](const double* R, const double* t, int i) {
    const double X0[3] = {(double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]};
    double X[3], pu, pv;
    mat3_vec_add(R, X0, t, X);
    if (!project_brown(in, X, pu, pv)) return false;
    const double du = pu - w[2 * i], dv = pv - w[2 * i + 1];
    return du * du + dv * dv < thr2;
  };
  RansacState st; ransac_init(st, it);
  double bR[9] = {0}, bt[3] = {0};
  if (n >= 4) {
    for (int h = 0; ransac_continue(st, it); ++h) {
      int idx[4];
      ransac_sample4(seed, b, (unsigned long long)h, n, idx);
      double ab4[4][2], X4[4][3], R[9], t[3];
      bool ok = true;
      for (int k = 0; k < 4; ++k) {
        ab4[k][0] = ab[2 * idx[k]]; ab4[k][1] = ab[2 * idx[k] + 1]; ok = ok && okv[idx[k]];
        for (int c = 0; c < 3; ++c) X4[k][c] = (double)xyz[3 * idx[k] + c];
      }
      const double w4[2] = {w[2 * idx[3]], w[2 * idx[3] + 1]};
      ok = ok && p3p_hypothesis(ab4, X4, in, w4, R, t);
      int count = 0;
      if (ok) for (int i = 0; i < n; ++i) count += (okv[i] && inlier(R, t, i)) ? 1 : 0;
      const int prev = st.best_h;
      ransac_update(st, h, count, n, p);
      if (st.best_h != prev) { for (int k = 0; k < 9; ++k) bR[k] = R[k]; for (int k = 0; k < 3; ++k) bt[k] = t[k]; }
    }
  }
  int cnt = 0;
  for (int i = 0; i < n; ++i) { mask[i] = (st.best_h >= 0 && okv[i] && inlier(bR, bt, i)) ? 1 : 0; cnt += mask[i]; }
  double x[6] = {0, 0, 0, 0, 0, 0};
  if (st.best_h >= 0) { rot_to_rvec(bR, x); x[3] = bt[0]; x[4] = bt[1]; x[5] = bt[2]; }
  for (int k = 0; k < 6; ++k) pose0[k] = x[k];
  auto normal_eq = [&](const double* q, double* H, double* g, double& cost) {
    double prep[PREP]; prep_block(q, prep);
    for (int k = 0; k < 21; ++k) H[k] = 0; for (int k = 0; k < 6; ++k) g[k] = 0; cost = 0;
    for (int i = 0; i < n; ++i) {
      if (!mask[i]) continue;
      const double X0[3] = {(double)xyz[3 * i], (double)xyz[3 * i + 1], (double)xyz[3 * i + 2]};
      double r[2], Ju[6], Jv[6];
      if (!pnp_residual_jac(prep, in, X0, w[2 * i], w[2 * i + 1], r, Ju, Jv)) { cost += 1e12; continue; }
      cost += r[0] * r[0] + r[1] * r[1];
      int k = 0;
      for (int a = 0; a < 6; ++a) { for (int c = a; c < 6; ++c) H[k++] += Ju[a] * Ju[c] + Jv[a] * Jv[c]; g[a] += Ju[a] * r[0] + Jv[a] * r[1]; }
    }
  };
  if (refine && st.best_h >= 0 && cnt >= 4) {
    PnpLm lm; lm.init(x);
    for (int iter = 0; iter <= refine_it && !lm.fin; ++iter) {
      double cand[6], H2[21], g2[6], cost2;
      lm.propose(cand);
      if (lm.fin) break;
      normal_eq(cand, H2, g2, cost2);
      lm.update(cand, H2, g2, cost2);
    }
    bool fin = true;
    for (int k = 0; k < 6; ++k) fin = fin && is_fin(lm.pose[k]);
    if (fin) for (int k = 0; k < 6; ++k) x[k] = lm.pose[k];
  }
  for (int k = 0; k < 6; ++k) pose[k] = x[k];
  *best_h = st.best_h; *n_hyp = st.countit;
  return cnt;
}

// ---- the product's trust-region bookkeeping (csrc/mcba_lm.cuh), driven from a list of per-iteration inputs ----
#include "../../mc-calib_b200/csrc/mcba_lm.cuh"
// steps: n_steps rows of {solver_ok, mcc, step_norm, cand_norm, cand_cost, accepted_cost, accepted_grad_max}
// (the last two are what the re-evaluation at an accepted point returns). Writes one mcba_trace_row per recorded
// iteration; returns their number; *termination = mcba_termination or -1 when the inputs ran out first.
extern "C" int lmtest_replay(const mcba_options* opt, double cost0, double grad0, double xnorm0, int n_steps,
                             const double* steps, mcba_trace_row* rows, int rows_cap, int* termination,
                             int* n_successful, int* n_unsuccessful, double* final_cost) {
  LmState s; lm_init(s, *opt);
  lm_iteration_zero(s, cost0, grad0, xnorm0);
  int n_rows = 0, k = 0;
  *termination = -1;
  for (;;) {
    mcba_trace_row rec;
    const int cont = lm_finalize(s, *opt, &rec);
    if (n_rows < rows_cap) rows[n_rows] = rec;
    ++n_rows;
    if (!cont) { *termination = s.termination; break; }
    if (k >= n_steps) break;
    const double* in = steps + 7 * k++;
    LmStep st; st.solver_ok = in[0] != 0.0; st.mcc = in[1]; st.step_norm = in[2]; st.cand_norm = in[3]; st.cand_cost = in[4];
    const int d = lm_decide(s, *opt, st);
    if (d == LM_TERMINATED) { *termination = s.termination; break; }
    if (d == LM_ACCEPT) lm_accepted_eval(s, *opt, in[5], in[6]);
  }
  *n_successful = s.num_successful; *n_unsuccessful = s.num_unsuccessful; *final_cost = s.x_cost;
  return n_rows;
}

// ---- fused observation kernels of round 2 (csrc/mcba_obs_fused.cuh): the two pieces that are plain arithmetic ----
// Copy-free context (CtxRaw): raw prep records (or the identity record for a block that is not refined) + intrinsics as
// the kernels copy them, the 63 products in the kernels' order, then obs_jacobian on that layout.
template <int STAGE>
static double run_raw(int model, int flags, const double* cam, const double* pose, const double* board,
                      const double* intr, const double* xyz, const double* uv, double huber_a, double* r, double* J) {
  typedef StageTraits<STAGE> T;
  double pc[PREP], po[PREP], pb[PREP], ident[PREP] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, ctx[CtxRaw::SIZE] = {0};
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  prep_block(cam ? cam : zero6, pc);
  prep_block(pose, po);
  prep_block(board ? board : zero6, pb);
  const bool act_c = T::CAM && !(flags & 1), act_b = T::BOARD && !(flags & 2);
  for (int e = 0; e < PREP; ++e) {
    ctx[CtxRaw::PC + e] = act_c ? pc[e] : ident[e];
    ctx[CtxRaw::PO + e] = po[e];
    ctx[CtxRaw::PB + e] = act_b ? pb[e] : ident[e];
  }
  for (int e = 0; e < 9; ++e) ctx[CtxRaw::IN + e] = intr[e];
  for (int e = 0; e < 63; ++e) {          // Mco (e < 9) before the Nb products that read it, as in the kernels' rounds
    int m, d, o;
    ctx_raw_operands(e, m, d, o);
    ctx[o] = ctx[m] * ctx[d] + ctx[m + 1] * ctx[d + 3] + ctx[m + 2] * ctx[d + 6];
  }
  ArraySink<STAGE> sink{J, r};
  return obs_jacobian<STAGE, ArraySink<STAGE>, CtxRaw>(ctx, act_c, act_b, model, xyz, uv[0], uv[1], huber_a, sink);
}
extern "C" double mathtest_obs_raw(int stage, int model, int flags, const double* cam, const double* pose,
                                   const double* board, const double* intr, const double* xyz, const double* uv,
                                   double huber_a, double* r, double* J) {
  switch (stage) {
    case 1: return run_raw<1>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J);
    case 2: return run_raw<2>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J);
    case 3: return run_raw<3>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J);
    case 4: return run_raw<4>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J);
    case 5: return run_raw<5>(model, flags, cam, pose, board, intr, xyz, uv, huber_a, r, J);
  }
  return -1;
}
// Group-pair cover of the triangle: how many (product, lane, entry) triples of the DMMA accumulators land on each entry
// of the compact record, plus, per product, the row / column groups; returns the record length (NCOMP).
template <int STAGE> static int cover_counts(int* count, int* n_products, int* groups) {
  typedef Cols<STAGE> C;
  typedef RectCover<(C::NC + 3) / 4> RCV;
  for (int i = 0; i < C::NCOMP; ++i) count[i] = 0;
  for (int r = 0; r < RCV::NR; ++r) {
    groups[4 * r] = RCV::r0(r); groups[4 * r + 1] = RCV::r1(r); groups[4 * r + 2] = RCV::c0(r); groups[4 * r + 3] = RCV::c1(r);
    for (int lane = 0; lane < 32; ++lane)
      for (int h = 0; h < 2; ++h) {
        const int ci = k1_store_index<STAGE>(r, lane, h);
        if (ci >= C::NCOMP) return -1;
        if (ci >= 0) count[ci]++;
      }
  }
  *n_products = RCV::NR;
  return C::NCOMP;
}
extern "C" int mathtest_cover(int stage, int* count, int* n_products, int* groups) {
  switch (stage) {
    case 1: return cover_counts<1>(count, n_products, groups);
    case 2: return cover_counts<2>(count, n_products, groups);
    case 3: return cover_counts<3>(count, n_products, groups);
    case 4: return cover_counts<4>(count, n_products, groups);
    case 5: return cover_counts<5>(count, n_products, groups);
  }
  return -1;
}
// Which (column a, column b) of J~ a record entry holds (inverse of Cols<STAGE>::cidx), for the test's bookkeeping
extern "C" int mathtest_cidx(int stage, int a, int b) {
  switch (stage) {
    case 1: return Cols<1>::cidx(a, b); case 2: return Cols<2>::cidx(a, b); case 3: return Cols<3>::cidx(a, b);
    case 4: return Cols<4>::cidx(a, b); case 5: return Cols<5>::cidx(a, b);
  }
  return -2;
}
