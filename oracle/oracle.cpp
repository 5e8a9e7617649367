// oracle.cpp — CPU oracle: TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle.h).
//
// Restates, without Ceres/Eigen/OpenCV:
//   * the five reprojection functors of /root/reference/McCalib/include/OptimizationCeres.h
//     (ReprojectionError :11-117, _3DObjRef :120-251, _CameraGroupRef :255-382,
//      _CameraGroupAndObjectRef :385-525, _CameraGroupAndObjectRefAndIntrinsics :528-665),
//     evaluated with forward-mode dual numbers as ceres::AutoDiffCostFunction does;
//   * ceres-solver 2.2.0 (Dockerfile:48) semantics the reference relies on through its five
//     ceres::Solve call sites (Camera.cpp:301, Object3D.cpp:223, CameraGroup.cpp:273,468,577):
//     AngleAxisRotatePoint, HuberLoss(1.0) + Corrector, ProgramEvaluator, Jacobi scaling,
//     LevenbergMarquardtStrategy, SchurEliminator (e-blocks = the per-frame poses that the
//     automatic SPARSE_SCHUR ordering eliminates), TrustRegionMinimizer with default options.
// PARITY UNPINNED by the reference itself (no golden vectors exist). Pinned by
// tests/ against mpmath, cv2 and scipy instead.
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual number with K partials (the role ceres::Jet<double,K> plays in the reference).
template <int K>
struct Dual {
  double a;
  double v[K];
  Dual() : a(0.0) { for (int i = 0; i < K; ++i) v[i] = 0.0; }
  explicit Dual(double c) : a(c) { for (int i = 0; i < K; ++i) v[i] = 0.0; }
  Dual(double c, int k) : a(c) { for (int i = 0; i < K; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int K> inline Dual<K> operator+(const Dual<K>& f, const Dual<K>& g) {
  Dual<K> h; h.a = f.a + g.a; for (int i = 0; i < K; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int K> inline Dual<K> operator-(const Dual<K>& f, const Dual<K>& g) {
  Dual<K> h; h.a = f.a - g.a; for (int i = 0; i < K; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int K> inline Dual<K> operator*(const Dual<K>& f, const Dual<K>& g) {
  Dual<K> h; h.a = f.a * g.a; for (int i = 0; i < K; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int K> inline Dual<K> operator/(const Dual<K>& f, const Dual<K>& g) {
  // [recalled] ceres/jet.h: f/g = (f.a/g.a, (f.v - (f.a/g.a) g.v) / g.a)
  Dual<K> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < K; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int K> inline Dual<K>& operator+=(Dual<K>& f, const Dual<K>& g) { f = f + g; return f; }
template <int K> inline Dual<K>& operator/=(Dual<K>& f, const Dual<K>& g) { f = f / g; return f; }
template <int K> inline bool operator>(const Dual<K>& f, const Dual<K>& g) { return f.a > g.a; }
template <int K> inline Dual<K> sqrt(const Dual<K>& f) {
  Dual<K> h; const double t = std::sqrt(f.a); const double d = 1.0 / (2.0 * t); h.a = t;
  for (int i = 0; i < K; ++i) h.v[i] = f.v[i] * d; return h; }
template <int K> inline Dual<K> atan(const Dual<K>& f) {
  Dual<K> h; const double d = 1.0 / (1.0 + f.a * f.a); h.a = std::atan(f.a);
  for (int i = 0; i < K; ++i) h.v[i] = f.v[i] * d; return h; }
template <int K> inline Dual<K> sin(const Dual<K>& f) {
  Dual<K> h; const double c = std::cos(f.a); h.a = std::sin(f.a);
  for (int i = 0; i < K; ++i) h.v[i] = f.v[i] * c; return h; }
template <int K> inline Dual<K> cos(const Dual<K>& f) {
  Dual<K> h; const double s = -std::sin(f.a); h.a = std::cos(f.a);
  for (int i = 0; i < K; ++i) h.v[i] = f.v[i] * s; return h; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double atan(double x) { return std::atan(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
template <typename T> struct Lift { static T c(double x) { return T(x); } };
template <> struct Lift<double> { static double c(double x) { return x; } };

// [recalled] ceres/rotation.h AngleAxisRotatePoint: Rodrigues' formula when theta^2 > epsilon,
// first-order Taylor expansion (pt + w x pt) otherwise so that derivatives stay finite at 0.
template <typename T>
inline void angle_axis_rotate(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > Lift<T>::c(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T ct = cos(theta);
    const T st = sin(theta);
    const T ti = Lift<T>::c(1.0) / theta;
    const T w[3] = {aa[0] * ti, aa[1] * ti, aa[2] * ti};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (Lift<T>::c(1.0) - ct);
    out[0] = pt[0] * ct + wxp[0] * st + w[0] * tmp;
    out[1] = pt[1] * ct + wxp[1] * st + w[1] * tmp;
    out[2] = pt[2] * ct + wxp[2] * st + w[2] * tmp;
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + wxp[0];
    out[1] = pt[1] + wxp[1];
    out[2] = pt[2] + wxp[2];
  }
}

// ---------------------------------------------------------------------------------------------
// Which parameter blocks a stage has and where they sit in the residual block's argument list.
// Argument order at the reference call sites: stage 1 (pose, intr) Camera.cpp:290; stage 2
// (obs pose, board) Object3D.cpp:208-209; stage 3 (cam, obj) CameraGroup.cpp:253-256; stage 4
// (cam, obj, board) :445-451; stage 5 (cam, obj, board, intr) :553-560.
struct StageInfo {
  bool has_cam, has_board, has_intr;
  int K, off_cam, off_pose, off_board, off_intr;
};
inline StageInfo stage_info(int stage) {
  switch (stage) {
    case 1: return {false, false, true, 15, -1, 0, -1, 6};
    case 2: return {false, true, false, 12, -1, 0, 6, -1};
    case 3: return {true, false, false, 12, 0, 6, -1, -1};
    case 4: return {true, true, false, 18, 0, 6, 12, -1};
    case 5: return {true, true, true, 27, 0, 6, 12, 18};
    default: return {false, false, false, 0, -1, -1, -1, -1};
  }
}

// One reprojection residual. The chain is the union of the five functors: absent blocks are the
// identity. board (if refined) -> pose (always) -> camera (if refined) -> normalise -> distort
// -> pixels. Follows OptimizationCeres.h:401-409 (board), :413-417 (object/board-obs pose),
// :420-427 (camera), :430-431 (normalisation, no z>0 guard), :436-443 (Brown), :460-477
// (Kannala incl. the r > 1e-8 guard), :480-486 (pixels). Intrinsics: [fx fy u0 v0 d4..d8] with
// Brown (k1,k2,p1,p2,k3) = d4,d5,d6,d7,d8 (:541-549, :37-41) and Kannala (k1..k4) = d4..d7
// (:67-70; the fixed-intrinsics functors pass them as k1,k2,p1,p2, :471).
template <typename T>
inline void reproject(const T* cam, bool use_cam, const T* pose, const T* board, bool use_board,
                      const T* in, int model, const double xyz[3], const double uv[2], T res[2]) {
  T p[3] = {Lift<T>::c(xyz[0]), Lift<T>::c(xyz[1]), Lift<T>::c(xyz[2])};
  T q[3];
  if (use_board) {
    angle_axis_rotate(board, p, q);
    p[0] = q[0] + board[3]; p[1] = q[1] + board[4]; p[2] = q[2] + board[5];
  }
  angle_axis_rotate(pose, p, q);
  p[0] = q[0] + pose[3]; p[1] = q[1] + pose[4]; p[2] = q[2] + pose[5];
  if (use_cam) {
    angle_axis_rotate(cam, p, q);
    p[0] = q[0] + cam[3]; p[1] = q[1] + cam[4]; p[2] = q[2] + cam[5];
  }
  p[0] /= p[2];
  p[1] /= p[2];
  const T one = Lift<T>::c(1.0), two = Lift<T>::c(2.0);
  T xd, yd;
  if (model == 0) {
    const T k1 = in[4], k2 = in[5], k3 = in[8], p1 = in[6], p2 = in[7];
    const T r2 = p[0] * p[0] + p[1] * p[1];
    const T r4 = r2 * r2;
    const T r6 = r4 * r2;
    const T rc = (one + k1 * r2 + k2 * r4 + k3 * r6);
    xd = p[0] * rc + two * p1 * p[0] * p[1] + p2 * (r2 + two * p[0] * p[0]);
    yd = p[1] * rc + p1 * (r2 + two * (p[1] * p[1])) + two * p2 * p[0] * p[1];
  } else {
    const T k1 = in[4], k2 = in[5], k3 = in[6], k4 = in[7];
    const T r2 = p[0] * p[0] + p[1] * p[1];
    const T r = sqrt(r2);
    const T th = atan(r);
    const T th2 = th * th, th3 = th2 * th, th4 = th2 * th2, th5 = th4 * th;
    const T th6 = th3 * th3, th7 = th6 * th, th8 = th4 * th4, th9 = th8 * th;
    const T thd = th + k1 * th3 + k2 * th5 + k3 * th7 + k4 * th9;
    const bool far = r > Lift<T>::c(1e-8);
    const T inv_r = far ? one / r : one;
    const T cd = far ? thd * inv_r : one;
    xd = p[0] * cd;
    yd = p[1] * cd;
  }
  res[0] = in[0] * xd + in[2] - Lift<T>::c(uv[0]);
  res[1] = in[1] * yd + in[3] - Lift<T>::c(uv[1]);
}

// Autodiff evaluation of one residual block: value + 2xK Jacobian in the reference argument order.
template <int STAGE>
inline void functor_autodiff(const double* cam, bool cam_ref, const double* pose, const double* board,
                             bool board_ref, const double* intr, int model, const double xyz[3],
                             const double uv[2], double r[2], double* J /*[2][K]*/) {
  const StageInfo si = stage_info(STAGE);
  constexpr int K = (STAGE == 1) ? 15 : (STAGE == 2 || STAGE == 3) ? 12 : (STAGE == 4) ? 18 : 27;
  typedef Dual<K> D;
  D dcam[6], dpose[6], dboard[6], din[9];
  for (int i = 0; i < 6; ++i) {
    dpose[i] = D(pose[i], si.off_pose + i);
    if (si.has_cam) dcam[i] = D(cam[i], si.off_cam + i);
    if (si.has_board) dboard[i] = D(board[i], si.off_board + i);
  }
  for (int i = 0; i < 9; ++i) din[i] = si.has_intr ? D(intr[i], si.off_intr + i) : D(intr[i]);
  D res[2];
  reproject<D>(dcam, si.has_cam && !cam_ref, dpose, dboard, si.has_board && !board_ref, din, model, xyz, uv, res);
  r[0] = res[0].a; r[1] = res[1].a;
  for (int k = 0; k < K; ++k) { J[k] = res[0].v[k]; J[K + k] = res[1].v[k]; }
}

typedef void (*functor_fn)(const double*, bool, const double*, const double*, bool, const double*, int,
                           const double*, const double*, double*, double*);
inline functor_fn functor_for(int stage) {
  switch (stage) {
    case 1: return &functor_autodiff<1>;
    case 2: return &functor_autodiff<2>;
    case 3: return &functor_autodiff<3>;
    case 4: return &functor_autodiff<4>;
    case 5: return &functor_autodiff<5>;
  }
  return nullptr;
}

// [recalled] ceres::HuberLoss::Evaluate (loss_function.cc): b = a^2.
inline void huber(double a, double s, double rho[3]) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

inline bool finite_all(const double* p, int n) {
  for (int i = 0; i < n; ++i) if (!std::isfinite(p[i])) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
struct Layout {   // state vector: e-blocks first, then f-blocks (Ceres puts the eliminated blocks first)
  StageInfo si;
  int Wc, Wb, n_e, n_f, n_x;
  int cam_col(int c) const { return c * Wc; }                 // within f: [pose(6)] [intr(9)]
  int board_col(int b, int n_cam) const { return n_cam * Wc + b * Wb; }
};
inline Layout make_layout(const mcba_problem* P) {
  Layout L; L.si = stage_info(P->stage);
  L.Wc = (L.si.has_cam ? 6 : 0) + (L.si.has_intr ? 9 : 0);
  L.Wb = L.si.has_board ? 6 : 0;
  L.n_e = P->n_pose * 6;
  L.n_f = P->n_cam * L.Wc + P->n_board * L.Wb;
  L.n_x = L.n_e + L.n_f;
  return L;
}

struct Work {
  const mcba_problem* P;
  Layout L;
  mcba_options opt;
  int n_threads;
  std::vector<int32_t> obs_bobs;       // obs -> bobs
  std::vector<int64_t> pose_bobs_ptr;  // e-block -> bobs range
  std::vector<double> x, cand, res, jac, cost_obs, grad, scale, diag;
  // parameter views used by evaluation
  std::vector<double> cam_pose, pose, board_pose, intr;
};

inline void unpack_x(const Work& W, const std::vector<double>& x, std::vector<double>& cam_pose,
                     std::vector<double>& pose, std::vector<double>& board_pose, std::vector<double>& intr) {
  const mcba_problem* P = W.P; const Layout& L = W.L;
  for (int i = 0; i < L.n_e; ++i) pose[i] = x[i];
  for (int c = 0; c < P->n_cam; ++c) {
    const double* f = &x[L.n_e + L.cam_col(c)];
    int o = 0;
    if (L.si.has_cam) { for (int i = 0; i < 6; ++i) cam_pose[c * 6 + i] = f[i]; o = 6; }
    if (L.si.has_intr) for (int i = 0; i < 9; ++i) intr[c * 9 + i] = f[o + i];
  }
  if (L.si.has_board)
    for (int b = 0; b < P->n_board; ++b)
      for (int i = 0; i < 6; ++i) board_pose[b * 6 + i] = x[L.n_e + L.board_col(b, P->n_cam) + i];
}

// ProgramEvaluator::Evaluate [recalled]: per residual block evaluate the cost function, apply the
// loss Corrector (Jacobian first, then residuals), cost += 0.5 rho(s). Returns false on non-finite output.
bool evaluate(Work& W, const std::vector<double>& x, bool want_jac, double* cost_out) {
  const mcba_problem* P = W.P; const Layout& L = W.L; const int K = L.si.K;
  unpack_x(W, x, W.cam_pose, W.pose, W.board_pose, W.intr);
  functor_fn fn = functor_for(P->stage);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  bool ok = true;
  const double a = W.opt.huber_a;
#pragma omp parallel for schedule(static) num_threads(W.n_threads)
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    const int c = P->bobs_cam[b], o = P->bobs_pose[b];
    const int bd = L.si.has_board ? P->bobs_board[b] : 0;
    const double* cam = L.si.has_cam ? &W.cam_pose[c * 6] : zero6;
    const double* brd = L.si.has_board ? &W.board_pose[bd * 6] : zero6;
    const bool cam_ref = L.si.has_cam && P->cam_is_ref && P->cam_is_ref[c];
    const bool brd_ref = L.si.has_board && P->board_is_ref && P->board_is_ref[bd];
    double Jloc[2 * 27];
    for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
      const float* pt = &P->pts_xyz[3 * (int64_t)P->pt_idx[i]];
      const double xyz[3] = {(double)pt[0], (double)pt[1], (double)pt[2]};
      const double uv[2] = {(double)P->u[i], (double)P->v[i]};
      double r[2];
      if (want_jac) {
        fn(cam, cam_ref, &W.pose[o * 6], brd, brd_ref, &W.intr[c * 9], P->cam_model[c], xyz, uv, r, Jloc);
      } else {   // cost-only evaluation runs the functor on plain doubles, as Ceres does
        reproject<double>(cam, L.si.has_cam && !cam_ref, &W.pose[o * 6], brd, L.si.has_board && !brd_ref,
                          &W.intr[c * 9], P->cam_model[c], xyz, uv, r);
      }
      if (!finite_all(r, 2) || (want_jac && !finite_all(Jloc, 2 * K))) {
#pragma omp atomic write
        ok = false;
      }
      const double s = r[0] * r[0] + r[1] * r[1];
      double rho[3] = {s, 1.0, 0.0};
      if (a > 0) huber(a, s, rho);
      // Corrector [recalled]: rho'' <= 0 for Huber, so both are scaled by sqrt(rho') only.
      const double sr = std::sqrt(rho[1]);
      if (want_jac) {
        double* J = &W.jac[(size_t)i * 2 * K];
        for (int k = 0; k < 2 * K; ++k) J[k] = sr * Jloc[k];
        W.res[2 * i] = sr * r[0];
        W.res[2 * i + 1] = sr * r[1];
      }
      W.cost_obs[i] = 0.5 * rho[0];
    }
  }
  double cost = 0.0;   // single-thread Ceres sums residual blocks in problem order
  for (int64_t i = 0; i < P->n_obs; ++i) cost += W.cost_obs[i];
  *cost_out = cost;
  return ok && std::isfinite(cost);
}

// Maps the local Jacobian columns of a bobs to state-vector columns.
struct ColMap { int e_col; int nf; int fcol[21]; int floc[21]; };
inline ColMap col_map(const Work& W, int64_t b) {
  const mcba_problem* P = W.P; const Layout& L = W.L; ColMap m; m.nf = 0;
  m.e_col = P->bobs_pose[b] * 6;
  const int c = P->bobs_cam[b];
  if (L.si.has_cam) for (int i = 0; i < 6; ++i) { m.fcol[m.nf] = L.cam_col(c) + i; m.floc[m.nf++] = L.si.off_cam + i; }
  if (L.si.has_intr) for (int i = 0; i < 9; ++i) { m.fcol[m.nf] = L.cam_col(c) + (L.si.has_cam ? 6 : 0) + i; m.floc[m.nf++] = L.si.off_intr + i; }
  if (L.si.has_board) for (int i = 0; i < 6; ++i) { m.fcol[m.nf] = L.board_col(P->bobs_board[b], P->n_cam) + i; m.floc[m.nf++] = L.si.off_board + i; }
  return m;
}

void gradient_and_colnorm(Work& W, std::vector<double>* grad, std::vector<double>* sqnorm) {
  const mcba_problem* P = W.P; const Layout& L = W.L; const int K = L.si.K;
  if (grad) grad->assign(L.n_x, 0.0);
  if (sqnorm) sqnorm->assign(L.n_x, 0.0);
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    const ColMap m = col_map(W, b);
    for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
      const double* J = &W.jac[(size_t)i * 2 * K];
      for (int r = 0; r < 2; ++r) {
        const double rr = W.res[2 * i + r];
        for (int k = 0; k < 6; ++k) {
          const double j = J[r * K + L.si.off_pose + k];
          if (grad) (*grad)[m.e_col + k] += j * rr;
          if (sqnorm) (*sqnorm)[m.e_col + k] += j * j;
        }
        for (int k = 0; k < m.nf; ++k) {
          const double j = J[r * K + m.floc[k]];
          if (grad) (*grad)[L.n_e + m.fcol[k]] += j * rr;
          if (sqnorm) (*sqnorm)[L.n_e + m.fcol[k]] += j * j;
        }
      }
    }
  }
}

void scale_columns(Work& W, const std::vector<double>& s) {
  const mcba_problem* P = W.P; const Layout& L = W.L; const int K = L.si.K;
#pragma omp parallel for schedule(static) num_threads(W.n_threads)
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    const ColMap m = col_map(W, b);
    for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
      double* J = &W.jac[(size_t)i * 2 * K];
      for (int r = 0; r < 2; ++r) {
        for (int k = 0; k < 6; ++k) J[r * K + L.si.off_pose + k] *= s[m.e_col + k];
        for (int k = 0; k < m.nf; ++k) J[r * K + m.floc[k]] *= s[L.n_e + m.fcol[k]];
      }
    }
  }
}

// In-place lower Cholesky of a dense symmetric matrix (row-major, lower part used). false if not PD.
bool cholesky(double* A, int n) {
  for (int i = 0; i < n; ++i) {
    double* Li = A + (size_t)i * n;
    for (int j = 0; j <= i; ++j) {
      const double* Lj = A + (size_t)j * n;
      double s = Li[j];
      for (int k = 0; k < j; ++k) s -= Li[k] * Lj[k];
      if (i == j) {
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        Li[j] = std::sqrt(s);
      } else {
        Li[j] = s / Lj[j];
      }
    }
  }
  return true;
}
void cholesky_solve(const double* Lm, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i]; const double* Li = Lm + (size_t)i * n;
    for (int k = 0; k < i; ++k) s -= Li[k] * b[k];
    b[i] = s / Li[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= Lm[(size_t)k * n + i] * b[k];
    b[i] = s / Lm[(size_t)i * n + i];
  }
}

// SchurComplementSolver::SolveImpl [recalled]: eliminate the e-blocks chunk by chunk
// (ete = sum E^T E + D_e^2, inverse by LLT; S = F^T F + D_f^2 - sum (E^T F)^T ete^-1 (E^T F);
// rhs = F^T b - sum (E^T F)^T ete^-1 E^T b), factor the reduced system, back-substitute.
// Solves (J^T J + D^2) y = J^T r for the (already column-scaled) stored Jacobian.
bool schur_solve(Work& W, const std::vector<double>& D, std::vector<double>& y) {
  const mcba_problem* P = W.P; const Layout& L = W.L; const int K = L.si.K;
  const int nf = L.n_f, ne_blocks = P->n_pose;
  const int T = std::max(1, W.n_threads);
  std::vector<std::vector<double>> S_t(T), rhs_t(T);
  std::vector<double> inv_ete((size_t)ne_blocks * 36), ge((size_t)ne_blocks * 6);
  bool ok = true;
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    std::vector<double>& S = S_t[tid]; std::vector<double>& rhs = rhs_t[tid];
    S.assign((size_t)nf * nf, 0.0); rhs.assign(nf, 0.0);
    std::vector<double> etf((size_t)6 * nf, 0.0);
    std::vector<int> seg_start, seg_len; std::vector<char> mark(nf + 1, 0);
#pragma omp for schedule(dynamic, 16)
    for (int o = 0; o < ne_blocks; ++o) {
      double ete[36] = {0}, g[6] = {0};
      seg_start.clear(); seg_len.clear();
      for (int64_t b = W.pose_bobs_ptr[o]; b < W.pose_bobs_ptr[o + 1]; ++b) {
        const ColMap m = col_map(W, b);
        // touched f segments of this chunk
        if (L.Wc) { int s0 = L.cam_col(P->bobs_cam[b]); if (!mark[s0]) { mark[s0] = 1; seg_start.push_back(s0); seg_len.push_back(L.Wc); } }
        if (L.Wb) { int s0 = L.board_col(P->bobs_board[b], P->n_cam); if (!mark[s0]) { mark[s0] = 1; seg_start.push_back(s0); seg_len.push_back(L.Wb); } }
        for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
          const double* J = &W.jac[(size_t)i * 2 * K];
          for (int r = 0; r < 2; ++r) {
            const double* Jr = J + r * K; const double* E = Jr + L.si.off_pose; const double rr = W.res[2 * i + r];
            double fv[21];
            for (int k = 0; k < m.nf; ++k) fv[k] = Jr[m.floc[k]];
            for (int p = 0; p < 6; ++p) {
              g[p] += E[p] * rr;
              for (int q = 0; q < 6; ++q) ete[p * 6 + q] += E[p] * E[q];
              double* etfp = &etf[(size_t)p * nf];
              for (int k = 0; k < m.nf; ++k) etfp[m.fcol[k]] += E[p] * fv[k];
            }
            for (int k = 0; k < m.nf; ++k) {
              rhs[m.fcol[k]] += fv[k] * rr;
              double* Sk = &S[(size_t)m.fcol[k] * nf];
              for (int l = 0; l < m.nf; ++l) Sk[m.fcol[l]] += fv[k] * fv[l];
            }
          }
        }
      }
      for (int p = 0; p < 6; ++p) ete[p * 6 + p] += D[o * 6 + p] * D[o * 6 + p];
      // inverse via LLT (Eigen selfadjointView.llt().solve(I) in Ceres' InvertPSDMatrix)
      double Lc[36]; std::memcpy(Lc, ete, sizeof(Lc));
      double* inv = &inv_ete[(size_t)o * 36];
      if (!cholesky(Lc, 6)) {
#pragma omp atomic write
        ok = false;
        for (int k = 0; k < 36; ++k) inv[k] = 0.0;
      } else {
        for (int c = 0; c < 6; ++c) {
          double e[6] = {0, 0, 0, 0, 0, 0}; e[c] = 1.0; cholesky_solve(Lc, 6, e);
          for (int r = 0; r < 6; ++r) inv[r * 6 + c] = e[r];
        }
      }
      for (int p = 0; p < 6; ++p) ge[(size_t)o * 6 + p] = g[p];
      double ig[6];
      for (int p = 0; p < 6; ++p) { double s = 0; for (int q = 0; q < 6; ++q) s += inv[p * 6 + q] * g[q]; ig[p] = s; }
      const int ns = (int)seg_start.size();
      std::vector<double> tmp;  // inv_ete * etf on touched segments, computed per segment pair
      for (int a = 0; a < ns; ++a) {
        const int sa = seg_start[a], la = seg_len[a];
        for (int i = 0; i < la; ++i) {
          double s = 0; for (int p = 0; p < 6; ++p) s += etf[(size_t)p * nf + sa + i] * ig[p];
          rhs[sa + i] -= s;
        }
        // w = inv_ete * etf[:, seg a]  (6 x la)
        double w[6 * 15];
        for (int p = 0; p < 6; ++p) for (int i = 0; i < la; ++i) {
          double s = 0; for (int q = 0; q < 6; ++q) s += inv[p * 6 + q] * etf[(size_t)q * nf + sa + i];
          w[p * la + i] = s;
        }
        for (int b2 = 0; b2 < ns; ++b2) {
          const int sb = seg_start[b2], lb = seg_len[b2];
          for (int i = 0; i < la; ++i) {
            double* Srow = &S[(size_t)(sa + i) * nf + sb];
            for (int j = 0; j < lb; ++j) {
              double s = 0; for (int p = 0; p < 6; ++p) s += w[p * la + i] * etf[(size_t)p * nf + sb + j];
              Srow[j] -= s;
            }
          }
        }
      }
      for (int a = 0; a < ns; ++a) {
        mark[seg_start[a]] = 0;
        for (int p = 0; p < 6; ++p) std::memset(&etf[(size_t)p * nf + seg_start[a]], 0, sizeof(double) * seg_len[a]);
      }
    }
  }
  if (!ok) return false;
  std::vector<double>& S = S_t[0]; std::vector<double>& rhs = rhs_t[0];
  for (int t = 1; t < T; ++t) {
    if (S_t[t].empty()) continue;
    for (size_t k = 0; k < S.size(); ++k) S[k] += S_t[t][k];
    for (int k = 0; k < nf; ++k) rhs[k] += rhs_t[t][k];
  }
  for (int k = 0; k < nf; ++k) S[(size_t)k * nf + k] += D[L.n_e + k] * D[L.n_e + k];
  std::vector<double> z(rhs);
  if (nf > 0) {
    if (!cholesky(S.data(), nf)) return false;
    cholesky_solve(S.data(), nf, z.data());
  }
  // BackSubstitute: y_e = ete^-1 (E^T b - E^T F z)
  y.assign(L.n_x, 0.0);
  for (int k = 0; k < nf; ++k) y[L.n_e + k] = z[k];
#pragma omp parallel for schedule(static) num_threads(T)
  for (int o = 0; o < ne_blocks; ++o) {
    double g[6]; for (int p = 0; p < 6; ++p) g[p] = ge[(size_t)o * 6 + p];
    for (int64_t b = W.pose_bobs_ptr[o]; b < W.pose_bobs_ptr[o + 1]; ++b) {
      const ColMap m = col_map(W, b);
      for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
        const double* J = &W.jac[(size_t)i * 2 * K];
        for (int r = 0; r < 2; ++r) {
          const double* Jr = J + r * K; double fz = 0;
          for (int k = 0; k < m.nf; ++k) fz += Jr[m.floc[k]] * z[m.fcol[k]];
          for (int p = 0; p < 6; ++p) g[p] -= Jr[L.si.off_pose + p] * fz;
        }
      }
    }
    const double* inv = &inv_ete[(size_t)o * 36];
    for (int p = 0; p < 6; ++p) { double s = 0; for (int q = 0; q < 6; ++q) s += inv[p * 6 + q] * g[q]; y[o * 6 + p] = s; }
  }
  return finite_all(y.data(), L.n_x);
}

// model_cost_change = -(J step)^T (r + J step / 2)   (TrustRegionMinimizer::ComputeTrustRegionStep)
double model_cost_change(const Work& W, const std::vector<double>& step) {
  const mcba_problem* P = W.P; const Layout& L = W.L; const int K = L.si.K;
  double acc = 0.0;
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    const ColMap m = col_map(W, b);
    for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) {
      const double* J = &W.jac[(size_t)i * 2 * K];
      for (int r = 0; r < 2; ++r) {
        const double* Jr = J + r * K; double mr = 0;
        for (int k = 0; k < 6; ++k) mr += Jr[L.si.off_pose + k] * step[m.e_col + k];
        for (int k = 0; k < m.nf; ++k) mr += Jr[m.floc[k]] * step[L.n_e + m.fcol[k]];
        acc += mr * (W.res[2 * i + r] + mr / 2.0);
      }
    }
  }
  return -acc;
}

inline double norm2(const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); }

bool init_work(Work& W, const mcba_problem* P, const mcba_params* prm, const mcba_options* opt, int n_threads) {
  W.P = P; W.L = make_layout(P); W.opt = *opt;
  if (W.L.si.K == 0) return false;
  W.n_threads = std::max(1, n_threads);
  const Layout& L = W.L;
  W.obs_bobs.resize(P->n_obs);
  W.pose_bobs_ptr.assign(P->n_pose + 1, 0);
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    if (b > 0 && P->bobs_pose[b] < P->bobs_pose[b - 1]) return false;
    W.pose_bobs_ptr[P->bobs_pose[b] + 1] = b + 1;
    for (int64_t i = P->bobs_ptr[b]; i < P->bobs_ptr[b + 1]; ++i) W.obs_bobs[i] = (int32_t)b;
  }
  for (int o = 0; o < P->n_pose; ++o) W.pose_bobs_ptr[o + 1] = std::max(W.pose_bobs_ptr[o + 1], W.pose_bobs_ptr[o]);
  W.x.assign(L.n_x, 0.0);
  for (int i = 0; i < L.n_e; ++i) W.x[i] = prm->pose[i];
  for (int c = 0; c < P->n_cam; ++c) {
    double* f = &W.x[L.n_e + L.cam_col(c)]; int o = 0;
    if (L.si.has_cam) { for (int i = 0; i < 6; ++i) f[i] = prm->cam_pose[c * 6 + i]; o = 6; }
    if (L.si.has_intr) for (int i = 0; i < 9; ++i) f[o + i] = prm->intrinsics[c * 9 + i];
  }
  if (L.si.has_board) for (int b = 0; b < P->n_board; ++b) for (int i = 0; i < 6; ++i)
    W.x[L.n_e + L.board_col(b, P->n_cam) + i] = prm->board_pose[b * 6 + i];
  W.cam_pose.assign((size_t)P->n_cam * 6, 0.0); W.pose.assign((size_t)P->n_pose * 6, 0.0);
  W.board_pose.assign((size_t)std::max(1, P->n_board) * 6, 0.0); W.intr.assign((size_t)P->n_cam * 9, 0.0);
  for (int i = 0; i < P->n_cam * 9; ++i) W.intr[i] = prm->intrinsics[i];
  W.res.assign((size_t)P->n_obs * 2, 0.0); W.jac.assign((size_t)P->n_obs * 2 * L.si.K, 0.0);
  W.cost_obs.assign(P->n_obs, 0.0);
  return true;
}

void write_back(const Work& W, const std::vector<double>& x, mcba_params* prm) {
  const mcba_problem* P = W.P; const Layout& L = W.L;
  for (int i = 0; i < L.n_e; ++i) prm->pose[i] = x[i];
  for (int c = 0; c < P->n_cam; ++c) {
    const double* f = &x[L.n_e + L.cam_col(c)]; int o = 0;
    if (L.si.has_cam) { for (int i = 0; i < 6; ++i) prm->cam_pose[c * 6 + i] = f[i]; o = 6; }
    if (L.si.has_intr) for (int i = 0; i < 9; ++i) prm->intrinsics[c * 9 + i] = f[o + i];
  }
  if (L.si.has_board) for (int b = 0; b < P->n_board; ++b) for (int i = 0; i < 6; ++i)
    prm->board_pose[b * 6 + i] = x[L.n_e + L.board_col(b, P->n_cam) + i];
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_huber(double a, double s, double* rho) { huber(a, s, rho); }

void oracle_angle_axis_rotate(const double* aa, const double* pt, double* result, double* d_aa, double* d_pt) {
  typedef Dual<6> D;
  D a[3], p[3], o[3];
  for (int i = 0; i < 3; ++i) { a[i] = D(aa[i], i); p[i] = D(pt[i], 3 + i); }
  angle_axis_rotate<D>(a, p, o);
  for (int i = 0; i < 3; ++i) {
    result[i] = o[i].a;
    for (int k = 0; k < 3; ++k) { if (d_aa) d_aa[i * 3 + k] = o[i].v[k]; if (d_pt) d_pt[i * 3 + k] = o[i].v[3 + k]; }
  }
}

int oracle_functor(int stage, int model, int flags, const double* cam_pose, const double* pose,
                   const double* board_pose, const double* intr, const double* xyz, const double* uv,
                   double* residual, double* jacobian) {
  functor_fn fn = functor_for(stage);
  if (!fn) return MCBA_ERR_INVALID;
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  fn(cam_pose ? cam_pose : zero6, (flags & 1) != 0, pose, board_pose ? board_pose : zero6, (flags & 2) != 0,
     intr, model, xyz, uv, residual, jacobian);
  return MCBA_OK;
}

int oracle_eval(const mcba_problem* prob, const mcba_params* params, const mcba_options* opt, double* cost,
                double* residuals, double* jacobian, double* gradient_max_norm) {
  Work W;
  if (!init_work(W, prob, params, opt, oracle_max_threads())) return MCBA_ERR_INVALID;
  double c = 0;
  const bool ok = evaluate(W, W.x, true, &c);
  if (cost) *cost = c;
  if (residuals) std::memcpy(residuals, W.res.data(), sizeof(double) * W.res.size());
  if (jacobian) std::memcpy(jacobian, W.jac.data(), sizeof(double) * W.jac.size());
  if (gradient_max_norm) {
    std::vector<double> g; gradient_and_colnorm(W, &g, nullptr);
    double m = 0; for (double e : g) m = std::max(m, std::fabs(e)); *gradient_max_norm = m;
  }
  return ok ? MCBA_OK : MCBA_ERR_NUMERIC;
}

// TrustRegionMinimizer::Minimize [recalled, ceres 2.2.0 trust_region_minimizer.cc] with
// LevenbergMarquardtStrategy and the monotonic TrustRegionStepEvaluator; default options.
int oracle_solve(const mcba_problem* prob, mcba_params* params, const mcba_options* opt, mcba_summary* summary,
                 mcba_trace_row* trace, int trace_cap, int n_threads) {
  Work W;
  if (!init_work(W, prob, params, opt, n_threads)) return MCBA_ERR_INVALID;
  const Layout& L = W.L; const int n = L.n_x;
  const double t0 = now_s();
  mcba_summary S; std::memset(&S, 0, sizeof(S));
  int n_rows = 0;
  auto record = [&](const mcba_trace_row& r) {
    if (trace && n_rows < trace_cap) trace[n_rows] = r;
    ++n_rows;
    if (opt->verbose)
      std::printf("%4d % .6e % .2e % .2e % .2e % .2e % .2e\n", r.iteration, r.cost, r.cost_change,
                  r.gradient_max_norm, r.step_norm, r.relative_decrease, r.trust_region_radius);
  };

  // Init
  const std::vector<double> x_start(W.x);
  double x_norm = norm2(W.x);
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0;
  std::vector<double> scale(n, 1.0), diag(n, 0.0), D(n), y, step(n), delta(n), cand(n), grad, sq;
  double x_cost = 0, grad_max = 0;

  auto eval_grad_jac = [&](bool first) -> bool {
    if (!evaluate(W, W.x, true, &x_cost)) return false;
    gradient_and_colnorm(W, &grad, first && opt->jacobi_scaling ? &sq : nullptr);   // gradient of the unscaled J
    if (opt->jacobi_scaling) {
      if (first) for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(sq[i]));
      scale_columns(W, scale);
    }
    // |x - Plus(x, -g)|_inf, computed the way Ceres does (through Plus)
    double m = 0;
    for (int i = 0; i < n; ++i) { const double pg = W.x[i] + (-grad[i]); m = std::max(m, std::fabs(W.x[i] - pg)); }
    grad_max = m;
    return true;
  };

  // IterationZero
  mcba_trace_row it; std::memset(&it, 0, sizeof(it));
  if (!eval_grad_jac(true)) {
    S.termination = MCBA_FAILURE; if (summary) *summary = S; return MCBA_ERR_NUMERIC;
  }
  S.initial_cost = x_cost;
  it.iteration = 0; it.step_is_valid = 1; it.step_is_successful = 1; it.cost = x_cost; it.gradient_max_norm = grad_max;
  double minimum_cost = std::numeric_limits<double>::max();
  std::vector<double> best_x(W.x);
  double ref_cost = x_cost, cur_cost = x_cost;   // TrustRegionStepEvaluator, monotonic: reference == current
  bool at_least_one_success = false;
  int termination = MCBA_NO_CONVERGENCE;

  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it.step_is_successful) {
      ++S.num_successful_steps;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; best_x = W.x; }
    } else {
      ++S.num_unsuccessful_steps;
    }
    it.trust_region_radius = radius;
    record(it);
    if (it.iteration >= opt->max_num_iterations) { termination = MCBA_NO_CONVERGENCE; break; }
    if (it.step_is_successful && it.gradient_max_norm <= opt->gradient_tolerance) { termination = MCBA_CONVERGENCE; break; }
    if (it.trust_region_radius <= opt->min_trust_region_radius) { termination = MCBA_CONVERGENCE; break; }

    const double prev_grad_max = it.gradient_max_norm;
    const int iter = it.iteration + 1;
    std::memset(&it, 0, sizeof(it));
    it.iteration = iter;

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      gradient_and_colnorm(W, nullptr, &diag);   // of the scaled Jacobian
      for (int i = 0; i < n; ++i) diag[i] = std::min(std::max(diag[i], opt->min_lm_diagonal), opt->max_lm_diagonal);
    }
    for (int i = 0; i < n; ++i) D[i] = std::sqrt(diag[i] / radius);
    bool solved = schur_solve(W, D, y);
    reuse_diagonal = true;
    double mcc = 0.0;
    if (solved) {
      for (int i = 0; i < n; ++i) step[i] = -y[i];
      mcc = model_cost_change(W, step);
      it.step_is_valid = mcc > 0.0;
    }
    if (!it.step_is_valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) { termination = MCBA_FAILURE; break; }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;   // StepIsInvalid -> StepRejected(0)
      it.cost = x_cost; it.cost_change = 0; it.gradient_max_norm = prev_grad_max; it.step_norm = 0; it.relative_decrease = 0;
      continue;
    }
    for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i];
    num_consecutive_invalid = 0;

    // ComputeCandidatePointAndEvaluateCost
    for (int i = 0; i < n; ++i) cand[i] = W.x[i] + delta[i];
    double cand_cost;
    if (!evaluate(W, cand, false, &cand_cost)) cand_cost = std::numeric_limits<double>::max();

    if (at_least_one_success) {
      // ParameterToleranceReached
      double sn = 0; for (int i = 0; i < n; ++i) { const double d = W.x[i] - cand[i]; sn += d * d; }
      it.step_norm = std::sqrt(sn);
      if (it.step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) { termination = MCBA_CONVERGENCE; break; }
      // FunctionToleranceReached
      it.cost_change = x_cost - cand_cost;
      if (std::fabs(it.cost_change) <= opt->function_tolerance * x_cost) { termination = MCBA_CONVERGENCE; break; }
    }

    // IsStepSuccessful (TrustRegionStepEvaluator::StepQuality, monotonic)
    double rel;
    if (cand_cost >= std::numeric_limits<double>::max()) rel = std::numeric_limits<double>::lowest();
    else rel = std::max((cur_cost - cand_cost) / mcc, (ref_cost - cand_cost) / mcc);
    it.relative_decrease = rel;
    if (rel > opt->min_relative_decrease) {
      at_least_one_success = true;
      // HandleSuccessfulStep
      W.x = cand;
      x_norm = norm2(W.x);
      if (!eval_grad_jac(false)) { termination = MCBA_FAILURE; break; }
      it.step_is_successful = 1; it.cost = x_cost; it.gradient_max_norm = grad_max;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(opt->max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      cur_cost = cand_cost; ref_cost = cand_cost;
    } else {
      it.step_is_successful = 0; it.cost = cand_cost; it.gradient_max_norm = prev_grad_max;
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }

  // Minimizer returns the best point seen (parameters_ = x at minimum cost); ceres::Solve (solver.cc, Minimize) then
  // copies it to the user's parameter blocks only if the summary is usable - on FAILURE the ORIGINAL values are restored.
  write_back(W, termination == MCBA_FAILURE ? x_start : best_x, params);
  S.termination = termination;
  S.num_iterations = n_rows;
  S.final_cost = minimum_cost;
  {  // unrobustified RMS at the returned parameters
    mcba_options o2 = *opt; o2.huber_a = 0; Work W2; init_work(W2, prob, params, &o2, n_threads);
    double c2 = 0; evaluate(W2, W2.x, false, &c2);
    S.rms_px = std::sqrt(2.0 * c2 / (double)std::max<int64_t>(1, prob->n_obs));
  }
  S.t_total_ms = (now_s() - t0) * 1e3;
  if (summary) *summary = S;
  return MCBA_OK;
}

int oracle_time_iteration(const mcba_problem* prob, const mcba_params* params, const mcba_options* opt, int n_iters,
                          int n_threads, double* sec_per_iter, double* sec_eval, double* sec_linear) {
  Work W;
  if (!init_work(W, prob, params, opt, n_threads)) return MCBA_ERR_INVALID;
  const int n = W.L.n_x;
  std::vector<double> scale(n), diag, D(n), y, grad, sq, step(n), cand(n);
  double te = 0, tl = 0; const double t0 = now_s();
  for (int k = 0; k < n_iters; ++k) {
    double a = now_s(), cost;
    if (!evaluate(W, W.x, true, &cost)) return MCBA_ERR_NUMERIC;
    gradient_and_colnorm(W, &grad, &sq);
    for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(sq[i]));
    scale_columns(W, scale);
    double b = now_s(); te += b - a;
    gradient_and_colnorm(W, nullptr, &diag);
    for (int i = 0; i < n; ++i) D[i] = std::sqrt(std::min(std::max(diag[i], opt->min_lm_diagonal), opt->max_lm_diagonal) / opt->initial_trust_region_radius);
    if (!schur_solve(W, D, y)) return MCBA_ERR_NUMERIC;
    for (int i = 0; i < n; ++i) step[i] = -y[i];
    volatile double mcc = model_cost_change(W, step); (void)mcc;
    double c = now_s(); tl += c - b;
    for (int i = 0; i < n; ++i) cand[i] = W.x[i] + step[i] * scale[i];
    double cc; evaluate(W, cand, false, &cc);
    te += now_s() - c;
  }
  const double tt = now_s() - t0;
  if (sec_per_iter) *sec_per_iter = tt / n_iters;
  if (sec_eval) *sec_eval = te / n_iters;
  if (sec_linear) *sec_linear = tl / n_iters;
  return MCBA_OK;
}

}  // extern "C"
