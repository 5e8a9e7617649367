// mcba_pnp_math.cuh — arithmetic of the pose initialisation that feeds the bundle adjustment:
// point undistortion, a three-point pose solver with fourth-point disambiguation, the pinhole/Brown projection used
// for inlier counting, and the pieces of the pose-only Levenberg-Marquardt refinement. __host__ __device__ so that the
// arithmetic (not the kernels) can be unit-tested on a CPU-only box against OpenCV (tests/test_pnp_math.py through
// tests/host_math); the shipped library only calls it from device code.
//
// What it replaces in the reference (/root/reference/McCalib/src/geometrytools.cpp):
//   ransacP3P            :233-326   cv::solvePnP(4 points, SOLVEPNP_P3P) per hypothesis, cv::projectPoints + threshold for
//                                   the inliers, cv::solvePnP(inliers, useExtrinsicGuess, SOLVEPNP_ITERATIVE) to refine
//   ransacP3PDistortion  :709-751   Kannala cameras: cv::fisheye::undistortPoints, re-projection through K, then the
//                                   same RANSAC with a zero distortion vector
// The OpenCV routines themselves are a third-party dependency (OpenCV 4.x, absent from /root/reference); the P3P solver
// here is not a transcription of OpenCV's: it intersects the two conics of the depth constraints through a degenerate
// member of their pencil (a real root of a cubic), which yields the same up-to-four poses.
#pragma once
#include "mcba_math.cuh"

namespace mcba {

MCBA_HD bool is_fin(double x) { return fabs(x) <= DBL_MAX; }   // false for NaN and +-inf
MCBA_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
MCBA_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Pixel -> normalised image coordinates (a, b) = (X/Z, Y/Z) of the ray.
//   Brown: fixed-point iteration of cv::undistortPoints (x <- (x0 - tangential(x)) / radial(x)), run to convergence
//          instead of OpenCV's fixed 5 sweeps.
//   Kannala: cv::fisheye::undistortPoints: theta_d = |pw|, Newton on theta (1 + k1 t^2 + k2 t^4 + k3 t^6 + k4 t^8) = theta_d,
//          (a, b) = pw * tan(theta) / theta_d.
// Returns false when the iteration leaves the model's domain (the observation then never counts as an inlier).
MCBA_HD bool undistort_pixel(int model, const double* in, double u, double v, double& a, double& b) {
  const double x0 = (u - in[2]) / in[0], y0 = (v - in[3]) / in[1];
  if (model == 0) {
    const double k1 = in[4], k2 = in[5], p1 = in[6], p2 = in[7], k3 = in[8];
    double x = x0, y = y0;
    for (int it = 0; it < 50; ++it) {
      const double r2 = x * x + y * y;
      const double rad = 1.0 + ((k3 * r2 + k2) * r2 + k1) * r2;
      if (!(rad > 0.0)) return false;
      const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
      const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
      const double xn = (x0 - dx) / rad, yn = (y0 - dy) / rad;
      const double ch = fabs(xn - x) + fabs(yn - y);
      x = xn; y = yn;
      if (ch < 1e-15) break;
    }
    a = x; b = y;
    return is_fin(x) && is_fin(y);
  }
  const double k1 = in[4], k2 = in[5], k3 = in[6], k4 = in[7];
  double thd = sqrt(x0 * x0 + y0 * y0);
  const double half_pi = 1.5707963267948966;
  thd = fmin(fmax(-half_pi, thd), half_pi);
  if (!(thd > 1e-8)) { a = x0; b = y0; return true; }
  double th = thd;
  bool conv = false;
  for (int it = 0; it < 20; ++it) {
    const double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
    const double k0t2 = k1 * t2, k1t4 = k2 * t4, k2t6 = k3 * t6, k3t8 = k4 * t8;
    const double fix = (th * (1.0 + k0t2 + k1t4 + k2t6 + k3t8) - thd) / (1.0 + 3.0 * k0t2 + 5.0 * k1t4 + 7.0 * k2t6 + 9.0 * k3t8);
    th -= fix;
    if (fabs(fix) < 1e-14) { conv = true; break; }
  }
  if (!conv || !(th > 0.0) || !(th < half_pi)) return false;   // theta flipped or did not converge (OpenCV marks these -1e6)
  const double sc = tan(th) / thd;
  a = x0 * sc; b = y0 * sc;
  return is_fin(a) && is_fin(b);
}

// Pixel of a camera-frame point under the Brown model (all-zero distortion = pinhole). Used for inlier counting.
MCBA_HD bool project_brown(const double* in, const double* X, double& u, double& v) {
  if (!(X[2] > 0.0)) { u = v = 1e30; return false; }
  const double iz = 1.0 / X[2], a = X[0] * iz, b = X[1] * iz;
  const double r2 = a * a + b * b;
  const double rc = 1.0 + ((in[8] * r2 + in[5]) * r2 + in[4]) * r2;
  const double xd = a * rc + 2.0 * in[6] * a * b + in[7] * (r2 + 2.0 * a * a);
  const double yd = b * rc + in[6] * (r2 + 2.0 * b * b) + 2.0 * in[7] * a * b;
  u = in[0] * xd + in[2]; v = in[1] * yd + in[3];
  return true;
}

// Real roots of c3 x^3 + c2 x^2 + c1 x + c0 (c3 may vanish); returns their number (0..3).
MCBA_HD int real_cubic_roots(double c3, double c2, double c1, double c0, double* r) {
  const double sc = fmax(fmax(fabs(c3), fabs(c2)), fmax(fabs(c1), fabs(c0)));
  if (!(sc > 0.0)) return 0;
  int n = 0;
  if (fabs(c3) < 1e-14 * sc) {          // quadratic (or lower)
    if (fabs(c2) < 1e-14 * sc) { if (fabs(c1) > 0.0) r[n++] = -c0 / c1; return n; }
    const double d = c1 * c1 - 4.0 * c2 * c0;
    if (d < 0.0) return 0;
    const double s = sqrt(d), qq = -0.5 * (c1 + (c1 >= 0.0 ? s : -s));
    r[n++] = qq / c2; if (qq != 0.0) r[n++] = c0 / qq;
    return n;
  }
  const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
  const double Q = (a * a - 3.0 * b) / 9.0, R = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  const double Q3 = Q * Q * Q;
  if (R * R < Q3) {
    const double th = acos(fmin(1.0, fmax(-1.0, R / sqrt(Q3)))), sq = -2.0 * sqrt(Q);
    r[0] = sq * cos(th / 3.0) - a / 3.0;
    r[1] = sq * cos((th + 6.283185307179586) / 3.0) - a / 3.0;
    r[2] = sq * cos((th - 6.283185307179586) / 3.0) - a / 3.0;
    n = 3;
  } else {
    const double A = -(R >= 0.0 ? 1.0 : -1.0) * cbrt(fabs(R) + sqrt(fmax(0.0, R * R - Q3)));
    const double B = (A != 0.0) ? Q / A : 0.0;
    r[0] = (A + B) - a / 3.0;
    n = 1;
  }
  for (int i = 0; i < n; ++i) {          // Newton polish on the original polynomial
    double x = r[i];
    for (int it = 0; it < 3; ++it) {
      const double f = ((c3 * x + c2) * x + c1) * x + c0, df = (3.0 * c3 * x + 2.0 * c2) * x + c1;
      if (df == 0.0) break;
      const double xn = x - f / df;
      if (!is_fin(xn)) break;
      x = xn;
    }
    r[i] = x;
  }
  return n;
}

// symmetric 3x3 stored as {m00, m01, m02, m11, m12, m22}
MCBA_HD double sym_det(const double* m) {
  return m[0] * (m[3] * m[5] - m[4] * m[4]) - m[1] * (m[1] * m[5] - m[4] * m[2]) + m[2] * (m[1] * m[4] - m[3] * m[2]);
}
MCBA_HD void sym_adj(const double* m, double* a) {   // adjugate (symmetric)
  a[0] = m[3] * m[5] - m[4] * m[4]; a[1] = m[2] * m[4] - m[1] * m[5]; a[2] = m[1] * m[4] - m[2] * m[3];
  a[3] = m[0] * m[5] - m[2] * m[2]; a[4] = m[1] * m[2] - m[0] * m[4]; a[5] = m[0] * m[3] - m[1] * m[1];
}
MCBA_HD double sym_trace_prod(const double* a, const double* b) {   // tr(A B)
  return a[0] * b[0] + a[3] * b[3] + a[5] * b[5] + 2.0 * (a[1] * b[1] + a[2] * b[2] + a[4] * b[4]);
}
MCBA_HD double sym_quad(const double* m, const double* x, const double* y) {   // x^T M y
  return x[0] * (m[0] * y[0] + m[1] * y[1] + m[2] * y[2]) + x[1] * (m[1] * y[0] + m[3] * y[1] + m[4] * y[2]) +
         x[2] * (m[2] * y[0] + m[4] * y[1] + m[5] * y[2]);
}
// unit eigenvector of the symmetric matrix m for eigenvalue e (largest cross product of two rows of m - e I)
MCBA_HD bool sym_eigvec(const double* m, double e, double* v) {
  const double r0[3] = {m[0] - e, m[1], m[2]}, r1[3] = {m[1], m[3] - e, m[4]}, r2[3] = {m[2], m[4], m[5] - e};
  double c[3][3];
  cross3(r0, r1, c[0]); cross3(r0, r2, c[1]); cross3(r1, r2, c[2]);
  int best = 0; double bn = -1.0;
  for (int i = 0; i < 3; ++i) { const double n2 = dot3(c[i], c[i]); if (n2 > bn) { bn = n2; best = i; } }
  if (!(bn > 0.0)) return false;
  const double s = 1.0 / sqrt(bn);
  v[0] = c[best][0] * s; v[1] = c[best][1] * s; v[2] = c[best][2] * s;
  return true;
}

struct P3PSolutions { int n; double R[4][9]; double t[4][3]; };

// Three-point pose. f[i]: unit bearing vectors in the camera frame; P[i]: the points in the world (board) frame.
// Depths l_i > 0 with |l_i f_i - l_j f_j|^2 = |P_i - P_j|^2: two homogeneous conics in (l_1 : l_2 : l_3); a real root of
// det(D1 + g D2) = 0 gives a line pair through their (up to four) common points; each line meets D2 in two points.
// Each candidate is polished with Gauss-Newton on the three distance equations, then (R, t) aligns the two triangles.
MCBA_HD void p3p_solve(const double f[3][3], const double P[3][3], P3PSolutions& out) {
  out.n = 0;
  double d12[3], d13[3], d23[3];
  for (int k = 0; k < 3; ++k) { d12[k] = P[0][k] - P[1][k]; d13[k] = P[0][k] - P[2][k]; d23[k] = P[1][k] - P[2][k]; }
  const double a12 = dot3(d12, d12), a13 = dot3(d13, d13), a23 = dot3(d23, d23);
  double nx[3]; cross3(d12, d13, nx);
  if (!(dot3(nx, nx) > 1e-12 * a12 * a13) || !(a23 > 0.0)) return;     // (nearly) collinear or coincident points
  const double b12 = dot3(f[0], f[1]), b13 = dot3(f[0], f[2]), b23 = dot3(f[1], f[2]);
  // M12 = [1 -b12 0; -b12 1 0; 0 0 0], M13 = [1 0 -b13; 0 0 0; -b13 0 1], M23 = [0 0 0; 0 1 -b23; 0 -b23 1]
  const double D1[6] = {a23, -b12 * a23, 0.0, a23 - a12, b23 * a12, -a12};          // M12 a23 - M23 a12
  const double D2[6] = {a23, 0.0, -b13 * a23, -a13, b23 * a13, a23 - a13};          // M13 a23 - M23 a13
  double A1[6], A2[6];
  sym_adj(D1, A1); sym_adj(D2, A2);
  double roots[3];
  const int nr = real_cubic_roots(sym_det(D2), sym_trace_prod(A2, D1), sym_trace_prod(A1, D2), sym_det(D1), roots);
  // choose the root whose member of the pencil is the most clearly a real line pair (eigenvalues of opposite sign)
  double D0[6], e1 = 0.0, e2 = 0.0; bool have = false; double best_q = 0.0;
  for (int i = 0; i < nr; ++i) {
    double T[6];
    for (int k = 0; k < 6; ++k) T[k] = D1[k] + roots[i] * D2[k];
    const double tr = T[0] + T[3] + T[5];
    const double pm = (T[0] * T[3] - T[1] * T[1]) + (T[0] * T[5] - T[2] * T[2]) + (T[3] * T[5] - T[4] * T[4]);   // e1 e2
    if (!(pm < 0.0)) continue;
    const double q = -pm / (tr * tr - 2.0 * pm);          // |e1 e2| / (e1^2 + e2^2): scale-free, in (0, 1/2]
    if (!have || q > best_q) {
      have = true; best_q = q;
      for (int k = 0; k < 6; ++k) D0[k] = T[k];
      const double disc = sqrt(fmax(0.0, tr * tr - 4.0 * pm));
      e1 = 0.5 * (tr + disc); e2 = 0.5 * (tr - disc);
      // the smaller-magnitude root of e^2 - tr e + pm from the product, for accuracy
      if (fabs(e1) >= fabs(e2)) e2 = pm / e1; else e1 = pm / e2;
    }
  }
  if (!have) return;
  double v1[3], v2[3];
  if (!sym_eigvec(D0, e1, v1) || !sym_eigvec(D0, e2, v2)) return;
  const double s1 = sqrt(fabs(e1)), s2 = sqrt(fabs(e2));
  for (int sgn = 0; sgn < 2; ++sgn) {
    double l[3];
    for (int k = 0; k < 3; ++k) l[k] = s1 * v1[k] + (sgn ? -s2 : s2) * v2[k];
    // orthonormal basis (u1, u2) of the plane l . x = 0
    int kmin = 0;
    if (fabs(l[1]) < fabs(l[kmin])) kmin = 1;
    if (fabs(l[2]) < fabs(l[kmin])) kmin = 2;
    double ek[3] = {0.0, 0.0, 0.0}; ek[kmin] = 1.0;
    double u1[3], u2[3];
    cross3(l, ek, u1);
    double n1 = sqrt(dot3(u1, u1));
    if (!(n1 > 0.0)) continue;
    for (int k = 0; k < 3; ++k) u1[k] /= n1;
    cross3(l, u1, u2);
    n1 = sqrt(dot3(u2, u2));
    if (!(n1 > 0.0)) continue;
    for (int k = 0; k < 3; ++k) u2[k] /= n1;
    const double A = sym_quad(D2, u1, u1), B = sym_quad(D2, u1, u2), C = sym_quad(D2, u2, u2);
    double disc = B * B - A * C;
    if (disc < 0.0) {
      if (disc > -1e-12 * (B * B + fabs(A * C))) disc = 0.0; else continue;
    }
    const double sq = sqrt(disc);
    for (int pm = 0; pm < 2; ++pm) {
      if (pm == 1 && sq == 0.0) break;
      double al, be;
      if (fabs(A) >= fabs(C)) { if (A == 0.0) continue; be = 1.0; al = (-B + (pm ? -sq : sq)) / A; }
      else { al = 1.0; be = (-B + (pm ? -sq : sq)) / C; }
      double lam[3];
      for (int k = 0; k < 3; ++k) lam[k] = al * u1[k] + be * u2[k];
      // all depths of one sign
      if (lam[0] < 0.0) { lam[0] = -lam[0]; lam[1] = -lam[1]; lam[2] = -lam[2]; }
      if (!(lam[0] > 0.0 && lam[1] > 0.0 && lam[2] > 0.0)) continue;
      // scale from the sum of the three constraints
      const double q12 = lam[0] * lam[0] + lam[1] * lam[1] - 2.0 * b12 * lam[0] * lam[1];
      const double q13 = lam[0] * lam[0] + lam[2] * lam[2] - 2.0 * b13 * lam[0] * lam[2];
      const double q23 = lam[1] * lam[1] + lam[2] * lam[2] - 2.0 * b23 * lam[1] * lam[2];
      const double qs = q12 + q13 + q23;
      if (!(qs > 0.0)) continue;
      const double s = sqrt((a12 + a13 + a23) / qs);
      lam[0] *= s; lam[1] *= s; lam[2] *= s;
      // Gauss-Newton on the three distance equations
      for (int it = 0; it < 5; ++it) {
        const double r0 = lam[0] * lam[0] + lam[1] * lam[1] - 2.0 * b12 * lam[0] * lam[1] - a12;
        const double r1 = lam[0] * lam[0] + lam[2] * lam[2] - 2.0 * b13 * lam[0] * lam[2] - a13;
        const double r2 = lam[1] * lam[1] + lam[2] * lam[2] - 2.0 * b23 * lam[1] * lam[2] - a23;
        if (fabs(r0) + fabs(r1) + fabs(r2) < 1e-15 * (a12 + a13 + a23)) break;
        const double j00 = 2.0 * (lam[0] - b12 * lam[1]), j01 = 2.0 * (lam[1] - b12 * lam[0]);
        const double j10 = 2.0 * (lam[0] - b13 * lam[2]), j12 = 2.0 * (lam[2] - b13 * lam[0]);
        const double j21 = 2.0 * (lam[1] - b23 * lam[2]), j22 = 2.0 * (lam[2] - b23 * lam[1]);
        // J = [j00 j01 0; j10 0 j12; 0 j21 j22]
        const double det = -j00 * j12 * j21 - j01 * j10 * j22;
        if (!(fabs(det) > 1e-300)) break;
        const double id = 1.0 / det;
        const double dx0 = (-j12 * j21 * r0 - j01 * j22 * r1 + j01 * j12 * r2) * id;
        const double dx1 = (-j10 * j22 * r0 + j00 * j22 * r1 - j00 * j12 * r2) * id;
        const double dx2 = (j10 * j21 * r0 - j00 * j21 * r1 - j01 * j10 * r2) * id;
        lam[0] -= dx0; lam[1] -= dx1; lam[2] -= dx2;
      }
      if (!(lam[0] > 0.0 && lam[1] > 0.0 && lam[2] > 0.0)) continue;
      // align the triangles: R [x1 x2 x3] = [y1 y2 y3]
      double Y[3][3];
      for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) Y[i][k] = lam[i] * f[i][k];
      double x1[3], x2[3], x3[3], y1[3], y2[3], y3[3];
      for (int k = 0; k < 3; ++k) { x1[k] = P[1][k] - P[0][k]; x2[k] = P[2][k] - P[0][k]; y1[k] = Y[1][k] - Y[0][k]; y2[k] = Y[2][k] - Y[0][k]; }
      cross3(x1, x2, x3); cross3(y1, y2, y3);
      // inverse of X = [x1 x2 x3] (columns): rows are (x2 x x3, x3 x x1, x1 x x2) / det
      double c23[3], c31[3], c12[3];
      cross3(x2, x3, c23); cross3(x3, x1, c31); cross3(x1, x2, c12);
      const double detX = dot3(x1, c23);
      if (!(fabs(detX) > 0.0)) continue;
      const double idx = 1.0 / detX;
      double* R = out.R[out.n]; double* t = out.t[out.n];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = (y1[i] * c23[j] + y2[i] * c31[j] + y3[i] * c12[j]) * idx;
      for (int i = 0; i < 3; ++i) t[i] = Y[0][i] - (R[i * 3] * P[0][0] + R[i * 3 + 1] * P[0][1] + R[i * 3 + 2] * P[0][2]);
      bool fin = true;
      for (int i = 0; i < 9; ++i) fin = fin && is_fin(R[i]);
      if (!fin) continue;
      // drop duplicates (a double intersection found from both lines)
      bool dup = false;
      for (int s0 = 0; s0 < out.n && !dup; ++s0) {
        double dd = 0.0;
        for (int i = 0; i < 9; ++i) dd += fabs(out.R[s0][i] - R[i]);
        for (int i = 0; i < 3; ++i) dd += fabs(out.t[s0][i] - t[i]) / (1e-300 + fabs(t[i]) + fabs(out.t[s0][i]));
        dup = dd < 1e-9;
      }
      if (dup) continue;
      if (++out.n == 4) return;
    }
  }
}

// The hypothesis of one RANSAC draw: pose from the first three of four points; the fourth picks among the (up to four)
// solutions by the squared pixel distance of its re-projection through the working camera model (what cv::solvePnP
// does with SOLVEPNP_P3P and 4 points: cv::projectPoints of the fourth point, smallest error first).
// ab[i]: normalised coordinates of the observations, X[i]: their board points, in: working intrinsics, w4: working
// pixel of the fourth observation.
MCBA_HD bool p3p_hypothesis(const double ab[4][2], const double X[4][3], const double* in, const double* w4, double* R, double* t) {
  double f[3][3], P[3][3];
  for (int i = 0; i < 3; ++i) {
    const double n = 1.0 / sqrt(ab[i][0] * ab[i][0] + ab[i][1] * ab[i][1] + 1.0);
    f[i][0] = ab[i][0] * n; f[i][1] = ab[i][1] * n; f[i][2] = n;
    for (int k = 0; k < 3; ++k) P[i][k] = X[i][k];
  }
  P3PSolutions sol;
  p3p_solve(f, P, sol);
  double best = 1e300; int bi = -1;
  for (int s = 0; s < sol.n; ++s) {
    double Xc[3], pu, pv;
    mat3_vec_add(sol.R[s], X[3], sol.t[s], Xc);
    if (!project_brown(in, Xc, pu, pv)) continue;
    const double du = pu - w4[0], dv = pv - w4[1];
    const double e = du * du + dv * dv;
    if (e < best) { best = e; bi = s; }
  }
  if (bi < 0) return false;
  for (int i = 0; i < 9; ++i) R[i] = sol.R[bi][i];
  for (int i = 0; i < 3; ++i) t[i] = sol.t[bi][i];
  return true;
}

// Rotation matrix -> angle-axis (what cv::Rodrigues returns for a rotation matrix; near pi the axis comes from the
// symmetric part).
MCBA_HD void rot_to_rvec(const double* R, double* rv) {
  const double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];   // 2 sin(theta) axis
  const double s = 0.5 * sqrt(rx * rx + ry * ry + rz * rz);
  double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  c = fmin(1.0, fmax(-1.0, c));
  const double th = atan2(s, c);
  if (s < 1e-5) {
    if (c > 0.0) { rv[0] = 0.5 * rx; rv[1] = 0.5 * ry; rv[2] = 0.5 * rz; return; }   // first order around the identity
    // theta ~ pi: (R + R^T)/2 = c I + (1 - c) a a^T; the column of a a^T with the largest diagonal gives the axis
    const double t = 1.0 / (1.0 - c);
    const double m[3][3] = {{(R[0] - c) * t, 0.5 * (R[1] + R[3]) * t, 0.5 * (R[2] + R[6]) * t},
                            {0.5 * (R[1] + R[3]) * t, (R[4] - c) * t, 0.5 * (R[5] + R[7]) * t},
                            {0.5 * (R[2] + R[6]) * t, 0.5 * (R[5] + R[7]) * t, (R[8] - c) * t}};
    int k = 0;
    if (m[1][1] > m[k][k]) k = 1;
    if (m[2][2] > m[k][k]) k = 2;
    double a[3] = {m[0][k], m[1][k], m[2][k]};
    double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (!(n > 0.0)) { rv[0] = th; rv[1] = 0.0; rv[2] = 0.0; return; }
    if (a[0] * rx + a[1] * ry + a[2] * rz < 0.0) n = -n;
    rv[0] = a[0] * th / n; rv[1] = a[1] * th / n; rv[2] = a[2] * th / n;
    return;
  }
  const double k = 0.5 * th / s;
  rv[0] = rx * k; rv[1] = ry * k; rv[2] = rz * k;
}

// Unrobustified residual and its 2x6 Jacobian w.r.t. the pose (angle-axis, translation) for the pose-only refinement
// (cv::solvePnP SOLVEPNP_ITERATIVE minimises the same sum of squared pixel errors). prep = prep_block(pose).
MCBA_HD bool pnp_residual_jac(const double* prep, const double* in, const double* X0, double u, double v,
                              double* r, double* Ju, double* Jv) {
  double X[3];
  mat3_vec_add(prep, X0, prep + 36, X);
  if (!(X[2] > 0.0)) return false;
  const double iz = 1.0 / X[2], a = X[0] * iz, b = X[1] * iz;
  Distort d;
  distort<true>(0, in, a, b, d);
  r[0] = in[0] * d.xd + in[2] - u;
  r[1] = in[1] * d.yd + in[3] - v;
  const double fx = in[0], fy = in[1];
  const double Pu[3] = {fx * d.dxa * iz, fx * d.dxb * iz, -fx * (d.dxa * a + d.dxb * b) * iz};
  const double Pv[3] = {fy * d.dya * iz, fy * d.dyb * iz, -fy * (d.dya * a + d.dyb * b) * iz};
  for (int k = 0; k < 3; ++k) {
    double g[3];
    mat3_vec(prep + 9 + 9 * k, X0, g);
    Ju[k] = Pu[0] * g[0] + Pu[1] * g[1] + Pu[2] * g[2];
    Jv[k] = Pv[0] * g[0] + Pv[1] * g[1] + Pv[2] * g[2];
    Ju[3 + k] = Pu[k]; Jv[3 + k] = Pv[k];
  }
  return true;
}

// 6x6 SPD solve for the damped normal equations (H + lambda diag(H)) d = -g, H given as its upper triangle (21).
// Returns false when the damped matrix is not positive definite.
MCBA_HD bool pnp_solve6(const double* H, double lambda, const double* g, double* d) {
  double L[21], inv[6];          // lower factor, row-major packed: L[i(i+1)/2 + j], j <= i
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int i = j; i < 6; ++i) {
      // H upper index of (j, i), j <= i: j*6 - j(j-1)/2 + (i - j)
      double s = H[j * 6 - j * (j - 1) / 2 + (i - j)];
      if (i == j) s *= (1.0 + lambda);
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      if (i == j) {
        if (!(s > 0.0) || !(s <= DBL_MAX)) return false;
        const double l = sqrt(s);
        inv[j] = rcp(l);
        L[j * (j + 1) / 2 + j] = l;
      } else {
        L[i * (i + 1) / 2 + j] = s * inv[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {          // forward
    double s = -g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * (i + 1) / 2 + k] * d[k];
    d[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {         // backward
    double s = d[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k * (k + 1) / 2 + i] * d[k];
    d[i] = s * inv[i];
  }
  return true;
}

// Levenberg-Marquardt driver of the pose-only refinement (damping H + lambda diag(H), lambda /10 on success and x10
// on failure as cv::solvePnP's iterative solver; stops on a relative step below 1e-8 or when the cost can no longer be
// improved beyond rounding, instead of OpenCV's FLT_EPSILON / 20 iterations). One evaluation (H, g, cost at a point)
// per iteration: propose() returns the point to evaluate, update() consumes the evaluation.
struct PnpLm {
  double pose[6], H[21], g[6], cost, lambda;
  bool have, fin, stepok;
  MCBA_HD void init(const double* p0) {
    for (int k = 0; k < 6; ++k) pose[k] = p0[k];
    cost = 0.0; lambda = 1e-3; have = false; fin = false; stepok = true;
  }
  MCBA_HD void propose(double* cand) {
    stepok = true;
    if (!have || fin) { for (int k = 0; k < 6; ++k) cand[k] = pose[k]; return; }
    double d[6];
    if (!pnp_solve6(H, lambda, g, d)) { stepok = false; for (int k = 0; k < 6; ++k) cand[k] = pose[k]; return; }
    double dn = 0.0, xn = 0.0;
    for (int k = 0; k < 6; ++k) { cand[k] = pose[k] + d[k]; dn += d[k] * d[k]; xn += pose[k] * pose[k]; }
    if (dn <= 1e-16 * (xn + 1e-16)) {      // |d| <= 1e-8 |x|: take the step, converged
      for (int k = 0; k < 6; ++k) pose[k] = cand[k];
      fin = true;
    }
  }
  MCBA_HD void update(const double* cand, const double* H2, const double* g2, double cost2) {
    if (fin) return;
    if (!have) {
      for (int k = 0; k < 21; ++k) H[k] = H2[k];
      for (int k = 0; k < 6; ++k) g[k] = g2[k];
      cost = cost2; have = true;
      if (!(cost2 <= DBL_MAX)) fin = true;   // the starting pose cannot be evaluated
      return;
    }
    if (stepok && cost2 < cost) {
      for (int k = 0; k < 21; ++k) H[k] = H2[k];
      for (int k = 0; k < 6; ++k) { g[k] = g2[k]; pose[k] = cand[k]; }
      if (cost - cost2 <= 1e-12 * cost) fin = true;
      cost = cost2;
      lambda = fmax(lambda * 0.1, 1e-16);
    } else {
      if (stepok && cost2 - cost <= 1e-12 * cost) fin = true;   // at the minimum up to rounding
      lambda *= 10.0;
      if (lambda > 1e16) fin = true;
    }
  }
};

// splitmix64: the counter-based generator behind the RANSAC draws (the reference seeds std::mt19937 from
// std::random_device at every draw, i.e. it is not reproducible; here draw h of board observation b under seed s is a
// pure function of (s, b, h)).
MCBA_HD unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
// four distinct indices in [0, n), n >= 4, in draw order
MCBA_HD void ransac_sample4(unsigned long long seed, unsigned long long b, unsigned long long h, int n, int* idx) {
  unsigned long long st = splitmix64(seed ^ splitmix64(b * 0x100000001B3ULL + h));
  int sorted[4];
  for (int k = 0; k < 4; ++k) {
    st = splitmix64(st);
    int r = (int)(st % (unsigned long long)(n - k));
    int pos = 0;
    while (pos < k && r >= sorted[pos]) { ++r; ++pos; }
    for (int q = k; q > pos; --q) sorted[q] = sorted[q - 1];
    sorted[pos] = r;
    idx[k] = r;
  }
}

// The sequential stopping rule of ransacP3P (:243-300) applied to the inlier count of the next hypothesis.
struct RansacState { int N, trialcount, countit, best_count, best_h; };
MCBA_HD void ransac_init(RansacState& s, int it) { s.N = it; s.trialcount = 0; s.countit = 0; s.best_count = 0; s.best_h = -1; }
MCBA_HD bool ransac_continue(const RansacState& s, int it) { return s.N > s.trialcount && s.countit < it; }
MCBA_HD void ransac_update(RansacState& s, int h, int count, int total, double p) {
  s.trialcount++;
  if (count > s.best_count) {
    s.best_count = count; s.best_h = h;
    const double eps = 0.00001;
    const double frac = (double)count / (double)total;
    double pno = 1.0 - frac * frac * frac;
    if (pno == 0.0) pno = eps;
    if (pno > 1.0 - eps) pno = 1.0 - eps;
    const double est = log(1.0 - p) / log(pno);
    s.N = (int)floor(fmin(est, 2.0e9) + 0.5);   // int(round(est)), est > 0
    s.trialcount = 0;
  }
  s.countit++;
}

}  // namespace mcba
