// mcba_math.cuh — per-block and per-observation arithmetic of the reprojection model, analytic
// (no dual numbers), shared by every observation kernel. Functions are __host__ __device__ so the
// arithmetic (not the kernels) can also be unit-tested against the oracle on a CPU-only box
// (tests/test_math_host.py builds csrc/math_host_test.cpp); the shipped library only calls them
// from device code.
//
// Model (/root/reference/McCalib/include/OptimizationCeres.h):
//   X1 = board refined ? R(wb) X0 + tb : X0          :401-409
//   X2 = R(wo) X1 + to                                :413-417   (e-block, always present)
//   X3 = camera refined ? R(wc) X2 + tc : X2          :420-427
//   a = X3x/X3z, b = X3y/X3z                           :430-431
//   Brown :436-443 / Kannala :460-477, pixels :480-486, HuberLoss(1.0) at the call sites.
#pragma once
#include <cmath>
#include <cfloat>

#ifdef __CUDACC__
#define MCBA_HD __host__ __device__ __forceinline__
#else
#define MCBA_HD inline
#endif

namespace mcba {

// IEEE-rounded reciprocal without the slow path of the generic fp64 division (device); plain division on the host
MCBA_HD double rcp(double x) {
#ifdef __CUDA_ARCH__
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// Local column order of a board observation's Jacobian: [cam pose 6][intrinsics 9][board 6] | [e-block 6] | [residual]
template <int STAGE> struct StageTraits;
template <> struct StageTraits<1> { static constexpr bool CAM = false, INTR = true,  BOARD = false; };
template <> struct StageTraits<2> { static constexpr bool CAM = false, INTR = false, BOARD = true;  };
template <> struct StageTraits<3> { static constexpr bool CAM = true,  INTR = false, BOARD = false; };
template <> struct StageTraits<4> { static constexpr bool CAM = true,  INTR = false, BOARD = true;  };
template <> struct StageTraits<5> { static constexpr bool CAM = true,  INTR = true,  BOARD = true;  };
template <int STAGE> struct Cols {
  typedef StageTraits<STAGE> T;
  static constexpr int C0 = 0;                              // camera pose
  static constexpr int I0 = C0 + (T::CAM ? 6 : 0);          // intrinsics
  static constexpr int B0 = I0 + (T::INTR ? 9 : 0);         // board pose
  static constexpr int WF = B0 + (T::BOARD ? 6 : 0);        // width of the f part
  static constexpr int E0 = WF;                             // e-block pose
  static constexpr int R0 = WF + 6;                         // residual column
  static constexpr int NC = WF + 7;                         // columns incl. residual
  static constexpr int NBLK = (NC + 7) / 8;                 // 8-wide column blocks
  static constexpr int NT = NBLK * (NBLK + 1) / 2;          // upper-triangle 8x8 tiles
  static constexpr int K = WF + 6;                          // Jacobian width (reference: 15,12,12,18,27)
  // Compact per-bobs record of J~^T J~ (only what is consumed, in the order the consumers read it):
  //   [ f x f upper (NFF) | f x r (WF) | f x e (6 WF) | e x e upper (21) | e x r (6) ]
  // The per-pair reduction reads the first NFF + WF doubles contiguously, the e-block assembly the rest.
  static constexpr int NFF = WF * (WF + 1) / 2;
  static constexpr int CI_FR = NFF, CI_FE = NFF + WF, CI_EE = CI_FE + 6 * WF, CI_ER = CI_EE + 21, NCOMP = CI_ER + 6;
  MCBA_HD static int cidx(int a, int b) {                   // entry (a <= b) -> index in the compact record, -1 if unused
    if (b < WF) return a * WF - a * (a - 1) / 2 + (b - a);
    if (b == R0) return a < WF ? CI_FR + a : (a < R0 ? CI_ER + (a - E0) : -1);
    if (a < WF) return CI_FE + a * 6 + (b - E0);
    const int p = a - E0, r = b - E0;
    return CI_EE + p * 6 - p * (p - 1) / 2 + (r - p);
  }
  // reference parameter-block order at the call sites (Camera.cpp:290, Object3D.cpp:208-209,
  // CameraGroup.cpp:253-256, :445-451, :553-560): [cam][pose][board][intr] with absent blocks removed
  static constexpr int REF_C = 0;
  static constexpr int REF_E = T::CAM ? 6 : 0;
  static constexpr int REF_B = REF_E + 6;
  static constexpr int REF_I = REF_B + (T::BOARD ? 6 : 0);
  MCBA_HD static int to_ref(int local) {   // local column -> column in the reference order
    if (T::CAM && local < I0) return REF_C + local;
    if (T::INTR && local >= I0 && local < B0) return REF_I + (local - I0);
    if (T::BOARD && local >= B0 && local < WF) return REF_B + (local - B0);
    return REF_E + (local - E0);
  }
};
MCBA_HD int tile_index(int bi, int bj, int nblk) { return bi * nblk - bi * (bi - 1) / 2 + (bj - bi); }

// Per-block rotation data: R (row-major 3x3), the three matrices dR/dw_k and t. 39 doubles.
//   prep[0..8] = R, prep[9+9k .. 9+9k+8] = dR/dw_k, prep[36..38] = t.
// Reproduces what autodiff of ceres::AngleAxisRotatePoint yields [recalled, ceres/rotation.h]:
// Rodrigues' formula differentiated in ambient R^3 when theta^2 > DBL_EPSILON; R = I + [w]x,
// dR/dw_k = [e_k]x otherwise.
constexpr int PREP = 39;
MCBA_HD void prep_block(const double* pose, double* out) {
  const double wx = pose[0], wy = pose[1], wz = pose[2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double* R = out; double* dR = out + 9;
  if (th2 > DBL_EPSILON) {
    const double th = sqrt(th2), ti = 1.0 / th;
    const double c = cos(th), s = sin(th), oc = 1.0 - c;
    const double w[3] = {wx * ti, wy * ti, wz * ti};
    // R = c I + s [w]x + (1-c) w w^T
    R[0] = c + oc * w[0] * w[0];        R[1] = -s * w[2] + oc * w[0] * w[1]; R[2] = s * w[1] + oc * w[0] * w[2];
    R[3] = s * w[2] + oc * w[1] * w[0]; R[4] = c + oc * w[1] * w[1];         R[5] = -s * w[0] + oc * w[1] * w[2];
    R[6] = -s * w[1] + oc * w[2] * w[0]; R[7] = s * w[0] + oc * w[2] * w[1]; R[8] = c + oc * w[2] * w[2];
    for (int k = 0; k < 3; ++k) {
      // d theta/d w_k = w_k ; d w/d w_k = (e_k - w_k w)/theta
      const double dth = w[k];
      double dw[3] = {-w[k] * w[0] * ti, -w[k] * w[1] * ti, -w[k] * w[2] * ti};
      dw[k] += ti;
      const double dc = -s * dth, ds = c * dth, doc = s * dth;
      double* D = dR + 9 * k;
      // d/dw_k [c I]
      // d/dw_k [s [w]x] = ds [w]x + s [dw]x
      // d/dw_k [(1-c) w w^T] = doc w w^T + (1-c)(dw w^T + w dw^T)
      D[0] = dc + doc * w[0] * w[0] + oc * 2.0 * dw[0] * w[0];
      D[4] = dc + doc * w[1] * w[1] + oc * 2.0 * dw[1] * w[1];
      D[8] = dc + doc * w[2] * w[2] + oc * 2.0 * dw[2] * w[2];
      const double s01 = doc * w[0] * w[1] + oc * (dw[0] * w[1] + w[0] * dw[1]);
      const double s02 = doc * w[0] * w[2] + oc * (dw[0] * w[2] + w[0] * dw[2]);
      const double s12 = doc * w[1] * w[2] + oc * (dw[1] * w[2] + w[1] * dw[2]);
      const double a0 = ds * w[0] + s * dw[0], a1 = ds * w[1] + s * dw[1], a2 = ds * w[2] + s * dw[2];
      D[1] = s01 - a2; D[3] = s01 + a2;
      D[2] = s02 + a1; D[6] = s02 - a1;
      D[5] = s12 - a0; D[7] = s12 + a0;
    }
  } else {
    R[0] = 1.0; R[1] = -wz; R[2] = wy;
    R[3] = wz;  R[4] = 1.0; R[5] = -wx;
    R[6] = -wy; R[7] = wx;  R[8] = 1.0;
    for (int i = 0; i < 27; ++i) dR[i] = 0.0;
    // [e_0]x, [e_1]x, [e_2]x
    dR[0 * 9 + 5] = -1.0; dR[0 * 9 + 7] = 1.0;
    dR[1 * 9 + 2] = 1.0;  dR[1 * 9 + 6] = -1.0;
    dR[2 * 9 + 1] = -1.0; dR[2 * 9 + 3] = 1.0;
  }
  out[36] = pose[3]; out[37] = pose[4]; out[38] = pose[5];
}

// Per board-observation context (shared by all its corners), 144 doubles:
constexpr int CTX_RB = 0, CTX_TB = 9, CTX_RO = 12, CTX_TO = 21, CTX_RC = 24, CTX_TC = 33;
constexpr int CTX_DRC = 36;   // 27: dRc_k
constexpr int CTX_MC = 63;    // 9 : Mc  = camera refined ? Rc : I          (d X3 / d X2)
constexpr int CTX_MCO = 72;   // 9 : Mco = Mc Ro                             (d X3 / d X1)
constexpr int CTX_NO = 81;    // 27: Mc  dRo_k
constexpr int CTX_NB = 108;   // 27: Mco dRb_k
constexpr int CTX_IN = 135;   // 9 : intrinsics
constexpr int CTX_SIZE = 144;
// The same context as a layout type (obs_jacobian is written against one):
struct CtxBuilt {   // entries computed / copied one by one (ctx_entry_a / ctx_entry_b)
  static constexpr int RB = CTX_RB, TB = CTX_TB, RO = CTX_RO, TO = CTX_TO, RC = CTX_RC, TC = CTX_TC, DRC = CTX_DRC,
                       MC = CTX_MC, MCO = CTX_MCO, NO = CTX_NO, NB = CTX_NB, IN = CTX_IN;
};
// Copy-free layout: the three prep records (camera, e-block pose, board) and the intrinsics land in shared memory as
// they are (asynchronous copies, issued one board observation ahead), only Mco / No / Nb are computed. A block that is
// not refined at this stage, or is the reference camera / board, points at the identity record (R = I, dR = 0, t = 0).
struct CtxRaw {
  static constexpr int PC = 0, PO = PREP, PB = 2 * PREP, IN = 3 * PREP;      // raw part: 3 x 39 + 9 = 126 doubles
  static constexpr int RAW = 3 * PREP + 9;
  static constexpr int MCO = RAW, NO = RAW + 9, NB = RAW + 36, SIZE = 192;  // 189 used
  static constexpr int RB = PB, TB = PB + 36, RO = PO, TO = PO + 36, RC = PC, TC = PC + 36, DRC = PC + 9, MC = PC;
};

// Phase A entries [0,81) + [135,144): copies, Mc, Mco. Each entry depends only on the prep records.
MCBA_HD double ctx_entry_a(int e, const double* pc, bool act_c, const double* po, const double* pb, bool act_b,
                           const double* intr) {
  if (e < CTX_TB) return act_b ? pb[e] : ((e % 4 == 0) ? 1.0 : 0.0);          // Rb (identity when not refined)
  if (e < CTX_RO) return act_b ? pb[36 + e - CTX_TB] : 0.0;                    // tb
  if (e < CTX_TO) return po[e - CTX_RO];                                       // Ro
  if (e < CTX_RC) return po[36 + e - CTX_TO];                                  // to
  if (e < CTX_TC) return act_c ? pc[e - CTX_RC] : (((e - CTX_RC) % 4 == 0) ? 1.0 : 0.0);   // Rc
  if (e < CTX_DRC) return act_c ? pc[36 + e - CTX_TC] : 0.0;                   // tc
  if (e < CTX_MC) return act_c ? pc[9 + e - CTX_DRC] : 0.0;                    // dRc
  if (e < CTX_MCO) { const int i = e - CTX_MC; return act_c ? pc[i] : ((i % 4 == 0) ? 1.0 : 0.0); }   // Mc
  if (e < CTX_NO) {                                                            // Mco = Mc Ro
    const int i = (e - CTX_MCO) / 3, j = (e - CTX_MCO) % 3;
    if (!act_c) return po[i * 3 + j];
    return pc[i * 3 + 0] * po[0 * 3 + j] + pc[i * 3 + 1] * po[1 * 3 + j] + pc[i * 3 + 2] * po[2 * 3 + j];
  }
  return intr[e - CTX_IN];
}
// Phase B entries [81,135): No = Mc dRo_k, Nb = Mco dRb_k; read Mc/Mco from the context itself.
MCBA_HD double ctx_entry_b(int e, const double* ctx, const double* po, const double* pb, bool act_b) {
  const bool is_o = e < CTX_NB;
  const int q = e - (is_o ? CTX_NO : CTX_NB);
  const int k = q / 9, i = (q % 9) / 3, j = q % 3;
  if (!is_o && !act_b) return 0.0;
  const double* M = ctx + (is_o ? CTX_MC : CTX_MCO);
  const double* D = (is_o ? po : pb) + 9 + 9 * k;
  return M[i * 3 + 0] * D[0 * 3 + j] + M[i * 3 + 1] * D[1 * 3 + j] + M[i * 3 + 2] * D[2 * 3 + j];
}

// Cover of the upper triangle (in 4-column groups) by products rows {r0, r1} x columns {c0, c1}; own = which of the four
// blocks (bit 2 * row_half + col_half) this product stores (a block may be computed twice in the small covers).
template <int NG> struct RectCover;
template <> struct RectCover<7> {   // Fano plane
  static constexpr int NR = 7;
  MCBA_HD static constexpr int r0(int i) { return i; }
  MCBA_HD static constexpr int r1(int i) { return (i + 1) % 7; }
  MCBA_HD static constexpr int c0(int i) { return i; }
  MCBA_HD static constexpr int c1(int i) { return (i + 3) % 7; }
  MCBA_HD static constexpr int own(int) { return 15; }
};
template <> struct RectCover<5> {   // (0,0)(0,2)(1,0)(1,2) | (3,3)(3,0)(4,3)(4,0) | (1,3)(1,4)(2,3)(2,4) | (1,1)(2,2) | (4,4)
  static constexpr int NR = 5;
  MCBA_HD static constexpr int r0(int i) { return i == 0 ? 0 : i == 1 ? 3 : i == 2 ? 1 : i == 3 ? 1 : 4; }
  MCBA_HD static constexpr int r1(int i) { return i == 0 ? 1 : i == 1 ? 4 : i == 2 ? 2 : i == 3 ? 2 : 4; }
  MCBA_HD static constexpr int c0(int i) { return i == 0 ? 0 : i == 1 ? 3 : i == 2 ? 3 : i == 3 ? 1 : 4; }
  MCBA_HD static constexpr int c1(int i) { return i == 0 ? 2 : i == 1 ? 0 : i == 2 ? 4 : i == 3 ? 2 : 4; }
  MCBA_HD static constexpr int own(int i) { return i < 3 ? 15 : i == 3 ? 9 : 1; }
};
template <> struct RectCover<4> {   // aligned 8x8 tiles: (0,0)(0,1)(1,1) | (0,2)(0,3)(1,2)(1,3) | (2,2)(2,3)(3,3)
  static constexpr int NR = 3;
  MCBA_HD static constexpr int r0(int i) { return i == 2 ? 2 : 0; }
  MCBA_HD static constexpr int r1(int i) { return i == 2 ? 3 : 1; }
  MCBA_HD static constexpr int c0(int i) { return i == 0 ? 0 : 2; }
  MCBA_HD static constexpr int c1(int i) { return i == 0 ? 1 : 3; }
  MCBA_HD static constexpr int own(int i) { return i == 1 ? 15 : 11; }   // not the mirrored (1,0) / (3,2)
};

// Record index (Cols<STAGE>::cidx) of entry h of the accumulator pair that lane `lane` holds for product r of the cover
// (m8n8k4 accumulator layout: row lane / 4, columns 2 (lane % 4) + h; the lower half-warp's rows belong to the product's
// first row group, the upper half-warp's to the second; columns 0-3 to its first column group, 4-7 to the second);
// -1: nothing to store (mirrored half of a diagonal block, padding column, residual x residual, block owned elsewhere).
template <int STAGE> MCBA_HD int k1_store_index(int r, int lane, int h) {
  typedef Cols<STAGE> C;
  typedef RectCover<(C::NC + 3) / 4> RCV;
  const int g = lane >> 2, q = lane & 3, hi = lane >> 4, c4 = g & 3, ch = q >> 1;
  int gr = 0, gc = 0, own = 0;
  for (int rr = 0; rr < RCV::NR; ++rr)
    if (rr == r) { gr = hi ? RCV::r1(rr) : RCV::r0(rr); gc = ch ? RCV::c1(rr) : RCV::c0(rr); own = RCV::own(rr); }
  const int ea = 4 * gr + c4, eb = 4 * gc + ((2 * q + h) & 3);
  if (!((own >> (2 * hi + ch)) & 1) || ea >= C::NC || eb >= C::NC || (gr == gc && ea > eb)) return -1;
  return C::cidx(ea < eb ? ea : eb, ea < eb ? eb : ea);
}
// The 63 three-term products of the copy-free context: product e writes ctx[o] = sum_n ctx[m + n] * ctx[d + 3 n].
// e < 9: Mco = Rc Ro; e < 36: No_k = Mc dRo_k; else Nb_k = Mco dRb_k (needs Mco: computed in an earlier round).
MCBA_HD void ctx_raw_operands(int e, int& m, int& d, int& o) {
  if (e < 9) { m = CtxRaw::RC + 3 * (e / 3); d = CtxRaw::RO + e % 3; o = CtxRaw::MCO + e; }
  else if (e < 36) { const int f = e - 9, k = f / 9, i = (f % 9) / 3, j = f % 3; m = CtxRaw::MC + 3 * i; d = CtxRaw::PO + 9 + 9 * k + j; o = CtxRaw::NO + f; }
  else { const int f = e - 36, k = f / 9, i = (f % 9) / 3, j = f % 3; m = CtxRaw::MCO + 3 * i; d = CtxRaw::PB + 9 + 9 * k + j; o = CtxRaw::NB + f; }
}

MCBA_HD void mat3_vec(const double* M, const double* x, double* y) {
  y[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
  y[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
  y[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
}
MCBA_HD void mat3_vec_add(const double* M, const double* x, const double* t, double* y) {
  y[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2] + t[0];
  y[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2] + t[1];
  y[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2] + t[2];
}

// Distortion: (a,b) -> (xd,yd) and optionally d(xd,yd)/d(a,b) and the distortion-coefficient columns.
struct Distort {
  double xd, yd, dxa, dxb, dya, dyb;
  double cu[5], cv[5];   // d xd / d intr[4..8], d yd / d intr[4..8]
};
template <bool JAC>
MCBA_HD void distort(int model, const double* in, double a, double b, Distort& d) {
  const double r2 = a * a + b * b;
  if (model == 0) {
    const double k1 = in[4], k2 = in[5], p1 = in[6], p2 = in[7], k3 = in[8];
    const double r4 = r2 * r2, r6 = r4 * r2;
    const double rc = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
    d.xd = a * rc + 2.0 * p1 * a * b + p2 * (r2 + 2.0 * a * a);
    d.yd = b * rc + p1 * (r2 + 2.0 * (b * b)) + 2.0 * p2 * a * b;
    if (JAC) {
      const double drc = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;   // d rc / d r2
      d.dxa = rc + 2.0 * a * a * drc + 2.0 * p1 * b + 6.0 * p2 * a;
      d.dxb = 2.0 * a * b * drc + 2.0 * p1 * a + 2.0 * p2 * b;
      d.dya = 2.0 * a * b * drc + 2.0 * p1 * a + 2.0 * p2 * b;
      d.dyb = rc + 2.0 * b * b * drc + 6.0 * p1 * b + 2.0 * p2 * a;
      d.cu[0] = a * r2; d.cu[1] = a * r4; d.cu[2] = 2.0 * a * b;       d.cu[3] = r2 + 2.0 * a * a; d.cu[4] = a * r6;
      d.cv[0] = b * r2; d.cv[1] = b * r4; d.cv[2] = r2 + 2.0 * b * b;  d.cv[3] = 2.0 * a * b;      d.cv[4] = b * r6;
    }
  } else {
    const double k1 = in[4], k2 = in[5], k3 = in[6], k4 = in[7];
    const double r = sqrt(r2);
    const double th = atan(r);
    const double t2 = th * th, t3 = t2 * th, t4 = t2 * t2, t5 = t4 * th, t6 = t3 * t3, t7 = t6 * th, t8 = t4 * t4, t9 = t8 * th;
    const double thd = th + k1 * t3 + k2 * t5 + k3 * t7 + k4 * t9;
    const bool far = r > 1e-8;
    const double inv_r = far ? 1.0 / r : 1.0;
    const double cd = far ? thd * inv_r : 1.0;
    d.xd = a * cd;
    d.yd = b * cd;
    if (JAC) {
      if (far) {
        const double dthd = 1.0 + 3.0 * k1 * t2 + 5.0 * k2 * t4 + 7.0 * k3 * t6 + 9.0 * k4 * t8;
        // d cd / d r = (dthd / (1+r^2) - cd) / r ; d r / d a = a / r
        const double dcd = (dthd / (1.0 + r2) - cd) * inv_r * inv_r;   // (d cd / d r) / r
        d.dxa = cd + a * a * dcd; d.dxb = a * b * dcd;
        d.dya = a * b * dcd;      d.dyb = cd + b * b * dcd;
        d.cu[0] = a * t3 * inv_r; d.cu[1] = a * t5 * inv_r; d.cu[2] = a * t7 * inv_r; d.cu[3] = a * t9 * inv_r; d.cu[4] = 0.0;
        d.cv[0] = b * t3 * inv_r; d.cv[1] = b * t5 * inv_r; d.cv[2] = b * t7 * inv_r; d.cv[3] = b * t9 * inv_r; d.cv[4] = 0.0;
      } else {
        d.dxa = 1.0; d.dxb = 0.0; d.dya = 0.0; d.dyb = 1.0;
        for (int i = 0; i < 5; ++i) { d.cu[i] = 0.0; d.cv[i] = 0.0; }
      }
    }
  }
}

// ceres::HuberLoss(a) + Corrector [recalled]: rho'' <= 0 everywhere, so residual and Jacobian are both
// scaled by sqrt(rho'); block cost = rho(s)/2.
MCBA_HD void huber_weight(double a, double s, double& w, double& half_rho) {
  if (a > 0.0 && s > a * a) {
    const double r = sqrt(s);
    half_rho = 0.5 * (2.0 * a * r - a * a);
#ifdef __CUDA_ARCH__
    w = sqrt(a) * rsqrt(r);          // sqrt(a / r) with one dependent special-function op instead of div + sqrt
#else
    w = sqrt(fmax(DBL_MIN, a / r));
#endif
  } else {
    half_rho = 0.5 * s;
    w = 1.0;
  }
}

// Residual only (candidate-cost pass): returns 1/2 rho(|r|^2); ru, rv are the raw residuals.
MCBA_HD double obs_cost(const double* ctx, int model, const double* X0, double u, double v, double huber_a,
                        double& ru, double& rv) {
  double X1[3], X2[3], X3[3];
  mat3_vec_add(ctx + CTX_RB, X0, ctx + CTX_TB, X1);
  mat3_vec_add(ctx + CTX_RO, X1, ctx + CTX_TO, X2);
  mat3_vec_add(ctx + CTX_RC, X2, ctx + CTX_TC, X3);
  const double iz = rcp(X3[2]);
  Distort d;
  distort<false>(model, ctx + CTX_IN, X3[0] * iz, X3[1] * iz, d);
  const double* in = ctx + CTX_IN;
  ru = in[0] * d.xd + in[2] - u;
  rv = in[1] * d.yd + in[3] - v;
  double w, hr;
  huber_weight(huber_a, ru * ru + rv * rv, w, hr);
  return hr;
}

// Reprojection error of one corner with the reference's intermediate roundings (the error outputs of MC-Calib, not its
// cost functors): the object-pose transform is stored in cv::Point3f (transform3DPts, McCalib/src/geometrytools.cpp:
// 328-352: t + r1 x + r2 y + r3 z evaluated in double, assigned to float), cv::projectPoints / cv::fisheye::projectPoints
// compute in double from those float points and return cv::Point2f, and computeDistanceBetweenPoints
// (McCalib/src/McCalib.cpp:2150-2160) subtracts the two cv::Point2f in float, squares and adds in double (std::pow
// promotes), takes the double sqrt and stores a float.
MCBA_HD float obs_error_ref(const double* ctx, int model, const double* X0, float u, float v) {
  double X1[3], X3[3];
  mat3_vec_add(ctx + CTX_RB, X0, ctx + CTX_TB, X1);
  const double* R = ctx + CTX_RO; const double* t = ctx + CTX_TO;
  double X2[3];
  for (int i = 0; i < 3; ++i) X2[i] = (double)(float)(t[i] + R[3 * i] * X1[0] + R[3 * i + 1] * X1[1] + R[3 * i + 2] * X1[2]);
  mat3_vec_add(ctx + CTX_RC, X2, ctx + CTX_TC, X3);
  const double iz = 1.0 / X3[2];
  Distort d;
  distort<false>(model, ctx + CTX_IN, X3[0] * iz, X3[1] * iz, d);
  const double* in = ctx + CTX_IN;
  const float pu = (float)(in[0] * d.xd + in[2]), pv = (float)(in[1] * d.yd + in[3]);
#ifdef __CUDA_ARCH__
  const float dx = __fsub_rn(u, pu), dy = __fsub_rn(v, pv);
#else
  const volatile float dxv = u - pu, dyv = v - pv;
  const float dx = dxv, dy = dyv;
#endif
  return (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy);
}

// Residual + Jacobian of one corner. Emits loss-corrected columns through sink.put(local_col, ju, jv)
// in the local order of Cols<STAGE>, the corrected residual as column R0, and returns 1/2 rho.
template <int STAGE, class Sink, class L = CtxBuilt>
MCBA_HD double obs_jacobian(const double* ctx, bool act_c, bool act_b, int model, const double* X0, double u,
                            double v, double huber_a, Sink& sink) {
  typedef StageTraits<STAGE> T;
  typedef Cols<STAGE> C;
  double X1[3], X2[3], X3[3];
  mat3_vec_add(ctx + L::RB, X0, ctx + L::TB, X1);
  mat3_vec_add(ctx + L::RO, X1, ctx + L::TO, X2);
  mat3_vec_add(ctx + L::RC, X2, ctx + L::TC, X3);
  const double iz = rcp(X3[2]);
  const double a = X3[0] * iz, b = X3[1] * iz;
  Distort d;
  distort<true>(model, ctx + L::IN, a, b, d);
  const double* in = ctx + L::IN;
  const double fx = in[0], fy = in[1];
  const double ru = fx * d.xd + in[2] - u;
  const double rv = fy * d.yd + in[3] - v;
  double w, hr;
  huber_weight(huber_a, ru * ru + rv * rv, w, hr);
  // d residual / d X3 (2x3), loss weight folded in
  const double wfx = w * fx, wfy = w * fy;
  const double Pu[3] = {wfx * d.dxa * iz, wfx * d.dxb * iz, -wfx * (d.dxa * a + d.dxb * b) * iz};
  const double Pv[3] = {wfy * d.dya * iz, wfy * d.dyb * iz, -wfy * (d.dya * a + d.dyb * b) * iz};
  double g[3];
  if (T::CAM) {
    if (act_c) {
      for (int k = 0; k < 3; ++k) {
        mat3_vec(ctx + L::DRC + 9 * k, X2, g);
        sink.put(C::C0 + k, Pu[0] * g[0] + Pu[1] * g[1] + Pu[2] * g[2], Pv[0] * g[0] + Pv[1] * g[1] + Pv[2] * g[2]);
      }
      for (int k = 0; k < 3; ++k) sink.put(C::C0 + 3 + k, Pu[k], Pv[k]);
    } else {
      for (int k = 0; k < 6; ++k) sink.put(C::C0 + k, 0.0, 0.0);
    }
  }
  if (T::INTR) {
    sink.put(C::I0 + 0, w * d.xd, 0.0);
    sink.put(C::I0 + 1, 0.0, w * d.yd);
    sink.put(C::I0 + 2, w, 0.0);
    sink.put(C::I0 + 3, 0.0, w);
    for (int k = 0; k < 5; ++k) sink.put(C::I0 + 4 + k, wfx * d.cu[k], wfy * d.cv[k]);
  }
  if (T::BOARD) {
    if (act_b) {
      const double* M = ctx + L::MCO;
      for (int k = 0; k < 3; ++k) {
        mat3_vec(ctx + L::NB + 9 * k, X0, g);
        sink.put(C::B0 + k, Pu[0] * g[0] + Pu[1] * g[1] + Pu[2] * g[2], Pv[0] * g[0] + Pv[1] * g[1] + Pv[2] * g[2]);
      }
      for (int k = 0; k < 3; ++k)
        sink.put(C::B0 + 3 + k, Pu[0] * M[k] + Pu[1] * M[3 + k] + Pu[2] * M[6 + k],
                 Pv[0] * M[k] + Pv[1] * M[3 + k] + Pv[2] * M[6 + k]);
    } else {
      for (int k = 0; k < 6; ++k) sink.put(C::B0 + k, 0.0, 0.0);
    }
  }
  {
    const double* M = ctx + L::MC;
    for (int k = 0; k < 3; ++k) {
      mat3_vec(ctx + L::NO + 9 * k, X1, g);
      sink.put(C::E0 + k, Pu[0] * g[0] + Pu[1] * g[1] + Pu[2] * g[2], Pv[0] * g[0] + Pv[1] * g[1] + Pv[2] * g[2]);
    }
    for (int k = 0; k < 3; ++k)
      sink.put(C::E0 + 3 + k, Pu[0] * M[k] + Pu[1] * M[3 + k] + Pu[2] * M[6 + k],
               Pv[0] * M[k] + Pv[1] * M[3 + k] + Pv[2] * M[6 + k]);
  }
  sink.put(C::R0, w * ru, w * rv);
  return hr;
}

// 6x6 Cholesky (lower, row-major packed in a 36 array; upper part untouched). Returns false if not PD.
MCBA_HD bool chol6(double* A) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= A[j * 6 + k] * A[j * 6 + k];
    if (!(s > 0.0) || !(s < DBL_MAX)) return false;
    const double l = sqrt(s), il = 1.0 / l;
    A[j * 6 + j] = l;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = t * il;
    }
  }
  return true;
}
// Same factorisation without the early return (the array then stays in registers on the device): on a non-positive
// pivot the flag is cleared and the factorisation continues with 1 in its place; inv[j] = 1 / L[j][j].
MCBA_HD bool chol6_inv(double (&A)[36], double (&inv)[6]) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= A[j * 6 + k] * A[j * 6 + k];
    const bool bad = !(s > 0.0) || !(s < DBL_MAX);
    ok = ok && !bad;
    const double l = bad ? 1.0 : sqrt(s);
    const double il = 1.0 / l;
    A[j * 6 + j] = l; inv[j] = il;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = t * il;
    }
  }
  return ok;
}
MCBA_HD void chol6_fwd_inv(const double (&L)[36], const double (&inv)[6], double (&b)[6]) {   // b <- L^-1 b, no division
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * b[k];
    b[i] = s * inv[i];
  }
}
MCBA_HD void chol6_fwd(const double* L, double* b) {   // b <- L^-1 b
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * b[k];
    b[i] = s / L[i * 6 + i];
  }
}
MCBA_HD void chol6_bwd(const double* L, double* b) {   // b <- L^-T b
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * b[k];
    b[i] = s / L[i * 6 + i];
  }
}

}  // namespace mcba
