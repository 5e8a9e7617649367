// mcba_batched.cuh — many independent stage-1 problems in one set of launches (BASELINE config 5:
// Camera::refineIntrinsicCalibration for thousands of cameras, /root/reference/McCalib/src/Camera.cpp:268-305 called
// per camera from Calibration::refineIntrinsicAndPoseAllCam, McCalib.cpp:713-716).
//
// Every camera is its own Levenberg-Marquardt problem: e-blocks = its board-observation poses (one bobs each),
// f-block = its 9 intrinsics, reduced system 9x9. The observation kernels are the shared k_observations<1,*> with a
// per-camera gate; everything after them is per-camera / per-e-block small dense work, and the accept / reject /
// terminate decisions are taken on the device by one thread per problem with the same functions (mcba_lm.cuh) the
// single-problem host driver uses. The host only polls the number of still-active problems.
#pragma once
#include "mcba_kernels.cuh"
#include "mcba_lm.cuh"

namespace mcba {

typedef Cols<1> C1;   // local columns: intrinsics 0..8 | e-block 9..14 | residual 15
constexpr int B_REC = C1::NCOMP;
__device__ __forceinline__ int b_off(int a, int b) { return C1::cidx(a, b); }   // a <= b, index in the compact record
// per-camera scalar slots
enum { BC_COST = 0, BC_XE2, BC_GMAX, BC_ZG, BC_ZHZ, BC_SN, BC_XN, BC_FAIL, BC_N };

__global__ void k_b_init(LmState* lm, mcba_options opt, int n_cam, uint8_t* gate_jac, uint8_t* active) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  lm_init(lm[c], opt);
  gate_jac[c] = 1; active[c] = 1;
}

// Jacobi scale of the e-block columns (iteration 0 only)
__global__ void k_b_escale(const int4* __restrict__ bobs, const uint8_t* __restrict__ gate, const double* __restrict__ brec,
                           int n_pose, int jacobi, double* __restrict__ scale_e) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_pose || !gate[bobs[o].y]) return;
  for (int k = 0; k < 6; ++k) {
    const double h = brec[(size_t)o * B_REC + b_off(C1::E0 + k, C1::E0 + k)];
    scale_e[(size_t)o * 6 + k] = jacobi ? 1.0 / (1.0 + sqrt(h)) : 1.0;
  }
}

// Per camera, fixed-order sums over its board observations: F^T F (9x9), F^T r, cost, |x_e|^2, max |x_e - (x_e - g_e)|
__global__ void __launch_bounds__(64) k_b_cam_reduce(const uint8_t* __restrict__ gate, const int* __restrict__ cam_ptr,
                                                     const double* __restrict__ brec, const double* __restrict__ cost_bobs,
                                                     const double* __restrict__ pose, double* __restrict__ FFc,
                                                     double* __restrict__ gfc, double* __restrict__ sc) {
  const int c = blockIdx.x, t = threadIdx.x;
  if (!gate[c]) return;
  const int b0 = cam_ptr[c], b1 = cam_ptr[c + 1];
  if (t < 45) {
    int a = 0, rem = t;
    while (rem >= 9 - a) { rem -= 9 - a; ++a; }
    const int b = a + rem, off = b_off(a, b);
    double s = 0.0;
    for (int i = b0; i < b1; ++i) s += brec[(size_t)i * B_REC + off];
    FFc[(size_t)c * 81 + a * 9 + b] = s; FFc[(size_t)c * 81 + b * 9 + a] = s;
  } else if (t < 54) {
    const int a = t - 45, off = b_off(a, C1::R0);
    double s = 0.0;
    for (int i = b0; i < b1; ++i) s += brec[(size_t)i * B_REC + off];
    gfc[(size_t)c * 9 + a] = s;
  } else if (t == 54) {
    double s = 0.0;
    for (int i = b0; i < b1; ++i) s += cost_bobs[i];
    sc[(size_t)c * BC_N + BC_COST] = s;
  } else if (t == 55) {
    double s = 0.0, m = 0.0;
    for (int i = b0; i < b1; ++i)
      for (int k = 0; k < 6; ++k) {
        const double x = pose[(size_t)i * 6 + k], g = brec[(size_t)i * B_REC + b_off(C1::E0 + k, C1::R0)];
        s += x * x; m = fmax(m, fabs(x - (x + (-g))));
      }
    sc[(size_t)c * BC_N + BC_XE2] = s; sc[(size_t)c * BC_N + BC_GMAX] = m;
  }
}

// After an evaluation at the current point: cost / gradient norm into the LM state (iteration 0 or accepted step),
// Jacobi scale of the intrinsics columns at iteration 0.
__global__ void k_b_after_eval(LmState* lm, mcba_options opt, const uint8_t* __restrict__ gate, int n_cam, int first,
                               const double* __restrict__ FFc, const double* __restrict__ gfc, const double* __restrict__ sc,
                               const double* __restrict__ intr, double* __restrict__ scale_f) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam || !gate[c]) return;
  double xf2 = 0.0, gmax = sc[(size_t)c * BC_N + BC_GMAX];
  for (int k = 0; k < 9; ++k) {
    const double x = intr[(size_t)c * 9 + k], g = gfc[(size_t)c * 9 + k];
    xf2 += x * x; gmax = fmax(gmax, fabs(x - (x + (-g))));
    if (first) scale_f[(size_t)c * 9 + k] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(FFc[(size_t)c * 81 + k * 10])) : 1.0;
  }
  const double cost = sc[(size_t)c * BC_N + BC_COST];
  LmState s = lm[c];
  if (!(cost < DBL_MAX) || !(cost == cost)) { s.termination = MCBA_FAILURE; s.done = 1; lm[c] = s; return; }
  if (first) lm_iteration_zero(s, cost, gmax, sqrt(sc[(size_t)c * BC_N + BC_XE2] + xf2));
  else lm_accepted_eval(s, opt, cost, gmax);
  lm[c] = s;
}

__global__ void k_b_finalize(LmState* lm, mcba_options opt, int n_cam, uint8_t* active, int* n_active) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  LmState s = lm[c];
  int cont = 0;
  if (!s.done) { cont = lm_finalize(s, opt, nullptr); lm[c] = s; }
  active[c] = (uint8_t)cont;
  if (cont) atomicAdd(n_active, 1);
}

// Per e-block: LM diagonal, (E^T E + D^2) = L L^T, Zt = L^-1 s_e [E^T F s_f | E^T r]  (6 x 10)
__global__ void __launch_bounds__(128) k_b_eblock_factor(const int4* __restrict__ bobs, const uint8_t* __restrict__ active,
                                                         const LmState* __restrict__ lm, mcba_options opt,
                                                         const double* __restrict__ brec, const double* __restrict__ scale_e,
                                                         const double* __restrict__ scale_f, double* __restrict__ diag_e,
                                                         int n_pose, double* __restrict__ Lout, double* __restrict__ Zb,
                                                         int* __restrict__ cam_fail) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_pose) return;
  const int c = bobs[o].y;
  if (!active[c]) return;
  const double* rec = brec + (size_t)o * B_REC;
  const double radius = lm[c].radius;
  const bool reuse = lm[c].reuse_diagonal != 0;
  double L[36], se[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) se[k] = scale_e[(size_t)o * 6 + k];
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int r = 0; r <= p; ++r) { const double v = se[p] * rec[b_off(C1::E0 + r, C1::E0 + p)] * se[r]; L[p * 6 + r] = v; L[r * 6 + p] = v; }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double d;
    if (reuse) d = diag_e[(size_t)o * 6 + k];
    else { d = fmin(fmax(L[k * 7], opt.min_lm_diagonal), opt.max_lm_diagonal); diag_e[(size_t)o * 6 + k] = d; }
    L[k * 7] += d / radius;
  }
  const bool ok = chol6(L);
  if (!ok) atomicExch(cam_fail + c, 1);
#pragma unroll
  for (int e = 0; e < 36; ++e) Lout[(size_t)o * 36 + e] = ok ? L[e] : 0.0;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    double w[6];
    const double sf = (j < 9) ? scale_f[(size_t)c * 9 + j] : 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k] = se[k] * rec[(j < 9) ? b_off(j, C1::E0 + k) : b_off(C1::E0 + k, C1::R0)] * sf;
    if (ok) chol6_fwd(L, w);
#pragma unroll
    for (int k = 0; k < 6; ++k) Zb[(size_t)o * 60 + k * 10 + j] = ok ? w[k] : 0.0;
  }
}

// Per camera: reduced 9x9 system, Cholesky, z_f, candidate intrinsics and the f-part of the step scalars
__global__ void __launch_bounds__(64) k_b_schur_solve(const uint8_t* __restrict__ active, LmState* lm, mcba_options opt,
                                                      const int* __restrict__ cam_ptr, const double* __restrict__ FFc,
                                                      const double* __restrict__ gfc, const double* __restrict__ Zb,
                                                      const double* __restrict__ scale_f, double* __restrict__ diag_f,
                                                      const double* __restrict__ intr, double* __restrict__ intr_cand,
                                                      double* __restrict__ zf, double* __restrict__ sc, int* __restrict__ cam_fail) {
  __shared__ double S[9][9], rhs[9], sf[9], df[9];
  const int c = blockIdx.x, t = threadIdx.x;
  if (!active[c]) return;
  const int b0 = cam_ptr[c], b1 = cam_ptr[c + 1];
  const double radius = lm[c].radius;
  if (t < 9) {
    sf[t] = scale_f[(size_t)c * 9 + t];
    double d;
    if (lm[c].reuse_diagonal) d = diag_f[(size_t)c * 9 + t];
    else { d = fmin(fmax(sf[t] * sf[t] * FFc[(size_t)c * 81 + t * 10], opt.min_lm_diagonal), opt.max_lm_diagonal); diag_f[(size_t)c * 9 + t] = d; }
    df[t] = d;
  }
  __syncthreads();
  if (t < 45) {
    int a = 0, rem = t;
    while (rem >= 9 - a) { rem -= 9 - a; ++a; }
    const int b = a + rem;
    double s = 0.0;
    for (int i = b0; i < b1; ++i) {
      const double* z = Zb + (size_t)i * 60;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += z[k * 10 + a] * z[k * 10 + b];
    }
    double v = sf[a] * FFc[(size_t)c * 81 + a * 9 + b] * sf[b] - s;
    if (a == b) v += df[a] / radius;
    S[a][b] = v; S[b][a] = v;
  } else if (t < 54) {
    const int a = t - 45;
    double s = 0.0;
    for (int i = b0; i < b1; ++i) {
      const double* z = Zb + (size_t)i * 60;
#pragma unroll
      for (int k = 0; k < 6; ++k) s += z[k * 10 + a] * z[k * 10 + 9];
    }
    rhs[a] = sf[a] * gfc[(size_t)c * 9 + a] - s;
  }
  __syncthreads();
  if (t == 0) {
    bool ok = true;
    for (int j = 0; j < 9 && ok; ++j) {   // in-place lower Cholesky
      double d = S[j][j];
      for (int k = 0; k < j; ++k) d -= S[j][k] * S[j][k];
      if (!(d > 0.0) || !(d < DBL_MAX)) { ok = false; break; }
      const double l = sqrt(d);
      S[j][j] = l;
      for (int i = j + 1; i < 9; ++i) { double v = S[i][j]; for (int k = 0; k < j; ++k) v -= S[i][k] * S[j][k]; S[i][j] = v / l; }
    }
    double z[9];
    for (int i = 0; i < 9; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= S[i][k] * z[k]; z[i] = ok ? v / S[i][i] : 0.0; }
    for (int i = 8; i >= 0; --i) { double v = z[i]; for (int k = i + 1; k < 9; ++k) v -= S[k][i] * z[k]; z[i] = ok ? v / S[i][i] : 0.0; }
    if (!ok) atomicExch(cam_fail + c, 1);
    double zg = 0, zhz = 0, sn = 0, xn = 0;
    for (int i = 0; i < 9; ++i) {
      zf[(size_t)c * 9 + i] = z[i];
      double hz = 0.0;
      for (int k = 0; k < 9; ++k) hz += FFc[(size_t)c * 81 + i * 9 + k] * sf[k] * z[k];
      zg += z[i] * sf[i] * gfc[(size_t)c * 9 + i];
      zhz += z[i] * sf[i] * hz;
      const double x = intr[(size_t)c * 9 + i], xc = x + sf[i] * (-z[i]);
      intr_cand[(size_t)c * 9 + i] = xc;
      sn += (x - xc) * (x - xc); xn += xc * xc;
    }
    double* o = sc + (size_t)c * BC_N;
    o[BC_ZG] = zg; o[BC_ZHZ] = zhz; o[BC_SN] = sn; o[BC_XN] = xn;
    lm[c].reuse_diagonal = 1;
  }
}

// Per e-block back-substitution and candidate pose; part[o] as in k_backsubst
__global__ void __launch_bounds__(128) k_b_backsubst(const int4* __restrict__ bobs, const uint8_t* __restrict__ active,
                                                     const LmState* __restrict__ lm, const double* __restrict__ Zb,
                                                     const double* __restrict__ Lall, const double* __restrict__ zf,
                                                     const double* __restrict__ scale_e, const double* __restrict__ diag_e,
                                                     const double* __restrict__ pose, int n_pose,
                                                     double* __restrict__ pose_cand, double* __restrict__ part) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_pose) return;
  const int c = bobs[o].y;
  if (!active[c]) return;
  const double radius = lm[c].radius;
  const double* z = Zb + (size_t)o * 60;
  double zc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) zc[j] = zf[(size_t)c * 9 + j];
  double L[36], p[6], zo[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int q = 0; q <= r; ++q) L[r * 6 + q] = Lall[(size_t)o * 36 + r * 6 + q];
  double py = 0, pp = 0, pq = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double q = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) q += z[k * 10 + j] * zc[j];
    const double y = z[k * 10 + 9];
    p[k] = y - q; zo[k] = p[k]; py += p[k] * y; pp += p[k] * p[k]; pq += p[k] * q;
  }
  chol6_bwd(L, zo);
  double dz = 0, sn = 0, xn = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    dz += diag_e[(size_t)o * 6 + k] / radius * zo[k] * zo[k];
    const double x = pose[(size_t)o * 6 + k], xc = x + scale_e[(size_t)o * 6 + k] * (-zo[k]);
    pose_cand[(size_t)o * 6 + k] = xc;
    sn += (x - xc) * (x - xc); xn += xc * xc;
  }
  part[(size_t)o * 4 + 0] = py; part[(size_t)o * 4 + 1] = pp - dz + 2.0 * pq;
  part[(size_t)o * 4 + 2] = sn; part[(size_t)o * 4 + 3] = xn;
}

// Per camera: fixed-order sums of the e-block pieces and of the candidate cost, then the accept / reject decision.
__global__ void __launch_bounds__(32) k_b_decide(const uint8_t* __restrict__ active, LmState* lm, mcba_options opt,
                                                 const int* __restrict__ cam_ptr, const double* __restrict__ part,
                                                 const double* __restrict__ cost_bobs, const double* __restrict__ sc,
                                                 int* __restrict__ cam_fail, uint8_t* __restrict__ gate_jac) {
  __shared__ double r[5];
  const int c = blockIdx.x, t = threadIdx.x;
  if (!active[c]) { if (t == 0) gate_jac[c] = 0; return; }
  const int b0 = cam_ptr[c], b1 = cam_ptr[c + 1];
  if (t < 4) { double s = 0.0; for (int i = b0; i < b1; ++i) s += part[(size_t)i * 4 + t]; r[t] = s; }
  else if (t == 4) { double s = 0.0; for (int i = b0; i < b1; ++i) s += cost_bobs[i]; r[4] = s; }
  __syncwarp();
  if (t != 0) return;
  const double* f = sc + (size_t)c * BC_N;
  LmStep st;
  st.mcc = (r[0] + f[BC_ZG]) - 0.5 * (r[1] + f[BC_ZHZ]);
  st.step_norm = sqrt(r[2] + f[BC_SN]);
  st.cand_norm = sqrt(r[3] + f[BC_XN]);
  st.cand_cost = (r[4] == r[4] && r[4] < DBL_MAX) ? r[4] : DBL_MAX;
  st.solver_ok = (cam_fail[c] == 0) && (st.mcc == st.mcc) && (fabs(st.mcc) < DBL_MAX) && (st.step_norm == st.step_norm);
  cam_fail[c] = 0;
  LmState s = lm[c];
  const int d = lm_decide(s, opt, st);
  lm[c] = s;
  gate_jac[c] = (d == LM_ACCEPT) ? 1 : 0;
}

// Accepted problems: candidate becomes current
__global__ void k_b_commit(const int4* __restrict__ bobs, const uint8_t* __restrict__ gate, int n_pose, int n_cam,
                           const double* __restrict__ pose_cand, double* __restrict__ pose,
                           const double* __restrict__ intr_cand, double* __restrict__ intr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pose) {
    if (gate[bobs[i].y]) for (int k = 0; k < 6; ++k) pose[(size_t)i * 6 + k] = pose_cand[(size_t)i * 6 + k];
  } else if (i < n_pose + n_cam) {
    const int c = i - n_pose;
    if (gate[c]) for (int k = 0; k < 9; ++k) intr[(size_t)c * 9 + k] = intr_cand[(size_t)c * 9 + k];
  }
}

__global__ void k_b_cam_cost(const int* __restrict__ cam_ptr, const double* __restrict__ cost_bobs, int n_cam,
                             double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cam) return;
  double s = 0.0;
  for (int i = cam_ptr[c]; i < cam_ptr[c + 1]; ++i) s += cost_bobs[i];
  out[c] = s;
}

}  // namespace mcba
