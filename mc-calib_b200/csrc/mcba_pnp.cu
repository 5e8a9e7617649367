// mcba_pnp.cu — batched P3P RANSAC + pose-only refinement, one warp per board (object) observation (include/mcba_pnp.h).
//
//   k_pnp_prepare   one thread per corner: pixel -> normalised ray (a, b) and the "working" pixel the inlier test is
//                   made in (Brown: the detection itself; Kannala: the pinhole re-projection of the undistorted ray,
//                   geometrytools.cpp:728-742), plus the per-camera working intrinsics (Kannala: zero distortion)
//   k_pnp_ransac    one warp per observation: lanes evaluate 32 hypotheses at a time (draw 4 corners, three-point pose +
//                   fourth-point disambiguation, inlier count over all corners of the observation); the sequential
//                   stopping rule of ransacP3P (geometrytools.cpp:243-300) is then replayed over the 32 counts in draw
//                   order, so the outcome is the one a sequential run over the same draws would produce. Then the inlier
//                   mask of the winner and a Levenberg-Marquardt refinement of the pose on the inliers (lanes stride
//                   over corners, 6x6 normal equations reduced by shuffles in a fixed order).
// fp64 throughout; latency/compute bound (a few hundred bytes per observation are read once).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mcba_pnp.h"
#include "mcba_pnp_math.cuh"

namespace mcba {

struct PnpArgs {
  const float* u; const float* v; const int32_t* pt; const float* pts;
  const long long* bobs_ptr; const int32_t* bobs_cam; const int32_t* cam_model; const double* intr;
  double2* ab; double2* w; uint8_t* ok; double* wintr;
  long long n_obs; int n_bobs; int n_cam;
  double thr2; int it; double p; int refine; int refine_it; unsigned long long seed;
  double* pose; int* n_inl; uint8_t* mask; int* best_h; int* n_hyp; double* pose0;
};

__global__ void k_pnp_wintr(PnpArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_cam) return;
  for (int k = 0; k < 9; ++k) a.wintr[9 * c + k] = (k < 4 || a.cam_model[c] == 0) ? a.intr[9 * c + k] : 0.0;
}

__global__ void k_pnp_prepare(PnpArgs a) {
  // one warp per observation run so that the camera lookup is per run; lanes stride over its corners
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= a.n_bobs) return;
  const int cam = a.bobs_cam[b];
  const int model = a.cam_model[cam];
  double in[9];
  for (int k = 0; k < 9; ++k) in[k] = a.intr[9 * cam + k];
  for (long long i = a.bobs_ptr[b] + lane; i < a.bobs_ptr[b + 1]; i += 32) {
    const double uu = (double)a.u[i], vv = (double)a.v[i];
    double x = 0.0, y = 0.0;
    const bool ok = undistort_pixel(model, in, uu, vv, x, y);
    a.ab[i] = make_double2(x, y);
    a.w[i] = (model == 0) ? make_double2(uu, vv) : make_double2(x * in[0] + in[2], y * in[1] + in[3]);
    a.ok[i] = ok ? 1 : 0;
  }
}

__device__ __forceinline__ bool pnp_is_inlier(const double* R, const double* t, const double* in, const float* p3,
                                              double2 w, double thr2) {
  const double X0[3] = {(double)p3[0], (double)p3[1], (double)p3[2]};
  double X[3], pu, pv;
  mat3_vec_add(R, X0, t, X);
  if (!project_brown(in, X, pu, pv)) return false;
  const double du = pu - w.x, dv = pv - w.y;
  return du * du + dv * dv < thr2;
}

// H (upper 21) / g (6) / cost of the pose-only problem at `pose`, over the inliers of the run; all lanes get the sums
__device__ void pnp_normal_eq(const PnpArgs& a, long long o0, int n, const double* in, const double* pose, int lane,
                              double* H, double* g, double& cost) {
  double prep[PREP];
  prep_block(pose, prep);
  for (int k = 0; k < 21; ++k) H[k] = 0.0;
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  cost = 0.0;
  for (int i = lane; i < n; i += 32) {
    if (!a.mask[o0 + i]) continue;
    const float* p3 = a.pts + 3 * (size_t)a.pt[o0 + i];
    const double X0[3] = {(double)p3[0], (double)p3[1], (double)p3[2]};
    const double2 w = a.w[o0 + i];
    double r[2], Ju[6], Jv[6];
    if (!pnp_residual_jac(prep, in, X0, w.x, w.y, r, Ju, Jv)) { cost += 1e12; continue; }   // behind the camera: reject the step
    cost += r[0] * r[0] + r[1] * r[1];
    int k = 0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
#pragma unroll
      for (int q = p; q < 6; ++q) H[k++] += Ju[p] * Ju[q] + Jv[p] * Jv[q];
      g[p] += Ju[p] * r[0] + Jv[p] * r[1];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    for (int k = 0; k < 21; ++k) H[k] += __shfl_xor_sync(0xffffffffu, H[k], off);
    for (int k = 0; k < 6; ++k) g[k] += __shfl_xor_sync(0xffffffffu, g[k], off);
    cost += __shfl_xor_sync(0xffffffffu, cost, off);
  }
}

constexpr int PNP_WARPS = 4;
__global__ void __launch_bounds__(PNP_WARPS * 32) k_pnp_ransac(PnpArgs a) {
  const int b = blockIdx.x * PNP_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= a.n_bobs) return;
  const long long o0 = a.bobs_ptr[b];
  const int n = (int)(a.bobs_ptr[b + 1] - o0);
  const int cam = a.bobs_cam[b];
  double in[9];
  for (int k = 0; k < 9; ++k) in[k] = a.wintr[9 * cam + k];
  RansacState st;
  ransac_init(st, a.it);
  double bR[9], bt[3];
  for (int k = 0; k < 9; ++k) bR[k] = 0.0;
  for (int k = 0; k < 3; ++k) bt[k] = 0.0;
  if (n >= 4) {
    for (int round = 0; ransac_continue(st, a.it); ++round) {
      const int h = round * 32 + lane;
      double R[9], t[3];
      bool ok = h < a.it;
      if (ok) {
        int idx[4];
        ransac_sample4(a.seed, (unsigned long long)b, (unsigned long long)h, n, idx);
        double ab4[4][2], X4[4][3];
        for (int k = 0; k < 4; ++k) {
          const double2 q = a.ab[o0 + idx[k]];
          ab4[k][0] = q.x; ab4[k][1] = q.y;
          ok = ok && a.ok[o0 + idx[k]];
          const float* p3 = a.pts + 3 * (size_t)a.pt[o0 + idx[k]];
          X4[k][0] = (double)p3[0]; X4[k][1] = (double)p3[1]; X4[k][2] = (double)p3[2];
        }
        const double2 w4 = a.w[o0 + idx[3]];
        const double w4a[2] = {w4.x, w4.y};
        ok = ok && p3p_hypothesis(ab4, X4, in, w4a, R, t);
      }
      int count = 0;
      if (ok)
        for (int i = 0; i < n; ++i)
          count += (a.ok[o0 + i] && pnp_is_inlier(R, t, in, a.pts + 3 * (size_t)a.pt[o0 + i], a.w[o0 + i], a.thr2)) ? 1 : 0;
      // replay the sequential rule over this round's 32 counts, in draw order (identical on every lane)
      int new_best_lane = -1;
      for (int l = 0; l < 32; ++l) {
        const int cl = __shfl_sync(0xffffffffu, count, l);
        if (!ransac_continue(st, a.it)) continue;
        const int prev = st.best_h;
        ransac_update(st, round * 32 + l, cl, n, a.p);
        if (st.best_h != prev) new_best_lane = l;
      }
      if (new_best_lane >= 0) {
        for (int k = 0; k < 9; ++k) bR[k] = __shfl_sync(0xffffffffu, R[k], new_best_lane);
        for (int k = 0; k < 3; ++k) bt[k] = __shfl_sync(0xffffffffu, t[k], new_best_lane);
      }
    }
  }
  // inlier mask of the selected hypothesis
  int cnt = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool inl = false;
    if (i < n && st.best_h >= 0)
      inl = a.ok[o0 + i] && pnp_is_inlier(bR, bt, in, a.pts + 3 * (size_t)a.pt[o0 + i], a.w[o0 + i], a.thr2);
    if (i < n) a.mask[o0 + i] = inl ? 1 : 0;
    cnt += __popc(__ballot_sync(0xffffffffu, inl));
  }
  __syncwarp();
  double pose[6] = {0, 0, 0, 0, 0, 0};
  if (st.best_h >= 0) { rot_to_rvec(bR, pose); pose[3] = bt[0]; pose[4] = bt[1]; pose[5] = bt[2]; }
  if (a.pose0 && lane < 6) a.pose0[6 * (size_t)b + lane] = pose[lane];
  if (a.refine && st.best_h >= 0 && cnt >= 4) {
    // Levenberg-Marquardt on the 6 pose parameters, damping H + lambda diag(H) as cv::solvePnP's iterative solver
    double H[21], g[6], cost;
    pnp_normal_eq(a, o0, n, in, pose, lane, H, g, cost);
    double lambda = 1e-3;
    for (int iter = 0; iter < a.refine_it; ++iter) {
      double A[36], d[6];
      int k = 0;
      for (int p = 0; p < 6; ++p)
        for (int q = p; q < 6; ++q) { A[q * 6 + p] = H[k]; A[p * 6 + q] = H[k]; ++k; }
      for (int p = 0; p < 6; ++p) { A[p * 6 + p] *= (1.0 + lambda); d[p] = -g[p]; }
      if (!chol6(A)) { lambda *= 10.0; if (lambda > 1e16) break; continue; }
      chol6_fwd(A, d); chol6_bwd(A, d);
      double cand[6], dn = 0.0, xn = 0.0;
      for (int p = 0; p < 6; ++p) { cand[p] = pose[p] + d[p]; dn += d[p] * d[p]; xn += pose[p] * pose[p]; }
      double H2[21], g2[6], cost2;
      pnp_normal_eq(a, o0, n, in, cand, lane, H2, g2, cost2);
      if (cost2 < cost) {
        for (int p = 0; p < 6; ++p) { pose[p] = cand[p]; g[p] = g2[p]; }
        for (int q = 0; q < 21; ++q) H[q] = H2[q];
        const bool small = (cost - cost2) <= 1e-14 * cost || dn <= 1e-24 * (xn + 1e-24);
        cost = cost2;
        lambda = fmax(lambda * 0.1, 1e-16);
        if (small) break;
      } else {
        if (!(dn > 1e-30 * (xn + 1e-30))) break;     // the step no longer changes the pose
        lambda *= 10.0;
        if (lambda > 1e16) break;
      }
    }
  }
  if (lane < 6) a.pose[6 * (size_t)b + lane] = pose[lane];
  if (lane == 0) {
    a.n_inl[b] = cnt;
    if (a.best_h) a.best_h[b] = st.best_h;
    if (a.n_hyp) a.n_hyp[b] = st.countit;
  }
}

static thread_local std::string g_pnp_err;

}  // namespace mcba

using namespace mcba;

#define PCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_pnp_err = std::string(#x) + ": " + cudaGetErrorString(e_); goto fail; } } while (0)

extern "C" {

void mcba_pnp_default_options(mcba_pnp_options* o) {
  o->threshold_px = 10.0; o->max_iterations = 1000; o->confidence = 0.99; o->refine = 1; o->refine_max_iterations = 20; o->seed = 0;
}

const char* mcba_pnp_last_error(void) { return g_pnp_err.c_str(); }

int mcba_pnp_ransac(const mcba_pnp_problem* P, const mcba_pnp_options* opt, int device, mcba_pnp_result* res) {
  g_pnp_err.clear();
  if (!P || !opt || !res || !res->pose || !res->n_inliers || !res->inlier_mask) { g_pnp_err = "null argument"; return MCBA_ERR_INVALID; }
  if (P->n_obs < 0 || P->n_bobs < 0 || P->n_cam <= 0 || P->n_pts <= 0 || P->n_bobs > 0x7fffffffLL) { g_pnp_err = "bad sizes"; return MCBA_ERR_INVALID; }
  if (!(opt->threshold_px > 0.0) || opt->max_iterations < 0 || !(opt->confidence > 0.0 && opt->confidence < 1.0)) { g_pnp_err = "bad options"; return MCBA_ERR_INVALID; }
  if (P->n_bobs > 0 && (P->bobs_ptr[0] != 0 || P->bobs_ptr[P->n_bobs] != P->n_obs)) { g_pnp_err = "bobs_ptr must start at 0 and end at n_obs"; return MCBA_ERR_INVALID; }
  for (int64_t b = 0; b < P->n_bobs; ++b) {
    if (P->bobs_ptr[b + 1] < P->bobs_ptr[b]) { g_pnp_err = "bobs_ptr must be non-decreasing"; return MCBA_ERR_INVALID; }
    if (P->bobs_cam[b] < 0 || P->bobs_cam[b] >= P->n_cam) { g_pnp_err = "bobs_cam out of range"; return MCBA_ERR_INVALID; }
  }
  for (int64_t i = 0; i < P->n_obs; ++i)
    if (P->pt_idx[i] < 0 || P->pt_idx[i] >= P->n_pts) { g_pnp_err = "pt_idx out of range"; return MCBA_ERR_INVALID; }
  for (int c = 0; c < P->n_cam; ++c)
    if (P->cam_model[c] != 0 && P->cam_model[c] != 1) { g_pnp_err = "cam_model must be 0 (Brown) or 1 (Kannala)"; return MCBA_ERR_INVALID; }
  res->kernel_ms = 0.0; res->total_ms = 0.0; res->kernel_launches = 0;
  if (P->n_bobs == 0) return MCBA_OK;
  const auto t0 = std::chrono::steady_clock::now();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); g_pnp_err = "no such CUDA device (libmcba has no CPU fallback)"; return MCBA_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) { cudaGetLastError(); g_pnp_err = "libmcba is built for sm_100a only"; return MCBA_ERR_CUDA; }
  PnpArgs a; std::memset(&a, 0, sizeof(a));
  const size_t no = (size_t)P->n_obs, nb = (size_t)P->n_bobs, nc = (size_t)P->n_cam;
  char* d_all = nullptr; cudaStream_t stream = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  // one device allocation, carved up (256-byte aligned pieces)
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_u = carve(no * 4), o_v = carve(no * 4), o_pt = carve(no * 4), o_pts = carve((size_t)P->n_pts * 12);
  const size_t o_bp = carve((nb + 1) * 8), o_bc = carve(nb * 4), o_cm = carve(nc * 4), o_in = carve(nc * 72), o_wi = carve(nc * 72);
  const size_t o_ab = carve(no * 16), o_w = carve(no * 16), o_ok = carve(no), o_mask = carve(no);
  const size_t o_pose = carve(nb * 48), o_pose0 = carve(nb * 48), o_ni = carve(nb * 4), o_bh = carve(nb * 4), o_nh = carve(nb * 4);
  float ms = 0.f;
  PCU(cudaSetDevice(device));
  PCU(cudaMalloc((void**)&d_all, off));
  PCU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  PCU(cudaEventCreate(&e0)); PCU(cudaEventCreate(&e1));
  PCU(cudaMemcpyAsync(d_all + o_u, P->u, no * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_v, P->v, no * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_pt, P->pt_idx, no * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_pts, P->pts_xyz, (size_t)P->n_pts * 12, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_bp, P->bobs_ptr, (nb + 1) * 8, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_bc, P->bobs_cam, nb * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_cm, P->cam_model, nc * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_in, P->intrinsics, nc * 72, cudaMemcpyHostToDevice, stream));
  a.u = (const float*)(d_all + o_u); a.v = (const float*)(d_all + o_v); a.pt = (const int32_t*)(d_all + o_pt);
  a.pts = (const float*)(d_all + o_pts); a.bobs_ptr = (const long long*)(d_all + o_bp); a.bobs_cam = (const int32_t*)(d_all + o_bc);
  a.cam_model = (const int32_t*)(d_all + o_cm); a.intr = (const double*)(d_all + o_in); a.wintr = (double*)(d_all + o_wi);
  a.ab = (double2*)(d_all + o_ab); a.w = (double2*)(d_all + o_w); a.ok = (uint8_t*)(d_all + o_ok); a.mask = (uint8_t*)(d_all + o_mask);
  a.pose = (double*)(d_all + o_pose); a.pose0 = (double*)(d_all + o_pose0); a.n_inl = (int*)(d_all + o_ni);
  a.best_h = (int*)(d_all + o_bh); a.n_hyp = (int*)(d_all + o_nh);
  a.n_obs = P->n_obs; a.n_bobs = (int)P->n_bobs; a.n_cam = P->n_cam;
  a.thr2 = opt->threshold_px * opt->threshold_px; a.it = opt->max_iterations; a.p = opt->confidence;
  a.refine = opt->refine; a.refine_it = opt->refine_max_iterations; a.seed = opt->seed;
  PCU(cudaEventRecord(e0, stream));
  k_pnp_wintr<<<(P->n_cam + 127) / 128, 128, 0, stream>>>(a);
  k_pnp_prepare<<<(unsigned)((nb * 32 + 255) / 256), 256, 0, stream>>>(a);
  k_pnp_ransac<<<(unsigned)((nb + PNP_WARPS - 1) / PNP_WARPS), PNP_WARPS * 32, 0, stream>>>(a);
  PCU(cudaGetLastError());
  PCU(cudaEventRecord(e1, stream));
  res->kernel_launches = 3;
  PCU(cudaMemcpyAsync(res->pose, a.pose, nb * 48, cudaMemcpyDeviceToHost, stream));
  PCU(cudaMemcpyAsync(res->n_inliers, a.n_inl, nb * 4, cudaMemcpyDeviceToHost, stream));
  PCU(cudaMemcpyAsync(res->inlier_mask, a.mask, no, cudaMemcpyDeviceToHost, stream));
  if (res->best_hypothesis) PCU(cudaMemcpyAsync(res->best_hypothesis, a.best_h, nb * 4, cudaMemcpyDeviceToHost, stream));
  if (res->n_hypotheses) PCU(cudaMemcpyAsync(res->n_hypotheses, a.n_hyp, nb * 4, cudaMemcpyDeviceToHost, stream));
  if (res->unrefined_pose) PCU(cudaMemcpyAsync(res->unrefined_pose, a.pose0, nb * 48, cudaMemcpyDeviceToHost, stream));
  PCU(cudaStreamSynchronize(stream));
  PCU(cudaEventElapsedTime(&ms, e0, e1));
  res->kernel_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(stream); cudaFree(d_all);
  res->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return MCBA_OK;
fail:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (stream) cudaStreamDestroy(stream);
  if (d_all) cudaFree(d_all);
  cudaGetLastError();
  return MCBA_ERR_CUDA;
}

}  // extern "C"
