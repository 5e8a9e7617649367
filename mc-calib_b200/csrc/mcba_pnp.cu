// mcba_pnp.cu — batched P3P RANSAC + pose-only refinement, one warp per board (object) observation (include/mcba_pnp.h).
//
//   k_pnp_prepare   one thread per corner: pixel -> normalised ray (a, b) and the "working" pixel the inlier test is
//                   made in (Brown: the detection itself; Kannala: the pinhole re-projection of the undistorted ray,
//                   geometrytools.cpp:728-742), plus the per-camera working intrinsics (Kannala: zero distortion)
//   k_pnp_ransac<G> G lanes per observation (32/G observations per warp): the lanes evaluate G hypotheses at a time (draw 4
//                   corners, three-point pose + fourth-point disambiguation, inlier count over all corners of the
//                   observation); the sequential stopping rule of ransacP3P (geometrytools.cpp:243-300) is then replayed
//                   over the G counts in draw order, so the outcome is the one a sequential run over the same draws
//                   would produce. Then the inlier mask of the winner.
//   k_pnp_refine<G> Levenberg-Marquardt refinement of the pose on the inliers (lanes stride over corners, 6x6 normal
//                   equations reduced over the group by shuffles in a fixed order, driver PnpLm).
// fp64 throughout; latency/compute bound (a few hundred bytes per observation are read once).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

// Sub-warp groups: G lanes (a power of two <= 32) work on one observation run, 32/G runs per warp, so that a warp
// instruction is shared by several runs: a run of 48 corners needs 1-2 draws and 6 pose parameters, far too little to
// occupy 32 lanes. Shuffles use the full mask with width G; loops are made warp-uniform with __any_sync.
template <int G> __device__ __forceinline__ double group_sum(double x) {
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off, G);
  return x;
}

constexpr int PNP_THREADS = 128;

// P3P RANSAC: selected hypothesis, its inlier mask and its (unrefined) pose.
template <int G>
__global__ void __launch_bounds__(PNP_THREADS) k_pnp_ransac(PnpArgs a) {
  const int gtid = blockIdx.x * PNP_THREADS + threadIdx.x;
  const int b = gtid / G, gl = threadIdx.x & (G - 1);
  const bool has = b < a.n_bobs;
  const long long o0 = has ? a.bobs_ptr[b] : 0;
  const int n = has ? (int)(a.bobs_ptr[b + 1] - o0) : 0;
  const int cam = has ? a.bobs_cam[b] : 0;
  double in[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) in[k] = a.wintr[9 * cam + k];
  RansacState st;
  ransac_init(st, a.it);
  double bR[9], bt[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) bR[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) bt[k] = 0.0;
  const bool live = has && n >= 4;
#pragma unroll 1
  for (int round = 0;; ++round) {
    const bool go = live && ransac_continue(st, a.it);
    if (!__any_sync(0xffffffffu, go)) break;
    const int h = round * G + gl;
    double R[9], t[3];
    bool ok = go && h < a.it;
    if (ok) {
      int idx[4];
      ransac_sample4(a.seed, (unsigned long long)b, (unsigned long long)h, n, idx);
      double ab4[4][2], X4[4][3];
#pragma unroll
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
    if (ok) {
#pragma unroll 1
      for (int i = 0; i < n; ++i)
        count += (a.ok[o0 + i] && pnp_is_inlier(R, t, in, a.pts + 3 * (size_t)a.pt[o0 + i], a.w[o0 + i], a.thr2)) ? 1 : 0;
    }
    // replay the sequential rule over this round's G counts, in draw order (identical on every lane of the group)
    int new_best_lane = -1;
#pragma unroll 1
    for (int l = 0; l < G; ++l) {
      const int cl = __shfl_sync(0xffffffffu, count, l, G);
      if (!go || !ransac_continue(st, a.it)) continue;
      const int prev = st.best_h;
      ransac_update(st, round * G + l, cl, n, a.p);
      if (st.best_h != prev) new_best_lane = l;
    }
    const int src = new_best_lane >= 0 ? new_best_lane : 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double x = __shfl_sync(0xffffffffu, R[k], src, G); if (new_best_lane >= 0) bR[k] = x; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double x = __shfl_sync(0xffffffffu, t[k], src, G); if (new_best_lane >= 0) bt[k] = x; }
  }
  // inlier mask of the selected hypothesis: lanes of the group stride over the corners
  int cnt = 0;
  for (int i = gl; i < n; i += G) {
    const bool inl = st.best_h >= 0 && a.ok[o0 + i] &&
                     pnp_is_inlier(bR, bt, in, a.pts + 3 * (size_t)a.pt[o0 + i], a.w[o0 + i], a.thr2);
    a.mask[o0 + i] = inl ? 1 : 0;
    cnt += inl ? 1 : 0;
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off, G);
  if (has && gl == 0) {
    double pose[6] = {0, 0, 0, 0, 0, 0};
    if (st.best_h >= 0) { rot_to_rvec(bR, pose); pose[3] = bt[0]; pose[4] = bt[1]; pose[5] = bt[2]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) { a.pose0[6 * (size_t)b + k] = pose[k]; a.pose[6 * (size_t)b + k] = pose[k]; }
    a.n_inl[b] = cnt; a.best_h[b] = st.best_h; a.n_hyp[b] = st.countit;
  }
}

// Pose-only Levenberg-Marquardt on the inliers (geometrytools.cpp:314-324), driver in PnpLm.
template <int G>
__global__ void __launch_bounds__(PNP_THREADS) k_pnp_refine(PnpArgs a) {
  const int gtid = blockIdx.x * PNP_THREADS + threadIdx.x;
  const int b = gtid / G, gl = threadIdx.x & (G - 1);
  const bool has = b < a.n_bobs && a.best_h[b] >= 0 && a.n_inl[b] >= 4;
  const long long o0 = has ? a.bobs_ptr[b] : 0;
  const int n = has ? (int)(a.bobs_ptr[b + 1] - o0) : 0;
  const int cam = has ? a.bobs_cam[b] : 0;
  double in[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) in[k] = a.wintr[9 * cam + k];
  PnpLm lm;
  {
    double p0[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) p0[k] = has ? a.pose[6 * (size_t)b + k] : 0.0;
    lm.init(p0);
    if (!has) lm.fin = true;
  }
#pragma unroll 1
  for (int iter = 0; iter <= a.refine_it; ++iter) {
    if (!__any_sync(0xffffffffu, !lm.fin)) break;
    double cand[6];
    lm.propose(cand);
    // H (upper 21) / g (6) / cost at cand over the inliers of the run, summed over the group in a fixed order
    double prep[PREP];
    prep_block(cand, prep);
    double H2[21], g2[6], cost2 = 0.0;
#pragma unroll
    for (int k = 0; k < 21; ++k) H2[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) g2[k] = 0.0;
    if (!lm.fin) {
#pragma unroll 1
      for (int i = gl; i < n; i += G) {
        if (!a.mask[o0 + i]) continue;
        const float* p3 = a.pts + 3 * (size_t)a.pt[o0 + i];
        const double X0[3] = {(double)p3[0], (double)p3[1], (double)p3[2]};
        const double2 w = a.w[o0 + i];
        double r[2], Ju[6], Jv[6];
        if (!pnp_residual_jac(prep, in, X0, w.x, w.y, r, Ju, Jv)) { cost2 += 1e12; continue; }   // behind the camera: reject the step
        cost2 += r[0] * r[0] + r[1] * r[1];
        int k = 0;
#pragma unroll
        for (int p = 0; p < 6; ++p) {
#pragma unroll
          for (int q = p; q < 6; ++q) H2[k++] += Ju[p] * Ju[q] + Jv[p] * Jv[q];
          g2[p] += Ju[p] * r[0] + Jv[p] * r[1];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 21; ++k) H2[k] = group_sum<G>(H2[k]);
#pragma unroll
    for (int k = 0; k < 6; ++k) g2[k] = group_sum<G>(g2[k]);
    cost2 = group_sum<G>(cost2);
    lm.update(cand, H2, g2, cost2);
  }
  if (has && gl == 0) {
    bool fin = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) fin = fin && is_fin(lm.pose[k]);
    if (fin) {
#pragma unroll
      for (int k = 0; k < 6; ++k) a.pose[6 * (size_t)b + k] = lm.pose[k];
    }
  }
}

template <int G> static void launch_pnp(const PnpArgs& a, cudaStream_t stream, int64_t* launches) {
  const long long threads = (long long)a.n_bobs * G;
  const unsigned grid = (unsigned)((threads + PNP_THREADS - 1) / PNP_THREADS);
  k_pnp_ransac<G><<<grid, PNP_THREADS, 0, stream>>>(a);
  ++*launches;
  if (a.refine) { k_pnp_refine<G><<<grid, PNP_THREADS, 0, stream>>>(a); ++*launches; }
}

static thread_local std::string g_pnp_err;

}  // namespace mcba

using namespace mcba;

#define PCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_pnp_err = std::string(#x) + ": " + cudaGetErrorString(e_); goto fail; } } while (0)

namespace mcba {   // mcba_api.cu: device-resident observations shared between calls (mcba_obs_register)
bool obs_cache_acquire(const float* u, const float* v, const int32_t* pt, int64_t n, int device, float** d_u, float** d_v, int32_t** d_pt);
void obs_cache_release(const float* d_u);
}
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
  bool shared_obs = false; float *c_u = nullptr, *c_v = nullptr; int32_t* c_pt = nullptr;
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
  // corners already resident (mcba_obs_register, mcba.h): shared with the bundle adjustment that follows, no upload
  shared_obs = mcba::obs_cache_acquire(P->u, P->v, P->pt_idx, P->n_obs, device, &c_u, &c_v, &c_pt);
  if (!shared_obs) {
    PCU(cudaMemcpyAsync(d_all + o_u, P->u, no * 4, cudaMemcpyHostToDevice, stream));
    PCU(cudaMemcpyAsync(d_all + o_v, P->v, no * 4, cudaMemcpyHostToDevice, stream));
    PCU(cudaMemcpyAsync(d_all + o_pt, P->pt_idx, no * 4, cudaMemcpyHostToDevice, stream));
  }
  PCU(cudaMemcpyAsync(d_all + o_pts, P->pts_xyz, (size_t)P->n_pts * 12, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_bp, P->bobs_ptr, (nb + 1) * 8, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_bc, P->bobs_cam, nb * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_cm, P->cam_model, nc * 4, cudaMemcpyHostToDevice, stream));
  PCU(cudaMemcpyAsync(d_all + o_in, P->intrinsics, nc * 72, cudaMemcpyHostToDevice, stream));
  a.u = shared_obs ? c_u : (const float*)(d_all + o_u); a.v = shared_obs ? c_v : (const float*)(d_all + o_v);
  a.pt = shared_obs ? c_pt : (const int32_t*)(d_all + o_pt);
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
  res->kernel_launches = 2;
  {
    // lanes per observation run: every lane gets ~10 corners or more (measured on 48-corner boards: G = 4 2.2 ms, 8 2.6 ms,
    // 16 3.9 ms, 32 6.8 ms for 129 k runs)
    const double avg = (double)P->n_obs / (double)P->n_bobs;
    const char* env = getenv("MCBA_PNP_GROUP");
    int G = avg <= 64.0 ? 4 : (avg <= 128.0 ? 8 : (avg <= 256.0 ? 16 : 32));
    if (env && (atoi(env) == 4 || atoi(env) == 8 || atoi(env) == 16 || atoi(env) == 32)) G = atoi(env);
    if (G == 4) launch_pnp<4>(a, stream, &res->kernel_launches);
    else if (G == 8) launch_pnp<8>(a, stream, &res->kernel_launches);
    else if (G == 16) launch_pnp<16>(a, stream, &res->kernel_launches);
    else launch_pnp<32>(a, stream, &res->kernel_launches);
  }
  PCU(cudaGetLastError());
  PCU(cudaEventRecord(e1, stream));
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
  if (shared_obs) mcba::obs_cache_release(c_u);
  res->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return MCBA_OK;
fail:
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (stream) cudaStreamDestroy(stream);
  if (d_all) cudaFree(d_all);
  if (shared_obs) mcba::obs_cache_release(c_u);
  cudaGetLastError();
  return MCBA_ERR_CUDA;
}

}  // extern "C"
