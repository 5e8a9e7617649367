// mcba_obs_stream.cuh — K1 (fused residual + Jacobian + J^T J), streaming variant.
//
// k_observations<S, MODE_FUSED> gives every board observation its own warp passes of 32 corners: a 48-corner board
// takes two passes, the second one half empty, and the per-corner Jacobian phase - latency-bound, ~8 k cycles per pass
// whatever the number of active lanes - is what the kernel's time is made of. Here a warp owns a CONTIGUOUS range of
// board observations and treats their corners as one stream cut into passes of 32 lanes: the tail of one board
// observation shares its pass with the head of the next (two 48-corner observations = 3 full passes instead of 4).
// Within a pass the lanes of the second observation use their own context (two context slots per warp, alternating
// with the observation's parity) and start on an even lane, so that no DMMA k-step (4 tile rows = 2 corners) mixes
// the two; the accumulators are flushed to the first observation's record at the boundary. At most two observations
// share a pass. Same arithmetic per corner and same row order per observation as the per-observation kernel: the
// per-observation records are bit-identical to it; only the lane that holds a corner changes (and with it the
// shuffle tree of the per-observation cost sum).
#pragma once
#include "mcba_kernels.cuh"

namespace mcba {

template <int STAGE>
__global__ void __launch_bounds__(OBS_WARPS * 32, 3) k_observations_stream(ObsArgs a) {
  typedef Cols<STAGE> C;
  typedef StageTraits<STAGE> ST;
  constexpr int TILE_DOUBLES = C::NC * TS;
  constexpr int WARP_DOUBLES = 2 * CTX_SIZE + TILE_DOUBLES;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ctx0 = smem + warp * WARP_DOUBLES;       // two context slots: observation b uses slot b & 1
  double* T = ctx0 + 2 * CTX_SIZE;
  const int g = lane >> 2, q = lane & 3;
  const int total_warps = gridDim.x * OBS_WARPS, gw = blockIdx.x * OBS_WARPS + warp;
  // contiguous share of the board observations
  const int b_lo = (int)(((long long)a.nb * gw) / total_warps), b_hi = (int)(((long long)a.nb * (gw + 1)) / total_warps);
  if (b_lo >= b_hi) return;
  for (int e = lane; e < TILE_DOUBLES; e += 32) T[e] = 0.0;
  double acc[C::NT][2];
#pragma unroll
  for (int t = 0; t < C::NT; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
  double costA = 0.0, costB = 0.0;                  // per-lane cost of the current / the next observation
  int b = b_lo;                                     // current observation and the position inside it
  int4 rec = a.bobs[b];
  int o_beg = rec.x, o_end = a.bobs[b + 1].x, p = 0;
  int built = b_lo - 1;                             // contexts are built for observations <= built
  auto flush = [&](int bb, double cost) {
    double* out = a.brec + (size_t)bb * C::NCOMP;
    int t = 0;
#pragma unroll
    for (int bi = 0; bi < C::NBLK; ++bi)
#pragma unroll
      for (int bj = bi; bj < C::NBLK; ++bj) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ea = 8 * bi + g, eb = 8 * bj + 2 * q + h;
          if (ea <= eb && eb < C::NC) { const int ci = C::cidx(ea, eb); if (ci >= 0) out[ci] = acc[t][h]; }
        }
        acc[t][0] = 0.0; acc[t][1] = 0.0;
        ++t;
      }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
    if (lane == 0) a.cost_bobs[bb] = cost;
  };
  auto build_ctx = [&](int bb, const int4& r) {
    const int cam = r.y;
    const bool act_c = ST::CAM && !(a.cam_is_ref && a.cam_is_ref[cam]);
    const bool act_b = ST::BOARD && !(a.board_is_ref && a.board_is_ref[r.w]);
    const double* pc = ST::CAM ? a.prep_cam + (size_t)PREP * cam : nullptr;
    const double* po = a.prep_pose + (size_t)PREP * r.z;
    const double* pb = ST::BOARD ? a.prep_board + (size_t)PREP * r.w : nullptr;
    const double* in = a.intr + 9 * (size_t)cam;
    double* ctx = ctx0 + (bb & 1) * CTX_SIZE;
    for (int e = lane; e < CTX_SIZE; e += 32)
      if (e < CTX_NO || e >= CTX_IN) ctx[e] = ctx_entry_a(e, pc, act_c, po, pb, act_b, in);
    __syncwarp();
    for (int e = CTX_NO + lane; e < CTX_IN; e += 32) ctx[e] = ctx_entry_b(e, ctx, po, pb, act_b);
    __syncwarp();
  };
  while (b < b_hi) {
    // ---- plan of this pass: segment 1 = observation b from corner p, segment 2 = head of observation b + 1
    const int n1 = min(32, o_end - o_beg - p);
    const int s2 = (n1 + 1) & ~1;
    int n2 = 0; int4 rec2 = rec; int o2_beg = o_end;
    if (n1 < 32 && p + n1 == o_end - o_beg && s2 < 32 && b + 1 < b_hi) {
      rec2 = a.bobs[b + 1];
      o2_beg = rec2.x;
      n2 = min(32 - s2, a.bobs[b + 2].x - o2_beg);
    }
    const bool in1 = lane < n1, in2 = lane >= s2 && lane < s2 + n2;
    float cu = 0.f, cv = 0.f; int cpt = 0;
    if (in1) { const int i = o_beg + p + lane; cu = a.u[i]; cv = a.v[i]; cpt = a.pt[i]; }
    else if (in2) { const int i = o2_beg + lane - s2; cu = a.u[i]; cv = a.v[i]; cpt = a.pt[i]; }
    if (built < b) { build_ctx(b, rec); built = b; }
    if (n2 > 0 && built < b + 1) { build_ctx(b + 1, rec2); built = b + 1; }
    // ---- Jacobian phase
    if (in1 || in2) {
      const int4 r = in1 ? rec : rec2;
      const int bb = in1 ? b : b + 1;
      const double* ctx = ctx0 + (bb & 1) * CTX_SIZE;
      const bool act_c = ST::CAM && !(a.cam_is_ref && a.cam_is_ref[r.y]);
      const bool act_b = ST::BOARD && !(a.board_is_ref && a.board_is_ref[r.w]);
      const double* p3 = a.pts + 3 * (size_t)cpt;
      const double X0[3] = {p3[0], p3[1], p3[2]};
      TileSink sink{T, lane};
      const double cst = obs_jacobian<STAGE>(ctx, act_c, act_b, a.cam_model[r.y], X0, (double)cu, (double)cv, a.huber_a, sink);
      if (in1) costA += cst; else costB += cst;
    } else {
#pragma unroll
      for (int col = 0; col < C::NC; ++col) *reinterpret_cast<double2*>(T + col * TS + 2 * lane) = make_double2(0.0, 0.0);
    }
    __syncwarp();
    // ---- DMMA phase: k-steps [0, k1) belong to observation b, [k1, k2) to observation b + 1
    const int k1 = s2 >> 1, k2 = (s2 + n2 + 1) >> 1;
    const bool done1 = p + n1 == o_end - o_beg;
    const bool done2 = n2 > 0 && n2 == a.bobs[b + 2].x - o2_beg;
    for (int seg = 0; seg < 2; ++seg) {
      const int ka = seg == 0 ? 0 : k1, kb = seg == 0 ? min(k1, (n1 + 1) >> 1) : k2;
      if (kb > ka) {
        double f[C::NBLK], fn[C::NBLK];
#pragma unroll
        for (int c = 0; c < C::NBLK; ++c) f[c] = (8 * c + g < C::NC) ? T[(8 * c + g) * TS + 4 * ka + q] : 0.0;
        for (int ks = ka; ks < kb; ++ks) {
          const int kn = min(ks + 1, kb - 1);
#pragma unroll
          for (int c = 0; c < C::NBLK; ++c) fn[c] = (8 * c + g < C::NC) ? T[(8 * c + g) * TS + 4 * kn + q] : 0.0;
          int t = 0;
#pragma unroll
          for (int bi = 0; bi < C::NBLK; ++bi)
#pragma unroll
            for (int bj = bi; bj < C::NBLK; ++bj) { dmma884(acc[t][0], acc[t][1], f[bi], f[bj]); ++t; }
#pragma unroll
          for (int c = 0; c < C::NBLK; ++c) f[c] = fn[c];
        }
      }
      if (seg == 0 && done1) { flush(b, costA); costA = costB; costB = 0.0; }
      if (seg == 1 && done2) { flush(b + 1, costA); costA = 0.0; }
    }
    __syncwarp();
    // ---- advance
    if (!done1) { p += n1; }
    else if (n2 == 0) { ++b; p = 0; if (b < b_hi) { rec = a.bobs[b]; o_beg = rec.x; o_end = a.bobs[b + 1].x; } }
    else if (done2) { b += 2; p = 0; if (b < b_hi) { rec = a.bobs[b]; o_beg = rec.x; o_end = a.bobs[b + 1].x; } }
    else { ++b; p = n2; rec = rec2; o_beg = o2_beg; o_end = a.bobs[b + 1].x; }
  }
}

}  // namespace mcba
