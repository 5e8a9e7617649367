// mcba_obs_fused.cuh — K1, fused residual + Jacobian + J~^T J~ (the dominant kernel of an LM iteration), second version.
//
// Same arithmetic per corner as k_observations<S, MODE_FUSED> (obs_jacobian, mcba_math.cuh) and the same per-board-
// observation record; what changes is everything around it. Source-level ncu samples of the first version
// (profiles/r02_ncu_k1_source_hotspots.txt) put 24 % of the warp time into building the per-board-observation context
// (three dependent global round trips: record -> prep records -> products, behind a long if-chain per entry), 30 %
// into the DMMA loop (10 tiles for 28 columns: 37 % of the products are padding or the mirrored half of a diagonal
// tile) and 40 % into the per-corner Jacobian. Here:
//
//  * The context is copy-free and one board observation ahead. The prep records of the camera, the e-block pose and
//    the board and the intrinsics are copied as they are into a double-buffered shared-memory slot with cp.async while
//    the previous board observation is being processed (CtxRaw); a block that is not refined points at an identity
//    record. Only Mco = Rc Ro, No_k = Rc dRo_k and Nb_k = Mco dRb_k (63 three-term products, two rounds) are computed.
//    The board-observation record, the camera model and the first pass's pixels / corner coordinates are prefetched
//    the same way, so no global load sits on the critical path of a warp.
//  * The upper triangle of J~^T J~ is covered by 8x8 DMMA products of *two 4-column groups by two 4-column groups*
//    (rows {a,b} x columns {c,d}) instead of aligned 8x8 tiles. With the 28 columns of stage 5 as 7 groups, the 7 lines
//    {i, i+1, i+3} (mod 7) of the Fano plane cover every pair of groups exactly once; product i = rows {i, i+1} x
//    columns {i, i+3} yields the blocks (i,i), (i,i+3), (i+1,i), (i+1,i+3): 7 DMMAs per k-step instead of 10, 448
//    products for the 406 entries of the triangle (no padding column, one mirrored 4x4 half per product). The lower
//    half-warp holds rows/columns of the first group of a product, the upper half-warp those of the second: every lane
//    loads its column of all groups once per k-step (one shared-memory wavefront per group) and selects.
//  * Tail passes only clear the one tile row pair a DMMA k-step can still read, not every idle lane's.
#pragma once
#include "mcba_kernels.cuh"

namespace mcba {

// Identity prep record (R = I, dR/dw_k = 0, t = 0) for blocks that are not refined.
__device__ const double g_prep_identity[PREP + 1] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// PASS corners per warp pass (tile rows = 2 PASS, row pitch 2 PASS + 4: pitch mod 16 == 4 for 32 and 24), DBUF: two
// context slots (copy for the next board observation issued during the first pass) or one (issued after the last
// Jacobian phase, lands during the last DMMA phase).
template <int PASS> struct TileSinkP {
  static constexpr int PITCH = 2 * PASS + 4;
  double* T; int lane;
  __device__ __forceinline__ void put(int col, double ju, double jv) {
    *reinterpret_cast<double2*>(T + col * PITCH + 2 * lane) = make_double2(ju, jv);
  }
};
template <int STAGE, int PASS = OBS_PASS, bool DBUF = true> struct FusedSmem {
  typedef Cols<STAGE> C;
  static constexpr int NG = (C::NC + 3) / 4;
  static constexpr int PITCH = 2 * PASS + 4;
  static constexpr int TILE = 4 * NG * PITCH;                   // pad columns stay zero
  static constexpr int WARP_DOUBLES = (DBUF ? 2 : 1) * CtxRaw::SIZE + TILE;
  static constexpr int NR = RectCover<NG>::NR;
  static constexpr int IDX_BYTES = NR * 32 * 4;                 // per CTA: record index of every accumulator entry (short2 per product and lane)
  static constexpr int BYTES = OBS_WARPS * WARP_DOUBLES * 8 + IDX_BYTES;
};

// ACC: the warp walks a contiguous range of the items in (camera, board) pair order (ObsArgs::srec / saux) and keeps the
// f x f / f x r entries of J~^T J~ in its accumulators across the board observations of one pair: per board observation
// only the e-block part of the record (f x e, e x e, e x r: 153 of 405 doubles at stage 5) goes to HBM, per run ("chunk")
// one partial pair record. k_pair_reduce - a second pass over 62 % of every record - disappears; k_pair_final sums the
// chunks of a pair in order. The sums are deterministic (fixed ranges, fixed order) but ordered differently from the
// per-record path: pair records agree with it to rounding.
template <int STAGE, int PASS = OBS_PASS, int MINB = 3, bool DBUF = true, bool ACC = false>
__global__ void __launch_bounds__(OBS_WARPS * 32, MINB) k_observations_fused(ObsArgs a) {
  typedef Cols<STAGE> C;
  typedef StageTraits<STAGE> ST;
  typedef FusedSmem<STAGE, PASS, DBUF> SM;
  constexpr int TSP = SM::PITCH;
  typedef RectCover<SM::NG> RCV;
  constexpr int NG = SM::NG, NR = RCV::NR;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* cbuf = smem + warp * SM::WARP_DOUBLES;
  double* T = cbuf + (DBUF ? 2 : 1) * CtxRaw::SIZE;
  const int total_warps = gridDim.x * OBS_WARPS;
  const int g = lane >> 2, q = lane & 3, hi = lane >> 4, c4 = g & 3;
  // Where each accumulator entry goes in the compact record (-1: not stored): lane constants, kept in shared memory -
  // left to the compiler they are hoisted out of the board-observation loop and spilled (14 dependent local-memory
  // round trips per board observation in the epilogue, 30 % of the kernel: profiles/r02_ncu_k1_source_hotspots.txt)
  short2* sidx = reinterpret_cast<short2*>(smem + OBS_WARPS * SM::WARP_DOUBLES);
  for (int r = warp; r < NR; r += OBS_WARPS)
    sidx[r * 32 + lane] = make_short2((short)k1_store_index<STAGE>(r, lane, 0), (short)k1_store_index<STAGE>(r, lane, 1));
  __syncthreads();
  // work items: ACC: positions [it, it_end) of the pair-ordered list; otherwise board observations it, it + total_warps, ...
  int it, it_end, it_step;
  if (ACC) {
    const long long s = blockIdx.x * OBS_WARPS + warp;
    it = (int)((long long)a.nb * s / a.n_streams); it_end = (int)((long long)a.nb * (s + 1) / a.n_streams); it_step = 1;
  } else { it = blockIdx.x * OBS_WARPS + warp; it_end = a.nb; it_step = total_warps; }
  if (it >= it_end) return;
  auto fetch = [&](int i, int4& r, int& end, int& f, int& ob, int& ch) {
    if (ACC) { r = a.srec[i]; const int4 x = a.saux[i]; end = x.x; f = x.y; ob = x.z; ch = x.w; }
    else { r = a.bobs[i]; end = a.bobs[i + 1].x; f = a.bobs_flags[i]; ob = i; ch = 0; }
  };
  for (int e = lane; e < SM::TILE; e += 32) T[e] = 0.0;
  // lane-constant operands of the two rounds of context products: out = M[3 i + .] . D[. + j]
  int m1, d1, o1, m2, d2, o2;   // products 0..31 in the first round (Mco and No[0..22]), 32..62 in the second (No[23..26], Nb)
  ctx_raw_operands(lane, m1, d1, o1);
  ctx_raw_operands(min(32 + lane, 62), m2, d2, o2);
  // flags of a board observation (ObsArgs::bobs_flags: bit 0 reference camera, bit 1 reference board, camera model
  // << 2): one word per board observation, so that nothing has to be looked up through the record
  auto issue_ctx = [&](const int4& r, int f, double* dst) {   // raw context of one board observation -> dst (asynchronous)
    const bool act_c = ST::CAM && !(f & 1), act_b = ST::BOARD && !(f & 2);
    const double* pc = act_c ? a.prep_cam + (size_t)PREP * r.y : g_prep_identity;
    const double* po = a.prep_pose + (size_t)PREP * r.z;
    const double* pb = act_b ? a.prep_board + (size_t)PREP * r.w : g_prep_identity;
    const double* in = a.intr + 9 * (size_t)r.y;
#pragma unroll
    for (int e = lane; e < CtxRaw::RAW; e += 32) {
      const double* src = e < CtxRaw::PO ? pc + e : e < CtxRaw::PB ? po + (e - CtxRaw::PO) : e < CtxRaw::IN ? pb + (e - CtxRaw::PB) : in + (e - CtxRaw::IN);
      cp_async8(dst + e, src);
    }
  };
  // prologue: this warp's first board observation
  int obs_begin, obs_end, fl, ob, ch;
  {
    int4 rec;
    fetch(it, rec, obs_end, fl, ob, ch);
    obs_begin = rec.x;
    issue_ctx(rec, fl, cbuf);
  }
  double acc[NR][2];
#pragma unroll
  for (int t = 0; t < NR; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
  float pf_u = 0.f, pf_v = 0.f; int pf_pt = 0;
  if (lane < PASS && obs_begin + lane < obs_end) { pf_u = a.u[obs_begin + lane]; pf_v = a.v[obs_begin + lane]; pf_pt = a.pt[obs_begin + lane]; }
  int par = 0;
  while (true) {
    double* ctx = cbuf + (DBUF ? par : 0) * CtxRaw::SIZE;
    const int itn = it + it_step;
    const bool has_n = itn < it_end;
    int4 rec_n = make_int4(0, 0, 0, 0); int end_n = 0, fl_n = 0, ob_n = 0, ch_n = -1;   // the next item: consumed after the first Jacobian phase at the earliest
    if (has_n) fetch(itn, rec_n, end_n, fl_n, ob_n, ch_n);
    cp_async_wait_all();
    __syncwarp();
    ctx[o1] = ctx[m1] * ctx[d1] + ctx[m1 + 1] * ctx[d1 + 3] + ctx[m1 + 2] * ctx[d1 + 6];
    __syncwarp();
    if (lane < 31) ctx[o2] = ctx[m2] * ctx[d2] + ctx[m2 + 1] * ctx[d2 + 3] + ctx[m2 + 2] * ctx[d2 + 6];
    __syncwarp();
    const bool act_c = ST::CAM && !(fl & 1), act_b = ST::BOARD && !(fl & 2);
    const int model = fl >> 2;
    if (!ACC) {
#pragma unroll
      for (int t = 0; t < NR; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    }
    double cost = 0.0;
    for (int base = obs_begin; base < obs_end; base += PASS) {
      const int nact = min(PASS, obs_end - base);
      const bool active = lane < nact;
      const float cur_u = pf_u, cur_v = pf_v;
      const double* p3 = a.pts + 3 * (size_t)pf_pt;       // the point table is small: L1 hits
      const double X0[3] = {p3[0], p3[1], p3[2]};
      // pixels / corner index of the following pass: of this board observation, or the first one of the next
      const bool last_pass = base + PASS >= obs_end;
      const int nxt = last_pass ? rec_n.x + lane : base + PASS + lane;
      const int nxt_end = last_pass ? end_n : obs_end;
      const bool nxt_ok = lane < PASS && (has_n || !last_pass) && nxt < nxt_end;
      if (nxt_ok) { pf_u = a.u[nxt]; pf_v = a.v[nxt]; pf_pt = a.pt[nxt]; }
      if (active) {
        TileSinkP<PASS> sink{T, lane};
        cost += obs_jacobian<STAGE, TileSinkP<PASS>, CtxRaw>(ctx, act_c, act_b, model, X0, (double)cur_u, (double)cur_v, a.huber_a, sink);
      } else if (lane == nact && (nact & 1)) {           // the odd corner's partner rows of the last k-step
#pragma unroll
        for (int col = 0; col < C::NC; ++col)
          *reinterpret_cast<double2*>(T + col * TSP + 2 * lane) = make_double2(0.0, 0.0);
      }
      if (DBUF && base == obs_begin && has_n) issue_ctx(rec_n, fl_n, cbuf + (par ^ 1) * CtxRaw::SIZE);
      __syncwarp();
      if (!DBUF && last_pass && has_n) issue_ctx(rec_n, fl_n, cbuf);   // nobody reads the context any more in this iteration
      const int ksteps = (2 * nact + 3) >> 2;
      const double* Tl = T + c4 * TSP + q;
      double G[NG], Gn[NG];
#pragma unroll
      for (int j = 0; j < NG; ++j) G[j] = Tl[4 * j * TSP];
      for (int ks = 0; ks < ksteps; ++ks) {
        const int kn = min(ks + 1, ksteps - 1);
#pragma unroll
        for (int j = 0; j < NG; ++j) Gn[j] = Tl[4 * j * TSP + 4 * kn];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double fa = hi ? G[RCV::r1(r)] : G[RCV::r0(r)];
          const double fb = hi ? G[RCV::c1(r)] : G[RCV::c0(r)];
          dmma884(acc[r][0], acc[r][1], fa, fb);
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = Gn[j];
      }
      __syncwarp();
    }
    if (obs_begin >= obs_end && has_n) {                  // empty board observation: no pass prefetched for the next one
      issue_ctx(rec_n, fl_n, cbuf + (DBUF ? (par ^ 1) : 0) * CtxRaw::SIZE);
      if (lane < PASS && rec_n.x + lane < end_n) { pf_u = a.u[rec_n.x + lane]; pf_v = a.v[rec_n.x + lane]; pf_pt = a.pt[rec_n.x + lane]; }
    }
    {
      double* out = a.brec + (size_t)ob * C::NCOMP;
      if (!ACC) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const short2 ci = sidx[r * 32 + lane];
          if (ci.x >= 0) out[ci.x] = acc[r][0];
          if (ci.y >= 0) out[ci.y] = acc[r][1];
        }
      } else {
        // e-block part: stored and cleared per board observation; f part: kept until the run of this pair ends
        const bool run_ends = !has_n || ch_n != ch;
        double* part = a.partial + (size_t)ch * (size_t)C::CI_FE;
        // (one predicated store and one select per entry: the class of an entry varies from lane to lane, branches
        // here serialise the warp - 15 % of the kernel in the first build)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const short2 ci = sidx[r * 32 + lane];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = h ? ci.y : ci.x;
            const bool is_e = c >= C::CI_FE;
            double* dst = (is_e ? out : part) + max(c, 0);
            if (is_e || (run_ends && c >= 0)) *dst = acc[r][h];
            acc[r][h] = (is_e || run_ends) ? 0.0 : acc[r][h];
          }
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
      if (lane == 0) a.cost_bobs[ob] = cost;
    }
    if (!has_n) break;
    it = itn; obs_begin = rec_n.x; obs_end = end_n; fl = fl_n; ob = ob_n; ch = ch_n; par ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Third version: TWO board observations per warp, one per half-warp, 16 corners of each per pass. A 48-corner board
// takes 3 full passes per pair instead of 2 x (one full + one half-empty) = 4 - the per-corner Jacobian phase costs the
// same whatever the number of active lanes - and a 16-corner board one pass per pair instead of two. Every lane keeps
// the scalars of "its" board observation (record, range, flags, prefetched pixels), so nothing is held twice; the two
// contexts sit 16 banks apart; tile rows 0..31 are the first observation's, 32..63 the second's; the DMMA loop runs
// the two observations' k-steps side by side (two independent sets of accumulators: twice the products in flight per
// shared-memory round trip). The context is single-buffered: the asynchronous copy for the next pair is issued after
// the last Jacobian phase of the current one and lands during its last DMMA phase. Same per-observation arithmetic and
// row order: the records are bit-identical to the other versions'.
constexpr int HP = 16;                          // corners per half-warp pass
constexpr int CTXP = CtxRaw::SIZE + 8;          // context pitch: 400 words = 16 banks apart

template <int STAGE> struct PairSmem {
  typedef Cols<STAGE> C;
  static constexpr int NG = (C::NC + 3) / 4;
  static constexpr int TILE = 4 * NG * TS;
  static constexpr int WARP_DOUBLES = 2 * CTXP + TILE;
  static constexpr int NR = RectCover<NG>::NR;
  static constexpr int IDX_BYTES = NR * 32 * 4;
  static constexpr int CTAB_BYTES = 64 * 4;     // operand offsets of the 63 context products, 4 rounds of 16
  static constexpr int BYTES = OBS_WARPS * WARP_DOUBLES * 8 + IDX_BYTES + CTAB_BYTES;
};

// ACC (see k_observations_fused): every half-warp is a stream with its own contiguous range of the pair-ordered items
// and its own accumulator set; otherwise half h of warp w takes board observations 2 (w + k total_warps) + h.
template <int STAGE, bool ACC = false>
__global__ void __launch_bounds__(OBS_WARPS * 32, 3) k_observations_pair(ObsArgs a) {
  typedef Cols<STAGE> C;
  typedef StageTraits<STAGE> ST;
  typedef PairSmem<STAGE> SM;
  typedef RectCover<SM::NG> RCV;
  constexpr int NG = SM::NG, NR = RCV::NR;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3, hi = lane >> 4, c4 = g & 3, l16 = lane & 15;
  double* cbuf = smem + warp * SM::WARP_DOUBLES;
  double* ctx = cbuf + hi * CTXP;                // this half-warp's context
  double* T = cbuf + 2 * CTXP;
  short2* sidx = reinterpret_cast<short2*>(smem + OBS_WARPS * SM::WARP_DOUBLES);
  int* ctab = reinterpret_cast<int*>(sidx + NR * 32);
  for (int r = warp; r < NR; r += OBS_WARPS)
    sidx[r * 32 + lane] = make_short2((short)k1_store_index<STAGE>(r, lane, 0), (short)k1_store_index<STAGE>(r, lane, 1));
  if (threadIdx.x < 64) {   // context products (ctx_raw_operands): Mco (9), No (27), Nb (27, needs Mco: rounds 2, 3)
    int m, d, o;
    ctx_raw_operands(min((int)threadIdx.x, 62), m, d, o);
    ctab[threadIdx.x] = m | (d << 8) | (o << 16);
  }
  __syncthreads();
  const int total_warps = gridDim.x * OBS_WARPS;
  // this half-warp's items
  int it, it_end, it_step;
  if (ACC) {
    const long long s = 2 * (blockIdx.x * OBS_WARPS + warp) + hi;
    it = (int)((long long)a.nb * s / a.n_streams); it_end = (int)((long long)a.nb * (s + 1) / a.n_streams); it_step = 1;
  } else { it = 2 * (blockIdx.x * OBS_WARPS + warp) + hi; it_end = a.nb; it_step = 2 * total_warps; }
  bool cur_ok = it < it_end;
  if (!__any_sync(0xffffffffu, cur_ok)) return;
  for (int e = lane; e < SM::TILE; e += 32) T[e] = 0.0;
  auto fetch = [&](int i, bool ok, int4& r, int& end, int& f, int& ob, int& ch) {   // an exhausted half sees empty items
    if (!ok) { r = make_int4(0, 0, 0, 0); end = 0; f = 0; ob = 0; ch = -1; }
    else if (ACC) { r = a.srec[i]; const int4 x = a.saux[i]; end = x.x; f = x.y; ob = x.z; ch = x.w; }
    else { r = a.bobs[i]; end = a.bobs[i + 1].x; f = a.bobs_flags[i]; ob = i; ch = 0; }
  };
  auto issue_ctx = [&](const int4& r, int f) {
    const bool act_c = ST::CAM && !(f & 1), act_b = ST::BOARD && !(f & 2);
    const double* pc = act_c ? a.prep_cam + (size_t)PREP * r.y : g_prep_identity;
    const double* po = a.prep_pose + (size_t)PREP * r.z;
    const double* pb = act_b ? a.prep_board + (size_t)PREP * r.w : g_prep_identity;
    const double* in = a.intr + 9 * (size_t)r.y;
#pragma unroll
    for (int e = l16; e < CtxRaw::RAW; e += 16) {
      const double* src = e < CtxRaw::PO ? pc + e : e < CtxRaw::PB ? po + (e - CtxRaw::PO) : e < CtxRaw::IN ? pb + (e - CtxRaw::PB) : in + (e - CtxRaw::IN);
      cp_async8(ctx + e, src);
    }
  };
  // batched problems: board observations of cameras whose gate is off are skipped (nothing computed, nothing stored)
  int my_begin, obs_end, fl, ob, ch, gate_on = 1;
  {
    int4 rec;
    fetch(it, cur_ok, rec, obs_end, fl, ob, ch);
    my_begin = rec.x;
    issue_ctx(rec, fl);
    if (a.gate && cur_ok) gate_on = a.gate[rec.y];
  }
  float pf_u = 0.f, pf_v = 0.f; int pf_pt = 0;
  if (my_begin + l16 < obs_end) { pf_u = a.u[my_begin + l16]; pf_v = a.v[my_begin + l16]; pf_pt = a.pt[my_begin + l16]; }
  double accA[NR][2], accB[NR][2];
#pragma unroll
  for (int t = 0; t < NR; ++t) { accA[t][0] = 0.0; accA[t][1] = 0.0; accB[t][0] = 0.0; accB[t][1] = 0.0; }
  while (true) {
    const int itn = it + it_step;
    const bool has_n = cur_ok && itn < it_end;
    int4 rec_n; int end_n, fl_n, ob_n, ch_n;                   // consumed in the last pass
    fetch(itn, has_n, rec_n, end_n, fl_n, ob_n, ch_n);
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int rd = 0; rd < 4; ++rd) {
      const int pk = ctab[rd * 16 + l16];
      const int m = pk & 255, d = (pk >> 8) & 255, o = pk >> 16;
      ctx[o] = ctx[m] * ctx[d] + ctx[m + 1] * ctx[d + 3] + ctx[m + 2] * ctx[d + 6];
      __syncwarp();
    }
    const bool act_c = ST::CAM && !(fl & 1), act_b = ST::BOARD && !(fl & 2);
    const int model = fl >> 2;
    if (!ACC) {
#pragma unroll
      for (int t = 0; t < NR; ++t) { accA[t][0] = 0.0; accA[t][1] = 0.0; accB[t][0] = 0.0; accB[t][1] = 0.0; }
    }
    double cost = 0.0;
    const int n_mine = gate_on ? obs_end - my_begin : 0;
    int gate_n = 1;
    const int nA = __shfl_sync(0xffffffffu, n_mine, 0), nB = __shfl_sync(0xffffffffu, n_mine, 16);
    const int npass = (max(nA, nB) + HP - 1) / HP;
    for (int pass = 0; pass < npass; ++pass) {
      const int off = pass * HP;
      const int nact = min(HP, max(0, n_mine - off));
      const bool active = l16 < nact;
      const float cur_u = pf_u, cur_v = pf_v;
      const double* p3 = a.pts + 3 * (size_t)pf_pt;       // the point table is small: L1 hits
      const double X0[3] = {p3[0], p3[1], p3[2]};
      const bool last = pass + 1 == npass;
      const int nxt = last ? rec_n.x + l16 : my_begin + off + HP + l16;
      const int nxt_end = last ? end_n : obs_end;
      const bool nxt_ok = (has_n || !last) && nxt < nxt_end;
      if (nxt_ok) { pf_u = a.u[nxt]; pf_v = a.v[nxt]; pf_pt = a.pt[nxt]; }
      if (active) {
        TileSink sink{T, lane};
        cost += obs_jacobian<STAGE, TileSink, CtxRaw>(ctx, act_c, act_b, model, X0, (double)cur_u, (double)cur_v, a.huber_a, sink);
      } else if (l16 == nact && (nact & 1)) {
#pragma unroll
        for (int col = 0; col < C::NC; ++col)
          *reinterpret_cast<double2*>(T + col * TS + 2 * lane) = make_double2(0.0, 0.0);
      }
      __syncwarp();
      if (last && has_n && a.gate) gate_n = a.gate[rec_n.y];   // consumed at the top of the next iteration
      if (last && has_n) issue_ctx(rec_n, fl_n);          // the context is not read any more in this iteration
      const int kA = (2 * min(HP, max(0, nA - off)) + 3) >> 2, kB = (2 * min(HP, max(0, nB - off)) + 3) >> 2;
      // the two observations' k-steps one after the other (side by side, with both sets of fragments live, the loop
      // runs out of registers at stage 5 and the DMMAs serialise on two temporaries: 41 % of the kernel's samples)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int kh = half ? kB : kA;
        const double* Tl = T + c4 * TS + q + half * 2 * HP;
        double G[NG], Gn[NG];
#pragma unroll
        for (int j = 0; j < NG; ++j) G[j] = Tl[4 * j * TS];
        for (int ks = 0; ks < kh; ++ks) {
          const int kn = min(ks + 1, kh - 1);
#pragma unroll
          for (int j = 0; j < NG; ++j) Gn[j] = Tl[4 * j * TS + 4 * kn];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const double fa = hi ? G[RCV::r1(r)] : G[RCV::r0(r)];
            const double fb = hi ? G[RCV::c1(r)] : G[RCV::c0(r)];
            if (half) dmma884(accB[r][0], accB[r][1], fa, fb); else dmma884(accA[r][0], accA[r][1], fa, fb);
          }
#pragma unroll
          for (int j = 0; j < NG; ++j) G[j] = Gn[j];
        }
      }
      __syncwarp();
    }
    if (npass == 0 && has_n) {                             // both board observations empty: nothing prefetched for the next pair
      issue_ctx(rec_n, fl_n);
      if (a.gate) gate_n = a.gate[rec_n.y];
      if (rec_n.x + l16 < end_n) { pf_u = a.u[rec_n.x + l16]; pf_v = a.v[rec_n.x + l16]; pf_pt = a.pt[rec_n.x + l16]; }
    }
    {
      const int obA = __shfl_sync(0xffffffffu, ob, 0), obB = __shfl_sync(0xffffffffu, ob, 16);
      const bool ok = cur_ok && gate_on;
      const bool okA = __shfl_sync(0xffffffffu, (int)ok, 0) != 0, okB = __shfl_sync(0xffffffffu, (int)ok, 16) != 0;
      double* outA = a.brec + (size_t)obA * C::NCOMP;
      double* outB = a.brec + (size_t)obB * C::NCOMP;
      if (!ACC) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const short2 ci = sidx[r * 32 + lane];
          if (ci.x >= 0) { if (okA) outA[ci.x] = accA[r][0]; if (okB) outB[ci.x] = accB[r][0]; }
          if (ci.y >= 0) { if (okA) outA[ci.y] = accA[r][1]; if (okB) outB[ci.y] = accB[r][1]; }
        }
      } else {
        const bool ends = cur_ok && (!has_n || ch_n != ch);
        const bool endA = __shfl_sync(0xffffffffu, (int)ends, 0) != 0, endB = __shfl_sync(0xffffffffu, (int)ends, 16) != 0;
        const int chA = __shfl_sync(0xffffffffu, ch, 0), chB = __shfl_sync(0xffffffffu, ch, 16);
        double* partA = a.partial + (size_t)max(chA, 0) * (size_t)C::CI_FE;
        double* partB = a.partial + (size_t)max(chB, 0) * (size_t)C::CI_FE;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const short2 ci = sidx[r * 32 + lane];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = h ? ci.y : ci.x;
            const bool is_e = c >= C::CI_FE;
            double* dA = (is_e ? outA : partA) + max(c, 0);
            double* dB = (is_e ? outB : partB) + max(c, 0);
            if (is_e ? okA : (endA && c >= 0)) *dA = accA[r][h];
            if (is_e ? okB : (endB && c >= 0)) *dB = accB[r][h];
            accA[r][h] = (is_e || endA) ? 0.0 : accA[r][h];
            accB[r][h] = (is_e || endB) ? 0.0 : accB[r][h];
          }
        }
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
      if (l16 == 0 && ok) a.cost_bobs[ob] = cost;
    }
    if (!__any_sync(0xffffffffu, has_n)) break;
    it = itn; cur_ok = has_n; my_begin = rec_n.x; obs_end = end_n; fl = fl_n; ob = ob_n; ch = ch_n; gate_on = gate_n;
  }
}

}  // namespace mcba
