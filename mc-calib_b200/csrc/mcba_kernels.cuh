// mcba_kernels.cuh — sm_100a kernels of the bundle-adjustment hot path (fp64 throughout).
//
//   k_prep_blocks        K0  R, dR/dw_k, t per parameter block (replaces the per-residual Rodrigues setup
//                            inside ceres::AngleAxisRotatePoint)
//   k_observations<MODE> K1  one warp per board observation. MODE_FUSED: residual + Jacobian + J^T J / J^T r
//                            accumulation with fp64 tensor-core MMA (DMMA m8n8k4) without materialising J;
//                            MODE_MATERIALIZE: writes residuals + Jacobian planes to HBM (the HBM-roofline
//                            kernel of BASELINE.json); MODE_FROM_J: reads those planes back and accumulates;
//                            MODE_COST: residual-only pass for the candidate point.
//   k_eblock_assemble    K2a per e-block: E^T E, E^T r and the dense row block [E^T F | E^T r]
//   k_pair_reduce/final  K2b per (camera, board) pair: F^T F blocks and F^T r, fixed-order (deterministic)
//   k_eblock_factor      K2c (E^T E + D^2) = L L^T, Z = L^-1 [E^T F | E^T r] (Jacobi-scaled)
//   k_syrk / k_syrk_sum  K2d Schur complement sum_o Z_o^T Z_o with DMMA, split-K with fixed-order sum
//   k_build_reduced      K2e S = F^T F + D^2 - Z^T Z, rhs
//   k_chol_coop          K3  cooperative persistent blocked Cholesky + both triangular solves of the reduced system
//   k_backsubst, k_fstep K4  e-block back-substitution, candidate point, model-cost-change pieces
// Replaces ProgramEvaluator / SchurEliminator / CHOLMOD of ceres-solver 2.2.0 as used by the reference
// (SURVEY.md §2.2, §8a rows a12-a15).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math_constants.h>
#include "mcba_math.cuh"

namespace mcba {

constexpr int OBS_PASS = 32;      // corners per warp pass (64 tile rows = 16 DMMA k-steps); 24 (4 CTAs/SM) measured no faster
constexpr int TS = 2 * OBS_PASS + 4;   // row stride (doubles) of the per-warp J^T tile: 48 rows + 4 (bank spread: TS mod 16 == 4)
constexpr int OBS_WARPS = 4;      // warps per CTA in the observation kernels
enum { MODE_FUSED = 0, MODE_MATERIALIZE = 1, MODE_FROM_J = 2, MODE_COST = 3 };

struct ObsArgs {
  const float* u; const float* v; const int32_t* pt; const double* pts;   // observations (SoA) + point table
  const int4* bobs;                 // [nb+1] {obs_begin, cam, pose, board}
  const double* prep_cam; const double* prep_pose; const double* prep_board; const double* intr;
  const int32_t* cam_model; const uint8_t* cam_is_ref; const uint8_t* board_is_ref;
  double huber_a;
  int nb;
  double* brec;                     // [nb][NCOMP] per-bobs J~^T J~, compact (Cols<STAGE>::cidx)
  double* cost_bobs;                // [nb]
  double* jmat; double* resmat;     // materialised Jacobian blocks (see BlockSink); resmat unused
  int64_t npad;
  const int* bobs_flags;            // [nb] bit 0: reference camera, bit 1: reference board, camera model << 2 (the fused kernels of round 2)
  // accumulating variants of the fused kernels: items in (camera, board) pair order - srec {first obs, cam, pose, board},
  // saux {end obs, flags, board-observation index, chunk} - n_streams contiguous ranges, one partial pair record per chunk
  const int4* srec; const int4* saux; int n_streams; double* partial;
  const uint8_t* gate;              // batched problems: per-camera flag, board observations of gated-off cameras are skipped
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_prep_blocks(const double* __restrict__ cam_pose, int n_cam, const double* __restrict__ pose,
                              int n_pose, const double* __restrict__ board_pose, int n_board,
                              double* __restrict__ prep_cam, double* __restrict__ prep_pose,
                              double* __restrict__ prep_board) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* src; double* dst;
  if (i < n_pose) { src = pose + 6 * (size_t)i; dst = prep_pose + PREP * (size_t)i; }
  else if (i < n_pose + n_cam) { src = cam_pose + 6 * (size_t)(i - n_pose); dst = prep_cam + PREP * (size_t)(i - n_pose); }
  else if (i < n_pose + n_cam + n_board) { const int b = i - n_pose - n_cam; src = board_pose + 6 * (size_t)b; dst = prep_board + PREP * (size_t)b; }
  else return;
  double p[6], out[PREP];
  for (int k = 0; k < 6; ++k) p[k] = src[k];
  prep_block(p, out);
  for (int k = 0; k < PREP; ++k) dst[k] = out[k];
}

// Sinks for obs_jacobian
struct TileSink {            // J^T tile in shared memory: T[col][2*lane + {0,1}]
  double* T; int lane;
  __device__ __forceinline__ void put(int col, double ju, double jv) {
    *reinterpret_cast<double2*>(T + col * TS + 2 * lane) = make_double2(ju, jv);
  }
};
// Materialised Jacobian layout (MODE_MATERIALIZE writes, MODE_FROM_J reads): per board observation one contiguous
// block of n_obs(bobs) x NC x 2 doubles starting at jmat + 2*NC*obs_begin, column-major inside the block:
// element (local column c, corner j) = double2{u-row, v-row} at ((c * n_obs(bobs) + j) * 2). Column R0 holds the
// corrected residual. A warp pass writes 512 contiguous bytes per column, all inside one ~21 KB region.
struct BlockSink {
  double* blk; int nobs, j;
  __device__ __forceinline__ void put(int col, double ju, double jv) {
    *reinterpret_cast<double2*>(blk + ((size_t)col * nobs + j) * 2) = make_double2(ju, jv);
  }
};

template <int STAGE, int MODE>
__global__ void __launch_bounds__(OBS_WARPS * 32, (MODE == MODE_FUSED || MODE == MODE_FROM_J) ? 3 : (MODE == MODE_COST ? 8 : 4)) k_observations(ObsArgs a) {
  typedef Cols<STAGE> C;
  typedef StageTraits<STAGE> ST;
  constexpr bool ACC = (MODE == MODE_FUSED || MODE == MODE_FROM_J);
  constexpr int TILE_DOUBLES = ACC ? C::NC * TS : 0;   // only the NC real columns are stored
  constexpr int PASS = ACC ? OBS_PASS : 32;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ctx = smem + warp * (CTX_SIZE + TILE_DOUBLES);
  double* T = ctx + CTX_SIZE;
  if (ACC) {
    for (int e = lane; e < TILE_DOUBLES; e += 32) T[e] = 0.0;
  }
  const int g = lane >> 2, q = lane & 3;
  const int total_warps = gridDim.x * OBS_WARPS;
  for (int b = blockIdx.x * OBS_WARPS + warp; b < a.nb; b += total_warps) {
    const int4 rec = a.bobs[b];
    const int obs_begin = rec.x, obs_end = a.bobs[b + 1].x;
    const int cam = rec.y;
    if (a.gate && !a.gate[cam]) continue;
    const bool act_c = ST::CAM && !(a.cam_is_ref && a.cam_is_ref[cam]);
    const bool act_b = ST::BOARD && !(a.board_is_ref && a.board_is_ref[rec.w]);
    const int model = a.cam_model[cam];
    const double* pc = ST::CAM ? a.prep_cam + (size_t)PREP * cam : nullptr;
    const double* po = a.prep_pose + (size_t)PREP * rec.z;
    const double* pb = ST::BOARD ? a.prep_board + (size_t)PREP * rec.w : nullptr;
    const double* in = a.intr + 9 * (size_t)cam;
    // observation loads run one pass ahead: the first pass's are issued before the context is built
    float pf_u = 0.f, pf_v = 0.f; int pf_pt = 0;
    if (MODE != MODE_FROM_J && lane < PASS && obs_begin + lane < obs_end) {
      pf_u = a.u[obs_begin + lane]; pf_v = a.v[obs_begin + lane]; pf_pt = a.pt[obs_begin + lane];
    }
    __syncwarp();
    if (MODE != MODE_FROM_J) {
      if (MODE == MODE_COST) {   // residual only: the three (R, t) pairs and the intrinsics
        for (int e = lane; e < CTX_DRC + 9; e += 32) {
          const int ee = (e < CTX_DRC) ? e : CTX_IN + (e - CTX_DRC);
          ctx[ee] = ctx_entry_a(ee, pc, act_c, po, pb, act_b, in);
        }
      } else {
        for (int e = lane; e < CTX_SIZE; e += 32)
          if (e < CTX_NO || e >= CTX_IN) ctx[e] = ctx_entry_a(e, pc, act_c, po, pb, act_b, in);
      }
      __syncwarp();
      if (MODE != MODE_COST) {
        for (int e = CTX_NO + lane; e < CTX_IN; e += 32) ctx[e] = ctx_entry_b(e, ctx, po, pb, act_b);
        __syncwarp();
      }
    }
    double acc[ACC ? C::NT : 1][2];
#pragma unroll
    for (int t = 0; t < (ACC ? C::NT : 1); ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    double cost = 0.0;
    for (int base = obs_begin; base < obs_end; base += PASS) {
      const int i = base + lane;
      const bool active = lane < PASS && i < obs_end;
      const float cur_u = pf_u, cur_v = pf_v; const int cur_pt = pf_pt;
      if (MODE != MODE_FROM_J && lane < PASS && i + PASS < obs_end) {
        pf_u = a.u[i + PASS]; pf_v = a.v[i + PASS]; pf_pt = a.pt[i + PASS];
      }
      if (active) {
        if (MODE == MODE_FROM_J) {
          const double* blk = a.jmat + (size_t)2 * C::NC * obs_begin;
          const int nobs = obs_end - obs_begin, j = i - obs_begin;
#pragma unroll
          for (int col = 0; col < C::NC; ++col)
            *reinterpret_cast<double2*>(T + col * TS + 2 * lane) =
                *reinterpret_cast<const double2*>(blk + ((size_t)col * nobs + j) * 2);
        } else {
          const double* p3 = a.pts + 3 * (size_t)cur_pt;
          const double X0[3] = {p3[0], p3[1], p3[2]};
          const double uu = (double)cur_u, vv = (double)cur_v;
          if (MODE == MODE_COST) {
            double ru, rv;
            cost += obs_cost(ctx, model, X0, uu, vv, a.huber_a, ru, rv);
          } else if (MODE == MODE_FUSED) {
            TileSink sink{T, lane};
            cost += obs_jacobian<STAGE>(ctx, act_c, act_b, model, X0, uu, vv, a.huber_a, sink);
          } else {
            BlockSink sink{a.jmat + (size_t)2 * C::NC * obs_begin, obs_end - obs_begin, i - obs_begin};
            cost += obs_jacobian<STAGE>(ctx, act_c, act_b, model, X0, uu, vv, a.huber_a, sink);
          }
        }
      } else if (ACC && lane < PASS) {
#pragma unroll
        for (int col = 0; col < C::NC; ++col)
          *reinterpret_cast<double2*>(T + col * TS + 2 * lane) = make_double2(0.0, 0.0);
      }
      if (ACC) {
        __syncwarp();
        const int nrows = 2 * min(PASS, obs_end - base);
        const int ksteps = (nrows + 3) >> 2;
        double f[C::NBLK], fn[C::NBLK];   // fragments of the next k-step are fetched while this one's DMMAs issue
#pragma unroll
        for (int c = 0; c < C::NBLK; ++c) f[c] = (8 * c + g < C::NC) ? T[(8 * c + g) * TS + q] : 0.0;
        for (int ks = 0; ks < ksteps; ++ks) {
          const int kn = min(ks + 1, ksteps - 1);
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
        __syncwarp();
      }
    }
    if (ACC) {
      double* out = a.brec + (size_t)b * C::NCOMP;
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
          ++t;
        }
    }
    if (MODE != MODE_FROM_J) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, off);
      if (lane == 0) a.cost_bobs[b] = cost;
    }
  }
}

// Unrobustified per-observation reprojection error sqrt(du^2+dv^2) (McCalib.cpp:2168-2245 averages these).
// err_ref != nullptr: the same error with the reference's float roundings (obs_error_ref), what its output files hold.
template <int STAGE>
__global__ void __launch_bounds__(OBS_WARPS * 32) k_reproj_err(ObsArgs a, double* __restrict__ err, float* __restrict__ err_ref) {
  typedef StageTraits<STAGE> ST;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ctx = smem + warp * CTX_SIZE;
  const int total_warps = gridDim.x * OBS_WARPS;
  for (int b = blockIdx.x * OBS_WARPS + warp; b < a.nb; b += total_warps) {
    const int4 rec = a.bobs[b];
    const int obs_begin = rec.x, obs_end = a.bobs[b + 1].x, cam = rec.y;
    const bool act_c = ST::CAM && !(a.cam_is_ref && a.cam_is_ref[cam]);
    const bool act_b = ST::BOARD && !(a.board_is_ref && a.board_is_ref[rec.w]);
    const double* pc = ST::CAM ? a.prep_cam + (size_t)PREP * cam : nullptr;
    const double* pb = ST::BOARD ? a.prep_board + (size_t)PREP * rec.w : nullptr;
    __syncwarp();
    for (int e = lane; e < CTX_SIZE; e += 32)
      if (e < CTX_NO || e >= CTX_IN) ctx[e] = ctx_entry_a(e, pc, act_c, a.prep_pose + (size_t)PREP * rec.z, pb, act_b, a.intr + 9 * (size_t)cam);
    __syncwarp();
    for (int i = obs_begin + lane; i < obs_end; i += 32) {
      const double* p3 = a.pts + 3 * (size_t)a.pt[i];
      const double X0[3] = {p3[0], p3[1], p3[2]};
      if (err_ref) { err_ref[i] = obs_error_ref(ctx, a.cam_model[cam], X0, a.u[i], a.v[i]); continue; }
      double ru, rv;
      obs_cost(ctx, a.cam_model[cam], X0, (double)a.u[i], (double)a.v[i], 0.0, ru, rv);
      err[i] = sqrt(ru * ru + rv * rv);
    }
  }
}

// SC_FAIL scalar: numerical-failure flag of this rank, plus P2P_ERR_UNIT if the peer-memory reduction timed out
constexpr double P2P_ERR_UNIT = 1048576.0;
__global__ void k_flag_to_scalar(const int* __restrict__ flag, const int* __restrict__ p2p_err, double* __restrict__ out) {
  *out = (double)*flag + ((p2p_err && *p2p_err) ? P2P_ERR_UNIT : 0.0);
}

// ---------------------------------------------------------------------------------------------
// Deterministic reductions: out[c] = op_c over rows of in[n][m]; columns < m_sum are summed, the others maxed.
// Stage 1: fixed contiguous chunks; stage 2: one CTA. Summation order depends only on (n, grid).
constexpr int RED_T = 256;
__device__ __forceinline__ double red_op(double x, double y, bool is_max) { return is_max ? fmax(x, y) : x + y; }
__global__ void __launch_bounds__(RED_T) k_reduce_rows(const double* __restrict__ in, int64_t n, int m, int m_sum,
                                                       int64_t rows_per_block, double* __restrict__ out) {
  __shared__ double sh[RED_T];
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(n, r0 + rows_per_block);
  for (int c = 0; c < m; ++c) {
    const bool is_max = c >= m_sum;
    double acc = is_max ? -DBL_MAX : 0.0;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += RED_T) acc = red_op(acc, in[r * m + c], is_max);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = RED_T / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] = red_op(sh[threadIdx.x], sh[threadIdx.x + s], is_max);
      __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)blockIdx.x * m + c] = sh[0];
    __syncthreads();
  }
}

// Column layout of the e-block row blocks Wt / Z and of Z^T Z: every camera block is padded to a multiple of 8 columns
// (Wcp), every board block to Wbp, and the right-hand side has an 8-column block of its own, so that the 8-wide DMMA
// tiles of the Schur SYRK never straddle two parameter blocks: a tile is then structurally zero for every frame that
// does not see its camera / board, and the SYRK skips it (k_syrk, activity masks). F^T F, S, z_f stay unpadded.
struct FLayout {
  int n_cam, Wc, Wcp, n_board, Wb, Wbp, n_f, ldz, rhs_col;
  __host__ __device__ int pcol(int j) const {
    if (Wcp == Wc && Wbp == Wb) return j;                  // unpadded (dense SYRK): identity
    const int ncw = n_cam * Wc;
    if (j < ncw) { const int c = j / Wc; return c * Wcp + (j - c * Wc); }
    j -= ncw;
    const int b = j / Wb;
    return n_cam * Wcp + b * Wbp + (j - b * Wb);
  }
};

// ---------------------------------------------------------------------------------------------
// K2a: per e-block sums over its board observations (fixed order): Hoo (6x6 full), and the dense row block
// Wt[o] = [E^T F (6 x n_f) | E^T r] with leading dimension ldz.
constexpr int EA_PAD = 6;   // row pitch of the e-block assembly's shared tile: with ldz a multiple of 16 doubles the six rows of a column would share a bank
template <int STAGE>
__global__ void __launch_bounds__(192) k_eblock_assemble(const int4* __restrict__ bobs, const int* __restrict__ pose_bobs_ptr,
                                                         const double* __restrict__ brec, FLayout L,
                                                         const unsigned* __restrict__ tile_mask, int mask_words,
                                                         double* __restrict__ Hoo, double* __restrict__ Wt) {
  typedef Cols<STAGE> C;
  extern __shared__ double W[];   // [6][lds], lds = ldz + EA_PAD: the six e-rows of one column land in different banks
  const int ldz = L.ldz, lds = L.ldz + EA_PAD;
  const int o = blockIdx.x, tid = threadIdx.x;
  // only the 8-column tiles of the blocks this e-block observes are ever non-zero (tile_mask, static per stage): the
  // others are neither cleared here nor written to Wt, which was zeroed once when it was allocated
  const unsigned* mk = tile_mask ? tile_mask + (size_t)o * mask_words : nullptr;
  for (int c = tid; c < ldz; c += blockDim.x) {
    const int t = c >> 3;
    if (!mk || ((mk[t >> 5] >> (t & 31)) & 1u)) {
#pragma unroll
      for (int k = 0; k < 6; ++k) W[k * lds + c] = 0.0;
    }
  }
  __syncthreads();
  const int b0 = pose_bobs_ptr[o], b1 = pose_bobs_ptr[o + 1];
  // every thread owns one entry of the record and walks the e-block's board observations in order (fixed summation
  // order); the loads of EA_U records are issued together - one dependent HBM round trip per record otherwise
  constexpr int EA_U = 1;   // records in flight per thread (8 measured slower: 127 vs 115 us; the kernel is not latency-bound)
  if (tid < 6 * C::WF) {
    const int j = tid / 6, k = tid % 6;                   // local f column j, e-row k: consecutive threads read consecutive
    const int off = C::cidx(j, C::E0 + k);                // doubles of the record's f x e part (CI_FE + 6 j + k = CI_FE + tid)
    for (int b = b0; b < b1; b += EA_U) {
      double v[EA_U]; int gc[EA_U];
#pragma unroll
      for (int u = 0; u < EA_U; ++u)
        if (b + u < b1) {
          const int4 rec = bobs[b + u];
          gc[u] = (j < C::B0) ? rec.y * L.Wcp + j : L.n_cam * L.Wcp + rec.w * L.Wbp + (j - C::B0);
          v[u] = brec[(size_t)(b + u) * C::NCOMP + off];
        }
#pragma unroll
      for (int u = 0; u < EA_U; ++u)
        if (b + u < b1) W[k * lds + gc[u]] += v[u];
    }
  } else if (tid < 6 * C::WF + 42) {
    const int e = tid - 6 * C::WF;
    int off;
    if (e < 36) { const int p = e / 6, r = e % 6; off = C::cidx(C::E0 + min(p, r), C::E0 + max(p, r)); }
    else off = C::cidx(C::E0 + (e - 36), C::R0);
    double s = 0.0;
    for (int b = b0; b < b1; b += EA_U) {
      double v[EA_U];
#pragma unroll
      for (int u = 0; u < EA_U; ++u) v[u] = (b + u < b1) ? brec[(size_t)(b + u) * C::NCOMP + off] : 0.0;
#pragma unroll
      for (int u = 0; u < EA_U; ++u) if (b + u < b1) s += v[u];
    }
    if (e < 36) Hoo[(size_t)o * 36 + e] = s;
    else W[(e - 36) * lds + L.rhs_col] = s;
  }
  __syncthreads();
  double* dst = Wt + (size_t)o * 6 * ldz;
  for (int c = tid; c < ldz; c += blockDim.x) {
    const int t = c >> 3;
    if (!mk || ((mk[t >> 5] >> (t & 31)) & 1u)) {
#pragma unroll
      for (int k = 0; k < 6; ++k) dst[k * ldz + c] = W[k * lds + c];
    }
  }
}

// K2b: per (camera[,board]) pair, fixed-order sum over the pair's board observations of the f x f upper
// triangle and the f x r column. Entry t of a pair record: a <= b < WF -> a*WF - a(a-1)/2 + (b-a);
// then WF(WF+1)/2 + a for (a, r).
template <int STAGE> struct PairRec {
  typedef Cols<STAGE> C;
  static constexpr int NFF = C::WF * (C::WF + 1) / 2;
  static constexpr int NE = NFF + C::WF;
  __host__ __device__ static int ff(int a, int b) { return a * C::WF - a * (a - 1) / 2 + (b - a); }   // a <= b
};
struct PairChunk { int pair, begin, end, first; };   // [begin,end) in pair_bobs; first = 1 for the pair's first chunk

template <int STAGE>
__global__ void __launch_bounds__(256) k_pair_reduce(const PairChunk* __restrict__ chunks, const int* __restrict__ pair_bobs,
                                                     const double* __restrict__ brec, double* __restrict__ partial) {
  typedef Cols<STAGE> C;
  typedef PairRec<STAGE> PR;
  const int t = threadIdx.x;
  if (t >= PR::NE) return;
  // entry t of a pair record is entry t of the compact per-bobs record (f x f upper, then f x r): contiguous reads
  const PairChunk ch = chunks[blockIdx.x];
  double s = 0.0;
  constexpr int PU = 8;   // loads of PU records in flight per thread; summation order unchanged
  for (int i = ch.begin; i < ch.end; i += PU) {
    double v[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) v[u] = (i + u < ch.end) ? brec[(size_t)pair_bobs[i + u] * C::NCOMP + t] : 0.0;
#pragma unroll
    for (int u = 0; u < PU; ++u) if (i + u < ch.end) s += v[u];
  }
  partial[(size_t)blockIdx.x * PR::NE + t] = s;
}
template <int STAGE>
__global__ void k_pair_final(const int* __restrict__ pair_chunk_ptr, int n_pair, const double* __restrict__ partial,
                             double* __restrict__ pair_rec) {
  typedef PairRec<STAGE> PR;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pair * PR::NE) return;
  const int p = idx / PR::NE, t = idx % PR::NE;
  double s = 0.0;
  for (int c = pair_chunk_ptr[p]; c < pair_chunk_ptr[p + 1]; ++c) s += partial[(size_t)c * PR::NE + t];
  pair_rec[idx] = s;
}

// Dense F^T F (n_f x n_f, full symmetric, unscaled) and F^T r from the pair records.
template <int STAGE>
__global__ void k_ff_assemble(const double* __restrict__ pair_rec, int n_cam, int n_board, int Wc, int Wb, int n_f,
                              double* __restrict__ FF, double* __restrict__ gf) {
  typedef Cols<STAGE> C;
  typedef PairRec<STAGE> PR;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nbp = C::T::BOARD ? n_board : 1;   // pairs per camera
  if (idx < n_f * n_f) {
    const int i = idx / n_f, j = idx % n_f;
    const int lo = min(i, j), hi = max(i, j);
    const int ncw = n_cam * Wc;
    double s = 0.0;
    if (hi < ncw) {                         // camera x camera
      const int c = lo / Wc;
      if (hi / Wc == c) { const int e = PR::ff(lo - c * Wc, hi - c * Wc); for (int b = 0; b < nbp; ++b) s += pair_rec[(size_t)(c * nbp + b) * PR::NE + e]; }
    } else if (lo >= ncw) {                 // board x board
      const int b = (lo - ncw) / Wb;
      if ((hi - ncw) / Wb == b) { const int e = PR::ff(C::B0 + lo - ncw - b * Wb, C::B0 + hi - ncw - b * Wb); for (int c = 0; c < n_cam; ++c) s += pair_rec[(size_t)(c * nbp + b) * PR::NE + e]; }
    } else {                                // camera x board
      const int c = lo / Wc, b = (hi - ncw) / Wb;
      s = pair_rec[(size_t)(c * nbp + b) * PR::NE + PR::ff(lo - c * Wc, C::B0 + hi - ncw - b * Wb)];
    }
    FF[idx] = s;
  } else if (idx < n_f * n_f + n_f) {
    const int i = idx - n_f * n_f;
    const int ncw = n_cam * Wc;
    double s = 0.0;
    if (i < ncw) { const int c = i / Wc; for (int b = 0; b < nbp; ++b) s += pair_rec[(size_t)(c * nbp + b) * PR::NE + PR::NFF + (i - c * Wc)]; }
    else { const int b = (i - ncw) / Wb; for (int c = 0; c < n_cam; ++c) s += pair_rec[(size_t)(c * nbp + b) * PR::NE + PR::NFF + C::B0 + (i - ncw - b * Wb)]; }
    gf[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Jacobi scaling (TrustRegionMinimizer, iteration 0): s_j = 1/(1+sqrt(|J_j|^2)); LM diagonal
// (LevenbergMarquardtStrategy::ComputeStep): diag_j = clamp(s_j^2 |J_j|^2, min, max).
__global__ void k_scale_diag(const double* __restrict__ Hoo, int n_pose, const double* __restrict__ FF, int n_f,
                             int compute_scale, int jacobi, double min_diag, double max_diag,
                             double* __restrict__ scale_e, double* __restrict__ scale_f,
                             double* __restrict__ diag_e, double* __restrict__ diag_f) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ne = n_pose * 6;
  if (idx >= ne + n_f) return;
  const bool is_e = idx < ne;
  const int j = is_e ? idx : idx - ne;
  const double h = is_e ? Hoo[(size_t)(j / 6) * 36 + (j % 6) * 7] : FF[(size_t)j * n_f + j];
  double* sc = is_e ? scale_e : scale_f;
  if (compute_scale) sc[j] = jacobi ? 1.0 / (1.0 + sqrt(h)) : 1.0;
  const double s = sc[j];
  const double d = fmin(fmax(s * s * h, min_diag), max_diag);
  (is_e ? diag_e : diag_f)[j] = d;
}

// Gradient max-norm pieces |x - (x + (-g))| (computed through Plus as Ceres does) and |x|^2, per e-block.
// part[o] = {sum x^2, max |x - (x - g)|}. A non-finite gradient entry or Jacobian column (its squared norm is the
// diagonal of E^T E) turns the max into +inf: Ceres' evaluator rejects non-finite residuals, gradients AND Jacobian
// values (IsArrayValid), not only a non-finite cost.
__global__ void k_eblock_grad(const double* __restrict__ pose, const double* __restrict__ Wt, const double* __restrict__ Hoo,
                              int n_pose, int n_f, int ldz, double* __restrict__ part) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_pose) return;
  double s = 0.0, m = 0.0;
  bool finite = true;
  for (int k = 0; k < 6; ++k) {
    const double x = pose[(size_t)o * 6 + k], gk = Wt[((size_t)o * 6 + k) * ldz + n_f];
    const double pg = x + (-gk);
    s += x * x; m = fmax(m, fabs(x - pg));
    finite = finite && isfinite(gk) && isfinite(Hoo[(size_t)o * 36 + k * 7]);
  }
  part[(size_t)o * 2] = s; part[(size_t)o * 2 + 1] = finite ? m : CUDART_INF;
}

// K2c: factor the e-block, Z[o] = L^-1 (s_e [W | g] s_f). One CTA per e-block; every thread factors its own
// copy of the 6x6 in registers (cheaper than a barrier), then handles columns j = tid, tid+T, ...
__global__ void __launch_bounds__(128) k_eblock_factor(const double* __restrict__ Hoo, const double* __restrict__ Wt,
                                                       const double* __restrict__ scale_e, const double* __restrict__ scale_f,
                                                       const double* __restrict__ diag_e, const double* __restrict__ radius_p, FLayout Lo,
                                                       const unsigned* __restrict__ tile_mask, int mask_words,
                                                       double* __restrict__ Lout, double* __restrict__ Z,
                                                       int* __restrict__ fail) {
  const int o = blockIdx.x;
  const double radius = *radius_p;   // device-resident: the step is replayed as a CUDA graph, only this value changes
  const unsigned* mk = tile_mask ? tile_mask + (size_t)o * mask_words : nullptr;   // columns outside the observed blocks are zero in Wt and stay zero in Z
  const int n_f = Lo.n_f, ldz = Lo.ldz;
  double L[36], se[6], inv[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) se[k] = scale_e[(size_t)o * 6 + k];
#pragma unroll
  for (int p = 0; p < 6; ++p)
#pragma unroll
    for (int r = 0; r <= p; ++r) L[p * 6 + r] = se[p] * Hoo[(size_t)o * 36 + p * 6 + r] * se[r];
#pragma unroll
  for (int k = 0; k < 6; ++k) L[k * 7] += diag_e[(size_t)o * 6 + k] / radius;
  const bool ok = chol6_inv(L, inv);    // every thread its own copy, in registers (cheaper than a barrier)
  if (threadIdx.x == 0) {
    if (!ok) atomicExch(fail, 1);
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int r = 0; r < 6; ++r) Lout[(size_t)o * 36 + p * 6 + r] = (ok && r <= p) ? L[p * 6 + r] : 0.0;
  }
  const double* W = Wt + (size_t)o * 6 * ldz;
  double* Zo = Z + (size_t)o * 6 * ldz;
  // the 15 strictly-lower entries as scalars: the column loop below must not read the factor back from local memory
  const double l10 = L[6], l20 = L[12], l21 = L[13], l30 = L[18], l31 = L[19], l32 = L[20], l40 = L[24], l41 = L[25], l42 = L[26],
               l43 = L[27], l50 = L[30], l51 = L[31], l52 = L[32], l53 = L[33], l54 = L[34];
  const double i0 = inv[0], i1 = inv[1], i2 = inv[2], i3 = inv[3], i4 = inv[4], i5 = inv[5];
  // one real column (f index j, then the rhs as j = n_f) per thread and pass; the padding columns of Z stay zero
  for (int j = threadIdx.x; j <= n_f; j += blockDim.x) {
    const int pc = (j < n_f) ? Lo.pcol(j) : Lo.rhs_col;
    if (mk && !((mk[pc >> 8] >> ((pc >> 3) & 31)) & 1u)) continue;      // inactive tile (sparse layout only)
    const double sf = (j < n_f) ? scale_f[j] : 1.0;
    const double b0 = se[0] * W[pc] * sf, b1 = se[1] * W[ldz + pc] * sf, b2 = se[2] * W[2 * ldz + pc] * sf,
                 b3 = se[3] * W[3 * ldz + pc] * sf, b4 = se[4] * W[4 * ldz + pc] * sf, b5 = se[5] * W[5 * ldz + pc] * sf;
    const double w0 = b0 * i0;
    const double w1 = (b1 - l10 * w0) * i1;
    const double w2 = (b2 - l20 * w0 - l21 * w1) * i2;
    const double w3 = (b3 - l30 * w0 - l31 * w1 - l32 * w2) * i3;
    const double w4 = (b4 - l40 * w0 - l41 * w1 - l42 * w2 - l43 * w3) * i4;
    const double w5 = (b5 - l50 * w0 - l51 * w1 - l52 * w2 - l53 * w3 - l54 * w4) * i5;
    Zo[pc] = ok ? w0 : 0.0; Zo[ldz + pc] = ok ? w1 : 0.0; Zo[2 * ldz + pc] = ok ? w2 : 0.0;
    Zo[3 * ldz + pc] = ok ? w3 : 0.0; Zo[4 * ldz + pc] = ok ? w4 : 0.0; Zo[5 * ldz + pc] = ok ? w5 : 0.0;
  }
}

// K2d: C = Z^T Z restricted to upper tile pairs, split over row ranges. 64x64 tile per CTA, 4 warps (2x2),
// each warp 32x32 = 4x4 DMMA tiles; K chunk of 16 rows staged through shared memory with register prefetch.
constexpr int SY_T = 128, SY_K = 32, SY_LD = SY_T + 4, SY_STAGES = 3, SY_THREADS = 256;
constexpr int SY_SMEM = SY_STAGES * 2 * SY_K * SY_LD * 8;   // dynamic shared memory of k_syrk (202,752 B: one CTA per SM)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {   // src_bytes 0 -> zero fill
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// 128x128 output tile per CTA (halves the L2 traffic of a 64x64 tiling: every CTA streams its two 128-column strips
// of Z), 8 warps (4x2) of 32x64 = 4x8 DMMA tiles each; rows of Z stream through a 3-stage cp.async ring of 32-row
// chunks (one barrier per 256 DMMAs per warp, loads two chunks ahead of the math). On diagonal tiles the B strip is
// the A strip and warps entirely below the diagonal skip the math.
__global__ void __launch_bounds__(SY_THREADS) k_syrk(const double* __restrict__ Z, int64_t n_rows, int ldz, int nt,
                                                     int64_t rows_per_split, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sy_sm[];
  int ti = 0, rem = blockIdx.x;   // tile pair from linear index (ti <= tj)
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  const int tj = ti + rem;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * rows_per_split, r1 = min(n_rows, r0 + rows_per_split);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
  const bool diag = (ti == tj);
  const bool skip = diag && (wm >= wn + 64);   // warp tile strictly below the diagonal of a diagonal tile
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  const int n_chunks = (r1 > r0) ? (int)((r1 - r0 + SY_K - 1) / SY_K) : 0;
  // per-thread copy slots: SY_K rows x 128 cols = SY_K*64 16-byte pieces per operand, SY_K/4 per thread
  const int c2 = (tid & 63) * 2, row0 = tid >> 6;            // piece i: row = row0 + 4 i
  const bool ca_ok = ti * SY_T + c2 < ldz, cb_ok = tj * SY_T + c2 < ldz;
  const double* ga = Z + (r0 + row0) * ldz + ti * SY_T + c2;
  const double* gb = Z + (r0 + row0) * ldz + tj * SY_T + c2;
  auto issue = [&](int chunk) {
    if (chunk < n_chunks) {
      double* As = sy_sm + (size_t)(chunk % SY_STAGES) * (2 * SY_K * SY_LD) + row0 * SY_LD + c2;
      double* Bs = As + SY_K * SY_LD;
      const int64_t rbase = r0 + (int64_t)chunk * SY_K + row0;
      const double* pa = ga + (int64_t)chunk * SY_K * ldz;
      const double* pb = gb + (int64_t)chunk * SY_K * ldz;
#pragma unroll
      for (int i = 0; i < SY_K / 4; ++i) {
        const bool rv = rbase + 4 * i < r1;
        const bool va = rv && ca_ok, vb = rv && cb_ok;
        cp_async16(As + 4 * i * SY_LD, va ? (const void*)(pa + (int64_t)4 * i * ldz) : (const void*)Z, va ? 16 : 0);
        if (!diag) cp_async16(Bs + 4 * i * SY_LD, vb ? (const void*)(pb + (int64_t)4 * i * ldz) : (const void*)Z, vb ? 16 : 0);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < SY_STAGES - 1; ++s) issue(s);
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    cp_async_wait<SY_STAGES - 2>();
    __syncthreads();
    issue(chunk + SY_STAGES - 1);
    if (skip) continue;
    const double* As = sy_sm + (size_t)(chunk % SY_STAGES) * (2 * SY_K * SY_LD);
    const double* Bp = diag ? As : As + SY_K * SY_LD;
#pragma unroll
    for (int ks = 0; ks < SY_K / 4; ++ks) {
      double fa[4], fb[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[i] = As[(4 * ks + q) * SY_LD + wm + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 8; ++j) fb[j] = Bp[(4 * ks + q) * SY_LD + wn + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  }
  cp_async_wait<0>();
  double* out = partial + ((size_t)split * gridDim.x + blockIdx.x) * (SY_T * SY_T);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<double2*>(out + (wm + 8 * i + g) * SY_T + wn + 8 * j + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
}
// Fixed-order sum of the split-K partials into the dense upper part of ZtZ (ldz x ldz).
__global__ void k_syrk_sum(const double* __restrict__ partial, int n_tiles, int nt, int nsplit, int ldz,
                           double* __restrict__ ZtZ) {
  const int tile = blockIdx.x;
  int ti = 0, rem = tile;
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  const int tj = ti + rem;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < SY_T * SY_T; e += gridDim.y * blockDim.x) {
    const int i = ti * SY_T + e / SY_T, j = tj * SY_T + e % SY_T;
    if (i >= ldz || j >= ldz || j < i) continue;   // only the upper triangle is consumed
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((size_t)sp * n_tiles + tile) * (SY_T * SY_T) + e];
    ZtZ[(size_t)i * ldz + j] = s;
  }
}

// K2d, multi-GPU: the fixed-order split-K sum fused with the cross-GPU reduction of the Schur complement, over
// NVLink peer memory instead of an NCCL allreduce - one kernel, three phases:
//   1. every CTA sums this rank's split-K partials (fixed order) into this rank's staging buffer (tile layout
//      [tile][SY_T*SY_T]); the last CTA to finish publishes "staged" (epoch) into every rank's flag array;
//   2. rank r owns the contiguous slice [r, r+1) * ceil(n_tiles*SY_T^2 / n_ranks) of the staging layout: it loads that
//      slice from EVERY rank's staging buffer (independent peer loads, all in flight together), sums in rank order - so
//      all ranks end up with bit-identical sums - and stores the result into every rank's ZtZ (peer stores): a
//      reduce-scatter and an all-gather in one pass;
//   3. the last CTA publishes "written" and waits until every rank has written its slice into this rank's ZtZ, so
//      that the next kernel in the stream may consume ZtZ and the next iteration may overwrite the staging buffer.
// Spins are bounded (P2P_TIMEOUT cycles -> *err, reported through the SC_FAIL scalar). All CTAs must be co-resident
// (grid = 2 CTAs per SM).
constexpr long long P2P_TIMEOUT = 6000000000LL;
constexpr int P2P_MAX_RANKS = 8;
struct P2PArgs {
  double* stage[P2P_MAX_RANKS];                // staging buffers (peer-mapped), [n_tiles][SY_T*SY_T]
  double* ZtZ[P2P_MAX_RANKS];                  // [ldz][ldz]
  unsigned long long* flags[P2P_MAX_RANKS];    // each [2][P2P_MAX_RANKS]: "staged" and "written" epochs per source rank
  const double* partial;                       // this rank's split-K partials
  int nsplit, rank, n_ranks;
  const unsigned long long* epoch_p;   // launch counter in device memory (bumped before every launch: graph replay)
  int n_tiles, nt, ldz;
  unsigned int* counter;                       // [2], zero between launches
  int* err;
};
__device__ __forceinline__ bool p2p_wait(const volatile unsigned long long* flags, int n, unsigned long long epoch, int* err) {
  const long long t0 = clock64();
  for (int p = 0; p < n; ++p)
    while (flags[p] < epoch) {
      if (clock64() - t0 > P2P_TIMEOUT) { *err = 1; return false; }
      __nanosleep(32);
    }
  __threadfence_system();
  return true;
}
__global__ void k_bump_epoch64(unsigned long long* e) { *e += 1ull; }
__device__ __forceinline__ void p2p_publish(const P2PArgs& a, int slot) {
  __threadfence_system();
  const unsigned long long epoch = *a.epoch_p;
  for (int p = 0; p < a.n_ranks; ++p) ((volatile unsigned long long*)a.flags[p])[slot * P2P_MAX_RANKS + a.rank] = epoch;
}
// (row, column) in ZtZ of element idx of the staging layout; false for entries nobody consumes
__device__ __forceinline__ bool p2p_locate(int64_t idx, int nt, int ldz, int* i, int* j) {
  const int tile = (int)(idx / (SY_T * SY_T)), e = (int)(idx % (SY_T * SY_T));
  int ti = 0, rem = tile;
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  *i = ti * SY_T + e / SY_T; *j = (ti + rem) * SY_T + e % SY_T;
  return *i < ldz && *j < ldz && *j >= *i;
}
__global__ void __launch_bounds__(256) k_syrk_sum_p2p(P2PArgs a) {
  __shared__ int ok_s;
  const unsigned long long epoch = *a.epoch_p;
  const int64_t total = (int64_t)a.n_tiles * SY_T * SY_T;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, gstride = (int64_t)gridDim.x * blockDim.x;
  // phase 1
  double* my_stage = a.stage[a.rank];
  for (int64_t idx = gtid; idx < total; idx += gstride) {
    int i, j;
    if (!p2p_locate(idx, a.nt, a.ldz, &i, &j)) continue;
    double s = 0.0;
    for (int sp = 0; sp < a.nsplit; ++sp) s += a.partial[(size_t)sp * total + idx];
    my_stage[idx] = s;
  }
  __threadfence_system();
  __syncthreads();
  volatile unsigned long long* my_flags = a.flags[a.rank];
  if (threadIdx.x == 0) {
    if (atomicAdd(a.counter, 1u) == gridDim.x - 1) { a.counter[0] = 0; p2p_publish(a, 0); }
    ok_s = p2p_wait(my_flags, a.n_ranks, epoch, a.err) ? 1 : 0;
  }
  __syncthreads();
  // phase 2
  if (ok_s) {
    const int64_t slice = (total + a.n_ranks - 1) / a.n_ranks;
    const int64_t lo = slice * a.rank, hi = (lo + slice < total) ? lo + slice : total;
    for (int64_t idx = lo + gtid; idx < hi; idx += gstride) {
      int i, j;
      if (!p2p_locate(idx, a.nt, a.ldz, &i, &j)) continue;
      double v[P2P_MAX_RANKS];
#pragma unroll
      for (int p = 0; p < P2P_MAX_RANKS; ++p) v[p] = (p < a.n_ranks) ? __ldcg(a.stage[p] + idx) : 0.0;
      double s = v[0];
#pragma unroll
      for (int p = 1; p < P2P_MAX_RANKS; ++p) if (p < a.n_ranks) s += v[p];
      const size_t o = (size_t)i * a.ldz + j;
#pragma unroll
      for (int p = 0; p < P2P_MAX_RANKS; ++p) if (p < a.n_ranks) a.ZtZ[p][o] = s;
    }
  }
  // phase 3
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(a.counter + 1, 1u) == gridDim.x - 1) {
    a.counter[1] = 0;
    p2p_publish(a, 1);
    p2p_wait(my_flags + P2P_MAX_RANKS, a.n_ranks, epoch, a.err);
  }
}

// The other two exchanges of an LM iteration over the same peer-memory mechanism (no NCCL on the iteration path):
// k_allreduce_p2p  sum of the pair records (F^T F blocks, F^T r; 1.3 MB on the dome) across ranks: every rank's input
//                  vector is peer-mapped; rank r sums slice r from all ranks in rank order (bit-identical results
//                  everywhere) and stores it into every rank's output vector; flag slots P2P_CH_PAIR, +1.
// k_gather_p2p     the 12 scalars of a phase (cost, norms, model-cost pieces, failure flag): every rank writes its row
//                  into every rank's table; double-buffered by the parity of the epoch; flag slot P2P_CH_GATHER.
constexpr int P2P_CH_PAIR = 2, P2P_CH_GATHER = 4, P2P_FLAG_SLOTS = 6, P2P_GATHER_W = 16;
struct P2PVecArgs {
  double* in[P2P_MAX_RANKS]; double* out[P2P_MAX_RANKS]; unsigned long long* flags[P2P_MAX_RANKS];
  int n, rank, n_ranks; const unsigned long long* epoch_p; unsigned int* counter; int* err;
};
__global__ void __launch_bounds__(256) k_allreduce_p2p(P2PVecArgs a) {
  __shared__ int ok_s;
  const unsigned long long epoch = *a.epoch_p;
  volatile unsigned long long* my_flags = a.flags[a.rank];
  // the input was written by the previous kernel of this stream: publish it, then wait for everybody's
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int p = 0; p < a.n_ranks; ++p) ((volatile unsigned long long*)a.flags[p])[P2P_CH_PAIR * P2P_MAX_RANKS + a.rank] = epoch;
    }
    ok_s = p2p_wait(my_flags + P2P_CH_PAIR * P2P_MAX_RANKS, a.n_ranks, epoch, a.err) ? 1 : 0;
  }
  __syncthreads();
  if (ok_s) {
    const int slice = (a.n + a.n_ranks - 1) / a.n_ranks;
    const int lo = slice * a.rank, hi = min(a.n, lo + slice);
    for (int idx = lo + blockIdx.x * blockDim.x + threadIdx.x; idx < hi; idx += gridDim.x * blockDim.x) {
      double v[P2P_MAX_RANKS];
#pragma unroll
      for (int p = 0; p < P2P_MAX_RANKS; ++p) v[p] = (p < a.n_ranks) ? __ldcg(a.in[p] + idx) : 0.0;
      double s = v[0];
#pragma unroll
      for (int p = 1; p < P2P_MAX_RANKS; ++p) if (p < a.n_ranks) s += v[p];
#pragma unroll
      for (int p = 0; p < P2P_MAX_RANKS; ++p) if (p < a.n_ranks) a.out[p][idx] = s;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(a.counter, 1u) == gridDim.x - 1) {
    a.counter[0] = 0;
    __threadfence_system();
    for (int p = 0; p < a.n_ranks; ++p) ((volatile unsigned long long*)a.flags[p])[(P2P_CH_PAIR + 1) * P2P_MAX_RANKS + a.rank] = epoch;
    p2p_wait(my_flags + (P2P_CH_PAIR + 1) * P2P_MAX_RANKS, a.n_ranks, epoch, a.err);
  }
}
struct P2PGatherArgs {
  double* table[P2P_MAX_RANKS]; unsigned long long* flags[P2P_MAX_RANKS];
  const double* local; int m, rank, n_ranks; const unsigned long long* epoch_p; int* err;
};
__global__ void __launch_bounds__(256) k_gather_p2p(P2PGatherArgs a) {
  const unsigned long long epoch = *a.epoch_p;
  const int region = (int)(epoch & 1ull) * P2P_MAX_RANKS * P2P_GATHER_W;
  for (int t = threadIdx.x; t < a.m * a.n_ranks; t += blockDim.x) {
    const int p = t / a.m, i = t % a.m;
    a.table[p][region + a.rank * P2P_GATHER_W + i] = a.local[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.n_ranks) ((volatile unsigned long long*)a.flags[threadIdx.x])[P2P_CH_GATHER * P2P_MAX_RANKS + a.rank] = epoch;
  if (threadIdx.x == 0) p2p_wait((volatile unsigned long long*)a.flags[a.rank] + P2P_CH_GATHER * P2P_MAX_RANKS, a.n_ranks, epoch, a.err);
}

// K2e: reduced system. S = s_f F^T F s_f + D_f^2 - Z^T Z (full symmetric, n_f x n_f, ld n_f); rhs = s_f g_f - Z^T y.
__global__ void k_build_reduced(const double* __restrict__ FF, const double* __restrict__ gf,
                                const double* __restrict__ ZtZ, const double* __restrict__ scale_f,
                                const double* __restrict__ diag_f, const double* __restrict__ radius_p, FLayout L,
                                double* __restrict__ S, double* __restrict__ rhs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const double radius = *radius_p;
  const int n_f = L.n_f, ldz = L.ldz;
  if (idx < n_f * n_f) {
    const int i = idx / n_f, j = idx % n_f;
    const int lo = min(i, j), hi = max(i, j);
    double s = scale_f[i] * FF[idx] * scale_f[j] - ZtZ[(size_t)L.pcol(lo) * ldz + L.pcol(hi)];
    if (i == j) s += diag_f[i] / radius;
    S[idx] = s;
  } else if (idx < n_f * n_f + n_f) {
    const int i = idx - n_f * n_f;
    const double v = scale_f[i] * gf[i] - ZtZ[(size_t)L.pcol(i) * ldz + L.rhs_col];
    rhs[i] = v;
    S[(size_t)n_f * n_f + i] = v;   // extra row n_f: the factorisation turns it into L^-1 rhs
  }
}

// ---------------------------------------------------------------------------------------------
// K3: one cooperative persistent kernel for the whole factorisation + both triangular solves.
// S is (n+1) x n row-major: rows 0..n-1 the symmetric reduced matrix (lower part used), row n the rhs.
// Right-looking with 64-wide panels. Per panel: every CTA factors the 64x64 diagonal block redundantly in shared
// memory (no broadcast needed), warps solve one row each of the panel below it (the rhs row included, which
// performs the forward substitution for free), grid sync, DMMA trailing update on 64x64 lower tiles, grid sync.
// CTA 0 finishes with the backward substitution L^T z = y.
constexpr int CC_B = 64, CC_LD = 68, CC_DL = 65, CC_THREADS = 256;
constexpr int CC_SMEM = (2 * CC_B * CC_DL + 2 * CC_B * CC_LD + CC_B + 2 * CC_B + 4 * CC_B) * 8;
__global__ void __launch_bounds__(CC_THREADS) k_chol_coop(double* __restrict__ S, int n, double* __restrict__ z,
                                                          int* __restrict__ fail, long long* __restrict__ prof,
                                                          int backsolve_mode) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double sm[];
  double* D = sm;                        // [64][65] working copy of the diagonal block
  double* Lm = D + CC_B * CC_DL;         // [64][65] its Cholesky factor
  double* Pa = Lm + CC_B * CC_DL;        // [64][68] panel rows of tile i
  double* Pb = Pa + CC_B * CC_LD;        // [64][68] panel rows of tile j
  double* invd = Pb + CC_B * CC_LD;      // [64] 1 / L_jj
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  long long tprev = clock64();
#define CC_PROF(slot) do { if (prof && blockIdx.x == 0 && tid == 0) { const long long tn = clock64(); prof[slot] += tn - tprev; tprev = tn; } } while (0)
  for (int j0 = 0; j0 < n; j0 += CC_B) {
    const int jb = min(CC_B, n - j0);
    // ---- diagonal block in shared memory, factored in 8-wide sub-panels: (a) all threads apply the previous columns
    // to the sub-panel, (b) every row-thread factors the 8x8 diagonal sub-block redundantly in registers (no barrier
    // inside the 8 dependent pivots) and solves its own row of the sub-panel.
    for (int e = tid; e < CC_B * CC_B; e += CC_THREADS) {
      const int r = e >> 6, c = e & 63;
      D[r * CC_DL + c] = (r < jb && c <= r) ? S[(size_t)(j0 + r) * n + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int c0 = 0; c0 < jb; c0 += 8) {
      if (c0 > 0) {
        for (int e = tid; e < (CC_B - c0) * 8; e += CC_THREADS) {
          const int i = c0 + (e >> 3), c = c0 + (e & 7);
          if (c <= i) {
            double s0 = D[i * CC_DL + c], s1 = 0.0;
            for (int k = 0; k < c0; k += 2) {
              s0 -= Lm[i * CC_DL + k] * Lm[c * CC_DL + k];
              s1 -= Lm[i * CC_DL + k + 1] * Lm[c * CC_DL + k + 1];
            }
            D[i * CC_DL + c] = s0 + s1;
          }
        }
        __syncthreads();
      }
      if (tid < CC_B - c0) {
        const int i = c0 + tid;
        double l[8][8], inv[8], x[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) l[r][c] = D[(c0 + r) * CC_DL + c0 + c];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double d = l[j][j];
#pragma unroll
          for (int k = 0; k < j; ++k) d -= l[j][k] * l[j][k];
          const bool b = !(d > 0.0) || !(d < DBL_MAX);
          bad = bad || b;
          inv[j] = b ? 1.0 : rsqrt(d);
          l[j][j] = d * inv[j];
#pragma unroll
          for (int r = j + 1; r < 8; ++r) {
            double t = l[r][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t -= l[r][k] * l[j][k];
            l[r][j] = t * inv[j];
          }
        }
        if (bad && tid == 0 && blockIdx.x == 0) atomicExch(fail, 1);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          double t = D[i * CC_DL + c0 + c];
#pragma unroll
          for (int k = 0; k < c; ++k) t -= x[k] * l[c][k];
          x[c] = t * inv[c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) Lm[i * CC_DL + c0 + c] = x[c];
        if (tid == 0) {
#pragma unroll
          for (int c = 0; c < 8; ++c) invd[c0 + c] = inv[c];
        }
      }
      __syncthreads();
    }
    CC_PROF(0);
    // ---- panel below (rows base..n, the rhs row n included): one row per warp, column-oriented substitution
    const int base = j0 + jb;
    for (int r = base + blockIdx.x * (CC_THREADS / 32) + warp; r <= n; r += gridDim.x * (CC_THREADS / 32)) {
      double* row = S + (size_t)r * n + j0;
      double a0 = (lane < jb) ? row[lane] : 0.0, a1 = (lane + 32 < jb) ? row[lane + 32] : 0.0;
      for (int j = 0; j < jb; ++j) {
        const double xj = __shfl_sync(0xffffffffu, (j < 32) ? a0 : a1, j & 31) * invd[j];
        if (lane == (j & 31)) { if (j < 32) a0 = xj; else a1 = xj; }
        if (lane > j && lane < jb) a0 -= xj * Lm[lane * CC_DL + j];
        if (lane + 32 > j && lane + 32 < jb) a1 -= xj * Lm[(lane + 32) * CC_DL + j];
      }
      if (lane < jb) row[lane] = a0;
      if (lane + 32 < jb) row[lane + 32] = a1;
    }
    CC_PROF(1);
    __threadfence();
    grid.sync();
    CC_PROF(2);
    // every CTA has loaded the diagonal block by now: CTA 0 may overwrite it with its factor
    if (blockIdx.x == 0)
      for (int e = tid; e < jb * jb; e += CC_THREADS) { const int r = e / jb, c = e % jb; if (c <= r) S[(size_t)(j0 + r) * n + j0 + c] = Lm[r * CC_DL + c]; }
    if (base >= n) break;   // last panel: only the rhs row was below it
    // ---- trailing update on lower 64x64 tiles: C -= Pa Pb^T with DMMA; rows base..n (incl. rhs row), cols base..n-1
    const int ntr = (n + 1 - base + CC_B - 1) / CC_B, ntc = (n - base + CC_B - 1) / CC_B;
    int ntiles = 0;
    for (int ti = 0; ti < ntr; ++ti) ntiles += min(ti, ntc - 1) + 1;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int ti = 0, rem = t;
      while (rem >= min(ti, ntc - 1) + 1) { rem -= min(ti, ntc - 1) + 1; ++ti; }
      const int tj = rem;
      const int row0 = base + ti * CC_B, col0 = base + tj * CC_B;
      __syncthreads();
      {
        double va[16], vb[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int e = tid + it * CC_THREADS, r = e >> 6, c = e & 63;
          va[it] = (row0 + r <= n && c < jb) ? S[(size_t)(row0 + r) * n + j0 + c] : 0.0;
          vb[it] = (col0 + r < n && c < jb) ? S[(size_t)(col0 + r) * n + j0 + c] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int e = tid + it * CC_THREADS, r = e >> 6, c = e & 63;
          Pa[r * CC_LD + c] = va[it];
          Pb[r * CC_LD + c] = vb[it];
        }
      }
      __syncthreads();
      const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
      double acc[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
      const int ksteps = (jb + 3) >> 2;
      for (int ks = 0; ks < ksteps; ++ks) {
        double fa[4], fb[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[i] = Pa[(wm + 8 * i + g) * CC_LD + 4 * ks + q];
#pragma unroll
        for (int j = 0; j < 2; ++j) fb[j] = Pb[(wn + 8 * j + g) * CC_LD + 4 * ks + q];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
      }
      {
        double cur[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int gr = row0 + wm + 8 * i + g, gc = col0 + wn + 8 * j + 2 * q + h;
              cur[i][j][h] = (gr <= n && gc < n && gc <= gr) ? S[(size_t)gr * n + gc] : 0.0;
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int gr = row0 + wm + 8 * i + g, gc = col0 + wn + 8 * j + 2 * q + h;
              if (gr <= n && gc < n && gc <= gr) S[(size_t)gr * n + gc] = cur[i][j][h] - acc[i][j][h];
            }
      }
    }
    CC_PROF(3);
    __threadfence();
    grid.sync();
    CC_PROF(4);
  }
  // ---- backward substitution L^T z = y (y = row n of S).
  // backsolve_mode 1 (default): CTA 0 alone, no grid sync. y lives in shared memory; per 64-block b (from the last one)
  // warp 0 runs the 64-step solve of the diagonal block while warps 1-7 fetch the next diagonal block and apply the
  // DEFERRED part of the previous block's update (columns left of the next chunk); then all warps apply block b's
  // update to the one chunk the next solve needs. No grid sync, but all update traffic (n^2/2 doubles) goes through
  // one SM: chosen by the host for small systems only (mcba_api.cu, cholesky_reduced).
  if (backsolve_mode == 1) {
    if (blockIdx.x != 0) return;
    double* yrow = S + (size_t)n * n;
    double* ysm = Pa;                     // [n <= 2*CC_B*CC_LD]: Pa and Pb are contiguous
    double* zbuf = invd + CC_B;           // [2][64] z of the block being solved / of the previous block
    double* red = invd + 3 * CC_B;        // [4][64]
    double* Lcur = Lm; double* Lnext = D;
    const int nblk = (n + CC_B - 1) / CC_B;
    __syncthreads();
    for (int i = tid; i < n; i += CC_THREADS) ysm[i] = yrow[i];
    {
      const int j0 = (nblk - 1) * CC_B, jb = min(CC_B, n - j0);
      for (int e = tid; e < CC_B * CC_B; e += CC_THREADS) {
        const int r = e >> 6, c = e & 63;
        Lcur[r * CC_DL + c] = (r < jb && c <= r) ? S[(size_t)(j0 + r) * n + j0 + c] : 0.0;
      }
    }
    __syncthreads();
    for (int b = nblk - 1; b >= 0; --b) {
      const int j0 = b * CC_B, jb = min(CC_B, n - j0);
      double* zb = zbuf + (b & 1) * CC_B;
      const double* zprev = zbuf + ((b + 1) & 1) * CC_B;
      if (warp == 0) {
        if (lane < jb) invd[lane] = 1.0 / Lcur[lane * CC_DL + lane];
        if (lane + 32 < jb) invd[lane + 32] = 1.0 / Lcur[(lane + 32) * CC_DL + lane + 32];
        __syncwarp();
        double a0 = (lane < jb) ? ysm[j0 + lane] : 0.0, a1 = (lane + 32 < jb) ? ysm[j0 + lane + 32] : 0.0;
        for (int j = jb - 1; j >= 0; --j) {
          const double zj = __shfl_sync(0xffffffffu, (j < 32) ? a0 : a1, j & 31) * invd[j];
          if (lane == (j & 31)) { if (j < 32) a0 = zj; else a1 = zj; }
          if (lane < j) a0 -= zj * Lcur[j * CC_DL + lane];
          if (lane + 32 < j) a1 -= zj * Lcur[j * CC_DL + lane + 32];
        }
        zb[lane] = (lane < jb) ? a0 : 0.0;
        zb[lane + 32] = (lane + 32 < jb) ? a1 : 0.0;
        if (lane < jb) ysm[j0 + lane] = a0;
        if (lane + 32 < jb) ysm[j0 + lane + 32] = a1;
      } else {
        const int t = tid - 32, nt = CC_THREADS - 32;
        if (b > 0) {                      // next diagonal block
          const int p0 = (b - 1) * CC_B;
          for (int e = t; e < CC_B * CC_B; e += nt) {
            const int r = e >> 6, c = e & 63;
            Lnext[r * CC_DL + c] = (c <= r) ? S[(size_t)(p0 + r) * n + p0 + c] : 0.0;
          }
        }
        if (b + 1 < nblk) {               // deferred update of block b+1 on the columns left of chunk b
          const int q0 = (b + 1) * CC_B, qb = min(CC_B, n - q0);
          for (int i = t; i < j0; i += nt) {
            double s0 = 0.0, s1 = 0.0;
            const double* col = S + (size_t)q0 * n + i;
            // 16 loads in flight per thread: one SM has to pull up to 0.5 MB per block out of L2
#pragma unroll 1
            for (int k = 0; k < CC_B; k += 16) {
              double v[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) v[u] = (k + u < qb) ? col[(size_t)(k + u) * n] : 0.0;
#pragma unroll
              for (int u = 0; u < 16; u += 2) { s0 += v[u] * zprev[k + u]; s1 += v[u + 1] * zprev[k + u + 1]; }
            }
            ysm[i] -= s0 + s1;
          }
        }
      }
      __syncthreads();
      if (b > 0) {                        // block b's update of chunk b-1, all threads
        const int i = j0 - CC_B + (tid & 63), kg = tid >> 6;
        double v[16], s0 = 0.0;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) { const int k = kg + 4 * kk; v[kk] = (k < jb) ? S[(size_t)(j0 + k) * n + i] : 0.0; }
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) s0 += v[kk] * zb[kg + 4 * kk];
        red[kg * 64 + (tid & 63)] = s0;
        __syncthreads();
        if (tid < 64) ysm[j0 - CC_B + tid] -= (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
      }
      __syncthreads();
      double* tmp = Lcur; Lcur = Lnext; Lnext = tmp;
    }
    const bool failed = *fail != 0;
    for (int i = tid; i < n; i += CC_THREADS) z[i] = failed ? 0.0 : ysm[i];
    CC_PROF(5);
    return;
  }
  // ---- backward substitution L^T z = y (y = row n of S, in place). Per 64-block, from the last one: CTA 0 solves the
  // diagonal block and publishes z_b; one grid sync; then the update y[i] -= sum_k L[j0+k][i] z_b[k] of the columns to
  // the left is spread over CTAs by 64-column chunk. CTA 0 keeps the chunk it needs next for itself, so the critical
  // path per block is one small update + one 64-step warp solve + one grid sync.
  double* yrow = S + (size_t)n * n;
  double* zb = Pa;            // [64] current z block (shared)
  double* red = Pa + 64;      // [4][64] partial sums
  const int nblk = (n + CC_B - 1) / CC_B;
  for (int b = nblk - 1; b >= 0; --b) {
    const int j0 = b * CC_B, jb = min(CC_B, n - j0);
    if (blockIdx.x == 0) {
      __syncthreads();
      {
        double v[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int e = tid + it * CC_THREADS, r = e >> 6, c = e & 63;
          v[it] = (r < jb && c <= r) ? S[(size_t)(j0 + r) * n + j0 + c] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) { const int e = tid + it * CC_THREADS; Lm[(e >> 6) * CC_DL + (e & 63)] = v[it]; }
      }
      __syncthreads();
      if (tid < jb) invd[tid] = 1.0 / Lm[tid * CC_DL + tid];
      __syncthreads();
      if (warp == 0) {
        double a0 = (lane < jb) ? yrow[j0 + lane] : 0.0, a1 = (lane + 32 < jb) ? yrow[j0 + lane + 32] : 0.0;
        for (int j = jb - 1; j >= 0; --j) {
          const double zj = __shfl_sync(0xffffffffu, (j < 32) ? a0 : a1, j & 31) * invd[j];
          if (lane == (j & 31)) { if (j < 32) a0 = zj; else a1 = zj; }
          if (lane < j) a0 -= zj * Lm[j * CC_DL + lane];
          if (lane + 32 < j) a1 -= zj * Lm[j * CC_DL + lane + 32];
        }
        if (lane < jb) z[j0 + lane] = a0;
        if (lane + 32 < jb) z[j0 + lane + 32] = a1;
        __threadfence();
      }
    }
    grid.sync();
    if (b == 0) break;
    for (int c = b - 1; c >= 0; --c) {
      const int owner = (c == b - 1) ? 0 : 1 + (c % (gridDim.x - 1));
      if ((int)blockIdx.x != owner) continue;
      __syncthreads();
      if (tid < CC_B) zb[tid] = (tid < jb) ? z[j0 + tid] : 0.0;
      __syncthreads();
      const int i = c * CC_B + (tid & 63), kg = tid >> 6;
      double v[16], s0 = 0.0;
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) { const int k = kg + 4 * kk; v[kk] = (k < jb) ? S[(size_t)(j0 + k) * n + i] : 0.0; }
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) s0 += v[kk] * zb[kg + 4 * kk];
      red[kg * 64 + (tid & 63)] = s0;
      __syncthreads();
      if (tid < 64) yrow[c * CC_B + tid] -= (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
    }
    __threadfence();
  }
  if (blockIdx.x == 0 && *fail != 0) {
    __syncthreads();
    for (int i = tid; i < n; i += CC_THREADS) z[i] = 0.0;
  }
  CC_PROF(5);
#undef CC_PROF
}

// ---------------------------------------------------------------------------------------------
// K4: e-block back-substitution z_o = L^-T (y_o - Z_o z_f); candidate x+ = x + s_e * (-z_o); pieces of
// model_cost_change and of the norms. part[o] = {p.y, |p|^2 - sum D^2 z^2 + 2 p.q, sum (x - x+)^2, sum x+^2}
__global__ void __launch_bounds__(256) k_backsubst(const double* __restrict__ Z, const double* __restrict__ Lall,
                                                   const double* __restrict__ zf, const double* __restrict__ scale_e,
                                                   const double* __restrict__ diag_e, const double* __restrict__ radius_p,
                                                   const double* __restrict__ pose, int n_pose, FLayout Lo,
                                                   const unsigned* __restrict__ tile_mask, int mask_words,
                                                   double* __restrict__ pose_cand, double* __restrict__ part) {
  const int n_f = Lo.n_f, ldz = Lo.ldz;
  const double radius = *radius_p;
  // one warp per e-block: q = Z_o z_f by lanes striding the columns (coalesced), xor-shuffle tree (fixed order)
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= n_pose) return;
  const double* Zo = Z + (size_t)o * 6 * ldz;
  double q[6] = {0, 0, 0, 0, 0, 0};
  const unsigned* mk = tile_mask ? tile_mask + (size_t)o * mask_words : nullptr;
  for (int j = lane; j < n_f; j += 32) {
    const int pc = Lo.pcol(j);
    if (mk && !((mk[pc >> 8] >> ((pc >> 3) & 31)) & 1u)) continue;      // structurally zero tile of Z_o
    const double zj = zf[j];
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] += Zo[k * ldz + pc] * zj;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], off);
  }
  if (lane != 0) return;
  double L[36], p[6], zo[6];
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) L[r * 6 + c] = Lall[(size_t)o * 36 + r * 6 + c];
  double py = 0, pp = 0, pq = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double y = Zo[k * ldz + Lo.rhs_col];
    p[k] = y - q[k]; zo[k] = p[k]; py += p[k] * y; pp += p[k] * p[k]; pq += p[k] * q[k];
  }
  chol6_bwd(L, zo);
  double dz = 0, sn = 0, xn = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double d2 = diag_e[(size_t)o * 6 + k] / radius;
    dz += d2 * zo[k] * zo[k];
    const double x = pose[(size_t)o * 6 + k];
    const double delta = scale_e[(size_t)o * 6 + k] * (-zo[k]);
    const double xc = x + delta;
    pose_cand[(size_t)o * 6 + k] = xc;
    sn += (x - xc) * (x - xc); xn += xc * xc;
  }
  part[(size_t)o * 4 + 0] = py; part[(size_t)o * 4 + 1] = pp - dz + 2.0 * pq;
  part[(size_t)o * 4 + 2] = sn; part[(size_t)o * 4 + 3] = xn;
}

// f part of the step (single CTA): candidate f-blocks, z_f.gs_f, z_f^T (s FF s) z_f, sum (x-x+)^2, sum x+^2,
// and at Jacobian evaluation time (zf == nullptr): sum x_f^2 and max |x - (x - g)| only.
// State layout of f: camera c -> [pose 6][intr 9] (whichever exist), then boards.
struct FState { double* cam_pose; double* intr; double* board; };
__device__ __forceinline__ double* f_slot(const FState& s, int j, int n_cam, int Wc, int Wb, bool has_cam) {
  const int ncw = n_cam * Wc;
  if (j < ncw) { const int c = j / Wc, l = j % Wc; return (has_cam && l < 6) ? s.cam_pose + 6 * c + l : s.intr + 9 * c + (l - (has_cam ? 6 : 0)); }
  const int b = (j - ncw) / Wb, l = (j - ncw) % Wb;
  return s.board + 6 * b + l;
}
// hz = (s FF s) z_f, one warp per row (FF is symmetric: row j read contiguously), xor-shuffle tree
__global__ void __launch_bounds__(256) k_ffz(const double* __restrict__ FF, const double* __restrict__ scale_f,
                                             const double* __restrict__ zf, int n_f, double* __restrict__ hz) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= n_f) return;
  double s = 0.0;
  for (int k = lane; k < n_f; k += 32) s += FF[(size_t)j * n_f + k] * scale_f[k] * zf[k];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) hz[j] = s * scale_f[j];
}
__global__ void __launch_bounds__(1024) k_fstep(FState cur, FState cand, const double* __restrict__ zf,
                                                const double* __restrict__ hzv, const double* __restrict__ gf,
                                                const double* __restrict__ scale_f, int n_f, int n_cam, int Wc, int Wb,
                                                int has_cam, double* __restrict__ out /*[4]*/) {
  __shared__ double red[4][1024];
  const int t = threadIdx.x;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int j = t; j < n_f; j += 1024) {
    const double x = *f_slot(cur, j, n_cam, Wc, Wb, has_cam);
    if (zf) {
      const double z = zf[j], sj = scale_f[j];
      const double xc = x + sj * (-z);
      *f_slot(cand, j, n_cam, Wc, Wb, has_cam) = xc;
      a0 += z * sj * gf[j]; a1 += z * hzv[j]; a2 += (x - xc) * (x - xc); a3 += xc * xc;
    } else {   // hzv is F^T F here: a non-finite gradient entry or Jacobian column -> +inf (see k_eblock_grad)
      const double pg = x + (-gf[j]);
      a0 += x * x; a1 = fmax(a1, fabs(x - pg));
      if (!isfinite(gf[j]) || !isfinite(hzv[(size_t)j * n_f + j])) a1 = CUDART_INF;
    }
  }
  red[0][t] = a0; red[1][t] = a1; red[2][t] = a2; red[3][t] = a3;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (t < s) {
      red[0][t] += red[0][t + s]; red[2][t] += red[2][t + s]; red[3][t] += red[3][t + s];
      red[1][t] = zf ? red[1][t] + red[1][t + s] : fmax(red[1][t], red[1][t + s]);
    }
    __syncthreads();
  }
  if (t < 4) out[t] = red[t][0];
}

}  // namespace mcba
