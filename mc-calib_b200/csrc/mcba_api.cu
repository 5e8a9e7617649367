// mcba_api.cu — libmcba: C ABI (include/mcba.h), host packer and the Levenberg-Marquardt driver.
//
// The driver restates ceres-solver 2.2.0's TrustRegionMinimizer + LevenbergMarquardtStrategy with the
// options the reference uses (SPARSE_SCHUR, max_num_iterations, all else default;
// /root/reference/McCalib/src/CameraGroup.cpp:463-468, Object3D.cpp:218-223, Camera.cpp:296-301):
// SURVEY.md Appendix B lists the rules. All arithmetic of the path runs in the CUDA kernels of
// mcba_kernels.cuh; the host only takes the accept/reject/terminate decisions from a handful of scalars.
// There is no CPU fallback: every entry point fails with MCBA_ERR_CUDA when no sm_100 device is usable.
#include "../../include/mcba.h"
#include "mcba_kernels.cuh"
#include "mcba_lm.cuh"
#include "mcba_batched.cuh"
#include "mcba_chol_tiles.cuh"
#include "mcba_syrk.cuh"
#include "mcba_obs_stream.cuh"
#include "mcba_obs_fused.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

using namespace mcba;

// ---- NCCL through dlopen (no link-time dependency; torch's bundled libnccl is reused when loaded) -------
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
    GetUniqueId = (int (*)(nccl_uid_t*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid_t, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
    CommDestroy = (int (*)(nccl_comm_t))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && AllGather && CommDestroy;
  }
};
static NcclApi g_nccl;
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;

// ---- kernel families for timing -------------------------------------------------------------------------
enum { KF_PREP = 0, KF_FUSED, KF_MATERIALIZE, KF_FROM_J, KF_COST, KF_EASSEMBLE, KF_PAIR, KF_EFACTOR, KF_SYRK,
       KF_REDUCED, KF_CHOL, KF_BACKSUBST, KF_REDUCE, KF_NCCL, KF_COUNT };
static const char* kKernelNames[KF_COUNT] = {
    "prep_blocks", "resid_jac_fused", "resid_jac_materialize", "accumulate_from_jacobian", "cost_only",
    "eblock_assemble", "pair_reduce", "eblock_factor", "schur_syrk", "build_reduced", "cholesky", "backsubst",
    "reduce_misc", "nccl"};

struct Handle {
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_b = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // F^T F assembly overlaps the e-block assembly
  // problem (host copies of the small things)
  int stage = 0, n_cam = 0, n_pose = 0, n_board = 0, n_pts = 0, nb = 0;
  int64_t n_obs = 0, npad = 0;
  std::vector<int32_t> h_bobs_cam, h_bobs_pose, h_bobs_board, h_bobs_flags;
  std::vector<int64_t> h_bobs_ptr;
  // static device data
  float *d_u = nullptr, *d_v = nullptr; int32_t* d_pt = nullptr; double* d_pts = nullptr;
  int4* d_bobs = nullptr; int* d_pose_bobs_ptr = nullptr; int* d_bobs_flags = nullptr;
  int32_t* d_cam_model = nullptr; uint8_t *d_cam_is_ref = nullptr, *d_board_is_ref = nullptr;
  // stage-dependent
  FLayout fl{};   // padded column layout of Wt / Z / ZtZ
  int Wc = 0, Wb = 0, n_f = 0, ldz = 0, n_pair = 0, n_chunks = 0, NT = 0, NE = 0, K = 0;
  int* d_pair_bobs = nullptr; PairChunk* d_chunks = nullptr; int* d_pair_chunk_ptr = nullptr;
  double *d_brec = nullptr, *d_brec_alt = nullptr, *brec_target = nullptr, *d_cost_bobs = nullptr, *d_Hoo = nullptr, *d_Wt = nullptr, *d_Z = nullptr, *d_L = nullptr;
  double *d_scale_e = nullptr, *d_diag_e = nullptr, *d_scale_f = nullptr, *d_diag_f = nullptr;
  double *d_FF = nullptr, *d_gf = nullptr, *d_S = nullptr, *d_rhs = nullptr, *d_zf = nullptr, *d_ZtZ = nullptr, *d_hz = nullptr;
  double *d_syrk_partial = nullptr, *d_pair_partial = nullptr, *d_pair_rec = nullptr, *d_part_e = nullptr;
  double *pair_in = nullptr, *pair_out = nullptr;   // where k_pair_final writes / k_ff_assemble reads (d_pair_rec, or the exported buffer)
  double *d_red_scratch = nullptr, *d_scalars = nullptr, *d_gather = nullptr;
  int* d_fail = nullptr;
  bool obs_shared = false;   // d_u / d_v / d_pt belong to the observation cache (mcba_obs_register)
  double *d_jmat = nullptr, *d_resmat = nullptr;   // lazily allocated
  int syrk_nt = 0, syrk_tiles = 0, syrk_nsplit = 0; int64_t syrk_rows_per_split = 0;
  unsigned* d_tile_mask = nullptr; int mask_words = 0; bool syrk_sparse = false, syrk_tma = true;   // k_syrk2: which 8-column tiles of Z an e-block touches
  int chol_grid = 1, n_sm = 148;
  long long* d_chol_prof = nullptr;
  unsigned* d_ct_flags = nullptr; double *d_ct_dinv = nullptr, *d_ct_pbuf = nullptr; unsigned ct_epoch = 0; int ct_capacity = 0; long long* d_ct_prof = nullptr;   // k_chol_tiles
  // batched independent problems (config 5)
  bool batched = false;
  const uint8_t* obs_gate = nullptr;
  int* d_cam_ptr = nullptr; LmState* d_lm = nullptr; uint8_t *d_gate_jac = nullptr, *d_active = nullptr;
  int *d_cam_fail = nullptr, *d_n_active = nullptr; double *d_cam_sc = nullptr, *d_cam_cost = nullptr;
  std::vector<int> h_cam_ptr;
  // parameters: two copies (current / candidate), swapped on acceptance
  double *d_cam_pose[2] = {nullptr, nullptr}, *d_pose[2] = {nullptr, nullptr}, *d_board[2] = {nullptr, nullptr},
         *d_intr[2] = {nullptr, nullptr};
  double *d_prep_cam = nullptr, *d_prep_pose = nullptr, *d_prep_board = nullptr;
  int cur = 0;
  // comm
  int rank = 0, n_ranks = 1; nccl_comm_t comm = nullptr;
  // peer-memory path of the Schur reduction (k_syrk_sum_p2p)
  bool p2p = false; unsigned long long p2p_epoch = 0; int p2p_grid = 0;
  unsigned long long* d_p2p_flags = nullptr; unsigned int* d_p2p_counter = nullptr; int* d_p2p_err = nullptr;
  double* d_p2p_stage = nullptr; bool ZtZ_pooled = false;
  double* peer_stage[P2P_MAX_RANKS] = {nullptr}; double* peer_ZtZ[P2P_MAX_RANKS] = {nullptr}; unsigned long long* peer_flags[P2P_MAX_RANKS] = {nullptr};
  // pair records and phase scalars over the same mechanism: one more exported buffer per rank [pair in | pair out | scalar table x 2]
  double* d_p2p_x = nullptr; double* peer_x[P2P_MAX_RANKS] = {nullptr}; size_t x_pair = 0; unsigned long long pair_epoch = 0, gather_epoch = 0;
  unsigned int* d_p2p_counter2 = nullptr;
  unsigned long long* d_epochs = nullptr;   // [3] launch counters of the three exchanges (device memory: the kernels sit in replayed graphs)
  std::vector<void*> ipc_opened;
  // LM state
  mcba_options opt;
  bool solving = false, have_scale = false, cand_has_jac = false;
  LmState lm;
  double ev_cost = 0, ev_x_norm = 0, ev_grad_max = 0;   // outputs of the last eval_jacobian
  mcba_summary summ;
  std::vector<mcba_trace_row> trace;
  std::vector<double> x0_cam, x0_pose, x0_board, x0_intr;   // start point of the solve in flight (returned on FAILURE)
  int64_t n_obs_global = 0;
  // CUDA graphs of the two per-iteration launch sequences (single-rank handles): key = 2 * cur + brec_par
  bool use_graph = false, graph_env = false, capturing = false; int brec_par = 0;
  int k1_variant = -1;               // fused observation kernel of this handle (launch_obs)
  bool k1_pair_saves = true;         // 16-corner half-warp passes of two board observations < 0.9 x the 32-corner passes of one
  // K1 with in-kernel accumulation of the f x f / f x r part over runs of one (camera, board) pair (k1_acc): the board
  // observations in pair order with everything a warp needs per item, the runs ("chunks") cut at the streams' range
  // boundaries, one partial pair record per chunk for each of the two tile sets
  bool k1_acc = false; int k1_grid = 0, k1_streams = 0, n_k1_chunks = 0;
  int4 *d_k1_rec = nullptr, *d_k1_aux = nullptr; int* d_k1_chunk_ptr = nullptr;
  double *d_k1_partial = nullptr, *d_k1_partial_alt = nullptr;
  cudaGraphExec_t g_step[4] = {nullptr, nullptr, nullptr, nullptr}, g_asm[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t g_step_launches[4] = {0, 0, 0, 0}, g_asm_launches[4] = {0, 0, 0, 0};
  mcba_options graph_opt; bool graph_opt_valid = false;
  double* h_pin = nullptr;      // pinned host: [0] radius, [8 .. 8 + SC_N) the scalars of the last phase
  double* d_radius = nullptr; unsigned* d_ct_epoch = nullptr;
  // timing
  struct Pending { int fam; cudaEvent_t a, b; };
  std::vector<Pending> pending; std::vector<cudaEvent_t> free_events;
  int64_t kf_launches[KF_COUNT] = {0}; double kf_ms[KF_COUNT] = {0};
  double kf_ms_at_begin[KF_COUNT] = {0};   // snapshot taken by mcba_solve_begin: mcba_summary::t_* are per-solve deltas
  int64_t n_launches = 0;   // kernels launched by this handle
  cudaEvent_t tm_a = nullptr, tm_b = nullptr;
};

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { H->err = std::string(#x) + ": " + cudaGetErrorString(e_); return MCBA_ERR_CUDA; } } while (0)
#define CHECK_LAUNCH() CU(cudaGetLastError())

static cudaEvent_t get_event(Handle* H) {
  if (!H->free_events.empty()) { cudaEvent_t e = H->free_events.back(); H->free_events.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
static void flush_timers(Handle* H) {
  if (H->pending.empty()) return;
  cudaStreamSynchronize(H->stream);
  if (H->stream_b) cudaStreamSynchronize(H->stream_b);
  for (auto& p : H->pending) {
    float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
    H->kf_ms[p.fam] += ms; H->kf_launches[p.fam] += 1;
    H->free_events.push_back(p.a); H->free_events.push_back(p.b);
  }
  H->pending.clear();
}
struct Timed {   // RAII: brackets the launches of one kernel family with CUDA events on the stream they are issued to
  Handle* H; int fam; cudaEvent_t a; cudaStream_t st;
  bool on;
  Timed(Handle* h, int f, cudaStream_t s = nullptr) : H(h), fam(f), st(s ? s : h->stream), on(!h->use_graph) { if (on) { a = get_event(H); cudaEventRecord(a, st); } }
  ~Timed() { if (!on) return; cudaEvent_t b = get_event(H); cudaEventRecord(b, st); H->pending.push_back({fam, a, b}); if (H->pending.size() > 4096) flush_timers(H); }
};

// Device memory comes from the device's default stream-ordered pool, with the release threshold lifted: what
// mcba_destroy gives back stays in the pool, so a create / solve / destroy cycle (one per refine* stage and camera
// group in the reference's workflow) does not pay cudaMalloc / cudaFree (each a device-wide synchronisation) again.
static void keep_pool(int device) {
  static std::mutex m; static bool done[64] = {false};
  std::lock_guard<std::mutex> g(m);
  if (device < 0 || device >= 64 || done[device]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
  done[device] = true;
}
template <typename T> static int dalloc(Handle* H, T** p, size_t n) {
  if (*p) { cudaFreeAsync(*p, H->stream); *p = nullptr; }
  if (n == 0) n = 1;
  CU(cudaMallocAsync((void**)p, n * sizeof(T), H->stream));
  return 0;
}
template <typename T> static void dfree(Handle* H, T** p) { if (*p) { cudaFreeAsync(*p, H->stream); *p = nullptr; } }

struct StageDims { bool cam, intr, board; int K, NT, NE, WF, NCOMP; };
static StageDims stage_dims(int stage) {
  switch (stage) {
    case 1: return {false, true, false, Cols<1>::K, Cols<1>::NT, PairRec<1>::NE, Cols<1>::WF, Cols<1>::NCOMP};
    case 2: return {false, false, true, Cols<2>::K, Cols<2>::NT, PairRec<2>::NE, Cols<2>::WF, Cols<2>::NCOMP};
    case 3: return {true, false, false, Cols<3>::K, Cols<3>::NT, PairRec<3>::NE, Cols<3>::WF, Cols<3>::NCOMP};
    case 4: return {true, false, true, Cols<4>::K, Cols<4>::NT, PairRec<4>::NE, Cols<4>::WF, Cols<4>::NCOMP};
    case 5: return {true, true, true, Cols<5>::K, Cols<5>::NT, PairRec<5>::NE, Cols<5>::WF, Cols<5>::NCOMP};
  }
  return {false, false, false, 0, 0, 0, 0, 0};
}

#define DISPATCH_STAGE(stage, ...)                    \
  switch (stage) {                                    \
    case 1: { constexpr int ST = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int ST = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int ST = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int ST = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int ST = 5; __VA_ARGS__; } break; \
  }

static int setup_batched(Handle* H) {
  H->stage = 1; H->K = Cols<1>::K; H->NT = Cols<1>::NT; H->NE = PairRec<1>::NE;
  H->Wc = 9; H->Wb = 0; H->n_f = 9 * H->n_cam; H->ldz = 16;
  const size_t np = H->n_pose, nc = H->n_cam;
  H->h_cam_ptr.assign(nc + 1, 0);
  for (int b = 0; b < H->nb; ++b) H->h_cam_ptr[H->h_bobs_cam[b] + 1] = b + 1;
  for (size_t c = 0; c < nc; ++c) H->h_cam_ptr[c + 1] = std::max(H->h_cam_ptr[c + 1], H->h_cam_ptr[c]);
  if (dalloc(H, &H->d_cam_ptr, nc + 1)) return MCBA_ERR_CUDA;
  CU(cudaMemcpy(H->d_cam_ptr, H->h_cam_ptr.data(), (nc + 1) * sizeof(int), cudaMemcpyHostToDevice));
  if (dalloc(H, &H->d_brec, (size_t)H->nb * B_REC) || dalloc(H, &H->d_cost_bobs, (size_t)H->nb)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_L, np * 36) || dalloc(H, &H->d_Z, np * 60) || dalloc(H, &H->d_part_e, np * 4)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_scale_e, np * 6) || dalloc(H, &H->d_diag_e, np * 6)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_scale_f, nc * 9) || dalloc(H, &H->d_diag_f, nc * 9) || dalloc(H, &H->d_zf, nc * 9)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_FF, nc * 81) || dalloc(H, &H->d_gf, nc * 9)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_cam_sc, nc * BC_N) || dalloc(H, &H->d_cam_cost, nc)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_lm, nc) || dalloc(H, &H->d_gate_jac, nc) || dalloc(H, &H->d_active, nc)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_cam_fail, nc) || dalloc(H, &H->d_n_active, (size_t)1)) return MCBA_ERR_CUDA;
  CU(cudaMemset(H->d_cam_fail, 0, nc * sizeof(int)));
  const int sm = OBS_WARPS * (CTX_SIZE + Cols<1>::NC * TS) * 8;
  CU(cudaFuncSetAttribute(k_observations<1, MODE_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  return MCBA_OK;
}

static void drop_graphs(Handle* H);
// Fused observation kernel of a stage (launch_obs): the handle's variant, or the default per stage.
// (variant + 32 = that variant with the in-kernel pair accumulation, + 16 = without it, whatever MCBA_K1_ACC says: test hook)
// Default: the pair kernel (3) when pairing saves Jacobian passes - corner counts whose last 32-corner pass is at most
// half full (48-corner boards: 3 passes of 16 + 16 per pair instead of 4; 16-corner boards: 1 instead of 2) - else the
// one-observation kernel (0): with 63-corner boards both need 2 passes per board observation and the pair kernel only
// adds its bookkeeping (config 3, stage 5: 1.81 vs 1.70 ms per iteration).
static int k1_variant_of(const Handle* H, int stage) { (void)stage; return H->k1_variant >= 0 ? (H->k1_variant & 15) : (H->k1_pair_saves ? 3 : 0); }
// Grid of the two kernels that can accumulate pair records (k_observations_fused / k_observations_pair): 3 CTAs per SM
// unless the shared memory of the stage allows fewer; the streams (warps / half-warps) of this grid own fixed ranges.
static int k1_grid_of(const Handle* H, int stage, int k1v, int nb) {
  int sm = 0;
  DISPATCH_STAGE(stage, { sm = k1v == 3 ? PairSmem<ST>::BYTES : FusedSmem<ST>::BYTES; });
  const int per_sm = std::min(3, std::max(1, (227 * 1024) / (sm + 1024)));
  const int units = k1v == 3 ? (nb + 1) / 2 : nb;
  return std::max(1, std::min((units + OBS_WARPS - 1) / OBS_WARPS, H->n_sm * per_sm));
}
static bool k1_acc_on(const Handle* H) { return H->k1_acc && !H->opt.materialize_jacobian; }
static int setup_stage(Handle* H, int stage) {
  if (H->batched) return setup_batched(H);
  drop_graphs(H);   // every buffer the captured launches point to is reallocated below
  const StageDims sd = stage_dims(stage);
  if (sd.K == 0) { H->err = "stage must be 1..5"; return MCBA_ERR_INVALID; }
  if (sd.board && H->n_board <= 0) { H->err = "stage needs board blocks but n_board == 0"; return MCBA_ERR_INVALID; }
  H->stage = stage; H->K = sd.K; H->NT = sd.NT; H->NE = sd.NE;
  H->Wc = (sd.cam ? 6 : 0) + (sd.intr ? 9 : 0);
  H->Wb = sd.board ? 6 : 0;
  H->n_f = H->n_cam * H->Wc + (sd.board ? H->n_board * H->Wb : 0);
  {
    // MCBA_SYRK_SPARSE=1: block-sparse Schur SYRK on a layout padded to the DMMA tiles (measured slower than the dense
    // kernel on the dome: DESIGN.md); default: dense, unpadded
    static const bool sparse_env = []() { const char* e = getenv("MCBA_SYRK_SPARSE"); return e && atoi(e) != 0; }();
    H->syrk_sparse = sparse_env;
    const int nbd = sd.board ? H->n_board : 0;
    const int Wcp = H->syrk_sparse ? ((H->Wc + 7) / 8) * 8 : H->Wc, Wbp = H->syrk_sparse ? ((H->Wb + 7) / 8) * 8 : H->Wb;
    const int ncm = H->Wc > 0 ? H->n_cam : 0;     // stage 2 has no per-camera f-block
    const int rhs_col = ncm * Wcp + nbd * Wbp;
    H->ldz = H->syrk_sparse ? rhs_col + 8 : ((rhs_col + 1 + 7) / 8) * 8;
    H->fl = FLayout{ncm, std::max(1, H->Wc), Wcp, nbd, std::max(1, H->Wb), Wbp, H->n_f, H->ldz, rhs_col};
  }
  // pair lists: pair = cam * nbp + board (board = 0 when the stage has no board block)
  const int nbp = sd.board ? H->n_board : 1;
  H->n_pair = H->n_cam * nbp;
  std::vector<int> cnt(H->n_pair + 1, 0);
  for (int b = 0; b < H->nb; ++b) cnt[H->h_bobs_cam[b] * nbp + (sd.board ? H->h_bobs_board[b] : 0) + 1]++;
  for (int p = 0; p < H->n_pair; ++p) cnt[p + 1] += cnt[p];
  std::vector<int> pair_bobs(std::max(1, H->nb)), fill(cnt.begin(), cnt.end() - 1);
  for (int b = 0; b < H->nb; ++b) pair_bobs[fill[H->h_bobs_cam[b] * nbp + (sd.board ? H->h_bobs_board[b] : 0)]++] = b;
  constexpr int CHUNK = 64;
  std::vector<PairChunk> chunks; std::vector<int> pcp(H->n_pair + 1, 0);
  for (int p = 0; p < H->n_pair; ++p) {
    pcp[p] = (int)chunks.size();
    for (int s = cnt[p]; s < cnt[p + 1]; s += CHUNK) chunks.push_back({p, s, std::min(cnt[p + 1], s + CHUNK), s == cnt[p]});
  }
  pcp[H->n_pair] = (int)chunks.size();
  H->n_chunks = (int)chunks.size();
  if (dalloc(H, &H->d_pair_bobs, pair_bobs.size())) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_chunks, chunks.size())) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_pair_chunk_ptr, pcp.size())) return MCBA_ERR_CUDA;
  CU(cudaMemcpy(H->d_pair_bobs, pair_bobs.data(), pair_bobs.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (!chunks.empty()) CU(cudaMemcpy(H->d_chunks, chunks.data(), chunks.size() * sizeof(PairChunk), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(H->d_pair_chunk_ptr, pcp.data(), pcp.size() * sizeof(int), cudaMemcpyHostToDevice));
  {
    // K1 accumulating the f x f / f x r part itself (variants 0 and 3): items in pair order, streams own contiguous
    // ranges of them, a chunk = a run of one pair inside one range; k_pair_final sums the chunks of a pair in order
    const int k1v = k1_variant_of(H, stage);
    // Off by default - measured slower: the runs of a pair are spread over all frames, so a warp's consecutive items
    // share nothing but the camera and the board (pose records, pixels and records scattered over HBM instead of
    // streaming), and K1 loses 0.05 ms where the assembly gains 0.04. MCBA_K1_ACC=1 turns it on.
    static const bool acc_env = []() { const char* e = getenv("MCBA_K1_ACC"); return e && atoi(e) != 0; }();
    const bool want = H->k1_variant >= 0 && (H->k1_variant & 32) ? true : (H->k1_variant >= 0 && (H->k1_variant & 16) ? false : acc_env);
    H->k1_acc = want && (k1v == 0 || k1v == 3) && H->nb > 0;
    if (H->k1_acc) {
      H->k1_grid = k1_grid_of(H, stage, k1v, H->nb);
      H->k1_streams = H->k1_grid * OBS_WARPS * (k1v == 3 ? 2 : 1);
      std::vector<int4> rec(H->nb + 1), aux(H->nb + 1);
      std::vector<int> k1pcp(H->n_pair + 1, 0);
      int n_ch = 0, s_cur = 0;
      long long s_hi = (long long)H->nb * 1 / H->k1_streams;   // end of stream 0's range
      int prev_pair = -1;
      for (int i = 0; i < H->nb; ++i) {
        while (i >= s_hi && s_cur + 1 < H->k1_streams) { ++s_cur; s_hi = (long long)H->nb * (s_cur + 1) / H->k1_streams; prev_pair = -1; }
        const int b = pair_bobs[i];
        const int pr = H->h_bobs_cam[b] * nbp + (sd.board ? H->h_bobs_board[b] : 0);
        if (pr != prev_pair) { ++n_ch; prev_pair = pr; }
        const int bf = H->h_bobs_flags[b];
        rec[i] = make_int4((int)H->h_bobs_ptr[b], H->h_bobs_cam[b], H->h_bobs_pose[b], H->h_bobs_board[b]);
        aux[i] = make_int4((int)H->h_bobs_ptr[b + 1], bf, b, n_ch - 1);
        k1pcp[pr + 1] = n_ch;
      }
      rec[H->nb] = make_int4(0, 0, 0, 0); aux[H->nb] = make_int4(0, 0, 0, -1);
      for (int p = 0; p < H->n_pair; ++p) k1pcp[p + 1] = std::max(k1pcp[p + 1], k1pcp[p]);   // pairs without board observations
      H->n_k1_chunks = n_ch;
      if (dalloc(H, &H->d_k1_rec, rec.size()) || dalloc(H, &H->d_k1_aux, aux.size()) || dalloc(H, &H->d_k1_chunk_ptr, k1pcp.size())) return MCBA_ERR_CUDA;
      if (dalloc(H, &H->d_k1_partial, (size_t)n_ch * H->NE) || dalloc(H, &H->d_k1_partial_alt, (size_t)n_ch * H->NE)) return MCBA_ERR_CUDA;
      CU(cudaMemcpy(H->d_k1_rec, rec.data(), rec.size() * sizeof(int4), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(H->d_k1_aux, aux.data(), aux.size() * sizeof(int4), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(H->d_k1_chunk_ptr, k1pcp.data(), k1pcp.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
  }
  const size_t np = H->n_pose, nf = H->n_f, ldz = H->ldz;
  if (dalloc(H, &H->d_brec, (size_t)H->nb * sd.NCOMP) || dalloc(H, &H->d_brec_alt, (size_t)H->nb * sd.NCOMP)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_cost_bobs, (size_t)H->nb)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_Hoo, np * 36) || dalloc(H, &H->d_L, np * 36)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_Wt, np * 6 * ldz) || dalloc(H, &H->d_Z, np * 6 * ldz)) return MCBA_ERR_CUDA;
  CU(cudaMemsetAsync(H->d_Z, 0, np * 6 * ldz * sizeof(double), H->stream));   // padding columns and unobserved blocks are never written
  CU(cudaMemsetAsync(H->d_Wt, 0, np * 6 * ldz * sizeof(double), H->stream));
  if (dalloc(H, &H->d_scale_e, np * 6) || dalloc(H, &H->d_diag_e, np * 6)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_scale_f, nf) || dalloc(H, &H->d_diag_f, nf)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_FF, nf * nf) || dalloc(H, &H->d_gf, nf) || dalloc(H, &H->d_S, (nf + 1) * nf)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_hz, nf) || dalloc(H, &H->d_rhs, nf) || dalloc(H, &H->d_zf, nf) || dalloc(H, &H->d_ZtZ, ldz * ldz)) return MCBA_ERR_CUDA;
  CU(cudaMemset(H->d_ZtZ, 0, ldz * ldz * sizeof(double)));   // the lower triangle is never written
  H->syrk_nt = (H->ldz + SY_T - 1) / SY_T;
  H->syrk_tiles = H->syrk_nt * (H->syrk_nt + 1) / 2;
  const int64_t n_rows = 6 * (int64_t)H->n_pose;
  static const bool syrk_cpasync_env = []() { const char* e = getenv("MCBA_SYRK_CPASYNC"); return e && atoi(e) != 0; }();
  H->syrk_tma = !syrk_cpasync_env || H->syrk_sparse;         // bulk-copy (TMA) pipeline by default; MCBA_SYRK_CPASYNC=1: round-1 kernel
  int syrk_per_sm = 1;
  CU(cudaFuncSetAttribute(k_syrk, cudaFuncAttributeMaxDynamicSharedMemorySize, SY_SMEM));
  CU(cudaFuncSetAttribute(k_syrk2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SY2_SMEM));
  CU(cudaFuncSetAttribute(k_syrk2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SY2_SMEM));
  if (H->syrk_tma) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&syrk_per_sm, k_syrk2<false>, SY2_THREADS, SY2_SMEM);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&syrk_per_sm, k_syrk, SY_THREADS, SY_SMEM);
  const int kq = H->syrk_tma ? SY2_K : SY_K;                 // rows per chunk: splits are whole chunks (whole e-blocks for k_syrk2)
  int nsplit = std::max(1, (2 * std::max(1, syrk_per_sm) * H->n_sm) / H->syrk_tiles);   // about two full waves of equal CTAs
  nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(nsplit, (n_rows + 63) / 64));
  H->syrk_rows_per_split = (((n_rows + nsplit - 1) / nsplit + kq - 1) / kq) * kq;
  if (H->syrk_rows_per_split == 0) H->syrk_rows_per_split = kq;
  H->syrk_nsplit = (int)std::max<int64_t>(1, (n_rows + H->syrk_rows_per_split - 1) / H->syrk_rows_per_split);
  if (H->syrk_sparse) {
    // activity masks: e-block o touches the tiles of the cameras / boards its board observations name, and the rhs block
    const FLayout& L = H->fl;
    H->mask_words = (H->ldz / 8 + 31) / 32;
    std::vector<unsigned> mask((size_t)std::max(1, H->n_pose) * H->mask_words, 0u);
    auto set_range = [&](int o, int col0, int width) { for (int t = col0 / 8; t < (col0 + width) / 8; ++t) mask[(size_t)o * H->mask_words + (t >> 5)] |= 1u << (t & 31); };
    for (int o = 0; o < H->n_pose; ++o) set_range(o, L.rhs_col, 8);
    for (int b = 0; b < H->nb; ++b) {
      const int o = H->h_bobs_pose[b];
      if (L.n_cam > 0) set_range(o, H->h_bobs_cam[b] * L.Wcp, L.Wcp);
      if (L.n_board > 0) set_range(o, L.n_cam * L.Wcp + H->h_bobs_board[b] * L.Wbp, L.Wbp);
    }
    if (dalloc(H, &H->d_tile_mask, mask.size())) return MCBA_ERR_CUDA;
    CU(cudaMemcpyAsync(H->d_tile_mask, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice, H->stream));
    CU(cudaStreamSynchronize(H->stream));
  } else {
    dfree(H, &H->d_tile_mask); H->mask_words = 0;
  }
  if (dalloc(H, &H->d_syrk_partial, (size_t)H->syrk_nsplit * H->syrk_tiles * SY_T * SY_T)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_pair_partial, (size_t)H->n_chunks * H->NE)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_pair_rec, (size_t)H->n_pair * H->NE)) return MCBA_ERR_CUDA;
  H->pair_in = H->d_pair_rec; H->pair_out = H->d_pair_rec;
  if (dalloc(H, &H->d_part_e, np * 4)) return MCBA_ERR_CUDA;
  dfree(H, &H->d_jmat); dfree(H, &H->d_resmat);
  H->have_scale = false;
  // opt in to > 48 KB dynamic shared memory
  DISPATCH_STAGE(stage, {
    const int sm = OBS_WARPS * (CTX_SIZE + Cols<ST>::NC * TS) * 8;
    CU(cudaFuncSetAttribute(k_observations<ST, MODE_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaFuncSetAttribute(k_observations<ST, MODE_FROM_J>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    CU(cudaFuncSetAttribute(k_observations_stream<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, OBS_WARPS * (2 * CTX_SIZE + Cols<ST>::NC * TS) * 8));
    CU(cudaFuncSetAttribute(k_observations_fused<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem<ST>::BYTES));
    CU(cudaFuncSetAttribute(k_observations_fused<ST, OBS_PASS, 3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem<ST>::BYTES));
    CU(cudaFuncSetAttribute(k_observations_pair<ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem<ST>::BYTES));
    CU(cudaFuncSetAttribute(k_observations_pair<ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem<ST>::BYTES));
    CU(cudaFuncSetAttribute(k_observations_fused<ST, 24, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedSmem<ST, 24, false>::BYTES));
    CU(cudaFuncSetAttribute(k_eblock_assemble<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (H->ldz + EA_PAD) * 8));
  });
  return MCBA_OK;
}

// out_dev[m] = column-wise sum (cols < m_sum) / max (others) of in[n][m]; deterministic two-stage
static int reduce_rows(Handle* H, const double* in, int64_t n, int m, int m_sum, double* out_dev) {
  Timed t(H, KF_REDUCE);
  int nblk = (int)std::min<int64_t>(512, std::max<int64_t>(1, (n + 2047) / 2048));
  const int64_t rpb = (n + nblk - 1) / std::max(1, nblk);
  if (nblk == 1) {
    ++H->n_launches;
    k_reduce_rows<<<1, RED_T, 0, H->stream>>>(in, n, m, m_sum, std::max<int64_t>(n, 1), out_dev);
  } else {
    ++H->n_launches;
    k_reduce_rows<<<nblk, RED_T, 0, H->stream>>>(in, n, m, m_sum, rpb, H->d_red_scratch);
    ++H->n_launches;
    k_reduce_rows<<<1, RED_T, 0, H->stream>>>(H->d_red_scratch, nblk, m, m_sum, nblk, out_dev);
  }
  CHECK_LAUNCH();
  return MCBA_OK;
}

static ObsArgs obs_args(Handle* H, int which, double huber_a) {
  ObsArgs a;
  a.u = H->d_u; a.v = H->d_v; a.pt = H->d_pt; a.pts = H->d_pts; a.bobs = H->d_bobs;
  a.prep_cam = H->d_prep_cam; a.prep_pose = H->d_prep_pose; a.prep_board = H->d_prep_board; a.intr = H->d_intr[which];
  a.cam_model = H->d_cam_model; a.cam_is_ref = H->d_cam_is_ref; a.board_is_ref = H->d_board_is_ref;
  a.huber_a = huber_a; a.nb = H->nb; a.brec = H->brec_target ? H->brec_target : H->d_brec; a.cost_bobs = H->d_cost_bobs;
  a.jmat = H->d_jmat; a.resmat = H->d_resmat; a.npad = H->npad; a.gate = H->obs_gate; a.bobs_flags = H->d_bobs_flags;
  a.srec = H->d_k1_rec; a.saux = H->d_k1_aux; a.n_streams = H->k1_streams;
  a.partial = (H->brec_target && H->brec_target == H->d_brec_alt) ? H->d_k1_partial_alt : H->d_k1_partial;
  return a;
}

static int launch_prep(Handle* H, int which) {
  Timed t(H, KF_PREP);
  const int n = H->n_pose + H->n_cam + H->n_board;
  ++H->n_launches;
  k_prep_blocks<<<(n + 127) / 128, 128, 0, H->stream>>>(H->d_cam_pose[which], H->n_cam, H->d_pose[which], H->n_pose,
                                                        H->d_board[which], H->n_board, H->d_prep_cam, H->d_prep_pose,
                                                        H->d_prep_board);
  CHECK_LAUNCH();
  return MCBA_OK;
}

static int g_k1_variant = -1;   // mcba_debug_set_k1: overrides MCBA_K1 for handles created afterwards
static int grid_for(const Handle* H, int nb, int ctas_per_sm) { return std::max(1, std::min((nb + OBS_WARPS - 1) / OBS_WARPS, H->n_sm * ctas_per_sm)); }

template <int MODE> static int launch_obs(Handle* H, int which, double huber_a) {
  if (H->nb == 0) return MCBA_OK;
  const int fam = MODE == MODE_FUSED ? KF_FUSED : MODE == MODE_MATERIALIZE ? KF_MATERIALIZE : MODE == MODE_FROM_J ? KF_FROM_J : KF_COST;
  Timed t(H, fam);
  ObsArgs a = obs_args(H, which, huber_a);
  // Variant of the fused kernel, fixed per handle at mcba_create (MCBA_K1 / mcba_debug_set_k1):
  //   0 = k_observations_fused: one board observation per warp, prefetched copy-free context, group-pair cover of the
  //       triangle (7 DMMAs per k-step instead of 10 at stage 5);
  //   3 = k_observations_pair: the same with two board observations per warp (one per half-warp, 16 corners per pass);
  //   1 = the first version k_observations<S, MODE_FUSED> (what gated / batched handles run unless the variant is 3);
  //   2 = the streaming experiment (tails of consecutive board observations share a warp pass; slower than version 1);
  //   4 = k_observations_fused with 24-corner passes, one context slot and 4 CTAs (16 warps) per SM at 128 registers:
  //       0.91 ms at stage 5 (spills), 0.60 ms at stage 4 (= variant 0): the kernel is bound by the pipe the DMMA and DFMA
  //       instructions share (63 % busy) and the shared-memory wavefronts (56 %), not by the number of resident warps;
  //  -1 = default: 3 at every stage (at stage 5 the two kernels are within 1 % of each other: 0.726 vs 0.734 ms).
  //  Variants 0-3 write bit-identical records. MCBA_K1_ACC=1 (or variant + 32): variants 0 and 3 accumulate the f x f /
  //  f x r part over runs of one (camera, board) pair and write partial pair records instead (setup_stage).
  const int k1v = k1_variant_of(H, H->stage);
  if (MODE == MODE_FUSED && k1v == 2 && !a.gate) {
    DISPATCH_STAGE(H->stage, {
      const int sm = OBS_WARPS * (2 * CTX_SIZE + Cols<ST>::NC * TS) * 8;
      const int per_sm = std::min(3, std::max(1, (227 * 1024) / (sm + 1024)));
      const int grid = std::max(1, std::min((H->nb + 2 * OBS_WARPS - 1) / (2 * OBS_WARPS), H->n_sm * per_sm));
      ++H->n_launches;
      k_observations_stream<ST><<<grid, OBS_WARPS * 32, sm, H->stream>>>(a);
    });
    CHECK_LAUNCH();
    return MCBA_OK;
  }
  if (MODE == MODE_FUSED && k1v == 3) {   // the pair kernel honours ObsArgs::gate (batched problems)
    const bool acc = k1_acc_on(H) && !a.gate;
    DISPATCH_STAGE(H->stage, {
      const int sm = PairSmem<ST>::BYTES;
      ++H->n_launches;
      if (acc) k_observations_pair<ST, true><<<H->k1_grid, OBS_WARPS * 32, sm, H->stream>>>(a);
      else k_observations_pair<ST, false><<<k1_grid_of(H, H->stage, 3, H->nb), OBS_WARPS * 32, sm, H->stream>>>(a);
    });
    CHECK_LAUNCH();
    return MCBA_OK;
  }
  if (MODE == MODE_FUSED && k1v == 4 && !a.gate) {   // experiment: 24-corner passes, one context slot, 4 CTAs (16 warps) per SM
    DISPATCH_STAGE(H->stage, {
      const int sm = FusedSmem<ST, 24, false>::BYTES;
      const int grid = grid_for(H, H->nb, std::min(4, std::max(1, (227 * 1024) / (sm + 1024))));
      ++H->n_launches;
      k_observations_fused<ST, 24, 4, false><<<grid, OBS_WARPS * 32, sm, H->stream>>>(a);
    });
    CHECK_LAUNCH();
    return MCBA_OK;
  }
  if (MODE == MODE_FUSED && k1v == 0 && !a.gate) {
    const bool acc = k1_acc_on(H);
    DISPATCH_STAGE(H->stage, {
      const int sm = FusedSmem<ST>::BYTES;
      ++H->n_launches;
      if (acc) k_observations_fused<ST, OBS_PASS, 3, true, true><<<H->k1_grid, OBS_WARPS * 32, sm, H->stream>>>(a);
      else k_observations_fused<ST><<<k1_grid_of(H, H->stage, 0, H->nb), OBS_WARPS * 32, sm, H->stream>>>(a);
    });
    CHECK_LAUNCH();
    return MCBA_OK;
  }
  DISPATCH_STAGE(H->stage, {
    constexpr bool ACC = (MODE == MODE_FUSED || MODE == MODE_FROM_J);
    const int sm = OBS_WARPS * (CTX_SIZE + (ACC ? Cols<ST>::NC * TS : 0)) * 8;
    const int grid = grid_for(H, H->nb, ACC ? std::min(4, std::max(1, (227 * 1024) / (sm + 1024))) : 8);
    ++H->n_launches;
    k_observations<ST, MODE><<<grid, OBS_WARPS * 32, sm, H->stream>>>(a);
  });
  CHECK_LAUNCH();
  return MCBA_OK;
}

static int ensure_jmat(Handle* H) {
  if (H->d_jmat) return MCBA_OK;
  if (dalloc(H, &H->d_jmat, (size_t)2 * (H->K + 1) * H->npad) || dalloc(H, &H->d_resmat, (size_t)2)) return MCBA_ERR_CUDA;
  return MCBA_OK;
}

// The scalar gather over peer memory in two halves: the launch part (capturable: bump the device-side launch counter,
// k_gather_p2p, copy BOTH halves of the double-buffered table to pinned host memory) and the host part (synchronise,
// pick the half by the parity of the mirrored launch count).
constexpr int PIN_GATHER = 32;
static int gather_p2p_launch(Handle* H, const double* d_local, int m) {
  Timed t(H, KF_NCCL);
  P2PGatherArgs ga;
  const size_t tab = 2 * H->x_pair;
  for (int r = 0; r < P2P_MAX_RANKS; ++r) { ga.table[r] = H->peer_x[r] ? H->peer_x[r] + tab : nullptr; ga.flags[r] = H->peer_flags[r]; }
  ga.local = d_local; ga.m = m; ga.rank = H->rank; ga.n_ranks = H->n_ranks; ga.epoch_p = H->d_epochs + 2; ga.err = H->d_p2p_err;
  H->n_launches += 2;
  k_bump_epoch64<<<1, 1, 0, H->stream>>>(H->d_epochs + 2);
  k_gather_p2p<<<1, 256, 0, H->stream>>>(ga);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(H->h_pin + PIN_GATHER, H->d_p2p_x + tab, (size_t)2 * P2P_MAX_RANKS * P2P_GATHER_W * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  return MCBA_OK;
}
static int gather_p2p_read(Handle* H, int m, std::vector<double>& host) {
  ++H->gather_epoch;                       // mirrors the device-side counter: one gather executed per call
  CU(cudaStreamSynchronize(H->stream));
  host.assign((size_t)m * H->n_ranks, 0.0);
  const double* tabp = H->h_pin + PIN_GATHER + (H->gather_epoch & 1ull) * P2P_MAX_RANKS * P2P_GATHER_W;
  for (int r = 0; r < H->n_ranks; ++r) for (int i = 0; i < m; ++i) host[(size_t)r * m + i] = tabp[(size_t)r * P2P_GATHER_W + i];
  return MCBA_OK;
}

// gather [m] doubles from every rank (rank-major) to the host; world 1: plain D2H
static int gather_scalars(Handle* H, const double* d_local, int m, std::vector<double>& host) {
  host.assign((size_t)m * H->n_ranks, 0.0);
  if (H->n_ranks == 1) {
    CU(cudaMemcpyAsync(host.data(), d_local, m * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  } else if (H->p2p && m <= P2P_GATHER_W) {
    int rc;
    if ((rc = gather_p2p_launch(H, d_local, m))) return rc;
    return gather_p2p_read(H, m, host);
  } else {
    { Timed t(H, KF_NCCL);
      const int rc = g_nccl.AllGather(d_local, H->d_gather, m, NCCL_DOUBLE, H->comm, H->stream);
      if (rc != 0) { H->err = std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return MCBA_ERR_NCCL; } }
    CU(cudaMemcpyAsync(host.data(), H->d_gather, (size_t)m * H->n_ranks * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  }
  CU(cudaStreamSynchronize(H->stream));
  return MCBA_OK;
}
static int allreduce_sum(Handle* H, double* d_buf, size_t n) {
  if (H->n_ranks == 1) return MCBA_OK;
  Timed t(H, KF_NCCL);
  const int rc = g_nccl.AllReduce(d_buf, d_buf, n, NCCL_DOUBLE, NCCL_SUM, H->comm, H->stream);
  if (rc != 0) { H->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return MCBA_ERR_NCCL; }
  return MCBA_OK;
}

// d_scalars slots
enum { SC_COST = 0, SC_XE2 = 1, SC_GMAX_E = 2, SC_F0 = 3 /*..6*/, SC_PE0 = 7 /*..10*/, SC_FAIL = 11, SC_N = 12 };

static FState fstate(Handle* H, int which) { return FState{H->d_cam_pose[which], H->d_intr[which], H->d_board[which]}; }

// ProgramEvaluator::Evaluate with Jacobian, in two halves. eval_observations: residual + Jacobian + per-bobs J^T J
// tiles at parameter copy `which`, written to `target` (the current tile set or the alternate one), and the cost.
// eval_assemble: everything SchurEliminator and the trust-region loop need from the tiles of the *current* set.
// The split lets the candidate point be evaluated with its Jacobian straight away (the cost comes with it): when
// the step is accepted - the usual case - the tiles are simply swapped in instead of running a cost-only pass
// followed by a second, full pass at the same point as Ceres does; a rejected step leaves the current set untouched.
static int eval_observations(Handle* H, int which, double* target) {
  int rc;
  if ((rc = launch_prep(H, which))) return rc;
  H->brec_target = target;
  if (H->opt.materialize_jacobian) {
    if (!(rc = ensure_jmat(H)) && !(rc = launch_obs<MODE_MATERIALIZE>(H, which, H->opt.huber_a)))
      rc = launch_obs<MODE_FROM_J>(H, which, H->opt.huber_a);
  } else {
    rc = launch_obs<MODE_FUSED>(H, which, H->opt.huber_a);
  }
  H->brec_target = nullptr;
  if (rc) return rc;
  return reduce_rows(H, H->d_cost_bobs, H->nb, 1, 1, H->d_scalars + SC_COST);
}

static void drop_graphs(Handle* H) {
  for (int k = 0; k < 4; ++k) {
    if (H->g_step[k]) { cudaGraphExecDestroy(H->g_step[k]); H->g_step[k] = nullptr; }
    if (H->g_asm[k]) { cudaGraphExecDestroy(H->g_asm[k]); H->g_asm[k] = nullptr; }
  }
}
// Run `launch` once through stream capture and keep the instantiated graph in *slot; later calls replay it.
// Falls back to direct launches (use_graph = false) if the capture is refused.
template <class F> static int run_graphed(Handle* H, cudaGraphExec_t* slot, int64_t* n_nodes, F launch) {
  if (!*slot) {
    const int64_t l0 = H->n_launches;
    H->capturing = true;
    cudaError_t e = cudaStreamBeginCapture(H->stream, cudaStreamCaptureModeThreadLocal);
    int rc = MCBA_OK;
    if (e == cudaSuccess) rc = launch();
    cudaGraph_t g = nullptr;
    if (e == cudaSuccess) e = cudaStreamEndCapture(H->stream, &g);
    H->capturing = false;
    if (e == cudaSuccess && rc == MCBA_OK && g) e = cudaGraphInstantiate(slot, g, 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess || rc != MCBA_OK || !*slot) {     // not capturable here: direct launches from now on
      cudaGetLastError();
      *slot = nullptr; H->use_graph = false; drop_graphs(H);
      H->n_launches = l0;
      return launch();
    }
    *n_nodes = H->n_launches - l0;
    H->n_launches = l0;
  }
  H->n_launches += *n_nodes;
  CU(cudaGraphLaunch(*slot, H->stream));
  return MCBA_OK;
}

static int eval_assemble_launch(Handle* H, bool first) {
  int rc;
  const StageDims sd = stage_dims(H->stage);
  // Two independent readers of the per-board-observation records: the e-block assembly (main stream) and the
  // F^T F / F^T r assembly (pair reduction -> [allreduce] -> dense F^T F; stream_b). They meet before k_fstep.
  // (single GPU only for now: with ranks the pair-record allreduce would move to stream_b as well, which has not been
  // measured on a multi-GPU box yet; there everything stays on the main stream, in this order)
  const bool ov = H->stream_b != nullptr && (H->n_ranks == 1 || H->p2p);
  cudaStream_t sb = ov ? H->stream_b : H->stream;
  if (ov) { CU(cudaEventRecord(H->ev_fork, H->stream)); CU(cudaStreamWaitEvent(H->stream_b, H->ev_fork, 0)); }
  {
    Timed t(H, KF_PAIR, sb);
    const bool acc = k1_acc_on(H);   // the fused kernel already left one partial pair record per chunk of its streams
    if (!acc && H->n_chunks > 0) {
      ++H->n_launches;
      DISPATCH_STAGE(H->stage, (k_pair_reduce<ST><<<H->n_chunks, 256, 0, sb>>>(H->d_chunks, H->d_pair_bobs, H->d_brec, H->d_pair_partial)));
      CHECK_LAUNCH();
    }
    const int n = H->n_pair * H->NE;
    ++H->n_launches;
    DISPATCH_STAGE(H->stage, (k_pair_final<ST><<<(n + 255) / 256, 256, 0, sb>>>(acc ? H->d_k1_chunk_ptr : H->d_pair_chunk_ptr, H->n_pair,
                                                                              acc ? H->d_k1_partial : H->d_pair_partial, H->pair_in)));
    CHECK_LAUNCH();
  }
  if (H->n_ranks > 1 && H->p2p) {   // sum of the pair records across ranks over peer memory (k_allreduce_p2p)
    Timed t(H, KF_NCCL, sb);
    P2PVecArgs va;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) { va.in[r] = H->peer_x[r]; va.out[r] = H->peer_x[r] ? H->peer_x[r] + H->x_pair : nullptr; va.flags[r] = H->peer_flags[r]; }
    k_bump_epoch64<<<1, 1, 0, sb>>>(H->d_epochs + 1);
    va.n = H->n_pair * H->NE; va.rank = H->rank; va.n_ranks = H->n_ranks; va.epoch_p = H->d_epochs + 1; va.counter = H->d_p2p_counter2; va.err = H->d_p2p_err;
    ++H->n_launches;   // no grid-wide barrier inside (the CTAs only wait for other ranks' flags): an ordinary launch
    k_allreduce_p2p<<<std::min(H->n_sm, std::max(1, (va.n / H->n_ranks + 255) / 256)), 256, 0, sb>>>(va);
    CHECK_LAUNCH();
  } else if (H->n_ranks > 1) {   // NCCL fallback
    Timed t(H, KF_NCCL, sb);
    const int nrc = g_nccl.AllReduce(H->d_pair_rec, H->d_pair_rec, (size_t)H->n_pair * H->NE, NCCL_DOUBLE, NCCL_SUM, H->comm, sb);
    if (nrc != 0) { H->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?"); return MCBA_ERR_NCCL; }
  }
  {
    Timed t(H, KF_PAIR, sb);
    const int n = H->n_f * H->n_f + H->n_f;
    ++H->n_launches;
    DISPATCH_STAGE(H->stage, (k_ff_assemble<ST><<<(n + 255) / 256, 256, 0, sb>>>(H->pair_out, H->n_cam, H->n_board, H->Wc, H->Wb, H->n_f, H->d_FF, H->d_gf)));
    CHECK_LAUNCH();
  }
  if (ov) CU(cudaEventRecord(H->ev_join, H->stream_b));
  if (H->n_pose > 0) {
    Timed t(H, KF_EASSEMBLE);
    ++H->n_launches;
    DISPATCH_STAGE(H->stage, (k_eblock_assemble<ST><<<H->n_pose, 192, 6 * (H->ldz + EA_PAD) * 8, H->stream>>>(
        H->d_bobs, H->d_pose_bobs_ptr, H->d_brec, H->fl, H->d_tile_mask, H->mask_words, H->d_Hoo, H->d_Wt)));
    CHECK_LAUNCH();
  }
  {
    Timed t(H, KF_REDUCE);
    if (H->n_pose > 0) {
      ++H->n_launches;
      k_eblock_grad<<<(H->n_pose + 127) / 128, 128, 0, H->stream>>>(H->d_pose[H->cur], H->d_Wt, H->d_Hoo, H->n_pose, H->fl.rhs_col, H->ldz, H->d_part_e);
      CHECK_LAUNCH();
    }
    if (ov) CU(cudaStreamWaitEvent(H->stream, H->ev_join, 0));   // F^T F and F^T r are ready
    ++H->n_launches;
    k_fstep<<<1, 1024, 0, H->stream>>>(fstate(H, H->cur), fstate(H, H->cur), nullptr, H->d_FF, H->d_gf, H->d_scale_f, H->n_f, H->n_cam, H->Wc, H->Wb, sd.cam ? 1 : 0, H->d_scalars + SC_F0);
    CHECK_LAUNCH();
  }
  if ((rc = reduce_rows(H, H->d_part_e, H->n_pose, 2, 1, H->d_scalars + SC_XE2))) return rc;
  if (first || !H->have_scale) {
    Timed t(H, KF_REDUCE);
    const int n = H->n_pose * 6 + H->n_f;
    ++H->n_launches;
    k_scale_diag<<<(n + 255) / 256, 256, 0, H->stream>>>(H->d_Hoo, H->n_pose, H->d_FF, H->n_f, 1, H->opt.jacobi_scaling, H->opt.min_lm_diagonal, H->opt.max_lm_diagonal, H->d_scale_e, H->d_scale_f, H->d_diag_e, H->d_diag_f);
    CHECK_LAUNCH();
  }
  if (H->n_ranks == 1) CU(cudaMemcpyAsync(H->h_pin + 8, H->d_scalars, SC_N * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  else if (H->p2p && (rc = gather_p2p_launch(H, H->d_scalars, SC_N))) return rc;
  return MCBA_OK;
}

static int eval_assemble(Handle* H, bool first) {
  int rc;
  const bool with_scale = first || !H->have_scale;
  if (H->use_graph && !with_scale) {
    const int key = 2 * H->cur + H->brec_par;
    rc = run_graphed(H, &H->g_asm[key], &H->g_asm_launches[key], [&]() { return eval_assemble_launch(H, false); });
  } else {
    rc = eval_assemble_launch(H, with_scale);
  }
  if (rc) return rc;
  if (with_scale) { H->have_scale = true; H->lm.reuse_diagonal = 1; }   // the diagonal was just computed from this Jacobian
  else H->lm.reuse_diagonal = 0;
  std::vector<double> hs;
  if (H->n_ranks == 1) {
    CU(cudaStreamSynchronize(H->stream));
    hs.assign(H->h_pin + 8, H->h_pin + 8 + SC_N);
  } else if (H->p2p) { if ((rc = gather_p2p_read(H, SC_N, hs))) return rc; }
  else if ((rc = gather_scalars(H, H->d_scalars, SC_N, hs))) return rc;
  double cost = 0, xe2 = 0, gmax = 0;
  for (int r = 0; r < H->n_ranks; ++r) {
    const double* s = &hs[(size_t)r * SC_N];
    cost += s[SC_COST];
    if (H->n_pose > 0 || H->n_ranks > 1) { xe2 += s[SC_XE2]; gmax = std::max(gmax, s[SC_GMAX_E]); }
  }
  if (H->n_pose == 0 && H->n_ranks == 1) { xe2 = 0; }
  gmax = std::max(gmax, hs[SC_F0 + 1]);
  H->ev_cost = cost;
  H->ev_x_norm = std::sqrt(xe2 + hs[SC_F0 + 0]);
  H->ev_grad_max = gmax;
  // Evaluator validity as in Ceres: cost, gradient and every Jacobian column must be finite
  if (!std::isfinite(cost) || !std::isfinite(gmax)) return MCBA_ERR_NUMERIC;
  return MCBA_OK;
}

static int eval_jacobian(Handle* H, bool first) {
  const int rc = eval_observations(H, H->cur, H->d_brec);
  return rc ? rc : eval_assemble(H, first);
}

static int cholesky_reduced(Handle* H) {
  Timed t(H, KF_CHOL);
  int n = H->n_f;
  // dataflow kernel (one CTA per 64x64 tile of the factor) whenever all tiles fit on the device at once
  static const bool tiles_off_env = []() { const char* e = getenv("MCBA_CHOL_TILES"); return e && atoi(e) == 0; }();
  bool tiles_off = tiles_off_env;
  if (const char* once = getenv("MCBA_CHOL_TILES_ONCE")) tiles_off = atoi(once) == 0;   // mcba_debug_cholesky
  if (!tiles_off && n > 0 && ct_num_tiles(n) <= H->ct_capacity && (n + CT_B - 1) / CT_B <= CT_MAXNT) {
    static const bool ct_prof = getenv("MCBA_CHOL_PROF") != nullptr;
    if (ct_prof && !H->d_ct_prof) { CU(cudaMalloc((void**)&H->d_ct_prof, (size_t)CT_MAXT * 16 * sizeof(long long))); }
    if (ct_prof) CU(cudaMemsetAsync(H->d_ct_prof, 0, (size_t)CT_MAXT * 16 * sizeof(long long), H->stream));
    ++H->n_launches;
    k_bump_epoch<<<1, 1, 0, H->stream>>>(H->d_ct_epoch);
    CholTilesArgs a{H->d_S, n, H->d_zf, H->d_fail, H->d_ct_flags, H->d_ct_epoch, H->d_ct_dinv, H->d_ct_pbuf, H->d_ct_prof};
    void* args[] = {&a};
    ++H->n_launches;
    CU(cudaLaunchCooperativeKernel((void*)k_chol_tiles, dim3(ct_num_tiles(n)), dim3(CT_THREADS), args, CT_SMEM, H->stream));
    return MCBA_OK;
  }
  if (n > 2 * CC_B * CC_LD) { H->err = "reduced system too large for the cooperative Cholesky (n_f > 8704)"; return MCBA_ERR_INVALID; }
  double* S = H->d_S; double* z = H->d_zf; int* fail = H->d_fail;
  static const bool want_prof = getenv("MCBA_CHOL_PROF") != nullptr;
  if (want_prof && !H->d_chol_prof) { CU(cudaMalloc((void**)&H->d_chol_prof, 16 * sizeof(long long))); CU(cudaMemset(H->d_chol_prof, 0, 16 * sizeof(long long))); }
  long long* prof = H->d_chol_prof;
  // back-substitution: 1 = CTA 0 alone, pipelined, no grid sync (its update traffic, n^2/2 doubles through one SM, only
  // pays off for small systems: 0.532 vs 0.539 ms at n = 1020); 0 = updates spread over the CTAs, one grid sync per block
  static const int bs_env = []() { const char* e = getenv("MCBA_CHOL_BACKSOLVE"); return e ? atoi(e) : -1; }();
  int mode = bs_env >= 0 ? bs_env : (n <= 1536 ? 1 : 0);
  void* args[] = {&S, &n, &z, &fail, &prof, &mode};
  ++H->n_launches;
  CU(cudaLaunchCooperativeKernel((void*)k_chol_coop, dim3(H->chol_grid), dim3(CC_THREADS), args, CC_SMEM, H->stream));
  return MCBA_OK;
}

// The launch sequence of one trust-region step (no host synchronisation inside): what a CUDA graph captures.
// Everything that changes from one iteration to the next reaches the kernels through device memory (the radius) or
// through the graph key (which parameter copy / tile set is current).
static int compute_step_launch(Handle* H) {
  int rc;
  const StageDims sd = stage_dims(H->stage);
  CU(cudaMemcpyAsync(H->d_radius, H->h_pin, sizeof(double), cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemsetAsync(H->d_fail, 0, sizeof(int), H->stream));
  if (H->n_pose > 0) {
    Timed t(H, KF_EFACTOR);
    ++H->n_launches;
    k_eblock_factor<<<H->n_pose, 128, 0, H->stream>>>(H->d_Hoo, H->d_Wt, H->d_scale_e, H->d_scale_f, H->d_diag_e, H->d_radius, H->fl, H->d_tile_mask, H->mask_words, H->d_L, H->d_Z, H->d_fail);
    CHECK_LAUNCH();
  }
  {
    Timed t(H, KF_SYRK);
    dim3 grid(H->syrk_tiles, H->syrk_nsplit);
    ++H->n_launches;
    if (H->syrk_sparse)
      k_syrk2<true><<<grid, SY2_THREADS, SY2_SMEM, H->stream>>>(H->d_Z, 6 * (int64_t)H->n_pose, H->ldz, H->syrk_nt, H->syrk_rows_per_split, H->d_tile_mask, H->mask_words, H->d_syrk_partial);
    else if (H->syrk_tma)
      k_syrk2<false><<<grid, SY2_THREADS, SY2_SMEM, H->stream>>>(H->d_Z, 6 * (int64_t)H->n_pose, H->ldz, H->syrk_nt, H->syrk_rows_per_split, nullptr, 0, H->d_syrk_partial);
    else
      k_syrk<<<grid, SY_THREADS, SY_SMEM, H->stream>>>(H->d_Z, 6 * (int64_t)H->n_pose, H->ldz, H->syrk_nt, H->syrk_rows_per_split, H->d_syrk_partial);
    ++H->n_launches;
    if (H->p2p) {   // split-K sum + cross-GPU reduction in one kernel over NVLink peer memory
      P2PArgs pa;
      for (int r = 0; r < P2P_MAX_RANKS; ++r) { pa.stage[r] = H->peer_stage[r]; pa.ZtZ[r] = H->peer_ZtZ[r]; pa.flags[r] = H->peer_flags[r]; }
      pa.partial = H->d_syrk_partial; pa.nsplit = H->syrk_nsplit;
      k_bump_epoch64<<<1, 1, 0, H->stream>>>(H->d_epochs);
      pa.rank = H->rank; pa.n_ranks = H->n_ranks; pa.epoch_p = H->d_epochs; pa.n_tiles = H->syrk_tiles; pa.nt = H->syrk_nt; pa.ldz = H->ldz;
      pa.counter = H->d_p2p_counter; pa.err = H->d_p2p_err;
      // the CTAs of this kernel wait for one another (and for the peers' kernels): a cooperative launch makes the
      // driver guarantee co-residency of the whole grid (it fails with an error instead of deadlocking)
      void* pargs[] = {&pa};
      CU(cudaLaunchCooperativeKernel((void*)k_syrk_sum_p2p, dim3(H->p2p_grid), dim3(256), pargs, 0, H->stream));
    } else {
      k_syrk_sum<<<dim3(H->syrk_tiles, 8), 256, 0, H->stream>>>(H->d_syrk_partial, H->syrk_tiles, H->syrk_nt, H->syrk_nsplit, H->ldz, H->d_ZtZ);
    }
    CHECK_LAUNCH();
  }
  if (!H->p2p && (rc = allreduce_sum(H, H->d_ZtZ, (size_t)H->ldz * H->ldz))) return rc;
  {
    Timed t(H, KF_REDUCED);
    const int n = H->n_f * H->n_f + H->n_f;
    ++H->n_launches;
    k_build_reduced<<<(n + 255) / 256, 256, 0, H->stream>>>(H->d_FF, H->d_gf, H->d_ZtZ, H->d_scale_f, H->d_diag_f, H->d_radius, H->fl, H->d_S, H->d_rhs);
    CHECK_LAUNCH();
  }
  if ((rc = cholesky_reduced(H))) return rc;
  const int cand = 1 - H->cur;
  {
    Timed t(H, KF_BACKSUBST);
    if (H->n_pose > 0) {
      ++H->n_launches;
      k_backsubst<<<(H->n_pose + 7) / 8, 256, 0, H->stream>>>(H->d_Z, H->d_L, H->d_zf, H->d_scale_e, H->d_diag_e, H->d_radius, H->d_pose[H->cur], H->n_pose, H->fl, H->d_tile_mask, H->mask_words, H->d_pose[cand], H->d_part_e);
    }
    ++H->n_launches;
    k_ffz<<<(H->n_f + 7) / 8, 256, 0, H->stream>>>(H->d_FF, H->d_scale_f, H->d_zf, H->n_f, H->d_hz);
    ++H->n_launches;
    k_fstep<<<1, 1024, 0, H->stream>>>(fstate(H, H->cur), fstate(H, cand), H->d_zf, H->d_hz, H->d_gf, H->d_scale_f, H->n_f, H->n_cam, H->Wc, H->Wb, sd.cam ? 1 : 0, H->d_scalars + SC_F0);
    CHECK_LAUNCH();
  }
  if ((rc = reduce_rows(H, H->d_part_e, H->n_pose, 4, 4, H->d_scalars + SC_PE0))) return rc;
  // candidate point: residual + Jacobian tiles into the alternate set (the cost comes with them), see eval_observations
  if ((rc = eval_observations(H, cand, H->d_brec_alt))) return rc;
  ++H->n_launches;
  k_flag_to_scalar<<<1, 1, 0, H->stream>>>(H->d_fail, H->p2p ? H->d_p2p_err : nullptr, H->d_scalars + SC_FAIL);
  CHECK_LAUNCH();
  if (H->n_ranks == 1) CU(cudaMemcpyAsync(H->h_pin + 8, H->d_scalars, SC_N * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  else if (H->p2p && (rc = gather_p2p_launch(H, H->d_scalars, SC_N))) return rc;
  return MCBA_OK;
}

// LevenbergMarquardtStrategy::ComputeStep + SchurComplementSolver + candidate point + candidate cost.
// Outputs (host): step validity ingredients.
static int compute_step(Handle* H, LmStep* out) {
  int rc;
  if (!H->lm.reuse_diagonal) {
    Timed t(H, KF_REDUCE);
    const int n = H->n_pose * 6 + H->n_f;
    ++H->n_launches;
    k_scale_diag<<<(n + 255) / 256, 256, 0, H->stream>>>(H->d_Hoo, H->n_pose, H->d_FF, H->n_f, 0, H->opt.jacobi_scaling, H->opt.min_lm_diagonal, H->opt.max_lm_diagonal, H->d_scale_e, H->d_scale_f, H->d_diag_e, H->d_diag_f);
    CHECK_LAUNCH();
  }
  H->lm.reuse_diagonal = 1;
  H->h_pin[0] = H->lm.radius;
  if (H->use_graph) {
    const int key = 2 * H->cur + H->brec_par;
    rc = run_graphed(H, &H->g_step[key], &H->g_step_launches[key], [&]() { return compute_step_launch(H); });
  } else {
    rc = compute_step_launch(H);
  }
  if (rc) return rc;
  H->cand_has_jac = true;
  std::vector<double> hs;
  if (H->n_ranks == 1) {
    CU(cudaStreamSynchronize(H->stream));
    hs.assign(H->h_pin + 8, H->h_pin + 8 + SC_N);
  } else if (H->p2p) { if ((rc = gather_p2p_read(H, SC_N, hs))) return rc; }
  else if ((rc = gather_scalars(H, H->d_scalars, SC_N, hs))) return rc;
  double py = 0, he = 0, sn = 0, xn = 0, cost = 0, h_fail = 0;
  for (int r = 0; r < H->n_ranks; ++r) {
    const double* s = &hs[(size_t)r * SC_N];
    cost += s[SC_COST]; h_fail += s[SC_FAIL];
    if (H->n_pose > 0 || H->n_ranks > 1) { py += s[SC_PE0]; he += s[SC_PE0 + 1]; sn += s[SC_PE0 + 2]; xn += s[SC_PE0 + 3]; }
  }
  if (h_fail >= P2P_ERR_UNIT) { H->err = "peer-memory Schur reduction timed out waiting for another rank"; return MCBA_ERR_NCCL; }
  const double* f = &hs[SC_F0];
  const double zg = py + f[0], zhz = he + f[1];
  out->mcc = zg - 0.5 * zhz;
  out->step_norm = std::sqrt(sn + f[2]);
  out->cand_norm = std::sqrt(xn + f[3]);
  out->cand_cost = std::isfinite(cost) ? cost : std::numeric_limits<double>::max();
  out->solver_ok = ((h_fail == 0.0) && std::isfinite(out->mcc) && std::isfinite(out->step_norm)) ? 1 : 0;
  return MCBA_OK;
}

static int upload_params(Handle* H, const mcba_params* p) {
  const StageDims sd = stage_dims(H->stage);
  if (!p || !p->pose || !p->intrinsics || (sd.cam && !p->cam_pose) || (sd.board && !p->board_pose)) { H->err = "missing parameter array for this stage"; return MCBA_ERR_INVALID; }
  for (int w = 0; w < 2; ++w) {
    CU(cudaMemcpyAsync(H->d_pose[w], p->pose, (size_t)H->n_pose * 6 * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    CU(cudaMemcpyAsync(H->d_intr[w], p->intrinsics, (size_t)H->n_cam * 9 * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    if (p->cam_pose) CU(cudaMemcpyAsync(H->d_cam_pose[w], p->cam_pose, (size_t)H->n_cam * 6 * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    else CU(cudaMemsetAsync(H->d_cam_pose[w], 0, (size_t)H->n_cam * 6 * sizeof(double), H->stream));
    if (p->board_pose && H->n_board > 0) CU(cudaMemcpyAsync(H->d_board[w], p->board_pose, (size_t)H->n_board * 6 * sizeof(double), cudaMemcpyHostToDevice, H->stream));
  }
  H->cur = 0;
  return MCBA_OK;
}
static int download_params(Handle* H, mcba_params* p) {
  const StageDims sd = stage_dims(H->stage);
  const int w = H->cur;
  CU(cudaMemcpyAsync(p->pose, H->d_pose[w], (size_t)H->n_pose * 6 * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  if (sd.cam) CU(cudaMemcpyAsync(p->cam_pose, H->d_cam_pose[w], (size_t)H->n_cam * 6 * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  if (sd.board) CU(cudaMemcpyAsync(p->board_pose, H->d_board[w], (size_t)H->n_board * 6 * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  if (sd.intr) CU(cudaMemcpyAsync(p->intrinsics, H->d_intr[w], (size_t)H->n_cam * 9 * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  return MCBA_OK;
}

// Peer-memory setup for k_syrk_sum_p2p: exchange CUDA IPC handles of the staging buffer, ZtZ and the flag array
// (through the NCCL communicator), open the peers' handles and keep the pointer tables. Collective: every rank calls it
// (mcba_comm_init, and mcba_set_stage after the buffers were reallocated). Falls back to the NCCL allreduce
// (p2p = false) if anything is refused on any rank; MCBA_P2P=0 disables it.
struct IpcRecord { cudaIpcMemHandle_t h[4]; long long off[4]; int ok; int pad; };
// Buffers whose CUDA IPC handles go to other ranks are never cudaFree'd while the process lives: freeing an exported
// allocation before every importer has closed its mapping is undefined behaviour, and mcba_destroy is not a
// collective. They are recycled through this process-wide pool instead (one allocation of >= 2 MB each, so that an IPC
// handle covers exactly one buffer).
struct ExportPool {
  struct Item { void* p; size_t bytes; int device; bool used; };
  std::mutex m; std::vector<Item> items;
  void* acquire(int device, size_t bytes) {
    bytes = std::max<size_t>(bytes, (size_t)2 << 20);
    std::lock_guard<std::mutex> g(m);
    for (auto& it : items)
      if (!it.used && it.device == device && it.bytes >= bytes && it.bytes <= 2 * bytes) { it.used = true; return it.p; }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    items.push_back({p, bytes, device, true});
    return p;
  }
  void release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(m);
    for (auto& it : items) if (it.p == p) it.used = false;
  }
};
static ExportPool g_export_pool;
static int close_p2p(Handle* H) {
  for (void* p : H->ipc_opened) cudaIpcCloseMemHandle(p);
  H->ipc_opened.clear(); H->p2p = false;
  return MCBA_OK;
}
// give the exported buffers of this handle back to the pool (after every rank has unmapped them, or at destruction,
// when no kernel of any rank can touch them any more: k_syrk_sum_p2p only returns once all peers are done with them)
static void release_exported(Handle* H) {
  g_export_pool.release(H->d_p2p_stage); H->d_p2p_stage = nullptr;
  g_export_pool.release(H->d_p2p_flags); H->d_p2p_flags = nullptr;
  g_export_pool.release(H->d_p2p_x); H->d_p2p_x = nullptr; H->pair_in = H->d_pair_rec; H->pair_out = H->d_pair_rec;
  if (H->ZtZ_pooled) { g_export_pool.release(H->d_ZtZ); H->d_ZtZ = nullptr; H->ZtZ_pooled = false; }
}
static int setup_p2p(Handle* H) {
  close_p2p(H);
  if (H->n_ranks < 2 || H->n_ranks > P2P_MAX_RANKS || H->batched) return MCBA_OK;
  const char* env = getenv("MCBA_P2P");
  if (env && atoi(env) == 0) return MCBA_OK;
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn get_range = []() -> range_fn { void* l = dlopen("libcuda.so.1", RTLD_NOW); return l ? (range_fn)dlsym(l, "cuMemGetAddressRange_v2") : nullptr; }();
  const int N = H->n_ranks;
  if (!H->d_p2p_counter) {
    if (dalloc(H, &H->d_p2p_counter, (size_t)2) || dalloc(H, &H->d_p2p_counter2, (size_t)2) || dalloc(H, &H->d_p2p_err, (size_t)1) || dalloc(H, &H->d_epochs, (size_t)4)) return MCBA_ERR_CUDA;
    CU(cudaMemset(H->d_p2p_counter, 0, 2 * sizeof(unsigned int))); CU(cudaMemset(H->d_p2p_counter2, 0, 2 * sizeof(unsigned int))); CU(cudaMemset(H->d_p2p_err, 0, sizeof(int)));
  }
  // exported buffers come from the pool and are cleared BEFORE their handles are published (a peer may start writing
  // flags as soon as it has opened them); the epochs restart with them
  g_export_pool.release(H->d_p2p_stage); g_export_pool.release(H->d_p2p_flags); g_export_pool.release(H->d_p2p_x);
  H->x_pair = (((size_t)H->n_pair * H->NE + 31) / 32) * 32;
  const size_t x_doubles = 2 * H->x_pair + (size_t)2 * P2P_MAX_RANKS * P2P_GATHER_W;
  H->d_p2p_x = (double*)g_export_pool.acquire(H->device, x_doubles * sizeof(double));
  const size_t zz_bytes = (size_t)H->ldz * H->ldz * sizeof(double);
  H->d_p2p_stage = (double*)g_export_pool.acquire(H->device, (size_t)H->syrk_tiles * SY_T * SY_T * sizeof(double));
  H->d_p2p_flags = (unsigned long long*)g_export_pool.acquire(H->device, (size_t)P2P_FLAG_SLOTS * P2P_MAX_RANKS * sizeof(unsigned long long));
  if (!H->ZtZ_pooled) {
    double* zz = (double*)g_export_pool.acquire(H->device, zz_bytes);
    if (zz) { dfree(H, &H->d_ZtZ); H->d_ZtZ = zz; H->ZtZ_pooled = true; }
  }
  if (!H->d_p2p_stage || !H->d_p2p_flags || !H->d_p2p_x || !H->ZtZ_pooled) { H->err = "cudaMalloc(exported buffers)"; return MCBA_ERR_CUDA; }
  CU(cudaMemset(H->d_ZtZ, 0, zz_bytes));   // the lower triangle is never written
  CU(cudaMemset(H->d_p2p_flags, 0, (size_t)P2P_FLAG_SLOTS * P2P_MAX_RANKS * sizeof(unsigned long long)));
  CU(cudaMemset(H->d_p2p_x, 0, x_doubles * sizeof(double)));
  CU(cudaDeviceSynchronize());
  H->p2p_epoch = 0; H->pair_epoch = 0; H->gather_epoch = 0;
  CU(cudaMemset(H->d_epochs, 0, 4 * sizeof(unsigned long long)));
  IpcRecord mine; std::memset(&mine, 0, sizeof(mine));
  void* ptrs[4] = {H->d_p2p_stage, H->d_ZtZ, H->d_p2p_flags, H->d_p2p_x};
  mine.ok = get_range ? 1 : 0;
  for (int k = 0; k < 4 && mine.ok; ++k) {
    unsigned long long base = 0; size_t size = 0;
    if (get_range(&base, &size, (unsigned long long)ptrs[k]) != 0) { mine.ok = 0; break; }
    mine.off[k] = (long long)((unsigned long long)ptrs[k] - base);
    if (cudaIpcGetMemHandle(&mine.h[k], (void*)base) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
  }
  // allgather the records (as bytes) through NCCL
  char *d_send = nullptr, *d_recv = nullptr;
  if (dalloc(H, &d_send, sizeof(IpcRecord)) || dalloc(H, &d_recv, sizeof(IpcRecord) * (size_t)N)) return MCBA_ERR_CUDA;
  CU(cudaMemcpy(d_send, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  if (g_nccl.AllGather(d_send, d_recv, sizeof(IpcRecord), /*ncclChar*/ 0, H->comm, H->stream) != 0) { cudaFreeAsync(d_send, H->stream); cudaFreeAsync(d_recv, H->stream); H->err = "ncclAllGather(ipc handles)"; return MCBA_ERR_NCCL; }
  std::vector<IpcRecord> all(N);
  CU(cudaMemcpyAsync(all.data(), d_recv, sizeof(IpcRecord) * (size_t)N, cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  cudaFreeAsync(d_send, H->stream); cudaFreeAsync(d_recv, H->stream);
  bool ok = true;
  for (int r = 0; r < N; ++r) ok = ok && all[r].ok;
  for (int r = 0; r < P2P_MAX_RANKS; ++r) { H->peer_stage[r] = nullptr; H->peer_ZtZ[r] = nullptr; H->peer_flags[r] = nullptr; H->peer_x[r] = nullptr; }
  for (int r = 0; r < N && ok; ++r) {
    if (r == H->rank) { H->peer_stage[r] = H->d_p2p_stage; H->peer_ZtZ[r] = H->d_ZtZ; H->peer_flags[r] = H->d_p2p_flags; H->peer_x[r] = H->d_p2p_x; continue; }
    void* opened[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4 && ok; ++k) {
      bool dup = false;   // two buffers may live in the same underlying allocation
      for (int q = 0; q < k; ++q) if (std::memcmp(&all[r].h[q], &all[r].h[k], sizeof(cudaIpcMemHandle_t)) == 0) { opened[k] = opened[q]; dup = true; }
      if (dup) continue;
      if (cudaIpcOpenMemHandle(&opened[k], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      H->ipc_opened.push_back(opened[k]);
    }
    if (!ok) break;
    H->peer_stage[r] = (double*)((char*)opened[0] + all[r].off[0]);
    H->peer_ZtZ[r] = (double*)((char*)opened[1] + all[r].off[1]);
    H->peer_flags[r] = (unsigned long long*)((char*)opened[2] + all[r].off[2]);
    H->peer_x[r] = (double*)((char*)opened[3] + all[r].off[3]);
  }
  // every rank must take the same path: agree through a sum over ranks of the local verdict
  double verdict = ok ? 1.0 : 0.0;
  CU(cudaMemcpy(H->d_scalars, &verdict, sizeof(double), cudaMemcpyHostToDevice));
  if (g_nccl.AllReduce(H->d_scalars, H->d_scalars, 1, NCCL_DOUBLE, NCCL_SUM, H->comm, H->stream) != 0) { H->err = "ncclAllReduce(p2p verdict)"; return MCBA_ERR_NCCL; }
  CU(cudaMemcpyAsync(&verdict, H->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  if (verdict < N - 0.5) { close_p2p(H); return MCBA_OK; }
  { int per_sm = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_syrk_sum_p2p, 256, 0));
    H->p2p_grid = std::min(2, std::max(1, per_sm)) * H->n_sm; }
  H->p2p = true;
  H->pair_in = H->d_p2p_x; H->pair_out = H->d_p2p_x + H->x_pair;   // the pair records live in the exported buffer from now on
  return MCBA_OK;
}

// ---- device-resident observations shared between calls (mcba_obs_register) ----------------------------------------
// The pose initialisation (mcba_pnp_ransac) and the bundle adjustment that follows (mcba_create) read the same u / v /
// pt_idx arrays. A caller that registers them pays the H2D copy (12 B per corner: 74 MB for a 6.18 M-corner shard) once;
// every later call whose u pointer, length and device match a registered entry uses the resident copy.
struct ObsEntry { const float* u; const float* v; const int32_t* pt; int64_t n; int device; float* d_u; float* d_v; int32_t* d_pt; int users; };
static std::mutex g_obs_mutex;
static std::vector<ObsEntry> g_obs;
namespace mcba {
bool obs_cache_acquire(const float* u, const float* v, const int32_t* pt, int64_t n, int device, float** d_u, float** d_v, int32_t** d_pt) {
  std::lock_guard<std::mutex> g(g_obs_mutex);
  for (auto& e : g_obs)
    if (e.u == u && e.v == v && e.pt == pt && e.n == n && e.device == device) { ++e.users; *d_u = e.d_u; *d_v = e.d_v; *d_pt = e.d_pt; return true; }
  return false;
}
void obs_cache_release(const float* d_u) {
  std::lock_guard<std::mutex> g(g_obs_mutex);
  for (auto& e : g_obs) if (e.d_u == d_u && e.users > 0) --e.users;
}
}  // namespace mcba

// pt_idx range check on the device (replaces a single-threaded 6 M-element host loop in mcba_create)
__global__ void k_validate_pt(const int32_t* __restrict__ pt, int64_t n, int n_pts, int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool b = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const int p = pt[i]; b = b || p < 0 || p >= n_pts; }
  if (b) *bad = 1;
}

extern "C" {

int mcba_obs_register(const float* u, const float* v, const int32_t* pt_idx, int64_t n_obs, int device) {
  if (!u || !v || !pt_idx || n_obs <= 0) return MCBA_ERR_INVALID;
  { float *a, *b; int32_t* c;
    if (mcba::obs_cache_acquire(u, v, pt_idx, n_obs, device, &a, &b, &c)) { mcba::obs_cache_release(a); return MCBA_OK; } }
  if (cudaSetDevice(device) != cudaSuccess) return MCBA_ERR_CUDA;
  const int64_t npad = ((n_obs + 31) / 32) * 32;
  ObsEntry e{u, v, pt_idx, n_obs, device, nullptr, nullptr, nullptr, 0};
  if (cudaMalloc((void**)&e.d_u, npad * 4) != cudaSuccess || cudaMalloc((void**)&e.d_v, npad * 4) != cudaSuccess ||
      cudaMalloc((void**)&e.d_pt, npad * 4) != cudaSuccess ||
      cudaMemcpy(e.d_u, u, n_obs * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(e.d_v, v, n_obs * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(e.d_pt, pt_idx, n_obs * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError(); cudaFree(e.d_u); cudaFree(e.d_v); cudaFree(e.d_pt); return MCBA_ERR_CUDA;
  }
  std::lock_guard<std::mutex> g(g_obs_mutex);
  g_obs.push_back(e);
  return MCBA_OK;
}
int mcba_obs_release(const float* u, int device) {
  std::lock_guard<std::mutex> g(g_obs_mutex);
  for (size_t i = 0; i < g_obs.size(); ++i)
    if (g_obs[i].u == u && g_obs[i].device == device) {
      if (g_obs[i].users > 0) return MCBA_ERR_INVALID;   // a live handle still reads it
      cudaSetDevice(device); cudaFree(g_obs[i].d_u); cudaFree(g_obs[i].d_v); cudaFree(g_obs[i].d_pt);
      g_obs.erase(g_obs.begin() + i);
      return MCBA_OK;
    }
  return MCBA_ERR_INVALID;
}

/* 0: one rank; 1: NCCL allreduce of the reduced system; 2: fused split-K sum + reduction over NVLink peer memory */
int mcba_schur_reduce_mode(void* h) { Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID; return H->n_ranks <= 1 ? 0 : (H->p2p ? 2 : 1); }

void mcba_default_options(mcba_options* o) {
  o->max_num_iterations = 50;   // Ceres default; the reference overrides it with YAML number_iterations
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->huber_a = 1.0;
  o->jacobi_scaling = 1; o->max_num_consecutive_invalid_steps = 5; o->verbose = 0; o->materialize_jacobian = 0;
}

const char* mcba_last_error(void* h) { return h ? ((Handle*)h)->err.c_str() : "null handle"; }
int mcba_num_kernels(void) { return KF_COUNT; }
const char* mcba_kernel_name(int k) { return (k >= 0 && k < KF_COUNT) ? kKernelNames[k] : ""; }
int mcba_kernel_stats(void* h, int k, int64_t* launches, double* total_ms) {
  Handle* H = (Handle*)h; if (!H || k < 0 || k >= KF_COUNT) return MCBA_ERR_INVALID;
  flush_timers(H);
  if (launches) *launches = H->kf_launches[k];
  if (total_ms) *total_ms = H->kf_ms[k];
  return MCBA_OK;
}
int mcba_reset_kernel_stats(void* h) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  flush_timers(H);
  for (int k = 0; k < KF_COUNT; ++k) { H->kf_launches[k] = 0; H->kf_ms[k] = 0; H->kf_ms_at_begin[k] = 0; }
  return MCBA_OK;
}
int mcba_jacobian_width(int stage) { return stage_dims(stage).K; }
int64_t mcba_kernel_launches(void* h) { return h ? ((Handle*)h)->n_launches : 0; }
int mcba_timer_start(void* h) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  if (!H->tm_a) { CU(cudaEventCreate(&H->tm_a)); CU(cudaEventCreate(&H->tm_b)); }
  CU(cudaStreamSynchronize(H->stream));
  CU(cudaEventRecord(H->tm_a, H->stream));
  return MCBA_OK;
}
int mcba_timer_stop(void* h, double* ms) {
  Handle* H = (Handle*)h; if (!H || !H->tm_a) return MCBA_ERR_INVALID;
  CU(cudaEventRecord(H->tm_b, H->stream));
  CU(cudaEventSynchronize(H->tm_b));
  float f = 0; CU(cudaEventElapsedTime(&f, H->tm_a, H->tm_b));
  if (ms) *ms = f;
  return MCBA_OK;
}

void mcba_destroy(void* h) {
  Handle* H = (Handle*)h; if (!H) return;
  cudaSetDevice(H->device);
  if (H->stream) cudaStreamSynchronize(H->stream);
  if (H->d_chol_prof) {
    long long p[8]; cudaMemcpy(p, H->d_chol_prof, sizeof(p), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[mcba chol prof, CTA0 cycles] diag %lld trsm %lld sync1 %lld update %lld sync2 %lld backsolve %lld\n", p[0], p[1], p[2], p[3], p[4], p[5]);
    cudaFree(H->d_chol_prof);
  }
  close_p2p(H);
  release_exported(H);
  if (H->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(H->comm);
  dfree(H, &H->d_p2p_counter); dfree(H, &H->d_p2p_counter2); dfree(H, &H->d_p2p_err); dfree(H, &H->d_epochs);
  for (auto& p : H->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : H->free_events) cudaEventDestroy(e);
  if (H->tm_a) { cudaEventDestroy(H->tm_a); cudaEventDestroy(H->tm_b); }
  if (H->obs_shared) { obs_cache_release(H->d_u); H->d_u = nullptr; H->d_v = nullptr; H->d_pt = nullptr; }
  dfree(H, &H->d_u); dfree(H, &H->d_v); dfree(H, &H->d_pt); dfree(H, &H->d_pts); dfree(H, &H->d_bobs); dfree(H, &H->d_pose_bobs_ptr); dfree(H, &H->d_bobs_flags);
  dfree(H, &H->d_k1_rec); dfree(H, &H->d_k1_aux); dfree(H, &H->d_k1_chunk_ptr); dfree(H, &H->d_k1_partial); dfree(H, &H->d_k1_partial_alt);
  dfree(H, &H->d_cam_model); dfree(H, &H->d_cam_is_ref); dfree(H, &H->d_board_is_ref);
  dfree(H, &H->d_pair_bobs); dfree(H, &H->d_chunks); dfree(H, &H->d_pair_chunk_ptr);
  dfree(H, &H->d_brec); dfree(H, &H->d_brec_alt); dfree(H, &H->d_cost_bobs); dfree(H, &H->d_Hoo); dfree(H, &H->d_Wt); dfree(H, &H->d_Z); dfree(H, &H->d_L);
  dfree(H, &H->d_scale_e); dfree(H, &H->d_diag_e); dfree(H, &H->d_scale_f); dfree(H, &H->d_diag_f);
  dfree(H, &H->d_FF); dfree(H, &H->d_gf); dfree(H, &H->d_S); dfree(H, &H->d_rhs); dfree(H, &H->d_zf); dfree(H, &H->d_ZtZ); dfree(H, &H->d_hz);
  dfree(H, &H->d_syrk_partial); dfree(H, &H->d_pair_partial); dfree(H, &H->d_pair_rec); dfree(H, &H->d_part_e);
  dfree(H, &H->d_red_scratch); dfree(H, &H->d_scalars); dfree(H, &H->d_gather); dfree(H, &H->d_fail);
  dfree(H, &H->d_jmat); dfree(H, &H->d_resmat); dfree(H, &H->d_ct_flags); dfree(H, &H->d_ct_dinv); dfree(H, &H->d_ct_pbuf); dfree(H, &H->d_tile_mask); dfree(H, &H->d_ct_epoch); dfree(H, &H->d_radius);
  drop_graphs(H);
  if (H->h_pin) { cudaFreeHost(H->h_pin); H->h_pin = nullptr; }
  dfree(H, &H->d_cam_ptr); dfree(H, &H->d_lm); dfree(H, &H->d_gate_jac); dfree(H, &H->d_active); dfree(H, &H->d_cam_fail); dfree(H, &H->d_n_active);
  dfree(H, &H->d_cam_sc); dfree(H, &H->d_cam_cost);
  for (int w = 0; w < 2; ++w) { dfree(H, &H->d_cam_pose[w]); dfree(H, &H->d_pose[w]); dfree(H, &H->d_board[w]); dfree(H, &H->d_intr[w]); }
  dfree(H, &H->d_prep_cam); dfree(H, &H->d_prep_pose); dfree(H, &H->d_prep_board);
  if (H->stream_b) { cudaStreamSynchronize(H->stream_b); cudaStreamDestroy(H->stream_b); cudaEventDestroy(H->ev_fork); cudaEventDestroy(H->ev_join); }
  if (H->stream) { cudaStreamSynchronize(H->stream); cudaStreamDestroy(H->stream); }
  delete H;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int mcba_create(void** handle, const mcba_problem* P, int device) {
  if (!handle) return MCBA_ERR_INVALID;
  static const bool prof = getenv("MCBA_CREATE_PROF") != nullptr;
  const double t_begin = now_ms(); double t_mark = t_begin;
  auto mark = [&](const char* what) { if (prof) { const double t = now_ms(); std::fprintf(stderr, "[mcba_create] %-28s %8.3f ms\n", what, t - t_mark); t_mark = t; } };
  Handle* H = new Handle();
  *handle = H;
  if (!P || P->n_obs < 0 || P->n_bobs < 0 || P->n_cam <= 0 || P->n_pose < 0) { H->err = "bad problem sizes"; return MCBA_ERR_INVALID; }
  if (!P->cam_model || !P->bobs_ptr || (P->n_bobs > 0 && (!P->bobs_cam || !P->bobs_pose)) || (P->n_obs > 0 && (!P->u || !P->v || !P->pt_idx)) ||
      (P->n_pts > 0 && !P->pts_xyz) || P->n_pts < 0) { H->err = "missing array in mcba_problem"; return MCBA_ERR_INVALID; }
  if (P->n_obs >= (int64_t)1 << 31 || P->n_bobs >= (int64_t)1 << 30) { H->err = "shard too large for 32-bit indices: split across handles"; return MCBA_ERR_INVALID; }
  if (P->cam_problem || P->n_problems > 1) {
    // batched independent problems: stage 1, one problem per camera, one board observation per e-block, camera-major
    bool ok = P->stage == 1 && P->cam_problem && P->n_problems == P->n_cam && P->n_pose == P->n_bobs;
    for (int c = 0; ok && c < P->n_cam; ++c) ok = P->cam_problem[c] == c;
    for (int64_t b = 0; ok && b < P->n_bobs; ++b) ok = P->bobs_pose[b] == b && (b == 0 || P->bobs_cam[b] >= P->bobs_cam[b - 1]);
    if (!ok) { H->err = "batched mode needs stage 1, cam_problem[c] == c, bobs_pose[b] == b and bobs sorted by camera"; return MCBA_ERR_INVALID; }
    H->batched = true;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { H->err = "no CUDA device: libmcba has no CPU fallback"; return MCBA_ERR_CUDA; }
  int cc_major = 0, n_sm = 0, coop = 0;   // cudaGetDeviceProperties costs ~6 ms per call: query the three attributes instead
  CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (cc_major != 10) { H->err = "libmcba is built for sm_100a (B200) only"; return MCBA_ERR_CUDA; }
  CU(cudaSetDevice(device));
  H->device = device;
  keep_pool(device);
  CU(cudaStreamCreate(&H->stream));
  {
    static const bool overlap = []() { const char* e = getenv("MCBA_OVERLAP"); return !e || atoi(e) != 0; }();
    if (overlap) { CU(cudaStreamCreateWithFlags(&H->stream_b, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&H->ev_join, cudaEventDisableTiming)); }
  }
  H->n_cam = P->n_cam; H->n_pose = P->n_pose; H->n_board = std::max(0, P->n_board); H->n_pts = P->n_pts;
  H->n_obs = P->n_obs; H->nb = (int)P->n_bobs; H->npad = ((P->n_obs + 31) / 32) * 32; H->n_obs_global = P->n_obs;
  mark("device + stream");
  // validate + pack the board-observation table (host, small)
  for (int c = 0; c < P->n_cam; ++c)
    if (P->cam_model[c] != 0 && P->cam_model[c] != 1) { H->err = "cam_model must be 0 (Brown) or 1 (Kannala)"; return MCBA_ERR_INVALID; }
  std::vector<int4> bobs(H->nb + 1);
  std::vector<int32_t>& bflags = H->h_bobs_flags;
  bflags.assign(H->nb + 1, 0);
  std::vector<int> ppt(H->n_pose + 1, 0);
  H->h_bobs_cam.resize(H->nb); H->h_bobs_pose.resize(H->nb); H->h_bobs_board.resize(H->nb);
  H->h_bobs_ptr.assign(P->bobs_ptr, P->bobs_ptr + H->nb + 1);
  {
    long long p32 = 0, p16 = 0;      // Jacobian passes: one observation per warp / two per warp (a pair's passes = the longer half's)
    for (int b = 0; b < H->nb; ++b) {
      const long long c = P->bobs_ptr[b + 1] - P->bobs_ptr[b];
      if (c > 0) { p32 += (c + 31) / 32; if (b % 2 == 0) { const long long c2 = b + 1 < H->nb ? P->bobs_ptr[b + 2] - P->bobs_ptr[b + 1] : 0; p16 += (std::max(c, c2) + 15) / 16; } }
      else if (b % 2 == 0 && b + 1 < H->nb) { p16 += (P->bobs_ptr[b + 2] - P->bobs_ptr[b + 1] + 15) / 16; }
    }
    H->k1_pair_saves = H->batched || 10 * p16 < 9 * p32;
  }
  for (int b = 0; b < H->nb; ++b) {
    const int c = P->bobs_cam[b], o = P->bobs_pose[b], bd = P->bobs_board ? P->bobs_board[b] : 0;
    if (c < 0 || c >= P->n_cam || o < 0 || o >= P->n_pose || (b > 0 && o < P->bobs_pose[b - 1]) ||
        P->bobs_ptr[b] > P->bobs_ptr[b + 1] || P->bobs_ptr[b] < 0 || P->bobs_ptr[b + 1] > P->n_obs ||
        (H->n_board > 0 && (bd < 0 || bd >= H->n_board))) { H->err = "inconsistent board-observation table"; return MCBA_ERR_INVALID; }
    bobs[b] = make_int4((int)P->bobs_ptr[b], c, o, H->n_board > 0 ? bd : 0);
    bflags[b] = ((P->cam_is_ref && P->cam_is_ref[c]) ? 1 : 0) | ((P->board_is_ref && H->n_board > 0 && P->board_is_ref[bd]) ? 2 : 0) | (P->cam_model[c] << 2);
    H->h_bobs_cam[b] = c; H->h_bobs_pose[b] = o; H->h_bobs_board[b] = H->n_board > 0 ? bd : 0;
    ppt[o + 1] = b + 1;
  }
  if (H->nb > 0 && (P->bobs_ptr[0] != 0 || P->bobs_ptr[H->nb] != P->n_obs)) { H->err = "bobs_ptr must cover [0, n_obs)"; return MCBA_ERR_INVALID; }
  bobs[H->nb] = make_int4((int)P->n_obs, 0, 0, 0);
  for (int o = 0; o < H->n_pose; ++o) ppt[o + 1] = std::max(ppt[o + 1], ppt[o]);
  // the three big uploads run behind the host work that follows (point table, pair lists of setup_stage)
  if (dalloc(H, &H->d_fail, (size_t)2)) return MCBA_ERR_CUDA;
  CU(cudaMemsetAsync(H->d_fail, 0, 2 * sizeof(int), H->stream));
  if (P->n_obs > 0 && obs_cache_acquire(P->u, P->v, P->pt_idx, P->n_obs, device, &H->d_u, &H->d_v, &H->d_pt)) {
    H->obs_shared = true;
  } else {
    if (dalloc(H, &H->d_u, (size_t)H->npad) || dalloc(H, &H->d_v, (size_t)H->npad) || dalloc(H, &H->d_pt, (size_t)H->npad)) return MCBA_ERR_CUDA;
    CU(cudaMemcpyAsync(H->d_u, P->u, P->n_obs * sizeof(float), cudaMemcpyHostToDevice, H->stream));
    CU(cudaMemcpyAsync(H->d_v, P->v, P->n_obs * sizeof(float), cudaMemcpyHostToDevice, H->stream));
    CU(cudaMemcpyAsync(H->d_pt, P->pt_idx, P->n_obs * sizeof(int32_t), cudaMemcpyHostToDevice, H->stream));
  }
  if (P->n_obs > 0) {
    ++H->n_launches;
    k_validate_pt<<<std::min<int64_t>(4 * n_sm, (P->n_obs + 255) / 256), 256, 0, H->stream>>>(H->d_pt, P->n_obs, P->n_pts, H->d_fail + 1);
    CHECK_LAUNCH();
  }
  std::vector<double> pts((size_t)std::max(1, P->n_pts) * 3);
  for (int i = 0; i < P->n_pts * 3; ++i) pts[i] = (double)P->pts_xyz[i];
  mark("validate + pack (host)");
  if (dalloc(H, &H->d_pts, pts.size()) || dalloc(H, &H->d_bobs, bobs.size()) || dalloc(H, &H->d_pose_bobs_ptr, ppt.size()) || dalloc(H, &H->d_bobs_flags, bflags.size())) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_cam_model, (size_t)H->n_cam) || dalloc(H, &H->d_cam_is_ref, (size_t)H->n_cam) || dalloc(H, &H->d_board_is_ref, (size_t)std::max(1, H->n_board))) return MCBA_ERR_CUDA;
  // small tables from vector storage of this call: pageable copies are staged before cudaMemcpyAsync returns
  CU(cudaMemcpyAsync(H->d_pts, pts.data(), pts.size() * sizeof(double), cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemcpyAsync(H->d_bobs, bobs.data(), bobs.size() * sizeof(int4), cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemcpyAsync(H->d_pose_bobs_ptr, ppt.data(), ppt.size() * sizeof(int), cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemcpyAsync(H->d_bobs_flags, bflags.data(), bflags.size() * sizeof(int), cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemcpyAsync(H->d_cam_model, P->cam_model, H->n_cam * sizeof(int32_t), cudaMemcpyHostToDevice, H->stream));
  std::vector<uint8_t> zc((size_t)std::max(H->n_cam, std::max(1, H->n_board)), 0);
  CU(cudaMemcpyAsync(H->d_cam_is_ref, P->cam_is_ref ? P->cam_is_ref : zc.data(), H->n_cam, cudaMemcpyHostToDevice, H->stream));
  CU(cudaMemcpyAsync(H->d_board_is_ref, (P->board_is_ref && H->n_board > 0) ? P->board_is_ref : zc.data(), std::max(1, H->n_board), cudaMemcpyHostToDevice, H->stream));
  mark("observation alloc + H2D");
  for (int w = 0; w < 2; ++w) {
    if (dalloc(H, &H->d_cam_pose[w], (size_t)H->n_cam * 6) || dalloc(H, &H->d_pose[w], (size_t)H->n_pose * 6) ||
        dalloc(H, &H->d_board[w], (size_t)std::max(1, H->n_board) * 6) || dalloc(H, &H->d_intr[w], (size_t)H->n_cam * 9)) return MCBA_ERR_CUDA;
    CU(cudaMemset(H->d_board[w], 0, (size_t)std::max(1, H->n_board) * 6 * sizeof(double)));
  }
  if (dalloc(H, &H->d_prep_cam, (size_t)H->n_cam * PREP) || dalloc(H, &H->d_prep_pose, (size_t)H->n_pose * PREP) ||
      dalloc(H, &H->d_prep_board, (size_t)std::max(1, H->n_board) * PREP)) return MCBA_ERR_CUDA;
  if (dalloc(H, &H->d_red_scratch, (size_t)512 * 8) || dalloc(H, &H->d_scalars, (size_t)SC_N) || dalloc(H, &H->d_gather, (size_t)SC_N * 64)) return MCBA_ERR_CUDA;
  CU(cudaMemset(H->d_scalars, 0, SC_N * sizeof(double)));
  H->n_sm = n_sm;
  CU(cudaFuncSetAttribute(k_chol_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, CC_SMEM));
  { int per_sm = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_coop, CC_THREADS, CC_SMEM));
    if (per_sm < 1 || !coop) { H->err = "cooperative launch unavailable"; return MCBA_ERR_CUDA; }
    H->chol_grid = H->n_sm; }
  CU(cudaFuncSetAttribute(k_chol_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
  { int per_sm = 0; CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_tiles, CT_THREADS, CT_SMEM));
    H->ct_capacity = std::min(CT_MAXT, per_sm * H->n_sm); }
  if (dalloc(H, &H->d_ct_flags, (size_t)2 * CT_MAXT + 64) || dalloc(H, &H->d_ct_dinv, (size_t)CT_MAXNT * 512) || dalloc(H, &H->d_ct_pbuf, (size_t)CT_MAXT * 64)) return MCBA_ERR_CUDA;
  CU(cudaMemsetAsync(H->d_ct_flags, 0, ((size_t)2 * CT_MAXT + 64) * sizeof(unsigned), H->stream));
  if (dalloc(H, &H->d_ct_epoch, (size_t)1) || dalloc(H, &H->d_radius, (size_t)1)) return MCBA_ERR_CUDA;
  CU(cudaMemsetAsync(H->d_ct_epoch, 0, sizeof(unsigned), H->stream));
  CU(cudaMallocHost((void**)&H->h_pin, (PIN_GATHER + 2 * P2P_MAX_RANKS * P2P_GATHER_W) * sizeof(double)));
  H->k1_variant = g_k1_variant >= 0 ? g_k1_variant : []() { const char* e = getenv("MCBA_K1"); return e ? atoi(e) : -1; }();
  { const char* e = getenv("MCBA_GRAPH"); H->graph_env = !H->batched && !(e && atoi(e) == 0); H->use_graph = H->graph_env; }   // replay the per-iteration launch sequences as CUDA graphs
  mcba_default_options(&H->opt);
  mark("small allocs");
  const int rc = setup_stage(H, P->stage);
  if (rc) return rc;
  mark("setup_stage (pair lists + allocs)");
  int bad_pt = 0;   // the stream sync below also ends the lifetime requirement of the host vectors above
  CU(cudaMemcpyAsync(&bad_pt, H->d_fail + 1, sizeof(int), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  if (bad_pt) { H->err = "pt_idx out of range"; return MCBA_ERR_INVALID; }
  mark("device sync");
  if (prof) std::fprintf(stderr, "[mcba_create] total %8.3f ms\n", now_ms() - t_begin);
  return MCBA_OK;
}

int mcba_set_stage(void* h, int stage) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  if (H->batched) { H->err = "batched handles are stage 1 only"; return MCBA_ERR_INVALID; }
  CU(cudaSetDevice(H->device));
  if (stage == H->stage) return MCBA_OK;
  int rc;
  if (H->n_ranks > 1) {   // every rank unmaps its peers' buffers before anybody recycles them
    close_p2p(H);
    if ((rc = allreduce_sum(H, H->d_scalars, 1))) return rc;
    CU(cudaStreamSynchronize(H->stream));
    release_exported(H);
  }
  rc = setup_stage(H, stage);
  if (rc) return rc;
  CU(cudaDeviceSynchronize());
  if (H->n_ranks > 1 && (rc = setup_p2p(H))) return rc;   // collective: every rank changes stage
  H->use_graph = H->graph_env && (H->n_ranks == 1 || H->p2p);
  return MCBA_OK;
}

int mcba_nccl_unique_id(void* out) {
  if (!g_nccl.load()) return MCBA_ERR_NCCL;
  nccl_uid_t id;
  if (g_nccl.GetUniqueId(&id) != 0) return MCBA_ERR_NCCL;
  std::memcpy(out, &id, sizeof(id));
  return MCBA_OK;
}

int mcba_comm_init(void* h, int rank, int n_ranks, const void* uid) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || n_ranks > 64) { H->err = "bad rank / n_ranks"; return MCBA_ERR_INVALID; }
  H->rank = rank; H->n_ranks = n_ranks;
  if (n_ranks == 1) return MCBA_OK;
  H->use_graph = false; drop_graphs(H);   // decided again below: graphs with ranks need every exchange to go over peer memory
  if (!g_nccl.load()) { H->err = "libnccl.so.2 not found"; return MCBA_ERR_NCCL; }
  CU(cudaSetDevice(H->device));
  nccl_uid_t id; std::memcpy(&id, uid, sizeof(id));
  const int rc = g_nccl.CommInitRank(&H->comm, n_ranks, id, rank);
  if (rc != 0) { H->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return MCBA_ERR_NCCL; }
  // global observation count for the RMS
  double n_local = (double)H->n_obs;
  CU(cudaMemcpy(H->d_scalars, &n_local, sizeof(double), cudaMemcpyHostToDevice));
  if (g_nccl.AllReduce(H->d_scalars, H->d_scalars, 1, NCCL_DOUBLE, NCCL_SUM, H->comm, H->stream) != 0) { H->err = "ncclAllReduce(n_obs)"; return MCBA_ERR_NCCL; }
  double n_glob = 0;
  CU(cudaMemcpyAsync(&n_glob, H->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  H->n_obs_global = (int64_t)(n_glob + 0.5);
  // NCCL sets up its channels for a message size on first use (~100 ms for the 8 MB reduced-system allreduce):
  // pay that here, once per communicator, not inside the first LM iteration.
  if (!H->batched) {
    CU(cudaMemsetAsync(H->d_ZtZ, 0, (size_t)H->ldz * H->ldz * sizeof(double), H->stream));
    CU(cudaMemsetAsync(H->d_pair_rec, 0, (size_t)H->n_pair * H->NE * sizeof(double), H->stream));
    int rc;
    if ((rc = allreduce_sum(H, H->d_ZtZ, (size_t)H->ldz * H->ldz))) return rc;
    if ((rc = allreduce_sum(H, H->d_pair_rec, (size_t)H->n_pair * H->NE))) return rc;
    std::vector<double> hs;
    if ((rc = gather_scalars(H, H->d_scalars, SC_N, hs))) return rc;
    flush_timers(H);
    for (int k = 0; k < KF_COUNT; ++k) { H->kf_launches[k] = 0; H->kf_ms[k] = 0; }
    if ((rc = setup_p2p(H))) return rc;
    H->use_graph = H->graph_env && H->p2p;
  }
  return MCBA_OK;
}

int mcba_solve_begin(void* h, const mcba_params* p, const mcba_options* opt) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  if (H->batched) { H->err = "batched handle: use mcba_solve_batched"; return MCBA_ERR_INVALID; }
  CU(cudaSetDevice(H->device));
  mcba_options o; if (opt) o = *opt; else mcba_default_options(&o);
  if (H->graph_opt_valid && std::memcmp(&H->graph_opt, &o, sizeof(o)) != 0) drop_graphs(H);   // options are baked into the captured arguments
  H->graph_opt = o; H->graph_opt_valid = true;
  H->opt = o;
  int rc;
  if (o.materialize_jacobian && (rc = ensure_jmat(H))) return rc;   // no allocation inside a stream capture
  if ((rc = upload_params(H, p))) return rc;
  H->x0_pose.assign(p->pose, p->pose + (size_t)H->n_pose * 6);
  H->x0_intr.assign(p->intrinsics, p->intrinsics + (size_t)H->n_cam * 9);
  if (p->cam_pose) H->x0_cam.assign(p->cam_pose, p->cam_pose + (size_t)H->n_cam * 6); else H->x0_cam.clear();
  if (p->board_pose && H->n_board > 0) H->x0_board.assign(p->board_pose, p->board_pose + (size_t)H->n_board * 6); else H->x0_board.clear();
  flush_timers(H);
  for (int k = 0; k < KF_COUNT; ++k) H->kf_ms_at_begin[k] = H->kf_ms[k];
  H->solving = true; H->trace.clear(); H->have_scale = false;
  std::memset(&H->summ, 0, sizeof(H->summ));
  lm_init(H->lm, o);
  rc = eval_jacobian(H, true);   // IterationZero
  if (rc == MCBA_ERR_NUMERIC) { H->err = "Initial residual and Jacobian evaluation failed."; H->summ.termination = MCBA_FAILURE; H->lm.termination = MCBA_FAILURE; H->lm.done = 1; return rc; }
  if (rc) return rc;
  lm_iteration_zero(H->lm, H->ev_cost, H->ev_grad_max, H->ev_x_norm);
  H->summ.initial_cost = H->lm.initial_cost;
  return MCBA_OK;
}

// One pass of TrustRegionMinimizer's while loop (decision rules in mcba_lm.cuh): FinalizeIterationAndCheck... for the
// iteration in flight, then one trust-region iteration. Returns 1 to continue, 0 when terminated.
int mcba_solve_step(void* h) {
  Handle* H = (Handle*)h; if (!H || !H->solving) return MCBA_ERR_INVALID;
  if (H->lm.done) return 0;
  CU(cudaSetDevice(H->device));
  mcba_trace_row row;
  const int cont = lm_finalize(H->lm, H->opt, &row);
  H->trace.push_back(row);
  if (H->opt.verbose && H->rank == 0)
    std::printf("%4d % .6e % .2e % .2e % .2e % .2e % .2e\n", row.iteration, row.cost, row.cost_change, row.gradient_max_norm,
                row.step_norm, row.relative_decrease, row.trust_region_radius);
  if (!cont) return 0;
  LmStep so;
  H->cand_has_jac = false;
  int rc = compute_step(H, &so);
  if (rc) return rc;
  const int d = lm_decide(H->lm, H->opt, so);
  if (d == LM_TERMINATED) return 0;
  if (d == LM_ACCEPT) {   // HandleSuccessfulStep
    H->cur = 1 - H->cur;
    if (H->cand_has_jac) { std::swap(H->d_brec, H->d_brec_alt); std::swap(H->d_k1_partial, H->d_k1_partial_alt); H->brec_par ^= 1; rc = eval_assemble(H, false); }
    else rc = eval_jacobian(H, false);
    if (rc == MCBA_ERR_NUMERIC) { H->lm.termination = MCBA_FAILURE; H->lm.done = 1; return 0; }
    if (rc) return rc;
    lm_accepted_eval(H->lm, H->opt, H->ev_cost, H->ev_grad_max);
  }
  return 1;
}

int mcba_solve_end(void* h, mcba_params* p, mcba_summary* summary) {
  Handle* H = (Handle*)h; if (!H || !H->solving) return MCBA_ERR_INVALID;
  CU(cudaSetDevice(H->device));
  int rc;
  H->summ.termination = H->lm.termination;
  H->summ.num_iterations = H->lm.num_recorded;
  H->summ.num_successful_steps = H->lm.num_successful; H->summ.num_unsuccessful_steps = H->lm.num_unsuccessful;
  H->summ.final_cost = H->lm.minimum_cost < DBL_MAX ? H->lm.minimum_cost : H->lm.x_cost;
  if (H->lm.termination == MCBA_FAILURE) {
    // ceres::Solve (solver.cc, Minimize): when the summary is not usable (FAILURE) the parameter blocks get the
    // ORIGINAL values back, whatever the minimizer had reached. Same here: the start point is re-uploaded and returned.
    mcba_params x0 = {H->x0_cam.empty() ? nullptr : H->x0_cam.data(), H->x0_pose.data(), H->x0_board.empty() ? nullptr : H->x0_board.data(), H->x0_intr.data()};
    const StageDims sd = stage_dims(H->stage);
    if ((!sd.cam || x0.cam_pose) && (!sd.board || x0.board_pose) && (rc = upload_params(H, &x0))) return rc;
  }
  // unrobustified RMS at the returned parameters: cost pass with the loss disabled
  if ((rc = launch_prep(H, H->cur))) return rc;
  if ((rc = launch_obs<MODE_COST>(H, H->cur, 0.0))) return rc;
  if ((rc = reduce_rows(H, H->d_cost_bobs, H->nb, 1, 1, H->d_scalars + SC_COST))) return rc;
  std::vector<double> hs;
  if ((rc = gather_scalars(H, H->d_scalars, SC_N, hs))) return rc;
  double c = 0; for (int r = 0; r < H->n_ranks; ++r) c += hs[(size_t)r * SC_N + SC_COST];
  H->summ.rms_px = std::sqrt(2.0 * c / (double)std::max<int64_t>(1, H->n_obs_global));
  flush_timers(H);
  double dt[KF_COUNT];   // this solve only (a handle is reused across stages and mcba_eval calls)
  for (int k = 0; k < KF_COUNT; ++k) dt[k] = H->kf_ms[k] - H->kf_ms_at_begin[k];
  H->summ.t_jacobian_ms = dt[KF_FUSED] + dt[KF_MATERIALIZE] + dt[KF_FROM_J] + dt[KF_PREP];
  H->summ.t_schur_ms = dt[KF_EASSEMBLE] + dt[KF_PAIR] + dt[KF_EFACTOR] + dt[KF_SYRK] + dt[KF_REDUCED] + dt[KF_BACKSUBST];
  H->summ.t_chol_ms = dt[KF_CHOL]; H->summ.t_cost_ms = dt[KF_COST]; H->summ.t_comm_ms = dt[KF_NCCL];
  double tot = 0; for (int k = 0; k < KF_COUNT; ++k) tot += dt[k];
  H->summ.t_total_ms = tot;
  if (p && (rc = download_params(H, p))) return rc;
  if (summary) *summary = H->summ;
  H->solving = false;
  return MCBA_OK;
}

int mcba_solve(void* h, mcba_params* p, const mcba_options* opt, mcba_summary* summary, mcba_trace_row* trace, int trace_cap) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  int rc = mcba_solve_begin(h, p, opt);
  if (rc) { if (summary) *summary = H->summ; return rc; }
  for (;;) { rc = mcba_solve_step(h); if (rc <= 0) break; }
  if (rc < 0) return rc;
  rc = mcba_solve_end(h, p, summary);
  if (trace) for (int i = 0; i < (int)H->trace.size() && i < trace_cap; ++i) trace[i] = H->trace[i];
  return rc;
}

// Jacobian phase of the batched driver for the cameras flagged in d_gate_jac
static int batched_eval(Handle* H, bool first) {
  int rc;
  if ((rc = launch_prep(H, 0))) return rc;
  H->obs_gate = H->d_gate_jac;
  rc = launch_obs<MODE_FUSED>(H, 0, H->opt.huber_a);
  H->obs_gate = nullptr;
  if (rc) return rc;
  Timed t(H, KF_PAIR);
  const int nc = H->n_cam, np = H->n_pose;
  if (first) { ++H->n_launches; k_b_escale<<<(np + 127) / 128, 128, 0, H->stream>>>(H->d_bobs, H->d_gate_jac, H->d_brec, np, H->opt.jacobi_scaling, H->d_scale_e); }
  ++H->n_launches;
  k_b_cam_reduce<<<nc, 64, 0, H->stream>>>(H->d_gate_jac, H->d_cam_ptr, H->d_brec, H->d_cost_bobs, H->d_pose[0], H->d_FF, H->d_gf, H->d_cam_sc);
  ++H->n_launches;
  k_b_after_eval<<<(nc + 127) / 128, 128, 0, H->stream>>>(H->d_lm, H->opt, H->d_gate_jac, nc, first ? 1 : 0, H->d_FF, H->d_gf, H->d_cam_sc, H->d_intr[0], H->d_scale_f);
  CHECK_LAUNCH();
  return MCBA_OK;
}

// Config 5: all problems advance together; each takes its own trust-region decisions on the device.
int mcba_solve_batched(void* h, mcba_params* p, const mcba_options* opt, mcba_summary* summaries) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  if (!H->batched) { H->err = "not a batched handle (mcba_problem::cam_problem was NULL)"; return MCBA_ERR_INVALID; }
  CU(cudaSetDevice(H->device));
  mcba_options o; if (opt) o = *opt; else mcba_default_options(&o);
  H->opt = o;
  int rc;
  if ((rc = upload_params(H, p))) return rc;
  const int nc = H->n_cam, np = H->n_pose;
  ++H->n_launches;
  k_b_init<<<(nc + 127) / 128, 128, 0, H->stream>>>(H->d_lm, o, nc, H->d_gate_jac, H->d_active);
  CHECK_LAUNCH();
  if ((rc = batched_eval(H, true))) return rc;
  for (int64_t guard = 0; guard < (int64_t)o.max_num_iterations * 8 + 64; ++guard) {
    CU(cudaMemsetAsync(H->d_n_active, 0, sizeof(int), H->stream));
    ++H->n_launches;
    k_b_finalize<<<(nc + 127) / 128, 128, 0, H->stream>>>(H->d_lm, o, nc, H->d_active, H->d_n_active);
    int n_active = 0;
    CU(cudaMemcpyAsync(&n_active, H->d_n_active, sizeof(int), cudaMemcpyDeviceToHost, H->stream));
    CU(cudaStreamSynchronize(H->stream));
    if (n_active == 0) break;
    {
      Timed t(H, KF_EFACTOR);
      ++H->n_launches;
      k_b_eblock_factor<<<(np + 127) / 128, 128, 0, H->stream>>>(H->d_bobs, H->d_active, H->d_lm, o, H->d_brec, H->d_scale_e, H->d_scale_f, H->d_diag_e, np, H->d_L, H->d_Z, H->d_cam_fail);
    }
    {
      Timed t(H, KF_SYRK);
      ++H->n_launches;
      k_b_schur_solve<<<nc, 64, 0, H->stream>>>(H->d_active, H->d_lm, o, H->d_cam_ptr, H->d_FF, H->d_gf, H->d_Z, H->d_scale_f, H->d_diag_f, H->d_intr[0], H->d_intr[1], H->d_zf, H->d_cam_sc, H->d_cam_fail);
    }
    {
      Timed t(H, KF_BACKSUBST);
      ++H->n_launches;
      k_b_backsubst<<<(np + 127) / 128, 128, 0, H->stream>>>(H->d_bobs, H->d_active, H->d_lm, H->d_Z, H->d_L, H->d_zf, H->d_scale_e, H->d_diag_e, H->d_pose[0], np, H->d_pose[1], H->d_part_e);
    }
    CHECK_LAUNCH();
    if ((rc = launch_prep(H, 1))) return rc;
    H->obs_gate = H->d_active;
    rc = launch_obs<MODE_COST>(H, 1, o.huber_a);
    H->obs_gate = nullptr;
    if (rc) return rc;
    {
      Timed t(H, KF_REDUCE);
      ++H->n_launches;
      k_b_decide<<<nc, 32, 0, H->stream>>>(H->d_active, H->d_lm, o, H->d_cam_ptr, H->d_part_e, H->d_cost_bobs, H->d_cam_sc, H->d_cam_fail, H->d_gate_jac);
      ++H->n_launches;
      k_b_commit<<<(np + nc + 127) / 128, 128, 0, H->stream>>>(H->d_bobs, H->d_gate_jac, np, nc, H->d_pose[1], H->d_pose[0], H->d_intr[1], H->d_intr[0]);
    }
    CHECK_LAUNCH();
    if ((rc = batched_eval(H, false))) return rc;
  }
  // per-problem RMS of the unrobustified residuals at the returned parameters
  if ((rc = launch_prep(H, 0))) return rc;
  if ((rc = launch_obs<MODE_COST>(H, 0, 0.0))) return rc;
  ++H->n_launches;
  k_b_cam_cost<<<(nc + 127) / 128, 128, 0, H->stream>>>(H->d_cam_ptr, H->d_cost_bobs, nc, H->d_cam_cost);
  CHECK_LAUNCH();
  std::vector<LmState> lm(nc); std::vector<double> cc(nc);
  CU(cudaMemcpyAsync(lm.data(), H->d_lm, nc * sizeof(LmState), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaMemcpyAsync(cc.data(), H->d_cam_cost, nc * sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  H->cur = 0;
  if (p && (rc = download_params(H, p))) return rc;
  flush_timers(H);
  if (summaries) {
    for (int c = 0; c < nc; ++c) {
      mcba_summary& S = summaries[c];
      std::memset(&S, 0, sizeof(S));
      S.termination = lm[c].termination; S.num_successful_steps = lm[c].num_successful; S.num_unsuccessful_steps = lm[c].num_unsuccessful;
      S.num_iterations = lm[c].num_recorded; S.initial_cost = lm[c].initial_cost;
      S.final_cost = lm[c].minimum_cost < DBL_MAX ? lm[c].minimum_cost : lm[c].x_cost;
      const int64_t n_obs_c = H->h_bobs_ptr[H->h_cam_ptr[c + 1]] - H->h_bobs_ptr[H->h_cam_ptr[c]];
      S.rms_px = std::sqrt(2.0 * cc[c] / (double)std::max<int64_t>(1, n_obs_c));
    }
  }
  return MCBA_OK;
}

int mcba_eval(void* h, const mcba_params* p, const mcba_options* opt, double* cost, double* residuals, double* jacobian) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  CU(cudaSetDevice(H->device));
  mcba_options o; if (opt) o = *opt; else mcba_default_options(&o);
  int rc;
  if ((rc = upload_params(H, p))) return rc;
  if ((rc = ensure_jmat(H))) return rc;
  if ((rc = launch_prep(H, 0))) return rc;
  if ((rc = launch_obs<MODE_MATERIALIZE>(H, 0, o.huber_a))) return rc;
  if ((rc = reduce_rows(H, H->d_cost_bobs, H->nb, 1, 1, H->d_scalars + SC_COST))) return rc;
  double c = 0;
  CU(cudaMemcpyAsync(&c, H->d_scalars + SC_COST, sizeof(double), cudaMemcpyDeviceToHost, H->stream));
  CU(cudaStreamSynchronize(H->stream));
  if (cost) *cost = c;
  const int K = H->K, NC = K + 1;
  if (residuals || jacobian) {
    std::vector<double> blk((size_t)2 * NC * H->n_obs);
    CU(cudaMemcpy(blk.data(), H->d_jmat, blk.size() * sizeof(double), cudaMemcpyDeviceToHost));
    int ref_of[32];
    DISPATCH_STAGE(H->stage, { for (int c = 0; c < K; ++c) ref_of[c] = Cols<ST>::to_ref(c); });
    for (int b = 0; b < H->nb; ++b) {
      const int64_t o0 = H->h_bobs_ptr[b], nobs = H->h_bobs_ptr[b + 1] - o0;
      const double* base = blk.data() + (size_t)2 * NC * o0;
      for (int64_t j = 0; j < nobs; ++j) {
        const int64_t i = o0 + j;
        if (jacobian)
          for (int c = 0; c < K; ++c) {
            jacobian[((size_t)i * 2 + 0) * K + ref_of[c]] = base[((size_t)c * nobs + j) * 2];
            jacobian[((size_t)i * 2 + 1) * K + ref_of[c]] = base[((size_t)c * nobs + j) * 2 + 1];
          }
        if (residuals) { residuals[2 * i] = base[((size_t)K * nobs + j) * 2]; residuals[2 * i + 1] = base[((size_t)K * nobs + j) * 2 + 1]; }
      }
    }
  }
  return std::isfinite(c) ? MCBA_OK : MCBA_ERR_NUMERIC;
}

static int reprojection_errors(void* h, const mcba_params* p, double* err, float* err_ref) {
  Handle* H = (Handle*)h; if (!H || (!err && !err_ref)) return MCBA_ERR_INVALID;
  CU(cudaSetDevice(H->device));
  int rc;
  if ((rc = upload_params(H, p))) return rc;
  if ((rc = launch_prep(H, 0))) return rc;
  double* d_err = nullptr; float* d_ref = nullptr;
  if (err_ref ? dalloc(H, &d_ref, (size_t)H->npad) : dalloc(H, &d_err, (size_t)H->npad)) return MCBA_ERR_CUDA;
  ObsArgs a = obs_args(H, 0, 0.0);
  if (H->nb > 0) {
    ++H->n_launches;
    DISPATCH_STAGE(H->stage, (k_reproj_err<ST><<<grid_for(H, H->nb, 8), OBS_WARPS * 32, OBS_WARPS * CTX_SIZE * 8, H->stream>>>(a, d_err, d_ref)));
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = err_ref ? cudaMemcpyAsync(err_ref, d_ref, H->n_obs * sizeof(float), cudaMemcpyDeviceToHost, H->stream)
                                    : cudaMemcpyAsync(err, d_err, H->n_obs * sizeof(double), cudaMemcpyDeviceToHost, H->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(H->stream);
  if (d_err) cudaFreeAsync(d_err, H->stream);
  if (d_ref) cudaFreeAsync(d_ref, H->stream);
  if (e != cudaSuccess) { H->err = cudaGetErrorString(e); return MCBA_ERR_CUDA; }
  return MCBA_OK;
}
int mcba_reprojection_errors(void* h, const mcba_params* p, double* err) { return reprojection_errors(h, p, err, nullptr); }
int mcba_reprojection_errors_f32(void* h, const mcba_params* p, float* err) { return reprojection_errors(h, p, nullptr, err); }

// Test / benchmark hook: factor and solve one dense system with the reduced-system kernels on their own.
// S_host is (n+1) x n row-major (symmetric matrix, then the rhs row). variant 0 = k_chol_coop, 1 = k_chol_tiles,
// -1 = what the solver would pick. Returns the mean kernel time over `iters` runs (the system is re-uploaded each time).
int mcba_debug_cholesky(void* h, int n, const double* S_host, double* z_out, double* L_out, int variant, int iters, double* ms_out, int* fail_out) {
  Handle* H = (Handle*)h; if (!H || n <= 0 || !S_host) return MCBA_ERR_INVALID;
  CU(cudaSetDevice(H->device));
  double *S_keep = H->d_S, *z_keep = H->d_zf; const int n_keep = H->n_f;
  double *dS = nullptr, *dz = nullptr;
  CU(cudaMalloc((void**)&dS, (size_t)(n + 1) * n * sizeof(double))); CU(cudaMalloc((void**)&dz, (size_t)n * sizeof(double)));
  H->d_S = dS; H->d_zf = dz; H->n_f = n;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double total = 0; int rc = MCBA_OK;
  if (variant == 0) setenv("MCBA_CHOL_TILES_ONCE", "0", 1); else if (variant == 1) setenv("MCBA_CHOL_TILES_ONCE", "1", 1); else unsetenv("MCBA_CHOL_TILES_ONCE");
  for (int it = 0; it < std::max(1, iters) && rc == MCBA_OK; ++it) {
    cudaMemcpyAsync(dS, S_host, (size_t)(n + 1) * n * sizeof(double), cudaMemcpyHostToDevice, H->stream);
    cudaMemsetAsync(H->d_fail, 0, sizeof(int), H->stream);
    cudaEventRecord(e0, H->stream);
    rc = cholesky_reduced(H);
    cudaEventRecord(e1, H->stream);
    if (cudaStreamSynchronize(H->stream) != cudaSuccess) { H->err = cudaGetErrorString(cudaGetLastError()); rc = MCBA_ERR_CUDA; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); total += ms;
  }
  unsetenv("MCBA_CHOL_TILES_ONCE");
  if (rc == MCBA_OK) {
    if (z_out) cudaMemcpy(z_out, dz, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
    if (L_out) cudaMemcpy(L_out, dS, (size_t)(n + 1) * n * sizeof(double), cudaMemcpyDeviceToHost);
    if (fail_out) cudaMemcpy(fail_out, H->d_fail, sizeof(int), cudaMemcpyDeviceToHost);
    if (ms_out) *ms_out = total / std::max(1, iters);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(dS); cudaFree(dz);
  H->d_S = S_keep; H->d_zf = z_keep; H->n_f = n_keep;
  flush_timers(H);
  return rc;
}

// Debug/test hook: copy an internal device buffer to the host. Names: "FF","gf","S","rhs","zf","ZtZ","Hoo","Wt","Z","brec","scale_e","scale_f","diag_e","diag_f"
int mcba_debug_set_k1(int variant) { g_k1_variant = variant; return MCBA_OK; }

int mcba_debug_fetch(void* h, const char* name, double* out, int64_t cap, int64_t* n_out) {
  Handle* H = (Handle*)h; if (!H) return MCBA_ERR_INVALID;
  CU(cudaSetDevice(H->device));
  const std::string s(name);
  const double* src = nullptr; int64_t n = 0;
  const int64_t nf = H->n_f, np = H->n_pose, ldz = H->ldz;
  if (s == "FF") { src = H->d_FF; n = nf * nf; } else if (s == "gf") { src = H->d_gf; n = nf; }
  else if (s == "S") { src = H->d_S; n = nf * nf; } else if (s == "rhs") { src = H->d_rhs; n = nf; }
  else if (s == "zf") { src = H->d_zf; n = nf; } else if (s == "ZtZ") { src = H->d_ZtZ; n = ldz * ldz; }
  else if (s == "Hoo") { src = H->d_Hoo; n = np * 36; } else if (s == "Wt") { src = H->d_Wt; n = np * 6 * ldz; }
  else if (s == "Z") { src = H->d_Z; n = np * 6 * ldz; } else if (s == "brec") { src = H->d_brec; n = (int64_t)H->nb * stage_dims(H->stage).NCOMP; }
  else if (s == "pair_rec") { src = H->pair_out; n = (int64_t)H->n_pair * H->NE; }
  else if (s == "scale_e") { src = H->d_scale_e; n = np * 6; } else if (s == "scale_f") { src = H->d_scale_f; n = nf; }
  else if (s == "diag_e") { src = H->d_diag_e; n = np * 6; } else if (s == "diag_f") { src = H->d_diag_f; n = nf; }
  else if (s == "chol_prof") {   // %globaltimer stamps of the last k_chol_tiles launch, as doubles (ns)
    if (!H->d_ct_prof) { H->err = "run with MCBA_CHOL_PROF=1"; return MCBA_ERR_INVALID; }
    n = (int64_t)CT_MAXT * 16;
    if (n_out) *n_out = n;
    if (out) { std::vector<long long> t(n); CU(cudaStreamSynchronize(H->stream)); CU(cudaMemcpy(t.data(), H->d_ct_prof, n * sizeof(long long), cudaMemcpyDeviceToHost)); for (int64_t i = 0; i < std::min(n, cap); ++i) out[i] = (double)t[i]; }
    return MCBA_OK;
  }
  else { H->err = "unknown buffer"; return MCBA_ERR_INVALID; }
  if (n_out) *n_out = n;
  if (out) { CU(cudaStreamSynchronize(H->stream)); CU(cudaMemcpy(out, src, std::min(n, cap) * sizeof(double), cudaMemcpyDeviceToHost)); }
  return MCBA_OK;
}

}  // extern "C"
