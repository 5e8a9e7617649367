// mcba_syrk.cuh — K2d, second generation: the Schur complement sum_o Z_o^T Z_o as a block-sparse, split-K SYRK
// (what SchurEliminator::Eliminate accumulates chunk by chunk in ceres-solver 2.2.0, as one tall-skinny product).
//
// Z is [6 n_pose][ldz] row-major with the padded column layout of FLayout: every 8-column DMMA tile belongs to exactly
// one camera / board block (or to the rhs block). The 6 rows of e-block o are non-zero only in the tiles of the blocks
// that e-block observes: `tile_mask[o]` (one bit per 8-column tile, built once per stage on the host from the
// board-observation table) says which. A 128x128 output tile per CTA as before; per k-step (4 rows of Z) a warp issues
// the DMMA of an 8x8 sub-tile pair only if the product can be non-zero for one of the (at most two) e-blocks the four
// rows belong to. 12 rows = 2 e-blocks = 3 k-steps: [o rows 0-3] [o rows 4-5, o+1 rows 0-1] [o+1 rows 2-5].
// For the 64-camera dome (a frame sees 37 of 64 cameras) that removes about 55 % of the DMMA issue slots.
//
// Data movement: one producer warp streams 24-row chunks (4 e-blocks) of the two 128-column strips into a 4-stage
// shared-memory ring with bulk asynchronous copies (cp.async.bulk global -> shared, the TMA engine; one 1 KB copy per
// row so that the rows keep their bank-conflict-free 132-double pitch) completing on an mbarrier per stage; the eight
// consumer warps wait on "full", run their k-steps and arrive on "empty". No CTA-wide barrier in the main loop.
#pragma once
#include "mcba_kernels.cuh"

namespace mcba {

constexpr int SY2_K = 24, SY2_STAGES = 4, SY2_CONSUMERS = 8, SY2_THREADS = (SY2_CONSUMERS + 1) * 32;
constexpr int SY2_STAGE_DOUBLES = 2 * SY2_K * SY_LD;
constexpr int SY2_SMEM = SY2_STAGES * SY2_STAGE_DOUBLES * 8 + SY2_STAGES * 4 * 8 + 2 * SY2_STAGES * 8;

__device__ __forceinline__ unsigned sm_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(sm_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sm_addr(dst)), "l"(src), "r"(bytes), "r"(sm_addr(bar)) : "memory");
}

// Four DMMAs (a 2 x 2 group of 8x8 sub-tiles) behind ONE warp-uniform branch. `go` is the same on every lane (it comes
// from activity words all lanes read from the same shared-memory address); bra.uni tells ptxas so: it keeps a real
// branch instead of predicating every mma.sync (which costs a WARPSYNC + a register shuffle per instruction).
__device__ __forceinline__ void dmma_group4(double& a0, double& a1, double& b0, double& b1, double& c0, double& c1,
                                            double& d0, double& d1, double fa0, double fa1, double fb0, double fb1, unsigned go) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, %12, 0;\n\t"
      "@p bra.uni DMMA_SKIP;\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%8}, {%10}, {%0,%1};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {%8}, {%11}, {%2,%3};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%4,%5}, {%9}, {%10}, {%4,%5};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%6,%7}, {%9}, {%11}, {%6,%7};\n\t"
      "DMMA_SKIP:\n\t"
      "}\n"
      : "+d"(a0), "+d"(a1), "+d"(b0), "+d"(b1), "+d"(c0), "+d"(c1), "+d"(d0), "+d"(d1)
      : "d"(fa0), "d"(fa1), "d"(fb0), "d"(fb1), "r"(go));
}

// tile_mask: [n_pose][mask_words] one bit per 8-column tile of Z (bit t of word t/32)
template <bool SPARSE>
__global__ void __launch_bounds__(SY2_THREADS, 1) k_syrk2(const double* __restrict__ Z, int64_t n_rows, int ldz, int nt,
                                                          int64_t rows_per_split, const unsigned* __restrict__ tile_mask,
                                                          int mask_words, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sy_sm[];
  unsigned* smask = reinterpret_cast<unsigned*>(sy_sm + SY2_STAGES * SY2_STAGE_DOUBLES);       // [stage][4 e-blocks][2]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smask + SY2_STAGES * 8);    // [stage]
  unsigned long long* empty = full + SY2_STAGES;
  int ti = 0, rem = blockIdx.x;   // tile pair from linear index (ti <= tj)
  while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
  const int tj = ti + rem;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)split * rows_per_split, r1 = min(n_rows, r0 + rows_per_split);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool diag = (ti == tj);
  const int n_chunks = (r1 > r0) ? (int)((r1 - r0 + SY2_K - 1) / SY2_K) : 0;
  // the ring starts zeroed: the bulk copies only ever write the valid columns of a row
  for (int e = tid; e < SY2_STAGES * SY2_STAGE_DOUBLES; e += SY2_THREADS) sy_sm[e] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < SY2_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, SY2_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros before async-proxy writes
  __syncthreads();
  if (warp == SY2_CONSUMERS) {
    // ---- producer warp
    const int wa = min(SY_T, ldz - ti * SY_T), wb = min(SY_T, ldz - tj * SY_T);      // valid columns of the two strips
    const unsigned bytes_a = (unsigned)wa * 8u, bytes_b = diag ? 0u : (unsigned)wb * 8u;
    const int mw_r = (16 * ti) >> 5, ms_r = (16 * ti) & 31, mw_c = (16 * tj) >> 5, ms_c = (16 * tj) & 31;
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % SY2_STAGES;
      mbar_wait(empty + s, ((c / SY2_STAGES) & 1) ^ 1);
      double* As = sy_sm + (size_t)s * SY2_STAGE_DOUBLES;
      double* Bs = As + SY2_K * SY_LD;
      const int64_t rbase = r0 + (int64_t)c * SY2_K;
      const int valid = (int)min((int64_t)SY2_K, r1 - rbase);
      if (SPARSE && lane < 4) {                                // activity of the chunk's four e-blocks for this CTA's strips
        const int64_t o = rbase / 6 + lane;
        unsigned mr = 0, mc = 0;
        if (o * 6 < r1) {
          mr = (tile_mask[o * mask_words + mw_r] >> ms_r) & 0xFFFFu;
          mc = (tile_mask[o * mask_words + mw_c] >> ms_c) & 0xFFFFu;
        }
        smask[s * 8 + lane * 2] = mr; smask[s * 8 + lane * 2 + 1] = mc;
      }
      if (valid < SY2_K && lane >= valid && lane < SY2_K) {     // ragged last chunk: rows past the end must read as zero
        for (int e = 0; e < SY_T; ++e) { As[lane * SY_LD + e] = 0.0; if (!diag) Bs[lane * SY_LD + e] = 0.0; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(full + s, (unsigned)valid * (bytes_a + bytes_b));
      __syncwarp();
      if (lane < valid) {
        const double* src = Z + (rbase + lane) * ldz;
        bulk_g2s(As + lane * SY_LD, src + ti * SY_T, bytes_a, full + s);
        if (!diag) bulk_g2s(Bs + lane * SY_LD, src + tj * SY_T, bytes_b, full + s);
      }
    }
    return;
  }
  // ---- consumer warps: 4 x 2 layout, 32 x 64 per warp
  const int g = lane >> 2, q = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
  const bool skip = diag && (wm >= wn + 64);   // warp tile strictly below the diagonal of a diagonal tile
  const int sh_r = (warp >> 1) * 4, sh_c = (warp & 1) * 8;
  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (int c = 0; c < n_chunks; ++c) {
    const int s = c % SY2_STAGES;
    mbar_wait(full + s, (c / SY2_STAGES) & 1);
    if (!skip) {
      const double* As = sy_sm + (size_t)s * SY2_STAGE_DOUBLES;
      const double* Bp = diag ? As : As + SY2_K * SY_LD;
      if (!SPARSE) {
#pragma unroll
        for (int ks = 0; ks < SY2_K / 4; ++ks) {
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
      } else {
#pragma unroll
      for (int p = 0; p < 2; ++p) {                            // two pairs of e-blocks per chunk
        const unsigned ra0 = (smask[s * 8 + (2 * p) * 2] >> sh_r) & 0xFu, cb0 = (smask[s * 8 + (2 * p) * 2 + 1] >> sh_c) & 0xFFu;
        const unsigned ra1 = (smask[s * 8 + (2 * p + 1) * 2] >> sh_r) & 0xFu, cb1 = (smask[s * 8 + (2 * p + 1) * 2 + 1] >> sh_c) & 0xFFu;
        unsigned a0 = 0, a1 = 0;                               // bit 8 i + j: sub-tile (i, j) can be non-zero for e-block 0 / 1
#pragma unroll
        for (int i = 0; i < 4; ++i) { a0 |= ((ra0 >> i) & 1u) ? (cb0 << (8 * i)) : 0u; a1 |= ((ra1 >> i) & 1u) ? (cb1 << (8 * i)) : 0u; }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const unsigned act = (kk == 0) ? a0 : (kk == 1 ? (a0 | a1) : a1);
          if (act == 0u) continue;
          const int ks = 3 * p + kk;
          const unsigned rany = (kk == 0) ? ra0 : (kk == 1 ? (ra0 | ra1) : ra1);
          const unsigned cany = (kk == 0) ? cb0 : (kk == 1 ? (cb0 | cb1) : cb1);
          double fa[4], fb[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) fa[i] = ((rany >> i) & 1u) ? As[(4 * ks + q) * SY_LD + wm + 8 * i + g] : 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) fb[j] = ((cany >> j) & 1u) ? Bp[(4 * ks + q) * SY_LD + wn + 8 * j + g] : 0.0;
          // one warp-uniform branch per group of 2 x 2 sub-tiles (= camera block x camera block at stage 5)
#pragma unroll
          for (int rc = 0; rc < 2; ++rc)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const unsigned m4 = (0x3u << (16 * rc + 2 * cc)) | (0x3u << (16 * rc + 8 + 2 * cc));
              dmma_group4(acc[2 * rc][2 * cc][0], acc[2 * rc][2 * cc][1], acc[2 * rc][2 * cc + 1][0], acc[2 * rc][2 * cc + 1][1],
                          acc[2 * rc + 1][2 * cc][0], acc[2 * rc + 1][2 * cc][1], acc[2 * rc + 1][2 * cc + 1][0], acc[2 * rc + 1][2 * cc + 1][1],
                          fa[2 * rc], fa[2 * rc + 1], fb[2 * cc], fb[2 * cc + 1], act & m4);
            }
        }
      }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  double* out = partial + ((size_t)split * gridDim.x + blockIdx.x) * (SY_T * SY_T);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<double2*>(out + (wm + 8 * i + g) * SY_T + wn + 8 * j + 2 * q) = make_double2(acc[i][j][0], acc[i][j][1]);
}

}  // namespace mcba
