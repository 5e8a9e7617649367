// mcba_chol_tiles.cuh — K3, dataflow variant: dense Cholesky + both triangular solves of the reduced system as ONE
// persistent kernel in which every 64x64 tile of the factor has its own CTA (replaces CHOLMOD's supernodal
// factorisation of the reduced camera/board system, SparseSchurComplementSolver in ceres-solver 2.2.0).
//
// S is (n+1) x n row-major: rows 0..n-1 the symmetric reduced matrix (lower part used), row n the right-hand side.
// NT = ceil(n / 64) tile columns; tile rows 0..NT-1 plus row NT = the rhs row. Tile (i, j), j <= i, is owned by one
// CTA for the whole kernel (left-looking):
//   1. acc = sum_{k<j} L(i,k) L(j,k)^T, accumulated with DMMA as the tiles of the columns to the left become ready
//      (per-tile ready flags in global memory, epoch-stamped so that nothing is reset between launches);
//   2. X = A(i,j) - acc. Diagonal tile: factor X = L L^T in shared memory (8-wide sub-panels: redundant in-register
//      8x8 factorisation per row thread, DMMA trailing update), publish L(j,j) and the inverses of its 8x8 diagonal
//      blocks. Other tiles: wait for L(j,j), X <- X L(j,j)^-T by blocked substitution (DMMA for the block products,
//      the 8x8 inverses for the diagonal blocks), publish. The rhs row rides along as tile row NT: forward substitution.
//   3. Backward substitution L^T z = y with the tiles still resident in shared memory: the owner of (i,j) contributes
//      L(i,j)^T z_i as soon as z_i is published, the owner of (j,j) sums the contributions in a fixed order and solves
//      its block.
// Nothing waits at a grid-wide barrier: the critical path is potrf(j) -> trsm(j+1,j) -> update of (j+1,j+1) -> potrf(j+1),
// about 16 x 7 us at n = 1020 instead of 16 x (redundant potrf + panel solve + 2 grid syncs + trailing update).
// All CTAs must be co-resident (cooperative launch); every wait is bounded (timeout -> *fail = 2).
// Summation orders are fixed: results are bit-reproducible.
#pragma once
#include "mcba_kernels.cuh"

namespace mcba {

constexpr int CT_B = 64, CT_LD = 68, CT_THREADS = 256, CT_MAXT = 512, CT_MAXNT = 30;
constexpr int CT_SMEM = (2 * CT_B * CT_LD + 512 + 64 + 64 + 64 + 4 * 64) * 8;
constexpr long long CT_TIMEOUT = 4000000000LL;

struct CholTilesArgs {
  double* S; int n; double* z; int* fail;
  unsigned* flags;        // [2 * CT_MAXT + 64]: tile ready | partial ready | z ready
  const unsigned* epoch_p; // launch counter in device memory, bumped by k_bump_epoch before every launch (graph replay)
  double* dinv;           // [CT_MAXNT][512] inverses of the 8x8 diagonal blocks of L(j,j)
  double* pbuf;           // [CT_MAXT][64] backward-substitution contributions
  long long* prof;        // optional [CT_MAXT][8] %globaltimer stamps per tile (MCBA_CHOL_PROF)
};

__device__ __forceinline__ void ct_post(unsigned* f, unsigned epoch) {
  __threadfence();
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__device__ __forceinline__ bool ct_wait(const unsigned* f, unsigned epoch, int* fail) {
  const long long t0 = clock64();
  unsigned v;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == epoch) return true;
    if (clock64() - t0 > CT_TIMEOUT) { atomicExch(fail, 2); return false; }
  }
}
__device__ __forceinline__ long long ct_now() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CT_STAMP(k) do { if (a.prof && tid == 0) a.prof[(size_t)my_tile * 16 + (k)] = ct_now(); } while (0)
__host__ __device__ inline int ct_tile_id(int i, int j, int NT) { return j * (NT + 1) - j * (j - 1) / 2 + (i - j); }
__host__ __device__ inline int ct_num_tiles(int n) { const int NT = (n + CT_B - 1) / CT_B; return NT * (NT + 3) / 2; }

__global__ void k_bump_epoch(unsigned* e) { *e += 1u; }

__global__ void __launch_bounds__(CT_THREADS, 2) k_chol_tiles(CholTilesArgs a) {
  extern __shared__ __align__(16) double sm[];
  double* T0 = sm;                       // [64][68] operand A of the updates, then this CTA's own tile X
  double* T1 = T0 + CT_B * CT_LD;        // [64][68] operand B of the updates, then L(j,j)
  double* Dinv = T1 + CT_B * CT_LD;      // [8][8][8] inverses of the 8x8 diagonal blocks of L(j,j)
  double* wv = Dinv + 512;               // [64]
  double* zv = wv + 64;                  // [64]
  double* invd = zv + 64;                // [64]
  double* red = invd + 64;               // [4][64]
  const int n = a.n, NT = (n + CT_B - 1) / CT_B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  int tj = 0, rem = blockIdx.x;
  while (tj < NT && rem >= NT + 1 - tj) { rem -= NT + 1 - tj; ++tj; }
  if (tj >= NT) return;
  const int ti = tj + rem;                                   // tj <= ti <= NT
  const bool is_diag = ti == tj, is_rhs = ti == NT;
  const int row0 = is_rhs ? n : ti * CT_B, col0 = tj * CT_B;
  const int rv = is_rhs ? 1 : min(CT_B, n - row0);           // valid rows / columns of this tile
  const int cv = min(CT_B, n - col0);
  double* S = a.S;
  unsigned* ready = a.flags; unsigned* pready = a.flags + CT_MAXT; unsigned* zready = a.flags + 2 * CT_MAXT;
  const unsigned epoch = *a.epoch_p;
  const int my_tile = ct_tile_id(ti, tj, NT);
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
  // this tile of A, in accumulator-fragment layout (issued now, consumed after the updates)
  double orig[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = wm + 8 * i + g, c = wn + 8 * j + 2 * q + h;
        orig[i][j][h] = (r < rv && c < cv) ? S[(size_t)(row0 + r) * n + col0 + c] : 0.0;
      }
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  bool alive = true;
  CT_STAMP(0);
  // ---- 1. left-looking updates
  for (int k = 0; k < tj; ++k) {
    if (tid == 0) alive = ct_wait(ready + ct_tile_id(ti, k, NT), epoch, a.fail);
    if (tid == 32 && !is_diag) alive = ct_wait(ready + ct_tile_id(tj, k, NT), epoch, a.fail);
    __syncthreads();                                         // also: everybody is done with the previous operands
    {
      double va[16], vb[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int e = tid + it * CT_THREADS, r = e >> 6, c = e & 63;
        va[it] = (r < rv) ? __ldcg(S + (size_t)(row0 + r) * n + k * CT_B + c) : 0.0;
        vb[it] = (!is_diag && r < cv) ? __ldcg(S + (size_t)(col0 + r) * n + k * CT_B + c) : 0.0;
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int e = tid + it * CT_THREADS, r = e >> 6, c = e & 63;
        T0[r * CT_LD + c] = va[it];
        if (!is_diag) T1[r * CT_LD + c] = vb[it];
      }
    }
    __syncthreads();
    const double* Bt = is_diag ? T0 : T1;
#pragma unroll 4
    for (int ks = 0; ks < CT_B / 4; ++ks) {
      double fa[4], fb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[i] = T0[(wm + 8 * i + g) * CT_LD + 4 * ks + q];
#pragma unroll
      for (int j = 0; j < 2; ++j) fb[j] = Bt[(wn + 8 * j + g) * CT_LD + 4 * ks + q];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  }
  __syncthreads();
  CT_STAMP(1);
  // ---- 2. X = A - acc into T0 (identity padding beyond the valid part of a diagonal tile)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = wm + 8 * i + g, c = wn + 8 * j + 2 * q + h;
        double v = orig[i][j][h] - acc[i][j][h];
        if (is_diag && (r >= cv || c >= cv)) v = (r == c) ? 1.0 : 0.0;
        T0[r * CT_LD + c] = v;
      }
  __syncthreads();
  if (is_diag) {
    // ---- potrf of the 64x64 block: D = T0 (lower part), factor -> T1, right-looking in 8-wide sub-panels
    for (int c0 = 0; c0 < CT_B; c0 += 8) {
      if (tid < CT_B - c0 || (tid >= 64 && tid < 72)) {
        double l[8][8], inv[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) l[r][c] = T0[(c0 + r) * CT_LD + c0 + c];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double d = l[j][j];
          const bool b = !(d > 0.0) || !(d < DBL_MAX);
          bad = bad || b;
          inv[j] = b ? 1.0 : rsqrt(d);
          l[j][j] = d * inv[j];
#pragma unroll
          for (int r = j + 1; r < 8; ++r) l[r][j] *= inv[j];
#pragma unroll
          for (int r = j + 1; r < 8; ++r)
#pragma unroll
            for (int c = j + 1; c <= r; ++c) l[r][c] -= l[r][j] * l[c][j];
        }
        if (tid < CT_B - c0) {
          const int i = c0 + tid;
          double x[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double t = T0[i * CT_LD + c0 + c];
#pragma unroll
            for (int k = 0; k < c; ++k) t -= x[k] * l[c][k];
            x[c] = (c0 + c <= i) ? t * inv[c] : 0.0;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) T1[i * CT_LD + c0 + c] = x[c];
          if (tid == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) invd[c0 + c] = inv[c];
            if (bad) atomicExch(a.fail, 1);
          }
        } else {
          const int cc = tid - 64;                            // column cc of the inverse of the 8x8 block
          double y[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            double t = (r == cc) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < r; ++k) t -= l[r][k] * y[k];
            y[r] = t * inv[r];
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) Dinv[(c0 >> 3) * 64 + r * 8 + cc] = y[r];
        }
      }
      if (c0 == 0 && tid >= 72) {
        // rows above the first sub-panel's reach: clear the strictly upper part of the factor once
        for (int e = tid - 72; e < CT_B * CT_B; e += CT_THREADS - 72) { const int r = e >> 6, c = e & 63; if (c > r && c >= 8) T1[r * CT_LD + c] = 0.0; }
      }
      __syncthreads();
      const int m = 7 - (c0 >> 3);                            // trailing 8x8 tiles per side
      const int ntile = m * (m + 1) / 2;
      for (int t = warp; t < ntile; t += CT_THREADS / 32) {
        int rt = 0, rr = t;
        while (rr >= rt + 1) { rr -= rt + 1; ++rt; }
        const int R = c0 + 8 + 8 * rt, Cc = c0 + 8 + 8 * rr;
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int s = 0; s < 2; ++s)
          dmma884(d0, d1, T1[(R + g) * CT_LD + c0 + 4 * s + q], T1[(Cc + g) * CT_LD + c0 + 4 * s + q]);
        double2* p = reinterpret_cast<double2*>(T0 + (R + g) * CT_LD + Cc + 2 * q);
        double2 cur = *p;
        cur.x -= d0; cur.y -= d1;
        *p = cur;
      }
      __syncthreads();
    }
    CT_STAMP(2);
    // publish L(j,j) (in place in S) and the 8x8 inverses
    for (int e = tid; e < CT_B * CT_B; e += CT_THREADS) {
      const int r = e >> 6, c = e & 63;
      if (r < cv && c <= r) S[(size_t)(row0 + r) * n + col0 + c] = T1[r * CT_LD + c];
    }
    for (int e = tid; e < 512; e += CT_THREADS) a.dinv[(size_t)tj * 512 + e] = Dinv[e];
    __threadfence();
    __syncthreads();
    if (tid == 0) ct_post(ready + my_tile, epoch);
    CT_STAMP(3);
  } else {
    // ---- trsm: X <- X L(j,j)^-T. Each warp owns 8 rows and walks the 8 column blocks; no CTA-wide barrier inside.
    if (tid == 0) alive = ct_wait(ready + ct_tile_id(tj, tj, NT), epoch, a.fail);
    __syncthreads();
    CT_STAMP(4);
    {
      double vb[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int e = tid + it * CT_THREADS, r = e >> 6, c = e & 63;
        vb[it] = (r < cv && c <= r) ? __ldcg(S + (size_t)(col0 + r) * n + col0 + c) : 0.0;
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) { const int e = tid + it * CT_THREADS; T1[(e >> 6) * CT_LD + (e & 63)] = vb[it]; }
      for (int e = tid; e < 512; e += CT_THREADS) Dinv[e] = __ldcg(a.dinv + (size_t)tj * 512 + e);
    }
    __syncthreads();
    CT_STAMP(5);
    const int R = 8 * warp;
    for (int b = 0; b < 8; ++b) {
      double c0a = 0.0, c1a = 0.0, c0b = 0.0, c1b = 0.0;      // two accumulators: half the dependent DMMA chain
      for (int s = 0; s < 2 * b; s += 2) {
        dmma884(c0a, c1a, T0[(R + g) * CT_LD + 4 * s + q], T1[(8 * b + g) * CT_LD + 4 * s + q]);
        dmma884(c0b, c1b, T0[(R + g) * CT_LD + 4 * s + 4 + q], T1[(8 * b + g) * CT_LD + 4 * s + 4 + q]);
      }
      const double2 xb = *reinterpret_cast<const double2*>(T0 + (R + g) * CT_LD + 8 * b + 2 * q);
      const double w0 = xb.x - (c0a + c0b), w1 = xb.y - (c1a + c1b);   // W[g][2q], W[g][2q+1]
      // W as an A operand: lane (g, q) needs W[g][q] and W[g][4+q]
      const int src0 = (lane & ~3) | (q >> 1), src1 = src0 + 2;
      const double a0x = __shfl_sync(0xffffffffu, w0, src0), a0y = __shfl_sync(0xffffffffu, w1, src0);
      const double a1x = __shfl_sync(0xffffffffu, w0, src1), a1y = __shfl_sync(0xffffffffu, w1, src1);
      const double fa0 = (q & 1) ? a0y : a0x, fa1 = (q & 1) ? a1y : a1x;
      double x0 = 0.0, x1 = 0.0;                              // X_b = W Linv_bb^T
      dmma884(x0, x1, fa0, Dinv[b * 64 + g * 8 + q]);
      dmma884(x0, x1, fa1, Dinv[b * 64 + g * 8 + 4 + q]);
      *reinterpret_cast<double2*>(T0 + (R + g) * CT_LD + 8 * b + 2 * q) = make_double2(x0, x1);
      __syncwarp();
    }
    __syncthreads();
    CT_STAMP(2);
    for (int e = tid; e < CT_B * CT_B; e += CT_THREADS) {
      const int r = e >> 6, c = e & 63;
      if (r < rv && c < cv) S[(size_t)(row0 + r) * n + col0 + c] = T0[r * CT_LD + c];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) ct_post(ready + my_tile, epoch);
    CT_STAMP(3);
  }
  if (is_rhs) return;
  // ---- 3. backward substitution
  if (!is_diag) {
    if (ti == tj + 1) return;            // the owner of (j,j) applies L(j+1,j)^T z_{j+1} itself from its own copy of the tile
    // contribution of this tile: p = L(i,j)^T z_i
    if (tid == 0) alive = ct_wait(zready + ti, epoch, a.fail);
    __syncthreads();
    CT_STAMP(6);
    if (tid < CT_B) zv[tid] = (tid < rv) ? __ldcg(a.z + row0 + tid) : 0.0;
    __syncthreads();
    {
      const int c = tid & 63, kq = tid >> 6;
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) { const int rr = kq * 16 + r; s += (rr < rv) ? T0[rr * CT_LD + c] * zv[rr] : 0.0; }
      red[kq * 64 + c] = s;
    }
    __syncthreads();
    if (tid < CT_B) a.pbuf[(size_t)my_tile * 64 + tid] = (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
    __threadfence();
    __syncthreads();
    if (tid == 0) ct_post(pready + my_tile, epoch);
    CT_STAMP(7);
    return;
  }
  // diagonal owner: w = y_j - sum_{i>j} L(i,j)^T z_i in the fixed order i = NT-1 .. j+1, then L(j,j)^T z_j = w.
  // The contributions of the tiles further down arrive early (their z_i are long published when z_{j+1} is); the one of
  // the sub-diagonal tile is the critical one: this CTA keeps its own copy of L(j+1,j) in T0 (loaded while idle) and
  // applies it as soon as z_{j+1} is visible - no second hop through another CTA.
  const int rv1 = (tj + 1 < NT) ? min(CT_B, n - (tj + 1) * CT_B) : 0;
  if (tj + 1 < NT) {
    if (tid == 0) alive = ct_wait(ready + ct_tile_id(tj + 1, tj, NT), epoch, a.fail);
    __syncthreads();
    double vb[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int e = tid + it * CT_THREADS, r = e >> 6, c = e & 63;
      vb[it] = (r < rv1 && c < cv) ? __ldcg(S + (size_t)((tj + 1) * CT_B + r) * n + col0 + c) : 0.0;
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) { const int e = tid + it * CT_THREADS; T0[(e >> 6) * CT_LD + (e & 63)] = vb[it]; }
  }
  if (tid == 0) {
    alive = ct_wait(ready + ct_tile_id(NT, tj, NT), epoch, a.fail);
    for (int i = NT - 1; i >= tj + 2; --i) alive = ct_wait(pready + ct_tile_id(i, tj, NT), epoch, a.fail) && alive;
  }
  __syncthreads();
  double w = 0.0;
  if (tid < CT_B) {
    w = (tid < cv) ? __ldcg(S + (size_t)n * n + col0 + tid) : 0.0;
    double pv[8];
#pragma unroll 1
    for (int i0 = NT - 1; i0 >= tj + 2; i0 -= 8) {            // loads of up to 8 contributions in flight
#pragma unroll
      for (int u = 0; u < 8; ++u) if (i0 - u >= tj + 2) pv[u] = __ldcg(a.pbuf + (size_t)ct_tile_id(i0 - u, tj, NT) * 64 + tid);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (i0 - u >= tj + 2) w -= pv[u];
    }
  }
  if (tj + 1 < NT) {
    if (tid == 0) alive = ct_wait(zready + tj + 1, epoch, a.fail);
    __syncthreads();
    if (tid < CT_B) zv[tid] = (tid < rv1) ? __ldcg(a.z + (tj + 1) * CT_B + tid) : 0.0;
    __syncthreads();
    {
      const int c = tid & 63, kq = tid >> 6;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int r = 0; r < 16; r += 2) {
        const int rr = kq * 16 + r;
        s0 = fma(T0[rr * CT_LD + c], zv[rr], s0);
        s1 = fma(T0[(rr + 1) * CT_LD + c], zv[rr + 1], s1);
      }
      red[kq * 64 + c] = s0 + s1;
    }
    __syncthreads();
    if (tid < CT_B) w -= (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
  }
  if (tid < CT_B) wv[tid] = w;
  __syncthreads();
  CT_STAMP(6);
  if (warp == 0) {
    for (int b = 7; b >= 0; --b) {
      if (lane < 8) {                                         // z_b = Linv_bb^T w_b
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          s0 = fma((k >= lane) ? Dinv[b * 64 + k * 8 + lane] : 0.0, wv[8 * b + k], s0);
          s1 = fma((k + 1 >= lane) ? Dinv[b * 64 + (k + 1) * 8 + lane] : 0.0, wv[8 * b + k + 1], s1);
        }
        zv[8 * b + lane] = s0 + s1;
      }
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = lane + 32 * half;
        if (c < 8 * b) {
          double s0 = wv[c], s1 = 0.0;
#pragma unroll
          for (int r = 0; r < 8; r += 2) {
            s0 = fma(-T1[(8 * b + r) * CT_LD + c], zv[8 * b + r], s0);
            s1 = fma(-T1[(8 * b + r + 1) * CT_LD + c], zv[8 * b + r + 1], s1);
          }
          wv[c] = s0 + s1;
        }
      }
      __syncwarp();
    }
    const bool failed = *((volatile int*)a.fail) != 0;
    for (int c = lane; c < cv; c += 32) a.z[col0 + c] = failed ? 0.0 : zv[c];
    __threadfence();
    __syncwarp();
    if (lane == 0) ct_post(zready + tj, epoch);
    CT_STAMP(7);
  }
  (void)alive;
}

}  // namespace mcba
