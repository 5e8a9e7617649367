// Latency microbenchmarks behind the reduced-system Cholesky's pivot chain (B200): dependent DFMA, Newton reciprocal,
// library rsqrt / division, DMMA m8n8k4, shared-memory load, and the two in-register 8x8 factorisations.
#include <cstdio>
#include <cuda_runtime.h>
#include <cfloat>
__device__ __forceinline__ double ct_rcp(double d) {
  double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0); x = fma(x, e, x); e = fma(-d, x, 1.0); x = fma(x, e, x); return x;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void k_lat(double* out, long long* cyc, double seed, int n) {
  __shared__ double sm[64 * 68];
  for (int i = threadIdx.x; i < 64 * 68; i += blockDim.x) sm[i] = 1.0 + 1e-3 * i;
  __syncthreads();
  double x = seed + threadIdx.x * 1e-9, y = 0.5, c0 = 0, c1 = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) x = fma(x, y, 1e-3);
    if (MODE == 1) x = ct_rcp(x) + 1.5;
    if (MODE == 2) x = rsqrt(x) + 1.5;
    if (MODE == 3) x = 1.0 / x + 1.5;
    if (MODE == 4) { dmma884(c0, c1, x, y); }
    if (MODE == 5) x = sm[((int)x) & 63] + 1e-9;           // dependent LDS (address from value)
    if (MODE == 6) x = sqrt(x) + 1.5;
    if (MODE == 7) x = x * y;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + c0 + c1;
}
// 8x8 factorisations in registers, input from shared memory, result to shared memory; one warp, all lanes redundant
template <int VAR> __global__ void k_f8(double* out, long long* cyc, int reps) {
  __shared__ double T0[8 * 12], F8[64];
  if (threadIdx.x < 64) { const int r = threadIdx.x >> 3, c = threadIdx.x & 7; T0[r * 12 + c] = (r == c) ? 10.0 + r : 0.1 * (r + c + 1); }
  __syncthreads();
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double m[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) m[r][c] = T0[r * 12 + c];
    if (VAR == 0) {          // Cholesky form with library rsqrt (first version)
      double inv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double d = m[j][j]; inv[j] = rsqrt(d); m[j][j] = d * inv[j];
#pragma unroll
        for (int r = j + 1; r < 8; ++r) m[r][j] *= inv[j];
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
          for (int c = j + 1; c <= r; ++c) m[r][c] -= m[r][j] * m[c][j];
      }
    } else {                 // square-root-free with Newton reciprocal on the chain
      double dd[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double d = m[j][j]; dd[j] = d;
        const double rj = (VAR == 1) ? ct_rcp(d) : 1.0 / d;
        double tr[8];
        if (j < 7) m[j + 1][j + 1] = fma(-(m[j + 1][j] * m[j + 1][j]), rj, m[j + 1][j + 1]);
#pragma unroll
        for (int r = j + 1; r < 8; ++r) tr[r] = m[r][j] * rj;
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
          for (int c = j + 1; c <= r; ++c) if (!(r == j + 1 && c == j + 1)) m[r][c] = fma(-tr[r], m[c][j], m[r][c]);
#pragma unroll
        for (int r = j + 1; r < 8; ++r) m[r][j] = tr[r];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j][j] = rsqrt(dd[j]);
    }
    if (threadIdx.x == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) F8[r * 8 + c] = m[r][c];
    }
    __syncwarp();
    if (threadIdx.x < 8) T0[threadIdx.x * 12 + threadIdx.x] += F8[threadIdx.x * 9] * 1e-9;   // dependency between reps
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = F8[threadIdx.x & 63];
}
// lane-per-element variant: lane = 4 r + cq holds m[r][2cq], m[r][2cq+1]; per pivot one broadcast of d, the Newton
// reciprocal (redundant on all lanes), three shuffles for the column entries and two FMAs per lane
__global__ void k_f8_lanes(double* out, long long* cyc, int reps) {
  __shared__ double T0[8 * 12], F8[64];
  if (threadIdx.x < 64) { const int r = threadIdx.x >> 3, c = threadIdx.x & 7; T0[r * 12 + c] = (r == c) ? 10.0 + r : 0.1 * (r + c + 1); }
  __syncthreads();
  const int lane = threadIdx.x, r = lane >> 2, cq = lane & 3, c0 = 2 * cq, c1 = c0 + 1;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double m0 = T0[r * 12 + c0], m1 = T0[r * 12 + c1];
    double dj = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double colj = (j & 1) ? m1 : m0;                 // this lane's entry of column j if it holds one (cq == j/2)
      const double d = __shfl_sync(0xffffffffu, colj, j * 4 + (j >> 1));
      const double colr = __shfl_sync(0xffffffffu, colj, (lane & ~3) | (j >> 1));
      const double cc0 = __shfl_sync(0xffffffffu, colj, c0 * 4 + (j >> 1));
      const double cc1 = __shfl_sync(0xffffffffu, colj, c1 * 4 + (j >> 1));
      const double p0 = colr * cc0, p1 = colr * cc1;
      const double rj = ct_rcp(d);
      if (r > j && c0 > j && c0 <= r) m0 = fma(-p0, rj, m0);
      if (r > j && c1 > j && c1 <= r) m1 = fma(-p1, rj, m1);
      const double t = colr * rj;
      if (r > j && c0 == j) m0 = t;
      if (r > j && c1 == j) m1 = t;
      if (r == j) dj = d;
    }
    const double si = rsqrt(dj);
    if (c0 <= r) F8[r * 8 + c0] = (c0 == r) ? si : m0;
    if (c1 <= r) F8[r * 8 + c1] = (c1 == r) ? si : m1;
    __syncwarp();
    if (threadIdx.x < 8) T0[threadIdx.x * 12 + threadIdx.x] += F8[threadIdx.x * 9] * 1e-9;
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = F8[threadIdx.x & 63];
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"dfma", "newton_rcp(+add)", "rsqrt(+add)", "div(+add)", "dmma884 chain", "lds dependent", "sqrt(+add)", "dmul"};
  long long h; const int n = 4096;
#define RUN(M) k_lat<M><<<1, 32>>>(out, cyc, 1.25, n); k_lat<M><<<1, 32>>>(out, cyc, 1.25, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-20s %.1f cycles/op\n", names[M], (double)h / n);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
  const char* fn[] = {"8x8 cholesky rsqrt", "8x8 ldl newton-rcp", "8x8 ldl division"};
#define RUNF(V) k_f8<V><<<1, 32>>>(out, cyc, 64); k_f8<V><<<1, 32>>>(out, cyc, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-20s %.0f cycles per 8x8 block (1 warp)\n", fn[V], (double)h / 64);
  RUNF(0) RUNF(1) RUNF(2)
  k_f8_lanes<<<1, 32>>>(out, cyc, 64); k_f8_lanes<<<1, 32>>>(out, cyc, 64); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("8x8 ldl lane/element %.0f cycles per 8x8 block (1 warp)\n", (double)h / 64);
  { double hv[64], hw[64]; k_f8<1><<<1, 32>>>(out, cyc, 1); cudaMemcpy(hv, out, 64 * 8, cudaMemcpyDeviceToHost); k_f8_lanes<<<1, 32>>>(out, cyc, 1); cudaMemcpy(hw, out, 64 * 8, cudaMemcpyDeviceToHost); double md = 0; for (int r = 0; r < 8; ++r) for (int c = 0; c <= r; ++c) { double dd = hv[r * 8 + c] - hw[r * 8 + c]; if (dd < 0) dd = -dd; if (dd > md) md = dd; } printf("max diff lane/element vs redundant: %.3e\n", md); }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
