// Micro-benchmark: fp64 vector (DFMA) and fp64 tensor (DMMA m8n8k4) peak, plus an HBM copy,
// on one B200. Test infrastructure: gives the fp64 roofline denominator that MEASURED_PEAKS.json lacks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peak_fp64 peak_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0; 
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NT>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; i++) { c[i][0] = 0; c[i][1] = 0; }
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NT; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// DMMA with a 4x4 outer-product operand pattern (16 accumulator tiles, 4 A and 4 B fragment registers per k-step
// refreshed every step from shared memory), the operand pattern of k_syrk / k_observations.
__global__ void __launch_bounds__(128) dmma_outer_kernel(double* out, int iters) {
  __shared__ double sm[16 * 68];
  for (int i = threadIdx.x; i < 16 * 68; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double fa[4], fb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { fa[i] = sm[(4 * ks + q) * 68 + 8 * i + g]; fb[i] = sm[(4 * ks + q) * 68 + 32 + 8 * i + g]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(fa[i]), "d"(fb[j]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Same, but the fragment loads cannot be hoisted (row offset depends on the iteration), and a variant with a
// 4x8 warp tile (12 LDS per 32 DMMA instead of 8 per 16).
template <int NJ>
__global__ void __launch_bounds__(128) dmma_lds_kernel(double* out, int iters) {
  __shared__ double sm[64 * 68];
  for (int i = threadIdx.x; i < 64 * 68; i += blockDim.x) sm[i] = 1.0 + i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double c[4][NJ][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
    const double* base = sm + ((it & 3) * 16) * 68;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double fa[4], fb[NJ];
#pragma unroll
      for (int i = 0; i < 4; i++) fa[i] = base[(4 * ks + q) * 68 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < NJ; j++) fb[j] = base[(4 * ks + q) * 68 + ((32 + 8 * j) & 63) + g];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                       : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(fa[i]), "d"(fb[j]));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < NJ; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void write_kernel(double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = make_double2(1.0, 2.0);
}

template<typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\",\n", p.name, sms, p.major, p.minor);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
  {
    const int iters = 20000, ILP = 8;
    for (int bps : {4, 8}) {
      float ms = time_ms([&] { dfma_kernel<ILP><<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
      double flops = 2.0 * ILP * iters * 256.0 * sms * bps;
      printf(" \"dfma_tflops_bps%d\": %.3f,\n", bps, flops / ms * 1e-9);
    }
  }
  {
    const int iters = 5000;
    for (int bps : {2, 4}) {
      float ms = time_ms([&] { dmma_kernel<8><<<sms * bps, 256>>>(out, iters); }, 5);
      double flops = 2.0 * 256.0 * 8 * iters * 8.0 * sms * bps;  // 8x8x4 FMAs per warp-instr, 8 warps/CTA
      printf(" \"dmma_tflops_bps%d\": %.3f,\n", bps, flops / ms * 1e-9);
    }
  }
  {
    const int iters = 2000;
    for (int bps : {2, 4, 6}) {
      float ms = time_ms([&] { dmma_outer_kernel<<<sms * bps, 128>>>(out, iters); }, 5);
      double flops = 2.0 * 256.0 * 64 * iters * 4.0 * sms * bps;   // 64 DMMA per iteration per warp, 4 warps/CTA
      printf(" \"dmma_outer4x4_tflops_bps%d\": %.3f,\n", bps, flops / ms * 1e-9);
    }
  }
  {
    const int iters = 2000;
    for (int bps : {2, 3, 4}) {
      float ms = time_ms([&] { dmma_lds_kernel<4><<<sms * bps, 128>>>(out, iters); }, 5);
      printf(" \"dmma_lds_4x4_tflops_bps%d\": %.3f,\n", bps, 2.0 * 256.0 * 64 * iters * 4.0 * sms * bps / ms * 1e-9);
      ms = time_ms([&] { dmma_lds_kernel<8><<<sms * bps, 128>>>(out, iters); }, 5);
      printf(" \"dmma_lds_4x8_tflops_bps%d\": %.3f,\n", bps, 2.0 * 256.0 * 128 * iters * 4.0 * sms * bps / ms * 1e-9);
    }
  }
  {
    size_t n = (size_t)1 << 27;  // 2 GiB per buffer of double2
    double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
    CK(cudaMemset(a, 1, n * 16));
    float ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
    printf(" \"copy_gbs\": %.1f,\n", 2.0 * n * 16 / ms * 1e-6);
    ms = time_ms([&] { write_kernel<<<sms * 16, 512>>>(b, n); }, 5);
    printf(" \"write_gbs\": %.1f,\n", 1.0 * n * 16 / ms * 1e-6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); }, 5);
    printf(" \"memcpy_d2d_gbs\": %.1f\n}\n", 2.0 * n * 16 / ms * 1e-6);
  }
  return 0;
}
