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
