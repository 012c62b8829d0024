// fp64 peak probe for the compute-bound smoother kernels (K6/K7): DFMA pipe vs DMMA
// (mma.sync.m8n8k4.f64) throughput on this GPU.  Build: nvcc -arch=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters) {
  double a[16];
  const double x = 1.0000001, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double *out, int iters) {
  double c[8][2];
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8, iters = 20000;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double fl_fma = 2.0 * 16 * iters * (double)blocks * threads;
  const double tf_fma = fl_fma / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  const double fl_mma = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
  const double tf_mma = fl_mma / (ms * 1e-3) / 1e12;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.2f, \"fp64_dmma_m8n8k4_tflops\": %.2f}\n", p.name, sms, tf_fma, tf_mma);
  return 0;
}
