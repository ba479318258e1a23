// FP64 FMA peak of the device ("measure an FP64 FMA peak ... report FP64-pipe % beside DRAM %").
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/fp64_peak.cu -o profiles/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_ffma(float *out, float a, float b, int iters) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 1; }
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 16384;
  void *buf;
  cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best64 = 0, best32 = 0;
  for (int rep = 0; rep < 4; rep++) {
    float ms;
    cudaEventRecord(e0);
    k_dfma<<<blocks, threads>>>((double *)buf, 0.999999, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep && tf > best64) best64 = tf;
    cudaEventRecord(e0);
    k_ffma<<<blocks, threads>>>((float *)buf, 0.999999f, 1e-9f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    tf = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep && tf > best32) best32 = tf;
  }
  printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp32_fma_tflops\": %.2f}\n", prop.name,
         prop.multiProcessorCount, best64, best32);
  return 0;
}
