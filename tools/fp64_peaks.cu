// FP64 pipe microbenchmarks on B200 ("FP64 peak is not in MEASURED_PEAKS.json -- measure it"):
// DFMA (CUDA-core) and DMMA (mma.sync.m8n8k4.f64, FP64 tensor core) throughput with independent chains, and the
// dependent-issue latency of each.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, double a, double b) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d,\n", p.name, sms, clk_khz / 1000);
  {  // throughput: 8 CTAs x 256 threads per SM, 8 independent chains
    const int blocks = sms * 8, threads = 256;
    float ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 8 * iters * (double)blocks * threads;
    printf(" \"dfma_tflops\": %.2f,\n", flops / ms / 1e9);
    ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    flops = 512.0 * 8 * iters * (double)blocks * (threads / 32);
    printf(" \"dmma_m8n8k4_tflops\": %.2f,\n", flops / ms / 1e9);
  }
  {  // latency: one warp, one dependent chain
    float ms = time_ms([&] { k_dfma<1><<<1, 32>>>(out, iters * 10, 1.0000001, 1e-9); });
    printf(" \"dfma_dependent_ns\": %.2f,\n", ms * 1e6 / (iters * 10));
    ms = time_ms([&] { k_dmma<1><<<1, 32>>>(out, iters * 10, 1.0000001, 1e-9); });
    printf(" \"dmma_dependent_ns\": %.2f,\n", ms * 1e6 / (iters * 10));
    // per-SM issue rate with 1..4 warps per SMSP, single chain per warp (what a latency-bound kernel sees)
    for (int w = 4; w <= 16; w *= 2) {
      ms = time_ms([&] { k_dmma<1><<<sms, 32 * w>>>(out, iters, 1.0000001, 1e-9); });
      printf(" \"dmma_ns_per_mma_per_sm_%dwarps_1chain\": %.3f,\n", w, ms * 1e6 / ((double)iters * w));
    }
  }
  printf(" \"note\": \"flops: DFMA = 2 per lane, DMMA m8n8k4 = 512 per warp instruction\"}\n");
  return 0;
}
