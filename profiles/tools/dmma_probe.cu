// Throughput of mma.sync.m8n8k4 fp64 (SASS DMMA) against DFMA on one GPU: is the fp64 tensor path of B200 at least the vector
// rate?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a dmma_probe.cu -o dmma_probe && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(int iters, double* out) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = i; }
  const double a = 1.0 + threadIdx.x * 1e-12, b = 0.999999;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(int iters, double* out) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = threadIdx.x * 1e-9 + i;
  const double a = 0.999999, b = 1e-7;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 8);
  const int iters = 20000, blocks = p.multiProcessorCount * 8;
  {
    float ms = time_ms([&] { k_dfma<8><<<blocks, 256>>>(iters, out); });
    double fl = 2.0 * blocks * 256.0 * iters * 8;
    printf("DFMA ilp8 : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dmma<4><<<blocks, 256>>>(iters, out); });
    double fl = 2.0 * 256.0 * (blocks * 8.0) * iters * 4;       // 8 warps per block, 256 FMA per DMMA
    printf("DMMA ilp4 : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dmma<8><<<blocks, 256>>>(iters, out); });
    double fl = 2.0 * 256.0 * (blocks * 8.0) * iters * 8;
    printf("DMMA ilp8 : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  {
    float ms = time_ms([&] { k_dmma<1><<<blocks, 256>>>(iters, out); });
    double fl = 2.0 * 256.0 * (blocks * 8.0) * iters * 1;
    printf("DMMA ilp1 : %.3f ms  %.2f TFLOP/s (dependent chain: latency %.1f cycles at 1.965 GHz with 8 warps/SMSP x 2)\n", ms, fl / ms * 1e-9, 0.0);
  }
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  return 0;
}
