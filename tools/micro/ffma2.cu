// Microbenchmark: packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.  Per thread 8 independent
// accumulator chains; reports FMAs per clock per SM for 1, 2, 4, 8 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long xv, av, bv;
        asm("mov.b64 %0, {%1, %2};" : "=l"(xv) : "f"(x[i]), "f"(x[i + 1]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xv) : "l"(av), "l"(bv));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x[i]), "=f"(x[i + 1]) : "l"(xv));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    for (int packed = 0; packed < 2; ++packed) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        if (packed) k<true><<<p.multiProcessorCount, threads>>>(out, iters, 1.0001f, 0.5f);
        else k<false><<<p.multiProcessorCount, threads>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
      }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      const double fmas = 16.0 * iters * threads;                       // per SM
      int clk = 0;
      cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
      printf("%4d threads/SM %-6s: %.3f ms, %.1f FMA/ns/SM (%.1f FMA per clock at %d MHz nominal)\n", threads,
             packed ? "FFMA2" : "FFMA", ms, fmas / (ms * 1e6), fmas / (ms * 1e6) / (clk / 1e6), clk / 1000);
    }
  }
  return 0;
}
