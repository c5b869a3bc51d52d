// FP32 issue-rate microbenchmark for B200: scalar FFMA (3-register form) vs packed fma.rn.f32x2.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma ffma.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, NACC = 16;

__global__ void k_ffma(float* out, float a, float b) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long acc[NACC / 2], a2, b2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) {
    float x = threadIdx.x * 0.001f + i, y = x + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(x), "f"(y));
  }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < NACC / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a2), "l"(b2));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC / 2; ++i) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
    s += x + y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K> float run(K k, float* d, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, 256>>>(d, 1.0001f, 0.0001f);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<<<blocks, 256>>>(d, 1.0001f, 0.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int blocks = 148 * 8;
  float* d; cudaMalloc(&d, blocks * 256 * sizeof(float));
  double flops = 2.0 * ITER * NACC * 256.0 * blocks;
  float t1 = run(k_ffma, d, blocks), t2 = run(k_ffma2, d, blocks);
  printf("FFMA  (3-reg scalar): %.3f ms  %.1f TFLOP/s\n", t1, flops / t1 / 1e9);
  printf("FFMA2 (f32x2 packed): %.3f ms  %.1f TFLOP/s\n", t2, flops / t2 / 1e9);
  return 0;
}
