// Issue / pipe throughput of scalar FFMA against packed FFMA2 (fma.rn.f32x2) on sm_100a: 8 independent chains per thread, 1024 threads
// per CTA, 4 CTAs per SM worth of blocks.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__global__ void k_scalar(float* out, float m, float c, int iters) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], m, c);
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float m, float c, int iters) {
  f2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = pk(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
  const f2 mm = pk(m, m), cc = pk(c, c);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma2(x[i], mm, cc);
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 2, threads = 1024, iters = 4096;
  float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int variant = 0; variant < 2; ++variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (variant == 0) k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f, iters); else k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    const double fmas = (double)blocks * threads * 16.0 * iters;       // fp32 FMAs (lane-ops), the same in both variants
    printf("%s: %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)  %.1f fp32 FMA / clk / SM at 1.965 GHz\n", variant ? "FFMA2 (packed)" : "FFMA (scalar)", best, fmas / best / 1e9,
           2 * fmas / best / 1e9, fmas / (best * 1e-3) / sms / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
