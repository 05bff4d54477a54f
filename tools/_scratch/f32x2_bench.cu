// Micro-benchmark: scalar FMUL/FADD chains vs packed FMUL2/FFMA2(x1.0) chains on sm_100a (issue-bound, registers only).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b, u64 one) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(one), "l"(b)); return r; }
constexpr int ITERS = 4096, ILP = 8;
__global__ void k_scalar(float* o, float m, float c) {
  float v[2 * ILP];
  for (int i = 0; i < 2 * ILP; ++i) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 2 * ILP; ++i) v[i] = __fadd_rn(__fmul_rn(v[i], m), c);
  float s = 0; for (int i = 0; i < 2 * ILP; ++i) s += v[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* o, float m, float c, u64 one) {
  u64 v[ILP]; const u64 M = pk(m, m), C = pk(c, c);
  for (int i = 0; i < ILP; ++i) v[i] = pk(threadIdx.x * 1e-3f + 2 * i, threadIdx.x * 1e-3f + 2 * i + 1);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = add2(mul2(v[i], M), C, one);
  float s = 0; for (int i = 0; i < ILP; ++i) { float a, b; upk(v[i], a, b); s += a; s += b; }
  o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* o; cudaMalloc(&o, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const u64 one = 0x3f8000003f800000ull;
  for (int rep = 0; rep < 3; ++rep) {
    float ms1, ms2;
    cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(o, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
    cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(o, 0.999f, 0.001f, one); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1);
    const double flops = 148.0 * 8 * 256 * ITERS * 2 * ILP * 2;
    printf("scalar FMUL+FADD %.3f ms (%.1f TFLOP/s)   packed FMUL2+FFMA2 %.3f ms (%.1f TFLOP/s)   ratio %.2f\n", ms1, flops / ms1 * 1e-9, ms2, flops / ms2 * 1e-9, ms1 / ms2);
  }
  float h[4]; cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost); printf("%g %s\n", h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
