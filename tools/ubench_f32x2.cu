// micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a (not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 1.0001f, c = 0.5f;
  if (MODE == 0) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
      }
    }
  } else {
    unsigned long long p0, p1, p2, p3, mm, cc;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(mm) : "f"(m), "f"(m));
    asm("mov.b64 %0, {%1,%2};" : "=l"(cc) : "f"(c), "f"(c));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(mm), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(mm), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(mm), "l"(cc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(mm), "l"(cc));
      }
    }
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters); else k<1><<<148 * 2, 1024>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 2 * 1024 * (double)iters * 64;
      printf("mode %d (%s): %.3f ms  %.2f TFMA/s (%.1f TFLOP/s) err=%s\n", mode, mode ? "fma.rn.f32x2" : "scalar fmaf", ms,
             fma / ms / 1e9, 2 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
