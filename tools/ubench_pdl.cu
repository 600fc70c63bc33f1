// development: what does programmatic dependent launch save per kernel boundary on this GPU?
// A chain of dependent streaming kernels (each reads the previous one's output), launched (a) plainly, (b) with the
// programmatic-stream-serialization attribute, griddepcontrol.wait as the first instruction and launch_dependents right after.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build_tmp/ubench_pdl tools/ubench_pdl.cu && ./build_tmp/ubench_pdl
#include <cstdio>
#include <cuda_runtime.h>
// PDL: 0 plain, 1 wait + trigger at the top, 2 wait only (implicit trigger when the grid completes), 3 wait + trigger after the loop
template <int PDL>
__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (PDL == 1) asm volatile("griddepcontrol.launch_dependents;");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = in[i]; v.x += 1.f; out[i] = v;
  }
  if (PDL == 3) asm volatile("griddepcontrol.launch_dependents;");
}
template <int PDL>
static void launch(const float4* in, float4* out, size_t n, int grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k_copy<PDL>, in, out, n);
}
int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int chain = 40;
  for (size_t mb : {1, 4, 16, 64, 256}) {
    const size_t n = mb * (1 << 20) / 16;
    float4 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 0, n * 16);
    for (int grid : {148 * 2, 148 * 8, 148 * 32}) {
      float ms[4];
      for (int pdl = 0; pdl < 4; ++pdl) {
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0, s);
          for (int i = 0; i < chain; ++i) {
            const float4* src = i & 1 ? b : a; float4* dst = i & 1 ? a : b;
            if (pdl == 0) launch<0>(src, dst, n, grid, s); else if (pdl == 1) launch<1>(src, dst, n, grid, s);
            else if (pdl == 2) launch<2>(src, dst, n, grid, s); else launch<3>(src, dst, n, grid, s);
          }
          cudaEventRecord(e1, s); cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms[pdl], e0, e1);
        }
      }
      printf("%4zu MB  grid %5d  us/kernel: plain %7.2f  early %7.2f  wait-only %7.2f  late %7.2f  (%s)\n", mb, grid, ms[0] * 1000 / chain,
             ms[1] * 1000 / chain, ms[2] * 1000 / chain, ms[3] * 1000 / chain, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(a); cudaFree(b);
  }
  return 0;
}
