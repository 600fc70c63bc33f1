// micro-benchmark (not part of the product): data movement skeleton of a SINGLE-SWEEP column FFT of length R = Ra * Rb
// done by a thread-block cluster of K CTAs with the four-step exchange over distributed shared memory.
//   load : CTA j pulls rows {Rb*a + b : a < Ra, b in its Rb/K slice} x W columns of both planes with ONE 3-D TMA box each
//   work : `passes` read+write sweeps over the tile in shared memory (stand-in for the FFT stages)
//   xchg : element (ka, b) goes to the CTA that owns ka (st.shared::cluster), held in registers across a cluster barrier
//   work : `passes` sweeps again, then ONE 3-D TMA store per plane of rows {Rb*ka + kb : ka in its Ra/K slice, kb < Rb}
// Reports algorithmic GB/s = (read + write of both planes) / time.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_cluster tools/ubench_cluster_fft_skeleton.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args {
  int Ra, Rb, K, W;       // R = Ra*Rb, cluster size, columns per tile
  int nbA, naB;           // Rb/K, Ra/K
  int passes;
};

template <int NPT, int T>
__global__ void __launch_bounds__(T) k_skeleton(const __grid_constant__ CUtensorMap ld_re, const __grid_constant__ CUtensorMap ld_im,
                                                const __grid_constant__ CUtensorMap st_re, const __grid_constant__ CUtensorMap st_im,
                                                const __grid_constant__ Args a) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const int W = a.W;
  const int tile_floats = a.Ra * a.nbA * W;          // per plane
  float* s_re = smem;
  float* s_im = smem + tile_floats;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tile = blockIdx.x / a.K;
  const int tid = threadIdx.x;
  const int lanes = W / 2;                            // column pairs
  const int c = tid % lanes, s = tid / lanes, S = T / lanes;
  const uint32_t bytes = (uint32_t)tile_floats * 4u;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(2u * bytes) : "memory");
    const int c0 = tile * W, c1 = (int)rank * a.nbA, c2 = 0;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(s_re)), "l"(&ld_re), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&mbar)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(s_im)), "l"(&ld_im), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&mbar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  float2 vr[NPT], vi[NPT];
  // ---- stand-in for the stages of sweep A: read everything, barrier, write back (in place, like the real stages)
  for (int p = 0; p < a.passes; ++p) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int e = s + i * S;
      vr[i] = *reinterpret_cast<const float2*>(s_re + e * W + 2 * c);
      vi[i] = *reinterpret_cast<const float2*>(s_im + e * W + 2 * c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int e = s + i * S;
      *reinterpret_cast<float2*>(s_re + e * W + 2 * c) = make_float2(vr[i].x + vi[i].y, vr[i].y - vi[i].x);
      *reinterpret_cast<float2*>(s_im + e * W + 2 * c) = make_float2(vi[i].x * 0.5f, vi[i].y * 0.5f);
    }
    __syncthreads();
  }
  // ---- exchange: hold the tile in registers, barrier, scatter to the owners
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int e = s + i * S;
    vr[i] = *reinterpret_cast<const float2*>(s_re + e * W + 2 * c);
    vi[i] = *reinterpret_cast<const float2*>(s_im + e * W + 2 * c);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int s_ka = s / a.nbA, bl = s - s_ka * a.nbA, ka_step = S / a.nbA;     // S is a multiple of nbA
  const unsigned inv_naB = (65536u + a.naB - 1) / a.naB;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int ka = s_ka + i * ka_step;
    const int dst = (int)(((unsigned)ka * inv_naB) >> 16), kal = ka - dst * a.naB;
    const int off = (kal * a.Rb + (int)rank * a.nbA + bl) * W + 2 * c;
    uint32_t ra_re, ra_im;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra_re) : "r"(smem_u32(s_re + off)), "r"(dst));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra_im) : "r"(smem_u32(s_im + off)), "r"(dst));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(ra_re), "f"(vr[i].x), "f"(vr[i].y) : "memory");
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(ra_im), "f"(vi[i].x), "f"(vi[i].y) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---- stand-in for the stages of sweep B
  for (int p = 0; p < a.passes; ++p) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int e = s + i * S;
      vr[i] = *reinterpret_cast<const float2*>(s_re + e * W + 2 * c);
      vi[i] = *reinterpret_cast<const float2*>(s_im + e * W + 2 * c);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int e = s + i * S;
      *reinterpret_cast<float2*>(s_re + e * W + 2 * c) = make_float2(vr[i].x + vi[i].y, vr[i].y - vi[i].x);
      *reinterpret_cast<float2*>(s_im + e * W + 2 * c) = make_float2(vi[i].x * 0.5f, vi[i].y * 0.5f);
    }
    __syncthreads();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    const int c0 = tile * W, c1 = 0, c2 = (int)rank * a.naB;
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(&st_re), "r"(smem_u32(s_re)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 :: "l"(&st_im), "r"(smem_u32(s_im)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn enc, float* plane, int P, int Ra, int Rb, int bw, int bb, int ba) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)Rb, (cuuint64_t)Ra};
  cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)P * 4 * Rb};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bb, (cuuint32_t)ba};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, plane, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box %d %d %d)\n", (int)r, bw, bb, ba); exit(1); }
  return m;
}

template <int NPT, int T>
static float run(EncodeFn enc, float* re, float* im, int P, int Ch, Args a, int reps, bool check) {
  const int R = a.Ra * a.Rb;
  CUtensorMap lr = make_map(enc, re, P, a.Ra, a.Rb, a.W, a.nbA, a.Ra), li = make_map(enc, im, P, a.Ra, a.Rb, a.W, a.nbA, a.Ra);
  CUtensorMap sr = make_map(enc, re, P, a.Ra, a.Rb, a.W, a.Rb, a.naB), si = make_map(enc, im, P, a.Ra, a.Rb, a.W, a.Rb, a.naB);
  const int ntiles = (Ch + 1 + a.W - 1) / a.W;
  const size_t smem = (size_t)2 * a.Ra * a.nbA * a.W * 4;
  CK(cudaFuncSetAttribute(k_skeleton<NPT, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (a.K > 8) CK(cudaFuncSetAttribute(k_skeleton<NPT, T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ntiles * a.K); cfg.blockDim = dim3(T); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = a.K; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int maxc = 0;
  cudaOccupancyMaxActiveClusters(&maxc, k_skeleton<NPT, T>, &cfg);
  if (check) {
    // correctness of the index plumbing with passes = 0: out[Rb*ka + kb] must equal in[Rb*a + b] with (a, b) = (ka, kb)
    // (the skeleton moves element (a, b) to position (ka = a, kb = b): an identity on the plane)
    std::vector<float> h((size_t)R * P);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1000003);
    CK(cudaMemcpy(re, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    Args a0 = a; a0.passes = 0;
    CK(cudaLaunchKernelEx(&cfg, k_skeleton<NPT, T>, lr, li, sr, si, a0));
    CK(cudaDeviceSynchronize());
    std::vector<float> g(h.size());
    CK(cudaMemcpy(g.data(), re, g.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int r = 0; r < R; ++r) for (int cc = 0; cc < ntiles * a.W && cc < P; ++cc) if (g[(size_t)r * P + cc] != h[(size_t)r * P + cc]) ++bad;
    printf("  identity check: %zu mismatches\n", bad);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) CK(cudaLaunchKernelEx(&cfg, k_skeleton<NPT, T>, lr, li, sr, si, a));
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) CK(cudaLaunchKernelEx(&cfg, k_skeleton<NPT, T>, lr, li, sr, si, a));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 2.0 * 2.0 * (double)R * (Ch + 1) * 4.0;     // both planes, read + write
  printf("R=%d (Ra=%d Rb=%d) Ch=%d K=%d W=%d T=%d NPT=%d passes=%d smem=%zu KB max_active_clusters=%d : %.1f us/launch, %.0f GB/s\n",
         R, a.Ra, a.Rb, Ch, a.K, a.W, T, NPT, a.passes, smem / 1024, maxc, 1000.0 * ms / reps, bytes / (ms / reps * 1e-3) / 1e9);
  return ms / reps;
}

int main(int argc, char** argv) {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  // largest case: R = 28672, C = 8192 -> P = 4128
  const size_t maxfl = (size_t)28672 * 4128 > (size_t)4096 * 7200 ? (size_t)28672 * 4128 : (size_t)4096 * 7200;
  float *re, *im;
  CK(cudaMalloc(&re, maxfl * 4)); CK(cudaMalloc(&im, maxfl * 4));
  CK(cudaMemset(re, 0, maxfl * 4)); CK(cudaMemset(im, 0, maxfl * 4));
  const int reps = 10;
  for (int passes = 0; passes <= 2; passes += 2) {
    // R = 14336 = 112 x 128, C = 4096 (gate / up): K = 8, W = 8 -> 112 KB per CTA, 14 complex-pairs per thread at T = 512
    { Args a{112, 128, 8, 8, 16, 14, passes}; run<14, 512>(enc, re, im, 2080, 2048, a, reps, passes == 0); }
    // same with W = 4 (16-byte rows): K = 4 -> 112 KB per CTA; elements per CTA = 112*32*2 pairs = 7168 / 256 thr = 28... use T = 512: 14
    { Args a{112, 128, 4, 4, 32, 28, passes}; run<14, 512>(enc, re, im, 2080, 2048, a, reps, passes == 0); }
    // R = 4096 = 64 x 64, C = 14336 (down): P = 7200
    { Args a{64, 64, 4, 8, 16, 16, passes}; run<16, 256>(enc, re, im, 7200, 7168, a, reps, passes == 0); }    // 64 KB per CTA
    { Args a{64, 64, 2, 8, 32, 32, passes}; run<16, 512>(enc, re, im, 7200, 7168, a, reps, passes == 0); }    // 128 KB per CTA
    { Args a{64, 64, 8, 16, 8, 8, passes}; run<8, 512>(enc, re, im, 7200, 7168, a, reps, passes == 0); }      // W = 16, 64 KB per CTA
    { Args a{64, 64, 4, 16, 16, 16, passes}; run<16, 512>(enc, re, im, 7200, 7168, a, reps, passes == 0); }   // W = 16, 128 KB per CTA
    // R = 1024 = 32 x 32, C = 4096: no cluster
    { Args a{32, 32, 1, 8, 32, 32, passes}; run<16, 256>(enc, re, im, 2080, 2048, a, reps, passes == 0); }    // 64 KB
    { Args a{32, 32, 1, 16, 32, 32, passes}; run<16, 512>(enc, re, im, 2080, 2048, a, reps, passes == 0); }   // 128 KB
    // R = 8192 = 64 x 128, C = 8192: P = 4128
    { Args a{64, 128, 8, 8, 16, 8, passes}; run<16, 256>(enc, re, im, 4128, 4096, a, reps, passes == 0); }    // 64 KB
    { Args a{64, 128, 8, 16, 16, 8, passes}; run<16, 512>(enc, re, im, 4128, 4096, a, reps, passes == 0); }   // 128 KB
    // R = 28672 = 224 x 128, C = 8192: K = 16 (non-portable), W = 8 -> 112 KB per CTA
    { Args a{224, 128, 16, 8, 8, 14, passes}; run<14, 512>(enc, re, im, 4128, 4096, a, reps, passes == 0); }
    { Args a{224, 128, 8, 4, 16, 28, passes}; run<14, 512>(enc, re, im, 4128, 4096, a, reps, passes == 0); }   // W = 4, K = 8
  }
  // reference point: plain device-to-device copy of one plane pair
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t n = (size_t)14336 * 2080 * 4;
    cudaMemcpy(im, re, n, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(im, re, n, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("memcpy D2D %zu MB: %.0f GB/s (read + write)\n", n >> 20, 2.0 * n * 10 / (ms * 1e-3) / 1e9);
  }
  return 0;
}
