// kernels_elem.cu -- the element-wise strategies of the reference as single streaming kernels.
//
//   mode 0  AdditionMerge._merge_layer      (shard/merge/addition.py:44-83):  out = 0; out += (ft_k - base) for every model
//   mode 1  TaskAdditionMerge._merge_layer  (shard/merge/taskaddition.py:44-83): deltas whose sign differs from the
//           majority sign (sign of the sum of signs) are zeroed, the rest summed
//
// The reference does every step in the tensors' own dtype; for bf16 that is "compute in fp32, round to bf16 after each
// torch op", and torch.sum over bf16 accumulates in fp32 in model order and rounds once.  The kernels reproduce exactly
// that rounding sequence, so the results are bit-identical to the reference's (tests/golden/elem_*.npz), NaN payloads
// aside.  HBM-bound: (M + 1) reads and one write of 2 bytes per element; 8 elements (16 bytes per tensor) per thread.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "sm_internal.h"

namespace {

constexpr int kMaxModels = 8;
struct ElemArgs { const uint4* base; const uint4* ft[kMaxModels]; uint4* out; int M; size_t n8; size_t n; };

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// one bf16 rounding (RNE), widened again: what every torch bf16 op leaves behind
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
// torch.sign for floats: NaN stays NaN
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : (x == x ? 0.f : x)); }

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// M is a compile-time constant: all M + 1 16-byte loads of a thread are issued before the first is consumed, and the
// sign-agreement mode keeps the raw words in registers instead of reading the finetunes twice
template <int MODE, int M>
__device__ __forceinline__ void merge8(const uint4 (&raw)[M], const float (&b)[8], float (&o)[8]) {
  if (MODE == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = bf16r(o[e] + bf16r(f[e] - b[e]));      // addition.py:72-73
    }
  } else {
    float ssum[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ssum[e] = 0.f; o[e] = 0.f; }
#pragma unroll
    for (int k = 0; k < M; ++k) {                      // taskaddition.py:68-73: bf16 deltas, sum of their signs
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) ssum[e] += sgnf(bf16r(f[e] - b[e]));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) ssum[e] = sgnf(ssum[e]);
#pragma unroll
    for (int k = 0; k < M; ++k) {                      // :75-78: mask by the majority sign, fp32 sum in model order
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = bf16r(f[e] - b[e]);
        const float m = (sgnf(d) == ssum[e]) ? 1.f : 0.f;
        o[e] += bf16r(d * m);
      }
    }
  }
}

template <int MODE, int M>
__global__ void __launch_bounds__(256) k_elem_merge(const __grid_constant__ ElemArgs a) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n8; i += stride) {
    uint4 raw[M];
    const uint4 rb = __ldg(a.base + i);
#pragma unroll
    for (int k = 0; k < M; ++k) raw[k] = __ldg(a.ft[k] + i);
    float b[8], o[8];
    unpack8(rb, b);
    merge8<MODE, M>(raw, b, o);
    a.out[i] = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
  }
  // tail (n not a multiple of 8): one element per thread of the first CTA
  if (blockIdx.x == 0) {
    const uint16_t* base16 = reinterpret_cast<const uint16_t*>(a.base);
    uint16_t* out16 = reinterpret_cast<uint16_t*>(a.out);
    for (size_t j = a.n8 * 8 + threadIdx.x; j < a.n; j += blockDim.x) {
      const float b = __uint_as_float((uint32_t)base16[j] << 16);
      float o = 0.f, ssum = 0.f;
      if (MODE == 0) {
        for (int k = 0; k < M; ++k) {
          const float f = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(a.ft[k])[j] << 16);
          o = bf16r(o + bf16r(f - b));
        }
      } else {
        for (int k = 0; k < M; ++k) {
          const float f = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(a.ft[k])[j] << 16);
          ssum += sgnf(bf16r(f - b));
        }
        for (int k = 0; k < M; ++k) {
          const float f = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(a.ft[k])[j] << 16);
          const float d = bf16r(f - b);
          o += bf16r(d * ((sgnf(d) == sgnf(ssum)) ? 1.f : 0.f));
        }
      }
      __nv_bfloat16 h = __float2bfloat16_rn(o);
      out16[j] = *reinterpret_cast<uint16_t*>(&h);
    }
  }
}

// ---- any storage dtype, up to 64 models (the reference takes whatever the checkpoints hold; the compile-time kernels above
// cover the common case, bf16 and <= 8 finetunes).  Run-time loop over the models, four consecutive elements per thread;
// the sign-agreement mode reads the finetunes twice (sum of signs, then the masked sum).  torch.sum on the CPU (where
// taskaddition.py:69 puts the stack) keeps fp32 accumulators in model order and, from 16 rows on, folds every 16 rows
// into a second accumulator that is added at the end (ATen SumKernel.cpp multi_row_sum; oracle_np._cascade_sum_rows):
// reproduced here so that fp32 checkpoints come out bit-identical as well.
constexpr int kMaxAny = 64;
struct ElemAnyArgs { const void* base; const void* ft[kMaxAny]; void* out; int M; size_t n; };

template <int DT> struct ElemCodec;
template <> struct ElemCodec<0> {                    // fp32
  __device__ static float load(const void* p, size_t i) { return __ldg(reinterpret_cast<const float*>(p) + i); }
  __device__ static void load4(const void* p, size_t i4, float (&f)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i4); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static float rnd(float x) { return x; }
  __device__ static void store(void* p, size_t i, float v) { reinterpret_cast<float*>(p)[i] = v; }
  __device__ static void store4(void* p, size_t i4, const float (&o)[4]) { reinterpret_cast<float4*>(p)[i4] = make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct ElemCodec<1> {                    // bf16
  __device__ static float load(const void* p, size_t i) { return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const uint16_t*>(p) + i) << 16); }
  __device__ static void load4(const void* p, size_t i4, float (&f)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  }
  __device__ static float rnd(float x) { return bf16r(x); }
  __device__ static void store(void* p, size_t i, float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); reinterpret_cast<uint16_t*>(p)[i] = *reinterpret_cast<uint16_t*>(&h); }
  __device__ static void store4(void* p, size_t i4, const float (&o)[4]) { reinterpret_cast<uint2*>(p)[i4] = make_uint2(pack2(o[0], o[1]), pack2(o[2], o[3])); }
};
template <> struct ElemCodec<2> {                    // fp16
  __device__ static float h2f(uint16_t u) { return __half2float(*reinterpret_cast<const __half*>(&u)); }
  __device__ static uint16_t f2h(float x) { __half h = __float2half_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
  __device__ static float load(const void* p, size_t i) { return h2f(__ldg(reinterpret_cast<const uint16_t*>(p) + i)); }
  __device__ static void load4(const void* p, size_t i4, float (&f)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    f[0] = h2f((uint16_t)(v.x & 0xffffu)); f[1] = h2f((uint16_t)(v.x >> 16)); f[2] = h2f((uint16_t)(v.y & 0xffffu)); f[3] = h2f((uint16_t)(v.y >> 16));
  }
  __device__ static float rnd(float x) { return h2f(f2h(x)); }
  __device__ static void store(void* p, size_t i, float v) { reinterpret_cast<uint16_t*>(p)[i] = f2h(v); }
  __device__ static void store4(void* p, size_t i4, const float (&o)[4]) {
    reinterpret_cast<uint2*>(p)[i4] = make_uint2((uint32_t)f2h(o[0]) | ((uint32_t)f2h(o[1]) << 16), (uint32_t)f2h(o[2]) | ((uint32_t)f2h(o[3]) << 16));
  }
};

// W elements (4: vector body, 1: tail) starting at element e0; `get(p, f)` loads them from tensor p
template <int MODE, int DT, int W, class Get>
__device__ __forceinline__ void merge_any(const ElemAnyArgs& a, const Get& get, float (&o)[W]) {
  using Cd = ElemCodec<DT>;
  float b[W], f[W];
  get(a.base, b);
  if (MODE == 0) {
#pragma unroll
    for (int e = 0; e < W; ++e) o[e] = 0.f;
    for (int k = 0; k < a.M; ++k) {
      get(a.ft[k], f);
#pragma unroll
      for (int e = 0; e < W; ++e) o[e] = Cd::rnd(o[e] + Cd::rnd(f[e] - b[e]));          // addition.py:72-73
    }
  } else {
    float ssum[W], a0[W], a1[W];
#pragma unroll
    for (int e = 0; e < W; ++e) { ssum[e] = 0.f; a0[e] = 0.f; a1[e] = 0.f; }
    for (int k = 0; k < a.M; ++k) {                      // taskaddition.py:68-73
      get(a.ft[k], f);
#pragma unroll
      for (int e = 0; e < W; ++e) ssum[e] += sgnf(Cd::rnd(f[e] - b[e]));
    }
#pragma unroll
    for (int e = 0; e < W; ++e) ssum[e] = sgnf(ssum[e]);
    for (int k = 0; k < a.M; ++k) {                      // :75-78
      get(a.ft[k], f);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const float d = Cd::rnd(f[e] - b[e]);
        a0[e] += Cd::rnd(d * ((sgnf(d) == ssum[e]) ? 1.f : 0.f));
      }
      if (((k + 1) & 15) == 0) {                         // a full group of 16 rows: fold into the next level
#pragma unroll
        for (int e = 0; e < W; ++e) { a1[e] += a0[e]; a0[e] = 0.f; }
      }
    }
#pragma unroll
    for (int e = 0; e < W; ++e) o[e] = a0[e] + a1[e];
  }
}

template <int MODE, int DT>
__global__ void __launch_bounds__(256) k_elem_merge_any(const __grid_constant__ ElemAnyArgs a) {
  using Cd = ElemCodec<DT>;
  const size_t n4 = a.n / 4, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float o[4];
    merge_any<MODE, DT, 4>(a, [&](const void* p, float (&f)[4]) { Cd::load4(p, i, f); }, o);
    if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = Cd::rnd(o[e]);
    }
    Cd::store4(a.out, i, o);
  }
  if (blockIdx.x == 0) {
    for (size_t j = n4 * 4 + threadIdx.x; j < a.n; j += blockDim.x) {
      float o[1];
      merge_any<MODE, DT, 1>(a, [&](const void* p, float (&f)[1]) { f[0] = Cd::load(p, j); }, o);
      Cd::store(a.out, j, o[0]);
    }
  }
}

// ---- correlate_pairs (shard/tensor/functions.py:304-314): mean over the columns of the cosine similarity of two [R][C]
// tensors along dim 0.  One CTA per 32 adjacent columns walks all rows (coalesced 128-byte row segments for fp32, 64 for
// bf16), three fp32 sums per column, one atomic per CTA.  torch: (x / max(||x||, eps)) . (y / max(||y||, eps)), eps = 1e-8,
// NaN -> 0 (nan_to_num), then .mean().  HBM-bound: both tensors are read once.
template <class T> __device__ __forceinline__ float ld_as_f32(const T* p, size_t i);
template <> __device__ __forceinline__ float ld_as_f32<float>(const float* p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ld_as_f32<uint16_t>(const uint16_t* p, size_t i) { return __uint_as_float((uint32_t)__ldg(p + i) << 16); }

template <class T>
__global__ void __launch_bounds__(256) k_cosine_cols(const T* __restrict__ a, const T* __restrict__ b, int R, int C, double* out_sum) {
  __shared__ float sab[8][33], saa[8][33], sbb[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  if (c < C) {
    for (int r = w; r < R; r += 8) {
      const float x = ld_as_f32<T>(a, (size_t)r * C + c), y = ld_as_f32<T>(b, (size_t)r * C + c);
      ab = fmaf(x, y, ab); aa = fmaf(x, x, aa); bb = fmaf(y, y, bb);
    }
  }
  sab[w][lane] = ab; saa[w][lane] = aa; sbb[w][lane] = bb;
  __syncthreads();
  if (w == 0) {
    float cs = 0.f;
    if (c < C) {
      for (int k = 1; k < 8; ++k) { ab += sab[k][lane]; aa += saa[k][lane]; bb += sbb[k][lane]; }
      cs = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
      if (cs != cs) cs = 0.f;
    }
    double v = (double)cs;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(out_sum, v);
  }
}

}  // namespace

extern "C" int sm_cosine_cols(int dtype, int R, int C, const void* a, const void* b, double* out_sum, void* stream) {
  if (R < 1 || C < 1) { sm_set_error("cosine_cols: empty tensor"); return -2; }
  cudaStream_t st = (cudaStream_t)stream;
  SM_CUDA_CHECK(cudaMemsetAsync(out_sum, 0, sizeof(double), st));
  const unsigned int grid = (unsigned int)((C + 31) / 32);
  if (dtype == 0) k_cosine_cols<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, R, C, out_sum);
  else if (dtype == 1) k_cosine_cols<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)a, (const uint16_t*)b, R, C, out_sum);
  else { sm_set_error("cosine_cols: dtype 0 (fp32) or 1 (bf16)"); return -2; }
  SM_LAUNCH_CHECK();
  return 0;
}

static int elem_grid(size_t vec_items) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  size_t need = (vec_items + 255) / 256;
  if (need < 1) need = 1;
  const size_t cap = (size_t)sms * 16;                  // two resident waves of 8 CTAs per SM, grid-stride over the rest
  return (int)(need < cap ? need : cap);
}

extern "C" int sm_elem_merge(int mode, int dtype, size_t n, const void* base, const void* const* fts, int n_models, void* out,
                             void* stream) {
  if (mode < 0 || mode > 1) { sm_set_error("elem_merge: unknown mode %d", mode); return -2; }
  if (dtype < 0 || dtype > 2) { sm_set_error("elem_merge: dtype 0 (fp32), 1 (bf16) or 2 (fp16), got %d", dtype); return -2; }
  if (n_models < 1 || n_models > kMaxAny) { sm_set_error("elem_merge: 1..%d models, got %d", kMaxAny, n_models); return -2; }
  bool aligned = ((uintptr_t)base % 16 == 0) && ((uintptr_t)out % 16 == 0);
  for (int k = 0; k < n_models; ++k) aligned = aligned && ((uintptr_t)fts[k] % 16 == 0);
  if (!aligned) { sm_set_error("elem_merge: tensors must be 16-byte aligned"); return -2; }
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 1 && n_models <= kMaxModels) {            // the common case: compile-time model count, 8 elements per thread
    ElemArgs a{};
    a.base = reinterpret_cast<const uint4*>(base); a.out = reinterpret_cast<uint4*>(out); a.M = n_models;
    a.n = n; a.n8 = n / 8;
    for (int k = 0; k < n_models; ++k) a.ft[k] = reinterpret_cast<const uint4*>(fts[k]);
    const unsigned int grid = (unsigned int)elem_grid(a.n8);
#define SM_ELEM_CASE(MM)                                                 \
  case MM:                                                               \
    if (mode == 0) k_elem_merge<0, MM><<<grid, 256, 0, st>>>(a);         \
    else k_elem_merge<1, MM><<<grid, 256, 0, st>>>(a);                   \
    break;
    switch (n_models) {
      SM_ELEM_CASE(1) SM_ELEM_CASE(2) SM_ELEM_CASE(3) SM_ELEM_CASE(4)
      SM_ELEM_CASE(5) SM_ELEM_CASE(6) SM_ELEM_CASE(7) SM_ELEM_CASE(8)
    }
#undef SM_ELEM_CASE
    SM_LAUNCH_CHECK();
    return 0;
  }
  ElemAnyArgs a{};
  a.base = base; a.out = out; a.M = n_models; a.n = n;
  for (int k = 0; k < n_models; ++k) a.ft[k] = fts[k];
  const unsigned int grid = (unsigned int)elem_grid(n / 4);
#define SM_ELEM_ANY(DT)                                                  \
  case DT:                                                               \
    if (mode == 0) k_elem_merge_any<0, DT><<<grid, 256, 0, st>>>(a);     \
    else k_elem_merge_any<1, DT><<<grid, 256, 0, st>>>(a);               \
    break;
  switch (dtype) { SM_ELEM_ANY(0) SM_ELEM_ANY(1) SM_ELEM_ANY(2) }
#undef SM_ELEM_ANY
  SM_LAUNCH_CHECK();
  return 0;
}

extern "C" int sm_elem_merge_bf16(int mode, size_t n, const void* base, const void* const* fts, int n_models, void* out,
                                  void* stream) {
  return sm_elem_merge(mode, 1, n, base, fts, n_models, out, stream);
}
