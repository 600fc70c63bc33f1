// fft_core.cuh -- register butterflies and the Stockham stage used by every FFT kernel.
//
// Everything in this header is __host__ __device__ and free of CUDA intrinsics, so the
// exact same index math / butterflies can be driven from a CPU loop in tests/hostemu
// (one emulated CTA at a time) as well as from the sm_100a kernels in kernels_fft.cu.
//
// Conventions
//   * forward transform only: X[k] = sum_n x[n] * exp(-2*pi*i*n*k/N).  The inverse is
//     obtained by swapping re/im at the input and at the output of the forward engine
//     (ifft(z) = swap(fft(swap(z))), unnormalised), so no second set of twiddles exists.
//   * Stockham autosort, decimation in frequency: natural order in, natural order out.
//   * twiddle tables hold W_M^j = (cos(2*pi*j/M), -sin(2*pi*j/M)) as float2, computed
//     in double precision (kernels_fft.cu: k_init_twiddles).
//
// Replaces torch.fft.fft / fftn / ifft / ifftn as called by the reference at
// shard/tensor/functions.py:55-58 and :70-73 (library calls into MKL / cuFFT there).
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>
#if defined(__CUDACC__)
#include <cuda_bf16.h>
#endif

#if defined(__CUDACC__)
#define SM_HD __host__ __device__ __forceinline__
#define SM_CX __host__ __device__ constexpr
#else
#define SM_HD inline
#define SM_CX constexpr
#endif

namespace smfft {

struct cf { float x, y; };   // plain complex<float>, layout-compatible with float2

// read-only global loads: ld.global.nc on the device lets the compiler hoist them above stores
SM_HD cf ldg_cf(const cf* p) {
#if defined(__CUDA_ARCH__)
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  cf r; r.x = v.x; r.y = v.y; return r;
#else
  return *p;
#endif
}
SM_HD float ldg_f32(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
// streaming load that never allocates in L1 (ld.global.cg): the column sweeps read data another CTA of the same
// launch may have just written (fused two-sweep kernel), so a stale L1 line must not be hit
SM_HD float ldcg_f32(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
SM_HD uint32_t ldg_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ---------------------------------------------------------------------------------
// compile-time trigonometry (only used to bake small-radix constants into immediates)
// ---------------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

SM_CX double c_reduce(double x) {  // to [-pi, pi]
  while (x > kPi) x -= 2.0 * kPi;
  while (x < -kPi) x += 2.0 * kPi;
  return x;
}
SM_CX double c_sin(double x0) {
  double x = c_reduce(x0), x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 20; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
  return sum;
}
SM_CX double c_cos(double x0) {
  double x = c_reduce(x0), x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 20; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
  return sum;
}
// cos / sin of 2*pi*k/n with the exact values at multiples of an eighth turn.
SM_CX double c_cos2pi(int k, int n) {
  k %= n; if (k < 0) k += n;
  if ((4 * k) % n == 0) { int q = (4 * k) / n; return q == 0 ? 1.0 : (q == 2 ? -1.0 : 0.0); }
  if ((8 * k) % n == 0) { int q = (8 * k) / n; return (q == 1 || q == 7) ? 0.70710678118654752440 : -0.70710678118654752440; }
  return c_cos(2.0 * kPi * double(k) / double(n));
}
SM_CX double c_sin2pi(int k, int n) {
  k %= n; if (k < 0) k += n;
  if ((4 * k) % n == 0) { int q = (4 * k) / n; return q == 1 ? 1.0 : (q == 3 ? -1.0 : 0.0); }
  if ((8 * k) % n == 0) { int q = (8 * k) / n; return (q == 1 || q == 3) ? 0.70710678118654752440 : -0.70710678118654752440; }
  return c_sin(2.0 * kPi * double(k) / double(n));
}

template <int I, int N, class F>
SM_HD void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---------------------------------------------------------------------------------
// value types of the butterflies: float, or pf = two independent fp32 lanes that go through
// identical arithmetic (two adjacent columns of a column sweep).  On sm_100a a pf lives in an
// aligned register pair and every operation is ONE packed instruction (FADD2 / FMUL2 / FFMA2,
// PTX add/sub/mul/fma.rn.f32x2): the FFT kernels are instruction-issue bound (profiles/r01), and
// packing halves the FP, load/store and index instructions per element.  Each lane rounds exactly
// like the scalar operation, so a pf butterfly is bit-identical to two float butterflies.
// ---------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
struct pf { unsigned long long v; };
SM_HD pf pf_make(float a, float b) { pf r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
SM_HD float pf_lo(pf x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return a; }
SM_HD float pf_hi(pf x) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); return b; }
SM_HD pf operator+(pf a, pf b) { pf r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SM_HD pf operator-(pf a, pf b) { pf r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SM_HD pf operator*(pf a, pf b) { pf r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SM_HD pf pf_fma(pf a, pf b, pf c) { pf r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct pf { float a, b; };
SM_HD pf pf_make(float a, float b) { pf r; r.a = a; r.b = b; return r; }
SM_HD float pf_lo(pf x) { return x.a; }
SM_HD float pf_hi(pf x) { return x.b; }
SM_HD pf operator+(pf x, pf y) { return pf_make(x.a + y.a, x.b + y.b); }
SM_HD pf operator-(pf x, pf y) { return pf_make(x.a - y.a, x.b - y.b); }
SM_HD pf operator*(pf x, pf y) { return pf_make(x.a * y.a, x.b * y.b); }
SM_HD pf pf_fma(pf x, pf y, pf z) { return pf_make(fmaf(x.a, y.a, z.a), fmaf(x.b, y.b, z.b)); }
#endif
SM_HD pf pf_bcast(float c) { return pf_make(c, c); }
SM_HD pf& operator+=(pf& a, pf b) { a = a + b; return a; }

// scalar-coefficient helpers, overloaded for both value types
SM_HD float mulc(float x, float c) { return x * c; }
SM_HD float fmac(float c, float x, float acc) { return acc + c * x; }        // contracted to one FFMA by nvcc
SM_HD float zero_of(float) { return 0.f; }
SM_HD pf mulc(pf x, float c) { return x * pf_bcast(c); }
SM_HD pf fmac(float c, pf x, pf acc) { return pf_fma(pf_bcast(c), x, acc); }
SM_HD pf zero_of(pf) { return pf_make(0.f, 0.f); }

// y = x * W_N^K, W_N = exp(-2*pi*i/N), K and N compile-time.
template <int K, int N, class T>
SM_HD void twmul(T xr, T xi, T& yr, T& yi) {
  constexpr int k = ((K % N) + N) % N;
  if constexpr (k == 0) { yr = xr; yi = xi; }
  else if constexpr (4 * k == N) { yr = xi; yi = zero_of(xr) - xr; }          // -i
  else if constexpr (2 * k == N) { yr = zero_of(xr) - xr; yi = zero_of(xi) - xi; }         // -1
  else if constexpr (4 * k == 3 * N) { yr = zero_of(xi) - xi; yi = xr; }      // +i
  else {
    constexpr float c = float(c_cos2pi(k, N));
    constexpr float s = float(c_sin2pi(k, N));
    // (xr + i xi)(c - i s)
    yr = fmac(s, xi, mulc(xr, c));
    yi = fmac(-s, xr, mulc(xi, c));
  }
}

SM_HD void cmul(float& xr, float& xi, float wr, float wi) {
  float tr = xr * wr - xi * wi;
  float ti = xr * wi + xi * wr;
  xr = tr; xi = ti;
}
SM_HD void cmul(pf& xr, pf& xi, float wr, float wi) {       // one twiddle for both lanes
  const pf pwr = pf_bcast(wr), pwi = pf_bcast(wi);
  const pf tr = pf_fma(xr, pwr, zero_of(xr) - xi * pwi);
  const pf ti = pf_fma(xr, pwi, xi * pwr);
  xr = tr; xi = ti;
}

// ---------------------------------------------------------------------------------
// in-register forward DFTs of small length
// ---------------------------------------------------------------------------------
template <int N> struct Dft;

template <> struct Dft<1> {
  template <class T> static SM_HD void run(T (&)[1], T (&)[1]) {}
};

template <> struct Dft<2> {
  template <class T> static SM_HD void run(T (&re)[2], T (&im)[2]) {
    T ar = re[0], ai = im[0], br = re[1], bi = im[1];
    re[0] = ar + br; im[0] = ai + bi;
    re[1] = ar - br; im[1] = ai - bi;
  }
};

template <> struct Dft<4> {
  template <class T> static SM_HD void run(T (&re)[4], T (&im)[4]) {
    T t0r = re[0] + re[2], t0i = im[0] + im[2];
    T t1r = re[0] - re[2], t1i = im[0] - im[2];
    T t2r = re[1] + re[3], t2i = im[1] + im[3];
    T t3r = re[1] - re[3], t3i = im[1] - im[3];
    re[0] = t0r + t2r; im[0] = t0i + t2i;
    re[2] = t0r - t2r; im[2] = t0i - t2i;
    // X1 = t1 - i*t3 ; X3 = t1 + i*t3
    re[1] = t1r + t3i; im[1] = t1i - t3r;
    re[3] = t1r - t3i; im[3] = t1i + t3r;
  }
};

// odd prime length, symmetric (real cos/sin) formulation: (P-1)^2 FMAs.
template <int P> struct DftPrime {
  template <class T> static SM_HD void run(T (&re)[P], T (&im)[P]) {
    constexpr int H = (P - 1) / 2;
    T sr[H], si[H], dr[H], di[H];
    static_for<0, H>([&](auto j_) {
      constexpr int j = decltype(j_)::value;
      sr[j] = re[j + 1] + re[P - 1 - j]; si[j] = im[j + 1] + im[P - 1 - j];
      dr[j] = re[j + 1] - re[P - 1 - j]; di[j] = im[j + 1] - im[P - 1 - j];
    });
    const T r0 = re[0], i0 = im[0];
    T sumr = r0, sumi = i0;
    static_for<0, H>([&](auto j_) { constexpr int j = decltype(j_)::value; sumr = sumr + sr[j]; sumi = sumi + si[j]; });
    static_for<1, H + 1>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      T ar = r0, ai = i0, br = zero_of(r0), bi = zero_of(r0);
      static_for<0, H>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr float c = float(c_cos2pi((j + 1) * k, P));
        constexpr float s = float(c_sin2pi((j + 1) * k, P));
        ar = fmac(c, sr[j], ar); ai = fmac(c, si[j], ai);
        br = fmac(s, dr[j], br); bi = fmac(s, di[j], bi);
      });
      // X[k] = a - i*b ; X[P-k] = a + i*b   with a=(ar,ai), b=(br,bi)
      re[k] = ar + bi; im[k] = ai - br;
      re[P - k] = ar - bi; im[P - k] = ai + br;
    });
    re[0] = sumr; im[0] = sumi;
  }
};
template <> struct Dft<3> : DftPrime<3> {};
template <> struct Dft<5> : DftPrime<5> {};
template <> struct Dft<7> : DftPrime<7> {};
template <> struct Dft<11> : DftPrime<11> {};
template <> struct Dft<13> : DftPrime<13> {};

// composite length A*B, one Cooley-Tukey level in registers.
// input index n = B*a + b, output index k = ka + A*kb.
template <int A, int B> struct DftCT {
  template <class T> static SM_HD void run(T (&re)[A * B], T (&im)[A * B]) {
    T tr[A * B], ti[A * B];
    static_for<0, B>([&](auto b_) {
      constexpr int b = decltype(b_)::value;
      T xr[A], xi[A];
      static_for<0, A>([&](auto a_) { constexpr int a = decltype(a_)::value; xr[a] = re[B * a + b]; xi[a] = im[B * a + b]; });
      Dft<A>::run(xr, xi);
      static_for<0, A>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        twmul<b * k, A * B>(xr[k], xi[k], tr[k * B + b], ti[k * B + b]);
      });
    });
    static_for<0, A>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      T zr[B], zi[B];
      static_for<0, B>([&](auto b_) { constexpr int b = decltype(b_)::value; zr[b] = tr[k * B + b]; zi[b] = ti[k * B + b]; });
      Dft<B>::run(zr, zi);
      static_for<0, B>([&](auto q_) { constexpr int q = decltype(q_)::value; re[k + A * q] = zr[q]; im[k + A * q] = zi[q]; });
    });
  }
};
template <> struct Dft<8> : DftCT<4, 2> {};
template <> struct Dft<16> : DftCT<4, 4> {};

// ---------------------------------------------------------------------------------
// one Stockham (DIF, autosort) butterfly of radix r
//
//   stage state: total length N, stride s = product of the radices of earlier stages.
//   butterfly b in [0, N/r):  q = b % s, p = b / s
//     reads   x[b + j*N/r]                 j = 0..r-1
//     writes  y[q + s*(r*p + k)] * W_N^(s*p*k)   k = 0..r-1
//   The twiddle table is W_M with M = N*tw_mul (so W_N^e = tab[e*tw_mul]).
//   In the last stage (s*r == N) p is always 0 and the twiddles vanish (kLast).
// ---------------------------------------------------------------------------------
template <int r, bool kLast, class T = float, class Src, class Dst>
SM_HD void stockham_bfly(int b, int N, int s, int tw_mul, const cf* tw, const Src& src, const Dst& dst) {
  T re[r], im[r];
  const int Nr = N / r;
  static_for<0, r>([&](auto j_) {
    constexpr int j = decltype(j_)::value;
    src.load(b + j * Nr, re[j], im[j]);
  });
  Dft<r>::run(re, im);
  if constexpr (kLast) {
    static_for<0, r>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      dst.store(b + k * s, re[k], im[k]);   // p == 0, q == b
    });
  } else {
    const int p = b / s;
    const int q = b - p * s;
    const int obase = q + s * r * p;
    const int tstep = s * p * tw_mul;
    dst.store(obase, re[0], im[0]);
    static_for<1, r>([&](auto k_) {
      constexpr int k = decltype(k_)::value;
      const cf w = ldg_cf(tw + tstep * k);
      T xr = re[k], xi = im[k];
      cmul(xr, xi, w.x, w.y);
      dst.store(obase + k * s, xr, xi);
    });
  }
}

// ---------------------------------------------------------------------------------
// first stage of a row pass (s == 1, p == b): every butterfly needs its own r-1 twiddles
// W_N^(b*k).  Looked up one by one in the W table they are 32 scattered sectors per warp
// request (measured: 90 % of the L1 traffic of k_row_fwd, profiles/r01).  Instead each
// butterfly reads one 32-byte "quad" {W^b, W^2b, W^4b, W^8b} (coalesced: 1 KB per warp)
// and forms the other powers as products of at most three table values.
// ---------------------------------------------------------------------------------
SM_HD void ldg_quad(const cf* quad, int b, cf (&q)[4]) {
#if defined(__CUDA_ARCH__)
  const float4 lo = __ldg(reinterpret_cast<const float4*>(quad + 4 * (size_t)b));
  const float4 hi = __ldg(reinterpret_cast<const float4*>(quad + 4 * (size_t)b) + 1);
  q[0].x = lo.x; q[0].y = lo.y; q[1].x = lo.z; q[1].y = lo.w;
  q[2].x = hi.x; q[2].y = hi.y; q[3].x = hi.z; q[3].y = hi.w;
#else
  for (int i = 0; i < 4; ++i) q[i] = quad[4 * (size_t)b + i];
#endif
}

SM_CX int c_top_bit(int k) { int t = 1; while (2 * t <= k) t *= 2; return t; }

// qmul: the table was built for a transform qmul times longer (W_N = W_{qmul N}^qmul): read entry qmul * b
template <int r>
SM_HD void quad_twiddles(const cf* quad, int b, float (&wr)[r], float (&wi)[r], int qmul = 1) {
  cf q[4];
  ldg_quad(quad, b * qmul, q);
  wr[0] = 1.f; wi[0] = 0.f;
  if constexpr (r > 1) { wr[1] = q[0].x; wi[1] = q[0].y; }
  if constexpr (r > 2) { wr[2] = q[1].x; wi[2] = q[1].y; }
  if constexpr (r > 4) { wr[4] = q[2].x; wi[4] = q[2].y; }
  if constexpr (r > 8) { wr[8] = q[3].x; wi[8] = q[3].y; }
  static_for<3, r>([&](auto k_) {
    constexpr int k = decltype(k_)::value;
    constexpr int t = c_top_bit(k);
    if constexpr (t != k) {
      float xr = wr[t], xi = wi[t];
      cmul(xr, xi, wr[k - t], wi[k - t]);
      wr[k] = xr; wi[k] = xi;
    }
  });
}

// non-last first stage: s == 1, so q == 0, p == b, outputs go to r*b + k
template <int r, class V = float, class Src, class Dst>
SM_HD void stockham_bfly_first(int b, int N, const cf* quad, const Src& src, const Dst& dst, int qmul = 1) {
  V re[r], im[r];
  const int Nr = N / r;
  static_for<0, r>([&](auto j_) {
    constexpr int j = decltype(j_)::value;
    src.load(b + j * Nr, re[j], im[j]);
  });
  float wr[r], wi[r];
  quad_twiddles<r>(quad, b, wr, wi, qmul);
  Dft<r>::run(re, im);
  const int obase = r * b;
  dst.store(obase, re[0], im[0]);
  static_for<1, r>([&](auto k_) {
    constexpr int k = decltype(k_)::value;
    V xr = re[k], xi = im[k];
    cmul(xr, xi, wr[k], wi[k]);
    dst.store(obase + k, xr, xi);
  });
}

template <class Src, class Dst>
SM_HD void stockham_bfly_first_rt(int r, int b, int N, const cf* quad, const Src& src, const Dst& dst) {
  switch (r) {
    case 1:  stockham_bfly_first<1>(b, N, quad, src, dst); break;     // identity stage in front of a generic radix
    case 2:  stockham_bfly_first<2>(b, N, quad, src, dst); break;
    case 3:  stockham_bfly_first<3>(b, N, quad, src, dst); break;
    case 4:  stockham_bfly_first<4>(b, N, quad, src, dst); break;
    case 5:  stockham_bfly_first<5>(b, N, quad, src, dst); break;
    case 7:  stockham_bfly_first<7>(b, N, quad, src, dst); break;
    case 8:  stockham_bfly_first<8>(b, N, quad, src, dst); break;
    case 11: stockham_bfly_first<11>(b, N, quad, src, dst); break;
    case 13: stockham_bfly_first<13>(b, N, quad, src, dst); break;
    case 16: stockham_bfly_first<16>(b, N, quad, src, dst); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------
// generic radix: any r (the plan uses it for prime factors above 13, e.g. 37 in 18944 = 2^9 * 37 or 167 in
// 128256 = 2^8 * 3 * 167).  One call computes ONE output k of butterfly b as an r-term sum,
//   y[q + s*(r*p + k)] = W_N^(s*p*k) * sum_j x[b + j*N/r] * W_r^(j*k),      W_r^e = tab[e * (N/r) * tw_mul],
// re-reading its inputs from the source (global memory through L1, or shared memory): O(r) per element instead of
// O(log r), which is what a length that does not factor costs without a Bluestein / Rader stage.  Work items are
// (b, k) pairs, so a stage has as many independent items as the transform has points.  Sources must be pure (the
// plan keeps generic radices out of the first row stage, whose delta source accumulates the sum of squares).
// ---------------------------------------------------------------------------------
SM_HD bool sm_radix_is_generic(int r) {
  return !(r == 1 || r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 11 || r == 13 || r == 16);
}

template <bool kLast, class Src, class Dst>
SM_HD void stockham_generic_output(int r, int item, int N, int s, int tw_mul, const cf* tw, const Src& src, const Dst& dst) {
  const int Nr = N / r;
  const int b = item / r, k = item - b * r;
  const int p = kLast ? 0 : b / s;
  const int q = b - p * s;
  const size_t wstep = (size_t)Nr * tw_mul;
  double accr = 0.0, acci = 0.0;                   // r terms: accumulate in fp64, one rounding at the end
  int e = 0;                                       // (j * k) mod r
  for (int j = 0; j < r; ++j) {
    float xr, xi;
    src.load(b + j * Nr, xr, xi);
    const cf w = ldg_cf(tw + (size_t)e * wstep);
    accr += (double)xr * w.x - (double)xi * w.y;
    acci += (double)xr * w.y + (double)xi * w.x;
    e += k; if (e >= r) e -= r;
  }
  float ar = (float)accr, ai = (float)acci;
  if (!kLast && k > 0) {
    const cf w = ldg_cf(tw + (size_t)s * p * tw_mul * k);
    cmul(ar, ai, w.x, w.y);
  }
  dst.store(q + s * r * p + k * s, ar, ai);
}

// runtime-radix dispatch (the plan is data, the butterflies are code)
template <bool kLast, class Src, class Dst>
SM_HD void stockham_bfly_rt(int r, int b, int N, int s, int tw_mul, const cf* tw, const Src& src, const Dst& dst) {
  switch (r) {
    case 1:  stockham_bfly<1, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 2:  stockham_bfly<2, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 3:  stockham_bfly<3, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 4:  stockham_bfly<4, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 5:  stockham_bfly<5, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 7:  stockham_bfly<7, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 8:  stockham_bfly<8, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 11: stockham_bfly<11, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 13: stockham_bfly<13, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    case 16: stockham_bfly<16, kLast>(b, N, s, tw_mul, tw, src, dst); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------------
// bf16 helpers (bit-level, host/device identical)
// ---------------------------------------------------------------------------------
SM_HD float bf16_bits_to_f32(uint32_t hi16) {
  union { uint32_t u; float f; } v; v.u = hi16 << 16; return v.f;
}
SM_HD uint32_t f32_bits(float f) { union { uint32_t u; float f; } v; v.f = f; return v.u; }
SM_HD float bits_f32(uint32_t u) { union { uint32_t u; float f; } v; v.u = u; return v.f; }
// round-to-nearest-even fp32 -> bf16 (same result as torch's .to(torch.bfloat16))
SM_HD uint32_t f32_to_bf16_rne(float f) {
  uint32_t u = f32_bits(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x0040u;  // quiet NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return u >> 16;
}

// pack two fp32 values into bf16x2 (low half = a), round to nearest even
SM_HD uint32_t pack_bf16x2_rne(float a, float b) {
#if defined(__CUDA_ARCH__)
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);     // one cvt.rn.bf16x2.f32
  return *reinterpret_cast<uint32_t*>(&h);
#else
  return f32_to_bf16_rne(a) | (f32_to_bf16_rne(b) << 16);
#endif
}
// true for NaN and +-Inf (one compare)
SM_HD bool not_finite(float v) { return !(fabsf(v) <= 3.4028234663852886e38f); }

}  // namespace smfft
