// kernels_fstats.cu -- fused statistics for the pair-merge chain (csrc/pipeline.cu).
//
// interpolate_fft_components (shard/tensor/functions.py:90-162) needs, between the forward and the
// inverse transform: the cutoff order statistic of cat(|Re X0|, |Re X1|) (:113-122), three masked
// sums for slerp() (:36-41 under the masks of :124-129), the blend (:134-136) and the cull order
// statistic of |Re R| (:138-148).  kernels_stats.cu does each of these as its own pass (that is
// the step-by-step API, 25 launches, 16N bytes); here they are folded into TWO streaming passes:
//
//   k_fs_sample<0>   64 K random keys of the full (mirrored) spectra; the last CTA to finish turns
//                    them into a key window [lo, hi] that brackets the cutoff statistic with ~6 sigma
//                    of sampling noise, and into the bin width of a 2048-bin histogram over it
//   k_fs_pass<0>     ONE pass over Re X0, Re X1 (4N bytes): keys below the window are counted, the
//                    ~1 % inside it are appended to per-bin buckets; SLERP sums are accumulated for
//                    every element whose mask the window already decides, the few with
//                    lo <= |re1| <= hi go to per-bin side buckets (+ per-bin partial sums).
//                    The last CTA finds the bin of the statistic, resolves the exact key inside that
//                    one bucket, closes the sums and computes dot / cos / sin / ||rel||.
//   k_fs_sample<1>   the same sampling for |blend(re0, re1)| (the blend is cheap to evaluate at a
//                    sample position) -> window for the cull statistic
//   k_fs_pass<1>     the blend itself (read 4N, write 2N) with the count / bucket step folded into its
//                    epilogue; the last CTA produces the exact cull threshold.
//
// Exactness: the thresholds are the bit patterns of actual elements at the exact rank (Hermitian
// multiplicities included), as in kernels_stats.cu; the masks therefore reproduce the reference's.
// A window that misses, is too wide, or overflows a bucket (degenerate distributions) sets status
// bits and the thresholds become NaN; the caller re-runs the tensor on the step-by-step path.
// Only tensors with more than 2^20 elements come here (sm_fstats_supported); smaller ones are launch
// bound and use the step-by-step kernels.
#include <cstddef>
#include <cstdlib>
#include "sm_internal.h"

namespace {

constexpr int kBins = 2048;
constexpr unsigned int kNS = 1u << 16;

struct FsState {                     // SM_FS_STATE_BYTES device bytes
  unsigned long long rank;           // rank of the statistic in the full key multiset
  unsigned long long below;          // keys below the window
  unsigned int lo, hi;               // window [lo, hi] (key bit patterns)
  unsigned int shift;                // bucket = (key - lo) >> shift
  unsigned int status;               // bit0 window missed, bit1 bucket overflow, bit2 window too wide   (SM_FS_STATUS_OFF)
  unsigned int ticket;
  unsigned int key;
  float value;
  unsigned int bstar;
  double s_in[3];                    // decided part of the SLERP sums (s00, s11, s01)
  unsigned int sticky;               // OR of every status this state has ended with (kernels never clear it)   (SM_FS_STICKY_OFF)
  unsigned int pad[13];
};
static_assert(sizeof(FsState) == SM_FS_STATE_BYTES, "FsState layout");
static_assert(offsetof(FsState, status) == SM_FS_STATUS_OFF, "FsState status offset");
static_assert(offsetof(FsState, sticky) == SM_FS_STICKY_OFF, "FsState sticky offset");

struct FsWs {                        // carved out of the caller's workspace
  unsigned long long* hc;            // [kBins]  hi 32 bits: keys (with multiplicity), lo 32 bits: bucket entries
  unsigned int* scnt;                // [kBins]  side bucket entries
  double* hs;                        // [kBins][3] per-bin partial SLERP sums of the side entries
  unsigned int* shist;               // [kSampleBins] sample histogram (zeroed by a memset node before k_fs_sample)
  unsigned int* bkt;                 // [kBins][bcap]  key | (multiplicity - 1) << 31
  float4* sbkt;                      // [kBins][scap]  (re0, re1, multiplicity, -)
  unsigned int bcap, scap;
  unsigned long long* dbg;           // development: per-CTA phase timestamps (NULL in production)
};

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FS_STAMP(slot) do { if (ws.dbg && threadIdx.x == 0) ws.dbg[(blockIdx.x) * 8 + (slot)] = gtimer(); } while (0)

__device__ __forceinline__ unsigned int absbits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ int sgn(float v) { return (v > 0.f) - (v < 0.f); }

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// the blend of one element (functions.py:124-136; one rounding per torch op)
struct BlendScal { float thr, dot, ct, sn, rn, t_sum; };
__device__ __forceinline__ float blend1(float a, float b, const BlendScal& s) {
  const bool same = (sgn(a) == sgn(b));
  if (same) {
    if (!(fabsf(b) < s.thr)) {
      const float rel = __fsub_rn(b, __fmul_rn(a, s.dot));
      return __fadd_rn(__fmul_rn(a, s.ct), __fmul_rn(__fdiv_rn(rel, s.rn), s.sn));
    }
    return __fadd_rn(a, __fmul_rn(s.t_sum, b));
  }
  return (fabsf(a) > fabsf(b)) ? a : b;
}

// ---------------------------------------------------------------------------------------------
// block-wide search of a histogram for the bin holding 0-based `rank` (NT threads).
// load(b) returns the count of bin b.
// ---------------------------------------------------------------------------------------------
struct Pick { int bin; unsigned long long rank_in_bin; };

template <int NT, class Load>
__device__ void block_pick(const Load& load, int nbins, unsigned long long rank, Pick* out /*shared*/) {
  __shared__ unsigned long long wtot[NT / 32];
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const int per = (nbins + NT - 1) / NT;
  if (t == 0) { out->bin = -1; out->rank_in_bin = 0ull; }
  unsigned long long v = 0ull;
  for (int j = 0; j < per; ++j) { const int b = t * per + j; if (b < nbins) v += load(b); }
  unsigned long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const unsigned long long w = lane < NT / 32 ? wtot[lane] : 0ull;
    unsigned long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long n = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += n;
    }
    if (lane < NT / 32) wtot[lane] = wi - w;
  }
  __syncthreads();
  unsigned long long excl = wtot[wid] + incl - v;
  if (rank >= excl && rank < excl + v) {
    for (int j = 0; j < per; ++j) {
      const int b = t * per + j;
      if (b >= nbins) break;
      const unsigned long long h = load(b);
      if (rank < excl + h) { out->bin = b; out->rank_in_bin = rank - excl; break; }
      excl += h;
    }
  }
  __syncthreads();
}

struct LoadSmem32 { const unsigned int* h; __device__ unsigned long long operator()(int b) const { return h[b]; } };
struct LoadGlobalHi { const unsigned long long* h; __device__ unsigned long long operator()(int b) const { return __ldcg(h + b) >> 32; } };

__device__ __forceinline__ double block_sum_d(double v, double* sh /* >= 32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < nw ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;                           // valid in warp 0
}

// scalars of slerp() from the three masked sums (functions.py:36-43); same arithmetic as k_slerp_scalars
__device__ void fs_scalars(const double* s, double t, float* scal4) {
  const double s00 = s[0], s11 = s[1], s01 = s[2];
  const float n0 = (float)sqrt(s00), n1 = (float)sqrt(s11);
  float dot = (float)s01 / (n0 * n1);
  dot = fminf(fmaxf(dot, -1.0f), 1.0f);
  if (!(dot == dot)) dot = __uint_as_float(0x7fc00000u);
  const float theta = (float)acos((double)dot) * (float)t;
  const float ct = (float)cos((double)theta), sn = (float)sin((double)theta);
  const double d = (double)dot;
  double rn2 = s11 - 2.0 * d * s01 + d * d * s00;
  if (rn2 < 0.0) rn2 = 0.0;
  float rn = (float)sqrt(rn2);
  if (rn < 1e-12f) rn = 1e-12f;
  if (!(d == d)) rn = __uint_as_float(0x7fc00000u);
  scal4[0] = dot; scal4[1] = ct; scal4[2] = sn; scal4[3] = rn;
}

// returns true in every thread of the last CTA to arrive
__device__ __forceinline__ bool fs_last_block(FsState* st) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int n = gridDim.x * gridDim.y;
    s_last = (atomicAdd(&st->ticket, 1u) == n - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ---------------------------------------------------------------------------------------------
// sampling + window
// ---------------------------------------------------------------------------------------------
struct FsCommon {
  const float* reX; const float* reY; const int* sel;     // role pick: *sel != 0 -> (re0, re1) = (reY, reX)
  const float* thr_cut; const float* scal4; float t_sum;  // MODE 1 only
};

constexpr int kSampleShift = 18;                         // sample histogram digit: key bits [30:18] (exponent + 5 mantissa bits)
constexpr int kSampleBins = 1 << (31 - kSampleShift);    // 8192

struct LoadGlobal32 { const unsigned int* h; __device__ unsigned long long operator()(int b) const { return __ldcg(h + b); } };

// 64 K random keys of the full (mirrored) spectrum are histogrammed by their top 14 key bits with global
// reductions (spread over the L2 slices, no single-SM bottleneck); the last CTA to finish copies the 8 K bins to
// shared memory (all loads in flight at once: single-CTA epilogues must not chain L2 round trips), finds the bins
// holding the sample ranks k_lo / k_hi and takes their outer edges as the key window: at most one bin (3 % of the
// key value) wider per side than the exact sample statistics would give.
template <int MODE>
__global__ void __launch_bounds__(1024) k_fs_sample(const __grid_constant__ SmPlan pl, const __grid_constant__ FsCommon c,
                                                    FsState* st, const __grid_constant__ FsWs ws, unsigned long long rank,
                                                    long long k_lo, long long k_hi) {
  __shared__ Pick out_a, out_b;
  __shared__ unsigned int s_h[kSampleBins];
  const unsigned int ns = kNS;
  const unsigned long long per_plane = (unsigned long long)pl.R * (unsigned long long)pl.C;
  const unsigned long long total = per_plane * (MODE == 0 ? 2ull : 1ull);
  const bool sw = (c.sel != nullptr && *c.sel != 0);
  const float* re0 = sw ? c.reY : c.reX;
  const float* re1 = sw ? c.reX : c.reY;
  BlendScal bs{};
  if (MODE == 1) { bs.thr = *c.thr_cut; bs.dot = c.scal4[0]; bs.ct = c.scal4[1]; bs.sn = c.scal4[2]; bs.rn = c.scal4[3]; bs.t_sum = c.t_sum; }
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  FS_STAMP(0);
#pragma unroll
  for (int j = 0; j < kSampleBins / 1024; ++j) s_h[threadIdx.x + 1024 * j] = 0u;
  __syncthreads();
  // this kernel also zeroes the bucket counters / per-bin sums of the pass that follows
  for (unsigned int z = i; z < (unsigned int)kBins; z += gridDim.x * blockDim.x) {
    ws.hc[z] = 0ull; ws.scnt[z] = 0u; ws.hs[3 * z] = 0.0; ws.hs[3 * z + 1] = 0.0; ws.hs[3 * z + 2] = 0.0;
  }
  if (i < ns) {
    unsigned long long j = __umul64hi(mix64(i + (MODE ? 0x51ed27ull : 0ull)), total);
    const float* pp = re0;
    if (MODE == 0 && j >= per_plane) { j -= per_plane; pp = re1; }
    const unsigned int row = (unsigned int)(j / (unsigned long long)pl.C);
    unsigned int col = (unsigned int)(j - (unsigned long long)row * pl.C);
    if (col > (unsigned int)pl.Ch) col = pl.C - col;
    const size_t off = (size_t)row * pl.P + col;
    const unsigned int key = (MODE == 0) ? absbits(pp[off]) : absbits(blend1(re0[off], re1[off], bs));
    // CTA-local histogram first: the populated bins are few, and same-address reductions serialise in L2
    // (measured: 64 K direct global REDs took 12-17 us; one RED per CTA and non-empty bin takes < 1 us)
    atomicAdd(&s_h[key >> kSampleShift], 1u);
  }
  __syncthreads();
  FS_STAMP(1);
#pragma unroll
  for (int j = 0; j < kSampleBins / 1024; ++j) {
    const unsigned int v = s_h[threadIdx.x + 1024 * j];
    if (v) atomicAdd(ws.shist + threadIdx.x + 1024 * j, v);
  }
  FS_STAMP(2);
  const bool last = fs_last_block(st);
  FS_STAMP(3);
  if (!last) return;
  const bool want_a = (k_lo >= 0 && k_lo < (long long)ns), want_b = (k_hi >= 0 && k_hi < (long long)ns);
  {
    unsigned int v[kSampleBins / 1024];
#pragma unroll
    for (int j = 0; j < kSampleBins / 1024; ++j) v[j] = __ldcg(ws.shist + threadIdx.x + 1024 * j);
#pragma unroll
    for (int j = 0; j < kSampleBins / 1024; ++j) s_h[threadIdx.x + 1024 * j] = v[j];
  }
  __syncthreads();
  block_pick<1024>(LoadSmem32{s_h}, kSampleBins, want_a ? (unsigned long long)k_lo : 0ull, &out_a);
  block_pick<1024>(LoadSmem32{s_h}, kSampleBins, want_b ? (unsigned long long)k_hi : 0ull, &out_b);
  const int bin_a = out_a.bin, bin_b = out_b.bin;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int lo = (want_a && bin_a >= 0) ? ((unsigned int)bin_a << kSampleShift) : 0u;
    unsigned int hi = (want_b && bin_b >= 0) ? ((((unsigned int)bin_b + 1u) << kSampleShift) - 1u) : 0x7f800000u;
    if (hi > 0x7f800000u) hi = 0x7f800000u;             // NaN keys stay above every window
    unsigned int status = 0u;
    const unsigned long long width = hi >= lo ? (unsigned long long)hi - lo + 1ull : 0ull;
    if (width == 0ull) status |= 4u;
    unsigned int l2 = 0;
    while ((1ull << l2) < width) ++l2;
    st->rank = rank; st->below = 0ull; st->lo = lo; st->hi = hi;
    st->shift = l2 > 11 ? l2 - 11 : 0;
    st->status = status; st->key = 0u; st->value = 0.f; st->bstar = 0u;
    st->s_in[0] = 0.0; st->s_in[1] = 0.0; st->s_in[2] = 0.0;
    st->ticket = 0u;
  }
  FS_STAMP(4);
}

// ---------------------------------------------------------------------------------------------
// the streaming pass
// ---------------------------------------------------------------------------------------------
constexpr int kCandStage = 2048, kCandFlush = 1024;   // staged per CTA; flushed (uniformly) once half full
constexpr int kSideStage = 512, kSideFlush = 256;

struct PassCtx {
  float lo_f, hi_f;                  // the window as floats: for non-NaN keys, float order == bit-pattern order
  unsigned int lo, span, shift, bcap, scap;
  unsigned long long* hc; unsigned int* bkt;
  unsigned int* scnt; float4* sbkt; double* hs;
};

// one key of multiplicity w: count it below the window, or stage it for its bucket (a shared-memory atomic; the
// global bucket append -- a returning L2 atomic -- is batched in fs_flush so that no warp stalls on it per key)
// CTA staging of the pass kernel, at namespace scope so that the out-of-line push helpers need no pointer arguments
__shared__ unsigned int g_fs_hist[kBins];               // streaming phase: staged candidates; last CTA: bucket histogram
__shared__ float4 g_fs_side[kSideStage];
__shared__ unsigned int g_fs_ncand, g_fs_nside;
static_assert(kCandStage <= kBins, "candidate staging aliases the histogram");

__device__ __noinline__ void fs_push_cand(unsigned int e) {
  const unsigned int pos = atomicAdd(&g_fs_ncand, 1u);
  if (pos < (unsigned int)kCandStage) g_fs_hist[pos] = e;
}
__device__ __noinline__ void fs_push_side(float a, float b, float wf) {
  const unsigned int pos = atomicAdd(&g_fs_nside, 1u);
  if (pos < (unsigned int)kSideStage) g_fs_side[pos] = make_float4(a, b, wf, 0.f);
}
// torch.sign(a) == torch.sign(b) (sign(+-0) = sign(NaN) = 0).  A non-zero product decides it at once; the exact
// comparison only runs for zeros, NaNs and products that underflow (a rarely taken, warp-coherent branch).
__device__ __forceinline__ bool same_sign(float a, float b) {
  const float p = a * b;
  bool same = p > 0.f;
  if (!(p > 0.f) && !(p < 0.f)) same = (sgn(a) == sgn(b));
  return same;
}

// one element pair (MODE 0: statistics, MODE 1: blend), multiplicity w (0 for the zero-filled padding columns).
// The common work is predicated / select based -- the three masks split a warp roughly 46 / 4 / 50 %, so branching
// on them would run every side for every warp anyway -- and only the rare events (a key or an element inside the
// window) branch, out of line.  Window tests are float compares on |v| (free abs modifier); NaN keys compare
// false everywhere, i.e. they sit above the window like their bit patterns do.
template <int MODE>
__device__ __forceinline__ float fs_elem(const PassCtx& x, const BlendScal& bs, float a, float b, unsigned int w,
                                         unsigned int& below, float& p00, float& p11, float& p01) {
  const bool same = same_sign(a, b);
  if (MODE == 0) {
    below += (fabsf(a) < x.lo_f ? w : 0u) + (fabsf(b) < x.lo_f ? w : 0u);
    if (fabsf(a) >= x.lo_f && fabsf(a) <= x.hi_f && w) fs_push_cand(absbits(a) | ((w - 1u) << 31));
    const bool b_in_window = fabsf(b) >= x.lo_f && fabsf(b) <= x.hi_f && w;
    if (b_in_window) fs_push_cand(absbits(b) | ((w - 1u) << 31));
    const float wf = (float)w;
    if (same && !(fabsf(b) <= x.hi_f)) {                 // |re1| >= thr for every thr in the window (NaN: never < thr)
      p00 = fmaf(wf * a, a, p00); p11 = fmaf(wf * b, b, p11); p01 = fmaf(wf * a, b, p01);
    }
    if (same && b_in_window) fs_push_side(a, b, wf);                // undecided until the exact threshold is known
    return 0.f;
  } else {
    // functions.py:134-136, one rounding per torch op (same arithmetic as blend1 / k_blend)
    const float rel = __fsub_rn(b, __fmul_rn(a, bs.dot));
    const float o_slerp = __fadd_rn(__fmul_rn(a, bs.ct), __fmul_rn(__fdiv_rn(rel, bs.rn), bs.sn));
    const float o_sum = __fadd_rn(a, __fmul_rn(bs.t_sum, b));
    const float o_big = (fabsf(a) > fabsf(b)) ? a : b;
    const float o = same ? ((fabsf(b) < bs.thr) ? o_sum : o_slerp) : o_big;
    below += fabsf(o) < x.lo_f ? w : 0u;
    if (fabsf(o) >= x.lo_f && fabsf(o) <= x.hi_f && w) fs_push_cand(absbits(o) | ((w - 1u) << 31));
    return o;
  }
}

// all threads of the CTA: append the staged entries to their global buckets; returns the overflow flag
template <int MODE>
__device__ unsigned int fs_flush(const PassCtx& x) {
  unsigned int* s_cand = g_fs_hist; float4* s_side = g_fs_side;
  unsigned int ovf = 0u;
  unsigned int nc = g_fs_ncand, nsd = (MODE == 0) ? g_fs_nside : 0u;
  if (nc > (unsigned int)kCandStage) { nc = kCandStage; ovf = 2u; }
  if (nsd > (unsigned int)kSideStage) { nsd = kSideStage; ovf = 2u; }
  for (unsigned int i = threadIdx.x; i < nc; i += blockDim.x) {
    const unsigned int e = s_cand[i], k = e & 0x7fffffffu, w = 1u + (e >> 31);
    const unsigned int bin = (k - x.lo) >> x.shift;
    const unsigned int pos = (unsigned int)atomicAdd(x.hc + bin, ((unsigned long long)w << 32) | 1ull);
    if (pos < x.bcap) x.bkt[(size_t)bin * x.bcap + pos] = e; else ovf = 2u;
  }
  if (MODE == 0) {
    for (unsigned int i = threadIdx.x; i < nsd; i += blockDim.x) {
      const float4 e = s_side[i];
      const unsigned int bin = (absbits(e.y) - x.lo) >> x.shift;
      const unsigned int pos = atomicAdd(x.scnt + bin, 1u);
      if (pos < x.scap) x.sbkt[(size_t)bin * x.scap + pos] = e; else ovf = 2u;
      const double da = (double)e.x, db = (double)e.y, dw = (double)e.z;
      atomicAdd(x.hs + 3 * bin, dw * da * da); atomicAdd(x.hs + 3 * bin + 1, dw * db * db);
      atomicAdd(x.hs + 3 * bin + 2, dw * da * db);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { g_fs_ncand = 0u; g_fs_nside = 0u; }
  __syncthreads();
  return ovf;
}

template <int MODE>
__global__ void __launch_bounds__(SM_EW_THREADS, 4) k_fs_pass(const __grid_constant__ SmPlan pl, const __grid_constant__ FsCommon c,
                                                           FsState* st, const __grid_constant__ FsWs ws, float* out,
                                                           double t, float* thr_out, float* scal4, double* sums_out) {
  __shared__ unsigned long long s_below[SM_EW_THREADS / 32];
  __shared__ double s_red[32];
  __shared__ Pick pick;
  __shared__ unsigned long long s_rank;
  __shared__ int s_ok;
  unsigned int* const s_hist = g_fs_hist;
  if (threadIdx.x == 0) { g_fs_ncand = 0u; g_fs_nside = 0u; }
  __syncthreads();
  FS_STAMP(0);
  const bool sw = (c.sel != nullptr && *c.sel != 0);
  const float* __restrict__ re0 = sw ? c.reY : c.reX;
  const float* __restrict__ re1 = sw ? c.reX : c.reY;
  PassCtx x;
  x.lo = st->lo; x.span = st->hi - x.lo; x.shift = st->shift; x.bcap = ws.bcap; x.scap = ws.scap;
  x.lo_f = __uint_as_float(x.lo); x.hi_f = __uint_as_float(st->hi);       // hi <= +inf (k_fs_sample)
  x.hc = ws.hc; x.bkt = ws.bkt; x.scnt = ws.scnt; x.sbkt = ws.sbkt; x.hs = ws.hs;
  const bool dead = st->status != 0u;
  BlendScal bs{};
  if (MODE == 1) { bs.thr = *c.thr_cut; bs.dot = c.scal4[0]; bs.ct = c.scal4[1]; bs.sn = c.scal4[2]; bs.rn = c.scal4[3]; bs.t_sum = c.t_sum; }
  const int Ch = pl.Ch;
  unsigned int below32 = 0u;
  unsigned int ovf = 0u;
  double d00 = 0.0, d11 = 0.0, d01 = 0.0;
  // Work items are float4 column groups, numbered row-major over rows x G groups (G covers columns 0..Ch); the grid
  // strides over them so every thread gets the same share (no idle column blocks, no tail wave), two items in flight.
  const unsigned int G = (unsigned int)(Ch + 4) / 4u;
  const unsigned int total = (unsigned int)pl.R * G;
  const unsigned int stride = gridDim.x * blockDim.x;
  const unsigned int step_row = stride / G, step_g = stride - step_row * G;
  if (!dead) {
    int it = 0;
    unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int row = idx / G, g = idx - row * G;
    for (unsigned int base = blockIdx.x * blockDim.x; base < total; base += 2u * stride, ++it) {   // uniform trip count
      // second item of this iteration: one grid stride further
      unsigned int row1 = row + step_row, g1 = g + step_g;
      if (g1 >= G) { g1 -= G; ++row1; }
      const int n_item = idx < total ? (idx + stride < total ? 2 : 1) : 0;
      if (n_item > 0) {
        const size_t off[2] = {(size_t)row * pl.P + 4u * g, n_item == 2 ? (size_t)row1 * pl.P + 4u * g1 : (size_t)row * pl.P + 4u * g};
        float4 a4[2], b4[2];                             // both items in flight before the first is consumed
        a4[0] = *reinterpret_cast<const float4*>(re0 + off[0]); b4[0] = *reinterpret_cast<const float4*>(re1 + off[0]);
        a4[1] = *reinterpret_cast<const float4*>(re0 + off[1]); b4[1] = *reinterpret_cast<const float4*>(re1 + off[1]);
#pragma unroll 1
        for (int r = 0; r < n_item; ++r) {               // rolled: one copy of the element code keeps the loop in the I-cache
          const float4 av = r == 0 ? a4[0] : a4[1], bv = r == 0 ? b4[0] : b4[1];
          const int c0 = 4 * (int)(r == 0 ? g : g1);
          const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
          float p00 = 0.f, p11 = 0.f, p01 = 0.f;         // fp32 over one float4, fp64 across
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // multiplicity: 1 at columns 0 and Ch, 2 inside, 0 for the padding (zero-filled by the plane allocator and
            // never written by any kernel, so padding elements contribute nothing)
            const int col = c0 + i;
            const unsigned int w = col > Ch ? 0u : ((col == 0 || col == Ch) ? 1u : 2u);
            o[i] = fs_elem<MODE>(x, bs, a[i], b[i], w, below32, p00, p11, p01);
          }
          if (MODE == 0) { d00 += (double)p00; d11 += (double)p11; d01 += (double)p01; }
          else *reinterpret_cast<float4*>(out + (r == 0 ? off[0] : off[1])) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      // advance this thread by two grid strides
      idx += 2u * stride;
      row = row1 + step_row; g = g1 + step_g;
      if (g >= G) { g -= G; ++row; }
      if ((it & 7) == 7) {                               // every 8 iterations: flush the staging if it is half full
        __syncthreads();
        if (g_fs_ncand >= (unsigned int)kCandFlush || g_fs_nside >= (unsigned int)kSideFlush)
          ovf |= fs_flush<MODE>(x);
        else __syncthreads();                            // nobody appends before everybody has read the counters
      }
    }
    __syncthreads();
    FS_STAMP(1);
    ovf |= fs_flush<MODE>(x);
    FS_STAMP(2);
  }
  unsigned long long below = below32;
  if (ovf) atomicOr(&st->status, ovf);
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if ((threadIdx.x & 31) == 0) s_below[threadIdx.x >> 5] = below;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long bsum = 0ull;
      for (int i = 0; i < SM_EW_THREADS / 32; ++i) bsum += s_below[i];
      if (bsum) atomicAdd(&st->below, bsum);
    }
  }
  if (MODE == 0) {
    const double r0 = block_sum_d(d00, s_red), r1 = block_sum_d(d11, s_red), r2 = block_sum_d(d01, s_red);
    if (threadIdx.x == 0 && (r0 != 0.0 || r1 != 0.0 || r2 != 0.0)) {
      atomicAdd(&st->s_in[0], r0); atomicAdd(&st->s_in[1], r1); atomicAdd(&st->s_in[2], r2);
    }
  }
  FS_STAMP(3);
  const bool last_cta = fs_last_block(st);
  FS_STAMP(4);
  if (!last_cta) return;

  // ---- last CTA: bin of the statistic -> exact key inside that bucket -> (MODE 0) close the sums
  constexpr int NT = SM_EW_THREADS;
  if (threadIdx.x == 0) {
    const unsigned long long r = st->rank, bl = *((volatile unsigned long long*)&st->below);
    unsigned int status = *((volatile unsigned int*)&st->status);
    if (r < bl) { status |= 1u; s_rank = 0ull; } else s_rank = r - bl;
    st->status = status;
    s_ok = status == 0u ? 1 : 0;
  }
  __syncthreads();
  bool ok = s_ok != 0;                                    // block-uniform from here on
  unsigned int key = 0u;
  int bstar = -1;
  if (ok) {
    {  // key counts per bucket -> shared memory, every load in flight at once
      unsigned long long v[kBins / NT];
#pragma unroll
      for (int j = 0; j < kBins / NT; ++j) v[j] = __ldcg(ws.hc + threadIdx.x + NT * j);
#pragma unroll
      for (int j = 0; j < kBins / NT; ++j) s_hist[threadIdx.x + NT * j] = (unsigned int)(v[j] >> 32);
    }
    __syncthreads();
    block_pick<NT>(LoadSmem32{s_hist}, kBins, s_rank, &pick);
    bstar = pick.bin;
    if (bstar < 0) ok = false;                            // rank beyond the window: the sample window missed
  }
  if (ok) {
    // descend inside the bucket: 2048-bin shared histograms until one bin is one key (one level when the
    // window spans <= 2^22 bit patterns, the usual case)
    unsigned long long rin = pick.rank_in_bin;
    unsigned int n_b = (unsigned int)__ldcg(ws.hc + bstar);
    if (n_b > x.bcap) n_b = x.bcap;                       // overflow is already flagged in status
    unsigned int cur_lo = x.lo + ((unsigned int)bstar << x.shift), wlog = x.shift;
    while (true) {
      const unsigned int sh2 = wlog > 11 ? wlog - 11 : 0;
      for (int b = threadIdx.x; b < kBins; b += NT) s_hist[b] = 0u;
      __syncthreads();
      for (unsigned int i = threadIdx.x; i < n_b; i += NT) {
        const unsigned int e = __ldcg(ws.bkt + (size_t)bstar * x.bcap + i);
        const unsigned int d = (e & 0x7fffffffu) - cur_lo;              // unsigned wrap: keys below cur_lo fall out
        if ((d >> wlog) == 0u) atomicAdd(&s_hist[d >> sh2], 1u + (e >> 31));
      }
      __syncthreads();
      block_pick<NT>(LoadSmem32{s_hist}, kBins, rin, &pick);
      if (pick.bin < 0) { ok = false; break; }
      cur_lo += (unsigned int)pick.bin << sh2;
      rin = pick.rank_in_bin;
      if (sh2 == 0u) { key = cur_lo; break; }
      wlog = sh2;
    }
  }
  if (MODE == 0 && ok) {
    // side entries: bins above the statistic's bin are "in" (per-bin partial sums); inside the bin decide per entry
    double acc[3] = {0.0, 0.0, 0.0};
    {
      double h[kBins / NT][3];
#pragma unroll
      for (int j = 0; j < kBins / NT; ++j) {
        const int b = threadIdx.x + NT * j;
#pragma unroll
        for (int q = 0; q < 3; ++q) h[j][q] = (b > bstar) ? __ldcg(ws.hs + 3 * b + q) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < kBins / NT; ++j) { acc[0] += h[j][0]; acc[1] += h[j][1]; acc[2] += h[j][2]; }
    }
    unsigned int ns_b = __ldcg(ws.scnt + bstar);
    if (ns_b > x.scap) ns_b = x.scap;
    for (unsigned int i = threadIdx.x; i < ns_b; i += NT) {
      const float4 e = __ldcg(ws.sbkt + (size_t)bstar * x.scap + i);
      const double a = (double)e.x, b = (double)e.y, w = (double)e.z;
      if (absbits(e.y) >= key) { acc[0] += w * a * a; acc[1] += w * b * b; acc[2] += w * a * b; }
      // entries below the key were added to the bin's partial sums but bins <= bstar are not summed above
    }
    for (int j = 0; j < 3; ++j) {
      const double r = block_sum_d(acc[j], s_red);
      if (threadIdx.x == 0) st->s_in[j] = *((volatile double*)&st->s_in[j]) + r;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!ok && st->status == 0u) st->status = 1u;
    st->sticky |= st->status;
    st->ticket = 0u;
    if (st->status != 0u) {
      st->value = __uint_as_float(0x7fc00000u);
      if (thr_out) *thr_out = st->value;
    } else {
      st->key = key; st->value = __uint_as_float(key); st->bstar = (unsigned int)bstar;
      if (thr_out) *thr_out = st->value;
      if (MODE == 0) {
        double s[3] = {st->s_in[0], st->s_in[1], st->s_in[2]};
        if (sums_out) { sums_out[0] = s[0]; sums_out[1] = s[1]; sums_out[2] = s[2]; }
        if (scal4) fs_scalars(s, t, scal4);
      }
    }
  }
  FS_STAMP(5);
}

inline dim3 fs_grid(const SmPlan& p) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)p.R * ((p.Ch + 4) / 4);
  long long ctas = (long long)sms * 4;                   // one resident wave at 4 CTAs of 256 threads per SM (launch bounds)
  const long long need = (items + 2 * SM_EW_THREADS - 1) / (2 * SM_EW_THREADS);
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  return dim3((unsigned int)ctas, 1);
}

void fs_sample_ranks(unsigned long long rank, unsigned long long total, long long* k_lo, long long* k_hi) {
  const double pq = (double)rank / (double)total;
  const double ks = pq * (double)kNS;
  const double delta = 6.0 * sqrt((double)kNS * pq * (1.0 - pq)) + 16.0;
  *k_lo = (long long)floor(ks - delta); *k_hi = (long long)ceil(ks + delta);
}

// bucket capacities: the window holds ~frac of the keys, spread over >= half of the 2048 buckets; x4 headroom.
// Sized for the widest window any rank can need (p = 0.5).
double fs_max_frac() { return 2.0 * (6.0 * sqrt((double)kNS * 0.25) + 16.0) / (double)kNS; }
unsigned int fs_bcap(const SmPlan& p) {
  const double entries = fs_max_frac() * (double)p.R * (double)(p.Ch + 1) * 2.0;      // two planes, one entry per stored bin
  return (unsigned int)(entries / 1024.0 * 4.0) + 64u;
}
unsigned int fs_scap(const SmPlan& p) {
  const double entries = fs_max_frac() * (double)p.R * (double)(p.Ch + 1);
  return (unsigned int)(entries / 1024.0 * 4.0) + 64u;
}

}  // namespace

// ================================================================== C ABI
extern "C" int sm_fstats_supported(const sm_plan* plan) {
  const SmPlan& p = plan->p;
  const unsigned long long n = (unsigned long long)p.R * p.C;
  return (n > (1ull << 20) && 2ull * n < 0xfffffff0ull) ? 1 : 0;
}

extern "C" size_t sm_fstats_ws_bytes(const sm_plan* plan) {
  const SmPlan& p = plan->p;
  if (!sm_fstats_supported(plan)) return 256;
  size_t b = (size_t)kBins * (8 + 24 + 4) + (size_t)(1 << 13) * 4;
  b += (size_t)kBins * fs_scap(p) * 16;
  b += (size_t)kBins * fs_bcap(p) * 4;
  return b + 512 + 65536;        // + development timestamps (SM_FS_STAMPS)
}

static int fs_carve(const sm_plan* plan, void* wsp, size_t ws_bytes, FsWs* w) {
  const SmPlan& p = plan->p;
  if (!sm_fstats_supported(plan)) { sm_set_error("fstats: tensor too small / too large for the fused statistics"); return -2; }
  if (ws_bytes < sm_fstats_ws_bytes(plan)) { sm_set_error("fstats: workspace too small"); return -3; }
  char* b = reinterpret_cast<char*>(wsp);
  b = reinterpret_cast<char*>(((uintptr_t)b + 63) / 64 * 64);
  w->hc = reinterpret_cast<unsigned long long*>(b); b += (size_t)kBins * 8;
  w->hs = reinterpret_cast<double*>(b); b += (size_t)kBins * 3 * 8;
  w->scnt = reinterpret_cast<unsigned int*>(b); b += (size_t)kBins * 4;
  w->shist = reinterpret_cast<unsigned int*>(b); b += (size_t)(1 << 13) * 4;
  w->bcap = fs_bcap(p); w->scap = fs_scap(p);
  w->sbkt = reinterpret_cast<float4*>(b); b += (size_t)kBins * w->scap * 16;
  w->bkt = reinterpret_cast<unsigned int*>(b);
  b += (size_t)kBins * w->bcap * 4;
  b = reinterpret_cast<char*>(((uintptr_t)b + 63) / 64 * 64);
  w->dbg = getenv("SM_FS_STAMPS") ? reinterpret_cast<unsigned long long*>(b) : nullptr;   // development only
  return 0;
}

// Cutoff statistic + SLERP sums + scalars in one streaming pass (see the header of this file).
extern "C" int sm_fstats_cutoff(const sm_plan* plan, const float* reX, const float* reY, const int* sel, uint64_t rank,
                                double t, void* fs_state, void* ws, size_t ws_bytes, float* thr_cut_out,
                                float* scal4_out, double* sums3_out, void* stream) {
  const SmPlan& p = plan->p;
  cudaStream_t s = (cudaStream_t)stream;
  FsWs w;
  int rc = fs_carve(plan, ws, ws_bytes, &w);
  if (rc) return rc;
  const unsigned long long total = 2ull * (unsigned long long)p.R * p.C;
  if (rank >= total) rank = total - 1;
  FsState* st = reinterpret_cast<FsState*>(fs_state);
  FsCommon c{reX, reY, sel, nullptr, nullptr, 1.f};
  long long k_lo = 0, k_hi = 0;
  fs_sample_ranks(rank, total, &k_lo, &k_hi);
  SM_CUDA_CHECK(cudaMemsetAsync(w.shist, 0, (size_t)kSampleBins * 4, s));
  k_fs_sample<0><<<(int)(kNS / 1024), 1024, 0, s>>>(p, c, st, w, rank, k_lo, k_hi);
  SM_LAUNCH_CHECK();
  if (getenv("SM_FS_ONLY_SAMPLE")) return 0;
  k_fs_pass<0><<<fs_grid(p), SM_EW_THREADS, 0, s>>>(p, c, st, w, nullptr, t, thr_cut_out, scal4_out, sums3_out);
  SM_LAUNCH_CHECK();
  return 0;
}

// Blend (SLERP mode) + cull statistic of its output in one streaming pass.
extern "C" int sm_fstats_blend_cull(const sm_plan* plan, const float* reX, const float* reY, const int* sel,
                                    const float* thr_cut, const float* scal4, float t_sum, float* out_re, uint64_t rank,
                                    void* fs_state, void* ws, size_t ws_bytes, float* thr_cull_out, void* stream) {
  const SmPlan& p = plan->p;
  cudaStream_t s = (cudaStream_t)stream;
  FsWs w;
  int rc = fs_carve(plan, ws, ws_bytes, &w);
  if (rc) return rc;
  const unsigned long long total = (unsigned long long)p.R * p.C;
  if (rank >= total) rank = total - 1;
  FsState* st = reinterpret_cast<FsState*>(fs_state);
  FsCommon c{reX, reY, sel, thr_cut, scal4, t_sum};
  long long k_lo = 0, k_hi = 0;
  fs_sample_ranks(rank, total, &k_lo, &k_hi);
  SM_CUDA_CHECK(cudaMemsetAsync(w.shist, 0, (size_t)kSampleBins * 4, s));
  k_fs_sample<1><<<(int)(kNS / 1024), 1024, 0, s>>>(p, c, st, w, rank, k_lo, k_hi);
  SM_LAUNCH_CHECK();
  k_fs_pass<1><<<fs_grid(p), SM_EW_THREADS, 0, s>>>(p, c, st, w, out_re, 0.0, thr_cull_out, nullptr, nullptr);
  SM_LAUNCH_CHECK();
  return 0;
}
