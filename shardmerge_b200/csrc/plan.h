// plan.h -- host-side FFT plan for one tensor shape [R][C] (C contiguous).
//
// The 2-D real FFT that the reference obtains from torch.fft.fftn(dim=(-2,-1))
// (shard/tensor/functions.py:58) is decomposed as
//   row pass : real FFT of length C on every row  -> Hermitian half spectrum, Ch+1 = C/2+1 bins
//   col pass : complex FFT of length R down every stored column, done either in one
//              sweep (R <= 256) or as a four-step split R = Ra*Rb in two in-place sweeps.
// With the two-sweep split the spectrum is left in a *permuted row order*:
//   stored row  Rb*ka + kb   holds frequency   ka + Ra*kb.
// The blend is element-wise plus global statistics, so the order is irrelevant to it,
// and the inverse sweeps consume exactly this order and hand back natural rows.
#pragma once
#include <cstdint>
#include <cstring>

#define SM_MAX_STAGES 8
#define SM_COL_TILE 32          // columns per column-pass CTA (one 128 B segment)
#define SM_COL_LMAX 256         // longest FFT done inside one column-pass CTA

struct SmPlan {
  int R, C, Ch, P;              // P: row pitch of the spectrum planes, in floats
  // row pass: complex FFT of length Ch (packed even/odd) + untangle
  int n_row; int row_rad[SM_MAX_STAGES];
  int row_threads; int row_pad;  // row_pad: 1 -> padded smem indexing
  int row_smem_fwd, row_smem_inv;
  // column passes
  int col_passes;               // 0 (R==1), 1, or 2
  int Ra, Rb;                   // sweep A: length Ra, row stride Rb; sweep B: length Rb, contiguous rows
  int nA; int radA[SM_MAX_STAGES];
  int nB; int radB[SM_MAX_STAGES];
  int thrA, thrB; int smemA, smemB;
};

#define SM_GENERIC_MAX 1021     // largest prime factor handled by the generic O(r) radix stage (fft_core.cuh)

static inline bool sm_is_small_radix(int r) {
  return r == 1 || r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 11 || r == 13 || r == 16;
}

static inline int sm_factor(int L, int* rad) {
  // odd primes <= 13 first (they make the stride-1 first stage bank-conflict free), then a balanced split of the power
  // of two into radices <= 16, then every larger prime factor as a generic radix (37 in 18944, 167 in 128256, ...).
  // A generic radix re-reads its inputs, so it must not be the first stage (whose source may accumulate, or be the very
  // global memory the stage writes): if nothing else precedes it, an identity stage of radix 1 does.  returns #stages or -1.
  static const int odd[] = {13, 11, 7, 5, 3};
  int n = 0;
  for (int i = 0; i < 5; ++i)
    while (L % odd[i] == 0) { if (n >= SM_MAX_STAGES) return -1; rad[n++] = odd[i]; L /= odd[i]; }
  int e = 0;
  while (L % 2 == 0) { L /= 2; ++e; }
  if (e > 0) {
    int ns = (e + 3) / 4;
    if (n + ns > SM_MAX_STAGES) return -1;
    int base = e / ns, extra = e % ns;
    for (int i = 0; i < ns; ++i) rad[n++] = 1 << (base + (i < extra ? 1 : 0));
  }
  if (L != 1) {
    if (n == 0) rad[n++] = 1;
    for (int f = 17; L > 1; ) {
      if (f > SM_GENERIC_MAX) return -1;
      if ((long long)f * f > L) f = L;             // what is left is prime
      if (L % f == 0) {
        if (f > SM_GENERIC_MAX || n >= SM_MAX_STAGES) return -1;
        rad[n++] = f; L /= f;
      } else {
        f += 2;
      }
    }
  }
  if (n == 0) { rad[n++] = 1; }  // length-1 transform: identity stage
  return n;
}

static inline int sm_count_generic(const int* rad, int n) {
  int g = 0;
  for (int i = 0; i < n; ++i) if (!sm_is_small_radix(rad[i])) ++g;
  return g;
}

static inline int sm_round_up(int x, int m) { return (x + m - 1) / m * m; }

// returns 0 on success, negative on unsupported shape
static inline int sm_make_plan(int R, int C, SmPlan* pl) {
  std::memset(pl, 0, sizeof(*pl));
  if (R < 1 || C < 2 || (C & 1)) return -1;
  pl->R = R; pl->C = C; pl->Ch = C / 2;
  pl->P = sm_round_up(pl->Ch + 1, 32);
  pl->n_row = sm_factor(pl->Ch, pl->row_rad);
  if (pl->n_row < 0) return -2;
  // threads: enough for the widest stage at ~1 butterfly per thread, capped at 512
  {
    int rmin = 16;
    for (int i = 0; i < pl->n_row; ++i) if (pl->row_rad[i] < rmin) rmin = pl->row_rad[i];
    (void)rmin;
    int t = sm_round_up((pl->Ch + 15) / 16, 32);
    if (t < 32) t = 32;
    if (t > 512) t = 512;
    pl->row_threads = t;
  }
  {
    const int kMaxSmem = 227 * 1024;
    int nbuf_f = pl->n_row >= 2 ? 2 : 1;
    int nbuf_i = pl->n_row >= 3 ? 2 : 1;
    int padded = pl->Ch + (pl->Ch >> 4) + 1;
    if ((long long)padded * 8 * nbuf_f <= kMaxSmem) {
      pl->row_pad = 1;
      pl->row_smem_fwd = padded * 8 * nbuf_f;
      pl->row_smem_inv = padded * 8 * nbuf_i;
    } else {
      pl->row_pad = 0;
      pl->row_smem_fwd = pl->Ch * 8 * nbuf_f;
      pl->row_smem_inv = pl->Ch * 8 * nbuf_i;
      if (pl->row_smem_fwd > kMaxSmem) return -3;
    }
  }
  if (R == 1) { pl->col_passes = 0; pl->Ra = 1; pl->Rb = 1; return 0; }
  int radR[SM_MAX_STAGES];
  int nR = sm_factor(R, radR);
  if (nR < 0) return -4;
  const int kMaxSmemCol = 227 * 1024;
  auto col_smem = [](int L, int nst) { return (long long)L * SM_COL_TILE * 8 * (nst >= 3 ? 2 : 1); };
  if (R <= SM_COL_LMAX && nR <= 3) {
    pl->col_passes = 1; pl->Ra = R; pl->Rb = 1;
    pl->nA = nR; std::memcpy(pl->radA, radR, sizeof(radR));
    pl->nB = 0;
  } else {
    // four-step split R = a * b.  Lengths up to SM_COL_LMAX with small radices are what the specialised kernels cover and
    // always win; longer sweeps (as far as one instance x 32 columns fits in shared memory) and generic radices only
    // serve shapes that have no such split (128256 = 501 x 256 with 501 = 3 x 167).
    int best = -1; long long bestcost = 1ll << 60;
    for (int a = 2; a <= 1024; ++a) {
      if (R % a) continue;
      int b = R / a;
      if (b > 1024 || b < 2) continue;
      int ra[SM_MAX_STAGES], rb[SM_MAX_STAGES];
      int na = sm_factor(a, ra), nb = sm_factor(b, rb);
      if (na < 0 || nb < 0) continue;
      if (col_smem(a, na) > kMaxSmemCol || col_smem(b, nb) > kMaxSmemCol) continue;
      // two register stages per sweep first (one smem exchange, and the shapes the specialised
      // kernels cover), then the fewest stages, then the most balanced split
      int off2 = (na > 2 ? na - 2 : 2 - na) + (nb > 2 ? nb - 2 : 2 - nb);
      long long cost = (long long)off2 * 65536 + (na + nb) * 1024 + (a > b ? a - b : b - a);
      cost += (long long)(sm_count_generic(ra, na) + sm_count_generic(rb, nb)) * (1ll << 24);
      if (a > SM_COL_LMAX) cost += 1ll << 28;
      if (b > SM_COL_LMAX) cost += 1ll << 28;
      if (cost < bestcost) { bestcost = cost; best = a; }
    }
    if (best < 0) return -5;
    pl->col_passes = 2; pl->Ra = best; pl->Rb = R / best;
    pl->nA = sm_factor(pl->Ra, pl->radA);
    pl->nB = sm_factor(pl->Rb, pl->radB);
  }
  auto threads_for = [](int L, const int* rad, int n) {
    int rmax = 1;
    for (int i = 0; i < n; ++i) if (rad[i] > rmax) rmax = rad[i];
    if (sm_count_generic(rad, n) > 0) rmax = 4;   // generic stages work per output, not per butterfly: as many warps as fit
    int warps = (L + rmax - 1) / rmax;      // butterflies of the widest-radix stage
    if (warps < 2) warps = 2;
    if (warps > 16) warps = 16;
    return warps * 32;
  };
  pl->thrA = threads_for(pl->Ra, pl->radA, pl->nA);
  pl->smemA = pl->Ra * SM_COL_TILE * 8 * (pl->nA >= 3 ? 2 : 1);
  if (pl->col_passes == 2) {
    pl->thrB = threads_for(pl->Rb, pl->radB, pl->nB);
    pl->smemB = pl->Rb * SM_COL_TILE * 8 * (pl->nB >= 3 ? 2 : 1);
  }
  return 0;
}

// stored row -> frequency index along R
static inline int sm_row_freq(const SmPlan* pl, int stored_row) {
  if (pl->col_passes < 2) return stored_row;
  int ka = stored_row / pl->Rb, kb = stored_row % pl->Rb;
  return ka + pl->Ra * kb;
}
