// hostemu.cpp -- TEST-ONLY CPU emulation of the FFT sweep bodies.
//
// Compiles shardmerge_b200/csrc/fft_bodies.cuh with g++ and runs every CTA of every sweep
// as a sequential loop over emulated threads, so the index math, twiddles, butterflies,
// tangle/untangle and epilogue are checked against numpy on a machine without a GPU.
// Nothing in the product imports or links this file; it exists only for
// tests/test_hostemu.py (pytest -m "not gpu").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../shardmerge_b200/csrc/fft_bodies.cuh"

using namespace smfft;

struct HostExec {
  int T;
  int nthreads() const { return T; }
  int tid_begin() const { return 0; }
  int tid_end() const { return T; }
  void sync() const {}
};

static void make_tw(int M, std::vector<cf>& tw) {
  tw.resize(M > 0 ? M : 1);
  for (int j = 0; j < M; ++j) {
    // exact at multiples of a quarter turn, like sincospi on the device
    double a = 2.0 * (double)j / (double)M;   // in units of pi
    double c = std::cos(M_PI * a), s = std::sin(M_PI * a);
    if ((4LL * j) % M == 0) {
      int q = (int)((4LL * j) / M);
      c = (q == 0) ? 1.0 : (q == 2 ? -1.0 : 0.0);
      s = (q == 1) ? 1.0 : (q == 3 ? -1.0 : 0.0);
    }
    tw[j].x = (float)c; tw[j].y = (float)(-s);
  }
}

static void make_quads(const SmPlan& pl, std::vector<cf>& q) {
  const int nb = pl.n_row > 1 ? pl.Ch / pl.row_rad[0] : 0, N = pl.Ch;
  q.resize(4 * (size_t)nb + 4);
  for (int b = 0; b < nb; ++b)
    for (int j = 0; j < 4; ++j) {
      const long long e = ((long long)b << j) % N;
      const double a = 2.0 * M_PI * (double)e / (double)N;
      q[4 * (size_t)b + j].x = (float)std::cos(a); q[4 * (size_t)b + j].y = (float)(-std::sin(a));
    }
}

extern "C" {

int emu_plan(int R, int C, int* out /* >= 64 ints */) {
  SmPlan pl;
  int rc = sm_make_plan(R, C, &pl);
  if (rc) return rc;
  std::memcpy(out, &pl, sizeof(pl) < 64 * sizeof(int) ? sizeof(pl) : 64 * sizeof(int));
  return 0;
}
int emu_plan_pitch(int R, int C) { SmPlan pl; if (sm_make_plan(R, C, &pl)) return -1; return pl.P; }
int emu_row_freq(int R, int C, int stored) { SmPlan pl; if (sm_make_plan(R, C, &pl)) return -1; return sm_row_freq(&pl, stored); }

// forward: rows (delta or fp32) then column sweeps, spectrum scaled by `scale`
int emu_forward(int R, int C, int mode, const uint16_t* base, const uint16_t* ft, const float* x32,
                float m1, float m2, float scale, float* re, float* im, double* sumsq) {
  SmPlan pl;
  int rc = sm_make_plan(R, C, &pl);
  if (rc) return rc;
  std::vector<cf> twC, twR, twQ;
  make_tw(C, twC); make_tw(R, twR); make_quads(pl, twQ);
  std::vector<cf> smem(2 * (size_t)(pl.Ch + (pl.Ch >> 4) + 1) + 16);
  RowFwdArgs a{};
  a.mode = mode; a.base = base; a.ft = ft; a.x32 = x32; a.m1 = m1; a.m2 = m2; a.re = re; a.im = im;
  double acc = 0.0;
  for (int row = 0; row < R; ++row) {
    HostExec ex{pl.row_threads};
    float part = 0.f;                      // the device widens one fp32 partial per thread; here per row
    row_fwd_body(ex, pl, row, a, twC.data(), twQ.data(), smem.data(), &part);
    acc += (double)part;
  }
  *sumsq = acc;
  for (int sweep = 0; sweep < pl.col_passes; ++sweep) {
    ColArgs ca{};
    int n_inst = 0;
    sm_col_args(pl, sweep, 0, &ca, &n_inst);
    ca.re = re; ca.im = im; ca.cull_thr = nullptr; ca.write_im = 1;
    const bool lastsweep = (sweep == pl.col_passes - 1);
    ca.use_scale = lastsweep ? 1 : 0; ca.scale_ptr = nullptr; ca.scale_host = scale;
    std::vector<cf> cs(2 * (size_t)ca.L * SM_COL_TILE);
    int ntiles = (pl.Ch + 1 + SM_COL_TILE - 1) / SM_COL_TILE;
    for (int inst = 0; inst < n_inst; ++inst)
      for (int tile = 0; tile < ntiles; ++tile) {
        HostExec ex{sweep == 0 ? pl.thrA : pl.thrB};
        col_body(ex, pl, tile, inst, ca, twR.data(), cs.data());
      }
  }
  if (pl.col_passes == 0 && scale != 1.0f) {
    for (int k = 0; k <= pl.Ch; ++k) { re[k] *= scale; im[k] *= scale; }
  }
  return 0;
}

// inverse: column sweeps (cull on load) then rows + epilogue
int emu_inverse(int R, int C, float* re, float* im, float cull_thr, int out_mode, const uint16_t* base,
                uint16_t* out_bf16, float* out_f32, float scale, unsigned int* flags4) {
  SmPlan pl;
  int rc = sm_make_plan(R, C, &pl);
  if (rc) return rc;
  std::vector<cf> twC, twR, twQ;
  make_tw(C, twC); make_tw(R, twR); make_quads(pl, twQ);
  for (int i = 0; i < pl.col_passes; ++i) {
    int sweep = pl.col_passes - 1 - i;       // B first, then A
    ColArgs ca{};
    int n_inst = 0;
    sm_col_args(pl, sweep, 1, &ca, &n_inst);
    ca.re = re; ca.im = im; ca.write_im = 1; ca.use_scale = 0; ca.scale_ptr = nullptr; ca.scale_host = 1.f;
    ca.cull_thr = (i == 0) ? &cull_thr : nullptr;
    std::vector<cf> cs(2 * (size_t)ca.L * SM_COL_TILE);
    int ntiles = (pl.Ch + 1 + SM_COL_TILE - 1) / SM_COL_TILE;
    for (int inst = 0; inst < n_inst; ++inst)
      for (int tile = 0; tile < ntiles; ++tile) {
        HostExec ex{sweep == 0 ? pl.thrA : pl.thrB};
        col_body(ex, pl, tile, inst, ca, twR.data(), cs.data());
      }
  }
  std::vector<cf> smem(2 * (size_t)(pl.Ch + (pl.Ch >> 4) + 1) + 16);
  RowInvArgs a{};
  a.re = re; a.im = im; a.cull_thr = (pl.col_passes == 0) ? &cull_thr : nullptr;
  a.out_mode = out_mode; a.base = base; a.out_bf16 = out_bf16; a.out_f32 = out_f32;
  a.inv_n = (float)(1.0 / ((double)R * (double)C));
  a.scale_ptr = nullptr; a.scale_host = scale; a.flags = flags4; a.check_ifft = 1;
  for (int i = 0; i < 4; ++i) flags4[i] = 0;
  for (int row = 0; row < R; ++row) {
    HostExec ex{pl.row_threads};
    row_inv_body(ex, pl, row, a, twC.data(), twQ.data(), smem.data());
  }
  return 0;
}

}  // extern "C"
