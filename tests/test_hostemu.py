"""CPU emulation of the FFT sweep bodies (tests/hostemu, the very code the kernels run, driven by
a sequential loop over emulated threads) against numpy: index math, twiddles, butterflies of every
radix, four-step row permutation, tangle/untangle, bf16 delta load and bf16 epilogue."""
import ctypes
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle_np as O
from tests.parity_util import bf16_ulp_distance, rel_l2

LIB = Path(__file__).resolve().parent / "hostemu" / "libsm_hostemu.so"
c_fp = ctypes.POINTER(ctypes.c_float)
c_u16 = ctypes.POINTER(ctypes.c_uint16)


@pytest.fixture(scope="module")
def emu():
    if not LIB.exists():
        import subprocess
        subprocess.run(["make", "hostemu"], cwd=LIB.parent.parent.parent, check=True)
    return ctypes.CDLL(str(LIB))


def P(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def forward(emu, R, C, x=None, base=None, ft=None, scale=1.0):
    pitch = emu.emu_plan_pitch(R, C)
    assert pitch > 0
    re = np.zeros((R, pitch), np.float32); im = np.zeros((R, pitch), np.float32)
    ss = ctypes.c_double(0)
    rc = emu.emu_forward(R, C, 1 if x is not None else 0, P(base, c_u16), P(ft, c_u16), P(x, c_fp),
                         ctypes.c_float(1), ctypes.c_float(1), ctypes.c_float(scale), P(re, c_fp), P(im, c_fp), ctypes.byref(ss))
    assert rc == 0
    return re, im, ss.value


def natural(emu, a, R, C):
    idx = np.array([emu.emu_row_freq(R, C, i) for i in range(R)])
    out = np.empty_like(a); out[idx] = a
    return out


SHAPES = [(1, 2), (1, 4), (1, 16), (1, 2048), (2, 4), (4, 4), (8, 8), (16, 16), (32, 64), (3, 10), (7, 22), (256, 64),
          (512, 64), (1024, 32), (4096, 8), (5632, 8), (896, 8), (96, 40), (1, 5632), (143, 26), (300, 12), (1, 7168),
          # generic radix stages (prime factors above 13): 37, 167, 17 * 19, in rows and in one- and two-sweep columns
          (1, 74), (1, 296), (1, 18944), (37, 8), (74, 26), (167, 4), (2004, 4), (18944, 4), (128256, 2), (323, 646)]


@pytest.mark.parametrize("shape", SHAPES)
def test_forward_and_roundtrip(emu, shape):
    R, C = shape
    rng = np.random.default_rng(R * 131 + C)
    x = rng.standard_normal((R, C)).astype(np.float32)
    re, im, ss = forward(emu, R, C, x=x)
    Ch = C // 2
    ref = np.fft.rfft2(x.astype(np.float64)) if R > 1 else np.fft.rfft(x.astype(np.float64), axis=1)
    got = natural(emu, re[:, : Ch + 1] + 1j * im[:, : Ch + 1], R, C)
    assert rel_l2(got, ref) < 5e-7
    # fp32 partial sums (per thread on the device, per row in the emulation), widened to fp64
    # (the emulation keeps ONE fp32 partial per row where the device keeps one per thread: a long single row needs slack)
    assert abs(ss / float((x.astype(np.float64) ** 2).sum()) - 1) < (2e-6 if C <= 8192 else 1e-5)
    out = np.zeros((R, C), np.float32); fl = (ctypes.c_uint * 4)()
    rc = emu.emu_inverse(R, C, P(re, c_fp), P(im, c_fp), ctypes.c_float(0.0), 1, None, None, P(out, c_fp), ctypes.c_float(1.0), fl)
    assert rc == 0 and list(fl) == [0, 0, 0, 0]
    assert rel_l2(out, x) < 1e-6


def test_bf16_delta_and_epilogue(emu):
    R, C = 64, 128
    rng = np.random.default_rng(9)
    base = O.f32_to_bf16((0.02 * rng.standard_normal((R, C))).astype(np.float32))
    ft = O.f32_to_bf16((O.bf16_to_f32(base) + 0.002 * rng.standard_normal((R, C))).astype(np.float32))
    delta = (O.bf16_to_f32(ft) - O.bf16_to_f32(base)).astype(np.float32)
    re, im, ss = forward(emu, R, C, base=base, ft=ft)
    got = natural(emu, re[:, : C // 2 + 1] + 1j * im[:, : C // 2 + 1], R, C)
    assert rel_l2(got, np.fft.rfft2(delta.astype(np.float64))) < 5e-7
    # inverse with the bf16 epilogue: base + delta*scale, RNE (a non-dyadic scale: with a dyadic one
    # base + delta*scale lands on exact bf16 rounding ties for ~2 % of these grid-valued inputs)
    out = np.zeros((R, C), np.uint16); fl = (ctypes.c_uint * 4)()
    rc = emu.emu_inverse(R, C, P(re, c_fp), P(im, c_fp), ctypes.c_float(0.0), 0, P(base, c_u16), P(out, c_u16), None,
                         ctypes.c_float(1.2345), fl)
    assert rc == 0
    expect = O.f32_to_bf16((O.bf16_to_f32(base) + (delta * np.float32(1.2345)).astype(np.float32)).astype(np.float32))
    u = bf16_ulp_distance(out, expect)
    # results that cancel to ~0 have a tiny ulp: allow a handful beyond 1 ulp, bound them absolutely
    assert float((u <= 1).mean()) >= 0.9995 and float((u == 0).mean()) > 0.995
    assert np.abs(O.bf16_to_f32(out) - O.bf16_to_f32(expect)).max() < 2e-4


def test_cull_on_load(emu):
    """|re| < thr is read as zero by the first inverse sweep (functions.py:146), 2-D and 1-D plans."""
    for R, C in ((32, 64), (1, 256)):
        rng = np.random.default_rng(R + C)
        x = rng.standard_normal((R, C)).astype(np.float32)
        re, im, _ = forward(emu, R, C, x=x)
        Ch = C // 2
        thr = np.float32(np.quantile(np.abs(re[:, : Ch + 1]), 0.3))
        spec = natural(emu, np.where(np.abs(re[:, : Ch + 1]) < thr, 0, re[:, : Ch + 1]) + 1j * im[:, : Ch + 1], R, C)
        expect = np.fft.irfft2(spec, s=(R, C)) if R > 1 else np.fft.irfft(spec, n=C, axis=1)
        out = np.zeros((R, C), np.float32); fl = (ctypes.c_uint * 4)()
        assert emu.emu_inverse(R, C, P(re, c_fp), P(im, c_fp), ctypes.c_float(thr), 1, None, None, P(out, c_fp),
                               ctypes.c_float(1.0), fl) == 0
        assert rel_l2(out, expect) < 1e-6
