"""Shared metrics for the parity tests (TEST INFRASTRUCTURE).

The reference algorithm is discontinuous (sign masks, percentile thresholds, a cull that
zeroes bins strictly below an order statistic), so two correct implementations whose FFTs
differ in the last bit can disagree on a handful of spectrum bins ("flips").
`flip_accounted` measures the relative L2 error after removing the K largest bins of the
difference spectrum, and reports how much of the error energy those bins carried.
"""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a = a.astype(dt); b = b.astype(dt)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def flip_accounted(ours, ref, k=16):
    """-> (raw rel L2, residual rel L2 without the k largest difference bins, their energy share)."""
    ours = np.asarray(ours, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    d = ours - ref
    D = np.fft.fftn(d) if d.ndim > 1 else np.fft.fft(d)
    E = (np.abs(D) ** 2).ravel()
    tot = float(E.sum())
    refE = float(np.sum(ref ** 2)) * ref.size           # Parseval: sum |FFT(ref)|^2 without a second transform
    if tot == 0.0:
        return 0.0, 0.0, 0.0
    k = min(k, E.size)
    top = float(np.partition(E, E.size - k)[E.size - k:].sum())
    return float(np.sqrt(tot / refE)), float(np.sqrt(max(tot - top, 0.0) / refE)), top / tot


def flipped_bins(ours, ref, tol=1e-5, kmax=64):
    """How many bins of the difference spectrum have to be set aside before the residual relative L2 error is <= tol,
    and which (row, column) frequencies they are: those are the mask / threshold decisions that went the other way.
    -> dict(n=..., bins=[[row, col, share of the error energy], ...]); n = -1 if kmax bins do not suffice."""
    ours = np.asarray(ours, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    d = ours - ref
    D = np.fft.fftn(d) if d.ndim > 1 else np.fft.fft(d)
    E = (np.abs(D) ** 2)
    tot = float(E.sum())
    refE = float(np.sum(ref ** 2)) * ref.size           # Parseval: sum |FFT(ref)|^2
    if tot == 0.0 or np.sqrt(tot / refE) <= tol:
        return dict(n=0, bins=[])
    flat = E.ravel()
    k = min(kmax, flat.size)
    idx = np.argpartition(flat, flat.size - k)[flat.size - k:]
    idx = idx[np.argsort(-flat[idx])]
    rest = tot
    bins = []
    for n, i in enumerate(idx, 1):
        rest -= float(flat[i])
        pos = np.unravel_index(int(i), E.shape)
        bins.append([int(p) for p in pos] + [round(float(flat[i]) / tot, 6)])
        if np.sqrt(max(rest, 0.0) / refE) <= tol:
            return dict(n=n, bins=bins)
    return dict(n=-1, bins=bins)


def bf16_ulp_distance(a_bits, b_bits):
    """ULP distance between two arrays of bf16 bit patterns (uint16)."""
    def key(u):
        u = u.astype(np.int32)
        return np.where(u & 0x8000, 0x8000 - (u & 0x7FFF), u + 0x8000)
    return np.abs(key(np.asarray(a_bits)) - key(np.asarray(b_bits)))


# ---- tests/golden/elem_*.npz (oracle/make_golden_elem.py): the element-wise strategies' fixtures ----------------------------
def _bf16_to_f32(u):
    return (np.asarray(u, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


ELEM_CASES = ["elem_addition_3x_64x128", "elem_addition_2x_special_32x64", "elem_taskaddition_3x_64x128",
              "elem_taskaddition_4x_96x40", "elem_taskaddition_2x_special_32x64",
              # other storage dtypes, more than 8 models (torch.sum on the CPU cascades from 16 rows on)
              "elem_addition_f16_3x_64x128", "elem_addition_f32_2x_special_32x64", "elem_addition_10x_48x64",
              "elem_taskaddition_f16_3x_special_64x128", "elem_taskaddition_f32_4x_96x40", "elem_taskaddition_12x_48x64",
              "elem_taskaddition_20x_48x64", "elem_taskaddition_f32_35x_24x64"]


def elem_case(golden_dir, name):
    """-> (dtype tag, base, finetunes, expected) of a tests/golden/elem_*.npz fixture (bf16 as uint16 bit patterns)"""
    d = np.load(golden_dir / f"{name}.npz")
    dt = str(d["dtype"]) if "dtype" in d.files else "bf16"
    return dt, d["base"], [d[f"ft{k}"] for k in range(int(d["n"]))], d["out"]


def elem_same(dt, got, want):
    """bit-identical, NaN == NaN"""
    wide = _bf16_to_f32 if dt == "bf16" else (lambda a: np.asarray(a, dtype=np.float32))
    fa, fb = wide(got), wide(want)
    return (got == want) | ((fa != fa) & (fb != fb))
