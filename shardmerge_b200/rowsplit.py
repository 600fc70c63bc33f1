"""Row-split spectral merge of ONE tensor across the GPUs of a box (BASELINE config 5: the 128256 x 8192 embedding /
lm_head shape with its rows spread over 8 B200s).

The reference has no counterpart (its model-level path passes embed / lm_head through, shard/merge/fast_fourier.py:104-130);
this is `merge_tensors_fft2_slerp` + the tail of `_merge_layer` (shard/tensor/functions.py:164-221,
shard/merge/fast_fourier.py:147-165, 209-243, 269-276) for two bf16 finetunes of one base whose rows live on different GPUs.
Unlike whole-tensor partitioning this is NOT collective-free: the column FFT couples all rows and every
statistic is global.  One process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch):

    rows   (local)   delta + row FFT of this rank's rows, sum of squares          -> all-reduce of 2 doubles
    A2A              half spectrum [R/G][P] -> column slabs [R][w]  (4 planes; each rank ends up with all rows of w columns)
    cols   (local)   column FFT of the slab, x 1/||delta||
    stats            cutoff order statistic: exact radix select over the key bits, three rounds of local weighted
                     histograms (Hermitian multiplicities) + all-reduce; SLERP sums: local fp64 sums + all-reduce
    blend  (local)   sm_blend on the slab; cull order statistic like the cutoff
    cols   (local)   inverse column FFT, cull on load
    A2A              slabs -> rows (2 planes)
    rows   (local)   inverse row FFT + x target_norm + base + NaN / Inf policy + bf16 RNE

The FFT / blend / epilogue kernels are the single-GPU ones (the slab is just a narrower spectrum plane); the order statistics
and masked sums are torch reductions on the device here -- this path exists for one or two tensors per model, and the
128256 x 8192 shape also fits ONE B200 (0.33 s through FourierMerge.merge_sources), which is what merge() uses.
Only the SLERP branch is implemented (what the BASELINE inputs take); other branches raise NotImplementedError.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import engine as E


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _abs_bits(plane: torch.Tensor) -> torch.Tensor:
    """|x| as its int32 bit pattern: the order of non-negative floats is the order of their bits (NaN above Inf)."""
    return plane.view(torch.int32) & 0x7FFFFFFF


def dist_select_kth(keys: Sequence[torch.Tensor], mult: torch.Tensor, rank: int, group=None) -> int:
    """Bit pattern of the 0-based `rank`-th smallest key of the multiset spread over all ranks of `group`.
    keys: int32 [rows][w] tensors (non-negative); mult: int64 [w] copies of every key in a column (0: padding column).
    Exact radix select: bits [30:20], [19:9], [8:0]; per round one weighted histogram per rank and one all-reduce."""
    prefix, prefix_mask = 0, 0
    k = int(rank)
    wts = mult.to(torch.float64)
    for shift, bits in ((20, 11), (9, 11), (0, 9)):
        nb = 1 << bits
        hist = torch.zeros(nb, dtype=torch.float64, device=mult.device)
        for key in keys:
            bucket = ((key >> shift) & (nb - 1)).to(torch.int64)
            w_ = wts.expand(key.shape)
            if prefix_mask:
                w_ = w_ * ((key & prefix_mask) == prefix)
            hist += torch.bincount(bucket.reshape(-1), weights=w_.reshape(-1), minlength=nb)   # counts < 2^53: exact in fp64
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(hist, group=group)
        cum = torch.cumsum(hist, 0)
        b = int(torch.searchsorted(cum, torch.tensor(float(k), dtype=torch.float64, device=cum.device), right=True).item())
        b = min(b, nb - 1)
        if b > 0:
            k -= int(cum[b - 1].item())
        prefix |= b << shift
        prefix_mask |= (nb - 1) << shift
    return prefix


def merge_rowsplit(base_rows: torch.Tensor, ft_rows: List[torch.Tensor], alphas: Sequence[float], group=None,
                   cutoff_pct: float = 0.08, cull_pct: float = 0.20, target_norm_offset: float = 1e-10,
                   info: Optional[dict] = None) -> torch.Tensor:
    """This rank's rows of the merged tensor (bf16 [R/G][C]).  `base_rows` / `ft_rows[k]`: this rank's contiguous block of
    rows (rank g holds rows g*R/G .. (g+1)*R/G - 1) of the base and of the two finetunes, bf16 on this rank's GPU."""
    if len(ft_rows) != 2:
        raise NotImplementedError("merge_rowsplit merges two finetunes (one pair merge)")
    G = dist.get_world_size(group) if dist.is_initialized() else 1
    g = dist.get_rank(group) if dist.is_initialized() else 0
    dev = E._require_cuda(base_rows.device)
    if base_rows.ndim != 2 or base_rows.dtype != torch.bfloat16 or any(t.dtype != torch.bfloat16 or t.shape != base_rows.shape for t in ft_rows):
        raise ValueError("merge_rowsplit needs bf16 [rows][C] blocks of one shape")
    Rl, C = base_rows.shape
    R, Ch = Rl * G, C // 2
    N = R * C
    base_rows = base_rows.contiguous()
    ws_row = E.get_workspace(Rl, C, dev, n_spectra=2, lane="rowsplit")
    P = ws_row.plan.P
    w = _round_up(-(-(Ch + 1) // G), 32)                      # columns per slab (the last slabs may be partly / all padding)
    ws_col = E.get_workspace(R, 2 * (w - 1), dev, n_spectra=2, lane="rowsplit")
    assert ws_col.plan.P == w

    # ---- rows: delta, row FFT, sums of squares (global) ---------------------------------------------------------------
    ws_row.ctl.zero_(); ws_col.ctl.zero_()
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    srcs = [E.Source(base=base_rows, ft=ft.contiguous(), weight=float(a)) for ft, a in zip(ft_rows, alphas)]
    for i, src in enumerate(srcs):
        E.fwd_rows_ptr(ws_row, i, src, sums.data_ptr() + 8 * i)
    if G > 1:
        dist.all_reduce(sums, group=group)
    nx, ny = (E.f32(v ** 0.5) for v in sums.tolist())
    # the host decisions of fast_fourier.py:165, 209-232 (what k_prepare does on the device for the one-GPU chain)
    swap = abs(nx) < abs(ny)
    na, nb = (ny, nx) if swap else (nx, ny)
    target_norm = torch.tensor([nx, ny], dtype=torch.float32).mean().item() + target_norm_offset
    cnorm_a, cnorm_b = abs(na / target_norm), abs(nb / target_norm)
    if cnorm_a < 1e-6 or cnorm_b < 1e-6 or cnorm_b / (cnorm_a + 1e-10) < 0.1 or na < 1e-4 or nb < 1e-4:
        raise NotImplementedError("merge_rowsplit implements the SLERP branch only (norms %.3e / %.3e)" % (nx, ny))
    a_w, b_w = float(alphas[0]), float(alphas[1])
    t = a_w / (a_w + b_w)                                      # config order, never swapped (fast_fourier.py:234)
    ia, ib = (1, 0) if swap else (0, 1)                        # slots of the models in roles v0 / v1

    # ---- rows -> column slabs, column FFT ------------------------------------------------------------------------------
    def to_slab(row_plane: torch.Tensor, slab_plane: torch.Tensor):
        send = torch.zeros((Rl, G * w), dtype=torch.float32, device=dev)
        send[:, :P] = row_plane
        send = send.view(Rl, G, w).permute(1, 0, 2).contiguous()               # [peer][my rows][w]
        recv = slab_plane.view(G, Rl, w)                                          # [peer rows][w] = rows in global order
        if G > 1:
            dist.all_to_all_single(recv, send, group=group)
        else:
            recv.copy_(send)

    for i in range(2):
        to_slab(ws_row.re[i], ws_col.re[i])
        to_slab(ws_row.im[i], ws_col.im[i])
    E.fwd_cols(ws_col, ia, scale=E.inv_norm_f32(na), write_im=True)
    E.fwd_cols(ws_col, ib, scale=E.inv_norm_f32(nb), write_im=False)
    re0, im0, re1 = ws_col.re[ia], ws_col.im[ia], ws_col.re[ib]

    # ---- statistics over the whole tensor ------------------------------------------------------------------------------
    col = g * w + torch.arange(w, device=dev, dtype=torch.int64)                # global column of every slab column
    mult = torch.where(col <= Ch, torch.where((col == 0) | (col == Ch), 1, 2), 0).to(torch.int64)
    total2 = 2 * N
    rank_cut = min(int(total2 * cutoff_pct), total2 - 1)                         # functions.py:113-120
    thr_cut = 0.0
    if cutoff_pct > 0:
        bits = dist_select_kth([_abs_bits(re0), _abs_bits(re1)], mult, rank_cut, group)
        thr_cut = torch.tensor([bits], dtype=torch.int32).view(torch.float32).item()
    ws_col.flt[E.F_THR_CUT] = thr_cut
    m = (torch.sign(re0) == torch.sign(re1)) & ~(re1.abs() < thr_cut)            # :124-127 (both "small" masks test re1)
    wm = m * mult.to(torch.float64)
    a64, b64 = re0.double(), re1.double()
    s = torch.stack([(a64 * a64 * wm).sum(), (b64 * b64 * wm).sum(), (a64 * b64 * wm).sum()])
    del a64, b64, wm, m
    if G > 1:
        dist.all_reduce(s, group=group)
    ws_col.dbl[E.D_S00:E.D_S00 + 3] = s
    E.slerp_scalars(ws_col, t)                                                   # dot, cos, sin, ||rel|| (functions.py:36-43)
    E.blend(ws_col, 0, True, re0, re1, 1.0, re0)                                 # :134-136, in place
    cull = cull_pct > 0
    if cull:
        rank_cull = min(int(N * cull_pct), N - 1)                                # :138-141
        bits = dist_select_kth([_abs_bits(re0)], mult, rank_cull, group)
        ws_col.flt[E.F_THR_CULL] = torch.tensor([bits], dtype=torch.int32).view(torch.float32).item()

    # ---- inverse: columns on the slab, back to rows, rows + epilogue ----------------------------------------------------
    E.inv_cols(ws_col, re0, im0, cull)

    def to_rows(slab_plane: torch.Tensor, row_plane: torch.Tensor):
        send = slab_plane.view(G, Rl, w)                                          # [peer's rows][w], contiguous already
        recv = torch.empty((G, Rl, w), dtype=torch.float32, device=dev)          # [peer = slab][my rows][w]
        if G > 1:
            dist.all_to_all_single(recv, send, group=group)
        else:
            recv.copy_(send)
        row_plane.copy_(recv.permute(1, 0, 2).reshape(Rl, G * w)[:, :P])

    to_rows(re0, ws_row.re[0])
    to_rows(im0, ws_row.im[0])
    out = torch.empty((Rl, C), dtype=torch.bfloat16, device=dev)
    # the row kernel divides by its own plan's Rl * C; the transform spans R = G * Rl rows
    E.inv_rows(ws_row, ws_row.re[0], ws_row.im[0], False, E.f32(target_norm) / G, base_rows, out, check_ifft=True)
    _, _, flags, _ = ws_row.read_ctl()
    bad = torch.tensor([int(flags[1]), int(flags[3])], dtype=torch.int64, device=dev)
    if G > 1:
        dist.all_reduce(bad, group=group)
    if int(bad[0]) > 0:
        raise ValueError("Inf in ifft output")                                   # functions.py:215-217
    if int(bad[1]) > 0:
        raise ValueError("Inf in merged tensor")                                 # fast_fourier.py:273-274
    if info is not None:
        info.update(norms=[nx, ny], target_norm=target_norm, swap=int(swap), thr_cut=thr_cut,
                    thr_cull=float(ws_col.flt[E.F_THR_CULL]) if cull else None, slab_columns=w, ranks=G)
    return out
