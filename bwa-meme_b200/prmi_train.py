"""P-RMI trainer, file-format compatible with the reference's Rust trainer.

The reference trains `pwl{BIT},linear,linear_spline` with `RMI/` (Rust; crates are not vendored,
so it cannot be built offline -- SURVEY.md §8c).  Alignment results do not depend on the
parameter values (the last-mile search is self-correcting, src/LearnedIndex_seeding.cpp:L2277-2357),
only lookup cost does, so this is a restatement of the published structure, not of its numerics:

  top layer   : leaf = key >> (64 - BIT)                      (RMI/rmi_lib/src/models/piecewiselinear.rs:L23-27)
  leaf        : pred = fma(beta, (double)key, alpha)           (src/LearnedIndex_seeding.cpp:L186-210)
  dense leaves: (> PARTIAL_MIN keys) get a partial third layer of ~n/20 models; the leaf's
                line then predicts the sub-model index          (RMI/rmi_lib/src/train/two_layer.rs:L373-443)
  error word  : leaf/sub-model  (below << 32) | above           (two_layer.rs:L1521; decoded at
                                                                 src/LearnedIndex_seeding.cpp:L2145-2146)
                partial pointer (1<<63) | start << 32 | count    (two_layer.rs:L1536)

Files: <base>_L0_PARAMETERS (one u64 = BIT), <base>_L1_PARAMETERS (sub-models, >= one 24-byte record),
<base>_L2_PARAMETERS (2^BIT x {f64 alpha, f64 beta, u64 err}).
"""
from __future__ import annotations

import numpy as np
import torch

PARTIAL_MIN = 1000  # leaves with more keys get a third layer (two_layer.rs:L373)
PARTIAL_DIV = 20
SLACK = 1           # added to both error bounds (the reference's window code underflows with a zero window at position 0, L2262-2264)
TWO32 = 4294967296.0


def default_bits(sa_num: int) -> int:
    """build_rmis_dna.sh:L68-79 picks BIT by `.suffixarray_uint64` file size (8 B per key)."""
    size = sa_num * 8 + 8
    if size > 8e9:
        return 28
    if size > 1e9:
        return 26
    return 24


def u64_to_f64(key: torch.Tensor) -> torch.Tensor:
    """Correctly rounded (double)uint64 for int64 bit patterns."""
    hi = (key >> 32) & 0xFFFFFFFF
    lo = key & 0xFFFFFFFF
    return hi.to(torch.float64) * TWO32 + lo.to(torch.float64)


def _segment_bounds(seg: torch.Tensor, nseg: int):
    """first/last index (inclusive) of each segment id in sorted `seg`; count."""
    cnt = torch.bincount(seg, minlength=nseg)
    end = torch.cumsum(cnt, 0)
    start = end - cnt
    return start, end, cnt


def _fit_lines(kd: torch.Tensor, n: int, first: torch.Tensor, last: torch.Tensor, nonempty: torch.Tensor,
               empty_rank: torch.Tensor):
    """One line per segment through its two OUTSIDE neighbours: (key[first-1], first-1) and
    (key[last+1], last+1), `first`/`last` being global ranks of the segment's first/last key.
    Every key that can fall into the segment lies between those neighbours, so the prediction
    stays inside [first-1, last+1] for unseen keys too (the reference's trainer also feeds one
    neighbour point per side into its least-squares fit, two_layer.rs:L52-83).  Empty segments
    are constants = rank of the next key (two_layer.rs:L1453-1463)."""
    f = torch.where(nonempty, first, torch.zeros_like(first))
    l = torch.where(nonempty, last, torch.zeros_like(last))
    has_prev = f > 0
    has_next = l < n - 1
    k_prev = torch.where(has_prev, kd[torch.clamp(f - 1, min=0)], torch.zeros_like(kd[f]))
    k_next = torch.where(has_next, kd[torch.clamp(l + 1, max=n - 1)], torch.full_like(kd[l], 18446744073709551616.0))
    t_prev = (f - 1).to(torch.float64)
    t_next = (l + 1).to(torch.float64)
    dk = k_next - k_prev
    ok = dk > 0
    beta = torch.where(ok, (t_next - t_prev) / torch.where(ok, dk, torch.ones_like(dk)), torch.zeros_like(dk))
    alpha = torch.where(ok, t_prev - beta * k_prev, (t_prev + t_next) * 0.5)
    alpha = torch.where(nonempty, alpha, empty_rank)
    beta = torch.where(nonempty, beta, torch.zeros_like(beta))
    return alpha, beta


def _errors(kd, rank, seg, nseg, alpha, beta, n):
    pred = torch.floor(torch.clamp(beta[seg] * kd + alpha[seg], 0.0, float(n - 1)))
    diff = (pred - rank).to(torch.int64)
    dev = kd.device
    below = torch.zeros(nseg, dtype=torch.int64, device=dev).scatter_reduce_(0, seg, diff, "amax", include_self=True)
    above = torch.zeros(nseg, dtype=torch.int64, device=dev).scatter_reduce_(0, seg, -diff, "amax", include_self=True)
    return below, above


def train(keys: torch.Tensor, bits: int, verbose: bool = False):
    """keys: sorted int64 bit patterns of the uint64 RMI keys.  Returns numpy (L1, L2) record arrays."""
    dev = keys.device
    n = keys.numel()
    nleaf = 1 << bits
    kd = u64_to_f64(keys)
    rank = torch.arange(n, dtype=torch.float64, device=dev)
    leaf = ((keys >> (64 - bits)) & (nleaf - 1)).to(torch.int64)
    start, end, cnt = _segment_bounds(leaf, nleaf)
    next_rank = torch.clamp(end, max=n - 1).to(torch.float64)
    alpha, beta = _fit_lines(kd, n, start, end - 1, cnt > 0, next_rank)
    below, above = _errors(kd, rank, leaf, nleaf, alpha, beta, n)
    err = ((below + SLACK) << 32) | (above + SLACK)

    # ---- partial third layer for dense leaves -------------------------------------------------
    dense = torch.nonzero(cnt > PARTIAL_MIN).squeeze(1)
    l1_alpha = torch.zeros(1, dtype=torch.float64, device=dev)
    l1_beta = torch.zeros(1, dtype=torch.float64, device=dev)
    l1_err = torch.zeros(1, dtype=torch.int64, device=dev)
    if dense.numel() > 0:
        nsub = cnt[dense] // PARTIAL_DIV
        sub_start = torch.cumsum(nsub, 0) - nsub
        total_sub = int(nsub.sum())
        assert total_sub < (1 << 31)
        # the leaf's line now predicts the sub-model index in [0, nsub) across the leaf's key span
        s, e = start[dense], end[dense] - 1
        k0, k1 = kd[s], kd[e]
        dk = k1 - k0
        b2 = torch.where(dk > 0, nsub.to(torch.float64) / torch.where(dk > 0, dk, torch.ones_like(dk)),
                         torch.zeros_like(dk))
        a2 = -b2 * k0
        alpha[dense] = a2
        beta[dense] = b2
        err[dense] = (torch.ones_like(nsub) << 63) | (sub_start << 32) | nsub
        # assign every key of a dense leaf to its sub-model exactly as the lookup will
        is_dense = torch.zeros(nleaf, dtype=torch.bool, device=dev)
        is_dense[dense] = True
        dense_slot = torch.zeros(nleaf, dtype=torch.int64, device=dev)
        dense_slot[dense] = torch.arange(dense.numel(), device=dev)
        kidx = torch.nonzero(is_dense[leaf]).squeeze(1)          # global ranks of dense-leaf keys
        kl = dense_slot[leaf[kidx]]
        p = beta[leaf[kidx]] * kd[kidx] + alpha[leaf[kidx]]
        sub = torch.clamp(torch.floor(torch.clamp(p, min=0.0)).to(torch.int64), max=(nsub[kl] - 1))
        gsub = sub_start[kl] + sub  # non-decreasing because keys are sorted within a leaf
        st2, en2, cn2 = _segment_bounds(gsub, total_sub)
        K = kidx.numel()
        sub_leaf = torch.repeat_interleave(torch.arange(dense.numel(), device=dev), nsub)
        first = kidx[torch.clamp(st2, max=K - 1)]
        last = kidx[torch.clamp(en2 - 1, min=0)]
        nxt_in_leaf = (en2 < K) & (kl[torch.clamp(en2, max=K - 1)] == sub_leaf)
        empty_rank = torch.where(nxt_in_leaf, kidx[torch.clamp(en2, max=K - 1)], end[dense][sub_leaf])
        empty_rank = torch.clamp(empty_rank, max=n - 1).to(torch.float64)
        l1_alpha, l1_beta = _fit_lines(kd, n, first, last, cn2 > 0, empty_rank)
        bl, ab = _errors(kd[kidx], rank[kidx], gsub, total_sub, l1_alpha, l1_beta, n)
        l1_err = ((bl + SLACK) << 32) | (ab + SLACK)
    stats = {
        "n_leaf": nleaf,
        "n_dense_leaf": int(dense.numel()),
        "n_sub": int(l1_alpha.numel()),
        "mean_window": float((below + above + 2 * SLACK).to(torch.float64)[cnt > 0].mean()) if n else 0.0,
        "max_window": int((below + above).max()),
    }
    if verbose:
        print("[prmi]", stats, flush=True)

    def pack(a, b, e):
        rec = np.empty(a.numel(), dtype=[("alpha", "<f8"), ("beta", "<f8"), ("err", "<u8")])
        rec["alpha"] = a.cpu().numpy()
        rec["beta"] = b.cpu().numpy()
        rec["err"] = e.cpu().numpy().view(np.uint64)
        return rec

    return pack(l1_alpha, l1_beta, l1_err), pack(alpha, beta, err), stats


def train_and_write(base: str, keys: torch.Tensor, bits: int, verbose: bool = False) -> dict:
    l1, l2, stats = train(keys, bits, verbose=verbose)
    np.array([bits], dtype=np.uint64).tofile(base + "_L0_PARAMETERS")
    l1.tofile(base + "_L1_PARAMETERS")
    l2.tofile(base + "_L2_PARAMETERS")
    return stats


def lookup_numpy(l1: np.ndarray, l2: np.ndarray, keys_u64: np.ndarray, sa_num: int):
    """Vectorised restatement of learned_index_lookup (src/LearnedIndex_seeding.cpp:L186-210),
    using mul+add instead of fma (used only by tests to sanity-check the error bounds)."""
    nleaf = l2.shape[0]
    shift = 64 - (int(nleaf).bit_length() - 1)
    m = (keys_u64 >> np.uint64(shift)).astype(np.int64)
    kd = keys_u64.astype(np.float64)
    pred = l2["beta"][m] * kd + l2["alpha"][m]
    err = l2["err"][m].copy()
    part = (err >> np.uint64(63)) != 0
    if part.any():
        st = ((err[part] >> np.uint64(32)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
        cn = (err[part] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        p = np.clip(np.where(pred[part] < 0, 0, pred[part]), None, cn - 1).astype(np.int64)
        m2 = st + p
        pred[part] = l1["beta"][m2] * kd[part] + l1["alpha"][m2]
        err[part] = l1["err"][m2]
    pos = np.clip(np.where(pred < 0, 0, pred), None, sa_num - 1).astype(np.int64)
    below = ((err >> np.uint64(32)) & np.uint64(0x3FFFFFFF)).astype(np.int64)
    above = (err & np.uint64(0x7FFFFFFF)).astype(np.int64)
    return pos, below, above
