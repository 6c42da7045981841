"""Index pipeline in the reference's on-disk formats.

Produces, for a genome given as uint8 codes, every file `bwa-meme mem -7` (and this
repo's `meme_index_load`) opens:

  <p>.pac .ann .amb          bntseq files        (reference: src/bntseq.cpp:L82-113, L313-372)
  <p>.0123                   fwd+RC, 1 B/base    (reference: src/Learnedindex.cpp:L150-223)
  <p>.pos_packed             5-byte SA entries   (reference: src/Learnedindex.cpp:L457-544)
  <p>.suffixarray_uint64     u64 n + n keys      (trainer input, same place)
  <p>.suffixarray_uint64_L{0,1,2}_PARAMETERS     (reference trainer is Rust and cannot be built
                                                  offline; see prmi_train.py)

On a GPU the suffix array and the keys come from the library's own builder (meme_sa_build,
csrc/meme_sa.cu: prefix doubling with device radix sorts between small kernels); the same algorithm
on torch tensors (suffix_array / rmi_keys below) serves CPU-only machines and pins the native one in
tests/test_index_gpu.py.  Ordering is the reference's:
the text `fwd + RC` is padded with P = max(longest A run, longest T run)+1 `T`s, suffixes
compare with the usual end-of-string-is-smallest rule, and suffixes that start inside the
padding are dropped (src/Learnedindex.cpp:L190-229, L242, L469-473).
"""
from __future__ import annotations

import os
import numpy as np
import torch


def fwd_rc_text(genome: np.ndarray) -> np.ndarray:
    """fwd + reverse complement, codes 0..3 (the `.0123` content)."""
    g = np.ascontiguousarray(genome, dtype=np.uint8)
    return np.concatenate([g, (3 - g[::-1]).astype(np.uint8)])


def _longest_run(t: torch.Tensor, val: int) -> int:
    m = (t == val).to(torch.int64)
    if int(m.sum()) == 0:
        return 0
    # run length at each position via cumsum trick
    idx = torch.arange(t.numel(), device=t.device, dtype=torch.int64)
    last_zero = torch.cummax(torch.where(m == 0, idx, torch.full_like(idx, -1)), 0).values
    return int(((idx - last_zero) * m).max())


def pad_len_of(text: torch.Tensor) -> int:
    return max(_longest_run(text, 0), _longest_run(text, 3)) + 1


def suffix_array(text_codes, device: str | torch.device = "cpu", verbose: bool = False) -> torch.Tensor:
    """Suffix array (int64, len(text)) of `text_codes` in the reference's order."""
    dev = torch.device(device)
    t8 = torch.as_tensor(text_codes, dtype=torch.uint8).to(dev)
    n = t8.numel()
    P = pad_len_of(t8)
    N = n + P
    assert N < (1 << 31), "prefix-doubling composite key assumes < 2^31 suffixes"
    sym = torch.empty(N + 32, dtype=torch.int64, device=dev)
    sym[:n] = t8.to(torch.int64) + 1
    sym[n:N] = 4
    sym[N:] = 0  # end of string: smaller than every base
    K0 = 27  # 5^27 < 2^63
    key = torch.zeros(N, dtype=torch.int64, device=dev)
    for r in range(K0):
        key.mul_(5).add_(sym[r:r + N])
    del sym
    order = torch.argsort(key)
    skey = key[order]
    del key
    head = torch.ones(N, dtype=torch.bool, device=dev)
    head[1:] = skey[1:] != skey[:-1]
    del skey
    slots = torch.arange(N, dtype=torch.int64, device=dev)
    grp = torch.cummax(torch.where(head, slots, torch.zeros_like(slots)), 0).values  # bucket head slot
    rank = torch.empty(N, dtype=torch.int64, device=dev)
    rank[order] = grp
    sa = order  # sa[slot] = suffix start
    # active = slots whose bucket has > 1 member
    nxt_head = torch.ones(N, dtype=torch.bool, device=dev)
    nxt_head[:-1] = head[1:]
    active = torch.nonzero(~(head & nxt_head)).squeeze(1)
    del head, nxt_head, grp, slots
    h = K0
    rnd = 0
    while active.numel() > 0:
        rnd += 1
        if verbose:
            print(f"[sa] round {rnd}: h={h} active={active.numel()}", flush=True)
        suf = sa[active]
        nb = suf + h
        sec = torch.where(nb < N, rank[torch.clamp(nb, max=N - 1)] + 1, torch.zeros_like(nb))
        comp = rank[suf] * (N + 1) + sec
        del nb, sec
        o = torch.argsort(comp)
        comp = comp[o]
        suf = suf[o]
        del o
        sa[active] = suf
        hd = torch.ones(active.numel(), dtype=torch.bool, device=dev)
        hd[1:] = comp[1:] != comp[:-1]
        del comp
        newrank = torch.cummax(torch.where(hd, active, torch.zeros_like(active)), 0).values
        rank[suf] = newrank
        nh = torch.ones_like(hd)
        nh[:-1] = hd[1:]
        keep = ~(hd & nh)
        active = active[keep]
        del hd, nh, keep, newrank, suf
        h *= 2
    del rank
    return sa[sa < n]


def rmi_keys(text_codes, sa: torch.Tensor, pad: int) -> torch.Tensor:
    """`.suffixarray_uint64` keys: 32 bases from T[SA[i]..], first base in bits 63-62,
    wrapping cyclically over the padded text (src/Learnedindex.cpp:L481-500).  Returned as
    int64 bit patterns."""
    dev = sa.device
    t8 = torch.as_tensor(text_codes, dtype=torch.uint8).to(dev)
    n = t8.numel()
    padded = torch.cat([t8, torch.full((pad,), 3, dtype=torch.uint8, device=dev)]).to(torch.int64)
    N = n + pad
    key = torch.zeros(sa.numel(), dtype=torch.int64, device=dev)
    for r in range(32):
        key = (key << 2) | padded[(sa + r) % N]
    return key


def write_bntseq(prefix: str, genome: np.ndarray, name: str = "chrS") -> None:
    """.pac (forward strand only, MSB-first 2-bit), .ann, .amb  (src/bntseq.cpp:L82-113, L347-363)."""
    g = np.ascontiguousarray(genome, dtype=np.uint8)
    l_pac = g.shape[0]
    padded = np.zeros((l_pac + 3) // 4 * 4, dtype=np.uint8)
    padded[:l_pac] = g
    q = padded.reshape(-1, 4)
    pac = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
    with open(prefix + ".pac", "wb") as f:
        f.write(pac.astype(np.uint8).tobytes())
        if l_pac % 4 == 0:
            f.write(b"\x00")
        f.write(bytes([l_pac % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write(f"{l_pac} 1 11\n")
        f.write(f"0 {name} (null)\n")
        f.write(f"0 {l_pac} 0\n")
    with open(prefix + ".amb", "w") as f:
        f.write(f"{l_pac} 1 0\n")


def write_pos_packed(path: str, sa: torch.Tensor) -> None:
    """5 bytes per entry: uint32 LE (pos>>8), uint8 (pos&0xff)."""
    s = sa.cpu().numpy().astype(np.uint64)
    out = np.empty((s.shape[0], 5), dtype=np.uint8)
    hi = (s >> 8).astype(np.uint32)
    out[:, :4] = hi.view(np.uint8).reshape(-1, 4)
    out[:, 4] = (s & 0xFF).astype(np.uint8)
    out.tofile(path)


def read_pos_packed(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8).reshape(-1, 5)
    hi = raw[:, :4].copy().view(np.uint32).reshape(-1).astype(np.uint64)
    return (hi << 8) | raw[:, 4].astype(np.uint64)


def build_index(prefix: str, genome: np.ndarray, bits: int | None = None, device: str = "cpu",
                name: str = "chrS", write_keys_file: bool = True, verbose: bool = False) -> dict:
    """Build every index file for `genome` under `prefix`. Returns a small summary dict."""
    from . import prmi_train  # local import: keeps this module importable on its own

    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    text = fwd_rc_text(genome)
    write_bntseq(prefix, genome, name)
    text.tofile(prefix + ".0123")
    dev = torch.device(device)
    if dev.type == "cuda" and not os.environ.get("MEME_INDEX_TORCH"):
        # the library's own builder (meme_sa_build: prefix doubling with device radix sorts, csrc/meme_sa.cu)
        from . import capi
        r = capi.sa_build(text, device=dev.index or 0)
        sa = torch.from_numpy(r["sa"])
        pad = r["pad"]
        keys = torch.from_numpy(r["keys"].view(np.int64)).to(dev)
    else:
        sa = suffix_array(text, device=device, verbose=verbose)
        t8 = torch.as_tensor(text, dtype=torch.uint8).to(sa.device)
        pad = pad_len_of(t8)
        keys = rmi_keys(text, sa, pad)
    write_pos_packed(prefix + ".pos_packed", sa)
    n = int(sa.numel())
    if write_keys_file:
        with open(prefix + ".suffixarray_uint64", "wb") as f:
            f.write(np.uint64(n).tobytes())
            f.write(keys.cpu().numpy().tobytes())
    if bits is None:
        bits = prmi_train.default_bits(n)
    if dev.type == "cuda" and not os.environ.get("MEME_INDEX_TORCH"):
        from . import capi                      # the library's own trainer (meme_prmi_train, csrc/meme_prmi.cu)
        l1, l2, stats = capi.prmi_train(r["keys"], bits, device=dev.index or 0)
        base = prefix + ".suffixarray_uint64"
        np.array([bits], dtype=np.uint64).tofile(base + "_L0_PARAMETERS")
        l1.tofile(base + "_L1_PARAMETERS")
        l2.tofile(base + "_L2_PARAMETERS")
    else:
        stats = prmi_train.train_and_write(prefix + ".suffixarray_uint64", keys, bits, verbose=verbose)
    stats.update({"sa_num": n, "l_pac": int(genome.shape[0]), "pad": pad, "bits": bits})
    return stats
