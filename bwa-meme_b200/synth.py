"""Seeded synthetic genomes and reads (SURVEY.md §8d generator spec).

No genomes exist on disk and there is no network, so every config of
BASELINE.json is stood in for by a deterministic generator:
  * genome: uniform ACGT + planted repeat families (copies diverged by a few %),
  * reads: uniform start, random strand, substitutions, occasional indel, occasional N.
Only numpy is used; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np

_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)
_ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)


def make_genome(length: int, seed: int = 20261017, n_families: int = 5, copies: int = 40,
                rep_len: int = 1000, divergence: float = 0.02) -> np.ndarray:
    """Return uint8 codes (0..3) of a synthetic genome of `length` bases."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=length, dtype=np.uint8)
    if length < 4 * rep_len:
        n_families = 0
    for _ in range(n_families):
        unit = rng.integers(0, 4, size=rep_len, dtype=np.uint8)
        for _ in range(copies):
            pos = int(rng.integers(0, length - rep_len))
            cp = unit.copy()
            mut = rng.random(rep_len) < divergence
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            g[pos:pos + rep_len] = cp
    return g


def revcomp(codes: np.ndarray) -> np.ndarray:
    return _COMP[codes[::-1]]


def make_reads(genome: np.ndarray, n_reads: int, read_len: int = 150, seed: int = 7,
               sub_rate: float = 0.01, indel_frac: float = 0.05, n_frac: float = 0.005,
               len_jitter: int = 0):
    """Return (codes, offsets): concatenated read codes (0..3, 4=N) and int64 offsets[n+1].

    len_jitter>0 draws each read length uniformly in [read_len-len_jitter, read_len].
    """
    rng = np.random.default_rng(seed)
    L = genome.shape[0]
    if len_jitter > 0:
        lens = rng.integers(read_len - len_jitter, read_len + 1, size=n_reads)
    else:
        lens = np.full(n_reads, read_len, dtype=np.int64)
    starts = rng.integers(0, L - read_len - 8, size=n_reads)
    strand = rng.random(n_reads) < 0.5
    has_indel = rng.random(n_reads) < indel_frac
    has_n = rng.random(n_reads) < n_frac
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    out = np.empty(int(offsets[-1]), dtype=np.uint8)
    # vectorised fast path for reads without indels
    for i in range(n_reads):
        ln = int(lens[i])
        s = int(starts[i])
        if has_indel[i] and ln > 30:
            k = int(rng.integers(1, 4))
            p = int(rng.integers(10, ln - 10))
            if rng.random() < 0.5:  # deletion from the read
                seg = np.concatenate([genome[s:s + p], genome[s + p + k:s + ln + k]])
            else:  # insertion into the read
                ins = rng.integers(0, 4, size=k, dtype=np.uint8)
                seg = np.concatenate([genome[s:s + p], ins, genome[s + p:s + ln - k]])
        else:
            seg = genome[s:s + ln].copy()
        out[offsets[i]:offsets[i + 1]] = seg
    # substitutions
    mut = rng.random(out.shape[0]) < sub_rate
    out[mut] = (out[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    # strand flip + N
    for i in np.nonzero(strand)[0]:
        a, b = offsets[i], offsets[i + 1]
        out[a:b] = _COMP[out[a:b][::-1]]
    for i in np.nonzero(has_n)[0]:
        a, b = offsets[i], offsets[i + 1]
        out[a + int(rng.integers(0, b - a))] = 4
    return out, offsets


def make_reads_fast(genome: np.ndarray, n_reads: int, read_len: int = 150, seed: int = 7,
                    sub_rate: float = 0.01, n_frac: float = 0.005) -> tuple[np.ndarray, np.ndarray]:
    """Fully vectorised variant for bench-scale read sets (no indels; fixed length)."""
    rng = np.random.default_rng(seed)
    L = genome.shape[0]
    starts = rng.integers(0, L - read_len, size=n_reads)
    idx = starts[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    mut = rng.random(reads.shape) < sub_rate
    reads[mut] = (reads[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    strand = rng.random(n_reads) < 0.5
    reads[strand] = _COMP[reads[strand][:, ::-1]]
    has_n = np.nonzero(rng.random(n_reads) < n_frac)[0]
    reads[has_n, rng.integers(0, read_len, size=has_n.shape[0])] = 4
    offsets = np.arange(n_reads + 1, dtype=np.int64) * read_len
    return reads.reshape(-1), offsets


def write_fasta(path: str, genome: np.ndarray, name: str = "chrS", width: int = 60) -> None:
    s = _ASCII[genome].tobytes()
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(s), width):
            f.write(s[i:i + width] + b"\n")


def write_fastq(path: str, codes: np.ndarray, offsets: np.ndarray, prefix: str = "r") -> None:
    s = _ASCII[codes].tobytes()
    n = len(offsets) - 1
    lens = np.diff(offsets)
    if n > 20000 and (lens == lens[0]).all():
        # fixed-length reads: one byte matrix per block of reads (names padded to a fixed width with a space-free scheme)
        L = int(lens[0])
        seq = np.frombuffer(s, dtype=np.uint8).reshape(n, L)
        width = len(str(n - 1))
        with open(path, "wb") as f:
            for a in range(0, n, 1 << 18):
                b = min(n, a + (1 << 18))
                names = np.array([("@%s%0*d\n" % (prefix, width, i)).encode() for i in range(a, b)], dtype=f"S{2 + len(prefix) + width}")
                nm = np.frombuffer(names.tobytes(), dtype=np.uint8).reshape(b - a, -1)
                rec = np.concatenate([nm, seq[a:b], np.frombuffer(b"\n+\n", dtype=np.uint8)[None, :].repeat(b - a, 0),
                                      np.full((b - a, L), ord("I"), np.uint8), np.full((b - a, 1), 10, np.uint8)], axis=1)
                f.write(rec.tobytes())
        return
    with open(path, "wb") as f:
        for i in range(len(offsets) - 1):
            a, b = int(offsets[i]), int(offsets[i + 1])
            f.write(b"@%s%d\n" % (prefix.encode(), i))
            f.write(s[a:b])
            f.write(b"\n+\n")
            f.write(b"I" * (b - a))
            f.write(b"\n")


def make_genome_large(length: int, seed: int = 20261017, n_families: int = 300, copies: int = 200,
                      rep_len: int = 1000, divergence: float = 0.08) -> np.ndarray:
    """chr1-scale stand-in: uniform ACGT plus many diverged repeat families
    (~n_families*copies*rep_len bases of repeats), so that seeding meets both unique and repetitive loci."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=length, dtype=np.uint8)
    for _ in range(n_families):
        unit = rng.integers(0, 4, size=rep_len, dtype=np.uint8)
        pos = rng.integers(0, length - rep_len, size=copies)
        div = rng.random() * divergence
        for p in pos:
            cp = unit.copy()
            mut = rng.random(rep_len) < div
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            g[p:p + rep_len] = cp
    return g


def make_reads_torch(genome_t, n_reads: int, read_len: int = 150, seed: int = 7, sub_rate: float = 0.01,
                     n_frac: float = 0.005, indel_frac: float = 0.05, chunk: int = 1 << 20):
    """Bench-scale read generator on a torch device (plumbing): uniform start, random strand, 1 % substitutions,
    5 % of reads with one 1-3 bp insertion or deletion, 0.5 % of reads with one N.
    Returns a uint8 tensor [n_reads, read_len] on genome_t's device."""
    import torch
    dev = genome_t.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    L = genome_t.numel()
    out = torch.empty((n_reads, read_len), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for a in range(0, n_reads, chunk):
        b = min(n_reads, a + chunk)
        m = b - a
        starts = torch.randint(0, L - read_len - 4, (m,), generator=gen, device=dev)
        # one indel per selected read: a deletion skips k source bases at p, an insertion puts k random bases there
        has_indel = torch.rand((m,), generator=gen, device=dev) < indel_frac
        is_del = torch.rand((m,), generator=gen, device=dev) < 0.5
        ipos = torch.randint(10, max(11, read_len - 10), (m,), generator=gen, device=dev)
        ilen = torch.randint(1, 4, (m,), generator=gen, device=dev)
        dele = (has_indel & is_del)[:, None]
        ins = (has_indel & ~is_del)[:, None]
        after = ar[None, :] >= ipos[:, None]
        inside = ins & after & (ar[None, :] < (ipos + ilen)[:, None])
        shift = torch.where(dele & after, ilen[:, None], torch.zeros_like(after, dtype=torch.int64))
        shift = shift - torch.where(ins & (ar[None, :] >= (ipos + ilen)[:, None]), ilen[:, None], torch.zeros_like(shift))
        r = genome_t[(starts[:, None] + ar[None, :] + shift)]
        rnd = torch.randint(0, 4, (m, read_len), generator=gen, device=dev, dtype=torch.uint8)
        r = torch.where(inside, rnd, r)
        mut = torch.rand((m, read_len), generator=gen, device=dev) < sub_rate
        delta = torch.randint(1, 4, (m, read_len), generator=gen, device=dev, dtype=torch.uint8)
        r = torch.where(mut, (r + delta) & 3, r)
        flip = torch.rand((m,), generator=gen, device=dev) < 0.5
        rc = (3 - r).flip(1)
        r = torch.where(flip[:, None], rc, r)
        has_n = torch.rand((m,), generator=gen, device=dev) < n_frac
        npos = torch.randint(0, read_len, (m,), generator=gen, device=dev)
        nmask = has_n[:, None] & (ar[None, :] == npos[:, None])
        r = torch.where(nmask, torch.full_like(r, 4), r)
        out[a:b] = r
    return out


def make_pairs_torch(genome_t, n_pairs: int, read_len: int = 150, seed: int = 11, sub_rate: float = 0.01,
                     ins_mean: float = 400.0, ins_std: float = 50.0, chunk: int = 1 << 20):
    """Paired-end reads on a torch device (plumbing): fragments of length N(400, 50^2), FR orientation, the fragment's
    strand random, 1 % substitutions.  Returns two uint8 tensors [n_pairs, read_len]."""
    import torch
    dev = genome_t.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    L = genome_t.numel()
    r1 = torch.empty((n_pairs, read_len), dtype=torch.uint8, device=dev)
    r2 = torch.empty((n_pairs, read_len), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for a in range(0, n_pairs, chunk):
        b = min(n_pairs, a + chunk)
        m = b - a
        ins = (torch.randn((m,), generator=gen, device=dev) * ins_std + ins_mean).round().clamp(read_len + 20, 1000).to(torch.int64)
        starts = (torch.rand((m,), generator=gen, device=dev, dtype=torch.float64) * (L - 1001)).to(torch.int64)
        left = genome_t[starts[:, None] + ar[None, :]]                          # fragment start, forward
        right = genome_t[(starts + ins)[:, None] - 1 - ar[None, :]]             # fragment end, read backwards ...
        right = 3 - right                                                       # ... and complemented
        for r in (left, right):
            mut = torch.rand((m, read_len), generator=gen, device=dev) < sub_rate
            delta = torch.randint(1, 4, (m, read_len), generator=gen, device=dev, dtype=torch.uint8)
            r[:] = torch.where(mut, (r + delta) & 3, r)
        flip = torch.rand((m,), generator=gen, device=dev) < 0.5                # the fragment comes from the other strand
        r1[a:b] = torch.where(flip[:, None], right, left)
        r2[a:b] = torch.where(flip[:, None], left, right)
    return r1, r2


def make_genome_textend(length: int = 30000, seed: int = 301, start_a_run: int = 6, end_t_run: int = 5) -> np.ndarray:
    """Fixture genome whose fwd+RC text has homopolymer ends: the forward strand starts with `start_a_run` A's (so the
    text ENDS with that many T's) and ends with `end_t_run` T's (the forward/reverse junction sits in a T|A run)."""
    g = make_genome(length, seed=seed, n_families=3, copies=12, rep_len=400)
    g[:start_a_run] = 0
    g[start_a_run] = 1
    g[-end_t_run:] = 3
    g[-end_t_run - 1] = 2
    return g


def make_textend_reads(genome: np.ndarray, pad: int, start_a_run: int = 6, read_len: int = 150, seed: int = 1):
    """Reads that run into the END of the fwd+RC text (the last k text bases followed by a tail: random bases, or a run
    of T's shorter than / equal to / longer than the index's padding run `pad`, or T's to the read end), both strands;
    reads spanning the forward/reverse junction; and reads whose tail is a long T (or A) run elsewhere.
    Returns (codes, offsets)."""
    rng = np.random.default_rng(seed)
    text = np.concatenate([genome, (3 - genome[::-1]).astype(np.uint8)])
    n = len(text)
    reads = []

    def add(r):
        r = np.asarray(r, dtype=np.uint8)[:read_len]
        reads.append(r)
        reads.append(revcomp(r))

    for k in list(range(1, 40)) + [50, 64, 65, 96, 100, 128, 140, 149]:
        pre, m = text[n - k:], read_len - k
        for c in (0, 1, pad - start_a_run - 1, pad - start_a_run, pad - 1, pad, pad + 1, pad + 5, m):
            c = max(0, min(m, c))
            rest = m - c
            tail = [np.full(c, 3, np.uint8)]
            if rest > 0:
                tail.append(rng.integers(0, 3, 1).astype(np.uint8))          # the run ends with a non-T base
                tail.append(rng.integers(0, 4, rest - 1).astype(np.uint8))
            add(np.concatenate([pre] + tail))
    L = len(genome)
    for k in (20, 75, 130):
        add(text[L - k:L - k + read_len])
    for _ in range(40):
        st, k = int(rng.integers(100, L - 200)), int(rng.integers(20, 140))
        add(np.concatenate([genome[st:st + k], np.full(read_len - k, 3, np.uint8)]))
    add(np.full(read_len, 3, np.uint8))
    codes = np.concatenate(reads)
    return codes, np.arange(len(reads) + 1, dtype=np.int64) * read_len
