"""Mate-rescue SW test cases: windows of a genome against reads as mem_matesw_batch_pre pairs them."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000


def make_matesw_cases(n: int, seed: int, a: int = 1, long_reads: float = 0.15, rec_dtype=None):
    """n pairs: a reference window of 30..900 bases (mate-rescue windows are insert-size wide) and a read that is a mutated
    copy of part of it (substitutions, indels, N, clipped ends), of a repeat inside it (second-best hits), or unrelated.
    xtra as mem_matesw_batch_pre sets it (src/bwamem_pair.cpp:L1128): XSUBO | XSTART | (l_ms * a < 250 ? XBYTE : 0) | min_seed_len * a."""
    rng = np.random.default_rng(seed)
    ref, qer, recs = [], [], []
    ro = qo = 0
    for k in range(n):
        l1 = int(rng.choice([30, 64, 150, 300, 480, 626, 627, 900])) + int(rng.integers(0, 20))
        win = rng.integers(0, 4, size=l1).astype(np.uint8)
        if rng.random() < 0.3 and l1 > 200:                                     # a diverged repeat of part of the window
            ln = int(rng.integers(40, 120)); s0 = int(rng.integers(0, l1 - 2 * ln)); s1 = int(rng.integers(s0 + ln, l1 - ln))
            win[s1:s1 + ln] = win[s0:s0 + ln]
            mm = rng.random(ln) < 0.04
            win[s1:s1 + ln][mm] = (win[s1:s1 + ln][mm] + 1) & 3
        l2 = int(rng.integers(250, 301)) if rng.random() < long_reads else int(rng.choice([20, 36, 76, 100, 150, 151, 160, 249 // a]))
        kind = rng.random()
        if kind < 0.12 or l1 < l2 + 5:
            q = rng.integers(0, 4, size=l2).astype(np.uint8)                    # unrelated
        else:
            st = int(rng.integers(0, l1 - l2))
            q = list(win[st:st + l2])
            mm = rng.random(len(q)) < rng.choice([0.0, 0.01, 0.03, 0.08])
            q = [int((b + rng.integers(1, 4)) & 3) if m else int(b) for b, m in zip(q, mm)]
            for _ in range(int(rng.choice([0, 0, 0, 1, 1, 2]))):
                p = int(rng.integers(3, max(4, len(q) - 3))); ln = int(rng.choice([1, 1, 2, 3, 6, 10]))
                if rng.random() < 0.5:
                    del q[p:p + ln]
                else:
                    q[p:p] = [int(x) for x in rng.integers(0, 4, size=ln)]
            if rng.random() < 0.25:                                             # soft-clipped ends
                c = int(rng.integers(3, 30))
                q[:c] = [int(x) for x in rng.integers(0, 4, size=c)]
            if rng.random() < 0.25:
                c = int(rng.integers(3, 30))
                q[-c:] = [int(x) for x in rng.integers(0, 4, size=c)]
            q = np.array(q[:300] if len(q) >= 20 else q + [0] * 20, np.uint8)
        if rng.random() < 0.1:
            q[int(rng.integers(0, len(q)))] = 4
        if rng.random() < 0.03:
            win[int(rng.integers(0, l1))] = 4                                   # the aligner's text has no N, the C ABI allows it
        l2 = len(q)
        xtra = XSUBO | XSTART | (XBYTE if l2 * a < 250 else 0) | (19 * a)
        recs.append((ro, qo, l1, l2, xtra, 0))
        ref.append(win); qer.append(q)
        ro += l1; qo += l2
    dt = rec_dtype if rec_dtype is not None else np.dtype([("r_off", "<i8"), ("q_off", "<i8"), ("len1", "<i4"), ("len2", "<i4"), ("xtra", "<i4"), ("pad", "<i4")])
    return np.array(recs, dtype=dt), np.concatenate(ref), np.concatenate(qer)


def shim_path():
    p = os.path.join(ROOT, "oracle", "_ref", "matesw_shim_mode3.avx512bw")
    return p if os.path.exists(p) and " avx512bw" in open("/proc/cpuinfo").read() else None


def run_shim(O, recs, ref, qer, opt, tmpdir, blk=None):
    inp, outp = os.path.join(tmpdir, "msw_in.bin"), os.path.join(tmpdir, "msw_out.bin")
    O.matesw_write_cases(inp, recs, ref, qer, opt)
    env = dict(os.environ)
    if blk:
        env["MATESW_BLK"] = str(blk)
    subprocess.run([shim_path(), inp, outp], check=True, env=env)
    return np.fromfile(outp, dtype=O.KSWR_DTYPE)
