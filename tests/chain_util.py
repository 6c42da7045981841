"""Helpers shared by the chaining tests: the repeat-rich fixture genome, hit sampling as the device lists have it, parsing
of the reference's chain dumps."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rep_meta() -> dict:
    return json.load(open(os.path.join(GOLDEN, "chains_rep.json")))


def rep_genome(synth) -> np.ndarray:
    """Must match tests/golden/make_golden.py: chains_rep_genome."""
    m = rep_meta()
    g = synth.make_genome(m["length"], seed=m["seed"], n_families=m["n_families"], copies=m["copies"], rep_len=m["rep_len"],
                          divergence=m["divergence"])
    rng = np.random.default_rng(m["seed"] + 1)
    unit = rng.integers(0, 4, size=m["tandem_unit"], dtype=np.uint8)
    arr = np.tile(unit, m["tandem_copies"])
    mut = rng.random(arr.size) < 0.01
    arr[mut] = (arr[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    g[m["tandem_at"]:m["tandem_at"] + arr.size] = arr
    return g


def rep_reads(synth, g):
    m = rep_meta()
    return synth.make_reads(g, m["n_reads"], m["read_len"], seed=m["reads_seed"], n_frac=0.02)


def rep_contigs(dtype) -> np.ndarray:
    return np.array([(off, ln, alt) for name, off, ln, alt in rep_meta()["contigs"]], dtype=dtype)


def sample_hits(res: dict, max_occ: int) -> dict:
    """A seeding result with FULL hit lists -> the compact form the device lists have when the seeding run samples with
    max_occ (src/bwamem.cpp:L1154-1155: k = 0, step, 2*step, ... < hitcount, at most max_occ)."""
    sm = res["smems"].copy()
    so, ho, hits = res["smem_off"], res["hit_off"], res["hits"]
    n = len(so) - 1
    new_hits, new_ho = [], np.zeros(n + 1, np.int64)
    for r in range(n):
        hb = 0
        for i in range(int(so[r]), int(so[r + 1])):
            c = int(sm["hitcount"][i])
            step = c // max_occ if c > max_occ else 1
            take = min((c + step - 1) // step, max_occ)
            src = int(ho[r]) + int(sm["hitbeg"][i])
            new_hits.append(hits[src:src + take * step:step][:take])
            sm["hitbeg"][i] = hb
            hb += take
        new_ho[r + 1] = new_ho[r] + hb
    h = np.concatenate(new_hits).astype(np.uint64) if new_hits else np.zeros(0, np.uint64)
    return {"smems": sm, "smem_off": so.copy(), "hit_off": new_ho, "hits": h}


def read_dump(name: str) -> list[str]:
    return open(os.path.join(GOLDEN, name)).read().splitlines()


# ---- SAM-stage alignment cases (tests/golden/cigar_cases.json, oracle/ref_cigar_shim.cpp) ----------------------------------------
CIGAR_CASE_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("read_off", "<i8"), ("l_query", "<i4"), ("qb", "<i4"), ("qe", "<i4"),
                             ("truesc", "<i4"), ("reg_w", "<i4"), ("pad", "<i4")])


def make_cigar_cases(g: np.ndarray, contigs, n: int, seed: int):
    """Regions as the extension stage hands them to mem_reg2aln: a read segment [qb, qe) against a text range [rb, re) on
    either strand, with substitutions, indels of 1-12 bases, ragged ends, and local scores that make the band-doubling loop
    stop at its first, second and third try.  Returns (codes, cases)."""
    rng = np.random.default_rng(seed)
    l_pac = len(g)
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    codes, cases, off = [], [], 0
    while len(cases) < n:
        name, c_off, c_len, _ = contigs[int(rng.integers(0, len(contigs)))]
        tl = int(rng.integers(25, 220))
        st = int(rng.integers(c_off + 5, c_off + c_len - tl - 5))
        seg = g[st:st + tl].copy()
        q = list(seg)
        mm = rng.random(len(q)) < rng.choice([0.0, 0.01, 0.05])
        q = [int((b + rng.integers(1, 4)) & 3) if m else int(b) for b, m in zip(q, mm)]
        for _ in range(int(rng.choice([0, 0, 1, 1, 2, 3]))):
            p = int(rng.integers(3, max(4, len(q) - 3)))
            ln = int(rng.choice([1, 1, 2, 3, 5, 8, 12]))
            if rng.random() < 0.5:
                del q[p:p + ln]
            else:
                q[p:p] = [int(x) for x in rng.integers(0, 4, size=ln)]
        if len(q) < 20 or len(q) > 300:
            continue
        if rng.random() < 0.1:
            q[int(rng.integers(0, len(q)))] = 4                                   # an N in the read
        lead, trail = int(rng.choice([0, 0, 3, 17])), int(rng.choice([0, 0, 5, 22]))
        read = [int(x) for x in rng.integers(0, 4, size=lead)] + q + [int(x) for x in rng.integers(0, 4, size=trail)]
        if len(read) > 300:
            continue
        qb, qe = lead, lead + len(q)
        a0 = st + int(rng.choice([0, 0, 0, 1, -2]))                               # ragged reference ends
        a1 = st + tl + int(rng.choice([0, 0, 0, -1, 3]))
        a0, a1 = max(a0, c_off), min(a1, c_off + c_len)
        if a1 - a0 < 10:
            continue
        rd = np.array(read, np.uint8)
        if rng.random() < 0.5:                                                   # reverse strand: text coordinates beyond l_pac
            rb, re = 2 * l_pac - a1, 2 * l_pac - a0
            rd = comp[rd[::-1]]
            qb, qe = len(read) - qe, len(read) - qb
        else:
            rb, re = a0, a1
        approx = (qe - qb) - 5 * int(mm.sum())
        truesc = int(rng.choice([approx, approx - 10, qe - qb, (qe - qb) // 2, 30]))
        cases.append((rb, re, off, len(read), qb, qe, max(truesc, 1), int(rng.choice([100, 100, 200, 13])), 0))
        codes.append(rd)
        off += len(read)
    return np.concatenate(codes), np.array(cases, dtype=CIGAR_CASE_DTYPE)


def cigar_tasks(cases: np.ndarray, infer_w, task_dtype) -> np.ndarray:
    t = np.zeros(len(cases), dtype=task_dtype)
    t["rb"], t["re"] = cases["rb"], cases["re"]
    t["q_off"] = cases["read_off"] + cases["qb"]
    t["qlen"] = cases["qe"] - cases["qb"]
    t["truesc"] = cases["truesc"]
    t["w"] = [infer_w(int(c["qe"] - c["qb"]), int(c["re"] - c["rb"]), int(c["truesc"]), int(c["reg_w"])) for c in cases]
    return t


def run_cigar_shim(binary: str, prefix: str, codes: np.ndarray, cases: np.ndarray, workdir: str) -> list[dict]:
    import subprocess
    path = os.path.join(workdir, "cases.bin")
    with open(path, "wb") as f:
        f.write(np.array([len(cases), len(codes)], np.int64).tobytes())
        f.write(np.ascontiguousarray(codes, np.uint8).tobytes())
        f.write(np.ascontiguousarray(cases).tobytes())
    p = subprocess.run([binary, prefix, path], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-800:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
