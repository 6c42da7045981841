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
