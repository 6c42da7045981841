"""Kernel B on REAL extension pairs (dumped by the patched aligner with MEME_B200_DUMP_PAIRS, see profiles/e2e_align.py
DUMP=...): GCUPS / pairs per second of the GPU kernel with the device-side band retries, per side."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import capi
import torch


def load_dump(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n, nr, nq = np.frombuffer(raw[:24].tobytes(), dtype=np.int64)
    par = np.frombuffer(raw[24:72].tobytes(), dtype=np.int32)
    o = 72
    pairs = np.frombuffer(raw[o:o + n * 56].tobytes(), dtype=capi.SEQPAIR_DTYPE).copy(); o += n * 56
    ref = raw[o:o + nr].copy(); o += nr
    qer = raw[o:o + nq].copy()
    return pairs, ref, qer, par


def bench_side(path, reps=5):
    pairs, ref, qer, par = load_dump(path)
    w, tries, prev_is_h0 = int(par[0]), int(par[1]), int(par[2])
    opt = capi.BswOpt(*[int(x) for x in par[3:11]])
    b = capi.MemeBsw(opt)
    b.upload(pairs, ref, qer)
    lib = capi.lib()
    import ctypes as C
    for _ in range(2):
        capi._check(lib.meme_bsw_run_ext(b.h, w, tries, prev_is_h0, C.byref(opt)), "run_ext")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        capi._check(lib.meme_bsw_run_ext(b.h, w, tries, prev_is_h0, C.byref(opt)), "run_ext")
    cells = b.cells()
    dt = (time.perf_counter() - t0) / reps
    b.close()
    return {"file": os.path.basename(path), "pairs": int(len(pairs)), "ms": dt * 1e3, "pairs_per_s": len(pairs) / dt, "gcups": cells / dt / 1e9,
            "cells_per_pair": cells / max(1, len(pairs)), "mean_len2": float(pairs["len2"].mean()), "mean_len1": float(pairs["len1"].mean()),
            "w": w, "tries": tries}


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(json.dumps(bench_side(p)), flush=True)
