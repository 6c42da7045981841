"""End-to-end `bwa-meme mem -7` runs for measurement (bench.py `align_e2e`, profiles/e2e_align.py): the unmodified
reference aligner and the patched aligner on the same index and FASTQ, timed by the aligner's own
"Processed N reads in ... real sec" lines (src/bwamem.cpp:L1968, index load excluded on both sides), SAM compared minus @PG.
The binaries come from oracle/_ref (built where the reference sources are mounted; they travel to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_isa() -> str:
    flags = open("/proc/cpuinfo").read()
    return "avx512bw" if " avx512bw" in flags else "avx2"


def binaries(mode: int = 3):
    isa = host_isa()
    ref = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme_mode{mode}.{isa}")
    b200 = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme-b200_mode{mode}.{isa}")
    shim = os.path.join(ROOT, "oracle", "_ref", f"bsw_shim_mode3.{isa}")
    return ref, b200, shim


def run_mem(exe: str, prefix: str, fastqs: list[str], out_sam: str, threads: int, env_extra: dict | None = None,
            chunk_bases: int = 100_000_000) -> dict:
    env = dict(os.environ)
    env.update(env_extra or {})
    env["MEME_B200_LIB"] = os.path.join(ROOT, "bwa-meme_b200", "libmeme_b200.so")
    t0 = time.time()
    p = subprocess.run([exe, "mem", "-7", "-t", str(threads), "-K", str(chunk_bases), prefix, *fastqs, "-o", out_sam],
                       capture_output=True, text=True, env=env)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{os.path.basename(exe)} failed: {p.stderr[-1500:]}")
    proc = [l for l in p.stderr.splitlines() if "Processed" in l and "real sec" in l]
    real = sum(float(l.split("CPU sec,")[1].split("real sec")[0]) for l in proc)
    logs = [l for l in p.stderr.splitlines() if l.startswith("[meme_b200]")]
    return {"real_s": real, "wall_s": wall, "chunks": len(proc), "log": logs, "stderr_tail": p.stderr[-6000:]}


def sam_body(path: str) -> bytes:
    with open(path, "rb") as f:
        return b"".join(l for l in f if not l.startswith(b"@PG"))


def compare_arms(prefix: str, fastqs: list[str], n_reads: int, read_len: int, threads: int, workdir: str, mode: int = 3,
                 gpu_env: dict | None = None, tag: str = "se") -> dict:
    """reference CPU vs patched aligner with GPU seeding + chaining + seed extension on the same FASTQ."""
    ref, b200, _ = binaries(mode)
    gpu_env = gpu_env or {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1"}
    out = {}
    r = run_mem(ref, prefix, fastqs, os.path.join(workdir, f"{tag}_ref.sam"), threads)
    g = run_mem(b200, prefix, fastqs, os.path.join(workdir, f"{tag}_gpu.sam"), threads, gpu_env)
    same = sam_body(os.path.join(workdir, f"{tag}_ref.sam")) == sam_body(os.path.join(workdir, f"{tag}_gpu.sam"))
    bases = n_reads * read_len
    out["reference_cpu"] = {"Mbases_per_s": bases / r["real_s"] / 1e6, "real_s": r["real_s"], "threads": threads}
    out["gpu"] = {"Mbases_per_s": bases / g["real_s"] / 1e6, "real_s": g["real_s"], "threads": threads, "env": gpu_env,
                  "log_tail": g["log"][-2:]}
    out["sam_identical"] = bool(same)
    out["reads"] = n_reads
    out["speedup"] = r["real_s"] / g["real_s"]
    return out


def load_pair_dump(path: str):
    from . import capi
    raw = np.fromfile(path, dtype=np.uint8)
    n, nr, nq = (int(x) for x in np.frombuffer(raw[:24].tobytes(), dtype=np.int64))
    par = np.frombuffer(raw[24:72].tobytes(), dtype=np.int32)
    o = 72
    pairs = np.frombuffer(raw[o:o + n * 56].tobytes(), dtype=capi.SEQPAIR_DTYPE).copy()
    o += n * 56
    ref = raw[o:o + nr].copy()
    o += nr
    qer = raw[o:o + nq].copy()
    return pairs, ref, qer, par


def bench_bsw_dump(path: str, reps: int = 5, device: int = 0) -> dict:
    """Kernel B on one side of a pair dump (device-resident inputs, band retries on the device): GCUPS, pairs/s."""
    import ctypes as C
    import torch
    from . import capi
    pairs, ref, qer, par = load_pair_dump(path)
    w, tries, prev_is_h0 = int(par[0]), int(par[1]), int(par[2])
    opt = capi.BswOpt(*[int(x) for x in par[3:11]])
    b = capi.MemeBsw(opt, device=device)
    b.upload(pairs, ref, qer)
    lib = capi.lib()
    for _ in range(2):
        capi._check(lib.meme_bsw_run_ext(b.h, w, tries, prev_is_h0, C.byref(opt)), "meme_bsw_run_ext")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        capi._check(lib.meme_bsw_run_ext(b.h, w, tries, prev_is_h0, C.byref(opt)), "meme_bsw_run_ext")
    cells = b.cells()                       # synchronises
    dt = (time.perf_counter() - t0) / reps
    b.close()
    return {"pairs": int(len(pairs)), "ms": dt * 1e3, "pairs_per_s": len(pairs) / dt, "gcups": cells / dt / 1e9, "cells": int(cells),
            "mean_query_len": float(pairs["len2"].mean()), "mean_ref_len": float(pairs["len1"].mean()), "w": w, "tries": tries,
            "bytes": int(pairs["len1"].sum() + pairs["len2"].sum() + 80 * len(pairs))}


def bench_bsw_cpu(path: str, threads: int, reps: int = 2) -> dict | None:
    """The reference's getScores8/16 on the same dump, all host threads (oracle/_ref/bsw_shim `bench`)."""
    import json
    _, _, shim = binaries(3)
    if not os.path.exists(shim):
        return None
    p = subprocess.run([shim, path, "-", "bench", str(threads), str(reps)], capture_output=True, text=True)
    if p.returncode != 0:
        return {"error": p.stderr[-300:]}
    return json.loads(p.stdout.strip().splitlines()[-1])
