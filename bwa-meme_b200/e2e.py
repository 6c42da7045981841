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
    chunk_s = [float(l.split("CPU sec,")[1].split("real sec")[0]) for l in proc]
    return {"real_s": real, "wall_s": wall, "chunks": len(proc), "chunk_s": chunk_s, "log": logs, "stderr_tail": p.stderr[-6000:]}


def sam_body(path: str) -> bytes:
    with open(path, "rb") as f:
        return b"".join(l for l in f if not l.startswith(b"@PG"))


def compare_arms(prefix: str, fastqs: list[str], n_reads: int, read_len: int, threads: int, workdir: str, mode: int = 3,
                 gpu_env: dict | None = None, tag: str = "se") -> dict:
    """reference CPU vs patched aligner with every GPU stage on (seeding + chaining, seed extension, the global alignments of
    the SAM stage, mate-rescue Smith-Waterman) on the same FASTQ."""
    ref, b200, _ = binaries(mode)
    gpu_env = gpu_env or {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1", "MEME_B200_MATE": "1"}
    out = {}
    # The patched aligner's glue raises glibc's malloc trim / mmap thresholds (meme_glue.cpp: ensure_index); the unmodified reference
    # gets the same settings through glibc's environment tunables, and the faster of its two runs is the baseline.
    tuned = {"MALLOC_TRIM_THRESHOLD_": str(0x7fffffff), "MALLOC_MMAP_THRESHOLD_": str(32 << 20), "MALLOC_TOP_PAD_": str(64 << 20)}
    r0 = run_mem(ref, prefix, fastqs, os.path.join(workdir, f"{tag}_ref.sam"), threads)
    r1 = run_mem(ref, prefix, fastqs, os.path.join(workdir, f"{tag}_ref_tuned.sam"), threads, tuned)
    r = r1 if r1["real_s"] < r0["real_s"] else r0
    g = run_mem(b200, prefix, fastqs, os.path.join(workdir, f"{tag}_gpu.sam"), threads, gpu_env)
    same = sam_body(os.path.join(workdir, f"{tag}_ref.sam")) == sam_body(os.path.join(workdir, f"{tag}_gpu.sam")) == sam_body(os.path.join(workdir, f"{tag}_ref_tuned.sam"))
    bases = n_reads * read_len
    out["reference_cpu"] = {"Mbases_per_s": bases / r["real_s"] / 1e6, "real_s": r["real_s"], "chunk_s": r["chunk_s"], "threads": threads,
                            "malloc": "glibc defaults" if r is r0 else "glibc tunables as in the patched aligner's glue",
                            "real_s_glibc_defaults": r0["real_s"], "real_s_glibc_tuned": r1["real_s"]}
    out["gpu"] = {"Mbases_per_s": bases / g["real_s"] / 1e6, "real_s": g["real_s"], "chunk_s": g["chunk_s"], "threads": threads, "env": gpu_env,
                  "log_tail": g["log"][-6:]}
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


def make_rescue_pairs(n: int, read_len: int = 150, window: int = 626, seed: int = 9, a: int = 1):
    """n mate-rescue pairs of the shape mem_matesw_batch_pre builds for 150 bp reads with a 400 +- 50 insert (window = high - low +
    read length, src/bwamem_pair.cpp:L1108-1121): the mate is a copy of part of the window with 2 % substitutions (70 % of the
    pairs; a fifth of those with a 3-base deletion) or unrelated to it.  Vectorised; returns (SEQPAIR array, ref bytes, query bytes)."""
    from . import capi
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, size=(n, window), dtype=np.uint8)
    st = rng.integers(0, window - read_len - 3, size=n)
    idx = st[:, None] + np.arange(read_len)[None, :]
    dele = rng.random(n) < 0.14
    idx = idx + (dele[:, None] & (np.arange(read_len)[None, :] >= read_len // 2)) * 3
    q = np.take_along_axis(ref, idx, axis=1)
    mm = rng.random((n, read_len)) < 0.02
    q = np.where(mm, (q + rng.integers(1, 4, size=(n, read_len), dtype=np.uint8)) & 3, q).astype(np.uint8)
    unrelated = rng.random(n) < 0.3
    q[unrelated] = rng.integers(0, 4, size=(int(unrelated.sum()), read_len), dtype=np.uint8)
    pairs = np.zeros(n, capi.SEQPAIR_DTYPE)
    pairs["idr"] = np.arange(n, dtype=np.int64) * window
    pairs["idq"] = np.arange(n, dtype=np.int64) * read_len
    pairs["len1"] = window; pairs["len2"] = read_len
    pairs["h0"] = capi.KSW_XSUBO | capi.KSW_XSTART | (capi.KSW_XBYTE if read_len * a < 250 else 0) | (19 * a)
    pairs["regid"] = np.arange(n)
    return pairs, ref.reshape(-1), q.reshape(-1)


def bench_matesw(ix, n: int = 400_000, threads: int = 16, reps: int = 3, oracle=None) -> dict:
    """Row f2: meme_matesw_batch on n rescue pairs (host buffers in, kswr_t out) with the reference's own batch step (sort_classify +
    mem_sam_pe_batch = kswv, oracle/_ref/matesw_shim `bench`, AVX512 hosts only) on the same pairs and all host threads beside it."""
    import json
    import tempfile
    pairs, ref, qer = make_rescue_pairs(n)
    segs = [(pairs, ref, qer)]
    ix.matesw_batch(segs)
    wall, kms = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = ix.matesw_batch(segs)
        wall.append((time.perf_counter() - t0) * 1e3)
        kms.append(out["ms"])
    cells = int(out["cells"])
    rec = {"pairs": n, "window": int(pairs["len1"][0]), "read_len": int(pairs["len2"][0]), "kernel_ms": float(np.mean(kms)), "call_ms": float(np.mean(wall)),
           "cells": cells, "gcups_kernel": cells / (float(np.mean(kms)) / 1e3) / 1e9, "pairs_per_s_call": n / (float(np.mean(wall)) / 1e3),
           "h2d_bytes": int(len(ref) + len(qer) + 56 * n), "d2h_bytes": 28 * n,
           "with_start_positions": int((out["aln"][0]["tb"] >= 0).sum()), "with_second_hit": int((out["aln"][0]["score2"] > 0).sum()),
           "what": "kswv phases 0 + 1 (full local alignment of the mate against the rescue window, score2 / te2, start positions) for every pair, "
                   "one thread per pair; cells counts both phases over the query padded as the reference pads it"}
    if oracle is not None:
        ns = 2000
        want = oracle.matesw_batch(np.array([(int(p["idr"]), int(p["idq"]), int(p["len1"]), int(p["len2"]), int(p["h0"]), 0) for p in pairs[:ns]],
                                            dtype=oracle.MATESW_REC_DTYPE), ref, qer)
        rec["parity_sample"] = {"pairs": ns, "identical_to_oracle": bool(out["aln"][0][:ns].tolist() == want.tolist())}
    shim = os.path.join(ROOT, "oracle", "_ref", "matesw_shim_mode3.avx512bw")
    if host_isa() == "avx512bw" and os.path.exists(shim):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "pairs.bin")
            with open(path, "wb") as f:
                f.write(np.array([n, len(ref), len(qer)], np.int64).tobytes())
                f.write(np.array([6, 1, 6, 1, 1, 4, 0, 0], np.int32).tobytes())
                r = np.zeros(n, dtype=[("r_off", "<i8"), ("q_off", "<i8"), ("len1", "<i4"), ("len2", "<i4"), ("xtra", "<i4"), ("pad", "<i4")])
                r["r_off"] = pairs["idr"].astype(np.int64); r["q_off"] = pairs["idq"].astype(np.int64)
                r["r_off"] = np.arange(n, dtype=np.int64) * int(pairs["len1"][0]); r["q_off"] = np.arange(n, dtype=np.int64) * int(pairs["len2"][0])
                r["len1"] = pairs["len1"]; r["len2"] = pairs["len2"]; r["xtra"] = pairs["h0"]
                f.write(r.tobytes()); f.write(ref.tobytes()); f.write(qer.tobytes())
            p = subprocess.run([shim, path, "-", "bench", str(threads), "2"], capture_output=True, text=True)
            if p.returncode == 0:
                c = json.loads(p.stdout.strip().splitlines()[-1])
                rec["cpu_baseline"] = {"ms_mean": c["ms_mean"], "pairs_per_s": n / (c["ms_mean"] / 1e3), "cores": threads, "kind": "reference",
                                       "what": "sort_classify + mem_sam_pe_batch (kswv512_u8/16, both phases) in blocks of 512 pairs per host thread"}
                rec["speedup_call_vs_cpu"] = c["ms_mean"] / float(np.mean(wall))
            else:
                rec["cpu_baseline"] = {"error": p.stderr[-300:]}
    return rec
