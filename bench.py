#!/usr/bin/env python
"""bench.py -- seeding throughput of the B200 hot path (BASELINE.json metric) with roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm  (one process per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU seeding path

Workload (N=1): BASELINE.json configs[1] -- chr1-scale reference (248,956,422 bp synthetic stand-in, no genomes
exist offline), 10 M synthetic 150 bp SE reads, mode-3 P-RMI index, seeding only (the semantics of
test/Learned_seeding_big_read.cpp with steps=3: SMEMs + reseeding + third round, every hit materialised).
A "step" is one pass of the seeding path over the whole 10 M-read batch.

value : reads/s with reads and index resident in HBM (meme_seed_run: seeding kernel + offset scans + hit gather)
e2e   : reads/s through meme_seed_batch with HOST (pinned) buffers: H2D of the reads, kernels, D2H of SMEM/hit lists
The index and reads are larger than L2 (11 GB / 1.5 GB vs 126 MB), so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

CACHE = os.environ.get("MEME_BENCH_CACHE", "/tmp/meme_b200_bench")
CHR1_LEN = 248_956_422


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-len", type=int, default=int(os.environ.get("MEME_BENCH_GENOME", CHR1_LEN)))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("MEME_BENCH_READS", 10_000_000)))
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--bits", type=int, default=0, help="P-RMI leaf bits (0 = build_rmis_dna.sh rule)")
    ap.add_argument("--mode", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wide", action="store_true", help="skip the 64-bit-index sub-record")
    ap.add_argument("--no-chain", action="store_true", help="skip the GPU chaining sub-record")
    ap.add_argument("--no-align", action="store_true", help="skip the end-to-end `mem -7` and kernel-B records")
    ap.add_argument("--align-genome", type=int, default=int(os.environ.get("MEME_BENCH_ALIGN_GENOME", 20_000_000)))
    ap.add_argument("--align-reads", type=int, default=int(os.environ.get("MEME_BENCH_ALIGN_READS", 2_000_000)))
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def workload_key(a) -> str:
    s = f"g{a.genome_len}-b{a.bits}-v5"
    return hashlib.sha1(s.encode()).hexdigest()[:12]


def prepare_index(a, device: str) -> tuple[str, dict]:
    """Build (once per box) the chr1-scale index in the reference's file formats; return its prefix."""
    ge.load_pkg()
    from bwa_meme_b200 import synth, index_build, prmi_train
    d = os.path.join(CACHE, workload_key(a))
    prefix = os.path.join(d, "ref.fa")
    meta_p = os.path.join(d, "meta.json")
    if os.path.exists(meta_p):
        return prefix, json.load(open(meta_p))
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    if a.genome_len >= 20_000_000:
        g = synth.make_genome_large(a.genome_len)
    else:
        g = synth.make_genome(a.genome_len)
    g.tofile(os.path.join(d, "genome.u8"))
    log(f"genome {a.genome_len} bp generated in {time.time() - t0:.1f}s; building index on {device}")
    bits = a.bits or prmi_train.default_bits(2 * a.genome_len)
    st = index_build.build_index(prefix, g, bits=bits, device=device, write_keys_file=False, verbose=False)
    st["build_s"] = time.time() - t0
    json.dump(st, open(meta_p, "w"))
    log(f"index built in {st['build_s']:.1f}s: {st}")
    return prefix, st


def make_reads(a, d: str, seed: int, device: str):
    """10 M x 150 bp synthetic reads (uint8 codes [n, L]) as a pinned host tensor."""
    import torch
    from bwa_meme_b200 import synth
    g = np.fromfile(os.path.join(d, "genome.u8"), dtype=np.uint8)
    gt = torch.from_numpy(g).to(device)
    r = synth.make_reads_torch(gt, a.reads, a.read_len, seed=seed)
    host = torch.empty(r.shape, dtype=torch.uint8, pin_memory=(device != "cpu"))
    host.copy_(r)
    del r, gt
    if device != "cpu":
        torch.cuda.empty_cache()
    return host


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        out = self.p.communicate()[0]
        mhz, mx, reasons, power = [], [], set(), []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                mhz.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not mhz:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(mhz)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(mhz), "power_w_max": max(power)}


def host_isa() -> str:
    flags = open("/proc/cpuinfo").read()
    return "avx512bw" if " avx512bw" in flags else ("avx2" if " avx2" in flags else "")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def ref_seed_binary(mode: int):
    isa = host_isa()
    for cand in ([isa] if isa else []) + ["avx2"]:
        p = os.path.join(ROOT, "oracle", "_ref", f"ref_seed_bench_mode{mode}.{cand}")
        if os.path.exists(p) and (cand != "avx512bw" or isa == "avx512bw"):
            return p
    return None


def run_reference_cpu(a, prefix: str, sample: np.ndarray, warmup: int, steps: int) -> dict:
    """Time the reference's own seeding functions on host cores (oracle/_ref), else the oracle port."""
    n, L = sample.shape
    exe = ref_seed_binary(a.mode)
    cores = host_cores()
    if exe:
        rb = os.path.join(os.path.dirname(prefix), f"sample_{os.getpid()}.bin")
        with open(rb, "wb") as f:
            f.write(np.array([n, L], dtype=np.int64).tobytes())
            f.write(np.ascontiguousarray(sample).tobytes())
        t0 = time.time()
        p = subprocess.run([exe, prefix, rb, str(cores), str(warmup), str(steps), "3"], capture_output=True, text=True)
        os.remove(rb)
        if p.returncode == 0:
            passes = [l.split() for l in p.stdout.splitlines() if l.startswith("PASS")]
            ms = [float(x[2]) for x in passes][warmup:]
            tot = [(int(x[3]), int(x[4])) for x in passes]
            return {"kind": "reference", "cores": cores, "ms": ms, "n_reads": n, "totals": tot[-1],
                    "wall_s": time.time() - t0, "binary": os.path.basename(exe)}
        log("reference binary failed:", p.stderr[-400:])
    # oracle port, one core
    O = ge.load_oracle()
    oix = O.OracleIndex(prefix, a.mode)
    offs = np.arange(n + 1, dtype=np.int64) * L
    ms = []
    for _ in range(max(1, min(steps, 2))):
        t0 = time.time()
        r = oix.seed(sample.reshape(-1), offs)
        ms.append((time.time() - t0) * 1e3)
    oix.close()
    return {"kind": "port", "cores": 1, "ms": ms, "n_reads": n, "totals": (len(r["smems"]), len(r["hits"])), "binary": "oracle/seed_oracle.c"}


def measured_peak_hbm() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    device = "cuda:0" if torch.cuda.is_available() else "cpu"
    prefix, meta = prepare_index(a, device)
    d = os.path.dirname(prefix)
    cores = host_cores()
    n_sample = a.cpu_sample or min(a.reads, max(200_000, 60_000 * cores))
    keep, a.reads = a.reads, n_sample
    sample = make_reads(a, d, seed=7, device=device).numpy()
    a.reads = keep
    res = run_reference_cpu(a, prefix, sample, a.warmup, a.steps)
    ms = float(np.mean(res["ms"]))
    value = res["n_reads"] / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "seeding_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"chr1-scale synthetic ref ({a.genome_len} bp, sa_num {meta['sa_num']}), "
                               f"{a.read_len} bp SE reads, mode-{a.mode} P-RMI (2^{meta['bits']} leaves), seeding only",
                   "step": f"one pass over a bounded sample of {res['n_reads']} reads of the 10 M-read workload"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": f"{res['n_reads']} reads x {a.steps} passes, {res['binary']}"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    ge.build_lib()
    ge.load_pkg()
    from bwa_meme_b200 import capi, dist as mdist

    # ---- index (rank 0 builds once per box, everyone loads) ----
    if rank == 0:
        prefix, meta = prepare_index(a, device)
    if world > 1:
        dist.barrier()
    if rank != 0:
        prefix, meta = prepare_index(a, device)
    d = os.path.dirname(prefix)
    torch.cuda.empty_cache()
    t0 = time.time()
    ix = capi.MemeIndex(prefix, mode=a.mode, device=local)
    log(f"rank {rank}: index resident in HBM: {ix.device_bytes / 1e9:.2f} GB, loaded in {time.time() - t0:.1f}s")
    # a non-default torch stream: the library launches on it and the torch events below are recorded on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ix.set_stream(stream.cuda_stream)

    # ---- reads: every rank owns its shard of the job (weak scaling: 10 M reads per GPU) ----
    host = make_reads(a, d, seed=7 + rank, device=device)
    n, L = host.shape
    offs = torch.arange(n + 1, dtype=torch.int64) * L
    offs_pinned = torch.empty_like(offs, pin_memory=True)
    offs_pinned.copy_(offs)
    opt = capi.default_seed_opt()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- HBM-resident timing: `value` ----
    ix.seed_upload_ptr(host.data_ptr(), offs_pinned.data_ptr(), n)
    # work counters (searches, block probes, ...) are compiled out of the timed kernels: one untimed pass collects them
    os.environ["MEME_SEED_STATS"] = "1"
    ix.seed_run(opt)
    counters = ix.seed_counters()
    del os.environ["MEME_SEED_STATS"]
    for _ in range(a.warmup):
        ix.seed_run(opt)
        if world > 1:                           # the record gather is part of a step: warm it up too (NCCL channels)
            parts = mdist.device_results(ix, n, device)
            gathered, _ = mdist.gather_records(parts, 0, device)
            del parts, gathered
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ix.seed_counters()["launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seed_ms = []
    gather_bytes = 0
    ev0.record()
    for _ in range(a.steps):
        ix.seed_run(opt)
        seed_ms.append(ix.seed_last_ms())
        if world > 1:
            # the one exchange step of the path, INSIDE the timed region: every rank's record lists, device-resident,
            # travel over NCCL (NVLink) into one buffer on rank 0 -- exact sizes, no padding, no host bounce
            parts = mdist.device_results(ix, n, device)
            gathered, moved = mdist.gather_records(parts, 0, device)
            gather_bytes = moved
            del parts, gathered
    ev1.record()
    sync_all()
    dev_ms = ev0.elapsed_time(ev1)
    launches = ix.seed_counters()["launches"] - l0
    n_smems, n_hits = ix.seed_totals()

    # ---- end-to-end through the C ABI with host buffers: `e2e` ----
    def e2e_step():
        # one meme_seed_batch call on host buffers: sub-batches pipeline H2D / kernels / D2H inside the library
        return ix.seed_batch_ptr(host.data_ptr(), offs_pinned.data_ptr(), n, opt)

    e2e_step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        tot = e2e_step()
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
    clocks = sampler.stop()
    h2d = n * L + (n + 1) * 8
    d2h = tot["n_smems"] * 24 + tot["n_hits"] * 8 + 2 * (n + 1) * 8

    # ---- the 64-bit-index kernel instantiations (sa_num >= 2^32: every GRCh38-scale run) on the same workload ----
    wide = None
    if rank == 0 and not a.no_wide:
        os.environ["MEME_FORCE_WIDE"] = "1"
        wix = capi.MemeIndex(prefix, mode=a.mode, device=local)
        del os.environ["MEME_FORCE_WIDE"]
        wix.set_stream(stream.cuda_stream)
        wix.seed_upload_ptr(host.data_ptr(), offs_pinned.data_ptr(), n)
        for _ in range(2):
            wix.seed_run(opt)
        torch.cuda.synchronize()
        w0e, w1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wms = []
        w0e.record()
        for _ in range(3):
            wix.seed_run(opt)
            wms.append(wix.seed_last_ms()["seed"])
        w1e.record()
        torch.cuda.synchronize()
        wtot = wix.seed_totals()
        wide = {"value": n * 3 / (w0e.elapsed_time(w1e) / 1e3), "unit": "reads/s", "k_seed_ms": float(np.mean(wms)), "steps": 3,
                "index_hbm_gb": round(wix.device_bytes / 1e9, 2), "same_results_as_narrow": bool(wtot == (n_smems, n_hits)),
                "how": "MEME_FORCE_WIDE=1: 5-byte positions / inverse SA and the WIDE kernel instantiations on the same index and reads"}
        wix.close()

    # ---- chaining + chain filtering on the device-resident lists: seeds -> filtered chains ----
    chain = None
    if rank == 0 and world == 1 and not a.no_chain:
        copt = capi.default_chain_opt()
        sopt = capi.default_seed_opt(max_occ=copt.max_occ)
        l_pac = int(ix.sa_num // 2)
        ix.chain_set_contigs(np.array([(0, l_pac, 0)], dtype=capi.CONTIG_DTYPE), l_pac)     # the synthetic reference is one contig
        ix.seed_upload_ptr(host.data_ptr(), offs_pinned.data_ptr(), n)
        ix.seed_run(sopt)
        ix.chain_run(copt)
        torch.cuda.synchronize()
        cms = []
        c0e, c1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0e.record()
        for _ in range(3):
            ix.seed_run(sopt)
            ix.chain_run(copt)
            cms.append(ix.chain_last_ms())
        c1e.record()
        torch.cuda.synchronize()
        ctot = ix.chain_download(copy=False)
        ls, lh = ix.seed_totals()
        chain = {"k_chain_ms": float(np.mean(cms)), "chained_reads_per_s": n / (float(np.mean(cms)) / 1e3), "seed_plus_chain_reads_per_s": n * 3 / (c0e.elapsed_time(c1e) / 1e3),
                 "chains": ctot["n_chains"], "chain_seeds": ctot["n_seeds"],
                 "d2h_bytes_chains": ctot["n_chains"] * 40 + ctot["n_seeds"] * 16 + (n + 1) * 8 + n * 4,
                 "d2h_bytes_lists_it_replaces": ls * 24 + lh * 8 + 2 * (n + 1) * 8,
                 "what": "mem_chain_Learned + mem_chain_flt (max_occ 500 sampling) for every read of the batch, one thread per read, on the "
                         "SMEM / hit lists meme_seed_run left in HBM; only chains and their seeds would travel to the host"}

    # ---- totals per rank (counts only; the record gather itself ran inside the timed region above) ----
    summary = mdist.gather_summary(ix, n, world, device)
    summary["record_gather"] = {"inside_timed_region": world > 1, "bytes_into_rank0_per_step": int(gather_bytes) if rank == 0 else None,
                                "how": "all_gather of sizes + NCCL send/recv of device-resident lists into one buffer on rank 0"}

    # ---- max over ranks ----
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_reads = n * world
    value = total_reads * a.steps / (dev_ms / 1e3)
    e2e_value = total_reads * a.steps / (e2e_ms / 1e3)

    if rank == 0:
        kseed_ms = float(np.mean([m["seed"] for m in seed_ms]))
        gather_ms = float(np.mean([m["gather"] for m in seed_ms]))
        line = {
            "metric": "seeding_reads_per_s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"chr1-scale synthetic ref ({a.genome_len} bp, sa_num {meta['sa_num']}), "
                                   f"{n} x {L} bp SE reads per GPU, mode-{a.mode} P-RMI (2^{meta['bits']} leaves), "
                                   "seeding only (BASELINE.json configs[1])",
                       "l2": "inputs larger than L2 (index %.1f GB, reads %.2f GB); no flush" % (ix.device_bytes / 1e9, n * L / 1e9),
                       "index_hbm_gb": round(ix.device_bytes / 1e9, 2), "parallelism": f"reads sharded x{world}, index replicated"},
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / a.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "results": {"smems_per_read": n_smems / n, "hits_per_read": n_hits / n, "searches_per_read": counters["searches"] / n,
                        "sa_blocks_per_read": counters["block_probes"] / n, "kary_rounds_per_read": counters["kary_rounds"] / n,
                        "ext_rounds_per_read": counters["ext_rounds"] / n, "gather": summary},
            "kernel_ms": {"k_seed": kseed_ms, "k_gather": gather_ms, "offsets": float(np.mean([m["offsets"] for m in seed_ms])),
                          "k_seed_share_of_step": kseed_ms / (dev_ms / a.steps)},
        }
        # ---- roofline + CPU baseline (rank 0, N=1) ----
        peak, peak_src = measured_peak_hbm()
        bytes_per_read = None
        if world == 1 and not a.no_cpu_baseline:
            O = ge.load_oracle()
            cores = host_cores()
            # (a) algorithmic bytes per read + full-scale parity check on a sample, via the oracle
            ns = min(n, 20000)
            sample = host[:ns].numpy()
            soffs = np.arange(ns + 1, dtype=np.int64) * L
            t0 = time.time()
            oix = O.OracleIndex(prefix, a.mode)
            want = oix.seed(sample.reshape(-1), soffs)
            oix.close()
            got = ix.seed(sample.reshape(-1), soffs)
            same = all((got[k] == want[k]).all() for k in ("smem_off", "hit_off", "hits")) and \
                all((got["smems"][f] == want["smems"][f]).all() for f in ("start", "end", "hitcount", "hitbeg"))
            bytes_per_read = O.bytes_seed(want["counters"], ns * L, a.mode) / ns
            line["parity_sample"] = {"reads": ns, "identical_to_oracle": bool(same), "oracle_s": time.time() - t0,
                                     "oracle_counters_per_read": {k: v / ns for k, v in want["counters"].items()}}
            if chain is not None:        # chains of the same sample: C ABI vs the chaining oracle on the seeding oracle's lists
                l_pac = int(ix.sa_num // 2)
                ct = np.array([(0, l_pac, 0)], dtype=O.CONTIG_DTYPE)
                cw = O.chain_lines(O.chain_reads(want, np.full(ns, L, np.int32), ct, l_pac))
                cg = O.chain_lines(ix.chain(sample.reshape(-1), soffs, capi.default_seed_opt(max_occ=500), capi.default_chain_opt()))
                chain["parity_sample"] = {"reads": ns, "identical_to_oracle": bool(cw == cg)}
            # (b) the reference's own CPU seeding on all host cores, bounded sample
            n_cpu = a.cpu_sample or min(n, max(200_000, 60_000 * cores))
            res = run_reference_cpu(a, prefix, host[:n_cpu].numpy(), 1, 2)
            cms = float(np.mean(res["ms"]))
            line["cpu_baseline"] = {"value": res["n_reads"] / (cms / 1e3), "unit": "reads/s", "cores": res["cores"],
                                    "kind": res["kind"], "sample": f"{res['n_reads']} reads x 2 passes after 1 warm-up, {res['binary']}"}
        else:
            line["cpu_baseline"] = None
        if bytes_per_read is None:
            ab = os.path.join(ROOT, "profiles", "algorithmic_bytes.json")
            if os.path.exists(ab):
                bytes_per_read = float(json.load(open(ab))["bytes_seed_per_read"])
        # DRAM traffic of the seeding kernels: not measurable from inside this process (it needs ncu's replay), so it is
        # the per-read figure of the committed ncu capture of THIS command's kernels, scaled to the reads of one launch
        # sequence; null when no capture of the current kernels is committed (profiles/traffic.json names its source)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("seed_kernels_dram_bytes_per_read") and tj.get("kernel") == "k_flat" and tj.get("workload_reads") == n:
                    traffic = tj["seed_kernels_dram_bytes_per_read"] * n
                    traffic_src = tj.get("source")
            except Exception:
                traffic = None
        if bytes_per_read:
            achieved = bytes_per_read * n / (kseed_ms / 1e3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                "traffic": traffic, "traffic_source": traffic_src, "kernel": "seeding kernels k_flat<1|2|3> (T1 zig-zag, T2 reseed, T3 third round) + k_pack_reads",
                                "algorithmic_bytes_per_read": bytes_per_read,
                                "peak_source": peak_src,
                                "note": "algorithmic bytes = SURVEY §8d formula with counters from oracle/seed_oracle.c on a "
                                        "sample of this workload; accesses are 8-24 B random reads, so the sector-granular DRAM "
                                        "ceiling is well below the streaming peak"}
        else:
            line["roofline"] = None
        line["wide"] = wide
        line["chain"] = chain
        if world == 1 and not a.no_align:
            try:
                from bwa_meme_b200 import e2e as _e2e
                line["mate_rescue"] = _e2e.bench_matesw(ix, threads=host_cores(), oracle=None if a.no_cpu_baseline else ge.load_oracle())
            except Exception as e:
                line["mate_rescue"] = {"error": repr(e)[:400]}
        if world == 1 and not a.no_align:
            try:
                line["align_e2e"], line["bsw"] = bench_align_and_bsw(a, device, local)
            except Exception as e:      # the headline (seeding) line must survive a failure of the secondary records
                line["align_e2e"] = {"error": repr(e)[:400]}
        emit(line)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


def bench_align_and_bsw(a, device: int, local: int) -> tuple[dict, dict]:
    """The second half of BASELINE.json's metric and kernel B, measured next to the seeding numbers (rank 0, N=1):
    `align_e2e` = Mbases aligned / s of `bwa-meme mem -7` end to end -- unmodified reference on all host cores vs the
    patched aligner with GPU seeding + GPU seed extension -- on the same FASTQ, SE and PE, SAM compared;
    `bsw` = kernel B on the extension pairs the aligner really builds (dumped from a short run), with the reference's
    getScores8/16 on the same pairs and all host cores beside it."""
    import torch
    from bwa_meme_b200 import synth, index_build, e2e
    cores = host_cores()
    d = os.path.join(CACHE, f"align-g{a.align_genome}-v2")
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, "ref.fa")
    ref_exe, b200_exe, _ = e2e.binaries(3)
    if not (os.path.exists(ref_exe) and os.path.exists(b200_exe)):
        return {"unavailable": "oracle/_ref aligner binaries are not in this tree"}, {"unavailable": "no pair dump without the aligner"}
    t0 = time.time()
    if not os.path.exists(prefix + ".pos_packed"):
        g = synth.make_genome_large(a.align_genome, n_families=30, copies=100)
        g.tofile(os.path.join(d, "genome.u8"))
        index_build.build_index(prefix, g, device=device, write_keys_file=False)
    g = np.fromfile(os.path.join(d, "genome.u8"), dtype=np.uint8)
    gt = torch.from_numpy(g).to(device)
    L = a.read_len
    n_se, n_pe = a.align_reads, a.align_reads // 2
    se = os.path.join(d, f"se{n_se}.fq")
    if not os.path.exists(se):
        r = synth.make_reads_torch(gt, n_se, L, seed=3).cpu().numpy()
        synth.write_fastq(se, r.reshape(-1), np.arange(n_se + 1, dtype=np.int64) * L)
        synth.write_fastq(os.path.join(d, "dump.fq"), r[:200000].reshape(-1), np.arange(200001, dtype=np.int64) * L)
    pe = [os.path.join(d, f"pe{n_pe}_1.fq"), os.path.join(d, f"pe{n_pe}_2.fq")]
    if not os.path.exists(pe[1]):
        r1, r2 = synth.make_pairs_torch(gt, n_pe, L, seed=5)
        offs = np.arange(n_pe + 1, dtype=np.int64) * L
        synth.write_fastq(pe[0], r1.cpu().numpy().reshape(-1), offs, prefix="p")
        synth.write_fastq(pe[1], r2.cpu().numpy().reshape(-1), offs, prefix="p")
    del gt
    torch.cuda.empty_cache()
    log(f"align_e2e inputs ready in {time.time() - t0:.1f}s")
    os.environ.setdefault("MEME_B200_DEVICE", str(local))
    align = {"workload": f"{a.align_genome} bp synthetic genome, mode-3 index, {L} bp reads, -K 100000000, {cores} host threads; "
                         "time = the aligner's own 'Processed N reads in ... real sec' (index load excluded on both sides)",
             "unit": "Mbases/s"}
    align["se"] = e2e.compare_arms(prefix, [se], n_se, L, cores, d, tag="se")
    align["pe"] = e2e.compare_arms(prefix, pe, 2 * n_pe, L, cores, d, tag="pe")
    # ---- kernel B on real pairs ----
    dump = os.path.join(d, "pairs")
    e2e.run_mem(b200_exe, prefix, [os.path.join(d, "dump.fq")], os.path.join(d, "dump.sam"), cores,
                {"MEME_B200": "1", "MEME_B200_BSW": "1", "MEME_B200_DUMP_PAIRS": dump, "MEME_B200_WARM_READS": "0"})
    bsw = {"pairs_from": "left/right extension pairs of 200000 reads, dumped by the patched aligner (real h0, real windows)",
           "band": "w=100 then 200 for pairs that are not final (MAX_BAND_TRY), decided on the device"}
    tot_cells = tot_ms = tot_pairs = 0
    cpu_ms = 0.0
    for side in ("left", "right"):
        r = e2e.bench_bsw_dump(f"{dump}.{side}", device=local)
        c = e2e.bench_bsw_cpu(f"{dump}.{side}", cores)
        bsw[side] = {"gpu": r, "cpu_getScores": c}
        tot_cells += r["cells"]; tot_ms += r["ms"]; tot_pairs += r["pairs"]
        if c and "ms_mean" in c:
            cpu_ms += c["ms_mean"]
    bsw["gcups"] = tot_cells / (tot_ms / 1e3) / 1e9
    bsw["pairs_per_s"] = tot_pairs / (tot_ms / 1e3)
    if cpu_ms:
        bsw["cpu_baseline"] = {"pairs_per_s": tot_pairs / (cpu_ms / 1e3), "gcups_same_cells": tot_cells / (cpu_ms / 1e3) / 1e9, "cores": cores,
                               "kind": "reference", "what": "BandedPairWiseSW::getScores8/16 (oracle/_ref/bsw_shim bench), pairs classed and "
                               "sorted as the extension driver does, blocks of 4096 pairs per host thread"}
    return align, bsw


def emit(line: dict):
    """The ONE JSON line of the contract, on the real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    args = parse()
    # Libraries print to stdout too (NCCL announces its version there): keep fd 1 for the JSON line alone and send
    # everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        main_reference(args)
    else:
        main_ours(args)
