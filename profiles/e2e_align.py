"""End-to-end `mem -7` timing on the GPU box: unmodified reference vs the patched aligner with GPU seeding
(and optionally GPU seed extension).  Prints one JSON line per arm; SAM outputs are compared (minus @PG)."""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import synth, index_build
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = int(os.environ.get("G", "20000000")); N = int(os.environ.get("N", "1000000")); T = os.environ.get("T", str(len(os.sched_getaffinity(0))))
d = "/tmp/meme_e2e"; os.makedirs(d, exist_ok=True)
prefix = d + "/ref.fa"
if not os.path.exists(prefix + ".pos_packed"):
    g = synth.make_genome_large(G, n_families=30, copies=100)
    g.tofile(d + "/genome.u8")
    index_build.build_index(prefix, g, device="cuda", write_keys_file=False)
g = np.fromfile(d + "/genome.u8", dtype=np.uint8)
fq = d + "/reads.fq"
if not os.path.exists(fq):
    reads = synth.make_reads_torch(torch.from_numpy(g).cuda(), N, 150, seed=3).cpu().numpy()
    synth.write_fastq(fq, reads.reshape(-1), np.arange(N + 1, dtype=np.int64) * 150)
isa = "avx512bw" if " avx512bw" in open("/proc/cpuinfo").read() else "avx2"
MODE = os.environ.get("MODE", "3")          # index layout the binaries are compiled for (1: positions only, 3: keys + ISA)
ref = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme_mode{MODE}.{isa}")
b200 = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme-b200_mode{MODE}.{isa}")
lib = os.path.join(ROOT, "bwa-meme_b200", "libmeme_b200.so")
sams = {}
for name, exe, env in (("reference_cpu", ref, {}), ("gpu_seed", b200, {"MEME_B200": "1"}),
                       ("gpu_seed_bsw", b200, {"MEME_B200": "1", "MEME_B200_BSW": "1"})):
    e = dict(os.environ); e.update(env); e["MEME_B200_LIB"] = lib
    t0 = time.time()
    p = subprocess.run([exe, "mem", "-7", "-t", T, "-K", "100000000", prefix, fq, "-o", f"{d}/{name}.sam"], capture_output=True, text=True, env=e)
    wall = time.time() - t0
    assert p.returncode == 0, p.stderr[-1500:]
    proc = [l for l in p.stderr.splitlines() if "Processed" in l and "real sec" in l]
    real = sum(float(l.split("CPU sec,")[1].split("real sec")[0]) for l in proc)
    sams[name] = "".join(l for l in open(f"{d}/{name}.sam") if not l.startswith("@PG"))
    out = {"arm": name, "mode": int(MODE), "genome_bp": G, "threads": int(T), "reads": N, "mem_process_seqs_real_s": real, "wall_s": wall,
           "Mbases_per_s": N * 150 / real / 1e6, "sam_identical_to_reference": sams[name] == sams["reference_cpu"]}
    gpu_lines = [l for l in p.stderr.splitlines() if "[meme_b200] seeded" in l]
    if gpu_lines: out["gpu_seed_log"] = gpu_lines[-1]
    if os.environ.get("E2E_VERBOSE"):
        print("\n".join(l for l in p.stderr.splitlines() if "[meme" in l), file=sys.stderr)
    print(json.dumps(out), flush=True)
