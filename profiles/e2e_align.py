"""End-to-end `mem -7` timing on the GPU box: unmodified reference vs the patched aligner with GPU seeding and GPU seed
extension (chunk-level driver).  Prints one JSON line per arm; SAM outputs are compared (minus @PG).
env: G genome bp, N reads (or pairs with PE=1), T threads, MODE index mode, PE=1 paired-end, DUMP=path (pair dump)."""
import json, os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import synth, index_build
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = int(os.environ.get("G", "20000000")); N = int(os.environ.get("N", "1000000")); T = os.environ.get("T", str(len(os.sched_getaffinity(0))))
PE = os.environ.get("PE", "0") == "1"
d = "/tmp/meme_e2e"; os.makedirs(d, exist_ok=True)
prefix = d + "/ref.fa"
if not os.path.exists(prefix + ".pos_packed"):
    g = synth.make_genome_large(G, n_families=30, copies=100)
    g.tofile(d + "/genome.u8")
    index_build.build_index(prefix, g, device="cuda", write_keys_file=False)
g = np.fromfile(d + "/genome.u8", dtype=np.uint8)
gt = torch.from_numpy(g).cuda()
if PE:
    fqs = [d + f"/pe{N}_1.fq", d + f"/pe{N}_2.fq"]
    if not os.path.exists(fqs[1]):
        r1, r2 = synth.make_pairs_torch(gt, N, 150, seed=5)
        offs = np.arange(N + 1, dtype=np.int64) * 150
        synth.write_fastq(fqs[0], r1.cpu().numpy().reshape(-1), offs, prefix="p")
        synth.write_fastq(fqs[1], r2.cpu().numpy().reshape(-1), offs, prefix="p")
else:
    fqs = [d + f"/se{N}.fq"]
    if not os.path.exists(fqs[0]):
        reads = synth.make_reads_torch(gt, N, 150, seed=3).cpu().numpy()
        synth.write_fastq(fqs[0], reads.reshape(-1), np.arange(N + 1, dtype=np.int64) * 150)
del gt
isa = "avx512bw" if " avx512bw" in open("/proc/cpuinfo").read() else "avx2"
MODE = os.environ.get("MODE", "3")          # index layout the binaries are compiled for (1: positions only, 3: keys + ISA)
ref = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme_mode{MODE}.{isa}")
b200 = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme-b200_mode{MODE}.{isa}")
lib = os.path.join(ROOT, "bwa-meme_b200", "libmeme_b200.so")
sams = {}
arms = [("reference_cpu", ref, {}), ("gpu_seed", b200, {"MEME_B200": "1"}),
        ("gpu_seed_bsw", b200, {"MEME_B200": "1", "MEME_B200_BSW": "1"}),
        ("gpu_seed_chain_bsw", b200, {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1"}),
        ("gpu_all", b200, {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1", "MEME_B200_MATE": "1"})]
if os.environ.get("ARMS"):
    arms = [a for a in arms if a[0] in os.environ["ARMS"].split(",") or a[0] == "reference_cpu"]
nreads = N * (2 if PE else 1)
for name, exe, env in arms:
    e = dict(os.environ); e.update(env); e["MEME_B200_LIB"] = lib
    if os.environ.get("DEVICES") and name != "reference_cpu":
        e["MEME_B200_DEVICES"] = os.environ["DEVICES"]      # reads and extension pairs shard across these GPUs
    if (name.endswith("_bsw") or name == "gpu_all") and os.environ.get("DUMP"):
        e["MEME_B200_DUMP_PAIRS"] = os.environ["DUMP"]
    t0 = time.time()
    p = subprocess.run([exe, "mem", "-7", "-t", T, "-K", "100000000", prefix, *fqs, "-o", f"{d}/{name}.sam"], capture_output=True, text=True, env=e)
    wall = time.time() - t0
    assert p.returncode == 0, p.stderr[-1500:]
    proc = [l for l in p.stderr.splitlines() if "Processed" in l and "real sec" in l]
    real = sum(float(l.split("CPU sec,")[1].split("real sec")[0]) for l in proc)
    sams[name] = "".join(l for l in open(f"{d}/{name}.sam") if not l.startswith("@PG"))
    out = {"arm": name, "mode": int(MODE), "paired": PE, "genome_bp": G, "threads": int(T), "reads": nreads, "mem_process_seqs_real_s": real,
           "wall_s": wall, "Mbases_per_s": nreads * 150 / real / 1e6, "sam_identical_to_reference": sams[name] == sams["reference_cpu"]}
    gpu_lines = [l for l in p.stderr.splitlines() if "[meme_b200] seeded" in l]
    out["seed_stage_ms_per_chunk"] = [float(l.split("meme_")[-1].split()[1]) for l in gpu_lines if "this chunk" in l]
    if gpu_lines: out["gpu_seed_log"] = gpu_lines[-1]
    cig_lines = [l for l in p.stderr.splitlines() if "[meme_b200] global alignments" in l]
    if cig_lines: out["gpu_cigar_log"] = cig_lines[-1]
    msw_lines = [l for l in p.stderr.splitlines() if "[meme_b200] mate rescue:" in l]
    if msw_lines: out["gpu_matesw_log"] = msw_lines[-2:]
    ext_lines = [l for l in p.stderr.splitlines() if "[meme_b200] extended" in l]
    if ext_lines: out["gpu_extend_log"] = ext_lines[-1]
    if os.environ.get("E2E_VERBOSE"):
        print("\n".join(l for l in p.stderr.splitlines() if "[meme" in l), file=sys.stderr)
    if os.environ.get("E2E_PROFILE"):      # the aligner's own profile table (src/profiling.cpp)
        print(f"---- {name}\n" + p.stderr[-5000:], file=sys.stderr)
    print(json.dumps(out), flush=True)
