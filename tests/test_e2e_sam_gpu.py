"""End-to-end drop-in check: the reference aligner with the 13-line integration patch, seeding (and seed
extension) routed through libmeme_b200.so, must print the same SAM as the unmodified reference CPU run
(golden files written by tests/golden/make_golden.py).  Needs oracle/_ref (built where the reference is
mounted; travels to the GPU box)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _binary():
    flags = open("/proc/cpuinfo").read()
    for isa in (["avx512bw"] if " avx512bw" in flags else []) + ["avx2"]:
        p = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme-b200_mode3.{isa}")
        if os.path.exists(p):
            return p
    return None


def _run(exe, args, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    env["MEME_B200_LIB"] = os.path.join(ROOT, "bwa-meme_b200", "libmeme_b200.so")
    p = subprocess.run([exe, "mem", "-7", "-t", "4", *args], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    sam = "".join(l + "\n" for l in p.stdout.splitlines() if not l.startswith("@PG"))
    return sam, p.stderr


@pytest.mark.parametrize("hooks", [{"MEME_B200": "1"}, {"MEME_B200": "1", "MEME_B200_BSW": "1"}, {"MEME_B200": "1", "MEME_B200_CHAIN": "1"},
                                   {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1"},
                                   {"MEME_B200": "1", "MEME_B200_SAM": "1"},
                                   {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1"}])
def test_single_end_sam_identical(pkg, golden_index, golden_reads, tmp_path, hooks):
    from bwa_meme_b200 import synth
    exe = _binary()
    if exe is None:
        pytest.skip("oracle/_ref/bwa-meme-b200_* not built (needs the reference sources)")
    codes, offs = golden_reads[150]
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, codes, offs)
    sam, err = _run(exe, [golden_index, fq], hooks)
    assert ("seeded and chained 300 reads on the GPU" if "MEME_B200_CHAIN" in hooks else "seeded 300 reads on the GPU") in err
    if "MEME_B200_SAM" in hooks:
        assert "global alignments of" in err
    assert sam == open(os.path.join(GOLDEN, "mem7.sam")).read()


def test_paired_end_sam_identical(pkg, golden_index, tmp_path):
    from bwa_meme_b200 import synth
    exe = _binary()
    if exe is None:
        pytest.skip("oracle/_ref/bwa-meme-b200_* not built (needs the reference sources)")
    pe = np.load(os.path.join(GOLDEN, "pe_reads.npy"))
    n = pe.shape[1] // 150
    offs = np.arange(n + 1, dtype=np.int64) * 150
    f1, f2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
    synth.write_fastq(f1, pe[0], offs, prefix="p")
    synth.write_fastq(f2, pe[1], offs, prefix="p")
    sam, err = _run(exe, [golden_index, f1, f2], {"MEME_B200": "1", "MEME_B200_BSW": "1"})
    assert "on the GPU" in err
    assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()
    sam, err = _run(exe, [golden_index, f1, f2], {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1"})
    assert "seeded and chained" in err
    assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()
    # + the global alignments of the SAM stage; regions created later by mate rescue take the reference's own code
    sam, err = _run(exe, [golden_index, f1, f2], {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1"})
    assert "global alignments of" in err
    assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()
    # + the mate-rescue Smith-Waterman of the whole chunk (row f2): alone, and with every other stage
    for hooks in ({"MEME_B200": "1", "MEME_B200_MATE": "1"},
                  {"MEME_B200": "1", "MEME_B200_MATE": "1", "MEME_B200_MATE_RESTORE": "1"},      # worker_sam skips its own collection pass
                  {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1", "MEME_B200_MATE": "1"}):
        sam, err = _run(exe, [golden_index, f1, f2], hooks)
        assert "mate rescue:" in err and "Smith-Waterman pairs" in err
        if " avx512bw" in open("/proc/cpuinfo").read():        # kswv (and the seam in worker_sam) exists in AVX512 builds only
            taken = [l for l in err.splitlines() if "Smith-Waterman results in" in l]
            assert taken and int(taken[-1].split("mate rescue:")[1].split()[0]) > 0, err[-2000:]
        assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()


def _ref_binary():
    flags = open("/proc/cpuinfo").read()
    for isa in (["avx512bw"] if " avx512bw" in flags else []) + ["avx2"]:
        p = os.path.join(ROOT, "oracle", "_ref", f"bwa-meme_mode3.{isa}")
        if os.path.exists(p):
            return p
    return None


def test_high_copy_repeats_sampled_hits_sam_identical(pkg, tmp_path):
    """SMEMs with more than max_occ (500) occurrences: the glue asks the kernel for the sampled hits only
    (src/bwamem.cpp:L1154-1155) and scatters them to the strided slots mem_chain_Learned reads.  The SAM must equal the
    unmodified reference binary's, run here on the same index and reads."""
    from bwa_meme_b200 import synth, index_build
    exe, ref = _binary(), _ref_binary()
    if exe is None or ref is None:
        pytest.skip("oracle/_ref binaries not built (needs the reference sources)")
    rng = np.random.default_rng(41)
    g = synth.make_genome(300000, seed=41, n_families=2, copies=10, rep_len=300)
    unit = rng.integers(0, 4, size=70, dtype=np.uint8)
    for p in rng.choice(np.arange(1000, 299000, 90), size=900, replace=False):      # 900 exact copies of a 70-mer
        g[p:p + 70] = unit
    prefix = str(tmp_path / "rep.fa")
    index_build.build_index(prefix, g, bits=14)
    codes, offs = synth.make_reads(g, 400, 150, seed=8)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, codes, offs)
    want, _ = _run(ref, [prefix, fq], {})
    got, err = _run(exe, [prefix, fq], {"MEME_B200": "1"})
    assert "seeded 400 reads on the GPU" in err
    assert got == want
    got, err = _run(exe, [prefix, fq], {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1"})
    assert "seeded and chained 400 reads on the GPU" in err
    assert got == want
    got, err = _run(exe, [prefix, fq], {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1"})
    assert "global alignments of" in err            # reads in 900-copy repeats: many regions per read, XA candidates
    assert got == want
    # the case is really exercised: some SMEM has more than 500 hits
    from bwa_meme_b200 import capi
    ix = capi.MemeIndex(prefix, mode=3)
    r = ix.seed(codes, offs)
    ix.close()
    assert int(r["smems"]["hitcount"].max()) > 500


def test_two_gpus_sam_identical(pkg, golden_index, tmp_path):
    """Reads (pair-aligned) and extension pairs shard across two GPUs through the C ABI (meme_index_load_devices,
    meme_seed_batch_multi, segments dealt to one kernel-B handle per device): SAM must not depend on the GPU count."""
    import torch
    from bwa_meme_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = _binary()
    if exe is None:
        pytest.skip("oracle/_ref/bwa-meme-b200_* not built (needs the reference sources)")
    pe = np.load(os.path.join(GOLDEN, "pe_reads.npy"))
    n = pe.shape[1] // 150
    offs = np.arange(n + 1, dtype=np.int64) * 150
    f1, f2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
    synth.write_fastq(f1, pe[0], offs, prefix="p")
    synth.write_fastq(f2, pe[1], offs, prefix="p")
    sam, err = _run(exe, [golden_index, f1, f2], {"MEME_B200": "1", "MEME_B200_BSW": "1", "MEME_B200_DEVICES": "0,1"})
    assert "resident on 2 GPU(s)" in err
    assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()
    sam, err = _run(exe, [golden_index, f1, f2], {"MEME_B200": "1", "MEME_B200_CHAIN": "1", "MEME_B200_BSW": "1", "MEME_B200_SAM": "1",
                                                  "MEME_B200_MATE": "1", "MEME_B200_DEVICES": "0,1"})
    assert "seeded and chained" in err and "global alignments of" in err and "mate rescue:" in err
    assert sam == open(os.path.join(GOLDEN, "mem7_pe.sam")).read()


def test_chain_dumps_identical_to_reference_code(pkg, tmp_path):
    """chain_ar after the chaining stage, printed by the same binary: GPU chains (MEME_B200_CHAIN=1) against the reference's
    own mem_chain_Learned + mem_chain_flt (every hook off), on the repeat-rich three-contig genome of the chaining
    fixtures, with default options and with -c 8 -W 40 -N 3; and both against the committed dump of the reference."""
    from bwa_meme_b200 import synth, index_build
    from chain_util import rep_genome, rep_reads, rep_meta, read_dump
    exe = _binary()
    if exe is None:
        pytest.skip("oracle/_ref/bwa-meme-b200_* not built (needs the reference sources)")
    m = rep_meta()
    g = rep_genome(synth)
    prefix = str(tmp_path / "rep.fa")
    index_build.build_index(prefix, g, bits=m["bits"])
    # our builder writes one contig; the fixture has three (the middle one ALT): rewrite .ann / .alt as bwa index would
    with open(prefix + ".ann", "w") as f:
        f.write(f"{len(g)} {len(m['contigs'])} 11\n")
        for name, off, ln, alt in m["contigs"]:
            f.write(f"0 {name} (null)\n{off} {ln} 0\n")
    open(prefix + ".amb", "w").write(f"{len(g)} {len(m['contigs'])} 0\n")
    open(prefix + ".alt", "w").write("".join(f"{name}\t0\t*\t0\t0\t*\t*\t0\t0\t*\t*\n" for name, off, ln, alt in m["contigs"] if alt))
    codes, offs = rep_reads(synth, g)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, codes, offs)

    def dump(hooks, extra, tag):
        out = str(tmp_path / f"chains_{tag}.txt")
        env = dict(hooks, MEME_B200_DUMP_CHAINS=out)
        sam, err = _run(exe, [*extra, prefix, fq], env)
        lines = sorted(open(out).read().splitlines(), key=lambda l: int(l.split()[0][1:]))
        return lines, sam, err

    for extra, golden in (([], "chains_rep.txt"), (["-c", "8", "-W", "40", "-N", "3"], "chains_rep_c8_W40_N3.txt")):
        cpu, sam_cpu, _ = dump({}, extra, "cpu")
        gpu, sam_gpu, err = dump({"MEME_B200": "1", "MEME_B200_CHAIN": "1"}, extra, "gpu")
        assert "seeded and chained 500 reads on the GPU" in err
        assert cpu == read_dump(golden)
        assert gpu == cpu
        assert sam_gpu == sam_cpu
