"""Generate tests/golden/* from the UNMODIFIED reference compiled under oracle/_ref/.

Run here (where /root/reference is mounted):  python tests/golden/make_golden.py
The fixtures are small and committed; the GPU box only reads them.

  index_sha.json        sha256 of the files `bwa-meme index -a meme` writes for the golden genome
                        (pins bwa-meme_b200/index_build.py against the reference's own builder)
  smem_len*.txt         SMEM dumps of test/Learned_seeding_big_read.cpp (steps=4, 1 thread, MODE=3)
  smem_len*_s{1,2}.json per-read SMEM counts with steps=1 / steps=2 (rounds 1 and 1+2 only)
  bsw_*.npy             {score,tle,gtle,qle,gscore,max_off} of BandedPairWiseSW::scalarBandedSWAWrapper
  mem7.sam              `bwa-meme mem -7` SAM (minus @PG) for the len150 reads  (end-to-end reference)
  pe_reads.npy, mem7_pe.sam   250 read pairs (codes) and the reference's paired-end SAM for them
  chains_len*.txt, chains_rep*.txt   chain_ar after mem_chain_Learned + mem_chain_flt (the patched aligner with every hook off runs
                        the reference's own code; MEME_B200_DUMP_CHAINS prints it): golden read sets, and a repeat-rich
                        three-contig genome (chains_rep.json) with default options and with -c 8 -W 40 -N 3
  cigar_ref.json        CIGAR / NM / MD / position of the reference's mem_reg2aln (oracle/ref_cigar_shim.cpp) for 1500 regions of the
                        repeat-rich genome (tests/chain_util.py: make_cigar_cases)
  smem_textend.txt      SMEM dump for reads that run into the END of the fwd+RC text of a genome with homopolymer ends
                        (synth.make_genome_textend / make_textend_reads): compare_read_and_ref_binary's text-end branch,
                        src/LearnedIndex_seeding.cpp:L266-273
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_pkg()
O = ge.load_oracle()
from bwa_meme_b200 import synth, index_build  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
ISA = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
GENOME_LEN, GENOME_SEED = 50000, 101
READ_SETS = [(150, 300, 0.05, 21), (101, 120, 0.3, 22), (250, 100, 0.05, 23), (33, 80, 0.1, 24)]
BSW_SETS = [(0, 400, 150, 110, 120), (1, 300, 60, 200, 40), (2, 200, 300, 120, 250)]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


TEXTEND = {"length": 30000, "seed": 301, "start_a_run": 6, "end_t_run": 5, "bits": 16, "reads_seed": 1}


MATESW_SETS = (("default", 1200, 11, {}), ("a2b3", 600, 12, {"a": 2, "b": 3}), ("gaps", 600, 13, {"o_del": 1, "e_del": 3, "o_ins": 12, "e_ins": 2}))


def make_matesw(work):
    """mem_sam_pe_batch of the reference (oracle/ref_matesw_shim.cpp; kswv exists in AVX512 builds only) on synthetic rescue pairs."""
    import matesw_util as MU
    if MU.shim_path() is None:
        print("matesw: no AVX512 shim here, fixtures left as they are")
        return
    for name, n, seed, kw in MATESW_SETS:
        recs, ref, qer = MU.make_matesw_cases(n, seed, a=kw.get("a", 1))
        out = MU.run_shim(O, recs, ref, qer, O.default_matesw_opt(**kw), work)
        np.save(os.path.join(HERE, f"matesw_ref_{name}.npy"), out)
        print("matesw", name, len(out), "pairs,", int((out["tb"] >= 0).sum()), "with start positions,", int((out["score2"] > 0).sum()), "with a second hit")


def make_textend(work):
    """Reads crossing the end of the text (and the forward/reverse junction) of a genome whose text ends in a T run."""
    g = synth.make_genome_textend(TEXTEND["length"], TEXTEND["seed"], TEXTEND["start_a_run"], TEXTEND["end_t_run"])
    fa = os.path.join(work, "textend.fa")
    synth.write_fasta(fa, g)
    subprocess.run([os.path.join(REF, f"bwa-meme_mode3.{ISA}"), "index", "-a", "meme", "-t", "4", fa], check=True, capture_output=True,
                   timeout=300)
    mine = os.path.join(work, "textend_mine.fa")
    st = index_build.build_index(mine, g, bits=TEXTEND["bits"])
    for ext in ("_L0_PARAMETERS", "_L1_PARAMETERS", "_L2_PARAMETERS"):
        open(fa + ".suffixarray_uint64" + ext, "wb").write(open(mine + ".suffixarray_uint64" + ext, "rb").read())
    codes, offs = synth.make_textend_reads(g, st["pad"], TEXTEND["start_a_run"], seed=TEXTEND["reads_seed"])
    fq = os.path.join(work, "textend.fq")
    synth.write_fastq(fq, codes, offs)
    out = os.path.join(HERE, "smem_textend.txt")
    with open(out, "w") as fo:
        # (a hard limit: with longer homopolymer ends the reference's own search does not terminate)
        subprocess.run(["timeout", "-s", "KILL", "120", os.path.join(REF, f"learned_seeding_mode3.{ISA}"), fa, fq, "512", "1", "4"],
                       stdout=fo, stderr=subprocess.DEVNULL, check=True)
    json.dump({**TEXTEND, "pad": int(st["pad"]), "n_reads": int(len(offs) - 1)}, open(os.path.join(HERE, "textend.json"), "w"), indent=1)


CHAINS_REP = {"length": 120000, "seed": 909, "n_families": 4, "copies": 150, "rep_len": 200, "divergence": 0.02,
              "tandem_unit": 37, "tandem_copies": 60, "tandem_at": 70000, "contigs": [["c0", 0, 40000, 0], ["c1", 40000, 35000, 1], ["c2", 75000, 45000, 0]],
              "n_reads": 500, "read_len": 150, "reads_seed": 31, "bits": 14}


def chains_rep_genome():
    """Repeat-rich three-contig genome (the middle contig ALT) for the chaining fixtures: 150-copy dispersed repeats and a
    tandem array, so that reads carry tens to hundreds of chains (B-tree splits, equal positions, ties in the weight sort)."""
    m = CHAINS_REP
    g = synth.make_genome(m["length"], seed=m["seed"], n_families=m["n_families"], copies=m["copies"], rep_len=m["rep_len"],
                          divergence=m["divergence"])
    rng = np.random.default_rng(m["seed"] + 1)
    unit = rng.integers(0, 4, size=m["tandem_unit"], dtype=np.uint8)
    arr = np.tile(unit, m["tandem_copies"])
    mut = rng.random(arr.size) < 0.01
    arr[mut] = (arr[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
    g[m["tandem_at"]:m["tandem_at"] + arr.size] = arr
    return g


def chains_rep_reads(g):
    m = CHAINS_REP
    codes, offs = synth.make_reads(g, m["n_reads"], m["read_len"], seed=m["reads_seed"], n_frac=0.02)
    return codes, offs


def dump_chains(prefix, fq, out, extra=()):
    """chain_ar of every read as the UNMODIFIED reference code leaves it: the patched aligner with every hook off."""
    env = dict(os.environ, MEME_B200_DUMP_CHAINS=out)
    for k in ("MEME_B200", "MEME_B200_BSW", "MEME_B200_CHAIN"):
        env.pop(k, None)
    subprocess.run([os.path.join(REF, f"bwa-meme-b200_mode3.{ISA}"), "mem", "-7", "-t", "1", *extra, prefix, fq, "-o", "/dev/null"], check=True,
                   capture_output=True, env=env, timeout=600)
    lines = sorted(open(out).read().splitlines(), key=lambda l: int(l.split()[0][1:]))
    open(out, "w").write("".join(l + "\n" for l in lines))


def make_chains(work, fa, g):
    """Chain dumps (mem_chain_Learned + mem_chain_flt of the reference) for the golden read sets and the repeat-rich genome."""
    for rl, n, nf, seed in READ_SETS:
        dump_chains(fa, os.path.join(work, f"r{rl}.fq"), os.path.join(HERE, f"chains_len{rl}.txt"))
    m = CHAINS_REP
    gr = chains_rep_genome()
    far = os.path.join(work, "rep.fa")
    with open(far, "wb") as f:
        for name, off, ln, alt in m["contigs"]:
            tmp = os.path.join(work, "_c.fa")
            synth.write_fasta(tmp, gr[off:off + ln], name=name)
            f.write(open(tmp, "rb").read())
    subprocess.run([os.path.join(REF, f"bwa-meme_mode3.{ISA}"), "index", "-a", "meme", "-t", "4", far], check=True, capture_output=True, timeout=600)
    open(far + ".alt", "w").write("".join(f"{name}\t0\t*\t0\t0\t*\t*\t0\t0\t*\t*\n" for name, off, ln, alt in m["contigs"] if alt))
    mine = os.path.join(work, "rep_mine.fa")
    index_build.build_index(mine, gr, bits=m["bits"])
    assert open(mine + ".pos_packed", "rb").read() == open(far + ".pos_packed", "rb").read()
    for ext in ("_L0_PARAMETERS", "_L1_PARAMETERS", "_L2_PARAMETERS"):
        open(far + ".suffixarray_uint64" + ext, "wb").write(open(mine + ".suffixarray_uint64" + ext, "rb").read())
    codes, offs = chains_rep_reads(gr)
    fq = os.path.join(work, "rep.fq")
    synth.write_fastq(fq, codes, offs)
    dump_chains(far, fq, os.path.join(HERE, "chains_rep.txt"))
    # the same reads with a small max_occ (hit sampling), a chain-weight floor and a chain-extension cap
    dump_chains(far, fq, os.path.join(HERE, "chains_rep_c8_W40_N3.txt"), extra=("-c", "8", "-W", "40", "-N", "3"))
    json.dump(m, open(os.path.join(HERE, "chains_rep.json"), "w"), indent=1)


def make_cigar(work):
    """mem_reg2aln of the reference (oracle/ref_cigar_shim.cpp) on regions of the repeat-rich genome of the chaining fixtures."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import chain_util as U
    gr = chains_rep_genome()
    codes, cases = U.make_cigar_cases(gr, CHAINS_REP["contigs"], 1500, 4)
    out = U.run_cigar_shim(os.path.join(REF, f"cigar_shim_mode3.{ISA}"), os.path.join(work, "rep.fa"), codes, cases, work)
    assert len(out) == len(cases)
    json.dump(out, open(os.path.join(HERE, "cigar_ref.json"), "w"))


def main():
    work = "/tmp/meme_golden"
    os.makedirs(work, exist_ok=True)
    g = synth.make_genome(GENOME_LEN, seed=GENOME_SEED, n_families=3, copies=12, rep_len=400)
    fa = os.path.join(work, "ref.fa")
    synth.write_fasta(fa, g)
    subprocess.run([os.path.join(REF, f"bwa-meme_mode3.{ISA}"), "index", "-a", "meme", "-t", "4", fa], check=True,
                   capture_output=True)
    shas = {ext: sha(fa + ext) for ext in (".pac", ".ann", ".amb", ".0123", ".pos_packed", ".suffixarray_uint64")}
    json.dump({"genome_len": GENOME_LEN, "genome_seed": GENOME_SEED, "sha256": shas},
              open(os.path.join(HERE, "index_sha.json"), "w"), indent=1)
    # parameters from our trainer, written next to the reference-built files
    mine = os.path.join(work, "mine.fa")
    index_build.build_index(mine, g, bits=16)
    for ext in ("_L0_PARAMETERS", "_L1_PARAMETERS", "_L2_PARAMETERS"):
        src = mine + ".suffixarray_uint64" + ext
        open(fa + ".suffixarray_uint64" + ext, "wb").write(open(src, "rb").read())
    for rl, n, nf, seed in READ_SETS:
        codes, offs = synth.make_reads(g, n, rl, seed=seed, n_frac=nf)
        fq = os.path.join(work, f"r{rl}.fq")
        synth.write_fastq(fq, codes, offs)
        out = os.path.join(HERE, f"smem_len{rl}.txt")
        with open(out, "w") as fo:
            subprocess.run([os.path.join(REF, f"learned_seeding_mode3.{ISA}"), fa, fq, "512", "1", "4"], stdout=fo,
                           stderr=subprocess.DEVNULL, check=True)
        for steps in (1, 2):
            p = subprocess.run([os.path.join(REF, f"learned_seeding_mode3.{ISA}"), fa, fq, "512", "1", str(steps)],
                               capture_output=True, text=True, check=True)
            tot = [l for l in p.stderr.splitlines() if l.startswith("totalSmems")]
            json.dump({"totalSmems": int(tot[0].split("=")[1])}, open(os.path.join(HERE, f"smem_len{rl}_s{steps}.json"), "w"))
        if rl == 150:
            sam = subprocess.run([os.path.join(REF, f"bwa-meme_mode3.{ISA}"), "mem", "-7", "-t", "1", fa, fq],
                                 capture_output=True, text=True, check=True).stdout
            open(os.path.join(HERE, "mem7.sam"), "w").write("".join(l + "\n" for l in sam.splitlines() if not l.startswith("@PG")))
    # paired-end: 250 pairs, insert ~ N(400, 50), FR orientation  (exercises mem_pestat + mate rescue downstream)
    rng = np.random.default_rng(77)
    n_pairs, rl = 250, 150
    r1, r2 = [], []
    for _ in range(n_pairs):
        ins = int(np.clip(rng.normal(400, 50), 2 * rl // 2 + 20, 700))
        st = int(rng.integers(0, GENOME_LEN - ins))
        frag = g[st:st + ins].copy()
        a, b = frag[:rl].copy(), synth.revcomp(frag[-rl:])
        for rd in (a, b):
            m = rng.random(rl) < 0.01
            rd[m] = (rd[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        if rng.random() < 0.5:
            a, b = b, a
        r1.append(a); r2.append(b)
    offs = np.arange(n_pairs + 1, dtype=np.int64) * rl
    synth.write_fastq(os.path.join(work, "pe_1.fq"), np.concatenate(r1), offs, prefix="p")
    synth.write_fastq(os.path.join(work, "pe_2.fq"), np.concatenate(r2), offs, prefix="p")
    np.save(os.path.join(HERE, "pe_reads.npy"), np.stack([np.concatenate(r1), np.concatenate(r2)]))
    sam = subprocess.run([os.path.join(REF, f"bwa-meme_mode3.{ISA}"), "mem", "-7", "-t", "1", fa, os.path.join(work, "pe_1.fq"),
                          os.path.join(work, "pe_2.fq")], capture_output=True, text=True, check=True).stdout
    open(os.path.join(HERE, "mem7_pe.sam"), "w").write("".join(l + "\n" for l in sam.splitlines() if not l.startswith("@PG")))
    make_textend(work)
    make_chains(work, fa, g)
    make_cigar(work)
    make_matesw(work)
    shim = os.path.join(REF, f"bsw_shim_mode3.{ISA}")
    for seed, n, mq, mx, h0m in BSW_SETS:
        pairs, ref, qer = O.make_pairs(n, seed=seed, max_q=mq, max_extra_ref=mx, h0_max=h0m)
        for w in (100, 200, 7):
            r = O.run_ref_bsw_shim(shim, pairs, ref, qer, w, "scalar")
            res = np.stack([r[f] for f in O.RESULT_FIELDS], axis=1).astype(np.int32)
            np.save(os.path.join(HERE, f"bsw_scalar_s{seed}_w{w}.npy"), res)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
