"""The seeding oracle against the reference's own outputs (golden SMEM dumps) -- CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_READ_SETS, ROOT


@pytest.mark.parametrize("rl", [s[0] for s in GOLDEN_READ_SETS])
@pytest.mark.parametrize("mode", [3, 2, 1])
def test_oracle_matches_reference_dump(oracle, golden_index, golden_reads, rl, mode):
    codes, offs = golden_reads[rl]
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, f"smem_len{rl}.txt"), n)
    ix = oracle.OracleIndex(golden_index, mode)
    r = ix.seed(codes, offs)
    got = oracle.canon_lines(r["smems"], r["smem_off"], r["hit_off"], r["hits"], n)
    ix.close()
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, f"{len(bad)} reads differ from the reference dump, first: {bad[:5]}"
    assert sum(len(x) for x in want) > 0


@pytest.mark.parametrize("rl", [150, 101])
@pytest.mark.parametrize("steps", [1, 2])
def test_oracle_round_counts(oracle, golden_index, golden_reads, rl, steps):
    """steps=1 (step1 only) and steps=2 (+reseeding) totals printed by the reference harness."""
    codes, offs = golden_reads[rl]
    want = json.load(open(os.path.join(GOLDEN, f"smem_len{rl}_s{steps}.json")))["totalSmems"]
    ix = oracle.OracleIndex(golden_index, 3)
    r = ix.seed(codes, offs, oracle.default_opt(rounds=steps))
    ix.close()
    assert len(r["smems"]) == want


def test_oracle_edge_reads(oracle, golden_index, golden_genome):
    """empty read, all-N read, read shorter than min_seed_len, exact full-length match."""
    g = golden_genome
    reads = [np.zeros(0, np.uint8), np.full(50, 4, np.uint8), g[100:110].copy(), g[500:650].copy(),
             np.concatenate([g[700:760], [4], g[761:850]]).astype(np.uint8)]
    offs = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    codes = np.concatenate(reads).astype(np.uint8)
    ix = oracle.OracleIndex(golden_index, 3)
    r = ix.seed(codes, offs)
    ix.close()
    so = r["smem_off"]
    assert so[1] - so[0] == 0 and so[2] - so[1] == 0 and so[3] - so[2] == 0
    full = r["smems"][so[3]:so[4]]
    assert any(s["start"] == 0 and s["end"] == 150 for s in full)
    hits = r["hits"][r["hit_off"][3]:r["hit_off"][4]]
    assert 500 in hits.tolist()
    withn = r["smems"][so[4]:so[5]]
    assert all(not (s["start"] <= 60 < s["end"]) for s in withn), "an SMEM spans an N"


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")), reason="oracle/_ref not built")
def test_oracle_vs_live_reference(oracle, pkg, tmp_path):
    """Fresh genome + reads through the compiled reference harness (only where oracle/_ref exists)."""
    import subprocess
    from bwa_meme_b200 import synth, index_build
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    exe = os.path.join(ROOT, "oracle", "_ref", f"learned_seeding_mode3.{isa}")
    if not os.path.exists(exe):
        pytest.skip("harness binary missing")
    g = synth.make_genome(120000, seed=901)
    prefix = str(tmp_path / "live.fa")
    index_build.build_index(prefix, g, bits=14)
    codes, offs = synth.make_reads(g, 400, 150, seed=902, n_frac=0.1)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, codes, offs)
    dump = str(tmp_path / "dump.txt")
    with open(dump, "w") as fo:
        subprocess.run([exe, prefix, fq, "512", "1", "4"], stdout=fo, stderr=subprocess.DEVNULL, check=True)
    n = len(offs) - 1
    want = oracle.parse_ref_dump(dump, n)
    ix = oracle.OracleIndex(prefix, 3)
    r = ix.seed(codes, offs)
    ix.close()
    got = oracle.canon_lines(r["smems"], r["smem_off"], r["hit_off"], r["hits"], n)
    assert got == want


def test_oracle_matches_reference_at_the_text_end(oracle, textend_case):
    """Reads that run into the end of the fwd+RC text (homopolymer T end), across the forward/reverse junction, and with
    T tails longer than the index's padding run: compare_read_and_ref_binary's text-end branch
    (src/LearnedIndex_seeding.cpp:L266-273).  Fixture = the reference's own dump (tests/golden/make_golden.py)."""
    prefix, codes, offs = textend_case
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, "smem_textend.txt"), n)
    for mode in (3, 1):
        oix = oracle.OracleIndex(prefix, mode)
        got = oix.seed(codes, offs)
        oix.close()
        a = oracle.canon_lines(got["smems"], got["smem_off"], got["hit_off"], got["hits"], n)
        bad = [i for i in range(n) if a[i] != want[i]]
        assert not bad, f"mode {mode}: {len(bad)} reads differ, first {bad[:5]}"
