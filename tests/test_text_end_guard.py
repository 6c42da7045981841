"""The corner the reference does not survive: a fwd+RC text that ENDS WITH ITS LONGEST T RUN.

`bwa-meme index` pads the text with (longest A/T run + 1) T's, so the suffixes inside that final run sort behind every
other suffix, while compare_read_and_ref_binary caps their match at the text end (src/LearnedIndex_seeding.cpp:L266-273).
For a read that matches into the text end the longest match then hides behind the comparator's insertion point, the
right search of the zig-zag comes back shorter than the left one, and Learned_getSMEMsOnePosOneThread_step1
(L1724-1849) never advances: the reference's own seeding harness does not terminate (checked live below where the
reference is mounted).  No output exists to be bit-exact with; what is pinned here is that THIS repo terminates -- oracle
and kernels carry the same progress guard -- and that oracle and kernels agree with each other on those reads.
With a shorter final run (tests/golden/smem_textend.txt) the reference terminates and all three agree with it."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def longest_run_case(pkg, tmp_path_factory):
    from bwa_meme_b200 import synth, index_build
    g = synth.make_genome_textend(30000, seed=301, start_a_run=12, end_t_run=5)      # text ends with T^12
    d = tmp_path_factory.mktemp("textend_guard")
    prefix = str(d / "g.fa")
    st = index_build.build_index(prefix, g, bits=16)
    assert st["pad"] == 13, st                              # the final run IS the longest run
    codes, offs = synth.make_textend_reads(g, st["pad"], 12, seed=2)
    return prefix, g, codes, offs


@pytest.mark.timeout(300)
def test_oracle_terminates_on_the_reference_hang(oracle, longest_run_case):
    prefix, _, codes, offs = longest_run_case
    oix = oracle.OracleIndex(prefix, 3)
    r = oix.seed(codes, offs)
    oix.close()
    assert len(r["smem_off"]) == len(offs)


@pytest.mark.timeout(600)
def test_kernel_source_on_cpu_terminates_and_matches_oracle(oracle, longest_run_case):
    import test_flat_host_sim as hs
    prefix, _, codes, offs = longest_run_case
    n = len(offs) - 1
    oix = oracle.OracleIndex(prefix, 3)
    want = oix.seed(codes, offs)
    oix.close()
    b = oracle.canon_lines(want["smems"], want["smem_off"], want["hit_off"], want["hits"], n)
    for impl in ("flat", "grp"):
        sim = hs.make_sim(impl)
        got = sim(prefix, 3, codes, offs)
        a = oracle.canon_lines(got["smems"], got["smem_off"], got["hit_off"], got["hits"], n)
        bad = [i for i in range(n) if a[i] != b[i]]
        assert not bad, f"{impl}: {len(bad)} reads differ from the oracle, first {bad[:5]}"


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_cuda_terminates_and_matches_oracle(oracle, longest_run_case):
    from bwa_meme_b200 import capi
    prefix, _, codes, offs = longest_run_case
    n = len(offs) - 1
    oix = oracle.OracleIndex(prefix, 3)
    want = oix.seed(codes, offs)
    oix.close()
    b = oracle.canon_lines(want["smems"], want["smem_off"], want["hit_off"], want["hits"], n)
    for mode in (3, 1):
        ix = capi.MemeIndex(prefix, mode=mode)
        got = ix.seed(codes, offs)
        ix.close()
        a = oracle.canon_lines(got["smems"], got["smem_off"], got["hit_off"], got["hits"], n)
        bad = [i for i in range(n) if a[i] != b[i]]
        assert not bad, f"mode {mode}: {len(bad)} reads differ from the oracle, first {bad[:5]}"


def test_live_reference_does_not_terminate_here(pkg, longest_run_case, tmp_path):
    """Where the reference is built (oracle/_ref): its seeding harness hangs on ONE read that matches into the text end
    (killed after 5 s; the whole index loads in a fraction of a second), and finishes at once on the same read of a
    genome whose final run is short."""
    from bwa_meme_b200 import synth
    isa = "avx512bw" if " avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    exe = os.path.join(ROOT, "oracle", "_ref", f"learned_seeding_mode3.{isa}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/learned_seeding_* not in this tree (git-ignored, not shipped to the GPU box)")
    prefix, g, _, _ = longest_run_case
    text = np.concatenate([g, (3 - g[::-1]).astype(np.uint8)])
    rng = np.random.default_rng(1)
    read = np.concatenate([text[-20:], np.full(3, 3, np.uint8), [1], rng.integers(0, 4, 126).astype(np.uint8)]).astype(np.uint8)
    fq = str(tmp_path / "one.fq")
    synth.write_fastq(fq, read, np.array([0, 150], dtype=np.int64))
    p = subprocess.run(["timeout", "-s", "KILL", "5", exe, prefix, fq, "512", "1", "4"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert p.returncode != 0, "the reference terminated: the corner is gone, revisit DESIGN.md section 2"
