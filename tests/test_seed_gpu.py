"""Parity of the CUDA seeding kernel (through the C ABI) with the oracle and the reference's golden dumps."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_READ_SETS

pytestmark = pytest.mark.gpu


def _canon(O, r, n):
    return O.canon_lines(r["smems"], r["smem_off"], r["hit_off"], r["hits"], n)


@pytest.fixture(scope="module")
def gidx(pkg, golden_index):
    from bwa_meme_b200 import capi
    out = {m: capi.MemeIndex(golden_index, mode=m) for m in (1, 2, 3)}
    yield out
    for v in out.values():
        v.close()


@pytest.mark.parametrize("rl", [s[0] for s in GOLDEN_READ_SETS])
@pytest.mark.parametrize("mode", [3, 2, 1])
def test_cuda_matches_reference_dump(oracle, gidx, golden_reads, rl, mode):
    codes, offs = golden_reads[rl]
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, f"smem_len{rl}.txt"), n)
    got = _canon(oracle, gidx[mode].seed(codes, offs), n)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, f"{len(bad)} reads differ, first {bad[:5]}: {got[bad[0]][:4]} vs {want[bad[0]][:4]}"


@pytest.mark.parametrize("rounds", [1, 2, 3])
def test_cuda_matches_oracle_rounds(oracle, pkg, gidx, golden_index, golden_reads, rounds):
    from bwa_meme_b200 import capi
    codes, offs = golden_reads[150]
    n = len(offs) - 1
    oix = oracle.OracleIndex(golden_index, 3)
    want = oix.seed(codes, offs, oracle.default_opt(rounds=rounds))
    oix.close()
    got = gidx[3].seed(codes, offs, capi.default_seed_opt(rounds=rounds))
    assert _canon(oracle, got, n) == _canon(oracle, want, n)
    # identical ordering too (sorted by start,end; ties carry identical hit lists)
    assert (got["smems"]["start"] == want["smems"]["start"]).all() and (got["smems"]["end"] == want["smems"]["end"]).all()
    assert (got["smems"]["hitcount"] == want["smems"]["hitcount"]).all()
    assert (got["smem_off"] == want["smem_off"]).all() and (got["hit_off"] == want["hit_off"]).all()
    assert (got["hits"] == want["hits"]).all()


def test_wide_windows_repeats_and_all_lengths(oracle, pkg, mid_index, monkeypatch):
    """400 kb genome, 2^12 leaves (RMI windows of hundreds of entries -> 32-ary narrowing), 60-copy repeat
    families (intervals larger than one 32-entry block -> galloping), ragged read lengths 20..320."""
    from bwa_meme_b200 import capi, synth
    prefix, g = mid_index
    codes, offs = synth.make_reads(g, 3000, 320, seed=5, n_frac=0.05, len_jitter=300)
    n = len(offs) - 1
    monkeypatch.setenv("MEME_SEED_STATS", "1")          # the work counters are off by default (they cost registers)
    for mode in (3, 1):
        ix = capi.MemeIndex(prefix, mode=mode)
        got = ix.seed(codes, offs)
        cnt = ix.seed_counters()
        ix.close()
        oix = oracle.OracleIndex(prefix, mode)
        want = oix.seed(codes, offs)
        oix.close()
        a, b = _canon(oracle, got, n), _canon(oracle, want, n)
        bad = [i for i in range(n) if a[i] != b[i]]
        assert not bad, f"mode {mode}: {len(bad)} reads differ, first {bad[:5]}"
        assert cnt["kary_rounds"] > 0 and cnt["ext_rounds"] > 0, cnt


def test_edge_reads(oracle, pkg, gidx, golden_index, golden_genome):
    g = golden_genome
    reads = [np.zeros(0, np.uint8), np.full(50, 4, np.uint8), g[100:110].copy(), g[500:650].copy(),
             np.concatenate([g[700:760], [4], g[761:850]]).astype(np.uint8), np.full(150, 0, np.uint8),
             np.tile(np.array([0, 1], np.uint8), 75), g[-150:].copy(), (3 - g[:150][::-1]).astype(np.uint8),
             np.full(1, 2, np.uint8), np.concatenate([[4], g[900:1000], [4]]).astype(np.uint8)]
    offs = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=offs[1:])
    codes = np.concatenate(reads).astype(np.uint8)
    n = len(reads)
    oix = oracle.OracleIndex(golden_index, 3)
    want = oix.seed(codes, offs)
    oix.close()
    got = gidx[3].seed(codes, offs)
    assert _canon(oracle, got, n) == _canon(oracle, want, n)


def test_empty_batch(pkg, gidx):
    r = gidx[3].seed(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(r["smems"]) == 0 and r["smem_off"].tolist() == [0]


def test_too_long_read_is_rejected(pkg, gidx):
    from bwa_meme_b200 import capi
    with pytest.raises(capi.MemeError):
        gidx[3].seed(np.zeros(400, np.uint8), np.array([0, 400], np.int64))


def test_max_occ_sampling(oracle, pkg, mid_index):
    """max_occ > 0 returns exactly the hits mem_chain_Learned would read (bwamem.cpp:L1154-1155)."""
    from bwa_meme_b200 import capi, synth
    prefix, g = mid_index
    codes, offs = synth.make_reads(g, 500, 150, seed=9)
    ix = capi.MemeIndex(prefix, mode=3)
    full = ix.seed(codes, offs)
    occ = 7
    samp = ix.seed(codes, offs, capi.default_seed_opt(max_occ=occ))
    ix.close()
    assert (full["smems"]["hitcount"] == samp["smems"]["hitcount"]).all()
    n = len(offs) - 1
    saw_big = False
    for r in range(n):
        for k in range(full["smem_off"][r], full["smem_off"][r + 1]):
            c = int(full["smems"]["hitcount"][k])
            step = c // occ if c > occ else 1
            idx = list(range(0, c, step))[:occ]
            saw_big |= c > occ
            fh = full["hits"][full["hit_off"][r] + full["smems"]["hitbeg"][k]:][:c]
            sh = samp["hits"][samp["hit_off"][r] + samp["smems"]["hitbeg"][k]:][:len(idx)]
            assert sh.tolist() == fh[idx].tolist()
    assert saw_big


def test_invariants_at_scale(oracle, pkg, mid_index):
    """size-independent properties on a batch the oracle is not run on in full: every hit really matches
    the read substring on the text; SMEMs sorted; repeated run is bit-identical; shard-and-concatenate equals
    the whole batch (what the multi-GPU path relies on)."""
    from bwa_meme_b200 import capi, synth, index_build
    prefix, g = mid_index
    text = index_build.fwd_rc_text(g)
    codes, offs = synth.make_reads_fast(g, 60000, 150, seed=31)
    n = len(offs) - 1
    ix = capi.MemeIndex(prefix, mode=3)
    a = ix.seed(codes, offs)
    b = ix.seed(codes, offs)
    for k in ("smems", "smem_off", "hit_off", "hits"):
        assert (a[k] == b[k]).all()
    half = n // 2
    p1 = ix.seed(codes[:offs[half]], offs[:half + 1])
    p2 = ix.seed(codes[offs[half]:], offs[half:] - offs[half])
    ix.close()
    assert (np.concatenate([p1["smems"], p2["smems"]]) == a["smems"]).all()
    assert (np.concatenate([p1["hits"], p2["hits"]]) == a["hits"]).all()
    sm = a["smems"]
    rng = np.random.default_rng(0)
    for r in rng.integers(0, n, 400):
        s = sm[a["smem_off"][r]:a["smem_off"][r + 1]]
        key = s["start"].astype(np.int64) * 1000 + s["end"]
        assert (np.diff(key) >= 0).all()
        read = codes[offs[r]:offs[r + 1]]
        for m in s:
            hits = a["hits"][a["hit_off"][r] + m["hitbeg"]:][:m["hitcount"]]
            sub = read[m["start"]:m["end"]]
            for h in hits[:3]:
                assert (text[int(h):int(h) + len(sub)] == sub).all()
    # checksum of checksums against the oracle on a sample
    sel = rng.choice(n, 1500, replace=False)
    sel.sort()
    c2 = np.concatenate([codes[offs[i]:offs[i + 1]] for i in sel])
    o2 = np.arange(len(sel) + 1, dtype=np.int64) * 150
    oix = oracle.OracleIndex(prefix, 3)
    want = oix.seed(c2, o2)
    oix.close()
    got_cnt = np.array([a["smem_off"][i + 1] - a["smem_off"][i] for i in sel])
    assert (got_cnt == np.diff(want["smem_off"])).all()
    got_hits = np.array([a["hit_off"][i + 1] - a["hit_off"][i] for i in sel])
    assert (got_hits == np.diff(want["hit_off"])).all()


def test_pipelined_batch_equals_one_shot(oracle, pkg, gidx, golden_index, golden_genome, monkeypatch):
    """meme_seed_batch cuts large batches into sub-batches that overlap copies and kernels (MEME_SEED_CHUNK reads
    each): the assembled lists must equal the one-shot result and the oracle, including a ragged last sub-batch."""
    from bwa_meme_b200 import synth
    codes, offs = synth.make_reads(golden_genome, 5300, 150, seed=77, n_frac=0.05, len_jitter=60)
    n = len(offs) - 1
    monkeypatch.setenv("MEME_SEED_CHUNK", "100000000")
    one = gidx[3].seed(codes, offs)
    monkeypatch.setenv("MEME_SEED_CHUNK", "1024")
    for _ in range(2):                                   # second call reuses the pinned arrays and both buffer sets
        got = gidx[3].seed(codes, offs)
        for k in ("smem_off", "hit_off", "hits"):
            assert (got[k] == one[k]).all(), k
        assert (got["smems"] == one["smems"]).all()
    oix = oracle.OracleIndex(golden_index, 3)
    want = oix.seed(codes, offs)
    oix.close()
    assert _canon(oracle, got, n) == _canon(oracle, want, n)
    monkeypatch.setenv("MEME_SEED_CHUNK", "100000000")
    again = gidx[3].seed(codes[: offs[700]], offs[:701])   # back to the one-shot path on the same handle
    assert (again["hits"] == one["hits"][: one["hit_off"][700]]).all()


@pytest.mark.parametrize("mode", [3, 1])
def test_wide_index_kernels_on_golden_fixtures(oracle, pkg, golden_index, golden_reads, mid_index, mode, monkeypatch):
    """GRCh38-scale indexes (sa_num >= 2^32) use 5-byte positions and the 64-bit-index kernel instantiations.
    MEME_FORCE_WIDE runs exactly that code on the golden fixtures: results must equal the reference dumps / oracle."""
    from bwa_meme_b200 import capi, synth
    monkeypatch.setenv("MEME_FORCE_WIDE", "1")
    ix = capi.MemeIndex(golden_index, mode=mode)
    try:
        for rl in (150, 250):
            codes, offs = golden_reads[rl]
            n = len(offs) - 1
            want = oracle.parse_ref_dump(os.path.join(GOLDEN, f"smem_len{rl}.txt"), n)
            assert _canon(oracle, ix.seed(codes, offs), n) == want
    finally:
        ix.close()
    prefix, g = mid_index                                  # wide RMI windows + repeats: k-ary narrowing and galloping
    codes, offs = synth.make_reads(g, 1500, 320, seed=9, n_frac=0.05, len_jitter=300)
    n = len(offs) - 1
    ix = capi.MemeIndex(prefix, mode=mode)
    got = ix.seed(codes, offs)
    ix.close()
    oix = oracle.OracleIndex(prefix, mode)
    want = oix.seed(codes, offs)
    oix.close()
    assert _canon(oracle, got, n) == _canon(oracle, want, n)


@pytest.mark.parametrize("wide", ["0", "1"])
def test_regrouped_kernel_matches_reference_dump(oracle, pkg, golden_index, mid_index, wide):
    """The regrouped seeding kernel (MEME_SEED_IMPL=grp, meme_seed_grp.cuh) is selectable and must give the reference's
    results too.  The choice is latched per process, so it runs in its own interpreter."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    meta = __import__("json").load(open(os.path.join(GOLDEN, "index_sha.json")))
    mid_prefix, _ = mid_index
    code = f"""
import sys, os, numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
ge.load_pkg(); O = ge.load_oracle()
from bwa_meme_b200 import capi, synth
g = synth.make_genome({meta["genome_len"]}, seed={meta["genome_seed"]}, n_families=3, copies=12, rep_len=400)
for mode in (3, 1):
    ix = capi.MemeIndex({golden_index!r}, mode=mode)
    for rl, nr, nf, seed in {GOLDEN_READ_SETS!r}:
        codes, offs = synth.make_reads(g, nr, rl, seed=seed, n_frac=nf)
        n = len(offs) - 1
        got = ix.seed(codes, offs)
        want = O.parse_ref_dump(os.path.join({GOLDEN!r}, f"smem_len{{rl}}.txt"), n)
        assert O.canon_lines(got["smems"], got["smem_off"], got["hit_off"], got["hits"], n) == want, (mode, rl)
    ix.close()
g2 = synth.make_genome(400000, seed=77, n_families=6, copies=60, rep_len=800, divergence=0.01)
codes, offs = synth.make_reads(g2, 1500, 320, seed=9, n_frac=0.05, len_jitter=300)
n = len(offs) - 1
ix = capi.MemeIndex({mid_prefix!r}, mode=3); got = ix.seed(codes, offs); ix.close()
oix = O.OracleIndex({mid_prefix!r}, 3); want = oix.seed(codes, offs); oix.close()
c = lambda r: O.canon_lines(r["smems"], r["smem_off"], r["hit_off"], r["hits"], n)
assert c(got) == c(want)
print("OK")
"""
    e = dict(os.environ)
    e.update({"MEME_SEED_IMPL": "grp", "MEME_FORCE_WIDE": wide})
    if wide == "0":
        e.pop("MEME_FORCE_WIDE")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=900)
    assert p.returncode == 0 and "OK" in p.stdout, p.stderr[-2000:]


def test_rmi_inference_matches_reference_formula(oracle, pkg, golden_index, mid_index):
    """A4 pinned: the device's P-RMI prediction and error word equal learned_index_lookup
    (src/LearnedIndex_seeding.cpp:L186-210) for every key -- keys of real suffixes, random keys, the extremes.
    (Seeding results are canonical whatever the prediction, so only a direct comparison can see a wrong model.)"""
    from bwa_meme_b200 import capi
    rng = np.random.default_rng(3)
    for prefix in (golden_index, mid_index[0]):
        ix = capi.MemeIndex(prefix, mode=3)
        oix = oracle.OracleIndex(prefix, 3)
        text = np.fromfile(prefix + ".0123", dtype=np.uint8)
        starts = rng.integers(0, len(text) - 40, size=20000)
        k = np.zeros(len(starts), np.uint64)
        for r in range(32):                                   # first base in the most significant bits (Tokenization)
            k |= text[starts + r].astype(np.uint64) << np.uint64(62 - 2 * r)
        keys = np.concatenate([k, rng.integers(0, 1 << 63, size=20000, dtype=np.uint64) * np.uint64(2),
                               np.array([0, 1, (1 << 64) - 1, (1 << 63), (1 << 63) - 1], dtype=np.uint64)])
        gp, ge_ = ix.rmi_lookup(keys)
        wp, we = oix.rmi_lookup(keys)
        ix.close(); oix.close()
        assert (gp == wp).all() and (ge_ == we).all()


def test_probe_counts_stay_near_one_block_per_search(oracle, pkg, golden_index, golden_reads, monkeypatch):
    """A3/A4 guarded from the other side: with a well-trained index a search needs about one 8-entry block probe and
    hardly any k-ary narrowing round.  A wrong RMI key or prediction would not change the (canonical) results, but it
    would show here as extra probes."""
    from bwa_meme_b200 import capi
    monkeypatch.setenv("MEME_SEED_STATS", "1")
    codes, offs = golden_reads[150]
    n = len(offs) - 1
    ix = capi.MemeIndex(golden_index, mode=3)
    ix.seed(codes, offs, capi.default_seed_opt(rounds=1))      # round 1: every search starts from the RMI
    c = ix.seed_counters()
    ix.close()
    assert c["searches"] > 5 * n
    assert c["block_probes"] <= 1.15 * c["searches"], c
    assert c["kary_rounds"] <= 0.05 * c["searches"], c


def test_text_end_reads_match_reference_dump(oracle, pkg, textend_case):
    """Reads running into the end of the fwd+RC text (T-run end), across the forward/reverse junction and with T tails
    longer than the padding run: the CUDA kernel equals the reference's dump (src/LearnedIndex_seeding.cpp:L266-273)."""
    from bwa_meme_b200 import capi
    prefix, codes, offs = textend_case
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, "smem_textend.txt"), n)
    for mode in (3, 2, 1):
        ix = capi.MemeIndex(prefix, mode=mode)
        got = _canon(oracle, ix.seed(codes, offs), n)
        ix.close()
        bad = [i for i in range(n) if got[i] != want[i]]
        assert not bad, f"mode {mode}: {len(bad)} reads differ, first {bad[:5]}"
