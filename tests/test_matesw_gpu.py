"""Mate-rescue Smith-Waterman on the GPU (meme_matesw_batch) through the C ABI against the reference's
own sort_classify + mem_sam_pe_batch outputs (tests/golden/matesw_ref_*.npy) and the oracle."""
import os

import numpy as np
import pytest

import matesw_util as U

pytestmark = pytest.mark.gpu
SETS = (("default", 1200, 11, {}), ("a2b3", 600, 12, {"a": 2, "b": 3}), ("gaps", 600, 13, {"o_del": 1, "e_del": 3, "o_ins": 12, "e_ins": 2}))


@pytest.fixture(scope="module")
def ix(pkg, tmp_path_factory):
    from bwa_meme_b200 import synth, index_build, capi
    prefix = str(tmp_path_factory.mktemp("msw_index") / "g.fa")
    index_build.build_index(prefix, synth.make_genome(30000, seed=2), bits=14)
    h = capi.MemeIndex(prefix, mode=3)
    yield h
    h.close()


def to_segments(capi, recs, ref, qer, cuts):
    """The pairs as several host threads would hand them over: contiguous shares, each with its own buffers and offsets."""
    segs, bounds = [], [0] + list(cuts) + [len(recs)]
    for b, e in zip(bounds[:-1], bounds[1:]):
        p = np.zeros(e - b, capi.SEQPAIR_DTYPE)
        r0 = int(recs["r_off"][b]) if e > b else 0
        q0 = int(recs["q_off"][b]) if e > b else 0
        r1 = int(recs["r_off"][e - 1] + recs["len1"][e - 1]) if e > b else 0
        q1 = int(recs["q_off"][e - 1] + recs["len2"][e - 1]) if e > b else 0
        p["idr"] = recs["r_off"][b:e] - r0; p["idq"] = recs["q_off"][b:e] - q0
        p["len1"] = recs["len1"][b:e]; p["len2"] = recs["len2"][b:e]; p["h0"] = recs["xtra"][b:e]
        p["regid"] = np.arange(e - b)
        segs.append((p, ref[r0:r1].copy(), qer[q0:q1].copy()))
    return segs


@pytest.mark.parametrize("name", [s[0] for s in SETS])
def test_cuda_matches_reference_mem_sam_pe_batch(pkg, oracle, ix, name):
    from bwa_meme_b200 import capi
    _, n, seed, kw = next(s for s in SETS if s[0] == name)
    recs, ref, qer = U.make_matesw_cases(n, seed, a=kw.get("a", 1), rec_dtype=oracle.MATESW_REC_DTYPE)
    want = np.load(os.path.join(U.GOLDEN, f"matesw_ref_{name}.npy"))
    for cuts in ((), (1, 400, 401)):
        out = ix.matesw_batch(to_segments(capi, recs, ref, qer, cuts), capi.default_matesw_opt(**kw))
        got = np.concatenate(out["aln"])
        assert got.tolist() == want.tolist(), cuts
    assert out["cells"] > 0 and out["ms"] > 0


def test_cuda_matches_oracle_many_pairs_edges_and_limits(pkg, oracle, ix):
    """More pairs than one launch slice (2^17), warm-up in between, empty batch / segments, and inputs beyond the limits."""
    from bwa_meme_b200 import capi
    recs, ref, qer = U.make_matesw_cases(3000, 21, rec_dtype=oracle.MATESW_REC_DTYPE)
    want = oracle.matesw_batch(recs, ref, qer)
    reps = (1 << 17) // len(recs) + 2
    segs = to_segments(capi, recs, ref, qer, (1000,))
    ix.matesw_warmup(5000)
    out = ix.matesw_batch(segs * reps)
    for k in (0, 1, reps, 2 * reps - 1):
        assert np.concatenate(out["aln"][2 * (k // 2):2 * (k // 2) + 2]).tolist() == want.tolist()
    out = ix.matesw_batch([])
    assert out["aln"] == []
    empty = (np.zeros(0, capi.SEQPAIR_DTYPE), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    out = ix.matesw_batch([empty, segs[0], empty])
    assert out["aln"][1].tolist() == want[:1000].tolist() and len(out["aln"][0]) == 0
    bad = segs[0][0][:4].copy()
    bad["len2"][1] = 600
    with pytest.raises(RuntimeError, match="limits|outside"):
        ix.matesw_batch([(bad, segs[0][1], segs[0][2])])
    with pytest.raises(RuntimeError, match="range"):
        ix.matesw_batch([segs[0]], capi.default_matesw_opt(e_del=0))


def test_cuda_matches_oracle_at_the_limits(pkg, oracle, ix):
    """The longest mates the kernel takes (496 bases in the 16-bit class, 255 in the 8-bit class with a match score that keeps
    KSW_XBYTE legal), windows of several thousand bases, and a poly-A window against a poly-A mate (every row maximum equal)."""
    from bwa_meme_b200 import capi
    rng = np.random.default_rng(31)
    base = U.XSUBO | U.XSTART
    ref_parts, qer_parts, rows = [], [], []
    ro = qo = 0

    def add(win, q, xtra):
        nonlocal ro, qo
        rows.append((ro, qo, len(win), len(q), xtra, 0))
        ref_parts.append(win); qer_parts.append(q)
        ro += len(win); qo += len(q)

    for l1, l2, xtra in ((5000, 496, base | 19), (3000, 255, base | U.XBYTE | 19), (31999, 150, base | U.XBYTE | 19), (700, 249, base | U.XBYTE | 19),
                         (2000, 400, base | 30), (640, 150, base | U.XBYTE | 19)):
        win = rng.integers(0, 4, size=l1).astype(np.uint8)
        st = int(rng.integers(0, l1 - l2))
        q = win[st:st + l2].copy()
        mm = rng.random(l2) < 0.03
        q[mm] = (q[mm] + 1) & 3
        add(win, q, xtra)
    add(np.zeros(600, np.uint8), np.zeros(150, np.uint8), base | U.XBYTE | 19)
    add(np.full(600, 3, np.uint8), np.zeros(150, np.uint8), base | U.XBYTE | 19)
    recs = np.array(rows, dtype=oracle.MATESW_REC_DTYPE)
    ref, qer = np.concatenate(ref_parts), np.concatenate(qer_parts)
    want = oracle.matesw_batch(recs, ref, qer)
    out = ix.matesw_batch(to_segments(capi, recs, ref, qer, (3,)))
    assert np.concatenate(out["aln"]).tolist() == want.tolist()
    assert want["score"][0] > 255 and want["tb"][0] >= 0
