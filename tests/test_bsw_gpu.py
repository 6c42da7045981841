"""Parity of the CUDA banded-SW kernel (through the C ABI) with the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_BSW_SETS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bsw(pkg):
    from bwa_meme_b200 import capi
    b = capi.MemeBsw()
    yield b
    b.close()


@pytest.mark.parametrize("cfg", GOLDEN_BSW_SETS)
@pytest.mark.parametrize("w", [100, 200, 7])
def test_cuda_matches_reference_vectors(oracle, bsw, cfg, w):
    seed, n, mq, mx, h0m = cfg
    pairs, ref, qer = oracle.make_pairs(n, seed=seed, max_q=mq, max_extra_ref=mx, h0_max=h0m)
    want = np.load(os.path.join(GOLDEN, f"bsw_scalar_s{seed}_w{w}.npy"))
    got = bsw.get_scores(pairs, ref, qer, w)
    res = np.stack([got[f] for f in oracle.RESULT_FIELDS], axis=1)
    bad = np.nonzero((res != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} pairs differ; first {bad[:3]}: {res[bad[0]]} vs {want[bad[0]]} len {pairs[bad[0]]}"


@pytest.mark.parametrize("mq,mx,h0m,n", [(150, 110, 120, 20000), (31, 40, 30, 5000), (64, 10, 200, 5000),
                                         (320, 200, 300, 3000), (1, 5, 5, 500)])
def test_cuda_matches_oracle_random(oracle, bsw, mq, mx, h0m, n):
    pairs, ref, qer = oracle.make_pairs(n, seed=mq + n, max_q=mq, max_extra_ref=mx, h0_max=h0m)
    for w in (100, 13):
        want = oracle.bsw_batch(pairs, ref, qer, w)
        got = bsw.get_scores(pairs, ref, qer, w)
        for f in oracle.RESULT_FIELDS:
            bad = np.nonzero(got[f] != want[f])[0]
            assert len(bad) == 0, f"{f}: {len(bad)} differ, first {bad[:3]} got {got[f][bad[:3]]} want {want[f][bad[:3]]} {pairs[bad[0]]}"


def test_long_queries_take_generic_path(oracle, bsw):
    pairs, ref, qer = oracle.make_pairs(40, seed=4, max_q=900, max_extra_ref=300, h0_max=400)
    assert (pairs["len2"] > 383).any()
    want = oracle.bsw_batch(pairs, ref, qer, 100)
    got = bsw.get_scores(pairs, ref, qer, 100)
    for f in oracle.RESULT_FIELDS:
        assert (got[f] == want[f]).all(), f


def test_non_default_penalties(oracle, pkg):
    from bwa_meme_b200 import capi
    opt = capi.BswOpt(5, 2, 7, 3, 50, 3, 2, 5)
    oopt = oracle.BswOpt(5, 2, 7, 3, 50, 3, 2, 5)
    b = capi.MemeBsw(opt)
    pairs, ref, qer = oracle.make_pairs(4000, seed=12)
    want = oracle.bsw_batch(pairs, ref, qer, 60, oopt)
    got = b.get_scores(pairs, ref, qer, 60)
    b.close()
    for f in oracle.RESULT_FIELDS:
        assert (got[f] == want[f]).all(), f


def test_empty_and_bad_input(pkg, bsw, oracle):
    from bwa_meme_b200 import capi
    out = bsw.get_scores(np.zeros(0, capi.SEQPAIR_DTYPE), np.zeros(0, np.uint8), np.zeros(0, np.uint8), 100)
    assert len(out) == 0
    pairs, ref, qer = oracle.make_pairs(4, seed=1)
    pairs["idr"][2] = len(ref)          # points past the buffer
    with pytest.raises(capi.MemeError):
        bsw.get_scores(pairs, ref, qer, 100)


def test_cells_counter_and_determinism(oracle, bsw):
    pairs, ref, qer = oracle.make_pairs(3000, seed=21)
    before = oracle.bsw_cells()
    want = oracle.bsw_batch(pairs, ref, qer, 100)
    cpu_cells = oracle.bsw_cells() - before
    a = bsw.get_scores(pairs, ref, qer, 100)
    assert bsw.cells() == cpu_cells
    b = bsw.get_scores(pairs, ref, qer, 100)
    assert (a == b).all() and (a["score"] == want["score"]).all()


def test_band_retry_on_device(oracle, bsw):
    """w, then 2w when the first result is not final (src/bwamem.cpp:L2968-3000), decided by the kernel: emulated here
    with two oracle runs.  Covers the inter-sequence kernel and (long queries, huge h0) the overflow classes."""
    for mq, mx, h0m, n in ((150, 110, 120, 20000), (700, 200, 300, 300), (100, 60, 70000, 500)):
        pairs, ref, qer = oracle.make_pairs(n, seed=77 + mq, max_q=mq, max_extra_ref=mx, h0_max=h0m)
        for w, prev_is_h0 in ((3, False), (3, True), (100, True)):
            a = oracle.bsw_batch(pairs, ref, qer, w)
            b = oracle.bsw_batch(pairs, ref, qer, 2 * w)
            prev = pairs["h0"] if prev_is_h0 else np.full(len(pairs), -1)
            final = (a["score"] == prev) | (a["max_off"] < (w >> 1) + (w >> 2))
            got, wu = bsw.get_scores_ext(pairs, ref, qer, w, tries=2, prev_is_h0=prev_is_h0)
            for f in oracle.RESULT_FIELDS:
                bad = np.nonzero(got[f] != np.where(final, a[f], b[f]))[0]
                assert len(bad) == 0, (mq, w, prev_is_h0, f, bad[:3], pairs[bad[0]])
            assert (wu == np.where(final, w, 2 * w)).all()


def test_sorted_launch_order_covers_every_pair(oracle, bsw):
    """Pairs are processed in a device-sorted order and in per-length classes: every pair must be written exactly once,
    whatever its position (ragged lengths, empty queries, empty references)."""
    pairs, ref, qer = oracle.make_pairs(50000, seed=5, max_q=318, max_extra_ref=60, h0_max=100)
    pairs = pairs.copy()
    pairs["len2"][::97] = 0
    pairs["len1"][::89] = 0
    pairs["score"] = -12345
    want = oracle.bsw_batch(pairs, ref, qer, 100)
    got = bsw.get_scores(pairs, ref, qer, 100)
    for f in oracle.RESULT_FIELDS:
        assert (got[f] == want[f]).all(), f


def test_reference_by_position_in_the_resident_text(oracle, pkg, golden_index):
    """Pairs whose reference side is a position in the index's HBM-resident 2-bit text (no reference upload;
    meme_bsw_attach_text): forwards and backwards, inter-sequence kernel and overflow classes (a long query, a huge h0),
    with the band retries.  Equal to the oracle on the bytes those positions spell."""
    from bwa_meme_b200 import capi
    rng = np.random.default_rng(19)
    text = np.fromfile(golden_index + ".0123", dtype=np.uint8)
    n = 6000
    pairs = np.zeros(n, capi.SEQPAIR_DTYPE)
    byte_pairs = np.zeros(n, capi.SEQPAIR_DTYPE)
    refs, qers = [], []
    ro = qo = 0
    for i in range(n):
        ql = int(rng.integers(1, 151)) if i % 500 else int(rng.integers(330, 420))
        tl = ql + int(rng.integers(0, 100))
        d = 1 if rng.random() < 0.5 else -1
        p = int(rng.integers(600, len(text) - 600))
        t = text[p:p + tl] if d > 0 else text[p - tl + 1:p + 1][::-1]
        q = t[:ql].copy()
        m = rng.random(ql) < 0.04
        q[m] = (q[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        h0 = int(rng.integers(19, 120)) if i % 333 else 70000
        pairs[i] = (p & 0xffffffff, qo, (p >> 32) | (-(1 << 31) if d < 0 else 0), tl, ql, h0, 0, 0, 0, 0, 0, 0, 0, 0)
        byte_pairs[i] = (ro, qo, 0, tl, ql, h0, 0, 0, 0, 0, 0, 0, 0, 0)
        refs.append(t); qers.append(q); ro += tl; qo += ql
    ref, qer = np.concatenate(refs), np.concatenate(qers)
    ix = capi.MemeIndex(golden_index, mode=3)
    b = capi.MemeBsw()
    try:
        for w, tries in ((100, 1), (3, 2)):
            a = oracle.bsw_batch(byte_pairs, ref, qer, w)
            if tries == 2:
                a2 = oracle.bsw_batch(byte_pairs, ref, qer, 2 * w)
                final = (a["score"] == -1) | (a["max_off"] < (w >> 1) + (w >> 2))
            got = b.get_scores_text(ix, pairs, qer, w, tries=tries)
            for f in oracle.RESULT_FIELDS:
                want = a[f] if tries == 1 else np.where(final, a[f], a2[f])
                bad = np.nonzero(got[f] != want)[0]
                assert len(bad) == 0, (w, tries, f, bad[:3], pairs[bad[0]])
    finally:
        b.close(); ix.close()
