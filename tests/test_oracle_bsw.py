"""The banded-SW oracle against BandedPairWiseSW::scalarBandedSWA outputs (golden vectors) -- CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_BSW_SETS, ROOT


@pytest.mark.parametrize("cfg", GOLDEN_BSW_SETS)
@pytest.mark.parametrize("w", [100, 200, 7])
def test_bsw_oracle_matches_reference(oracle, cfg, w):
    seed, n, mq, mx, h0m = cfg
    pairs, ref, qer = oracle.make_pairs(n, seed=seed, max_q=mq, max_extra_ref=mx, h0_max=h0m)
    want = np.load(os.path.join(GOLDEN, f"bsw_scalar_s{seed}_w{w}.npy"))
    got = oracle.bsw_batch(pairs, ref, qer, w)
    res = np.stack([got[f] for f in oracle.RESULT_FIELDS], axis=1)
    assert (res == want).all(), f"{int((res != want).any(axis=1).sum())} pairs differ"


def test_bsw_oracle_simple_cases(oracle):
    """perfect match extends to the end with score h0 + qlen*a; empty band / N handling."""
    q = np.array([0, 1, 2, 3] * 10, np.uint8)
    t = np.concatenate([q, np.array([0, 0, 0], np.uint8)])
    pairs = np.zeros(3, oracle.SEQPAIR_DTYPE)
    pairs["len1"] = [len(t), len(t), 5]
    pairs["len2"] = [len(q), len(q), len(q)]
    pairs["h0"] = [10, 1, 10]
    pairs["idr"] = [0, 0, 0]
    out = oracle.bsw_batch(pairs, t, q, 100)
    assert out["score"][0] == 10 + 40 and out["qle"][0] == 40 and out["tle"][0] == 40
    assert out["gscore"][0] == 50 and out["gtle"][0] == 40
    assert out["score"][2] == 15 and out["tle"][2] == 5 and out["gscore"][2] == 0


@pytest.mark.skipif(not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")), reason="oracle/_ref not built")
def test_bsw_oracle_vs_live_reference(oracle):
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    shim = os.path.join(ROOT, "oracle", "_ref", f"bsw_shim_mode3.{isa}")
    if not os.path.exists(shim):
        pytest.skip("shim missing")
    pairs, ref, qer = oracle.make_pairs(1500, seed=77, max_q=200, max_extra_ref=150, h0_max=180)
    for w in (100, 33):
        want = oracle.run_ref_bsw_shim(shim, pairs, ref, qer, w, "scalar")
        got = oracle.bsw_batch(pairs, ref, qer, w)
        for f in oracle.RESULT_FIELDS:
            assert (got[f] == want[f]).all(), f
