"""The inter-sequence banded-SW pair function of the product (bwa-meme_b200/csrc/meme_bsw_is.cuh), compiled with g++
and run on the CPU against the oracle and the reference's golden vectors (test infrastructure; the `-m gpu` tests cover
the CUDA build of the same source)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_BSW_SETS, ROOT

SIM_DIR = os.path.join(ROOT, "tests", "host_sim")
CSRC = os.path.join(ROOT, "bwa-meme_b200", "csrc")


@pytest.fixture(scope="module")
def sim(pkg):
    inc = "/usr/local/cuda/include"
    if not shutil.which("g++") or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are missing")
    out_dir = os.path.join(SIM_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libbsw_host.so")
    deps = [os.path.join(SIM_DIR, "bsw_host.cpp"), os.path.join(CSRC, "meme_bsw_is.cuh"), os.path.join(CSRC, "meme_common.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I", inc, "-o", lib,
                        os.path.join(SIM_DIR, "bsw_host.cpp")], check=True)
    L = C.CDLL(lib)
    L.bsw_is_sim.restype = C.c_int
    from bwa_meme_b200 import capi

    def run(pairs, ref, qer, w, tries=1, prev_is_h0=0, opt=None):
        opt = opt or capi.default_bsw_opt()
        out = np.ascontiguousarray(pairs).copy().view(capi.SEQPAIR_DTYPE)
        wu = np.zeros(len(out), np.int32)
        cells = C.c_uint64()
        rc = L.bsw_is_sim(C.c_void_p(out.ctypes.data), C.c_int64(len(out)), C.c_void_p(ref.ctypes.data), C.c_void_p(qer.ctypes.data),
                          C.c_int(w), C.c_int(tries), C.c_int(prev_is_h0), C.byref(opt), C.c_void_p(wu.ctypes.data), C.byref(cells))
        assert rc == 0
        return out, wu, cells.value

    return run


@pytest.mark.parametrize("cfg", GOLDEN_BSW_SETS)
@pytest.mark.parametrize("w", [100, 200, 7])
def test_pair_function_matches_reference_vectors(oracle, sim, cfg, w):
    seed, n, mq, mx, h0m = cfg
    pairs, ref, qer = oracle.make_pairs(n, seed=seed, max_q=mq, max_extra_ref=mx, h0_max=h0m)
    want = np.load(os.path.join(GOLDEN, f"bsw_scalar_s{seed}_w{w}.npy"))
    got, _, _ = sim(pairs, ref, qer, w)
    res = np.stack([got[f] for f in oracle.RESULT_FIELDS], axis=1)
    assert (res == want).all()


@pytest.mark.parametrize("mq,mx,h0m,n", [(150, 110, 120, 6000), (31, 40, 30, 3000), (320, 200, 300, 1500), (1, 5, 5, 300)])
def test_pair_function_matches_oracle_and_counts_cells(oracle, sim, mq, mx, h0m, n):
    pairs, ref, qer = oracle.make_pairs(n, seed=mq + n, max_q=mq, max_extra_ref=mx, h0_max=h0m)
    for w in (100, 13):
        before = oracle.bsw_cells()
        want = oracle.bsw_batch(pairs, ref, qer, w)
        cpu_cells = oracle.bsw_cells() - before
        got, _, cells = sim(pairs, ref, qer, w)
        for f in oracle.RESULT_FIELDS:
            assert (got[f] == want[f]).all(), f
        assert cells == cpu_cells


def test_band_retry_rule(oracle, sim):
    """w, then 2w when the first result is not final (src/bwamem.cpp:L2968-3000): emulated with two oracle runs."""
    pairs, ref, qer = oracle.make_pairs(6000, seed=77, max_q=150, max_extra_ref=110, h0_max=120)
    for w, prev_is_h0 in ((3, 0), (3, 1), (100, 1)):
        a = oracle.bsw_batch(pairs, ref, qer, w)
        b = oracle.bsw_batch(pairs, ref, qer, 2 * w)
        prev = pairs["h0"] if prev_is_h0 else np.full(len(pairs), -1)
        final = (a["score"] == prev) | (a["max_off"] < (w >> 1) + (w >> 2))
        got, wu, _ = sim(pairs, ref, qer, w, tries=2, prev_is_h0=prev_is_h0)
        for f in oracle.RESULT_FIELDS:
            assert (got[f] == np.where(final, a[f], b[f])).all(), f
        assert (wu == np.where(final, w, 2 * w)).all()
        if w == 3:
            assert (~final).any() and final.any()


def test_non_default_penalties(oracle, sim, pkg):
    from bwa_meme_b200 import capi
    pairs, ref, qer = oracle.make_pairs(3000, seed=12)
    want = oracle.bsw_batch(pairs, ref, qer, 60, oracle.BswOpt(5, 2, 7, 3, 50, 3, 2, 5))
    got, _, _ = sim(pairs, ref, qer, 60, opt=capi.BswOpt(5, 2, 7, 3, 50, 3, 2, 5))
    for f in oracle.RESULT_FIELDS:
        assert (got[f] == want[f]).all(), f


def test_reference_by_position_equals_reference_bytes(oracle, sim, pkg):
    """The pair function reading its reference side from a 2-bit text by position (forwards for right extensions,
    backwards for left ones) gives what the byte form gives."""
    import ctypes as C
    from bwa_meme_b200 import capi
    rng = np.random.default_rng(9)
    text = rng.integers(0, 4, size=60000, dtype=np.uint8)
    n = 3000
    pairs = np.zeros(n, capi.SEQPAIR_DTYPE)
    pos = np.zeros(n, np.int64); dirs = np.zeros(n, np.int32)
    refs, qers = [], []
    ro = qo = 0
    for i in range(n):
        ql, tl = int(rng.integers(1, 151)), 0
        tl = ql + int(rng.integers(0, 100))
        d = 1 if rng.random() < 0.5 else -1
        p = int(rng.integers(400, len(text) - 400))
        t = text[p:p + tl] if d > 0 else text[p - tl + 1:p + 1][::-1]
        q = t[:ql].copy()
        m = rng.random(ql) < 0.04
        q[m] = (q[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        pairs[i] = (ro, qo, 0, tl, ql, int(rng.integers(19, 120)), 0, 0, 0, 0, 0, 0, 0, 0)
        pos[i], dirs[i] = p, d
        refs.append(t); qers.append(q); ro += tl; qo += ql
    ref, qer = np.concatenate(refs), np.concatenate(qers)
    want = oracle.bsw_batch(pairs, ref, qer, 100)
    L = C.CDLL(os.path.join(SIM_DIR, "_build", "libbsw_host.so"))
    out = pairs.copy()
    opt = capi.default_bsw_opt()
    rc = L.bsw_is_sim_text(C.c_void_p(out.ctypes.data), C.c_int64(n), C.c_void_p(text.ctypes.data), C.c_int64(len(text)),
                           C.c_void_p(pos.ctypes.data), C.c_void_p(dirs.ctypes.data), C.c_void_p(qer.ctypes.data), C.c_int(100),
                           C.c_int(1), C.c_int(0), C.byref(opt), None)
    assert rc == 0
    for f in oracle.RESULT_FIELDS:
        assert (out[f] == want[f]).all(), f
