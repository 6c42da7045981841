"""The mate-rescue Smith-Waterman oracle (oracle/matesw_oracle.c) against the reference's own
sort_classify + mem_sam_pe_batch (tests/golden/matesw_ref_*.npy, written by tests/golden/make_golden.py through
oracle/ref_matesw_shim.cpp; re-run live, in differently cut batches, where oracle/_ref has the AVX512 shim), and the product's
per-pair procedure (bwa-meme_b200/csrc/meme_matesw_core.cuh compiled with g++: tests/host_sim/matesw_host.cpp) against both."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import matesw_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETS = (("default", 1200, 11, {}), ("a2b3", 600, 12, {"a": 2, "b": 3}), ("gaps", 600, 13, {"o_del": 1, "e_del": 3, "o_ins": 12, "e_ins": 2}))   # = make_golden.MATESW_SETS


def cases(oracle, name):
    _, n, seed, kw = next(s for s in SETS if s[0] == name)
    recs, ref, qer = U.make_matesw_cases(n, seed, a=kw.get("a", 1), rec_dtype=oracle.MATESW_REC_DTYPE)
    return recs, ref, qer, oracle.default_matesw_opt(**kw), np.load(os.path.join(U.GOLDEN, f"matesw_ref_{name}.npy"))


@pytest.mark.parametrize("name", [s[0] for s in SETS])
def test_oracle_matches_reference_mem_sam_pe_batch(oracle, name):
    recs, ref, qer, opt, want = cases(oracle, name)
    got = oracle.matesw_batch(recs, ref, qer, opt)
    assert got.tolist() == want.tolist()
    if name == "default":                  # the cases cover both classes, second hits, found and missing start positions, empty hits
        byte = (recs["xtra"] & U.XBYTE) != 0
        assert byte.sum() > 200 and (~byte).sum() > 50
        assert (want["score2"] > 0).sum() > 50 and (want["tb"] >= 0).sum() > 200 and (want["tb"] < 0).sum() > 50
        assert (want["score"][~byte] > 255).any()


def test_live_reference_agrees_with_committed_outputs_in_any_batching(oracle, tmp_path):
    """A lane of the vector kernels does not depend on its neighbours: the reference gives the same kswr_t for a pair whether the
    pairs are run as one block or cut into blocks of 64, 37 or 1 (which changes every batch-wide loop bound of kswv512_*)."""
    if U.shim_path() is None:
        pytest.skip("oracle/_ref/matesw_shim_mode3.avx512bw not built, or no AVX512 here")
    recs, ref, qer, opt, want = cases(oracle, "default")
    for blk in (None, 64, 37, 1):
        got = U.run_shim(oracle, recs, ref, qer, opt, str(tmp_path), blk=blk)
        assert got.tolist() == want.tolist(), blk


@pytest.fixture(scope="module")
def sim(pkg):
    inc = "/usr/local/cuda/include"
    if not shutil.which("g++") or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are missing")
    sim_dir = os.path.join(ROOT, "tests", "host_sim")
    csrc = os.path.join(ROOT, "bwa-meme_b200", "csrc")
    os.makedirs(os.path.join(sim_dir, "_build"), exist_ok=True)
    lib = os.path.join(sim_dir, "_build", "libmatesw_host.so")
    deps = [os.path.join(sim_dir, "matesw_host.cpp"), os.path.join(csrc, "meme_matesw_core.cuh"), os.path.join(csrc, "meme_bsw_is.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-w", "-I", inc, "-o", lib, os.path.join(sim_dir, "matesw_host.cpp")], check=True)
    L = C.CDLL(lib)

    def run(oracle, recs, ref, qer, opt, stride=1, lane=0):
        out = np.zeros(len(recs), oracle.KSWR_DTYPE)
        cells = C.c_uint64()
        rc = L.matesw_sim(C.c_void_p(recs.ctypes.data), C.c_int64(len(recs)), C.c_void_p(ref.ctypes.data), C.c_void_p(qer.ctypes.data), C.byref(opt),
                          C.c_void_p(out.ctypes.data), C.byref(cells), C.c_int(stride), C.c_int(lane))
        assert rc == 0, rc
        return out, cells.value
    return run


@pytest.mark.parametrize("name", [s[0] for s in SETS])
def test_kernel_source_on_the_host_matches_reference_and_oracle(oracle, sim, name):
    recs, ref, qer, opt, want = cases(oracle, name)
    got, cells = sim(oracle, recs, ref, qer, opt)
    assert got.tolist() == want.tolist()
    got32, _ = sim(oracle, recs, ref, qer, opt, stride=32, lane=5)         # the device layout: 32 pairs interleaved
    assert got32.tolist() == want.tolist()
    assert cells > int((recs["len1"].astype(np.int64) * recs["len2"]).sum())           # phase 0 alone covers len1 x quanta(len2)


def test_kernel_source_edge_cases(oracle, sim):
    """Empty and one-base sequences, a window shorter than the read, zero gap-open penalties, a minimum score of 0."""
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 4, size=4000).astype(np.uint8)
    qer = np.concatenate([ref[100:250], rng.integers(0, 4, size=300).astype(np.uint8)])
    base = U.XSUBO | U.XSTART
    rows = [(0, 0, 0, 150, base | U.XBYTE | 19), (0, 0, 300, 0, base | U.XBYTE | 19), (100, 0, 1, 1, base | U.XBYTE | 0), (90, 0, 40, 150, base | U.XBYTE | 19),
            (0, 0, 900, 150, base | U.XBYTE | 0), (0, 150, 500, 260, base | 19), (0, 0, 900, 150, U.XBYTE | 19), (0, 0, 900, 150, base | 19),
            (3, 7, 5, 3, base | U.XBYTE | 1)]
    recs = np.array([(r, q, l1, l2, x, 0) for r, q, l1, l2, x in rows], dtype=oracle.MATESW_REC_DTYPE)
    for kw in ({}, {"o_del": 0, "o_ins": 0}, {"a": 3, "b": 1}):
        opt = oracle.default_matesw_opt(**kw)
        want = oracle.matesw_batch(recs, ref, qer, opt)
        got, _ = sim(oracle, recs, ref, qer, opt, stride=32, lane=31)
        assert got.tolist() == want.tolist(), kw
    if U.shim_path() is not None:
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            nz = recs[recs["len1"] > 0]                                             # the reference writes out of bounds for an empty window
            live = U.run_shim(oracle, nz, ref, qer, oracle.default_matesw_opt(), d)
            assert live.tolist() == oracle.matesw_batch(nz, ref, qer, oracle.default_matesw_opt()).tolist()
