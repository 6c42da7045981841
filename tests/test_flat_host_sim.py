"""The flat seeding kernel's own source (bwa-meme_b200/csrc/meme_seed_flat.cuh), compiled with g++ through
tests/host_sim/cuda_shim.h and run ONE LANE AT A TIME on the CPU, against the reference dumps and the oracle.

This is test infrastructure, not a CPU path of the product: it exists so that a change to the kernel's control logic
(state machine, search engine, task kinds) can be checked without a GPU.  What it cannot see is anything warp-level
(reconvergence, cp.async ordering, shared-memory layout across lanes) -- the `-m gpu` tests cover that.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_READ_SETS, ROOT

SIM_DIR = os.path.join(ROOT, "tests", "host_sim")
CSRC = os.path.join(ROOT, "bwa-meme_b200", "csrc")


def _cuda_include():
    for base in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if base and os.path.exists(os.path.join(base, "include", "cuda_runtime.h")):
            return os.path.join(base, "include")
    return None


def make_sim(which: str):
    """Build (if stale) and bind the host simulation of one kernel: "flat" or "grp"."""
    inc = _cuda_include()
    if not shutil.which("g++") or not inc:
        pytest.skip("g++ or the CUDA headers are missing")
    out_dir = os.path.join(SIM_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libflat_host.so")
    deps = [os.path.join(SIM_DIR, f) for f in ("flat_host.cpp", "cuda_shim.h")] + \
           [os.path.join(CSRC, f) for f in ("meme_seed_flat.cuh", "meme_seed_grp.cuh", "meme_seed_task.cuh", "meme_common.cuh")] + \
           [os.path.join(ROOT, "include", "meme_b200.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", "-I", inc, "-o", lib,
                        os.path.join(SIM_DIR, "flat_host.cpp")], check=True)
    L = C.CDLL(lib)
    L.flat_sim_seed.restype = C.c_int
    impl = 1 if which == "grp" else 0
    from bwa_meme_b200 import capi

    def seed(prefix, mode, codes, offs, opt=None, wide=False):
        opt = opt or capi.default_seed_opt()
        f = lambda ext: np.fromfile(prefix + ext, dtype=np.uint8)
        text, pos = f(".0123"), f(".pos_packed")
        l1, l2 = f(".suffixarray_uint64_L1_PARAMETERS"), f(".suffixarray_uint64_L2_PARAMETERS")
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        res = capi.SeedResult()
        counters = (C.c_uint64 * 8)()
        vp = lambda a: C.c_void_p(a.ctypes.data)
        L.flat_sim_set_impl(C.c_int(impl))
        rc = L.flat_sim_seed(vp(text), C.c_uint64(len(text)), vp(pos), vp(l1), C.c_uint64(len(l1) // 24), vp(l2),
                             C.c_uint64(len(l2) // 24), C.c_int(mode), C.c_int(int(wide)), vp(codes), vp(offs),
                             C.c_int64(len(offs) - 1), C.byref(opt), C.byref(res), counters)
        assert rc == 0, rc
        r = capi.MemeIndex._wrap(res)
        r["counters"] = dict(zip(("searches", "blocks", "kary", "ext", "pac", "round_trips", "t2_tasks", "pool"), counters))
        return r

    return seed


@pytest.fixture(scope="module", params=["grp", "flat"])
def sim(pkg, request):
    return make_sim(request.param)


def _canon(O, r, n):
    return O.canon_lines(r["smems"], r["smem_off"], r["hit_off"], r["hits"], n)


@pytest.mark.parametrize("rl", [s[0] for s in GOLDEN_READ_SETS])
@pytest.mark.parametrize("mode", [3, 2, 1])
def test_kernel_source_on_cpu_matches_reference_dump(oracle, sim, golden_index, golden_reads, rl, mode):
    codes, offs = golden_reads[rl]
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, f"smem_len{rl}.txt"), n)
    got = _canon(oracle, sim(golden_index, mode, codes, offs), n)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, f"{len(bad)} reads differ, first {bad[:5]}"


def test_wide_instantiation_on_cpu(oracle, sim, golden_index, golden_reads):
    codes, offs = golden_reads[150]
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, "smem_len150.txt"), n)
    assert _canon(oracle, sim(golden_index, 3, codes, offs, wide=True), n) == want


@pytest.mark.parametrize("rounds", [1, 2, 3])
def test_rounds_on_cpu(oracle, sim, pkg, golden_index, golden_reads, rounds):
    from bwa_meme_b200 import capi
    codes, offs = golden_reads[150]
    n = len(offs) - 1
    oix = oracle.OracleIndex(golden_index, 3)
    want = oix.seed(codes, offs, oracle.default_opt(rounds=rounds))
    oix.close()
    got = sim(golden_index, 3, codes, offs, capi.default_seed_opt(rounds=rounds))
    assert _canon(oracle, got, n) == _canon(oracle, want, n)
    assert (got["hits"] == want["hits"]).all() and (got["smem_off"] == want["smem_off"]).all()


@pytest.mark.parametrize("mode", [3, 1])
def test_wide_windows_and_repeats_on_cpu(oracle, sim, pkg, mid_index, mode):
    """2^12 leaves (k-ary narrowing), 60-copy repeat families (galloping), ragged read lengths 20..320."""
    from bwa_meme_b200 import synth
    prefix, g = mid_index
    codes, offs = synth.make_reads(g, 1200, 320, seed=5, n_frac=0.05, len_jitter=300)
    n = len(offs) - 1
    got = sim(prefix, mode, codes, offs)
    oix = oracle.OracleIndex(prefix, mode)
    want = oix.seed(codes, offs)
    oix.close()
    a, b = _canon(oracle, got, n), _canon(oracle, want, n)
    bad = [i for i in range(n) if a[i] != b[i]]
    assert not bad, f"mode {mode}: {len(bad)} reads differ, first {bad[:5]}"
    assert got["counters"]["kary"] > 0 and got["counters"]["ext"] > 0, got["counters"]


def test_max_occ_sampling_on_cpu(sim, pkg, mid_index):
    """max_occ > 0 returns exactly the hits mem_chain_Learned would read (bwamem.cpp:L1154-1155)."""
    from bwa_meme_b200 import capi, synth
    prefix, g = mid_index
    codes, offs = synth.make_reads(g, 300, 150, seed=9)
    n = len(offs) - 1
    occ = 7
    full = sim(prefix, 3, codes, offs)
    samp = sim(prefix, 3, codes, offs, capi.default_seed_opt(max_occ=occ))
    assert (full["smems"]["hitcount"] == samp["smems"]["hitcount"]).all()
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


def test_text_end_reads_on_cpu(oracle, sim, textend_case):
    """The kernels' control code on reads that run into the end of the text (tests/golden/smem_textend.txt)."""
    prefix, codes, offs = textend_case
    n = len(offs) - 1
    want = oracle.parse_ref_dump(os.path.join(GOLDEN, "smem_textend.txt"), n)
    for mode in (3, 1):
        got = _canon(oracle, sim(prefix, mode, codes, offs), n)
        bad = [i for i in range(n) if got[i] != want[i]]
        assert not bad, f"mode {mode}: {len(bad)} reads differ, first {bad[:5]}"
