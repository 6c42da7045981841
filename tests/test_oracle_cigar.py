"""The SAM-stage alignment oracle (oracle/cigar_oracle.c) against the reference's own mem_reg2aln (tests/golden/cigar_ref.json,
written by tests/golden/make_golden.py through oracle/ref_cigar_shim.cpp; re-run live where oracle/_ref has the shim), and the
product's per-region procedure (bwa-meme_b200/csrc/meme_cigar_core.cuh compiled with g++: tests/host_sim/cigar_host.cpp) against
both."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import chain_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CASES, SEED = 1500, 4            # must match tests/golden/make_golden.py: make_cigar


@pytest.fixture(scope="module")
def cigar_case(pkg, oracle):
    from bwa_meme_b200 import synth, index_build
    g = U.rep_genome(synth)
    codes, cases = U.make_cigar_cases(g, U.rep_meta()["contigs"], N_CASES, SEED)
    text = index_build.fwd_rc_text(g)
    tasks = U.cigar_tasks(cases, oracle.infer_w, oracle.CIGAR_TASK_DTYPE)
    ref = json.load(open(os.path.join(U.GOLDEN, "cigar_ref.json")))
    return g, text, codes, cases, tasks, ref


def check_against_reference(oracle, g, cases, raw, ref):
    assert len(raw) == len(ref) == len(cases)
    for i, (c, r, w) in enumerate(zip(cases, raw, ref)):
        s = oracle.sam_cigar(r, int(c["qb"]), int(c["qe"]), int(c["l_query"]), int(c["rb"]) >= len(g))
        assert (s["cigar"], s["md"], s["nm"]) == (w["cigar"], w["md"], w["nm"]), (i, c, s, w)


def test_oracle_matches_reference_mem_reg2aln(oracle, cigar_case):
    g, text, codes, cases, tasks, ref = cigar_case
    raw = oracle.cigar_regions(text, len(g), codes, tasks)
    check_against_reference(oracle, g, cases, raw, ref)
    # the cases cover the fast path, real matrices, long CIGARs and all three tries of the band loop
    assert sum(1 for r in raw if r["n_cigar"] == 1) > 50 and max(r["n_cigar"] for r in raw) > 20
    assert any("^" in r["md"] for r in raw)


def test_live_reference_agrees_with_committed_outputs(oracle, cigar_case, tmp_path):
    g, text, codes, cases, tasks, ref = cigar_case
    isa = "avx512bw" if " avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    shim = os.path.join(ROOT, "oracle", "_ref", f"cigar_shim_mode3.{isa}")
    if not os.path.exists(shim):
        pytest.skip("oracle/_ref/cigar_shim_* not built (needs the reference sources)")
    from bwa_meme_b200 import synth, index_build
    prefix = str(tmp_path / "rep.fa")
    m = U.rep_meta()
    index_build.write_bntseq(prefix, g)
    with open(prefix + ".ann", "w") as f:
        f.write(f"{len(g)} {len(m['contigs'])} 11\n")
        for name, off, ln, alt in m["contigs"]:
            f.write(f"0 {name} (null)\n{off} {ln} 0\n")
    open(prefix + ".amb", "w").write(f"{len(g)} {len(m['contigs'])} 0\n")
    live = U.run_cigar_shim(shim, prefix, codes, cases, str(tmp_path))
    assert [(x["cigar"], x["md"], x["nm"]) for x in live] == [(x["cigar"], x["md"], x["nm"]) for x in ref]


@pytest.fixture(scope="module")
def cigar_sim():
    build = os.path.join(ROOT, "tests", "host_sim", "_build")
    os.makedirs(build, exist_ok=True)
    if not shutil.which("g++"):
        pytest.skip("g++ is missing")
    lib = os.path.join(build, "libcigar_host.so")
    srcs = [os.path.join(ROOT, "tests", "host_sim", "cigar_host.cpp"), os.path.join(ROOT, "bwa-meme_b200", "csrc", "meme_cigar_core.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", lib, srcs[0]], check=True)
    L = C.CDLL(lib)
    L.cigar_sim.restype = C.c_int
    L.cigar_sim.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_int]
    return L


def run_sim(L, O, text, l_pac, codes, tasks, opt, stride=1, lane=0):
    n = len(tasks)
    recs = np.zeros(n, O.CIGAR_REC_DTYPE)
    cig = np.zeros(int((2 * tasks["qlen"].astype(np.int64) + 2).sum()) + 8, np.uint32)
    md = C.create_string_buffer(int((4 * (tasks["re"] - tasks["rb"]) + 16).sum()) + 64)
    cells = C.c_uint64()
    rc = L.cigar_sim(text.ctypes.data, l_pac, codes.ctypes.data, len(codes), tasks.ctypes.data, n, C.addressof(opt), recs.ctypes.data, cig.ctypes.data,
                     C.addressof(md), C.byref(cells), stride, lane)
    assert rc == 0, rc
    return O.unpack_cigar_result(recs, cig, md.raw), int(cells.value)


def test_kernel_source_on_host_matches_reference_and_oracle(oracle, cigar_case, cigar_sim):
    g, text, codes, cases, tasks, ref = cigar_case
    raw, cells = run_sim(cigar_sim, oracle, text, len(g), codes, tasks, oracle.default_cigar_opt())
    check_against_reference(oracle, g, cases, raw, ref)
    assert raw == oracle.cigar_regions(text, len(g), codes, tasks)
    assert cells > 0
    # the device layout: lane 5 of 32 interleaved regions (and the other lanes' bytes stay untouched)
    raw32, _ = run_sim(cigar_sim, oracle, text, len(g), codes, tasks[:400], oracle.default_cigar_opt(), stride=32, lane=5)
    assert raw32 == raw[:400]


def test_kernel_source_on_host_option_corners(oracle, cigar_case, cigar_sim):
    """Scoring and band options the reference outputs do not cover (they come from mem_opt_init defaults), rejected regions,
    starting bands beyond 4w: against the oracle."""
    g, text, codes, cases, tasks, ref = cigar_case
    for kw in ({"a": 2, "b": 3}, {"o_del": 1, "e_del": 3, "o_ins": 12, "e_ins": 2}, {"w": 5}, {"w": 1}, {"a": 1, "b": 1, "o_del": 0, "o_ins": 0}):
        opt = oracle.default_cigar_opt(**kw)
        raw, _ = run_sim(cigar_sim, oracle, text, len(g), codes, tasks, opt)
        assert raw == oracle.cigar_regions(text, len(g), codes, tasks, opt), kw
    odd = tasks[:8].copy()
    odd["w"][0] = 1000                                     # beyond 4 * opt.w: clamped by the loop
    odd["qlen"][1] = 0                                     # empty query: rejected
    odd["rb"][2], odd["re"][2] = len(g) - 20, len(g) + 20  # bridges the two strands: rejected
    odd["re"][3] = odd["rb"][3]                            # empty reference range: rejected
    odd["truesc"][4] = 10000                               # never reached: all three tries
    raw, _ = run_sim(cigar_sim, oracle, text, len(g), codes, odd, oracle.default_cigar_opt())
    assert raw == oracle.cigar_regions(text, len(g), codes, odd)
    assert [r["n_cigar"] for r in raw[1:4]] == [-1, -1, -1]
