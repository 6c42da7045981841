"""The chaining oracle (oracle/chain_oracle.c) against chain dumps of the reference's own mem_chain_Learned + mem_chain_flt
(tests/golden/chains_*.txt, written by tests/golden/make_golden.py from the patched aligner with every hook off), and the
product's per-read chaining procedure (bwa-meme_b200/csrc/meme_chain_core.cuh compiled with g++: tests/host_sim/chain_host.cpp)
against both."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from chain_util import rep_genome, rep_reads, rep_contigs, sample_hits, read_dump, rep_meta
from conftest import GOLDEN_READ_SETS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [("chains_rep.txt", {}), ("chains_rep_c8_W40_N3.txt", {"max_occ": 8, "min_chain_weight": 40, "max_chain_extend": 3})]


@pytest.fixture(scope="module")
def rep_case(pkg, oracle, tmp_path_factory):
    from bwa_meme_b200 import synth, index_build
    g = rep_genome(synth)
    prefix = str(tmp_path_factory.mktemp("rep_index") / "rep.fa")
    index_build.build_index(prefix, g, bits=rep_meta()["bits"])
    codes, offs = rep_reads(synth, g)
    ix = oracle.OracleIndex(prefix, 3)
    res = ix.seed(codes, offs)
    ix.close()
    return g, codes, offs, res


@pytest.fixture(scope="module")
def chain_sim():
    build = os.path.join(ROOT, "tests", "host_sim", "_build")
    os.makedirs(build, exist_ok=True)
    if not shutil.which("g++"):
        pytest.skip("g++ is missing")
    lib = os.path.join(build, "libchain_host.so")
    srcs = [os.path.join(ROOT, "tests", "host_sim", "chain_host.cpp"), os.path.join(ROOT, "bwa-meme_b200", "csrc", "meme_chain_core.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", lib, srcs[0]], check=True)
    L = C.CDLL(lib)
    L.chain_sim.restype = C.c_int
    L.chain_sim.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p] + [C.c_void_p] * 4
    return L


def run_sim(L, O, sampled, read_len, contigs, l_pac, opt):
    n = len(read_len)
    nh = int(sampled["hit_off"][-1])
    chain_off = np.zeros(n + 1, np.int64)
    frac = np.zeros(max(n, 1), np.float32)
    chains = np.zeros(nh + 1, O.CHAIN_DTYPE)
    seeds = np.zeros(nh + 1, O.CHAIN_SEED_DTYPE)
    sm = np.ascontiguousarray(sampled["smems"])
    rl = np.ascontiguousarray(read_len, dtype=np.int32)
    ct = np.ascontiguousarray(contigs)
    rc = L.chain_sim(sm.ctypes.data, sampled["smem_off"].ctypes.data, sampled["hits"].ctypes.data, sampled["hit_off"].ctypes.data, rl.ctypes.data,
                     n, ct.ctypes.data, len(ct), l_pac, C.addressof(opt), chain_off.ctypes.data, frac.ctypes.data, chains.ctypes.data,
                     seeds.ctypes.data)
    assert rc == 0
    nc = int(chain_off[-1])
    ns = int(chains["seed_beg"][nc - 1] + chains["n"][nc - 1]) if nc else 0
    return {"chain_off": chain_off, "frac_rep": frac[:n], "chains": chains[:nc], "seeds": seeds[:ns]}


@pytest.mark.parametrize("dump,kw", VARIANTS)
def test_oracle_matches_reference_dump_repeats(oracle, rep_case, dump, kw):
    g, codes, offs, res = rep_case
    opt = oracle.default_chain_opt(**kw)
    out = oracle.chain_reads(res, np.diff(offs), rep_contigs(oracle.CONTIG_DTYPE), len(g), opt)
    lines = oracle.chain_lines(out, names=[f"r{i}" for i in range(len(offs) - 1)])
    ref = read_dump(dump)
    assert len(lines) == len(ref)
    bad = [i for i, (a, b) in enumerate(zip(lines, ref)) if a != b]
    assert not bad, (bad[:5], lines[bad[0]][:300], ref[bad[0]][:300])
    # the fixture does exercise what makes chaining order-sensitive
    if not kw:
        n_per_read = np.diff(out["chain_off"])
        assert n_per_read.max() > 16                  # beyond one B-tree node and beyond the insertion-sort cut-off
        assert (out["chains"]["n"] > 1).any() and (out["chains"]["is_alt"] == 1).any() and (out["chains"]["kept"] == 1).any()


@pytest.mark.parametrize("rl", [rs[0] for rs in GOLDEN_READ_SETS])
def test_oracle_matches_reference_dump_golden(oracle, golden_index, golden_genome, golden_reads, rl):
    codes, offs = golden_reads[rl]
    ix = oracle.OracleIndex(golden_index, 3)
    res = ix.seed(codes, offs)
    ix.close()
    contigs = np.array([(0, len(golden_genome), 0)], dtype=oracle.CONTIG_DTYPE)
    out = oracle.chain_reads(res, np.diff(offs), contigs, len(golden_genome))
    lines = oracle.chain_lines(out, names=[f"r{i}" for i in range(len(offs) - 1)])
    assert lines == read_dump(f"chains_len{rl}.txt")


@pytest.mark.parametrize("dump,kw", VARIANTS)
def test_kernel_source_on_host_matches_reference_dump(oracle, rep_case, chain_sim, dump, kw):
    g, codes, offs, res = rep_case
    opt = oracle.default_chain_opt(**kw)
    out = run_sim(chain_sim, oracle, sample_hits(res, opt.max_occ), np.diff(offs), rep_contigs(oracle.CONTIG_DTYPE), len(g), opt)
    lines = oracle.chain_lines(out, names=[f"r{i}" for i in range(len(offs) - 1)])
    ref = read_dump(dump)
    bad = [i for i, (a, b) in enumerate(zip(lines, ref)) if a != b]
    assert not bad, (bad[:5], lines[bad[0]][:300], ref[bad[0]][:300])


def test_kernel_source_on_host_matches_oracle_random_options(oracle, rep_case, chain_sim):
    """Option corners the dumps do not cover: narrow chaining band, small gap limit, strict masking, drop ratio, every chain
    lighter than the floor, reads shorter than the seed length."""
    g, codes, offs, res = rep_case
    contigs = rep_contigs(oracle.CONTIG_DTYPE)
    rl = np.diff(offs)
    for kw in ({"w": 5}, {"max_chain_gap": 60}, {"mask_level": 0.1, "drop_ratio": 0.9}, {"min_chain_weight": 10000}, {"max_occ": 3, "max_chain_extend": 1},
               {"min_seed_len": 151}, {"mask_level": 0.95, "drop_ratio": 0.05, "max_occ": 50}):
        opt = oracle.default_chain_opt(**kw)
        a = oracle.chain_reads(res, rl, contigs, len(g), opt)
        b = run_sim(chain_sim, oracle, sample_hits(res, opt.max_occ), rl, contigs, len(g), opt)
        assert oracle.chain_lines(a) == oracle.chain_lines(b), kw


def test_tree_and_sort_beyond_one_node(oracle, chain_sim):
    """Synthetic seed lists with hundreds of chains per read, many at EQUAL reference positions and equal weights: the
    B-tree grows to three levels, kb_intervalp meets duplicates at every level, the introsort partitions and ties abound."""
    rng = np.random.default_rng(5)
    n_reads, l_pac = 40, 1 << 20
    smems, hits, so, ho = [], [], [0], [0]
    for r in range(n_reads):
        ns = int(rng.integers(3, 40))
        starts = np.sort(rng.integers(0, 120, size=ns))
        hb = 0
        pool = rng.integers(0, 2 * l_pac - 200, size=12)                     # few distinct positions -> duplicates
        for s in starts:
            ln = int(rng.integers(19, 31))
            c = int(rng.integers(1, 40))
            pos = np.where(rng.random(c) < 0.5, rng.choice(pool, size=c), rng.integers(0, 2 * l_pac - 200, size=c))
            smems.append((int(s), int(s) + ln, hb, c, 0))
            hits.append(pos.astype(np.uint64))
            hb += c
        so.append(len(smems)); ho.append(ho[-1] + hb)
    res = {"smems": np.array(smems, dtype=oracle.SMEM_DTYPE), "smem_off": np.array(so, np.int64), "hit_off": np.array(ho, np.int64),
           "hits": np.concatenate(hits)}
    # records sorted by (start, end) within a read, as the seeding stage delivers them
    for r in range(n_reads):
        a, b = so[r], so[r + 1]
        order = np.lexsort((res["smems"]["end"][a:b], res["smems"]["start"][a:b]))
        res["smems"][a:b] = res["smems"][a:b][order]
    contigs = np.array([(0, l_pac // 2, 0), (l_pac // 2, l_pac // 2, 1)], dtype=oracle.CONTIG_DTYPE)
    rl = np.full(n_reads, 150, np.int32)
    for kw in ({}, {"max_occ": 7}, {"w": 1000000, "max_chain_gap": 1 << 30}, {"mask_level": 2.0}):
        opt = oracle.default_chain_opt(**kw)
        a = oracle.chain_reads(res, rl, contigs, l_pac, opt)
        if kw == {"mask_level": 2.0}:                 # nothing counts as overlap: every chain survives, in sorted order
            assert np.diff(a["chain_off"]).max() > 300
        b = run_sim(chain_sim, oracle, sample_hits(res, opt.max_occ), rl, contigs, l_pac, opt)
        assert oracle.chain_lines(a) == oracle.chain_lines(b), kw
