"""Chaining on the GPU (meme_chain_run / meme_chain_batch_multi) through the C ABI against the
chaining oracle and the committed chain dumps of the reference's own mem_chain_Learned + mem_chain_flt."""
import os

import numpy as np
import pytest

from chain_util import rep_genome, rep_reads, rep_contigs, rep_meta, read_dump
from conftest import GOLDEN_READ_SETS

pytestmark = pytest.mark.gpu

VARIANTS = [("chains_rep.txt", {}), ("chains_rep_c8_W40_N3.txt", {"max_occ": 8, "min_chain_weight": 40, "max_chain_extend": 3})]


@pytest.fixture(scope="module")
def rep_case(pkg, tmp_path_factory):
    from bwa_meme_b200 import synth, index_build, capi
    g = rep_genome(synth)
    prefix = str(tmp_path_factory.mktemp("rep_index") / "rep.fa")
    index_build.build_index(prefix, g, bits=rep_meta()["bits"])
    codes, offs = rep_reads(synth, g)
    ix = capi.MemeIndex(prefix, mode=3)
    ix.chain_set_contigs(rep_contigs(capi.CONTIG_DTYPE), len(g))
    yield g, codes, offs, ix, prefix
    ix.close()


def _opts(capi, kw):
    co = capi.default_chain_opt(**kw)
    so = capi.default_seed_opt(max_occ=co.max_occ)
    return so, co


@pytest.mark.parametrize("dump,kw", VARIANTS)
def test_cuda_matches_reference_dump(pkg, oracle, rep_case, dump, kw):
    from bwa_meme_b200 import capi
    g, codes, offs, ix, _ = rep_case
    so, co = _opts(capi, kw)
    ix.seed_upload(codes, offs)
    ix.seed_run(so)
    ix.chain_run(co)
    out = ix.chain_download()
    lines = oracle.chain_lines(out, names=[f"r{i}" for i in range(len(offs) - 1)])
    ref = read_dump(dump)
    bad = [i for i, (a, b) in enumerate(zip(lines, ref)) if a != b]
    assert len(lines) == len(ref) and not bad, (bad[:5], lines[bad[0]][:300], ref[bad[0]][:300])
    assert ix.chain_last_ms() > 0


@pytest.mark.parametrize("rl", [rs[0] for rs in GOLDEN_READ_SETS])
def test_cuda_matches_reference_dump_golden(pkg, oracle, golden_index, golden_genome, golden_reads, rl):
    from bwa_meme_b200 import capi
    codes, offs = golden_reads[rl]
    ix = capi.MemeIndex(golden_index, mode=3)
    ix.chain_set_contigs(np.array([(0, len(golden_genome), 0)], dtype=capi.CONTIG_DTYPE), len(golden_genome))
    out = ix.chain(codes, offs, *_opts(capi, {}))
    ix.close()
    assert oracle.chain_lines(out, names=[f"r{i}" for i in range(len(offs) - 1)]) == read_dump(f"chains_len{rl}.txt")


def test_cuda_matches_oracle_option_corners_and_modes(pkg, oracle, rep_case):
    """Options the dumps do not cover, on index layouts 1-3 and with the 64-bit-index kernels, against the oracle fed with
    the seeding oracle's lists."""
    from bwa_meme_b200 import capi
    g, codes, offs, ix, prefix = rep_case
    oix = oracle.OracleIndex(prefix, 3)
    res = oix.seed(codes, offs)
    oix.close()
    contigs = rep_contigs(oracle.CONTIG_DTYPE)
    rl = np.diff(offs)
    for kw in ({"w": 5}, {"max_chain_gap": 60}, {"mask_level": 0.1, "drop_ratio": 0.9}, {"min_chain_weight": 10000}, {"max_occ": 3, "max_chain_extend": 1},
               {"min_seed_len": 151}, {"mask_level": 2.0}):
        so, co = _opts(capi, kw)
        want = oracle.chain_lines(oracle.chain_reads(res, rl, contigs, len(g), oracle.default_chain_opt(**kw)))
        got = oracle.chain_lines(ix.chain(codes, offs, so, co))
        assert got == want, kw
    want = oracle.chain_lines(oracle.chain_reads(res, rl, contigs, len(g)))
    for mode in (1, 2):
        ix2 = capi.MemeIndex(prefix, mode=mode)
        ix2.chain_set_contigs(rep_contigs(capi.CONTIG_DTYPE), len(g))
        assert oracle.chain_lines(ix2.chain(codes, offs, *_opts(capi, {}))) == want, mode
        ix2.close()


def test_sub_batches_and_empty_inputs(pkg, oracle, rep_case, monkeypatch):
    """meme_chain_batch_multi cuts a device's share into sub-batches of MEME_SEED_CHUNK reads and appends their chains (offsets
    and seed_beg rebased): same result as one batch.  Empty batch, reads shorter than a seed, all-N reads."""
    from bwa_meme_b200 import capi
    g, codes, offs, ix, _ = rep_case
    so, co = _opts(capi, {})
    whole = oracle.chain_lines(ix.chain(codes, offs, so, co))
    monkeypatch.setenv("MEME_SEED_CHUNK", "1024")      # the floor: sub-batches of 1024 reads
    reps = 6                                            # 3000 reads -> 3 sub-batches (1024, 1024, 952)
    codes6 = np.tile(codes, reps)
    offs6 = np.concatenate([[0], np.cumsum(np.tile(np.diff(offs), reps))]).astype(np.int64)
    out = ix.chain(codes6, offs6, so, co)
    assert oracle.chain_lines(out) == [f"{i} " + l.split(" ", 1)[1] for i, l in enumerate(whole * reps)]
    monkeypatch.delenv("MEME_SEED_CHUNK")
    empty = ix.chain(np.zeros(0, np.uint8), np.zeros(1, np.int64), so, co)
    assert len(empty["chains"]) == 0 and list(empty["chain_off"]) == [0]
    odd_codes = np.concatenate([codes[:10], np.full(150, 4, np.uint8), codes[:150]]).astype(np.uint8)
    odd_offs = np.array([0, 10, 160, 310], np.int64)
    odd = ix.chain(odd_codes, odd_offs, so, co)
    assert list(np.diff(odd["chain_off"])[:2]) == [0, 0] and np.diff(odd["chain_off"])[2] >= 1


def test_chain_run_argument_checks(pkg, rep_case):
    from bwa_meme_b200 import capi
    g, codes, offs, ix, prefix = rep_case
    ix.seed_upload(codes, offs)
    ix.seed_run(capi.default_seed_opt(max_occ=0))              # unsampled lists: chaining must refuse them
    with pytest.raises(capi.MemeError, match="max_occ"):
        ix.chain_run(capi.default_chain_opt())
    ix2 = capi.MemeIndex(prefix, mode=3)
    ix2.seed_upload(codes, offs)
    ix2.seed_run(capi.default_seed_opt(max_occ=500))
    with pytest.raises(capi.MemeError, match="contigs"):
        ix2.chain_run(capi.default_chain_opt())
    ix2.close()


def test_two_gpus_same_chains(pkg, oracle, rep_case):
    import torch
    from bwa_meme_b200 import capi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g, codes, offs, ix, prefix = rep_case
    so, co = _opts(capi, {})
    whole = oracle.chain_lines(ix.chain(codes, offs, so, co))
    ix1 = capi.MemeIndex(prefix, mode=3, device=1)
    ix1.chain_set_contigs(rep_contigs(capi.CONTIG_DTYPE), len(g))
    outs, sb = capi.chain_batch_multi([ix, ix1], codes, offs, so, co, pair_aligned=True)
    ix1.close()
    assert sb[0] == 0 and sb[2] == len(offs) - 1 and sb[1] % 2 == 0
    got = []
    for d, o in enumerate(outs):
        got += [l.split(" ", 1)[1] for l in oracle.chain_lines(o)]
    assert got == [l.split(" ", 1)[1] for l in whole]
