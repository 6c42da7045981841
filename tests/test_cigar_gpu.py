"""The SAM-stage global alignment on the GPU (meme_cigar_batch) through the C ABI against the
reference's own mem_reg2aln outputs (tests/golden/cigar_ref.json) and the oracle."""
import json
import os

import numpy as np
import pytest

import chain_util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cigar_case(pkg, oracle, tmp_path_factory):
    from bwa_meme_b200 import synth, index_build, capi
    g = U.rep_genome(synth)
    prefix = str(tmp_path_factory.mktemp("rep_index") / "rep.fa")
    index_build.build_index(prefix, g, bits=U.rep_meta()["bits"])
    codes, cases = U.make_cigar_cases(g, U.rep_meta()["contigs"], 1500, 4)
    tasks = U.cigar_tasks(cases, oracle.infer_w, oracle.CIGAR_TASK_DTYPE)
    ref = json.load(open(os.path.join(U.GOLDEN, "cigar_ref.json")))
    ix = capi.MemeIndex(prefix, mode=3)
    yield g, index_build.fwd_rc_text(g), codes, cases, tasks, ref, ix
    ix.close()


def test_cuda_matches_reference_mem_reg2aln(pkg, oracle, cigar_case):
    g, text, codes, cases, tasks, ref, ix = cigar_case
    out = ix.cigar_batch(tasks, codes, len(g))
    raw = oracle.unpack_cigar_result(out["recs"], out["cigars"], out["md"])
    assert len(raw) == len(ref)
    for i, (c, r, w) in enumerate(zip(cases, raw, ref)):
        s = oracle.sam_cigar(r, int(c["qb"]), int(c["qe"]), int(c["l_query"]), int(c["rb"]) >= len(g))
        assert (s["cigar"], s["md"], s["nm"]) == (w["cigar"], w["md"], w["nm"]), (i, c, s, w)
    assert raw == oracle.cigar_regions(text, len(g), codes, tasks)          # scores too
    assert out["cells"] > 0 and out["ms"] > 0


def test_cuda_matches_oracle_options_slices_and_rejects(pkg, oracle, cigar_case):
    """Non-default scoring / band options; more tasks than one slice of 2^18 (offsets rebased across slices); rejected and
    empty inputs."""
    from bwa_meme_b200 import capi
    g, text, codes, cases, tasks, ref, ix = cigar_case
    ix.cigar_warmup(5000)                                                   # dummy regions through every kernel; results below unaffected
    for kw in ({"a": 2, "b": 3}, {"o_del": 1, "e_del": 3, "o_ins": 12, "e_ins": 2}, {"w": 5}):
        out = ix.cigar_batch(tasks, codes, len(g), capi.default_cigar_opt(**kw))
        raw = oracle.unpack_cigar_result(out["recs"], out["cigars"], out["md"])
        assert raw == oracle.cigar_regions(text, len(g), codes, tasks, oracle.default_cigar_opt(**kw)), kw
    reps = (1 << 18) // len(tasks) + 2                                      # just over one slice
    big = np.tile(tasks, reps)
    out = ix.cigar_batch(big, codes, len(g))
    want = oracle.cigar_regions(text, len(g), codes, tasks)
    got = oracle.unpack_cigar_result(out["recs"], out["cigars"], out["md"])
    assert len(got) == len(big)
    for k in (0, 1, reps // 2, reps - 1):
        assert got[k * len(tasks):(k + 1) * len(tasks)] == want
    odd = tasks[:8].copy()
    odd["w"][0] = 1000
    odd["qlen"][1] = 0
    odd["rb"][2], odd["re"][2] = len(g) - 20, len(g) + 20
    odd["re"][3] = odd["rb"][3]
    odd["truesc"][4] = 10000
    out = ix.cigar_batch(odd, codes, len(g))
    assert oracle.unpack_cigar_result(out["recs"], out["cigars"], out["md"]) == oracle.cigar_regions(text, len(g), codes, odd)
    out = ix.cigar_batch(np.zeros(0, capi.CIGAR_TASK_DTYPE), codes, len(g))
    assert len(out["recs"]) == 0
