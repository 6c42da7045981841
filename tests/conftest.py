import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return ge.load_pkg()


@pytest.fixture(scope="session")
def oracle():
    O = ge.load_oracle()
    O.build()
    return O


@pytest.fixture(scope="session")
def golden_meta():
    return json.load(open(os.path.join(GOLDEN, "index_sha.json")))


@pytest.fixture(scope="session")
def golden_genome(pkg, golden_meta):
    from bwa_meme_b200 import synth
    return synth.make_genome(golden_meta["genome_len"], seed=golden_meta["genome_seed"], n_families=3, copies=12, rep_len=400)


@pytest.fixture(scope="session")
def golden_index(pkg, golden_genome, tmp_path_factory):
    """The golden genome indexed by OUR builder (files in the reference's formats)."""
    from bwa_meme_b200 import index_build
    d = tmp_path_factory.mktemp("golden_index")
    prefix = str(d / "ref.fa")
    index_build.build_index(prefix, golden_genome, bits=16)
    return prefix


# (read_len, n_reads, n_frac, seed) -- must match tests/golden/make_golden.py
GOLDEN_READ_SETS = [(150, 300, 0.05, 21), (101, 120, 0.3, 22), (250, 100, 0.05, 23), (33, 80, 0.1, 24)]
GOLDEN_BSW_SETS = [(0, 400, 150, 110, 120), (1, 300, 60, 200, 40), (2, 200, 300, 120, 250)]


@pytest.fixture(scope="session")
def golden_reads(pkg, golden_genome):
    from bwa_meme_b200 import synth
    return {rl: synth.make_reads(golden_genome, n, rl, seed=seed, n_frac=nf) for rl, n, nf, seed in GOLDEN_READ_SETS}


@pytest.fixture(scope="session")
def mid_index(pkg, tmp_path_factory):
    """A 400 kb genome with repeats: big enough for k-ary narrowing and interval galloping."""
    from bwa_meme_b200 import synth, index_build
    d = tmp_path_factory.mktemp("mid_index")
    prefix = str(d / "mid.fa")
    g = synth.make_genome(400000, seed=77, n_families=6, copies=60, rep_len=800, divergence=0.01)
    index_build.build_index(prefix, g, bits=12)   # few leaves -> wide RMI windows
    return prefix, g


@pytest.fixture(scope="session")
def textend_case(pkg, tmp_path_factory):
    """Genome whose fwd+RC text ends in a T run, indexed by our builder, and the reads of tests/golden/smem_textend.txt
    (reads running into the text end / across the forward-reverse junction / with long T tails)."""
    import json
    from bwa_meme_b200 import synth, index_build
    meta = json.load(open(os.path.join(GOLDEN, "textend.json")))
    g = synth.make_genome_textend(meta["length"], meta["seed"], meta["start_a_run"], meta["end_t_run"])
    d = tmp_path_factory.mktemp("textend_index")
    prefix = str(d / "textend.fa")
    st = index_build.build_index(prefix, g, bits=meta["bits"])
    assert st["pad"] == meta["pad"]
    codes, offs = synth.make_textend_reads(g, meta["pad"], meta["start_a_run"], seed=meta["reads_seed"])
    assert len(offs) - 1 == meta["n_reads"]
    return prefix, codes, offs


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
