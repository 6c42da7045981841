"""The C-ABI library loads and exports every symbol include/meme_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
import __graft_entry__ as ge


@pytest.fixture(scope="module")
def built():
    ge.build_lib()
    return os.path.join(ROOT, "bwa-meme_b200", "libmeme_b200.so")


def test_header_symbols_exported(built, pkg):
    hdr = open(os.path.join(ROOT, "include", "meme_b200.h")).read()
    declared = set(re.findall(r"\b(meme_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = C.CDLL(built)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    from bwa_meme_b200 import capi
    assert set(capi.EXPORTS) == declared


def test_struct_layouts_match_reference(pkg):
    from bwa_meme_b200 import capi
    assert capi.SMEM_DTYPE.itemsize == 24       # mem_tl, src/LearnedIndex_seeding.h:L121-127
    assert capi.SEQPAIR_DTYPE.itemsize == 56    # SeqPair, src/bandedSWA.h:L90-99
    assert C.sizeof(capi.SeedOpt) == 24 and C.sizeof(capi.BswOpt) == 32


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU error path")
def test_fails_loudly_without_gpu(built, pkg, golden_index):
    from bwa_meme_b200 import capi
    with pytest.raises(capi.MemeError) as e:
        capi.MemeIndex(golden_index, mode=3)
    assert "no CPU path" in str(e.value) or "-6" in str(e.value)
    with pytest.raises(capi.MemeError):
        capi.MemeBsw()
    # the index pipeline has no CPU path behind the C ABI either (index_build.py's torch code is separate tooling)
    import numpy as np
    with pytest.raises(capi.MemeError):
        capi.sa_build(np.zeros(64, np.uint8))
    with pytest.raises(capi.MemeError):
        capi.prmi_train(np.arange(100, dtype=np.uint64), 4)
    assert capi.KSWR_DTYPE.itemsize == 28       # kswr_t, src/ksw.h:L45-50


def test_version_string(built, pkg):
    from bwa_meme_b200 import capi
    assert b"sm_100a" in capi.lib().meme_version()
