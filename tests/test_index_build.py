"""Index pipeline (suffix array, file formats, P-RMI trainer) -- CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch


def test_files_match_reference_builder(golden_index, golden_meta):
    """Every file `bwa-meme index -a meme` writes is reproduced byte for byte (sha256 from the reference run)."""
    for ext, want in golden_meta["sha256"].items():
        got = hashlib.sha256(open(golden_index + ext, "rb").read()).hexdigest()
        assert got == want, ext


def test_suffix_array_small_bruteforce(pkg):
    from bwa_meme_b200 import index_build
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = int(rng.integers(20, 400))
        g = rng.integers(0, 4 if trial % 2 else 2, size=n, dtype=np.uint8)   # low-entropy texts stress the doubling
        text = index_build.fwd_rc_text(g)
        sa = index_build.suffix_array(text).numpy()
        pad = index_build.pad_len_of(torch.as_tensor(text))
        padded = bytes(text.tolist()) + bytes([3] * pad)
        want = sorted(range(len(text)), key=lambda i: padded[i:])
        assert sa.tolist() == want


def test_text_ending_in_t_run(pkg):
    """suffixes that run into the T padding sort as the reference's sais does (shorter one first)."""
    from bwa_meme_b200 import index_build
    g = np.array([0, 0, 0, 1, 2, 3, 0, 0], np.uint8)      # RC ends in TTT
    text = index_build.fwd_rc_text(g)
    sa = index_build.suffix_array(text).numpy()
    pad = index_build.pad_len_of(torch.as_tensor(text))
    padded = bytes(text.tolist()) + bytes([3] * pad)
    assert sa.tolist() == sorted(range(len(text)), key=lambda i: padded[i:])


@pytest.mark.parametrize("bits", [4, 10, 16])
def test_trainer_error_bounds_hold(pkg, golden_index, bits, tmp_path):
    """for every training key AND for unseen keys the true rank lies inside [pred-below, pred+above]."""
    from bwa_meme_b200 import prmi_train
    raw = np.fromfile(golden_index + ".suffixarray_uint64", dtype=np.uint64)
    n, keys = int(raw[0]), raw[1:]
    base = str(tmp_path / "t")
    prmi_train.train_and_write(base, torch.from_numpy(keys.view(np.int64).copy()), bits)
    rec = np.dtype([("alpha", "<f8"), ("beta", "<f8"), ("err", "<u8")])
    l1 = np.fromfile(base + "_L1_PARAMETERS", dtype=rec)
    l2 = np.fromfile(base + "_L2_PARAMETERS", dtype=rec)
    assert l2.shape[0] == 1 << bits and l1.shape[0] >= 1
    pos, below, above = prmi_train.lookup_numpy(l1, l2, keys, n)
    d = pos - np.arange(n)
    assert (d <= below).all() and (-d <= above).all()
    rng = np.random.default_rng(3)
    q = keys[rng.integers(0, n, 20000)] ^ (np.uint64(1) << rng.integers(0, 50, 20000).astype(np.uint64))
    pos, below, above = prmi_train.lookup_numpy(l1, l2, q, n)
    t = np.searchsorted(keys, q)
    d = pos - t
    assert (d <= below + 1).all() and (-d <= above + 1).all()


def test_pos_packed_roundtrip(pkg, tmp_path):
    from bwa_meme_b200 import index_build
    sa = torch.tensor([0, 255, 256, 65535, 1 << 31, (1 << 39) + 12345], dtype=torch.int64)
    p = str(tmp_path / "x.pos_packed")
    index_build.write_pos_packed(p, sa)
    assert os.path.getsize(p) == 5 * len(sa)
    assert index_build.read_pos_packed(p).tolist() == sa.tolist()
