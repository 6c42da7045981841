"""On-GPU runtime index build (inline keys, inverse SA) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_keys_pos_isa(pkg, golden_index):
    from bwa_meme_b200 import capi, index_build
    ix = capi.MemeIndex(golden_index, mode=3)
    keys, pos, isa = ix.download()
    ix.close()
    want_pos = index_build.read_pos_packed(golden_index + ".pos_packed")
    assert (pos == want_pos).all()
    n = len(pos)
    assert (isa[pos.astype(np.int64)] == np.arange(n, dtype=np.uint64)).all()
    text = np.fromfile(golden_index + ".0123", dtype=np.uint8)
    padded = np.concatenate([text, np.full(32, 3, np.uint8)]).astype(np.uint64)
    want = np.zeros(n, np.uint64)
    p = pos.astype(np.int64)
    for r in range(32):     # get_key_of_ref: base r in bits 2r..2r+1, 3 past the end
        want |= padded[p + r] << np.uint64(2 * r)
    assert (keys == want).all()


def test_from_host_equals_load(pkg, golden_index, golden_reads):
    from bwa_meme_b200 import capi
    rec = np.dtype([("alpha", "<f8"), ("beta", "<f8"), ("err", "<u8")])
    arrays = {"text": np.fromfile(golden_index + ".0123", dtype=np.uint8),
              "pos_packed": np.fromfile(golden_index + ".pos_packed", dtype=np.uint8),
              "l1": np.fromfile(golden_index + ".suffixarray_uint64_L1_PARAMETERS", dtype=rec),
              "l2": np.fromfile(golden_index + ".suffixarray_uint64_L2_PARAMETERS", dtype=rec)}
    a = capi.MemeIndex(host_arrays=arrays, mode=3)
    b = capi.MemeIndex(golden_index, mode=3)
    codes, offs = golden_reads[150]
    ra, rb = a.seed(codes, offs), b.seed(codes, offs)
    a.close(); b.close()
    for k in ("smems", "hits", "smem_off", "hit_off"):
        assert (ra[k] == rb[k]).all()


def test_missing_files_error(pkg, tmp_path):
    from bwa_meme_b200 import capi
    with pytest.raises(capi.MemeError) as e:
        capi.MemeIndex(str(tmp_path / "nope.fa"), mode=3)
    assert "cannot open" in str(e.value)


def test_native_suffix_array_matches_reference_files_and_torch(pkg, golden_index, golden_meta, tmp_path):
    """meme_sa_build (row f3: prefix doubling with device radix sorts) gives the suffix array and keys the reference's own builder
    wrote for the golden genome (byte-equal files, sha256 from `bwa-meme index -a meme`), and agrees with the torch restatement
    on low-entropy texts, homopolymer ends and a repeat-rich genome."""
    import hashlib
    import torch
    from bwa_meme_b200 import capi, index_build, synth
    text = np.fromfile(golden_index + ".0123", dtype=np.uint8)
    r = capi.sa_build(text)
    assert (r["sa"] == index_build.read_pos_packed(golden_index + ".pos_packed").astype(np.int64)).all()
    raw = np.fromfile(golden_index + ".suffixarray_uint64", dtype=np.uint64)
    assert int(raw[0]) == len(text) and (raw[1:] == r["keys"]).all()
    # the whole pipeline through the native builder: every file byte-equal to the reference builder's
    g = text[:len(text) // 2]
    prefix = str(tmp_path / "native.fa")
    index_build.build_index(prefix, g, bits=16, device="cuda")
    for ext, want in golden_meta["sha256"].items():
        if "PARAMETERS" in ext:
            continue                       # the trainer is the repo's own (format-compatible), pinned by test_index_build.py
        assert hashlib.sha256(open(prefix + ext, "rb").read()).hexdigest() == want, ext
    rng = np.random.default_rng(7)
    texts = [index_build.fwd_rc_text(rng.integers(0, 2, size=int(rng.integers(20, 3000)), dtype=np.uint8)) for _ in range(4)]
    texts.append(index_build.fwd_rc_text(np.array([0, 0, 0, 1, 2, 3, 0, 0], np.uint8)))
    texts.append(index_build.fwd_rc_text(np.zeros(500, np.uint8)))                         # one long A run + one long T run
    texts.append(index_build.fwd_rc_text(synth.make_genome(200000, seed=9, n_families=6, copies=40, rep_len=800)))
    big = index_build.fwd_rc_text(synth.make_genome_large(3000000, n_families=10, copies=30))          # several doubling rounds at 6 M suffixes
    want = index_build.suffix_array(big, device="cuda")
    got = capi.sa_build(big)
    assert got["rounds"] >= 2 and (got["sa"] == want.cpu().numpy()).all()
    for t in texts:
        want = index_build.suffix_array(t)
        pad = index_build.pad_len_of(torch.as_tensor(t))
        got = capi.sa_build(t)
        assert got["pad"] == pad
        assert (got["sa"] == want.numpy()).all()
        assert (got["keys"].view(np.int64) == index_build.rmi_keys(t, want, pad).numpy()).all()


@pytest.mark.parametrize("bits", [4, 10, 16])
def test_native_trainer_error_bounds_hold(pkg, golden_index, mid_index, bits):
    """meme_prmi_train (row f3): for every training key the true rank lies inside [pred - below, pred + above] of the model the
    reference's lookup selects (learned_index_lookup, restated in prmi_train.lookup_numpy), with plain and dense leaves, and
    unseen keys stay within one of it; the seeding kernel gives the same SMEMs with these models as with the torch-trained ones."""
    from bwa_meme_b200 import capi, prmi_train
    for prefix in (golden_index, mid_index[0]):
        raw = np.fromfile(prefix + ".suffixarray_uint64", dtype=np.uint64)
        n, keys = int(raw[0]), raw[1:]
        l1, l2, stats = capi.prmi_train(keys, bits)
        assert l2.shape[0] == 1 << bits and l1.shape[0] >= 1
        if bits == 4:
            assert stats["n_dense_leaf"] > 0 and (l2["err"] >> np.uint64(63)).any()
        pos, below, above = prmi_train.lookup_numpy(l1, l2, keys, n)
        d = pos - np.arange(n)
        assert (d <= below).all() and (-d <= above).all()
        rng = np.random.default_rng(3)
        q = keys[rng.integers(0, n, 20000)] ^ (np.uint64(1) << rng.integers(0, 50, 20000).astype(np.uint64))
        pos, below, above = prmi_train.lookup_numpy(l1, l2, q, n)
        d = pos - np.searchsorted(keys, q)
        assert (d <= below + 1).all() and (-d <= above + 1).all()


def test_native_index_seeds_like_torch_index(pkg, golden_genome, golden_index, golden_reads, tmp_path):
    from bwa_meme_b200 import capi, index_build
    prefix = str(tmp_path / "native.fa")
    index_build.build_index(prefix, golden_genome, bits=10, device="cuda")       # few leaves: dense ones with sub-models
    a, b = capi.MemeIndex(prefix, mode=3), capi.MemeIndex(golden_index, mode=3)
    codes, offs = golden_reads[150]
    ra, rb = a.seed(codes, offs), b.seed(codes, offs)
    a.close(); b.close()
    for k in ("smems", "hits", "smem_off", "hit_off"):
        assert (ra[k] == rb[k]).all()
