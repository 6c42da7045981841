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
