"""ctypes binding of include/meme_b200.h (the C-ABI shared library libmeme_b200.so).

This is the only way Python code in this repo reaches the CUDA kernels: tests, bench.py and the
multi-GPU driver all call through the same `extern "C"` entry points a C++ host would.  There is
no CPU fallback here -- if the library is missing or no GPU is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEME_B200_LIB") or os.path.join(HERE, "libmeme_b200.so")   # same override as the C++ glue

SMEM_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("hitbeg", "<i4"), ("hitcount", "<i4"), ("cache_refpos", "<u8")])
SEQPAIR_DTYPE = np.dtype([("idr", "<i4"), ("idq", "<i4"), ("id", "<i4"), ("len1", "<i4"), ("len2", "<i4"),
                          ("h0", "<i4"), ("seqid", "<i4"), ("regid", "<i4"), ("score", "<i4"), ("tle", "<i4"),
                          ("gtle", "<i4"), ("qle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
RESULT_FIELDS = ("score", "tle", "gtle", "qle", "gscore", "max_off")

EXPORTS = [
    "meme_index_load", "meme_rmi_lookup", "meme_index_load_devices", "meme_seed_batch_multi", "meme_index_from_host", "meme_index_free", "meme_index_sa_num", "meme_index_device_bytes",
    "meme_index_set_stream", "meme_index_download", "meme_seed_batch", "meme_seed_upload", "meme_seed_run",
    "meme_seed_warmup", "meme_seed_copy_device", "meme_seed_download", "meme_seed_totals", "meme_seed_counters", "meme_seed_last_ms", "meme_bsw_create", "meme_bsw_free",
    "meme_bsw_set_stream", "meme_bsw_batch", "meme_bsw_upload", "meme_bsw_run", "meme_bsw_download", "meme_bsw_cells",
    "meme_bsw_attach_text", "meme_bsw_upload_segments", "meme_bsw_download_segments", "meme_bsw_batch_ext", "meme_bsw_set_match", "meme_bsw_run_ext", "meme_bsw_download_w", "meme_pinned_alloc", "meme_pinned_free",
    "meme_chain_set_contigs", "meme_chain_run", "meme_chain_download", "meme_chain_last_ms", "meme_chain_batch_multi", "meme_chain_warmup",
    "meme_cigar_batch", "meme_cigar_last", "meme_cigar_warmup",
    "meme_matesw_batch", "meme_matesw_last", "meme_matesw_warmup", "meme_sa_build", "meme_prmi_train",
    "meme_last_error", "meme_device_count", "meme_version",
]

CONTIG_DTYPE = np.dtype([("offset", "<i8"), ("len", "<i4"), ("is_alt", "<i4")])
CHAIN_DTYPE = np.dtype([("pos", "<i8"), ("seed_beg", "<i8"), ("n", "<i4"), ("rid", "<i4"), ("w", "<i4"), ("first", "<i4"),
                        ("kept", "<i4"), ("is_alt", "<i4")])
CHAIN_SEED_DTYPE = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4")])
CIGAR_TASK_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("q_off", "<i8"), ("qlen", "<i4"), ("w", "<i4"), ("truesc", "<i4"), ("pad", "<i4")])
CIGAR_REC_DTYPE = np.dtype([("score", "<i4"), ("nm", "<i4"), ("n_cigar", "<i4"), ("md_len", "<i4"), ("cigar_off", "<i8"), ("md_off", "<i8")])


class SeedOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int32), ("split_len", C.c_int32), ("split_width", C.c_int32),
                ("max_mem_intv", C.c_int32), ("rounds", C.c_int32), ("max_occ", C.c_int32)]


class SeedResult(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_smems", C.c_int64), ("n_hits", C.c_int64), ("smems", C.c_void_p),
                ("smem_off", C.c_void_p), ("hit_off", C.c_void_p), ("hits", C.c_void_p)]


class BswOpt(C.Structure):
    _fields_ = [("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32),
                ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


class CigarOpt(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("pad", C.c_int8 * 3), ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32),
                ("e_ins", C.c_int32), ("w", C.c_int32), ("a", C.c_int32)]


class CigarResult(C.Structure):
    _fields_ = [("n_tasks", C.c_int64), ("n_ops", C.c_int64), ("n_md", C.c_int64), ("recs", C.c_void_p), ("cigars", C.c_void_p),
                ("md", C.c_void_p)]


KSWR_DTYPE = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4")])
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000


class MateswOpt(C.Structure):
    _fields_ = [("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


def default_matesw_opt(a: int = 1, b: int = 4, o_del: int = 6, e_del: int = 1, o_ins: int = 6, e_ins: int = 1) -> MateswOpt:
    return MateswOpt(o_del, e_del, o_ins, e_ins, a, b)


def default_cigar_opt(a: int = 1, b: int = 4, o_del: int = 6, e_del: int = 1, o_ins: int = 6, e_ins: int = 1, w: int = 100) -> CigarOpt:
    """bwa_fill_scmat (reference src/bwa.cpp) + mem_opt_init defaults."""
    o = CigarOpt()
    k = 0
    for i in range(4):
        for j in range(4):
            o.mat[k] = a if i == j else -b
            k += 1
        o.mat[k] = -1
        k += 1
    for j in range(5):
        o.mat[k] = -1
        k += 1
    o.o_del, o.e_del, o.o_ins, o.e_ins, o.w, o.a = o_del, e_del, o_ins, e_ins, w, a
    return o


class ChainOpt(C.Structure):
    _fields_ = [("w", C.c_int32), ("max_chain_gap", C.c_int32), ("max_occ", C.c_int32), ("min_seed_len", C.c_int32),
                ("min_chain_weight", C.c_int32), ("max_chain_extend", C.c_int32), ("mask_level", C.c_float), ("drop_ratio", C.c_float)]


class ChainResult(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_chains", C.c_int64), ("n_seeds", C.c_int64), ("chain_off", C.c_void_p),
                ("frac_rep", C.c_void_p), ("chains", C.c_void_p), ("seeds", C.c_void_p)]


def default_chain_opt(w: int = 100, max_chain_gap: int = 10000, max_occ: int = 500, min_seed_len: int = 19, min_chain_weight: int = 0,
                      max_chain_extend: int = 1 << 30, mask_level: float = 0.5, drop_ratio: float = 0.5) -> ChainOpt:
    """mem_opt_init defaults (reference src/bwamem.cpp:L126-162)."""
    return ChainOpt(w, max_chain_gap, max_occ, min_seed_len, min_chain_weight, max_chain_extend, mask_level, drop_ratio)


def default_seed_opt(rounds: int = 3, min_seed_len: int = 19, split_factor: float = 1.5, split_width: int = 10,
                     max_mem_intv: int = 20, max_occ: int = 0) -> SeedOpt:
    """mem_opt_init defaults (reference src/bwamem.cpp:L126-162)."""
    return SeedOpt(min_seed_len, int(min_seed_len * split_factor + .499), split_width, max_mem_intv, rounds, max_occ)


def default_bsw_opt() -> BswOpt:
    return BswOpt(6, 1, 6, 1, 100, 5, 1, 4)


class MemeError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MemeError(f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.meme_index_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
        L.meme_index_from_host.argtypes = [vp, u64, vp, vp, u64, vp, u64, C.c_int, C.c_int, C.POINTER(vp)]
        L.meme_index_free.argtypes = [vp]
        L.meme_index_free.restype = None
        L.meme_index_sa_num.argtypes = [vp]
        L.meme_index_sa_num.restype = u64
        L.meme_index_device_bytes.argtypes = [vp]
        L.meme_index_device_bytes.restype = u64
        L.meme_index_set_stream.argtypes = [vp, vp]
        L.meme_index_download.argtypes = [vp, vp, vp, vp]
        L.meme_seed_batch.argtypes = [vp, vp, vp, i64, C.POINTER(SeedOpt), C.POINTER(SeedResult)]
        L.meme_seed_upload.argtypes = [vp, vp, vp, i64]
        L.meme_seed_run.argtypes = [vp, C.POINTER(SeedOpt)]
        L.meme_seed_download.argtypes = [vp, C.POINTER(SeedResult)]
        L.meme_seed_totals.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.meme_seed_counters.argtypes = [vp, C.POINTER(u64 * 8)]
        L.meme_seed_last_ms.argtypes = [vp, C.POINTER(C.c_float * 3)]
        L.meme_bsw_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.meme_bsw_free.argtypes = [vp]
        L.meme_bsw_free.restype = None
        L.meme_bsw_set_stream.argtypes = [vp, vp]
        L.meme_bsw_batch.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, C.POINTER(BswOpt)]
        L.meme_bsw_upload.argtypes = [vp, vp, i64, vp, i64, vp, i64]
        L.meme_bsw_run.argtypes = [vp, i32, C.POINTER(BswOpt)]
        L.meme_bsw_download.argtypes = [vp, vp]
        L.meme_bsw_cells.argtypes = [vp, C.POINTER(u64)]
        L.meme_bsw_batch_ext.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, i32, i32, C.POINTER(BswOpt), vp]
        L.meme_bsw_set_match.argtypes = [vp, i32]
        L.meme_bsw_run_ext.argtypes = [vp, i32, i32, i32, C.POINTER(BswOpt)]
        L.meme_bsw_download_w.argtypes = [vp, vp]
        L.meme_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        L.meme_pinned_free.argtypes = [vp]
        L.meme_pinned_free.restype = None
        L.meme_chain_set_contigs.argtypes = [vp, vp, i32, i64]
        L.meme_chain_run.argtypes = [vp, C.POINTER(ChainOpt)]
        L.meme_chain_download.argtypes = [vp, C.POINTER(ChainResult)]
        L.meme_chain_last_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.meme_chain_batch_multi.argtypes = [vp, C.c_int, vp, vp, i64, C.c_int, C.POINTER(SeedOpt), C.POINTER(ChainOpt), vp, vp]
        L.meme_cigar_batch.argtypes = [vp, vp, i64, vp, i64, i64, C.POINTER(CigarOpt), C.POINTER(CigarResult)]
        L.meme_cigar_last.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u64)]
        L.meme_cigar_warmup.argtypes = [vp, i64, C.c_int]
        L.meme_matesw_batch.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp, C.POINTER(MateswOpt), vp]
        L.meme_matesw_last.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(u64)]
        L.meme_matesw_warmup.argtypes = [vp, i64, C.c_int, C.c_int]
        L.meme_sa_build.argtypes = [C.c_int, vp, i64, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.meme_prmi_train.argtypes = [C.c_int, vp, i64, C.c_int32, vp, vp, i64, C.POINTER(i64), vp]
        L.meme_last_error.restype = C.c_char_p
        L.meme_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        raise MemeError(f"{what} failed with {rc}: {lib().meme_last_error().decode()}")


def sa_build(text: np.ndarray, device: int = 0, want_keys: bool = True) -> dict:
    """meme_sa_build: suffix array (int64) of the fwd+RC text in the reference's order, the suffixes' 64-bit keys, the padding."""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(text)
    sa = np.empty(n, np.int64)
    keys = np.empty(n, np.uint64) if want_keys else None
    pad, rounds = C.c_int32(), C.c_int32()
    _check(lib().meme_sa_build(int(device), text.ctypes.data, n, sa.ctypes.data, keys.ctypes.data if want_keys else None, C.byref(pad), C.byref(rounds)),
           "meme_sa_build")
    return {"sa": sa, "keys": keys, "pad": pad.value, "rounds": rounds.value}


RMI_REC_DTYPE = np.dtype([("alpha", "<f8"), ("beta", "<f8"), ("err", "<u8")])


def prmi_train(keys: np.ndarray, bits: int, device: int = 0):
    """meme_prmi_train: (l1 records, l2 records, stats) from the sorted uint64 keys."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    n = len(keys)
    l2 = np.zeros(1 << bits, RMI_REC_DTYPE)
    cap = n // 20 + 1
    l1 = np.zeros(cap, RMI_REC_DTYPE)
    n_l1 = C.c_int64()
    stats = np.zeros(2, np.int64)
    _check(lib().meme_prmi_train(int(device), keys.ctypes.data, n, int(bits), l2.ctypes.data, l1.ctypes.data, cap, C.byref(n_l1), stats.ctypes.data),
           "meme_prmi_train")
    return l1[:n_l1.value].copy(), l2, {"n_leaf": 1 << bits, "n_dense_leaf": int(stats[0]), "n_sub": int(n_l1.value), "max_window": int(stats[1])}


class MemeIndex:
    """Host mirror of the reference's index ownership in worker_t (sa_position / ref2sa / rc_pac,
    src/bwamem.h:L231-268) -- here a handle to HBM-resident arrays."""

    def __init__(self, prefix: str | None = None, mode: int = 3, device: int = 0, host_arrays: dict | None = None):
        L = lib()
        h = C.c_void_p()
        if host_arrays is not None:
            a = host_arrays
            self._keep = a
            _check(L.meme_index_from_host(a["text"].ctypes.data, int(a["text"].shape[0]), a["pos_packed"].ctypes.data,
                                          a["l1"].ctypes.data, int(a["l1"].shape[0]), a["l2"].ctypes.data,
                                          int(a["l2"].shape[0]), mode, device, C.byref(h)), "meme_index_from_host")
        else:
            _check(L.meme_index_load(prefix.encode(), mode, device, C.byref(h)), "meme_index_load")
        self.h = h
        self.mode = mode
        self.device = device
        self.sa_num = int(L.meme_index_sa_num(h))
        self.device_bytes = int(L.meme_index_device_bytes(h))

    def close(self):
        if self.h:
            lib().meme_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        _check(lib().meme_index_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)), "meme_index_set_stream")

    def download(self, keys=True, pos=True, isa=True):
        n = self.sa_num
        k = np.empty(n, np.uint64) if keys else None
        p = np.empty(n, np.uint64) if pos else None
        i = np.empty(n, np.uint64) if isa else None
        _check(lib().meme_index_download(self.h, k.ctypes.data if keys else None, p.ctypes.data if pos else None,
                                         i.ctypes.data if isa else None), "meme_index_download")
        return k, p, i

    # ---- seeding ----
    @staticmethod
    def _wrap(res: SeedResult) -> dict:
        n, ns, nh = int(res.n_reads), int(res.n_smems), int(res.n_hits)

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt).copy()

        return {"smems": arr(res.smems, ns, SMEM_DTYPE), "smem_off": arr(res.smem_off, n + 1, np.int64),
                "hit_off": arr(res.hit_off, n + 1, np.int64), "hits": arr(res.hits, nh, np.uint64)}

    def seed(self, codes: np.ndarray, offsets: np.ndarray, opt: SeedOpt | None = None) -> dict:
        opt = opt or default_seed_opt()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        res = SeedResult()
        _check(lib().meme_seed_batch(self.h, codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, C.byref(opt),
                                     C.byref(res)), "meme_seed_batch")
        return self._wrap(res)

    def seed_batch_ptr(self, codes_ptr: int, offsets_ptr: int, n_reads: int, opt: SeedOpt) -> dict:
        """meme_seed_batch on raw host pointers; the lists stay in the handle's pinned arrays (no numpy copy)."""
        res = SeedResult()
        _check(lib().meme_seed_batch(self.h, C.c_void_p(codes_ptr), C.c_void_p(offsets_ptr), n_reads, C.byref(opt), C.byref(res)),
               "meme_seed_batch")
        return {"n_reads": int(res.n_reads), "n_smems": int(res.n_smems), "n_hits": int(res.n_hits), "_res": res}

    def wrap_result(self, r: dict) -> dict:
        return self._wrap(r["_res"])

    def seed_upload(self, codes: np.ndarray, offsets: np.ndarray):
        _check(lib().meme_seed_upload(self.h, codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1), "meme_seed_upload")

    def seed_upload_ptr(self, codes_ptr: int, offsets_ptr: int, n_reads: int):
        _check(lib().meme_seed_upload(self.h, C.c_void_p(codes_ptr), C.c_void_p(offsets_ptr), n_reads), "meme_seed_upload")

    def seed_run(self, opt: SeedOpt):
        _check(lib().meme_seed_run(self.h, C.byref(opt)), "meme_seed_run")

    def seed_download(self, copy: bool = True) -> dict:
        res = SeedResult()
        _check(lib().meme_seed_download(self.h, C.byref(res)), "meme_seed_download")
        if not copy:
            return {"n_reads": int(res.n_reads), "n_smems": int(res.n_smems), "n_hits": int(res.n_hits)}
        return self._wrap(res)

    def seed_copy_device(self, d_smems: int, d_hits: int, d_smem_off: int, d_hit_off: int):
        vp = C.c_void_p
        lib().meme_seed_copy_device.argtypes = [vp, vp, vp, vp, vp]
        _check(lib().meme_seed_copy_device(self.h, vp(d_smems), vp(d_hits), vp(d_smem_off), vp(d_hit_off)), "meme_seed_copy_device")

    def rmi_lookup(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        pred = np.zeros(len(keys), np.uint64)
        err = np.zeros(len(keys), np.uint64)
        lib().meme_rmi_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _check(lib().meme_rmi_lookup(self.h, keys.ctypes.data, len(keys), pred.ctypes.data, err.ctypes.data), "meme_rmi_lookup")
        return pred, err

    def seed_totals(self):
        a, b = C.c_int64(), C.c_int64()
        _check(lib().meme_seed_totals(self.h, C.byref(a), C.byref(b)), "meme_seed_totals")
        return int(a.value), int(b.value)

    def seed_last_ms(self):
        ms = (C.c_float * 3)()
        _check(lib().meme_seed_last_ms(self.h, C.byref(ms)), "meme_seed_last_ms")
        return {"seed": float(ms[0]), "offsets": float(ms[1]), "gather": float(ms[2])}

    def seed_counters(self) -> dict:
        c = (C.c_uint64 * 8)()
        _check(lib().meme_seed_counters(self.h, C.byref(c)), "meme_seed_counters")
        return {"searches": int(c[0]), "block_probes": int(c[1]), "kary_rounds": int(c[2]), "ext_rounds": int(c[3]),
                "pac_rounds": int(c[4]), "launches": int(c[5]), "reseed_tasks": int(c[6]), "round_trips": int(c[7])}


    # ---- chaining (mem_chain_Learned + mem_chain_flt on the device-resident lists of the last seed_run) ----
    def chain_set_contigs(self, contigs: np.ndarray, l_pac: int):
        ct = np.ascontiguousarray(contigs, dtype=CONTIG_DTYPE)
        _check(lib().meme_chain_set_contigs(self.h, ct.ctypes.data, len(ct), int(l_pac)), "meme_chain_set_contigs")

    def chain_run(self, opt: ChainOpt):
        _check(lib().meme_chain_run(self.h, C.byref(opt)), "meme_chain_run")

    @staticmethod
    def _wrap_chains(res: ChainResult) -> dict:
        n, nc, ns = int(res.n_reads), int(res.n_chains), int(res.n_seeds)

        def arr(ptr, count, dt):
            if count == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            return np.frombuffer(C.string_at(ptr, count * np.dtype(dt).itemsize), dtype=dt).copy()
        return {"chain_off": arr(res.chain_off, n + 1, np.int64) if n else np.zeros(1, np.int64), "frac_rep": arr(res.frac_rep, n, np.float32),
                "chains": arr(res.chains, nc, CHAIN_DTYPE), "seeds": arr(res.seeds, ns, CHAIN_SEED_DTYPE)}

    def chain_download(self, copy: bool = True) -> dict:
        res = ChainResult()
        _check(lib().meme_chain_download(self.h, C.byref(res)), "meme_chain_download")
        if not copy:
            return {"n_reads": int(res.n_reads), "n_chains": int(res.n_chains), "n_seeds": int(res.n_seeds)}
        return self._wrap_chains(res)

    def chain_last_ms(self) -> float:
        ms = C.c_float()
        _check(lib().meme_chain_last_ms(self.h, C.byref(ms)), "meme_chain_last_ms")
        return float(ms.value)

    # ---- global alignment of the SAM stage (mem_reg2aln's loop around bwa_gen_cigar2) ----
    # ---- mate-rescue Smith-Waterman (mem_sam_pe_batch) ----
    def matesw_warmup(self, n_pairs: int, len1: int = 640, len2: int = 150) -> None:
        _check(lib().meme_matesw_warmup(self.h, int(n_pairs), int(len1), int(len2)), "meme_matesw_warmup")

    def matesw_batch(self, segments, opt: "MateswOpt | None" = None) -> dict:
        """segments: list of (pairs [SEQPAIR_DTYPE: idr, idq, len1, len2, h0 used], ref bytes, query bytes).
        Returns {"aln": [KSWR_DTYPE array per segment], "ms", "cells"}."""
        opt = opt or default_matesw_opt()
        ns = len(segments)
        ps = [np.ascontiguousarray(s[0], dtype=SEQPAIR_DTYPE) for s in segments]
        rs = [np.ascontiguousarray(s[1], dtype=np.uint8) for s in segments]
        qs = [np.ascontiguousarray(s[2], dtype=np.uint8) for s in segments]
        outs = [np.zeros(len(p), dtype=KSWR_DTYPE) for p in ps]
        ptr = lambda arrs: (C.c_void_p * max(ns, 1))(*[a.ctypes.data for a in arrs])
        cnt = lambda arrs: (C.c_int64 * max(ns, 1))(*[len(a) for a in arrs])
        _check(lib().meme_matesw_batch(self.h, ns, ptr(ps), cnt(ps), ptr(rs), cnt(rs), ptr(qs), cnt(qs), C.byref(opt), ptr(outs)), "meme_matesw_batch")
        ms, cells = C.c_float(), C.c_uint64()
        _check(lib().meme_matesw_last(self.h, C.byref(ms), C.byref(cells)), "meme_matesw_last")
        return {"aln": outs, "ms": ms.value, "cells": cells.value}

    def cigar_warmup(self, n_regions: int, qlen: int = 150) -> None:
        _check(lib().meme_cigar_warmup(self.h, int(n_regions), int(qlen)), "meme_cigar_warmup")

    def cigar_batch(self, tasks: np.ndarray, codes: np.ndarray, l_pac: int, opt: CigarOpt | None = None) -> dict:
        opt = opt or default_cigar_opt()
        tasks = np.ascontiguousarray(tasks, dtype=CIGAR_TASK_DTYPE)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        res = CigarResult()
        _check(lib().meme_cigar_batch(self.h, tasks.ctypes.data, len(tasks), codes.ctypes.data, len(codes), int(l_pac), C.byref(opt), C.byref(res)),
               "meme_cigar_batch")
        n, no, nm = int(res.n_tasks), int(res.n_ops), int(res.n_md)
        ms, cells = C.c_float(), C.c_uint64()
        _check(lib().meme_cigar_last(self.h, C.byref(ms), C.byref(cells)), "meme_cigar_last")
        return {"recs": np.frombuffer(C.string_at(res.recs, n * CIGAR_REC_DTYPE.itemsize), dtype=CIGAR_REC_DTYPE).copy() if n else np.zeros(0, CIGAR_REC_DTYPE),
                "cigars": np.frombuffer(C.string_at(res.cigars, no * 4), dtype=np.uint32).copy() if no else np.zeros(0, np.uint32),
                "md": C.string_at(res.md, nm) if nm else b"", "ms": float(ms.value), "cells": int(cells.value)}

    def chain(self, codes: np.ndarray, offsets: np.ndarray, seed_opt: SeedOpt, chain_opt: ChainOpt) -> dict:
        """Seeding + chaining of a batch through meme_chain_batch_multi on this one device."""
        return chain_batch_multi([self], codes, offsets, seed_opt, chain_opt)[0][0]


def chain_batch_multi(indexes: list, codes: np.ndarray, offsets: np.ndarray, seed_opt: SeedOpt, chain_opt: ChainOpt, pair_aligned: bool = False):
    """meme_chain_batch_multi: returns ([result dict per device], shard_begin)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    nd = len(indexes)
    hs = (C.c_void_p * nd)(*[ix.h for ix in indexes])
    outs = (ChainResult * nd)()
    sb = (C.c_int64 * (nd + 1))()
    _check(lib().meme_chain_batch_multi(hs, nd, codes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, 1 if pair_aligned else 0,
                                        C.byref(seed_opt), C.byref(chain_opt), outs, sb), "meme_chain_batch_multi")
    return [MemeIndex._wrap_chains(outs[d]) for d in range(nd)], [int(x) for x in sb]


class MemeBsw:
    """Host mirror of BandedPairWiseSW (reference src/bandedSWA.h:L118-135): construct once, then
    getScores(pairs, seqBufRef, seqBufQer, numPairs, w)."""

    def __init__(self, opt: BswOpt | None = None, device: int = 0):
        self.opt = opt or default_bsw_opt()
        h = C.c_void_p()
        _check(lib().meme_bsw_create(device, C.byref(h)), "meme_bsw_create")
        self.h = h

    def close(self):
        if self.h:
            lib().meme_bsw_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        _check(lib().meme_bsw_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)), "meme_bsw_set_stream")

    def get_scores(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int) -> np.ndarray:
        out = np.ascontiguousarray(pairs.copy())
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        qer = np.ascontiguousarray(qer, dtype=np.uint8)
        _check(lib().meme_bsw_batch(self.h, out.ctypes.data, len(out), ref.ctypes.data, len(ref), qer.ctypes.data,
                                    len(qer), w, C.byref(self.opt)), "meme_bsw_batch")
        return out

    def get_scores_ext(self, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int, tries: int = 2, prev_is_h0: bool = False):
        """The extension driver's band retries on the device (src/bwamem.cpp:L2968-3376): returns (pairs, w_used)."""
        out = np.ascontiguousarray(pairs.copy())
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        qer = np.ascontiguousarray(qer, dtype=np.uint8)
        wu = np.zeros(len(out), np.int32)
        _check(lib().meme_bsw_batch_ext(self.h, out.ctypes.data, len(out), ref.ctypes.data, len(ref), qer.ctypes.data, len(qer), w,
                                        tries, int(prev_is_h0), C.byref(self.opt), wu.ctypes.data), "meme_bsw_batch_ext")
        return out, wu

    def get_scores_text(self, ix: "MemeIndex", pairs: np.ndarray, qer: np.ndarray, w: int, tries: int = 1, prev_is_h0: bool = False):
        """Reference by position: pairs carry (id, idr) = position in ix's text (id < 0: backwards); no reference upload."""
        out = np.ascontiguousarray(pairs.copy())
        qer = np.ascontiguousarray(qer, dtype=np.uint8)
        vp = C.c_void_p
        L = lib()
        L.meme_bsw_attach_text.argtypes = [vp, vp]
        _check(L.meme_bsw_attach_text(self.h, ix.h), "meme_bsw_attach_text")
        _check(L.meme_bsw_set_match(self.h, self.opt.a), "meme_bsw_set_match")
        pp = (vp * 1)(out.ctypes.data); np_ = (C.c_int64 * 1)(len(out)); pr = (vp * 1)(None); nr = (C.c_int64 * 1)(0)
        pq = (vp * 1)(qer.ctypes.data); nq = (C.c_int64 * 1)(len(qer))
        L.meme_bsw_upload_segments.argtypes = [vp, C.c_int32, vp, vp, vp, vp, vp, vp]
        _check(L.meme_bsw_upload_segments(self.h, 1, pp, np_, pr, nr, pq, nq), "meme_bsw_upload_segments")
        _check(L.meme_bsw_run_ext(self.h, w, tries, int(prev_is_h0), C.byref(self.opt)), "meme_bsw_run_ext")
        _check(L.meme_bsw_download(self.h, out.ctypes.data), "meme_bsw_download")
        _check(L.meme_bsw_attach_text(self.h, None), "meme_bsw_attach_text")
        return out

    def upload(self, pairs, ref, qer):
        _check(lib().meme_bsw_set_match(self.h, self.opt.a), "meme_bsw_set_match")
        _check(lib().meme_bsw_upload(self.h, pairs.ctypes.data, len(pairs), ref.ctypes.data, len(ref), qer.ctypes.data,
                                     len(qer)), "meme_bsw_upload")

    def run(self, w: int):
        _check(lib().meme_bsw_run(self.h, w, C.byref(self.opt)), "meme_bsw_run")

    def download(self, pairs: np.ndarray):
        _check(lib().meme_bsw_download(self.h, pairs.ctypes.data), "meme_bsw_download")

    def cells(self) -> int:
        c = C.c_uint64()
        _check(lib().meme_bsw_cells(self.h, C.byref(c)), "meme_bsw_cells")
        return int(c.value)
