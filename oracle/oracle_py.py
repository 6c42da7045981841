"""ctypes bindings for the CPU oracle (oracle/seed_oracle.c, oracle/bsw_oracle.c, oracle/chain_oracle.c, oracle/cigar_oracle.c,
oracle/matesw_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(force: bool = False) -> None:
    os.makedirs(BUILD, exist_ok=True)
    for src, lib in (("seed_oracle.c", "liboracle_seed.so"), ("bsw_oracle.c", "liboracle_bsw.so"), ("chain_oracle.c", "liboracle_chain.so"), ("cigar_oracle.c", "liboracle_cigar.so"),
                     ("matesw_oracle.c", "liboracle_matesw.so")):
        s, o = os.path.join(HERE, src), os.path.join(BUILD, lib)
        if not os.path.exists(s):
            continue
        if force or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(s):
            subprocess.run(["gcc", "-O2", "-g", "-Wall", "-shared", "-fPIC", "-o", o, s, "-lm"], check=True)


class SeedOpt(C.Structure):
    _fields_ = [("min_seed_len", C.c_int), ("split_len", C.c_int), ("split_width", C.c_int),
                ("max_mem_intv", C.c_int), ("rounds", C.c_int)]


class SeedResult(C.Structure):
    _fields_ = [("n_smems", C.c_int64), ("n_hits", C.c_int64), ("smems", C.c_void_p),
                ("smem_off", C.POINTER(C.c_int64)), ("hit_off", C.POINTER(C.c_int64)),
                ("hits", C.POINTER(C.c_uint64)), ("counters", C.c_uint64 * 8)]


SMEM_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("hitbeg", "<i4"), ("hitcount", "<i4"), ("sa_lo", "<u8")])
COUNTER_NAMES = ("Lk", "Lp", "P", "W", "I", "H", "S", "searches")


def default_opt(rounds: int = 3, min_seed_len: int = 19, split_factor: float = 1.5, split_width: int = 10,
                max_mem_intv: int = 20) -> SeedOpt:
    return SeedOpt(min_seed_len, int(min_seed_len * split_factor + .499), split_width, max_mem_intv, rounds)


_seed_lib = None


def seed_lib():
    global _seed_lib
    if _seed_lib is None:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_seed.so"))
        L.oracle_index_load.restype = C.c_void_p
        L.oracle_index_load.argtypes = [C.c_char_p, C.c_int]
        L.oracle_index_free.argtypes = [C.c_void_p]
        L.oracle_index_sa_num.restype = C.c_uint64
        L.oracle_index_sa_num.argtypes = [C.c_void_p]
        L.oracle_seed_reads.restype = C.POINTER(SeedResult)
        L.oracle_seed_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(SeedOpt)]
        L.oracle_seed_result_free.argtypes = [C.POINTER(SeedResult)]
        _seed_lib = L
    return _seed_lib


class OracleIndex:
    def __init__(self, prefix: str, mode: int = 3):
        self.lib = seed_lib()
        self.h = self.lib.oracle_index_load(prefix.encode(), mode)
        if not self.h:
            raise FileNotFoundError(f"oracle could not load index {prefix}")
        self.sa_num = int(self.lib.oracle_index_sa_num(self.h))
        self.mode = mode

    def close(self):
        if self.h:
            self.lib.oracle_index_free(self.h)
            self.h = None

    def rmi_lookup(self, keys: np.ndarray):
        """learned_index_lookup (src/LearnedIndex_seeding.cpp:L186-210) for a batch of keys: (prediction, error word)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        pred = np.zeros(len(keys), np.uint64)
        err = np.zeros(len(keys), np.uint64)
        self.lib.oracle_rmi_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.oracle_rmi_lookup.restype = None
        self.lib.oracle_rmi_lookup(self.h, keys.ctypes.data, len(keys), pred.ctypes.data, err.ctypes.data)
        return pred, err

    def seed(self, codes: np.ndarray, offsets: np.ndarray, opt: SeedOpt | None = None) -> dict:
        """Returns dict(smems=record array, smem_off, hit_off, hits, counters)."""
        opt = opt or default_opt()
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = offsets.shape[0] - 1
        rp = self.lib.oracle_seed_reads(self.h, codes.ctypes.data, offsets.ctypes.data, n, C.byref(opt))
        r = rp.contents
        ns, nh = int(r.n_smems), int(r.n_hits)
        smems = np.frombuffer(C.string_at(r.smems, ns * SMEM_DTYPE.itemsize), dtype=SMEM_DTYPE).copy() if ns else np.zeros(0, SMEM_DTYPE)
        out = {
            "smems": smems,
            "smem_off": np.ctypeslib.as_array(r.smem_off, shape=(n + 1,)).copy(),
            "hit_off": np.ctypeslib.as_array(r.hit_off, shape=(n + 1,)).copy(),
            "hits": np.ctypeslib.as_array(r.hits, shape=(nh,)).copy() if nh else np.zeros(0, np.uint64),
            "counters": {k: int(r.counters[i]) for i, k in enumerate(COUNTER_NAMES)},
        }
        self.lib.oracle_seed_result_free(rp)
        return out


def bytes_seed(counters: dict, total_read_bases: int, mode: int = 3) -> int:
    """Algorithmic bytes of the seeding path, SURVEY.md §8d:
    R + 24(Lk+Lp) + SASIZE*P + 8W + 5I + 8H + 24S."""
    sasize = 5 if mode == 1 else 13
    c = counters
    return (total_read_bases + 24 * (c["Lk"] + c["Lp"]) + sasize * c["P"] + 8 * c["W"] + 5 * c["I"]
            + 8 * c["H"] + 24 * c["S"])


# ---- canonical text form shared by the oracle, the CUDA path and the reference harness dump ----
def canon_lines(smems: np.ndarray, smem_off: np.ndarray, hit_off: np.ndarray, hits: np.ndarray, n_reads: int):
    """Per read, a sorted list of (start, end, tuple(hits)) -- order-insensitive to ties."""
    out = []
    for i in range(n_reads):
        a, b = int(smem_off[i]), int(smem_off[i + 1])
        hb = int(hit_off[i])
        rec = []
        for k in range(a, b):
            s = smems[k]
            h0 = hb + int(s["hitbeg"])
            rec.append((int(s["start"]), int(s["end"]), tuple(int(x) for x in hits[h0:h0 + int(s["hitcount"])])))
        rec.sort()
        out.append(rec)
    return out


def parse_ref_dump(path: str, n_reads: int):
    """Parse the steps>=4 dump of test/Learned_seeding_big_read.cpp:L288-300:
    `rid:` line, then `[start,end] [pos,pos,...,]` per SMEM."""
    out = [[] for _ in range(n_reads)]
    cur = None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.endswith(":") and line[:-1].isdigit():
                cur = int(line[:-1])
                continue
            if line.startswith("[") and cur is not None:
                a, b = line.split("] [")
                s, e = a[1:].split(",")
                hs = tuple(int(x) for x in b.rstrip("]").split(",") if x)
                out[cur].append((int(s), int(e), hs))
    for rec in out:
        rec.sort()
    return out


# ------------------------------------------------------------------------------------------------
# banded Smith-Waterman oracle
# ------------------------------------------------------------------------------------------------
SEQPAIR_DTYPE = np.dtype([("idr", "<i4"), ("idq", "<i4"), ("id", "<i4"), ("len1", "<i4"), ("len2", "<i4"),
                          ("h0", "<i4"), ("seqid", "<i4"), ("regid", "<i4"), ("score", "<i4"), ("tle", "<i4"),
                          ("gtle", "<i4"), ("qle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
assert SEQPAIR_DTYPE.itemsize == 56
RESULT_FIELDS = ("score", "tle", "gtle", "qle", "gscore", "max_off")


class BswOpt(C.Structure):
    _fields_ = [("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32),
                ("zdrop", C.c_int32), ("end_bonus", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


def default_bsw_opt() -> BswOpt:
    """mem_opt_init defaults, src/bwamem.cpp:L126-162."""
    return BswOpt(6, 1, 6, 1, 100, 5, 1, 4)


_bsw_lib = None


def bsw_lib():
    global _bsw_lib
    if _bsw_lib is None:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_bsw.so"))
        L.oracle_bsw_batch.argtypes = [C.POINTER(BswOpt), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _bsw_lib = L
    return _bsw_lib


def bsw_batch(pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int, opt: BswOpt | None = None) -> np.ndarray:
    """Run the oracle over `pairs` (SEQPAIR_DTYPE); returns a copy with the six result fields filled."""
    opt = opt or default_bsw_opt()
    out = np.ascontiguousarray(pairs.copy())
    ref = np.ascontiguousarray(ref, dtype=np.uint8)
    qer = np.ascontiguousarray(qer, dtype=np.uint8)
    bsw_lib().oracle_bsw_batch(C.byref(opt), out.ctypes.data, ref.ctypes.data, qer.ctypes.data, len(out), w)
    return out


def bsw_cells() -> int:
    return int(C.c_uint64.in_dll(bsw_lib(), "oracle_bsw_cells").value)


def make_pairs(n: int, seed: int = 1, max_q: int = 150, max_extra_ref: int = 110, sub: float = 0.03,
               indel: float = 0.2, with_n: float = 0.02, h0_max: int = 120, lens=None):
    """Seeded extension-like pairs: query = mutated prefix of the reference window (as the left/right
    extension pairs of mem_chain2aln_across_reads_V2 are), h0 = a seed score."""
    rng = np.random.default_rng(seed)
    pairs = np.zeros(n, dtype=SEQPAIR_DTYPE)
    refs, qers = [], []
    ro = qo = 0
    for i in range(n):
        ql = int(rng.integers(1, max_q + 1)) if lens is None else int(lens[i][0])
        tl = ql + int(rng.integers(0, max_extra_ref + 1)) if lens is None else int(lens[i][1])
        t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        kind = rng.random()
        if kind < 0.1:
            q = rng.integers(0, 4, size=ql, dtype=np.uint8)          # unrelated
        else:
            q = t[:ql].copy() if tl >= ql else np.concatenate([t, rng.integers(0, 4, size=ql - tl, dtype=np.uint8)])
            m = rng.random(ql) < sub
            q[m] = (q[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            if rng.random() < indel and ql > 12:
                p = int(rng.integers(3, ql - 6)); k = int(rng.integers(1, 6))
                if rng.random() < 0.5:
                    q = np.concatenate([q[:p], q[p + k:], rng.integers(0, 4, size=k, dtype=np.uint8)])[:ql]
                else:
                    q = np.concatenate([q[:p], rng.integers(0, 4, size=k, dtype=np.uint8), q[p:]])[:ql]
            if kind > 0.9:   # diverge after a while (exercises z-drop)
                p = int(rng.integers(0, ql))
                q[p:] = rng.integers(0, 4, size=ql - p, dtype=np.uint8)
        if rng.random() < with_n:
            q[int(rng.integers(0, ql))] = 4
        if rng.random() < with_n:
            t[int(rng.integers(0, tl))] = 4
        pairs[i]["idr"], pairs[i]["idq"], pairs[i]["id"] = ro, qo, i
        pairs[i]["len1"], pairs[i]["len2"] = tl, ql
        pairs[i]["h0"] = int(rng.integers(1, h0_max + 1))
        pairs[i]["seqid"], pairs[i]["regid"] = i, 0
        refs.append(t); qers.append(q)
        ro += tl; qo += ql
    return pairs, np.concatenate(refs), np.concatenate(qers)


def run_ref_bsw_shim(binary: str, pairs: np.ndarray, ref: np.ndarray, qer: np.ndarray, w: int, mode: str,
                     opt: BswOpt | None = None, tmpdir: str = "/tmp") -> np.ndarray:
    """Run oracle/_ref/bsw_shim_* (the unmodified reference class) and return pairs with results."""
    opt = opt or default_bsw_opt()
    fin = os.path.join(tmpdir, f"bsw_in_{os.getpid()}.bin")
    fout = os.path.join(tmpdir, f"bsw_out_{os.getpid()}.bin")
    hdr = np.array([len(pairs), w, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.zdrop, opt.end_bonus,
                    opt.a, opt.b, len(ref), len(qer)], dtype=np.int32)
    with open(fin, "wb") as f:
        f.write(hdr.tobytes()); f.write(pairs.tobytes()); f.write(ref.tobytes()); f.write(qer.tobytes())
    subprocess.run([binary, fin, fout, mode], check=True, capture_output=True)
    res = np.fromfile(fout, dtype=np.int32).reshape(-1, 6)
    out = pairs.copy()
    for k, name in enumerate(RESULT_FIELDS):
        out[name] = res[:, k]
    os.remove(fin); os.remove(fout)
    return out


# ---- chaining oracle (oracle/chain_oracle.c) ----------------------------------------------------------------------------------
class ChainOpt(C.Structure):
    _fields_ = [("w", C.c_int32), ("max_chain_gap", C.c_int32), ("max_occ", C.c_int32), ("min_seed_len", C.c_int32),
                ("min_chain_weight", C.c_int32), ("max_chain_extend", C.c_int32), ("mask_level", C.c_float), ("drop_ratio", C.c_float)]


class ChainResult(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_chains", C.c_int64), ("n_seeds", C.c_int64), ("chain_off", C.POINTER(C.c_int64)),
                ("frac_rep", C.POINTER(C.c_float)), ("chains", C.c_void_p), ("seeds", C.c_void_p)]


CONTIG_DTYPE = np.dtype([("offset", "<i8"), ("len", "<i4"), ("is_alt", "<i4")])
CHAIN_DTYPE = np.dtype([("pos", "<i8"), ("seed_beg", "<i8"), ("n", "<i4"), ("rid", "<i4"), ("w", "<i4"), ("first", "<i4"),
                        ("kept", "<i4"), ("is_alt", "<i4")])
CHAIN_SEED_DTYPE = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4")])


def default_chain_opt(w: int = 100, max_chain_gap: int = 10000, max_occ: int = 500, min_seed_len: int = 19, min_chain_weight: int = 0,
                      max_chain_extend: int = 1 << 30, mask_level: float = 0.5, drop_ratio: float = 0.5) -> ChainOpt:
    """mem_opt_init defaults (src/bwamem.cpp:L126-162)."""
    return ChainOpt(w, max_chain_gap, max_occ, min_seed_len, min_chain_weight, max_chain_extend, mask_level, drop_ratio)


_chain_lib = None


def chain_lib():
    global _chain_lib
    if _chain_lib is None:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_chain.so"))
        L.oracle_chain_reads.restype = C.POINTER(ChainResult)
        L.oracle_chain_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                         C.c_int64, C.POINTER(ChainOpt)]
        L.oracle_chain_result_free.argtypes = [C.POINTER(ChainResult)]
        _chain_lib = L
    return _chain_lib


def chain_reads(seed_res: dict, read_len: np.ndarray, contigs: np.ndarray, l_pac: int, opt: ChainOpt | None = None) -> dict:
    """mem_chain_Learned + mem_chain_flt for every read of a seeding result with FULL hit lists (the dict OracleIndex.seed
    returns; hitbeg relative to the read's list)."""
    opt = opt or default_chain_opt()
    L = chain_lib()
    sm = np.ascontiguousarray(seed_res["smems"])
    so = np.ascontiguousarray(seed_res["smem_off"], dtype=np.int64)
    hi = np.ascontiguousarray(seed_res["hits"], dtype=np.uint64)
    ho = np.ascontiguousarray(seed_res["hit_off"], dtype=np.int64)
    rl = np.ascontiguousarray(read_len, dtype=np.int32)
    ct = np.ascontiguousarray(contigs, dtype=CONTIG_DTYPE)
    assert sm.dtype.itemsize == 24
    n = len(rl)
    r = L.oracle_chain_reads(sm.ctypes.data, so.ctypes.data, hi.ctypes.data, ho.ctypes.data, rl.ctypes.data, n, ct.ctypes.data, len(ct),
                             int(l_pac), C.byref(opt))
    R = r.contents
    nc, nsd = int(R.n_chains), int(R.n_seeds)
    out = {"chain_off": np.ctypeslib.as_array(R.chain_off, shape=(n + 1,)).copy(),
           "frac_rep": np.ctypeslib.as_array(R.frac_rep, shape=(max(n, 1),))[:n].copy(),
           "chains": np.frombuffer(C.string_at(R.chains, nc * CHAIN_DTYPE.itemsize), dtype=CHAIN_DTYPE).copy() if nc else np.zeros(0, CHAIN_DTYPE),
           "seeds": np.frombuffer(C.string_at(R.seeds, nsd * CHAIN_SEED_DTYPE.itemsize), dtype=CHAIN_SEED_DTYPE).copy() if nsd else np.zeros(0, CHAIN_SEED_DTYPE)}
    L.oracle_chain_result_free(r)
    return out


def chain_lines(res: dict, names=None) -> list[str]:
    """Canonical text form of a chaining result, one line per read -- the format MEME_B200_DUMP_CHAINS writes
    (bwa-meme_b200/host/meme_glue.cpp: meme_glue_dump_chains)."""
    out = []
    co, ch, sd, fr = res["chain_off"], res["chains"], res["seeds"], res["frac_rep"]
    for r in range(len(co) - 1):
        parts = [names[r] if names is not None else str(r), f"{len(range(int(co[r]), int(co[r + 1])))}"]
        for c in ch[int(co[r]):int(co[r + 1])]:
            seeds = sd[int(c["seed_beg"]):int(c["seed_beg"]) + int(c["n"])]
            frac = fr[r]
            parts.append(f"[pos={int(c['pos'])} rid={int(c['rid'])} n={int(c['n'])} w={int(c['w'])} kept={int(c['kept'])} alt={int(c['is_alt'])} "
                         f"first={int(c['first'])} frac={float(frac):.6f} " +
                         " ".join(f"{int(s['qbeg'])},{int(s['len'])},{int(s['rbeg'])}" for s in seeds) + "]")
        out.append(" ".join(parts))
    return out


# ---- SAM-stage alignment oracle (oracle/cigar_oracle.c) -------------------------------------------------------------------------
class CigarOpt(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("pad", C.c_int8 * 3), ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32),
                ("e_ins", C.c_int32), ("w", C.c_int32), ("a", C.c_int32)]


class CigarResult(C.Structure):
    _fields_ = [("n_tasks", C.c_int64), ("n_ops", C.c_int64), ("n_md", C.c_int64), ("recs", C.c_void_p), ("cigars", C.c_void_p),
                ("md", C.c_void_p)]


CIGAR_TASK_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("q_off", "<i8"), ("qlen", "<i4"), ("w", "<i4"), ("truesc", "<i4"), ("pad", "<i4")])
CIGAR_REC_DTYPE = np.dtype([("score", "<i4"), ("nm", "<i4"), ("n_cigar", "<i4"), ("md_len", "<i4"), ("cigar_off", "<i8"), ("md_off", "<i8")])


def default_cigar_opt(a: int = 1, b: int = 4, o_del: int = 6, e_del: int = 1, o_ins: int = 6, e_ins: int = 1, w: int = 100) -> CigarOpt:
    """bwa_fill_scmat (src/bwa.cpp) + mem_opt_init defaults."""
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


_cigar_lib = None


def cigar_lib():
    global _cigar_lib
    if _cigar_lib is None:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_cigar.so"))
        L.oracle_cigar_regions.restype = C.POINTER(CigarResult)
        L.oracle_cigar_regions.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(CigarOpt)]
        L.oracle_cigar_result_free.argtypes = [C.POINTER(CigarResult)]
        L.oracle_infer_w.restype = C.c_int
        L.oracle_infer_w.argtypes = [C.c_int] * 10
        _cigar_lib = L
    return _cigar_lib


def infer_w(qlen: int, tlen: int, truesc: int, reg_w: int, opt: CigarOpt | None = None) -> int:
    opt = opt or default_cigar_opt()
    return int(cigar_lib().oracle_infer_w(qlen, tlen, truesc, opt.a, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.w, reg_w))


def unpack_cigar_result(recs: np.ndarray, cigars: np.ndarray, md: bytes) -> list[dict]:
    out = []
    for r in recs:
        n = int(r["n_cigar"])
        if n < 0:
            out.append({"n_cigar": -1})
            continue
        c0, m0 = int(r["cigar_off"]), int(r["md_off"])
        out.append({"score": int(r["score"]), "nm": int(r["nm"]), "n_cigar": n, "cigar": [int(x) for x in cigars[c0:c0 + n]],
                    "md": md[m0:m0 + int(r["md_len"]) - 1].decode()})
    return out


def cigar_regions(text: np.ndarray, l_pac: int, codes: np.ndarray, tasks: np.ndarray, opt: CigarOpt | None = None) -> list[dict]:
    """The band-doubling loop of mem_reg2aln around bwa_gen_cigar2 for every task; text = the .0123 codes."""
    opt = opt or default_cigar_opt()
    L = cigar_lib()
    text = np.ascontiguousarray(text, dtype=np.uint8)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    tasks = np.ascontiguousarray(tasks, dtype=CIGAR_TASK_DTYPE)
    r = L.oracle_cigar_regions(text.ctypes.data, int(l_pac), codes.ctypes.data, tasks.ctypes.data, len(tasks), C.byref(opt))
    R = r.contents
    n, no, nm = int(R.n_tasks), int(R.n_ops), int(R.n_md)
    recs = np.frombuffer(C.string_at(R.recs, n * CIGAR_REC_DTYPE.itemsize), dtype=CIGAR_REC_DTYPE).copy() if n else np.zeros(0, CIGAR_REC_DTYPE)
    cig = np.frombuffer(C.string_at(R.cigars, no * 4), dtype=np.uint32).copy() if no else np.zeros(0, np.uint32)
    md = C.string_at(R.md, nm) if nm else b""
    L.oracle_cigar_result_free(r)
    return unpack_cigar_result(recs, cig, md)


def sam_cigar(raw: dict, qb: int, qe: int, l_query: int, rev: bool) -> dict:
    """What mem_reg2aln does to bwa_gen_cigar2's CIGAR before it is printed (src/bwamem.cpp:L2354-2379): squeeze a leading or
    trailing deletion, add the soft clips.  Returns {"cigar", "md", "nm", "pos_shift"}."""
    cig = list(raw["cigar"])
    shift = 0
    if cig:
        if (cig[0] & 0xf) == 2:
            shift = cig[0] >> 4
            cig = cig[1:]
        elif (cig[-1] & 0xf) == 2:
            cig = cig[:-1]
    if qb != 0 or qe != l_query:
        clip5 = l_query - qe if rev else qb
        clip3 = qb if rev else l_query - qe
        if clip5:
            cig = [clip5 << 4 | 3] + cig
        if clip3:
            cig = cig + [clip3 << 4 | 3]
    return {"cigar": cig, "md": raw["md"], "nm": raw["nm"], "pos_shift": shift}


# ---- mate-rescue Smith-Waterman oracle (oracle/matesw_oracle.c) ------------------------------------------------------------------
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000
MATESW_REC_DTYPE = np.dtype([("r_off", "<i8"), ("q_off", "<i8"), ("len1", "<i4"), ("len2", "<i4"), ("xtra", "<i4"), ("pad", "<i4")])
KSWR_DTYPE = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4")])


class MateswOpt(C.Structure):
    _fields_ = [("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


def default_matesw_opt(a: int = 1, b: int = 4, o_del: int = 6, e_del: int = 1, o_ins: int = 6, e_ins: int = 1) -> MateswOpt:
    return MateswOpt(o_del, e_del, o_ins, e_ins, a, b)


_matesw_lib = None


def matesw_lib():
    global _matesw_lib
    if _matesw_lib is None:
        build()
        L = C.CDLL(os.path.join(BUILD, "liboracle_matesw.so"))
        L.oracle_matesw.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(MateswOpt), C.c_void_p]
        _matesw_lib = L
    return _matesw_lib


def matesw_batch(recs: np.ndarray, ref: np.ndarray, qer: np.ndarray, opt: MateswOpt | None = None) -> np.ndarray:
    """mem_sam_pe_batch (kswv phases 0 and 1) for every pair; returns KSWR_DTYPE records."""
    opt = opt or default_matesw_opt()
    recs = np.ascontiguousarray(recs, dtype=MATESW_REC_DTYPE)
    ref = np.ascontiguousarray(ref, dtype=np.uint8); qer = np.ascontiguousarray(qer, dtype=np.uint8)
    out = np.zeros(len(recs), dtype=KSWR_DTYPE)
    matesw_lib().oracle_matesw(recs.ctypes.data, len(recs), ref.ctypes.data, qer.ctypes.data, C.byref(opt), out.ctypes.data)
    return out


def matesw_write_cases(path: str, recs: np.ndarray, ref: np.ndarray, qer: np.ndarray, opt: MateswOpt | None = None) -> None:
    """The input file of oracle/ref_matesw_shim.cpp."""
    opt = opt or default_matesw_opt()
    with open(path, "wb") as f:
        f.write(np.array([len(recs), len(ref), len(qer)], np.int64).tobytes())
        f.write(np.array([opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.a, opt.b, 0, 0], np.int32).tobytes())
        f.write(np.ascontiguousarray(recs, dtype=MATESW_REC_DTYPE).tobytes())
        f.write(np.ascontiguousarray(ref, np.uint8).tobytes()); f.write(np.ascontiguousarray(qer, np.uint8).tobytes())
