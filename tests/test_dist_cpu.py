"""world_size-2 gloo test of the sharding + final gather (bwa-meme_b200/dist.py) on CPU.

The per-rank engine is stood in for by the oracle (there is no GPU here); what is under test is the host
logic: contiguous pair-aligned shards, rank-ordered concatenation, offset rebasing."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, prefix, codes, offs, paired, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_pkg()
    O = ge.load_oracle()
    from bwa_meme_b200 import dist as mdist, capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c, o, (b, e) = mdist.shard_reads(codes, offs, world, rank, paired)
    oix = O.OracleIndex(prefix, 3)
    r = oix.seed(c, o)
    oix.close()
    sm = np.zeros(len(r["smems"]), capi.SMEM_DTYPE)
    for f in ("start", "end", "hitbeg", "hitcount"):
        sm[f] = r["smems"][f]
    res = {"smems": sm, "hits": r["hits"], "smem_off": r["smem_off"], "hit_off": r["hit_off"]}
    full = mdist.gather_seed_results(res, root=0)
    if rank == 0:
        q.put((full["smems"].tobytes(), full["hits"].tobytes(), full["smem_off"].tobytes(), full["hit_off"].tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("paired", [False, True])
def test_sharded_equals_whole(oracle, pkg, golden_index, golden_reads, paired):
    from bwa_meme_b200 import capi
    codes, offs = golden_reads[150]
    codes, offs = codes[:offs[151]], offs[:152]          # 151 reads: odd count exercises the pair alignment
    oix = oracle.OracleIndex(golden_index, 3)
    want = oix.seed(codes, offs)
    oix.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if paired else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, golden_index, codes, offs, paired, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    smems = np.frombuffer(got[0], dtype=capi.SMEM_DTYPE)
    hits = np.frombuffer(got[1], dtype=np.uint64)
    so = np.frombuffer(got[2], dtype=np.int64)
    ho = np.frombuffer(got[3], dtype=np.int64)
    assert (so == want["smem_off"]).all() and (ho == want["hit_off"]).all()
    assert (hits == want["hits"]).all()
    for f in ("start", "end", "hitbeg", "hitcount"):
        assert (smems[f] == want["smems"][f]).all()


def test_shard_ranges_properties(pkg):
    from bwa_meme_b200 import dist as mdist
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 4, 8):
            for paired in (False, True):
                rs = mdist.shard_ranges(n, world, paired)
                assert rs[0][0] == 0 and rs[-1][1] == n
                for (a, b), (c, d) in zip(rs, rs[1:]):
                    assert b == c and a <= b
                if paired:
                    assert all(b % 2 == 0 for _, b in rs[:-1])
