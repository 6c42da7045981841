"""Multi-GPU plumbing of the hot path: reads are independent units, so a chunk is cut
into `world` contiguous read ranges (mates kept together), the index is replicated in every GPU's HBM and
NO collective sits on the data path.  The single exchange step is the final gather of per-read result
records to the rank that feeds the host stage; it is written against torch.distributed so that NCCL
(NVLink/NVSwitch) carries it on the GPU box and gloo carries it in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(n_reads: int, world: int, paired: bool = False) -> list[tuple[int, int]]:
    """Contiguous [begin, end) read ranges, one per rank; with paired=True boundaries fall on even
    indices so that both mates of a pair land on the same rank (interleaved R1/R2 order as in
    bseq_read_orig, reference src/fastmap.cpp:L743)."""
    unit = 2 if paired else 1
    units = n_reads // unit
    out = []
    for r in range(world):
        b = (units * r // world) * unit
        e = (units * (r + 1) // world) * unit if r < world - 1 else n_reads
        out.append((b, e))
    return out


def shard_reads(codes: np.ndarray, offsets: np.ndarray, world: int, rank: int, paired: bool = False):
    b, e = shard_ranges(len(offsets) - 1, world, paired)[rank]
    off = offsets[b:e + 1] - offsets[b]
    return codes[offsets[b]:offsets[e]], off, (b, e)


def _all_gather_i64(vals: list[int], device) -> np.ndarray:
    t = torch.tensor(vals, dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        return torch.stack(out).cpu().numpy()
    return t.cpu().numpy()[None, :]


def gather_summary(ix, n_reads: int, world: int, device) -> dict:
    """Per-rank totals of the last seeding run gathered on every rank (the cheap form of the final
    hit gather: counts only)."""
    ns, nh = ix.seed_totals()
    m = _all_gather_i64([n_reads, ns, nh], device)
    return {"reads": int(m[:, 0].sum()), "smems": int(m[:, 1].sum()), "hits": int(m[:, 2].sum()),
            "per_rank": m.tolist()}


def _gather_var(t: torch.Tensor, root: int, device) -> list[torch.Tensor] | None:
    """Gather 1-D tensors of different lengths to `root` (all_gather of sizes + padded gather)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = _all_gather_i64([int(t.numel())], device)[:, 0]
    mx = int(sizes.max())
    pad = torch.zeros(mx, dtype=t.dtype, device=device)
    pad[:t.numel()] = t.to(device)
    bufs = [torch.empty(mx, dtype=t.dtype, device=device) for _ in range(world)] if rank == root else None
    dist.gather(pad, bufs, dst=root)
    if rank != root:
        return None
    return [bufs[r][:int(sizes[r])] for r in range(world)]


def gather_seed_results(res: dict, root: int = 0, device="cpu") -> dict | None:
    """Concatenate every rank's seeding result (the dict MemeIndex.seed returns, for that rank's
    contiguous shard) on `root`, in rank order = read order.  Returns None on the other ranks.
    Record layouts travel as raw bytes; offsets are rebased on the root."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return res
    from . import capi
    parts = {}
    parts["smems"] = _gather_var(torch.from_numpy(res["smems"].view(np.uint8).reshape(-1).copy()), root, device)
    parts["hits"] = _gather_var(torch.from_numpy(res["hits"].view(np.int64).copy()), root, device)
    parts["smem_off"] = _gather_var(torch.from_numpy(res["smem_off"].copy()), root, device)
    parts["hit_off"] = _gather_var(torch.from_numpy(res["hit_off"].copy()), root, device)
    if dist.get_rank() != root:
        return None
    smems = np.concatenate([p.cpu().numpy().view(capi.SMEM_DTYPE) for p in parts["smems"]])
    hits = np.concatenate([p.cpu().numpy().view(np.uint64) for p in parts["hits"]])
    so, ho = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    sb = hb = 0
    for a, b in zip(parts["smem_off"], parts["hit_off"]):
        a, b = a.cpu().numpy(), b.cpu().numpy()
        so.append(a[1:] + sb)
        ho.append(b[1:] + hb)
        sb += int(a[-1])
        hb += int(b[-1])
    return {"smems": smems, "hits": hits, "smem_off": np.concatenate(so), "hit_off": np.concatenate(ho)}
