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


def _gather_exact(t: torch.Tensor, root: int, device) -> tuple[torch.Tensor | None, np.ndarray]:
    """Gather 1-D tensors of different lengths to `root` without padding: all_gather of the sizes, then every peer
    sends its tensor straight into its slice of ONE buffer on the root (NCCL send/recv over NVLink on the GPU box, gloo
    in the CPU tests).  Returns (concatenation on root | None, sizes per rank)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = _all_gather_i64([int(t.numel())], device)[:, 0]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    t = t.to(device)
    if rank == root:
        out = torch.empty(int(starts[-1]), dtype=t.dtype, device=device)
        out[int(starts[root]):int(starts[root + 1])] = t
        ops = [dist.P2POp(dist.irecv, out[int(starts[r]):int(starts[r + 1])], r) for r in range(world) if r != root and sizes[r] > 0]
    else:
        out = None
        ops = [dist.P2POp(dist.isend, t, root)] if t.numel() > 0 else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out, sizes


def gather_records(parts: dict, root: int = 0, device="cpu") -> tuple[dict | None, int]:
    """The final hit gather: every rank's record arrays -- `smems` (raw bytes of the 24-byte records),
    `hits` (int64), `smem_off`, `hit_off` (int64, n+1 entries) as 1-D torch tensors ALREADY ON `device` -- concatenated on
    `root` in rank order = read order, offsets rebased there.  No host bounce and no padding: device-resident tensors
    travel device to device.  Returns (dict of tensors on root | None, bytes this rank sent or received)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return parts, 0
    rank = dist.get_rank()
    moved = 0
    out = {}
    sizes = {}
    for k in ("smems", "hits", "smem_off", "hit_off"):
        out[k], sizes[k] = _gather_exact(parts[k], root, device)
        esz = parts[k].element_size()
        moved += int((sizes[k].sum() - sizes[k][root]) * esz) if rank == root else int(parts[k].numel() * esz)
    if rank != root:
        return None, moved
    # offsets: drop the leading 0 of every shard but the first, add the running totals
    world = dist.get_world_size()
    for k in ("smem_off", "hit_off"):
        st = np.concatenate([[0], np.cumsum(sizes[k])])
        pieces, base = [], 0
        for r in range(world):
            seg = out[k][int(st[r]):int(st[r + 1])]
            pieces.append(seg + base if r == 0 else seg[1:] + base)
            base = base + int(seg[-1]) if seg.numel() else base
        out[k] = torch.cat(pieces)
    return out, moved


def device_results(ix, n_reads: int, device) -> dict:
    """The lists of the last meme_seed_run as torch tensors on `device` (device-to-device copies, no host bounce)."""
    ns, nh = ix.seed_totals()
    parts = {"smems": torch.empty(ns * 24, dtype=torch.uint8, device=device), "hits": torch.empty(nh, dtype=torch.int64, device=device),
             "smem_off": torch.empty(n_reads + 1, dtype=torch.int64, device=device),
             "hit_off": torch.empty(n_reads + 1, dtype=torch.int64, device=device)}
    ix.seed_copy_device(parts["smems"].data_ptr(), parts["hits"].data_ptr(), parts["smem_off"].data_ptr(), parts["hit_off"].data_ptr())
    return parts


def gather_seed_results(res: dict, root: int = 0, device="cpu") -> dict | None:
    """Host-array form of gather_records (the dict MemeIndex.seed returns, for that rank's contiguous shard): used by
    the CPU tests (gloo).  Returns None on the other ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return res
    from . import capi
    parts = {"smems": torch.from_numpy(res["smems"].view(np.uint8).reshape(-1).copy()),
             "hits": torch.from_numpy(res["hits"].view(np.int64).copy()),
             "smem_off": torch.from_numpy(res["smem_off"].copy()), "hit_off": torch.from_numpy(res["hit_off"].copy())}
    out, _ = gather_records(parts, root, device)
    if out is None:
        return None
    return {"smems": out["smems"].cpu().numpy().view(capi.SMEM_DTYPE), "hits": out["hits"].cpu().numpy().view(np.uint64),
            "smem_off": out["smem_off"].cpu().numpy(), "hit_off": out["hit_off"].cpu().numpy()}
