"""Multi-GPU host logic: one process per GPU, windows sharded, database replicated.

Windows are the reference's unit of parallelism (one pool task per window, Plugins/GetVariants.java:138-150); they
share bases (k-1 at tiling boundaries) but never results, so the path shards without a data-path collective: every
rank screens a contiguous range of the window list, balanced on the number of positions, and the per-window rows are
concatenated in window order (the KCF writer wants .faidx order, GetVariants.java:169-179).  The only communication
is the gather of the 48-byte result rows.
"""
from __future__ import annotations

import numpy as np

from ._lib import RESULT_DTYPE, SEGMENT_DTYPE, WINDOW_DTYPE


def window_lengths(wins: np.ndarray, segs: np.ndarray) -> np.ndarray:
    """positions per window = Σ segment lengths (what the screening kernel walks)"""
    seg_len = segs["len"].astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(seg_len)])
    first = wins["first_seg"].astype(np.int64)
    return csum[first + wins["n_segs"].astype(np.int64)] - csum[first]


def partition(lengths: np.ndarray, world: int) -> list[tuple[int, int]]:
    """contiguous window ranges [begin, end) per rank with near-equal Σ length; every window in exactly one range"""
    n = int(lengths.size)
    csum = np.concatenate([[0], np.cumsum(lengths.astype(np.int64))])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        c = int(np.searchsorted(csum, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def grid_layout(rank: int, world: int, table_parts: int) -> tuple[int, int, list[int]]:
    """scan placement on world = table_parts x window_shards ranks: rank r keeps table slice r % table_parts and screens
    window shard r // table_parts; the ranks that share a window shard (consecutive ranks, so neighbours on the NVSwitch)
    reduce their bitmaps among themselves.  Returns (slice, shard, ranks of this rank's reduction group)."""
    if table_parts < 1 or world % table_parts:
        raise ValueError(f"table_parts {table_parts} must divide the world size {world}")
    shard_id = rank // table_parts
    return rank % table_parts, shard_id, list(range(shard_id * table_parts, (shard_id + 1) * table_parts))


def local_slice(wins: np.ndarray, segs: np.ndarray, begin: int, end: int) -> tuple[np.ndarray, np.ndarray]:
    """window / segment arrays of one rank's range, segment indices re-based"""
    w = np.ascontiguousarray(wins[begin:end]).astype(WINDOW_DTYPE, copy=True)
    if w.size == 0:
        return w, np.zeros(0, SEGMENT_DTYPE)
    s0 = int(w["first_seg"][0])
    s1 = int(w["first_seg"][-1]) + int(w["n_segs"][-1])
    w["first_seg"] -= np.uint32(s0)
    return w, np.ascontiguousarray(segs[s0:s1])


def screen_sharded(screen_fn, wins: np.ndarray, segs: np.ndarray, group=None) -> np.ndarray:
    """screen_fn(wins, segs) -> RESULT_DTYPE rows for the given windows (Context.screen bound to a database on this
    rank's GPU).  Returns all rows, in window order, on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return screen_fn(wins, segs)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ranges = partition(window_lengths(wins, segs), world)
    b, e = ranges[rank]
    lw, ls = local_slice(wins, segs, b, e)
    mine = screen_fn(lw, ls) if lw.size else np.zeros(0, RESULT_DTYPE)
    assert mine.dtype == RESULT_DTYPE and mine.size == e - b
    # rows are plain bytes: gather them as uint8 tensors of the (known) per-rank sizes
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    parts = [torch.empty((re - rb) * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev) for (rb, re) in ranges]
    src = torch.from_numpy(mine.view(np.uint8).copy()).to(dev)
    most = max(p.numel() for p in parts)
    # all_gather wants equal shapes: pad to the largest shard
    padded = torch.zeros(most, dtype=torch.uint8, device=dev)
    padded[:src.numel()] = src
    bufs = [torch.empty(most, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    out = np.concatenate([bufs[r][:parts[r].numel()].cpu().numpy() for r in range(world)]).view(RESULT_DTYPE)
    return out


# ---- C5: many databases against one reference, SAMPLES sharded over the GPUs -----------------------------------------
def assign_samples(n_samples: int, world: int) -> list[list[int]]:
    """contiguous blocks of samples per rank (64 databases on 8 GPUs: 8 each); every sample on exactly one rank"""
    cuts = [n_samples * r // world for r in range(world + 1)]
    return [list(range(cuts[r], cuts[r + 1])) for r in range(world)]


def cohort_sharded(screen_sample_fn, n_samples: int, n_windows: int, group=None) -> np.ndarray:
    """screen_sample_fn(sample) -> CELL_DTYPE array of n_windows cells (Cohort.fetch of the column this rank filled with
    Cohort.add_plan after screening database `sample` against the resident reference).  Every rank screens its block of
    samples; the columns are gathered so that every rank returns the whole [n_samples, n_windows] matrix, in sample
    order — what Plugins/Cohort.java builds by re-reading one KCF per sample.  The only communication is this gather."""
    import torch
    import torch.distributed as dist
    from ._lib import CELL_DTYPE
    if not (dist.is_available() and dist.is_initialized()):
        cols = [np.ascontiguousarray(screen_sample_fn(s), CELL_DTYPE) for s in range(n_samples)]
        return np.stack(cols) if cols else np.zeros((0, n_windows), CELL_DTYPE)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = assign_samples(n_samples, world)
    mine = [np.ascontiguousarray(screen_sample_fn(s), CELL_DTYPE) for s in blocks[rank]]
    assert all(c.shape == (n_windows,) for c in mine)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    row = n_windows * CELL_DTYPE.itemsize
    most = max(len(b) for b in blocks) * row
    padded = torch.zeros(max(most, 1), dtype=torch.uint8, device=dev)
    if mine:
        src = torch.from_numpy(np.stack(mine).view(np.uint8).reshape(-1).copy()).to(dev)
        padded[:src.numel()] = src
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    parts = [bufs[r][:len(blocks[r]) * row].cpu().numpy().view(CELL_DTYPE).reshape(len(blocks[r]), n_windows) for r in range(world)]
    return np.concatenate(parts) if parts else np.zeros((0, n_windows), CELL_DTYPE)
