"""Screening against a database that is partitioned over the GPUs of one node (placement 1).

Each rank holds 1/world of the table (cut by home line) and a shard of the windows.  A rank can no longer answer its
own k-mers, so the path gets its one real exchange step: per batch of tiles the k-mers are grouped by owning rank
(kcf_xchg_extract), moved with an all-to-all, looked up by their owners (kcf_xchg_lookup), and the counts travel back
with a second all-to-all before the per-window statistics are folded (kcf_xchg_fold, kcf_plan_finalize).

Second strategy, `screen_partitioned_scan*` ("scan placement"): no k-mer leaves its GPU.  Every rank holds ALL windows,
walks all of them, probes only the k-mers whose home line it owns (kcf_scan_owned) and the per-position hit bits and
per-tile count sums are sum-reduced over the ranks (one bit per k-mer instead of 12-16 bytes each way), then folded
(kcf_scan_fold, kcf_plan_finalize).  Every rank ends up with every row.

`screen_partitioned` is the k-mer exchange over PEER MEMORY (kcf_xg_*: every rank's workspace is mapped by its peers through
CUDA IPC; the screening kernel writes k-mers straight into their owners' inboxes over NVLink, the owners write counts straight
back); `screen_partitioned_a2a` is the same exchange as NCCL all-to-all collectives (round 1, kept for comparison).  The
`*_local` variants drive several contexts of ONE process in lockstep — the single-GPU tests of exactly the same library
calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import RESULT_DTYPE

BATCH_TILES = 1 << 16  # 2048 positions each: 1.3e8 positions, ~2 GB of exchange buffers per rank and batch


def _extract(ctx, db, plan, t0, t1, world, torch, dev):
    npos = max(0, min(t1, plan.n_tiles) - t0) * 2048
    keys = torch.empty(max(npos, 1), dtype=torch.int64, device=dev)
    homes = torch.empty(max(npos, 1), dtype=torch.int32, device=dev)
    src = torch.empty(max(npos, 1), dtype=torch.int32, device=dev)
    counts = (C.c_uint64 * world)()
    ctx._check(ctx._lib.kcf_xchg_extract(ctx._h, db._h, plan._h, t0, t1, world, keys.data_ptr(), homes.data_ptr(), src.data_ptr(), npos, counts))
    sc = [int(c) for c in counts]
    n = sum(sc)
    return keys[:n], homes[:n], src[:n], sc


def _lookup(ctx, db, keys, homes, torch, dev):
    out = torch.empty(max(keys.numel(), 1), dtype=torch.int32, device=dev)
    ctx._check(ctx._lib.kcf_xchg_lookup(ctx._h, db._h, keys.data_ptr(), homes.data_ptr(), keys.numel(), out.data_ptr()))
    return out[:keys.numel()]


def _fold(ctx, plan, t0, t1, counts_back, src, min_count):
    ctx._check(ctx._lib.kcf_xchg_fold(ctx._h, plan._h, t0, t1, counts_back.data_ptr(), src.data_ptr(), src.numel(), min_count))


def _finish(ctx, plan, weights):
    w = (C.c_double * 3)(*weights)
    ctx._check(ctx._lib.kcf_plan_finalize(ctx._h, plan._h, w))
    return plan.fetch()


def screen_partitioned_a2a(ctx, db, plan, group=None, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = BATCH_TILES,
                           phases: dict | None = None) -> np.ndarray:
    """the k-mer exchange as NCCL all-to-all collectives over caller-owned buffers (round-1 path, kept for comparison with the
    peer-memory exchange below: five host-synchronous phases per batch, 12 B out + 4 B back per k-mer).  One rank of a
    torch.distributed job: `db` was opened with placement=1 after ctx.set_partition(rank, world), `plan` holds THIS rank's
    windows.  Returns this rank's rows.  `phases` (optional dict) accumulates seconds per phase — a device synchronisation is
    then inserted after each — and the bytes this rank put on the wire (`_bytes_out`, `_bytes_back`; `_n` counts the calls)."""
    import torch
    import torch.distributed as dist
    world, dev = dist.get_world_size(group), torch.device("cuda", ctx.device)
    # every rank takes part in every all-to-all: loop over the largest batch count
    import os
    import time
    me = dist.get_rank(group)
    trace = phases is not None or (os.environ.get("KCF_PART_TRACE") and me == 0)
    tt = phases if phases is not None else {}
    for k_ in ("extract", "a2a_keys", "lookup", "a2a_counts", "fold"):
        tt.setdefault(k_, 0.0)
    tt["_n"] = tt.get("_n", 0) + 1

    def lap(name, t):
        if trace:
            torch.cuda.synchronize(dev)
            tt[name] += time.perf_counter() - t
        return time.perf_counter()
    nb = torch.tensor([(plan.n_tiles + batch_tiles - 1) // batch_tiles], device=dev)
    dist.all_reduce(nb, op=dist.ReduceOp.MAX, group=group)
    for b in range(int(nb.item())):
        t0, t1 = b * batch_tiles, (b + 1) * batch_tiles
        t = time.perf_counter()
        keys, homes, src, sc = _extract(ctx, db, plan, t0, t1, world, torch, dev)
        t = lap("extract", t)
        send = torch.tensor(sc, dtype=torch.int64, device=dev)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        rc = [int(x) for x in recv.tolist()]
        rkeys = torch.empty(sum(rc), dtype=torch.int64, device=dev)
        rhomes = torch.empty(sum(rc), dtype=torch.int32, device=dev)
        dist.all_to_all_single(rkeys, keys, rc, sc, group=group)
        dist.all_to_all_single(rhomes, homes, rc, sc, group=group)
        off_rank = sum(sc) - sc[me]  # what stays on this rank does not cross NVLink
        tt["_bytes_out"] = tt.get("_bytes_out", 0) + 12 * off_rank
        tt["_bytes_back"] = tt.get("_bytes_back", 0) + 4 * off_rank
        torch.cuda.current_stream(dev).synchronize()  # NCCL ran on torch's stream, the library has its own
        t = lap("a2a_keys", t)
        rcounts = _lookup(ctx, db, rkeys, rhomes, torch, dev)
        t = lap("lookup", t)
        back = torch.empty(sum(sc), dtype=torch.int32, device=dev)
        dist.all_to_all_single(back, rcounts, sc, rc, group=group)
        torch.cuda.current_stream(dev).synchronize()
        t = lap("a2a_counts", t)
        _fold(ctx, plan, t0, t1, back, src, min_count)
        t = lap("fold", t)
    if phases is None and trace:
        print("[partitioned] ms per phase:", {k: round(1e3 * v, 3) for k, v in tt.items() if not k.startswith("_")}, flush=True)
    return _finish(ctx, plan, weights)


def screen_partitioned_a2a_local(ranks, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = BATCH_TILES) -> list[np.ndarray]:
    """`ranks` = [(ctx, db, plan), ...]: every slice of one database, each with its own window shard, all on GPUs this
    process can see (typically the same one).  Runs the ranks in lockstep; the all-to-all is done by slicing."""
    import torch
    world = len(ranks)
    devs = [torch.device("cuda", r[0].device) for r in ranks]
    nb = max((r[2].n_tiles + batch_tiles - 1) // batch_tiles for r in ranks)
    for b in range(nb):
        t0, t1 = b * batch_tiles, (b + 1) * batch_tiles
        ext = [_extract(ctx, db, plan, t0, t1, world, torch, devs[i]) for i, (ctx, db, plan) in enumerate(ranks)]
        offs = [np.concatenate([[0], np.cumsum(e[3])]) for e in ext]
        looked = []
        for o, (ctx, db, plan) in enumerate(ranks):  # owner o receives from every sender s
            rk = torch.cat([ext[s][0][offs[s][o]:offs[s][o + 1]].to(devs[o]) for s in range(world)])
            rh = torch.cat([ext[s][1][offs[s][o]:offs[s][o + 1]].to(devs[o]) for s in range(world)])
            looked.append(_lookup(ctx, db, rk, rh, torch, devs[o]))
        for s, (ctx, db, plan) in enumerate(ranks):  # counts travel back in the order they were sent
            parts = []
            for o in range(world):
                before = sum(ext[q][3][o] for q in range(s))
                parts.append(looked[o][before:before + ext[s][3][o]].to(devs[s]))
            back = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int32, device=devs[s])
            _fold(ctx, plan, t0, t1, back, ext[s][2], min_count)
    return [_finish(ctx, plan, weights) for (ctx, db, plan) in ranks]


# ---- k-mer exchange over peer memory (kcf_xg_*) ----------------------------------------------------------------------------
class Exchange:
    """this rank's exchange workspace, connected to its peers'"""

    def __init__(self, ctx, db, rank: int, world: int, batch_tiles: int):
        self.ctx, self.rank, self.world, self.batch_tiles = ctx, rank, world, batch_tiles
        self._h = C.c_void_p()
        ctx._check(ctx._lib.kcf_xg_create(ctx._h, db._h, rank, world, batch_tiles, C.byref(self._h)))

    def export(self) -> tuple[bytes, int]:
        h = (C.c_uint8 * 64)()
        p = C.c_void_p()
        self.ctx._check(self.ctx._lib.kcf_xg_export(self._h, h, C.byref(p), None))
        return bytes(h), p.value

    def connect(self, handles: list[bytes] | None = None, pointers: list[int] | None = None):
        if pointers is not None:
            arr = (C.c_void_p * self.world)(*pointers)
            self.ctx._check(self.ctx._lib.kcf_xg_connect(self._h, None, arr))
        else:
            blob = b"".join(handles)
            self.ctx._check(self.ctx._lib.kcf_xg_connect(self._h, blob, None))

    def pipeline(self, send_ctas_per_sm: int):
        """sends of this workspace run on their own stream from now on, on at most that many resident CTAs per SM"""
        self.ctx._check(self.ctx._lib.kcf_xg_pipeline(self._h, send_ctas_per_sm))

    def join(self):
        """the context's stream waits for this workspace's last send"""
        self.ctx._check(self.ctx._lib.kcf_xg_join(self.ctx._h, self._h))

    def status(self) -> tuple[int, int, int]:
        """synchronises; raises on a workspace overflow.  Returns (bytes out per run, bytes back per run, runs sent in the last batch)"""
        a, b, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.ctx._check(self.ctx._lib.kcf_xg_status(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.kcf_xg_destroy(self._h)
            self._h = C.c_void_p()


def _exchange_of(ctx, db, plan, rank, world, batch_tiles, connect, n: int = 1):
    """the workspaces are kept on the plan: created and connected once, reused by every later call"""
    xs = getattr(plan, "_exchange", None)
    if xs is None or len(xs) < n or xs[0].world != world or xs[0].batch_tiles != batch_tiles:
        for x in xs or []:
            x.close()
        xs = []
        for _ in range(n):
            x = Exchange(ctx, db, rank, world, batch_tiles)
            connect(x)
            xs.append(x)
        plan._exchange = xs
    return xs


SEND_CTAS = 10  # of the 20 CTAs per SM the screening kernel would take: the other half of the SM answers the batch before


def _pipelined(parts, nb, batch_tiles, min_count, barrier):
    """The batch loop with the send of batch b + 1 beside the answers of batch b.  `parts` = [(ctx, db, plan, (x0, x1)), ...]
    (one entry per rank this process drives: one under torch.distributed, all of them in the lockstep harness), `barrier()` =
    every rank's queued sends and answers have landed (queued on the library's stream, or a device synchronisation)."""
    def send(b, which):
        for ctx, db, plan, xs in parts:
            ctx._check(ctx._lib.kcf_xg_send(ctx._h, db._h, plan._h, xs[which]._h, b * batch_tiles, (b + 1) * batch_tiles))

    send(0, 0)
    for ctx, db, plan, xs in parts:
        xs[0].join()
    barrier()
    for b in range(nb):
        cur, nxt = b % 2, (b + 1) % 2
        if b + 1 < nb:
            send(b + 1, nxt)  # ordered after the fold of batch b - 1, which last read that workspace
        for ctx, db, plan, xs in parts:
            ctx._check(ctx._lib.kcf_xg_answer(ctx._h, db._h, xs[cur]._h))
            if b + 1 < nb:
                xs[nxt].join()
        barrier()
        for ctx, db, plan, xs in parts:
            ctx._check(ctx._lib.kcf_xg_fold(ctx._h, plan._h, xs[cur]._h, b * batch_tiles, (b + 1) * batch_tiles, min_count))


def screen_partitioned(ctx, db, plan, group=None, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = BATCH_TILES,
                       phases: dict | None = None, pipelined: bool = False) -> np.ndarray:
    """one rank of a torch.distributed job (NCCL, one process per GPU): `db` was opened with placement=1 after
    ctx.set_partition(rank, world), `plan` holds THIS rank's windows.  Returns this rank's rows.

    Per batch of tiles: kcf_xg_send (the screening kernel appends every RUN of k-mers sharing a home line — 16 bytes for up to
    11 k-mers — to its owner's inbox over NVLink), a barrier, kcf_xg_answer (owners fetch a run's line once, look its k-mers up,
    store the counts into the requesters' workspaces), a barrier, kcf_xg_fold.  Everything is
    queued on the library's stream — the barriers are one-element all-reduces issued on that same stream — so the host
    never waits inside the loop.  `pipelined=True` runs a job of more than one batch over two workspaces (_pipelined: the
    send of batch b + 1 on half of every SM beside the answers of batch b, one barrier per batch) — measured SLOWER than the
    plain sequence on 2 x B200 (47.3 against 39.2 ms per 3e9-k-mer job, profiles/r2x): the two kernels compete for the same
    memory pipeline, they do not fill each other's gaps; kept as an option, off by default.  `phases` (optional dict) accumulates seconds per phase (a device synchronisation is then
    inserted after each) and `_bytes_out` / `_bytes_back`, the bytes this rank put on the wire."""
    import time
    import torch
    import torch.distributed as dist
    world, me, dev = dist.get_world_size(group), dist.get_rank(group), torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    batch_tiles = int(min(batch_tiles, max(plan.n_tiles, 1)))
    nb = torch.tensor([plan.n_tiles, batch_tiles], device=dev)
    dist.all_reduce(nb, op=dist.ReduceOp.MAX, group=group)  # every rank takes part in every barrier: the largest tile count and batch
    max_tiles, batch_tiles = int(nb[0].item()), int(nb[1].item())

    def connect(x):
        handle, _ = x.export()
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=group)
        x.connect(handles=handles)
    pipelined = pipelined and phases is None and max_tiles > batch_tiles
    xs = _exchange_of(ctx, db, plan, me, world, batch_tiles, connect, 2 if pipelined else 1)
    token = torch.zeros(1, device=dev)
    if pipelined:  # two workspaces: batch b + 1 is sent while batch b is answered, one barrier per batch
        for x_ in xs[:2]:
            x_.pipeline(SEND_CTAS)
        with torch.cuda.stream(stream):
            _pipelined([(ctx, db, plan, xs)], (max_tiles + batch_tiles - 1) // batch_tiles, batch_tiles, min_count,
                       lambda: dist.all_reduce(token, group=group))
        for x_ in xs[:2]:
            x_.status()
        return _finish(ctx, plan, weights)
    x = xs[0]
    x.pipeline(0)
    tt = phases if phases is not None else {}
    for k_ in ("send", "barrier_1", "answer", "barrier_2", "fold"):
        tt.setdefault(k_, 0.0)
    tt["_n"] = tt.get("_n", 0) + 1

    def lap(name, t):
        if phases is not None:
            torch.cuda.synchronize(dev)
            tt[name] += time.perf_counter() - t
        return time.perf_counter()
    with torch.cuda.stream(stream):
        for t0 in range(0, max(max_tiles, 1), batch_tiles):
            t1 = t0 + batch_tiles
            t = time.perf_counter()
            ctx._check(ctx._lib.kcf_xg_send(ctx._h, db._h, plan._h, x._h, t0, t1))
            x.join()
            t = lap("send", t)
            if phases is not None:  # what crosses NVLink: the runs of this batch that other ranks answer, there and back
                out_b, back_b, runs = x.status()
                tt["_runs"] = tt.get("_runs", 0) + runs
                tt["_bytes_out"] = tt.get("_bytes_out", 0) + out_b * runs * (world - 1) // world
                tt["_bytes_back"] = tt.get("_bytes_back", 0) + back_b * runs * (world - 1) // world
                t = time.perf_counter()
            dist.all_reduce(token, group=group)
            t = lap("barrier_1", t)
            ctx._check(ctx._lib.kcf_xg_answer(ctx._h, db._h, x._h))
            t = lap("answer", t)
            dist.all_reduce(token, group=group)
            t = lap("barrier_2", t)
            ctx._check(ctx._lib.kcf_xg_fold(ctx._h, plan._h, x._h, t0, t1, min_count))
            t = lap("fold", t)
    x.status()
    return _finish(ctx, plan, weights)


def screen_partitioned_local(ranks, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = BATCH_TILES,
                             pipelined: bool = False) -> list[np.ndarray]:
    """`ranks` = [(ctx, db, plan), ...]: every slice of one database, each with its own window shard, all contexts on GPUs
    of THIS process (typically the same one).  The same library calls as screen_partitioned, the workspaces connected by
    plain device pointers, the barriers replaced by device synchronisations."""
    import torch
    world = len(ranks)
    batch_tiles = int(min(batch_tiles, max(max(r[2].n_tiles for r in ranks), 1)))
    n_tiles = max(r[2].n_tiles for r in ranks)
    pairs = []
    for _ in range(2 if pipelined else 1):
        xs = [Exchange(ctx, db, i, world, batch_tiles) for i, (ctx, db, plan) in enumerate(ranks)]
        ptrs = [x.export()[1] for x in xs]
        for x in xs:
            x.connect(pointers=ptrs)
        pairs.append(xs)
    try:
        if pipelined:
            for xs in pairs:
                for x in xs:
                    x.pipeline(SEND_CTAS)
            parts = [(ctx, db, plan, (pairs[0][i], pairs[1][i])) for i, (ctx, db, plan) in enumerate(ranks)]
            _pipelined(parts, (max(n_tiles, 1) + batch_tiles - 1) // batch_tiles, batch_tiles, min_count, torch.cuda.synchronize)
        else:
            xs = pairs[0]
            for t0 in range(0, max(n_tiles, 1), batch_tiles):
                t1 = t0 + batch_tiles
                for x, (ctx, db, plan) in zip(xs, ranks):
                    ctx._check(ctx._lib.kcf_xg_send(ctx._h, db._h, plan._h, x._h, t0, t1))
                torch.cuda.synchronize()
                for x, (ctx, db, plan) in zip(xs, ranks):
                    ctx._check(ctx._lib.kcf_xg_answer(ctx._h, db._h, x._h))
                torch.cuda.synchronize()
                for x, (ctx, db, plan) in zip(xs, ranks):
                    ctx._check(ctx._lib.kcf_xg_fold(ctx._h, plan._h, x._h, t0, t1, min_count))
        for xs in pairs:
            for x in xs:
                x.status()
        return [_finish(ctx, plan, weights) for (ctx, db, plan) in ranks]
    finally:
        for xs in pairs:
            for x in xs:
                x.close()


# ---- scan placement ---------------------------------------------------------------------------------------------------
def _scan_owned(ctx, db, plan, t0, t1, min_count, torch, dev):
    nt = max(0, min(t1, plan.n_tiles) - t0)
    hit = torch.empty(max(nt * 64, 1), dtype=torch.int32, device=dev)  # one word per 32 positions
    sums = torch.empty(max(nt, 1), dtype=torch.int64, device=dev)
    ctx._check(ctx._lib.kcf_scan_owned(ctx._h, db._h, plan._h, t0, t1, min_count, hit.data_ptr(), sums.data_ptr()))
    return hit[:nt * 64], sums[:nt]


def _scan_fold(ctx, plan, t0, t1, hit, sums):
    ctx._check(ctx._lib.kcf_scan_fold(ctx._h, plan._h, t0, t1, hit.data_ptr(), sums.data_ptr()))


def screen_partitioned_scan(ctx, db, plan, group=None, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = 1 << 20,
                            phases: dict | None = None) -> np.ndarray:
    """one rank of a torch.distributed job: `db` was opened with placement=1 after ctx.set_partition(rank, world), `plan`
    holds ALL windows (the same plan on every rank).  Returns all rows (identical on every rank).  `phases` (optional dict)
    accumulates seconds per phase; `_n` counts the calls."""
    import time
    import torch
    import torch.distributed as dist
    on_gpu = dist.get_backend(group) == "nccl"  # gloo: the host-logic test on CPU (tests/test_shard.py)
    dev = torch.device("cuda", ctx.device) if on_gpu else torch.device("cpu")
    tt = phases if phases is not None else {}
    for k_ in ("scan_owned", "all_reduce", "fold"):
        tt.setdefault(k_, 0.0)
    tt["_n"] = tt.get("_n", 0) + 1
    for t0 in range(0, plan.n_tiles, batch_tiles):
        t1 = t0 + batch_tiles
        ta = time.perf_counter()
        hit, sums = _scan_owned(ctx, db, plan, t0, t1, min_count, torch, dev)  # returns after the kernel (host-synchronous)
        tb = time.perf_counter()
        # the owners' bitmaps are disjoint, so the sum is the union (no carries; NCCL has no bitwise reduction)
        dist.all_reduce(hit, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        if on_gpu:
            torch.cuda.current_stream(dev).synchronize()  # NCCL ran on torch's stream, the library has its own
        tc = time.perf_counter()
        _scan_fold(ctx, plan, t0, t1, hit, sums)
        tt["scan_owned"] += tb - ta
        tt["all_reduce"] += tc - tb
        tt["fold"] += time.perf_counter() - tc
    return _finish(ctx, plan, weights)


def screen_partitioned_scan_local(ranks, min_count: int = 1, weights=(0.3, 0.3, 0.4), batch_tiles: int = 1 << 20) -> list[np.ndarray]:
    """`ranks` = [(ctx, db, plan), ...]: every slice of one database, every plan holding ALL windows, all on GPUs this
    process can see.  Lockstep version of screen_partitioned_scan: the all-reduce is a sum of the ranks' tensors."""
    import torch
    devs = [torch.device("cuda", r[0].device) for r in ranks]
    n_tiles = ranks[0][2].n_tiles
    assert all(r[2].n_tiles == n_tiles for r in ranks), "scan placement: every rank plans the same windows"
    for t0 in range(0, n_tiles, batch_tiles):
        t1 = t0 + batch_tiles
        parts = [_scan_owned(ctx, db, plan, t0, t1, min_count, torch, devs[i]) for i, (ctx, db, plan) in enumerate(ranks)]
        for i, (ctx, db, plan) in enumerate(ranks):
            hit = torch.stack([p[0].to(devs[i]) for p in parts]).sum(0, dtype=torch.int32)
            sums = torch.stack([p[1].to(devs[i]) for p in parts]).sum(0)
            _scan_fold(ctx, plan, t0, t1, hit, sums)
    return [_finish(ctx, plan, weights) for (ctx, db, plan) in ranks]
