"""N > 1 host logic on CPU: 2-rank gloo job, windows sharded, rows gathered in window order.  The per-rank compute is
the CPU oracle here (no GPU in this test); on the GPU box the same function wraps Context.screen (bench.py, test_gpu_*)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from common import Scenario, assert_results_equal
from kcftools_b200 import shard
from kcftools_b200.api import fixed_windows


def test_partition_is_contiguous_balanced_and_complete():
    rng = np.random.default_rng(3)
    for n, world in [(0, 2), (1, 4), (7, 8), (1000, 2), (1000, 3), (18012, 8)]:
        lengths = rng.integers(31, 50_001, n)
        ranges = shard.partition(lengths, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:])) and all(b <= e for b, e in ranges)
        if n >= 100 * world:
            sums = [int(lengths[b:e].sum()) for b, e in ranges]
            assert max(sums) - min(sums) <= 2 * int(lengths.max())


def test_local_slice_rebases_segments():
    from common import windows_from_lists
    wins, segs = windows_from_lists([[(0, 0, 10)], [(0, 5, 7), (1, 0, 9)], [(1, 3, 4)], [(0, 1, 2), (0, 9, 2), (1, 1, 1)]])
    assert list(shard.window_lengths(wins, segs)) == [10, 16, 4, 5]
    w, s = shard.local_slice(wins, segs, 1, 3)
    assert list(w["first_seg"]) == [0, 2] and list(w["n_segs"]) == [2, 1] and s.size == 3 and tuple(s[0]) == (0, 5, 7)
    w, s = shard.local_slice(wins, segs, 2, 2)
    assert w.size == 0 and s.size == 0


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = Scenario(seq_lens=(60_000, 23_457, 40), seed=4, n_bins=16)
    wins, segs, *_ = fixed_windows(sc.seq_lens, 3000, 0, 31)
    odb = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    calls = []

    def screen(w, s):
        calls.append(int(w.size))
        rc, res = odb.screen(sc.seqs(), w, s, threads=2)
        assert rc == 0
        return res
    got = shard.screen_sharded(screen, wins, segs)
    rc, want = odb.screen(sc.seqs(), wins, segs, threads=2)
    assert_results_equal(got, want)
    assert len(calls) == 1 and 0 < calls[0] < wins.size  # this rank screened only its share
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), got.view(np.uint8))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_screen(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert (a == b).all() and a.size > 0


def test_grid_layout_and_sample_blocks():
    # scan placement on N = T x S ranks: slice r % T, window shard r // T, reduction group = the T ranks of the shard
    assert shard.grid_layout(5, 8, 2) == (1, 2, [4, 5])
    assert shard.grid_layout(3, 4, 4) == (3, 0, [0, 1, 2, 3])
    assert shard.grid_layout(0, 1, 1) == (0, 0, [0])
    seen = set()
    for r in range(8):
        sl, sh, grp = shard.grid_layout(r, 8, 4)
        assert r in grp and len(grp) == 4 and (sl, sh) not in seen
        seen.add((sl, sh))
    with pytest.raises(ValueError):
        shard.grid_layout(0, 8, 3)
    for n, world in [(64, 8), (5, 2), (3, 4), (0, 2)]:
        blocks = shard.assign_samples(n, world)
        assert len(blocks) == world and [s for b in blocks for s in b] == list(range(n))
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1


def _cohort_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from kcftools_b200._lib import CELL_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_samples, n_windows = 5, 37
    calls = []

    def column(s):  # stands for: open database s, plan.run, cohort.add_plan, cohort.fetch
        calls.append(s)
        c = np.zeros(n_windows, CELL_DTYPE)
        c["obs"] = np.arange(n_windows) * 10 + s
        c["ibs"] = -1
        c["kmer_count"] = (np.arange(n_windows, dtype=np.int64) + 1) * (s + 1) * 1_000_003
        c["score"] = s + np.arange(n_windows) / 64.0
        return c
    got = shard.cohort_sharded(column, n_samples, n_windows)
    assert calls == shard.assign_samples(n_samples, world)[rank]  # only this rank's databases were screened here
    assert got.shape == (n_samples, n_windows)
    for s in range(n_samples):
        want = column(s)
        assert (got[s] == want).all()
    np.save(os.path.join(out_dir, f"cohort{rank}.npy"), got.view(np.uint8))
    dist.destroy_process_group()


def test_two_rank_gloo_samples_sharded_cohort(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_cohort_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "cohort0.npy")
    b = np.load(tmp_path / "cohort1.npy")
    assert (a == b).all() and a.size == 5 * 37 * 40


def _scan_worker(rank, world, port, out_dir):
    """host logic of the scan placement (kcftools_b200.partitioned.screen_partitioned_scan) with the three library calls
    replaced by CPU stand-ins: rank r "owns" the hit bits b with (word + b) % T == slice, every rank of a window shard's
    group must see the union after the all-reduce, batch by batch, and only its own shard"""
    import types
    import numpy as np
    import torch
    import torch.distributed as dist
    from kcftools_b200 import partitioned, shard as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 2
    groups = [dist.new_group(list(range(s * T, (s + 1) * T))) for s in range(world // T)]
    slice_id, shard_id, members = sh.grid_layout(rank, world, T)
    assert rank in members
    n_tiles = 11 + shard_id  # the shards differ in size
    rng = np.random.default_rng(77 + shard_id)  # the same "truth" on every rank of a shard
    truth = rng.integers(0, 2**32, n_tiles * 64, dtype=np.uint64).astype(np.uint32)
    tsum = rng.integers(0, 10**12, n_tiles).astype(np.int64)
    bit = np.arange(32, dtype=np.uint64)
    word = np.arange(n_tiles * 64, dtype=np.uint64)
    own_mask = np.zeros(n_tiles * 64, np.uint32)
    for b in range(32):
        own_mask |= (((word + bit[b]) % T == slice_id).astype(np.uint32) << np.uint32(b))
    seen = {"owned": [], "fold": []}

    def fake_owned(ctx, db, plan, t0, t1, min_count, torch_, dev):
        t1 = min(t1, plan.n_tiles)
        seen["owned"].append((t0, t1))
        hit = (truth[t0 * 64:t1 * 64] & own_mask[t0 * 64:t1 * 64]).view(np.int32).copy()
        sums = np.where(np.arange(t0, t1) % T == slice_id, tsum[t0:t1], 0).astype(np.int64)
        return torch_.from_numpy(hit), torch_.from_numpy(sums)

    def fake_fold(ctx, plan, t0, t1, hit, sums):
        t1 = min(t1, plan.n_tiles)
        seen["fold"].append((t0, t1))
        assert (hit.numpy().view(np.uint32) == truth[t0 * 64:t1 * 64]).all()  # sum of disjoint bit sets = their union
        assert (sums.numpy() == tsum[t0:t1]).all()

    partitioned._scan_owned, partitioned._scan_fold = fake_owned, fake_fold
    partitioned._finish = lambda ctx, plan, weights: ("rows of shard", shard_id)
    plan = types.SimpleNamespace(n_tiles=n_tiles)
    ctx = types.SimpleNamespace(device=0)
    out = partitioned.screen_partitioned_scan(ctx, None, plan, group=groups[shard_id], batch_tiles=4)
    assert out == ("rows of shard", shard_id)
    want = [(t, min(t + 4, n_tiles)) for t in range(0, n_tiles, 4)]
    assert seen["owned"] == want and seen["fold"] == want
    open(os.path.join(out_dir, f"scan{rank}.ok"), "w").write("ok")
    dist.destroy_process_group()


def test_four_rank_gloo_scan_placement_groups(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_scan_worker, args=(4, port, str(tmp_path)), nprocs=4, join=True)
    assert all((tmp_path / f"scan{r}.ok").exists() for r in range(4))


def test_library_cut_equals_python_cut():
    """kcf_shard_windows (what kcf_screen_sharded and `getVariations --devices` use) makes the cut shard.partition makes (what
    the torchrun bench uses): same ranges for fixed windows, multi-segment windows, more shards than windows.  Host arithmetic
    only: no device is touched."""
    from kcftools_b200 import shard
    from kcftools_b200.api import fixed_windows, shard_windows
    from common import windows_from_lists
    rng = np.random.default_rng(9)
    cases = [fixed_windows([1_000_003, 77, 250_000], 50_000, 0, 31)[:2],
             windows_from_lists([[(0, int(rng.integers(0, 1000)), int(rng.integers(1, 5000))) for _ in range(int(rng.integers(1, 6)))] for _ in range(257)]),
             windows_from_lists([[(0, 0, 10)], [(0, 5, 7)]])]
    for wins, segs in cases:
        for n in (1, 2, 3, 8, 16):
            got = [int(b) for b in shard_windows(wins, segs, n)]
            want = shard.partition(shard.window_lengths(wins, segs), n)
            assert got == [r[0] for r in want] + [wins.size]
            assert all(a <= b for a, b in zip(got, got[1:]))


def test_pipelined_exchange_schedule_orders_workspaces_and_barriers():
    """host logic of the pipelined exchange (partitioned._pipelined) against a recording stand-in for the library: batch
    b + 1 is sent into the OTHER workspace before batch b is answered, the context's stream joins that send before the
    barrier, every batch is folded after its barrier from the workspace it was sent into, and a workspace is never sent
    into again before the batch it held has been folded."""
    from kcftools_b200 import partitioned

    log = []

    class Lib:
        def kcf_xg_send(self, c, d, p, x, t0, t1):
            log.append(("send", x, t0, t1))
            return 0

        def kcf_xg_answer(self, c, d, x):
            log.append(("answer", x))
            return 0

        def kcf_xg_fold(self, c, p, x, t0, t1, mc):
            log.append(("fold", x, t0, t1, mc))
            return 0

    class Ctx:
        _lib, _h = Lib(), "ctx"

        def _check(self, rc):
            assert rc == 0

    class H:
        def __init__(self, h):
            self._h = h

    class X(H):
        def join(self):
            log.append(("join", self._h))

    for nb in (1, 2, 5):
        log.clear()
        partitioned._pipelined([(Ctx(), H("db"), H("plan"), (X("x0"), X("x1")))], nb, 7, 2, lambda: log.append(("barrier",)))
        sends = [e for e in log if e[0] == "send"]
        assert [(e[1], e[2], e[3]) for e in sends] == [("x%d" % (b % 2), 7 * b, 7 * b + 7) for b in range(nb)]
        folds = [e for e in log if e[0] == "fold"]
        assert [(e[1], e[2], e[3], e[4]) for e in folds] == [("x%d" % (b % 2), 7 * b, 7 * b + 7, 2) for b in range(nb)]
        assert sum(e[0] == "barrier" for e in log) == nb + 1  # one per batch + the one after the first send
        for b in range(nb):
            i_ans = [i for i, e in enumerate(log) if e[0] == "answer"][b]
            i_fold = log.index(folds[b])
            assert log[i_ans][1] == "x%d" % (b % 2)
            assert ("barrier",) in log[i_ans:i_fold]  # the answers of all ranks land before the fold reads them
            i_send = log.index(sends[b])
            assert i_send < i_ans and ("join", "x%d" % (b % 2)) in log[i_send:i_ans + 2] and ("barrier",) in log[i_send:i_ans + 3 if b else i_ans]
            if b + 1 < nb:  # the next batch goes out before this one is answered ...
                assert log.index(sends[b + 1]) < i_ans
            if b >= 2:      # ... and never into a workspace whose batch has not been folded
                assert log.index(folds[b - 2]) < i_send
