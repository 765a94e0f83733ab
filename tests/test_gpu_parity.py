"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.
Bar: integer columns bit-exact, score within 1e-9 relative (north_star)."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from tools import synth
from common import Scenario, assert_results_equal, windows_from_lists
from kcftools_b200.api import KMC, KcfError, fixed_windows

pytestmark = pytest.mark.gpu


def _oracle_screen(sc, wins, segs, min_count=1, w=(0.3, 0.3, 0.4), threads=8):
    odb = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    rc, want = odb.screen(sc.seqs(), wins, segs, min_count=min_count, w=w, threads=threads)
    return rc, want


@pytest.fixture(scope="module")
def sc_main():
    return Scenario(seq_lens=(300_000, 123_457, 40), seed=1, n_bins=64)


def test_db_info_and_counts(ctx, sc_main):
    sc = sc_main
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    assert db.getKmerLength() == 31 and db.getPrefixLength() == 7 and db.isBothStrands()
    assert db.info.total_kmers == sc.kmc.total
    assert db.info.resident_kmers == sc.kmc.total and db.info.unreachable_kmers == 0
    odb = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    raw, lb, lw, sl = sc.seqs()[0]
    rc, text = ob.get_sequence(raw, lb, lw, sl, 0, 20000)
    text = text.decode().upper()
    kms = [text[i:i + 31] for i in range(0, len(text) - 31) if set(text[i:i + 31]) <= set("ACGT")]
    got = db.getCounts(kms)
    want = np.array([odb.count(s) for s in kms], np.int32)
    assert (got == want).all() and (want > 0).any() and (want == 0).any()
    # lower case input is upper-cased like Fasta.getKmersList does
    assert db.getCount(kms[5].lower()) == want[5]
    db.close()


def test_fixed_windows_parity(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    for (W, step) in [(50_000, 0), (5000, 0), (3000, 1000), (700, 2500)]:
        wins, segs, *_ = fixed_windows(sc.seq_lens, W, step, 31)
        rc, want = _oracle_screen(sc, wins, segs)
        assert rc == 0
        got = ctx.screen(db, wins, segs)
        assert_results_equal(got, want)
        assert want["obs"].sum() > 0 and (want["variations"] > 0).any() and (want["eff_len"] != want["total_kmers"] + 30).any()
    db.close()


def test_per_kmer_counts_of_a_window(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    wins, segs, starts, ends, sids = fixed_windows(sc.seq_lens, 5000, 0, 31)
    plan = ctx.plan(31, wins, segs)
    plan.run(db)
    plan.fetch()
    odb = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    for w in (0, 7, len(starts) - 1):
        raw, lb, lw, sl = sc.seqs()[sids[w]]
        rc, text = ob.get_sequence(raw, lb, lw, sl, int(starts[w]), int(ends[w] - starts[w]))
        rc, res, counts = odb.process_window(text, want_counts=True)
        got = plan.window_counts(db, w)
        assert got.size == counts.size and (got == counts).all()
    plan.close()
    db.close()


def test_min_count_and_weights(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    wins, segs, *_ = fixed_windows(sc.seq_lens, 10_000, 0, 31)
    for mc, w in [(1, (0.3, 0.3, 0.4)), (5, (0.5, 0.25, 0.25)), (12, (0.0, 0.0, 1.0)), (300, (0.3, 0.3, 0.4))]:
        rc, want = _oracle_screen(sc, wins, segs, min_count=mc, w=w)
        assert rc == 0
        got = ctx.screen(db, wins, segs, min_count=mc, weights=w)
        assert_results_equal(got, want)
    # min_count 300 > max counter (255): no hit anywhere -> variations 1, right = total (Q5)
    assert (got["obs"] == 0).all() and (got["variations"][got["total_kmers"] > 0] == 1).all()
    assert (got["right"] == got["total_kmers"]).all() and (got["score"] == 0).all()
    # weights that do not sum to 1.0 are fatal in the reference once a score is computed
    with pytest.raises(KcfError) as e:
        ctx.screen(db, wins, segs, weights=(0.3, 0.3, 0.5))
    assert e.value.code == -8
    with pytest.raises(KcfError) as e:
        ctx.screen(db, wins, segs, min_count=0)
    assert e.value.code == -5
    db.close()


def test_multi_segment_windows(ctx, sc_main):
    """gene / transcript style windows: concatenated loci, k-mers span the junctions (GTF.java:240-244)."""
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    rng = np.random.default_rng(5)
    lists = []
    for _ in range(300):
        sid = int(rng.integers(0, 2))
        n = sc.seq_lens[sid]
        ns = int(rng.integers(1, 9))
        pos = int(rng.integers(0, n - 30_000))
        sl = []
        for _ in range(ns):
            ln = int(rng.integers(1, 2500))
            sl.append((sid, pos, ln))
            pos += ln + int(rng.integers(0, 800))  # 0 => abutting loci
        lists.append(sl)
    lists.append([(0, 0, 10), (1, 5, 12), (0, 100, 9)])       # shorter than k in total: 31 bases, exactly one k-mer
    lists.append([(0, 0, 5), (0, 5, 5)])                       # 10 bases: no k-mer
    lists.append([(2, 0, 40)])                                  # the 40-base sequence
    lists.append([(0, 0, 300_000), (1, 0, 123_457)])           # very long window (many tiles)
    wins, segs = windows_from_lists(lists)
    rc, want = _oracle_screen(sc, wins, segs)
    assert rc == 0
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    assert want["total_kmers"][-3] == 0 and want["total_kmers"][-4] <= 1
    db.close()


@pytest.mark.parametrize("k,P,L,cs,both,line", [(32, 8, 9, 2, True, 70), (21, 5, 7, 1, False, 60), (13, 1, 5, 3, True, 33),
                                                 (5, 1, 3, 1, True, 7), (31, 3, 9, 1, True, 1000), (16, 4, 5, 0, True, 60)])
def test_other_k_and_layouts(ctx, k, P, L, cs, both, line):
    sc = Scenario(seq_lens=(30_000, 2_000), k=k, P=P, L=L, n_bins=8, counter_size=cs, both_strands=both, seed=k, line=line,
                  coverage=8.0 if cs else 1.0)
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    assert db.info.unreachable_kmers == 0
    wins, segs, *_ = fixed_windows(sc.seq_lens, 4000, 0, k)
    rc, want = _oracle_screen(sc, wins, segs)
    assert rc == 0
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    if cs == 0:
        assert (got["obs"] == 0).all()  # Q7: counter_size 0 => nothing is ever observed
    db.close()


def test_high_load_factor_uses_displacement_and_stash(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    ctx.set_load_factor(0.9)
    try:
        db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    finally:
        ctx.set_load_factor(0.0)
    assert db.info.resident_kmers == sc.kmc.total
    wins, segs, *_ = fixed_windows(sc.seq_lens, 20_000, 0, 31)
    rc, want = _oracle_screen(sc, wins, segs)
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    db.close()


@pytest.mark.parametrize("m,lf", [(16, 0.5), (12, 0.3), (4, 0.5), (1, 0.9), (10, 0.9), (18, 0.3), (19, 0.3), (24, 0.6)])
def test_minimizer_length_and_load_factor_never_change_results(ctx, sc_main, m, lf):
    """the home-line function (minimizer length) and the load factor are layout knobs only.  m = 4 / 1 pile thousands
    of keys onto few minimizers: displaced lines, continuation fetches and the stash all get exercised."""
    sc = sc_main
    sc.add_to(ctx)
    ctx.set_minimizer_length(m)
    ctx.set_load_factor(lf)
    try:
        db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    finally:
        ctx.set_minimizer_length(0)
        ctx.set_load_factor(0.0)
    assert db.info.resident_kmers == sc.kmc.total
    if m <= 4:
        assert db.info.stash_kmers > 0
    wins, segs, *_ = fixed_windows(sc.seq_lens, 20_000, 0, 31)
    rc, want = _oracle_screen(sc, wins, segs)
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    db.close()


@pytest.mark.parametrize("k,m", [(21, 18), (21, 19), (21, 9), (21, 8), (13, 10), (16, 3)])
def test_sliding_minimum_paths_at_their_boundaries(ctx, k, m):
    """w = k - m + 1 = 4 and 13 bound the register sliding minimum of the screening kernel, 3 and 14 fall to the doubling path:
    (21, 18) w = 4, (21, 19) w = 3, (21, 9) w = 13, (21, 8) w = 14, (13, 10) w = 4, (16, 3) w = 14; windows long enough for
    several 512-position chunks, so the carried hashes between chunks are exercised too"""
    sc = Scenario(seq_lens=(9_000, 700), k=k, P=5 if (k - 5) % 4 == 0 else (k % 4 if k % 4 else 4), L=5, n_bins=8, seed=100 + k + m, n_runs=4)
    sc.add_to(ctx)
    ctx.set_minimizer_length(m)
    try:
        db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    finally:
        ctx.set_minimizer_length(0)
    assert db.info.resident_kmers == sc.kmc.total
    wins, segs, *_ = fixed_windows(sc.seq_lens, 3000, 1700, k)
    rc, want = _oracle_screen(sc, wins, segs)
    assert rc == 0
    assert_results_equal(ctx.screen(db, wins, segs), want)
    db.close()


def test_unreachable_records_are_ignored_like_the_reference(ctx):
    """records sitting in a bin their signature does not map to can never be returned by KMC.getCount."""
    sc = Scenario(seq_lens=(20_000,), seed=9, n_bins=16)
    # rotate the signature map: most records now live in the "wrong" bin for both implementations
    img = sc.kmc
    L = img.L
    nmap = (1 << (2 * L)) + 1
    pre = img.pre.copy()
    map_start = pre.size - 68 - 8 - nmap * 4
    m = pre[map_start:map_start + 4 * nmap].view("<u4").copy()
    m[::3] = (m[::3] + 1) % 16
    pre[map_start:map_start + 4 * nmap] = m.view(np.uint8)
    sc.kmc = synth.KmcImage(pre, img.suf, img.k, img.P, img.L, img.n_bins, img.counter_size, img.total, img.both_strands)
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    assert 0 < db.info.unreachable_kmers < img.total
    wins, segs, *_ = fixed_windows(sc.seq_lens, 2000, 0, 31)
    rc, want = _oracle_screen(sc, wins, segs)
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    db.close()


def test_error_codes(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    # segment outside its sequence (FastaIndex.java:132-135)
    wins, segs = windows_from_lists([[(0, 299_990, 100)]])
    with pytest.raises(KcfError) as e:
        ctx.screen(db, wins, segs)
    assert e.value.code == -6
    wins, segs = windows_from_lists([[(5, 0, 100)]])
    with pytest.raises(KcfError) as e:
        ctx.screen(db, wins, segs)
    assert e.value.code == -6
    # bad version
    pre = sc.kmc.pre.copy()
    pre[pre.size - 8 - 4:pre.size - 8] = 0
    with pytest.raises(KcfError) as e:
        KMC(ctx, pre=pre, suf=sc.kmc.suf)
    assert e.value.code == -3
    # unsorted records inside one (bin, prefix) range: the reference's binary search is undefined -> refused
    sc2 = Scenario(seq_lens=(20_000,), seed=2, n_bins=2, P=3)
    KMC(ctx, pre=sc2.kmc.pre, suf=sc2.kmc.suf).close()
    suf = sc2.kmc.suf.copy()
    suf[4:12], suf[12:20] = sc2.kmc.suf[12:20].copy(), sc2.kmc.suf[4:12].copy()  # 7 suffix bytes + 1 counter byte
    with pytest.raises(KcfError) as e:
        KMC(ctx, pre=sc2.kmc.pre, suf=suf)
    assert e.value.code == -10
    # missing trailing newline: reading the final bases is fatal in the reference (Q9)
    g = synth.random_genome(500, 4)
    rec = synth.fasta_record(g, "x", line=60, trailing_newline=False)
    ctx.ref_clear()
    ctx.ref_add(rec[3:], 60, 61, 500)
    wins, segs = windows_from_lists([[(0, 400, 100)]])
    with pytest.raises(KcfError) as e:
        ctx.screen(db, wins, segs)
    assert e.value.code == -7
    wins, segs = windows_from_lists([[(0, 300, 100)]])
    ctx.screen(db, wins, segs)
    db.close()


def test_empty_inputs(ctx, sc_main):
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    wins, segs = windows_from_lists([])
    assert ctx.screen(db, wins, segs).size == 0
    db.close()
    # empty database: everything is a miss
    import torch
    z = torch.zeros(0, dtype=torch.int64)
    img = synth.kmc_image_from_kmers(z, z, z, k=31, P=7, L=9, n_bins=4, counter_size=1)
    db = KMC(ctx, pre=img.pre, suf=img.suf)
    wins, segs, *_ = fixed_windows(sc.seq_lens, 50_000, 0, 31)
    got = ctx.screen(db, wins, segs)
    assert (got["obs"] == 0).all() and (got["right"] == got["total_kmers"]).all()
    db.close()


@pytest.mark.parametrize("world,batch_tiles,how", [(1, 1 << 16, "peer"), (2, 1 << 16, "peer"), (3, 7, "peer"), (4, 3, "peer"), (2, 1 << 16, "a2a"), (3, 7, "a2a"),
                                                   (1, 5, "pipelined"), (2, 1 << 16, "pipelined"), (3, 7, "pipelined"), (4, 3, "pipelined")])
def test_partitioned_database_exchange_path(sc_main, world, batch_tiles, how):
    """placement 1: every rank holds 1/world of the table and a shard of the windows; k-mers travel to their owners and
    the counts come back.  "peer": the exchange over peer memory (kcf_xg_*: the screening kernel appends into the owners'
    inboxes, owners store the counts into the requesters' workspaces); "a2a": the same as all-to-all collectives over
    caller-owned buffers (kcf_xchg_*); "pipelined": the peer exchange over two workspaces, the send of batch b + 1 on its own
    stream beside the answers of batch b (kcf_xg_pipeline / kcf_xg_join).  Here all ranks live on cuda:0 — workspaces connected by plain pointers, the
    all-to-all done by slicing — the library calls are the ones the NCCL job makes."""
    from kcftools_b200 import shard
    from kcftools_b200.api import Context
    from kcftools_b200.partitioned import screen_partitioned_a2a_local, screen_partitioned_local
    if how == "a2a":
        screen_partitioned_local = screen_partitioned_a2a_local
    elif how == "pipelined":
        import functools
        screen_partitioned_local = functools.partial(screen_partitioned_local, pipelined=True)
    sc = sc_main
    wins, segs, *_ = fixed_windows(sc.seq_lens, 20_000, 0, 31)
    rc, want = _oracle_screen(sc, wins, segs, min_count=2, w=(0.2, 0.3, 0.5))
    ranges = shard.partition(shard.window_lengths(wins, segs), world)
    ranks, ctxs = [], []
    total_resident = 0
    for r in range(world):
        c = Context(0)
        ctxs.append(c)
        sc.add_to(c)
        c.set_partition(r, world)
        db = KMC(c, pre=sc.kmc.pre, suf=sc.kmc.suf, placement=1)
        total_resident += db.info.resident_kmers
        assert db.info.resident_kmers + db.info.elsewhere_kmers == sc.kmc.total
        lw, ls = shard.local_slice(wins, segs, *ranges[r])
        ranks.append((c, db, c.plan(31, lw, ls)))
    assert total_resident == sc.kmc.total  # every record lives on exactly one rank
    if world > 1:
        assert all(r[1].info.resident_kmers > 0 for r in ranks)
        with pytest.raises(KcfError):  # a slice cannot be screened alone
            ranks[0][2].run(ranks[0][1])
    parts = screen_partitioned_local(ranks, min_count=2, weights=(0.2, 0.3, 0.5), batch_tiles=batch_tiles)
    got = np.concatenate(parts)
    assert_results_equal(got, want)
    for (c, db, plan) in ranks:
        plan.close()
        db.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("world,batch_tiles,kind", [(1, 1 << 20, "fixed"), (2, 1 << 20, "fixed"), (3, 5, "fixed"), (4, 1 << 20, "multi")])
def test_partitioned_database_scan_path(sc_main, world, batch_tiles, kind):
    """placement 1, scan strategy: every rank holds 1/world of the table and ALL windows, probes only the k-mers whose
    home line it owns, and the hit bitmaps / count sums are summed over the ranks before the fold (here all ranks live
    on cuda:0 and the all-reduce is a tensor sum — the library calls are the ones the NCCL job makes)."""
    from kcftools_b200.api import Context
    from kcftools_b200.partitioned import screen_partitioned_scan_local
    sc = sc_main
    if kind == "fixed":
        wins, segs, *_ = fixed_windows(sc.seq_lens, 20_000, 0, 31)
    else:  # multi-segment windows whose k-mers span junctions, one shorter than k, one crossing several tiles
        wins, segs = windows_from_lists([[(0, 100, 500), (0, 5_000, 40), (1, 10, 3_000)], [(1, 0, 20)], [(0, 0, 30_000), (1, 2_000, 9_000)],
                                         [(1, 100, 31)], [(0, 50_000, 2_048), (0, 60_000, 2_048)]])
    rc, want = _oracle_screen(sc, wins, segs, min_count=2, w=(0.2, 0.3, 0.5))
    assert rc == 0
    ranks, ctxs = [], []
    for r in range(world):
        c = Context(0)
        ctxs.append(c)
        sc.add_to(c)
        c.set_partition(r, world)
        db = KMC(c, pre=sc.kmc.pre, suf=sc.kmc.suf, placement=1)
        ranks.append((c, db, c.plan(31, wins, segs)))
    parts = screen_partitioned_scan_local(ranks, min_count=2, weights=(0.2, 0.3, 0.5), batch_tiles=batch_tiles)
    for got in parts:  # every rank ends up with every row
        assert_results_equal(got, want)
    for (c, db, plan) in ranks:
        plan.close()
        db.close()
    for c in ctxs:
        c.close()


def test_large_reference_windows_reproduce_the_small_run():
    """size-independent property used at full C4 reference size by tools/scale_c4ref.py (profiles/r1f_scale_c4_reference.json:
    1.5e10 positions, 300,069 windows): a reference whose sequences START with the chromosomes of a smaller run must give
    the smaller run's rows for every window inside those cores, and exact totals / no hits in the unrelated random tails."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "scale_c4ref.py"), "--core", "c2s", "--seqs", "14", "--seq-len", "12000000"],
                       capture_output=True, text=True, cwd=root)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    out = json.loads(p.stdout.strip().split("\n")[-1])
    assert out["ok"] and out["core_windows_identical"] and out["core_windows_compared"] == 14 * 150 and out["tail_totals_exact"]
    assert out["windows"] == 14 * 241 and out["tail_observed_kmers"] < 10


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_random_window_layouts(ctx, sc_main, seed):
    """seeded fuzz over window shapes: 1-9 segments per window, segment lengths from 1 base to several tiles, segments that
    overlap, repeat, abut, sit at sequence ends, windows shorter than k, junctions every few bases (k-mers spanning several
    segments), all mixed in one call; min_count and weights vary with the seed"""
    rng = np.random.default_rng(seed)
    sc = sc_main
    lists = []
    for _ in range(160):
        nseg = int(rng.integers(1, 10))
        segs = []
        for _ in range(nseg):
            sid = int(rng.integers(0, 3))
            n = sc.seq_lens[sid]
            kind = rng.random()
            length = int(rng.integers(1, 12)) if kind < 0.25 else int(rng.integers(12, 400)) if kind < 0.7 else int(rng.integers(400, 9000))
            length = min(length, n)
            start = int(rng.integers(0, n - length + 1)) if rng.random() < 0.9 else (0 if rng.random() < 0.5 else n - length)
            segs.append((sid, start, length))
        lists.append(segs)
    wins, segs = windows_from_lists(lists)
    mc = int(rng.integers(1, 4))
    wts = [(0.3, 0.3, 0.4), (0.5, 0.25, 0.25), (0.0, 0.0, 1.0), (0.125, 0.125, 0.75)][seed % 4]
    rc, want = _oracle_screen(sc, wins, segs, min_count=mc, w=wts)
    assert rc == 0
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    sc.add_to(ctx)
    got = ctx.screen(db, wins, segs, min_count=mc, weights=wts)
    assert_results_equal(got, want)
    assert (want["total_kmers"] == 0).any() and (want["obs"] > 0).any()  # the mix contains empty and observed windows
    db.close()


@pytest.mark.parametrize("n_ctx,kind", [(1, "fixed"), (3, "fixed"), (2, "multi"), (4, "tiny")])
def test_sharded_job_over_several_contexts(sc_main, n_ctx, kind):
    """kcf_screen_sharded: ONE job cut over n contexts (here all on cuda:0; on the box one per GPU), sequences given as host
    bytes, each context uploading only the line-aligned stretches its windows touch.  Rows must equal the oracle's for the
    whole job, in window order, whatever the cut."""
    from kcftools_b200.api import Context, screen_sharded, shard_windows
    from kcftools_b200 import shard
    sc = sc_main
    if kind == "fixed":
        wins, segs, *_ = fixed_windows(sc.seq_lens, 7_000, 0, 31)
    elif kind == "tiny":  # fewer windows than contexts: some shards are empty
        wins, segs = windows_from_lists([[(1, 100, 5_000)], [(0, 299_000, 1_000)]])
    else:  # windows of several segments, out of order, across sequences, one shorter than k
        rng = np.random.default_rng(3)
        lists = [[(0, 100, 500), (0, 5_000, 40), (1, 10, 3_000)], [(1, 0, 20)], [(0, 0, 30_000), (1, 2_000, 9_000)], [(2, 0, 40)]]
        for _ in range(60):
            sid = int(rng.integers(0, 2))
            n = sc.seq_lens[sid]
            lists.append([(sid, int(rng.integers(0, n - 3_000)), int(rng.integers(1, 3_000))) for _ in range(int(rng.integers(1, 5)))])
        wins, segs = windows_from_lists(lists)
    rc, want = _oracle_screen(sc, wins, segs, min_count=2, w=(0.2, 0.3, 0.5))
    assert rc == 0
    bounds = shard_windows(wins, segs, n_ctx)
    assert bounds[0] == 0 and bounds[-1] == wins.size and (np.diff(bounds.astype(np.int64)) >= 0).all()
    # the library cuts where the Python host logic cuts (the torchrun bench shards with the latter)
    assert [int(b) for b in bounds] == [r[0] for r in shard.partition(shard.window_lengths(wins, segs), n_ctx)] + [wins.size]
    ctxs = [Context(0) for _ in range(n_ctx)]
    dbs = [KMC(c, pre=sc.kmc.pre, suf=sc.kmc.suf) for c in ctxs]
    try:
        got = screen_sharded(ctxs, dbs, sc.seqs(), wins, segs, min_count=2, weights=(0.2, 0.3, 0.5))
        assert_results_equal(got, want)
        # errors surface with the reference's wording, whichever context hits them
        bad_w, bad_s = windows_from_lists([[(0, 0, 100)], [(1, 123_400, 100)]])
        with pytest.raises(KcfError) as e:
            screen_sharded(ctxs, dbs, sc.seqs(), bad_w, bad_s)
        assert e.value.code == -6 and "Invalid range" in e.value.msg
    finally:
        for d in dbs:
            d.close()
        for c in ctxs:
            c.close()


def test_sharded_job_pieces_of_a_long_sequence(ctx):
    """a sequence longer than one upload piece (kcf_set_upload_piece): windows straddling the piece boundaries become two
    segments inside the library; lines of odd width, a final line without its newline is still fatal (Q9)"""
    from kcftools_b200.api import screen_sharded
    sc = Scenario(seq_lens=(3_000_000,), seed=77, n_bins=16, line=77, n_runs=6, snp=0.02)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    piece = 1 << 20  # 13,617 whole lines of 77 bases: boundaries at 1,048,509 and 2,097,018
    edge = piece // 77 * 77
    starts = np.concatenate([np.arange(0, 2_990_000, 130_000), [edge - 40, edge - 1, 2 * edge - 9_999, 2_990_000]])
    wins, segs = windows_from_lists([[(0, int(s), 10_000)] for s in starts] + [[(0, 100, 2_500_000)]])  # the last one spans all three pieces
    rc, want = _oracle_screen(sc, wins, segs)
    assert rc == 0
    ctx.set_upload_piece(piece)
    try:
        got = screen_sharded([ctx], [db], sc.seqs(), wins, segs)
    finally:
        ctx.set_upload_piece(0)
    assert_results_equal(got, want)
    assert_results_equal(screen_sharded([ctx], [db], sc.seqs(), wins, segs), want)  # default piece: the whole sequence at once
    db.close()


def test_plans_do_not_outlive_their_sequences(ctx, sc_main):
    """a plan addresses the sequences that were resident when it was created: after kcf_ref_clear it is refused instead of
    reading recycled device memory; several plans queued before their fetches keep their own weight checks"""
    sc = sc_main
    sc.add_to(ctx)
    db = KMC(ctx, pre=sc.kmc.pre, suf=sc.kmc.suf)
    wins, segs, *_ = fixed_windows(sc.seq_lens, 50_000, 0, 31)
    rc, want = _oracle_screen(sc, wins, segs)
    a, b = ctx.plan(31, wins, segs), ctx.plan(31, wins[:2], segs[:2])
    a.run(db, weights=(0.3, 0.3, 0.5))  # bad weights, queued first
    b.run(db)                           # good weights, queued second
    with pytest.raises(KcfError) as e:
        a.fetch()
    assert e.value.code == -8
    assert_results_equal(b.fetch(), want[:2])
    a.run(db)
    assert_results_equal(a.fetch(), want)
    sc.add_to(ctx)  # kcf_ref_clear + the same sequences again: the old plans are stale all the same
    with pytest.raises(KcfError) as e:
        a.run(db)
    assert e.value.code == -5 and "kcf_ref_clear" in e.value.msg
    a.close()
    b.close()
    c = ctx.plan(31, wins, segs)
    c.run(db)
    assert_results_equal(c.fetch(), want)
    c.close()
    db.close()


@pytest.mark.parametrize("k,P,L,cs,both", [(33, 5, 9, 1, True), (48, 8, 9, 1, True), (63, 7, 9, 2, True), (64, 8, 9, 1, True), (64, 4, 7, 3, False),
                                            (40, 0, 5, 1, True)])
def test_kmers_longer_than_32_bases(ctx, k, P, L, cs, both):
    """k = 33 .. 64 on the device (two 64-bit bit planes per key, 7 / 6 slots per line, home line by a hash of
    the key) against the C oracle, which follows Kmer.java's long[] words (Kmer.java:232-252, 300-338, 406-414): windows with N
    runs, lower case, an inverted stretch found through the other strand, several segments, per-k-mer counts, getCount of every
    record's own text; k = 65 is refused."""
    rng = np.random.default_rng(2000 + k + cs)
    n = 9000
    ref = "".join("ACGT"[c] for c in rng.integers(0, 4, n))
    ref = ref[:700] + "NNNN" + ref[704:1500].lower() + "R" + ref[1501:5000] + "N" * 70 + ref[5070:]
    qry = list(ref.upper().replace("N", "A").replace("R", "G"))
    for pos in rng.integers(0, len(qry), 60):
        qry[pos] = "ACGT"[(("ACGT".index(qry[pos])) + 1 + int(rng.integers(0, 3))) % 4]
    qry = "".join(qry[:1800] + qry[1900:])
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    qry = qry[:400] + "".join(comp[c] for c in reversed(qry[400:900])) + qry[900:]
    img = synth.kmc_image_from_strings([qry], k=k, P=P, L=L, n_bins=8, counter_size=cs, both_strands=both, seed=k)
    rec = np.concatenate([np.frombuffer(b">chrL\n", np.uint8), np.frombuffer("".join(ref[i:i + 61] + "\n" for i in range(0, n, 61)).encode(), np.uint8)])
    raw = rec[6:]
    ctx.ref_clear()
    ctx.ref_add(raw, 61, 62, n)
    db = KMC(ctx, pre=img.pre, suf=img.suf)
    assert db.getKmerLength() == k and db.info.resident_kmers == img.total and db.info.unreachable_kmers == 0
    odb = ob.OracleKMC(img.pre, img.suf)
    lists = [[(0, s, 1500)] for s in range(0, n - 1500, 1100)] + [[(0, 0, n)], [(0, 100, 40), (0, 300, 45), (0, 2000, 700)], [(0, 10, k - 1)], [(0, 20, k)]]
    wins, segs = windows_from_lists(lists)
    rc, want = odb.screen([(raw, 61, 62, n)], wins, segs, min_count=1, threads=4)
    assert rc == 0 and 0 < want["obs"].sum() < want["total_kmers"].sum()
    got = ctx.screen(db, wins, segs)
    assert_results_equal(got, want)
    rc, want2 = odb.screen([(raw, 61, 62, n)], wins, segs, min_count=3, w=(0.5, 0.25, 0.25), threads=4)
    assert_results_equal(ctx.screen(db, wins, segs, min_count=3, weights=(0.5, 0.25, 0.25)), want2)
    # per-k-mer counts of one window, and KMC.getCount of records under their own text (either strand)
    plan = ctx.plan(k, wins, segs)
    plan.run(db)
    plan.fetch()
    rc, text = ob.get_sequence(raw, 61, 62, n, 0, 1500)
    rc, res, counts = odb.process_window(text, want_counts=True)
    gotc = plan.window_counts(db, 0)
    assert gotc.size == counts.size and (gotc == counts).all() and (counts > 0).any() and (counts == 0).any()
    plan.close()
    kms = [qry[i:i + k] for i in range(0, len(qry) - k, 53)]
    wantc = np.array([odb.count(s) for s in kms], np.int32)
    assert (db.getCounts(kms) == wantc).all()
    if both:
        assert (db.getCounts(["".join(comp[c] for c in reversed(s)) for s in kms]) == wantc).all() and (wantc > 0).all()
    db.close()


def test_k_above_64_is_refused(ctx):
    img = synth.kmc_image_from_strings(["ACGT" * 100], k=65, P=5, L=9, n_bins=4, counter_size=1)
    with pytest.raises(KcfError) as e:
        KMC(ctx, pre=img.pre, suf=img.suf)
    assert e.value.code == -4
