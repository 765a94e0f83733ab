#!/usr/bin/env python
"""Full-size check of the reference side of BASELINE.json configs[3] (C4: 15 Gb, 21 sequences, ~3e5 windows of 50 kb)
on ONE GPU, through a size-independent property instead of the CPU oracle (which would need hours):

  sequence j of the big reference = chromosome (j mod 12) of the c2 workload (75 Mb) followed by unrelated random bases
  up to --seq-len.  Tiling windows start at 0, so every window that ends inside the 75 Mb core is THE SAME window as in
  the c2 run and must produce the same row, bit for bit (all integer columns and the score); the windows of the random
  tail contain no N, so TOTAL_KMERS = length - k + 1 and EFFLEN = length exactly, and (random 31-mers against a 9e8-record
  database) essentially nothing is observed.

Also times the 1.5e10-position screening launch.  The DATABASE stays at c2 size (8.96e8 records): a 1.5e10-record image
cannot be synthesised in the box's host memory; table-side scale is covered by the partitioned-placement runs.

  python tools/scale_c4ref.py [--seqs 21] [--seq-len 714000000]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=21)
    ap.add_argument("--seq-len", type=int, default=714_000_000)
    ap.add_argument("--core", default="c2", help="bench workload whose chromosomes are the cores (c2 or c2s)")
    ap.add_argument("--window", type=int, default=50_000)
    args = ap.parse_args()
    import torch
    import bench
    from kcftools_b200.api import KMC, Context, fixed_windows
    from tools import synth
    dev = "cuda:0"
    wl_ = bench.build_workload(args.core, dev)
    fasta, kmc, window = wl_.fasta, wl_.kmc, wl_.window
    assert window == args.window
    core_len = fasta.lengths[0]
    n_core = len(fasta.names)
    assert core_len % 60 == 0 and args.seq_len >= core_len and (args.seq_len - core_len) % 60 == 0
    out = {"core": args.core, "seqs": args.seqs, "seq_len": args.seq_len}
    with Context(0) as ctx:
        db = KMC(ctx, pre=kmc.pre, suf=kmc.suf)
        for i in range(n_core):
            ctx.ref_add(fasta.seq_bytes(i), fasta.line_bases[i], fasta.line_width[i], fasta.lengths[i])
        wins, segs, starts, ends, sids = fixed_windows(fasta.lengths, window, 0, 31)
        plan = ctx.plan(31, wins, segs)
        plan.run(db)
        core_rows = plan.fetch().copy()
        plan.close()
        core_first = np.searchsorted(sids, np.arange(n_core + 1))
        ctx.ref_clear()
        t0 = time.time()
        for j in range(args.seqs):
            c = j % n_core
            tail_len = args.seq_len - core_len
            body = fasta.seq_bytes(c)[:core_len // 60 * 61]  # the core's folded lines only (the mapped slice runs on into the next header)
            if tail_len:
                tail = synth.fasta_record(synth.random_genome(tail_len, 9100 + j, dev), "t", line=60)[3:]  # drop ">t\n"
                raw = np.concatenate([body, tail])
                del tail
            else:
                raw = body
            ctx.ref_add(raw, 60, 61, args.seq_len)
            del raw
        out["reference_build_upload_s"] = round(time.time() - t0, 1)
        bw, bs, bstarts, bends, bsids = fixed_windows([args.seq_len] * args.seqs, window, 0, 31)
        plan = ctx.plan(31, bw, bs)
        ctx.set_profiling(True)
        for _ in range(2):
            plan.run(db)
        rows = plan.fetch()
        ms = ctx.last_kernel_ms()[0]
        total = int(rows["total_kmers"].astype(np.int64).sum())
        out.update({"windows": int(bw.size), "positions": int(plan.n_positions), "kmers_screened": total, "screen_kernel_ms": round(float(ms), 3),
                    "kmers_per_s": total / (ms * 1e-3)})
        # property 1: windows inside the core equal the c2 rows
        big_first = np.searchsorted(bsids, np.arange(args.seqs + 1))
        compared = 0
        same = True
        for j in range(args.seqs):
            c = j % n_core
            inside = int(np.sum(bends[big_first[j]:big_first[j + 1]] <= core_len))
            n_cmp = min(inside, core_first[c + 1] - core_first[c] - 1)  # the core's own last window is cut at its end
            a = rows[big_first[j]:big_first[j] + n_cmp]
            b = core_rows[core_first[c]:core_first[c] + n_cmp]
            assert (bstarts[big_first[j]:big_first[j] + n_cmp] == starts[core_first[c]:core_first[c] + n_cmp]).all()
            same = same and bool((a == b).all())
            compared += n_cmp
        out["core_windows_compared"] = compared
        out["core_windows_identical"] = same
        # property 2: windows entirely inside the random tail (no N, nothing of the database)
        tail_mask = bstarts >= core_len
        tl = (bends - bstarts)[tail_mask].astype(np.int64)
        tr = rows[tail_mask]
        out["tail_windows"] = int(tail_mask.sum())
        out["tail_totals_exact"] = bool((tr["total_kmers"] == tl - 30).all() and (tr["eff_len"] == tl).all())
        out["tail_observed_kmers"] = int(tr["obs"].astype(np.int64).sum())
        out["tail_rows_consistent"] = bool(((tr["obs"] > 0) | ((tr["right"] == tr["total_kmers"]) & (tr["variations"] == 1) & (tr["score"] == 0))).all())
        plan.close()
        db.close()
    out["ok"] = bool(out["core_windows_identical"] and out["tail_totals_exact"] and out["tail_rows_consistent"] and out["tail_observed_kmers"] < 1000)
    print(json.dumps(out), flush=True)
    return 0 if out["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
