"""profiling helper: the partitioned-database path with `world` ranks in lockstep on ONE GPU (c2s workload), so that a
single-process ncu launch list shows what every phase costs.  Not a benchmark."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kcftools_b200 import shard  # noqa: E402
from kcftools_b200.api import Context, KMC, fixed_windows  # noqa: E402
from kcftools_b200.partitioned import screen_partitioned_local  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl_ = bench.build_workload("c2s", "cuda:0")
fasta, kmc, window, desc = wl_.fasta, wl_.kmc, wl_.window, wl_.desc
wins, segs, *_ = fixed_windows(fasta.lengths, window, 0, 31)
ranges = shard.partition(shard.window_lengths(wins, segs), world)
ranks = []
for r in range(world):
    c = Context(0)
    for i in range(len(fasta.names)):
        c.ref_add(fasta.seq_bytes(i), fasta.line_bases[i], fasta.line_width[i], fasta.lengths[i])
    c.set_partition(r, world)
    db = KMC(c, pre=kmc.pre, suf=kmc.suf, placement=1)
    lw, ls = shard.local_slice(wins, segs, *ranges[r])
    ranks.append((c, db, c.plan(31, lw, ls)))
for it in range(3):
    parts = screen_partitioned_local(ranks)
torch.cuda.synchronize()
print("k-mers", int(sum(p["total_kmers"].sum() for p in parts)), "observed", int(sum(p["obs"].sum() for p in parts)))
