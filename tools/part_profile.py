"""profiling helper: the partitioned-database path (k-mer exchange over peer memory) with `world` ranks in lockstep on ONE
GPU, so that a single-process ncu run sees every kernel and the phases can be timed without NVLink in the picture.  The
"peer" regions are plain device pointers here.  Not a benchmark: the ranks share one GPU.

  python tools/part_profile.py [world] [workload]      # default 2 c2s; prints ms per phase summed over the ranks
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kcftools_b200 import shard  # noqa: E402
from kcftools_b200.api import Context, KMC, fixed_windows  # noqa: E402
from kcftools_b200.partitioned import BATCH_TILES, Exchange, _finish  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "c2s"
wl_ = bench.build_workload(name, "cuda:0")
fasta, kmc, window = wl_.fasta, wl_.kmc, wl_.window
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

n_tiles = max(p.n_tiles for _, _, p in ranks)
batch = int(min(BATCH_TILES, max(n_tiles, 1)))
xs = [Exchange(c, db, i, world, batch) for i, (c, db, p) in enumerate(ranks)]
ptrs = [x.export()[1] for x in xs]
for x in xs:
    x.connect(pointers=ptrs)
phases = {"send": 0.0, "answer": 0.0, "fold": 0.0}
for it in range(3):
    if it == 1:
        phases = dict.fromkeys(phases, 0.0)  # the first pass warms up
    for t0 in range(0, max(n_tiles, 1), batch):
        t1 = t0 + batch
        for ph, call in (("send", lambda c, db, p, x: c._lib.kcf_xg_send(c._h, db._h, p._h, x._h, t0, t1)),
                         ("answer", lambda c, db, p, x: c._lib.kcf_xg_answer(c._h, db._h, x._h)),
                         ("fold", lambda c, db, p, x: c._lib.kcf_xg_fold(c._h, p._h, x._h, t0, t1, 1))):
            torch.cuda.synchronize()
            ta = time.perf_counter()
            for x, (c, db, p) in zip(xs, ranks):
                c._check(call(c, db, p, x))
            torch.cuda.synchronize()
            phases[ph] += time.perf_counter() - ta
for x in xs:
    x.status()
parts = [_finish(c, p, (0.3, 0.3, 0.4)) for (c, db, p) in ranks]
for x in xs:
    x.close()
kmers = int(sum(p["total_kmers"].sum() for p in parts))
print("workload", name, "world", world, "k-mers", kmers, "observed", int(sum(p["obs"].sum() for p in parts)),
      "ms per job, summed over the ranks sharing the GPU:", {k: round(v / 2 * 1e3, 3) for k, v in phases.items()},
      "lib", os.environ.get("KCF_LIB_PATH", "default"))
