"""Occupancy of the table lines of the bench database (kcf_db_line_histogram): how many lines hold n keys, and where the keys live.

  python tools/line_histogram.py        (on a GPU box; prints the histogram and the key shares)
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from kcftools_b200.api import Context, KMC
wl_ = bench.build_workload("c2", "cuda:0")
fasta, kmc, window = wl_.fasta, wl_.kmc, wl_.window
ctx = Context(0); db = KMC(ctx, pre=kmc.pre, suf=kmc.suf)
h = np.zeros(16, np.uint64)
ctx._check(ctx._lib.kcf_db_line_histogram(db._h, h.ctypes.data))
h = h.astype(np.float64); n = np.arange(16)
keys = h * n
print("lines by occupancy:", [int(x) for x in h[:14]])
print("share of lines:", np.round(h[:14] / h.sum(), 4).tolist())
print("share of KEYS living in lines with n keys:", np.round(keys[:14] / keys.sum(), 4).tolist())
for p1 in (4, 5, 6, 7, 8):
    # a present k-mer needs phase 2 when it sits in slot >= p1; an absent one when slot p1-1 is occupied
    in_late_slot = sum(h[m] * (m - p1) for m in range(p1 + 1, 14)) / keys.sum()
    line_full = sum(keys[m] for m in range(p1, 14)) / keys.sum()
    print(f"phase-1 slots {p1}: present k-mer beyond phase 1 {in_late_slot:.4f}; probe lands in a line with >= {p1} keys (size-biased) {line_full:.4f}")
