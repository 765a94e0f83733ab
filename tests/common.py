"""shared scenario builders for the tests (synthetic; seeds fixed)."""
from __future__ import annotations

import numpy as np
import torch

from tools import synth
from kcftools_b200._lib import SEGMENT_DTYPE, WINDOW_DTYPE


class Scenario:
    """a reference FASTA (several sequences, N runs, lower case, IUPAC), a mutated query, its KMC image."""

    def __init__(self, seq_lens=(60_000, 23_457), k=31, P=7, L=9, n_bins=64, counter_size=1, seed=1, line=60,
                 both_strands=True, snp=0.01, indel=0.001, n_runs=3, coverage=8.0, device="cpu"):
        self.k = k
        recs, self.codes = [], []
        qs = []
        for i, n in enumerate(seq_lens):
            g = synth.random_genome(n, seed * 1000 + i, device)
            self.codes.append(g)
            qs.append(synth.mutate(g, seed * 1000 + 500 + i, snp=snp, indel=indel, big_deletions=1 if n > 40_000 else 0,
                                   big_len=max(200, n // 20), replace_len=max(100, n // 30) if n > 4000 else 0))
            lower = synth.random_intervals(n, 4, 10, max(11, n // 50), seed + 10 + i)
            nr = synth.random_intervals(n, n_runs, 1, max(2, min(400, n // 20)), seed + 20 + i)
            other = synth.random_intervals(n, 2, 1, 3, seed + 30 + i)
            rec = synth.fasta_record(g, f"chr{i + 1} synthetic", line=line, lower=lower, n_runs=nr, other=other)
            recs.append((f"chr{i + 1} synthetic", rec, n, line))
        self.fasta = synth.fasta_image(recs)
        # header name length includes the description: offsets computed from the full header line
        self.kmc = synth.kmc_image_from_genomes(qs, k=k, P=P, L=L, n_bins=n_bins, counter_size=counter_size,
                                                both_strands=both_strands, coverage=coverage, seed=seed + 99)
        self.seq_lens = list(seq_lens)

    def seqs(self):
        f = self.fasta
        return [(f.seq_bytes(i), f.line_bases[i], f.line_width[i], f.lengths[i]) for i in range(len(f.names))]

    def add_to(self, ctx):
        ctx.ref_clear()
        for (raw, lb, lw, sl) in self.seqs():
            ctx.ref_add(raw, lb, lw, sl)


def windows_from_lists(seg_lists):
    """seg_lists: per window a list of (seq_id, start0, len)."""
    wins = np.zeros(len(seg_lists), WINDOW_DTYPE)
    segs = []
    for i, sl in enumerate(seg_lists):
        wins[i] = (len(segs), len(sl))
        segs.extend(sl)
    s = np.zeros(len(segs), SEGMENT_DTYPE)
    for i, t in enumerate(segs):
        s[i] = t
    return wins, s


INT_FIELDS = ("total_kmers", "eff_len", "obs", "variations", "inner", "left", "right", "kmer_count_sum")


def assert_results_equal(got, want, rtol=1e-9):
    for f in INT_FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{f}: first mismatch at window {bad[0]}: got {got[f][bad[0]]} want {want[f][bad[0]]} ({bad.size} windows differ)"
    # score: |got - want| <= rtol * |want| (north_star: 1e-9 relative)
    err = np.abs(got["score"] - want["score"])
    tol = rtol * np.abs(want["score"])
    assert np.all(err <= tol), f"score mismatch: max abs err {err.max()}"
