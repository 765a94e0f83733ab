#!/usr/bin/env python
"""C5-shaped pipeline through the command line, scaled down (BASELINE.json configs[4]: N query databases against one
reference, feeding cohort -> findIBS -> kcf2gt):

  getVariations -k db0,...,dbN-1   (reference resident once, every database screened in turn, cohort matrix filled device
                                    to device, ONE cohort KCF written)
  findIBS --summary --bed, kcf2gt  on that file

and, for comparison, the file pipeline the reference prescribes: N x getVariations -> cohort.  Prints one JSON line with
the wall times (process start to exit, files in the page cache) and checks that both routes write the same rows.

  python tools/pipeline_c5s.py [--samples 4] [--chrom-len 7500000] [--chroms 12] [--keep DIR]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--chroms", type=int, default=12)
    ap.add_argument("--chrom-len", type=int, default=7_500_000)
    ap.add_argument("--window", type=int, default=50_000)
    ap.add_argument("--keep", default=None)
    ap.add_argument("--devices", default=None, help="e.g. 0,1,2,3: also time the multi-database run with the samples shared out over these GPUs")
    args = ap.parse_args()
    import torch
    from tools import synth
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    cli = os.path.join(ROOT, "kcftools_b200", "host", "kcftools_b200")
    d = args.keep or tempfile.mkdtemp(prefix="kcfc5s")
    os.makedirs(d, exist_ok=True)
    recs, genomes = [], []
    for i in range(args.chroms):
        g = synth.random_genome(args.chrom_len, 5001 + i, dev)
        genomes.append(g)
        nr = synth.random_intervals(args.chrom_len, 5, 100, 5_000, 5101 + i)
        recs.append((f"chr{i + 1:02d}", synth.fasta_record(g, f"chr{i + 1:02d}", line=60, n_runs=nr), args.chrom_len, 60))
    fa = os.path.join(d, "ref.fa")
    synth.fasta_image(recs).write(fa)
    del recs
    prefixes, names = [], []
    for s in range(args.samples):
        snp = 0.002 * (1 + 3 * s)  # samples of increasing divergence
        qs = [synth.mutate(g, 6000 + 100 * s + i, snp=snp) for i, g in enumerate(genomes)]
        pre = os.path.join(d, f"sample{s}")
        synth.kmc_image_from_genomes(qs, k=31, P=7, L=9, n_bins=512, counter_size=1, coverage=8.0, seed=700 + s).write(pre)
        del qs
        prefixes.append(pre)
        names.append(f"sample{s}")
    del genomes
    if dev != "cpu":
        torch.cuda.empty_cache()

    def run(*a):
        t = time.perf_counter()
        subprocess.run([cli, *a], check=True, stdout=subprocess.DEVNULL)
        return time.perf_counter() - t

    w = ["-f", "window", "-w", str(args.window)]
    run("getVariations", "-r", fa, "-k", prefixes[0], "-o", os.path.join(d, "warm.kcf"), "-s", "warm", *w)  # .faidx + page cache
    direct = os.path.join(d, "cohort_direct.kcf")
    t_direct = run("getVariations", "-r", fa, "-k", ",".join(prefixes), "-o", direct, "-s", ",".join(names), *w)
    t_spread = None
    if args.devices:
        spread = os.path.join(d, "cohort_spread.kcf")
        t_spread = run("getVariations", "-r", fa, "-k", ",".join(prefixes), "-o", spread, "-s", ",".join(names), "--devices", args.devices, *w)
    singles, t_single = [], []
    for pre, nm in zip(prefixes, names):
        o = os.path.join(d, nm + ".kcf")
        t_single.append(run("getVariations", "-r", fa, "-k", pre, "-o", o, "-s", nm, *w))
        singles.append(o)
    merged = os.path.join(d, "cohort_files.kcf")
    t_cohort = run("cohort", "-i", ",".join(singles), "-o", merged)
    t_ibs = run("findIBS", "-i", direct, "-o", os.path.join(d, "ibs.kcf"), "--summary", "--bed")
    t_gt = run("kcf2gt", "-i", direct, "-o", os.path.join(d, "gt.tsv"))

    def body(p):
        return [l for l in open(p) if not l.startswith("##date=") and not l.startswith("##CMD=")]
    rows = [l for l in body(direct) if not l.startswith("#")]
    out = {"samples": args.samples, "reference_bp": args.chroms * args.chrom_len, "windows": len(rows),
           "kmers_screened": sum(int(r.split("\t")[4]) for r in rows) * args.samples,
           "wall_s": {"getVariations_all_databases_to_cohort": round(t_direct, 3), "getVariations_per_sample": [round(t, 3) for t in t_single],
                      "cohort_from_files": round(t_cohort, 3), "findIBS_summary_bed": round(t_ibs, 3), "kcf2gt": round(t_gt, 3)},
           "direct_equals_file_pipeline": body(direct) == body(merged),
           "wall_s_samples_over_devices": None if t_spread is None else {"devices": args.devices, "getVariations_all_databases_to_cohort": round(t_spread, 3),
                                                                          "equals_file_pipeline": body(os.path.join(d, "cohort_spread.kcf")) == body(merged)},
           "ibs_blocks": len(open(os.path.join(d, "ibs.summary.tsv")).read().strip().split("\n")) - 1,
           "genotype_rows": len(open(os.path.join(d, "gt.tsv")).read().strip().split("\n")) - 2}
    print(json.dumps(out), flush=True)
    if not args.keep:
        shutil.rmtree(d, ignore_errors=True)
    return 0 if out["direct_equals_file_pipeline"] else 1


if __name__ == "__main__":
    sys.exit(main())
