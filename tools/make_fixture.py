"""Regenerates tests/fixtures_oracle/small.npz: a small, fully materialised getVariations case (FASTA image, KMC image, GTF,
window lists) together with the rows the CPU oracle computes for it and the KCF text lines oracle/pyhost.py renders.

The reference itself cannot produce vectors here (Java, no JVM in the image; it ships no tests or fixtures), so these
are ORACLE-generated: they pin the oracle, the GPU path and the host formatting against regressions and against each
other, not against the reference (parity unpinned).

    python tools/make_fixture.py
    python tools/make_fixture.py --write-files DIR    # the case's input FILES (ref.fa, sample.kmc_pre/.kmc_suf, ann.gtf) for a
                                                      # run of the real reference (oracle/build_ref.sh), nothing else
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import windows_from_lists  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import pyhost  # noqa: E402
from tools import synth  # noqa: E402


def main():
    lens = (12_000, 3_000)
    recs, qs = [], []
    for i, n in enumerate(lens):
        g = synth.random_genome(n, 4100 + i)
        qs.append(synth.mutate(g, 4200 + i, snp=0.012, indel=0.002, big_deletions=1 if i == 0 else 0, big_len=600, replace_len=200 if i == 0 else 0))
        nr = synth.random_intervals(n, 2, 3, 120, 4300 + i)
        low = synth.random_intervals(n, 3, 10, 300, 4400 + i)
        oth = synth.random_intervals(n, 2, 1, 2, 4500 + i)
        name = f"chr{i + 1} golden"
        recs.append((name, synth.fasta_record(g, name, line=60, lower=low, n_runs=nr, other=oth), n, 60))
    img = synth.fasta_image(recs)
    kmc = synth.kmc_image_from_genomes(qs, k=31, P=3, L=5, n_bins=4, counter_size=1, coverage=6.0, seed=4600)
    names = ["chr1", "chr2"]
    gtf_text = synth.synthetic_gtf([("chr1", lens[0]), ("chr2", lens[1])], 3, 4700, max_tx=2, max_exons=4, exon_lo=40, exon_hi=400)
    gtf = pyhost.Gtf(gtf_text)
    seqs = [(img.seq_bytes(i), img.line_bases[i], img.line_width[i], img.lengths[i]) for i in range(len(lens))]
    odb = ob.OracleKMC(kmc.pre, kmc.suf)
    out = {"fasta": img.data, "kmc_pre": kmc.pre, "kmc_suf": kmc.suf, "gtf": np.frombuffer(gtf_text.encode(), np.uint8),
           "lens": np.array(lens), "offsets": np.array(img.offsets)}
    if len(sys.argv) == 3 and sys.argv[1] == "--write-files":
        d = sys.argv[2]
        os.makedirs(d, exist_ok=True)
        img.write(os.path.join(d, "ref.fa"))
        kmc.write(os.path.join(d, "sample"))
        open(os.path.join(d, "ann.gtf"), "w").write(gtf_text)
        print(d)
        return
    w = (0.3, 0.3, 0.4)
    for mode, kw in [("window", dict(window=2000, step=0)), ("sliding", dict(window=1500, step=700)), ("gene", {}), ("transcript", {})]:
        feat = "window" if mode in ("window", "sliding") else mode
        ws = pyhost.windows_of(feat, names, list(lens), 31, gtf=gtf, **kw)
        wins, segs = windows_from_lists([x[4] for x in ws])
        rc, res = odb.screen(seqs, wins, segs, min_count=1, w=w, threads=2)
        assert rc == 0
        out[f"{mode}_wins"] = wins
        out[f"{mode}_segs"] = segs
        out[f"{mode}_rows"] = res
        text = "\n".join(pyhost.kcf_row(x[1], x[2], x[3], x[0], res[i], w) for i, x in enumerate(ws)) + "\n"
        out[f"{mode}_kcf"] = np.frombuffer(text.encode(), np.uint8)
    path = os.path.join(ROOT, "tests", "fixtures_oracle", "small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", kmc.total, "records")


if __name__ == "__main__":
    main()
