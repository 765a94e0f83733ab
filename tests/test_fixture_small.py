"""tests/fixtures_oracle/small.npz (tools/make_fixture.py): a fully materialised small case with the rows and KCF lines the CPU
oracle produced for it.  NOT reference-authored golden data — the reference ships none and no JVM exists here (DESIGN.md §2,
"parity unpinned") — it pins the oracle and the CUDA path against silent drift between commits, nothing more."""
import os

import numpy as np
import pytest

from common import INT_FIELDS, assert_results_equal
from kcftools_b200._lib import RESULT_DTYPE, SEGMENT_DTYPE, WINDOW_DTYPE
from oracle import binding as ob
from oracle import pyhost

G = np.load(os.path.join(os.path.dirname(__file__), "fixtures_oracle", "small.npz"))
MODES = ["window", "sliding", "gene", "transcript"]


def _seqs():
    lens, offs, data = G["lens"], G["offsets"], G["fasta"]
    ends = list(offs[1:]) + [data.size]
    out = []
    for i, n in enumerate(lens):
        a = int(offs[i])
        b = int(ends[i]) if i + 1 == len(lens) else int(np.flatnonzero(data[:int(ends[i])] == ord(">"))[-1])
        out.append((data[a:b], 60, 61, int(n)))
    return out


@pytest.mark.parametrize("mode", MODES)
def test_oracle_reproduces_fixture_rows_and_text(mode):
    odb = ob.OracleKMC(G["kmc_pre"], G["kmc_suf"])
    wins, segs = G[f"{mode}_wins"].view(WINDOW_DTYPE), G[f"{mode}_segs"].view(SEGMENT_DTYPE)
    rc, res = odb.screen(_seqs(), wins, segs, threads=2)
    assert rc == 0
    want = G[f"{mode}_rows"].view(RESULT_DTYPE)
    assert_results_equal(res, want)
    assert want["obs"].sum() > 0 and (want["variations"] > 0).any()
    # the KCF text is a pure function of the integer columns (Data.java:70-107, Window.java:125-152)
    lines = G[f"{mode}_kcf"].tobytes().decode().strip("\n").split("\n")
    assert len(lines) == want.size
    for ln, r in zip(lines, want):
        f = ln.split("\t")
        assert pyhost.kcf_row(f[0], int(f[1]), int(f[2]), f[3], r) == ln
        assert int(f[4]) == r["total_kmers"] and f[6] == "GT:VA:OB:ID:LD:RD:KD:SC"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", MODES)
def test_gpu_reproduces_fixture_rows(ctx, mode):
    from kcftools_b200.api import KMC
    ctx.ref_clear()
    for (raw, lb, lw, n) in _seqs():
        ctx.ref_add(raw, lb, lw, n)
    db = KMC(ctx, pre=G["kmc_pre"], suf=G["kmc_suf"])
    got = ctx.screen(db, G[f"{mode}_wins"].view(WINDOW_DTYPE), G[f"{mode}_segs"].view(SEGMENT_DTYPE))
    assert_results_equal(got, G[f"{mode}_rows"].view(RESULT_DTYPE))
    db.close()
