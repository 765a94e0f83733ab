"""CPU tests: the C oracle against the hand-derived known answers of SURVEY.md §8(c), the doc snippets the
reference publishes, and the independent Python restatement.  PARITY UNPINNED (no reference tests exist)."""
import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyoracle
from tools import synth
from common import Scenario


@pytest.mark.parametrize("L,special,total,distinct", [(5, 170, 535916, 428), (7, 2904, 143896476, 6741),
                                                      (9, 53560, 38671249344, 104293)])
def test_norm_table_kat(L, special, total, distinct):
    n = ob.norm_table(L)
    assert int((n == (1 << (2 * L))).sum()) == special
    assert int(n.astype(np.int64).sum()) == total
    assert len(np.unique(n)) == distinct
    assert (n == synth.norm_table(L)).all()


def test_norm_table_vs_python_restatement():
    for L in (3, 4, 5, 6):
        assert (np.array(pyoracle.norm_table(L)) == ob.norm_table(L)).all()


def test_norm_rc_symmetric():
    # Q8: norm[m] == norm[rc(m)], which is why the signature can be taken on either strand
    L = 7
    n = ob.norm_table(L)
    m = np.arange(1 << (2 * L))
    rev = np.zeros_like(m)
    t = m.copy()
    for _ in range(L):
        rev = (rev << 2) | ((~t) & 3)
        t >>= 2
    assert (n[m] == n[rev]).all()


def test_gap_machine_kats():
    assert ob.gap_machine([0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0], 5) == (12, 5, 4, 2, 2, 1)
    assert ob.gap_machine([0] * 7, 5) == (7, 0, 1, 0, 0, 7)
    assert ob.gap_machine([1] * 7, 5) == (7, 7, 0, 0, 0, 0)
    assert ob.gap_machine([], 5) == (0, 0, 0, 0, 0, 0)


def test_get_distance_kat_k31():
    # Q4: g=1->28, 29->0, 30->1, 31->1, 32->2, 61->31 (inner gap between two hits)
    for g, d in [(1, 28), (29, 0), (30, 1), (31, 1), (32, 2), (61, 31)]:
        assert ob.gap_machine([1] + [0] * g + [1], 31)[3] == d
        assert pyoracle.get_distance(31, g) == d
    # GetVariants.java:265 comment: k=3, 3 missing k-mers -> distance 1
    assert pyoracle.get_distance(3, 3) == 1


def test_tiling_windows_doc_snippets():
    s, e = ob.windows_fixed(3000, 1000, 0, 31)
    assert list(zip(s, e)) == [(0, 1000), (970, 1970), (1940, 2940), (2910, 3000)]
    # docs/formats/gttable.md:25-28 (W=1000, k=31): starts 0, 970, 1940, 2910
    assert list(ob.windows_fixed(100000, 1000, 0, 31)[0][:4]) == [0, 970, 1940, 2910]
    # docs/formats/attributes.md:27-29 (W=5000, k=32): chr1_0, chr1_4969, chr1_9938
    assert list(ob.windows_fixed(100000, 5000, 0, 32)[0][:3]) == [0, 4969, 9938]
    # C1 = 10 Mb, W=50000, k=31 -> 201 windows, last two (9944030-9994030), (9994000-10000000)
    s, e = ob.windows_fixed(10_000_000, 50_000, 0, 31)
    assert len(s) == 201 and (s[-2], e[-2], s[-1], e[-1]) == (9944030, 9994030, 9994000, 10000000)
    # sliding
    s, e = ob.windows_fixed(1000, 400, 300, 31)
    assert list(zip(s, e)) == [(0, 400), (300, 700), (600, 1000), (900, 1000)]
    assert pyoracle.windows_fixed(1000, 400, 300, 31) == list(zip(s, e))
    assert pyoracle.windows_fixed(3000, 1000, 0, 31) == [(0, 1000), (970, 1970), (1940, 2940), (2910, 3000)]


def test_score_kat():
    rc, s = ob.compute_score(49000, 49970, 50000, 120, 30, 5)
    assert rc == 0 and s == 99.13053412047228
    assert ob.compute_score(0, 10, 10, 0, 0, 10) == (0, 0.0)
    rc, _ = ob.compute_score(5, 10, 40, 0, 0, 0, (0.3, 0.3, 0.5))
    assert rc != 0  # "Weights should sum to 1.0" is fatal in the reference


def test_get_sequence_newlines_and_trailing():
    g = synth.random_genome(250, 3)
    rec = synth.fasta_record(g, "s", line=60)
    raw = rec[3:]
    txt = "".join("ACGT"[c] for c in g.tolist())
    for (st, ln) in [(0, 250), (59, 2), (60, 60), (119, 131), (240, 10), (0, 1)]:
        rc, got = ob.get_sequence(raw, 60, 61, 250, st, ln)
        assert rc == 0 and got.decode() == txt[st:st + ln]
        assert pyoracle.get_sequence(bytes(raw), 60, 61, 250, st, ln) == txt[st:st + ln]
    # Q9: without the trailing newline the final bases are unreadable (fatal in the reference)
    raw2 = synth.fasta_record(g, "s", line=60, trailing_newline=False)[3:]
    assert ob.get_sequence(raw2, 60, 61, 250, 240, 10)[0] != 0
    assert ob.get_sequence(raw2, 60, 61, 250, 200, 10)[0] == 0
    with pytest.raises(ValueError):
        pyoracle.get_sequence(bytes(raw2), 60, 61, 250, 240, 10)
    # invalid ranges
    assert ob.get_sequence(raw, 60, 61, 250, 10, 0)[0] != 0
    assert ob.get_sequence(raw, 60, 61, 250, 245, 10)[0] != 0


@pytest.mark.parametrize("k,P,L,cs,both", [(31, 7, 9, 1, True), (32, 8, 9, 2, True), (21, 5, 7, 1, False), (13, 1, 5, 4, True)])
def test_db_writer_roundtrip_and_python_agreement(k, P, L, cs, both):
    sc = Scenario(seq_lens=(4000, 1500), k=k, P=P, L=L, n_bins=16, counter_size=cs, both_strands=both, seed=7)
    db = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    py = pyoracle.PyKMC(sc.kmc.pre.tobytes(), sc.kmc.suf.tobytes())
    assert db.info.kmer_length == k and db.info.total_kmers == sc.kmc.total == py.total
    assert bool(db.info.both_strands) == both == py.both_strands
    seqs = sc.seqs()
    raw, lb, lw, sl = seqs[0]
    rc, text = ob.get_sequence(raw, lb, lw, sl, 0, sl)
    assert rc == 0
    text = text.decode()
    rc, res, counts = db.process_window(text.encode(), want_counts=True)
    assert rc == 0
    want = pyoracle.process_window(py, text)
    for f in ("total_kmers", "eff_len", "obs", "variations", "inner", "left", "right", "kmer_count_sum"):
        assert getattr(res, f) == want[f], f
    assert res.score == want["score"]
    kms = pyoracle.kmers_list(text, k)
    assert len(kms) == res.total_kmers
    pyc = [py.get_count(py.canonical(s)) for s in kms[:400]]
    assert list(counts[:400]) == pyc
    assert res.obs > 0 and res.obs < res.total_kmers  # the scenario exercises both hits and misses


def test_every_db_kmer_is_found_with_its_count():
    # the writer's signature / bin / LUT layout is exactly what the reference's lookup walks
    g = synth.random_genome(3000, 11)
    img = synth.kmc_image_from_genomes([g], k=31, P=7, L=9, n_bins=32, seed=5)
    db = ob.OracleKMC(img.pre, img.suf)
    text = "".join("ACGT"[c] for c in g.tolist())
    rc, res, counts = db.process_window(text.encode(), want_counts=True)
    assert rc == 0 and res.total_kmers == 3000 - 30
    # Poisson(8) zero-drops are the only misses
    assert (counts == 0).sum() < 10 and res.obs == (counts > 0).sum()
    assert res.kmer_count_sum == counts.sum()


def test_screen_threads_equal_single():
    sc = Scenario(seq_lens=(9000, 5000), seed=3, n_bins=16)
    db = ob.OracleKMC(sc.kmc.pre, sc.kmc.suf)
    from kcftools_b200.api import fixed_windows
    wins, segs, *_ = fixed_windows(sc.seq_lens, 1000, 0, 31)
    rc1, a = db.screen(sc.seqs(), wins, segs, threads=1)
    rc2, b = db.screen(sc.seqs(), wins, segs, threads=4)
    assert rc1 == 0 and rc2 == 0 and (a == b).all() and a["total_kmers"].sum() > 0


@pytest.mark.parametrize("k,P,L,cs,both", [(33, 5, 9, 1, True), (41, 5, 7, 2, True), (64, 8, 9, 1, True), (65, 5, 9, 1, False), (100, 8, 11, 3, True),
                                            (31, 7, 9, 1, True)])
def test_multiword_kmers_c_oracle_vs_python(k, P, L, cs, both):
    """SURVEY §8 row f4 (k > 32; the CUDA path follows up to k = 64, tests/test_gpu_parity.py): the C restatement follows Kmer.java's long[]
    words (32 bases per word, word-by-word unsigned canonical compare) and agrees with the string-based Python restatement
    on databases written by the any-k generator — canonical orientation across word boundaries, signature, prefix / suffix
    bytes, binary search, gap statistics and score."""
    rng = np.random.default_rng(1000 + k)
    ref = "".join("ACGT"[c] for c in rng.integers(0, 4, 2600))
    ref = ref[:700] + "NNNN" + ref[700:1500].lower() + "R" + ref[1500:]
    # the sample: the reference with SNPs, a deletion, and a stretch in the opposite orientation
    qry = list(ref.upper().replace("N", "A").replace("R", "G"))
    for pos in rng.integers(0, len(qry), 25):
        qry[pos] = "ACGT"[(("ACGT".index(qry[pos])) + 1 + int(rng.integers(0, 3))) % 4]
    qry = "".join(qry[:1800] + qry[1900:])
    qry = qry[:400] + pyoracle.revcomp(qry[400:900]) + qry[900:]
    img = synth.kmc_image_from_strings([qry], k=k, P=P, L=L, n_bins=8, counter_size=cs, both_strands=both, seed=k)
    db = ob.OracleKMC(img.pre, img.suf)
    py = pyoracle.PyKMC(img.pre.tobytes(), img.suf.tobytes())
    assert db.info.kmer_length == k == py.k and db.info.total_kmers == img.total == py.total
    rc, res, counts = db.process_window(ref.encode(), min_count=1, want_counts=True)
    assert rc == 0
    want = pyoracle.process_window(py, ref)
    for f in ("total_kmers", "eff_len", "obs", "variations", "inner", "left", "right", "kmer_count_sum"):
        assert getattr(res, f) == want[f], f
    assert res.score == want["score"]
    kms = pyoracle.kmers_list(ref, k)
    assert list(counts[:len(kms)]) == [py.get_count(py.canonical(s)) for s in kms]
    assert 0 < res.obs < res.total_kmers
    if both:  # the inverted stretch is found through the reverse-complement orientation
        inv = pyoracle.kmers_list(ref[450:850], k)
        assert sum(py.get_count(py.canonical(s)) > 0 for s in inv) > len(inv) // 2
    # every record of the database is found under its own text with its own count
    some = list(pyoracle.kmers_list(qry, k))[::37]
    for s in some:
        c = db.count(py.canonical(s))
        assert c == py.get_count(py.canonical(s)) and (c > 0 or cs == 0)
