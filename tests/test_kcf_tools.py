"""Rows f1-f3 of SURVEY §8(f): the KCF reader and the cohort / findIBS / kcf2gt consumers of getVariations output.

CPU part (not gpu): the Python restatement (oracle/pykcf.py) against hand-derived known answers — the reference ships no
tests for these paths — and the C++ header parser through a CLI hook.  GPU part: the C++ commands (numbers computed by
the kcf_cohort_* kernels) against the restatement, byte for byte except the ##date / ##CMD lines."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyhost, pykcf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "kcftools_b200", "host")
CLI = os.path.join(HOST_DIR, "kcftools_b200")
W = (0.3, 0.3, 0.4)


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(os.path.join(ROOT, "kcftools_b200", "libkcfgpu.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "kcftools_b200", "csrc"), "-j4"])
    subprocess.check_call(["make", "-C", HOST_DIR], stdout=subprocess.DEVNULL)
    return CLI


def run(cli, *args, check=True):
    p = subprocess.run([cli, *args], capture_output=True, text=True)
    if check:
        assert p.returncode == 0, p.stdout + p.stderr
    return p


def strip_volatile(text: str) -> str:
    return "".join(l + "\n" for l in text.split("\n")[:-1] if not l.startswith("##date=") and not l.startswith("##CMD="))


# ---------------------------------------------------------------------------------------------- synthetic KCF files
CHROMS = [("chr1", 60_000), ("chr2", 45_000), ("chr10", 30_000), ("scaffold_7", 20_000), ("chrUn", 12_000)]


def synth_results(seed: int, n: int, good: float):
    """random but self-consistent per-window integers: `good` = share of windows that look identical to the reference"""
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=[("total_kmers", "i4"), ("eff_len", "i4"), ("obs", "i4"), ("variations", "i4"), ("inner", "i4"), ("left", "i4"),
                           ("right", "i4"), ("kmer_count_sum", "i8")])
    for i in range(n):
        total = int(rng.integers(500, 5000))
        kind = rng.random()
        if kind < good:
            miss = int(rng.integers(0, total // 50 + 1))
        elif kind < good + 0.1:
            miss = total                          # nothing observed: score 0
        else:
            miss = int(rng.integers(total // 20, total))
        obs = total - miss
        left = int(rng.integers(0, miss + 1)) if obs else 0
        right = int(rng.integers(0, miss - left + 1)) if obs else total
        inner = (miss - left - right) if obs else 0
        var = (1 if left else 0) + (1 if right else 0) + (int(rng.integers(1, 8)) if inner else 0)
        r[i] = (total, total + 30, obs, var, inner, left, right, int(obs * rng.uniform(3.0, 12.0)))
    return r


def windows_list(window: int = 5000):
    out = []
    for name, n in CHROMS:
        j = 0
        for s in range(0, n, window):
            out.append((name, s, min(s + window, n), f"{name}_{j}"))
            j += 1
    return out


def sample_kcf(sample: str, seed: int, good: float, window: int = 5000, step: int = 0, weights=W, totals_from=None) -> str:
    wl = windows_list(window)
    res = synth_results(seed, len(wl), good)
    if totals_from is not None:  # the samples of a cohort share TOTAL_KMERS / EFFLEN (same reference windows)
        for f in ("total_kmers", "eff_len"):
            res[f] = totals_from[f]
        res["obs"] = np.minimum(res["obs"], res["total_kmers"])
    hdr = pykcf.KcfHeader(reference="ref.fa", contigs=dict(CHROMS), cmds=[f"kcftools getVariations -s {sample}"], samples=[sample])
    vals = [str(window), str(step), "31", "false", str(len(wl)), pykcf.java_double_str(weights[0]), pykcf.java_double_str(weights[1]),
            pykcf.java_double_str(weights[2])]
    hdr.params = [(k, v) for k, v in zip(pykcf.PARAM_KEYS, vals)]
    rows = "".join(pyhost.kcf_row(c, s, e, wid, res[i], weights) + "\n" for i, (c, s, e, wid) in enumerate(wl))
    return hdr.text("2026-01-01") + rows, res


@pytest.fixture(scope="module")
def cohort_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("kcf")
    texts, paths = [], []
    base = None
    for i, (name, good) in enumerate([("alpha", 0.8), ("beta", 0.5), ("gamma", 0.2), ("delta", 0.95)]):
        t, res = sample_kcf(name, 40 + i, good, totals_from=base)
        base = res if base is None else base
        p = str(d / f"{name}.kcf")
        open(p, "w").write(t)
        texts.append(t)
        paths.append(p)
    return {"dir": d, "texts": texts, "paths": paths}


# ---------------------------------------------------------------------------------------------- CPU: known answers
def test_java_hashmap_order_known_answers():
    # String.hashCode: "a" = 97, "b" = 98, ...; bucket = (h ^ h >>> 16) & 15: single letters keep their alphabetical order
    assert pykcf.java_hashmap_order(["c", "a", "b"]) == ["a", "b", "c"]
    # by hand: "chr1".hashCode() = ((99*31 + 104)*31 + 114)*31 + 49 = 3052836 = 0x2E9524; ^ (>>> 16 = 0x2E) = 0x2E950A; & 15 = 10
    def bucket(s, cap=16):
        h = 0
        for ch in s:
            h = (31 * h + ord(ch)) & 0xFFFFFFFF
        return (h ^ (h >> 16)) & (cap - 1)
    assert bucket("chr1") == 10 and bucket("chr2") == 11 and bucket("Aa") == bucket("BB") == 0
    keys = ["chr1", "chr2", "chr10", "scaffold_7", "chrUn"]
    got = pykcf.java_hashmap_order(keys)
    assert got == ["scaffold_7", "chrUn", "chr10", "chr1", "chr2"]  # buckets 1, 2, 8, 10, 11
    assert [bucket(k) for k in got] == sorted(bucket(k) for k in keys)
    # 13 keys force the 16 -> 32 resize: the order follows the 32-bucket table
    many = [f"chr{i}" for i in range(1, 14)]
    got = pykcf.java_hashmap_order(many)
    assert [bucket(k, 32) for k in got] == sorted(bucket(k, 32) for k in many)
    # equal buckets keep insertion order: "Aa" and "BB" collide (hashCode 2112 both)
    assert pykcf.java_hashmap_order(["BB", "Aa"]) == ["BB", "Aa"] and pykcf.java_hashmap_order(["Aa", "BB"]) == ["Aa", "BB"]


def test_java_round_and_kd_round_trip():
    assert [pykcf.java_round(x) for x in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 0.49999999999999994)] == [1, 2, 3, 0, -1, 2, 0]
    # KD is written with two decimals and read back as round(KD * OB) (Window.java:70): lossy, and the reference lives with it
    cell = pykcf.parse_cell("N:3:7:10:2:1:8.57:91.23", 100, 130, W)
    assert cell.obs == 7 and round(cell.mean_kmer_count * 7) == 60 and pyhost.java_format_2f(cell.mean_kmer_count) == "8.57"
    cell = pykcf.parse_cell("4:0:3:0:0:0:0.10:50.00", 100, 130, W)
    assert cell.ibs == 4 and cell.mean_kmer_count == 0.0  # round(0.1 * 3) = 0 -> "kmerCount > 0" fails -> 0.00


def test_header_parse_and_text_round_trip():
    t, _ = sample_kcf("s1", 1, 0.5, step=2500)
    hdr, rows = pykcf.parse_kcf(t)
    assert hdr.window_size == 5000 and hdr.step_size == 2500 and hdr.kmer_size == 31 and not hdr.is_ibs
    assert hdr.window_count == len(rows) == 34 and hdr.weights == W and hdr.samples == ["s1"]
    assert list(hdr.contigs.items()) == CHROMS and hdr.cmds == ["kcftools getVariations -s s1"]
    assert pykcf.kcf_text(hdr, rows, "2026-01-01") == t  # a one-sample file survives the reader + writer unchanged
    other = pykcf.KcfHeader.parse(hdr.text("x").replace("value=31", "value=21"))
    assert hdr.mismatch(other) == "Kmer size mismatch between the KCFs" and hdr.mismatch(hdr) is None


def test_find_ibs_block_numbering_known_answer():
    """hand-walked FindIBS.java:124-158: chromosomes visited in HashMap order (here b < a is impossible: 'a' = bucket 1,
    'b' = bucket 2), min = 1: a gap of more than one non-IBS window or a chromosome change opens a new block"""
    hdr = pykcf.KcfHeader(reference="r", contigs={"b": 100, "a": 100}, samples=["s"])
    hdr.params = [(k, v) for k, v in zip(pykcf.PARAM_KEYS, ["10", "0", "31", "false", "12", "0.3", "0.3", "0.4"])]

    def row(c, i, ok):
        obs = 100 if ok else 10  # score 100.0 or 10*0.4 + ... < 95
        r = dict(total_kmers=100, eff_len=130, obs=obs, variations=0 if ok else 1, inner=0, left=0, right=100 - obs, kmer_count_sum=obs * 5)
        return pyhost.kcf_row(c, 10 * i, 10 * i + 10, f"{c}{i}", r, W)
    pattern = {"b": [0, 1, 1, 0, 1, 0, 0, 1], "a": [1, 0, 1, 0]}
    text = hdr.text("d") + "".join(row(c, i, ok) + "\n" for c in ("b", "a") for i, ok in enumerate(pattern[c]))
    out, summ, beds = pykcf.find_ibs(text, "cmd", "d", min_consecutive=1, summary=True, bed=True)
    h2, rows = pykcf.parse_kcf(out)
    assert h2.is_ibs and [r.seq for r in rows] == ["a"] * 4 + ["b"] * 8  # HashMap order, not file order
    got = {c: [r.data["s"].ibs for r in rows if r.seq == c] for c in "ab"}
    assert got["a"] == [1, -1, 1, -1]                       # one non-IBS window in between: same block
    assert got["b"] == [-1, 2, 2, -1, 2, -1, -1, 3]          # chromosome change -> 2; two non-IBS windows -> 3
    assert summ.split("\n")[1].split("\t")[:8] == ["1", "s", "a", "0", "30", "30", "3", "2"]
    assert beds["s"] == "a\t0\t30\nb\t10\t50\nb\t70\t80\n"
    # --var flips the test
    out2, _, _ = pykcf.find_ibs(text, "cmd", "d", detect_var=True, min_consecutive=1)
    rows2 = pykcf.parse_kcf(out2)[1]
    assert [r.data["s"].ibs for r in rows2 if r.seq == "a"] == [-1, 1, -1, 1]


def test_kcf2gt_known_answer(cohort_files):
    merged = pykcf.cohort(cohort_files["texts"][:2], ["a", "b"], "cmd", "d")
    table, cmap = pykcf.kcf2gt(merged)
    lines = table.split("\n")
    assert lines[0] == "# Genotype Table 0:95.0 - 100.00, 2:60.0 - 95.0, 1:30.0 - 60.0, -1: <=30.0"
    assert lines[1] == "ID\tCHR\tSTART\tEND\talpha\tbeta"
    assert cmap == "contigName\tcontigID\nchr1\t1\nchr2\t2\nchr10\t3\nscaffold_7\t4\nchrUn\t5\n"
    hdr, rows = pykcf.parse_kcf(merged)
    assert len(lines) - 3 == len(rows)  # no filter active: every window is written
    for l, r in zip(lines[2:], rows):
        f = l.split("\t")
        assert f[0] == r.wid and f[2:4] == [str(r.start), str(r.end)]
        for a, s in zip(f[4:], hdr.samples):
            sc = r.data[s].score
            assert int(a) == (0 if sc >= 95 else 2 if sc >= 60 else -1 if sc <= 30 else 1)
    # with a filter the monomorphic windows go
    t2, _ = pykcf.kcf2gt(merged, min_maf=0.01)
    assert 2 < len(t2.split("\n")) < len(lines)
    with pytest.raises(pykcf.KcfError):
        pykcf.kcf2gt(merged, score_a=50.0, score_b=60.0)


def test_cohort_restatement(cohort_files):
    texts = cohort_files["texts"]
    merged = pykcf.cohort(texts, ["a", "b", "c", "d"], "kcftools cohort", "2026-01-01")
    hdr, rows = pykcf.parse_kcf(merged)
    assert hdr.samples == ["alpha", "beta", "gamma", "delta"] and len(hdr.cmds) == 5 and hdr.cmds[-1] == "kcftools cohort"
    singles = [pykcf.parse_kcf(t)[1] for t in texts]
    for i, r in enumerate(rows):
        assert [r.data[s].text() for s in hdr.samples] == [singles[j][i].data[n].text() for j, n in enumerate(hdr.samples)]
        scores = [r.data[s].score for s in hdr.samples]
        info = dict(kv.split("=") for kv in r.text().split("\t")[5].split(";"))
        assert info["IS"] == pyhost.java_format_2f(min(scores)) and info["XS"] == pyhost.java_format_2f(max(max(scores), 1.401298464324817e-45))
    bad = texts[1].replace("<ID=kmer,value=31>", "<ID=kmer,value=25>")
    with pytest.raises(pykcf.KcfError, match="Kmer size mismatch"):
        pykcf.cohort([texts[0], bad], ["a", "b"], "c", "d")
    with pytest.raises(pykcf.KcfError, match="already exists"):
        pykcf.cohort([texts[0], texts[0]], ["a", "b"], "c", "d")


def test_cli_usage_of_the_new_commands(cli):
    for cmd, first in (("cohort", "Usage: kcftools cohort"), ("findIBS", "Usage: kcftools findIBS"), ("kcf2gt", "Usage: kcftools kcf2gt")):
        p = run(cli, cmd, "--help")
        assert p.stdout.startswith(first)
        p = run(cli, cmd, "--nonsense", check=False)
        assert p.returncode == 2 and "Unknown option: '--nonsense'" in p.stderr and first in p.stderr
    p = run(cli, "kcf2gt", "-i", "x.kcf", check=False)
    assert p.returncode == 2 and "Missing required options: '--output'" in p.stderr


def test_cli_header_hook(cli, cohort_files):
    """the C++ KCF header parser / writer on CPU (no device needed for this hook)"""
    out = run(cli, "_kcfheader", cohort_files["paths"][0]).stdout
    out = "".join(l + "\n" for l in out.split("\n")[:-1] if l.startswith("#"))  # Logger lines share stdout
    hdr, _ = pykcf.parse_kcf(cohort_files["texts"][0])
    assert strip_volatile(out) == strip_volatile(hdr.text("x"))
    assert [l for l in out.split("\n") if l.startswith("##CMD=")] == ["##CMD=kcftools getVariations -s alpha"]


# ---------------------------------------------------------------------------------------------- GPU: the commands
@pytest.mark.gpu
def test_cli_cohort_matches_restatement(cli, cohort_files):
    d = cohort_files["dir"]
    out = str(d / "cohort.kcf")
    run(cli, "cohort", "-i", ",".join(cohort_files["paths"]), "-o", out)
    want = pykcf.cohort(cohort_files["texts"], cohort_files["paths"], "x", "x")
    assert strip_volatile(open(out).read()) == strip_volatile(want)
    lst = str(d / "list.txt")
    open(lst, "w").write("\n".join(cohort_files["paths"][:3]) + "\n")
    out2 = str(d / "cohort3.kcf")
    run(cli, "cohort", "-l", lst, "-o", out2)
    assert strip_volatile(open(out2).read()) == strip_volatile(pykcf.cohort(cohort_files["texts"][:3], cohort_files["paths"], "x", "x"))
    # the reference's fatal conditions
    bad = str(d / "bad.kcf")
    open(bad, "w").write(cohort_files["texts"][1].replace("<ID=window,value=5000>", "<ID=window,value=4000>"))
    p = run(cli, "cohort", "-i", cohort_files["paths"][0] + "," + bad, "-o", str(d / "x.kcf"), check=False)
    assert p.returncode == 1 and "Window size mismatch between the KCFs" in p.stdout + p.stderr
    p = run(cli, "cohort", "-o", str(d / "x.kcf"), check=False)
    assert p.returncode == 1 and "No input files provided" in p.stdout + p.stderr
    p = run(cli, "cohort", "-i", cohort_files["paths"][0], check=False)
    assert p.returncode == 2 and "Missing required options: '--output'" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(), dict(detect_var=True, min_consecutive=1, score_cutoff=80.5), dict(min_consecutive=0, score_cutoff=60.0)])
def test_cli_find_ibs_matches_restatement(cli, cohort_files, opts):
    d = cohort_files["dir"]
    merged = pykcf.cohort(cohort_files["texts"], cohort_files["paths"], "cohort-cmd", "x")
    inp = str(d / "merged.kcf")
    open(inp, "w").write(merged)
    out = str(d / "ibs_out")  # ".kcf" is appended
    args = ["findIBS", "-i", inp, "-o", out, "--summary", "--bed"]
    if opts.get("detect_var"):
        args.append("--var")
    if "min_consecutive" in opts:
        args += ["--min", str(opts["min_consecutive"])]
    if "score_cutoff" in opts:
        args += ["--score", str(opts["score_cutoff"])]
    run(cli, *args)
    want, summ, beds = pykcf.find_ibs(merged, "x", "x", summary=True, bed=True, **opts)
    got = open(out + ".kcf").read()
    assert strip_volatile(got) == strip_volatile(want)
    assert open(out + ".summary.tsv").read() == summ
    for s, b in beds.items():
        assert open(out + f".{s}.bed").read() == b
    labels = [f.split(":")[0] for l in got.split("\n") if l and not l.startswith("#") for f in l.split("\t")[7:]]
    assert "N" in labels and any(x not in ("N", "1") for x in labels)  # several blocks were numbered


@pytest.mark.gpu
def test_cli_find_ibs_sliding_windows_override_min(cli, cohort_files):
    d = cohort_files["dir"]
    t, _ = sample_kcf("slide", 77, 0.7, window=5000, step=1000)
    inp = str(d / "slide.kcf")
    open(inp, "w").write(t)
    p = run(cli, "findIBS", "-i", inp, "-o", str(d / "slide_ibs.kcf"), "--min", "2")
    assert "--min = windowSize/stepSize [5]" in p.stdout + p.stderr
    want, _, _ = pykcf.find_ibs(t, "x", "x", min_consecutive=2)
    assert strip_volatile(open(str(d / "slide_ibs.kcf")).read()) == strip_volatile(want)


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(), dict(min_maf=0.3), dict(max_missing=0.5, score_a=90.0, score_b=50.0, score_n=20.0), dict(chrs=True)])
def test_cli_kcf2gt_matches_restatement(cli, cohort_files, opts):
    d = cohort_files["dir"]
    merged = pykcf.cohort(cohort_files["texts"], cohort_files["paths"], "cohort-cmd", "x")
    inp = str(d / "merged_gt.kcf")
    open(inp, "w").write(merged)
    out = str(d / "gt.tsv")
    args = ["kcf2gt", "-i", inp, "-o", out]
    kw = {}
    for k, flag in (("score_a", "--score_a"), ("score_b", "--score_b"), ("score_n", "--score_n"), ("min_maf", "--maf"), ("max_missing", "--max-missing")):
        if k in opts:
            args += [flag, str(opts[k])]
            kw[k] = opts[k]
    if opts.get("chrs"):
        cf = str(d / "chrs.txt")
        open(cf, "w").write("# keep\nchr2\n\n  chrUn  \n")
        args += ["--chrs", cf]
        kw["chrs"] = {"chr2", "chrUn"}
    run(cli, *args)
    table, cmap = pykcf.kcf2gt(merged, **kw)
    assert open(out).read() == table and open(out + ".contigsMap.tsv").read() == cmap
    p = run(cli, "kcf2gt", "-i", inp, "-o", out, "--score_a", "50", "--score_b", "60", check=False)
    assert p.returncode == 1 and "Score A must be greater than Score B" in p.stdout + p.stderr


@pytest.mark.gpu
def test_cohort_fed_from_device_results_equals_the_file_pipeline(cli, tmp_path):
    """row f1: `getVariations -k a,b,c` fills the cohort matrix device-to-device and writes what
    getVariations x 3 -> cohort writes (rows byte for byte; header up to ##CMD / ##date)."""
    from tools import synth
    lens = (40_000, 12_345)
    recs, codes = [], []
    for i, n in enumerate(lens):
        g = synth.random_genome(n, 900 + i)
        codes.append(g)
        recs.append((f"chr{i + 1}", synth.fasta_record(g, f"chr{i + 1}", line=60, lower=synth.random_intervals(n, 3, 5, 400, 920 + i),
                                                        n_runs=synth.random_intervals(n, 2, 1, 300, 910 + i)), n, 60))
    fa = str(tmp_path / "ref.fa")
    synth.fasta_image(recs).write(fa)
    prefixes, names, singles = [], [], []
    for j, snp in enumerate((0.002, 0.02, 0.08)):
        qs = [synth.mutate(g, 950 + 10 * j + i, snp=snp, big_deletions=1 if i == 0 else 0, big_len=1500, replace_len=0) for i, g in enumerate(codes)]
        pre = str(tmp_path / f"db{j}")
        synth.kmc_image_from_genomes(qs, k=31, P=7, L=9, n_bins=32, counter_size=1, coverage=8.0, seed=60 + j).write(pre)
        prefixes.append(pre)
        names.append(f"s{j}")
        one = str(tmp_path / f"s{j}.kcf")
        run(cli, "getVariations", "-r", fa, "-k", pre, "-o", one, "-s", f"s{j}", "-f", "window", "-w", "2500")
        singles.append(one)
    merged = str(tmp_path / "merged.kcf")
    run(cli, "cohort", "-i", ",".join(singles), "-o", merged)
    direct = str(tmp_path / "direct.kcf")
    run(cli, "getVariations", "-r", fa, "-k", ",".join(prefixes), "-o", direct, "-s", ",".join(names), "-f", "window", "-w", "2500")
    assert strip_volatile(open(direct).read()) == strip_volatile(open(merged).read())
    # the same with the databases shared out over several devices (here two contexts on the one GPU: the thread-per-device
    # path of configs[4], samples sharded over the GPUs of a box)
    spread = str(tmp_path / "spread.kcf")
    p = run(cli, "getVariations", "-r", fa, "-k", ",".join(prefixes), "-o", spread, "-s", ",".join(names), "-f", "window", "-w", "2500", "--devices", "0,0")
    assert "Sample s1 screened on device 0" in p.stdout and strip_volatile(open(spread).read()) == strip_volatile(open(merged).read())
    rows = [l for l in open(direct).read().split("\n") if l and not l.startswith("#")]
    assert len(rows) == 22 and all(len(r.split("\t")) == 10 for r in rows)
    sc = np.array([[float(f.split(":")[7]) for f in r.split("\t")[7:]] for r in rows])
    assert sc[:, 0].mean() > sc[:, 1].mean() > sc[:, 2].mean() > 0  # more divergence, lower identity
    # and the same matrix through ctypes: Cohort.add_plan + scores reproduce the library's own per-window scores
    from kcftools_b200.api import KMC, Cohort, Context, fixed_windows
    with Context(0) as ctx:
        img = synth.fasta_image(recs)
        for i in range(len(lens)):
            ctx.ref_add(img.seq_bytes(i), img.line_bases[i], img.line_width[i], img.lengths[i])
        wins, segs, *_ = fixed_windows(list(lens), 2500, 0, 31)
        plan = ctx.plan(31, wins, segs)
        co = Cohort(ctx, wins.size, len(prefixes))
        per = []
        for j, pre in enumerate(prefixes):
            db = KMC(ctx, prefix=pre)
            plan.run(db)
            per.append(plan.fetch().copy())
            co.add_plan(j, plan)
            db.close()
        co.scores(W)
        for j in range(len(prefixes)):
            cells, tot, eff = co.fetch(j)
            assert (tot == per[j]["total_kmers"]).all() and (eff == per[j]["eff_len"]).all()
            for f in ("obs", "variations", "inner", "left", "right"):
                assert (cells[f] == per[j][f]).all()
            assert (cells["kmer_count"] == per[j]["kmer_count_sum"]).all() and (cells["ibs"] == -1).all()
            assert (cells["score"] == per[j]["score"]).all()  # bit-identical: same formula, same rounding
        al, bad = co.genotypes()
        assert al.shape == (wins.size, 3) and set(np.unique(al)) <= {-1, 0, 1, 2}
        co.close()
        plan.close()
