"""The C++ host side (kcftools_b200/host): .faidx generation, window / GTF logic, option validation and KCF text,
against the Python restatement of the reference's host code (oracle/pyhost.py).  The getVariations run itself needs
a GPU (marked gpu); everything else runs on CPU through the CLI's test hooks."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import pyhost
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "kcftools_b200", "host")
CLI = os.path.join(HOST_DIR, "kcftools_b200")


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


@pytest.fixture(scope="module")
def genome(tmp_path_factory):
    """three sequences (different line widths are not allowed inside one sequence, but differ between sequences via
    separate records), N runs, lower case; written to disk with its GTF"""
    d = tmp_path_factory.mktemp("host")
    lens = (40_000, 12_345, 80)
    recs, codes = [], []
    for i, n in enumerate(lens):
        g = synth.random_genome(n, 700 + i)
        codes.append(g)
        nr = synth.random_intervals(n, 2, 1, max(2, min(300, n // 20)), 710 + i)
        low = synth.random_intervals(n, 3, 5, max(6, n // 40), 720 + i)
        name = f"chr{i + 1} some description"
        recs.append((name, synth.fasta_record(g, name, line=60 if i != 1 else 71, lower=low, n_runs=nr), n, 60 if i != 1 else 71))
    img = synth.fasta_image(recs)
    fa = str(d / "ref.fa")
    img.write(fa)
    gtf_text = synth.synthetic_gtf([("chr1", lens[0]), ("chr2", lens[1])], 6, 99, max_tx=3, max_exons=6, exon_lo=40, exon_hi=900)
    # extra shapes the reference handles: a transcript line before its gene line, same-start exons on both strands,
    # duplicated exon, comment and blank lines, an unknown feature type
    gtf_text = ("# comment\n\n"
                'chr2\tx\ttranscript\t100\t900\t.\t+\t.\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\texon\t100\t300\t.\t+\t.\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\texon\t100\t250\t.\t-\t.\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\texon\t100\t300\t.\t+\t.\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\texon\t301\t500\t.\t+\t.\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\tCDS\t120\t280\t.\t+\t0\tgene_id "GX"; transcript_id "GX.t1";\n'
                'chr2\tx\tgene\t100\t900\t.\t+\t.\tgene_id "GX";\n') + gtf_text
    gtf = str(d / "ann.gtf")
    open(gtf, "w").write(gtf_text)
    return {"dir": d, "fa": fa, "gtf": gtf, "gtf_text": gtf_text, "img": img, "lens": lens, "codes": codes,
            "names": ["chr1", "chr2", "chr3"]}


def test_java_number_formatting(cli):
    rng = np.random.default_rng(1)
    vals = [0.0, 0.125, 2.675, 0.005, 0.015, 99.995, 99.13053412047228, 1e7, 12345678.0, 0.001, 1.0, 100.0, 7.0, 255.0, 3.4028234663852886e38,
            1.401298464324817e-45, 33.335, 49000 / 49970 * 40] + list(rng.random(200) * 100) + [float(x) for x in rng.integers(0, 60000, 50)]
    hexes = [struct.pack(">d", v).hex() for v in vals]
    out = run(cli, "_format", *hexes).stdout.strip().split("\n")
    assert len(out) == len(vals)
    for v, line in zip(vals, out):
        f2, dstr, fstr = line.split("\t")
        assert f2 == pyhost.java_format_2f(v), (v, f2)
        assert fstr == pyhost.java_float_str(v), (v, fstr)
    assert out[1].split("\t")[0] == "0.13" and out[2].split("\t")[0] == "2.68"      # HALF_UP on the shortest repr (Q13)
    assert out[0].split("\t")[1] == "0.0" and out[9].split("\t")[1] == "0.001" and out[7].split("\t")[1] == "1.0E7"


def test_faidx_matches_reference_arithmetic(cli, genome):
    fai = genome["fa"] + ".faidx"
    if os.path.exists(fai):
        os.unlink(fai)
    out = run(cli, "_faidx", genome["fa"]).stdout
    rows = [tuple(l.split("\t")) for l in out.strip().split("\n") if "\t" in l and " - " not in l]
    want = pyhost.faidx_generate(open(genome["fa"]).read())
    assert [(n, int(a), int(b), int(c), int(d)) for (n, a, b, c, d) in rows] == want
    assert [r[0] for r in rows] == genome["names"]  # header name = text up to the first space (FastaIndex.java:268)
    assert open(fai).read() == "".join("\t".join(map(str, r)) + "\n" for r in want)
    assert "Generating/Updating index file" in out
    assert "Using existing index file" in run(cli, "_faidx", genome["fa"]).stdout
    img = genome["img"]
    assert [w[2] for w in want] == list(img.offsets) and [w[1] for w in want] == list(img.lengths)


def _parse_windows(stdout):
    rows = []
    for l in stdout.split("\n"):
        if not l.startswith("W\t"):
            continue
        f = l.split("\t")
        segs = [tuple(int(x) for x in s.split(":")) for s in f[6:]]
        rows.append((f[1], f[2], int(f[3]), int(f[4]), None if f[5] == "1" else segs))
    return rows


@pytest.mark.parametrize("window,step,k", [(5000, 0, 31), (1000, 400, 21), (700, 2500, 31), (64, 0, 32)])
def test_fixed_windows_match_reference_loops(cli, genome, window, step, k):
    got = _parse_windows(run(cli, "_windows", "-r", genome["fa"], "-f", "window", "-w", str(window), "-p", str(step), "--kmer-size", str(k)).stdout)
    want = pyhost.windows_of("window", genome["names"], list(genome["lens"]), k, window, step)
    assert got == want and len(got) > 0


@pytest.mark.parametrize("feature", ["gene", "transcript"])
def test_gtf_windows_match_reference_rules(cli, genome, feature):
    got = _parse_windows(run(cli, "_windows", "-r", genome["fa"], "-f", feature, "-g", genome["gtf"], "--kmer-size", "31").stdout)
    gtf = pyhost.Gtf(genome["gtf_text"])
    want = pyhost.windows_of(feature, genome["names"], list(genome["lens"]), 31, gtf=gtf)
    assert got == want and len(got) >= 12
    gx = [w for w in got if w[0] in ("GX", "GX.t1")][0]
    # duplicate exon dropped; + and - exons with the same start stay separate (merge needs the same strand); abutting
    # + exons 100-300 / 301-500 do NOT merge (301 <= 300 is false)
    assert sorted(gx[4]) == sorted([(1, 99, 201), (1, 99, 151), (1, 300, 200)])
    assert gx[2:4] == (100, 900)  # raw GTF coordinates, 1-based inclusive (Q11)


def test_validation_messages_and_exit_codes(cli, genome):
    base = ["getVariations", "-r", genome["fa"], "-k", "nodb", "-o", "/dev/null", "-s", "S"]
    cases = [
        (["-f", "window"], "Window size is required for window model"),
        (["-f", "window", "-w", "100", "-g", genome["gtf"]], "GTF file is not valid for window model"),
        (["-f", "gene"], "GTF file is required for targeted model"),
        (["-f", "gene", "-g", genome["gtf"], "-w", "5"], "Window size is not valid for targeted model"),
        (["-f", "exon", "-w", "5"], "Invalid model type: exon. Supported models are 'window' or 'gene' or 'transcript'"),
        (["-f", "window", "-w", "100", "-t", "0"], "Number of threads should be greater than 0"),
        (["-f", "window", "-w", "100", "-c", "0"], "Minimum kmer count should be at least 1"),
    ]
    for extra, msg in cases:
        p = run(cli, *base, *extra, check=False)
        assert p.returncode == 1 and msg in p.stdout and " - ERROR    - GetVariants" in p.stdout, (extra, p.stdout, p.stderr)
        assert "CMD options - GetVariants" in p.stdout and "--min-k-count" in p.stdout
    p = run(cli, "getVariations", "-r", genome["fa"], check=False)
    assert p.returncode == 2 and "Missing required options" in p.stderr and "Usage:" in p.stderr
    p = run(cli, *base, "-f", "window", "-w", "x", check=False)
    assert p.returncode == 2 and "is not an int" in p.stderr
    p = run(cli, *base, "-f", "window", "--bogus", check=False)
    assert p.returncode == 2 and "Unknown option" in p.stderr
    assert run(cli, "getVariations", "--help").stdout.startswith("Usage: kcftools getVariations")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["window", "sliding", "gene", "transcript"])
def test_getvariations_cli_writes_the_reference_kcf(cli, genome, mode):
    """end to end through the CLI: files on disk -> KCF text, compared row by row with the CPU oracle's integers pushed
    through the reference's text formatting (keyed by window id: row order among equal starts is free, Q12)."""
    from oracle import binding as ob
    from common import windows_from_lists
    d = genome["dir"]
    qs = [synth.mutate(g, 800 + i, big_deletions=1 if i == 0 else 0, big_len=1500, replace_len=0) for i, g in enumerate(genome["codes"])]
    kmc = synth.kmc_image_from_genomes(qs, k=31, P=7, L=9, n_bins=32, counter_size=1, coverage=8.0, seed=5)
    prefix = str(d / "sample")
    kmc.write(prefix)
    out = str(d / f"out_{mode}.kcf")
    w = (0.25, 0.35, 0.4)
    args = ["getVariations", "-r", genome["fa"], "-k", prefix, "-o", out, "-s", "my:sample", "--wi", str(w[0]), "--wt", str(w[1]), "--wr", str(w[2]),
            "-c", "2", "-t", "4", "-m"]
    if mode == "window":
        args += ["-f", "window", "-w", "5000"]
        want_w = pyhost.windows_of("window", genome["names"], list(genome["lens"]), 31, 5000, 0)
    elif mode == "sliding":
        args += ["-f", "window", "-w", "3000", "-p", "1000"]
        want_w = pyhost.windows_of("window", genome["names"], list(genome["lens"]), 31, 3000, 1000)
    else:
        args += ["-f", mode, "-g", genome["gtf"]]
        want_w = pyhost.windows_of(mode, genome["names"], list(genome["lens"]), 31, gtf=pyhost.Gtf(genome["gtf_text"]))
    p = run(cli, *args)
    assert "Sample name contains invalid characters, changed to: my_sample" in p.stdout
    wins, segs = windows_from_lists([ws[4] for ws in want_w])
    img = genome["img"]
    seqs = [(img.seq_bytes(i), img.line_bases[i], img.line_width[i], img.lengths[i]) for i in range(3)]
    rc, res = ob.OracleKMC(kmc.pre, kmc.suf).screen(seqs, wins, segs, min_count=2, w=w, threads=4)
    assert rc == 0
    lines = open(out).read().split("\n")
    assert lines[-1] == ""
    header = [l for l in lines if l.startswith("#")]
    rows = [l for l in lines if l and not l.startswith("#")]
    assert header[0] == "##format=KCF0.4.0" and header[2] == "##source=kcftools" and header[3] == "##reference=" + genome["fa"]
    assert header[4:7] == [f"##contig=<ID=chr{i + 1},length={n}>" for i, n in enumerate(genome["lens"])]
    params = [l for l in header if l.startswith("##PARAM")]
    wsize = {"window": 5000, "sliding": 3000}.get(mode, 0)
    step = 1000 if mode == "sliding" else 0
    assert params == [f"##PARAM=<ID=window,value={wsize}>", f"##PARAM=<ID=step,value={step}>", "##PARAM=<ID=kmer,value=31>",
                      "##PARAM=<ID=IBS,value=false>", f"##PARAM=<ID=nwindow,value={len(want_w)}>", "##PARAM=<ID=wti,value=0.25>",
                      "##PARAM=<ID=wtt,value=0.35>", "##PARAM=<ID=wtk,value=0.4>"]
    assert len([l for l in header if l.startswith("##INFO=")]) == 10 and len([l for l in header if l.startswith("##FORMAT=")]) == 8
    assert header[-1] == "#CHROM\tSTART\tEND\tID\tTOTAL_KMERS\tINFO\tFORMAT\tmy_sample"
    want_rows = {ws[0]: pyhost.kcf_row(ws[1], ws[2], ws[3], ws[0], res[i], w) for i, ws in enumerate(want_w)}
    assert len(rows) == len(want_rows)
    for r in rows:
        wid = r.split("\t")[3]
        assert r == want_rows[wid]
    # rows of a contig are sorted by start; contigs follow the .faidx order
    starts = [(r.split("\t")[0], int(r.split("\t")[1])) for r in rows]
    order = {n: i for i, n in enumerate(genome["names"])}
    assert starts == sorted(starts, key=lambda t: (order[t[0]], t[1]))
    assert any(":0.00" not in r for r in rows) and res["obs"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["window", "gene"])
def test_getvariations_one_database_over_several_devices(cli, genome, mode):
    """--devices with ONE database: the windows are cut over the devices (here three contexts on device 0; on the box one per
    GPU) and the KCF must be the single-device file byte for byte, ##date / ##CMD aside (GetVariants.java:129-151, 169-179)"""
    d = genome["dir"]
    qs = [synth.mutate(g, 900 + i, big_deletions=0, replace_len=0) for i, g in enumerate(genome["codes"])]
    kmc = synth.kmc_image_from_genomes(qs, k=31, P=7, L=9, n_bins=32, counter_size=1, coverage=8.0, seed=6)
    prefix = str(d / "sample_md")
    kmc.write(prefix)
    common = ["getVariations", "-r", genome["fa"], "-k", prefix, "-s", "S"]
    common += ["-f", "window", "-w", "4000"] if mode == "window" else ["-f", "gene", "-g", genome["gtf"]]
    one, three = str(d / f"md1_{mode}.kcf"), str(d / f"md3_{mode}.kcf")
    run(cli, *common, "-o", one, "--device", "0")
    p = run(cli, *common, "-o", three, "--devices", "0,0,0")
    assert "on 3 devices" in p.stdout

    def body(path):
        return [l for l in open(path).read().split("\n") if not l.startswith("##date") and not l.startswith("##CMD")]
    assert body(one) == body(three) and len(body(one)) > 30


def test_java_string_hash_counts_utf16_units(cli):
    """String.hashCode() runs over UTF-16 code units: known answers worked out by hand (31 * h + unit), ASCII, Latin-1, BMP and
    a supplementary character (U+1D538 = D835 DD38), and the C++ host against the Python restatement"""
    names = ["chr1", "é", "chrÄ1", "染色体1", "\U0001d538x"]
    want = [pyhost.java_string_hash(s) for s in names]
    assert want[0] == ((((ord("c") * 31 + ord("h")) * 31 + ord("r")) * 31 + ord("1")) & 0xFFFFFFFF)
    assert want[1] == 233 and want[4] == ((31 * 0xD835 + 0xDD38) * 31 + ord("x"))
    got = [int(x) for x in run(cli, "_hash", *names).stdout.split()]
    assert got == want
