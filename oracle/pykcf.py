"""CPU restatement of the reference's KCF reader and of the three plugins that consume getVariations output —
test infrastructure only (like everything under oracle/: nothing under kcftools_b200/ may import this).
PARITY UNPINNED: the reference ships no tests or fixtures for these paths and cannot run here (no JVM); what pins this
file are hand-derived known answers (tests/test_kcf_tools.py) and the reference's documented examples.

  KcfHeader            Data/KCFHeader.java:44-96 (parse), :291-330 (toString), :333-370 (equals), :420-432 (mergeHeader)
  parse_kcf, Row       Data/KCFReader.java:31-105; Data/Window.java:42-83 (row -> Window / Data; KD -> kmerCount via Math.round)
  row_text             Data/Window.java:125-152, 170-214 (INFO statistics); Data/Data.java:120-132
  cohort               Plugins/Cohort.java:71-119
  find_ibs             Plugins/FindIBS.java:74-273 (block numbering :124-158, summary :172-272, bed :222-233)
  kcf2gt               Plugins/KCFToGenotypeTable.java:66-197
  java_hashmap_order   java.util.HashMap iteration order of String keys (FindIBS iterates HashMaps of chromosome names)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .pyhost import compute_score, java_float_str, java_format_2f, java_lines, java_split, java_string_hash

KCF_VERSION = "0.4.0"
INFO_LINES = [
    '<ID=EFFLEN,Type=Integer,Description="Effective length of the window">',
    '<ID=IS,Type=Float,Description="Minimum score for the window">',
    '<ID=XS,Type=Float,Description="Maximum score for the window">',
    '<ID=MS,Type=Float,Description="Mean score for the window">',
    '<ID=IO,Type=Integer,Description="Minimum observed kmers in the window">',
    '<ID=XO,Type=Integer,Description="Maximum observed kmers in the window">',
    '<ID=MO,Type=Integer,Description="Mean observed kmers in the window">',
    '<ID=IV,Type=Integer,Description="Minimum variations in the window">',
    '<ID=XV,Type=Integer,Description="Maximum variations in the window">',
    '<ID=MV,Type=Integer,Description="Mean variations in the window">',
]
FORMAT_LINES = [
    '<ID=IB,Type=Integer,Description="IBS number">',
    '<ID=VA,Type=Integer,Description="Variations">',
    '<ID=OB,Type=Integer,Description="Observed kmers">',
    '<ID=ID,Type=Integer,Description="Inner Distance">',
    '<ID=LD,Type=Integer,Description="Kmer Variation Distance at the leftTail">',
    '<ID=RD,Type=Integer,Description="Kmer Variation Distance at the rightTail">',
    '<ID=KD,Type=Float,Description="Mean Kmer Depth">',
    '<ID=SC,Type=Float,Description="Score">',
]
PARAM_KEYS = ["window", "step", "kmer", "IBS", "nwindow", "wti", "wtt", "wtk"]  # KCFHeader.java:26, 63-90


class KcfError(Exception):
    """Logger.error in the reference (print + System.exit(1))"""


def java_double_str(x: float) -> str:
    """Double.toString for the magnitudes these tools print (1e-3 <= |x| < 1e7 or 0)"""
    if x == 0:
        return "0.0"
    r = repr(float(x))
    if "e" in r or "E" in r:
        raise NotImplementedError(r)
    return r if "." in r else r + ".0"


def java_round(x: float) -> int:
    """Math.round(double): the closest long, ties towards positive infinity (exact: not floor(x + 0.5) in floating point)"""
    f = math.floor(x)
    return int(f) + (1 if (x - f) >= 0.5 else 0)


def java_hashmap_order(keys: list[str]) -> list[str]:
    """iteration order of a java.util.HashMap<String, ?> after inserting `keys` (distinct) in order: buckets of the final
    table ascending, insertion order inside a bucket (resizes split a bucket order-preservingly; no bin here reaches the
    treeify threshold)"""
    cap, n = 16, 0
    for _ in keys:
        n += 1
        if n > cap * 3 // 4:
            cap *= 2

    def spread(s):
        h = java_string_hash(s) & 0xFFFFFFFF
        return (h ^ (h >> 16)) & (cap - 1)
    return [k for _, _, k in sorted((spread(k), i, k) for i, k in enumerate(keys))]


@dataclass
class KcfHeader:
    reference: str = ""
    contigs: dict | None = None          # name -> length, insertion ordered (LinkedHashMap)
    cmds: list | None = None
    samples: list | None = None
    params: list = field(default_factory=lambda: [None] * 8)  # (key, value-string) pairs in PARAM_KEYS order

    @staticmethod
    def parse(text: str) -> "KcfHeader":
        h = KcfHeader()
        for line in java_split(text, "\n"):
            if line.startswith("##reference="):
                h.reference = line[12:]
            elif line.startswith("##contig="):
                f = java_split(line[10:len(line) - 1], ",")
                if h.contigs is None:
                    h.contigs = {}
                h.contigs[f[0][3:]] = int(f[1][7:])
            elif line.startswith("##CMD="):
                h.cmds = (h.cmds or []) + [line[6:]]
            elif line.startswith("#CHROM"):
                h.samples = java_split(line, "\t")[7:]
            elif line.startswith("##PARAM="):
                f = java_split(line[9:len(line) - 1], ",")
                key, value = f[0][3:], f[1][6:]
                if key in PARAM_KEYS:
                    h.params[PARAM_KEYS.index(key)] = (key, value)
        return h

    def _p(self, i):
        return self.params[i][1] if self.params[i] is not None else None

    def _int(self, i):
        return int(self._p(i)) if self._p(i) is not None else 0

    def _dbl(self, i):
        return float(self._p(i)) if self._p(i) is not None else 0.0

    window_size = property(lambda s: s._int(0))
    step_size = property(lambda s: s._int(1))
    kmer_size = property(lambda s: s._int(2))
    is_ibs = property(lambda s: s._p(3) is not None and s._p(3).lower() == "true")  # Boolean.parseBoolean
    window_count = property(lambda s: s._int(4))
    weights = property(lambda s: (s._dbl(5), s._dbl(6), s._dbl(7)))  # KCFHeader.getWeights: wti, wtt, wtk

    def mismatch(self, o: "KcfHeader") -> str | None:
        """KCFHeader.equals: the first differing property (the reference logs it as a fatal error), else None"""
        checks = [("Window size", self.window_size, o.window_size), ("Kmer size", self.kmer_size, o.kmer_size),
                  ("IBS processing", self.is_ibs, o.is_ibs), ("Number of windows", self.window_count, o.window_count),
                  ("Weight Inner Distance", self.weights[0], o.weights[0]), ("Weight Tail Distance", self.weights[1], o.weights[1]),
                  ("Weight Kmer Ratio", self.weights[2], o.weights[2]), ("Step size", self.step_size, o.step_size)]
        for name, a, b in checks:
            if a != b:
                return f"{name} mismatch between the KCFs"
        return None

    def text(self, date: str) -> str:
        sb = [f"##format=KCF{KCF_VERSION}", f"##date={date}", "##source=kcftools", f"##reference={self.reference}"]
        for n, l in (self.contigs or {}).items():
            sb.append(f"##contig=<ID={n},length={l}>")
        sb += ["##INFO=" + l for l in INFO_LINES] + ["##FORMAT=" + l for l in FORMAT_LINES]
        for p in self.params:
            if p is not None:
                sb.append(f"##PARAM=<ID={p[0]},value={p[1]}>")
        for c in self.cmds or []:
            sb.append("##CMD=" + c)
        sb.append("\t".join(["#CHROM", "START", "END", "ID", "TOTAL_KMERS", "INFO", "FORMAT"] + list(self.samples or [])))
        return "\n".join(sb) + "\n"


@dataclass
class Cell:
    """Data.java: one sample in one window"""
    obs: int
    variations: int
    inner: int
    left: int
    right: int
    mean_kmer_count: float
    score: float
    ibs: int

    def text(self) -> str:  # Data.toString
        return ":".join(["N" if self.ibs == -1 else str(self.ibs), str(self.variations), str(self.obs), str(self.inner), str(self.left),
                         str(self.right), java_format_2f(self.mean_kmer_count), java_format_2f(self.score)])


@dataclass
class Row:
    """Window.java as read from a KCF line"""
    seq: str
    start: int
    end: int
    wid: str
    total: int
    eff: int
    data: dict  # sample -> Cell, insertion ordered

    def text(self) -> str:  # Window.toString with calculateStats (Window.java:125-214)
        f32 = np.float32
        cells = list(self.data.values())
        mn_o, mx_o, mn_v, mx_v = 2**31 - 1, -2**31, 2**31 - 1, -2**31
        mean_o, mean_v = f32(0), f32(0)
        mn_s, mx_s, mean_s = float(np.finfo(np.float32).max), float(f32(1.401298464324817e-45)), 0.0
        for d in cells:
            mn_o, mx_o = min(mn_o, d.obs), max(mx_o, d.obs)
            mean_o = f32(mean_o + f32(d.obs))          # float += int
            mn_v, mx_v = min(mn_v, d.variations), max(mx_v, d.variations)
            mean_v = f32(mean_v + f32(d.variations))
            mn_s = d.score if d.score < mn_s else mn_s
            mx_s = d.score if d.score > mx_s else mx_s
            mean_s += d.score
        n = len(cells)
        mean_o, mean_v, mean_s = f32(mean_o / f32(n)), f32(mean_v / f32(n)), mean_s / n
        info = (f"EFFLEN={self.eff};IS={java_format_2f(mn_s)};XS={java_format_2f(mx_s)};MS={java_format_2f(mean_s)};IO={mn_o};XO={mx_o};"
                f"MO={java_format_2f(float(mean_o))};IV={mn_v};XV={mx_v};MV={java_float_str(float(mean_v))}")
        return "\t".join([self.seq, str(self.start), str(self.end), self.wid, str(self.total), info, "GT:VA:OB:ID:LD:RD:KD:SC"]
                         + [d.text() for d in cells])


def parse_cell(fld: str, total: int, eff: int, w) -> Cell:  # Window.parseSampleData + Data ctor
    s = java_split(fld, ":")
    ibs = -1 if s[0] == "N" else int(s[0])
    var, obs, inner, left, right = (int(x) for x in s[1:6])
    kmer_count = java_round(float(s[6]) * obs)
    mean = (kmer_count / obs) if kmer_count > 0 else 0.0
    return Cell(obs, var, inner, left, right, mean, compute_score_w(obs, total, eff, inner, left, right, w), ibs)


def compute_score_w(obs, total, eff, inner, left, right, w) -> float:
    if obs == 0 or total == 0 or eff == 0:
        return 0.0
    if w[0] + w[1] + w[2] != 1.0:
        raise KcfError("Weights should sum to 1.0")
    return compute_score(obs, total, eff, inner, left, right, w)


def parse_kcf(text: str) -> tuple[KcfHeader, list[Row]]:
    lines = java_lines(text)
    i = 0
    while i < len(lines) and lines[i].startswith("##"):
        i += 1
    hdr = KcfHeader.parse("\n".join(lines[:i + 1]) + "\n")
    rows = []
    w = hdr.weights
    for line in lines[i + 1:]:
        f = java_split(line, "\t")
        info = dict(kv.split("=")[:2] for kv in java_split(f[5], ";"))
        total, eff = int(f[4]), int(info["EFFLEN"])
        data = {}
        for j in range(7, len(f)):
            data[hdr.samples[j - 7]] = parse_cell(f[j], total, eff, w)
        rows.append(Row(f[0], int(f[1]), int(f[2]), f[3], total, eff, data))
    return hdr, rows


def kcf_text(hdr: KcfHeader, rows: list[Row], date: str) -> str:
    return hdr.text(date) + "".join(r.text() + "\n" for r in rows)


# ------------------------------------------------------------------------------------------------ cohort
def cohort(texts: list[str], names: list[str], cmdline: str, date: str) -> str:
    header, windows = None, {}
    for i, t in enumerate(texts):
        hdr, rows = parse_kcf(t)
        if i == 0:
            header = hdr
            for r in rows:
                windows[r.wid] = r      # LinkedHashMap.put: a repeated id keeps its first position, takes the last row
        else:
            mm = header.mismatch(hdr)
            if mm:
                raise KcfError(mm)
            if hdr.samples is not None:
                header.samples = (header.samples or []) + list(hdr.samples)
            for c in hdr.cmds or []:
                header.cmds = (header.cmds or []) + [c]
            for r in rows:
                if r.wid not in windows:
                    raise KcfError(f"Windows mismatch found in sample: {names[i]}")
                for s, d in r.data.items():
                    if s in windows[r.wid].data:
                        raise KcfError(f"Sample {s} already exists in window {r.wid}")
                    windows[r.wid].data[s] = d
    header.cmds = (header.cmds or []) + [cmdline]
    for r in windows.values():  # alignSamplesWithHeader; a missing sample is a null Data -> NullPointerException on write
        if any(s not in r.data for s in header.samples):
            raise KcfError(f"window {r.wid} lacks a sample of the header")
        r.data = {s: r.data[s] for s in header.samples}
    return kcf_text(header, list(windows.values()), date)


# ------------------------------------------------------------------------------------------------ findIBS
def find_ibs(text: str, cmdline: str, date: str, detect_var: bool = False, min_consecutive: int = 4, score_cutoff: float = 95.0,
             summary: bool = False, bed: bool = False):
    """returns (kcf text, summary tsv text or None, {sample: bed text})"""
    hdr, rows = parse_kcf(text)
    cutoff = float(np.float32(score_cutoff))  # the option is a Java float; compared against the double score
    if hdr.step_size > 0:
        min_consecutive = hdr.window_size // hdr.step_size
    names = []
    by_chrom = {}
    for r in rows:
        if r.seq not in by_chrom:
            names.append(r.seq)
            by_chrom[r.seq] = []
        by_chrom[r.seq].append(r)
    order = java_hashmap_order(names)  # HashMap<String, Window[]>.keySet()
    for sample in hdr.samples:
        block_num, block_chrom, first_found = 0, None, False
        for chrom in order:
            num_na = 0
            for r in by_chrom[chrom]:
                if sample not in r.data:
                    continue
                sc = r.data[sample].score
                is_ibs = (sc < cutoff) if detect_var else (sc >= cutoff)
                if is_ibs:
                    if not first_found:
                        block_num, first_found = 1, True
                    elif num_na > min_consecutive or (block_chrom is not None and block_chrom != chrom):
                        block_num += 1
                    block_chrom = chrom
                    r.data[sample].ibs = block_num
                    num_na = 0
                else:
                    num_na += 1
                    r.data[sample].ibs = -1
    hdr.params[3] = ("IBS", "true")
    hdr.cmds = (hdr.cmds or []) + [cmdline]
    out = hdr.text(date) + "".join(r.text() + "\n" for chrom in order for r in by_chrom[chrom])
    summ, beds = None, {}
    if summary:
        sb = ["Block\tSample\tChromosome\tStart\tEnd\tLength\tTotalBlocks\tIBSBlocks\tIBSProportion\tMeanScore\n"]
        for sample in hdr.samples:
            blocks = {}
            for chrom in order:
                na = []
                for r in by_chrom[chrom]:
                    v = r.data[sample].ibs
                    if v == -1:
                        na.append(r)
                    elif v in blocks:
                        blocks[v].extend(na)
                        blocks[v].append(r)
                        na = []
                    else:
                        blocks[v] = [r]
                        na = []
            if bed:
                beds[sample] = "".join(f"{b[0].seq}\t{b[0].start}\t{b[-1].end}\n" for b in blocks.values() if b)
            for bn, b in blocks.items():
                if not b:
                    continue
                mean = np.float32(0)
                ibs_blocks = 0
                for r in b:
                    mean = np.float32(np.float64(mean) + r.data[sample].score)  # float += double
                    if r.data[sample].ibs != -1:
                        ibs_blocks += 1
                mean = np.float32(mean / np.float32(len(b)))
                prop = np.float32(np.float32(ibs_blocks) / np.float32(len(b)))
                sb.append(f"{bn}\t{sample}\t{b[0].seq}\t{b[0].start}\t{b[-1].end}\t{b[-1].end - b[0].start}\t{len(b)}\t{ibs_blocks}\t"
                          f"{java_format_2f(float(prop))}\t{java_format_2f(float(mean))}\n")
        summ = "".join(sb)
    return out, summ, beds


# ------------------------------------------------------------------------------------------------ kcf2gt
def kcf2gt(text: str, score_a: float = 95.0, score_b: float = 60.0, score_n: float = 30.0, min_maf: float = 0.0, max_missing: float = 1.0,
           chrs: set | None = None):
    """returns (genotype table text, contigs map text)"""
    for nm, v in (("A", score_a), ("B", score_b), ("N", score_n)):
        if v < 0.0 or v > 100.0:
            raise KcfError(f"Score {nm} must be between 0.0 and 100.0")
    if score_a <= score_b:
        raise KcfError("Score A must be greater than Score B")
    if score_b == 0.0 and score_n != 0.0:
        score_n = 0.0
    hdr, rows = parse_kcf(text)
    samples = hdr.samples
    contig_names = list((hdr.contigs or {}).keys())
    out = [f"# Genotype Table 0:{java_double_str(score_a)} - 100.00, 2:{java_double_str(score_b)} - {java_double_str(score_a)}, "
           f"1:{java_double_str(score_n)} - {java_double_str(score_b)}, -1: <={java_double_str(score_n)}\n",
           "ID\tCHR\tSTART\tEND" + "".join("\t" + s for s in samples) + "\n"]
    cmap = []
    for r in rows:
        if r.seq not in contig_names:
            raise KcfError(f"Contig {r.seq} not found in the KCF header")
        cid = contig_names.index(r.seq) + 1
        ent = f"{r.seq}\t{cid}"
        if ent not in cmap:
            cmap.append(ent)
        if chrs is not None and r.seq not in chrs:
            continue
        al = []
        for s in samples:
            sc = r.data[s].score
            al.append(0 if sc >= score_a else (2 if sc >= score_b else (-1 if sc <= score_n else 1)))
        n = len(al)
        c0, c1, c2, cn = al.count(0), al.count(1), al.count(2), al.count(-1)
        valid = n - cn
        bad = ((c0 == n or c1 == n or c2 == n or cn == n)
               or (valid > 0 and (c0 <= min_maf * valid or c2 <= min_maf * valid))
               or (cn >= max_missing * n or (cn + c1) >= max_missing * n))
        if bad and (min_maf > 0.0 or max_missing < 1.0):
            continue
        out.append("\t".join([r.wid, str(cid), str(r.start), str(r.end)] + [str(a) for a in al]) + "\n")
    return "".join(out), "contigName\tcontigID\n" + "".join(e + "\n" for e in cmap)
