"""CPU restatement of the reference's HOST logic around the hot path — test infrastructure only (see oracle/README in
the header of kcf_oracle.c): nothing under kcftools_b200/ may import this.  PARITY UNPINNED like the rest of oracle/:
the reference ships no tests or fixtures and cannot run here (no JVM).

  faidx_generate / faidx_offsets   Data/FastaIndex.java:239-299 (generateIndexFile), :54-68 (per-sequence slice)
  parse_gtf, Gtf.children/loci     Data/GTF.java:26-100, 156-163, 207-217, 278-293
  merged_loci                      Data/GTF.java:223-248, 278-293, 372-444 (HashSet order, stable sorts, same-strand merge)
  windows_of                       Plugins/GetVariants.java:278-352
  java_format_2f / java_float_str  java.util.Formatter "%.2f" (HALF_UP on the shortest repr), Float.toString
  kcf_row / kcf_header             Data/Window.java:125-152, 170-214; Data/Data.java:70-107, 120-132; Data/KCFHeader.java:291-330
"""
from __future__ import annotations

import struct
from decimal import ROUND_HALF_UP, Decimal, localcontext

import numpy as np

VALID_FASTA = set("ACGTYRWSMKHBVDNacgtyrwsmkhbvdn")


def java_lines(text: str) -> list[str]:
    """BufferedReader.readLine: \\n, \\r or \\r\\n end a line; no empty last line after a final terminator."""
    out, cur, i = [], [], 0
    while i < len(text):
        c = text[i]
        if c == "\n" or c == "\r":
            out.append("".join(cur))
            cur = []
            if c == "\r" and i + 1 < len(text) and text[i + 1] == "\n":
                i += 1
        else:
            cur.append(c)
        i += 1
    if cur:
        out.append("".join(cur))
    return out


def java_split(s: str, sep: str) -> list[str]:
    parts = s.split(sep)
    while parts and parts[-1] == "":
        parts.pop()
    return parts if parts or s != "" else [""]


def faidx_generate(fasta_text: str) -> list[tuple[str, int, int, int, int]]:
    """rows (name, length, offset, lineBases, lineWidth) as FastaIndex.generateIndexFile writes them"""
    rows = []
    offset = 0
    name = None
    seq_len = line_bases = line_width = start = 0
    for line in java_lines(fasta_text):
        if line.startswith(">"):
            if name is not None:
                rows.append((name, seq_len, start, line_bases, line_width))
            name = line[1:].split(" ")[0]
            offset += len(line) + 1
            start = offset
            seq_len = 0
        else:
            assert all(c in VALID_FASTA for c in line)
            if seq_len == 0:
                line_bases, line_width = len(line), len(line) + 1
            seq_len += len(line)
            offset += len(line) + 1
    if name is not None:
        rows.append((name, seq_len, start, line_bases, line_width))
    return rows


# ---------------------------------------------------------------------------------------------- GTF
def _i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def java_string_hash(s: str) -> int:
    """String.hashCode(): over the UTF-16 code units (a surrogate pair for code points above U+FFFF)"""
    h = 0
    units = s.encode("utf-16-le", "surrogatepass")
    for i in range(0, len(units), 2):
        h = (31 * h + (units[i] | (units[i + 1] << 8))) & 0xFFFFFFFF
    return _i32(h)


class Gtf:
    TRANSCRIPT_TYPES = {"transcript", "mRNA", "RNA", "lnc_RNA", "rRNA", "tRNA", "snRNA", "snoRNA"}

    def __init__(self, text: str):
        self.children: dict[str, list[str]] = {}   # insertion-ordered adjacency, no parallel edges
        self.features: dict[str, list] = {}        # id -> [chrom, start, end, strand, type]
        exon_counts: dict[str, int] = {}
        for line in java_lines(text):
            if line.startswith("#") or all(ord(c) <= 32 for c in line):  # line.trim().isEmpty()
                continue
            f = java_split(line, "\t")
            assert len(f) >= 9, line
            attrs = {}
            for a in java_split(f[8], ";"):
                t = a.strip("".join(chr(c) for c in range(33))).replace('"', "")
                pair = java_split(t, " ")
                if len(pair) == 2:
                    attrs[pair[0]] = pair[1]
            typ, chrom = f[2], f[0]
            self._vertex(chrom)
            if typ in ("gene", "pseudogene"):
                fid, parent = attrs.get("gene_id"), chrom
            elif typ in self.TRANSCRIPT_TYPES:
                fid, parent = attrs["transcript_id"], attrs["gene_id"]
                assert fid != parent
                s, e = int(f[3]), int(f[4])
                if parent not in self.children:
                    self._vertex(parent)
                    self._edge(chrom, parent)
                    self.features[parent] = [chrom, s, e, f[6][0], "gene"]
                g = self.features.get(parent)
                if g is not None:
                    g[1] = min(g[1], s)
                    g[2] = max(g[2], e)
            elif typ == "exon":
                parent = attrs.get("transcript_id")
                key = parent if parent is not None else "null"
                exon_counts[key] = exon_counts.get(key, 0) + 1
                fid = f"{key}-e-{exon_counts[key]}"
            else:
                continue
            self.features[fid] = [f[0], int(f[3]), int(f[4]), f[6][0], typ]
            self._vertex(fid)
            if parent is not None:
                self._edge(parent, fid)

    def _vertex(self, v):
        self.children.setdefault(v, [])

    def _edge(self, a, b):
        self._vertex(a)
        self._vertex(b)
        if b not in self.children[a]:
            self.children[a].append(b)

    def kids(self, v: str) -> list[str]:
        return [c for c in self.children.get(v, []) if c != v]

    def loci(self, fid: str):
        c, s, e, st, _ = self.features[fid]
        return (c, s, e, st)

    @staticmethod
    def hashset_order(ins: list[tuple]) -> list[tuple]:
        uniq = []
        for l in ins:
            if l not in uniq:
                uniq.append(l)
        cap = 16
        while len(uniq) > cap * 3 // 4:
            cap *= 2
        keyed = []
        for i, (c, s, e, st) in enumerate(uniq):
            h = java_string_hash(c) & 0xFFFFFFFF
            h = (31 * h + (s & 0xFFFFFFFF)) & 0xFFFFFFFF
            h = (31 * h + (e & 0xFFFFFFFF)) & 0xFFFFFFFF
            h = (31 * h + (java_string_hash(st) & 0xFFFFFFFF)) & 0xFFFFFFFF
            h ^= h >> 16
            keyed.append((h & (cap - 1), i))
        keyed.sort()
        return [uniq[i] for _, i in keyed]

    @staticmethod
    def _cmp_key(l):
        return (l[0], l[1])

    def merged_loci(self, fid: str, is_gene: bool) -> list[tuple]:
        if fid not in self.children:
            return []
        ins = []
        for t in self.kids(fid):
            for ex in (self.kids(t) if is_gene else [t]):
                if ex in self.features:
                    c, s, e, st, _ = self.features[ex]
                    ins.append((c, s, e, st))
        if not ins:
            return []
        srt = sorted(self.hashset_order(ins), key=self._cmp_key)  # stable, like Collections.sort
        merged = []
        for cur in srt:
            if merged:
                last = merged[-1]
                if last[0] == cur[0] and last[3] == cur[3] and last[1] <= cur[2] and cur[1] <= last[2]:
                    merged[-1] = (last[0], min(last[1], cur[1]), max(last[2], cur[2]), last[3])
                    continue
            merged.append(cur)
        return sorted(merged, key=self._cmp_key)


def windows_of(feature: str, seq_names: list[str], seq_lens: list[int], k: int, window: int = 0, step: int = 0, gtf: Gtf | None = None):
    """[(window_id, seq_name, start, end, [(seq_id, start0, len), ...] or None)] in the reference's generation order"""
    sid = {n: i for i, n in enumerate(seq_names)}
    out = []
    for name, n in zip(seq_names, seq_lens):
        if feature == "window":
            if step > 0:
                pos = 0
                while pos < n:
                    s, e = pos, min(pos + window, n)
                    if e - s >= k:
                        out.append((f"{name}_{s}", name, s, e, [(sid[name], s, e - s)]))
                    pos += step
            else:
                last_end = 0
                while last_end < n:
                    s = max(0, last_end - k + 1)
                    e = min(s + window, n)
                    if e - s >= k:
                        out.append((f"{name}_{s}", name, s, e, [(sid[name], s, e - s)]))
                    last_end = e
        else:
            is_gene = feature == "gene"
            feats = []
            for g in gtf.kids(name):
                feats += [g] if is_gene else gtf.kids(g)
            for fid in feats:
                c, s, e, _ = gtf.loci(fid)
                merged = gtf.merged_loci(fid, is_gene)
                segs = [(sid[mc], ms - 1, me - ms + 1) for (mc, ms, me, _) in merged] if merged else None
                out.append((fid, c, s, e, segs))
    return out


# ---------------------------------------------------------------------------------------------- KCF text
def java_format_2f(v: float) -> str:
    d = Decimal(repr(float(v)))  # repr = the shortest decimal that round-trips, what Formatter starts from
    with localcontext() as ctx:
        ctx.prec = 400
        q = d.quantize(Decimal("0.01"), rounding=ROUND_HALF_UP)
    s = f"{q:.2f}"
    if s.startswith("-") and float(s) == 0 and not str(v).startswith("-"):
        s = s[1:]
    return s


def java_float_str(x: float) -> str:
    f = float(np.float32(x))
    if f == 0:
        return "0.0"
    # shortest digits that round-trip as float32
    for p in range(1, 10):
        s = f"{abs(f):.{p - 1}e}"
        with np.errstate(over="ignore"):
            if float(np.float32(float(s))) == abs(f):
                break
    mant, ex = s.split("e")
    digits = mant.replace(".", "").rstrip("0") or "0"
    e10 = int(ex) + 1  # value = 0.digits * 10^e10
    sign = "-" if f < 0 else ""
    if -3 < e10 <= 7:
        if e10 <= 0:
            return sign + "0." + "0" * (-e10) + digits
        if e10 >= len(digits):
            return sign + digits + "0" * (e10 - len(digits)) + ".0"
        return sign + digits[:e10] + "." + digits[e10:]
    return sign + digits[0] + "." + (digits[1:] or "0") + "E" + str(e10 - 1)


def compute_score(obs, total, eff, inner, left, right, w=(0.3, 0.3, 0.4)) -> float:
    if obs == 0 or total == 0 or eff == 0:
        return 0.0
    assert w[0] + w[1] + w[2] == 1.0
    return float(((np.float64(w[2]) * (np.float64(obs) / np.float64(total)))
                  + (np.float64(w[0]) * (np.float64(1.0) - (np.float64(inner) / np.float64(eff))))
                  + (np.float64(w[1]) * (np.float64(1.0) - (np.float64(left + right) / np.float64(eff))))) * np.float64(100.0))


def kcf_row(seq_name: str, start: int, end: int, wid: str, r, w=(0.3, 0.3, 0.4)) -> str:
    """r: mapping with the integer fields of one window (total_kmers, eff_len, obs, variations, inner, left, right, kmer_count_sum)"""
    obs, var = int(r["obs"]), int(r["variations"])
    sc = compute_score(obs, int(r["total_kmers"]), int(r["eff_len"]), int(r["inner"]), int(r["left"]), int(r["right"]), w)
    f32max = float(np.finfo(np.float32).max)
    f32min = float(np.float32(1.401298464324817e-45))
    kd = (int(r["kmer_count_sum"]) / obs) if obs > 0 else 0.0
    info = (f"EFFLEN={int(r['eff_len'])};IS={java_format_2f(min(f32max, sc))};XS={java_format_2f(max(f32min, sc))};MS={java_format_2f(sc)};"
            f"IO={obs};XO={obs};MO={java_format_2f(float(np.float32(obs)))};IV={var};XV={var};MV={java_float_str(float(var))}")
    data = f"N:{var}:{obs}:{int(r['inner'])}:{int(r['left'])}:{int(r['right'])}:{java_format_2f(kd)}:{java_format_2f(sc)}"
    return "\t".join([seq_name, str(start), str(end), wid, str(int(r["total_kmers"])), info, "GT:VA:OB:ID:LD:RD:KD:SC", data])
