"""Second, independent restatement of the getVariations hot path in plain Python (tiny inputs only).

TEST INFRASTRUCTURE ONLY (see oracle/kcf_oracle.c).  PARITY UNPINNED for the same reason: the
reference has no tests and cannot run here.  This file exists so that the C oracle is checked by
something written differently: it works on strings and Python ints where the C file works on packed
words, and it parses the KMC files with `struct`.

Citations use the abbreviations of SURVEY.md (P/ = Plugins, D/ = Data, U/ = Utils).
"""
from __future__ import annotations

import struct

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_BITS = {"A": 0, "C": 1, "G": 2, "T": 3}


# ---- D/Signature.java ----------------------------------------------------------------------------
def _mmer_str(v: int, L: int) -> str:
    return "".join("ACGT"[(v >> (2 * (L - 1 - i))) & 3] for i in range(L))


def is_allowed(m: str) -> bool:
    """D/Signature.java:42-76 restated on the m-mer's text."""
    L = len(m)
    if m.endswith("TTT") or m.endswith("TGT"):
        return False
    if m[-3:-1] == "TT":                      # (signature & 0x3C) == 0x3C: the code tests 'TT*' (its comment says TG*)
        return False
    # loop j = 0 .. L-4 checks (sig >> 2j) & 0xF == 0  => bases (L-2-j, L-1-j) == 'AA'
    for j in range(L - 3):
        if m[L - 2 - j] == "A" and m[L - 1 - j] == "A":
            return False
    head = m[:3]                              # what is left after L-3 shifts
    if head == "AAA" or head == "ACA":
        return False
    if head[1:] == "AA":
        return False
    return True


def revcomp(s: str) -> str:
    return "".join(_COMP[c] for c in reversed(s))


def to_int(s: str) -> int:
    v = 0
    for c in s:
        v = (v << 2) | _BITS[c]
    return v


def norm_table(L: int) -> list[int]:
    """D/Signature.java:23-37"""
    special = 1 << (2 * L)
    out = []
    for i in range(special):
        m = _mmer_str(i, L)
        r = revcomp(m)
        a = i if is_allowed(m) else special
        b = to_int(r) if is_allowed(r) else special
        out.append(min(a, b))
    return out


# ---- D/KMC.java -----------------------------------------------------------------------------------
class PyKMC:
    def __init__(self, pre: bytes, suf: bytes):
        """D/KMC.java:107-168 readPrefixFile + :84-102"""
        size = len(pre)
        (header_offset,) = struct.unpack_from("<i", pre, size - 8)
        hp = size - header_offset - 8
        (self.k, self.mode, self.counter_size, self.P, self.L, self.min_count, self.max_count,
         self.total) = struct.unpack_from("<7iq", pre, hp)
        self.both_strands = pre[hp + 36] == 0
        (self.version,) = struct.unpack_from("<i", pre, hp + 36 + 1 + 3 + 24)
        if self.version != 0x200:
            raise ValueError("KMC version is not 0x200")
        nmap = (1 << (2 * self.L)) + 1
        map_start = size - header_offset - 8 - nmap * 4
        self.sigmap = struct.unpack_from(f"<{nmap}i", pre, map_start)
        self.lut_size = 1 << (2 * self.P)
        n_arrays = (map_start - 8 - 4) // (self.lut_size * 8)
        self.prefix_array = struct.unpack_from(f"<{n_arrays * self.lut_size}q", pre, 4)
        self.nsb = (self.k - self.P) // 4
        self.rec = self.counter_size + self.nsb
        self.suf = suf[4:]
        self.norm = norm_table(self.L)

    def signature(self, kmer: str) -> int:
        """D/Kmer.java:105-118 on the (already canonical) k-mer"""
        return min(self.norm[to_int(kmer[i:i + self.L])] for i in range(self.k - self.L + 1))

    def get_count(self, kmer: str) -> int:
        """D/KMC.java:292-326; kmer already canonicalised by the caller (P/GetVariants.java:222)"""
        sig = self.signature(kmer)
        prefix = to_int(kmer[:self.P]) if self.P else 0
        sfx = kmer[self.P:]
        suffix = bytes(to_int(sfx[4 * i:4 * i + 4]) for i in range(self.nsb))
        idx = self.sigmap[sig] * self.lut_size + prefix
        start = self.prefix_array[idx]
        end = self.total - 1 if idx + 1 >= len(self.prefix_array) else self.prefix_array[idx + 1] - 1
        while start <= end:
            mid = (start + end) // 2
            e = self.suf[mid * self.rec:(mid + 1) * self.rec]
            es = e[:self.nsb]
            if suffix < es:          # bytes compare unsigned-lexicographically (U/HelperFunctions.java:232-243)
                end = mid - 1
            elif suffix > es:
                start = mid + 1
            else:
                c = int.from_bytes(e[self.nsb:], "little")
                return c - (1 << 32) if c >= (1 << 31) else c   # Java int
        return 0

    def canonical(self, kmer: str) -> str:
        """D/Kmer.java:72-79: smaller of forward / reverse complement, tie keeps forward"""
        if not self.both_strands:
            return kmer
        r = revcomp(kmer)
        return r if r < kmer else kmer


# ---- D/FastaIndex.java:122-182 ---------------------------------------------------------------------
def get_sequence(raw: bytes, line_bases: int, line_width: int, seq_len: int, start: int, length: int) -> str:
    end = start + length
    if start < 0 or end > seq_len or start >= end:
        raise ValueError("Invalid range")
    pos = (start // line_bases) * line_width + start % line_bases
    col = start % line_bases
    out = []
    todo = end - start
    while todo > 0:
        n = min(todo, line_bases - col)
        if pos + n > len(raw):
            raise ValueError("Error reading sequence")
        out.append(raw[pos:pos + n].decode("latin-1"))
        pos += n + (line_width - line_bases)
        if pos > len(raw):
            raise ValueError("Error reading sequence")
        todo -= n
        col = 0
    return "".join(out)


# ---- D/Fasta.java:90-134 + P/GetVariants.java:202-273 + D/Data.java:95-107 ------------------------------
def kmers_list(seq: str, k: int) -> list[str]:
    out = []
    run = ""
    for ch in seq:
        b = ch.upper() if "a" <= ch <= "z" else ch
        if b not in "ACGT":
            run = ""
            continue
        run += b
        if len(run) >= k:
            out.append(run[-k:])
    return out


def effective_atgc(seq: str, k: int) -> int:
    """D/Fasta.java:140-167"""
    total = 0
    stretch = 0
    for ch in seq + "N":
        if ch.upper() in "ACGT" and ch.isascii():
            stretch += 1
        else:
            if stretch >= k:
                total += stretch
            stretch = 0
    return total


def get_distance(k: int, gap: int) -> int:
    d = gap - (k - 1)
    return abs(d + 1) if d <= 0 else d


def process_window(db: PyKMC, seq: str, min_count: int = 1, w=(0.3, 0.3, 0.4)) -> dict:
    total = obs = var = inner = gap = left = right = 0
    ksum = 0
    tail = True
    for km in kmers_list(seq, db.k):
        total += 1
        c = db.get_count(db.canonical(km))
        if c >= min_count:
            ksum += c
            obs += 1
            if gap > 0:
                var += 1
                if tail:
                    left += gap
                else:
                    inner += get_distance(db.k, gap)
            tail = False
            gap = 0
        else:
            gap += 1
    if gap > 0:
        var += 1
        right += gap
    eff = effective_atgc(seq, db.k)
    if obs == 0 or total == 0 or eff == 0:
        score = 0.0
    else:
        if w[0] + w[1] + w[2] != 1.0:
            raise ValueError("Weights should sum to 1.0")
        score = ((w[2] * (obs / total)) + (w[0] * (1.0 - (inner / eff))) + (w[1] * (1.0 - ((left + right) / eff)))) * 100.0
    return dict(total_kmers=total, eff_len=eff, obs=obs, variations=var, inner=inner, left=left, right=right,
                kmer_count_sum=ksum, score=score)


def windows_fixed(seq_len: int, window: int, step: int, k: int) -> list[tuple[int, int]]:
    """P/GetVariants.java:292-320"""
    out = []
    if step > 0:
        pos = 0
        while pos < seq_len:
            e = min(pos + window, seq_len)
            if e - pos >= k:
                out.append((pos, e))
            pos += step
    else:
        last_end = 0
        while last_end < seq_len:
            s = max(0, last_end - k + 1)
            e = min(s + window, seq_len)
            if e - s >= k:
                out.append((s, e))
            last_end = e
    return out
