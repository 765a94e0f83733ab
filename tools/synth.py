"""Deterministic synthetic inputs for the getVariations hot path (test / bench infrastructure).

Nothing here is on the product path.  All heavy steps are torch tensor ops that
run unchanged on CPU (tests, small sizes) and on a B200 (bench, full sizes):

* random genomes and SNP / indel / deletion mutated copies,
* FASTA images with fixed line width, soft-masked (lower-case) stretches and N runs,
* a KMC 0x200 database writer (`.kmc_pre` / `.kmc_suf` byte images) laid out exactly
  as the reference parses it (D/KMC.java:107-168, 84-102, 292-326, 386-401),
* a synthetic GTF with overlapping exons across transcripts.

Data = "synthetic": there is no KMC binary or real genome in the sandbox.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np
import torch

_SIGN = -(1 << 63)  # int64 with only the top bit set: xor gives unsigned ordering


# ----------------------------------------------------------------------------------------
# genomes
# ----------------------------------------------------------------------------------------

def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def random_genome(n: int, seed: int, device="cpu") -> torch.Tensor:
    """iid uniform ACGT as uint8 codes (A=0 C=1 G=2 T=3)."""
    return torch.randint(0, 4, (n,), dtype=torch.uint8, device=device, generator=_gen(seed, device))


def mutate(codes: torch.Tensor, seed: int, snp: float = 0.01, indel: float = 0.001,
           big_deletions: int = 3, big_len: int = 100_000, replace_len: int = 200_000) -> torch.Tensor:
    """SNP / short indel / large deletion / random replacement copy of `codes` (uint8 codes)."""
    dev = codes.device
    g = _gen(seed, dev)
    n = codes.numel()
    out = codes.clone()
    # SNPs: add 1..3 mod 4
    m = torch.rand(n, device=dev, generator=g) < snp
    delta = torch.randint(1, 4, (n,), dtype=torch.uint8, device=dev, generator=g)
    out = torch.where(m, (out + delta) & 3, out)
    # one random replacement block
    if replace_len > 0 and n > 4 * replace_len:
        s = int(torch.randint(0, n - replace_len, (1,), generator=_gen(seed + 7, "cpu")))
        out[s:s + replace_len] = torch.randint(0, 4, (replace_len,), dtype=torch.uint8, device=dev, generator=g)
    # deletions (short + large) through a +1/-1 difference array
    diff = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    starts = torch.nonzero(torch.rand(n, device=dev, generator=g) < indel / 2).flatten()
    if starts.numel():
        lens = torch.clamp(torch.empty(starts.numel(), device=dev).geometric_(1.0 / 3.0, generator=g), max=20).to(torch.int64)
        diff.index_add_(0, starts, torch.ones_like(starts, dtype=torch.int32))
        diff.index_add_(0, torch.clamp(starts + lens, max=n), -torch.ones_like(starts, dtype=torch.int32))
    if big_deletions > 0 and n > 8 * big_len:
        cg = _gen(seed + 11, "cpu")
        for _ in range(big_deletions):
            s = int(torch.randint(0, n - big_len, (1,), generator=cg))
            diff[s] += 1
            diff[s + big_len] -= 1
    keep = torch.cumsum(diff[:n], 0) <= 0
    out = out[keep]
    # short insertions
    n2 = out.numel()
    ins = torch.rand(n2, device=dev, generator=g) < indel / 2
    reps = torch.ones(n2, dtype=torch.int64, device=dev)
    k = int(ins.sum())
    if k:
        lens = torch.clamp(torch.empty(k, device=dev).geometric_(1.0 / 3.0, generator=g), max=20).to(torch.int64)
        reps[ins] += lens
        first = torch.cumsum(reps, 0) - reps
        big = torch.repeat_interleave(out, reps)
        pos = torch.arange(big.numel(), device=dev) - torch.repeat_interleave(first, reps)
        rnd = torch.randint(0, 4, (big.numel(),), dtype=torch.uint8, device=dev, generator=g)
        out = torch.where(pos > 0, rnd, big)
    return out


_ASCII = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)  # A C G T


def fasta_record(codes: torch.Tensor, name: str, line: int = 60, lower: list[tuple[int, int]] | None = None,
                 n_runs: list[tuple[int, int]] | None = None, other: list[tuple[int, int]] | None = None,
                 trailing_newline: bool = True) -> np.ndarray:
    """FASTA bytes (header + folded sequence) for one sequence.  `lower`, `n_runs`, `other` are
    (start, length) intervals turned to lower case, 'N', and IUPAC 'R' respectively."""
    dev = codes.device
    seq = _ASCII.to(dev)[codes.long()]
    for (s, l) in (lower or []):
        seq[s:s + l] |= 32  # overlapping intervals must not shift a byte twice
    for (s, l) in (n_runs or []):
        seq[s:s + l] = 78
    for (s, l) in (other or []):
        seq[s:s + l] = 82
    n = seq.numel()
    full = n // line
    parts = []
    if full:
        body = seq[:full * line].view(full, line)
        nl = torch.full((full, 1), 10, dtype=torch.uint8, device=dev)
        parts.append(torch.cat([body, nl], 1).flatten())
    rem = n - full * line
    if rem:
        parts.append(seq[full * line:])
        if trailing_newline:
            parts.append(torch.tensor([10], dtype=torch.uint8, device=dev))
    elif not trailing_newline and parts:
        parts[-1] = parts[-1][:-1]
    head = np.frombuffer((">" + name + "\n").encode(), dtype=np.uint8)
    body = torch.cat(parts).cpu().numpy() if parts else np.zeros(0, np.uint8)
    return np.concatenate([head, body])


def random_intervals(n: int, count: int, lo: int, hi: int, seed: int) -> list[tuple[int, int]]:
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        l = int(rng.integers(lo, hi + 1))
        if l >= n:
            continue
        s = int(rng.integers(0, n - l))
        out.append((s, l))
    return out


# ----------------------------------------------------------------------------------------
# k-mer arithmetic on tensors (right-aligned 2k-bit values in int64; k <= 32)
# ----------------------------------------------------------------------------------------

def kmers_fwd(codes: torch.Tensor, k: int) -> torch.Tensor:
    """value of the k bases starting at every position (first base most significant)."""
    n = codes.numel()
    if n < k:
        return torch.zeros(0, dtype=torch.int64, device=codes.device)
    pw, pwlen = codes.to(torch.int64), 1
    res, reslen = None, 0
    kk = k
    while kk:
        if kk & 1:
            if res is None:
                res, reslen = pw, pwlen
            else:
                m = n - reslen - pwlen + 1
                res = (res[:m] << (2 * pwlen)) | pw[reslen:reslen + m]
                reslen += pwlen
        kk >>= 1
        if kk:
            m = n - 2 * pwlen + 1
            if m <= 0:
                # remaining powers cannot be needed (k <= n guarantees the used ones fit)
                pw = pw[:0]
            else:
                pw = (pw[:m] << (2 * pwlen)) | pw[pwlen:pwlen + m]
            pwlen *= 2
    return res


def kmers_canonical(codes: torch.Tensor, k: int, both_strands: bool = True) -> torch.Tensor:
    f = kmers_fwd(codes, k)
    if not both_strands or f.numel() == 0:
        return f
    rc = torch.flip(kmers_fwd(3 - torch.flip(codes, [0]), k), [0])
    if k == 32:
        return torch.where((f ^ _SIGN) <= (rc ^ _SIGN), f, rc)
    return torch.minimum(f, rc)


def norm_table(L: int) -> np.ndarray:
    """KMC signature 'norm' table, vectorised restatement of D/Signature.java:23-95."""
    special = 1 << (2 * L)
    m = np.arange(special, dtype=np.int64)

    def allowed(s):
        ok = (s & 0x3F) != 0x3F
        ok &= (s & 0x3F) != 0x3B
        ok &= (s & 0x3C) != 0x3C
        t = s.copy()
        for _ in range(L - 3):
            ok &= (t & 0xF) != 0
            t >>= 2
        ok &= t != 0
        ok &= t != 4
        ok &= (t & 0xF) != 0
        return ok

    rev = np.zeros_like(m)
    t = m.copy()
    for _ in range(L):
        rev = (rev << 2) | ((~t) & 3)
        t >>= 2
    a = np.where(allowed(m), m, special)
    b = np.where(allowed(rev), rev, special)
    return np.minimum(a, b).astype(np.int32)


def sliding_min(x: torch.Tensor, w: int) -> torch.Tensor:
    """min over every length-w window (doubling)."""
    n = x.numel()
    cur, cl = x, 1
    while cl * 2 <= w:
        m = n - 2 * cl + 1
        cur = torch.minimum(cur[:m], cur[cl:cl + m])
        cl *= 2
    if cl < w:
        m = n - w + 1
        cur = torch.minimum(cur[:m], cur[w - cl:w - cl + m])
    return cur


def signatures(codes: torch.Tensor, k: int, L: int, norm: torch.Tensor) -> torch.Tensor:
    """signature of every k-mer position; uses the RC symmetry of norm."""
    mm = kmers_fwd(codes, L)
    nv = norm[mm]
    return sliding_min(nv, k - L + 1)


# ----------------------------------------------------------------------------------------
# KMC 0x200 writer
# ----------------------------------------------------------------------------------------

@dataclass
class KmcImage:
    pre: np.ndarray          # uint8 image of .kmc_pre
    suf: np.ndarray          # uint8 image of .kmc_suf
    k: int
    P: int
    L: int
    n_bins: int
    counter_size: int
    total: int
    both_strands: bool

    def write(self, prefix: str) -> None:
        self.pre.tofile(prefix + ".kmc_pre")
        self.suf.tofile(prefix + ".kmc_suf")


def default_sigmap(L: int, n_bins: int) -> np.ndarray:
    """any total map signature -> bin is legal for the reference (it trusts the file)."""
    s = np.arange((1 << (2 * L)) + 1, dtype=np.int64)
    return (s % n_bins).astype(np.uint32)


def kmc_image_from_kmers(kmers: torch.Tensor, counts: torch.Tensor, sigs: torch.Tensor, *, k: int, P: int, L: int,
                         n_bins: int, counter_size: int, both_strands: bool = True, sigmap: np.ndarray | None = None,
                         min_count: int = 1, max_count: int = 255) -> KmcImage:
    """kmers: distinct right-aligned values (int64), counts: int64, sigs: signature per k-mer."""
    dev = kmers.device
    if sigmap is None:
        sigmap = default_sigmap(L, n_bins)
    smap = torch.from_numpy(sigmap.astype(np.int64)).to(dev)
    bins = smap[sigs.long()]
    # order: bin, then k-mer value (unsigned)
    key = kmers ^ _SIGN if k == 32 else kmers
    o1 = torch.argsort(key)
    kmers, counts, bins = kmers[o1], counts[o1], bins[o1]
    o2 = torch.sort(bins, stable=True).indices
    kmers, counts, bins = kmers[o2], counts[o2], bins[o2]
    del o1, o2, key
    N = kmers.numel()
    sbits = 2 * (k - P)
    if P > 0:
        prefix = (kmers >> sbits) & ((1 << (2 * P)) - 1)
    else:
        prefix = torch.zeros_like(kmers)
    idx = bins * (1 << (2 * P)) + prefix
    hist = torch.bincount(idx, minlength=n_bins << (2 * P))
    lut = (torch.cumsum(hist, 0) - hist).cpu().numpy().astype("<u8")
    nsb = (k - P) // 4
    assert (k - P) % 4 == 0
    rec = torch.empty((N, nsb + counter_size), dtype=torch.uint8, device=dev)
    suffix = kmers & ((1 << sbits) - 1) if sbits < 64 else kmers
    for j in range(nsb):
        rec[:, j] = ((suffix >> (8 * (nsb - 1 - j))) & 0xFF).to(torch.uint8)
    for j in range(counter_size):
        rec[:, nsb + j] = ((counts >> (8 * j)) & 0xFF).to(torch.uint8)
    mark = torch.tensor(list(b"KMCS"), dtype=torch.uint8, device=dev)
    suf = torch.cat([mark, rec.flatten(), mark]).cpu().numpy()  # one device-side assembly, one D2H
    del rec
    header = struct.pack("<7IQB3x24xI", k, 0, counter_size, P, L, min_count, max_count, N, 0 if both_strands else 1, 0x200)
    assert len(header) == 68
    pre = np.concatenate([
        np.frombuffer(b"KMCP", np.uint8),
        lut.view(np.uint8),
        np.frombuffer(struct.pack("<Q", N), np.uint8),          # guard word after the LUTs
        sigmap.astype("<u4").view(np.uint8),
        np.frombuffer(header, np.uint8),
        np.frombuffer(struct.pack("<I", 68), np.uint8),
        np.frombuffer(b"KMCP", np.uint8),
    ])
    return KmcImage(pre=pre, suf=suf, k=k, P=P, L=L, n_bins=n_bins, counter_size=counter_size, total=N,
                    both_strands=both_strands)


def kmc_image_from_genomes(genomes: list[torch.Tensor], *, k: int = 31, P: int = 7, L: int = 9, n_bins: int = 512,
                           counter_size: int = 1, both_strands: bool = True, coverage: float = 8.0, seed: int = 1,
                           sigmap: np.ndarray | None = None) -> KmcImage:
    """KMC DB of all k-mers of `genomes` (uint8 code tensors, no N) with Poisson counts."""
    dev = genomes[0].device
    norm = torch.from_numpy(norm_table(L)).to(dev)
    ks, ss = [], []
    for gseq in genomes:
        if gseq.numel() < k:
            continue
        f = kmers_fwd(gseq, k)
        if both_strands:
            rc = torch.flip(kmers_fwd(3 - torch.flip(gseq, [0]), k), [0])
            can = torch.where((f ^ _SIGN) <= (rc ^ _SIGN), f, rc) if k == 32 else torch.minimum(f, rc)
            del rc
        else:
            can = f
        del f
        ks.append(can)
        ss.append(signatures(gseq, k, L, norm).to(torch.int32))
    allk = torch.cat(ks)
    alls = torch.cat(ss)
    del ks, ss
    key = allk ^ _SIGN if k == 32 else allk
    order = torch.argsort(key)
    del key
    allk, alls = allk[order], alls[order]
    del order
    first = torch.ones(allk.numel(), dtype=torch.bool, device=dev)
    first[1:] = allk[1:] != allk[:-1]
    pos = torch.nonzero(first).flatten()
    mult = torch.diff(pos, append=torch.tensor([allk.numel()], device=dev))
    uk, us = allk[pos], alls[pos]
    del allk, alls, first, pos
    g = _gen(seed, dev)
    maxc = (1 << (8 * counter_size)) - 1 if counter_size > 0 else 0
    cnt = torch.poisson(mult.to(torch.float32) * coverage, generator=g).to(torch.int64)
    if counter_size > 0:
        cnt = torch.clamp(cnt, max=min(maxc, 255 if counter_size == 1 else maxc))
        keep = cnt > 0
        uk, us, cnt = uk[keep], us[keep], cnt[keep]
    else:
        cnt = torch.zeros_like(cnt)
    return kmc_image_from_kmers(uk, cnt, us, k=k, P=P, L=L, n_bins=n_bins, counter_size=counter_size,
                                both_strands=both_strands, sigmap=sigmap, max_count=max(maxc, 1))


# ----------------------------------------------------------------------------------------
# FASTA file + .faidx (same columns as D/FastaIndex.java:239-299 writes)
# ----------------------------------------------------------------------------------------

@dataclass
class FastaImage:
    data: np.ndarray                                  # whole file bytes
    names: list[str]
    lengths: list[int]
    offsets: list[int]                                # byte offset of first base of each sequence
    line_bases: list[int]
    line_width: list[int]

    def seq_bytes(self, i: int) -> np.ndarray:
        """the slice the reference maps for sequence i (D/FastaIndex.java:54-68)."""
        end = self.offsets[i + 1] - 0 if i + 1 < len(self.names) else self.data.size
        if i + 1 < len(self.names):
            end = self.offsets[i + 1]
        return self.data[self.offsets[i]:end]

    def write(self, path: str) -> None:
        self.data.tofile(path)


def fasta_image(records: list[tuple[str, np.ndarray, int, int]]) -> FastaImage:
    """records: (name, record bytes incl. header, seq_len, line)"""
    names, lengths, offsets, lbs, lws, chunks = [], [], [], [], [], []
    off = 0
    for (name, rec, seqlen, line) in records:
        hdr = len(name) + 2
        names.append(name)
        lengths.append(seqlen)
        offsets.append(off + hdr)
        lb = min(line, seqlen) if seqlen > 0 else 0
        lbs.append(lb)
        lws.append(lb + 1)
        chunks.append(rec)
        off += rec.size
    return FastaImage(np.concatenate(chunks), names, lengths, offsets, lbs, lws)


# ----------------------------------------------------------------------------------------
# synthetic GTF
# ----------------------------------------------------------------------------------------

def synthetic_gtf(chroms: list[tuple[str, int]], genes_per_chrom: int, seed: int, max_tx: int = 3, max_exons: int = 12,
                  exon_lo: int = 50, exon_hi: int = 2000) -> str:
    """gene / transcript / exon lines with gene_id / transcript_id attributes; exons of different
    transcripts of a gene deliberately overlap; coordinates 1-based inclusive."""
    rng = np.random.default_rng(seed)
    lines = []
    gid = 0
    for (chrom, clen) in chroms:
        span = clen // max(genes_per_chrom, 1)
        for gi in range(genes_per_chrom):
            gid += 1
            gname = f"G{gid:06d}"
            strand = "+" if rng.random() < 0.5 else "-"
            base = gi * span + 1
            budget = span - 10
            ntx = int(rng.integers(1, max_tx + 1))
            txs = []
            for t in range(ntx):
                nex = int(rng.integers(1, max_exons + 1))
                pos = base + int(rng.integers(0, max(1, budget // 8)))
                exons = []
                for _ in range(nex):
                    el = int(rng.integers(exon_lo, exon_hi + 1))
                    if pos + el >= base + budget:
                        break
                    exons.append((pos, pos + el - 1))
                    pos += el + int(rng.integers(0, max(2, budget // (4 * max_exons))))  # gap 0 => abutting exons
                if not exons:
                    el = min(exon_lo, budget - 1)
                    exons = [(base, base + el - 1)]
                txs.append(exons)
            gs = min(e[0] for ex in txs for e in ex)
            ge = max(e[1] for ex in txs for e in ex)
            lines.append(f'{chrom}\tsynth\tgene\t{gs}\t{ge}\t.\t{strand}\t.\tgene_id "{gname}";')
            for t, exons in enumerate(txs):
                tname = f"{gname}.t{t + 1}"
                lines.append(f'{chrom}\tsynth\ttranscript\t{exons[0][0]}\t{exons[-1][1]}\t.\t{strand}\t.\t'
                             f'gene_id "{gname}"; transcript_id "{tname}";')
                for (s, e) in exons:
                    lines.append(f'{chrom}\tsynth\texon\t{s}\t{e}\t.\t{strand}\t.\t'
                                 f'gene_id "{gname}"; transcript_id "{tname}";')
    return "\n".join(lines) + "\n"


# ----------------------------------------------------------------------------------------
# any k (test scale): the same file layout from Python strings and ints, for k-mers wider than one word
# ----------------------------------------------------------------------------------------
def kmc_image_from_strings(seqs: list[str], *, k: int, P: int, L: int, n_bins: int, counter_size: int = 1, both_strands: bool = True,
                           coverage: float = 8.0, seed: int = 1, sigmap: np.ndarray | None = None) -> KmcImage:
    """KMC 0x200 image of all k-mers of `seqs` (strings over ACGT; other characters break the runs), any k with
    (k - P) % 4 == 0.  Pure Python: thousands of k-mers, not millions."""
    assert (k - P) % 4 == 0 and 0 <= P <= 15 and L <= k
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    norm = norm_table(L)
    if sigmap is None:
        sigmap = default_sigmap(L, n_bins)
    rng = np.random.default_rng(seed)
    counts: dict[str, int] = {}
    for s in seqs:
        s = s.upper()
        run = 0
        for i, ch in enumerate(s):
            run = run + 1 if ch in code else 0
            if run >= k:
                km = s[i - k + 1:i + 1]
                if both_strands:
                    rc = "".join(comp[c] for c in reversed(km))
                    if rc < km:
                        km = rc
                counts[km] = counts.get(km, 0) + int(rng.poisson(coverage))
    cmax = (1 << (8 * counter_size)) - 1 if counter_size else 0
    recs = []
    for km, c in counts.items():
        if counter_size and c == 0:
            continue  # KMC drops k-mers below its minimum count
        val = 0
        for ch in km:
            val = (val << 2) | code[ch]
        sig = min(int(norm[(val >> (2 * (k - L - j))) & ((1 << (2 * L)) - 1)]) for j in range(k - L + 1))
        recs.append((int(sigmap[sig]), val, min(c, cmax)))
    recs.sort()  # bin, then k-mer value: inside a bin ascending (prefix, suffix)
    N = len(recs)
    sbits = 2 * (k - P)
    nsb = (k - P) // 4
    hist = np.zeros(n_bins << (2 * P), np.int64)
    body = bytearray()
    for b, val, c in recs:
        hist[(b << (2 * P)) + (val >> sbits)] += 1
        body += (val & ((1 << sbits) - 1)).to_bytes(nsb, "big") + c.to_bytes(counter_size, "little")
    lut = (np.cumsum(hist) - hist).astype("<u8")
    suf = np.frombuffer(b"KMCS" + bytes(body) + b"KMCS", np.uint8).copy()
    header = struct.pack("<7IQB3x24xI", k, 0, counter_size, P, L, 1, max(cmax, 1), N, 0 if both_strands else 1, 0x200)
    pre = np.concatenate([
        np.frombuffer(b"KMCP", np.uint8),
        lut.view(np.uint8),
        np.frombuffer(struct.pack("<Q", N), np.uint8),
        sigmap.astype("<u4").view(np.uint8),
        np.frombuffer(header, np.uint8),
        np.frombuffer(struct.pack("<I", 68), np.uint8),
        np.frombuffer(b"KMCP", np.uint8),
    ])
    return KmcImage(pre=pre, suf=suf, k=k, P=P, L=L, n_bins=n_bins, counter_size=counter_size, total=N, both_strands=both_strands)


# ----------------------------------------------------------------------------------------
# large databases (3e9 records and more): the same image, built one group of bins at a time so that the sort buffers of
# a group — not of the whole database — have to fit the device
# ----------------------------------------------------------------------------------------
def kmc_image_from_genomes_grouped(genomes: list[torch.Tensor], *, k: int = 31, P: int = 7, L: int = 9, n_bins: int = 512,
                                   counter_size: int = 1, coverage: float = 8.0, seed: int = 1, groups: int = 4,
                                   both_strands: bool = True) -> KmcImage:
    """as kmc_image_from_genomes (counter_size >= 1, default signature map), but the bins are processed in `groups`
    contiguous ranges: every pass walks all genomes, keeps the k-mers whose bin falls in the range, sorts / deduplicates /
    draws counts for those, and appends their records.  Bins are contiguous in `.kmc_suf` (KMC.java:61, 300-307), so the
    concatenation of the groups IS the file."""
    assert counter_size >= 1 and (k - P) % 4 == 0 and k <= 32
    dev = genomes[0].device
    norm = torch.from_numpy(norm_table(L)).to(dev)
    sigmap = default_sigmap(L, n_bins)
    smap = torch.from_numpy(sigmap.astype(np.int64)).to(dev)
    sbits = 2 * (k - P)
    nsb = (k - P) // 4
    maxc = min((1 << (8 * counter_size)) - 1, 255 if counter_size == 1 else (1 << 31) - 1)
    hist_all = torch.zeros(n_bins << (2 * P), dtype=torch.int64, device=dev)
    rec_size = nsb + counter_size
    upper = sum(max(int(gq.numel()) - k + 1, 0) for gq in genomes)  # distinct k-mers cannot outnumber positions
    buf = np.empty(8 + upper * rec_size, np.uint8)  # one buffer, filled group by group (untouched pages cost nothing)
    buf[:4] = np.frombuffer(b"KMCS", np.uint8)
    used = 4
    total = 0
    for gi in range(groups):
        b_lo, b_hi = n_bins * gi // groups, n_bins * (gi + 1) // groups
        ks, bs = [], []
        for gseq in genomes:
            if gseq.numel() < k:
                continue
            f = kmers_fwd(gseq, k)
            if both_strands:
                rc = torch.flip(kmers_fwd(3 - torch.flip(gseq, [0]), k), [0])
                can = torch.where((f ^ _SIGN) <= (rc ^ _SIGN), f, rc) if k == 32 else torch.minimum(f, rc)
                del rc
            else:
                can = f
            del f
            bins = smap[signatures(gseq, k, L, norm).long()]
            sel = (bins >= b_lo) & (bins < b_hi)
            ks.append(can[sel])
            bs.append(bins[sel].to(torch.int16))
            del can, bins, sel
        allk, allb = torch.cat(ks), torch.cat(bs)
        del ks, bs
        order = torch.argsort(allk ^ _SIGN if k == 32 else allk)
        allk, allb = allk[order], allb[order]
        del order
        first = torch.ones(allk.numel(), dtype=torch.bool, device=dev)
        first[1:] = allk[1:] != allk[:-1]
        pos = torch.nonzero(first).flatten()
        del first
        mult = torch.diff(pos, append=torch.tensor([allk.numel()], device=dev))
        uk, ub = allk[pos], allb[pos].to(torch.int64)
        del allk, allb, pos
        cnt = torch.clamp(torch.poisson(mult.to(torch.float32) * coverage, generator=_gen(seed + 1000 * gi, dev)).to(torch.int64), max=maxc)
        del mult
        keep = cnt > 0
        uk, ub, cnt = uk[keep], ub[keep], cnt[keep]
        del keep
        o2 = torch.sort(ub, stable=True).indices  # bin, then k-mer value (already ascending)
        uk, ub, cnt = uk[o2], ub[o2], cnt[o2]
        del o2
        prefix = (uk >> sbits) & ((1 << (2 * P)) - 1) if P > 0 else torch.zeros_like(uk)
        hist_all += torch.bincount(ub * (1 << (2 * P)) + prefix, minlength=n_bins << (2 * P))
        del prefix, ub
        n = uk.numel()
        rec = torch.empty((n, nsb + counter_size), dtype=torch.uint8, device=dev)
        suffix = uk & ((1 << sbits) - 1) if sbits < 64 else uk
        for j in range(nsb):
            rec[:, j] = ((suffix >> (8 * (nsb - 1 - j))) & 0xFF).to(torch.uint8)
        for j in range(counter_size):
            rec[:, nsb + j] = ((cnt >> (8 * j)) & 0xFF).to(torch.uint8)
        buf[used:used + n * rec_size] = rec.flatten().cpu().numpy()
        used += n * rec_size
        total += n
        del rec, suffix, uk, cnt
    buf[used:used + 4] = np.frombuffer(b"KMCS", np.uint8)
    suf = buf[:used + 4]
    lut = (torch.cumsum(hist_all, 0) - hist_all).cpu().numpy().astype("<u8")
    header = struct.pack("<7IQB3x24xI", k, 0, counter_size, P, L, 1, max(maxc, 1), total, 0 if both_strands else 1, 0x200)
    pre = np.concatenate([
        np.frombuffer(b"KMCP", np.uint8),
        lut.view(np.uint8),
        np.frombuffer(struct.pack("<Q", total), np.uint8),
        sigmap.astype("<u4").view(np.uint8),
        np.frombuffer(header, np.uint8),
        np.frombuffer(struct.pack("<I", 68), np.uint8),
        np.frombuffer(b"KMCP", np.uint8),
    ])
    return KmcImage(pre=pre, suf=suf, k=k, P=P, L=L, n_bins=n_bins, counter_size=counter_size, total=total, both_strands=both_strands)
