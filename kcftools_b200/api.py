"""Python host-side handles over the C ABI (include/kcf_b200.h).

Names follow the reference's objects for this path: `KMC` is the database handle
(Data/KMC.java: getKmerLength / getPrefixLength / isBothStrands / getCount / close), a `Context`
owns one GPU and the resident reference sequences (what FastaIndex maps, Data/FastaIndex.java:26-77),
and `Context.screen` is the replacement of the per-window fan-out of
Plugins/GetVariants.java:126-159.  Everything computes on the GPU through libkcfgpu.so; there is no
CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RESULT_DTYPE, SEGMENT_DTYPE, WINDOW_DTYPE, DbInfo, KcfError


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


class Context:
    """one GPU (CUDA ordinal `device`) + its stream + the resident reference sequences."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.kcf_init(device, C.byref(self._h))
        if rc:
            raise KcfError(rc, self._lib.kcf_last_error(None).decode())
        self.device = device
        self.n_seqs = 0

    # -- plumbing --
    def _check(self, rc: int):
        if rc:
            raise KcfError(rc, self._lib.kcf_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kcf_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self) -> int:
        """cudaStream_t every call of this context is ordered on."""
        return self._lib.kcf_stream(self._h) or 0

    def pinned(self, n_bytes: int) -> np.ndarray:
        """uint8 view of page-locked host memory (freed with the context's process)."""
        p = C.c_void_p()
        self._check(self._lib.kcf_host_alloc(self._h, n_bytes, C.byref(p)))
        buf = (C.c_uint8 * max(n_bytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=n_bytes)
        arr.flags.writeable = True
        return arr

    def set_load_factor(self, lf: float):
        self._check(self._lib.kcf_set_load_factor(self._h, lf))

    def set_minimizer_length(self, m: int):
        """0 = automatic; a tuning / test knob of the table layout, results never depend on it"""
        self._check(self._lib.kcf_set_minimizer_length(self._h, m))

    def set_upload_piece(self, bases: int):
        """0 = default; upper bound of the stretches a sharded job uploads one at a time (a tuning / test knob)"""
        self._check(self._lib.kcf_set_upload_piece(self._h, bases))

    def set_partition(self, rank: int, world: int):
        """slice kept by databases opened afterwards with placement=1 (kcftools_b200/partitioned.py)"""
        self._check(self._lib.kcf_set_partition(self._h, rank, world))

    def set_profiling(self, on: bool):
        self._check(self._lib.kcf_set_profiling(self._h, int(on)))

    def last_kernel_ms(self) -> tuple[float, float]:
        a, b = C.c_float(), C.c_float()
        self._check(self._lib.kcf_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def random_sector_gbps(self, n_bytes: int, n_loads: int, repeats: int = 5) -> float:
        out = C.c_double()
        self._check(self._lib.kcf_measure_random_sector_gbps(self._h, n_bytes, n_loads, repeats, C.byref(out)))
        return out.value

    def random_line_rate(self, n_bytes: int, n_lines_read: int, repeats: int = 5) -> float:
        """random 128-byte lines per second this GPU serves (one coalesced request per line)"""
        out = C.c_double()
        self._check(self._lib.kcf_measure_random_line_rate(self._h, n_bytes, n_lines_read, repeats, C.byref(out)))
        return out.value

    # -- reference sequences --
    def ref_add(self, seq_bytes: np.ndarray, line_bases: int, line_width: int, seq_len: int) -> int:
        """seq_bytes = the file bytes the reference maps for one sequence (FastaIndex.java:54-68)."""
        a = np.ascontiguousarray(seq_bytes, np.uint8)
        sid = C.c_int()
        self._check(self._lib.kcf_ref_add(self._h, _ptr(a), a.size, line_bases, line_width, seq_len, C.byref(sid)))
        self.n_seqs += 1
        return sid.value

    def ref_add_async(self, seq_bytes: np.ndarray, line_bases: int, line_width: int, seq_len: int) -> int:
        """as ref_add, but only queued: `seq_bytes` (ideally pinned, see `pinned`) must stay untouched until ref_sync /
        Plan.fetch / screen returns."""
        a = seq_bytes
        assert a.dtype == np.uint8 and a.flags.c_contiguous
        sid = C.c_int()
        self._check(self._lib.kcf_ref_add_async(self._h, _ptr(a), a.size, line_bases, line_width, seq_len, C.byref(sid)))
        self.n_seqs += 1
        return sid.value

    def ref_sync(self):
        self._check(self._lib.kcf_ref_sync(self._h))

    def ref_clear(self):
        self._check(self._lib.kcf_ref_clear(self._h))
        self.n_seqs = 0

    # -- screening --
    def screen(self, db: "KMC", wins: np.ndarray, segs: np.ndarray, min_count: int = 1,
               weights=(0.3, 0.3, 0.4)) -> np.ndarray:
        """weights = (wi, wt, wr) as getWeights() orders them (GetVariants.java:388-390)."""
        wins = np.ascontiguousarray(wins, WINDOW_DTYPE)
        segs = np.ascontiguousarray(segs, SEGMENT_DTYPE)
        out = np.zeros(wins.size, RESULT_DTYPE)
        w = (C.c_double * 3)(*weights)
        self._check(self._lib.kcf_screen(self._h, db._h, _ptr(wins), wins.size, _ptr(segs), segs.size, min_count, w,
                                         _ptr(out)))
        return out

    def plan(self, k: int, wins: np.ndarray, segs: np.ndarray) -> "Plan":
        return Plan(self, k, wins, segs)


class KMC:
    """KMC database resident in HBM (replaces Data/KMC.java)."""

    def __init__(self, ctx: Context, prefix: str | None = None, pre: np.ndarray | None = None,
                 suf: np.ndarray | None = None, placement: int = 0):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = C.c_void_p()
        if prefix is not None:
            rc = self._lib.kcf_db_open(ctx._h, prefix.encode(), placement, C.byref(self._h))
        else:
            pre = np.ascontiguousarray(pre, np.uint8)
            suf = np.ascontiguousarray(suf, np.uint8)
            rc = self._lib.kcf_db_open_mem(ctx._h, _ptr(pre), pre.size, _ptr(suf), suf.size, placement, C.byref(self._h))
        ctx._check(rc)
        self.info = DbInfo()
        ctx._check(self._lib.kcf_db_info(self._h, C.byref(self.info)))

    def getKmerLength(self) -> int:
        return self.info.kmer_length

    def getPrefixLength(self) -> int:
        return self.info.lut_prefix_length

    def isBothStrands(self) -> bool:
        return bool(self.info.both_strands)

    def getCount(self, kmer: str) -> int:
        return int(self.getCounts([kmer])[0])

    def getCounts(self, kmers) -> np.ndarray:
        """batch KMC.getCount; kmers: list of str or an (n, k) uint8 ASCII array."""
        k = self.info.kmer_length
        if isinstance(kmers, np.ndarray):
            a = np.ascontiguousarray(kmers, np.uint8)
            assert a.ndim == 2 and a.shape[1] == k
        else:
            assert all(len(s) == k for s in kmers)
            a = np.frombuffer("".join(kmers).encode(), np.uint8).reshape(-1, k)
        out = np.zeros(a.shape[0], np.int32)
        self.ctx._check(self._lib.kcf_db_count(self.ctx._h, self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kcf_db_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Plan:
    """a window list resident on the device; run it against any database of the same k."""

    def __init__(self, ctx: Context, k: int, wins: np.ndarray, segs: np.ndarray):
        self.ctx = ctx
        self._lib = ctx._lib
        self.wins = np.ascontiguousarray(wins, WINDOW_DTYPE)
        self.segs = np.ascontiguousarray(segs, SEGMENT_DTYPE)
        self._h = C.c_void_p()
        ctx._check(self._lib.kcf_plan_create(ctx._h, k, _ptr(self.wins), self.wins.size, _ptr(self.segs), self.segs.size,
                                             C.byref(self._h)))
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint32()
        ctx._check(self._lib.kcf_plan_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_tiles, self.n_positions, self.kernels_per_run = a.value, b.value, c.value

    def run(self, db: KMC, min_count: int = 1, weights=(0.3, 0.3, 0.4)):
        """asynchronous on ctx.stream"""
        w = (C.c_double * 3)(*weights)
        self.ctx._check(self._lib.kcf_plan_run(self.ctx._h, db._h, self._h, min_count, w))

    def fetch(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.zeros(self.wins.size, RESULT_DTYPE)
        self.ctx._check(self._lib.kcf_plan_fetch(self.ctx._h, self._h, _ptr(out)))
        return out

    def window_counts(self, db: KMC, window: int) -> np.ndarray:
        """Java-int count of every valid k-mer of one window, in order (parity helper)."""
        cap = int(self.segs["len"][self.wins["first_seg"][window]:][:self.wins["n_segs"][window]].sum())
        out = np.zeros(max(cap, 1), np.int32)
        n = C.c_uint64()
        self.ctx._check(self._lib.kcf_window_counts(self.ctx._h, db._h, self._h, window, _ptr(out), cap, C.byref(n)))
        return out[:n.value]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kcf_plan_destroy(self._h)
            self._h = C.c_void_p()


class Cohort:
    """windows x samples matrix resident on the device (Plugins/Cohort.java, FindIBS.java, KCFToGenotypeTable.java)."""

    def __init__(self, ctx: Context, n_windows: int, n_samples: int, total_kmers: np.ndarray | None = None, eff_len: np.ndarray | None = None):
        self.ctx, self._lib = ctx, ctx._lib
        self.n_windows, self.n_samples = int(n_windows), int(n_samples)
        self._h = C.c_void_p()
        t = None if total_kmers is None else np.ascontiguousarray(total_kmers, np.int32)
        e = None if eff_len is None else np.ascontiguousarray(eff_len, np.int32)
        ctx._check(self._lib.kcf_cohort_create(ctx._h, self.n_windows, self.n_samples, None if t is None else _ptr(t), None if e is None else _ptr(e),
                                               C.byref(self._h)))

    def add_plan(self, sample: int, plan: "Plan", window_offset: int = 0):
        self.ctx._check(self._lib.kcf_cohort_add_plan(self.ctx._h, self._h, sample, window_offset, plan._h))

    def set_sample(self, sample: int, cells: np.ndarray):
        from ._lib import CELL_DTYPE
        cells = np.ascontiguousarray(cells, CELL_DTYPE)
        assert cells.size == self.n_windows
        self.ctx._check(self._lib.kcf_cohort_set_sample(self.ctx._h, self._h, sample, _ptr(cells)))

    def scores(self, weights=(0.3, 0.3, 0.4)):
        self.ctx._check(self._lib.kcf_cohort_scores(self.ctx._h, self._h, (C.c_double * 3)(*weights)))

    def find_ibs(self, order: np.ndarray, chrom: np.ndarray, detect_var: bool = False, min_consecutive: int = 4, score_cutoff: float = 95.0):
        order = np.ascontiguousarray(order, np.uint32)
        chrom = np.ascontiguousarray(chrom, np.uint32)
        assert order.size == chrom.size
        self.ctx._check(self._lib.kcf_cohort_find_ibs(self.ctx._h, self._h, _ptr(order), _ptr(chrom), order.size, int(detect_var), min_consecutive,
                                                      C.c_float(score_cutoff)))

    def genotypes(self, score_a=95.0, score_b=60.0, score_n=30.0, min_maf=0.0, max_missing=1.0):
        al = np.zeros((self.n_windows, self.n_samples), np.int8)
        bad = np.zeros(self.n_windows, np.uint8)
        self.ctx._check(self._lib.kcf_cohort_genotypes(self.ctx._h, self._h, score_a, score_b, score_n, min_maf, max_missing, _ptr(al), _ptr(bad)))
        return al, bad

    def fetch(self, sample: int):
        from ._lib import CELL_DTYPE
        cells = np.zeros(self.n_windows, CELL_DTYPE)
        tot = np.zeros(self.n_windows, np.int32)
        eff = np.zeros(self.n_windows, np.int32)
        self.ctx._check(self._lib.kcf_cohort_fetch(self.ctx._h, self._h, sample, _ptr(cells), _ptr(tot), _ptr(eff)))
        return cells, tot, eff

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kcf_cohort_destroy(self._h)
            self._h = C.c_void_p()


def host_seqs(seqs) -> "C.Array":
    """kcf_host_seq_t array over (bytes, line_bases, line_width, seq_len) tuples; the byte arrays must stay alive"""
    arr = (_lib.HostSeq * max(len(seqs), 1))()
    for i, (raw, lb, lw, sl) in enumerate(seqs):
        assert raw.dtype == np.uint8 and raw.flags.c_contiguous
        arr[i] = _lib.HostSeq(_ptr(raw), raw.size, int(lb), int(lw), int(sl))
    return arr


def shard_windows(wins: np.ndarray, segs: np.ndarray, n_shards: int) -> np.ndarray:
    """bounds[n_shards + 1] of the contiguous window ranges kcf_screen_sharded gives its contexts (balanced on bases)"""
    wins = np.ascontiguousarray(wins, WINDOW_DTYPE)
    segs = np.ascontiguousarray(segs, SEGMENT_DTYPE)
    out = np.zeros(n_shards + 1, np.uint64)
    rc = _lib.load().kcf_shard_windows(_ptr(wins), wins.size, _ptr(segs), segs.size, n_shards, _ptr(out))
    if rc:
        raise KcfError(rc, "kcf_shard_windows: bad window / segment arrays")
    return out


def screen_sharded(ctxs, dbs, seqs, wins: np.ndarray, segs: np.ndarray, min_count: int = 1, weights=(0.3, 0.3, 0.4),
                   out: np.ndarray | None = None) -> np.ndarray:
    """ONE job over the GPUs of this process (GetVariants.java:129-151): `dbs[g]` is the same database opened on `ctxs[g]`;
    `seqs` = [(fasta bytes, line_bases, line_width, seq_len), ...] in HOST memory (pinned for full-rate copies).  Every
    context uploads only the stretches its window range touches.  len(ctxs) == 1 is the plain host-buffer call."""
    wins = np.ascontiguousarray(wins, WINDOW_DTYPE)
    segs = np.ascontiguousarray(segs, SEGMENT_DTYPE)
    if out is None:
        out = np.zeros(wins.size, RESULT_DTYPE)
    n = len(ctxs)
    assert n >= 1 and len(dbs) == n
    hs = seqs if not isinstance(seqs, (list, tuple)) else host_seqs(seqs)
    n_seqs = len(seqs) if isinstance(seqs, (list, tuple)) else len(hs)
    ch = (C.c_void_p * n)(*[c._h for c in ctxs])
    dh = (C.c_void_p * n)(*[d._h for d in dbs])
    w = (C.c_double * 3)(*weights)
    lib = ctxs[0]._lib
    ctxs[0]._check(lib.kcf_screen_sharded(ch, dh, n, hs, n_seqs, _ptr(wins), wins.size, _ptr(segs), segs.size, min_count, w, _ptr(out)))
    for c in ctxs:
        c.n_seqs = 0  # the call replaced the contexts' sequences
    return out


def fixed_windows(seq_lens, window: int, step: int, k: int):
    """window / segment arrays of the `-f window` mode (GetVariants.java:292-320) for sequences 0..n-1.
    Returns (wins, segs, starts, ends, seq_ids)."""
    starts, ends, sids = [], [], []
    for sid, n in enumerate(seq_lens):
        if step > 0:
            pos = 0
            while pos < n:
                s, e = pos, min(pos + window, n)
                if e - s >= k:
                    starts.append(s); ends.append(e); sids.append(sid)
                pos += step
        else:
            if window <= k - 1:
                raise ValueError("window must exceed k-1 in tiling mode (the reference loop does not terminate)")
            last_end = 0
            while last_end < n:
                s = max(0, last_end - k + 1)
                e = min(s + window, n)
                if e - s >= k:
                    starts.append(s); ends.append(e); sids.append(sid)
                last_end = e
    starts = np.asarray(starts, np.int32)
    ends = np.asarray(ends, np.int32)
    sids = np.asarray(sids, np.int32)
    nw = starts.size
    wins = np.zeros(nw, WINDOW_DTYPE)
    wins["first_seg"] = np.arange(nw, dtype=np.uint32)
    wins["n_segs"] = 1
    segs = np.zeros(nw, SEGMENT_DTYPE)
    segs["seq_id"] = sids
    segs["start0"] = starts
    segs["len"] = ends - starts
    return wins, segs, starts, ends, sids
