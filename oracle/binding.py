"""ctypes binding of the CPU oracle (oracle/kcf_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of kcf_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
the product package kcftools_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkcforacle.so")


class OrcResult(C.Structure):
    _fields_ = [("total_kmers", C.c_int32), ("eff_len", C.c_int32), ("obs", C.c_int32), ("variations", C.c_int32),
                ("inner", C.c_int32), ("left", C.c_int32), ("right", C.c_int32), ("_pad", C.c_int32),
                ("kmer_count_sum", C.c_int64), ("score", C.c_double)]


RESULT_DTYPE = np.dtype([("total_kmers", "<i4"), ("eff_len", "<i4"), ("obs", "<i4"), ("variations", "<i4"),
                         ("inner", "<i4"), ("left", "<i4"), ("right", "<i4"), ("_pad", "<i4"),
                         ("kmer_count_sum", "<i8"), ("score", "<f8")])
WINDOW_DTYPE = np.dtype([("first_seg", "<u4"), ("n_segs", "<u4")])
SEGMENT_DTYPE = np.dtype([("seq_id", "<i4"), ("start0", "<i4"), ("len", "<i4")])


class OrcSeq(C.Structure):
    _fields_ = [("raw", C.c_void_p), ("raw_len", C.c_int64), ("line_bases", C.c_int32), ("line_width", C.c_int32),
                ("seq_len", C.c_int32), ("_pad", C.c_int32)]


class OrcInfo(C.Structure):
    _fields_ = [("kmer_length", C.c_int32), ("lut_prefix_length", C.c_int32), ("signature_length", C.c_int32),
                ("counter_size", C.c_int32), ("both_strands", C.c_int32), ("min_count", C.c_int32),
                ("max_count", C.c_int32), ("n_bins", C.c_int32), ("total_kmers", C.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kcf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkcforacle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_norm_table.argtypes = [C.c_int, C.c_void_p]
        L.orc_kmc_open_mem.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
        L.orc_kmc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.orc_kmc_close.argtypes = [C.c_void_p]
        L.orc_kmc_close.restype = None
        L.orc_kmc_get_info.argtypes = [C.c_void_p, C.POINTER(OrcInfo)]
        L.orc_kmc_get_info.restype = None
        L.orc_kmc_count.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_kmc_count.restype = C.c_int32
        L.orc_kmc_signature.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_kmc_signature.restype = C.c_int32
        L.orc_get_sequence.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_windows_fixed.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_windows_fixed.restype = C.c_int64
        L.orc_compute_score.argtypes = [C.c_int32] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_process_window.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_double),
                                         C.POINTER(OrcResult), C.c_void_p]
        L.orc_screen.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                 C.POINTER(C.c_double), C.c_int32, C.c_void_p]
        L.orc_gap_machine.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_gap_machine.restype = None
        _lib = L
    return _lib


def norm_table(L: int) -> np.ndarray:
    out = np.zeros(1 << (2 * L), np.int32)
    rc = lib().orc_norm_table(L, out.ctypes.data)
    if rc:
        raise ValueError(f"orc_norm_table rc={rc}")
    return out


def windows_fixed(seq_len: int, window: int, step: int, k: int):
    n = lib().orc_windows_fixed(seq_len, window, step, k, None, None, 0)
    if n < 0:
        raise ValueError(f"orc_windows_fixed rc={n}")
    s = np.zeros(n, np.int32)
    e = np.zeros(n, np.int32)
    lib().orc_windows_fixed(seq_len, window, step, k, s.ctypes.data, e.ctypes.data, n)
    return s, e


def gap_machine(hits, k: int):
    h = np.ascontiguousarray(hits, np.uint8)
    out = np.zeros(6, np.int32)
    lib().orc_gap_machine(h.ctypes.data, h.size, k, out.ctypes.data)
    return tuple(int(x) for x in out)


def compute_score(obs, total, eff, inner, left, right, w=(0.3, 0.3, 0.4)):
    ww = (C.c_double * 3)(*w)
    s = C.c_double()
    rc = lib().orc_compute_score(obs, total, eff, inner, left, right, ww, C.byref(s))
    return rc, s.value


def get_sequence(raw: np.ndarray, line_bases: int, line_width: int, seq_len: int, start: int, length: int):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros(max(length, 1), np.uint8)
    rc = lib().orc_get_sequence(raw.ctypes.data, raw.size, line_bases, line_width, seq_len, start, length, out.ctypes.data)
    return rc, out[:max(length, 0)].tobytes()


class OracleKMC:
    """the reference's KMC object (D/KMC.java) restated on the CPU."""

    def __init__(self, pre: np.ndarray | None = None, suf: np.ndarray | None = None, prefix: str | None = None):
        self._h = C.c_void_p()
        if prefix is not None:
            rc = lib().orc_kmc_open(prefix.encode(), C.byref(self._h))
        else:
            self._pre = np.ascontiguousarray(pre, np.uint8)
            self._suf = np.ascontiguousarray(suf, np.uint8)  # borrowed by the C side
            rc = lib().orc_kmc_open_mem(self._pre.ctypes.data, self._pre.size, self._suf.ctypes.data, self._suf.size,
                                        0, C.byref(self._h))
        if rc:
            raise ValueError(f"oracle KMC open failed rc={rc}")
        info = OrcInfo()
        lib().orc_kmc_get_info(self._h, C.byref(info))
        self.info = info
        self.k = info.kmer_length

    def close(self):
        if self._h:
            lib().orc_kmc_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def count(self, kmer: str) -> int:
        assert len(kmer) == self.k
        return lib().orc_kmc_count(self._h, kmer.encode())

    def signature(self, kmer: str) -> int:
        return lib().orc_kmc_signature(self._h, kmer.encode())

    def process_window(self, seq: bytes, min_count: int = 1, w=(0.3, 0.3, 0.4), want_counts: bool = False):
        ww = (C.c_double * 3)(*w)
        res = OrcResult()
        counts = np.zeros(max(len(seq), 1), np.int32) if want_counts else None
        rc = lib().orc_process_window(self._h, seq, len(seq), min_count, ww, C.byref(res),
                                      counts.ctypes.data if want_counts else None)
        if want_counts:
            return rc, res, counts[:res.total_kmers]
        return rc, res

    def screen(self, seqs, wins: np.ndarray, segs: np.ndarray, min_count: int = 1, w=(0.3, 0.3, 0.4), threads: int = 1):
        """seqs: list of (raw uint8 array, line_bases, line_width, seq_len). Returns (rc, results)."""
        arr = (OrcSeq * len(seqs))()
        keep = []
        for i, (raw, lb, lw, sl) in enumerate(seqs):
            raw = np.ascontiguousarray(raw, np.uint8)
            keep.append(raw)
            arr[i] = OrcSeq(raw.ctypes.data, raw.size, lb, lw, sl, 0)
        wins = np.ascontiguousarray(wins, WINDOW_DTYPE)
        segs = np.ascontiguousarray(segs, SEGMENT_DTYPE)
        out = np.zeros(wins.size, RESULT_DTYPE)
        ww = (C.c_double * 3)(*w)
        rc = lib().orc_screen(self._h, C.byref(arr), len(seqs), wins.ctypes.data, wins.size, segs.ctypes.data,
                              min_count, ww, threads, out.ctypes.data)
        return rc, out
