#!/usr/bin/env python
"""bench.py — reference k-mers screened per second by the getVariations hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c2s|c3s|c3st]

One "step" = one pass of the hot path (K3 screening kernel + K4/K5 finalize) over every window of the workload.
The workload of the headline `value` is BASELINE.json configs[1] ("c2": synthetic 900 Mb / 12 chromosome reference,
k=31, 50 kb tiling windows, one KMC database of a SNP/indel-mutated copy at ~8x) at every N.

N = 1      `value`  database, packed reference and window list resident in HBM (kcf_plan_run)
           `e2e`    the same job through the host-buffer call of the C ABI (kcf_screen_sharded over one context): FASTA bytes
                    in pinned host memory -> H2D -> pack -> screen -> rows D2H, every step; database resident
           `e2e_cold`  what ONE getVariations invocation pays below the process: kcf_db_open_mem (ingest) + that call
           `cli`    the getVariations command line on files, process start to exit (BASELINE's second metric)
           `c1`     configs[0] in full: every window compared with the CPU restatement
           `c3`     configs[2] at its stated size: 2.5 Gb reference, >= 2.5e9-record database, gene + transcript windows
           `cpu_baseline`, `roofline` as the contract asks
N > 1      (torchrun) ONE c2 job on N GPUs, strong scaling: the window list is cut into N contiguous ranges (the cut of
           kcf_shard_windows / shard.partition), the database is replicated, no data-path collective; `value` = the job's
           k-mers / max-over-ranks time.  `e2e`: every rank uploads only the stretches of the reference its range touches.
           `placements`: the same kind of job on a >= 3e9-record database (configs[3] / 5) under every placement of the
           table this library has — replicated + sharded windows, scan placement with the table cut in 2 and 4 slices,
           k-mer exchange by NCCL all-to-all — each with per-phase times and the bytes that cross NVLink.

`cpu_baseline` / `--impl reference` time the CPU restatement of the reference algorithm (oracle/, kind "port": the
reference is Java and no JDK exists on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_KMER = 32.375  # one 32-B DRAM sector of the table + 2-bit base + 1-bit validity

WORKLOADS = {
    # name: (n_chrom, chrom_len, window, description); window 0 / -1 = gene / transcript windows from a synthetic GTF
    "c2": (12, 75_000_000, 50_000, "configs[1]: synthetic 900 Mb / 12 chr reference, k=31, 50 kb tiling windows, KMC DB of a mutated copy (~8x)"),
    "c2s": (12, 7_500_000, 50_000, "configs[1] / 10: synthetic 90 Mb / 12 chr"),
    "c1": (1, 10_000_000, 50_000, "configs[0]: synthetic 10 Mb single chromosome, k=31, 50 kb tiling windows"),
    "c3s": (9, 10_000_000, 0, "configs[2] / 28: synthetic 90 Mb / 9 chr reference + GTF (1,500 genes per chromosome), gene windows"),
    "c3st": (9, 10_000_000, -1, "configs[2] / 28: synthetic 90 Mb / 9 chr reference + GTF (1,500 genes per chromosome), transcript windows"),
    "c3": (9, 278_000_000, 0, "configs[2]: synthetic 2.5 Gb / 9 chr reference + GTF (40,500 genes), gene and transcript windows, KMC DB of a mutated copy"),
    "c4s": (12, 250_000_000, 50_000, "configs[3] / 5: synthetic 3.0 Gb / 12 chr reference, 50 kb tiling windows, KMC DB of a mutated copy (~3e9 records)"),
    "c4": (21, 714_000_000, 50_000, "configs[3]: synthetic 15 Gb / 21 chr reference, 50 kb tiling windows, KMC DB of a mutated copy (~1.5e10 records); tools/c4_full.py"),
}
GENES_PER_CHROM = {"c3s": 1500, "c3st": 1500, "c3": 4500}
BIG = {"c3": 4, "c4s": 4, "c4": 16}  # databases built in this many groups of bins (tools/synth.kmc_image_from_genomes_grouped)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, name, fasta, kmc, window, desc):
        self.name, self.fasta, self.kmc, self.window, self.desc = name, fasta, kmc, window, desc
        self.wins = self.segs = self.sids = None

    def seqs(self):
        f = self.fasta
        return [(f.seq_bytes(i), f.line_bases[i], f.line_width[i], f.lengths[i]) for i in range(len(f.names))]


def build_workload(name: str, device, rank: int = 0) -> Workload:
    """synthetic reference FASTA image + KMC image (tools/synth.py), generated on `device`; identical on every rank."""
    import torch
    from tools import synth
    n_chrom, clen, window, desc = WORKLOADS[name]
    t0 = time.time()
    recs, queries = [], []
    for i in range(n_chrom):
        g = synth.random_genome(clen, 2001 + i, device)
        nr = synth.random_intervals(clen, 10, 100, 10_000, 3001 + i)
        lower = synth.random_intervals(clen, 20, 1000, clen // 400, 4001 + i)
        rec = synth.fasta_record(g, f"chr{i + 1:02d}", line=60, lower=lower, n_runs=nr)
        recs.append((f"chr{i + 1:02d}", rec, clen, 60))
        queries.append(synth.mutate(g, 2501 + i))
        del g
    fasta = synth.fasta_image(recs)
    del recs
    log(f"[bench r{rank}] {name} reference: {n_chrom} x {clen} bp, FASTA {fasta.data.size / 1e6:.0f} MB ({time.time() - t0:.1f}s)")
    t1 = time.time()
    if name in BIG:
        kmc = synth.kmc_image_from_genomes_grouped(queries, k=31, P=7, L=9, n_bins=512, counter_size=1, coverage=8.0, seed=77, groups=BIG[name])
    else:
        kmc = synth.kmc_image_from_genomes(queries, k=31, P=7, L=9, n_bins=512, counter_size=1, coverage=8.0, seed=77)
    del queries
    if device != "cpu":
        torch.cuda.empty_cache()
    log(f"[bench r{rank}] {name} KMC image: {kmc.total} records, .kmc_suf {kmc.suf.size / 1e9:.2f} GB, .kmc_pre {kmc.pre.size / 1e6:.0f} MB ({time.time() - t1:.1f}s)")
    wl = Workload(name, fasta, kmc, window, desc)
    if window > 0:
        from kcftools_b200.api import fixed_windows
        wl.wins, wl.segs, wl.starts, wl.ends, wl.sids = fixed_windows(fasta.lengths, window, 0, 31)
    else:
        wl.wins, wl.segs = gtf_windows(fasta, "gene" if window == 0 else "transcript", GENES_PER_CHROM[name])
    return wl


def shared_workload(name: str, device, rank: int, world: int, dist) -> tuple[Workload, str | None]:
    """N > 1, large workloads: rank 0 builds the images once and the other ranks map them from tmpfs — 8 ranks holding a 21 GB
    .kmc_suf each (twice while it is assembled) would take a third of a terabyte of host memory.  Falls back to every rank
    building its own copy when /dev/shm is too small.  Returns (workload, directory to remove at the end or None)."""
    from tools import synth
    n_chrom, clen, window, desc = WORKLOADS[name]
    need = int(n_chrom * clen * 1.02) + int(n_chrom * clen * 8.5)  # FASTA + ~1 record of 8 bytes per base
    ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need + (8 << 30)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    if not flag[0]:
        return build_workload(name, device, rank), None
    d = f"/dev/shm/kcfbench_{os.environ.get('MASTER_PORT', '0')}_{name}"
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        w = build_workload(name, device, rank)
        np.save(os.path.join(d, "fasta.npy"), w.fasta.data)
        np.save(os.path.join(d, "pre.npy"), w.kmc.pre)
        np.save(os.path.join(d, "suf.npy"), w.kmc.suf)
        json.dump({"names": w.fasta.names, "lengths": w.fasta.lengths, "offsets": w.fasta.offsets, "line_bases": w.fasta.line_bases,
                   "line_width": w.fasta.line_width, "k": w.kmc.k, "P": w.kmc.P, "L": w.kmc.L, "n_bins": w.kmc.n_bins, "counter_size": w.kmc.counter_size,
                   "total": w.kmc.total, "both_strands": w.kmc.both_strands}, open(os.path.join(d, "meta.json"), "w"))
        del w
    dist.barrier()
    m = json.load(open(os.path.join(d, "meta.json")))
    fasta = synth.FastaImage(np.load(os.path.join(d, "fasta.npy"), mmap_mode="r"), m["names"], m["lengths"], m["offsets"], m["line_bases"], m["line_width"])
    kmc = synth.KmcImage(pre=np.load(os.path.join(d, "pre.npy"), mmap_mode="r"), suf=np.load(os.path.join(d, "suf.npy"), mmap_mode="r"), k=m["k"], P=m["P"],
                         L=m["L"], n_bins=m["n_bins"], counter_size=m["counter_size"], total=m["total"], both_strands=m["both_strands"])
    w = Workload(name, fasta, kmc, window, desc)
    from kcftools_b200.api import fixed_windows
    w.wins, w.segs, w.starts, w.ends, w.sids = fixed_windows(fasta.lengths, window, 0, 31)
    log(f"[bench r{rank}] {name}: images mapped from {d}")
    return w, (d if rank == 0 else None)


def gtf_windows(fasta, feature: str, genes_per_chrom: int, seed: int = 31337):
    """gene / transcript window and segment arrays from the C++ host (kcftools_b200/host `_windows` hook: the product's own GTF
    logic) for a synthetic GTF over the workload's chromosomes.  The hook needs the sequence NAMES and LENGTHS only, so it gets
    the workload's .faidx next to an empty placeholder FASTA instead of gigabytes of bases."""
    from tools import synth
    from kcftools_b200._lib import SEGMENT_DTYPE, WINDOW_DTYPE
    cli = cli_path()
    d = tempfile.mkdtemp(prefix="kcfbench")
    try:
        fa, gtf = os.path.join(d, "ref.fa"), os.path.join(d, "ann.gtf")
        open(fa, "w").write(">placeholder\n")
        time.sleep(0.01)
        with open(fa + ".faidx", "w") as f:  # name, length, offset, lineBases, lineWidth (FastaIndex.java:239-299)
            for i, n in enumerate(fasta.names):
                f.write(f"{n}\t{fasta.lengths[i]}\t{fasta.offsets[i]}\t{fasta.line_bases[i]}\t{fasta.line_width[i]}\n")
        os.utime(fa + ".faidx", (time.time() + 5, time.time() + 5))  # newer than the FASTA: the index is used as it is
        open(gtf, "w").write(synth.synthetic_gtf(list(zip(fasta.names, fasta.lengths)), genes_per_chrom, seed, max_tx=3, max_exons=12))
        out = subprocess.run([cli, "_windows", "-r", fa, "-f", feature, "-g", gtf, "--kmer-size", "31"], capture_output=True, text=True, check=True).stdout
    finally:
        shutil.rmtree(d, ignore_errors=True)
    wl, sl = [], []
    for line in out.split("\n"):
        if not line.startswith("W\t"):
            continue
        f = line.split("\t")
        wl.append((len(sl), len(f) - 6))
        sl.extend(tuple(int(x) for x in t.split(":")) for t in f[6:])
    return np.array(wl, dtype=WINDOW_DTYPE), np.array(sl, dtype=SEGMENT_DTYPE)


def cli_path():
    cli = os.path.join(ROOT, "kcftools_b200", "host", "kcftools_b200")
    if not os.path.exists(cli):
        subprocess.check_call(["make", "-C", os.path.dirname(cli)], stdout=subprocess.DEVNULL)
    return cli


def cpu_leg(wl: Workload, n_windows: int, threads: int, first: int = 0):
    """time the CPU restatement (oracle, kind=port) on windows [first, first + n_windows); returns (kmers, seconds, rows)."""
    from oracle import binding as ob
    odb = ob.OracleKMC(wl.kmc.pre, wl.kmc.suf)
    w = wl.wins[first:first + n_windows].copy()
    t0 = time.perf_counter()
    rc, res = odb.screen(wl.seqs(), w, wl.segs, min_count=1, threads=threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    odb.close()
    return int(res["total_kmers"].sum()), dt, res


INT_FIELDS = ("total_kmers", "eff_len", "obs", "variations", "inner", "left", "right", "kmer_count_sum")


def rows_equal(a, b) -> bool:
    return a.size == b.size and all((a[f] == b[f]).all() for f in INT_FIELDS) and bool(np.all(np.abs(a["score"] - b["score"]) <= 1e-9 * np.abs(b["score"])))


_STDOUT_FD = None


def quiet_stdout():
    """everything libraries print to fd 1 (NCCL's version banner, ...) goes to stderr: stdout carries the one JSON line"""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def base_config(wl: Workload, placement: str) -> dict:
    return {"workload": f"{wl.name}: {wl.desc}", "k": 31, "window": wl.window, "windows": int(wl.wins.size),
            "db_records": int(wl.kmc.total), "reference_bp": int(sum(wl.fasta.lengths)), "db_placement": placement,
            "l2_policy": "inputs_exceed_l2 (hash table >> 126 MB L2; no flush needed)"}


# ---------------------------------------------------------------------------------------------------------------------
# legs shared by the single- and the multi-GPU run
# ---------------------------------------------------------------------------------------------------------------------
def pin_sequences(ctx, wl: Workload):
    """pinned host copies of the FASTA bytes of every sequence (what a host that wants full-rate copies hands the library)"""
    out = []
    for (raw, lb, lw, sl) in wl.seqs():
        pb = ctx.pinned(raw.size)
        pb[:] = raw
        out.append((pb, lb, lw, sl))
    return out


def timed_resident(torch, ctx, stream, step, steps: int, warmup: int, barrier):
    """W untimed + K timed steps, CUDA events on the library's stream, barrier + synchronize on both sides"""
    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    barrier()
    return ev0.elapsed_time(ev1)


def e2e_leg(torch, ctx, db, pinned, wins, segs, steps: int, barrier, gather=None):
    """the job through the host-buffer call (kcf_screen_sharded, one context): every step uploads the stretches of the reference
    the windows touch, screens them behind the uploads, and brings the rows back.  Wall clock per step, copies inside."""
    from kcftools_b200.api import host_seqs, screen_sharded
    hs = host_seqs(pinned)
    ms, out = [], None
    for it in range(steps + 1):
        barrier()
        t1 = time.perf_counter()
        out = screen_sharded([ctx], [db], hs, wins, segs)
        if gather is not None:
            out = gather(out)
        dt = (time.perf_counter() - t1) * 1e3
        if it > 0:  # the first pass allocates the library's staging buffers
            ms.append(dt)
    return ms, out


def h2d_floor(torch, device, pinned, lo_hi=None):
    """the floor of the e2e leg on this box: the same pinned FASTA bytes copied to the device and nothing else"""
    try:
        srcs = [torch.from_numpy(p[0]) for p in pinned]
        if lo_hi is not None:  # the byte ranges a rank's shard uploads
            srcs = [s[a:b] for s, (a, b) in zip(srcs, lo_hi) if b > a]
        dst = [torch.empty(s.numel(), dtype=torch.uint8, device=device) for s in srcs]
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for d_, s_ in zip(dst, srcs):
                d_.copy_(s_, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t1) * 1e3
            best = dt if best is None else min(best, dt)
        return best, int(sum(s.numel() for s in srcs))
    except Exception as e:  # measurement helper only
        log(f"[bench] h2d floor measurement failed: {e}")
        return None, 0


def shard_byte_ranges(wl: Workload, wins, segs):
    """per sequence the FASTA byte range [a, b) the windows touch (what kcf_screen_sharded uploads, up to line rounding)"""
    out = []
    sid = segs["seq_id"]
    for i in range(len(wl.fasta.names)):
        m = sid == i
        if not m.any():
            out.append((0, 0))
            continue
        lb, lw = wl.fasta.line_bases[i], wl.fasta.line_width[i]
        lo = int(segs["start0"][m].min()) // lb * lw
        hi = (int((segs["start0"][m].astype(np.int64) + segs["len"][m]).max()) - 1) // lb * lw + lw
        out.append((lo, min(hi, wl.fasta.seq_bytes(i).size)))
    return out


# ---------------------------------------------------------------------------------------------------------------------
def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("KCF_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-windows", type=int, default=0, help="windows in the CPU baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the getVariations command-line leg (writes the workload to a temporary directory)")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0] leg (every window against the CPU restatement)")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] leg (2.5 Gb reference, 2.5e9-record database, gene + transcript windows)")
    ap.add_argument("--no-placements", action="store_true", help="N > 1: skip the placements of the >= 3e9-record database")
    ap.add_argument("--placement-workload", default="c4s", choices=sorted(WORKLOADS), help="N > 1: workload of the `placements` records")
    ap.add_argument("--only", default="", help="comma list of legs to keep (resident,e2e,cold,cli,c1,c3,cpu,rand,placements): the others are skipped")
    ap.add_argument("--lf", type=float, default=0.0, help="table load factor (0 = library default)")
    ap.add_argument("--m", type=int, default=0, help="minimizer length (0 = automatic)")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)

    def want(leg: str, default: bool = True) -> bool:
        return (leg in only) if only else default

    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        if args.impl == "reference" and rank != 0:
            return 0  # the CPU arm runs on rank 0 alone
        if args.impl == "ours":
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line and nothing else
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    have_gpu = torch.cuda.is_available()
    if args.impl == "ours" and not have_gpu:
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    device = f"cuda:{local_rank}" if have_gpu else "cpu"
    if have_gpu:
        torch.cuda.set_device(local_rank)
    host_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    wl = build_workload(args.workload, device, rank)
    n_wins = wl.wins.size
    config = base_config(wl, "replicated")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        threads = host_cores
        # bounded sample per step: 50 windows (2.5e6 k-mers) per host core so K+W steps end within minutes
        nwin = args.cpu_windows or min(n_wins, 50 * threads)
        times, kmers = [], 0
        for it in range(args.warmup + args.steps):
            kmers, dt, _ = cpu_leg(wl, nwin, threads)
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        v = kmers / (ms * 1e-3)
        line = {"impl": "reference", "metric": "ref k-mers screened/s", "value": v, "unit": "kmers/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "kmers/s", "cores": threads, "kind": "port",
                                 "sample": f"first {nwin} windows ({kmers} k-mers) of the workload per step; CPU restatement of the reference algorithm "
                                           "(oracle/kcf_oracle.c, pthreads over windows like GetVariants.java:129-151); the Java reference cannot run (no JDK)"},
                "e2e": {"value": v, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    from kcftools_b200 import shard
    from kcftools_b200.api import Context, KMC
    ctx = Context(local_rank)
    if args.lf > 0:
        ctx.set_load_factor(args.lf)
    if args.m > 0:
        ctx.set_minimizer_length(args.m)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    warmup = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: int) -> int:
        if dist is None:
            return x
        t = torch.tensor([x], device=device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    pinned = pin_sequences(ctx, wl)
    line = {}

    # ---- cold: ONE invocation's work below the process = open the database (ingest) + screen from host buffers.
    # Taken twice: the very first open of the process also pays the lazy loading of the kernels and the first pinned
    # allocations; the second is the steady cost of "database not resident".
    cold = None
    t0 = time.time()
    if world == 1 and want("cold"):
        from kcftools_b200.api import screen_sharded
        cold = {"runs": []}
        for it in range(2):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            db_ = KMC(ctx, pre=wl.kmc.pre, suf=wl.kmc.suf)
            t2 = time.perf_counter()
            rows_ = screen_sharded([ctx], [db_], pinned, wl.wins, wl.segs)
            t3 = time.perf_counter()
            cold["runs"].append({"seconds": t3 - t1, "db_open_s": t2 - t1, "screen_s": t3 - t2, "db_load_seconds_library": db_.info.load_seconds,
                                 "db_load_phase_s": {"setup": db_.info.load_phase_s[0], "stream": db_.info.load_phase_s[1], "drain": db_.info.load_phase_s[2],
                                                     "device_ingest_kernels": db_.info.load_phase_s[3]}})
            cold_rows = rows_
            db_.close()
        k_ = int(cold_rows["total_kmers"].sum())
        best = min(cold["runs"][1:], key=lambda r: r["seconds"])
        cold.update({"seconds": best["seconds"], "value": k_ / best["seconds"], "unit": "kmers/s", "db_open_s": best["db_open_s"], "screen_s": best["screen_s"],
                     "first_call_seconds": cold["runs"][0]["seconds"], "input_bytes": int(wl.kmc.suf.size + wl.kmc.pre.size + sum(p[0].size for p in pinned)),
                     "what": "kcf_db_open_mem (parse .kmc_pre, stream .kmc_suf through pinned staging, device ingest into the line table) + "
                             "kcf_screen_sharded from pinned host FASTA bytes (H2D, pack, screen, rows D2H); context and process already up; "
                             "second of two consecutive cold passes (the first also loads the kernels)"})
    t0 = time.time()
    db = KMC(ctx, pre=wl.kmc.pre, suf=wl.kmc.suf)
    db_load_s = time.time() - t0
    log(f"[bench r{rank}] db resident: {db.info.resident_kmers} records in {db.info.n_buckets} lines "
        f"({db.info.table_bytes / 1e9:.2f} GB, {db.info.table_bytes / max(db.info.resident_kmers, 1):.1f} B/record, stash {db.info.stash_kmers}) in {db_load_s:.3f}s")
    table = {"table_bytes": int(db.info.table_bytes), "table_bytes_per_record": db.info.table_bytes / max(db.info.resident_kmers, 1),
             "lines": int(db.info.n_buckets), "stash_kmers": int(db.info.stash_kmers), "db_records": int(db.info.resident_kmers)}
    checks = {}

    # ---- resident: the job's windows cut over the ranks (one range each), packed reference and plan on the device
    ranges = shard.partition(shard.window_lengths(wl.wins, wl.segs), world)
    w0, w1 = ranges[rank]
    my_wins, my_segs = shard.local_slice(wl.wins, wl.segs, w0, w1)
    ctx.ref_clear()  # the host-buffer calls above left their upload pieces as the context's sequences
    for (pb, lb, lw, sl) in pinned:
        ctx.ref_add(pb, lb, lw, sl)
    plan = ctx.plan(31, my_wins, my_segs)
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    total_ms = timed_resident(torch, ctx, stream, lambda: plan.run(db), args.steps, warmup, barrier)
    kernel_ms, finalize_ms = ctx.last_kernel_ms()  # the LAST step of the timed region: a launch in steady state
    clocks = sampler.stop()
    res = plan.fetch()
    my_kmers = int(res["total_kmers"].sum())
    total_ms = allmax(total_ms)
    job_kmers = allsum(my_kmers)
    ms_per_step = total_ms / args.steps
    value = job_kmers / (ms_per_step * 1e-3)
    if cold is not None:
        assert rows_equal(cold_rows, res), "cold pass rows differ from the resident run"
        cold["rows_match_resident_run"] = True

    # ---- e2e: host buffers through the C ABI, copies inside the timed region; rank r uploads only what range r touches
    gather = None
    if dist is not None:
        def gather(mine):  # the job's rows, in window order, on every rank (48 B per window)
            most = max(b - a for a, b in ranges) * 48
            pad = torch.zeros(most, dtype=torch.uint8, device=device)
            pad[:mine.size * 48] = torch.from_numpy(mine.view(np.uint8)).to(device)
            bufs = [torch.empty(most, dtype=torch.uint8, device=device) for _ in range(world)]
            dist.all_gather(bufs, pad)
            return np.concatenate([bufs[r][:(ranges[r][1] - ranges[r][0]) * 48].cpu().numpy() for r in range(world)]).view(res.dtype)
    e2e_ms, e2e_rows = ([], None)
    if want("e2e") and args.e2e_steps > 0:
        e2e_ms, e2e_rows = e2e_leg(torch, ctx, db, pinned, my_wins, my_segs, args.e2e_steps, barrier, gather)
    byte_ranges = shard_byte_ranges(wl, my_wins, my_segs)
    h2d = int(sum(b - a for a, b in byte_ranges) + my_wins.nbytes + my_segs.nbytes)
    d2h = int(my_wins.size * 48)
    e2e_step = allmax(float(np.mean(e2e_ms))) if e2e_ms else None
    e2e_value = job_kmers / (e2e_step * 1e-3) if e2e_step else None
    floor_ms, floor_bytes = h2d_floor(torch, device, pinned, byte_ranges) if e2e_ms else (None, 0)
    if floor_ms is not None and dist is not None:
        barrier()
        floor_ms = allmax(h2d_floor(torch, device, pinned, byte_ranges)[0])  # all ranks copying at once: the host's PCIe fabric is shared
    all_rows = gather(res) if dist is not None else res  # the whole job's rows, in window order
    if e2e_rows is not None:
        assert rows_equal(e2e_rows, all_rows), "e2e rows differ from the resident run"
        checks["e2e_rows_equal_resident"] = True
    if dist is not None and rank == 0:
        # strong scaling moves no result: rank 0 screens the WHOLE job alone and compares every row
        ctx.ref_clear()
        for (pb, lb, lw, sl) in pinned:
            ctx.ref_add(pb, lb, lw, sl)
        whole = ctx.screen(db, wl.wins, wl.segs)
        assert rows_equal(all_rows, whole), "sharded rows differ from the single-GPU run of the same job"
        checks["sharded_rows_equal_single_gpu"] = True

    # ---- roofline of the dominant kernel (kcf_screen_kernel), rank 0's launch
    peak, peak_src = measured_peaks()
    achieved = my_kmers * ALGO_BYTES_PER_KMER / (kernel_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "kernel": "kcf_screen_kernel", "kernel_ms": kernel_ms, "finalize_ms": finalize_ms,
            "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kmers_per_launch": my_kmers,
            "kernel_ms_how": "CUDA events around kcf_screen_kernel on the library's own stream (kcf_last_kernel_ms), read for the last step of the "
                             "timed region; the region is K back-to-back steps timed by events on the same stream, ms_per_step = kernel_ms + "
                             "finalize_ms + launch gaps"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    tj = None
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            if tj.get("workload") == args.workload and world == 1:
                roof["traffic"] = tj["dram_bytes_per_launch"]
            else:
                tj = None
        except Exception:
            tj = None
    if rank == 0 and want("rand"):
        try:
            free_b = torch.cuda.mem_get_info(local_rank)[0]
            buf = min(64 << 30, max(1 << 30, (free_b - (8 << 30)) // (1 << 30) * (1 << 30)))
            rnd = ctx.random_sector_gbps(buf, 1 << 28, 5)
            roof["rand_peak"] = rnd
            roof["rand_frac"] = achieved / rnd
            roof["rand_peak_how"] = f"2^28 independent 32-B loads at uniformly random sector addresses of a {buf >> 30} GiB buffer, best of 5, same process"
            lr = ctx.random_line_rate(buf, 1 << 28, 5)
            roof["rand_line_rate"] = lr
            roof["rand_line_rate_how"] = (f"2^28 random 128-B lines of a {buf >> 30} GiB buffer, each asked for by ONE coalesced request of 4 lanes (the table's "
                                          "access pattern), best of 5, same process")
            if tj is not None and tj.get("dram_read_bytes_per_launch"):
                lines = tj["dram_read_bytes_per_launch"] / 128.0
                roof["line_frac"] = lines / (kernel_ms * 1e-3) / lr
                roof["line_frac_how"] = ("DRAM line reads per launch (ncu dram__bytes_read.sum / 128, profiles/traffic_bytes_per_launch.json) / kernel time, "
                                         "over the measured random-line rate: the memory-side fraction against a bound that is one")
        except Exception as e:  # measurement helper only
            roof["rand_peak"] = None
            log(f"[bench] random access microbenchmarks failed: {e}")

    line = {"metric": "ref k-mers screened/s", "value": value, "unit": "kmers/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "kmers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_step, "ms_each_step": [round(x, 3) for x in e2e_ms], "h2d_copy_alone_ms": floor_ms, "h2d_copy_alone_bytes": floor_bytes,
                    "what": "kcf_screen_sharded (one context per rank) on this rank's window range: line-aligned pieces of the pinned FASTA bytes "
                            "-> H2D -> pack, the windows ending in a piece planned and screened behind it, rows D2H"
                            + ("; rows all-gathered over the ranks inside the timed region; h2d_* are rank 0's, h2d_copy_alone_ms = all ranks copying "
                               "their ranges at once, max over ranks" if world > 1 else "") + "; database resident (db_load_s)"},
            "gpu_launches": int(args.steps * plan.kernels_per_run),
            "roofline": roof, "clocks": clocks, "db_load_s": db_load_s, "db_load_phase_s": [round(x, 4) for x in db.info.load_phase_s], "table": table, "checks": checks,
            "kmers_per_step": job_kmers, "kmers_per_step_this_rank": my_kmers, "obs_fraction": float(res["obs"].sum() / max(my_kmers, 1))}
    if cold is not None:
        line["e2e_cold"] = cold

    # ---- getVariations wall time through the command line, files to KCF (BASELINE's second metric)
    if rank == 0 and world == 1 and wl.window > 0 and want("cli", not args.no_cli):
        line["cli"] = cli_leg(wl, res, local_rank)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    if rank == 0 and world == 1 and want("cpu", not args.no_cpu_baseline):
        nwin = args.cpu_windows or min(n_wins, 250 * host_cores)
        kmers, dt, cres = cpu_leg(wl, nwin, host_cores)
        line["cpu_baseline"] = {"value": kmers / dt, "unit": "kmers/s", "cores": host_cores, "kind": "port",
                                "sample": f"first {nwin} windows ({kmers} k-mers) of the same workload, {dt:.1f} s; CPU restatement of the reference "
                                          "algorithm (oracle/kcf_oracle.c), pthreads over windows; Java reference not runnable (no JDK)",
                                "gpu_matches_cpu_on_sample": rows_equal(res[:nwin], cres)}
    plan.close()
    db.close()
    ctx.ref_clear()

    # ---- configs[0] in full and configs[2] at its stated size (N = 1)
    if rank == 0 and world == 1 and args.workload == "c2":
        if want("c1", not args.no_c1):
            line["c1"] = small_config_leg(torch, ctx, stream, device, host_cores, "c1")
        if want("c3", not args.no_c3):
            try:
                line["c3"] = c3_leg(torch, ctx, stream, device, host_cores, rank)
            except Exception as e:
                line["c3"] = {"error": repr(e)}
                log(f"[bench] c3 leg failed: {e!r}")

    # ---- N > 1: the placements of a table that does not have to fit one GPU
    if world > 1 and want("placements", not args.no_placements):
        del pinned
        wl.fasta = wl.kmc = None  # the headline workload's images are no longer needed: give the host memory back
        import gc
        gc.collect()
        try:
            line["placements"] = placements_leg(torch, dist, ctx, stream, device, rank, world, args, barrier, allmax, allsum)
        except Exception as e:
            line["placements"] = {"error": repr(e)}
            log(f"[bench r{rank}] placements failed: {e!r}")
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        emit(line)  # last: nothing may follow the JSON line on stdout
    return 0


def cli_leg(wl: Workload, res, local_rank: int) -> dict:
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * (wl.kmc.suf.size + wl.fasta.data.size) else None
    d = tempfile.mkdtemp(prefix="kcfcli", dir=base)
    try:
        fa, pref, outp = os.path.join(d, "ref.fa"), os.path.join(d, "sample"), os.path.join(d, "out.kcf")
        wl.fasta.write(fa)
        wl.kmc.write(pref)
        cli = cli_path()
        walls, cli_log = [], []
        for _ in range(2):  # first run also builds the .faidx; both runs read the files from the page cache
            t1 = time.perf_counter()
            pr = subprocess.run([cli, "getVariations", "-r", fa, "-k", pref, "-o", outp, "-s", "bench", "-f", "window", "-w", str(wl.window),
                                 "--device", str(local_rank)], check=True, capture_output=True, text=True)
            walls.append(time.perf_counter() - t1)
            cli_log = [l[11:23] + l[33:] for l in pr.stdout.split("\n") if " - INFO " in l and "CMD" not in l and "--" not in l][-12:]
        rows = [l.split("\t") for l in open(outp) if not l.startswith("#")]
        ok = len(rows) == wl.wins.size and all(int(r[4]) == int(res["total_kmers"][i]) and r[7].split(":")[2] == str(int(res["obs"][i]))
                                                for i, r in enumerate(rows))
        total_kmers = int(res["total_kmers"].sum())
        return {"wall_s_first_run": walls[0], "wall_s": walls[1], "kmers_per_s": total_kmers / walls[1],
                "rows_match_library": bool(ok), "log": cli_log, "input_bytes": int(wl.fasta.data.size + wl.kmc.pre.size + wl.kmc.suf.size),
                "files_on": "tmpfs (/dev/shm)" if base else "the temporary directory's file system",
                "what": "kcftools_b200 getVariations on files in the page cache: CUDA context, mmap + KMC ingest (H2D, table build), .faidx, "
                        "FASTA H2D + pack, screening, KCF text; process start to exit"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def small_config_leg(torch, ctx, stream, device, host_cores: int, name: str) -> dict:
    """configs[0]: every window through the GPU path AND the CPU restatement, all rows compared"""
    from kcftools_b200.api import KMC, screen_sharded
    w = build_workload(name, device)
    pinned = pin_sequences(ctx, w)
    t1 = time.perf_counter()
    db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf)
    rows = screen_sharded([ctx], [db], pinned, w.wins, w.segs)
    cold_s = time.perf_counter() - t1
    ms, rows2 = e2e_leg(torch, ctx, db, pinned, w.wins, w.segs, 5, torch.cuda.synchronize)
    ctx.ref_clear()
    for (pb, lb, lw, sl) in pinned:
        ctx.ref_add(pb, lb, lw, sl)
    plan = ctx.plan(31, w.wins, w.segs)
    total_ms = timed_resident(torch, ctx, stream, lambda: plan.run(db), 20, 3, torch.cuda.synchronize)
    res = plan.fetch()
    kmers, dt, cres = cpu_leg(w, w.wins.size, host_cores)
    out = {"workload": f"{name}: {w.desc}", "windows": int(w.wins.size), "kmers": kmers, "db_records": int(w.kmc.total),
           "value": kmers / (total_ms / 20 * 1e-3), "ms_per_step": total_ms / 20, "e2e_value": kmers / (float(np.mean(ms)) * 1e-3), "e2e_ms_per_step": float(np.mean(ms)),
           "e2e_cold_seconds": cold_s, "cpu_seconds_all_windows": dt, "cpu_value": kmers / dt, "cpu_cores": host_cores,
           "gpu_matches_cpu_all_windows": bool(rows_equal(res, cres) and rows_equal(rows, cres) and rows_equal(rows2, cres)),
           "table_bytes": int(db.info.table_bytes)}
    plan.close()
    db.close()
    ctx.ref_clear()
    return out


def c3_leg(torch, ctx, stream, device, host_cores: int, rank: int) -> dict:
    """configs[2] at its stated size: 2.5 Gb reference, database of a mutated copy (>= 2.5e9 records: the table no longer fits
    at the sparse default density, the loader's ladder picks a denser one), gene AND transcript windows from a synthetic GTF"""
    from kcftools_b200.api import KMC
    from oracle import binding as ob
    torch.cuda.empty_cache()
    w = build_workload("c3", device, rank)
    pinned = pin_sequences(ctx, w)
    t1 = time.perf_counter()
    db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf)
    load_s = time.perf_counter() - t1
    out = {"workload": f"c3: {w.desc}", "db_records": int(w.kmc.total), "reference_bp": int(sum(w.fasta.lengths)), "db_load_s": load_s,
           "table_bytes": int(db.info.table_bytes), "table_bytes_per_record": db.info.table_bytes / max(db.info.resident_kmers, 1),
           "stash_kmers": int(db.info.stash_kmers), "suf_bytes": int(w.kmc.suf.size)}
    log(f"[bench] c3 db resident: {db.info.resident_kmers} records, {db.info.table_bytes / 1e9:.1f} GB "
        f"({out['table_bytes_per_record']:.1f} B/record) in {load_s:.2f}s")
    for (pb, lb, lw, sl) in pinned:
        ctx.ref_add(pb, lb, lw, sl)
    odb = ob.OracleKMC(w.kmc.pre, w.kmc.suf)
    for feature in ("gene", "transcript"):
        wins, segs = (w.wins, w.segs) if feature == "gene" else gtf_windows(w.fasta, "transcript", GENES_PER_CHROM["c3"])
        plan = ctx.plan(31, wins, segs)
        total_ms = timed_resident(torch, ctx, stream, lambda: plan.run(db), 10, 3, torch.cuda.synchronize)
        res = plan.fetch()
        kmers = int(res["total_kmers"].sum())
        plan.close()
        ms, rows = e2e_leg(torch, ctx, db, pinned, wins, segs, 3, torch.cuda.synchronize)
        ctx.ref_clear()
        for (pb, lb, lw, sl) in pinned:  # the e2e call replaced the resident sequences
            ctx.ref_add(pb, lb, lw, sl)
        nchk = min(wins.size, 3000)
        t2 = time.perf_counter()
        rc, cres = odb.screen(w.seqs(), wins[:nchk].copy(), segs, min_count=1, threads=host_cores)
        cdt = time.perf_counter() - t2
        out[feature] = {"windows": int(wins.size), "segments": int(segs.size), "kmers": kmers, "value": kmers / (total_ms / 10 * 1e-3), "ms_per_step": total_ms / 10,
                        "e2e_value": kmers / (float(np.mean(ms)) * 1e-3), "e2e_ms_per_step": float(np.mean(ms)),
                        "obs_fraction": float(res["obs"].sum() / max(kmers, 1)),
                        "cpu_windows_checked": int(nchk), "cpu_value": int(cres["total_kmers"].sum()) / cdt if rc == 0 else None,
                        "gpu_matches_cpu_on_sample": bool(rc == 0 and rows_equal(res[:nchk], cres) and rows_equal(rows[:nchk], cres))}
    odb.close()
    db.close()
    ctx.ref_clear()
    return out


def placements_leg(torch, dist, ctx, stream, device, rank, world, args, barrier, allmax, allsum) -> dict:
    """ONE job on a database of >= 3e9 records under every placement of its table (strong scaling over the N GPUs)"""
    from kcftools_b200 import shard
    from kcftools_b200.api import KMC
    from kcftools_b200.partitioned import screen_partitioned, screen_partitioned_a2a, screen_partitioned_scan
    torch.cuda.empty_cache()
    w, shm_dir = shared_workload(args.placement_workload, device, rank, world, dist)
    seqs = w.seqs()
    out = {"workload": f"{w.name}: {w.desc}", "db_records": int(w.kmc.total), "windows": int(w.wins.size), "reference_bp": int(sum(w.fasta.lengths)),
           "steps": 3, "warmup": 1}
    lengths = shard.window_lengths(w.wins, w.segs)
    ranges = shard.partition(lengths, world)

    def load_reference():
        ctx.ref_clear()
        for (raw, lb, lw, sl) in seqs:
            ctx.ref_add(raw, lb, lw, sl)

    def timed(step):
        r = step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t1 = time.perf_counter()
        ev0.record(stream)
        for _ in range(3):
            r = step()
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t1) * 1e3
        # host-synchronous placements (exchange, scan) are timed by the wall clock between the barriers; the resident one by events
        return r, allmax(wall_ms) / 3, allmax(ev0.elapsed_time(ev1)) / 3

    load_reference()
    baseline_rows = None
    # (1) replicated table, windows cut over the ranks
    try:
        t1 = time.perf_counter()
        db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf)
        load_s = allmax(time.perf_counter() - t1)
        lw_, ls_ = shard.local_slice(w.wins, w.segs, *ranges[rank])
        plan = ctx.plan(31, lw_, ls_)
        _, wall_ms, ev_ms = timed(lambda: plan.run(db))
        rows = plan.fetch()
        kmers = allsum(int(rows["total_kmers"].sum()))
        baseline_rows = rows
        out["job_kmers"] = kmers
        out["job_observed_kmers"] = allsum(int(rows["obs"].sum()))
        out["replicated_sharded"] = {"value": kmers / (ev_ms * 1e-3), "ms_per_step": ev_ms, "db_load_s": load_s, "table_bytes_per_gpu": int(db.info.table_bytes),
                                     "table_bytes_per_record": db.info.table_bytes / max(db.info.resident_kmers, 1), "nvlink_bytes_per_step": 0,
                                     "phases_ms": {"screen": ev_ms}, "what": "whole table on every GPU, windows cut into N ranges, no collective"}
        plan.close()
        db.close()
    except Exception as e:
        out["replicated_sharded"] = {"error": repr(e)}
    # (2) scan placement: table cut in T slices, N / T window shards
    for T in (2, 4):
        key = f"scan_T{T}"
        if world % T:
            continue
        try:
            n_shards = world // T
            groups = [dist.new_group(list(range(s * T, (s + 1) * T))) for s in range(n_shards)]  # every rank creates every group
            part_rank, shard_id, _ = shard.grid_layout(rank, world, T)
            ctx.set_partition(part_rank, T)
            t1 = time.perf_counter()
            db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf, placement=1)
            load_s = allmax(time.perf_counter() - t1)
            rng = shard.partition(lengths, n_shards)[shard_id]
            lw_, ls_ = shard.local_slice(w.wins, w.segs, *rng)
            plan = ctx.plan(31, lw_, ls_)
            phases = {}
            rows, wall_ms, _ = timed(lambda: screen_partitioned_scan(ctx, db, plan, group=groups[shard_id], phases=phases))
            kmers = allsum(int(rows["total_kmers"].sum())) // T
            same = "job_observed_kmers" not in out or (allsum(int(rows["obs"].sum())) // T == out["job_observed_kmers"] and kmers == out["job_kmers"])
            words = plan.n_tiles * 64
            out[key] = {"value": kmers / (wall_ms * 1e-3), "ms_per_step": wall_ms, "db_load_s": load_s, "table_bytes_per_gpu": int(db.info.table_bytes),
                        "table_slices": T, "window_shards": n_shards, "totals_equal_replicated": bool(same),
                        "nvlink_bytes_per_step": int(2 * (T - 1) / T * (words * 4 + plan.n_tiles * 8) * world),
                        "phases_ms": {k_: allmax(v) / max(phases.get("_n", 1), 1) * 1e3 for k_, v in phases.items() if not k_.startswith("_")},
                        "what": "every rank of a group walks the group's windows, probes the k-mers whose home line lies in its slice; hit bitmaps (1 bit / "
                                "position) and count sums all-reduced inside the group (ring estimate for the NVLink bytes: 2 (T-1)/T x payload per rank)"}
            plan.close()
            db.close()
        except Exception as e:
            out[key] = {"error": repr(e)}
    # (3) k-mer exchange: table cut in N slices, windows cut in N ranges, k-mers routed to their owners by all-to-all
    try:
        ctx.set_partition(rank, world)
        t1 = time.perf_counter()
        db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf, placement=1)
        load_s = allmax(time.perf_counter() - t1)
        lw_, ls_ = shard.local_slice(w.wins, w.segs, *ranges[rank])
        plan = ctx.plan(31, lw_, ls_)
        rows, wall_ms, _ = timed(lambda: screen_partitioned(ctx, db, plan))
        phases = {}
        screen_partitioned(ctx, db, plan, phases=phases)
        kmers = allsum(int(rows["total_kmers"].sum()))
        ok = baseline_rows is None or rows_equal(rows, baseline_rows)
        out["a2a"] = {"value": kmers / (wall_ms * 1e-3), "ms_per_step": wall_ms, "db_load_s": load_s, "table_bytes_per_gpu": int(db.info.table_bytes),
                      "rows_equal_replicated": bool(ok),
                      "nvlink_bytes_per_step": int(allsum(int(phases.get("_bytes_out", 0)) + int(phases.get("_bytes_back", 0))) / max(phases.get("_n", 1), 1)),
                      "phases_ms": {k_: allmax(v) / max(phases.get("_n", 1), 1) * 1e3 for k_, v in phases.items() if not k_.startswith("_")},
                      "phases_note": "per-phase times come from a run with a device synchronisation after every phase; ms_per_step from one without",
                      "runs_per_step": int(allsum(int(phases.get("_runs", 0))) / max(phases.get("_n", 1), 1)),
                      "what": "table cut by home line in N slices, windows in N ranges; exchange over peer memory (CUDA IPC + NVLink): per batch the "
                              "screening kernel appends every RUN of k-mers sharing a home line (16 B for up to 11 k-mers: the bases they span) to "
                              "its owner's inbox, barrier, owners fetch the run's line once, look the k-mers up and store a 16-B slot of counts into "
                              "the requester's workspace, barrier, fold; nothing crosses the host inside a step"}
        # the same exchange as NCCL all-to-all collectives over caller-owned buffers (round 1), for comparison
        phases2 = {}
        rows2, wall2, _ = timed(lambda: screen_partitioned_a2a(ctx, db, plan, phases=phases2))
        out["a2a_nccl_collectives"] = {"value": kmers / (wall2 * 1e-3), "ms_per_step": wall2, "rows_equal_replicated": bool(baseline_rows is None or rows_equal(rows2, baseline_rows)),
                                       "nvlink_bytes_per_step": int(allsum(int(phases2.get("_bytes_out", 0)) + int(phases2.get("_bytes_back", 0))) / max(phases2.get("_n", 1), 1)),
                                       "phases_ms": {k_: allmax(v) / max(phases2.get("_n", 1), 1) * 1e3 for k_, v in phases2.items() if not k_.startswith("_")}}
        for x_ in getattr(plan, "_exchange", None) or []:
            x_.close()
        plan.close()
        db.close()
    except Exception as e:
        out["a2a"] = {"error": repr(e)}
    ctx.set_partition(0, 1)
    ctx.ref_clear()
    barrier()
    if shm_dir:
        shutil.rmtree(shm_dir, ignore_errors=True)
    return out


if __name__ == "__main__":
    sys.exit(main())
