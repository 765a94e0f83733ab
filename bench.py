#!/usr/bin/env python
"""bench.py — reference k-mers screened per second by the getVariations hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c2s] [--impl ours|reference]

One "step" = one pass of the hot path (kcf_plan_run: K3 screening kernel + K4/K5 finalize) over every
window of the workload.  Workload at N=1 = BASELINE.json configs[1] ("c2": synthetic 900 Mb / 12 chromosome
reference, k=31, 50 kb tiling windows, one KMC database of a SNP/indel-mutated copy at ~8x).  Under torchrun
each rank screens its own c2-sized shard against a replicated database (weak scaling, no data-path
collective).

JSON keys follow the driver contract: `value` is device-resident throughput (database, 2-bit reference and
window list already in HBM), `e2e` goes through the C ABI with host buffers: FASTA bytes in pinned host
memory -> kcf_ref_add (H2D + pack) -> kcf_screen (descriptor H2D, kernels, result D2H).  `roofline` is the
screening kernel against the measured HBM copy bandwidth with 32.375 algorithmic bytes per k-mer
(DESIGN.md §6); `roofline.rand_*` is the same against a random 32-byte-sector gather microbenchmark run on
the same GPU in the same process.  `cpu_baseline` / `--impl reference` time the CPU restatement of the
reference algorithm (oracle/, kind "port": the reference is Java and no JDK exists on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_KMER = 32.375  # one 32-B DRAM sector of the table + 2-bit base + 1-bit validity

WORKLOADS = {
    # name: (n_chrom, chrom_len, window, description)
    "c2": (12, 75_000_000, 50_000, "configs[1]: synthetic 900 Mb / 12 chr reference, k=31, 50 kb tiling windows, KMC DB of a mutated copy (~8x)"),
    "c2s": (12, 7_500_000, 50_000, "configs[1] / 10: synthetic 90 Mb / 12 chr"),
    "c1": (1, 10_000_000, 50_000, "configs[0]: synthetic 10 Mb single chromosome, k=31, 50 kb tiling windows"),
    # gene / transcript windows (configs[2] scaled down): window size 0 = windows come from a synthetic GTF
    "c3s": (9, 10_000_000, 0, "configs[2] / 28: synthetic 90 Mb / 9 chr reference + GTF (1,500 genes per chromosome), gene windows"),
    "c3st": (9, 10_000_000, -1, "configs[2] / 28: synthetic 90 Mb / 9 chr reference + GTF (1,500 genes per chromosome), transcript windows"),
}


def gtf_windows(fasta, feature: str, seed: int = 31337):
    """gene / transcript window and segment arrays from the C++ host (kcftools_b200/host `_windows` hook: the
    product's own GTF logic), for a synthetic GTF over the workload's chromosomes"""
    from tools import synth
    from kcftools_b200._lib import SEGMENT_DTYPE, WINDOW_DTYPE
    cli = os.path.join(ROOT, "kcftools_b200", "host", "kcftools_b200")
    if not os.path.exists(cli):
        subprocess.check_call(["make", "-C", os.path.dirname(cli)], stdout=subprocess.DEVNULL)
    d = tempfile.mkdtemp(prefix="kcfbench")
    fa, gtf = os.path.join(d, "ref.fa"), os.path.join(d, "ann.gtf")
    fasta.write(fa)
    open(gtf, "w").write(synth.synthetic_gtf(list(zip(fasta.names, fasta.lengths)), 1500, seed, max_tx=3, max_exons=12))
    out = subprocess.run([cli, "_windows", "-r", fa, "-f", feature, "-g", gtf, "--kmer-size", "31"], capture_output=True, text=True, check=True).stdout
    wl, sl = [], []
    for line in out.split("\n"):
        if not line.startswith("W\t"):
            continue
        f = line.split("\t")
        wl.append((len(sl), len(f) - 6))
        sl.extend(tuple(int(x) for x in t.split(":")) for t in f[6:])
    for f_ in (fa, fa + ".faidx", gtf):
        os.unlink(f_)
    os.rmdir(d)
    wins = np.array(wl, dtype=WINDOW_DTYPE)
    segs = np.array(sl, dtype=SEGMENT_DTYPE)
    return wins, segs


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


def build_workload(name: str, device, rank: int = 0):
    """synthetic reference FASTA image + KMC image (tools/synth.py), generated on `device`."""
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
    log(f"[bench r{rank}] reference: {n_chrom} x {clen} bp, FASTA {fasta.data.size / 1e6:.0f} MB ({time.time() - t0:.1f}s)")
    t1 = time.time()
    kmc = synth.kmc_image_from_genomes(queries, k=31, P=7, L=9, n_bins=512, counter_size=1, coverage=8.0, seed=77)
    del queries
    if device != "cpu":
        torch.cuda.empty_cache()
    log(f"[bench r{rank}] KMC image: {kmc.total} records, .kmc_suf {kmc.suf.size / 1e9:.2f} GB, .kmc_pre {kmc.pre.size / 1e6:.0f} MB ({time.time() - t1:.1f}s)")
    return fasta, kmc, window, desc


def cpu_leg(fasta, kmc, wins, segs, n_windows: int, threads: int):
    """time the CPU restatement (oracle, kind=port) on the first n_windows windows; returns (kmers, seconds)."""
    from oracle import binding as ob
    odb = ob.OracleKMC(kmc.pre, kmc.suf)
    seqs = [(fasta.seq_bytes(i), fasta.line_bases[i], fasta.line_width[i], fasta.lengths[i]) for i in range(len(fasta.names))]
    w = wins[:n_windows].copy()
    s = segs
    t0 = time.perf_counter()
    rc, res = odb.screen(seqs, w, s, min_count=1, threads=threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    odb.close()
    return int(res["total_kmers"].sum()), dt, res


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


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("KCF_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-pieces", type=int, default=1,
                    help="e2e leg: line-aligned pieces per chromosome, uploaded and screened in turn (measured on C2: 18.9 / 23.6 / 30.5 ms per "
                         "step for 1 / 4 / 8 pieces: the per-upload host cost outweighs the shorter tail)")
    ap.add_argument("--cpu-windows", type=int, default=0, help="windows in the CPU baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cli", action="store_true",
                    help="also time the getVariations command line end to end on files (FASTA + KMC database written to a "
                         "temporary directory): BASELINE.json's second metric, wall time vs the host CPU")
    ap.add_argument("--placement", default="replicated", choices=["replicated", "partitioned", "partitioned-scan"],
                    help="partitioned: every rank keeps 1/N of the table, k-mers are routed by NCCL all-to-all (needs --gpus > 1); "
                         "partitioned-scan: same table slices, every rank walks ALL windows and probes what it owns, hit bitmaps are "
                         "all-reduced (total work fixed: strong scaling)")
    ap.add_argument("--table-parts", type=int, default=0,
                    help="partitioned-scan only: slices of the table (default = N); N / table-parts window shards, each screened by a "
                         "group of table-parts GPUs (e.g. --gpus 8 --table-parts 2: half a table per GPU, 4 window shards)")
    ap.add_argument("--lf", type=float, default=0.0, help="table load factor (0 = library default)")
    ap.add_argument("--m", type=int, default=0, help="minimizer length (0 = automatic)")
    args = ap.parse_args()

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

    fasta, kmc, window, desc = build_workload(args.workload, device, rank)
    from kcftools_b200.api import fixed_windows
    if window > 0:
        wins, segs, starts, ends, sids = fixed_windows(fasta.lengths, window, 0, 31)
    else:
        wins, segs = gtf_windows(fasta, "gene" if window == 0 else "transcript")
        sids = segs["seq_id"][wins["first_seg"]]  # every locus of a synthetic gene lies on the gene's chromosome
    n_wins = wins.size
    config = {"workload": f"{args.workload}: {desc}", "k": 31, "window": window, "windows": int(n_wins),
              "db_records": int(kmc.total), "reference_bp": int(sum(fasta.lengths)), "db_placement": "replicated",
              "l2_policy": "inputs_exceed_l2 (hash table >> 126 MB L2; no flush needed)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        threads = host_cores
        # bounded sample per step: 50 windows (2.5e6 k-mers) per host core so K+W steps end within minutes
        nwin = args.cpu_windows or min(n_wins, 50 * threads)
        times, kmers = [], 0
        for it in range(args.warmup + args.steps):
            kmers, dt, _ = cpu_leg(fasta, kmc, wins, segs, nwin, threads)
            if it >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times))
        v = kmers / (ms * 1e-3)
        line = {"impl": "reference", "metric": "ref k-mers screened/s", "value": v, "unit": "kmers/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "kmers/s", "cores": threads, "kind": "port",
                                 "sample": f"first {nwin} windows ({kmers} k-mers) of the workload per step; CPU restatement of the reference algorithm "
                                           "(oracle/kcf_oracle.c, pthreads over windows like GetVariants.java:129-151); the Java reference cannot run (no JDK)"},
                "e2e": {"value": v, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    from kcftools_b200.api import Context, KMC
    ctx = Context(local_rank)
    if args.lf > 0:
        ctx.set_load_factor(args.lf)
    if args.m > 0:
        ctx.set_minimizer_length(args.m)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    t0 = time.time()
    partitioned = args.placement.startswith("partitioned") and world > 1
    scan = partitioned and args.placement == "partitioned-scan"
    tparts = (args.table_parts or world) if scan else world
    scan_group, scan_shard, n_shards = None, 0, 1
    if scan:
        from kcftools_b200 import shard as _shard
        n_shards = world // tparts
        for s_ in range(n_shards):  # every rank creates every group (torch.distributed rule)
            g_ = dist.new_group(list(range(s_ * tparts, (s_ + 1) * tparts)))
            if s_ == rank // tparts:
                scan_group = g_
        part_rank, scan_shard, _ = _shard.grid_layout(rank, world, tparts)
    if partitioned:
        ctx.set_partition(part_rank if scan else rank, tparts)
        config["db_placement"] = (f"partitioned by home line, 1/{tparts} per GPU, {n_shards} window shard(s); every rank walks its shard's windows and "
                                  "probes the k-mers it owns, hit bitmaps + count sums all-reduced inside the shard's group (NCCL)" if scan else
                                  f"partitioned by home line, 1/{world} per GPU, k-mers routed by NCCL all-to-all")
    db = KMC(ctx, pre=kmc.pre, suf=kmc.suf, placement=1 if partitioned else 0)
    db_load_s = time.time() - t0
    log(f"[bench r{rank}] db resident: {db.info.resident_kmers} records in {db.info.n_buckets} buckets "
        f"({db.info.table_bytes / 1e9:.2f} GB, stash {db.info.stash_kmers}) in {db_load_s:.3f}s")
    # reference sequences: pinned host copies (the e2e leg re-uploads them every step)
    pinned = []
    for i in range(len(fasta.names)):
        raw = fasta.seq_bytes(i)
        pb = ctx.pinned(raw.size)
        pb[:] = raw
        pinned.append(pb)
    for i, pb in enumerate(pinned):
        ctx.ref_add(pb, fasta.line_bases[i], fasta.line_width[i], fasta.lengths[i])
    all_wins = wins
    if scan and n_shards > 1:  # this rank's group screens one contiguous shard of the window list
        rng = _shard.partition(_shard.window_lengths(wins, segs), n_shards)[scan_shard]
        wins, segs = _shard.local_slice(wins, segs, *rng)
    plan = ctx.plan(31, wins, segs)
    ctx.set_profiling(True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if partitioned:
        from kcftools_b200.partitioned import screen_partitioned, screen_partitioned_scan

        def step():
            return screen_partitioned_scan(ctx, db, plan, group=scan_group) if scan else screen_partitioned(ctx, db, plan)
    else:
        def step():
            plan.run(db)
    for _ in range(max(args.warmup, 3)):
        r_ = step()
    res = r_ if partitioned else plan.fetch()
    total_kmers = int(res["total_kmers"].sum())

    sampler = ClockSampler(local_rank)
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ms = []
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    # duration of the screening kernel: the library brackets it with CUDA events on its own stream in every step
    # (profiling on); the pair read here belongs to the LAST step of the timed region, i.e. a launch in steady state
    if partitioned:
        kernel_ms.append((total_ms / args.steps, 0.0))
    else:
        kernel_ms.append(ctx.last_kernel_ms())
    clocks = sampler.stop()
    res2 = step() if partitioned else plan.fetch()
    assert (res2 == res).all(), "results changed between runs"
    if dist is not None:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        k_all = torch.tensor([total_kmers], device=device, dtype=torch.int64)
        dist.all_reduce(k_all, op=dist.ReduceOp.SUM)
        # scan placement: ONE copy of the workload screened by all ranks together (every rank of a group reports its shard)
        job_kmers = int(k_all.item()) // tparts if scan else int(k_all.item())
    else:
        job_kmers = total_kmers
    ms_per_step = total_ms / args.steps
    value = job_kmers / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region.  The host walks the sequences the way
    # GetVariants.java:117-121 does: chromosome i is uploaded (pinned FASTA bytes, asynchronously) and its windows are
    # screened while chromosome i+1 is on the PCIe bus; results are read back at the end of the step.
    e2e_ms = []
    h2d = int(sum(p.size for p in pinned) + wins.nbytes + segs.nbytes)
    d2h = int(n_wins * 48)
    out = res
    # Every chromosome goes up in `--e2e-pieces` line-aligned pieces, each registered as its own sequence (kcf_ref_add_async)
    # with the windows that END in it planned right behind it; a window that straddles a piece boundary is the
    # concatenation of two segments (the ABI's window model, GTF.java:240-244).  The screening of a chromosome then trails
    # its upload by a piece, not by the whole chromosome.
    pieces = max(1, args.e2e_pieces) if window > 0 else 1
    uploads = []  # (pinned view, line_bases, line_width, bases, wins, segs) in upload order
    if not partitioned and args.e2e_steps > 0:
        from kcftools_b200._lib import SEGMENT_DTYPE, WINDOW_DTYPE
        bounds = np.searchsorted(sids, np.arange(len(pinned) + 1))
        gid = 0
        for i, pb in enumerate(pinned):
            n, lb, lw = int(fasta.lengths[i]), int(fasta.line_bases[i]), int(fasta.line_width[i])
            w0, w1 = int(bounds[i]), int(bounds[i + 1])
            if pieces == 1 or window <= 0:
                from kcftools_b200 import shard
                lw_, ls_ = shard.local_slice(wins, segs, w0, w1)
                ls_ = ls_.copy()
                ls_["seq_id"] = gid
                uploads.append((pb, lb, lw, n, lw_, ls_))
                gid += 1
                continue
            cuts = [min(n, (n * j // pieces) // lb * lb) for j in range(pieces)] + [n]
            ws, we = starts[w0:w1].astype(np.int64), ends[w0:w1].astype(np.int64)
            last_piece = np.searchsorted(np.asarray(cuts[1:]), we - 1, side="right")  # piece holding the window's last base
            for j in range(pieces):
                b0, b1 = cuts[j], cuts[j + 1]
                byte0 = b0 // lb * lw
                byte1 = pb.size if j == pieces - 1 else b1 // lb * lw
                sel = np.nonzero(last_piece == j)[0]
                pw = np.zeros(sel.size, WINDOW_DTYPE)
                ps = []
                for t, wi in enumerate(sel):
                    s_, e_ = int(ws[wi]), int(we[wi])
                    first = len(ps)
                    jj = j
                    while jj > 0 and cuts[jj] > s_:
                        jj -= 1  # the window starts in an earlier piece
                    for q in range(jj, j + 1):
                        a_, z_ = max(s_, cuts[q]), min(e_, cuts[q + 1])
                        if z_ > a_:
                            ps.append((gid - (j - q), a_ - cuts[q], z_ - a_))
                    pw[t] = (first, len(ps) - first)
                psa = np.zeros(len(ps), SEGMENT_DTYPE)
                for t, sg in enumerate(ps):
                    psa[t] = sg
                uploads.append((pb[byte0:byte1], lb, lw, b1 - b0, pw, psa))
                gid += 1
    for it in range(args.e2e_steps + 1 if (args.e2e_steps > 0 and not partitioned) else 0):
        barrier()
        t1 = time.perf_counter()
        ctx.ref_clear()
        plans = []
        for (buf, lb, lw, nb_, pw, psa) in uploads:
            ctx.ref_add_async(buf, lb, lw, nb_)
            if pw.size:
                pl = ctx.plan(31, pw, psa)
                pl.run(db)
                plans.append(pl)
        out = np.concatenate([pl.fetch() for pl in plans])
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t1) * 1e3
        for pl in plans:
            pl.close()
        if it > 0:
            e2e_ms.append(dt)
    assert (out == res).all(), "e2e results differ from the resident run"
    # the floor of that leg on this box: the same pinned FASTA bytes copied to the device and nothing else
    h2d_floor_ms = None
    if e2e_ms:
        try:
            dst = [torch.empty(p.size, dtype=torch.uint8, device=device) for p in pinned]
            srcs = [torch.from_numpy(p) for p in pinned]
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                for d_, s_ in zip(dst, srcs):
                    d_.copy_(s_, non_blocking=True)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t1) * 1e3
                best = dt if best is None else min(best, dt)
            h2d_floor_ms = best
            del dst
        except Exception as e:  # measurement helper only
            log(f"[bench] h2d floor measurement failed: {e}")
    e2e_step = float(np.mean(e2e_ms)) if e2e_ms else None
    if dist is not None and e2e_step is not None:
        t = torch.tensor([e2e_step], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_step = float(t.item())
    e2e_value = job_kmers / (e2e_step * 1e-3) if e2e_step else None

    # ---- roofline of the dominant kernel (kcf_screen_kernel)
    peak, peak_src = measured_peaks()
    screen_ms = float(np.mean([a for a, _ in kernel_ms]))
    finalize_ms = float(np.mean([b for _, b in kernel_ms]))
    achieved = total_kmers * ALGO_BYTES_PER_KMER / (screen_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "kernel": "kcf_screen_kernel", "kernel_ms": screen_ms, "finalize_ms": finalize_ms,
            "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER, "kmers_per_launch": total_kmers,
            "kernel_ms_how": "CUDA events around kcf_screen_kernel on the library's own stream (kcf_last_kernel_ms), read for the last step of the "
                             "timed region; the region is K back-to-back steps timed by events on the same stream, ms_per_step = kernel_ms + "
                             "finalize_ms + launch gaps"}
    if rank == 0:
        try:
            rnd = ctx.random_sector_gbps(min(16 << 30, max(1 << 30, 2 * db.info.table_bytes)), 1 << 28, 5)
            roof["rand_peak"] = rnd
            roof["rand_frac"] = achieved / rnd
            roof["rand_peak_how"] = "2^28 independent 32-B loads at uniformly random sector addresses of a 16 GiB buffer, best of 5, same process"
        except Exception as e:  # measurement helper only
            roof["rand_peak"] = None
            log(f"[bench] random sector microbenchmark failed: {e}")
    traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            if tj.get("workload") == args.workload:
                roof["traffic"] = tj["dram_bytes_per_launch"]
        except Exception:
            pass

    line = {"metric": "ref k-mers screened/s", "value": value, "unit": "kmers/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if scan else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "kmers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_step, "ms_each_step": [round(x, 3) for x in e2e_ms], "h2d_copy_alone_ms": h2d_floor_ms, "pieces_per_chromosome": pieces,
                    "what": "per piece of a chromosome: kcf_ref_add_async (pinned FASTA bytes -> H2D -> 2-bit pack) + kcf_plan_create (window "
                            "H2D) + kcf_plan_run; then kcf_plan_fetch (result D2H) of every plan; database resident (loaded once: db_load_s)"},
            "gpu_launches": int(args.steps * plan.kernels_per_run),
            "roofline": roof, "clocks": clocks, "db_load_s": db_load_s,
            "kmers_per_step_per_gpu": total_kmers, "obs_fraction": float(res["obs"].sum() / max(total_kmers, 1))}

    # ---- getVariations wall time through the command line, files to KCF (opt-in: writes the workload to disk)
    if args.cli and rank == 0 and world == 1 and window > 0:
        d = tempfile.mkdtemp(prefix="kcfcli")
        try:
            fa, pref, outp = os.path.join(d, "ref.fa"), os.path.join(d, "sample"), os.path.join(d, "out.kcf")
            fasta.write(fa)
            kmc.write(pref)
            cli = os.path.join(ROOT, "kcftools_b200", "host", "kcftools_b200")
            if not os.path.exists(cli):
                subprocess.check_call(["make", "-C", os.path.dirname(cli)], stdout=subprocess.DEVNULL)
            walls = []
            for _ in range(2):  # first run also builds the .faidx; both runs read the files from the page cache
                t1 = time.perf_counter()
                pr = subprocess.run([cli, "getVariations", "-r", fa, "-k", pref, "-o", outp, "-s", "bench", "-f", "window", "-w", str(window),
                                     "--device", str(local_rank)], check=True, capture_output=True, text=True)
                walls.append(time.perf_counter() - t1)
                cli_log = [l[11:23] + l[33:] for l in pr.stdout.split("\n") if " - INFO " in l and "CMD" not in l and "--" not in l][-12:]
            rows = [l.split("\t") for l in open(outp) if not l.startswith("#")]
            ok = len(rows) == n_wins and all(int(r[4]) == int(res["total_kmers"][i]) and r[7].split(":")[2] == str(int(res["obs"][i]))
                                              for i, r in enumerate(rows))
            line["cli"] = {"wall_s_first_run": walls[0], "wall_s": walls[1], "kmers_per_s": total_kmers / walls[1],
                           "rows_match_library": bool(ok), "log": cli_log, "input_bytes": int(fasta.data.size + kmc.pre.size + kmc.suf.size),
                           "what": "kcftools_b200 getVariations on files in the page cache: mmap + KMC ingest (H2D, table build), .faidx, "
                                   "FASTA H2D + pack, screening, KCF text; process start to exit"}
        finally:
            import shutil
            shutil.rmtree(d, ignore_errors=True)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nwin = args.cpu_windows or min(n_wins, 250 * host_cores)
        kmers, dt, cres = cpu_leg(fasta, kmc, wins, segs, nwin, host_cores)
        same = all((cres[f] == res[:nwin][f]).all() for f in ("total_kmers", "eff_len", "obs", "variations", "inner", "left", "right", "kmer_count_sum"))
        line["cpu_baseline"] = {"value": kmers / dt, "unit": "kmers/s", "cores": host_cores, "kind": "port",
                                "sample": f"first {nwin} windows ({kmers} k-mers) of the same workload, {dt:.1f} s; CPU restatement of the reference "
                                          "algorithm (oracle/kcf_oracle.c), pthreads over windows; Java reference not runnable (no JDK)",
                                "gpu_matches_cpu_on_sample": bool(same)}
    plan.close()
    db.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        emit(line)  # last: nothing may follow the JSON line on stdout
    return 0


if __name__ == "__main__":
    sys.exit(main())
