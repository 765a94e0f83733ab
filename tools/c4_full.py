"""configs[3] at its stated size on the 8 GPUs of one box (measurement tool, run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c4_full.py

A 15 Gb / 21-chromosome synthetic reference (300,069 tiling windows of 50 kb, 1.5e10 k-mers) against the KMC image of a
mutated copy (~1.5e10 records, ~105 GB of .kmc_suf, built once by rank 0 and mapped from tmpfs by the others).  The table
does not fit one GPU at any density, so the job runs under the partitioned placements only: the k-mer
exchange over peer memory (table cut in 8 slices) and the scan placement with 4 slices x 2 window shards.  No oracle can
follow at this size: the two placements have to agree with each other row by row (gathered per window range), and the
windows inside... are checked for exact TOTAL_KMERS / EFFLEN against the window geometry.  Rank 0 prints one JSON line.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from kcftools_b200 import shard
    from kcftools_b200.api import Context, KMC
    from kcftools_b200.partitioned import screen_partitioned, screen_partitioned_scan
    bench.quiet_stdout()
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    t0 = time.time()
    w, shm_dir = bench.shared_workload(name, device, rank, world, dist)
    build_s = time.time() - t0
    out = {"workload": f"{name}: {w.desc}", "db_records": int(w.kmc.total), "suf_bytes": int(w.kmc.suf.size), "windows": int(w.wins.size),
           "reference_bp": int(sum(w.fasta.lengths)), "n_gpus": world, "build_s": build_s}
    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], device=device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    t0 = time.time()
    for (raw, lb, lw, sl) in w.seqs():
        ctx.ref_add(raw, lb, lw, sl)
    out["reference_upload_s"] = allmax(time.time() - t0)
    lengths = shard.window_lengths(w.wins, w.segs)

    def timed(step, n=2):
        r = step()
        barrier()
        t1 = time.perf_counter()
        for _ in range(n):
            r = step()
        barrier()
        return r, allmax((time.perf_counter() - t1) * 1e3) / n

    # ---- k-mer exchange over peer memory: table in `world` slices, windows in `world` ranges
    ranges = shard.partition(lengths, world)
    ctx.set_partition(rank, world)
    t0 = time.time()
    db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf, placement=1)
    load_s = allmax(time.time() - t0)
    lw_, ls_ = shard.local_slice(w.wins, w.segs, *ranges[rank])
    plan = ctx.plan(31, lw_, ls_)
    rows_x, ms = timed(lambda: screen_partitioned(ctx, db, plan))
    kmers = allsum(int(rows_x["total_kmers"].sum()))
    obs = allsum(int(rows_x["obs"].sum()))
    # TOTAL_KMERS / EFFLEN follow from the window geometry alone where a window holds no N run: at most window - k + 1
    out["exchange"] = {"value": kmers / (ms * 1e-3), "ms_per_step": ms, "seconds_per_job": ms * 1e-3, "db_load_s": load_s, "table_bytes_per_gpu": int(db.info.table_bytes),
                       "table_bytes_per_record": db.info.table_bytes / max(db.info.resident_kmers, 1), "records_per_gpu": int(db.info.resident_kmers),
                       "job_kmers": kmers, "job_observed_kmers": obs, "stash_kmers": int(db.info.stash_kmers)}
    for x_ in getattr(plan, "_exchange", None) or []:
        x_.close()
    plan.close()
    db.close()
    # ---- scan placement: 4 slices x (world / 4) window shards
    T = 4
    n_shards = world // T
    groups = [dist.new_group(list(range(s * T, (s + 1) * T))) for s in range(n_shards)]
    part_rank, shard_id, _ = shard.grid_layout(rank, world, T)
    ctx.set_partition(part_rank, T)
    torch.cuda.empty_cache()
    bench.log(f"[c4 r{rank}] before the second open: {torch.cuda.mem_get_info(local_rank)[0] / 1e9:.1f} GB free")
    t0 = time.time()
    err = None
    try:
        db = KMC(ctx, pre=w.kmc.pre, suf=w.kmc.suf, placement=1)
    except Exception as e:  # every rank has to learn of it: the ranks that opened their slice would wait in the collectives for ever
        err = repr(e)
        bench.log(f"[c4 r{rank}] scan placement: {err}")
    if allsum(int(err is not None)):
        out["scan_T4"] = {"error": err or "another rank could not open its slice"}
        ctx.close()
        barrier()
        if shm_dir:
            import shutil
            shutil.rmtree(shm_dir, ignore_errors=True)
        dist.destroy_process_group()
        if rank == 0:
            bench.emit(out)
        return
    load_s = allmax(time.time() - t0)
    rng = shard.partition(lengths, n_shards)[shard_id]
    lw_, ls_ = shard.local_slice(w.wins, w.segs, *rng)
    plan = ctx.plan(31, lw_, ls_)
    rows_s, ms = timed(lambda: screen_partitioned_scan(ctx, db, plan, group=groups[shard_id]))
    kmers_s = allsum(int(rows_s["total_kmers"].sum())) // T
    obs_s = allsum(int(rows_s["obs"].sum())) // T
    # the exchange rows of this rank's window range against the scan rows of the same windows (this rank holds the shard's rows)
    a, b = ranges[rank]
    same = bench.rows_equal(rows_x, rows_s[a - rng[0]:b - rng[0]]) if (a >= rng[0] and b <= rng[1]) else None
    n_same = allsum(int(bool(same))) if same is not None else allsum(0)
    n_checked = allsum(int(same is not None))
    out["scan_T4"] = {"value": kmers_s / (ms * 1e-3), "ms_per_step": ms, "seconds_per_job": ms * 1e-3, "db_load_s": load_s, "table_bytes_per_gpu": int(db.info.table_bytes),
                      "table_bytes_per_record": db.info.table_bytes / max(db.info.resident_kmers, 1), "records_per_gpu": int(db.info.resident_kmers),
                      "job_kmers": kmers_s, "job_observed_kmers": obs_s, "window_shards": n_shards, "stash_kmers": int(db.info.stash_kmers)}
    out["placements_agree"] = {"totals_equal": bool(kmers == kmers_s and obs == obs_s), "ranks_whose_rows_were_compared": n_checked, "ranks_with_identical_rows": n_same}
    plan.close()
    db.close()
    ctx.close()
    barrier()
    if shm_dir:
        import shutil
        shutil.rmtree(shm_dir, ignore_errors=True)
    dist.destroy_process_group()
    if rank == 0:
        bench.emit(out)


if __name__ == "__main__":
    main()
