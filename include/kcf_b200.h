/*
 * kcf_b200.h — C ABI of libkcfgpu.so, the B200 (sm_100a) implementation of kcftools'
 * `getVariations` hot path.  Plain C types only; no C++/torch types, no callbacks.
 *
 * What it replaces in the reference (sivasubramanics/kcftools v0.4.0; paths under
 * src/main/java/nl/wur/bis/kcftools/): the thread-pool fan-out and per-window scan in
 * Plugins/GetVariants.java:126-159 (processWindow :202-261), together with everything
 * that scan calls: Data/KMC.java (database open :56-189, getCount :292-326),
 * Data/Kmer.java, Data/Signature.java, Data/Fasta.java:90-167 and the substring
 * extraction of Data/FastaIndex.java:122-182.  The CLI, window generation, GTF parsing
 * and KCF emission stay on the host side (INTEGRATION.md shows the JNI / FFM binding).
 *
 * Conventions
 *   - every function returns KCF_OK (0) or a negative kcf_status; the message for the
 *     last failure on a context is kcf_last_error(ctx).  Where the reference calls
 *     Logger.error (print + System.exit(1), Utils/Logger.java:29-31) this library
 *     returns an error code instead; the Java shim maps non-zero to Logger.error.
 *   - the caller owns every input and output buffer; the library copies what it needs
 *     before returning and keeps no caller pointer.  Handles are opaque and are freed
 *     only by the matching close/destroy call.
 *   - there is NO CPU fallback: without a CUDA device, or for a database outside the
 *     supported envelope, the call fails.
 *   - one context drives one GPU (one process or thread per GPU); a context serialises
 *     its own stream.  Different contexts may be used concurrently.
 */
#ifndef KCF_B200_H
#define KCF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum kcf_status {
    KCF_OK = 0,
    KCF_ERR_CUDA = -1,          /* CUDA runtime failure or no device */
    KCF_ERR_IO = -2,            /* cannot read .kmc_pre / .kmc_suf */
    KCF_ERR_DB_FORMAT = -3,     /* not a KMC 0x200 database (KMC.java:139-141) or inconsistent sizes / LUT */
    KCF_ERR_UNSUPPORTED = -4,   /* valid database outside the envelope (k > 64, (k-P) % 4 != 0, counter > 4 B, k > 32 with a partitioned table, ...) */
    KCF_ERR_ARG = -5,           /* bad argument (null handle, min_count < 1 as GetVariants.java:383-385, ...) */
    KCF_ERR_RANGE = -6,         /* window segment outside its sequence (FastaIndex.java:132-135) or no segment (GetVariants.java:213-216) */
    KCF_ERR_FASTA = -7,         /* read past the mapped sequence bytes, e.g. no trailing newline (FastaIndex.java:175-177) */
    KCF_ERR_WEIGHTS = -8,       /* wi + wt + wr != 1.0 when a score had to be computed (Data.java:101-103) */
    KCF_ERR_NOMEM = -9,         /* device or host allocation failed */
    KCF_ERR_DB_ORDER = -10      /* records of one (bin, prefix) range not strictly ascending: the reference's binary search is undefined */
} kcf_status;

typedef struct kcf_ctx kcf_ctx;   /* one GPU + its stream + the resident reference sequences */
typedef struct kcf_db kcf_db;     /* one KMC database resident in HBM (replaces the KMC object, Data/KMC.java) */
typedef struct kcf_plan kcf_plan; /* a window list resident on the device, reusable across databases */

/* One output row = one window (what processWindow delivers through Window.addTotalKmers / setEffLength /
 * addData, GetVariants.java:254-258).  48 bytes, natural alignment. */
typedef struct kcf_result_t {
    int32_t total_kmers;     /* TOTAL_KMERS column */
    int32_t eff_len;         /* EFFLEN, Fasta.getEffectiveATGCCount (Fasta.java:140-167) */
    int32_t obs;             /* OB */
    int32_t variations;      /* VA */
    int32_t inner;           /* ID */
    int32_t left;            /* LD */
    int32_t right;           /* RD */
    int32_t _pad;
    int64_t kmer_count_sum;  /* Σ count over observed k-mers; KD = sum / obs (Data.java:87) */
    double score;            /* SC, Data.computeScore (Data.java:95-107); Java recomputes it from the integers */
} kcf_result_t;

/* A window is the concatenation of n_segs segments starting at segs[first_seg]
 * (fixed / sliding window: 1 segment, Window.java:224-226; gene / transcript: the merged exon loci in
 * concatenation order, GTF.java:223-248). */
typedef struct kcf_window_t {
    uint32_t first_seg;
    uint32_t n_segs;
} kcf_window_t;

typedef struct kcf_segment_t {
    int32_t seq_id;   /* id returned by kcf_ref_add */
    int32_t start0;   /* 0-based start on the sequence */
    int32_t len;      /* number of bases (> 0) */
} kcf_segment_t;

typedef struct kcf_db_info_t {
    int32_t kmer_length;        /* KMC.getKmerLength() */
    int32_t lut_prefix_length;  /* KMC.getPrefixLength() */
    int32_t signature_length;
    int32_t counter_size;
    int32_t both_strands;       /* KMC.isBothStrands(): 1 when the stored flag byte is 0 */
    int32_t min_count;
    int32_t max_count;
    int32_t n_bins;
    int64_t total_kmers;        /* records in .kmc_suf */
    int64_t resident_kmers;     /* records a reference getCount() can return (inserted in the HBM table) */
    int64_t unreachable_kmers;  /* records no reference lookup can ever reach (wrong bin / non-canonical / before LUT[0]) */
    int64_t stash_kmers;        /* resident records living in the overflow stash */
    int64_t table_bytes;        /* HBM bytes of the lookup structure */
    int64_t n_buckets;          /* 128-byte table lines */
    double load_seconds;        /* wall time of the open call */
    int64_t elsewhere_kmers;    /* placement 1: reachable records whose home line belongs to another rank's slice */
    double load_phase_s[4];     /* where load_seconds went: [0] allocations + table reset queued, [1] records streamed (host staging,
                                 * H2D copies and ingest kernels queued), [2] wait for the device + stash build, [3] device time from the
                                 * first to the last ingest kernel (CUDA events) */
} kcf_db_info_t;

/* ---- context ------------------------------------------------------------------------- */
/* device: CUDA ordinal (LOCAL_RANK in a one-process-per-GPU job). */
int kcf_init(int device, kcf_ctx **out);
void kcf_shutdown(kcf_ctx *ctx);
const char *kcf_last_error(kcf_ctx *ctx); /* ctx may be NULL: message of the last failed kcf_init */
/* The CUDA stream (cudaStream_t) every call on this context is ordered on; lets a caller time the
 * kernels with events on the launching stream. */
void *kcf_stream(kcf_ctx *ctx);
/* Pinned host memory for callers that want asynchronous full-rate copies (optional). */
int kcf_host_alloc(kcf_ctx *ctx, uint64_t n_bytes, void **out);
void kcf_host_free(kcf_ctx *ctx, void *p);

/* ---- database: replaces `new KMC(prefix, inMemory)` (KMC.java:56-78) ------------------- */
/* placement: 0 = whole database resident on this context's GPU (replicated across contexts);
 *            1 = this context keeps only its slice of the table's line space (kcf_set_partition; screened through
 *                the kcf_xchg_* calls below). */
int kcf_db_open(kcf_ctx *ctx, const char *kmc_prefix, int placement, kcf_db **out);
/* Same, from the byte images of the two files (what the reference mmaps, KMC.java:112, 173-189). */
int kcf_db_open_mem(kcf_ctx *ctx, const uint8_t *pre, uint64_t pre_len, const uint8_t *suf, uint64_t suf_len,
                    int placement, kcf_db **out);
int kcf_db_info(kcf_db *db, kcf_db_info_t *out);
void kcf_db_close(kcf_db *db);
/* Load-factor target for subsequently opened databases (0 < lf <= 0.9; 0 = automatic, the default: 0.15, denser when
 * the table would take more than 40 % of the device memory). */
int kcf_set_load_factor(kcf_ctx *ctx, double lf);
/* Minimizer length of the home-line function for subsequently opened databases (1..24; 0 = chosen from the
 * database size and k).  A tuning / test knob: results never depend on it. */
int kcf_set_minimizer_length(kcf_ctx *ctx, int m);

/* KMC.getCount for a batch of k-mers given as ASCII (n * k bytes, upper or lower case ACGT), canonicalised
 * per the database's both_strands flag like GetVariants.java:222-223.  counts_out[i] is the Java int.
 * Exists for parity tests of the lookup structure. */
int kcf_db_count(kcf_ctx *ctx, kcf_db *db, const char *kmers_ascii, uint64_t n, int32_t *counts_out);

/* ---- reference sequences: replaces the per-sequence mmap of FastaIndex (FastaIndex.java:26-77) */
/* bytes = the raw file bytes of one sequence from its .faidx offset to the next sequence's offset (or EOF),
 * newlines included; line_bases / line_width / seq_len are the .faidx columns. */
int kcf_ref_add(kcf_ctx *ctx, const uint8_t *fasta_seq_bytes, uint64_t n_bytes, uint32_t line_bases,
                uint32_t line_width, uint64_t seq_len, int *seq_id_out);
/* Same without the final synchronisation: the copy is queued on the context's copy stream (it overlaps the 2-bit
 * packing of the previous sequence and any screening already queued), and `fasta_seq_bytes` must stay valid and
 * unmodified until kcf_ref_sync, kcf_plan_fetch or kcf_screen returns.  Lets a host upload chromosome i+1 while
 * chromosome i is being screened (the reference walks the sequences one after the other, GetVariants.java:117-121). */
int kcf_ref_add_async(kcf_ctx *ctx, const uint8_t *fasta_seq_bytes, uint64_t n_bytes, uint32_t line_bases,
                      uint32_t line_width, uint64_t seq_len, int *seq_id_out);
int kcf_ref_sync(kcf_ctx *ctx);
/* Forget all sequences (their device memory is kept for reuse by later kcf_ref_add calls).  Plans created before the
 * call refer to sequences that no longer exist: running them afterwards fails with KCF_ERR_ARG. */
int kcf_ref_clear(kcf_ctx *ctx);

/* ---- screening: replaces GetVariants.java:126-159 ---------------------------------------- */
/* One call screens all windows against the database and fills out[n_wins].
 * w = {wi, wt, wr} = {--wi, --wt, --wr} in the order of getWeights() (GetVariants.java:388-390). */
int kcf_screen(kcf_ctx *ctx, kcf_db *db, const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs,
               uint64_t n_segs, int32_t min_count, const double w[3], kcf_result_t *out);

/* ---- one job on several GPUs of one process: replaces the pool over all windows, GetVariants.java:129-151 ---------
 * The host side of the reference hands every window to one pool (`-t` threads); here the window list is cut into n
 * contiguous ranges balanced on the number of bases (kcf_shard_windows: bounds_out[n_shards + 1]) and range g is screened
 * on ctxs[g] against dbs[g] — the same database opened (placement 0) on every context.  The reference sequences are
 * given as HOST bytes (seqs[i] = what kcf_ref_add takes for sequence id i): each context uploads only the line-aligned
 * stretches its windows touch, screening them while the next stretch crosses PCIe, and the rows land in out[] at their
 * window's index.  One host thread per context; no communication library, no device-to-device traffic.  The call
 * replaces whatever sequences the contexts held (kcf_ref_clear).  n = 1 is the plain "screen from host buffers" call.
 * Errors: the first failing context's code; its message is copied to ctxs[0]. */
typedef struct kcf_host_seq_t {
    const uint8_t *bytes;  /* raw FASTA bytes of the sequence, newlines included (FastaIndex.java:54-68) */
    uint64_t n_bytes;
    uint32_t line_bases;   /* .faidx columns */
    uint32_t line_width;
    uint64_t seq_len;
} kcf_host_seq_t;
/* Upper bound, in bases, of the stretches kcf_screen_sharded uploads one at a time (0 = default, 96 Mbases).  A tuning / test knob:
 * results never depend on it. */
int kcf_set_upload_piece(kcf_ctx *ctx, uint64_t bases);
int kcf_shard_windows(const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs, uint64_t n_segs, int n_shards,
                      uint64_t *bounds_out);
int kcf_screen_sharded(kcf_ctx *const *ctxs, kcf_db *const *dbs, int n, const kcf_host_seq_t *seqs, uint32_t n_seqs,
                       const kcf_window_t *wins, uint64_t n_wins, const kcf_segment_t *segs, uint64_t n_segs, int32_t min_count,
                       const double w[3], kcf_result_t *out);

/* The same in three steps, for callers that screen one window list against many databases (cohorts)
 * or want the device-resident part timed alone. */
int kcf_plan_create(kcf_ctx *ctx, int32_t kmer_length, const kcf_window_t *wins, uint64_t n_wins,
                    const kcf_segment_t *segs, uint64_t n_segs, kcf_plan **out);
int kcf_plan_run(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, int32_t min_count, const double w[3]); /* asynchronous on kcf_stream */
int kcf_plan_fetch(kcf_ctx *ctx, kcf_plan *plan, kcf_result_t *out);                              /* D2H + synchronise */
void kcf_plan_destroy(kcf_plan *plan);
/* Σ total_kmers of the last run and the number of kernels launched by it. */
int kcf_plan_stats(kcf_plan *plan, uint64_t *n_tiles, uint64_t *n_positions, uint32_t *kernels_per_run);
/* Per-valid-k-mer counts of one window of the last run (debug / parity; runs a separate pass). */
int kcf_window_counts(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t window, int32_t *counts_out, uint64_t cap,
                      uint64_t *n_out);

/* ---- partitioned database (placement 1): for a database beyond one GPU's HBM ---------------------------------------
 * The line space of the table is cut in `world` equal ranges; the context opened after kcf_set_partition(rank, world)
 * keeps the records whose home line falls in range `rank` (kcf_db_info_t.elsewhere_kmers counts the others).  A rank
 * can then not answer its own k-mers: screening becomes extract -> all-to-all -> lookup -> all-to-all -> fold, with
 * the two all-to-all steps done by the HOST over caller-owned device buffers (NCCL; kcftools_b200/partitioned.py).
 * This is the one place where the getVariations path has a real exchange step; the reference has no
 * counterpart (its database always lives in one JVM, KMC.java:56-78). */
int kcf_set_partition(kcf_ctx *ctx, int rank, int world);
/* Requester, step 1: canonical k-mers of tiles [tile_begin, tile_end) of the plan (a tile = 2048 window positions,
 * kcf_plan_stats), grouped by owning rank.  d_keys_out (uint64), d_homes_out (uint32: global home line), d_src_out
 * (uint32: position inside the batch) are DEVICE buffers of `cap` entries (cap >= positions of the batch is always
 * enough); send_counts[world] (host) receives how many consecutive entries go to each rank. */
int kcf_xchg_extract(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, int world,
                     void *d_keys_out, void *d_homes_out, void *d_src_out, uint64_t cap, uint64_t *send_counts);
/* Owner, step 2: KMC.getCount (KMC.java:292-326) of n received k-mers against this rank's slice -> uint32 counts (device). */
int kcf_xchg_lookup(kcf_ctx *ctx, kcf_db *db, const void *d_keys, const void *d_homes, uint64_t n, void *d_counts_out);
/* Requester, step 3: the counts of the n entries sent in step 1 (same order; d_src = step 1's d_src_out) -> gap
 * summaries of tiles [tile_begin, tile_end) (GetVariants.java:217-252). */
int kcf_xchg_fold(kcf_ctx *ctx, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, const void *d_counts_back,
                  const void *d_src, uint64_t n, int32_t min_count);
/* Requester, last: scores and rows from the tile summaries (Data.java:70-107); then kcf_plan_fetch. */
int kcf_plan_finalize(kcf_ctx *ctx, kcf_plan *plan, const double w[3]);

/* The same exchange over PEER MEMORY (NVLink): no host in the loop, no grouping pass, one byte back per k-mer.  Every rank
 * creates a workspace (kcf_xg_create, sized for batches of batch_tiles tiles), exports it (kcf_xg_export: a 64-byte CUDA IPC
 * handle for ranks in other processes, or the raw device pointer for ranks of the same process) and connects to the others'
 * (kcf_xg_connect: handles = world x 64 bytes, or same_process_ptrs[world]).  Per batch of tiles, stream-ordered, no host
 * synchronisation inside the library:
 *   kcf_xg_send    requester: Fasta.getKmersList + home line of every k-mer (Fasta.java:90-127); RUNS of up to 11 consecutive k-mers
 *                  sharing a home line (they share their minimizer) go as one 16-byte entry — the bases they span — straight
 *                  into their owner's inbox
 *   -- barrier over all ranks on kcf_stream (the caller's: e.g. a one-element NCCL all-reduce) --
 *   kcf_xg_answer  owner: canonical k-mers of every run (Kmer.java:57-79) and KMC.getCount (KMC.java:292-326) against the run's line,
 *                  fetched once; counts stored straight into the requesters' workspaces
 *   -- barrier --
 *   kcf_xg_fold    requester: gap summaries of the batch's tiles (GetVariants.java:217-252)
 * then kcf_plan_finalize / kcf_plan_fetch.  kcf_xg_status synchronises and reports a workspace overflow (KCF_ERR_NOMEM). */
typedef struct kcf_xg kcf_xg;
int kcf_xg_create(kcf_ctx *ctx, kcf_db *db, int rank, int world, uint64_t batch_tiles, kcf_xg **out);
int kcf_xg_export(kcf_xg *x, void *handle64_out, void **device_ptr_out, uint64_t *bytes_out);
int kcf_xg_connect(kcf_xg *x, const void *handles, void *const *same_process_ptrs);
void kcf_xg_destroy(kcf_xg *x);
int kcf_xg_send(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, kcf_xg *x, uint64_t tile_begin, uint64_t tile_end);
int kcf_xg_answer(kcf_ctx *ctx, kcf_db *db, kcf_xg *x);
int kcf_xg_fold(kcf_ctx *ctx, kcf_plan *plan, kcf_xg *x, uint64_t tile_begin, uint64_t tile_end, int32_t min_count);
int kcf_xg_status(kcf_xg *x, uint64_t *bytes_out_per_run, uint64_t *bytes_back_per_run, uint64_t *runs_sent_last_batch);
/* Pipelining with two workspaces: after kcf_xg_pipeline the sends of this workspace run on a stream of their own, on at most
 * send_ctas_per_sm resident CTAs per SM (0 = no cap), ordered after what the context's stream held when kcf_xg_send was
 * called; kcf_xg_join makes the context's stream wait for the workspace's last send.  Per batch b a rank then queues
 * send(b + 1, other workspace), answer(b), join(other), barrier, fold(b): one barrier per batch, the send of the next batch
 * beside the answers of this one. */
int kcf_xg_pipeline(kcf_xg *x, uint32_t send_ctas_per_sm);
int kcf_xg_join(kcf_ctx *ctx, kcf_xg *x);

/* Scan placement — the second way to screen against a partitioned table, without moving k-mers: the plan holds ALL
 * windows on every rank (the 2-bit reference is replicated; it is small next to the table), every rank runs
 * Fasta.getKmersList (Fasta.java:90-127) over tiles [tile_begin, tile_end) but does KMC.getCount (KMC.java:292-326)
 * only for the k-mers whose home line lies in its slice.  Output (device buffers of the caller): d_hit_out = one uint32
 * per 32 positions (bit = k-mer observed with count >= min_count BY THIS RANK), d_sum_out = one uint64 per tile (Σcount
 * of those hits).  The caller sum-reduces both over the ranks (the owners' bitmaps are disjoint: + is OR; NCCL
 * all-reduce) and hands the totals to kcf_scan_fold, which rebuilds the gap summaries (GetVariants.java:217-252);
 * then kcf_plan_finalize / kcf_plan_fetch as above. */
int kcf_scan_owned(kcf_ctx *ctx, kcf_db *db, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, int32_t min_count,
                   void *d_hit_out, void *d_sum_out);
int kcf_scan_fold(kcf_ctx *ctx, kcf_plan *plan, uint64_t tile_begin, uint64_t tile_end, const void *d_hit, const void *d_sum);

/* ---- windows x samples matrix: cohort, findIBS, kcf2gt ---------------------------------------
 * The consumers of getVariations output, kept on the device: a cohort is the matrix [sample][window] of the per-window
 * integers, filled either straight from the rows of a kcf_plan (no KCF text in between; replaces writing and re-reading
 * one file per sample before Plugins/Cohort.java:71-119) or from rows the host parsed out of KCF files
 * (Data/KCFReader.java:31-105, Data/Window.java:42-83). */
typedef struct kcf_cohort kcf_cohort;
typedef struct kcf_cell_t {   /* one sample in one window = Data/Data.java:16-24; 40 bytes */
    int32_t obs, variations, inner, left, right;
    int32_t ibs;              /* -1 = "N" */
    int64_t kmer_count;       /* Σ count (from a plan) or Math.round(KD * obs) (from a KCF row, Window.java:70) */
    double score;
} kcf_cell_t;
/* total_kmers / eff_len: per window (host arrays), or both NULL: the first sample added with kcf_cohort_add_plan then
 * defines them, and has to be added completely (all its plans) before any other sample. */
int kcf_cohort_create(kcf_ctx *ctx, uint64_t n_windows, uint32_t n_samples, const int32_t *total_kmers, const int32_t *eff_len,
                      kcf_cohort **out);
void kcf_cohort_destroy(kcf_cohort *c);
/* Column `sample`, rows [window_offset, window_offset + plan windows) <- the results of a plan that has run (device to
 * device, asynchronous).  A sample whose TOTAL_KMERS / EFFLEN differ from the cohort's is reported as "Windows mismatch"
 * by kcf_cohort_scores / kcf_cohort_fetch.  This is an extra check of this path, NOT reference behaviour: Cohort.java:92-94
 * raises the mismatch only for a window id missing from the first file and otherwise keeps the first file's totals
 * silently (the file-based `cohort` of the host program does exactly that); plans made from one window list cannot differ
 * there, so the check only fires on a caller's mistake. */
int kcf_cohort_add_plan(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, uint64_t window_offset, kcf_plan *plan);
/* Column `sample` <- n_windows cells parsed by the host (needs the totals given to kcf_cohort_create). */
int kcf_cohort_set_sample(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, const kcf_cell_t *cells);
/* Every cell's score from its integers and the weights {wi, wt, wr}, as the reference does whenever it reads a KCF row
 * (Window.java:42-83 -> Data.java:41-67, 95-107); KCF_ERR_WEIGHTS as in kcf_plan_fetch. */
int kcf_cohort_scores(kcf_ctx *ctx, kcf_cohort *c, const double w[3]);
/* FindIBS.java:118-158 for every sample: order[p] = window index at traversal position p (the reference walks a HashMap
 * of chromosomes, windows in file order inside each), chrom[p] = ordinal of that window's chromosome; a window is IBS
 * when score >= cutoff (detect_var: score < cutoff), blocks are numbered per sample; result in kcf_cell_t.ibs. */
int kcf_cohort_find_ibs(kcf_ctx *ctx, kcf_cohort *c, const uint32_t *order, const uint32_t *chrom, uint64_t n, int detect_var,
                        int32_t min_consecutive, float score_cutoff);
/* KCFToGenotypeTable.java:116-133, 159-172: alleles_out[window][sample] in {0, 2, 1, -1}, bad_out[window] = badWindow(). */
int kcf_cohort_genotypes(kcf_ctx *ctx, kcf_cohort *c, double score_a, double score_b, double score_n, double min_maf, double max_missing,
                         int8_t *alleles_out, uint8_t *bad_out);
/* Column `sample` (and the per-window totals; any pointer may be NULL) back to the host. */
int kcf_cohort_fetch(kcf_ctx *ctx, kcf_cohort *c, uint32_t sample, kcf_cell_t *cells_out, int32_t *total_kmers_out, int32_t *eff_len_out);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* Random 32-byte-sector gather bandwidth of this GPU (the random-access roofline of SURVEY §8(d)):
 * n_loads independent 32-B loads from uniformly random sector addresses of a buffer of n_bytes. */
int kcf_measure_random_sector_gbps(kcf_ctx *ctx, uint64_t n_bytes, uint64_t n_loads, int repeats, double *gbps_out);
/* Random 128-byte LINE gather rate of this GPU (lines per second): 4 lanes read the 4 sectors of one uniformly random line
 * with one coalesced request — the access pattern of this library's table, hence its memory-side bound. */
int kcf_measure_random_line_rate(kcf_ctx *ctx, uint64_t n_bytes, uint64_t n_lines_read, int repeats, double *lines_per_s_out);
/* Layout statistics: hist_out[n] = table lines (home + overflow) holding n keys, n = 0 .. 15. */
int kcf_db_line_histogram(kcf_db *db, uint64_t hist_out[16]);
/* Milliseconds of the screening kernel in the last kcf_plan_run when profiling is on. */
int kcf_set_profiling(kcf_ctx *ctx, int on);
int kcf_last_kernel_ms(kcf_ctx *ctx, float *screen_ms, float *finalize_ms);
const char *kcf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* KCF_B200_H */
