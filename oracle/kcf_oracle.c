/*
 * kcf_oracle.c — CPU restatement of kcftools' `getVariations` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kcftools_b200/ (the product) may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker and
 * as the reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (sivasubramanics/kcftools v0.4.0, Java 17)
 * ships no tests, fixtures or golden files, and no JDK exists in the build
 * container, so this restatement cannot be checked against reference-authored
 * vectors or against the reference itself.  It is pinned only by (a) the
 * hand-derived known answers listed in SURVEY.md §8(c), (b) an independent
 * pure-Python restatement (oracle/pyoracle.py) and (c) the doc snippets the
 * reference publishes (window tiling ids).
 *
 * Each function cites the reference lines it follows.  Path abbreviations:
 *   P/ = src/main/java/nl/wur/bis/kcftools/Plugins/
 *   D/ = src/main/java/nl/wur/bis/kcftools/Data/
 *   U/ = src/main/java/nl/wur/bis/kcftools/Utils/
 *
 * The algorithmic shape is deliberately the reference's, not a fast one:
 * every k-mer is re-packed from characters in O(k), reverse-complemented in an
 * O(k) loop, its signature is the minimum of k-L+1 table lookups, and the
 * count comes from a per-(bin,prefix) binary search over the on-disk records
 * with an unsigned byte-wise comparison.  Only the Java object allocations are
 * gone.  k up to 256 (long[] words like Kmer.java); the CUDA path covers k <= 32.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -shared -fPIC -pthread)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define ORC_OK 0
#define ORC_ERR_IO (-1)
#define ORC_ERR_FORMAT (-2)
#define ORC_ERR_ARG (-3)
#define ORC_ERR_FATAL (-4) /* a condition on which the reference calls Logger.error => System.exit(1) */

/* ------------------------------------------------------------------ */
/* D/Signature.java                                                   */
/* ------------------------------------------------------------------ */

/* D/Signature.java:42-76 isAllowed */
static int sig_is_allowed(uint32_t signature, int sign_len)
{
    if ((signature & 0x3F) == 0x3F) return 0; /* TTT suffix */
    if ((signature & 0x3F) == 0x3B) return 0; /* TGT suffix */
    if ((signature & 0x3C) == 0x3C) return 0; /* TG* suffix */
    for (int j = 0; j < sign_len - 3; ++j) {
        if ((signature & 0xF) == 0) return 0; /* AA inside */
        signature >>= 2;
    }
    if (signature == 0) return 0;    /* AAA prefix */
    if (signature == 0x04) return 0; /* ACA prefix */
    if ((signature & 0xF) == 0) return 0; /* *AA prefix */
    return 1;
}

/* D/Signature.java:82-95 getRev */
static uint32_t sig_get_rev(uint32_t sequence, int length)
{
    uint32_t rc = 0;
    for (int i = 0; i < length; i++) {
        uint32_t base = sequence & 3u;
        base = (~base) & 3u;
        rc = (rc << 2) | base;
        sequence >>= 2;
    }
    return rc;
}

/* D/Signature.java:23-37 initNorm.  out has 4^sign_len entries. */
int orc_norm_table(int sign_len, int32_t *out)
{
    if (sign_len < 3 || sign_len > 13) return ORC_ERR_ARG;
    uint32_t special = 1u << (sign_len * 2);
    for (uint32_t i = 0; i < special; ++i) {
        uint32_t rev = sig_get_rev(i, sign_len);
        uint32_t str_val = sig_is_allowed(i, sign_len) ? i : special;
        uint32_t rev_val = sig_is_allowed(rev, sign_len) ? rev : special;
        out[i] = (int32_t)(str_val < rev_val ? str_val : rev_val);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* D/Kmer.java (long[] words, 32 bases per word, k <= 256)            */
/* ------------------------------------------------------------------ */
#define ORC_MAX_K 256
#define ORC_KWORDS (ORC_MAX_K / 32)
typedef struct kword { uint64_t w[ORC_KWORDS]; } kword; /* Kmer.kmerLong; words past ceil(2k/64) stay 0 */

/* D/Kmer.java:286-294 baseToBits; input already upper-cased ACGT */
static inline uint64_t base_to_bits(char b)
{
    switch (b) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    default:  return 3; /* 'T' */
    }
}

/* base i of a packed k-mer: word i/32, bits 62-2(i%32) (bit offsets are always even, so the reference's
 * "split between two longs" branches, D/Kmer.java:216-220, 245-248, 315-319, never run) */
static inline uint32_t kword_base(const kword *x, int i) { return (uint32_t)((x->w[i >> 5] >> (62 - 2 * (i & 31))) & 3u); }

/* D/Kmer.java:232-252 kmerToLong: base i at bits 62-2(i%32) of word i/32 (MSB first, left aligned) */
static kword kmer_to_long(const char *kmer, int k)
{
    kword r;
    memset(&r, 0, sizeof r);
    for (int i = 0; i < k; i++) r.w[i >> 5] |= base_to_bits(kmer[i]) << (62 - 2 * (i & 31));
    return r;
}

/* D/Kmer.java:300-338 getReverseComplement */
static kword kmer_revcomp(const kword *x, int k)
{
    kword rev;
    memset(&rev, 0, sizeof rev);
    for (int i = 0; i < k; i++) {
        uint64_t comp = (~(uint64_t)kword_base(x, i)) & 3u;
        int j = k - i - 1;
        rev.w[j >> 5] |= comp << (62 - 2 * (j & 31));
    }
    return rev;
}

/* D/Kmer.java:72-79 getCanonical + :406-414 compareLongArrays (word by word, unsigned; tie keeps forward) */
static kword kmer_canonical(kword fwd, int k, int both_strands)
{
    if (both_strands) {
        kword rc = kmer_revcomp(&fwd, k);
        int nw = (2 * k + 63) / 64;
        for (int i = 0; i < nw; i++) {
            if (fwd.w[i] != rc.w[i]) return fwd.w[i] > rc.w[i] ? rc : fwd;
        }
    }
    return fwd;
}

/* D/Kmer.java:208-226 extractIntFromBits */
static uint32_t extract_int_from_bits(const kword *x, int start_base, int length_bases)
{
    uint32_t result = 0;
    for (int i = 0; i < length_bases; i++) result = (result << 2) | kword_base(x, start_base + i);
    return result;
}

/* D/Kmer.java:105-118 getSignature */
static int32_t kmer_signature(const kword *x, int k, int sign_len, const int32_t *norm)
{
    uint32_t cur = extract_int_from_bits(x, 0, sign_len);
    int32_t min_sig = norm[cur];
    uint32_t mask = (1u << (2 * sign_len)) - 1u;
    for (int i = 1; i <= k - sign_len; i++) {
        cur = ((cur << 2) & mask) | extract_int_from_bits(x, i + sign_len - 1, 1);
        if (norm[cur] < min_sig) min_sig = norm[cur];
    }
    return min_sig;
}

/* D/Kmer.java:143-170 extractSuffix: bases P..k-1 packed 4 per byte, first base in bits 7..6 */
static void kmer_extract_suffix(const kword *x, int k, int prefix_len, uint8_t *suffix /* (k-P+3)/4 bytes */)
{
    int suffix_len = k - prefix_len;
    memset(suffix, 0, (size_t)(suffix_len + 3) / 4);
    for (int i = 0; i < suffix_len; ++i) suffix[i / 4] |= (uint8_t)(kword_base(x, prefix_len + i) << ((3 - i % 4) * 2));
}

/* ------------------------------------------------------------------ */
/* D/KMC.java                                                         */
/* ------------------------------------------------------------------ */

typedef struct orc_kmc {
    int32_t kmer_length, mode, counter_size, lut_prefix_length, signature_length;
    int32_t min_count, max_count;
    int64_t total_kmers;
    int32_t both_strands; /* 1 when the stored flag byte is 0 (D/KMC.java:133) */
    int32_t version;
    int32_t suffix_length;        /* kmerLength - lutPrefixLength, D/KMC.java:128 */
    int32_t record_size;          /* counterSize + sufixLength/4,  D/KMC.java:61 */
    int32_t lut_prefix_array_size;/* 4^P */
    int64_t prefix_array_len;     /* numPrefixArrays * 4^P */
    uint64_t *prefix_array;
    int64_t signature_map_len;    /* 4^L + 1 */
    int32_t *signature_map;
    int32_t *norm;                /* Signature table */
    uint8_t *suf;                 /* records, after the 4-byte marker */
    int64_t suf_len;
    int owns_suf;
} orc_kmc;

static uint32_t rd_u32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint64_t rd_u64(const uint8_t *p) { return (uint64_t)rd_u32(p) | ((uint64_t)rd_u32(p + 4) << 32); }

void orc_kmc_close(orc_kmc *db)
{
    if (!db) return;
    free(db->prefix_array);
    free(db->signature_map);
    free(db->norm);
    if (db->owns_suf == 1) free(db->suf);
    else if (db->owns_suf == 2) free(db->suf - 4); /* whole-file image from orc_kmc_open */
    free(db);
}

/* D/KMC.java:107-168 readPrefixFile, :56-63 ctor, :84-102 preloadSuffixBuffers (paging removed:
 * pages hold whole records, so a flat array addresses the same bytes). The suffix image is
 * borrowed when copy_suf == 0. */
int orc_kmc_open_mem(const uint8_t *pre, int64_t pre_len, const uint8_t *suf, int64_t suf_len,
                     int copy_suf, orc_kmc **out)
{
    *out = NULL;
    if (pre_len < 16) return ORC_ERR_FORMAT;
    orc_kmc *db = (orc_kmc *)calloc(1, sizeof(orc_kmc));
    if (!db) return ORC_ERR_IO;
    int64_t file_size = pre_len;
    int32_t header_offset = (int32_t)rd_u32(pre + file_size - 8);              /* :117-118 */
    int64_t hpos = file_size - header_offset - 8;                                /* :121 */
    if (hpos < 4 || hpos + 68 > file_size) { free(db); return ORC_ERR_FORMAT; }
    const uint8_t *h = pre + hpos;
    db->kmer_length = (int32_t)rd_u32(h + 0);
    db->mode = (int32_t)rd_u32(h + 4);
    db->counter_size = (int32_t)rd_u32(h + 8);
    db->lut_prefix_length = (int32_t)rd_u32(h + 12);
    db->suffix_length = db->kmer_length - db->lut_prefix_length;
    db->signature_length = (int32_t)rd_u32(h + 16);
    db->min_count = (int32_t)rd_u32(h + 20);
    db->max_count = (int32_t)rd_u32(h + 24);
    db->total_kmers = (int64_t)rd_u64(h + 28);
    db->both_strands = (h[36] == 0);                                              /* :133 */
    db->version = (int32_t)rd_u32(h + 36 + 1 + 3 + 24);                          /* :134-138 */
    if (db->version != 0x200) { free(db); return ORC_ERR_FATAL; }                /* :139-141 */
    if (db->kmer_length < 1 || db->kmer_length > ORC_MAX_K || db->signature_length < 3 || db->signature_length > 13 ||
        db->lut_prefix_length < 0 || db->lut_prefix_length > db->kmer_length || db->lut_prefix_length > 15 || db->counter_size < 0 ||
        db->counter_size > 4 || (db->suffix_length % 4) != 0) {
        free(db);
        return ORC_ERR_ARG; /* outside what this restatement (and the CUDA path) covers */
    }
    int64_t sig_map_size = ((int64_t)1 << (2 * db->signature_length)) + 1;       /* :145 */
    int64_t sig_map_start = file_size - header_offset - 8 - sig_map_size * 4;    /* :146 */
    if (sig_map_start < 4) { free(db); return ORC_ERR_FORMAT; }
    db->signature_map_len = sig_map_size;
    db->signature_map = (int32_t *)malloc((size_t)sig_map_size * 4);
    for (int64_t i = 0; i < sig_map_size; i++) db->signature_map[i] = (int32_t)rd_u32(pre + sig_map_start + 4 * i);
    db->lut_prefix_array_size = 1 << (2 * db->lut_prefix_length);                /* :154 */
    int64_t single_lut_size = (int64_t)db->lut_prefix_array_size * 8;            /* :155 */
    int64_t num_prefix_arrays = (sig_map_start - 8 - 4) / single_lut_size;       /* :156 */
    if (num_prefix_arrays < 0) num_prefix_arrays = 0;
    db->prefix_array_len = num_prefix_arrays * db->lut_prefix_array_size;
    db->prefix_array = (uint64_t *)malloc((size_t)(db->prefix_array_len ? db->prefix_array_len : 1) * 8);
    for (int64_t i = 0; i < db->prefix_array_len; i++) db->prefix_array[i] = rd_u64(pre + 4 + 8 * i); /* :153,159-163 */
    db->norm = (int32_t *)malloc(((size_t)1 << (2 * db->signature_length)) * 4);
    orc_norm_table(db->signature_length, db->norm);                              /* :60 */
    db->record_size = db->counter_size + db->suffix_length / 4;                  /* :61 */
    /* :94 — first 4 bytes of .kmc_suf are the KMCS marker */
    int64_t need = 4 + db->total_kmers * db->record_size;
    if (suf_len < need) { orc_kmc_close(db); return ORC_ERR_FORMAT; }
    if (copy_suf) {
        db->suf = (uint8_t *)malloc((size_t)(need - 4 ? need - 4 : 1));
        memcpy(db->suf, suf + 4, (size_t)(need - 4));
        db->owns_suf = 1;
    } else {
        db->suf = (uint8_t *)(uintptr_t)(suf + 4);
        db->owns_suf = 0;
    }
    db->suf_len = need - 4;
    *out = db;
    return ORC_OK;
}

static uint8_t *read_whole(const char *path, int64_t *len)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    int64_t n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = (uint8_t *)malloc((size_t)(n ? n : 1));
    if (b && fread(b, 1, (size_t)n, f) != (size_t)n) { free(b); b = NULL; }
    fclose(f);
    *len = n;
    return b;
}

/* D/KMC.java:56-59: prefix + ".kmc_pre" / ".kmc_suf", -m (in-memory) mode */
int orc_kmc_open(const char *prefix, orc_kmc **out)
{
    char path[4096];
    int64_t pre_len = 0, suf_len = 0;
    snprintf(path, sizeof path, "%s.kmc_pre", prefix);
    uint8_t *pre = read_whole(path, &pre_len);
    if (!pre) return ORC_ERR_IO;
    snprintf(path, sizeof path, "%s.kmc_suf", prefix);
    uint8_t *suf = read_whole(path, &suf_len);
    if (!suf) { free(pre); return ORC_ERR_IO; }
    int rc = orc_kmc_open_mem(pre, pre_len, suf, suf_len, 0, out);
    free(pre);
    if (rc != ORC_OK) { free(suf); return rc; }
    (*out)->owns_suf = 2; /* db->suf == suf + 4; close frees the whole image */
    return ORC_OK;
}

typedef struct {
    int32_t kmer_length, lut_prefix_length, signature_length, counter_size, both_strands;
    int32_t min_count, max_count, n_bins;
    int64_t total_kmers;
} orc_kmc_info;

void orc_kmc_get_info(const orc_kmc *db, orc_kmc_info *o)
{
    o->kmer_length = db->kmer_length;
    o->lut_prefix_length = db->lut_prefix_length;
    o->signature_length = db->signature_length;
    o->counter_size = db->counter_size;
    o->both_strands = db->both_strands;
    o->min_count = db->min_count;
    o->max_count = db->max_count;
    o->n_bins = (int32_t)(db->prefix_array_len / db->lut_prefix_array_size);
    o->total_kmers = db->total_kmers;
}

/* U/HelperFunctions.java:232-243 compareByteArray (unsigned lexicographic) */
static int compare_byte_array(const uint8_t *a, const uint8_t *b, int n)
{
    int i;
    for (i = 0; i < n; ++i)
        if (a[i] != b[i]) break;
    if (i == n) return 0;
    return a[i] < b[i] ? -1 : 1;
}

/* D/KMC.java:292-326 getCount, :366-401 getEntry/getSuffixFromEntry/getCountFromEntry.
 * `w` is the already-canonicalised k-mer word (P/GetVariants.java:222-223).
 * Returns the Java int count (may be negative for 4-byte counters >= 2^31). */
static int32_t kmc_get_count_word(const orc_kmc *db, const kword *w)
{
    int k = db->kmer_length, P = db->lut_prefix_length;
    int32_t signature = kmer_signature(w, k, db->signature_length, db->norm);
    int32_t prefix = (int32_t)extract_int_from_bits(w, 0, P);                    /* D/Kmer.java:123-128 */
    uint8_t suffix[ORC_MAX_K / 4 + 1];
    kmer_extract_suffix(w, k, P, suffix);
    int nsb = db->suffix_length / 4;
    int64_t signature_index = (int64_t)db->signature_map[signature] * db->lut_prefix_array_size; /* :300 */
    int64_t start, end;
    if (signature_index + prefix < 0 || signature_index + prefix >= db->prefix_array_len) return 0; /* Java would throw; not reachable for well-formed DBs */
    start = (int64_t)db->prefix_array[signature_index + prefix];                 /* :301 */
    if (signature_index + prefix + 1 >= db->prefix_array_len) end = db->total_kmers - 1; /* :302-304 */
    else end = (int64_t)db->prefix_array[signature_index + prefix + 1] - 1;      /* :305-307 */
    while (start <= end) {                                                       /* :310-323 */
        int64_t mid = (start + end) / 2;
        const uint8_t *entry = db->suf + mid * db->record_size;
        int c = compare_byte_array(suffix, entry, nsb);
        if (c < 0) end = mid - 1;
        else if (c > 0) start = mid + 1;
        else {
            int32_t count = 0;                                                   /* :395-401 */
            for (int i = 0; i < db->counter_size; i++) count |= (int32_t)((uint32_t)entry[nsb + i] << (i * 8));
            return count;
        }
    }
    return 0;
}

/* count of one k-mer given as ASCII (upper-case ACGT), canonicalised per the DB's flag */
int32_t orc_kmc_count(const orc_kmc *db, const char *kmer_ascii)
{
    kword w = kmer_to_long(kmer_ascii, db->kmer_length);
    w = kmer_canonical(w, db->kmer_length, db->both_strands);
    return kmc_get_count_word(db, &w);
}

int32_t orc_kmc_signature(const orc_kmc *db, const char *kmer_ascii)
{
    kword w = kmer_to_long(kmer_ascii, db->kmer_length);
    w = kmer_canonical(w, db->kmer_length, db->both_strands);
    return kmer_signature(&w, db->kmer_length, db->signature_length, db->norm);
}

/* ------------------------------------------------------------------ */
/* D/FastaIndex.java:122-182 getSequence                              */
/* ------------------------------------------------------------------ */
/* raw = the bytes mmapped for one sequence (from its .faidx offset up to the next
 * sequence's offset or EOF, D/FastaIndex.java:54-68).  Returns ORC_ERR_FATAL where
 * the reference logs an error (invalid range, or running off the mapped buffer). */
int orc_get_sequence(const uint8_t *raw, int64_t raw_len, int32_t line_bases, int32_t line_width,
                     int32_t seq_len, int32_t start, int32_t length, char *out)
{
    int32_t end = start + length;
    if (start < 0 || end > seq_len || start >= end) return ORC_ERR_FATAL;        /* :132-135 */
    int32_t start_line = start / line_bases;                                     /* :147 */
    int32_t start_line_base_index = start % line_bases;                          /* :148 */
    int64_t pos = (int64_t)start_line * line_width + start_line_base_index;      /* :151 */
    if (pos > raw_len) return ORC_ERR_FATAL;                                     /* buf.position() throws */
    int32_t to_extract = end - start;
    int32_t o = 0;
    while (to_extract > 0) {                                                     /* :157-174 */
        int32_t remaining_in_line = line_bases - (start_line_base_index % line_bases);
        int32_t to_read = to_extract < remaining_in_line ? to_extract : remaining_in_line;
        for (int32_t i = 0; i < to_read; i++) {
            if (pos >= raw_len) return ORC_ERR_FATAL;                            /* BufferUnderflow */
            out[o++] = (char)raw[pos++];
        }
        pos += (line_width - line_bases);                                        /* :169 */
        if (pos > raw_len) return ORC_ERR_FATAL;                                 /* :175-177 (Q9: no trailing newline) */
        to_extract -= to_read;
        start_line_base_index = 0;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* P/GetVariants.java:278-352 getWindows ("window" feature)           */
/* ------------------------------------------------------------------ */
/* Returns the number of windows; fills starts/ends up to cap. */
int64_t orc_windows_fixed(int32_t seq_len, int32_t window_size, int32_t step, int32_t k,
                          int32_t *starts, int32_t *ends, int64_t cap)
{
    int64_t n = 0;
    if (step > 0) {                                                              /* :295-306 sliding */
        int32_t last_pos = 0;
        while (last_pos < seq_len) {
            int32_t start = last_pos;
            int32_t end = start + window_size < seq_len ? start + window_size : seq_len;
            if (end - start >= k) {
                if (n < cap) { starts[n] = start; ends[n] = end; }
                n++;
            }
            last_pos += step;
        }
    } else {                                                                     /* :307-320 tiling */
        if (window_size <= k - 1) return ORC_ERR_ARG; /* Q10: the reference loop never terminates */
        int32_t last_end = 0;
        while (last_end < seq_len) {
            int32_t start = last_end - k + 1 > 0 ? last_end - k + 1 : 0;
            int32_t end = start + window_size < seq_len ? start + window_size : seq_len;
            if (end - start >= k) {
                if (n < cap) { starts[n] = start; ends[n] = end; }
                n++;
            }
            last_end = end;
        }
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* P/GetVariants.java:202-273 processWindow + getDistance             */
/* D/Fasta.java:90-134 getKmersList, :140-167 getEffectiveATGCCount   */
/* D/Data.java:70-107 update + computeScore                           */
/* ------------------------------------------------------------------ */

/* Same layout as kcf_result_t in include/kcf_b200.h (48 bytes). */
typedef struct {
    int32_t total_kmers, eff_len, obs, variations, inner, left, right, _pad;
    int64_t kmer_count_sum;
    double score;
} orc_result;

/* P/GetVariants.java:267-273 */
static int32_t get_distance(int32_t k, int32_t gap_size)
{
    int32_t distance = gap_size - (k - 1);
    if (distance <= 0) distance = abs(distance + 1);
    return distance;
}

static inline int is_valid_base_upper(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; } /* D/Fasta.java:132-134 */
static inline char to_upper_ascii(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }       /* Character.toUpperCase, ASCII */

/* D/Fasta.java:140-167 */
static int32_t effective_atgc_count(const char *seq, int32_t n, int32_t k)
{
    int32_t count = 0, stretch = 0;
    for (int32_t i = 0; i < n; i++) {
        char b = to_upper_ascii(seq[i]);
        if (b == 'A' || b == 'T' || b == 'G' || b == 'C') stretch++;
        else {
            if (stretch >= k) count += stretch;
            stretch = 0;
        }
    }
    if (stretch >= k) count += stretch;
    return count;
}

/* D/Data.java:95-107 computeScore; weights = {wi, wt, wr} (P/GetVariants.java:388-390).
 * Returns ORC_ERR_FATAL when the reference would Logger.error ("Weights should sum to 1.0"). */
int orc_compute_score(int32_t obs, int32_t total, int32_t eff, int32_t inner, int32_t left, int32_t right,
                      const double *w, double *score)
{
    if (obs == 0 || total == 0 || eff == 0) { *score = 0; return ORC_OK; }
    int rc = ORC_OK;
    if (w[0] + w[1] + w[2] != 1.0) rc = ORC_ERR_FATAL;
    int32_t tail = left + right;
    *score = ((w[2] * ((double)obs / total))
              + (w[0] * (1.0 - ((double)inner / eff)))
              + (w[1] * (1.0 - ((double)tail / eff)))) * 100.0;
    return rc;
}

/* One window given its already-extracted sequence string (not NUL-terminated).
 * hits_out (optional, may be NULL): per emitted k-mer the Java int count, capacity n. */
int orc_process_window(const orc_kmc *db, const char *seq, int32_t n, int32_t min_kmer_count,
                       const double *weights, orc_result *res, int32_t *counts_out)
{
    int k = db->kmer_length;
    int32_t total = 0, obs = 0, variation = 0, inner = 0, gap = 0, left = 0, right = 0;
    int is_tail = 1;
    int64_t kmer_count_sum = 0;
    char chars[ORC_MAX_K + 8];
    int have_kmer = 0;
    int32_t valid_start = 0;
    memset(chars, 0, sizeof chars);
    /* D/Fasta.java:96-124 fused with P/GetVariants.java:220-245 (the list is consumed in order) */
    for (int32_t i = 0; i < n; i++) {
        char base = to_upper_ascii(seq[i]);
        if (!is_valid_base_upper(base)) {
            valid_start = i + 1;
            have_kmer = 0;
            continue;
        }
        int32_t offset = i - valid_start;
        if (offset < k) {
            chars[offset] = base;
            if (offset == k - 1) have_kmer = 1;
        } else if (have_kmer) {
            memmove(chars, chars + 1, (size_t)(k - 1));
            chars[k - 1] = base;
        }
        if (!have_kmer) continue;
        kword w = kmer_to_long(chars, k);                          /* new Kmer(char[]) per position, D/Fasta.java:108,118 */
        w = kmer_canonical(w, k, db->both_strands);                /* new Kmer(k, kmc.isBothStrands()), P/GetVariants.java:222 */
        int32_t cnt = kmc_get_count_word(db, &w);                  /* :223 */
        if (counts_out) counts_out[total] = cnt;
        total++;                                                   /* :221 */
        if (cnt >= min_kmer_count) {                               /* :224 */
            kmer_count_sum += cnt;
            obs++;
            if (gap > 0) {
                variation++;
                if (is_tail) left += gap;
                else inner += get_distance(k, gap);
            }
            is_tail = 0;
            gap = 0;
        } else {
            gap++;
        }
    }
    if (gap > 0) {                                                 /* :247-251 */
        variation++;
        right += gap;
    }
    res->total_kmers = total;
    res->eff_len = effective_atgc_count(seq, n, k);                /* :256 */
    res->obs = obs;
    res->variations = variation;
    res->inner = inner;
    res->left = left;
    res->right = right;
    res->_pad = 0;
    res->kmer_count_sum = kmer_count_sum;
    return orc_compute_score(obs, total, res->eff_len, inner, left, right, weights, &res->score);
}

/* ------------------------------------------------------------------ */
/* Whole-job driver with the C-ABI's window/segment description        */
/* (mirrors kcf_screen in include/kcf_b200.h; used for parity tests    */
/* and as the threaded CPU baseline, P/GetVariants.java:129-151)       */
/* ------------------------------------------------------------------ */

typedef struct { uint32_t first_seg, n_segs; } orc_window;
typedef struct { int32_t seq_id, start0, len; } orc_segment;
typedef struct {
    const uint8_t *raw; int64_t raw_len; int32_t line_bases, line_width, seq_len, _pad;
} orc_seq;

typedef struct {
    const orc_kmc *db;
    const orc_seq *seqs; int32_t n_seqs;
    const orc_window *wins; int64_t n_wins;
    const orc_segment *segs;
    int32_t min_count;
    const double *weights;
    orc_result *out;
    volatile int64_t next;
    volatile int status;
    pthread_mutex_t mu;
} screen_job;

static void *screen_worker(void *arg)
{
    screen_job *job = (screen_job *)arg;
    size_t cap = 1 << 16;
    char *buf = (char *)malloc(cap);
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int64_t w = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (w >= job->n_wins) break;
        const orc_window *win = &job->wins[w];
        size_t need = 0;
        for (uint32_t s = 0; s < win->n_segs; s++) need += (size_t)job->segs[win->first_seg + s].len;
        if (need + 1 > cap) { cap = need + 1; buf = (char *)realloc(buf, cap); }
        size_t o = 0;
        int rc = ORC_OK;
        /* gene/transcript: concatenate the merged loci in order, D/GTF.java:240-244;
         * fixed window: one segment, D/Window.java:224-226 */
        for (uint32_t s = 0; s < win->n_segs && rc == ORC_OK; s++) {
            const orc_segment *sg = &job->segs[win->first_seg + s];
            if (sg->seq_id < 0 || sg->seq_id >= job->n_seqs) { rc = ORC_ERR_ARG; break; }
            const orc_seq *sq = &job->seqs[sg->seq_id];
            rc = orc_get_sequence(sq->raw, sq->raw_len, sq->line_bases, sq->line_width, sq->seq_len,
                                  sg->start0, sg->len, buf + o);
            o += (size_t)sg->len;
        }
        if (rc == ORC_OK && win->n_segs == 0) rc = ORC_ERR_FATAL; /* fasta == null, P/GetVariants.java:213-216 */
        if (rc == ORC_OK) rc = orc_process_window(job->db, buf, (int32_t)o, job->min_count, job->weights, &job->out[w], NULL);
        if (rc != ORC_OK) job->status = rc;
    }
    free(buf);
    return NULL;
}

int orc_screen(const orc_kmc *db, const orc_seq *seqs, int32_t n_seqs,
               const orc_window *wins, int64_t n_wins, const orc_segment *segs,
               int32_t min_count, const double *weights, int32_t n_threads, orc_result *out)
{
    if (min_count < 1 || n_threads < 1) return ORC_ERR_FATAL; /* P/GetVariants.java:379-385 validateCMD */
    screen_job job;
    memset(&job, 0, sizeof job);
    job.db = db; job.seqs = seqs; job.n_seqs = n_seqs; job.wins = wins; job.n_wins = n_wins; job.segs = segs;
    job.min_count = min_count; job.weights = weights; job.out = out; job.next = 0; job.status = ORC_OK;
    pthread_mutex_init(&job.mu, NULL);
    if (n_threads == 1) {
        screen_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
        for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, screen_worker, &job);
        for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&job.mu);
    return job.status;
}

/* gap state machine alone over a 0/1 hit vector (KATs ii/iii in SURVEY §8c) */
void orc_gap_machine(const uint8_t *hits, int32_t n, int32_t k, int32_t *out6 /* total,obs,var,inner,left,right */)
{
    int32_t total = 0, obs = 0, variation = 0, inner = 0, gap = 0, left = 0, right = 0;
    int is_tail = 1;
    for (int32_t i = 0; i < n; i++) {
        total++;
        if (hits[i]) {
            obs++;
            if (gap > 0) {
                variation++;
                if (is_tail) left += gap;
                else inner += get_distance(k, gap);
            }
            is_tail = 0;
            gap = 0;
        } else gap++;
    }
    if (gap > 0) { variation++; right += gap; }
    out6[0] = total; out6[1] = obs; out6[2] = variation; out6[3] = inner; out6[4] = left; out6[5] = right;
}
