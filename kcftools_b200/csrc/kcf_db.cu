// kcf_db.cu — K1: KMC 0x200 database ingest and re-layout into the HBM-resident table.
//
// Replaces `new KMC(prefix, inMemory)` (KMC.java:56-189).  The file layout is parsed exactly as
// readPrefixFile does (KMC.java:107-168); the records of .kmc_suf are streamed through pinned
// staging buffers and re-keyed on the device.  A record is inserted only if a reference lookup
// could return it, i.e. if it sits in the (bin, prefix) range that getCount (KMC.java:292-326)
// would search for that very k-mer: its signature (Kmer.java:105-118, Signature.java:23-37) must
// map to the bin that holds it and, for a both-strands database, it must be its own canonical
// form (Kmer.java:72-79).  That proof is done once per record here, so the screening kernel needs
// neither signatures nor the per-bin LUTs.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "kcf_internal.cuh"
#include "kcf_lookup.cuh"
#include <atomic>
#include <thread>
#include <vector>

// The records reach the device through a ring of pinned staging buffers kept by the context (pinned allocations cost
// tens of milliseconds; a getVariations run opens one database, a cohort run many).  Filling a slot is a host memcpy out of
// the page cache (or the caller's array): KCF_INGEST_FILLERS threads each fill whole chunks, KCF_INGEST_SLOTS chunks ahead
// of the device, so the copy engine never waits for the host (one thread alone moves ~10 GB/s, PCIe 55).
#ifndef KCF_INGEST_CHUNK_BYTES
#define KCF_INGEST_CHUNK_BYTES (16ULL << 20)
#endif
#define KCF_INGEST_SLOTS 8
#define KCF_INGEST_FILLERS 8
static_assert(KCF_INGEST_SLOTS <= 8 && KCF_INGEST_FILLERS <= KCF_INGEST_SLOTS, "the context holds 8 staging slots; a filler owns a slot while it fills it");

// ---- Signature.java:23-95 on the device: one thread per m-mer -----------------------------------
__device__ __forceinline__ bool sig_allowed(uint32_t s, int L)
{
    if ((s & 0x3F) == 0x3F) return false; // TTT suffix
    if ((s & 0x3F) == 0x3B) return false; // TGT suffix
    if ((s & 0x3C) == 0x3C) return false; // TG* suffix
    for (int j = 0; j < L - 3; ++j) {
        if ((s & 0xF) == 0) return false; // AA inside
        s >>= 2;
    }
    if (s == 0) return false;    // AAA prefix
    if (s == 0x04) return false; // ACA prefix
    if ((s & 0xF) == 0) return false; // *AA prefix
    return true;
}

__global__ void kcf_norm_kernel(int L, uint32_t *__restrict__ norm)
{
    uint32_t special = 1u << (2 * L);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= special) return;
    uint32_t rev = 0, t = i;
    for (int j = 0; j < L; ++j) {
        rev = (rev << 2) | ((~t) & 3u);
        t >>= 2;
    }
    uint32_t a = sig_allowed(i, L) ? i : special;
    uint32_t b = sig_allowed(rev, L) ? rev : special;
    norm[i] = min(a, b);
}

// bit i = Signature.isAllowed(i): 32 m-mers per thread.  4^9 bits = 32 KB: the ingest kernel keeps it in shared memory and
// computes norm(m) = min(allowed(m) ? m : 4^L, allowed(rc m) ? rc m : 4^L) (Signature.java:28-35) from it, instead of
// gathering 23 words per record from the 1 MB norm table through L2.
__global__ void kcf_allowed_kernel(int L, uint32_t *__restrict__ bits)
{
    const uint32_t n_words = (1u << (2 * L)) >> 5;
    const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= (n_words ? n_words : 1u)) return;
    uint32_t v = 0;
    for (uint32_t b = 0; b < 32; ++b) {
        const uint32_t i = wi * 32 + b;
        if (i < (1u << (2 * L)) && sig_allowed(i, L)) v |= 1u << b;
    }
    bits[wi] = v;
}

// signature of a k-mer (the reference's value): min norm over its k-L+1 m-mers, first base most significant
// (Kmer.java:105-118), both strands of every m-mer rolled along
template <typename BitsPtr>
__device__ __forceinline__ uint32_t kcf_signature_from_bits(uint64_t kmer, uint32_t k, uint32_t L, BitsPtr A)
{
    const uint32_t special = 1u << (2 * L), mmask = special - 1u;
    uint32_t m = (uint32_t)(kmer >> (2 * (k - L))) & mmask, rc = 0;
    for (uint32_t j = 0, t = m; j < L; ++j, t >>= 2) rc = (rc << 2) | ((~t) & 3u);
    uint32_t sig = 0xFFFFFFFFu;
    for (uint32_t j = 0;; ++j) {
        const uint32_t a = ((A[m >> 5] >> (m & 31u)) & 1u) ? m : special;
        const uint32_t b = ((A[rc >> 5] >> (rc & 31u)) & 1u) ? rc : special;
        sig = min(sig, min(a, b));
        if (j == k - L) break;
        const uint32_t nb = (uint32_t)(kmer >> (2 * (k - L - j - 1))) & 3u; // the base entering on the right
        m = ((m << 2) | nb) & mmask;
        rc = (rc >> 2) | ((3u - nb) << (2 * (L - 1)));
    }
    return sig;
}

// LUT must be non-decreasing and bounded by total (else the reference reads garbage ranges)
__global__ void kcf_lut_check_kernel(const uint64_t *__restrict__ lut, uint64_t n, uint64_t total, uint32_t *flags)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = lut[i];
    bool bad = v > total;
    if (i + 1 < n && lut[i + 1] < v) bad = true;
    if (bad) atomicOr(&flags[FLAG_LUT_BAD], 1u);
}

// bound[c] = first LUT index whose value exceeds record rec0 + 256 c (c = 0 .. n_cta; the last one bounds the chunk's last record)
__global__ void kcf_lut_bounds_kernel(const uint64_t *__restrict__ lut, uint64_t lut_len, uint64_t rec0, uint64_t n_rec, uint32_t n_cta,
                                      uint64_t *__restrict__ bound)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cta) return;
    const uint64_t ib = rec0 + min((uint64_t)c * 256, n_rec - 1);
    uint64_t lo = 0, hi = lut_len;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (lut[mid] <= ib) lo = mid + 1;
        else hi = mid;
    }
    bound[c] = lo;
}

struct KcfStaged;
struct KcfIngestParams {
    const uint8_t *rec;      // staged records of this chunk
    uint64_t rec0;           // global index of the first staged record
    uint64_t n_rec;
    uint8_t prev[16];        // record rec0-1 (valid when rec0 > 0)
    const uint64_t *lut;
    uint64_t lut_len;
    const uint64_t *cta_bound; // per CTA of this chunk (+1): first LUT index whose value exceeds the CTA's first record
    const uint32_t *sigmap;
    const uint32_t *allowed;  // Signature.isAllowed bitmap, 4^L bits
    uint32_t n_groups;        // groups of 256 consecutive records in this chunk
    unsigned int *group_counter; // zeroed per chunk: next group to take
    struct KcfStaged *staged;    // per record of the chunk: proof kernel -> insert kernel (k <= 32)
    uint32_t P, L, nsb, cs, rec_size;
    uint8_t *table;
    KcfStashEntry *ovf;      // overflow list
    uint64_t ovf_cap;
    unsigned long long *counters; // [0] inserted, [1] unreachable, [2] overflow (also the list's cursor), [3] owned by another rank
    uint32_t part_rank, part_world;
    uint32_t *flags;
};

// record in the home line's filter that `key` lives outside it (bits are cleared: the filter is stored inverted)
__device__ __forceinline__ void kcf_filter_add(uint8_t *home_line, uint64_t key, const KcfTableGeom &g)
{
    const uint32_t h = kcf_filter_hash(key);
    uint32_t *f = reinterpret_cast<uint32_t *>(home_line + g.foff);
    if (g.fbits == 64) {
        const uint32_t b1 = h >> 26, b2 = (h >> 20) & 63u;
        atomicAnd(f + (b1 >> 5), ~(1u << (b1 & 31u)));
        atomicAnd(f + (b2 >> 5), ~(1u << (b2 & 31u)));
    } else {
        atomicAnd(f, ~((1u << (h >> 27)) | (1u << ((h >> 22) & 31u))));
    }
}

// A record between its proof and its insert (k <= 32): see kcf_ingest_kernel / kcf_insert_kernel.
struct KcfStaged {
    uint64_t key;   // table key
    uint32_t home;  // global home line; 0xFFFFFFFF: nothing to insert
    uint32_t count;
};

// 8 record bytes starting at byte offset `off` of the staged chunk (rec_size <= 8 fast path), little endian
__device__ __forceinline__ uint64_t kcf_load_rec8(const uint8_t *rec, uint64_t off)
{
    const uint64_t *q = reinterpret_cast<const uint64_t *>(rec + (off & ~7ULL));
    const uint32_t sh = (uint32_t)(off & 7ULL) * 8u;
    const uint64_t lo = __ldg(q);
    return sh ? (lo >> sh) | (__ldg(q + 1) << (64u - sh)) : lo;
}

// proof of reachability of one record (k <= 32).  Returns what becomes of it — 0 to be inserted (or resident without storage
// when the counters have 0 bytes), 1 unreachable, 3 homed in another rank's slice — and fills `out`.
template <typename BitsPtr>
__device__ __forceinline__ int kcf_prove_one(const KcfIngestParams &p, const KcfTableGeom &g, uint64_t t, uint32_t group, BitsPtr allowed, KcfStaged &out)
{
    out.key = 0;
    out.home = 0xFFFFFFFFu;
    out.count = 0;
    // (bin, prefix) range holding record i: last idx with lut[idx] <= i.  The LUT is monotone, so the answers of a group's
    // 256 consecutive records lie between those of its first record and of the next group's first record, which
    // kcf_lut_bounds_kernel searched beforehand (23 dependent loads for 512 bins x 4^7 prefixes, once per group instead of
    // once per record); what is left here is a search over the few entries in between.
    const uint64_t i = p.rec0 + t;
    uint64_t lo = p.cta_bound[group], hi = p.cta_bound[group + 1];
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (p.lut[mid] <= i) lo = mid + 1;
        else hi = mid;
    }
    if (lo == 0) return 1;
    const uint64_t idx = lo - 1;
    const uint32_t bin = (uint32_t)(idx >> (2 * p.P));
    const uint64_t prefix = idx & ((1ULL << (2 * p.P)) - 1);
    uint64_t suffix = 0, ps = 0;
    uint32_t count = 0;
    const bool need_prev = i > p.lut[idx]; // strict ascending order inside the range, which the reference's binary search presumes
    if (p.rec_size <= 8) { // the usual geometry (7 suffix bytes + 1 counter byte): one 64-bit load per record
        const uint64_t v = kcf_load_rec8(p.rec, t * p.rec_size);
        const uint64_t be = ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123); // big-endian suffix bytes (Kmer.java:166)
        suffix = p.nsb ? be >> (8u * (8u - p.nsb)) : 0ULL;
        if (p.cs) count = (uint32_t)((v >> (8u * p.nsb)) & (p.cs >= 4 ? 0xFFFFFFFFULL : ((1ULL << (8u * p.cs)) - 1ULL))); // KMC.java:395-401
        if (need_prev) {
            if (t > 0) {
                const uint64_t pv = kcf_load_rec8(p.rec, (t - 1) * p.rec_size);
                const uint64_t pbe = ((uint64_t)__byte_perm((uint32_t)pv, 0, 0x0123) << 32) | __byte_perm((uint32_t)(pv >> 32), 0, 0x0123);
                ps = p.nsb ? pbe >> (8u * (8u - p.nsb)) : 0ULL;
            } else {
                for (uint32_t j = 0; j < p.nsb; ++j) ps = (ps << 8) | p.prev[j];
            }
        }
    } else {
        const uint8_t *r = p.rec + t * p.rec_size;
        for (uint32_t j = 0; j < p.nsb; ++j) suffix = (suffix << 8) | r[j];
        for (uint32_t j = 0; j < p.cs; ++j) count |= (uint32_t)r[p.nsb + j] << (8 * j);
        if (need_prev) {
            const uint8_t *q = (t > 0) ? (r - p.rec_size) : p.prev;
            for (uint32_t j = 0; j < p.nsb; ++j) ps = (ps << 8) | q[j];
        }
    }
    if (need_prev && ps >= suffix) atomicOr(&p.flags[FLAG_ORDER_BAD], 1u);
    const uint32_t sbits = 8 * p.nsb;
    const uint64_t kmer = (p.P > 0 ? (prefix << sbits) : 0ULL) | suffix;
    if (g.both_strands && kmer > kcf_revcomp(kmer, g.kshift)) return 1; // a query is canonicalised first (GetVariants.java:222)
    if (p.sigmap[kcf_signature_from_bits(kmer, g.k, p.L, allowed)] != bin) return 1; // KMC.java:300
    if (p.cs == 0) return 0; // Q7: a 0-byte counter reads as 0, never a hit; nothing to store
    // proof done in the reference's encoding; from here on the record is its table key (bit planes, kcf_lookup.cuh)
    out.key = kcf_table_key(kmer, g);
    const uint32_t home = kcf_home_line(kcf_minimizer_of_key(out.key, g), g);
    if (p.part_world > 1 && kcf_line_owner(home, g.n_lines, p.part_world) != p.part_rank) return 3; // another rank's slice
    out.home = home;
    out.count = count;
    return 0;
}

// insert of one proven record (k <= 32).  Returns 0 resident in a line, 2 resident in the overflow list (stash).
// store the count of slot s of a line: the slot's bytes belong to the thread that claimed it (plain byte-masked store)
__device__ __forceinline__ void kcf_store_count(uint8_t *line, uint32_t s, uint32_t count, const KcfTableGeom &g)
{
    const uint32_t off = g.coff + g.cw * s;
    if ((off >> 2) == 31u) { // the 13th 1-byte count shares word 31 with the mask, which other threads update atomically:
                             // a plain store there would race with their read-modify-write
        const uint32_t sh = 8 * (off & 3u);
        const uint32_t field = ((1u << (8 * g.cw)) - 1u) << sh;
        atomicAnd(reinterpret_cast<uint32_t *>(line) + 31, ~field | (count << sh)); // cleared out of the all-ones initial image
        return;
    }
    uint8_t *c = line + off;
    if (g.cw == 1) *c = (uint8_t)count;
    else if (g.cw == 2) *reinterpret_cast<uint16_t *>(c) = (uint16_t)count;
    else *reinterpret_cast<uint32_t *>(c) = count;
}

// insert of one proven record (k <= 32) into the lines dmin .. 14 of its home's sequence, then the overflow list.
// Returns 0 resident in a line, 2 resident in the overflow list (stash).
__device__ __noinline__ int kcf_insert_one(const KcfIngestParams &p, const KcfTableGeom &g, uint64_t tkey, uint32_t home, uint32_t count, uint32_t dmin)
{
    uint8_t *home_line = p.table + (uint64_t)kcf_line_wrap(home, 0, g) * KCF_LINE_BYTES;
    uint32_t *home_w31 = reinterpret_cast<uint32_t *>(home_line) + 31;
    if (KCF_KEY_IN_LINES(tkey)) {
        const uint32_t lo = (uint32_t)tkey, hi = (uint32_t)(tkey >> 32);
        for (uint32_t d = dmin; d <= KCF_MAX_DISP; ++d) {
            uint8_t *line = p.table + (uint64_t)kcf_line_wrap(home, d, g) * KCF_LINE_BYTES;
            uint32_t *w = reinterpret_cast<uint32_t *>(line);
            // One look at the line's low key words (four 16-byte loads past L1), then claim the first slot seen empty.
            // Slots fill in order and never change once written, so every slot before the one claimed has been compared
            // with its final content: either in the snapshot, or through the value a failed CAS returns.
            uint32_t snap[16];
            {
                const uint4 *q = reinterpret_cast<const uint4 *>(line);
                const uint4 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), e = __ldcg(q + 3);
                snap[0] = a.x; snap[1] = a.y; snap[2] = a.z; snap[3] = a.w; snap[4] = b.x; snap[5] = b.y; snap[6] = b.z; snap[7] = b.w;
                snap[8] = c.x; snap[9] = c.y; snap[10] = c.z; snap[11] = c.w; snap[12] = e.x; snap[13] = e.y; snap[14] = e.z; snap[15] = e.w;
            }
            bool dup = false;
            uint32_t s0 = g.S;
#pragma unroll
            for (uint32_t s = 0; s < 13; ++s) {
                if (s < g.S && s < s0) {
                    if (snap[s] == lo) dup = true;
                    else if (snap[s] == KCF_EMPTY_LO) s0 = s;
                }
            }
            // a key with the same low word already lives in this line: keep low words unique per line (a probe then has
            // one candidate at most) and move on to the next line
            for (uint32_t s = s0; !dup && s < g.S; ++s) {
                const uint32_t v = atomicCAS(&w[s], KCF_EMPTY_LO, lo); // claiming a slot and publishing the low word are one step
                if (v == KCF_EMPTY_LO) {
                    w[g.S + s] = hi;
                    kcf_store_count(line, s, count, g);
                    // mask bits are cleared out of the all-ones initial image (stored inverted); bits only ever get cleared, so a
                    // bit seen cleared needs no atomic
                    if ((__ldcg(home_w31) >> (16 + d)) & 1u) atomicAnd(home_w31, ~(1u << (16 + d)));
                    if (d > 0) kcf_filter_add(home_line, tkey, g);
                    return 0;
                }
                if (v == lo) dup = true;
            }
        }
    }
    atomicAnd(home_w31, ~(1u << (16 + KCF_STASH_BIT)));
    kcf_filter_add(home_line, tkey, g);
    unsigned long long pos = atomicAdd(&p.counters[2], 1ULL);
    if (pos < p.ovf_cap) {
        p.ovf[pos].key = tkey;
        p.ovf[pos].key_hi = 0;
        p.ovf[pos].meta = (1ULL << 63) | count;
    }
    return 2;
}

// ---- the same for k = 33 .. 64: the reference's value is 2k <= 128 bits (Kmer.java:232-252 keeps it in long[] words, first
// base most significant; canonical = the lexicographically smaller strand, :72-79, 406-414 — for equal lengths the numeric
// order of the right-aligned values), the table key two 64-bit planes, the home line a hash of the whole key
__device__ __forceinline__ unsigned __int128 kcf_revcomp128(unsigned __int128 x, uint32_t k)
{
    const uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
    const unsigned __int128 rev = ((unsigned __int128)kcf_pair_reverse64(~lo) << 64) | kcf_pair_reverse64(~hi);
    return rev >> (128u - 2u * k); // the complemented padding falls off the low end
}

// record in the home line's filter that a 128-bit key lives outside it (32-bit filter; absent when fbits == 0)
__device__ __forceinline__ void kcf_filter_add2(uint8_t *home_line, const KcfKey2 &key, const KcfTableGeom &g)
{
    if (g.fbits == 0) return;
    const uint32_t h = kcf_filter_hash(key.p0 ^ (key.p1 * 0x9E3779B97F4A7C15ULL));
    atomicAnd(reinterpret_cast<uint32_t *>(home_line + g.foff), ~((1u << (h >> 27)) | (1u << ((h >> 22) & 31u))));
}

template <typename BitsPtr>
__device__ __forceinline__ int kcf_ingest_one2(const KcfIngestParams &p, const KcfTableGeom &g, uint64_t t, uint32_t group, BitsPtr allowed)
{
    typedef unsigned __int128 u128;
    const uint64_t i = p.rec0 + t;
    uint64_t lo = p.cta_bound[group], hi = p.cta_bound[group + 1];
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (p.lut[mid] <= i) lo = mid + 1;
        else hi = mid;
    }
    if (lo == 0) return 1;
    const uint64_t idx = lo - 1;
    const uint32_t bin = (uint32_t)(idx >> (2 * p.P));
    const uint64_t prefix = idx & ((1ULL << (2 * p.P)) - 1);
    const uint8_t *r = p.rec + t * p.rec_size;
    u128 suffix = 0;
    for (uint32_t j = 0; j < p.nsb; ++j) suffix = (suffix << 8) | r[j]; // big-endian suffix bytes (Kmer.java:166), up to 16
    uint32_t count = 0;
    for (uint32_t j = 0; j < p.cs; ++j) count |= (uint32_t)r[p.nsb + j] << (8 * j); // KMC.java:395-401
    if (i > p.lut[idx]) { // strict ascending order inside the range
        const uint8_t *q = (t > 0) ? (r - p.rec_size) : p.prev;
        u128 ps = 0;
        for (uint32_t j = 0; j < p.nsb; ++j) ps = (ps << 8) | q[j];
        if (ps >= suffix) atomicOr(&p.flags[FLAG_ORDER_BAD], 1u);
    }
    const uint32_t sbits = 8 * p.nsb;
    const u128 kmer = (p.P > 0 ? ((u128)prefix << sbits) : (u128)0) | suffix;
    bool reachable = true;
    if (g.both_strands && kmer > kcf_revcomp128(kmer, g.k)) reachable = false; // a query is canonicalised first (GetVariants.java:222)
    if (reachable) {
        // signature (Kmer.java:105-118): min norm over the k-L+1 m-mers, both strands of each rolled along
        const uint32_t L = p.L, special = 1u << (2 * L), mmask = special - 1u;
        uint32_t m = (uint32_t)(kmer >> (2 * (g.k - L))) & mmask, rc = 0;
        for (uint32_t j = 0, tt = m; j < L; ++j, tt >>= 2) rc = (rc << 2) | ((~tt) & 3u);
        uint32_t sig = 0xFFFFFFFFu;
        for (uint32_t j = 0;; ++j) {
            const uint32_t a = ((allowed[m >> 5] >> (m & 31u)) & 1u) ? m : special;
            const uint32_t b = ((allowed[rc >> 5] >> (rc & 31u)) & 1u) ? rc : special;
            sig = min(sig, min(a, b));
            if (j == g.k - L) break;
            const uint32_t nb = (uint32_t)(kmer >> (2 * (g.k - L - j - 1))) & 3u;
            m = ((m << 2) | nb) & mmask;
            rc = (rc >> 2) | ((3u - nb) << (2 * (L - 1)));
        }
        if (p.sigmap[sig] != bin) reachable = false; // KMC.java:300
    }
    if (!reachable) return 1;
    if (p.cs == 0) return 0;
    const KcfKey2 key = kcf_table_key2(kmer, g);
    const uint32_t home = kcf_home_line2(key, g);
    if (p.part_world > 1 && kcf_line_owner(home, g.n_lines, p.part_world) != p.part_rank) return 3;
    uint8_t *home_line = p.table + (uint64_t)kcf_line_wrap(home, 0, g) * KCF_LINE_BYTES;
    uint32_t *home_w31 = reinterpret_cast<uint32_t *>(home_line) + 31;
    const uint32_t klo = (uint32_t)key.p0;
    if (klo != KCF_EMPTY_LO) {
        for (uint32_t d = 0; d <= KCF_MAX_DISP; ++d) {
            uint8_t *line = p.table + (uint64_t)kcf_line_wrap(home, d, g) * KCF_LINE_BYTES;
            uint32_t *w = reinterpret_cast<uint32_t *>(line);
            bool dup = false;
            for (uint32_t s2 = 0; !dup && s2 < g.S; ++s2) {
                uint32_t v = __ldcg(w + s2);
                if (v == KCF_EMPTY_LO) v = atomicCAS(&w[s2], KCF_EMPTY_LO, klo); // claiming a slot and publishing the low word are one step
                if (v == KCF_EMPTY_LO) {
                    uint32_t *rw = w + g.S + 3 * s2;
                    rw[0] = (uint32_t)(key.p0 >> 32);
                    rw[1] = (uint32_t)key.p1;
                    rw[2] = (uint32_t)(key.p1 >> 32);
                    const uint32_t off = g.coff + g.cw * s2;
                    const uint32_t sh = 8 * (off & 3u);
                    const uint32_t field = g.cw == 4 ? 0xFFFFFFFFu : (((1u << (8 * g.cw)) - 1u) << sh);
                    atomicAnd(w + (off >> 2), ~field | (count << sh));
                    atomicAnd(home_w31, ~(1u << (16 + d)));
                    if (d > 0) kcf_filter_add2(home_line, key, g);
                    return 0;
                }
                if (v == klo) dup = true; // low words stay unique per line: try the next line
            }
        }
    }
    atomicAnd(home_w31, ~(1u << (16 + KCF_STASH_BIT)));
    kcf_filter_add2(home_line, key, g);
    unsigned long long pos = atomicAdd(&p.counters[2], 1ULL);
    if (pos < p.ovf_cap) {
        p.ovf[pos].key = key.p0;
        p.ovf[pos].key_hi = key.p1;
        p.ovf[pos].meta = (1ULL << 63) | count;
    }
    return 2;
}

// The database load runs as two kernels per chunk (k <= 32).  The proof (LUT range, order, canonical form, signature, table
// key, minimizer: ~1,800 instructions per record, shared-memory lookups, hardly any global traffic) and the insert (a random
// 128-byte line, one CAS, two stores: latency, hardly any arithmetic) have opposite needs.  Measured on C2 (16 MB chunks of
// 2.4e6 records, profiles/README.md "ingest"): one kernel doing both per thread 0.61 ms per chunk at a quarter of the issue
// slots; the same with the insert delayed by one group behind a bulk L2 prefetch 0.57 ms; proof 0.14 ms + insert 0.29 ms as
// two kernels, the insert on its own stream so that it runs under the next chunk's proof.
//
// Proof kernel: CTAs take the chunk's groups of 256 consecutive records from a counter (no tail); the allowed-m-mer bitmap
// is loaded into shared memory once per CTA (SMEM_BITS: 4^L bits fit, i.e. L <= 9); the per-record outcomes are counted in
// shared memory and reach the global counters with one atomic per CTA and outcome.  KW = 1: the proven records go to
// `staged` for the insert kernel; KW = 2 (k > 32): proof and insert in one step.
template <bool SMEM_BITS, int KW>
__global__ void __launch_bounds__(256) kcf_ingest_kernel(KcfIngestParams p, KcfTableGeom g)
{
    extern __shared__ uint32_t s_allowed[];
    __shared__ unsigned int s_cnt[4];
    __shared__ unsigned int s_group;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    if (SMEM_BITS) {
        const uint32_t n_words = (1u << (2 * p.L)) >> 5;
        for (uint32_t i = threadIdx.x; i < (n_words ? n_words : 1u); i += blockDim.x) s_allowed[i] = p.allowed[i];
    }
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_group = atomicAdd(p.group_counter, 1u);
        __syncthreads();
        const uint32_t group = s_group;
        if (group >= p.n_groups) break;
        const uint64_t t = (uint64_t)group * 256 + threadIdx.x;
        if (t >= p.n_rec) continue;
        int what;
        if (KW == 2) {
            what = SMEM_BITS ? kcf_ingest_one2(p, g, t, group, (const uint32_t *)s_allowed) : kcf_ingest_one2(p, g, t, group, p.allowed);
        } else {
            KcfStaged st;
            what = SMEM_BITS ? kcf_prove_one(p, g, t, group, (const uint32_t *)s_allowed, st) : kcf_prove_one(p, g, t, group, p.allowed, st);
            p.staged[t] = st;
            if (st.home != 0xFFFFFFFFu) what = 2; // counted by the insert kernel
        }
        if (what != 2) atomicAdd(&s_cnt[what], 1u); // the overflow list counts its own entries
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&p.counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// one 32-byte sector past L1 with a single 256-bit load
__device__ __forceinline__ void kcf_ld_sector_cg(const uint8_t *p, uint32_t (&w)[8])
{
    uint64_t a, b, c, d;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32);
    w[4] = (uint32_t)c; w[5] = (uint32_t)(c >> 32); w[6] = (uint32_t)d; w[7] = (uint32_t)(d >> 32);
}

// Insert kernel (k <= 32): one thread per proven record.  The home line is fetched the way this part serves a random line
// best (profiles/README.md, pattern T3): the four lanes of a quad read its four sectors with ONE instruction — four rounds,
// one per lane of the quad — so the whole line, mask word included, costs one DRAM access; asking for the first two sectors
// and touching the fourth afterwards (count byte, mask bit) made every insert two (0.45 ms per chunk instead of 0.29).
// The owner gets the S low words and the mask word by shuffles, claims the first empty slot by CAS, and writes high word and
// count; the mask atomic is skipped when the snapshot shows the bit already cleared.  Overflow lines go the
// thread-at-a-time way.
__global__ void __launch_bounds__(256) kcf_insert_kernel(KcfIngestParams p, KcfTableGeom g)
{
    __shared__ unsigned int s_ins;
    if (threadIdx.x == 0) s_ins = 0;
    __syncthreads();
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, qb = lane & ~3u, sub = lane & 3u;
    KcfStaged st;
    st.key = 0;
    st.home = 0xFFFFFFFFu;
    st.count = 0;
    if (t < p.n_rec) st = p.staged[t];
    const bool active = st.home != 0xFFFFFFFFu;
    uint32_t snap[13], w31 = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        const uint32_t src = qb + j;
        const uint32_t hj = __shfl_sync(0xffffffffu, st.home, src);
        uint32_t sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (hj != 0xFFFFFFFFu) kcf_ld_sector_cg(p.table + (uint64_t)kcf_line_wrap(hj, 0, g) * KCF_LINE_BYTES + 32u * sub, sec);
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) { // words 0 .. 7 from the quad's lane 0, 8 .. 12 from lane 1, 31 from lane 3
            const uint32_t v0 = __shfl_sync(0xffffffffu, sec[i], qb);
            const uint32_t v1 = i < 5 ? __shfl_sync(0xffffffffu, sec[i], qb + 1) : 0u;
            if (lane == src) {
                snap[i] = v0;
                if (i < 5) snap[8 + i] = v1;
            }
        }
        const uint32_t v3 = __shfl_sync(0xffffffffu, sec[7], qb + 3);
        if (lane == src) w31 = v3;
    }
    if (active) {
        bool placed = false;
        if (KCF_KEY_IN_LINES(st.key)) {
            const uint32_t lo = (uint32_t)st.key, hi = (uint32_t)(st.key >> 32);
            uint8_t *line = p.table + (uint64_t)kcf_line_wrap(st.home, 0, g) * KCF_LINE_BYTES;
            uint32_t *w = reinterpret_cast<uint32_t *>(line);
            bool dup = false;
            uint32_t s0 = g.S;
#pragma unroll
            for (uint32_t s2 = 0; s2 < 13; ++s2) {
                if (s2 < g.S && s2 < s0) {
                    if (snap[s2] == lo) dup = true;
                    else if (snap[s2] == KCF_EMPTY_LO) s0 = s2;
                }
            }
            for (uint32_t s2 = s0; !dup && s2 < g.S; ++s2) {
                const uint32_t v = atomicCAS(&w[s2], KCF_EMPTY_LO, lo); // claiming a slot and publishing the low word are one step
                if (v == KCF_EMPTY_LO) {
                    w[g.S + s2] = hi;
                    kcf_store_count(line, s2, st.count, g);
                    if ((w31 >> 16) & 1u) atomicAnd(w + 31, ~(1u << 16)); // "a key homed here lives here" (stored inverted: 1 = not yet)
                    placed = true;
                    break;
                }
                if (v == lo) dup = true;
            }
        }
        if (placed || kcf_insert_one(p, g, st.key, st.home, st.count, 1) == 0) atomicAdd(&s_ins, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_ins) atomicAdd(&p.counters[0], (unsigned long long)s_ins);
}

__global__ void kcf_stash_build_kernel(const KcfStashEntry *__restrict__ ovf, uint64_t n, KcfStashEntry *stash, KcfTableGeom g)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    KcfStashEntry e = ovf[t];
    uint64_t i = kcf_mix64(e.key ^ (e.key_hi * 0x9E3779B97F4A7C15ULL));
    for (uint64_t k = 0;; ++k) {
        KcfStashEntry *s = &stash[(i + k) & g.stash_mask];
        if (atomicCAS((unsigned long long *)&s->meta, 0ULL, (unsigned long long)e.meta) == 0ULL) {
            s->key = e.key; // keys are distinct; nobody reads them before the build kernel ends
            s->key_hi = e.key_hi;
            return;
        }
    }
}

// ---- KMC.getCount for a batch of ASCII k-mers (parity helper) ----------------------------------
__global__ void kcf_count_kernel(const char *__restrict__ ascii, uint64_t n, const uint8_t *__restrict__ table,
                                 const KcfStashEntry *__restrict__ stash, KcfTableGeom g, int32_t *__restrict__ out)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const char *s = ascii + t * g.k;
    unsigned __int128 v = 0;
    bool ok = true;
    for (uint32_t j = 0; j < g.k; ++j) {
        uint32_t b = (uint8_t)s[j];
        uint32_t u = b & 0xDFu;
        ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
        v = (v << 2) | (((b >> 1) & 3u) ^ ((b >> 2) & 1u));
    }
    if (!ok) {
        out[t] = 0;
        return;
    }
    // the table keys are strand symmetric for a both-strands database: no canonicalisation needed here
    if (g.kw == 2) {
        const KcfKey2 key = kcf_table_key2(v, g);
        out[t] = (int32_t)kcf_lookup2(table, stash, g, key, kcf_home_line2(key, g));
    } else {
        out[t] = (int32_t)kcf_lookup(table, stash, g, kcf_table_key((uint64_t)v, g));
    }
}

// ---- host side ----------------------------------------------------------------------------------
static uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd_u64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

static int floor_log2_u64(uint64_t v)
{
    int r = -1;
    while (v) { v >>= 1; ++r; }
    return r;
}

// overflow lines per home line for a table of load factor lf holding (a slice of) a database of n_total records
static double kcf_overflow_share(double lf, uint64_t n_total)
{
    const double colliding = 1.0 - std::exp(-(double)n_total / 6.0 / 4294967296.0);
    return std::max(0.125, 0.55 * lf * lf) + colliding * 0.35 * lf / 0.6;
}

extern "C" int kcf_set_load_factor(kcf_ctx *ctx, double lf)
{
    if (!ctx) return KCF_ERR_ARG;
    if (lf < 0.0 || lf > 0.9) return kcf_fail(ctx, KCF_ERR_ARG, "load factor must be in (0, 0.9], or 0 for automatic");
    ctx->load_factor = lf;
    return KCF_OK;
}

extern "C" int kcf_set_minimizer_length(kcf_ctx *ctx, int m)
{
    if (!ctx) return KCF_ERR_ARG;
    if (m < 0 || m > 24) return kcf_fail(ctx, KCF_ERR_ARG, "minimizer length must be 0 (auto) or 1..24");
    ctx->minimizer_len = m;
    return KCF_OK;
}

extern "C" int kcf_set_partition(kcf_ctx *ctx, int rank, int world)
{
    if (!ctx) return KCF_ERR_ARG;
    if (world < 1 || rank < 0 || rank >= world) return kcf_fail(ctx, KCF_ERR_ARG, "partition %d of %d", rank, world);
    ctx->part_rank = rank;
    ctx->part_world = world;
    return KCF_OK;
}

extern "C" int kcf_db_open_mem(kcf_ctx *ctx, const uint8_t *pre, uint64_t pre_len, const uint8_t *suf, uint64_t suf_len,
                               int placement, kcf_db **out)
{
    if (!ctx || !pre || !suf || !out) return KCF_ERR_ARG;
    *out = nullptr;
    if (placement != 0 && placement != 1) return kcf_fail(ctx, KCF_ERR_ARG, "placement %d (0 = whole database, 1 = this context's slice)", placement);
    const uint32_t part_world = placement == 1 ? (uint32_t)ctx->part_world : 1u, part_rank = placement == 1 ? (uint32_t)ctx->part_rank : 0u;
    auto t0 = std::chrono::steady_clock::now();
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    // --- KMC.java:107-168 readPrefixFile ---
    if (pre_len < 16) return kcf_fail(ctx, KCF_ERR_DB_FORMAT, ".kmc_pre too short");
    const uint64_t file_size = pre_len;
    const uint32_t header_offset = rd_u32(pre + file_size - 8);
    if ((uint64_t)header_offset + 8 + 4 > file_size || header_offset < 68)
        return kcf_fail(ctx, KCF_ERR_DB_FORMAT, "bad header offset %u", header_offset);
    const uint8_t *h = pre + (file_size - header_offset - 8);
    kcf_db_info_t info{};
    info.kmer_length = (int32_t)rd_u32(h + 0);
    /* mode */
    info.counter_size = (int32_t)rd_u32(h + 8);
    info.lut_prefix_length = (int32_t)rd_u32(h + 12);
    info.signature_length = (int32_t)rd_u32(h + 16);
    info.min_count = (int32_t)rd_u32(h + 20);
    info.max_count = (int32_t)rd_u32(h + 24);
    info.total_kmers = (int64_t)rd_u64(h + 28);
    info.both_strands = (h[36] == 0) ? 1 : 0; // KMC.java:133
    const uint32_t version = rd_u32(h + 64);
    if (version != 0x200) return kcf_fail(ctx, KCF_ERR_DB_FORMAT, "KMC version is not 0x200 (found 0x%x)", version);
    const int k = info.kmer_length, P = info.lut_prefix_length, L = info.signature_length, cs = info.counter_size;
    if (k < 1 || P < 0 || P > k || L < 1 || cs < 0 || info.total_kmers < 0)
        return kcf_fail(ctx, KCF_ERR_DB_FORMAT, "inconsistent header (k=%d P=%d L=%d counter=%d)", k, P, L, cs);
    if (k > 64) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "k=%d > 64 is not supported by this build", k);
    if (k > 32 && part_world > 1) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "k=%d > 32 with a partitioned table: the exchange moves 64-bit keys", k);
    if ((k - P) % 4 != 0) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "(k - lut_prefix_length) %% 4 != 0 (k=%d P=%d)", k, P);
    if (cs > 4) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "counter_size %d > 4", cs);
    if (L < 3 || L > 12 || L > k) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "signature length %d", L);
    if (P > 15) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "lut_prefix_length %d > 15", P);
    const uint64_t sig_map_size = (1ULL << (2 * L)) + 1;
    if (sig_map_size * 4 + header_offset + 8 + 4 > file_size) return kcf_fail(ctx, KCF_ERR_DB_FORMAT, "signature map does not fit the file");
    const uint64_t sig_map_start = file_size - header_offset - 8 - sig_map_size * 4;
    const uint64_t lut_size = 1ULL << (2 * P);
    const uint64_t single_lut_bytes = lut_size * 8;
    const uint64_t n_bins = (sig_map_start >= 12) ? (sig_map_start - 8 - 4) / single_lut_bytes : 0; // KMC.java:156
    const uint64_t lut_len = n_bins * lut_size;
    info.n_bins = (int32_t)n_bins;
    const uint32_t nsb = (uint32_t)(k - P) / 4, rec_size = nsb + (uint32_t)cs; // KMC.java:61
    const uint64_t N = (uint64_t)info.total_kmers;
    if (suf_len < 4 + N * rec_size) return kcf_fail(ctx, KCF_ERR_DB_FORMAT, ".kmc_suf holds fewer than total_kmers records");
    // signature map entries must name existing bins, else the reference indexes outside prefixArray
    for (uint64_t i = 0; i < sig_map_size; ++i)
        if (rd_u32(pre + sig_map_start + 4 * i) >= (n_bins ? n_bins : 1) && N > 0)
            return kcf_fail(ctx, KCF_ERR_DB_FORMAT, "signature map entry %llu names bin %u of %llu", (unsigned long long)i,
                            rd_u32(pre + sig_map_start + 4 * i), (unsigned long long)n_bins);

    // --- table geometry ---
    const int kk2 = 2 * k;
    KcfTableGeom g{};
    g.k = (uint32_t)k;
    g.kw = k > 32 ? 2u : 1u;
    g.kshift = k > 32 ? 0u : 64 - kk2;
    g.kmask = kk2 >= 64 ? ~0ULL : ((1ULL << kk2) - 1);
    g.km = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    g.km64 = k >= 64 ? ~0ULL : ((1ULL << k) - 1ULL);
    g.cw = cs <= 1 ? 1u : (cs == 2 ? 2u : 4u);
    if (g.kw == 2) { // 128-bit keys: 7 / 7 / 6 slots of 16 bytes (kcf_internal.cuh)
        g.S = g.cw == 4 ? 6u : 7u;
        g.coff = 16u * g.S;
        g.foff = 120u;
        g.fbits = g.cw == 2 ? 0u : 32u;
    } else {
        g.S = g.cw == 1 ? 13u : (g.cw == 2 ? 12u : 10u);
        g.coff = g.cw == 1 ? 112u : 8u * g.S;
        g.foff = g.cw == 1 ? 104u : 120u;
        g.fbits = g.cw == 1 ? 64u : 32u;
    }
    g.both_strands = (uint32_t)info.both_strands;
    if (g.kw == 2) {
        g.m = g.w = 0; // no minimizer: the home line is a hash of the whole key
        g.mm = 0;
    } else {
        // minimizer length: long enough that one m-mer value rarely names more than one locus of the sampled genome
        // (4^m >= 4 N) and that a run of k-mers sharing it fits one line (w = k-m+1 < S: a group larger than a line always
        // overflows), short enough that consecutive k-mers share it at all
        int m = ctx->minimizer_len;
        if (m <= 0) {
            m = 8;
            while (m < 16 && (1ULL << (2 * m)) < 4 * N) ++m;
            m = std::max(m, k - (int)g.S + 3); // w = S - 2: a run leaves two slots of its line free (C2: 78.9 / 79.9 / 79.9 G k-mers/s at w = 12 / 11 / 10)
        }
        m = std::min(std::min(m, 24), k);
        m = std::max(m, 1);
        if (k - m + 1 > 32) m = k - 31;
        g.m = (uint32_t)m;
        g.w = (uint32_t)(k - m + 1);
        g.mm = m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
    }
    double lf = ctx->load_factor;
    if (lf <= 0.0) { // automatic: sparse tables are faster (fewer keys outside their home line: 72.8 / 74.9 / 76.9 / 78.9 / 79.7 G
                     // k-mers/s at 0.3 / 0.25 / 0.2 / 0.15 / 0.125 on C2); start at 0.15 and give memory back when it is scarce
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const double share = part_world > 1 ? 1.0 / part_world : 1.0;
        const double steps[] = {0.15, 0.2, 0.3, 0.4, 0.5, 0.6, 0.75, 0.9};
        // against the device's TOTAL memory: every rank of a partitioned database must derive the same geometry
        for (double cand : steps) {
            lf = cand;
            const double ov = kcf_overflow_share(lf, N); // the overflow region grows with the density (below)
            if ((double)N * share / (g.S * lf) * KCF_LINE_BYTES * (1.0 + ov) <= 0.4 * (double)total_b) break;
        }
    }
    uint64_t nb = cs == 0 ? 1 : (uint64_t)((double)N / (g.S * lf)) + 1;
    nb = std::max<uint64_t>(nb, cs == 0 ? 1 : 64);
    if (nb >= 0xFFFFFFFFULL) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "%llu records need more than 2^32 table lines; partition the database", (unsigned long long)N);
    g.n_lines = nb;
    g.line_lo = 0;
    g.n_local = nb;
    if (part_world > 1) { // home lines [lo, hi) of the global line space
        const uint64_t lo = (nb * part_rank + part_world - 1) / part_world, hi = (nb * (part_rank + 1) + part_world - 1) / part_world;
        g.line_lo = lo;
        g.n_local = hi - lo;
    }
    // overflow region.  The k-mers of one minimizer run arrive together (1 to w of them, 6 on average), so the share of keys that
    // find their home line full grows with the density: 5 % at 0.15, 10 % at 0.3, 17 % at 0.5, 23 % at 0.7, 29 % at 0.9
    // (tools/line_occupancy_model.py) — about 0.33 lf; held in lines filled to ~0.6 that is 0.55 lf^2 of the home lines,
    // and never less than 1/8 of them
    // and never less than 1/8 of them.  A second term covers what the 32-bit minimizer hash costs very large databases: two
    // different minimizers with the same hash share a home line for good, and with N / 6 minimizer runs in 2^32 values the
    // share of runs that do is 1 - exp(-N / 6 / 2^32): 3 % at C2's 9e8 records, 9 % at 2.5e9, 44 % at C4's 1.5e10 — measured
    // there (profiles/r2u_c4_full_8gpu.json) as 6-15 % of the records in the stash with the overflow region sized by density alone
    g.n_ov = cs == 0 ? 1 : std::max<uint64_t>((uint64_t)((double)g.n_local * kcf_overflow_share(lf, N)), 32);
    nb = g.n_local + g.n_ov;                                       // lines allocated below
    if (nb >= 0xFFFFFFFFULL) return kcf_fail(ctx, KCF_ERR_UNSUPPORTED, "%llu table lines do not fit a 32-bit line index; partition the database", (unsigned long long)nb);
    g.stash_mask = 0;

    kcf_db *db = new kcf_db();
    db->ctx = ctx;
    uint64_t *d_lut = nullptr;
    uint32_t *d_sigmap = nullptr, *d_allowed = nullptr;
    KcfStashEntry *d_ovf = nullptr;
    unsigned long long *d_counters = nullptr;
    uint64_t *d_bound = nullptr; // per-group LUT bounds of the chunk in flight
    unsigned int *d_group_counter = nullptr; // one per chunk: the proof kernel's CTAs draw their groups from it
    KcfStaged *d_staged = nullptr;           // proven records of the two chunks in flight (k <= 32)
    cudaEvent_t ev_first = nullptr, ev_last = nullptr; // device time of the ingest kernels (kcf_db_info_t.load_phase_s)
    std::chrono::steady_clock::time_point t_setup = t0, t_streamed = t0;
    int rc = KCF_OK;
    const uint64_t chunk_rec = std::max<uint64_t>(256, (KCF_INGEST_CHUNK_BYTES) / std::max<uint32_t>(rec_size, 1) / 256 * 256);
    const uint64_t chunk_bytes = chunk_rec * std::max<uint32_t>(rec_size, 1);
    const uint64_t n_chunks = (N + chunk_rec - 1) / chunk_rec;
    uint64_t ovf_cap = N / 64 + 4096; // a second pass follows if the list turns out longer (dense tables, colliding minimizers)
    unsigned long long counters[4] = {0, 0, 0, 0};
    uint32_t flags[FLAG_COUNT] = {0};
    const bool smem_bits = L <= 9; // 4^9 bits = 32 KB of shared memory
    const size_t bits_bytes = std::max<size_t>(((size_t)1 << (2 * L)) / 8, 4);
    db->table_bytes = nb * KCF_LINE_BYTES;

#define DB_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            rc = kcf_fail(ctx, e__ == cudaErrorMemoryAllocation ? KCF_ERR_NOMEM : KCF_ERR_CUDA, "%s: %s (%s:%d)", \
                          #call, cudaGetErrorString(e__), __FILE__, __LINE__);                               \
            goto done;                                                                                        \
        }                                                                                                     \
    } while (0)

    // the table comes from the context's block pool: a host that opens one database after the other (a cohort run) gets the
    // previous table's memory back without a cudaFree / cudaMalloc round trip
    // a pooled table of another size is of no use to this database and would sit next to the new one (two 70 GB tables do
    // not fit one GPU): give large pooled blocks that cannot serve this request back to the driver first
    {
        bool freed = false;
        for (size_t i = 0; i < ctx->pool.size();) {
            const size_t b = ctx->pool[i].bytes;
            const bool fits = b >= db->table_bytes && b <= db->table_bytes + db->table_bytes / 8 + 4096;
            if (b >= (256ULL << 20) && !fits) {
                if (!freed) cudaStreamSynchronize(ctx->stream);
                freed = true;
                cudaFree(ctx->pool[i].p);
                ctx->pool.erase(ctx->pool.begin() + i);
            } else {
                ++i;
            }
        }
    }
    db->table = (uint8_t *)kcf_pool_get(ctx, db->table_bytes);
    if (!db->table) {
        rc = kcf_fail(ctx, KCF_ERR_NOMEM, "device memory for a table of %llu lines (%.1f GB)", (unsigned long long)nb, (double)db->table_bytes / 1e9);
        goto done;
    }
    DB_CUDA(cudaMemsetAsync(db->table, 0xFF, nb * KCF_LINE_BYTES, ctx->stream)); // empty keys, zero counts and masks (stored inverted)
    if (ctx->ing_slot_bytes < chunk_bytes) { // the staging ring, allocated by the context's first open
        for (int j = 0; j < KCF_INGEST_SLOTS; ++j) {
            if (ctx->ing_h[j]) cudaFreeHost(ctx->ing_h[j]);
            if (ctx->ing_d[j]) cudaFree(ctx->ing_d[j]);
            ctx->ing_h[j] = ctx->ing_d[j] = nullptr;
        }
        ctx->ing_slot_bytes = 0;
        for (int j = 0; j < KCF_INGEST_SLOTS; ++j) {
            DB_CUDA(cudaHostAlloc(&ctx->ing_h[j], chunk_bytes, cudaHostAllocDefault));
            DB_CUDA(cudaMalloc(&ctx->ing_d[j], chunk_bytes + 16)); // the 64-bit record loads may touch the word past the last record
            if (!ctx->ing_free[j]) DB_CUDA(cudaEventCreateWithFlags(&ctx->ing_free[j], cudaEventDisableTiming));
            if (!ctx->ing_copied[j]) DB_CUDA(cudaEventCreateWithFlags(&ctx->ing_copied[j], cudaEventDisableTiming));
        }
        ctx->ing_slot_bytes = chunk_bytes;
    }
    DB_CUDA(cudaMalloc(&d_lut, std::max<uint64_t>(lut_len, 1) * 8));
    DB_CUDA(cudaMalloc(&d_sigmap, sig_map_size * 4));
    DB_CUDA(cudaMalloc(&d_allowed, bits_bytes));
    DB_CUDA(cudaMalloc(&d_counters, 4 * sizeof(unsigned long long)));
    DB_CUDA(cudaMalloc(&d_bound, (chunk_rec / 256 + 2) * sizeof(uint64_t)));
    DB_CUDA(cudaMalloc(&d_group_counter, std::max<uint64_t>(n_chunks, 1) * sizeof(unsigned int)));
    if (g.kw == 1) DB_CUDA(cudaMalloc(&d_staged, 2 * chunk_rec * sizeof(KcfStaged))); // two buffers: the insert of chunk c runs under the proof of chunk c + 1
    if (!ctx->ing_stream) {
        DB_CUDA(cudaStreamCreateWithFlags(&ctx->ing_stream, cudaStreamNonBlocking));
        for (int j = 0; j < 2; ++j) {
            DB_CUDA(cudaEventCreateWithFlags(&ctx->ing_proved[j], cudaEventDisableTiming));
            DB_CUDA(cudaEventCreateWithFlags(&ctx->ing_inserted[j], cudaEventDisableTiming));
        }
    }

    DB_CUDA(cudaMemsetAsync(d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    DB_CUDA(cudaMemsetAsync(ctx->d_flags, 0, FLAG_COUNT * sizeof(uint32_t), ctx->stream));
    DB_CUDA(cudaMalloc(&d_ovf, ovf_cap * sizeof(KcfStashEntry)));
    if (lut_len) DB_CUDA(cudaMemcpyAsync(d_lut, pre + 4, lut_len * 8, cudaMemcpyHostToDevice, ctx->stream)); // KMC.java:153,159-163
    DB_CUDA(cudaMemcpyAsync(d_sigmap, pre + sig_map_start, sig_map_size * 4, cudaMemcpyHostToDevice, ctx->stream)); // :145-151
    {
        const uint32_t n_words = (uint32_t)(bits_bytes / 4);
        kcf_allowed_kernel<<<(n_words + 255) / 256, 256, 0, ctx->stream>>>(L, d_allowed);
        if (lut_len) kcf_lut_check_kernel<<<(unsigned)((lut_len + 255) / 256), 256, 0, ctx->stream>>>(d_lut, lut_len, N, ctx->d_flags);
        DB_CUDA(cudaGetLastError());
        if (smem_bits) DB_CUDA(cudaFuncSetAttribute(g.kw == 2 ? kcf_ingest_kernel<true, 2> : kcf_ingest_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    }
    cudaEventCreate(&ev_first);
    cudaEventCreate(&ev_last);
    t_setup = t_streamed = std::chrono::steady_clock::now();
    for (int pass = 0; pass < 4; ++pass) {
        if (pass > 0) {
            // the overflow list was too short (a table whose minimizers collide massively): the device counted
            // how long it has to be; reset the table and stream the records once more
            cudaFree(d_ovf);
            d_ovf = nullptr;
            ovf_cap = counters[2] + counters[2] / 8 + 4096; // placement races make the count vary a little between passes
            DB_CUDA(cudaMalloc(&d_ovf, ovf_cap * sizeof(KcfStashEntry)));
            DB_CUDA(cudaMemsetAsync(db->table, 0xFF, nb * KCF_LINE_BYTES, ctx->stream));
            DB_CUDA(cudaMemsetAsync(d_counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
        }
        const uint8_t *recs = suf + 4; // KMC.java:94 — skip the KMCS marker
        DB_CUDA(cudaMemsetAsync(d_group_counter, 0, std::max<uint64_t>(n_chunks, 1) * sizeof(unsigned int), ctx->stream));
        // ---- host side of the pipeline: filler threads copy whole chunks into the pinned ring, this thread queues the H2D
        // copy (copy stream) and the ingest kernel (main stream) of every chunk in order.  Slot s = chunk % SLOTS is free
        // again when the kernel of chunk - SLOTS has run (event ing_free[s], recorded by this thread: `launched` tells the
        // fillers that it has been).
        std::atomic<uint64_t> launched{0};
        std::vector<std::atomic<int>> ready(n_chunks);
        for (auto &r : ready) r.store(0, std::memory_order_relaxed);
        std::atomic<int> abort_fill{0};
        const unsigned n_fill = (unsigned)std::min<uint64_t>(KCF_INGEST_FILLERS, std::max<uint64_t>(n_chunks, 1));
        std::vector<std::thread> fillers;
        for (unsigned f = 0; f < n_fill; ++f)
            fillers.emplace_back([&, f] {
                cudaSetDevice(ctx->device);
                for (uint64_t c = f; c < n_chunks; c += n_fill) {
                    const int sl = (int)(c % KCF_INGEST_SLOTS);
                    if (c >= KCF_INGEST_SLOTS) {
                        while (launched.load(std::memory_order_acquire) + KCF_INGEST_SLOTS <= c && !abort_fill.load()) std::this_thread::yield();
                        if (abort_fill.load()) return;
                        cudaEventSynchronize(ctx->ing_free[sl]); // the kernel that last read this slot's device buffer (its H2D copy came before)
                    }
                    const uint64_t r0 = c * chunk_rec, n = std::min<uint64_t>(chunk_rec, N - r0);
                    memcpy(ctx->ing_h[sl], recs + r0 * rec_size, n * rec_size); // page cache / caller memory -> pinned staging
                    ready[c].store(1, std::memory_order_release);
                }
            });
        cudaError_t perr = cudaSuccess;
        for (uint64_t c = 0; c < n_chunks && perr == cudaSuccess; ++c) {
            const int sl = (int)(c % KCF_INGEST_SLOTS);
            const uint64_t r0 = c * chunk_rec, n = std::min<uint64_t>(chunk_rec, N - r0);
            while (!ready[c].load(std::memory_order_acquire)) std::this_thread::yield();
            // the copy rides the copy stream: later chunks cross PCIe while the ingest kernel of this one runs
            perr = cudaMemcpyAsync(ctx->ing_d[sl], ctx->ing_h[sl], n * rec_size, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (perr == cudaSuccess) perr = cudaEventRecord(ctx->ing_copied[sl], ctx->copy_stream);
            if (perr == cudaSuccess) perr = cudaStreamWaitEvent(ctx->stream, ctx->ing_copied[sl], 0);
            if (perr != cudaSuccess) break;
            KcfIngestParams p{};
            p.rec = ctx->ing_d[sl];
            p.rec0 = r0;
            p.n_rec = n;
            memset(p.prev, 0, sizeof p.prev);
            if (r0 > 0) memcpy(p.prev, recs + (r0 - 1) * rec_size, std::min<uint32_t>(rec_size, 16));
            p.lut = d_lut;
            p.lut_len = lut_len;
            p.n_groups = (uint32_t)((n + 255) / 256);
            kcf_lut_bounds_kernel<<<(p.n_groups + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_lut, lut_len, r0, n, p.n_groups, d_bound);
            p.cta_bound = d_bound;
            p.sigmap = d_sigmap;
            p.allowed = d_allowed;
            p.P = (uint32_t)P;
            p.L = (uint32_t)L;
            p.nsb = nsb;
            p.cs = (uint32_t)cs;
            p.rec_size = rec_size;
            p.table = db->table;
            p.ovf = d_ovf;
            p.ovf_cap = ovf_cap;
            p.counters = d_counters;
            p.part_rank = part_rank;
            p.part_world = part_world;
            p.flags = ctx->d_flags;
            p.group_counter = d_group_counter + c;
            if (c == 0) cudaEventRecord(ev_first, ctx->stream);
            if (g.kw == 2) {
                const unsigned grid = (unsigned)std::min<uint64_t>(p.n_groups, (uint64_t)ctx->sm_count * 6);
                if (smem_bits) kcf_ingest_kernel<true, 2><<<grid, 256, 32768, ctx->stream>>>(p, g);
                else kcf_ingest_kernel<false, 2><<<grid, 256, 0, ctx->stream>>>(p, g);
            } else {
                // proof on the main stream, insert on its own: the insert of chunk c may run under the proof of chunk c + 1
                // (measured: 0.170 -> 0.159 s for C2; giving the proof only half of every SM and the insert stream priority
                // made it 0.201 s — the two kernels want different shared-memory carve-outs and mostly take turns)
                const unsigned grid = (unsigned)std::min<uint64_t>(p.n_groups, (uint64_t)ctx->sm_count * 6);
                const int sb = (int)(c & 1);
                p.staged = d_staged + (size_t)sb * chunk_rec;
                if (c >= 2) perr = cudaStreamWaitEvent(ctx->stream, ctx->ing_inserted[sb], 0); // staged buffer sb free again
                if (smem_bits) kcf_ingest_kernel<true, 1><<<grid, 256, 32768, ctx->stream>>>(p, g);
                else kcf_ingest_kernel<false, 1><<<grid, 256, 0, ctx->stream>>>(p, g);
                if (perr == cudaSuccess) perr = cudaEventRecord(ctx->ing_proved[sb], ctx->stream);
                if (perr == cudaSuccess) perr = cudaStreamWaitEvent(ctx->ing_stream, ctx->ing_proved[sb], 0);
                kcf_insert_kernel<<<p.n_groups, 256, 0, ctx->ing_stream>>>(p, g);
                if (perr == cudaSuccess) perr = cudaEventRecord(ctx->ing_inserted[sb], ctx->ing_stream);
            }
            if (perr == cudaSuccess) perr = cudaGetLastError();
            if (perr == cudaSuccess) perr = cudaEventRecord(ctx->ing_free[sl], ctx->stream);
            launched.store(c + 1, std::memory_order_release);
        }
        if (g.kw == 1 && n_chunks) { // the main stream goes on only when the last inserts have landed
            cudaStreamWaitEvent(ctx->stream, ctx->ing_inserted[0], 0);
            if (n_chunks > 1) cudaStreamWaitEvent(ctx->stream, ctx->ing_inserted[1], 0);
        }
        cudaEventRecord(ev_last, ctx->stream);
        t_streamed = std::chrono::steady_clock::now();
        if (perr != cudaSuccess) abort_fill.store(1);
        launched.store(n_chunks + KCF_INGEST_SLOTS, std::memory_order_release); // releases fillers waiting on a chunk that will not be launched
        for (std::thread &t : fillers) t.join();
        if (perr != cudaSuccess) {
            cudaStreamSynchronize(ctx->stream);
            rc = kcf_fail(ctx, KCF_ERR_CUDA, "database ingest: %s", cudaGetErrorString(perr));
            goto done;
        }
        DB_CUDA(cudaMemcpyAsync(counters, d_counters, sizeof counters, cudaMemcpyDeviceToHost, ctx->stream));
        DB_CUDA(cudaMemcpyAsync(flags, ctx->d_flags, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
        DB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (counters[2] <= ovf_cap || flags[FLAG_LUT_BAD] || flags[FLAG_ORDER_BAD]) break;
    }
    if (flags[FLAG_LUT_BAD]) { rc = kcf_fail(ctx, KCF_ERR_DB_FORMAT, "prefix LUT is not monotone or exceeds total_kmers"); goto done; }
    if (flags[FLAG_ORDER_BAD]) { rc = kcf_fail(ctx, KCF_ERR_DB_ORDER, "records inside a (bin, prefix) range are not strictly ascending"); goto done; }
    if (counters[2] > ovf_cap) { rc = kcf_fail(ctx, KCF_ERR_NOMEM, "hash overflow list exhausted (%llu entries); lower the load factor", counters[2]); goto done; }
    if (counters[2] > 0) {
        uint64_t cap = 64;
        while (cap < 2 * counters[2]) cap <<= 1;
        g.stash_mask = cap - 1;
        if (cudaMalloc(&db->stash, cap * sizeof(KcfStashEntry)) != cudaSuccess) {
            (void)cudaGetLastError();
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            rc = kcf_fail(ctx, KCF_ERR_NOMEM, "device memory for a stash of %llu records (%.2f GB; %.1f of %.1f GB free; table %.1f GB at load factor %.2f)",
                          counters[2], (double)(cap * sizeof(KcfStashEntry)) / 1e9, (double)free_b / 1e9, (double)total_b / 1e9, (double)db->table_bytes / 1e9, lf);
            goto done;
        }
        DB_CUDA(cudaMemsetAsync(db->stash, 0, cap * sizeof(KcfStashEntry), ctx->stream));
        kcf_stash_build_kernel<<<(unsigned)((counters[2] + 255) / 256), 256, 0, ctx->stream>>>(d_ovf, counters[2], db->stash, g);
        DB_CUDA(cudaGetLastError());
        DB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    info.resident_kmers = (int64_t)(counters[0] + counters[2]);
    info.unreachable_kmers = (int64_t)counters[1];
    info.stash_kmers = (int64_t)counters[2];
    info.n_buckets = (int64_t)nb;
    info.elsewhere_kmers = (int64_t)counters[3];
    db->part_rank = (int)part_rank;
    db->part_world = (int)part_world;
    info.table_bytes = (int64_t)(nb * KCF_LINE_BYTES + (db->stash ? (g.stash_mask + 1) * sizeof(KcfStashEntry) : 0));
    info.load_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    info.load_phase_s[0] = std::chrono::duration<double>(t_setup - t0).count();
    info.load_phase_s[1] = std::chrono::duration<double>(t_streamed - t_setup).count();
    info.load_phase_s[2] = info.load_seconds - info.load_phase_s[0] - info.load_phase_s[1];
    {
        float ms = 0;
        if (N > 0 && cudaEventElapsedTime(&ms, ev_first, ev_last) == cudaSuccess) info.load_phase_s[3] = ms * 1e-3;
        else (void)cudaGetLastError();
    }
    db->info = info;
    db->geom = g;

done:
    if (ev_first) cudaEventDestroy(ev_first);
    if (ev_last) cudaEventDestroy(ev_last);
    if (d_lut) cudaFree(d_lut);
    if (d_sigmap) cudaFree(d_sigmap);
    if (d_allowed) cudaFree(d_allowed);
    if (d_ovf) cudaFree(d_ovf);
    if (d_counters) cudaFree(d_counters);
    if (d_bound) cudaFree(d_bound);
    if (d_group_counter) cudaFree(d_group_counter);
    if (d_staged) cudaFree(d_staged);
    if (rc != KCF_OK) {
        if (db->table) {
            cudaStreamSynchronize(ctx->stream);
            kcf_pool_put(ctx, db->table, db->table_bytes);
        }
        if (db->stash) cudaFree(db->stash);
        delete db;
        return rc;
    }
    *out = db;
    return KCF_OK;
#undef DB_CUDA
}

namespace {
struct Mapped {
    const uint8_t *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const std::string &path)
    {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) return false;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = (const uint8_t *)m;
        return true;
    }
    ~Mapped()
    {
        if (p) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
    }
};
} // namespace

extern "C" int kcf_db_open(kcf_ctx *ctx, const char *kmc_prefix, int placement, kcf_db **out)
{
    if (!ctx || !kmc_prefix || !out) return KCF_ERR_ARG;
    *out = nullptr;
    Mapped pre, suf;
    std::string base(kmc_prefix);
    if (!pre.open(base + ".kmc_pre")) return kcf_fail(ctx, KCF_ERR_IO, "Error reading prefix file %s.kmc_pre", kmc_prefix); // KMC.java:165-167
    if (!suf.open(base + ".kmc_suf")) return kcf_fail(ctx, KCF_ERR_IO, "Error reading suffix buffers from file %s.kmc_suf", kmc_prefix); // :186-188
    return kcf_db_open_mem(ctx, pre.p, pre.n, suf.p, suf.n, placement, out);
}

extern "C" int kcf_db_info(kcf_db *db, kcf_db_info_t *out)
{
    if (!db || !out) return KCF_ERR_ARG;
    *out = db->info;
    return KCF_OK;
}

extern "C" void kcf_db_close(kcf_db *db)
{
    if (!db) return;
    cudaSetDevice(db->ctx->device);
    if (db->table) {
        cudaStreamSynchronize(db->ctx->stream); // nothing queued may still read the table once its block can be handed out again
        kcf_pool_put(db->ctx, db->table, db->table_bytes);
    }
    if (db->stash) cudaFree(db->stash);
    delete db;
}

extern "C" int kcf_db_count(kcf_ctx *ctx, kcf_db *db, const char *kmers_ascii, uint64_t n, int32_t *counts_out)
{
    if (!ctx || !db || (!kmers_ascii && n) || (!counts_out && n)) return KCF_ERR_ARG;
    if (n == 0) return KCF_OK;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    char *d_in = nullptr;
    int32_t *d_out = nullptr;
    const uint64_t k = db->geom.k;
    KCF_CUDA(ctx, cudaMalloc(&d_in, n * k));
    cudaError_t e = cudaMalloc(&d_out, n * 4);
    if (e != cudaSuccess) { cudaFree(d_in); return kcf_fail(ctx, KCF_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    cudaMemcpyAsync(d_in, kmers_ascii, n * k, cudaMemcpyHostToDevice, ctx->stream);
    kcf_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_in, n, db->table, db->stash, db->geom, d_out);
    cudaMemcpyAsync(counts_out, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "kcf_db_count: %s", cudaGetErrorString(e));
    return KCF_OK;
}

// ---- layout statistics (measurement helper): how many home / overflow lines hold 0, 1, ..., S keys --------------------
__global__ void kcf_line_hist_kernel(const uint8_t *__restrict__ table, uint64_t n_lines, uint32_t S, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[16];
    if (threadIdx.x < 16) sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(table + i * KCF_LINE_BYTES);
        uint32_t n = 0;
        for (uint32_t s = 0; s < S; ++s) n += w[s] != KCF_EMPTY_LO;
        atomicAdd(&sh[n], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

extern "C" int kcf_db_line_histogram(kcf_db *db, uint64_t hist_out[16])
{
    if (!db || !hist_out) return KCF_ERR_ARG;
    kcf_ctx *ctx = db->ctx;
    KCF_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr;
    KCF_CUDA(ctx, cudaMalloc(&d, 16 * 8));
    cudaMemsetAsync(d, 0, 16 * 8, ctx->stream);
    const uint64_t n = db->geom.n_local + db->geom.n_ov;
    kcf_line_hist_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(db->table, n, db->geom.S, d);
    cudaError_t e = cudaMemcpyAsync(hist_out, d, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return kcf_fail(ctx, KCF_ERR_CUDA, "kcf_db_line_histogram: %s", cudaGetErrorString(e));
    return KCF_OK;
}
