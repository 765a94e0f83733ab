// kcf_lookup.cuh — device-side probe of the HBM-resident k-mer table.
// Replaces KMC.getCount (KMC.java:292-326): one 32-byte sector read in the common case.
#pragma once
#include "kcf_internal.cuh"

// one whole bucket = one DRAM sector, fetched with a single 256-bit load that bypasses L1
// allocation (the probes are uniformly random; L1 is kept for the reference bases)
__device__ __forceinline__ void kcf_ld_bucket(const uint64_t *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
                 : "l"(p));
}

// reverse complement of a right-aligned 2k-bit k-mer value
__device__ __forceinline__ uint64_t kcf_revcomp(uint64_t x, uint32_t kshift)
{
    uint64_t z = __brevll(~x);
    z = ((z >> 1) & 0x5555555555555555ULL) | ((z & 0x5555555555555555ULL) << 1);
    return z >> kshift; // the complemented padding bits fall off the low end
}

// base-order reversal of a 2k-bit value packed LSB-first (base j in bits 2j) into the
// first-base-most-significant value the reference uses (Kmer.java:232-252)
__device__ __forceinline__ uint64_t kcf_pair_reverse(uint64_t x, uint32_t kshift)
{
    uint64_t z = __brevll(x);
    z = ((z >> 1) & 0x5555555555555555ULL) | ((z & 0x5555555555555555ULL) << 1);
    return z >> kshift;
}

// Match `tag` (valid|disp|rem) against the 4 slots of one bucket.  Returns true on a match and sets
// count; `full` tells whether the probe sequence must continue.
__device__ __forceinline__ bool kcf_match4(uint64_t s0, uint64_t s1, uint64_t s2, uint64_t s3, uint64_t tag,
                                           const KcfTableGeom &g, uint32_t &count, bool &full)
{
    const uint32_t cb = g.cbits;
    bool m0 = (s0 >> cb) == tag, m1 = (s1 >> cb) == tag, m2 = (s2 >> cb) == tag, m3 = (s3 >> cb) == tag;
    uint64_t hit = m0 ? s0 : (m1 ? s1 : (m2 ? s2 : s3));
    bool any = m0 | m1 | m2 | m3;
    count = (uint32_t)(hit & g.cmask);
    full = (s0 != 0) & (s1 != 0) & (s2 != 0) & (s3 != 0);
    return any;
}

// Continue a probe sequence after the home bucket was full and held no match (rare path).
static __device__ __noinline__ uint32_t kcf_lookup_tail(const uint64_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                                 const KcfTableGeom &g, uint64_t key, uint64_t h, uint64_t home)
{
    const uint64_t rem = h & g.rmask;
    for (uint32_t d = 1; d <= KCF_MAX_DISP; ++d) {
        uint64_t b = home + d;
        if (b >= g.n_buckets) b -= g.n_buckets;
        uint64_t s0, s1, s2, s3;
        kcf_ld_bucket(table + 4 * b, s0, s1, s2, s3);
        uint64_t tag = ((((uint64_t)(8u | d)) << g.rbits) | rem);
        uint32_t c;
        bool full;
        if (kcf_match4(s0, s1, s2, s3, tag, g, c, full)) return c;
        if (!full) return 0;
    }
    if (stash == nullptr) return 0;
    // stash: linear probing over 16-byte entries
    uint64_t i = (kcf_mix(key, g) * 0x9E3779B97F4A7C15ULL) >> 20;
    for (uint64_t n = 0; n <= g.stash_mask; ++n) {
        const KcfStashEntry e = stash[(i + n) & g.stash_mask];
        if (e.meta == 0) return 0;
        if (e.key == key) return (uint32_t)e.meta;
    }
    return 0;
}

// Full lookup of one canonical k-mer value (used by the count kernel and the slow paths).
__device__ __forceinline__ uint32_t kcf_lookup(const uint64_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                               const KcfTableGeom &g, uint64_t key)
{
    uint64_t h = kcf_mix(key, g);
    uint64_t home = kcf_home_bucket(h, g);
    uint64_t s0, s1, s2, s3;
    kcf_ld_bucket(table + 4 * home, s0, s1, s2, s3);
    uint64_t tag = (((uint64_t)8u << g.rbits) | (h & g.rmask));
    uint32_t c;
    bool full;
    if (kcf_match4(s0, s1, s2, s3, tag, g, c, full)) return c;
    if (!full) return 0;
    return kcf_lookup_tail(table, stash, g, key, h, home);
}
