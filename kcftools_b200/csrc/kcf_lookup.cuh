// kcf_lookup.cuh — device-side pieces shared by the database loader and the screening kernel:
// the two k-mer encodings, the minimizer that picks a k-mer's home line, and the generic (global-memory)
// probe that replaces KMC.getCount (KMC.java:292-326) outside the tiled fast path.
#pragma once
#include "kcf_internal.cuh"

// ------------------------------------------------------------------------------------------------------------
// Two encodings of a k-mer meet in this library.
//
//   * the REFERENCE's value (Kmer.java:232-252): 2 bits per base, first base most significant, right aligned in
//     64 bits.  The database records arrive in it and the loader's reachability proof (canonical form, signature,
//     bin) is stated in it.
//   * the TABLE KEY: two BIT PLANES of k bits each, base j in bit j — plane 0 holds bit 0 of every base code, plane 1
//     bit 1 (A = 00, C = 01, G = 10, T = 11); key = plane1 << 32 | plane0.  The packed reference sequences use the same
//     planes (32 bases per word pair), so the screening kernel cuts a k-mer or an m-mer out of them with one funnel
//     shift per plane, and the reverse complement is ONE bit reversal per plane plus a complement (A<->T, C<->G flips
//     both bits) — no pair swaps.  For a both-strands database the key is the smaller of the two strands' plane
//     forms: any strand-symmetric injective choice works, since the table only has to answer "is the record that
//     spells this k-mer's canonical form present" and the loader stores exactly that record under this key.
// ------------------------------------------------------------------------------------------------------------

// bit-reverse a 64-bit word keeping each 2-bit base code intact: base j moves to pair 31-j
__device__ __forceinline__ uint64_t kcf_pair_reverse64(uint64_t x)
{
    uint64_t z = __brevll(x);
    return ((z >> 1) & 0x5555555555555555ULL) | ((z & 0x5555555555555555ULL) << 1);
}

// reverse complement of a right-aligned 2k-bit k-mer value (first base most significant)
__device__ __forceinline__ uint64_t kcf_revcomp(uint64_t x, uint32_t kshift)
{
    return kcf_pair_reverse64(~x) >> kshift; // the complemented padding bits fall off the low end
}

// base-order reversal of a 2k-bit value: LSB-first packing (base j in bits 2j) <-> the
// first-base-most-significant value the reference uses (Kmer.java:232-252)
__device__ __forceinline__ uint64_t kcf_pair_reverse(uint64_t x, uint32_t kshift)
{
    return kcf_pair_reverse64(x) >> kshift;
}

// the even bits of x, compacted
__host__ __device__ __forceinline__ uint32_t kcf_even_bits(uint64_t x)
{
    x &= 0x5555555555555555ULL;
    x = (x | (x >> 1)) & 0x3333333333333333ULL;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)x;
}

// reverse complement of one bit plane of n bases (n = 1..32; `nm` = its n one-bits)
__device__ __forceinline__ uint32_t kcf_plane_rc(uint32_t f, uint32_t n, uint32_t nm)
{
    return (__brev(f) >> (32u - n)) ^ nm;
}

// the smaller of a sequence's two strands in plane form, compared as plane1:plane0 (strand symmetric)
__device__ __forceinline__ void kcf_plane_canonical(uint32_t f0, uint32_t f1, uint32_t r0, uint32_t r1, uint32_t &c0, uint32_t &c1)
{
    const bool rev = r1 < f1 || (r1 == f1 && r0 < f0);
    c0 = rev ? r0 : f0;
    c1 = rev ? r1 : f1;
}

// table key of a k-mer given as the reference's value
__device__ __forceinline__ uint64_t kcf_table_key(uint64_t kmer, const KcfTableGeom &g)
{
    const uint64_t E = kcf_pair_reverse(kmer, g.kshift); // base j in bits 2j, 2j+1
    uint32_t f0 = kcf_even_bits(E), f1 = kcf_even_bits(E >> 1);
    if (g.both_strands) kcf_plane_canonical(f0, f1, kcf_plane_rc(f0, g.k, g.km), kcf_plane_rc(f1, g.k, g.km), f0, f1);
    return ((uint64_t)f1 << 32) | f0;
}

// Order hash of an m-mer from the plane form (c0, c1) of its canonical strand: what ranks the m-mers of a k-mer when
// its minimizer is chosen.  Two multiplies fold the planes and carry every input bit upwards, one xor-shift / multiply
// round spreads them; the rank is decided by the high bits.
__host__ __device__ __forceinline__ uint32_t kcf_order_hash(uint32_t c0, uint32_t c1)
{
    uint32_t h = c0 * 0x9E3779B1u + c1 * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x846ca68bU;
    return h;
}

// order hash of the m-mer whose forward-strand planes are x0, x1 (bits above m may hold anything)
__device__ __forceinline__ uint32_t kcf_mmer_order(uint32_t x0, uint32_t x1, uint32_t m, uint32_t mm)
{
    const uint32_t r0 = (~__brev(x0)) >> (32u - m), r1 = (~__brev(x1)) >> (32u - m); // the shift drops the unknown high bits
    uint32_t c0, c1;
    kcf_plane_canonical(x0 & mm, x1 & mm, r0, r1, c0, c1);
    return kcf_order_hash(c0, c1);
}

__device__ __forceinline__ uint32_t kcf_home_line(uint32_t mu, const KcfTableGeom &g)
{
    return __umulhi(kcf_mix32(mu ^ 0x9E3779B9u), (uint32_t)g.n_lines);
}

// minimizer value of a k-mer given as its table key (either strand's plane form gives the same value: the order
// hash is strand symmetric and the k-mer's m-mers are the same set)
__device__ __forceinline__ uint32_t kcf_minimizer_of_key(uint64_t key, const KcfTableGeom &g)
{
    const uint32_t f0 = (uint32_t)key, f1 = (uint32_t)(key >> 32);
    uint32_t mu = 0xFFFFFFFFu;
    for (uint32_t j = 0; j < g.w; ++j) mu = min(mu, kcf_mmer_order(f0 >> j, f1 >> j, g.m, g.mm));
    return mu;
}

// local index of the line that mask bit d of home line `home` names: d = 0 the home line itself, d >= 1 the (d-1)-th
// line of the home's probe sequence in the overflow region (stored after the local home lines)
__device__ __forceinline__ uint32_t kcf_line_wrap(uint32_t home, uint32_t d, const KcfTableGeom &g)
{
    if (d == 0) return (uint32_t)(home - g.line_lo);
    uint64_t o = __umulhi(kcf_mix32(home ^ 0x85EBCA6Bu), (uint32_t)g.n_ov) + (d - 1);
    if (o >= g.n_ov) o -= g.n_ov;
    return (uint32_t)(g.n_local + o);
}

// rank that owns a home line when the line space is cut in `world` equal ranges
__host__ __device__ __forceinline__ uint32_t kcf_line_owner(uint32_t home, uint64_t n_lines, uint32_t world)
{
    return (uint32_t)(((uint64_t)home * world) / n_lines);
}

// the same value from the precomputed map of an exchange workspace (exact: the estimate is a lower bound, the slice bounds decide)
__device__ __forceinline__ uint32_t kcf_line_owner_mapped(uint32_t home, const KcfXgDev &x)
{
    uint32_t e = __umulhi(home, x.own_mul);
    while (home >= x.own_bound[e + 1]) ++e;
    return e;
}

// the 16-bit mask of a home line (stored inverted so that the table can be initialised with 0xFF bytes)
__device__ __forceinline__ uint32_t kcf_mask_from_word31(uint32_t w31) { return (~w31) >> 16; }

// does the home line's filter (stored inverted, as read from memory) admit that `key` lives outside its home line?
__device__ __forceinline__ bool kcf_filter_pass64(uint64_t stored, uint64_t key)
{
    const uint32_t h = kcf_filter_hash(key);
    const uint64_t f = ~stored;
    return ((f >> (h >> 26)) & (f >> ((h >> 20) & 63u)) & 1ULL) != 0;
}
__device__ __forceinline__ bool kcf_filter_pass32(uint32_t stored, uint64_t key)
{
    const uint32_t h = kcf_filter_hash(key);
    const uint32_t f = ~stored;
    return ((f >> (h >> 27)) & (f >> ((h >> 22) & 31u)) & 1u) != 0;
}
__device__ __forceinline__ bool kcf_filter_pass(const uint8_t *home_line, uint64_t key, const KcfTableGeom &g)
{
    if (g.fbits == 64) return kcf_filter_pass64(__ldg(reinterpret_cast<const unsigned long long *>(home_line + g.foff)), key);
    return kcf_filter_pass32(__ldg(reinterpret_cast<const uint32_t *>(home_line + g.foff)), key);
}

// count of slot s of a line image (global or shared memory)
__device__ __forceinline__ uint32_t kcf_slot_count(const uint8_t *line, uint32_t s, const KcfTableGeom &g)
{
    const uint8_t *c = line + g.coff + g.cw * s;
    if (g.cw == 1) return *c;
    if (g.cw == 2) return *reinterpret_cast<const uint16_t *>(c);
    return *reinterpret_cast<const uint32_t *>(c);
}

// search one line in global memory.  Inside a line the low words of the live keys are distinct (the loader
// guarantees it), so the first low-word match is the only candidate.
__device__ __forceinline__ bool kcf_line_find(const uint8_t *line, uint64_t key, const KcfTableGeom &g, uint32_t &count)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(line);
    const uint32_t lo = (uint32_t)key, hi = (uint32_t)(key >> 32);
    for (uint32_t s = 0; s < g.S; ++s) {
        const uint32_t v = __ldg(w + s);
        if (v == lo) {
            if (__ldg(w + g.S + s) != hi) return false;
            count = kcf_slot_count(line, s, g);
            return true;
        }
        if (v == KCF_EMPTY_LO) return false; // occupied slots form a prefix of the line
    }
    return false;
}

static __device__ __noinline__ uint32_t kcf_stash_find(const KcfStashEntry *__restrict__ stash, const KcfTableGeom &g, uint64_t key, uint64_t key_hi = 0)
{
    if (stash == nullptr) return 0;
    uint64_t i = kcf_mix64(key ^ (key_hi * 0x9E3779B97F4A7C15ULL));
    for (uint64_t n = 0; n <= g.stash_mask; ++n) {
        const KcfStashEntry e = stash[(i + n) & g.stash_mask];
        if (e.meta == 0) return 0;
        if (e.key == key && e.key_hi == key_hi) return (uint32_t)e.meta;
    }
    return 0;
}

// Probe the lines named by `mask` bits [dmin, 14] of home line `home`, then the stash if bit 15 is set.
static __device__ __noinline__ uint32_t kcf_probe_lines(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                                        const KcfTableGeom &g, uint64_t key, uint32_t home, uint32_t mask, uint32_t dmin)
{
    if (!KCF_KEY_IN_LINES(key)) dmin = KCF_MAX_DISP + 1; // low word doubles as a slot marker: such keys live in the stash
    for (uint32_t d = dmin; d <= KCF_MAX_DISP; ++d) {
        if (!((mask >> d) & 1u)) continue;
        uint32_t c;
        if (kcf_line_find(table + (uint64_t)kcf_line_wrap(home, d, g) * KCF_LINE_BYTES, key, g, c)) return c;
    }
    if ((mask >> KCF_STASH_BIT) & 1u) return kcf_stash_find(stash, g, key);
    return 0;
}

// Lookup of one canonical k-mer value whose (global) home line is known.
__device__ __forceinline__ uint32_t kcf_lookup_at(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                                  const KcfTableGeom &g, uint64_t key, uint32_t home)
{
    const uint8_t *L = table + (uint64_t)kcf_line_wrap(home, 0, g) * KCF_LINE_BYTES;
    uint32_t c;
    if (KCF_KEY_IN_LINES(key) && kcf_line_find(L, key, g, c)) return c;
    if (!kcf_filter_pass(L, key, g)) return 0; // no key homed here that lives elsewhere looks like this one
    const uint32_t w31 = __ldg(reinterpret_cast<const uint32_t *>(L) + 31);
    return kcf_probe_lines(table, stash, g, key, home, kcf_mask_from_word31(w31), 1);
}

// Full lookup of one canonical k-mer value (count kernel and slow paths).
__device__ __forceinline__ uint32_t kcf_lookup(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash,
                                               const KcfTableGeom &g, uint64_t key)
{
    return kcf_lookup_at(table, stash, g, key, kcf_home_line(kcf_minimizer_of_key(key, g), g));
}

// Search the S low key words of a line (already in registers: a, b, c and the 13th in d0) for `key`; live low words of a line
// are distinct, so the low-word match is the only candidate and is confirmed on the high word.  The line is in L1 by then:
// high word and count depend on the slot only and are read together (one L1 round trip, not two).
template <int S>
__device__ __forceinline__ bool kcf_match_line(const uint4 a, const uint4 b, const uint4 c, const uint32_t d0, const uint8_t *line, uint64_t key,
                                               uint32_t &count)
{
    const uint32_t lo = (uint32_t)key;
    int idx = -1;
    if (a.x == lo) idx = 0;
    if (a.y == lo) idx = 1;
    if (a.z == lo) idx = 2;
    if (a.w == lo) idx = 3;
    if (b.x == lo) idx = 4;
    if (b.y == lo) idx = 5;
    if (b.z == lo) idx = 6;
    if (b.w == lo) idx = 7;
    if (c.x == lo) idx = 8;
    if (c.y == lo) idx = 9;
    if (S > 10 && c.z == lo) idx = 10;
    if (S > 11 && c.w == lo) idx = 11;
    if (S > 12 && d0 == lo) idx = 12;
    if (idx < 0) return false;
    constexpr int CW = S == 13 ? 1 : (S == 12 ? 2 : 4);
    constexpr int COFF = S == 13 ? 112 : 8 * S;
    const uint8_t *cp = line + COFF + CW * idx;
    const uint32_t hiw = __ldg(reinterpret_cast<const uint32_t *>(line) + S + idx);
    const uint32_t cntw = CW == 1 ? (uint32_t)__ldg(cp) : (CW == 2 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(cp)) : __ldg(reinterpret_cast<const uint32_t *>(cp)));
    if (hiw != (uint32_t)(key >> 32)) return false;
    count = cntw;
    return true;
}

// Probe one table line for `key`: the S low key words sit in the line's first two 32-byte sectors (4 loads of up to 16 bytes).
template <int S>
__device__ __forceinline__ bool kcf_probe_line(const uint8_t *line, uint64_t key, uint32_t &count)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(line);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const uint32_t d0 = S > 12 ? __ldg(reinterpret_cast<const uint32_t *>(line) + 12) : 0u;
    return kcf_match_line<S>(a, b, c, d0, line, key, count);
}

// ------------------------------------------------------------------------------------------------------------
// k = 33 .. 64 (kw = 2): 128-bit keys.  Same line discipline as above (low words distinct inside a line, overflow
// lines named by the home line's mask, filter over the keys that live elsewhere, stash), different geometry.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t kcf_plane_rc64(uint64_t f, uint32_t n, uint64_t nm)
{
    return (__brevll(f) >> (64u - n)) ^ nm;
}

// table key of a k-mer given as the reference's value (2k bits, right aligned in 128)
__device__ __forceinline__ KcfKey2 kcf_table_key2(unsigned __int128 kmer, const KcfTableGeom &g)
{
    // base-order reversal of the 64 pairs, then the 2k used bits moved down: base j in bits 2j, 2j+1
    const uint64_t hi = (uint64_t)(kmer >> 64), lo = (uint64_t)kmer;
    const unsigned __int128 rev = ((unsigned __int128)kcf_pair_reverse64(lo) << 64) | kcf_pair_reverse64(hi);
    const unsigned __int128 E = rev >> (128u - 2u * g.k);
    const uint64_t e0 = (uint64_t)E, e1 = (uint64_t)(E >> 64);
    KcfKey2 f;
    f.p0 = (uint64_t)kcf_even_bits(e0) | ((uint64_t)kcf_even_bits(e1) << 32);
    f.p1 = (uint64_t)kcf_even_bits(e0 >> 1) | ((uint64_t)kcf_even_bits(e1 >> 1) << 32);
    if (g.both_strands) {
        const uint64_t r0 = kcf_plane_rc64(f.p0, g.k, g.km64), r1 = kcf_plane_rc64(f.p1, g.k, g.km64);
        if (r1 < f.p1 || (r1 == f.p1 && r0 < f.p0)) {
            f.p0 = r0;
            f.p1 = r1;
        }
    }
    return f;
}

__device__ __forceinline__ uint32_t kcf_key2_hash(const KcfKey2 &key)
{
    return kcf_filter_hash(key.p0) * 0x2545F491u ^ kcf_filter_hash(key.p1 ^ 0xD6E8FEB86659FD93ULL);
}

__device__ __forceinline__ uint32_t kcf_home_line2(const KcfKey2 &key, const KcfTableGeom &g)
{
    return __umulhi(kcf_mix32(kcf_key2_hash(key) ^ 0x9E3779B9u), (uint32_t)g.n_lines);
}

__device__ __forceinline__ bool kcf_filter_pass2(const uint8_t *home_line, const KcfKey2 &key, const KcfTableGeom &g)
{
    if (g.fbits == 0) return true; // no room for a filter in this geometry: every miss follows the mask
    return kcf_filter_pass32(__ldg(reinterpret_cast<const uint32_t *>(home_line + g.foff)), key.p0 ^ (key.p1 * 0x9E3779B97F4A7C15ULL));
}

// search one line for a 128-bit key; the S low words sit in front, slot s keeps its other three words at 4 S + 12 s
__device__ __forceinline__ bool kcf_line_find2(const uint8_t *line, const KcfKey2 &key, const KcfTableGeom &g, uint32_t &count)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(line);
    const uint32_t lo = (uint32_t)key.p0;
    for (uint32_t s = 0; s < g.S; ++s) {
        const uint32_t v = __ldg(w + s);
        if (v == lo) {
            const uint32_t *r = w + g.S + 3 * s;
            if (__ldg(r) != (uint32_t)(key.p0 >> 32) || __ldg(r + 1) != (uint32_t)key.p1 || __ldg(r + 2) != (uint32_t)(key.p1 >> 32)) return false;
            count = kcf_slot_count(line, s, g);
            return true;
        }
        if (v == KCF_EMPTY_LO) return false; // occupied slots form a prefix of the line
    }
    return false;
}

// full lookup of one 128-bit table key
__device__ __forceinline__ uint32_t kcf_lookup2(const uint8_t *__restrict__ table, const KcfStashEntry *__restrict__ stash, const KcfTableGeom &g,
                                                const KcfKey2 &key, uint32_t home)
{
    const uint8_t *L = table + (uint64_t)kcf_line_wrap(home, 0, g) * KCF_LINE_BYTES;
    const bool inl = (uint32_t)key.p0 != KCF_EMPTY_LO;
    uint32_t c;
    if (inl && kcf_line_find2(L, key, g, c)) return c;
    if (!kcf_filter_pass2(L, key, g)) return 0;
    const uint32_t mask = kcf_mask_from_word31(__ldg(reinterpret_cast<const uint32_t *>(L) + 31));
    if (inl)
        for (uint32_t d = 1; d <= KCF_MAX_DISP; ++d)
            if (((mask >> d) & 1u) && kcf_line_find2(table + (uint64_t)kcf_line_wrap(home, d, g) * KCF_LINE_BYTES, key, g, c)) return c;
    if ((mask >> KCF_STASH_BIT) & 1u) return kcf_stash_find(stash, g, key.p0, key.p1);
    return 0;
}

// Probe one table line for a 128-bit key (screening kernel): the S <= 7 low words arrive with two 16-byte loads; the
// low-word match is the only candidate (low words are distinct inside a line) and is confirmed on its other three words,
// which are read together with the count.
template <int S>
__device__ __forceinline__ bool kcf_probe_line2(const uint8_t *line, const KcfKey2 &key, const KcfTableGeom &g, uint32_t &count)
{
    const uint4 *q = reinterpret_cast<const uint4 *>(line);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    const uint32_t lo = (uint32_t)key.p0;
    int idx = -1;
    if (a.x == lo) idx = 0;
    if (a.y == lo) idx = 1;
    if (a.z == lo) idx = 2;
    if (a.w == lo) idx = 3;
    if (b.x == lo) idx = 4;
    if (b.y == lo) idx = 5;
    if (S > 6 && b.z == lo) idx = 6;
    if (idx < 0) return false;
    const uint32_t *r = reinterpret_cast<const uint32_t *>(line) + S + 3 * idx;
    const uint32_t r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    const uint32_t c = kcf_slot_count(line, (uint32_t)idx, g);
    if (r0 != (uint32_t)(key.p0 >> 32) || r1 != (uint32_t)key.p1 || r2 != (uint32_t)(key.p1 >> 32)) return false;
    count = c;
    return true;
}
