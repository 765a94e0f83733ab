// kcf_gap.cuh — the gap monoid (K4), shared by every kernel that folds hit / valid bitmaps into the per-window
// statistics of GetVariants.processWindow (GetVariants.java:217-252, getDistance :267-273).
//
// A KcfGap (kcf_internal.cuh) summarises a run of valid k-mers, each hit or miss; kcf_gap_combine is the in-order
// concatenation of two runs and is associative, so 32 positions (bit tricks) -> warp (kcf_gap_fold_warp) -> tile -> window
// reproduces the reference's sequential state machine bit for bit.  ONE definition: the replicated kernel
// (kcf_screen.cu), the exchange fold and the scan fold (kcf_part.cu) all include this file.
#pragma once
#include "kcf_internal.cuh"

// GetVariants.java:267-273 getDistance
__device__ __forceinline__ uint32_t kcf_gap_distance(uint32_t gap, uint32_t k)
{
    int32_t d = (int32_t)gap - ((int32_t)k - 1);
    if (d <= 0) d = abs(d + 1);
    return (uint32_t)d;
}

__device__ __forceinline__ KcfGap kcf_gap_zero()
{
    KcfGap r;
    r.n = r.obs = r.lead = r.trail = r.vin = r.inner = r.has = r.starts = 0;
    r.sum = 0;
    return r;
}

// in-order concatenation of two summaries
__device__ __forceinline__ KcfGap kcf_gap_combine(const KcfGap &a, const KcfGap &b, uint32_t k)
{
    if (b.n == 0) return a;
    if (a.n == 0) return b;
    KcfGap r;
    r.n = a.n + b.n;
    r.obs = a.obs + b.obs;
    r.sum = a.sum + b.sum;
    r.starts = a.starts + b.starts;
    r.vin = a.vin + b.vin;
    r.inner = a.inner + b.inner;
    r.has = a.has | b.has;
    if (a.has && b.has) {
        uint32_t g = a.trail + b.lead; // a miss run closed by hits on both sides (GetVariants.java:227-238)
        if (g > 0) {
            r.vin += 1;
            r.inner += kcf_gap_distance(g, k);
        }
        r.lead = a.lead;
        r.trail = b.trail;
    } else if (a.has) {
        r.lead = a.lead;
        r.trail = a.trail + b.n;
    } else if (b.has) {
        r.lead = a.n + b.lead;
        r.trail = b.trail;
    } else {
        r.lead = r.n;
        r.trail = r.n;
    }
    return r;
}

// gap summary of 32 consecutive positions from their bitmaps (bit i = position i): `vw` marks the positions where a
// k-mer ends, `hw` (a subset) the observed ones, `sw` the k-mers that open a valid stretch.  Positions without a k-mer
// are transparent: a miss run continues across them (GetVariants.java:217-245 runs over the compacted k-mer list).
__device__ __forceinline__ KcfGap kcf_gap_from_bits(uint32_t hw, uint32_t vw, uint32_t sw, uint32_t k)
{
    KcfGap a;
    a.n = __popc(vw);
    a.obs = __popc(hw);
    a.starts = __popc(sw);
    a.sum = 0;
    a.vin = a.inner = 0;
    a.has = hw != 0;
    if (!hw) {
        a.lead = a.trail = a.n;
        return a;
    }
    const uint32_t first = __ffs(hw) - 1, last = 31 - __clz(hw);
    a.lead = __popc(vw & ((1u << first) - 1u));
    a.trail = __popc(vw & ~(0xFFFFFFFFu >> (31 - last)));
    uint32_t zr = ~hw & (0xFFFFFFFFu >> (31 - last)) & ~((1u << first) - 1u); // non-hit positions between two hits
    while (zr) {
        const uint32_t s = __ffs(zr) - 1;
        const uint32_t e = __ffs(~(zr >> s)) - 1; // length of this run of non-hit positions (ends before bit `last`)
        const uint32_t gm = ((1u << e) - 1u) << s;
        const uint32_t glen = __popc(vw & gm);
        if (glen) {
            a.vin += 1;
            a.inner += kcf_gap_distance(glen, k);
        }
        zr &= ~gm;
    }
    return a;
}

// Ordered reduction of the per-lane summaries of a warp (lane order = position order), without a combine tree: lane j holds the bitmaps of positions [32 j, 32 j + 32) (zero words for lanes past
// the data).  The counting fields are plain warp sums; a miss run that crosses lanes is closed by the lane holding its left
// hit, which fetches the next hit lane's leading misses and the k-mers in between (prefix sums of the per-lane k-mer
// counts) — exactly what kcf_gap_combine does pairwise, since lanes without a hit only lengthen the open run.  About 50
// warp instructions instead of five shuffle-and-combine rounds; the result is uniform over the warp (sum = 0).
__device__ __forceinline__ KcfGap kcf_gap_fold_warp(uint32_t hw, uint32_t vw, uint32_t sw, uint32_t lane, uint32_t k)
{
    const KcfGap a = kcf_gap_from_bits(hw, vw, sw, k);
    uint32_t incl = a.n; // inclusive prefix sum of the k-mer counts over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    const uint32_t excl = incl - a.n, total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t hasmask = __ballot_sync(0xffffffffu, a.has != 0);
    KcfGap r;
    r.n = total;
    r.obs = __reduce_add_sync(0xffffffffu, a.obs);
    r.starts = __reduce_add_sync(0xffffffffu, a.starts);
    r.sum = 0;
    r.has = hasmask != 0;
    if (!hasmask) { // warp uniform: no hit in the chunk
        r.lead = r.trail = total;
        r.vin = r.inner = 0;
        return r;
    }
    const uint32_t above = hasmask & (0xFFFFFFFEu << lane);     // hit lanes after this one
    const uint32_t nxt = above ? __ffs(above) - 1 : lane;       // the next of them (this lane when there is none)
    const uint32_t lead_n = __shfl_sync(0xffffffffu, a.lead, nxt), excl_n = __shfl_sync(0xffffffffu, excl, nxt);
    uint32_t vin = a.vin, inner = a.inner;
    if (a.has && above) {
        const uint32_t g = a.trail + (excl_n - incl) + lead_n; // misses between this lane's last hit and the next lane's first
        if (g > 0) {
            vin += 1;
            inner += kcf_gap_distance(g, k);
        }
    }
    r.vin = __reduce_add_sync(0xffffffffu, vin);
    r.inner = __reduce_add_sync(0xffffffffu, inner);
    const uint32_t f = __ffs(hasmask) - 1, l = 31 - __clz(hasmask);
    r.lead = __shfl_sync(0xffffffffu, excl + a.lead, f);
    r.trail = __shfl_sync(0xffffffffu, a.trail + (total - incl), l);
    return r;
}
