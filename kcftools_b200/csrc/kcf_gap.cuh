// kcf_gap.cuh — the gap monoid (K4), shared by every kernel that folds hit / valid bitmaps into the per-window
// statistics of GetVariants.processWindow (GetVariants.java:217-252, getDistance :267-273).
//
// A KcfGap (kcf_internal.cuh) summarises a run of valid k-mers, each hit or miss; kcf_gap_combine is the in-order
// concatenation of two runs and is associative, so 32 positions (bit tricks) -> warp (shuffle tree) -> tile -> window
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

// every field but `sum` (Σ count is not tied to positions: callers reduce it as a plain sum)
__device__ __forceinline__ KcfGap kcf_gap_shfl_down(const KcfGap &a, int delta)
{
    KcfGap r;
    r.n = __shfl_down_sync(0xffffffffu, a.n, delta);
    r.obs = __shfl_down_sync(0xffffffffu, a.obs, delta);
    r.lead = __shfl_down_sync(0xffffffffu, a.lead, delta);
    r.trail = __shfl_down_sync(0xffffffffu, a.trail, delta);
    r.vin = __shfl_down_sync(0xffffffffu, a.vin, delta);
    r.inner = __shfl_down_sync(0xffffffffu, a.inner, delta);
    r.has = __shfl_down_sync(0xffffffffu, a.has, delta);
    r.starts = __shfl_down_sync(0xffffffffu, a.starts, delta);
    r.sum = 0;
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

// ordered reduction of 32 per-lane summaries (lane order = position order); the result is meaningful in lane 0
__device__ __forceinline__ KcfGap kcf_gap_warp_reduce(KcfGap a, uint32_t lane, uint32_t k)
{
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const KcfGap b = kcf_gap_shfl_down(a, d);
        if (lane + d < 32) a = kcf_gap_combine(a, b, k);
    }
    return a;
}
