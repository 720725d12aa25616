// Single-pass chained exclusive scan (decoupled look-back) of 32-bit counts: out[c] = sum of counts[0..c), out[nc] =
// the total.  One launch of ceil(nc_max / kScanChunk) CTAs of kScanT threads; CTAs number themselves through `ticket`
// in the order they start, so a CTA only ever waits for CTAs that are already running.  `ticket` and `desc` must be
// zero at launch.  Used by the grid build (cell table of the target, grid_build.cu) and by the bulk NN query (cell
// table of the query batch, nn_bulk.cu).
#pragma once
#include <cuda_runtime.h>

namespace ficp {

constexpr int kScanT = 256;
constexpr int kScanPer = 8;
constexpr int kScanChunk = kScanT * kScanPer;  // 2048 counts per CTA
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagIncl = 2ull << 62, kValMask = (1ull << 62) - 1;

__device__ __forceinline__ void chained_scan_block(const unsigned* __restrict__ counts, long long nc, unsigned* ticket,
                                                   unsigned* max_out, unsigned long long* __restrict__ desc,
                                                   unsigned* __restrict__ out) {
    __shared__ unsigned wsum[kScanT / 32];
    __shared__ unsigned s_block, s_prefix;
    if (threadIdx.x == 0) s_block = atomicAdd(ticket, 1u);   // blocks are numbered in the order they start
    __syncthreads();
    const unsigned blk = s_block;
    const long long base = (long long)blk * kScanChunk + (long long)threadIdx.x * kScanPer;
    if ((long long)blk * kScanChunk >= nc) return;
    unsigned v[kScanPer];
    unsigned s = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        v[j] = (base + j < nc) ? counts[base + j] : 0u;
        s += v[j];
        mx = max(mx, v[j]);
    }
    // block-exclusive scan of the per-thread sums
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (l >= o) inc += t;
    }
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if (l == 31) wsum[w] = inc;
    if (max_out && l == 0 && mx) atomicMax(max_out, mx);
    __syncthreads();
    if (w == 0) {
        unsigned ws = (l < kScanT / 32) ? wsum[l] : 0u;
#pragma unroll
        for (int o = 1; o < kScanT / 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, ws, o);
            if (l >= o) ws += t;
        }
        if (l < kScanT / 32) wsum[l] = ws;  // inclusive over warps
    }
    __syncthreads();
    const unsigned total = wsum[kScanT / 32 - 1];
    const unsigned in_block = ((w > 0) ? wsum[w - 1] : 0u) + inc - s;
    // publish the aggregate, then warp 0 looks back for the exclusive prefix of this block, 32 predecessors per step:
    // sum the aggregates down to (and including) the nearest block that already knows its inclusive prefix
    if (w == 0) {
        unsigned prefix = 0;
        if (blk == 0) {
            if (l == 0) atomicExch(desc + 0, kFlagIncl | (unsigned long long)total);
        } else {
            if (l == 0) atomicExch(desc + blk, kFlagAgg | (unsigned long long)total);
            long long hi = (long long)blk - 1;     // nearest predecessor not yet accounted for
            for (;;) {
                const long long p = hi - l;
                unsigned long long d = kFlagIncl;   // lanes before block 0 read as "inclusive 0"
                if (p >= 0) {
                    do { d = *reinterpret_cast<volatile unsigned long long*>(desc + p); } while ((d >> 62) == 0);
                }
                const unsigned incl = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                const int first = incl ? (__ffs(incl) - 1) : 32;      // nearest predecessor with an inclusive prefix
                unsigned v2 = (l <= first) ? (unsigned)(d & kValMask) : 0u;
                v2 = __reduce_add_sync(0xFFFFFFFFu, v2);
                prefix += v2;
                if (incl) break;
                hi -= 32;
            }
            if (l == 0) atomicExch(desc + blk, kFlagIncl | (unsigned long long)(prefix + total));
        }
        if (l == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned run = s_prefix + in_block;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        if (base + j < nc) out[base + j] = run;
        run += v[j];
    }
    if (base <= nc - 1 && nc - 1 < base + kScanPer) out[nc] = run;  // total, written by the owner of the last count
}

}  // namespace ficp
