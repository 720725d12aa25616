// Kernel 4: the whole two-stage Fractional ICP loop on-device, batched over plots x start-pose
// hypotheses.  Fuses kernels 1b/2/3 (NN query, FRMSD trimming, closed-form rigid fit):
//
//   run()/_iterate()             /root/reference/ficp.py:122-154   -> icp_run_hypothesis
//   find_correspondences          ficp.py:65-71                     -> nn phase of icp_pass
//   find_optimal_fraction, frmsd  ficp.py:54-60,73-86               -> trim phase of icp_pass
//   get_n_first_elements          ficp.py:62-63                     -> (d2, index) threshold of the k-th element
//   compute_optimal_transform_2d  ficp.py:89-110                    -> icp_fit (no SVD: normalised (H00+H11, H01-H10))
//   apply_transform_2d_xy_only    ficp.py:112-119                   -> composed into (M, c); Z never touched
//
// Mapping (DESIGN.md "persistent kernel"):
//   * one WARP owns one (plot, hypothesis) ICP from start to convergence - no block barrier inside
//     the loop, so the 5..70-pass spread between hypotheses costs nothing;
//   * ELASTIC: warps with no ICP of their own (the batch is smaller than the machine - one stand sharded over 8
//     GPUs, one start pose per plot - or the plot is running out of hypotheses) HELP the ICPs in flight in
//     their CTA: a lead hands out the nearest-neighbour rounds of its pass through a shared-memory ticket;
//     trimming and fit stay with the lead in unchanged arithmetic, so results are bit-identical either way;
//   * SKIP TEST: a full search also yields a lower bound on the distance to every OTHER target point; the
//     difference to the winner's distance is the query's SLACK.  On later passes the slack shrinks by the distance
//     the pose update moves the query (triangle inequality, directed rounding); while the winner's new distance is
//     still below it, the winner provably is the unique nearest neighbour again and the query costs ONE distance
//     evaluation instead of a search (measured: ~80 % of all queries).  Only the others are compacted into a list
//     and searched; the result is bit-identical to searching every query on every pass;
//   * each lane owns E source points (N <= 32*E); the trim order is a register-resident bitonic
//     sort of packed 32-bit keys (quantised d2 | point index) followed by an exact fix-up on the
//     fp64 (d2, index) keys, a warp prefix scan of d2 and an arg-min of FRMSD(k);
//   * a CTA is bound to one plot at a time: the plot's source points, its FRMSD weight tables and
//     the WINDOW of grid cells its hypotheses can reach are staged in shared memory once and
//     shared by all warps; a query whose search leaves the window falls back to the global grid
//     (same exact search), so results never depend on the window;
//   * CTAs pull (plot, slice) work from a global counter and warps pull hypotheses of the plot
//     from a per-plot counter (dynamic load balance across the whole GPU).
#include <algorithm>
#include <climits>
#include "icp_shared.cuh"

namespace ficp {

namespace {

template <int E>
struct LaneCfg {
    static constexpr int kNPad = 32 * E;
    static constexpr int kIdxBits = (E == 1) ? 5 : (E == 2) ? 6 : (E == 4) ? 7 : (E == 8) ? 8 : (E == 16) ? 9 : 10;
    static constexpr unsigned kIdxMask = (1u << kIdxBits) - 1u;
};

// Elastic mode: what a lead warp publishes for the helper warps of its CTA (warps that have no ICP of their own).
struct __align__(16) SlotCtrl {
    Pose pose;         // pose of the pass in flight
    int ticket;        // (epoch << 8) | next work unit to hand out; a unit field >= `count` means "closed"
    int count;         // work units of the current epoch (rounds of 32 queries, or chunks of 32 deferred queries)
    int done;          // units of the current epoch completed (by anyone)
    int mode;          // 0: nearest-neighbour rounds, 1: chunks of the deferred list
    int have_prev;     // the pass has a previous pass's neighbours to seed from
    int ndef;          // length of the deferred list (mode 1)
    int nglob;         // helpers' count of queries that ran on the global grid
    int nlist;         // mode 0: entries of the search list (queries that failed the skip test)
};
constexpr int kTicketClosed = 0xFF;

// ---- shared-memory carve-up (host and device use the same function) -------------------------------
struct SmemLayout {
    size_t s_u, w_xy, s_z, w_z, s_g, sd2, snn, sord, ssl, ctrl, w_cell, rowoff, rowdelta, rowg, total;
};
// `slots` = ICPs in flight per CTA (= lead warps); the other warps of the CTA, if any, are helpers
__host__ __device__ inline SmemLayout smem_layout(int npad, bool z3, int slots, int wcap_pts, int wcap_cells,
                                                  int wcap_rows) {
    SmemLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~size_t(15); return r; };
    L.s_u = take((size_t)npad * 16);
    L.w_xy = take((size_t)wcap_pts * 16);
    L.s_z = take(z3 ? (size_t)npad * 8 : 0);
    L.w_z = take(z3 ? (size_t)wcap_pts * 8 : 0);
    L.s_g = take((size_t)kMaxStages * npad * 8);
    L.sd2 = take((size_t)slots * npad * 8);
    L.snn = take((size_t)slots * npad * 4);
    L.sord = take((size_t)slots * npad * 2);
    L.ssl = take((size_t)slots * npad * 2);
    L.ctrl = take((size_t)slots * sizeof(SlotCtrl));
    L.w_cell = take((size_t)wcap_cells * 4);
    L.rowoff = take((size_t)(wcap_rows + 1) * 4);
    L.rowdelta = take((size_t)wcap_rows * 4);
    L.rowg = take((size_t)wcap_rows * 4);
    L.total = o;
    return L;
}

// ---- register-resident bitonic sort of 32*E keys across one warp -----------------------------------
// Logical position p = lane*E + r.  Steps with stride < E are compare-exchanges between registers
// of one lane; larger strides exchange with lane ^ (stride/E) through shuffles.  Validated against
// a numpy emulation of the same loops.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(unsigned (&key)[E], int lane) {
    // phases k = 2 .. E/2: compare-exchanges between registers of one lane, directions known at compile time
#pragma unroll
    for (int k = 2; k < E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int l = r ^ j;
                if (l > r) {
                    const bool desc = (r & k) != 0;
                    const unsigned a = key[r], b = key[l];
                    const unsigned mn = min(a, b), mx = max(a, b);
                    key[r] = desc ? mx : mn;
                    key[l] = desc ? mn : mx;
                }
            }
        }
    }
    // phases k = E << m, m = 0..5: direction = bit m of the lane (bit 5 is always 0: the last phase ascends).
    // Kept as RUNTIME loops (one copy of the shuffle step and one of the in-lane steps) to keep the kernel's
    // instruction footprint small - 16 warps in different phases share one instruction cache.
#pragma unroll 1
    for (int m = 0; m <= 5; ++m) {
        const bool desc = ((lane >> m) & 1) != 0;
#pragma unroll 1
        for (int lj = (1 << m) >> 1; lj > 0; lj >>= 1) {
            const bool takemin = (((lane & lj) != 0) == desc);
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const unsigned o = __shfl_xor_sync(kFull, key[r], lj);
                key[r] = takemin ? min(key[r], o) : max(key[r], o);
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int l = r ^ j;
                if (l > r) {
                    const unsigned a = key[r], b = key[l];
                    const unsigned mn = min(a, b), mx = max(a, b);
                    key[r] = desc ? mx : mn;
                    key[l] = desc ? mn : mx;
                }
            }
        }
    }
}

// Hand out the next work unit of a slot's epoch in flight: -1 when none is left (or the slot is closed).
__device__ __forceinline__ int grab_unit(SlotCtrl* c, int lane) {
    int e = -1;
    if (lane == 0) {
        for (;;) {
            const int t = ld_volatile(&c->ticket);
            if ((t & 0xFF) >= ld_volatile(&c->count)) break;   // count belongs to t's epoch iff the CAS below succeeds
            if (atomicCAS(&c->ticket, t, t + 1) == t) { e = t & 0xFF; break; }
        }
    }
    return __shfl_sync(kFull, e, 0);
}

// Lead: open an epoch of `count` units for the helpers.
__device__ __forceinline__ void open_epoch(SlotCtrl* c, int& epoch, int mode, int count, int lane) {
    if (lane == 0) {
        c->mode = mode;
        c->done = 0;
        *reinterpret_cast<volatile int*>(&c->ticket) = (epoch << 8) | kTicketClosed;  // nobody grabs while count changes
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&c->count) = count;
        __threadfence_block();
        epoch = (epoch + 1) & 0x7FFFFF;
        *reinterpret_cast<volatile int*>(&c->ticket) = epoch << 8;
    }
    __syncwarp();
}

// Lead: wait until every unit of the epoch is done, then close the slot.
__device__ __forceinline__ void close_epoch(SlotCtrl* c, int epoch, int count, int lane) {
    if (lane == 0) {
        while (ld_volatile(&c->done) < count) __nanosleep(40);
        *reinterpret_cast<volatile int*>(&c->ticket) = (epoch << 8) | kTicketClosed;
    }
    __syncwarp();
    __threadfence_block();
}

__device__ __forceinline__ void unit_done(SlotCtrl* c, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicAdd(&c->done, 1);
}

// Everything a warp needs to work on the plot staged in this CTA, rebuilt from the kernel parameters and the plot's
// metadata (so that the out-of-line slot_work below takes three pointers instead of references to the caller's
// structures, which would otherwise be forced into local memory).
template <bool Z3>
struct PlotView {
    WindowAcc W;
    PlotCtx pc;
    __device__ __forceinline__ PlotView(const IcpParams& P, const PlotMeta& pm, unsigned char* smem, const SmemLayout& L)
        : W{reinterpret_cast<double2*>(smem + L.w_xy), reinterpret_cast<double*>(smem + L.w_z),
            reinterpret_cast<unsigned*>(smem + L.w_cell), reinterpret_cast<int*>(smem + L.rowoff),
            reinterpret_cast<int*>(smem + L.rowdelta), P.grid.orig, P.grid.rec,
            pm.wx0, pm.wy0, pm.wx1, pm.wy1, pm.wx1 - pm.wx0, pm.wy1 - pm.wy0},
          pc{reinterpret_cast<double2*>(smem + L.s_u), reinterpret_cast<double*>(smem + L.s_z), pm.n, pm.fixed_k,
             pm.ubx, pm.uby} {}
};

// Take work units of slot `s` until its epoch has none left (called by helpers and by the slot's own lead).  Out of
// line on purpose: the one-warp-per-ICP code around the call sites stays exactly what it is without helpers.
template <bool Z3>
__device__ __noinline__ void slot_work(const IcpParams* Pp, const PlotMeta* pmp, unsigned char* smem, int npad,
                                       int win_ok, int s, int lane) {
    const IcpParams& P = *Pp;
    const PlotMeta pm = *pmp;
    const SmemLayout L = smem_layout(npad, Z3, P.slots, P.wcap_pts, P.wcap_cells, P.wcap_rows);
    const PlotView<Z3> V(P, pm, smem, L);
    const GridView& G = P.grid;
    SlotCtrl* c = reinterpret_cast<SlotCtrl*>(smem + L.ctrl) + s;
    double* sd2 = reinterpret_cast<double*>(smem + L.sd2) + (size_t)s * npad;
    int* snn = reinterpret_cast<int*>(smem + L.snn) + (size_t)s * npad;
    __half* ssl = reinterpret_cast<__half*>(smem + L.ssl) + (size_t)s * npad;
    const unsigned short* sord = reinterpret_cast<const unsigned short*>(smem + L.sord) + (size_t)s * npad;
#pragma unroll 1
    for (;;) {
        const int e = grab_unit(c, lane);
        if (e < 0) break;
        __threadfence_block();
        const Pose pose = c->pose;
        if (c->mode == 0) {
            (void)nn_round<Z3, true>(G, V.W, win_ok != 0, V.pc, pose, sd2, snn, ssl, sord, e, c->nlist, lane,
                                     c->have_prev != 0);
        } else {
            unsigned ng = 0;
            nn_deferred_chunk<Z3, true>(G, V.W, V.pc, pose, sd2, snn, sord, e * 32, c->ndef, lane, ng);
            ng = __reduce_add_sync(kFull, ng);
            if (lane == 0 && ng) atomicAdd(&c->nglob, (int)ng);
        }
        unit_done(c, lane);
    }
}

// Optional per-pass trace (IcpParams::trace_cap > 0; tests only): for every tree the ORIGINAL target row of its nearest
// neighbour (what `tree.query` returns at ficp.py:70), the squared distance, whether the tree is in the trimmed subset
// (`argsort(d)[:k]`, ficp.py:62-63,133), plus k and FRMSD of the pass.  Out of line and behind one warp-uniform branch
// per pass, so the kernel is unchanged when the trace is off.
template <bool Z3>
__device__ __noinline__ void trace_pass(const IcpParams* Pp, const PlotMeta* pmp, unsigned char* smem, int npad, int slot,
                                        int lane, int pass_no, long long icp, int k, double f, double thr, int thr_idx) {
    const IcpParams& P = *Pp;
    if (pass_no >= P.trace_cap) return;
    const PlotMeta pm = *pmp;
    const SmemLayout L = smem_layout(npad, Z3, P.slots, P.wcap_pts, P.wcap_cells, P.wcap_rows);
    const PlotView<Z3> V(P, pm, smem, L);
    const double* sd2 = reinterpret_cast<const double*>(smem + L.sd2) + (size_t)slot * npad;
    const int* snn = reinterpret_cast<const int*>(smem + L.snn) + (size_t)slot * npad;
    const size_t rec = ((size_t)icp * P.trace_cap + pass_no);
    const size_t base = rec * P.trace_stride;
    for (int i = lane; i < pm.n; i += 32) {
        const int code = snn[i];
        int orig = -1;
        if (code != -1) {
            const int gpos = (code < 0) ? (code & 0x7FFFFFFF) : V.W.global_pos(code & 0xFFFF);
            orig = grid_orig(P.grid, gpos);
        }
        const double d2 = sd2[i];
        P.tr_idx[base + i] = orig;
        P.tr_d2[base + i] = d2;
        P.tr_in[base + i] = (k > 0 && (d2 < thr || (d2 == thr && i <= thr_idx))) ? 1 : 0;
    }
    if (lane == 0) { P.tr_k[rec] = k; P.tr_f[rec] = f; }
}

// Nearest neighbours of one pass (ficp.py:65-71) for the ICP of slot `ctrl`: skip test for every query (passes after
// the first), then the search of the queries that failed it, in rounds of 32 list entries.  ELASTIC: when enough warps
// of the CTA have no ICP of their own (`*sh_active <= dyn_leads`: the batch is smaller than the machine, or the plot is
// running out of hypotheses) the search rounds are handed out through the slot's ticket so that those warps take some;
// skip test, trimming and fit stay with the lead warp in unchanged arithmetic, so results do not depend on who
// computed a round.
template <int E, bool Z3, bool ELASTIC>
__device__ __forceinline__ void icp_nn_phase(const GridView& G, const WindowAcc& W, bool win_ok, const PlotCtx& pc,
                                             const Pose& P, const Pose& D, double* __restrict__ sd2,
                                             int* __restrict__ snn, __half* __restrict__ ssl,
                                             unsigned short* __restrict__ sord, SlotCtrl* ctrl, const int* sh_active,
                                             const IcpParams* Pp, const PlotMeta* pmp, unsigned char* smem, int slot,
                                             int& epoch, int lane, bool have_prev, unsigned& n_global,
                                             unsigned& n_searched, unsigned& n_deferred) {
    const int n = pc.n;
    const unsigned lt_mask = (1u << lane) - 1u;
    // ---- search list: every query on the first pass (identity list, written when the ICP starts), afterwards the
    // queries whose slack does not cover the pose update
    int n_list = n;
    if (have_prev) {
        n_list = 0;
#pragma unroll 1
        for (int e = 0; e * 32 < n; ++e) {
            const int need = nn_test_round<Z3>(W, pc, P, D, sd2, snn, ssl, e, lane);
            const unsigned m = __ballot_sync(kFull, need >= 0);
            if (need >= 0) sord[n_list + __popc(m & lt_mask)] = (unsigned short)need;
            n_list += __popc(m);
        }
        __syncwarp();
    }
    n_searched += (unsigned)n_list;
    const int rounds = (n_list + 31) >> 5;
    int n_def = 0;
    if (ELASTIC && rounds > 1 && ld_volatile_uniform(sh_active, lane) <= Pp->dyn_leads) {
        if (lane == 0) {
            ctrl->pose = P;
            ctrl->have_prev = have_prev ? 1 : 0;
            ctrl->nlist = n_list;
        }
        open_epoch(ctrl, epoch, 0, rounds, lane);
        slot_work<Z3>(Pp, pmp, smem, 32 * E, win_ok ? 1 : 0, slot, lane);
        // wait for the rounds the helpers took, then compact the flagged queries into `sord` (all consumed by now)
        close_epoch(ctrl, epoch, rounds, lane);
#pragma unroll 1
        for (int e = 0; e < rounds; ++e) {
            const int p = e * 32 + lane;
            int defer = -1;
            if (p < n_list) {
                const int i = sord[p];
                const long long b = __double_as_longlong(sd2[i]);
                if (b < 0) defer = i | ((b == __double_as_longlong(-kInf)) ? 0x8000 : 0);
            }
            const unsigned m = __ballot_sync(kFull, defer >= 0);
            if (defer >= 0) sord[n_def + __popc(m & lt_mask)] = (unsigned short)defer;
            n_def += __popc(m);
        }
        __syncwarp();
        if (n_def > 32) {
            // the deferred list is handed out as well, in chunks of 32
            const int chunks = (n_def + 31) >> 5;
            if (lane == 0) ctrl->ndef = n_def;
            n_deferred += (unsigned)n_def;
            open_epoch(ctrl, epoch, 1, chunks, lane);
            slot_work<Z3>(Pp, pmp, smem, 32 * E, win_ok ? 1 : 0, slot, lane);
            close_epoch(ctrl, epoch, chunks, lane);
            n_def = 0;
        }
    } else {
        // in-order rounds: the deferred list reuses the already-consumed entries of `sord`
#pragma unroll 1
        for (int e = 0; e < rounds; ++e) {
            const int defer = nn_round<Z3, false>(G, W, win_ok, pc, P, sd2, snn, ssl, sord, e, n_list, lane, have_prev);
            const unsigned m = __ballot_sync(kFull, defer >= 0);
            FICP_ASSERT(n_def + __popc(m) <= e * 32 + 32);
            if (defer >= 0) sord[n_def + __popc(m & lt_mask)] = (unsigned short)defer;
            n_def += __popc(m);
        }
    }
    __syncwarp();
    n_deferred += (unsigned)n_def;
#pragma unroll 1
    for (int base = 0; base < n_def; base += 32)
        nn_deferred_chunk<Z3, ELASTIC>(G, W, pc, P, sd2, snn, sord, base, n_def, lane, n_global);
    __syncwarp();
}

// A warp without an ICP of its own: take work units from the passes in flight in this CTA until every lead is done.
template <bool Z3>
__device__ __forceinline__ void icp_help(const IcpParams* Pp, const PlotMeta* pmp, unsigned char* smem, SlotCtrl* ctrls,
                                         int npad, int win_ok, int slots, const int* sh_active, int warp, int lane) {
    int s = warp % slots;
    unsigned idle = 0;
#pragma unroll 1
    while (ld_volatile_uniform(sh_active, lane) > 0) {
        bool open = false;
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < slots; ++k) {
                const int t = ld_volatile(&ctrls[s].ticket);
                if ((t & 0xFF) < ld_volatile(&ctrls[s].count)) { open = true; break; }
                s = (s + 1 == slots) ? 0 : s + 1;
            }
        }
        open = __shfl_sync(kFull, (int)open, 0) != 0;
        s = __shfl_sync(kFull, s, 0);
        if (open) {
            slot_work<Z3>(Pp, pmp, smem, npad, win_ok, s, lane);
            idle = 0;
        } else {
            idle = min(idle + 1, 8u);
            __nanosleep(100u << (idle >> 1));  // back off (up to 1.6 us): spinning warps compete with the leads for issue slots
        }
    }
}

// FRMSD trimming of one pass (ficp.py:62-63,73-86), by the ICP's (lead) warp.
template <int E>
__device__ __forceinline__ PassOut icp_trim_phase(const PlotCtx& pc, const double* __restrict__ s_g,
                                                  const double* __restrict__ g_c, const double* __restrict__ sd2,
                                                  int lane, unsigned& n_fix) {
    using C = LaneCfg<E>;
    const int n = pc.n;

    // ---- trim order: packed keys (monotone 32-IDXBITS-bit code of d2 | index), warp bitonic sort ----
    unsigned key[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        const unsigned fb = __float_as_uint(__double2float_rd(sd2[i]));
        key[e] = (i < n) ? (((fb >> (C::kIdxBits - 1)) << C::kIdxBits) | (unsigned)i) : 0xFFFFFFFFu;
    }
    warp_bitonic_sort<E>(key, lane);
    double dd[E];
#pragma unroll
    for (int r = 0; r < E; ++r) dd[r] = sd2[key[r] & C::kIdxMask];

    // ---- exact fix-up: points whose quantised codes collide are re-ordered by the true (d2, index) ----
    // The code is a monotone function of d2, so the sorted keys are already in exact order wherever neighbouring codes
    // differ; only runs of EQUAL codes can be out of order.  Most passes have none: one integer compare per neighbour pair
    // decides, and the fix-up (5 KB of straight-line fp64 compare-exchanges - instruction fetch is this kernel's top
    // stall) is skipped.
    bool collide = false;
#pragma unroll
    for (int r = 0; r + 1 < E; ++r)      // (padding slots all carry the key 0xFFFFFFFF: not a collision)
        collide = collide || (((key[r] ^ key[r + 1]) >> C::kIdxBits) == 0u && key[r + 1] != 0xFFFFFFFFu);
    {
        const unsigned nk0 = __shfl_down_sync(kFull, key[0], 1);
        collide = collide || (lane < 31 && ((key[E - 1] ^ nk0) >> C::kIdxBits) == 0u && nk0 != 0xFFFFFFFFu);
    }
    if (__any_sync(kFull, collide))
    for (;;) {
        bool sw = false;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int r = par; r + 1 < E; r += 2) {
                if (key_greater(dd[r], key[r], dd[r + 1], key[r + 1])) {
                    const double td = dd[r]; dd[r] = dd[r + 1]; dd[r + 1] = td;
                    const unsigned tk = key[r]; key[r] = key[r + 1]; key[r + 1] = tk;
                    sw = true;
                }
            }
        }
        if (E >= 2) {
            // boundary pair: my last element vs the next lane's first
            const double nd = __shfl_down_sync(kFull, dd[0], 1);
            const unsigned nk = __shfl_down_sync(kFull, key[0], 1);
            const double pd = __shfl_up_sync(kFull, dd[E - 1], 1);
            const unsigned pk = __shfl_up_sync(kFull, key[E - 1], 1);
            const bool hi = (lane < 31) && key_greater(dd[E - 1], key[E - 1], nd, nk);
            const bool lo = (lane > 0) && key_greater(pd, pk, dd[0], key[0]);
            if (hi) { dd[E - 1] = nd; key[E - 1] = nk; sw = true; }
            if (lo) { dd[0] = pd; key[0] = pk; sw = true; }
        } else {
            // one element per lane: odd-even transposition across lanes
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const bool left = ((lane & 1) == par);  // left member of the pair (lane, lane + 1)
                const int partner = left ? lane + 1 : lane - 1;
                const bool valid = (partner >= 0) && (partner < 32);
                const double od = __shfl_sync(kFull, dd[0], partner & 31);
                const unsigned ok2 = __shfl_sync(kFull, key[0], partner & 31);
                if (valid) {
                    const bool doswap = left ? key_greater(dd[0], key[0], od, ok2) : key_greater(od, ok2, dd[0], key[0]);
                    if (doswap) { dd[0] = od; key[0] = ok2; sw = true; }
                }
            }
        }
        if (!__any_sync(kFull, sw)) break;
        ++n_fix;
    }

    // ---- inclusive prefix sums S_k of d2 in trim order (in place: dd[r] becomes S at position lane*E + r) ----
    double run = 0.0;
#pragma unroll
    for (int r = 0; r < E; ++r) { run += dd[r]; dd[r] = run; }
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    double excl = __shfl_up_sync(kFull, inc, 1);
    if (lane == 0) excl = 0.0;
#pragma unroll
    for (int r = 0; r < E; ++r) dd[r] = excl + dd[r];
    double (&s)[E] = dd;

    // ---- subset size ----
    PassOut out;
    int kstar;
    double fstar = kInf, rstar = 0.0;
    if (pc.fixed_k > 0) {
        kstar = pc.fixed_k;
    } else {
        // filter with G(k) = S_k * (c_k^2 / k) ~ FRMSD(k)^2 (one multiply per k) ...
        double gbest = kInf;
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int p = lane * E + r;
            const double gr = s[r] * s_g[r * 32 + lane];
            if (p < n && gr < gbest) gbest = gr;     // no NaN handling needed: a NaN (0 * inf of a padded slot) never wins
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gbest = fmin(gbest, __shfl_xor_sync(kFull, gbest, o));
        // ... then evaluate the reference's exact expression c_k * sqrt(S_k / k) only for the k whose G is
        // within rounding distance of the minimum; first strict minimum wins (ficp.py:84)
        const double gthr = gbest * (1.0 + 1e-12);
        unsigned cand = 0u;  // bit r: position lane*E + r is within rounding distance of the minimum
#pragma unroll
        for (int r = 0; r < E; ++r) {
            const int p = lane * E + r;
            if (p < n && s[r] * s_g[r * 32 + lane] <= gthr) cand |= (1u << r);
        }
        int kb = INT_MAX;
        // one copy of the divide + square root: lanes walk their (almost always single) candidates in ascending k
        while (__any_sync(kFull, cand != 0u)) {
            if (cand != 0u) {
                const int r = __ffs(cand) - 1;
                cand &= cand - 1u;
                double sr = s[0];
#pragma unroll
                for (int rr = 1; rr < E; ++rr)
                    if (rr == r) sr = s[rr];
                const int k = lane * E + r + 1;
                const double rm = sqrt(sr / (double)k);
                const double f = __ldg(g_c + r * 32 + lane) * rm;
                if (f < fstar) { fstar = f; kb = k; rstar = rm; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double of = __shfl_xor_sync(kFull, fstar, o);
            const int ok = __shfl_xor_sync(kFull, kb, o);
            const double orr = __shfl_xor_sync(kFull, rstar, o);
            if (of < fstar || (of == fstar && ok < kb)) { fstar = of; kb = ok; rstar = orr; }
        }
        kstar = (kb == INT_MAX) ? 0 : kb;
    }
    out.k = kstar;
    if (kstar == 0) {
        out.f = kInf; out.rmse = 0.0; out.thr = -1.0; out.thr_idx = -1;
        return out;
    }
    // the k-th element in trim order defines the inlier set {(d2, i) <= (thr, thr_idx)}
    const int pstar = kstar - 1, lstar = pstar / E, rsel = pstar % E;
    double ssel = s[0];
    unsigned ksel = key[0];
#pragma unroll
    for (int r = 1; r < E; ++r)
        if (r == rsel) { ksel = key[r]; ssel = s[r]; }
    out.thr_idx = (int)(__shfl_sync(kFull, ksel, lstar) & C::kIdxMask);
    FICP_ASSERT(out.thr_idx < n && kstar >= 1 && kstar <= n);
    out.thr = sd2[out.thr_idx];
    if (pc.fixed_k > 0) {
        const double sk = __shfl_sync(kFull, ssel, lstar);
        rstar = sqrt(sk / (double)kstar);
        fstar = __ldg(g_c + rsel * 32 + lstar) * rstar;
    }
    out.f = fstar;
    out.rmse = rstar;
    return out;
}

// Closed-form rigid fit on the trimmed subset and composition into the pose (arithmetic: icp_shared.cuh).
template <int E, bool Z3>
__device__ __forceinline__ void icp_fit(const GridView& G, const WindowAcc& W, const PlotCtx& pc, Pose& P, Pose& D,
                                        const PassOut& po, const double* __restrict__ sd2,
                                        const int* __restrict__ snn, int lane, int allow_reflection) {
    const int n = pc.n;
    // shift point: the plot centroid under the current pose (keeps the running sums well conditioned)
    double ax, ay;
    fit_shift(P, pc.ubx, pc.uby, ax, ay);
    FitSums s = fit_zero();
#pragma unroll 2
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        if (i < n) {
            const double d2 = sd2[i];
            if (d2 < po.thr || (d2 == po.thr && i <= po.thr_idx)) {
                double qx, qy;
                pose_apply(P, pc.s_u[i], qx, qy);
                const int code = snn[i];
                FICP_ASSERT(code != -1 && ((code < 0) ? ((code & 0x7FFFFFFF) < G.m) : ((code & 0xFFFF) < W.rowoff[W.wh])));
                const double2 t = corr_xy(G, W, code);
                fit_term(s, qx, qy, t.x, t.y, ax, ay);
            }
        }
    }
    fit_reduce(s);
    fit_solve(s, po.k, allow_reflection, ax, ay, P, D);
}

template <int E, bool Z3, int NT, bool ELASTIC>
__global__ void __launch_bounds__(NT, 1) icp_kernel(const __grid_constant__ IcpParams P) {
    using C = LaneCfg<E>;
    constexpr int NPAD = C::kNPad;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // ELASTIC: the first `slots` warps of the CTA lead an ICP each; the others - and every lead that finds no
    // hypothesis left - help with the nearest-neighbour rounds of the passes in flight
    const int slots = ELASTIC ? P.slots : nwarps;
    const bool is_lead = warp < slots;
    const int slot = is_lead ? warp : 0;
    const SmemLayout L = smem_layout(NPAD, Z3, slots, P.wcap_pts, P.wcap_cells, P.wcap_rows);
    double2* s_u = reinterpret_cast<double2*>(smem + L.s_u);
    double2* w_xy = reinterpret_cast<double2*>(smem + L.w_xy);
    double* s_z = reinterpret_cast<double*>(smem + L.s_z);
    double* w_z = reinterpret_cast<double*>(smem + L.w_z);
    double* s_g = reinterpret_cast<double*>(smem + L.s_g);
    double* sd2 = reinterpret_cast<double*>(smem + L.sd2) + (size_t)slot * NPAD;
    int* snn = reinterpret_cast<int*>(smem + L.snn) + (size_t)slot * NPAD;
    unsigned short* sord = reinterpret_cast<unsigned short*>(smem + L.sord) + (size_t)slot * NPAD;
    __half* ssl = reinterpret_cast<__half*>(smem + L.ssl) + (size_t)slot * NPAD;
    SlotCtrl* ctrl = reinterpret_cast<SlotCtrl*>(smem + L.ctrl) + slot;
    unsigned* w_cell = reinterpret_cast<unsigned*>(smem + L.w_cell);
    int* rowoff = reinterpret_cast<int*>(smem + L.rowoff);
    int* rowdelta = reinterpret_cast<int*>(smem + L.rowdelta);
    int* rowg = reinterpret_cast<int*>(smem + L.rowg);
    __shared__ int sh_slice;
    __shared__ int sh_exhausted;
    __shared__ int sh_win_ok;
    __shared__ int sh_active;  // ELASTIC: warps leading an ICP or still entitled to pull one

    const GridView& G = P.grid;
    if (ELASTIC && is_lead && lane == 0) { ctrl->ticket = kTicketClosed; ctrl->count = 0; }
    int epoch = 0;
    int staged_plot = -1;
    unsigned long long acc_passes = 0, acc_global = 0, acc_fix = 0, acc_queries = 0, acc_searched = 0, acc_deferred = 0;

    for (;;) {
        __syncthreads();  // everyone is done with the previous slice (and with sh_slice)
        if (threadIdx.x == 0) {
            const int t = atomicAdd(P.slice_counter, 1);
            sh_slice = t;
            sh_exhausted = (t < P.n_slices) ? (atomicAdd(P.hyp_counter + (t % P.n_plots), 0) >= P.n_hyp_local) : 0;
            sh_active = slots;
        }
        __syncthreads();
        const int slice = sh_slice;
        if (slice >= P.n_slices) break;
        // tickets are dealt round-robin over the plots: every CTA may help any plot, and a ticket for a plot
        // whose hypotheses are all taken costs one atomic read (no staging)
        const int plot = slice % P.n_plots;
        if (sh_exhausted) continue;
        const PlotMeta pm = P.plots[plot];

        if (plot != staged_plot) {
            // ---- stage the plot: source rows, weight tables, window of grid cells ----
            for (int i = threadIdx.x; i < NPAD; i += blockDim.x) {
                s_u[i] = (i < pm.n) ? P.src_u[pm.off + i] : make_double2(0.0, 0.0);
                if (Z3) s_z[i] = (i < pm.n) ? P.src_z[pm.off + i] : 0.0;
            }
            const double* tab = P.tabs + (size_t)pm.tab * P.n_stages * 2 * NPAD;
            for (int i = threadIdx.x; i < P.n_stages * NPAD; i += blockDim.x)
                s_g[i] = tab[(size_t)(i / NPAD) * 2 * NPAD + (i % NPAD)];
            const int ww = pm.wx1 - pm.wx0, wh = pm.wy1 - pm.wy0;
            bool ok = (ww > 0 && wh > 0 && (long long)ww * wh <= P.wcap_cells && wh <= P.wcap_rows && G.m > 0);
            if (ok) {
                for (int r = threadIdx.x; r < wh; r += blockDim.x) {
                    const size_t rowbase = (size_t)(pm.wy0 + r) * G.g.gw;
                    const unsigned gs = G.cell_start[rowbase + pm.wx0], ge = G.cell_start[rowbase + pm.wx1];
                    rowg[r] = (int)gs;
                    rowdelta[r] = (int)(ge - gs);  // temporarily: the row's point count
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (ok) {
                    int o = 0;
                    for (int r = 0; r < wh; ++r) {
                        const int cnt = rowdelta[r];
                        rowoff[r] = o;
                        rowdelta[r] = rowg[r] - o;
                        o += cnt;
                    }
                    rowoff[wh] = o;
                    if (o > P.wcap_pts || o > 32767) ok = false;  // window positions are packed in 15 bits
                }
                sh_win_ok = ok ? 1 : 0;
                if (!ok) atomicAdd(P.stats + 2, 1ull);
            }
            __syncthreads();
            ok = (sh_win_ok != 0);
            if (ok) {
                for (int c = threadIdx.x; c < ww * wh; c += blockDim.x) {
                    const int r = c / ww, col = c - r * ww;
                    const size_t g = (size_t)(pm.wy0 + r) * G.g.gw + pm.wx0 + col;
                    const unsigned a = G.cell_start[g], b = G.cell_start[g + 1];
                    FICP_ASSERT(b - a < 65536u && rowoff[r] + (int)(a - (unsigned)rowg[r]) < 32768);
                    w_cell[c] = (unsigned)(rowoff[r] + (int)(a - (unsigned)rowg[r])) | ((b - a) << 16);
                }
                for (int r = warp; r < wh; r += nwarps) {
                    const int cnt = rowoff[r + 1] - rowoff[r], gs = rowg[r], lo = rowoff[r];
                    for (int j = lane; j < cnt; j += 32) {
                        w_xy[lo + j] = grid_xy(G, gs + j);
                        if (Z3) w_z[lo + j] = grid_z(G, gs + j);
                    }
                }
            }
            staged_plot = plot;
            __syncthreads();
        }
        const bool win_ok = (sh_win_ok != 0);
        const WindowAcc W{w_xy, w_z, w_cell, rowoff, rowdelta, G.orig, G.rec,
                          pm.wx0, pm.wy0, pm.wx1, pm.wy1, pm.wx1 - pm.wx0, pm.wy1 - pm.wy0};
        const PlotCtx pc{s_u, s_z, pm.n, pm.fixed_k, pm.ubx, pm.uby};
        const double* g_ctab = P.tabs + (size_t)pm.tab * P.n_stages * 2 * NPAD;  // [stage][0]=g [stage][1]=c

        // ---- lead warps pull hypotheses of this plot until none are left ----
        for (;;) {
            if (!is_lead) break;
            int j = 0;
            if (lane == 0) j = atomicAdd(P.hyp_counter + plot, 1);
            j = __shfl_sync(kFull, j, 0);
            if (j >= P.n_hyp_local) break;
            const int h = P.hyp_begin + j * P.hyp_stride;
            const double* hr = P.hyp + (size_t)h * 6;
            Pose pose{hr[0], hr[1], hr[2], hr[3], dadd(pm.cinx, hr[4]), dadd(pm.ciny, hr[5])};
            unsigned n_global = 0, n_fix = 0, n_searched = 0, n_deferred = 0;
            int passes = 0;
            Pose dpose{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // pose of the coming pass minus pose of the previous pass
            if (ELASTIC && lane == 0) ctrl->nglob = 0;
            // first pass: identity order; padding slots never change
            for (int i = lane; i < NPAD; i += 32) {
                sord[i] = (unsigned short)i;
                if (i >= pm.n) { sd2[i] = kInf; snn[i] = -1; }
            }
            __syncwarp();
            PassOut po{0, kInf, 0.0, -1.0, -1};
            for (int st = 0; st < P.n_stages; ++st) {
                const double* sg = s_g + (size_t)st * NPAD;
                const double* gc = g_ctab + ((size_t)st * 2 + 1) * NPAD;
                // ficp.py:122-147 with ONE call site for the pass:
                //   pass; [first: k==0 -> stage ends | else: converged -> stage ends]; it == max -> ends; fit; repeat
                double cur = 0.0;
                int it = 0;
                bool first = true;
                for (;;) {
                    icp_nn_phase<E, Z3, ELASTIC>(G, W, win_ok, pc, pose, dpose, sd2, snn, ssl, sord, ctrl, &sh_active, &P,
                                                 P.plots + plot, smem, slot, epoch, lane, passes > 0, n_global, n_searched, n_deferred);
                    dpose = Pose{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // a stage may end without a fit: same pose again
                    po = icp_trim_phase<E>(pc, sg, gc, sd2, lane, n_fix);
                    if (P.trace_cap > 0)
                        trace_pass<Z3>(&P, P.plots + plot, smem, NPAD, slot, lane, passes, (long long)plot * P.n_hyp_local + j,
                                       po.k, po.f, po.thr, po.thr_idx);
                    ++passes;
                    if (first) {
                        if (po.k == 0) break;  // ficp.py:125-126
                        cur = po.f;
                        first = false;
                    } else {
                        if (cur - po.f <= P.threshold) break;  // also stops on a regression, keeping the pose (ficp.py:142)
                        cur = po.f;
                        ++it;
                    }
                    if (it >= P.max_iter) break;
                    icp_fit<E, Z3>(G, W, pc, pose, dpose, po, sd2, snn, lane, P.allow_reflection);
                    __syncwarp();
                }
                __syncwarp();
            }
            // ---- results ----
            n_global = __reduce_add_sync(kFull, n_global);  // counted per lane
            if (ELASTIC) n_global += (unsigned)ld_volatile_uniform(&ctrl->nglob, lane);
            const size_t ridx = (size_t)plot * P.n_hyp_local + j;
            if (lane == 0) {
                HypResult r;
                r.m00 = pose.m00; r.m01 = pose.m01; r.m10 = pose.m10; r.m11 = pose.m11;
                r.cx = pose.cx; r.cy = pose.cy;
                r.frmsd = po.f; r.rmse = po.rmse; r.k = po.k; r.passes = passes;
                r.flags = (n_global ? 1 : 0) | (win_ok ? 0 : 2);
                r.pad = 0;
                P.results[ridx] = r;
                const float score = (po.k >= P.min_k && po.k > 0) ? (float)po.f : __int_as_float(0x7F800000);
                const unsigned long long bk = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)h;
                atomicMin(P.best_key + plot, bk);
            }
            if (P.final_xy && P.n_hyp_local == 1) {
                for (int i = lane; i < pm.n; i += 32) {
                    double qx, qy;
                    pose_apply(pose, s_u[i], qx, qy);
                    P.final_xy[(pm.off + i) * 2] = qx;
                    P.final_xy[(pm.off + i) * 2 + 1] = qy;
                }
            }
            acc_passes += passes; acc_global += n_global; acc_fix += n_fix; acc_searched += n_searched; acc_deferred += n_deferred;
            acc_queries += (unsigned long long)passes * pm.n;
            __syncwarp();
        }
        if (ELASTIC) {
            // no hypothesis left for this warp: help the passes still in flight in this CTA
            if (is_lead && lane == 0) atomicSub(&sh_active, 1);
            __syncwarp();
            if (P.dyn_leads >= 0)
                icp_help<Z3>(&P, P.plots + plot, smem, reinterpret_cast<SlotCtrl*>(smem + L.ctrl), NPAD, win_ok ? 1 : 0, slots,
                             &sh_active, warp, lane);
        }
    }
    if (lane == 0 && acc_passes) {
        atomicAdd(P.stats + 0, acc_passes);
        atomicAdd(P.stats + 1, acc_global);
        atomicAdd(P.stats + 3, acc_fix);
        atomicAdd(P.stats + 4, acc_queries);
        atomicAdd(P.stats + 5, acc_searched);
        atomicAdd(P.stats + 6, acc_deferred);
    }
}

template <int E, bool Z3>
struct KernelFor {
    // threads per CTA the kernel is compiled for (register budget = 64K / NT)
#ifndef FICP_NT16
#define FICP_NT16 512
#endif
    static constexpr int kNT = (E >= 32) ? 384 : (E >= 16) ? FICP_NT16 : 512;
};

template <int E, bool Z3>
int launch_one(const IcpParams& p, const IcpLaunch& l, cudaStream_t stream) {
    constexpr int NT = KernelFor<E, Z3>::kNT;
    if (l.warps * 32 > NT || l.slots < 1 || l.slots > l.warps || (!l.elastic && l.slots != l.warps)) {
        set_error("launch_icp: bad warps per CTA / lead warps for this instantiation");
        return kErrInvalid;
    }
    auto kern = l.elastic ? icp_kernel<E, Z3, NT, true> : icp_kernel<E, Z3, NT, false>;
    FICP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem));
    kern<<<l.ctas, l.warps * 32, l.smem, stream>>>(p);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

template <int E, bool Z3>
int occupancy_one(int warps, bool elastic, size_t smem, int* out) {
    constexpr int NT = KernelFor<E, Z3>::kNT;
    auto kern = elastic ? icp_kernel<E, Z3, NT, true> : icp_kernel<E, Z3, NT, false>;
    FICP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, warps * 32, smem));
    return kOk;
}

}  // namespace

size_t icp_smem_bytes(int e, bool z3, int slots, int wcap_pts, int wcap_cells, int wcap_rows) {
    return smem_layout(32 * e, z3, slots, wcap_pts, wcap_cells, wcap_rows).total;
}

int icp_max_warps(int e) { return (e >= 32) ? 12 : (e >= 16) ? FICP_NT16 / 32 : 16; }  // = KernelFor<E>::kNT / 32

#define FICP_DISPATCH(FN, ...)                                                                    \
    switch (e) {                                                                                  \
        case 1: return z3 ? FN<1, true>(__VA_ARGS__) : FN<1, false>(__VA_ARGS__);                 \
        case 2: return z3 ? FN<2, true>(__VA_ARGS__) : FN<2, false>(__VA_ARGS__);                 \
        case 4: return z3 ? FN<4, true>(__VA_ARGS__) : FN<4, false>(__VA_ARGS__);                 \
        case 8: return z3 ? FN<8, true>(__VA_ARGS__) : FN<8, false>(__VA_ARGS__);                 \
        case 16: return z3 ? FN<16, true>(__VA_ARGS__) : FN<16, false>(__VA_ARGS__);              \
        case 32: return z3 ? FN<32, true>(__VA_ARGS__) : FN<32, false>(__VA_ARGS__);              \
        default: set_error("launch_icp: unsupported elements-per-lane"); return kErrInvalid;      \
    }

int icp_max_ctas_per_sm(int e, bool z3, int warps, bool elastic, size_t smem, int* out) {
    FICP_DISPATCH(occupancy_one, warps, elastic, smem, out)
}

int launch_icp(const IcpParams& p, const IcpLaunch& l, cudaStream_t stream) {
    const int e = l.e;
    const bool z3 = l.z3;
    FICP_DISPATCH(launch_one, p, l, stream)
}

}  // namespace ficp
