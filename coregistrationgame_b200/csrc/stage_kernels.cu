// Stand-alone stage kernels behind the per-method API of FractionalICP (the persistent kernel in
// icp_persistent.cu fuses the same stages for the batched path):
//
//   select_fraction  <- find_optimal_fraction / get_n_first_elements  (/root/reference/ficp.py:62-63,73-86)
//   fit_rigid2d      <- compute_optimal_transform_2d                  (ficp.py:89-110)
//   apply_xy         <- apply_transform_2d_xy_only                    (ficp.py:112-119)
//   sumsq            <- the sum inside frmsd                          (ficp.py:58-59)
//
// Each runs as ONE CTA: the inputs are a single plot (N <= 8192 trees); K = 2 Procrustes is eight
// running sums, so there is no dense contraction for tensor cores here.
#include "ficp_internal.h"

namespace ficp {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ double block_sum(double v, double* sh /* >= 32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += sh[i];  // same order in every thread -> identical bits
    return s;
}

__device__ __forceinline__ double sqdiff(const double* a, const double* b, int md) {
    const double dx = dsub(a[0], b[0]), dy = dsub(a[1], b[1]);
    double s = dadd(dmul(dx, dx), dmul(dy, dy));
    if (md == 3) {
        const double dz = dsub(a[2], b[2]);
        s = dadd(s, dmul(dz, dz));
    }
    return s;
}

// ---- select_fraction ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) select_fraction_kernel(const double* __restrict__ src, int ld_s,
                                                               const double* __restrict__ corr, int ld_c,
                                                               const double* __restrict__ dist, int n, int npad,
                                                               int md, const double* __restrict__ w, int fixed_k,
                                                               long long* __restrict__ k_out,
                                                               double* __restrict__ f_out,
                                                               int* __restrict__ order_out) {
    extern __shared__ double sm[];
    double* key = sm;
    double* pre = sm + npad;
    int* idx = reinterpret_cast<int*>(sm + 2 * npad);
    __shared__ double red[32];
    __shared__ double red_f[32];
    __shared__ int red_k[32];
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int i = tid; i < npad; i += nt) {
        key[i] = (i < n) ? dist[i] : kInf;
        idx[i] = i;
    }
    __syncthreads();
    // bitonic sort of (distance, index): the index makes the order stable (lowest index first)
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += nt) {
                const int i = ((t / j) * 2 * j) + (t % j);
                const int l = i + j;
                const bool asc = ((i & k) == 0);
                const double ka = key[i], kb = key[l];
                const int ia = idx[i], ib = idx[l];
                const bool gt = (ka > kb) || (ka == kb && ia > ib);
                if (gt == asc) {
                    key[i] = kb; key[l] = ka;
                    idx[i] = ib; idx[l] = ia;
                }
            }
            __syncthreads();
        }
    }
    if (order_out)
        for (int i = tid; i < n; i += nt) order_out[i] = idx[i];
    if (!src) return;

    // squared residuals in trim order, then an inclusive scan (thread t owns a contiguous chunk)
    for (int i = tid; i < npad; i += nt)
        pre[i] = (i < n) ? sqdiff(src + (size_t)idx[i] * ld_s, corr + (size_t)idx[i] * ld_c, md) : 0.0;
    __syncthreads();
    const int chunk = (npad + nt - 1) / nt;
    const int b0 = tid * chunk;
    double loc = 0.0;
    for (int j = 0; j < chunk; ++j)
        if (b0 + j < npad) loc += pre[b0 + j];
    // exclusive scan of `loc` over the block
    double inc = loc;
    const int l = tid & 31, wq = tid >> 5, nw = (nt + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, inc, o);
        if (l >= o) inc += t;
    }
    if (l == 31) red[wq] = inc;
    __syncthreads();
    double base = 0.0;
    for (int i = 0; i < wq && i < nw; ++i) base += red[i];
    double run = base + inc - loc;
    double fbest = kInf;
    int kbest = 0;
    for (int j = 0; j < chunk; ++j) {
        const int i = b0 + j;
        if (i < npad) {
            run += pre[i];
            if (i < n) {
                const int k = i + 1;
                const double f = w[i] * sqrt(run / (double)k);
                if (fixed_k > 0) {
                    if (k == fixed_k) { fbest = f; kbest = k; }
                } else if (f < fbest) {  // first strict minimum (ficp.py:84)
                    fbest = f;
                    kbest = k;
                }
            }
        }
    }
    // block arg-min: smallest value, ties to the smallest k
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(kFull, fbest, o);
        const int ok = __shfl_xor_sync(kFull, kbest, o);
        if (ok != 0 && (kbest == 0 || of < fbest || (of == fbest && ok < kbest))) { fbest = of; kbest = ok; }
    }
    if (l == 0) { red_f[wq] = fbest; red_k[wq] = kbest; }
    __syncthreads();
    if (tid == 0) {
        double f = kInf; int k = 0;
        for (int i = 0; i < nw; ++i)
            if (red_k[i] != 0 && (k == 0 || red_f[i] < f || (red_f[i] == f && red_k[i] < k))) { f = red_f[i]; k = red_k[i]; }
        *k_out = k;
        *f_out = (k == 0) ? kInf : f;
    }
}

// ---- select_fraction for plots above kSelectMaxN rows: the same three steps over GLOBAL scratch arrays ----------------
// (the reference accepts any N, ficp.py:73-86; np.argsort + np.cumsum there).  Sort: bitonic network on (distance, index)
// pairs - phases / steps whose partners lie within one tile of kSfTile elements run in shared memory (one launch for all
// phases k <= kSfTile, one per later phase), the wider steps are one grid-wide compare-exchange launch each
// (45 launches for 2^20 rows).  Scan + first strict minimum: one CTA, every thread a contiguous chunk, as above.
constexpr int kSfTile = 4096;

__device__ __forceinline__ void sf_cex(double& ka, int& ia, double& kb, int& ib, bool asc) {
    const bool gt = (ka > kb) || (ka == kb && ia > ib);
    if (gt == asc) {
        const double tk = ka; ka = kb; kb = tk;
        const int ti = ia; ia = ib; ib = ti;
    }
}

__global__ void __launch_bounds__(256) sf_init_kernel(const double* __restrict__ dist, int n, int npad, double* __restrict__ key,
                                                      int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    key[i] = (i < n) ? dist[i] : kInf;
    idx[i] = i;
}

// Steps j = j_hi, j_hi/2, ..., 1 of the phases k = k_lo, 2 k_lo, ..., k_hi inside tiles of kSfTile elements (j_hi < kSfTile;
// for k_lo < k_hi every phase starts at j = k/2, i.e. k_hi <= kSfTile).
__global__ void __launch_bounds__(1024) sf_tile_kernel(double* __restrict__ key, int* __restrict__ idx, int k_lo, int k_hi, int j_hi) {
    __shared__ double sk[kSfTile];
    __shared__ int si[kSfTile];
    const int base = blockIdx.x * kSfTile, tid = threadIdx.x;
    for (int i = tid; i < kSfTile; i += 1024) { sk[i] = key[base + i]; si[i] = idx[base + i]; }
    __syncthreads();
    for (int k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = (k_lo == k_hi) ? j_hi : (k >> 1); j > 0; j >>= 1) {
            for (int t = tid; t < kSfTile / 2; t += 1024) {
                const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
                sf_cex(sk[i], si[i], sk[l], si[l], ((base + i) & k) == 0);
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < kSfTile; i += 1024) { key[base + i] = sk[i]; idx[base + i] = si[i]; }
}

// One step (k, j) with j >= kSfTile: partners are j apart in global memory.
__global__ void __launch_bounds__(256) sf_step_kernel(double* __restrict__ key, int* __restrict__ idx, int npad, int k, int j) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (npad >> 1)) return;
    const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
    double ka = key[i], kb = key[l];
    int ia = idx[i], ib = idx[l];
    const bool gt = (ka > kb) || (ka == kb && ia > ib);
    if (gt == ((i & k) == 0)) { key[i] = kb; key[l] = ka; idx[i] = ib; idx[l] = ia; }
}

__global__ void __launch_bounds__(256) sf_residual_kernel(const double* __restrict__ src, int ld_s, const double* __restrict__ corr,
                                                          int ld_c, const int* __restrict__ idx, int n, int md,
                                                          double* __restrict__ pre, int* __restrict__ order_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = idx[i];
    if (order_out) order_out[i] = r;
    if (src) pre[i] = sqdiff(src + (size_t)r * ld_s, corr + (size_t)r * ld_c, md);
}

__global__ void __launch_bounds__(1024) sf_scan_min_kernel(const double* __restrict__ pre, int n, const double* __restrict__ w,
                                                           int fixed_k, long long* __restrict__ k_out, double* __restrict__ f_out) {
    __shared__ double red[32];
    __shared__ double red_f[32];
    __shared__ int red_k[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int chunk = (n + nt - 1) / nt;
    const int b0 = tid * chunk;
    double loc = 0.0;
    for (int j = 0; j < chunk; ++j)
        if (b0 + j < n) loc += pre[b0 + j];
    double inc = loc;
    const int l = tid & 31, wq = tid >> 5, nw = (nt + 31) >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, inc, o);
        if (l >= o) inc += t;
    }
    if (l == 31) red[wq] = inc;
    __syncthreads();
    double base = 0.0;
    for (int i = 0; i < wq && i < nw; ++i) base += red[i];
    double run = base + inc - loc;
    double fbest = kInf;
    int kbest = 0;
    for (int j = 0; j < chunk; ++j) {
        const int i = b0 + j;
        if (i < n) {
            run += pre[i];
            const int k = i + 1;
            const double f = w[i] * sqrt(run / (double)k);
            if (fixed_k > 0) {
                if (k == fixed_k) { fbest = f; kbest = k; }
            } else if (f < fbest) {  // first strict minimum (ficp.py:84)
                fbest = f;
                kbest = k;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(kFull, fbest, o);
        const int ok = __shfl_xor_sync(kFull, kbest, o);
        if (ok != 0 && (kbest == 0 || of < fbest || (of == fbest && ok < kbest))) { fbest = of; kbest = ok; }
    }
    if (l == 0) { red_f[wq] = fbest; red_k[wq] = kbest; }
    __syncthreads();
    if (tid == 0) {
        double f = kInf; int k = 0;
        for (int i = 0; i < nw; ++i)
            if (red_k[i] != 0 && (k == 0 || red_f[i] < f || (red_f[i] == f && red_k[i] < k))) { f = red_f[i]; k = red_k[i]; }
        *k_out = k;
        *f_out = (k == 0) ? kInf : f;
    }
}

// ---- rigid 2-D fit -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fit_rigid2d_kernel(const double* __restrict__ src, int ld_s,
                                                          const double* __restrict__ tgt, int ld_t,
                                                          const int* __restrict__ sel, int k, int allow_reflection,
                                                          double* __restrict__ T9) {
    __shared__ double red[32];
    double sx = 0, sy = 0, tx = 0, ty = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t r = sel ? (size_t)sel[i] : (size_t)i;
        sx += src[r * ld_s]; sy += src[r * ld_s + 1];
        tx += tgt[r * ld_t]; ty += tgt[r * ld_t + 1];
    }
    const double csx = block_sum(sx, red) / k, csy = block_sum(sy, red) / k;
    const double ctx = block_sum(tx, red) / k, cty = block_sum(ty, red) / k;
    double h00 = 0, h01 = 0, h10 = 0, h11 = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t r = sel ? (size_t)sel[i] : (size_t)i;
        const double ux = src[r * ld_s] - csx, uy = src[r * ld_s + 1] - csy;
        const double vx = tgt[r * ld_t] - ctx, vy = tgt[r * ld_t + 1] - cty;
        h00 += ux * vx; h01 += ux * vy; h10 += uy * vx; h11 += uy * vy;
    }
    h00 = block_sum(h00, red); h01 = block_sum(h01, red); h10 = block_sum(h10, red); h11 = block_sum(h11, red);
    if (threadIdx.x == 0) {
        double r00, r01, r10, r11;
        // reflection only when det(H) is negative beyond rounding noise (det == 0: SVD's choice is arbitrary)
    if (allow_reflection && (h00 * h11 - h01 * h10) < -1e-14 * (fabs(h00 * h11) + fabs(h01 * h10))) {
            const double a = h00 - h11, b = h01 + h10, nrm = hypot(a, b);
            const double c = (nrm == 0.0) ? 1.0 : a / nrm, s = (nrm == 0.0) ? 0.0 : b / nrm;
            r00 = c; r01 = s; r10 = s; r11 = -c;
        } else {
            const double a = h00 + h11, b = h01 - h10, nrm = hypot(a, b);
            const double c = (nrm == 0.0) ? 1.0 : a / nrm, s = (nrm == 0.0) ? 0.0 : b / nrm;
            r00 = c; r01 = -s; r10 = s; r11 = c;
        }
        T9[0] = r00; T9[1] = r01; T9[2] = ctx - (r00 * csx + r01 * csy);
        T9[3] = r10; T9[4] = r11; T9[5] = cty - (r10 * csx + r11 * csy);
        T9[6] = 0.0; T9[7] = 0.0; T9[8] = 1.0;
    }
}

__global__ void __launch_bounds__(256) apply_xy_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                       long long n, int ld, const double* __restrict__ T9) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = in[i * ld], y = in[i * ld + 1];
    for (int c = 2; c < ld; ++c) out[i * ld + c] = in[i * ld + c];  // other columns: bit-identical copy
    out[i * ld] = dadd(dadd(dmul(T9[0], x), dmul(T9[1], y)), T9[2]);
    out[i * ld + 1] = dadd(dadd(dmul(T9[3], x), dmul(T9[4], y)), T9[5]);
}

__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ a, int ld_a,
                                                    const double* __restrict__ b, int ld_b,
                                                    const int* __restrict__ sel, int k, int md,
                                                    double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t r = sel ? (size_t)sel[i] : (size_t)i;
        s += sqdiff(a + r * ld_a, b + r * ld_b, md);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
}

// Per-plot best registration of a finished batch, packed for the multi-GPU exchange (dist.py): kPackWords words per plot:
// [0] the packed key (fp32 score bits << 32 | hypothesis id), [1..10] the 80-byte result row of that hypothesis,
// [11] the hypothesis-iterations this GPU ran in the launch, [12..13] the translation of the WORLD-frame transform
// b = c - M centre (final = M p + b), evaluated like batch.compose_world_transforms (elementwise, no FMA) so that a
// receiver needs no plot centres.  One all_gather of these records is the whole exchange.
__global__ void __launch_bounds__(256) pack_best_kernel(const unsigned long long* __restrict__ best,
                                                        const HypResult* __restrict__ results, const PlotMeta* __restrict__ plots,
                                                        int n_plots, int n_hyp_local, int hyp_begin, int hyp_stride,
                                                        const unsigned long long* __restrict__ stats,
                                                        unsigned long long* __restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = t / kPackWords, w = t - p * kPackWords;
    if (p >= n_plots) return;
    const unsigned long long key = best[p];
    long long j = ((long long)(unsigned)(key & 0xFFFFFFFFull) - hyp_begin) / hyp_stride;
    if (j < 0 || j >= n_hyp_local) j = 0;
    const HypResult* row = results + (size_t)p * n_hyp_local + j;
    unsigned long long v;
    if (w == 0) {
        v = key;
    } else if (w == 11) {
        v = stats[0];
    } else if (w >= 12) {
        const double c0 = plots[p].cinx, c1 = plots[p].ciny;
        const double bw = (w == 12) ? __dsub_rn(row->cx, __dadd_rn(__dmul_rn(row->m00, c0), __dmul_rn(row->m01, c1)))
                                    : __dsub_rn(row->cy, __dadd_rn(__dmul_rn(row->m10, c0), __dmul_rn(row->m11, c1)));
        v = (unsigned long long)__double_as_longlong(bw);
    } else {
        v = reinterpret_cast<const unsigned long long*>(row)[w - 1];
    }
    dst[t] = v;
}

}  // namespace

// Raw plot rows (as the caller holds them: `ld` doubles per row) -> the batch's layout: u = (x, y) - centre of the row's plot
// as double2, z apart.  One warp per plot; the subtraction is the single IEEE operation the host pass of batch_prep.h (and the
// oracle's pre_transform) performs, so both routes give the same bits.
namespace {
template <bool Z3>
__global__ void __launch_bounds__(256) split_rows_kernel(const double* __restrict__ raw, int ld, const PlotMeta* __restrict__ plots,
                                                         int n_plots, double2* __restrict__ u, double* __restrict__ z) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_plots) return;
    const PlotMeta pm = plots[p];
    for (int i = threadIdx.x & 31; i < pm.n; i += 32) {
        const double* r = raw + (size_t)(pm.off + i) * ld;
        u[pm.off + i] = make_double2(r[0] - pm.cinx, r[1] - pm.ciny);
        if (Z3) z[pm.off + i] = r[2];
    }
}
}  // namespace

int launch_split_rows(const double* d_raw, int ld, const PlotMeta* d_plots, int n_plots, bool z3, double2* d_u, double* d_z,
                      cudaStream_t stream) {
    if (n_plots <= 0) return kOk;
    const int blocks = (n_plots + 7) / 8;
    if (z3) split_rows_kernel<true><<<blocks, 256, 0, stream>>>(d_raw, ld, d_plots, n_plots, d_u, d_z);
    else split_rows_kernel<false><<<blocks, 256, 0, stream>>>(d_raw, ld, d_plots, n_plots, d_u, d_z);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_pack_best(const unsigned long long* d_best, const HypResult* d_results, const PlotMeta* d_plots, int n_plots,
                     int n_hyp_local, int hyp_begin, int hyp_stride, const unsigned long long* d_stats,
                     unsigned long long* d_dst, cudaStream_t stream) {
    static_assert(sizeof(HypResult) == 80, "pack_best_kernel copies ten 8-byte words per row");
    if (n_plots <= 0) return kOk;
    const int n = n_plots * kPackWords;
    pack_best_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_best, d_results, d_plots, n_plots, n_hyp_local, hyp_begin, hyp_stride, d_stats, d_dst);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_select_fraction(const double* d_src, int ld_s, const double* d_corr, int ld_c, const double* d_dist,
                           int n, int md, const double* d_weights, int fixed_k, long long* d_k_out,
                           double* d_frmsd_out, int* d_order_out, cudaStream_t stream) {
    if (n <= 0) return kOk;
    if (n > kSelectLargeMaxN) {
        set_error("select_fraction: more than 2^24 points per plot are not supported");
        return kErrTooLarge;
    }
    int npad = 2;
    while (npad < n) npad <<= 1;
    if (n > kSelectMaxN) {
        // global-scratch path: (distance, index) pairs padded to a power of two, residuals in trim order
        double *key = nullptr, *pre = nullptr;
        int* idx = nullptr;
        cudaError_t e;
        if ((e = dev_alloc_t(&key, (size_t)npad, stream)) != cudaSuccess || (e = dev_alloc_t(&idx, (size_t)npad, stream)) != cudaSuccess ||
            (e = dev_alloc_t(&pre, (size_t)n, stream)) != cudaSuccess) {
            dev_free(key, stream); dev_free(idx, stream); dev_free(pre, stream);
            return cuda_fail(e, "select_fraction scratch", __FILE__, __LINE__);
        }
        auto done = [&](int rc) { dev_free(key, stream); dev_free(idx, stream); dev_free(pre, stream); return rc; };
        sf_init_kernel<<<(npad + 255) / 256, 256, 0, stream>>>(d_dist, n, npad, key, idx);
        const int tiles = npad / kSfTile;   // npad >= 16384
        sf_tile_kernel<<<tiles, 1024, 0, stream>>>(key, idx, 2, kSfTile, 0);
        for (int k = 2 * kSfTile; k <= npad; k <<= 1) {
            int j = k >> 1;
            for (; j >= kSfTile; j >>= 1) sf_step_kernel<<<((npad >> 1) + 255) / 256, 256, 0, stream>>>(key, idx, npad, k, j);
            sf_tile_kernel<<<tiles, 1024, 0, stream>>>(key, idx, k, k, j);
        }
        sf_residual_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_src, ld_s, d_corr, ld_c, idx, n, md, pre, d_order_out);
        if (d_src) sf_scan_min_kernel<<<1, 1024, 0, stream>>>(pre, n, d_weights, fixed_k, d_k_out, d_frmsd_out);
        const cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return done(cuda_fail(le, "select_fraction (large)", __FILE__, __LINE__));
        return done(kOk);
    }
    const size_t smem = (size_t)npad * (2 * sizeof(double) + sizeof(int));
    // per launch, not once per process: the attribute is per device (ficp_set_device may have moved us)
    FICP_CUDA(cudaFuncSetAttribute(select_fraction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kSelectMaxN * (2 * sizeof(double) + sizeof(int))));
    int nt = npad / 2;
    if (nt < 32) nt = 32;
    if (nt > 1024) nt = 1024;
    select_fraction_kernel<<<1, nt, smem, stream>>>(d_src, ld_s, d_corr, ld_c, d_dist, n, npad, md, d_weights, fixed_k,
                                                    d_k_out, d_frmsd_out, d_order_out);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_fit_rigid2d(const double* d_src, int ld_s, const double* d_tgt, int ld_t, const int* d_sel, int k,
                       int allow_reflection, double* d_T9, cudaStream_t stream) {
    fit_rigid2d_kernel<<<1, 256, 0, stream>>>(d_src, ld_s, d_tgt, ld_t, d_sel, k, allow_reflection, d_T9);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_apply_xy(const double* d_in, double* d_out, long long n, int ld, const double* d_T9, cudaStream_t stream) {
    if (n <= 0) return kOk;
    apply_xy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_in, d_out, n, ld, d_T9);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_sumsq(const double* d_a, int ld_a, const double* d_b, int ld_b, const int* d_sel, int k, int md,
                 double* d_out, cudaStream_t stream) {
    sumsq_kernel<<<1, 256, 0, stream>>>(d_a, ld_a, d_b, ld_b, d_sel, k, md, d_out);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ficp
