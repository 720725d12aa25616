// Latency micro-benchmarks for the design of the CTA-per-ICP pass (B200): dependent FP64 chains, 64-bit shuffles,
// shared-memory round trips, CTA barriers at several CTA sizes, fp64 div / sqrt sequences.  One warp (or CTA) measured
// with clock64().   nvcc -O3 -gencode arch=compute_100a,code=sm_100a lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(double* out, long long* cyc, int mode) {
    double a = out[threadIdx.x], b = out[threadIdx.x + 32] + 1.0000001, c = 1e-9;
    __shared__ double sh[1024];
    sh[threadIdx.x] = a;
    __syncthreads();
    const int N = 512;
    long long t0 = clock64();
    if (mode == 0) { for (int i = 0; i < N; ++i) a = __dadd_rn(a, b); }
    else if (mode == 1) { for (int i = 0; i < N; ++i) a = __fma_rn(a, b, c); }
    else if (mode == 2) { for (int i = 0; i < N; ++i) a = __dmul_rn(a, b); }
    else if (mode == 3) { for (int i = 0; i < N; ++i) a = __dadd_rn(a, __shfl_xor_sync(0xFFFFFFFFu, a, 1)); }
    else if (mode == 4) { int j = threadIdx.x; for (int i = 0; i < N; ++i) { a = __dadd_rn(a, sh[j]); j = (j + (int)a) & 1023; } }
    else if (mode == 5) { for (int i = 0; i < N; ++i) a = __ddiv_rn(a, b); }
    else if (mode == 6) { for (int i = 0; i < N; ++i) a = __dsqrt_rn(a + 2.0); }
    else if (mode == 7) { for (int i = 0; i < N; ++i) { a = __dadd_rn(a, b); __syncthreads(); } }
    else if (mode == 8) { float f = (float)a; for (int i = 0; i < N; ++i) f = __fmaf_rn(f, 1.0000001f, 1e-9f); a = f; }
    else if (mode == 9) { for (int i = 0; i < N; ++i) { sh[threadIdx.x] = a; __syncthreads(); a = sh[threadIdx.x ^ 32] + 1.0; __syncthreads(); } }
    else if (mode == 10) { for (int i = 0; i < N; ++i) { bool p = a > (double)i; a += __syncthreads_or(p) ? 1.0 : 0.5; } }
    else if (mode == 11) { for (int i = 0; i < N; ++i) a = fmin(a, b) + 1e-3; }
    else if (mode == 12) { int v = (int)a; for (int i = 0; i < N; ++i) v = __reduce_min_sync(0xFFFFFFFFu, v + i); a = v; }
    else if (mode == 13) { for (int i = 0; i < N; ++i) { a = (a > b) ? a - 1.0 : a + 1.0; } }
    else if (mode == 14) { for (int i = 0; i < N; ++i) a = __drcp_rn(a + 3.0); }
    else if (mode == 15) { __shared__ int cnt; if (threadIdx.x == 0) cnt = 0; __syncthreads(); int v = 0; for (int i = 0; i < N; ++i) v += atomicAdd(&cnt, (v & 1) + 1); a = v; }
    else if (mode == 16) { float f = (float)a + 2.f; for (int i = 0; i < N; ++i) f = __fsqrt_ru(f + 1.5f); a = f; }
    else if (mode == 17) { const double* g = out; int j = threadIdx.x; for (int i = 0; i < N; ++i) { const double v = __ldcg(g + j); j = (j + 32 + (int)v) & 1023; a += v; } }
    else if (mode == 18) { for (int i = 0; i < N; ++i) { a = __dadd_rn(a, b); asm volatile("bar.sync 1, 128;"); } }
    else if (mode == 19) { for (int i = 0; i < N; ++i) { unsigned m = __ballot_sync(0xFFFFFFFFu, a > (double)i); a += __popc(m); } }
    else if (mode == 20) { float f = (float)a; for (int i = 0; i < N; ++i) { f = __double2float_ru(a); a = a + (double)f; } }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = (t1 - t0);
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 2048 * 8); cudaMalloc(&c, 8); cudaMemset(d, 0, 2048 * 8);
    const char* names[] = {"DADD chain", "DFMA chain", "DMUL chain", "DADD+SHFL64", "DADD+LDS (dependent addr)", "DDIV chain", "DSQRT chain", "DADD+__syncthreads",
                           "FFMA chain", "STS+sync+LDS+sync", "__syncthreads_or", "fmin+dadd", "redux.min.s32", "DSETP+sel+DADD", "DRCP chain", "smem atomicAdd (1 lane active/all lanes)", "fsqrt_ru chain", "L2 load (ldcg, dependent addr)", "DADD+bar.sync 1,128 (4 warps)", "ballot+popc+I2F+DADD", "F2F ru + F2D + DADD"};
    for (int mode = 0; mode < 21; ++mode)
        for (int nt : {32, 128, 256, 512, 1024}) {
            if (mode == 18 && nt != 128) continue;
            if (nt > 32 && !(mode == 7 || mode == 9 || mode == 10 || mode == 0 || mode == 1 || mode == 15 || mode == 18)) continue;
            k_chain<<<1, nt>>>(d, c, mode); cudaDeviceSynchronize();
            k_chain<<<1, nt>>>(d, c, mode); cudaDeviceSynchronize();
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%-28s threads %4d : %7.1f cycles per step\n", names[mode], nt, h / 512.0);
        }
    return 0;
}
