// Shared definitions for the B200 Fractional-ICP kernels.
//
// Numerics contract (DESIGN.md "Numerics"): every quantity the reference computes in IEEE
// float64 (ficp.py:34-35) is float64 here as well - B200's FP64 pipe runs at half the FP32
// rate, so the hot path keeps bit-level parity of the NN indices with the reference instead of
// approximating in fp32.  Squared distances are evaluated WITHOUT fused multiply-add in the
// order ((dx*dx)+(dy*dy))+(dz*dz), which is bit-identical to scipy's cKDTree (ficp.py:69-70).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define FICP_HD __host__ __device__ __forceinline__
#else
#define FICP_HD inline
#endif

// Bounds / invariant checks of the kernels: compiled in only with -DFICP_DEBUG (compute-sanitizer is not
// available on the target pool, so the debug build is how out-of-range indices are hunted).
#if defined(FICP_DEBUG) && defined(__CUDA_ARCH__)
#include <cassert>
#define FICP_ASSERT(c) assert(c)
#else
#define FICP_ASSERT(c) ((void)0)
#endif

namespace ficp {

#if defined(__CUDA_ARCH__)
FICP_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
FICP_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
FICP_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
// host build (tests/hostcheck): compiled with -ffp-contract=off
FICP_HD double dmul(double a, double b) { return a * b; }
FICP_HD double dadd(double a, double b) { return a + b; }
FICP_HD double dsub(double a, double b) { return a - b; }
#endif

constexpr double kInf = __builtin_huge_val();  // +inf (nvcc and gcc both accept the builtin in constant expressions)

// Geometry of the uniform grid over the Layer-2 (CHM) points.
struct GridGeom {
    double x0, y0;     // lower-left corner of the grid
    double h, inv_h;   // cell edge and its reciprocal
    double eps;        // conservative slack for cell-box bounds (covers binning round-off)
    int gw, gh;        // cells per row / number of rows
    // True bounding box of ALL target points.  Normally the grid spans it.  For skewed targets (a stray placeholder
    // row at (0, 0) in UTM data, a far outlier, long thin tails) the grid spans a robust core extent instead and the
    // points outside are CLAMPED into the border cells (`clamped` != 0): a border cell then reaches outward without
    // bound, which the search's box tests account for; results stay exact either way.
    double tx0, tx1, ty0, ty1;
    int clamped, pad;
};

// Device view of a built target index (cell-sorted copy of the target + CSR cell table).  Two layouts:
//   XY  targets: xy[M] (16 B) + orig[M] (4 B)
//   XYZ targets: rec[M] = {x, y, z, bits(original index)} - 32 B, 32 B aligned, i.e. exactly one L2 sector per
//                candidate (the split 16 B + 8 B + 4 B arrays pulled 2-3 sectors each; profiles/r01_summary.md)
struct GridView {
    GridGeom g;
    const double2* xy;          // XY layout only
    const double4* rec;         // XYZ layout only (nullptr otherwise)
    const int* orig;            // XY layout only
    const unsigned* cell_start; // [gw*gh + 1] exclusive prefix of per-cell counts
    long long m;
};

FICP_HD double index_to_bits(int i) {
    long long v = i;
    double d;
#if defined(__CUDA_ARCH__)
    d = __longlong_as_double(v);
#else
    __builtin_memcpy(&d, &v, sizeof d);
#endif
    return d;
}
FICP_HD int bits_to_index(double d) {
#if defined(__CUDA_ARCH__)
    return (int)__double_as_longlong(d);
#else
    long long v;
    __builtin_memcpy(&v, &d, sizeof v);
    return (int)v;
#endif
}

FICP_HD int clamp_cell(double f, int n) {
    // f = (coord - origin) * inv_h ; robust to huge magnitudes (clamped before conversion)
    if (!(f > 0.0)) return 0;
    if (f >= (double)(n - 1)) return n - 1;
    return (int)f;
}

}  // namespace ficp
