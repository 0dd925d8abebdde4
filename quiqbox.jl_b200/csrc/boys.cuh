// Boys function F_m(T) on the device.
//
// Replaces computeBoysSequence (src/Integration/Engines/BoysFunction.jl:67-77), which the
// reference evaluates per primitive component quartet through SpecialFunctions.gamma_inc.
// Two evaluators:
//   * boys_table<L>   -- hot path of the class kernels (L <= 8): 8-term Taylor expansion
//                        about the nearest point of a 1/8-spaced table whose 9 needed
//                        columns are staged in shared memory once per (persistent) block,
//                        exp(-T) from the table + a Taylor tail (no SFU/exp call), then the
//                        reference's own downward recursion (BoysFunction.jl:33-40).
//                        T >= QBX_BOYS_TMAX switches to the asymptotic form by a select, not
//                        a branch (a warp mixes near and far kets).
//   * boys_generic    -- any order (generic per-function kernel, golden-vector orders up
//                        to 100): convergent series at the top order + downward recursion,
//                        or erf + upward recursion when T is large against the order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define QBX_BOYS_TMAX 64.0
#define QBX_BOYS_STEP_INV 8.0
#define QBX_BOYS_NCOL 16                       // row i = [F_0(T_i) ... F_15(T_i)], 128 B;
                                               // exp(-T_i) lives in a second array
#define QBX_BOYS_NROW ((int)(QBX_BOYS_TMAX * QBX_BOYS_STEP_INV) + 2)

struct BoysTable {
    const double *f;     // [NROW][16]
    const double *e;     // [NROW]  exp(-T_i)
};

// (2L-1)!! as a compile-time constant
__host__ __device__ constexpr double qbx_dfact(int n) { return n <= 1 ? 1.0 : n * qbx_dfact(n - 2); }

// Shared-memory copy of the columns one class needs: tab[k][row], k = 0..7 -> F_{L+k}(T_row),
// k = 8 -> exp(-T_row).  Column-major so that a warp's 32 row indices spread over the banks.
#define QBX_BOYS_SMEM_COLS 9
#define QBX_BOYS_SMEM_BYTES (QBX_BOYS_SMEM_COLS * QBX_BOYS_NROW * 8)

template <int L>
__device__ __forceinline__ void boys_stage_smem(const BoysTable &tb, double *tab)
{
    for (int e = threadIdx.x; e < 8 * QBX_BOYS_NROW; e += blockDim.x) {
        const int k = e / QBX_BOYS_NROW, row = e - k * QBX_BOYS_NROW;
        tab[e] = __ldg(tb.f + row * QBX_BOYS_NCOL + L + k);
    }
    for (int row = threadIdx.x; row < QBX_BOYS_NROW; row += blockDim.x) tab[8 * QBX_BOYS_NROW + row] = __ldg(tb.e + row);
}

// F[0..L] = F_m(T) * scale, branch-free: below QBX_BOYS_TMAX the top order comes from the
// 8-term Taylor expansion about the nearest table row, above it from the asymptotic form
// F_L = (2L-1)!! sqrt(pi) / (2 (2T)^L sqrt(T)) (exp(-T) < 2e-28 is dropped); both then run the
// same downward recursion F_{m-1} = (2T F_m + exp(-T)) / (2m-1)  (BoysFunction.jl:33-40).
template <int L>
__device__ __forceinline__ void boys_table(const double *tab, double T, double scale, double (&F)[L + 1])
{
    const bool big = T >= QBX_BOYS_TMAX;
    const int i = big ? (QBX_BOYS_NROW - 2) : __double2int_rn(T * QBX_BOYS_STEP_INV);
    const double mx = fma((double)i, 1.0 / QBX_BOYS_STEP_INV, -T);          // -(T - T_i), |mx| <= 1/16
    const double *c = tab + i;
    double r = c[7 * QBX_BOYS_NROW];
    r = fma(r, mx * (1.0 / 7.0), c[6 * QBX_BOYS_NROW]);
    r = fma(r, mx * (1.0 / 6.0), c[5 * QBX_BOYS_NROW]);
    r = fma(r, mx * (1.0 / 5.0), c[4 * QBX_BOYS_NROW]);
    r = fma(r, mx * (1.0 / 4.0), c[3 * QBX_BOYS_NROW]);
    r = fma(r, mx * (1.0 / 3.0), c[2 * QBX_BOYS_NROW]);
    r = fma(r, mx * (1.0 / 2.0), c[1 * QBX_BOYS_NROW]);
    r = fma(r, mx, c[0]);
    const double rt = rsqrt(T);                                             // inf at T = 0, unused there
    if constexpr (L == 0) {
        F[0] = (big ? 0.88622692545275801365 * rt : r) * scale;
    } else {
        double ex = 1.0 / 5040.0;
        ex = fma(ex, mx, 1.0 / 720.0);
        ex = fma(ex, mx, 1.0 / 120.0);
        ex = fma(ex, mx, 1.0 / 24.0);
        ex = fma(ex, mx, 1.0 / 6.0);
        ex = fma(ex, mx, 0.5);
        ex = fma(ex, mx, 1.0);
        ex = fma(ex, mx, 1.0);
        ex *= c[8 * QBX_BOYS_NROW];
        const double h = 0.5 * rt * rt;                                     // 1 / (2T)
        double as = 0.88622692545275801365 * qbx_dfact(2 * L - 1) * rt;
#pragma unroll
        for (int m = 0; m < L; ++m) as *= h;
        F[L] = (big ? as : r) * scale;
        ex = big ? 0.0 : ex * scale;
        const double t2 = 2.0 * T;
#pragma unroll
        for (int m = L; m >= 1; --m) F[m - 1] = fma(t2, F[m], ex) * (1.0 / (2.0 * m - 1.0));
    }
}

// table evaluation straight from global memory (qbx_boys with table = 1: pins the tabulated path)
template <int L>
__device__ __forceinline__ void boys_table_global(const BoysTable &tb, double T, double scale, double (&F)[L + 1])
{
    // gather the 9 values this T needs into a 1-row "table" and reuse the evaluator above
    const bool big = T >= QBX_BOYS_TMAX;
    const int i = big ? (QBX_BOYS_NROW - 2) : __double2int_rn(T * QBX_BOYS_STEP_INV);
    double loc[9];
#pragma unroll
    for (int k = 0; k < 8; ++k) loc[k] = __ldg(tb.f + i * QBX_BOYS_NCOL + L + k);
    loc[8] = __ldg(tb.e + i);
    const double mx = fma((double)i, 1.0 / QBX_BOYS_STEP_INV, -T);
    double r = loc[7];
    r = fma(r, mx * (1.0 / 7.0), loc[6]);
    r = fma(r, mx * (1.0 / 6.0), loc[5]);
    r = fma(r, mx * (1.0 / 5.0), loc[4]);
    r = fma(r, mx * (1.0 / 4.0), loc[3]);
    r = fma(r, mx * (1.0 / 3.0), loc[2]);
    r = fma(r, mx * (1.0 / 2.0), loc[1]);
    r = fma(r, mx, loc[0]);
    const double rt = rsqrt(T);
    double ex = 1.0 / 5040.0;
    ex = fma(ex, mx, 1.0 / 720.0);
    ex = fma(ex, mx, 1.0 / 120.0);
    ex = fma(ex, mx, 1.0 / 24.0);
    ex = fma(ex, mx, 1.0 / 6.0);
    ex = fma(ex, mx, 0.5);
    ex = fma(ex, mx, 1.0);
    ex = fma(ex, mx, 1.0);
    ex *= loc[8];
    const double h = 0.5 * rt * rt;
    double as = 0.88622692545275801365 * qbx_dfact(2 * L - 1) * rt;
#pragma unroll
    for (int m = 0; m < L; ++m) as *= h;
    F[L] = (big ? as : r) * scale;
    ex = big ? 0.0 : ex * scale;
    const double t2 = 2.0 * T;
#pragma unroll
    for (int m = L; m >= 1; --m) F[m - 1] = fma(t2, F[m], ex) * (1.0 / (2.0 * m - 1.0));
}

// F[0..mtop] for any mtop >= 0 (runtime), T >= 0.
__device__ inline void boys_generic(double T, int mtop, double *F)
{
    const double tsw = fmax(30.0, (double)mtop + 20.0);
    if (T < tsw) {
        const double e = exp(-T);
        double term = 1.0 / (2.0 * mtop + 1.0), sum = term;
        for (int k = 1; k < 4000; ++k) {
            term *= 2.0 * T / (2.0 * mtop + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        F[mtop] = e * sum;
        for (int m = mtop; m >= 1; --m) F[m - 1] = (2.0 * T * F[m] + e) / (2.0 * m - 1.0);
    } else {
        const double e = exp(-T), st = sqrt(T);
        F[0] = 0.88622692545275801365 * erf(st) / st;
        for (int m = 0; m < mtop; ++m) F[m + 1] = ((2.0 * m + 1.0) * F[m] - e) / (2.0 * T);
    }
}
