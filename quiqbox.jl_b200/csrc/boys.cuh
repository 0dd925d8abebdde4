// Boys function F_m(T) on the device.
//
// Replaces computeBoysSequence (src/Integration/Engines/BoysFunction.jl:67-77), which the
// reference evaluates per primitive component quartet through SpecialFunctions.gamma_inc.
// Two evaluators:
//   * boys_table<L>   -- hot path of the class kernels (L <= 8): degree-5 Taylor expansion
//                        about the nearest point of a 1/32-spaced table; only F_{L+5}(T_i)
//                        and exp(-T_i) are staged in shared memory (once per persistent
//                        block), the other coefficients follow by downward recursion in
//                        registers; exp(-T) from the table + a Taylor tail (no SFU/exp
//                        call), then the reference's own downward recursion
//                        (BoysFunction.jl:33-40).
//                        T >= QBX_BOYS_TMAX switches to the asymptotic form by a select, not
//                        a branch (a warp mixes near and far kets).
//   * boys_generic    -- any order (generic per-function kernel, golden-vector orders up
//                        to 100): convergent series at the top order + downward recursion,
//                        or erf + upward recursion when T is large against the order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define QBX_BOYS_TMAX 64.0
#define QBX_BOYS_STEP_INV 32.0
#define QBX_BOYS_DEG 5                         // Taylor degree about the nearest table point
#define QBX_BOYS_NCOL 16                       // row i = [F_0(T_i) ... F_15(T_i)], 128 B;
                                               // exp(-T_i) lives in a second array
#define QBX_BOYS_NROW ((int)(QBX_BOYS_TMAX * QBX_BOYS_STEP_INV) + 2)

struct BoysTable {
    const double *f;     // [NROW][16]
    const double *e;     // [NROW]  exp(-T_i)
};

// (2L-1)!! as a compile-time constant
__host__ __device__ constexpr double qbx_dfact(int n) { return n <= 1 ? 1.0 : n * qbx_dfact(n - 2); }

// Shared-memory copy of what one class needs: tab[row] = F_{L+DEG}(T_row), tab[NROW + row] =
// exp(-T_row).  Two columns only: the lower Taylor coefficients F_{L+k}(T_row), k < DEG, are NOT
// tabulated but recomputed in registers from those two numbers by the (stable) downward recursion
// at the table point.  The first version staged all eight coefficient columns and did nine
// shared-memory loads with random row indices per primitive quartet; ncu showed the s/p kernels
// L1TEX-bound on exactly those loads and their bank conflicts (profiles/r01/ncu_final_summary.md:
// L1TEX 64-87 % against 43-63 % FP64 pipe).  Now: 2 loads, and with the 1/32 grid a degree-5
// expansion, i.e. about the same number of FP64 operations as before.
#define QBX_BOYS_SMEM_COLS 2
#define QBX_BOYS_SMEM_BYTES (QBX_BOYS_SMEM_COLS * QBX_BOYS_NROW * 8)

template <int L>
__device__ __forceinline__ void boys_stage_smem(const BoysTable &tb, double *tab)
{
    for (int row = threadIdx.x; row < QBX_BOYS_NROW; row += blockDim.x) {
        tab[row] = __ldg(tb.f + row * QBX_BOYS_NCOL + L + QBX_BOYS_DEG);
        tab[QBX_BOYS_NROW + row] = __ldg(tb.e + row);
    }
}

// F[0..L] = F_m(T) * scale from ctop = F_{L+DEG}(T_i) and ei = exp(-T_i), i = nearest table row.
// Branch-free: below QBX_BOYS_TMAX the top order F_L comes from the Taylor expansion about T_i,
//     F_L(T) = sum_k F_{L+k}(T_i) (T_i - T)^k / k!,   |T_i - T| <= 1/64, k <= DEG = 5
// (remainder < 1/64^6 / 6! = 2e-14 times F_{L+6} < 0.1), above it from the asymptotic form
// F_L = (2L-1)!! sqrt(pi) / (2 (2T)^L sqrt(T)) (exp(-T) < 2e-28 is dropped); both then run the
// same downward recursion F_{m-1} = (2T F_m + exp(-T)) / (2m-1)  (BoysFunction.jl:33-40), which
// is also what produces the coefficients F_{L+k}(T_i) here.
template <int L>
__device__ __forceinline__ void boys_eval(double ctop, double ei, int i, bool big, double T, double scale, double (&F)[L + 1])
{
    const double Ti = (double)i * (1.0 / QBX_BOYS_STEP_INV);                 // exact
    const double mx = Ti - T;                                               // -(T - T_i), |mx| <= 1/64
    const double t2i = 2.0 * Ti;
    double c[QBX_BOYS_DEG + 1];
    c[QBX_BOYS_DEG] = ctop;
#pragma unroll
    for (int k = QBX_BOYS_DEG - 1; k >= 0; --k) c[k] = fma(t2i, c[k + 1], ei) * (1.0 / (2.0 * (L + k) + 1.0));
    double r = ctop;
#pragma unroll
    for (int k = QBX_BOYS_DEG - 1; k >= 0; --k) r = fma(r, mx * (1.0 / (k + 1.0)), c[k]);
    const double rt = rsqrt(T);                                             // inf at T = 0, unused there
    if constexpr (L == 0) {
        F[0] = (big ? 0.88622692545275801365 * rt : r) * scale;
    } else {
        double ex = 1.0 / 720.0;                                            // exp(mx), |mx| <= 1/64: 1e-16
        ex = fma(ex, mx, 1.0 / 120.0);
        ex = fma(ex, mx, 1.0 / 24.0);
        ex = fma(ex, mx, 1.0 / 6.0);
        ex = fma(ex, mx, 0.5);
        ex = fma(ex, mx, 1.0);
        ex = fma(ex, mx, 1.0);
        ex *= ei;
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

template <int L>
__device__ __forceinline__ void boys_table(const double *tab, double T, double scale, double (&F)[L + 1])
{
    const bool big = T >= QBX_BOYS_TMAX;
    const int i = big ? (QBX_BOYS_NROW - 2) : __double2int_rn(T * QBX_BOYS_STEP_INV);
    boys_eval<L>(tab[i], tab[QBX_BOYS_NROW + i], i, big, T, scale, F);
}

// table evaluation straight from global memory (qbx_boys with table = 1: pins the tabulated path)
template <int L>
__device__ __forceinline__ void boys_table_global(const BoysTable &tb, double T, double scale, double (&F)[L + 1])
{
    const bool big = T >= QBX_BOYS_TMAX;
    const int i = big ? (QBX_BOYS_NROW - 2) : __double2int_rn(T * QBX_BOYS_STEP_INV);
    boys_eval<L>(__ldg(tb.f + i * QBX_BOYS_NCOL + L + QBX_BOYS_DEG), __ldg(tb.e + i), i, big, T, scale, F);
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
