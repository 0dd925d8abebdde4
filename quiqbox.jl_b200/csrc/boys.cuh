// Boys function F_m(T) on the device.
//
// Replaces computeBoysSequence (src/Integration/Engines/BoysFunction.jl:67-77), which the
// reference evaluates per primitive component quartet through SpecialFunctions.gamma_inc.
// Two evaluators:
//   * boys_table<L>   -- hot path of the class kernels (L <= 8): 8-term Taylor expansion
//                        about the nearest point of a 1/8-spaced table (one 128-byte row
//                        = F_0..F_15 at that point, read through the read-only path),
//                        exp(-T) from the same row + a Taylor tail (no SFU/exp call), then
//                        the reference's own downward recursion (BoysFunction.jl:33-40).
//                        T >= QBX_BOYS_TMAX uses the asymptotic series (exp(-T) < 2e-28).
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

// F[0..L] = F_m(T) * scale.  L is a compile-time constant <= 8.
template <int L>
__device__ __forceinline__ void boys_table(const BoysTable &tb, double T, double scale, double (&F)[L + 1])
{
    if (T < QBX_BOYS_TMAX) {
        const int i = __double2int_rn(T * QBX_BOYS_STEP_INV);
        const double x = fma(-(double)i, 1.0 / QBX_BOYS_STEP_INV, T);        // T - T_i, |x| <= 1/16
        const double mx = -x;
        const double *row = tb.f + i * QBX_BOYS_NCOL + L;
        // F_L(T) = sum_k F_{L+k}(T_i) (-x)^k / k!,  k = 0..7  (Horner)
        double r = __ldg(row + 7);
        r = fma(r, mx * (1.0 / 7.0), __ldg(row + 6));
        r = fma(r, mx * (1.0 / 6.0), __ldg(row + 5));
        r = fma(r, mx * (1.0 / 5.0), __ldg(row + 4));
        r = fma(r, mx * (1.0 / 4.0), __ldg(row + 3));
        r = fma(r, mx * (1.0 / 3.0), __ldg(row + 2));
        r = fma(r, mx * (1.0 / 2.0), __ldg(row + 1));
        r = fma(r, mx, __ldg(row));
        if (L > 0) {
            // exp(-T) = exp(-T_i) * exp(-x), 8-term Taylor for the second factor
            double ex = 1.0 / 5040.0;
            ex = fma(ex, mx, 1.0 / 720.0);
            ex = fma(ex, mx, 1.0 / 120.0);
            ex = fma(ex, mx, 1.0 / 24.0);
            ex = fma(ex, mx, 1.0 / 6.0);
            ex = fma(ex, mx, 0.5);
            ex = fma(ex, mx, 1.0);
            ex = fma(ex, mx, 1.0);
            ex *= __ldg(tb.e + i) * scale;
            F[L] = r * scale;
            const double t2 = 2.0 * T;
#pragma unroll
            for (int m = L; m >= 1; --m) F[m - 1] = fma(t2, F[m], ex) * (1.0 / (2.0 * m - 1.0));
        } else {
            F[0] = r * scale;
        }
    } else {
        const double it = 1.0 / T;
        double f = 0.88622692545275801365 * sqrt(it) * scale;       // sqrt(pi)/2 / sqrt(T)
        F[0] = f;
        const double h = 0.5 * it;
#pragma unroll
        for (int m = 0; m < L; ++m) { f *= (2.0 * m + 1.0) * h; F[m + 1] = f; }
    }
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
