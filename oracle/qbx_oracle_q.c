/*
 * qbx_oracle_q.c -- HIGHER-PRECISION ARBITER of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * The reference's primitive ERI routine (computePGTOrbTwoBodyRepulsion!, GaussianOrbitals.jl:594-663) is
 * numerically orientation dependent in Float64: for (s s|d d)-type entries with a tight primitive the value
 * obtained in the reference's index order differs from the value of the permuted (l-canonical) call by up to
 * 2.8e-5 on (H2O)2/cc-pVDZ.  To say which one is RIGHT, the same recurrences
 * (qbx_prim_eri.inc, shared text with a double instantiation that is checked bit for bit against
 * orc_prim_eri) are evaluated here in __float128 (113-bit significand: a cancellation of 11 digits still
 * leaves 23), with a quad-precision Boys function of its own.  In quad precision every orientation agrees to
 * ~1e-30, so the result is the exact integral of the given double-precision primitives, rounded once.
 *
 * Entry points: orcq_prim_eri (quad, rounded to double), orcd_prim_eri (double instantiation of the shared
 * text), orcq_eri_list (contracted integrals over the boundary's basis format, Framework.jl:526-554).
 * Needs libquadmath (gcc).
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q128;

/* ---- Boys function in quad precision.  F_n(x) = exp(-x) sum_k (2x)^k / ((2n+1)(2n+3)...(2n+2k+1)) (all terms
 * positive, converges for every x; used at the top order for x < 200), then the stable downward recursion
 * F_{m-1} = (2x F_m + exp(-x))/(2m-1) (BoysFunction.jl:33-40).  For x >= 200, exp(-x) < 1e-86 of F and
 * F_m = (2m-1)!!/(2x)^m sqrt(pi/x)/2 upward from F_0 = sqrt(pi/x) erf(sqrt x)/2 is exact to quad rounding. */
static void orcq_boys_sequence(q128 x, int n, q128 *out)
{
    if (x < 200.0Q) {
        q128 term = 1.0Q / (2 * n + 1), sum = term;
        for (int k = 1; k < 4000; ++k) {
            term *= 2.0Q * x / (2 * n + 2 * k + 1);
            sum += term;
            if (term < sum * 1e-36Q) break;
        }
        const q128 ex = expq(-x);
        out[n] = ex * sum;
        for (int m = n; m >= 1; --m) out[m - 1] = (2.0Q * x * out[m] + ex) / (2 * m - 1);
    } else {
        const q128 ex = expq(-x);
        out[0] = sqrtq(M_PIq / x) * erfq(sqrtq(x)) / 2.0Q;
        for (int m = 0; m < n; ++m) out[m + 1] = ((2 * m + 1) * out[m] - ex) / (2.0Q * x);
    }
}

#define ORQ_MAXL 96
#define ORQ_REAL q128
#define ORQ_FN(name) orcq_##name
#define ORQ_EXP(x) expq(x)
#define ORQ_SQRT(x) sqrtq(x)
#define ORQ_PREFAC (2.0Q * powq(M_PIq, 2.5Q))
#define ORQ_BOYS_SEQ(x, n, out) orcq_boys_sequence((x), (n), (out))
#include "qbx_prim_eri.inc"
#undef ORQ_REAL
#undef ORQ_FN
#undef ORQ_EXP
#undef ORQ_SQRT
#undef ORQ_PREFAC
#undef ORQ_BOYS_SEQ

/* double instantiation of the same text, on the double oracle's own Boys function: must reproduce
 * orc_prim_eri bit for bit (tests/test_oracle_arbiter.py) -- which proves the shared text IS the oracle's
 * algorithm, so that the quad instantiation differs from it in precision only */
void orc_boys_sequence(double x, int n, double *out);      /* qbx_oracle.c (linked into the same library) */
#define ORQ_REAL double
#define ORQ_FN(name) orcd_##name
#define ORQ_EXP(x) exp(x)
#define ORQ_SQRT(x) sqrt(x)
#define ORQ_PREFAC (2.0 * (double)powl(3.14159265358979323846264338327950288L, 2.5L))
#define ORQ_BOYS_SEQ(x, n, out) orc_boys_sequence((x), (n), (out))
#include "qbx_prim_eri.inc"

double orcq_prim_eri_d(const double *cen, const double *xpn, const int *ang) { return (double)orcq_prim_eri(cen, xpn, ang); }
double orcd_prim_eri_d(const double *cen, const double *xpn, const int *ang) { return orcd_prim_eri(cen, xpn, ang); }
double orcq_boys_d(double x, int n)
{
    q128 f[ORQ_MAXL + 1];
    orcq_boys_sequence((q128)x, n, f);
    return (double)f[n];
}

typedef struct {
    int64_t nprim, nbf;
    const double *cen, *xpn;
    const int32_t *ang;
    const int64_t *bf_off, *bf_prim;
    const double *bf_w;
} orc_basis;

/* getOrbLayoutIntegralCore! two-body (Framework.jl:526-554), accumulated in quad */
void orcq_eri_list(const orc_basis *b, int64_t n, const int64_t *ijkl, double *out)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t t = 0; t < n; ++t) {
        const int64_t i = ijkl[4 * t], j = ijkl[4 * t + 1], k = ijkl[4 * t + 2], l = ijkl[4 * t + 3];
        q128 res = 0;
        double cen[12], xpn[4];
        int ang[12];
        for (int64_t s = b->bf_off[l]; s < b->bf_off[l + 1]; ++s)
            for (int64_t r = b->bf_off[k]; r < b->bf_off[k + 1]; ++r)
                for (int64_t q = b->bf_off[j]; q < b->bf_off[j + 1]; ++q)
                    for (int64_t p = b->bf_off[i]; p < b->bf_off[i + 1]; ++p) {
                        const int64_t id[4] = {b->bf_prim[p], b->bf_prim[q], b->bf_prim[r], b->bf_prim[s]};
                        for (int u = 0; u < 4; ++u) {
                            memcpy(cen + 3 * u, b->cen + 3 * id[u], 3 * sizeof(double));
                            xpn[u] = b->xpn[id[u]];
                            for (int d = 0; d < 3; ++d) ang[3 * u + d] = b->ang[3 * id[u] + d];
                        }
                        res += orcq_prim_eri(cen, xpn, ang) * ((q128)b->bf_w[p] * b->bf_w[q] * b->bf_w[r] * b->bf_w[s]);
                    }
        out[t] = (double)res;
    }
}
