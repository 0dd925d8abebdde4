// Warp-cooperative, table-driven ERI kernel for the LARGE shell-quartet classes.
//
// The thread-per-quartet kernels of eri_class.cuh keep every intermediate of a quartet in one
// thread; from (dp|dp) upwards that is 250-1000 accumulators plus up to 2500 recurrence
// intermediates, which spill to local memory and run at < 1 TFLOP/s.  Here one WARP owns one
// contracted shell quartet, all intermediates live in shared memory, and the lanes share the
// entries of every recurrence level (entries of a level are independent):
//
//   per primitive quartet   Boys (every lane, redundantly) -> S[m]
//                           VRR level e -> e+1          (cf. vertTransfer,  GaussianOrbitals.jl:386)
//                           transfer level f -> f+1     (cf. modeTransfer,  :529)
//                           ACC += needed [e0|f0]
//   per contracted quartet  HRR bra, HRR ket            (cf. horiTransfer,  :394), weights, store
//
// The recurrences are not unrolled: the host emits, per class, a "program" -- for every entry
// the shared-memory addresses of its sources and its integer coefficients -- that the lanes
// read coalesced.  One kernel therefore serves every class (and compiles in seconds); the
// arithmetic is the same sequence of operations as in eri_class.cuh.
#include <algorithm>
#include <cstdio>
#include <map>
#include <vector>

#include "engine.h"

namespace {

struct Op {                     // 16 bytes, read coalesced (one 128-bit load per lane)
    short dst, a, b, c, d;      // addresses in the warp's buffer (< 32768 doubles); -1 = absent
    short ax, n1, n2;           // axis, integer coefficients / vector selector
};
static_assert(sizeof(Op) == 16, "Op layout");

struct Level { int first, count; };

struct Program {
    std::vector<Op> ops;
    std::vector<Level> vrr, xfer, hrr;      // level boundaries into ops
    Level acc{0, 0}, store{0, 0};
    int L = 0, buf_doubles = 0, acc_off = 0, acc_n = 0, ncomp = 0, y_off = 0;
    int NA = 1, NB = 1, NCc = 1, ND = 1;
    // device copies
    Op *d_ops = nullptr;
};

int cidx(int j, int k) { return (j + k) * (j + k + 1) / 2 + k; }
int s1(int n) { return n * (n + 1) * (n + 2) / 6; }
int ncsum(int lo, int hi) { return s1(hi + 1) - s1(lo); }

// Builds the program for class (la lb|lc ld); mirrors the loops of EriClass<...>::primitive/finish.
Program *build_program(int LA, int LB, int LC, int LD)
{
    Program *P = new Program;
    const int E = LA + LB, F = LC + LD, L = E + F;
    P->L = L;
    P->NA = NC(LA); P->NB = NC(LB); P->NCc = NC(LC); P->ND = NC(LD);
    P->ncomp = P->NA * P->NB * P->NCc * P->ND;
    auto voff = [&](int e) { return (L + 1) * s1(e) - (e - 1) * e * (e + 1) * (e + 2) / 8; };
    const int vtot = voff(L + 1);
    auto elo = [&](int f) { return std::max(0, LA - (F - f)); };
    auto ehi = [&](int f) { return E + (F - f); };
    std::vector<std::vector<int>> woff(F + 2, std::vector<int>(L + 3, -1));
    int wtot = 0;
    for (int f = 1; f <= F; ++f)
        for (int e = elo(f); e <= ehi(f); ++e) { woff[f][e] = vtot + wtot; wtot += NC(f) * NC(e); }
    auto saddr = [&](int f, int e, int ce, int cf) { return f == 0 ? voff(e) + ce : woff[f][e] + cf * NC(e) + ce; };
    const int s_total = vtot + wtot;

    // ---- VRR: level e -> e + 1, all orders m <= L - e - 1
    for (int e = 0; e < L; ++e) {
        Level lv{(int)P->ops.size(), 0};
        for (int m = 0; m <= L - e - 1; ++m)
            for (int r = 0; r <= e + 1; ++r)
                for (int k = 0; k <= r; ++k) {
                    const int i = e + 1 - r, j = r - k;
                    const int ax = i > 0 ? 0 : (j > 0 ? 1 : 2);
                    const int j1 = j - (ax == 1), k1 = k - (ax == 2), i1 = i - (ax == 0);
                    const int n1 = ax == 0 ? i1 : (ax == 1 ? j1 : k1);
                    Op o{};
                    o.dst = voff(e + 1) + m * NC(e + 1) + cidx(j, k);
                    o.a = voff(e) + m * NC(e) + cidx(j1, k1);
                    o.b = voff(e) + (m + 1) * NC(e) + cidx(j1, k1);
                    o.c = o.d = -1;
                    if (n1 > 0) {
                        const int c2 = cidx(j1 - (ax == 1), k1 - (ax == 2));
                        o.c = voff(e - 1) + m * NC(e - 1) + c2;
                        o.d = voff(e - 1) + (m + 1) * NC(e - 1) + c2;
                    }
                    o.ax = (short)ax; o.n1 = (short)n1;
                    P->ops.push_back(o); ++lv.count;
                }
        P->vrr.push_back(lv);
    }
    // ---- electron transfer: level f -> f + 1 at m = 0
    for (int f = 0; f < F; ++f) {
        Level lv{(int)P->ops.size(), 0};
        for (int rf = 0; rf <= f + 1; ++rf)
            for (int kf = 0; kf <= rf; ++kf) {
                const int fi = f + 1 - rf, fj = rf - kf;
                const int ax = fi > 0 ? 0 : (fj > 0 ? 1 : 2);
                const int fj1 = fj - (ax == 1), fk1 = kf - (ax == 2);
                const int nf = (ax == 0 ? fi : (ax == 1 ? fj : kf)) - 1;
                const int cf = cidx(fj, kf), cf1 = cidx(fj1, fk1);
                for (int e = elo(f + 1); e <= ehi(f + 1); ++e)
                    for (int re = 0; re <= e; ++re)
                        for (int ke = 0; ke <= re; ++ke) {
                            const int ei = e - re, ej = re - ke, ce = cidx(ej, ke);
                            const int na = ax == 0 ? ei : (ax == 1 ? ej : ke);
                            Op o{};
                            o.dst = saddr(f + 1, e, ce, cf);
                            o.a = saddr(f, e, ce, cf1);
                            o.b = saddr(f, e + 1, cidx(ej + (ax == 1), ke + (ax == 2)), cf1);
                            o.c = na > 0 ? saddr(f, e - 1, cidx(ej - (ax == 1), ke - (ax == 2)), cf1) : -1;
                            o.d = nf > 0 ? saddr(f - 1, e, ce, cidx(fj1 - (ax == 1), fk1 - (ax == 2))) : -1;
                            o.ax = (short)ax; o.n1 = (short)na; o.n2 = (short)nf;
                            P->ops.push_back(o); ++lv.count;
                        }
            }
        P->xfer.push_back(lv);
    }
    // ---- accumulate [e0|f0], e in [LA,E], f in [LC,F]
    const int nae = ncsum(LA, E), nkf = ncsum(LC, F);
    P->acc_off = s_total;
    P->acc_n = nae * nkf;
    auto aoff = [&](int f, int e) { return P->acc_off + (s1(f) - s1(LC)) * nae + NC(f) * (s1(e) - s1(LA)); };
    P->acc.first = (int)P->ops.size();
    for (int f = LC; f <= F; ++f)
        for (int e = LA; e <= E; ++e)
            for (int cf = 0; cf < NC(f); ++cf)
                for (int ce = 0; ce < NC(e); ++ce) {
                    Op o{};
                    o.dst = aoff(f, e) + cf * NC(e) + ce;
                    o.a = saddr(f, e, ce, cf);
                    o.b = o.c = o.d = -1;
                    P->ops.push_back(o); ++P->acc.count;
                }
    // ---- HRR: temporaries re-use the S region (dead after the primitive loop); X after ACC
    const int x_off = P->acc_off + P->acc_n;              // X[kk][a*NB + b]
    const int NA = P->NA, NB = P->NB, NCc = P->NCc, ND = P->ND;
    const int y_off = x_off + nkf * NA * NB;
    // temporaries of the two d-type HRR passes: inside the (dead) S region when they fit
    const int need_bra = LB == 2 ? nkf * (NC(LA) + NC(LA + 1)) * 3 : 0;
    const int need_ket = LD == 2 ? NA * NB * (NC(LC) + NC(LC + 1)) * 3 : 0;
    const int tneed = std::max(need_bra, need_ket);
    const int tbase = tneed <= s_total ? 0 : y_off + P->ncomp;
    int tmp = tbase;                                      // bump allocator
    auto hrr_pass = [&](int LX, int LY, int nouter, auto in_addr, auto out_addr, int vec, std::vector<Level> &levels) {
        // emits up to LY levels; `vec` selects AB (0) or CD (1), encoded in n2
        std::vector<std::vector<int>> t1(nouter);         // address of t1[o][d][c][x]
        if (LY == 0) {
            Level lv{(int)P->ops.size(), 0};
            for (int o = 0; o < nouter; ++o)
                for (int c = 0; c < NC(LX); ++c) {
                    Op op{}; op.dst = out_addr(o, c, 0); op.a = in_addr(o, LX, c); op.b = op.c = op.d = -1; op.ax = -1; op.n2 = (short)vec;
                    P->ops.push_back(op); ++lv.count;
                }
            levels.push_back(lv);
            return;
        }
        Level l1{(int)P->ops.size(), 0};
        for (int o = 0; o < nouter; ++o) {
            t1[o].assign(LY * NC(LX + LY - 1) * 3, -1);
            for (int d = 0; d < LY; ++d)
                for (int r = 0; r <= LX + d; ++r)
                    for (int k = 0; k <= r; ++k) {
                        const int j = r - k, c = cidx(j, k);
                        for (int x = 0; x < 3; ++x) {
                            Op op{};
                            const bool final1 = (LY == 1);
                            op.dst = final1 ? out_addr(o, c, x) : (t1[o][(d * NC(LX + LY - 1) + c) * 3 + x] = tmp++);
                            op.a = in_addr(o, LX + d + 1, cidx(j + (x == 1), k + (x == 2)));     // hi
                            op.b = in_addr(o, LX + d, c);                                       // lo
                            op.c = op.d = -1; op.ax = (short)x; op.n2 = (short)vec;
                            P->ops.push_back(op); ++l1.count;
                        }
                    }
        }
        levels.push_back(l1);
        if (LY == 2) {
            Level l2{(int)P->ops.size(), 0};
            static const int dax[6] = {0, 0, 0, 1, 1, 2}, drem[6] = {0, 1, 2, 1, 2, 2};   // xx xy xz yy yz zz
            for (int o = 0; o < nouter; ++o)
                for (int r = 0; r <= LX; ++r)
                    for (int k = 0; k <= r; ++k) {
                        const int j = r - k, c = cidx(j, k);
                        for (int b = 0; b < 6; ++b) {
                            const int ax = dax[b], rem = drem[b];
                            Op op{};
                            op.dst = out_addr(o, c, b);
                            op.a = t1[o][(1 * NC(LX + 1) + cidx(j + (ax == 1), k + (ax == 2))) * 3 + rem];
                            op.b = t1[o][(0 * NC(LX + 1) + c) * 3 + rem];
                            op.c = op.d = -1; op.ax = (short)ax; op.n2 = (short)vec;
                            P->ops.push_back(op); ++l2.count;
                        }
                    }
            levels.push_back(l2);
        }
    };
    // bra: outer index = stacked ket component kk; input ACC[(f,cf)][e][ce]
    std::vector<int> kk_f(nkf), kk_cf(nkf);
    { int kk = 0; for (int f = LC; f <= F; ++f) for (int cf = 0; cf < NC(f); ++cf) { kk_f[kk] = f; kk_cf[kk] = cf; ++kk; } }
    hrr_pass(LA, LB, nkf,
             [&](int kk, int e, int c) { return aoff(kk_f[kk], e) + kk_cf[kk] * NC(e) + c; },
             [&](int kk, int a, int b) { return x_off + kk * NA * NB + a * NB + b; }, 0, P->hrr);
    // ket: outer index = ab; input X[(f,c)][ab]; output goes to the Y region (then stored)
    tmp = tbase;                                          // bra temporaries are dead once X is complete
    hrr_pass(LC, LD, NA * NB,
             [&](int ab, int f, int c) { return x_off + (ncsum(LC, f - 1) + c) * NA * NB + ab; },
             [&](int ab, int c, int d) { return y_off + (ab * NCc + c) * ND + d; }, 1, P->hrr);
    P->y_off = y_off;
    P->buf_doubles = y_off + P->ncomp + (tbase ? tneed : 0);
    if (P->buf_doubles >= 32768) { delete P; return nullptr; }
    return P;
}

struct CoopArgs {
    PairSet bra, ket;
    const int2 *tasks;
    int64_t ntasks;
    double *out;
    const double *shell_scale;
    BoysTable boys;
    const Op *ops;
    int L, buf, acc_off, acc_n, y_off, ncomp, NA, NB, NCc, ND;
    int nvrr, nxfer, nhrr;
    Level vrr[9], xfer[5], hrr[4], acc;
    double boys_inv[QBX_BOYS_DEG];     // 1 / (2 (L + k) + 1)
};

// Boys function with a run-time order: same evaluator as boys_eval<L> (boys.cuh) -- two table
// loads, the Taylor coefficients by downward recursion at the table point -- with the
// reciprocals 1/(2(L+k)+1) passed in (kernel parameters, constant bank).
__device__ __forceinline__ void boys_rt(const BoysTable &tb, const double *inv_odd, double T, double scale, int L, double *F)
{
    const bool big = T >= QBX_BOYS_TMAX;
    const int i = big ? (QBX_BOYS_NROW - 2) : __double2int_rn(T * QBX_BOYS_STEP_INV);
    const double ctop = __ldg(tb.f + i * QBX_BOYS_NCOL + L + QBX_BOYS_DEG), ei = __ldg(tb.e + i);
    const double Ti = (double)i * (1.0 / QBX_BOYS_STEP_INV), mx = Ti - T, t2i = 2.0 * Ti;
    double c[QBX_BOYS_DEG + 1];
    c[QBX_BOYS_DEG] = ctop;
#pragma unroll
    for (int k = QBX_BOYS_DEG - 1; k >= 0; --k) c[k] = fma(t2i, c[k + 1], ei) * inv_odd[k];
    double r = ctop;
#pragma unroll
    for (int k = QBX_BOYS_DEG - 1; k >= 0; --k) r = fma(r, mx * (1.0 / (k + 1.0)), c[k]);
    double ex = 1.0 / 720.0;
    ex = fma(ex, mx, 1.0 / 120.0);
    ex = fma(ex, mx, 1.0 / 24.0);
    ex = fma(ex, mx, 1.0 / 6.0);
    ex = fma(ex, mx, 0.5);
    ex = fma(ex, mx, 1.0);
    ex = fma(ex, mx, 1.0);
    ex *= ei;
    const double rt = rsqrt(T), h = 0.5 * rt * rt;
    double as = 0.88622692545275801365 * rt;
    for (int m = 0; m < L; ++m) as *= (2.0 * m + 1.0) * h;
    double f = (big ? as : r) * scale;
    ex = big ? 0.0 : ex * scale;
    F[L] = f;
    const double t2 = 2.0 * T;
    for (int m = L; m >= 1; --m) { f = fma(t2, f, ex) * (1.0 / (2.0 * m - 1.0)); F[m - 1] = f; }
}

#ifndef COOP_WARPS
#define COOP_WARPS 4
#endif
#ifndef QBX_COOP_ONE_PASS
#define QBX_COOP_ONE_PASS 1
#endif
#ifndef QBX_COOP_MINB
#define QBX_COOP_MINB 3                 // resident blocks of 128 threads the register allocation aims at: 168 registers.  A/B on a B200
                                        // (tools/gpu_ab_coop.sh, profiles/r02/ab_coop_launch_bounds.log): five d-rich classes 2.92 ms unbounded,
                                        // 2.41 ms at 3 blocks, 2.47 ms at 4 (128 registers, more spills)
#endif
#ifndef COOP_BATCH
#define COOP_BATCH 4
#endif

// op record i of a level, or a harmless self-referencing dummy past the end
__device__ __forceinline__ Op ld_op(const Op *ops, int i, int n)
{
    Op o;
    if (i < n) {
        const int4 r = __ldg(reinterpret_cast<const int4 *>(ops + i));
        o = *reinterpret_cast<const Op *>(&r);
    } else {
        o.dst = 0; o.a = 0; o.b = 0; o.c = -1; o.d = -1; o.ax = 0; o.n1 = 0; o.n2 = 0;
    }
    return o;
}

// Vertical recurrence of one primitive quartet, table-driven (shared by both cooperative kernels):
// B[0..L] holds the scaled Boys values on entry, the V region [e0|00]^(m) on exit.
__device__ __forceinline__ void coop_vrr(const CoopArgs &p, double *B, int lane, const double (&PA)[3], const double (&WP)[3],
                                         double i2z, double rz, int nlev = 99)
{
    // Entries of one level are independent, so each lane takes COOP_BATCH of them at a
    // time: all op records first (global, L1/L2), then all sources (shared), then the
    // arithmetic and the stores -- otherwise every entry pays the full load latency.
    for (int lv = 0; lv < p.nvrr && lv < nlev; ++lv) {
        const Op *ops = p.ops + p.vrr[lv].first;
        const int n = p.vrr[lv].count;
        for (int i0 = lane; i0 < n; i0 += 32 * COOP_BATCH) {
            Op o[COOP_BATCH]; double sa[COOP_BATCH], sb[COOP_BATCH], sc[COOP_BATCH], sd[COOP_BATCH];
#pragma unroll
            for (int u = 0; u < COOP_BATCH; ++u) o[u] = ld_op(ops, i0 + 32 * u, n);
#pragma unroll
            for (int u = 0; u < COOP_BATCH; ++u) {
                sa[u] = B[o[u].a]; sb[u] = B[o[u].b];
                sc[u] = o[u].c >= 0 ? B[o[u].c] : 0.0; sd[u] = o[u].c >= 0 ? B[o[u].d] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < COOP_BATCH; ++u)
                if (i0 + 32 * u < n)
                    B[o[u].dst] = fma(o[u].n1 * i2z, fma(-rz, sd[u], sc[u]), fma(PA[o[u].ax], sa[u], WP[o[u].ax] * sb[u]));
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Second-generation cooperative kernel (compiled per class): only the vertical recurrence is still
// table-driven.  ncu on the interpreter above (profiles/r01): issue slots 61 % busy with the FP64
// pipe at 12 % -- ~40 instructions per recurrence entry, most of them fetching and decoding the op
// record.  Here the electron transfer, the contraction and both horizontal recurrences are
// compile-time code in which a LANE owns a stacked bra component g = (e, i, j, k) and the ket
// components are unrolled:
//   * transfer  [g|cf] <- k0[ax] [g|cf1] - (zeta/eta) [g+1_ax|cf1] + n_ax/(2 eta) [g-1_ax|cf1]
//                        + nf/(2 eta) [g|cf2]
//     (cf. modeTransfer, GaussianOrbitals.jl:529-538; same recurrence as EriClass::primitive):
//     ax, cf1, cf2, nf depend on the ket component only and fold to constants, the neighbours
//     g+-1_ax are per-lane shared-memory addresses computed once per kernel, so an entry costs
//     4 LDS + 1 STS + the FP64 work and nothing else;
//   * the contraction accumulates in registers: the lane that produces [g|cf] adds it to its own
//     acc[cf] when g is in the contracted range (<= 32 stacked components for every d-rich class);
//   * horizontal recurrences run on register rows with the thread kernels' hrr_apply: bra with one
//     lane per stacked ket component, ket with one lane per (a,b) pair, transposed through shared memory.
template <int LA, int LB, int LC, int LD>
struct Coop2 {
    static constexpr int E = LA + LB, F = LC + LD, L = E + F;
    static constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD), NAB = NA * NB;
    static constexpr int NACC = NCSUM(LA, E), NKET = NCSUM(LC, F);
    static constexpr int GT = S1(E + 1);                       // stacked bra components that take part: degrees 0..E
    static constexpr int GB = S1(LA);                          // first contracted stacked component
    static_assert(F >= 1 && NACC <= 32 && NKET <= 32, "Coop2: class outside the layout's assumptions");
    // ket level f keeps degrees e in [elo(f), E] = global stacked range [glo(f), ghi(f)), at orders m = 0..F-f
    static __host__ __device__ constexpr int elo(int f) { return (LA - (F - f)) > 0 ? (LA - (F - f)) : 0; }
    static __host__ __device__ constexpr int ehi(int) { return E; }
    static __host__ __device__ constexpr int nm(int f) { return F - f + 1; }
    static __host__ __device__ constexpr int glo(int f) { return S1(elo(f)); }
    static __host__ __device__ constexpr int ghi(int f) { return S1(ehi(f) + 1); }
    static __host__ __device__ constexpr int rl(int f) { return ghi(f) - glo(f); }
    static constexpr int VT = VLay<L>::total;                  // V region: [e0|00]^(m), layout of VLay<L>
    static __host__ __device__ constexpr int woff(int f)       // level f: [m][cf][g - glo(f)]; level 0 = V, orders 0..F, stacked
    {
        int o = VT;
        for (int x = 0; x < f; ++x) o += nm(x) * NC(x) * rl(x);
        return o;
    }
    static __host__ __device__ constexpr int row(int f, int m, int cf) { return woff(f) + (m * NC(f) + cf) * rl(f) - glo(f); }
    static constexpr int WT = woff(F + 1);
    // after the primitive loops the W levels are dead: ACC2[kk][ASTR] and X2[ab][XSTR] (odd strides) go there
    static constexpr int ASTR = NACC | 1, XSTR = NKET | 1;
    static constexpr int acc2_off = VT, x2_off = VT + NKET * ASTR;
    static constexpr int BUF = (WT > x2_off + NAB * XSTR ? WT : x2_off + NAB * XSTR);
    // Lane layout: the GT stacked bra components are cut into passes of 32 lanes FROM THE TOP: the high passes
    // q < NPH own g = P0 + 32 q + lane, a last, short pass owns the P0 = GT mod 32 lowest components (none if GT is
    // a multiple of 32).  (dp|dp): GT = 20 -> one pass, every recurrence entry is one instruction stream with 20 lanes
    // (it used to be 16 + 4 lanes in two: 1.21 -> 0.94 ms on the B200).  (dd|..): GT = 35 -> 32 + 3 lanes, and the short
    // pass only holds degrees 0 and 1, which the upper ket levels do not need: it drops out of them at compile time
    // (the earlier split at the first contracted component, 25 + 10 lanes, kept both passes in every level; measured:
    // the same 0.52 / 0.40 / 0.23 ms for (dd|pp), (dd|ds), (dd|dp) either way -- profiles/r02/ab_coop_launch_bounds.log).
    // The contracted components g >= GB accumulate in the lane that owns them (always pass 0: GT - GB <= 32).
    static constexpr int P0 = QBX_COOP_ONE_PASS ? (GT > 32 ? GT % 32 : 0) : GB;     // QBX_COOP_ONE_PASS=0: the old layout, split at GB
    static constexpr int NPH = (GT - P0 + 31) / 32;
    static constexpr int NP = NPH + (P0 > 0 ? 1 : 0);
    static_assert(NPH == 1, "Coop2: the contracted components must sit in pass 0");
    static __host__ __device__ constexpr int pass_lo(int q) { return q < NPH ? P0 + 32 * q : 0; }
    static __host__ __device__ constexpr int pass_hi(int q) { return q < NPH ? (P0 + 32 * q + 32 < GT ? P0 + 32 * q + 32 : GT) : P0; }
    // does pass q hold targets of transfer level f?  (decided at compile time: whole passes drop out of a level)
    static __host__ __device__ constexpr bool pass_in_level(int q, int f) { return pass_lo(q) < ghi(f) && pass_hi(q) > glo(f); }

    struct LaneInfo {            // per pass
        int g;                   // own stacked index (or -1: lane idle in this pass)
        const double *v0;        // [g|0]^(0) in the V region, order m at v0[m * vs]
        int vs;                  // NC(degree of g)
        double *gp;              // B + g and B + (g - 1_ax): a level's row start (a constant) is added at the point of
        const double *dn[3];     //   use, so every access is one LDS/STS with an immediate offset
        double n[3];             // (i, j, k) of g
    };

    static __device__ __forceinline__ void lane_info(int q, int lane, double *B, LaneInfo &I)
    {
        const int g = pass_lo(q) + lane;
        I.g = g < pass_hi(q) ? g : -1;
        int e = 0;
        while (e < L && S1(e + 1) <= g) ++e;
        const int c = g - S1(e);
        int r = 0;
        while ((r + 1) * (r + 2) / 2 <= c) ++r;
        const int k = c - r * (r + 1) / 2, j = r - k, i = e - r;
        I.v0 = B + VLay<L>::off(e) + c;
        I.vs = NC(e);
        I.gp = B + g;
        I.dn[0] = B + (i > 0 ? S1(e - 1) + CIDX(j, k) : g);
        I.dn[1] = B + (j > 0 ? S1(e - 1) + CIDX(j - 1, k) : g);
        I.dn[2] = B + (k > 0 ? S1(e - 1) + CIDX(j, k - 1) : g);
        I.n[0] = i; I.n[1] = j; I.n[2] = k;
        if (I.g < 0) { I.v0 = B; I.vs = 0; I.gp = B; for (int x = 0; x < 3; ++x) { I.dn[x] = B; I.n[x] = 0.0; } }
    }

    // Vertical recurrence on the KET (Obara-Saika, centre C), level FL -> FL + 1 at orders m = 0..F-FL-1:
    //   [g|cf]^(m) = QC_ax [g|cf1]^(m) + WQ_ax [g|cf1]^(m+1) + nf/(2 eta) ([g|cf2]^(m) - (rho/eta) [g|cf2]^(m+1))
    //                + n_ax(g) / (2 (zeta + eta)) [g - 1_ax|cf1]^(m+1)
    // ax, cf1, cf2, nf depend on the ket component only (constants after unrolling); everything except the last
    // term is the lane's own column.  Unlike the electron transfer of the thread kernels and of the interpreter
    // (decision 7) nothing here is multiplied by zeta/eta, so contracted tight d shells keep their digits.
    template <int FL>
    static __device__ __forceinline__ void ket_vrr(const LaneInfo (&I)[NP], const double (&nh)[NP][3], const double (&QC)[3],
                                                   const double (&WQ)[3], double i2e, double rhoe, double (&acc)[NKET], int lane)
    {
        if constexpr (FL < F) {
            constexpr int f1 = FL + 1;
#pragma unroll
            for (int rf = 0; rf <= f1; ++rf)
#pragma unroll
                for (int kf = 0; kf <= rf; ++kf) {
                    const int fi = f1 - rf, fj = rf - kf;
                    const int ax = fi > 0 ? 0 : (fj > 0 ? 1 : 2);
                    const int fj1 = fj - (ax == 1), fk1 = kf - (ax == 2);
                    const int nf = (ax == 0 ? fi : (ax == 1 ? fj : kf)) - 1;
                    const int cf = CIDX(fj, kf), cf1 = CIDX(fj1, fk1);
                    const int cf2 = nf > 0 ? CIDX(fj1 - (ax == 1), fk1 - (ax == 2)) : 0;
#pragma unroll
                    for (int m = 0; m < nm(f1); ++m) {
                        const int s0 = row(FL, m, cf1), s1 = row(FL, m + 1, cf1), dst = row(f1, m, cf);
                        const int t0 = row(FL > 0 ? FL - 1 : 0, m, cf2), t1 = row(FL > 0 ? FL - 1 : 0, m + 1, cf2);
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (!pass_in_level(q, f1)) continue;
                            const int g = I[q].g;
                            if (g >= glo(f1) && g < ghi(f1)) {
                                double v = fma(QC[ax], I[q].gp[s0], WQ[ax] * I[q].gp[s1]);
                                if (nf > 0) v = fma(nf * i2e, fma(-rhoe, I[q].gp[t1], I[q].gp[t0]), v);
                                v = fma(nh[q][ax], I[q].dn[ax][s1], v);
                                I[q].gp[dst] = v;
                                if (m == 0 && f1 >= LC && q == 0 && g >= GB) acc[NCSUM(LC, f1 - 1) + cf] += v;
                            }
                        }
                    }
                }
            __syncwarp();
            ket_vrr<FL + 1>(I, nh, QC, WQ, i2e, rhoe, acc, lane);
        }
    }
};

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(COOP_WARPS * 32, QBX_COOP_MINB) eri_coop2_kernel(CoopArgs p)
{
    using C2 = Coop2<LA, LB, LC, LD>;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *B = smem + (size_t)wib * C2::BUF;
    typename C2::LaneInfo I[C2::NP];
#pragma unroll
    for (int q = 0; q < C2::NP; ++q) C2::lane_info(q, lane, B, I[q]);
    const int64_t warp0 = (int64_t)blockIdx.x * COOP_WARPS + wib, nwarps = (int64_t)gridDim.x * COOP_WARPS;
    for (int64_t qt = warp0; qt < p.ntasks; qt += nwarps) {
        const int2 t = p.tasks[qt];
        const double *gb = p.bra.geom + 8 * (int64_t)t.x, *gk = p.ket.geom + 8 * (int64_t)t.y;
        const double A[3] = {gb[0], gb[1], gb[2]}, AB[3] = {gb[3], gb[4], gb[5]};
        const double Cc[3] = {gk[0], gk[1], gk[2]}, CD[3] = {gk[3], gk[4], gk[5]};
        const int pb0 = p.bra.prim_off[t.x], pb1 = p.bra.prim_off[t.x + 1];
        const int pk0 = p.ket.prim_off[t.y], pk1 = p.ket.prim_off[t.y + 1];
        double acc[C2::NKET];
#pragma unroll
        for (int i = 0; i < C2::NKET; ++i) acc[i] = 0.0;
        for (int pb = pb0; pb < pb1; ++pb) {
            const double *rb = p.bra.prim + 8 * (int64_t)pb;
            const double zeta = __ldg(rb), P[3] = {__ldg(rb + 1), __ldg(rb + 2), __ldg(rb + 3)};
            const double Kab = __ldg(rb + 4), bx = __ldg(rb + 5), i2z = __ldg(rb + 6);
            const double PA[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
            for (int pk = pk0; pk < pk1; ++pk) {
                const double *rk = p.ket.prim + 8 * (int64_t)pk;
                const double eta = __ldg(rk), Q[3] = {__ldg(rk + 1), __ldg(rk + 2), __ldg(rk + 3)};
                const double Kcd = __ldg(rk + 4), dx = __ldg(rk + 5), i2e = __ldg(rk + 6);
                const double rs = rsqrt(zeta + eta), inv = rs * rs, rz = eta * inv;
                const double PQ[3] = {P[0] - Q[0], P[1] - Q[1], P[2] - Q[2]};
                const double T = zeta * rz * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
                const double WP[3] = {-rz * PQ[0], -rz * PQ[1], -rz * PQ[2]};
                const double rhoe = 1.0 - rz, hinv = 0.5 * inv;       // rho/eta = zeta/(zeta+eta); 1/(2(zeta+eta))
                const double QC[3] = {Q[0] - Cc[0], Q[1] - Cc[1], Q[2] - Cc[2]};
                const double WQ[3] = {rhoe * PQ[0], rhoe * PQ[1], rhoe * PQ[2]};
                double Fm[9];
                boys_rt(p.boys, p.boys_inv, T, Kab * Kcd * rs, C2::L, Fm);
                __syncwarp();                              // the previous primitive's level-0 copy has read V
                if (lane <= C2::L) B[lane] = Fm[lane];
                __syncwarp();
                coop_vrr(p, B, lane, PA, WP, i2z, rz, C2::E);          // bra degrees 0..E only, all orders
                // level 0 of the ket recurrence = V at orders 0..F, copied into the stacked layout
                double nh[C2::NP][3];
#pragma unroll
                for (int q = 0; q < C2::NP; ++q) {
                    const int g = I[q].g;
                    if (g >= C2::glo(0)) {
#pragma unroll
                        for (int m = 0; m < C2::nm(0); ++m) I[q].gp[C2::row(0, m, 0)] = I[q].v0[m * I[q].vs];
                    }
#pragma unroll
                    for (int x = 0; x < 3; ++x) nh[q][x] = I[q].n[x] * hinv;
                }
                if constexpr (LC == 0) {
                    if (I[0].g >= C2::GB) acc[0] += *I[0].v0;
                }
                __syncwarp();
                C2::template ket_vrr<0>(I, nh, QC, WQ, i2e, rhoe, acc, lane);
            }
        }
        // contracted [e0|f0] -> shared, one row per stacked ket component
        __syncwarp();
        if (I[0].g >= C2::GB) {                                 // the lane that owns a contracted component holds its sums
#pragma unroll
            for (int kk = 0; kk < C2::NKET; ++kk) B[C2::acc2_off + kk * C2::ASTR + (I[0].g - C2::GB)] = acc[kk];
        }
        __syncwarp();
        // bra horizontal recurrence: lane = stacked ket component
        if (lane < C2::NKET) {
            const double *row = B + C2::acc2_off + lane * C2::ASTR;
            hrr_apply<LA, LB>([&](int e, int c) { return row[S1(e) - S1(LA) + c]; }, AB,
                              [&](int a, int b, double v) { B[C2::x2_off + (a * C2::NB + b) * C2::XSTR + lane] = v; });
        }
        __syncwarp();
        // ket horizontal recurrence: lane = (a, b); weights; store (component-major)
        const int2 sb = p.bra.shells[t.x], sk = p.ket.shells[t.y];
        const double *sA = p.shell_scale + 6 * sb.x, *sB = p.shell_scale + 6 * sb.y;
        const double *sC = p.shell_scale + 6 * sk.x, *sD = p.shell_scale + 6 * sk.y;
        for (int ab = lane; ab < C2::NAB; ab += 32) {
            const double *row = B + C2::x2_off + ab * C2::XSTR;
            const double sab = sA[ab / C2::NB] * sB[ab % C2::NB];
            hrr_apply<LC, LD>([&](int f, int c) { return row[NCSUM(LC, f - 1) + c]; }, CD,
                              [&](int c, int d, double v) {
                                  p.out[(int64_t)((ab * C2::NCc + c) * C2::ND + d) * p.ntasks + qt] = v * sab * sC[c] * sD[d];
                              });
        }
        __syncwarp();
    }
}

typedef int (*Coop2Launch)(const CoopArgs &, int64_t, cudaStream_t);

template <int LA, int LB, int LC, int LD>
int launch_coop2(const CoopArgs &c, int64_t ntasks, cudaStream_t s)
{
    using C2 = Coop2<LA, LB, LC, LD>;
    const size_t smem = (size_t)COOP_WARPS * C2::BUF * sizeof(double);
    static int per_sm = 0, sms = 0;
    if (per_sm == 0) {
        int dev = 0;
        QBX_CUDA(cudaFuncSetAttribute(eri_coop2_kernel<LA, LB, LC, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QBX_CUDA(cudaGetDevice(&dev));
        QBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        QBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eri_coop2_kernel<LA, LB, LC, LD>, COOP_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
    }
    const int64_t need = (ntasks + COOP_WARPS - 1) / COOP_WARPS, cap = (int64_t)sms * per_sm;
    static const bool trace = getenv("QBX_TRACE_COOP") != nullptr;
    if (trace) fprintf(stderr, "[qbx] coop2 (%d%d|%d%d): %lld quartets, %zu B shared per block, %d blocks/SM\n", LA, LB, LC, LD,
                       (long long)ntasks, smem, per_sm);
    eri_coop2_kernel<LA, LB, LC, LD><<<(unsigned)std::min(need, cap), COOP_WARPS * 32, smem, s>>>(c);
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {                      // a launch-configuration error is reported synchronously: nothing ran
        fprintf(stderr, "[qbx] eri_coop2_kernel<%d,%d,%d,%d> could not be launched (%s); using the table-driven kernel\n", LA, LB, LC,
                LD, cudaGetErrorString(le));
        return -2;
    }
    return QBX_OK;
}

// classes compiled for the second-generation kernel: everything the thread kernels do not serve, plus
// the thread-kernel classes that spill at 255 registers ((dp|pp), (dp|ds), (dd|ps)); those are routed here only when
// QBX_COOP_MIN_ACC is lowered (A/B runs: by instruction count the thread kernels should still win)
Coop2Launch coop2_for(int la, int lb, int lc, int ld)
{
    const int key = la * 1000 + lb * 100 + lc * 10 + ld;
    switch (key) {
        case 2121: return launch_coop2<2, 1, 2, 1>;
        case 2211: return launch_coop2<2, 2, 1, 1>;
        case 2220: return launch_coop2<2, 2, 2, 0>;
        case 2221: return launch_coop2<2, 2, 2, 1>;
        case 2222: return launch_coop2<2, 2, 2, 2>;
        case 2111: return launch_coop2<2, 1, 1, 1>;
        case 2120: return launch_coop2<2, 1, 2, 0>;
        case 2210: return launch_coop2<2, 2, 1, 0>;     // reached only with QBX_COOP_MIN_ACC <= 93 (A/B against the thread kernel)
        default: return nullptr;
    }
}

std::map<int, Program *> g_programs;

}  // namespace

// Returns QBX_OK after launching, or -1 if this class is not served by the cooperative kernel.
int qbx_launch_eri_coop(int la, int lb, int lc, int ld, const ClassArgs &a, cudaStream_t s)
{
    const int key = la * 1000 + lb * 100 + lc * 10 + ld;
    auto it = g_programs.find(key);
    if (it == g_programs.end()) {
        Program *P = build_program(la, lb, lc, ld);
        if (!P) return -1;
        QBX_CUDA(cudaMalloc(&P->d_ops, P->ops.size() * sizeof(Op)));
        QBX_CUDA(cudaMemcpy(P->d_ops, P->ops.data(), P->ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
        it = g_programs.emplace(key, P).first;
    }
    Program *P = it->second;
    if (a.ntasks <= 0) return QBX_OK;
    CoopArgs c{};
    c.bra = a.bra; c.ket = a.ket; c.tasks = a.tasks; c.ntasks = a.ntasks; c.out = a.out;
    c.shell_scale = a.shell_scale; c.boys = a.boys; c.ops = P->d_ops;
    c.L = P->L; c.buf = P->buf_doubles; c.acc_off = P->acc_off; c.acc_n = P->acc_n;
    c.ncomp = P->ncomp; c.y_off = P->y_off;
    c.NA = P->NA; c.NB = P->NB; c.NCc = P->NCc; c.ND = P->ND;
    c.nvrr = (int)P->vrr.size(); c.nxfer = (int)P->xfer.size(); c.nhrr = (int)P->hrr.size();
    for (int i = 0; i < c.nvrr; ++i) c.vrr[i] = P->vrr[i];
    for (int i = 0; i < c.nxfer; ++i) c.xfer[i] = P->xfer[i];
    for (int i = 0; i < c.nhrr; ++i) c.hrr[i] = P->hrr[i];
    c.acc = P->acc;
    for (int k = 0; k < QBX_BOYS_DEG; ++k) c.boys_inv[k] = 1.0 / (2.0 * (P->L + k) + 1.0);
    // The table-driven first-generation kernel (an interpreter over the op table, 1.5-4.4x slower per class: 30.98 vs
    // 28.63 ms for the (H2O)16 ERI pass, profiles/r02/probe_ab_head_of_round1.log) was deleted in round 2; its op table
    // survives as the vertical-recurrence program of the compiled kernel.  A class the compiled kernel does not
    // cover goes to its thread-per-quartet kernel (-1).
    if (Coop2Launch f = coop2_for(la, lb, lc, ld)) {
        const int rc2 = f(c, a.ntasks, s);
        if (rc2 == -2) { qbx_set_error("eri_coop2_kernel could not be launched"); return QBX_ERR_CUDA; }
        return rc2;
    }
    return -1;
}
