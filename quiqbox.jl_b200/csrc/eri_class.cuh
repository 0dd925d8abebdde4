// Shell-quartet ERI kernels for the s/p/d classes (la lb|lc ld), la >= lb, lc >= ld.
//
// What the reference computes one primitive *component* quartet at a time
// (computePGTOrbTwoBodyRepulsion!, src/Integration/Engines/GaussianOrbitals.jl:594-663, then
// K^4 contraction in getOrbLayoutIntegralCore!, src/Integration/Framework.jl:526-554) is done
// here once per primitive *shell* quartet with everything shared between the components:
//
//   per primitive quartet   Boys F_0..F_L (boys.cuh)
//                           vertical recurrence   [e0|00]^(m), e <= L      (cf. vertTransfer :386)
//                           electron transfer     [e0|f0],   f <= lc+ld    (cf. modeTransfer :529)
//                           accumulate [e0|f0], la <= e <= la+lb, lc <= f <= lc+ld
//   per contracted quartet  horizontal recurrences on bra and ket          (cf. horiTransfer :394)
//                           per-component weights, store
//
// One thread owns one contracted shell quartet.  All recurrences are unrolled at compile
// time over Cartesian components, so every index below is a constant in the SASS and small
// classes live entirely in registers.  Values are written component-major
// (out[comp * ntasks + q]) so that a warp's stores are coalesced.
#pragma once
#include "qbx_internal.h"

// ---- compile-time Cartesian bookkeeping (component order = SubshellXYZs, src/Lexicons.jl:17-35:
//      i descending, then j descending; index within degree = r(r+1)/2 + k with r = j + k) ----
__host__ __device__ constexpr int NC(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ constexpr int CIDX(int j, int k) { return (j + k) * (j + k + 1) / 2 + k; }
__host__ __device__ constexpr int S1(int n) { return n * (n + 1) * (n + 2) / 6; }      // # components of degree < n
__host__ __device__ constexpr int NCSUM(int lo, int hi) { return S1(hi + 1) - S1(lo); }

// All layout offsets are closed-form arithmetic so that they fold to literals once the
// component loops are unrolled (every array below is then statically indexed).
// [e0|00]^(m): degree e keeps orders m = 0..L-e;  off(e) = sum_{x<e} NC(x) (L + 1 - x)
template <int L> struct VLay {
    static __host__ __device__ constexpr int off(int e) { return (L + 1) * S1(e) - (e - 1) * e * (e + 1) * (e + 2) / 8; }
    static constexpr int total = off(L + 1);
};
// [e0|f0], f = 1..F: level f keeps e in [elo(f), ehi(f)], stored [cf][ce] per (f, e)
template <int LA, int E, int F> struct WLay {
    static __host__ __device__ constexpr int elo(int f) { return (LA - (F - f)) > 0 ? (LA - (F - f)) : 0; }
    static __host__ __device__ constexpr int ehi(int f) { return E + (F - f); }
    static __host__ __device__ constexpr int lsize(int f) { return NC(f) * (S1(ehi(f) + 1) - S1(elo(f))); }
    static __host__ __device__ constexpr int off(int f, int e) {
        return (f > 1 ? lsize(1) : 0) + (f > 2 ? lsize(2) : 0) + (f > 3 ? lsize(3) : 0) + (f > 4 ? lsize(4) : 0) +
               NC(f) * (S1(e) - S1(elo(f)));
    }
    static constexpr int total = F > 0 ? off(F, ehi(F) + 1) : 1;
};
// contracted accumulators: f in [LC,F], e in [LA,E], stored [cf][ce] per (f, e)
template <int LA, int E, int LC, int F> struct ALay {
    static __host__ __device__ constexpr int off(int f, int e) {
        return (S1(f) - S1(LC)) * NCSUM(LA, E) + NC(f) * (S1(e) - S1(LA));
    }
    static constexpr int total = NCSUM(LA, E) * NCSUM(LC, F);
};

// primitive-pair record (8 doubles, 64 B): zeta, Px, Py, Pz, K, xr, 1/(2 zeta), 1/zeta
//   K  = sqrt(2) pi^(5/4) c_a c_b exp(-a b |AB|^2 / zeta) / zeta      (GaussianOrbitals.jl:627-629)
//   xr = exponent of the right-hand primitive (b or d), needed by the electron transfer
// pair geometry record (8 doubles): A(3), A-B(3), pad(2)
// The same primitive data also exists transposed ("SoA") for the ket side: pairs are sorted
// by their number of primitive pairs, and inside a run of g pairs with equal count the
// field f of primitive p of the j-th pair of the run sits at soa[base + (p*7 + f)*g + j], so
// that the lanes of a warp (consecutive kets) read consecutive addresses.
// Fields: eta, Qx, Qy, Qz, K, d, 1/(2 eta).
#define QBX_SOA_NF 7
struct PairSet {
    const int2 *shells;        // (A, B) shell ids per pair
    const int *prim_off;       // [npair + 1]
    const double *geom;        // [npair][8]
    const double *prim;        // [nprimpair][8]   AoS (bra side: warp-uniform loads)
    const double *soa;         // transposed copy  (ket side: coalesced loads)
    const int2 *soa_idx;       // [npair] (base + j, g)
    int npair;
};

struct ClassArgs {
    PairSet bra, ket;
    const int2 *tasks;         // (bra pair, ket pair) per contracted shell quartet
    int64_t ntasks;
    double *out;               // [ncomp][ntasks]
    const double *shell_scale; // [nshell][6] per-component weights
    BoysTable boys;
    unsigned int *counter;     // work queue head for this launch (zeroed by the host)
    const int *order;          // 32-task chunks in processing order (heaviest first), or null = list order
};

__device__ __forceinline__ double4 ldg4(const double *p)
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ---- horizontal recurrence (a b+1| = (a+1 b| + AB (a b|, from stacked degrees LX..LX+LY ----
template <int LX, int LY, class In, class Out>
__device__ __forceinline__ void hrr_apply(const In &in, const double (&AB)[3], const Out &out)
{
    double t0[LY + 1][NC(LX + LY)];
    // (loop bounds are kept independent of outer loop variables and guarded instead: a bound
    //  like NC(e) is quadratic in the outer index, which stops nvcc from fully unrolling the
    //  inner loop and would push every array touched here into local memory)
#pragma unroll
    for (int d = 0; d <= LY; ++d)
#pragma unroll
        for (int c = 0; c < NC(LX + LY); ++c)
            if (c < NC(LX + d)) t0[d][c] = in(LX + d, c);
    if constexpr (LY == 0) {
#pragma unroll
        for (int c = 0; c < NC(LX); ++c) out(c, 0, t0[0][c]);
    } else {
        double t1[LY][NC(LX + LY - 1)][3];
#pragma unroll
        for (int d = 0; d < LY; ++d)
#pragma unroll
            for (int r = 0; r <= LX + d; ++r)
#pragma unroll
                for (int k = 0; k <= r; ++k) {
                    const int j = r - k, c = CIDX(j, k);
                    t1[d][c][0] = fma(AB[0], t0[d][c], t0[d + 1][CIDX(j, k)]);
                    t1[d][c][1] = fma(AB[1], t0[d][c], t0[d + 1][CIDX(j + 1, k)]);
                    t1[d][c][2] = fma(AB[2], t0[d][c], t0[d + 1][CIDX(j, k + 1)]);
                }
        if constexpr (LY == 1) {
#pragma unroll
            for (int c = 0; c < NC(LX); ++c)
#pragma unroll
                for (int x = 0; x < 3; ++x) out(c, x, t1[0][c][x]);
        } else {
            static_assert(LY == 2, "class kernels stop at d");
#pragma unroll
            for (int r = 0; r <= LX; ++r)
#pragma unroll
                for (int k = 0; k <= r; ++k) {
                    const int j = r - k, c = CIDX(j, k);
                    // d components in SubshellXYZs order: xx xy xz yy yz zz, lowering axis = first non-zero
                    out(c, 0, fma(AB[0], t1[0][c][0], t1[1][CIDX(j, k)][0]));         // xx = x + x
                    out(c, 1, fma(AB[0], t1[0][c][1], t1[1][CIDX(j, k)][1]));         // xy = x + y
                    out(c, 2, fma(AB[0], t1[0][c][2], t1[1][CIDX(j, k)][2]));         // xz = x + z
                    out(c, 3, fma(AB[1], t1[0][c][1], t1[1][CIDX(j + 1, k)][1]));     // yy = y + y
                    out(c, 4, fma(AB[1], t1[0][c][2], t1[1][CIDX(j + 1, k)][2]));     // yz = y + z
                    out(c, 5, fma(AB[2], t1[0][c][2], t1[1][CIDX(j, k + 1)][2]));     // zz = z + z
                }
        }
    }
}

template <int LA, int LB, int LC, int LD>
struct EriClass {
    static constexpr int E = LA + LB, F = LC + LD, L = E + F;
    static constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD);
    static constexpr int NCOMP = NA * NB * NCc * ND;
    using VL = VLay<L>;
    using WL = WLay<LA, E, F>;
    using AL = ALay<LA, E, LC, F>;
    static constexpr int NKET = NCSUM(LC, F);       // stacked ket components after contraction

    // one primitive shell quartet: acc += [e0|f0]
    static __device__ __forceinline__ void primitive(double (&acc)[AL::total], const double *tb, double zeta,
                                                     const double (&P)[3], double Kab, const double (&PA)[3],
                                                     double i2z, const double (&bAB)[3], double eta,
                                                     const double (&Q)[3], double Kcd, double i2e,
                                                     const double (&dCD)[3])
    {
        const double zpe = zeta + eta;
        const double rs = rsqrt(zpe), inv = rs * rs;
        const double rz = eta * inv;                      // rho / zeta
        const double PQ[3] = {P[0] - Q[0], P[1] - Q[1], P[2] - Q[2]};
        const double T = zeta * rz * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
        double Fm[L + 1];
        boys_table<L>(tb, T, Kab * Kcd * rs, Fm);
        if constexpr (L == 0) {
            acc[0] += Fm[0];
        } else {
            double V[VL::total];
#pragma unroll
            for (int m = 0; m <= L; ++m) V[m] = Fm[m];
            const double WP[3] = {-rz * PQ[0], -rz * PQ[1], -rz * PQ[2]};
            // vertical recurrence on centre A
#pragma unroll
            for (int e = 0; e < L; ++e)
#pragma unroll
                for (int m = 0; m <= L - e - 1; ++m)
#pragma unroll
                    for (int r = 0; r <= e + 1; ++r)
#pragma unroll
                        for (int k = 0; k <= r; ++k) {
                            const int i = e + 1 - r, j = r - k;
                            const int ax = i > 0 ? 0 : (j > 0 ? 1 : 2);
                            const int i1 = i - (ax == 0), j1 = j - (ax == 1), k1 = k - (ax == 2);
                            const int n1 = ax == 0 ? i1 : (ax == 1 ? j1 : k1);
                            const int c1 = CIDX(j1, k1);
                            double v = fma(PA[ax], V[VL::off(e) + m * NC(e) + c1],
                                           WP[ax] * V[VL::off(e) + (m + 1) * NC(e) + c1]);
                            if (n1 > 0) {
                                const int c2 = CIDX(j1 - (ax == 1), k1 - (ax == 2));
                                const int ee = e > 0 ? e - 1 : 0;
                                v = fma(n1 * i2z, fma(-rz, V[VL::off(ee) + (m + 1) * NC(ee) + c2],
                                                      V[VL::off(ee) + m * NC(ee) + c2]), v);
                            }
                            V[VL::off(e + 1) + m * NC(e + 1) + CIDX(j, k)] = v;
                        }
            if constexpr (F == 0) {
#pragma unroll
                for (int e = LA; e <= E; ++e)
#pragma unroll
                    for (int c = 0; c < NC(E); ++c)
                        if (c < NC(e)) acc[AL::off(0, e) + c] += V[VL::off(e) + c];
            } else {
                // electron transfer to centre C at m = 0
                double W[WL::total];
                const double ie = 2.0 * i2e, zoe = zeta * ie;
                const double k0[3] = {-(bAB[0] + dCD[0]) * ie, -(bAB[1] + dCD[1]) * ie, -(bAB[2] + dCD[2]) * ie};
#pragma unroll
                for (int f = 0; f < F; ++f)                      // build level f + 1
#pragma unroll
                    for (int rf = 0; rf <= f + 1; ++rf)
#pragma unroll
                        for (int kf = 0; kf <= rf; ++kf) {
                            const int fi = f + 1 - rf, fj = rf - kf;
                            const int ax = fi > 0 ? 0 : (fj > 0 ? 1 : 2);
                            const int fj1 = fj - (ax == 1), fk1 = kf - (ax == 2);
                            const int nf = (ax == 0 ? fi : (ax == 1 ? fj : kf)) - 1;
                            const int cf = CIDX(fj, kf), cf1 = CIDX(fj1, fk1);
                            const int cf2 = nf > 0 ? CIDX(fj1 - (ax == 1), fk1 - (ax == 2)) : 0;
#pragma unroll
                            for (int e = WL::elo(f + 1); e <= WL::ehi(f + 1); ++e)
#pragma unroll
                                for (int re = 0; re <= e; ++re)
#pragma unroll
                                    for (int ke = 0; ke <= re; ++ke) {
                                        const int ei = e - re, ej = re - ke;
                                        const int ce = CIDX(ej, ke);
                                        const int na = ax == 0 ? ei : (ax == 1 ? ej : ke);
                                        const int cup = CIDX(ej + (ax == 1), ke + (ax == 2));
                                        double v;
                                        if (f == 0) {
                                            v = fma(k0[ax], V[VL::off(e) + ce], -zoe * V[VL::off(e + 1) + cup]);
                                            if (na > 0)
                                                v = fma(na * i2e, V[VL::off(e > 0 ? e - 1 : 0) + CIDX(ej - (ax == 1), ke - (ax == 2))], v);
                                        } else {
                                            v = fma(k0[ax], W[WL::off(f, e) + cf1 * NC(e) + ce],
                                                    -zoe * W[WL::off(f, e + 1) + cf1 * NC(e + 1) + cup]);
                                            if (na > 0)
                                                v = fma(na * i2e,
                                                        W[WL::off(f, e > 0 ? e - 1 : 0) + cf1 * NC(e > 0 ? e - 1 : 0) +
                                                          CIDX(ej - (ax == 1), ke - (ax == 2))], v);
                                            if (nf > 0) {
                                                if (f == 1) v = fma(nf * i2e, V[VL::off(e) + ce], v);
                                                else v = fma(nf * i2e, W[WL::off(f > 1 ? f - 1 : 1, e) + cf2 * NC(e) + ce], v);
                                            }
                                        }
                                        W[WL::off(f + 1, e) + cf * NC(e) + ce] = v;
                                    }
                        }
#pragma unroll
                for (int f = LC; f <= F; ++f)
#pragma unroll
                    for (int e = LA; e <= E; ++e)
#pragma unroll
                        for (int c = 0; c < NC(F) * NC(E); ++c) {
                            if (f == 0) { if (c < NC(e)) acc[AL::off(0, e) + c] += V[VL::off(e) + c]; }
                            else if (c < NC(f) * NC(e)) acc[AL::off(f, e) + c] += W[WL::off(f, e) + c];
                        }
            }
        }
    }

    // contracted [e0|f0] -> (ab|cd), weights, store
    static __device__ __forceinline__ void finish(const double (&acc)[AL::total], const double (&AB)[3],
                                                  const double (&CD)[3], const double *sA, const double *sB,
                                                  const double *sC, const double *sD, double *out, int64_t stride)
    {
        // bra: for every stacked ket component kk, X[kk][a*NB + b]
        double X[NKET][NA * NB];
#pragma unroll
        for (int f = LC; f <= F; ++f)
#pragma unroll
            for (int cf = 0; cf < NC(F); ++cf) {
                if (cf >= NC(f)) continue;
                const int kk = NCSUM(LC, f - 1) + cf;
                hrr_apply<LA, LB>([&](int e, int c) { return acc[AL::off(f, e) + cf * NC(e) + c]; }, AB,
                                  [&](int a, int b, double v) { X[kk][a * NB + b] = v; });
            }
#pragma unroll
        for (int ab = 0; ab < NA * NB; ++ab) {
            const double sab = sA[ab / NB] * sB[ab % NB];
            hrr_apply<LC, LD>([&](int f, int c) { return X[NCSUM(LC, f - 1) + c][ab]; }, CD,
                              [&](int c, int d, double v) {
                                  out[(int64_t)((ab * NCc + c) * ND + d) * stride] = v * sab * sC[c] * sD[d];
                              });
        }
    }
};

#ifndef QBX_ERI_THREADS
#define QBX_ERI_THREADS 256
#endif

// Persistent blocks: the Boys columns of this class are staged in shared memory once, then
// its warps pull chunks of 32 consecutive tasks from a global work queue.  The list is
// ordered heavy-first (rows by bra contraction length, kets likewise), and a single task spans
// 1 to 6561 primitive quartets, so a static grid-stride assignment leaves most of the GPU idle
// behind the few threads that drew a heavy task (it halved the throughput at 1/8 of the list,
// i.e. on 8 GPUs); with the queue the light chunks at the end fill the tail.
// resident blocks the register allocation aims at, by total angular momentum (A/B builds: tools/gpu_ab_eri.sh)
// (B200, (H2O)16, profiles/r02/ab_eri_launch_bounds.log: L = 3 classes at 2 blocks of 256 threads = 128 registers:
// (pp|ps) 2.80 -> 2.50 ms, (ds|ps) 2.19 -> 1.75 ms; at 3 blocks 4.56 / 2.71 ms; L = 4 classes at 2 blocks: (pp|pp) 0.59 -> 1.50 ms.)
#ifndef QBX_ERI_MINB_L2
#define QBX_ERI_MINB_L2 3
#endif
#ifndef QBX_ERI_MINB_L3
#define QBX_ERI_MINB_L3 2
#endif
#ifndef QBX_ERI_MINB_L4
#define QBX_ERI_MINB_L4 1
#endif
__host__ __device__ constexpr int eri_min_blocks(int L) { return L <= 2 ? QBX_ERI_MINB_L2 : (L == 3 ? QBX_ERI_MINB_L3 : (L == 4 ? QBX_ERI_MINB_L4 : 1)); }

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(QBX_ERI_THREADS, eri_min_blocks(LA + LB + LC + LD)) eri_class_kernel(ClassArgs p)
{
    using EC = EriClass<LA, LB, LC, LD>;
    extern __shared__ double boys_smem[];
    boys_stage_smem<EC::L>(p.boys, boys_smem);
    __syncthreads();
    // every WARP pulls its own 32-task chunks: a block-wide queue needs a barrier per chunk, and
    // the warps then wait for the slowest task of the block (ncu: 5 of 9 stall cycles per issue)
    const int64_t nchunk = (p.ntasks + 31) / 32;
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned int c = 0;
        if (lane == 0) c = atomicAdd(p.counter, 1u);
        const int64_t k = __shfl_sync(0xffffffffu, c, 0);
        if (k >= nchunk) break;
        const int64_t chunk = p.order ? p.order[k] : k;
        const int64_t q = chunk * 32 + lane;
        if (q >= p.ntasks) continue;
        const int2 t = p.tasks[q];
        const double *gb = p.bra.geom + 8 * (int64_t)t.x, *gk = p.ket.geom + 8 * (int64_t)t.y;
        const double A[3] = {gb[0], gb[1], gb[2]}, AB[3] = {gb[3], gb[4], gb[5]};
        const double CD[3] = {gk[3], gk[4], gk[5]};
        const int pb0 = p.bra.prim_off[t.x], pb1 = p.bra.prim_off[t.x + 1];
        const int nk = p.ket.prim_off[t.y + 1] - p.ket.prim_off[t.y];
        const int2 si = p.ket.soa_idx[t.y];
        double acc[EC::AL::total];
#pragma unroll
        for (int i = 0; i < EC::AL::total; ++i) acc[i] = 0.0;
        for (int pb = pb0; pb < pb1; ++pb) {
            const double4 b0 = ldg4(p.bra.prim + 8 * (int64_t)pb);
            const double4 b1 = ldg4(p.bra.prim + 8 * (int64_t)pb + 4);
            const double zeta = b0.x, P[3] = {b0.y, b0.z, b0.w}, Kab = b1.x, i2z = b1.z;
            const double PA[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
            const double bAB[3] = {b1.y * AB[0], b1.y * AB[1], b1.y * AB[2]};
            const double *kp = p.ket.soa + si.x;
            for (int pk = 0; pk < nk; ++pk, kp += QBX_SOA_NF * si.y) {
                const double eta = __ldg(kp);
                const double Q[3] = {__ldg(kp + si.y), __ldg(kp + 2 * si.y), __ldg(kp + 3 * si.y)};
                const double Kcd = __ldg(kp + 4 * si.y), dx = __ldg(kp + 5 * si.y), i2e = __ldg(kp + 6 * si.y);
                const double dCD[3] = {dx * CD[0], dx * CD[1], dx * CD[2]};
                EC::primitive(acc, boys_smem, zeta, P, Kab, PA, i2z, bAB, eta, Q, Kcd, i2e, dCD);
            }
        }
        const int2 sb = p.bra.shells[t.x], sk = p.ket.shells[t.y];
        EC::finish(acc, AB, CD, p.shell_scale + 6 * sb.x, p.shell_scale + 6 * sb.y, p.shell_scale + 6 * sk.x,
                   p.shell_scale + 6 * sk.y, p.out + q, p.ntasks);
    }
}

// One WARP per task, the lanes split the ket primitives and meet in a shuffle reduction: for SHORT lists of
// HEAVY tasks, i.e. the Schwarz diagonals (ab|ab).  A lone thread needs ~800 cycles per primitive quartet (one
// dependent chain), so with one task per thread the 6561 primitive quartets of a (9s 9s|9s 9s) diagonal set the
// duration of the whole Schwarz pass (3.3 ms of a 61 ms create -> store -> Fock build step) while the GPU idles.
template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(QBX_ERI_THREADS) eri_class_split_kernel(ClassArgs p)
{
    using EC = EriClass<LA, LB, LC, LD>;
    extern __shared__ double boys_smem[];
    boys_stage_smem<EC::L>(p.boys, boys_smem);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = warp0; q < p.ntasks; q += nwarps) {
        const int2 t = p.tasks[q];
        const double *gb = p.bra.geom + 8 * (int64_t)t.x, *gk = p.ket.geom + 8 * (int64_t)t.y;
        const double A[3] = {gb[0], gb[1], gb[2]}, AB[3] = {gb[3], gb[4], gb[5]};
        const double CD[3] = {gk[3], gk[4], gk[5]};
        const int pb0 = p.bra.prim_off[t.x], pb1 = p.bra.prim_off[t.x + 1];
        const int nk = p.ket.prim_off[t.y + 1] - p.ket.prim_off[t.y];
        const int2 si = p.ket.soa_idx[t.y];
        double acc[EC::AL::total];
#pragma unroll
        for (int i = 0; i < EC::AL::total; ++i) acc[i] = 0.0;
        for (int pk = lane; pk < nk; pk += 32) {
            const double *kp = p.ket.soa + si.x + (int64_t)pk * QBX_SOA_NF * si.y;
            const double eta = __ldg(kp);
            const double Q[3] = {__ldg(kp + si.y), __ldg(kp + 2 * si.y), __ldg(kp + 3 * si.y)};
            const double Kcd = __ldg(kp + 4 * si.y), dx = __ldg(kp + 5 * si.y), i2e = __ldg(kp + 6 * si.y);
            const double dCD[3] = {dx * CD[0], dx * CD[1], dx * CD[2]};
            for (int pb = pb0; pb < pb1; ++pb) {
                const double4 b0 = ldg4(p.bra.prim + 8 * (int64_t)pb);
                const double4 b1 = ldg4(p.bra.prim + 8 * (int64_t)pb + 4);
                const double zeta = b0.x, P[3] = {b0.y, b0.z, b0.w}, Kab = b1.x, i2z = b1.z;
                const double PA[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
                const double bAB[3] = {b1.y * AB[0], b1.y * AB[1], b1.y * AB[2]};
                EC::primitive(acc, boys_smem, zeta, P, Kab, PA, i2z, bAB, eta, Q, Kcd, i2e, dCD);
            }
        }
#pragma unroll
        for (int i = 0; i < EC::AL::total; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (lane == 0) {
            const int2 sb = p.bra.shells[t.x], sk = p.ket.shells[t.y];
            EC::finish(acc, AB, CD, p.shell_scale + 6 * sb.x, p.shell_scale + 6 * sb.y, p.shell_scale + 6 * sk.x,
                       p.shell_scale + 6 * sk.y, p.out + q, p.ntasks);
        }
    }
}
