// Consumers of packed shell-quartet blocks: J/K digestion (= getGcore,
// src/HartreeFock.jl:305-319) and the dense-tensor scatter (= the 8 permutational stores of
// getOrbVectorIntegralCore!, src/Integration/Framework.jl:659-665).
//
// getGcore sums over all ordered index quartets,
//     G[mu,nu] = sum DJ[sg,lm] (mu nu|lm sg) - sum DK[lm,sg] (mu lm|sg nu).
// A unique shell quartet (AB|CD) (A >= B, C >= D, AB >= CD) stands for 8 ordered images; with
// the block factor f = (A==B ? 1/2 : 1)(C==D ? 1/2 : 1)(AB==CD ? 1/2 : 1) and symmetric
// densities every value v = (ab|cd) of the block contributes 6 updates to the half
// accumulators
//     Jt[ab] += 2 f DJ[cd] v     Jt[cd] += 2 f DJ[ab] v
//     Kt[ad] += f DK[bc] v   Kt[bd] += f DK[ac] v   Kt[ac] += f DK[bd] v   Kt[bc] += f DK[ad] v
// and G = (Jt + Jt^T) - (Kt + Kt^T) (k_finish_G in engine.cu), so either element of a
// transposed pair may be written.
//
// Everything here works in the engine's INTERNAL function numbering: shells are sorted by
// (l, contraction length), functions are numbered shell by shell, shell pairs (C,D) are listed
// C-major with D ascending, and tasks are bra-major.  A warp's 32 consecutive quartets then
// (almost always) share A, B and C and run over consecutive D, so that
//   * J[ab], K[ac], K[bc] are warp-uniform addresses  -> reduced with shuffles, one RED per warp;
//   * J[cd], K[ad], K[bd] and the densities D[cd], D[bd], D[ad] are indexed by the D function,
//     which is consecutive over the lanes           -> coalesced loads and REDs
//     (always written/read as element [d-side + N * other]);
//   * D[ab], D[bc], D[ac] are warp-uniform loads.
// One thread owns one quartet and first reduces its block into per-thread partial sums.
#pragma once
#include "engine.h"

#ifndef QBX_DIGEST_SPREAD
#define QBX_DIGEST_SPREAD 384
#endif

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Segmented warp reduction: a segment is a run of consecutive lanes that update the same address
// (same bra pair, or same bra pair and same shell C); `end` is the last lane of the caller's run.
// After the five steps the FIRST lane of every run holds the run's sum.  A warp that sits inside
// one run (the common case) does exactly the work of a butterfly sum; a warp that straddles two or
// three runs -- every other warp on a screened list, where a (bra pair, C) run is some tens of kets
// long, and every warp of the general-contraction classes -- still issues one RED per run instead
// of one per lane (ncu, profiles/r01: the kernels are bound by the L2 RED request rate).
__device__ __forceinline__ double seg_sum(double v, int end, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o <= end) v += y;
    }
    return v;
}

// runs of equal keys over the lanes: first-lane flag and last lane of the caller's run
__device__ __forceinline__ void seg_runs(int key0, int key1, int lane, bool &head, int &end)
{
    const int p0 = __shfl_up_sync(0xffffffffu, key0, 1), p1 = __shfl_up_sync(0xffffffffu, key1, 1);
    head = lane == 0 || p0 != key0 || p1 != key1;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u));
    end = above ? __ffs(above) - 2 : 31;
}

// K updates of one bra-side function x (= a or b) against the ket functions: K[x c] is one address
// per run of lanes that share the bra pair and C (segmented shuffle reduction, one RED per run),
// K[x d] is consecutive over the lanes.
template <int NCc, int ND>
__device__ __forceinline__ void digest_flush_row(double *Krow, int ic, int id, const double *kxc,
                                                 const double *kxd, double f, bool headC, int endC, bool valid, int lane)
{
#pragma unroll
    for (int c = 0; c < NCc; ++c) {
        const double x = seg_sum(f * kxc[c], endC, lane);
        if (headC && x != 0.0) atomicAdd(Krow + ic + c, x);
    }
    if (valid) {
#pragma unroll
        for (int d = 0; d < ND; ++d) atomicAdd(Krow + id + d, f * kxd[d]);
    }
}

// Kernel structure (ncu, profiles/r01: nothing is saturated, the kernel waits on memory):
//   * the values of a quartet do not depend on its task record, so the first slab of value loads
//     is issued before anything else and overlaps the chain task -> pair info -> densities;
//   * values are consumed in slabs of up to 64 independent loads (whole quartet, one `a`, or one
//     (a,b) slice) -- the bytes in flight per SM are what decides the HBM rate here;
//   * accumulators of row a (K[ac], K[ad]) are flushed as soon as a is finished, those of row b
//     stay in registers only when they are few (QBX_DIGEST_KEEP_B), so that no class spills;
//   * the block factor f multiplies the sums at flush time, not the values.
// Values are read once and feed J and K together (the second exchange density of a UHF build
// re-reads them).
#ifndef QBX_DIGEST_SLAB
#define QBX_DIGEST_SLAB 64
#endif
#ifndef QBX_DIGEST_KEEP_B
#define QBX_DIGEST_KEEP_B 18
#endif

// resident blocks per SM the register allocation aims at (128 threads each)
__host__ __device__ constexpr int digest_min_blocks(int ncomp)
{
    return ncomp <= 3 ? 10 : (ncomp <= 9 ? 6 : (ncomp <= 36 ? 4 : (ncomp <= 162 ? 3 : 2)));
}

// J/K updates of one quartet per lane.  v holds slab 0 of the quartet's values on entry.
template <int LA, int LB, int LC, int LD, int SL>
__device__ __forceinline__ void digest_quartet(const DigestArgs &p, const double *__restrict__ vq, bool valid, int2 t, int4 rb,
                                               int4 rk, double (&v)[SL * NC(LC) * NC(LD)])
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD), NCD = NCc * ND;
    constexpr bool KEEP_B = NB * (NCc + ND) <= QBX_DIGEST_KEEP_B;     // K[bc], K[bd] live across a
    constexpr bool KEEP_DCD = NCD <= 18;                              // D[cd] in registers
    const int64_t nt = p.ntasks;
    double f = valid ? 1.0 : 0.0;
    if (rb.x == rb.y) f *= 0.5;
    if (rk.x == rk.y) f *= 0.5;
    if (p.same_class && t.x == t.y) f *= 0.5;
    const int lane = threadIdx.x & 31;
    bool headAB, headC;
    int endAB, endC;
    seg_runs(t.x, 0, lane, headAB, endAB);                    // runs of one bra pair: J[ab]
    seg_runs(t.x, rk.x, lane, headC, endC);                   // runs of one bra pair and one shell C: K[ac], K[bc]
    const int N = p.nbf;                                      // internal dimension
    const int ia = rb.z, ib = rb.w, ic = rk.z, id = rk.w;
    const double *__restrict__ DJ = p.DJ;

    for (int m = 0; m < p.nmat; ++m) {
        const double *__restrict__ dA = p.DK + (int64_t)m * N * N + (int64_t)N * ia;
        const double *__restrict__ dB = p.DK + (int64_t)m * N * N + (int64_t)N * ib;
        double *kA = p.Kt + (int64_t)m * N * N + (int64_t)N * ia;
        double *kB = p.Kt + (int64_t)m * N * N + (int64_t)N * ib;
        double *jab = p.Jt + ib + (int64_t)N * ia;
        const int jsa = N;
        const bool coul = (m == 0);
        double dcd[KEEP_DCD ? NCD : 1], jcd[NCD];
        double dbc[KEEP_B ? NB * NCc : NCc], dbd[KEEP_B ? NB * ND : ND], kbc[KEEP_B ? NB * NCc : NCc], kbd[KEEP_B ? NB * ND : ND];
#pragma unroll
        for (int cd = 0; cd < NCD; ++cd) jcd[cd] = 0.0;
        if (KEEP_DCD) {
#pragma unroll
            for (int c = 0; c < NCc; ++c)
#pragma unroll
                for (int d = 0; d < ND; ++d) dcd[c * ND + d] = coul ? DJ[(id + d) + N * (ic + c)] : 0.0;
        }
        if (KEEP_B) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int c = 0; c < NCc; ++c) { dbc[b * NCc + c] = dB[(ic + c) + N * b]; kbc[b * NCc + c] = 0.0; }
#pragma unroll
                for (int d = 0; d < ND; ++d) { dbd[b * ND + d] = dB[(id + d) + N * b]; kbd[b * ND + d] = 0.0; }
            }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            double dac[NCc], dad[ND], kac[NCc], kad[ND];
#pragma unroll
            for (int c = 0; c < NCc; ++c) { dac[c] = dA[(ic + c) + N * a]; kac[c] = 0.0; }
#pragma unroll
            for (int d = 0; d < ND; ++d) { dad[d] = dA[(id + d) + N * a]; kad[d] = 0.0; }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int ab = a * NB + b, s0 = (ab % SL) * NCD;
                if (ab % SL == 0 && (ab > 0 || m > 0)) {             // next slab of values
#pragma unroll
                    for (int i = 0; i < SL * NCD; ++i) v[i] = __ldg(vq + (int64_t)(ab * NCD + i) * nt);
                }
                const int ob = KEEP_B ? b : 0;
                if (!KEEP_B) {
#pragma unroll
                    for (int c = 0; c < NCc; ++c) { dbc[c] = dB[(ic + c) + N * b]; kbc[c] = 0.0; }
#pragma unroll
                    for (int d = 0; d < ND; ++d) { dbd[d] = dB[(id + d) + N * b]; kbd[d] = 0.0; }
                }
                const double dab = coul ? DJ[(ib + b) + N * (ia + a)] : 0.0;
                double j = 0.0;
#pragma unroll
                for (int c = 0; c < NCc; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double x = v[s0 + c * ND + d];
                        const double dd = KEEP_DCD ? dcd[c * ND + d] : (coul ? DJ[(id + d) + N * (ic + c)] : 0.0);
                        j = fma(dd, x, j);
                        jcd[c * ND + d] = fma(dab, x, jcd[c * ND + d]);
                        kac[c] = fma(dbd[ob * ND + d], x, kac[c]);
                        kad[d] = fma(dbc[ob * NCc + c], x, kad[d]);
                        kbc[ob * NCc + c] = fma(dad[d], x, kbc[ob * NCc + c]);
                        kbd[ob * ND + d] = fma(dac[c], x, kbd[ob * ND + d]);
                    }
                if (coul) {                                   // J[ab]: one address per warp
                    j = seg_sum(j * 2.0 * f, endAB, lane);
                    if (headAB && j != 0.0) atomicAdd(jab + b + jsa * a, j);
                }
                if (!KEEP_B) digest_flush_row<NCc, ND>(kB + N * b, ic, id, kbc, kbd, f, headC, endC, valid, lane);
            }
            digest_flush_row<NCc, ND>(kA + N * a, ic, id, kac, kad, f, headC, endC, valid, lane);
        }
        if (KEEP_B) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
                digest_flush_row<NCc, ND>(kB + N * b, ic, id, kbc + b * NCc, kbd + b * ND, f, headC, endC, valid, lane);
        }
        if (coul && valid) {                                  // J[cd]: consecutive over the lanes
#pragma unroll
            for (int c = 0; c < NCc; ++c)
#pragma unroll
                for (int d = 0; d < ND; ++d) atomicAdd(p.Jt + (id + d) + N * (ic + c), 2.0 * f * jcd[c * ND + d]);
        }
    }
}

// (A three-stage register pipeline over several tiles per block -- task records of tile i+2, pair
// info and values of tile i+1 in flight while tile i is digested -- was measured and is not
// faster: 13.9 ms against 13.1 ms for the whole (H2O)16 build.  The kernel is bound by the L2
// RED rate and the L1TEX request rate of the density gathers, not by the latency chain.)
template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(128, digest_min_blocks(NC(LA) * NC(LB) * NC(LC) * NC(LD))) digest_kernel(DigestArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCD = NC(LC) * NC(LD), NAB = NA * NB;
    // (a,b) slices per slab of value loads
    constexpr int SL = (NAB * NCD <= QBX_DIGEST_SLAB) ? NAB : ((NB * NCD <= QBX_DIGEST_SLAB) ? NB : 1);
    // Block order is transposed against task order: the blocks resident at one time sit
    // spread-th of the list apart, i.e. on different bra rows.  In task order they would
    // all update the same few hundred G elements, and fp64 REDs on a hot set that small run at
    // 1e10/s instead of 2e11/s on B200 (tools/redbench.cu; profiles/r01/redbench.md).
    const unsigned nblk = (unsigned)((p.ntasks + 127) / 128);
    const unsigned R = nblk < (unsigned)p.spread ? nblk : (unsigned)p.spread, C = (nblk + R - 1) / R;
    const unsigned blk = (blockIdx.x % R) * C + blockIdx.x / R;
    if (blk >= nblk) return;
    const int64_t q0 = (int64_t)blk * 128 + threadIdx.x;
    if (q0 - (threadIdx.x & 31) >= p.ntasks) return;          // whole warp past the end
    const int64_t q = q0 < p.ntasks ? q0 : p.ntasks - 1;
    const double *__restrict__ vq = p.vals + q;
    double v[SL * NCD];
#pragma unroll
    for (int i = 0; i < SL * NCD; ++i) v[i] = __ldg(vq + (int64_t)i * p.ntasks);    // slab 0, before the dependent chain
    int2 t = __ldg(p.tasks + q);
    const bool valid = q0 < p.ntasks && t.y >= 0;              // ket = -1: unused slot of a group task
    if (t.y < 0) t.y = 0;
    const int4 rb = __ldg(p.bra_info + t.x), rk = __ldg(p.ket_info + t.y);   // (shell A, shell B, first A, first B)
    digest_quartet<LA, LB, LC, LD, SL>(p, vq, valid, t, rb, rk, v);
}


// ------------------------------------------------------------------------------------------------
// Span digestion (the stored-mode and direct-mode Fock build; digest_kernel above remains the fallback
// for bases whose rows do not fit shared memory).
//
// Measured on (H2O)16/cc-pVDZ (profiles/r01, r02): digest_kernel sends 0.66 fp64 REDs per stored value
// to L2 (8.75e8 per build), gathers ~10 density elements per quartet through L1 and spends most of its
// instructions on segmented shuffle reductions: 12.3 ms against 1.65 ms for the bytes.  A first span
// kernel that only moved the K rows to shared memory removed the REDs and got SLOWER (17.5 ms): 4.3e9
// warp instructions per build, most of them the shuffle reductions of K[x,c] over runs of lanes with the
// same shell C, with the warps stalled on the shuffle / shared-memory latency chain (short scoreboard 4.5
// of the issue slots).  The reductions exist only because consecutive kets of the list share C; they
// disappear with the ORDER of the shell-pair lists:
//   * shell pairs of a class are listed diagonal by diagonal ((C_i, D_(i+k)), i = 0, 1, ..; engine.cu:
//     Engine::upload), so the 32 kets of a warp have 32 different shells C and 32 different shells D
//     (Schwarz screening only thins the sequence out);
//   * a WARP walks a contiguous span of the bra-major list and keeps, for its current bra pair, the K rows
//     of the bra functions AND their exchange-density rows in its private slice of shared memory, cut
//     down to the columns the class can touch (the functions of angular momentum lc and ld are two
//     contiguous ranges of the internal numbering): nmat 2 (NA + NB) W doubles, 3.6 KB for (ss|ss), 7 KB for
//     (ps|ss), 29 KB for (pp|ps) at (H2O)16;
//   * all four exchange updates of a value are then plain LDS / DFMA / STS on addresses no other lane
//     of the warp touches -- no atomics (fp64 shared atomics are CAS loops), no shuffles, no barriers
//     other than __syncwarp.  Lanes that do hold the same shell (a run boundary, the member slots of a
//     general-contraction task) are found with __match_any_sync and take turns;
//   * J[ab] is a sum over the whole row: per-lane partial sums in registers (small bra classes) or one
//     shuffle sum per tile into shared memory, flushed with the rows.
// Only J[cd] += 2 f DJ[ab] v remains a global RED per ket component and quartet, ~2e8 per build.  Rows are
// flushed with REDs of their non-zero entries when the warp moves to another bra pair and at the end of
// a span.  The task record and the ket info of the NEXT tile, and for the classes with few values per
// quartet also its values, are loaded before the current tile is digested.
#define QBX_SPAN_MAX_SMEM (224 * 1024)
#ifndef QBX_SPAN_PREFETCH_VALUES
#define QBX_SPAN_PREFETCH_VALUES 9
#endif

struct SpanRows {
    double *k;           // this warp's K rows:  element (m, x, j) at k[(m * NX + x) * W + j], x = a or NA + b, j = local column
    const double *d;     // exchange-density rows, same layout
    double *jab;         // NA * NB (used when the bra class is too large for registers)
    int W;
};

// Lanes of a tile that hold the same shell C take turns in the K[x,c] updates (turn = number of lower lanes with
// that C), lanes with the same shell D in the K[x,d] updates.  With the diagonal pair order both counts are 1
// except at a run boundary and in the general-contraction classes (member slots (c_m, d_n): 3 and 3).
struct SpanLanes {
    bool mine;
    int turnC, nturnC, turnD, nturnD;
};

// K[x, col .. col + n) += f * val[0 .. n) on the warp's own shared-memory row
template <int NV_>
__device__ __forceinline__ void span_rmw(double *row, int col, const double *val, double f, bool mine, int turn, int nturn)
{
    if (nturn == 1) {
        if (mine) {
#pragma unroll
            for (int i = 0; i < NV_; ++i) row[col + i] += f * val[i];
        }
    } else {
        for (int r = 0; r < nturn; ++r) {
            if (mine && turn == r) {
#pragma unroll
                for (int i = 0; i < NV_; ++i) row[col + i] += f * val[i];
            }
            __syncwarp();
        }
    }
}

// both halves of one bra function's K row; the barrier between them covers a lane's C against another lane's D
template <int NCc, int ND>
__device__ __forceinline__ void span_flush_row(double *row, int jc, int jd, const double *kxc, const double *kxd, double f,
                                               const SpanLanes &L)
{
    span_rmw<ND>(row, jd, kxd, f, L.mine, L.turnD, L.nturnD);
    __syncwarp();
    span_rmw<NCc>(row, jc, kxc, f, L.mine, L.turnC, L.nturnC);
    __syncwarp();
}

// J/K updates of one quartet per lane (f = 0 for lanes that sit out); v holds slab 0 of the values on entry.
// ia, ib, ic, id: first internal function of the four shells; jc, jd: local columns of C and D in the rows.
template <int LA, int LB, int LC, int LD, int SL, bool JREG>
__device__ __forceinline__ void span_quartet(const DigestArgs &p, const double *__restrict__ vq, int64_t nt, double f, int ia, int ib,
                                             int ic, int id, int jc, int jd, const SpanLanes &L, double (&v)[SL * NC(LC) * NC(LD)],
                                             const SpanRows &R, double (&jreg)[JREG ? NC(LA) * NC(LB) : 1], int lane)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD), NCD = NCc * ND, NX = NA + NB;
    constexpr bool KEEP_B = NB * (NCc + ND) <= QBX_DIGEST_KEEP_B;
    constexpr bool KEEP_DCD = NCD <= 18;
    const int N = p.nbf, W = R.W;
    const double *__restrict__ DJ = p.DJ;
    for (int m = 0; m < p.nmat; ++m) {
        const double *dA = R.d + (int64_t)m * NX * W, *dB = dA + NA * W;
        double *kA = R.k + (int64_t)m * NX * W, *kB = kA + NA * W;
        const bool coul = (m == 0);
        double dcd[KEEP_DCD ? NCD : 1], jcd[NCD];
        double dbc[KEEP_B ? NB * NCc : NCc], dbd[KEEP_B ? NB * ND : ND], kbc[KEEP_B ? NB * NCc : NCc], kbd[KEEP_B ? NB * ND : ND];
#pragma unroll
        for (int cd = 0; cd < NCD; ++cd) jcd[cd] = 0.0;
        if (KEEP_DCD) {
#pragma unroll
            for (int c = 0; c < NCc; ++c)
#pragma unroll
                for (int d = 0; d < ND; ++d) dcd[c * ND + d] = coul ? __ldg(DJ + (id + d) + N * (ic + c)) : 0.0;
        }
        if (KEEP_B) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
                for (int c = 0; c < NCc; ++c) { dbc[b * NCc + c] = dB[b * W + jc + c]; kbc[b * NCc + c] = 0.0; }
#pragma unroll
                for (int d = 0; d < ND; ++d) { dbd[b * ND + d] = dB[b * W + jd + d]; kbd[b * ND + d] = 0.0; }
            }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            double dac[NCc], dad[ND], kac[NCc], kad[ND];
#pragma unroll
            for (int c = 0; c < NCc; ++c) { dac[c] = dA[a * W + jc + c]; kac[c] = 0.0; }
#pragma unroll
            for (int d = 0; d < ND; ++d) { dad[d] = dA[a * W + jd + d]; kad[d] = 0.0; }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int ab = a * NB + b, s0 = (ab % SL) * NCD;
                if (ab % SL == 0 && (ab > 0 || m > 0)) {              // next slab of values
#pragma unroll
                    for (int i = 0; i < SL * NCD; ++i) v[i] = __ldg(vq + (int64_t)(ab * NCD + i) * nt);
                }
                const int ob = KEEP_B ? b : 0;
                if (!KEEP_B) {
#pragma unroll
                    for (int c = 0; c < NCc; ++c) { dbc[c] = dB[b * W + jc + c]; kbc[c] = 0.0; }
#pragma unroll
                    for (int d = 0; d < ND; ++d) { dbd[d] = dB[b * W + jd + d]; kbd[d] = 0.0; }
                }
                const double dab = coul ? __ldg(DJ + (ib + b) + N * (ia + a)) : 0.0;
                double j = 0.0;
#pragma unroll
                for (int c = 0; c < NCc; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double x = v[s0 + c * ND + d];
                        const double dd = KEEP_DCD ? dcd[c * ND + d] : (coul ? __ldg(DJ + (id + d) + N * (ic + c)) : 0.0);
                        j = fma(dd, x, j);
                        jcd[c * ND + d] = fma(dab, x, jcd[c * ND + d]);
                        kac[c] = fma(dbd[ob * ND + d], x, kac[c]);
                        kad[d] = fma(dbc[ob * NCc + c], x, kad[d]);
                        kbc[ob * NCc + c] = fma(dad[d], x, kbc[ob * NCc + c]);
                        kbd[ob * ND + d] = fma(dac[c], x, kbd[ob * ND + d]);
                    }
                if (coul) {                                   // J[ab]: a sum over the whole bra row
                    j *= 2.0 * f;
                    if (JREG) jreg[JREG ? ab : 0] += j;
                    else {
                        j = warp_sum(j);
                        if (lane == 0) R.jab[ab] += j;
                    }
                }
                if (!KEEP_B) span_flush_row<NCc, ND>(kB + b * W, jc, jd, kbc, kbd, f, L);
            }
            span_flush_row<NCc, ND>(kA + a * W, jc, jd, kac, kad, f, L);
        }
        if (KEEP_B) {
#pragma unroll
            for (int b = 0; b < NB; ++b) span_flush_row<NCc, ND>(kB + b * W, jc, jd, kbc + b * NCc, kbd + b * ND, f, L);
        }
        if (coul && L.mine) {                                 // J[cd]: the one global update per ket component
#pragma unroll
            for (int c = 0; c < NCc; ++c)
#pragma unroll
                for (int d = 0; d < ND; ++d) atomicAdd(p.Jt + (id + d) + N * (ic + c), 2.0 * f * jcd[c * ND + d]);
        }
    }
}

// shared memory per warp, in doubles: K rows and density rows of both bra shells over the W cached columns, for every
// exchange density, + J[ab]
template <int LA, int LB>
__host__ __device__ constexpr int digest_span_doubles(int W, int nmat) { return 2 * nmat * (NC(LA) + NC(LB)) * W + ((NC(LA) * NC(LB) + 1) & ~1); }

// Occupancy is set by the shared-memory rows; the register budget follows the bra class (ss: 64, ps: 128, else 255)
#define QBX_SPAN_MAX_WARPS 8
__host__ __device__ constexpr int digest_span_min_blocks(int la, int lb) { return la + lb == 0 ? 4 : (la + lb == 1 ? 2 : 1); }

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(32 * QBX_SPAN_MAX_WARPS, digest_span_min_blocks(LA, LB)) digest_span_kernel(DigestArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCD = NC(LC) * NC(LD), NAB = NA * NB, NX = NA + NB;
    constexpr int SL = (NAB * NCD <= QBX_DIGEST_SLAB) ? NAB : ((NB * NCD <= QBX_DIGEST_SLAB) ? NB : 1);
    constexpr int NV = SL * NCD;
    constexpr bool PFV = NV <= QBX_SPAN_PREFETCH_VALUES;       // next tile's values in flight while this one is digested
    constexpr bool JREG = NAB <= 9;
    extern __shared__ double span_smem[];
    const int N = p.nbf, nm = p.nmat;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int W = p.wC + p.wD;                                 // cached columns: [c0, c0 + wC) then [d0, d0 + wD)
    const int rowd = digest_span_doubles<LA, LB>(W, nm), nrow = nm * NX * W;
    SpanRows R;
    R.W = W;
    double *sk = span_smem + (size_t)warp * rowd, *sd = sk + nrow;
    R.k = sk; R.d = sd; R.jab = sd + nrow;
    for (int i = lane; i < rowd; i += 32) sk[i] = 0.0;
    __syncwarp();
    double jreg[JREG ? NAB : 1];
#pragma unroll
    for (int i = 0; i < (JREG ? NAB : 1); ++i) jreg[i] = 0.0;
    const int64_t span = p.span, nspan = (p.ntasks + span - 1) / span, nt = p.ntasks;
    const int64_t nwarp = (int64_t)gridDim.x * wpb;
    auto gcol = [&](int j) { return j < p.wC ? p.c0 + j : p.d0 + (j - p.wC); };   // local column -> internal function

    for (int64_t sp = (int64_t)blockIdx.x * wpb + warp; sp < nspan; sp += nwarp) {
        const int64_t s0 = sp * span, s1 = s0 + span < nt ? s0 + span : nt;
        int cur = -1, ia = 0, ib = 0;                          // bra pair whose rows are resident, its first functions
        double fab = 1.0;
        auto flush = [&]() {                                   // non-zero entries of the resident K rows -> Kt, Jt; rows zeroed
            if (cur < 0) return;
            __syncwarp();
            for (int m = 0; m < nm; ++m)
                for (int x = 0; x < NX; ++x) {
                    double *row = sk + (int64_t)(m * NX + x) * W;
                    double *g = p.Kt + (int64_t)m * N * N + (int64_t)N * (x < NA ? ia + x : ib + x - NA);
                    for (int j = lane; j < W; j += 32) {
                        const double y = row[j];
                        if (y != 0.0) { atomicAdd(g + gcol(j), y); row[j] = 0.0; }
                    }
                }
            if (JREG) {
#pragma unroll
                for (int i = 0; i < NAB; ++i) {
                    const double y = warp_sum(jreg[JREG ? i : 0]);
                    if (lane == 0 && y != 0.0) atomicAdd(p.Jt + (ib + i % NB) + (int64_t)N * (ia + i / NB), y);
                    jreg[JREG ? i : 0] = 0.0;
                }
            } else {
                for (int i = lane; i < NAB; i += 32) {
                    const double y = R.jab[i];
                    if (y != 0.0) { atomicAdd(p.Jt + (ib + i % NB) + (int64_t)N * (ia + i / NB), y); R.jab[i] = 0.0; }
                }
            }
            __syncwarp();
        };
        auto load_rows = [&]() {                               // exchange-density rows of the new bra pair
            for (int m = 0; m < nm; ++m)
                for (int x = 0; x < NX; ++x) {
                    double *row = sd + (int64_t)(m * NX + x) * W;
                    const double *g = p.DK + (int64_t)m * N * N + (int64_t)N * (x < NA ? ia + x : ib + x - NA);
                    for (int j = lane; j < W; j += 32) row[j] = __ldg(g + gcol(j));
                }
            __syncwarp();
        };
        auto task_at = [&](int64_t tile) { const int64_t q = tile + lane; return __ldg(p.tasks + (q < s1 ? q : s1 - 1)); };
        int2 t = task_at(s0);
        int2 tn = s0 + 32 < s1 ? task_at(s0 + 32) : t;
        int4 rk = __ldg(p.ket_info + (t.y < 0 ? 0 : t.y));
        double v[NV], vn[PFV ? NV : 1];
        if (PFV) {
            const double *vq = p.vals + (s0 + lane < s1 ? s0 + lane : s1 - 1);
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = __ldg(vq + (int64_t)i * nt);
        }
        for (int64_t tile = s0; tile < s1; tile += 32) {
            const int64_t q0 = tile + lane, q = q0 < s1 ? q0 : s1 - 1;
            const double *__restrict__ vq = p.vals + q;
            // loads for the tiles ahead: task record of tile + 2, ket info (and values) of tile + 1
            const int2 tn2 = tile + 64 < s1 ? task_at(tile + 64) : tn;
            const int4 rkn = __ldg(p.ket_info + (tn.y < 0 ? 0 : tn.y));
            if (PFV) {
                if (tile + 32 < s1) {
                    const double *vqn = p.vals + (q0 + 32 < s1 ? q0 + 32 : s1 - 1);
#pragma unroll
                    for (int i = 0; i < NV; ++i) vn[PFV ? i : 0] = __ldg(vqn + (int64_t)i * nt);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NV; ++i) v[i] = __ldg(vq + (int64_t)i * nt);
            }
            const bool valid = q0 < s1 && t.y >= 0;            // ket = -1: unused slot of a group task
            unsigned todo = __ballot_sync(0xffffffffu, valid);
            int round = 0;
            while (todo) {                                     // one round per bra pair in the tile (almost always one)
                const int bra = __shfl_sync(0xffffffffu, t.x, __ffs(todo) - 1);
                SpanLanes L;
                L.mine = valid && t.x == bra;
                todo &= ~__ballot_sync(0xffffffffu, L.mine);
                if (bra != cur) {
                    flush();
                    const int4 rb = __ldg(p.bra_info + bra);
                    cur = bra; ia = rb.z; ib = rb.w; fab = rb.x == rb.y ? 0.5 : 1.0;
                    load_rows();
                }
                double f = L.mine ? fab : 0.0;
                if (rk.x == rk.y) f *= 0.5;
                if (p.same_class && t.x == t.y) f *= 0.5;
                // lanes that hold the same shell C (or D) as another lane of the tile would collide in the shared-memory
                // rows: they take turns (rare with the diagonal pair order)
                const unsigned below = (1u << lane) - 1u;
                L.turnC = __popc(__match_any_sync(0xffffffffu, L.mine ? rk.x : -1 - lane) & below);
                L.turnD = __popc(__match_any_sync(0xffffffffu, L.mine ? rk.y : -1 - lane) & below);
                L.nturnC = L.nturnD = 1;
                while (__any_sync(0xffffffffu, L.turnC >= L.nturnC)) ++L.nturnC;
                while (__any_sync(0xffffffffu, L.turnD >= L.nturnD)) ++L.nturnD;
                if (round > 0 && NV < NAB * NCD) {             // a later round: slab 0 again
#pragma unroll
                    for (int i = 0; i < NV; ++i) v[i] = __ldg(vq + (int64_t)i * nt);
                }
                const int jc = rk.z - p.c0, jd = p.wD ? p.wC + (rk.w - p.d0) : rk.w - p.c0;
                span_quartet<LA, LB, LC, LD, SL, JREG>(p, vq, nt, f, ia, ib, rk.z, rk.w, jc, jd, L, v, R, jreg, lane);
                ++round;
            }
            t = tn; tn = tn2; rk = rkn;
            if (PFV) {
#pragma unroll
                for (int i = 0; i < NV; ++i) v[i] = vn[PFV ? i : 0];
            }
        }
        flush();
    }
}

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(128) scatter_kernel(ScatterArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD);
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= p.ntasks) return;
    const int2 t = p.tasks[q];
    if (t.y < 0) return;
    const int2 sb = p.bra_shells[t.x], sk = p.ket_shells[t.y];
    const int64_t N = p.nbf;
    double *T = p.tensor;
    for (int a = 0; a < NA; ++a)
        for (int b = 0; b < NB; ++b)
            for (int c = 0; c < NCc; ++c)
                for (int d = 0; d < ND; ++d) {
                    const int64_t i = p.shell_bf[6 * sb.x + a], j = p.shell_bf[6 * sb.y + b];
                    const int64_t k = p.shell_bf[6 * sk.x + c], l = p.shell_bf[6 * sk.y + d];
                    if (i < 0 || j < 0 || k < 0 || l < 0) continue;
                    const double v = p.vals[(int64_t)(((a * NB + b) * NCc + c) * ND + d) * p.ntasks + q];
#define AT(w, x, y, z) T[(w) + N * ((x) + N * ((y) + N * (z)))]
                    AT(i, j, k, l) = v; AT(j, i, k, l) = v; AT(i, j, l, k) = v; AT(j, i, l, k) = v;
                    AT(k, l, i, j) = v; AT(k, l, j, i) = v; AT(l, k, i, j) = v; AT(l, k, j, i) = v;
#undef AT
                }
}
