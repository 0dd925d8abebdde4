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
//
// Round 2 tried five restructurings of this kernel on the B200 -- K rows of the bra pair in shared memory with
// shuffle sums, shell pairs listed diagonal by diagonal with turn-taking, U consecutive quartets per lane with
// shared-memory CAS-adds, and a run table with per-run butterflies -- and every one of them lost against this one
// (17.5 / 22 / 25 / 46 ms against 11-12 ms for the (H2O)16 build).  The numbers, the ncu counters and the reason --
// the list's runs of equal shell C are 5-11 quartets long, not tens, because the pair lists are sorted by
// primitive count for the ERI kernels -- are in profiles/r02/digest_history.md.  What did win in round 2: the value slab
// and kept-row sizes below by A/B (12.2 -> 11.4 ms), and a kernel of its own for the general-contraction classes
// (ss|ss), (ps|ss), one lane per group task (eri_group.cu: digest_group_kernel; 11.3 -> 10.9 ms).
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
// (A/B on a B200, tools/gpu_ab_digest.sh, Fock build of (H2O)16: slab 64 / keep 18 -- the round-1 guess -- 12.2 ms;
// slab 32 11.9, slab 16 12.0, slab 128 12.6; keep 0 14.6, keep 36 11.6, keep 54 11.6; slab 32 + keep 36 11.4 ms.
// One register tier up or down in digest_min_blocks: 13.9 / 11.85 ms, two down 13.0.)
#ifndef QBX_DIGEST_SLAB
#define QBX_DIGEST_SLAB 32
#endif
#ifndef QBX_DIGEST_KEEP_B
#define QBX_DIGEST_KEEP_B 36
#endif

// resident blocks per SM the register allocation aims at (128 threads each), by components per quartet
// (A/B builds: QBX_NVCC_DEFS="-DQBX_DIGEST_MB2=6"; tools/gpu_ab_digest.sh)
#ifndef QBX_DIGEST_MB0
#define QBX_DIGEST_MB0 10
#define QBX_DIGEST_MB1 6
#define QBX_DIGEST_MB2 4
#define QBX_DIGEST_MB3 3
#define QBX_DIGEST_MB4 2
#endif
__host__ __device__ constexpr int digest_min_blocks(int ncomp)
{
    return ncomp <= 3 ? QBX_DIGEST_MB0 : (ncomp <= 9 ? QBX_DIGEST_MB1 : (ncomp <= 36 ? QBX_DIGEST_MB2 : (ncomp <= 162 ? QBX_DIGEST_MB3 : QBX_DIGEST_MB4)));
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
