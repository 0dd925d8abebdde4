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
//     Jt[a,b] += 2 f DJ[c,d] v     Jt[c,d] += 2 f DJ[a,b] v
//     Kt[a,d] += f DK[b,c] v   Kt[b,d] += f DK[a,c] v   Kt[a,c] += f DK[b,d] v   Kt[b,c] += f DK[a,d] v
// and G = (Jt + Jt^T) - (Kt + Kt^T) (k_finish_G in engine.cu).  One thread owns one quartet,
// reduces its block into per-thread partial sums first and only then issues atomics, so a
// block of n_a n_b n_c n_d values costs n_a n_b + n_c n_d + (n_a + n_b)(n_c + n_d) atomics.
#pragma once
#include "engine.h"

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Lanes of a warp that update the same G block (same ket shell C or D) are summed first:
// 32 same-address atomics in one instruction serialise in L2, one atomic per group does not.
struct KeyGroup {
    unsigned peers;
    int rank, steps;
    int src[5];
};
__device__ __forceinline__ KeyGroup make_group(int key)
{
    const int lane = threadIdx.x & 31;
    KeyGroup g;
    g.peers = __match_any_sync(0xffffffffu, key);
    g.rank = __popc(g.peers & ((1u << lane) - 1u));
    int size = __popc(g.peers);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) size = max(size, __shfl_xor_sync(0xffffffffu, size, o));
    g.steps = size > 1 ? 32 - __clz(size - 1) : 0;             // warp-uniform tree depth
#pragma unroll
    for (int k = 0; k < 5; ++k) g.src[k] = (int)__fns(g.peers, lane, (1 << k) + 1);   // (1<<k)-th peer above me or -1
    return g;
}
__device__ __forceinline__ double group_sum(double v, const KeyGroup &g)
{
    for (int k = 0; k < g.steps; ++k) {
        const double up = __shfl_sync(0xffffffffu, v, g.src[k] & 31);
        if (g.src[k] >= 0 && (g.rank & ((2 << k) - 1)) == 0) v += up;
    }
    return v;                                                   // total lives in the rank-0 lane
}

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(128) digest_kernel(DigestArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD);
    const int64_t q0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q0 - (threadIdx.x & 31) >= p.ntasks) return;          // whole warp past the end
    const bool valid = q0 < p.ntasks;
    const int64_t q = valid ? q0 : p.ntasks - 1;
    const int2 t = p.tasks[q];
    const int2 sb = p.bra_shells[t.x], sk = p.ket_shells[t.y];
    double f = valid ? 1.0 : 0.0;
    if (sb.x == sb.y) f *= 0.5;
    if (sk.x == sk.y) f *= 0.5;
    if (p.same_class && t.x == t.y) f *= 0.5;
    // tasks are bra-major, so a warp nearly always shares its bra pair: J_AB is then reduced
    // over the warp and added once (the same-address atomics were the bottleneck otherwise)
    const bool uni = __all_sync(0xffffffffu, t.x == __shfl_sync(0xffffffffu, t.x, 0));
    int fa[NA], fb[NB], fc[NCc], fd[ND];
#pragma unroll
    for (int i = 0; i < NA; ++i) fa[i] = p.shell_bf[6 * sb.x + i];
#pragma unroll
    for (int i = 0; i < NB; ++i) fb[i] = p.shell_bf[6 * sb.y + i];
#pragma unroll
    for (int i = 0; i < NCc; ++i) fc[i] = p.shell_bf[6 * sk.x + i];
#pragma unroll
    for (int i = 0; i < ND; ++i) fd[i] = p.shell_bf[6 * sk.y + i];
    const int64_t N = p.nbf, N2 = N * N;
    const double *vq = p.vals + q;

    // Coulomb part
    double jcd[NCc * ND];
#pragma unroll
    for (int i = 0; i < NCc * ND; ++i) jcd[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const bool okab = fa[a] >= 0 && fb[b] >= 0;
            const double dab = okab ? p.DJ[fa[a] + N * fb[b]] : 0.0;
            double jab = 0.0;
#pragma unroll
            for (int c = 0; c < NCc; ++c)
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    if (!okab || fc[c] < 0 || fd[d] < 0) continue;
                    const double v = vq[(int64_t)(((a * NB + b) * NCc + c) * ND + d) * p.ntasks];
                    jab = fma(p.DJ[fc[c] + N * fd[d]], v, jab);
                    jcd[c * ND + d] = fma(dab, v, jcd[c * ND + d]);
                }
            jab *= 2.0 * f;
            if (uni) {
                jab = warp_sum(jab);
                if ((threadIdx.x & 31) == 0 && okab) atomicAdd(p.Jt + fa[a] + N * fb[b], jab);
            } else if (okab && valid) {
                atomicAdd(p.Jt + fa[a] + N * fb[b], jab);
            }
        }
    if (valid) {
#pragma unroll
        for (int c = 0; c < NCc; ++c)
#pragma unroll
            for (int d = 0; d < ND; ++d)
                if (fc[c] >= 0 && fd[d] >= 0) atomicAdd(p.Jt + fc[c] + N * fd[d], 2.0 * f * jcd[c * ND + d]);
    }
    // exchange blocks K[A,C], K[B,C] are shared by the lanes with the same C (given the common
    // bra), K[A,D], K[B,D] by those with the same D; without a common bra nothing is merged
    const int lane = threadIdx.x & 31;
    const KeyGroup gC = make_group(uni && valid ? sk.x : -1 - lane);
    const KeyGroup gD = make_group(uni && valid ? sk.y : -1 - lane);

    // exchange part, one density at a time
    for (int m = 0; m < p.nmat; ++m) {
        const double *DK = p.DK + m * N2;
        double *Kt = p.Kt + m * N2;
        double kac[NA * NCc], kad[NA * ND], kbc[NB * NCc], kbd[NB * ND];
#pragma unroll
        for (int i = 0; i < NA * NCc; ++i) kac[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NA * ND; ++i) kad[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB * NCc; ++i) kbc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB * ND; ++i) kbd[i] = 0.0;
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (fa[a] < 0 || fb[b] < 0) continue;
#pragma unroll
                for (int c = 0; c < NCc; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        if (fc[c] < 0 || fd[d] < 0) continue;
                        const double v = vq[(int64_t)(((a * NB + b) * NCc + c) * ND + d) * p.ntasks];
                        kac[a * NCc + c] = fma(DK[fb[b] + N * fd[d]], v, kac[a * NCc + c]);
                        kad[a * ND + d] = fma(DK[fb[b] + N * fc[c]], v, kad[a * ND + d]);
                        kbc[b * NCc + c] = fma(DK[fa[a] + N * fd[d]], v, kbc[b * NCc + c]);
                        kbd[b * ND + d] = fma(DK[fa[a] + N * fc[c]], v, kbd[b * ND + d]);
                    }
            }
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int c = 0; c < NCc; ++c) {
                const double v = group_sum(f * kac[a * NCc + c], gC);
                if (valid && gC.rank == 0 && fa[a] >= 0 && fc[c] >= 0) atomicAdd(Kt + fa[a] + N * fc[c], v);
            }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int c = 0; c < NCc; ++c) {
                const double v = group_sum(f * kbc[b * NCc + c], gC);
                if (valid && gC.rank == 0 && fb[b] >= 0 && fc[c] >= 0) atomicAdd(Kt + fb[b] + N * fc[c], v);
            }
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double v = group_sum(f * kad[a * ND + d], gD);
                if (valid && gD.rank == 0 && fa[a] >= 0 && fd[d] >= 0) atomicAdd(Kt + fa[a] + N * fd[d], v);
            }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double v = group_sum(f * kbd[b * ND + d], gD);
                if (valid && gD.rank == 0 && fb[b] >= 0 && fd[d] >= 0) atomicAdd(Kt + fb[b] + N * fd[d], v);
            }
    }
}

template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(128) scatter_kernel(ScatterArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD);
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= p.ntasks) return;
    const int2 t = p.tasks[q];
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
