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

// -----------------------------------------------------------------------------------------
// Row-block digestion (stored mode).  One block owns a SEGMENT of the task list: consecutive
// quartets that share the bra pair (A,B).  That turns almost every scattered global access of
// digest_kernel into shared-memory or coalesced traffic:
//   * K: the density rows D[a,:], D[b,:] of the bra shells are staged in shared memory, and the
//     exchange blocks K[a,:], K[b,:] are accumulated in per-warp private shared-memory rows
//     (lanes sharing a ket shell are summed first, so each phase writes distinct addresses),
//     reduced over the warps and flushed once per segment with coalesced REDs (into the
//     transposed element, which is equivalent because G = Kt + Kt^T);
//   * J: densities and results are "pair vectors" indexed [component][pair], so a warp's 32
//     consecutive kets read dket[..][t.y] and update jket[..][t.y] at consecutive addresses,
//     and J_AB is a block-uniform address (warp-reduced first).
// Global traffic left per quartet: its values (the HBM stream), one task, one ket shell pair.
// -----------------------------------------------------------------------------------------
template <int LA, int LB, int LC, int LD, int W>
__global__ void __launch_bounds__(W * 32) digest2_kernel(Digest2Args p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD), NR = NA + NB;
    extern __shared__ double sm[];
    const int N = p.nbf;
    double *Drow = sm;                                   // [NR][N]
    double *Kall = sm + NR * N;                          // [W][NR][N]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *Kp = Kall + (size_t)wib * NR * N;
    for (int seg = blockIdx.x; seg < p.nsegs; seg += gridDim.x) {
        const int2 sg = p.segs[seg];
        const int ib = p.tasks[sg.x].x;
        const int2 sb = p.bra_shells[ib];
        int fa[NA], fb[NB];
#pragma unroll
        for (int i = 0; i < NA; ++i) fa[i] = p.shell_bf[6 * sb.x + i];
#pragma unroll
        for (int i = 0; i < NB; ++i) fb[i] = p.shell_bf[6 * sb.y + i];
        for (int m = 0; m < p.nmat; ++m) {
            const double *DK = p.DK + (size_t)m * N * N;
            double *Kt = p.Kt + (size_t)m * N * N;
            for (int idx = threadIdx.x; idx < NR * N; idx += W * 32) {
                const int r = idx / N, n = idx - r * N;
                int f = -1;
#pragma unroll
                for (int i = 0; i < NA; ++i) if (r == i) f = fa[i];
#pragma unroll
                for (int i = 0; i < NB; ++i) if (r == NA + i) f = fb[i];
                Drow[idx] = f >= 0 ? DK[n + (size_t)N * f] : 0.0;       // D is symmetric: row = column
            }
            for (int idx = threadIdx.x; idx < W * NR * N; idx += W * 32) Kall[idx] = 0.0;
            __syncthreads();
            for (int base = 0; base < sg.y; base += W * 32) {
                const bool valid = base + (int)threadIdx.x < sg.y;
                const int64_t q = sg.x + (valid ? base + (int)threadIdx.x : 0);
                const int2 t = p.tasks[q];
                const int2 sk = p.ket_shells[t.y];
                double f = valid ? 1.0 : 0.0;
                if (sb.x == sb.y) f *= 0.5;
                if (sk.x == sk.y) f *= 0.5;
                if (p.same_class && t.x == t.y) f *= 0.5;
                int fc[NCc], fd[ND];
#pragma unroll
                for (int i = 0; i < NCc; ++i) fc[i] = p.shell_bf[6 * sk.x + i];
#pragma unroll
                for (int i = 0; i < ND; ++i) fd[i] = p.shell_bf[6 * sk.y + i];
                const double *vq = p.vals + q;
                if (m == 0) {                                            // Coulomb part, pair vectors
                    double jcd[NCc * ND];
#pragma unroll
                    for (int i = 0; i < NCc * ND; ++i) jcd[i] = 0.0;
                    double dk[NCc * ND];
#pragma unroll
                    for (int i = 0; i < NCc * ND; ++i) dk[i] = p.dket[(size_t)i * p.nket + t.y];
#pragma unroll
                    for (int ab = 0; ab < NA * NB; ++ab) {
                        const double dab = p.dbra[(size_t)ab * p.nbra + ib];
                        double jab = 0.0;
#pragma unroll
                        for (int cd = 0; cd < NCc * ND; ++cd) {
                            const double v = vq[(int64_t)(ab * NCc * ND + cd) * p.ntasks];
                            jab = fma(dk[cd], v, jab);
                            jcd[cd] = fma(dab, v, jcd[cd]);
                        }
                        jab = warp_sum(2.0 * f * jab);
                        if (lane == 0) atomicAdd(p.jbra + (size_t)ab * p.nbra + ib, jab);
                    }
                    if (valid) {
#pragma unroll
                        for (int cd = 0; cd < NCc * ND; ++cd) atomicAdd(p.jket + (size_t)cd * p.nket + t.y, 2.0 * f * jcd[cd]);
                    }
                }
                // exchange part: per-thread block sums with the staged density rows
                double kac[NA * NCc], kad[NA * ND], kbc[NB * NCc], kbd[NB * ND];
#pragma unroll
                for (int i = 0; i < NA * NCc; ++i) kac[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NA * ND; ++i) kad[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NB * NCc; ++i) kbc[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NB * ND; ++i) kbd[i] = 0.0;
                double dac[NA * NCc], dad[NA * ND], dbc[NB * NCc], dbd[NB * ND];
#pragma unroll
                for (int a = 0; a < NA; ++a) {
#pragma unroll
                    for (int c = 0; c < NCc; ++c) dac[a * NCc + c] = fc[c] >= 0 ? Drow[a * N + fc[c]] : 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) dad[a * ND + d] = fd[d] >= 0 ? Drow[a * N + fd[d]] : 0.0;
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
#pragma unroll
                    for (int c = 0; c < NCc; ++c) dbc[b * NCc + c] = fc[c] >= 0 ? Drow[(NA + b) * N + fc[c]] : 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) dbd[b * ND + d] = fd[d] >= 0 ? Drow[(NA + b) * N + fd[d]] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int c = 0; c < NCc; ++c)
#pragma unroll
                            for (int d = 0; d < ND; ++d) {
                                const double v = f * vq[(int64_t)(((a * NB + b) * NCc + c) * ND + d) * p.ntasks];
                                kac[a * NCc + c] = fma(dbd[b * ND + d], v, kac[a * NCc + c]);
                                kad[a * ND + d] = fma(dbc[b * NCc + c], v, kad[a * ND + d]);
                                kbc[b * NCc + c] = fma(dad[a * ND + d], v, kbc[b * NCc + c]);
                                kbd[b * ND + d] = fma(dac[a * NCc + c], v, kbd[b * ND + d]);
                            }
                const KeyGroup gC = make_group(valid ? sk.x : -1 - lane);
                const KeyGroup gD = make_group(valid ? sk.y : -1 - lane);
                // phase C: leaders of the C groups own distinct columns fc[.]
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int c = 0; c < NCc; ++c) {
                        const double v = group_sum(kac[a * NCc + c], gC);
                        if (valid && gC.rank == 0 && fc[c] >= 0) Kp[a * N + fc[c]] += v;
                    }
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int c = 0; c < NCc; ++c) {
                        const double v = group_sum(kbc[b * NCc + c], gC);
                        if (valid && gC.rank == 0 && fc[c] >= 0) Kp[(NA + b) * N + fc[c]] += v;
                    }
                __syncwarp();
                // phase D
#pragma unroll
                for (int a = 0; a < NA; ++a)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double v = group_sum(kad[a * ND + d], gD);
                        if (valid && gD.rank == 0 && fd[d] >= 0) Kp[a * N + fd[d]] += v;
                    }
#pragma unroll
                for (int b = 0; b < NB; ++b)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double v = group_sum(kbd[b * ND + d], gD);
                        if (valid && gD.rank == 0 && fd[d] >= 0) Kp[(NA + b) * N + fd[d]] += v;
                    }
                __syncwarp();
            }
            __syncthreads();
            for (int idx = threadIdx.x; idx < NR * N; idx += W * 32) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < W; ++w) tot += Kall[(size_t)w * NR * N + idx];
                if (tot != 0.0) {
                    const int r = idx / N, n = idx - r * N;
                    int fr = -1;
#pragma unroll
                    for (int i = 0; i < NA; ++i) if (r == i) fr = fa[i];
#pragma unroll
                    for (int i = 0; i < NB; ++i) if (r == NA + i) fr = fb[i];
                    if (fr >= 0) atomicAdd(Kt + n + (size_t)N * fr, tot);
                }
            }
            __syncthreads();
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
