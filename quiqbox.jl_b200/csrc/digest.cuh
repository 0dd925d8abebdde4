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

#define QBX_DIGEST_SPREAD 384

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Latency notes (ncu, profiles/r01): the kernel is bound by exposed memory latency, not by
// any throughput, so the body is organised as few dependent steps as possible:
//   task -> one int4 record per pair (shells + first function indices) -> all densities and
//   the values of one (a,b) slice as independent loads -> FMAs -> all shuffles/REDs at the end.
// Values are read once and feed J and K together (the second exchange density of a UHF build
// re-reads them).
template <int LA, int LB, int LC, int LD>
__global__ void __launch_bounds__(128) digest_kernel(DigestArgs p)
{
    constexpr int NA = NC(LA), NB = NC(LB), NCc = NC(LC), ND = NC(LD), NCD = NCc * ND;
    // Block order is transposed against task order: the blocks resident at one time sit
    // spread-th of the list apart, i.e. on different bra rows.  In task order they would
    // all update the same few hundred G elements, and fp64 REDs on a hot set that small run at
    // 1e10/s instead of 2e11/s on B200 (tools/redbench.cu; profiles/r01/redbench.md).
    const int64_t nblk = (p.ntasks + 127) / 128;
    const int64_t R = nblk < p.spread ? nblk : p.spread, C = (nblk + R - 1) / R;
    const int64_t blk = (blockIdx.x % R) * C + blockIdx.x / R;
    if (blk >= nblk) return;
    const int64_t q0 = blk * 128 + threadIdx.x;
    if (q0 - (threadIdx.x & 31) >= p.ntasks) return;          // whole warp past the end
    const int64_t q = q0 < p.ntasks ? q0 : p.ntasks - 1;
    int2 t = p.tasks[q];
    const bool valid = q0 < p.ntasks && t.y >= 0;              // ket = -1: unused slot of a group task
    if (t.y < 0) t.y = 0;
    const int4 rb = __ldg(p.bra_info + t.x), rk = __ldg(p.ket_info + t.y);   // (shell A, shell B, first A, first B)
    double f = valid ? 1.0 : 0.0;
    if (rb.x == rb.y) f *= 0.5;
    if (rk.x == rk.y) f *= 0.5;
    if (p.same_class && t.x == t.y) f *= 0.5;
    const bool uniAB = __all_sync(0xffffffffu, t.x == __shfl_sync(0xffffffffu, t.x, 0));
    const bool uniC = uniAB && __all_sync(0xffffffffu, rk.x == __shfl_sync(0xffffffffu, rk.x, 0));
    const bool lane0 = (threadIdx.x & 31) == 0;
    const int64_t N = p.nbf;                                  // internal dimension
    const int ia = rb.z, ib = rb.w, ic = rk.z, id = rk.w;
    const double *vq = p.vals + q;

    for (int m = 0; m < p.nmat; ++m) {
        const double *DK = p.DK + m * N * N;
        double *Kt = p.Kt + m * N * N;
        const bool coul = (m == 0);
        double dcd[NCD], jcd[NCD], jab[NA * NB];
        double dac[NA * NCc], dad[NA * ND], dbc[NB * NCc], dbd[NB * ND];
        double kac[NA * NCc], kad[NA * ND], kbc[NB * NCc], kbd[NB * ND];
#pragma unroll
        for (int c = 0; c < NCc; ++c)
#pragma unroll
            for (int d = 0; d < ND; ++d) { dcd[c * ND + d] = coul ? p.DJ[(id + d) + N * (ic + c)] : 0.0; jcd[c * ND + d] = 0.0; }
#pragma unroll
        for (int a = 0; a < NA; ++a) {
#pragma unroll
            for (int c = 0; c < NCc; ++c) { dac[a * NCc + c] = DK[(ic + c) + N * (ia + a)]; kac[a * NCc + c] = 0.0; }
#pragma unroll
            for (int d = 0; d < ND; ++d) { dad[a * ND + d] = DK[(id + d) + N * (ia + a)]; kad[a * ND + d] = 0.0; }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int c = 0; c < NCc; ++c) { dbc[b * NCc + c] = DK[(ic + c) + N * (ib + b)]; kbc[b * NCc + c] = 0.0; }
#pragma unroll
            for (int d = 0; d < ND; ++d) { dbd[b * ND + d] = DK[(id + d) + N * (ib + b)]; kbd[b * ND + d] = 0.0; }
        }
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const double dab = coul ? p.DJ[(ib + b) + N * (ia + a)] : 0.0;
                double v[NCD];
#pragma unroll
                for (int cd = 0; cd < NCD; ++cd) v[cd] = f * vq[(int64_t)((a * NB + b) * NCD + cd) * p.ntasks];
                double j = 0.0;
#pragma unroll
                for (int c = 0; c < NCc; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double x = v[c * ND + d];
                        j = fma(dcd[c * ND + d], x, j);
                        jcd[c * ND + d] = fma(dab, x, jcd[c * ND + d]);
                        kac[a * NCc + c] = fma(dbd[b * ND + d], x, kac[a * NCc + c]);
                        kad[a * ND + d] = fma(dbc[b * NCc + c], x, kad[a * ND + d]);
                        kbc[b * NCc + c] = fma(dad[a * ND + d], x, kbc[b * NCc + c]);
                        kbd[b * ND + d] = fma(dac[a * NCc + c], x, kbd[b * ND + d]);
                    }
                jab[a * NB + b] = j;
            }
        // ---- updates (f is already folded into the values)
        if (coul) {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    double j = 2.0 * jab[a * NB + b];
                    if (uniAB) {
                        j = warp_sum(j);
                        if (lane0) atomicAdd(p.Jt + (ib + b) + N * (ia + a), j);
                    } else if (valid) {
                        atomicAdd(p.Jt + (ib + b) + N * (ia + a), j);
                    }
                }
            if (valid) {
#pragma unroll
                for (int c = 0; c < NCc; ++c)
#pragma unroll
                    for (int d = 0; d < ND; ++d) atomicAdd(p.Jt + (id + d) + N * (ic + c), 2.0 * jcd[c * ND + d]);
            }
        }
        if (uniC) {                                           // K[ac], K[bc]: one address per warp
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NCc; ++c) {
                    const double x = warp_sum(kac[a * NCc + c]);
                    if (lane0) atomicAdd(Kt + (ic + c) + N * (ia + a), x);
                }
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int c = 0; c < NCc; ++c) {
                    const double x = warp_sum(kbc[b * NCc + c]);
                    if (lane0) atomicAdd(Kt + (ic + c) + N * (ib + b), x);
                }
        } else if (valid) {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int c = 0; c < NCc; ++c) atomicAdd(Kt + (ic + c) + N * (ia + a), kac[a * NCc + c]);
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int c = 0; c < NCc; ++c) atomicAdd(Kt + (ic + c) + N * (ib + b), kbc[b * NCc + c]);
        }
        if (valid) {                                          // K[ad], K[bd]: consecutive over the lanes
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int d = 0; d < ND; ++d) atomicAdd(Kt + (id + d) + N * (ia + a), kad[a * ND + d]);
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int d = 0; d < ND; ++d) atomicAdd(Kt + (id + d) + N * (ib + b), kbd[b * ND + d]);
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
