// General-contraction sharing on the ket side of the (xs|ss) classes.
//
// cc-pVDZ-type basis sets build several contracted s functions from ONE primitive set (O: two
// 9-term s shells and a 1-term one on the same 9 exponents; H: 4 + 1 on 4).  The reference
// exploits this by de-duplicating primitives (indexGetOrbCore!, src/OrbitalBases.jl:337-365) and
// memoising primitive integrals (TwoBodyIntegralValCache, src/Integration/Framework.jl:68-89).
// Here the s shells of one centre that share exponents form a primitive GROUP; a ket "group
// pair" (P,Q) stands for all its member shell pairs (up to 3 x 3 = 9), the primitive integrals
// over P x Q are evaluated once per bra pair and contracted with the members' coefficient
// products.  (ss|ss), (ps|ss) and (ds|ss) are 49 % of the (H2O)16 ERI time; for O-O kets this cuts
// the primitive work 361 -> 81.
//
// Everything downstream (packed store, digestion, scatter) keeps seeing ordinary tasks: the task
// list of such a class is laid out group task by group task, one slot per member, and slots of
// members that are screened out or violate the bra >= ket uniqueness rule carry ket = -1.
#include <algorithm>
#include <map>
#include <vector>

#include "engine.h"
#include "digest.cuh"

#define QBX_GRP_MAXMEM 9
#define QBX_GRP_NF (5 + QBX_GRP_MAXMEM)      // eta, Qx, Qy, Qz, Kgeom, cc[9]
#define QBX_GRP_CHUNK 32                     // sharding granule in group tasks

namespace {

// ------------------------------------------------------------------ task building
__device__ __forceinline__ bool member_ok(int j, int i, int same, double qi, const double *Qk, double tol)
{
    return j >= 0 && (!same || j <= i) && qi * Qk[j] >= tol;
}

// The three task-building kernels give one WARP to a bra row: lanes take groups g = lane, lane+32, ..
// and positions inside the row come from ballots / a shuffle scan, in group order.
__device__ __forceinline__ bool group_any(const int *members, int g, int i, int same, double qi, const double *Qk, double tol)
{
    bool any = false;
#pragma unroll
    for (int m = 0; m < QBX_GRP_MAXMEM; ++m) any |= member_ok(members[g * QBX_GRP_MAXMEM + m], i, same, qi, Qk, tol);
    return any;
}

__global__ void k_gcount(const double *Qb, const double *Qk, int nb, int same, double tol, int ng, const int *members,
                         int *cnt)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= nb) return;
    const double qi = Qb[i];
    int c = 0;
    for (int g0 = 0; g0 < ng; g0 += 32) {
        const int g = g0 + lane;
        const bool any = g < ng && group_any(members, g, i, same, qi, Qk, tol);
        c += __popc(__ballot_sync(0xffffffffu, any));
    }
    if (lane == 0) cnt[i] = c;
}

__global__ void k_gcount_owned(const double *Qb, const double *Qk, int nb, int same, double tol, int ng, const int *members,
                               const int *nmem, const int64_t *growoff, int rank, int nranks, int *cnt_tasks, int *cnt_slots)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= nb) return;
    const double qi = Qb[i];
    int64_t idx = growoff[i];
    int ct = 0, cs = 0;
    for (int g0 = 0; g0 < ng; g0 += 32) {
        const int g = g0 + lane;
        const bool any = g < ng && group_any(members, g, i, same, qi, Qk, tol);
        const unsigned mask = __ballot_sync(0xffffffffu, any);
        if (any) {
            const int64_t my = idx + __popc(mask & ((1u << lane) - 1u));
            if ((my / QBX_GRP_CHUNK) % nranks == rank) { ++ct; cs += nmem[g]; }
        }
        idx += __popc(mask);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ct += __shfl_xor_sync(0xffffffffu, ct, o); cs += __shfl_xor_sync(0xffffffffu, cs, o); }
    if (lane == 0) { cnt_tasks[i] = ct; cnt_slots[i] = cs; }
}

__global__ void k_gfill(const double *Qb, const double *Qk, int nb, int same, double tol, int ng, const int *members,
                        const int *nmem, const int64_t *growoff, int rank, int nranks, const int64_t *toff,
                        const int64_t *soff, int *gt_bra, int *gt_grp, int *gt_off, int2 *tasks)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
    if (i >= nb) return;
    const double qi = Qb[i];
    int64_t idx = growoff[i], t = toff[i], s = soff[i];
    for (int g0 = 0; g0 < ng; g0 += 32) {
        const int g = g0 + lane;
        const bool any = g < ng && group_any(members, g, i, same, qi, Qk, tol);
        const unsigned mask = __ballot_sync(0xffffffffu, any);
        const int64_t my = idx + __popc(mask & ((1u << lane) - 1u));
        const bool own = any && (my / QBX_GRP_CHUNK) % nranks == rank;
        const unsigned omask = __ballot_sync(0xffffffffu, own);
        const int nm = own ? nmem[g] : 0;
        int incl = nm;                                        // inclusive scan of the slot counts of the owned lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (own) {
            const int64_t tt = t + __popc(omask & ((1u << lane) - 1u)), ss = s + incl - nm;
            gt_bra[tt] = i; gt_grp[tt] = g; gt_off[tt] = (int)ss;
            for (int m = 0; m < nm; ++m) {
                const int j = members[g * QBX_GRP_MAXMEM + m];
                tasks[ss + m] = make_int2(i, member_ok(j, i, same, qi, Qk, tol) ? j : -1);
            }
        }
        idx += __popc(mask);
        t += __popc(omask);
        s += __shfl_sync(0xffffffffu, incl, 31);
    }
}

__global__ void k_gstats(const int2 *tasks, int64_t nslots, const int *gt_bra, const int *gt_grp, int ntasks,
                         const int *poffb, const int *poffg, double *out /*[2]: valid slots, prim quartets*/)
{
    __shared__ double r0[256], r1[256];
    double a = 0, b = 0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nslots; q += (int64_t)gridDim.x * blockDim.x) a += tasks[q].y >= 0;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < ntasks; q += (int64_t)gridDim.x * blockDim.x)
        b += (double)(poffb[gt_bra[q] + 1] - poffb[gt_bra[q]]) * (double)(poffg[gt_grp[q] + 1] - poffg[gt_grp[q]]);
    r0[threadIdx.x] = a; r1[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { r0[threadIdx.x] += r0[threadIdx.x + s]; r1[threadIdx.x] += r1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { atomicAdd(out, r0[0]); atomicAdd(out + 1, r1[0]); }
}

// ------------------------------------------------------------------ the ERI kernel
struct GroupArgs {
    PairSet bra, ket;          // ket = the regular (ss) pair set (member shells -> weights)
    const int2 *tasks;         // slots
    int64_t nslots;
    const int *gt_bra, *gt_grp, *gt_off;
    int ntasks;
    const int *grp_nmem, *grp_prim_off;
    const double *grp_soa;
    const int2 *grp_soa_idx;
    double *out;
    const double *shell_scale;
    BoysTable boys;
    unsigned int *counter;
    const int *order;
    int nheavy;                // leading chunks of `order` that are handed out one TASK per warp
};

// block size / resident blocks the register allocation aims at, per bra class (A/B builds: QBX_NVCC_DEFS)
#ifndef QBX_GRP_THREADS_S
#define QBX_GRP_THREADS_S QBX_ERI_THREADS
#endif
#ifndef QBX_GRP_THREADS_P
#define QBX_GRP_THREADS_P QBX_ERI_THREADS
#endif
#ifndef QBX_GRP_MINB_S
#define QBX_GRP_MINB_S 3
#endif
#ifndef QBX_GRP_MINB_P
#define QBX_GRP_MINB_P 2
#endif
template <int LA> constexpr int grp_threads() { return LA == 0 ? QBX_GRP_THREADS_S : (LA == 1 ? QBX_GRP_THREADS_P : QBX_ERI_THREADS); }
template <int LA> constexpr int grp_minb() { return LA == 0 ? QBX_GRP_MINB_S : (LA == 1 ? QBX_GRP_MINB_P : 1); }

template <int LA>
__global__ void __launch_bounds__(grp_threads<LA>(), grp_minb<LA>()) eri_group_kernel(GroupArgs p)
{
    using EC = EriClass<LA, 0, 0, 0>;
    constexpr int NA = NC(LA);
    extern __shared__ double boys_smem[];
    boys_stage_smem<EC::L>(p.boys, boys_smem);
    __syncthreads();
    const int nchunk = (p.ntasks + 31) / 32;
    const int lane = threadIdx.x & 31;
    // Work queue per warp (see eri_class.cuh).  The first 32 * nheavy positions hand out the tasks
    // of the heaviest chunks ONE AT A TIME: a task with 81 x 81 primitive quartets keeps a lone
    // thread busy for ~3 ms (its ~800-cycle dependent chain per primitive quartet is only hidden
    // while other warps are resident), which sets the run time of the whole launch once the list
    // is split over 8 GPUs.  Such a task is shared by the warp -- the lanes split the ket
    // primitives, partial sums meet in a shuffle reduction.  All other chunks: one task per lane.
    const int nq = 32 * p.nheavy + (nchunk - p.nheavy);
    const double zero3[3] = {0.0, 0.0, 0.0};
    for (;;) {
        unsigned int cq = 0;
        if (lane == 0) cq = atomicAdd(p.counter, 1u);
        const int kq = (int)__shfl_sync(0xffffffffu, cq, 0);
        if (kq >= nq) break;
        const bool coop = kq < 32 * p.nheavy;
        const int chunk = coop ? p.order[kq >> 5] : (p.order ? p.order[p.nheavy + (kq - 32 * p.nheavy)] : kq);
        const int t = chunk * 32 + (coop ? (kq & 31) : lane);
        if (t >= p.ntasks) continue;
        const int ib = p.gt_bra[t], g = p.gt_grp[t], off = p.gt_off[t];
        const double *gb = p.bra.geom + 8 * (int64_t)ib;
        const double A[3] = {gb[0], gb[1], gb[2]};
        const int pb0 = p.bra.prim_off[ib], pb1 = p.bra.prim_off[ib + 1];
        const int nk = p.grp_prim_off[g + 1] - p.grp_prim_off[g];
        const int2 si = p.grp_soa_idx[g];
        const int nmem = p.grp_nmem[g];
        double acc[QBX_GRP_MAXMEM][NA];
#pragma unroll
        for (int m = 0; m < QBX_GRP_MAXMEM; ++m)
#pragma unroll
            for (int c = 0; c < NA; ++c) acc[m][c] = 0.0;
        // ket primitive outermost: its record and the members' coefficient products are loaded once,
        // the bra primitives run inside, and the coefficient contraction happens once per ket
        // primitive instead of once per primitive quartet
        for (int pk = coop ? lane : 0; pk < nk; pk += coop ? 32 : 1) {
            const double *kp = p.grp_soa + si.x + (int64_t)pk * QBX_GRP_NF * si.y;
            const double eta = __ldg(kp);
            const double Q[3] = {__ldg(kp + si.y), __ldg(kp + 2 * si.y), __ldg(kp + 3 * si.y)};
            const double Kg = __ldg(kp + 4 * si.y);
            double v[NA];
#pragma unroll
            for (int c = 0; c < NA; ++c) v[c] = 0.0;
            for (int pb = pb0; pb < pb1; ++pb) {
                const double4 b0 = ldg4(p.bra.prim + 8 * (int64_t)pb);
                const double4 b1 = ldg4(p.bra.prim + 8 * (int64_t)pb + 4);
                const double P[3] = {b0.y, b0.z, b0.w};
                const double PA[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
                EC::primitive(v, boys_smem, b0.x, P, b1.x, PA, b1.z, zero3, eta, Q, Kg, 0.0, zero3);
            }
#pragma unroll
            for (int m = 0; m < QBX_GRP_MAXMEM; ++m) {
                if (m < nmem) {
                    const double cc = __ldg(kp + (5 + m) * si.y);
#pragma unroll
                    for (int c = 0; c < NA; ++c) acc[m][c] = fma(cc, v[c], acc[m][c]);
                }
            }
        }
        if (coop) {
#pragma unroll
            for (int m = 0; m < QBX_GRP_MAXMEM; ++m)
#pragma unroll
                for (int c = 0; c < NA; ++c)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc[m][c] += __shfl_xor_sync(0xffffffffu, acc[m][c], o);
            if (lane != 0) continue;
        }
        const int2 sb = p.bra.shells[ib];
        const double *sA = p.shell_scale + 6 * sb.x;
        const double sB = p.shell_scale[6 * sb.y];
#pragma unroll
        for (int m = 0; m < QBX_GRP_MAXMEM; ++m) {
            if (m >= nmem) break;
            const int j = p.tasks[off + m].y;
            double w = 0.0;
            if (j >= 0) {
                const int2 sk = p.ket.shells[j];
                w = sB * p.shell_scale[6 * sk.x] * p.shell_scale[6 * sk.y];
            }
#pragma unroll
            for (int c = 0; c < NA; ++c) p.out[(int64_t)c * p.nslots + off + m] = acc[m][c] * sA[c] * w;
        }
    }
}

template <int LA>
int launch_group(GroupArgs a, cudaStream_t s)
{
    static int max_blocks = 0;
    if (max_blocks == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        QBX_CUDA(cudaGetDevice(&dev));
        QBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        QBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eri_group_kernel<LA>, grp_threads<LA>(), QBX_BOYS_SMEM_BYTES));
        max_blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int need = (a.ntasks + grp_threads<LA>() - 1) / grp_threads<LA>();
    // task-granular hand-out of the heavy chunks only when the launch is short (few chunks per
    // resident warp), i.e. when the tail decides; on a long list the one-task-per-lane mode is
    // ~8 % faster and the tail is filled by the light chunks anyway
    if ((a.ntasks + 31) / 32 >= 8 * max_blocks * (grp_threads<LA>() / 32)) a.nheavy = 0;
    eri_group_kernel<LA><<<need < max_blocks ? need : max_blocks, grp_threads<LA>(), QBX_BOYS_SMEM_BYTES, s>>>(a);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}


// ---- group-pair primitive records on the device (see engine.cu: k_pair_count / k_pair_fill)
struct GroupTables {
    const double *gcen;          // [npg][3] centre of a primitive group
    const int *gxoff;            // [npg + 1] exponents of group g at gxpn[gxoff[g] ..]
    const double *gxpn;
    const int *scoef_off;        // [nshell + 1]: coefficients of s shell s over its group's primitives
    const double *scoef;
    const int *spg;              // [nshell] group of an s shell
    const int2 *gp;              // [ng] (P, Q)
    const int *gmem;             // [ng][9] member pair indices (regular (ss) pairs), -1 = none
    const int2 *ss;              // regular (ss) pairs: shells (C, D)
};

__device__ __forceinline__ double group_prim(const GroupTables &T, int g, int P, int Q, int a, int b, double pref, double pq2,
                                             double (&v)[QBX_GRP_NF])
{
    const double x = T.gxpn[a], y = T.gxpn[b], z = x + y;
    v[0] = z;
    for (int d = 0; d < 3; ++d) v[1 + d] = (x * T.gcen[3 * P + d] + y * T.gcen[3 * Q + d]) / z;
    v[4] = pref * exp(-x * y / z * pq2) / z;
    const int pa = a - T.gxoff[P], pb = b - T.gxoff[Q];
    double big = 0.0;
    for (int m = 0; m < QBX_GRP_MAXMEM; ++m) {
        double cc = 0.0;
        const int j = T.gmem[g * QBX_GRP_MAXMEM + m];
        if (j >= 0) {
            const int2 cd = T.ss[j];
            if (T.spg[cd.x] == P) cc = T.scoef[T.scoef_off[cd.x] + pa] * T.scoef[T.scoef_off[cd.y] + pb];
            else cc = T.scoef[T.scoef_off[cd.y] + pa] * T.scoef[T.scoef_off[cd.x] + pb];
        }
        v[5 + m] = cc;
        big = fmax(big, fabs(cc * v[4]));
    }
    return big;
}

__global__ void k_group_count(GroupTables T, int ng, double pref, int *cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    const int P = T.gp[g].x, Q = T.gp[g].y;
    double pq2 = 0;
    for (int d = 0; d < 3; ++d) { const double t = T.gcen[3 * P + d] - T.gcen[3 * Q + d]; pq2 += t * t; }
    int c = 0;
    double v[QBX_GRP_NF];
    for (int a = T.gxoff[P]; a < T.gxoff[P + 1]; ++a)
        for (int b = T.gxoff[Q]; b < T.gxoff[Q + 1]; ++b)
            if (group_prim(T, g, P, Q, a, b, pref, pq2, v) >= 1e-24) ++c;
    cnt[g] = c;
}

__global__ void k_group_fill(GroupTables T, int ng, double pref, const int *order, const int2 *soa_idx, double *soa)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ng) return;
    const int g = order[n];
    const int P = T.gp[g].x, Q = T.gp[g].y;
    double pq2 = 0;
    for (int d = 0; d < 3; ++d) { const double t = T.gcen[3 * P + d] - T.gcen[3 * Q + d]; pq2 += t * t; }
    const int64_t b0 = soa_idx[n].x, gs = soa_idx[n].y;
    int c = 0;
    double v[QBX_GRP_NF];
    for (int a = T.gxoff[P]; a < T.gxoff[P + 1]; ++a)
        for (int b = T.gxoff[Q]; b < T.gxoff[Q + 1]; ++b) {
            if (group_prim(T, g, P, Q, a, b, pref, pq2, v) < 1e-24) continue;
            for (int k = 0; k < QBX_GRP_NF; ++k) soa[b0 + ((int64_t)c * QBX_GRP_NF + k) * gs] = v[k];
            ++c;
        }
}


#ifndef QBX_DGRP_MINB_S
#define QBX_DGRP_MINB_S 5               // resident blocks of 128 threads the register allocation aims at (A/B: tools/gpu_ab_digest.sh)
#endif
#ifndef QBX_DGRP_MINB_P
#define QBX_DGRP_MINB_P 3
#endif
// ------------------------------------------------------------------ J/K digestion of the group classes: lane = group task
// The slot list of an (x s|ss) class is laid out group task by group task: up to 9 member quartets (3 shells C of one
// centre x 3 shells D of another) behind ONE bra pair.  The per-quartet kernel of digest.cuh sees them as 32 unrelated
// lanes whose C repeats in runs of 3 and whose D repeats with stride 3: segmented sums three lanes long, REDs of one
// instruction hitting the same address three times, and a quarter of the slots empty (screened members) -- the (ps|ss)
// launch was the longest of the whole Fock build.  Here ONE LANE digests a whole group task: the <= 3 distinct C and the
// <= 3 distinct D of its members are found on the fly, K[x,c] and K[x,d] (x = the bra functions a.., b) are summed over the
// members in registers and leave as one RED per distinct shell, the density elements D[x,c], D[x,d] are loaded once per
// distinct shell, empty slots cost one compare.  J[ab] is summed over the lanes that share the bra pair as before.
template <int LA>
__global__ void __launch_bounds__(128, LA == 0 ? QBX_DGRP_MINB_S : QBX_DGRP_MINB_P) digest_group_kernel(DigestArgs p, const int *__restrict__ gt_bra,
                                                                          const int *__restrict__ gt_grp, const int *__restrict__ gt_off,
                                                                          const int *__restrict__ grp_nmem, const int *__restrict__ grp_flip, int ngt)
{
    constexpr int NA = NC(LA), NX = NA + 1;
    const unsigned nblk = (unsigned)((ngt + 127) / 128);
    const unsigned R = nblk < (unsigned)p.spread ? nblk : (unsigned)p.spread, Cb = (nblk + R - 1) / R;
    const unsigned blk = (blockIdx.x % R) * Cb + blockIdx.x / R;          // blocks resident together sit on different bra rows
    if (blk >= nblk) return;
    const int t0 = (int)(blk * 128 + threadIdx.x), lane = threadIdx.x & 31;
    if (t0 - lane >= ngt) return;
    const bool live = t0 < ngt;
    const int t = live ? t0 : ngt - 1;
    const int ib = __ldg(gt_bra + t), off = __ldg(gt_off + t);
    const int grp = __ldg(gt_grp + t);
    const int nmem = live ? __ldg(grp_nmem + grp) : 0;
    const int flip = __ldg(grp_flip + grp);                     // members listed as (shell of Q, shell of P): swapped below, so
                                                                // that "c" is always one of P's <= 3 shells and "d" one of Q's
    const int4 rb = __ldg(p.bra_info + ib);
    const int N = p.nbf, ia = rb.z, ibf = rb.w;
    const double fab = rb.x == rb.y ? 0.5 : 1.0;
    const int64_t nt = p.ntasks;
    bool headAB;
    int endAB;
    seg_runs(ib, 0, lane, headAB, endAB);
    for (int mm = 0; mm < p.nmat; ++mm) {
        const double *__restrict__ DKm = p.DK + (int64_t)mm * N * N;
        double *Ktm = p.Kt + (int64_t)mm * N * N;
        const bool coul = mm == 0;
        int cfun[3] = {-1, -1, -1}, dfun[3] = {-1, -1, -1};
        double dxc[NX][3], dxd[NX][3], kxc[NX][3], kxd[NX][3], jab[NA], dab[NA];
#pragma unroll
        for (int x = 0; x < NX; ++x)
#pragma unroll
            for (int k = 0; k < 3; ++k) { dxc[x][k] = dxd[x][k] = 0.0; kxc[x][k] = kxd[x][k] = 0.0; }
#pragma unroll
        for (int a = 0; a < NA; ++a) { jab[a] = 0.0; dab[a] = coul ? __ldg(p.DJ + ibf + (int64_t)N * (ia + a)) : 0.0; }
#pragma unroll
        for (int m = 0; m < QBX_GRP_MAXMEM; ++m) {
            if (m >= nmem) break;
            const int64_t slot = (int64_t)off + m;
            const int j = __ldg(&p.tasks[slot].y);
            if (j < 0) continue;                                           // member screened out / not unique
            const int4 rk = __ldg(p.ket_info + j);
            const bool sw = (flip >> m) & 1;
            const int ic = sw ? rk.w : rk.z, id = sw ? rk.z : rk.w;       // (ab|cd) = (ab|dc): every update below is symmetric in it
            double f = fab;
            if (rk.x == rk.y) f *= 0.5;
            if (p.same_class && j == ib) f *= 0.5;
            // position of C and D among the distinct shells seen so far; a new one brings its density elements along
            bool pc[3], pd[3];
            {
                const bool h0 = cfun[0] == ic, h1 = cfun[1] == ic, h2 = cfun[2] == ic;
                const bool fresh = !(h0 || h1 || h2);
                pc[0] = h0 || (fresh && cfun[0] < 0);
                pc[1] = h1 || (fresh && cfun[0] >= 0 && cfun[1] < 0);
                pc[2] = h2 || (fresh && cfun[0] >= 0 && cfun[1] >= 0);
                if (fresh) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (pc[k]) {
                            cfun[k] = ic;
#pragma unroll
                            for (int x = 0; x < NX; ++x) dxc[x][k] = __ldg(DKm + ic + (int64_t)N * (x < NA ? ia + x : ibf));
                        }
                }
            }
            {
                const bool h0 = dfun[0] == id, h1 = dfun[1] == id, h2 = dfun[2] == id;
                const bool fresh = !(h0 || h1 || h2);
                pd[0] = h0 || (fresh && dfun[0] < 0);
                pd[1] = h1 || (fresh && dfun[0] >= 0 && dfun[1] < 0);
                pd[2] = h2 || (fresh && dfun[0] >= 0 && dfun[1] >= 0);
                if (fresh) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (pd[k]) {
                            dfun[k] = id;
#pragma unroll
                            for (int x = 0; x < NX; ++x) dxd[x][k] = __ldg(DKm + id + (int64_t)N * (x < NA ? ia + x : ibf));
                        }
                }
            }
            double dsc[NX], dsd[NX];                                       // D[x,c], D[x,d] of this member
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                dsc[x] = pc[0] ? dxc[x][0] : (pc[1] ? dxc[x][1] : dxc[x][2]);
                dsd[x] = pd[0] ? dxd[x][0] : (pd[1] ? dxd[x][1] : dxd[x][2]);
            }
            const double dcd = coul ? __ldg(p.DJ + id + (int64_t)N * ic) : 0.0;
            double jcd = 0.0, kbc = 0.0, kbd = 0.0;
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double x = f * __ldg(p.vals + (int64_t)a * nt + slot);
                jab[a] = fma(dcd, x, jab[a]);
                jcd = fma(dab[a], x, jcd);
                const double kac = dsd[NA] * x, kad = dsc[NA] * x;         // K[a,c] += D[b,d] v   K[a,d] += D[b,c] v
                kbc = fma(dsd[a], x, kbc);                                 // K[b,c] += D[a,d] v
                kbd = fma(dsc[a], x, kbd);                                 // K[b,d] += D[a,c] v
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (pc[k]) kxc[a][k] += kac;
                    if (pd[k]) kxd[a][k] += kad;
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (pc[k]) kxc[NA][k] += kbc;
                if (pd[k]) kxd[NA][k] += kbd;
            }
            if (coul) atomicAdd(p.Jt + id + (int64_t)N * ic, 2.0 * jcd);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (cfun[k] >= 0) {
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    if (kxc[x][k] != 0.0) atomicAdd(Ktm + cfun[k] + (int64_t)N * (x < NA ? ia + x : ibf), kxc[x][k]);
            }
            if (dfun[k] >= 0) {
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    if (kxd[x][k] != 0.0) atomicAdd(Ktm + dfun[k] + (int64_t)N * (x < NA ? ia + x : ibf), kxd[x][k]);
            }
        }
        if (coul) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
                const double js = seg_sum(2.0 * jab[a], endAB, lane);
                if (headAB && js != 0.0) atomicAdd(p.Jt + ibf + (int64_t)N * (ia + a), js);
            }
        }
    }
}

template <int LA>
int launch_digest_group(const DigestArgs &a, const TaskList &tl, const GroupSet &G, cudaStream_t s)
{
    const unsigned nblk = (unsigned)((tl.ngt + 127) / 128);
    const unsigned R = nblk < (unsigned)a.spread ? nblk : (unsigned)a.spread, Cb = (nblk + R - 1) / R;
    digest_group_kernel<LA><<<R * Cb, 128, 0, s>>>(a, tl.gt_bra, tl.gt_grp, tl.gt_off, G.nmem, G.flip, tl.ngt);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}
}  // namespace

// ------------------------------------------------------------------ host side
// Primitive groups of the s shells and the (ss) group pairs.  `ss_pairs` = shells of every
// regular (ss) pair in the order of the device pair set.
int qbx_group_build(const std::vector<HostShell> &sh, const std::vector<int2> &ss_pairs, GroupSet &out, const int2 *d_ss_pairs)
{
    // 1. primitive groups: s shells on one centre whose exponents are a subset of the group's
    struct PG { double cen[3]; std::vector<double> xpn; std::vector<int> shells; };
    std::vector<PG> pgs;
    std::vector<int> pg_of(sh.size(), -1);
    std::vector<std::vector<double>> coef(sh.size());         // per shell, over its group's primitives
    for (size_t s = 0; s < sh.size(); ++s) {
        if (sh[s].l != 0) continue;
        int found = -1;
        for (size_t g = 0; g < pgs.size() && found < 0; ++g) {
            PG &G = pgs[g];
            if (G.shells.size() >= 3 || G.cen[0] != sh[s].cen[0] || G.cen[1] != sh[s].cen[1] || G.cen[2] != sh[s].cen[2]) continue;
            bool sub = true;
            for (double x : sh[s].xpn) sub &= std::find(G.xpn.begin(), G.xpn.end(), x) != G.xpn.end();
            if (sub) found = (int)g;
        }
        if (found < 0) {
            PG G;
            for (int d = 0; d < 3; ++d) G.cen[d] = sh[s].cen[d];
            G.xpn = sh[s].xpn;                               // shells come longest contraction first
            pgs.push_back(G);
            found = (int)pgs.size() - 1;
        }
        PG &G = pgs[found];
        G.shells.push_back((int)s);
        pg_of[s] = found;
        coef[s].assign(G.xpn.size(), 0.0);
        for (size_t k = 0; k < sh[s].xpn.size(); ++k) {
            const size_t pos = std::find(G.xpn.begin(), G.xpn.end(), sh[s].xpn[k]) - G.xpn.begin();
            coef[s][pos] += sh[s].coef[k];
        }
    }
    // 2. group pairs and their members (regular pair indices)
    const size_t npg = pgs.size();
    std::vector<int> gp_index(npg * npg, -1);                // (P, Q) -> group pair, in order of first appearance
    std::vector<int> members, nmem;                          // flat: QBX_GRP_MAXMEM slots per group pair
    std::vector<std::pair<int, int>> gp_pq;
    members.reserve(npg * (npg + 1) / 2 * QBX_GRP_MAXMEM);
    for (size_t j = 0; j < ss_pairs.size(); ++j) {
        int P = pg_of[ss_pairs[j].x], Q = pg_of[ss_pairs[j].y];
        if (P < Q) std::swap(P, Q);
        int &gi = gp_index[(size_t)P * npg + Q];
        if (gi < 0) {
            gi = (int)nmem.size();
            nmem.push_back(0);
            members.resize(members.size() + QBX_GRP_MAXMEM, -1);
            gp_pq.push_back({P, Q});
        }
        if (nmem[gi] >= QBX_GRP_MAXMEM) { qbx_set_error("internal: group pair with more than 9 members"); return QBX_ERR_STATE; }
        members[(size_t)gi * QBX_GRP_MAXMEM + nmem[gi]++] = (int)j;
    }
    const size_t ng0 = nmem.size();
    const double pref = sqrt(2.0) * pow(M_PI, 1.25);
    if (ng0 > 0) {
        // ---- records on the device: count -> host sort by count -> fill
        cudaStream_t st = qbx_stream();
        std::vector<double> gcen(3 * pgs.size()), gxpn, scoef;
        std::vector<int> gxoff(pgs.size() + 1, 0), scoef_off(sh.size() + 1, 0);
        const std::vector<int> &gmem = members;
        std::vector<int2> gp(ng0);
        for (size_t g = 0; g < pgs.size(); ++g) {
            for (int d = 0; d < 3; ++d) gcen[3 * g + d] = pgs[g].cen[d];
            gxpn.insert(gxpn.end(), pgs[g].xpn.begin(), pgs[g].xpn.end());
            gxoff[g + 1] = (int)gxpn.size();
        }
        for (size_t i = 0; i < sh.size(); ++i) {
            scoef.insert(scoef.end(), coef[i].begin(), coef[i].end());
            scoef_off[i + 1] = (int)scoef.size();
        }
        for (size_t g = 0; g < ng0; ++g) gp[g] = make_int2(gp_pq[g].first, gp_pq[g].second);
        void *d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const void *src[8] = {gcen.data(), gxoff.data(), gxpn.data(), scoef_off.data(), scoef.data(), pg_of.data(), gp.data(), gmem.data()};
        const size_t bytes[8] = {gcen.size() * 8, gxoff.size() * 4, gxpn.size() * 8, scoef_off.size() * 4, scoef.size() * 8,
                                 pg_of.size() * 4, gp.size() * sizeof(int2), gmem.size() * 4};
        for (int i = 0; i < 8; ++i) {
            QBX_CUDA(qbx_pool_malloc(&d[i], std::max<size_t>(8, bytes[i])));
            if (bytes[i]) QBX_CUDA(cudaMemcpyAsync(d[i], src[i], bytes[i], cudaMemcpyHostToDevice, st));
        }
        GroupTables T{(const double *)d[0], (const int *)d[1], (const double *)d[2], (const int *)d[3], (const double *)d[4],
                      (const int *)d[5], (const int2 *)d[6], (const int *)d[7], d_ss_pairs};
        int *d_cnt = nullptr, *d_order = nullptr;
        QBX_CUDA(qbx_dmalloc(&d_cnt, ng0 * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&d_order, ng0 * sizeof(int)));
        std::vector<int> cnt(ng0);
        k_group_count<<<(unsigned)((ng0 + 127) / 128), 128, 0, st>>>(T, (int)ng0, pref, d_cnt);
        QBX_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, ng0 * sizeof(int), cudaMemcpyDeviceToHost, st));
        QBX_CUDA(cudaStreamSynchronize(st));
        std::vector<int> order(ng0);
        {   // stable, descending count (counting sort)
            int cmax = 0;
            for (int c : cnt) cmax = std::max(cmax, c);
            std::vector<int> start(cmax + 2, 0);
            for (int c : cnt) ++start[cmax - c + 1];
            for (int c = 0; c <= cmax; ++c) start[c + 1] += start[c];
            for (size_t g = 0; g < ng0; ++g) order[start[cmax - cnt[g]]++] = (int)g;
        }
        out.ng = (int)ng0;
        out.h_nprim.resize(ng0); out.h_nmem.resize(ng0);
        std::vector<int> mem(ng0 * QBX_GRP_MAXMEM, -1), poff(ng0 + 1, 0), flip(ng0, 0);
        for (size_t n = 0; n < ng0; ++n) {
            const int g = order[n];
            out.h_nprim[n] = cnt[g];
            out.h_nmem[n] = nmem[g];
            for (int m = 0; m < nmem[g]; ++m) {
                const int j = members[(size_t)g * QBX_GRP_MAXMEM + m];
                mem[n * QBX_GRP_MAXMEM + m] = j;
                if (pg_of[ss_pairs[j].x] != gp_pq[g].first) flip[n] |= 1 << m;
            }
            poff[n + 1] = poff[n] + cnt[g];
        }
        std::vector<int2> soa_idx(ng0);
        size_t base = 0, g0 = 0;
        while (g0 < ng0) {
            size_t g1 = g0;
            while (g1 < ng0 && out.h_nprim[g1] == out.h_nprim[g0]) ++g1;
            const size_t gs = g1 - g0;
            for (size_t n = g0; n < g1; ++n) soa_idx[n] = make_int2((int)(base + (n - g0)), (int)gs);
            base += gs * (size_t)out.h_nprim[g0] * QBX_GRP_NF;
            g0 = g1;
        }
        const size_t n_soa = (size_t)QBX_GRP_NF * poff.back();
        QBX_CUDA(qbx_dmalloc(&out.nmem, ng0 * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&out.members, mem.size() * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&out.flip, ng0 * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&out.prim_off, poff.size() * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&out.soa, std::max<size_t>(1, n_soa) * sizeof(double)));
        QBX_CUDA(qbx_dmalloc(&out.soa_idx, ng0 * sizeof(int2)));
        QBX_CUDA(cudaMemcpyAsync(out.nmem, out.h_nmem.data(), ng0 * sizeof(int), cudaMemcpyHostToDevice, st));
        QBX_CUDA(cudaMemcpyAsync(out.members, mem.data(), mem.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        QBX_CUDA(cudaMemcpyAsync(out.flip, flip.data(), ng0 * sizeof(int), cudaMemcpyHostToDevice, st));
        QBX_CUDA(cudaMemcpyAsync(out.soa_idx, soa_idx.data(), ng0 * sizeof(int2), cudaMemcpyHostToDevice, st));
        QBX_CUDA(cudaMemcpyAsync(out.prim_off, poff.data(), poff.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        QBX_CUDA(cudaMemcpyAsync(d_order, order.data(), ng0 * sizeof(int), cudaMemcpyHostToDevice, st));
        k_group_fill<<<(unsigned)((ng0 + 127) / 128), 128, 0, st>>>(T, (int)ng0, pref, d_order, out.soa_idx, out.soa);
        QBX_CUDA(cudaGetLastError());
        QBX_CUDA(cudaStreamSynchronize(st));                 // host vectors go out of scope
        for (void *q : d) qbx_pool_free_async(q);
        qbx_pool_free_async(d_cnt); qbx_pool_free_async(d_order);
        return QBX_OK;
    }
    out.ng = 0;                                              // no (ss) pairs: nothing to share
    (void)pref;
    return QBX_OK;
}

void qbx_group_free(GroupSet &g)
{
    qbx_pool_free(g.nmem); qbx_pool_free(g.members); qbx_pool_free(g.flip); qbx_pool_free(g.prim_off); qbx_pool_free(g.soa); qbx_pool_free(g.soa_idx);
    g = GroupSet();
}

// Builds the slot list (tl.tasks / tl.n) and the group tasks of one (x s|ss) class for this rank.
int qbx_group_count(const GroupSet &G, const DevPairSet &B, const DevPairSet &K, bool same, double tol, int rank, int nranks,
                    TaskScratch &ts, int64_t *d_total, cudaStream_t s)
{
    if (B.npair == 0 || G.ng == 0) return QBX_OK;
    const int nb = B.npair, grid = (int)(((int64_t)nb * 32 + 127) / 128);
    for (int i = 0; i < 3; ++i) {
        QBX_CUDA(qbx_dmalloc(&ts.cnt[i], nb * sizeof(int)));
        QBX_CUDA(qbx_dmalloc(&ts.off[i], (nb + 1) * sizeof(int64_t)));
    }
    k_gcount<<<grid, 128, 0, s>>>(B.schwarz, K.schwarz, nb, same, tol, G.ng, G.members, ts.cnt[0]);
    qbx_scan_counts(ts.cnt[0], nb, ts.off[0], d_total, s);
    k_gcount_owned<<<grid, 128, 0, s>>>(B.schwarz, K.schwarz, nb, same, tol, G.ng, G.members, G.nmem, ts.off[0], rank, nranks,
                                        ts.cnt[1], ts.cnt[2]);
    qbx_scan_counts(ts.cnt[1], nb, ts.off[1], d_total + 1, s);
    qbx_scan_counts(ts.cnt[2], nb, ts.off[2], d_total + 2, s);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int qbx_group_fill(const GroupSet &G, const DevPairSet &B, const DevPairSet &K, bool same, double tol, int rank, int nranks,
                   TaskScratch &ts, const int64_t *h_total, TaskList &tl, double *d_stat, int *d_nheavy, cudaStream_t s)
{
    tl = TaskList();
    if (B.npair == 0 || G.ng == 0) return QBX_OK;
    const int nb = B.npair, grid = (int)(((int64_t)nb * 32 + 127) / 128);
    tl.ngt = (int)h_total[1];
    tl.n = h_total[2];
    if (tl.ngt <= 0) { tl.ngt = 0; tl.n = 0; return QBX_OK; }
    QBX_CUDA(qbx_dmalloc(&tl.tasks, tl.n * sizeof(int2)));
    QBX_CUDA(qbx_dmalloc(&tl.gt_bra, tl.ngt * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&tl.gt_grp, tl.ngt * sizeof(int)));
    QBX_CUDA(qbx_dmalloc(&tl.gt_off, tl.ngt * sizeof(int)));
    k_gfill<<<grid, 128, 0, s>>>(B.schwarz, K.schwarz, nb, same, tol, G.ng, G.members, G.nmem, ts.off[0], rank, nranks, ts.off[1],
                                 ts.off[2], tl.gt_bra, tl.gt_grp, tl.gt_off, tl.tasks);
    k_gstats<<<296, 256, 0, s>>>(tl.tasks, tl.n, tl.gt_bra, tl.gt_grp, tl.ngt, B.prim_off, G.prim_off, d_stat);
    QBX_CUDA(cudaGetLastError());
    return qbx_chunk_order(nullptr, tl.gt_bra, tl.gt_grp, tl.ngt, B.prim_off, G.prim_off, &tl.order, d_nheavy, s);
}

int qbx_group_eri(int la, const GroupSet &G, const ClassArgs &a, const TaskList &tl, cudaStream_t s)
{
    if (tl.ngt <= 0) return QBX_OK;
    GroupArgs g;
    g.bra = a.bra; g.ket = a.ket; g.tasks = tl.tasks; g.nslots = tl.n;
    g.gt_bra = tl.gt_bra; g.gt_grp = tl.gt_grp; g.gt_off = tl.gt_off; g.ntasks = tl.ngt;
    g.grp_nmem = G.nmem; g.grp_prim_off = G.prim_off; g.grp_soa = G.soa; g.grp_soa_idx = G.soa_idx;
    g.out = a.out; g.shell_scale = a.shell_scale; g.boys = a.boys; g.counter = a.counter; g.order = tl.order; g.nheavy = tl.order ? tl.nheavy : 0;
    switch (la) {
    case 0: return launch_group<0>(g, s);
    case 1: return launch_group<1>(g, s);
    case 2: return launch_group<2>(g, s);
    }
    qbx_set_error("internal: no group kernel for this class");
    return QBX_ERR_STATE;
}

// J/K digestion of a stored group class with one lane per group task; -1 = not served (the per-quartet kernel does it).
int qbx_group_digest(int la, const GroupSet &G, const DigestArgs &a, const TaskList &tl, cudaStream_t s)
{
    if (tl.ngt <= 0) return QBX_OK;
    switch (la) {
    case 0: return launch_digest_group<0>(a, tl, G, s);
    case 1: return launch_digest_group<1>(a, tl, G, s);
    }
    return -1;                                               // (ds|ss): 7 x 6 accumulators per lane do not fit
}
