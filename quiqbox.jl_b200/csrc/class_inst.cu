// One translation unit per quartet class: compiled 21 times with -DQLA= -DQLB= -DQLC= -DQLD=
// (build.py), so the classes build in parallel.
#include <stdlib.h>

#include <mutex>

#include "digest.cuh"

#define QBX_CAT2(a, b, c, d) qbx_ops_##a##b##c##d
#define QBX_CAT(a, b, c, d) QBX_CAT2(a, b, c, d)

// Classes with >= QBX_COOP_ACC contracted accumulators per quartet are served by the
// warp-cooperative kernel (eri_coop.cu) only: the thread-per-quartet kernel is not instantiated
// for them (it spills, and NVVM needs minutes to tens of minutes to optimise it).
static int launch_eri(const ClassArgs &a, cudaStream_t s)
{
    if constexpr (NCSUM(QLA, QLA + QLB) * NCSUM(QLC, QLC + QLD) >= QBX_COOP_ACC) {
        (void)a; (void)s;
        qbx_set_error("internal: this class is served by the cooperative kernel");
        return QBX_ERR_STATE;
    } else {
    if (a.ntasks <= 0) return QBX_OK;
    static int max_blocks = 0;           // resident blocks on the whole device for this kernel
    if (max_blocks == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        QBX_CUDA(cudaGetDevice(&dev));
        QBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        QBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eri_class_kernel<QLA, QLB, QLC, QLD>,
                                                               QBX_ERI_THREADS, QBX_BOYS_SMEM_BYTES));
        max_blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    // A/B knob (QBX_ERI_SPILL_THREADS=128 or 64, default off): kernels that spill to local memory run with smaller
    // blocks, so that the spill working set of an SM fits its L1 (ncu, profiles/r01: (dp|pp) writes 1.2 GB of spills to DRAM)
    static int threads = 0;
    if (threads == 0) {
        threads = QBX_ERI_THREADS;
        const char *e = getenv("QBX_ERI_SPILL_THREADS");
        cudaFuncAttributes fa;
        if (e && atoi(e) >= 32 && atoi(e) <= QBX_ERI_THREADS && atoi(e) % 32 == 0 &&
            cudaFuncGetAttributes(&fa, eri_class_kernel<QLA, QLB, QLC, QLD>) == cudaSuccess && fa.localSizeBytes > 256)
            threads = atoi(e);
    }
    const int64_t need = (a.ntasks + threads - 1) / threads;
    const unsigned grid = (unsigned)(need < max_blocks ? need : max_blocks);
    eri_class_kernel<QLA, QLB, QLC, QLD><<<grid, threads, QBX_BOYS_SMEM_BYTES, s>>>(a);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
    }
}
// warp-per-task variant, instantiated for the diagonal classes (ab|ab) only: the Schwarz pass
static int launch_eri_split(const ClassArgs &a, cudaStream_t s)
{
    if constexpr (QLA == QLC && QLB == QLD && NCSUM(QLA, QLA + QLB) * NCSUM(QLC, QLC + QLD) < QBX_COOP_ACC) {
        if (a.ntasks <= 0) return QBX_OK;
        int dev = 0, sms = 0;
        QBX_CUDA(cudaGetDevice(&dev));
        QBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t need = (a.ntasks + QBX_ERI_THREADS / 32 - 1) / (QBX_ERI_THREADS / 32);
        const int64_t cap = (int64_t)sms * 4;
        eri_class_split_kernel<QLA, QLB, QLC, QLD><<<(unsigned)(need < cap ? need : cap), QBX_ERI_THREADS, QBX_BOYS_SMEM_BYTES, s>>>(a);
        QBX_CUDA(cudaGetLastError());
        return QBX_OK;
    } else {
        (void)a; (void)s;
        qbx_set_error("internal: no warp-per-task kernel for this class");
        return QBX_ERR_STATE;
    }
}
static int launch_digest(const DigestArgs &a, cudaStream_t s)
{
    if (a.ntasks <= 0) return QBX_OK;
    if (a.ntasks > (int64_t)128 * 0x7fffffff) { qbx_set_error("digest: task list too long for one launch"); return QBX_ERR_ARG; }
    const int64_t nblk = (a.ntasks + 127) / 128;
    const int64_t R = nblk < a.spread ? nblk : a.spread, C = (nblk + R - 1) / R;
    digest_kernel<QLA, QLB, QLC, QLD><<<(unsigned)(R * C), 128, 0, s>>>(a);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}
// span digestion (digest.cuh); -1 = the K rows of a warp do not fit shared memory: use launch_digest
static int launch_digest_span(const DigestArgs &a0, cudaStream_t s)
{
    if (a0.ntasks <= 0) return QBX_OK;
    constexpr int NCOMP = EriClass<QLA, QLB, QLC, QLD>::NCOMP;
    const size_t per_warp = (size_t)digest_span_doubles<QLA, QLB>(a0.wC + a0.wD, a0.nmat) * sizeof(double);
    // warps per block: 8 when two or more such blocks fit an SM, else as many warps as fit (one block per SM)
    int wpb = QBX_SPAN_MAX_WARPS;
    if (2 * wpb * per_warp > QBX_SPAN_MAX_SMEM) wpb = (int)(QBX_SPAN_MAX_SMEM / per_warp) < wpb ? (int)(QBX_SPAN_MAX_SMEM / per_warp) : wpb;
    if (wpb < 1) return -1;
    const size_t smem = wpb * per_warp;
    static std::mutex mu;
    static int sms = 0, per_sm[QBX_SPAN_MAX_WARPS + 1] = {0};
    static size_t attr = 0, occ_smem[QBX_SPAN_MAX_WARPS + 1] = {0};
    int blocks_per_sm;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (sms == 0) {
            int dev = 0;
            QBX_CUDA(cudaGetDevice(&dev));
            QBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        if (smem > attr) {
            QBX_CUDA(cudaFuncSetAttribute(digest_span_kernel<QLA, QLB, QLC, QLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = smem;
        }
        if (occ_smem[wpb] != smem) {
            QBX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[wpb], digest_span_kernel<QLA, QLB, QLC, QLD>, wpb * 32, smem));
            if (per_sm[wpb] < 1) per_sm[wpb] = 1;
            occ_smem[wpb] = smem;
        }
        blocks_per_sm = per_sm[wpb];
    }
    DigestArgs a = a0;
    // span: long enough to amortise flushing the rows, short enough that every resident warp gets several
    const int64_t warps = (int64_t)sms * blocks_per_sm * wpb;
    const int64_t span_min = NCOMP >= 100 ? 32 : (NCOMP >= 27 ? 64 : 256);
    int64_t span = a.ntasks / (warps * 4);
    span = span > 4096 ? 4096 : (span < span_min ? span_min : span / 32 * 32);
    a.span = (int)span;
    const int64_t nspan = (a.ntasks + span - 1) / span, need = (nspan + wpb - 1) / wpb, cap = (int64_t)sms * blocks_per_sm;
    digest_span_kernel<QLA, QLB, QLC, QLD><<<(unsigned)(need < cap ? need : cap), wpb * 32, smem, s>>>(a);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}
static int launch_scatter(const ScatterArgs &a, cudaStream_t s)
{
    if (a.ntasks <= 0) return QBX_OK;
    scatter_kernel<QLA, QLB, QLC, QLD><<<(unsigned)((a.ntasks + 127) / 128), 128, 0, s>>>(a);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

extern const ClassOps QBX_CAT(QLA, QLB, QLC, QLD);
const ClassOps QBX_CAT(QLA, QLB, QLC, QLD) = {QLA, QLB, QLC, QLD,
                                              EriClass<QLA, QLB, QLC, QLD>::NCOMP,
                                              launch_eri, launch_digest, launch_scatter, launch_digest_span,
                                              (QLA == QLC && QLB == QLD && NCSUM(QLA, QLA + QLB) * NCSUM(QLC, QLC + QLD) < QBX_COOP_ACC)
                                                  ? launch_eri_split : nullptr};
