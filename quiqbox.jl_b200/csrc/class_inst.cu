// One translation unit per quartet class: compiled 21 times with -DQLA= -DQLB= -DQLC= -DQLD=
// (build.py), so the classes build in parallel.
#include <stdlib.h>

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
    const int threads = QBX_ERI_THREADS;
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
                                              launch_eri, launch_digest, launch_scatter,
                                              (QLA == QLC && QLB == QLD && NCSUM(QLA, QLA + QLB) * NCSUM(QLC, QLC + QLD) < QBX_COOP_ACC)
                                                  ? launch_eri_split : nullptr};
