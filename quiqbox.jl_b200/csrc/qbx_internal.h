// Internal declarations shared by the translation units of libqbx.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qbx.h"
#include "boys.cuh"

#define QBX_MAX_L 2                 // class kernels cover s, p, d
#define QBX_NPAIRCLS 6              // (ss) (ps) (pp) (ds) (dp) (dd)
#define QBX_NCLASS 21               // canonical quartet classes, bra pair class >= ket pair class
#define QBX_COOP_ACC 180             // >= this many [e0|f0] accumulators: warp-cooperative ERI kernel
#define QBX_GEN_MAXL 64             // generic kernel: max total angular momentum of a quartet
#define QBX_GEN_MAXAX 32            // generic kernel: max angular momentum sum on one axis

void qbx_set_error(const std::string &msg);
#define QBX_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            qbx_set_error(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " at " + \
                          __FILE__ + ":" + std::to_string(__LINE__));                         \
            return QBX_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

// pool.cu: device allocations of the library go through a size-keyed pool
cudaError_t qbx_pool_malloc(void **p, size_t bytes);
cudaError_t qbx_pool_free(void *p);
cudaError_t qbx_pool_free_async(void *p);
// One device synchronisation for a batch of qbx_pool_free calls (a handle's destructor returns ~100 blocks): while a
// scope is alive on this thread qbx_pool_free does not synchronise again.  Nothing may be enqueued inside the scope.
struct QbxPoolFreeScope {
    QbxPoolFreeScope();
    ~QbxPoolFreeScope();
    QbxPoolFreeScope(const QbxPoolFreeScope &) = delete;
    QbxPoolFreeScope &operator=(const QbxPoolFreeScope &) = delete;
};
// While a scope is alive on this thread qbx_pool_free_async only remembers the blocks; release() returns them to the
// pool.  For phases that run on the side streams: a block freed by one stream must not be handed to another before the
// streams have been joined (release() is called after the join has been enqueued).
struct QbxPoolDeferScope {
    QbxPoolDeferScope();
    ~QbxPoolDeferScope();               // releases if release() was not called
    void release();
    QbxPoolDeferScope(const QbxPoolDeferScope &) = delete;
    QbxPoolDeferScope &operator=(const QbxPoolDeferScope &) = delete;
};
// Scratch block of an entry point: returned to the pool when the holder goes out of scope, also on the early returns of
// QBX_CUDA (the blocking free: the work that uses the block has finished when the pool gets it back).
template <class T>
struct QbxScratch {
    T *p = nullptr;
    QbxScratch() = default;
    ~QbxScratch() { if (p) qbx_pool_free(p); }
    QbxScratch(const QbxScratch &) = delete;
    QbxScratch &operator=(const QbxScratch &) = delete;
    cudaError_t alloc(size_t bytes) { return qbx_pool_malloc((void **)&p, bytes); }
    operator T *() const { return p; }
};
void *qbx_pinned(size_t bytes);
void *qbx_staging(size_t bytes);
std::mutex &qbx_staging_mutex();
void qbx_pool_release();
void qbx_pool_counts(int64_t *hits, int64_t *misses, int64_t *idle_bytes);
template <class T> inline cudaError_t qbx_dmalloc(T **p, size_t bytes) { return qbx_pool_malloc((void **)p, bytes); }

// f(lo, hi) over [0, n) on a few host threads, `grain` items at a time.  The workers are persistent
// (pool.cu): a qbx_basis_create makes a dozen of these calls, and spawning seven std::threads for
// each cost more than the work of the smaller pair classes.
void qbx_host_workers_run(unsigned nthreads, void (*fn)(void *), void *ctx);     // runs fn(ctx) on nthreads threads (caller included)
unsigned qbx_host_workers();                                                      // threads available (<= 8)
template <class F>
inline void qbx_parallel_for(size_t n, size_t grain, F f)
{
    const unsigned nt = (unsigned)std::min<size_t>(qbx_host_workers(), (n + grain - 1) / grain);
    if (nt <= 1) { f(0, n); return; }
    struct Ctx { std::atomic<size_t> next; size_t n, grain; F *f; } ctx{{0}, n, grain, &f};
    qbx_host_workers_run(nt, [](void *p) {
        Ctx &c = *static_cast<Ctx *>(p);
        for (;;) {
            const size_t lo = c.next.fetch_add(c.grain);
            if (lo >= c.n) return;
            (*c.f)(lo, std::min(c.n, lo + c.grain));
        }
    }, &ctx);
}

__host__ __device__ constexpr int qbx_nc(int l) { return (l + 1) * (l + 2) / 2; }

// flat primitive table + CSR on the device (generic kernels)
struct DevFlat {
    int64_t nprim, nbf, nnz;
    double *cen, *xpn;      // 3 x nprim, nprim
    int32_t *ang;           // 3 x nprim
    int64_t *bf_off, *bf_prim;
    double *bf_w;
};

BoysTable qbx_boys_table();         // device pointers of the Taylor table (built on first use)
cudaStream_t qbx_stream();

// generic.cu
int qbx_launch_generic_quartets(const DevFlat &f, int64_t n, const int64_t *d_ijkl, double *d_out, cudaStream_t s);
int qbx_launch_generic_tensor(const DevFlat &f, double *d_tensor, cudaStream_t s);
int qbx_launch_one_body(const DevFlat &f, int kind, int64_t nnuc, const double *dZ, const double *dR, double *d_out,
                        cudaStream_t s);
int qbx_launch_dense_gcore(int64_t n, const double *dH, int nmat, const double *dDJ, const double *dDK, double *dG,
                           cudaStream_t s);
int qbx_launch_boys(int64_t n, const double *dT, int mmax, int table, double *d_out, cudaStream_t s);
