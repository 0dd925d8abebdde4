// Device-memory pool of libqbx.so.
//
// Every step of the reference's workflow (one geometry of a scan, one SCF of an optimisation,
// src/HartreeFock.jl:583-606) builds a new basis, a new ERI store of the same few sizes, and
// drops them again.  cudaMalloc/cudaFree of the ~10 GB packed store cost 0.1-0.7 s per step on
// B200 (page mapping + the implicit device synchronisation of cudaFree) -- several times the
// 32 ms the ERI kernels need -- so freed blocks are kept here, keyed by size, and handed out again.
//   QBX_POOL_GB=<n>  upper bound of cached (not in use) bytes, default 64; 0 switches pooling off.
// qbx_pool_trim() / qbx_shutdown() give everything back to the driver.
#include <map>
#include <mutex>
#include <unordered_map>

#include "qbx_internal.h"

namespace {
struct Pool {
    std::mutex mu;
    std::multimap<size_t, void *> idle;              // size -> block
    std::unordered_map<void *, size_t> size_of;      // every block that came from qbx_pool_malloc
    size_t idle_bytes = 0, cap = 0;
    bool cap_read = false;
    int64_t hits = 0, misses = 0;
};
Pool g_pool;

size_t pool_cap(Pool &P)
{
    if (!P.cap_read) {
        const char *e = getenv("QBX_POOL_GB");
        const double gb = e ? atof(e) : 64.0;
        P.cap = gb <= 0 ? 0 : (size_t)(gb * (double)(1ull << 30));
        P.cap_read = true;
    }
    return P.cap;
}

void release_idle(Pool &P)
{
    for (auto &kv : P.idle) { P.size_of.erase(kv.second); cudaFree(kv.second); }
    P.idle.clear();
    P.idle_bytes = 0;
}
}   // namespace

cudaError_t qbx_pool_malloc(void **p, size_t bytes)
{
    Pool &P = g_pool;
    const size_t sz = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(P.mu);
    if (pool_cap(P)) {
        auto it = P.idle.lower_bound(sz);
        // a block is reused when it wastes at most a quarter of itself (or less than 1 MiB)
        if (it != P.idle.end() && (it->first - sz <= (1u << 20) || it->first - sz <= it->first / 4)) {
            *p = it->second;
            P.idle_bytes -= it->first;
            P.idle.erase(it);
            ++P.hits;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, sz);
    if (e != cudaSuccess && !P.idle.empty()) {       // give the cache back and try again
        cudaGetLastError();
        release_idle(P);
        e = cudaMalloc(p, sz);
    }
    if (e == cudaSuccess) { P.size_of[*p] = sz; ++P.misses; }
    return e;
}

static thread_local int tl_free_scope = 0;
QbxPoolFreeScope::QbxPoolFreeScope() { if (tl_free_scope++ == 0) cudaDeviceSynchronize(); }
QbxPoolFreeScope::~QbxPoolFreeScope() { --tl_free_scope; }

static cudaError_t pool_free(void *p, bool sync)
{
    if (!p) return cudaSuccess;
    if (tl_free_scope > 0) sync = false;
    Pool &P = g_pool;
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.size_of.find(p);
    if (it == P.size_of.end()) return cudaFree(p);    // not ours
    const size_t sz = it->second;
    if (pool_cap(P) == 0 || P.idle_bytes + sz > P.cap) {
        P.size_of.erase(it);
        return cudaFree(p);
    }
    cudaError_t e = sync ? cudaDeviceSynchronize() : cudaSuccess;
    P.idle.emplace(sz, p);
    P.idle_bytes += sz;
    return e;
}

// same contract as cudaFree: work that still uses the block has finished when this returns
cudaError_t qbx_pool_free(void *p) { return pool_free(p, true); }

// Stream-ordered variant for scratch blocks: the caller guarantees that every later user of the
// block is ordered after the work enqueued so far.  That holds inside the library because all of
// it runs on the library's one stream (side streams fork from and join into it) and
// qbx_set_stream synchronises the device when the stream changes.
static thread_local std::vector<void *> *tl_deferred = nullptr;
cudaError_t qbx_pool_free_async(void *p)
{
    if (p && tl_deferred) { tl_deferred->push_back(p); return cudaSuccess; }
    return pool_free(p, false);
}

QbxPoolDeferScope::QbxPoolDeferScope() { if (!tl_deferred) tl_deferred = new std::vector<void *>(); }
QbxPoolDeferScope::~QbxPoolDeferScope() { release(); }
void QbxPoolDeferScope::release()
{
    if (!tl_deferred) return;
    std::vector<void *> *list = tl_deferred;
    tl_deferred = nullptr;
    for (void *p : *list) pool_free(p, false);
    delete list;
}

// Pinned host scratch for the small read-backs (counts, totals); grows, never shrinks.
void *qbx_pinned(size_t bytes)
{
    static void *buf = nullptr;
    static size_t cap = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (bytes > cap) {
        if (buf) cudaFreeHost(buf);
        cap = std::max<size_t>(bytes, 1 << 16);
        if (cudaMallocHost(&buf, cap) != cudaSuccess) { cudaGetLastError(); buf = nullptr; cap = 0; }
    }
    return buf;
}

void qbx_pool_release()
{
    std::lock_guard<std::mutex> lk(g_pool.mu);
    release_idle(g_pool);
}

void qbx_pool_counts(int64_t *hits, int64_t *misses, int64_t *idle_bytes)
{
    std::lock_guard<std::mutex> lk(g_pool.mu);
    *hits = g_pool.hits; *misses = g_pool.misses; *idle_bytes = (int64_t)g_pool.idle_bytes;
}

// Pinned staging area for the host-built tables of a basis (primitive-pair records): resident pages
// (no first-touch faults on every qbx_basis_create) and DMA-able without a bounce buffer.
// Grow-only; the caller holds qbx_staging_mutex() while it uses the pointer.
std::mutex &qbx_staging_mutex()
{
    static std::mutex mu;
    return mu;
}

void *qbx_staging(size_t bytes)
{
    static void *buf = nullptr;
    static size_t cap = 0;
    if (bytes > cap) {
        if (buf) cudaFreeHost(buf);
        cap = bytes + bytes / 2;
        if (cudaMallocHost(&buf, cap) != cudaSuccess) { cudaGetLastError(); buf = nullptr; cap = 0; }
    }
    return buf;
}

// ------------------------------------------------------------------ persistent host workers
// A fixed set of threads that sleep on a condition variable between calls; one job at a time
// (qbx_parallel_for is only entered under the library's handle mutex or the staging mutex, a second
// caller simply waits).  The threads are detached and live until the process ends.
#include <condition_variable>
namespace {
struct HostWorkers {
    std::mutex job_mu;                        // one job at a time
    std::mutex mu;
    std::condition_variable cv_start, cv_done;
    void (*fn)(void *) = nullptr;
    void *ctx = nullptr;
    unsigned wanted = 0, started = 0, running = 0;
    uint64_t epoch = 0;
    unsigned nthreads = 0;
    HostWorkers()
    {
        nthreads = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        for (unsigned t = 1; t < nthreads; ++t) std::thread([this] { loop(); }).detach();
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            void (*f)(void *);
            void *c;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_start.wait(lk, [&] { return epoch != seen && started < wanted; });
                seen = epoch;
                ++started;
                f = fn; c = ctx;
            }
            f(c);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
    void run(unsigned nt, void (*f)(void *), void *c)
    {
        std::lock_guard<std::mutex> job(job_mu);
        nt = std::min(nt, nthreads);
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = f; ctx = c;
            wanted = nt - 1; started = 0; running = nt - 1;
            ++epoch;
        }
        if (nt > 1) cv_start.notify_all();
        f(c);                                 // the caller works too
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return running == 0; });
        wanted = 0;
    }
};
HostWorkers &host_workers()
{
    static HostWorkers *w = new HostWorkers;  // never destroyed: detached threads may outlive static destructors
    return *w;
}
}   // namespace

unsigned qbx_host_workers() { return host_workers().nthreads; }
void qbx_host_workers_run(unsigned nthreads, void (*fn)(void *), void *ctx) { host_workers().run(nthreads, fn, ctx); }
