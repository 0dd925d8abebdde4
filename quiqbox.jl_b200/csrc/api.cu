// C ABI (include/qbx.h) and host-side orchestration of libqbx.so.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>

#include "engine.h"
#include "qbx_internal.h"

// ------------------------------------------------------------------ error + globals
static thread_local std::string g_err;
void qbx_set_error(const std::string &msg) { g_err = msg; }
extern "C" const char *qbx_last_error(void) { return g_err.c_str(); }

static std::mutex g_mu;
static int g_device = -1;
static cudaStream_t g_stream = nullptr;
static double *g_boys_f = nullptr, *g_boys_e = nullptr;

static cudaStream_t g_own_stream = nullptr;
cudaStream_t qbx_stream() { return g_stream; }
BoysTable qbx_boys_table() { return BoysTable{g_boys_f, g_boys_e}; }

// F_m(T) in long double: convergent series at the top order, downward recursion
static void host_boys_row(long double T, int mtop, long double *F)
{
    long double e = expl(-T), term = 1.0L / (2 * mtop + 1), sum = term;
    for (int k = 1; k < 100000; ++k) {
        term *= 2.0L * T / (2 * mtop + 2 * k + 1);
        sum += term;
        if (term < 1e-22L * sum) break;
    }
    F[mtop] = e * sum;
    for (int m = mtop; m >= 1; --m) F[m - 1] = (2.0L * T * F[m] + e) / (2 * m - 1);
}

static int build_boys_table()
{
    std::vector<double> f((size_t)QBX_BOYS_NROW * QBX_BOYS_NCOL), e(QBX_BOYS_NROW);
    long double row[QBX_BOYS_NCOL + 8];
    for (int i = 0; i < QBX_BOYS_NROW; ++i) {
        long double T = (long double)i / (long double)QBX_BOYS_STEP_INV;
        host_boys_row(T, QBX_BOYS_NCOL + 7, row);
        for (int m = 0; m < QBX_BOYS_NCOL; ++m) f[(size_t)i * QBX_BOYS_NCOL + m] = (double)row[m];
        e[i] = (double)expl(-T);
    }
    QBX_CUDA(qbx_dmalloc(&g_boys_f, f.size() * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&g_boys_e, e.size() * sizeof(double)));
    QBX_CUDA(cudaMemcpy(g_boys_f, f.data(), f.size() * sizeof(double), cudaMemcpyHostToDevice));
    QBX_CUDA(cudaMemcpy(g_boys_e, e.data(), e.size() * sizeof(double), cudaMemcpyHostToDevice));
    return QBX_OK;
}

extern "C" int qbx_init(int device, int *n_dev_out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    QBX_CUDA(cudaGetDeviceCount(&n));
    if (n_dev_out) *n_dev_out = n;
    if (n == 0) { qbx_set_error("qbx_init: no CUDA device visible; this library has no CPU fallback"); return QBX_ERR_CUDA; }
    if (device < 0 || device >= n) { qbx_set_error("qbx_init: device index out of range"); return QBX_ERR_ARG; }
    if (g_device == device) return QBX_OK;
    if (g_device >= 0) { qbx_set_error("qbx_init: this process is already bound to another device"); return QBX_ERR_STATE; }
    QBX_CUDA(cudaSetDevice(device));
    QBX_CUDA(cudaStreamCreateWithFlags(&g_own_stream, cudaStreamNonBlocking));
    g_stream = g_own_stream;
    int rc = build_boys_table();
    if (rc) return rc;
    g_device = device;
    return QBX_OK;
}

extern "C" int qbx_pool_trim(int64_t *counts)
{
    if (counts) qbx_pool_counts(counts, counts + 1, counts + 2);
    qbx_pool_release();
    return QBX_OK;
}

extern "C" int qbx_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) return QBX_OK;
    cudaSetDevice(g_device);
    qbx_pool_free(g_boys_f); qbx_pool_free(g_boys_e);
    g_boys_f = g_boys_e = nullptr;
    qbx_pool_release();
    cudaStreamDestroy(g_own_stream);
    g_stream = g_own_stream = nullptr;
    g_device = -1;
    return QBX_OK;
}

int qbx_ensure_init()
{
    if (g_device < 0) {
        int rc = qbx_init(0, nullptr);
        if (rc) return rc;
    }
    QBX_CUDA(cudaSetDevice(g_device));     // Julia tasks migrate between OS threads
    return QBX_OK;
}

#include "handle.h"

template <class T>
static int to_device(T **dst, const std::vector<T> &src)
{
    QBX_CUDA(qbx_dmalloc(dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
    // enqueued: `src` is a member of the handle and outlives the copy; every later use is on the same stream
    if (!src.empty()) QBX_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, g_stream));
    return QBX_OK;
}

extern "C" int qbx_basis_create(int64_t nprim, const double *cen, const double *xpn, const int32_t *ang, int64_t nbf,
                                const int64_t *bf_off, const int64_t *bf_prim, const double *bf_w, qbx_basis **out)
{
    if (!out || nprim <= 0 || nbf <= 0 || !cen || !xpn || !ang || !bf_off || !bf_prim || !bf_w) {
        qbx_set_error("qbx_basis_create: null or empty argument");
        return QBX_ERR_ARG;
    }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    const int64_t nnz = bf_off[nbf];
    if (bf_off[0] != 0 || nnz <= 0) { qbx_set_error("qbx_basis_create: bf_off must start at 0 and be non-empty"); return QBX_ERR_ARG; }
    for (int64_t i = 0; i < nbf; ++i)
        if (bf_off[i + 1] <= bf_off[i]) { qbx_set_error("qbx_basis_create: every basis function needs >= 1 primitive"); return QBX_ERR_ARG; }
    for (int64_t k = 0; k < nnz; ++k)
        if (bf_prim[k] < 0 || bf_prim[k] >= nprim) { qbx_set_error("qbx_basis_create: primitive index out of range"); return QBX_ERR_RANGE; }
    for (int64_t p = 0; p < nprim; ++p) {
        if (!(xpn[p] > 0.0)) { qbx_set_error("qbx_basis_create: exponents must be positive"); return QBX_ERR_ARG; }
        for (int d = 0; d < 3; ++d)
            if (ang[3 * p + d] < 0) { qbx_set_error("qbx_basis_create: negative angular momentum"); return QBX_ERR_ARG; }
    }
    std::unique_ptr<qbx_basis> b(new qbx_basis);
    b->nprim = nprim; b->nbf = nbf;
    b->cen.assign(cen, cen + 3 * nprim);
    b->xpn.assign(xpn, xpn + nprim);
    b->ang.assign(ang, ang + 3 * nprim);
    b->bf_off.assign(bf_off, bf_off + nbf + 1);
    b->bf_prim.assign(bf_prim, bf_prim + nnz);
    b->bf_w.assign(bf_w, bf_w + nnz);
    b->flat.nprim = nprim; b->flat.nbf = nbf; b->flat.nnz = nnz;
    if ((rc = to_device(&b->flat.cen, b->cen))) return rc;
    if ((rc = to_device(&b->flat.xpn, b->xpn))) return rc;
    if ((rc = to_device(&b->flat.ang, b->ang))) return rc;
    if ((rc = to_device(&b->flat.bf_off, b->bf_off))) return rc;
    if ((rc = to_device(&b->flat.bf_prim, b->bf_prim))) return rc;
    if ((rc = to_device(&b->flat.bf_w, b->bf_w))) return rc;
    b->eng.reset(Engine::create(nprim, cen, xpn, ang, nbf, bf_off, bf_prim, bf_w));   // null if irregular
    *out = b.release();
    return QBX_OK;
}

static void free_store(qbx_basis *b)
{
    qbx_pool_free(b->d_dense); b->d_dense = nullptr;
    if (b->eng) b->eng->release_store();
    b->mode = -1;
}

extern "C" int qbx_basis_destroy(qbx_basis *b)
{
    if (!b) return QBX_OK;
    if (g_device >= 0) cudaSetDevice(g_device);
    QbxPoolFreeScope one_sync;
    free_store(b);
    qbx_pool_free(b->flat.cen); qbx_pool_free(b->flat.xpn); qbx_pool_free(b->flat.ang);
    qbx_pool_free(b->flat.bf_off); qbx_pool_free(b->flat.bf_prim); qbx_pool_free(b->flat.bf_w);
    qbx_pool_free(b->d_DJ); qbx_pool_free(b->d_DK); qbx_pool_free(b->d_G);
    delete b;
    return QBX_OK;
}

extern "C" int qbx_basis_info(qbx_basis *b, int64_t *info)
{
    if (!b || !info) { qbx_set_error("qbx_basis_info: null argument"); return QBX_ERR_ARG; }
    for (int i = 0; i < 16; ++i) info[i] = 0;
    info[0] = b->nbf;
    if (b->eng) b->eng->info(info);
    return QBX_OK;
}

extern "C" int qbx_eri_quartets(qbx_basis *b, int64_t n, const int64_t *ijkl, double *out)
{
    if (!b || n < 0 || (n > 0 && (!ijkl || !out))) { qbx_set_error("qbx_eri_quartets: bad argument"); return QBX_ERR_ARG; }
    if (n == 0) return QBX_OK;
    for (int64_t t = 0; t < 4 * n; ++t)
        if (ijkl[t] < 0 || ijkl[t] >= b->nbf) { qbx_set_error("qbx_eri_quartets: function index out of range"); return QBX_ERR_RANGE; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    QbxScratch<int64_t> d_idx;                               // (freed on every return path)
    QbxScratch<double> d_out;
    QBX_CUDA(d_idx.alloc(4 * n * sizeof(int64_t)));
    QBX_CUDA(d_out.alloc(n * sizeof(double)));
    QBX_CUDA(cudaMemcpyAsync(d_idx, ijkl, 4 * n * sizeof(int64_t), cudaMemcpyHostToDevice, g_stream));
    rc = qbx_launch_generic_quartets(b->flat, n, d_idx, d_out, g_stream);
    if (!rc) {
        b->stats[0] += 1;
        QBX_CUDA(cudaMemcpyAsync(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        QBX_CUDA(cudaStreamSynchronize(g_stream));
    }
    if (!rc)
        for (int64_t t = 0; t < n; ++t)
            if (out[t] != out[t]) { qbx_set_error("qbx_eri_quartets: angular momentum beyond the generic kernel's range"); return QBX_ERR_RANGE; }
    return rc;
}

extern "C" int qbx_eri_tensor(qbx_basis *b, double *out, int64_t out_bytes)
{
    if (!b || !out) { qbx_set_error("qbx_eri_tensor: null argument"); return QBX_ERR_ARG; }
    const int64_t N = b->nbf, need = N * N * N * N * (int64_t)sizeof(double);
    if (out_bytes < need) { qbx_set_error("qbx_eri_tensor: output buffer smaller than nbf^4 * 8 bytes"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    QbxScratch<double> d_t;
    QBX_CUDA(d_t.alloc(need));
    if (b->eng) rc = b->eng->fill_tensor(d_t, g_stream, b->stats);
    else { rc = qbx_launch_generic_tensor(b->flat, d_t, g_stream); b->stats[0] += 1; }
    if (!rc) {
        QBX_CUDA(cudaMemcpyAsync(out, d_t, need, cudaMemcpyDeviceToHost, g_stream));
        QBX_CUDA(cudaStreamSynchronize(g_stream));
    }
    if (!rc && !b->eng)
        for (int64_t t = 0; t < N * N * N * N; ++t)
            if (out[t] != out[t]) { qbx_set_error("qbx_eri_tensor: angular momentum beyond the generic kernel's range"); return QBX_ERR_RANGE; }
    return rc;
}

extern "C" int qbx_eri_store(qbx_basis *b, double screen_tol, int mode, int rank, int nranks)
{
    if (!b || mode < 0 || mode > 2 || nranks < 1 || rank < 0 || rank >= nranks || !(screen_tol >= 0.0)) {
        qbx_set_error("qbx_eri_store: bad argument");
        return QBX_ERR_ARG;
    }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    if (nranks > 1 && qbx_comm_size() == nranks && qbx_comm_rank() != rank) {
        qbx_set_error("qbx_eri_store: rank differs from the rank of this process in the communicator (qbx_comm_init)");
        return QBX_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(b->mu);
    free_store(b);
    b->nranks = nranks;
    if (mode == 2 || !b->eng) {
        const int64_t N = b->nbf, need = N * N * N * N * (int64_t)sizeof(double);
        if (!b->eng && mode != 2) {
            // A basis with functions outside the s/p/d shell classes (l > 2, contractions over two centres) has no shell
            // quartet lists: the packed store and the direct mode do not exist for it.  Small bases are served by the dense
            // tensor (same results, documented in qbx.h); a large one gets an error that says so instead of an N^4 allocation
            // that cannot succeed or a shard request that is silently ignored.
            if (nranks != 1 || need > ((int64_t)16 << 30)) {
                qbx_set_error("qbx_eri_store: this basis contains functions outside the s/p/d shell classes (l > 2 or a contraction over "
                              "several centres), so the packed store / direct mode and the sharding over ranks are not available; the dense "
                              "fallback needs nbf^4 * 8 = " + std::to_string((long long)(need >> 20)) + " MiB on one GPU (limit 16 GiB). "
                              "Use mode 2 explicitly on one rank, or remove the irregular functions.");
                return QBX_ERR_STATE;
            }
        }
        if (nranks != 1) { qbx_set_error("qbx_eri_store: the dense mode does not shard"); return QBX_ERR_ARG; }
        QBX_CUDA(qbx_dmalloc(&b->d_dense, need));
        if (b->eng) rc = b->eng->fill_tensor(b->d_dense, g_stream, b->stats);
        else { rc = qbx_launch_generic_tensor(b->flat, b->d_dense, g_stream); b->stats[0] += 1; }
        if (rc) return rc;
        QBX_CUDA(cudaStreamSynchronize(g_stream));
        b->mode = 2;
        return QBX_OK;
    }
    rc = b->eng->store(screen_tol, mode, rank, nranks, g_stream, b->stats);
    if (rc) return rc;
    b->mode = mode;
    return QBX_OK;
}

extern "C" int qbx_eri_recompute_async(qbx_basis *b)
{
    if (!b) { qbx_set_error("qbx_eri_recompute: null handle"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->mode != 0 || !b->eng) { qbx_set_error("qbx_eri_recompute: call qbx_eri_store(mode = 0) first"); return QBX_ERR_STATE; }
    return b->eng->recompute(g_stream, b->stats);
}

extern "C" int qbx_eri_recompute(qbx_basis *b)
{
    int rc = qbx_eri_recompute_async(b);
    if (rc) return rc;
    QBX_CUDA(cudaStreamSynchronize(g_stream));
    return QBX_OK;
}

extern "C" int qbx_set_stream(void *stream)
{
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_mu);
    QBX_CUDA(cudaStreamSynchronize(g_stream));
    g_stream = stream ? (cudaStream_t)stream : g_own_stream;
    return QBX_OK;
}

extern "C" int qbx_class_stats(qbx_basis *b, double *out)
{
    if (!b || !out) { qbx_set_error("qbx_class_stats: null argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    if (!b->eng) { qbx_set_error("qbx_class_stats: basis has no shell-class path"); return QBX_ERR_STATE; }
    return b->eng->class_stats(g_stream, b->stats, out);
}

// register-resident DFMA chains: 8 independent accumulators per thread
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int qbx_fp64_peak(double *tflops)
{
    if (!tflops) { qbx_set_error("qbx_fp64_peak: null argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    cudaDeviceProp prop;
    QBX_CUDA(cudaGetDeviceProperties(&prop, g_device));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
    double *d = nullptr;
    QBX_CUDA(qbx_dmalloc(&d, (size_t)blocks * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    QBX_CUDA(cudaEventCreate(&e0));
    QBX_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        QBX_CUDA(cudaEventRecord(e0, g_stream));
        k_dfma_peak<<<blocks, 256, 0, g_stream>>>(d, iters, 0.999999, 1e-9);
        QBX_CUDA(cudaEventRecord(e1, g_stream));
        QBX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        QBX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); qbx_pool_free(d);
    *tflops = best;
    return QBX_OK;
}

int qbx_fock_device(qbx_basis *b, int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s)
{
    if (b->mode < 0) { qbx_set_error("qbx_fock_build: call qbx_eri_store first"); return QBX_ERR_STATE; }
    if (b->mode == 2) {
        b->stats[0] += 1;
        return qbx_launch_dense_gcore(b->nbf, b->d_dense, nmat, dDJ, dDK, dG, s);
    }
    int rc = b->eng->fock(nmat, dDJ, dDK, dG, s, b->stats);
    // the collective inside the boundary: with a communicator of the store's size the partial G of the ranks is summed
    // here (one ncclAllReduce on the same stream), and getGcore's contract -- the result is the full G -- holds
    if (!rc && b->nranks > 1 && qbx_comm_size() == b->nranks) rc = qbx_comm_allreduce(dG, (size_t)nmat * b->nbf * b->nbf, s);
    return rc;
}

extern "C" int qbx_fock_build_device(qbx_basis *b, int nmat, const double *dDJ, const double *dDK, double *dG, void *stream)
{
    if (!b || nmat < 1 || nmat > 2 || !dDJ || !dDK || !dG) { qbx_set_error("qbx_fock_build_device: bad argument"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    return qbx_fock_device(b, nmat, dDJ, dDK, dG, stream ? (cudaStream_t)stream : g_stream);
}

extern "C" int qbx_fock_build(qbx_basis *b, int nmat, const double *DJ, const double *DK, double *G)
{
    if (!b || nmat < 1 || nmat > 2 || !DJ || !DK || !G) { qbx_set_error("qbx_fock_build: bad argument (nmat must be 1 or 2)"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    const size_t n2 = (size_t)b->nbf * b->nbf * sizeof(double);
    if (b->staged_nmat < nmat) {
        qbx_pool_free(b->d_DJ); qbx_pool_free(b->d_DK); qbx_pool_free(b->d_G);
        b->d_DJ = b->d_DK = b->d_G = nullptr; b->staged_nmat = 0;      // (an early return below must not leave dangling pointers)
        QBX_CUDA(qbx_dmalloc(&b->d_DJ, n2));
        QBX_CUDA(qbx_dmalloc(&b->d_DK, n2 * 2));
        QBX_CUDA(qbx_dmalloc(&b->d_G, n2 * 2));
        b->staged_nmat = 2;
    }
    // The caller's matrices are pageable: they go through the library's pinned staging area (one memcpy each, then a DMA
    // that runs while the host checks the symmetry), and G comes back the same way -- a pageable cudaMemcpyAsync of
    // these 1.3 MB matrices costs 0.2-0.3 ms each and blocks the host.
    std::lock_guard<std::mutex> stage_lock(qbx_staging_mutex());
    char *stage = (char *)qbx_staging(n2 * (1 + 2 * (size_t)nmat));
    if (!stage) { qbx_set_error("qbx_fock_build: pinned staging allocation failed"); return QBX_ERR_NOMEM; }
    double *hDJ = (double *)stage, *hDK = (double *)(stage + n2), *hG = (double *)(stage + n2 * (1 + (size_t)nmat));
    memcpy(hDJ, DJ, n2);
    QBX_CUDA(cudaMemcpyAsync(b->d_DJ, hDJ, n2, cudaMemcpyHostToDevice, g_stream));
    memcpy(hDK, DK, n2 * nmat);
    QBX_CUDA(cudaMemcpyAsync(b->d_DK, hDK, n2 * nmat, cudaMemcpyHostToDevice, g_stream));
    if (b->mode == 0 || b->mode == 1) {
        // the packed-store digestion is only valid for symmetric densities (include/qbx.h); reject anything else here
        // instead of returning a mode-dependent result
        const int64_t N = b->nbf, T = 32;                        // tiles: both D[i,j] and D[j,i] stay in cache
        for (int m = 0; m <= nmat; ++m) {
            const double *D = m == 0 ? DJ : DK + (size_t)(m - 1) * N * N;
            bool ok = true;
            for (int64_t j0 = 0; j0 < N; j0 += T)
                for (int64_t i0 = 0; i0 <= j0; i0 += T)
                    for (int64_t j = j0; j < std::min(N, j0 + T); ++j)
                        for (int64_t i = i0; i < std::min(j, i0 + T); ++i) {
                            const double v = D[i + N * j], w = D[j + N * i];
                            ok &= fabs(v - w) <= 1e-10 * std::max(1.0, std::max(fabs(v), fabs(w)));   // (false for NaN)
                        }
            if (!ok) {
                cudaStreamSynchronize(g_stream);                 // the staging area is in flight
                qbx_set_error("qbx_fock_build: DJ and DK must be symmetric matrices (stored and direct modes)");
                return QBX_ERR_ARG;
            }
        }
    }
    rc = qbx_fock_device(b, nmat, b->d_DJ, b->d_DK, b->d_G, g_stream);
    if (rc) { cudaStreamSynchronize(g_stream); return rc; }
    QBX_CUDA(cudaMemcpyAsync(hG, b->d_G, n2 * nmat, cudaMemcpyDeviceToHost, g_stream));
    QBX_CUDA(cudaStreamSynchronize(g_stream));
    memcpy(G, hG, n2 * nmat);
    return QBX_OK;
}

extern "C" int qbx_one_body(qbx_basis *b, int kind, int64_t nnuc, const double *Z, const double *R, double *out)
{
    if (!b || kind < 0 || kind > 2 || !out || (kind == 2 && nnuc > 0 && (!Z || !R)) || nnuc < 0) {
        qbx_set_error("qbx_one_body: bad argument");
        return QBX_ERR_ARG;
    }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(b->mu);
    const size_t n2 = (size_t)b->nbf * b->nbf * sizeof(double);
    QbxScratch<double> dZ, dR, dO;                           // (freed on every return path)
    QBX_CUDA(dO.alloc(n2));
    if (kind == 2 && nnuc > 0) {
        QBX_CUDA(dZ.alloc(nnuc * sizeof(double)));
        QBX_CUDA(dR.alloc(3 * nnuc * sizeof(double)));
        QBX_CUDA(cudaMemcpyAsync(dZ, Z, nnuc * sizeof(double), cudaMemcpyHostToDevice, g_stream));
        QBX_CUDA(cudaMemcpyAsync(dR, R, 3 * nnuc * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    }
    rc = qbx_launch_one_body(b->flat, kind, nnuc, dZ, dR, dO, g_stream);
    b->stats[0] += 1;
    if (!rc) {
        QBX_CUDA(cudaMemcpyAsync(out, dO, n2, cudaMemcpyDeviceToHost, g_stream));
        QBX_CUDA(cudaStreamSynchronize(g_stream));
    }
    if (!rc)
        for (int64_t t = 0; t < b->nbf * b->nbf; ++t)
            if (out[t] != out[t]) { qbx_set_error("qbx_one_body: angular momentum beyond the generic kernel's range"); return QBX_ERR_RANGE; }
    return rc;
}

extern "C" int qbx_boys(int64_t n, const double *T, int mmax, int table, double *out)
{
    if (n < 0 || mmax < 0 || mmax > 128 || (table && mmax > 8) || (n > 0 && (!T || !out))) {
        qbx_set_error("qbx_boys: bad argument (mmax <= 128; <= 8 for the tabulated path)");
        return QBX_ERR_ARG;
    }
    if (n == 0) return QBX_OK;
    for (int64_t i = 0; i < n; ++i)
        if (!(T[i] >= 0.0)) { qbx_set_error("qbx_boys: T must be >= 0"); return QBX_ERR_ARG; }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    QbxScratch<double> dT, dO;
    QBX_CUDA(dT.alloc(n * sizeof(double)));
    QBX_CUDA(dO.alloc(n * (mmax + 1) * sizeof(double)));
    QBX_CUDA(cudaMemcpyAsync(dT, T, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    rc = qbx_launch_boys(n, dT, mmax, table, dO, g_stream);
    if (!rc) {
        QBX_CUDA(cudaMemcpyAsync(out, dO, n * (mmax + 1) * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
        QBX_CUDA(cudaStreamSynchronize(g_stream));
    }
    return rc;
}

extern "C" int qbx_prim_batch(int la, int lb, int lc, int ld, int K, int64_t nquartets, uint64_t seed, double *secs,
                              double *checksum, double *prim_quartets, int64_t nsample, double *sample_out,
                              double *sample_geom)
{
    if (la < 0 || la > QBX_MAX_L || lb < 0 || lb > la || lc < 0 || lc > QBX_MAX_L || ld < 0 || ld > lc || K < 1 ||
        K > 16 || nquartets <= 0 || !secs || !checksum) {
        qbx_set_error("qbx_prim_batch: bad argument (need la >= lb, lc >= ld, l <= 2, 1 <= K <= 16)");
        return QBX_ERR_ARG;
    }
    int rc = qbx_ensure_init();
    if (rc) return rc;
    return Engine::synthetic(la, lb, lc, ld, K, nquartets, seed, secs, checksum, prim_quartets, nsample, sample_out, sample_geom,
                             g_stream);
}

extern "C" int qbx_stats(qbx_basis *b, double *out, int reset)
{
    if (!b) { qbx_set_error("qbx_stats: null handle"); return QBX_ERR_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (out) memcpy(out, b->stats, sizeof(b->stats));
    if (reset) memset(b->stats, 0, sizeof(b->stats));
    return QBX_OK;
}
