// cuBLAS / cuSOLVER bound at run time (dlopen; the copy a host framework has already loaded is re-used), for the two
// places SURVEY.md 8(f) allows a library: the symmetric eigenproblem of an SCF step (`eigen`, HartreeFock.jl:46-56) and
// the plain FP64 GEMMs of the density / MO transforms.  libqbx.so has no link-time dependency on either library and a
// process that only builds integrals never loads them.
#include <dlfcn.h>

#include <mutex>

#include "linalg.h"

namespace {
std::mutex g_lmu;
struct Lib {
    void *blas = nullptr, *solver = nullptr;
    void *hb = nullptr, *hs = nullptr;            // cublasHandle_t, cusolverDnHandle_t
    int (*blasCreate)(void **) = nullptr;
    int (*blasSetStream)(void *, cudaStream_t) = nullptr;
    int (*dgemm)(void *, int, int, int, int, int, const double *, const double *, int, const double *, int, const double *, double *, int) = nullptr;
    int (*dgemmSB)(void *, int, int, int, int, int, const double *, const double *, int, long long, const double *, int, long long,
                   const double *, double *, int, long long, int) = nullptr;
    int (*solverCreate)(void **) = nullptr;
    int (*solverSetStream)(void *, cudaStream_t) = nullptr;
    int (*syevdBuf)(void *, int, int, int, const double *, int, const double *, int *) = nullptr;
    int (*syevd)(void *, int, int, int, double *, int, double *, double *, int, int *) = nullptr;
} L;

void *open_first(const char *const *names)
{
    for (int pass = 0; pass < 2; ++pass)
        for (const char *const *n = names; *n; ++n)
            if (void *h = dlopen(*n, RTLD_NOW | (pass == 0 ? RTLD_NOLOAD : RTLD_GLOBAL))) return h;
    return nullptr;
}

int bind()
{
    if (L.hb && L.hs) return QBX_OK;
    static const char *const bn[] = {"libcublas.so.12", "libcublas.so.13", "libcublas.so", nullptr};
    static const char *const sn[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so", nullptr};
    L.blas = open_first(bn);
    L.solver = open_first(sn);
    if (!L.blas || !L.solver) { qbx_set_error(std::string("linalg: cannot load cuBLAS / cuSOLVER: ") + (dlerror() ? dlerror() : "")); return QBX_ERR_STATE; }
#define QBX_SYM(field, lib, name) L.field = (decltype(L.field))dlsym(lib, name); if (!L.field) { qbx_set_error("linalg: missing " name); return QBX_ERR_STATE; }
    QBX_SYM(blasCreate, L.blas, "cublasCreate_v2")
    QBX_SYM(blasSetStream, L.blas, "cublasSetStream_v2")
    QBX_SYM(dgemm, L.blas, "cublasDgemm_v2")
    QBX_SYM(dgemmSB, L.blas, "cublasDgemmStridedBatched")
    QBX_SYM(solverCreate, L.solver, "cusolverDnCreate")
    QBX_SYM(solverSetStream, L.solver, "cusolverDnSetStream")
    QBX_SYM(syevdBuf, L.solver, "cusolverDnDsyevd_bufferSize")
    QBX_SYM(syevd, L.solver, "cusolverDnDsyevd")
#undef QBX_SYM
    if (L.blasCreate(&L.hb) != 0) { L.hb = nullptr; qbx_set_error("linalg: cublasCreate failed"); return QBX_ERR_CUDA; }
    if (L.solverCreate(&L.hs) != 0) { L.hs = nullptr; qbx_set_error("linalg: cusolverDnCreate failed"); return QBX_ERR_CUDA; }
    return QBX_OK;
}
}   // namespace

// C (m x n) = alpha op(A) op(B) + beta C, column-major, ta / tb: 0 = N, 1 = T
int qbx_gemm(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B, int ldb, double beta,
             double *C, int ldc, cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_lmu);
    int rc = bind();
    if (rc) return rc;
    L.blasSetStream(L.hb, s);
    if (L.dgemm(L.hb, ta, tb, m, n, k, &alpha, A, lda, B, ldb, &beta, C, ldc) != 0) { qbx_set_error("cublasDgemm failed"); return QBX_ERR_CUDA; }
    return QBX_OK;
}

int qbx_gemm_batched(int ta, int tb, int m, int n, int k, double alpha, const double *A, int lda, long long sa, const double *B, int ldb,
                     long long sb, double beta, double *C, int ldc, long long sc, int batch, cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_lmu);
    int rc = bind();
    if (rc) return rc;
    L.blasSetStream(L.hb, s);
    if (L.dgemmSB(L.hb, ta, tb, m, n, k, &alpha, A, lda, sa, B, ldb, sb, &beta, C, ldc, sc, batch) != 0) {
        qbx_set_error("cublasDgemmStridedBatched failed");
        return QBX_ERR_CUDA;
    }
    return QBX_OK;
}

// eigen-decomposition of the symmetric n x n matrix in A (column-major, overwritten by the eigenvectors), ascending w
int qbx_syevd(int n, double *A, double *w, cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_lmu);
    int rc = bind();
    if (rc) return rc;
    L.solverSetStream(L.hs, s);
    int lwork = 0;
    if (L.syevdBuf(L.hs, 1 /*vectors*/, 0 /*lower*/, n, A, n, w, &lwork) != 0) { qbx_set_error("cusolverDnDsyevd_bufferSize failed"); return QBX_ERR_CUDA; }
    double *work = nullptr; int *info = nullptr;
    QBX_CUDA(qbx_dmalloc(&work, (size_t)std::max(lwork, 1) * sizeof(double)));
    QBX_CUDA(qbx_dmalloc(&info, sizeof(int)));
    const int st = L.syevd(L.hs, 1, 0, n, A, n, w, work, lwork, info);
    qbx_pool_free_async(work); qbx_pool_free_async(info);
    if (st != 0) { qbx_set_error("cusolverDnDsyevd failed"); return QBX_ERR_CUDA; }
    return QBX_OK;
}
